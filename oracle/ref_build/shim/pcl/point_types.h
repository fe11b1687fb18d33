#include "../ref_shim_ros.hpp"
