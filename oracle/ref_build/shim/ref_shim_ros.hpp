// =================================================================================================
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Inert stand-ins for the ROS / PCL / OpenCV types the
// UNMODIFIED reference sources mention (include/mlmap.h:7-35, include/rviz_vis.h, src/mlmap.cpp,
// src/rviz_vis.cpp).  None of them computes anything on the mapping path: messages are plain structs,
// publishers keep the last message of every topic so the driver (ref_capi.cpp) can read the clouds the
// reference publishes, timers and subscribers never fire.  ROS/nodelet plumbing is out of scope (north star).
// =================================================================================================
#ifndef MLM_REF_SHIM_ROS_HPP
#define MLM_REF_SHIM_ROS_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeindex>
#include <vector>

// boost::bind(&mlmap::depth_odom_input_callback, this, _1, _2, _3)  (src/mlmap.cpp:143)
using namespace std::placeholders;
namespace boost {
template <class... A>
auto bind(A &&... a) -> decltype(std::bind(std::forward<A>(a)...)) {
  return std::bind(std::forward<A>(a)...);
}
template <class T>
using shared_ptr = std::shared_ptr<T>;
}  // namespace boost

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)

namespace ref_shim {
// last message published on every topic (type erased)
struct Captured {
  std::shared_ptr<void> msg;
  std::type_index type = std::type_index(typeid(void));
  long count = 0;
};
inline std::map<std::string, Captured> &topics() {
  static std::map<std::string, Captured> t;
  return t;
}
inline std::map<std::string, std::string> &params() {
  static std::map<std::string, std::string> p;
  return p;
}
template <class M>
const M *last(const std::string &topic) {
  auto it = topics().find(topic);
  if (it == topics().end() || it->second.type != std::type_index(typeid(M))) return nullptr;
  return static_cast<const M *>(it->second.msg.get());
}
}  // namespace ref_shim

namespace ros {
struct Duration {
  double s = 0;
  Duration() {}
  explicit Duration(double v) : s(v) {}
  double toSec() const { return s; }
};
struct Time {
  double s = 0;
  Time() {}
  explicit Time(double v) : s(v) {}
  static Time now() { return Time(); }
  double toSec() const { return s; }
};
inline Duration operator-(const Time &a, const Time &b) { return Duration(a.s - b.s); }
struct Timer {};
class Publisher {
 public:
  std::string topic;
  template <class M>
  void publish(const M &m) const {
    auto &c = ref_shim::topics()[topic];
    c.msg = std::make_shared<M>(m);
    c.type = std::type_index(typeid(M));
    c.count++;
  }
};
class NodeHandle {
 public:
  bool getParam(const std::string &key, std::string &out) const {
    auto it = ref_shim::params().find(key);
    if (it == ref_shim::params().end()) return false;
    out = it->second;
    return true;
  }
  template <class M>
  Publisher advertise(const std::string &topic, unsigned) {
    Publisher p;
    p.topic = topic;
    return p;
  }
  template <class F>
  Timer createTimer(Duration, F &&) {
    return Timer();
  }
  template <class F, class O>
  Timer createTimer(Duration, F, O *) {
    return Timer();
  }
};
}  // namespace ros

namespace std_msgs {
struct Header {
  ros::Time stamp;
  std::string frame_id;
  uint32_t seq = 0;
};
struct ColorRGBA {
  float r = 0, g = 0, b = 0, a = 0;
};
}  // namespace std_msgs

namespace geometry_msgs {
struct Point {
  double x = 0, y = 0, z = 0;
};
struct Vector3 {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 0;
};
struct Pose {
  Point position;
  Quaternion orientation;
};
struct PoseWithCovariance {
  Pose pose;
};
struct Twist {
  Vector3 linear, angular;
};
struct TwistWithCovariance {
  Twist twist;
};
struct PoseStamped {
  std_msgs::Header header;
  Pose pose;
};
struct Transform {
  Vector3 translation;
  Quaternion rotation;
};
struct TransformStamped {
  std_msgs::Header header;
  std::string child_frame_id;
  Transform transform;
};
}  // namespace geometry_msgs

namespace sensor_msgs {
namespace image_encodings {
const std::string TYPE_32FC1 = "32FC1";
const std::string TYPE_16UC1 = "16UC1";
}  // namespace image_encodings
struct Image {
  typedef std::shared_ptr<const Image> ConstPtr;
  typedef std::shared_ptr<Image> Ptr;
  std_msgs::Header header;
  uint32_t height = 0, width = 0, step = 0;
  std::string encoding;
  std::vector<uint8_t> data;
};
struct Imu {
  typedef std::shared_ptr<const Imu> ConstPtr;
  std_msgs::Header header;
  geometry_msgs::Vector3 angular_velocity, linear_acceleration;
  geometry_msgs::Quaternion orientation;
};
struct PointCloud2 {
  typedef std::shared_ptr<const PointCloud2> ConstPtr;
  std_msgs::Header header;
  uint32_t height = 0, width = 0, point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
};
}  // namespace sensor_msgs

namespace nav_msgs {
struct Odometry {
  typedef std::shared_ptr<const Odometry> ConstPtr;
  std_msgs::Header header;
  std::string child_frame_id;
  geometry_msgs::PoseWithCovariance pose;
  geometry_msgs::TwistWithCovariance twist;
};
struct OccupancyGrid {
  std_msgs::Header header;
  std::vector<int8_t> data;
};
}  // namespace nav_msgs

namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::Header header;
  std::string ns;
  int32_t id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
  std::vector<std_msgs::ColorRGBA> colors;
};
struct MarkerArray {
  std::vector<Marker> markers;
};
}  // namespace visualization_msgs

namespace tf2_ros {
struct TransformBroadcaster {
  void sendTransform(const geometry_msgs::TransformStamped &) {}
};
}  // namespace tf2_ros

namespace message_filters {
template <class M>
class Subscriber {
 public:
  void subscribe(ros::NodeHandle &, const std::string &, uint32_t) {}
};
namespace sync_policies {
template <class... M>
struct ApproximateTime {
  explicit ApproximateTime(uint32_t) {}
};
template <class... M>
struct ExactTime {
  explicit ExactTime(uint32_t) {}
};
}  // namespace sync_policies
template <class Policy>
class Synchronizer {
 public:
  template <class... S>
  Synchronizer(const Policy &, S &...) {}
  template <class F>
  void registerCallback(const F &) {}
};
}  // namespace message_filters

// ---- OpenCV / cv_bridge: a dense row-major image is all project_depth needs (src/mlmap.cpp:311-349,476-482) ----
#define CV_16UC1 2
#define CV_32FC1 5
namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0, type_ = CV_16UC1;
  std::shared_ptr<std::vector<uint8_t>> buf;
  size_t elem() const { return type_ == CV_32FC1 ? 4 : 2; }
  void create(int r, int c, int t) {
    rows = r;
    cols = c;
    type_ = t;
    buf = std::make_shared<std::vector<uint8_t>>((size_t)r * c * elem());
  }
  template <class T>
  T *ptr(int r) {
    return reinterpret_cast<T *>(buf->data() + (size_t)r * cols * elem());
  }
  template <class T>
  const T *ptr(int r) const {
    return reinterpret_cast<const T *>(buf->data() + (size_t)r * cols * elem());
  }
  void copyTo(Mat &dst) const {
    dst.rows = rows;
    dst.cols = cols;
    dst.type_ = type_;
    dst.buf = buf ? std::make_shared<std::vector<uint8_t>>(*buf) : nullptr;
  }
  // CV_32FC1 -> CV_16UC1 with a scale: saturate_cast<ushort>(cvRound(v * alpha)), round half to even
  void convertTo(Mat &dst, int rtype, double alpha = 1.0) const {
    Mat out;
    out.create(rows, cols, rtype);
    if (type_ == CV_32FC1 && rtype == CV_16UC1) {
      for (int r = 0; r < rows; r++) {
        const float *s = ptr<float>(r);
        uint16_t *d = out.ptr<uint16_t>(r);
        for (int c = 0; c < cols; c++) {
          long v = std::lrint((double)s[c] * alpha);
          d[c] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
        }
      }
    } else if (type_ == rtype && buf) {
      *out.buf = *buf;
    }
    dst = out;
  }
};
}  // namespace cv
namespace cv_bridge {
struct CvImage {
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
};
typedef std::shared_ptr<CvImage> CvImagePtr;
inline CvImagePtr toCvCopy(const sensor_msgs::Image::ConstPtr &src, const std::string &encoding = std::string()) {
  CvImagePtr p = std::make_shared<CvImage>();
  p->header = src->header;
  p->encoding = encoding.empty() ? src->encoding : encoding;
  const bool f32 = src->encoding == sensor_msgs::image_encodings::TYPE_32FC1;
  p->image.create((int)src->height, (int)src->width, f32 ? CV_32FC1 : CV_16UC1);
  const size_t row_bytes = (size_t)src->width * (f32 ? 4 : 2);
  for (uint32_t r = 0; r < src->height; r++)
    memcpy(p->image.ptr<uint8_t>(0) + r * row_bytes, src->data.data() + (size_t)r * src->step, row_bytes);
  return p;
}
}  // namespace cv_bridge

// ---- PCL: PointXYZ is 16 bytes {x, y, z, 1.0f} (pcl/impl/point_types.hpp); clouds are vectors ----
namespace pcl {
struct alignas(16) PointXYZ {
  float x, y, z, pad;
  PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad(1.0f) {}
};
struct PointXYZRGB : PointXYZ {
  uint32_t rgba = 0;
};
struct PointXYZRGBA : PointXYZ {
  uint32_t rgba = 0;
};
struct PointXYZI : PointXYZ {
  float intensity = 0;
};
struct PCLHeader {
  uint32_t seq = 0;
  uint64_t stamp = 0;
  std::string frame_id;
};
template <class P>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
  PCLHeader header;
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
};
// wire format of the reference's PointCloud2 topics: the cloud's points, point_step = sizeof(P)
template <class P>
void toROSMsg(const PointCloud<P> &cloud, sensor_msgs::PointCloud2 &msg) {
  msg.header.frame_id = cloud.header.frame_id;
  msg.height = cloud.height;
  msg.width = cloud.width;
  msg.point_step = sizeof(P);
  msg.row_step = (uint32_t)(sizeof(P) * cloud.points.size());
  msg.data.resize(sizeof(P) * cloud.points.size());
  if (!cloud.points.empty()) memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
}
}  // namespace pcl

#endif
