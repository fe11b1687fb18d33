"""mlmapping_b200 — B200-native (sm_100a CUDA) implementation of MLMapping's per-frame map-update
hot path and batched queries behind the reference's ``mlmap`` method surface.

The compute lives in ``csrc/`` (CUDA kernels + a C ABI, ``include/mlmap_b200.h``).  This Python
package is only the host-side mirror used by the ROS-free harness, tests and ``bench.py``; it fails
loudly if the CUDA library is missing — there is no CPU fallback.
"""
from .capi import (  # noqa: F401
    MLMap,
    MlmConfig,
    FrameStats,
    MlmError,
    load_library,
    library_path,
    build_library,
    default_config,
    config_cfg_a,
    config_cfg_b,
    config_cfg_c,
    compensate_pose,
)
