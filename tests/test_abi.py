"""CPU: the C-ABI library builds, loads and exports every symbol include/mlmap_b200.h declares;
ctypes struct layouts match; with no GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

import mlmapping_b200 as mlm
from mlmapping_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not mlm.library_path().exists():
        mlm.build_library()
    return mlm.load_library()


def _declared_symbols():
    text = (ROOT / "include" / "mlmap_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mlmap_b200.h but not exported"
    assert set(capi.ABI_SYMBOLS) == set(declared)


def test_struct_layouts(lib):
    assert lib.mlm_sizeof_config() == C.sizeof(capi.MlmConfig)
    assert lib.mlm_sizeof_frame_stats() == C.sizeof(capi.FrameStats)
    assert lib.mlm_abi_version() == 1


def test_default_config_is_config_sim_yaml(lib):
    c = mlm.default_config()  # reference launch/config/config_sim.yaml
    assert (c.am_d_rho, c.am_d_phi_deg, c.am_d_z) == (0.20, 5.0, 0.20)
    assert (c.am_n_rho, c.am_n_z_below, c.am_n_z_over) == (40, 20, 20)
    assert c.subbox_n == 10 and abs(c.subbox_d_xyz - 0.2) < 1e-15
    assert abs(c.log_odds_max - 4.2) < 1e-6 and abs(c.log_odds_miss + 0.9) < 1e-6
    assert c.cam_fx == pytest.approx(347.99755859375)
    assert c.sample_cnt == 500
    assert list(c.T_bs) == [0.12, 0.0, 0.0, 0.5, -0.5, 0.5, -0.5]


def test_invalid_config_and_no_silent_cpu_fallback(lib):
    import torch
    cfg = mlm.config_cfg_a()
    h = C.c_void_p()
    bad = cfg.copy()
    bad.am_d_rho = -1.0
    assert lib.mlm_create(C.byref(bad), 0, C.byref(h)) == 2  # MLM_ERR_INVALID_CONFIG
    bad = cfg.copy()
    bad.max_points = 0
    assert lib.mlm_create(C.byref(bad), 0, C.byref(h)) == 2
    assert lib.mlm_create(None, 0, C.byref(h)) == 1           # MLM_ERR_INVALID_ARG
    if not torch.cuda.is_available():
        rc = lib.mlm_create(C.byref(cfg), 0, C.byref(h))
        assert rc == 7 and b"no CPU fallback" in lib.mlm_last_error()
        with pytest.raises(mlm.MlmError):
            mlm.MLMap(cfg)
    # null handles are rejected, not dereferenced
    assert lib.mlm_sync(None) == 1
    assert lib.mlm_destroy(None) == 1
    n = C.c_size_t()
    assert lib.mlm_export_cloud(None, 0, None, 0, C.byref(n)) == 1
    assert lib.mlm_export_odds_slice(None, 0.7, None, 0, C.byref(n)) == 1
    assert lib.mlm_checkpoint_size(None, C.byref(n)) == 1
    assert lib.mlm_checkpoint_save(None, None, 0, C.byref(n)) == 1
    assert lib.mlm_checkpoint_restore(None, None, 0) == 1
    out7 = (C.c_double * 7)()
    assert lib.mlm_compensate_pose(None, None, None, None, 0.0, 0.0, 0.0, out7) == 1


def test_rand_stream_is_glibc_rand(lib):
    """sampled project_depth replays glibc's rand() (TYPE_3 random_r, seed 1) from a per-handle state"""
    import numpy as np
    out = np.zeros(4096, dtype=np.int32)
    assert lib.mlm_debug_rand(None, out.ctypes.data, out.size) == 0
    libc = C.CDLL("libc.so.6")
    libc.srand(1)
    ref = np.array([libc.rand() for _ in range(out.size)], dtype=np.int32)
    assert np.array_equal(out, ref)


def test_multi_gpu_and_layer_entry_points_reject_bad_arguments_without_a_device(lib):
    """the sharded-map, replicated-map, two-call, asynchronous-frame and index-query entry points validate their handle
    and arguments before they touch a device (MLM_ERR_INVALID_ARG = 1)"""
    blob = (C.c_ubyte * capi.SHARD_BLOB_BYTES)()
    pose = (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0)
    n32 = C.c_int32()
    assert lib.mlm_shard_open(None, 0, 1, blob) == 1
    assert lib.mlm_shard_connect(None, blob) == 1
    assert lib.mlm_shard_submit_points_f64(None, None, 0, pose) == 1
    assert lib.mlm_shard_submit_points_f64_device(None, None, 0, pose) == 1
    assert lib.mlm_shard_submit_points_slice_f64(None, None, 0, 0, 0, pose) == 1
    assert lib.mlm_shard_finish(None, None) == 1
    assert lib.mlm_shard_integrate_points_f64(None, None, 0, pose, None) == 1
    assert lib.mlm_shard_last_exchange(None, None) == 1
    assert lib.mlm_shard_last_kernel_ms(None, None) == 1
    assert lib.mlm_shard_close(None) == 1
    assert lib.mlm_replica_open(None, 0, 1, 0, blob) == 1
    assert lib.mlm_replica_connect(None, blob) == 1
    assert lib.mlm_replica_publish(None, C.byref(n32)) == 1
    assert lib.mlm_replica_apply(None, C.byref(n32)) == 1
    assert lib.mlm_replica_close(None) == 1
    assert lib.mlm_dirty_count(None, None, None) == 1
    assert lib.mlm_awareness_input_pc_pose_f64(None, None, 0, pose, None) == 1
    assert lib.mlm_awareness_input_depth_u16(None, None, 0, 0, 0, pose, None) == 1
    assert lib.mlm_local_input_pc_pose_direct(None, None) == 1
    assert lib.mlm_set_sm_budget(None, 4) == 1
    assert lib.mlm_frame_submit_depth_u16_device(None, None, 0, 0, pose) == 1
    assert lib.mlm_frame_submit_points_f64_device(None, None, 0, pose) == 1
    assert lib.mlm_frame_finish(None, None) == 1
    assert lib.mlm_get_odd_at(None, None, None, 1, None) == 1
    assert lib.mlm_get_odd_at_device(None, None, None, 1, None) == 1
