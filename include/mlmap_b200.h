/*
 * mlmap_b200.h — C ABI of the B200-native MLMapping hot path (the drop-in boundary).
 *
 * The reference has no FFI layer: its boundary is the C++ class `mlmap`
 * (reference include/mlmap.h:105-139).  A maintainer binds these entry points
 * from that class; every function below cites the
 * reference method it replaces.  Plain pointers and sizes only; no C++/torch
 * types cross this boundary; nothing throws.
 *
 * Threading contract: one handle owns one CUDA stream.  Calls on a
 * handle are stream-ordered and must be externally serialised; distinct handles
 * are independent.  Queries observe every previously submitted update.
 *
 * Pose layout everywhere: double[7] = { tx, ty, tz, qw, qx, qy, qz }.  The
 * quaternion is normalised on entry exactly as Sophus::SO3(Quaterniond) does
 * (reference 3rdPartLib/Sophus/sophus/so3.cpp:42-47).
 */
#ifndef MLMAP_B200_H
#define MLMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MLM_ABI_VERSION 1

/* status codes returned by every entry point */
enum {
  MLM_OK = 0,
  MLM_ERR_INVALID_ARG = 1,    /* null pointer, bad size, bad handle */
  MLM_ERR_INVALID_CONFIG = 2, /* violates a precondition of the reference's tables */
  MLM_ERR_CUDA = 3,           /* a CUDA runtime call failed; see mlm_last_error() */
  MLM_ERR_POOL_EXHAUSTED = 4, /* submap pool or hash table full */
  MLM_ERR_CAPACITY = 5,       /* more points/records in one frame than the handle was sized for */
  MLM_ERR_UNSUPPORTED = 6,    /* option not implemented on the GPU path (never a silent CPU fallback) */
  MLM_ERR_NO_DEVICE = 7       /* no CUDA device / wrong architecture */
};

/* occupancy codes, reference include/mlmap.h:109-114 */
enum { MLM_FREE = 1, MLM_OCCUPIED = 0, MLM_UNKNOWN = -1 };

/*
 * Configuration = the YAML keys mlmap::init_map reads (reference src/mlmap.cpp:10-33,37-53,75-85)
 * as a plain struct (yamlRead.h is out of scope).
 */
typedef struct mlm_config {
  /* awareness_map_cylindrical::init_map arguments (reference src/map_awareness.cpp:19) */
  double am_d_rho;      /* mlmapping_am_d_Rho      */
  double am_d_phi_deg;  /* mlmapping_am_d_Phi_deg  */
  double am_d_z;        /* mlmapping_am_d_Z        */
  int32_t am_n_rho;     /* mlmapping_am_n_Rho      */
  int32_t am_n_z_below; /* mlmapping_am_n_Z_below  */
  int32_t am_n_z_over;  /* mlmapping_am_n_Z_over   */
  int32_t use_raycasting; /* mlmapping_use_raycasting */
  int32_t _pad0;
  double depth_noise_coe; /* mlmapping_depth_noise_coe */

  /* local_map_cartesian::init_map arguments (reference src/map_local.cpp:46-53) */
  double subbox_d_xyz;  /* mlmapping_subbox_d_xyz */
  int32_t subbox_n;     /* mlmapping_subbox_n     */
  float log_odds_min;   /* mlmapping_lm_log_odds_min */
  float log_odds_max;   /* mlmapping_lm_log_odds_max */
  float log_odds_hit;   /* mlmapping_lm_measurement_hit (read, stored, never used: map_local.cpp:128) */
  float log_odds_miss;  /* mlmapping_lm_measurement_miss */
  float log_odds_occupied_sh; /* mlmapping_lm_occupied_sh */
  int32_t use_exploration_frontiers; /* use_exploration_frontiers */
  int32_t _pad1;

  /* camera intrinsics, float members in the reference (include/mlmap.h:92) */
  float cam_cx, cam_cy, cam_fx, cam_fy;

  /* T_B_S (sensor in body) as pose[7]; reference src/mlmap.cpp:22-25 */
  double T_bs[7];

  /* inflation layer (reference src/mlmap.cpp:10,84-85; include/map_local.h:63-65) */
  int32_t inflate_n;
  int32_t inflate_global_n;
  int32_t apply_inflate;
  int32_t _pad2;
  double inflate_height; /* local_map_cartesian::flate_height, default 0.1 */

  /* project_depth: 0 = full-frame mode (every non-zero pixel, row-major);
   * >0 = mlmapping_sample_cnt random pixels via glibc rand() (reference src/mlmap.cpp:321-326) */
  int32_t sample_cnt;

  /* capacities of the GPU-resident structures (no reference equivalent: the reference mallocs) */
  int32_t max_points;     /* max points (pixels) per frame */
  int32_t pool_submaps;   /* capacity of the submap block pool */
  int32_t _pad3;
} mlm_config;

typedef struct mlm_map *mlm_handle;

/* per-frame counters (reference public counters ram_expand_cnt/obs_cnt, include/map_local.h:79-80) */
typedef struct mlm_frame_stats {
  int32_t n_points;        /* points fed to input_pc_pose */
  int32_t n_inside;        /* points inside the awareness range (update_hits calls) */
  int32_t n_cast;          /* rays cast (can_do_cast && visibility_check) */
  int32_t n_hit_cells;     /* distinct keys of hit_idx_odds_hashmap */
  int32_t n_miss_cells;    /* distinct entries of miss_idx_set */
  int32_t n_touched_voxels;/* distinct local-map cells receiving any update */
  int32_t n_new_submaps;   /* allocate_ram() insertions this frame */
  int32_t hit_bucket_count;/* emulated libstdc++ bucket count of the hit map after this frame */
  int32_t ordering_slow_path; /* 1 if the frame crossed a libstdc++ rehash */
  int32_t status;          /* MLM_OK or the error raised on device */
  int64_t ram_expand_cnt;  /* cumulative */
  int64_t obs_cnt;         /* cumulative */
} mlm_frame_stats;

/* ---- lifetime ------------------------------------------------------------------------------- */

/* Fill cfg with the live reference configuration (launch/config/config_sim.yaml) */
int mlm_default_config(mlm_config *cfg);

/* mlmap::init_map (reference src/mlmap.cpp:3-149) minus ROS: builds the awareness tables
 * (map_awareness.cpp:19-82) and local-map tables (map_local.cpp:46-139), allocates device state. */
int mlm_create(const mlm_config *cfg, int device, mlm_handle *out);
int mlm_destroy(mlm_handle h);
const char *mlm_last_error(void);
int mlm_abi_version(void);
/* sizeof(mlm_config) / sizeof(mlm_frame_stats) as compiled, for binding self-checks */
size_t mlm_sizeof_config(void);
size_t mlm_sizeof_frame_stats(void);

/* ---- per-frame update (the north-star path) --------------------------------------------------- */

/* project_depth() + update_map() (reference src/mlmap.cpp:311-349,382-386) on a HOST depth image
 * (uint16 millimetres, row stride in bytes).  H2D copy, all kernels and the stats read-back are
 * inside the call. */
int mlm_integrate_depth_u16(mlm_handle h, const uint16_t *img, int rows, int cols, size_t stride_bytes,
                            const double T_wb[7], mlm_frame_stats *stats /* may be NULL */);
/* same with the image already resident in device memory (dense rows) */
int mlm_integrate_depth_u16_device(mlm_handle h, const uint16_t *d_img, int rows, int cols,
                                   const double T_wb[7], mlm_frame_stats *stats);

/* awareness_map_cylindrical::input_pc_pose + local_map_cartesian::input_pc_pose_direct
 * (reference src/map_awareness.cpp:173, src/map_local.cpp:143) on sensor-frame points, xyz doubles. */
int mlm_integrate_points_f64(mlm_handle h, const double *xyz, int n, const double T_wb[7],
                             mlm_frame_stats *stats);
int mlm_integrate_points_f64_device(mlm_handle h, const double *d_xyz, int n, const double T_wb[7],
                                    mlm_frame_stats *stats);

/* The two layer entry points behind mlmap's public `awareness_map` / `local_map` pointers, which mlmap::update_map calls
 * one after the other (reference src/mlmap.cpp:382-386, include/mlmap.h:107-108):
 *   mlm_awareness_input_*            awareness_map_cylindrical::input_pc_pose (include/map_awareness.h:74,
 *                                    src/map_awareness.cpp:173-282): the frame's hit map and miss set, readable with
 *                                    mlm_last_frame_hits / mlm_last_frame_misses; the depth variant runs project_depth first
 *   mlm_local_input_pc_pose_direct   local_map_cartesian::input_pc_pose_direct(awareness_map) (include/map_local.h:114,
 *                                    src/map_local.cpp:143-237): fuses the staged sets into the local / global map
 * One difference is visible between the two calls: the subboxes the frame touches are allocated by the first call already
 * (all-unknown cells, so every query still answers like the reference; only mlm_export_map_count is ahead).
 * mlm_integrate_* = both calls as ONE cooperative launch. */
int mlm_awareness_input_pc_pose_f64(mlm_handle h, const double *xyz, int n, const double T_wb[7], mlm_frame_stats *stats);
int mlm_awareness_input_depth_u16(mlm_handle h, const uint16_t *img, int rows, int cols, size_t stride_bytes,
                                  const double T_wb[7], mlm_frame_stats *stats);
int mlm_local_input_pc_pose_direct(mlm_handle h, mlm_frame_stats *stats);

/* Asynchronous frames, for several maps of one process on one GPU (BASELINE config 5: 8 agent maps): submit enqueues the
 * frame on the handle's stream and returns; mlm_frame_finish waits for it, settles a rehash frame and returns the
 * counters.  mlm_set_sm_budget(h, n) makes the handle's frame kernel use n CTAs instead of one per SM (0 restores that),
 * so that the cooperative launches of several handles are co-resident and overlap. */
int mlm_set_sm_budget(mlm_handle h, int n_sms);
int mlm_frame_submit_depth_u16_device(mlm_handle h, const uint16_t *d_img, int rows, int cols, const double T_wb[7]);
int mlm_frame_submit_points_f64_device(mlm_handle h, const double *d_xyz, int n, const double T_wb[7]);
int mlm_frame_finish(mlm_handle h, mlm_frame_stats *stats /* may be NULL */);

/* mlmap::setFree_map_in_bound (reference src/mlmap.cpp:388-407) */
int mlm_set_free_in_bound(mlm_handle h, const double box_min[3], const double box_max[3]);

/* mlmap::inflate_map (reference src/mlmap.cpp:286-309) around centre position ct_pos */
int mlm_inflate_map(mlm_handle h, const double ct_pos[3]);

/* ---- batched queries (reference include/mlmap.h:142-295); pos = n x 3 doubles ------------------ */

int mlm_get_occupancy(mlm_handle h, const double *pos, size_t n, int32_t *out);
int mlm_get_occupancy_inflate(mlm_handle h, const double *pos, size_t n, float inflate, int32_t *out);
int mlm_get_inflate_occupancy(mlm_handle h, const double *pos, size_t n, int32_t *out);
int mlm_get_odd(mlm_handle h, const double *pos, size_t n, float *out);
int mlm_get_odd_grad(mlm_handle h, const double *pos, size_t n, size_t max_iter, double *out3n);
/* getOdd(const Vec3I &glb_id, size_t subbox_id) (include/mlmap.h:128,227-235): glb3 = n x 3 subbox indices, sub = n cell ids */
int mlm_get_odd_at(mlm_handle h, const int32_t *glb3, const int32_t *sub, size_t n, float *out);
int mlm_get_odd_at_device(mlm_handle h, const int32_t *d_glb3, const int32_t *d_sub, size_t n, float *d_out);
/* device-resident variants: pos/out are device pointers, the call only enqueues on the handle's stream */
int mlm_get_occupancy_device(mlm_handle h, const double *d_pos, size_t n, int32_t *d_out);
int mlm_get_odd_device(mlm_handle h, const double *d_pos, size_t n, float *d_out);
int mlm_get_odd_grad_device(mlm_handle h, const double *d_pos, size_t n, size_t max_iter, double *d_out3n);

/* ---- stream control / timing (CUDA events on the handle's own stream) ------------------------- */

int mlm_sync(mlm_handle h);
int mlm_timer_start(mlm_handle h);
int mlm_timer_stop_ms(mlm_handle h, float *ms); /* records + synchronises; ms since mlm_timer_start */
/* device scratch helpers for harnesses that keep inputs resident */
int mlm_device_alloc(mlm_handle h, size_t bytes, void **d_ptr);
int mlm_device_free(mlm_handle h, void *d_ptr);
/* page-locked host buffers: images / point arrays placed here are DMA'd without the staging copy */
int mlm_host_alloc(mlm_handle h, size_t bytes, void **ptr);
int mlm_host_free(mlm_handle h, void *ptr);
int mlm_copy_to_device(mlm_handle h, void *d_dst, const void *src, size_t bytes);
int mlm_copy_to_host(mlm_handle h, void *dst, const void *d_src, size_t bytes);
int mlm_flush_l2(mlm_handle h); /* writes a buffer larger than L2 (bench hygiene) */
/* per-kernel timing of the frame pipeline (CUDA events between launches on the handle's stream).
 * enable=1 records events in every following frame; mlm_last_frame_kernel_ms returns the durations
 * of the MLM_NUM_FRAME_KERNELS stages of the last frame in launch order:
 * 0 k_project, 1 k_column, 2 k_fuse */
#define MLM_NUM_FRAME_KERNELS 3
int mlm_set_profiling(mlm_handle h, int enable);
int mlm_last_frame_kernel_ms(mlm_handle h, float ms[MLM_NUM_FRAME_KERNELS]);
/* per-column phase clocks of the last k_column launch (only filled by -DMLM_PHASE_TIMING builds) */
int mlm_debug_phase_cycles(mlm_handle h, long long *out, size_t cap);
/* sampled project_depth (cfg.sample_cnt > 0) draws pixels from the handle's own glibc-compatible rand() stream
 * (seed 1 like a process that never calls srand); mlm_srand reseeds it, mlm_debug_rand draws from it
 * (h == NULL: from a fresh seed-1 stream) so tests can compare with libc's rand() */
int mlm_srand(mlm_handle h, unsigned seed);
int mlm_debug_rand(mlm_handle h, int32_t *out, size_t n);
/* number of kernels launched by this handle since creation */
int mlm_kernel_launch_count(mlm_handle h, int64_t *count);

/* ---- parity / debug exports ------------------------------------------------------------------- */

/* Last frame's hit_idx_odds_hashmap in the reference's iteration order: keys = n x 3 ints
 * (rho,phi,z), p = n floats.  Returns the number of entries through *n_out (cap = array capacity). */
int mlm_last_frame_hits(mlm_handle h, int32_t *keys3, float *p, size_t cap, size_t *n_out);
/* Last frame's miss_idx_set (awareness cell indices, ascending). */
int mlm_last_frame_misses(mlm_handle h, uint64_t *idx, size_t cap, size_t *n_out);

/* Whole-map export.  mlm_export_map_count gives the number of submaps; mlm_export_map fills
 * glb3 (n x 3 int32), collapsed (n bytes), and per submap cells = subbox_n^3 entries of
 * occupancy chars ('u','f','o'), inflate chars and float log-odds (collapsed submaps: entry 0 valid). */
int mlm_export_map_count(mlm_handle h, size_t *n_submaps);
int mlm_export_map(mlm_handle h, size_t cap_submaps, int32_t *glb3, uint8_t *collapsed,
                   char *occupancy, char *inflate_occupancy, float *log_odds, size_t *n_out);

/* Exploration mode: frontier set of every subbox (reference struct subbox::frontier, include/map_local.h:58) as a
 * bitmask of ceil(subbox_n^3 / 32) 32-bit words per subbox (bit c = cell id c), with its own glb3 list. */
int mlm_export_frontier(mlm_handle h, size_t cap_submaps, int32_t *glb3, uint32_t *frontier_words, size_t *n_out);

/* ---- pose forwarding of the ROS callback, host-only, no handle ---------------------------------------
 * mlmap::depth_odom_input_callback (src/mlmap.cpp:470-498) forwards the odometry pose to the image stamp with a linear
 * model before the update:  time_gap = gap_imu - latency;  rot_cp = log(R) + time_gap * (R * omega);
 * T_wb = (exp(rot_cp), p + (gap_odom - latency) * v), with Sophus' SO3::log / SO3::exp (so3.cpp:127-199).
 * gap_* = image stamp - odometry / IMU stamp in seconds; quaternion order w,x,y,z; T_wb_out = {x,y,z,qw,qx,qy,qz},
 * ready for mlm_integrate_*. */
int mlm_compensate_pose(const double odom_pos[3], const double odom_quat_wxyz[4], const double odom_lin_vel[3],
                        const double imu_ang_vel[3], double gap_odom_s, double gap_imu_s, double camera2odom_latency_s,
                        double T_wb_out[7]);

/* ---- map clouds for consumers ---------------------------------------------------------------------
 * Whole-map compaction on the device into float4 points {x, y, z, w} (16-byte stride = pcl::PointXYZ, the wire format
 * of the reference's PointCloud2 topics, include/common.h:53):
 *   MLM_CLOUD_INFLATED  cells with inflate_occupancy == 'o'  (rviz_vis::pub_global_local_map, src/rviz_vis.cpp:296-327)
 *   MLM_CLOUD_OCCUPIED  cells with occupancy == 'o'
 *   MLM_CLOUD_FRONTIER  the frontier sets                     (rviz_vis::pub_frontier, src/rviz_vis.cpp:267-294)
 * w is 1.0f.  mlm_export_odds_slice writes the cells whose centre height is within 1e-3 of `height` as
 * {x, y, z, odd}, odd = logit_inv(log_odds)  (mlmap::visualize_odds, src/mlmap.cpp:200-284).
 * `xyzw` holds `cap` points (host memory for the plain calls, device memory for *_device); *n_out is the number of
 * points the map has, which may exceed cap (then only cap points were written).  Point order is unspecified. */
#define MLM_CLOUD_INFLATED 0
#define MLM_CLOUD_OCCUPIED 1
#define MLM_CLOUD_FRONTIER 2
int mlm_export_cloud(mlm_handle h, int kind, float *xyzw, size_t cap, size_t *n_out);
int mlm_export_cloud_device(mlm_handle h, int kind, float *d_xyzw, size_t cap, size_t *n_out);
int mlm_export_odds_slice(mlm_handle h, double height, float *xyzw, size_t cap, size_t *n_out);

/* ---- checkpoint / restore of the map (the reference keeps its map in process memory only) ------------
 * A checkpoint is a self-describing byte image: a header (format version, the map-defining configuration fields,
 * the cumulative counters, the emulated bucket counts of the per-frame containers and the rand() stream of the
 * sampled projection) followed by one fixed-size record per subbox (index, collapsed flag, log_odds, occupancy,
 * inflate_occupancy and, in exploration mode, the frontier set).  Restoring it into a handle created with the same
 * map-defining configuration (any pool size that holds the subboxes) replaces that handle's map; frames integrated
 * afterwards give bit for bit what the original handle would have produced.
 *   mlm_checkpoint_size     bytes mlm_checkpoint_save will write for the current map
 *   mlm_checkpoint_save     writes at most `cap` bytes into host memory `buf`; MLM_ERR_CAPACITY if cap is too small
 *   mlm_checkpoint_restore  MLM_ERR_INVALID_ARG for a foreign / truncated image, MLM_ERR_INVALID_CONFIG when the image
 *                           was taken with a different map configuration, MLM_ERR_POOL_EXHAUSTED if the pool is too small */
int mlm_checkpoint_size(mlm_handle h, size_t *bytes);
int mlm_checkpoint_save(mlm_handle h, void *buf, size_t cap, size_t *written);
int mlm_checkpoint_restore(mlm_handle h, const void *buf, size_t bytes);

/* ---- one logical map sharded over `world` ranks (large LiDAR scans) -------------------------------------
 * The sharded form of awareness_map_cylindrical::input_pc_pose + local_map_cartesian::input_pc_pose_direct
 * (reference src/map_awareness.cpp:173-282, src/map_local.cpp:143-207) for scans too large / too frequent for one GPU:
 * every rank gets the SAME points and pose, casts the phi columns `phi % world == rank`, and owns the subboxes whose
 * index hashes to it.  Between the two stages the library itself exchanges (a) every rank's distinct hit keys with
 * their first-insert stamps (all-gather, so that all ranks derive the same unordered_map iteration order) and (b) one
 * 24-byte update record per touched voxel and hit key, sent to the voxel's owner (all-to-all).  Both exchanges are
 * plain stores into the destination's exchange arena over NVLink peer memory followed by a system-scope flag; the
 * owner waits on device.  There is no host synchronisation between mlm_shard_submit_* and mlm_shard_finish and no
 * collective-library call on the path.  The union of the ranks' subboxes equals the single-GPU map bit for bit.
 *
 * Setup (once):  mlm_shard_open on every rank -> exchange the MLM_SHARD_BLOB_BYTES blobs by any means (MPI_Allgather,
 * torch.distributed.all_gather, a file) -> mlm_shard_connect with all `world` blobs in rank order.  Ranks may be
 * processes (one GPU each; arenas mapped with CUDA IPC) or handles of one process (peer access / same device).
 * Per scan:  mlm_shard_submit_points_f64[_device] enqueues the whole scan on the handle's stream and returns;
 * mlm_shard_finish waits for it, runs the rare rehash path (a scan whose distinct hit keys exceed the emulated bucket
 * count: each rank re-sequences the gathered list on its own) and returns the counters.  One process driving several
 * handles must submit on ALL of them before finishing any.  mlm_shard_integrate_points_f64 = submit + finish.
 * A peer that fails or does not signal within MLM_SHARD_TIMEOUT_MS (default 10000) makes finish return MLM_ERR_CUDA. */
#define MLM_SHARD_BLOB_BYTES 128
#define MLM_SHARD_MAX_WORLD 16
typedef struct mlm_shard_exchange {
  int32_t world;
  int32_t n_hit_total;      /* distinct hit keys of the scan over all ranks */
  int32_t n_hit_local;      /* ... cast by this rank */
  int32_t records_received; /* update records this rank ingested (all sources) */
  int32_t records_from_self; /* -1: not tracked (sources reserve straight in the owner's inbox) */
  int32_t rehash_path;      /* 1: the scan took the rehash path */
  int64_t wait_ns;          /* time this rank's wait kernel spun for its peers */
  int64_t arena_bytes;      /* size of this rank's exchange arena */
} mlm_shard_exchange;
int mlm_shard_open(mlm_handle h, int rank, int world, void *blob_out /* MLM_SHARD_BLOB_BYTES */);
int mlm_shard_connect(mlm_handle h, const void *blobs /* world * MLM_SHARD_BLOB_BYTES, rank order */);
int mlm_shard_submit_points_f64(mlm_handle h, const double *xyz, int n, const double T_wb[7]);
int mlm_shard_submit_points_f64_device(mlm_handle h, const double *d_xyz, int n, const double T_wb[7]);
/* the same scan handed over in slices: rank r passes points [first, first + n_slice) of the n_total points (the ranks'
 * slices partition the scan; xyz_slice is host memory, page-locked for a direct copy).  Every rank copies only its
 * slice from the host; the slices reach the other ranks' arenas over NVLink peer memory and every rank waits on the
 * device until the scan is complete. */
int mlm_shard_submit_points_slice_f64(mlm_handle h, const double *xyz_slice, int first, int n_slice, int n_total, const double T_wb[7]);
int mlm_shard_finish(mlm_handle h, mlm_frame_stats *stats /* may be NULL */);
int mlm_shard_integrate_points_f64(mlm_handle h, const double *xyz, int n, const double T_wb[7], mlm_frame_stats *stats);
int mlm_shard_last_exchange(mlm_handle h, mlm_shard_exchange *out);
/* with mlm_set_profiling(h, 1): durations of the stages of the last scan between CUDA events on the handle's stream:
 * 0 resets (+ H2D copy), 1 k_project, 2 k_column (the rank's phi columns), 3 k_shard_push (keys + records to the owners, signal),
 * 4 k_shard_act_ingest (wait for the sources, activation stamps, staging of the received records), 5 k_fuse */
#define MLM_NUM_SHARD_KERNELS 6
int mlm_shard_last_kernel_ms(mlm_handle h, float ms[MLM_NUM_SHARD_KERNELS]);
int mlm_shard_close(mlm_handle h); /* unmaps the peers' arenas; call on every rank before any rank is destroyed */

/* ---- replicated map for split query streams: after a frame the updating rank exports the subbox
 * blocks that frame touched (mlm_dirty_count gives their number and the record size), the caller broadcasts the
 * buffer (NCCL), replicas apply it with mlm_dirty_import; queries on a replica then equal the owner's. */
int mlm_dirty_count(mlm_handle h, int32_t *n_blocks, size_t *record_bytes);
int mlm_dirty_export(mlm_handle h, void *d_out, int32_t n_blocks);
int mlm_dirty_import(mlm_handle h, const void *d_in, int32_t n_blocks);

/* The same replication without a collective of the caller: every rank opens an inbox arena (mlm_replica_open fills
 * a MLM_SHARD_BLOB_BYTES setup blob), the `world` blobs are exchanged once (any transport) and handed to
 * mlm_replica_connect in rank order.  After each frame the source rank calls mlm_replica_publish: one kernel packs the
 * frame's dirty blocks straight into every replica's inbox over NVLink peer memory (CUDA IPC between processes) and
 * raises the frame's flag; every replica calls mlm_replica_apply: one kernel waits for the flag (bounded device-side
 * spin, MLM_SHARD_TIMEOUT_MS), applies the records and acknowledges to the source.  Inboxes are double-buffered by
 * frame parity: the source may be two frames ahead of the slowest replica.  Calls come in lock step, one publish and
 * one apply per frame.  n_blocks_out (may be NULL) receives the number of blocks shipped / applied. */
int mlm_replica_open(mlm_handle h, int rank, int world, int src_rank, void *blob_out /* MLM_SHARD_BLOB_BYTES */);
int mlm_replica_connect(mlm_handle h, const void *blobs /* world * MLM_SHARD_BLOB_BYTES, rank order */);
int mlm_replica_publish(mlm_handle h, int32_t *n_blocks_out);
int mlm_replica_apply(mlm_handle h, int32_t *n_blocks_out);
int mlm_replica_close(mlm_handle h); /* unmaps the peers' arenas; call on every rank before any rank is destroyed */

/* glibc-2.39 log10f as evaluated on device (parity test hook, SURVEY §7 hard part 3) */
int mlm_debug_log10f(mlm_handle h, const float *x, size_t n, float *out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MLMAP_B200_H */
