// ORACLE — TEST INFRASTRUCTURE ONLY (see mlmap_oracle.hpp header).  C ABI over the CPU
// restatement so tests/ and bench.py's cpu_baseline leg can drive it through ctypes.
// Pinned against oracle/_ref (the reference's own sources, oracle/ref_build) by tests/test_reference_pin.py.
#include "mlmap_oracle.hpp"

#include <chrono>

using namespace orc;

extern "C" {

void *orc_create(const mlm_config *cfg) { return new orc::mlmap(*cfg); }
void orc_destroy(void *h) { delete static_cast<orc::mlmap *>(h); }
void orc_set_bookkeeping(void *h, int on) { static_cast<orc::mlmap *>(h)->local->bookkeeping = on != 0; }

// project_depth() + update_map(), the region the reference times (src/mlmap.cpp:474-511).
// Returns the wall time of exactly that region in seconds.
double orc_integrate_depth_u16(void *h, const uint16_t *img, int rows, int cols, size_t stride_bytes,
                               const double T_wb[7]) {
  auto *m = static_cast<orc::mlmap *>(h);
  m->pc_eigen.clear();  // depth_odom_input_callback, src/mlmap.cpp:470
  m->T_wb = se3_from_pose7(T_wb);
  m->ct_pos = Vec3(T_wb[0], T_wb[1], T_wb[2]);
  auto t0 = std::chrono::steady_clock::now();
  m->project_depth(img, rows, cols, stride_bytes);
  m->update_map();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

double orc_integrate_points_f64(void *h, const double *xyz, int n, const double T_wb[7]) {
  auto *m = static_cast<orc::mlmap *>(h);
  m->pc_eigen.clear();
  m->pc_eigen.reserve(n);
  for (int i = 0; i < n; i++) m->pc_eigen.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  m->T_wb = se3_from_pose7(T_wb);
  m->ct_pos = Vec3(T_wb[0], T_wb[1], T_wb[2]);
  auto t0 = std::chrono::steady_clock::now();
  m->update_map();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

void orc_frame_stats(void *h, mlm_frame_stats *s) {
  auto *m = static_cast<orc::mlmap *>(h);
  memset(s, 0, sizeof(*s));
  s->n_points = (int32_t)m->pc_eigen.size();
  s->n_inside = (int32_t)m->awareness->n_inside;
  s->n_cast = (int32_t)m->awareness->n_cast;
  s->n_hit_cells = (int32_t)m->awareness->hit_idx_odds_hashmap.size();
  s->n_miss_cells = (int32_t)m->awareness->miss_idx_set.size();
  s->n_touched_voxels = (int32_t)m->local->n_touched_last;
  s->hit_bucket_count = (int32_t)m->awareness->hit_idx_odds_hashmap.bucket_count();
  s->ram_expand_cnt = m->local->ram_expand_cnt;
  s->obs_cnt = m->local->obs_cnt;
}

size_t orc_num_points(void *h) { return static_cast<orc::mlmap *>(h)->pc_eigen.size(); }
size_t orc_get_points(void *h, double *xyz, size_t cap) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t n = std::min(cap, m->pc_eigen.size());
  for (size_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) xyz[3 * i + k] = m->pc_eigen[i][k];
  return m->pc_eigen.size();
}

// hit map in its iteration order (the order map_local.cpp:147 consumes it in)
size_t orc_last_hits(void *h, int32_t *keys3, float *p, size_t cap) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t i = 0;
  for (auto &kv : m->awareness->hit_idx_odds_hashmap) {
    if (i < cap) {
      keys3[3 * i] = kv.first[0];
      keys3[3 * i + 1] = kv.first[1];
      keys3[3 * i + 2] = kv.first[2];
      p[i] = kv.second;
    }
    i++;
  }
  return i;
}
size_t orc_last_misses(void *h, uint64_t *idx, size_t cap) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t i = 0;
  for (auto v : m->awareness->miss_idx_set) {
    if (i < cap) idx[i] = v;
    i++;
  }
  return i;
}

void orc_set_log_inserts(void *h, int on) { static_cast<orc::mlmap *>(h)->awareness->log_inserts = on != 0; }
size_t orc_insert_log(void *h, int32_t *keys3, size_t cap) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t n = m->awareness->insert_log.size();
  for (size_t i = 0; i < n && i < cap; i++)
    for (int k = 0; k < 3; k++) keys3[3 * i + k] = m->awareness->insert_log[i][k];
  return n;
}

void orc_set_free_in_bound(void *h, const double mn[3], const double mx[3]) {
  static_cast<orc::mlmap *>(h)->setFree_map_in_bound(Vec3(mn[0], mn[1], mn[2]), Vec3(mx[0], mx[1], mx[2]));
}
void orc_inflate_map(void *h, const double ct[3]) {
  auto *m = static_cast<orc::mlmap *>(h);
  m->ct_pos = Vec3(ct[0], ct[1], ct[2]);
  m->inflate_map();
}

void orc_get_occupancy(void *h, const double *pos, size_t n, int32_t *out) {
  auto *m = static_cast<orc::mlmap *>(h);
  for (size_t i = 0; i < n; i++) out[i] = m->getOccupancy(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void orc_get_occupancy_inflate(void *h, const double *pos, size_t n, float inflate, int32_t *out) {
  auto *m = static_cast<orc::mlmap *>(h);
  for (size_t i = 0; i < n; i++)
    out[i] = m->getOccupancy(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), inflate);
}
void orc_get_inflate_occupancy(void *h, const double *pos, size_t n, int32_t *out) {
  auto *m = static_cast<orc::mlmap *>(h);
  for (size_t i = 0; i < n; i++)
    out[i] = m->getInflateOccupancy(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void orc_get_odd(void *h, const double *pos, size_t n, float *out) {
  auto *m = static_cast<orc::mlmap *>(h);
  for (size_t i = 0; i < n; i++) out[i] = m->getOdd(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void orc_get_odd_grad(void *h, const double *pos, size_t n, size_t max_iter, double *out) {
  auto *m = static_cast<orc::mlmap *>(h);
  for (size_t i = 0; i < n; i++) {
    Vec3 g = m->getOddGrad(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), max_iter);
    out[3 * i] = g[0];
    out[3 * i + 1] = g[1];
    out[3 * i + 2] = g[2];
  }
}

size_t orc_export_map_count(void *h) { return static_cast<orc::mlmap *>(h)->local->observed_group_map.size(); }
// Same layout as mlm_export_map.  Unordered; callers sort by glb index.
size_t orc_export_map(void *h, size_t cap, int32_t *glb3, uint8_t *collapsed, char *occ, char *infl,
                      float *lo) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t cells = m->local->cell_num_subbox;
  size_t i = 0;
  for (auto &kv : m->local->observed_group_map) {
    if (i < cap) {
      glb3[3 * i] = kv.first[0];
      glb3[3 * i + 1] = kv.first[1];
      glb3[3 * i + 2] = kv.first[2];
      const auto &sb = kv.second;
      collapsed[i] = sb.occupancy.size() == 1;
      memset(occ + i * cells, 0, cells);
      memset(infl + i * cells, 0, cells);
      memset(lo + i * cells, 0, cells * sizeof(float));
      memcpy(occ + i * cells, sb.occupancy.data(), sb.occupancy.size());
      memcpy(infl + i * cells, sb.inflate_occupancy.data(), sb.inflate_occupancy.size());
      memcpy(lo + i * cells, sb.log_odds.data(), sb.log_odds.size() * sizeof(float));
    }
    i++;
  }
  return i;
}

// frontier sets (exploration mode): per exported subbox (same order as orc_export_map) a bitmask of cells/8 bytes
size_t orc_export_frontier(void *h, size_t cap, uint8_t *bits) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t cells = m->local->cell_num_subbox, nb = (cells + 7) / 8;
  size_t i = 0;
  for (auto &kv : m->local->observed_group_map) {
    if (i < cap) {
      memset(bits + i * nb, 0, nb);
      for (int c : kv.second.frontier) bits[i * nb + (size_t)c / 8] |= (uint8_t)(1u << (c & 7));
    }
    i++;
  }
  return i;
}
// map clouds, reference src/rviz_vis.cpp:267-327 (pub_frontier, pub_global_local_map) and the same loop over
// `occupancy`; points as float x,y,z,1 (pcl::PointXYZ).  kind 0 inflate_occupancy == 'o', 1 occupancy == 'o', 2 frontier
size_t orc_export_cloud(void *h, int kind, float *xyzw, size_t cap) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t n = 0;
  auto emit = [&](const Vec3I &g, int id) {
    Vec3 p = m->local->subbox_id2xyz_glb_vec(g, id);  // same expression as subbox_id2xyz_glb, converted to float by PointP
    if (n < cap) {
      xyzw[4 * n] = (float)p[0];
      xyzw[4 * n + 1] = (float)p[1];
      xyzw[4 * n + 2] = (float)p[2];
      xyzw[4 * n + 3] = 1.0f;
    }
    n++;
  };
  for (auto &kv : m->local->observed_group_map) {
    if (kind == 2) {
      for (int c : kv.second.frontier) emit(kv.first, c);
      continue;
    }
    const auto &v = kind == 0 ? kv.second.inflate_occupancy : kv.second.occupancy;
    int id = 0;
    for (auto it = v.begin(); it != v.end(); it++, id++)
      if (*it == 'o') emit(kv.first, id);
  }
  return n;
}
// odds slice, reference mlmap::visualize_odds src/mlmap.cpp:200-284: x,y,z of the cell and logit_inv(log_odds)
size_t orc_export_odds_slice(void *h, double height, float *xyzw, size_t cap) {
  auto *m = static_cast<orc::mlmap *>(h);
  size_t n = 0;
  for (auto &kv : m->local->observed_group_map) {
    int id = 0;
    for (auto it = kv.second.log_odds.begin(); it != kv.second.log_odds.end(); it++, id++) {
      Vec3 pt = m->local->subbox_id2xyz_glb_vec(kv.first, id);
      if (pt[2] < height + 1e-3 && pt[2] > height - 1e-3) {
        if (n < cap) {
          xyzw[4 * n] = (float)pt[0];
          xyzw[4 * n + 1] = (float)pt[1];
          xyzw[4 * n + 2] = (float)pt[2];
          xyzw[4 * n + 3] = (float)ORC_logit_inv(*it);
        }
        n++;
      }
    }
  }
  return n;
}
int orc_released_last(void *h) { return static_cast<orc::mlmap *>(h)->local->n_released_last; }

// table / scalar probes used by known-answer tests
float orc_odds_table(void *h, int diff, int r) {
  auto *m = static_cast<orc::mlmap *>(h);
  return m->awareness->get_odds_table[diff + m->awareness->diff_range][r];
}
float orc_three_sigma(void *h, int r) { return 3 * static_cast<orc::mlmap *>(h)->awareness->sigma_in_dr(r); }
double orc_fast_atan2(void *h, double y, double x) { return static_cast<orc::mlmap *>(h)->awareness->fast_atan2(y, x); }
float orc_logit(float p) { return ORC_logit(p); }
float orc_logit_inv(float lo) { return ORC_logit_inv(lo); }
void orc_log10f_array(const float *x, size_t n, float *out) {
  for (size_t i = 0; i < n; i++) out[i] = log10(x[i]);  // std::log10(float) == log10f
}
double orc_pow2(double x) { return pow(x, 2); }
int orc_vector_hash(int a, int b, int c) { return orc::awareness_map::VectorHasher()(Vec3I(a, b, c)); }
void orc_transform_point(const double T_wb[7], const double T_bs[7], const double p_s[3], double p_l[3]) {
  // input_pc_pose prologue, map_awareness.cpp:184-186 + :222
  SE3 Twb = se3_from_pose7(T_wb), Tbs = se3_from_pose7(T_bs);
  SE3 T_wa = SE3(SO3(Quat{1, 0, 0, 0}), Twb.translation());
  SE3 T_ws = Twb * Tbs;
  SE3 T_ls = T_wa.inverse() * T_ws;
  Vec3 r = T_ls * Vec3(p_s[0], p_s[1], p_s[2]);
  p_l[0] = r[0];
  p_l[1] = r[1];
  p_l[2] = r[2];
}
void orc_T_ls(const double T_wb[7], const double T_bs[7], double out7[7]) {
  SE3 Twb = se3_from_pose7(T_wb), Tbs = se3_from_pose7(T_bs);
  SE3 T_wa = SE3(SO3(Quat{1, 0, 0, 0}), Twb.translation());
  SE3 T_ls = T_wa.inverse() * (Twb * Tbs);
  out7[0] = T_ls.t[0];
  out7[1] = T_ls.t[1];
  out7[2] = T_ls.t[2];
  out7[3] = T_ls.so3.q.w;
  out7[4] = T_ls.so3.q.x;
  out7[5] = T_ls.so3.q.y;
  out7[6] = T_ls.so3.q.z;
}
void orc_compensate_pose(const double pos[3], const double quat_wxyz[4], const double lin_vel[3], const double ang_vel[3],
                         double gap_odom, double gap_imu, double latency, double out7[7]) {
  SE3 T = compensate_pose(Vec3(pos[0], pos[1], pos[2]), Quat{quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]},
                          Vec3(lin_vel[0], lin_vel[1], lin_vel[2]), Vec3(ang_vel[0], ang_vel[1], ang_vel[2]), gap_odom, gap_imu,
                          latency);
  out7[0] = T.t[0];
  out7[1] = T.t[1];
  out7[2] = T.t[2];
  out7[3] = T.so3.q.w;
  out7[4] = T.so3.q.x;
  out7[5] = T.so3.q.y;
  out7[6] = T.so3.q.z;
}
// libstdc++ growth chain probe: bucket_count of an empty unordered_set after rehash(n)
size_t orc_next_bucket_count(size_t n) {
  std::unordered_set<size_t> s;
  s.rehash(n);
  return s.bucket_count();
}

}  // extern "C"
