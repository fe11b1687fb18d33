// =================================================================================================
// TEST INFRASTRUCTURE ONLY — driver of oracle/_ref/libmlmap_ref.so.
//
// This translation unit is the only code of ours inside that library's mapping path: everything it calls
// is the UNMODIFIED reference, compiled from the sources where they lie under /root/reference
// (src/map_awareness.cpp, src/map_local.cpp, src/mlmap.cpp, src/rviz_vis.cpp, include/*.h, the vendored
// 3rdPartLib/Sophus/sophus/{so3,se3}.cpp and 3rdPartLib/yaml-cpp-0.6.2/src/*.cpp) against the inert
// ROS/PCL/OpenCV stand-ins and the Eigen subset of oracle/ref_build/shim/ (see mini_eigen.hpp for what
// that means for pinning).  It exports the same orc_* C surface as oracle/oracle_capi.cpp, so
// tests/oracle_binding.py can drive either the restatement or the reference itself.
//
// What is NOT the reference here (harness additions, each marked below):
//   * full-frame projection (every non-zero pixel, row-major) — the reference only samples pixels with rand();
//     the per-pixel expression is the one of mlmap::project_depth (src/mlmap.cpp:329-347) on the class's own members;
//   * the frame counters n_inside / n_cast / n_touched_voxels / released, recomputed after the update by calling
//     the reference's own xyz2RhoPhiZwithBoderCheck / get_global_idx on the containers it left behind;
//   * the "occupied cells" cloud (kind 1): the loop of rviz_vis::pub_global_local_map over `occupancy`.
// =================================================================================================
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <typeindex>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <fcntl.h>
#include <unistd.h>

#include "shim/ref_shim_ros.hpp"
#include "shim/mini_eigen/mini_eigen.hpp"
#include "../3rdPartLib/yaml-cpp-0.6.2/include/yaml-cpp/yaml.h"

// the driver reads / sets private members of the reference classes (pc_eigen, T_wb, ct_pos, depth_image_, cx_ ...)
// exactly as depth_odom_input_callback does; the class layout does not depend on access specifiers
#define private public
#define protected public
#include <mlmap.h>
#undef private
#undef protected

#include "../../include/mlmap_b200.h"

// mlmap::tf_timerCb (src/mlmap.cpp:440-451) mentions the two legacy publishers (2-D grid, ESDF; dead modules, SURVEY §2
// #11/#12, their .cpp files are not part of this build); the timer never fires here, the linker just needs the symbols
void Local2OccupancyGrid2D::pub_occupancy_grid_2D_from_localmap(local_map_cartesian *, ros::Time) {}
void Local2ESDFsBatch::pub_ESDF_3D_from_localmap(local_map_cartesian *, ros::Time) {}

namespace {

struct RefMap {
  mlmap *m = nullptr;
  mlm_config cfg;
  bool bookkeeping = true;
  size_t n_inside = 0, n_cast = 0, n_touched = 0;
  int n_released = 0;
  std::vector<uint16_t> img;  // dense copy for sampled mode
};

SE3 se3_from_pose7(const double p[7]) { return SE3(SO3(Quaterniond(p[3], p[4], p[5], p[6])), Vec3(p[0], p[1], p[2])); }

std::string fmt(double v) {
  char b[64];
  snprintf(b, sizeof(b), "%.17g", v);
  std::string s(b);
  if (s.find_first_of(".eEni") == std::string::npos) s += ".0";
  return s;
}

// the YAML file mlmap::init_map reads (src/mlmap.cpp:9-85,102)
std::string write_yaml(const mlm_config &c) {
  static int serial = 0;
  char path[256];
  snprintf(path, sizeof(path), "/tmp/mlmap_ref_%d_%d.yaml", (int)getpid(), serial++);
  Matrix3d Rm = Quaterniond(c.T_bs[3], c.T_bs[4], c.T_bs[5], c.T_bs[6]).normalized().toRotationMatrix();
  std::ofstream f(path);
  f << "mlmapping_am_d_Rho: " << fmt(c.am_d_rho) << "\n";
  f << "mlmapping_am_d_Phi_deg: " << fmt(c.am_d_phi_deg) << "\n";
  f << "mlmapping_am_d_Z: " << fmt(c.am_d_z) << "\n";
  f << "mlmapping_am_n_Rho: " << c.am_n_rho << "\n";
  f << "mlmapping_am_n_Z_below: " << c.am_n_z_below << "\n";
  f << "mlmapping_am_n_Z_over: " << c.am_n_z_over << "\n";
  f << "mlmapping_subbox_d_xyz: " << fmt(c.subbox_d_xyz) << "\n";
  f << "mlmapping_subbox_n: " << c.subbox_n << "\n";
  f << "mlmapping_lm_log_odds_min: " << fmt(c.log_odds_min) << "\n";
  f << "mlmapping_lm_log_odds_max: " << fmt(c.log_odds_max) << "\n";
  f << "mlmapping_lm_measurement_hit: " << fmt(c.log_odds_hit) << "\n";
  f << "mlmapping_lm_measurement_miss: " << fmt(c.log_odds_miss) << "\n";
  f << "mlmapping_lm_occupied_sh: " << fmt(c.log_odds_occupied_sh) << "\n";
  f << "mlmapping_cam_cx: " << fmt(c.cam_cx) << "\n";
  f << "mlmapping_cam_cy: " << fmt(c.cam_cy) << "\n";
  f << "mlmapping_cam_fx: " << fmt(c.cam_fx) << "\n";
  f << "mlmapping_cam_fy: " << fmt(c.cam_fy) << "\n";
  f << "mlmapping_depth_noise_coe: " << fmt(c.depth_noise_coe) << "\n";
  f << "camera2odom_latency: 0.0\n";
  f << "use_exploration_frontiers: " << (c.use_exploration_frontiers ? "true" : "false") << "\n";
  f << "visualize_odds: false\n";
  f << "mlmapping_sample_cnt: " << c.sample_cnt << "\n";
  f << "mlmapping_inflate_n: " << c.inflate_n << "\n";
  f << "mlmapping_inflate_global_n: " << c.inflate_global_n << "\n";
  f << "mlmapping_apply_inflate: " << (c.apply_inflate ? "true" : "false") << "\n";
  f << "mlmapping_use_raycasting: " << (c.use_raycasting ? "true" : "false") << "\n";
  f << "publish_T_wb: false\npublish_T_bs: false\n";
  f << "sensor_frame_id: \"sensor\"\nbody_frame_id: \"body\"\nawareness_frame_id: \"awareness\"\n";
  f << "local_frame_id: \"local\"\nworld_frame_id: \"map\"\nvisulize_raycasting: false\n";
  f << "T_B_S: [";
  for (int i = 0; i < 3; i++) f << fmt(Rm(i, 0)) << ", " << fmt(Rm(i, 1)) << ", " << fmt(Rm(i, 2)) << ", " << fmt(c.T_bs[i]) << ", ";
  f << "0.0, 0.0, 0.0, 1.0]\n";
  f.close();
  return path;
}

RefMap *create(const mlm_config &c) {
  std::cout.setstate(std::ios_base::failbit);  // the reference prints per un-castable point (map_awareness.cpp:277-278)
  RefMap *r = new RefMap();
  r->cfg = c;
  std::string path = write_yaml(c);
  ref_shim::params()["/mlmapping_configfile"] = path;
  r->m = new mlmap();
  ros::NodeHandle nh;
  r->m->init_map(nh);  // the reference's own initialisation: awareness tables, local tables, float casts of the yaml values
  unlink(path.c_str());
  // T_B_S: the configuration struct carries it as a pose (quaternion); hand exactly that to the public setter
  // (awareness_map_cylindrical::setTbs, map_awareness.cpp:14-17) instead of the matrix -> quaternion conversion
  r->m->awareness_map->setTbs(se3_from_pose7(c.T_bs));
  r->m->local_map->flate_height = c.inflate_height;  // public member, default 0.1 (map_local.h:65)
  return r;
}

// harness counters, recomputed with the reference's own functions from what the frame left behind
void bookkeeping_before(RefMap *r, std::set<std::array<int, 3>> &collapsed) {
  collapsed.clear();
  if (!r->bookkeeping) return;
  for (auto &kv : r->m->local_map->observed_group_map)
    if (kv.second.occupancy.size() == 1) collapsed.insert({kv.first[0], kv.first[1], kv.first[2]});
}
void bookkeeping_after(RefMap *r, const std::set<std::array<int, 3>> &collapsed_before) {
  if (!r->bookkeeping) return;
  mlmap *m = r->m;
  awareness_map_cylindrical *a = m->awareness_map;
  local_map_cartesian *l = m->local_map;
  // input_pc_pose prologue (map_awareness.cpp:184-186)
  SE3 T_wa = SE3(SO3(Quaterniond(1, 0, 0, 0)), m->T_wb.translation());
  SE3 T_ws = m->T_wb * a->T_bs;
  SE3 T_ls = T_wa.inverse() * T_ws;
  r->n_inside = r->n_cast = 0;
  for (auto &p_s : m->pc_eigen) {
    Vec3I rpz;
    bool can;
    bool inside = a->xyz2RhoPhiZwithBoderCheck(T_ls * p_s, rpz, can);
    if (inside) r->n_inside++;
    if (can && a->visibility_check) r->n_cast++;
  }
  std::set<std::array<long, 4>> touched;
  auto touch = [&](const Vec3 &p_w) {
    Vec3I g;
    size_t s;
    l->get_global_idx(p_w, g, s);
    if (collapsed_before.count({g[0], g[1], g[2]})) return;  // allocate_ram returned false for it during the frame
    touched.insert({g[0], g[1], g[2], (long)s});
  };
  for (auto &kv : a->hit_idx_odds_hashmap) touch(a->T_wa * a->map->at(a->mapIdx(kv.first)).center_pt);
  for (auto idx : a->miss_idx_set) touch(a->T_wa * a->map->at(idx).center_pt);
  r->n_touched = touched.size();
  int now_collapsed = 0;
  for (auto &kv : l->observed_group_map)
    if (kv.second.occupancy.size() == 1) now_collapsed++;
  r->n_released = now_collapsed - (int)collapsed_before.size();
}

double run_update(RefMap *r, const double T_wb[7]) {
  mlmap *m = r->m;
  m->T_wb = se3_from_pose7(T_wb);
  m->ct_pos = Vec3(T_wb[0], T_wb[1], T_wb[2]);
  std::set<std::array<int, 3>> collapsed;
  bookkeeping_before(r, collapsed);
  auto t0 = std::chrono::steady_clock::now();
  m->update_map();  // src/mlmap.cpp:382-386: input_pc_pose + input_pc_pose_direct
  auto t1 = std::chrono::steady_clock::now();
  bookkeeping_after(r, collapsed);
  return std::chrono::duration<double>(t1 - t0).count();
}

RefMap *scratch_map() {  // for the handle-free probes
  static RefMap *s = nullptr;
  if (!s) {
    mlm_config c;
    memset(&c, 0, sizeof(c));
    c.am_d_rho = 0.1; c.am_d_phi_deg = 45; c.am_d_z = 0.1; c.am_n_rho = 4; c.am_n_z_below = 1; c.am_n_z_over = 1;
    c.use_raycasting = 1; c.depth_noise_coe = 0.00375; c.subbox_d_xyz = 0.1; c.subbox_n = 2;
    c.log_odds_min = -2.f; c.log_odds_max = 4.2f; c.log_odds_hit = 0.7f; c.log_odds_miss = -0.9f; c.log_odds_occupied_sh = 3.f;
    c.cam_cx = c.cam_cy = 1; c.cam_fx = c.cam_fy = 1;
    c.T_bs[3] = 1; c.inflate_n = 1; c.inflate_global_n = 1; c.inflate_height = 0.1;
    s = create(c);
  }
  return s;
}

void out7(const SE3 &T, double o[7]) {
  o[0] = T.translation()[0];
  o[1] = T.translation()[1];
  o[2] = T.translation()[2];
  o[3] = T.unit_quaternion().w();
  o[4] = T.unit_quaternion().x();
  o[5] = T.unit_quaternion().y();
  o[6] = T.unit_quaternion().z();
}

}  // namespace

extern "C" {

const char *orc_kind(void) { return "reference"; }

void *orc_create(const mlm_config *cfg) { return create(*cfg); }
void orc_destroy(void *h) {
  RefMap *r = static_cast<RefMap *>(h);
  if (!r) return;
  delete r->m->awareness_map;
  delete r->m->local_map;
  // the reference never frees its publishers (src/mlmap.cpp:125-133); neither is mlmap itself destructible twice
  delete r->m;
  delete r;
}
void orc_set_bookkeeping(void *h, int on) { static_cast<RefMap *>(h)->bookkeeping = on != 0; }

double orc_integrate_depth_u16(void *h, const uint16_t *img, int rows, int cols, size_t stride_bytes, const double T_wb[7]) {
  RefMap *r = static_cast<RefMap *>(h);
  mlmap *m = r->m;
  m->pc_eigen.clear();  // depth_odom_input_callback, src/mlmap.cpp:470
  double t_proj;
  if (m->pc_sample_cnt > 0) {
    // the reference's own sampled projection on its own cv::Mat member (src/mlmap.cpp:311-349), libc rand()
    m->depth_image_.create(rows, cols, CV_16UC1);
    for (int v = 0; v < rows; v++)
      memcpy(m->depth_image_.ptr<uint16_t>(v), reinterpret_cast<const uint8_t *>(img) + (size_t)v * stride_bytes, (size_t)cols * 2);
    auto t0 = std::chrono::steady_clock::now();
    m->project_depth();
    t_proj = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } else {
    // HARNESS ADDITION: full-frame mode, the per-pixel statements of src/mlmap.cpp:329-347 for every pixel, v outer / u inner
    auto t0 = std::chrono::steady_clock::now();
    uint16_t *row_ptr;
    size_t u, v;
    double depth;
    Vec3 pt_cur;
    for (v = 0; v < (size_t)rows; v++)
      for (u = 0; u < (size_t)cols; u++) {
        row_ptr = reinterpret_cast<uint16_t *>(const_cast<uint8_t *>(reinterpret_cast<const uint8_t *>(img)) + v * stride_bytes) + u;
        depth = (*row_ptr) * m->inv_factor;
        if (*row_ptr == 0) continue;
        pt_cur(0) = (u - m->cx_) * depth / m->fx_;
        pt_cur(1) = (v - m->cy_) * depth / m->fy_;
        pt_cur(2) = depth;
        m->pc_eigen.emplace_back(pt_cur);
      }
    t_proj = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return t_proj + run_update(r, T_wb);
}

double orc_integrate_points_f64(void *h, const double *xyz, int n, const double T_wb[7]) {
  RefMap *r = static_cast<RefMap *>(h);
  mlmap *m = r->m;
  m->pc_eigen.clear();
  m->pc_eigen.reserve(n);
  for (int i = 0; i < n; i++) m->pc_eigen.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  return run_update(r, T_wb);
}

void orc_frame_stats(void *h, mlm_frame_stats *s) {
  RefMap *r = static_cast<RefMap *>(h);
  mlmap *m = r->m;
  memset(s, 0, sizeof(*s));
  s->n_points = (int32_t)m->pc_eigen.size();
  s->n_inside = (int32_t)r->n_inside;
  s->n_cast = (int32_t)r->n_cast;
  s->n_hit_cells = (int32_t)m->awareness_map->hit_idx_odds_hashmap.size();
  s->n_miss_cells = (int32_t)m->awareness_map->miss_idx_set.size();
  s->n_touched_voxels = (int32_t)r->n_touched;
  s->hit_bucket_count = (int32_t)m->awareness_map->hit_idx_odds_hashmap.bucket_count();
  s->ram_expand_cnt = m->local_map->ram_expand_cnt;
  s->obs_cnt = m->local_map->obs_cnt;
}

size_t orc_num_points(void *h) { return static_cast<RefMap *>(h)->m->pc_eigen.size(); }
size_t orc_get_points(void *h, double *xyz, size_t cap) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  size_t n = std::min(cap, m->pc_eigen.size());
  for (size_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) xyz[3 * i + k] = m->pc_eigen[i][k];
  return m->pc_eigen.size();
}
size_t orc_last_hits(void *h, int32_t *keys3, float *p, size_t cap) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  size_t i = 0;
  for (auto &kv : m->awareness_map->hit_idx_odds_hashmap) {
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
  mlmap *m = static_cast<RefMap *>(h)->m;
  size_t i = 0;
  for (auto v : m->awareness_map->miss_idx_set) {
    if (i < cap) idx[i] = v;
    i++;
  }
  return i;
}
void orc_set_log_inserts(void *, int) {}  // instrumentation of the restatement only
size_t orc_insert_log(void *, int32_t *, size_t) { return 0; }

void orc_set_free_in_bound(void *h, const double mn[3], const double mx[3]) {
  static_cast<RefMap *>(h)->m->setFree_map_in_bound(Vec3(mn[0], mn[1], mn[2]), Vec3(mx[0], mx[1], mx[2]));
}
void orc_inflate_map(void *h, const double ct[3]) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  m->ct_pos = Vec3(ct[0], ct[1], ct[2]);
  m->inflate_map();
}
void orc_get_occupancy(void *h, const double *pos, size_t n, int32_t *out) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  for (size_t i = 0; i < n; i++) out[i] = m->getOccupancy(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void orc_get_occupancy_inflate(void *h, const double *pos, size_t n, float inflate, int32_t *out) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  for (size_t i = 0; i < n; i++) out[i] = m->getOccupancy(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), inflate);
}
void orc_get_inflate_occupancy(void *h, const double *pos, size_t n, int32_t *out) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  for (size_t i = 0; i < n; i++) out[i] = m->getInflateOccupancy(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void orc_get_odd(void *h, const double *pos, size_t n, float *out) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  for (size_t i = 0; i < n; i++) out[i] = m->getOdd(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void orc_get_odd_at(void *h, const int32_t *glb3, const int32_t *sub, size_t n, float *out) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  for (size_t i = 0; i < n; i++) out[i] = m->getOdd(Vec3I(glb3[3 * i], glb3[3 * i + 1], glb3[3 * i + 2]), (size_t)sub[i]);
}
void orc_get_odd_grad(void *h, const double *pos, size_t n, size_t max_iter, double *out) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  for (size_t i = 0; i < n; i++) {
    Vec3 g = m->getOddGrad(Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), max_iter);
    out[3 * i] = g[0];
    out[3 * i + 1] = g[1];
    out[3 * i + 2] = g[2];
  }
}

size_t orc_export_map_count(void *h) { return static_cast<RefMap *>(h)->m->local_map->observed_group_map.size(); }
size_t orc_export_map(void *h, size_t cap, int32_t *glb3, uint8_t *collapsed, char *occ, char *infl, float *lo) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  size_t cells = m->local_map->cell_num_subbox;
  size_t i = 0;
  for (auto &kv : m->local_map->observed_group_map) {
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
size_t orc_export_frontier(void *h, size_t cap, uint8_t *bits) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  size_t cells = m->local_map->cell_num_subbox, nb = (cells + 7) / 8;
  size_t i = 0;
  for (auto &kv : m->local_map->observed_group_map) {
    if (i < cap) {
      memset(bits + i * nb, 0, nb);
      for (int c : kv.second.frontier) bits[i * nb + (size_t)c / 8] |= (uint8_t)(1u << (c & 7));
    }
    i++;
  }
  return i;
}
// kind 0: what rviz_vis::pub_global_local_map publishes on /global_map (src/rviz_vis.cpp:296-327), read back from the
// captured PointCloud2; kind 2: rviz_vis::pub_frontier on /frontier (:267-294); kind 1: harness loop over `occupancy`
size_t orc_export_cloud(void *h, int kind, float *xyzw, size_t cap) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  if (kind == 1) {
    size_t n = 0;
    for (auto &kv : m->local_map->observed_group_map) {
      int id = 0;
      for (auto it = kv.second.occupancy.begin(); it != kv.second.occupancy.end(); it++, id++)
        if (*it == 'o') {
          PointP p = m->local_map->subbox_id2xyz_glb(kv.first, id);
          if (n < cap) {
            xyzw[4 * n] = p.x;
            xyzw[4 * n + 1] = p.y;
            xyzw[4 * n + 2] = p.z;
            xyzw[4 * n + 3] = 1.0f;
          }
          n++;
        }
    }
    return n;
  }
  const char *topic = kind == 0 ? "/global_map" : "/frontier";
  ref_shim::topics().erase(topic);
  if (kind == 0) m->visualize_map(); else m->visualize_frontier();
  const sensor_msgs::PointCloud2 *msg = ref_shim::last<sensor_msgs::PointCloud2>(topic);
  if (!msg) return 0;
  size_t n = msg->data.size() / msg->point_step;
  for (size_t i = 0; i < n && i < cap; i++) memcpy(xyzw + 4 * i, msg->data.data() + i * msg->point_step, 16);
  return n;
}
// mlmap::visualize_odds (src/mlmap.cpp:200-284) publishes the slice's cell centres (and a colour derived from the odd)
// on /odds; the points come from that message, the odd is logit_inv of the same cell in the same iteration order
size_t orc_export_odds_slice(void *h, double height, float *xyzw, size_t cap) {
  mlmap *m = static_cast<RefMap *>(h)->m;
  ref_shim::topics().erase("/odds");
  m->visualize_odds((float)height);
  const visualization_msgs::MarkerArray *msg = ref_shim::last<visualization_msgs::MarkerArray>("/odds");
  std::vector<float> odd;
  const float hf = (float)height;  // visualize_odds takes a float height
  for (auto &kv : m->local_map->observed_group_map) {
    int id = 0;
    for (auto it = kv.second.log_odds.begin(); it != kv.second.log_odds.end(); it++, id++) {
      Vec3 pt = m->local_map->subbox_id2xyz_glb_vec(kv.first, id);
      if (pt[2] < hf + 1e-3 && pt[2] > hf - 1e-3) odd.push_back((float)logit_inv(*it));
    }
  }
  if (!msg) return 0;
  const auto &pts = msg->markers[0].points;
  if (pts.size() != odd.size()) return (size_t)-1;
  for (size_t i = 0; i < pts.size() && i < cap; i++) {
    xyzw[4 * i] = (float)pts[i].x;
    xyzw[4 * i + 1] = (float)pts[i].y;
    xyzw[4 * i + 2] = (float)pts[i].z;
    xyzw[4 * i + 3] = odd[i];
  }
  return pts.size();
}
int orc_released_last(void *h) { return static_cast<RefMap *>(h)->n_released; }

// table / scalar probes used by known-answer tests (private members of the reference classes)
float orc_odds_table(void *h, int diff, int r) {
  awareness_map_cylindrical *a = static_cast<RefMap *>(h)->m->awareness_map;
  return a->get_odds_table[diff + a->diff_range][r];
}
float orc_three_sigma(void *h, int r) { return 3 * static_cast<RefMap *>(h)->m->awareness_map->sigma_in_dr(r); }
double orc_fast_atan2(void *h, double y, double x) { return static_cast<RefMap *>(h)->m->awareness_map->fast_atan2(y, x); }
float orc_logit(float p) { return logit(p); }
float orc_logit_inv(float lo) { return logit_inv(lo); }
void orc_log10f_array(const float *x, size_t n, float *out) {
  for (size_t i = 0; i < n; i++) out[i] = log10(x[i]);
}
double orc_pow2(double x) { return pow(x, 2); }
int orc_vector_hash(int a, int b, int c) { return awareness_map_cylindrical::VectorHasher()(Vec3I(a, b, c)); }
void orc_transform_point(const double T_wb[7], const double T_bs[7], const double p_s[3], double p_l[3]) {
  SE3 Twb = se3_from_pose7(T_wb), Tbs = se3_from_pose7(T_bs);
  SE3 T_wa = SE3(SO3(Quaterniond(1, 0, 0, 0)), Twb.translation());
  SE3 T_ws = Twb * Tbs;
  SE3 T_ls = T_wa.inverse() * T_ws;
  Vec3 r = T_ls * Vec3(p_s[0], p_s[1], p_s[2]);
  p_l[0] = r[0];
  p_l[1] = r[1];
  p_l[2] = r[2];
}
void orc_T_ls(const double T_wb[7], const double T_bs[7], double o[7]) {
  SE3 Twb = se3_from_pose7(T_wb), Tbs = se3_from_pose7(T_bs);
  SE3 T_wa = SE3(SO3(Quaterniond(1, 0, 0, 0)), Twb.translation());
  SE3 T_ls = T_wa.inverse() * (Twb * Tbs);
  out7(T_ls, o);
}
// the pose forwarding of mlmap::depth_odom_input_callback (src/mlmap.cpp:462-498), by calling that callback on a scratch
// map with an all-zero 1x1 image (nothing to integrate) and reading the T_wb it computed
void orc_compensate_pose(const double pos[3], const double quat_wxyz[4], const double lin_vel[3], const double ang_vel[3],
                         double gap_odom, double gap_imu, double latency, double o[7]) {
  mlmap *m = scratch_map()->m;
  auto img = std::make_shared<sensor_msgs::Image>();
  img->height = img->width = 1;
  img->step = 2;
  img->encoding = sensor_msgs::image_encodings::TYPE_16UC1;
  img->data.assign(2, 0);
  img->header.stamp = ros::Time(0.0);  // stamps chosen so that the callback's stamp differences are exactly gap_odom / gap_imu
  auto odom = std::make_shared<nav_msgs::Odometry>();
  odom->header.stamp = ros::Time(-gap_odom);
  odom->pose.pose.position.x = pos[0];
  odom->pose.pose.position.y = pos[1];
  odom->pose.pose.position.z = pos[2];
  odom->pose.pose.orientation.w = quat_wxyz[0];
  odom->pose.pose.orientation.x = quat_wxyz[1];
  odom->pose.pose.orientation.y = quat_wxyz[2];
  odom->pose.pose.orientation.z = quat_wxyz[3];
  odom->twist.twist.linear.x = lin_vel[0];
  odom->twist.twist.linear.y = lin_vel[1];
  odom->twist.twist.linear.z = lin_vel[2];
  auto imu = std::make_shared<sensor_msgs::Imu>();
  imu->header.stamp = ros::Time(-gap_imu);
  imu->angular_velocity.x = ang_vel[0];
  imu->angular_velocity.y = ang_vel[1];
  imu->angular_velocity.z = ang_vel[2];
  m->camera2odom_latency = latency;
  m->pc_sample_cnt = 0;  // project_depth's loop does not run: libc's rand() stream stays untouched
  // the callback printf()s its timing every 10th call (src/mlmap.cpp:524-525): keep that off the test output
  fflush(stdout);
  int saved = dup(1), devnull = open("/dev/null", O_WRONLY);
  dup2(devnull, 1);
  m->depth_odom_input_callback(img, odom, imu);
  fflush(stdout);
  dup2(saved, 1);
  close(saved);
  close(devnull);
  out7(m->T_wb, o);
}
size_t orc_next_bucket_count(size_t n) {
  std::unordered_set<size_t> s;
  s.rehash(n);
  return s.bucket_count();
}

}  // extern "C"
