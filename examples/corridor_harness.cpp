// ROS-free harness in C++ over the reference-shaped façade (mlmapping_b200/include/mlmap.hpp):
// integrates a few synthetic depth frames of a box corridor and runs the planner-facing queries.
// Build: see __graft_entry__.build() (g++ -std=c++17 examples/corridor_harness.cpp -Lmlmapping_b200/lib -lmlmap_b200)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../mlmapping_b200/include/mlmap.hpp"

using namespace mlmap_b200;

// depth image of a 60 x 2.4 x 2.4 m corridor seen from (x,0,1.2) looking along +x (camera level)
static void make_frame(std::vector<uint16_t> &img, int rows, int cols, float cx, float cy, float fx, float fy, double x0) {
  img.resize((size_t)rows * cols);
  for (int v = 0; v < rows; v++)
    for (int u = 0; u < cols; u++) {
      // optical frame: x right, y down, z forward  ->  world: forward +x, right -y, down -z
      double dx = 1.0, dy = -(u - cx) / fx, dz = -(v - cy) / fy;
      double t = (60.0 - (x0 + 0.12)) / dx;
      if (dy > 0) t = std::fmin(t, 1.2 / dy);
      if (dy < 0) t = std::fmin(t, -1.2 / dy);
      if (dz > 0) t = std::fmin(t, 1.2 / dz);
      if (dz < 0) t = std::fmin(t, -1.2 / dz);
      double mm = std::round(t * 1000.0);
      img[(size_t)v * cols + u] = mm > 65535 ? 0 : (uint16_t)mm;
    }
}

int main() {
  mlm_config cfg = mlmap::default_config();
  // CFG-A of SURVEY §8d: 0.1 m cells, 65 x 360 x 41 awareness grid
  cfg.am_d_rho = 0.1, cfg.am_d_phi_deg = 1.0, cfg.am_d_z = 0.1;
  cfg.am_n_rho = 65, cfg.am_n_z_below = 20, cfg.am_n_z_over = 20;
  cfg.depth_noise_coe = 0.00375, cfg.subbox_d_xyz = 0.1;
  cfg.cam_cx = 320, cfg.cam_cy = 240;
  cfg.max_points = 640 * 480, cfg.pool_submaps = 8192;
  mlmap map;
  try {
    map.init_map(cfg);
  } catch (const std::exception &e) {
    std::printf("init_map failed: %s\n", e.what());
    return 2;
  }
  std::vector<uint16_t> img;
  double total_ms = 0;
  const int frames = 20;
  for (int k = 0; k < frames; k++) {
    const double x = 5.0 + 0.05 * k;
    make_frame(img, 480, 640, cfg.cam_cx, cfg.cam_cy, cfg.cam_fx, cfg.cam_fy, x);
    auto t0 = std::chrono::steady_clock::now();
    const mlm_frame_stats &st = map.depth_odom_input(img.data(), 480, 640, 640 * 2, SE3(1, 0, 0, 0, Vec3(x, 0, 1.2)));
    total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (k % 5 == 0)
      std::printf("frame %2d: rays %d hit cells %d miss cells %d touched voxels %d submaps %lld\n", k, st.n_points,
                  st.n_hit_cells, st.n_miss_cells, st.n_touched_voxels, (long long)st.ram_expand_cnt);
  }
  std::printf("[mlmapping] ave-time cost: %.4f ms per frame (host wall clock, H2D included)\n", total_ms / frames);
  const Vec3 probe(7.0, 0.0, 1.2), wall(7.0, 1.22, 1.2), far(100, 0, 0);
  std::printf("getOccupancy free=%d wall=%d unknown=%d  getOdd wall=%.4f\n", map.getOccupancy(probe), map.getOccupancy(wall),
              map.getOccupancy(far), map.getOdd(wall));
  Vec3 g = map.getOddGrad(Vec3(7.0, 1.12, 1.2));
  std::printf("getOddGrad near wall = (%.4f, %.4f, %.4f)\n", g[0], g[1], g[2]);
  // what the reference publishes after an update (occupied-cell cloud, odds slice) and a checkpoint round trip
  const auto cloud = map.map_cloud(MLM_CLOUD_OCCUPIED);
  const auto slice = map.odds_slice(1.25);
  const auto image = map.checkpoint();
  mlmap copy;
  copy.init_map(cfg);
  copy.restore(image);
  std::printf("occupied cloud %zu points, odds slice %zu cells, checkpoint %zu bytes, restored copy: wall=%d (%zu points)\n",
              cloud.size(), slice.size(), image.size(), copy.getOccupancy(wall), copy.map_cloud(MLM_CLOUD_OCCUPIED).size());
  // the layer-level calls of update_map (src/mlmap.cpp:382-386) and the index overload of getOdd
  std::vector<Vec3> pc;
  for (int i = 0; i < 2000; i++) pc.emplace_back(0.002 * (i % 50) - 0.05, 0.002 * (i / 50) - 0.04, 1.0 + 0.0005 * i);
  copy.awareness_map->input_pc_pose(pc, SE3(1, 0, 0, 0, Vec3(6.0, 0, 1.2)));
  const mlm_frame_stats &st2 = copy.local_map->input_pc_pose_direct(copy.awareness_map);
  const float odd_idx = map.getOdd(Vec3I(7, 1, 1), 2 + 10 * 2 + 100 * 2);   // cell (2,2,2) of subbox (7,1,1): centre (7.25, 1.25, 1.25)
  const float odd_pos = map.getOdd(Vec3(7.25, 1.25, 1.25));
  std::printf("two-call update: %d points, %d hit cells; getOdd by index %.6f == by position %.6f\n", st2.n_points, st2.n_hit_cells,
              odd_idx, odd_pos);
  // self-checks: the C++ mirror must behave on a device, not just link
  int bad = 0;
  bad += map.getOccupancy(probe) != mlmap::FREE;
  bad += map.getOccupancy(wall) != mlmap::OCCUPIED;
  bad += map.getOccupancy(far) != mlmap::UNKNOWN;
  bad += !(map.getOdd(wall) > 0.99f);
  bad += copy.getOccupancy(wall) != mlmap::OCCUPIED;
  bad += cloud.empty() || slice.empty();
  bad += st2.n_points != 2000 || st2.n_hit_cells <= 0;
  bad += odd_idx != odd_pos;
  std::printf("harness self-checks: %s\n", bad ? "FAILED" : "ok");
  return bad ? 1 : 0;
}
