// =================================================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement of the MLMapping per-frame map-update path and its point queries, following
// the reference (/root/reference) function by function.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call this code; the
// product (mlmapping_b200/csrc) never links or includes anything from oracle/.
//
// PINNED AGAINST THE REFERENCE ITSELF: the reference ships no tests, golden vectors or fixtures for any mapping
// function, but its mapping sources compile unmodified against a small Eigen subset and inert
// ROS/PCL/OpenCV stand-ins (oracle/ref_build -> oracle/_ref/libmlmap_ref.so).  tests/test_reference_pin.py and the
// known-answer tests run this restatement and that library on the same seeded inputs and require identical bits
// (hit map in iteration order, miss set in iteration order, map, queries, clouds, forwarded pose);
// tests/golden/golden.json is generated from that library.  This restatement keeps the reference's own containers
// (std::unordered_map / std::unordered_set with the reference's hashers, same insert/clear sequence) so libstdc++
// iteration order is inherited, and is compiled with the reference's flags (-std=c++17 -O3, no -march, no
// fast-math; reference CMakeLists.txt:4).
//
// What stays a restatement on BOTH sides is Eigen3 (the reference's un-vendored, unpinned system dependency, absent
// from this image): Quaternion product / normalize / conjugate / _transformVector and fixed-size vector +,-,* follow
// Eigen 3.3's evaluation order for an x86-64 SSE2 build (see the notes at quat_mul below and in
// oracle/ref_build/shim/mini_eigen/mini_eigen.hpp).
// =================================================================================================
#ifndef MLMAP_ORACLE_HPP
#define MLMAP_ORACLE_HPP

#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <array>
#include <algorithm>
#include <functional>
#include <unordered_map>
#include <unordered_set>

#include "../include/mlmap_b200.h"

namespace orc {
using namespace std;  // the reference has `using namespace std` (include/common.h:17): overloads resolve alike

// ---- POD stand-ins for Eigen::Matrix<double,3,1>, Matrix<int,3,1> (include/common.h:22-23) -------
struct Vec3 {
  double v[3];
  Vec3() : v{0, 0, 0} {}
  Vec3(double a, double b, double c) : v{a, b, c} {}
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return Vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return Vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vec3 operator*(const Vec3 &a, double s) { return Vec3(a[0] * s, a[1] * s, a[2] * s); }

struct Vec3I {
  int v[3];
  Vec3I() : v{0, 0, 0} {}
  Vec3I(int a, int b, int c) : v{a, b, c} {}
  // Eigen Matrix<int,3,1>(double,double,double) converts each coefficient to int
  Vec3I(double a, double b, double c) : v{static_cast<int>(a), static_cast<int>(b), static_cast<int>(c)} {}
  int &operator()(int i) { return v[i]; }
  int operator()(int i) const { return v[i]; }
  int &operator[](int i) { return v[i]; }
  int operator[](int i) const { return v[i]; }
  int size() const { return 3; }
  bool operator==(const Vec3I &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};
inline Vec3I operator+(const Vec3I &a, const Vec3I &b) { return Vec3I(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }

// ---- Eigen::Quaterniond subset ---------------------------------------------------------------------
struct Quat {
  double w, x, y, z;
};
// Eigen evaluates these as an x86-64 SSE2 build does (the reference's flags: -O3, no -march; packets of 2 doubles):
// the quaternion product is the kernel of Eigen/src/Geometry/arch/Geometry_SSE.h, a 4-coefficient squaredNorm
// adds the two packets first, (x^2 + z^2) + (y^2 + w^2), a 3-coefficient one is (a0^2 + a1^2) + a2^2.
// oracle/ref_build/shim/mini_eigen/mini_eigen.hpp states the same rules for the build of the reference's own sources.
inline Quat quat_mul(const Quat &a, const Quat &b) {
  Quat r;
  r.x = (a.w * b.x + a.y * b.z) - (a.z * b.y - a.x * b.w);
  r.y = (a.w * b.y + a.y * b.w) + (a.z * b.x - a.x * b.z);
  r.z = (a.w * b.z - a.y * b.x) + (a.z * b.w + a.x * b.y);
  r.w = (a.w * b.w - a.y * b.y) - (a.z * b.z + a.x * b.x);
  return r;
}
inline Quat quat_normalized(const Quat &q) {  // MatrixBase::normalize(): z = squaredNorm(); if (z > 0) coeffs /= sqrt(z)
  double n2 = (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w);
  if (!(n2 > 0)) return q;
  double n = sqrt(n2);
  return Quat{q.w / n, q.x / n, q.y / n, q.z / n};
}
inline Quat quat_conj(const Quat &q) { return Quat{q.w, -q.x, -q.y, -q.z}; }
inline Vec3 quat_rotate(const Quat &q, const Vec3 &v) {  // Eigen _transformVector
  double uvx = q.y * v[2] - q.z * v[1];
  double uvy = q.z * v[0] - q.x * v[2];
  double uvz = q.x * v[1] - q.y * v[0];
  uvx += uvx;
  uvy += uvy;
  uvz += uvz;
  double cx = q.y * uvz - q.z * uvy;
  double cy = q.z * uvx - q.x * uvz;
  double cz = q.x * uvy - q.y * uvx;
  return Vec3(v[0] + q.w * uvx + cx, v[1] + q.w * uvy + cy, v[2] + q.w * uvz + cz);
}

// ---- Sophus SO3 / SE3 subset (3rdPartLib/Sophus/sophus/so3.cpp:42-90, se3.cpp:59-95) ---------------
struct SO3 {
  Quat q{1, 0, 0, 0};
  SO3() {}
  explicit SO3(const Quat &quat) : q(quat_normalized(quat)) {}  // so3.cpp:42-47
  void mul_assign(const SO3 &o) {                                // so3.cpp:73-78
    q = quat_mul(q, o.q);
    q = quat_normalized(q);
  }
  Vec3 operator*(const Vec3 &xyz) const { return quat_rotate(q, xyz); }  // so3.cpp:80-84
  SO3 inverse() const { return SO3(quat_conj(q)); }                       // so3.cpp:86-90
};
struct SE3 {
  SO3 so3;
  Vec3 t;
  SE3() {}
  SE3(const SO3 &r, const Vec3 &tr) : so3(r), t(tr) {}
  SE3 operator*(const SE3 &o) const {  // se3.cpp:59-66
    SE3 result(*this);
    result.t = result.t + (so3 * o.t);
    result.so3.mul_assign(o.so3);
    return result;
  }
  SE3 inverse() const {  // se3.cpp:76-83
    SE3 ret;
    ret.so3 = so3.inverse();
    ret.t = ret.so3 * (t * -1.);
    return ret;
  }
  Vec3 operator*(const Vec3 &xyz) const { return (so3 * xyz) + t; }  // se3.cpp:91-95
  const Vec3 &translation() const { return t; }
};
inline SE3 se3_from_pose7(const double p[7]) {
  return SE3(SO3(Quat{p[3], p[4], p[5], p[6]}), Vec3(p[0], p[1], p[2]));
}

// Sophus SO3::log / SO3::exp (3rdPartLib/Sophus/sophus/so3.cpp:127-199), SMALL_EPS = 1e-10 (so3.h:35)
inline Vec3 so3_log(const SO3 &r) {
  const double n = sqrt(r.q.x * r.q.x + r.q.y * r.q.y + r.q.z * r.q.z);
  const double w = r.q.w;
  const double squared_w = w * w;
  double two_atan_nbyw_by_n;
  if (n < 1e-10) {
    two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * squared_w);
  } else {
    // the |w| < SMALL_EPS branch of the reference is overwritten by the next statement (so3.cpp:152-165)
    two_atan_nbyw_by_n = 2 * atan(n / w) / n;
  }
  return Vec3(two_atan_nbyw_by_n * r.q.x, two_atan_nbyw_by_n * r.q.y, two_atan_nbyw_by_n * r.q.z);
}
inline SO3 so3_exp(const Vec3 &omega) {
  const double theta = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  const double half_theta = 0.5 * theta;
  double imag_factor;
  const double real_factor = cos(half_theta);
  if (theta < 1e-10) {
    const double theta_sq = theta * theta;
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
  } else {
    imag_factor = sin(half_theta) / theta;
  }
  return SO3(Quat{real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]});
}
// Eigen Quaternion::toRotationMatrix() times a vector (rot_og.matrix() * w), src/mlmap.cpp:492
inline Vec3 quat_matrix_times(const Quat &q, const Vec3 &v) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  const double m00 = 1 - (tyy + tzz), m01 = txy - twz, m02 = txz + twy;
  const double m10 = txy + twz, m11 = 1 - (txx + tzz), m12 = tyz - twx;
  const double m20 = txz - twy, m21 = tyz + twx, m22 = 1 - (txx + tyy);
  // coefficient-based product into a column vector: rows 0-1 form one packet and accumulate in order, the odd last
  // row is a scalar dot product reduced as a halving tree
  return Vec3((v[0] * m00 + v[1] * m01) + v[2] * m02, (v[0] * m10 + v[1] * m11) + v[2] * m12, m20 * v[0] + (m21 * v[1] + m22 * v[2]));
}
// pose forwarded to the image stamp, mlmap::depth_odom_input_callback src/mlmap.cpp:470-498:
// time_gap = gap_imu - latency; rot_cp = log(R) + time_gap * (R * omega); T_wb = (exp(rot_cp), p + (gap_odom - latency) * v)
inline SE3 compensate_pose(const Vec3 &pos, const Quat &quat, const Vec3 &lin_vel, const Vec3 &ang_vel, double gap_odom,
                           double gap_imu, double camera2odom_latency) {
  const double time_gap = gap_imu - camera2odom_latency;
  const SO3 rot_og(quat);
  const Vec3 rot_dot = quat_matrix_times(rot_og.q, ang_vel);
  const Vec3 rot_cp = so3_log(rot_og) + rot_dot * time_gap;
  return SE3(so3_exp(rot_cp), pos + lin_vel * (gap_odom - camera2odom_latency));
}

// ---- awareness_map_cylindrical (include/map_awareness.h, src/map_awareness.cpp) ------------------
#define ORC_deg2rad M_PI / 180 /* unparenthesised on purpose, map_awareness.h:7 */
class awareness_map {
 public:
  int nRho_x_nPhi;
  SE3 T_bs;
  bool visibility_check;
  double map_dRho, map_dPhi;
  int map_nRho, map_nPhi, map_center_z_idx;
  double noise_coe_;
  vector<vector<float>> get_odds_table;
  int diff_range;
  struct VectorHasher {  // map_awareness.h:31-41
    int operator()(const Vec3I &V) const {
      int hash = V.size();
      hash ^= V[0] + 0x9e3779b9 + (hash << 6) + (hash >> 2);
      hash ^= V[1] + 0x9e3779b9 + (hash << 6) + (hash >> 2);
      hash ^= V[2] + 0x9e3779b9 + (hash << 6) + (hash >> 2);
      return hash;
    }
  };
  SE3 T_wa;
  int map_nZ;
  double map_dZ;
  double z_border_min;
  // The reference materialises an 80-byte CYLINDRICAL_CELL per cell (data_type.h:6-15); only
  // center_pt and raycasting_z_over_rho are read on the path.  Both are separable, so the oracle
  // keeps the same values in factored tables: centre xy per (phi,rho), centre z per z
  // (map_awareness.cpp:57-62) and the rate recomputed from the identical expression (:64-71).
  vector<double> center_x, center_y;  // [phi*nRho + rho]
  vector<double> center_z;            // [z]
  unordered_map<Vec3I, float, VectorHasher> hit_idx_odds_hashmap;  // map_awareness.h:56
  unordered_set<size_t> miss_idx_set;                              // map_awareness.h:57
  size_t n_inside = 0, n_cast = 0;
  // harness-only: distinct hit keys in first-insert order (for validating the iteration-order model)
  bool log_inserts = false;
  vector<Vec3I> insert_log;

  inline size_t mapIdx(int Rho, int Phi, int z) {  // map_awareness.h:81-84
    return static_cast<size_t>(z * this->nRho_x_nPhi + Phi * this->map_nRho + Rho);
  }
  inline size_t mapIdx(Vec3I r) { return mapIdx(r(0), r(1), r(2)); }
  inline Vec3 center_pt(size_t idx) const {
    size_t z = idx / nRho_x_nPhi, rem = idx % nRho_x_nPhi;
    return Vec3(center_x.at(rem), center_y.at(rem), center_z.at(z));
  }
  inline double raycasting_z_over_rho(int rho, int z) const {  // map_awareness.cpp:64-71
    if (rho > 0) return (z - map_center_z_idx) / (rho * 1.0);
    return 0;
  }
  inline double fast_atan(double x) { return x * (45 - (x - 1) * (14 + 3.83 * x)); }  // h:115-118
  inline double fast_atan2(double y, double x) {                                       // h:86-113
    double input = y / x;
    double a_input = abs(input);
    double res;
    if (a_input > 1) {
      res = copysign(ORC_deg2rad * (90 - fast_atan(1 / a_input)), input);
    } else {
      res = copysign(ORC_deg2rad * fast_atan(a_input), input);
    }
    if (x > 0) {
      return res;
    } else if (y >= 0) {
      return res + M_PI;
    } else {
      return res - M_PI;
    }
  }
  inline float sigma_in_dr(size_t x) {  // h:120-124
    float dis = (x * this->map_dRho);
    return noise_coe_ * dis * dis / this->map_dRho;
  }
  inline float standard_ND(float x) {  // h:126-146 (A&S 7.1.26)
    double a1 = 0.254829592;
    double a2 = -0.284496736;
    double a3 = 1.421413741;
    double a4 = -1.453152027;
    double a5 = 1.061405429;
    double p = 0.3275911;
    int sign = 1;
    if (x < 0) sign = -1;
    x = fabs(x) / sqrt(2.0);
    double t = 1.0 / (1.0 + p * x);
    double y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * exp(-x * x);
    return 0.5 * (1.0 + sign * y);
  }
  float get_odds(int diff, size_t r) {  // map_awareness.cpp:119-132
    if (r == 0) {
      r = 1;
    }
    float up = standard_ND(static_cast<float>(diff + 0.5) / sigma_in_dr(r));
    float down = standard_ND(static_cast<float>(diff - 0.5) / sigma_in_dr(r));
    float res = up - down < 0.001 ? 0.001 : up - down;
    res = res >= 0.999 ? 0.999 : res;
    return res;
  }
  inline void update_odds_hashmap(Vec3I rpz_idx, float odd) {  // h:147-154
    if (hit_idx_odds_hashmap.find(rpz_idx) == hit_idx_odds_hashmap.end()) {
      if (log_inserts) insert_log.push_back(rpz_idx);
      hit_idx_odds_hashmap[rpz_idx] = odd;
    } else
      hit_idx_odds_hashmap[rpz_idx] = 1 - (1 - hit_idx_odds_hashmap[rpz_idx]) * (1 - odd);
  }

  void init_map(double d_Rho, double d_Phi_deg, double d_Z, int n_Rho, int n_z_below, int n_z_over,
                bool apply_raycasting, double noise_coe) {  // map_awareness.cpp:19-82
    this->noise_coe_ = noise_coe;
    this->map_dRho = d_Rho;
    this->map_dPhi = d_Phi_deg * M_PI / 180;
    this->map_dZ = d_Z;
    this->map_nRho = n_Rho;
    this->map_nPhi = static_cast<int>(360 / d_Phi_deg);
    this->map_nZ = n_z_below + n_z_over + 1;
    this->map_center_z_idx = n_z_below;
    this->z_border_min = -(n_z_below * d_Z) - 0.5 * d_Z;
    this->nRho_x_nPhi = map_nRho * map_nPhi;
    diff_range = 10;
    for (int diff = -diff_range; diff < diff_range + 1; diff++) {
      vector<float> line;
      for (int r = 0; r < n_Rho; r++) {
        line.emplace_back(get_odds(diff, r));
      }
      get_odds_table.emplace_back(line);
    }
    center_x.resize(nRho_x_nPhi);
    center_y.resize(nRho_x_nPhi);
    center_z.resize(map_nZ);
    for (int z = 0; z < this->map_nZ; z++) {
      center_z[z] = this->z_border_min + (this->map_dZ / 2) + (z * this->map_dZ);
    }
    for (int phi = 0; phi < this->map_nPhi; phi++) {
      for (int rho = 0; rho < this->map_nRho; rho++) {
        double center_rho = this->map_dRho / 2 + (rho * this->map_dRho);
        double center_phi = this->map_dPhi / 2 + (phi * this->map_dPhi);
        center_x[phi * map_nRho + rho] = center_rho * cos(center_phi);
        center_y[phi * map_nRho + rho] = center_rho * sin(center_phi);
      }
    }
    visibility_check = apply_raycasting;
  }

  bool xyz2RhoPhiZwithBoderCheck(Vec3 xyz_l, Vec3I &rhophiz, bool &can_do_cast) {  // cpp:84-107
    double rho = sqrt(pow(xyz_l(0), 2) + pow(xyz_l(1), 2));
    int rho_idx = static_cast<int>(rho / this->map_dRho);
    double phi = fast_atan2(xyz_l(1), xyz_l(0));
    if (phi < 0) phi += 2 * M_PI;
    int phi_idx = static_cast<int>(phi / this->map_dPhi);
    double z = xyz_l(2) - this->z_border_min;
    int z_idx = static_cast<int>(floor(z / this->map_dZ));
    rhophiz = Vec3I(rho_idx, phi_idx, z_idx);
    can_do_cast = (rho_idx >= 0 && phi_idx >= 0 && phi_idx < this->map_nPhi);
    if (can_do_cast && z_idx >= 0 && rho_idx < this->map_nRho && z_idx < this->map_nZ) {
      return true;
    }
    return false;
  }

  void update_hits(Vec3 /*p_l*/, Vec3I rpz_idx, size_t /*map_idx*/) {  // cpp:135-171
    int raycasting_z;
    double raycasting_rate = raycasting_z_over_rho(rpz_idx[0], rpz_idx[2]);
    float odd;
    Vec3I neighbor;
    update_odds_hashmap(rpz_idx, get_odds_table[0 + diff_range][rpz_idx[0]]);
    for (auto diff_r = 1; diff_r < 3 * sigma_in_dr(rpz_idx[0]) && (rpz_idx[0] + diff_r < this->map_nRho);
         diff_r++) {
      raycasting_z = static_cast<int>(round(rpz_idx[2] + (diff_r * raycasting_rate)));
      odd = get_odds_table[diff_r + diff_range][rpz_idx[0]];
      if (0 <= raycasting_z && raycasting_z < map_nZ) {
        neighbor = {rpz_idx[0] + diff_r, rpz_idx[1], raycasting_z};
        update_odds_hashmap(neighbor, odd);
      }
      odd = get_odds_table[-diff_r + diff_range][rpz_idx[0]];
      raycasting_z = static_cast<int>(round(rpz_idx[2] - (diff_r * raycasting_rate)));
      if (0 <= raycasting_z && raycasting_z < map_nZ) {
        neighbor = {rpz_idx[0] - diff_r, rpz_idx[1], raycasting_z};
        update_odds_hashmap(neighbor, odd);
      }
    }
  }

  void input_pc_pose(vector<Vec3> PC_s, SE3 T_wb) {  // cpp:173-282 (PC_s by value, like the reference)
    this->hit_idx_odds_hashmap.clear();
    this->miss_idx_set.clear();
    insert_log.clear();
    n_inside = n_cast = 0;
    T_wa = SE3(SO3(Quat{1, 0, 0, 0}), T_wb.translation());
    SE3 T_ws = T_wb * this->T_bs;
    SE3 T_ls = T_wa.inverse() * T_ws;
    for (auto p_s : PC_s) {
      auto p_l = T_ls * p_s;
      Vec3I rpz_idx;
      bool can_do_cast;
      size_t map_idx = 0;
      bool inside_range = xyz2RhoPhiZwithBoderCheck(p_l, rpz_idx, can_do_cast);
      if (inside_range) {
        map_idx = mapIdx(rpz_idx);
        update_hits(p_l, rpz_idx, map_idx);
        n_inside++;
      }
      if (can_do_cast && visibility_check) {
        n_cast++;
        double raycasting_rate;
        if (inside_range) {
          raycasting_rate = raycasting_z_over_rho(rpz_idx[0], rpz_idx[2]);
        } else {
          if (rpz_idx[0] > 0) {
            raycasting_rate = (rpz_idx[2] - this->map_center_z_idx) / (rpz_idx[0] * 1.0);
          } else {
            raycasting_rate = 0;
          }
        }
        if (rpz_idx[0] >= map_nRho) {
          rpz_idx[2] = static_cast<int>(round(rpz_idx[2] - ((rpz_idx[0] - map_nRho + 1) * raycasting_rate)));
          rpz_idx[0] = map_nRho - 1;
        }
        for (int r = rpz_idx[0] - 1; r > 0; r--) {
          int diff_r = rpz_idx[0] - r;
          int raycasting_z = static_cast<int>(round(rpz_idx[2] - (diff_r * raycasting_rate)));
          if (0 <= raycasting_z && raycasting_z < map_nZ)
            miss_idx_set.emplace(this->mapIdx(Vec3I(r, rpz_idx[1], raycasting_z)));
        }
      }
      // else: the reference prints "point out range" (cpp:277-278); output only, dropped here
    }
  }
};

// ---- local_map_cartesian (include/map_local.h, src/map_local.cpp) ---------------------------------
#define ORC_logit(x) (log10((x) / (1 - (x)))) /* map_local.h:8 */
struct Mat6x4I {
  int m[6][4];
};
class local_map {
 public:
  vector<Vec3I> nbr_disp;
  vector<Vec3> nbr_disp_real;
  struct VectorHasher {  // map_local.h:42-52
    int operator()(const Vec3I &V) const {
      int hash = V.size();
      hash ^= V[0] + 0x9e3779b9 + (hash << 6) + (hash >> 2);
      hash ^= V[1] + 0x9e3779b9 + (hash << 6) + (hash >> 2);
      hash ^= V[2] + 0x9e3779b9 + (hash << 6) + (hash >> 2);
      return hash;
    }
  };
  struct subbox {  // map_local.h:53-60
    vector<char> occupancy;
    vector<char> inflate_occupancy;
    vector<float> log_odds;
    unordered_set<int> frontier;
  };
  float log_odds_hit, log_odds_miss, log_odds_occupied_sh;
  int inflate_n = 3;
  bool apply_inflate = false;
  double flate_height = 0.1;
  bool apply_explored_area;
  float log_odds_max, log_odds_min;
  double map_dxyz_obv_glb, map_dxyz_obv_sub, map_dxyz_obv_sub_half;
  size_t cell_num_subbox;
  int subbox_nxyz;
  int ram_expand_cnt = 0;
  int obs_cnt = 0;
  float map_reso_inv;
  vector<double> global_bd = vector<double>(6);
  unordered_map<Vec3I, size_t, VectorHasher> subbox_cell_id_table;
  vector<Vec3I> subbox_id2xyz_table;
  unordered_map<Vec3I, subbox, VectorHasher> observed_group_map;
  vector<Mat6x4I> subbox_neighbors;
  unordered_set<Vec3I, VectorHasher> observed_subboxes;
  // bookkeeping for the parity harness only (not in the reference): distinct cells updated this frame
  size_t n_touched_last = 0;
  int n_released_last = 0;
  bool bookkeeping = true;  // harness-only counters; switched off while the oracle is being timed

  void init_map(double d_xyz_in, unsigned int subbox_n, float log_odds_min_in, float log_odds_max_in,
                float log_odds_hit_in, float log_odds_miss_in, float log_odds_occupied_sh_in,
                bool if_apply_explor) {  // map_local.cpp:46-139
    map_dxyz_obv_sub = d_xyz_in;
    map_dxyz_obv_sub_half = map_dxyz_obv_sub * 0.5;
    map_reso_inv = 1 / map_dxyz_obv_sub;
    subbox_nxyz = subbox_n;
    map_dxyz_obv_glb = map_dxyz_obv_sub * subbox_nxyz;
    cell_num_subbox = pow(subbox_nxyz, 3);
    int cnt = 0;
    for (auto i = 0; i < subbox_nxyz; i++) {
      for (auto j = 0; j < subbox_nxyz; j++) {
        for (auto k = 0; k < subbox_nxyz; k++) {
          subbox_cell_id_table[Vec3I(k, j, i)] = cnt++;
          subbox_id2xyz_table.emplace_back(k, j, i);
        }
      }
    }
    Vec3I temp_id;
    nbr_disp.emplace_back(Vec3I(0, 0, 1));
    nbr_disp.emplace_back(Vec3I(0, 0, -1));
    nbr_disp.emplace_back(Vec3I(0, 1, 0));
    nbr_disp.emplace_back(Vec3I(0, -1, 0));
    nbr_disp.emplace_back(Vec3I(1, 0, 0));
    nbr_disp.emplace_back(Vec3I(-1, 0, 0));
    nbr_disp_real.emplace_back(Vec3(0, 0, map_dxyz_obv_sub));
    nbr_disp_real.emplace_back(Vec3(0, 0, -map_dxyz_obv_sub));
    nbr_disp_real.emplace_back(Vec3(0, map_dxyz_obv_sub, 0));
    nbr_disp_real.emplace_back(Vec3(0, -map_dxyz_obv_sub, 0));
    nbr_disp_real.emplace_back(Vec3(map_dxyz_obv_sub, 0, 0));
    nbr_disp_real.emplace_back(Vec3(-map_dxyz_obv_sub, 0, 0));
    for (auto i = 0; i < subbox_nxyz; i++) {
      for (auto j = 0; j < subbox_nxyz; j++) {
        for (auto k = 0; k < subbox_nxyz; k++) {
          Mat6x4I nbrs;
          for (auto n = 0; n < 6; n++) {
            Vec3I glb_disp(0, 0, 0);
            temp_id = nbr_disp[n] + Vec3I(k, j, i);
            for (auto m = 0; m < 3; m++) {
              if (temp_id[m] >= subbox_nxyz) {
                glb_disp[m] = 1;
                temp_id[m] = 0;
              } else if (temp_id[m] < 0) {
                glb_disp[m] = -1;
                temp_id[m] = subbox_nxyz - 1;
              }
            }
            nbrs.m[n][0] = glb_disp[0];
            nbrs.m[n][1] = glb_disp[1];
            nbrs.m[n][2] = glb_disp[2];
            nbrs.m[n][3] = subbox_cell_id_table[temp_id];
          }
          subbox_neighbors.emplace_back(nbrs);
        }
      }
    }
    apply_explored_area = if_apply_explor;
    global_bd = {-30, 30, -30, 30, 0, 5};
    log_odds_min = log_odds_min_in;
    log_odds_max = log_odds_max_in;
    log_odds_hit = log_odds_hit_in;
    log_odds_miss = log_odds_miss_in;
    log_odds_occupied_sh = log_odds_occupied_sh_in;
  }

  inline void get_global_idx(const Vec3 &pt_w, Vec3I &glb_idx, size_t &subbox_id) {  // map_local.h:148-152
    glb_idx = Vec3I(floor(pt_w[0] / map_dxyz_obv_glb), floor(pt_w[1] / map_dxyz_obv_glb),
                    floor(pt_w[2] / map_dxyz_obv_glb));
    subbox_id = get_subbox_id(pt_w, glb_idx);
  }
  inline bool inside_exp_bd(Vec3 pt_w) {  // map_local.h:160-165
    return (pt_w[0] >= global_bd[0] && pt_w[0] < global_bd[1] && pt_w[1] >= global_bd[2] &&
            pt_w[1] < global_bd[3] && pt_w[2] >= global_bd[4] && pt_w[2] < global_bd[5]);
  }
  inline size_t get_subbox_id(const Vec3 &pt_w, const Vec3I &glb_idx) {  // map_local.h:167-173
    return subbox_cell_id_table[Vec3I(floor(pt_w[0] / map_dxyz_obv_sub) - glb_idx[0] * subbox_nxyz,
                                      floor(pt_w[1] / map_dxyz_obv_sub) - glb_idx[1] * subbox_nxyz,
                                      floor(pt_w[2] / map_dxyz_obv_sub) - glb_idx[2] * subbox_nxyz)];
  }
  inline Vec3 subbox_id2xyz_glb_vec(const Vec3I &origin, int idx) {  // map_local.h:208-213
    return Vec3(origin[0] * map_dxyz_obv_glb + subbox_id2xyz_table[idx][0] * map_dxyz_obv_sub + map_dxyz_obv_sub_half,
                origin[1] * map_dxyz_obv_glb + subbox_id2xyz_table[idx][1] * map_dxyz_obv_sub + map_dxyz_obv_sub_half,
                origin[2] * map_dxyz_obv_glb + subbox_id2xyz_table[idx][2] * map_dxyz_obv_sub + map_dxyz_obv_sub_half);
  }
  inline bool allocate_ram(Vec3I &glb_idx) {  // map_local.h:215-231
    if (observed_group_map.find(glb_idx) == observed_group_map.end()) {
      observed_group_map[glb_idx].occupancy.resize(cell_num_subbox, 'u');
      observed_group_map[glb_idx].inflate_occupancy.resize(cell_num_subbox, 'u');
      observed_group_map[glb_idx].log_odds.resize(cell_num_subbox, 0);
      observed_group_map[glb_idx].frontier.clear();
      ram_expand_cnt++;
      return true;
    } else if (observed_group_map[glb_idx].occupancy.size() == 1)
      return false;
    return true;
  }
  inline void inflate_atpos(const Vec3I &glb_idx, size_t subbox_id) {  // map_local.h:233-264
    Vec3I off;
    for (off(0) = -inflate_n; off(0) <= inflate_n; off(0)++)
      for (off(1) = -inflate_n; off(1) <= inflate_n; off(1)++)
        for (off(2) = -inflate_n; off(2) <= inflate_n; off(2)++) {
          if (abs(off(0)) + abs(off(1)) + abs(off(2)) > inflate_n) continue;  // lpNorm<1>
          Vec3I subbox_id_inflate = off + subbox_id2xyz_table[subbox_id];
          bool expanded = false;
          Vec3I glb_idx_inflate = glb_idx;
          for (auto m = 0; m < 3; m++) {
            if (subbox_id_inflate[m] >= subbox_nxyz) {
              glb_idx_inflate[m] += 1;
              subbox_id_inflate[m] = subbox_id_inflate[m] - subbox_nxyz;
              expanded = true;
            } else if (subbox_id_inflate[m] < 0) {
              glb_idx_inflate[m] += -1;
              subbox_id_inflate[m] = subbox_nxyz + subbox_id_inflate[m];
              expanded = true;
            }
          }
          if ((expanded && allocate_ram(glb_idx_inflate)) || !expanded)
            observed_group_map[glb_idx_inflate].inflate_occupancy[subbox_cell_id_table[subbox_id_inflate]] = 'o';
        }
  }

  void update_observation(Vec3I glb_idx, size_t subbox_id, Vec3 pt_w) {  // map_local.cpp:7-33
    if (!inside_exp_bd(pt_w)) return;
    observed_subboxes.emplace(glb_idx);
    Vec3I glb_idx_nb;
    size_t subbox_id_nb;
    for (auto i = 0; i < 6; i++) {
      Vec3 pt_w_nb = pt_w + nbr_disp_real[i];
      const Mat6x4I &nb = subbox_neighbors[subbox_id];
      glb_idx_nb = glb_idx + Vec3I(nb.m[i][0], nb.m[i][1], nb.m[i][2]);
      subbox_id_nb = nb.m[i][3];
      if (inside_exp_bd(pt_w_nb) && allocate_ram(glb_idx_nb) &&
          (observed_group_map[glb_idx_nb].occupancy[subbox_id_nb] == 'u')) {
        observed_group_map[glb_idx_nb].frontier.emplace(subbox_id_nb);
        break;
      }
    }
    return;
  }

  void input_pc_pose_direct(awareness_map *a_map) {  // map_local.cpp:143-237
    SE3 T_wa = a_map->T_wa;
    unordered_set<uint64_t> touched;  // harness bookkeeping only
    auto touch = [&](const Vec3I &g, size_t s) {
      if (!bookkeeping) return;
      // pack for counting distinct updated cells; not part of the reference
      uint64_t k = (uint64_t)(uint32_t)observed_group_slot(g) * (uint64_t)cell_num_subbox + s;
      touched.insert(k);
    };
    for (auto pair_ : a_map->hit_idx_odds_hashmap) {
      Vec3I glb_idx;
      size_t subbox_id;
      Vec3 p_w = T_wa * a_map->center_pt(a_map->mapIdx(pair_.first));
      get_global_idx(p_w, glb_idx, subbox_id);
      if (allocate_ram(glb_idx)) {
        touch(glb_idx, subbox_id);
        if (observed_group_map[glb_idx].log_odds[subbox_id] < log_odds_max) {
          observed_group_map[glb_idx].log_odds[subbox_id] += ORC_logit(pair_.second);
          observed_group_map[glb_idx].log_odds[subbox_id] =
              observed_group_map[glb_idx].log_odds[subbox_id] > log_odds_max
                  ? log_odds_max
                  : observed_group_map[glb_idx].log_odds[subbox_id];
        }
        if (observed_group_map[glb_idx].log_odds[subbox_id] > log_odds_occupied_sh &&
            observed_group_map[glb_idx].occupancy[subbox_id] != 'o') {
          observed_group_map[glb_idx].occupancy[subbox_id] = 'o';
          if (apply_explored_area) observed_group_map[glb_idx].frontier.erase(subbox_id);
          obs_cnt++;
        }
      }
    }
    for (auto idx : a_map->miss_idx_set) {
      Vec3I glb_idx;
      size_t subbox_id;
      Vec3 p_w = T_wa * a_map->center_pt(idx);
      get_global_idx(p_w, glb_idx, subbox_id);
      if (allocate_ram(glb_idx)) {
        touch(glb_idx, subbox_id);
        if (observed_group_map[glb_idx].log_odds[subbox_id] >= log_odds_min) {
          observed_group_map[glb_idx].log_odds[subbox_id] += log_odds_miss;
          observed_group_map[glb_idx].log_odds[subbox_id] =
              observed_group_map[glb_idx].log_odds[subbox_id] < log_odds_min
                  ? log_odds_min
                  : observed_group_map[glb_idx].log_odds[subbox_id];
        }
        if (observed_group_map[glb_idx].log_odds[subbox_id] < log_odds_occupied_sh &&
            observed_group_map[glb_idx].occupancy[subbox_id] != 'f') {
          if (observed_group_map[glb_idx].occupancy[subbox_id] == 'u' && apply_explored_area)
            update_observation(glb_idx, subbox_id, p_w);
          observed_group_map[glb_idx].occupancy[subbox_id] = 'f';
          if (apply_explored_area) observed_group_map[glb_idx].frontier.erase(subbox_id);
        }
      }
    }
    n_released_last = 0;
    for (auto glb_idx : observed_subboxes) {
      if (observed_group_map.find(glb_idx) != observed_group_map.end() &&
          observed_group_map[glb_idx].occupancy.size() > 1 && observed_group_map[glb_idx].frontier.empty()) {
        if (std::adjacent_find(observed_group_map[glb_idx].occupancy.begin(),
                               observed_group_map[glb_idx].occupancy.end(),
                               std::not_equal_to<char>()) == observed_group_map[glb_idx].occupancy.end()) {
          observed_group_map[glb_idx].occupancy.resize(1);
          observed_group_map[glb_idx].occupancy.shrink_to_fit();
          observed_group_map[glb_idx].inflate_occupancy.resize(1);
          observed_group_map[glb_idx].inflate_occupancy.shrink_to_fit();
          observed_group_map[glb_idx].log_odds.resize(1);
          observed_group_map[glb_idx].log_odds.shrink_to_fit();
          n_released_last++;  // reference prints "memory release!" (map_local.cpp:228)
        }
      }
    }
    observed_subboxes.clear();
    n_touched_last = touched.size();
  }

 private:
  // harness bookkeeping: stable small id per subbox for the touched-cell counter
  unordered_map<Vec3I, uint32_t, VectorHasher> slot_ids_;
  uint32_t observed_group_slot(const Vec3I &g) {
    auto it = slot_ids_.find(g);
    if (it != slot_ids_.end()) return it->second;
    uint32_t id = (uint32_t)slot_ids_.size();
    slot_ids_[g] = id;
    return id;
  }
};

// ---- class mlmap (include/mlmap.h, src/mlmap.cpp), ROS-free ----------------------------------------
#define ORC_logit_inv(x) (pow(10, x) / (1 + pow(10, x))) /* mlmap.h:40 */
class mlmap {
 public:
  awareness_map *awareness;
  local_map *local;
  enum { FREE = 1, OCCUPIED = 0, UNKNOWN = -1 };  // mlmap.h:109-114
  const double k_depth_scaling_factor_ = 1000.0;
  const double inv_factor = 1.0 / k_depth_scaling_factor_;
  size_t pc_sample_cnt = 0;
  vector<Vec3> pc_eigen;
  float cx_, cy_, fx_, fy_;
  int inflate_global_n = 2;
  vector<Vec3I> glb_idx_nb_list = vector<Vec3I>(6);
  vector<size_t> subbox_id_nb_list = vector<size_t>(6);
  Vec3 ct_pos;
  SE3 T_wb;

  explicit mlmap(const mlm_config &c) {  // mlmap::init_map, src/mlmap.cpp:3-149 minus ROS
    inflate_global_n = c.inflate_global_n;
    pc_sample_cnt = c.sample_cnt;
    cx_ = c.cam_cx;
    cy_ = c.cam_cy;
    fx_ = c.cam_fx;
    fy_ = c.cam_fy;
    awareness = new awareness_map();
    awareness->T_bs = se3_from_pose7(c.T_bs);
    awareness->init_map(c.am_d_rho, c.am_d_phi_deg, c.am_d_z, c.am_n_rho, c.am_n_z_below, c.am_n_z_over,
                        c.use_raycasting != 0, c.depth_noise_coe);
    local = new local_map();
    local->init_map(c.subbox_d_xyz, static_cast<unsigned int>(c.subbox_n), c.log_odds_min, c.log_odds_max,
                    c.log_odds_hit, c.log_odds_miss, c.log_odds_occupied_sh, c.use_exploration_frontiers != 0);
    local->inflate_n = c.inflate_n;
    local->apply_inflate = c.apply_inflate != 0;
    local->flate_height = c.inflate_height;
  }
  ~mlmap() {
    delete awareness;
    delete local;
  }

  // project_depth, src/mlmap.cpp:311-349.  sample_cnt == 0 selects the harness' full-frame mode
  // (every pixel, v outer / u inner, same per-pixel formula).
  void project_depth(const uint16_t *img, int rows, int cols, size_t stride_bytes) {
    uint16_t *row_ptr;
    size_t u, v;
    double depth;
    Vec3 pt_cur;
    auto row_of = [&](size_t vv) {
      return reinterpret_cast<uint16_t *>(const_cast<uint8_t *>(reinterpret_cast<const uint8_t *>(img)) +
                                          vv * stride_bytes);
    };
    if (pc_sample_cnt == 0) {
      for (v = 0; v < (size_t)rows; v++)
        for (u = 0; u < (size_t)cols; u++) {
          row_ptr = row_of(v) + u;
          depth = (*row_ptr) * inv_factor;
          if (*row_ptr == 0) continue;
          pt_cur(0) = (u - cx_) * depth / fx_;
          pt_cur(1) = (v - cy_) * depth / fy_;
          pt_cur(2) = depth;
          pc_eigen.emplace_back(pt_cur);
        }
      return;
    }
    int cnt = 0;
    int max_iter = 2 * pc_sample_cnt;
    while (pc_eigen.size() < pc_sample_cnt && cnt < max_iter) {
      cnt++;
      v = static_cast<size_t>(rand() % rows);
      u = static_cast<size_t>(rand() % cols);
      row_ptr = row_of(v) + u;
      depth = (*row_ptr) * inv_factor;
      if (*row_ptr == 0) {
        continue;
      }
      pt_cur(0) = (u - cx_) * depth / fx_;
      pt_cur(1) = (v - cy_) * depth / fy_;
      pt_cur(2) = depth;
      pc_eigen.emplace_back(pt_cur);
    }
  }
  void update_map() {  // src/mlmap.cpp:382-386
    awareness->input_pc_pose(pc_eigen, T_wb);
    local->input_pc_pose_direct(awareness);
  }
  void setFree_map_in_bound(Vec3 box_min, Vec3 box_max) {  // src/mlmap.cpp:388-407
    Vec3I glb_id;
    size_t subbox_id;
    for (double x = box_min[0]; x <= box_max[0]; x += local->map_dxyz_obv_sub) {
      for (double y = box_min[1]; y <= box_max[1]; y += local->map_dxyz_obv_sub) {
        for (double z = box_min[2]; z <= box_max[2]; z += local->map_dxyz_obv_sub) {
          local->get_global_idx(Vec3(x, y, z), glb_id, subbox_id);
          if (local->observed_group_map.find(glb_id) != local->observed_group_map.end() &&
              local->observed_group_map[glb_id].occupancy.size() > 1) {
            local->observed_group_map[glb_id].occupancy[subbox_id] = 'f';
            local->observed_group_map[glb_id].log_odds[subbox_id] = 0;
          }
        }
      }
    }
  }
  void inflate_map() {  // src/mlmap.cpp:286-309
    Vec3I ct_glb;
    size_t subbox_id;
    local->get_global_idx(ct_pos, ct_glb, subbox_id);
    Vec3I off;
    for (off(0) = -inflate_global_n; off(0) <= inflate_global_n; off(0)++)
      for (off(1) = -inflate_global_n; off(1) <= inflate_global_n; off(1)++)
        for (off(2) = -inflate_global_n; off(2) <= inflate_global_n; off(2)++) {
          Vec3I temp_glb = off + ct_glb;
          if (local->observed_group_map.find(temp_glb) != local->observed_group_map.end() &&
              local->observed_group_map[temp_glb].occupancy.size() > 1) {
            local->observed_group_map[temp_glb].inflate_occupancy.clear();
            local->observed_group_map[temp_glb].inflate_occupancy.resize(local->cell_num_subbox, 'u');
            for (size_t it = 0; it < local->observed_group_map[temp_glb].occupancy.size(); it++)
              if (local->observed_group_map[temp_glb].occupancy[it] == 'o' &&
                  local->subbox_id2xyz_glb_vec(temp_glb, it)(2) > local->flate_height) {
                local->inflate_atpos(temp_glb, it);
              }
          }
        }
  }

  inline int getOccupancy(const Vec3 &pos_w) {  // mlmap.h:170-193
    Vec3I glb_id;
    size_t subbox_id;
    char res;
    local->get_global_idx(pos_w, glb_id, subbox_id);
    if (local->observed_group_map.find(glb_id) == local->observed_group_map.end())
      return UNKNOWN;
    else if (local->observed_group_map[glb_id].occupancy.size() == 1)
      res = local->observed_group_map[glb_id].occupancy[0];
    else
      res = local->observed_group_map[glb_id].occupancy[subbox_id];
    if (res == 'o')
      return OCCUPIED;
    else if (res == 'f')
      return FREE;
    else
      return UNKNOWN;
  }
  inline int getOccupancy(const Vec3 &pos_w, float inflate) {  // mlmap.h:142-169
    if (getOccupancy(pos_w) != OCCUPIED && getOccupancy(pos_w + Vec3(0, 0, inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(0, 0, -inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(0, inflate, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(0, -inflate, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(inflate, 0, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(-inflate, 0, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(-inflate, inflate, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(-inflate, -inflate, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(inflate, inflate, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(inflate, -inflate, 0)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(0, -inflate, inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(0, -inflate, -inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(0, inflate, inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(0, inflate, -inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(-inflate, 0, inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(-inflate, 0, -inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(inflate, 0, inflate)) != OCCUPIED &&
        getOccupancy(pos_w + Vec3(inflate, 0, -inflate)) != OCCUPIED)
      return FREE;
    else
      return OCCUPIED;
  }
  inline int getInflateOccupancy(const Vec3 &pos_w) {  // mlmap.h:195-211
    Vec3I glb_id;
    size_t subbox_id;
    local->get_global_idx(pos_w, glb_id, subbox_id);
    if (local->observed_group_map.find(glb_id) == local->observed_group_map.end())
      return UNKNOWN;
    else if (local->observed_group_map[glb_id].occupancy.size() == 1)
      return UNKNOWN;
    else {
      if (local->observed_group_map[glb_id].inflate_occupancy[subbox_id] == 'o')
        return OCCUPIED;
      else
        return UNKNOWN;
    }
  }
  inline float getOdd(const Vec3 &pos_w) {  // mlmap.h:213-225
    Vec3I glb_id;
    size_t subbox_id;
    local->get_global_idx(pos_w, glb_id, subbox_id);
    if (local->observed_group_map.find(glb_id) == local->observed_group_map.end())
      return 0.5;
    else if (local->observed_group_map[glb_id].log_odds.size() == 1)
      return ORC_logit_inv(local->observed_group_map[glb_id].log_odds[0]);
    else
      return ORC_logit_inv(local->observed_group_map[glb_id].log_odds[subbox_id]);
  }
  inline float getOdd(const Vec3I &glb_id, size_t subbox_id) {  // mlmap.h:227-235
    if (local->observed_group_map.find(glb_id) == local->observed_group_map.end())
      return 0.5;
    else if (local->observed_group_map[glb_id].log_odds.size() == 1)
      return ORC_logit_inv(local->observed_group_map[glb_id].log_odds[0]);
    else
      return ORC_logit_inv(local->observed_group_map[glb_id].log_odds[subbox_id]);
  }
  inline Vec3 getOddGrad(const Vec3 &pos_w, size_t max_iter = 5) {  // mlmap.h:237-295
    Vec3I glb_id;
    size_t subbox_id;
    local->get_global_idx(pos_w, glb_id, subbox_id);
    Vec3I glb_idx_nb, glb_idx_nb_min;
    size_t subbox_id_nb, subbox_id_nb_min = 0;
    float min_odd = getOdd(glb_id, subbox_id);
    float ori_odd = min_odd;
    float tmp_odd;
    bool flag = false;
    size_t iter;
    for (iter = 0; iter < max_iter && !flag; iter++) {
      for (auto i = 0; i < 6; i++) {
        if (iter == 0) {
          const Mat6x4I &nb = local->subbox_neighbors[subbox_id];
          glb_idx_nb = glb_id + Vec3I(nb.m[i][0], nb.m[i][1], nb.m[i][2]);
          subbox_id_nb = nb.m[i][3];
        } else {
          const Mat6x4I &nb = local->subbox_neighbors[subbox_id_nb_list[i]];
          glb_idx_nb = glb_idx_nb_list[i] + Vec3I(nb.m[i][0], nb.m[i][1], nb.m[i][2]);
          subbox_id_nb = nb.m[i][3];
        }
        glb_idx_nb_list[i] = glb_idx_nb;
        subbox_id_nb_list[i] = subbox_id_nb;
        tmp_odd = getOdd(glb_idx_nb, subbox_id_nb);
        if (tmp_odd < min_odd) {
          min_odd = tmp_odd;
          glb_idx_nb_min = glb_idx_nb;
          subbox_id_nb_min = subbox_id_nb;
          flag = true;
        }
      }
    }
    if (flag) {
      return (local->subbox_id2xyz_glb_vec(glb_idx_nb_min, subbox_id_nb_min) - pos_w) * (ori_odd - min_odd);
    } else {
      return Vec3(0.0, 0.0, 0.0);
    }
  }
};

}  // namespace orc
#endif
