// common.cuh -- shared declarations for the B200-native CoverageControl grid hot path.
// Compiled for sm_100a only (see build.py); no fallbacks, no other back-ends.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ccb200.h"

namespace ccb {

// ---- error plumbing: nothing throws across the C ABI -------------------------
extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

inline int fail(const char *what, const char *detail = nullptr) {
  g_last_error = what;
  if (detail) {
    g_last_error += ": ";
    g_last_error += detail;
  }
  return 1;
}

#define CCB_CUDA(expr)                                                      \
  do {                                                                      \
    cudaError_t err__ = (expr);                                             \
    if (err__ != cudaSuccess) {                                             \
      char buf__[256];                                                      \
      snprintf(buf__, sizeof buf__, "%s at %s:%d", cudaGetErrorString(err__), \
               __FILE__, __LINE__);                                         \
      return ::ccb::fail("CUDA error", buf__);                              \
    }                                                                       \
  } while (0)

#define CCB_LAUNCH_CHECK()                       \
  do {                                           \
    ::ccb::g_launches.fetch_add(1);              \
    CCB_CUDA(cudaGetLastError());                \
  } while (0)

constexpr double kEps = 1e-10;      // reference constants.h:48
constexpr double kLargeEps = 1e-4;  // reference constants.h:50

// ---- device view of a batch of environments ----------------------------------
// State layout in HBM (per environment e; N = world size, R robots, NW = ceil(N / 32)):
//   world  f32 [E][N][N]        importance map, row-major [x][y]
//   seenb  u32 [E][R][N][NW]    one BIT PLANE per robot: bit (y & 31) of word [x][y >> 5] is set once
//                               the cell has been inside the robot's sensor window
//                               (robot_map_r = seen ? world : fill, robot_model.h:98-107,180-187)
//   xseenb u32 [E][R][N][NW]    same for the per-robot exploration map (only if pUpdateExplorationMap)
//   sysb   u32 [E][2][N][NW]    plane 0: cell is in explored_idf_map_, plane 1: exploration_map_
//                               cleared (coverage_system.h:104-120)
// instead of the reference's (3 + 2R) dense N*N float maps: a robot's 256 x 256 local window is
// 8 KB of mask instead of 256 KB of floats, and world values are only read where a bit is set.
struct EnvDev {
  int E, R, N, S, L, H, NW;
  double res;
  float fill;       // pUnknownImportance * pNorm, or 0
  int use_sensor;   // pUpdateSensorView
  int upd_robot;    // pUpdateRobotMap
  int upd_expl;     // pUpdateExplorationMap
  int upd_sys;      // pUpdateSystemMap
  int check_osc;    // pCheckOscillations
  int obs_prefetch; // observation kernel: L2 prefetch of the seen box (test option, default on)
  double comm_range, max_speed, dt;
  float *world;
  uint32_t *seenb;
  uint32_t *xseenb;
  uint32_t *sysb;
  int *seen_box;   // [E][R][4]: row_min, row_max + 1, col_min, col_max + 1 of the robot-map bits ever set (a bound,
                   // kept by k_mark_windows; the window kernels only stage the mask words inside it)
  double *pos_g;   // [E][R][2] global current position
  double *pos_s;   // [E][R][2] global start position
  double *pos_l;   // [E][R][2] local current position (robot_model.h:66)
  double *hist;    // [E][R][H][2] ring buffer (coverage_system.h:88)
  int *hist_start; // [E][R]
  int *hist_len;   // [E][R]
  int *status;     // [E] sticky error bits (bit0: decentralized goal out of bounds)
};

__host__ __device__ inline size_t cell_count(const EnvDev &d) { return (size_t)d.N * d.N; }
__host__ __device__ inline size_t plane_words(const EnvDev &d) { return (size_t)d.N * d.NW; }
__host__ __device__ inline uint32_t *robot_plane(const EnvDev &d, uint32_t *base, int e, int r) {
  return base + ((size_t)e * d.R + r) * plane_words(d);
}
__host__ __device__ inline uint32_t *sys_plane(const EnvDev &d, int e, int which) {
  return d.sysb + ((size_t)e * 2 + which) * plane_words(d);
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t plane_bit(const uint32_t *plane, int NW, int row, int col) {
  return (__ldg(plane + (size_t)row * NW + (col >> 5)) >> (col & 31)) & 1u;
}
#endif

// map_utils.h:51-87 ComputeOffsets, reduced to the clipped window [x0,x1) x [y0,y1) and the
// unclipped origin (ox, oy) = (index.left, index.bottom).
struct Window {
  int ox, oy;          // unclipped origin in map cells
  int x0, x1, y0, y1;  // clipped
};

__host__ __device__ inline Window make_window(double res, double px, double py, int sub, int map_size) {
  Window w;
  int cx = (int)round(px / res);  // std::round: half away from zero (map_utils.h:60-63)
  int cy = (int)round(py / res);
  w.ox = cx - sub / 2;
  w.oy = cy - sub / 2;
  int r = cx + sub / 2, t = cy + sub / 2;
  w.x0 = w.ox < 0 ? 0 : w.ox;
  w.y0 = w.oy < 0 ? 0 : w.oy;
  w.x1 = r > map_size ? map_size : r;
  w.y1 = t > map_size ? map_size : t;
  if (w.x1 < w.x0) w.x1 = w.x0;
  if (w.y1 < w.y0) w.y1 = w.y0;
  return w;
}

}  // namespace ccb
