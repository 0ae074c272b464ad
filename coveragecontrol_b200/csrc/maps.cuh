// maps.cuh -- robot kinematics, sensor-window map updates, Voronoi cell finalisation, CVT actions
// and dense map materialisation.
//
// Replaces reference robot_model.h:87-122,180-345, coverage_system.h:104-130,279-363 and
// coverage_system.cpp:201-291.  The reference keeps (3 + 2R) dense N x N float maps per
// environment and rewrites system_map_ = explored - exploration over the full map every step;
// here the state is the world map plus per-cell "seen" bit masks (common.cuh) and every other map
// is a view computed on demand.
#pragma once

#include "common.cuh"

namespace ccb {

// ---- value of the derived maps at one cell -----------------------------------------------------
__device__ __forceinline__ float robot_map_value(const EnvDev &d, int e, int r, int row, int col) {
  if (plane_bit(robot_plane(d, d.seenb, e, r), d.NW, row, col))
    return d.use_sensor ? d.world[(size_t)e * cell_count(d) + (size_t)row * d.N + col] : 0.f;
  return d.fill;
}
__device__ __forceinline__ float explored_value(const EnvDev &d, int e, int row, int col) {
  if (plane_bit(sys_plane(d, e, 0), d.NW, row, col))
    return d.use_sensor ? d.world[(size_t)e * cell_count(d) + (size_t)row * d.N + col] : 0.f;
  return 0.f;
}
__device__ __forceinline__ float exploration_value(const EnvDev &d, int e, int row, int col) {
  return plane_bit(sys_plane(d, e, 1), d.NW, row, col) ? 0.f : 1.f;
}
// robot_model.h:137-148: per-robot exploration map (1 = unexplored), all zeros when disabled
__device__ __forceinline__ float robot_exploration_value(const EnvDev &d, int e, int r, int row, int col) {
  if (!d.upd_expl) return 0.f;
  return plane_bit(robot_plane(d, d.xseenb, e, r), d.NW, row, col) ? 0.f : 1.f;
}

// ---- exploration (frontier) features --------------------------------------------------------------
// robot_model.h:347-403 GetExplorationFeatures.  Candidate points q on a res-spaced lattice within
// 4 time-step distances of the robot; for each, over the sensor window centred at q:
//   unexplored = sum of the robot's exploration map, benefit = unexplored (+ 2 x the number of non-zero
//   robot-map cells if unexplored > 0.1 S^2), value = benefit / max(|q - p|, time-step distance);
// the pNumFrontiers best candidates are kept in a std::priority_queue (worst on top) and popped worst
// first into [dx, dy] pairs.  One CTA per (robot, environment): both indicator maps over the union of
// the candidate windows become integral images in shared memory (a window sum = four look-ups), the
// candidates are scored in parallel, and one thread replays the queue in the reference's order with
// libstdc++'s heap algorithms (bits/stl_heap.h __push_heap / __adjust_heap), so ties resolve as there.
// Dynamic smem: 2 (nxr + 1)(nyr + 1) ints + nc doubles, sized by the host from the same bounds.
constexpr int EXPL_MAXC = 96;   // candidates per axis
constexpr int EXPL_MAXK = 64;   // pNumFrontiers

struct ExplHeap {
  double v[EXPL_MAXK];
  int id[EXPL_MAXK];
  int n;
};
// comp(a, b) of FrontierCompare (typedefs.h:79-83): a.value > b.value
__host__ __device__ __forceinline__ void expl_push_heap(ExplHeap &h, int hole, int top, double val, int id) {
  int parent = (hole - 1) / 2;
  while (hole > top && h.v[parent] > val) {
    h.v[hole] = h.v[parent], h.id[hole] = h.id[parent];
    hole = parent, parent = (hole - 1) / 2;
  }
  h.v[hole] = val, h.id[hole] = id;
}
__host__ __device__ __forceinline__ void expl_adjust_heap(ExplHeap &h, int hole, int len, double val, int id) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (h.v[child] > h.v[child - 1]) --child;
    h.v[hole] = h.v[child], h.id[hole] = h.id[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h.v[hole] = h.v[child - 1], h.id[hole] = h.id[child - 1];
    hole = child - 1;
  }
  expl_push_heap(h, hole, top, val, id);
}
// priority_queue::pop: the top leaves, the last element is sifted in from the root
__host__ __device__ __forceinline__ void expl_pop(ExplHeap &h) {
  if (h.n > 1) {
    const double val = h.v[h.n - 1];
    const int id = h.id[h.n - 1];
    h.v[h.n - 1] = h.v[0], h.id[h.n - 1] = h.id[0];
    expl_adjust_heap(h, 0, h.n - 1, val, id);
  }
  --h.n;
}

// robot_model.h:383-391: keep the K best of val[0 .. n), the worst on top
__host__ __device__ __forceinline__ void expl_select(ExplHeap &h, const double *val, int n, int K) {
  h.n = 0;
  for (int t = 0; t < n; ++t) {
    const double bcr = val[t];
    if (h.n < K) {
      ++h.n;  // push_back + push_heap
      expl_push_heap(h, h.n - 1, 0, bcr, t);
    } else if (h.v[0] < bcr) {
      expl_pop(h);
      ++h.n;
      expl_push_heap(h, h.n - 1, 0, bcr, t);
    }
  }
}

// the reference's candidate loop along one axis: for (q = max(p - reach, kLargeEps); q < min(p + reach, size); q += res)
__device__ __forceinline__ int expl_axis(double p, double reach, double size, double res, double *q) {
  const double lim = fmin(p + reach, size);
  int n = 0;
  for (double v = fmax(p - reach, kLargeEps); v < lim && n < EXPL_MAXC; v += res) q[n++] = v;
  return n;
}

__global__ void __launch_bounds__(256) k_exploration_features(EnvDev d, int K, double *__restrict__ out) {
  extern __shared__ int s_dyn[];
  __shared__ double s_qx[EXPL_MAXC], s_qy[EXPL_MAXC];
  __shared__ int s_cx[EXPL_MAXC], s_cy[EXPL_MAXC];
  __shared__ int s_nx, s_ny;
  __shared__ ExplHeap s_heap;
  const int r = blockIdx.x, e = blockIdx.y, tid = threadIdx.x;
  const size_t er = (size_t)e * d.R + r;
  const double px = d.pos_g[er * 2], py = d.pos_g[er * 2 + 1];
  const double res = d.res;
  const double tsd = d.max_speed * d.dt * res;  // time_step_dist_ (robot_model.h:150-151)
  const double reach = 4.0 * tsd, size = (double)d.N;  // (the reference compares with pWorldMapSize, not its extent)
  const int S = d.S, half = S / 2;
  if (tid == 0) {
    s_nx = expl_axis(px, reach, size, res, s_qx);
    for (int k = 0; k < s_nx; ++k) s_cx[k] = (int)round(s_qx[k] / res);
  }
  if (tid == 32) {
    s_ny = expl_axis(py, reach, size, res, s_qy);
    for (int k = 0; k < s_ny; ++k) s_cy[k] = (int)round(s_qy[k] / res);
  }
  __syncthreads();
  const int nx = s_nx, ny = s_ny;
  double *o = out + er * 2 * K;
  if (nx == 0 || ny == 0) {
    for (int t = tid; t < 2 * K; t += blockDim.x) o[t] = 0.0;
    return;
  }
  // union of the candidate windows (the centres are monotone along an axis), clipped to the map
  const int x0 = max(s_cx[0] - half, 0), x1 = min(s_cx[nx - 1] + half, d.N);
  const int y0 = max(s_cy[0] - half, 0), y1 = min(s_cy[ny - 1] + half, d.N);
  const int nxr = max(x1 - x0, 0), nyr = max(y1 - y0, 0), ld = nyr + 1;
  int *s_un = s_dyn, *s_nz = s_dyn + (size_t)(nxr + 1) * ld;
  double *s_val = reinterpret_cast<double *>(s_dyn + (((size_t)2 * (nxr + 1) * ld + 1) & ~(size_t)1));
  // indicators, shifted by one row and one column; row 0 and column 0 stay zero
  for (int t = tid; t < (nxr + 1) * ld; t += blockDim.x) {
    const int a = t / ld, b = t - a * ld;
    int un = 0, nz = 0;
    if (a > 0 && b > 0) {
      un = robot_exploration_value(d, e, r, x0 + a - 1, y0 + b - 1) != 0.f;
      nz = robot_map_value(d, e, r, x0 + a - 1, y0 + b - 1) != 0.f;
    }
    s_un[t] = un, s_nz[t] = nz;
  }
  __syncthreads();
  for (int a = 1 + tid; a <= nxr; a += blockDim.x) {  // prefix along a row
    int su = 0, sn = 0;
    for (int b = 1; b <= nyr; ++b) su += s_un[a * ld + b], sn += s_nz[a * ld + b], s_un[a * ld + b] = su, s_nz[a * ld + b] = sn;
  }
  __syncthreads();
  for (int b = 1 + tid; b <= nyr; b += blockDim.x) {  // then down the columns
    int su = 0, sn = 0;
    for (int a = 1; a <= nxr; ++a) su += s_un[a * ld + b], sn += s_nz[a * ld + b], s_un[a * ld + b] = su, s_nz[a * ld + b] = sn;
  }
  __syncthreads();
  const double area = (double)(S * S);
  for (int t = tid; t < nx * ny; t += blockDim.x) {
    const int kx = t / ny, ky = t - kx * ny;
    const double dx = s_qx[kx] - px, dy = s_qy[ky] - py;
    const double dist = fmax(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), tsd);
    // map_utils.h:51-87 window of q, clipped, in integral-image coordinates
    const int a0 = max(s_cx[kx] - half, 0) - x0, a1 = min(s_cx[kx] + half, d.N) - x0;
    const int b0 = max(s_cy[ky] - half, 0) - y0, b1 = min(s_cy[ky] + half, d.N) - y0;
    const int un = s_un[a1 * ld + b1] - s_un[a0 * ld + b1] - s_un[a1 * ld + b0] + s_un[a0 * ld + b0];
    double benefit = (double)un;
    if ((double)un > 0.1 * area) {
      const int nz = s_nz[a1 * ld + b1] - s_nz[a0 * ld + b1] - s_nz[a1 * ld + b0] + s_nz[a0 * ld + b0];
      benefit += 2.0 * (double)nz;
    }
    s_val[t] = benefit / dist;
  }
  __syncthreads();
  if (tid == 0) {
    ExplHeap &h = s_heap;
    expl_select(h, s_val, nx * ny, K);  // kx outer, ky inner: the reference's order
    int count = 0;
    while (h.n > 0) {
      const int t = h.id[0], kx = t / ny, ky = t - kx * ny;
      o[count++] = s_qx[kx] - px, o[count++] = s_qy[ky] - py;
      expl_pop(h);
    }
    while (count < 2 * K) o[count++] = 0.0;
  }
}

// ---- sensor window marking ---------------------------------------------------------------------
// robot_model.h:98-122 (UpdateRobotMap / UpdateExplorationMap) and coverage_system.h:104-118
// (explored_idf / exploration windows).  One WARP per (robot, environment), MARK_WARPS of them per CTA;
// env_sel / robot_sel >= 0 select one.  A lane ORs one word of one window row: S x S cells are S x (S / 32 + 1)
// word updates.  A robot moves a cell or so per step, so all but one row / one bit column of its window is marked
// already: every word is READ first and only written when the OR would change it (the robot's own planes with a
// plain store -- no other warp writes them -- the shared system planes with an atomic).
// move_actions (optional, [E][R][2]): StepActions in the same launch -- the robot's warp first applies its action
// (robot_model.h:200-226, move_robot below), then marks the window at the new position.
constexpr int MARK_WARPS = 8;
__device__ inline void move_robot(const EnvDev &d, int e, int r, double ax, double ay, bool post, double &gx, double &gy);
__global__ void __launch_bounds__(MARK_WARPS * 32) k_mark_windows(EnvDev d, int env_sel, int robot_sel, int mark_robot, int mark_sys,
                                                                  const double *__restrict__ move_actions) {
  const int nr = robot_sel >= 0 ? 1 : d.R;
  const long wid = (long)blockIdx.x * MARK_WARPS + (threadIdx.x >> 5);
  const long total_w = (long)nr * (env_sel >= 0 ? 1 : d.E);
  if (wid >= total_w) return;
  const int e = env_sel >= 0 ? env_sel : (int)(wid / nr);
  const int r = robot_sel >= 0 ? robot_sel : (int)(wid % nr);
  const int lane = threadIdx.x & 31;
  const double *p = d.pos_g + ((size_t)e * d.R + r) * 2;
  double px, py;
  if (move_actions) {
    px = py = 0.0;
    if (lane == 0) move_robot(d, e, r, move_actions[2 * ((size_t)e * d.R + r)], move_actions[2 * ((size_t)e * d.R + r) + 1], true, px, py);
    px = __shfl_sync(0xffffffffu, px, 0), py = __shfl_sync(0xffffffffu, py, 0);
  } else {
    px = p[0], py = p[1];
  }
  const Window w = make_window(d.res, px, py, d.S, d.N);
  if (w.x1 <= w.x0 || w.y1 <= w.y0) return;
  const int w0 = w.y0 >> 5, nwords = ((w.y1 - 1) >> 5) - w0 + 1;
  const int total = (w.x1 - w.x0) * nwords;
  uint32_t *seen = robot_plane(d, d.seenb, e, r);
  uint32_t *xseen = d.xseenb ? robot_plane(d, d.xseenb, e, r) : nullptr;
  uint32_t *sys0 = sys_plane(d, e, 0), *sys1 = sys_plane(d, e, 1);
  if (mark_robot && d.upd_robot && lane == 0) {
    int *box = d.seen_box + ((size_t)e * d.R + r) * 4;  // only this warp touches this robot's box
    box[0] = min(box[0], w.x0), box[1] = max(box[1], w.x1), box[2] = min(box[2], w.y0), box[3] = max(box[3], w.y1);
  }
  const bool do_robot = mark_robot && d.upd_robot, do_expl = mark_robot && d.upd_expl, do_sys = mark_sys && d.upd_sys;
  for (int t = lane; t < total; t += 32) {
    const int a = t / nwords, wd = w0 + (t - a * nwords);
    const int lo = max(w.y0, wd * 32) - wd * 32, hi = min(w.y1, wd * 32 + 32) - wd * 32;  // bits [lo, hi)
    const uint32_t m = (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
    const size_t idx = (size_t)(w.x0 + a) * d.NW + wd;
    // all the reads first: four independent loads per word in flight
    const uint32_t o_seen = do_robot ? seen[idx] : m, o_x = do_expl ? xseen[idx] : m;
    const uint32_t o_s0 = do_sys ? sys0[idx] : m, o_s1 = do_sys ? sys1[idx] : m;
    if ((o_seen & m) != m) seen[idx] = o_seen | m;
    if ((o_x & m) != m) xseen[idx] = o_x | m;
    if ((o_s0 & m) != m) atomicOr(sys0 + idx, m);
    if ((o_s1 & m) != m) atomicOr(sys1 + idx, m);
  }
}

// ---- history (coverage_system.cpp:263-271) -----------------------------------------------------
__device__ inline void history_post_step(const EnvDev &d, int e, int r, double gx, double gy) {
  const size_t er = (size_t)e * d.R + r;
  double *h = d.hist + er * d.H * 2;
  int start = d.hist_start[er], len = d.hist_len[er];
  if (len > 0 && len == d.H) {
    // pop_front WITHOUT pushing, exactly as the reference does when the history is full
    start = (start + 1) % d.H;
    len -= 1;
  } else {
    int slot = (start + len) % d.H;
    h[2 * slot] = gx, h[2 * slot + 1] = gy;
    len += 1;
  }
  d.hist_start[er] = start, d.hist_len[er] = len;
}

// coverage_system.h:341-363
__device__ inline bool check_oscillation(const EnvDev &d, int e, int r) {
  if (!d.check_osc) return false;
  const size_t er = (size_t)e * d.R + r;
  const int len = d.hist_len[er], start = d.hist_start[er];
  if (len < 3) return false;
  const double *h = d.hist + er * d.H * 2;
  const int last = (start + len - 1) % d.H;
  const double lx = h[2 * last], ly = h[2 * last + 1];
  const int span = min(6, len - 1);
  int count = 0;
  for (int k = 0; k < span; ++k) {
    int s = (start + len - 1 - k) % d.H;
    double dx = h[2 * s] - lx, dy = h[2 * s + 1] - ly;
    if (sqrt(dx * dx + dy * dy) < kLargeEps) ++count;
  }
  return count > 2;
}

// robot_model.h:229-243
__device__ __forceinline__ void clamp_global(const EnvDev &d, double &gx, double &gy) {
  if (gx <= 0) gx = 0 + kLargeEps;
  if (gy <= 0) gy = 0 + kLargeEps;
  const double max_xy = d.N * d.res;
  if (gx >= max_xy) gx = max_xy - kLargeEps;
  if (gy >= max_xy) gy = max_xy - kLargeEps;
}

// Eigen normalized(): v / sqrt(v.v) when v.v > 0
__device__ __forceinline__ void eigen_normalize(double &x, double &y) {
  double z = x * x + y * y;
  if (z > 0) {
    double s = sqrt(z);
    x = x / s;
    y = y / s;
  }
}

// coverage_system.h:279-291 StepActions + robot_model.h:200-256 StepControl/SetRobotPosition.
// One thread per (env, robot).  limit[e] (may be null) = number of leading robots to move when the
// host found an invalid control (the reference returns before PostStepCommands in that case).
// One robot of StepActions (coverage_system.h:236-247 -> robot_model.h:200-226): speed = |action| capped at
// pMaxRobotSpeed, direction normalised, local position advanced, global position clamped into the map;
// post: PostStepCommands' history push.
__device__ inline void move_robot(const EnvDev &d, int e, int r, double ax, double ay, bool post, double &gx, double &gy) {
  const size_t er = (size_t)e * d.R + r;
  double sp = sqrt(ax * ax + ay * ay);
  double dx = ax, dy = ay;
  eigen_normalize(dx, dy);
  if (sp > d.max_speed) sp = d.max_speed;
  eigen_normalize(dx, dy);
  double lx = d.pos_l[2 * er], ly = d.pos_l[2 * er + 1];
  if (sp > kEps) {
    lx = lx + dx * sp * d.dt;
    ly = ly + dy * sp * d.dt;
  }
  gx = lx + d.pos_s[2 * er], gy = ly + d.pos_s[2 * er + 1];
  clamp_global(d, gx, gy);
  d.pos_l[2 * er] = gx - d.pos_s[2 * er];
  d.pos_l[2 * er + 1] = gy - d.pos_s[2 * er + 1];
  d.pos_g[2 * er] = gx;
  d.pos_g[2 * er + 1] = gy;
  if (post) history_post_step(d, e, r, gx, gy);
}

// robot_model.h:228-286 SetRobotPosition (local != 0) / SetGlobalRobotPosition; positions are
// R x 2 for env e (robot < 0) or one pair.  post != 0 pushes the history (PostStepCommands).
__global__ void k_set_positions(EnvDev d, int e, int robot, int local, const double *__restrict__ pos,
                                int post) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int count = robot < 0 ? d.R : 1;
  if (t >= count) return;
  const int r = robot < 0 ? t : robot;
  const size_t er = (size_t)e * d.R + r;
  double gx = pos[2 * t], gy = pos[2 * t + 1];
  if (local) gx = gx + d.pos_s[2 * er], gy = gy + d.pos_s[2 * er + 1];
  clamp_global(d, gx, gy);
  d.pos_l[2 * er] = gx - d.pos_s[2 * er];
  d.pos_l[2 * er + 1] = gy - d.pos_s[2 * er + 1];
  d.pos_g[2 * er] = gx;
  d.pos_g[2 * er + 1] = gy;
  if (post) history_post_step(d, e, r, gx, gy);
}

// RobotModel::Initialize (robot_model.h:124-153) positions part + InitSetup history push
// env_sel >= 0: only that environment, pos = its R x 2 positions; otherwise all, pos = E x R x 2
__global__ void k_init_positions(EnvDev d, const double *__restrict__ pos, int env_sel = -1) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (env_sel >= 0 ? d.R : d.E * d.R)) return;
  const double *src = pos + 2 * (size_t)idx;
  if (env_sel >= 0) idx += env_sel * d.R;
  const size_t er = (size_t)idx;
  double gx = src[0], gy = src[1];
  d.pos_s[2 * er] = gx, d.pos_s[2 * er + 1] = gy;
  d.pos_g[2 * er] = gx, d.pos_g[2 * er + 1] = gy;
  d.pos_l[2 * er] = 0, d.pos_l[2 * er + 1] = 0;
  d.hist_start[er] = 0, d.hist_len[er] = 0;
  d.seen_box[4 * er] = d.seen_box[4 * er + 2] = 0x7fffffff, d.seen_box[4 * er + 1] = d.seen_box[4 * er + 3] = 0;  // nothing seen yet
  history_post_step(d, idx / d.R, idx % d.R, gx, gy);
}

// ---- convex cell polygon: rectangle clipped by the bisector half-planes --------------------------
// Stands in for the CGAL arrangement face of voronoi.cpp:184-326; same vertex set.
// sites: x_j = sx[j * stride], y_j = sy[j * stride]
__device__ inline int clip_cell(const double *sx, const double *sy, int stride, int n, int k, double w, double h,
                                double (*poly)[2], double (*tmp)[2]) {
  int na = 4;
  poly[0][0] = 0, poly[0][1] = 0, poly[1][0] = w, poly[1][1] = 0;
  poly[2][0] = w, poly[2][1] = h, poly[3][0] = 0, poly[3][1] = h;
  const double kx = sx[k * stride], ky = sy[k * stride];
  for (int j = 0; j < n && na > 0; ++j) {
    if (j == k) continue;
    const double jx = sx[j * stride], jy = sy[j * stride];
    const double nx = jx - kx, ny = jy - ky;
    if (nx == 0 && ny == 0) continue;
    const double c = ((jx * jx + jy * jy) - (kx * kx + ky * ky)) / 2.;
    // most bisectors do not reach the cell: when every vertex is inside, the polygon stays as it is
    bool cuts = false;
    for (int v = 0; v < na; ++v) cuts = cuts || (nx * poly[v][0] + ny * poly[v][1] - c > 0);
    if (!cuts) continue;
    int nb = 0;
    for (int v = 0; v < na; ++v) {
      int u = (v + 1 == na) ? 0 : v + 1;
      double vx = poly[v][0], vy = poly[v][1], ux = poly[u][0], uy = poly[u][1];
      double dv = nx * vx + ny * vy - c, du = nx * ux + ny * uy - c;
      bool vin = dv <= 0, uin = du <= 0;
      if (vin && nb < CCB_MAX_CELL_VERTS) tmp[nb][0] = vx, tmp[nb][1] = vy, ++nb;
      if (vin != uin && nb < CCB_MAX_CELL_VERTS) {
        double t = dv / (dv - du);
        tmp[nb][0] = vx + t * (ux - vx), tmp[nb][1] = vy + t * (uy - vy), ++nb;
      }
    }
    int nc = 0;
    for (int v = 0; v < nb; ++v) {
      int pv = (v + nb - 1) % nb;
      if (nb > 1 && tmp[v][0] == tmp[pv][0] && tmp[v][1] == tmp[pv][1]) continue;
      poly[nc][0] = tmp[v][0], poly[nc][1] = tmp[v][1], ++nc;
    }
    na = nc;
  }
  return na;
}

struct CellsOut {
  double *mass;      // [E][R]
  double *centroid;  // [E][R][2]
  double *obj;       // [E][R]  sum |c - site|^2 v
  int *nverts;       // [E][R]
  double *verts;     // [E][R][CCB_MAX_CELL_VERTS][2]
};

// The same clipping with one WARP per cell: lane v holds vertex v, a half-plane is tested by all lanes at
// once (most bisectors do not reach the cell and cost a dozen instructions), a cutting half-plane
// is one prefix sum + one pass through the warp's shared-memory scratch.  Same expressions in the same
// order as clip_cell, hence the same vertices bit for bit.  sx/sy: n sites (any address space, read by
// all lanes alike); buf: this warp's scratch, 2 x 32 x 2 doubles; returns the vertex count with the
// polygon in buf[0..na), or -1 if a polygon outgrew 32 vertices (the caller then uses clip_cell).
constexpr int CLIPW_MAX = 32;
__device__ inline int clip_cell_warp(const double *sx, const double *sy, int stride, int n, int k, double w, double h,
                                     double (*buf)[2]) {
  const int lane = threadIdx.x & 31;
  double (*poly)[2] = buf, (*tmp)[2] = buf + CLIPW_MAX;
  int na = 4;
  double vx = (lane == 1 || lane == 2) ? w : 0.0, vy = (lane == 2 || lane == 3) ? h : 0.0;
  const double kx = sx[k * stride], ky = sy[k * stride];
  // The sites 32 at a time: one coalesced load per chunk, then shuffles (a load per site inside this sequential
  // loop costs an L2 latency each: 60 us at R = 256).  A site farther from k than twice the polygon's
  // circumradius about k (plus a margin above the rounding error of the half-plane test below) cannot cut it:
  // the lanes test their own site of the chunk against that bound, and only the others go through the
  // sequential loop, in ascending order as before.  Skipping is exactly "the bisector does not reach the cell",
  // so the vertices stay bit for bit those of clip_cell; after each cut the bound shrinks and the test is repeated.
  const double margin = 1e-6 * (w + h);
  double reach2;  // (2 circumradius + margin)^2, warp-uniform
  auto update_reach = [&]() {
    double r2 = 0.0;
    if (lane < na) r2 = (vx - kx) * (vx - kx) + (vy - ky) * (vy - ky);
#pragma unroll
    for (int o = 16; o; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    const double rr = 2.0 * sqrt(r2) + margin;
    reach2 = rr * rr;
  };
  update_reach();
  for (int base = 0; base < n && na > 0; base += 32) {
    const int jj = min(base + lane, n - 1);
    const double cjx = sx[jj * stride], cjy = sy[jj * stride];
    const double d2 = (cjx - kx) * (cjx - kx) + (cjy - ky) * (cjy - ky);
    unsigned todo = __ballot_sync(0xffffffffu, base + lane < n && base + lane != k && d2 <= reach2);
    while (todo && na > 0) {
      const int jl = __ffs(todo) - 1;
      todo &= todo - 1;
      const double jx = __shfl_sync(0xffffffffu, cjx, jl), jy = __shfl_sync(0xffffffffu, cjy, jl);
      const double nx = jx - kx, ny = jy - ky;
      if (nx == 0 && ny == 0) continue;
      const double c = ((jx * jx + jy * jy) - (kx * kx + ky * ky)) / 2.;
      const double dv = nx * vx + ny * vy - c;
      const bool act = lane < na;
      if (!__any_sync(0xffffffffu, act && dv > 0)) continue;  // the bisector does not reach the cell
      const int u = (lane + 1 == na) ? 0 : lane + 1;
      const double ux = __shfl_sync(0xffffffffu, vx, u & 31), uy = __shfl_sync(0xffffffffu, vy, u & 31);
      const double du = __shfl_sync(0xffffffffu, dv, u & 31);
      const bool vin = act && dv <= 0, uin = du <= 0;
      const bool cross = act && ((dv <= 0) != uin);
      const int cnt = (vin ? 1 : 0) + (cross ? 1 : 0);
      int incl = cnt;
  #pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const int nb = __shfl_sync(0xffffffffu, incl, 31);
      if (nb > CLIPW_MAX) return -1;
      int off = incl - cnt;
      if (vin) tmp[off][0] = vx, tmp[off][1] = vy, ++off;
      if (cross) {
        const double t = dv / (dv - du);
        tmp[off][0] = vx + t * (ux - vx), tmp[off][1] = vy + t * (uy - vy);
      }
      __syncwarp();
      // drop a vertex equal to its predecessor (cyclically), keep the order
      bool keep = false;
      double tx = 0, ty = 0;
      if (lane < nb) {
        tx = tmp[lane][0], ty = tmp[lane][1];
        const int pv = lane == 0 ? nb - 1 : lane - 1;
        keep = !(nb > 1 && tx == tmp[pv][0] && ty == tmp[pv][1]);
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = __popc(km & ((1u << lane) - 1u));
        poly[pos][0] = tx, poly[pos][1] = ty;
      }
      na = __popc(km);
      __syncwarp();
      vx = lane < na ? poly[lane][0] : 0.0, vy = lane < na ? poly[lane][1] : 0.0;
      __syncwarp();
      update_reach();
      todo &= __ballot_sync(0xffffffffu, d2 <= reach2);
    }
  }
  if (lane < na) poly[lane][0] = vx, poly[lane][1] = vy;
  __syncwarp();
  return na;
}

constexpr int FIN_WARPS = 8;

// Finalise the global Voronoi cells of every environment from the reduced moments
// (voronoi.cpp:41-53 ComputeFinalCentroid incl. the vertex-mean fallback; voronoi.h:89-95 sums).
// One warp per cell; the polygon is only built for cells without mass (or when the caller wants the
// vertex lists).  grid = ceil(E R / FIN_WARPS), block = 32 FIN_WARPS.
__global__ void __launch_bounds__(FIN_WARPS * 32) k_finalize_cells(EnvDev d, const double *__restrict__ sums, CellsOut out,
                                                                   int want_polys) {
  __shared__ double s_buf[FIN_WARPS][2 * CLIPW_MAX][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t er = (size_t)blockIdx.x * FIN_WARPS + warp;
  if (er >= (size_t)d.E * d.R) return;
  const int e = (int)(er / d.R), r = (int)(er - (size_t)e * d.R);
  const double *pos = d.pos_g + (size_t)e * d.R * 2;
  const double *s = sums + er * 4;
  const double m = s[0];
  const double px = pos[2 * r], py = pos[2 * r + 1];
  const double res = d.res;
  const double mx = res * s[1], my = res * s[2];
  double cx = 0, cy = 0;
  int nv = 0;
  if (m < kEps || want_polys) {  // warp-uniform
    // box = (pWorldMapSize, pWorldMapSize), not scaled by the resolution (coverage_system.h:373-375)
    double (*poly)[2] = s_buf[warp];
    nv = clip_cell_warp(pos, pos + 1, 2, d.R, r, (double)d.N, (double)d.N, poly);
    if (nv >= 0) {
      double ax = 0, ay = 0;
      for (int v = 0; v < nv; ++v) ax += poly[v][0], ay += poly[v][1];
      cx = nv > 0 ? ax / nv : px;
      cy = nv > 0 ? ay / nv : py;
      if (want_polys && lane < nv) {
        double *dst = out.verts + er * CCB_MAX_CELL_VERTS * 2;
        dst[2 * lane] = poly[lane][0], dst[2 * lane + 1] = poly[lane][1];
      }
    } else if (lane == 0) {  // more than 32 vertices: the sequential version
      double big[CCB_MAX_CELL_VERTS][2], tmp[CCB_MAX_CELL_VERTS][2];
      nv = clip_cell(pos, pos + 1, 2, d.R, r, (double)d.N, (double)d.N, big, tmp);
      double ax = 0, ay = 0;
      for (int v = 0; v < nv; ++v) ax += big[v][0], ay += big[v][1];
      cx = nv > 0 ? ax / nv : px;
      cy = nv > 0 ? ay / nv : py;
      if (want_polys) {
        double *dst = out.verts + er * CCB_MAX_CELL_VERTS * 2;
        for (int v = 0; v < nv; ++v) dst[2 * v] = big[v][0], dst[2 * v + 1] = big[v][1];
      }
    }
  }
  if (lane == 0) {
    out.mass[er] = m;
    out.obj[er] = res * res * s[3] - 2. * (px * mx + py * my) + (px * px + py * py) * m;
    if (!(m < kEps)) cx = mx / m, cy = my / m;
    out.centroid[2 * er] = cx, out.centroid[2 * er + 1] = cy;
    if (want_polys) out.nverts[er] = nv;
  }
}

// Goals of DecentralizedCVT (decentralized_cvt.h:85-143) from the per-robot window moments.
// One warp per robot; the clipped cell is only built when the robot's own cell has no mass.
// grid = ceil(E R / FIN_WARPS), block = 32 FIN_WARPS.
__global__ void __launch_bounds__(FIN_WARPS * 32) k_finalize_single(EnvDev d, const double *__restrict__ sums,
                                                                    double *__restrict__ goals) {
  __shared__ double s_buf[FIN_WARPS][2 * CLIPW_MAX][2];
  __shared__ double s_lx[FIN_WARPS][CCB_MAX_CELL_VERTS], s_ly[FIN_WARPS][CCB_MAX_CELL_VERTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t er = (size_t)blockIdx.x * FIN_WARPS + warp;
  if (er >= (size_t)d.E * d.R) return;
  const int e = (int)(er / d.R), r = (int)(er - (size_t)e * d.R);
  const double *pos = d.pos_g + (size_t)e * d.R * 2;
  const double *s = sums + er * 4;
  const double m = s[0], res = d.res;
  double gx, gy;
  if (m < kEps) {  // warp-uniform
    // vertex mean of (window rectangle ∩ bisector half-planes), in window coordinates
    const double sxr = pos[2 * r], syr = pos[2 * r + 1];
    const Window w = make_window(res, sxr, syr, d.L, d.N);
    const double tx = w.x0 * res, ty = w.y0 * res;
    double *lx = s_lx[warp], *ly = s_ly[warp];
    if (lane == 0) lx[0] = sxr - tx, ly[0] = syr - ty;
    int n = 1;
    // the robot, then its neighbours within communication range in ascending index (at most
    // CCB_MAX_CELL_VERTS sites, as the sequential version)
    for (int base = 0; base < d.R && n < CCB_MAX_CELL_VERTS; base += 32) {
      const int j = base + lane;
      bool keep = false;
      double jx = 0, jy = 0;
      if (j < d.R && j != r) {
        jx = pos[2 * j], jy = pos[2 * j + 1];
        const double dx = jx - sxr, dy = jy - syr;
        keep = sqrt(dx * dx + dy * dy) < d.comm_range;
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      const int slot = n + __popc(km & ((1u << lane) - 1u));
      if (keep && slot < CCB_MAX_CELL_VERTS) lx[slot] = jx - tx, ly[slot] = jy - ty;
      n = min(n + __popc(km), CCB_MAX_CELL_VERTS);
    }
    __syncwarp();
    double (*poly)[2] = s_buf[warp];
    int nv = clip_cell_warp(lx, ly, 1, n, 0, (double)(w.x1 - w.x0), (double)(w.y1 - w.y0), poly);
    double ax = 0, ay = 0;
    if (nv >= 0) {
      for (int v = 0; v < nv; ++v) ax += poly[v][0], ay += poly[v][1];
    } else {
      double big[CCB_MAX_CELL_VERTS][2], tmp[CCB_MAX_CELL_VERTS][2];
      nv = clip_cell(lx, ly, 1, n, 0, (double)(w.x1 - w.x0), (double)(w.y1 - w.y0), big, tmp);
      for (int v = 0; v < nv; ++v) ax += big[v][0], ay += big[v][1];
    }
    // voronoi.cpp:49: centroid_ += origin_shift (= -site0); decentralized_cvt.h:126 adds it back
    const double l0x = sxr - tx, l0y = syr - ty;
    const double cxs = (nv > 0 ? ax / nv : l0x) + (-l0x);
    const double cys = (nv > 0 ? ay / nv : l0y) + (-l0y);
    gx = cxs + l0x + tx;
    gy = cys + l0y + ty;
  } else {
    gx = res * s[1] / m;
    gy = res * s[2] / m;
  }
  if (lane == 0) {
    goals[2 * er] = gx, goals[2 * er + 1] = gy;
    if (gx < 0 || gx > d.N || gy < 0 || gy > d.N) atomicOr(d.status + e, 1);  // decentralized_cvt.h:127-141
  }
}

// clairvoyant_cvt.h:100-127 ComputeActions (identical in centralized_cvt.h / decentralized_cvt.h)
// goals_out (optional): the controller's own copy of the goals when `goals` are the cells' centroids
__global__ void k_actions(EnvDev d, const double *__restrict__ goals, double *__restrict__ actions,
                          int *__restrict__ converged, double *__restrict__ goals_out) {
  const int e = blockIdx.x;
  int all_zero = 1;
  for (int r = threadIdx.x; r < d.R; r += blockDim.x) {
    const size_t er = (size_t)e * d.R + r;
    double ax = 0, ay = 0;
    const double gx = goals[2 * er], gy = goals[2 * er + 1];
    if (goals_out) goals_out[2 * er] = gx, goals_out[2 * er + 1] = gy;
    double dx = gx - d.pos_g[2 * er], dy = gy - d.pos_g[2 * er + 1];
    double dist = sqrt(dx * dx + dy * dy);
    if (!(dist < kEps) && !check_oscillation(d, e, r)) {
      double speed = dist / d.dt;
      speed = speed < d.max_speed ? speed : d.max_speed;
      eigen_normalize(dx, dy);
      ax = speed * dx, ay = speed * dy;
      all_zero = 0;
    }
    actions[2 * er] = ax, actions[2 * er + 1] = ay;
  }
  int conv = __syncthreads_and(all_zero);
  if (threadIdx.x == 0) converged[e] = conv;
}

// objective = sum over robots of obj (voronoi.h:179-185), one CTA per environment
__global__ void k_objective(EnvDev d, const double *__restrict__ obj, double *__restrict__ out) {
  const int e = blockIdx.x;
  if (threadIdx.x == 0) {
    double o = 0;
    for (int r = 0; r < d.R; ++r) o = o + obj[(size_t)e * d.R + r];  // reference order
    out[e] = o;
  }
}

// ---- dense map materialisation (getters) ---------------------------------------------------------
// kind: ccb_map_kind.  N x N kinds write N*N floats; window kinds write dim*dim floats.
__global__ void k_materialize(EnvDev d, int e, int kind, int robot, float *__restrict__ out) {
  const size_t nn = cell_count(d);
  if (kind <= CCB_MAP_ROBOT) {
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nn; c += (size_t)gridDim.x * blockDim.x) {
      const int row = (int)(c / d.N), col = (int)(c - (size_t)row * d.N);
      float v;
      switch (kind) {
        case CCB_MAP_WORLD: v = d.world[(size_t)e * nn + c]; break;
        case CCB_MAP_SYSTEM: v = explored_value(d, e, row, col) - exploration_value(d, e, row, col); break;  // coverage_system.h:119
        case CCB_MAP_EXPLORED_IDF: v = explored_value(d, e, row, col); break;
        case CCB_MAP_EXPLORATION: v = exploration_value(d, e, row, col); break;
        default: v = robot_map_value(d, e, robot, row, col); break;
      }
      out[c] = v;
    }
    return;
  }
  const int dim = (kind == CCB_MAP_SENSOR_VIEW) ? d.S : d.L;
  const double *p = d.pos_g + ((size_t)e * d.R + robot) * 2;
  const Window w = make_window(d.res, p[0], p[1], dim, d.N);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < dim * dim; t += gridDim.x * blockDim.x) {
    int a = t / dim, b = t - a * dim;
    int gx = w.ox + a, gy = w.oy + b;
    bool in = gx >= 0 && gx < d.N && gy >= 0 && gy < d.N;
    size_t c = in ? (size_t)gx * d.N + gy : 0;
    if (!in) gx = gy = 0;
    float v;
    switch (kind) {
      case CCB_MAP_SENSOR_VIEW:  // robot_model.h:87-96: zero-padded world window
        v = (in && d.use_sensor) ? d.world[(size_t)e * nn + c] : 0.f;
        break;
      case CCB_MAP_LOCAL:  // robot_model.h:303-313
        v = in ? robot_map_value(d, e, robot, gx, gy) : 0.f;
        break;
      case CCB_MAP_OBSTACLE:  // robot_model.h:336-345
        v = in ? 0.f : 1.f;
        break;
      case CCB_MAP_ROBOT_EXPLORATION:  // robot_model.h:321-334: -1 padded
        v = in ? robot_exploration_value(d, e, robot, gx, gy) : -1.f;
        break;
      default:  // CCB_MAP_ROBOT_SYSTEM, robot_model.h:183-187: local - local_exploration
        v = (in ? robot_map_value(d, e, robot, gx, gy) : 0.f) - (in ? robot_exploration_value(d, e, robot, gx, gy) : -1.f);
        break;
    }
    out[t] = v;
  }
}

// coverage_system.h:557-571: [cells with exploration cleared, sum of explored IDF, sum of world]
__global__ void k_ratio_sums(EnvDev d, double *__restrict__ out /* [E][3] */) {
  const int e = blockIdx.y;
  const size_t nn = cell_count(d);
  double cnt = 0, ex = 0, tot = 0;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nn; c += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(c / d.N), col = (int)(c - (size_t)row * d.N);
    float w = d.world[(size_t)e * nn + c];
    cnt += plane_bit(sys_plane(d, e, 1), d.NW, row, col) ? 1.0 : 0.0;
    ex += plane_bit(sys_plane(d, e, 0), d.NW, row, col) ? (d.use_sensor ? (double)w : 0.0) : 0.0;
    tot += (double)w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 3 * e + 0, cnt);
    atomicAdd(out + 3 * e + 1, ex);
    atomicAdd(out + 3 * e + 2, tot);
  }
}

// robot_model.h:408-430 GetVoronoiFeatures: single-site "Voronoi" over the robot's clipped
// L-window of its own map.  One CTA per (robot, env).  out: [E][R][7].
// local != 0: CoverageSystem::GetLocalVoronoiFeatures (coverage_system.cpp:648-676): the same window and
// map, but only the robot's own cell among itself and its neighbours within communication range (the
// stated rule: a cell is lost when some neighbour is strictly nearer); dynamic smem: 2 R doubles.
__global__ void k_voronoi_features(EnvDev d, double *__restrict__ out, int local) {
  extern __shared__ double s_nb[];  // neighbour coordinates (local mode)
  __shared__ int s_nnb;
  const int r = blockIdx.x, e = blockIdx.y;
  const size_t er = (size_t)e * d.R + r;
  const double *p = d.pos_g + er * 2;
  const double res = d.res;
  const Window w = make_window(res, p[0], p[1], d.L, d.N);
  const double tx = w.x0 * res, ty = w.y0 * res;
  const double s0x = p[0] - tx, s0y = p[1] - ty;  // site in window coordinates
  const int ww = w.x1 - w.x0, wh = w.y1 - w.y0;
  if (threadIdx.x == 0) {
    int n = 0;
    if (local) {
      const double *pos = d.pos_g + (size_t)e * d.R * 2;
      for (int j = 0; j < d.R; ++j) {
        if (j == r) continue;
        const double dx = pos[2 * j] - p[0], dy = pos[2 * j + 1] - p[1];
        if (sqrt(dx * dx + dy * dy) < d.comm_range) s_nb[2 * n] = pos[2 * j], s_nb[2 * n + 1] = pos[2 * j + 1], ++n;
      }
    }
    s_nnb = n;
  }
  __syncthreads();
  const int nnb = s_nnb;
  double m = 0, cx = 0, cy = 0, sd = 0, sd2 = 0;
  for (int t = threadIdx.x; t < ww * wh; t += blockDim.x) {
    int a = t / wh, b = t - a * wh;
    float vf = robot_map_value(d, e, r, w.x0 + a, w.y0 + b);
    if (vf == 0.f) continue;
    if (nnb > 0) {
      const double ccx = __dmul_rn(__dadd_rn((double)(w.x0 + a), 0.5), res), ccy = __dmul_rn(__dadd_rn((double)(w.y0 + b), 0.5), res);
      const double ex = __dsub_rn(ccx, p[0]), ey = __dsub_rn(ccy, p[1]);
      const double own = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
      bool lost = false;
      for (int k = 0; k < nnb && !lost; ++k) {
        const double fx = __dsub_rn(ccx, s_nb[2 * k]), fy = __dsub_rn(ccy, s_nb[2 * k + 1]);
        lost = __dadd_rn(__dmul_rn(fx, fx), __dmul_rn(fy, fy)) < own;
      }
      if (lost) continue;
    }
    double v = vf;
    // voronoi.h:89-95 with origin_shift = -site0 (voronoi.h:171): pt is shifted, then (pt - site)
    // shifts a second time -- the reference's quirk, kept.
    double ptx = (a * res + res / 2.) - s0x, pty = (b * res + res / 2.) - s0y;
    double ddx = ptx - s0x, ddy = pty - s0y;
    double d2 = ddx * ddx + ddy * ddy;
    m += v, cx += ptx * v, cy += pty * v, sd += sqrt(d2) * v, sd2 += d2 * v;
  }
  __shared__ double s_red[5][32];
  double vals[5] = {m, cx, cy, sd, sd2};
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    double x = vals[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) s_red[q][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nw = blockDim.x >> 5;
    for (int q = 0; q < 5; ++q) {
      double x = 0;
      for (int k = 0; k < nw; ++k) x += s_red[q][k];
      vals[q] = x;
    }
    double *o = out + er * 7;
    if (vals[0] < kEps) {
      if (nnb > 0) {
        // vertex mean of (window rectangle cut by the bisectors) + origin_shift (voronoi.cpp:41-50)
        double lx[CCB_MAX_CELL_VERTS], ly[CCB_MAX_CELL_VERTS];
        int n = 1;
        lx[0] = s0x, ly[0] = s0y;
        for (int k = 0; k < nnb && n < CCB_MAX_CELL_VERTS; ++k) lx[n] = s_nb[2 * k] - tx, ly[n] = s_nb[2 * k + 1] - ty, ++n;
        double poly[CCB_MAX_CELL_VERTS][2], tmp[CCB_MAX_CELL_VERTS][2];
        const int nv = clip_cell(lx, ly, 1, n, 0, (double)ww, (double)wh, poly, tmp);
        double ax = 0, ay = 0;
        for (int v = 0; v < nv; ++v) ax += poly[v][0], ay += poly[v][1];
        o[0] = (nv > 0 ? ax / nv : s0x) - s0x, o[1] = (nv > 0 ? ay / nv : s0y) - s0y;
      } else {
        // vertex mean of the window rectangle + origin_shift (voronoi.cpp:41-50)
        o[0] = (0 + ww + ww + 0) / 4. - s0x, o[1] = (0 + 0 + wh + wh) / 4. - s0y;
      }
    } else {
      o[0] = vals[1] / vals[0], o[1] = vals[2] / vals[0];
    }
    o[2] = vals[0], o[3] = vals[4], o[4] = 0, o[5] = vals[3], o[6] = 0;  // voronoi.h:73-78
  }
}

}  // namespace ccb
