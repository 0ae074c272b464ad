// obs.cuh -- LPAC observation tensors produced directly on the device.
//
// Replaces the per-robot Python loop of reference python/coverage_control/coverage_env_utils.py:
// get_raw_local_maps (:97-117), get_raw_obstacle_maps (:119-141), get_communication_maps
// (:143-191), resize_maps (:200-225, torchvision antialias bilinear) and get_maps (:227-277),
// which call RobotModel::GetRobotLocalMap / GetObstacleMap (robot_model.h:303-345) R times per
// step, copy every L x L map twice and resize on the CPU.
#pragma once

#include "common.cuh"
#include "maps.cuh"

namespace ccb {

// E x R x L x L raw local maps (kind CCB_MAP_LOCAL) or obstacle maps (CCB_MAP_OBSTACLE).
// grid = (L, R, E): one CTA per output row, coalesced along y.
__global__ void k_raw_window_maps(EnvDev d, int kind, float *__restrict__ out) {
  const int a = blockIdx.x, r = blockIdx.y, e = blockIdx.z;
  const size_t er = (size_t)e * d.R + r;
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], d.L, d.N);
  const int gx = w.ox + a;
  float *dst = out + (er * d.L + a) * d.L;
  for (int b = threadIdx.x; b < d.L; b += blockDim.x) {
    const int gy = w.oy + b;
    const bool in = gx >= 0 && gx < d.N && gy >= 0 && gy < d.N;
    float v;
    if (kind == CCB_MAP_OBSTACLE) v = in ? 0.f : 1.f;
    else v = in ? robot_map_value(d, e, r, gx, gy) : 0.f;
    dst[b] = v;
  }
}

// Communication maps with the semantics LPAC consumes (coverage_env_utils.py:143-191): float32
// arithmetic, torch.round = round-half-even, duplicates summed.  out: E x R x 2 x C x C, or a
// channel slice of a wider tensor (chan_stride = floats between robots, chan_off = first channel).
// One CTA per (robot, env).
__global__ void k_comm_maps(EnvDev d, int C, float *__restrict__ out, size_t robot_stride, size_t chan_off) {
  const int r = blockIdx.x, e = blockIdx.y;
  const size_t er = (size_t)e * d.R + r;
  float *ox = out + er * robot_stride + chan_off;
  float *oy = ox + (size_t)C * C;
  for (int t = threadIdx.x; t < 2 * C * C; t += blockDim.x) ox[t] = 0.f;
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double *pos = d.pos_g + (size_t)e * d.R * 2;
  const float denom = (float)(d.comm_range * d.res * 2.0);
  const float center = (float)(C / 2.0 - d.res / 2.0);
  const float cr = (float)d.comm_range;
  const double sx = pos[2 * r], sy = pos[2 * r + 1];
  for (int j = 0; j < d.R; ++j) {
    if (j == r) continue;
    // coverage_system.cpp:274-291: relative position p_j - p_r (stored negated for j < r)
    double dx, dy;
    if (j > r) dx = pos[2 * j] - sx, dy = pos[2 * j + 1] - sy;
    else dx = -(sx - pos[2 * j]), dy = -(sy - pos[2 * j + 1]);
    if (!(sqrt(dx * dx + dy * dy) < d.comm_range)) continue;
    const float fx = (float)dx, fy = (float)dy;
    const float ix = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(fx, (float)C), denom), center));
    const float iy = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(fy, (float)C), denom), center));
    const int xi = (int)ix, yi = (int)iy;
    if (xi < 0 || xi >= C || yi < 0 || yi >= C) continue;
    ox[xi * C + yi] += __fdiv_rn(fx, cr);
    oy[xi * C + yi] += __fdiv_rn(fy, cr);
  }
}

// Separable antialiased bilinear down-sampling weights (what torchvision resize(antialias=True)
// applies): for output index o the taps are input indices [xmin[o], xmin[o]+xsize[o]).
struct ResizeTable {
  const int *xmin;    // [C]
  const int *xsize;   // [C]
  const float *w;     // [C][kmax]
  int kmax;
};

// Fused crop -> mask -> antialias resize for the local-map and obstacle channels, writing
// straight into the E x R x nch x C x C CNN input.  grid = (C, R, E): one CTA per output row ox.
//   pass 1 (width, as torch does first): tmp[a][oy] = sum_t w[oy][t] * in[xmin[ox]+a][ymin[oy]+t]
//   pass 2 (height):                      out[ox][oy] = sum_a w[ox][a] * tmp[a][oy]
// Dynamic smem: kmax * L floats (input rows) + kmax * C floats (tmp).
__global__ void k_cnn_local_obstacle(EnvDev d, int C, ResizeTable tab, float *__restrict__ out,
                                     size_t robot_stride, size_t local_off, size_t obst_off) {
  extern __shared__ float s_obs[];
  float *s_in = s_obs;                       // [kmax][L]
  float *s_tmp = s_obs + (size_t)tab.kmax * d.L;  // [kmax][C]
  const int ox = blockIdx.x, r = blockIdx.y, e = blockIdx.z;
  const size_t er = (size_t)e * d.R + r;
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], d.L, d.N);
  const int a0 = tab.xmin[ox], na = tab.xsize[ox];
  const float *wrow = tab.w + (size_t)ox * tab.kmax;

  for (int pass = 0; pass < 2; ++pass) {  // 0: local map, 1: obstacle map
    for (int t = threadIdx.x; t < na * d.L; t += blockDim.x) {
      const int a = t / d.L, b = t - a * d.L;
      const int gx = w.ox + a0 + a, gy = w.oy + b;
      const bool in = gx >= 0 && gx < d.N && gy >= 0 && gy < d.N;
      float v;
      if (pass == 0) v = in ? robot_map_value(d, e, r, gx, gy) : 0.f;
      else v = in ? 0.f : 1.f;
      s_in[a * d.L + b] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < na * C; t += blockDim.x) {
      const int a = t / C, oy = t - a * C;
      const int b0 = tab.xmin[oy], nb = tab.xsize[oy];
      const float *wc = tab.w + (size_t)oy * tab.kmax;
      float acc = 0.f;
      for (int q = 0; q < nb; ++q) acc = fmaf(wc[q], s_in[a * d.L + b0 + q], acc);
      s_tmp[a * C + oy] = acc;
    }
    __syncthreads();
    float *dst = out + er * robot_stride + (pass == 0 ? local_off : obst_off) + (size_t)ox * C;
    for (int oy = threadIdx.x; oy < C; oy += blockDim.x) {
      float acc = 0.f;
      for (int a = 0; a < na; ++a) acc = fmaf(wrow[a], s_tmp[a * C + oy], acc);
      dst[oy] = acc;
    }
    __syncthreads();
  }
}


// ---- specialised path for L = 8 C (LPAC: 256 -> 32) ----------------------------------------------
// One CTA per (robot, env), 8 warps; warp w owns output rows 4w .. 4w+3, i.e. window rows
// 32 w - 4 .. 32 w + 35 (40 rows, each used by at most two output rows), and lane = oy.
// The robot's bit plane says where its local map is non-zero: the warp first stages the 40 x 9 mask
// words of its rows in shared memory (1.4 KB instead of 40 KB of floats), then only rows with a set
// bit load world values (8 predicated loads per lane), go through a padded shared-memory line (one
// pad word per 8 cells -> the 16-tap reads of the 32 lanes hit 32 different banks) and the width
// pass; the height pass accumulates in registers in ascending row order, the order torchvision's
// separable antialias kernel uses.  The obstacle channel needs no map at all:
// obstacle(a, b) = 1 - in_x(a) in_y(b), so its width pass has only two possible values per output
// column.  wfull[o][k] = weight of input index 8 o - 4 + k (zero outside torch's tap range).
constexpr int CNN8_THREADS = 256;
constexpr int CNN8_LINE = 304;
constexpr int CNN8_ROWS = 40;   // window rows per warp
constexpr int CNN8_WORDS = 9;   // mask words a 256-cell window row can touch
// window column b in [-4, 260) -> word of the padded line (one pad word per 8 cells)
__device__ __forceinline__ int cnn8_pos(int b) { return (b + 8) + ((b + 8) >> 3); }

__global__ void __launch_bounds__(CNN8_THREADS, 3)
k_cnn_fused8(EnvDev d, const float *__restrict__ wfull /* [32][16] */, float *__restrict__ out,
             size_t robot_stride, size_t local_off, size_t obst_off) {
  __shared__ float s_line[CNN8_THREADS / 32][CNN8_LINE];
  __shared__ float s_w[32][16];
  __shared__ uint32_t s_mask[CNN8_THREADS / 32][CNN8_ROWS * CNN8_WORDS];
  const int r = blockIdx.x, e = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t er = (size_t)e * d.R + r;
  const size_t nn = cell_count(d);
  for (int t = threadIdx.x; t < 32 * 16; t += CNN8_THREADS) s_w[t >> 4][t & 15] = wfull[t];
  float *line = s_line[warp];
  uint32_t *smask = s_mask[warp];
  if (lane < 4) line[cnn8_pos(lane - 4)] = 0.f;         // taps left of the window
  else if (lane < 8) line[cnn8_pos(252 + lane)] = 0.f;  // taps right of the window (b = 256..259)
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], 256, d.N);
  const float *world = d.world + (size_t)e * nn;
  const uint32_t *plane = robot_plane(d, d.seenb, e, r);

  // This warp owns output rows ox0 .. ox0+3; they need window rows a0 = 8 ox0 - 4 .. 8 ox0 + 35.
  const int ox0 = warp * 4;
  const int a0 = 8 * ox0 - 4;
  // mask words of the 40 rows: word j of a row covers map columns 32 (wbase + j) .. + 31
  const int wbase = w.oy >> 5;  // arithmetic shift: floor for negative origins
  for (int t = lane; t < CNN8_ROWS * CNN8_WORDS; t += 32) {
    const int a = t / CNN8_WORDS, j = t - a * CNN8_WORDS;
    const int gx = w.ox + a0 + a, wd = wbase + j;
    uint32_t m = 0u;
    if (gx >= 0 && gx < d.N && wd >= 0 && wd < d.NW) m = __ldg(plane + (size_t)gx * d.NW + wd);
    smask[t] = m;
  }
  __syncthreads();

  // width pass of the obstacle indicator for this lane's output column
  float wc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) wc[q] = s_w[lane][q];
  float oy_out = 0.f, oy_all = 0.f;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int gy = w.oy + 8 * lane - 4 + q;
    const bool in_y = gy >= 0 && gy < d.N;
    oy_all = fmaf(wc[q], 1.f, oy_all);
    oy_out = fmaf(wc[q], in_y ? 0.f : 1.f, oy_out);
  }

  // Row a contributes to output row o = ox0 + i with weight s_w[o][a - (8 o - 4)] (two rows at most).
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc_obs[4] = {0.f, 0.f, 0.f, 0.f};
  const bool dense = d.fill != 0.f;  // unseen cells carry pUnknownImportance: nothing can be skipped
#pragma unroll 1
  for (int t = 0; t < CNN8_ROWS; ++t) {
    const int gx = w.ox + a0 + t;
    const bool in_x = gx >= 0 && gx < d.N;
    // weights of this row for the four output rows (zero outside torch's 16-tap range)
    float wr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = t - 8 * i;  // tap index of row a0 + t in output row ox0 + i
      wr[i] = (k >= 0 && k < 16) ? s_w[ox0 + i][k] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc_obs[i] = fmaf(wr[i], in_x ? oy_out : oy_all, acc_obs[i]);
    if (!in_x) continue;  // zero padding (warp-uniform)
    const uint32_t mine = lane < CNN8_WORDS ? smask[t * CNN8_WORDS + lane] : 0u;
    if (!dense && !__any_sync(0xffffffffu, mine != 0u)) continue;  // nothing seen in this row
    float vals[8];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int gy = w.oy + 32 * k + lane;
      const bool in = gy >= 0 && gy < d.N;
      const uint32_t word = smask[t * CNN8_WORDS + ((gy >> 5) - wbase)];
      float v = 0.f;
      if (in) {
        if ((word >> (gy & 31)) & 1u) v = d.use_sensor ? __ldg(world + (size_t)gx * d.N + gy) : 0.f;
        else v = d.fill;
      }
      vals[k] = v;
      any |= (v != 0.f);
    }
    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
      for (int k = 0; k < 8; ++k) line[cnn8_pos(32 * k + lane)] = vals[k];
      __syncwarp();
      float tmp = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) tmp = fmaf(wc[q], line[cnn8_pos(8 * lane - 4 + q)], tmp);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(wr[i], tmp, acc[i]);
      __syncwarp();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float *dst = out + er * robot_stride + (size_t)(ox0 + i) * 32 + lane;
    dst[local_off] = acc[i];
    dst[obst_off] = acc_obs[i];
  }
}

}  // namespace ccb
