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


// ---- specialised path for L = 8 C (LPAC: 256 -> 32), N % 4 == 0 -----------------------------------
// One CTA per (robot, env), 8 warps, lane = output column oy.  The robot's bit plane says where its
// local map is non-zero, and after a few hundred steps that is still a small part of the 256 x 256
// window, so the kernel is organised around the rows that have a seen cell:
//   1. the 264 x 9 mask words of the window rows a in [-4, 260) are staged in shared memory, clipped to
//      the window's columns (2.4 KB of global reads instead of 270 KB of floats), and the rows with a
//      set bit are compacted into a list;
//   2. width pass: the warps deal the listed rows between them in batches of CNN8_BATCH rows whose
//      world values are all requested before the first is used (one HBM latency per batch instead of
//      one per row).  A lane owns the 8 window columns 8 oy .. 8 oy + 7 of its output column: one byte
//      of the row's mask says which are seen, the values arrive as 16-byte loads (the window origin is
//      arbitrary, so the loads start at the aligned column below it and the misalignment S, uniform
//      over the CTA, is a template parameter), the 4 + 4 neighbour taps come from the adjacent lanes by
//      shuffle, and the 16-tap sum goes to s_tmp[a][oy];
//   3. height pass: warp w owns output rows 4w .. 4w+3 and sums its 16 taps in ascending row order,
//      the order torchvision's separable antialias kernel uses.
// The obstacle channel needs no map at all: obstacle(a, b) = 1 - in_x(a) in_y(b), so its width pass
// has only two possible values per output column.  wfull[o][k] = weight of input index 8 o - 4 + k
// (zero outside torch's tap range).
constexpr int CNN8_THREADS = 256;
constexpr int CNN8_WARPS = CNN8_THREADS / 32;
constexpr int CNN8_ROWS = 264;  // window rows -4 .. 259
constexpr int CNN8_WORDS = 9;   // mask words a 256-cell window row can touch
constexpr int CNN8_BATCH = 3;   // rows whose loads are in flight together

struct Cnn8Shared {
  float tmp[CNN8_ROWS][32];
  uint32_t mask[CNN8_ROWS * CNN8_WORDS + 1];
  float w[32][16];
  uint32_t flags[(CNN8_ROWS + 31) / 32 + 1];  // bit a + 4: window row a has a non-zero cell
  int act[CNN8_ROWS];
  int nact;
};
constexpr size_t CNN8_SMEM_BYTES = sizeof(Cnn8Shared);

// width pass over the listed rows; S = (window origin column) & 3
template <int S>
__device__ __forceinline__ void cnn8_width_pass(const EnvDev &d, Cnn8Shared &sh, const float *__restrict__ world, int ox_map,
                                                int oy_map, int wbase, const float (&wc)[16]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nact = sh.nact;
  const int gy0 = oy_map + 8 * lane;  // map column of this lane's first cell
  const int gyb = gy0 - S;            // 16-byte aligned (N % 4 == 0)
  constexpr int NV = S == 0 ? 2 : 3;  // aligned 4-cell groups covering the 8 cells
  bool grp_in[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) grp_in[j] = gyb + 4 * j >= 0 && gyb + 4 * j + 3 < d.N;
  unsigned inmap = 0u;  // bit i: column gy0 + i is inside the map
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (gy0 + i >= 0 && gy0 + i < d.N) inmap |= 1u << i;
  // the lane's 8 mask bits start at bit `off` of the row's 9 words
  const int off = gy0 - 32 * wbase;
  const int wi = off >> 5, shf = off & 31;
  const float fill = d.fill;
  const bool dense = fill != 0.f;
#pragma unroll 1
  for (int base = warp * CNN8_BATCH; base < nact; base += CNN8_WARPS * CNN8_BATCH) {
    float4 v[CNN8_BATCH][NV];
    unsigned vis[CNN8_BATCH];
#pragma unroll
    for (int b = 0; b < CNN8_BATCH; ++b) {
      vis[b] = 0u;
#pragma unroll
      for (int j = 0; j < NV; ++j) v[b][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + b < nact) {  // warp-uniform
        const int ai = sh.act[base + b];
        const uint32_t lo = sh.mask[ai * CNN8_WORDS + wi], hi = sh.mask[ai * CNN8_WORDS + wi + 1];
        vis[b] = __funnelshift_r(lo, hi, shf) & 0xffu;
        if (vis[b] != 0u && d.use_sensor) {
          const float4 *src = reinterpret_cast<const float4 *>(world + (size_t)(ox_map + ai - 4) * d.N + gyb);
#pragma unroll
          for (int j = 0; j < NV; ++j)
            if (grp_in[j]) v[b][j] = __ldg(src + j);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < CNN8_BATCH; ++b) {
      if (base + b >= nact) break;  // warp-uniform
      float x[4 * NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) x[4 * j] = v[b][j].x, x[4 * j + 1] = v[b][j].y, x[4 * j + 2] = v[b][j].z, x[4 * j + 3] = v[b][j].w;
      float c[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float unseen = (dense && ((inmap >> i) & 1u)) ? fill : 0.f;
        c[i] = ((vis[b] >> i) & 1u) ? x[S + i] : unseen;
      }
      // taps q = 0..15 are window columns 8 lane - 4 + q: the left neighbour's c[4..7], c[0..7], the
      // right neighbour's c[0..3]; beyond the window (lanes 0 / 31) the padding is zero
      float tmp = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float l = __shfl_up_sync(0xffffffffu, c[4 + q], 1);
        if (lane == 0) l = 0.f;
        tmp = fmaf(wc[q], l, tmp);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) tmp = fmaf(wc[4 + q], c[q], tmp);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float r = __shfl_down_sync(0xffffffffu, c[q], 1);
        if (lane == 31) r = 0.f;
        tmp = fmaf(wc[12 + q], r, tmp);
      }
      sh.tmp[sh.act[base + b]][lane] = tmp;
    }
  }
}

__global__ void __launch_bounds__(CNN8_THREADS, 3)
k_cnn_fused8(EnvDev d, const float *__restrict__ wfull /* [32][16] */, float *__restrict__ out,
             size_t robot_stride, size_t local_off, size_t obst_off) {
  extern __shared__ __align__(16) unsigned char cnn8_smem[];  // CNN8_SMEM_BYTES (opt-in: above 48 KB)
  Cnn8Shared &sh = *reinterpret_cast<Cnn8Shared *>(cnn8_smem);
  const int r = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t er = (size_t)e * d.R + r;
  const size_t nn = cell_count(d);
  for (int t = tid; t < 32 * 16; t += CNN8_THREADS) sh.w[t >> 4][t & 15] = wfull[t];
  if (tid < (CNN8_ROWS + 31) / 32 + 1) sh.flags[tid] = 0u;
  if (tid == 0) sh.nact = 0, sh.mask[CNN8_ROWS * CNN8_WORDS] = 0u;
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], 256, d.N);
  const float *world = d.world + (size_t)e * nn;
  const uint32_t *plane = robot_plane(d, d.seenb, e, r);
  const bool dense = d.fill != 0.f;  // unseen cells carry pUnknownImportance: every in-map row counts
  __syncthreads();

  // ---- 1. mask words, clipped to the window's map columns [cy0, cy1); all loads issued first ------
  const int wbase = w.oy >> 5;  // arithmetic shift: floor for negative origins
  const int cy0 = max(w.oy, 0), cy1 = min(w.oy + 256, d.N);
  constexpr int NMW = (CNN8_ROWS * CNN8_WORDS + CNN8_THREADS - 1) / CNN8_THREADS;
  uint32_t mw[NMW];
#pragma unroll
  for (int it = 0; it < NMW; ++it) {
    const int t = tid + it * CNN8_THREADS;
    const int ai = t / CNN8_WORDS, j = t - ai * CNN8_WORDS;  // ai = a + 4
    const int a = ai - 4;
    const int gx = w.ox + a, wd = wbase + j;
    uint32_t m = 0u;
    if (t < CNN8_ROWS * CNN8_WORDS && a >= 0 && a < 256 && gx >= 0 && gx < d.N && wd >= 0 && wd < d.NW) {
      const int cb = wd << 5;  // map column of bit 0
      const int lo = min(max(cy0 - cb, 0), 32), hi = min(max(cy1 - cb, 0), 32);
      if (hi > lo) {
        const uint32_t keep = ((hi - lo >= 32) ? 0xffffffffu : ((1u << (hi - lo)) - 1u)) << lo;
        m = __ldg(plane + (size_t)gx * d.NW + wd) & keep;
        if (dense) atomicOr(&sh.flags[ai >> 5], 1u << (ai & 31));
      }
    }
    mw[it] = m;
  }
#pragma unroll
  for (int it = 0; it < NMW; ++it) {
    const int t = tid + it * CNN8_THREADS;
    if (t < CNN8_ROWS * CNN8_WORDS) {
      sh.mask[t] = mw[it];
      const int ai = t / CNN8_WORDS;
      if (mw[it]) atomicOr(&sh.flags[ai >> 5], 1u << (ai & 31));
    }
  }
  __syncthreads();
  for (int ai = tid; ai < CNN8_ROWS; ai += CNN8_THREADS)
    if ((sh.flags[ai >> 5] >> (ai & 31)) & 1u) sh.act[atomicAdd(&sh.nact, 1)] = ai;
  __syncthreads();

  // ---- 2. width pass over the listed rows ----------------------------------------------------------
  float wc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) wc[q] = sh.w[lane][q];
  switch (w.oy & 3) {
    case 0: cnn8_width_pass<0>(d, sh, world, w.ox, w.oy, wbase, wc); break;
    case 1: cnn8_width_pass<1>(d, sh, world, w.ox, w.oy, wbase, wc); break;
    case 2: cnn8_width_pass<2>(d, sh, world, w.ox, w.oy, wbase, wc); break;
    default: cnn8_width_pass<3>(d, sh, world, w.ox, w.oy, wbase, wc); break;
  }

  // width pass of the obstacle indicator for this lane's output column
  float oy_out = 0.f, oy_all = 0.f;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int gy = w.oy + 8 * lane - 4 + q;
    const bool in_y = gy >= 0 && gy < d.N;
    oy_all = fmaf(wc[q], 1.f, oy_all);
    oy_out = fmaf(wc[q], in_y ? 0.f : 1.f, oy_out);
  }
  __syncthreads();

  // ---- 3. height pass: output rows 4 warp .. 4 warp + 3, taps in ascending row order --------------
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = warp * 4 + i;
    float acc = 0.f, acc_obs = 0.f;
    const int ai0 = 8 * ox;  // a + 4 of tap 0
    // the 16 row flags of this output row: bits ai0 .. ai0 + 15
    const uint32_t fl = __funnelshift_r(sh.flags[ai0 >> 5], sh.flags[(ai0 >> 5) + 1], ai0 & 31) & 0xffffu;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int a = ai0 - 4 + k;
      const float wr = sh.w[ox][k];  // zero outside torch's tap range
      const int gx = w.ox + a;
      const bool in_x = a >= 0 && a < 256 && gx >= 0 && gx < d.N;
      acc_obs = fmaf(wr, in_x ? oy_out : oy_all, acc_obs);
      if ((fl >> k) & 1u) acc = fmaf(wr, sh.tmp[ai0 + k][lane], acc);
    }
    float *dst = out + er * robot_stride + (size_t)ox * 32 + lane;
    dst[local_off] = acc;
    dst[obst_off] = acc_obs;
  }
}

}  // namespace ccb
