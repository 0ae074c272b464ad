// One-shot all-reduce(SUM) of the per-robot moment partials over peer memory (NVLink 5 / NVSwitch).
//
// Row bands / robot shards of ONE world (BASELINE config 4) exchange E x R x 4 doubles per step -- 8 KB at
// R = 256: pure latency.  Every rank maps every peer's exchange buffer (symmetric allocation, the pointers come from
// the host plumbing).  One kernel per rank and step:
//   1. the rank stores its vector into slot [parity][rank] of EVERY peer's buffer (plain stores over NVLink),
//      fences system-wide and raises flag [parity][rank] = seq on every peer (st.release.sys);
//   2. it waits until its own flags [parity][0 .. G) have reached seq (ld.acquire.sys);
//   3. it adds its own slots IN RANK ORDER (loads that bypass L1: the lines were written by the peers)
//      -> every rank gets bit-identical sums whatever the arrival order (NCCL's ring / tree order is not fixed).
// seq lives in device memory and is advanced by the kernel, so the launch can sit in a replayed CUDA graph.
// parity = seq & 1 double-buffers slots and flags: a rank can be at most one exchange ahead of the slowest (it
// cannot pass step 2 of exchange s + 1 before every peer has finished reading exchange s), so exchange s + 2 never
// overwrites slots still being read.  No reference equivalent (the reference's Voronoi is single-process).
#pragma once
#include <cstdint>

namespace ccb {

constexpr int PEER_MAX = 16;
constexpr int PEER_THREADS = 512;
struct PeerArgs {
  unsigned char *buf[PEER_MAX];  // the exchange buffer of every rank as mapped HERE: [2][G][count] doubles, then [2][G] flags
  int rank, world;
  unsigned long long count;
  unsigned long long *state;  // local: [0] = number of exchanges done, [1] = sticky error (a peer never arrived)
};
__host__ __device__ inline size_t peer_flags_offset(int world, size_t count) {
  return (sizeof(double) * 2 * (size_t)world * count + 127) / 128 * 128;
}
__host__ __device__ inline size_t peer_buffer_bytes(int world, size_t count) {
  return peer_flags_offset(world, count) + sizeof(unsigned long long) * 2 * (size_t)world;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_cg_f64(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(PEER_THREADS) k_peer_allreduce(PeerArgs a, const double *__restrict__ in, double *__restrict__ out) {
  const int tid = threadIdx.x, G = a.world;
  const unsigned long long seq = a.state[0] + 1ull, par = seq & 1ull;
  const size_t n = a.count, slot = (size_t)(par * G + a.rank) * n, foff = peer_flags_offset(G, n);
  __syncthreads();  // everyone has read state[0] before thread 0 advances it below
  for (size_t i = tid; i < n; i += PEER_THREADS) {
    const double v = in[i];
    for (int p = 0; p < G; ++p) reinterpret_cast<double *>(a.buf[p])[slot + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < G) {
    st_release_sys(reinterpret_cast<unsigned long long *>(a.buf[tid] + foff) + par * G + a.rank, seq);
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(a.buf[a.rank] + foff) + par * G + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < seq) {
      if (clock64() - t0 > 6000000000ll) {  // ~3 s: a peer is gone; give up loudly instead of hanging the GPU
        a.state[1] = seq;
        break;
      }
      __nanosleep(20);
    }
  }
  __syncthreads();
  const double *my = reinterpret_cast<const double *>(a.buf[a.rank]) + (size_t)par * G * n;
  for (size_t i = tid; i < n; i += PEER_THREADS) {
    double s = 0.0;
    for (int p = 0; p < G; ++p) s = __dadd_rn(s, ld_cg_f64(my + (size_t)p * n + i));
    out[i] = s;
  }
  if (tid == 0) a.state[0] = seq;
}

}  // namespace ccb
