// ccb200.cu -- host side of the C ABI declared in include/ccb200.h.
//
// One translation unit: kernels live in idf.cuh / voronoi.cuh / maps.cuh / obs.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false ... (build.py).
// -fmad=false keeps every implicit a*b+c un-contracted so the fp64 host-visible arithmetic
// (kinematics, distances, clipping) rounds like the reference's x86-64 build; FMAs are written
// explicitly where they are wanted.
#include <algorithm>
#include <array>
#include <map>
#include <memory>

#include "../../include/ccb200_test.h"
#include "common.cuh"
#include "idf.cuh"
#include "maps.cuh"
#include "obs.cuh"
#include "voronoi.cuh"
#include "vglobal.cuh"
#include "vwindow.cuh"
#include "peer.cuh"

namespace ccb {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace ccb

using namespace ccb;

// =================================================================================================
// WorldIDF
// =================================================================================================
struct ccb_idf {
  ccb_params p;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<std::array<double, 6>> bnds;  // mean_x, mean_y, sigma_x, sigma_y, scale, rho
  struct Poly {
    std::vector<double> xy;
    float imp;
  };
  std::vector<Poly> polys;
  float *d_map = nullptr;
  int *d_max = nullptr;
  float *d_factor = nullptr;
  float h_factor = 0.f;
  bool factor_valid = true;
};

namespace {

// Ear-clipping triangulation of a simple polygon.  The reference partitions each polygon into
// y-monotone pieces with CGAL (polygon_utils.cpp:38-68) before the GPU test; triangles are
// y-monotone pieces with the same union, so the per-cell result is the same.
static double cross2(double ax, double ay, double bx, double by, double cx, double cy) {
  return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}
static bool in_tri(double px, double py, const double *a, const double *b, const double *c) {
  double d1 = cross2(a[0], a[1], b[0], b[1], px, py);
  double d2 = cross2(b[0], b[1], c[0], c[1], px, py);
  double d3 = cross2(c[0], c[1], a[0], a[1], px, py);
  bool neg = (d1 < 0) || (d2 < 0) || (d3 < 0);
  bool pos = (d1 > 0) || (d2 > 0) || (d3 > 0);
  return !(neg && pos);
}
static void triangulate(const std::vector<double> &xy, std::vector<std::array<double, 6>> &tris) {
  int n = (int)xy.size() / 2;
  if (n < 3) return;
  std::vector<int> idx(n);
  double area = 0;
  for (int i = 0; i < n; ++i) {
    int j = (i + 1) % n;
    area += xy[2 * i] * xy[2 * j + 1] - xy[2 * j] * xy[2 * i + 1];
  }
  for (int i = 0; i < n; ++i) idx[i] = area >= 0 ? i : n - 1 - i;  // make CCW (polygon_utils.cpp:48-50)
  int guard = 0;
  while ((int)idx.size() > 3 && guard < 10 * n * n) {
    int m = (int)idx.size();
    bool clipped = false;
    for (int i = 0; i < m; ++i) {
      int ia = idx[(i + m - 1) % m], ib = idx[i], ic = idx[(i + 1) % m];
      const double *a = &xy[2 * ia], *b = &xy[2 * ib], *c = &xy[2 * ic];
      if (cross2(a[0], a[1], b[0], b[1], c[0], c[1]) <= 0) continue;  // reflex or degenerate
      bool ear = true;
      for (int k = 0; k < m && ear; ++k) {
        int ik = idx[k];
        if (ik == ia || ik == ib || ik == ic) continue;
        if (in_tri(xy[2 * ik], xy[2 * ik + 1], a, b, c)) ear = false;
      }
      if (!ear) continue;
      tris.push_back({a[0], a[1], b[0], b[1], c[0], c[1]});
      idx.erase(idx.begin() + i);
      clipped = true;
      break;
    }
    ++guard;
    if (!clipped) break;  // numerically degenerate: emit what is left as a fan
  }
  for (size_t i = 1; i + 1 < idx.size(); ++i) {
    const double *a = &xy[2 * idx[0]], *b = &xy[2 * idx[i]], *c = &xy[2 * idx[i + 1]];
    tris.push_back({a[0], a[1], b[0], b[1], c[0], c[1]});
  }
}

}  // namespace

extern "C" int ccb_abi_version(void) { return CCB_ABI_VERSION; }
extern "C" const char *ccb_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t ccb_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int ccb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
extern "C" int ccb_cuda_available(void) { return ccb_device_count() > 0 ? 1 : 0; }

extern "C" void ccb_params_default(ccb_params *p) {
  p->pNumRobots = 32;
  p->pResolution = 1.0;
  p->pWorldMapSize = 1024;
  p->pRobotMapSize = 1024;
  p->pLocalMapSize = 256;
  p->pUpdateRobotMap = 1;
  p->pUpdateExplorationMap = 1;
  p->pUpdateSensorView = 1;
  p->pUpdateSystemMap = 1;
  p->pTruncationBND = 2;
  p->pNorm = 1;
  p->pUnknownImportance = 0.5;
  p->pRobotMapUseUnknownImportance = 0;
  p->pSensorSize = 64;
  p->pCommunicationRange = 128;
  p->pMaxRobotSpeed = 5;
  p->pRobotPosHistorySize = 20;
  p->pTimeStep = 1;
  p->pCheckOscillations = 1;
}

static int require_gpu() {
  if (!ccb_cuda_available()) return fail("no CUDA device: this library has no CPU fallback");
  return 0;
}

extern "C" ccb_idf *ccb_idf_create(const ccb_params *p, int device, void *stream) {
  if (require_gpu()) return nullptr;
  if (p->pWorldMapSize <= 0) {
    fail("pWorldMapSize must be positive");
    return nullptr;
  }
  auto idf = std::make_unique<ccb_idf>();
  idf->p = *p;
  idf->device = device;
  idf->stream = (cudaStream_t)stream;
  if (cudaSetDevice(device) != cudaSuccess) {
    fail("cudaSetDevice failed");
    return nullptr;
  }
  size_t nn = (size_t)p->pWorldMapSize * p->pWorldMapSize;
  if (cudaMalloc(&idf->d_map, nn * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&idf->d_max, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&idf->d_factor, sizeof(float)) != cudaSuccess) {
    fail("cudaMalloc failed (WorldIDF map)");
    cudaGetLastError();
    ccb_idf_destroy(idf.release());  // frees whatever was allocated before the failure
    return nullptr;
  }
  cudaMemsetAsync(idf->d_map, 0, nn * sizeof(float), idf->stream);
  return idf.release();
}

extern "C" void ccb_idf_destroy(ccb_idf *idf) {
  if (!idf) return;
  cudaSetDevice(idf->device);
  cudaFree(idf->d_map);
  cudaFree(idf->d_max);
  cudaFree(idf->d_factor);
  delete idf;
}

extern "C" int ccb_idf_add_bnd(ccb_idf *idf, double mx, double my, double sx, double sy, double rho, double sc) {
  (void)rho;
  // bivariate_normal_distribution.h:103-112: the general constructor tests the member rho_ (always
  // 0) instead of its argument, so every distribution the reference ever holds has rho = 0.
  idf->bnds.push_back({mx, my, sx, sy, sc, 0.0});
  return 0;
}

extern "C" int ccb_idf_add_polygon(ccb_idf *idf, const double *xy, int n, float imp) {
  if (n < 3) return fail("polygon needs at least 3 vertices");
  ccb_idf::Poly poly;
  poly.xy.assign(xy, xy + 2 * n);
  poly.imp = imp;
  idf->polys.push_back(std::move(poly));
  return 0;
}

// Test hook, host only: the y-monotone pieces (triangles) the library cuts a polygon into before the kernel's
// point-in-piece test (world_idf.cpp:125-153 does this with CGAL): lets a test hand the REFERENCE kernel the
// same pieces.  tris_out: cap x 6 doubles; returns the number of triangles or -1.
extern "C" int ccb_selftest_triangulate(const double *xy, int n, double *tris_out, int cap) {
  if (n < 3) return -1;
  std::vector<std::array<double, 6>> tris;
  triangulate(std::vector<double>(xy, xy + 2 * n), tris);
  if ((int)tris.size() > cap) return -1;
  for (size_t i = 0; i < tris.size(); ++i) memcpy(tris_out + 6 * i, tris[i].data(), 6 * sizeof(double));
  return (int)tris.size();
}

extern "C" int ccb_idf_num_features(const ccb_idf *idf) { return (int)(idf->bnds.size() + idf->polys.size()); }
extern "C" int ccb_idf_num_bnds(const ccb_idf *idf) { return (int)idf->bnds.size(); }
extern "C" int ccb_idf_num_polygons(const ccb_idf *idf) { return (int)idf->polys.size(); }
extern "C" int ccb_idf_get_bnds(const ccb_idf *idf, double *out6, int cap) {
  int n = std::min<int>(cap, (int)idf->bnds.size());
  for (int i = 0; i < n; ++i) memcpy(out6 + 6 * i, idf->bnds[i].data(), 6 * sizeof(double));
  return n;
}
extern "C" int ccb_idf_get_polygon(const ccb_idf *idf, int k, double *xy, int cap, float *imp) {
  if (k < 0 || k >= (int)idf->polys.size()) return -1;
  const auto &poly = idf->polys[k];
  int n = (int)poly.xy.size() / 2;
  if (imp) *imp = poly.imp;
  if (xy) memcpy(xy, poly.xy.data(), sizeof(double) * 2 * std::min(n, cap));
  return n;
}

static int idf_finish(ccb_idf *idf, int row0, int row1) {
  const int N = idf->p.pWorldMapSize;
  cudaStream_t st = idf->stream;
  k_idf_factor<<<1, 1, 0, st>>>(idf->d_max, (float)idf->p.pNorm, idf->d_factor);
  CCB_LAUNCH_CHECK();
  size_t nn = (size_t)(row1 - row0) * N;
  if (nn > 0) {
    k_idf_normalize<<<148 * 4, 256, 0, st>>>(idf->d_map + (size_t)row0 * N, ((size_t)row0 * N) % 4 == 0 ? nn / 4 : 0, nn, idf->d_factor);
    CCB_LAUNCH_CHECK();
  }
  idf->factor_valid = false;
  return 0;
}

// everything the integration kernel reads, packed on the host as ONE blob (one allocation, one copy)
struct IdfInputs {
  std::vector<unsigned char> blob;
  size_t o_b = 0, o_bs = 0, o_bi = 0, o_px = 0, o_py = 0, o_imp = 0, o_start = 0, o_sz = 0, o_bounds = 0, o_pks = 0, o_pki = 0;
  int F = 0, NP = 0, bk_nb = 0, bk_size = 0;
};

static void idf_pack_inputs(const ccb_idf *idf, IdfInputs &in) {
  const ccb_params &p = idf->p;
  const int N = p.pWorldMapSize;
  const int F = (int)idf->bnds.size();
  // feature records, fp32 as the reference packs them (world_idf.cpp:111-123)
  std::vector<BndF> hb(std::max(F, 1));
  for (int i = 0; i < F; ++i) {
    const auto &b = idf->bnds[i];
    hb[i] = {(float)b[0], (float)b[1], (float)b[2], (float)b[3], (float)b[4], (float)b[5]};
  }
  // polygon pieces (world_idf.cpp:125-153)
  std::vector<float> px, py, imp;
  std::vector<int> start, sz;
  std::vector<float4> bounds;
  for (const auto &poly : idf->polys) {
    std::vector<std::array<double, 6>> tris;
    triangulate(poly.xy, tris);
    for (const auto &t : tris) {
      float4 b = make_float4(3.4028235e38f, 1.17549435e-38f, 3.4028235e38f, 1.17549435e-38f);  // generate_world_map.h:49-54
      start.push_back((int)px.size());
      for (int v = 0; v < 3; ++v) {
        float x = (float)t[2 * v], y = (float)t[2 * v + 1];
        px.push_back(x), py.push_back(y);
        if (x < b.x) b.x = x;
        if (x > b.y) b.y = x;
        if (y < b.z) b.z = y;
        if (y > b.w) b.w = y;
      }
      sz.push_back(3);
      imp.push_back(poly.imp);
      bounds.push_back(b);
    }
  }
  const int NP = (int)sz.size();

  // coarse bucket grid over the features: bucket (bx, by) lists, in ascending order, the features whose
  // truncation box (a little larger than the one the kernel tests per tile) touches it
  const int bk_size = 256;
  const int bk_nb = (N + bk_size - 1) / bk_size;
  std::vector<int> bk_start((size_t)bk_nb * bk_nb + 1, 0), bk_items;
  {
    const double T = std::sqrt((double)p.pTruncationBND * p.pTruncationBND) * 1.001 + 1e-2;
    const double cell = p.pResolution * bk_size;
    std::vector<std::array<int, 4>> span(F);
    for (int i = 0; i < F; ++i) {
      const auto &b = idf->bnds[i];
      const double rx = T * std::fabs(b[2]) + 2 * p.pResolution, ry = T * std::fabs(b[3]) + 2 * p.pResolution;
      int x0 = (int)std::floor((b[0] - rx) / cell), x1 = (int)std::floor((b[0] + rx) / cell);
      int y0 = (int)std::floor((b[1] - ry) / cell), y1 = (int)std::floor((b[1] + ry) / cell);
      x0 = std::max(x0, 0), y0 = std::max(y0, 0), x1 = std::min(x1, bk_nb - 1), y1 = std::min(y1, bk_nb - 1);
      span[i] = {x0, x1, y0, y1};
      for (int bx = x0; bx <= x1; ++bx)
        for (int by = y0; by <= y1; ++by) bk_start[(size_t)bx * bk_nb + by + 1]++;
    }
    for (size_t k = 1; k < bk_start.size(); ++k) bk_start[k] += bk_start[k - 1];
    bk_items.resize(std::max<size_t>(bk_start.back(), 1));
    std::vector<int> fill(bk_start.begin(), bk_start.end() - 1);
    for (int i = 0; i < F; ++i)
      for (int bx = span[i][0]; bx <= span[i][1]; ++bx)
        for (int by = span[i][2]; by <= span[i][3]; ++by) bk_items[fill[(size_t)bx * bk_nb + by]++] = i;
  }
  // the same grid for the polygon pieces, by bounding box (closed, one cell of margin)
  std::vector<int> pk_start((size_t)bk_nb * bk_nb + 1, 0), pk_items;
  {
    const double cell = p.pResolution * bk_size, mg = p.pResolution;
    std::vector<std::array<int, 4>> span(NP);
    for (int i = 0; i < NP; ++i) {
      const float4 &b = bounds[i];
      int x0 = (int)std::floor((b.x - mg) / cell), x1 = (int)std::floor((b.y + mg) / cell);
      int y0 = (int)std::floor((b.z - mg) / cell), y1 = (int)std::floor((b.w + mg) / cell);
      x0 = std::max(x0, 0), y0 = std::max(y0, 0), x1 = std::min(x1, bk_nb - 1), y1 = std::min(y1, bk_nb - 1);
      span[i] = {x0, x1, y0, y1};
      for (int bx = x0; bx <= x1; ++bx)
        for (int by = y0; by <= y1; ++by) pk_start[(size_t)bx * bk_nb + by + 1]++;
    }
    for (size_t k = 1; k < pk_start.size(); ++k) pk_start[k] += pk_start[k - 1];
    pk_items.resize(std::max<size_t>(pk_start.back(), 1));
    std::vector<int> fill(pk_start.begin(), pk_start.end() - 1);
    for (int i = 0; i < NP; ++i)
      for (int bx = span[i][0]; bx <= span[i][1]; ++bx)
        for (int by = span[i][2]; by <= span[i][3]; ++by) pk_items[fill[(size_t)bx * bk_nb + by]++] = i;
  }
  // everything the kernel reads goes up in ONE blob (one allocation, one copy): the order counter for the
  // maximum, the feature records, the bucket grid and the polygon pieces, each section 16-byte aligned
  std::vector<unsigned char> &blob = in.blob;
  auto put = [&blob](const void *src, size_t bytes) -> size_t {
    const size_t off = (blob.size() + 15) & ~(size_t)15;
    blob.resize(off + bytes);
    if (bytes) memcpy(blob.data() + off, src, bytes);
    return off;
  };
  const size_t o_b = put(hb.data(), sizeof(BndF) * hb.size());
  const size_t o_bs = put(bk_start.data(), sizeof(int) * bk_start.size());
  const size_t o_bi = put(bk_items.data(), sizeof(int) * bk_items.size());
  const size_t o_px = put(px.data(), sizeof(float) * px.size()), o_py = put(py.data(), sizeof(float) * py.size());
  const size_t o_imp = put(imp.data(), sizeof(float) * imp.size());
  const size_t o_start = put(start.data(), sizeof(int) * start.size()), o_sz = put(sz.data(), sizeof(int) * sz.size());
  const size_t o_bounds = put(bounds.data(), sizeof(float4) * bounds.size());
  in.o_pks = put(pk_start.data(), sizeof(int) * pk_start.size()), in.o_pki = put(pk_items.data(), sizeof(int) * pk_items.size());
  in.o_b = o_b, in.o_bs = o_bs, in.o_bi = o_bi, in.o_px = o_px, in.o_py = o_py, in.o_imp = o_imp;
  in.o_start = o_start, in.o_sz = o_sz, in.o_bounds = o_bounds;
  in.F = F, in.NP = NP, in.bk_nb = bk_nb, in.bk_size = bk_size;
}

static int idf_launch_kernel(ccb_idf *idf, const IdfInputs &in, const unsigned char *bl, int mode, int row0, int row1) {
  const ccb_params &p = idf->p;
  const int N = p.pWorldMapSize;
  PolyDev pd{reinterpret_cast<const float *>(bl + in.o_px), reinterpret_cast<const float *>(bl + in.o_py),
             reinterpret_cast<const int *>(bl + in.o_start), reinterpret_cast<const int *>(bl + in.o_sz),
             reinterpret_cast<const float *>(bl + in.o_imp), reinterpret_cast<const float4 *>(bl + in.o_bounds), in.NP,
             reinterpret_cast<const int *>(bl + in.o_pks), reinterpret_cast<const int *>(bl + in.o_pki)};
  const BndF *d_b = reinterpret_cast<const BndF *>(bl + in.o_b);
  const int *d_bk_start = reinterpret_cast<const int *>(bl + in.o_bs), *d_bk_items = reinterpret_cast<const int *>(bl + in.o_bi);
  const float res = (float)p.pResolution;
  const float trun = (float)p.pTruncationBND;
  const float trun_sq = trun * trun;                       // generate_world_map.cu:188-189
  const float one_by_sqrt2 = (float)(1. / std::sqrt(2.));  // generate_world_map.cu:190-192
  if (row1 <= row0) return 0;
  const int rows_t = (row1 - (row0 & ~(IDF_TX - 1)) + IDF_TX - 1) / IDF_TX;
  // 32 x 128 tiles once there are several waves of them (148 SMs x ~8 CTAs), 32 x 32 below that
  const bool wide = (long)rows_t * ((N + 4 * IDF_TY - 1) / (4 * IDF_TY)) >= 4L * 148 * 8;
  dim3 grid((N + (wide ? 4 : 1) * IDF_TY - 1) / ((wide ? 4 : 1) * IDF_TY), rows_t);
#define CCB_IDF_LAUNCH(MODE, NTY)                                                                                              \
  k_idf_generate<MODE, NTY><<<grid, IDF_THREADS, 0, idf->stream>>>(d_b, in.F, d_bk_start, d_bk_items, in.bk_nb, in.bk_size, pd, N, row0, \
                                                                    row1, res, trun_sq, one_by_sqrt2, idf->d_map, idf->d_max)
  if (mode == 0) {
    if (wide) CCB_IDF_LAUNCH(0, 4);
    else CCB_IDF_LAUNCH(0, 1);
  } else {
    if (wide) CCB_IDF_LAUNCH(1, 4);
    else CCB_IDF_LAUNCH(1, 1);
  }
#undef CCB_IDF_LAUNCH
  CCB_LAUNCH_CHECK();
  return 0;
}

// rows [row0, row1) of the map; finish = compute the factor from the map maximum and normalise
static int idf_generate_rows(ccb_idf *idf, int mode, int row0, int row1, bool finish) {
  if (mode != 0 && mode != 1) return fail("ccb_idf_generate: mode must be 0 or 1");
  CCB_CUDA(cudaSetDevice(idf->device));
  const int N = idf->p.pWorldMapSize;
  if (row0 < 0 || row1 > N || row0 > row1) return fail("ccb_idf_generate: bad row band");
  cudaStream_t st = idf->stream;
  IdfInputs in;
  idf_pack_inputs(idf, in);
  unsigned char *d_blob = nullptr;
  CCB_CUDA(cudaMallocAsync(&d_blob, in.blob.size(), st));
  CCB_CUDA(cudaMemcpyAsync(d_blob, in.blob.data(), in.blob.size(), cudaMemcpyHostToDevice, st));
  const int init = (int)0x80000000;
  CCB_CUDA(cudaMemcpyAsync(idf->d_max, &init, sizeof(int), cudaMemcpyHostToDevice, st));
  if (idf_launch_kernel(idf, in, d_blob, mode, row0, row1)) return 1;
  if (finish && idf_finish(idf, row0, row1)) return 1;
  CCB_CUDA(cudaFreeAsync(d_blob, st));
  return 0;
}

// Profiling hook: the integration kernel alone, `iters` times over the whole map (inputs uploaded before the timed
// region; no maximum / normalise pass: the map is left un-normalised).  *ms_out = mean kernel time, CUDA events.
extern "C" int ccb_bench_idf_kernel(ccb_idf *idf, int mode, int iters, float *ms_out) {
  if (mode != 0 && mode != 1) return fail("ccb_bench_idf_kernel: mode must be 0 or 1");
  if (iters < 1) return fail("ccb_bench_idf_kernel: iters must be positive");
  CCB_CUDA(cudaSetDevice(idf->device));
  cudaStream_t st = idf->stream;
  IdfInputs in;
  idf_pack_inputs(idf, in);
  unsigned char *d_blob = nullptr;
  CCB_CUDA(cudaMalloc(&d_blob, in.blob.size()));
  CCB_CUDA(cudaMemcpyAsync(d_blob, in.blob.data(), in.blob.size(), cudaMemcpyHostToDevice, st));
  cudaEvent_t e0, e1;
  CCB_CUDA(cudaEventCreate(&e0));
  CCB_CUDA(cudaEventCreate(&e1));
  const int N = idf->p.pWorldMapSize;
  int rc = idf_launch_kernel(idf, in, d_blob, mode, 0, N);  // warm-up
  if (!rc) {
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters && !rc; ++i) rc = idf_launch_kernel(idf, in, d_blob, mode, 0, N);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(d_blob);
  return rc;
}

// Profiling hook: the FP32 FMA rate this GPU sustains (the denominator of the IDF kernel's roofline; there is no
// FP32 figure in MEASURED_PEAKS.json).  148 x 8 CTAs of 256 threads, 8 independent FFMA chains per thread.
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int n, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b), x1 = fmaf(x1, a, b), x2 = fmaf(x2, a, b), x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b), x5 = fmaf(x5, a, b), x6 = fmaf(x6, a, b), x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[0] = s;  // keeps the chains alive
}
extern "C" int ccb_bench_fp32_peak(int device, double *tflops_out) {
  if (require_gpu()) return 1;
  CCB_CUDA(cudaSetDevice(device));
  float *d_out = nullptr;
  CCB_CUDA(cudaMalloc(&d_out, sizeof(float)));
  cudaEvent_t e0, e1;
  CCB_CUDA(cudaEventCreate(&e0));
  CCB_CUDA(cudaEventCreate(&e1));
  const int n = 4096, ctas = 148 * 8;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CCB_CUDA(cudaEventRecord(e0, 0));
    k_fp32_peak<<<ctas, 256>>>(d_out, n, 0.999f, 0.001f);
    CCB_CUDA(cudaEventRecord(e1, 0));
    CCB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8 * 16 * (double)n * 256 * ctas;
    if (rep >= 1) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  g_launches.fetch_add(6);
  cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(d_out);
  *tflops_out = best;
  return 0;
}

extern "C" int ccb_idf_generate(ccb_idf *idf, int mode) {
  return idf_generate_rows(idf, mode, 0, idf->p.pWorldMapSize, true);
}
extern "C" int ccb_idf_generate_band(ccb_idf *idf, int mode, int row0, int row1) {
  return idf_generate_rows(idf, mode, row0, row1, false);
}
extern "C" int *ccb_idf_band_max_dev(ccb_idf *idf) { return idf->d_max; }
extern "C" int ccb_idf_normalize_band(ccb_idf *idf, int row0, int row1) {
  CCB_CUDA(cudaSetDevice(idf->device));
  if (row0 < 0 || row1 > idf->p.pWorldMapSize || row0 > row1) return fail("ccb_idf_normalize_band: bad row band");
  return idf_finish(idf, row0, row1);
}
extern "C" float *ccb_idf_map_dev_mut(ccb_idf *idf) { return idf->d_map; }

extern "C" int ccb_idf_get_map(ccb_idf *idf, float *out) {
  CCB_CUDA(cudaSetDevice(idf->device));
  size_t nn = (size_t)idf->p.pWorldMapSize * idf->p.pWorldMapSize;
  CCB_CUDA(cudaMemcpyAsync(out, idf->d_map, nn * sizeof(float), cudaMemcpyDeviceToHost, idf->stream));
  CCB_CUDA(cudaStreamSynchronize(idf->stream));
  return 0;
}

extern "C" int ccb_idf_set_map(ccb_idf *idf, const float *in) {
  CCB_CUDA(cudaSetDevice(idf->device));
  size_t nn = (size_t)idf->p.pWorldMapSize * idf->p.pWorldMapSize;
  CCB_CUDA(cudaMemcpyAsync(idf->d_map, in, nn * sizeof(float), cudaMemcpyHostToDevice, idf->stream));
  CCB_CUDA(cudaStreamSynchronize(idf->stream));
  return 0;
}

extern "C" const float *ccb_idf_map_dev(const ccb_idf *idf) { return idf->d_map; }

extern "C" int ccb_idf_copy_map(ccb_idf *dst, const ccb_idf *src) {
  if (dst->p.pWorldMapSize != src->p.pWorldMapSize) return fail("ccb_idf_copy_map: map sizes differ");
  CCB_CUDA(cudaSetDevice(dst->device));
  const size_t nn = (size_t)src->p.pWorldMapSize * src->p.pWorldMapSize;
  if (src->stream != dst->stream) CCB_CUDA(cudaStreamSynchronize(src->stream));
  CCB_CUDA(cudaMemcpyAsync(dst->d_map, src->d_map, nn * sizeof(float), cudaMemcpyDefault, dst->stream));
  CCB_CUDA(cudaMemcpyAsync(dst->d_factor, src->d_factor, sizeof(float), cudaMemcpyDefault, dst->stream));
  dst->h_factor = src->h_factor, dst->factor_valid = src->factor_valid;
  return 0;
}

extern "C" int ccb_idf_norm_factor(ccb_idf *idf, float *out) {
  if (!idf->factor_valid) {
    CCB_CUDA(cudaSetDevice(idf->device));
    CCB_CUDA(cudaMemcpyAsync(&idf->h_factor, idf->d_factor, sizeof(float), cudaMemcpyDeviceToHost, idf->stream));
    CCB_CUDA(cudaStreamSynchronize(idf->stream));
    idf->factor_valid = true;
  }
  *out = idf->h_factor;
  return 0;
}

// =================================================================================================
// CoverageSystem batch
// =================================================================================================
struct ResizeTab {
  int C = 0, kmax = 0;
  int *d_xmin = nullptr, *d_xsize = nullptr;
  float *d_w = nullptr;
  float *d_wfull8 = nullptr;  // [C][16] weights in the (8 o - 4) frame, only when L == 8 C == 256
  bool interior_uniform = false;  // rows 1 .. C - 2 of d_wfull8 are identical (k_obs_fused keeps them in registers)
};

struct ccb_env {
  ccb_params p;
  int E = 0, R = 0, device = 0;
  cudaStream_t stream = nullptr;
  EnvDev d{};
  double *d_sums = nullptr;     // [E][R][4] global Voronoi moments (world map)
  CellsOut cells{};             // env-level cells (ComputeVoronoiCells)
  double *d_objective = nullptr;  // [E]
  double *d_stage = nullptr;    // [E][R][2] staging for host-provided actions / positions
  double *d_feat = nullptr;     // [E][R][7]
  double *d_ratio = nullptr;    // [E][3]
  float *d_scratch = nullptr;   // max(N*N, L*L) floats for host getters
  double *h_stage[2] = {nullptr, nullptr};  // pinned host -> device staging (actions / positions), used in turn
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};  // recorded after the copy out of the slot
  int stage_next = 0;
  double *h_pin = nullptr;      // pinned host staging, E*R*CCB_MAX_CELL_VERTS*2 doubles
  size_t h_pin_bytes = 0;
  uint64_t pos_version = 1;     // bumped whenever positions or the world map change
  uint64_t cells_version = 0;   // pos_version the env-level cells were computed for
  bool cells_have_polys = false;
  bool initialised = false;
  std::map<int, ResizeTab> resize;
  std::map<const void *, size_t> func_smem;  // opt-in dynamic shared memory already granted, per kernel (on this device)
  int vor_rpw_override = 0;
  // second stream: the objective chain (global Voronoi -> finalise -> sum) reads only positions and the
  // world map, so it runs beside the observation kernels and the next step's window Voronoi.
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_state = nullptr;  // recorded on `stream` after every change of positions / maps
  cudaEvent_t ev_aux = nullptr;    // recorded on `aux` after the last chain launched there
  cudaEvent_t ev_cells = nullptr;  // recorded on `stream` after cells were (re)computed there
  cudaEvent_t ev_vg = nullptr;     // recorded on `stream` after every whole-map Voronoi launched there: the partition
                                   // tables / flag list / spans / items are ONE scratch set shared with `aux`
  bool no_aux = false;             // CCB_NO_AUX: keep everything on `stream` (debugging)
  bool no_graph = false;           // CCB_NO_GRAPH: never replay roll-outs from a CUDA graph
  // roll-out graphs are captured on and launched into `cap` (the caller's stream may be the legacy
  // default stream, which cannot be captured); every replay is bracketed by ev_fork / ev_join so that
  // it stays ordered with the work on `stream` before and after it
  cudaStream_t cap = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool cells_recomputed = false;   // set whenever ensure_cells launches (read by the roll-out graph capture)
  // DecentralizedCVT's window moments as a by-product of the observation kernel (obs.cuh, VOR variant):
  // valid for the positions of wsums_version; the next ComputeGoals picks them up instead of launching
  // k_voronoi_window over the same windows
  double *d_wsums = nullptr;       // [E][R][4]
  uint64_t wsums_version = 0;
  int n_decentralized = 0;         // live DecentralizedCVT controllers on this environment
  bool no_wfuse = false;           // CCB_NO_WFUSE: keep the two kernels apart (tests)
  bool aux_pending = false;        // `stream` has not yet waited for ev_aux
  uint32_t *d_vg_table = nullptr;  // [E][R][N] row intervals of the site-ordered global Voronoi (vglobal.cuh)
  int *d_vg_flag = nullptr;     // [1 + E*N]: count, then the flagged rows
  int *d_vg_span = nullptr;     // [E][R][2] row span of every site
  double *d_xfeat = nullptr;    // [E][R][2 K] exploration features (ccb_env_exploration_features)
  size_t feat_cap = 0;
  int *d_vg_items = nullptr;    // [E][R + 1] first (site, chunk) item of every site, then their total
  int vg_warps = 1;             // CCB_VG_WARPS: warps per CTA of k_vg_reduce (<= VG_RED_WARPS)
  int vg_griddiv = 4;           // CCB_VG_GRIDDIV: grid = largest possible item count / this
  int vg_mode = -1;             // CCB_VG_GENERIC: 1 forces the tile-ordered kernel, 0 the site-ordered one (tests)
  bool vw_generic = false;      // CCB_VW_GENERIC: force the generic window kernel (tests)
  bool band_mode = false;       // this env holds one row band of a world tiled over several GPUs
  int vg_pf = 1;                // test option: L2 prefetch distance of the fused partition walk (items)
  int vg_part_rpt = 1;          // CCB_VG_PART_RPT (opt-in, 1 / 2 / 4): rows per thread of k_vg_partition
  bool vg_prefix = true;        // block prefix sums of the world maps for the site-ordered world-map moments (vglobal.cuh)
  VgPfx *d_vg_pfx = nullptr;    // [E][N][N / 8 + 1] x 128 B, allocated on first use, rebuilt per environment when its world map is set
  std::vector<char> pfx_valid;  // [E]: the entries of environment e match its world map
  int *d_pfx_list = nullptr;    // [E] environments to rebuild
  double *d_vg_band = nullptr;  // [E][ceil(N / 128)][R][4] band moments of the fused partition + moments kernel
  bool vg_bulk = false;         // A/B: k_vg_reduce_bulk (cp.async.bulk + mbarrier ring) instead of k_vg_reduce
  bool vg_fused = true;         // world-map moments inside the row partition kernel (off: interval table + k_vg_lookup)
  bool cnn_generic = false;     // test option: the generic observation kernel even where the fused one applies
  int *d_obs_order = nullptr;   // [2 E R + 4]: CTA order of k_obs_fused, class words, class counters
  bool obs_sorted = true;       // test option: launch k_obs_fused's CTAs grouped by window misalignment
  bool obs_v1 = false;          // test option: the row-ordered fused observation kernel of round 1 (A/B timing)
};

// kernels that need more than 48 KB of dynamic shared memory: raise the limit once per (device, kernel)
template <class K>
static int ensure_smem(ccb_env *env, K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  size_t &have = env->func_smem[reinterpret_cast<const void *>(kernel)];
  if (bytes > have) {
    CCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return 0;
}

// every change of the robot state or the maps goes through here: new version + fork point for `aux`
static int state_changed(ccb_env *env) {
  env->pos_version++;
  CCB_CUDA(cudaEventRecord(env->ev_state, env->stream));
  return 0;
}
// make `stream` wait for whatever was launched on `aux` (no-op when nothing is pending)
static int join_aux(ccb_env *env) {
  if (env->aux_pending) {
    CCB_CUDA(cudaStreamWaitEvent(env->stream, env->ev_aux, 0));
    env->aux_pending = false;
  }
  return 0;
}
struct AuxScope {  // launches issued through env->stream go to `aux` while this is alive
  ccb_env *env;
  cudaStream_t saved;
  explicit AuxScope(ccb_env *e) : env(e), saved(e->stream) { e->stream = e->aux; }
  ~AuxScope() { env->stream = saved; }
};

// rows per warp region: as large as keeps >= ~3 CTAs per SM in flight over the whole launch
static int pick_rpw(const ccb_env *env, int region_rows, int region_cols, int regions) {
  if (env->vor_rpw_override > 0) return std::max(2, std::min(env->vor_rpw_override, VOR_MAXRPW) & ~1);
  const long tiles_y = (region_cols + VOR_WX * VOR_TW - 1) / (VOR_WX * VOR_TW);
  for (int rpw = VOR_MAXRPW; rpw > 4; rpw >>= 1) {
    long tiles_x = (region_rows + VOR_WY * rpw - 1) / (VOR_WY * rpw);
    if ((long)regions * tiles_x * tiles_y >= 3L * 148) return rpw;
  }
  return 4;
}

template <int SRC, bool SINGLE, bool LABELS>
static int run_voronoi(ccb_env *env, const VorArgs &a, dim3 grid) {
  const size_t smem = vor_smem_bytes(env->d.R);
  if (ensure_smem(env, k_voronoi<SRC, SINGLE, LABELS>, smem)) return 1;
  k_voronoi<SRC, SINGLE, LABELS><<<grid, VOR_THREADS, smem, env->stream>>>(a);
  CCB_LAUNCH_CHECK();
  return 0;
}

static int ensure_prefix(ccb_env *env);

// Which whole-map Voronoi kernel (measured: profiles/r01_sweep_c5.md).  The site-ordered path pays one sequential
// sweep per map row (~30 us whatever the batch) and then walks every site's intervals; the tile-ordered kernel is
// faster where most 64 x 64 regions have a single owner (cells of more than ~128 K map cells: 4.7 TB/s at N = 16384,
// R = 64) and slower where they do not (1.5 TB/s at N = 1024, R = 32).
static bool use_site_ordered(const ccb_env *env, int rows) {
  const EnvDev &d = env->d;
  const bool by_shape = (long)d.E * rows >= 32L * 1024 && (long)d.N * d.N / d.R <= 128L * 1024;
  const bool wanted = env->vg_mode == 0 || (env->vg_mode < 0 && by_shape);
  return d.R <= VG_MAXR && (d.N & 3) == 0 && d.N < 65536 && wanted;
}

// Voronoi label + reduce over the full map of every environment; sums must be a [E][R][4] buffer.
// row0/row1 restrict the pass to a row band (row1 < 0: the whole map).
// stages: bit 0 row partition, bit 1 per-site reduce, bit 2 flagged-row fix-up (site-ordered path; the
// profiling hook times the reduce alone on the tables of a previous full call)
static int launch_voronoi_global(ccb_env *env, int src, double *sums, int *labels = nullptr, int row0 = 0, int row1 = -1,
                                 int stages = 7) {
  const EnvDev &d = env->d;
  if (row1 < 0) row1 = d.N;
  if (env->stream != env->aux && join_aux(env)) return 1;  // the partition tables are shared with `aux`
  CCB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * d.E * d.R, env->stream));
  if (row1 <= row0) return 0;
  if (use_site_ordered(env, row1 - row0)) {
    // site-ordered path: row partition -> per-site interval streaming -> exact rule on flagged rows
    if (!env->d_vg_table) {
      CCB_CUDA(cudaMalloc(&env->d_vg_table, sizeof(uint32_t) * (size_t)d.E * d.R * d.N));
      CCB_CUDA(cudaMalloc(&env->d_vg_flag, sizeof(int) * ((size_t)d.E * d.N + 1)));
      CCB_CUDA(cudaMalloc(&env->d_vg_span, sizeof(int) * 2 * (size_t)d.E * d.R));
      CCB_CUDA(cudaMalloc(&env->d_vg_items, sizeof(int) * (size_t)d.E * (d.R + 1)));
    }
    bool lookup = false;
    if (src == SRC_WORLD && !labels && env->vg_prefix) {
      if (ensure_prefix(env)) return 1;
      lookup = env->d_vg_pfx != nullptr;
    }
    const int rows = row1 - row0;
    const int nbands = (rows + VG_PART_THREADS - 1) / VG_PART_THREADS;
    const bool fused = lookup && env->vg_fused;
    if (fused && !env->d_vg_band)
      CCB_CUDA(cudaMalloc(&env->d_vg_band, sizeof(double) * 4 * d.E * d.R * ((d.N + VG_PART_THREADS - 1) / VG_PART_THREADS)));
    if ((stages & 1) && !fused) {
      k_vg_init_span<<<(d.E * d.R + 255) / 256, 256, 0, env->stream>>>(env->d_vg_span, d.E * d.R);
      CCB_LAUNCH_CHECK();
    }
    if (stages & 1) CCB_CUDA(cudaMemsetAsync(env->d_vg_flag, 0, sizeof(int), env->stream));
    VgArgs g;
    g.d = d, g.row0 = row0, g.row1 = row1, g.table = env->d_vg_table, g.nflag = env->d_vg_flag, g.flag_list = env->d_vg_flag + 1, g.span = env->d_vg_span, g.sums = sums, g.labels = labels;
    g.items = env->d_vg_items;
    g.pfx = env->d_vg_pfx, g.band = env->d_vg_band, g.pf = env->vg_pf;
    const size_t psmem = sizeof(double) * d.R * d.R;
    if (fused) {
      // world-map moments inside the row partition: one prefix entry + one map sector per boundary, band moments
      // stored, bands added in order (no table, no spans, no atomics)
      if (stages & 1) {
        // (static + dynamic shared memory pass 48 KB at R = 64: the opt-in covers the dynamic part)
        size_t &have = env->func_smem[reinterpret_cast<const void *>(k_vg_partition<1, true>)];
        if (have < psmem + 1) {
          CCB_CUDA(cudaFuncSetAttribute(k_vg_partition<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(psmem, 40 * 1024)));
          have = psmem + 1;
        }
        k_vg_partition<1, true><<<dim3(nbands, d.E), VG_PART_THREADS, psmem, env->stream>>>(g);
        CCB_LAUNCH_CHECK();
      }
      if (stages & 2) {
        k_vg_bandsum<<<(d.E * d.R * 4 + 255) / 256, 256, 0, env->stream>>>(env->d_vg_band, sums, d.E, d.R, nbands);
        CCB_LAUNCH_CHECK();
      }
    } else {
    if (stages & 1) {
      const int rpt = env->vg_part_rpt, band = rpt * VG_PART_THREADS;
      const dim3 pgrid((rows + band - 1) / band, d.E);
      if (rpt == 2) k_vg_partition<2><<<pgrid, VG_PART_THREADS, psmem, env->stream>>>(g);
      else if (rpt == 4) k_vg_partition<4><<<pgrid, VG_PART_THREADS, psmem, env->stream>>>(g);
      else k_vg_partition<1><<<pgrid, VG_PART_THREADS, psmem, env->stream>>>(g);
      CCB_LAUNCH_CHECK();
      if (!lookup) {
        k_vg_items<<<(d.E + 3) / 4, 128, 0, env->stream>>>(env->d_vg_span, env->d_vg_items, d.E, d.R, row0);
        CCB_LAUNCH_CHECK();
      }
    }
    // A site of a typical partition spans a fraction of the rows: a quarter of the largest possible item
    // count (every site on every chunk) covers it with about one item per warp; the warps stride beyond it.
    const int chunks = (rows + VG_CHUNK - 1) / VG_CHUNK;
    const int max_ctas = (chunks * d.R + env->vg_warps - 1) / env->vg_warps;
    dim3 rgrid(std::max(1, (max_ctas + env->vg_griddiv - 1) / env->vg_griddiv), d.E);
    if (stages & 2) {
      const int nt = env->vg_warps * 32;
      if (lookup) {
        // site-ordered look-ups on the interval table (the parity reference of the fused kernel)
        k_vg_lookup<<<dim3((d.R + VG_LOOK_WARPS - 1) / VG_LOOK_WARPS, d.E), VG_LOOK_WARPS * 32, 0, env->stream>>>(g, env->d_vg_pfx);
      } else if (src == SRC_WORLD && !labels && env->vg_bulk) {
        // A/B variant: bulk-asynchronous copies into a shared-memory ring (ccb_test_set_option "vg_bulk")
        if (ensure_smem(env, k_vg_reduce_bulk, VGB_SMEM)) return 1;
        const int items_max = chunks * d.R;
        dim3 bgrid(std::max(1, std::min(148 * 1 * 8, (items_max + VGB_WARPS - 1) / VGB_WARPS)), d.E);
        k_vg_reduce_bulk<<<bgrid, VGB_WARPS * 32, VGB_SMEM, env->stream>>>(g);
      } else if (src == SRC_WORLD) {
        if (labels) k_vg_reduce<SRC_WORLD, true><<<rgrid, nt, 0, env->stream>>>(g);
        else k_vg_reduce<SRC_WORLD, false, VG_U, VG_RED_MINB, 4><<<rgrid, nt, 0, env->stream>>>(g);  // 8 rows x 16 columns per pass
      } else {
        if (labels) k_vg_reduce<SRC_EXPLORED, true><<<rgrid, nt, 0, env->stream>>>(g);
        else k_vg_reduce<SRC_EXPLORED, false, VG_U, VG_RED_MINB, 4><<<rgrid, nt, 0, env->stream>>>(g);
      }
      CCB_LAUNCH_CHECK();
    }
    }
    const dim3 fgrid(148 * 2);
    if (stages & 4) {
      if (src == SRC_WORLD) {
        if (labels) k_vg_fixup<SRC_WORLD, true><<<fgrid, 256, 0, env->stream>>>(g);
        else k_vg_fixup<SRC_WORLD, false><<<fgrid, 256, 0, env->stream>>>(g);
      } else {
        if (labels) k_vg_fixup<SRC_EXPLORED, true><<<fgrid, 256, 0, env->stream>>>(g);
        else k_vg_fixup<SRC_EXPLORED, false><<<fgrid, 256, 0, env->stream>>>(g);
      }
      CCB_LAUNCH_CHECK();
    }
    if (env->stream != env->aux) CCB_CUDA(cudaEventRecord(env->ev_vg, env->stream));  // `aux` waits before reusing the scratch
    return 0;
  }
  VorArgs a;
  a.d = d, a.sums = sums, a.labels = labels;
  a.row0 = row0, a.row1 = row1;
  a.rpw = pick_rpw(env, row1 - row0, d.N, d.E);
  a.tiles_y = (d.N + VOR_WX * VOR_TW - 1) / (VOR_WX * VOR_TW);
  a.tiles_x = (row1 - row0 + VOR_WY * a.rpw - 1) / (VOR_WY * a.rpw);
  dim3 grid(a.tiles_x * a.tiles_y, 1, d.E);
  if (src == SRC_WORLD) return labels ? run_voronoi<SRC_WORLD, false, true>(env, a, grid) : run_voronoi<SRC_WORLD, false, false>(env, a, grid);
  return labels ? run_voronoi<SRC_EXPLORED, false, true>(env, a, grid) : run_voronoi<SRC_EXPLORED, false, false>(env, a, grid);
}

// per-robot L-window variant (DecentralizedCVT): only the robot's own cell is reduced
static int launch_voronoi_single(ccb_env *env, double *sums) {
  const EnvDev &d = env->d;
  if (env->wsums_version == env->pos_version) {  // the observation kernel already reduced these windows
    CCB_CUDA(cudaMemcpyAsync(sums, env->d_wsums, sizeof(double) * 4 * d.E * d.R, cudaMemcpyDeviceToDevice, env->stream));
    return 0;
  }
  if (d.L <= VW_MAXL && (d.N & 3) == 0 && !env->vw_generic) {
    const size_t smem = vw_smem_bytes(d.R);
    if (ensure_smem(env, k_voronoi_window, smem)) return 1;
    k_voronoi_window<<<dim3(d.R, d.E), VW_THREADS, smem, env->stream>>>(d, sums, 0);
    CCB_LAUNCH_CHECK();
    return 0;
  }
  CCB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * d.E * d.R, env->stream));
  VorArgs a;
  a.d = d, a.sums = sums, a.labels = nullptr;
  a.rpw = env->vor_rpw_override > 0 ? std::max(2, std::min(env->vor_rpw_override, VOR_MAXRPW) & ~1) : 32;
  a.tiles_x = a.tiles_y = 1;  // one CTA per robot window
  a.row0 = 0, a.row1 = d.N;
  return run_voronoi<SRC_ROBOT, true, false>(env, a, dim3(1, d.R, d.E));
}

static int launch_finalize(ccb_env *env, const double *sums, CellsOut out, int want_polys) {
  const EnvDev &d = env->d;
  const int warps = d.E * d.R;
  k_finalize_cells<<<(warps + FIN_WARPS - 1) / FIN_WARPS, FIN_WARPS * 32, 0, env->stream>>>(d, sums, out, want_polys);
  CCB_LAUNCH_CHECK();
  return 0;
}

static int ensure_cells(ccb_env *env, bool want_polys) {
  if (!env->initialised) return fail("environment has no robot positions yet (call ccb_env_reset)");
  if (env->stream != env->aux && join_aux(env)) return 1;  // the cells may still be in the making on `aux`
  if (env->cells_version != env->pos_version) {
    if (env->band_mode)
      return fail("row-band mode: all-reduce the partials and call ccb_env_finalize_from_partials_dev first");
    if (launch_voronoi_global(env, SRC_WORLD, env->d_sums)) return 1;
    if (launch_finalize(env, env->d_sums, env->cells, want_polys ? 1 : 0)) return 1;
    env->cells_recomputed = true;
    env->cells_version = env->pos_version;
    env->cells_have_polys = want_polys;
  } else if (want_polys && !env->cells_have_polys) {
    if (launch_finalize(env, env->d_sums, env->cells, 1)) return 1;
    env->cells_have_polys = true;
  }
  if (env->stream != env->aux) CCB_CUDA(cudaEventRecord(env->ev_cells, env->stream));
  return 0;
}

static int alloc_cells(CellsOut &c, int E, int R, bool with_polys) {
  size_t er = (size_t)E * R;
  CCB_CUDA(cudaMalloc(&c.mass, er * sizeof(double)));
  CCB_CUDA(cudaMalloc(&c.centroid, er * 2 * sizeof(double)));
  CCB_CUDA(cudaMalloc(&c.obj, er * sizeof(double)));
  if (with_polys) {
    CCB_CUDA(cudaMalloc(&c.nverts, er * sizeof(int)));
    CCB_CUDA(cudaMalloc(&c.verts, er * CCB_MAX_CELL_VERTS * 2 * sizeof(double)));
  }
  return 0;
}
static void free_cells(CellsOut &c) {
  cudaFree(c.mass), cudaFree(c.centroid), cudaFree(c.obj), cudaFree(c.nverts), cudaFree(c.verts);
  c = CellsOut{};
}

static int env_alloc(ccb_env *env) {
  EnvDev &d = env->d;
  const ccb_params &p = env->p;
  const size_t nn = (size_t)d.N * d.N, er = (size_t)d.E * d.R;
  // VOR_PAD_ROWS rows of padding: the Voronoi load ring prefetches past the last region of the last map
  CCB_CUDA(cudaMalloc(&d.world, sizeof(float) * (nn * d.E + (size_t)VOR_PAD_ROWS * d.N)));
  const size_t pw = plane_words(d);
  CCB_CUDA(cudaMalloc(&d.seenb, sizeof(uint32_t) * pw * er));
  if (p.pUpdateExplorationMap) CCB_CUDA(cudaMalloc(&d.xseenb, sizeof(uint32_t) * pw * er));
  CCB_CUDA(cudaMalloc(&d.sysb, sizeof(uint32_t) * pw * 2 * d.E));
  CCB_CUDA(cudaMalloc(&d.seen_box, sizeof(int) * 4 * er));
  CCB_CUDA(cudaMemsetAsync(d.seen_box, 0, sizeof(int) * 4 * er, env->stream));
  CCB_CUDA(cudaMalloc(&d.pos_g, sizeof(double) * 2 * er));
  CCB_CUDA(cudaMalloc(&d.pos_s, sizeof(double) * 2 * er));
  CCB_CUDA(cudaMalloc(&d.pos_l, sizeof(double) * 2 * er));
  CCB_CUDA(cudaMalloc(&d.hist, sizeof(double) * 2 * er * d.H));
  CCB_CUDA(cudaMalloc(&d.hist_start, sizeof(int) * er));
  CCB_CUDA(cudaMalloc(&d.hist_len, sizeof(int) * er));
  CCB_CUDA(cudaMalloc(&d.status, sizeof(int) * d.E));
  CCB_CUDA(cudaMalloc(&env->d_sums, sizeof(double) * 4 * er));
  CCB_CUDA(cudaMalloc(&env->d_wsums, sizeof(double) * 4 * er));
  if (alloc_cells(env->cells, d.E, d.R, true)) return 1;
  CCB_CUDA(cudaMalloc(&env->d_objective, sizeof(double) * d.E));
  CCB_CUDA(cudaMalloc(&env->d_stage, sizeof(double) * 2 * er));
  CCB_CUDA(cudaMalloc(&env->d_feat, sizeof(double) * 7 * er));
  CCB_CUDA(cudaMalloc(&env->d_ratio, sizeof(double) * 3 * d.E));
  CCB_CUDA(cudaMalloc(&env->d_scratch, sizeof(float) * std::max(nn, (size_t)d.L * d.L)));
  env->h_pin_bytes = sizeof(double) * std::max<size_t>(er * CCB_MAX_CELL_VERTS * 2, 64);
  CCB_CUDA(cudaMallocHost(&env->h_pin, env->h_pin_bytes));
  for (int k = 0; k < 2; ++k) {
    CCB_CUDA(cudaMallocHost(&env->h_stage[k], sizeof(double) * 2 * er));
    CCB_CUDA(cudaEventCreateWithFlags(&env->ev_stage[k], cudaEventDisableTiming));
  }
  CCB_CUDA(cudaMemsetAsync(d.world, 0, sizeof(float) * (nn * d.E + (size_t)VOR_PAD_ROWS * d.N), env->stream));
  env->pfx_valid.assign(d.E, 0);
  CCB_CUDA(cudaMemsetAsync(d.seenb, 0, sizeof(uint32_t) * pw * er, env->stream));
  if (d.xseenb) CCB_CUDA(cudaMemsetAsync(d.xseenb, 0, sizeof(uint32_t) * pw * er, env->stream));
  CCB_CUDA(cudaMemsetAsync(d.sysb, 0, sizeof(uint32_t) * pw * 2 * d.E, env->stream));
  CCB_CUDA(cudaMemsetAsync(d.status, 0, sizeof(int) * d.E, env->stream));
  CCB_CUDA(cudaStreamCreateWithFlags(&env->aux, cudaStreamNonBlocking));
  CCB_CUDA(cudaEventCreateWithFlags(&env->ev_state, cudaEventDisableTiming));
  CCB_CUDA(cudaEventCreateWithFlags(&env->ev_aux, cudaEventDisableTiming));
  CCB_CUDA(cudaEventCreateWithFlags(&env->ev_cells, cudaEventDisableTiming));
  CCB_CUDA(cudaEventCreateWithFlags(&env->ev_vg, cudaEventDisableTiming));
  CCB_CUDA(cudaStreamCreateWithFlags(&env->cap, cudaStreamNonBlocking));
  CCB_CUDA(cudaEventCreateWithFlags(&env->ev_fork, cudaEventDisableTiming));
  CCB_CUDA(cudaEventCreateWithFlags(&env->ev_join, cudaEventDisableTiming));
  CCB_CUDA(cudaEventRecord(env->ev_state, env->stream));
  CCB_CUDA(cudaEventRecord(env->ev_cells, env->stream));
  CCB_CUDA(cudaEventRecord(env->ev_vg, env->stream));
  return 0;
}

extern "C" ccb_env *ccb_env_create(const ccb_params *p, int num_envs, int num_robots, int device, void *stream) {
  if (require_gpu()) return nullptr;
  if (num_envs < 1 || num_robots < 1) {
    fail("ccb_env_create: need at least one environment and one robot");
    return nullptr;
  }
  if (p->pRobotMapSize != p->pWorldMapSize) {
    fail("pRobotMapSize must equal pWorldMapSize (robot maps share the world grid, robot_model.h:69)");
    return nullptr;
  }
  if (p->pSensorSize <= 0 || p->pLocalMapSize <= 0 || p->pWorldMapSize <= 0) {
    fail("map sizes must be positive");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    fail("cudaSetDevice failed");
    return nullptr;
  }
  auto env = std::make_unique<ccb_env>();
  env->p = *p;
  env->E = num_envs, env->R = num_robots, env->device = device;
  env->stream = (cudaStream_t)stream;
  EnvDev &d = env->d;
  d.E = num_envs, d.R = num_robots, d.N = p->pWorldMapSize, d.S = p->pSensorSize, d.L = p->pLocalMapSize;
  d.H = std::max(1, p->pRobotPosHistorySize);
  d.NW = (d.N + 31) / 32;
  d.res = p->pResolution;
  d.fill = p->pRobotMapUseUnknownImportance ? (float)(p->pUnknownImportance * p->pNorm) : 0.f;
  d.use_sensor = p->pUpdateSensorView, d.upd_robot = p->pUpdateRobotMap;
  d.upd_expl = p->pUpdateExplorationMap, d.upd_sys = p->pUpdateSystemMap;
  d.check_osc = p->pCheckOscillations;
  d.obs_prefetch = 1;
  d.comm_range = p->pCommunicationRange, d.max_speed = p->pMaxRobotSpeed, d.dt = p->pTimeStep;
  if (env_alloc(env.get())) {
    ccb_env_destroy(env.release());
    return nullptr;
  }
  return env.release();
}

extern "C" void ccb_env_destroy(ccb_env *env) {
  if (!env) return;
  cudaSetDevice(env->device);
  cudaStreamSynchronize(env->stream);
  if (env->aux) cudaStreamSynchronize(env->aux), cudaStreamDestroy(env->aux);
  if (env->ev_state) cudaEventDestroy(env->ev_state);
  if (env->ev_aux) cudaEventDestroy(env->ev_aux);
  if (env->ev_cells) cudaEventDestroy(env->ev_cells);
  if (env->ev_vg) cudaEventDestroy(env->ev_vg);
  if (env->cap) cudaStreamSynchronize(env->cap), cudaStreamDestroy(env->cap);
  if (env->ev_fork) cudaEventDestroy(env->ev_fork);
  if (env->ev_join) cudaEventDestroy(env->ev_join);
  EnvDev &d = env->d;
  cudaFree(d.world), cudaFree(d.seenb), cudaFree(d.xseenb), cudaFree(d.sysb), cudaFree(d.seen_box);
  cudaFree(d.pos_g), cudaFree(d.pos_s), cudaFree(d.pos_l);
  cudaFree(d.hist), cudaFree(d.hist_start), cudaFree(d.hist_len), cudaFree(d.status);
  cudaFree(env->d_sums), cudaFree(env->d_wsums);
  free_cells(env->cells);
  cudaFree(env->d_objective), cudaFree(env->d_stage), cudaFree(env->d_feat), cudaFree(env->d_ratio);
  cudaFree(env->d_scratch);
  cudaFree(env->d_vg_table), cudaFree(env->d_vg_flag), cudaFree(env->d_vg_span), cudaFree(env->d_vg_items), cudaFree(env->d_xfeat);
  cudaFree(env->d_vg_pfx), cudaFree(env->d_pfx_list), cudaFree(env->d_vg_band), cudaFree(env->d_obs_order);
  cudaFreeHost(env->h_pin);
  for (int k = 0; k < 2; ++k) {
    cudaFreeHost(env->h_stage[k]);
    if (env->ev_stage[k]) cudaEventDestroy(env->ev_stage[k]);
  }
  for (auto &kv : env->resize)
    cudaFree(kv.second.d_xmin), cudaFree(kv.second.d_xsize), cudaFree(kv.second.d_w), cudaFree(kv.second.d_wfull8);
  delete env;
}

// test / A-B hook (include/ccb200_test.h): the library never reads the process environment
extern "C" int ccb_test_set_option(ccb_env *env, const char *name, int value) {
  const std::string n(name ? name : "");
  if (n == "vor_rpw") env->vor_rpw_override = value;
  else if (n == "vw_generic") env->vw_generic = value != 0;
  else if (n == "vg_generic") env->vg_mode = value < 0 ? -1 : (value != 0 ? 1 : 0);
  else if (n == "vg_warps") env->vg_warps = std::min(std::max(value, 1), VG_RED_WARPS);
  else if (n == "vg_griddiv") env->vg_griddiv = std::max(value, 1);
  else if (n == "vg_prefix") env->vg_prefix = value != 0;
  else if (n == "vg_fused") env->vg_fused = value != 0;
  else if (n == "vg_bulk") env->vg_bulk = value != 0;
  else if (n == "vg_pf") env->vg_pf = std::max(0, std::min(value, 8));
  else if (n == "vg_part_rpt") env->vg_part_rpt = (value == 2 || value == 4) ? value : 1;
  else if (n == "no_aux") env->no_aux = value != 0;
  else if (n == "no_graph") env->no_graph = value != 0;
  else if (n == "no_wfuse") env->no_wfuse = value != 0;
  else if (n == "cnn_generic") env->cnn_generic = value != 0;
  else if (n == "obs_v1") env->obs_v1 = value != 0;
  else if (n == "obs_sorted") env->obs_sorted = value != 0;
  else if (n == "obs_prefetch") env->d.obs_prefetch = value != 0;
  else return fail("ccb_test_set_option: unknown option", name);
  return 0;
}

extern "C" int ccb_env_num_envs(const ccb_env *env) { return env->E; }
extern "C" int ccb_env_num_robots(const ccb_env *env) { return env->R; }
extern "C" const float *ccb_env_world_dev(const ccb_env *env) { return env->d.world; }
extern "C" const double *ccb_env_positions_dev(const ccb_env *env) { return env->d.pos_g; }

// the world map of environment e was just written on env->stream: its block prefix sums are stale
static int world_changed(ccb_env *env, int e) {
  env->pfx_valid[e] = 0;
  return 0;
}

// Block prefix sums of every environment's world map, (re)built where stale on env->stream.  Returns with
// env->d_vg_pfx == nullptr when the path is off or its memory cannot be had (the streaming reduce runs instead).
static int ensure_prefix(ccb_env *env) {
  const EnvDev &d = env->d;
  if (!env->vg_prefix || d.R > VG_MAXR || (d.N % VG_PB) != 0 || (d.N & 3) != 0 || d.N >= 65536) return 0;
  const size_t per_env = (size_t)d.N * vg_pfx_row_entries(d.N);
  if (!env->d_vg_pfx) {
    if (cudaMalloc(&env->d_vg_pfx, sizeof(VgPfx) * per_env * d.E) != cudaSuccess ||
        cudaMalloc(&env->d_pfx_list, sizeof(int) * d.E) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(env->d_vg_pfx);
      env->d_vg_pfx = nullptr;
      env->vg_prefix = false;  // 16 bytes per map cell were not available: keep streaming the maps
      return 0;
    }
    env->pfx_valid.assign(d.E, 0);
  }
  std::vector<int> todo;
  for (int e = 0; e < d.E; ++e)
    if (!env->pfx_valid[e]) todo.push_back(e);
  if (todo.empty()) return 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(env->stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return fail("block prefix sums are stale inside a graph capture (internal error)");
  const int *list = nullptr;
  if ((int)todo.size() != d.E) {
    // (pageable source: the copy has been staged when the call returns, so `todo` may go out of scope)
    CCB_CUDA(cudaMemcpyAsync(env->d_pfx_list, todo.data(), sizeof(int) * todo.size(), cudaMemcpyHostToDevice, env->stream));
    list = env->d_pfx_list;
  }
  const long threads = (long)todo.size() * d.N;  // one per map row: a sequential double-double scan
  k_vg_prefix<<<(unsigned)((threads + 127) / 128), 128, 0, env->stream>>>(d.world, env->d_vg_pfx, d.N, list, (int)todo.size());
  CCB_LAUNCH_CHECK();
  for (int e : todo) env->pfx_valid[e] = 1;
  return 0;
}

static int check_env_index(const ccb_env *env, int e) {
  if (e < 0 || e >= env->E) return fail("environment index out of range");
  return 0;
}
static int check_robot(const ccb_env *env, int r) {
  if (r < 0 || r >= env->R) return fail("Robot index more than the number of robots");  // coverage_system.h:365-369
  return 0;
}

extern "C" int ccb_env_set_world_dev(ccb_env *env, int e, const float *map_dev) {
  if (check_env_index(env, e)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  size_t nn = cell_count(env->d);
  CCB_CUDA(cudaMemcpyAsync(env->d.world + nn * e, map_dev, nn * sizeof(float), cudaMemcpyDeviceToDevice, env->stream));
  if (world_changed(env, e)) return 1;
  if (state_changed(env)) return 1;
  return 0;
}
extern "C" int ccb_env_set_world_idf(ccb_env *env, int e, const ccb_idf *idf) {
  if (idf->p.pWorldMapSize != env->p.pWorldMapSize) return fail("World map size does not match the specified size");
  if (idf->stream != env->stream) cudaStreamSynchronize(idf->stream);
  return ccb_env_set_world_dev(env, e, idf->d_map);
}
extern "C" int ccb_env_set_world_host(ccb_env *env, int e, const float *map) {
  if (check_env_index(env, e)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  size_t nn = cell_count(env->d);
  CCB_CUDA(cudaMemcpyAsync(env->d.world + nn * e, map, nn * sizeof(float), cudaMemcpyHostToDevice, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  if (world_changed(env, e)) return 1;
  if (state_changed(env)) return 1;
  return 0;
}

static int launch_mark(ccb_env *env, int e, int robot, int mark_robot, int mark_sys, const double *move_actions = nullptr) {
  const EnvDev &d = env->d;
  const long warps = (long)(robot >= 0 ? 1 : d.R) * (e >= 0 ? 1 : d.E);  // one warp per (robot, environment)
  k_mark_windows<<<(unsigned)((warps + MARK_WARPS - 1) / MARK_WARPS), MARK_WARPS * 32, 0, env->stream>>>(d, e, robot, mark_robot, mark_sys, move_actions);
  CCB_LAUNCH_CHECK();
  return 0;
}

// A caller's buffer that is page-locked itself (cudaHostAlloc / cudaHostRegister, a pinned torch tensor) is copied
// from / to directly: no pass through the library's own pinned slots.  Asked only for transfers large enough for the
// extra host memcpy to matter.
static bool user_buffer_is_pinned(const void *p, size_t bytes) {
  if (bytes < 32 * 1024) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

static int stage_to_device(ccb_env *env, const double *host, size_t count, double *dev) {
  if (user_buffer_is_pinned(host, count * sizeof(double))) {
    // the call still returns only when the buffer has been read (the caller may reuse it at once)
    CCB_CUDA(cudaMemcpyAsync(dev, host, count * sizeof(double), cudaMemcpyHostToDevice, env->stream));
    CCB_CUDA(cudaEventRecord(env->ev_stage[env->stage_next], env->stream));
    CCB_CUDA(cudaEventSynchronize(env->ev_stage[env->stage_next]));
    return 0;
  }
  // host -> pinned slot -> device on the env stream, without waiting for the copy: two slots are used in
  // turn and a slot is only rewritten once the copy that last read it has finished
  if (count > 2 * (size_t)env->E * env->R) return fail("staging buffer too small");
  const int k = env->stage_next;
  env->stage_next ^= 1;
  CCB_CUDA(cudaEventSynchronize(env->ev_stage[k]));
  memcpy(env->h_stage[k], host, count * sizeof(double));
  CCB_CUDA(cudaMemcpyAsync(dev, env->h_stage[k], count * sizeof(double), cudaMemcpyHostToDevice, env->stream));
  CCB_CUDA(cudaEventRecord(env->ev_stage[k], env->stream));
  return 0;
}

extern "C" int ccb_env_reset(ccb_env *env, const double *positions) {
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  const EnvDev &d = env->d;
  const size_t pw = plane_words(d), er = (size_t)d.E * d.R;
  if (stage_to_device(env, positions, 2 * er, env->d_stage)) return 1;
  CCB_CUDA(cudaMemsetAsync(d.seenb, 0, sizeof(uint32_t) * pw * er, env->stream));
  if (d.xseenb) CCB_CUDA(cudaMemsetAsync(d.xseenb, 0, sizeof(uint32_t) * pw * er, env->stream));
  CCB_CUDA(cudaMemsetAsync(d.sysb, 0, sizeof(uint32_t) * pw * 2 * d.E, env->stream));
  CCB_CUDA(cudaMemsetAsync(d.status, 0, sizeof(int) * d.E, env->stream));
  k_init_positions<<<(int)((er + 255) / 256), 256, 0, env->stream>>>(d, env->d_stage);
  CCB_LAUNCH_CHECK();
  if (launch_mark(env, -1, -1, 1, 1)) return 1;
  env->initialised = true;
  if (state_changed(env)) return 1;
  return 0;
}

// RobotModel::Initialize + InitSetup for ONE environment of the batch (a fresh episode in that slot:
// the world map is set separately with ccb_env_set_world_*); positions = R x 2
extern "C" int ccb_env_reset_env(ccb_env *env, int e, const double *positions) {
  if (check_env_index(env, e)) return 1;
  if (!env->initialised) return fail("environment has no robot positions yet (call ccb_env_reset first)");
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  const EnvDev &d = env->d;
  const size_t pw = plane_words(d);
  if (stage_to_device(env, positions, 2 * (size_t)d.R, env->d_stage)) return 1;
  CCB_CUDA(cudaMemsetAsync(robot_plane(d, d.seenb, e, 0), 0, sizeof(uint32_t) * pw * d.R, env->stream));
  if (d.xseenb) CCB_CUDA(cudaMemsetAsync(robot_plane(d, d.xseenb, e, 0), 0, sizeof(uint32_t) * pw * d.R, env->stream));
  CCB_CUDA(cudaMemsetAsync(sys_plane(d, e, 0), 0, sizeof(uint32_t) * pw * 2, env->stream));
  CCB_CUDA(cudaMemsetAsync(d.status + e, 0, sizeof(int), env->stream));
  k_init_positions<<<(d.R + 255) / 256, 256, 0, env->stream>>>(d, env->d_stage, e);
  CCB_LAUNCH_CHECK();
  if (launch_mark(env, e, -1, 1, 1)) return 1;
  if (state_changed(env)) return 1;
  return 0;
}

extern "C" int ccb_env_step_actions(ccb_env *env, const double *actions, int *bad_env) {
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  const EnvDev &d = env->d;
  const size_t er = (size_t)d.E * d.R;
  if (bad_env) *bad_env = -1;
  if (stage_to_device(env, actions, 2 * er, env->d_stage)) return 1;
  if (launch_mark(env, -1, -1, 1, 1, env->d_stage)) return 1;  // move + mark, one warp per robot
  if (state_changed(env)) return 1;
  return 0;
}

// device-resident variant used by the roll-out (actions already on the device)
static int step_actions_dev(ccb_env *env, const double *actions_dev) {
  if (join_aux(env)) return 1;
  const EnvDev &d = env->d;
  const size_t er = (size_t)d.E * d.R;
  if (launch_mark(env, -1, -1, 1, 1, actions_dev)) return 1;  // move + mark, one warp per robot
  if (state_changed(env)) return 1;
  return 0;
}

__global__ void k_reset_seen_rows(EnvDev d, int e) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.R) return;
  int *box = d.seen_box + ((size_t)e * d.R + r) * 4;
  box[0] = box[2] = 0x7fffffff, box[1] = box[3] = 0;
}
__global__ void k_step_control(EnvDev d, int e, int r, double dx, double dy, double sp) {
  // robot_model.h:200-226 after the validity test, then SetRobotPosition
  const size_t er = (size_t)e * d.R + r;
  eigen_normalize(dx, dy);
  double lx = d.pos_l[2 * er], ly = d.pos_l[2 * er + 1];
  if (sp > kEps) lx = lx + dx * sp * d.dt, ly = ly + dy * sp * d.dt;
  double gx = lx + d.pos_s[2 * er], gy = ly + d.pos_s[2 * er + 1];
  clamp_global(d, gx, gy);
  d.pos_l[2 * er] = gx - d.pos_s[2 * er], d.pos_l[2 * er + 1] = gy - d.pos_s[2 * er + 1];
  d.pos_g[2 * er] = gx, d.pos_g[2 * er + 1] = gy;
}
__global__ void k_post_history(EnvDev d, int e, int robot) {
  // coverage_system.cpp:263-271 for all robots of env e (robot < 0) or one robot (:228-255)
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int r = robot < 0 ? t : robot;
  if (t >= (robot < 0 ? d.R : 1)) return;
  const size_t er = (size_t)e * d.R + r;
  history_post_step(d, e, r, d.pos_g[2 * er], d.pos_g[2 * er + 1]);
}

extern "C" int ccb_env_step_control(ccb_env *env, int e, int robot, double dir_x, double dir_y, double speed) {
  if (check_env_index(env, e) || check_robot(env, robot)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  const EnvDev &d = env->d;
  double sp = speed;
  if (sp > d.max_speed) sp = d.max_speed;
  double dn = std::sqrt(dir_x * dir_x + dir_y * dir_y);
  bool invalid_dir = dn < kEps && sp >= kEps;
  if (sp < 0 || invalid_dir) return 2;  // robot_model.h:208-216 (the reference throws)
  k_step_control<<<1, 1, 0, env->stream>>>(d, e, robot, dir_x, dir_y, sp);
  CCB_LAUNCH_CHECK();
  if (launch_mark(env, e, robot, 1, 0)) return 1;
  // PostStepCommands() for all robots (coverage_system.h:326-333)
  if (launch_mark(env, e, -1, 0, 1)) return 1;
  k_post_history<<<(d.R + 127) / 128, 128, 0, env->stream>>>(d, e, -1);
  CCB_LAUNCH_CHECK();
  if (state_changed(env)) return 1;
  return 0;
}

extern "C" int ccb_env_set_positions(ccb_env *env, int e, int robot, int local, const double *positions) {
  if (check_env_index(env, e)) return 1;
  if (robot >= 0 && check_robot(env, robot)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  const EnvDev &d = env->d;
  const int count = robot < 0 ? d.R : 1;
  if (stage_to_device(env, positions, 2 * (size_t)count, env->d_stage)) return 1;
  // all robots: PostStepCommands() (history for all); one robot: PostStepCommands(robot_id)
  k_set_positions<<<(count + 127) / 128, 128, 0, env->stream>>>(d, e, robot, local, env->d_stage, 1);
  CCB_LAUNCH_CHECK();
  if (launch_mark(env, e, robot, 1, 1)) return 1;
  if (state_changed(env)) return 1;
  return 0;
}

extern "C" int ccb_env_get_positions(ccb_env *env, double *out) {
  CCB_CUDA(cudaSetDevice(env->device));
  const size_t n = (size_t)env->E * env->R * 2;
  CCB_CUDA(cudaMemcpyAsync(env->h_pin, env->d.pos_g, n * sizeof(double), cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  memcpy(out, env->h_pin, n * sizeof(double));
  return 0;
}

extern "C" int ccb_env_clear_robot_maps(ccb_env *env, int e) {
  // coverage_system.h:399-403 -> robot_model.h:180-196: refill, then re-sense at the current position
  if (check_env_index(env, e)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  const EnvDev &d = env->d;
  const size_t pw = plane_words(d);
  CCB_CUDA(cudaMemsetAsync(robot_plane(d, d.seenb, e, 0), 0, sizeof(uint32_t) * pw * d.R, env->stream));
  k_reset_seen_rows<<<(d.R + 127) / 128, 128, 0, env->stream>>>(d, e);
  CCB_LAUNCH_CHECK();
  // the per-robot exploration maps are NOT cleared by the reference; only the robot-map bits are
  // re-marked here (mark_robot also ORs xseen bits that are already set)
  if (launch_mark(env, e, -1, 1, 0)) return 1;
  // the robot maps changed: window moments cached by the observation kernel (d_wsums) and roll-out graphs keyed on
  // them are stale even though no robot moved
  return state_changed(env);
}

extern "C" int ccb_env_clear_explored_idf(ccb_env *env, int e) {
  // coverage_system.h:405-408: explored_idf_map_ back to zero, the exploration map is kept
  if (check_env_index(env, e)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  CCB_CUDA(cudaMemsetAsync(sys_plane(env->d, e, 0), 0, sizeof(uint32_t) * plane_words(env->d), env->stream));
  return state_changed(env);
}

extern "C" int ccb_env_map_dim(const ccb_env *env, int kind) {
  if (kind <= CCB_MAP_ROBOT) return env->d.N;
  if (kind == CCB_MAP_SENSOR_VIEW) return env->d.S;
  return env->d.L;
}

extern "C" int ccb_env_get_map_dev(ccb_env *env, int e, int kind, int robot, float *out_dev) {
  if (check_env_index(env, e)) return 1;
  if (kind < 0 || kind > CCB_MAP_ROBOT_SYSTEM) return fail("unknown map kind");
  if (kind >= CCB_MAP_ROBOT && check_robot(env, robot)) return 1;
  if (kind >= CCB_MAP_SENSOR_VIEW && !env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  int dim = ccb_env_map_dim(env, kind);
  int blocks = std::min(148 * 8, (dim * dim + 255) / 256);
  k_materialize<<<blocks, 256, 0, env->stream>>>(env->d, e, kind, robot, out_dev);
  CCB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ccb_env_get_map(ccb_env *env, int e, int kind, int robot, float *out) {
  if (ccb_env_get_map_dev(env, e, kind, robot, env->d_scratch)) return 1;
  int dim = ccb_env_map_dim(env, kind);
  CCB_CUDA(cudaMemcpyAsync(out, env->d_scratch, sizeof(float) * dim * dim, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  return 0;
}

extern "C" int ccb_env_local_maps_dev(ccb_env *env, float *out_dev) {
  CCB_CUDA(cudaSetDevice(env->device));
  const EnvDev &d = env->d;
  k_raw_window_maps<<<dim3(d.L, d.R, d.E), 256, 0, env->stream>>>(d, CCB_MAP_LOCAL, out_dev);
  CCB_LAUNCH_CHECK();
  return 0;
}
extern "C" int ccb_env_obstacle_maps_dev(ccb_env *env, float *out_dev) {
  CCB_CUDA(cudaSetDevice(env->device));
  const EnvDev &d = env->d;
  k_raw_window_maps<<<dim3(d.L, d.R, d.E), 256, 0, env->stream>>>(d, CCB_MAP_OBSTACLE, out_dev);
  CCB_LAUNCH_CHECK();
  return 0;
}
// coverage_env_utils.py:313-335 get_weights + the normalised positions of :368-369, left on the device
extern "C" int ccb_env_graph_dev(ccb_env *env, float *weights_dev, float *pos_norm_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  k_graph<<<env->d.E, 256, 0, env->stream>>>(env->d, weights_dev, pos_norm_dev);
  CCB_LAUNCH_CHECK();
  return 0;
}
extern "C" int ccb_env_comm_maps_dev(ccb_env *env, int C, float *out_dev) {
  CCB_CUDA(cudaSetDevice(env->device));
  const EnvDev &d = env->d;
  k_comm_maps<<<dim3(d.R, d.E), 128, 0, env->stream>>>(d, C, out_dev, (size_t)2 * C * C, 0);
  CCB_LAUNCH_CHECK();
  return 0;
}

// torch's antialias bilinear index/weight computation (ATen UpSampleKernel.cpp,
// _compute_indices_weights_aa with the triangle filter), float arithmetic.
static int resize_weights(int L, int C, std::vector<int> &xmin, std::vector<int> &xsize, std::vector<float> &w) {
  const float scale = (float)L / (float)C;
  const float support = scale >= 1.f ? scale : 1.f;  // interp_size/2 * scale
  const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
  const int kmax = (int)std::ceil(support) * 2 + 1;
  xmin.assign(C, 0), xsize.assign(C, 0);
  w.assign((size_t)C * kmax, 0.f);
  for (int i = 0; i < C; ++i) {
    float center = scale * ((float)i + 0.5f);
    int lo = std::max((int)(center - support + 0.5f), 0);
    int hi = std::min((int)(center + support + 0.5f), L);
    xmin[i] = lo, xsize[i] = hi - lo;
    float total = 0.f;
    for (int j = 0; j < xsize[i]; ++j) {
      float x = ((float)(j + lo) - center + 0.5f) * invscale;
      x = std::fabs(x);
      float v = x < 1.f ? 1.f - x : 0.f;
      w[(size_t)i * kmax + j] = v;
      total += v;
    }
    for (int j = 0; j < xsize[i]; ++j)
      if (total != 0.f) w[(size_t)i * kmax + j] /= total;
  }
  return kmax;
}

// Host only (CPU tests against the installed torchvision): the taps of output index o are input indices
// [xmin[o], xmin[o] + xsize[o]) with weights w[o][0 .. kmax); returns kmax, or -1 when w_cap floats are too few.
extern "C" int ccb_selftest_resize_weights(int L, int C, int *xmin_out, int *xsize_out, float *w_out, int w_cap) {
  if (L < 1 || C < 1) return -1;
  std::vector<int> xmin, xsize;
  std::vector<float> w;
  const int kmax = resize_weights(L, C, xmin, xsize, w);
  if ((size_t)w_cap < w.size()) return -1;
  memcpy(xmin_out, xmin.data(), sizeof(int) * C);
  memcpy(xsize_out, xsize.data(), sizeof(int) * C);
  memcpy(w_out, w.data(), sizeof(float) * w.size());
  return kmax;
}

static int get_resize_table(ccb_env *env, int C, ResizeTab **out) {
  auto it = env->resize.find(C);
  if (it != env->resize.end()) {
    *out = &it->second;
    return 0;
  }
  const int L = env->d.L;
  std::vector<int> xmin, xsize;
  std::vector<float> w;
  const int kmax = resize_weights(L, C, xmin, xsize, w);
  ResizeTab t;
  t.C = C, t.kmax = kmax;
  CCB_CUDA(cudaMalloc(&t.d_xmin, sizeof(int) * C));
  CCB_CUDA(cudaMalloc(&t.d_xsize, sizeof(int) * C));
  CCB_CUDA(cudaMalloc(&t.d_w, sizeof(float) * C * kmax));
  CCB_CUDA(cudaMemcpy(t.d_xmin, xmin.data(), sizeof(int) * C, cudaMemcpyHostToDevice));
  CCB_CUDA(cudaMemcpy(t.d_xsize, xsize.data(), sizeof(int) * C, cudaMemcpyHostToDevice));
  CCB_CUDA(cudaMemcpy(t.d_w, w.data(), sizeof(float) * C * kmax, cudaMemcpyHostToDevice));
  if (L == 256 && C == 32) {
    std::vector<float> wf((size_t)C * 16, 0.f);
    for (int o = 0; o < C; ++o)
      for (int k = 0; k < 16; ++k) {
        int col = 8 * o - 4 + k;
        if (col >= xmin[o] && col < xmin[o] + xsize[o]) wf[(size_t)o * 16 + k] = w[(size_t)o * kmax + (col - xmin[o])];
      }
    t.interior_uniform = true;
    for (int o = 2; o < C - 1; ++o) t.interior_uniform = t.interior_uniform && memcmp(&wf[(size_t)o * 16], &wf[16], sizeof(float) * 16) == 0;
    if (t.interior_uniform) CCB_CUDA(cudaMemcpyToSymbol(c_ob_w, &wf[16], sizeof(float) * 16));  // (the same values whatever the environment)
    CCB_CUDA(cudaMalloc(&t.d_wfull8, sizeof(float) * C * 16));
    CCB_CUDA(cudaMemcpy(t.d_wfull8, wf.data(), sizeof(float) * C * 16, cudaMemcpyHostToDevice));
  }
  env->resize[C] = t;
  *out = &env->resize[C];
  return 0;
}

extern "C" int ccb_env_cnn_input_dev(ccb_env *env, int C, int use_comm, float *out_dev) {
  if (C < 1) return fail("CNN map size must be positive");
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  const EnvDev &d = env->d;
  ResizeTab *t = nullptr;
  if (get_resize_table(env, C, &t)) return 1;
  const int nch = use_comm ? 4 : 2;
  const size_t cc = (size_t)C * C, stride = cc * nch;
  ResizeTable tab{t->d_xmin, t->d_xsize, t->d_w, t->kmax};
  size_t smem = sizeof(float) * ((size_t)t->kmax * d.L + (size_t)t->kmax * C);
  if (ensure_smem(env, k_cnn_local_obstacle, smem)) return 1;
  // channel order of coverage_env_utils.py:264-276: [local, comm_x, comm_y, obstacle]
  const bool fused_ok = t->d_wfull8 && (d.N & 3) == 0 && !env->cnn_generic;
  const bool vor = fused_ok && env->n_decentralized > 0 && !env->no_wfuse && d.L == 256 && env->wsums_version != env->pos_version;
  const bool dense = d.fill != 0.f;
  const size_t oo = cc * (nch - 1);
  const dim3 g8(d.R, d.E);
  bool comm_done = false;
  if (fused_ok && t->interior_uniform && !env->obs_v1) {
    // item-ordered kernel: local + obstacle (+ communication) channels and, for DecentralizedCVT, the window moments
    const size_t sm = obs_smem_bytes(d.R);
    double *ws = vor ? env->d_wsums : nullptr;
    const int n_er = d.E * d.R;
    int *order = nullptr;
    if (env->obs_sorted && n_er >= 4 * 148) {  // (a few CTAs per SM at least: below that the order does not matter)
      if (!env->d_obs_order) CCB_CUDA(cudaMalloc(&env->d_obs_order, sizeof(int) * (2 * (size_t)n_er + 4)));
      int *cls = env->d_obs_order + n_er, *counts = env->d_obs_order + 2 * n_er;
      order = env->d_obs_order;
      CCB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * 4, env->stream));
      k_obs_classify<<<(n_er + 255) / 256, 256, 0, env->stream>>>(d, cls, counts);
      CCB_LAUNCH_CHECK();
      k_obs_order<<<(n_er + 255) / 256, 256, 0, env->stream>>>(n_er, cls, counts, order);
      CCB_LAUNCH_CHECK();
    }
#define CCB_OBS_LAUNCH(V, D)                                                                                                   \
  do {                                                                                                                         \
    if (ensure_smem(env, k_obs_fused<V, D>, sm)) return 1;                                                                      \
    k_obs_fused<V, D><<<g8, OB_THREADS, sm, env->stream>>>(d, t->d_wfull8, out_dev, stride, 0, oo, cc, use_comm, ws, order);    \
  } while (0)
    if (vor) {
      if (dense) CCB_OBS_LAUNCH(true, true);
      else CCB_OBS_LAUNCH(true, false);
      env->wsums_version = env->pos_version;
    } else if (dense) {
      CCB_OBS_LAUNCH(false, true);
    } else {
      CCB_OBS_LAUNCH(false, false);
    }
#undef CCB_OBS_LAUNCH
    comm_done = true;
  } else if (fused_ok) {
    // row-ordered kernel of round 1 (kept for A/B timing: ccb_test_set_option "obs_v1")
    const size_t smem8 = vor ? cnn8_vor_smem_bytes(d.R) : CNN8_SMEM_BYTES;
    if (ensure_smem(env, k_cnn_fused8<false, false>, smem8) || ensure_smem(env, k_cnn_fused8<false, true>, smem8) ||
        ensure_smem(env, k_cnn_fused8<true, false>, smem8) || ensure_smem(env, k_cnn_fused8<true, true>, smem8))
      return 1;
    if (vor) {
      if (dense) k_cnn_fused8<true, true><<<g8, CNN8_THREADS, smem8, env->stream>>>(d, t->d_wfull8, out_dev, stride, 0, oo, env->d_wsums);
      else k_cnn_fused8<true, false><<<g8, CNN8_THREADS, smem8, env->stream>>>(d, t->d_wfull8, out_dev, stride, 0, oo, env->d_wsums);
      env->wsums_version = env->pos_version;
    } else {
      if (dense) k_cnn_fused8<false, true><<<g8, CNN8_THREADS, smem8, env->stream>>>(d, t->d_wfull8, out_dev, stride, 0, oo, nullptr);
      else k_cnn_fused8<false, false><<<g8, CNN8_THREADS, smem8, env->stream>>>(d, t->d_wfull8, out_dev, stride, 0, oo, nullptr);
    }
  } else {
    k_cnn_local_obstacle<<<dim3(C, d.R, d.E), 256, smem, env->stream>>>(d, C, tab, out_dev, stride, 0, cc * (nch - 1));
  }
  CCB_LAUNCH_CHECK();
  if (use_comm && !comm_done) {
    k_comm_maps<<<dim3(d.R, d.E), 128, 0, env->stream>>>(d, C, out_dev, stride, cc);
    CCB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int ccb_env_comm_maps_cpp(ccb_env *env, int e, int C, float *out) {
  // coverage_system.cpp:177-199, evaluated on the host from the current positions (R^2 work)
  if (check_env_index(env, e)) return 1;
  std::vector<double> pos((size_t)env->E * env->R * 2);
  if (ccb_env_get_positions(env, pos.data())) return 1;
  const int R = env->R;
  const double *p = pos.data() + (size_t)e * R * 2;
  const double res = env->p.pResolution, cr = env->p.pCommunicationRange;
  std::fill(out, out + (size_t)R * 2 * C * C, 0.f);
  const double center = C / 2. - res / 2.;
  for (int r = 0; r < R; ++r) {
    float *ox = out + (size_t)r * 2 * C * C, *oy = ox + (size_t)C * C;
    for (int j = 0; j < R; ++j) {
      if (j == r) continue;
      double dx, dy;
      if (j > r) dx = p[2 * j] - p[2 * r], dy = p[2 * j + 1] - p[2 * r + 1];
      else dx = -(p[2 * r] - p[2 * j]), dy = -(p[2 * r + 1] - p[2 * j + 1]);
      if (!(std::sqrt(dx * dx + dy * dy) < cr)) continue;
      int ix = (int)std::round(dx * C / (cr * res * 2.) + center);
      int iy = (int)std::round(dy * C / (cr * res * 2.) + center);
      if (ix < 0 || ix >= C || iy < 0 || iy >= C) continue;
      ox[ix * C + iy] += (float)(dx / cr);
      oy[ix * C + iy] += (float)(dy / cr);
    }
  }
  return 0;
}

extern "C" int ccb_env_compute_voronoi(ccb_env *env) {
  CCB_CUDA(cudaSetDevice(env->device));
  return ensure_cells(env, false);
}

extern "C" int ccb_env_get_cells(ccb_env *env, int e, double *mass, double *centroid, double *site,
                                 double *sum_dist_sqr, int *nverts, double *verts) {
  if (check_env_index(env, e)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  const bool polys = nverts != nullptr || verts != nullptr;
  if (ensure_cells(env, polys)) return 1;
  const int R = env->R;
  const size_t off = (size_t)e * R;
  cudaStream_t st = env->stream;
  if (mass) CCB_CUDA(cudaMemcpyAsync(mass, env->cells.mass + off, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
  if (centroid) CCB_CUDA(cudaMemcpyAsync(centroid, env->cells.centroid + 2 * off, sizeof(double) * 2 * R, cudaMemcpyDeviceToHost, st));
  if (site) CCB_CUDA(cudaMemcpyAsync(site, env->d.pos_g + 2 * off, sizeof(double) * 2 * R, cudaMemcpyDeviceToHost, st));
  if (sum_dist_sqr) CCB_CUDA(cudaMemcpyAsync(sum_dist_sqr, env->cells.obj + off, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
  if (nverts) CCB_CUDA(cudaMemcpyAsync(nverts, env->cells.nverts + off, sizeof(int) * R, cudaMemcpyDeviceToHost, st));
  if (verts)
    CCB_CUDA(cudaMemcpyAsync(verts, env->cells.verts + off * CCB_MAX_CELL_VERTS * 2,
                             sizeof(double) * R * CCB_MAX_CELL_VERTS * 2, cudaMemcpyDeviceToHost, st));
  CCB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// mass, centroid and sum |c - site|^2 v of every cell of every environment in one go (the Lloyd
// iterations of NearOptimalCVT read all tries at once); any of the outputs may be null
extern "C" int ccb_env_get_cells_all(ccb_env *env, double *mass, double *centroid, double *sum_dist_sqr) {
  CCB_CUDA(cudaSetDevice(env->device));
  if (ensure_cells(env, false)) return 1;
  const size_t er = (size_t)env->E * env->R;
  cudaStream_t st = env->stream;
  if (mass) CCB_CUDA(cudaMemcpyAsync(mass, env->cells.mass, sizeof(double) * er, cudaMemcpyDeviceToHost, st));
  if (centroid) CCB_CUDA(cudaMemcpyAsync(centroid, env->cells.centroid, sizeof(double) * 2 * er, cudaMemcpyDeviceToHost, st));
  if (sum_dist_sqr) CCB_CUDA(cudaMemcpyAsync(sum_dist_sqr, env->cells.obj, sizeof(double) * er, cudaMemcpyDeviceToHost, st));
  CCB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int ccb_env_voronoi_labels(ccb_env *env, int src, int32_t *out /* E x N x N, host */) {
  // parity-test accessor: the label the fused kernel assigns to every cell (never stored otherwise)
  if (!env->initialised) return fail("environment has no robot positions yet");
  if (src != SRC_WORLD && src != SRC_EXPLORED) return fail("labels: src must be 0 (world) or 1 (explored)");
  CCB_CUDA(cudaSetDevice(env->device));
  const EnvDev &d = env->d;
  const size_t n = cell_count(d) * d.E;
  int *d_labels = nullptr;
  double *d_tmp = nullptr;
  CCB_CUDA(cudaMalloc(&d_labels, sizeof(int) * n));
  CCB_CUDA(cudaMalloc(&d_tmp, sizeof(double) * 4 * d.E * d.R));
  CCB_CUDA(cudaMemsetAsync(d_labels, 0xff, sizeof(int) * n, env->stream));
  int st = launch_voronoi_global(env, src, d_tmp, d_labels);
  if (!st) {
    cudaError_t err = cudaMemcpyAsync(out, d_labels, sizeof(int) * n, cudaMemcpyDeviceToHost, env->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(env->stream);
    if (err != cudaSuccess) st = fail("CUDA error", cudaGetErrorString(err));
  }
  cudaFree(d_labels);
  cudaFree(d_tmp);
  return st;
}

// ---- row bands (multi-GPU tiling of one large world) ---------------------------------
extern "C" int ccb_env_voronoi_partials_dev(ccb_env *env, int src, int row0, int row1, double *sums_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  if (src != SRC_WORLD && src != SRC_EXPLORED) return fail("partials: src must be 0 (world) or 1 (explored)");
  if (row0 < 0 || row1 > env->d.N || row0 > row1) return fail("partials: bad row band");
  CCB_CUDA(cudaSetDevice(env->device));
  return launch_voronoi_global(env, src, sums_dev, nullptr, row0, row1);
}

extern "C" int ccb_env_finalize_from_partials_dev(ccb_env *env, const double *sums_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  if (sums_dev != env->d_sums)
    CCB_CUDA(cudaMemcpyAsync(env->d_sums, sums_dev, sizeof(double) * 4 * env->E * env->R, cudaMemcpyDeviceToDevice, env->stream));
  if (launch_finalize(env, env->d_sums, env->cells, 0)) return 1;
  env->cells_version = env->pos_version;
  env->cells_have_polys = false;
  CCB_CUDA(cudaEventRecord(env->ev_cells, env->stream));
  return 0;
}

// DecentralizedCVT over several GPUs: the robots' windows are independent, so they are dealt to the ranks by ROBOT.
// Window moments (decentralized_cvt.h:85-143) of robots [robot0, robot1) into sums_dev (E x R x 4, zeros elsewhere)...
extern "C" int ccb_env_window_partials_dev(ccb_env *env, int robot0, int robot1, double *sums_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  const EnvDev &d = env->d;
  if (robot0 < 0 || robot1 > d.R || robot0 > robot1) return fail("window partials: bad robot range");
  if (!(d.L <= VW_MAXL && (d.N & 3) == 0)) return fail("window partials: needs pLocalMapSize <= 256 and pWorldMapSize % 4 == 0");
  CCB_CUDA(cudaSetDevice(env->device));
  if (join_aux(env)) return 1;
  CCB_CUDA(cudaMemsetAsync(sums_dev, 0, sizeof(double) * 4 * d.E * d.R, env->stream));
  if (robot1 > robot0) {
    const size_t smem = vw_smem_bytes(d.R);
    if (ensure_smem(env, k_voronoi_window, smem)) return 1;
    k_voronoi_window<<<dim3(robot1 - robot0, d.E), VW_THREADS, smem, env->stream>>>(d, sums_dev, robot0);
    CCB_LAUNCH_CHECK();
  }
  return 0;
}
// ... and, once all-reduced, handed back as the window moments of the current positions: the next
// DecentralizedCVT ComputeActions uses them instead of reducing the windows itself
extern "C" int ccb_env_set_window_sums_dev(ccb_env *env, const double *sums_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  CCB_CUDA(cudaMemcpyAsync(env->d_wsums, sums_dev, sizeof(double) * 4 * env->E * env->R, cudaMemcpyDeviceToDevice, env->stream));
  env->wsums_version = env->pos_version;
  return 0;
}

extern "C" int ccb_env_set_band_mode(ccb_env *env, int on) {
  env->band_mode = on != 0;
  env->cells_version = 0;
  return 0;
}

// ---- one-shot all-reduce over peer memory (csrc/peer.cuh) --------------------------------------------------
extern "C" size_t ccb_peer_buffer_bytes(int world, size_t count) { return peer_buffer_bytes(world, count); }
extern "C" int ccb_peer_allreduce_sum(int device, void *stream, const double *in_dev, double *out_dev, size_t count,
                                      const void *const *peer_bufs, int rank, int world, void *state_dev) {
  if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world) return fail("ccb_peer_allreduce_sum: bad rank / world size");
  if (!in_dev || !out_dev || !peer_bufs || !state_dev || count == 0) return fail("ccb_peer_allreduce_sum: null argument");
  CCB_CUDA(cudaSetDevice(device));
  PeerArgs a;
  for (int p = 0; p < world; ++p) {
    if (!peer_bufs[p]) return fail("ccb_peer_allreduce_sum: a peer buffer is not mapped");
    a.buf[p] = static_cast<unsigned char *>(const_cast<void *>(peer_bufs[p]));
  }
  a.rank = rank, a.world = world, a.count = count, a.state = static_cast<unsigned long long *>(state_dev);
  k_peer_allreduce<<<1, PEER_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, in_dev, out_dev);
  CCB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ccb_env_step_actions_dev(ccb_env *env, const double *actions_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  return step_actions_dev(env, actions_dev);
}

static int objective_dev(ccb_env *env, double *out_dev) {
  if (ensure_cells(env, false)) return 1;
  k_objective<<<env->E, 32, 0, env->stream>>>(env->d, env->cells.obj, out_dev);
  CCB_LAUNCH_CHECK();
  return 0;
}

// The objective chain on `aux`: it starts as soon as the last state change is done and runs beside
// whatever `stream` does next (observation kernels, the next step's window Voronoi).
static int objective_aux(ccb_env *env, double *out_dev) {
  if (env->no_aux) return objective_dev(env, out_dev);
  CCB_CUDA(cudaStreamWaitEvent(env->aux, env->ev_state, 0));
  CCB_CUDA(cudaStreamWaitEvent(env->aux, env->ev_cells, 0));
  CCB_CUDA(cudaStreamWaitEvent(env->aux, env->ev_vg, 0));  // a CentralizedCVT pass on `stream` may still be using the scratch
  int rc;
  {
    AuxScope scope(env);
    rc = objective_dev(env, out_dev);
  }
  if (rc) return rc;
  CCB_CUDA(cudaEventRecord(env->ev_aux, env->aux));
  env->aux_pending = true;
  return 0;
}

// GetObjectiveValue for every environment, left on the device (E doubles), on the env stream
extern "C" int ccb_env_objective_dev(ccb_env *env, double *out_dev) {
  CCB_CUDA(cudaSetDevice(env->device));
  return objective_dev(env, out_dev);
}

extern "C" int ccb_env_objective(ccb_env *env, double *out) {
  CCB_CUDA(cudaSetDevice(env->device));
  if (objective_aux(env, env->d_objective)) return 1;
  cudaStream_t st = env->no_aux ? env->stream : env->aux;
  CCB_CUDA(cudaMemcpyAsync(env->h_pin, env->d_objective, sizeof(double) * env->E, cudaMemcpyDeviceToHost, st));
  CCB_CUDA(cudaStreamSynchronize(st));
  memcpy(out, env->h_pin, sizeof(double) * env->E);
  return 0;
}

extern "C" int ccb_env_voronoi_features(ccb_env *env, double *out) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  k_voronoi_features<<<dim3(env->R, env->E), 256, sizeof(double) * 2 * env->R, env->stream>>>(env->d, env->d_feat, 0);
  CCB_LAUNCH_CHECK();
  const size_t n = (size_t)env->E * env->R * 7;
  CCB_CUDA(cudaMemcpyAsync(env->h_pin, env->d_feat, sizeof(double) * n, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  memcpy(out, env->h_pin, sizeof(double) * n);
  return 0;
}

extern "C" int ccb_env_voronoi_features_dev(ccb_env *env, double *out_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  k_voronoi_features<<<dim3(env->R, env->E), 256, sizeof(double) * 2 * env->R, env->stream>>>(env->d, out_dev, 0);
  CCB_LAUNCH_CHECK();
  return 0;
}

// coverage_system.cpp:648-676 GetLocalVoronoiFeatures for every robot: E x R x 7, host
extern "C" int ccb_env_local_voronoi_features(ccb_env *env, double *out) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  k_voronoi_features<<<dim3(env->R, env->E), 256, sizeof(double) * 2 * env->R, env->stream>>>(env->d, env->d_feat, 1);
  CCB_LAUNCH_CHECK();
  const size_t n = (size_t)env->E * env->R * 7;
  CCB_CUDA(cudaMemcpyAsync(env->h_pin, env->d_feat, sizeof(double) * n, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  memcpy(out, env->h_pin, sizeof(double) * n);
  return 0;
}

// Host run of the kernel's frontier queue (the same expl_* code, compiled for the host): lets the CPU
// tests check it against std::priority_queue without a GPU.  Not a compute path: no map is touched.
extern "C" int ccb_selftest_frontier_select(const double *values, int n, int num_frontiers, int *order_out) {
  if (num_frontiers < 1 || num_frontiers > EXPL_MAXK) return -1;
  ExplHeap h;
  expl_select(h, values, n, num_frontiers);
  int count = 0;
  while (h.n > 0) {
    order_out[count++] = h.id[0];
    expl_pop(h);
  }
  return count;
}

// Host run of make_window (common.cuh; map_utils.h:51-87 as every window kernel uses it): out6 = ox, oy, x0, x1, y0, y1
extern "C" int ccb_selftest_make_window(double res, double px, double py, int sub, int map_size, int *out6) {
  const Window w = make_window(res, px, py, sub, map_size);
  out6[0] = w.ox, out6[1] = w.oy, out6[2] = w.x0, out6[3] = w.x1, out6[4] = w.y0, out6[5] = w.y1;
  return 0;
}

// Host run of k_vg_partition for one environment: the row sweep is the kernel's own vg_partition_row compiled
// for the host; the CTA preamble (sort by (py, index), twins, band cull, 1 / (2 dy) table) is restated
// sequentially with the kernel's expressions, band by band of VG_PART_THREADS rows.  Lets the CPU tests hold the
// envelope intervals of every unflagged row against the stated rule without a GPU.  Not a compute path: no map
// is read.  table[R][N] = lo | hi << 16, flag_rows[<= N], span[R][2]; returns the number of flagged rows.
extern "C" int ccb_selftest_vg_partition(const double *pos, int R, int N, double res, int row0, int row1, int band_rows,
                                         uint32_t *table, int *flag_rows, int *span) {
  if (band_rows <= 0) band_rows = VG_PART_THREADS;  // rows per CTA band: VG_PART_THREADS x CCB_VG_PART_RPT
  if (R < 1 || R > VG_MAXR || N < 1 || N >= 65536 || row0 < 0 || row1 > N || row0 > row1 || !(res > 0.0)) return -1;
  VgArgs a{};
  a.d.E = 1, a.d.R = R, a.d.N = N, a.d.res = res;
  a.row0 = row0, a.row1 = row1;
  a.table = table;
  int nflag = 0;
  a.flag_list = flag_rows, a.nflag = &nflag;
  for (int k = 0; k < R; ++k) span[2 * k] = 0x7fffffff, span[2 * k + 1] = 0;
  std::vector<int> t_ord(R), s_ord(R), s_rmin(R), s_rmax(R);
  std::vector<double> t_x(R), t_y(R), s_x(R), s_y(R), s_inv((size_t)R * R);
  std::vector<unsigned char> keep(R);
  for (int k = 0; k < R; ++k) {
    const double yk = pos[2 * k + 1];
    int rank = 0;
    for (int j = 0; j < R; ++j) {
      const double yj = pos[2 * j + 1];
      rank += (yj < yk || (yj == yk && j < k)) ? 1 : 0;
    }
    t_ord[rank] = k, t_x[rank] = pos[2 * k], t_y[rank] = yk;
  }
  for (int rlo = row0; rlo < row1; rlo += band_rows) {
    const int rhi = std::min(rlo + band_rows, row1) - 1;
    const double cxs[2] = {cell_centre(rlo, res), cell_centre(rhi, res)};
    const double cys[2] = {cell_centre(0, res), cell_centre(N - 1, res)};
    for (int q = 0; q < R; ++q) {
      bool dup = false;
      for (int p = 0; p < R; ++p) dup = dup || (t_ord[p] < t_ord[q] && t_x[p] == t_x[q] && t_y[p] == t_y[q]);
      keep[q] = dup ? 0 : 1;
    }
    for (int q = 0; q < R; ++q) {
      double fq[4];
      for (int c = 0; c < 4; ++c) fq[c] = d2_exact(cxs[c & 1], cys[c >> 1], t_x[q], t_y[q]);
      for (int p = 0; p < R; ++p) {
        if (p == q || (t_x[p] == t_x[q] && t_y[p] == t_y[q])) continue;
        bool all = true;
        for (int c = 0; c < 4; ++c) {
          const double fp = d2_exact(cxs[c & 1], cys[c >> 1], t_x[p], t_y[p]);
          all = all && (fq[c] - fp > 1e-7 * (1.0 + fp));
        }
        if (all) {
          keep[q] = 0;
          break;
        }
      }
    }
    int Rs = 0;
    for (int q = 0; q < R; ++q)
      if (keep[q]) s_ord[Rs] = t_ord[q], s_x[Rs] = t_x[q], s_y[Rs] = t_y[q], ++Rs;
    for (int idx = 0; idx < Rs * Rs; ++idx) {
      const int p = idx / Rs, q = idx - p * Rs;
      const double dy = s_y[q] - s_y[p];
      s_inv[idx] = dy > 0.0 ? 1.0 / (2.0 * dy) : 0.0;
    }
    for (int k = 0; k < R; ++k) s_rmin[k] = 0x7fffffff, s_rmax[k] = 0;
    for (int row = rlo; row <= rhi; ++row)
      vg_partition_row<false>(a, 0, row, true, Rs, s_ord.data(), s_x.data(), s_y.data(), s_inv.data(), s_rmin.data(), s_rmax.data(), nullptr);
    for (int k = 0; k < R; ++k)
      if (s_rmax[k] > s_rmin[k]) {
        span[2 * k] = std::min(span[2 * k], s_rmin[k]);
        span[2 * k + 1] = std::max(span[2 * k + 1], s_rmax[k]);
      }
  }
  return nflag;
}

// Host run of the window kernels' own-cell row interval (vw_row_interval of k_voronoi_window / k_cnn_fused8<VOR>,
// the same code compiled for the host): px / py = the robot first, then its K neighbours.  For every row of
// [row0, row1): lo[], hi[] (columns y0 + lo .. y0 + hi - 1) and exact[] = 1 when the kernel applies the per-cell
// rule inside that interval.  Also returns the kernel's mask words for the interval of the first row through
// vw_range_bits when words_out is given (9 words from column word `wbase`).  Not a compute path.
extern "C" int ccb_selftest_vw_row_interval(double res, int row0, int row1, int y0, int ncols, int K, const double *px,
                                            const double *py, int *lo, int *hi, int *exact, int wbase, uint32_t *words_out) {
  if (K < 0 || ncols < 0 || row1 < row0 || !(res > 0.0)) return -1;
  for (int row = row0; row < row1; ++row) {
    int l = 0, h = ncols;
    bool ex = false;
    if (K > 0) ex = vw_row_interval(res, row, y0, ncols, K, px, py, l, h);
    lo[row - row0] = l, hi[row - row0] = h > l ? h : l, exact[row - row0] = ex ? 1 : 0;
  }
  if (words_out && row1 > row0)
    for (int j = 0; j < VW_WORDS; ++j) words_out[j] = vw_range_bits(y0 + lo[0], y0 + hi[0], (wbase + j) << 5);
  return 0;
}

// coverage_system.h:603-610 GetRobotExplorationFeatures for every robot: E x R x 2 K, host
extern "C" int ccb_env_exploration_features(ccb_env *env, int num_frontiers, double *out) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  const EnvDev &d = env->d;
  const int K = num_frontiers;
  if (K < 1 || K > EXPL_MAXK) return fail("exploration features: pNumFrontiers must be in [1, 64]");
  // the kernel's bounds, from the same quantities: candidates per axis, union of their sensor windows
  const double tsd = d.max_speed * d.dt * d.res;
  const double per_axis = 8.0 * tsd / d.res + 2.0;
  if (!(per_axis <= (double)EXPL_MAXC)) return fail("exploration features: more than 96 candidate points per axis (4 x pMaxRobotSpeed x pTimeStep too large)");
  const size_t side = (size_t)per_axis + (size_t)d.S + 2;
  const size_t words = (2 * side * side + 1) & ~(size_t)1;
  const size_t smem = sizeof(int) * words + sizeof(double) * (size_t)per_axis * (size_t)per_axis;
  if (smem > 200 * 1024) return fail("exploration features: candidate windows do not fit in shared memory (pSensorSize too large)");
  CCB_CUDA(cudaSetDevice(env->device));
  const size_t n = (size_t)d.E * d.R * 2 * K;
  if (n > env->feat_cap) {
    if (env->d_xfeat) cudaFree(env->d_xfeat);
    env->d_xfeat = nullptr;
    CCB_CUDA(cudaMalloc(&env->d_xfeat, sizeof(double) * n));
    env->feat_cap = n;
  }
  CCB_CUDA(cudaFuncSetAttribute(k_exploration_features, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_exploration_features<<<dim3(d.R, d.E), 256, smem, env->stream>>>(d, K, env->d_xfeat);
  CCB_LAUNCH_CHECK();
  CCB_CUDA(cudaMemcpyAsync(out, env->d_xfeat, sizeof(double) * n, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  return 0;
}

extern "C" int ccb_env_exploration_ratios(ccb_env *env, double *out) {
  CCB_CUDA(cudaSetDevice(env->device));
  const EnvDev &d = env->d;
  CCB_CUDA(cudaMemsetAsync(env->d_ratio, 0, sizeof(double) * 3 * d.E, env->stream));
  k_ratio_sums<<<dim3(148, d.E), 256, 0, env->stream>>>(d, env->d_ratio);
  CCB_LAUNCH_CHECK();
  CCB_CUDA(cudaMemcpyAsync(env->h_pin, env->d_ratio, sizeof(double) * 3 * d.E, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  for (int e = 0; e < d.E; ++e) {
    const double *s = env->h_pin + 3 * e;
    // coverage_system.h:557-571; exploration_map_ counts cells NOT yet explored
    out[2 * e] = 1.0 - ((double)d.N * d.N - s[0]) / ((double)d.N * d.N);
    out[2 * e + 1] = s[1] / s[2];
  }
  return 0;
}

__global__ void k_check_osc(EnvDev d, int e, int r, int *out) { *out = check_oscillation(d, e, r) ? 1 : 0; }
extern "C" int ccb_env_check_oscillation(ccb_env *env, int e, int robot, int *out) {
  if (check_env_index(env, e) || check_robot(env, robot)) return 1;
  CCB_CUDA(cudaSetDevice(env->device));
  int *d_flag = reinterpret_cast<int *>(env->d_ratio);
  k_check_osc<<<1, 1, 0, env->stream>>>(env->d, e, robot, d_flag);
  CCB_LAUNCH_CHECK();
  CCB_CUDA(cudaMemcpyAsync(env->h_pin, d_flag, sizeof(int), cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  *out = *reinterpret_cast<int *>(env->h_pin);
  return 0;
}

// diagnostics: rows the last site-ordered whole-map Voronoi handed to the exact per-cell rule
extern "C" int ccb_env_flagged_rows(ccb_env *env, int *count, int *rows, int cap) {
  CCB_CUDA(cudaSetDevice(env->device));
  *count = 0;
  if (!env->d_vg_flag) return 0;
  CCB_CUDA(cudaStreamSynchronize(env->aux));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  CCB_CUDA(cudaMemcpy(count, env->d_vg_flag, sizeof(int), cudaMemcpyDeviceToHost));
  const int n = std::min(*count, cap);
  if (rows && n > 0) CCB_CUDA(cudaMemcpy(rows, env->d_vg_flag + 1, sizeof(int) * n, cudaMemcpyDeviceToHost));
  return 0;
}

// Diagnostics for the roofline arithmetic (include/ccb200_test.h): what = 0 -> cells of the robots' local windows
// that their robot maps have seen (the cells the observation / window-Voronoi kernels must read), summed over all
// robots and environments; what = 1 -> (site, row) pairs with cells in the last site-ordered partition.
__global__ void k_count_window_seen(EnvDev d, unsigned long long *out) {
  const int r = blockIdx.x, e = blockIdx.y;
  const size_t er = (size_t)e * d.R + r;
  const Window w = make_window(d.res, d.pos_g[2 * er], d.pos_g[2 * er + 1], d.L, d.N);
  const uint32_t *plane = robot_plane(d, d.seenb, e, r);
  unsigned long long n = 0;
  const int w0 = w.y0 >> 5, w1 = (w.y1 + 31) >> 5;
  for (int t = threadIdx.x; t < (w.x1 - w.x0) * (w1 - w0); t += blockDim.x) {
    const int row = w.x0 + t / (w1 - w0), wd = w0 + t % (w1 - w0);
    n += __popc(__ldg(plane + (size_t)row * d.NW + wd) & vw_range_bits(w.y0, w.y1, wd << 5));
  }
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, n);
}
__global__ void k_count_pairs(const uint32_t *table, size_t n, unsigned long long *out) {
  unsigned long long c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t iv = __ldg(table + i);
    c += (iv >> 16) > (iv & 0xffffu) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
extern "C" int ccb_test_count(ccb_env *env, int what, uint64_t *out) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  CCB_CUDA(cudaStreamSynchronize(env->aux));
  unsigned long long *d_n = reinterpret_cast<unsigned long long *>(env->d_ratio);
  CCB_CUDA(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), env->stream));
  const EnvDev &d = env->d;
  if (what == 0) k_count_window_seen<<<dim3(d.R, d.E), 256, 0, env->stream>>>(d, d_n);
  else if (what == 1 && env->d_vg_table) k_count_pairs<<<148 * 8, 256, 0, env->stream>>>(env->d_vg_table, (size_t)d.E * d.R * d.N, d_n);
  else return fail("ccb_test_count: unknown counter (or no partition yet)");
  CCB_LAUNCH_CHECK();
  unsigned long long h = 0;
  CCB_CUDA(cudaMemcpyAsync(&h, d_n, sizeof h, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  *out = h;
  return 0;
}

extern "C" int ccb_env_sync(ccb_env *env) {
  CCB_CUDA(cudaSetDevice(env->device));
  CCB_CUDA(cudaStreamSynchronize(env->aux));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  return 0;
}

// =================================================================================================
// CVT controllers
// =================================================================================================
struct ccb_ctrl {
  ccb_env *env = nullptr;
  int kind = 0;
  double *d_goals = nullptr;    // [E][R][2]
  double *d_actions = nullptr;  // [E][R][2]
  int *d_conv = nullptr;        // [E]
  double *d_sums = nullptr;     // [E][R][4] (centralized / decentralized)
  CellsOut cells{};             // centralized
  // A roll-out call is a fixed sequence of launches on two streams: the second identical call is
  // captured into a CUDA graph and later ones replay it (one launch per call instead of ~15 per step).
  struct RolloutGraph {
    int steps = 0, with_objective = 0, C = 0, use_comm = 0, entry_valid = 0;
    double *obj = nullptr;
    float *obs = nullptr;
    size_t stride = 0;
    int seen = 0;                 // eager calls with this signature so far
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;        // kernel launches inside the graph
    uint64_t pos_delta = 0;       // host bookkeeping the replay must reproduce
    bool cells_recomputed = false;
    uint64_t cells_lag = 0;
    bool wsums_valid_after = false;
  } rg;
};

// `centroids` (optional): where the goals are left when they are a cell set's centroids -- the caller (k_actions)
// then copies them into d_goals itself instead of a separate device-to-device copy.
static int ctrl_compute_goals(ccb_ctrl *c, const double **centroids = nullptr) {
  ccb_env *env = c->env;
  const EnvDev &d = env->d;
  const size_t er = (size_t)d.E * d.R;
  const double *src = nullptr;
  if (c->kind == CCB_CLAIRVOYANT) {
    // clairvoyant_cvt.h:89-98: Voronoi on the world map == the env-level cells
    if (ensure_cells(env, false)) return 1;
    src = env->cells.centroid;
  } else if (c->kind == CCB_CENTRALIZED && env->band_mode) {
    // row-band mode: the caller all-reduced the explored-IDF partials into the env-level cells
    if (ensure_cells(env, false)) return 1;
    src = env->cells.centroid;
  } else if (c->kind == CCB_CENTRALIZED) {
    // centralized_cvt.h:90-100: Voronoi on the explored-IDF map
    if (launch_voronoi_global(env, SRC_EXPLORED, c->d_sums)) return 1;
    if (launch_finalize(env, c->d_sums, c->cells, 0)) return 1;
    src = c->cells.centroid;
  }
  if (src) {
    if (centroids) *centroids = src;
    else CCB_CUDA(cudaMemcpyAsync(c->d_goals, src, sizeof(double) * 2 * er, cudaMemcpyDeviceToDevice, env->stream));
  } else {
    if (launch_voronoi_single(env, c->d_sums)) return 1;
    const int warps = d.E * d.R;
    k_finalize_single<<<(warps + FIN_WARPS - 1) / FIN_WARPS, FIN_WARPS * 32, 0, env->stream>>>(d, c->d_sums, c->d_goals);
    CCB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" ccb_ctrl *ccb_ctrl_create(ccb_env *env, int kind) {
  if (kind < CCB_CLAIRVOYANT || kind > CCB_DECENTRALIZED) {
    fail("unknown controller kind");
    return nullptr;
  }
  if (!env->initialised) {
    fail("environment has no robot positions yet (call ccb_env_reset)");
    return nullptr;
  }
  cudaSetDevice(env->device);
  auto c = std::make_unique<ccb_ctrl>();
  c->env = env, c->kind = kind;
  const size_t er = (size_t)env->E * env->R;
  bool ok = cudaMalloc(&c->d_goals, sizeof(double) * 2 * er) == cudaSuccess &&
            cudaMalloc(&c->d_actions, sizeof(double) * 2 * er) == cudaSuccess &&
            cudaMalloc(&c->d_conv, sizeof(int) * env->E) == cudaSuccess &&
            cudaMalloc(&c->d_sums, sizeof(double) * 4 * er) == cudaSuccess;
  if (ok && kind == CCB_CENTRALIZED) ok = alloc_cells(c->cells, env->E, env->R, false) == 0;
  if (!ok) {
    fail("cudaMalloc failed (controller)");
    ccb_ctrl_destroy(c.release());
    return nullptr;
  }
  if (kind == CCB_DECENTRALIZED) env->n_decentralized++;
  cudaMemsetAsync(c->d_actions, 0, sizeof(double) * 2 * er, env->stream);
  cudaMemsetAsync(c->d_conv, 0, sizeof(int) * env->E, env->stream);
  if (ctrl_compute_goals(c.get())) {  // the reference constructors call ComputeGoals()
    ccb_ctrl_destroy(c.release());
    return nullptr;
  }
  return c.release();
}

extern "C" void ccb_ctrl_destroy(ccb_ctrl *c) {
  if (!c) return;
  cudaSetDevice(c->env->device);
  cudaStreamSynchronize(c->env->stream);
  if (c->rg.exec) cudaGraphExecDestroy(c->rg.exec);
  if (c->kind == CCB_DECENTRALIZED && c->env->n_decentralized > 0) c->env->n_decentralized--;
  cudaFree(c->d_goals), cudaFree(c->d_actions), cudaFree(c->d_conv), cudaFree(c->d_sums);
  free_cells(c->cells);
  delete c;
}

static int ctrl_compute_actions_async(ccb_ctrl *c) {
  ccb_env *env = c->env;
  const EnvDev &d = env->d;
  const double *centroids = nullptr;
  if (ctrl_compute_goals(c, &centroids)) return 1;
  int threads = std::min(1024, ((d.R + 31) / 32) * 32);
  k_actions<<<d.E, threads, 0, env->stream>>>(d, centroids ? centroids : c->d_goals, c->d_actions, c->d_conv,
                                              centroids ? c->d_goals : nullptr);
  CCB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ccb_ctrl_compute_actions(ccb_ctrl *c) {
  CCB_CUDA(cudaSetDevice(c->env->device));
  return ctrl_compute_actions_async(c);
}

static int d2h_doubles(ccb_env *env, const double *dev, double *out, size_t n) {
  if (n * sizeof(double) > env->h_pin_bytes) return fail("staging buffer too small");
  const bool direct = user_buffer_is_pinned(out, n * sizeof(double));
  CCB_CUDA(cudaMemcpyAsync(direct ? out : env->h_pin, dev, n * sizeof(double), cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  if (!direct) memcpy(out, env->h_pin, n * sizeof(double));
  return 0;
}

extern "C" int ccb_ctrl_get_actions(ccb_ctrl *c, double *out) {
  CCB_CUDA(cudaSetDevice(c->env->device));
  ccb_env *env = c->env;
  const size_t n = (size_t)env->E * env->R * 2;
  const bool dec = c->kind == CCB_DECENTRALIZED;
  // the actions and (DecentralizedCVT) the environments' status words in ONE round trip
  if (n * sizeof(double) + (dec ? sizeof(int) * env->E : 0) > env->h_pin_bytes) return fail("staging buffer too small");
  const bool direct = user_buffer_is_pinned(out, n * sizeof(double));
  CCB_CUDA(cudaMemcpyAsync(direct ? out : env->h_pin, c->d_actions, n * sizeof(double), cudaMemcpyDeviceToHost, env->stream));
  int *h_status = reinterpret_cast<int *>(env->h_pin + n);
  if (dec) CCB_CUDA(cudaMemcpyAsync(h_status, env->d.status, sizeof(int) * env->E, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  if (!direct) memcpy(out, env->h_pin, n * sizeof(double));
  if (dec) {
    // surface the reference's "Goal out of bounds" exception (decentralized_cvt.h:127-141)
    for (int e = 0; e < env->E; ++e)
      if (h_status[e] & 1) {
        fail("Goal out of bounds: this should not have happened for convex environments");
        return 2;
      }
  }
  return 0;
}
extern "C" int ccb_ctrl_get_goals(ccb_ctrl *c, double *out) {
  CCB_CUDA(cudaSetDevice(c->env->device));
  return d2h_doubles(c->env, c->d_goals, out, (size_t)c->env->E * c->env->R * 2);
}
extern "C" int ccb_ctrl_is_converged(ccb_ctrl *c, int *out) {
  CCB_CUDA(cudaSetDevice(c->env->device));
  ccb_env *env = c->env;
  CCB_CUDA(cudaMemcpyAsync(env->h_pin, c->d_conv, sizeof(int) * env->E, cudaMemcpyDeviceToHost, env->stream));
  CCB_CUDA(cudaStreamSynchronize(env->stream));
  memcpy(out, env->h_pin, sizeof(int) * env->E);
  return 0;
}
extern "C" const double *ccb_ctrl_actions_dev(const ccb_ctrl *c) { return c->d_actions; }

static int rollout_obs_eager(ccb_ctrl *c, int steps, int with_objective, double *objective_dev_out, int C, int use_comm,
                             float *obs_dev, size_t obs_step_stride) {
  ccb_env *env = c->env;
  // fork points of this call (inside a graph capture they must belong to the capture)
  CCB_CUDA(cudaEventRecord(env->ev_state, env->stream));
  CCB_CUDA(cudaEventRecord(env->ev_cells, env->stream));
  CCB_CUDA(cudaEventRecord(env->ev_vg, env->stream));
  for (int s = 0; s < steps; ++s) {
    if (ctrl_compute_actions_async(c)) return 1;
    if (step_actions_dev(env, c->d_actions)) return 1;
    if (with_objective) {  // on `aux`, beside the observation below and the next step's controller
      double *dst = objective_dev_out ? objective_dev_out + (size_t)s * env->E : env->d_objective;
      if (objective_aux(env, dst)) return 1;
    }
    if (obs_dev && ccb_env_cnn_input_dev(env, C, use_comm, obs_dev + (size_t)s * obs_step_stride)) return 1;
  }
  return join_aux(env);  // everything this call produced is ordered on the env stream when it returns
}

extern "C" int ccb_ctrl_rollout_obs(ccb_ctrl *c, int steps, int with_objective, double *objective_dev_out, int C,
                                    int use_comm, float *obs_dev, size_t obs_step_stride) {
  CCB_CUDA(cudaSetDevice(c->env->device));
  ccb_env *env = c->env;
  ccb_ctrl::RolloutGraph &g = c->rg;
  const int entry_valid = (env->cells_version == env->pos_version ? 1 : 0) | (env->wsums_version == env->pos_version ? 2 : 0);
  const bool same = g.steps == steps && g.with_objective == with_objective && g.C == C && g.use_comm == use_comm &&
                    g.obj == objective_dev_out && g.obs == obs_dev && g.stride == obs_step_stride &&
                    g.entry_valid == entry_valid;
  if (env->no_graph || env->no_aux || env->band_mode || steps < 1 || steps > 64)
    return rollout_obs_eager(c, steps, with_objective, objective_dev_out, C, use_comm, obs_dev, obs_step_stride);
  if (!same) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g = ccb_ctrl::RolloutGraph{};
    g.steps = steps, g.with_objective = with_objective, g.C = C, g.use_comm = use_comm, g.obj = objective_dev_out;
    g.obs = obs_dev, g.stride = obs_step_stride, g.entry_valid = entry_valid;
  }
  if (join_aux(env)) return 1;
  if (!env->band_mode && use_site_ordered(env, env->d.N) && ensure_prefix(env)) return 1;  // never inside the captured region: world maps change between replays
  auto launch_graph = [&]() -> int {
    CCB_CUDA(cudaEventRecord(env->ev_fork, env->stream));
    CCB_CUDA(cudaStreamWaitEvent(env->cap, env->ev_fork, 0));
    CCB_CUDA(cudaGraphLaunch(g.exec, env->cap));
    CCB_CUDA(cudaEventRecord(env->ev_join, env->cap));
    CCB_CUDA(cudaStreamWaitEvent(env->stream, env->ev_join, 0));
    // the events were last recorded inside the captured region: make them real again
    CCB_CUDA(cudaEventRecord(env->ev_state, env->stream));
    CCB_CUDA(cudaEventRecord(env->ev_cells, env->stream));
    CCB_CUDA(cudaEventRecord(env->ev_vg, env->stream));
    env->aux_pending = false;
    return 0;
  };
  if (g.exec) {  // replay
    if (launch_graph()) return 1;
    g_launches.fetch_add(g.launches);
    env->pos_version += g.pos_delta;
    if (g.cells_recomputed) env->cells_version = env->pos_version - g.cells_lag, env->cells_have_polys = false;
    if (g.wsums_valid_after) env->wsums_version = env->pos_version;
    return 0;
  }
  if (g.seen < 1) {  // first call: eager (also allocates and caches whatever the kernels need)
    g.seen++;
    return rollout_obs_eager(c, steps, with_objective, objective_dev_out, C, use_comm, obs_dev, obs_step_stride);
  }
  // second identical call: capture it
  const uint64_t pos0 = env->pos_version, launches0 = g_launches.load();
  env->cells_recomputed = false;
  cudaGraph_t graph = nullptr;
  int rc;
  cudaError_t ce;
  {
    cudaStream_t user = env->stream;
    env->stream = env->cap;  // every launch of the captured pass goes to `cap` (and, forked from it, `aux`)
    ce = cudaStreamBeginCapture(env->cap, cudaStreamCaptureModeRelaxed);
    if (ce != cudaSuccess) {
      env->stream = user;
      env->no_graph = true;
      cudaGetLastError();
      return rollout_obs_eager(c, steps, with_objective, objective_dev_out, C, use_comm, obs_dev, obs_step_stride);
    }
    rc = rollout_obs_eager(c, steps, with_objective, objective_dev_out, C, use_comm, obs_dev, obs_step_stride);
    ce = cudaStreamEndCapture(env->cap, &graph);
    env->stream = user;
  }
  if (rc || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    env->no_graph = true;  // do not try again; the state may be inconsistent with the (unexecuted) capture
    return fail("roll-out graph capture failed", rc ? ccb_last_error() : cudaGetErrorString(ce));
  }
  g.launches = g_launches.load() - launches0;
  g.pos_delta = env->pos_version - pos0;
  g.cells_recomputed = env->cells_recomputed;
  g.cells_lag = env->pos_version - env->cells_version;
  g.wsums_valid_after = env->wsums_version == env->pos_version;
  cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    g.exec = nullptr;
    env->no_graph = true;
    return fail("cudaGraphInstantiate failed", cudaGetErrorString(ie));
  }
  // nothing ran during the capture: launch the graph for this call (host bookkeeping is already done)
  return launch_graph();
}

extern "C" int ccb_ctrl_rollout(ccb_ctrl *c, int steps, int with_objective, double *objective_dev_out) {
  return ccb_ctrl_rollout_obs(c, steps, with_objective, objective_dev_out, 0, 0, nullptr, 0);
}

// =================================================================================================
// Profiling hook: launch one kernel of the step `iters` times back to back on the env stream,
// bypassing the position-version cache.  bench.py brackets it with CUDA events for the per-kernel
// roofline; it is not part of the reference-facing API.
// =================================================================================================
extern "C" int ccb_bench_kernel(ccb_env *env, int which, int iters, float *scratch_dev) {
  if (!env->initialised) return fail("environment has no robot positions yet");
  CCB_CUDA(cudaSetDevice(env->device));
  const EnvDev &d = env->d;
  for (int i = 0; i < iters; ++i) {
    switch (which) {
      case 0: if (launch_voronoi_global(env, SRC_WORLD, env->d_sums)) return 1; env->cells_version = 0; break;
      case 1: if (launch_voronoi_global(env, SRC_EXPLORED, env->d_sums)) return 1; env->cells_version = 0; break;
      case 2: env->wsums_version = 0; if (launch_voronoi_single(env, env->d_sums)) return 1; env->cells_version = 0; break;
      case 3: if (launch_mark(env, -1, -1, 1, 1)) return 1; break;
      case 4: env->wsums_version = env->pos_version; if (ccb_env_cnn_input_dev(env, 32, 1, scratch_dev)) return 1; break;
      case 8: {  // the fused observation + window-Voronoi variant
        env->wsums_version = 0;
        env->n_decentralized++;
        const int rc = ccb_env_cnn_input_dev(env, 32, 1, scratch_dev);
        env->n_decentralized--;
        if (rc) return 1;
        break;
      }
      case 5: if (launch_finalize(env, env->d_sums, env->cells, 0)) return 1; env->cells_version = 0; break;
      case 6: if (launch_voronoi_global(env, SRC_WORLD, env->d_sums, nullptr, 0, -1, 2)) return 1; env->cells_version = 0; break;
      case 7: if (launch_voronoi_global(env, SRC_WORLD, env->d_sums, nullptr, 0, -1, 1)) return 1; env->cells_version = 0; break;
      case 9: if (launch_voronoi_global(env, SRC_WORLD, env->d_sums, nullptr, 0, -1, 4)) return 1; env->cells_version = 0; break;
      default: return fail("ccb_bench_kernel: unknown kernel id");
    }
  }
  (void)d;
  return 0;
}
