// Thin extern "C" wrapper around the REFERENCE's own generate_world_map_cuda()
// (cppsrc/core/src/cuda/generate_world_map.cu:175-280), compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libref_idf.so.
// TEST INFRASTRUCTURE ONLY: it is the GPU IDF oracle and the "kernel to beat".
#include <vector>

#include "CoverageControl/generate_world_map.h"

extern "C" int ref_idf_generate(const float *bnd6, int num_dists,
                                const float *px, const float *py, const int *sz,
                                const float *imp, int num_poly, int map_size,
                                float resolution, float truncation, float norm,
                                float *host_out, float *factor_out) {
  using namespace CoverageControl;
  std::vector<BND_Cuda> dists(num_dists > 0 ? num_dists : 1);
  for (int i = 0; i < num_dists; ++i) {
    dists[i].mean_x = bnd6[6 * i + 0];
    dists[i].mean_y = bnd6[6 * i + 1];
    dists[i].sigma_x = bnd6[6 * i + 2];
    dists[i].sigma_y = bnd6[6 * i + 3];
    dists[i].scale = bnd6[6 * i + 4];
    dists[i].rho = bnd6[6 * i + 5];
  }
  Polygons_Cuda_Host polys;
  int start = 0;
  for (int k = 0; k < num_poly; ++k) {
    Bounds b;  // same bounds construction as world_idf.cpp:125-147
    for (int v = 0; v < sz[k]; ++v) {
      float x = px[start + v], y = py[start + v];
      polys.x.push_back(x);
      polys.y.push_back(y);
      if (x < b.xmin) b.xmin = x;
      if (y < b.ymin) b.ymin = y;
      if (x > b.xmax) b.xmax = x;
      if (y > b.ymax) b.ymax = y;
    }
    polys.imp.push_back(imp[k]);
    polys.sz.push_back(sz[k]);
    polys.bounds.push_back(b);
    start += sz[k];
  }
  polys.num_pts = static_cast<int>(polys.x.size());
  polys.num_polygons = static_cast<int>(polys.imp.size());
  float f = 0;
  try {
    generate_world_map_cuda(dists.data(), polys, num_dists, map_size, resolution,
                            truncation, norm, host_out, f);
  } catch (...) {
    return 1;
  }
  *factor_out = f;
  return 0;
}
