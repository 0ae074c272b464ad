/*
 * ccb200_test.h -- test, diagnostics and profiling hooks of libccb200.so.
 *
 * NOT part of the drop-in boundary (include/ccb200.h): nothing a binding of the reference's API needs is
 * declared here.  tests/, bench.py and scripts/ use these to (a) run the kernels' host-compilable pieces
 * without a GPU, (b) time one kernel of the step alone, (c) force a kernel choice the library would not make
 * on its own for the given shape.
 */
#ifndef CCB200_TEST_H_
#define CCB200_TEST_H_

#include "ccb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Force a kernel choice / diagnostics switch on one environment batch (tests and A/B timing only; the library
 * itself never reads the process environment).  Returns 0, or 1 for an unknown name.  Names:
 *   "vg_generic"  1: tile-ordered whole-map Voronoi (csrc/voronoi.cuh), 0: site-ordered (csrc/vglobal.cuh), -1: by shape
 *   "vg_prefix"   1/0: block prefix sums for the world-map moments (csrc/vglobal.cuh) on / off (off: k_vg_reduce streams the map)
 *   "vg_bulk"     1: (prefix off) k_vg_reduce_bulk -- cp.async.bulk + mbarrier ring -- instead of k_vg_reduce (A/B)
 *   "vg_fused"    1/0: with prefix sums, moments inside the row partition kernel / interval table + k_vg_lookup
 *   "vw_generic"  1: generic window Voronoi kernel      "cnn_generic" 1: generic observation kernel
 *   "obs_sorted"  0: launch the observation kernel's CTAs in robot order instead of grouped by window misalignment
 *   "obs_v1"      1: the row-ordered fused observation kernel of round 1 instead of the item-ordered one
 *   "vor_rpw"     rows per warp region of the tile-ordered kernel   "vg_warps" / "vg_griddiv" / "vg_part_rpt"
 *   "no_aux" / "no_graph" / "no_wfuse"   1: single stream / no CUDA-graph replay / no fused window moments */
int ccb_test_set_option(ccb_env *env, const char *name, int value);

/* Test hook, host only: the triangles the library cuts a polygon feature into before the IDF kernel's
 * point-in-piece test (the reference partitions with CGAL, world_idf.cpp:125-153; polygon_utils.cpp:38-68);
 * xy = n x 2, tris_out = cap x 6 doubles (x0 y0 x1 y1 x2 y2); returns the number of triangles or -1. */
int ccb_selftest_triangulate(const double *xy, int n, double *tris_out, int cap);
/* Test hook, host only: the frontier queue of ccb_env_exploration_features (the kernel's own heap code
 * compiled for the host) over values[n]; candidate numbers as popped, worst first; returns their count
 * or -1.  Touches no map and no device. */
int ccb_selftest_frontier_select(const double *values, int n, int num_frontiers, int *order_out);
/* Test hook, host only: the separable antialiased bilinear down-sampling taps the observation kernels apply
 * (what torchvision.transforms.functional.resize(BILINEAR, antialias=True) does, coverage_env_utils.py:97-141):
 * output index o reads inputs [xmin[o], xmin[o] + xsize[o]) with weights w[o][0 .. kmax).  Returns kmax or -1. */
int ccb_selftest_resize_weights(int L, int C, int *xmin_out, int *xsize_out, float *w_out, int w_cap);
/* Test hook, host only: the window of size `sub` around a position as every window kernel derives it
 * (map_utils.h:51-87 ComputeOffsets): out6 = unclipped origin (ox, oy), clipped rows [x0, x1), columns [y0, y1). */
int ccb_selftest_make_window(double res, double px, double py, int sub, int map_size, int *out6);
/* Test hook, host only: k_vg_partition (the whole-map Voronoi's row partition, replacing voronoi.cpp:55-165
 * for the global cells) for ONE environment, its row sweep being the kernel's own code compiled for the
 * host.  pos[R][2] sites, rows [row0, row1) of an N x N map; table[R][N] = lo | hi << 16 (columns [lo, hi)
 * of the row belong to the site), flag_rows[<= N] = rows left to the per-cell rule, span[R][2] = first row /
 * last row + 1 on which the site owns a cell; band_rows = rows per CTA band (0: the kernel's default, 128).  Returns the number of flagged rows or -1.  Touches no map
 * and no device. */
int ccb_selftest_vg_partition(const double *pos, int R, int N, double res, int row0, int row1, int band_rows,
                              uint32_t *table, int *flag_rows, int *span);
/* Test hook, host only: the own-cell column interval of window rows [row0, row1) as the DecentralizedCVT
 * window kernels compute it (replacing voronoi.cpp:55-150 for decentralized_cvt.h:85-143), their own code
 * compiled for the host.  px / py[K + 1]: the robot, then its K neighbours; window columns y0 .. y0 + ncols - 1.
 * lo / hi / exact[row1 - row0]; exact = the kernel applies the per-cell rule inside [lo, hi) on that row.
 * words_out (optional, 9 words): the first row's interval as mask words from column word wbase.
 * Touches no map and no device. */
int ccb_selftest_vw_row_interval(double res, int row0, int row1, int y0, int ncols, int K, const double *px,
                                 const double *py, int *lo, int *hi, int *exact, int wbase, uint32_t *words_out);
/* diagnostics: how many map rows the last whole-map Voronoi pass resolved with the exact per-cell rule
 * (csrc/vglobal.cuh k_vg_fixup) and which (entries e * N + row, at most cap) */
int ccb_env_flagged_rows(ccb_env *env, int *count, int *rows, int cap);
/* launch one kernel of the step `iters` times back to back on the env stream, bypassing caches:
 * which = 0 Voronoi label+reduce (world map), 1 same on the explored-IDF map, 2 per-robot window
 * variant, 3 sensor-window marking, 4 CNN input (C = 32, scratch_dev = E x R x 4 x 32 x 32 floats),
 * 5 cell finalisation, 6 the per-site reduce of the whole-map Voronoi alone (on the row partition a
 * previous which = 0 left), 7 its row partition alone, 8 the CNN input fused with the window Voronoi */
int ccb_bench_kernel(ccb_env *env, int which, int iters, float *scratch_dev);
/* the IDF integration kernel alone, `iters` times over the whole map (inputs uploaded before the timed region,
 * no maximum / normalise pass; leaves the map un-normalised); *ms_out = mean kernel time by CUDA events */
int ccb_bench_idf_kernel(ccb_idf *idf, int mode, int iters, float *ms_out);
/* the FP32 FMA rate the GPU sustains (TFLOP/s, best of 5): denominator of the IDF kernel's FP32 roofline */
int ccb_bench_fp32_peak(int device, double *tflops_out);
/* roofline arithmetic: what = 0 -> cells of all robots' local windows that their robot maps have seen (what
 * the observation / window-Voronoi kernels must read); 1 -> (site, row) pairs with cells in the last
 * site-ordered row partition (what k_vg_lookup / k_vg_reduce walk) */
int ccb_test_count(ccb_env *env, int what, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* CCB200_TEST_H_ */
