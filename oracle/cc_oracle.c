/*
 * cc_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see cc_oracle.h).
 *
 * Plain-C restatement of the reference's grid hot path.  Paths cited below are
 * relative to the reference tree (KumarRobotics/CoverageControl).
 * Compile with -O2 -ffp-contract=off (see oracle/Makefile): the reference's
 * x86-64 Release build has no FMA contraction either.
 *
 * CGAL (Delaunay -> arrangement -> face polygon, voronoi.cpp:184-326) is
 * replaced by clipping the map rectangle with the n-1 bisector half-planes;
 * the resulting convex polygon has the same vertex set.  Everything downstream
 * of the polygon (scan-line rasterisation, fp64 sums, vertex-mean fallback)
 * follows the reference statement by statement.
 */
#include "cc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define K_EPS 1e-10       /* constants.h:48 kEps */
#define K_LARGE_EPS 1e-4  /* constants.h:50 kLargeEps */

void orc_params_default(orc_params *p) {
  /* parameters.h:60-160 class defaults */
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

/* ========================================================================== */
/* IDF                                                                         */
/* ========================================================================== */

/* geometry_utils.cuh:44-54 */
static int orientation_f(float px, float py, float qx, float qy, float rx,
                         float ry) {
  float qpx = qx - px, qpy = qy - py;
  float rpx = rx - px, rpy = ry - py;
  float c1 = qpx * rpy;
  float c2 = qpy * rpx;
  if (c1 > c2) return 1;
  if (c1 < c2) return -1;
  return 0;
}

/* geometry_utils.cuh:60-101 */
int orc_point_in_monotone_polygon(const float *x, const float *y, int sz,
                                  float rx, float ry) {
  int left = 0, right = 0;
  for (int ct = 0; ct < sz; ++ct) {
    int ctp1 = ct + 1;
    if (ctp1 == sz) ctp1 = 0;
    float yct = y[ct], yctp1 = y[ctp1];
    float px, py, qx, qy;
    if (yct < yctp1) {
      if (ry < yct || ry > yctp1) continue;
      px = x[ct], py = yct, qx = x[ctp1], qy = yctp1;
    } else {
      if (ry < yctp1 || ry > yct) continue;
      qx = x[ct], qy = yct, px = x[ctp1], py = yctp1;
    }
    int o = orientation_f(px, py, qx, qy, rx, ry);
    if (o == 0) return 1;
    if (o == 1) left = 1;
    if (o == -1) right = 1;
    if (left & right) return 1;
  }
  return 0;
}

/* strict-interior crossing-number test (stands in for CGAL bounded_side ==
 * ON_BOUNDED_SIDE, world_idf.cpp:66-70) */
static int point_strictly_inside(const double *px, const double *py, int n,
                                 double x, double y) {
  int inside = 0;
  for (int i = 0, j = n - 1; i < n; j = i++) {
    double xi = px[i], yi = py[i], xj = px[j], yj = py[j];
    /* on-boundary -> not strictly inside */
    double cross = (xj - xi) * (y - yi) - (yj - yi) * (x - xi);
    if (cross == 0 && x >= fmin(xi, xj) && x <= fmax(xi, xj) &&
        y >= fmin(yi, yj) && y <= fmax(yi, yj))
      return 0;
    if ((yi > y) != (yj > y)) {
      double xc = xi + (y - yi) * (xj - xi) / (yj - yi);
      if (x < xc) inside = !inside;
    }
  }
  return inside;
}

float orc_idf_generate(const orc_params *p, const double *bnd, int F,
                       const double *px, const double *py, const int *sz,
                       const float *imp, int num_poly, int mode, float *out) {
  const int N = p->pWorldMapSize;
  const float res = (float)p->pResolution;
  float *fx = NULL, *fy = NULL;
  int npts = 0;
  for (int k = 0; k < num_poly; ++k) npts += sz[k];
  if (npts > 0) {
    fx = (float *)malloc(sizeof(float) * npts);
    fy = (float *)malloc(sizeof(float) * npts);
    for (int i = 0; i < npts; ++i) fx[i] = (float)px[i], fy[i] = (float)py[i];
  }
  float max_importance = 0;
  int max_init = 0;
  if (mode == 2) {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)N * N);
    orc_idf_fp64(p, bnd, F, 0.0, tmp, NULL);
    for (size_t i = 0; i < (size_t)N * N; ++i) out[i] = (float)tmp[i];
    free(tmp);
  }
  const float trun_f = (float)p->pTruncationBND;
  const float trun_sq = trun_f * trun_f;
  const float one_by_sqrt2_cpu = 1.f / sqrtf(2.f);              /* constants.h:53 */
  const float one_by_sqrt2_cuda = (float)(1. / sqrt(2.));       /* generate_world_map.cu:192 */
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) {
      float importance = 0;
      float midx, midy;
      if (mode == 0) {
        /* world_idf.cpp:54-63, world_idf.h:73-95 */
        float x1 = res * i, x2 = x1 + res;
        float y1 = res * j, y2 = y1 + res;
        midx = (x1 + x2) / 2.f;
        midy = (y1 + y2) / 2.f;
        for (int f = 0; f < F; ++f) {
          const double *b = bnd + 6 * f;
          float mx = (float)b[0], my = (float)b[1], sx = (float)b[2];
          float sc = (float)b[4];
          /* bivariate_normal_distribution.h:152-154: circular branch divides
           * both axes by sigma.x (rho_ is always 0, :103-112) */
          float tx = (midx - mx) / sx, ty = (midy - my) / sx;
          if (tx * tx + ty * ty > trun_sq) continue;
          float ax1 = (x1 - mx) / sx, ax2 = (x2 - mx) / sx;
          float ay1 = (y1 - my) / sx, ay2 = (y2 - my) / sx;
          /* bivariate_normal_distribution.h:176-181; order world_idf.h:89-92 */
          importance += sc * erfcf(ax1 * one_by_sqrt2_cpu) * erfcf(ay1 * one_by_sqrt2_cpu) / 4.f;
          importance -= sc * erfcf(ax2 * one_by_sqrt2_cpu) * erfcf(ay1 * one_by_sqrt2_cpu) / 4.f;
          importance -= sc * erfcf(ax1 * one_by_sqrt2_cpu) * erfcf(ay2 * one_by_sqrt2_cpu) / 4.f;
          importance += sc * erfcf(ax2 * one_by_sqrt2_cpu) * erfcf(ay2 * one_by_sqrt2_cpu) / 4.f;
        }
      } else {
        /* generate_world_map.cu:146-164 */
        float x1 = i * res, y1 = j * res;
        float x2 = i * res + res, y2 = j * res + res;
        midx = (float)((x1 + x2) / 2.);
        midy = (float)((y1 + y2) / 2.);
        if (mode == 1) {
          for (int f = 0; f < F; ++f) {
            const double *b = bnd + 6 * f;
            float mx = (float)b[0], my = (float)b[1];
            float sx = (float)b[2], sy = (float)b[3];
            float sc = (float)b[4];
            /* rho == 0 branch, generate_world_map.cu:102-115 */
            float tx = (midx - mx) / sx, ty = (midy - my) / sy;
            if (tx * tx + ty * ty > trun_sq) continue;
            float ax1 = (x1 - mx) / sx, ax2 = (x2 - mx) / sx;
            float ay1 = (y1 - my) / sy, ay2 = (y2 - my) / sy;
            /* generate_world_map.cu:76-93,116-119 */
            importance += sc * erfcf(ax1 * one_by_sqrt2_cuda) * erfcf(ay1 * one_by_sqrt2_cuda) / 4.f;
            importance -= sc * erfcf(ax2 * one_by_sqrt2_cuda) * erfcf(ay1 * one_by_sqrt2_cuda) / 4.f;
            importance -= sc * erfcf(ax1 * one_by_sqrt2_cuda) * erfcf(ay2 * one_by_sqrt2_cuda) / 4.f;
            importance += sc * erfcf(ax2 * one_by_sqrt2_cuda) * erfcf(ay2 * one_by_sqrt2_cuda) / 4.f;
          }
        } else {
          importance = out[(size_t)i * N + j];
        }
      }
      float poly_imp = 0;
      int start = 0;
      for (int k = 0; k < num_poly; ++k) {
        int n = sz[k];
        if (mode == 0) {
          if (point_strictly_inside(px + start, py + start, n, midx, midy))
            poly_imp = poly_imp > imp[k] ? poly_imp : imp[k];
        } else {
          /* generate_world_map.cu:124-144 incl. the bounding-box pre-test */
          float xmin = 3.4028235e38f, ymin = 3.4028235e38f;
          float xmax = 1.17549435e-38f, ymax = 1.17549435e-38f; /* generate_world_map.h:49-54 */
          for (int v = 0; v < n; ++v) {
            float vx = fx[start + v], vy = fy[start + v];
            if (vx < xmin) xmin = vx;
            if (vy < ymin) ymin = vy;
            if (vx > xmax) xmax = vx;
            if (vy > ymax) ymax = vy;
          }
          if (!(midx < xmin || midx > xmax || midy < ymin || midy > ymax)) {
            if (orc_point_in_monotone_polygon(fx + start, fy + start, n, midx, midy))
              poly_imp = fmaxf(poly_imp, imp[k]);
          }
        }
        start += n;
      }
      importance += poly_imp;
      out[(size_t)i * N + j] = importance;
      if (mode == 0) {
        if (importance > max_importance) max_importance = importance; /* world_idf.cpp:78-80 */
      } else {
        /* thrust::max_element, generate_world_map.cu:248-250 (true max) */
        if (!max_init || importance > max_importance) max_importance = importance, max_init = 1;
      }
    }
  }
  free(fx);
  free(fy);
  float factor;
  if (max_importance < K_EPS)
    factor = (float)p->pNorm;
  else
    factor = (float)p->pNorm / max_importance;
  if (!(factor > 1e-5)) return factor; /* world_idf.cpp:90, generate_world_map.cu:257 */
  for (size_t i = 0; i < (size_t)N * N; ++i) out[i] *= factor;
  return factor;
}

void orc_idf_fp64(const orc_params *p, const double *bnd, int F, double band,
                  double *out, unsigned char *borderline) {
  const int N = p->pWorldMapSize;
  const double res = p->pResolution;
  const double T = p->pTruncationBND, T2 = T * T;
  const double rs2 = 1.0 / sqrt(2.0);
  memset(out, 0, sizeof(double) * (size_t)N * N);
  if (borderline) memset(borderline, 0, (size_t)N * N);
  double *ex = (double *)malloc(sizeof(double) * (N + 1));
  double *ey = (double *)malloc(sizeof(double) * (N + 1));
  for (int f = 0; f < F; ++f) {
    const double *b = bnd + 6 * f;
    /* the product stores fp32 feature records (generate_world_map.h:42-46) */
    double mx = (float)b[0], my = (float)b[1], sx = (float)b[2], sy = (float)b[3];
    double sc = (float)b[4];
    for (int i = 0; i <= N; ++i) {
      ex[i] = erfc((i * res - mx) / sx * rs2);
      ey[i] = erfc((i * res - my) / sy * rs2);
    }
    int i0 = (int)floor((mx - (T + 1e-3) * sx) / res) - 1, i1 = (int)ceil((mx + (T + 1e-3) * sx) / res) + 1;
    int j0 = (int)floor((my - (T + 1e-3) * sy) / res) - 1, j1 = (int)ceil((my + (T + 1e-3) * sy) / res) + 1;
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (i1 > N) i1 = N;
    if (j1 > N) j1 = N;
    for (int i = i0; i < i1; ++i) {
      double tx = ((i + 0.5) * res - mx) / sx;
      for (int j = j0; j < j1; ++j) {
        double ty = ((j + 0.5) * res - my) / sy;
        double r2 = tx * tx + ty * ty;
        if (borderline && fabs(r2 - T2) < band) borderline[(size_t)i * N + j] += 1;
        if (r2 > T2) continue;
        out[(size_t)i * N + j] += sc * 0.25 * (ex[i] - ex[i + 1]) * (ey[j] - ey[j + 1]);
      }
    }
  }
  free(ex);
  free(ey);
}

/* ========================================================================== */
/* Windows                                                                     */
/* ========================================================================== */

/* map_utils.h:51-87 */
void orc_compute_offsets(double res, double px, double py, int submap,
                         int map_size, orc_bounds *index, orc_bounds *offset) {
  int pos_idx = (int)round(px / res); /* std::round: half away from zero */
  int pos_idy = (int)round(py / res);
  index->left = pos_idx - submap / 2;
  index->right = pos_idx + submap / 2;
  index->bottom = pos_idy - submap / 2;
  index->top = pos_idy + submap / 2;
  index->width = index->height = 0;
  offset->left = offset->right = offset->bottom = offset->top = 0;
  if (index->left < 0) offset->left = -index->left;
  if (index->bottom < 0) offset->bottom = -index->bottom;
  if (index->right > map_size) offset->right = map_size - index->right;
  if (index->top > map_size) offset->top = map_size - index->top;
  offset->width = index->right + offset->right - (index->left + offset->left);
  offset->height = index->top + offset->top - (index->bottom + offset->bottom);
}

/* map_utils.h:140-155 */
static int is_point_outside_boundary(double res, double px, double py,
                                     int sensor, int boundary) {
  if (px <= -sensor * res / 2.) return 1;
  if (py <= -sensor * res / 2.) return 1;
  if (px >= boundary * res + sensor * res / 2.) return 1;
  if (py >= boundary * res + sensor * res / 2.) return 1;
  return 0;
}

/* ========================================================================== */
/* Voronoi                                                                     */
/* ========================================================================== */

#define ORC_MAX_VERTS 64

int orc_clip_cell(const double *sites, int n, int k, double w, double h,
                  double *verts, int max_verts) {
  /* rectangle in CCW order, voronoi.cpp:171-173 */
  int cap = 4 + n + 8;
  double *a = (double *)malloc(sizeof(double) * 2 * cap);
  double *b = (double *)malloc(sizeof(double) * 2 * cap);
  int na = 4;
  a[0] = 0, a[1] = 0, a[2] = w, a[3] = 0, a[4] = w, a[5] = h, a[6] = 0, a[7] = h;
  double kx = sites[2 * k], ky = sites[2 * k + 1];
  for (int j = 0; j < n && na > 0; ++j) {
    if (j == k) continue;
    double jx = sites[2 * j], jy = sites[2 * j + 1];
    double nx = jx - kx, ny = jy - ky;
    if (nx == 0 && ny == 0) continue; /* coincident sites: no bisector */
    /* keep { x : n.x <= c },  c = (|pj|^2 - |pk|^2)/2 */
    double c = ((jx * jx + jy * jy) - (kx * kx + ky * ky)) / 2.;
    int nb = 0;
    for (int v = 0; v < na; ++v) {
      int u = (v + 1) % na;
      double vx = a[2 * v], vy = a[2 * v + 1];
      double ux = a[2 * u], uy = a[2 * u + 1];
      double dv = nx * vx + ny * vy - c;
      double du = nx * ux + ny * uy - c;
      int vin = dv <= 0, uin = du <= 0;
      if (vin) {
        b[2 * nb] = vx, b[2 * nb + 1] = vy, ++nb;
      }
      if (vin != uin) {
        double t = dv / (dv - du);
        b[2 * nb] = vx + t * (ux - vx), b[2 * nb + 1] = vy + t * (uy - vy), ++nb;
      }
    }
    /* drop exact consecutive duplicates (plane through a vertex) */
    int nc = 0;
    for (int v = 0; v < nb; ++v) {
      int pv = (v + nb - 1) % nb;
      if (nb > 1 && b[2 * v] == b[2 * pv] && b[2 * v + 1] == b[2 * pv + 1]) continue;
      a[2 * nc] = b[2 * v], a[2 * nc + 1] = b[2 * v + 1], ++nc;
    }
    na = nc;
  }
  int ret = na;
  if (na > max_verts) ret = -1;
  else memcpy(verts, a, sizeof(double) * 2 * na);
  free(a);
  free(b);
  return ret;
}

typedef struct {
  double mass, cx, cy, sd, sd2;
} acc_t;

/* voronoi.cpp:55-150 (CellNavigator) with voronoi.h:89-95 (MassCentroidFunctional)
 * as the evaluator.  Returns 0, or 1 for the reference's "Unexpected error!". */
static int cell_navigate(const double *cell, int n, const float *map, int stride,
                         double map_w, double map_h, double res, double shiftx,
                         double shifty, double sitex, double sitey, acc_t *acc) {
  int left_id = 0, right_id = 0;
  for (int i = 1; i < n; ++i) {
    if (cell[2 * i] < cell[2 * left_id]) left_id = i;
    if (cell[2 * i] > cell[2 * right_id]) right_id = i;
  }
  int min_i = (int)round(cell[2 * left_id] / res);
  min_i = min_i < 0 ? 0 : min_i;
  int max_i = (int)round(cell[2 * right_id] / res);
  max_i = max_i > map_w ? (int)map_w : max_i;
#define NEXT_ID(id) (((id) + 1) % n)
#define PREV_ID(id) ((id) == 0 ? (n - 1) : ((id)-1))
  int cc = NEXT_ID(left_id);
  int c = PREV_ID(left_id);
  for (int i = min_i; i < max_i; ++i) {
    double x = i * res + res / 2.;
    while (1) {
      if (cc == left_id) break;
      if (cell[2 * cc] < x) cc = NEXT_ID(cc);
      else break;
    }
    while (1) {
      if (c == left_id) break;
      if (cell[2 * c] < x) c = PREV_ID(c);
      else break;
    }
    if (cell[2 * cc] < x) break;
    if (cell[2 * c] < x) break;
    int pc = PREV_ID(cc);
    double x1 = cell[2 * pc], y1 = cell[2 * pc + 1];
    double x2 = cell[2 * cc], y2 = cell[2 * cc + 1];
    if ((x2 - x1) < K_EPS) return 1;
    double y_lower = y1 + (x - x1) * (y2 - y1) / (x2 - x1);
    int nx = NEXT_ID(c);
    x1 = cell[2 * nx], y1 = cell[2 * nx + 1];
    x2 = cell[2 * c], y2 = cell[2 * c + 1];
    if ((x2 - x1) < K_EPS) return 1;
    double y_upper = y1 + (x - x1) * (y2 - y1) / (x2 - x1);
    int min_j = (int)round(y_lower / res);
    min_j = min_j < 0 ? 0 : min_j;
    int max_j = (int)round(y_upper / res);
    max_j = max_j > map_h ? (int)map_h : max_j;
    for (int j = min_j; j < max_j; ++j) {
      double y = j * res + res / 2.;
      double v = map[(size_t)i * stride + j];
      double ptx = x + shiftx, pty = y + shifty;
      acc->mass += v;
      acc->cx += ptx * v;
      acc->cy += pty * v;
      double dx = ptx - sitex, dy = pty - sitey;
      double d2 = dx * dx + dy * dy;
      acc->sd += sqrt(d2) * v;
      acc->sd2 += d2 * v;
    }
  }
#undef NEXT_ID
#undef PREV_ID
  return 0;
}

/* voronoi.cpp:152-165 + :41-53 */
static int compute_mass_centroid(orc_cell *vc, const float *map, int stride,
                                 double map_w, double map_h, double res,
                                 double shiftx, double shifty) {
  acc_t acc = {0, 0, 0, 0, 0};
  int err = cell_navigate(vc->verts, vc->nverts, map, stride, map_w, map_h, res,
                          shiftx, shifty, vc->site[0], vc->site[1], &acc);
  if (err) return err;
  vc->mass = acc.mass;
  vc->sum_idf_site_dist = acc.sd;
  vc->sum_idf_site_dist_sqr = acc.sd2;
  if (acc.mass < K_EPS) {
    double sx = 0, sy = 0;
    for (int v = 0; v < vc->nverts; ++v) sx += vc->verts[2 * v], sy += vc->verts[2 * v + 1];
    vc->centroid[0] = sx / vc->nverts + shiftx;
    vc->centroid[1] = sy / vc->nverts + shifty;
  } else {
    vc->centroid[0] = acc.cx / acc.mass;
    vc->centroid[1] = acc.cy / acc.mass;
  }
  return 0;
}

int orc_voronoi(const double *sites, int n, const float *map, int stride,
                double map_w, double map_h, double res, int single,
                int robot_id, orc_cell *cells) {
  double shiftx = 0, shifty = 0;
  if (single) shiftx = -sites[0], shifty = -sites[1]; /* voronoi.h:171 */
  if (single) {
    orc_cell *vc = &cells[0];
    vc->site[0] = sites[2 * robot_id], vc->site[1] = sites[2 * robot_id + 1];
    vc->nverts = orc_clip_cell(sites, n, robot_id, map_w, map_h, vc->verts, ORC_MAX_VERTS);
    if (vc->nverts < 3) return 2;
    return compute_mass_centroid(vc, map, stride, map_w, map_h, res, shiftx, shifty);
  }
  int err = 0;
  /* voronoi.cpp:289 -- one thread per site, as the reference */
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < n; ++k) {
    orc_cell *vc = &cells[k];
    vc->site[0] = sites[2 * k], vc->site[1] = sites[2 * k + 1];
    vc->nverts = orc_clip_cell(sites, n, k, map_w, map_h, vc->verts, ORC_MAX_VERTS);
    int e = 0;
    if (vc->nverts < 3) e = 2;
    else e = compute_mass_centroid(vc, map, stride, map_w, map_h, res, 0, 0);
    if (e) {
#pragma omp atomic write
      err = e;
    }
  }
  return err;
}

void orc_voronoi_bruteforce(const double *sites, int n, const float *map,
                            int stride, int W, int H, double res, int *labels,
                            double *sums) {
  memset(sums, 0, sizeof(double) * 4 * n);
  for (int i = 0; i < W; ++i) {
    double cx = (i + 0.5) * res;
    for (int j = 0; j < H; ++j) {
      double cy = (j + 0.5) * res;
      int best = 0;
      double bd = 0;
      for (int k = 0; k < n; ++k) {
        double dx = cx - sites[2 * k], dy = cy - sites[2 * k + 1];
        double d2 = dx * dx + dy * dy;
        if (k == 0 || d2 < bd) bd = d2, best = k;
      }
      if (labels) labels[(size_t)i * H + j] = best;
      double v = map[(size_t)i * stride + j];
      sums[4 * best + 0] += v;
      sums[4 * best + 1] += cx * v;
      sums[4 * best + 2] += cy * v;
      sums[4 * best + 3] += bd * v;
    }
  }
}

/* Labels only, rows dealt to the OpenMP threads: the same expression and tie rule as orc_voronoi_bruteforce
 * (argmin of dx*dx + dy*dy in fp64 without contraction, strict <, ascending site index), for maps too large
 * for the single-threaded loop (BASELINE configs 4 and 5: 8192^2 x 256 sites, 16384^2 x 64). */
void orc_voronoi_labels_omp(const double *sites, int n, int W, int H, double res, int *labels) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < W; ++i) {
    double cx = (i + 0.5) * res;
    for (int j = 0; j < H; ++j) {
      double cy = (j + 0.5) * res;
      int best = 0;
      double bd = 0;
      for (int k = 0; k < n; ++k) {
        double dx = cx - sites[2 * k], dy = cy - sites[2 * k + 1];
        double d2 = dx * dx + dy * dy;
        if (k == 0 || d2 < bd) bd = d2, best = k;
      }
      labels[(size_t)i * H + j] = best;
    }
  }
}

/* ========================================================================== */
/* Environment: dense maps exactly as the reference keeps them                 */
/* ========================================================================== */

struct orc_env {
  orc_params p;
  int R, N, S, L, H;
  float *world;        /* N*N */
  float *robot_map;    /* R*N*N  robot_model.h:69 */
  float *robot_expl;   /* R*N*N  robot_model.h:76 (only if pUpdateExplorationMap) */
  float *sensor_view;  /* R*S*S  robot_model.h:71 */
  float *system_map, *exploration_map, *explored_idf; /* coverage_system.h:82-87 */
  double *gstart, *gcur, *lcur;                       /* R*2 each */
  double *hist;                                       /* R*H*2 (ring as list) */
  int *hist_len;
  double *gpos; /* robot_global_positions_ */
};

static void env_update_sensor_view(orc_env *e, int r) {
  /* robot_model.h:87-96 */
  const orc_params *p = &e->p;
  float *sv = e->sensor_view + (size_t)r * e->S * e->S;
  memset(sv, 0, sizeof(float) * e->S * e->S);
  double px = e->gcur[2 * r], py = e->gcur[2 * r + 1];
  if (is_point_outside_boundary(p->pResolution, px, py, e->S, e->N)) return;
  orc_bounds ix, of;
  orc_compute_offsets(p->pResolution, px, py, e->S, e->N, &ix, &of);
  for (int a = 0; a < of.width; ++a)
    memcpy(sv + (size_t)(of.left + a) * e->S + of.bottom,
           e->world + (size_t)(ix.left + of.left + a) * e->N + ix.bottom + of.bottom,
           sizeof(float) * of.height);
}

static void env_update_robot_map(orc_env *e, int r) {
  /* robot_model.h:98-107 */
  const orc_params *p = &e->p;
  orc_bounds ix, of;
  orc_compute_offsets(p->pResolution, e->gcur[2 * r], e->gcur[2 * r + 1], e->S, e->N, &ix, &of);
  float *rm = e->robot_map + (size_t)r * e->N * e->N;
  const float *sv = e->sensor_view + (size_t)r * e->S * e->S;
  for (int a = 0; a < of.width; ++a)
    memcpy(rm + (size_t)(ix.left + of.left + a) * e->N + ix.bottom + of.bottom,
           sv + (size_t)(of.left + a) * e->S + of.bottom, sizeof(float) * of.height);
}

static void env_update_robot_expl(orc_env *e, int r) {
  /* robot_model.h:109-122 */
  const orc_params *p = &e->p;
  double px = e->gcur[2 * r], py = e->gcur[2 * r + 1];
  if (is_point_outside_boundary(p->pResolution, px, py, e->S, e->N)) return;
  orc_bounds ix, of;
  orc_compute_offsets(p->pResolution, px, py, e->S, e->N, &ix, &of);
  float *em = e->robot_expl + (size_t)r * e->N * e->N;
  for (int a = 0; a < of.width; ++a)
    memset(em + (size_t)(ix.left + of.left + a) * e->N + ix.bottom + of.bottom, 0,
           sizeof(float) * of.height);
}

static void env_after_move(orc_env *e, int r) {
  if (e->p.pUpdateSensorView) env_update_sensor_view(e, r);
  if (e->p.pUpdateRobotMap) env_update_robot_map(e, r);
  if (e->p.pUpdateExplorationMap) env_update_robot_expl(e, r);
}

static void clamp_global(const orc_env *e, double *g) {
  /* robot_model.h:229-243 */
  if (g[0] <= 0) g[0] = 0 + K_LARGE_EPS;
  if (g[1] <= 0) g[1] = 0 + K_LARGE_EPS;
  double max_xy = e->N * e->p.pResolution;
  if (g[0] >= max_xy) g[0] = max_xy - K_LARGE_EPS;
  if (g[1] >= max_xy) g[1] = max_xy - K_LARGE_EPS;
}

/* RobotModel::SetRobotPosition, robot_model.h:228-256 */
static void env_set_local(orc_env *e, int r, double lx, double ly) {
  double g[2] = {lx + e->gstart[2 * r], ly + e->gstart[2 * r + 1]};
  clamp_global(e, g);
  e->lcur[2 * r] = g[0] - e->gstart[2 * r];
  e->lcur[2 * r + 1] = g[1] - e->gstart[2 * r + 1];
  e->gcur[2 * r] = g[0];
  e->gcur[2 * r + 1] = g[1];
  env_after_move(e, r);
}

/* RobotModel::SetGlobalRobotPosition, robot_model.h:258-286 */
static void env_set_global(orc_env *e, int r, double gx, double gy) {
  double g[2] = {gx, gy};
  clamp_global(e, g);
  e->lcur[2 * r] = g[0] - e->gstart[2 * r];
  e->lcur[2 * r + 1] = g[1] - e->gstart[2 * r + 1];
  e->gcur[2 * r] = g[0];
  e->gcur[2 * r + 1] = g[1];
  env_after_move(e, r);
}

static void env_post_step(orc_env *e) {
  /* coverage_system.cpp:257-272 */
  const orc_params *p = &e->p;
  memcpy(e->gpos, e->gcur, sizeof(double) * 2 * e->R);
  if (p->pUpdateSystemMap) {
    /* coverage_system.h:104-120 */
    for (int r = 0; r < e->R; ++r) {
      orc_bounds ix, of;
      orc_compute_offsets(p->pResolution, e->gpos[2 * r], e->gpos[2 * r + 1], e->S, e->N, &ix, &of);
      const float *sv = e->sensor_view + (size_t)r * e->S * e->S;
      for (int a = 0; a < of.width; ++a) {
        size_t dst = (size_t)(ix.left + of.left + a) * e->N + ix.bottom + of.bottom;
        memcpy(e->explored_idf + dst, sv + (size_t)(of.left + a) * e->S + of.bottom,
               sizeof(float) * of.height);
        memset(e->exploration_map + dst, 0, sizeof(float) * of.height);
      }
    }
    size_t nn = (size_t)e->N * e->N;
#pragma omp parallel for
    for (size_t i = 0; i < nn; ++i) e->system_map[i] = e->explored_idf[i] - e->exploration_map[i];
  }
  for (int r = 0; r < e->R; ++r) {
    int *len = &e->hist_len[r];
    double *h = e->hist + (size_t)r * e->H * 2;
    if (*len > 0 && *len == p->pRobotPosHistorySize) {
      memmove(h, h + 2, sizeof(double) * 2 * (*len - 1)); /* pop_front, no push */
      --*len;
    } else {
      h[2 * *len] = e->gpos[2 * r], h[2 * *len + 1] = e->gpos[2 * r + 1];
      ++*len;
    }
  }
}

orc_env *orc_env_create(const orc_params *p, const float *world_map,
                        const double *positions, int R) {
  orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
  e->p = *p;
  e->R = R;
  e->N = p->pWorldMapSize;
  e->S = p->pSensorSize;
  e->L = p->pLocalMapSize;
  e->H = p->pRobotPosHistorySize > 0 ? p->pRobotPosHistorySize : 1;
  size_t nn = (size_t)e->N * e->N;
  e->world = (float *)malloc(sizeof(float) * nn);
  memcpy(e->world, world_map, sizeof(float) * nn);
  e->robot_map = (float *)malloc(sizeof(float) * nn * R);
  e->robot_expl = p->pUpdateExplorationMap ? (float *)malloc(sizeof(float) * nn * R) : NULL;
  e->sensor_view = (float *)calloc((size_t)R * e->S * e->S, sizeof(float));
  e->system_map = (float *)calloc(nn, sizeof(float));
  e->exploration_map = (float *)malloc(sizeof(float) * nn);
  e->explored_idf = (float *)calloc(nn, sizeof(float));
  e->gstart = (double *)malloc(sizeof(double) * 2 * R);
  e->gcur = (double *)malloc(sizeof(double) * 2 * R);
  e->lcur = (double *)calloc(2 * R, sizeof(double));
  e->gpos = (double *)malloc(sizeof(double) * 2 * R);
  e->hist = (double *)calloc((size_t)R * e->H * 2, sizeof(double));
  e->hist_len = (int *)calloc(R, sizeof(int));
  /* RobotModel::Initialize, robot_model.h:124-153 */
  float fill = p->pRobotMapUseUnknownImportance ? (float)(p->pUnknownImportance * p->pNorm) : 0.f;
  for (int r = 0; r < R; ++r) {
    e->gstart[2 * r] = e->gcur[2 * r] = positions[2 * r];
    e->gstart[2 * r + 1] = e->gcur[2 * r + 1] = positions[2 * r + 1];
    float *rm = e->robot_map + nn * r;
    for (size_t i = 0; i < nn; ++i) rm[i] = fill;
    if (p->pUpdateSensorView) env_update_sensor_view(e, r);
    if (p->pUpdateRobotMap) env_update_robot_map(e, r);
    if (p->pUpdateExplorationMap) {
      float *em = e->robot_expl + nn * r;
      for (size_t i = 0; i < nn; ++i) em[i] = 1.f;
      env_update_robot_expl(e, r);
    }
  }
  /* CoverageSystem::InitSetup, coverage_system.cpp:201-226 */
  for (size_t i = 0; i < nn; ++i) e->exploration_map[i] = 1.f;
  env_post_step(e);
  return e;
}

void orc_env_destroy(orc_env *e) {
  if (!e) return;
  free(e->world), free(e->robot_map), free(e->robot_expl), free(e->sensor_view);
  free(e->system_map), free(e->exploration_map), free(e->explored_idf);
  free(e->gstart), free(e->gcur), free(e->lcur), free(e->gpos), free(e->hist), free(e->hist_len);
  free(e);
}

/* Eigen normalized()/normalize(): v / sqrt(v.v) when v.v > 0 */
static void eigen_normalize(double *v) {
  double z = v[0] * v[0] + v[1] * v[1];
  if (z > 0) {
    double s = sqrt(z);
    v[0] = v[0] / s;
    v[1] = v[1] / s;
  }
}

int orc_env_step_actions(orc_env *e, const double *actions) {
  /* coverage_system.h:279-291 + robot_model.h:200-226 */
  const orc_params *p = &e->p;
  for (int r = 0; r < e->R; ++r) {
    double ax = actions[2 * r], ay = actions[2 * r + 1];
    double speed = sqrt(ax * ax + ay * ay);
    double dir[2] = {ax, ay};
    eigen_normalize(dir);
    double sp = speed;
    if (sp > p->pMaxRobotSpeed) sp = p->pMaxRobotSpeed;
    double dn = sqrt(dir[0] * dir[0] + dir[1] * dir[1]);
    int invalid_dir = dn < K_EPS && sp >= K_EPS;
    if (sp < 0 || invalid_dir) return 1;
    eigen_normalize(dir);
    double nx = e->lcur[2 * r], ny = e->lcur[2 * r + 1];
    if (sp > K_EPS) {
      nx = e->lcur[2 * r] + dir[0] * sp * p->pTimeStep;
      ny = e->lcur[2 * r + 1] + dir[1] * sp * p->pTimeStep;
    }
    env_set_local(e, r, nx, ny);
  }
  env_post_step(e);
  return 0;
}

/* CoverageSystem::ClearRobotMaps, coverage_system.h:399-403 -> RobotModel::ClearRobotMap, robot_model.h:180-195:
 * refill every robot map, then re-sense at the current position (the exploration maps are kept) */
void orc_env_clear_robot_maps(orc_env *e) {
  const orc_params *p = &e->p;
  size_t nn = (size_t)e->N * e->N;
  float fill = p->pRobotMapUseUnknownImportance ? (float)(p->pUnknownImportance * p->pNorm) : 0.f;
  for (int r = 0; r < e->R; ++r) {
    float *rm = e->robot_map + nn * r;
    for (size_t i = 0; i < nn; ++i) rm[i] = fill;
    if (p->pUpdateSensorView) env_update_sensor_view(e, r);
    if (p->pUpdateRobotMap) env_update_robot_map(e, r);
  }
}

/* CoverageSystem::ClearExploredIDF, coverage_system.h:405-408 (system_map_ is only refreshed by the next step) */
void orc_env_clear_explored_idf(orc_env *e) {
  memset(e->explored_idf, 0, sizeof(float) * (size_t)e->N * e->N);
}

void orc_env_set_global_positions(orc_env *e, const double *pos) {
  /* coverage_system.h:235-244 */
  for (int r = 0; r < e->R; ++r) env_set_global(e, r, pos[2 * r], pos[2 * r + 1]);
  env_post_step(e);
}

void orc_env_get_positions(const orc_env *e, double *out) {
  memcpy(out, e->gcur, sizeof(double) * 2 * e->R);
}

const float *orc_env_world_map(const orc_env *e) { return e->world; }
const float *orc_env_system_map(const orc_env *e) { return e->system_map; }
const float *orc_env_explored_map(const orc_env *e) { return e->explored_idf; }
const float *orc_env_exploration_map(const orc_env *e) { return e->exploration_map; }
const float *orc_env_robot_map(const orc_env *e, int r) { return e->robot_map + (size_t)r * e->N * e->N; }
const float *orc_env_sensor_view(const orc_env *e, int r) { return e->sensor_view + (size_t)r * e->S * e->S; }

void orc_env_compute_voronoi(orc_env *e, orc_cell *cells) {
  /* coverage_system.h:371-377 (always over the true world map).  Note the box
   * is (pWorldMapSize, pWorldMapSize), not scaled by the resolution. */
  memcpy(e->gpos, e->gcur, sizeof(double) * 2 * e->R);
  /* voronoi.h:153 copies the whole map */
  size_t nn = (size_t)e->N * e->N;
  float *copy = (float *)malloc(sizeof(float) * nn);
  memcpy(copy, e->world, sizeof(float) * nn);
  orc_voronoi(e->gpos, e->R, copy, e->N, e->N, e->N, e->p.pResolution, 0, 0, cells);
  free(copy);
}

double orc_env_objective(orc_env *e) {
  /* coverage_system.h:598-601 + voronoi.h:179-185 */
  orc_cell *cells = (orc_cell *)malloc(sizeof(orc_cell) * e->R);
  orc_env_compute_voronoi(e, cells);
  double obj = 0;
  for (int r = 0; r < e->R; ++r) obj = obj + cells[r].sum_idf_site_dist_sqr;
  free(cells);
  return obj;
}

void orc_env_local_map(orc_env *e, int r, float *out) {
  /* robot_model.h:303-313 + map_utils.h:88-97 */
  const orc_params *p = &e->p;
  int L = e->L;
  memset(out, 0, sizeof(float) * L * L);
  double px = e->gcur[2 * r], py = e->gcur[2 * r + 1];
  if (is_point_outside_boundary(p->pResolution, px, py, L, e->N)) return;
  orc_bounds ix, of;
  orc_compute_offsets(p->pResolution, px, py, L, p->pRobotMapSize, &ix, &of);
  const float *rm = e->robot_map + (size_t)r * e->N * e->N;
  for (int a = 0; a < of.width; ++a)
    memcpy(out + (size_t)(of.left + a) * L + of.bottom,
           rm + (size_t)(ix.left + of.left + a) * e->N + ix.bottom + of.bottom,
           sizeof(float) * of.height);
}

void orc_env_obstacle_map(orc_env *e, int r, float *out) {
  /* robot_model.h:336-345 */
  const orc_params *p = &e->p;
  int L = e->L;
  for (int i = 0; i < L * L; ++i) out[i] = 1.f;
  orc_bounds ix, of;
  orc_compute_offsets(p->pResolution, e->gcur[2 * r], e->gcur[2 * r + 1], L, p->pRobotMapSize, &ix, &of);
  for (int a = 0; a < of.width; ++a)
    memset(out + (size_t)(of.left + a) * L + of.bottom, 0, sizeof(float) * of.height);
}

int orc_env_neighbors(const orc_env *e, int r, double *rel) {
  /* coverage_system.cpp:274-291: for robot r the list is ordered by the
   * (i<j) pair loop, i.e. ascending neighbour index */
  int n = 0;
  for (int j = 0; j < e->R; ++j) {
    if (j == r) continue;
    double dx, dy;
    if (j > r) {
      dx = e->gpos[2 * j] - e->gpos[2 * r], dy = e->gpos[2 * j + 1] - e->gpos[2 * r + 1];
    } else { /* stored as -(p_r - p_j) */
      dx = -(e->gpos[2 * r] - e->gpos[2 * j]), dy = -(e->gpos[2 * r + 1] - e->gpos[2 * j + 1]);
    }
    if (sqrt(dx * dx + dy * dy) < e->p.pCommunicationRange) rel[2 * n] = dx, rel[2 * n + 1] = dy, ++n;
  }
  return n;
}

void orc_env_comm_maps_cpp(orc_env *e, int r, int ms, float *ox, float *oy) {
  /* coverage_system.cpp:177-199 */
  const orc_params *p = &e->p;
  memset(ox, 0, sizeof(float) * ms * ms);
  memset(oy, 0, sizeof(float) * ms * ms);
  double *rel = (double *)malloc(sizeof(double) * 2 * e->R);
  int n = orc_env_neighbors(e, r, rel);
  double center = ms / 2. - p->pResolution / 2.;
  for (int k = 0; k < n; ++k) {
    double sx = rel[2 * k] * ms / (p->pCommunicationRange * p->pResolution * 2.) + center;
    double sy = rel[2 * k + 1] * ms / (p->pCommunicationRange * p->pResolution * 2.) + center;
    int ix = (int)round(sx), iy = (int)round(sy);
    /* the reference writes comm_map(ix, iy) unchecked (undefined behaviour once pResolution < 1 pushes the index
     * past the map); the restatement skips such neighbours instead of corrupting the heap */
    if (ix < 0 || ix >= ms || iy < 0 || iy >= ms) continue;
    ox[ix * ms + iy] += (float)(rel[2 * k] / p->pCommunicationRange);
    oy[ix * ms + iy] += (float)(rel[2 * k + 1] / p->pCommunicationRange);
  }
  free(rel);
}

void orc_env_comm_maps_py(orc_env *e, int r, int ms, float *ox, float *oy) {
  /* coverage_env_utils.py:143-191: float32 tensors, torch.round = half-to-even */
  const orc_params *p = &e->p;
  memset(ox, 0, sizeof(float) * ms * ms);
  memset(oy, 0, sizeof(float) * ms * ms);
  double *rel = (double *)malloc(sizeof(double) * 2 * e->R);
  int n = orc_env_neighbors(e, r, rel);
  float denom = (float)(p->pCommunicationRange * p->pResolution * 2.0);
  float center = (float)(ms / 2.0 - p->pResolution / 2.0);
  float cr = (float)p->pCommunicationRange;
  for (int k = 0; k < n; ++k) {
    float fx = (float)rel[2 * k], fy = (float)rel[2 * k + 1];
    float sx = fx * (float)ms / denom + center;
    float sy = fy * (float)ms / denom + center;
    long ix = (long)rintf(sx), iy = (long)rintf(sy); /* default rounding mode: ties-to-even */
    /* torch.sparse_coo_tensor raises for an index outside the map (pResolution < 1 can produce one): skipped here,
     * as the CUDA path does, instead of writing past the buffer */
    if (ix < 0 || ix >= ms || iy < 0 || iy >= ms) continue;
    ox[ix * ms + iy] += fx / cr;
    oy[ix * ms + iy] += fy / cr;
  }
  free(rel);
}

int orc_env_check_oscillation(const orc_env *e, int r) {
  /* coverage_system.h:341-363 */
  if (!e->p.pCheckOscillations) return 0;
  int len = e->hist_len[r];
  if (len < 3) return 0;
  const double *h = e->hist + (size_t)r * e->H * 2;
  double lx = h[2 * (len - 1)], ly = h[2 * (len - 1) + 1];
  int span = 6 < len - 1 ? 6 : len - 1;
  int count = 0;
  for (int k = 0; k < span; ++k) {
    double dx = h[2 * (len - 1 - k)] - lx, dy = h[2 * (len - 1 - k) + 1] - ly;
    if (sqrt(dx * dx + dy * dy) < K_LARGE_EPS) ++count;
  }
  return count > 2;
}

void orc_env_voronoi_features(orc_env *e, int r, double *out7) {
  /* robot_model.h:408-430 + voronoi.h:73-78 */
  const orc_params *p = &e->p;
  orc_bounds ix, of;
  double px = e->gcur[2 * r], py = e->gcur[2 * r + 1];
  orc_compute_offsets(p->pResolution, px, py, e->L, e->N, &ix, &of);
  const float *rm = e->robot_map + (size_t)r * e->N * e->N +
                    (size_t)(ix.left + of.left) * e->N + ix.bottom + of.bottom;
  double tx = (ix.left + of.left) * p->pResolution, ty = (ix.bottom + of.bottom) * p->pResolution;
  double site[2] = {px - tx, py - ty};
  orc_cell c;
  orc_voronoi(site, 1, rm, e->N, of.width, of.height, p->pResolution, 1, 0, &c);
  out7[0] = c.centroid[0], out7[1] = c.centroid[1], out7[2] = c.mass;
  out7[3] = c.sum_idf_site_dist_sqr, out7[4] = 0, out7[5] = c.sum_idf_site_dist, out7[6] = 0;
}

/* frontier_pq.cpp: the real std::priority_queue of robot_model.h:383-401 */
int orc_frontier_select(const double *values, int n, int num_frontiers, int *order_out);

void orc_env_exploration_features(orc_env *e, int r, int num_frontiers, double *out) {
  /* robot_model.h:347-403 GetExplorationFeatures */
  const orc_params *p = &e->p;
  const double posx = e->gcur[2 * r], posy = e->gcur[2 * r + 1];
  const double tsd = p->pMaxRobotSpeed * p->pTimeStep * p->pResolution; /* :150-151 */
  const double sensor_area = (double)(p->pSensorSize * p->pSensorSize);  /* :152 */
  const double factor = 4;
  const double pos_x_lim = fmin(posx + factor * tsd, (double)p->pWorldMapSize);
  const double pos_y_lim = fmin(posy + factor * tsd, (double)p->pWorldMapSize);
  const float *rm = e->robot_map + (size_t)r * e->N * e->N;
  const float *ex = e->robot_expl ? e->robot_expl + (size_t)r * e->N * e->N : NULL;
  int cap = 1024, n = 0;
  double *val = (double *)malloc(sizeof(double) * cap), *qx = (double *)malloc(sizeof(double) * cap),
         *qy = (double *)malloc(sizeof(double) * cap);
  for (double px = fmax(posx - factor * tsd, 1e-4); px < pos_x_lim; px += 1 * p->pResolution) {
    for (double py = fmax(posy - factor * tsd, 1e-4); py < pos_y_lim; py += 1 * p->pResolution) {
      double dx = px - posx, dy = py - posy;
      double dist = sqrt(dx * dx + dy * dy);
      dist = fmax(dist, tsd);
      orc_bounds ix, of;
      orc_compute_offsets(p->pResolution, px, py, p->pSensorSize, p->pWorldMapSize, &ix, &of);
      float unexplored_f = 0.f; /* Eigen block().sum() over floats 0/1: exact */
      if (ex)
        for (int a = 0; a < of.width; ++a)
          for (int b = 0; b < of.height; ++b)
            unexplored_f += ex[(size_t)(ix.left + of.left + a) * e->N + ix.bottom + of.bottom + b];
      double unexplored = unexplored_f;
      double benefit = unexplored;
      if (unexplored > 0.1 * sensor_area) {
        long cnt = 0; /* block().count(): non-zero coefficients */
        for (int a = 0; a < of.width; ++a)
          for (int b = 0; b < of.height; ++b)
            cnt += rm[(size_t)(ix.left + of.left + a) * e->N + ix.bottom + of.bottom + b] != 0.f;
        double idf_value = (double)cnt;
        benefit += 2 * idf_value;
      }
      if (n == cap) {
        cap *= 2;
        val = (double *)realloc(val, sizeof(double) * cap);
        qx = (double *)realloc(qx, sizeof(double) * cap);
        qy = (double *)realloc(qy, sizeof(double) * cap);
      }
      val[n] = benefit / dist, qx[n] = px, qy[n] = py, ++n;
    }
  }
  int *order = (int *)malloc(sizeof(int) * (num_frontiers > 0 ? num_frontiers : 1));
  int got = orc_frontier_select(val, n, num_frontiers, order);
  for (int k = 0; k < 2 * num_frontiers; ++k) out[k] = 0.0;
  for (int k = 0; k < got; ++k) out[2 * k] = qx[order[k]] - posx, out[2 * k + 1] = qy[order[k]] - posy;
  free(order), free(val), free(qx), free(qy);
}

/* ========================================================================== */
/* Controllers                                                                 */
/* ========================================================================== */

static int decentralized_goals(orc_env *e, double *goals) {
  /* decentralized_cvt.h:85-143 */
  const orc_params *p = &e->p;
  int err = 0;
  const double *pos = e->gcur;
  size_t nn = (size_t)e->N * e->N;
#pragma omp parallel for schedule(dynamic, 1)
  for (int r = 0; r < e->R; ++r) {
    orc_bounds ix, of;
    orc_compute_offsets(p->pResolution, pos[2 * r], pos[2 * r + 1], e->L, e->N, &ix, &of);
    /* :92 copies the robot's whole map; kept because it is the reference's cost */
    float *copy = (float *)malloc(sizeof(float) * nn);
    memcpy(copy, e->robot_map + nn * r, sizeof(float) * nn);
    const float *local = copy + (size_t)(ix.left + of.left) * e->N + ix.bottom + of.bottom;
    double tx = (ix.left + of.left) * p->pResolution, ty = (ix.bottom + of.bottom) * p->pResolution;
    double *sites = (double *)malloc(sizeof(double) * 2 * e->R);
    int n = 1;
    sites[0] = pos[2 * r] - tx, sites[1] = pos[2 * r + 1] - ty;
    for (int j = 0; j < e->R; ++j) {
      if (j == r) continue;
      double dx = pos[2 * j] - pos[2 * r], dy = pos[2 * j + 1] - pos[2 * r + 1];
      if (sqrt(dx * dx + dy * dy) < p->pCommunicationRange) {
        sites[2 * n] = pos[2 * j] - tx, sites[2 * n + 1] = pos[2 * j + 1] - ty, ++n;
      }
    }
    orc_cell c;
    int ve = orc_voronoi(sites, n, local, e->N, of.width, of.height, p->pResolution, 1, 0, &c);
    goals[2 * r] = c.centroid[0] + sites[0] + tx;
    goals[2 * r + 1] = c.centroid[1] + sites[1] + ty;
    if (ve || goals[2 * r] < 0 || goals[2 * r] > e->N || goals[2 * r + 1] < 0 || goals[2 * r + 1] > e->N) {
#pragma omp atomic write
      err = 1;
    }
    free(sites);
    free(copy);
  }
  return err;
}

int orc_ctrl_compute_actions(orc_env *e, int kind, double *actions, double *goals) {
  /* clairvoyant_cvt.h:89-127, centralized_cvt.h:90-130, decentralized_cvt.h:145-173 */
  const orc_params *p = &e->p;
  int R = e->R;
  memcpy(e->gpos, e->gcur, sizeof(double) * 2 * R);
  if (kind == ORC_DECENTRALIZED) {
    if (decentralized_goals(e, goals)) return -1;
  } else {
    orc_cell *cells = (orc_cell *)malloc(sizeof(orc_cell) * R);
    const float *src = kind == ORC_CLAIRVOYANT ? e->world : e->explored_idf;
    size_t nn = (size_t)e->N * e->N;
    float *copy = (float *)malloc(sizeof(float) * nn); /* voronoi.h:153 */
    memcpy(copy, src, sizeof(float) * nn);
    int ve = orc_voronoi(e->gpos, R, copy, e->N, e->N, e->N, p->pResolution, 0, 0, cells);
    free(copy);
    for (int r = 0; r < R; ++r) goals[2 * r] = cells[r].centroid[0], goals[2 * r + 1] = cells[r].centroid[1];
    free(cells);
    if (ve) return -1;
  }
  int converged = 1;
  for (int r = 0; r < R; ++r) {
    actions[2 * r] = actions[2 * r + 1] = 0;
    double dx = goals[2 * r] - e->gpos[2 * r], dy = goals[2 * r + 1] - e->gpos[2 * r + 1];
    double dist = sqrt(dx * dx + dy * dy);
    if (dist < K_EPS) continue;
    if (orc_env_check_oscillation(e, r)) continue;
    double speed = dist / p->pTimeStep;
    speed = speed < p->pMaxRobotSpeed ? speed : p->pMaxRobotSpeed;
    double dir[2] = {dx, dy};
    eigen_normalize(dir);
    actions[2 * r] = speed * dir[0];
    actions[2 * r + 1] = speed * dir[1];
    converged = 0;
  }
  return converged;
}
