/*
 * cc_oracle.h -- CPU oracle for the CoverageControl grid hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (KumarRobotics/CoverageControl), written from the reference's
 * behaviour, each function citing the reference file:line it follows.  It is
 * the checker for the CUDA path in coveragecontrol_b200/ and the "port" CPU
 * baseline timed by bench.py.  Nothing in the product path may call it.
 *
 * Parity pin: the three known answers of the reference's own
 * python/tests/test_coverage.py:31-115 (98.90 %/166, 58.13 %/104, 8.87 %/17)
 * are reproduced by tests/test_oracle_kat.py through this file.
 * Polygon-feature IDF values and the 32x32 resize have no in-tree golden
 * vector in the reference: "parity unpinned" for those two.
 */
#ifndef CC_ORACLE_H_
#define CC_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors CoverageControl::Parameters (parameters.h:49-175); only the fields
 * the hot path reads. */
typedef struct orc_params {
  int pNumRobots;
  double pResolution;
  int pWorldMapSize;
  int pRobotMapSize;
  int pLocalMapSize;
  int pUpdateRobotMap;
  int pUpdateExplorationMap;
  int pUpdateSensorView;
  int pUpdateSystemMap;
  double pTruncationBND;
  double pNorm;
  double pUnknownImportance;
  int pRobotMapUseUnknownImportance;
  int pSensorSize;
  double pCommunicationRange;
  double pMaxRobotSpeed;
  int pRobotPosHistorySize;
  double pTimeStep;
  int pCheckOscillations;
} orc_params;

void orc_params_default(orc_params *p);

/* ---- IDF ---------------------------------------------------------------- */
/* bnd: F x 6 doubles {mean_x, mean_y, sigma_x, sigma_y, scale, rho}
 * polygons (y-monotone pieces, CUDA semantics): SoA px/py (doubles), sz[], imp[]
 * mode 0: reference CPU arithmetic (world_idf.cpp:41-101; fp32, sigma.x on both axes)
 * mode 1: reference CUDA arithmetic restated on the CPU (generate_world_map.cu:76-164;
 *         fp32, per-axis sigma, no FMA contraction, glibc erfcf)
 * mode 2: fp64 separable form of the same integral (truth for tolerance tests)
 * out: float32[N*N] row-major [x][y]; returns normalization factor. */
float orc_idf_generate(const orc_params *p, const double *bnd, int F,
                       const double *px, const double *py, const int *sz,
                       const float *imp, int num_poly, int mode, float *out);
/* fp64 map (unnormalised) + count of features per cell whose truncation test is
 * within `band` of the threshold (borderline cells). */
void orc_idf_fp64(const orc_params *p, const double *bnd, int F, double band,
                  double *out, unsigned char *borderline);
int orc_point_in_monotone_polygon(const float *x, const float *y, int sz,
                                  float rx, float ry);

/* ---- windows ------------------------------------------------------------ */
typedef struct orc_bounds {
  int left, right, bottom, top, width, height;
} orc_bounds;
void orc_compute_offsets(double res, double px, double py, int submap,
                         int map_size, orc_bounds *index, orc_bounds *offset);

/* ---- Voronoi ------------------------------------------------------------ */
typedef struct orc_cell {
  double site[2];
  double mass;
  double centroid[2];
  double sum_idf_site_dist;
  double sum_idf_site_dist_sqr;
  int nverts;
  double verts[2 * 64]; /* polygon vertices, CCW */
} orc_cell;

/* Reference-faithful: convex cell polygon + scan-line rasterisation
 * (voronoi.cpp:55-150,167-327).  map: float32 [W][H] with row stride `stride`.
 * single != 0 : compute_single mode for site `robot_id` (origin shift -sites[0]).
 * cells: n entries (or 1 when single).  Returns 0 on success. */
int orc_voronoi(const double *sites, int n, const float *map, int stride,
                double map_w, double map_h, double res, int single,
                int robot_id, orc_cell *cells);
/* Stated tie rule of the build: label = argmin_k d2, lowest index wins, d2 in
 * fp64 without FMA.  labels: int32 [W][H].  sums: n x 4 {mass, sum cx v, sum cy v,
 * sum d2 v}. */
void orc_voronoi_bruteforce(const double *sites, int n, const float *map,
                            int stride, int W, int H, double res, int *labels,
                            double *sums);
void orc_voronoi_labels_omp(const double *sites, int n, int W, int H, double res, int *labels);
/* Convex polygon of site k: rectangle [0,w]x[0,h] clipped by the bisector
 * half-planes.  Returns vertex count. */
int orc_clip_cell(const double *sites, int n, int k, double w, double h,
                  double *verts, int max_verts);

/* ---- environment (structure-faithful CPU baseline) ------------------------ */
typedef struct orc_env orc_env;
orc_env *orc_env_create(const orc_params *p, const float *world_map,
                        const double *positions, int num_robots);
void orc_env_destroy(orc_env *e);
int orc_env_step_actions(orc_env *e, const double *actions);
void orc_env_set_global_positions(orc_env *e, const double *pos);
void orc_env_clear_robot_maps(orc_env *e);
void orc_env_clear_explored_idf(orc_env *e);
void orc_env_get_positions(const orc_env *e, double *out);
double orc_env_objective(orc_env *e);
void orc_env_compute_voronoi(orc_env *e, orc_cell *cells);
const float *orc_env_world_map(const orc_env *e);
const float *orc_env_system_map(const orc_env *e);
const float *orc_env_explored_map(const orc_env *e);
const float *orc_env_exploration_map(const orc_env *e);
const float *orc_env_robot_map(const orc_env *e, int r);
const float *orc_env_sensor_view(const orc_env *e, int r);
void orc_env_local_map(orc_env *e, int r, float *out);     /* L x L */
void orc_env_obstacle_map(orc_env *e, int r, float *out);  /* L x L */
/* C++ semantics (coverage_system.cpp:177-199) */
void orc_env_comm_maps_cpp(orc_env *e, int r, int map_size, float *out_x,
                           float *out_y);
/* Python/LPAC semantics (coverage_env_utils.py:143-191): fp32 + round-half-even */
void orc_env_comm_maps_py(orc_env *e, int r, int map_size, float *out_x,
                          float *out_y);
int orc_env_neighbors(const orc_env *e, int r, double *rel /* R x 2 */);
void orc_env_voronoi_features(orc_env *e, int r, double *out7);
/* robot_model.h:347-403: 2 * num_frontiers doubles [dx, dy, ...], worst frontier first, zero-padded;
 * the frontier queue is the real std::priority_queue (frontier_pq.cpp) */
void orc_env_exploration_features(orc_env *e, int r, int num_frontiers, double *out);
int orc_env_check_oscillation(const orc_env *e, int r);

/* ---- controllers -------------------------------------------------------- */
enum { ORC_CLAIRVOYANT = 0, ORC_CENTRALIZED = 1, ORC_DECENTRALIZED = 2 };
/* returns 1 if converged (all actions zero), fills actions and goals (R x 2) */
int orc_ctrl_compute_actions(orc_env *e, int kind, double *actions,
                             double *goals);

#ifdef __cplusplus
}
#endif
#endif /* CC_ORACLE_H_ */
