/*
 * ccb200.h -- C ABI of the B200-native CoverageControl grid hot path.
 *
 * This is the drop-in boundary: the entry points below are what a
 * `coverage_control._core` shim binds in place of the reference's pybind11
 * module (reference cppsrc/python_bindings/core_binds.h).  Plain pointers and
 * sizes only; every call returns an int status (0 = ok) and never throws or
 * exits across the ABI; ccb_last_error() holds the message of the last failure
 * on the calling thread.
 *
 * Conventions (reference typedefs.h:44-48, world_idf.cpp:54-77):
 *   maps are float32, row-major [x_index][y_index] (y fastest), N = pWorldMapSize;
 *   points are float64 pairs (x, y); "E" is the number of environments held by
 *   one ccb_env (E = 1 reproduces one reference CoverageSystem), "R" robots each.
 * Pointers named *_dev are device pointers on the env's GPU; all others are host.
 * All work is issued on the stream given at creation; host-returning calls
 * synchronise that stream, nothing else does.
 */
#ifndef CCB200_H_
#define CCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCB_ABI_VERSION 1
#define CCB_MAX_CELL_VERTS 64

/* reference: CoverageControl::Parameters, parameters.h:49-175 */
typedef struct ccb_params {
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
} ccb_params;

typedef struct ccb_idf ccb_idf;   /* reference: class WorldIDF, world_idf.h:61 */
typedef struct ccb_env ccb_env;   /* reference: class CoverageSystem, coverage_system.h:66 (x E) */
typedef struct ccb_ctrl ccb_ctrl; /* reference: Clairvoyant/Centralized/DecentralizedCVT */

/* ---- library ------------------------------------------------------------ */
int ccb_abi_version(void);
const char *ccb_last_error(void);
/* reference: CudaUtils::IsCudaAvailable, cuda_utils.h:59-61.  1 if a usable GPU exists. */
int ccb_cuda_available(void);
int ccb_device_count(void);
void ccb_params_default(ccb_params *p); /* parameters.h class defaults */

/* ---- WorldIDF ----------------------------------------------------------- */
/* world_idf.h:100-113 WorldIDF(params); map zero-filled (the reference leaves it uninitialised) */
ccb_idf *ccb_idf_create(const ccb_params *p, int device, void *stream);
void ccb_idf_destroy(ccb_idf *idf);
/* world_idf.h:207-218 AddNormalDistribution; rho is stored but, as in the reference
 * (bivariate_normal_distribution.h:103-112), the general-form constructor forces rho = 0 */
int ccb_idf_add_bnd(ccb_idf *idf, double mean_x, double mean_y, double sigma_x,
                    double sigma_y, double rho, double scale);
/* world_idf.h:202-204 AddUniformDistributionPolygon; xy = n x 2 doubles */
int ccb_idf_add_polygon(ccb_idf *idf, const double *xy, int n, float importance);
int ccb_idf_num_features(const ccb_idf *idf); /* world_idf.h:262-264 */
/* world_idf.h:227-241 GenerateMap / world_idf.cpp:104-163 GenerateMapCuda.
 * mode 0: arithmetic of the reference CUDA kernel (generate_world_map.cu:76-164), bit for bit
 * mode 1: fast separable form (erfc x1 - erfc x2)(erfc y1 - erfc y2), same truncation */
int ccb_idf_generate(ccb_idf *idf, int mode);
/* row-band variant for a world tiled over several GPUs: generate rows [row0,row1)
 * un-normalised; *ccb_idf_band_max_dev = the band maximum as an order-preserving int32 (all-reduce
 * it with MAX), then ccb_idf_normalize_band applies generate_world_map.cu:252-259 to the band. */
int ccb_idf_generate_band(ccb_idf *idf, int mode, int row0, int row1);
int *ccb_idf_band_max_dev(ccb_idf *idf);
int ccb_idf_normalize_band(ccb_idf *idf, int row0, int row1);
float *ccb_idf_map_dev_mut(ccb_idf *idf);
int ccb_idf_get_map(ccb_idf *idf, float *out);       /* world_idf.h:258  GetWorldMap (D2H copy) */
int ccb_idf_set_map(ccb_idf *idf, const float *in);  /* world_idf.h:115-127 / GetWorldMapMutable */
const float *ccb_idf_map_dev(const ccb_idf *idf);
/* world_idf.h:173-196 copy constructor: the generated map and its normalisation factor, device to device */
int ccb_idf_copy_map(ccb_idf *dst, const ccb_idf *src);
int ccb_idf_norm_factor(ccb_idf *idf, float *out);   /* world_idf.h:256 */
/* feature records back out, for WriteDistributions (world_idf.cpp:166-195) */
int ccb_idf_get_bnds(const ccb_idf *idf, double *out6 /* F x 6 */, int cap);
int ccb_idf_num_bnds(const ccb_idf *idf);
int ccb_idf_num_polygons(const ccb_idf *idf);
int ccb_idf_get_polygon(const ccb_idf *idf, int k, double *xy, int cap, float *imp);

/* ---- CoverageSystem (batch of E environments) ----------------------------- */
ccb_env *ccb_env_create(const ccb_params *p, int num_envs, int num_robots,
                        int device, void *stream);
void ccb_env_destroy(ccb_env *env);
int ccb_env_num_envs(const ccb_env *env);
int ccb_env_num_robots(const ccb_env *env);
/* coverage_system.h:261-264 SetWorldIDF (deep copy of the map) */
int ccb_env_set_world_idf(ccb_env *env, int e, const ccb_idf *idf);
int ccb_env_set_world_host(ccb_env *env, int e, const float *map);
int ccb_env_set_world_dev(ccb_env *env, int e, const float *map_dev);
/* RobotModel::Initialize (robot_model.h:124-153) + CoverageSystem::InitSetup
 * (coverage_system.cpp:201-226) for every environment; positions = E x R x 2 */
int ccb_env_reset(ccb_env *env, const double *positions);
/* the same for ONE environment of the batch (a fresh episode in slot e; set its world map first) */
int ccb_env_reset_env(ccb_env *env, int e, const double *positions /* R x 2 */);
/* coverage_system.h:279-291 StepActions; actions = E x R x 2.  The reference splits an action into (direction,
 * speed = |action|) and StepControl rejects speed < 0 or a zero direction with speed > 0 (robot_model.h:208-216):
 * neither can happen for an action vector, so the batch call never reports an invalid control -- it returns 0 (or 1
 * with ccb_last_error for a CUDA failure) and leaves *bad_env = -1.  ccb_env_step_control keeps the check. */
int ccb_env_step_actions(ccb_env *env, const double *actions, int *bad_env);
/* coverage_system.h:303-333 StepAction / StepControl (single robot of env e) */
int ccb_env_step_control(ccb_env *env, int e, int robot, double dir_x, double dir_y, double speed);
/* coverage_system.h:218-257: local=1 SetRobotPositions/SetLocalRobotPositions, local=0
 * SetGlobalRobotPositions; robot < 0 means all robots (positions = R x 2) */
int ccb_env_set_positions(ccb_env *env, int e, int robot, int local, const double *positions);
int ccb_env_get_positions(ccb_env *env, double *out /* E x R x 2 */); /* coverage_system.h:514-525 */
int ccb_env_clear_robot_maps(ccb_env *env, int e);   /* coverage_system.h:399-403 */
int ccb_env_clear_explored_idf(ccb_env *env, int e); /* coverage_system.h:405-408 */

enum ccb_map_kind {
  CCB_MAP_WORLD = 0,             /* GetWorldMap                 N x N */
  CCB_MAP_SYSTEM = 1,            /* GetSystemMap                N x N */
  CCB_MAP_EXPLORED_IDF = 2,      /* GetSystemExploredIDFMap     N x N */
  CCB_MAP_EXPLORATION = 3,       /* GetSystemExplorationMap     N x N */
  CCB_MAP_ROBOT = 4,             /* GetRobotMap(id)             N x N */
  CCB_MAP_SENSOR_VIEW = 5,       /* GetRobotSensorView(id)      S x S */
  CCB_MAP_LOCAL = 6,             /* GetRobotLocalMap(id)        L x L */
  CCB_MAP_OBSTACLE = 7,          /* GetRobotObstacleMap(id)     L x L */
  CCB_MAP_ROBOT_EXPLORATION = 8, /* GetRobotExplorationMap(id)  L x L */
  CCB_MAP_ROBOT_SYSTEM = 9       /* GetRobotSystemMap(id)       L x L */
};
/* dense materialisation of one map (coverage_system.h:471-583, robot_model.h:293-345) */
int ccb_env_get_map(ccb_env *env, int e, int kind, int robot, float *out);
int ccb_env_get_map_dev(ccb_env *env, int e, int kind, int robot, float *out_dev);
/* dimension (rows == cols) of a map kind */
int ccb_env_map_dim(const ccb_env *env, int kind);

/* batched LPAC observation tensors on the device (coverage_env_utils.py:97-277):
 *   local/obstacle: E x R x L x L ; comm: E x R x 2 x C x C (python float32 semantics);
 *   cnn_input: E x R x (use_comm ? 4 : 2) x C x C = [local, comm_x, comm_y, obstacle],
 *   resized with torchvision's antialias bilinear weights. */
int ccb_env_local_maps_dev(ccb_env *env, float *out_dev);
int ccb_env_obstacle_maps_dev(ccb_env *env, float *out_dev);
int ccb_env_comm_maps_dev(ccb_env *env, int C, float *out_dev);
int ccb_env_cnn_input_dev(ccb_env *env, int C, int use_comm, float *out_dev);
/* LPAC's communication graph (coverage_env_utils.py:313-335 get_weights, :368-369 normalised positions),
 * float32 on the device: weights E x R x R (exp(-d^2 / Rc^2), 0 below 1/e and on the diagonal),
 * pos_norm E x R x 2 ((p + N/2) / N); either pointer may be NULL */
int ccb_env_graph_dev(ccb_env *env, float *weights_dev, float *pos_norm_dev);
/* coverage_system.cpp:177-199 GetCommunicationMaps (C++ float64 semantics), host, env e: R x 2 x C x C */
int ccb_env_comm_maps_cpp(ccb_env *env, int e, int C, float *out);

/* coverage_system.h:371-377 ComputeVoronoiCells (always on the world map) */
int ccb_env_compute_voronoi(ccb_env *env);
/* coverage_system.h:635-638 GetVoronoiCells / voronoi.h:52-103.  Arrays of R entries for env e;
 * verts = R x CCB_MAX_CELL_VERTS x 2. */
int ccb_env_get_cells(ccb_env *env, int e, double *mass, double *centroid, double *site,
                      double *sum_dist_sqr, int *nverts, double *verts);
/* the same for every environment at once: mass E x R, centroid E x R x 2, sum_dist_sqr E x R (any may be NULL) */
int ccb_env_get_cells_all(ccb_env *env, double *mass, double *centroid, double *sum_dist_sqr);
int ccb_env_objective(ccb_env *env, double *out /* E */); /* coverage_system.h:598-601 */
int ccb_env_objective_dev(ccb_env *env, double *out_dev); /* the same, left on the device (env stream) */
/* parity-test accessor: the site index the fused label+reduce kernel assigns to every cell
 * (src 0: world map, 1: explored-IDF map); out = E x N x N int32 on the host.  The product path
 * never stores labels. */
int ccb_env_voronoi_labels(ccb_env *env, int src, int32_t *out);
/* coverage_system.h:612-619 GetRobotVoronoiFeatures: E x R x 7 */
int ccb_env_voronoi_features(ccb_env *env, double *out);
int ccb_env_voronoi_features_dev(ccb_env *env, double *out_dev); /* the same, left on the device */
/* coverage_system.cpp:648-676 GetLocalVoronoiFeatures (own cell among the robot and its neighbours): E x R x 7 */
int ccb_env_local_voronoi_features(ccb_env *env, double *out);
/* coverage_system.h:603-610 GetRobotExplorationFeatures -> robot_model.h:347-403 GetExplorationFeatures:
 * the num_frontiers (pNumFrontiers) best candidate points around every robot, worst first, as
 * [dx, dy] pairs relative to the robot, zero-padded: E x R x (2 num_frontiers) */
int ccb_env_exploration_features(ccb_env *env, int num_frontiers, double *out);
/* coverage_system.h:557-571 (ratio of explored cells / IDF-weighted): 2 doubles per env */
int ccb_env_exploration_ratios(ccb_env *env, double *out /* E x 2 */);
int ccb_env_check_oscillation(ccb_env *env, int e, int robot, int *out); /* coverage_system.h:341-363 */
int ccb_env_sync(ccb_env *env);
/* raw device state for zero-copy consumers (world: E x N x N floats) */
const float *ccb_env_world_dev(const ccb_env *env);
const double *ccb_env_positions_dev(const ccb_env *env);

/* ---- row bands: one large world tiled over the GPUs of a box ------------------------
 * Every rank holds the full-size buffers and the replicated robot state but labels/reduces only
 * its rows.  Per step: partials (each rank, its band) -> all-reduce(SUM) of E x R x 4 doubles by the
 * caller (NCCL) -> finalize (each rank) -> controllers / objective read the finalised cells.
 * No reference equivalent: the reference's Voronoi (voronoi.cpp:55-327) is single-process. */
int ccb_env_set_band_mode(ccb_env *env, int on);
/* moments (sum v, sum x v, sum y v, sum (x^2+y^2) v; cell units) of rows [row0,row1) per site;
 * src 0: world map, 1: explored-IDF map; sums_dev = E x R x 4 doubles on the device */
int ccb_env_voronoi_partials_dev(ccb_env *env, int src, int row0, int row1, double *sums_dev);
/* voronoi.cpp:41-53 + voronoi.h:89-95 from all-reduced moments; the env-level cells become current */
int ccb_env_finalize_from_partials_dev(ccb_env *env, const double *sums_dev);
/* DecentralizedCVT over several GPUs: robots' windows are independent, so the ranks share them out by robot.
 * Window moments (decentralized_cvt.h:85-143) of robots [robot0, robot1) -> sums_dev (E x R x 4, zeros elsewhere);
 * after the caller's all-reduce(SUM), ccb_env_set_window_sums_dev makes them the window moments of the current
 * positions (the next DecentralizedCVT ComputeActions uses them instead of reducing the windows itself). */
int ccb_env_window_partials_dev(ccb_env *env, int robot0, int robot1, double *sums_dev);
int ccb_env_set_window_sums_dev(ccb_env *env, const double *sums_dev);
/* The all-reduce(SUM) of those partials as ONE kernel over peer memory (NVLink / NVSwitch; csrc/peer.cuh): every rank
 * stores its vector into its slot of every peer's exchange buffer, raises a flag, waits for the others' flags and adds
 * the slots in RANK ORDER -- bit-identical sums on every rank and from run to run (NCCL's order is not
 * fixed), capturable in a CUDA graph (the exchange counter lives in state_dev).  peer_bufs: `world` device pointers,
 * entry p = rank p's exchange buffer (ccb_peer_buffer_bytes(world, count) bytes, zeroed before the first call, the
 * same symmetric allocation on every rank) as mapped in THIS process; state_dev: 16 zeroed bytes of local device
 * memory ([0] exchanges done, [1] non-zero once a peer failed to arrive within ~3 s).  in_dev may equal out_dev. */
size_t ccb_peer_buffer_bytes(int world, size_t count);
int ccb_peer_allreduce_sum(int device, void *stream, const double *in_dev, double *out_dev, size_t count,
                           const void *const *peer_bufs, int rank, int world, void *state_dev);
/* coverage_system.h:279-291 StepActions with the actions already on the device (E x R x 2) */
int ccb_env_step_actions_dev(ccb_env *env, const double *actions_dev);

/* ---- CVT controllers ---------------------------------------------------- */
enum ccb_ctrl_kind { CCB_CLAIRVOYANT = 0, CCB_CENTRALIZED = 1, CCB_DECENTRALIZED = 2 };
/* clairvoyant_cvt.h:68-83 etc.: constructor runs ComputeGoals once */
ccb_ctrl *ccb_ctrl_create(ccb_env *env, int kind);
void ccb_ctrl_destroy(ccb_ctrl *c);
/* ComputeActions (clairvoyant_cvt.h:100-127, centralized_cvt.h:102-130, decentralized_cvt.h:145-173);
 * returns 0, or 2 if a decentralized goal left the world (the reference throws) */
int ccb_ctrl_compute_actions(ccb_ctrl *c);
/* Host buffers of 32 KB and more that are page-locked (cudaHostAlloc / cudaHostRegister) are copied from / to
 * directly by ccb_ctrl_get_actions, ccb_ctrl_get_goals and ccb_env_step_actions; pageable ones pass through the
 * library's pinned slots.  Either way the call returns once the buffer has been written / read. */
int ccb_ctrl_get_actions(ccb_ctrl *c, double *out /* E x R x 2 */);
int ccb_ctrl_get_goals(ccb_ctrl *c, double *out /* E x R x 2 */);
int ccb_ctrl_is_converged(ccb_ctrl *c, int *out /* E */);
/* device-resident roll-out: `steps` x (ComputeActions; StepActions(GetActions())) for every
 * environment with no host round trip; with_objective != 0 adds GetObjectiveValue per step and
 * writes steps x E doubles to objective_dev (may be NULL). */
int ccb_ctrl_rollout(ccb_ctrl *c, int steps, int with_objective, double *objective_dev);
/* the same roll-out that also produces the LPAC observation (ccb_env_cnn_input_dev) after every step:
 * obs_dev + s * obs_step_stride floats receives step s (stride 0: only the last step is kept).  The
 * objective chain and the observation kernels of a step run on two streams side by side; everything is
 * ordered on the env stream when the call returns. */
int ccb_ctrl_rollout_obs(ccb_ctrl *c, int steps, int with_objective, double *objective_dev, int C, int use_comm,
                         float *obs_dev, size_t obs_step_stride);
const double *ccb_ctrl_actions_dev(const ccb_ctrl *c);

/* kernels launched by this library so far (bench.py reports the count of its timed region) */
uint64_t ccb_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CCB200_H_ */
