/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.cpp header). C interface for ctypes. */
#ifndef LEGOLOAM_ORACLE_H_
#define LEGOLOAM_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Packed point: x, y, z, intensity (16 B). The reference's pcl::PointXYZI is the same four
 * floats at byte 0/4/8/16 of a 32-byte struct (utility.h:45); tests repack. */
typedef struct { float x, y, z, i; } orc_p4;

/* Hot-path knobs; defaults = the reference's hard-coded values. */
typedef struct {
  float corner_leaf;      /* 0.2  mapOptmization.cpp:92 */
  float surf_leaf;        /* 0.4  mapOptmization.cpp:93 */
  float max_sq_dist;      /* 1.0  :1279 / :1389 */
  float eig_ratio;        /* 3    :1325 */
  double plane_tol;       /* 0.2  :1416  (double literal in the reference: float is promoted) */
  double weight_k;        /* 0.9  :1363 / :1425  (double literal) */
  double min_weight;      /* 0.1  :1370 / :1435  (double literal) */
  int   min_residuals;    /* 50   :1453 */
  float eig_thresh;       /* 100  :1519 */
  int   iter_count_thres; /* 200  loam_config.yaml:35 */
  float step_size;        /* 1.0  loam_config.yaml:32, used at :1500 */
  float stop_thres;       /* 0.05 loam_config.yaml:33, used at :1554 */
  int   apply_update;     /* 0 = as shipped (:1539-1545 commented out), 1 = upstream */
  int   eigen_mode;       /* 0 = as shipped (row 0 / row-masked V), 1 = upstream semantics */
  int   map_downsample;   /* 1 = VoxelGrid on the surrounding maps (:1225-1231) */
  int   min_corner_map;   /* 10   :1573 */
  int   min_surf_map;     /* 100  :1573 */
} orc_params;

/* State that persists across scans in the reference (members of MapOptimization). */
typedef struct {
  int   is_degenerate;    /* mapOptimization.h: isDegenerate */
  float matP[36];         /* row-major 6x6 */
  float min_lambda;
  float cf_mean;
} orc_state;

typedef struct {
  int   status;           /* 0 = optimised, 1 = skipped by the map-size guard (:1573) */
  int   iter_num;
  float min_lambda;
  float cf_mean;
  int   is_degenerate;
  int   converged;
  int   sel_num;          /* laserCloudSelNum of the last iteration */
  float pose[6];          /* transformTobeMapped after the loop */
  float last_x[6];        /* matX of the last iteration (after the degeneracy projection) */
} orc_result;

void orc_default_params(orc_params* p);
void orc_init_state(orc_state* s);

/* k-NN backend: 0 = brute force (canonical, ties by index), 1 = own uniform grid (same
 * canonical result), 2 = external (the reference's vendored nanoflann via oracle/_ref). */
typedef void* (*orc_ext_build_fn)(const orc_p4* pts, int n);
typedef void  (*orc_ext_knn5_fn)(void* tree, const float q[3], int* idx5, float* d2_5);
typedef void  (*orc_ext_free_fn)(void* tree);
void orc_set_knn_backend(int kind);
void orc_set_ext_knn(orc_ext_build_fn b, orc_ext_knn5_fn q, orc_ext_free_fn f);
/* radius search of the external tree (the reference's nanoflann): indices sorted by distance; returns the count */
typedef int (*orc_ext_radius_fn)(void* tree, const float q[3], double radius, int* idx, float* d2, int cap);
void orc_set_ext_radius(orc_ext_radius_fn f);

/* a2  transformPointCloud  mapOptmization.cpp:637-667; pose6 = [rx ry rz tx ty tz]. */
void orc_transform_cloud(const orc_p4* in, int n, const float pose6[6], orc_p4* out);

/* a3  pcl::VoxelGrid<PointXYZI>::filter. member[i] = output slot of input point i (or -1 for
 * non-finite points); returns 0, or 1 when the int32 voxel index would overflow (PCL then
 * passes the input through unchanged; so do we). */
int orc_voxel_downsample(const orc_p4* in, int n, float leaf, orc_p4* out, int* n_out, int* member);

/* a5  exact 5-NN. idx/d2 are [nq][5], ascending by (d2, idx). Missing neighbours: idx -1. */
void orc_knn5(const orc_p4* map, int n, const orc_p4* q, int nq, int* idx, float* d2, int backend);

/* a6  pointAssociateToMap  :606-620 */
void orc_point_associate_to_map(const orc_p4* in, int n, const float pose6[6], orc_p4* out);

/* a7 / a8. Per query (original order): mask[i]=1 if the residual is kept, coeff[i] = the
 * 4-float record, knn_idx[i][5] / knn_d2[i][5], accept[i] = (d2_5 < max_sq_dist). */
void orc_corner_residuals(const orc_p4* map, int n, const orc_p4* q, int nq, const float pose6[6],
                          const orc_params* prm, float* coeff, int* mask, int* accept,
                          int* knn_idx, float* knn_d2);
void orc_surf_residuals(const orc_p4* map, int n, const orc_p4* q, int nq, const float pose6[6],
                        const orc_params* prm, float* coeff, int* mask, int* accept,
                        int* knn_idx, float* knn_d2);

/* a9  LMOptimization(iterCount). ori/coeff are the compacted residual records. Outputs the
 * 6x6 AtA (row-major), Atb, x; updates pose (if apply_update) and state. Returns 1 when the
 * iteration reports convergence, 0 otherwise, -1 when skipped (n < min_residuals). */
int orc_lm_step(const orc_p4* ori, const float* coeff, int n, int iter, float pose6[6],
                const orc_params* prm, orc_state* st, float* AtA36, float* Atb6, float* x6,
                float* deltaR, float* deltaT);

/* a10 scan2MapOptimization on already-downsampled maps. */
void orc_scan2map(const orc_p4* corner_map, int nc, const orc_p4* surf_map, int ns,
                  const orc_p4* corner_q, int nqc, const orc_p4* surf_q, int nqs,
                  float pose6[6], const orc_params* prm, orc_state* st, orc_result* res);

/* a1  extractSurroundingKeyFrames tail: concatenate the selected keyframes (corner | surf then
 * outlier per keyframe, cache order = ids order), transformed by their poses, then VoxelGrid.
 * Clouds are given as one packed array + offsets (n_kf + 1 entries each). */
void orc_build_local_map(int n_kf, const float* poses6, const orc_p4* corner, const int* corner_off,
                         const orc_p4* surf, const int* surf_off, const orc_p4* outlier,
                         const int* outlier_off, const orc_params* prm, orc_p4* corner_ds,
                         int* n_corner_ds, orc_p4* surf_ds, int* n_surf_ds, int* n_corner_raw,
                         int* n_surf_raw);

/* Prior: transformAssociateToMap  mapOptmization.cpp:458-581 */
void orc_transform_associate_to_map(const float sum[6], const float bef[6], const float aft[6],
                                    float tobe[6]);

/* Replay harness (config 2): the per-scan body of MapOptimization::run (:1878-1888) with
 * loop closure on (sliding window of the last `window` keyframes, :1101-1151) and the pose
 * graph as pass-through. */
typedef struct orc_replay orc_replay;
orc_replay* orc_replay_create(const orc_params* prm, int window);
/* loop_closure_enabled = 0: the reference's DEFAULT local-map branch (mapOptmization.cpp:1152-1223,
 * loam_config.yaml:23): radius search on the key poses around currentRobotPosPoint, 1 m VoxelGrid on the hits (the
 * keyframe id is the (int) of the AVERAGED intensity, :1175/:1197/:1205), cached id list updated by erase / add. */
orc_replay* orc_replay_create2(const orc_params* prm, int window, int loop_closure_enabled, float search_radius);
int orc_replay_get_map_ids(orc_replay* r, int* ids, int cap);
int orc_select_surrounding(const orc_p4* key_poses, int n, const float cur_pos[3], float radius, int* existing_inout,
                           int n_existing, int cap, int use_ext_radius);
void orc_replay_destroy(orc_replay* r);
void orc_replay_set_params(orc_replay* r, const orc_params* prm);
/* One frame: clouds as the reference receives them after downsampleCurrentScan
 * (corner_scan raw = keyframe corner cloud, corner_scan_ds = corner queries, surf_last_ds and
 * outlier_last_ds = keyframe surf/outlier clouds, surf_total_ds = surf queries). odom = transformSum. */
void orc_replay_step(orc_replay* r, const float odom6[6], const orc_p4* corner_scan, int n_cs,
                     const orc_p4* corner_scan_ds, int n_csd, const orc_p4* surf_last_ds, int n_sl,
                     const orc_p4* outlier_last_ds, int n_ol, const orc_p4* surf_total_ds, int n_st,
                     orc_result* res, int* n_corner_map_ds, int* n_surf_map_ds);
void orc_replay_get_map(orc_replay* r, int which, orc_p4* out, int cap, int* n);

/* f-1: FeatureAssociation's normal-equation solves (featureAssociation.cpp:1925-1997, :2068-2131, :2208-2271)
 * with a column mask over [rx ry rz tx ty tz]; 3 or 6 unknowns. Returns 0, or -1 for other counts. */
int orc_solve_normal_eq(const float* rows6, const float* b, int n, int dof_mask, int iter, float eig_thresh,
                        int eigen_mode, float* x_out, float* matP_inout, int* is_degenerate_inout);
/* f-3: FeatureAssociation::findCorrespondingCornerFeatures / findCorrespondingSurfFeatures
 * (featureAssociation.cpp:1580-1697, :1699-1844) with TransformToStart (:1389-1412) inside.
 * which = 0 corner (two tripod points), 1 surf (three). q = cornerPointsSharp / surfPointsFlat, last =
 * laserCloudCornerLast / laserCloudSurfLast (intensity = ring + column/10000). ind1/2/3 [nq] are
 * pointSearch*Ind*: rewritten when iter % 5 == 0, read otherwise. coeff [nq][4] and keep [nq] describe the
 * (laserCloudOri, coeffSel) pair each query would push_back (keep = 0: nothing pushed). knn_backend: 1-NN by
 * brute force (ties: lowest index) or by the reference's nanoflann when the external backend is installed. */
void orc_fa_correspondences(int which, const orc_p4* q, int nq, const orc_p4* last, int n_last, const float transform_cur[6],
                            int iter, float nearest_feature_dist_sqr, int* ind1, int* ind2, int* ind3, float* coeff,
                            int* keep, int use_ext_knn);
void orc_fa_transform_to_start(const orc_p4* in, int n, const float transform_cur[6], orc_p4* out);

/* f-1 + f-3 as the reference runs them: FeatureAssociation::updateTransformation (featureAssociation.cpp:2505-2535) —
 * up to 100 iterations of findCorrespondingSurfFeatures + calculateTransformationSurf (:1846-2010; unknowns rx, rz, ty),
 * then up to 100 of findCorrespondingCornerFeatures + calculateTransformationCorner (:2013-2143; unknowns ry, tx, tz).
 * transform_cur is transformCur (in / out). State that outlives a call in the reference: isDegenerate (a class member,
 * shared by both loops, :223) and — by accident of the stack — matP, a function-LOCAL that is only assigned at
 * iterCount == 0 but read at every iteration while isDegenerate is set (:1961-1990): reading it later is undefined
 * behaviour upstream; here it is kept per function from its last iteration 0 (zero before the first), which is what a
 * reused stack slot gives. iters[0] / iters[1] = iterCount1 / iterCount2 at loop exit (100 = never converged).
 * trace_* (nullable, [100][4]): matX and the row count of every iteration. Returns 1 when the :2506 guard skipped. */
typedef struct orc_fa_state {
  int is_degenerate;
  float matP_surf[9];
  float matP_corner[9];
} orc_fa_state;
void orc_fa_init_state(orc_fa_state* st);
int orc_fa_update_transformation(const orc_p4* corner_sharp, int n_cs, const orc_p4* surf_flat, int n_sf,
                                 const orc_p4* corner_last, int n_cl, const orc_p4* surf_last, int n_sl,
                                 float transform_cur[6], float nearest_feature_dist_sqr, int eigen_mode, int use_ext_knn,
                                 orc_fa_state* st, int iters[2], float* trace_surf, float* trace_corner);

void orc_test_self_adjoint_eig(const float* A, int n, float* evals, float* evecs);
void orc_test_colpiv_qr_solve(const float* A, const float* b, int rows, int cols, float* x);
void orc_test_inverse6(const float* A, float* inv);

#ifdef __cplusplus
}
#endif
#endif
