/* legoloam_b200.h — C ABI of the B200-native scan-to-map registration path of LeGO-LOAM-SR.
 *
 * The reference has no plugin / FFI layer for this path: the "API" is a set of private,
 * argument-less member functions of MapOptimization that communicate through class members
 * (LeGO-LOAM/src/mapOptimization.h:311,327-333; only caller MapOptimization::run,
 * mapOptmization.cpp:1882-1886). This header is the boundary those member functions call after the
 * swap: plain pointers and sizes, no C++ or torch types, every export returns int (0 = ok,
 * negative = ll_status). INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - Host point clouds are read IN PLACE as (base, count, stride_bytes): x,y,z at byte 0/4/8 and
 *    intensity at byte 16 of each record (pcl::PointXYZI, utility.h:45, stride 32). stride 16 means
 *    packed {x,y,z,intensity}. The library never owns or frees host memory.
 *  - `*_dev` variants take DEVICE pointers to packed float4 {x,y,z,intensity}.
 *  - A pose is transformTobeMapped's layout [rx, ry, rz, tx, ty, tz] (mapOptimization.h:248),
 *    R = R_y(ry) R_x(rx) R_z(rz) (mapOptmization.cpp:606-620). A keyframe pose uses the same layout
 *    (PointTypePose roll,pitch,yaw,x,y,z = rx,ry,rz,tx,ty,tz; mapOptmization.cpp:1700-1706).
 *  - A handle is externally synchronised (one caller thread, like the reference's _run_thread,
 *    mapOptmization.cpp:175) and owns one CUDA stream; different handles are independent.
 *  - There is no CPU fallback. If no CUDA device is usable ll_create fails with LL_ERR_CUDA.
 */
#ifndef LEGOLOAM_B200_H_
#define LEGOLOAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  LL_OK = 0,
  LL_ERR_INVALID_ARG = -1,
  LL_ERR_CUDA = -2,
  LL_ERR_NOMEM = -3,
  LL_ERR_NOT_FOUND = -4,      /* unknown keyframe id */
  LL_ERR_CAPACITY = -5,       /* output buffer too small */
  LL_ERR_GRID_TOO_LARGE = -6, /* the 1 m search grid over the local map exceeds the cell budget */
  LL_ERR_STATE = -7           /* call order (e.g. scan2map before a local map exists) */
} ll_status;

typedef struct ll_handle_s* ll_handle;

/* Hot-path knobs. Defaults (ll_default_params) are the reference's values. Thresholds the
 * reference writes as double literals are doubles here so that comparisons promote the same way. */
typedef struct {
  float  corner_leaf;      /* 0.2  downSizeFilterCorner, mapOptmization.cpp:92 */
  float  surf_leaf;        /* 0.4  downSizeFilterSurf,   mapOptmization.cpp:93 */
  float  max_sq_dist;      /* 1.0  pointSearchSqDis[4] < 1.0, :1279 / :1389 (must be <= 1.0) */
  float  eig_ratio;        /* 3    matD1[2] > 3 * matD1[1], :1325 */
  double plane_tol;        /* 0.2  :1416 */
  double weight_k;         /* 0.9  :1363 / :1425 */
  double min_weight;       /* 0.1  :1370 / :1435 */
  int    min_residuals;    /* 50   :1453 */
  float  eig_thresh;       /* 100  :1519 */
  int    iter_count_thres; /* 200  loam_config.yaml:35 */
  float  step_size;        /* 1.0  loam_config.yaml:32 (used at :1500) */
  float  stop_thres;       /* 0.05 loam_config.yaml:33 (used at :1554) */
  int    apply_update;     /* 0 = as shipped (pose update commented out, :1539-1545); 1 = upstream */
  int    eigen_mode;       /* 0 = as shipped (row 0 of V, row-masked V, :1329-1334,:1520-1530);
                              1 = upstream semantics (principal eigenvector, projector sum v v^T) */
  int    map_downsample;   /* 1 = VoxelGrid on the surrounding maps (:1225-1231); 0 = bypass */
  int    min_corner_map;   /* 10   :1573 */
  int    min_surf_map;     /* 100  :1573 */
  int    lm_solver;        /* how matAtA x = matAtB is solved (:1505). 0 (default) = Eigen's ColPivHouseholderQR step
                              sequence in float; 1 = LDL^T in double (same system solved to double accuracy; falls
                              back to 0 when AtA is rank deficient). Pose parity holds for both (tests). */
  int    persistent_lm;    /* 1 (default) = the whole iteration loop of scan2MapOptimization in one cooperative launch
                              (one block per SM, grid barrier per iteration); 0 = one launch per iteration (use this
                              when several handles share a GPU and should interleave, e.g. batched replay). */
  int    lm_max_blocks;    /* 0 (default) = the scan kernel uses every SM. > 0: at most this many blocks (one per SM), so
                              that several handles can keep a persistent scan kernel in flight on one GPU at the same time
                              (batched replay: e.g. half the SMs each); maps too large for that many 8192-point tiles
                              are assembled by the stand-alone kernels instead. */
} ll_params;

/* What scan2MapOptimization leaves behind (mapOptmization.cpp:1602-1607 and members). */
typedef struct {
  int   status;        /* 0 = optimised; 1 = skipped by the map-size guard (:1573), pose unchanged; 2 = NOT registered
                          because the 1 m search grid over the map's bounding box exceeds the cell budget (2^24 cells,
                          flags bit 2) — the reference's KD-tree has no such limit, so callers must treat 2 as a failure
                          of this library, not as the reference's skip (the C++ mirror counts and reports it) */
  int   iter_num;      /* iterations executed (localInfo.x) */
  float min_lambda;    /* smallest eigenvalue of AtA at iteration 0 (localInfo.y) */
  float cf_mean;       /* mean |residual| of the last iteration (localInfo.z) */
  int   is_degenerate; /* persists in the handle across scans, like the member */
  int   converged;     /* LMOptimization returned true */
  int   sel_num;       /* laserCloudSelNum of the last iteration */
  int   flags;         /* bit 0: corner VoxelGrid index overflow (PCL pass-through reproduced);
                          bit 1: same for surf; bit 2: search grid too large (scan skipped) */
  int   corner_map_ds; /* laserCloudCornerFromMapDSNum */
  int   surf_map_ds;   /* laserCloudSurfFromMapDSNum */
  float pose[6];       /* transformTobeMapped after the loop */
  float last_x[6];     /* matX of the last iteration */
} ll_result;

int ll_default_params(ll_params* p);
int ll_create(const ll_params* p, int device_id, ll_handle* out);
int ll_destroy(ll_handle h);
int ll_set_params(ll_handle h, const ll_params* p);
/* Pre-size the device buffers (points of the raw surrounding corner / surf clouds, query points per scan) so
 * that no call on the per-frame path has to reallocate; buffers still grow on demand beyond this. */
int ll_reserve(ll_handle h, size_t max_corner_raw, size_t max_surf_raw, size_t max_queries);
/* Text of the last failure on this handle (CUDA error strings included); never NULL. */
const char* ll_last_error_string(ll_handle h);
/* The handle's CUDA stream (cudaStream_t as void*), for callers that record their own events. */
int ll_get_stream(ll_handle h, void** stream);
/* Number of kernel launches this handle has issued so far (bench.py's gpu_launches). */
int ll_get_launch_count(ll_handle h, uint64_t* n);

/* Optional per-kernel-class timing with CUDA events on the handle's stream (bench.py's roofline). Classes:
 * 0 transform+bbox, 1 voxel keys, 2 voxel sort pass, 3 voxel reduce, 4 grid keys, 5 grid sort pass,
 * 6 grid finalize, 7 LM prologue, 8 LM iteration. ll_profile_read synchronises, returns the summed
 * milliseconds and launch counts since the last read, and clears them. */
int ll_profile_enable(ll_handle h, int on);
int ll_profile_read(ll_handle h, double* ms_per_class, uint64_t* launches_per_class, int n_classes);

/* ---- keyframe store: cornerCloudKeyFrames / surfCloudKeyFrames / outlierCloudKeyFrames
 *      (mapOptimization.h:134-136, filled at mapOptmization.cpp:1744-1752) --------------------- */
int ll_add_keyframe(ll_handle h, int id, const void* corner, size_t n_corner, const void* surf,
                    size_t n_surf, const void* outlier, size_t n_outlier, size_t stride_bytes,
                    const float pose6[6]);
int ll_add_keyframe_dev(ll_handle h, int id, const void* d_corner, size_t n_corner, const void* d_surf,
                        size_t n_surf, const void* d_outlier, size_t n_outlier, const float pose6[6]);
/* Optional overlap of the keyframe upload with the scan in flight: the three clouds saveKeyFramesAndFactor will store
 * (:1744-1752) are inputs of the frame, only their pose comes out of scan2MapOptimization. Called between
 * ll_scan2map_async and ll_scan2map_wait, this packs and uploads them on the handle's copy stream while the scan
 * kernel runs; the following ll_add_keyframe with the same id AND the same three point counts then only attaches the
 * pose (its cloud pointers are not read again); with different counts the prefetch is dropped and the clouds passed
 * to ll_add_keyframe are stored. A second prefetch before the add replaces the first (its block is reclaimed).
 * Without a prefetch ll_add_keyframe uploads as before.
 * _dev variant: the device clouds are copied on the handle's COPY stream, which is not ordered after the stream that
 * produced them — they must be complete when the call is made (true for ll_downsample_scan's results, which
 * synchronises before it returns; a caller with its own producer stream synchronises it first). */
int ll_prefetch_keyframe(ll_handle h, int id, const void* corner, size_t n_corner, const void* surf, size_t n_surf,
                         const void* outlier, size_t n_outlier, size_t stride_bytes);
int ll_prefetch_keyframe_dev(ll_handle h, int id, const void* d_corner, size_t n_corner, const void* d_surf, size_t n_surf,
                             const void* d_outlier, size_t n_outlier);
/* correctPoses (mapOptmization.cpp:1757-1785) */
int ll_set_keyframe_pose(ll_handle h, int id, const float pose6[6]);
/* Memory management of the store. The reference keeps every keyframe cloud for the whole run (:1750-1752); on the
 * device that is ~240 KB per VLP-16 keyframe, 256-byte-aligned blocks in 256 MiB slabs that only grow. A caller that
 * knows a keyframe will not be listed in ll_build_local_map again (e.g. it left the sliding window and loop closure
 * is off) may remove it: its block goes to a free list and is reused by later keyframes (first fit). Re-adding an
 * existing id and dropping an unattached prefetch reclaim their blocks the same way. ll_get_memory_info reports the
 * number of stored keyframes, the arena bytes handed out so far and the bytes waiting in the free list. */
int ll_remove_keyframe(ll_handle h, int id);
int ll_get_memory_info(ll_handle h, size_t* keyframes, size_t* arena_bytes, size_t* free_list_bytes);

/* ---- a1: extractSurroundingKeyFrames tail (mapOptmization.cpp:1147-1151 / :1218-1231):
 * transform the listed keyframes by their poses, concatenate (per keyframe corner | surf, outlier),
 * VoxelGrid 0.2 / 0.4 (unless map_downsample = 0), then build the 1 m search grids that replace
 * kdtreeCornerFromMap / kdtreeSurfFromMap (:1575-1576). Asynchronous on the handle's stream. With persistent_lm = 1 and a
 * map of up to 2^20 raw points the work is only prepared here and runs at the head of the next ll_scan2map launch (one
 * cooperative kernel for the whole scan); any other consumer (ll_get_map_ds, ll_knn5, ...) triggers the stand-alone
 * kernels first, so the observable results are the same either way (bit for bit; tests). */
int ll_build_local_map(ll_handle h, const int* ids, int n_ids);
/* Test / bench entry that skips the keyframe store: raw surrounding clouds already in the map frame. */
int ll_set_local_map_raw(ll_handle h, const void* corner, size_t n_corner, const void* surf,
                         size_t n_surf, size_t stride_bytes);
/* Copy back laserCloud{Corner,Surf}FromMapDS (which = 0 corner, 1 surf). Synchronises. */
int ll_get_map_ds(ll_handle h, int which, void* out, size_t cap_points, size_t stride_bytes, size_t* n);

/* ---- a10: scan2MapOptimization (mapOptmization.cpp:1572-1610) with a7/a8/a9 inside ------------
 * corner_scan_ds = laserCloudCornerScanDS (:1273), surf_total_ds = laserCloudSurfTotalLastDS (:1381).
 * pose6_inout: transformTobeMapped in, refined pose out. Synchronous. */
int ll_scan2map(ll_handle h, const void* corner_scan_ds, size_t n_corner, const void* surf_total_ds,
                size_t n_surf, size_t stride_bytes, float pose6_inout[6], ll_result* res);
int ll_scan2map_dev(ll_handle h, const void* d_corner_scan_ds, size_t n_corner, const void* d_surf_total_ds,
                    size_t n_surf, float pose6_inout[6], ll_result* res);
/* Asynchronous pair: enqueue, then collect. Between the two the caller may drive other handles.
 * Host clouds are normally packed into the library's own staging memory before the call returns. Clouds that lie in
 * PAGE-LOCKED host memory (cudaHostAlloc / cudaHostRegister by the caller) with a stride that is a multiple of 16 bytes
 * (pcl::PointXYZI: 32) are instead copied by DMA straight from the caller's buffers, which must then stay untouched until
 * ll_scan2map_wait returns (ll_scan2map: until it returns). Same results either way (tests). */
int ll_scan2map_async(ll_handle h, const void* corner_scan_ds, size_t n_corner, const void* surf_total_ds,
                      size_t n_surf, size_t stride_bytes, const float pose6_in[6]);
int ll_scan2map_dev_async(ll_handle h, const void* d_corner_scan_ds, size_t n_corner,
                          const void* d_surf_total_ds, size_t n_surf, const float pose6_in[6]);
int ll_scan2map_wait(ll_handle h, float pose6_out[6], ll_result* res);

/* ---- stage-level entry points (parity tests; each synchronous) --------------------------------- */
/* a3: pcl::VoxelGrid<PointXYZI>::filter as called at mapOptmization.cpp:1225-1231 / :1236-1265.
 * member (nullable, n ints): output slot of each input point. flags bit 0 = index overflow. */
int ll_voxel_downsample(ll_handle h, const void* in, size_t n, size_t stride_bytes, float leaf, void* out,
                        size_t cap_points, size_t out_stride_bytes, size_t* n_out, int* member, int* flags);
/* f-2 in one call: downsampleCurrentScan (mapOptmization.cpp:1234-1267). clouds[5] = laserCloudCornerLast,
 * laserCloudSurfLast, laserCloudCornerScan, laserCloudSurfScan, laserCloudOutlierLast (host, stride_bytes per point);
 * the six results are, in order, laserCloudCornerLastDS, SurfLastDS, CornerScanDS, SurfScanDS, OutlierLastDS
 * (params.corner_leaf for the corner clouds, params.surf_leaf for the others, :92-94) and SurfTotalLastDS
 * (= VoxelGrid(SurfLastDS + OutlierLastDS), :1259-1263), bit-identical to six ll_voxel_downsample calls.
 * *dev receives packed float4 DEVICE copies of the inputs and of the results — what ll_scan2map_dev_async and
 * ll_prefetch_keyframe_dev / ll_add_keyframe_dev take — valid until the next ll_downsample_scan on this handle.
 * out may be NULL (results stay on the device) or hold six host buffers (NULL entries skipped) of cap_points[k]
 * points, out_stride_bytes per point. flags[k] = 1: PCL's "leaf size too small" case, cloud k passed through. */
typedef struct ll_scan_clouds {
  const void* d_raw[5];
  size_t      n_raw[5];
  const void* d_ds[6];
  size_t      n_ds[6];
  int         flags[6];
} ll_scan_clouds;
int ll_downsample_scan(ll_handle h, const void* const clouds[5], const size_t n[5], size_t stride_bytes,
                       ll_scan_clouds* dev, void* const out[6], const size_t cap_points[6], size_t out_stride_bytes);
/* a4 + a5: nearestKSearch(point, 5) against the current local map (which = 0 corner, 1 surf) for
 * queries already in the map frame. accept[i] = (5 neighbours found with d2 < max_sq_dist); for
 * accepted queries idx5/d2_5 are the exact 5-NN ascending by (d2, index). */
int ll_knn5(ll_handle h, int which, const void* queries, size_t nq, size_t stride_bytes, int* idx5,
            float* d2_5, int* accept);
/* a6 + a7 / a8 per query, original order: coeff4 = the coeffSel record, mask = kept (s > 0.1). */
int ll_corner_residuals(ll_handle h, const void* queries, size_t nq, size_t stride_bytes, const float pose6[6],
                        float* coeff4, int* mask, int* accept, int* knn_idx5, float* knn_d2_5);
int ll_surf_residuals(ll_handle h, const void* queries, size_t nq, size_t stride_bytes, const float pose6[6],
                      float* coeff4, int* mask, int* accept, int* knn_idx5, float* knn_d2_5);
/* a9: LMOptimization(iter) on explicit residual records (laserCloudOri, coeffSel). ret: 1 converged,
 * 0 not, -1 skipped (n < min_residuals). AtA36 row-major. Uses and updates the handle's persistent
 * isDegenerate / matP. */
int ll_lm_step(ll_handle h, const void* ori, const float* coeff4, size_t n, size_t stride_bytes, int iter,
               float pose6_inout[6], float* AtA36, float* Atb6, float* x6, float* deltaR, float* deltaT,
               int* ret);
/* Persistent degeneracy state (mapOptimization.h:279,281). matP36 row-major. */
int ll_get_degeneracy(ll_handle h, int* is_degenerate, float* matP36);
int ll_reset_state(ll_handle h);

/* ---- second consumer: FeatureAssociation::calculateTransformationSurf / Corner /
 * calculateTransformation (featureAssociation.cpp:1846-2010, :2013-2143, :2145-2289): the same
 * normal-equation reduce + small solve + degeneracy kernel with a column mask. rows = n x 6
 * Jacobian rows (only the masked columns are used, in mask order), b = n right-hand sides.
 * dof_mask: bit k set = unknown k of [rx ry rz tx ty tz]. x_out gets the masked unknowns in order.
 * On iter == 0 the eigen / degeneracy projector is recomputed into matP_inout (m x m, row-major) and
 * *is_degenerate_inout; otherwise they are applied as given. */
int ll_solve_normal_eq(ll_handle h, const float* rows6, const float* b, size_t n, int dof_mask, int iter,
                       float eig_thresh, float* x_out, float* matP_inout, int* is_degenerate_inout);

/* ---- f-3: FeatureAssociation's correspondence search (featureAssociation.cpp:1580-1697 corner, :1699-1844 surf,
 * TransformToStart :1389-1412 inside). which = 0 corner (two tripod points), 1 surf (three).
 * ll_fa_set_last replaces kdtree{Corner,Surf}Last.setInputCloud(laserCloud{Corner,Surf}Last) (:2311-2312, :2715-2716):
 * the cloud (intensity = ring + column / 10000) is uploaded once per sweep.
 * ll_fa_find_correspondences replaces one call of findCorresponding{Corner,Surf}Features(iterCount): queries =
 * cornerPointsSharp / surfPointsFlat; ind1/ind2/ind3 [nq] are pointSearch*Ind1/2/3 (host arrays, rewritten when
 * iter_count % 5 == 0 and read otherwise; ind3 may be NULL for which = 0); coeff4 [nq][4] and keep [nq] describe the
 * (laserCloudOri = queries[i], coeffSel) pair each query pushes (keep[i] = 0: none) — the caller appends the kept ones in
 * order and hands them to ll_solve_normal_eq as today. Synchronous. */
int ll_fa_set_last(ll_handle h, int which, const void* cloud, size_t n, size_t stride_bytes);
int ll_fa_find_correspondences(ll_handle h, int which, const void* queries, size_t nq, size_t stride_bytes,
                               const float transform_cur[6], int iter_count, float nearest_feature_dist_sqr,
                               int* ind1, int* ind2, int* ind3, float* coeff4, int* keep);

/* ---- f-1 + f-3 as the reference runs them: FeatureAssociation::updateTransformation (featureAssociation.cpp:2505-2535),
 * the whole of it in one call and one launch — up to 100 iterations of findCorrespondingSurfFeatures +
 * calculateTransformationSurf (:1846-2010; unknowns rx, rz, ty), then up to 100 of findCorrespondingCornerFeatures +
 * calculateTransformationCorner (:2013-2143; unknowns ry, tx, tz): correspondence search every fifth iteration, Jacobian
 * rows, b = -0.05 d, 3x3 normal equations, eigen threshold 10 at iterCount == 0, NaN reset, the deltaR / deltaT < 0.1 exit.
 * corner_sharp / surf_flat = cornerPointsSharp / surfPointsFlat of the current sweep; the previous sweep's clouds are those
 * of ll_fa_set_last. transform_cur_inout = transformCur. isDegenerate (a class member upstream, :223) and the matP of each
 * function persist in the handle from call to call (ll_fa_reset_state clears them); params.eigen_mode selects the
 * as-shipped or upstream projector as for LMOptimization. trace (nullable, [2][100][4] floats): matX and the row count of
 * every iteration, surf loop first. Synchronous. */
typedef struct ll_fa_result {
  float transform_cur[6];
  int   iters_surf;     /* iterCount1 at loop exit (100: the loop ran out) */
  int   iters_corner;   /* iterCount2 */
  int   is_degenerate;
  int   skipped;        /* 1: the :2506 guard (laserCloudCornerLastNum < 10 || laserCloudSurfLastNum < 100) returned */
  int   n_sel_surf;     /* laserCloudOri size of the last iteration of each loop */
  int   n_sel_corner;
} ll_fa_result;
int ll_fa_update_transformation(ll_handle h, const void* corner_sharp, size_t n_corner, const void* surf_flat, size_t n_surf,
                                size_t stride_bytes, float transform_cur_inout[6], float nearest_feature_dist_sqr,
                                ll_fa_result* res, float* trace);
int ll_fa_reset_state(ll_handle h);

#ifdef __cplusplus
}
#endif
#endif /* LEGOLOAM_B200_H_ */
