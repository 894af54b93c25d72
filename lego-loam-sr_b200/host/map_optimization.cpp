// See map_optimization.hpp. Every member body is marshalling plus calls into the C ABI; the only
// arithmetic kept on the host is the scalar odometry-to-map prior (twelve floats per frame) and the
// key-pose bookkeeping, both of which the reference also does outside its hot loops.
#include "map_optimization.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace lego_loam_b200 {

namespace {

struct M3 {
  double m[3][3];
};

// R = R_y(ry) R_x(rx) R_z(rz): the rotation pointAssociateToMap applies (mapOptmization.cpp:606-620).
M3 rot(const float* p) {
  const double cx = std::cos((double)p[0]), sx = std::sin((double)p[0]);
  const double cy = std::cos((double)p[1]), sy = std::sin((double)p[1]);
  const double cz = std::cos((double)p[2]), sz = std::sin((double)p[2]);
  M3 R = {{{cy * cz + sy * sx * sz, -cy * sz + sy * sx * cz, sy * cx},
           {cx * sz, cx * cz, -sx},
           {-sy * cz + cy * sx * sz, sy * sz + cy * sx * cz, cy * cx}}};
  return R;
}
M3 mul(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return c;
}
M3 transpose(const M3& a) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[j][i];
  return c;
}

}  // namespace

// transformAssociateToMap (mapOptmization.cpp:458-581). The reference expands the products of sines
// and cosines by hand; what it computes is
//     R_tobe = R_aft * R_bef^T * R_sum,      t_tobe = t_aft - R_tobe * (R_sum^T (t_bef - t_sum)),
// evaluated here with 3x3 matrices in double and rounded to float once.
void associateToMap(const float sum[6], const float bef[6], const float aft[6], float incre[6], float tobe[6]) {
  const M3 Rs = rot(sum), Rb = rot(bef), Ra = rot(aft);
  const double d[3] = {(double)bef[3] - sum[3], (double)bef[4] - sum[4], (double)bef[5] - sum[5]};
  double inc[3];
  for (int i = 0; i < 3; ++i) inc[i] = Rs.m[0][i] * d[0] + Rs.m[1][i] * d[1] + Rs.m[2][i] * d[2];
  incre[3] = (float)inc[0];
  incre[4] = (float)inc[1];
  incre[5] = (float)inc[2];
  const M3 R = mul(Ra, mul(transpose(Rb), Rs));
  const double rx = std::asin(std::max(-1.0, std::min(1.0, -R.m[1][2])));
  const double ry = std::atan2(R.m[0][2], R.m[2][2]);
  const double rz = std::atan2(R.m[1][0], R.m[1][1]);
  tobe[0] = (float)rx;
  tobe[1] = (float)ry;
  tobe[2] = (float)rz;
  const float ang[3] = {tobe[0], tobe[1], tobe[2]};
  const M3 Rt = rot(ang);
  for (int i = 0; i < 3; ++i)
    tobe[3 + i] = (float)((double)aft[3 + i] - (Rt.m[i][0] * inc[0] + Rt.m[i][1] * inc[1] + Rt.m[i][2] * inc[2]));
}

MapOptimization::MapOptimization(const ll_params& params, int device_id, bool loop_closure_enabled,
                                 int surrounding_keyframe_search_num, float surrounding_keyframe_search_radius,
                                 size_t reserve_corner_raw, size_t reserve_surf_raw, size_t reserve_queries)
    : prm_(params),
      loop_closure_enabled_(loop_closure_enabled),
      search_num_(surrounding_keyframe_search_num),
      search_radius_(surrounding_keyframe_search_radius) {
  for (int i = 0; i < 6; ++i)  // mapOptmization.cpp:267-274
    transformLast[i] = transformSum[i] = transformIncre[i] = transformTobeMapped[i] = transformBefMapped[i] =
        transformAftMapped[i] = 0;
  auto mk = [] { return std::make_shared<PointCloud>(); };
  laserCloudCornerLast = mk(); laserCloudSurfLast = mk(); laserCloudCornerScan = mk(); laserCloudSurfScan = mk();
  laserCloudOutlierLast = mk(); laserCloudCornerLastDS = mk(); laserCloudSurfLastDS = mk();
  laserCloudCornerScanDS = mk(); laserCloudSurfScanDS = mk(); laserCloudOutlierLastDS = mk();
  laserCloudSurfTotalLast = mk(); laserCloudSurfTotalLastDS = mk(); laserCloudOri = mk(); coeffSel = mk();
  std::memset(&lastResult, 0, sizeof(lastResult));
  const int rc = ll_create(&prm_, device_id, &h_);
  if (rc != LL_OK) throw std::runtime_error("ll_create failed (" + std::to_string(rc) + "): no usable CUDA device; there is no CPU fallback");
  // A 50-keyframe VLP-16 window is ~0.15 M corner + ~0.45 M surf raw points; the defaults reserve well above that once.
  check(ll_reserve(h_, reserve_corner_raw, reserve_surf_raw, reserve_queries), "ll_reserve");
}

MapOptimization::~MapOptimization() { ll_destroy(h_); }

void MapOptimization::check(int rc, const char* what) {
  if (rc == LL_OK) return;
  err_ = std::string(what) + ": " + ll_last_error_string(h_);
  throw std::runtime_error(err_);
}

void MapOptimization::transformAssociateToMap() {
  associateToMap(transformSum, transformBefMapped, transformAftMapped, transformIncre, transformTobeMapped);
}

void MapOptimization::transformUpdate() {  // :583-589
  for (int i = 0; i < 6; i++) {
    transformBefMapped[i] = transformSum[i];
    transformAftMapped[i] = transformTobeMapped[i];
  }
}

PointCloud::Ptr MapOptimization::voxel(const PointCloud& in, float leaf) {
  auto out = std::make_shared<PointCloud>();
  if (in.empty()) return out;
  out->points.resize(in.size());
  size_t n_out = 0;
  int flags = 0;
  check(ll_voxel_downsample(h_, in.points.data(), in.size(), sizeof(PointType), leaf, out->points.data(),
                            out->points.size(), sizeof(PointType), &n_out, nullptr, &flags),
        "ll_voxel_downsample");
  out->points.resize(n_out);
  return out;
}

void MapOptimization::extractSurroundingKeyFrames() {  // :1096-1232
  mapBuilt_ = false;
  if (cloudKeyPoses3D.empty()) return;
  mapIds_.clear();
  if (loop_closure_enabled_) {  // :1099-1151: the last `search_num_` keyframes
    const int numPoses = (int)cloudKeyPoses3D.size();
    if ((int)recentKeyFrameIDs_.size() < search_num_) {
      recentKeyFrameIDs_.clear();
      for (int i = numPoses - 1; i >= 0; --i) {
        recentKeyFrameIDs_.push_front((int)cloudKeyPoses3D[i].intensity);
        if ((int)recentKeyFrameIDs_.size() >= search_num_) break;
      }
    } else if (latestFrameID_ != numPoses - 1) {
      recentKeyFrameIDs_.pop_front();
      latestFrameID_ = numPoses - 1;
      recentKeyFrameIDs_.push_back(latestFrameID_);
    }
    mapIds_.assign(recentKeyFrameIDs_.begin(), recentKeyFrameIDs_.end());
  } else {  // :1152-1223: key poses within the search radius, thinned by a 1 m VoxelGrid
    struct Hit { float d2; int idx; };
    std::vector<Hit> hits;
    const float r2 = search_radius_ * search_radius_;
    for (int i = 0; i < (int)cloudKeyPoses3D.size(); ++i) {
      float dx = currentRobotPosPoint.x - cloudKeyPoses3D[i].x, dy = currentRobotPosPoint.y - cloudKeyPoses3D[i].y,
            dz = currentRobotPosPoint.z - cloudKeyPoses3D[i].z;
      float d2 = dx * dx;
      d2 += dy * dy;
      d2 += dz * dz;
      if (d2 < r2) hits.push_back({d2, i});
    }
    std::stable_sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) { return a.d2 < b.d2; });
    PointCloud surrounding;
    for (const Hit& hh : hits) surrounding.push_back(cloudKeyPoses3D[hh.idx]);
    PointCloud::Ptr ds = voxel(surrounding, 1.0f);  // downSizeFilterSurroundingKeyPoses :99
    // drop cached keyframes that left the region, then append new ones (:1169-1216); the keyframe id is
    // recovered as (int) of the AVERAGED intensity, exactly like the reference (:1175, :1197, :1205)
    for (size_t i = 0; i < surroundingExistingKeyPosesID_.size(); ++i) {
      bool existing = false;
      for (const PointType& p : ds->points)
        if (surroundingExistingKeyPosesID_[i] == (int)p.intensity) { existing = true; break; }
      if (!existing) {
        surroundingExistingKeyPosesID_.erase(surroundingExistingKeyPosesID_.begin() + i);
        --i;
      }
    }
    for (const PointType& p : ds->points) {
      const int id = (int)p.intensity;
      if (std::find(surroundingExistingKeyPosesID_.begin(), surroundingExistingKeyPosesID_.end(), id) ==
          surroundingExistingKeyPosesID_.end())
        surroundingExistingKeyPosesID_.push_back(id);
    }
    mapIds_ = surroundingExistingKeyPosesID_;
  }
  // transformPointCloud of every listed keyframe + concatenation + VoxelGrid 0.2 / 0.4 (:1218-1231) and the
  // search structures of :1575-1576, all on the device
  check(ll_build_local_map(h_, mapIds_.data(), (int)mapIds_.size()), "ll_build_local_map");
  mapBuilt_ = true;
}

void MapOptimization::downsampleCurrentScan() {  // :1234-1267
  // The six filters in one call (ll_downsample_scan): the raw clouds go up once, the results come back as the
  // reference's host members AND stay on the device for scan2MapOptimization / saveKeyFramesAndFactor.
  const PointCloud* in[5] = {laserCloudCornerLast.get(), laserCloudSurfLast.get(), laserCloudCornerScan.get(),
                             laserCloudSurfScan.get(), laserCloudOutlierLast.get()};
  PointCloud::Ptr* res[6] = {&laserCloudCornerLastDS, &laserCloudSurfLastDS, &laserCloudCornerScanDS,
                             &laserCloudSurfScanDS, &laserCloudOutlierLastDS, &laserCloudSurfTotalLastDS};
  const void* clouds[5];
  size_t n[5], cap[6];
  void* out[6];
  for (int k = 0; k < 5; ++k) {
    clouds[k] = in[k]->points.data();
    n[k] = in[k]->size();
  }
  for (int k = 0; k < 6; ++k) {  // landing buffers sized for the worst case (output = input); they only ever grow
    cap[k] = k < 5 ? n[k] : n[1] + n[4];
    if (scanLanding_[k].size() < cap[k]) scanLanding_[k].resize(cap[k] + cap[k] / 2);
    out[k] = scanLanding_[k].data();
  }
  check(ll_downsample_scan(h_, clouds, n, sizeof(PointType), &scanDev_, out, cap, sizeof(PointType)), "ll_downsample_scan");
  for (int k = 0; k < 6; ++k) {  // like filter(*cloudDS): the member clouds are reused, only the n_ds points are written
    if (!*res[k] || res[k]->use_count() != 1) *res[k] = std::make_shared<PointCloud>();  // (never write into a caller's cloud)
    (*res[k])->points.assign(scanLanding_[k].begin(), scanLanding_[k].begin() + (long)scanDev_.n_ds[k]);
  }
  scanDevValid_ = true;
  laserCloudSurfTotalLast->clear();  // :1259-1260 (its downsampled form came out of the call above)
  *laserCloudSurfTotalLast += *laserCloudSurfLastDS;
  *laserCloudSurfTotalLast += *laserCloudOutlierLastDS;
}

static void fill_residual_clouds(const PointCloud& q, const std::vector<float>& coeff, const std::vector<int>& mask,
                                 PointCloud& ori, PointCloud& sel) {
  for (size_t i = 0; i < q.size(); ++i) {
    if (!mask[i]) continue;
    ori.push_back(q.points[i]);
    PointType c;
    c.x = coeff[4 * i];
    c.y = coeff[4 * i + 1];
    c.z = coeff[4 * i + 2];
    c.intensity = coeff[4 * i + 3];
    sel.push_back(c);
  }
}

void MapOptimization::cornerOptimization() {  // :1269-1377
  const PointCloud& q = *laserCloudCornerScanDS;
  std::vector<float> coeff(4 * q.size() + 4);
  std::vector<int> mask(q.size() + 1);
  check(ll_corner_residuals(h_, q.points.data(), q.size(), sizeof(PointType), transformTobeMapped, coeff.data(),
                            mask.data(), nullptr, nullptr, nullptr),
        "ll_corner_residuals");
  fill_residual_clouds(q, coeff, mask, *laserCloudOri, *coeffSel);
}

void MapOptimization::surfOptimization() {  // :1379-1442
  const PointCloud& q = *laserCloudSurfTotalLastDS;
  std::vector<float> coeff(4 * q.size() + 4);
  std::vector<int> mask(q.size() + 1);
  check(ll_surf_residuals(h_, q.points.data(), q.size(), sizeof(PointType), transformTobeMapped, coeff.data(),
                          mask.data(), nullptr, nullptr, nullptr),
        "ll_surf_residuals");
  fill_residual_clouds(q, coeff, mask, *laserCloudOri, *coeffSel);
}

bool MapOptimization::LMOptimization(int iterCount) {  // :1444-1570
  std::vector<float> coeff(4 * coeffSel->size() + 4);
  for (size_t i = 0; i < coeffSel->size(); ++i) {
    coeff[4 * i] = coeffSel->points[i].x;
    coeff[4 * i + 1] = coeffSel->points[i].y;
    coeff[4 * i + 2] = coeffSel->points[i].z;
    coeff[4 * i + 3] = coeffSel->points[i].intensity;
  }
  int ret = 0;
  check(ll_lm_step(h_, laserCloudOri->points.data(), coeff.data(), laserCloudOri->size(), sizeof(PointType), iterCount,
                   transformTobeMapped, nullptr, nullptr, nullptr, nullptr, nullptr, &ret),
        "ll_lm_step");
  int deg = 0;
  check(ll_get_degeneracy(h_, &deg, nullptr), "ll_get_degeneracy");
  isDegenerate = deg != 0;
  return ret == 1;
}

void MapOptimization::scan2MapOptimization() {  // :1572-1610
  scan2MapLaunch();
  scan2MapCollect();
}

void MapOptimization::scan2MapLaunch() {
  scanInFlight_ = false;
  std::memset(&lastResult, 0, sizeof(lastResult));
  lastResult.status = 1;
  std::memcpy(lastResult.pose, transformTobeMapped, sizeof(lastResult.pose));
  std::memcpy(lastPrior, transformTobeMapped, sizeof(lastPrior));
  if (!mapBuilt_) return;  // first frame: laserCloud*FromMapDSNum are 0, the guard at :1573 fails
  if (view_ && view_->device) {
    check(ll_scan2map_dev_async(h_, view_->corner_scan_ds, view_->n_corner_scan_ds, view_->surf_total_ds,
                                view_->n_surf_total_ds, transformTobeMapped),
          "ll_scan2map_dev_async");
    check(ll_prefetch_keyframe_dev(h_, (int)cloudKeyPoses3D.size(), view_->corner_scan, view_->n_corner_scan,
                                   view_->surf_last_ds, view_->n_surf_last_ds, view_->outlier_last_ds,
                                   view_->n_outlier_last_ds),
          "ll_prefetch_keyframe_dev");
  } else if (view_) {
    // While the scan runs on the GPU, pack and upload the three clouds saveKeyFramesAndFactor is about to store
    // (every frame becomes a keyframe, :1622-1631; only its pose comes out of the scan).
    check(ll_scan2map_async(h_, view_->corner_scan_ds, view_->n_corner_scan_ds, view_->surf_total_ds,
                            view_->n_surf_total_ds, view_->stride_bytes, transformTobeMapped),
          "ll_scan2map_async");
    check(ll_prefetch_keyframe(h_, (int)cloudKeyPoses3D.size(), view_->corner_scan, view_->n_corner_scan,
                               view_->surf_last_ds, view_->n_surf_last_ds, view_->outlier_last_ds,
                               view_->n_outlier_last_ds, view_->stride_bytes),
          "ll_prefetch_keyframe");
  } else if (scanDevValid_) {  // downsampleCurrentScan left everything the scan and the keyframe need on the device
    const ll_scan_clouds& d = scanDev_;
    check(ll_scan2map_dev_async(h_, d.d_ds[2], d.n_ds[2], d.d_ds[5], d.n_ds[5], transformTobeMapped), "ll_scan2map_dev_async");
    check(ll_prefetch_keyframe_dev(h_, (int)cloudKeyPoses3D.size(), d.d_raw[2], d.n_raw[2], d.d_ds[1], d.n_ds[1], d.d_ds[4],
                                   d.n_ds[4]),
          "ll_prefetch_keyframe_dev");
  } else {
    check(ll_scan2map_async(h_, laserCloudCornerScanDS->points.data(), laserCloudCornerScanDS->size(),
                            laserCloudSurfTotalLastDS->points.data(), laserCloudSurfTotalLastDS->size(),
                            sizeof(PointType), transformTobeMapped),
          "ll_scan2map_async");
    check(ll_prefetch_keyframe(h_, (int)cloudKeyPoses3D.size(), laserCloudCornerScan->points.data(),
                               laserCloudCornerScan->size(), laserCloudSurfLastDS->points.data(),
                               laserCloudSurfLastDS->size(), laserCloudOutlierLastDS->points.data(),
                               laserCloudOutlierLastDS->size(), sizeof(PointType)),
          "ll_prefetch_keyframe");
  }
  scanInFlight_ = true;
}

void MapOptimization::scan2MapCollect() {
  if (!scanInFlight_) return;  // (first frame: no map yet, nothing was launched)
  scanInFlight_ = false;
  check(ll_scan2map_wait(h_, transformTobeMapped, &lastResult), "ll_scan2map_wait");
  laserCloudCornerFromMapDSNum = lastResult.corner_map_ds;
  laserCloudSurfFromMapDSNum = lastResult.surf_map_ds;
  if (lastResult.status == 2) {  // not the reference's skip: the search grid did not fit (see gridOverflowScans)
    if (gridOverflowScans++ == 0)
      std::fprintf(stderr, "legoloam_b200: frame %d NOT registered: the 1 m search grid over the local map exceeds the cell "
                           "budget (map bounding box too large); pose left at the odometry prior\n", frame_idx);
    err_ = "scan2MapOptimization: search grid too large, scan not registered";
  }
  if (lastResult.status != 0) return;
  iter_num = lastResult.iter_num;
  min_lambda = lastResult.min_lambda;
  CF_mean = lastResult.cf_mean;
  isDegenerate = lastResult.is_degenerate != 0;
  PointType info;  // :1602-1607
  info.x = (float)iter_num;
  info.y = min_lambda;
  info.z = CF_mean;
  info.intensity = (float)frame_idx;
  localInfo.push_back(info);
  transformUpdate();
}

void MapOptimization::saveKeyFramesAndFactor() {  // :1612-1755, pose graph = pass-through
  currentRobotPosPoint.x = transformAftMapped[3];
  currentRobotPosPoint.y = transformAftMapped[4];
  currentRobotPosPoint.z = transformAftMapped[5];
  previousRobotPosPoint = currentRobotPosPoint;  // every frame is a keyframe (:1622-1631)
  const bool first = cloudKeyPoses3D.empty();
  float pose[6];
  if (first) {  // PriorFactor at transformTobeMapped (:1639-1652)
    for (int i = 0; i < 6; ++i) pose[i] = transformLast[i] = transformTobeMapped[i];
  } else {      // BetweenFactor transformLast -> transformAftMapped; the estimate is its end point
    for (int i = 0; i < 6; ++i) pose[i] = transformAftMapped[i];
  }
  PointType p3;
  p3.x = pose[3]; p3.y = pose[4]; p3.z = pose[5];
  p3.intensity = (float)cloudKeyPoses3D.size();  // "this can be used as index"
  cloudKeyPoses3D.push_back(p3);
  PointTypePose p6;
  p6.x = p3.x; p6.y = p3.y; p6.z = p3.z; p6.intensity = p3.intensity;
  p6.roll = pose[0]; p6.pitch = pose[1]; p6.yaw = pose[2];
  cloudKeyPoses6D.push_back(p6);
  if (cloudKeyPoses3D.size() > 1) {  // :1724-1736
    for (int i = 0; i < 6; ++i) {
      transformAftMapped[i] = pose[i];
      transformLast[i] = transformAftMapped[i];
      transformTobeMapped[i] = transformAftMapped[i];
    }
  }
  // :1744-1752: corner keyframe = un-downsampled laserCloudCornerScan, surf = laserCloudSurfLastDS,
  // outlier = laserCloudOutlierLastDS. Uploaded once; the device keeps them for the whole run.
  if (view_ && view_->device) {
    check(ll_add_keyframe_dev(h_, (int)cloudKeyPoses3D.size() - 1, view_->corner_scan, view_->n_corner_scan,
                              view_->surf_last_ds, view_->n_surf_last_ds, view_->outlier_last_ds,
                              view_->n_outlier_last_ds, pose),
          "ll_add_keyframe_dev");
    return;
  }
  if (view_) {
    check(ll_add_keyframe(h_, (int)cloudKeyPoses3D.size() - 1, view_->corner_scan, view_->n_corner_scan,
                          view_->surf_last_ds, view_->n_surf_last_ds, view_->outlier_last_ds, view_->n_outlier_last_ds,
                          view_->stride_bytes, pose),
          "ll_add_keyframe");
    return;
  }
  if (scanDevValid_) {
    const ll_scan_clouds& d = scanDev_;
    check(ll_add_keyframe_dev(h_, (int)cloudKeyPoses3D.size() - 1, d.d_raw[2], d.n_raw[2], d.d_ds[1], d.n_ds[1], d.d_ds[4],
                              d.n_ds[4], pose),
          "ll_add_keyframe_dev");
    return;
  }
  check(ll_add_keyframe(h_, (int)cloudKeyPoses3D.size() - 1, laserCloudCornerScan->points.data(),
                        laserCloudCornerScan->size(), laserCloudSurfLastDS->points.data(), laserCloudSurfLastDS->size(),
                        laserCloudOutlierLastDS->points.data(), laserCloudOutlierLastDS->size(), sizeof(PointType), pose),
        "ll_add_keyframe");
}

void MapOptimization::processFrameView(const float laser_odometry[6], const FrameView& frame) {
  processFrameViewBegin(laser_odometry, frame);
  processFrameViewEnd();
}

void MapOptimization::processFrameViewBegin(const float laser_odometry[6], const FrameView& frame) {
  frameStart_ = std::chrono::steady_clock::now();
  frame_idx++;
  scanDevValid_ = false;
  view_ = &frame;
  try {
    for (int i = 0; i < 6; ++i) transformSum[i] = laser_odometry[i];
    transformAssociateToMap();
    extractSurroundingKeyFrames();
    scan2MapLaunch();
  } catch (...) {
    view_ = nullptr;
    throw;
  }
}

void MapOptimization::processFrameViewEnd() {
  try {
    scan2MapCollect();
    saveKeyFramesAndFactor();
    correctPoses();
    clearCloud();
  } catch (...) {
    view_ = nullptr;
    throw;
  }
  view_ = nullptr;
  lastFrameMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - frameStart_).count();
}

void MapOptimization::correctPoses() {  // :1757-1785
  if (!aLoopIsClosed) return;
  recentKeyFrameIDs_.clear();
  for (size_t i = 0; i < cloudKeyPoses6D.size(); ++i) {
    const PointTypePose& p = cloudKeyPoses6D[i];
    const float pose[6] = {p.roll, p.pitch, p.yaw, p.x, p.y, p.z};
    check(ll_set_keyframe_pose(h_, (int)i, pose), "ll_set_keyframe_pose");
  }
  aLoopIsClosed = false;
}

void MapOptimization::clearCloud() {}  // the device maps are overwritten by the next build (:1787-1792)

PointCloud::Ptr MapOptimization::getLaserCloudCornerFromMapDS() {
  auto out = std::make_shared<PointCloud>();
  size_t n = 0;
  check(ll_get_map_ds(h_, 0, nullptr, 0, sizeof(PointType), &n), "ll_get_map_ds");
  out->points.resize(n);
  if (n) check(ll_get_map_ds(h_, 0, out->points.data(), n, sizeof(PointType), &n), "ll_get_map_ds");
  return out;
}

PointCloud::Ptr MapOptimization::getLaserCloudSurfFromMapDS() {
  auto out = std::make_shared<PointCloud>();
  size_t n = 0;
  check(ll_get_map_ds(h_, 1, nullptr, 0, sizeof(PointType), &n), "ll_get_map_ds");
  out->points.resize(n);
  if (n) check(ll_get_map_ds(h_, 1, out->points.data(), n, sizeof(PointType), &n), "ll_get_map_ds");
  return out;
}

void MapOptimization::processFrame(const AssociationOut& a, bool scan_clouds_are_downsampled,
                                   PointCloud::Ptr corner_scan_ds) {  // run(): :1866-1896
  const auto t_frame = std::chrono::steady_clock::now();
  frame_idx++;
  laserCloudCornerLast = a.cloud_corner_last;
  laserCloudSurfLast = a.cloud_surf_last;
  laserCloudCornerScan = a.cloud_corner_scan;
  laserCloudSurfScan = a.cloud_surf_scan;
  laserCloudOutlierLast = a.cloud_outlier_last;
  for (int i = 0; i < 6; ++i) transformSum[i] = a.laser_odometry[i];
  transformAssociateToMap();
  extractSurroundingKeyFrames();
  scanDevValid_ = false;
  if (scan_clouds_are_downsampled) {
    laserCloudCornerScanDS = corner_scan_ds ? corner_scan_ds : a.cloud_corner_scan;
    laserCloudSurfLastDS = a.cloud_surf_last;
    laserCloudOutlierLastDS = a.cloud_outlier_last;
    laserCloudSurfTotalLastDS = a.cloud_surf_scan;
  } else {
    downsampleCurrentScan();
  }
  scan2MapOptimization();
  saveKeyFramesAndFactor();
  correctPoses();
  clearCloud();
  lastFrameMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_frame).count();
}

}  // namespace lego_loam_b200

// ---- C exports so that Python tests / bench.py can drive the C++ mirror --------------------------------
using lego_loam_b200::AssociationOut;
using lego_loam_b200::MapOptimization;
using lego_loam_b200::PointCloud;
using lego_loam_b200::PointType;

namespace lego_loam_b200 {
void associateToMap(const float sum[6], const float bef[6], const float aft[6], float incre[6], float tobe[6]);
}

static PointCloud::Ptr cloud_from_packed(const float* p4, size_t n) {
  auto c = std::make_shared<PointCloud>();
  c->points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    c->points[i].x = p4[4 * i];
    c->points[i].y = p4[4 * i + 1];
    c->points[i].z = p4[4 * i + 2];
    c->points[i].intensity = p4[4 * i + 3];
  }
  return c;
}

extern "C" {

void llh_transform_associate_to_map(const float sum[6], const float bef[6], const float aft[6], float tobe[6]) {
  float incre[6];
  lego_loam_b200::associateToMap(sum, bef, aft, incre, tobe);
}

void* llh_create(const ll_params* prm, int device, int loop_closure_enabled, int search_num, float search_radius,
                 char* err, size_t err_cap) {
  try {
    return new MapOptimization(*prm, device, loop_closure_enabled != 0, search_num, search_radius);
  } catch (const std::exception& e) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", e.what());
    return nullptr;
  }
}

void* llh_create2(const ll_params* prm, int device, int loop_closure_enabled, int search_num, float search_radius,
                  size_t reserve_corner_raw, size_t reserve_surf_raw, size_t reserve_queries, char* err, size_t err_cap) {
  try {
    return new MapOptimization(*prm, device, loop_closure_enabled != 0, search_num, search_radius, reserve_corner_raw,
                               reserve_surf_raw, reserve_queries);
  } catch (const std::exception& e) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", e.what());
    return nullptr;
  }
}

void llh_destroy(void* mo) { delete static_cast<MapOptimization*>(mo); }

// One frame from packed float4 clouds. With pre_ds != 0 the clouds are: corner_scan (raw), corner_scan_ds,
// surf_last_ds, outlier_last_ds, surf_total_ds; otherwise corner_scan, (unused), surf_last, outlier_last, (unused)
// are raw and downsampleCurrentScan runs on the device.
int llh_process_frame(void* mo, const float odom6[6], const float* corner_scan, size_t n_cs, const float* corner_scan_ds,
                      size_t n_csd, const float* surf_last, size_t n_sl, const float* outlier_last, size_t n_ol,
                      const float* surf_total_ds, size_t n_st, int pre_ds, ll_result* res, float pose_out[6],
                      char* err, size_t err_cap) {
  MapOptimization* M = static_cast<MapOptimization*>(mo);
  try {
    AssociationOut a;
    a.cloud_corner_scan = cloud_from_packed(corner_scan, n_cs);
    a.cloud_corner_last = pre_ds ? a.cloud_corner_scan : a.cloud_corner_scan;
    a.cloud_surf_last = cloud_from_packed(surf_last, n_sl);
    a.cloud_outlier_last = cloud_from_packed(outlier_last, n_ol);
    a.cloud_surf_scan = pre_ds ? cloud_from_packed(surf_total_ds, n_st) : a.cloud_surf_last;
    std::memcpy(a.laser_odometry, odom6, sizeof(a.laser_odometry));
    M->processFrame(a, pre_ds != 0, pre_ds ? cloud_from_packed(corner_scan_ds, n_csd) : nullptr);
    if (res) *res = M->lastResult;
    if (pose_out) std::memcpy(pose_out, M->lastResult.pose, 6 * sizeof(float));
    return 0;
  } catch (const std::exception& e) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", e.what());
    return -1;
  }
}

// One frame on caller-owned, already downsampled clouds (no copies into PointCloud objects): device != 0 —
// packed float4 clouds resident on the device; device == 0 — host clouds read in place with `stride` bytes per point.
int llh_process_frame_view(void* mo, const float odom6[6], const void* corner_scan, size_t n_cs,
                           const void* corner_scan_ds, size_t n_csd, const void* surf_last_ds, size_t n_sl,
                           const void* outlier_last_ds, size_t n_ol, const void* surf_total_ds, size_t n_st, int device,
                           size_t stride, ll_result* res, float pose_out[6], char* err, size_t err_cap) {
  MapOptimization* M = static_cast<MapOptimization*>(mo);
  try {
    MapOptimization::FrameView f;
    f.corner_scan = corner_scan; f.n_corner_scan = n_cs;
    f.corner_scan_ds = corner_scan_ds; f.n_corner_scan_ds = n_csd;
    f.surf_last_ds = surf_last_ds; f.n_surf_last_ds = n_sl;
    f.outlier_last_ds = outlier_last_ds; f.n_outlier_last_ds = n_ol;
    f.surf_total_ds = surf_total_ds; f.n_surf_total_ds = n_st;
    f.device = device != 0;
    f.stride_bytes = stride;
    M->processFrameView(odom6, f);
    if (res) *res = M->lastResult;
    if (pose_out) std::memcpy(pose_out, M->lastResult.pose, 6 * sizeof(float));
    return 0;
  } catch (const std::exception& e) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", e.what());
    return -1;
  }
}

int llh_num_keyframes(void* mo) { return (int)static_cast<MapOptimization*>(mo)->cloudKeyPoses3D.size(); }

void llh_get_key_pose(void* mo, int i, float pose6[6]) {
  const auto& p = static_cast<MapOptimization*>(mo)->cloudKeyPoses6D[(size_t)i];
  pose6[0] = p.roll; pose6[1] = p.pitch; pose6[2] = p.yaw; pose6[3] = p.x; pose6[4] = p.y; pose6[5] = p.z;
}

// transformTobeMapped as it entered the last scan2MapOptimization (the odometry-to-map prior)
void llh_get_last_prior(void* mo, float pose6[6]) {
  std::memcpy(pose6, static_cast<MapOptimization*>(mo)->lastPrior, 6 * sizeof(float));
}

void* llh_abi_handle(void* mo) { return static_cast<MapOptimization*>(mo)->handle(); }

// keyframe ids selected by the last extractSurroundingKeyFrames, in concatenation order; returns their number
int llh_get_map_ids(void* mo, int* ids, int cap) {
  const auto& v = static_cast<MapOptimization*>(mo)->mapIds();
  for (int i = 0; i < (int)v.size() && i < cap; ++i) ids[i] = v[(size_t)i];
  return (int)v.size();
}

int llh_grid_overflow_scans(void* mo) { return static_cast<MapOptimization*>(mo)->gridOverflowScans; }

double llh_last_frame_ms(void* mo) { return static_cast<MapOptimization*>(mo)->lastFrameMs; }

}  // extern "C"
