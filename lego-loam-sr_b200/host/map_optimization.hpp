// C++ host mirror of the reference's MapOptimization hot-path members, written over the C ABI only
// (include/legoloam_b200.h). Member names, call order and data flow follow
// LeGO-LOAM/src/mapOptimization.h:311,327-333 and MapOptimization::run (mapOptmization.cpp:1854-1930)
// so that the ROS 2 node can replace the bodies one for one. ROS, PCL, GTSAM and
// Eigen do not appear: clouds are std::vector of a 32-byte PointXYZI with pcl's field offsets, and the
// pose graph is a pass-through (loop closure disabled => iSAM2 returns the odometry chain).
#pragma once
#include <cstddef>
#include <cstdint>
#include <chrono>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "legoloam_b200.h"

namespace lego_loam_b200 {

// pcl::PointXYZI (utility.h:45): x,y,z at bytes 0/4/8, intensity at byte 16, 32-byte stride.
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, data3 = 1.0f;
  float intensity = 0, pad[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZI) == 32, "PointXYZI must match pcl::PointXYZI's 32-byte layout");
using PointType = PointXYZI;

struct PointCloud {
  std::vector<PointType> points;
  using Ptr = std::shared_ptr<PointCloud>;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void push_back(const PointType& p) { points.push_back(p); }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};

// PointXYZIRPYT (utility.h:118-134)
struct PointTypePose {
  float x = 0, y = 0, z = 0, intensity = 0, roll = 0, pitch = 0, yaw = 0;
  double time = 0;
};

// AssociationOut (utility.h:75-83) with the odometry already converted by OdometryToTransform (utility.h:99-113).
struct AssociationOut {
  PointCloud::Ptr cloud_outlier_last, cloud_corner_last, cloud_surf_last, cloud_corner_scan, cloud_surf_scan;
  float laser_odometry[6] = {0, 0, 0, 0, 0, 0};  // transformSum
  double stamp = 0;
};

class MapOptimization {
 public:
  // loop_closure_enabled / surrounding_keyframe_search_* : loam_config.yaml:23,26,27
  // reserve_*: points of the raw surrounding corner / surf clouds and queries per scan the device buffers are
  // pre-sized for (ll_reserve; they still grow on demand). The defaults hold a 50-keyframe VLP-16 window several
  // times over (~0.5 GB per handle, cell tables included); batched replay with many handles per GPU passes less.
  MapOptimization(const ll_params& params, int device_id, bool loop_closure_enabled = true,
                  int surrounding_keyframe_search_num = 50, float surrounding_keyframe_search_radius = 50.0f,
                  size_t reserve_corner_raw = 1u << 20, size_t reserve_surf_raw = 1u << 21, size_t reserve_queries = 1u << 17);
  ~MapOptimization();
  MapOptimization(const MapOptimization&) = delete;
  MapOptimization& operator=(const MapOptimization&) = delete;

  // ---- the per-frame body of run() (mapOptmization.cpp:1866-1896) -----------------------------------
  // scan_clouds_are_downsampled: the five clouds are already what downsampleCurrentScan would produce
  // (cloud_corner_scan stays raw: it is the corner keyframe cloud, :1746; `corner_scan_ds` then carries
  // laserCloudCornerScanDS and cloud_surf_last / cloud_outlier_last carry *LastDS, cloud_surf_scan carries
  // laserCloudSurfTotalLastDS).
  void processFrame(const AssociationOut& association, bool scan_clouds_are_downsampled = false,
                    PointCloud::Ptr corner_scan_ds = nullptr);

  // Same frame body on caller-owned memory, without copying into PointCloud objects: the clouds are
  // already what downsampleCurrentScan would produce. Two uses: (a) device = true — packed float4 clouds
  // resident in device memory (what bench.py times as "inputs resident in HBM"; only pointers and counts
  // cross the boundary); (b) device = false — host clouds read in place with the given stride (32 for
  // pcl::PointXYZI storage, i.e. cloud->points.data()), which is what the ROS node's clouds look like.
  struct FrameView {
    const void* corner_scan = nullptr;     size_t n_corner_scan = 0;      // keyframe corner cloud (:1746)
    const void* corner_scan_ds = nullptr;  size_t n_corner_scan_ds = 0;   // laserCloudCornerScanDS
    const void* surf_last_ds = nullptr;    size_t n_surf_last_ds = 0;     // keyframe surf cloud (:1744)
    const void* outlier_last_ds = nullptr; size_t n_outlier_last_ds = 0;  // keyframe outlier cloud (:1745)
    const void* surf_total_ds = nullptr;   size_t n_surf_total_ds = 0;    // laserCloudSurfTotalLastDS
    bool device = false;
    size_t stride_bytes = 32;  // host clouds only
  };
  void processFrameView(const float laser_odometry[6], const FrameView& frame);
  // The same frame in two halves, for a caller that drives several nodes from one thread (host/batch_replay.cpp): Begin
  // returns as soon as the scan kernel and the keyframe upload are enqueued, End waits for the scan and finishes the frame.
  // `frame` must stay alive until End returns; nothing else may be called on the node in between.
  void processFrameViewBegin(const float laser_odometry[6], const FrameView& frame);
  void processFrameViewEnd();

  // ---- members with the reference's names (mapOptimization.h:311,327-333 and :458,:583) ----------------
  void transformAssociateToMap();
  void transformUpdate();
  void extractSurroundingKeyFrames();
  void downsampleCurrentScan();
  void cornerOptimization();        // stage-level: fills laserCloudOri / coeffSel for the current pose
  void surfOptimization();          // stage-level
  bool LMOptimization(int iterCount);  // stage-level, on laserCloudOri / coeffSel
  void scan2MapOptimization();      // the whole loop in one ABI call
  void scan2MapLaunch();            // ... its first half: enqueue the scan and the keyframe upload
  void scan2MapCollect();           // ... its second half: wait, then :1588-1608
  void saveKeyFramesAndFactor();
  void correctPoses();
  void clearCloud();

  // ---- state with the reference's names -------------------------------------------------------------
  float transformLast[6], transformSum[6], transformIncre[6], transformTobeMapped[6], transformBefMapped[6],
      transformAftMapped[6];
  PointCloud::Ptr laserCloudCornerLast, laserCloudSurfLast, laserCloudCornerScan, laserCloudSurfScan,
      laserCloudOutlierLast;
  PointCloud::Ptr laserCloudCornerLastDS, laserCloudSurfLastDS, laserCloudCornerScanDS, laserCloudSurfScanDS,
      laserCloudOutlierLastDS, laserCloudSurfTotalLast, laserCloudSurfTotalLastDS;
  PointCloud::Ptr laserCloudOri, coeffSel;
  std::vector<PointType> cloudKeyPoses3D;
  std::vector<PointTypePose> cloudKeyPoses6D;
  PointType currentRobotPosPoint, previousRobotPosPoint;
  int laserCloudCornerFromMapDSNum = 0, laserCloudSurfFromMapDSNum = 0;
  bool isDegenerate = false;
  float min_lambda = 0, CF_mean = 0;
  int iter_num = 0;
  int frame_idx = 0;
  bool aLoopIsClosed = false;
  std::vector<PointType> localInfo;  // (iter_num, min_lambda, CF_mean, frame_idx), :1602-1607
  ll_result lastResult;
  float lastPrior[6] = {0, 0, 0, 0, 0, 0};  // transformTobeMapped on entry to the last scan2MapOptimization
  // Scans this library could NOT register because the 1 m search grid over the map's bounding box exceeded the cell
  // budget (ll_result.status == 2). The reference's KD-tree has no such limit: this is a failure to report, not the
  // reference's map-size skip (:1573). The first occurrence is also written to stderr.
  int gridOverflowScans = 0;
  // Wall time of the last processFrame / processFrameView body, milliseconds (steady_clock, the way the reference
  // times the same body for mapt.txt, mapOptmization.cpp:1877,1900-1903).
  double lastFrameMs = 0.0;
  const std::vector<int>& mapIds() const { return mapIds_; }  // keyframe ids of the last extractSurroundingKeyFrames

  // laserCloud{Corner,Surf}FromMapDS on request (publishKeyPosesAndFrames reads the surf one, :746-752)
  PointCloud::Ptr getLaserCloudCornerFromMapDS();
  PointCloud::Ptr getLaserCloudSurfFromMapDS();

  ll_handle handle() const { return h_; }
  const std::string& lastError() const { return err_; }

 private:
  void check(int rc, const char* what);
  PointCloud::Ptr voxel(const PointCloud& in, float leaf);

  ll_handle h_ = nullptr;
  ll_params prm_;
  bool loop_closure_enabled_;
  int search_num_;
  float search_radius_;
  std::string err_;
  // loop-closure-on branch: ids of the keyframes in the recent deque (:1101-1151)
  std::deque<int> recentKeyFrameIDs_;
  int latestFrameID_ = 0;
  // loop-closure-off branch: surroundingExistingKeyPosesID (:1171-1216)
  std::vector<int> surroundingExistingKeyPosesID_;
  std::vector<int> mapIds_;
  bool mapBuilt_ = false;
  bool scanInFlight_ = false;  // between scan2MapLaunch and scan2MapCollect
  std::chrono::steady_clock::time_point frameStart_;
  const FrameView* view_ = nullptr;  // non-null while processFrameView runs
  ll_scan_clouds scanDev_{};  // device copies left by downsampleCurrentScan (valid for the current frame)
  bool scanDevValid_ = false;
  std::vector<PointType> scanLanding_[6];  // host landing buffers of downsampleCurrentScan's six results
};

}  // namespace lego_loam_b200
