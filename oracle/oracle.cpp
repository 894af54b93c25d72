// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of LeGO-LOAM-SR's scan-to-map registration hot path, used (a) as the parity
// checker for the CUDA path in tests/ and __graft_entry__.smoke(), and (b) as the CPU baseline
// that bench.py times next to the GPU (cpu_baseline / --impl reference). Nothing under
// lego-loam-sr_b200/ or include/ may include, link, import or execute this file.
//
// PARITY STATUS: "parity unpinned" for the third-party pieces. The reference has no tests,
// golden vectors or fixtures for this path and its own translation unit
// (mapOptmization.cpp) needs rclcpp + PCL + GTSAM + Eigen, none of which exist here, so it
// cannot be compiled. What IS pinned: the 5-NN search is cross-checked against the reference's
// own vendored nanoflann (oracle/_ref, compiled from /root/reference in place) and against
// brute force; the Jacobian against finite differences; eigenvalues against a closed form.
//
// Every function cites the reference lines it restates (paths relative to
// /root/reference/LeGO-LOAM/src/ unless noted). Conventions fixed here because the reference
// leaves them open (SURVEY.md §7 "hard parts"):
//   * unqualified cos/sin/sqrt/asin/atan2 on float arguments are evaluated in double and
//     rounded to float on assignment (C `double` functions);
//   * VoxelGrid within-voxel accumulation order = ascending original index;
//   * k-NN ties are broken by the lower map index (BASELINE.json north_star).
// Build: g++ -O3 -std=c++17 -ffp-contract=off (same -O3/-std as LeGO-LOAM/CMakeLists.txt:6,10;
// baseline x86-64 has no FMA, the flag only makes that explicit).
#include "oracle.h"
#include "eigen_restate.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <unordered_map>
#include <vector>

using orc_eig::Mat;

namespace {

int g_knn_backend = 1;
orc_ext_build_fn g_ext_build = nullptr;
orc_ext_knn5_fn g_ext_knn5 = nullptr;
orc_ext_free_fn g_ext_free = nullptr;

struct Trig {
  float cr, sr, cp, sp, cy, sy, tx, ty, tz;
};

// updatePointAssociateToMapSinCos :591-604 / updateTransformPointCloudSinCos :622-635
inline Trig make_trig(const float pose6[6]) {
  Trig t;
  t.cr = (float)std::cos((double)pose6[0]);
  t.sr = (float)std::sin((double)pose6[0]);
  t.cp = (float)std::cos((double)pose6[1]);
  t.sp = (float)std::sin((double)pose6[1]);
  t.cy = (float)std::cos((double)pose6[2]);
  t.sy = (float)std::sin((double)pose6[2]);
  t.tx = pose6[3];
  t.ty = pose6[4];
  t.tz = pose6[5];
  return t;
}

// pointAssociateToMap :606-620 == transformPointCloud body :649-664 (same three stages)
inline orc_p4 rigid(const Trig& t, const orc_p4& p) {
  float x1 = t.cy * p.x - t.sy * p.y;
  float y1 = t.sy * p.x + t.cy * p.y;
  float z1 = p.z;
  float x2 = x1;
  float y2 = t.cr * y1 - t.sr * z1;
  float z2 = t.sr * y1 + t.cr * z1;
  orc_p4 o;
  o.x = t.cp * x2 + t.sp * z2 + t.tx;
  o.y = y2 + t.ty;
  o.z = -t.sp * x2 + t.cp * z2 + t.tz;
  o.i = p.i;
  return o;
}

// L2_Simple_Adaptor::evalMetric, include/lego_loam/nanoflann.hpp:431-439
inline float sqdist(const orc_p4& q, const orc_p4& m) {
  float r = 0.f;
  float d = q.x - m.x;
  r += d * d;
  d = q.y - m.y;
  r += d * d;
  d = q.z - m.z;
  r += d * d;
  return r;
}

struct Top5 {
  int idx[5];
  float d2[5];
  int cnt = 0;
  Top5() {
    for (int k = 0; k < 5; ++k) { idx[k] = -1; d2[k] = FLT_MAX; }
  }
  inline void offer(float d, int i) {  // canonical order: ascending (d2, idx)
    if (cnt == 5 && !(d < d2[4] || (d == d2[4] && i < idx[4]))) return;
    int k = cnt < 5 ? cnt : 4;
    while (k > 0 && (d2[k - 1] > d || (d2[k - 1] == d && idx[k - 1] > i))) {
      d2[k] = d2[k - 1];
      idx[k] = idx[k - 1];
      --k;
    }
    d2[k] = d;
    idx[k] = i;
    if (cnt < 5) ++cnt;
  }
};

// ---- own exact k-NN index: 1 m uniform grid with ring expansion -------------------------------
struct GridIndex {
  const orc_p4* pts = nullptr;
  int n = 0;
  std::unordered_map<uint64_t, std::vector<int>> cells;
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  static uint64_t key(int x, int y, int z) {
    return ((uint64_t)(uint32_t)(x + (1 << 20)) << 42) | ((uint64_t)(uint32_t)(y + (1 << 20)) << 21) |
           (uint64_t)(uint32_t)(z + (1 << 20));
  }
  void build(const orc_p4* p, int count) {
    pts = p;
    n = count;
    cells.clear();
    cells.reserve((size_t)count);
    for (int i = 0; i < count; ++i) {
      int c[3] = {(int)std::floor(p[i].x), (int)std::floor(p[i].y), (int)std::floor(p[i].z)};
      for (int a = 0; a < 3; ++a) {
        if (i == 0 || c[a] < lo[a]) lo[a] = c[a];
        if (i == 0 || c[a] > hi[a]) hi[a] = c[a];
      }
      cells[key(c[0], c[1], c[2])].push_back(i);
    }
  }
  void knn5(const orc_p4& q, Top5& t) const {
    if (n == 0) return;
    int c[3] = {(int)std::floor(q.x), (int)std::floor(q.y), (int)std::floor(q.z)};
    int maxr = 0;
    for (int a = 0; a < 3; ++a) {
      maxr = std::max(maxr, std::abs(c[a] - lo[a]));
      maxr = std::max(maxr, std::abs(c[a] - hi[a]));
    }
    for (int r = 0; r <= maxr; ++r) {
      // every point in a cell at Chebyshev ring >= r+1... a point in ring r has |d| > r-1 on
      // some axis, so once the 5th distance is strictly below (r-1)^2 nothing in ring r can
      // enter (strict: an equal distance with a lower index still could).
      if (t.cnt == 5 && r >= 2 && t.d2[4] < (float)(r - 1) * (float)(r - 1)) break;
      for (int dz = -r; dz <= r; ++dz)
        for (int dy = -r; dy <= r; ++dy) {
          bool shell = (std::abs(dz) == r) || (std::abs(dy) == r);
          int step = shell ? 1 : 2 * r;
          if (step == 0) step = 1;
          for (int dx = -r; dx <= r; dx += step) {
            auto it = cells.find(key(c[0] + dx, c[1] + dy, c[2] + dz));
            if (it == cells.end()) continue;
            for (int i : it->second) t.offer(sqdist(q, pts[i]), i);
          }
        }
    }
  }
};

struct KnnIndex {
  int backend;
  const orc_p4* pts;
  int n;
  GridIndex grid;
  void* ext = nullptr;
  KnnIndex(const orc_p4* p, int count, int be) : backend(be), pts(p), n(count) {
    if (backend == 2 && !(g_ext_build && g_ext_knn5)) backend = 1;
    if (backend == 1) grid.build(p, count);
    if (backend == 2) ext = g_ext_build(p, count);
  }
  ~KnnIndex() {
    if (ext && g_ext_free) g_ext_free(ext);
  }
  void query(const orc_p4& q, int idx[5], float d2[5]) const {
    if (backend == 2) {
      // kdtree.nearestKSearch(pointSel, 5, ...) :1276 / :1386 through the reference's nanoflann
      float v[3] = {q.x, q.y, q.z};
      for (int k = 0; k < 5; ++k) { idx[k] = -1; d2[k] = FLT_MAX; }
      g_ext_knn5(ext, v, idx, d2);
      return;
    }
    Top5 t;
    if (backend == 0) {
      for (int i = 0; i < n; ++i) t.offer(sqdist(q, pts[i]), i);
    } else {
      grid.knn5(q, t);
    }
    for (int k = 0; k < 5; ++k) { idx[k] = t.idx[k]; d2[k] = t.d2[k]; }
  }
};

// ---- a7 per-point body: cornerOptimization :1279-1373 -------------------------------------------
inline bool corner_fit(const orc_p4 nb[5], const orc_p4& sel, const orc_params& prm, float coeff[4]) {
  float cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < 5; ++j) { cx += nb[j].x; cy += nb[j].y; cz += nb[j].z; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
  for (int j = 0; j < 5; ++j) {
    float ax = nb[j].x - cx, ay = nb[j].y - cy, az = nb[j].z - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
    a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  Mat<3, 3> A, V;
  A(0, 0) = a11; A(0, 1) = a12; A(0, 2) = a13;
  A(1, 0) = a12; A(1, 1) = a22; A(1, 2) = a23;
  A(2, 0) = a13; A(2, 1) = a23; A(2, 2) = a33;
  float D[3];
  orc_eig::self_adjoint_eig<3>(A, D, V);  // ascending eigenvalues, eigenvectors in columns
  if (!(D[2] > prm.eig_ratio * D[1])) return false;
  float vx, vy, vz;
  if (prm.eigen_mode == 0) {  // as shipped: ROW 0 of the column-eigenvector matrix :1329-1334
    vx = V(0, 0); vy = V(0, 1); vz = V(0, 2);
  } else {  // upstream semantics: the principal eigenvector
    vx = V(0, 2); vy = V(1, 2); vz = V(2, 2);
  }
  float x0 = sel.x, y0 = sel.y, z0 = sel.z;
  float x1 = (float)(cx + 0.1 * vx), y1 = (float)(cy + 0.1 * vy), z1 = (float)(cz + 0.1 * vz);
  float x2 = (float)(cx - 0.1 * vx), y2 = (float)(cy - 0.1 * vy), z2 = (float)(cz - 0.1 * vz);
  float cxy = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
  float cxz = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
  float cyz = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
  float a012 = (float)std::sqrt((double)(cxy * cxy + cxz * cxz + cyz * cyz));
  float l12 = (float)std::sqrt((double)((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2)));
  float la = ((y1 - y2) * cxy + (z1 - z2) * cxz) / a012 / l12;
  float lb = -((x1 - x2) * cxy - (z1 - z2) * cyz) / a012 / l12;
  float lc = -((x1 - x2) * cxz + (y1 - y2) * cyz) / a012 / l12;
  float ld2 = a012 / l12;
  float s = (float)(1 - prm.weight_k * std::fabs((double)ld2));
  coeff[0] = s * la;
  coeff[1] = s * lb;
  coeff[2] = s * lc;
  coeff[3] = s * ld2;
  return s > prm.min_weight;
}

// ---- a8 per-point body: surfOptimization :1389-1438 ---------------------------------------------
inline bool surf_fit(const orc_p4 nb[5], const orc_p4& sel, const orc_params& prm, float coeff[4]) {
  Mat<5, 3> A0;
  for (int j = 0; j < 5; ++j) { A0(j, 0) = nb[j].x; A0(j, 1) = nb[j].y; A0(j, 2) = nb[j].z; }
  const float b0[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};  // matB0.fill(-1) :278
  float X0[3];
  orc_eig::colpiv_qr_solve<5, 3>(A0, b0, X0);
  float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
  float ps = (float)std::sqrt((double)(pa * pa + pb * pb + pc * pc));
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  for (int j = 0; j < 5; ++j)
    if (std::fabs(pa * nb[j].x + pb * nb[j].y + pc * nb[j].z + pd) > prm.plane_tol) return false;
  float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
  float nrm2 = sel.x * sel.x + sel.y * sel.y + sel.z * sel.z;
  float s = (float)(1 - prm.weight_k * std::fabs((double)pd2) / std::sqrt(std::sqrt((double)nrm2)));
  coeff[0] = s * pa;
  coeff[1] = s * pb;
  coeff[2] = s * pc;
  coeff[3] = s * pd2;
  return s > prm.min_weight;
}

template <bool kCorner>
void residuals(const KnnIndex& index, const orc_p4* map, const orc_p4* q, int nq, const float pose6[6],
               const orc_params& prm, float* coeff, int* mask, int* accept, int* knn_idx, float* knn_d2) {
  Trig t = make_trig(pose6);
  for (int i = 0; i < nq; ++i) {
    orc_p4 sel = rigid(t, q[i]);
    int idx[5];
    float d2[5];
    index.query(sel, idx, d2);
    if (knn_idx) std::memcpy(knn_idx + 5 * i, idx, sizeof(idx));
    if (knn_d2) std::memcpy(knn_d2 + 5 * i, d2, sizeof(d2));
    bool acc = idx[4] >= 0 && d2[4] < prm.max_sq_dist;
    if (accept) accept[i] = acc;
    float c[4] = {0, 0, 0, 0};
    bool keep = false;
    if (acc) {
      orc_p4 nb[5];
      for (int j = 0; j < 5; ++j) nb[j] = map[idx[j]];
      keep = kCorner ? corner_fit(nb, sel, prm, c) : surf_fit(nb, sel, prm, c);
      if (!keep) c[0] = c[1] = c[2] = c[3] = 0.f;
    }
    mask[i] = keep;
    std::memcpy(coeff + 4 * i, c, sizeof(c));
  }
}

// ---- a9 LMOptimization :1444-1570 ----------------------------------------------------------------
int lm_step(const orc_p4* ori, const float* coeff, int n, int iter, float pose[6], const orc_params& prm,
            orc_state& st, float* AtA36, float* Atb6, float* x6, float* pdeltaR, float* pdeltaT) {
  float srx = (float)std::sin((double)pose[0]), crx = (float)std::cos((double)pose[0]);
  float sry = (float)std::sin((double)pose[1]), cry = (float)std::cos((double)pose[1]);
  float srz = (float)std::sin((double)pose[2]), crz = (float)std::cos((double)pose[2]);
  if (n < prm.min_residuals) return -1;  // :1453
  Mat<6, 6> AtA;
  AtA.zero();
  float Atb[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const float px = ori[i].x, py = ori[i].y, pz = ori[i].z;
    const float nx = coeff[4 * i], ny = coeff[4 * i + 1], nz = coeff[4 * i + 2], d = coeff[4 * i + 3];
    float row[6];
    row[0] = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * nx +
             (-srx * srz * px - crz * srx * py - crx * pz) * ny +
             (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * nz;
    row[1] = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * nx +
             ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * nz;
    row[2] = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * nx +
             (crx * crz * px - crx * srz * py) * ny +
             ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * nz;
    row[3] = nx;
    row[4] = ny;
    row[5] = nz;
    float b = -prm.step_size * d;  // :1500
    for (int r = 0; r < 6; ++r) {
      for (int c = 0; c < 6; ++c) AtA(r, c) += row[r] * row[c];
      Atb[r] += row[r] * b;
    }
  }
  float x[6];
  orc_eig::colpiv_qr_solve<6, 6>(AtA, Atb, x);  // :1505
  if (iter == 0) {                              // :1507-1531
    float E[6];
    Mat<6, 6> V, V2, P;
    orc_eig::self_adjoint_eig<6>(AtA, E, V);
    st.min_lambda = E[0];
    st.is_degenerate = 0;
    if (prm.eigen_mode == 0) {
      // as shipped: ascending eigenvalues, walk from i = 5, zero ROW i of the column matrix
      V2 = V;
      for (int i = 5; i >= 0; --i) {
        if (E[i] < prm.eig_thresh) {
          for (int j = 0; j < 6; ++j) V2(i, j) = 0.f;
          st.is_degenerate = 1;
        } else {
          break;
        }
      }
      Mat<6, 6> Vinv;
      orc_eig::inverse_lu<6>(V, Vinv);
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          float acc = 0.f;
          for (int k = 0; k < 6; ++k) acc += Vinv(r, k) * V2(k, c);
          P(r, c) = acc;
        }
    } else {
      // upstream semantics (OpenCV row eigenvectors, descending): drop the eigenvectors whose
      // eigenvalue is below the threshold, smallest first; matP = sum over kept v v^T.
      int first_kept = 0;
      while (first_kept < 6 && E[first_kept] < prm.eig_thresh) ++first_kept;
      st.is_degenerate = first_kept > 0;
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          float acc = 0.f;
          for (int k = first_kept; k < 6; ++k) acc += V(r, k) * V(c, k);
          P(r, c) = acc;
        }
    }
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) st.matP[6 * r + c] = P(r, c);
  }
  if (st.is_degenerate) {  // :1533-1537
    float x2[6];
    for (int r = 0; r < 6; ++r) {
      float acc = 0.f;
      for (int c = 0; c < 6; ++c) acc += st.matP[6 * r + c] * x[c];
      x2[r] = acc;
    }
    std::memcpy(x, x2, sizeof(x));
  }
  if (prm.apply_update)  // :1540-1545 (commented out as shipped)
    for (int k = 0; k < 6; ++k) pose[k] += x[k];
  // :1547-1552  pcl::rad2deg(float) = alpha * 57.29578f; pow(float, int) evaluates in double
  double r0 = (double)(x[0] * 57.29578f), r1 = (double)(x[1] * 57.29578f), r2 = (double)(x[2] * 57.29578f);
  float deltaR = (float)std::sqrt(r0 * r0 + r1 * r1 + r2 * r2);
  double t0 = (double)(x[3] * 100), t1 = (double)(x[4] * 100), t2 = (double)(x[5] * 100);
  float deltaT = (float)std::sqrt(t0 * t0 + t1 * t1 + t2 * t2);
  float cf_all = 0.f;  // :1555-1568
  for (int i = 0; i < n; ++i) cf_all += std::fabs(coeff[4 * i + 3]);
  st.cf_mean = cf_all / n;
  if (AtA36)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) AtA36[6 * r + c] = AtA(r, c);
  if (Atb6) std::memcpy(Atb6, Atb, sizeof(Atb));
  if (x6) std::memcpy(x6, x, sizeof(x));
  if (pdeltaR) *pdeltaR = deltaR;
  if (pdeltaT) *pdeltaT = deltaT;
  return (deltaR < prm.stop_thres && deltaT < prm.stop_thres) ? 1 : 0;
}

// ---- a3 VoxelGrid -----------------------------------------------------------------------------------
int voxel_downsample(const orc_p4* in, int n, float leaf, std::vector<orc_p4>& out, int* member) {
  out.clear();
  if (n == 0) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int finite = 0;
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(in[i].x) || !std::isfinite(in[i].y) || !std::isfinite(in[i].z)) continue;
    ++finite;
    mn[0] = std::min(mn[0], in[i].x); mx[0] = std::max(mx[0], in[i].x);
    mn[1] = std::min(mn[1], in[i].y); mx[1] = std::max(mx[1], in[i].y);
    mn[2] = std::min(mn[2], in[i].z); mx[2] = std::max(mx[2], in[i].z);
  }
  if (finite == 0) {
    if (member) for (int i = 0; i < n; ++i) member[i] = -1;
    return 0;
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT_MAX) {  // PCL warns and returns the input unchanged
    out.assign(in, in + n);
    if (member) for (int i = 0; i < n; ++i) member[i] = i;
    return 1;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    max_b[a] = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<uint64_t> keys;
  keys.reserve(n);
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(in[i].x) || !std::isfinite(in[i].y) || !std::isfinite(in[i].z)) {
      if (member) member[i] = -1;
      continue;
    }
    int i0 = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
    int i1 = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
    int i2 = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
    int idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
    keys.push_back(((uint64_t)(uint32_t)idx << 32) | (uint32_t)i);
  }
  std::sort(keys.begin(), keys.end());  // (voxel idx, original index): canonical stable order
  size_t k = 0;
  while (k < keys.size()) {
    uint32_t v = (uint32_t)(keys[k] >> 32);
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    size_t first = k;
    while (k < keys.size() && (uint32_t)(keys[k] >> 32) == v) {
      const orc_p4& p = in[(uint32_t)keys[k]];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
      if (member) member[(uint32_t)keys[k]] = (int)out.size();
      ++k;
    }
    float cnt = (float)(k - first);
    orc_p4 o = {sx / cnt, sy / cnt, sz / cnt, si / cnt};
    out.push_back(o);
  }
  return 0;
}

void scan2map(const orc_p4* cmap, int nc, const orc_p4* smap, int ns, const orc_p4* cq, int nqc,
              const orc_p4* sq, int nqs, float pose[6], const orc_params& prm, orc_state& st,
              orc_result& res) {
  std::memset(&res, 0, sizeof(res));
  res.status = 1;
  res.min_lambda = st.min_lambda;
  res.cf_mean = st.cf_mean;
  res.is_degenerate = st.is_degenerate;
  std::memcpy(res.pose, pose, 6 * sizeof(float));
  if (!(nc > prm.min_corner_map && ns > prm.min_surf_map)) return;  // :1573
  res.status = 0;
  KnnIndex ci(cmap, nc, g_knn_backend), si(smap, ns, g_knn_backend);  // :1575-1576
  std::vector<float> coeff_c(4 * (size_t)nqc), coeff_s(4 * (size_t)nqs);
  std::vector<int> mask_c(nqc), mask_s(nqs);
  std::vector<orc_p4> ori;
  std::vector<float> coeff;
  int iter_num = 0;
  for (int it = 0; it < prm.iter_count_thres; ++it) {  // :1578-1586
    ori.clear();
    coeff.clear();
    iter_num++;
    residuals<true>(ci, cmap, cq, nqc, pose, prm, coeff_c.data(), mask_c.data(), nullptr, nullptr, nullptr);
    residuals<false>(si, smap, sq, nqs, pose, prm, coeff_s.data(), mask_s.data(), nullptr, nullptr, nullptr);
    for (int i = 0; i < nqc; ++i)
      if (mask_c[i]) { ori.push_back(cq[i]); coeff.insert(coeff.end(), &coeff_c[4 * i], &coeff_c[4 * i] + 4); }
    for (int i = 0; i < nqs; ++i)
      if (mask_s[i]) { ori.push_back(sq[i]); coeff.insert(coeff.end(), &coeff_s[4 * i], &coeff_s[4 * i] + 4); }
    res.sel_num = (int)ori.size();
    int r = lm_step(ori.data(), coeff.data(), (int)ori.size(), it, pose, prm, st, nullptr, nullptr,
                    res.last_x, nullptr, nullptr);
    if (r == 1) { res.converged = 1; break; }
  }
  res.iter_num = iter_num;
  res.min_lambda = st.min_lambda;
  res.cf_mean = st.cf_mean;
  res.is_degenerate = st.is_degenerate;
  std::memcpy(res.pose, pose, 6 * sizeof(float));
}

void build_local_map(int n_kf, const float* poses6, const orc_p4* corner, const int* coff, const orc_p4* surf,
                     const int* soff, const orc_p4* outl, const int* ooff, const orc_params& prm,
                     std::vector<orc_p4>& cds, std::vector<orc_p4>& sds, int* ncraw, int* nsraw) {
  std::vector<orc_p4> craw, sraw;
  for (int k = 0; k < n_kf; ++k) {  // :1147-1151 / :1218-1222 (cache order)
    Trig t = make_trig(poses6 + 6 * k);
    for (int i = coff[k]; i < coff[k + 1]; ++i) craw.push_back(rigid(t, corner[i]));
    for (int i = soff[k]; i < soff[k + 1]; ++i) sraw.push_back(rigid(t, surf[i]));
    for (int i = ooff[k]; i < ooff[k + 1]; ++i) sraw.push_back(rigid(t, outl[i]));
  }
  if (ncraw) *ncraw = (int)craw.size();
  if (nsraw) *nsraw = (int)sraw.size();
  if (prm.map_downsample) {  // :1225-1231
    voxel_downsample(craw.data(), (int)craw.size(), prm.corner_leaf, cds, nullptr);
    voxel_downsample(sraw.data(), (int)sraw.size(), prm.surf_leaf, sds, nullptr);
  } else {
    cds = craw;
    sds = sraw;
  }
}

}  // namespace

struct orc_replay {
  orc_params prm;
  orc_state st;
  int window;
  bool loop_closure = true;   // _loop_closure_enabled (loam_config.yaml:23; the VLP-16 block ships false)
  float search_radius = 50.f; // _surrounding_keyframe_search_radius (loam_config.yaml:26)
  float cur_pos[3] = {0.f, 0.f, 0.f};  // currentRobotPosPoint (set at :1613-1615)
  std::vector<int> existing_ids;       // surroundingExistingKeyPosesID (:1171-1216)
  std::vector<int> map_ids;            // keyframes concatenated into the last local map, in order
  float sum[6], bef[6], aft[6], tobe[6];
  std::vector<std::vector<orc_p4>> kf_corner, kf_surf, kf_outlier;
  std::vector<float> kf_pose;  // 6 per keyframe
  std::vector<orc_p4> cds, sds;
};

// kdtreeSurroundingKeyPoses.radiusSearch (nanoflann_pcl.h:154-174 over nanoflann.hpp RadiusResultSet): every point
// with squared distance (float, ((dx^2)+dy^2)+dz^2, query minus point) STRICTLY below radius^2 (the double product
// rounded to float), sorted ascending by distance. The reference sorts with std::sort (IndexDist_Sorter compares the
// distance only), so the order among exactly equal distances is unspecified there; here it is ascending index.
static orc_ext_radius_fn g_ext_radius = nullptr;
static void radius_search(const std::vector<orc_p4>& pts, const float q[3], float radius, std::vector<int>& idx, bool use_ext) {
  idx.clear();
  if (use_ext && g_ext_radius && g_ext_build && g_ext_free && !pts.empty()) {
    void* tree = g_ext_build(pts.data(), (int)pts.size());
    std::vector<int> tmp(pts.size());
    std::vector<float> d2(pts.size());
    const int n = g_ext_radius(tree, q, (double)radius, tmp.data(), d2.data(), (int)pts.size());
    g_ext_free(tree);
    idx.assign(tmp.begin(), tmp.begin() + n);
    return;
  }
  const float r2 = static_cast<float>((double)radius * (double)radius);
  std::vector<std::pair<float, int>> hits;
  for (int i = 0; i < (int)pts.size(); ++i) {
    float df = q[0] - pts[i].x;
    float d = df * df;
    df = q[1] - pts[i].y;
    d += df * df;
    df = q[2] - pts[i].z;
    d += df * df;
    if (d < r2) hits.emplace_back(d, i);
  }
  std::stable_sort(hits.begin(), hits.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
  for (const auto& h : hits) idx.push_back(h.second);
}

extern "C" {

void orc_default_params(orc_params* p) {
  p->corner_leaf = 0.2f;
  p->surf_leaf = 0.4f;
  p->max_sq_dist = 1.0f;
  p->eig_ratio = 3.f;
  p->plane_tol = 0.2;
  p->weight_k = 0.9;
  p->min_weight = 0.1;
  p->min_residuals = 50;
  p->eig_thresh = 100.f;
  p->iter_count_thres = 200;
  p->step_size = 1.0f;
  p->stop_thres = 0.05f;
  p->apply_update = 0;
  p->eigen_mode = 0;
  p->map_downsample = 1;
  p->min_corner_map = 10;
  p->min_surf_map = 100;
}

void orc_init_state(orc_state* s) {  // :267-302
  std::memset(s, 0, sizeof(*s));
}

void orc_set_knn_backend(int kind) { g_knn_backend = kind; }
void orc_set_ext_knn(orc_ext_build_fn b, orc_ext_knn5_fn q, orc_ext_free_fn f) {
  g_ext_build = b;
  g_ext_knn5 = q;
  g_ext_free = f;
}

void orc_transform_cloud(const orc_p4* in, int n, const float pose6[6], orc_p4* out) {
  Trig t = make_trig(pose6);
  for (int i = 0; i < n; ++i) out[i] = rigid(t, in[i]);
}

void orc_point_associate_to_map(const orc_p4* in, int n, const float pose6[6], orc_p4* out) {
  orc_transform_cloud(in, n, pose6, out);
}

int orc_voxel_downsample(const orc_p4* in, int n, float leaf, orc_p4* out, int* n_out, int* member) {
  std::vector<orc_p4> o;
  int flag = voxel_downsample(in, n, leaf, o, member);
  if (out) std::memcpy(out, o.data(), o.size() * sizeof(orc_p4));
  *n_out = (int)o.size();
  return flag;
}

void orc_knn5(const orc_p4* map, int n, const orc_p4* q, int nq, int* idx, float* d2, int backend) {
  KnnIndex index(map, n, backend);
  for (int i = 0; i < nq; ++i) index.query(q[i], idx + 5 * i, d2 + 5 * i);
}

void orc_corner_residuals(const orc_p4* map, int n, const orc_p4* q, int nq, const float pose6[6],
                          const orc_params* prm, float* coeff, int* mask, int* accept, int* knn_idx,
                          float* knn_d2) {
  KnnIndex index(map, n, g_knn_backend);
  residuals<true>(index, map, q, nq, pose6, *prm, coeff, mask, accept, knn_idx, knn_d2);
}

void orc_surf_residuals(const orc_p4* map, int n, const orc_p4* q, int nq, const float pose6[6],
                        const orc_params* prm, float* coeff, int* mask, int* accept, int* knn_idx,
                        float* knn_d2) {
  KnnIndex index(map, n, g_knn_backend);
  residuals<false>(index, map, q, nq, pose6, *prm, coeff, mask, accept, knn_idx, knn_d2);
}

int orc_lm_step(const orc_p4* ori, const float* coeff, int n, int iter, float pose6[6], const orc_params* prm,
                orc_state* st, float* AtA36, float* Atb6, float* x6, float* deltaR, float* deltaT) {
  return lm_step(ori, coeff, n, iter, pose6, *prm, *st, AtA36, Atb6, x6, deltaR, deltaT);
}

void orc_scan2map(const orc_p4* cmap, int nc, const orc_p4* smap, int ns, const orc_p4* cq, int nqc,
                  const orc_p4* sq, int nqs, float pose6[6], const orc_params* prm, orc_state* st,
                  orc_result* res) {
  scan2map(cmap, nc, smap, ns, cq, nqc, sq, nqs, pose6, *prm, *st, *res);
}

void orc_build_local_map(int n_kf, const float* poses6, const orc_p4* corner, const int* corner_off,
                         const orc_p4* surf, const int* surf_off, const orc_p4* outlier, const int* outlier_off,
                         const orc_params* prm, orc_p4* corner_ds, int* n_corner_ds, orc_p4* surf_ds,
                         int* n_surf_ds, int* n_corner_raw, int* n_surf_raw) {
  std::vector<orc_p4> cds, sds;
  build_local_map(n_kf, poses6, corner, corner_off, surf, surf_off, outlier, outlier_off, *prm, cds, sds,
                  n_corner_raw, n_surf_raw);
  if (corner_ds) std::memcpy(corner_ds, cds.data(), cds.size() * sizeof(orc_p4));
  if (surf_ds) std::memcpy(surf_ds, sds.data(), sds.size() * sizeof(orc_p4));
  *n_corner_ds = (int)cds.size();
  *n_surf_ds = (int)sds.size();
}

// transformAssociateToMap :458-581. The trigonometric identity is the composition
// R_tobe = R_aft * R_bef^-1 * R_sum; the expressions below follow the reference's expansion.
void orc_transform_associate_to_map(const float sum[6], const float bef[6], const float aft[6], float tobe[6]) {
  auto S = [](float a) { return (float)std::sin((double)a); };
  auto C = [](float a) { return (float)std::cos((double)a); };
  float x1 = C(sum[1]) * (bef[3] - sum[3]) - S(sum[1]) * (bef[5] - sum[5]);
  float y1 = bef[4] - sum[4];
  float z1 = S(sum[1]) * (bef[3] - sum[3]) + C(sum[1]) * (bef[5] - sum[5]);
  float x2 = x1;
  float y2 = C(sum[0]) * y1 + S(sum[0]) * z1;
  float z2 = -S(sum[0]) * y1 + C(sum[0]) * z1;
  float inc3 = C(sum[2]) * x2 + S(sum[2]) * y2;
  float inc4 = -S(sum[2]) * x2 + C(sum[2]) * y2;
  float inc5 = z2;

  float sbcx = S(sum[0]), cbcx = C(sum[0]), sbcy = S(sum[1]), cbcy = C(sum[1]), sbcz = S(sum[2]), cbcz = C(sum[2]);
  float sblx = S(bef[0]), cblx = C(bef[0]), sbly = S(bef[1]), cbly = C(bef[1]), sblz = S(bef[2]), cblz = C(bef[2]);
  float salx = S(aft[0]), calx = C(aft[0]), saly = S(aft[1]), caly = C(aft[1]), salz = S(aft[2]), calz = C(aft[2]);

  float srx = -sbcx * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz) -
              cbcx * sbcy *
                  (calx * calz * (cbly * sblz - cblz * sblx * sbly) -
                   calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
              cbcx * cbcy *
                  (calx * salz * (cblz * sbly - cbly * sblx * sblz) -
                   calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx);
  tobe[0] = (float)(-std::asin((double)srx));

  float srycrx = sbcx * (cblx * cblz * (caly * salz - calz * salx * saly) -
                         cblx * sblz * (caly * calz + salx * saly * salz) + calx * saly * sblx) -
                 cbcx * cbcy *
                     ((caly * calz + salx * saly * salz) * (cblz * sbly - cbly * sblx * sblz) +
                      (caly * salz - calz * salx * saly) * (sbly * sblz + cbly * cblz * sblx) -
                      calx * cblx * cbly * saly) +
                 cbcx * sbcy *
                     ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) +
                      (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) +
                      calx * cblx * saly * sbly);
  float crycrx = sbcx * (cblx * sblz * (calz * saly - caly * salx * salz) -
                         cblx * cblz * (saly * salz + caly * calz * salx) + calx * caly * sblx) +
                 cbcx * cbcy *
                     ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) +
                      (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) +
                      calx * caly * cblx * cbly) -
                 cbcx * sbcy *
                     ((saly * salz + caly * calz * salx) * (cbly * sblz - cblz * sblx * sbly) +
                      (calz * saly - caly * salx * salz) * (cbly * cblz + sblx * sbly * sblz) -
                      calx * caly * cblx * sbly);
  float c0 = C(tobe[0]);
  tobe[1] = (float)std::atan2((double)(srycrx / c0), (double)(crycrx / c0));

  float srzcrx = (cbcz * sbcy - cbcy * sbcx * sbcz) *
                     (calx * salz * (cblz * sbly - cbly * sblx * sblz) -
                      calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) -
                 (cbcy * cbcz + sbcx * sbcy * sbcz) *
                     (calx * calz * (cbly * sblz - cblz * sblx * sbly) -
                      calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) +
                 cbcx * sbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  float crzcrx = (cbcy * sbcz - cbcz * sbcx * sbcy) *
                     (calx * calz * (cbly * sblz - cblz * sblx * sbly) -
                      calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                 (sbcy * sbcz + cbcy * cbcz * sbcx) *
                     (calx * salz * (cblz * sbly - cbly * sblx * sblz) -
                      calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) +
                 cbcx * cbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  tobe[2] = (float)std::atan2((double)(srzcrx / c0), (double)(crzcrx / c0));

  x1 = C(tobe[2]) * inc3 - S(tobe[2]) * inc4;
  y1 = S(tobe[2]) * inc3 + C(tobe[2]) * inc4;
  z1 = inc5;
  x2 = x1;
  y2 = C(tobe[0]) * y1 - S(tobe[0]) * z1;
  z2 = S(tobe[0]) * y1 + C(tobe[0]) * z1;
  tobe[3] = aft[3] - (C(tobe[1]) * x2 + S(tobe[1]) * z2);
  tobe[4] = aft[4] - y2;
  tobe[5] = aft[5] - (-S(tobe[1]) * x2 + C(tobe[1]) * z2);
}

// ---- FeatureAssociation::calculateTransformationSurf / Corner / calculateTransformation ----------------
// (featureAssociation.cpp:1925-1997, :2068-2131, :2208-2271): rows restricted to the masked unknowns,
// matAtA / matAtB as float products, colPivHouseholderQr solve, and on iterCount == 0 the eigen / degeneracy
// projector with the given threshold (10 there). matP / isDegenerate are the caller's (function-local in the
// reference). M = number of unknowns (3 or 6).
}  // extern "C" (templates need C++ linkage)

template <int M>
static void solve_normal_eq_t(const float* rows6, const float* b, int n, const int* cols, int iter, float eig_thresh,
                              int eigen_mode, float* x_out, float* matP, int* is_degenerate) {
  Mat<M, M> AtA;
  AtA.zero();
  float Atb[M];
  for (int k = 0; k < M; ++k) Atb[k] = 0.f;
  for (int i = 0; i < n; ++i) {
    float r[M];
    for (int k = 0; k < M; ++k) r[k] = rows6[6 * i + cols[k]];
    for (int a = 0; a < M; ++a) {
      for (int c = 0; c < M; ++c) AtA(a, c) += r[a] * r[c];
      Atb[a] += r[a] * b[i];
    }
  }
  float x[M];
  orc_eig::colpiv_qr_solve<M, M>(AtA, Atb, x);
  if (iter == 0) {
    float E[M];
    Mat<M, M> V, V2, P;
    orc_eig::self_adjoint_eig<M>(AtA, E, V);
    int deg = 0;
    if (eigen_mode == 0) {
      V2 = V;
      for (int i = M - 1; i >= 0; --i) {
        if (E[i] < eig_thresh) {
          for (int j = 0; j < M; ++j) V2(i, j) = 0.f;
          deg = 1;
        } else {
          break;
        }
      }
      Mat<M, M> Vinv;
      if constexpr (M == 3) orc_eig::inverse3_cofactor(V, Vinv);
      else orc_eig::inverse_lu<M>(V, Vinv);
      for (int r = 0; r < M; ++r)
        for (int c = 0; c < M; ++c) {
          float acc = 0.f;
          for (int k = 0; k < M; ++k) acc += Vinv(r, k) * V2(k, c);
          P(r, c) = acc;
        }
    } else {
      int first_kept = 0;
      while (first_kept < M && E[first_kept] < eig_thresh) ++first_kept;
      deg = first_kept > 0;
      for (int r = 0; r < M; ++r)
        for (int c = 0; c < M; ++c) {
          float acc = 0.f;
          for (int k = first_kept; k < M; ++k) acc += V(r, k) * V(c, k);
          P(r, c) = acc;
        }
    }
    *is_degenerate = deg;
    for (int r = 0; r < M; ++r)
      for (int c = 0; c < M; ++c) matP[M * r + c] = P(r, c);
  }
  if (*is_degenerate) {
    float x2[M];
    for (int r = 0; r < M; ++r) {
      float acc = 0.f;
      for (int c = 0; c < M; ++c) acc += matP[M * r + c] * x[c];
      x2[r] = acc;
    }
    for (int r = 0; r < M; ++r) x[r] = x2[r];
  }
  for (int r = 0; r < M; ++r) x_out[r] = x[r];
}

extern "C" {

int orc_solve_normal_eq(const float* rows6, const float* b, int n, int dof_mask, int iter, float eig_thresh,
                        int eigen_mode, float* x_out, float* matP_inout, int* is_degenerate_inout) {
  int cols[6], m = 0;
  for (int k = 0; k < 6; ++k)
    if (dof_mask & (1 << k)) cols[m++] = k;
  if (m == 3) solve_normal_eq_t<3>(rows6, b, n, cols, iter, eig_thresh, eigen_mode, x_out, matP_inout, is_degenerate_inout);
  else if (m == 6) solve_normal_eq_t<6>(rows6, b, n, cols, iter, eig_thresh, eigen_mode, x_out, matP_inout, is_degenerate_inout);
  else return -1;
  return 0;
}

// ---- direct access to the restated Eigen pieces, for the oracle's own cross-checks -------------------
// A: row-major n x n symmetric (n = 3 or 6); evals ascending, evecs row-major with eigenvectors in columns.
void orc_test_self_adjoint_eig(const float* A, int n, float* evals, float* evecs) {
  if (n == 3) {
    Mat<3, 3> M, V;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M(r, c) = A[3 * r + c];
    orc_eig::self_adjoint_eig<3>(M, evals, V);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) evecs[3 * r + c] = V(r, c);
  } else {
    Mat<6, 6> M, V;
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) M(r, c) = A[6 * r + c];
    orc_eig::self_adjoint_eig<6>(M, evals, V);
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) evecs[6 * r + c] = V(r, c);
  }
}

// A: row-major (5x3 or 6x6), b: rows; x: cols.
void orc_test_colpiv_qr_solve(const float* A, const float* b, int rows, int cols, float* x) {
  if (rows == 5 && cols == 3) {
    Mat<5, 3> M;
    for (int r = 0; r < 5; ++r)
      for (int c = 0; c < 3; ++c) M(r, c) = A[3 * r + c];
    orc_eig::colpiv_qr_solve<5, 3>(M, b, x);
  } else {
    Mat<6, 6> M;
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) M(r, c) = A[6 * r + c];
    orc_eig::colpiv_qr_solve<6, 6>(M, b, x);
  }
}

void orc_test_inverse6(const float* A, float* inv) {
  Mat<6, 6> M, I;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) M(r, c) = A[6 * r + c];
  orc_eig::inverse_lu<6>(M, I);
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) inv[6 * r + c] = I(r, c);
}

orc_replay* orc_replay_create(const orc_params* prm, int window) {
  orc_replay* r = new orc_replay();
  r->prm = *prm;
  orc_init_state(&r->st);
  r->window = window;
  for (int i = 0; i < 6; ++i) r->sum[i] = r->bef[i] = r->aft[i] = r->tobe[i] = 0.f;  // :267-274
  return r;
}

// loop_closure_enabled = 0 selects the reference's DEFAULT local-map branch (:1152-1223): key poses within
// search_radius of currentRobotPosPoint, thinned by a 1 m VoxelGrid, cached set updated by erase / add.
orc_replay* orc_replay_create2(const orc_params* prm, int window, int loop_closure_enabled, float search_radius) {
  orc_replay* r = orc_replay_create(prm, window);
  r->loop_closure = loop_closure_enabled != 0;
  r->search_radius = search_radius;
  return r;
}

void orc_set_ext_radius(orc_ext_radius_fn f) { g_ext_radius = f; }

int orc_replay_get_map_ids(orc_replay* r, int* ids, int cap) {
  for (int i = 0; i < (int)r->map_ids.size() && i < cap; ++i) ids[i] = r->map_ids[(size_t)i];
  return (int)r->map_ids.size();
}

// Stand-alone form of the selection of :1152-1216 for tests: key poses (x, y, z, intensity = index) -> the updated
// surroundingExistingKeyPosesID. existing_inout holds n_existing ids on entry and the new list on return (capacity
// cap); returns the new count.
int orc_select_surrounding(const orc_p4* key_poses, int n, const float cur_pos[3], float radius, int* existing_inout,
                           int n_existing, int cap, int use_ext_radius) {
  orc_replay tmp;
  tmp.existing_ids.assign(existing_inout, existing_inout + n_existing);
  std::vector<orc_p4> kp(key_poses, key_poses + n);
  std::vector<int> hit;
  radius_search(kp, cur_pos, radius, hit, use_ext_radius != 0);
  std::vector<orc_p4> sur, sur_ds(hit.size());
  for (int i : hit) sur.push_back(kp[(size_t)i]);
  int n_ds = 0;
  if (!sur.empty()) orc_voxel_downsample(sur.data(), (int)sur.size(), 1.0f, sur_ds.data(), &n_ds, nullptr);
  std::vector<int>& ex = tmp.existing_ids;
  for (size_t i = 0; i < ex.size(); ++i) {
    bool existing = false;
    for (int j = 0; j < n_ds; ++j)
      if (ex[i] == (int)sur_ds[(size_t)j].i) { existing = true; break; }
    if (!existing) { ex.erase(ex.begin() + (long)i); --i; }
  }
  for (int i = 0; i < n_ds; ++i) {
    const int id = (int)sur_ds[(size_t)i].i;
    if (std::find(ex.begin(), ex.end(), id) == ex.end()) ex.push_back(id);
  }
  for (int i = 0; i < (int)ex.size() && i < cap; ++i) existing_inout[i] = ex[(size_t)i];
  return (int)ex.size();
}

void orc_replay_destroy(orc_replay* r) { delete r; }

// (bench: the as-shipped replay fills its window with iter_count_thres = 1 — same state, the pose never moves — then
// switches to the shipped 200 for the timed scans)
void orc_replay_set_params(orc_replay* r, const orc_params* prm) { r->prm = *prm; }

// The body of MapOptimization::run for one frame (:1878-1888) with loop closure enabled so the
// local map is the last `window` keyframes (:1101-1151), every frame a keyframe (:1622-1631) and
// the iSAM2 estimate equal to the odometry chain (pose graph pass-through).
void orc_replay_step(orc_replay* r, const float odom6[6], const orc_p4* corner_scan, int n_cs,
                     const orc_p4* corner_scan_ds, int n_csd, const orc_p4* surf_last_ds, int n_sl,
                     const orc_p4* outlier_last_ds, int n_ol, const orc_p4* surf_total_ds, int n_st,
                     orc_result* res, int* n_corner_map_ds, int* n_surf_map_ds) {
  std::memcpy(r->sum, odom6, sizeof(r->sum));                                // :1878
  orc_transform_associate_to_map(r->sum, r->bef, r->aft, r->tobe);           // :1880
  r->cds.clear();
  r->sds.clear();
  const int nkf = (int)r->kf_corner.size();
  r->map_ids.clear();
  if (nkf > 0 && r->loop_closure) {  // :1097, :1101-1151
    for (int k = std::max(0, nkf - r->window); k < nkf; ++k) r->map_ids.push_back(k);
  } else if (nkf > 0) {  // :1152-1216, the default branch
    std::vector<orc_p4> kp((size_t)nkf);
    for (int k = 0; k < nkf; ++k)  // cloudKeyPoses3D: position + intensity = index (:1692-1697)
      kp[(size_t)k] = orc_p4{r->kf_pose[6 * k + 3], r->kf_pose[6 * k + 4], r->kf_pose[6 * k + 5], (float)k};
    std::vector<int> ex = r->existing_ids;
    ex.resize((size_t)nkf + ex.size());
    const int n_ex = orc_select_surrounding(kp.data(), nkf, r->cur_pos, r->search_radius, ex.data(), (int)r->existing_ids.size(),
                                            (int)ex.size(), g_ext_radius != nullptr);
    ex.resize((size_t)n_ex);
    r->existing_ids = ex;
    r->map_ids = ex;
  }
  if (!r->map_ids.empty()) {
    std::vector<orc_p4> c, s, o;
    std::vector<int> co{0}, so{0}, oo{0};
    std::vector<float> poses;
    for (int k : r->map_ids) {
      c.insert(c.end(), r->kf_corner[k].begin(), r->kf_corner[k].end());
      s.insert(s.end(), r->kf_surf[k].begin(), r->kf_surf[k].end());
      o.insert(o.end(), r->kf_outlier[k].begin(), r->kf_outlier[k].end());
      co.push_back((int)c.size());
      so.push_back((int)s.size());
      oo.push_back((int)o.size());
      poses.insert(poses.end(), &r->kf_pose[6 * k], &r->kf_pose[6 * k] + 6);
    }
    build_local_map((int)r->map_ids.size(), poses.data(), c.data(), co.data(), s.data(), so.data(), o.data(), oo.data(),
                    r->prm, r->cds, r->sds, nullptr, nullptr);
  }
  if (n_corner_map_ds) *n_corner_map_ds = (int)r->cds.size();
  if (n_surf_map_ds) *n_surf_map_ds = (int)r->sds.size();
  scan2map(r->cds.data(), (int)r->cds.size(), r->sds.data(), (int)r->sds.size(), corner_scan_ds, n_csd,
           surf_total_ds, n_st, r->tobe, r->prm, r->st, *res);               // :1886
  if (res->status == 0) {                                                     // transformUpdate :583-589
    for (int i = 0; i < 6; ++i) { r->bef[i] = r->sum[i]; r->aft[i] = r->tobe[i]; }
  }
  for (int i = 0; i < 3; ++i) r->cur_pos[i] = r->aft[3 + i];  // currentRobotPosPoint, :1613-1615
  // saveKeyFramesAndFactor :1612-1755 with the pose graph as pass-through: the key pose is
  // transformAftMapped; the keyframe clouds are corner_scan (un-downsampled, :1746),
  // surf_last_ds (:1744) and outlier_last_ds (:1745).
  r->kf_corner.emplace_back(corner_scan, corner_scan + n_cs);
  r->kf_surf.emplace_back(surf_last_ds, surf_last_ds + n_sl);
  r->kf_outlier.emplace_back(outlier_last_ds, outlier_last_ds + n_ol);
  // First key pose = transformTobeMapped (:1640-1652); later ones = transformAftMapped (:1660-1670),
  // after which transformTobeMapped = transformAftMapped (:1730-1733, only once > 1 key poses exist).
  const bool first_kf = r->kf_pose.empty();
  const float* key_pose = first_kf ? r->tobe : r->aft;
  r->kf_pose.insert(r->kf_pose.end(), key_pose, key_pose + 6);
  if (!first_kf)
    for (int i = 0; i < 6; ++i) r->tobe[i] = r->aft[i];
}

void orc_replay_get_map(orc_replay* r, int which, orc_p4* out, int cap, int* n) {
  const std::vector<orc_p4>& v = which == 0 ? r->cds : r->sds;
  *n = (int)v.size();
  std::memcpy(out, v.data(), std::min<size_t>(cap, v.size()) * sizeof(orc_p4));
}


// ---- f-3: FeatureAssociation correspondence search ---------------------------------------------------------------
// TransformToStart (featureAssociation.cpp:1389-1412). cos / sin are the unqualified C functions on float arguments:
// taken as the double functions (repo-wide convention), so every product below is evaluated in double
// and rounded to float on assignment.
static orc_p4 fa_transform_to_start(const orc_p4& pi, const float* tc) {
  const float s = 10 * (pi.i - (float)(int)pi.i);
  const float rx = s * tc[0], ry = s * tc[1], rz = s * tc[2];
  const float tx = s * tc[3], ty = s * tc[4], tz = s * tc[5];
  const float x1 = (float)(std::cos((double)rz) * (double)(pi.x - tx) + std::sin((double)rz) * (double)(pi.y - ty));
  const float y1 = (float)(-std::sin((double)rz) * (double)(pi.x - tx) + std::cos((double)rz) * (double)(pi.y - ty));
  const float z1 = (pi.z - tz);
  const float x2 = x1;
  const float y2 = (float)(std::cos((double)rx) * (double)y1 + std::sin((double)rx) * (double)z1);
  const float z2 = (float)(-std::sin((double)rx) * (double)y1 + std::cos((double)rx) * (double)z1);
  orc_p4 po;
  po.x = (float)(std::cos((double)ry) * (double)x2 - std::sin((double)ry) * (double)z2);
  po.y = y2;
  po.z = (float)(std::sin((double)ry) * (double)x2 + std::cos((double)ry) * (double)z2);
  po.i = pi.i;
  return po;
}

void orc_fa_transform_to_start(const orc_p4* in, int n, const float transform_cur[6], orc_p4* out) {
  for (int i = 0; i < n; ++i) out[i] = fa_transform_to_start(in[i], transform_cur);
}

static inline float fa_sqdist(const orc_p4& a, const orc_p4& p) {
  return (a.x - p.x) * (a.x - p.x) + (a.y - p.y) * (a.y - p.y) + (a.z - p.z) * (a.z - p.z);
}

void orc_fa_correspondences(int which, const orc_p4* q, int nq, const orc_p4* last, int n_last, const float transform_cur[6],
                            int iter, float nearest_feature_dist_sqr, int* ind1, int* ind2, int* ind3, float* coeff,
                            int* keep, int use_ext_knn) {
  void* tree = nullptr;
  if (use_ext_knn && g_ext_build && g_ext_knn5 && n_last >= 5) tree = g_ext_build(last, n_last);
  for (int i = 0; i < nq; ++i) {
    const orc_p4 sel = fa_transform_to_start(q[i], transform_cur);
    keep[i] = 0;
    coeff[4 * i] = coeff[4 * i + 1] = coeff[4 * i + 2] = coeff[4 * i + 3] = 0.f;
    if (iter % 5 == 0) {  // :1588 / :1721
      int closest = -1, min2 = -1, min3 = -1;
      float best = FLT_MAX;
      int besti = -1;
      if (tree) {  // kdtree*Last.nearestKSearch(pointSel, 1, ...)
        int idx5[5];
        float d5[5];
        const float v[3] = {sel.x, sel.y, sel.z};
        g_ext_knn5(tree, v, idx5, d5);  // five nearest, ascending: the first is nearestKSearch(pointSel, 1)
        besti = idx5[0];
        best = d5[0];
      } else {
        for (int j = 0; j < n_last; ++j) {  // exact 1-NN, ties to the lowest index
          float df = sel.x - last[j].x;
          float d = df * df;
          df = sel.y - last[j].y;
          d += df * df;
          df = sel.z - last[j].z;
          d += df * df;
          if (d < best) { best = d; besti = j; }
        }
      }
      if (besti >= 0 && best < nearest_feature_dist_sqr) {
        closest = besti;
        const int scan = (int)last[closest].i;
        float d2min = nearest_feature_dist_sqr, d3min = nearest_feature_dist_sqr;
        // NB the reference bounds the forward scan by the QUERY count (:1599 / :1742), not by the size of the cloud
        // it indexes; reproduced (and clamped to the cloud so that the restatement never reads out of bounds).
        const int fwd_end = nq < n_last ? nq : n_last;
        for (int j = closest + 1; j < fwd_end; ++j) {
          if ((double)(int)last[j].i > (double)scan + 2.5) break;
          const float d = fa_sqdist(last[j], sel);
          if (which == 0) {
            if ((int)last[j].i > scan && d < d2min) { d2min = d; min2 = j; }
          } else if ((int)last[j].i <= scan) {
            if (d < d2min) { d2min = d; min2 = j; }
          } else if (d < d3min) {
            d3min = d; min3 = j;
          }
        }
        for (int j = closest - 1; j >= 0; --j) {
          if ((double)(int)last[j].i < (double)scan - 2.5) break;
          const float d = fa_sqdist(last[j], sel);
          if (which == 0) {
            if ((int)last[j].i < scan && d < d2min) { d2min = d; min2 = j; }
          } else if ((int)last[j].i >= scan) {
            if (d < d2min) { d2min = d; min2 = j; }
          } else if (d < d3min) {
            d3min = d; min3 = j;
          }
        }
      }
      ind1[i] = closest;
      ind2[i] = min2;
      if (ind3) ind3[i] = min3;
    }
    if (which == 0) {
      if (ind2[i] < 0) continue;  // :1644
      const orc_p4 t1 = last[ind1[i]], t2 = last[ind2[i]];
      const float x0 = sel.x, y0 = sel.y, z0 = sel.z, x1 = t1.x, y1 = t1.y, z1 = t1.z, x2 = t2.x, y2 = t2.y, z2 = t2.z;
      const float m11 = ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1));
      const float m22 = ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1));
      const float m33 = ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1));
      const float a012 = (float)std::sqrt((double)(m11 * m11 + m22 * m22 + m33 * m33));
      const float l12 = (float)std::sqrt((double)((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2)));
      const float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
      const float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
      const float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
      const float ld2 = a012 / l12;
      float s = 1;
      if (iter >= 5) s = (float)(1 - 1.8 * std::fabs((double)ld2));
      if ((double)s > 0.1 && ld2 != 0) {
        coeff[4 * i] = s * la; coeff[4 * i + 1] = s * lb; coeff[4 * i + 2] = s * lc; coeff[4 * i + 3] = s * ld2;
        keep[i] = 1;
      }
    } else {
      if (ind2[i] < 0 || ind3[i] < 0) continue;  // :1785
      const orc_p4 t1 = last[ind1[i]], t2 = last[ind2[i]], t3 = last[ind3[i]];
      float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
      float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
      float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
      float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
      const float ps = (float)std::sqrt((double)(pa * pa + pb * pb + pc * pc));
      pa /= ps; pb /= ps; pc /= ps; pd /= ps;
      const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
      float s = 1;
      if (iter >= 5)
        s = (float)(1 - 1.8 * std::fabs((double)pd2) / std::sqrt(std::sqrt((double)(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z))));
      if ((double)s > 0.1 && pd2 != 0) {
        coeff[4 * i] = s * pa; coeff[4 * i + 1] = s * pb; coeff[4 * i + 2] = s * pc; coeff[4 * i + 3] = s * pd2;
        keep[i] = 1;
      }
    }
  }
  if (tree && g_ext_free) g_ext_free(tree);
}

}  // extern "C"

// ---- f-1 + f-3 together: FeatureAssociation::updateTransformation (featureAssociation.cpp:2505-2535) ----------------
// Row of matA and entry of matB of calculateTransformationSurf (:1884-1949): unknowns [rx, rz, ty]. sin / cos are the
// unqualified C functions on float arguments (double functions, rounded to float on assignment: DESIGN.md §3); every
// other expression is float arithmetic as written.
namespace {

struct FaTrig {
  float srx, crx, sry, cry, srz, crz, tx, ty, tz;
};

inline FaTrig fa_trig(const float* tc) {
  FaTrig t;
  t.srx = (float)std::sin((double)tc[0]);
  t.crx = (float)std::cos((double)tc[0]);
  t.sry = (float)std::sin((double)tc[1]);
  t.cry = (float)std::cos((double)tc[1]);
  t.srz = (float)std::sin((double)tc[2]);
  t.crz = (float)std::cos((double)tc[2]);
  t.tx = tc[3];
  t.ty = tc[4];
  t.tz = tc[5];
  return t;
}

struct FaSurfConst {  // :1898-1923
  float a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, b1, b2, b5, b6, c1, c2, c3, c4, c5, c6, c7, c8, c9, crx;
};

inline FaSurfConst fa_surf_const(const FaTrig& g) {
  FaSurfConst k;
  const float srx = g.srx, crx = g.crx, sry = g.sry, cry = g.cry, srz = g.srz, crz = g.crz, tx = g.tx, ty = g.ty, tz = g.tz;
  k.a1 = crx * sry * srz;
  k.a2 = crx * crz * sry;
  k.a3 = srx * sry;
  k.a4 = tx * k.a1 - ty * k.a2 - tz * k.a3;
  k.a5 = srx * srz;
  k.a6 = crz * srx;
  k.a7 = ty * k.a6 - tz * crx - tx * k.a5;
  k.a8 = crx * cry * srz;
  k.a9 = crx * cry * crz;
  k.a10 = cry * srx;
  k.a11 = tz * k.a10 + ty * k.a9 - tx * k.a8;
  k.b1 = -crz * sry - cry * srx * srz;
  k.b2 = cry * crz * srx - sry * srz;
  k.b5 = cry * crz - srx * sry * srz;
  k.b6 = cry * srz + crz * srx * sry;
  k.c1 = -k.b6;
  k.c2 = k.b5;
  k.c3 = tx * k.b6 - ty * k.b5;
  k.c4 = -crx * crz;
  k.c5 = crx * srz;
  k.c6 = ty * k.c5 + tx * -k.c4;
  k.c7 = k.b2;
  k.c8 = -k.b1;
  k.c9 = tx * -k.b2 - ty * -k.b1;
  k.crx = crx;
  return k;
}

inline void fa_surf_row(const FaSurfConst& k, const orc_p4& o, const float* c, float r[3], float& b) {  // :1926-1948
  r[0] = (-k.a1 * o.x + k.a2 * o.y + k.a3 * o.z + k.a4) * c[0] + (k.a5 * o.x - k.a6 * o.y + k.crx * o.z + k.a7) * c[1] +
         (k.a8 * o.x - k.a9 * o.y - k.a10 * o.z + k.a11) * c[2];
  r[1] = (k.c1 * o.x + k.c2 * o.y + k.c3) * c[0] + (k.c4 * o.x - k.c5 * o.y + k.c6) * c[1] + (k.c7 * o.x + k.c8 * o.y + k.c9) * c[2];
  r[2] = -k.b6 * c[0] + k.c4 * c[1] + k.b2 * c[2];
  b = (float)(-0.05 * (double)c[3]);
}

struct FaCornerConst {  // :2034-2044
  float b1, b2, b3, b4, b5, b6, b7, b8, c5, srx;
};

inline FaCornerConst fa_corner_const(const FaTrig& g) {
  FaCornerConst k;
  const float srx = g.srx, crx = g.crx, sry = g.sry, cry = g.cry, srz = g.srz, crz = g.crz, tx = g.tx, ty = g.ty, tz = g.tz;
  k.b1 = -crz * sry - cry * srx * srz;
  k.b2 = cry * crz * srx - sry * srz;
  k.b3 = crx * cry;
  k.b4 = tx * -k.b1 + ty * -k.b2 + tz * k.b3;
  k.b5 = cry * crz - srx * sry * srz;
  k.b6 = cry * srz + crz * srx * sry;
  k.b7 = crx * sry;
  k.b8 = tz * k.b7 - ty * k.b6 - tx * k.b5;
  k.c5 = crx * srz;
  k.srx = srx;
  return k;
}

inline void fa_corner_row(const FaCornerConst& k, const orc_p4& o, const float* c, float r[3], float& b) {  // :2050-2066
  r[0] = (k.b1 * o.x + k.b2 * o.y - k.b3 * o.z + k.b4) * c[0] + (k.b5 * o.x + k.b6 * o.y - k.b7 * o.z + k.b8) * c[2];
  r[1] = -k.b5 * c[0] + k.c5 * c[1] + k.b1 * c[2];
  r[2] = k.b7 * c[0] - k.srx * c[1] - k.b3 * c[2];
  b = (float)(-0.05 * (double)c[3]);
}

}  // namespace

extern "C" {

// One of the two loops of updateTransformation (:2508-2518 surf, :2521-2530 corner). Returns the loop counter at exit.
static int fa_loop(int which, const orc_p4* q, int nq, const orc_p4* last, int n_last, float tc[6], float dist_sqr,
                   int eigen_mode, int use_ext_knn, orc_fa_state* st, float* trace) {
  std::vector<int> ind1((size_t)std::max(nq, 1), -1), ind2(ind1), ind3(ind1), keep(ind1);
  std::vector<float> coeff((size_t)std::max(nq, 1) * 4), rows, rhs;
  float* matP = which ? st->matP_surf : st->matP_corner;
  int it = 0;
  for (; it < 100; ++it) {
    orc_fa_correspondences(which, q, nq, last, n_last, tc, it, dist_sqr, ind1.data(), ind2.data(), ind3.data(), coeff.data(),
                           keep.data(), use_ext_knn);
    rows.clear();
    rhs.clear();
    const FaTrig g = fa_trig(tc);
    const FaSurfConst ks = fa_surf_const(g);
    const FaCornerConst kc = fa_corner_const(g);
    for (int i = 0; i < nq; ++i) {
      if (!keep[i]) continue;
      float r[3], b;
      if (which) fa_surf_row(ks, q[i], &coeff[4 * (size_t)i], r, b);
      else fa_corner_row(kc, q[i], &coeff[4 * (size_t)i], r, b);
      rows.insert(rows.end(), {r[0], r[1], r[2], 0.f, 0.f, 0.f});
      rhs.push_back(b);
    }
    const int n = (int)rhs.size();
    if (trace) { trace[4 * it] = trace[4 * it + 1] = trace[4 * it + 2] = 0.f; trace[4 * it + 3] = (float)n; }
    if (n < 10) continue;  // :2514 / :2527
    const int cols[3] = {0, 1, 2};
    float x[3];
    solve_normal_eq_t<3>(rows.data(), rhs.data(), n, cols, it, 10.f, eigen_mode, x, matP, &st->is_degenerate);
    if (trace) { trace[4 * it] = x[0]; trace[4 * it + 1] = x[1]; trace[4 * it + 2] = x[2]; }
    float deltaR, deltaT;
    if (which) {  // :1992-2004
      tc[0] += x[0];
      tc[2] += x[1];
      tc[4] += x[2];
      const double r0 = (double)(57.29578f * x[0]), r1 = (double)(57.29578f * x[1]), t0 = (double)(x[2] * 100.f);
      deltaR = (float)std::sqrt(r0 * r0 + r1 * r1);
      deltaT = (float)std::sqrt(t0 * t0);
    } else {      // :2127-2137
      tc[1] += x[0];
      tc[3] += x[1];
      tc[5] += x[2];
      const double r0 = (double)(57.29578f * x[0]), t0 = (double)(x[1] * 100.f), t1 = (double)(x[2] * 100.f);
      deltaR = (float)std::sqrt(r0 * r0);
      deltaT = (float)std::sqrt(t0 * t0 + t1 * t1);
    }
    for (int k = 0; k < 6; ++k)
      if (std::isnan(tc[k])) tc[k] = 0.f;
    if ((double)deltaR < 0.1 && (double)deltaT < 0.1) break;  // calculateTransformation* returned false
  }
  return it;
}

void orc_fa_init_state(orc_fa_state* st) {  // isDegenerate = false (:223); matP is a function-local there (see oracle.h)
  std::memset(st, 0, sizeof(*st));
}

int orc_fa_update_transformation(const orc_p4* corner_sharp, int n_cs, const orc_p4* surf_flat, int n_sf,
                                 const orc_p4* corner_last, int n_cl, const orc_p4* surf_last, int n_sl,
                                 float transform_cur[6], float nearest_feature_dist_sqr, int eigen_mode, int use_ext_knn,
                                 orc_fa_state* st, int iters[2], float* trace_surf, float* trace_corner) {
  iters[0] = iters[1] = 0;
  if (n_cl < 10 || n_sl < 100) return 1;  // :2506
  iters[0] = fa_loop(1, surf_flat, n_sf, surf_last, n_sl, transform_cur, nearest_feature_dist_sqr, eigen_mode, use_ext_knn, st, trace_surf);
  iters[1] = fa_loop(0, corner_sharp, n_cs, corner_last, n_cl, transform_cur, nearest_feature_dist_sqr, eigen_mode, use_ext_knn, st, trace_corner);
  return 0;
}

}  // extern "C"
