// FeatureAssociation's correspondence search on the device —
//   findCorrespondingCornerFeatures (featureAssociation.cpp:1580-1697) and findCorrespondingSurfFeatures (:1699-1844)
//   with TransformToStart (:1389-1412) inside.
// One warp per feature point of the current sweep:
//   * TransformToStart (de-skew by the fraction of the sweep stored behind the decimal point of `intensity`);
//   * every fifth iteration: the exact nearest neighbour in laserCloud{Corner,Surf}Last (what kdtree*Last.nearestKSearch(.,1)
//     returns) — the clouds hold a few thousand points, so the lanes simply stride over all of them and a redux.sync
//     pair picks the minimum by (d2, index); the ring-window walks around it (+-2.5 scan lines, by the integer part of
//     `intensity`; ordered: strict '<' with break) advance 32 points per step (fa_window_scan), with the reference's
//     forward bound by the QUERY count (:1599, :1742 — an upstream quirk, clamped to the cloud);
//   * the point-to-line / point-to-plane residual, weight and keep test of the (laserCloudOri, coeffSel) pair.
// Arithmetic follows DESIGN.md §3: -fmad=false, unqualified cos / sin / sqrt / fabs on floats are the double functions.
#pragma once
#include <float.h>
#include "common.cuh"

namespace ll {

constexpr int FA_THREADS = 256;

struct FaArgs {
  const float4* q;     // cornerPointsSharp / surfPointsFlat
  const float4* last;  // laserCloudCornerLast / laserCloudSurfLast
  int nq, n_last;
  float tc[6];         // transformCur
  int which;           // 0 corner, 1 surf
  int iter;
  float dist_sqr;      // _nearest_feature_dist_sqr
  int* ind1;           // pointSearch*Ind1/2/3 [nq]
  int* ind2;
  int* ind3;
  float* coeff;        // [nq][4]
  int* keep;           // [nq]
};

__device__ __forceinline__ float4 fa_transform_to_start(const float4 pi, const float* tc) {
  const float s = 10 * (pi.w - (float)(int)pi.w);
  const float rx = s * tc[0], ry = s * tc[1], rz = s * tc[2];
  const float tx = s * tc[3], ty = s * tc[4], tz = s * tc[5];
  double srx, crx, sry, cry, srz, crz;
  sincos((double)rx, &srx, &crx);
  sincos((double)ry, &sry, &cry);
  sincos((double)rz, &srz, &crz);
  const float x1 = (float)(crz * (double)(pi.x - tx) + srz * (double)(pi.y - ty));
  const float y1 = (float)(-srz * (double)(pi.x - tx) + crz * (double)(pi.y - ty));
  const float z1 = (pi.z - tz);
  const float x2 = x1;
  const float y2 = (float)(crx * (double)y1 + srx * (double)z1);
  const float z2 = (float)(-srx * (double)y1 + crx * (double)z1);
  float4 po;
  po.x = (float)(cry * (double)x2 - sry * (double)z2);
  po.y = y2;
  po.z = (float)(sry * (double)x2 + cry * (double)z2);
  po.w = pi.w;
  return po;
}

__device__ __forceinline__ float fa_sqdist(const float4 a, const float4 p) {
  return (a.x - p.x) * (a.x - p.x) + (a.y - p.y) * (a.y - p.y) + (a.z - p.z) * (a.z - p.z);
}

// One direction of the ring-window walk around the nearest neighbour (forward :1599-1616 / :1742-1762, backward
// :1617-1633 / :1764-1784). The reference walks the cloud point by point, stops at the first point more than 2.5 scan
// lines away, and keeps the first strict minimum per category; the running minima carry over from the forward to the
// backward walk. The same result in warp-wide steps: 32 points per step in walking order, a ballot finds the first
// stopping point, every lane keeps its own first strict minimum, and the warp takes the minimum by (d2, walking
// index) — the first of the smallest, which replaces the carried minimum only if strictly smaller.
// Loads of the previous sweep's cloud: the read-only global path, or plain loads when the cloud was staged in shared memory.
template <bool kSmem>
__device__ __forceinline__ float4 fa_ld(const float4* p) {
  if constexpr (kSmem) return *p;
  else return __ldg(p);
}

template <bool kForward, bool kSmem = false>
__device__ __forceinline__ void fa_window_scan(const FaArgs& A, const float4 sel, const int closest, const int scan,
                                               const int lane, float& d2min, int& min2, float& d3min, int& min3) {
  // forward bound by the QUERY count (:1599, :1742 — an upstream quirk, clamped to the cloud)
  const int count = kForward ? min(A.nq, A.n_last) - (closest + 1) : closest;
  uint32_t bd2 = 0x7f800000u, be2 = 0xffffffffu, bd3 = 0x7f800000u, be3 = 0xffffffffu;
  for (int base = 0; base < count; base += 32) {  // (warp-uniform)
    const int e = base + lane;
    const bool in = e < count;
    bool stop = false;
    float d = 0.f;
    int sj = 0;
    if (in) {
      const float4 m = fa_ld<kSmem>(A.last + (kForward ? closest + 1 + e : closest - 1 - e));
      sj = (int)m.w;
      stop = kForward ? ((double)sj > (double)scan + 2.5) : ((double)sj < (double)scan - 2.5);
      d = fa_sqdist(m, sel);
    }
    const unsigned stops = __ballot_sync(0xffffffffu, stop);
    const int first_stop = stops ? __ffs((int)stops) - 1 : 32;
    if (in && lane < first_stop) {
      const uint32_t db = __float_as_uint(d);  // d >= +0, or NaN whose bit patterns are above +inf and never win
      bool c2, c3;
      if (A.which == 0) {  // corner: the second line point lies on ANOTHER scan line
        c2 = kForward ? sj > scan : sj < scan;
        c3 = false;
      } else {             // surf: the same or the nearer-side line -> second point, the far side -> third point
        c2 = kForward ? sj <= scan : sj >= scan;
        c3 = !c2;
      }
      if (c2 && db < bd2) { bd2 = db; be2 = (uint32_t)e; }
      if (c3 && db < bd3) { bd3 = db; be3 = (uint32_t)e; }
    }
    if (stops) break;
  }
  const uint32_t m2 = __reduce_min_sync(0xffffffffu, bd2);
  const uint32_t e2 = __reduce_min_sync(0xffffffffu, bd2 == m2 ? be2 : 0xffffffffu);
  if (e2 != 0xffffffffu && __uint_as_float(m2) < d2min) {
    d2min = __uint_as_float(m2);
    min2 = kForward ? closest + 1 + (int)e2 : closest - 1 - (int)e2;
  }
  if (A.which != 0) {
    const uint32_t m3 = __reduce_min_sync(0xffffffffu, bd3);
    const uint32_t e3 = __reduce_min_sync(0xffffffffu, bd3 == m3 ? be3 : 0xffffffffu);
    if (e3 != 0xffffffffu && __uint_as_float(m3) < d3min) {
      d3min = __uint_as_float(m3);
      min3 = kForward ? closest + 1 + (int)e3 : closest - 1 - (int)e3;
    }
  }
}

// The search of one feature point by one warp (every fifth iteration): nearestKSearch(pointSel, 1) over the previous
// sweep's cloud, then the two ring-window walks. All 32 lanes; the results are warp-uniform.
template <bool kSmem = false>
__device__ __forceinline__ void fa_search_warp(const FaArgs& A, const float4 sel, const int lane, int& closest, int& min2,
                                               int& min3) {
  closest = -1;
  min2 = -1;
  min3 = -1;
  // exact 1-NN: distance as nanoflann's L2_Simple_Adaptor (((dx^2) + dy^2) + dz^2), ties to the lowest index
  uint32_t bd = 0x7f800000u;  // +inf
  uint32_t bj = 0xffffffffu;
  auto consider = [&](const float4 m, const int j) {
    float df = sel.x - m.x;
    float d = df * df;
    df = sel.y - m.y;
    d += df * df;
    df = sel.z - m.z;
    d += df * df;
    const uint32_t db = __float_as_uint(d);
    if (db < bd) { bd = db; bj = (uint32_t)j; }  // (NaN distances have bit patterns above +inf and never win)
  };
  int j = lane;
  for (; j + 96 < A.n_last; j += 128) {  // four loads in flight, candidates taken in ascending index order
    const float4 m0 = fa_ld<kSmem>(A.last + j), m1 = fa_ld<kSmem>(A.last + j + 32), m2 = fa_ld<kSmem>(A.last + j + 64),
                 m3 = fa_ld<kSmem>(A.last + j + 96);
    consider(m0, j);
    consider(m1, j + 32);
    consider(m2, j + 64);
    consider(m3, j + 96);
  }
  for (; j < A.n_last; j += 32) consider(fa_ld<kSmem>(A.last + j), j);
  const uint32_t md = __reduce_min_sync(0xffffffffu, bd);
  const uint32_t mj = __reduce_min_sync(0xffffffffu, bd == md ? bj : 0xffffffffu);
  if (mj != 0xffffffffu && __uint_as_float(md) < A.dist_sqr) {  // (warp-uniform)
    closest = (int)mj;
    const int scan = (int)fa_ld<kSmem>(A.last + closest).w;
    float d2min = A.dist_sqr, d3min = A.dist_sqr;
    fa_window_scan<true, kSmem>(A, sel, closest, scan, lane, d2min, min2, d3min, min3);
    fa_window_scan<false, kSmem>(A, sel, closest, scan, lane, d2min, min2, d3min, min3);
  }
}

// The (laserCloudOri, coeffSel) pair of one feature point from its tripod (:1644-1693 corner, :1785-1840 surf), one
// thread. Returns 1 when the pair is pushed (c = the coeffSel record), 0 otherwise (c = 0).
template <bool kSmem = false>
__device__ __forceinline__ int fa_record(const FaArgs& A, const float4 sel, const int closest, const int min2, const int min3,
                                         float (&c)[4]) {
  c[0] = c[1] = c[2] = c[3] = 0.f;
  int kp = 0;
  if (A.which == 0) {
    if (min2 >= 0) {
      const float4 t1 = fa_ld<kSmem>(A.last + closest), t2 = fa_ld<kSmem>(A.last + min2);
      const float x0 = sel.x, y0 = sel.y, z0 = sel.z, x1 = t1.x, y1 = t1.y, z1 = t1.z, x2 = t2.x, y2 = t2.y, z2 = t2.z;
      const float m11 = ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1));
      const float m22 = ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1));
      const float m33 = ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1));
      const float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);  // sqrt(double(float)) rounded == sqrtf
      const float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
      const float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
      const float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
      const float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
      const float ld2 = a012 / l12;
      float s = 1;
      if (A.iter >= 5) s = (float)(1 - 1.8 * fabs((double)ld2));
      if ((double)s > 0.1 && ld2 != 0) {
        c[0] = s * la; c[1] = s * lb; c[2] = s * lc; c[3] = s * ld2;
        kp = 1;
      }
    }
  } else if (min2 >= 0 && min3 >= 0) {
    const float4 t1 = fa_ld<kSmem>(A.last + closest), t2 = fa_ld<kSmem>(A.last + min2), t3 = fa_ld<kSmem>(A.last + min3);
    float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
    float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
    float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
    float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
    const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
    const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
    float s = 1;
    if (A.iter >= 5)
      s = (float)(1 - 1.8 * fabs((double)pd2) / sqrt(sqrt((double)(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z))));
    if ((double)s > 0.1 && pd2 != 0) {
      c[0] = s * pa; c[1] = s * pb; c[2] = s * pc; c[3] = s * pd2;
      kp = 1;
    }
  }
  return kp;
}

__global__ void __launch_bounds__(FA_THREADS) fa_corr_kernel(const __grid_constant__ FaArgs A) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (FA_THREADS / 32) + (threadIdx.x >> 5);
  if (i >= A.nq) return;  // warp-uniform
  const float4 sel = fa_transform_to_start(__ldg(A.q + i), A.tc);
  int closest = -1, min2 = -1, min3 = -1;
  if (A.iter % 5 == 0) {
    fa_search_warp(A, sel, lane, closest, min2, min3);
    if (lane == 0) {
      A.ind1[i] = closest;
      A.ind2[i] = min2;
      if (A.ind3) A.ind3[i] = min3;
    }
  } else if (lane == 0) {
    closest = A.ind1[i];
    min2 = A.ind2[i];
    min3 = A.ind3 ? A.ind3[i] : -1;
  }
  if (lane != 0) return;
  float c[4];
  const int kp = fa_record(A, sel, closest, min2, min3, c);
#pragma unroll
  for (int k = 0; k < 4; ++k) A.coeff[4 * i + k] = c[k];
  A.keep[i] = kp;
}

}  // namespace ll
