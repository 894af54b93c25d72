// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Thin C shim over the REFERENCE'S OWN k-NN code: the vendored nanoflann 1.3.0 header
// /root/reference/LeGO-LOAM/include/lego_loam/nanoflann.hpp, compiled from where it lies (the
// Makefile passes -I to that directory; no reference source is copied into this repo). Output goes
// to oracle/_ref/libnanoflann_ref.so. The adaptor below exposes the same three methods as the
// reference's PCL adaptor (nanoflann_pcl.h:91-98), the tree type is the one at
// nanoflann_pcl.h:100-102, and the query is the body of nearestKSearch (nanoflann_pcl.h:140-152).
// nanoflann_pcl.h itself cannot be used: it includes PCL headers (:45-46), which are not installed.
#include <nanoflann.hpp>

#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

namespace {

struct P4 { float x, y, z, i; };

struct Adaptor {
  const P4* pts;
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, int dim) const {
    const P4& p = pts[idx];
    if (dim == 0) return p.x;
    else if (dim == 1) return p.y;
    else if (dim == 2) return p.z;
    else return 0.0;
  }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};

typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::SO3_Adaptor<float, Adaptor>, Adaptor, 3, int> Tree;

struct Index {
  Adaptor adaptor;
  Tree tree;
  Index(const P4* p, size_t n) : adaptor{p, n}, tree(3, adaptor) { tree.buildIndex(); }
};

}  // namespace

extern "C" {

void* ref_kdtree_build(const void* pts, int n) { return new Index((const P4*)pts, (size_t)n); }

void ref_kdtree_knn5(void* tree, const float q[3], int* idx5, float* d2_5) {
  Index* ix = (Index*)tree;
  nanoflann::KNNResultSet<float, int> rs(5);
  rs.init(idx5, d2_5);
  ix->tree.findNeighbors(rs, q, nanoflann::SearchParams());
  for (size_t k = rs.size(); k < 5; ++k) { idx5[k] = -1; }
}

// Body of KdTreeFLANN::radiusSearch (nanoflann_pcl.h:154-174) with its default _params (sorted = true).
int ref_kdtree_radius(void* tree, const float q[3], double radius, int* idx, float* d2, int cap) {
  Index* ix = (Index*)tree;
  std::vector<std::pair<int, float>> indices_dist;
  indices_dist.reserve(128);
  nanoflann::RadiusResultSet<float, int> resultSet(static_cast<float>(radius * radius), indices_dist);
  nanoflann::SearchParams params;
  ix->tree.findNeighbors(resultSet, q, params);
  const size_t nFound = resultSet.size();
  if (params.sorted) std::sort(indices_dist.begin(), indices_dist.end(), nanoflann::IndexDist_Sorter());
  for (size_t i = 0; i < nFound && (int)i < cap; i++) {
    idx[i] = indices_dist[i].first;
    d2[i] = indices_dist[i].second;
  }
  return (int)nFound;
}

void ref_kdtree_free(void* tree) { delete (Index*)tree; }

}  // extern "C"
