// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// C entry points around the reference's OWN geometry code: src/common/adjacency_ransac.cpp and maximum_clique.cpp are
// compiled unmodified from /root/reference (oracle/build_ref.py) and linked with this file into
// oracle/_ref/libtod_ref.so.  Nothing here re-implements the reference's algorithms; the only restated lines are the
// per-object driver loop of GuessGenerator::process (src/detection/GuessGenerator.cpp:170-235, an ecto cell that
// cannot compile without ecto) in ref_process().
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include <boost/foreach.hpp>
#include <boost/shared_ptr.hpp>
#include <opencv2/core/core.hpp>

// The harness needs to read private state (valid_indices_, samples_) of the reference classes; the class layout is
// unchanged by the access specifier.
#define private public
#define protected public
#include "adjacency_ransac.h"
#include "ransac.h"
#include "sac_model_registration_graph.h"
#undef private
#undef protected

// ---- sampler stream: same generator as tod_rng_* of the product (restated, not linked) -------------------------------
static uint64_t g_rng_state = 1;
extern "C" int tod_oracle_rand(void) {
  g_rng_state = g_rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return int(g_rng_state >> 33);
}
static uint64_t rng_seed(uint64_t seed, uint32_t object_index, uint32_t round) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t(object_index) + 1) + 0xBF58476D1CE4E5B9ull * (uint64_t(round) + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

using tod::AdjacencyRansac;
using tod::SampleConsensusModelRegistrationGraph;

// RandomSampleConsensus with a settable distance threshold: the reference never sets it (sac.h:70 leaves DBL_MAX),
// the protected member is reachable from a subclass without touching the reference files.
class ThresholdRansac : public pcl::RandomSampleConsensus {
 public:
  ThresholdRansac(const pcl::SampleConsensus::SampleConsensusModelPtr &m, double thr) : pcl::RandomSampleConsensus(m) {
    static_cast<pcl::SampleConsensus *>(this)->threshold_ = thr;
  }
};

extern "C" {

void ref_set_rng(uint64_t state) { g_rng_state = state; }
uint64_t ref_rng_seed(uint64_t seed, uint32_t object_index, uint32_t round) { return rng_seed(seed, object_index, round); }

// ---- maximum_clique::Graph ---------------------------------------------------------------------------------------
int ref_find_clique(int n, const int *edges, int n_edges, int sorted, const int *deleted, int n_deleted,
                    unsigned minimal_size, unsigned *out) {
  tod::maximum_clique::Graph g(n);
  for (int e = 0; e < n_edges; ++e) {
    if (sorted) g.AddEdgeSorted(edges[2 * e], edges[2 * e + 1]);
    else g.AddEdge(edges[2 * e], edges[2 * e + 1]);
  }
  for (int e = 0; e < n_deleted; ++e) g.DeleteEdge(deleted[2 * e], deleted[2 * e + 1]);
  tod::maximum_clique::Graph::Vertices v;
  if (minimal_size == 0xFFFFFFFFu) g.FindMaximumClique(v);
  else g.FindClique(v, minimal_size);
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
  return int(v.size());
}

// ---- AdjacencyRansac ---------------------------------------------------------------------------------------------
void *ref_ar_new() { return new AdjacencyRansac(); }
void ref_ar_free(void *h) { delete static_cast<AdjacencyRansac *>(h); }

void ref_ar_add(void *h, const float *train, const float *query, unsigned query_index) {
  static_cast<AdjacencyRansac *>(h)->AddPoints(cv::Vec3f(train[0], train[1], train[2]),
                                               cv::Vec3f(query[0], query[1], query[2]), query_index);
}

static std::vector<cv::KeyPoint> make_keypoints(const float *xy, int n) {
  std::vector<cv::KeyPoint> k(n);
  for (int i = 0; i < n; ++i) { k[i].pt.x = xy[2 * i]; k[i].pt.y = xy[2 * i + 1]; }
  return k;
}

void ref_ar_fill(void *h, const float *kp_xy, int n_kp, float span, float sensor_error) {
  static_cast<AdjacencyRansac *>(h)->FillAdjacency(make_keypoints(kp_xy, n_kp), span, sensor_error);
}

int ref_ar_size(void *h) { return int(static_cast<AdjacencyRansac *>(h)->query_indices().size()); }

int ref_ar_neighbors(void *h, int which, unsigned i, unsigned *out, int cap) {
  AdjacencyRansac *a = static_cast<AdjacencyRansac *>(h);
  const std::vector<unsigned> &r = (which ? a->sample_adjacency_ : a->physical_adjacency_).neighbors(i);
  for (size_t k = 0; k < r.size() && int(k) < cap; ++k) out[k] = r[k];
  return int(r.size());
}

int ref_ar_valid(void *h, unsigned *out, int cap) {
  AdjacencyRansac *a = static_cast<AdjacencyRansac *>(h);
  for (size_t k = 0; k < a->valid_indices_.size() && int(k) < cap; ++k) out[k] = a->valid_indices_[k];
  return int(a->valid_indices_.size());
}

void ref_ar_invalidate_query(void *h, const unsigned *q, int n) {
  std::vector<unsigned> v(q, q + n);
  static_cast<AdjacencyRansac *>(h)->InvalidateQueryIndices(v);
}

static SampleConsensusModelRegistrationGraph::Ptr make_model(AdjacencyRansac *a, float sensor_error) {
  return SampleConsensusModelRegistrationGraph::Ptr(new SampleConsensusModelRegistrationGraph(
      a->query_points_, a->training_points_, a->valid_indices_, sensor_error, a->physical_adjacency_,
      a->sample_adjacency_));
}

// The reference's own sampler (getSamples, sac_model_registration_graph.h:141-168) run n_hyp times on the stream
// `rng_state`: produces the shared hypothesis list.  Returns the number of triples produced (stops at the first empty).
int ref_ar_get_samples(void *h, uint64_t rng_state, int n_hyp, unsigned *triples) {
  AdjacencyRansac *a = static_cast<AdjacencyRansac *>(h);
  SampleConsensusModelRegistrationGraph::Ptr model = make_model(a, 0.01f);
  g_rng_state = rng_state;
  int it = 0, n = 0;
  std::vector<unsigned> sel;
  for (; n < n_hyp; ++n) {
    model->getSamples(it, sel);
    if (sel.size() != 3) break;
    for (int k = 0; k < 3; ++k) triples[3 * n + k] = sel[k];
  }
  return n;
}

// One RANSAC iteration body for a given triple: computeModelCoefficients + selectWithinDistance (gate included).
// pre_gate_count receives |inliers| before the clique gate (recomputed by calling with the gate disabled is not
// possible without editing the reference, so it is derived from the candidate set the same way the function does).
int ref_ar_select(void *h, const unsigned *triple, double threshold, unsigned *inliers_out, int cap, float *R9,
                  float *T3) {
  AdjacencyRansac *a = static_cast<AdjacencyRansac *>(h);
  SampleConsensusModelRegistrationGraph::Ptr model = make_model(a, 0.01f);
  std::vector<unsigned> sel(triple, triple + 3);
  model->samples_ = sel;
  cv::Matx33f R;
  cv::Vec3f T;
  if (!model->computeModelCoefficients(sel, R, T)) return -1;
  std::vector<unsigned> inl;
  model->selectWithinDistance(R, T, threshold, inl);
  for (size_t k = 0; k < inl.size() && int(k) < cap; ++k) inliers_out[k] = inl[k];
  if (R9) for (int i = 0; i < 9; ++i) R9[i] = R.val[i];
  if (T3) for (int i = 0; i < 3; ++i) T3[i] = T.val[i];
  return int(inl.size());
}

// estimateRigidTransformationSVD over an arbitrary index list (sac_model_registration_graph.h:304-347)
int ref_ar_kabsch(void *h, const unsigned *idx, int n, float *R9, float *T3) {
  AdjacencyRansac *a = static_cast<AdjacencyRansac *>(h);
  SampleConsensusModelRegistrationGraph::Ptr model = make_model(a, 0.01f);
  std::vector<unsigned> v(idx, idx + n);
  cv::Matx33f R;
  cv::Vec3f T;
  if (!model->estimateRigidTransformationSVD(v, R, T)) return -1;
  for (int i = 0; i < 9; ++i) R9[i] = R.val[i];
  for (int i = 0; i < 3; ++i) T3[i] = T.val[i];
  return 0;
}

// RANSAC loop only (pcl::RandomSampleConsensus::computeModel, ransac.h:80-143) with an optional finite threshold.
int ref_ar_compute_model(void *h, unsigned max_iterations, uint64_t rng_state, double threshold, unsigned *inliers_out,
                         int cap, int *iterations) {
  AdjacencyRansac *a = static_cast<AdjacencyRansac *>(h);
  SampleConsensusModelRegistrationGraph::Ptr model = make_model(a, 0.01f);
  ThresholdRansac sc(model, threshold);
  sc.setMaxIterations(int(max_iterations));
  g_rng_state = rng_state;
  sc.computeModel();
  std::vector<unsigned> inl;
  sc.getInliers(inl);
  for (size_t k = 0; k < inl.size() && int(k) < cap; ++k) inliers_out[k] = inl[k];
  if (iterations) *iterations = static_cast<pcl::SampleConsensus &>(sc).iterations_;
  return int(inl.size());
}

// AdjacencyRansac::Ransac (adjacency_ransac.cpp:234-309) on the stream `rng_state`.
int ref_ar_ransac(void *h, float sensor_error, unsigned n_iterations, uint64_t rng_state, unsigned *inliers_out, int cap,
                  float *R9, float *T3) {
  AdjacencyRansac *a = static_cast<AdjacencyRansac *>(h);
  g_rng_state = rng_state;
  std::vector<unsigned> inl;
  cv::Matx33f R;
  cv::Vec3f T;
  a->Ransac(sensor_error, n_iterations, inl, R, T);
  for (size_t k = 0; k < inl.size() && int(k) < cap; ++k) inliers_out[k] = inl[k];
  for (int i = 0; i < 9; ++i) R9[i] = R.val[i];
  for (int i = 0; i < 3; ++i) T3[i] = T.val[i];
  return int(inl.size());
}

struct RefMatch { int queryIdx, trainIdx, imgIdx; float distance; };
struct RefPose { float R[9]; float T[3]; int object_index; int n_inliers; };

// ClusterPerObject (adjacency_ransac.cpp:176-205) + the per-object loop of GuessGenerator::process
// (GuessGenerator.cpp:170-235), re-seeding the sampler stream per (object, round) as the product does.
int ref_process(const float *kp_xy, int n_kp, const float *cloud, int height, int width, const RefMatch *matches,
                const int *counts, int k, const float *points3d, const float *spans, int n_objects,
                unsigned min_inliers, unsigned n_ransac_iterations, float sensor_error, uint64_t seed, RefPose *poses,
                int max_poses, int *inlier_keypoints, int max_inlier_total) {
  std::vector<cv::KeyPoint> keypoints = make_keypoints(kp_xy, n_kp);
  cv::Mat point_cloud(height, width, 3);
  std::memcpy(point_cloud.d.data(), cloud, size_t(height) * width * 3 * sizeof(float));
  std::vector<std::vector<cv::DMatch> > m(n_kp);
  std::vector<cv::Mat> m3d(n_kp);
  for (int q = 0; q < n_kp; ++q) {
    m[q].resize(counts[q]);
    m3d[q] = cv::Mat(1, counts[q], 3);
    for (int j = 0; j < counts[q]; ++j) {
      const RefMatch &r = matches[size_t(q) * k + j];
      m[q][j].queryIdx = r.queryIdx; m[q][j].trainIdx = r.trainIdx; m[q][j].imgIdx = r.imgIdx; m[q][j].distance = r.distance;
      m3d[q].at<cv::Vec3f>(0, j) = cv::Vec3f(points3d[(size_t(q) * k + j) * 3], points3d[(size_t(q) * k + j) * 3 + 1],
                                             points3d[(size_t(q) * k + j) * 3 + 2]);
    }
  }
  tod::OpenCVIdToObjectPoints all;
  tod::ClusterPerObject(keypoints, point_cloud, m, m3d, all);
  int n_poses = 0, n_inl = 0;
  while (!all.empty()) {
    AdjacencyRansac &ar = all.begin()->second;
    const size_t object = all.begin()->first;
    if (int(object) >= n_objects) return -2;
    ar.FillAdjacency(keypoints, spans[object], sensor_error);
    unsigned round = 0;
    while (true) {
      std::vector<unsigned> query_inliers;
      cv::Matx33f R;
      cv::Vec3f T;
      g_rng_state = rng_seed(seed, uint32_t(object), round++);
      ar.Ransac(sensor_error, n_ransac_iterations, query_inliers, R, T);
      if (query_inliers.size() < min_inliers) break;
      ar.InvalidateQueryIndices(query_inliers);
      if (n_poses < max_poses) {
        RefPose &p = poses[n_poses];
        for (int i = 0; i < 9; ++i) p.R[i] = R.val[i];
        for (int i = 0; i < 3; ++i) p.T[i] = T.val[i];
        p.object_index = int(object);
        p.n_inliers = int(query_inliers.size());
        for (size_t i = 0; i < query_inliers.size() && n_inl < max_inlier_total; ++i) inlier_keypoints[n_inl++] = int(query_inliers[i]);
      }
      ++n_poses;
    }
    all.erase(object);
  }
  return n_poses;
}

}  // extern "C"
