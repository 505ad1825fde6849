// tod_b200.hpp — header-only C++ host layer over the C-ABI (tod_b200.h): the two cells of TOD's detection hot path with
// the reference's names, parameters, inputs and outputs, minus the ecto / OpenCV types.
//
//   tod_b200::DescriptorMatcher   <->  tod::DescriptorMatcher   src/detection/DescriptorMatcher.cpp:58-270
//   tod_b200::GuessGenerator      <->  tod::GuessGenerator      src/detection/GuessGenerator.cpp:69-276
//
// An ecto cell body is a few lines on top of these: copy cv::Mat / cv::DMatch / cv::KeyPoint
// contents in and out (tod_match == cv::DMatch and tod_keypoint == cv::KeyPoint field for field) and forward.
// Errors are std::runtime_error carrying tod_last_error() — the reference reports through C++ exceptions too (or
// std::terminate for an unknown search type, DescriptorMatcher.cpp:182-186; here that is an exception as well).
// Like the reference's cells, objects are not thread-safe; the ecto scheduler calls them serially.
#ifndef TOD_B200_HPP_
#define TOD_B200_HPP_

#include <algorithm>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tod_b200.h"

namespace tod_b200 {

inline void check(int rc) {
  if (rc != TOD_OK) throw std::runtime_error(std::string("tod_b200: ") + tod_last_error());
}

typedef std::string ObjectId;  // object_recognition_core::db::ObjectId

// One DB document of method "TOD" (DescriptorMatcher.cpp:70-86): descriptors N x 32 u8, points N x 3 f32.
struct Document {
  ObjectId object_id;
  const uint8_t *descriptors;
  const float *points;
  int32_t n;
};

class DescriptorMatcher {
 public:
  // outputs of process() (DescriptorMatcher.cpp:147-151, :246-249)
  struct Outputs {
    std::vector<std::vector<tod_match> > matches;        // vector<vector<cv::DMatch>>
    std::vector<std::vector<float> > matches_3d;         // per query 1 x m CV_32FC3, flattened xyz
    std::vector<ObjectId> object_ids;
    std::map<ObjectId, float> spans;
  };

  DescriptorMatcher() : h_(nullptr) { tod_matcher_default_params(&params_); }
  ~DescriptorMatcher() { tod_matcher_destroy(h_); }
  DescriptorMatcher(const DescriptorMatcher &) = delete;
  DescriptorMatcher &operator=(const DescriptorMatcher &) = delete;

  // configure(): params["search_json_params"] (DescriptorMatcher.cpp:159-181).  device / shard: this build's extras.
  void configure(const std::string &search_json_params, int device = 0, int shard_rank = 0, int shard_count = 1) {
    check(tod_matcher_params_from_json(search_json_params.c_str(), &params_));
    params_.device = device;
    params_.shard_rank = shard_rank;
    params_.shard_count = shard_count;
    tod_matcher_destroy(h_);
    h_ = nullptr;
    check(tod_matcher_create(&params_, &h_));
  }

  // parameter_callback(db_documents) (:60-129): replaces the whole model set and retrains.
  void parameter_callback(const std::vector<Document> &documents) {
    require_configured();
    check(tod_matcher_clear(h_));
    object_ids_.clear();
    for (size_t i = 0; i < documents.size(); ++i) {
      const Document &d = documents[i];
      check(tod_matcher_add_object(h_, d.object_id.c_str(), d.descriptors, d.points, d.n));
      object_ids_.push_back(d.object_id);
    }
    check(tod_matcher_train(h_));
    spans_.clear();
    for (int32_t i = 0; i < tod_matcher_num_objects(h_); ++i) spans_[object_ids_[size_t(i)]] = tod_matcher_span(h_, i);
  }

  // process(): inputs["descriptors"] nq x 32 u8 (:195-252).  Also fills the flat arrays GuessGenerator takes.
  const Outputs &process(const uint8_t *descriptors, int32_t nq) {
    require_configured();
    const int32_t k = tod_matcher_k(h_);
    flat_.assign(size_t(nq) * k, tod_match());
    counts_.assign(size_t(nq), 0);
    points3d_.assign(size_t(nq) * k * 3, 0.f);
    check(tod_matcher_knn(h_, descriptors, nq, flat_.data(), counts_.data(), points3d_.data()));
    out_.matches.assign(size_t(nq), std::vector<tod_match>());
    out_.matches_3d.assign(size_t(nq), std::vector<float>());
    for (int32_t q = 0; q < nq; ++q) {
      const size_t o = size_t(q) * k;
      out_.matches[size_t(q)].assign(flat_.begin() + o, flat_.begin() + o + counts_[size_t(q)]);
      out_.matches_3d[size_t(q)].assign(points3d_.begin() + o * 3, points3d_.begin() + (o + counts_[size_t(q)]) * 3);
    }
    out_.object_ids = object_ids_;
    out_.spans = spans_;
    return out_;
  }

  int32_t k() const { return h_ ? tod_matcher_k(h_) : params_.k; }
  const std::vector<tod_match> &flat_matches() const { return flat_; }
  const std::vector<int32_t> &counts() const { return counts_; }
  const std::vector<float> &flat_points3d() const { return points3d_; }
  const std::vector<ObjectId> &object_ids() const { return object_ids_; }
  std::vector<float> spans_by_index() const {
    std::vector<float> v;
    for (int32_t i = 0; h_ && i < tod_matcher_num_objects(h_); ++i) v.push_back(tod_matcher_span(h_, i));
    return v;
  }
  tod_matcher *handle() const { return h_; }

 private:
  void require_configured() const {
    if (!h_) throw std::runtime_error("tod_b200::DescriptorMatcher used before configure()");
  }
  tod_matcher_params params_;
  tod_matcher *h_;
  std::vector<ObjectId> object_ids_;
  std::map<ObjectId, float> spans_;
  std::vector<tod_match> flat_;
  std::vector<int32_t> counts_;
  std::vector<float> points3d_;
  Outputs out_;
};

class GuessGenerator {
 public:
  // outputs of process() (GuessGenerator.cpp:96-98): pose_results (R, T, object id), Rs, Ts
  struct PoseResult {
    float R[9];  // 3 x 3 row-major, object -> camera
    float T[3];
    ObjectId object_id;
    int32_t object_index;
    std::vector<int32_t> inlier_keypoints;
  };

  GuessGenerator() : h_(nullptr) { tod_guess_default_params(&params_); }
  ~GuessGenerator() { tod_guess_destroy(h_); }
  GuessGenerator(const GuessGenerator &) = delete;
  GuessGenerator &operator=(const GuessGenerator &) = delete;

  // configure(): params min_inliers, n_ransac_iterations, sensor_error (GuessGenerator.cpp:74-80, :101-120)
  void configure(unsigned min_inliers = 15, unsigned n_ransac_iterations = 1000, float sensor_error = 0.01f,
                 int device = 0, uint64_t seed = 0) {
    params_.min_inliers = min_inliers;
    params_.n_ransac_iterations = n_ransac_iterations;
    params_.sensor_error = sensor_error;
    params_.device = device;
    params_.seed = seed;
    tod_guess_destroy(h_);
    h_ = nullptr;
    check(tod_guess_create(&params_, &h_));
  }

  // process(): inputs keypoints, points3d (H x W x 3 f32, NaN where invalid), matches, matches_3d, spans, object_ids
  // (GuessGenerator.cpp:86-94, :127-250), taken straight from a DescriptorMatcher that processed the same frame.
  std::vector<PoseResult> process(const std::vector<tod_keypoint> &keypoints, const float *points3d, int32_t height,
                                  int32_t width, const DescriptorMatcher &matcher) {
    if (!h_) throw std::runtime_error("tod_b200::GuessGenerator used before configure()");
    const int32_t n = int32_t(keypoints.size());
    if (size_t(n) != matcher.counts().size()) throw std::runtime_error("keypoints and matches disagree in size");
    const std::vector<float> spans = matcher.spans_by_index();
    std::vector<tod_pose> poses(256);
    // a keypoint can be an inlier of one pose per object it matched: at most min(k, n_objects) times
    std::vector<int32_t> inl(size_t(n) * size_t(std::max<int32_t>(1, std::min<int32_t>(matcher.k(), int32_t(spans.size())))) + 1);
    int32_t n_poses = 0;
    check(tod_guess_process(h_, keypoints.data(), n, points3d, height, width, matcher.flat_matches().data(),
                            matcher.counts().data(), matcher.k(), matcher.flat_points3d().data(), spans.data(),
                            int32_t(spans.size()), poses.data(), int32_t(poses.size()), &n_poses, inl.data(),
                            int32_t(inl.size())));
    std::vector<PoseResult> out;
    out.resize(size_t(n_poses));
    size_t o = 0;
    for (int32_t i = 0; i < n_poses; ++i) {
      PoseResult &p = out[size_t(i)];
      for (int j = 0; j < 9; ++j) p.R[j] = poses[size_t(i)].R[j];
      for (int j = 0; j < 3; ++j) p.T[j] = poses[size_t(i)].T[j];
      p.object_index = poses[size_t(i)].object_index;
      p.object_id = matcher.object_ids()[size_t(p.object_index)];  // GuessGenerator.cpp:175-176, :228
      p.inlier_keypoints.assign(inl.begin() + o, inl.begin() + o + poses[size_t(i)].n_inliers);
      o += size_t(poses[size_t(i)].n_inliers);
    }
    return out;
  }

  tod_guess *handle() const { return h_; }

 private:
  tod_guess_params params_;
  tod_guess *h_;
};

}  // namespace tod_b200
#endif  // TOD_B200_HPP_
