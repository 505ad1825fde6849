// C++ host code driving the hot path through include/tod_b200.hpp (the cell-shaped layer over the C-ABI), the way the
// reference's ecto cells would: configure -> parameter_callback -> process (DescriptorMatcher) -> process
// (GuessGenerator).  Checks the matches against an in-test brute force (cv::BFMatcher semantics: k smallest by
// (distance, imgIdx, trainIdx), radius cut) and the recovered pose against the planted one.  Prints "OK".
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <tuple>

#include "tod_b200.hpp"

static uint64_t g_state = 0x70D;
static uint32_t rnd() {
  g_state = g_state * 6364136223846793005ull + 1442695040888963407ull;
  return uint32_t(g_state >> 33);
}
static float unit() { return float(rnd() & 0xFFFFFF) / float(0x1000000); }

int main() {
  const int n_obj = 3, rows = 400, H = 480, W = 640, K = 5;
  const unsigned radius = 35;
  std::vector<std::vector<uint8_t> > desc(n_obj, std::vector<uint8_t>(size_t(rows) * 32));
  std::vector<std::vector<float> > pts(n_obj, std::vector<float>(size_t(rows) * 3));
  for (int o = 0; o < n_obj; ++o) {
    for (auto &b : desc[o]) b = uint8_t(rnd());
    for (auto &v : pts[o]) v = (unit() - 0.5f) * 0.14f;
  }
  // planted pose of object 1: rotation about z then x, translation in front of the camera
  const double az = 0.4, ax = -0.25;
  const double Rz[9] = {std::cos(az), -std::sin(az), 0, std::sin(az), std::cos(az), 0, 0, 0, 1};
  const double Rx[9] = {1, 0, 0, 0, std::cos(ax), -std::sin(ax), 0, std::sin(ax), std::cos(ax)};
  double R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rx[i * 3] * Rz[j] + Rx[i * 3 + 1] * Rz[3 + j] + Rx[i * 3 + 2] * Rz[6 + j];
  const double T[3] = {0.05, -0.02, 0.9};
  std::vector<float> cloud(size_t(H) * W * 3, std::numeric_limits<float>::quiet_NaN());
  std::vector<tod_keypoint> keypoints;
  std::vector<uint8_t> query;
  const double f = 525.0, cx = (W - 1) / 2.0, cy = (H - 1) / 2.0;
  for (int i = 0; i < rows; ++i) {
    const float *p = &pts[1][size_t(i) * 3];
    double c[3];
    for (int r = 0; r < 3; ++r) c[r] = R[r * 3] * p[0] + R[r * 3 + 1] * p[1] + R[r * 3 + 2] * p[2] + T[r];
    const int x = int(f * c[0] / c[2] + cx), y = int(f * c[1] / c[2] + cy);
    if (x < 0 || x >= W || y < 0 || y >= H || !std::isnan(cloud[(size_t(y) * W + x) * 3])) continue;
    for (int r = 0; r < 3; ++r) cloud[(size_t(y) * W + x) * 3 + r] = float(c[r]);
    tod_keypoint kp = {float(x) + 0.4f, float(y) + 0.4f, 31.f, 0.f, 1.f, 0, -1};
    keypoints.push_back(kp);
    for (int b = 0; b < 32; ++b) {
      uint8_t v = desc[1][size_t(i) * 32 + b];
      if ((rnd() & 7) == 0) v ^= uint8_t(1u << (rnd() & 7));  // ~4 flipped bits per descriptor
      query.push_back(v);
    }
  }
  for (int i = 0; i < 150; ++i) {  // clutter: random descriptors on random free pixels
    const int x = int(rnd() % W), y = int(rnd() % H);
    if (!std::isnan(cloud[(size_t(y) * W + x) * 3])) continue;
    cloud[(size_t(y) * W + x) * 3] = 0.3f * unit();
    cloud[(size_t(y) * W + x) * 3 + 1] = 0.3f * unit();
    cloud[(size_t(y) * W + x) * 3 + 2] = 0.8f + unit();
    tod_keypoint kp = {float(x) + 0.5f, float(y) + 0.5f, 31.f, 0.f, 1.f, 0, -1};
    keypoints.push_back(kp);
    for (int b = 0; b < 32; ++b) query.push_back(uint8_t(rnd()));
  }
  const int nq = int(keypoints.size());

  try {
    tod_b200::DescriptorMatcher matcher;
    // conf/detection.ork `search:` subtree as ORK core serialises it
    matcher.configure("{\"type\": \"LSH\", \"module\": \"ecto_opencv.features2d\", \"key_size\": 16, "
                      "\"multi_probe_level\": 1, \"n_tables\": 10, \"radius\": 35, \"ratio\": 0.8}");
    std::vector<tod_b200::Document> docs;
    for (int o = 0; o < n_obj; ++o) {
      tod_b200::Document d = {"object_" + std::to_string(o), desc[o].data(), pts[o].data(), rows};
      docs.push_back(d);
    }
    matcher.parameter_callback(docs);
    const tod_b200::DescriptorMatcher::Outputs &out = matcher.process(query.data(), nq);

    // brute-force check of every match list
    for (int q = 0; q < nq; ++q) {
      std::vector<std::tuple<int, int, int> > all;
      for (int o = 0; o < n_obj; ++o)
        for (int r = 0; r < rows; ++r) {
          int d = 0;
          for (int b = 0; b < 32; ++b) d += __builtin_popcount(unsigned(query[size_t(q) * 32 + b] ^ desc[o][size_t(r) * 32 + b]));
          all.emplace_back(d, o, r);
        }
      std::partial_sort(all.begin(), all.begin() + K, all.end());
      size_t want = 0;
      while (want < size_t(K) && unsigned(std::get<0>(all[want])) <= radius) ++want;
      if (out.matches[size_t(q)].size() != want) { printf("FAIL: query %d has %zu matches, expected %zu\n", q, out.matches[size_t(q)].size(), want); return 1; }
      for (size_t j = 0; j < want; ++j) {
        const tod_match &m = out.matches[size_t(q)][j];
        if (m.queryIdx != q || int(m.distance) != std::get<0>(all[j]) || m.imgIdx != std::get<1>(all[j]) || m.trainIdx != std::get<2>(all[j])) {
          printf("FAIL: query %d match %zu differs from the brute force\n", q, j);
          return 1;
        }
        const float *e = &pts[size_t(m.imgIdx)][size_t(m.trainIdx) * 3];
        const float *g = &out.matches_3d[size_t(q)][j * 3];
        if (e[0] != g[0] || e[1] != g[1] || e[2] != g[2]) { printf("FAIL: matches_3d of query %d\n", q); return 1; }
      }
    }
    if (out.object_ids.size() != size_t(n_obj) || out.spans.size() != size_t(n_obj)) { printf("FAIL: ids/spans\n"); return 1; }

    tod_b200::GuessGenerator guess;
    guess.configure(15, 500, 0.01f);
    std::vector<tod_b200::GuessGenerator::PoseResult> poses = guess.process(keypoints, cloud.data(), H, W, matcher);
    if (poses.size() != 1 || poses[0].object_id != "object_1") { printf("FAIL: %zu poses\n", poses.size()); return 1; }
    double err = 0;
    for (int i = 0; i < 9; ++i) err = std::max(err, std::fabs(double(poses[0].R[i]) - R[i]));
    for (int i = 0; i < 3; ++i) err = std::max(err, std::fabs(double(poses[0].T[i]) - T[i]));
    if (!(err < 1e-3)) { printf("FAIL: pose error %g\n", err); return 1; }
    printf("OK %d queries, %zu inlier keypoints, pose error %.2e\n", nq, poses[0].inlier_keypoints.size(), err);
  } catch (const std::exception &e) {
    printf("FAIL: %s\n", e.what());
    return 1;
  }
  return 0;
}
