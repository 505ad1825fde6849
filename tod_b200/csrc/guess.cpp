// C-ABI of the GuessGenerator half (include/tod_b200.h: tod_guess_*).  PLACEHOLDER — replaced by the full host
// driver (cluster -> K2 -> sampler -> K3 -> replay/gate/refine) in the next commit.
#include <cmath>
#include <cstring>
#include <limits>

#include "tod_internal.h"

struct tod_guess {
  tod_guess_params p{};
};

extern "C" {

void tod_guess_default_params(tod_guess_params *p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->min_inliers = 15;
  p->n_ransac_iterations = 1000;
  p->sensor_error = 0.01f;
  p->device = 0;
  p->ransac_threshold = std::numeric_limits<double>::infinity();
  p->seed = 0;
}

int tod_guess_create(const tod_guess_params *p, tod_guess **out) {
  TOD_REQUIRE(p && out, "null argument");
  tod_guess *g = new tod_guess();
  g->p = *p;
  *out = g;
  return TOD_OK;
}

void tod_guess_destroy(tod_guess *g) { delete g; }

int tod_guess_process(tod_guess *, const tod_keypoint *, int32_t, const float *, int32_t, int32_t, const tod_match *,
                      const int32_t *, int32_t, const float *, const float *, int32_t, tod_pose *, int32_t, int32_t *,
                      int32_t *, int32_t) {
  return tod::fail(TOD_ERR_STATE, "tod_guess_process: not built yet");
}

void tod_guess_last_stats(const tod_guess *, float *k2_ms, float *k3_ms, int64_t *n_hyp, int32_t *n_rounds) {
  if (k2_ms) *k2_ms = 0;
  if (k3_ms) *k3_ms = 0;
  if (n_hyp) *n_hyp = 0;
  if (n_rounds) *n_rounds = 0;
}

}  // extern "C"
