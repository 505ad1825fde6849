// Error channel, ABI version and launch counter of libtod_b200.so.
#include <cstring>

#include "json_min.h"
#include "tod_internal.h"

namespace tod {

static thread_local char g_error[1024] = "";
std::atomic<uint64_t> g_kernel_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace tod

extern "C" {

const char *tod_last_error(void) { return tod::g_error; }
int tod_abi_version(void) { return TOD_B200_ABI_VERSION; }
uint64_t tod_kernel_launch_count(void) { return tod::g_kernel_launches.load(); }

// splitmix64-seeded 64-bit LCG, 31 output bits — the sampler stream of the guess generator (stands in for the
// reference's unseeded libc rand(), sac_model_registration_graph.h:111).
uint64_t tod_rng_seed(uint64_t seed, uint32_t object_index, uint32_t round) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t(object_index) + 1) + 0xBF58476D1CE4E5B9ull * (uint64_t(round) + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int32_t tod_rng_next(uint64_t *state) {
  *state = *state * 6364136223846793005ull + 1442695040888963407ull;
  return int32_t(*state >> 33);
}

}  // extern "C"
