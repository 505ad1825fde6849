/*
 * oracle/hamming_knn.c — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the exact matcher the path is checked against: cv::BFMatcher(NORM_HAMMING).knnMatch as
 * called at src/detection/DescriptorMatcher.cpp:211 (reference tree).  OpenCV is not vendored; its published
 * algorithm is: for every query, distance to every train row = popcount(a XOR b) over the 32 descriptor bytes, keep
 * the k smallest under the key (distance, imgIdx, trainIdx).  With the objects concatenated in imgIdx order the key
 * is (distance, global_row).  Pinned against cv2 4.13.0 outputs in tests/test_oracle_knn.py.
 *
 * keys[q*k + j] = (distance << 32) | global_row, ascending; counts[q] = min(k, ndb).
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_knn_hamming(const uint8_t *query, int64_t nq, const uint8_t *db, int64_t ndb, int k, int threads,
                       uint64_t *keys, int32_t *counts) {
  if (k < 1 || k > 64) return 1;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t q = 0; q < nq; ++q) {
    uint64_t a[4];
    memcpy(a, query + 32 * q, 32);
    uint64_t best[64];
    int nb = 0;
    for (int64_t r = 0; r < ndb; ++r) {
      uint64_t b[4];
      memcpy(b, db + 32 * r, 32);
      uint64_t d = (uint64_t)(__builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
                              __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]));
      uint64_t key = (d << 32) | (uint64_t)r;
      if (nb == k && key >= best[k - 1]) continue;
      int j = nb < k ? nb++ : k - 1;
      while (j > 0 && best[j - 1] > key) {
        best[j] = best[j - 1];
        --j;
      }
      best[j] = key;
    }
    for (int j = 0; j < k; ++j) keys[q * k + j] = j < nb ? best[j] : ~(uint64_t)0;
    counts[q] = nb;
  }
  return 0;
}
