// Word-wise set operations of the host-side geometry code (clique gate, sampler) on 64-bit mask words, with AVX-512
// (VPOPCNTDQ) bodies picked once at run time: the inlier graphs are dense bit-rows of 12-64 words, and the gate and the
// sampler spend their time AND-ing and counting them.  Same results as the scalar loops, word for word.
#ifndef TOD_BITOPS_H_
#define TOD_BITOPS_H_

#include <cstdint>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace tod {
namespace bitops {

#if defined(__x86_64__)
inline bool have_avx512_popcnt() {
  static const bool yes = [] {
#if defined(TOD_BITOPS_SCALAR)   // variant builds only: the scalar loops, for A/B timing
    return false;
#endif
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") != 0 && __builtin_cpu_supports("avx512vpopcntdq") != 0;
  }();
  return yes;
}

#define TOD_AVX512 __attribute__((target("avx512f,avx512vpopcntdq")))

TOD_AVX512 inline int and_popcount_avx512(const uint64_t *a, const uint64_t *b, int n) {
  __m512i acc = _mm512_setzero_si512();
  int w = 0;
  for (; w + 8 <= n; w += 8)
    acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(_mm512_and_si512(_mm512_loadu_si512(a + w), _mm512_loadu_si512(b + w))));
  if (w < n) {
    const __mmask8 m = __mmask8((1u << (n - w)) - 1u);
    acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(_mm512_and_si512(_mm512_maskz_loadu_epi64(m, a + w),
                                                                     _mm512_maskz_loadu_epi64(m, b + w))));
  }
  return int(_mm512_reduce_add_epi64(acc));
}

TOD_AVX512 inline int and_store_popcount_avx512(uint64_t *dst, const uint64_t *a, const uint64_t *b, int n) {
  __m512i acc = _mm512_setzero_si512();
  int w = 0;
  for (; w + 8 <= n; w += 8) {
    const __m512i x = _mm512_and_si512(_mm512_loadu_si512(a + w), _mm512_loadu_si512(b + w));
    _mm512_storeu_si512(dst + w, x);
    acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(x));
  }
  if (w < n) {
    const __mmask8 m = __mmask8((1u << (n - w)) - 1u);
    const __m512i x = _mm512_and_si512(_mm512_maskz_loadu_epi64(m, a + w), _mm512_maskz_loadu_epi64(m, b + w));
    _mm512_mask_storeu_epi64(dst + w, m, x);
    acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(x));
  }
  return int(_mm512_reduce_add_epi64(acc));
}

// dst = a & ~b ; returns the number of set bits of the result
TOD_AVX512 inline int andnot_store_popcount_avx512(uint64_t *dst, const uint64_t *a, const uint64_t *b, int n) {
  __m512i acc = _mm512_setzero_si512();
  int w = 0;
  for (; w + 8 <= n; w += 8) {
    const __m512i x = _mm512_andnot_si512(_mm512_loadu_si512(b + w), _mm512_loadu_si512(a + w));
    _mm512_storeu_si512(dst + w, x);
    acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(x));
  }
  if (w < n) {
    const __mmask8 m = __mmask8((1u << (n - w)) - 1u);
    const __m512i x = _mm512_andnot_si512(_mm512_maskz_loadu_epi64(m, b + w), _mm512_maskz_loadu_epi64(m, a + w));
    _mm512_mask_storeu_epi64(dst + w, m, x);
    acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(x));
  }
  return int(_mm512_reduce_add_epi64(acc));
}

// dst = a & ~b ; returns the OR of all result words
TOD_AVX512 inline uint64_t andnot_store_any_avx512(uint64_t *dst, const uint64_t *a, const uint64_t *b, int n) {
  __m512i any = _mm512_setzero_si512();
  int w = 0;
  for (; w + 8 <= n; w += 8) {
    const __m512i x = _mm512_andnot_si512(_mm512_loadu_si512(b + w), _mm512_loadu_si512(a + w));
    _mm512_storeu_si512(dst + w, x);
    any = _mm512_or_si512(any, x);
  }
  if (w < n) {
    const __mmask8 m = __mmask8((1u << (n - w)) - 1u);
    const __m512i x = _mm512_andnot_si512(_mm512_maskz_loadu_epi64(m, b + w), _mm512_maskz_loadu_epi64(m, a + w));
    _mm512_mask_storeu_epi64(dst + w, m, x);
    any = _mm512_or_si512(any, x);
  }
  return uint64_t(_mm512_reduce_or_epi64(any));
}
#endif

// sum of popcount(a[w] & b[w])
inline int and_popcount(const uint64_t *a, const uint64_t *b, int n) {
#if defined(__x86_64__)
  if (n >= 4 && have_avx512_popcnt()) return and_popcount_avx512(a, b, n);
#endif
  int c = 0;
  for (int w = 0; w < n; ++w) c += __builtin_popcountll(a[w] & b[w]);
  return c;
}

// dst[w] = a[w] & b[w]; returns the number of set bits of dst
inline int and_store_popcount(uint64_t *dst, const uint64_t *a, const uint64_t *b, int n) {
#if defined(__x86_64__)
  if (n >= 4 && have_avx512_popcnt()) return and_store_popcount_avx512(dst, a, b, n);
#endif
  int c = 0;
  for (int w = 0; w < n; ++w) {
    dst[w] = a[w] & b[w];
    c += __builtin_popcountll(dst[w]);
  }
  return c;
}

// dst[w] = a[w] & ~b[w] (dst may alias a); returns the number of set bits of the result
inline int andnot_store_popcount(uint64_t *dst, const uint64_t *a, const uint64_t *b, int n) {
#if defined(__x86_64__)
  if (n >= 4 && have_avx512_popcnt()) return andnot_store_popcount_avx512(dst, a, b, n);
#endif
  int c = 0;
  for (int w = 0; w < n; ++w) {
    dst[w] = a[w] & ~b[w];
    c += __builtin_popcountll(dst[w]);
  }
  return c;
}

// dst[w] = a[w] & ~b[w] (dst may alias a); returns the OR of all result words
inline uint64_t andnot_store_any(uint64_t *dst, const uint64_t *a, const uint64_t *b, int n) {
#if defined(__x86_64__)
  if (n >= 4 && have_avx512_popcnt()) return andnot_store_any_avx512(dst, a, b, n);
#endif
  uint64_t any = 0;
  for (int w = 0; w < n; ++w) any |= (dst[w] = a[w] & ~b[w]);
  return any;
}

}  // namespace bitops
}  // namespace tod
#endif
