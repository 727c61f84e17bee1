// innr_ref.cpp -- CPU oracle (C++ restatement of innr 0.6.3) for the batch
// similarity-search hot path. TEST INFRASTRUCTURE, NOT PRODUCT: see innr_ref.h.
//
// Build: g++ -O3 -std=c++17 -ffp-contract=off -fPIC -shared (oracle/Makefile).
// -ffp-contract=off is mandatory: the Rust reference never contracts a*b+c, and
// the per-vector scores of the batch_* loops are defined by a strictly
// sequential, separately rounded multiply and add (src/batch.rs:257-265).
//
// Parity pinning: ports of the reference's own unit tests / examples / KATs live
// in tests/test_ref_*.py and tests/golden/reference_kats.json (tests/test_golden.py). Items recalled from Rust std and not verifiable
// offline are marked [RECALLED].

#include "innr_ref.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include <immintrin.h>

namespace {

int g_simd_mode = 0;  // 0 auto, 1 force emulation

inline bool use_avx512_intrin() {
  return g_simd_mode == 0 && __builtin_cpu_supports("avx512f");
}
inline bool use_avx2_intrin() {
  return g_simd_mode == 0 && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
}

// f32::total_cmp (Rust core::f32): bits ^= ((bits >> 31) as u32 >> 1); signed compare.
inline int32_t total_order_key(float x) {
  int32_t b;
  std::memcpy(&b, &x, 4);
  b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
  return b;
}
inline int total_cmp(float a, float b) {  // -1 Less, 0 Equal, 1 Greater
  int32_t ka = total_order_key(a), kb = total_order_key(b);
  return ka < kb ? -1 : (ka > kb ? 1 : 0);
}

// GCC's _mm512_reduce_add_ps tree (avx512fintrin.h __MM512_REDUCE_OP): hi256+lo256,
// hi128+lo128, +shuffle{2,3,0,1}, [0]+[1]. rustc/LLVM lowers to the same halving
// tree [RECALLED]; addition is commutative so operand order is immaterial.
inline float reduce16(const float* v) {
  float t[8], u[4];
  for (int j = 0; j < 8; ++j) t[j] = v[j + 8] + v[j];
  for (int j = 0; j < 4; ++j) u[j] = t[j + 4] + t[j];
  float w0 = u[0] + u[2], w1 = u[1] + u[3];
  return w0 + w1;
}
// hsum of a __m256 as written at src/arch/x86_64.rs:243-248: lo128+hi128, +movehl, +shuffle(1).
inline float hsum8(const float* v) {
  float s[4];
  for (int j = 0; j < 4; ++j) s[j] = v[j] + v[j + 4];
  float a0 = s[0] + s[2], a1 = s[1] + s[3];
  return a0 + a1;
}

// ---------------------------------------------------------------------------
// Scalar "virtual lane chain" emulations of the explicit SIMD kernels.
// ---------------------------------------------------------------------------

// src/arch/x86_64.rs:31-106
float dot_avx512_emul(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  float acc[64];
  for (int c = 0; c < 64; ++c) acc[c] = 0.0f;
  size_t chunks64 = n / 64;
  for (size_t i = 0; i < chunks64; ++i)
    for (int c = 0; c < 64; ++c) acc[c] = fmaf(a[i * 64 + c], b[i * 64 + c], acc[c]);
  float all[16];
  for (int j = 0; j < 16; ++j) {
    float s01 = acc[j] + acc[16 + j];
    float s23 = acc[32 + j] + acc[48 + j];
    all[j] = s01 + s23;
  }
  float result = reduce16(all);
  size_t rs = chunks64 * 64, remaining = n - rs;
  if (remaining > 0) {
    float rem[16];
    for (int j = 0; j < 16; ++j) rem[j] = 0.0f;
    size_t chunks16 = remaining / 16;
    for (size_t i = 0; i < chunks16; ++i)
      for (int j = 0; j < 16; ++j) rem[j] = fmaf(a[rs + i * 16 + j], b[rs + i * 16 + j], rem[j]);
    size_t tail = remaining % 16;
    if (tail > 0) {
      size_t off = rs + chunks16 * 16;
      for (int j = 0; j < 16; ++j) {
        float va = (size_t)j < tail ? a[off + j] : 0.0f;  // maskz load
        float vb = (size_t)j < tail ? b[off + j] : 0.0f;
        rem[j] = fmaf(va, vb, rem[j]);
      }
    }
    result += reduce16(rem);
  }
  return result;
}

// src/arch/x86_64.rs:183-265
float dot_avx2_emul(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  float acc[32];
  for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
  size_t chunks32 = n / 32;
  for (size_t i = 0; i < chunks32; ++i)
    for (int c = 0; c < 32; ++c) acc[c] = fmaf(a[i * 32 + c], b[i * 32 + c], acc[c]);
  float all[8];
  for (int j = 0; j < 8; ++j) all[j] = (acc[j] + acc[8 + j]) + (acc[16 + j] + acc[24 + j]);
  float result = hsum8(all);
  size_t rs = chunks32 * 32, remaining = n - rs, chunks8 = remaining / 8;
  float rem[8];
  for (int j = 0; j < 8; ++j) rem[j] = 0.0f;
  for (size_t i = 0; i < chunks8; ++i)
    for (int j = 0; j < 8; ++j) rem[j] = fmaf(a[rs + i * 8 + j], b[rs + i * 8 + j], rem[j]);
  result += hsum8(rem);
  for (size_t i = rs + chunks8 * 8; i < n; ++i) result += a[i] * b[i];
  return result;
}

inline float cosine_finish(float ab, float aa, float bb) {
  const float eps_sq = INNR_REF_NORM_EPSILON * INNR_REF_NORM_EPSILON;  // src/lib.rs:184
  if (aa > eps_sq && bb > eps_sq) return ab / (std::sqrt(aa) * std::sqrt(bb));
  return 0.0f;
}

// src/arch/x86_64.rs:681-786
float cosine_avx512_emul(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  float ab[64], aa[64], bb[64];
  for (int c = 0; c < 64; ++c) ab[c] = aa[c] = bb[c] = 0.0f;
  size_t chunks64 = n / 64;
  for (size_t i = 0; i < chunks64; ++i)
    for (int c = 0; c < 64; ++c) {
      float va = a[i * 64 + c], vb = b[i * 64 + c];
      ab[c] = fmaf(va, vb, ab[c]);
      aa[c] = fmaf(va, va, aa[c]);
      bb[c] = fmaf(vb, vb, bb[c]);
    }
  auto combine = [](const float* acc) {
    float all[16];
    for (int j = 0; j < 16; ++j) all[j] = (acc[j] + acc[16 + j]) + (acc[32 + j] + acc[48 + j]);
    return reduce16(all);
  };
  float rab = combine(ab), raa = combine(aa), rbb = combine(bb);
  size_t rs = chunks64 * 64, remaining = n - rs;
  if (remaining > 0) {
    float eab[16], eaa[16], ebb[16];
    for (int j = 0; j < 16; ++j) eab[j] = eaa[j] = ebb[j] = 0.0f;
    size_t chunks16 = remaining / 16;
    for (size_t i = 0; i < chunks16; ++i)
      for (int j = 0; j < 16; ++j) {
        float va = a[rs + i * 16 + j], vb = b[rs + i * 16 + j];
        eab[j] = fmaf(va, vb, eab[j]);
        eaa[j] = fmaf(va, va, eaa[j]);
        ebb[j] = fmaf(vb, vb, ebb[j]);
      }
    size_t tail = remaining % 16;
    if (tail > 0) {
      size_t off = rs + chunks16 * 16;
      for (int j = 0; j < 16; ++j) {
        float va = (size_t)j < tail ? a[off + j] : 0.0f;
        float vb = (size_t)j < tail ? b[off + j] : 0.0f;
        eab[j] = fmaf(va, vb, eab[j]);
        eaa[j] = fmaf(va, va, eaa[j]);
        ebb[j] = fmaf(vb, vb, ebb[j]);
      }
    }
    rab += reduce16(eab);
    raa += reduce16(eaa);
    rbb += reduce16(ebb);
  }
  return cosine_finish(rab, raa, rbb);
}

// src/arch/x86_64.rs:799-915
float cosine_avx2_emul(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  float ab[32], aa[32], bb[32];
  for (int c = 0; c < 32; ++c) ab[c] = aa[c] = bb[c] = 0.0f;
  size_t chunks32 = n / 32;
  for (size_t i = 0; i < chunks32; ++i)
    for (int c = 0; c < 32; ++c) {
      float va = a[i * 32 + c], vb = b[i * 32 + c];
      ab[c] = fmaf(va, vb, ab[c]);
      aa[c] = fmaf(va, va, aa[c]);
      bb[c] = fmaf(vb, vb, bb[c]);
    }
  auto combine = [](const float* acc) {
    float all[8];
    for (int j = 0; j < 8; ++j) all[j] = (acc[j] + acc[8 + j]) + (acc[16 + j] + acc[24 + j]);
    return hsum8(all);
  };
  float rab = combine(ab), raa = combine(aa), rbb = combine(bb);
  size_t rs = chunks32 * 32, remaining = n - rs, chunks8 = remaining / 8;
  float eab[8], eaa[8], ebb[8];
  for (int j = 0; j < 8; ++j) eab[j] = eaa[j] = ebb[j] = 0.0f;
  for (size_t i = 0; i < chunks8; ++i)
    for (int j = 0; j < 8; ++j) {
      float va = a[rs + i * 8 + j], vb = b[rs + i * 8 + j];
      eab[j] = fmaf(va, vb, eab[j]);
      eaa[j] = fmaf(va, va, eaa[j]);
      ebb[j] = fmaf(vb, vb, ebb[j]);
    }
  rab += hsum8(eab);
  raa += hsum8(eaa);
  rbb += hsum8(ebb);
  for (size_t i = rs + chunks8 * 8; i < n; ++i) {
    float ai = a[i], bi = b[i];
    rab += ai * bi;
    raa += ai * ai;
    rbb += bi * bi;
  }
  return cosine_finish(rab, raa, rbb);
}

// src/arch/x86_64.rs:928-1020
float dot_u8_f32_avx2_emul(const float* a, const uint8_t* b, size_t n) {
  if (n == 0) return 0.0f;
  float acc[32];
  for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
  size_t chunks32 = n / 32;
  for (size_t i = 0; i < chunks32; ++i)
    for (int c = 0; c < 32; ++c) acc[c] = fmaf(a[i * 32 + c], (float)b[i * 32 + c], acc[c]);
  float all[8];
  for (int j = 0; j < 8; ++j) all[j] = (acc[j] + acc[8 + j]) + (acc[16 + j] + acc[24 + j]);
  float result = hsum8(all);
  size_t rs = chunks32 * 32, remaining = n - rs, chunks8 = remaining / 8;
  float rem[8];
  for (int j = 0; j < 8; ++j) rem[j] = 0.0f;
  for (size_t i = 0; i < chunks8; ++i)
    for (int j = 0; j < 8; ++j) rem[j] = fmaf(a[rs + i * 8 + j], (float)b[rs + i * 8 + j], rem[j]);
  result += hsum8(rem);
  for (size_t i = rs + chunks8 * 8; i < n; ++i) result += a[i] * (float)b[i];
  return result;
}

// ---------------------------------------------------------------------------
// The same kernels with the reference's intrinsics in the reference's order.
// ---------------------------------------------------------------------------

__attribute__((target("avx512f"))) float dot_avx512_intrin(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  size_t chunks64 = n / 64;
  __m512 s0 = _mm512_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
  for (size_t i = 0; i < chunks64; ++i) {
    size_t base = i * 64;
    __m512 va0 = _mm512_loadu_ps(a + base), vb0 = _mm512_loadu_ps(b + base);
    __m512 va1 = _mm512_loadu_ps(a + base + 16), vb1 = _mm512_loadu_ps(b + base + 16);
    __m512 va2 = _mm512_loadu_ps(a + base + 32), vb2 = _mm512_loadu_ps(b + base + 32);
    __m512 va3 = _mm512_loadu_ps(a + base + 48), vb3 = _mm512_loadu_ps(b + base + 48);
    s0 = _mm512_fmadd_ps(va0, vb0, s0);
    s1 = _mm512_fmadd_ps(va1, vb1, s1);
    s2 = _mm512_fmadd_ps(va2, vb2, s2);
    s3 = _mm512_fmadd_ps(va3, vb3, s3);
  }
  __m512 s01 = _mm512_add_ps(s0, s1), s23 = _mm512_add_ps(s2, s3);
  float result = _mm512_reduce_add_ps(_mm512_add_ps(s01, s23));
  size_t rs = chunks64 * 64, remaining = n - rs;
  if (remaining > 0) {
    size_t chunks16 = remaining / 16;
    __m512 sr = _mm512_setzero_ps();
    for (size_t i = 0; i < chunks16; ++i) {
      size_t off = rs + i * 16;
      sr = _mm512_fmadd_ps(_mm512_loadu_ps(a + off), _mm512_loadu_ps(b + off), sr);
    }
    size_t tail = remaining % 16;
    if (tail > 0) {
      size_t off = rs + chunks16 * 16;
      __mmask16 m = (__mmask16)((1u << tail) - 1);
      sr = _mm512_fmadd_ps(_mm512_maskz_loadu_ps(m, a + off), _mm512_maskz_loadu_ps(m, b + off), sr);
    }
    result += _mm512_reduce_add_ps(sr);
  }
  return result;
}

__attribute__((target("avx2,fma"))) inline float hsum256(__m256 v) {
  __m128 hi = _mm256_extractf128_ps(v, 1);
  __m128 lo = _mm256_castps256_ps128(v);
  __m128 s128 = _mm_add_ps(lo, hi);
  __m128 s64 = _mm_add_ps(s128, _mm_movehl_ps(s128, s128));
  __m128 s32 = _mm_add_ss(s64, _mm_shuffle_ps(s64, s64, 1));
  return _mm_cvtss_f32(s32);
}

__attribute__((target("avx2,fma"))) float dot_avx2_intrin(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  size_t chunks32 = n / 32;
  __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
  for (size_t i = 0; i < chunks32; ++i) {
    size_t base = i * 32;
    s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base), _mm256_loadu_ps(b + base), s0);
    s1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base + 8), _mm256_loadu_ps(b + base + 8), s1);
    s2 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base + 16), _mm256_loadu_ps(b + base + 16), s2);
    s3 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base + 24), _mm256_loadu_ps(b + base + 24), s3);
  }
  float result = hsum256(_mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3)));
  size_t rs = chunks32 * 32, remaining = n - rs, chunks8 = remaining / 8;
  __m256 sum = _mm256_setzero_ps();
  for (size_t i = 0; i < chunks8; ++i) {
    size_t off = rs + i * 8;
    sum = _mm256_fmadd_ps(_mm256_loadu_ps(a + off), _mm256_loadu_ps(b + off), sum);
  }
  result += hsum256(sum);
  for (size_t i = rs + chunks8 * 8; i < n; ++i) result += a[i] * b[i];
  return result;
}

__attribute__((target("avx512f"))) float cosine_avx512_intrin(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  size_t chunks64 = n / 64;
  __m512 z = _mm512_setzero_ps();
  __m512 ab0 = z, ab1 = z, ab2 = z, ab3 = z, aa0 = z, aa1 = z, aa2 = z, aa3 = z, bb0 = z, bb1 = z,
         bb2 = z, bb3 = z;
  for (size_t i = 0; i < chunks64; ++i) {
    size_t base = i * 64;
    __m512 va0 = _mm512_loadu_ps(a + base), vb0 = _mm512_loadu_ps(b + base);
    __m512 va1 = _mm512_loadu_ps(a + base + 16), vb1 = _mm512_loadu_ps(b + base + 16);
    __m512 va2 = _mm512_loadu_ps(a + base + 32), vb2 = _mm512_loadu_ps(b + base + 32);
    __m512 va3 = _mm512_loadu_ps(a + base + 48), vb3 = _mm512_loadu_ps(b + base + 48);
    ab0 = _mm512_fmadd_ps(va0, vb0, ab0);
    ab1 = _mm512_fmadd_ps(va1, vb1, ab1);
    ab2 = _mm512_fmadd_ps(va2, vb2, ab2);
    ab3 = _mm512_fmadd_ps(va3, vb3, ab3);
    aa0 = _mm512_fmadd_ps(va0, va0, aa0);
    aa1 = _mm512_fmadd_ps(va1, va1, aa1);
    aa2 = _mm512_fmadd_ps(va2, va2, aa2);
    aa3 = _mm512_fmadd_ps(va3, va3, aa3);
    bb0 = _mm512_fmadd_ps(vb0, vb0, bb0);
    bb1 = _mm512_fmadd_ps(vb1, vb1, bb1);
    bb2 = _mm512_fmadd_ps(vb2, vb2, bb2);
    bb3 = _mm512_fmadd_ps(vb3, vb3, bb3);
  }
  float ab = _mm512_reduce_add_ps(_mm512_add_ps(_mm512_add_ps(ab0, ab1), _mm512_add_ps(ab2, ab3)));
  float aa = _mm512_reduce_add_ps(_mm512_add_ps(_mm512_add_ps(aa0, aa1), _mm512_add_ps(aa2, aa3)));
  float bb = _mm512_reduce_add_ps(_mm512_add_ps(_mm512_add_ps(bb0, bb1), _mm512_add_ps(bb2, bb3)));
  size_t rs = chunks64 * 64, remaining = n - rs;
  if (remaining > 0) {
    size_t chunks16 = remaining / 16;
    __m512 rab = z, raa = z, rbb = z;
    for (size_t i = 0; i < chunks16; ++i) {
      size_t off = rs + i * 16;
      __m512 va = _mm512_loadu_ps(a + off), vb = _mm512_loadu_ps(b + off);
      rab = _mm512_fmadd_ps(va, vb, rab);
      raa = _mm512_fmadd_ps(va, va, raa);
      rbb = _mm512_fmadd_ps(vb, vb, rbb);
    }
    size_t tail = remaining % 16;
    if (tail > 0) {
      size_t off = rs + chunks16 * 16;
      __mmask16 m = (__mmask16)((1u << tail) - 1);
      __m512 va = _mm512_maskz_loadu_ps(m, a + off), vb = _mm512_maskz_loadu_ps(m, b + off);
      rab = _mm512_fmadd_ps(va, vb, rab);
      raa = _mm512_fmadd_ps(va, va, raa);
      rbb = _mm512_fmadd_ps(vb, vb, rbb);
    }
    ab += _mm512_reduce_add_ps(rab);
    aa += _mm512_reduce_add_ps(raa);
    bb += _mm512_reduce_add_ps(rbb);
  }
  return cosine_finish(ab, aa, bb);
}

__attribute__((target("avx2,fma"))) float cosine_avx2_intrin(const float* a, const float* b, size_t n) {
  if (n == 0) return 0.0f;
  size_t chunks32 = n / 32;
  __m256 z = _mm256_setzero_ps();
  __m256 ab0 = z, ab1 = z, ab2 = z, ab3 = z, aa0 = z, aa1 = z, aa2 = z, aa3 = z, bb0 = z, bb1 = z,
         bb2 = z, bb3 = z;
  for (size_t i = 0; i < chunks32; ++i) {
    size_t base = i * 32;
    __m256 va0 = _mm256_loadu_ps(a + base), vb0 = _mm256_loadu_ps(b + base);
    __m256 va1 = _mm256_loadu_ps(a + base + 8), vb1 = _mm256_loadu_ps(b + base + 8);
    __m256 va2 = _mm256_loadu_ps(a + base + 16), vb2 = _mm256_loadu_ps(b + base + 16);
    __m256 va3 = _mm256_loadu_ps(a + base + 24), vb3 = _mm256_loadu_ps(b + base + 24);
    ab0 = _mm256_fmadd_ps(va0, vb0, ab0);
    ab1 = _mm256_fmadd_ps(va1, vb1, ab1);
    ab2 = _mm256_fmadd_ps(va2, vb2, ab2);
    ab3 = _mm256_fmadd_ps(va3, vb3, ab3);
    aa0 = _mm256_fmadd_ps(va0, va0, aa0);
    aa1 = _mm256_fmadd_ps(va1, va1, aa1);
    aa2 = _mm256_fmadd_ps(va2, va2, aa2);
    aa3 = _mm256_fmadd_ps(va3, va3, aa3);
    bb0 = _mm256_fmadd_ps(vb0, vb0, bb0);
    bb1 = _mm256_fmadd_ps(vb1, vb1, bb1);
    bb2 = _mm256_fmadd_ps(vb2, vb2, bb2);
    bb3 = _mm256_fmadd_ps(vb3, vb3, bb3);
  }
  float ab = hsum256(_mm256_add_ps(_mm256_add_ps(ab0, ab1), _mm256_add_ps(ab2, ab3)));
  float aa = hsum256(_mm256_add_ps(_mm256_add_ps(aa0, aa1), _mm256_add_ps(aa2, aa3)));
  float bb = hsum256(_mm256_add_ps(_mm256_add_ps(bb0, bb1), _mm256_add_ps(bb2, bb3)));
  size_t rs = chunks32 * 32, remaining = n - rs, chunks8 = remaining / 8;
  __m256 rab = z, raa = z, rbb = z;
  for (size_t i = 0; i < chunks8; ++i) {
    size_t off = rs + i * 8;
    __m256 va = _mm256_loadu_ps(a + off), vb = _mm256_loadu_ps(b + off);
    rab = _mm256_fmadd_ps(va, vb, rab);
    raa = _mm256_fmadd_ps(va, va, raa);
    rbb = _mm256_fmadd_ps(vb, vb, rbb);
  }
  ab += hsum256(rab);
  aa += hsum256(raa);
  bb += hsum256(rbb);
  for (size_t i = rs + chunks8 * 8; i < n; ++i) {
    float ai = a[i], bi = b[i];
    ab += ai * bi;
    aa += ai * ai;
    bb += bi * bi;
  }
  return cosine_finish(ab, aa, bb);
}

__attribute__((target("avx2,fma"))) inline __m256 load8_u8_as_f32(const uint8_t* p) {
  return _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)p)));
}

__attribute__((target("avx2,fma"))) float dot_u8_f32_avx2_intrin(const float* a, const uint8_t* b,
                                                                  size_t n) {
  if (n == 0) return 0.0f;
  size_t chunks32 = n / 32;
  __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
  for (size_t i = 0; i < chunks32; ++i) {
    size_t base = i * 32;
    __m256 b0 = load8_u8_as_f32(b + base), b1 = load8_u8_as_f32(b + base + 8);
    __m256 b2 = load8_u8_as_f32(b + base + 16), b3 = load8_u8_as_f32(b + base + 24);
    s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base), b0, s0);
    s1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base + 8), b1, s1);
    s2 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base + 16), b2, s2);
    s3 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base + 24), b3, s3);
  }
  float result = hsum256(_mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3)));
  size_t rs = chunks32 * 32, remaining = n - rs, chunks8 = remaining / 8;
  __m256 sum = _mm256_setzero_ps();
  for (size_t i = 0; i < chunks8; ++i) {
    size_t off = rs + i * 8;
    sum = _mm256_fmadd_ps(_mm256_loadu_ps(a + off), load8_u8_as_f32(b + off), sum);
  }
  result += hsum256(sum);
  for (size_t i = rs + chunks8 * 8; i < n; ++i) result += a[i] * (float)b[i];
  return result;
}

// Dispatch as the reference does on an AVX-512 host. When the host lacks the ISA
// the bit-identical emulation answers instead.
inline float k_dot_avx512(const float* a, const float* b, size_t n) {
  return use_avx512_intrin() ? dot_avx512_intrin(a, b, n) : dot_avx512_emul(a, b, n);
}
inline float k_dot_avx2(const float* a, const float* b, size_t n) {
  return use_avx2_intrin() ? dot_avx2_intrin(a, b, n) : dot_avx2_emul(a, b, n);
}
inline float k_cosine_avx512(const float* a, const float* b, size_t n) {
  return use_avx512_intrin() ? cosine_avx512_intrin(a, b, n) : cosine_avx512_emul(a, b, n);
}
inline float k_cosine_avx2(const float* a, const float* b, size_t n) {
  return use_avx2_intrin() ? cosine_avx2_intrin(a, b, n) : cosine_avx2_emul(a, b, n);
}
inline float k_dot_u8_f32_avx2(const float* a, const uint8_t* b, size_t n) {
  return use_avx2_intrin() ? dot_u8_f32_avx2_intrin(a, b, n) : dot_u8_f32_avx2_emul(a, b, n);
}

// ---------------------------------------------------------------------------
// Batch loops. No explicit SIMD in the reference (src/batch.rs:257-265,
// 290-296, 676-681): LLVM vectorises across vectors i; every per-vector result
// is the sequential unfused sum over d. target_clones lets GCC do the same
// across-i vectorisation on whatever ISA the host has; bits do not depend on it.
// ---------------------------------------------------------------------------
#define INNR_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))

INNR_CLONES void batch_l2_squared_impl(const float* q, const float* pdx, size_t n, size_t d,
                                       float* __restrict out) {
  for (size_t i = 0; i < n; ++i) out[i] = 0.0f;
  for (size_t dd = 0; dd < d; ++dd) {
    const float qd = q[dd];
    const float* __restrict row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i) {
      float diff = qd - row[i];
      out[i] += diff * diff;
    }
  }
}

INNR_CLONES void batch_dot_impl(const float* q, const float* pdx, size_t n, size_t d,
                                float* __restrict out) {
  for (size_t i = 0; i < n; ++i) out[i] = 0.0f;
  for (size_t dd = 0; dd < d; ++dd) {
    const float qd = q[dd];
    const float* __restrict row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i) out[i] += qd * row[i];
  }
}

INNR_CLONES void batch_norms_impl(const float* pdx, size_t n, size_t d, float* __restrict out) {
  for (size_t i = 0; i < n; ++i) out[i] = 0.0f;
  for (size_t dd = 0; dd < d; ++dd) {
    const float* __restrict row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i) out[i] += row[i] * row[i];
  }
  for (size_t i = 0; i < n; ++i) out[i] = std::sqrt(out[i]);
}

// src/batch.rs:705-728
void batch_cosine_impl(const float* q, const float* pdx, size_t n, size_t d, const float* norms,
                       float* out) {
  batch_dot_impl(q, pdx, n, d, out);
  float ss = 0.0f;  // query.iter().map(|x| x * x).sum::<f32>() -- sequential
  for (size_t dd = 0; dd < d; ++dd) ss += q[dd] * q[dd];
  float qn = std::sqrt(ss);
  if (qn < INNR_REF_NORM_EPSILON) {
    for (size_t i = 0; i < n; ++i) out[i] = 0.0f;
    return;
  }
  for (size_t i = 0; i < n; ++i) out[i] = norms[i] > INNR_REF_NORM_EPSILON ? out[i] / (qn * norms[i]) : 0.0f;
}

struct Pair {
  size_t idx;
  float score;
};

// indexed.sort_by(|a, b| b.1.total_cmp(&a.1)); truncate(k)  (stable; src/batch.rs:756-758)
size_t sort_desc_truncate(const float* scores, size_t n, size_t k, uint64_t* out_idx, float* out_score) {
  std::vector<Pair> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = {i, scores[i]};
  std::stable_sort(v.begin(), v.end(),
                   [](const Pair& a, const Pair& b) { return total_cmp(b.score, a.score) < 0; });
  for (size_t j = 0; j < k; ++j) {
    out_idx[j] = v[j].idx;
    out_score[j] = v[j].score;
  }
  return k;
}

}  // namespace

// ---------------------------------------------------------------------------
// TopK: src/topk.rs:47-187
// ---------------------------------------------------------------------------
struct innr_ref_topk {
  size_t k;
  std::vector<float> distances;  // sorted descending, [0] = worst
  std::vector<uint32_t> ids;
  size_t count;

  // slice::binary_search_by, Rust >= 1.82 [RECALLED]:
  //   while size > 1 { half = size/2; mid = base+half;
  //                    base = if f(mid) == Greater { base } else { mid }; size -= half }
  //   cmp = f(base); Equal -> Ok(base); else Err(base + (cmp == Less))
  // with f(d) = d.total_cmp(&distance).reverse()  (src/topk.rs:173-186)
  size_t find_insert_pos(float distance, size_t len) const {
    if (len == 0) return 0;
    auto f = [&](size_t i) { return -total_cmp(distances[i], distance); };
    size_t size = len, base = 0;
    while (size > 1) {
      size_t half = size / 2, mid = base + half;
      if (f(mid) != 1) base = mid;
      size -= half;
    }
    int c = f(base);
    if (c == 0) return base;
    return base + (c == -1 ? 1 : 0);
  }

  void insert(uint32_t id, float distance) {
    if (count < k) {  // :97-100
      size_t pos = find_insert_pos(distance, count);
      distances.insert(distances.begin() + pos, distance);
      ids.insert(ids.begin() + pos, id);
      ++count;
    } else if (total_cmp(distance, distances[0]) < 0) {  // :101
      distances.erase(distances.begin());                // copy_within(1.., 0)
      ids.erase(ids.begin());
      size_t pos = find_insert_pos(distance, k - 1);
      distances.insert(distances.begin() + pos, distance);
      ids.insert(ids.begin() + pos, id);
    }
  }
};

template <class F>
static void parallel_queries(size_t nq, int n_threads, F f) {
  if (n_threads <= 1) {
    for (size_t j = 0; j < nq; ++j) f(j);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t] {
      for (size_t j = t; j < nq; j += n_threads) f(j);
    });
  for (auto& x : th) x.join();
}

template <class F>
static void parallel_ranges(size_t total, int n_threads, F f) {
  if (n_threads <= 1 || total < 4096) {
    f((size_t)0, total);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (total + (size_t)n_threads - 1) / (size_t)n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const size_t lo = std::min(total, per * (size_t)t), hi = std::min(total, lo + per);
    if (lo < hi) th.emplace_back([=] { f(lo, hi); });
  }
  for (auto& x : th) x.join();
}
extern "C" {

int innr_ref_host_has_avx512(void) { return __builtin_cpu_supports("avx512f") ? 1 : 0; }
int innr_ref_host_has_avx2_fma(void) {
  return (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) ? 1 : 0;
}
void innr_ref_set_simd_mode(int mode) { g_simd_mode = mode; }

const char* innr_ref_dense_backend(size_t len) {  // src/backend.rs:46-67
  if (len >= INNR_REF_MIN_DIM_AVX512 && __builtin_cpu_supports("avx512f")) return "avx512";
  if (len >= INNR_REF_MIN_DIM_SIMD && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"))
    return "avx2+fma";
  return "portable";
}

float innr_ref_dot_portable(const float* a, const float* b, size_t n) {  // src/dense.rs:103-125
  size_t chunks = n / 4;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  for (size_t i = 0; i < chunks; ++i) {
    size_t base = i * 4;
    s0 += a[base] * b[base];
    s1 += a[base + 1] * b[base + 1];
    s2 += a[base + 2] * b[base + 2];
    s3 += a[base + 3] * b[base + 3];
  }
  float result = s0 + s1 + s2 + s3;
  for (size_t i = chunks * 4; i < n; ++i) result += a[i] * b[i];
  return result;
}

float innr_ref_cosine_portable(const float* a, const float* b, size_t n) {  // src/dense.rs:287-346
  size_t chunks = n / 4;
  float ab[4] = {0, 0, 0, 0}, aa[4] = {0, 0, 0, 0}, bb[4] = {0, 0, 0, 0};
  for (size_t i = 0; i < chunks; ++i)
    for (int j = 0; j < 4; ++j) {
      float x = a[i * 4 + j], y = b[i * 4 + j];
      ab[j] += x * y;
      aa[j] += x * x;
      bb[j] += y * y;
    }
  float rab = ab[0] + ab[1] + ab[2] + ab[3];
  float raa = aa[0] + aa[1] + aa[2] + aa[3];
  float rbb = bb[0] + bb[1] + bb[2] + bb[3];
  for (size_t i = chunks * 4; i < n; ++i) {
    rab += a[i] * b[i];
    raa += a[i] * a[i];
    rbb += b[i] * b[i];
  }
  return cosine_finish(rab, raa, rbb);
}

float innr_ref_dot(const float* a, const float* b, size_t n) {  // src/dense.rs:56-100
  if (n >= INNR_REF_MIN_DIM_AVX512) return k_dot_avx512(a, b, n);
  if (n >= INNR_REF_MIN_DIM_SIMD) return k_dot_avx2(a, b, n);
  return innr_ref_dot_portable(a, b, n);
}

float innr_ref_cosine(const float* a, const float* b, size_t n) {  // src/dense.rs:243-279
  if (n >= INNR_REF_MIN_DIM_AVX512) return k_cosine_avx512(a, b, n);
  if (n >= INNR_REF_MIN_DIM_SIMD) return k_cosine_avx2(a, b, n);
  return innr_ref_cosine_portable(a, b, n);
}

float innr_ref_dot_avx512_intrin(const float* a, const float* b, size_t n) { return dot_avx512_intrin(a, b, n); }
float innr_ref_dot_avx512_emul(const float* a, const float* b, size_t n) { return dot_avx512_emul(a, b, n); }
float innr_ref_dot_avx2_intrin(const float* a, const float* b, size_t n) { return dot_avx2_intrin(a, b, n); }
float innr_ref_dot_avx2_emul(const float* a, const float* b, size_t n) { return dot_avx2_emul(a, b, n); }
float innr_ref_cosine_avx512_intrin(const float* a, const float* b, size_t n) { return cosine_avx512_intrin(a, b, n); }
float innr_ref_cosine_avx512_emul(const float* a, const float* b, size_t n) { return cosine_avx512_emul(a, b, n); }
float innr_ref_cosine_avx2_intrin(const float* a, const float* b, size_t n) { return cosine_avx2_intrin(a, b, n); }
float innr_ref_cosine_avx2_emul(const float* a, const float* b, size_t n) { return cosine_avx2_emul(a, b, n); }

// ---- VerticalBatch ----------------------------------------------------------
void innr_ref_from_flat(const float* rows, size_t n, size_t d, float* pdx) {  // src/batch.rs:167-183
  for (size_t i = 0; i < n; ++i)
    for (size_t dd = 0; dd < d; ++dd) pdx[dd * n + i] = rows[i * d + dd];
}
void innr_ref_extract_vector(const float* pdx, size_t n, size_t d, size_t i, float* out) {
  for (size_t dd = 0; dd < d; ++dd) out[dd] = pdx[dd * n + i];
}

void innr_ref_batch_l2_squared(const float* q, const float* pdx, size_t n, size_t d, float* out) {
  batch_l2_squared_impl(q, pdx, n, d, out);
}
void innr_ref_batch_dot(const float* q, const float* pdx, size_t n, size_t d, float* out) {
  batch_dot_impl(q, pdx, n, d, out);
}
void innr_ref_batch_norms(const float* pdx, size_t n, size_t d, float* out) {
  batch_norms_impl(pdx, n, d, out);
}
void innr_ref_batch_cosine(const float* q, const float* pdx, size_t n, size_t d, const float* norms,
                           float* out) {
  batch_cosine_impl(q, pdx, n, d, norms, out);
}

// ---- kNN --------------------------------------------------------------------
size_t innr_ref_batch_knn(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                          uint64_t* out_idx, float* out_score) {  // src/batch.rs:385-411
  if (n == 0 || k == 0) return 0;
  k = std::min(k, n);
  std::vector<float> dist(n);
  batch_l2_squared_impl(q, pdx, n, d, dist.data());
  innr_ref_topk t{k, {}, {}, 0};
  t.distances.reserve(k + 1);
  t.ids.reserve(k + 1);
  for (size_t i = 0; i < n; ++i) t.insert((uint32_t)i, dist[i]);  // `i as u32` :403
  size_t m = t.count;
  for (size_t j = 0; j < m; ++j) {  // into_sorted reverses (src/topk.rs:140-145)
    out_idx[j] = t.ids[m - 1 - j];
    out_score[j] = t.distances[m - 1 - j];
  }
  return m;
}

size_t innr_ref_batch_knn_dot(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                              uint64_t* out_idx, float* out_score) {  // src/batch.rs:742-764
  if (n == 0 || k == 0) return 0;
  k = std::min(k, n);
  std::vector<float> s(n);
  batch_dot_impl(q, pdx, n, d, s.data());
  return sort_desc_truncate(s.data(), n, k, out_idx, out_score);
}

size_t innr_ref_batch_knn_cosine(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                                 uint64_t* out_idx, float* out_score) {  // src/batch.rs:777-800
  if (n == 0 || k == 0) return 0;
  k = std::min(k, n);
  std::vector<float> norms(n), s(n);
  batch_norms_impl(pdx, n, d, norms.data());  // recomputed on every call (:788)
  batch_cosine_impl(q, pdx, n, d, norms.data(), s.data());
  return sort_desc_truncate(s.data(), n, k, out_idx, out_score);
}

size_t innr_ref_batch_knn_filtered(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                                   const uint8_t* mask, uint64_t* out_idx, float* out_score) {
  if (n == 0 || k == 0) return 0;
  size_t passing = 0;
  for (size_t i = 0; i < n; ++i) passing += mask[i] ? 1 : 0;
  if (passing == 0) return 0;
  k = std::min(k, passing);
  std::vector<float> dist(n);
  for (size_t i = 0; i < n; ++i) dist[i] = mask[i] ? 0.0f : std::numeric_limits<float>::max();
  for (size_t dd = 0; dd < d; ++dd) {
    const float qd = q[dd];
    const float* row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i)
      if (mask[i]) {
        float diff = qd - row[i];
        dist[i] += diff * diff;
      }
  }
  std::vector<Pair> v;
  v.reserve(passing);
  for (size_t i = 0; i < n; ++i)
    if (mask[i]) v.push_back({i, dist[i]});
  std::stable_sort(v.begin(), v.end(),
                   [](const Pair& a, const Pair& b) { return total_cmp(a.score, b.score) < 0; });
  for (size_t j = 0; j < k; ++j) {
    out_idx[j] = v[j].idx;
    out_score[j] = v[j].score;
  }
  return k;
}

size_t innr_ref_batch_l2_squared_pruning(const float* q, const float* pdx, size_t n, size_t d,
                                         float threshold, uint64_t* out_idx, float* out_dist) {
  std::vector<float> dist(n, 0.0f);
  std::vector<char> alive(n, 1);
  size_t num_alive = n;
  for (size_t dd = 0; dd < d; ++dd) {
    if (num_alive == 0) break;
    const float qd = q[dd];
    const float* row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i) {
      if (!alive[i]) continue;
      float diff = qd - row[i];
      dist[i] += diff * diff;
      if (dist[i] > threshold) {
        alive[i] = 0;
        --num_alive;
      }
    }
  }
  size_t m = 0;
  for (size_t i = 0; i < n; ++i)
    if (alive[i]) {
      out_idx[m] = i;
      out_dist[m] = dist[i];
      ++m;
    }
  return m;
}

// batch_knn_adaptive (src/batch.rs:441-564): warm-up over the first `warmup` dimensions, a threshold extrapolated from
// the k-th partial distance, then dimension-by-dimension pruning (never below k candidates) with the threshold refreshed
// after every dimension d with d % 32 == 0; survivors sorted (stable) by their complete distances.
// warmup == 0 is the caller's assertion failure (`warmup_dims must be > 0`): returns SIZE_MAX.
size_t innr_ref_batch_knn_adaptive(const float* q, const float* pdx, size_t n, size_t d, size_t k, size_t warmup,
                                   uint64_t* out_idx, float* out_score) {
  if (warmup == 0) return (size_t)-1;
  if (n == 0 || k == 0) return 0;
  k = std::min(k, n);
  if (d == 0) {
    for (size_t j = 0; j < k; ++j) { out_idx[j] = j; out_score[j] = 0.0f; }
    return k;
  }
  warmup = std::min(warmup, d);
  std::vector<float> dist(n, 0.0f);
  std::vector<char> alive(n, 1);
  size_t alive_count = n;
  for (size_t dd = 0; dd < warmup; ++dd) {
    const float qd = q[dd];
    const float* row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i) {
      float diff = qd - row[i];
      dist[i] += diff * diff;
    }
  }
  const float ratio = (float)d / (float)warmup;
  float threshold;
  {
    std::vector<float> sorted(dist);
    std::stable_sort(sorted.begin(), sorted.end(), [](float a, float b) { return total_cmp(a, b) < 0; });
    threshold = sorted[k - 1] * ratio;  // k <= n always holds here
  }
  for (size_t i = 0; i < n; ++i) {
    const float estimated_full = dist[i] * ratio;
    if (alive_count > k && estimated_full > threshold * 1.5f) {
      alive[i] = 0;
      --alive_count;
    }
  }
  std::vector<float> buf(n);
  for (size_t dd = warmup; dd < d; ++dd) {
    const float qd = q[dd];
    const float* row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i) {
      if (!alive[i]) continue;
      float diff = qd - row[i];
      dist[i] += diff * diff;
      if (alive_count > k && dist[i] > threshold) {
        alive[i] = 0;
        --alive_count;
      }
    }
    if (dd % 32 == 0) {
      size_t count = 0;
      for (size_t i = 0; i < n; ++i)
        if (alive[i]) buf[count++] = dist[i];
      if (count >= k) {
        std::nth_element(buf.begin(), buf.begin() + (k - 1), buf.begin() + count,
                         [](float a, float b) { return total_cmp(a, b) < 0; });
        threshold = buf[k - 1];
      }
    }
  }
  std::vector<Pair> v;
  for (size_t i = 0; i < n; ++i)
    if (alive[i]) v.push_back({i, dist[i]});
  std::stable_sort(v.begin(), v.end(), [](const Pair& a, const Pair& b) { return total_cmp(a.score, b.score) < 0; });
  const size_t m = std::min(k, v.size());
  for (size_t j = 0; j < m; ++j) {
    out_idx[j] = v[j].idx;
    out_score[j] = v[j].score;
  }
  return m;
}

// batch_dimension_variance (src/batch.rs:572-592): per dimension row, mean = (sequential sum) / n, then the sequential
// sum of (x - mean) * (x - mean), / n; n <= 1 or d == 0 -> zeros.
void innr_ref_batch_dimension_variance(const float* pdx, size_t n, size_t d, float* out) {
  if (n <= 1 || d == 0) {
    for (size_t dd = 0; dd < d; ++dd) out[dd] = 0.0f;
    return;
  }
  const float nf = (float)n;
  for (size_t dd = 0; dd < d; ++dd) {
    const float* row = pdx + dd * n;
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += row[i];
    const float mean = sum / nf;
    float acc = 0.0f;
    for (size_t i = 0; i < n; ++i) acc += (row[i] - mean) * (row[i] - mean);
    out[dd] = acc / nf;
  }
}

// variance_order (src/batch.rs:599-603): stable sort of 0..d by decreasing variance under total_cmp
void innr_ref_variance_order(const float* variances, size_t d, uint64_t* order) {
  std::vector<size_t> o(d);
  for (size_t i = 0; i < d; ++i) o[i] = i;
  std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return total_cmp(variances[b], variances[a]) < 0; });
  for (size_t i = 0; i < d; ++i) order[i] = o[i];
}

// batch_knn_reordered (src/batch.rs:621-659): distances accumulated over the dimensions in variance order, then a
// stable ascending sort (ties -> lower index, unlike batch_knn's TopK) and truncate.
size_t innr_ref_batch_knn_reordered(const float* q, const float* pdx, size_t n, size_t d, size_t k,
                                    uint64_t* out_idx, float* out_score) {
  if (n == 0 || k == 0) return 0;
  k = std::min(k, n);
  std::vector<float> var(d);
  std::vector<uint64_t> order(d);
  innr_ref_batch_dimension_variance(pdx, n, d, var.data());
  innr_ref_variance_order(var.data(), d, order.data());
  std::vector<float> dist(n, 0.0f);
  for (size_t j = 0; j < d; ++j) {
    const size_t dd = (size_t)order[j];
    const float qd = q[dd];
    const float* row = pdx + dd * n;
    for (size_t i = 0; i < n; ++i) {
      float diff = qd - row[i];
      dist[i] += diff * diff;
    }
  }
  std::vector<Pair> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = {i, dist[i]};
  std::stable_sort(v.begin(), v.end(), [](const Pair& a, const Pair& b) { return total_cmp(a.score, b.score) < 0; });
  for (size_t j = 0; j < k; ++j) {
    out_idx[j] = v[j].idx;
    out_score[j] = v[j].score;
  }
  return k;
}

// ---- TopK -------------------------------------------------------------------
innr_ref_topk* innr_ref_topk_new(size_t k) {
  if (k == 0) return nullptr;  // reference: assert!(k > 0, "innr::TopK: k must be >= 1")
  auto* t = new innr_ref_topk{k, {}, {}, 0};
  t->distances.reserve(k + 1);
  t->ids.reserve(k + 1);
  return t;
}
void innr_ref_topk_free(innr_ref_topk* t) { delete t; }
void innr_ref_topk_insert(innr_ref_topk* t, uint32_t id, float distance) { t->insert(id, distance); }
float innr_ref_topk_threshold(const innr_ref_topk* t) {  // src/topk.rs:79-86
  return t->count < t->k ? std::numeric_limits<float>::infinity() : t->distances[0];
}
size_t innr_ref_topk_len(const innr_ref_topk* t) { return t->count; }
size_t innr_ref_topk_into_sorted(const innr_ref_topk* t, uint32_t* out_id, float* out_dist) {
  size_t m = t->count;
  for (size_t j = 0; j < m; ++j) {
    out_id[j] = t->ids[m - 1 - j];
    out_dist[j] = t->distances[m - 1 - j];
  }
  return m;
}

// ---- MaxSim -----------------------------------------------------------------
float innr_ref_maxsim(const float* q, size_t nq, const float* d, size_t nd, size_t dim) {
  if (nq == 0 || nd == 0) return 0.0f;  // src/maxsim.rs:97-99
  if (dim >= 64 || dim >= 16) {
    // maxsim_avx512 (src/arch/x86_64.rs:119-143) / maxsim_avx2 (:156-171): `>` running max, `+=` total
    float total = 0.0f;
    for (size_t i = 0; i < nq; ++i) {
      float mx = -std::numeric_limits<float>::infinity();
      for (size_t j = 0; j < nd; ++j) {
        float s = dim >= 64 ? k_dot_avx512(q + i * dim, d + j * dim, dim)
                            : k_dot_avx2(q + i * dim, d + j * dim, dim);
        if (s > mx) mx = s;
      }
      total += mx;
    }
    return total;
  }
  // maxsim_portable (src/maxsim.rs:140-152): fold(NEG_INFINITY, f32::max), .sum()
  float total = 0.0f;
  for (size_t i = 0; i < nq; ++i) {
    float mx = -std::numeric_limits<float>::infinity();
    for (size_t j = 0; j < nd; ++j) mx = fmaxf(mx, innr_ref_dot(q + i * dim, d + j * dim, dim));
    total += mx;
  }
  return total;
}

float innr_ref_maxsim_cosine(const float* q, size_t nq, const float* d, size_t nd, size_t dim) {
  if (nq == 0 || nd == 0) return 0.0f;  // src/maxsim.rs:169-171
  float total = 0.0f;                   // :185-193
  for (size_t i = 0; i < nq; ++i) {
    float mx = -std::numeric_limits<float>::infinity();
    for (size_t j = 0; j < nd; ++j) mx = fmaxf(mx, innr_ref_cosine(q + i * dim, d + j * dim, dim));
    total += mx;
  }
  return total;
}

void innr_ref_maxsim_corpus(const float* q, size_t nq, const float* tokens, const uint64_t* doc_offsets,
                            size_t n_docs, size_t dim, int cosine, float* out_scores, int n_threads) {
  auto work = [&](size_t lo, size_t hi) {
    for (size_t j = lo; j < hi; ++j) {
      size_t t0 = doc_offsets[j], t1 = doc_offsets[j + 1];
      out_scores[j] = cosine ? innr_ref_maxsim_cosine(q, nq, tokens + t0 * dim, t1 - t0, dim)
                             : innr_ref_maxsim(q, nq, tokens + t0 * dim, t1 - t0, dim);
    }
  };
  if (n_threads <= 1) {
    work(0, n_docs);
    return;
  }
  std::vector<std::thread> th;
  size_t per = (n_docs + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    size_t lo = std::min(n_docs, t * per), hi = std::min(n_docs, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& x : th) x.join();
}

// ---- binary -----------------------------------------------------------------
void innr_ref_packed_binary_mask(uint64_t* words, size_t dim_bits) {  // src/binary.rs:59-66
  size_t rem = dim_bits % 64, nw = (dim_bits + 63) / 64;
  if (rem != 0 && nw > 0) words[nw - 1] &= (1ULL << rem) - 1;
}
void innr_ref_encode_binary(const float* v, size_t n, float threshold, uint64_t* out_words) {
  size_t nw = (n + 63) / 64;
  for (size_t w = 0; w < nw; ++w) out_words[w] = 0;
  for (size_t i = 0; i < n; ++i)
    if (v[i] > threshold) out_words[i / 64] |= 1ULL << (i % 64);
}
uint32_t innr_ref_binary_hamming(const uint64_t* a, const uint64_t* b, size_t words) {
  uint32_t s = 0;
  for (size_t w = 0; w < words; ++w) s += (uint32_t)__builtin_popcountll(a[w] ^ b[w]);
  return s;
}
uint32_t innr_ref_binary_dot(const uint64_t* a, const uint64_t* b, size_t words) {  // src/binary.rs:178-185
  uint32_t s = 0;
  for (size_t w = 0; w < words; ++w) s += (uint32_t)__builtin_popcountll(a[w] & b[w]);
  return s;
}
float innr_ref_binary_jaccard(const uint64_t* a, const uint64_t* b, size_t words) {  // src/binary.rs:198-213
  const uint32_t inter = innr_ref_binary_dot(a, b, words);
  uint32_t uni = 0;
  for (size_t w = 0; w < words; ++w) uni += (uint32_t)__builtin_popcountll(a[w] | b[w]);
  return uni == 0 ? 1.0f : (float)inter / (float)uni;  // intersection as f32 / union as f32
}
// ---- ternary (src/ternary.rs) ---------------------------------------------------------------
void innr_ref_packed_ternary_mask(uint64_t* words, size_t dimension) {  // PackedTernary::new :72-79
  const size_t rem = dimension % 32;
  const size_t n = (dimension + 31) / 32;
  if (rem != 0 && n) words[n - 1] &= (1ULL << (rem * 2)) - 1;
}
void innr_ref_encode_ternary(const float* v, size_t n, float threshold, uint64_t* out_words) {  // :163-173
  const size_t words = (n + 31) / 32;
  for (size_t w = 0; w < words; ++w) out_words[w] = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t bits = 0;
    if (v[i] > threshold) bits = 1;        // set(i, 1)  -> 0b01
    else if (v[i] < -threshold) bits = 2;  // set(i, -1) -> 0b10
    out_words[i / 32] |= bits << ((i % 32) * 2);
  }
}
static const uint64_t T_ODD = 0x5555555555555555ULL, T_EVEN = 0xAAAAAAAAAAAAAAAAULL;
int32_t innr_ref_ternary_dot(const uint64_t* a, const uint64_t* b, size_t words) {  // :191-281 (popcnt path == portable)
  int64_t same = 0, diff = 0;
  for (size_t w = 0; w < words; ++w) {
    const uint64_t wa = a[w], wb = b[w];
    const uint64_t pos_a = wa & ~((wa & T_EVEN) >> 1) & T_ODD, pos_b = wb & ~((wb & T_EVEN) >> 1) & T_ODD;
    const uint64_t neg_a = ~wa & ((wa & T_EVEN) >> 1) & T_ODD, neg_b = ~wb & ((wb & T_EVEN) >> 1) & T_ODD;
    same += __builtin_popcountll((pos_a & pos_b) | (neg_a & neg_b));
    diff += __builtin_popcountll((pos_a & neg_b) | (neg_a & pos_b));
  }
  return (int32_t)(same - diff);
}
uint32_t innr_ref_ternary_hamming(const uint64_t* a, const uint64_t* b, size_t words) {  // :301-324
  uint32_t d = 0;
  for (size_t w = 0; w < words; ++w) {
    const uint64_t wa = a[w], wb = b[w];
    const uint64_t nz_a = (wa & T_ODD) | ((wa & T_EVEN) >> 1), nz_b = (wb & T_ODD) | ((wb & T_EVEN) >> 1);
    const uint64_t x = wa ^ wb;
    const uint64_t df = (x & T_ODD) | ((x & T_EVEN) >> 1);
    d += (uint32_t)__builtin_popcountll(df & nz_a & nz_b);
  }
  return d;
}
float innr_ref_ternary_asymmetric_dot(const float* q, const uint64_t* t, size_t dimension) {  // :286-296
  float sum = 0.0f;
  for (size_t i = 0; i < dimension; ++i) {
    const uint64_t bits = (t[i / 32] >> ((i % 32) * 2)) & 3;
    const float tv = bits == 1 ? 1.0f : (bits == 2 ? -1.0f : 0.0f);  // ternary.get(i) as f32
    sum += q[i] * tv;                                                // unfused (-ffp-contract=off)
  }
  return sum;
}
size_t innr_ref_hamming_topk(const uint64_t* q, const uint64_t* codes, size_t n, size_t words, size_t k,
                             uint64_t* out_idx, uint32_t* out_dist) {
  if (n == 0 || k == 0) return 0;
  k = std::min(k, n);
  struct P {
    size_t i;
    uint32_t h;
  };
  std::vector<P> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = {i, innr_ref_binary_hamming(q, codes + i * words, words)};
  std::stable_sort(v.begin(), v.end(), [](const P& a, const P& b) { return a.h < b.h; });  // sort_by_key
  for (size_t j = 0; j < k; ++j) {
    out_idx[j] = v[j].i;
    out_dist[j] = v[j].h;
  }
  return k;
}

// ---- scalar u8 ----------------------------------------------------------------
void innr_ref_qparams_from_range(float mn, float mx, float* alpha, float* offset) {
  float a = mx - mn;
  *alpha = a > 0.0f ? a : 1.0f;
  *offset = mn;
}
void innr_ref_qparams_fit(const float* v, size_t n, float* alpha, float* offset) {
  if (n == 0) {
    *alpha = 1.0f;
    *offset = 0.0f;
    return;
  }
  float mn = std::numeric_limits<float>::max(), mx = std::numeric_limits<float>::lowest();
  for (size_t i = 0; i < n; ++i) {
    if (v[i] < mn) mn = v[i];
    if (v[i] > mx) mx = v[i];
  }
  innr_ref_qparams_from_range(mn, mx, alpha, offset);
}
// QuantizationParams::fit_quantile (src/scalar.rs:104-137). quantile outside (0, 1] is the caller's assertion: returns 1.
int innr_ref_qparams_fit_quantile(const float* v, size_t n, float quantile, float* alpha, float* offset) {
  if (!(quantile > 0.0f && quantile <= 1.0f)) return 1;
  if (n == 0) {
    *alpha = 1.0f;
    *offset = 0.0f;
    return 0;
  }
  if (quantile >= 1.0f) {
    innr_ref_qparams_fit(v, n, alpha, offset);
    return 0;
  }
  std::vector<float> sorted;
  for (size_t i = 0; i < n; ++i)
    if (std::isfinite(v[i])) sorted.push_back(v[i]);
  std::stable_sort(sorted.begin(), sorted.end(), [](float a, float b) { return total_cmp(a, b) < 0; });
  if (sorted.empty()) {
    *alpha = 1.0f;
    *offset = 0.0f;
    return 0;
  }
  const float tail = (1.0f - quantile) / 2.0f;
  const float len = (float)sorted.size();
  size_t lo_idx = (size_t)std::floor(tail * len);
  size_t hi_idx = (size_t)std::ceil((1.0f - tail) * len);
  hi_idx = std::min(hi_idx, sorted.size() - 1);
  innr_ref_qparams_from_range(sorted[lo_idx], sorted[hi_idx], alpha, offset);
  return 0;
}
// asymmetric_dot_u8_precomputed (src/scalar.rs:286-300): the caller supplies query_sum
float innr_ref_asymmetric_dot_u8_precomputed(const float* q, const uint8_t* codes, size_t n, float alpha, float offset,
                                             float query_sum) {
  const float mixed = innr_ref_mixed_dot_u8_f32(q, codes, n);
  return (alpha / 255.0f) * mixed + offset * query_sum;
}
void innr_ref_quantize_u8(const float* v, size_t n, float alpha, float offset, uint8_t* out) {
  float inv_alpha = 255.0f / alpha;
  for (size_t i = 0; i < n; ++i) {
    float normalized = (v[i] - offset) * inv_alpha;
    float r = roundf(normalized);  // f32::round: half away from zero
    // clamp(0.0, 255.0) keeps NaN; `as u8` saturates and maps NaN to 0
    if (r != r) {
      out[i] = 0;
    } else {
      if (r < 0.0f) r = 0.0f;
      if (r > 255.0f) r = 255.0f;
      out[i] = (uint8_t)r;
    }
  }
}
float innr_ref_query_sum(const float* q, size_t n) {
  float s = 0.0f;  // iter().sum() from 0.0 (Rust >= 1.83 starts at -0.0; differs only for all-(-0.0) input)
  for (size_t i = 0; i < n; ++i) s += q[i];
  return s;
}
float innr_ref_mixed_dot_u8_f32_portable(const float* a, const uint8_t* b, size_t n) {
  float s = 0.0f;
  for (size_t i = 0; i < n; ++i) s += a[i] * (float)b[i];
  return s;
}
float innr_ref_mixed_dot_u8_f32(const float* a, const uint8_t* b, size_t n) {  // src/scalar.rs:327-349
  if (n >= 16) return k_dot_u8_f32_avx2(a, b, n);  // AVX2 kernel even on AVX-512 hosts
  return innr_ref_mixed_dot_u8_f32_portable(a, b, n);
}
float innr_ref_dot_u8_f32_avx2_intrin(const float* a, const uint8_t* b, size_t n) {
  return dot_u8_f32_avx2_intrin(a, b, n);
}
float innr_ref_dot_u8_f32_avx2_emul(const float* a, const uint8_t* b, size_t n) {
  return dot_u8_f32_avx2_emul(a, b, n);
}
static inline float asym_precomputed(const float* q, const uint8_t* codes, size_t n, float alpha,
                                     float offset, float query_sum) {
  float mixed = innr_ref_mixed_dot_u8_f32(q, codes, n);
  return (alpha / 255.0f) * mixed + offset * query_sum;  // src/scalar.rs:299, unfused
}
float innr_ref_asymmetric_dot_u8(const float* q, const uint8_t* codes, size_t n, float alpha, float offset) {
  return asym_precomputed(q, codes, n, alpha, offset, innr_ref_query_sum(q, n));
}
size_t innr_ref_batch_knn_u8(const float* q, const uint8_t* corpus, size_t n, size_t d, float alpha,
                             float offset, size_t k, uint64_t* out_idx, float* out_score) {
  if (n == 0 || k == 0) return 0;
  float qs = innr_ref_query_sum(q, d);
  k = std::min(k, n);
  std::vector<float> s(n);
  for (size_t i = 0; i < n; ++i) s[i] = asym_precomputed(q, corpus + i * d, d, alpha, offset, qs);
  return sort_desc_truncate(s.data(), n, k, out_idx, out_score);
}

// ---- generators -----------------------------------------------------------------
void innr_ref_generate_embedding(size_t dim, uint64_t seed, float* out) {
  for (size_t i = 0; i < dim; ++i) {
    uint64_t x = seed * 6364136223846793005ULL + (uint64_t)i * 1442695040888963407ULL;
    out[i] = ((float)(x >> 33) / (float)(1ULL << 31)) * 2.0f - 1.0f;
  }
}
void innr_ref_generate_normalized(size_t dim, uint64_t seed, float* out) {
  innr_ref_generate_embedding(dim, seed, out);
  float ss = 0.0f;
  for (size_t i = 0; i < dim; ++i) ss += out[i] * out[i];
  float norm = std::sqrt(ss);
  if (norm > std::numeric_limits<float>::epsilon())
    for (size_t i = 0; i < dim; ++i) out[i] /= norm;
}
uint64_t innr_ref_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
void innr_ref_ghash_f32(uint64_t salt, uint64_t first_idx, size_t count, float* out) {
  for (size_t i = 0; i < count; ++i) {
    uint64_t u = innr_ref_splitmix64(salt + first_idx + i);
    out[i] = (float)(u >> 40) * (1.0f / 8388608.0f) - 1.0f;
  }
}
void innr_ref_ghash_u64(uint64_t salt, uint64_t first_idx, size_t count, uint64_t* out) {
  for (size_t i = 0; i < count; ++i) out[i] = innr_ref_splitmix64(salt + first_idx + i);
}

// ---- multi-threaded corpus generators for bench.py's CPU legs (same values as the single-threaded generators above;
//      they exist so a full BASELINE-size corpus can be built on the host in seconds, directly in its final layout) ----
void innr_ref_ghash_f32_mt(uint64_t salt, uint64_t first_idx, size_t count, float* out, int n_threads) {
  parallel_ranges(count, n_threads, [=](size_t lo, size_t hi) { innr_ref_ghash_f32(salt, first_idx + lo, hi - lo, out + lo); });
}
void innr_ref_ghash_u64_mt(uint64_t salt, uint64_t first_idx, size_t count, uint64_t* out, int n_threads) {
  parallel_ranges(count, n_threads, [=](size_t lo, size_t hi) { innr_ref_ghash_u64(salt, first_idx + lo, hi - lo, out + lo); });
}
// G-hash rows [first_row, first_row + n) of dimension d written as a VerticalBatch: pdx[dd * n + i] = value(row i, dim dd)
// (what from_flat, src/batch.rs:167-183, makes of the row-major generator output -- without the row-major copy)
void innr_ref_ghash_pdx_mt(uint64_t salt, uint64_t first_row, size_t n, size_t d, float* pdx, int n_threads) {
  parallel_ranges(n, n_threads, [=](size_t lo, size_t hi) {
    constexpr size_t B = 256;  // rows per block: the d x B tile of writes stays cache-resident
    for (size_t i0 = lo; i0 < hi; i0 += B) {
      const size_t i1 = std::min(hi, i0 + B);
      for (size_t i = i0; i < i1; ++i) {
        const uint64_t base = salt + (first_row + i) * (uint64_t)d;
        for (size_t dd = 0; dd < d; ++dd) {
          uint64_t u = innr_ref_splitmix64(base + dd);
          pdx[dd * n + i] = (float)(u >> 40) * (1.0f / 8388608.0f) - 1.0f;
        }
      }
    }
  });
}
// quantize_u8 (src/scalar.rs:212-225) of G-hash f32 rows, row-major n x d codes, without the f32 intermediate
void innr_ref_ghash_u8_mt(uint64_t salt, uint64_t first_row, size_t n, size_t d, float alpha, float offset, uint8_t* out,
                          int n_threads) {
  parallel_ranges(n, n_threads, [=](size_t lo, size_t hi) {
    std::vector<float> row(d);
    for (size_t i = lo; i < hi; ++i) {
      innr_ref_ghash_f32(salt, (first_row + i) * (uint64_t)d, d, row.data());
      innr_ref_quantize_u8(row.data(), d, alpha, offset, out + i * d);
    }
  });
}

// ---- multi-threaded drivers -------------------------------------------------------
size_t innr_ref_batch_knn_many(int metric, const float* queries, size_t nq, const float* pdx, size_t n,
                               size_t d, size_t k, uint64_t* out_idx, float* out_score, int n_threads) {
  size_t kk = (n == 0 || k == 0) ? 0 : std::min(k, n);
  parallel_queries(nq, n_threads, [&](size_t j) {
    const float* q = queries + j * d;
    if (metric == 0) innr_ref_batch_knn_dot(q, pdx, n, d, k, out_idx + j * k, out_score + j * k);
    else if (metric == 1) innr_ref_batch_knn_cosine(q, pdx, n, d, k, out_idx + j * k, out_score + j * k);
    else innr_ref_batch_knn(q, pdx, n, d, k, out_idx + j * k, out_score + j * k);
  });
  return kk;
}
size_t innr_ref_hamming_topk_many(const uint64_t* queries, size_t nq, const uint64_t* codes, size_t n,
                                  size_t words, size_t k, uint64_t* out_idx, uint32_t* out_dist,
                                  int n_threads) {
  size_t kk = (n == 0 || k == 0) ? 0 : std::min(k, n);
  parallel_queries(nq, n_threads, [&](size_t j) {
    innr_ref_hamming_topk(queries + j * words, codes, n, words, k, out_idx + j * k, out_dist + j * k);
  });
  return kk;
}
size_t innr_ref_batch_knn_u8_many(const float* queries, size_t nq, const uint8_t* corpus, size_t n,
                                  size_t d, float alpha, float offset, size_t k, uint64_t* out_idx,
                                  float* out_score, int n_threads) {
  size_t kk = (n == 0 || k == 0) ? 0 : std::min(k, n);
  parallel_queries(nq, n_threads, [&](size_t j) {
    innr_ref_batch_knn_u8(queries + j * d, corpus, n, d, alpha, offset, k, out_idx + j * k,
                          out_score + j * k);
  });
  return kk;
}

}  // extern "C"
