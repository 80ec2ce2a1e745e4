// Bit-exact restatement of the two libm functions the reference decoder's arithmetic hinges on,
// usable from both host and device code.
//
// The reference's log_sum_exp<float> (native_client/ctcdecode/decoder_utils.h:46-53) and class
// log-prob (ctc_beam_search_decoder.cpp:355) call glibc's logf/expf.  CUDA's logf/expf differ from
// glibc in the last ulp on a fraction of inputs, which would change beam pruning.  glibc >= 2.27
// implements both with the ARM "optimized routines" algorithm: a small table + a short polynomial
// evaluated in double, rounded once to float.  On x86-64 with FMA (every Xeon the reference runs on)
// the ifunc-selected variant is the same C code compiled with -mfma, i.e. with the a*b+c expressions
// contracted.  We restate that algorithm with explicit fma placement; tests/test_hd_math.py checks
// it against the host libm EXHAUSTIVELY (every float) so the placement is pinned, not guessed.
//
// Published algorithm: glibc sysdeps/ieee754/flt-32/{e_logf.c,e_expf.c,math_config.h},
// tables sysdeps/ieee754/flt-32/{e_logf_data.c,e_exp2f_data.c} (not vendored in /root/reference;
// the reference links the system libm).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define STT_HD __host__ __device__ __forceinline__
#else
#define STT_HD static inline
#endif

namespace sttmath {

#define STT_LOGF_TAB                                              \
  {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2},                  \
  {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},                  \
  {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},                   \
  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},                  \
  {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3},                  \
  {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},                     \
  {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4},                  \
  {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},                  \
  {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},                  \
  {0x1p+0, 0x0p+0},                                               \
  {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},                   \
  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},                    \
  {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},                   \
  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},                    \
  {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},                   \
  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}

#define STT_EXP2F_TAB                                                                     \
  0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
  0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
  0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
  0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
  0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
  0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
  0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
  0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL

struct LogfEntry {
  double invc, logc;
};

static const LogfEntry kLogfTabHost[16] = {STT_LOGF_TAB};
static const uint64_t kExp2fTabHost[32] = {STT_EXP2F_TAB};
#if defined(__CUDACC__)
static __device__ const LogfEntry kLogfTabDev[16] = {STT_LOGF_TAB};
static __device__ const uint64_t kExp2fTabDev[32] = {STT_EXP2F_TAB};
#endif

STT_HD uint32_t as_u32(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
STT_HD float as_f32(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
STT_HD uint64_t as_u64(double f) {
  uint64_t u;
  memcpy(&u, &f, 8);
  return u;
}
STT_HD double as_f64(uint64_t u) {
  double f;
  memcpy(&f, &u, 8);
  return f;
}
STT_HD double fma64(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
STT_HD double mul64(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);  // never contracted by nvcc
#else
  return a * b;
#endif
}
STT_HD double add64(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

// glibc __logf (e_logf.c), FMA-contracted variant.  Special cases return what glibc returns
// (errno side effects aside): log(+0)=-inf, log(x<0)=nan, log(inf)=inf.
STT_HD float glibc_logf_t(float x, const LogfEntry* tab) {
  const double Ln2 = 0x1.62e42fefa39efp-1;
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
  uint32_t ix = as_u32(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return as_f32(0xff800000u);          // -inf
    if (ix == 0x7f800000u) return x;                      // +inf
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return as_f32(0x7fc00000u);  // nan
    ix = as_u32(x * 0x1p23f);  // subnormal: normalise
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> 19) & 15;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & (0x1ffu << 23));
  const double invc = tab[i].invc, logc = tab[i].logc;
  double z = (double)as_f32(iz);
  double r = fma64(z, invc, -1.0);
  double y0 = fma64((double)k, Ln2, logc);
  double r2 = mul64(r, r);
  double y = fma64(A1, r, A2);
  y = fma64(A0, r2, y);
  y = fma64(y, r2, add64(y0, r));
  return (float)y;
}

// glibc __expf (e_expf.c), FMA-contracted variant, non-TOINT_INTRINSICS path (x86-64).
STT_HD float glibc_expf_t(float x, const uint64_t* tab) {
  const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
  const double Shift = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0,
               C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
  double xd = (double)x;
  uint32_t abstop = (as_u32(x) >> 20) & 0x7ff;
  if (abstop >= 0x42b) {  // |x| >= 88 or nan
    if (as_u32(x) == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8) return x + x;
    if (x > 0x1.62e42ep6f) return as_f32(0x7f800000u);  // overflow -> +inf
    if (x < -0x1.9fe368p6f) return 0.0f;                // underflow -> +0
  }
  double z = mul64(InvLn2N, xd);
  double kd = add64(z, Shift);
  uint64_t ki = as_u64(kd);
  kd = add64(kd, -Shift);
  // glibc's -mfma build contracts the product into this subtraction (pinned exhaustively:
  // with r = z - kd two of the 2^32 inputs differ from libm, with the fma none do).
  double r = fma64(InvLn2N, xd, -kd);
  uint64_t t = tab[ki & 31];
  t += ki << (52 - 5);
  double s = as_f64(t);
  z = fma64(C0, r, C1);
  double r2 = mul64(r, r);
  double y = fma64(C2, r, 1.0);
  y = fma64(z, r2, y);
  y = mul64(y, s);
  return (float)y;
}

// Default-table wrappers (global-memory tables on the device, static tables on the host).
STT_HD float glibc_logf(float x) {
#if defined(__CUDA_ARCH__)
  return glibc_logf_t(x, kLogfTabDev);
#else
  return glibc_logf_t(x, kLogfTabHost);
#endif
}
STT_HD float glibc_expf(float x) {
#if defined(__CUDA_ARCH__)
  return glibc_expf_t(x, kExp2fTabDev);
#else
  return glibc_expf_t(x, kExp2fTabHost);
#endif
}

// The two tables packed for a shared-memory copy (kernels whose serial path is dominated by these functions).
struct MathTables {
  LogfEntry logf_tab[16];
  uint64_t exp2f_tab[32];
};
#if defined(__CUDACC__)
__device__ __forceinline__ void load_math_tables(MathTables* dst, int tid, int nthreads) {
  for (int i = tid; i < 16; i += nthreads) dst->logf_tab[i] = kLogfTabDev[i];
  for (int i = tid; i < 32; i += nthreads) dst->exp2f_tab[i] = kExp2fTabDev[i];
}
#endif

// log_sum_exp<float>, decoder_utils.h:46-53 (num_min = -FLT_MAX).
STT_HD float log_sum_exp_t(float x, float y, const MathTables* mt) {
  const float num_min = -3.402823466e+38f;
  if (x <= num_min) return y;
  if (y <= num_min) return x;
  float xmax = x > y ? x : y;  // std::max(x, y): returns x when equal
  return glibc_logf_t(glibc_expf_t(x - xmax, mt->exp2f_tab) + glibc_expf_t(y - xmax, mt->exp2f_tab), mt->logf_tab) + xmax;
}
STT_HD float log_sum_exp(float x, float y) {
  const float num_min = -3.402823466e+38f;
  if (x <= num_min) return y;
  if (y <= num_min) return x;
  // std::max(x, y) returns x when equal.  One of the two exponents is exactly 0 and expf(0) == 1.0f exactly (the
  // exhaustive libm check covers it), float addition commutes, so only the other expf is evaluated.
  const float xmax = x > y ? x : y;
  const float d = x > y ? y - xmax : x - xmax;
  return glibc_logf(1.0f + glibc_expf(d)) + xmax;
}

}  // namespace sttmath
