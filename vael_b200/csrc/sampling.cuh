// Device helpers shared by sampling.cu (the stand-alone facts-head / Gumbel kernels) and logic_fused.cu (the
// fused training-mode logic kernel): the counter-based generator, the Gumbel transform and the facts-head row.
#pragma once
#include "common.cuh"

namespace vael {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (row, world), key = seed -------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// log2 / reciprocal of a NORMAL positive float: the bare MUFU instruction.  __log2f / __logf / 1.0f/x without
// -use_fast_math wrap it in a denormal-range check and rescale (4-5 instructions; profiles/r02i: 6.5 instructions per
// variate where 3 were expected); every argument below is a normal number by construction.
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLn2f = 0.6931471805599453f;

// Gumbel(0,1) = -log(E), E ~ Exp(1) = -log(U)  (torch: -empty.exponential_().log())
// U = (k + 1/2) 2^-23 with k the top 23 random bits: exact in fp32 and strictly inside (0,1), so E > 0 without a
// clamp.  The noise only has to be Gumbel-distributed, not bit-reproducible against a CPU libm, so the logarithms are
// the fast ones — except where the fast one is not good enough: for U -> 1, E = -log U -> 0 and __logf's ABSOLUTE
// error (2^-21.4 on [0.5, 2]) would be a large relative error of E, i.e. a wrong, possibly huge Gumbel value for
// about one sample in 1e3 .. 1e7.  There E comes from the series of -log(1 - v), v = 1 - U exact (Sterbenz).
__device__ __forceinline__ float gumbel_from_bits(uint32_t bits) {
  const float u = (static_cast<float>(bits >> 9) + 0.5f) * (1.0f / 8388608.0f);
  const float v = 1.0f - u;
  const float e_series = v * fmaf(v, fmaf(v, 0.33333334f, 0.5f), 1.0f);   // v + v^2/2 + v^3/3, |rel err| < v^3/4
  const float e = (v < 1.0f / 1024.0f) ? e_series : -kLn2f * fast_lg2(u);
  return -kLn2f * fast_lg2(e);
}

constexpr float kLog2e = 1.4426950408889634f;
// log2 of the exponential variate E2 = -log2(U) behind a Gumbel variate, U = (k + 1/2) 2^-23 (sampling.cuh):
// E2 in [2^-24 / ln 2, 24]  ->  log2 E2 in [-23.48, 4.59]
constexpr float kLogE2Min = -23.5f, kLogE2Max = 4.6f;

// The noise in the form the kernels consume: L = log2(E2), E2 = -log2(U) = E / ln 2 with E ~ Exp(1), so that the Gumbel
// variate is g = -ln E = -ln2 (L + log2 ln2) and a perturbed logit in the log2 domain is
//     log2(e)/tau (log p + g) = s log p - L / tau + const,    s = log2(e) / tau
// (the constant shifts every world of the row alike: arg-max and softmax do not see it).  Same accuracy guard as
// gumbel_from_bits: for U -> 1 the series of -log(1 - v) replaces the fast logarithm.
// v = 1 - U below which the series replaces the fast logarithm (lg2.approx: absolute error 2^-22 on [0.5, 1), i.e. a
// relative error of E = -log U of at most 0.3 % at the switch-over, one variate in 16 384), and the same bound on the
// 23-bit integer for the per-call test
constexpr float kSeriesBelow = 1.0f / 16384.0f;
constexpr uint32_t kSeriesFromK = 8388608u - 512u;
__device__ __forceinline__ float log2_exp_variate(uint32_t bits) {
  const float u = (static_cast<float>(bits >> 9) + 0.5f) * (1.0f / 8388608.0f);
  const float v = 1.0f - u;
  const float series = v * fmaf(v, fmaf(v, 0.48089835f, 0.72134752f), 1.44269504f);   // (v + v^2/2 + v^3/3) / ln 2
  const float e2 = (v < kSeriesBelow) ? series : -fast_lg2(u);
  return fast_lg2(e2);
}
// The four variates of one Philox call.  The guard is taken per CALL, on the integers: the series is needed when one
// of the four 23-bit numbers lies in the top 2^9 (v < 2^-14) — one call in 4096 — and that rare path is a real branch
// (the same values as four log2_exp_variate calls; 1.25 instead of 5 guard instructions per variate).
// (k + 1/2) 2^-23 for a 23-bit k without an int->float conversion (I2F shares the quarter-rate XU pipe with the
// MUFU logarithms this routine is made of): [1,2) float with mantissa k, minus (1 - 2^-24) — exact, the same value
__device__ __forceinline__ float unit_from_23_bits(uint32_t k) {
  return __uint_as_float(0x3f800000u | k) - 0.99999994f;
}
__device__ __forceinline__ void log2_exp_variates4(uint4 r, float* L4) {
  const uint32_t k0 = r.x >> 9, k1 = r.y >> 9, k2 = r.z >> 9, k3 = r.w >> 9;
  if (max(max(k0, k1), max(k2, k3)) >= kSeriesFromK) {
    L4[0] = log2_exp_variate(r.x); L4[1] = log2_exp_variate(r.y);
    L4[2] = log2_exp_variate(r.z); L4[3] = log2_exp_variate(r.w);
    return;
  }
  L4[0] = fast_lg2(-fast_lg2(unit_from_23_bits(k0)));
  L4[1] = fast_lg2(-fast_lg2(unit_from_23_bits(k1)));
  L4[2] = fast_lg2(-fast_lg2(unit_from_23_bits(k2)));
  L4[3] = fast_lg2(-fast_lg2(unit_from_23_bits(k3)));
}
// The exponential variates themselves, base 2: E2 = -log2(U) = E / ln 2, E ~ Exp(1), four per Philox call, with the
// same per-call guard (series of -log(1 - v) where the fast logarithm's absolute error would be a large relative one).
__device__ __forceinline__ float exp2_variate(uint32_t bits) {
  const float u = unit_from_23_bits(bits >> 9);
  const float v = 1.0f - u;
  const float series = v * fmaf(v, fmaf(v, 0.48089835f, 0.72134752f), 1.44269504f);   // (v + v^2/2 + v^3/3) / ln 2
  return (v < kSeriesBelow) ? series : -fast_lg2(u);
}
__device__ __forceinline__ void exp2_variates4(uint4 r, float* E4) {
  const uint32_t k0 = r.x >> 9, k1 = r.y >> 9, k2 = r.z >> 9, k3 = r.w >> 9;
  if (max(max(k0, k1), max(k2, k3)) >= kSeriesFromK) {
    E4[0] = exp2_variate(r.x); E4[1] = exp2_variate(r.y); E4[2] = exp2_variate(r.z); E4[3] = exp2_variate(r.w);
    return;
  }
  E4[0] = -fast_lg2(unit_from_23_bits(k0)); E4[1] = -fast_lg2(unit_from_23_bits(k1));
  E4[2] = -fast_lg2(unit_from_23_bits(k2)); E4[3] = -fast_lg2(unit_from_23_bits(k3));
}
// 2^x, x <= 0 (MUFU.EX2; 2 ulp): the softmax runs in the log2 domain, the log2(e) factor folded into 1/tau
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// row of K floats of a [32, K] shared-memory tile <-> registers, with the widest access the row pitch allows
// (pitches of 80 / 120 / 72 / 36 bytes: 128 / 64 / 64 / 32-bit accesses are conflict-free per quarter / half warp)
template <int K>
__device__ __forceinline__ void tile_row_load(const float* row, float* x) {
  constexpr int V = (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1);
#pragma unroll
  for (int c = 0; c < K; c += V) {
    if constexpr (V == 4) { const float4 t = *reinterpret_cast<const float4*>(row + c); x[c] = t.x; x[c + 1] = t.y; x[c + 2] = t.z; x[c + 3] = t.w; }
    else if constexpr (V == 2) { const float2 t = *reinterpret_cast<const float2*>(row + c); x[c] = t.x; x[c + 1] = t.y; }
    else x[c] = row[c];
  }
}
template <int K>
__device__ __forceinline__ void tile_row_store(float* row, const float* x) {
  constexpr int V = (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1);
#pragma unroll
  for (int c = 0; c < K; c += V) {
    if constexpr (V == 4) *reinterpret_cast<float4*>(row + c) = make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]);
    else if constexpr (V == 2) *reinterpret_cast<float2*>(row + c) = make_float2(x[c], x[c + 1]);
    else row[c] = x[c];
  }
}

// One row of the facts head: per group softmax y (exponentials as ex2.approx of the log2-scaled, non-positive
// argument: 2 instructions instead of expf's ~10, relative error ~1e-6 on the terms that matter — the bar against the
// reference's softmax is 1e-5), and rz[g] = 1 / sum_i (y_i + eps) (the detached normaliser of
// models/vael.py:159-177, as a reciprocal).  Two divisions per GROUP — 1/s and 1/Z — and multiplications per element:
// within 1.5 ulp of the reference's element-wise divisions, and ~290 fewer instructions per row of 2 x 10 logits
// (the IEEE division is ~8 instructions; the stand-alone kernel was issue-limited at 0.66 of HBM with 40 of them).
template <int NG, int N>
__device__ __forceinline__ void facts_row(const float* x, float eps, float* y, float* rz, float* zs = nullptr) {
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float mx = x[g * N];
#pragma unroll
    for (int i = 1; i < N; ++i) mx = fmaxf(mx, x[g * N + i]);
    float e[N], s = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) { e[i] = fast_exp2((x[g * N + i] - mx) * kLog2e); s += e[i]; }   // exp of a value <= 0
    const float rs = 1.0f / s;
    float zsum = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) { y[g * N + i] = e[i] * rs; zsum += y[g * N + i] + eps; }
    rz[g] = 1.0f / zsum;
    if (zs != nullptr) zs[g] = zsum;
  }
}

}  // namespace vael
