// The training-mode logic path as ONE kernel per direction (SURVEY.md §7 hard part 8):
//
//   logits [B,2N] --facts head--> facts [B,2N] --outer product--> worlds (registers) --masked sum--> query [B]
//          --log + Gumbel noise + arg-max / softmax--> world index, y_soft[arg] --Herbrand row--> world_h [B,2N]
//
// Reference, line by line: MNISTPairsVAELModel.compute_facts_probability (models/vael.py:159-177),
// problog_inference / compute_query (:200-225), `logits = torch.log(self.worlds_prob)`,
// `F.gumbel_softmax(logits, tau=1, hard=True)` (:61-62), `herbrand` with define_herbrand_base (:82-84, :179-187:
// row i*N+k = [onehot(i), onehot(k)], so `world @ H` is two one-hot halves and needs no table here).
// The unfused chain (facts_tile_kernel -> ad2_forward_kernel -> gumbel_forward_kernel) writes worlds[B,W] and
// y_soft[B,W] to HBM and reads them back: 1.5 KB per row forward, 2.3 KB backward.  Here nothing [B,W]-shaped
// exists in memory: 252 B per row forward (logits, query index in; facts, query, world index, world_h out),
// 248 B backward (logits, query index, grad_world_h, grad_query in; grad_logits out).  The backward RE-DRAWS the
// noise from the same Philox counters instead of saving it, and recomputes the softmax.
//
// lane <-> row, everything in registers, no shared memory, no barriers; rows are fetched with 128-bit loads (a row is
// 2N contiguous floats).
//
// Arithmetic notes.  facts: the same facts_row as the stand-alone kernel (bit-identical).  worlds and query:
// __fmul_rn products folded in world order under the per-row query mask (bit-identical to ad2_forward_kernel).
// Ties of the arg-max go to the lowest world index.
// Adjoint: with y = softmax((logit + g)/tau), G_e = gh[i] + gh[N+k], d logit_e = y_e (G_e - sum y G) / tau;
// d/d p1[i] of sum_e d logit_e log(p1[i] p2[k]) = (sum_k d logit_ik) / p1[i] — no division per world;
// the query adds gq [e in query] p2[k] (resp. p1[i]); then the facts-head adjoint (detached normaliser).
//
// Two variants per direction.  GENERAL keeps the W perturbed logits in registers (two passes: max, then exp2 in the
// log2 domain) and takes any noise source, any tau, and produces y_soft: with INJECTED noise it is the reference's
// computation on the caller's Gumbel variates (logit(w) = log p1[i] + log p2[k]: 2N fast logarithms instead of W,
// absolute error <= 3e-6, y_soft within 1e-5 relative of the reference's).  FAST is the production path (in-kernel
// Philox noise, tau >= 0.5, no y_soft): the race times behind the Gumbel-max trick are drawn TOP-DOWN (see
// draw_world below), so the forward draws two uniforms per row and touches no world but the query's members, and the
// adjoint needs per-column sums of the softmax weights only — nothing W-sized lives in registers.
#include <cstdlib>
#include <type_traits>

#include "params.cuh"
#include "sampling.cuh"

namespace vael {

// ---- in-kernel noise (Philox mode): exponential races, drawn top-down ------------------------------------------------
// F.gumbel_softmax perturbs logit_w = log p_w with g_w = -log E_w, E_w ~ Exp(1): the perturbed logit is -log t_w with
// t_w = E_w / p_w, the sampled world is the arg-MIN of the race times t_w and y_soft_w is proportional to t_w^(-1/tau).
// The times can be drawn in another order with the same joint law (the Gumbel-max trick read backwards, Maddison et al.
// 2014): the winner w* ~ Categorical(p) — for the product distribution p_w = p1[i] p2[k] that is one inverse-CDF
// draw per annotated disjunction —, its time t* ~ Exp(1) / sum_w p_w, and by memorylessness every other
// t_w = t* + E'_w / p_w with fresh E'_w ~ Exp(1).  The FORWARD then needs two uniforms per row instead of W
// variates (the straight-through forward value (1 - y) + y of the selected component is 1 up to one ulp and is
// written as 1); only the adjoint, which needs every y_soft_w, walks the W worlds.  Counters: block e / 4 of a row
// holds the variates of worlds 4 (e / 4) .. + 3, block ceil(W / 4) the row's three draws; key = seed.
__device__ __forceinline__ uint4 philox_block(long long row, int blk, uint64_t seed) {
  return philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)blk, 0x5eedu),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

struct RowDraw {
  int arg;       // the sampled world i* N + k*
  float t_min;   // its race time, in units of ln 2 (t / ln 2): every other world's is t_min + E2'_w / p_w
};
template <int N>
__device__ __forceinline__ RowDraw draw_world(const float* f, long long row, uint64_t seed) {
  const uint4 r = philox_block(row, (N * N + 3) / 4, seed);
  float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) { s1 += f[i]; s2 += f[N + i]; }
  // inverse CDF per annotated disjunction: the number of proper prefix sums the target has reached
  const float ta = unit_from_23_bits(r.x >> 9) * s1, tb = unit_from_23_bits(r.y >> 9) * s2;
  int ia = 0, ka = 0;
  float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
  for (int i = 0; i < N - 1; ++i) {
    c1 += f[i]; c2 += f[N + i];
    ia += (ta >= c1) ? 1 : 0;
    ka += (tb >= c2) ? 1 : 0;
  }
  RowDraw d;
  d.arg = ia * N + ka;
  d.t_min = exp2_variate(r.z) / (s1 * s2);
  return d;
}

template <int F>
__device__ __forceinline__ void load_row(const float* __restrict__ base, long long row, float* x, bool al16) {
  const float* src = base + row * F;
  if (F % 4 == 0 && al16) {
#pragma unroll
    for (int c = 0; c < F / 4; ++c) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src) + c);
      x[4 * c] = t.x; x[4 * c + 1] = t.y; x[4 * c + 2] = t.z; x[4 * c + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < F; ++c) x[c] = __ldg(src + c);
  }
}
template <int F>
__device__ __forceinline__ void store_row(float* __restrict__ base, long long row, const float* x, bool al16) {
  float* dst = base + row * F;
  if (F % 4 == 0 && al16) {
#pragma unroll
    for (int c = 0; c < F / 4; ++c)
      reinterpret_cast<float4*>(dst)[c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
  } else {
#pragma unroll
    for (int c = 0; c < F; ++c) dst[c] = x[c];
  }
}

// GENERAL: perturbed logits z[e] of one row in the log2 domain (up to a row constant), their maximum and arg-max; the
// query sum under mask `m` when `qsum` is non-null.  Noise: injected (`noise` [B,W], z = s (log p + g): the reference's
// computation on the caller's Gumbel variates) or Philox (race times drawn top-down, z = -log2(t) / tau).
template <int N>
__device__ __forceinline__ void perturbed_logits(const FusedParams& P, long long row, uint64_t seed, const float* f,
                                                 const uint32_t* m, float inv_tau, float* z, float& mx, int& arg,
                                                 float* qsum) {
  constexpr int W = N * N;
  float acc = 0.0f;
  if (P.noise != nullptr) {
    const float s = inv_tau * kLog2e;
    float lps[2 * N];
#pragma unroll
    for (int c = 0; c < 2 * N; ++c) lps[c] = s * __logf(f[c]);   // absolute error of the log <= 3e-6: noise-level
    const bool nz_al = (W % 4 == 0) && aligned16(P.noise);
    mx = -INFINITY;
    arg = 0;
    float g4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int e = 0; e < W; ++e) {
      const int i = e / N, k = e % N;
      if ((e & 3) == 0) {
        if (nz_al) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(P.noise + row * W + e));
          g4[0] = t.x; g4[1] = t.y; g4[2] = t.z; g4[3] = t.w;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) g4[u] = (e + u < W) ? __ldg(P.noise + row * W + e + u) : 0.0f;
        }
      }
      if (qsum != nullptr && ((m[e >> 5] >> (e & 31)) & 1u)) acc = __fadd_rn(acc, __fmul_rn(f[i], f[N + k]));
      z[e] = fmaf(g4[e & 3], s, lps[i] + lps[N + k]);
      if (z[e] > mx) { mx = z[e]; arg = e; }
    }
  } else {
    const RowDraw d = draw_world<N>(f, row, seed);
    float r1[N], r2[N], E4[4];
#pragma unroll
    for (int i = 0; i < N; ++i) { r1[i] = fast_rcp(f[i]); r2[i] = fast_rcp(f[N + i]); }
#pragma unroll
    for (int e = 0; e < W; ++e) {
      const int i = e / N, k = e % N;
      if ((e & 3) == 0) exp2_variates4(philox_block(row, e >> 2, seed), E4);
      if (qsum != nullptr && ((m[e >> 5] >> (e & 31)) & 1u)) acc = __fadd_rn(acc, __fmul_rn(f[i], f[N + k]));
      float t = fmaf(E4[e & 3], r1[i] * r2[k], d.t_min);
      t = (e == d.arg) ? d.t_min : t;
      z[e] = -inv_tau * fast_lg2(t);
    }
    arg = d.arg;
    mx = -inv_tau * fast_lg2(d.t_min);   // = z[arg]: every other race time is t_min plus something non-negative
  }
  if (qsum != nullptr) *qsum = acc;
}

template <int MW>
__device__ __forceinline__ void fused_mask(uint32_t* m, const FusedParams& P, long long row) {
#pragma unroll
  for (int i = 0; i < MW; ++i) m[i] = 0u;
  const int32_t q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
  if (q >= 0 && q < P.n_queries) {
#pragma unroll
    for (int i = 0; i < MW; ++i) m[i] = __ldg(P.qmask + (long long)q * MW + i);
  }
}

// FMINB: resident CTAs per SM the register allocation is capped for (FAST: 6 -> 80 registers, the throughput shape;
// 1 -> no cap: more of the 25 independent Philox chains in flight per lane, the latency shape for small batches)
template <int N, bool FAST, int FMINB = (FAST ? 6 : 1)>
__global__ void __launch_bounds__(128, FMINB) fused_forward_kernel(const FusedParams P) {
  constexpr int F = 2 * N, W = N * N, MW = (W + 31) / 32;
  const uint64_t seed = P.seed_dev ? *P.seed_dev : P.seed;
  const float inv_tau = 1.0f / P.tau;
  const bool in_al = aligned16(P.logits);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < P.B; row += stride) {
    float x[F], y[F], rz[2], zs[2], f[F];
    load_row<F>(P.logits, row, x, in_al);
    facts_row<2, N>(x, P.eps, y, rz, zs);
#pragma unroll
    for (int c = 0; c < F; ++c) f[c] = (y[c] + P.eps) * rz[c / N];
    if (P.facts != nullptr) store_row<F>(P.facts, row, f, aligned16(P.facts));
    uint32_t m[MW];
    fused_mask<MW>(m, P, row);
    float ys;   // y_soft of the selected world
    int arg;
    if constexpr (FAST) {
      // two uniforms pick the world; no variate is drawn per world (the adjoint draws them)
      arg = draw_world<N>(f, row, seed).arg;
      if (P.query != nullptr) {
        float q = 0.0f;
#pragma unroll
        for (int e = 0; e < W; ++e)
          if ((m[e >> 5] >> (e & 31)) & 1u) q = __fadd_rn(q, __fmul_rn(f[e / N], f[N + e % N]));
        P.query[row] = q;
      }
      ys = 1.0f;   // stands for any y_soft: (1 - y) + y is written as 1 (the reference's value is 1 or 1 - 2^-24)
    } else {
      float z[W], mx, q;
      perturbed_logits<N>(P, row, seed, f, m, inv_tau, z, mx, arg, &q);
      if (P.query != nullptr) P.query[row] = q;
      float sum = 0.0f;
#pragma unroll
      for (int e = 0; e < W; ++e) {
        z[e] = fast_exp2(z[e] - mx);
        sum += z[e];
      }
      const float inv_sum = 1.0f / sum;
      if (P.y_soft != nullptr) {
#pragma unroll
        for (int e = 0; e < W; ++e) P.y_soft[row * W + e] = z[e] * inv_sum;
      }
      ys = 1.0f * inv_sum;   // exp2(0) / sum
    }
    // straight-through forward value of the selected component: (1 - y_soft) + y_soft
    const float one = (1.0f - ys) + ys;
    if (P.world_idx != nullptr) P.world_idx[row] = arg;
    if (P.world_h != nullptr) {
      const int ia = arg / N, ka = arg - ia * N;
      float h[F];
#pragma unroll
      for (int c = 0; c < F; ++c) h[c] = (c == ia || c == N + ka) ? one : 0.0f;
      store_row<F>(P.world_h, row, h, aligned16(P.world_h));
    }
  }
}

// The FAST forward for streaming batches: the same row computation, with the three 80-byte-per-row streams (logits in;
// facts and world_h out) moved as [32, 2N] tiles by bulk copies through warp-private rings (WarpTileLoader) instead of
// per-lane 128-bit global accesses.  The per-thread kernel above touches 20 different 128-byte lines with every
// warp-wide row access: L1/TEX 70 % busy, the limiter at 0.64 of HBM (profiles/r02i_fused_topdown_ncu.md).
constexpr int kFusedTileStages = 2;
template <int F>
__host__ __device__ constexpr int fused_tile_bytes() { return (kWarp * F * 4 + 127) / 128 * 128; }
template <int F>
__host__ __device__ constexpr int fused_warp_bytes() { return 128 + 3 * kFusedTileStages * fused_tile_bytes<F>(); }

template <int N>
__global__ void __launch_bounds__(128) fused_forward_tiles_kernel(const FusedParams P) {
  constexpr int F = 2 * N, W = N * N, MW = (W + 31) / 32, FT = kWarp * F, TB = fused_tile_bytes<F>();
  extern __shared__ __align__(128) uint8_t fused_smem[];
  const uint64_t seed = P.seed_dev ? *P.seed_dev : P.seed;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint8_t* wb = fused_smem + (size_t)warp * fused_warp_bytes<F>();
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* istage = reinterpret_cast<float*>(wb + 128);
  float* fstage = istage + kFusedTileStages * (TB / 4);
  float* hstage = fstage + kFusedTileStages * (TB / 4);
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool in_al = aligned16(P.logits);
  const bool f_al = P.facts != nullptr && aligned16(P.facts), h_al = P.world_h != nullptr && aligned16(P.world_h);
  const int n_out = (P.facts != nullptr ? 1 : 0) + (P.world_h != nullptr ? 1 : 0);
  WarpTileLoader<kFusedTileStages> lx;
  lx.init(istage, bars, TB / 4, lane);
  auto request = [&](long long sub, int s) {
    if (sub < n_sub) lx.issue(s, P.logits + sub * FT, (int)min((long long)kWarp, P.B - sub * kWarp) * F, in_al, lane, 0.0f);
  };
#pragma unroll
  for (int s = 0; s < kFusedTileStages; ++s) request(gw + s * GW, s);
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % kFusedTileStages);
    const int rows = (int)min((long long)kWarp, P.B - sub * kWarp);
    const long long row = sub * kWarp + lane;
    const bool valid = lane < rows;
    lx.wait(s);
    float x[F], y[F], rz[2], f[F];
    tile_row_load<F>(lx.stage(s) + lane * F, x);
    __syncwarp();
    request(sub + kFusedTileStages * GW, s);
    facts_row<2, N>(x, P.eps, y, rz);
#pragma unroll
    for (int c = 0; c < F; ++c) f[c] = (y[c] + P.eps) * rz[c / N];
    uint32_t m[MW];
#pragma unroll
    for (int i = 0; i < MW; ++i) m[i] = 0u;
    if (valid) fused_mask<MW>(m, P, row);
    const int arg = draw_world<N>(f, row, seed).arg;
    if (P.query != nullptr) {
      float q = 0.0f;
#pragma unroll
      for (int e = 0; e < W; ++e)
        if ((m[e >> 5] >> (e & 31)) & 1u) q = __fadd_rn(q, __fmul_rn(f[e / N], f[N + e % N]));
      if (valid) P.query[row] = q;
    }
    if (P.world_idx != nullptr && valid) P.world_idx[row] = arg;
    // the stores of two tiles ago have read their stages (n_out bulk groups per tile)
    if (lane == 0) {
      if (n_out == 2) bulk_wait_read<2>();
      else if (n_out == 1) bulk_wait_read<1>();
    }
    __syncwarp();
    if (P.facts != nullptr) tile_row_store<F>(fstage + s * (TB / 4) + lane * F, f);
    if (P.world_h != nullptr) {
      const int ia = arg / N, ka = arg - ia * N;
      float h[F];
#pragma unroll
      for (int c = 0; c < F; ++c) h[c] = (c == ia || c == N + ka) ? 1.0f : 0.0f;   // (1 - y) + y written as 1
      tile_row_store<F>(hstage + s * (TB / 4) + lane * F, h);
    }
    fence_proxy_async();
    __syncwarp();
    if (P.facts != nullptr) warp_tile_store(P.facts + sub * FT, fstage + s * (TB / 4), rows * F, FT, f_al, lane);
    if (P.world_h != nullptr) warp_tile_store(P.world_h + sub * FT, hstage + s * (TB / 4), rows * F, FT, h_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
}

// MINB: resident CTAs per SM the register allocation is capped for (4 -> 128 registers with a few spilled words,
// 1 -> 168 registers, no spills); VAEL_FUSED_BWD_MINB picks, the default is the measured optimum
template <int N, bool FAST, int MINB = 1>
__global__ void __launch_bounds__(128, MINB) fused_backward_kernel(const FusedParams P) {
  constexpr int F = 2 * N, W = N * N, MW = (W + 31) / 32;
  const uint64_t seed = P.seed_dev ? *P.seed_dev : P.seed;
  const float inv_tau = 1.0f / P.tau;
  const bool in_al = aligned16(P.logits);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < P.B; row += stride) {
    float x[F], y[F], rz[2], zs[2], f[F];
    load_row<F>(P.logits, row, x, in_al);
    facts_row<2, N>(x, P.eps, y, rz, zs);
#pragma unroll
    for (int c = 0; c < F; ++c) f[c] = (y[c] + P.eps) * rz[c / N];
    uint32_t m[MW];
    fused_mask<MW>(m, P, row);
    float df[F];
#pragma unroll
    for (int c = 0; c < F; ++c) df[c] = 0.0f;
    if (P.grad_world_h != nullptr && FAST) {
      // Softmax weights q_e = t_e^(-1/tau) of the race times (no maximum to subtract: t_e >= t_min > 0 and, for
      // tau >= 0.5, t^(-1/tau) stays inside fp32).  P[c] = sum of q over the worlds of column c, and A[c] = sum of
      // q_e G_e with G_e = gh[i] + gh[N + k] splits into gh[c] P[c] plus the cross term X[i] = sum_k q_ik gh[N + k]
      // (resp. X[N + k] = sum_i q_ik gh[i]): two additions and two FMAs per world.
      float gh[F], r1[N], r2[N], Pc[F], Xc[F];
      load_row<F>(P.grad_world_h, row, gh, aligned16(P.grad_world_h));
      const RowDraw d = draw_world<N>(f, row, seed);
#pragma unroll
      for (int i = 0; i < N; ++i) { r1[i] = fast_rcp(f[i]); r2[i] = fast_rcp(f[N + i]); }
#pragma unroll
      for (int c = 0; c < F; ++c) Pc[c] = Xc[c] = 0.0f;
      float E4[4];
      auto walk = [&](auto tau_is_one) {
#pragma unroll
        for (int e = 0; e < W; ++e) {
          const int i = e / N, k = e % N;
          if ((e & 3) == 0) exp2_variates4(philox_block(row, e >> 2, seed), E4);
          float t = fmaf(E4[e & 3], r1[i] * r2[k], d.t_min);
          t = (e == d.arg) ? d.t_min : t;
          float qe;
          if constexpr (decltype(tau_is_one)::value) qe = fast_rcp(t);
          else qe = fast_exp2(-inv_tau * fast_lg2(t));
          Pc[i] += qe; Pc[N + k] += qe;
          Xc[i] = fmaf(qe, gh[N + k], Xc[i]);
          Xc[N + k] = fmaf(qe, gh[i], Xc[N + k]);
        }
      };
      if (inv_tau == 1.0f) walk(std::true_type{});
      else walk(std::false_type{});
      float S = 0.0f, D = 0.0f;
#pragma unroll
      for (int i = 0; i < N; ++i) { S += Pc[i]; D += fmaf(gh[i], Pc[i], Xc[i]); }
      const float dot = D / S, scale = inv_tau / S;
#pragma unroll
      for (int c = 0; c < F; ++c) df[c] = fmaf(gh[c] - dot, Pc[c], Xc[c]) * scale * (c < N ? r1[c] : r2[c - N]);
    } else if (P.grad_world_h != nullptr) {
      float z[W], mx;
      int arg;
      perturbed_logits<N>(P, row, seed, f, m, inv_tau, z, mx, arg, nullptr);
      float gh[F];
      load_row<F>(P.grad_world_h, row, gh, aligned16(P.grad_world_h));
      float S = 0.0f, D = 0.0f;
#pragma unroll
      for (int e = 0; e < W; ++e) {
        z[e] = fast_exp2(z[e] - mx);
        S += z[e];
        D = fmaf(z[e], gh[e / N] + gh[N + e % N], D);
      }
      const float dot = D / S;
      float a[F];
#pragma unroll
      for (int c = 0; c < F; ++c) a[c] = 0.0f;
#pragma unroll
      for (int e = 0; e < W; ++e) {
        const float t = z[e] * ((gh[e / N] + gh[N + e % N]) - dot);
        a[e / N] += t;
        a[N + e % N] += t;
      }
      const float scale = inv_tau / S;
#pragma unroll
      for (int c = 0; c < F; ++c) df[c] = a[c] * scale / f[c];
    }
    if (P.grad_query != nullptr) {
      const float gq = __ldg(P.grad_query + row);
#pragma unroll
      for (int e = 0; e < W; ++e) {
        if ((m[e >> 5] >> (e & 31)) & 1u) {
          df[e / N] = fmaf(gq, f[N + e % N], df[e / N]);
          df[N + e % N] = fmaf(gq, f[e / N], df[N + e % N]);
        }
      }
    }
    if (P.grad_facts != nullptr) {
      float gf[F];
      load_row<F>(P.grad_facts, row, gf, aligned16(P.grad_facts));
#pragma unroll
      for (int c = 0; c < F; ++c) df[c] += gf[c];
    }
    // facts head: f = (y + eps) / Z with Z detached -> dy_i = df_i / Z; softmax: dx_i = y_i (dy_i - sum_j y_j dy_j)
    if constexpr (FAST) {   // y was not kept across the world loop (registers): y = f Z - eps, one rounding away
#pragma unroll
      for (int c = 0; c < F; ++c) y[c] = fmaf(f[c], zs[c / N], -P.eps);
    }
    float dx[F];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float dy[N], dot = 0.0f;
#pragma unroll
      for (int i = 0; i < N; ++i) { dy[i] = df[g * N + i] * rz[g]; dot = fmaf(y[g * N + i], dy[i], dot); }
#pragma unroll
      for (int i = 0; i < N; ++i) dx[g * N + i] = y[g * N + i] * (dy[i] - dot);
    }
    store_row<F>(P.grad_logits, row, dx, aligned16(P.grad_logits));
  }
}

bool fused_supported(int na, int nb) { return na == nb && (na == 10 || na == 9); }

// one row per thread.  Small batches (the B = 32 .. 4096 training configurations) use one-warp CTAs so that the
// rows spread over as many SMs as possible — a row is a long serial chain; large batches use four-warp CTAs.
static void fused_shape(long long B, int num_sms, int ctas_per_sm, int* grid, int* block) {
  *block = (B <= (long long)num_sms * 128) ? 32 : 128;
  // one row per thread, as many CTAs as rows need: no grid-stride loop (adjoint 0.216 -> 0.197 ms per Mi rows, forward
  // 0.063 -> 0.059; VAEL_FUSED_PERSISTENT=1 restores the resident-CTAs-only grid)
  static const bool persistent = getenv("VAEL_FUSED_PERSISTENT") != nullptr;
  long long g = (B + *block - 1) / *block;
  if (persistent) g = min(g, (long long)num_sms * ctas_per_sm * (128 / *block));
  *grid = (int)max(1LL, min(g, 0x7fffffffLL));
}

// FAST needs bounded (Philox) noise and a tau that keeps the softmax terms inside fp32's range; y_soft is a
// GENERAL-only output (it needs every term after the sum is known)
static bool fused_fast_ok(const FusedParams& P) {
  static const bool off = getenv("VAEL_FUSED_GENERAL") != nullptr;   // tests / A-B timing
  return !off && P.noise == nullptr && P.tau >= 0.5f && P.y_soft == nullptr;
}

template <bool BWD, bool FAST>
static cudaError_t fused_launch(int n, const FusedParams& P, int grid, int block, cudaStream_t st) {
  static const int minb_env = getenv("VAEL_FUSED_BWD_MINB") ? atoi(getenv("VAEL_FUSED_BWD_MINB")) : 0;
  static const int fminb_env = getenv("VAEL_FUSED_FWD_MINB") ? atoi(getenv("VAEL_FUSED_FWD_MINB")) : 0;
  const bool small = block == 32;   // fused_shape: fewer rows than one four-warp CTA per SM would take
  const int minb = minb_env ? minb_env : (small ? 1 : 4), fminb = fminb_env ? fminb_env : (small ? 1 : 6);
  if (n == 10) {
    if (!BWD && FAST && fminb == 1) fused_forward_kernel<10, FAST, 1><<<grid, block, 0, st>>>(P);
    else if (BWD && FAST && minb == 4) fused_backward_kernel<10, FAST, 4><<<grid, block, 0, st>>>(P);
    else if (BWD && FAST && minb == 5) fused_backward_kernel<10, FAST, 5><<<grid, block, 0, st>>>(P);
    else if (BWD && FAST && minb == 3) fused_backward_kernel<10, FAST, 3><<<grid, block, 0, st>>>(P);
    else if (BWD) fused_backward_kernel<10, FAST><<<grid, block, 0, st>>>(P);
    else fused_forward_kernel<10, FAST><<<grid, block, 0, st>>>(P);
  } else if (n == 9) {
    if (BWD) fused_backward_kernel<9, FAST><<<grid, block, 0, st>>>(P);
    else fused_forward_kernel<9, FAST><<<grid, block, 0, st>>>(P);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <int N>
static cudaError_t fused_forward_tiles(const FusedParams& P, int num_sms, cudaStream_t st) {
  static const int nw_env = getenv("VAEL_FUSED_TILE_WARPS") ? atoi(getenv("VAEL_FUSED_TILE_WARPS")) : 0;
  // measured on B200 (1 Mi rows): 2 warps per CTA 0.0586 ms, 3 0.0622, 4 0.0625; the per-thread kernel 0.0651
  const int nw = nw_env ? max(1, min(4, nw_env)) : 2, smem = nw * fused_warp_bytes<2 * N>();
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const int ctas = max(1, min(16, (227 * 1024) / (smem + 1024)));
  const int grid = (int)tile_grid(n_sub, nw, (long long)num_sms * ctas);
  cudaError_t e = cudaFuncSetAttribute(fused_forward_tiles_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  fused_forward_tiles_kernel<N><<<grid, nw * 32, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t fused_forward(int n, const FusedParams& P, int num_sms, cudaStream_t st) {
  int grid, block;
  const bool fast = fused_fast_ok(P);
  static const bool no_tiles = getenv("VAEL_FUSED_NO_TILES") != nullptr;   // A-B timing
  if (fast && !no_tiles && P.B > (long long)num_sms * 128) {   // streaming batches: tiles by bulk copies
    if (n == 10) return fused_forward_tiles<10>(P, num_sms, st);
    if (n == 9) return fused_forward_tiles<9>(P, num_sms, st);
  }
  fused_shape(P.B, num_sms, fast ? 6 : 3, &grid, &block);   // resident CTAs per SM by registers: 80 / 157
  return fast ? fused_launch<false, true>(n, P, grid, block, st) : fused_launch<false, false>(n, P, grid, block, st);
}
cudaError_t fused_backward(int n, const FusedParams& P, int num_sms, cudaStream_t st) {
  int grid, block;
  const bool fast = fused_fast_ok(P);
  fused_shape(P.B, num_sms, fast ? 4 : 2, &grid, &block);   // 128 / 255 registers
  return fast ? fused_launch<true, true>(n, P, grid, block, st) : fused_launch<true, false>(n, P, grid, block, st);
}

}  // namespace vael
