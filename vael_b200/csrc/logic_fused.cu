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
// Gumbel noise from the same Philox counters instead of saving it, and recomputes the softmax.
//
// lane <-> row, everything in registers (the W = N*N perturbed logits are a fully unrolled register array), no
// shared memory, no barriers; rows are fetched with 128-bit loads (a row is 2N contiguous floats).  The kernel is
// bound by FP32 / SFU issue, not by HBM: per row W x (a quarter Philox4x32-10 call, the Gumbel transform,
// one expf) plus 2N logf — about 5.5 k instructions against 252 B.
//
// Arithmetic notes.  facts: the same facts_row as the stand-alone kernel (bit-identical).  worlds and query:
// __fmul_rn products folded in world order under the per-row query mask (bit-identical to ad2_forward_kernel).
// logit(w) = log p1[i] + log p2[k] (2N logarithms instead of W; the reference takes log of the fp32 product —
// both are within an ulp of the exact value).  Ties of the arg-max go to the lowest world index.
// Adjoint: with y = softmax((logit + g)/tau), G_e = gh[i] + gh[N+k], d logit_e = y_e (G_e - sum y G) / tau;
// d/d p1[i] of sum_e d logit_e log(p1[i] p2[k]) = (sum_k d logit_ik) / p1[i] — no division per world;
// the query adds gq [e in query] p2[k] (resp. p1[i]); then the facts-head adjoint (detached normaliser).
#include "params.cuh"
#include "sampling.cuh"

namespace vael {

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

// Perturbed logits z[e] = (log p1[i] + log p2[k] + g_e) / tau of one row, its maximum and arg-max; the query sum
// under mask `m` when `qsum` is non-null.  The noise is injected (`noise`, [B,W]) or drawn from Philox4x32-10:
// one call per four consecutive worlds, counter = (row, e / 4), key = seed.
template <int N>
__device__ __forceinline__ void perturbed_logits(const FusedParams& P, long long row, uint64_t seed, const float* f,
                                                 const uint32_t* m, float inv_tau, float* z, float& mx, int& arg,
                                                 float* qsum) {
  constexpr int W = N * N;
  float lp[2 * N];
#pragma unroll
  for (int c = 0; c < 2 * N; ++c) lp[c] = logf(f[c]);
  const bool nz_al = P.noise != nullptr && (W % 4 == 0) && aligned16(P.noise);
  mx = -INFINITY;
  arg = 0;
  float acc = 0.0f;
  float g4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int e = 0; e < W; ++e) {
    const int i = e / N, k = e % N;
    if ((e & 3) == 0) {
      if (P.noise != nullptr) {
        if (nz_al) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(P.noise + row * W + e));
          g4[0] = t.x; g4[1] = t.y; g4[2] = t.z; g4[3] = t.w;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) g4[u] = (e + u < W) ? __ldg(P.noise + row * W + e + u) : 0.0f;
        }
      } else {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(e >> 2), 0x5eedu),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        g4[0] = gumbel_from_bits(r.x); g4[1] = gumbel_from_bits(r.y);
        g4[2] = gumbel_from_bits(r.z); g4[3] = gumbel_from_bits(r.w);
      }
    }
    if (qsum != nullptr && ((m[e >> 5] >> (e & 31)) & 1u)) acc = __fadd_rn(acc, __fmul_rn(f[i], f[N + k]));
    z[e] = (lp[i] + lp[N + k] + g4[e & 3]) * inv_tau;
    if (z[e] > mx) { mx = z[e]; arg = e; }
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

template <int N>
__global__ void __launch_bounds__(128) fused_forward_kernel(const FusedParams P) {
  constexpr int F = 2 * N, W = N * N, MW = (W + 31) / 32;
  const uint64_t seed = P.seed_dev ? *P.seed_dev : P.seed;
  const float inv_tau = 1.0f / P.tau;
  const bool in_al = aligned16(P.logits);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < P.B; row += stride) {
    float x[F], y[F], zs[2], f[F];
    load_row<F>(P.logits, row, x, in_al);
    facts_row<2, N>(x, P.eps, y, zs);
#pragma unroll
    for (int c = 0; c < F; ++c) f[c] = (y[c] + P.eps) / zs[c / N];
    if (P.facts != nullptr) store_row<F>(P.facts, row, f, aligned16(P.facts));
    uint32_t m[MW];
    fused_mask<MW>(m, P, row);
    float z[W], mx, q;
    int arg;
    perturbed_logits<N>(P, row, seed, f, m, inv_tau, z, mx, arg, &q);
    if (P.query != nullptr) P.query[row] = q;
    float sum = 0.0f;
#pragma unroll
    for (int e = 0; e < W; ++e) {
      z[e] = expf(z[e] - mx);
      sum += z[e];
    }
    const float inv_sum = 1.0f / sum;
    if (P.y_soft != nullptr) {
#pragma unroll
      for (int e = 0; e < W; ++e) P.y_soft[row * W + e] = z[e] * inv_sum;
    }
    // straight-through forward value of the selected component: (1 - y_soft) + y_soft, y_soft[arg] = exp(0) / sum
    const float ys = 1.0f * inv_sum;
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

template <int N>
__global__ void __launch_bounds__(128) fused_backward_kernel(const FusedParams P) {
  constexpr int F = 2 * N, W = N * N, MW = (W + 31) / 32;
  const uint64_t seed = P.seed_dev ? *P.seed_dev : P.seed;
  const float inv_tau = 1.0f / P.tau;
  const bool in_al = aligned16(P.logits);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < P.B; row += stride) {
    float x[F], y[F], zs[2], f[F];
    load_row<F>(P.logits, row, x, in_al);
    facts_row<2, N>(x, P.eps, y, zs);
#pragma unroll
    for (int c = 0; c < F; ++c) f[c] = (y[c] + P.eps) / zs[c / N];
    uint32_t m[MW];
    fused_mask<MW>(m, P, row);
    float df[F];
#pragma unroll
    for (int c = 0; c < F; ++c) df[c] = 0.0f;
    if (P.grad_world_h != nullptr) {
      float z[W], mx;
      int arg;
      perturbed_logits<N>(P, row, seed, f, m, inv_tau, z, mx, arg, nullptr);
      float gh[F];
      load_row<F>(P.grad_world_h, row, gh, aligned16(P.grad_world_h));
      float S = 0.0f, D = 0.0f;
#pragma unroll
      for (int e = 0; e < W; ++e) {
        z[e] = expf(z[e] - mx);
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
    float dx[F];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float dy[N], dot = 0.0f;
#pragma unroll
      for (int i = 0; i < N; ++i) { dy[i] = df[g * N + i] / zs[g]; dot = fmaf(y[g * N + i], dy[i], dot); }
#pragma unroll
      for (int i = 0; i < N; ++i) dx[g * N + i] = y[g * N + i] * (dy[i] - dot);
    }
    store_row<F>(P.grad_logits, row, dx, aligned16(P.grad_logits));
  }
}

bool fused_supported(int na, int nb) { return na == nb && (na == 10 || na == 9); }

// one row per thread.  Small batches (the B = 32 .. 4096 training configurations) use one-warp CTAs so that the
// rows spread over as many SMs as possible — a row is a ~5 k-instruction serial chain; large batches use
// four-warp CTAs, a few per SM, grid-striding.
static void fused_shape(long long B, int num_sms, int* grid, int* block) {
  *block = (B <= (long long)num_sms * 128) ? 32 : 128;
  *grid = (int)max(1LL, min((B + *block - 1) / *block, (long long)num_sms * 4 * (128 / *block)));
}

cudaError_t fused_forward(int n, const FusedParams& P, int num_sms, cudaStream_t st) {
  int grid, block;
  fused_shape(P.B, num_sms, &grid, &block);
  if (n == 10) fused_forward_kernel<10><<<grid, block, 0, st>>>(P);
  else if (n == 9) fused_forward_kernel<9><<<grid, block, 0, st>>>(P);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t fused_backward(int n, const FusedParams& P, int num_sms, cudaStream_t st) {
  int grid, block;
  fused_shape(P.B, num_sms, &grid, &block);
  if (n == 10) fused_backward_kernel<10><<<grid, block, 0, st>>>(P);
  else if (n == 9) fused_backward_kernel<9><<<grid, block, 0, st>>>(P);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace vael
