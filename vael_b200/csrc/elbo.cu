// ELBO terms and their gradients in one pass.
//
// Reference: loss_function(..., rec_loss='LAPLACE') — utils/mnist_utils/train.py:13-55
// (Mario: utils/mario_utils/train.py:11-69 calls the same terms once per frame):
//   recon = -mean_b sum_pix log Laplace(x | recon, 1) = mean_b ( D log 2 + sum_pix |x - recon| )
//   kl    =  mean_b -0.5 sum_l (1 + logvar - mu^2 - exp(logvar))
//   qbce  =  BCELoss(q, 1) = mean_b -max(log q, -100)
//   loss  =  recon_w recon + kl_w kl + query_w qbce
// rec_loss = 'MSE' / 'BCE' (train.py:16-21; the other two values of the reference's option): the reconstruction
// term is the mean over ALL B*D elements of (recon - x)^2, resp. of -(x max(log r, -100) + (1-x) max(log(1-r), -100))
// (torch BCELoss incl. its clamps; gradient (r - x) / max(r (1-r), 1e-12)).
// HBM-bound: reads recon and x once (2*D floats per row) and writes grad_recon.
// The reconstruction term is a sum over ALL pixels of the batch divided by B, so the kernel treats
// recon / x / grad_recon as one flat stream: every CTA takes a contiguous span of it (128-bit loads, four in
// flight per thread), the KL and query terms are grid-stride loops over their [B,L] / [B] elements.  Per-CTA
// partial sums go to a workspace in fp64 and the last CTA to finish adds them in index order, so the
// result is deterministic.
#include <cstdlib>

#include "params.cuh"

namespace vael {

__device__ __forceinline__ float sgn(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

// one element of the reconstruction term: returns its contribution, writes d(term)/d(recon) * gr to *g
template <int REC>
__device__ __forceinline__ float rec_elem(float x, float r, float gr, float* g) {
  if (REC == VAEL_REC_LAPLACE) {
    const float d = r - x;
    *g = gr * sgn(d);
    return fabsf(d);
  } else if (REC == VAEL_REC_MSE) {
    const float d = r - x;
    *g = gr * 2.0f * d;
    return d * d;
  } else {
    *g = gr * (r - x) / fmaxf((1.0f - r) * r, 1e-12f);
    return -(x * fmaxf(logf(r), -100.0f) + (1.0f - x) * fmaxf(logf(1.0f - r), -100.0f));
  }
}

template <int REC>
__device__ __forceinline__ float abs_and_grad4(const float4 x, const float4 r, float gr, float4* g) {
  float4 o;
  const float s = (rec_elem<REC>(x.x, r.x, gr, &o.x) + rec_elem<REC>(x.y, r.y, gr, &o.y)) +
                  (rec_elem<REC>(x.z, r.z, gr, &o.z) + rec_elem<REC>(x.w, r.w, gr, &o.w));
  if (g) *g = o;
  return s;
}

template <int REC>
__global__ void __launch_bounds__(256) elbo_terms_kernel(const ElboParams P) {
  __shared__ double red[3][8];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float invB = 1.0f / (float)P.B;
  const long long n = P.B * P.D;
  // LAPLACE: sum over pixels, mean over the batch; MSE / BCE: mean over all elements
  const float gr = REC == VAEL_REC_LAPLACE ? P.recon_w * invB : P.recon_w / (float)n;
  const float gk = P.kl_w * invB, gq = P.query_w * invB;
  const bool vec = (n % 4 == 0) && aligned16(P.recon) && aligned16(P.x) &&
                   (P.grad_recon == nullptr || aligned16(P.grad_recon));
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gsz = (long long)gridDim.x * blockDim.x;

  // ---- sum |x - recon| over the whole batch; grad_recon = recon_w / B * sign(recon - x) ----
  float s_abs = 0.0f;
  if (vec) {
    const long long n4 = n / 4;
    const long long span = (n4 + gridDim.x - 1) / gridDim.x;   // contiguous float4 span of this CTA
    const long long beg = (long long)blockIdx.x * span, end = min(n4, beg + span);
    const float4* x4 = reinterpret_cast<const float4*>(P.x);
    const float4* r4 = reinterpret_cast<const float4*>(P.recon);
    float4* g4 = reinterpret_cast<float4*>(P.grad_recon);
    long long i = beg + tid;
    for (; i + 3 * 256 < end; i += 4 * 256) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = __ldg(x4 + i + u * 256); b[u] = __ldg(r4 + i + u * 256); }
#pragma unroll
      for (int u = 0; u < 4; ++u) s_abs += abs_and_grad4<REC>(a[u], b[u], gr, g4 ? g4 + i + u * 256 : nullptr);
    }
    for (; i < end; i += 256) s_abs += abs_and_grad4<REC>(__ldg(x4 + i), __ldg(r4 + i), gr, g4 ? g4 + i : nullptr);
  } else {
    for (long long i = gtid; i < n; i += gsz) {
      float g;
      s_abs += rec_elem<REC>(__ldg(P.x + i), __ldg(P.recon + i), gr, &g);
      if (P.grad_recon) P.grad_recon[i] = g;
    }
  }
  // ---- KL over [B,L] ----
  float s_kl = 0.0f;
  for (long long i = gtid; i < P.B * P.L; i += gsz) {
    const float m = __ldg(P.mu + i), lv = __ldg(P.logvar + i);
    const float ev = expf(lv);
    s_kl += 1.0f + lv - m * m - ev;
    if (P.grad_mu) P.grad_mu[i] = gk * m;
    if (P.grad_logvar) P.grad_logvar[i] = gk * 0.5f * (ev - 1.0f);
  }
  // ---- query BCE over [B] ----
  float s_q = 0.0f;
  if (P.query_prob != nullptr) {
    for (long long i = gtid; i < P.B; i += gsz) {
      const float q = __ldg(P.query_prob + i);
      s_q += -fmaxf(logf(q), -100.0f);
      // torch binary_cross_entropy backward with target 1: (q - 1) / max((1 - q) q, 1e-12)
      if (P.grad_query) P.grad_query[i] = gq * (q - 1.0f) / fmaxf((1.0f - q) * q, 1e-12f);
    }
  }

  double da = (double)s_abs, dk = (double)s_kl, dq = (double)s_q;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    da += __shfl_xor_sync(0xffffffffu, da, o);
    dk += __shfl_xor_sync(0xffffffffu, dk, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
  if (lane == 0) { red[0][warp] = da; red[1][warp] = dk; red[2][warp] = dq; }
  __syncthreads();
  if (tid == 0) {
    double ta = 0.0, tk = 0.0, tq = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += red[0][w]; tk += red[1][w]; tq += red[2][w]; }
    P.partials[blockIdx.x * 3 + 0] = ta;
    P.partials[blockIdx.x * 3 + 1] = tk;
    P.partials[blockIdx.x * 3 + 2] = tq;
    __threadfence();
    const unsigned int done = atomicAdd(P.counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    // every partial is final; the 256 threads of the last CTA add them in a FIXED pattern (thread t takes
    // partials t, t + 256, ...; then a shuffle tree; then the eight warp sums in order), so the result does
    // not depend on which CTA happened to finish last
    __threadfence();
    double r = 0.0, k = 0.0, q = 0.0;
    const volatile double* part = P.partials;
    for (unsigned int b = tid; b < gridDim.x; b += blockDim.x) {
      r += part[b * 3 + 0];
      k += part[b * 3 + 1];
      q += part[b * 3 + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r += __shfl_xor_sync(0xffffffffu, r, o);
      k += __shfl_xor_sync(0xffffffffu, k, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) { red[0][warp] = r; red[1][warp] = k; red[2][warp] = q; }
    __syncthreads();
    if (tid == 0) {
      r = k = q = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { r += red[0][w]; k += red[1][w]; q += red[2][w]; }
      // LAPLACE: recon = mean_b ( D log 2 + sum_pix |x - recon| ); MSE / BCE: mean over all B*D elements
      const float recon = REC == VAEL_REC_LAPLACE ? (float)((double)P.D * 0.69314718055994531 + r / (double)P.B)
                                                  : (float)(r / ((double)P.B * (double)P.D));
      const float kl = (float)(-0.5 * k / (double)P.B), qb = (float)(q / (double)P.B);
      P.terms[1] = recon;
      P.terms[2] = kl;
      P.terms[3] = qb;
      P.terms[0] = P.recon_w * recon + P.kl_w * kl + P.query_w * qb;
      *P.counter = 0u;  // self-cleaning for the next call
    }
  }
}

constexpr int kElboMaxGrid = 8192;
size_t elbo_workspace_bytes() { return 16 + (size_t)kElboMaxGrid * 3 * sizeof(double); }

cudaError_t elbo_terms(ElboParams P, void* workspace, int num_sms, cudaStream_t st) {
  P.counter = reinterpret_cast<unsigned int*>(workspace);
  P.partials = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + 16);
  // enough CTAs to fill the chip, never more than one per 1024 float4s of the pixel stream
  const long long work = max(1LL, (P.B * P.D) / 4096);
  static const int grid_env = getenv("VAEL_ELBO_GRID") ? atoi(getenv("VAEL_ELBO_GRID")) : 0;   // experiments
  // more, shorter contiguous spans stream faster (B = 65 536: 1184 CTAs 5.89 TB/s, 2048 6.14, 4096 6.21, 8192 6.28)
  const int cap = grid_env > 0 ? min(kElboMaxGrid, grid_env) : kElboMaxGrid;
  const int grid = (int)max(1LL, min(work, (long long)cap));
  if (P.rec_loss == VAEL_REC_MSE) elbo_terms_kernel<VAEL_REC_MSE><<<grid, 256, 0, st>>>(P);
  else if (P.rec_loss == VAEL_REC_BCE) elbo_terms_kernel<VAEL_REC_BCE><<<grid, 256, 0, st>>>(P);
  else elbo_terms_kernel<VAEL_REC_LAPLACE><<<grid, 256, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace vael
