// ELBO terms and their gradients in one pass.
//
// Reference: loss_function(..., rec_loss='LAPLACE') — utils/mnist_utils/train.py:13-55
// (Mario: utils/mario_utils/train.py:11-69 calls the same terms once per frame):
//   recon = -mean_b sum_pix log Laplace(x | recon, 1) = mean_b ( D log 2 + sum_pix |x - recon| )
//   kl    =  mean_b -0.5 sum_l (1 + logvar - mu^2 - exp(logvar))
//   qbce  =  BCELoss(q, 1) = mean_b -max(log q, -100)
//   loss  =  recon_w recon + kl_w kl + query_w qbce
// HBM-bound: reads recon and x once (2*D floats per row) and writes grad_recon.
// One CTA per batch row (grid-stride); per-CTA partial sums go to a workspace and
// the last CTA to finish adds them in index order, so the result is deterministic.
#include "params.cuh"

namespace vael {


__device__ __forceinline__ float sgn(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

__global__ void __launch_bounds__(256) elbo_terms_kernel(const ElboParams P) {
  __shared__ float red[2][8];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float invB = 1.0f / (float)P.B;
  const float gr = P.recon_w * invB, gk = P.kl_w * invB, gq = P.query_w * invB;
  const bool vec = (P.D % 4 == 0) && aligned16(P.recon) && aligned16(P.x) &&
                   (P.grad_recon == nullptr || aligned16(P.grad_recon));
  double acc_recon = 0.0, acc_kl = 0.0, acc_q = 0.0;  // meaningful in thread 0 only

  for (long long row = blockIdx.x; row < P.B; row += gridDim.x) {
    const float* xr = P.x + row * P.D;
    const float* rr = P.recon + row * P.D;
    float* gr_out = P.grad_recon ? P.grad_recon + row * P.D : nullptr;
    float s_abs = 0.0f;
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(xr);
      const float4* r4 = reinterpret_cast<const float4*>(rr);
      float4* g4 = reinterpret_cast<float4*>(gr_out);
      for (long long i = tid; i < P.D / 4; i += blockDim.x) {
        const float4 a = __ldg(x4 + i), b = __ldg(r4 + i);
        const float d0 = b.x - a.x, d1 = b.y - a.y, d2 = b.z - a.z, d3 = b.w - a.w;
        s_abs += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
        if (g4) g4[i] = make_float4(gr * sgn(d0), gr * sgn(d1), gr * sgn(d2), gr * sgn(d3));
      }
    } else {
      for (long long i = tid; i < P.D; i += blockDim.x) {
        const float d = __ldg(rr + i) - __ldg(xr + i);
        s_abs += fabsf(d);
        if (gr_out) gr_out[i] = gr * sgn(d);
      }
    }
    float s_kl = 0.0f;
    for (int l = tid; l < P.L; l += blockDim.x) {
      const float m = __ldg(P.mu + row * P.L + l), lv = __ldg(P.logvar + row * P.L + l);
      const float ev = expf(lv);
      s_kl += 1.0f + lv - m * m - ev;
      if (P.grad_mu) P.grad_mu[row * P.L + l] = gk * m;
      if (P.grad_logvar) P.grad_logvar[row * P.L + l] = gk * 0.5f * (ev - 1.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_abs += __shfl_xor_sync(0xffffffffu, s_abs, o);
      s_kl += __shfl_xor_sync(0xffffffffu, s_kl, o);
    }
    if (lane == 0) { red[0][warp] = s_abs; red[1][warp] = s_kl; }
    __syncthreads();
    if (tid == 0) {
      float ta = 0.0f, tk = 0.0f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += red[0][w]; tk += red[1][w]; }
      acc_recon += (double)((float)P.D * 0.69314718055994531f + ta);
      acc_kl += (double)(-0.5f * tk);
      if (P.query_prob != nullptr) {
        const float q = __ldg(P.query_prob + row);
        acc_q += (double)(-fmaxf(logf(q), -100.0f));
        // torch binary_cross_entropy backward with target 1: (q - 1) / max((1 - q) q, 1e-12)
        if (P.grad_query) P.grad_query[row] = gq * (q - 1.0f) / fmaxf((1.0f - q) * q, 1e-12f);
      }
    }
    __syncthreads();
  }

  if (tid == 0) {
    P.partials[blockIdx.x * 3 + 0] = acc_recon;
    P.partials[blockIdx.x * 3 + 1] = acc_kl;
    P.partials[blockIdx.x * 3 + 2] = acc_q;
    __threadfence();
    const unsigned int done = atomicAdd(P.counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && tid == 0) {
    __threadfence();
    double r = 0.0, k = 0.0, q = 0.0;
    const volatile double* part = P.partials;
    for (unsigned int b = 0; b < gridDim.x; ++b) {
      r += part[b * 3 + 0];
      k += part[b * 3 + 1];
      q += part[b * 3 + 2];
    }
    const float recon = (float)(r / (double)P.B), kl = (float)(k / (double)P.B), qb = (float)(q / (double)P.B);
    P.terms[1] = recon;
    P.terms[2] = kl;
    P.terms[3] = qb;
    P.terms[0] = P.recon_w * recon + P.kl_w * kl + P.query_w * qb;
    *P.counter = 0u;  // self-cleaning for the next call
  }
}

constexpr int kElboMaxGrid = 2048;
size_t elbo_workspace_bytes() { return 16 + (size_t)kElboMaxGrid * 3 * sizeof(double); }

cudaError_t elbo_terms(ElboParams P, void* workspace, int num_sms, cudaStream_t st) {
  P.counter = reinterpret_cast<unsigned int*>(workspace);
  P.partials = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + 16);
  const int grid = (int)max(1LL, min(P.B, (long long)min(kElboMaxGrid, num_sms * 8)));
  elbo_terms_kernel<<<grid, 256, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace vael
