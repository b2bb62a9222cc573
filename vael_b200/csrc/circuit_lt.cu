// Literal-table circuits in the log semiring: world w = the product over ALL F fact columns of one literal
// per column, positive where table[w][f] is set and negative elsewhere — MarioVAELModel.admissible_worlds_logprob
// (models/vael.py:368-375: log p @ W_adm^T + log(1-p) @ (1-W_adm)^T).  The CSR the host compiler emits for it
// (vael_b200/compiler.py:circuit_from_world_table) lists, per world, the positive literals in column order and
// then the negative ones; vael_circuit_create recognises that shape and keeps one F-bit mask per world.
//
// lane <-> row, everything in registers: lp[f] = log p_f and ln[f] = log(1 - p_f) are computed once per row
// (the same logf of the same operands as the generic kernel's leaf ops), every world is the left fold the CSR
// holds — the set columns ascending, then the clear columns ascending — as 2 F warp-uniformly predicated FADDs
// over register operands.  Bit-exact against the generic CSR kernel.  The [32, F] facts tile and the [32, W]
// result tile go through shared memory so that global accesses are coalesced; the next tile's facts are
// requested one iteration ahead.  Instruction-bound (2 F logf + 2 F W predicated adds per row against
// (F + W) * 4 bytes), not HBM-bound.
#include "params.cuh"

namespace vael {

constexpr int kLtMaxWorlds = 64;

struct LtMasks {
  uint32_t m[kLtMaxWorlds];
};

template <int F>
__device__ __forceinline__ void lt_read_row(const float* fs, int lane, float* p) {
  if constexpr (F % 2 == 0) {
#pragma unroll
    for (int c = 0; c < F; c += 2) {
      const float2 t = *reinterpret_cast<const float2*>(fs + lane * F + c);
      p[c] = t.x; p[c + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < F; ++c) p[c] = fs[lane * F + c];
  }
}

template <int F>
__global__ void __launch_bounds__(256) lt_forward_kernel(const float* __restrict__ facts, long long B, int W,
                                                         const LtMasks M, float* __restrict__ worlds) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per_warp = ((kWarp * F * 4 + 127) / 128) * 128 + ((kWarp * W * 4 + 127) / 128) * 128;
  float* fs = reinterpret_cast<float*>(smem + (size_t)warp * per_warp);                                          // [32][F]
  float* ws = reinterpret_cast<float*>(smem + (size_t)warp * per_warp + ((kWarp * F * 4 + 127) / 128) * 128);     // [32][W]
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  float nx[F];
  auto fetch = [&](long long sub) {
    const long long r0 = sub * kWarp;
    const int n = (sub < n_sub) ? (int)min((long long)kWarp, B - r0) * F : 0;
    const float* src = facts + r0 * F;
#pragma unroll
    for (int k = 0; k < F; ++k) nx[k] = (k * kWarp + lane < n) ? __ldg(src + k * kWarp + lane) : 0.5f;
  };
  fetch(gw);
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, B - row0);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < F; ++k) fs[k * kWarp + lane] = nx[k];
    __syncwarp();
    fetch(sub + GW);
    float p[F], lp[F], ln[F];
    lt_read_row<F>(fs, lane, p);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      lp[f] = logf(p[f]);
      ln[f] = logf(__fsub_rn(1.0f, p[f]));
    }
#pragma unroll 2   // two worlds = two independent FADD chains in flight per lane
    for (int w = 0; w < W; ++w) {
      const uint32_t m = M.m[w];
      float acc = 0.0f;   // 0 + x is exact: the fold starts at its first literal
#pragma unroll
      for (int f = 0; f < F; ++f)
        if ((m >> f) & 1u) acc = __fadd_rn(acc, lp[f]);
#pragma unroll
      for (int f = 0; f < F; ++f)
        if (!((m >> f) & 1u)) acc = __fadd_rn(acc, ln[f]);
      ws[lane * W + w] = acc;
    }
    __syncwarp();
    float* dst = worlds + row0 * W;
    for (int i = lane; i < rows * W; i += kWarp) dst[i] = ws[i];
  }
}

// adjoint: d world_w / d p_f = 1 / p_f for a positive literal, -1 / (1 - p_f) for a negative one, so
// grad_facts[f] = (sum over worlds with f set of g_w) / p_f - (sum over the others of g_w) / (1 - p_f),
// the two sums accumulated in world order (the order of the generic kernel's parent records).
template <int F>
__global__ void __launch_bounds__(256) lt_backward_kernel(const float* __restrict__ facts,
                                                          const float* __restrict__ grad_worlds, long long B, int W,
                                                          const LtMasks M, float* __restrict__ grad_facts) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int fbytes = ((kWarp * F * 4 + 127) / 128) * 128;
  const int per_warp = fbytes + ((kWarp * W * 4 + 127) / 128) * 128;
  float* fs = reinterpret_cast<float*>(smem + (size_t)warp * per_warp);            // [32][F] facts in, grad_facts out
  float* gs = reinterpret_cast<float*>(smem + (size_t)warp * per_warp + fbytes);   // [32][W]
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, B - row0);
    __syncwarp();
    for (int i = lane; i < kWarp * F; i += kWarp) fs[i] = (i < rows * F) ? __ldg(facts + row0 * F + i) : 0.5f;
    for (int i = lane; i < kWarp * W; i += kWarp)
      gs[i] = (grad_worlds != nullptr && i < rows * W) ? __ldg(grad_worlds + row0 * W + i) : 0.0f;
    __syncwarp();
    float p[F], sp[F], sn[F];
    lt_read_row<F>(fs, lane, p);
#pragma unroll
    for (int f = 0; f < F; ++f) sp[f] = sn[f] = 0.0f;
    for (int w = 0; w < W; ++w) {
      const uint32_t m = M.m[w];
      const float g = gs[lane * W + w];
#pragma unroll
      for (int f = 0; f < F; ++f) {
        if ((m >> f) & 1u) sp[f] += g;
        else sn[f] += g;
      }
    }
    __syncwarp();
#pragma unroll
    for (int f = 0; f < F; ++f) fs[lane * F + f] = sp[f] / p[f] + (-sn[f] / (1.0f - p[f]));
    __syncwarp();
    float* dst = grad_facts + row0 * F;
    for (int i = lane; i < rows * F; i += kWarp) dst[i] = fs[i];
  }
}

bool lt_supported(int F, int W) { return F == 18 && W >= 1 && W <= kLtMaxWorlds; }

template <typename K, typename... Args>
static cudaError_t lt_launch(K kern, int F, int W, long long B, int num_sms, cudaStream_t st, Args... args) {
  const int per_warp = ((kWarp * F * 4 + 127) / 128) * 128 + ((kWarp * W * 4 + 127) / 128) * 128;
  const int nw = 8;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nw * per_warp);
  if (e != cudaSuccess) return e;
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long grid = max(1LL, min((n_sub + nw - 1) / nw, (long long)num_sms * 4));
  kern<<<(unsigned)grid, nw * 32, nw * per_warp, st>>>(args...);
  return cudaGetLastError();
}

cudaError_t lt_forward(int F, int W, const uint32_t* masks, const float* facts, long long B, float* worlds, int num_sms,
                       cudaStream_t st) {
  if (!lt_supported(F, W)) return cudaErrorInvalidValue;
  LtMasks M{};
  for (int w = 0; w < W; ++w) M.m[w] = masks[w];
  return lt_launch(lt_forward_kernel<18>, F, W, B, num_sms, st, facts, B, W, M, worlds);
}
cudaError_t lt_backward(int F, int W, const uint32_t* masks, const float* facts, const float* grad_worlds, long long B,
                        float* grad_facts, int num_sms, cudaStream_t st) {
  if (!lt_supported(F, W)) return cudaErrorInvalidValue;
  LtMasks M{};
  for (int w = 0; w < W; ++w) M.m[w] = masks[w];
  return lt_launch(lt_backward_kernel<18>, F, W, B, num_sms, st, facts, grad_worlds, B, W, M, grad_facts);
}

}  // namespace vael
