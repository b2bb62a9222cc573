// Literal-table circuits in the log semiring: world w = the product over ALL F fact columns of one literal
// per column, positive where table[w][f] is set and negative elsewhere — MarioVAELModel.admissible_worlds_logprob
// (models/vael.py:368-375: log p @ W_adm^T + log(1-p) @ (1-W_adm)^T).  The CSR the host compiler emits for it
// (vael_b200/compiler.py:circuit_from_world_table) lists, per world, the positive literals in column order and
// then the negative ones; vael_circuit_create recognises that shape and keeps one F-bit mask per world.
//
// Arithmetic (round 2).  Per row, once: lp[f] = log p_f, ln[f] = log(1 - p_f), d[f] = lp[f] - ln[f],
// base- = sum_f ln[f], base+ = sum_f lp[f].  A world with few positive literals (Mario: exactly two of 18) is
//     world_w = base- + sum_{f in pos(w)} d[f]            (popcount adds instead of F)
// and one with mostly positive literals base+ - sum_{f not in pos(w)} d[f].  The reference's two matmuls add the
// same F terms in an order nobody specifies; the parity bar is 1e-5 relative against its golden
// (tests/golden/mario.npz), not bit-equality with the generic CSR kernel (which folds in CSR order).
// The adjoint needs, per column, sp[f] = sum of g_w over the worlds where f is positive; the negative side is
// G - sp[f] with G = sum_w g_w, and grad_f = sp/p - (G - sp)/(1 - p) = (sp - G p) / (p (1 - p)): one division.
//
// Layout.  lane <-> row.  The [32,F] facts tile and the [32,W] world tile move between HBM and shared memory as
// contiguous spans (coalesced), at ODD row pitches in shared memory (F+1 = 19, W+1 = 25 words) so that lane <-> row
// reads and writes are conflict-free.  d[f] goes to a column-major [2F+1][32] array (rows F+1.. hold -d, row F
// zeros) so that "add the delta of column c", c a warp-uniform table entry, is one conflict-free LDS: the lists
// are padded to a common length with the zero row, every loop is branch-free and uniform.  The tables live in the
// kernel parameter block (constant bank).  The next tile's facts are requested one iteration ahead.
// Bound: 2 F accurate logf per row (fast __logf is not good enough: log(0.9999) would carry 0.4 % error) — about
// 850 instructions per row against 168 B: instruction- and HBM-bound meet near 4e10 rows/s.
#include "params.cuh"

namespace vael {

constexpr int kLtMaxWorlds = 64;
constexpr int kLtMaxList = 16;  // longest per-world delta list / per-column world list

struct LtTables {
  // forward: world w = (use_pos[w] ? base+ : base-) + sum_i D[ row[w][i] ]   (row index into the [2F+1][32] delta array)
  uint8_t row[kLtMaxWorlds][kLtMaxList];
  uint8_t use_pos[kLtMaxWorlds];
  int32_t list_len;  // common (padded) length of the per-world lists
  // backward: sp[f] = sum_i g[ world[f][i] ]   (index W = the zero pad column of the staged tile)
  uint8_t world[32][kLtMaxList];
  int32_t col_len;   // common (padded) length of the per-column lists
};

template <int F>
__device__ __forceinline__ void lt_read_row(const float* fs, int lane, float* p) {
#pragma unroll
  for (int c = 0; c < F; ++c) p[c] = fs[lane * (F + 1) + c];
}

template <int F>
__global__ void __launch_bounds__(256) lt_forward_kernel(const float* __restrict__ facts, long long B, int W,
                                                         const LtTables T, float* __restrict__ worlds) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int FP = F + 1, NR = 2 * F + 1;
  const int WP = (W + 1) | 1;   // odd pitch >= W + 1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per_warp = (kWarp * FP + NR * kWarp + kWarp * WP) * 4;
  float* fs = reinterpret_cast<float*>(smem + (size_t)warp * per_warp);   // [32][F+1]
  float* ds = fs + kWarp * FP;                                            // [2F+1][32]
  float* ws = ds + NR * kWarp;                                            // [32][W+1]
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
  ds[F * kWarp + lane] = 0.0f;   // the zero row
  const int r_first = lane / W, c_first = lane - r_first * W;
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, B - row0);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < F; ++k) {
      const int i = k * kWarp + lane, r = i / F, c = i - r * F;   // F compile-time: mul-shift
      fs[r * FP + c] = nx[k];
    }
    __syncwarp();
    fetch(sub + GW);
    float p[F];
    lt_read_row<F>(fs, lane, p);
    float base_neg = 0.0f, base_pos = 0.0f;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const float lp = logf(p[f]), ln = logf(1.0f - p[f]);
      base_pos += lp;
      base_neg += ln;
      const float d = lp - ln;
      ds[f * kWarp + lane] = d;
      ds[(F + 1 + f) * kWarp + lane] = -d;
    }
    __syncwarp();
    const int L = T.list_len;
#pragma unroll 4   // four worlds = four independent chains in flight per lane
    for (int w = 0; w < W; ++w) {
      float acc = T.use_pos[w] ? base_pos : base_neg;
      for (int i = 0; i < L; ++i) acc += ds[(int)T.row[w][i] * kWarp + lane];
      ws[lane * WP + w] = acc;
    }
    __syncwarp();
    float* dst = worlds + row0 * W;
    for (int i = lane, r = r_first, c = c_first; i < rows * W; i += kWarp) {
      dst[i] = ws[r * WP + c];
      for (c += kWarp; c >= W; c -= W) ++r;   // W is a run-time size: no division per element
    }
  }
}

template <int F>
__global__ void __launch_bounds__(256) lt_backward_kernel(const float* __restrict__ facts,
                                                          const float* __restrict__ grad_worlds, long long B, int W,
                                                          const LtTables T, float* __restrict__ grad_facts) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int FP = F + 1;
  const int WP = (W + 1) | 1;   // odd pitch >= W + 1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per_warp = (kWarp * FP + kWarp * WP) * 4;
  float* fs = reinterpret_cast<float*>(smem + (size_t)warp * per_warp);   // [32][F+1] facts in, grad_facts out
  float* gs = fs + kWarp * FP;                                            // [32][W+1], column W = 0
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  gs[lane * WP + W] = 0.0f;
  const int r_first = lane / W, c_first = lane - r_first * W;
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, B - row0);
    __syncwarp();
    for (int i = lane; i < kWarp * F; i += kWarp) {
      const int r = i / F;
      fs[i + r] = (i < rows * F) ? __ldg(facts + row0 * F + i) : 0.5f;
    }
    for (int i = lane, r = r_first, c = c_first; i < kWarp * W; i += kWarp) {
      gs[r * WP + c] = (grad_worlds != nullptr && i < rows * W) ? __ldg(grad_worlds + row0 * W + i) : 0.0f;
      for (c += kWarp; c >= W; c -= W) ++r;   // W is a run-time size: no division per element
    }
    __syncwarp();
    float p[F], sp[F];
    lt_read_row<F>(fs, lane, p);
    float G = 0.0f;
#pragma unroll 4
    for (int w = 0; w < W; ++w) G += gs[lane * WP + w];
    const int L = T.col_len;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      float acc = 0.0f;
      for (int i = 0; i < L; ++i) acc += gs[lane * WP + (int)T.world[f][i]];
      sp[f] = acc;
    }
    __syncwarp();
    // sp / p - (G - sp) / (1 - p) = (sp - G p) / (p (1 - p))
#pragma unroll
    for (int f = 0; f < F; ++f) fs[lane * FP + f] = fmaf(-G, p[f], sp[f]) / (p[f] * (1.0f - p[f]));
    __syncwarp();
    float* dst = grad_facts + row0 * F;
    for (int i = lane; i < rows * F; i += kWarp) {
      const int r = i / F;
      dst[i] = fs[i + r];
    }
  }
}

// The tables of a world-mask set; false when a list would not fit (the call then runs the generic kernel).
static bool lt_tables(int F, int W, const uint32_t* masks, LtTables* T) {
  memset(T, 0, sizeof(*T));
  int L = 0, Lc = 0;
  for (int w = 0; w < W; ++w) {
    const int pc = __builtin_popcount(masks[w]);
    L = max(L, min(pc, F - pc));
  }
  for (int f = 0; f < F; ++f) {
    int n = 0;
    for (int w = 0; w < W; ++w) n += (masks[w] >> f) & 1u;
    Lc = max(Lc, n);
  }
  if (L > kLtMaxList || Lc > kLtMaxList || F > 32 || W > kLtMaxWorlds || 2 * F + 1 > 255) return false;
  for (int w = 0; w < W; ++w) {
    const int pc = __builtin_popcount(masks[w]);
    const bool pos_base = pc > F - pc;   // mostly positive literals: start from base+ and subtract the others' deltas
    T->use_pos[w] = pos_base ? 1 : 0;
    int n = 0;
    for (int f = 0; f < F; ++f) {
      const bool set = (masks[w] >> f) & 1u;
      if (!pos_base && set) T->row[w][n++] = (uint8_t)f;
      if (pos_base && !set) T->row[w][n++] = (uint8_t)(F + 1 + f);
    }
    for (; n < L; ++n) T->row[w][n] = (uint8_t)F;   // the zero row
  }
  for (int f = 0; f < F; ++f) {
    int n = 0;
    for (int w = 0; w < W; ++w)
      if ((masks[w] >> f) & 1u) T->world[f][n++] = (uint8_t)w;
    for (; n < Lc; ++n) T->world[f][n] = (uint8_t)W;   // the zero column
  }
  T->list_len = L;
  T->col_len = Lc;
  return true;
}

bool lt_supported(int F, int W) { return F == 18 && W >= 1 && W <= kLtMaxWorlds; }

template <typename K, typename... Args>
static cudaError_t lt_launch(K kern, int per_warp, long long B, int num_sms, cudaStream_t st, Args... args) {
  const int nw = 8;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nw * per_warp);
  if (e != cudaSuccess) return e;
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long grid = max(1LL, min((n_sub + nw - 1) / nw, (long long)num_sms * 3));
  kern<<<(unsigned)grid, nw * 32, nw * per_warp, st>>>(args...);
  return cudaGetLastError();
}

cudaError_t lt_forward(int F, int W, const uint32_t* masks, const float* facts, long long B, float* worlds, int num_sms,
                       cudaStream_t st) {
  LtTables T;
  if (!lt_supported(F, W) || !lt_tables(F, W, masks, &T)) return cudaErrorInvalidValue;
  const int per_warp = (kWarp * (F + 1) + (2 * F + 1) * kWarp + kWarp * ((W + 1) | 1)) * 4;
  return lt_launch(lt_forward_kernel<18>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
}
cudaError_t lt_backward(int F, int W, const uint32_t* masks, const float* facts, const float* grad_worlds, long long B,
                        float* grad_facts, int num_sms, cudaStream_t st) {
  LtTables T;
  if (!lt_supported(F, W) || !lt_tables(F, W, masks, &T)) return cudaErrorInvalidValue;
  const int per_warp = (kWarp * (F + 1) + kWarp * ((W + 1) | 1)) * 4;
  return lt_launch(lt_backward_kernel<18>, per_warp, B, num_sms, st, facts, grad_worlds, B, W, T, grad_facts);
}

// does the mask set fit the table form?  (checked once at vael_circuit_create)
bool lt_masks_fit(int F, int W, const uint32_t* masks) {
  LtTables T;
  return lt_supported(F, W) && lt_tables(F, W, masks, &T);
}

}  // namespace vael
