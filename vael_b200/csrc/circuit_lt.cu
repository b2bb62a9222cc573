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
// same F terms in an order nobody specifies; the parity bar is 1e-5 against its golden (tests/golden/mario.npz),
// not bit-equality with the generic CSR kernel (which folds in CSR order).  Accuracy: base + delta cancels when a
// row is peaked (base- = -18.4, two deltas of +9.2, result -3.6e-4), so the result carries an ABSOLUTE error of a few
// ulp of the base (~2e-6) — i.e. the world PROBABILITY exp(result) carries that relative error, which is what the
// 1e-5 bar is about; tests hold log-probabilities to 1e-5 * max(1, |ref|).  With that error model the logarithms
// are the fast ones (MUFU.LG2: absolute error 2^-21.4 on [0.5, 2], 3 ulp elsewhere) — 2 instructions instead of 25.
// The adjoint needs, per column, sp[f] = sum of g_w over the worlds where f is positive; the negative side is
// G - sp[f] with G = sum_w g_w, and grad_f = sp/p - (G - sp)/(1 - p) = (sp - G p) / (p (1 - p)): one division.
//
// Layout.  lane <-> row, no staging tiles: a lane reads its own row (F = 18 floats = 72 contiguous bytes, rows
// back to back: 9 x LDG.64 per lane, the warp covers one contiguous 2304-byte span and every sector is consumed
// while it is L1-resident) one tile ahead into registers, and writes its own W results with 128-bit stores (the
// partially written sectors of one instruction are completed by the next ones of the same warp before L2 evicts
// them: DRAM traffic stays at the algorithmic bytes, checked with ncu).  Shared memory holds only what needs a
// run-time index: forward, the per-lane deltas d[f] as a column-major [2F+1][32] array (rows F+1.. hold -d, row F
// zeros) so that "add the delta of column c", c a warp-uniform table entry, is one conflict-free LDS; backward,
// the upstream gradients as [W+1][32] (row W zeros).  The tables (per world: the rows to add; per column: the
// worlds that set it) are packed four byte-indices to a word, padded to a common length L with the zero row,
// copied to shared memory once per CTA and read as broadcasts; the loops are unrolled over L (template),
// branch-free, and independent across worlds.
// Bound: HBM (168 B forward, 240 B backward per row against ~600 / ~550 instructions per 32-row tile).
#include "params.cuh"

namespace vael {

constexpr int kLtMaxWorlds = 64;
constexpr int kLtMaxList = 16;  // longest per-world delta list / per-column world list

struct LtTables {
  // forward: world w = (use_pos[w] ? base+ : base-) + sum_i D[ row[w][i] ]   (row index into the [2F+1][32] delta array)
  uint32_t row[kLtMaxWorlds][kLtMaxList / 4];   // four byte-indices per word
  uint32_t use_pos[2];                          // bit w
  int32_t list_len;  // common (padded) length of the per-world lists
  // backward: sp[f] = sum_i g[ world[f][i] ]   (index W = the zero pad column of the staged tile)
  uint32_t world[32][kLtMaxList / 4];
  int32_t col_len;   // common (padded) length of the per-column lists
};

__device__ __forceinline__ void lt_stage_tables(uint32_t* dst, const LtTables& T) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&T);
  for (int i = threadIdx.x; i < (int)(sizeof(LtTables) / 4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// one row of K floats (K even, rows 8-byte aligned) <-> registers
template <int K>
__device__ __forceinline__ void lt_load_row(const float* __restrict__ base, long long row, bool valid, float fill, float* x) {
  if (valid) {
    const float2* src = reinterpret_cast<const float2*>(base + row * K);
#pragma unroll
    for (int c = 0; c < K / 2; ++c) { const float2 t = __ldg(src + c); x[2 * c] = t.x; x[2 * c + 1] = t.y; }
  } else {
#pragma unroll
    for (int c = 0; c < K; ++c) x[c] = fill;
  }
}

// L = entries per list (padded with the zero row); list words hold four byte-indices each
template <int F, int L>
__global__ void __launch_bounds__(256) lt_forward_kernel(const float* __restrict__ facts, long long B, int W,
                                                         const LtTables T, float* __restrict__ worlds) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NR = 2 * F + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const LtTables* Ts = reinterpret_cast<const LtTables*>(smem);
  lt_stage_tables(reinterpret_cast<uint32_t*>(smem), T);
  float* ds = reinterpret_cast<float*>(smem + ((sizeof(LtTables) + 127) / 128) * 128) + (size_t)warp * NR * kWarp;   // [2F+1][32]
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  const bool out_vec = (W % 4 == 0) && aligned16(worlds);
  float nx[F];
  lt_load_row<F>(facts, gw * kWarp + lane, gw * kWarp + lane < B, 0.5f, nx);
  ds[F * kWarp + lane] = 0.0f;   // the zero row
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const long long row = sub * kWarp + lane;
    float p[F];
#pragma unroll
    for (int f = 0; f < F; ++f) p[f] = nx[f];
    lt_load_row<F>(facts, row + GW * kWarp, row + GW * kWarp < B, 0.5f, nx);   // the next tile, one iteration ahead
    float base_neg = 0.0f, base_pos = 0.0f;
    __syncwarp();
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const float lp = __logf(p[f]), ln = __logf(1.0f - p[f]);
      base_pos += lp;
      base_neg += ln;
      const float d = lp - ln;
      ds[f * kWarp + lane] = d;
      ds[(F + 1 + f) * kWarp + lane] = -d;
    }
    __syncwarp();
    float* dst = worlds + row * W;
    auto world_value = [&](int w) {
      float acc = ((Ts->use_pos[w >> 5] >> (w & 31)) & 1u) ? base_pos : base_neg;
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const uint32_t word = Ts->row[w][i >> 2];   // shared-memory broadcast
        acc += ds[((word >> (8 * (i & 3))) & 0xffu) * kWarp + lane];
      }
      return acc;
    };
    if (row < B) {
      if (out_vec) {
        for (int w = 0; w < W; w += 4)
          *reinterpret_cast<float4*>(dst + w) = make_float4(world_value(w), world_value(w + 1), world_value(w + 2), world_value(w + 3));
      } else {
        for (int w = 0; w < W; ++w) dst[w] = world_value(w);
      }
    }
  }
}

template <int F, int L>
__global__ void __launch_bounds__(256) lt_backward_kernel(const float* __restrict__ facts,
                                                          const float* __restrict__ grad_worlds, long long B, int W,
                                                          const LtTables T, float* __restrict__ grad_facts) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const LtTables* Ts = reinterpret_cast<const LtTables*>(smem);
  lt_stage_tables(reinterpret_cast<uint32_t*>(smem), T);
  float* gs = reinterpret_cast<float*>(smem + ((sizeof(LtTables) + 127) / 128) * 128) + (size_t)warp * (W + 1) * kWarp;   // [W+1][32]
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  const bool g_vec = grad_worlds != nullptr && (W % 4 == 0) && aligned16(grad_worlds);
  gs[W * kWarp + lane] = 0.0f;   // the zero row
  float nx[F];
  lt_load_row<F>(facts, gw * kWarp + lane, gw * kWarp + lane < B, 0.5f, nx);
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const long long row = sub * kWarp + lane;
    const bool valid = row < B;
    float p[F];
#pragma unroll
    for (int f = 0; f < F; ++f) p[f] = nx[f];
    lt_load_row<F>(facts, row + GW * kWarp, row + GW * kWarp < B, 0.5f, nx);
    // this lane's upstream gradients -> the column-major array (run-time world index below), and their sum
    __syncwarp();
    float G0 = 0.0f, G1 = 0.0f;
    if (g_vec) {
      const float4* src = reinterpret_cast<const float4*>(grad_worlds + row * W);
      for (int w = 0; w < W; w += 4) {
        const float4 t = valid ? __ldg(src + (w >> 2)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        gs[w * kWarp + lane] = t.x; gs[(w + 1) * kWarp + lane] = t.y;
        gs[(w + 2) * kWarp + lane] = t.z; gs[(w + 3) * kWarp + lane] = t.w;
        G0 += t.x + t.z; G1 += t.y + t.w;
      }
    } else {
      for (int w = 0; w < W; ++w) {
        const float t = (valid && grad_worlds != nullptr) ? __ldg(grad_worlds + row * W + w) : 0.0f;
        gs[w * kWarp + lane] = t;
        G0 += t;
      }
    }
    const float G = G0 + G1;
    __syncwarp();
    float out[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      float sp = 0.0f;
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const uint32_t word = Ts->world[f][i >> 2];
        sp += gs[((word >> (8 * (i & 3))) & 0xffu) * kWarp + lane];
      }
      // sp / p - (G - sp) / (1 - p) = (sp - G p) / (p (1 - p))
      out[f] = __fdividef(fmaf(-G, p[f], sp), p[f] * (1.0f - p[f]));
    }
    if (valid) {
      float2* dst = reinterpret_cast<float2*>(grad_facts + row * F);
#pragma unroll
      for (int c = 0; c < F / 2; ++c) dst[c] = make_float2(out[2 * c], out[2 * c + 1]);
    }
  }
}

static int lt_round_len(int L);
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
  L = lt_round_len(L);     // the instantiated lengths; lists are padded with the zero row
  Lc = lt_round_len(Lc);
  auto put = [](uint32_t* words, int i, int v) { words[i >> 2] |= (uint32_t)v << (8 * (i & 3)); };
  for (int w = 0; w < W; ++w) {
    const int pc = __builtin_popcount(masks[w]);
    const bool pos_base = pc > F - pc;   // mostly positive literals: start from base+ and subtract the others' deltas
    if (pos_base) T->use_pos[w >> 5] |= 1u << (w & 31);
    int n = 0;
    for (int f = 0; f < F; ++f) {
      const bool set = (masks[w] >> f) & 1u;
      if (!pos_base && set) put(T->row[w], n++, f);
      if (pos_base && !set) put(T->row[w], n++, F + 1 + f);
    }
    for (; n < L; ++n) put(T->row[w], n, F);   // the zero row
  }
  for (int f = 0; f < F; ++f) {
    int n = 0;
    for (int w = 0; w < W; ++w)
      if ((masks[w] >> f) & 1u) put(T->world[f], n++, w);
    for (; n < Lc; ++n) put(T->world[f], n, W);   // the zero column
  }
  T->list_len = L;
  T->col_len = Lc;
  return true;
}

bool lt_supported(int F, int W) { return F == 18 && W >= 1 && W <= kLtMaxWorlds; }

template <typename K, typename... Args>
static cudaError_t lt_launch(K kern, int per_warp, long long B, int num_sms, cudaStream_t st, Args... args) {
  const int nw = 8;
  const int smem = (int)((sizeof(LtTables) + 127) / 128) * 128 + nw * per_warp;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long grid = max(1LL, min((n_sub + nw - 1) / nw, (long long)num_sms * 4));
  kern<<<(unsigned)grid, nw * 32, smem, st>>>(args...);
  return cudaGetLastError();
}

// list lengths the kernels are instantiated for (a list is padded up to the next one)
static int lt_round_len(int L) { return L <= 2 ? 2 : (L <= 4 ? 4 : (L <= 8 ? 8 : 16)); }

cudaError_t lt_forward(int F, int W, const uint32_t* masks, const float* facts, long long B, float* worlds, int num_sms,
                       cudaStream_t st) {
  LtTables T;
  if (!lt_supported(F, W) || !lt_tables(F, W, masks, &T)) return cudaErrorInvalidValue;
  const int per_warp = (2 * F + 1) * kWarp * 4;
  switch (T.list_len) {
    case 2: return lt_launch(lt_forward_kernel<18, 2>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
    case 4: return lt_launch(lt_forward_kernel<18, 4>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
    case 8: return lt_launch(lt_forward_kernel<18, 8>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
    default: return lt_launch(lt_forward_kernel<18, 16>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
  }
}
cudaError_t lt_backward(int F, int W, const uint32_t* masks, const float* facts, const float* grad_worlds, long long B,
                        float* grad_facts, int num_sms, cudaStream_t st) {
  LtTables T;
  if (!lt_supported(F, W) || !lt_tables(F, W, masks, &T)) return cudaErrorInvalidValue;
  const int per_warp = (W + 1) * kWarp * 4;
  switch (T.col_len) {
    case 2: return lt_launch(lt_backward_kernel<18, 2>, per_warp, B, num_sms, st, facts, grad_worlds, B, W, T, grad_facts);
    case 4: return lt_launch(lt_backward_kernel<18, 4>, per_warp, B, num_sms, st, facts, grad_worlds, B, W, T, grad_facts);
    case 8: return lt_launch(lt_backward_kernel<18, 8>, per_warp, B, num_sms, st, facts, grad_worlds, B, W, T, grad_facts);
    default: return lt_launch(lt_backward_kernel<18, 16>, per_warp, B, num_sms, st, facts, grad_worlds, B, W, T, grad_facts);
  }
}

// does the mask set fit the table form?  (checked once at vael_circuit_create)
bool lt_masks_fit(int F, int W, const uint32_t* masks) {
  LtTables T;
  return lt_supported(F, W) && lt_tables(F, W, masks, &T);
}

}  // namespace vael
