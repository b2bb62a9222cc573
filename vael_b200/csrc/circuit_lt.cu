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
// Layout.  lane <-> row; warp-private tile pipelines (common.cuh: WarpTileLoader).  A [32,F] facts tile, a [32,W]
// upstream-gradient tile and the result tiles are contiguous spans of global memory: each moves by ONE 1-D bulk copy
// (cp.async.bulk, SASS UBLKCP) between HBM and a row-major shared-memory stage, two stages deep per warp, loads
// completing on mbarriers, stores draining while the next tile computes — no LSU wavefronts for global traffic
// (per-lane 64 / 128-bit global accesses at a 72 / 96-byte stride cost one L1 wavefront per 128-byte line touched:
// ~300 per tile, measured as the bound of the first round-2 version, profiles/r02b_fused_lt_ncu.md).  A lane reads
// its row from the stage with LDS.64 (72-byte row pitch: conflict-free per half-warp) and writes its results with
// STS.128 (96-byte pitch: two-way conflicts, 48 wavefronts per tile).  Shared memory also holds what needs a
// run-time index: forward, the per-lane deltas d[f] as a column-major [2F+1][32] array (rows F+1.. hold -d, row F
// zeros) so that "add the delta of column c", c a warp-uniform table entry, is one conflict-free LDS; backward,
// the upstream gradients as [W+1][32] (row W zeros).  The tables (per world: the rows to add; per column: the
// worlds that set it) are packed four byte-indices to a word, padded to a common length L with the zero row,
// copied to shared memory once per CTA and read as broadcasts; the loops are unrolled over L (template),
// branch-free, and independent across worlds.
// Bound: HBM (168 B forward, 240 B backward per row; ~180 shared-memory wavefronts and ~600 instructions per tile).
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

constexpr int kLtStages = 2;
__host__ __device__ constexpr int lt_pad128(int bytes) { return (bytes + 127) / 128 * 128; }
// per-warp shared memory: [barriers 128 B][facts stages][gw stages (backward)][column-major array][result stages]
__host__ __device__ constexpr int lt_fwd_warp_bytes(int F, int W) {
  return 128 + kLtStages * lt_pad128(kWarp * F * 4) + lt_pad128((2 * F + 1) * kWarp * 4) + kLtStages * lt_pad128(kWarp * W * 4);
}
__host__ __device__ constexpr int lt_bwd_warp_bytes(int F, int W) {
  return 128 + kLtStages * lt_pad128(kWarp * F * 4) + kLtStages * lt_pad128(kWarp * W * 4) + lt_pad128((W + 1) * kWarp * 4) +
         kLtStages * lt_pad128(kWarp * F * 4);
}

// ln x for x in [1e-5, 1): never denormal, so the bare MUFU.LG2 (no range fix-up) and one multiply
__device__ __forceinline__ float lt_log(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.69314718055994531f;
}

// this lane's row of the [32, K] stage (K even: 8-byte aligned rows)
template <int K>
__device__ __forceinline__ void lt_row_from_stage(const float* st, int lane, float* x) {
#pragma unroll
  for (int c = 0; c < K / 2; ++c) {
    const float2 t = *reinterpret_cast<const float2*>(st + lane * K + 2 * c);
    x[2 * c] = t.x; x[2 * c + 1] = t.y;
  }
}

// L = entries per list (padded with the zero row); list words hold four byte-indices each
// ANYPOS: some world starts from base+ (more positive than negative literals); Mario's admissible worlds never do
template <int F, int L, bool ANYPOS>
__global__ void __launch_bounds__(128) lt_forward_kernel(const float* __restrict__ facts, long long B, int W,
                                                         const LtTables T, float* __restrict__ worlds) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NR = 2 * F + 1, FT = kWarp * F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const LtTables* Ts = reinterpret_cast<const LtTables*>(smem);
  lt_stage_tables(reinterpret_cast<uint32_t*>(smem), T);
  uint8_t* wb = smem + lt_pad128(sizeof(LtTables)) + (size_t)warp * lt_fwd_warp_bytes(F, W);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* fstage = reinterpret_cast<float*>(wb + 128);
  float* ds = reinterpret_cast<float*>(wb + 128 + kLtStages * lt_pad128(FT * 4));                  // [2F+1][32]
  float* ostage = ds + lt_pad128(NR * kWarp * 4) / 4;                                               // kLtStages x [32][W]
  const int ostride = lt_pad128(kWarp * W * 4) / 4;
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool in_al = aligned16(facts), out_al = aligned16(worlds) && ((kWarp * W * 4) % 16 == 0);
  const bool out_vec = (W % 4 == 0);
  WarpTileLoader<kLtStages> ld;
  ld.init(fstage, bars, lt_pad128(FT * 4) / 4, lane);
  auto request = [&](long long sub, int s) {
    if (sub < n_sub) ld.issue(s, facts + sub * FT, (int)min((long long)kWarp, B - sub * kWarp) * F, in_al, lane, 0.5f);
  };
#pragma unroll
  for (int s = 0; s < kLtStages; ++s) request(gw + s * GW, s);
  ds[F * kWarp + lane] = 0.0f;   // the zero row
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % kLtStages);
    const int rows = (int)min((long long)kWarp, B - sub * kWarp);
    ld.wait(s);
    float p[F];
    lt_row_from_stage<F>(ld.stage(s), lane, p);
    __syncwarp();
    request(sub + kLtStages * GW, s);   // refill this stage behind the tile's arithmetic
    float base_neg = 0.0f, base_pos = 0.0f;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const float lp = lt_log(p[f]), ln = lt_log(1.0f - p[f]);
      base_pos += lp;
      base_neg += ln;
      const float d = lp - ln;
      ds[f * kWarp + lane] = d;
      ds[(F + 1 + f) * kWarp + lane] = -d;
    }
    if (lane == 0) bulk_wait_read<kLtStages - 1>();   // the store issued kLtStages tiles ago has read its stage
    __syncwarp();
    float* os = ostage + s * ostride + lane * W;
    auto world_value = [&](int w) {
      float acc = (ANYPOS && ((Ts->use_pos[w >> 5] >> (w & 31)) & 1u)) ? base_pos : base_neg;
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const uint32_t word = Ts->row[w][i >> 2];   // shared-memory broadcast
        acc += ds[((word >> (8 * (i & 3))) & 0xffu) * kWarp + lane];
      }
      return acc;
    };
    if (out_vec) {
      for (int w = 0; w < W; w += 4)
        *reinterpret_cast<float4*>(os + w) = make_float4(world_value(w), world_value(w + 1), world_value(w + 2), world_value(w + 3));
    } else {
      for (int w = 0; w < W; ++w) os[w] = world_value(w);
    }
    fence_proxy_async();
    __syncwarp();
    warp_tile_store(worlds + sub * kWarp * W, ostage + s * ostride, rows * W, kWarp * W, out_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
}

template <int F, int L>
__global__ void __launch_bounds__(128) lt_backward_kernel(const float* __restrict__ facts,
                                                          const float* __restrict__ grad_worlds, long long B, int W,
                                                          const LtTables T, float* __restrict__ grad_facts) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int FT = kWarp * F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const LtTables* Ts = reinterpret_cast<const LtTables*>(smem);
  lt_stage_tables(reinterpret_cast<uint32_t*>(smem), T);
  uint8_t* wb = smem + lt_pad128(sizeof(LtTables)) + (size_t)warp * lt_bwd_warp_bytes(F, W);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* fstage = reinterpret_cast<float*>(wb + 128);
  float* gstage = fstage + kLtStages * lt_pad128(FT * 4) / 4;
  const int gstride = lt_pad128(kWarp * W * 4) / 4;
  float* gs = gstage + kLtStages * gstride;                                  // [W+1][32]
  float* ostage = gs + lt_pad128((W + 1) * kWarp * 4) / 4;                   // kLtStages x [32][F]
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool f_al = aligned16(facts), g_al = grad_worlds != nullptr && aligned16(grad_worlds);
  const bool out_al = aligned16(grad_facts);
  const bool g_vec = (W % 4 == 0);
  WarpTileLoader<kLtStages> lf, lg;
  lf.init(fstage, bars, lt_pad128(FT * 4) / 4, lane);
  lg.init(gstage, bars + kLtStages, gstride, lane);
  auto request = [&](long long sub, int s) {
    if (sub >= n_sub) return;
    const int rows = (int)min((long long)kWarp, B - sub * kWarp);
    lf.issue(s, facts + sub * FT, rows * F, f_al, lane, 0.5f);
    if (grad_worlds != nullptr) lg.issue(s, grad_worlds + sub * kWarp * W, rows * W, g_al, lane, 0.0f);
  };
#pragma unroll
  for (int s = 0; s < kLtStages; ++s) request(gw + s * GW, s);
  gs[W * kWarp + lane] = 0.0f;   // the zero row
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % kLtStages);
    const int rows = (int)min((long long)kWarp, B - sub * kWarp);
    lf.wait(s);
    if (grad_worlds != nullptr) lg.wait(s);
    float p[F];
    lt_row_from_stage<F>(lf.stage(s), lane, p);
    // this lane's upstream gradients -> the column-major array (run-time world index below), and their sum
    float G0 = 0.0f, G1 = 0.0f;
    const float* grow = lg.stage(s) + lane * W;
    if (grad_worlds == nullptr) {
      for (int w = 0; w < W; ++w) gs[w * kWarp + lane] = 0.0f;
    } else if (g_vec) {
      for (int w = 0; w < W; w += 4) {
        const float4 t = *reinterpret_cast<const float4*>(grow + w);
        gs[w * kWarp + lane] = t.x; gs[(w + 1) * kWarp + lane] = t.y;
        gs[(w + 2) * kWarp + lane] = t.z; gs[(w + 3) * kWarp + lane] = t.w;
        G0 += t.x + t.z; G1 += t.y + t.w;
      }
    } else {
      for (int w = 0; w < W; ++w) {
        const float t = grow[w];
        gs[w * kWarp + lane] = t;
        G0 += t;
      }
    }
    const float G = G0 + G1;
    if (lane == 0) bulk_wait_read<kLtStages - 1>();
    __syncwarp();
    request(sub + kLtStages * GW, s);   // both input stages are consumed
    float* os = ostage + s * (lt_pad128(FT * 4) / 4) + lane * F;
#pragma unroll
    for (int f = 0; f < F; f += 2) {
      float o[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float sp = 0.0f;
#pragma unroll
        for (int i = 0; i < L; ++i) {
          const uint32_t word = Ts->world[f + h][i >> 2];
          sp += gs[((word >> (8 * (i & 3))) & 0xffu) * kWarp + lane];
        }
        // sp / p - (G - sp) / (1 - p) = (sp - G p) / (p (1 - p))
        o[h] = __fdividef(fmaf(-G, p[f + h], sp), p[f + h] * (1.0f - p[f + h]));
      }
      *reinterpret_cast<float2*>(os + f) = make_float2(o[0], o[1]);
    }
    fence_proxy_async();
    __syncwarp();
    warp_tile_store(grad_facts + sub * FT, ostage + s * (lt_pad128(FT * 4) / 4), rows * F, FT, out_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
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
  const int nw = 4;
  const int smem = lt_pad128((int)sizeof(LtTables)) + nw * per_warp;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const int ctas_per_sm = max(1, min(8, (227 * 1024) / (smem + 1024)));
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long grid = tile_grid(n_sub, nw, (long long)num_sms * ctas_per_sm);
  kern<<<(unsigned)grid, nw * 32, smem, st>>>(args...);
  return cudaGetLastError();
}

// list lengths the kernels are instantiated for (a list is padded up to the next one)
static int lt_round_len(int L) { return L <= 2 ? 2 : (L <= 4 ? 4 : (L <= 8 ? 8 : 16)); }

cudaError_t lt_forward(int F, int W, const uint32_t* masks, const float* facts, long long B, float* worlds, int num_sms,
                       cudaStream_t st) {
  LtTables T;
  if (!lt_supported(F, W) || !lt_tables(F, W, masks, &T)) return cudaErrorInvalidValue;
  const int per_warp = lt_fwd_warp_bytes(F, W);
  const bool anypos = (T.use_pos[0] | T.use_pos[1]) != 0u;
  if (!anypos && T.list_len == 2) return lt_launch(lt_forward_kernel<18, 2, false>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
  switch (T.list_len) {
    case 2: return lt_launch(lt_forward_kernel<18, 2, true>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
    case 4: return lt_launch(lt_forward_kernel<18, 4, true>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
    case 8: return lt_launch(lt_forward_kernel<18, 8, true>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
    default: return lt_launch(lt_forward_kernel<18, 16, true>, per_warp, B, num_sms, st, facts, B, W, T, worlds);
  }
}
cudaError_t lt_backward(int F, int W, const uint32_t* masks, const float* facts, const float* grad_worlds, long long B,
                        float* grad_facts, int num_sms, cudaStream_t st) {
  LtTables T;
  if (!lt_supported(F, W) || !lt_tables(F, W, masks, &T)) return cudaErrorInvalidValue;
  const int per_warp = lt_bwd_warp_bytes(F, W);
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
