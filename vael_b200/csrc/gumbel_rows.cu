// Gumbel-softmax world sampling + Herbrand projection, lane <-> row   (round 2)
//
// Reference: logits = log(worlds_prob); world = F.gumbel_softmax(logits, tau, hard=True);
// world_h = world @ herbrand_base      (models/vael.py:61-65, :82-84; Mario :335-338) and its autograd.
//
// The one-warp-per-row kernels in sampling.cu pay two five-step shuffle reductions, the arg-max shuffles and ~28
// idle lane slots per row of W = 100 worlds: 620 warp instructions per row, 0.21 / 0.25 of HBM
// (profiles/r02a_stage_ncu.md).  Here one LANE owns a row.  A [32, W] tile of scores arrives in a warp-private
// shared-memory buffer by ONE bulk copy (WarpTileLoader, SASS UBLKCP); the lane walks its row with 128-bit shared
// loads and rewrites it IN PLACE — scores -> perturbed logits z -> exp2(z - max) -> y_soft — so max / arg-max / sum
// are plain register chains, nothing W-sized lives in registers (the first version kept z[W] there: 190-255
// registers, 5 warps per SM, one instruction every 5 cycles), and the finished tile leaves by one bulk store.
// A tile buffer is single: ~12 resident warps per SM hide the load latency of one another.  These are
// THROUGHPUT kernels: the dispatcher takes them from 16 Ki rows up (below, a lane's ~4.6 k instructions are a latency
// floor the one-warp-per-row kernels do not have).
// Shapes: W a multiple of 4 whose tile fits (W <= 256), H <= 32; anything else stays on the one-warp-per-row
// kernels.  The Herbrand table is ANY [W, H] float table (0/1 in VAEL): the forward copies row `arg` of a
// shared-memory copy; the adjoint needs G_e = sum_h gh[h] H[e,h] and walks the non-zero columns of each world —
// a (column, weight) list per world built once per CTA, padded to 2 or 4 entries (VAEL's tables have two ones per
// row) so that the walk is branch-free, or a bit-mask loop for denser tables; ascending h either way, like the
// one-warp-per-row kernel.
//
// Arithmetic: the softmax runs in the log2 domain (z = log2(e)/tau (log p + g), exp2 of the difference to the
// maximum); log p through lg2.approx (absolute error <= 7e-6 on the logits VAEL produces, y_soft within 1e-5
// relative of the reference's); arg-max ties -> lowest world.  In-kernel noise: Philox4x32-10, one call per four
// consecutive worlds, counter = (row, world / 4, 1), key = seed.
#include "params.cuh"
#include "sampling.cuh"

namespace vael {

namespace {

__host__ __device__ constexpr int round128(int x) { return (x + 127) / 128 * 128; }

constexpr int kMaxPadEntries = 4;   // (column, weight) lists are padded to 2 or 4 entries; denser tables: bit masks

struct GrLayout {   // shared memory of one CTA, in bytes
  int tab, masks, entries, shared, per_warp, w_tile, h_tile;
};
__host__ __device__ inline GrLayout gr_layout(int W, int H, bool bwd, bool noise, bool need_scores) {
  GrLayout L;
  L.w_tile = round128(kWarp * W * 4);
  L.h_tile = round128(kWarp * H * 4);
  L.tab = round128(W * H * 4);
  L.masks = bwd ? round128(W * 4) : 0;
  L.entries = bwd ? round128(W * kMaxPadEntries * 8) : 0;
  L.shared = L.tab + L.masks + L.entries;
  if (!bwd) {
    // [bars 128][scores -> z -> y_soft][noise][world_h out]
    L.per_warp = 128 + L.w_tile * (noise ? 2 : 1) + L.h_tile;
  } else {
    // [bars 128][y_soft -> G -> grad_scores][scores -> y / p (is_log: y, no load)][grad_world_h][its transpose + one
    // zero row]
    (void)need_scores;
    L.per_warp = 128 + 2 * L.w_tile + 2 * L.h_tile + 128;
  }
  return L;
}

template <bool NOISE>
__global__ void __launch_bounds__(256, 2) gumbel_rows_forward_kernel(const GumbelParams P) {
  extern __shared__ __align__(128) uint8_t gr_smem[];
  const int W = P.W, H = P.H, W4 = W >> 2;
  const GrLayout L = gr_layout(W, H, false, NOISE, false);
  float* tab = reinterpret_cast<float*>(gr_smem);
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) tab[i] = __ldg(P.herbrand + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint8_t* wb = gr_smem + L.shared + (size_t)warp * L.per_warp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* t_buf = reinterpret_cast<float*>(wb + 128);
  float* n_buf = t_buf + L.w_tile / 4;
  float* h_out = (NOISE ? n_buf : t_buf) + L.w_tile / 4;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool s_al = aligned16(P.scores), n_al = NOISE && aligned16(P.noise);
  const bool y_al = P.y_soft != nullptr && aligned16(P.y_soft), h_al = P.world_h != nullptr && aligned16(P.world_h);
  WarpTileLoader<1> ls, ln;
  ls.init(t_buf, bars, L.w_tile / 4, lane);
  if (NOISE) ln.init(n_buf, bars + 1, L.w_tile / 4, lane);
  const int TW = kWarp * W, TH = kWarp * H;
  const float inv_tau = 1.0f / P.tau, sc_log = inv_tau * kLog2e;
  const bool is_log = P.is_log != 0;
  const float k = is_log ? sc_log : inv_tau;   // z = k * (is_log ? score : log2 score) = log2(e)/tau * ln p
  const uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
  float4* trow = reinterpret_cast<float4*>(t_buf + lane * W);
  const float4* nrow = reinterpret_cast<const float4*>(n_buf + lane * W);
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const int rows = (int)min((long long)kWarp, P.B - sub * kWarp);
    const long long row = sub * kWarp + lane;
    if (lane == 0) bulk_wait_read<0>();   // last tile's stores have read the buffers
    __syncwarp();
    ls.issue(0, P.scores + sub * TW, rows * W, s_al, lane, 1.0f);
    if (NOISE) ln.issue(0, P.noise + sub * TW, rows * W, n_al, lane, 0.0f);
    ls.wait(0);
    if (NOISE) ln.wait(0);
    // pass 1: scores -> perturbed logits (in place), running maximum and arg-max
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll 5
    for (int g = 0; g < W4; ++g) {
      const float4 sv = trow[g];
      const float sc[4] = {sv.x, sv.y, sv.z, sv.w};
      float nz[4], z[4];
      if (NOISE) {
        const float4 nv = nrow[g];
        nz[0] = nv.x; nz[1] = nv.y; nz[2] = nv.z; nz[3] = nv.w;
      } else {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)g, 1u), key);
        log2_exp_variates4(r, nz);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float l = k * (is_log ? sc[u] : fast_lg2(sc[u]));
        // injected g: z = l + log2(e)/tau g; Philox: z = l - L / tau (+ a row constant nobody sees)
        z[u] = NOISE ? fmaf(nz[u], sc_log, l) : fmaf(nz[u], -inv_tau, l);
        if (z[u] > mx) { mx = z[u]; arg = 4 * g + u; }
      }
      trow[g] = make_float4(z[0], z[1], z[2], z[3]);
    }
    // pass 2: exp2(z - max) in place, the sum
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 5
    for (int g = 0; g < W4; ++g) {
      float4 v = trow[g];
      v.x = fast_exp2(v.x - mx); v.y = fast_exp2(v.y - mx); v.z = fast_exp2(v.z - mx); v.w = fast_exp2(v.w - mx);
      s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
      if (P.y_soft != nullptr) trow[g] = v;
    }
    const float inv_sum = 1.0f / ((s0 + s1) + (s2 + s3));
    // pass 3: normalise in place
    if (P.y_soft != nullptr) {
#pragma unroll 5
      for (int g = 0; g < W4; ++g) {
        float4 v = trow[g];
        v.x *= inv_sum; v.y *= inv_sum; v.z *= inv_sum; v.w *= inv_sum;
        trow[g] = v;
      }
    }
    if (P.world_h != nullptr) {
      // straight-through forward value of the selected component: (1 - y_soft) + y_soft, y_soft[arg] = 1 / sum
      const float one = (1.0f - inv_sum) + inv_sum;
      const float* hsrc = tab + arg * H;
      float* hrow = h_out + lane * H;
      if ((H & 3) == 0) {
        for (int h = 0; h < H; h += 4) {
          const float4 t = *reinterpret_cast<const float4*>(hsrc + h);
          *reinterpret_cast<float4*>(hrow + h) = make_float4(one * t.x, one * t.y, one * t.z, one * t.w);
        }
      } else if ((H & 1) == 0) {
        for (int h = 0; h < H; h += 2) {
          const float2 t = *reinterpret_cast<const float2*>(hsrc + h);
          *reinterpret_cast<float2*>(hrow + h) = make_float2(one * t.x, one * t.y);
        }
      } else {
        for (int h = 0; h < H; ++h) hrow[h] = one * hsrc[h];
      }
    }
    if (P.world_idx != nullptr && lane < rows) P.world_idx[row] = arg;
    fence_proxy_async();
    __syncwarp();
    if (P.y_soft != nullptr) warp_tile_store(P.y_soft + sub * TW, t_buf, rows * W, TW, y_al, lane);
    if (P.world_h != nullptr) warp_tile_store(P.world_h + sub * TH, h_out, rows * H, TH, h_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
}

// G_e of one world from its padded (column offset, weight) list: LP entries, each a float2 {bit pattern of col * 32, w}
template <int LP>
__device__ __forceinline__ float world_grad(const float2* __restrict__ ent, int e, const float* __restrict__ gT_lane) {
  float acc = 0.0f;
#pragma unroll
  for (int l = 0; l < LP; l += 2) {
    const float4 en = *reinterpret_cast<const float4*>(ent + e * LP + l);   // two entries, one broadcast load
    acc = fmaf(gT_lane[__float_as_int(en.x)], en.y, acc);
    acc = fmaf(gT_lane[__float_as_int(en.z)], en.w, acc);
  }
  return acc;
}
// ... and from its bit mask (any number of non-zero columns)
__device__ __forceinline__ float world_grad_mask(uint32_t m, const float* __restrict__ tab_row,
                                                 const float* __restrict__ gT_lane) {
  float acc = 0.0f;
  while (m != 0u) {
    const int h = __ffs(m) - 1;
    m &= m - 1u;
    acc = fmaf(gT_lane[h * kWarp], tab_row[h], acc);
  }
  return acc;
}

__global__ void __launch_bounds__(256, 2) gumbel_rows_backward_kernel(const GumbelParams P) {
  extern __shared__ __align__(128) uint8_t gr_smem[];
  __shared__ int s_max_entries;
  const int W = P.W, H = P.H, W4 = W >> 2;
  const bool is_log = P.is_log != 0;
  const GrLayout L = gr_layout(W, H, true, false, !is_log);
  float* tab = reinterpret_cast<float*>(gr_smem);
  uint32_t* msk = reinterpret_cast<uint32_t*>(gr_smem + L.tab);
  float2* ent = reinterpret_cast<float2*>(gr_smem + L.tab + L.masks);
  if (threadIdx.x == 0) s_max_entries = 0;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) tab[i] = __ldg(P.herbrand + i);
  __syncthreads();
  for (int e = threadIdx.x; e < W; e += blockDim.x) {
    uint32_t m = 0u;
    for (int h = 0; h < H; ++h) m |= (tab[e * H + h] != 0.0f) ? (1u << h) : 0u;
    msk[e] = m;
    atomicMax(&s_max_entries, __popc(m));
  }
  __syncthreads();
  const int max_entries = s_max_entries;
  const int LP = max_entries <= 2 ? 2 : (max_entries <= kMaxPadEntries ? 4 : 0);   // 0: bit-mask walk
  if (LP != 0) {
    for (int e = threadIdx.x; e < W; e += blockDim.x) {
      uint32_t m = msk[e];
      for (int l = 0; l < LP; ++l) {
        // padding: the zero row below the transposed upstream gradient, weight 0
        int h = H;
        float w = 0.0f;
        if (m != 0u) { h = __ffs(m) - 1; m &= m - 1u; w = tab[e * H + h]; }
        ent[e * LP + l] = make_float2(__int_as_float(h * kWarp), w);
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint8_t* wb = gr_smem + L.shared + (size_t)warp * L.per_warp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* y_buf = reinterpret_cast<float*>(wb + 128);
  float* s_buf = y_buf + L.w_tile / 4;
  float* g_buf = s_buf + L.w_tile / 4;
  float* gT = g_buf + L.h_tile / 4;    // [H + 1][32]
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool y_al = aligned16(P.y_soft), s_al = aligned16(P.scores), g_al = aligned16(P.grad_world_h);
  const bool d_al = aligned16(P.grad_scores);
  WarpTileLoader<1> ly, lsc, lg;
  ly.init(y_buf, bars, L.w_tile / 4, lane);
  lg.init(g_buf, bars + 1, L.h_tile / 4, lane);
  if (!is_log) lsc.init(s_buf, bars + 2, L.w_tile / 4, lane);
  const int TW = kWarp * W, TH = kWarp * H;
  const float inv_tau = 1.0f / P.tau;
  float* gT_lane = gT + lane;
  gT_lane[H * kWarp] = 0.0f;
  float4* yrow = reinterpret_cast<float4*>(y_buf + lane * W);
  float4* srow = reinterpret_cast<float4*>(s_buf + lane * W);
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const int rows = (int)min((long long)kWarp, P.B - sub * kWarp);
    if (lane == 0) bulk_wait_read<0>();   // last tile's store has read its buffer
    __syncwarp();
    ly.issue(0, P.y_soft + sub * TW, rows * W, y_al, lane, 0.0f);
    lg.issue(0, P.grad_world_h + sub * TH, rows * H, g_al, lane, 0.0f);
    if (!is_log) lsc.issue(0, P.scores + sub * TW, rows * W, s_al, lane, 1.0f);
    lg.wait(0);
    // upstream gradient of this lane's row, column-major: a warp-uniform column index is then conflict-free
    {
      const float* grow = g_buf + lane * H;
      for (int h = 0; h < H; ++h) gT_lane[h * kWarp] = grow[h];
    }
    ly.wait(0);
    if (!is_log) lsc.wait(0);
    // G_e = sum_h gh[h] H[e,h] over the non-zero columns of world e (ascending h).  Pass 1: dot = sum_e y_e G_e, and
    // the row is rewritten in place: G_e over y_e, c_e = y_e / p_e (is_log: y_e) over the score.  Pass 2:
    // d logit_e [/ p_e] = c_e (G_e - dot) / tau over G_e — two shared loads and three flops per world instead of
    // evaluating G_e a second time (6 instructions per world: profiles/r02f, 39 % of the kernel) and dividing.
    auto G4 = [&](int g, float* G) {
      if (LP == 2) {
#pragma unroll
        for (int u = 0; u < 4; ++u) G[u] = world_grad<2>(ent, 4 * g + u, gT_lane);
      } else if (LP == 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) G[u] = world_grad<4>(ent, 4 * g + u, gT_lane);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) G[u] = world_grad_mask(msk[4 * g + u], tab + (4 * g + u) * H, gT_lane);
      }
    };
    float dot = 0.0f;
#pragma unroll 4
    for (int g = 0; g < W4; ++g) {
      const float4 yv = yrow[g];
      float G[4];
      G4(g, G);
      dot = fmaf(yv.x, G[0], dot); dot = fmaf(yv.y, G[1], dot);
      dot = fmaf(yv.z, G[2], dot); dot = fmaf(yv.w, G[3], dot);
      float4 c = yv;
      if (!is_log) {   // logit = log p: the gradient with respect to p carries 1 / p
        const float4 sv = srow[g];
        c = make_float4(__fdividef(yv.x, sv.x), __fdividef(yv.y, sv.y), __fdividef(yv.z, sv.z), __fdividef(yv.w, sv.w));
      }
      srow[g] = c;
      yrow[g] = make_float4(G[0], G[1], G[2], G[3]);
    }
#pragma unroll 5
    for (int g = 0; g < W4; ++g) {
      const float4 c = srow[g], G = yrow[g];
      yrow[g] = make_float4(c.x * (G.x - dot) * inv_tau, c.y * (G.y - dot) * inv_tau,
                            c.z * (G.z - dot) * inv_tau, c.w * (G.w - dot) * inv_tau);
    }
    fence_proxy_async();
    __syncwarp();
    warp_tile_store(P.grad_scores + sub * TW, y_buf, rows * W, TW, d_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
}

constexpr int kSmemBudget = 227 * 1024;

// warps per CTA and CTAs per SM that put the most warps on an SM within the shared-memory budget
struct GrShape { int nw, ctas, smem; };
GrShape gr_shape(const GrLayout& L) {
  GrShape best{0, 1, 0};
  for (int ctas = 1; ctas <= 4; ++ctas) {
    const int nw = min(8, (kSmemBudget / ctas - 1024 - L.shared) / L.per_warp);
    if (nw >= 1 && nw * ctas > best.nw * best.ctas) best = GrShape{nw, ctas, L.shared + nw * L.per_warp};
  }
  return best;
}

template <typename K>
cudaError_t launch_rows(K kernel, const GumbelParams& P, const GrLayout& L, int num_sms, cudaStream_t st) {
  const GrShape S = gr_shape(L);
  if (S.nw < 1) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S.smem);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const int grid = (int)tile_grid(n_sub, S.nw, (long long)num_sms * S.ctas);
  kernel<<<grid, S.nw * 32, S.smem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace

// Below ~16 Ki rows the one-warp-per-row kernels win: a lane here runs ~4.6 k instructions for its row whatever the
// batch, a ~14 us floor (scripts/gumbel_small_batch.py: 4096 rows 14.1 / 13.3 us against 5.9 / 6.9 us; 16 Ki rows
// 14.5 / 13.4 against 13.8 / 15.4; 256 Ki rows 76 / 84 against 174 / 214).  VAEL_GUMBEL_ROWS (read per call, so that
// tests can flip it): 0 = never, 1 = whenever the shape allows, unset = from kRowsMinBatch rows up.
constexpr long long kRowsMinBatch = 16384;

bool gumbel_rows_ok(const GumbelParams& P, bool backward) {
  const char* env = getenv("VAEL_GUMBEL_ROWS");
  const bool never = env != nullptr && env[0] == '0', always = env != nullptr && env[0] == '1';
  if (never || P.W < 4 || (P.W & 3) != 0 || P.W > 256 || P.H < 1 || P.H > 32 || P.B < 1) return false;
  if (!always && P.B < kRowsMinBatch) return false;
  const GrLayout L = gr_layout(P.W, P.H, backward, !backward && P.noise != nullptr, backward && !P.is_log);
  const GrShape S = gr_shape(L);
  return S.nw * S.ctas >= 4;   // fewer resident warps than that: the one-warp-per-row kernels
}

cudaError_t gumbel_rows_forward(const GumbelParams& P, int num_sms, cudaStream_t st) {
  const bool noise = P.noise != nullptr;
  const GrLayout L = gr_layout(P.W, P.H, false, noise, false);
  return noise ? launch_rows(gumbel_rows_forward_kernel<true>, P, L, num_sms, st)
               : launch_rows(gumbel_rows_forward_kernel<false>, P, L, num_sms, st);
}

cudaError_t gumbel_rows_backward(const GumbelParams& P, int num_sms, cudaStream_t st) {
  const GrLayout L = gr_layout(P.W, P.H, true, false, !P.is_log);
  return launch_rows(gumbel_rows_backward_kernel, P, L, num_sms, st);
}

}  // namespace vael
