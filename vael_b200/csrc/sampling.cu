// Gumbel-softmax world sampling + Herbrand projection, and the facts head.
//
// Reference: logits = log(worlds_prob); world = F.gumbel_softmax(logits, tau=1, hard=True);
// world_h = world @ herbrand_base      (models/vael.py:61-65, :82-84; Mario :335-338)
// and MNISTPairsVAELModel.compute_facts_probability (models/vael.py:159-177).
//
// One warp per batch row, lane <-> world: global accesses are coalesced and the
// max / argmax / sum reductions are warp shuffles; nothing is staged in shared
// memory except the small 0/1 Herbrand table.  These are the kernels of wide or odd shapes (W > 128, W % 4 != 0,
// H > 32: e.g. the 3-digit circuit's 1000 worlds); the shapes VAEL trains on take the lane <-> row kernels of
// gumbel_rows.cu.
#include "params.cuh"
#include "sampling.cuh"

namespace vael {

template <int ITEMS>
__global__ void __launch_bounds__(256) gumbel_forward_kernel(const GumbelParams P) {
  const int lane = threadIdx.x & 31;
  const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_tau = 1.0f / P.tau;
  // the next row's scores (and injected noise) are requested one iteration ahead: a warp works on one row at
  // a time, so without this every row would pay a full HBM round trip before its first instruction
  float s_nx[ITEMS], n_nx[ITEMS];
  auto fetch = [&](long long row) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int e = lane + j * kWarp;
      const bool on = row < P.B && e < P.W;
      s_nx[j] = on ? __ldg(P.scores + row * P.W + e) : 1.0f;
      n_nx[j] = (on && P.noise != nullptr) ? __ldg(P.noise + row * P.W + e) : 0.0f;
    }
  };
  fetch(gwarp);
  for (long long row = gwarp; row < P.B; row += nwarp) {
    float z[ITEMS], sc[ITEMS], nz[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) { sc[j] = s_nx[j]; nz[j] = n_nx[j]; }
    fetch(row + nwarp);
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    uint32_t rnd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int e = lane + j * kWarp;
      z[j] = -INFINITY;
      if (P.noise == nullptr && (j & 3) == 0) {
        // one Philox4x32-10 call = the four random words of this lane's items j..j+3 (worlds lane + 32 j'):
        // counter = (row, lane | block of four items), key = seed
        const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(lane | ((j >> 2) << 5)), 0u),
                                      make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
        rnd[0] = r.x; rnd[1] = r.y; rnd[2] = r.z; rnd[3] = r.w;
      }
      if (e < P.W) {
        const float logit = P.is_log ? sc[j] : logf(sc[j]);
        const float g = (P.noise != nullptr) ? nz[j] : gumbel_from_bits(rnd[j & 3]);
        z[j] = (logit + g) * inv_tau;
        if (z[j] > mx) { mx = z[j]; arg = e; }
      }
    }
    // warp arg-max, ties -> lowest world index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
      if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      z[j] = (lane + j * kWarp < P.W) ? expf(z[j] - mx) : 0.0f;
      sum += z[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float ys_arg = 0.0f;
    const float inv_sum = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int e = lane + j * kWarp;
      const float ys = z[j] * inv_sum;
      if (e < P.W) {
        if (P.y_soft != nullptr) P.y_soft[row * P.W + e] = ys;
        if (e == arg) ys_arg = ys;
      }
    }
    ys_arg = __shfl_sync(0xffffffffu, ys_arg, arg & 31);
    // straight-through forward value of the selected component: (1 - y_soft) + y_soft
    const float one = (1.0f - ys_arg) + ys_arg;
    if (P.world_idx != nullptr && lane == 0) P.world_idx[row] = arg;
    if (P.world_h != nullptr) {
      for (int h = lane; h < P.H; h += kWarp) P.world_h[row * P.H + h] = one * __ldg(P.herbrand + (long long)arg * P.H + h);
    }
  }
}

template <int ITEMS>
__global__ void __launch_bounds__(256) gumbel_backward_kernel(const GumbelParams P) {
  extern __shared__ float hb_s[];  // [W][H | 1] Herbrand table (when it fits; else read through L1/L2)
  const float* hb = P.herbrand;
  int hp = P.H;                    // row pitch of `hb`
  if (P.hb_in_smem) {
    hp = P.H | 1;                  // odd pitch: lanes (= consecutive worlds) read column h without bank conflicts
    for (int i = threadIdx.x; i < P.W * P.H; i += blockDim.x) {
      const int w = i / P.H, h = i - w * P.H;
      hb_s[w * hp + h] = __ldg(P.herbrand + i);
    }
    __syncthreads();
    hb = hb_s;
  }
  const int lane = threadIdx.x & 31;
  const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_tau = 1.0f / P.tau;
  for (long long row = gwarp; row < P.B; row += nwarp) {
    // grad wrt the (straight-through) world vector: G_e = sum_h gh[h] * H[e,h]
    float G[ITEMS], ys[ITEMS];
    float dot = 0.0f;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int e = lane + j * kWarp;
      G[j] = 0.0f;
      ys[j] = (e < P.W) ? __ldg(P.y_soft + row * P.W + e) : 0.0f;
    }
    // h outer: every upstream gradient component is fetched once per row (a warp-wide broadcast) and feeds the
    // ITEMS accumulation chains of the lane; each chain still adds in h order
    for (int h = 0; h < P.H; ++h) {
      const float gh = __ldg(P.grad_world_h + row * P.H + h);
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const int e = min(lane + j * kWarp, P.W - 1);
        G[j] = fmaf(gh, hb[e * hp + h], G[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if (lane + j * kWarp < P.W) dot = fmaf(ys[j], G[j], dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int e = lane + j * kWarp;
      if (e < P.W) {
        float d = ys[j] * (G[j] - dot) * inv_tau;  // d loss / d logit_e
        if (!P.is_log) d = d / __ldg(P.scores + row * P.W + e);  // logit = log p
        P.grad_scores[row * P.W + e] = d;
      }
    }
  }
}

static int pick_items(int W) {
  const int need = (W + kWarp - 1) / kWarp;
  const int opts[] = {1, 2, 3, 4, 8, 16, 32};
  for (int o : opts)
    if (need <= o) return o;
  return -1;
}

static int launch_grid(long long B, int num_sms) {
  const long long warps_per_block = 8;
  long long blocks = (B + warps_per_block - 1) / warps_per_block;
  return (int)max(1LL, min(blocks, (long long)num_sms * 8));
}

cudaError_t gumbel_forward(const GumbelParams& P, int num_sms, cudaStream_t st) {
  if (gumbel_rows_ok(P, false)) return gumbel_rows_forward(P, num_sms, st);
  const int grid = launch_grid(P.B, num_sms);
  switch (pick_items(P.W)) {
    case 1: gumbel_forward_kernel<1><<<grid, 256, 0, st>>>(P); break;
    case 2: gumbel_forward_kernel<2><<<grid, 256, 0, st>>>(P); break;
    case 3: gumbel_forward_kernel<3><<<grid, 256, 0, st>>>(P); break;
    case 4: gumbel_forward_kernel<4><<<grid, 256, 0, st>>>(P); break;
    case 8: gumbel_forward_kernel<8><<<grid, 256, 0, st>>>(P); break;
    case 16: gumbel_forward_kernel<16><<<grid, 256, 0, st>>>(P); break;
    case 32: gumbel_forward_kernel<32><<<grid, 256, 0, st>>>(P); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <int ITEMS>
static cudaError_t launch_gumbel_bwd(GumbelParams P, int grid, cudaStream_t st) {
  size_t smem = (size_t)P.W * (P.H | 1) * sizeof(float);
  P.hb_in_smem = smem <= (size_t)200 * 1024;
  if (!P.hb_in_smem) smem = 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gumbel_backward_kernel<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  gumbel_backward_kernel<ITEMS><<<grid, 256, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t gumbel_backward(const GumbelParams& P, int num_sms, cudaStream_t st) {
  if (gumbel_rows_ok(P, true)) return gumbel_rows_backward(P, num_sms, st);
  const int grid = launch_grid(P.B, num_sms);
  switch (pick_items(P.W)) {
    case 1: return launch_gumbel_bwd<1>(P, grid, st);
    case 2: return launch_gumbel_bwd<2>(P, grid, st);
    case 3: return launch_gumbel_bwd<3>(P, grid, st);
    case 4: return launch_gumbel_bwd<4>(P, grid, st);
    case 8: return launch_gumbel_bwd<8>(P, grid, st);
    case 16: return launch_gumbel_bwd<16>(P, grid, st);
    case 32: return launch_gumbel_bwd<32>(P, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------
// facts head: per group softmax, + eps, / detached sum   (models/vael.py:159-177)
//
// Structured kernels for the shapes VAEL uses (NG groups of N logits, N*NG floats per row): one warp per
// 32-row tile, the contiguous [32, NG*N] tile is staged in shared memory with coalesced 128-bit accesses,
// lane <-> row works on registers (compile-time indices), results go back through the same tile.
// HBM-bound: 2 * NG*N*4 bytes per row forward, 3 * NG*N*4 backward (logits re-read, softmax recomputed).
// Any other shape takes the generic one-thread-per-(row, group) kernels below.
// ---------------------------------------------------------------------------
// CLIP: out = clamp(softmax, eps, hi) instead of (softmax + eps) / sum — MarioVAELModel's facts
// (models/vael.py:319-322: Softmax then torch.clip(min=1e-5, max=0.9999)); its gradient passes where
// eps <= softmax <= hi (torch.clip's rule) and is zero elsewhere.
constexpr int kFactsStages = 2;
template <int F, bool BWD>
__host__ __device__ constexpr int facts_warp_bytes() {
  // [barriers 128][input stages: logits (+ grad_facts)][result stages]
  return 128 + kFactsStages * (BWD ? 2 : 1) * ((kWarp * F * 4 + 127) / 128 * 128) + kFactsStages * ((kWarp * F * 4 + 127) / 128 * 128);
}

// Round 2: the tiles move by bulk copies through a warp-private two-stage ring (WarpTileLoader) instead of being
// staged through the LSU at an odd pitch (ncu, profiles/r02a_stage_ncu.md: L1 45-55 % busy, 2.6 M bank conflicts).
template <int NG, int N, bool BWD, bool CLIP>
__global__ void __launch_bounds__(128) facts_tile_kernel(const float* __restrict__ logits,
                                                         const float* __restrict__ grad_facts, long long B,
                                                         float eps, float hi, float* __restrict__ out) {
  constexpr int F = NG * N, FT = kWarp * F, TB = (FT * 4 + 127) / 128 * 128;
  extern __shared__ __align__(128) uint8_t facts_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint8_t* wb = facts_smem + (size_t)warp * facts_warp_bytes<F, BWD>();
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* istage = reinterpret_cast<float*>(wb + 128);
  float* ostage = reinterpret_cast<float*>(wb + 128 + kFactsStages * (BWD ? 2 : 1) * TB);
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool in_al = aligned16(logits) && (!BWD || aligned16(grad_facts)), out_al = aligned16(out);
  WarpTileLoader<kFactsStages> lx, lg;
  lx.init(istage, bars, TB / 4, lane);
  if (BWD) lg.init(istage + kFactsStages * (TB / 4), bars + kFactsStages, TB / 4, lane);
  auto request = [&](long long sub, int s) {
    if (sub >= n_sub) return;
    const int n = (int)min((long long)kWarp, B - sub * kWarp) * F;
    lx.issue(s, logits + sub * FT, n, in_al, lane, 0.0f);
    if (BWD) lg.issue(s, grad_facts + sub * FT, n, in_al, lane, 0.0f);
  };
#pragma unroll
  for (int s = 0; s < kFactsStages; ++s) request(gw + s * GW, s);
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % kFactsStages);
    const int rows = (int)min((long long)kWarp, B - sub * kWarp);
    lx.wait(s);
    if (BWD) lg.wait(s);
    float x[F], y[F], rz[NG], gfr[BWD ? F : 1], o[F];
    tile_row_load<F>(lx.stage(s) + lane * F, x);
    if (BWD) tile_row_load<F>(lg.stage(s) + lane * F, gfr);
    __syncwarp();
    request(sub + kFactsStages * GW, s);
    facts_row<NG, N>(x, eps, y, rz);
    if (!BWD) {
#pragma unroll
      for (int c = 0; c < F; ++c) o[c] = CLIP ? fminf(fmaxf(y[c], eps), hi) : (y[c] + eps) * rz[c / N];
    } else {
      // f = (y + eps) / Z with Z detached -> dL/dy_i = gf_i / Z (CLIP: gf_i where eps <= y_i <= hi, else 0);
      // softmax: dx_i = y_i (dy_i - sum_j y_j dy_j)
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        float dy[N], dot = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float yi = y[g * N + i];
          dy[i] = CLIP ? ((yi >= eps && yi <= hi) ? gfr[g * N + i] : 0.0f) : gfr[g * N + i] * rz[g];
          dot = fmaf(yi, dy[i], dot);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) o[g * N + i] = y[g * N + i] * (dy[i] - dot);
      }
    }
    if (lane == 0) bulk_wait_read<kFactsStages - 1>();   // the store of two tiles ago has read its stage
    __syncwarp();
    tile_row_store<F>(ostage + s * (TB / 4) + lane * F, o);
    fence_proxy_async();
    __syncwarp();
    warp_tile_store(out + sub * FT, ostage + s * (TB / 4), rows * F, FT, out_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
}

template <int NG, int N, bool BWD, bool CLIP>
static cudaError_t launch_facts_tile(const float* logits, const float* gf, long long B, float eps, float hi, float* out,
                                     int num_sms, cudaStream_t st) {
  const int nw = 4;
  const int smem = nw * facts_warp_bytes<NG * N, BWD>();
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const int ctas = max(1, min(8, (227 * 1024) / (smem + 1024)));
  const int grid = (int)tile_grid(n_sub, nw, (long long)num_sms * ctas);
  cudaError_t e = cudaFuncSetAttribute(facts_tile_kernel<NG, N, BWD, CLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  facts_tile_kernel<NG, N, BWD, CLIP><<<grid, nw * 32, smem, st>>>(logits, gf, B, eps, hi, out);
  return cudaGetLastError();
}

template <bool BWD>
static bool facts_tile_dispatch(const float* logits, const float* gf, long long B, int n_groups, int group, float eps,
                                float hi, float* out, int num_sms, cudaStream_t st, cudaError_t* err) {
  if (hi > 0.0f) {  // clip mode: the Mario shapes
    if (n_groups == 2 && group == 9) { *err = launch_facts_tile<2, 9, BWD, true>(logits, gf, B, eps, hi, out, num_sms, st); return true; }
    if (n_groups == 1 && group == 9) { *err = launch_facts_tile<1, 9, BWD, true>(logits, gf, B, eps, hi, out, num_sms, st); return true; }
    return false;
  }
  if (n_groups == 2 && group == 10) { *err = launch_facts_tile<2, 10, BWD, false>(logits, gf, B, eps, hi, out, num_sms, st); return true; }
  if (n_groups == 3 && group == 10) { *err = launch_facts_tile<3, 10, BWD, false>(logits, gf, B, eps, hi, out, num_sms, st); return true; }
  if (n_groups == 2 && group == 9) { *err = launch_facts_tile<2, 9, BWD, false>(logits, gf, B, eps, hi, out, num_sms, st); return true; }
  if (n_groups == 1 && group == 9) { *err = launch_facts_tile<1, 9, BWD, false>(logits, gf, B, eps, hi, out, num_sms, st); return true; }
  return false;
}

// generic shapes: one thread per (row, group); a group is `group` contiguous floats
__global__ void __launch_bounds__(256) facts_forward_kernel(const float* __restrict__ logits, long long n_groups_total,
                                                            int group, float eps, float hi, float* __restrict__ facts) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < n_groups_total; gi += stride) {
    const float* x = logits + gi * group;
    float* y = facts + gi * group;
    float mx = -INFINITY;
    for (int i = 0; i < group; ++i) mx = fmaxf(mx, __ldg(x + i));
    float s = 0.0f;
    for (int i = 0; i < group; ++i) s += expf(__ldg(x + i) - mx);
    const float rs = 1.0f / s;   // reciprocals per group, as facts_row
    if (hi > 0.0f) {  // clip mode
      for (int i = 0; i < group; ++i) y[i] = fminf(fmaxf(expf(__ldg(x + i) - mx) * rs, eps), hi);
      continue;
    }
    float zsum = 0.0f;
    for (int i = 0; i < group; ++i) {
      const float v = expf(__ldg(x + i) - mx) * rs + eps;
      y[i] = v;
      zsum += v;
    }
    const float rz = 1.0f / zsum;
    for (int i = 0; i < group; ++i) y[i] = y[i] * rz;
  }
}

__global__ void __launch_bounds__(256) facts_backward_kernel(const float* __restrict__ logits,
                                                             const float* __restrict__ grad_facts,
                                                             long long n_groups_total, int group, float eps,
                                                             float hi, float* __restrict__ grad_logits) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < n_groups_total; gi += stride) {
    const float* x = logits + gi * group;
    const float* gf = grad_facts + gi * group;
    float* gx = grad_logits + gi * group;
    float mx = -INFINITY;
    for (int i = 0; i < group; ++i) mx = fmaxf(mx, __ldg(x + i));
    float s = 0.0f;
    for (int i = 0; i < group; ++i) s += expf(__ldg(x + i) - mx);
    const float rs = 1.0f / s;
    if (hi > 0.0f) {  // clip mode: dy_i = gf_i where eps <= y_i <= hi, else 0
      float dotc = 0.0f;
      for (int i = 0; i < group; ++i) {
        const float yi = expf(__ldg(x + i) - mx) * rs;
        dotc = fmaf(yi, (yi >= eps && yi <= hi) ? __ldg(gf + i) : 0.0f, dotc);
      }
      for (int i = 0; i < group; ++i) {
        const float yi = expf(__ldg(x + i) - mx) * rs;
        gx[i] = yi * (((yi >= eps && yi <= hi) ? __ldg(gf + i) : 0.0f) - dotc);
      }
      continue;
    }
    float zsum = 0.0f;
    for (int i = 0; i < group; ++i) zsum += expf(__ldg(x + i) - mx) * rs + eps;
    const float rz = 1.0f / zsum;
    // f = (y + eps) / Z with Z detached  ->  dL/dy_i = gf_i / Z ; softmax: dx_i = y_i (dy_i - sum_j y_j dy_j)
    float dot = 0.0f;
    for (int i = 0; i < group; ++i) dot = fmaf(expf(__ldg(x + i) - mx) * rs, __ldg(gf + i) * rz, dot);
    for (int i = 0; i < group; ++i) {
      const float yi = expf(__ldg(x + i) - mx) * rs;
      gx[i] = yi * (__ldg(gf + i) * rz - dot);
    }
  }
}

// hi <= 0: (softmax + eps) / sum (MNIST); hi > 0: clamp(softmax, eps, hi) (Mario)
cudaError_t facts_forward(const float* logits, long long B, int n_groups, int group, float eps, float hi, float* facts,
                          int num_sms, cudaStream_t st) {
  cudaError_t e;
  if (facts_tile_dispatch<false>(logits, nullptr, B, n_groups, group, eps, hi, facts, num_sms, st, &e)) return e;
  const long long total = B * n_groups;
  const int grid = (int)max(1LL, min((total + 255) / 256, (long long)num_sms * 8));
  facts_forward_kernel<<<grid, 256, 0, st>>>(logits, total, group, eps, hi, facts);
  return cudaGetLastError();
}
cudaError_t facts_backward(const float* logits, const float* grad_facts, long long B, int n_groups, int group,
                           float eps, float hi, float* grad_logits, int num_sms, cudaStream_t st) {
  cudaError_t e;
  if (facts_tile_dispatch<true>(logits, grad_facts, B, n_groups, group, eps, hi, grad_logits, num_sms, st, &e)) return e;
  const long long total = B * n_groups;
  const int grid = (int)max(1LL, min((total + 255) / 256, (long long)num_sms * 8));
  facts_backward_kernel<<<grid, 256, 0, st>>>(logits, grad_facts, total, group, eps, hi, grad_logits);
  return cudaGetLastError();
}

}  // namespace vael
