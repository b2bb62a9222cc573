// Structured fast path: circuits whose world level is the outer product of two
// annotated disjunctions (2-digit MNIST addition: 10x10 worlds; Mario: 9x9) and
// whose query nodes are sums over world nodes.
//
// Reference semantics: MNISTPairsVAELModel.problog_inference / compute_query
// (models/vael.py:200-225) == query-mode sdd.evaluate (models/vael.py:98-117);
// evidence mode == problog_inference_with_evidence (models/vael.py:119-135).
//
// Execution model (sm_100a): warp-autonomous pipelines.  Every warp owns a ring
// of shared-memory stages and streams 32-row sub-tiles of the batch-major
// tensors through them with 1-D bulk copies (cp.async.bulk, the TMA engine):
//   global facts[32,F] --bulk--> smem --LDS.128--> registers (one batch row per
//   lane) --FMUL/FADD--> STS.128 --> smem worlds[32,W] --bulk--> global.
// A 32-row sub-tile of a row-major [B,F] / [B,W] tensor is one contiguous span,
// so each transfer is a single bulk copy; lanes touch shared memory with
// conflict-free 128-bit accesses (row strides 80 B / 400 B land 8 consecutive
// lanes on 8 distinct 16-byte bank groups).  There is no block-level barrier.
#include <type_traits>

#include "params.cuh"

namespace vael {

template <int NA, int NB>
struct Ad2Shape {
  static constexpr int F = NA + NB;
  static constexpr int W = NA * NB;
  static constexpr int MW = (W + 31) / 32;
  static constexpr int FV = (F % 4 == 0) ? 4 : ((F % 2 == 0) ? 2 : 1);
  static constexpr int WV = (W % 4 == 0) ? 4 : ((W % 2 == 0) ? 2 : 1);
  static constexpr int FACT_BYTES = kWarp * F * 4;  // multiple of 128
  static constexpr int OUT_BYTES = kWarp * W * 4;   // multiple of 128
};


template <int V>
__device__ __forceinline__ void lds_vec(float* dst, const float* src) {
  if constexpr (V == 4) {
    float4 t = *reinterpret_cast<const float4*>(src);
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = *reinterpret_cast<const float2*>(src);
    dst[0] = t.x; dst[1] = t.y;
  } else {
    dst[0] = src[0];
  }
}
template <int V>
__device__ __forceinline__ void sts_vec(float* dst, const float* src) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(src[0], src[1]);
  } else {
    dst[0] = src[0];
  }
}

// plain (non-bulk) copies for tails / unaligned buffers
__device__ __forceinline__ void warp_copy(float* dst, const float* src, int n, int lane) {
  for (int i = lane; i < n; i += kWarp) dst[i] = src[i];
}

// normaliser Z = prod over ADs of (sum(p) + (1 - sum(p))): ProbLog's WMC of the AD
// constraints with the null-choice atom (ad_complement); == 1 up to rounding.
template <int NA, int NB>
__device__ __forceinline__ float ad2_normaliser(const float* p) {
  float sa = p[0];
#pragma unroll
  for (int i = 1; i < NA; ++i) sa = __fadd_rn(sa, p[i]);
  float sb = p[NA];
#pragma unroll
  for (int k = 1; k < NB; ++k) sb = __fadd_rn(sb, p[NA + k]);
  const float za = __fadd_rn(sa, __fsub_rn(1.0f, sa));
  const float zb = __fadd_rn(sb, __fsub_rn(1.0f, sb));
  return __fmul_rn(za, zb);
}

template <int MW>
__device__ __forceinline__ void load_mask(uint32_t* m, const Ad2Params& P, long long row, bool valid) {
#pragma unroll
  for (int i = 0; i < MW; ++i) m[i] = 0u;
  if (!valid) return;
  const int32_t q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
  if (q >= 0 && q < P.n_queries) {
#pragma unroll
    for (int i = 0; i < MW; ++i) m[i] = __ldg(P.qmask + (long long)q * MW + i);
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int NA, int NB, bool NORM, bool EVID, int S, int SO, bool STG = false>
__global__ void __launch_bounds__(256) ad2_forward_kernel(const Ad2Params P) {
  using Sh = Ad2Shape<NA, NB>;
  constexpr int F = Sh::F, W = Sh::W, MW = Sh::MW, FV = Sh::FV, WV = Sh::WV;
  constexpr int PER_WARP = S * Sh::FACT_BYTES + SO * Sh::OUT_BYTES + 128;
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  uint8_t* wbase = smem + (size_t)warp * PER_WARP;
  float* fact_s = reinterpret_cast<float*>(wbase);
  float* out_s = reinterpret_cast<float*>(wbase + S * Sh::FACT_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(wbase + S * Sh::FACT_BYTES + SO * Sh::OUT_BYTES);

  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;

  const bool in_al = aligned16(P.facts);
  const bool out_al = P.worlds != nullptr && aligned16(P.worlds);
  const uint64_t pol = l2_policy_evict_first();

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  uint32_t phase = 0;    // bit s: parity the next wait on stage s uses
  uint32_t via_tma = 0;  // bit s: the in-flight load of stage s went through the bulk engine

  auto load_sub = [&](long long sub, int s) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const uint32_t bytes = (uint32_t)rows * F * 4u;
    float* dst = fact_s + s * kWarp * F;
    const float* src = P.facts + row0 * F;
    if (in_al && (bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[s], bytes);
        if (P.l2_hint >= 2) bulk_g2s_hint(dst, src, bytes, &full[s], pol);
        else bulk_g2s(dst, src, bytes, &full[s]);
      }
      via_tma |= 1u << s;
    } else {
      warp_copy(dst, src, rows * F, lane);
      via_tma &= ~(1u << s);
    }
  };

  // prologue
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const long long sub = gw + (long long)s * GW;
    if (sub < n_sub) load_sub(sub, s);
  }

  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % S);
    const int so = (int)(it % SO);
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const long long row = row0 + lane;
    const bool valid = lane < rows;

    uint32_t m[MW];
    load_mask<MW>(m, P, row, valid);

    if (via_tma & (1u << s)) {
      mbar_wait(&full[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    } else {
      __syncwarp();
    }

    // this lane's batch row -> registers
    float p[F];
    {
      const float* src = fact_s + (s * kWarp + lane) * F;
#pragma unroll
      for (int c = 0; c < F / FV; ++c) lds_vec<FV>(p + c * FV, src + c * FV);
    }
    __syncwarp();
    // refill this stage with the sub-tile S iterations ahead
    {
      const long long nxt = sub + (long long)S * GW;
      if (nxt < n_sub) load_sub(nxt, s);
    }

    // the bulk store issued SO iterations ago must have finished reading out_s[so]
    if (lane == 0) bulk_wait_read<SO - 1>();
    __syncwarp();

    float z = 1.0f;
    if constexpr (NORM) z = ad2_normaliser<NA, NB>(p);

    float acc = 0.0f;  // value of the selected query node
    if constexpr (EVID) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const float v = __fmul_rn(p[w / NB], p[NA + (w % NB)]);
        if ((m[w >> 5] >> (w & 31)) & 1u) acc = __fadd_rn(acc, v);
      }
    }

    float* orow = out_s + (so * kWarp + lane) * W;
#pragma unroll
    for (int c = 0; c < W / WV; ++c) {
      float o[WV];
#pragma unroll
      for (int u = 0; u < WV; ++u) {
        const int w = c * WV + u;
        float v = __fmul_rn(p[w / NB], p[NA + (w % NB)]);
        const bool member = (m[w >> 5] >> (w & 31)) & 1u;
        if constexpr (EVID) {
          v = member ? v / acc : 0.0f;  // P(w | e) = [w |= e] P(w) / P(e)
        } else {
          if (member) acc = __fadd_rn(acc, v);
          if constexpr (NORM) v = v / z;
        }
        o[u] = v;
      }
      sts_vec<WV>(orow + c * WV, o);
    }

    if (P.query != nullptr && valid) {
      float qv = acc;  // WMC of the selected query node, sequential fp32 sum in world order
      if constexpr (NORM) qv = acc / z;
      P.query[row] = qv;
    }

    if (P.worlds != nullptr) {
      const uint32_t bytes = (uint32_t)rows * W * 4u;
      float* dst = P.worlds + row0 * W;
      const float* src = out_s + so * kWarp * W;
      if constexpr (STG) {
        // experiment: drain the tile with coalesced 128-bit stores instead of the bulk engine
        __syncwarp();
        if (out_al && (bytes & 15u) == 0) {
          const float4* s4 = reinterpret_cast<const float4*>(src);
          float4* d4 = reinterpret_cast<float4*>(dst);
          const int n4 = (int)(bytes >> 4);
#pragma unroll 5
          for (int i = lane; i < n4; i += kWarp) __stcs(d4 + i, s4[i]);
        } else {
          warp_copy(dst, src, rows * W, lane);
        }
        __syncwarp();
      } else if (out_al && (bytes & 15u) == 0) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (P.l2_hint >= 1) bulk_s2g_hint(dst, src, bytes, pol);
          else bulk_s2g(dst, src, bytes);
          bulk_commit();
        }
      } else {
        __syncwarp();
        warp_copy(dst, src, rows * W, lane);
      }
    }
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}

// ---------------------------------------------------------------------------
// backward (adjoint); values are recomputed from the facts, nothing was saved
// ---------------------------------------------------------------------------
template <int NA, int NB, bool NORM, int S, int SO>
__global__ void __launch_bounds__(256) ad2_backward_kernel(const Ad2Params P) {
  using Sh = Ad2Shape<NA, NB>;
  constexpr int F = Sh::F, W = Sh::W, MW = Sh::MW, FV = Sh::FV, WV = Sh::WV;
  constexpr int STAGE_IN = Sh::FACT_BYTES + Sh::OUT_BYTES;
  constexpr int PER_WARP = S * STAGE_IN + SO * Sh::FACT_BYTES + 128;
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  uint8_t* wbase = smem + (size_t)warp * PER_WARP;
  uint8_t* gout_base = wbase + S * STAGE_IN;
  uint64_t* full = reinterpret_cast<uint64_t*>(gout_base + SO * Sh::FACT_BYTES);

  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;

  const bool has_gw = P.grad_worlds != nullptr;
  const bool in_al = aligned16(P.facts) && (!has_gw || aligned16(P.grad_worlds));
  const bool out_al = aligned16(P.grad_facts);
  const uint64_t pol = l2_policy_evict_first();

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  uint32_t phase = 0, via_tma = 0;

  auto load_sub = [&](long long sub, int s) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const uint32_t fbytes = (uint32_t)rows * F * 4u;
    const uint32_t wbytes = has_gw ? (uint32_t)rows * W * 4u : 0u;
    float* fdst = reinterpret_cast<float*>(wbase + s * STAGE_IN);
    float* wdst = reinterpret_cast<float*>(wbase + s * STAGE_IN + Sh::FACT_BYTES);
    const float* fsrc = P.facts + row0 * F;
    const float* wsrc = has_gw ? P.grad_worlds + row0 * W : nullptr;
    if (in_al && ((fbytes | wbytes) & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[s], fbytes + wbytes);
        if (P.l2_hint >= 2) {
          bulk_g2s_hint(fdst, fsrc, fbytes, &full[s], pol);
          if (has_gw) bulk_g2s_hint(wdst, wsrc, wbytes, &full[s], pol);
        } else {
          bulk_g2s(fdst, fsrc, fbytes, &full[s]);
          if (has_gw) bulk_g2s(wdst, wsrc, wbytes, &full[s]);
        }
      }
      via_tma |= 1u << s;
    } else {
      warp_copy(fdst, fsrc, rows * F, lane);
      if (has_gw) warp_copy(wdst, wsrc, rows * W, lane);
      via_tma &= ~(1u << s);
    }
  };

#pragma unroll
  for (int s = 0; s < S; ++s) {
    const long long sub = gw + (long long)s * GW;
    if (sub < n_sub) load_sub(sub, s);
  }

  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % S);
    const int so = (int)(it % SO);
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const long long row = row0 + lane;
    const bool valid = lane < rows;

    uint32_t m[MW];
    load_mask<MW>(m, P, row, valid);
    float gq = 0.0f;
    if (P.grad_query != nullptr && valid) gq = __ldg(P.grad_query + row);

    if (via_tma & (1u << s)) {
      mbar_wait(&full[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    } else {
      __syncwarp();
    }

    const float* frow = reinterpret_cast<const float*>(wbase + s * STAGE_IN) + lane * F;
    const float* grow = reinterpret_cast<const float*>(wbase + s * STAGE_IN + Sh::FACT_BYTES) + lane * W;

    float p[F];
#pragma unroll
    for (int c = 0; c < F / FV; ++c) lds_vec<FV>(p + c * FV, frow + c * FV);

    float z = 1.0f;
    if constexpr (NORM) z = ad2_normaliser<NA, NB>(p);

    float g[F];
#pragma unroll
    for (int f = 0; f < F; ++f) g[f] = 0.0f;

#pragma unroll
    for (int c = 0; c < W / WV; ++c) {
      float a[WV];
      if (has_gw) {
        lds_vec<WV>(a, grow + c * WV);
      } else {
#pragma unroll
        for (int u = 0; u < WV; ++u) a[u] = 0.0f;
      }
#pragma unroll
      for (int u = 0; u < WV; ++u) {
        const int w = c * WV + u;
        float av = a[u];
        if ((m[w >> 5] >> (w & 31)) & 1u) av += gq;
        if constexpr (NORM) av = av / z;  // Z's own adjoint is exactly 0 (d/dp of s + (1 - s))
        g[w / NB] = fmaf(av, p[NA + (w % NB)], g[w / NB]);
        g[NA + (w % NB)] = fmaf(av, p[w / NB], g[NA + (w % NB)]);
      }
    }
    __syncwarp();
    {
      const long long nxt = sub + (long long)S * GW;
      if (nxt < n_sub) load_sub(nxt, s);
    }

    if (lane == 0) bulk_wait_read<SO - 1>();
    __syncwarp();

    float* gdst_s = reinterpret_cast<float*>(gout_base + so * Sh::FACT_BYTES);
    {
      float* orow = gdst_s + lane * F;
#pragma unroll
      for (int c = 0; c < F / FV; ++c) sts_vec<FV>(orow + c * FV, g + c * FV);
    }
    {
      const uint32_t bytes = (uint32_t)rows * F * 4u;
      float* dst = P.grad_facts + row0 * F;
      if (out_al && (bytes & 15u) == 0) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (P.l2_hint >= 1) bulk_s2g_hint(dst, gdst_s, bytes, pol);
          else bulk_s2g(dst, gdst_s, bytes);
          bulk_commit();
        }
      } else {
        __syncwarp();
        warp_copy(dst, gdst_s, rows * F, lane);
      }
    }
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}

// ---------------------------------------------------------------------------
// all query nodes at once: queries_all[b, q] = sum over the member worlds of q, in world order
// (the worlds[B,W] @ w_q[W,Q] mask contraction of utils/mnist_utils/metrics.py:71-75 and
// compute_query, models/vael.py:200-206, fused after the product level: the worlds never leave the SM).
// One warp per 32-row tile, lane <-> row (round 2: no world-major tile of all W products any more).  A query is the
// sum of its member worlds' products p1[i] p2[k]; only the MEMBERS are formed — Mario's five queries hold 48 of the
// 81 worlds' products between them.  World order is i-major, so the fold runs over i = 0 .. NA-1 as a COMPILE-TIME
// loop (p1[i] is a register operand) and, per (query, i), over that pair's k's: ONE shared-memory gather per member,
// from a column-major tile of the lane's p2 values (p2s[k][lane], plus an all-zero row for the padding) where "p2 of
// column k", k a warp-uniform table entry, is one conflict-free LDS.  The table — per (query, i) `len[q]` byte
// offsets, padded with the zero row — sits in the kernel's PARAMETER block (constant bank): its warp-uniform
// fetches cost no shared-memory wavefront, which together with instruction issue is what bounds this kernel
// (profiles/r02d_queries_ncu.md: the first round-2 version, both literals gathered and the table in shared memory,
// spent 251 wavefronts per tile).  The facts tile arrives by one bulk copy per tile (WarpTileLoader), the [32,Q]
// result tile leaves by one bulk store.  Every fold is sequential in world order with __fmul_rn / __fadd_rn
// (padding adds p1[i] * 0 = +0): bit-exact against the CSR fold.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float q_lds(uint32_t a) {
  float x;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
  return x;
}

constexpr int kQStages = 2;
template <int NA, int NB>
__host__ __device__ constexpr int queries_per_warp_bytes(int Q) {
  // [barriers 128][facts stages][p2s (NB+1) x 32][result stages 32 x Q]
  return 128 + kQStages * Ad2Shape<NA, NB>::FACT_BYTES + ((NB + 1) * kWarp * 4 + 127) / 128 * 128 +
         kQStages * ((kWarp * Q * 4 + 127) / 128 * 128);
}

template <int NA, int NB, bool NORM>
__global__ void __launch_bounds__(256) ad2_queries_kernel(const Ad2QParams P) {
  using Sh = Ad2Shape<NA, NB>;
  constexpr int F = Sh::F, FV = Sh::FV;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int Q = P.n_queries;
  const int per_warp = queries_per_warp_bytes<NA, NB>(Q);
  const int qbytes = (kWarp * Q * 4 + 127) / 128 * 128;
  uint8_t* wb = smem + (size_t)warp * per_warp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* fstage = reinterpret_cast<float*>(wb + 128);
  float* p2s = reinterpret_cast<float*>(wb + 128 + kQStages * Sh::FACT_BYTES);                          // [NB+1][32]
  float* qstage = reinterpret_cast<float*>(wb + 128 + kQStages * Sh::FACT_BYTES + ((NB + 1) * kWarp * 4 + 127) / 128 * 128);
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool in_al = aligned16(P.facts), out_al = aligned16(P.queries_all) && ((kWarp * Q * 4) % 16 == 0);
  WarpTileLoader<kQStages> ld;
  ld.init(fstage, bars, Sh::FACT_BYTES / 4, lane);
  auto request = [&](long long sub, int s) {
    if (sub < n_sub) ld.issue(s, P.facts + sub * kWarp * F, (int)min((long long)kWarp, P.B - sub * kWarp) * F, in_al, lane, 0.5f);
  };
#pragma unroll
  for (int s = 0; s < kQStages; ++s) request(gw + s * GW, s);
  p2s[NB * kWarp + lane] = 0.0f;   // the zero row (written once)
  const uint32_t pl = smem_u32(p2s) + lane * 4;
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % kQStages);
    const int rows = (int)min((long long)kWarp, P.B - sub * kWarp);
    ld.wait(s);
    float p[F];
#pragma unroll
    for (int c = 0; c < F / FV; ++c) lds_vec<FV>(p + c * FV, ld.stage(s) + lane * F + c * FV);
    __syncwarp();
    request(sub + kQStages * GW, s);
#pragma unroll
    for (int k = 0; k < NB; ++k) p2s[k * kWarp + lane] = p[NA + k];
    float z = 1.0f;
    if constexpr (NORM) z = ad2_normaliser<NA, NB>(p);
    if (lane == 0) bulk_wait_read<kQStages - 1>();   // the store of kQStages tiles ago has read its stage
    __syncwarp();
    float* qs = qstage + s * (qbytes / 4);
    // slots per (q, i): 1, 2 and 4 are unrolled (all gathers of a query in flight before its add chain starts)
    auto fold = [&](int t, auto Lc) {
      constexpr int LL = decltype(Lc)::value;
      float v[NA * LL];
#pragma unroll
      for (int j = 0; j < NA * LL; ++j) v[j] = q_lds(pl + P.slot[t + j]);
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < NA * LL; ++j) acc = __fadd_rn(acc, __fmul_rn(p[j / LL], v[j]));
      return acc;
    };
    for (int q = 0; q < Q; ++q) {
      const int L = P.len[q];
      int t = P.ptr[q];
      float acc = 0.0f;
      if (L == 1) acc = fold(t, std::integral_constant<int, 1>{});
      else if (L == 2) acc = fold(t, std::integral_constant<int, 2>{});
      else if (L == 4) acc = fold(t, std::integral_constant<int, 4>{});
      else {
#pragma unroll
        for (int i = 0; i < NA; ++i)
          for (int l = 0; l < L; ++l, ++t) acc = __fadd_rn(acc, __fmul_rn(p[i], q_lds(pl + P.slot[t])));
      }
      qs[lane * Q + q] = NORM ? acc / z : acc;
    }
    fence_proxy_async();
    __syncwarp();
    warp_tile_store(P.queries_all + sub * kWarp * Q, qs, rows * Q, kWarp * Q, out_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
}

// Addition-shaped queries (the MNIST-addition task itself: query s holds exactly the worlds (j, k) with j + k = s,
// recognised at vael_circuit_create): the member lists are the anti-diagonals of the NA x NB outer product, so the
// whole contraction unrolls at compile time over register operands — 100 FMUL + 100 FADD per row, no shared-memory
// world tile, no table fetches.  Each sum runs over j ascending = world order = CSR child order (bit-exact).  The
// [32, Q] result tile goes through a shared-memory tile so that the global store is coalesced.
template <int NA, int NB, bool NORM>
__global__ void __launch_bounds__(256) ad2_queries_diag_kernel(const Ad2QParams P) {
  using Sh = Ad2Shape<NA, NB>;
  constexpr int F = Sh::F, FV = Sh::FV, Q = NA + NB - 1;
  constexpr int PER_WARP = Sh::FACT_BYTES + ((kWarp * Q * 4 + 127) / 128) * 128;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* fs = reinterpret_cast<float*>(smem + (size_t)warp * PER_WARP);                      // [32][F]
  float* qs = reinterpret_cast<float*>(smem + (size_t)warp * PER_WARP + Sh::FACT_BYTES);     // [32][Q]
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long GW = (long long)gridDim.x * nwarps;
  float nx[F];
  auto fetch = [&](long long sub) {
    const long long r0 = sub * kWarp;
    const int n = (sub < n_sub) ? (int)min((long long)kWarp, P.B - r0) * F : 0;
    const float* src = P.facts + r0 * F;
#pragma unroll
    for (int k = 0; k < F; ++k) nx[k] = (k * kWarp + lane < n) ? __ldg(src + k * kWarp + lane) : 0.5f;
  };
  fetch(gw);
  for (long long sub = gw; sub < n_sub; sub += GW) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < F; ++k) fs[k * kWarp + lane] = nx[k];
    __syncwarp();
    fetch(sub + GW);
    float p[F];
#pragma unroll
    for (int c = 0; c < F / FV; ++c) lds_vec<FV>(p + c * FV, fs + lane * F + c * FV);
    float z = 1.0f;
    if constexpr (NORM) z = ad2_normaliser<NA, NB>(p);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        const int k = q - j;
        if (k >= 0 && k < NB) acc = __fadd_rn(acc, __fmul_rn(p[j], p[NA + (k >= 0 && k < NB ? k : 0)]));
      }
      qs[lane * Q + q] = NORM ? acc / z : acc;
    }
    __syncwarp();
    float* dst = P.queries_all + row0 * Q;
    for (int i = lane; i < rows * Q; i += kWarp) dst[i] = qs[i];
  }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return atoi(v);
}
int l2_hint_env() {
  static const int h = env_int("VAEL_L2_HINT", 0);
  return h;
}

// from these batches up the non-persistent grids below are used (1e5 rows: forward 12.6 us against 11.3 us persistent,
// adjoint 9.3 us against 10.9 us; 2^18 rows: 25.7 / 27.6 us against 27.5 / 32.4 us)
constexpr long long kOnceMinRows = 1 << 18, kOnceMinRowsBwd = 1 << 16;

// ONE tile per warp, CTAs of a few warps, as many CTAs as there are tiles (no persistent loop): the write front of a
// non-persistent grid advances compactly through memory, and write-dominated streams run measurably faster that way
// (scripts/micro/storebench.cu, 12.8 KB tiles + 2.5 KB loads: 6.7 TB/s against 6.1 TB/s for persistent warps striding
// over the tiles; pure stores 7.5 against 6.3)
template <int NA, int NB, bool NORM, bool EVID>
static cudaError_t launch_fwd_once(const Ad2Params& P, int nw, cudaStream_t st) {
  using Sh = Ad2Shape<NA, NB>;
  constexpr int PER_WARP = Sh::FACT_BYTES + Sh::OUT_BYTES + 128;
  cudaError_t e = cudaFuncSetAttribute(ad2_forward_kernel<NA, NB, NORM, EVID, 1, 1>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, nw * PER_WARP);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long grid = (n_sub + nw - 1) / nw;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  Ad2Params Q = P;
  Q.l2_hint = l2_hint_env();
  ad2_forward_kernel<NA, NB, NORM, EVID, 1, 1><<<(unsigned)grid, nw * 32, nw * PER_WARP, st>>>(Q);
  return cudaGetLastError();
}

template <int NA, int NB, bool NORM, bool EVID, int S, int SO, bool STG = false>
static cudaError_t launch_fwd_cfg(const Ad2Params& P, int num_sms, cudaStream_t st) {
  using Sh = Ad2Shape<NA, NB>;
  constexpr int PER_WARP = S * Sh::FACT_BYTES + SO * Sh::OUT_BYTES + 128;
  // warps per SM measured on B200 (scripts/ad2_warp_sweep.sh): the write-dominated forward wants FEW store streams
  // in flight for the 12.8 KB tiles of the 10 x 10 shape (5: 0.92 of the copy peak; 6-8: 0.87) and one more for
  // the 10.4 KB tiles of Mario's 9 x 9 (6: 0.89; 5: 0.83)
  int nw = max(1, min(8, env_int("VAEL_AD2_FWD_WARPS", Sh::W >= 100 ? 5 : 6)));
  while (nw > 1 && nw * PER_WARP > 227 * 1024) --nw;
  const int cps = max(1, min(8, env_int("VAEL_AD2_FWD_CTAS_PER_SM", (227 * 1024) / (nw * PER_WARP + 1024))));
  cudaError_t e = cudaFuncSetAttribute(ad2_forward_kernel<NA, NB, NORM, EVID, S, SO, STG>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, nw * PER_WARP);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  long long grid = (n_sub + nw - 1) / nw;
  grid = min(grid, (long long)num_sms * cps);
  Ad2Params Q = P;
  Q.l2_hint = l2_hint_env();
  ad2_forward_kernel<NA, NB, NORM, EVID, S, SO, STG><<<(unsigned)grid, nw * 32, nw * PER_WARP, st>>>(Q);
  return cudaGetLastError();
}

template <int NA, int NB, bool NORM, bool EVID>
static cudaError_t launch_fwd_t(const Ad2Params& P, int num_sms, cudaStream_t st) {
  // pipeline depth (load stages S, store stages SO) — tunable for experiments, default 2/2
  static const int cfg = env_int("VAEL_AD2_FWD_STAGES", 22);
  // streaming batches: one tile per warp, three warps per CTA (B = 2^23: 0.606 ms against 0.686 ms for the persistent
  // pipeline below; 1 / 2 / 4 / 6 / 8 warps: 0.611 / 0.610 / 0.613 / 0.640 / 0.797).  Small batches keep the
  // persistent kernel (B = 1e5: 11.3 us against 12.6).  VAEL_AD2_FWD_ONCE: 0 = never, n > 0 = always with n warps.
  static const int once_env = env_int("VAEL_AD2_FWD_ONCE", -1);
  if (once_env > 0 || (once_env < 0 && P.B >= kOnceMinRows))
    return launch_fwd_once<NA, NB, NORM, EVID>(P, once_env > 0 ? min(8, once_env) : 3, st);
  if constexpr (!NORM && !EVID) {
    if (cfg == 23) return launch_fwd_cfg<NA, NB, NORM, EVID, 2, 3>(P, num_sms, st);
    if (cfg == 33) return launch_fwd_cfg<NA, NB, NORM, EVID, 3, 3>(P, num_sms, st);
    if (cfg == 21) return launch_fwd_cfg<NA, NB, NORM, EVID, 2, 1, true>(P, num_sms, st);
    if (cfg == 31) return launch_fwd_cfg<NA, NB, NORM, EVID, 3, 1, true>(P, num_sms, st);
  }
  return launch_fwd_cfg<NA, NB, NORM, EVID, 2, 2>(P, num_sms, st);
}

template <int NA, int NB, bool NORM>
static cudaError_t launch_bwd_once(const Ad2Params& P, int nw, cudaStream_t st) {
  using Sh = Ad2Shape<NA, NB>;
  constexpr int PER_WARP = (Sh::FACT_BYTES + Sh::OUT_BYTES) + Sh::FACT_BYTES + 128;
  cudaError_t e = cudaFuncSetAttribute(ad2_backward_kernel<NA, NB, NORM, 1, 1>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, nw * PER_WARP);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long grid = (n_sub + nw - 1) / nw;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  Ad2Params Q = P;
  Q.l2_hint = l2_hint_env();
  ad2_backward_kernel<NA, NB, NORM, 1, 1><<<(unsigned)grid, nw * 32, nw * PER_WARP, st>>>(Q);
  return cudaGetLastError();
}

template <int NA, int NB, bool NORM>
static cudaError_t launch_bwd_t(const Ad2Params& P, int num_sms, cudaStream_t st) {
  // as the forward: 0.660 ms against 0.757 ms at B = 2^23 (2 / 3 / 4 / 6 / 8 warps: 0.660 / 0.660 / 0.660 / 0.661 / 0.726)
  static const int once_env = env_int("VAEL_AD2_BWD_ONCE", -1);
  if (once_env > 0 || (once_env < 0 && P.B >= kOnceMinRowsBwd))
    return launch_bwd_once<NA, NB, NORM>(P, once_env > 0 ? min(8, once_env) : 4, st);
  constexpr int S = 2, SO = 2;
  using Sh = Ad2Shape<NA, NB>;
  constexpr int PER_WARP = S * (Sh::FACT_BYTES + Sh::OUT_BYTES) + SO * Sh::FACT_BYTES + 128;
  // read-dominated: more loads in flight help (10 x 10: 6 warps 0.96; 9 x 9: 7 warps 0.88, 6 warps 0.79)
  int nw = max(1, min(8, env_int("VAEL_AD2_BWD_WARPS", Sh::W >= 100 ? 6 : 7)));
  while (nw > 1 && nw * PER_WARP > 227 * 1024) --nw;
  const int cps = max(1, min(8, env_int("VAEL_AD2_BWD_CTAS_PER_SM", (227 * 1024) / (nw * PER_WARP + 1024))));
  cudaError_t e = cudaFuncSetAttribute(ad2_backward_kernel<NA, NB, NORM, S, SO>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, nw * PER_WARP);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  long long grid = (n_sub + nw - 1) / nw;
  grid = min(grid, (long long)num_sms * cps);
  Ad2Params Q = P;
  Q.l2_hint = l2_hint_env();
  ad2_backward_kernel<NA, NB, NORM, S, SO><<<(unsigned)grid, nw * 32, nw * PER_WARP, st>>>(Q);
  return cudaGetLastError();
}

template <int NA, int NB>
static cudaError_t launch_fwd_shape(const Ad2Params& P, bool norm, bool evid, int num_sms, cudaStream_t st) {
  if (evid) return norm ? launch_fwd_t<NA, NB, true, true>(P, num_sms, st) : launch_fwd_t<NA, NB, false, true>(P, num_sms, st);
  return norm ? launch_fwd_t<NA, NB, true, false>(P, num_sms, st) : launch_fwd_t<NA, NB, false, false>(P, num_sms, st);
}

template <int NA, int NB>
static cudaError_t launch_queries(const Ad2QParams& P, bool norm, int num_sms, cudaStream_t st) {
  if (P.diagonal && P.n_queries == NA + NB - 1) {
    using Sh = Ad2Shape<NA, NB>;
    const int per_warp = Sh::FACT_BYTES + ((kWarp * (NA + NB - 1) * 4 + 127) / 128) * 128;
    const int nw = 8;
    auto kern = norm ? ad2_queries_diag_kernel<NA, NB, true> : ad2_queries_diag_kernel<NA, NB, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nw * per_warp);
    if (e != cudaSuccess) return e;
    const long long n_sub = (P.B + kWarp - 1) / kWarp;
    const long long grid = tile_grid(n_sub, nw, (long long)num_sms * 4);
    kern<<<(unsigned)grid, nw * 32, nw * per_warp, st>>>(P);
    return cudaGetLastError();
  }
  const int per_warp = queries_per_warp_bytes<NA, NB>(P.n_queries);
  int nw = 8;
  while (nw > 1 && nw * per_warp > 110 * 1024) --nw;  // two CTAs per SM
  auto kern = norm ? ad2_queries_kernel<NA, NB, true> : ad2_queries_kernel<NA, NB, false>;
  const int smem_bytes = nw * per_warp;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const int ctas_per_sm = max(1, min(8, (227 * 1024) / (smem_bytes + 1024)));
  const long long grid = tile_grid(n_sub, nw, (long long)num_sms * ctas_per_sm);
  kern<<<(unsigned)grid, nw * 32, smem_bytes, st>>>(P);
  return cudaGetLastError();
}

cudaError_t ad2_queries_all(int na, int nb, const Ad2QParams& P, bool norm, int num_sms, cudaStream_t st) {
  if (P.n_queries > 64) return cudaErrorInvalidValue;
  if (na == 10 && nb == 10) return launch_queries<10, 10>(P, norm, num_sms, st);
  if (na == 9 && nb == 9) return launch_queries<9, 9>(P, norm, num_sms, st);
  return cudaErrorInvalidValue;
}

bool ad2_supported(int na, int nb) { return (na == 10 && nb == 10) || (na == 9 && nb == 9); }

cudaError_t ad2_forward(int na, int nb, const Ad2Params& P, bool norm, bool evid, int num_sms, cudaStream_t st) {
  if (na == 10 && nb == 10) return launch_fwd_shape<10, 10>(P, norm, evid, num_sms, st);
  if (na == 9 && nb == 9) return launch_fwd_shape<9, 9>(P, norm, evid, num_sms, st);
  return cudaErrorInvalidValue;
}

cudaError_t ad2_backward(int na, int nb, const Ad2Params& P, bool norm, int num_sms, cudaStream_t st) {
  if (na == 10 && nb == 10) return norm ? launch_bwd_t<10, 10, true>(P, num_sms, st) : launch_bwd_t<10, 10, false>(P, num_sms, st);
  if (na == 9 && nb == 9) return norm ? launch_bwd_t<9, 9, true>(P, num_sms, st) : launch_bwd_t<9, 9, false>(P, num_sms, st);
  return cudaErrorInvalidValue;
}

}  // namespace vael
