// Structured path for WIDE world levels: circuits whose worlds are the outer product of
// three annotated disjunctions (3-digit MNIST addition: 10 x 10 x 10 = 1000 worlds, the
// "scaled logic program" of BASELINE.json configs[4]) and whose query nodes are sums over
// world nodes.
//
// Reference semantics: the k-digit generalisation of MNISTPairsVAELModel.problog_inference
// (models/vael.py:200-225) == query-mode sdd.evaluate (models/vael.py:98-117); evidence mode ==
// problog_inference_with_evidence (models/vael.py:119-135).  World w = (i*NA + j)*NB + k has the
// value (p1[i] * p2[j]) * p3[k] — the left-deep product the compiled CSR holds.
//
// Execution model (sm_100a): warp-autonomous pipelines like circuit_ad2.cu, but a 32-row x
// 1000-world tile is 128 KB, so the world level is swept in COLUMN CHUNKS of C = NA*NB worlds
// (one chunk per value i of the leading AD).  A 32-row x C-column chunk of the row-major [B,W]
// tensor is a strided box: it is moved by the TMA unit through a 2-D tensor map
// (cp.async.bulk.tensor.2d, SASS UTMALDG / UTMASTG), rows beyond B clipped / zero-filled by
// the hardware.  The [32,F] facts tile and the [32,F] grad_facts tile are contiguous and use
// 1-D bulk copies.  Lane <-> batch row; shared-memory accesses are conflict-free 128-bit
// (row stride C*4 = 400 B puts 8 consecutive lanes on 8 distinct 16-byte bank groups).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "params.cuh"

namespace vael {

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, const void* smem_src, int c0, int c1,
                                                  uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

template <int NP, int NA, int NB>
struct AdwShape {
  static constexpr int F = NP + NA + NB;
  static constexpr int C = NA * NB;   // worlds per chunk
  static constexpr int W = NP * C;
  static constexpr int MWC = (C + 31) / 32;  // mask words per (query, chunk)
  static constexpr int FV = (F % 4 == 0) ? 4 : ((F % 2 == 0) ? 2 : 1);
  static constexpr int CV = (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1);
  static constexpr int FACT_BYTES = ((kWarp * F * 4 + 127) / 128) * 128;
  static constexpr int CH_BYTES = kWarp * C * 4;
  static_assert(CH_BYTES % 128 == 0, "chunk tile must keep the stages 128-byte aligned");
  static_assert((C * 4) % 16 == 0, "TMA box rows are multiples of 16 bytes");
};

template <int V>
__device__ __forceinline__ void w_lds(float* dst, const float* src) {
  if constexpr (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(src);
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  } else if constexpr (V == 2) {
    const float2 t = *reinterpret_cast<const float2*>(src);
    dst[0] = t.x; dst[1] = t.y;
  } else {
    dst[0] = src[0];
  }
}
template <int V>
__device__ __forceinline__ void w_sts(float* dst, const float* src) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(src[0], src[1]);
  } else {
    dst[0] = src[0];
  }
}

// normaliser Z = ((za * zb) * zc), z_g = sum(p_g) + (1 - sum(p_g)): ProbLog's WMC of the AD constraints
template <int NP, int NA, int NB>
__device__ __forceinline__ float adw_normaliser(const float* frow, const float* pa, const float* pb) {
  float sp = frow[0];
#pragma unroll 1
  for (int i = 1; i < NP; ++i) sp = __fadd_rn(sp, frow[i]);
  float sa = pa[0];
#pragma unroll
  for (int j = 1; j < NA; ++j) sa = __fadd_rn(sa, pa[j]);
  float sb = pb[0];
#pragma unroll
  for (int k = 1; k < NB; ++k) sb = __fadd_rn(sb, pb[k]);
  const float zp = __fadd_rn(sp, __fsub_rn(1.0f, sp));
  const float za = __fadd_rn(sa, __fsub_rn(1.0f, sa));
  const float zb = __fadd_rn(sb, __fsub_rn(1.0f, sb));
  return __fmul_rn(__fmul_rn(zp, za), zb);
}

template <int MWC>
__device__ __forceinline__ void load_chunk_mask(uint32_t* m, const AdwParams& P, int q, int chunk) {
  if (q < 0) {
#pragma unroll
    for (int i = 0; i < MWC; ++i) m[i] = 0u;
    return;
  }
  const uint32_t* src = P.qmask_c + ((long long)q * P.n_chunks + chunk) * MWC;
  if constexpr (MWC == 4) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(src));
    m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < MWC; ++i) m[i] = __ldg(src + i);
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
// IC = values of the leading AD per stored box: the box is [32 rows, IC*C columns].  IC = 2 makes every
// row segment 800 B = 25 whole 32-byte sectors (a 400 B segment ends mid-sector), which is worth ~12 % of
// store bandwidth (scripts/micro/storebench.cu: 5.36 -> 5.91 TB/s at 3 warps/SM); the price is a 2-way
// bank conflict on the STS.128 (row stride 800 B), on a path that uses ~15 % of the shared-memory bandwidth.
template <int NP, int NA, int NB, bool NORM, bool EVID, int S, int SO, int IC>
__global__ void __launch_bounds__(256) adw_forward_kernel(const __grid_constant__ CUtensorMap tm_worlds,
                                                          const AdwParams P) {
  using Sh = AdwShape<NP, NA, NB>;
  constexpr int F = Sh::F, C = Sh::C, MWC = Sh::MWC, FV = Sh::FV, CV = Sh::CV;
  constexpr int BOX_BYTES = IC * Sh::CH_BYTES;
  constexpr int PER_WARP = S * Sh::FACT_BYTES + SO * BOX_BYTES + 128;
  static_assert(NP % IC == 0, "the leading AD must split into whole boxes");
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  uint8_t* wbase = smem + (size_t)warp * PER_WARP;
  uint8_t* out_base = wbase + S * Sh::FACT_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_base + SO * BOX_BYTES);

  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;

  const bool in_al = aligned16(P.facts);
  const uint64_t pol = l2_policy_evict_first();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  uint32_t phase = 0, via_tma = 0;
  auto load_facts = [&](long long sub, int s) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const uint32_t bytes = (uint32_t)rows * F * 4u;
    float* dst = reinterpret_cast<float*>(wbase + s * Sh::FACT_BYTES);
    const float* src = P.facts + row0 * F;
    if (in_al && (bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(dst, src, bytes, &full[s]);
      }
      via_tma |= 1u << s;
    } else {
      for (int i = lane; i < rows * F; i += kWarp) dst[i] = src[i];
      via_tma &= ~(1u << s);
    }
  };

#pragma unroll
  for (int s = 0; s < S; ++s) {
    const long long sub = gw + (long long)s * GW;
    if (sub < n_sub) load_facts(sub, s);
  }

  long long it = 0;
  long long nchunk = 0;  // chunks stored so far by this warp (selects the out stage)
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % S);
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const long long row = row0 + lane;
    const bool valid = lane < rows;

    int q = -1;
    if (valid) {
      q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
      if (q < 0 || q >= P.n_queries) q = -1;
    }

    if (via_tma & (1u << s)) {
      mbar_wait(&full[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    } else {
      __syncwarp();
    }

    const float* frow = reinterpret_cast<const float*>(wbase + s * Sh::FACT_BYTES) + lane * F;
    float pa[NA], pb[NB];
    {
      float tmp[F];  // only the [NP, F) part stays live
#pragma unroll
      for (int c = 0; c < F / FV; ++c) w_lds<FV>(tmp + c * FV, frow + c * FV);
#pragma unroll
      for (int j = 0; j < NA; ++j) pa[j] = tmp[NP + j];
#pragma unroll
      for (int k = 0; k < NB; ++k) pb[k] = tmp[NP + NA + k];
    }
    float z = 1.0f;
    if constexpr (NORM) z = adw_normaliser<NP, NA, NB>(frow, pa, pb);

    float acc = 0.0f;  // value of the selected query node: sequential fp32 sum in world order
    if constexpr (EVID) {
#pragma unroll 1
      for (int i = 0; i < NP; ++i) {
        uint32_t m[MWC];
        load_chunk_mask<MWC>(m, P, q, i);
        const float pre = frow[i];
#pragma unroll
        for (int w = 0; w < C; ++w) {
          const float v = __fmul_rn(__fmul_rn(pre, pa[w / NB]), pb[w % NB]);
          if ((m[w >> 5] >> (w & 31)) & 1u) acc = __fadd_rn(acc, v);
        }
      }
    }

#pragma unroll 1
    for (int i = 0; i < NP; i += IC, ++nchunk) {
      const int so = (int)(nchunk % SO);
      float* orow = reinterpret_cast<float*>(out_base + so * BOX_BYTES) + lane * (IC * C);
      if (P.worlds != nullptr) {
        // the tensor store issued SO boxes ago must have finished reading this stage
        if (lane == 0) bulk_wait_read<SO - 1>();
        __syncwarp();
      }
#pragma unroll
      for (int ii = 0; ii < IC; ++ii) {
        uint32_t m[MWC];
        load_chunk_mask<MWC>(m, P, q, i + ii);
        const float pre = frow[i + ii];
        float t[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) t[j] = __fmul_rn(pre, pa[j]);
#pragma unroll
        for (int c = 0; c < C / CV; ++c) {
          float o[CV];
#pragma unroll
          for (int u = 0; u < CV; ++u) {
            const int w = c * CV + u;
            float v = __fmul_rn(t[w / NB], pb[w % NB]);
            const bool member = (m[w >> 5] >> (w & 31)) & 1u;
            if constexpr (EVID) {
              v = member ? v / acc : 0.0f;  // P(w | e) = [w |= e] P(w) / P(e)
            } else {
              if (member) acc = __fadd_rn(acc, v);
              if constexpr (NORM) v = v / z;
            }
            o[u] = v;
          }
          if (P.worlds != nullptr) w_sts<CV>(orow + ii * C + c * CV, o);
        }
      }
      if (P.worlds != nullptr) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (P.l2_hint >= 1) tma_store_2d_hint(&tm_worlds, out_base + so * BOX_BYTES, i * C, (int)row0, pol);
          else tma_store_2d(&tm_worlds, out_base + so * BOX_BYTES, i * C, (int)row0);
          bulk_commit();
        }
      }
    }

    if (P.query != nullptr && valid) {
      float qv = acc;
      if constexpr (NORM) qv = acc / z;
      P.query[row] = qv;
    }

    // every lane is done with this facts stage: refill it with the tile S iterations ahead
    __syncwarp();
    {
      const long long nxt = sub + (long long)S * GW;
      if (nxt < n_sub) load_facts(nxt, s);
    }
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}

// ---------------------------------------------------------------------------
// backward (adjoint); node values are recomputed from the facts
// ---------------------------------------------------------------------------
template <int NP, int NA, int NB, bool NORM, int S>
__global__ void __launch_bounds__(256) adw_backward_kernel(const __grid_constant__ CUtensorMap tm_gw,
                                                           const AdwParams P) {
  using Sh = AdwShape<NP, NA, NB>;
  constexpr int F = Sh::F, C = Sh::C, MWC = Sh::MWC, FV = Sh::FV, CV = Sh::CV;
  constexpr int SF = 2, SG = 2;  // facts ring, grad_facts ring
  constexpr int PER_WARP = S * Sh::CH_BYTES + (SF + SG) * Sh::FACT_BYTES + 128;
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  uint8_t* wbase = smem + (size_t)warp * PER_WARP;
  uint8_t* fact_base = wbase + S * Sh::CH_BYTES;
  uint8_t* gout_base = fact_base + SF * Sh::FACT_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(gout_base + SG * Sh::FACT_BYTES);  // [S] chunk stages
  uint64_t* ffull = full + S;                                                      // [SF] facts stages

  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;

  const bool has_gw = P.grad_worlds != nullptr;
  const bool in_al = aligned16(P.facts);
  const bool out_al = aligned16(P.grad_facts);
  const uint64_t pol = l2_policy_evict_first();

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
#pragma unroll
    for (int s = 0; s < SF; ++s) mbar_init(&ffull[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  uint32_t fphase = 0, fvia = 0;
  auto load_facts = [&](long long sub, int fs) {
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const uint32_t bytes = (uint32_t)rows * F * 4u;
    float* dst = reinterpret_cast<float*>(fact_base + fs * Sh::FACT_BYTES);
    const float* src = P.facts + row0 * F;
    if (in_al && (bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&ffull[fs], bytes);
        bulk_g2s(dst, src, bytes, &ffull[fs]);
      }
      fvia |= 1u << fs;
    } else {
      for (int i = lane; i < rows * F; i += kWarp) dst[i] = src[i];
      fvia &= ~(1u << fs);
    }
  };
  // chunk n of this warp's sequence = (tile n / NP, leading-AD value n % NP)
  auto issue_chunk = [&](long long n) {
    const long long tt = n / NP;
    const int ii = (int)(n - tt * NP);
    const long long sub = gw + tt * GW;
    if (sub >= n_sub) return;
    if (ii == 0) load_facts(sub, (int)(tt % SF));
    if (has_gw && lane == 0) {
      const int s = (int)(n % S);
      mbar_arrive_expect_tx(&full[s], (uint32_t)Sh::CH_BYTES);  // the box is always delivered whole (OOB rows = 0)
      if (P.l2_hint >= 2) tma_load_2d_hint(wbase + s * Sh::CH_BYTES, &tm_gw, ii * C, (int)(sub * kWarp), &full[s], pol);
      else tma_load_2d(wbase + s * Sh::CH_BYTES, &tm_gw, ii * C, (int)(sub * kWarp), &full[s]);
    }
  };

#pragma unroll 1
  for (int n = 0; n < S; ++n) issue_chunk(n);

  long long n = 0, tt = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++tt) {
    const int fs = (int)(tt % SF), gs = (int)(tt % SG);
    const long long row0 = sub * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    const long long row = row0 + lane;
    const bool valid = lane < rows;

    int q = -1;
    float gq = 0.0f;
    if (valid && P.grad_query != nullptr) {
      q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
      if (q < 0 || q >= P.n_queries) q = -1;
      gq = __ldg(P.grad_query + row);
    }

    if (fvia & (1u << fs)) {
      mbar_wait(&ffull[fs], (fphase >> fs) & 1u);
      fphase ^= 1u << fs;
    } else {
      __syncwarp();
    }
    const float* frow = reinterpret_cast<const float*>(fact_base + fs * Sh::FACT_BYTES) + lane * F;
    float pa[NA], pb[NB];
    {
      float tmp[F];
#pragma unroll
      for (int c = 0; c < F / FV; ++c) w_lds<FV>(tmp + c * FV, frow + c * FV);
#pragma unroll
      for (int j = 0; j < NA; ++j) pa[j] = tmp[NP + j];
#pragma unroll
      for (int k = 0; k < NB; ++k) pb[k] = tmp[NP + NA + k];
    }
    float z = 1.0f;
    if constexpr (NORM) z = adw_normaliser<NP, NA, NB>(frow, pa, pb);

    // the bulk store of grad_facts issued SG tiles ago must have finished reading this stage
    if (lane == 0) bulk_wait_read<SG - 1>();
    __syncwarp();
    float* grow_out = reinterpret_cast<float*>(gout_base + gs * Sh::FACT_BYTES) + lane * F;

    float ga[NA], gb[NB];
#pragma unroll
    for (int j = 0; j < NA; ++j) ga[j] = 0.0f;
#pragma unroll
    for (int k = 0; k < NB; ++k) gb[k] = 0.0f;

#pragma unroll 1
    for (int i = 0; i < NP; ++i, ++n) {
      const int s = (int)(n % S);
      uint32_t m[MWC];
      load_chunk_mask<MWC>(m, P, q, i);
      const float pre = frow[i];
      float u[NB], v[NA];
#pragma unroll
      for (int k = 0; k < NB; ++k) u[k] = pre * pb[k];
#pragma unroll
      for (int j = 0; j < NA; ++j) v[j] = pre * pa[j];
      if (has_gw) mbar_wait(&full[s], (uint32_t)((n / S) & 1));
      const float* grow = reinterpret_cast<const float*>(wbase + s * Sh::CH_BYTES) + lane * C;
      float si = 0.0f;
#pragma unroll
      for (int c = 0; c < C / CV; ++c) {
        float a[CV];
        if (has_gw) {
          w_lds<CV>(a, grow + c * CV);
        } else {
#pragma unroll
          for (int x = 0; x < CV; ++x) a[x] = 0.0f;
        }
#pragma unroll
        for (int x = 0; x < CV; ++x) {
          const int w = c * CV + x;
          float av = a[x];
          if ((m[w >> 5] >> (w & 31)) & 1u) av += gq;
          if constexpr (NORM) av = av / z;  // Z's own adjoint is exactly 0
          si = fmaf(av, pa[w / NB] * pb[w % NB], si);
          ga[w / NB] = fmaf(av, u[w % NB], ga[w / NB]);
          gb[w % NB] = fmaf(av, v[w / NB], gb[w % NB]);
        }
      }
      grow_out[i] = si;
      __syncwarp();          // every lane has read chunk stage s
      issue_chunk(n + S);    // refill it (and, on a tile boundary, start the next facts tile)
    }
#pragma unroll
    for (int j = 0; j < NA; ++j) grow_out[NP + j] = ga[j];
#pragma unroll
    for (int k = 0; k < NB; ++k) grow_out[NP + NA + k] = gb[k];

    {
      const uint32_t bytes = (uint32_t)rows * F * 4u;
      float* dst = P.grad_facts + row0 * F;
      const float* src = reinterpret_cast<const float*>(gout_base + gs * Sh::FACT_BYTES);
      if (out_al && (bytes & 15u) == 0) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          bulk_s2g(dst, src, bytes);
          bulk_commit();
        }
      } else {
        __syncwarp();
        for (int i = lane; i < rows * F; i += kWarp) dst[i] = src[i];
        if (lane == 0) bulk_commit();  // keep one group per tile so that wait_group.read<SG-1> counts tiles
      }
    }
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tmap_encode_fn>(p);
  }
  return fn;
}

// [B,W] fp32 row-major tensor, box = [32 rows, C columns]
static bool make_map(CUtensorMap* map, const float* base, long long B, int W, int C) {
  tmap_encode_fn enc = tmap_encoder();
  if (!enc || !aligned16(base)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)B};
  const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)kWarp};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the same map for the generated column-chunked kernels (circuit_jit.cu keeps the driver types out of its sources):
// `map128` points at 128 bytes, 64-byte aligned
bool make_row_box_map(void* map128, const float* base, long long B, int W, int C) {
  return make_map(static_cast<CUtensorMap*>(map128), base, B, W, C);
}

static int adw_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Non-persistent grids (one tile per warp, as many CTAs as tiles) stream faster than persistent warps striding over
// the tiles (circuit_ad2.cu, scripts/micro/storebench.cu).  Measured here (3-digit circuit, 524 288 rows): forward
// 5.81 TB/s against 5.61 persistent; the adjoint, with its three-stage chunk ring, runs 6.38 persistent against 5.92
// non-persistent and stays persistent.  VAEL_ADW_FWD_ONCE / VAEL_ADW_BWD_ONCE override (0 / 1).
static bool adw_fwd_once() {
  static const bool v = adw_env("VAEL_ADW_FWD_ONCE", 1) != 0;
  return v;
}
static bool adw_bwd_once() {
  static const bool v = adw_env("VAEL_ADW_BWD_ONCE", 0) != 0;
  return v;
}

template <int NP, int NA, int NB, bool NORM, bool EVID, int S, int SO, int IC>
static cudaError_t launch_fwd_cfg(const AdwParams& P0, int num_sms, cudaStream_t st) {
  using Sh = AdwShape<NP, NA, NB>;
  constexpr int PER_WARP = S * Sh::FACT_BYTES + SO * IC * Sh::CH_BYTES + 128;
  AdwParams P = P0;
  P.l2_hint = l2_hint_env();
  alignas(64) CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (P.worlds != nullptr && !make_map(&map, P.worlds, P.B, Sh::W, IC * Sh::C)) return cudaErrorNotSupported;
  int nw = max(1, min(8, adw_env("VAEL_ADW_FWD_WARPS", 2)));  // fewer, longer store streams measure faster
  while (nw > 1 && nw * PER_WARP > 227 * 1024) --nw;
  auto kern = adw_forward_kernel<NP, NA, NB, NORM, EVID, S, SO, IC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nw * PER_WARP);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  long long grid = (n_sub + nw - 1) / nw;
  if (!adw_fwd_once()) grid = min(grid, (long long)num_sms);
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  kern<<<(unsigned)grid, nw * 32, nw * PER_WARP, st>>>(map, P);
  return cudaGetLastError();
}

template <int NP, int NA, int NB, bool NORM, bool EVID>
static cudaError_t launch_fwd(const AdwParams& P, int num_sms, cudaStream_t st) {
  // <load stages><store stages><leading-AD values per box>; measured on B200 (B = 2^20): 231 with 2 warps/SM
  // 5.68 TB/s, 232 5.63, 222 5.60, 231 with 3 / 4 warps 5.19 / 5.00 (more store streams in flight run slower)
  static const int cfg = adw_env("VAEL_ADW_FWD_CFG", 231);
  if constexpr (!NORM && !EVID) {
    if (cfg == 222) return launch_fwd_cfg<NP, NA, NB, NORM, EVID, 2, 2, 2>(P, num_sms, st);
    if (cfg == 221) return launch_fwd_cfg<NP, NA, NB, NORM, EVID, 2, 2, 1>(P, num_sms, st);
    if (cfg == 232) return launch_fwd_cfg<NP, NA, NB, NORM, EVID, 2, 3, 2>(P, num_sms, st);
    if (cfg == 241) return launch_fwd_cfg<NP, NA, NB, NORM, EVID, 2, 4, 1>(P, num_sms, st);
  }
  return launch_fwd_cfg<NP, NA, NB, NORM, EVID, 2, 3, 1>(P, num_sms, st);
}

template <int NP, int NA, int NB, bool NORM, int S>
static cudaError_t launch_bwd_cfg(const AdwParams& P0, int num_sms, cudaStream_t st) {
  using Sh = AdwShape<NP, NA, NB>;
  constexpr int PER_WARP = S * Sh::CH_BYTES + 4 * Sh::FACT_BYTES + 128;
  AdwParams P = P0;
  P.l2_hint = l2_hint_env();
  alignas(64) CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (P.grad_worlds != nullptr && !make_map(&map, P.grad_worlds, P.B, Sh::W, Sh::C)) return cudaErrorNotSupported;
  int nw = max(1, min(8, adw_env("VAEL_ADW_BWD_WARPS", adw_bwd_once() ? 2 : (227 * 1024 - 1024) / PER_WARP)));
  while (nw > 1 && nw * PER_WARP > 227 * 1024) --nw;
  auto kern = adw_backward_kernel<NP, NA, NB, NORM, S>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nw * PER_WARP);
  if (e != cudaSuccess) return e;
  const long long n_sub = (P.B + kWarp - 1) / kWarp;
  long long grid = (n_sub + nw - 1) / nw;
  if (!adw_bwd_once()) grid = min(grid, (long long)num_sms);
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  kern<<<(unsigned)grid, nw * 32, nw * PER_WARP, st>>>(map, P);
  return cudaGetLastError();
}

template <int NP, int NA, int NB, bool NORM>
static cudaError_t launch_bwd(const AdwParams& P, int num_sms, cudaStream_t st) {
  static const int cfg = adw_env("VAEL_ADW_BWD_STAGES", 3);
  if constexpr (!NORM) {
    if (cfg == 2) return launch_bwd_cfg<NP, NA, NB, NORM, 2>(P, num_sms, st);
    if (cfg == 4) return launch_bwd_cfg<NP, NA, NB, NORM, 4>(P, num_sms, st);
  }
  return launch_bwd_cfg<NP, NA, NB, NORM, 3>(P, num_sms, st);
}

// ---------------------------------------------------------------------------
// all query nodes at once for addition-shaped queries over three annotated disjunctions (3-digit addition:
// query s holds exactly the worlds (i, j, k) with i + j + k = s; recognised at vael_circuit_create).  Same idea as
// ad2_queries_diag_kernel: the whole contraction unrolls at compile time over register operands — world
// (i, j, k) = (p1[i] * p2[j]) * p3[k], the left-deep product the CSR holds, each sum over (i, j) ascending = world
// order = CSR child order (bit-exact); nothing but facts[B,30] in and queries[B,28] out touches memory.
// ---------------------------------------------------------------------------
template <int NP, int NA, int NB, bool NORM>
__global__ void __launch_bounds__(256) adw_queries_diag_kernel(const float* __restrict__ facts, long long B,
                                                               float* __restrict__ queries_all) {
  constexpr int F = NP + NA + NB, Q = NP + NA + NB - 2;
  constexpr int PER_WARP = ((kWarp * F * 4 + 127) / 128) * 128 + ((kWarp * Q * 4 + 127) / 128) * 128;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* fs = reinterpret_cast<float*>(smem + (size_t)warp * PER_WARP);                                        // [32][F]
  float* qs = reinterpret_cast<float*>(smem + (size_t)warp * PER_WARP + ((kWarp * F * 4 + 127) / 128) * 128);   // [32][Q]
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long GW = (long long)gridDim.x * nwarps;
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
    float p[F];
#pragma unroll
    for (int c = 0; c < F; c += 2) {   // F * 4 bytes per row: 8-byte aligned rows
      const float2 t = *reinterpret_cast<const float2*>(fs + lane * F + c);
      p[c] = t.x; p[c + 1] = t.y;
    }
    float z = 1.0f;
    if constexpr (NORM) z = adw_normaliser<NP, NA, NB>(p, p + NP, p + NP + NA);
    // Two-stage contraction (round 2): bc[t] = sum_{j+k=t} p2[j] p3[k] (NA NB FMAs), then query s =
    // sum_i p1[i] bc[s-i] (<= NP FMAs each) — 100 + 190 FMAs per row instead of the 100 + 2 x 1000 of folding every
    // world product (p1[i] p2[j]) p3[k] in world order.  Same terms, other association and order: agrees with the
    // CSR fold (the generic kernel, the oracle) to a few ulp; the test bar is 1e-6 relative, the north star's 1e-5.
    float bc[NA + NB - 1];
#pragma unroll
    for (int t = 0; t < NA + NB - 1; ++t) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        const int k = t - j;
        if (k >= 0 && k < NB) acc = fmaf(p[NP + j], p[NP + NA + (k >= 0 && k < NB ? k : 0)], acc);
      }
      bc[t] = acc;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float acc = 0.0f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int t = q - i;
        if (t >= 0 && t < NA + NB - 1) acc = fmaf(p[i], bc[(t >= 0 && t < NA + NB - 1) ? t : 0], acc);
      }
      qs[lane * Q + q] = NORM ? acc / z : acc;
    }
    __syncwarp();
    float* dst = queries_all + row0 * Q;
    for (int i = lane; i < rows * Q; i += kWarp) dst[i] = qs[i];
  }
}

cudaError_t adw_queries_diag(int np, int na, int nb, const float* facts, long long B, float* queries_all, bool norm,
                             int num_sms, cudaStream_t st) {
  if (!(np == 10 && na == 10 && nb == 10)) return cudaErrorInvalidValue;
  constexpr int F = 30, Q = 28;
  constexpr int PER_WARP = ((kWarp * F * 4 + 127) / 128) * 128 + ((kWarp * Q * 4 + 127) / 128) * 128;
  const int nw = 4;
  auto kern = norm ? adw_queries_diag_kernel<10, 10, 10, true> : adw_queries_diag_kernel<10, 10, 10, false>;
  const long long n_sub = (B + kWarp - 1) / kWarp;
  const long long grid = tile_grid(n_sub, nw, (long long)num_sms * 4);
  kern<<<(unsigned)grid, nw * 32, nw * PER_WARP, st>>>(facts, B, queries_all);
  return cudaGetLastError();
}

bool adw_supported(int np, int na, int nb) { return np == 10 && na == 10 && nb == 10; }

// can this call go through the tensor-map path?  (16-byte aligned [B,W] base, driver entry point present)
bool adw_usable(const float* wide_tensor) { return tmap_encoder() != nullptr && (wide_tensor == nullptr || aligned16(wide_tensor)); }

cudaError_t adw_forward(int np, int na, int nb, const AdwParams& P, bool norm, bool evid, int num_sms,
                        cudaStream_t st) {
  if (!adw_supported(np, na, nb)) return cudaErrorInvalidValue;
  if (evid) return norm ? launch_fwd<10, 10, 10, true, true>(P, num_sms, st) : launch_fwd<10, 10, 10, false, true>(P, num_sms, st);
  return norm ? launch_fwd<10, 10, 10, true, false>(P, num_sms, st) : launch_fwd<10, 10, 10, false, false>(P, num_sms, st);
}

cudaError_t adw_backward(int np, int na, int nb, const AdwParams& P, bool norm, int num_sms, cudaStream_t st) {
  if (!adw_supported(np, na, nb)) return cudaErrorInvalidValue;
  return norm ? launch_bwd<10, 10, 10, true>(P, num_sms, st) : launch_bwd<10, 10, 10, false>(P, num_sms, st);
}

}  // namespace vael
