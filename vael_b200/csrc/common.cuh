// Shared device helpers for the vael_b200 kernels (sm_100a).
// PTX wrappers for the bulk-copy engine (TMA 1-D bulk copies, SASS UBLKCP),
// mbarriers and the generic<->async proxy fence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>

#include <mutex>
#include <vector>

#include "../../include/vael_b200.h"

namespace vael {

constexpr int kWarp = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier --------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier initialisation visible to the async proxy (bulk-copy engine)
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- bulk copies (cp.async.bulk, 1-D "TMA") --------------------------------
// global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// same copies carrying an L2 eviction-priority hint (streaming data: evict_first)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// order this thread's generic-proxy shared-memory writes before later async-proxy reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- warp-private tile pipeline ----------------------------------------------
// A ring of S shared-memory stages owned by ONE warp and fed by 1-D bulk copies: a [32, K] tile of a row-major
// [B, K] tensor is one contiguous span of global memory, so a tile is a single cp.async.bulk (SASS UBLKCP) issued by
// lane 0 and completing on the stage's mbarrier — no LSU traffic, no per-lane address arithmetic.  Tails
// (fewer than 32 rows with a byte count that is not a multiple of 16) and unaligned bases take a plain coalesced
// copy by the whole warp.  All methods are called by the whole warp (converged).
template <int S>
struct WarpTileLoader {
  float* stages;        // S stages of `stage_floats` floats, 128-byte aligned
  uint64_t* bars;       // S mbarriers
  int stage_floats;
  uint32_t phase;       // bit s: parity the next wait on stage s uses
  uint32_t via_bulk;    // bit s: the in-flight load of stage s went through the bulk engine
  __device__ __forceinline__ void init(float* st, uint64_t* b, int floats_per_stage, int lane) {
    stages = st; bars = b; stage_floats = floats_per_stage; phase = 0u; via_bulk = 0u;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ float* stage(int s) const { return stages + s * stage_floats; }
  // request `n` floats from `src` into stage s (n <= stage_floats); `fill` pads the stage for tail tiles
  __device__ __forceinline__ void issue(int s, const float* src, int n, bool base_aligned, int lane, float fill) {
    float* dst = stage(s);
    const uint32_t bytes = (uint32_t)n * 4u;
    if (base_aligned && (bytes & 15u) == 0u && n > 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[s], bytes);
        bulk_g2s(dst, src, bytes, &bars[s]);
      }
      if (n < stage_floats)
        for (int i = n + lane; i < stage_floats; i += 32) dst[i] = fill;
      via_bulk |= 1u << s;
    } else {
      for (int i = lane; i < stage_floats; i += 32) dst[i] = (i < n) ? __ldg(src + i) : fill;
      via_bulk &= ~(1u << s);
    }
  }
  __device__ __forceinline__ void wait(int s) {
    if (via_bulk & (1u << s)) {
      mbar_wait(&bars[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    }
    __syncwarp();
  }
};

// Row-major [32, K] result tile in shared memory -> global: one bulk store when the tile is whole and aligned
// (the caller has fenced its generic-proxy writes: fence_proxy_async() by every lane + __syncwarp()), else a plain
// coalesced copy.  Lane 0 commits a bulk group; bulk_wait_read<N>() before the stage is rewritten.
__device__ __forceinline__ void warp_tile_store(float* gdst, const float* stage, int n, int n_full, bool base_aligned,
                                                int lane) {
  if (base_aligned && n == n_full) {
    if (lane == 0) {
      bulk_s2g(gdst, stage, (uint32_t)n * 4u);
      bulk_commit();
    }
  } else {
    for (int i = lane; i < n; i += 32) gdst[i] = stage[i];
    if (lane == 0) bulk_commit();   // keep the group count in step with the bulk path
  }
}

// Grid of a warp-tile kernel (warp w of the grid takes tiles w, w + GW, ...).  The kernels that move LARGE tiles (the
// [32, W] world tiles of circuit_ad2.cu / circuit_adw.cu / circuit_jit.cu) launch non-persistent grids — one tile per
// warp, as many CTAs as tiles: 10-15 % faster DRAM streams (scripts/micro/storebench.cu).  The kernels built on
// WarpTileLoader move small tiles and / or pay a per-CTA prologue (tables staged in shared memory), and measured the
// other way (scripts/gpu_r3t.sh: literal-table kernels 0.070 / 0.095 ms persistent against 0.143 / 0.159 ms, lane <->
// row Gumbel 0.229 / 0.264 against 0.269 / 0.379, facts head and all-queries kernels within 5 % either way): they keep
// the resident-CTAs-only grid.  VAEL_ONCE_TILES=1 flips them (A-B timing).
inline long long tile_grid(long long n_sub, int warps_per_cta, long long resident_ctas) {
  static const bool once = [] { const char* e = getenv("VAEL_ONCE_TILES"); return e != nullptr && e[0] == '1'; }();
  long long grid = (n_sub + warps_per_cta - 1) / warps_per_cta;
  if (!once) grid = grid < resident_ctas ? grid : resident_ctas;
  if (grid < 1) grid = 1;
  return grid > 0x7fffffffLL ? 0x7fffffffLL : grid;
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct GenProgram;
struct JitProgram;

}  // namespace vael

constexpr int kHostSlots = 4;  // stages of the host-buffer pipeline (one stream each)

// Opaque handle behind vael_circuit_t (include/vael_b200.h)
struct vael_host_slot {  // one stage of the host-buffer (e2e) pipeline
  cudaStream_t stream;       // kernels and device->host copies
  cudaStream_t up;           // host->device copies: never queued behind a download of the same chunk
  cudaEvent_t facts_up, grads_up, done;   // uploads landed (recorded on `up`), slot free again (on `stream`)
  float *facts, *worlds, *query, *gw, *gq, *gf;
  int32_t* qidx;
};

struct vael_circuit {
  int device;
  int num_sms;
  int32_t n_facts, n_nodes, n_levels, n_edges, n_worlds, n_queries, semiring, norm_node;
  uint64_t hash;
  vael::GenProgram* gen;     // schedule of the generic level-by-level kernels (circuit_generic.cu)
  uint32_t* d_qmask;         // [n_queries][mask_words] world membership of each query (SUM over world nodes), or null
  int32_t mask_words;
  // structured fast path
  int fast_kind;             // 0 none, 2 / 3 = worlds are the outer product of two / three annotated disjunctions,
                             // 4 = literal table in the log semiring (every world: one literal per fact column)
  int32_t ad_np, ad_na, ad_nb;  // AD sizes (np = leading AD, only for fast_kind 3); fact columns in that order
  bool norm_fast_ok;         // the normaliser node has the per-AD shape the fast paths recompute
  int32_t* d_qm_ptr;         // [n_queries+1] member worlds of every query node (CSR, world order) — set with d_qmask
  int32_t* d_qm_idx;
  int32_t qm_len;
  bool q_members_ok;         // fast_kind 2: the member lists below fit the all-queries kernel's parameter block
  uint16_t q_ptr[64];        // first slot of query q
  uint8_t q_len[64];         // slots per (q, i)
  uint16_t q_slot[1024];     // byte offset (k * 128) of the second literal's row, or NB * 128 (zero row)
  int32_t q_nslots;
  uint32_t lt_mask[64];      // fast_kind 4: per world, the fact columns that enter as positive literals
  bool queries_diagonal;     // fast_kind 2: query s = the worlds (j, k) with j + k = s (all-queries register kernel)
  uint32_t* d_qmask_c;       // fast_kind 3: [n_queries][np][ceil(na*nb/32)] chunked membership masks
  // circuit-specialised straight-line kernels (circuit_jit.cu), compiled with NVRTC on first use; the description is
  // kept for that (a few kB)
  mutable std::once_flag jit_once;
  mutable vael::JitProgram* jit;
  std::vector<uint8_t> h_node_kind;
  std::vector<int32_t> h_node_arg, h_level_ptr, h_child_ptr, h_child_idx, h_world_node, h_query_node;
  std::vector<float> h_const_val;
  int32_t n_consts;
  // VAEL_FLAG_CHECK_QUERY: device flag word + its pinned host mirror, one check at a time
  int32_t* d_bad_query;
  int32_t* h_bad_query;
  mutable std::mutex check_mutex;
  // host-buffer pipeline: the only mutable part of a circuit handle.  Allocated lazily by the *_host entry
  // points under host_mutex, which also serialises concurrent host callers of one handle (they share the
  // slot buffers and streams).
  mutable std::mutex host_mutex;
  mutable vael_host_slot slot[kHostSlots];
  mutable long long chunk_rows;
};
