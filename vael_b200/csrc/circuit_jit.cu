// Circuit-specialised kernels: the host compile step goes one stage further than the CSR arrays.
//
// The generic kernels of circuit_generic.cu INTERPRET the level-ordered CSR: node values live in shared memory,
// every edge of the circuit is a shared-memory access per batch row (1.9 KB per row forward, 4.7 KB backward for
// the 2-digit circuit against 0.5 KB of HBM traffic), so their roofline is the shared-memory pipe, not HBM
// (DESIGN.md §3.3).  Here the same CSR is turned, once per circuit, into STRAIGHT-LINE sm_100a code: lane <-> batch
// row, every node a register, the level-by-level sweep unrolled at code-generation time, the reverse-topological
// adjoint sweep likewise (scatter into the children's accumulators, values recomputed from the facts) — i.e. what
// circuit_ad2.cu is by hand for one circuit shape, emitted for ANY circuit.  The generated source is compiled with
// NVRTC for sm_100a and loaded through the driver API; tiles move exactly as in the hand-written kernels (warp-private
// rings of 1-D bulk copies, cp.async.bulk / SASS UBLKCP, completing on mbarriers).
//
// Reference semantics are the interpreter's (utils/graph_semiring.py:13-60 elementwise, models/vael.py:98-135): the
// same __fmul_rn / __fadd_rn folds in CSR child order, so forward results are bit-identical to the interpreter and
// the fp32 oracle; NORMALIZE / EVIDENCE / all-queries variants as there.
//
// Scope: probability semiring, circuits whose [32, W] world tile fits one shared-memory stage (W <= 128 — every
// circuit VAEL ships except the 3-digit scale-up, which has its own structured kernels), at most 2048 nodes.
// Anything else, or a box without libnvrtc / libcuda, runs the interpreter: the JIT is an accelerator, never a
// requirement (VAEL_JIT=0 switches it off).
#include <dlfcn.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "params.cuh"

namespace vael {

// ---- dynamically loaded NVRTC + driver entry points ---------------------------------------------------------------
namespace {
typedef struct _nvrtcProgram* nvrtcProgram;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUstream_st* CUstream;
struct JitApi {
  bool ok = false;        // NVRTC and the driver: kernels can be built and launched
  bool rtc_ok = false;    // NVRTC alone: sources can be compiled (no GPU needed)
  std::string why;
  int (*nvrtcCreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  int (*nvrtcCompileProgram)(nvrtcProgram, int, const char* const*);
  int (*nvrtcGetCUBINSize)(nvrtcProgram, size_t*);
  int (*nvrtcGetCUBIN)(nvrtcProgram, char*);
  int (*nvrtcGetProgramLogSize)(nvrtcProgram, size_t*);
  int (*nvrtcGetProgramLog)(nvrtcProgram, char*);
  int (*nvrtcDestroyProgram)(nvrtcProgram*);
  int (*nvrtcVersion)(int*, int*) = nullptr;   // optional: part of the cubin cache key
  int (*cuModuleLoadData)(CUmodule*, const void*);
  int (*cuModuleUnload)(CUmodule);
  int (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*);
  int (*cuFuncSetAttribute)(CUfunction, int, int);
  int (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**,
                        void**);
};

const JitApi& jit_api() {
  static JitApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* off = getenv("VAEL_JIT");
    if (off && off[0] == '0') { api.why = "VAEL_JIT=0"; return; }
    void* rtc = nullptr;
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"})
      if ((rtc = dlopen(n, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
    void* drv = nullptr;
    for (const char* n : {"libcuda.so.1", "libcuda.so"})
      if ((drv = dlopen(n, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
    if (!rtc) { api.why = "libnvrtc not found"; return; }
    bool all = true;
    auto sym = [&](void* lib, const char* name, auto& fn) {
      void* p = lib ? dlsym(lib, name) : nullptr;
      if (!p) all = false;
      fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
    };
    sym(rtc, "nvrtcCreateProgram", api.nvrtcCreateProgram);
    sym(rtc, "nvrtcCompileProgram", api.nvrtcCompileProgram);
    sym(rtc, "nvrtcGetCUBINSize", api.nvrtcGetCUBINSize);
    sym(rtc, "nvrtcGetCUBIN", api.nvrtcGetCUBIN);
    sym(rtc, "nvrtcGetProgramLogSize", api.nvrtcGetProgramLogSize);
    sym(rtc, "nvrtcGetProgramLog", api.nvrtcGetProgramLog);
    sym(rtc, "nvrtcDestroyProgram", api.nvrtcDestroyProgram);
    api.rtc_ok = all;
    api.nvrtcVersion = reinterpret_cast<int (*)(int*, int*)>(dlsym(rtc, "nvrtcVersion"));
    if (!drv) { api.why = "libcuda not found"; return; }
    sym(drv, "cuModuleLoadData", api.cuModuleLoadData);
    sym(drv, "cuModuleUnload", api.cuModuleUnload);
    sym(drv, "cuModuleGetFunction", api.cuModuleGetFunction);
    sym(drv, "cuFuncSetAttribute", api.cuFuncSetAttribute);
    sym(drv, "cuLaunchKernel", api.cuLaunchKernel);
    api.ok = all;
    if (!all) api.why = "missing NVRTC / driver symbol";
  });
  return api;
}

// ---- the part of the generated source that does not depend on the circuit -------------------------------------------
const char* kPrelude = R"PRELUDE(
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
struct JitParams {
  const float* facts; const int* qidx; int qscalar; int pad0; long long B; const unsigned* qmask;
  float* worlds; float* query; float* queries_all;
  const float* grad_worlds; const float* grad_query; float* grad_facts;
};
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool aligned16(const void* p) { return (((uint64_t)p) & 15ull) == 0; }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// 2-D tensor-map copies of a [32 rows, WC columns] box of a row-major [B,W] tensor (SASS UTMASTG / UTMALDG): rows
// beyond B are clipped on the way out and zero-filled on the way in
struct __align__(64) TensorMap { unsigned long long opaque[16]; };
__device__ __forceinline__ void tma_store_2d(const TensorMap* tm, const void* src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_u32(src)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const TensorMap* tm, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// warp-private ring of NST shared-memory stages fed by 1-D bulk copies (one [32, K] tile = one contiguous span)
template <int NST>
struct TileRing {
  float* stages; uint64_t* bars; int stage_floats; uint32_t phase, via_bulk;
  __device__ __forceinline__ void init(float* st, uint64_t* b, int fl, int lane) {
    stages = st; bars = b; stage_floats = fl; phase = 0u; via_bulk = 0u;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ float* stage(int s) const { return stages + s * stage_floats; }
  // up to two tensors land in one stage (facts, then upstream gradients) and complete on the same barrier
  __device__ __forceinline__ void issue(int s, const float* src0, int n0, int off1, const float* src1, int n1, bool al, int lane,
                                        float fill0) {
    float* dst = stage(s);
    const uint32_t b0 = (uint32_t)n0 * 4u, b1 = (uint32_t)n1 * 4u;
    if (al && ((b0 | b1) & 15u) == 0u && n0 > 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[s], b0 + b1);
        bulk_g2s(dst, src0, b0, &bars[s]);
        if (n1 > 0) bulk_g2s(dst + off1, src1, b1, &bars[s]);
      }
      via_bulk |= 1u << s;
    } else {
      for (int i = lane; i < n0; i += 32) dst[i] = __ldg(src0 + i);
      for (int i = lane; i < n1; i += 32) dst[off1 + i] = __ldg(src1 + i);
      via_bulk &= ~(1u << s);
    }
    (void)fill0;
  }
  __device__ __forceinline__ void wait(int s) {
    if (via_bulk & (1u << s)) { mbar_wait(&bars[s], (phase >> s) & 1u); phase ^= 1u << s; }
    __syncwarp();
  }
};
__device__ __forceinline__ void tile_store(float* gdst, const float* st, int n, int n_full, bool al, int lane) {
  if (al && n == n_full) {
    if (lane == 0) { bulk_s2g(gdst, st, (uint32_t)n * 4u); bulk_commit(); }
  } else {
    for (int i = lane; i < n; i += 32) gdst[i] = st[i];
    if (lane == 0) bulk_commit();
  }
}
)PRELUDE";

struct Gen {
  const vael_circuit_desc_t* d;
  std::ostringstream o;
  std::vector<char> emitted;
  std::vector<int> level_of;
  int kind(int n) const { return d->node_kind[n]; }
  int nch(int n) const { return d->child_ptr[n + 1] - d->child_ptr[n]; }
  int ch(int n, int i) const { return d->child_idx[d->child_ptr[n] + i]; }
  static std::string flit(float v) {
    char buf[64];
    uint32_t bits;
    memcpy(&bits, &v, 4);
    snprintf(buf, sizeof(buf), "__uint_as_float(0x%08xu)", bits);
    return buf;
  }
  // forward value of node n as `const float v<n>` (children first); everything at function scope
  void value(int n, const char* ind = "    ") {
    if (emitted[n]) return;
    for (int i = 0; i < nch(n); ++i) value(ch(n, i), ind);
    emitted[n] = 1;
    o << ind << "const float v" << n << " = ";
    switch (kind(n)) {
      case VAEL_NODE_LEAF_POS: o << "f" << d->node_arg[n]; break;
      case VAEL_NODE_LEAF_NEG: o << "__fsub_rn(1.0f, f" << d->node_arg[n] << ")"; break;
      case VAEL_NODE_CONST: o << flit(d->const_val[d->node_arg[n]]); break;
      case VAEL_NODE_PROD:
      case VAEL_NODE_SUM:
      case VAEL_NODE_COMPL: {
        const bool prod = kind(n) == VAEL_NODE_PROD;
        const int nc = nch(n);
        if (nc == 0) { o << (prod ? "1.0f" : "0.0f"); break; }
        std::string e = "v" + std::to_string(ch(n, 0));
        for (int i = 1; i < nc; ++i)
          e = std::string(prod ? "__fmul_rn(" : "__fadd_rn(") + e + ", v" + std::to_string(ch(n, i)) + ")";
        if (kind(n) == VAEL_NODE_COMPL) e = "__fsub_rn(1.0f, " + e + ")";
        o << e;
        break;
      }
    }
    o << ";\n";
  }
  // adjoint of node n (held in `an`) -> its children's accumulators a<c> / the fact-column gradients gf<col>
  void scatter(int n, const std::string& an, const std::vector<char>& live, const char* ind = "    ") {
    const int k = kind(n), nc = nch(n);
    if (k == VAEL_NODE_LEAF_POS) { o << ind << "gf" << d->node_arg[n] << " += " << an << ";\n"; return; }
    if (k == VAEL_NODE_LEAF_NEG) { o << ind << "gf" << d->node_arg[n] << " -= " << an << ";\n"; return; }
    if (k == VAEL_NODE_SUM || k == VAEL_NODE_COMPL) {
      for (int i = 0; i < nc; ++i)
        if (live[ch(n, i)]) o << ind << "a" << ch(n, i) << (k == VAEL_NODE_SUM ? " += " : " -= ") << an << ";\n";
      return;
    }
    if (k == VAEL_NODE_PROD) {
      for (int j = 0; j < nc; ++j) {
        const int c = ch(n, j);
        if (!live[c]) continue;
        std::string sib;
        for (int i = 0; i < nc; ++i) {
          if (i == j) continue;
          value(ch(n, i), ind);
          const std::string v = "v" + std::to_string(ch(n, i));
          sib = sib.empty() ? v : "(" + sib + " * " + v + ")";
        }
        if (sib.empty()) o << ind << "a" << c << " += " << an << ";\n";
        else o << ind << "a" << c << " = fmaf(" << an << ", " << sib << ", a" << c << ");\n";
      }
    }
  }
};

// facts row of this lane -> registers f0 .. f<F-1> (vector loads when the row pitch allows)
void emit_fact_loads(std::ostringstream& o, int F) {
  const int V = (F % 4 == 0) ? 4 : ((F % 2 == 0) ? 2 : 1);
  for (int c = 0; c < F; c += V) {
    if (V == 4) {
      o << "    const float4 t" << c << " = *reinterpret_cast<const float4*>(fr + " << c << ");\n";
      o << "    const float f" << c << " = t" << c << ".x, f" << c + 1 << " = t" << c << ".y, f" << c + 2 << " = t" << c << ".z, f"
        << c + 3 << " = t" << c << ".w;\n";
    } else if (V == 2) {
      o << "    const float2 t" << c << " = *reinterpret_cast<const float2*>(fr + " << c << ");\n";
      o << "    const float f" << c << " = t" << c << ".x, f" << c + 1 << " = t" << c << ".y;\n";
    } else {
      o << "    const float f" << c << " = fr[" << c << "];\n";
    }
  }
}

// row of K outputs from expressions e[0..K) -> shared-memory row `os` (vector stores when K % 4 / 2 == 0)
void emit_row_stores(std::ostringstream& o, const std::vector<std::string>& e, const char* os, const char* ind) {
  const int K = (int)e.size();
  const int V = (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1);
  for (int c = 0; c < K; c += V) {
    if (V == 4)
      o << ind << "*reinterpret_cast<float4*>(" << os << " + " << c << ") = make_float4(" << e[c] << ", " << e[c + 1] << ", "
        << e[c + 2] << ", " << e[c + 3] << ");\n";
    else if (V == 2)
      o << ind << "*reinterpret_cast<float2*>(" << os << " + " << c << ") = make_float2(" << e[c] << ", " << e[c + 1] << ");\n";
    else
      o << ind << os << "[" << c << "] = " << e[c] << ";\n";
  }
}

// Grid shape of the generated kernels.  Default: NON-persistent — one tile per warp, a few warps per CTA, as many CTAs
// as tiles, single-buffered stages: the access front of such a grid advances compactly through memory and DRAM
// streams run 10-15 % faster than with persistent warps striding over the tiles (scripts/micro/storebench.cu; the
// hand-written two-AD kernels: 0.69 -> 0.61 ms and 0.76 -> 0.66 ms).  VAEL_JIT_PERSISTENT=1: the earlier shape (one
// persistent CTA per SM, double-buffered stages), for A-B timing.
static bool jit_persistent() {
  static const bool v = [] { const char* e = getenv("VAEL_JIT_PERSISTENT"); return e != nullptr && e[0] == '1'; }();
  return v;
}
static int jit_stages() { return jit_persistent() ? 2 : 1; }

std::string generate_source(const vael_circuit_desc_t* d) {
  const int N = d->n_nodes, F = d->n_facts, W = d->n_worlds, Q = d->n_queries, MW = (W + 31) / 32;
  const int OC = std::max(std::max(W, Q), 1);
  Gen g;
  g.d = d;
  g.level_of.assign(N, 0);
  for (int L = 0; L < d->n_levels; ++L)
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n) g.level_of[n] = L;
  std::ostringstream& o = g.o;
  o << kPrelude;
  o << "#define F " << F << "\n#define W " << W << "\n#define NQ " << Q << "\n#define MW " << MW << "\n#define OC " << OC << "\n";
  o << "#define NST " << jit_stages() << "\n#define NSO " << jit_stages() << "\n";
  o << "#define FB ((32 * F * 4 + 127) / 128 * 128)\n#define OB ((32 * OC * 4 + 127) / 128 * 128)\n#define GB ((32 * W * 4 + 127) / 128 * 128)\n";

  // which nodes only the normaliser needs (emitted inside `if constexpr (NORM)`)
  std::vector<char> out_cone(N, 0);
  for (int w = 0; w < W; ++w) out_cone[d->world_node[w]] = 1;
  for (int q = 0; q < Q; ++q) out_cone[d->query_node[q]] = 1;
  for (int n = N - 1; n >= 0; --n)
    if (out_cone[n])
      for (int i = 0; i < g.nch(n); ++i) out_cone[g.ch(n, i)] = 1;
  std::vector<char> world_cone(N, 0);
  for (int w = 0; w < W; ++w) world_cone[d->world_node[w]] = 1;
  for (int n = N - 1; n >= 0; --n)
    if (world_cone[n])
      for (int i = 0; i < g.nch(n); ++i) world_cone[g.ch(n, i)] = 1;

  // ================================ forward ================================
  o << R"(
template <bool NORM, bool EVID, bool QALL>
__device__ __forceinline__ void fwd_body(const JitParams& P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned char* wb = smem + (size_t)warp * (128 + NST * FB + NSO * OB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* ostage = reinterpret_cast<float*>(wb + 128 + NST * FB);
  const long long n_sub = (P.B + 31) >> 5;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool in_al = aligned16(P.facts);
  float* const outp = QALL ? P.queries_all : P.worlds;
  constexpr int OK = QALL ? NQ : W;                       // columns of the output tile
  const bool out_al = outp != nullptr && aligned16(outp) && ((32 * OK * 4) % 16 == 0);
  const bool needq = QALL || EVID || P.query != nullptr;
  TileRing<NST> ld;
  ld.init(reinterpret_cast<float*>(wb + 128), bars, FB / 4, lane);
  auto request = [&](long long sub, int s) {
    if (sub < n_sub) ld.issue(s, P.facts + sub * 32 * F, (int)min(32LL, P.B - sub * 32) * F, 0, nullptr, 0, in_al, lane, 0.5f);
  };
#pragma unroll
  for (int s = 0; s < NST; ++s) request(gw + s * GW, s);
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % NST);
    const long long row0 = sub * 32, row = row0 + lane;
    const int rows = (int)min(32LL, P.B - row0);
    const bool valid = lane < rows;
    int q = -1;
    if (needq && !QALL) {
      q = valid ? (P.qidx ? __ldg(P.qidx + row) : P.qscalar) : -1;
      if (q < 0 || q >= NQ) q = -1;
    }
    unsigned m[MW > 0 ? MW : 1];
#pragma unroll
    for (int i = 0; i < MW; ++i) m[i] = (EVID && q >= 0) ? __ldg(P.qmask + (long long)q * MW + i) : 0u;
    ld.wait(s);
    const float* fr = ld.stage(s) + lane * F;
)";
  emit_fact_loads(o, F);
  o << "    __syncwarp();\n    request(sub + NST * GW, s);\n";
  g.emitted.assign(N, 0);
  // leaves first (plain register aliases), then the normaliser's private cone, the worlds' cones, the queries' cones
  for (int n = 0; n < N; ++n)
    if (g.nch(n) == 0 && out_cone[n]) g.value(n);
  o << "    float z = 1.0f;\n";
  if (d->norm_node >= 0) {
    // nodes of the normaliser's cone that outputs use too stay at function scope; its private nodes live inside
    // the `if constexpr (NORM)` block (and are forgotten afterwards: the block's names are not visible outside)
    std::vector<int> stack{d->norm_node};
    std::vector<char> in_norm(N, 0);
    while (!stack.empty()) {
      const int n = stack.back();
      stack.pop_back();
      if (in_norm[n]) continue;
      in_norm[n] = 1;
      for (int i = 0; i < g.nch(n); ++i) stack.push_back(g.ch(n, i));
    }
    for (int n = 0; n < N; ++n)
      if (in_norm[n] && out_cone[n]) g.value(n);
    o << "    if constexpr (NORM) {\n";
    const std::vector<char> before = g.emitted;
    g.value(d->norm_node, "      ");
    o << "      z = v" << d->norm_node << ";\n    }\n";
    for (int n = 0; n < N; ++n)
      if (g.emitted[n] && !before[n]) g.emitted[n] = 0;
  }
  for (int w = 0; w < W; ++w) g.value(d->world_node[w]);
  o << "    float qv = 0.0f;\n";
  if (Q > 0) {
    o << "    float qo[NQ > 0 ? NQ : 1];\n    if (needq) {\n";
    std::vector<char> before = g.emitted;
    for (int q = 0; q < Q; ++q) g.value(d->query_node[q], "      ");
    for (int q = 0; q < Q; ++q) o << "      qo[" << q << "] = v" << d->query_node[q] << ";\n";
    o << "      if (!QALL) {\n";
    for (int q = 0; q < Q; ++q) o << "        qv = (q == " << q << ") ? v" << d->query_node[q] << " : qv;\n";
    o << "      }\n    }\n";
    for (int n = 0; n < N; ++n)
      if (g.emitted[n] && !before[n]) g.emitted[n] = 0;
  }
  o << "    if (outp != nullptr) {\n      if (lane == 0) bulk_wait_read<NSO - 1>();\n      __syncwarp();\n";
  o << "      float* os = ostage + (it % NSO) * (OB / 4) + lane * OK;\n";
  o << "      if constexpr (QALL) {\n";
  {
    std::vector<std::string> e;
    for (int q = 0; q < Q; ++q) e.push_back("(NORM ? qo[" + std::to_string(q) + "] / z : qo[" + std::to_string(q) + "])");
    if (Q > 0) emit_row_stores(o, e, "os", "        ");
  }
  o << "      } else {\n";
  {
    std::vector<std::string> e;
    for (int w = 0; w < W; ++w) {
      const std::string v = "v" + std::to_string(d->world_node[w]);
      e.push_back("(EVID ? (((m[" + std::to_string(w >> 5) + "] >> " + std::to_string(w & 31) + ") & 1u) ? " + v + " / qv : 0.0f) : (NORM ? " +
                  v + " / z : " + v + "))");
    }
    if (W > 0) emit_row_stores(o, e, "os", "        ");
  }
  o << R"(      }
      fence_proxy_async();
      __syncwarp();
      tile_store(outp + row0 * OK, ostage + (it % NSO) * (OB / 4), rows * OK, 32 * OK, out_al, lane);
    }
    if (!QALL && P.query != nullptr && valid) P.query[row] = NORM ? qv / z : qv;
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}
extern "C" __global__ void __launch_bounds__(256) jit_fwd_00(const JitParams P) { fwd_body<false, false, false>(P); }
extern "C" __global__ void __launch_bounds__(256) jit_fwd_10(const JitParams P) { fwd_body<true, false, false>(P); }
extern "C" __global__ void __launch_bounds__(256) jit_fwd_01(const JitParams P) { fwd_body<false, true, false>(P); }
extern "C" __global__ void __launch_bounds__(256) jit_fwd_11(const JitParams P) { fwd_body<true, true, false>(P); }
extern "C" __global__ void __launch_bounds__(256) jit_qall_0(const JitParams P) { fwd_body<false, false, true>(P); }
extern "C" __global__ void __launch_bounds__(256) jit_qall_1(const JitParams P) { fwd_body<true, false, true>(P); }
)";

  // ================================ backward ================================
  // live nodes: those some output depends on AND that lead to a leaf (constants carry no gradient)
  std::vector<char> has_leaf(N, 0);
  for (int n = 0; n < N; ++n) {
    if (g.kind(n) == VAEL_NODE_LEAF_POS || g.kind(n) == VAEL_NODE_LEAF_NEG) has_leaf[n] = 1;
    for (int i = 0; i < g.nch(n); ++i) has_leaf[n] |= has_leaf[g.ch(n, i)];
  }
  std::vector<char> live(N, 0);
  for (int n = 0; n < N; ++n) live[n] = out_cone[n] && has_leaf[n];
  o << R"(
template <bool NORM>
__device__ __forceinline__ void bwd_body(const JitParams& P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned char* wb = smem + (size_t)warp * (128 + NST * (FB + GB) + NSO * FB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* ostage = reinterpret_cast<float*>(wb + 128 + NST * (FB + GB));
  const long long n_sub = (P.B + 31) >> 5;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool has_gw = P.grad_worlds != nullptr;
  const bool in_al = aligned16(P.facts) && (!has_gw || aligned16(P.grad_worlds));
  const bool out_al = aligned16(P.grad_facts);
  TileRing<NST> ld;
  ld.init(reinterpret_cast<float*>(wb + 128), bars, (FB + GB) / 4, lane);
  auto request = [&](long long sub, int s) {
    if (sub >= n_sub) return;
    const int rows = (int)min(32LL, P.B - sub * 32);
    ld.issue(s, P.facts + sub * 32 * F, rows * F, FB / 4, has_gw ? P.grad_worlds + sub * 32 * W : nullptr, has_gw ? rows * W : 0, in_al,
             lane, 0.5f);
  };
#pragma unroll
  for (int s = 0; s < NST; ++s) request(gw + s * GW, s);
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % NST);
    const long long row0 = sub * 32, row = row0 + lane;
    const int rows = (int)min(32LL, P.B - row0);
    const bool valid = lane < rows;
    int q = -1;
    float gq = 0.0f;
    if (P.grad_query != nullptr && valid) {
      q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
      if (q < 0 || q >= NQ) q = -1;
      gq = __ldg(P.grad_query + row);
    }
    ld.wait(s);
    const float* fr = ld.stage(s) + lane * F;
    const float* gr = ld.stage(s) + FB / 4 + lane * W;
)";
  emit_fact_loads(o, F);
  // upstream world gradients of this lane's row -> registers gw<w>
  {
    const int V = (W % 4 == 0) ? 4 : ((W % 2 == 0) ? 2 : 1);
    for (int c = 0; c < W; c += V) {
      if (V == 4) {
        o << "    const float4 u" << c << " = has_gw ? *reinterpret_cast<const float4*>(gr + " << c << ") : make_float4(0.f, 0.f, 0.f, 0.f);\n";
        o << "    const float gw" << c << " = u" << c << ".x, gw" << c + 1 << " = u" << c << ".y, gw" << c + 2 << " = u" << c << ".z, gw"
          << c + 3 << " = u" << c << ".w;\n";
      } else if (V == 2) {
        o << "    const float2 u" << c << " = has_gw ? *reinterpret_cast<const float2*>(gr + " << c << ") : make_float2(0.f, 0.f);\n";
        o << "    const float gw" << c << " = u" << c << ".x, gw" << c + 1 << " = u" << c << ".y;\n";
      } else {
        o << "    const float gw" << c << " = has_gw ? gr[" << c << "] : 0.0f;\n";
      }
    }
  }
  o << "    __syncwarp();\n    request(sub + NST * GW, s);\n";
  g.emitted.assign(N, 0);
  for (int n = 0; n < N; ++n)
    if (g.nch(n) == 0 && (out_cone[n] || (d->norm_node >= 0))) g.value(n);
  o << "    float iz = 1.0f;\n";
  if (d->norm_node >= 0) {
    // the normaliser's cone: values only (Z's own adjoint is exactly 0: the interpreter's and ad2's convention)
    o << "    if constexpr (NORM) {\n";
    std::vector<char> before = g.emitted;
    // shared interior nodes are re-derived inside the block (cheap: the cone is a handful of sums)
    g.value(d->norm_node, "      ");
    o << "      iz = v" << d->norm_node << ";\n    }\n";
    for (int n = 0; n < N; ++n)
      if (g.emitted[n] && !before[n]) g.emitted[n] = 0;
  }
  // adjoint accumulators of the live nodes and of the fact columns
  for (int n = 0; n < N; ++n)
    if (live[n]) o << "    float a" << n << " = 0.0f;\n";
  for (int c = 0; c < F; ++c) o << "    float gf" << c << " = 0.0f;\n";
  // seeds: worlds (upstream / Z), the row's query node (gq / Z)
  for (int w = 0; w < W; ++w)
    if (live[d->world_node[w]]) o << "    a" << d->world_node[w] << " += NORM ? gw" << w << " / iz : gw" << w << ";\n";
  if (Q > 0) {
    o << "    { const float gqz = NORM ? gq / iz : gq;\n";
    for (int q = 0; q < Q; ++q)
      if (live[d->query_node[q]]) o << "      a" << d->query_node[q] << " += (q == " << q << ") ? gqz : 0.0f;\n";
    o << "    }\n";
  }
  // reverse sweep: scatter into the children's accumulators, levels descending (every parent of a node lies in a
  // higher level, so a node's adjoint is complete when its level is reached)
  for (int L = d->n_levels - 1; L >= 0; --L)
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n)
      if (live[n]) g.scatter(n, "a" + std::to_string(n), live);
  o << "    if (lane == 0) bulk_wait_read<NSO - 1>();\n    __syncwarp();\n";
  o << "    float* os = ostage + (it % NSO) * (FB / 4) + lane * F;\n";
  {
    std::vector<std::string> e;
    for (int c = 0; c < F; ++c) e.push_back("gf" + std::to_string(c));
    emit_row_stores(o, e, "os", "    ");
  }
  o << R"(    fence_proxy_async();
    __syncwarp();
    tile_store(P.grad_facts + row0 * F, ostage + (it % NSO) * (FB / 4), rows * F, 32 * F, out_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}
extern "C" __global__ void __launch_bounds__(256) jit_bwd_0(const JitParams P) { bwd_body<false>(P); }
extern "C" __global__ void __launch_bounds__(256) jit_bwd_1(const JitParams P) { bwd_body<true>(P); }
)";
  return o.str();
}

// ---- wide circuits: the [32, W] world tensor does not fit a shared-memory stage (3-digit addition: W = 1000) ----
// The world level is swept in column chunks of WC worlds.  A chunk of a row-major [B, W] tensor is a strided box; here
// EVERY LANE moves its own row segment with its own 1-D bulk copy (WC * 4 contiguous bytes; 32 copies per chunk,
// completing on the warp's mbarrier inbound, tracked by the lane's own bulk group outbound) — no tensor map needed.
// Requirements (checked by wide_plan; anything else runs the interpreter): W a multiple of 4 with a divisor WC in
// [32, 128] that is one too; every world node on ONE level and parent-less apart from query nodes that are plain
// sums over world nodes in world order (their sums are then accumulated world by world, in CSR order, and their
// adjoints are gathered by the worlds instead of being scattered — which would keep W accumulators alive).
struct WidePlan {
  bool ok = false;
  int WC = 0, NCH = 0, world_level = -1;
  std::vector<int> world_of;                      // node -> world index or -1
  std::vector<std::vector<int>> sums_of_world;    // world -> queries (sum type) that contain it, ascending
  std::vector<int> query_world;                   // query -> world index when the query node IS a world node, else -1
};

WidePlan wide_plan(const vael_circuit_desc_t* d) {
  WidePlan P;
  const int N = d->n_nodes, W = d->n_worlds, Q = d->n_queries;
  if (d->semiring != VAEL_SEMIRING_PROB || W <= 128 || (W & 3) || N > 8192 || d->n_facts > 64 || Q > 64) return P;
  for (int wc = 128; wc >= 32; wc -= 4)
    if (W % wc == 0) { P.WC = wc; break; }
  if (!P.WC) return P;
  P.NCH = W / P.WC;
  std::vector<int> level_of(N, 0);
  for (int L = 0; L < d->n_levels; ++L)
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n) level_of[n] = L;
  P.world_of.assign(N, -1);
  for (int w = 0; w < W; ++w) {
    const int n = d->world_node[w];
    if (P.world_of[n] >= 0) return P;                                   // distinct world nodes
    P.world_of[n] = w;
    if (P.world_level < 0) P.world_level = level_of[n];
    if (level_of[n] != P.world_level) return P;
  }
  std::vector<char> is_qsum(N, 0);
  P.query_world.assign(Q, -1);
  P.sums_of_world.assign(W, {});
  std::vector<char> seen_q(N, 0);
  for (int q = 0; q < Q; ++q) {
    const int n = d->query_node[q];
    if (P.world_of[n] >= 0) { P.query_world[q] = P.world_of[n]; continue; }
    if (d->node_kind[n] != VAEL_NODE_SUM) return P;
    int prev = -1;
    for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) {
      const int w = P.world_of[d->child_idx[e]];
      if (w <= prev) return P;                                          // world nodes, strictly increasing
      prev = w;
      if (!seen_q[n]) P.sums_of_world[w].push_back(q);
    }
    if (seen_q[n]) {   // two queries sharing one sum node: list the second one too
      for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) P.sums_of_world[P.world_of[d->child_idx[e]]].push_back(q);
    }
    seen_q[n] = 1;
    is_qsum[n] = 1;
  }
  // world nodes and query sums have no other parents
  for (int n = 0; n < N; ++n) {
    if (is_qsum[n]) continue;
    for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) {
      const int c = d->child_idx[e];
      if (P.world_of[c] >= 0 || is_qsum[c]) {
        // tolerated only when n is outside every output's cone AND the normaliser's (dead code); keep it simple:
        return P;
      }
    }
  }
  P.ok = true;
  return P;
}

// store stages of the column-chunked forward: the box stores want few streams, each deep (the adw kernels' finding)
static int wide_store_stages() {
  const char* v = getenv("VAEL_JIT_WIDE_NSO");
  const int n = (v && *v) ? atoi(v) : 4;
  return std::max(2, std::min(6, n));
}

std::string generate_source_wide(const vael_circuit_desc_t* d, const WidePlan& PL) {
  const int N = d->n_nodes, F = d->n_facts, W = d->n_worlds, Q = d->n_queries, WC = PL.WC, NCH = PL.NCH;
  const int OC = std::max(WC, std::max(Q, 1));
  Gen g;
  g.d = d;
  std::ostringstream& o = g.o;
  o << kPrelude;
  o << "#define F " << F << "\n#define W " << W << "\n#define NQ " << Q << "\n#define WC " << WC << "\n#define NCH " << NCH
    << "\n#define OC " << OC << "\n#define NST " << jit_stages() << "\n#define NSO " << jit_stages() << "\n#define NSF " << wide_store_stages() << "\n#define NSG 2\n";
  o << "#define FB ((32 * F * 4 + 127) / 128 * 128)\n#define OB ((32 * OC * 4 + 127) / 128 * 128)\n#define GB ((32 * WC * 4 + 127) / 128 * 128)\n";
  std::vector<char> out_cone(N, 0);
  for (int w = 0; w < W; ++w) out_cone[d->world_node[w]] = 1;
  for (int q = 0; q < Q; ++q) out_cone[d->query_node[q]] = 1;
  for (int n = N - 1; n >= 0; --n)
    if (out_cone[n])
      for (int i = 0; i < g.nch(n); ++i) out_cone[g.ch(n, i)] = 1;
  auto norm_block = [&](const char* zname) {
    o << "    float " << zname << " = 1.0f;\n";
    if (d->norm_node < 0) return;
    std::vector<int> stack{d->norm_node};
    std::vector<char> in_norm(N, 0);
    while (!stack.empty()) {
      const int n = stack.back();
      stack.pop_back();
      if (in_norm[n]) continue;
      in_norm[n] = 1;
      for (int i = 0; i < g.nch(n); ++i) stack.push_back(g.ch(n, i));
    }
    for (int n = 0; n < N; ++n)
      if (in_norm[n] && out_cone[n]) g.value(n);
    o << "    if constexpr (NORM) {\n";
    const std::vector<char> before = g.emitted;
    g.value(d->norm_node, "      ");
    o << "      " << zname << " = v" << d->norm_node << ";\n    }\n";
    for (int n = 0; n < N; ++n)
      if (g.emitted[n] && !before[n]) g.emitted[n] = 0;
  };

  // ================================ forward ================================
  o << R"(
template <bool NORM, bool QALL>
__device__ __forceinline__ void fwd_body(const JitParams& P, const TensorMap* tm_w) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned char* wb = smem + (size_t)warp * (128 + NST * FB + NSF * OB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  float* ostage = reinterpret_cast<float*>(wb + 128 + NST * FB);
  const long long n_sub = (P.B + 31) >> 5;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool in_al = aligned16(P.facts);
  const bool qall_al = QALL && aligned16(P.queries_all) && ((32 * NQ * 4) % 16 == 0);
  const bool needq = QALL || P.query != nullptr;
  const bool want_w = !QALL && P.worlds != nullptr;
  TileRing<NST> ld;
  ld.init(reinterpret_cast<float*>(wb + 128), bars, FB / 4, lane);
  auto request = [&](long long sub, int s) {
    if (sub < n_sub) ld.issue(s, P.facts + sub * 32 * F, (int)min(32LL, P.B - sub * 32) * F, 0, nullptr, 0, in_al, lane, 0.5f);
  };
#pragma unroll
  for (int s = 0; s < NST; ++s) request(gw + s * GW, s);
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % NST);
    const long long row0 = sub * 32, row = row0 + lane;
    const int rows = (int)min(32LL, P.B - row0);
    const bool valid = lane < rows;
    int q = -1;
    if (needq && !QALL) {
      q = valid ? (P.qidx ? __ldg(P.qidx + row) : P.qscalar) : -1;
      if (q < 0 || q >= NQ) q = -1;
    }
    ld.wait(s);
    const float* fr = ld.stage(s) + lane * F;
)";
  emit_fact_loads(o, F);
  o << "    __syncwarp();\n    request(sub + NST * GW, s);\n";
  g.emitted.assign(N, 0);
  for (int n = 0; n < N; ++n)
    if (g.nch(n) == 0 && out_cone[n]) g.value(n);
  norm_block("z");
  for (int q = 0; q < Q; ++q) o << "    float qa" << q << " = 0.0f;\n";
  for (int c = 0; c < NCH; ++c) {
    const int c0 = c * WC;
    o << "    // ---- worlds " << c0 << " .. " << c0 + WC - 1 << "\n";
    for (int w = c0; w < c0 + WC; ++w) g.value(d->world_node[w]);
    o << "    if (needq) {\n";
    for (int w = c0; w < c0 + WC; ++w) {
      for (int q : PL.sums_of_world[w]) o << "      qa" << q << " = __fadd_rn(qa" << q << ", v" << d->world_node[w] << ");\n";
      for (int q = 0; q < Q; ++q)
        if (PL.query_world[q] == w) o << "      qa" << q << " = v" << d->world_node[w] << ";\n";
    }
    o << "    }\n";
    o << "    if (want_w) {\n      if (lane == 0) bulk_wait_read<NSF - 1>();   // the box store of NSF chunks ago has read its stage\n";
    o << "      __syncwarp();\n";
    o << "      float* ob = ostage + ((it * NCH + " << c << ") % NSF) * (OB / 4);\n      float* os = ob + lane * WC;\n";
    std::vector<std::string> e;
    for (int w = c0; w < c0 + WC; ++w) {
      const std::string v = "v" + std::to_string(d->world_node[w]);
      e.push_back("(NORM ? " + v + " / z : " + v + ")");
    }
    emit_row_stores(o, e, "os", "      ");
    o << "      fence_proxy_async();\n      __syncwarp();\n";
    o << "      if (lane == 0) { tma_store_2d(tm_w, ob, " << c0 << ", (int)row0); bulk_commit(); }\n    }\n";
  }
  o << "    float qv = 0.0f;\n";
  for (int q = 0; q < Q; ++q) o << "    qv = (q == " << q << ") ? qa" << q << " : qv;\n";
  o << "    if constexpr (QALL) {\n      if (lane == 0) bulk_wait_read<NSF - 1>();\n      __syncwarp();\n";
  o << "      float* os = ostage + (it % NSF) * (OB / 4) + lane * NQ;\n";
  {
    std::vector<std::string> e;
    for (int q = 0; q < Q; ++q) e.push_back("(NORM ? qa" + std::to_string(q) + " / z : qa" + std::to_string(q) + ")");
    if (Q > 0) emit_row_stores(o, e, "os", "      ");
  }
  o << R"(      fence_proxy_async();
      __syncwarp();
      tile_store(P.queries_all + row0 * NQ, ostage + (it % NSF) * (OB / 4), rows * NQ, 32 * NQ, qall_al, lane);
    }
    if (!QALL && P.query != nullptr && valid) P.query[row] = NORM ? qv / z : qv;
  }
  bulk_wait_all<0>();
  __syncwarp();
}
extern "C" __global__ void __launch_bounds__(256) jit_fwd_00(const JitParams P, const __grid_constant__ TensorMap tm) { fwd_body<false, false>(P, &tm); }
extern "C" __global__ void __launch_bounds__(256) jit_fwd_10(const JitParams P, const __grid_constant__ TensorMap tm) { fwd_body<true, false>(P, &tm); }
extern "C" __global__ void __launch_bounds__(256) jit_qall_0(const JitParams P, const __grid_constant__ TensorMap tm) { fwd_body<false, true>(P, &tm); }
extern "C" __global__ void __launch_bounds__(256) jit_qall_1(const JitParams P, const __grid_constant__ TensorMap tm) { fwd_body<true, true>(P, &tm); }
)";

  // ================================ backward ================================
  std::vector<char> has_leaf(N, 0);
  for (int n = 0; n < N; ++n) {
    if (g.kind(n) == VAEL_NODE_LEAF_POS || g.kind(n) == VAEL_NODE_LEAF_NEG) has_leaf[n] = 1;
    for (int i = 0; i < g.nch(n); ++i) has_leaf[n] |= has_leaf[g.ch(n, i)];
  }
  std::vector<char> live(N, 0);
  for (int n = 0; n < N; ++n) live[n] = out_cone[n] && has_leaf[n];
  o << R"(
template <bool NORM>
__device__ __forceinline__ void bwd_body(const JitParams& P, const TensorMap* tm_g) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned char* wb = smem + (size_t)warp * (128 + NST * FB + NSG * GB + NSO * FB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb);
  uint64_t* gbars = bars + NST;
  float* gstage = reinterpret_cast<float*>(wb + 128 + NST * FB);
  float* ostage = reinterpret_cast<float*>(wb + 128 + NST * FB + NSG * GB);
  const long long n_sub = (P.B + 31) >> 5;
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  if (gw >= n_sub) return;
  const bool has_gw = P.grad_worlds != nullptr;
  const bool in_al = aligned16(P.facts);
  const bool out_al = aligned16(P.grad_facts);
  TileRing<NST> ld;
  ld.init(reinterpret_cast<float*>(wb + 128), bars, FB / 4, lane);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSG; ++s) mbar_init(&gbars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto request = [&](long long sub, int s) {
    if (sub < n_sub) ld.issue(s, P.facts + sub * 32 * F, (int)min(32LL, P.B - sub * 32) * F, 0, nullptr, 0, in_al, lane, 0.5f);
  };
  // chunk ring: one [32, WC] box of the upstream gradient per chunk (rows beyond B arrive as zeros); the chunks of all
  // tiles of this warp form one sequence gc
  unsigned gc = 0;   // chunks consumed so far
  auto gissue = [&](long long sub, int c, unsigned k) {   // chunk c of tile `sub` as the k-th chunk of the sequence
    if (!has_gw || sub >= n_sub) return;
    __syncwarp();   // every lane has read the stage this box lands in
    if (lane == 0) {
      mbar_arrive_expect_tx(&gbars[k % NSG], 32u * WC * 4u);   // the box is always delivered whole
      tma_load_2d(gstage + (k % NSG) * (GB / 4), tm_g, c * WC, (int)(sub * 32), &gbars[k % NSG]);
    }
  };
#pragma unroll
  for (int s = 0; s < NST; ++s) request(gw + s * GW, s);
  gissue(gw, 0, 0);
  long long it = 0;
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int s = (int)(it % NST);
    const long long row0 = sub * 32, row = row0 + lane;
    const int rows = (int)min(32LL, P.B - row0);
    const bool valid = lane < rows;
    int q = -1;
    float gq = 0.0f;
    if (P.grad_query != nullptr && valid) {
      q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
      if (q < 0 || q >= NQ) q = -1;
      gq = __ldg(P.grad_query + row);
    }
    ld.wait(s);
    const float* fr = ld.stage(s) + lane * F;
)";
  emit_fact_loads(o, F);
  o << "    __syncwarp();\n    request(sub + NST * GW, s);\n";
  g.emitted.assign(N, 0);
  for (int n = 0; n < N; ++n)
    if (g.nch(n) == 0) g.value(n);
  norm_block("zb");
  // forward values the worlds' scatter needs (their children: siblings inside the products) — at function scope,
  // before the chunk blocks that use them
  for (int w = 0; w < W; ++w)
    for (int i = 0; i < g.nch(d->world_node[w]); ++i) g.value(g.ch(d->world_node[w], i));
  // accumulators: live nodes below the world level; fact-column gradients
  std::vector<int> level_of(N, 0);
  for (int L = 0; L < d->n_levels; ++L)
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n) level_of[n] = L;
  for (int n = 0; n < N; ++n)
    if (live[n] && level_of[n] < PL.world_level) o << "    float a" << n << " = 0.0f;\n";
  for (int c = 0; c < F; ++c) o << "    float gf" << c << " = 0.0f;\n";
  o << "    const float gqz = NORM ? gq / zb : gq;\n";
  for (int q = 0; q < Q; ++q) o << "    const float aq" << q << " = (q == " << q << ") ? gqz : 0.0f;\n";
  for (int c = 0; c < NCH; ++c) {
    const int c0 = c * WC;
    o << "    // ---- worlds " << c0 << " .. " << c0 + WC - 1 << "\n";
    // prefetch the next chunk (of this tile, or the first one of this warp's next tile), then wait for this one
    if (c + 1 < NCH) o << "    gissue(sub, " << c + 1 << ", gc + 1);\n";
    else o << "    gissue(sub + GW, 0, gc + 1);\n";
    o << "    if (has_gw) mbar_wait(&gbars[gc % NSG], (gc / NSG) & 1u);\n";
    o << "    { const float* gr = gstage + (gc % NSG) * (GB / 4) + lane * WC;\n";
    for (int w0 = c0; w0 < c0 + WC; w0 += 4) {
      o << "      { const float4 u = has_gw ? *reinterpret_cast<const float4*>(gr + " << (w0 - c0) << ") : make_float4(0.f, 0.f, 0.f, 0.f);\n";
      const char* comp[4] = {"u.x", "u.y", "u.z", "u.w"};
      for (int j = 0; j < 4; ++j) {
        const int w = w0 + j, n = d->world_node[w];
        if (!live[n]) continue;
        o << "        { float aw = NORM ? " << comp[j] << " / zb : " << comp[j] << ";\n";
        for (int qq : PL.sums_of_world[w]) o << "          aw += aq" << qq << ";\n";
        for (int qq = 0; qq < Q; ++qq)
          if (PL.query_world[qq] == w) o << "          aw += aq" << qq << ";\n";
        g.scatter(n, "aw", live, "          ");
        o << "        }\n";
      }
      o << "      }\n";
    }
    o << "    }\n    ++gc;\n";
  }
  // the levels below the worlds: plain scatter sweep
  for (int L = PL.world_level - 1; L >= 0; --L)
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n)
      if (live[n]) g.scatter(n, "a" + std::to_string(n), live);
  o << "    if (lane == 0) bulk_wait_read<NSO - 1>();\n    __syncwarp();\n";
  o << "    float* os = ostage + (it % NSO) * (FB / 4) + lane * F;\n";
  {
    std::vector<std::string> e;
    for (int c = 0; c < F; ++c) e.push_back("gf" + std::to_string(c));
    emit_row_stores(o, e, "os", "    ");
  }
  o << R"(    fence_proxy_async();
    __syncwarp();
    tile_store(P.grad_facts + row0 * F, ostage + (it % NSO) * (FB / 4), rows * F, 32 * F, out_al, lane);
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}
extern "C" __global__ void __launch_bounds__(256) jit_bwd_0(const JitParams P, const __grid_constant__ TensorMap tm) { bwd_body<false>(P, &tm); }
extern "C" __global__ void __launch_bounds__(256) jit_bwd_1(const JitParams P, const __grid_constant__ TensorMap tm) { bwd_body<true>(P, &tm); }
)";
  return o.str();
}
}  // namespace

// ---- public ---------------------------------------------------------------------------------------------------------
struct JitProgram {
  bool ok = false;
  std::string log;
  CUmodule mod = nullptr;
  CUfunction fwd[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [NORM][EVID]
  CUfunction qall[2] = {nullptr, nullptr};
  CUfunction bwd[2] = {nullptr, nullptr};
  int F = 0, W = 0, Q = 0, mask_words = 0;
  int fwd_warp_bytes = 0, bwd_warp_bytes = 0;
  bool wide = false;   // column-chunked world level: no EVIDENCE variant, the [B,W] tensors must be 16-byte aligned
  int WC = 0;          // wide: columns per chunk = box width of the 2-D tensor-map copies
};

struct JitParamsHost {
  const float* facts; const int* qidx; int qscalar; int pad0; long long B; const unsigned* qmask;
  float* worlds; float* query; float* queries_all;
  const float* grad_worlds; const float* grad_query; float* grad_facts;
};

static bool narrow_eligible(const vael_circuit_desc_t* d) {
  return d->semiring == VAEL_SEMIRING_PROB && d->n_worlds <= 128 && d->n_nodes <= 2048 && d->n_facts <= 64 &&
         d->n_queries <= 64 && d->n_worlds + d->n_queries > 0;
}
bool jit_eligible(const vael_circuit_desc_t* d) { return narrow_eligible(d) || wide_plan(d).ok; }

// the generated source of a circuit (narrow: whole world tile per stage; wide: column chunks)
static std::string source_for(const vael_circuit_desc_t* d, bool* wide, int* WC) {
  if (narrow_eligible(d)) { *wide = false; *WC = d->n_worlds; return generate_source(d); }
  const WidePlan P = wide_plan(d);
  *wide = true;
  *WC = P.WC;
  return generate_source_wide(d, P);
}
std::string jit_source(const vael_circuit_desc_t* d) {
  bool w; int wc;
  return jit_eligible(d) ? source_for(d, &w, &wc) : std::string();
}

// ---- opt-in cubin cache ------------------------------------------------------
static uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}
static std::string jit_cache_path(const JitApi& A, const std::string& src) {
  const char* dir = getenv("VAEL_JIT_CACHE_DIR");
  if (dir == nullptr || *dir == 0) return std::string();
  int ver[2] = {0, 0};
  if (A.nvrtcVersion) A.nvrtcVersion(&ver[0], &ver[1]);
  uint64_t h = fnv1a(src.data(), src.size());
  h = fnv1a(ver, sizeof(ver), h);
  static const char target[] = "sm_100a;c++17;fmad;lineinfo;abi2";
  h = fnv1a(target, sizeof(target), h);
  char name[64];
  snprintf(name, sizeof(name), "/vael_jit_%016llx.cubin", static_cast<unsigned long long>(h));
  return std::string(dir) + name;
}
static bool read_file(const std::string& path, std::vector<char>* out) {
  FILE* fh = fopen(path.c_str(), "rb");
  if (fh == nullptr) return false;
  bool ok = false;
  if (fseek(fh, 0, SEEK_END) == 0) {
    const long n = ftell(fh);
    if (n > 0 && fseek(fh, 0, SEEK_SET) == 0) {
      out->resize(static_cast<size_t>(n));
      ok = fread(out->data(), 1, static_cast<size_t>(n), fh) == static_cast<size_t>(n);
    }
  }
  fclose(fh);
  if (!ok) out->clear();
  return ok;
}
// write to a temporary name and rename: a concurrent reader (another rank compiling the same circuit) sees either no
// file or a complete one
static void write_file_atomically(const std::string& path, const std::vector<char>& data) {
  const std::string tmp = path + ".tmp" + std::to_string(static_cast<long long>(getpid()));
  FILE* fh = fopen(tmp.c_str(), "wb");
  if (fh == nullptr) return;
  const bool ok = fwrite(data.data(), 1, data.size(), fh) == data.size();
  if (fclose(fh) != 0 || !ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

JitProgram* jit_create(const vael_circuit_desc_t* d) {
  JitProgram* J = new JitProgram();
  const JitApi& A = jit_api();
  if (!A.ok) { J->log = A.why; return J; }
  if (!jit_eligible(d)) { J->log = "circuit outside the JIT's scope"; return J; }
  int WC = 0;
  const std::string src = source_for(d, &J->wide, &WC);
  if (const char* dump = getenv("VAEL_JIT_DUMP")) {
    if (FILE* fh = fopen(dump, "w")) { fputs(src.c_str(), fh); fclose(fh); }
  }
  // Opt-in disk cache of the compiled cubin (VAEL_JIT_CACHE_DIR): NVRTC needs ~2 s for the 2-digit circuit's source
  // and ~20 s for the 3-digit one's; the key is a hash of the generated source, the NVRTC version and the target.
  std::vector<char> cubin;
  const std::string cache_path = jit_cache_path(A, src);
  auto compile = [&]() -> bool {
    nvrtcProgram prog = nullptr;
    if (A.nvrtcCreateProgram(&prog, src.c_str(), "vael_jit.cu", 0, nullptr, nullptr) != 0) { J->log = "nvrtcCreateProgram failed"; return false; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    const int rc = A.nvrtcCompileProgram(prog, 4, opts);
    size_t logn = 0;
    A.nvrtcGetProgramLogSize(prog, &logn);
    J->log.clear();
    if (logn > 1) { J->log.resize(logn); A.nvrtcGetProgramLog(prog, &J->log[0]); }
    if (rc != 0) { A.nvrtcDestroyProgram(&prog); return false; }
    size_t n = 0;
    A.nvrtcGetCUBINSize(prog, &n);
    cubin.resize(n);
    A.nvrtcGetCUBIN(prog, cubin.data());
    A.nvrtcDestroyProgram(&prog);
    if (!cache_path.empty()) write_file_atomically(cache_path, cubin);   // best effort
    return true;
  };
  const bool cached = !cache_path.empty() && read_file(cache_path, &cubin);
  if (cached) J->log = "cubin from cache " + cache_path;
  else if (!compile()) return J;
  cudaFree(nullptr);   // the runtime's primary context is current from here on
  if (A.cuModuleLoadData(&J->mod, cubin.data()) != 0) {
    // a cache file that does not load (truncated, another toolkit's) is dropped and the circuit compiled afresh
    if (!cached) { J->log += " cuModuleLoadData failed"; return J; }
    remove(cache_path.c_str());
    J->mod = nullptr;
    if (!compile()) return J;
    if (A.cuModuleLoadData(&J->mod, cubin.data()) != 0) { J->log += " cuModuleLoadData failed"; return J; }
    J->log += " (unloadable cache file replaced)";
  }
  auto get = [&](CUfunction* f, const char* name) { return A.cuModuleGetFunction(f, J->mod, name) == 0; };
  bool all = get(&J->fwd[0][0], "jit_fwd_00") && get(&J->fwd[1][0], "jit_fwd_10") && get(&J->qall[0], "jit_qall_0") &&
             get(&J->qall[1], "jit_qall_1") && get(&J->bwd[0], "jit_bwd_0") && get(&J->bwd[1], "jit_bwd_1");
  if (!J->wide) all = all && get(&J->fwd[0][1], "jit_fwd_01") && get(&J->fwd[1][1], "jit_fwd_11");
  if (!all) { J->log += " kernel lookup failed"; return J; }
  J->F = d->n_facts; J->W = d->n_worlds; J->Q = d->n_queries; J->mask_words = (d->n_worlds + 31) / 32;
  J->WC = WC;
  auto pad = [](int b) { return (b + 127) / 128 * 128; };
  const int FB = pad(32 * J->F * 4), OB = pad(32 * std::max(std::max(WC, J->Q), 1) * 4), GB = pad(32 * WC * 4);
  const int ST = jit_stages();
  J->fwd_warp_bytes = 128 + ST * FB + (J->wide ? wide_store_stages() : ST) * OB;
  J->bwd_warp_bytes = J->wide ? 128 + ST * FB + 2 * GB + ST * FB : 128 + ST * (FB + GB) + ST * FB;
  J->ok = true;
  return J;
}

// Generate + compile (no module load: works without a GPU).  0 = compiled, 1 = out of scope, 2 = NVRTC unavailable,
// 3 = compile error; the NVRTC log (or the reason) lands in `log`.
int jit_compile_check(const vael_circuit_desc_t* d, std::string* log, size_t* cubin_bytes) {
  const JitApi& A = jit_api();
  if (!jit_eligible(d)) { *log = "circuit outside the JIT's scope"; return 1; }
  if (!A.rtc_ok) { *log = A.why; return 2; }
  bool wide = false;
  int WC = 0;
  const std::string src = source_for(d, &wide, &WC);
  if (const char* dump = getenv("VAEL_JIT_DUMP")) {
    if (FILE* fh = fopen(dump, "w")) { fputs(src.c_str(), fh); fclose(fh); }
  }
  nvrtcProgram prog = nullptr;
  if (A.nvrtcCreateProgram(&prog, src.c_str(), "vael_jit.cu", 0, nullptr, nullptr) != 0) { *log = "nvrtcCreateProgram failed"; return 3; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
  const int rc = A.nvrtcCompileProgram(prog, 4, opts);
  size_t logn = 0;
  A.nvrtcGetProgramLogSize(prog, &logn);
  if (logn > 1) { log->resize(logn); A.nvrtcGetProgramLog(prog, &(*log)[0]); }
  size_t n = 0;
  if (rc == 0) A.nvrtcGetCUBINSize(prog, &n);
  if (cubin_bytes) *cubin_bytes = n;
  A.nvrtcDestroyProgram(&prog);
  return rc == 0 ? 0 : 3;
}

void jit_destroy(JitProgram* J) {
  if (!J) return;
  if (J->mod) jit_api().cuModuleUnload(J->mod);
  delete J;
}
bool jit_ok(const JitProgram* J) { return J && J->ok; }
const char* jit_log(const JitProgram* J) { return J ? J->log.c_str() : ""; }

// `wide_tensor`: the [B,W] tensor a column-chunked (wide) kernel moves by 2-D tensor-map copies — worlds forward,
// grad_worlds backward — or null when the call does not touch it
static cudaError_t jit_launch(const JitProgram* J, CUfunction fn, int warp_bytes, int warps_per_sm, const GenCall& C, int num_sms,
                              cudaStream_t st, const float* wide_tensor = nullptr, bool resident_only = false) {
  const JitApi& A = jit_api();
  // one persistent CTA per SM holding `warps_per_sm` warps (the ad2 kernels' measured optimum: few, deep pipelines)
  int nw = std::max(1, std::min(8, warps_per_sm));
  while (nw > 1 && nw * warp_bytes > 227 * 1024) --nw;
  const int smem = nw * warp_bytes;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (A.cuFuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, smem) != 0) return cudaErrorInvalidValue;
  const int ctas_per_sm = std::max(1, std::min(4, (227 * 1024) / (smem + 1024)));
  const long long n_sub = (C.B + 31) / 32;
  long long grid = std::max(1LL, (n_sub + nw - 1) / nw);
  if (jit_persistent() || resident_only) grid = std::min(grid, (long long)num_sms * ctas_per_sm);
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  JitParamsHost P{};
  P.facts = C.facts; P.qidx = C.qidx; P.qscalar = C.qscalar; P.B = C.B; P.qmask = C.qmask;
  P.worlds = C.worlds; P.query = C.query; P.queries_all = C.queries_all;
  P.grad_worlds = C.grad_worlds; P.grad_query = C.grad_query; P.grad_facts = C.grad_facts;
  struct alignas(64) TensorMapHost { unsigned long long opaque[16]; } map{};
  if (J->wide && wide_tensor != nullptr && !make_row_box_map(&map, wide_tensor, C.B, J->W, J->WC)) return cudaErrorNotSupported;
  void* args[] = {&P, &map};   // the narrow kernels take P only: the second entry is never read for them
  if (A.cuLaunchKernel(fn, (unsigned)grid, 1, 1, (unsigned)(nw * 32), 1, 1, (unsigned)smem, reinterpret_cast<CUstream>(st), args,
                       nullptr) != 0)
    return cudaErrorLaunchFailure;
  return cudaSuccess;
}

static int jit_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// can the circuit's specialised kernels serve this call?  (wide: no EVIDENCE variant, aligned [B,W] tensors only)
bool jit_can(const JitProgram* J, uint32_t flags, const float* wide_tensor) {
  if (!J || !J->ok) return false;
  if (!J->wide) return true;
  if (flags & VAEL_FLAG_EVIDENCE) return false;
  return wide_tensor == nullptr || (reinterpret_cast<uintptr_t>(wide_tensor) & 15u) == 0;
}

cudaError_t jit_forward(const JitProgram* J, const GenCall& C, int num_sms, cudaStream_t st) {
  const bool norm = C.flags & VAEL_FLAG_NORMALIZE, evid = C.flags & VAEL_FLAG_EVIDENCE;
  // warps per SM (one persistent CTA), measured on B200 (scripts/gpu_r2j.sh): the write-dominated forward wants few
  // store streams in flight for 12.8 KB world tiles (5) and more for smaller tiles / column chunks (7)
  static const int forced = jit_env("VAEL_JIT_FWD_WARPS", 0);
  const int wps = forced ? forced : (jit_persistent() ? (J->wide ? 2 : (J->W >= 96 ? 5 : 7)) : (J->wide ? 2 : 3));
  // all-queries: small result tiles ([32, Q]) — a resident grid of many warps measures better than one tile per warp
  if (C.queries_all != nullptr)
    return jit_launch(J, J->qall[norm ? 1 : 0], J->fwd_warp_bytes, forced ? forced : 8, C, num_sms, st, nullptr, true);
  return jit_launch(J, J->fwd[norm ? 1 : 0][evid ? 1 : 0], J->fwd_warp_bytes, wps, C, num_sms, st, C.worlds);
}
cudaError_t jit_backward(const JitProgram* J, const GenCall& C, int num_sms, cudaStream_t st) {
  static const int wps = jit_env("VAEL_JIT_BWD_WARPS", jit_persistent() ? 7 : 4);   // persistent: more loads in flight
  return jit_launch(J, J->bwd[(C.flags & VAEL_FLAG_NORMALIZE) ? 1 : 0], J->bwd_warp_bytes, wps, C, num_sms, st, C.grad_worlds);
}

}  // namespace vael
