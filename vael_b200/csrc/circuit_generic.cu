// Generic evaluator: level-by-level sweep over the flat, level-ordered CSR circuit.
//
// Reference semantics: problog's sdd.evaluate(semiring=GraphSemiring) walking the compiled
// formula node by node and calling GraphSemiring.plus/times/negate/value/normalize for every
// node (utils/graph_semiring.py:13-60, models/vael.py:98-135), evaluated here ELEMENTWISE per
// batch row (the reference's batch-wide `.any()` identity shortcuts, utils/graph_semiring.py:27-49,
// are a documented divergence: see DESIGN.md "Parity domain").
//
// Execution model: a CTA owns a tile of 32 batch rows; lane <-> row.  Node values live in
// shared memory node-major / row-minor (val[node][33]): every warp-wide access touches 32
// consecutive words (conflict-free whatever the circuit shape; the odd pitch keeps the
// transposed reads of the output phase conflict-free too).  The circuit is turned ONCE, at
// vael_circuit_create, into a schedule: one 16-byte op per node with the children already
// converted to shared-memory word offsets, ops sorted by kind inside each level.  Warps take
// the ops of a level in groups of four; a group of four binary products / sums (the bulk of a
// compiled program) runs as eight independent shared-memory loads, four FP ops, four stores —
// no per-node branching — and the warps meet at one barrier per level.  Batch-major global
// tensors are transposed on the way in/out with coalesced accesses (lane <-> world for [B,W]).
//
// The backward is the reverse sweep in GATHER form over per-child parent records (built at
// creation, dead cones — nodes no output depends on — pruned), so it needs no atomics, no
// zero-fill of the adjoint tile, and is deterministic.
#include <algorithm>
#include <cstring>
#include <vector>

#include "params.cuh"

namespace vael {

constexpr int PITCH = 33;

enum : int {
  OP_LEAF_POS = 0, OP_LEAF_NEG, OP_CONST, OP_MUL2, OP_ADD2, OP_COMPL2, OP_COPY, OP_COMPL1, OP_PRODN, OP_SUMN,
  OP_COMPLN, OP_IDENT_ONE, OP_IDENT_ZERO
};
// parent-record kinds of the gather-form adjoint
enum : int { PR_ADD = 0, PR_SUB, PR_MUL_SIB, PR_MUL_SIBS, PR_LSE };

// ---- device-resident schedule ------------------------------------------------
struct GenProgram {
  int32_t n_nodes, n_facts, n_worlds, n_queries, n_levels, semiring, norm_node, n_bvals;
  // forward
  int4* ops;               // [n_nodes] {dst off, a, b, kind | nc << 8}; offsets are node * PITCH
  int32_t* edges;          // child offsets (node * PITCH) of the n-ary ops
  int32_t op_level_ptr[kMaxLevels + 1];
  int32_t* world_off;      // [W] node * PITCH
  int32_t* query_off;      // [Q]
  // backward: forward values of the nodes the adjoint needs (slot space), then the reverse sweep
  int4* bops;
  int32_t* bedges;
  int32_t bop_level_ptr[kMaxLevels + 1];
  int32_t norm_slot_off;   // slot * PITCH of Z, or -1
  int4* crec;              // live children, level by level: {adj off, first record, n records, seeded}
  int32_t cl_ptr[kMaxLevels + 1];
  int2* prec;              // parent records {parent adj off << 3 | kind, x} (+ one more int2 for kinds 3, 4)
  int32_t* qseed_off;      // [Q] adjoint offsets of the query nodes
  int32_t* zero_off;       // adjoint rows to clear per tile (query nodes that are not world nodes)
  int32_t n_zero;
  int2* leaf_rec;          // per fact column: live leaves {adj off, +1 / -1}
  int32_t* leaf_ptr;       // [F+1]
  size_t fwd_smem, bwd_smem;
};

struct GenArgs {
  GenProgram G;
  GenCall C;
};

// ---- forward op execution ------------------------------------------------------
template <bool LOG>
__device__ __forceinline__ float exec_op(const int4 d, const float* fs, const float* val, const int32_t* edges,
                                         int lane) {
  const int kind = d.w & 0xff, nc = d.w >> 8;
  switch (kind) {
    case OP_LEAF_POS: {
      const float p = fs[d.y + lane];
      return LOG ? logf(p) : p;
    }
    case OP_LEAF_NEG: {
      const float p = __fsub_rn(1.0f, fs[d.y + lane]);
      return LOG ? logf(p) : p;
    }
    case OP_CONST: return __int_as_float(d.y);
    case OP_MUL2: {
      const float x = val[d.y + lane], y = val[d.z + lane];
      return LOG ? __fadd_rn(x, y) : __fmul_rn(x, y);
    }
    case OP_COPY: return val[d.y + lane];
    case OP_COMPL1: return __fsub_rn(1.0f, val[d.y + lane]);
    case OP_IDENT_ONE: return LOG ? 0.0f : 1.0f;
    case OP_IDENT_ZERO: return LOG ? -INFINITY : 0.0f;
    case OP_PRODN: {
      float acc = val[__ldg(edges + d.y) + lane];
      for (int e = 1; e < nc; ++e) {
        const float v = val[__ldg(edges + d.y + e) + lane];
        acc = LOG ? __fadd_rn(acc, v) : __fmul_rn(acc, v);
      }
      return acc;
    }
    default: break;
  }
  // sums (ADD2 / SUMN) and complements (COMPL2 / COMPLN); children as a list either way
  if constexpr (!LOG) {
    float acc;
    if (kind == OP_ADD2 || kind == OP_COMPL2) {
      acc = __fadd_rn(val[d.y + lane], val[d.z + lane]);
    } else {
      acc = val[__ldg(edges + d.y) + lane];
      for (int e = 1; e < nc; ++e) acc = __fadd_rn(acc, val[__ldg(edges + d.y + e) + lane]);
    }
    return (kind == OP_COMPL2 || kind == OP_COMPLN) ? __fsub_rn(1.0f, acc) : acc;
  } else {  // logsumexp (COMPL is rejected for the log semiring at circuit creation)
    float mx = -INFINITY;
    for (int e = 0; e < nc; ++e) {
      const int c = (kind == OP_ADD2) ? (e == 0 ? d.y : d.z) : __ldg(edges + d.y + e);
      mx = fmaxf(mx, val[c + lane]);
    }
    if (mx == -INFINITY) return -INFINITY;
    float s = 0.0f;
    for (int e = 0; e < nc; ++e) {
      const int c = (kind == OP_ADD2) ? (e == 0 ? d.y : d.z) : __ldg(edges + d.y + e);
      s += expf(val[c + lane] - mx);
    }
    return mx + logf(s);
  }
}

// one level of ops: warps take groups of four; homogeneous groups of binary ops run branch-free
template <bool LOG>
__device__ __forceinline__ void run_level(const int4* __restrict__ ops, const int32_t* __restrict__ edges, int beg,
                                          int end, const float* fs, float* val, int warp, int nwarps, int lane) {
  for (int g = beg + 4 * warp; g < end; g += 4 * nwarps) {
    const int cnt = min(4, end - g);
    int4 d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u] = __ldg(ops + g + (u < cnt ? u : 0));
    const int k0 = d[0].w;
    const bool same = cnt == 4 && d[1].w == k0 && d[2].w == k0 && d[3].w == k0;
    if (same && (k0 & 0xff) == OP_MUL2) {
      float x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[u] = val[d[u].y + lane]; y[u] = val[d[u].z + lane]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) val[d[u].x + lane] = LOG ? __fadd_rn(x[u], y[u]) : __fmul_rn(x[u], y[u]);
    } else if (!LOG && same && (k0 & 0xff) == OP_ADD2) {
      float x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[u] = val[d[u].y + lane]; y[u] = val[d[u].z + lane]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) val[d[u].x + lane] = __fadd_rn(x[u], y[u]);
    } else if (same && (k0 & 0xff) <= OP_LEAF_NEG) {
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = fs[d[u].y + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v = (k0 & 0xff) == OP_LEAF_NEG ? __fsub_rn(1.0f, p[u]) : p[u];
        val[d[u].x + lane] = LOG ? logf(v) : v;
      }
    } else if (!LOG && cnt == 4 && ((k0 & 0xff) == OP_SUMN || (k0 & 0xff) == OP_PRODN) &&
               (d[1].w & 0xff) == (k0 & 0xff) && (d[2].w & 0xff) == (k0 & 0xff) && (d[3].w & 0xff) == (k0 & 0xff)) {
      // four left folds side by side (each one sequential in CSR order: bit-exact, yet four chains in flight)
      int nc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) nc[u] = d[u].w >> 8;
      const int ncmax = max(max(nc[0], nc[1]), max(nc[2], nc[3]));
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = val[__ldg(edges + d[u].y) + lane];
      for (int e = 1; e < ncmax; ++e) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (e < nc[u]) {
            const float v = val[__ldg(edges + d[u].y + e) + lane];
            acc[u] = (k0 & 0xff) == OP_SUMN ? __fadd_rn(acc[u], v) : __fmul_rn(acc[u], v);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) val[d[u].x + lane] = acc[u];
    } else {
      for (int u = 0; u < cnt; ++u) val[d[u].x + lane] = exec_op<LOG>(d[u], fs, val, edges, lane);
    }
  }
}

// transposed load of the facts tile: global [rows,F] row-major -> fs[f*PITCH + r]
__device__ __forceinline__ void load_facts_tile(float* fs, const float* facts, long long row0, int rows, int F) {
  const float* src = facts + row0 * F;
  const int n = rows * F;
  for (int i = threadIdx.x; i < kWarp * F; i += blockDim.x) {
    const int r = i / F, f = i - r * F;
    fs[f * PITCH + r] = (i < n) ? __ldg(src + i) : 0.5f;  // padding rows get a harmless in-domain value
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <bool LOG>
__global__ void __launch_bounds__(512) generic_forward_kernel(const GenArgs A) {
  const GenProgram& P = A.G;
  const GenCall& C = A.C;
  extern __shared__ __align__(16) float gsm[];
  float* fs = gsm;                        // [F][PITCH]
  float* val = gsm + P.n_facts * PITCH;   // [n_nodes][PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long n_tiles = (C.B + kWarp - 1) / kWarp;
  const bool norm = (C.flags & VAEL_FLAG_NORMALIZE) && P.norm_node >= 0;
  const bool evid = (C.flags & VAEL_FLAG_EVIDENCE) != 0;
  const int zoff = P.norm_node * PITCH;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kWarp;
    const int rows = (int)min((long long)kWarp, C.B - row0);
    load_facts_tile(fs, C.facts, row0, rows, P.n_facts);
    __syncthreads();
    for (int L = 0; L < P.n_levels; ++L) {
      run_level<LOG>(P.ops, P.edges, P.op_level_ptr[L], P.op_level_ptr[L + 1], fs, val, warp, nwarps, lane);
      __syncthreads();
    }

    // ---- outputs: one warp per batch row, lane <-> world (coalesced stores) ----
    if (C.worlds != nullptr) {
      for (int r = warp; r < rows; r += nwarps) {
        const long long row = row0 + r;
        float z = 1.0f, s = 1.0f;
        const uint32_t* mrow = nullptr;
        if (norm) z = val[zoff + r];
        if (evid) {
          const int32_t q = C.qidx ? __ldg(C.qidx + row) : C.qscalar;
          const bool qok = q >= 0 && q < P.n_queries;
          s = qok ? val[__ldg(P.query_off + q) + r] : 0.0f;
          mrow = qok ? C.qmask + (long long)q * C.mask_words : nullptr;
        }
        float* orow = C.worlds + row * P.n_worlds;
        for (int e0 = 0; e0 < P.n_worlds; e0 += kWarp) {
          const int e = e0 + lane;
          const uint32_t mword = (evid && mrow) ? __ldg(mrow + (e0 >> 5)) : 0u;
          if (e < P.n_worlds) {
            float v = val[__ldg(P.world_off + e) + r];
            if (evid) {
              v = ((mword >> lane) & 1u) ? v / s : 0.0f;
            } else if (norm) {
              v = v / z;
            }
            orow[e] = v;
          }
        }
      }
    }
    if (C.queries_all != nullptr) {
      for (int r = warp; r < rows; r += nwarps) {
        const long long row = row0 + r;
        const float z = norm ? val[zoff + r] : 1.0f;
        for (int q = lane; q < P.n_queries; q += kWarp) {
          float v = val[__ldg(P.query_off + q) + r];
          if (norm) v = v / z;
          C.queries_all[row * P.n_queries + q] = v;
        }
      }
    }
    if (C.query != nullptr && warp == nwarps - 1 && lane < rows) {
      const long long row = row0 + lane;
      const int32_t q = C.qidx ? __ldg(C.qidx + row) : C.qscalar;
      float v = 0.0f;
      if (q >= 0 && q < P.n_queries) {
        v = val[__ldg(P.query_off + q) + lane];
        if (norm) v = v / val[zoff + lane];
      }
      C.query[row] = v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// backward: forward values of the nodes the adjoint needs, then the reverse sweep
// ---------------------------------------------------------------------------
template <bool LOG>
__device__ __forceinline__ float apply_rec(float a, const int2* __restrict__ prec, int i, const float* val,
                                           const float* adj, const int32_t* __restrict__ bedges, int lane) {
  const int2 rec = __ldg(prec + i);
  const int kind = rec.x & 7;
  const float ap = adj[(rec.x >> 3) + lane];
  switch (kind) {
    case PR_ADD: return a + ap;
    case PR_SUB: return a - ap;
    case PR_MUL_SIB: return fmaf(ap, val[rec.y + lane], a);
    case PR_MUL_SIBS: {  // product of the other children of an n-ary product
      const int2 ext = __ldg(prec + i + 1);
      const int nc = ext.y & 0xffff, pos = ext.y >> 16;
      float sib = 1.0f;
      for (int e = 0; e < nc; ++e)
        if (e != pos) sib *= val[__ldg(bedges + ext.x + e) + lane];
      return fmaf(ap, sib, a);
    }
    default: {  // PR_LSE: softmax weight of the child inside a logsumexp parent
      const int2 ext = __ldg(prec + i + 1);
      const float vc = val[rec.y + lane], vp = val[ext.x + lane];
      return fmaf(ap, (vp == -INFINITY) ? 0.0f : expf(vc - vp), a);
    }
  }
}

template <bool LOG>
__global__ void __launch_bounds__(512) generic_backward_kernel(const GenArgs A) {
  const GenProgram& P = A.G;
  const GenCall& C = A.C;
  extern __shared__ __align__(16) float gsm[];
  float* fs = gsm;                               // [F][PITCH]
  float* val = fs + P.n_facts * PITCH;           // [n_bvals][PITCH]
  float* adj = val + P.n_bvals * PITCH;          // [n_nodes][PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long n_tiles = (C.B + kWarp - 1) / kWarp;
  const bool norm = (C.flags & VAEL_FLAG_NORMALIZE) && P.norm_slot_off >= 0;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kWarp;
    const int rows = (int)min((long long)kWarp, C.B - row0);
    load_facts_tile(fs, C.facts, row0, rows, P.n_facts);
    for (int i = threadIdx.x; i < P.n_zero * kWarp; i += blockDim.x) adj[__ldg(P.zero_off + (i >> 5)) + (i & 31)] = 0.0f;
    __syncthreads();

    for (int L = 0; L < P.n_levels; ++L) {
      if (P.bop_level_ptr[L + 1] > P.bop_level_ptr[L]) {
        run_level<LOG>(P.bops, P.bedges, P.bop_level_ptr[L], P.bop_level_ptr[L + 1], fs, val, warp, nwarps, lane);
        __syncthreads();
      }
    }

    // seed the world adjoints from the upstream gradient (assignment: no zero-fill needed)
    for (int r = warp; r < kWarp; r += nwarps) {
      const long long row = row0 + r;
      const bool live = r < rows && C.grad_worlds != nullptr;
      const float z = (norm && r < rows) ? val[P.norm_slot_off + r] : 1.0f;
      for (int e = lane; e < P.n_worlds; e += kWarp) {
        float g = live ? __ldg(C.grad_worlds + row * P.n_worlds + e) : 0.0f;
        if (norm) g = g / z;
        adj[__ldg(P.world_off + e) + r] = g;  // world nodes are distinct (checked at creation)
      }
    }
    __syncthreads();
    if (C.grad_query != nullptr && warp == 0 && lane < rows) {
      const long long row = row0 + lane;
      const int32_t q = C.qidx ? __ldg(C.qidx + row) : C.qscalar;
      if (q >= 0 && q < P.n_queries) {
        float g = __ldg(C.grad_query + row);
        if (norm) g = g / val[P.norm_slot_off + lane];
        adj[__ldg(P.qseed_off + q) + lane] += g;
      }
    }
    __syncthreads();

    // reverse sweep, gather form: adj[c] = seed[c] + sum over live parents p of adj[p] * d val[p] / d val[c]
    for (int L = P.n_levels - 2; L >= 0; --L) {
      const int beg = P.cl_ptr[L], end = P.cl_ptr[L + 1];
      for (int g = beg + 4 * warp; g < end; g += 4 * nwarps) {
        const int cnt = min(4, end - g);
        int4 cr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cr[u] = __ldg(P.crec + g + (u < cnt ? u : 0));
        const bool same = cnt == 4 && cr[1].z == cr[0].z && cr[2].z == cr[0].z && cr[3].z == cr[0].z &&
                          cr[1].w == cr[0].w && cr[2].w == cr[0].w && cr[3].w == cr[0].w;
        if (same && !(cr[0].w & 2)) {  // same record count, no two-word records: four chains side by side
          float a[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) a[u] = (cr[0].w & 1) ? adj[cr[u].x + lane] : 0.0f;
          for (int r = 0; r < cr[0].z; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = apply_rec<LOG>(a[u], P.prec, cr[u].y + r, val, adj, P.bedges, lane);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) adj[cr[u].x + lane] = a[u];
        } else {
          for (int u = 0; u < cnt; ++u) {
            float a = (cr[u].w & 1) ? adj[cr[u].x + lane] : 0.0f;
            for (int r = 0, i = cr[u].y; r < cr[u].z; ++r) {
              const int k = __ldg(&P.prec[i].x) & 7;
              a = apply_rec<LOG>(a, P.prec, i, val, adj, P.bedges, lane);
              i += (k == PR_MUL_SIBS || k == PR_LSE) ? 2 : 1;
            }
            adj[cr[u].x + lane] = a;
          }
        }
      }
      __syncthreads();
    }

    // live leaves -> grad_facts[row, f], coalesced
    for (int i = threadIdx.x; i < rows * P.n_facts; i += blockDim.x) {
      const int r = i / P.n_facts, f = i - r * P.n_facts;
      const float pf = fs[f * PITCH + r];
      float g = 0.0f;
      for (int k = __ldg(P.leaf_ptr + f); k < __ldg(P.leaf_ptr + f + 1); ++k) {
        const int2 lr = __ldg(P.leaf_rec + k);
        const float a = adj[lr.x + r];
        if (!LOG) {
          g += lr.y < 0 ? -a : a;
        } else {
          g += lr.y < 0 ? -a / (1.0f - pf) : a / pf;
        }
      }
      C.grad_facts[row0 * P.n_facts + i] = g;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// host: schedule construction
// ---------------------------------------------------------------------------
template <typename T>
static cudaError_t up(T** dptr, const std::vector<T>& h, cudaError_t prev) {
  *dptr = nullptr;
  if (prev != cudaSuccess) return prev;
  const size_t n = h.empty() ? 1 : h.size();
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dptr), n * sizeof(T));
  if (e != cudaSuccess || h.empty()) return e;
  return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

void gen_program_destroy(GenProgram* g) {
  if (!g) return;
  cudaFree(g->ops); cudaFree(g->edges); cudaFree(g->world_off); cudaFree(g->query_off); cudaFree(g->bops);
  cudaFree(g->bedges); cudaFree(g->crec); cudaFree(g->prec); cudaFree(g->qseed_off); cudaFree(g->zero_off);
  cudaFree(g->leaf_rec); cudaFree(g->leaf_ptr);
  delete g;
}

// One op per node; `off(n)` maps a node id to its value offset (node*PITCH forward, slot*PITCH backward).
template <typename OffFn>
static int4 make_op(const vael_circuit_desc_t* d, int n, OffFn off, std::vector<int32_t>& edges) {
  const int k = d->node_kind[n];
  const int cb = d->child_ptr[n], nc = d->child_ptr[n + 1] - cb;
  int4 op = make_int4(off(n), 0, 0, 0);
  auto kind = [&](int code) { op.w = code | (nc << 8); };
  if (k == VAEL_NODE_LEAF_POS || k == VAEL_NODE_LEAF_NEG) {
    op.y = d->node_arg[n] * PITCH;
    kind(k == VAEL_NODE_LEAF_POS ? OP_LEAF_POS : OP_LEAF_NEG);
  } else if (k == VAEL_NODE_CONST) {
    const float v = d->const_val[d->node_arg[n]];
    memcpy(&op.y, &v, 4);
    kind(OP_CONST);
  } else if (nc == 0) {
    kind(k == VAEL_NODE_SUM ? OP_IDENT_ZERO : OP_IDENT_ONE);  // empty product / complement of nothing = 1, empty sum = 0
  } else if (nc == 1) {
    op.y = off(d->child_idx[cb]);
    kind(k == VAEL_NODE_COMPL ? OP_COMPL1 : OP_COPY);
  } else if (nc == 2) {
    op.y = off(d->child_idx[cb]);
    op.z = off(d->child_idx[cb + 1]);
    kind(k == VAEL_NODE_PROD ? OP_MUL2 : (k == VAEL_NODE_SUM ? OP_ADD2 : OP_COMPL2));
  } else {
    op.y = (int)edges.size();
    for (int e = 0; e < nc; ++e) edges.push_back(off(d->child_idx[cb + e]));
    kind(k == VAEL_NODE_PROD ? OP_PRODN : (k == VAEL_NODE_SUM ? OP_SUMN : OP_COMPLN));
  }
  return op;
}

GenProgram* gen_program_create(const vael_circuit_desc_t* d, cudaError_t* err) {
  GenProgram* g = new GenProgram();
  memset(g, 0, sizeof(*g));
  g->n_nodes = d->n_nodes; g->n_facts = d->n_facts; g->n_worlds = d->n_worlds; g->n_queries = d->n_queries;
  g->n_levels = d->n_levels; g->semiring = d->semiring; g->norm_node = d->norm_node < 0 ? -1 : d->norm_node;
  const bool is_log = d->semiring == VAEL_SEMIRING_LOG;
  const int N = d->n_nodes;
  std::vector<int> level_of(N, 0);
  for (int L = 0; L < d->n_levels; ++L)
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n) level_of[n] = L;
  auto by_kind = [](std::vector<int4>& v, int beg, int end) {
    std::stable_sort(v.begin() + beg, v.begin() + end, [](const int4& a, const int4& b) { return a.w < b.w; });
  };

  // ---- forward schedule ----
  std::vector<int4> ops;
  std::vector<int32_t> edges;
  for (int L = 0; L < d->n_levels; ++L) {
    g->op_level_ptr[L] = (int)ops.size();
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n)
      ops.push_back(make_op(d, n, [](int x) { return x * PITCH; }, edges));
    by_kind(ops, g->op_level_ptr[L], (int)ops.size());
  }
  g->op_level_ptr[d->n_levels] = (int)ops.size();
  std::vector<int32_t> world_off(d->n_worlds), query_off(d->n_queries);
  for (int w = 0; w < d->n_worlds; ++w) world_off[w] = d->world_node[w] * PITCH;
  for (int q = 0; q < d->n_queries; ++q) query_off[q] = d->query_node[q] * PITCH;

  // ---- which node values does the adjoint need?  children of products (siblings), everything under the
  // normaliser, and for the log semiring sums and their children (softmax weights) ----
  std::vector<char> need(N, 0);
  for (int p = 0; p < N; ++p) {
    const int k = d->node_kind[p];
    const int nc = d->child_ptr[p + 1] - d->child_ptr[p];
    if (!is_log) {
      if (k == VAEL_NODE_PROD && nc >= 2)
        for (int e = d->child_ptr[p]; e < d->child_ptr[p + 1]; ++e) need[d->child_idx[e]] = 1;
    } else if (k == VAEL_NODE_SUM) {
      need[p] = 1;
      for (int e = d->child_ptr[p]; e < d->child_ptr[p + 1]; ++e) need[d->child_idx[e]] = 1;
    }
  }
  if (g->norm_node >= 0) need[g->norm_node] = 1;
  for (int n = N - 1; n >= 0; --n)  // closure: a kept node needs its children
    if (need[n])
      for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) need[d->child_idx[e]] = 1;
  std::vector<int32_t> bslot(N, -1);
  for (int n = 0; n < N; ++n)
    if (need[n]) bslot[n] = g->n_bvals++;
  std::vector<int4> bops;
  std::vector<int32_t> bedges;
  for (int L = 0; L < d->n_levels; ++L) {
    g->bop_level_ptr[L] = (int)bops.size();
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n)
      if (need[n]) bops.push_back(make_op(d, n, [&](int x) { return bslot[x] * PITCH; }, bedges));
    by_kind(bops, g->bop_level_ptr[L], (int)bops.size());
  }
  g->bop_level_ptr[d->n_levels] = (int)bops.size();
  g->norm_slot_off = g->norm_node >= 0 ? bslot[g->norm_node] * PITCH : -1;

  // ---- liveness: a node carries an adjoint iff it is an output (world / query node) or feeds a live node ----
  std::vector<char> seed(N, 0), live(N, 0), is_world(N, 0);
  for (int w = 0; w < d->n_worlds; ++w) { seed[d->world_node[w]] = 1; is_world[d->world_node[w]] = 1; }
  for (int q = 0; q < d->n_queries; ++q) seed[d->query_node[q]] = 1;
  for (int n = N - 1; n >= 0; --n) {
    if (seed[n]) live[n] = 1;
    if (live[n])
      for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) live[d->child_idx[e]] = 1;
  }
  // parent lists of the live part
  std::vector<std::vector<std::pair<int, int>>> parents(N);  // child -> {parent, position in parent}
  for (int p = 0; p < N; ++p)
    if (live[p])
      for (int e = d->child_ptr[p]; e < d->child_ptr[p + 1]; ++e)
        parents[d->child_idx[e]].push_back({p, e - d->child_ptr[p]});

  std::vector<int4> crec;
  std::vector<int2> prec;
  for (int L = 0; L < d->n_levels; ++L) {
    g->cl_ptr[L] = (int)crec.size();
    for (int c = d->level_ptr[L]; c < d->level_ptr[L + 1]; ++c) {
      if (!live[c] || parents[c].empty()) continue;
      int4 cr = make_int4(c * PITCH, (int)prec.size(), (int)parents[c].size(), seed[c] ? 1 : 0);
      for (auto& pr : parents[c]) {
        const int p = pr.first, pos = pr.second;
        const int k = d->node_kind[p], cb = d->child_ptr[p], nc = d->child_ptr[p + 1] - cb;
        const int padj = (p * PITCH) << 3;
        if (!is_log) {
          if (k == VAEL_NODE_SUM) prec.push_back(make_int2(padj | PR_ADD, 0));
          else if (k == VAEL_NODE_COMPL) prec.push_back(make_int2(padj | PR_SUB, 0));
          else if (nc == 1) prec.push_back(make_int2(padj | PR_ADD, 0));
          else if (nc == 2) prec.push_back(make_int2(padj | PR_MUL_SIB, bslot[d->child_idx[cb + 1 - pos]] * PITCH));
          else {
            prec.push_back(make_int2(padj | PR_MUL_SIBS, 0));
            prec.push_back(make_int2((int)bedges.size(), nc | (pos << 16)));
            for (int e = 0; e < nc; ++e) bedges.push_back(bslot[d->child_idx[cb + e]] * PITCH);
            cr.w |= 2;
          }
        } else {
          if (k == VAEL_NODE_PROD) prec.push_back(make_int2(padj | PR_ADD, 0));
          else {
            prec.push_back(make_int2(padj | PR_LSE, bslot[c] * PITCH));
            prec.push_back(make_int2(bslot[p] * PITCH, 0));
            cr.w |= 2;
          }
        }
      }
      crec.push_back(cr);
    }
    // children with the same shape next to each other -> homogeneous groups of four
    std::stable_sort(crec.begin() + g->cl_ptr[L], crec.end(), [](const int4& a, const int4& b) {
      return a.z != b.z ? a.z < b.z : a.w < b.w;
    });
  }
  g->cl_ptr[d->n_levels] = (int)crec.size();

  std::vector<int32_t> qseed(d->n_queries), zero_off;
  for (int q = 0; q < d->n_queries; ++q) {
    qseed[q] = d->query_node[q] * PITCH;
    if (!is_world[d->query_node[q]] &&
        std::find(zero_off.begin(), zero_off.end(), qseed[q]) == zero_off.end())
      zero_off.push_back(qseed[q]);
  }
  g->n_zero = (int)zero_off.size();
  std::vector<int32_t> leaf_ptr(d->n_facts + 1, 0);
  std::vector<int2> leaf_rec;
  for (int f = 0; f < d->n_facts; ++f) {
    for (int n = 0; n < d->level_ptr[1]; ++n) {
      const int k = d->node_kind[n];
      if ((k == VAEL_NODE_LEAF_POS || k == VAEL_NODE_LEAF_NEG) && d->node_arg[n] == f && live[n] &&
          (seed[n] || !parents[n].empty()))
        leaf_rec.push_back(make_int2(n * PITCH, k == VAEL_NODE_LEAF_NEG ? -1 : 1));
    }
    leaf_ptr[f + 1] = (int32_t)leaf_rec.size();
  }

  g->fwd_smem = (size_t)(d->n_facts + N) * PITCH * sizeof(float);
  g->bwd_smem = (size_t)(d->n_facts + g->n_bvals + N) * PITCH * sizeof(float);

  cudaError_t e = cudaSuccess;
  e = up(&g->ops, ops, e); e = up(&g->edges, edges, e); e = up(&g->world_off, world_off, e);
  e = up(&g->query_off, query_off, e); e = up(&g->bops, bops, e); e = up(&g->bedges, bedges, e);
  e = up(&g->crec, crec, e); e = up(&g->prec, prec, e); e = up(&g->qseed_off, qseed, e);
  e = up(&g->zero_off, zero_off, e); e = up(&g->leaf_rec, leaf_rec, e); e = up(&g->leaf_ptr, leaf_ptr, e);
  if (err) *err = e;
  if (e != cudaSuccess) {
    gen_program_destroy(g);
    return nullptr;
  }
  return g;
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
bool gen_fits(const GenProgram* g, bool backward) {
  return (backward ? g->bwd_smem : g->fwd_smem) <= (size_t)227 * 1024;
}

template <typename K>
static cudaError_t launch_generic(K kern, const GenProgram* g, const GenCall& C, size_t smem, int num_sms,
                                  cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // CTAs per SM by shared memory; warps per CTA so that an SM holds ~24 warps (latency hiding without
  // tensor-core-style pipelining: every warp is an independent interpreter of its share of a level)
  int per_sm = (int)std::min<size_t>(16, ((size_t)227 * 1024) / (smem + 1024));
  per_sm = std::max(1, per_sm);
  static const int want = [] { const char* v = getenv("VAEL_GEN_WARPS_PER_SM"); return (v && *v) ? atoi(v) : 24; }();
  int nw = std::max(1, std::min(16, (want + per_sm - 1) / per_sm));
  const long long n_tiles = (C.B + kWarp - 1) / kWarp;
  const long long grid = std::min<long long>(n_tiles, (long long)num_sms * per_sm);
  GenArgs A;
  A.G = *g;
  A.C = C;
  kern<<<(unsigned)grid, nw * 32, smem, st>>>(A);
  return cudaGetLastError();
}

cudaError_t gen_forward(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st) {
  if (g->semiring == VAEL_SEMIRING_LOG)
    return launch_generic(generic_forward_kernel<true>, g, C, g->fwd_smem, num_sms, st);
  return launch_generic(generic_forward_kernel<false>, g, C, g->fwd_smem, num_sms, st);
}
cudaError_t gen_backward(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st) {
  if (g->semiring == VAEL_SEMIRING_LOG)
    return launch_generic(generic_backward_kernel<true>, g, C, g->bwd_smem, num_sms, st);
  return launch_generic(generic_backward_kernel<false>, g, C, g->bwd_smem, num_sms, st);
}

}  // namespace vael
