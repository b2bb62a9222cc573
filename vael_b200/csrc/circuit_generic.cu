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
constexpr int PB = PITCH * 4;  // bytes per node row; every offset in the schedule is a BYTE offset into the tile

enum : int {
  OP_LEAF_POS = 0, OP_LEAF_NEG, OP_CONST, OP_MUL2, OP_ADD2, OP_COMPL2, OP_COPY, OP_COMPL1, OP_PRODN, OP_SUMN,
  OP_COMPLN, OP_IDENT_ONE, OP_IDENT_ZERO
};
// parent-record kinds of the gather-form adjoint
enum : int { PR_ADD = 0, PR_SUB, PR_MUL_SIB, PR_MUL_SIBS, PR_LSE };

// ---- device-resident schedule ------------------------------------------------
struct GenProgram {
  int32_t n_nodes, n_facts, n_worlds, n_queries, n_levels, semiring, norm_node, n_bvals;
  // forward.  The value tile has n_nodes + 2 rows: the last two hold the constants 0 and 1 the padded
  // child lists point at.
  int4* ops;               // {dst, a, b, kind | nc << 8}; n-ary ops: a = first edge (multiple of 4), b = padded length
  int32_t* edges;          // child offsets of the n-ary ops
  int32_t op_level_ptr[kMaxLevels + 1];
  int32_t* world_off;      // [W]
  int32_t* query_off;      // [Q]
  // backward: forward values of the nodes the adjoint needs (slot space, + the two constant rows), then the
  // reverse sweep over the adjoint tile (n_nodes rows + one scratch row for the padding records)
  int4* bops;
  int32_t* bedges;
  int32_t bop_level_ptr[kMaxLevels + 1];
  int32_t norm_slot_off;   // byte offset of Z in the slot space, or -1
  int4* crec;              // live children, level by level: {adj off, first record, n records, flags}
  int32_t cl_ptr[kMaxLevels + 1];
  int2* prec;              // parent records {parent adj off << 1 | kind, x} (+ one more int2 for kinds 3, 4)
  int32_t* wseed_off;      // [W] adjoint offsets of the world nodes
  int32_t* qseed_off;      // [Q] adjoint offsets of the query nodes
  int32_t* zero_off;       // adjoint rows to clear per tile (query nodes that are not world nodes)
  int32_t n_zero;
  int2* leaf_rec;          // per fact column: live leaves {adj off, +1 / -1}
  int32_t* leaf_ptr;       // [F+1]
  size_t fwd_smem, bwd_smem;
};
// crec flags
constexpr int CR_SEEDED = 1, CR_WIDE = 2, CR_UNIFORM_SHIFT = 2;  // bits 2.. : 1 + kind when every record has that kind

struct GenArgs {
  GenProgram G;
  GenCall C;
};

// value of row `idx` (lane or batch row) of the node at byte offset `off`
__device__ __forceinline__ float& VB(float* base, int off, int idx) {
  return *reinterpret_cast<float*>(reinterpret_cast<char*>(base + idx) + off);
}
__device__ __forceinline__ float VB(const float* base, int off, int idx) {
  return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(base + idx) + off);
}

// ---- forward op execution (any kind; the grouped fast paths are in run_level) ------------------
template <bool LOG>
__device__ __forceinline__ float exec_op(const int4 d, const float* fs, const float* val, const int32_t* edges,
                                         int lane) {
  const int kind = d.w & 0xff, nc = d.w >> 8;
  switch (kind) {
    case OP_LEAF_POS: {
      const float p = VB(fs, d.y, lane);
      return LOG ? logf(p) : p;
    }
    case OP_LEAF_NEG: {
      const float p = __fsub_rn(1.0f, VB(fs, d.y, lane));
      return LOG ? logf(p) : p;
    }
    case OP_CONST: return __int_as_float(d.y);
    case OP_MUL2: {
      const float x = VB(val, d.y, lane), y = VB(val, d.z, lane);
      return LOG ? __fadd_rn(x, y) : __fmul_rn(x, y);
    }
    case OP_COPY: return VB(val, d.y, lane);
    case OP_COMPL1: return __fsub_rn(1.0f, VB(val, d.y, lane));
    case OP_IDENT_ONE: return LOG ? 0.0f : 1.0f;
    case OP_IDENT_ZERO: return LOG ? -INFINITY : 0.0f;
    case OP_PRODN: {
      float acc = VB(val, __ldg(edges + d.y), lane);
      for (int e = 1; e < nc; ++e) {
        const float v = VB(val, __ldg(edges + d.y + e), lane);
        acc = LOG ? __fadd_rn(acc, v) : __fmul_rn(acc, v);
      }
      return acc;
    }
    default: break;
  }
  // sums (ADD2 / SUMN) and complements (COMPL2 / COMPLN)
  if constexpr (!LOG) {
    float acc;
    if (kind == OP_ADD2 || kind == OP_COMPL2) {
      acc = __fadd_rn(VB(val, d.y, lane), VB(val, d.z, lane));
    } else {
      acc = VB(val, __ldg(edges + d.y), lane);
      for (int e = 1; e < nc; ++e) acc = __fadd_rn(acc, VB(val, __ldg(edges + d.y + e), lane));
    }
    return (kind == OP_COMPL2 || kind == OP_COMPLN) ? __fsub_rn(1.0f, acc) : acc;
  } else {  // logsumexp (COMPL is rejected for the log semiring at circuit creation)
    float mx = -INFINITY;
    for (int e = 0; e < nc; ++e) {
      const int c = (kind == OP_ADD2) ? (e == 0 ? d.y : d.z) : __ldg(edges + d.y + e);
      mx = fmaxf(mx, VB(val, c, lane));
    }
    if (mx == -INFINITY) return -INFINITY;
    float s = 0.0f;
    for (int e = 0; e < nc; ++e) {
      const int c = (kind == OP_ADD2) ? (e == 0 ? d.y : d.z) : __ldg(edges + d.y + e);
      s += expf(VB(val, c, lane) - mx);
    }
    return mx + logf(s);
  }
}

// One level of ops.  Warps take the ops in groups of four (levels are padded to a multiple of four with copies
// of their last op: recomputing a node is harmless).  A group whose four ops have the same kind runs branch-free:
// eight independent shared-memory loads, four FP ops, four stores for binary ops; four left folds side by side
// for n-ary ones (each fold sequential in CSR order, so the result is bit-exact; the four child lists of such a
// group are padded to one length with pointers to the constant-0 / constant-1 rows).
template <bool LOG>
__device__ __forceinline__ void run_level(const int4* __restrict__ ops, const int32_t* __restrict__ edges, int beg,
                                          int end, const float* fs, float* val, int warp, int nwarps, int lane) {
  for (int g = beg + 4 * warp; g < end; g += 4 * nwarps) {
    int4 d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u] = __ldg(ops + g + u);
    const int k0 = d[0].w & 0xff;
    const bool same = (d[1].w & 0xff) == k0 && (d[2].w & 0xff) == k0 && (d[3].w & 0xff) == k0;
    if (same && k0 == OP_MUL2) {
      float x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[u] = VB(val, d[u].y, lane); y[u] = VB(val, d[u].z, lane); }
#pragma unroll
      for (int u = 0; u < 4; ++u) VB(val, d[u].x, lane) = LOG ? __fadd_rn(x[u], y[u]) : __fmul_rn(x[u], y[u]);
    } else if (!LOG && same && k0 == OP_ADD2) {
      float x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[u] = VB(val, d[u].y, lane); y[u] = VB(val, d[u].z, lane); }
#pragma unroll
      for (int u = 0; u < 4; ++u) VB(val, d[u].x, lane) = __fadd_rn(x[u], y[u]);
    } else if (same && k0 <= OP_LEAF_NEG) {
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = VB(fs, d[u].y, lane);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v = k0 == OP_LEAF_NEG ? __fsub_rn(1.0f, p[u]) : p[u];
        VB(val, d[u].x, lane) = LOG ? logf(v) : v;
      }
    } else if (same && (k0 == OP_PRODN || (!LOG && k0 == OP_SUMN))) {
      const bool adds = LOG || k0 == OP_SUMN;     // log-semiring product = sum of logs
      const int len = d[0].z;                     // group-uniform padded length (multiple of four)
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = adds ? 0.0f : 1.0f;  // 0 + x and 1 * x are exact
      for (int e = 0; e < len; e += 4) {
        int4 c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = __ldg(reinterpret_cast<const int4*>(edges + d[u].y + e));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float v0 = VB(val, c[u].x, lane), v1 = VB(val, c[u].y, lane), v2 = VB(val, c[u].z, lane),
                      v3 = VB(val, c[u].w, lane);
          if (adds) {
            acc[u] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc[u], v0), v1), v2), v3);
          } else {
            acc[u] = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(acc[u], v0), v1), v2), v3);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) VB(val, d[u].x, lane) = acc[u];
    } else {
      for (int u = 0; u < 4; ++u) VB(val, d[u].x, lane) = exec_op<LOG>(d[u], fs, val, edges, lane);
    }
  }
}

// transposed load of the facts tile: global [rows,F] row-major -> fs[f][r]; warps take rows, lanes take fact
// columns (a row is F contiguous floats, consecutive rows are contiguous: every sector is fully used)
__device__ __forceinline__ void load_facts_tile(float* fs, const float* facts, long long row0, int rows, int F,
                                                int warp, int nwarps, int lane) {
  const float* src = facts + (row0 + warp) * F + lane;
  float* dst = fs + lane * PITCH + warp;
  const int sstep = nwarps * F;
  for (int r = warp; r < kWarp; r += nwarps, src += sstep, dst += nwarps) {
    for (int f = lane, k = 0; f < F; f += kWarp, k += kWarp)
      dst[k * PITCH] = (r < rows) ? __ldg(src + k) : 0.5f;  // padding rows get a harmless in-domain value
  }
}

// the two constant rows behind the node rows (written once per CTA; nothing ever overwrites them)
__device__ __forceinline__ void init_const_rows(float* val, int n_rows) {
  if (threadIdx.x < kWarp) {
    val[n_rows * PITCH + threadIdx.x] = 0.0f;
    val[(n_rows + 1) * PITCH + threadIdx.x] = 1.0f;
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
  float* val = gsm + P.n_facts * PITCH;   // [n_nodes + 2][PITCH]
  __shared__ float es[kWarp];             // evidence mode: P(e) per row
  __shared__ int eq[kWarp];               //                query index per row (-1 = none)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long n_tiles = (C.B + kWarp - 1) / kWarp;
  const bool norm = (C.flags & VAEL_FLAG_NORMALIZE) && P.norm_node >= 0;
  const bool evid = (C.flags & VAEL_FLAG_EVIDENCE) != 0;
  const int zoff = P.norm_node * PB;
  const int W = P.n_worlds;
  init_const_rows(val, P.n_nodes);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kWarp;
    const int rows = (int)min((long long)kWarp, C.B - row0);
    load_facts_tile(fs, C.facts, row0, rows, P.n_facts, warp, nwarps, lane);
    __syncthreads();
    for (int L = 0; L < P.n_levels; ++L) {
      run_level<LOG>(P.ops, P.edges, P.op_level_ptr[L], P.op_level_ptr[L + 1], fs, val, warp, nwarps, lane);
      __syncthreads();
    }

    // ---- outputs: lane <-> world (chunks of 32 worlds), warps take rows; 128-byte coalesced stores ----
    if (C.worlds != nullptr) {
      if (evid) {  // per-row evidence value and query index, staged once per tile
        if (warp == 0) {
          const long long row = row0 + lane;
          const int32_t q = (lane < rows) ? (C.qidx ? __ldg(C.qidx + row) : C.qscalar) : -1;
          const bool qok = q >= 0 && q < P.n_queries;
          es[lane] = qok ? VB(val, __ldg(P.query_off + q), lane) : 0.0f;
          eq[lane] = qok ? q : -1;
        }
        __syncthreads();
      }
      const size_t ostep = (size_t)nwarps * W;
      for (int e0 = 0; e0 < W; e0 += kWarp) {
        const int e = e0 + lane;
        if (e < W) {
          const float* v = reinterpret_cast<const float*>(reinterpret_cast<const char*>(val) + __ldg(P.world_off + e)) + warp;
          float* o = C.worlds + (row0 + warp) * W + e;
          if (!norm && !evid) {
#pragma unroll 4
            for (int r = warp; r < rows; r += nwarps, v += nwarps, o += ostep) *o = *v;
          } else if (!evid) {
            const float* zr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(val) + zoff) + warp;
            for (int r = warp; r < rows; r += nwarps, v += nwarps, zr += nwarps, o += ostep) *o = *v / *zr;
          } else {
            for (int r = warp; r < rows; r += nwarps, v += nwarps, o += ostep) {
              const int q = eq[r];
              const uint32_t mword = q >= 0 ? __ldg(C.qmask + (long long)q * C.mask_words + (e0 >> 5)) : 0u;
              *o = ((mword >> lane) & 1u) ? *v / es[r] : 0.0f;
            }
          }
        }
      }
    }
    if (C.queries_all != nullptr) {
      for (int r = warp; r < rows; r += nwarps) {
        const long long row = row0 + r;
        const float z = norm ? VB(val, zoff, r) : 1.0f;
        for (int q = lane; q < P.n_queries; q += kWarp) {
          float v = VB(val, __ldg(P.query_off + q), r);
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
        v = VB(val, __ldg(P.query_off + q), lane);
        if (norm) v = v / VB(val, zoff, lane);
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
  const float ap = VB(adj, rec.x >> 1 & ~3, lane);
  switch (kind) {
    case PR_ADD: return a + ap;
    case PR_SUB: return a - ap;
    case PR_MUL_SIB: return fmaf(ap, VB(val, rec.y, lane), a);
    case PR_MUL_SIBS: {  // product of the other children of an n-ary product
      const int2 ext = __ldg(prec + i + 1);
      const int nc = ext.y & 0xffff, pos = ext.y >> 16;
      float sib = 1.0f;
      for (int e = 0; e < nc; ++e)
        if (e != pos) sib *= VB(val, __ldg(bedges + ext.x + e), lane);
      return fmaf(ap, sib, a);
    }
    default: {  // PR_LSE: softmax weight of the child inside a logsumexp parent
      const int2 ext = __ldg(prec + i + 1);
      const float vc = VB(val, rec.y, lane), vp = VB(val, ext.x, lane);
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
  float* val = fs + P.n_facts * PITCH;           // [n_bvals + 2][PITCH]
  float* adj = val + (P.n_bvals + 2) * PITCH;    // [n_nodes + 1][PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long n_tiles = (C.B + kWarp - 1) / kWarp;
  const bool norm = (C.flags & VAEL_FLAG_NORMALIZE) && P.norm_slot_off >= 0;
  const int W = P.n_worlds, F = P.n_facts;
  init_const_rows(val, P.n_bvals);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kWarp;
    const int rows = (int)min((long long)kWarp, C.B - row0);
    load_facts_tile(fs, C.facts, row0, rows, F, warp, nwarps, lane);
    for (int i = threadIdx.x; i < P.n_zero * kWarp; i += blockDim.x) VB(adj, __ldg(P.zero_off + (i >> 5)), i & 31) = 0.0f;
    __syncthreads();

    for (int L = 0; L < P.n_levels; ++L) {
      if (P.bop_level_ptr[L + 1] > P.bop_level_ptr[L]) {
        run_level<LOG>(P.bops, P.bedges, P.bop_level_ptr[L], P.bop_level_ptr[L + 1], fs, val, warp, nwarps, lane);
        __syncthreads();
      }
    }

    // seed the world adjoints from the upstream gradient (assignment: no zero-fill needed);
    // lane <-> world, warps take rows: 128-byte coalesced loads
    {
      const size_t gstep = (size_t)nwarps * W;
      for (int e0 = 0; e0 < W; e0 += kWarp) {
        const int e = e0 + lane;
        if (e < W) {
          float* a = reinterpret_cast<float*>(reinterpret_cast<char*>(adj) + __ldg(P.wseed_off + e)) + warp;
          if (C.grad_worlds == nullptr) {
            for (int r = warp; r < kWarp; r += nwarps, a += nwarps) *a = 0.0f;
          } else {
            const float* gsrc = C.grad_worlds + (row0 + warp) * W + e;
            if (!norm) {
#pragma unroll 4
              for (int r = warp; r < kWarp; r += nwarps, a += nwarps, gsrc += gstep) *a = (r < rows) ? __ldg(gsrc) : 0.0f;
            } else {
              for (int r = warp; r < kWarp; r += nwarps, a += nwarps, gsrc += gstep)
                *a = (r < rows) ? __ldg(gsrc) / VB(val, P.norm_slot_off, r) : 0.0f;
            }
          }
        }
      }
    }
    __syncthreads();
    if (C.grad_query != nullptr && warp == 0 && lane < rows) {
      const long long row = row0 + lane;
      const int32_t q = C.qidx ? __ldg(C.qidx + row) : C.qscalar;
      if (q >= 0 && q < P.n_queries) {
        float g = __ldg(C.grad_query + row);
        if (norm) g = g / VB(val, P.norm_slot_off, lane);
        VB(adj, __ldg(P.qseed_off + q), lane) += g;
      }
    }
    __syncthreads();

    // reverse sweep, gather form: adj[c] = seed[c] + sum over live parents p of adj[p] * d val[p] / d val[c].
    // Children come in groups of four of the same shape (same record count and kinds; levels padded with
    // records that write a scratch row), so four accumulation chains are in flight per warp.
    for (int L = P.n_levels - 2; L >= 0; --L) {
      const int beg = P.cl_ptr[L], end = P.cl_ptr[L + 1];
      for (int g = beg + 4 * warp; g < end; g += 4 * nwarps) {
        int4 cr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cr[u] = __ldg(P.crec + g + u);
        const bool same = cr[1].z == cr[0].z && cr[2].z == cr[0].z && cr[3].z == cr[0].z && cr[1].w == cr[0].w &&
                          cr[2].w == cr[0].w && cr[3].w == cr[0].w;
        if (same && !(cr[0].w & CR_WIDE)) {
          float a[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) a[u] = (cr[0].w & CR_SEEDED) ? VB(adj, cr[u].x, lane) : 0.0f;
          const int uk = cr[0].w >> CR_UNIFORM_SHIFT, n = cr[0].z;
          if (uk == 1 + PR_MUL_SIB) {
            for (int r = 0; r < n; ++r) {
              int2 rec[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) rec[u] = __ldg(P.prec + cr[u].y + r);
#pragma unroll
              for (int u = 0; u < 4; ++u) a[u] = fmaf(VB(adj, rec[u].x >> 1 & ~3, lane), VB(val, rec[u].y, lane), a[u]);
            }
          } else if (uk == 1 + PR_ADD) {
            for (int r = 0; r < n; ++r) {
#pragma unroll
              for (int u = 0; u < 4; ++u) a[u] += VB(adj, __ldg(&P.prec[cr[u].y + r].x) >> 1 & ~3, lane);
            }
          } else {
            for (int r = 0; r < n; ++r) {
#pragma unroll
              for (int u = 0; u < 4; ++u) a[u] = apply_rec<LOG>(a[u], P.prec, cr[u].y + r, val, adj, P.bedges, lane);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) VB(adj, cr[u].x, lane) = a[u];
        } else {
          for (int u = 0; u < 4; ++u) {
            float a = (cr[u].w & CR_SEEDED) ? VB(adj, cr[u].x, lane) : 0.0f;
            for (int r = 0, i = cr[u].y; r < cr[u].z; ++r) {
              const int k = __ldg(&P.prec[i].x) & 7;
              a = apply_rec<LOG>(a, P.prec, i, val, adj, P.bedges, lane);
              i += (k == PR_MUL_SIBS || k == PR_LSE) ? 2 : 1;
            }
            VB(adj, cr[u].x, lane) = a;
          }
        }
      }
      __syncthreads();
    }

    // live leaves -> grad_facts[row, f]: lanes take fact columns, warps take rows
    for (int f = lane; f < F; f += kWarp) {
      const int kb = __ldg(P.leaf_ptr + f), ke = __ldg(P.leaf_ptr + f + 1);
      float* o = C.grad_facts + (row0 + warp) * F + f;
      const size_t ostep = (size_t)nwarps * F;
      if (!LOG && ke - kb == 1) {  // the usual case: one (positive) literal per fact
        const int2 lr = __ldg(P.leaf_rec + kb);
        const float* a = reinterpret_cast<const float*>(reinterpret_cast<const char*>(adj) + lr.x) + warp;
        const float sgn = lr.y < 0 ? -1.0f : 1.0f;
#pragma unroll 4
        for (int r = warp; r < rows; r += nwarps, a += nwarps, o += ostep) *o = sgn * *a;
      } else {
        for (int r = warp; r < rows; r += nwarps, o += ostep) {
          float g = 0.0f;
          for (int k = kb; k < ke; ++k) {
            const int2 lr = __ldg(P.leaf_rec + k);
            const float a = VB(adj, lr.x, r);
            if (!LOG) {
              g += lr.y < 0 ? -a : a;
            } else {
              const float pf = fs[f * PITCH + r];
              g += lr.y < 0 ? -a / (1.0f - pf) : a / pf;
            }
          }
          *o = g;
        }
      }
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
  cudaFree(g->bedges); cudaFree(g->crec); cudaFree(g->prec); cudaFree(g->wseed_off); cudaFree(g->qseed_off);
  cudaFree(g->zero_off); cudaFree(g->leaf_rec); cudaFree(g->leaf_ptr);
  delete g;
}

namespace {
struct HostOp {
  int4 op;                    // dst / a / b filled; n-ary: a, b set when the level is laid out
  std::vector<int32_t> kids;  // n-ary child offsets
};

// One op per node; `off(n)` maps a node id to the byte offset of its value row.
template <typename OffFn>
HostOp make_op(const vael_circuit_desc_t* d, int n, OffFn off) {
  const int k = d->node_kind[n];
  const int cb = d->child_ptr[n], nc = d->child_ptr[n + 1] - cb;
  HostOp h;
  h.op = make_int4(off(n), 0, 0, 0);
  auto kind = [&](int code) { h.op.w = code | (nc << 8); };
  if (k == VAEL_NODE_LEAF_POS || k == VAEL_NODE_LEAF_NEG) {
    h.op.y = d->node_arg[n] * PB;
    kind(k == VAEL_NODE_LEAF_POS ? OP_LEAF_POS : OP_LEAF_NEG);
  } else if (k == VAEL_NODE_CONST) {
    const float v = d->const_val[d->node_arg[n]];
    memcpy(&h.op.y, &v, 4);
    kind(OP_CONST);
  } else if (nc == 0) {
    kind(k == VAEL_NODE_SUM ? OP_IDENT_ZERO : OP_IDENT_ONE);  // empty product / complement of nothing = 1, empty sum = 0
  } else if (nc == 1) {
    h.op.y = off(d->child_idx[cb]);
    kind(k == VAEL_NODE_COMPL ? OP_COMPL1 : OP_COPY);
  } else if (nc == 2) {
    h.op.y = off(d->child_idx[cb]);
    h.op.z = off(d->child_idx[cb + 1]);
    kind(k == VAEL_NODE_PROD ? OP_MUL2 : (k == VAEL_NODE_SUM ? OP_ADD2 : OP_COMPL2));
  } else {
    for (int e = 0; e < nc; ++e) h.kids.push_back(off(d->child_idx[cb + e]));
    kind(k == VAEL_NODE_PROD ? OP_PRODN : (k == VAEL_NODE_SUM ? OP_SUMN : OP_COMPLN));
  }
  return h;
}

// Sort the ops of one level by kind (then fan-in), pad the level to a multiple of four, lay out the child lists:
// every list starts at a multiple of four; the four lists of a group of same-kind n-ary ops are padded to one
// length with pointers to the constant row that is the identity of the fold (zero_off / one_off).
void emit_level(std::vector<HostOp>& lvl, bool is_log, int zero_off, int one_off, std::vector<int4>& ops,
                std::vector<int32_t>& edges) {
  if (lvl.empty()) return;
  std::stable_sort(lvl.begin(), lvl.end(), [](const HostOp& a, const HostOp& b) {
    return (a.op.w & 0xff) != (b.op.w & 0xff) ? (a.op.w & 0xff) < (b.op.w & 0xff) : (a.op.w >> 8) < (b.op.w >> 8);
  });
  while (lvl.size() % 4) lvl.push_back(lvl.back());
  for (size_t g = 0; g < lvl.size(); g += 4) {
    const int k0 = lvl[g].op.w & 0xff;
    bool same = true;
    size_t len = 0;
    for (int u = 0; u < 4; ++u) {
      same = same && (lvl[g + u].op.w & 0xff) == k0;
      len = std::max(len, (lvl[g + u].kids.size() + 3) / 4 * 4);
    }
    const bool fold = same && (k0 == OP_PRODN || (!is_log && k0 == OP_SUMN));
    for (int u = 0; u < 4; ++u) {
      HostOp& h = lvl[g + u];
      if (!h.kids.empty()) {
        const int code = h.op.w & 0xff;
        const int pad = (code == OP_PRODN && !is_log) ? one_off : zero_off;
        const size_t mylen = fold ? len : (h.kids.size() + 3) / 4 * 4;
        h.op.y = (int)edges.size();
        h.op.z = (int)mylen;
        for (size_t e = 0; e < mylen; ++e) edges.push_back(e < h.kids.size() ? h.kids[e] : pad);
      }
      ops.push_back(h.op);
    }
  }
}
}  // namespace

GenProgram* gen_program_create(const vael_circuit_desc_t* d, cudaError_t* err) {
  GenProgram* g = new GenProgram();
  memset(g, 0, sizeof(*g));
  g->n_nodes = d->n_nodes; g->n_facts = d->n_facts; g->n_worlds = d->n_worlds; g->n_queries = d->n_queries;
  g->n_levels = d->n_levels; g->semiring = d->semiring; g->norm_node = d->norm_node < 0 ? -1 : d->norm_node;
  const bool is_log = d->semiring == VAEL_SEMIRING_LOG;
  const int N = d->n_nodes;

  // ---- forward schedule ----
  std::vector<int4> ops;
  std::vector<int32_t> edges;
  for (int L = 0; L < d->n_levels; ++L) {
    g->op_level_ptr[L] = (int)ops.size();
    std::vector<HostOp> lvl;
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n) lvl.push_back(make_op(d, n, [](int x) { return x * PB; }));
    emit_level(lvl, is_log, N * PB, (N + 1) * PB, ops, edges);
  }
  g->op_level_ptr[d->n_levels] = (int)ops.size();
  std::vector<int32_t> world_off(d->n_worlds), query_off(d->n_queries);
  for (int w = 0; w < d->n_worlds; ++w) world_off[w] = d->world_node[w] * PB;
  for (int q = 0; q < d->n_queries; ++q) query_off[q] = d->query_node[q] * PB;

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
    std::vector<HostOp> lvl;
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n)
      if (need[n]) lvl.push_back(make_op(d, n, [&](int x) { return bslot[x] * PB; }));
    emit_level(lvl, is_log, g->n_bvals * PB, (g->n_bvals + 1) * PB, bops, bedges);
  }
  g->bop_level_ptr[d->n_levels] = (int)bops.size();
  g->norm_slot_off = g->norm_node >= 0 ? bslot[g->norm_node] * PB : -1;

  // ---- liveness: a node carries an adjoint iff it is an output (world / query node) or feeds a live node ----
  std::vector<char> seed(N, 0), live(N, 0), is_world(N, 0);
  for (int w = 0; w < d->n_worlds; ++w) { seed[d->world_node[w]] = 1; is_world[d->world_node[w]] = 1; }
  for (int q = 0; q < d->n_queries; ++q) seed[d->query_node[q]] = 1;
  for (int n = N - 1; n >= 0; --n) {
    if (seed[n]) live[n] = 1;
    if (live[n])
      for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) live[d->child_idx[e]] = 1;
  }
  std::vector<std::vector<std::pair<int, int>>> parents(N);  // child -> {live parent, position in parent}
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
      int4 cr = make_int4(c * PB, (int)prec.size(), (int)parents[c].size(), seed[c] ? CR_SEEDED : 0);
      int uniform = -1;
      bool first = true;
      for (auto& pr : parents[c]) {
        const int p = pr.first, pos = pr.second;
        const int k = d->node_kind[p], cb = d->child_ptr[p], nc = d->child_ptr[p + 1] - cb;
        const int padj = (p * PB) << 1;
        int kind;
        if (!is_log) {
          if (k == VAEL_NODE_SUM) { kind = PR_ADD; prec.push_back(make_int2(padj | kind, 0)); }
          else if (k == VAEL_NODE_COMPL) { kind = PR_SUB; prec.push_back(make_int2(padj | kind, 0)); }
          else if (nc == 1) { kind = PR_ADD; prec.push_back(make_int2(padj | kind, 0)); }
          else if (nc == 2) {
            kind = PR_MUL_SIB;
            prec.push_back(make_int2(padj | kind, bslot[d->child_idx[cb + 1 - pos]] * PB));
          } else {
            kind = PR_MUL_SIBS;
            prec.push_back(make_int2(padj | kind, 0));
            prec.push_back(make_int2((int)bedges.size(), nc | (pos << 16)));
            for (int e = 0; e < nc; ++e) bedges.push_back(bslot[d->child_idx[cb + e]] * PB);
            cr.w |= CR_WIDE;
          }
        } else {
          if (k == VAEL_NODE_PROD) { kind = PR_ADD; prec.push_back(make_int2(padj | kind, 0)); }
          else {
            kind = PR_LSE;
            prec.push_back(make_int2(padj | kind, bslot[c] * PB));
            prec.push_back(make_int2(bslot[p] * PB, 0));
            cr.w |= CR_WIDE;
          }
        }
        uniform = first ? kind : (uniform == kind ? kind : -2);
        first = false;
      }
      if (uniform >= 0) cr.w |= (1 + uniform) << CR_UNIFORM_SHIFT;
      crec.push_back(cr);
    }
    // children with the same shape next to each other -> homogeneous groups of four; pad with scratch writers
    std::stable_sort(crec.begin() + g->cl_ptr[L], crec.end(), [](const int4& a, const int4& b) {
      return a.w != b.w ? a.w < b.w : a.z < b.z;
    });
    while ((crec.size() - g->cl_ptr[L]) % 4) crec.push_back(make_int4(N * PB, 0, 0, 0));
  }
  g->cl_ptr[d->n_levels] = (int)crec.size();

  std::vector<int32_t> wseed(d->n_worlds), qseed(d->n_queries), zero_off;
  for (int w = 0; w < d->n_worlds; ++w) wseed[w] = d->world_node[w] * PB;
  for (int q = 0; q < d->n_queries; ++q) {
    qseed[q] = d->query_node[q] * PB;
    if (!is_world[d->query_node[q]] && std::find(zero_off.begin(), zero_off.end(), qseed[q]) == zero_off.end())
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
        leaf_rec.push_back(make_int2(n * PB, k == VAEL_NODE_LEAF_NEG ? -1 : 1));
    }
    leaf_ptr[f + 1] = (int32_t)leaf_rec.size();
  }

  g->fwd_smem = (size_t)(d->n_facts + N + 2) * PB;
  g->bwd_smem = (size_t)(d->n_facts + g->n_bvals + 2 + N + 1) * PB;

  cudaError_t e = cudaSuccess;
  e = up(&g->ops, ops, e); e = up(&g->edges, edges, e); e = up(&g->world_off, world_off, e);
  e = up(&g->query_off, query_off, e); e = up(&g->bops, bops, e); e = up(&g->bedges, bedges, e);
  e = up(&g->crec, crec, e); e = up(&g->prec, prec, e); e = up(&g->wseed_off, wseed, e);
  e = up(&g->qseed_off, qseed, e); e = up(&g->zero_off, zero_off, e); e = up(&g->leaf_rec, leaf_rec, e);
  e = up(&g->leaf_ptr, leaf_ptr, e);
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
