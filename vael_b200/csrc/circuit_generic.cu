// Generic evaluator: level-by-level sweep over the flat, level-ordered CSR circuit.
//
// Reference semantics: problog's sdd.evaluate(semiring=GraphSemiring) walking the compiled
// formula node by node and calling GraphSemiring.plus/times/negate/value/normalize for every
// node (utils/graph_semiring.py:13-60, models/vael.py:98-135), evaluated here ELEMENTWISE per
// batch row (the reference's batch-wide `.any()` identity shortcuts, utils/graph_semiring.py:27-49,
// are a documented divergence: see DESIGN.md "Parity domain").
//
// Execution model.  A CTA owns a tile of T = 32 R batch rows (R = 1, 2 or 4, chosen per circuit from its
// shared-memory footprint).  A lane holds R rows (lane, lane + 32, ...) as one float / float2 / float4, so
// node values live in shared memory node-major as val[node][lane] vectors: every warp-wide access is one
// conflict-free LDS/STS of 32 consecutive vectors whatever the circuit shape, and one decoded op does R
// rows' worth of arithmetic per lane — the interpreter's decode cost is amortised over 32 R rows.  The
// circuit is turned ONCE, at vael_circuit_create, into a schedule: one 16-byte op per node with the
// children already converted to shared-memory byte offsets, ops sorted by kind inside each level and every
// run of one kind described as a segment.  The kind is decoded once per segment; the segment's ops are swept
// by a tight loop specialised for that kind, G = 4 / R ops per warp at a time (binary products / sums: 2 G
// independent shared-memory loads, G R FP ops, G stores), and the warps meet at one barrier per level.  A
// small schedule is staged in shared memory once per CTA and read through side-effect-free ld.shared (Tab).
//
// Batch-major global tensors never meet the node-major layout in global memory: a [T, K] tile of a
// row-major [B, K] tensor is one contiguous span, moved by ONE bulk copy (cp.async.bulk, SASS UBLKCP)
// between HBM and a row-major staging tile in shared memory (row-by-row bulk copies for column chunks of
// wide tensors), loads completing on an mbarrier and prefetched one tile ahead, stores draining while the
// next tile computes.  The row-major <-> node-major transposition happens shared-to-shared with 128-bit
// accesses (lane <-> row on both sides).  Unaligned bases, K % 4 != 0 chunks and the tail tile take a
// plain coalesced copy through the same staging tile.
//
// The backward is the reverse sweep in GATHER form over per-child parent records (built at
// creation, dead cones — nodes no output depends on — pruned), so it needs no atomics, no
// zero-fill of the adjoint tile, and is deterministic.
#include <algorithm>
#include <cstring>
#include <vector>

#include "params.cuh"

namespace vael {

enum : int {
  OP_LEAF_POS = 0, OP_LEAF_NEG, OP_CONST, OP_MUL2, OP_ADD2, OP_COMPL2, OP_COPY, OP_COMPL1, OP_PRODN, OP_SUMN,
  OP_COMPLN, OP_IDENT_ONE, OP_IDENT_ZERO
};
// parent-record kinds of the gather-form adjoint
enum : int { PR_ADD = 0, PR_SUB, PR_MUL_SIB, PR_MUL_SIBS, PR_LSE };

// ---- device-resident schedule ------------------------------------------------
// Every offset in the schedule is a BYTE offset of a node row inside a value / adjoint tile whose rows are
// 128 R bytes; the forward and the backward kernel may run with different R (rf, rb).
struct GenProgram {
  int32_t n_nodes, n_facts, n_worlds, n_queries, n_levels, semiring, norm_node, n_bvals;
  int32_t rf, rb;          // rows per lane of the forward / backward kernel
  int32_t cf, cb;          // staged columns per chunk: worlds (forward) / grad_worlds (backward)
  int32_t cq;              // column capacity of the forward staging tile (>= cf; queries_all chunks)
  int32_t leaf_levels_f[2], leaf_levels_b[2];  // levels [0, leaf_levels) hold every leaf op of the schedule
  int32_t n_levels_f[2];   // levels of the forward schedule up to its last non-empty one
  // forward.  The value tile has n_nodes + 2 rows: the last two hold the constants 0 and 1 the padded
  // child lists point at.
  int4* ops;               // {dst, a, b, kind | nc << 8}; leaves: a = fact column; n-ary ops: a = first edge (multiple of 4), b = padded length
  int32_t* edges;          // child offsets of the n-ary ops
  int4* segs;              // runs of equal op kind inside a level: {kind, first op, end, 0}
  int32_t n_ops, n_edges, n_segs, sched_in_smem;  // forward schedule sizes; staged in shared memory when small
  int32_t seg_level_ptr[2][kMaxLevels + 1];  // [0]: every node; [1]: without the cone only the normaliser needs
  int32_t* world_off;      // [W] (padded to a multiple of 4)
  int32_t* query_off;      // [Q]
  // backward: forward values of the nodes the adjoint needs (slot space, + the two constant rows), then the
  // reverse sweep over the adjoint tile (n_nodes rows + one scratch row for the padding records)
  int4* bops;
  int32_t* bedges;
  int4* bsegs;
  int32_t n_bops, n_bedges, n_bsegs, n_crec, n_csegs, n_prec, bsched_in_smem;  // backward schedule sizes
  int32_t bseg_level_ptr[2][kMaxLevels + 1];  // [0]: with the normaliser's cone (NORMALIZE calls); [1]: without
  int32_t norm_slot_off;   // byte offset of Z in the slot space, or -1
  int4* crec;              // live children, level by level: {adj off, first record, n records, flags}
  int4* csegs;             // runs of children of the same shape inside a level: {flags, first child, end, n records}
  int32_t cseg_level_ptr[kMaxLevels + 1];
  int2* prec;              // parent records {parent adj off << 1 | kind, x} (+ one more int2 for kinds 3, 4)
  int32_t* wseed_off;      // [W] adjoint offsets of the world nodes (padded to a multiple of 4)
  int32_t* qseed_off;      // [Q] adjoint offsets of the query nodes, | 1 when the node is also a world node
  int32_t* qonly_off;      // adjoint rows of the query nodes that are not world nodes (seeded by assignment)
  int32_t n_qonly;
  int32_t q_world_any;     // some query node is also a world node (its seed is added after the world seeds)
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

// ---- R rows per lane ----------------------------------------------------------------------------------
template <int R> struct VecT;
template <> struct VecT<1> { float v[1]; };
template <> struct __align__(8) VecT<2> { float v[2]; };
template <> struct __align__(16) VecT<4> { float v[4]; };

// Node rows are addressed with 32-bit shared-window addresses (`lb` = tile base + lane * 4 R bytes, `off` = byte
// offset of the node row) through ld/st.shared directly: one IADD + one LDS/STS per access.
template <int R> __device__ __forceinline__ VecT<R> ldv(uint32_t lb, int off);
template <> __device__ __forceinline__ VecT<1> ldv<1>(uint32_t lb, int off) {
  VecT<1> x;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x.v[0]) : "r"(lb + off));
  return x;
}
template <> __device__ __forceinline__ VecT<2> ldv<2>(uint32_t lb, int off) {
  VecT<2> x;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x.v[0]), "=f"(x.v[1]) : "r"(lb + off));
  return x;
}
template <> __device__ __forceinline__ VecT<4> ldv<4>(uint32_t lb, int off) {
  VecT<4> x;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(x.v[0]), "=f"(x.v[1]), "=f"(x.v[2]), "=f"(x.v[3])
               : "r"(lb + off));
  return x;
}
template <int R> __device__ __forceinline__ void stv(uint32_t lb, int off, const VecT<R>& x);
template <> __device__ __forceinline__ void stv<1>(uint32_t lb, int off, const VecT<1>& x) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(lb + off), "f"(x.v[0]) : "memory");
}
template <> __device__ __forceinline__ void stv<2>(uint32_t lb, int off, const VecT<2>& x) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(lb + off), "f"(x.v[0]), "f"(x.v[1]) : "memory");
}
template <> __device__ __forceinline__ void stv<4>(uint32_t lb, int off, const VecT<4>& x) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lb + off), "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]),
               "f"(x.v[3])
               : "memory");
}
__device__ __forceinline__ float lds1(uint32_t a) {
  float x;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
  return x;
}
__device__ __forceinline__ void sts1(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
template <int R>
__device__ __forceinline__ VecT<R> vfill(float c) {
  VecT<R> o;
#pragma unroll
  for (int j = 0; j < R; ++j) o.v[j] = c;
  return o;
}
template <int R>
__device__ __forceinline__ VecT<R> vmul_rn(const VecT<R>& a, const VecT<R>& b) {
  VecT<R> o;
#pragma unroll
  for (int j = 0; j < R; ++j) o.v[j] = __fmul_rn(a.v[j], b.v[j]);
  return o;
}
template <int R>
__device__ __forceinline__ VecT<R> vadd_rn(const VecT<R>& a, const VecT<R>& b) {
  VecT<R> o;
#pragma unroll
  for (int j = 0; j < R; ++j) o.v[j] = __fadd_rn(a.v[j], b.v[j]);
  return o;
}
template <int R>
__device__ __forceinline__ VecT<R> vcompl_rn(const VecT<R>& a) {
  VecT<R> o;
#pragma unroll
  for (int j = 0; j < R; ++j) o.v[j] = __fsub_rn(1.0f, a.v[j]);
  return o;
}
template <int R>
__device__ __forceinline__ VecT<R> vfma(const VecT<R>& a, const VecT<R>& b, const VecT<R>& c) {
  VecT<R> o;
#pragma unroll
  for (int j = 0; j < R; ++j) o.v[j] = fmaf(a.v[j], b.v[j], c.v[j]);
  return o;
}

// A read-only schedule table: global memory through the non-coherent path (__ldg), or — when the schedule is small
// enough to be staged once per CTA — shared memory through ld.shared WITHOUT a memory clobber, so that the
// compiler may move the fetch across the value stores of the preceding op (the table never changes after staging).
template <bool SS, typename T>
struct Tab {
  const T* g;
  uint32_t s;
  // Construct AFTER the barrier that follows the staging copy: the shared address is passed through a volatile
  // asm there, so no (side-effect-free) table fetch can be scheduled above that point.
  __device__ __forceinline__ Tab(const T* gp, const void* sp) : g(gp), s(0u) {
    if constexpr (SS) {
      const uint32_t a = smem_u32(sp);
      asm volatile("mov.u32 %0, %1;" : "=r"(s) : "r"(a) : "memory");
    }
  }
  __device__ __forceinline__ int4 ld16(int i) const {   // 16 bytes starting at element i
    if constexpr (SS) {
      int4 v;
      asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(s + i * (int)sizeof(T)));
      return v;
    } else {
      return __ldg(reinterpret_cast<const int4*>(g + i));
    }
  }
  __device__ __forceinline__ int2 ld8(int i) const {
    if constexpr (SS) {
      int2 v;
      asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(s + i * (int)sizeof(T)));
      return v;
    } else {
      return __ldg(reinterpret_cast<const int2*>(g + i));
    }
  }
  __device__ __forceinline__ int32_t ld4(int i) const {
    if constexpr (SS) {
      int32_t v;
      asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(s + i * (int)sizeof(T)));
      return v;
    } else {
      return __ldg(reinterpret_cast<const int32_t*>(g + i));
    }
  }
};

// ---- staged I/O: [rows, cc] chunk (columns c0 .. c0+cc) of a row-major [B, K] tensor <-> stage[r * cc + c] ----
// mode 0: plain coalesced copy by all threads; 1: the chunk is the whole contiguous tile, one bulk copy;
// 2: one bulk copy per row.  The mode is CTA-uniform.
__device__ __forceinline__ int io_mode(const void* g, int rows, int T, int K, int c0, int cc) {
  if (!aligned16(g) || rows != T) return 0;
  if (cc == K) return 1;  // T * K * 4 bytes: a multiple of 128
  if (((K | c0 | cc) & 3) == 0) return 2;
  return 0;
}
// returns true when the data arrives asynchronously on `bar` (wait on it), false when the caller only needs a
// __syncthreads() before reading the stage
__device__ __forceinline__ bool stage_load(float* stage, const float* g, long long row0, int rows, int T, int K,
                                           int c0, int cc, uint64_t* bar, float fill) {
  const int mode = io_mode(g, rows, T, K, c0, cc);
  const uint32_t bytes = (uint32_t)T * cc * 4u;
  if (mode == 1) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(stage, g + row0 * K, bytes, bar);
    }
    return true;
  }
  if (mode == 2) {
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, bytes);
    for (int r = threadIdx.x; r < T; r += blockDim.x) bulk_g2s(stage + r * cc, g + (row0 + r) * K + c0, cc * 4u, bar);
    return true;
  }
  for (int r = threadIdx.x / cc, c = threadIdx.x - r * cc; r < T;) {
    stage[r * cc + c] = (r < rows) ? __ldg(g + (row0 + r) * K + c0 + c) : fill;
    c += blockDim.x;
    while (c >= cc) { c -= cc; ++r; }
  }
  return false;
}
// every thread has executed fence_proxy_async() + __syncthreads() after the last write to `stage`
__device__ __forceinline__ void stage_store(const float* stage, float* g, long long row0, int rows, int T, int K,
                                            int c0, int cc) {
  const int mode = io_mode(g, rows, T, K, c0, cc);
  if (mode == 1) {
    if (threadIdx.x == 0) {
      bulk_s2g(g + row0 * K, stage, (uint32_t)T * cc * 4u);
      bulk_commit();
    }
  } else if (mode == 2) {
    for (int r = threadIdx.x; r < T; r += blockDim.x) bulk_s2g(g + (row0 + r) * K + c0, stage + r * cc, cc * 4u);
    bulk_commit();
  } else {
    for (int r = threadIdx.x / cc, c = threadIdx.x - r * cc; r < rows;) {
      g[(row0 + r) * K + c0 + c] = stage[r * cc + c];
      c += blockDim.x;
      while (c >= cc) { c -= cc; ++r; }
    }
  }
}

// ---- forward op execution --------------------------------------------------------------------------
// `fl` = shared address of this lane's first row in the row-major facts staging tile (base + lane * F * 4)
template <int R, bool LOG>
__device__ __forceinline__ VecT<R> leaf_value(uint32_t fl, int F, int col, bool neg) {
  VecT<R> o;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float p = lds1(fl + (32 * j * F + col) * 4);
    if (neg) p = __fsub_rn(1.0f, p);
    o.v[j] = LOG ? logf(p) : p;
  }
  return o;
}

// any op kind (the slow path: kinds without a tight loop in run_level)
template <int R, bool LOG, bool SS>
__device__ __noinline__ VecT<R> exec_op(const int4 d, uint32_t fl, int F, uint32_t vl, const Tab<SS, int32_t> edges) {
  const int kind = d.w & 0xff, nc = d.w >> 8;
  switch (kind) {
    case OP_LEAF_POS: return leaf_value<R, LOG>(fl, F, d.y, false);
    case OP_LEAF_NEG: return leaf_value<R, LOG>(fl, F, d.y, true);
    case OP_CONST: return vfill<R>(__int_as_float(d.y));
    case OP_MUL2: {
      const VecT<R> x = ldv<R>(vl, d.y), y = ldv<R>(vl, d.z);
      return LOG ? vadd_rn<R>(x, y) : vmul_rn<R>(x, y);
    }
    case OP_COPY: return ldv<R>(vl, d.y);
    case OP_COMPL1: return vcompl_rn<R>(ldv<R>(vl, d.y));
    case OP_IDENT_ONE: return vfill<R>(LOG ? 0.0f : 1.0f);
    case OP_IDENT_ZERO: return vfill<R>(LOG ? -INFINITY : 0.0f);
    case OP_PRODN: {
      VecT<R> acc = ldv<R>(vl, edges.ld4(d.y));
      for (int e = 1; e < nc; ++e) {
        const VecT<R> v = ldv<R>(vl, edges.ld4(d.y + e));
        acc = LOG ? vadd_rn<R>(acc, v) : vmul_rn<R>(acc, v);
      }
      return acc;
    }
    default: break;
  }
  // sums (ADD2 / SUMN) and complements (COMPL2 / COMPLN)
  if constexpr (!LOG) {
    VecT<R> acc;
    if (kind == OP_ADD2 || kind == OP_COMPL2) {
      acc = vadd_rn<R>(ldv<R>(vl, d.y), ldv<R>(vl, d.z));
    } else {
      acc = ldv<R>(vl, edges.ld4(d.y));
      for (int e = 1; e < nc; ++e) acc = vadd_rn<R>(acc, ldv<R>(vl, edges.ld4(d.y + e)));
    }
    return (kind == OP_COMPL2 || kind == OP_COMPLN) ? vcompl_rn<R>(acc) : acc;
  } else {  // logsumexp (COMPL is rejected for the log semiring at circuit creation)
    VecT<R> mx = vfill<R>(-INFINITY);
    for (int e = 0; e < nc; ++e) {
      const int c = (kind == OP_ADD2) ? (e == 0 ? d.y : d.z) : edges.ld4(d.y + e);
      const VecT<R> v = ldv<R>(vl, c);
#pragma unroll
      for (int j = 0; j < R; ++j) mx.v[j] = fmaxf(mx.v[j], v.v[j]);
    }
    VecT<R> s = vfill<R>(0.0f);
    for (int e = 0; e < nc; ++e) {
      const int c = (kind == OP_ADD2) ? (e == 0 ? d.y : d.z) : edges.ld4(d.y + e);
      const VecT<R> v = ldv<R>(vl, c);
#pragma unroll
      for (int j = 0; j < R; ++j) s.v[j] += expf(v.v[j] - mx.v[j]);
    }
    VecT<R> o;
#pragma unroll
    for (int j = 0; j < R; ++j) o.v[j] = (mx.v[j] == -INFINITY) ? -INFINITY : mx.v[j] + logf(s.v[j]);
    return o;
  }
}

// One level of ops.  The schedule sorts the ops of a level by kind and describes every run of equal kind as a
// SEGMENT {kind, first op, end}; the kind is decoded once per segment and the segment's ops are then swept by
// a tight loop specialised for that kind — warps take G = 4 / R ops at a time (with R rows per lane one op
// already carries R independent chains), the last group of a segment re-executes the segment's last op instead
// of branching (recomputing a node is harmless).  Binary ops: 2 G independent shared-memory loads, G R FP ops,
// G stores.  N-ary folds: child lists padded to a multiple of four with pointers to the constant-0 / constant-1
// rows and fetched 128 bits at a time, every fold sequential in CSR order (bit-exact).
template <int R, int G, bool ADD, bool SS>
__device__ __forceinline__ void seg_binary(const Tab<SS, int4> ops, int beg, int end, uint32_t vl, int warp, int nwarps) {
  for (int i = beg + G * warp; i < end; i += G * nwarps) {
    int4 d[G];
#pragma unroll
    for (int u = 0; u < G; ++u) d[u] = ops.ld16(min(i + u, end - 1));
    VecT<R> x[G], y[G];
#pragma unroll
    for (int u = 0; u < G; ++u) { x[u] = ldv<R>(vl, d[u].y); y[u] = ldv<R>(vl, d[u].z); }
#pragma unroll
    for (int u = 0; u < G; ++u) stv<R>(vl, d[u].x, ADD ? vadd_rn<R>(x[u], y[u]) : vmul_rn<R>(x[u], y[u]));
  }
}
template <int R, int G, bool ADD, bool SS>
__device__ __forceinline__ void seg_fold(const Tab<SS, int4> ops, const Tab<SS, int32_t> edges, int beg, int end,
                                         uint32_t vl, int warp, int nwarps) {
  for (int i = beg + G * warp; i < end; i += G * nwarps) {
    int4 d[G];
    int len = 0;
#pragma unroll
    for (int u = 0; u < G; ++u) {
      d[u] = ops.ld16(min(i + u, end - 1));
      len = max(len, d[u].z);    // padded length (multiple of four); shorter lists stop early
    }
    VecT<R> acc[G];
#pragma unroll
    for (int u = 0; u < G; ++u) acc[u] = vfill<R>(ADD ? 0.0f : 1.0f);  // 0 + x and 1 * x are exact
    for (int e = 0; e < len; e += 4) {
#pragma unroll
      for (int u = 0; u < G; ++u) {
        if (G > 1 && e >= d[u].z) continue;
        const int4 c = edges.ld16(d[u].y + e);
        const VecT<R> v0 = ldv<R>(vl, c.x), v1 = ldv<R>(vl, c.y), v2 = ldv<R>(vl, c.z), v3 = ldv<R>(vl, c.w);
        if (ADD) {
          acc[u] = vadd_rn<R>(vadd_rn<R>(vadd_rn<R>(vadd_rn<R>(acc[u], v0), v1), v2), v3);
        } else {
          acc[u] = vmul_rn<R>(vmul_rn<R>(vmul_rn<R>(vmul_rn<R>(acc[u], v0), v1), v2), v3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < G; ++u) stv<R>(vl, d[u].x, acc[u]);
  }
}

template <int R, bool LOG, bool SS>
__device__ __forceinline__ void run_level(const Tab<SS, int4> segs, int sbeg, int send, const Tab<SS, int4> ops,
                                          const Tab<SS, int32_t> edges, uint32_t fl, int F, uint32_t vl, int warp,
                                          int nwarps) {
  constexpr int G = 4 / R;
  for (int s = sbeg; s < send; ++s) {
    const int4 sg = segs.ld16(s);
    const int beg = sg.y, end = sg.z;
    switch (sg.x) {
      case OP_LEAF_POS:
      case OP_LEAF_NEG:
        for (int i = beg + warp; i < end; i += nwarps) {
          const int4 d = ops.ld16(i);
          stv<R>(vl, d.x, leaf_value<R, LOG>(fl, F, d.y, sg.x == OP_LEAF_NEG));
        }
        break;
      case OP_MUL2:
        if (LOG) seg_binary<R, G, true, SS>(ops, beg, end, vl, warp, nwarps);
        else seg_binary<R, G, false, SS>(ops, beg, end, vl, warp, nwarps);
        break;
      case OP_PRODN:
        if (LOG) seg_fold<R, G, true, SS>(ops, edges, beg, end, vl, warp, nwarps);
        else seg_fold<R, G, false, SS>(ops, edges, beg, end, vl, warp, nwarps);
        break;
      default:
        if (!LOG && sg.x == OP_ADD2) {
          seg_binary<R, G, true, SS>(ops, beg, end, vl, warp, nwarps);
        } else if (!LOG && sg.x == OP_SUMN) {
          seg_fold<R, G, true, SS>(ops, edges, beg, end, vl, warp, nwarps);
        } else {
          for (int i = beg + warp; i < end; i += nwarps) {
            const int4 d = ops.ld16(i);
            stv<R>(vl, d.x, exec_op<R, LOG, SS>(d, fl, F, vl, edges));
          }
        }
        break;
    }
  }
}

// the two constant rows behind the node rows (written once per CTA; nothing ever overwrites them)
__device__ __forceinline__ void init_const_rows(float* val, int n_rows, int T) {
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    val[n_rows * T + i] = 0.0f;
    val[(n_rows + 1) * T + i] = 1.0f;
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int R, bool LOG, bool SS>
__global__ void __launch_bounds__(512) generic_forward_kernel(const GenArgs A) {
  constexpr int T = 32 * R;
  const GenProgram& P = A.G;
  const GenCall& C = A.C;
  extern __shared__ __align__(16) unsigned char gsm_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(gsm_raw);        // facts tile
  float* val = reinterpret_cast<float*>(gsm_raw + 16);          // [n_nodes + 2][T]
  float* fs = val + (P.n_nodes + 2) * T;                        // [T][F] row-major staging
  float* out = fs + T * P.n_facts;                              // [T][<= cf] row-major staging
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // opaque to the compiler from here on: otherwise it re-derives the three values from S2R / LDC at every use
  // (about a hundred instructions per warp and tile)
  asm volatile("" : "+r"(warp), "+r"(lane), "+r"(nwarps));
  const uint32_t vl = smem_u32(val) + lane * (4 * R);
  const uint32_t fl = smem_u32(fs) + lane * (4 * P.n_facts);
  const long long n_tiles = (C.B + T - 1) / T;
  const bool norm = (C.flags & VAEL_FLAG_NORMALIZE) && P.norm_node >= 0;
  const bool evid = (C.flags & VAEL_FLAG_EVIDENCE) != 0;
  const int zoff = P.norm_node * (T * 4);
  const int W = P.n_worlds, F = P.n_facts, Q = P.n_queries;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  init_const_rows(val, P.n_nodes, T);
  // a small schedule is copied into shared memory once per CTA: every op / child-list / segment fetch is then a
  // shared-memory broadcast instead of an L1 round trip in the middle of a dependent chain
  int4* s_ops = reinterpret_cast<int4*>(out + T * P.cq);
  int4* s_segs = s_ops + P.n_ops;
  int32_t* s_edges = reinterpret_cast<int32_t*>(s_segs + P.n_segs);
  int32_t* s_woff = s_edges + P.n_edges;
  if (SS) {
    for (int i = threadIdx.x; i < P.n_ops; i += blockDim.x) s_ops[i] = __ldg(P.ops + i);
    for (int i = threadIdx.x; i < P.n_segs; i += blockDim.x) s_segs[i] = __ldg(P.segs + i);
    for (int i = threadIdx.x; i < P.n_edges; i += blockDim.x) s_edges[i] = __ldg(P.edges + i);
    for (int i = threadIdx.x; i < ((W + 3) & ~3); i += blockDim.x) s_woff[i] = __ldg(P.world_off + i);
  }
  __syncthreads();
  const Tab<SS, int4> ops(P.ops, s_ops), segs(P.segs, s_segs);
  const Tab<SS, int32_t> edges(P.edges, s_edges), world_off(P.world_off, s_woff);

  // forward schedule: every node when Z is asked for, else without the cone only the normaliser needs
  const int sch = norm ? 0 : 1;
  const int32_t* lptr = P.seg_level_ptr[sch];
  const int n_lev = P.n_levels_f[sch], leaf_lev = P.leaf_levels_f[sch];
  const bool need_q = C.query != nullptr || evid;

  long long tile = blockIdx.x;
  bool f_async = false;
  uint32_t ph = 0;
  if (tile < n_tiles)
    f_async = stage_load(fs, C.facts, tile * T, (int)min((long long)T, C.B - tile * T), T, F, 0, F, bar, 0.5f);

  for (; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * T;
    const int rows = (int)min((long long)T, C.B - row0);
    // per-row query index (lane <-> rows lane + 32 j), requested before the levels so that its HBM latency hides
    int32_t qrow[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int r = lane + 32 * j;
      qrow[j] = (need_q && r < rows) ? (C.qidx ? __ldg(C.qidx + row0 + r) : C.qscalar) : -1;
    }
    if (f_async) { mbar_wait(bar, ph); ph ^= 1; } else __syncthreads();
    for (int L = 0; L < n_lev; ++L) {
      run_level<R, LOG, SS>(segs, lptr[L], lptr[L + 1], ops, edges, fl, F, vl, warp, nwarps);
      if (L == n_lev - 1) bulk_wait_read<0>();   // the previous tile's stores have finished reading the staging tile
      __syncthreads();
      if (L == leaf_lev - 1) {  // the facts tile is consumed: fetch the next one behind the remaining levels
        const long long nt = tile + gridDim.x;
        if (nt < n_tiles)
          f_async = stage_load(fs, C.facts, nt * T, (int)min((long long)T, C.B - nt * T), T, F, 0, F, bar, 0.5f);
      }
    }

    VecT<R> z = vfill<R>(1.0f);
    if (norm) z = ldv<R>(vl, zoff);
    float qv[R];   // value of the row's query node (0 when the row has no valid query)
    if (need_q) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (qrow[j] < 0 || qrow[j] >= Q) qrow[j] = -1;
        qv[j] = qrow[j] >= 0 ? lds1(vl + __ldg(P.query_off + qrow[j]) + 4 * j) : 0.0f;
      }
    }
    if (C.query != nullptr && warp == nwarps - 1) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int r = lane + 32 * j;
        if (r < rows) C.query[row0 + r] = norm ? qv[j] / z.v[j] : qv[j];
      }
    }

    // ---- worlds: node-major -> row-major staging (chunks of cf columns) -> bulk store ----
    bool stored = false;
    if (C.worlds != nullptr) {
      for (int c0 = 0; c0 < W; c0 += P.cf) {
        const int cc = min(P.cf, W - c0);
        if (stored) {   // the staging tile is still being read by the previous chunk's store
          bulk_wait_read<0>();
          __syncthreads();
        }
        if ((cc & 3) == 0) {
          for (int g = warp; g < (cc >> 2); g += nwarps) {
            const int w = c0 + 4 * g;
            const int4 off = world_off.ld16(w);
            const VecT<R> a0 = ldv<R>(vl, off.x), a1 = ldv<R>(vl, off.y), a2 = ldv<R>(vl, off.z), a3 = ldv<R>(vl, off.w);
#pragma unroll
            for (int j = 0; j < R; ++j) {
              float4 o = make_float4(a0.v[j], a1.v[j], a2.v[j], a3.v[j]);
              if (evid) {
                const uint32_t mword = qrow[j] >= 0 ? __ldg(C.qmask + (long long)qrow[j] * C.mask_words + (w >> 5)) : 0u;
                const uint32_t m = mword >> (w & 31);
                o.x = (m & 1u) ? o.x / qv[j] : 0.0f;
                o.y = (m & 2u) ? o.y / qv[j] : 0.0f;
                o.z = (m & 4u) ? o.z / qv[j] : 0.0f;
                o.w = (m & 8u) ? o.w / qv[j] : 0.0f;
              } else if (norm) {
                o.x = o.x / z.v[j]; o.y = o.y / z.v[j]; o.z = o.z / z.v[j]; o.w = o.w / z.v[j];
              }
              *reinterpret_cast<float4*>(out + (lane + 32 * j) * cc + 4 * g) = o;
            }
          }
        } else {
          for (int c = warp; c < cc; c += nwarps) {
            const int w = c0 + c;
            const VecT<R> a = ldv<R>(vl, world_off.ld4(w));
#pragma unroll
            for (int j = 0; j < R; ++j) {
              float o = a.v[j];
              if (evid) {
                const uint32_t mword = qrow[j] >= 0 ? __ldg(C.qmask + (long long)qrow[j] * C.mask_words + (w >> 5)) : 0u;
                o = ((mword >> (w & 31)) & 1u) ? o / qv[j] : 0.0f;
              } else if (norm) {
                o = o / z.v[j];
              }
              out[(lane + 32 * j) * cc + c] = o;
            }
          }
        }
        fence_proxy_async();
        __syncthreads();
        stage_store(out, C.worlds, row0, rows, T, W, c0, cc);
        stored = true;
      }
    }
    // ---- every query node of every row ----
    if (C.queries_all != nullptr) {
      const int qchunk = Q <= P.cq ? Q : (P.cq & ~3);
      for (int c0 = 0; c0 < Q; c0 += qchunk) {
        const int cc = min(qchunk, Q - c0);
        if (stored) {
          bulk_wait_read<0>();
          __syncthreads();
        }
        for (int c = warp; c < cc; c += nwarps) {
          const VecT<R> a = ldv<R>(vl, __ldg(P.query_off + c0 + c));
#pragma unroll
          for (int j = 0; j < R; ++j) out[(lane + 32 * j) * cc + c] = norm ? a.v[j] / z.v[j] : a.v[j];
        }
        fence_proxy_async();
        __syncthreads();
        stage_store(out, C.queries_all, row0, rows, T, Q, c0, cc);
        stored = true;
      }
    }
    // every read of the value tile lies before the barrier in front of the last store; without a store the
    // tile still has to be released for the next tile's leaves
    if (!stored) __syncthreads();
  }
  bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------
// backward: forward values of the nodes the adjoint needs, then the reverse sweep
// ---------------------------------------------------------------------------
template <int R, bool LOG, bool SS>
__device__ __forceinline__ VecT<R> apply_rec(VecT<R> a, const Tab<SS, int2> prec, int i, uint32_t vl, uint32_t al,
                                             const Tab<SS, int32_t> bedges) {
  const int2 rec = prec.ld8(i);
  const int kind = rec.x & 7;
  const VecT<R> ap = ldv<R>(al, rec.x >> 1 & ~3);
  switch (kind) {
    case PR_ADD: {
#pragma unroll
      for (int j = 0; j < R; ++j) a.v[j] += ap.v[j];
      return a;
    }
    case PR_SUB: {
#pragma unroll
      for (int j = 0; j < R; ++j) a.v[j] -= ap.v[j];
      return a;
    }
    case PR_MUL_SIB: return vfma<R>(ap, ldv<R>(vl, rec.y), a);
    case PR_MUL_SIBS: {  // product of the other children of an n-ary product
      const int2 ext = prec.ld8(i + 1);
      const int nc = ext.y & 0xffff, pos = ext.y >> 16;
      VecT<R> sib = vfill<R>(1.0f);
      for (int e = 0; e < nc; ++e)
        if (e != pos) {
          const VecT<R> v = ldv<R>(vl, bedges.ld4(ext.x + e));
#pragma unroll
          for (int j = 0; j < R; ++j) sib.v[j] *= v.v[j];
        }
      return vfma<R>(ap, sib, a);
    }
    default: {  // PR_LSE: softmax weight of the child inside a logsumexp parent
      const int2 ext = prec.ld8(i + 1);
      const VecT<R> vc = ldv<R>(vl, rec.y), vp = ldv<R>(vl, ext.x);
#pragma unroll
      for (int j = 0; j < R; ++j) a.v[j] = fmaf(ap.v[j], (vp.v[j] == -INFINITY) ? 0.0f : expf(vc.v[j] - vp.v[j]), a.v[j]);
      return a;
    }
  }
}

template <int R, bool LOG, bool SS>
__global__ void __launch_bounds__(512) generic_backward_kernel(const GenArgs A) {
  constexpr int T = 32 * R;
  const GenProgram& P = A.G;
  const GenCall& C = A.C;
  extern __shared__ __align__(16) unsigned char gsm_raw[];
  uint64_t* bar_f = reinterpret_cast<uint64_t*>(gsm_raw);       // facts tile
  uint64_t* bar_g = bar_f + 1;                                  // grad_worlds chunk
  float* val = reinterpret_cast<float*>(gsm_raw + 16);          // [n_bvals + 2][T]
  float* adj = val + (P.n_bvals + 2) * T;                       // [n_nodes + 1][T]
  float* fs = adj + (P.n_nodes + 1) * T;                        // [T][F] facts staging
  float* gfs = fs + T * P.n_facts;                              // [T][F] grad_facts staging
  float* gws = gfs + T * P.n_facts;                             // [T][<= cb] grad_worlds staging
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // opaque to the compiler from here on: otherwise it re-derives the three values from S2R / LDC at every use
  // (about a hundred instructions per warp and tile)
  asm volatile("" : "+r"(warp), "+r"(lane), "+r"(nwarps));
  const uint32_t vl = smem_u32(val) + lane * (4 * R);
  const uint32_t al = smem_u32(adj) + lane * (4 * R);
  const uint32_t fl = smem_u32(fs) + lane * (4 * P.n_facts);
  const long long n_tiles = (C.B + T - 1) / T;
  const bool norm = (C.flags & VAEL_FLAG_NORMALIZE) && P.norm_slot_off >= 0;
  const int W = P.n_worlds, F = P.n_facts, Q = P.n_queries;
  const bool have_gw = C.grad_worlds != nullptr;
  if (threadIdx.x == 0) {
    mbar_init(bar_f, 1);
    mbar_init(bar_g, 1);
    fence_mbar_init();
  }
  init_const_rows(val, P.n_bvals, T);
  // a small schedule is staged in shared memory once per CTA (see the forward kernel)
  int4* s_bops = reinterpret_cast<int4*>(gws + T * min(P.cb, max(W, 1)));
  int4* s_bsegs = s_bops + P.n_bops;
  int4* s_crec = s_bsegs + P.n_bsegs;
  int4* s_csegs = s_crec + P.n_crec;
  int32_t* s_bedges = reinterpret_cast<int32_t*>(s_csegs + P.n_csegs);   // 16-byte aligned: read 128 bits at a time
  int32_t* s_wseed = s_bedges + P.n_bedges;                               // n_bedges is a multiple of four
  int2* s_prec = reinterpret_cast<int2*>(s_wseed + ((W + 3) & ~3));
  if (SS) {
    for (int i = threadIdx.x; i < P.n_bops; i += blockDim.x) s_bops[i] = __ldg(P.bops + i);
    for (int i = threadIdx.x; i < P.n_bsegs; i += blockDim.x) s_bsegs[i] = __ldg(P.bsegs + i);
    for (int i = threadIdx.x; i < P.n_crec; i += blockDim.x) s_crec[i] = __ldg(P.crec + i);
    for (int i = threadIdx.x; i < P.n_csegs; i += blockDim.x) s_csegs[i] = __ldg(P.csegs + i);
    for (int i = threadIdx.x; i < P.n_prec; i += blockDim.x) s_prec[i] = __ldg(P.prec + i);
    for (int i = threadIdx.x; i < P.n_bedges; i += blockDim.x) s_bedges[i] = __ldg(P.bedges + i);
    for (int i = threadIdx.x; i < ((W + 3) & ~3); i += blockDim.x) s_wseed[i] = __ldg(P.wseed_off + i);
  }
  __syncthreads();
  const Tab<SS, int4> bops(P.bops, s_bops), bsegs(P.bsegs, s_bsegs), crec(P.crec, s_crec), csegs(P.csegs, s_csegs);
  const Tab<SS, int2> prec(P.prec, s_prec);
  const Tab<SS, int32_t> bedges(P.bedges, s_bedges), wseed_off(P.wseed_off, s_wseed);

  // recomputed values: with the normaliser's cone only when Z is asked for
  const int32_t* blptr = P.bseg_level_ptr[norm ? 0 : 1];
  const int leaf_lev_b = P.leaf_levels_b[norm ? 0 : 1];

  long long tile = blockIdx.x;
  bool f_async = false, g_async = false;
  uint32_t ph_f = 0, ph_g = 0;
  if (tile < n_tiles) {
    const int rows = (int)min((long long)T, C.B - tile * T);
    f_async = stage_load(fs, C.facts, tile * T, rows, T, F, 0, F, bar_f, 0.5f);
    if (have_gw) g_async = stage_load(gws, C.grad_worlds, tile * T, rows, T, W, 0, min(P.cb, W), bar_g, 0.0f);
  }

  for (; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * T;
    const int rows = (int)min((long long)T, C.B - row0);
    const long long nt = tile + gridDim.x;
    const int nrows = (int)min((long long)T, C.B - nt * T);
    // per-row query seed (lane <-> rows lane + 32 j), requested first so that its HBM latency hides
    int32_t qrow[R];
    float gq[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int r = lane + 32 * j;
      const bool on = C.grad_query != nullptr && r < rows;
      qrow[j] = on ? (C.qidx ? __ldg(C.qidx + row0 + r) : C.qscalar) : -1;
      gq[j] = on ? __ldg(C.grad_query + row0 + r) : 0.0f;
    }
    if (f_async) { mbar_wait(bar_f, ph_f); ph_f ^= 1; } else __syncthreads();

    for (int L = 0; L < P.n_levels; ++L) {
      if (blptr[L + 1] > blptr[L]) {
        run_level<R, LOG, SS>(bsegs, blptr[L], blptr[L + 1], bops, bedges, fl, F, vl, warp, nwarps);
        __syncthreads();
      }
      if (!LOG && L == leaf_lev_b - 1 && nt < n_tiles)
        f_async = stage_load(fs, C.facts, nt * T, nrows, T, F, 0, F, bar_f, 0.5f);
    }
    if (!LOG && leaf_lev_b == 0 && nt < n_tiles) {
      __syncthreads();   // every thread is past its wait on the facts barrier
      f_async = stage_load(fs, C.facts, nt * T, nrows, T, F, 0, F, bar_f, 0.5f);
    }

    // ---- seeds.  Query nodes that are not world nodes: assigned (no zero-fill of the adjoint tile) ----
    VecT<R> zv = vfill<R>(1.0f);
    if (norm) zv = ldv<R>(vl, P.norm_slot_off);
    int32_t qoff[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      qoff[j] = (qrow[j] >= 0 && qrow[j] < Q) ? __ldg(P.qseed_off + qrow[j]) : -1;
      if (norm) gq[j] = gq[j] / zv.v[j];
    }
    for (int i = warp; i < P.n_qonly; i += nwarps) {
      const int32_t off = __ldg(P.qonly_off + i);
      VecT<R> a;
#pragma unroll
      for (int j = 0; j < R; ++j) a.v[j] = (qoff[j] == off) ? gq[j] : 0.0f;
      stv<R>(al, off, a);
    }
    // ---- world adjoints from the upstream gradient: row-major staging -> node-major (assignment) ----
    if (!have_gw) {
      for (int w = warp; w < W; w += nwarps) stv<R>(al, wseed_off.ld4(w), vfill<R>(0.0f));
      __syncthreads();
    } else {
      for (int c0 = 0; c0 < W; c0 += P.cb) {
        const int cc = min(P.cb, W - c0);
        if (g_async) { mbar_wait(bar_g, ph_g); ph_g ^= 1; } else __syncthreads();
        if ((cc & 3) == 0) {
          for (int g = warp; g < (cc >> 2); g += nwarps) {
            const int w = c0 + 4 * g;
            const int4 off = wseed_off.ld16(w);
            VecT<R> a0, a1, a2, a3;
#pragma unroll
            for (int j = 0; j < R; ++j) {
              float4 gv = *reinterpret_cast<const float4*>(gws + (lane + 32 * j) * cc + 4 * g);
              if (norm) { gv.x = gv.x / zv.v[j]; gv.y = gv.y / zv.v[j]; gv.z = gv.z / zv.v[j]; gv.w = gv.w / zv.v[j]; }
              a0.v[j] = gv.x; a1.v[j] = gv.y; a2.v[j] = gv.z; a3.v[j] = gv.w;
            }
            stv<R>(al, off.x, a0); stv<R>(al, off.y, a1); stv<R>(al, off.z, a2); stv<R>(al, off.w, a3);
          }
        } else {
          for (int c = warp; c < cc; c += nwarps) {
            VecT<R> a;
#pragma unroll
            for (int j = 0; j < R; ++j) {
              const float gv = gws[(lane + 32 * j) * cc + c];
              a.v[j] = norm ? gv / zv.v[j] : gv;
            }
            stv<R>(al, wseed_off.ld4(c0 + c), a);
          }
        }
        __syncthreads();   // chunk consumed: the staging tile is free, the seeds are visible
        // next chunk of this tile, or the first chunk of the next tile (behind the reverse sweep)
        if (c0 + P.cb < W) {
          g_async = stage_load(gws, C.grad_worlds, row0, rows, T, W, c0 + P.cb, min(P.cb, W - c0 - P.cb), bar_g, 0.0f);
        } else if (nt < n_tiles) {
          g_async = stage_load(gws, C.grad_worlds, nt * T, nrows, T, W, 0, min(P.cb, W), bar_g, 0.0f);
        }
      }
    }
    if (P.q_world_any && C.grad_query != nullptr) {   // query nodes that are world nodes: added to the world seed
      if (warp == 0) {
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (qoff[j] >= 0 && (qoff[j] & 1)) {
            const uint32_t a = al + (qoff[j] & ~1) + 4 * j;
            sts1(a, lds1(a) + gq[j]);
          }
      }
      __syncthreads();
    }

    // reverse sweep, gather form: adj[c] = seed[c] + sum over live parents p of adj[p] * d val[p] / d val[c].
    // The children of a level are sorted by shape (record count and kinds) and every run of one shape is a
    // segment {flags, first child, end, n records}: the shape is decoded once per segment, the children are then
    // swept G = 4 / R at a time (G R accumulation chains in flight per warp; the last group of a segment
    // recomputes the segment's last child instead of branching).
    constexpr int G = 4 / R;
    for (int L = P.n_levels - 2; L >= 0; --L) {
      for (int sgi = P.cseg_level_ptr[L]; sgi < P.cseg_level_ptr[L + 1]; ++sgi) {
        const int4 sg = csegs.ld16(sgi);
        const int beg = sg.y, end = sg.z, n = sg.w, uk = sg.x >> CR_UNIFORM_SHIFT;
        const bool seeded = sg.x & CR_SEEDED;
        if (!(sg.x & CR_WIDE) && (uk == 1 + PR_MUL_SIB || uk == 1 + PR_ADD)) {
          for (int i = beg + G * warp; i < end; i += G * nwarps) {
            int4 cr[G];
            VecT<R> a[G];
#pragma unroll
            for (int u = 0; u < G; ++u) {
              cr[u] = crec.ld16(min(i + u, end - 1));
              a[u] = seeded ? ldv<R>(al, cr[u].x) : vfill<R>(0.0f);
            }
            if (uk == 1 + PR_MUL_SIB) {
              for (int r = 0; r < n; ++r) {
                int2 rec[G];
#pragma unroll
                for (int u = 0; u < G; ++u) rec[u] = prec.ld8(cr[u].y + r);
#pragma unroll
                for (int u = 0; u < G; ++u) a[u] = vfma<R>(ldv<R>(al, rec[u].x >> 1 & ~3), ldv<R>(vl, rec[u].y), a[u]);
              }
            } else {
              for (int r = 0; r < n; ++r) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                  const VecT<R> ap = ldv<R>(al, prec.ld8(cr[u].y + r).x >> 1 & ~3);
#pragma unroll
                  for (int j = 0; j < R; ++j) a[u].v[j] += ap.v[j];
                }
              }
            }
#pragma unroll
            for (int u = 0; u < G; ++u) stv<R>(al, cr[u].x, a[u]);
          }
        } else {
          for (int i = beg + warp; i < end; i += nwarps) {
            const int4 cr = crec.ld16(i);
            VecT<R> a = seeded ? ldv<R>(al, cr.x) : vfill<R>(0.0f);
            for (int r = 0, k = cr.y; r < cr.z; ++r) {
              const int kind = prec.ld8(k).x & 7;
              a = apply_rec<R, LOG, SS>(a, prec, k, vl, al, bedges);
              k += (kind == PR_MUL_SIBS || kind == PR_LSE) ? 2 : 1;
            }
            stv<R>(al, cr.x, a);
          }
        }
      }
      if (L == 0) bulk_wait_read<0>();   // the previous tile's store has finished reading the grad_facts staging
      __syncthreads();
    }
    if (P.n_levels < 2) {
      bulk_wait_read<0>();
      __syncthreads();
    }

    // ---- live leaves -> grad_facts staging (lane <-> row, warps take fact columns) -> bulk store ----
    for (int f = warp; f < F; f += nwarps) {
      const int kb = __ldg(P.leaf_ptr + f), ke = __ldg(P.leaf_ptr + f + 1);
      VecT<R> g = vfill<R>(0.0f);
      for (int k = kb; k < ke; ++k) {
        const int2 lr = __ldg(P.leaf_rec + k);
        const VecT<R> a = ldv<R>(al, lr.x);
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (!LOG) {
            g.v[j] += lr.y < 0 ? -a.v[j] : a.v[j];
          } else {
            const float pf = lds1(fl + (32 * j * F + f) * 4);
            g.v[j] += lr.y < 0 ? -a.v[j] / (1.0f - pf) : a.v[j] / pf;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < R; ++j) gfs[(lane + 32 * j) * F + f] = g.v[j];
    }
    fence_proxy_async();
    __syncthreads();
    stage_store(gfs, C.grad_facts, row0, rows, T, F, 0, F);
    if (LOG && nt < n_tiles) f_async = stage_load(fs, C.facts, nt * T, nrows, T, F, 0, F, bar_f, 0.5f);
  }
  bulk_wait_all<0>();
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
  cudaFree(g->segs); cudaFree(g->bsegs); cudaFree(g->csegs);
  cudaFree(g->ops); cudaFree(g->edges); cudaFree(g->world_off); cudaFree(g->query_off); cudaFree(g->bops);
  cudaFree(g->bedges); cudaFree(g->crec); cudaFree(g->prec); cudaFree(g->wseed_off); cudaFree(g->qseed_off);
  cudaFree(g->qonly_off); cudaFree(g->leaf_rec); cudaFree(g->leaf_ptr);
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
    h.op.y = d->node_arg[n];
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

// Sort the ops of one level by kind (then fan-in), describe every run of one kind as a segment, lay out the child
// lists: every list starts at a multiple of four and is padded to a multiple of four with pointers to the constant
// row that is the identity of the fold (zero_off / one_off).
void emit_level(std::vector<HostOp>& lvl, bool is_log, int zero_off, int one_off, std::vector<int4>& ops,
                std::vector<int32_t>& edges, std::vector<int4>& segs) {
  if (lvl.empty()) return;
  std::stable_sort(lvl.begin(), lvl.end(), [](const HostOp& a, const HostOp& b) {
    return (a.op.w & 0xff) != (b.op.w & 0xff) ? (a.op.w & 0xff) < (b.op.w & 0xff) : (a.op.w >> 8) < (b.op.w >> 8);
  });
  for (HostOp& h : lvl) {
    const int code = h.op.w & 0xff;
    if (!h.kids.empty()) {
      const int pad = (code == OP_PRODN && !is_log) ? one_off : zero_off;
      const size_t len = (h.kids.size() + 3) / 4 * 4;
      h.op.y = (int)edges.size();
      h.op.z = (int)len;
      for (size_t e = 0; e < len; ++e) edges.push_back(e < h.kids.size() ? h.kids[e] : pad);
    }
    if (segs.empty() || segs.back().x != code || segs.back().w != 0) segs.push_back(make_int4(code, (int)ops.size(), 0, 0));
    ops.push_back(h.op);
    segs.back().z = (int)ops.size();
  }
  segs.back().w = 1;   // closed: the next level starts a new segment even when it continues the same kind
}
}  // namespace


// Rows per lane (R) and staged chunk width for one kernel: the largest R whose tile — `node_rows` value /
// adjoint rows of 128 R bytes, `n_small` [T, F] staging tiles and one [T, chunk] staging tile — fits one CTA's
// shared memory with the whole [T, W] tensor staged at once; circuits too wide for that run R = 1 with
// column chunks.  VAEL_GEN_R forces R (experiments).
constexpr size_t kSmemLimit = (size_t)227 * 1024 - 64;
static size_t tile_smem(size_t node_rows, int F, int n_small, int cols, int R) {
  return 16 + (size_t)128 * R * (node_rows + (size_t)n_small * F + cols);
}
static bool choose_tile(size_t node_rows, int F, int n_small, int W, int Q, int* R_out, int* chunk, int* cap,
                        size_t* smem) {
  // measured on B200 (2-digit circuit): R = 2 with ~24 warps per SM is as fast as R = 4 forward and faster
  // backward (more CTAs per SM cover each other's level barriers); VAEL_GEN_R = 4 / 1 forces another width
  static const int forced = [] { const char* v = getenv("VAEL_GEN_R"); return (v && *v) ? atoi(v) : 2; }();
  // VAEL_GEN_CTAS = n: size the tile for n resident CTAs per SM, chunking the staged columns if need be
  static const int ctas = [] { const char* v = getenv("VAEL_GEN_CTAS"); return (v && *v) ? std::max(1, atoi(v)) : 1; }();
  const size_t limit = ctas > 1 ? ((size_t)227 * 1024) / ctas - 1024 - 64 : kSmemLimit;
  const int wcols = std::max(W, 1), qcols = std::min(std::max(Q, 1), 32);
  for (int R = 4; R >= 1; R >>= 1) {
    if (forced && R != forced && R != 1) continue;
    const size_t s = tile_smem(node_rows, F, n_small, std::max(wcols, qcols), R);
    if (s <= limit) { *R_out = R; *chunk = wcols; *cap = std::max(wcols, qcols); *smem = s; return true; }
    if (ctas > 1 && (W & 3) == 0) {   // chunked staging at this R: at least 32 columns per chunk
      const size_t fixed = tile_smem(node_rows, F, n_small, 0, R);
      if (fixed + (size_t)128 * R * 32 <= limit) {
        const int cols = (int)std::min<size_t>((limit - fixed) / (128 * R), (size_t)wcols) & ~3;
        *R_out = R; *chunk = *cap = cols; *smem = tile_smem(node_rows, F, n_small, cols, R);
        return true;
      }
    }
  }
  const size_t fixed = tile_smem(node_rows, F, n_small, 0, 1);
  if (fixed + 128 * 32 > kSmemLimit) { *R_out = 1; *chunk = *cap = 32; *smem = fixed + 128 * 32; return false; }
  const int cols = (int)std::min<size_t>((kSmemLimit - fixed) / 128, (size_t)wcols) & ~3;
  *R_out = 1; *chunk = *cap = cols; *smem = tile_smem(node_rows, F, n_small, cols, 1);
  return true;
}

GenProgram* gen_program_create(const vael_circuit_desc_t* d, cudaError_t* err) {
  GenProgram* g = new GenProgram();
  memset(g, 0, sizeof(*g));
  g->n_nodes = d->n_nodes; g->n_facts = d->n_facts; g->n_worlds = d->n_worlds; g->n_queries = d->n_queries;
  g->n_levels = d->n_levels; g->semiring = d->semiring; g->norm_node = d->norm_node < 0 ? -1 : d->norm_node;
  const bool is_log = d->semiring == VAEL_SEMIRING_LOG;
  const int N = d->n_nodes;

  // ---- forward schedule ----
  choose_tile((size_t)N + 2, d->n_facts, 1, d->n_worlds, d->n_queries, &g->rf, &g->cf, &g->cq,
              &g->fwd_smem);
  const int PBF = 128 * g->rf;
  // two schedules: [0] every node; [1] only the cones of the world and query nodes (what a call without
  // VAEL_FLAG_NORMALIZE needs: the normaliser's private cone is skipped)
  std::vector<char> out_cone(N, 0);
  for (int w = 0; w < d->n_worlds; ++w) out_cone[d->world_node[w]] = 1;
  for (int q = 0; q < d->n_queries; ++q) out_cone[d->query_node[q]] = 1;
  for (int n = N - 1; n >= 0; --n)
    if (out_cone[n])
      for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) out_cone[d->child_idx[e]] = 1;
  std::vector<int4> ops, segs;
  std::vector<int32_t> edges;
  for (int sch = 0; sch < 2; ++sch) {
    g->leaf_levels_f[sch] = 1;
    g->n_levels_f[sch] = 1;
    for (int L = 0; L < d->n_levels; ++L) {
      g->seg_level_ptr[sch][L] = (int)segs.size();
      std::vector<HostOp> lvl;
      for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n)
        if (sch == 0 || out_cone[n]) lvl.push_back(make_op(d, n, [&](int x) { return x * PBF; }));
      if (!lvl.empty()) g->n_levels_f[sch] = L + 1;
      for (const HostOp& h : lvl)
        if ((h.op.w & 0xff) <= OP_LEAF_NEG) g->leaf_levels_f[sch] = L + 1;
      emit_level(lvl, is_log, N * PBF, (N + 1) * PBF, ops, edges, segs);
    }
    g->seg_level_ptr[sch][d->n_levels] = (int)segs.size();
  }
  g->n_ops = (int)ops.size(); g->n_edges = (int)((edges.size() + 3) / 4 * 4); g->n_segs = (int)segs.size();
  edges.resize(g->n_edges, 0);
  {
    const size_t sched = (size_t)16 * (g->n_ops + g->n_segs) + (size_t)4 * (g->n_edges + (d->n_worlds + 3) / 4 * 4);
    static const int sched_limit = [] { const char* v = getenv("VAEL_GEN_SCHED_SMEM"); return (v && *v) ? atoi(v) : 12288; }();
    // staged behind the output staging tile (which is cq columns wide)
    if (sched <= (size_t)sched_limit && g->fwd_smem + sched <= kSmemLimit && g->cf == std::max(d->n_worlds, 1)) {
      g->sched_in_smem = 1;
      g->fwd_smem += sched;
    }
  }
  std::vector<int32_t> world_off((d->n_worlds + 3) / 4 * 4, 0), query_off(d->n_queries);
  for (int w = 0; w < d->n_worlds; ++w) world_off[w] = d->world_node[w] * PBF;
  for (int q = 0; q < d->n_queries; ++q) query_off[q] = d->query_node[q] * PBF;

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
  // two sets: what a call without VAEL_FLAG_NORMALIZE needs, and that plus the normaliser's cone.  Both use the
  // slot numbering of the larger set (the parent records refer to it).
  auto close = [&](std::vector<char>& set) {
    for (int n = N - 1; n >= 0; --n)  // closure: a kept node needs its children
      if (set[n])
        for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) set[d->child_idx[e]] = 1;
  };
  std::vector<char> need_plain = need;
  close(need_plain);
  if (g->norm_node >= 0) need[g->norm_node] = 1;
  close(need);
  std::vector<int32_t> bslot(N, -1);
  for (int n = 0; n < N; ++n)
    if (need[n]) bslot[n] = g->n_bvals++;
  int unused_cap = 0;
  choose_tile((size_t)g->n_bvals + 2 + N + 1, d->n_facts, 2, d->n_worlds, 0, &g->rb, &g->cb, &unused_cap,
              &g->bwd_smem);
  const int PBB = 128 * g->rb;
  std::vector<int4> bops, bsegs;
  std::vector<int32_t> bedges;
  for (int sch = 0; sch < 2; ++sch) {
    const std::vector<char>& set = sch == 0 ? need : need_plain;
    g->leaf_levels_b[sch] = 0;
    for (int L = 0; L < d->n_levels; ++L) {
      g->bseg_level_ptr[sch][L] = (int)bsegs.size();
      std::vector<HostOp> lvl;
      for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n)
        if (set[n]) lvl.push_back(make_op(d, n, [&](int x) { return bslot[x] * PBB; }));
      for (const HostOp& h : lvl)
        if ((h.op.w & 0xff) <= OP_LEAF_NEG) g->leaf_levels_b[sch] = L + 1;
      emit_level(lvl, is_log, g->n_bvals * PBB, (g->n_bvals + 1) * PBB, bops, bedges, bsegs);
    }
    g->bseg_level_ptr[sch][d->n_levels] = (int)bsegs.size();
  }
  g->norm_slot_off = g->norm_node >= 0 ? bslot[g->norm_node] * PBB : -1;

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

  std::vector<int4> crec, csegs;
  std::vector<int2> prec;
  for (int L = 0; L < d->n_levels; ++L) {
    g->cseg_level_ptr[L] = (int)csegs.size();
    const size_t lvl_beg = crec.size();
    for (int c = d->level_ptr[L]; c < d->level_ptr[L + 1]; ++c) {
      if (!live[c] || parents[c].empty()) continue;
      int4 cr = make_int4(c * PBB, (int)prec.size(), (int)parents[c].size(), seed[c] ? CR_SEEDED : 0);
      int uniform = -1;
      bool first = true;
      for (auto& pr : parents[c]) {
        const int p = pr.first, pos = pr.second;
        const int k = d->node_kind[p], cb = d->child_ptr[p], nc = d->child_ptr[p + 1] - cb;
        const int padj = (p * PBB) << 1;
        int kind;
        if (!is_log) {
          if (k == VAEL_NODE_SUM) { kind = PR_ADD; prec.push_back(make_int2(padj | kind, 0)); }
          else if (k == VAEL_NODE_COMPL) { kind = PR_SUB; prec.push_back(make_int2(padj | kind, 0)); }
          else if (nc == 1) { kind = PR_ADD; prec.push_back(make_int2(padj | kind, 0)); }
          else if (nc == 2) {
            kind = PR_MUL_SIB;
            prec.push_back(make_int2(padj | kind, bslot[d->child_idx[cb + 1 - pos]] * PBB));
          } else {
            kind = PR_MUL_SIBS;
            prec.push_back(make_int2(padj | kind, 0));
            prec.push_back(make_int2((int)bedges.size(), nc | (pos << 16)));
            for (int e = 0; e < nc; ++e) bedges.push_back(bslot[d->child_idx[cb + e]] * PBB);
            cr.w |= CR_WIDE;
          }
        } else {
          if (k == VAEL_NODE_PROD) { kind = PR_ADD; prec.push_back(make_int2(padj | kind, 0)); }
          else {
            kind = PR_LSE;
            prec.push_back(make_int2(padj | kind, bslot[c] * PBB));
            prec.push_back(make_int2(bslot[p] * PBB, 0));
            cr.w |= CR_WIDE;
          }
        }
        uniform = first ? kind : (uniform == kind ? kind : -2);
        first = false;
      }
      if (uniform >= 0) cr.w |= (1 + uniform) << CR_UNIFORM_SHIFT;
      crec.push_back(cr);
    }
    // children with the same shape next to each other -> one segment per shape
    std::stable_sort(crec.begin() + lvl_beg, crec.end(), [](const int4& a, const int4& b) {
      return a.w != b.w ? a.w < b.w : a.z < b.z;
    });
    for (size_t i = lvl_beg; i < crec.size(); ++i) {
      if (i == lvl_beg || csegs.back().x != crec[i].w || csegs.back().w != crec[i].z)
        csegs.push_back(make_int4(crec[i].w, (int)i, 0, crec[i].z));
      csegs.back().z = (int)i + 1;
    }
  }
  g->cseg_level_ptr[d->n_levels] = (int)csegs.size();

  std::vector<int32_t> wseed((d->n_worlds + 3) / 4 * 4, N * PBB), qseed(d->n_queries), qonly;
  for (int w = 0; w < d->n_worlds; ++w) wseed[w] = d->world_node[w] * PBB;
  for (int q = 0; q < d->n_queries; ++q) {
    qseed[q] = d->query_node[q] * PBB;
    if (is_world[d->query_node[q]]) {
      qseed[q] |= 1;
      g->q_world_any = 1;
    } else if (std::find(qonly.begin(), qonly.end(), qseed[q]) == qonly.end()) {
      qonly.push_back(qseed[q]);
    }
  }
  g->n_qonly = (int)qonly.size();
  std::vector<int32_t> leaf_ptr(d->n_facts + 1, 0);
  std::vector<int2> leaf_rec;
  for (int f = 0; f < d->n_facts; ++f) {
    for (int n = 0; n < d->level_ptr[1]; ++n) {
      const int k = d->node_kind[n];
      if ((k == VAEL_NODE_LEAF_POS || k == VAEL_NODE_LEAF_NEG) && d->node_arg[n] == f && live[n] &&
          (seed[n] || !parents[n].empty()))
        leaf_rec.push_back(make_int2(n * PBB, k == VAEL_NODE_LEAF_NEG ? -1 : 1));
    }
    leaf_ptr[f + 1] = (int32_t)leaf_rec.size();
  }


  g->n_bops = (int)bops.size(); g->n_bsegs = (int)bsegs.size(); g->n_crec = (int)crec.size(); g->n_csegs = (int)csegs.size();
  g->n_prec = (int)prec.size();
  bedges.resize((bedges.size() + 3) / 4 * 4, 0);
  g->n_bedges = (int)bedges.size();
  {
    const size_t sched = (size_t)16 * (g->n_bops + g->n_bsegs + g->n_crec + g->n_csegs) + (size_t)8 * g->n_prec +
                         (size_t)4 * (g->n_bedges + (d->n_worlds + 3) / 4 * 4);
    static const int sched_limit = [] { const char* v = getenv("VAEL_GEN_SCHED_SMEM"); return (v && *v) ? atoi(v) : 12288; }();
    if (sched <= (size_t)sched_limit && g->bwd_smem + sched <= kSmemLimit && g->cb == std::max(d->n_worlds, 1)) {
      g->bsched_in_smem = 1;
      g->bwd_smem += sched;
    }
  }

  cudaError_t e = cudaSuccess;
  e = up(&g->ops, ops, e); e = up(&g->edges, edges, e); e = up(&g->segs, segs, e); e = up(&g->bsegs, bsegs, e);
  e = up(&g->csegs, csegs, e); e = up(&g->world_off, world_off, e);
  e = up(&g->query_off, query_off, e); e = up(&g->bops, bops, e); e = up(&g->bedges, bedges, e);
  e = up(&g->crec, crec, e); e = up(&g->prec, prec, e); e = up(&g->wseed_off, wseed, e);
  e = up(&g->qseed_off, qseed, e); e = up(&g->qonly_off, qonly, e); e = up(&g->leaf_rec, leaf_rec, e);
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
  return (backward ? g->bwd_smem : g->fwd_smem) <= kSmemLimit && (backward ? g->cb : g->cf) >= 4;
}

template <typename K>
static cudaError_t launch_generic(K kern, const GenProgram* g, const GenCall& C, size_t smem, int R, int num_sms,
                                  cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // CTAs per SM by shared memory; warps per CTA so that an SM holds ~`want` warps (latency hiding without
  // tensor-core-style pipelining: every warp is an independent interpreter of its share of a level)
  int per_sm = (int)std::min<size_t>(16, ((size_t)227 * 1024) / (smem + 1024));
  per_sm = std::max(1, per_sm);
  static const int want = [] { const char* v = getenv("VAEL_GEN_WARPS_PER_SM"); return (v && *v) ? atoi(v) : 24; }();
  int nw = std::max(1, std::min(16, (want + per_sm - 1) / per_sm));
  const int T = 32 * R;
  const long long n_tiles = (C.B + T - 1) / T;
  const long long grid = std::min<long long>(n_tiles, (long long)num_sms * per_sm);
  GenArgs A;
  A.G = *g;
  A.C = C;
  kern<<<(unsigned)grid, nw * 32, smem, st>>>(A);
  return cudaGetLastError();
}

template <bool LOG, bool SS>
static cudaError_t gen_forward_r(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st) {
  switch (g->rf) {
    case 4: return launch_generic(generic_forward_kernel<4, LOG, SS>, g, C, g->fwd_smem, 4, num_sms, st);
    case 2: return launch_generic(generic_forward_kernel<2, LOG, SS>, g, C, g->fwd_smem, 2, num_sms, st);
    default: return launch_generic(generic_forward_kernel<1, LOG, SS>, g, C, g->fwd_smem, 1, num_sms, st);
  }
}
template <bool LOG, bool SS>
static cudaError_t gen_backward_r(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st) {
  switch (g->rb) {
    case 4: return launch_generic(generic_backward_kernel<4, LOG, SS>, g, C, g->bwd_smem, 4, num_sms, st);
    case 2: return launch_generic(generic_backward_kernel<2, LOG, SS>, g, C, g->bwd_smem, 2, num_sms, st);
    default: return launch_generic(generic_backward_kernel<1, LOG, SS>, g, C, g->bwd_smem, 1, num_sms, st);
  }
}
cudaError_t gen_forward(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st) {
  const bool lg = g->semiring == VAEL_SEMIRING_LOG;
  if (g->sched_in_smem) return lg ? gen_forward_r<true, true>(g, C, num_sms, st) : gen_forward_r<false, true>(g, C, num_sms, st);
  return lg ? gen_forward_r<true, false>(g, C, num_sms, st) : gen_forward_r<false, false>(g, C, num_sms, st);
}
cudaError_t gen_backward(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st) {
  const bool lg = g->semiring == VAEL_SEMIRING_LOG;
  if (g->bsched_in_smem) return lg ? gen_backward_r<true, true>(g, C, num_sms, st) : gen_backward_r<false, true>(g, C, num_sms, st);
  return lg ? gen_backward_r<true, false>(g, C, num_sms, st) : gen_backward_r<false, false>(g, C, num_sms, st);
}

}  // namespace vael
