// Generic evaluator: level-by-level sweep over the flat, level-ordered CSR circuit.
//
// Reference semantics: problog's sdd.evaluate(semiring=GraphSemiring) walking the
// compiled formula node by node and calling GraphSemiring.plus/times/negate/value/
// normalize for every node (utils/graph_semiring.py:13-60, models/vael.py:98-135),
// evaluated here ELEMENTWISE per batch row (the reference's batch-wide `.any()`
// identity shortcuts, utils/graph_semiring.py:27-49, are a documented divergence:
// see DESIGN.md "Parity domain").
//
// Execution model: a CTA owns a tile of 32 batch rows; lane <-> row.  Node values
// live in shared memory node-major / row-minor (val[node][33]): every warp-wide
// access touches 32 consecutive words (conflict-free whatever the circuit shape),
// node descriptors are warp-uniform 128-bit broadcast loads, warps split the
// nodes of a level and meet at one __syncthreads per level.  Batch-major global
// tensors are transposed on the way in/out with coalesced accesses (lane <-> world
// for [B,W]).  The backward is the reverse sweep in gather form over the parent
// lists (transposed CSR) built at circuit creation, so it needs no atomics and is
// deterministic.
#include "params.cuh"

namespace vael {

constexpr int RS = 33;  // row stride of the SoA tiles (odd -> transposed reads are conflict-free too)


__device__ __forceinline__ float leaf_value(int kind, float p, int semiring) {
  float v = (kind == VAEL_NODE_LEAF_NEG) ? __fsub_rn(1.0f, p) : p;
  if (semiring == VAEL_SEMIRING_LOG) v = logf(v);
  return v;
}

// value of an operator node from its children's values; VAL(c) reads child c for this lane
template <typename ValFn>
__device__ __forceinline__ float eval_op(int kind, int nc, const NodeDesc& d, const int32_t* edges, int semiring,
                                         ValFn VAL) {
  if (nc == 0) {
    if (kind == VAEL_NODE_PROD) return semiring == VAEL_SEMIRING_LOG ? 0.0f : 1.0f;
    if (kind == VAEL_NODE_COMPL) return 1.0f;
    return semiring == VAEL_SEMIRING_LOG ? -INFINITY : 0.0f;
  }
  if (semiring == VAEL_SEMIRING_PROB) {
    if (nc <= 2) {
      const float x = VAL(d.a);
      if (nc == 1) return kind == VAEL_NODE_COMPL ? __fsub_rn(1.0f, x) : x;
      const float y = VAL(d.b);
      if (kind == VAEL_NODE_PROD) return __fmul_rn(x, y);
      const float s = __fadd_rn(x, y);
      return kind == VAEL_NODE_COMPL ? __fsub_rn(1.0f, s) : s;
    }
    float acc = VAL(__ldg(edges + d.a));
    if (kind == VAEL_NODE_PROD) {
      for (int e = 1; e < nc; ++e) acc = __fmul_rn(acc, VAL(__ldg(edges + d.a + e)));
      return acc;
    }
    for (int e = 1; e < nc; ++e) acc = __fadd_rn(acc, VAL(__ldg(edges + d.a + e)));
    return kind == VAEL_NODE_COMPL ? __fsub_rn(1.0f, acc) : acc;
  }
  // log semiring: times = +, plus = logsumexp
  if (kind == VAEL_NODE_PROD) {
    if (nc <= 2) {
      const float x = VAL(d.a);
      return nc == 1 ? x : __fadd_rn(x, VAL(d.b));
    }
    float acc = VAL(__ldg(edges + d.a));
    for (int e = 1; e < nc; ++e) acc = __fadd_rn(acc, VAL(__ldg(edges + d.a + e)));
    return acc;
  }
  // SUM -> logsumexp (COMPL is rejected for the log semiring at circuit creation)
  float mx = -INFINITY;
  for (int e = 0; e < nc; ++e) {
    const int c = nc <= 2 ? (e == 0 ? d.a : d.b) : __ldg(edges + d.a + e);
    mx = fmaxf(mx, VAL(c));
  }
  if (mx == -INFINITY) return -INFINITY;
  float s = 0.0f;
  for (int e = 0; e < nc; ++e) {
    const int c = nc <= 2 ? (e == 0 ? d.a : d.b) : __ldg(edges + d.a + e);
    s += expf(VAL(c) - mx);
  }
  return mx + logf(s);
}

__device__ __forceinline__ NodeDesc load_desc(const NodeDesc* nodes, int n) {
  const int4 t = __ldg(reinterpret_cast<const int4*>(nodes) + n);
  NodeDesc d;
  d.kind_n = t.x; d.a = t.y; d.b = t.z; d.slot = t.w;
  return d;
}

// transposed load of the facts tile: global [rows,F] row-major -> fs[f*RS + r]
__device__ __forceinline__ void load_facts_tile(float* fs, const float* facts, long long row0, int rows, int F) {
  const float* src = facts + row0 * F;
  for (int i = threadIdx.x; i < kWarp * F; i += blockDim.x) {
    const int r = i / F, f = i - r * F;
    fs[f * RS + r] = (r < rows) ? __ldg(src + i) : 0.5f;  // padding rows get a harmless in-domain value
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) generic_forward_kernel(const GenParams P) {
  extern __shared__ __align__(16) float gsm[];
  float* fs = gsm;                       // [F][RS]
  float* val = gsm + P.n_facts * RS;     // [n_nodes][RS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long n_tiles = (P.B + kWarp - 1) / kWarp;
  const bool norm = (P.flags & VAEL_FLAG_NORMALIZE) && P.norm_node >= 0;
  const bool evid = (P.flags & VAEL_FLAG_EVIDENCE) != 0;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    load_facts_tile(fs, P.facts, row0, rows, P.n_facts);
    __syncthreads();

    for (int L = 0; L < P.n_levels; ++L) {
      for (int n = P.level_ptr[L] + warp; n < P.level_ptr[L + 1]; n += nwarps) {
        const NodeDesc d = load_desc(P.nodes, n);
        const int kind = d.kind_n & 0xff, nc = d.kind_n >> 8;
        float v;
        if (kind <= VAEL_NODE_LEAF_NEG) {
          v = leaf_value(kind, fs[d.a * RS + lane], P.semiring);
        } else if (kind == VAEL_NODE_CONST) {
          v = __int_as_float(d.a);
        } else {
          v = eval_op(kind, nc, d, P.edges, P.semiring, [&](int c) { return val[c * RS + lane]; });
        }
        val[n * RS + lane] = v;
      }
      __syncthreads();
    }

    // ---- outputs: one warp per batch row, lane <-> world (coalesced stores) ----
    if (P.worlds != nullptr) {
      for (int r = warp; r < rows; r += nwarps) {
        const long long row = row0 + r;
        float z = 1.0f, s = 1.0f;
        uint32_t mword = 0;
        const uint32_t* mrow = nullptr;
        if (norm) z = val[P.norm_node * RS + r];
        if (evid) {
          const int32_t q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
          const bool qok = q >= 0 && q < P.n_queries;
          s = qok ? val[__ldg(P.query_node + q) * RS + r] : 0.0f;
          mrow = qok ? P.qmask + (long long)q * P.mask_words : nullptr;
        }
        for (int e0 = 0; e0 < P.n_worlds; e0 += kWarp) {
          const int e = e0 + lane;
          if (evid) mword = mrow ? __ldg(mrow + (e0 >> 5)) : 0u;
          if (e < P.n_worlds) {
            float v = val[__ldg(P.world_node + e) * RS + r];
            if (evid) {
              v = ((mword >> lane) & 1u) ? v / s : 0.0f;
            } else if (norm) {
              v = v / z;
            }
            P.worlds[row * P.n_worlds + e] = v;
          }
        }
      }
    }
    if (P.queries_all != nullptr) {
      for (int r = warp; r < rows; r += nwarps) {
        const long long row = row0 + r;
        const float z = norm ? val[P.norm_node * RS + r] : 1.0f;
        for (int q = lane; q < P.n_queries; q += kWarp) {
          float v = val[__ldg(P.query_node + q) * RS + r];
          if (norm) v = v / z;
          P.queries_all[row * P.n_queries + q] = v;
        }
      }
    }
    if (P.query != nullptr && warp == nwarps - 1 && lane < rows) {
      const long long row = row0 + lane;
      const int32_t q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
      float v = 0.0f;
      if (q >= 0 && q < P.n_queries) {
        v = val[__ldg(P.query_node + q) * RS + lane];
        if (norm) v = v / val[P.norm_node * RS + lane];
      }
      P.query[row] = v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// backward: forward values of the nodes the adjoint needs, then reverse sweep
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) generic_backward_kernel(const GenParams P) {
  extern __shared__ __align__(16) float gsm[];
  float* fs = gsm;                               // [F][RS]
  float* val = fs + P.n_facts * RS;              // [n_bvals][RS]
  float* adj = val + P.n_bvals * RS;             // [n_nodes][RS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long n_tiles = (P.B + kWarp - 1) / kWarp;
  const bool norm = (P.flags & VAEL_FLAG_NORMALIZE) && P.norm_node >= 0;
  const bool is_log = P.semiring == VAEL_SEMIRING_LOG;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kWarp;
    const int rows = (int)min((long long)kWarp, P.B - row0);
    load_facts_tile(fs, P.facts, row0, rows, P.n_facts);
    for (int i = threadIdx.x; i < P.n_nodes * RS; i += blockDim.x) adj[i] = 0.0f;
    __syncthreads();

    // forward values of the kept nodes
    for (int L = 0; L < P.n_levels; ++L) {
      for (int n = P.level_ptr[L] + warp; n < P.level_ptr[L + 1]; n += nwarps) {
        const int slot = __ldg(P.bslot + n);
        if (slot < 0) continue;
        const NodeDesc d = load_desc(P.nodes, n);
        const int kind = d.kind_n & 0xff, nc = d.kind_n >> 8;
        float v;
        if (kind <= VAEL_NODE_LEAF_NEG) {
          v = leaf_value(kind, fs[d.a * RS + lane], P.semiring);
        } else if (kind == VAEL_NODE_CONST) {
          v = __int_as_float(d.a);
        } else {
          v = eval_op(kind, nc, d, P.edges, P.semiring,
                      [&](int c) { return val[__ldg(P.bslot + c) * RS + lane]; });
        }
        val[slot * RS + lane] = v;
      }
      __syncthreads();
    }

    // seed adjoints from the upstream gradients
    if (P.grad_worlds != nullptr) {
      for (int r = warp; r < rows; r += nwarps) {
        const long long row = row0 + r;
        const float z = norm ? val[__ldg(P.bslot + P.norm_node) * RS + r] : 1.0f;
        for (int e = lane; e < P.n_worlds; e += kWarp) {
          float g = __ldg(P.grad_worlds + row * P.n_worlds + e);
          if (norm) g = g / z;
          adj[__ldg(P.world_node + e) * RS + r] = g;  // world nodes are distinct (checked at creation)
        }
      }
    }
    __syncthreads();
    if (P.grad_query != nullptr && warp == 0 && lane < rows) {
      const long long row = row0 + lane;
      const int32_t q = P.qidx ? __ldg(P.qidx + row) : P.qscalar;
      if (q >= 0 && q < P.n_queries) {
        float g = __ldg(P.grad_query + row);
        if (norm) g = g / val[__ldg(P.bslot + P.norm_node) * RS + lane];
        adj[__ldg(P.query_node + q) * RS + lane] += g;
      }
    }
    __syncthreads();

    // reverse sweep, gather form: adj[c] += sum over parents p of adj[p] * d val[p] / d val[c]
    for (int L = P.n_levels - 2; L >= 0; --L) {
      for (int c = P.level_ptr[L] + warp; c < P.level_ptr[L + 1]; c += nwarps) {
        const int pb = __ldg(P.par_ptr + c), pe = __ldg(P.par_ptr + c + 1);
        if (pb == pe) continue;
        float a = adj[c * RS + lane];
        for (int i = pb; i < pe; ++i) {
          const int2 rec = __ldg(P.par_rec + i);
          const int p = rec.x;
          const float ap = adj[p * RS + lane];
          const NodeDesc d = load_desc(P.nodes, p);
          const int kind = d.kind_n & 0xff, nc = d.kind_n >> 8;
          if (!is_log) {
            if (kind == VAEL_NODE_SUM) {
              a += ap;
            } else if (kind == VAEL_NODE_COMPL) {
              a -= ap;
            } else {  // PROD: product of the siblings; rec.y = position of c among p's children
              float sib = 1.0f;
              if (nc == 2) {
                sib = val[__ldg(P.bslot + (rec.y == 0 ? d.b : d.a)) * RS + lane];
              } else if (nc > 2) {
                for (int e = 0; e < nc; ++e) {
                  if (e == rec.y) continue;
                  sib *= val[__ldg(P.bslot + __ldg(P.edges + d.a + e)) * RS + lane];
                }
              }
              a = fmaf(ap, sib, a);
            }
          } else {
            if (kind == VAEL_NODE_PROD) {
              a += ap;
            } else {  // logsumexp: softmax weight of c inside p
              const float vc = val[__ldg(P.bslot + c) * RS + lane];
              const float vp = val[__ldg(P.bslot + p) * RS + lane];
              a = fmaf(ap, (vp == -INFINITY) ? 0.0f : expf(vc - vp), a);
            }
          }
        }
        adj[c * RS + lane] = a;
      }
      __syncthreads();
    }

    // leaves -> grad_facts[row, f], coalesced
    for (int i = threadIdx.x; i < rows * P.n_facts; i += blockDim.x) {
      const int r = i / P.n_facts, f = i - r * P.n_facts;
      const float pf = fs[f * RS + r];
      float g = 0.0f;
      for (int k = __ldg(P.fact_leaf_ptr + f); k < __ldg(P.fact_leaf_ptr + f + 1); ++k) {
        const int leaf = __ldg(P.fact_leaf + k);
        const int kind = __ldg(reinterpret_cast<const int*>(P.nodes + leaf)) & 0xff;
        const float a = adj[leaf * RS + r];
        if (!is_log) {
          g += (kind == VAEL_NODE_LEAF_NEG) ? -a : a;
        } else {
          g += (kind == VAEL_NODE_LEAF_NEG) ? -a / (1.0f - pf) : a / pf;
        }
      }
      P.grad_facts[row0 * P.n_facts + i] = g;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static size_t gen_fwd_smem(const GenParams& P) { return (size_t)(P.n_facts + P.n_nodes) * RS * sizeof(float); }
static size_t gen_bwd_smem(const GenParams& P) {
  return (size_t)(P.n_facts + P.n_bvals + P.n_nodes) * RS * sizeof(float);
}

bool generic_fits(const GenParams& P, bool backward) {
  return (backward ? gen_bwd_smem(P) : gen_fwd_smem(P)) <= (size_t)227 * 1024;
}

static cudaError_t launch_generic(void (*kern)(const GenParams), const GenParams& P, size_t smem, int num_sms,
                                  cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int threads = 256;
  int per_sm = (int)min((size_t)8, ((size_t)227 * 1024) / (smem + 1024));
  per_sm = max(1, per_sm);
  const long long n_tiles = (P.B + kWarp - 1) / kWarp;
  const long long grid = min(n_tiles, (long long)num_sms * per_sm);
  kern<<<(unsigned)grid, threads, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t generic_forward(const GenParams& P, int num_sms, cudaStream_t st) {
  return launch_generic(generic_forward_kernel, P, gen_fwd_smem(P), num_sms, st);
}
cudaError_t generic_backward(const GenParams& P, int num_sms, cudaStream_t st) {
  return launch_generic(generic_backward_kernel, P, gen_bwd_smem(P), num_sms, st);
}

}  // namespace vael
