// C-ABI of libvael_b200.so (include/vael_b200.h): circuit upload + validation,
// kernel dispatch, host-buffer (e2e) variants.  No torch types anywhere.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "params.cuh"

using namespace vael;

// ---------------------------------------------------------------------------
// error / bookkeeping
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local const char* g_last_kernel = "";
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(VAEL_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                        \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return cuda_fail(_e, #call); \
  } while (0)

static int device_sms(int* out) {
  static int cached[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 64 && cached[dev]) { *out = cached[dev]; return VAEL_OK; }
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (dev < 64) cached[dev] = n;
  *out = n;
  return VAEL_OK;
}

extern "C" int vael_abi_version(void) { return VAEL_B200_ABI_VERSION; }
extern "C" const char* vael_last_error(void) { return g_err.c_str(); }
extern "C" int64_t vael_launch_count(void) { return g_launches.load(); }
extern "C" const char* vael_last_kernel(void) { return g_last_kernel; }

// ---------------------------------------------------------------------------
// circuit creation
// ---------------------------------------------------------------------------
template <typename T>
static cudaError_t upload(T** dptr, const std::vector<T>& h) {
  *dptr = nullptr;
  if (h.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dptr), h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static void fnv(uint64_t& h, const void* data, size_t n) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
}

// Is `z` the ProbLog normaliser of K annotated disjunctions over the fact column ranges
// [offs[a], offs[a]+sizes[a]):  PROD_a( SUM(leaves of AD a..., COMPL(leaves of AD a...)) ) ?
static bool norm_is_adk(const vael_circuit_desc_t* d, int z, int K, const int* sizes, const int* offs) {
  auto kind = [&](int n) { return (int)d->node_kind[n]; };
  auto nch = [&](int n) { return d->child_ptr[n + 1] - d->child_ptr[n]; };
  auto ch = [&](int n, int i) { return d->child_idx[d->child_ptr[n] + i]; };
  auto leaves_are = [&](int n, int first, int count, int off) {
    for (int i = 0; i < count; ++i) {
      const int c = ch(n, first + i);
      if (kind(c) != VAEL_NODE_LEAF_POS || d->node_arg[c] != off + i) return false;
    }
    return true;
  };
  if (kind(z) != VAEL_NODE_PROD || nch(z) != K) return false;
  for (int a = 0; a < K; ++a) {
    const int s = ch(z, a);
    if (kind(s) != VAEL_NODE_SUM || nch(s) != sizes[a] + 1) return false;
    if (!leaves_are(s, 0, sizes[a], offs[a])) return false;
    const int c = ch(s, sizes[a]);
    if (kind(c) != VAEL_NODE_COMPL || nch(c) != sizes[a]) return false;
    if (!leaves_are(c, 0, sizes[a], offs[a])) return false;
  }
  return true;
}

// Fact columns of the literals of a world node that is a left-deep product of positive literals,
// PROD(PROD(PROD(l1, l2), l3), ...) — the order the kernels multiply in.  false if it is anything else.
static bool world_literals(const vael_circuit_desc_t* d, int n, std::vector<int>& cols) {
  auto kind = [&](int x) { return (int)d->node_kind[x]; };
  auto nch = [&](int x) { return d->child_ptr[x + 1] - d->child_ptr[x]; };
  auto ch = [&](int x, int i) { return d->child_idx[d->child_ptr[x] + i]; };
  std::vector<int> rev;
  int cur = n;
  for (int depth = 0; depth < 16; ++depth) {
    if (kind(cur) != VAEL_NODE_PROD || nch(cur) < 2) return false;
    for (int e = nch(cur) - 1; e >= 1; --e) {
      const int c = ch(cur, e);
      if (kind(c) != VAEL_NODE_LEAF_POS) return false;
      rev.push_back(d->node_arg[c]);
    }
    const int first = ch(cur, 0);
    if (kind(first) == VAEL_NODE_LEAF_POS) {
      rev.push_back(d->node_arg[first]);
      cols.assign(rev.rbegin(), rev.rend());
      return true;
    }
    cur = first;
  }
  return false;
}

static void destroy_slot_streams(vael_host_slot& S) {
  if (S.facts_up) cudaEventDestroy(S.facts_up);
  if (S.grads_up) cudaEventDestroy(S.grads_up);
  if (S.done) cudaEventDestroy(S.done);
  if (S.up) cudaStreamDestroy(S.up);
  if (S.stream) cudaStreamDestroy(S.stream);
}

extern "C" void vael_circuit_destroy(vael_circuit_t* c) {
  if (!c) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  gen_program_destroy(c->gen);
  jit_destroy(c->jit);
  cudaFree(c->d_qmask); cudaFree(c->d_qm_ptr); cudaFree(c->d_qm_idx); cudaFree(c->d_qmask_c);
  cudaFree(c->d_bad_query);
  if (c->h_bad_query) cudaFreeHost(c->h_bad_query);
  for (auto& S : c->slot) {
    cudaFree(S.facts); cudaFree(S.worlds); cudaFree(S.query); cudaFree(S.gw); cudaFree(S.gq); cudaFree(S.gf);
    cudaFree(S.qidx);
    destroy_slot_streams(S);
  }
  cudaSetDevice(prev);
  delete c;
}

extern "C" int vael_circuit_create(const vael_circuit_desc_t* d, vael_circuit_t** out) {
  if (!d || !out) return fail(VAEL_E_INVALID, "null argument");
  *out = nullptr;
  if (d->n_facts <= 0 || d->n_nodes <= 0 || d->n_levels <= 0 || d->n_edges < 0 || d->n_worlds < 0 ||
      d->n_queries < 0)
    return fail(VAEL_E_INVALID, "non-positive circuit dimension");
  if (d->n_levels > kMaxLevels) return fail(VAEL_E_UNSUPPORTED, "too many levels (max 24)");
  if (d->semiring != VAEL_SEMIRING_PROB && d->semiring != VAEL_SEMIRING_LOG)
    return fail(VAEL_E_INVALID, "unknown semiring");
  if (!d->node_kind || !d->node_arg || !d->level_ptr || !d->child_ptr || (d->n_edges && !d->child_idx) ||
      (d->n_worlds && !d->world_node) || (d->n_queries && !d->query_node))
    return fail(VAEL_E_INVALID, "null circuit array");
  if (d->level_ptr[0] != 0 || d->level_ptr[d->n_levels] != d->n_nodes)
    return fail(VAEL_E_INVALID, "level_ptr does not cover the nodes");
  if (d->child_ptr[0] != 0 || d->child_ptr[d->n_nodes] != d->n_edges)
    return fail(VAEL_E_INVALID, "child_ptr does not cover the edges");
  std::vector<int> level_of(d->n_nodes, 0);
  for (int L = 0; L < d->n_levels; ++L) {
    if (d->level_ptr[L + 1] < d->level_ptr[L]) return fail(VAEL_E_INVALID, "level_ptr not monotone");
    for (int n = d->level_ptr[L]; n < d->level_ptr[L + 1]; ++n) level_of[n] = L;
  }
  for (int n = 0; n < d->n_nodes; ++n) {
    const int k = d->node_kind[n];
    const int nc = d->child_ptr[n + 1] - d->child_ptr[n];
    if (nc < 0) return fail(VAEL_E_INVALID, "child_ptr not monotone");
    if (k == VAEL_NODE_LEAF_POS || k == VAEL_NODE_LEAF_NEG) {
      if (d->node_arg[n] < 0 || d->node_arg[n] >= d->n_facts) return fail(VAEL_E_INVALID, "leaf fact column out of range");
      if (nc != 0 || level_of[n] != 0) return fail(VAEL_E_INVALID, "leaves must be childless level-0 nodes");
    } else if (k == VAEL_NODE_CONST) {
      if (d->node_arg[n] < 0 || d->node_arg[n] >= d->n_consts || !d->const_val) return fail(VAEL_E_INVALID, "const index out of range");
      if (nc != 0) return fail(VAEL_E_INVALID, "const node with children");
    } else if (k == VAEL_NODE_PROD || k == VAEL_NODE_SUM || k == VAEL_NODE_COMPL) {
      if (k == VAEL_NODE_COMPL && d->semiring == VAEL_SEMIRING_LOG)
        return fail(VAEL_E_UNSUPPORTED, "COMPL nodes are not defined for the log semiring");
      if (nc > 0xffffff) return fail(VAEL_E_UNSUPPORTED, "fan-in too large");
      for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) {
        const int c = d->child_idx[e];
        if (c < 0 || c >= d->n_nodes) return fail(VAEL_E_INVALID, "child index out of range");
        if (level_of[c] >= level_of[n]) return fail(VAEL_E_INVALID, "child is not in a strictly lower level");
      }
    } else {
      return fail(VAEL_E_INVALID, "unknown node kind");
    }
  }
  std::vector<int> world_pos(d->n_nodes, -1);
  for (int w = 0; w < d->n_worlds; ++w) {
    const int n = d->world_node[w];
    if (n < 0 || n >= d->n_nodes) return fail(VAEL_E_INVALID, "world node out of range");
    if (world_pos[n] >= 0 && d->node_kind[n] != VAEL_NODE_CONST)
      return fail(VAEL_E_INVALID, "world nodes must be distinct (constants excepted)");
    world_pos[n] = w;
  }
  for (int q = 0; q < d->n_queries; ++q)
    if (d->query_node[q] < 0 || d->query_node[q] >= d->n_nodes) return fail(VAEL_E_INVALID, "query node out of range");
  if (d->norm_node >= d->n_nodes) return fail(VAEL_E_INVALID, "norm node out of range");

  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  if (int rc = device_sms(&sms)) return rc;

  vael_circuit_t* c = new vael_circuit_t();   // value-initialised: every plain member is zero
  c->device = dev; c->num_sms = sms;
  c->n_facts = d->n_facts; c->n_nodes = d->n_nodes; c->n_levels = d->n_levels; c->n_edges = d->n_edges;
  c->n_worlds = d->n_worlds; c->n_queries = d->n_queries; c->semiring = d->semiring;
  c->norm_node = d->norm_node < 0 ? -1 : d->norm_node;
  c->n_consts = d->n_consts;
  c->h_node_kind.assign(d->node_kind, d->node_kind + d->n_nodes);
  c->h_node_arg.assign(d->node_arg, d->node_arg + d->n_nodes);
  c->h_level_ptr.assign(d->level_ptr, d->level_ptr + d->n_levels + 1);
  c->h_child_ptr.assign(d->child_ptr, d->child_ptr + d->n_nodes + 1);
  if (d->n_edges) c->h_child_idx.assign(d->child_idx, d->child_idx + d->n_edges);
  if (d->n_consts) c->h_const_val.assign(d->const_val, d->const_val + d->n_consts);
  if (d->n_worlds) c->h_world_node.assign(d->world_node, d->world_node + d->n_worlds);
  if (d->n_queries) c->h_query_node.assign(d->query_node, d->query_node + d->n_queries);

  // identity hash of the arrays as given
  uint64_t h = 1469598103934665603ull;
  const int32_t hdr[9] = {d->n_facts, d->n_nodes, d->n_levels, d->n_edges, d->n_worlds, d->n_queries, d->n_consts, d->semiring, c->norm_node};
  fnv(h, hdr, sizeof(hdr));
  fnv(h, d->node_kind, d->n_nodes);
  fnv(h, d->node_arg, sizeof(int32_t) * d->n_nodes);
  fnv(h, d->level_ptr, sizeof(int32_t) * (d->n_levels + 1));
  fnv(h, d->child_ptr, sizeof(int32_t) * (d->n_nodes + 1));
  if (d->n_edges) fnv(h, d->child_idx, sizeof(int32_t) * d->n_edges);
  if (d->n_consts) fnv(h, d->const_val, sizeof(float) * d->n_consts);
  if (d->n_worlds) fnv(h, d->world_node, sizeof(int32_t) * d->n_worlds);
  if (d->n_queries) fnv(h, d->query_node, sizeof(int32_t) * d->n_queries);
  c->hash = h;

  // query -> world membership bit masks (only if every query is a SUM over world nodes)
  c->mask_words = (d->n_worlds + 31) / 32;
  std::vector<uint32_t> qmask((size_t)d->n_queries * c->mask_words, 0u);
  bool masks_ok = d->n_queries > 0 && d->n_worlds > 0 && d->semiring == VAEL_SEMIRING_PROB;
  for (int q = 0; q < d->n_queries && masks_ok; ++q) {
    const int n = d->query_node[q];
    if (d->node_kind[n] == VAEL_NODE_CONST) continue;  // an unsatisfiable query: empty mask (const must be 0)
    if (world_pos[n] >= 0) {  // a query that holds in exactly one world is that world's node
      qmask[(size_t)q * c->mask_words + (world_pos[n] >> 5)] |= 1u << (world_pos[n] & 31);
      continue;
    }
    if (d->node_kind[n] != VAEL_NODE_SUM) { masks_ok = false; break; }
    for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) {
      const int w = world_pos[d->child_idx[e]];
      if (w < 0 || d->node_kind[d->child_idx[e]] == VAEL_NODE_CONST) { masks_ok = false; break; }
      qmask[(size_t)q * c->mask_words + (w >> 5)] |= 1u << (w & 31);
    }
  }
  // the fast path sums members in world order; the CSR children must be listed in that order too
  bool members_in_world_order = masks_ok;
  for (int q = 0; q < d->n_queries && members_in_world_order; ++q) {
    const int n = d->query_node[q];
    if (d->node_kind[n] != VAEL_NODE_SUM) continue;
    int prev = -1;
    for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1]; ++e) {
      const int w = world_pos[d->child_idx[e]];
      if (w <= prev) { members_in_world_order = false; break; }
      prev = w;
    }
  }

  // structured fast paths: world w = left-deep product of one positive literal per annotated disjunction,
  // worlds in mixed-radix order (first AD most significant), AD fact columns contiguous and in order
  c->fast_kind = 0;
  std::vector<uint32_t> qmask_c;
  if (d->semiring == VAEL_SEMIRING_PROB && masks_ok && members_in_world_order && d->n_worlds > 0) {
    std::vector<int> cols;
    int K = 0, sizes[4] = {0, 0, 0, 0}, offs[4] = {0, 0, 0, 0};
    bool ok = world_literals(d, d->world_node[0], cols) && cols.size() >= 2 && cols.size() <= 3 && cols[0] == 0;
    if (ok) {
      K = (int)cols.size();
      long long prod = 1;
      for (int k = 0; k < K; ++k) {
        offs[k] = cols[k];
        sizes[k] = (k + 1 < K ? cols[k + 1] : d->n_facts) - cols[k];
        ok = ok && sizes[k] > 0;
        prod *= sizes[k] > 0 ? sizes[k] : 1;
      }
      ok = ok && prod == d->n_worlds;
    }
    for (int w = 0; w < d->n_worlds && ok; ++w) {
      ok = world_literals(d, d->world_node[w], cols) && (int)cols.size() == K;
      int rem = w;
      for (int k = K - 1; k >= 0 && ok; --k) {
        ok = cols[k] == offs[k] + rem % sizes[k];
        rem /= sizes[k];
      }
    }
    if (ok && K == 2 && ad2_supported(sizes[0], sizes[1])) {
      c->fast_kind = 2; c->ad_na = sizes[0]; c->ad_nb = sizes[1];
    } else if (ok && K == 3 && adw_supported(sizes[0], sizes[1], sizes[2])) {
      c->fast_kind = 3; c->ad_np = sizes[0]; c->ad_na = sizes[1]; c->ad_nb = sizes[2];
      // chunked masks: chunk i = worlds [i*C, (i+1)*C), bit (w % C) of its MWC words
      const int C = sizes[1] * sizes[2], MWC = (C + 31) / 32;
      qmask_c.assign((size_t)d->n_queries * sizes[0] * MWC, 0u);
      for (int q = 0; q < d->n_queries; ++q)
        for (int w = 0; w < d->n_worlds; ++w)
          if ((qmask[(size_t)q * c->mask_words + (w >> 5)] >> (w & 31)) & 1u)
            qmask_c[((size_t)q * sizes[0] + w / C) * MWC + ((w % C) >> 5)] |= 1u << ((w % C) & 31);
    }
    if (c->fast_kind) c->norm_fast_ok = c->norm_node >= 0 && norm_is_adk(d, c->norm_node, K, sizes, offs);
  }

  // literal table in the log semiring (circuit_lt.cu): no queries, no normaliser, every world a product over all F
  // fact columns — the positive literals in column order, then the negative ones (compiler.py:circuit_from_world_table)
  if (c->fast_kind == 0 && d->semiring == VAEL_SEMIRING_LOG && d->n_queries == 0 && c->norm_node < 0 &&
      lt_supported(d->n_facts, d->n_worlds)) {
    bool ok = true;
    const uint32_t all = d->n_facts >= 32 ? 0xffffffffu : ((1u << d->n_facts) - 1u);
    for (int w = 0; w < d->n_worlds && ok; ++w) {
      const int n = d->world_node[w];
      ok = d->node_kind[n] == VAEL_NODE_PROD && d->child_ptr[n + 1] - d->child_ptr[n] == d->n_facts;
      uint32_t pos = 0, neg = 0;
      int prev_pos = -1, prev_neg = -1;
      for (int e = d->child_ptr[n]; e < d->child_ptr[n + 1] && ok; ++e) {
        const int ch = d->child_idx[e], f = d->node_arg[ch];
        if (d->node_kind[ch] == VAEL_NODE_LEAF_POS) {
          ok = neg == 0 && f > prev_pos;
          prev_pos = f; pos |= 1u << f;
        } else if (d->node_kind[ch] == VAEL_NODE_LEAF_NEG) {
          ok = f > prev_neg;
          prev_neg = f; neg |= 1u << f;
        } else {
          ok = false;
        }
      }
      ok = ok && (pos & neg) == 0 && (pos | neg) == all;
      c->lt_mask[w] = pos;
    }
    if (ok && lt_masks_fit(d->n_facts, d->n_worlds, c->lt_mask)) c->fast_kind = 4;
  }

  cudaError_t e = cudaSuccess;
  auto up = [&](auto** dp, const auto& v) { if (e == cudaSuccess) e = upload(dp, v); };
  if (masks_ok) {
    up(&c->d_qmask, qmask);
    // member worlds of every query node in world order, as byte offsets of the world rows of the all-queries
    // kernel's world-major tile ([W + 1][32] floats); every list is padded to a multiple of four with the
    // offset of the all-zero row W
    std::vector<int32_t> qm_ptr(d->n_queries + 1, 0), qm_idx;
    for (int q = 0; q < d->n_queries; ++q) {
      for (int w = 0; w < d->n_worlds; ++w)
        if ((qmask[(size_t)q * c->mask_words + (w >> 5)] >> (w & 31)) & 1u) qm_idx.push_back(w * 128);
      while (qm_idx.size() % 4) qm_idx.push_back(d->n_worlds * 128);
      qm_ptr[q + 1] = (int32_t)qm_idx.size();
    }
    if (qm_idx.empty()) qm_idx.assign(4, d->n_worlds * 128);
    up(&c->d_qm_ptr, qm_ptr); up(&c->d_qm_idx, qm_idx);
    c->qm_len = (int32_t)qm_idx.size();
    // the same lists in the form the two-AD all-queries kernel takes in its parameter block: grouped by the
    // first AD's value i, `len[q]` slots per (q, i) (circuit_ad2.cu: ad2_queries_kernel)
    c->q_members_ok = false;
    if (c->fast_kind == 2 && d->n_queries <= kQMaxQueries) {
      const int NA = c->ad_na, NB = c->ad_nb;
      std::vector<uint16_t> slots;
      bool fits = true;
      for (int q = 0; q < d->n_queries && fits; ++q) {
        int L = 0;
        for (int i = 0; i < NA; ++i) {
          int n = 0;
          for (int k = 0; k < NB; ++k) {
            const int w = i * NB + k;
            n += (qmask[(size_t)q * c->mask_words + (w >> 5)] >> (w & 31)) & 1u;
          }
          L = std::max(L, n);
        }
        c->q_ptr[q] = (uint16_t)slots.size();
        c->q_len[q] = (uint8_t)L;
        for (int i = 0; i < NA; ++i) {
          int n = 0;
          for (int k = 0; k < NB; ++k) {
            const int w = i * NB + k;
            if ((qmask[(size_t)q * c->mask_words + (w >> 5)] >> (w & 31)) & 1u) { slots.push_back((uint16_t)(k * 128)); ++n; }
          }
          for (; n < L; ++n) slots.push_back((uint16_t)(NB * 128));
        }
        fits = slots.size() <= (size_t)kQMaxSlots;
      }
      if (fits) {
        memcpy(c->q_slot, slots.data(), slots.size() * sizeof(uint16_t));
        c->q_nslots = (int32_t)slots.size();
        c->q_members_ok = true;
      }
    }
    // addition-shaped queries over a two-AD outer product: query s = { (j, k) : j + k = s }
    c->queries_diagonal = false;
    if (c->fast_kind == 3 && d->n_queries == c->ad_np + c->ad_na + c->ad_nb - 2) {
      bool diag = true;
      for (int q = 0; q < d->n_queries && diag; ++q)
        for (int w = 0; w < d->n_worlds && diag; ++w) {
          const bool in = (qmask[(size_t)q * c->mask_words + (w >> 5)] >> (w & 31)) & 1u;
          const int i = w / (c->ad_na * c->ad_nb), j = (w / c->ad_nb) % c->ad_na, k = w % c->ad_nb;
          diag = in == (i + j + k == q);
        }
      c->queries_diagonal = diag;
    }
    if (c->fast_kind == 2 && d->n_queries == c->ad_na + c->ad_nb - 1) {
      bool diag = true;
      for (int q = 0; q < d->n_queries && diag; ++q)
        for (int w = 0; w < d->n_worlds && diag; ++w) {
          const bool in = (qmask[(size_t)q * c->mask_words + (w >> 5)] >> (w & 31)) & 1u;
          diag = in == ((w / c->ad_nb) + (w % c->ad_nb) == q);
        }
      c->queries_diagonal = diag;
    }
  }
  if (!qmask_c.empty()) up(&c->d_qmask_c, qmask_c);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->d_bad_query), sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(c->d_bad_query, 0, sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&c->h_bad_query), sizeof(int32_t));
  if (e == cudaSuccess) c->gen = gen_program_create(d, &e);
  if (e != cudaSuccess) {
    vael_circuit_destroy(c);
    return cuda_fail(e, "circuit upload");
  }
  *out = c;
  return VAEL_OK;
}

extern "C" int vael_circuit_fast_kind(const vael_circuit_t* c) { return c ? c->fast_kind : 0; }
extern "C" uint64_t vael_circuit_hash(const vael_circuit_t* c) { return c ? c->hash : 0; }

// ---------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------
// VAEL_FLAG_CHECK_QUERY: count per-row query indices outside [-1, Q) (-1 = "no query for this row", value 0)
__global__ void check_query_idx_kernel(const int32_t* __restrict__ qidx, long long B, int32_t Q, int32_t* bad) {
  int32_t mine = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
    const int32_t q = __ldg(qidx + i);
    mine |= (q < -1 || q >= Q) ? 1 : 0;
  }
  if (__any_sync(0xffffffffu, mine) && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}

// Synchronises `st` (this is the one device entry-point option that is NOT CUDA-graph capturable).
static int check_query_indices(const vael_circuit_t* c, const int32_t* qidx, int64_t B, cudaStream_t st) {
  if (!qidx || B == 0) return VAEL_OK;
  std::lock_guard<std::mutex> lock(c->check_mutex);   // one flag word per circuit
  const int grid = (int)std::min<long long>((B + 255) / 256, (long long)c->num_sms * 8);
  check_query_idx_kernel<<<grid, 256, 0, st>>>(qidx, B, c->n_queries, c->d_bad_query);
  CK(cudaGetLastError());
  g_launches++;
  CK(cudaMemcpyAsync(c->h_bad_query, c->d_bad_query, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemsetAsync(c->d_bad_query, 0, sizeof(int32_t), st));
  CK(cudaStreamSynchronize(st));
  if (*c->h_bad_query) return fail(VAEL_E_INVALID, "per-row query index out of range (valid: -1 = none, 0..Q-1)");
  return VAEL_OK;
}

static vael_circuit_desc_t desc_of(const vael_circuit_t* c) {
  vael_circuit_desc_t d{};
  d.n_facts = c->n_facts; d.n_nodes = c->n_nodes; d.n_levels = c->n_levels; d.n_edges = c->n_edges;
  d.n_worlds = c->n_worlds; d.n_queries = c->n_queries; d.n_consts = c->n_consts; d.semiring = c->semiring;
  d.norm_node = c->norm_node;
  d.node_kind = c->h_node_kind.data(); d.node_arg = c->h_node_arg.data(); d.level_ptr = c->h_level_ptr.data();
  d.child_ptr = c->h_child_ptr.data(); d.child_idx = c->h_child_idx.data(); d.const_val = c->h_const_val.data();
  d.world_node = c->h_world_node.data(); d.query_node = c->h_query_node.data();
  return d;
}
// the circuit's specialised kernels, compiled on first use (thread-safe); null when they cannot be built
static const JitProgram* jit_of(const vael_circuit_t* c) {
  std::call_once(c->jit_once, [c] {
    const vael_circuit_desc_t d = desc_of(c);
    c->jit = jit_create(&d);
    if (!jit_ok(c->jit) && getenv("VAEL_JIT_VERBOSE")) fprintf(stderr, "vael_b200: no JIT kernels for this circuit: %s\n", jit_log(c->jit));
  });
  return jit_ok(c->jit) ? c->jit : nullptr;
}
extern "C" int vael_circuit_jit_status(const vael_circuit_t* c, char* log, int32_t n) {
  if (!c) return 0;
  const JitProgram* J = jit_of(c);
  if (log && n > 0) { strncpy(log, jit_log(c->jit), (size_t)n - 1); log[n - 1] = 0; }
  return J ? 1 : 0;
}
extern "C" int vael_jit_compile_check(const vael_circuit_desc_t* d, char* log, int32_t n) {
  if (!d) return 3;
  std::string l;
  size_t bytes = 0;
  const int rc = jit_compile_check(d, &l, &bytes);
  if (log && n > 0) { strncpy(log, l.c_str(), (size_t)n - 1); log[n - 1] = 0; }
  return rc;
}

static GenCall make_call(const vael_circuit_t* c, const float* facts, const int32_t* qidx, int32_t qs, int64_t B,
                         uint32_t flags) {
  GenCall G;
  memset(&G, 0, sizeof(G));
  G.facts = facts; G.qidx = qidx; G.qscalar = qs; G.B = B; G.flags = flags;
  G.qmask = c->d_qmask; G.mask_words = c->mask_words;
  return G;
}

static int check_common(const vael_circuit_t* c, const float* facts, const int32_t* qidx, int32_t qs, int64_t B,
                        uint32_t flags, bool need_query) {
  if (!c) return fail(VAEL_E_INVALID, "null circuit");
  if (B < 0) return fail(VAEL_E_INVALID, "negative batch size");
  if (B > 0 && !facts) return fail(VAEL_E_INVALID, "null facts");
  int dev = -1;
  CK(cudaGetDevice(&dev));
  if (dev != c->device) return fail(VAEL_E_INVALID, "circuit was created on another CUDA device");
  if ((flags & VAEL_FLAG_NORMALIZE) && c->norm_node < 0)
    return fail(VAEL_E_UNSUPPORTED, "NORMALIZE requested but the circuit has no normaliser node");
  if (need_query && !qidx && (qs < 0 || qs >= c->n_queries))
    return fail(VAEL_E_INVALID, "query index out of range");
  if ((flags & VAEL_FLAG_EVIDENCE) && !c->d_qmask)
    return fail(VAEL_E_UNSUPPORTED, "EVIDENCE needs query nodes that are sums over world nodes");
  return VAEL_OK;
}

// 0 = generic CSR kernel, 2 = ad2 (1-D bulk tiles), 3 = adw (2-D tensor-map tiles); `wide` is the [B,W] tensor
// of the call, whose base address the tensor map needs 16-byte aligned
// 5 = the circuit-specialised straight-line kernel (circuit_jit.cu)
static int pick_kernel(const vael_circuit_t* c, uint32_t flags, const float* wide) {
  if (flags & VAEL_FLAG_FORCE_GENERIC) return 0;
  if ((flags & VAEL_FLAG_FORCE_JIT) || c->fast_kind == 0)
    return jit_can(jit_of(c), flags, wide) ? 5 : ((flags & VAEL_FLAG_FORCE_JIT) ? -1 : 0);
  if (c->fast_kind == 4) return (flags & VAEL_FLAG_EVIDENCE) ? 0 : 4;   // no normaliser node: NORMALIZE is a no-op
  if ((flags & VAEL_FLAG_NORMALIZE) && !c->norm_fast_ok) return 0;
  if (c->fast_kind == 3 && !adw_usable(wide)) return 0;
  return c->fast_kind;
}

extern "C" int vael_circuit_forward(const vael_circuit_t* c, const float* facts, const int32_t* qidx, int32_t qs,
                                    int64_t B, float* worlds, float* query, uint32_t flags, void* stream) {
  const bool needq = query != nullptr || (flags & VAEL_FLAG_EVIDENCE);
  if (int rc = check_common(c, facts, qidx, qs, B, flags, needq)) return rc;
  if (B == 0) return VAEL_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((flags & VAEL_FLAG_CHECK_QUERY) && needq)
    if (int rc = check_query_indices(c, qidx, B, st)) return rc;
  const int kern = pick_kernel(c, flags, worlds);
  if (kern < 0) return fail(VAEL_E_UNSUPPORTED, std::string("no specialised kernel for this circuit: ") + jit_log(c->jit));
  if (kern == 5) {
    GenCall G = make_call(c, facts, qidx, needq ? qs : -1, B, flags);
    G.worlds = worlds; G.query = query;
    CK(jit_forward(jit_of(c), G, c->num_sms, st));
    g_last_kernel = "jit-straightline";
  } else if (kern == 2) {
    Ad2Params P{};
    P.facts = facts; P.qidx = qidx; P.qscalar = needq ? qs : -1; P.n_queries = c->n_queries; P.qmask = c->d_qmask;
    P.B = B; P.worlds = worlds; P.query = query;
    CK(ad2_forward(c->ad_na, c->ad_nb, P, flags & VAEL_FLAG_NORMALIZE, flags & VAEL_FLAG_EVIDENCE, c->num_sms, st));
    g_last_kernel = "ad2-tma";
  } else if (kern == 3) {
    AdwParams P{};
    P.facts = facts; P.qidx = qidx; P.qscalar = needq ? qs : -1; P.n_queries = c->n_queries; P.n_chunks = c->ad_np;
    P.qmask_c = c->d_qmask_c; P.B = B; P.worlds = worlds; P.query = query;
    CK(adw_forward(c->ad_np, c->ad_na, c->ad_nb, P, flags & VAEL_FLAG_NORMALIZE, flags & VAEL_FLAG_EVIDENCE, c->num_sms, st));
    g_last_kernel = "adw-tma2d";
  } else if (kern == 4 && worlds != nullptr && query == nullptr) {
    CK(lt_forward(c->n_facts, c->n_worlds, c->lt_mask, facts, B, worlds, c->num_sms, st));
    g_last_kernel = "lt-log";
  } else {
    GenCall G = make_call(c, facts, qidx, needq ? qs : -1, B, flags);
    G.worlds = worlds; G.query = query;
    if (!gen_fits(c->gen, false)) return fail(VAEL_E_UNSUPPORTED, "circuit too large for the shared-memory tile");
    CK(gen_forward(c->gen, G, c->num_sms, st));
    g_last_kernel = "generic-csr";
  }
  g_launches++;
  return VAEL_OK;
}

extern "C" int vael_circuit_queries_all(const vael_circuit_t* c, const float* facts, int64_t B, float* queries_all,
                                        uint32_t flags, void* stream) {
  if (int rc = check_common(c, facts, nullptr, 0, B, flags & ~VAEL_FLAG_EVIDENCE, false)) return rc;
  if (flags & VAEL_FLAG_EVIDENCE) return fail(VAEL_E_INVALID, "EVIDENCE is meaningless for queries_all");
  if (B == 0) return VAEL_OK;
  if (!queries_all) return fail(VAEL_E_INVALID, "null output");
  // Two-AD circuits whose queries are not the diagonals (Mario's overlapping, ragged member lists): the circuit's
  // generated kernel folds the same products in the same order as straight-line code — 0.70 of HBM against the table-
  // driven member-list kernel's 0.48 (scripts/stage_probe.py queries) — so it is preferred where it can be built
  // (first use compiles it; VAEL_QUERIES_MEMBERLIST=1, read per call, keeps the hand-written kernel: tests, A/B timing).
  if (pick_kernel(c, flags, nullptr) == 2 && !c->queries_diagonal && c->q_members_ok && c->n_queries <= kQMaxQueries) {
    const char* keep = getenv("VAEL_QUERIES_MEMBERLIST");
    if (!(keep && keep[0] == '1') && jit_can(jit_of(c), flags, nullptr)) {
      GenCall G = make_call(c, facts, nullptr, -1, B, flags);
      G.queries_all = queries_all;
      CK(jit_forward(jit_of(c), G, c->num_sms, static_cast<cudaStream_t>(stream)));
      g_last_kernel = "jit-straightline";
      g_launches++;
      return VAEL_OK;
    }
  }
  if (pick_kernel(c, flags, nullptr) == 2 && c->n_queries <= kQMaxQueries && (c->queries_diagonal || c->q_members_ok)) {
    Ad2QParams P{};
    P.facts = facts; P.B = B; P.n_queries = c->n_queries;
    P.diagonal = c->queries_diagonal ? 1 : 0;
    P.queries_all = queries_all;
    if (!c->queries_diagonal) {
      memcpy(P.ptr, c->q_ptr, sizeof(P.ptr));
      memcpy(P.len, c->q_len, sizeof(P.len));
      memcpy(P.slot, c->q_slot, sizeof(uint16_t) * c->q_nslots);
    }
    CK(ad2_queries_all(c->ad_na, c->ad_nb, P, flags & VAEL_FLAG_NORMALIZE, c->num_sms, static_cast<cudaStream_t>(stream)));
    g_last_kernel = "ad2-queries";
    g_launches++;
    return VAEL_OK;
  }
  if (pick_kernel(c, flags, nullptr) == 3 && c->queries_diagonal) {
    CK(adw_queries_diag(c->ad_np, c->ad_na, c->ad_nb, facts, B, queries_all, flags & VAEL_FLAG_NORMALIZE, c->num_sms,
                        static_cast<cudaStream_t>(stream)));
    g_last_kernel = "adw-queries";
    g_launches++;
    return VAEL_OK;
  }
  GenCall G = make_call(c, facts, nullptr, -1, B, flags);
  G.queries_all = queries_all;
  if (pick_kernel(c, flags, nullptr) == 5) {
    CK(jit_forward(jit_of(c), G, c->num_sms, static_cast<cudaStream_t>(stream)));
    g_last_kernel = "jit-straightline";
    g_launches++;
    return VAEL_OK;
  }
  if (!gen_fits(c->gen, false)) return fail(VAEL_E_UNSUPPORTED, "circuit too large for the shared-memory tile");
  CK(gen_forward(c->gen, G, c->num_sms, static_cast<cudaStream_t>(stream)));
  g_last_kernel = "generic-csr";
  g_launches++;
  return VAEL_OK;
}

extern "C" int vael_circuit_backward(const vael_circuit_t* c, const float* facts, const int32_t* qidx, int32_t qs,
                                     int64_t B, const float* grad_worlds, const float* grad_query,
                                     float* grad_facts, uint32_t flags, void* stream) {
  const bool needq = grad_query != nullptr;
  if (int rc = check_common(c, facts, qidx, qs, B, flags, needq)) return rc;
  if (flags & VAEL_FLAG_EVIDENCE) return fail(VAEL_E_UNSUPPORTED, "evidence mode has no adjoint (the reference only uses it under no_grad)");
  if (B == 0) return VAEL_OK;
  if (!grad_facts) return fail(VAEL_E_INVALID, "null grad_facts");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((flags & VAEL_FLAG_CHECK_QUERY) && needq)
    if (int rc = check_query_indices(c, qidx, B, st)) return rc;
  const int kern = pick_kernel(c, flags, grad_worlds);
  if (kern < 0) return fail(VAEL_E_UNSUPPORTED, std::string("no specialised kernel for this circuit: ") + jit_log(c->jit));
  if (kern == 5) {
    GenCall G = make_call(c, facts, qidx, needq ? qs : -1, B, flags);
    G.grad_worlds = grad_worlds; G.grad_query = grad_query; G.grad_facts = grad_facts;
    CK(jit_backward(jit_of(c), G, c->num_sms, st));
    g_last_kernel = "jit-straightline";
  } else if (kern == 2) {
    Ad2Params P{};
    P.facts = facts; P.qidx = qidx; P.qscalar = needq ? qs : -1; P.n_queries = c->n_queries; P.qmask = c->d_qmask;
    P.B = B; P.grad_worlds = grad_worlds; P.grad_query = grad_query; P.grad_facts = grad_facts;
    CK(ad2_backward(c->ad_na, c->ad_nb, P, flags & VAEL_FLAG_NORMALIZE, c->num_sms, st));
    g_last_kernel = "ad2-tma";
  } else if (kern == 3) {
    AdwParams P{};
    P.facts = facts; P.qidx = qidx; P.qscalar = needq ? qs : -1; P.n_queries = c->n_queries; P.n_chunks = c->ad_np;
    P.qmask_c = c->d_qmask_c; P.B = B; P.grad_worlds = grad_worlds; P.grad_query = grad_query; P.grad_facts = grad_facts;
    CK(adw_backward(c->ad_np, c->ad_na, c->ad_nb, P, flags & VAEL_FLAG_NORMALIZE, c->num_sms, st));
    g_last_kernel = "adw-tma2d";
  } else if (kern == 4) {   // no query nodes: grad_query has nothing to seed
    CK(lt_backward(c->n_facts, c->n_worlds, c->lt_mask, facts, grad_worlds, B, grad_facts, c->num_sms, st));
    g_last_kernel = "lt-log";
  } else {
    GenCall G = make_call(c, facts, qidx, needq ? qs : -1, B, flags);
    G.grad_worlds = grad_worlds; G.grad_query = grad_query; G.grad_facts = grad_facts;
    if (!gen_fits(c->gen, true)) return fail(VAEL_E_UNSUPPORTED, "circuit too large for the shared-memory tile");
    CK(gen_backward(c->gen, G, c->num_sms, st));
    g_last_kernel = "generic-csr";
  }
  g_launches++;
  return VAEL_OK;
}

// ---------------------------------------------------------------------------
// host-buffer variants: row chunks round-robin over kHostSlots slots (each with its own device buffers, a
// compute/download stream and an upload stream) so that the H2D copy of one chunk, the kernels of another and the
// D2H copy of a third overlap and both PCIe directions stay busy.  Uploads have their own stream per slot: in one
// stream the upload of a chunk's upstream gradients would queue behind the download of its worlds, which it does not
// depend on, and the upload direction idled part of the call (0.90 of the link; 0.93 now).  Chunks are LARGE
// (256 Ki rows by default, VAEL_HOST_CHUNK_ROWS): every chunk is seven copies, three of them 4 bytes per row, and the
// per-copy cost shows — 16 Ki-row chunks reach 0.71 of the link, 32 Ki 0.83, 128 Ki 0.928, 256 Ki 0.933
// (scripts/e2e_probe.py).  A call smaller than four full chunks is split in four so that all slots work.  A ramped
// schedule (32 / 64 / 128 Ki-row chunks at both ends, to shorten the single-direction head and tail) measured WORSE:
// 11.25 ms against 11.02 ms per Mi rows.
// ---------------------------------------------------------------------------
// Caller holds host_mutex.  All-or-nothing: a failure frees what this call created.
static int ensure_slots(const vael_circuit_t* c) {
  if (c->chunk_rows) return VAEL_OK;
  const char* env = getenv("VAEL_HOST_CHUNK_ROWS");
  long long rows = env && *env ? atoll(env) : (1 << 18);
  if (rows < 32) rows = 32;
  const size_t F = c->n_facts, W = c->n_worlds ? c->n_worlds : 1;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](auto** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(p), bytes); };
  auto event = [&](cudaEvent_t* ev) { if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming); };
  for (int s = 0; s < kHostSlots && e == cudaSuccess; ++s) {
    auto& S = c->slot[s];
    e = cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&S.up, cudaStreamNonBlocking);
    event(&S.facts_up); event(&S.grads_up); event(&S.done);
    alloc(&S.facts, rows * F * 4); alloc(&S.gf, rows * F * 4);
    alloc(&S.worlds, rows * W * 4); alloc(&S.gw, rows * W * 4);
    alloc(&S.query, rows * 4); alloc(&S.gq, rows * 4); alloc(&S.qidx, rows * 4);
  }
  if (e != cudaSuccess) {
    for (auto& S : c->slot) {
      cudaFree(S.facts); cudaFree(S.worlds); cudaFree(S.query); cudaFree(S.gw); cudaFree(S.gq); cudaFree(S.gf);
      cudaFree(S.qidx);
      destroy_slot_streams(S);
      S = vael_host_slot{};
    }
    return cuda_fail(e, "host pipeline allocation");
  }
  c->chunk_rows = rows;
  return VAEL_OK;
}

// worlds_h / query_h / gf_h select what is computed and copied back; `backward` runs the adjoint with the
// upstream gradients gw_h / gq_h (either may be null = 0; both null with `ones_gq` = d query / d facts).
static int host_pipeline_locked(const vael_circuit_t* c, const float* facts_h, const int32_t* qidx_h, int32_t qs,
                                int64_t B, const float* gw_h, const float* gq_h, float* worlds_h, float* query_h,
                                float* gf_h, uint32_t flags, bool backward) {
  const size_t F = c->n_facts, W = c->n_worlds;
  static const bool fixed_chunks = [] { const char* v = getenv("VAEL_HOST_FIXED_CHUNKS"); return v && v[0] == '1'; }();
  int64_t chunk = c->chunk_rows;
  if (!fixed_chunks) {
    const int64_t quarter = ((B / kHostSlots + 31) / 32) * 32;
    chunk = std::min<int64_t>(chunk, std::max<int64_t>(quarter, 32768));
  }
  int idx = 0;
  for (int64_t r0 = 0; r0 < B; r0 += chunk, ++idx) {
    const int64_t n = (B - r0 < chunk) ? (B - r0) : chunk;
    auto& S = c->slot[idx % kHostSlots];
    // uploads: after the chunk that used this slot before has finished with its buffers
    if (idx >= kHostSlots) CK(cudaStreamWaitEvent(S.up, S.done, 0));
    CK(cudaMemcpyAsync(S.facts, facts_h + r0 * F, n * F * 4, cudaMemcpyHostToDevice, S.up));
    if (qidx_h) CK(cudaMemcpyAsync(S.qidx, qidx_h + r0, n * 4, cudaMemcpyHostToDevice, S.up));
    CK(cudaEventRecord(S.facts_up, S.up));
    if (backward) {
      if (gw_h) CK(cudaMemcpyAsync(S.gw, gw_h + r0 * W, n * W * 4, cudaMemcpyHostToDevice, S.up));
      if (gq_h) CK(cudaMemcpyAsync(S.gq, gq_h + r0, n * 4, cudaMemcpyHostToDevice, S.up));
      CK(cudaEventRecord(S.grads_up, S.up));
    }
    CK(cudaStreamWaitEvent(S.stream, S.facts_up, 0));
    if (worlds_h || query_h) {
      if (int rc = vael_circuit_forward(c, S.facts, qidx_h ? S.qidx : nullptr, qs, n, worlds_h ? S.worlds : nullptr,
                                        query_h ? S.query : nullptr, flags, S.stream))
        return rc;
      if (worlds_h) CK(cudaMemcpyAsync(worlds_h + r0 * W, S.worlds, n * W * 4, cudaMemcpyDeviceToHost, S.stream));
      if (query_h) CK(cudaMemcpyAsync(query_h + r0, S.query, n * 4, cudaMemcpyDeviceToHost, S.stream));
    }
    if (backward) {
      CK(cudaStreamWaitEvent(S.stream, S.grads_up, 0));
      if (int rc = vael_circuit_backward(c, S.facts, qidx_h ? S.qidx : nullptr, qs, n, gw_h ? S.gw : nullptr,
                                         gq_h ? S.gq : nullptr, S.gf, flags & ~VAEL_FLAG_CHECK_QUERY, S.stream))
        return rc;
      CK(cudaMemcpyAsync(gf_h + r0 * F, S.gf, n * F * 4, cudaMemcpyDeviceToHost, S.stream));
    }
    CK(cudaEventRecord(S.done, S.stream));
  }
  return VAEL_OK;
}

static int host_pipeline(const vael_circuit_t* c, const float* facts_h, const int32_t* qidx_h, int32_t qs, int64_t B,
                         const float* gw_h, const float* gq_h, float* worlds_h, float* query_h, float* gf_h,
                         uint32_t flags, bool backward) {
  if (!c) return fail(VAEL_E_INVALID, "null circuit");
  if (B < 0 || (B > 0 && !facts_h)) return fail(VAEL_E_INVALID, "bad host buffers");
  if (backward && B > 0 && !gf_h) return fail(VAEL_E_INVALID, "null grad_facts host buffer");
  int dev = -1;
  CK(cudaGetDevice(&dev));
  if (dev != c->device) return fail(VAEL_E_INVALID, "circuit was created on another CUDA device");
  // the slot buffers and streams belong to the circuit handle: concurrent host callers of one handle take turns
  std::lock_guard<std::mutex> lock(c->host_mutex);
  if (int rc = ensure_slots(c)) return rc;
  const int rc = host_pipeline_locked(c, facts_h, qidx_h, qs, B, gw_h, gq_h, worlds_h, query_h, gf_h, flags, backward);
  // success or not, no copy may still be in flight when the caller gets its host buffers back
  cudaError_t first = cudaSuccess;
  for (int s = 0; s < kHostSlots; ++s) {
    for (cudaStream_t st : {c->slot[s].up, c->slot[s].stream}) {
      const cudaError_t e = cudaStreamSynchronize(st);
      if (first == cudaSuccess) first = e;
    }
  }
  if (rc) return rc;
  if (first != cudaSuccess) return cuda_fail(first, "cudaStreamSynchronize");
  return VAEL_OK;
}

extern "C" int vael_circuit_forward_host(const vael_circuit_t* c, const float* facts_h, const int32_t* qidx_h,
                                         int32_t qs, int64_t B, float* worlds_h, float* query_h, uint32_t flags) {
  return host_pipeline(c, facts_h, qidx_h, qs, B, nullptr, nullptr, worlds_h, query_h, nullptr, flags, false);
}
extern "C" int vael_circuit_fwdbwd_host(const vael_circuit_t* c, const float* facts_h, const int32_t* qidx_h,
                                        int32_t qs, int64_t B, const float* gw_h, const float* gq_h, float* worlds_h,
                                        float* query_h, float* gf_h, uint32_t flags) {
  return host_pipeline(c, facts_h, qidx_h, qs, B, gw_h, gq_h, worlds_h, query_h, gf_h, flags, true);
}
extern "C" int vael_circuit_query_grad_host(const vael_circuit_t* c, const float* facts_h, const int32_t* qidx_h,
                                            int32_t qs, int64_t B, const float* gq_h, float* query_h, float* gf_h,
                                            uint32_t flags) {
  if (B > 0 && (!gq_h || !query_h)) return fail(VAEL_E_INVALID, "null grad_query / query host buffer");
  return host_pipeline(c, facts_h, qidx_h, qs, B, nullptr, gq_h, nullptr, query_h, gf_h, flags, true);
}

// ---------------------------------------------------------------------------
// facts head, Gumbel sampling, ELBO terms
// ---------------------------------------------------------------------------
extern "C" int vael_facts_forward(const float* logits, int64_t B, int32_t n_groups, int32_t group, float eps,
                                  float* facts, void* stream) {
  if (B < 0 || n_groups <= 0 || group <= 0) return fail(VAEL_E_INVALID, "bad shape");
  if (B == 0) return VAEL_OK;
  if (!logits || !facts) return fail(VAEL_E_INVALID, "null pointer");
  int sms = 0;
  if (int rc = device_sms(&sms)) return rc;
  CK(facts_forward(logits, B, n_groups, group, eps, 0.0f, facts, sms, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return VAEL_OK;
}
extern "C" int vael_facts_backward(const float* logits, const float* grad_facts, int64_t B, int32_t n_groups,
                                   int32_t group, float eps, float* grad_logits, void* stream) {
  if (B < 0 || n_groups <= 0 || group <= 0) return fail(VAEL_E_INVALID, "bad shape");
  if (B == 0) return VAEL_OK;
  if (!logits || !grad_facts || !grad_logits) return fail(VAEL_E_INVALID, "null pointer");
  int sms = 0;
  if (int rc = device_sms(&sms)) return rc;
  CK(facts_backward(logits, grad_facts, B, n_groups, group, eps, 0.0f, grad_logits, sms, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return VAEL_OK;
}
extern "C" int vael_facts_clip_forward(const float* logits, int64_t B, int32_t n_groups, int32_t group, float lo,
                                       float hi, float* facts, void* stream) {
  if (B < 0 || n_groups <= 0 || group <= 0 || !(hi > 0.0f) || !(lo <= hi)) return fail(VAEL_E_INVALID, "bad shape or clip range");
  if (B == 0) return VAEL_OK;
  if (!logits || !facts) return fail(VAEL_E_INVALID, "null pointer");
  int sms = 0;
  if (int rc = device_sms(&sms)) return rc;
  CK(facts_forward(logits, B, n_groups, group, lo, hi, facts, sms, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return VAEL_OK;
}
extern "C" int vael_facts_clip_backward(const float* logits, const float* grad_facts, int64_t B, int32_t n_groups,
                                        int32_t group, float lo, float hi, float* grad_logits, void* stream) {
  if (B < 0 || n_groups <= 0 || group <= 0 || !(hi > 0.0f) || !(lo <= hi)) return fail(VAEL_E_INVALID, "bad shape or clip range");
  if (B == 0) return VAEL_OK;
  if (!logits || !grad_facts || !grad_logits) return fail(VAEL_E_INVALID, "null pointer");
  int sms = 0;
  if (int rc = device_sms(&sms)) return rc;
  CK(facts_backward(logits, grad_facts, B, n_groups, group, lo, hi, grad_logits, sms, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return VAEL_OK;
}

extern "C" int vael_gumbel_world_forward(const float* scores, const float* noise, uint64_t seed, int64_t B, int32_t W,
                                         int32_t is_log, float tau, const float* herbrand, int32_t H, float* y_soft,
                                         int32_t* world_idx, float* world_h, void* stream) {
  if (B < 0 || W <= 0 || W > 1024 || H < 0 || !(tau > 0.0f)) return fail(VAEL_E_INVALID, "bad shape or tau");
  if (B == 0) return VAEL_OK;
  if (!scores || (world_h && !herbrand)) return fail(VAEL_E_INVALID, "null pointer");
  int sms = 0;
  if (int rc = device_sms(&sms)) return rc;
  GumbelParams P{};
  P.scores = scores; P.noise = noise; P.seed = seed; P.B = B; P.W = W; P.H = H; P.is_log = is_log; P.tau = tau;
  P.herbrand = herbrand; P.y_soft = y_soft; P.world_idx = world_idx; P.world_h = world_h;
  CK(gumbel_forward(P, sms, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return VAEL_OK;
}
extern "C" int vael_gumbel_world_backward(const float* scores, const float* y_soft, const float* grad_world_h,
                                          int64_t B, int32_t W, int32_t is_log, float tau, const float* herbrand,
                                          int32_t H, float* grad_scores, void* stream) {
  if (B < 0 || W <= 0 || W > 1024 || H <= 0 || !(tau > 0.0f)) return fail(VAEL_E_INVALID, "bad shape or tau");
  if (B == 0) return VAEL_OK;
  if (!y_soft || !grad_world_h || !herbrand || !grad_scores || (!is_log && !scores)) return fail(VAEL_E_INVALID, "null pointer");
  int sms = 0;
  if (int rc = device_sms(&sms)) return rc;
  GumbelParams P{};
  P.scores = scores; P.B = B; P.W = W; P.H = H; P.is_log = is_log; P.tau = tau; P.herbrand = herbrand;
  P.y_soft = const_cast<float*>(y_soft); P.grad_world_h = grad_world_h; P.grad_scores = grad_scores;
  CK(gumbel_backward(P, sms, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return VAEL_OK;
}

// ---------------------------------------------------------------------------
// fused training-mode logic path (logic_fused.cu)
// ---------------------------------------------------------------------------
static int fused_check(const vael_circuit_t* c, const float* logits, const int32_t* qidx, int32_t qs, int64_t B,
                       float tau, bool need_query) {
  if (!c) return fail(VAEL_E_INVALID, "null circuit");
  if (B < 0) return fail(VAEL_E_INVALID, "negative batch size");
  if (B > 0 && !logits) return fail(VAEL_E_INVALID, "null logits");
  if (!(tau > 0.0f)) return fail(VAEL_E_INVALID, "tau must be positive");
  int dev = -1;
  CK(cudaGetDevice(&dev));
  if (dev != c->device) return fail(VAEL_E_INVALID, "circuit was created on another CUDA device");
  if (c->fast_kind != 2 || !fused_supported(c->ad_na, c->ad_nb) || !c->d_qmask)
    return fail(VAEL_E_UNSUPPORTED, "the fused logic kernel covers two equal annotated disjunctions of 9 or 10 values");
  if (need_query && !qidx && (qs < 0 || qs >= c->n_queries)) return fail(VAEL_E_INVALID, "query index out of range");
  return VAEL_OK;
}

extern "C" int vael_logic_fused_supported(const vael_circuit_t* c) {
  return c && c->fast_kind == 2 && fused_supported(c->ad_na, c->ad_nb) && c->d_qmask ? 1 : 0;
}

extern "C" int vael_logic_train_forward(const vael_circuit_t* c, const float* logits, const int32_t* qidx, int32_t qs,
                                        int64_t B, float eps, float tau, const float* noise, uint64_t seed,
                                        const uint64_t* seed_dev, float* facts, float* query, int32_t* world_idx,
                                        float* world_h, float* y_soft, void* stream) {
  if (int rc = fused_check(c, logits, qidx, qs, B, tau, query != nullptr)) return rc;
  if (B == 0) return VAEL_OK;
  FusedParams P{};
  P.logits = logits; P.qidx = qidx; P.qscalar = query ? qs : -1; P.n_queries = c->n_queries; P.qmask = c->d_qmask;
  P.B = B; P.eps = eps; P.tau = tau; P.noise = noise; P.seed = seed; P.seed_dev = seed_dev;
  P.facts = facts; P.query = query; P.world_idx = world_idx; P.world_h = world_h; P.y_soft = y_soft;
  CK(fused_forward(c->ad_na, P, c->num_sms, static_cast<cudaStream_t>(stream)));
  g_last_kernel = "logic-fused";
  g_launches++;
  return VAEL_OK;
}

extern "C" int vael_logic_train_backward(const vael_circuit_t* c, const float* logits, const int32_t* qidx, int32_t qs,
                                         int64_t B, float eps, float tau, const float* noise, uint64_t seed,
                                         const uint64_t* seed_dev, const float* grad_world_h, const float* grad_query,
                                         const float* grad_facts, float* grad_logits, void* stream) {
  if (int rc = fused_check(c, logits, qidx, qs, B, tau, grad_query != nullptr)) return rc;
  if (B == 0) return VAEL_OK;
  if (!grad_logits) return fail(VAEL_E_INVALID, "null grad_logits");
  FusedParams P{};
  P.logits = logits; P.qidx = qidx; P.qscalar = grad_query ? qs : -1; P.n_queries = c->n_queries; P.qmask = c->d_qmask;
  P.B = B; P.eps = eps; P.tau = tau; P.noise = noise; P.seed = seed; P.seed_dev = seed_dev;
  P.grad_world_h = grad_world_h; P.grad_query = grad_query; P.grad_facts = grad_facts; P.grad_logits = grad_logits;
  CK(fused_backward(c->ad_na, P, c->num_sms, static_cast<cudaStream_t>(stream)));
  g_last_kernel = "logic-fused";
  g_launches++;
  return VAEL_OK;
}

extern "C" int64_t vael_elbo_workspace_bytes(void) { return (int64_t)elbo_workspace_bytes(); }

extern "C" int vael_elbo_terms(const float* recon, const float* x, int64_t B, int64_t D, const float* mu,
                               const float* logvar, int32_t L, const float* query_prob, float recon_w, float kl_w,
                               float query_w, int32_t rec_loss, float* terms, float* grad_recon, float* grad_mu,
                               float* grad_logvar, float* grad_query, void* workspace, void* stream) {
  if (B <= 0 || D <= 0 || L < 0) return fail(VAEL_E_INVALID, "bad shape");
  if (rec_loss < VAEL_REC_LAPLACE || rec_loss > VAEL_REC_BCE) return fail(VAEL_E_INVALID, "unknown rec_loss");
  if (!recon || !x || !terms || !workspace || (L > 0 && (!mu || !logvar))) return fail(VAEL_E_INVALID, "null pointer");
  int sms = 0;
  if (int rc = device_sms(&sms)) return rc;
  ElboParams P{};
  P.recon = recon; P.x = x; P.B = B; P.D = D; P.mu = mu; P.logvar = logvar; P.L = L; P.query_prob = query_prob;
  P.recon_w = recon_w; P.kl_w = kl_w; P.query_w = query_w; P.rec_loss = rec_loss; P.terms = terms; P.grad_recon = grad_recon;
  P.grad_mu = grad_mu; P.grad_logvar = grad_logvar; P.grad_query = grad_query;
  CK(elbo_terms(P, workspace, sms, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return VAEL_OK;
}
