/*
 * vael_b200.h — C-ABI of the B200-native VAEL probabilistic-logic hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.
 * Every entry point names the reference interface (EleMisi/VAEL, file:line)
 * whose work it replaces.  The reference is pure Python/PyTorch, so the
 * "FFI" a maintainer binds is ctypes; the host-side
 * mirror of the reference classes lives in the vael_b200 python package.
 *
 * Conventions
 *  - all tensors are fp32, row-major ("batch-major": one batch row is
 *    contiguous), int32 for indices; B is the number of batch rows.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Device-pointer entry points only enqueue work: no allocation, no host
 *    synchronisation, CUDA-graph capturable.
 *  - *_host entry points take HOST buffers, do the H2D / D2H copies
 *    themselves in row chunks on internal streams, and return after the
 *    results are in the host buffers (also on an error return: no copy is
 *    still in flight when the caller gets its buffers back).
 *  - threading: a circuit handle is immutable after creation as far as the
 *    device entry points are concerned — any number of host threads may call
 *    them on one handle, each on its own stream.  The *_host entry points
 *    share the handle's staging buffers and streams (allocated on first use):
 *    concurrent *_host callers of ONE handle are serialised by a mutex inside
 *    the handle; use one handle per thread for parallel host pipelines.
 *  - return value: 0 = ok, otherwise a VAEL_E_* code; vael_last_error() gives
 *    the message for the calling thread.
 *  - there is no CPU fallback: without a CUDA device every compute entry
 *    point returns VAEL_E_CUDA.
 */
#ifndef VAEL_B200_H
#define VAEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAEL_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define VAEL_API __attribute__((visibility("default")))
#else
#define VAEL_API
#endif

/* ---- error codes ------------------------------------------------------- */
#define VAEL_OK 0
#define VAEL_E_INVALID 1     /* bad argument (shape, index, null pointer)   */
#define VAEL_E_CUDA 2        /* CUDA runtime error / no device              */
#define VAEL_E_UNSUPPORTED 3 /* valid request the compiled circuit cannot do */
#define VAEL_E_NOMEM 4

/* ---- node kinds of the level-ordered arithmetic circuit ---------------- */
#define VAEL_NODE_LEAF_POS 0 /* value = facts[b, arg]            (weight of a positive literal) */
#define VAEL_NODE_LEAF_NEG 1 /* value = 1 - facts[b, arg]        (GraphSemiring.negate, utils/graph_semiring.py:13-15) */
#define VAEL_NODE_CONST 2    /* value = const_val[arg]           (GraphSemiring.one/zero, :17-25) */
#define VAEL_NODE_PROD 3     /* semiring times over children     (GraphSemiring.times, :43-49) */
#define VAEL_NODE_SUM 4      /* semiring plus over children      (GraphSemiring.plus, :35-41) */
#define VAEL_NODE_COMPL 5    /* 1 - sum(children): ProbLog's AD null-choice weight (ad_complement) */

/* ---- semirings ---------------------------------------------------------- */
#define VAEL_SEMIRING_PROB 0 /* (+, *) on probabilities — GraphSemiring */
#define VAEL_SEMIRING_LOG 1  /* (logsumexp, +) on log-probabilities — MarioVAELModel.admissible_worlds_logprob, models/vael.py:368-375 */

/* ---- evaluation flags ---------------------------------------------------- */
#define VAEL_FLAG_NORMALIZE 1u /* divide every output by the normaliser node Z (GraphSemiring.normalize, :54-55) */
#define VAEL_FLAG_EVIDENCE 2u  /* condition the worlds on the selected query node: P(w | q) (models/vael.py:119-135) */
#define VAEL_FLAG_FORCE_GENERIC 4u /* run the generic level-by-level CSR kernel even if a structured fast path exists */
#define VAEL_FLAG_FORCE_JIT 16u    /* run the circuit-specialised straight-line kernel (compiled with NVRTC on first use)
                                      even if a structured fast path exists; VAEL_E_UNSUPPORTED when it cannot be built */
#define VAEL_FLAG_CHECK_QUERY 8u   /* validate the per-row query indices first: any index outside [-1, Q) makes the
                                      call return VAEL_E_INVALID before the circuit kernel is launched (-1 is the
                                      "no query for this row" sentinel: its query value is 0).  The check
                                      synchronises `stream`, so a call carrying this flag is NOT graph-capturable.
                                      Without the flag an out-of-range index silently yields 0. */

/* ---- reconstruction terms of vael_elbo_terms (rec_loss of loss_function, utils/mnist_utils/train.py:16-21) ---- */
#define VAEL_REC_LAPLACE 0 /* -mean_b sum_pix log Laplace(x | recon, 1)   (config.py:17,45: the shipped setting) */
#define VAEL_REC_MSE 1     /* torch.nn.MSELoss(reduction='mean') over the flattened images */
#define VAEL_REC_BCE 2     /* torch.nn.BCELoss(reduction='mean') over the flattened images */

/*
 * Host description of a compiled circuit.  Nodes are numbered level by level
 * (level 0 = leaves and constants); a node's children all live in strictly
 * lower levels.  Produced by vael_b200.compiler (the replacement of
 * build_model_dict, utils/mnist_utils/mnist_task_VAEL.py:203-219, which calls
 * LogicFormula -> LogicDAG -> SDD.create_from in ProbLog/PySDD).
 */
typedef struct vael_circuit_desc {
  int32_t n_facts;           /* F: columns of the facts matrix                        */
  int32_t n_nodes;
  int32_t n_levels;
  int32_t n_edges;
  int32_t n_worlds;          /* W: world outputs, reference order (d1*10+d2, models/vael.py:218-221,255-258) */
  int32_t n_queries;         /* Q: candidate query nodes, selected per row by query_idx */
  int32_t n_consts;
  int32_t semiring;          /* VAEL_SEMIRING_*                                       */
  int32_t norm_node;         /* node id of Z, or -1                                   */
  const uint8_t* node_kind;  /* [n_nodes]   VAEL_NODE_*                               */
  const int32_t* node_arg;   /* [n_nodes]   leaf: fact column; const: index in const_val; else 0 */
  const int32_t* level_ptr;  /* [n_levels+1] node id ranges per level                 */
  const int32_t* child_ptr;  /* [n_nodes+1] CSR row pointers                          */
  const int32_t* child_idx;  /* [n_edges]   CSR column indices (node ids)             */
  const float* const_val;    /* [n_consts]                                            */
  const int32_t* world_node; /* [n_worlds]                                            */
  const int32_t* query_node; /* [n_queries]                                           */
} vael_circuit_desc_t;

typedef struct vael_circuit vael_circuit_t; /* opaque; owns the HBM copy of the arrays */

/* ---- library ------------------------------------------------------------ */
VAEL_API int vael_abi_version(void);
VAEL_API const char* vael_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
VAEL_API int64_t vael_launch_count(void);
/* name of the kernel family the last forward/backward used: "ad2-tma", "adw-tma2d", "lt-log", "jit-straightline",
 * "generic-csr", "logic-fused" */
VAEL_API const char* vael_last_kernel(void);

/* ---- circuit lifetime ---------------------------------------------------- */
/* Validates the description (level order, index ranges), uploads it to the
 * current CUDA device and derives the backward (parent) tables and, when the
 * CSR has the annotated-disjunction outer-product shape, the structured
 * fast-path tables.  Replaces keeping 38 SDD objects in model_dict
 * (utils/mnist_utils/mnist_task_VAEL.py:208-217).                            */
VAEL_API int vael_circuit_create(const vael_circuit_desc_t* desc, vael_circuit_t** out);
VAEL_API void vael_circuit_destroy(vael_circuit_t* c);
/* 0 = generic CSR only; 2 / 3 = structured path for worlds that are the outer product of two / three
 * annotated disjunctions (1-D bulk-copy tiles / 2-D tensor-map tiles); 4 = literal table in the log
 * semiring (every world one literal per fact column: MarioVAELModel.admissible_worlds_logprob,
 * models/vael.py:368-375) */
VAEL_API int vael_circuit_fast_kind(const vael_circuit_t* c);
/* Circuits without a recognised shape run, when possible, a kernel SPECIALISED to the circuit: the level-ordered CSR
 * unrolled into straight-line register code (lane <-> row), generated once per circuit, compiled with NVRTC for
 * sm_100a on first use and loaded through the driver API; the level-by-level CSR interpreter is the fallback (no
 * libnvrtc, VAEL_JIT=0, log semiring, more than 128 worlds) and stays reachable with VAEL_FLAG_FORCE_GENERIC.
 * vael_circuit_jit_status compiles now if that has not happened yet and returns 1 when the kernels exist (the NVRTC
 * log / the reason otherwise is copied to `log`, up to n bytes).  vael_jit_compile_check generates and compiles
 * the source for a description WITHOUT loading it (works without a GPU): 0 = compiled, 1 = circuit outside the
 * generator's scope, 2 = NVRTC unavailable, 3 = compile error. */
VAEL_API int vael_circuit_jit_status(const vael_circuit_t* c, char* log, int32_t n);
VAEL_API int vael_jit_compile_check(const vael_circuit_desc_t* desc, char* log, int32_t n);
/* 64-bit FNV-1a hash of the uploaded arrays (bit-exact circuit identity) */
VAEL_API uint64_t vael_circuit_hash(const vael_circuit_t* c);

/* ---- circuit forward ------------------------------------------------------
 * Replaces sdd.evaluate(semiring=GraphSemiring) + extract_* (models/vael.py:98-135,
 * :245-266) and the dense override MNISTPairsVAELModel.problog_inference
 * (models/vael.py:200-225); for Mario compute_worlds_prob_conj / compute_query_prob /
 * admissible_worlds_logprob (models/vael.py:359-399).
 *   facts      [B,F]  device, contiguous
 *   query_idx  [B]    device int32 per-row query selector, or NULL -> query_scalar for every row
 *   worlds     [B,W]  device out (may be NULL to skip)
 *   query      [B]    device out (may be NULL to skip)                        */
VAEL_API int vael_circuit_forward(const vael_circuit_t* c, const float* facts, const int32_t* query_idx,
                         int32_t query_scalar, int64_t B, float* worlds, float* query, uint32_t flags,
                         void* stream);

/* ---- circuit backward (adjoint) ------------------------------------------
 * Replaces autograd through the recorded semiring ops (loss.backward(),
 * utils/mnist_utils/train.py:147).  Node values are recomputed from facts; nothing
 * is saved by forward.
 *   grad_worlds [B,W] device or NULL (=0);  grad_query [B] device or NULL (=0)
 *   grad_facts  [B,F] device out (overwritten)                                */
VAEL_API int vael_circuit_backward(const vael_circuit_t* c, const float* facts, const int32_t* query_idx,
                          int32_t query_scalar, int64_t B, const float* grad_worlds,
                          const float* grad_query, float* grad_facts, uint32_t flags, void* stream);

/* ---- all queries at once ---------------------------------------------------
 * queries_all[b,q] = value of query node q: the worlds[B,W] @ w_q[W,Q] mask
 * contraction of discriminative_ability (utils/mnist_utils/metrics.py:71-75),
 * compute_query (models/vael.py:200-206).                                     */
VAEL_API int vael_circuit_queries_all(const vael_circuit_t* c, const float* facts, int64_t B, float* queries_all,
                             uint32_t flags, void* stream);

/* ---- host-buffer variants (the e2e path: copies inside) ------------------- */
VAEL_API int vael_circuit_forward_host(const vael_circuit_t* c, const float* facts_h, const int32_t* query_idx_h,
                              int32_t query_scalar, int64_t B, float* worlds_h, float* query_h,
                              uint32_t flags);
/* forward + backward in one pass over host data: worlds/query out, grad_facts out */
VAEL_API int vael_circuit_fwdbwd_host(const vael_circuit_t* c, const float* facts_h, const int32_t* query_idx_h,
                             int32_t query_scalar, int64_t B, const float* grad_worlds_h,
                             const float* grad_query_h, float* worlds_h, float* query_h,
                             float* grad_facts_h, uint32_t flags);
/* What the reference's training loop consumes from the logic path (utils/mnist_utils/train.py:123-147: the
 * query probability enters the BCE term, its gradient flows back to the facts) without moving any [B,W] tensor
 * over PCIe: facts [B,F] + query index (+ upstream grad_query [B]) in; query [B] + grad_facts [B,F] =
 * grad_query * d query / d facts out.  172 B per row against 1000 B for vael_circuit_fwdbwd_host. */
VAEL_API int vael_circuit_query_grad_host(const vael_circuit_t* c, const float* facts_h, const int32_t* query_idx_h,
                                 int32_t query_scalar, int64_t B, const float* grad_query_h, float* query_h,
                                 float* grad_facts_h, uint32_t flags);

/* ---- facts from MLP logits --------------------------------------------------
 * MNISTPairsVAELModel.compute_facts_probability (models/vael.py:159-177):
 * per group of `group` logits: softmax, +eps, divide by the (detached) sum.
 *   logits [B, n_groups*group] -> facts same shape.                            */
VAEL_API int vael_facts_forward(const float* logits, int64_t B, int32_t n_groups, int32_t group, float eps,
                       float* facts, void* stream);
/* logits are re-read (the softmax is recomputed); grad_facts [B,n_groups*group] -> grad_logits */
VAEL_API int vael_facts_backward(const float* logits, const float* grad_facts, int64_t B, int32_t n_groups,
                        int32_t group, float eps, float* grad_logits, void* stream);

/* MarioVAELModel's facts (models/vael.py:319-322): per group softmax, then clamp to [lo, hi]
 * (torch.clip(min=1e-5, max=0.9999)); the gradient passes where lo <= softmax <= hi.          */
VAEL_API int vael_facts_clip_forward(const float* logits, int64_t B, int32_t n_groups, int32_t group, float lo,
                            float hi, float* facts, void* stream);
VAEL_API int vael_facts_clip_backward(const float* logits, const float* grad_facts, int64_t B, int32_t n_groups,
                             int32_t group, float lo, float hi, float* grad_logits, void* stream);

/* ---- Gumbel-softmax world sampling + Herbrand projection --------------------
 * Replaces logits=log(worlds_prob); F.gumbel_softmax(logits, tau, hard=True);
 * world @ herbrand_base (models/vael.py:61-65, :82-84; Mario :335-338).
 *   scores   [B,W]  world probabilities (is_log=0) or log-probabilities (is_log=1)
 *   noise    [B,W]  Gumbel(0,1) noise to add, or NULL -> drawn in-kernel from
 *                   Philox4x32-10 keyed by (seed, row, world)
 *   herbrand [W,H]  0/1 rows (define_herbrand_base, models/vael.py:179-187; W_adm)
 *   y_soft   [B,W]  out: softmax((log p + g)/tau)   (needed by backward)
 *   world_idx[B]    out: argmax
 *   world_h  [B,H]  out: ((1 - y_soft[idx]) + y_soft[idx]) * herbrand[idx]  (straight-through forward value) */
VAEL_API int vael_gumbel_world_forward(const float* scores, const float* noise, uint64_t seed, int64_t B, int32_t W,
                              int32_t is_log, float tau, const float* herbrand, int32_t H, float* y_soft,
                              int32_t* world_idx, float* world_h, void* stream);
/*   grad_world_h [B,H] -> grad_scores [B,W]  (gradient w.r.t. p when is_log=0, w.r.t. log p when is_log=1) */
VAEL_API int vael_gumbel_world_backward(const float* scores, const float* y_soft, const float* grad_world_h,
                               int64_t B, int32_t W, int32_t is_log, float tau, const float* herbrand,
                               int32_t H, float* grad_scores, void* stream);

/* ---- the training-mode logic path in one kernel per direction -----------------------
 * VAELModel.forward between the MLP and the decoder (models/vael.py:55-65) for the two-digit program family:
 * compute_facts_probability (:159-177) -> problog_inference (:200-225) -> log + F.gumbel_softmax(hard) (:61-62)
 * -> herbrand with define_herbrand_base (:82-84, :179-187).  No [B,W] tensor touches memory; the adjoint redraws
 * the noise from the same counters (Philox4x32-10 keyed by `seed`, or by the device word `seed_dev` when
 * given — a CUDA graph can refresh it per replay) instead of saving it.  With `noise` == NULL the race times behind
 * the Gumbel-max trick are drawn top-down (sampled world by inverse CDF, the others as winner + Exp / p): the same
 * joint law of (world, y_soft) as F.gumbel_softmax, two uniforms per row in the forward; the straight-through forward
 * value (1 - y) + y of the selected component (1 or 1 - 2^-24 in the reference) is then written as 1.  Covers circuits of fast_kind 2 whose two
 * annotated disjunctions have 9 or 10 values each (vael_logic_fused_supported); the Herbrand base is the
 * reference's: row i*N+k = [onehot(i), onehot(k)].
 *   logits   [B,2N]  MLP outputs                     noise  [B,W] injected Gumbel noise or NULL
 *   facts    [B,2N]  out or NULL                     query  [B]   out or NULL (selected per row by query_idx / scalar)
 *   world_idx[B]     out or NULL (arg-max world)     world_h[B,2N] out or NULL  ((1-y)+y at the two selected columns)
 *   y_soft   [B,W]   out or NULL (tests)                                                                   */
VAEL_API int vael_logic_fused_supported(const vael_circuit_t* c);
VAEL_API int vael_logic_train_forward(const vael_circuit_t* c, const float* logits, const int32_t* query_idx,
                             int32_t query_scalar, int64_t B, float eps, float tau, const float* noise,
                             uint64_t seed, const uint64_t* seed_dev, float* facts, float* query,
                             int32_t* world_idx, float* world_h, float* y_soft, void* stream);
/* autograd through all of the above (utils/mnist_utils/train.py:147): grad_world_h [B,2N], grad_query [B],
 * grad_facts [B,2N] (upstream gradient of the `facts` output) — each may be NULL (= 0) -> grad_logits [B,2N] */
VAEL_API int vael_logic_train_backward(const vael_circuit_t* c, const float* logits, const int32_t* query_idx,
                              int32_t query_scalar, int64_t B, float eps, float tau, const float* noise,
                              uint64_t seed, const uint64_t* seed_dev, const float* grad_world_h,
                              const float* grad_query, const float* grad_facts, float* grad_logits, void* stream);

/* ---- ELBO terms ---------------------------------------------------------------
 * loss_function (utils/mnist_utils/train.py:13-55; utils/mario_utils/train.py:11-69):
 *   recon = rec_loss LAPLACE: -mean_b sum_pix( -log 2 - |x - recon| );  MSE / BCE: torch's mean-reduced losses
 *   kl = mean_b -0.5*sum(1+logvar-mu^2-exp(logvar)),
 *   qbce  = mean_b -max(log q, -100),  loss = recon_w*recon + kl_w*kl + query_w*qbce.
 *   terms[4] device out: {loss, recon, kl, qbce}.
 * The same pass writes the gradients of `loss` (any of the grad_* may be NULL).
 * `workspace`: device scratch of vael_elbo_workspace_bytes() bytes, zero-filled once by
 * the caller before its first use (the kernel leaves it clean); per-CTA partial sums
 * are added in a fixed order by the last CTA, so the terms are deterministic.  A workspace
 * belongs to ONE call in flight: calls that may overlap (different streams) need their own. */
VAEL_API int64_t vael_elbo_workspace_bytes(void);
VAEL_API int vael_elbo_terms(const float* recon, const float* x, int64_t B, int64_t D, const float* mu,
                    const float* logvar, int32_t L, const float* query_prob, float recon_w, float kl_w,
                    float query_w, int32_t rec_loss, float* terms, float* grad_recon, float* grad_mu,
                    float* grad_logvar, float* grad_query, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAEL_B200_H */
