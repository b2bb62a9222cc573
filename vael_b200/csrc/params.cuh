// Parameter blocks and host launchers shared between the kernel translation units and capi.cu.
#pragma once
#include <string>

#include "common.cuh"

namespace vael {

// ---- circuit_ad2.cu --------------------------------------------------------
struct Ad2Params {
  const float* facts;        // [B,F]
  const int32_t* qidx;       // [B] or null
  int32_t qscalar;
  int32_t n_queries;
  const uint32_t* qmask;     // [Q][MW] world membership of every query node
  long long B;
  // forward
  float* worlds;             // [B,W] or null
  float* query;              // [B] or null
  // backward
  const float* grad_worlds;  // [B,W] or null
  const float* grad_query;   // [B] or null
  float* grad_facts;         // [B,F]
  int32_t l2_hint;           // 0 none, 1 = evict_first on the streaming stores, 2 = on loads and stores
};
constexpr int kQMaxSlots = 1024;    // (query, first-AD value, slot) entries the all-queries list kernel takes in its parameter block
constexpr int kQMaxQueries = 64;
struct Ad2QParams {           // all query nodes at once
  const float* facts;         // [B,F]
  long long B;
  int32_t n_queries;
  int32_t diagonal;           // query s holds exactly the worlds (j, k) with j + k = s (addition-shaped)
  float* queries_all;         // [B,Q]
  // Member lists, grouped by the FIRST annotated disjunction's value i (world order = i-major): for query q and
  // every i, `len[q]` slots holding the byte offset (k * 128) of the second literal's row in the kernel's
  // column-major p2 tile, padded with the tile's zero row (NB * 128).  p1[i] is a register operand (i is a
  // compile-time loop index): ONE gather per member.  The table lives in the PARAMETER block (constant bank):
  // its warp-uniform fetches cost no shared-memory wavefront.
  uint16_t ptr[kQMaxQueries];     // first slot of query q
  uint8_t len[kQMaxQueries];      // slots per (q, i)
  uint16_t slot[kQMaxSlots];
};
cudaError_t ad2_queries_all(int na, int nb, const Ad2QParams& P, bool norm, int num_sms, cudaStream_t st);
bool ad2_supported(int na, int nb);
int l2_hint_env();  // VAEL_L2_HINT (experiments)
cudaError_t ad2_forward(int na, int nb, const Ad2Params& P, bool norm, bool evid, int num_sms, cudaStream_t st);
cudaError_t ad2_backward(int na, int nb, const Ad2Params& P, bool norm, int num_sms, cudaStream_t st);

// ---- circuit_adw.cu (wide world levels: three annotated disjunctions, TMA 2-D tiles) ----
struct AdwParams {
  const float* facts;        // [B,F]
  const int32_t* qidx;       // [B] or null
  int32_t qscalar;
  int32_t n_queries;
  int32_t n_chunks;          // size of the leading AD = number of column chunks
  const uint32_t* qmask_c;   // [Q][n_chunks][MWC] world membership of every query node, chunk by chunk
  long long B;
  float* worlds;             // [B,W] or null
  float* query;              // [B] or null
  const float* grad_worlds;  // [B,W] or null
  const float* grad_query;   // [B] or null
  float* grad_facts;         // [B,F]
  int32_t l2_hint;           // as Ad2Params
};
bool adw_supported(int np, int na, int nb);
bool adw_usable(const float* wide_tensor);
cudaError_t adw_forward(int np, int na, int nb, const AdwParams& P, bool norm, bool evid, int num_sms, cudaStream_t st);
cudaError_t adw_backward(int np, int na, int nb, const AdwParams& P, bool norm, int num_sms, cudaStream_t st);
// all query nodes at once, addition-shaped queries (query s = worlds (i, j, k) with i + j + k = s)
cudaError_t adw_queries_diag(int np, int na, int nb, const float* facts, long long B, float* queries_all, bool norm,
                             int num_sms, cudaStream_t st);

// ---- circuit_lt.cu (literal-table circuits in the log semiring: Mario's admissible worlds) ----
bool lt_supported(int F, int W);
bool lt_masks_fit(int F, int W, const uint32_t* masks);  // the per-world / per-column lists fit the kernel's tables
cudaError_t lt_forward(int F, int W, const uint32_t* masks, const float* facts, long long B, float* worlds, int num_sms,
                       cudaStream_t st);
cudaError_t lt_backward(int F, int W, const uint32_t* masks, const float* facts, const float* grad_worlds, long long B,
                        float* grad_facts, int num_sms, cudaStream_t st);

// ---- circuit_generic.cu ----------------------------------------------------
constexpr int kMaxLevels = 24;
struct GenProgram;  // device-resident schedule of a circuit (ops per level, parent records), built once
GenProgram* gen_program_create(const vael_circuit_desc_t* d, cudaError_t* err);
void gen_program_destroy(GenProgram* g);
struct GenCall {  // the per-call part
  const float* facts;
  const int32_t* qidx;
  int32_t qscalar;
  long long B;
  uint32_t flags;
  const uint32_t* qmask;  // [Q][mask_words] or null (evidence mode unsupported then)
  int32_t mask_words;
  float* worlds;          // [B,W] or null
  float* query;           // [B] or null
  float* queries_all;     // [B,Q] or null
  const float* grad_worlds;
  const float* grad_query;
  float* grad_facts;
};
bool gen_fits(const GenProgram* g, bool backward);
cudaError_t gen_forward(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st);
cudaError_t gen_backward(const GenProgram* g, const GenCall& C, int num_sms, cudaStream_t st);

// ---- circuit_jit.cu (circuit-specialised straight-line kernels, NVRTC) ----
struct JitProgram;
bool jit_eligible(const vael_circuit_desc_t* d);
JitProgram* jit_create(const vael_circuit_desc_t* d);   // never null; jit_ok() tells whether the kernels exist
void jit_destroy(JitProgram* J);
bool jit_ok(const JitProgram* J);
bool jit_can(const JitProgram* J, uint32_t flags, const float* wide_tensor);
const char* jit_log(const JitProgram* J);
cudaError_t jit_forward(const JitProgram* J, const GenCall& C, int num_sms, cudaStream_t st);
cudaError_t jit_backward(const JitProgram* J, const GenCall& C, int num_sms, cudaStream_t st);
int jit_compile_check(const vael_circuit_desc_t* d, std::string* log, size_t* cubin_bytes);
std::string jit_source(const vael_circuit_desc_t* d);

// ---- sampling.cu -----------------------------------------------------------
struct GumbelParams {
  const float* scores;    // [B,W]
  const float* noise;     // [B,W] or null
  uint64_t seed;
  long long B;
  int32_t W, H, is_log;
  float tau;
  const float* herbrand;  // [W,H]
  float* y_soft;          // [B,W]
  int32_t* world_idx;     // [B]
  float* world_h;         // [B,H]
  const float* grad_world_h;  // [B,H]
  float* grad_scores;         // [B,W]
  int32_t hb_in_smem;         // set by the launcher: Herbrand table staged in shared memory
};
cudaError_t gumbel_forward(const GumbelParams& P, int num_sms, cudaStream_t st);
cudaError_t gumbel_backward(const GumbelParams& P, int num_sms, cudaStream_t st);
// circuit_adw.cu: 2-D tensor map of a row-major [B,W] fp32 tensor with a [32 rows, C columns] box, written to the
// 128 bytes (64-byte aligned) at `map128`; false when the driver entry point is missing or `base` is not 16-byte aligned
bool make_row_box_map(void* map128, const float* base, long long B, int W, int C);
// gumbel_rows.cu: lane <-> row kernels for W % 4 == 0, W <= 128, H <= 32 (the dispatchers above pick them)
bool gumbel_rows_ok(const GumbelParams& P, bool backward);
cudaError_t gumbel_rows_forward(const GumbelParams& P, int num_sms, cudaStream_t st);
cudaError_t gumbel_rows_backward(const GumbelParams& P, int num_sms, cudaStream_t st);
cudaError_t facts_forward(const float* logits, long long B, int n_groups, int group, float eps, float hi, float* facts,
                          int num_sms, cudaStream_t st);
cudaError_t facts_backward(const float* logits, const float* grad_facts, long long B, int n_groups, int group,
                           float eps, float hi, float* grad_logits, int num_sms, cudaStream_t st);

// ---- logic_fused.cu (training-mode logic path in one kernel per direction) ----
struct FusedParams {
  const float* logits;       // [B,2N]
  const int32_t* qidx;       // [B] or null
  int32_t qscalar;
  int32_t n_queries;
  const uint32_t* qmask;     // [Q][MW] world membership of every query node
  long long B;
  float eps, tau;
  const float* noise;        // [B,W] Gumbel noise to add, or null -> Philox4x32-10 keyed by the seed
  uint64_t seed;
  const uint64_t* seed_dev;  // device word holding the seed (overrides `seed`): lets a CUDA graph draw a fresh
                             // seed per replay with a captured generator op
  // forward outputs (each may be null)
  float* facts;              // [B,2N]
  float* query;              // [B]
  int32_t* world_idx;        // [B]
  float* world_h;            // [B,2N]
  float* y_soft;             // [B,W]  (tests)
  // backward
  const float* grad_world_h; // [B,2N] or null
  const float* grad_query;   // [B] or null
  const float* grad_facts;   // [B,2N] or null: upstream gradient of the `facts` output
  float* grad_logits;        // [B,2N]
};
bool fused_supported(int na, int nb);
cudaError_t fused_forward(int n, const FusedParams& P, int num_sms, cudaStream_t st);
cudaError_t fused_backward(int n, const FusedParams& P, int num_sms, cudaStream_t st);

// ---- elbo.cu ---------------------------------------------------------------
struct ElboParams {
  const float* recon;
  const float* x;
  long long B, D;
  const float* mu;
  const float* logvar;
  int32_t L;
  const float* query_prob;  // [B] or null
  float recon_w, kl_w, query_w;
  int32_t rec_loss;  // VAEL_REC_*
  float* terms;  // [4] {loss, recon, kl, qbce}
  float* grad_recon;
  float* grad_mu;
  float* grad_logvar;
  float* grad_query;
  unsigned int* counter;  // workspace[0]
  double* partials;       // workspace + 16 bytes: [grid][3]
};
size_t elbo_workspace_bytes();
cudaError_t elbo_terms(ElboParams P, void* workspace, int num_sms, cudaStream_t st);

}  // namespace vael
