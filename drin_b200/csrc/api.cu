// extern "C" boundary of libdrin_b200.so (declared in include/drin_b200.h).
#include "../../include/drin_b200.h"

#include "engine.cuh"
#include "kernels.cuh"
#include <string.h>

using namespace drin;

extern "C" {

const char* drin_last_error(void) { return drin::last_error(); }
int drin_version(void) { return 102; }
void drin_struct_sizes(int32_t* config_bytes, int32_t* inputs_bytes, int32_t* params_bytes) {
  if (config_bytes) *config_bytes = (int32_t)sizeof(drin_config);
  if (inputs_bytes) *inputs_bytes = (int32_t)sizeof(drin_inputs);
  if (params_bytes) *params_bytes = (int32_t)sizeof(drin_params);
}

int drin_split_planes(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  return split_planes((cudaStream_t)stream, x, (bf16*)hi, (bf16*)lo, n);
}

int drin_gemm(int32_t layout, const void* a_hi, const void* a_lo, int32_t lda, const void* b_hi, const void* b_lo,
              int32_t ldb, int64_t M, int32_t N, int64_t K, float* C, int32_t ldc, const float* bias, void* out_hi,
              void* out_lo, int32_t ld_planes, int32_t ksplit, float* partial, int32_t reference, void* stream) {
  if (layout < 0 || layout > 2) return fail(DRIN_ERR_ARG, "drin_gemm: bad layout %d", layout);
  Operand A, B;
  A.hi = (const bf16*)a_hi; A.lo = (const bf16*)a_lo; A.ld = lda;
  B.hi = (const bf16*)b_hi; B.lo = (const bf16*)b_lo; B.ld = ldb;
  if (layout == GEMM_TN) { A.rows = K; A.cols = (int)M; } else { A.rows = M; A.cols = (int)K; }
  if (layout == GEMM_NT) { B.rows = N; B.cols = (int)K; } else { B.rows = K; B.cols = N; }
  GemmEpilogue ep;
  ep.C = C; ep.ldc = ldc; ep.bias = bias;
  ep.out_hi = (bf16*)out_hi; ep.out_lo = (bf16*)out_lo; ep.ld_planes = ld_planes;
  if (reference) return gemm_reference_simt((cudaStream_t)stream, (GemmLayout)layout, A, B, M, N, K, ep);
  return gemm_tcgen05((cudaStream_t)stream, (GemmLayout)layout, A, B, M, N, K, ep, ksplit, partial);
}

void drin_gemm_debug_mn_desc(int32_t lbo_bytes, int32_t sbo_bytes) { gemm_debug_set_mn_desc(lbo_bytes, sbo_bytes); }

int drin_workspace_bytes(const drin_config* cfg, size_t* bytes) {
  if (!cfg || !bytes) return fail(DRIN_ERR_ARG, "drin_workspace_bytes: null argument");
  Workspace ws;
  DRIN_TRY(plan_workspace(*cfg, nullptr, nullptr, ws));
  *bytes = ws.bytes;
  return DRIN_OK;
}

int drin_forward(const drin_config* cfg, const drin_inputs* in, const drin_params* params, void* workspace,
                 size_t workspace_bytes, float* scores, void* stream) {
  if (!cfg || !in || !params) return fail(DRIN_ERR_ARG, "drin_forward: null argument");
  return forward(*cfg, *in, *params, workspace, workspace_bytes, scores, (cudaStream_t)stream);
}

int drin_backward(const drin_config* cfg, const drin_inputs* in, const drin_params* params, void* workspace,
                  size_t workspace_bytes, const float* dscores, const drin_params* grads, void* stream) {
  if (!cfg || !in || !params || !grads) return fail(DRIN_ERR_ARG, "drin_backward: null argument");
  return backward(*cfg, *in, *params, workspace, workspace_bytes, dscores, *grads, (cudaStream_t)stream);
}

int drin_backward_ex(const drin_config* cfg, const drin_inputs* in, const drin_params* params, void* workspace,
                     size_t workspace_bytes, const float* dscores, const drin_params* grads, void* layers_done_event,
                     void* stream) {
  if (!cfg || !in || !params || !grads) return fail(DRIN_ERR_ARG, "drin_backward_ex: null argument");
  return backward(*cfg, *in, *params, workspace, workspace_bytes, dscores, *grads, (cudaStream_t)stream,
                  (cudaEvent_t)layers_done_event);
}

int drin_loss_scratch_bytes(int32_t batch_global, int32_t candidates, size_t* bytes) {
  if (!bytes || batch_global <= 0) return fail(DRIN_ERR_ARG, "drin_loss_scratch_bytes: bad argument");
  *bytes = triplet_scratch_bytes(batch_global, candidates);
  return DRIN_OK;
}

int drin_triplet_loss(const float* scores_all, const uint8_t* labels_all, int32_t batch_global, int32_t candidates,
                      int32_t row_offset, int32_t rows_local, float margin, float* loss, float* dscores, void* scratch,
                      void* stream) {
  return triplet_loss((cudaStream_t)stream, scores_all, labels_all, batch_global, candidates, row_offset, rows_local,
                      margin, loss, dscores, scratch);
}

int drin_topk_hits(const float* scores, const uint8_t* labels, int32_t batch, int32_t candidates, const int32_t* topk,
                   int32_t n_k, int64_t* hits, void* stream) {
  return topk_hits((cudaStream_t)stream, scores, labels, batch, candidates, topk, n_k, (long long*)hits);
}

int drin_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* skip_mask,
                   int64_t n, int32_t step, float lr, float beta1, float beta2, float eps, void* stream) {
  return adam_step((cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, skip_mask, n, step, lr, beta1, beta2, eps);
}

int drin_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* skip_mask,
                       int64_t n, const int32_t* step_dev, float lr, float beta1, float beta2, float eps, void* stream) {
  return adam_step_dev((cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, skip_mask, n, step_dev, lr, beta1,
                       beta2, eps);
}

long long drin_launch_count(void) { return launch_count(); }
void drin_profile_enable(int32_t on) { prof::enable(on != 0); }
int drin_profile_collect(double* ms, double* flops, double* bytes, long long* count) {
  if (!ms || !flops || !bytes || !count) return fail(DRIN_ERR_ARG, "drin_profile_collect: null argument");
  prof::collect(ms, flops, bytes, count);
  return DRIN_OK;
}

int drin_frontend(const drin_config* cfg, const drin_inputs* in, float* span, float* mimean, float* epool,
                  float* edges, void* stream) {
  if (!cfg || !in) return fail(DRIN_ERR_ARG, "drin_frontend: null argument");
  DRIN_TRY(check_config(*cfg));
  const drin_config& c = *cfg;
  FrontendArgs fa{};
  fa.B = c.batch; fa.C = c.candidates; fa.Lm = c.mention_tokens; fa.Le = c.entity_tokens; fa.P = c.regions;
  fa.Om = c.mention_objects; fa.Oe = c.entity_objects; fa.D = c.embed_dim; fa.R = c.resnet_dim;
  fa.mtf = in->mention_text_feature; fa.start = (const long long*)in->mention_start_pos;
  fa.end = (const long long*)in->mention_end_pos; fa.mif = in->mention_image_feature;
  fa.mof = in->mention_object_feature; fa.mos = in->mention_object_score; fa.etf = in->entity_text_feature;
  fa.emask = (const long long*)in->entity_text_mask; fa.eif = in->entity_image_feature;
  fa.eof = in->entity_object_feature; fa.eos = in->entity_object_score; fa.miet = in->miet_similarity;
  fa.mtei = in->mtei_similarity;
  if (c.indexed) {
    fa.mention_index = (const long long*)in->mention_index;
    fa.entity_index = (const long long*)in->entity_index;
  }
  fa.span_f = span; fa.mim_f = mimean; fa.ep_f = epool; fa.edges = edges;
  return frontend((cudaStream_t)stream, fa, c.precision == DRIN_BF16);
}

// Test hook: force a kernel variant that is normally chosen from the problem size (-1 = automatic).
int drin_debug_option(const char* name, int32_t value) {
  if (!name) return fail(DRIN_ERR_ARG, "drin_debug_option: null name");
  if (!strcmp(name, "score_bwd_variant")) debug_set_score_bwd_variant(value);
  else if (!strcmp(name, "score_fwd_variant")) debug_set_score_fwd_variant(value);
  else if (!strcmp(name, "layer_fwd_variant")) debug_set_layer_fwd_variant(value);
  else if (!strcmp(name, "layer_bwd_variant")) debug_set_layer_bwd_variant(value);
  else if (!strcmp(name, "defer_reductions")) debug_set_defer_reductions(value);
  else if (!strcmp(name, "row_slice_min")) debug_set_row_slice_min(value);
  else if (!strcmp(name, "vec_bwd_width")) debug_set_vec_bwd_width(value);
  else if (!strcmp(name, "vec_ctas_per_sm")) debug_set_vec_ctas_per_sm(value);
  else if (!strcmp(name, "gemm_sm_cap")) debug_set_gemm_sm_cap(value);
  else if (!strcmp(name, "workspace_guard")) debug_set_workspace_guard(value);
  else if (!strcmp(name, "side_stream")) debug_set_side_stream(value);
  else return fail(DRIN_ERR_ARG, "drin_debug_option: unknown option '%s'", name);
  return DRIN_OK;
}

// Test hook: byte offsets of the guard bands of the workspace plan for cfg (see "workspace_guard"); each band is
// *guard_bytes long.  offsets may be NULL to query the count.
int drin_debug_guard_regions(const drin_config* cfg, size_t* offsets, int32_t max_regions, int32_t* count,
                             size_t* guard_bytes) {
  if (!cfg || !count || !guard_bytes) return fail(DRIN_ERR_ARG, "drin_debug_guard_regions: null argument");
  Workspace ws;
  DRIN_TRY(plan_workspace(*cfg, nullptr, nullptr, ws));
  const std::vector<size_t>& g = debug_workspace_guard_offsets();
  *count = (int32_t)g.size();
  *guard_bytes = debug_workspace_guard_bytes();
  if (offsets)
    for (int i = 0; i < (int)g.size() && i < max_regions; ++i) offsets[i] = g[i];
  return DRIN_OK;
}

// Test hook: device pointer and shape of a named intermediate inside a planned workspace.
int drin_debug_buffer(const drin_config* cfg, void* workspace, const char* name, int32_t layer, void** ptr,
                      int64_t* rows, int64_t* cols) {
  if (!cfg || !name || !ptr || !rows || !cols) return fail(DRIN_ERR_ARG, "drin_debug_buffer: null argument");
  Workspace ws;
  DRIN_TRY(plan_workspace(*cfg, nullptr, workspace, ws));
  const long long B = cfg->batch, BC = B * cfg->candidates, D = cfg->embed_dim;
  if (layer < 0 || layer >= cfg->gcn_layers) return fail(DRIN_ERR_ARG, "drin_debug_buffer: bad layer");
  const LayerWs& lw = ws.layer[layer];
  if (!strcmp(name, "edges0")) { *ptr = ws.edges0; *rows = 4; *cols = BC; }
  else if (!strcmp(name, "x0")) { *ptr = ws.x0; *rows = 2 * B + 2 * BC; *cols = D; }
  else if (!strcmp(name, "h")) { *ptr = lw.h; *rows = lw.rows; *cols = D; }
  else if (!strcmp(name, "xm")) { *ptr = lw.xm; *rows = 2 * B; *cols = D; }
  else if (!strcmp(name, "fu")) { *ptr = lw.fu; *rows = 2 * B; *cols = ws.vec ? D / 2 : D; }
  else if (!strcmp(name, "xa")) { *ptr = lw.xa; *rows = 2 * B + 2 * BC; *cols = D; }
  else if (!strcmp(name, "fv")) { *ptr = lw.fv; *rows = 2 * BC; *cols = D / 2; }
  else if (!strcmp(name, "q")) { *ptr = lw.q; *rows = 4 * BC; *cols = D; }
  else if (!strcmp(name, "edge_a")) { *ptr = lw.edge_a; *rows = 2 * B; *cols = D; }
  else if (!strcmp(name, "edge_bv")) { *ptr = lw.edge_bv; *rows = 2 * BC; *cols = D; }
  else if (!strcmp(name, "edge_w1")) { *ptr = lw.edge_w1; *rows = 1; *cols = D; }
  else if (!strcmp(name, "g")) { *ptr = lw.g; *rows = 2 * B; *cols = D; }
  else if (!strcmp(name, "edges_out")) { *ptr = lw.edges_out; *rows = 4; *cols = BC; }
  else if (!strcmp(name, "dz")) { *ptr = ws.dz; *rows = 2 * B + 2 * BC; *cols = D; }
  else return fail(DRIN_ERR_ARG, "drin_debug_buffer: unknown buffer '%s'", name);
  if (!*ptr) return fail(DRIN_ERR_ARG, "drin_debug_buffer: '%s' is not allocated for this config/layer", name);
  return DRIN_OK;
}

}  // extern "C"
