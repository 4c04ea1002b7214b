/* drin_b200 -- C ABI of the B200 (sm_100a) DRIN hot path.
 *
 * Plain pointers and sizes only: no torch types cross this boundary.  Every pointer is a DEVICE
 * pointer unless stated otherwise; `stream` is a cudaStream_t passed as void*.  All functions return
 * 0 on success, a non-zero drin_status otherwise; drin_last_error() returns the message for the
 * calling thread.  The library never allocates user-visible memory: outputs and the workspace are
 * owned by the caller (the PyTorch caching allocator in the Python shim).
 *
 * The reference (starreeze/drin) is pure Python/PyTorch and has no FFI of its own; each entry point
 * below names the reference code it replaces (file:line in the upstream repository).
 */
#ifndef DRIN_B200_H_
#define DRIN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum drin_status { DRIN_STATUS_OK = 0, DRIN_STATUS_BAD_ARG = 1, DRIN_STATUS_CUDA = 2, DRIN_STATUS_WORKSPACE = 3 };

/* numeric mode of the feature tensors and GEMM operands */
enum drin_precision {
  DRIN_FP32 = 0, /* fp32 features; GEMMs as split-bf16 (3 tensor-core passes), fp32-parity <= 1e-4 */
  DRIN_BF16 = 1  /* bf16 features; single-pass bf16 GEMMs, fp32 accumulation and fp32 tail         */
};

/* GEMM operand layouts (see drin_gemm) */
enum drin_gemm_layout { DRIN_GEMM_NT = 0, DRIN_GEMM_NN = 1, DRIN_GEMM_TN = 2 };

/* Problem description: the globals of common/args.py the hot path reads (args.py:25-36,45,52-57,72,85,101)
 * plus the batch size.  entity_tokens == 0 selects the WikiDiverse layout (entity text already pooled,
 * rank-3 tensors), > 0 the WikiMEL layout (rank-4 entity text + mask), exactly as drin/model.py:43-44,
 * 73-75,78-83 switch on tensor rank. */
typedef struct drin_config {
  int32_t batch;            /* B  mentions in this call                                   */
  int32_t candidates;       /* C  = num_candidates_model (real candidates + gold slot)    */
  int32_t mention_tokens;   /* Lm = max_mention_sentence_len (128)                        */
  int32_t entity_tokens;    /* Le = max_entity_attr_token_len (64) on WikiMEL, 0 on WikiDiverse */
  int32_t regions;          /* P  = resnet_num_region (49)                                */
  int32_t mention_objects;  /* Om = object_topk["mention"] (3)                            */
  int32_t entity_objects;   /* Oe = object_topk["entity"] (1)                             */
  int32_t embed_dim;        /* D  = gcn_embed_dim = bert_embed_dim (768)                  */
  int32_t resnet_dim;       /* R  = resnet_embed_dim (2048)                               */
  int32_t gcn_layers;       /* L  = num_gcn_layers (2)                                    */
  int32_t precision;        /* enum drin_precision                                        */
  int32_t training;         /* 1: keep what drin_backward needs in the workspace          */
  float edge_enabled[4];    /* gcn_edge_enabled, order tt, ti, it, ii                     */
  int32_t static_edges;     /* 0: gcn_edge_type "dynamic" (learned edge update, model.py:130-134);
                               1: "static" (masked edges pass through, model.py:135-136; w_u / w_v of every
                               layer receive no gradient)                                  */
  int32_t indexed;          /* 0: the 14 input tensors hold exactly this batch (row b = mention b);
                               1: they are resident feature TABLES and drin_inputs.mention_index /
                               entity_index select the rows of this batch (drin/data.py:85-108 on device) */
  int32_t vector_edges;     /* 0: gcn_edge_feature "scaler" (args.py:33): one scalar per edge;
                               1: "vector": every edge is a D-vector (model.py:202), messages are elementwise
                               products (model.py:139-146) and the dynamic edge update is
                               e' = sigmoid(W_m(cat[W_u u, W_v v] + e)) with W_u, W_v: D -> D/2 and
                               W_m: D -> D (model.py:112-116,133,148-152)                  */
} drin_config;

/* The 14 model inputs in the order of drin/model.py:164-180 (= drin/data.py:110-125).  Feature tensors
 * are fp32 (DRIN_FP32) or bf16 (DRIN_BF16); positions and masks are int64; scores and similarities fp32.
 * mention_text_mask is accepted for layout compatibility and never read (ghmfc.py:25-26). */
typedef struct drin_inputs {
  const void* mention_text_feature;    /* [B, Lm, D]                      */
  const int64_t* mention_text_mask;    /* [B, Lm]   (unused)              */
  const int64_t* mention_start_pos;    /* [B]                             */
  const int64_t* mention_end_pos;      /* [B]                             */
  const void* mention_image_feature;   /* [B, P, R]                       */
  const void* mention_object_feature;  /* [B, Om, 1, R]                   */
  const float* mention_object_score;   /* [B, Om]                         */
  const void* entity_text_feature;     /* WD [B, C, D]  | WM [B, C, Le, D] */
  const int64_t* entity_text_mask;     /* WD unused     | WM [B, C, Le]    */
  const void* entity_image_feature;    /* [B, C, R]  (WM [B, C, 1, R])    */
  const void* entity_object_feature;   /* [B, C, Oe, R] (WM [B, C, Oe, 1, R]) */
  const float* entity_object_score;    /* [B, C, Oe]                      */
  const float* miet_similarity;        /* [B, C]                          */
  const float* mtei_similarity;        /* [B, C]                          */
  /* Row selection when cfg->indexed != 0 (otherwise ignored; may be NULL).  What MELData.__getitem__ does on the
   * host (drin/data.py:85-108), done by the front-end kernel instead: the tensors above are whole-split tables
   * with a leading row dimension N instead of B.
   *   mention_index [B]    row m of mention b in every mention-side table and in miet / mtei similarity
   *   entity_index  [B, C] row of candidate (b, c) in the entity-side tables (WikiMEL: qid2idx lookup of
   *                        entity-name-raw, data.py:88-93).  NULL = WikiDiverse layout, row m * C + c (data.py:95-98). */
  const int64_t* mention_index;
  const int64_t* entity_index;
} drin_inputs;

/* Parameters (drin/model.py:21-24,111-119,159-162); all fp32.
 * Per GCN layer l: w_h, b_h, w_u, b_u, w_v, b_v, ln_w, ln_b [D, D] / [D]; with vector edges w_u, w_v are
 * [D/2, D] / [D/2] and w_m, b_m ([D, D], [D]; nn.Identity without parameters for scalar edges, model.py:112)
 * are read as well.  The same struct carries gradients. */
#define DRIN_MAX_LAYERS 8
typedef struct drin_layer_params {
  float* w_h; float* b_h; float* w_u; float* b_u; float* w_v; float* b_v; float* ln_w; float* ln_b;
  float* w_m; float* b_m;   /* vector edges only (NULL otherwise) */
} drin_layer_params;
typedef struct drin_params {
  float* w_mt; float* b_mt;   /* vertex_encoder.mention_text_encoder.final_layer.linear  [D, D], [D] */
  float* w_et; float* b_et;   /* vertex_encoder.entity_text_encoder.final_layer          [D, D], [D] */
  float* w_mi; float* b_mi;   /* vertex_encoder.mention_image_linear                     [D, R], [D] */
  float* w_ei; float* b_ei;   /* vertex_encoder.entity_image_linear                      [D, R], [D] */
  drin_layer_params layer[DRIN_MAX_LAYERS];
} drin_params;

const char* drin_last_error(void);
int drin_version(void);
/* sizeof(drin_config), sizeof(drin_inputs), sizeof(drin_params) as compiled: lets a binding check its struct mirrors */
void drin_struct_sizes(int32_t* config_bytes, int32_t* inputs_bytes, int32_t* params_bytes);

/* Bytes of caller-owned scratch drin_forward / drin_backward need for `cfg` (activations saved for the
 * backward pass live here when cfg->training != 0). */
int drin_workspace_bytes(const drin_config* cfg, size_t* bytes);

/* Model.forward (drin/model.py:164-209): vertex + edge encoders, L GCN layers, cosine scoring.
 * scores: [B, C] fp32. */
int drin_forward(const drin_config* cfg, const drin_inputs* in, const drin_params* params, void* workspace,
                 size_t workspace_bytes, float* scores, void* stream);

/* Backward of drin_forward through every parameter (what loss.backward() does after train.py:34).
 * dscores: [B, C] fp32.  grads: same layout as params; every tensor is OVERWRITTEN except the last
 * layer's w_u/b_u/w_v/b_v (and w_m/b_m with vector edges), which receive no gradient in the reference
 * (grad is None) and are left untouched.  Must follow a drin_forward with training = 1 on the same workspace. */
int drin_backward(const drin_config* cfg, const drin_inputs* in, const drin_params* params, void* workspace,
                  size_t workspace_bytes, const float* dscores, const drin_params* grads, void* stream);

/* Same as drin_backward; additionally records `layers_done_event` (a cudaEvent_t passed as void*, may be NULL) on
 * `stream` at the point where every gradient of the GCN layers and every bias gradient is final and only the four
 * input-projection weight gradients (w_mt, w_mi, w_et, w_ei) are still to come.  A data-parallel caller can start
 * all-reducing the layer gradients on another stream while those GEMMs run. */
int drin_backward_ex(const drin_config* cfg, const drin_inputs* in, const drin_params* params, void* workspace,
                     size_t workspace_bytes, const float* dscores, const drin_params* grads, void* layers_done_event,
                     void* stream);

/* TripletLoss (common/utils.py:35-43) forward and backward in one call, for the rows
 * [row_offset, row_offset + rows_local) of a global score matrix (data-parallel: scores of all ranks are
 * gathered first, because the loss couples every mention with every score of the batch).
 *   scores_all [B_glob, C] fp32, labels_all [B_glob, C-1] uint8 one-hot (all-zero row = gold not listed)
 *   loss       [1]   fp32: this call's share of the global loss, i.e. the sum over the local rows' scores
 *                    (the shares of all ranks add up to the reference loss; single rank: the loss itself)
 *   dscores    [rows_local, C] fp32 gradient of the GLOBAL loss w.r.t. the local rows (gold slot = 0)
 *   scratch    >= drin_loss_scratch_bytes(B_glob, C) bytes */
int drin_loss_scratch_bytes(int32_t batch_global, int32_t candidates, size_t* bytes);
int drin_triplet_loss(const float* scores_all, const uint8_t* labels_all, int32_t batch_global, int32_t candidates,
                      int32_t row_offset, int32_t rows_local, float margin, float* loss, float* dscores, void* scratch,
                      void* stream);

/* TopkAccuracy.update (common/utils.py:60-66): hits[k] += #rows whose gold score >= the k-th largest of the
 * row's C-1 real candidates (ties are hits).  topk: HOST array of n_k ints; hits: DEVICE int64[n_k]. */
int drin_topk_hits(const float* scores, const uint8_t* labels, int32_t batch, int32_t candidates, const int32_t* topk,
                   int32_t n_k, int64_t* hits, void* stream);

/* torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0) over one flat fp32 buffer
 * (train.py:55-56).  step is 1-based.  Elements with skip_mask[i] != 0 (may be NULL) are not updated
 * (parameters whose grad is None in the reference). */
int drin_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* skip_mask,
                   int64_t n, int32_t step, float lr, float beta1, float beta2, float eps, void* stream);

/* Same update with the 1-based step count read from DEVICE memory (int32), so the whole train step can be captured
 * in a CUDA graph and replayed: the caller increments *step_dev on the stream before each call. */
int drin_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* skip_mask,
                       int64_t n, const int32_t* step_dev, float lr, float beta1, float beta2, float eps, void* stream);

/* ---- stage-level entry points (unit parity tests, ncu) ------------------------------------------- */

/* fp32 -> split-bf16 planes (x ~= hi + lo); lo may be NULL (plain bf16 rounding). n % 4 == 0. */
int drin_split_planes(const float* x, void* hi, void* lo, int64_t n, void* stream);

/* One tcgen05 GEMM.  Operands are bf16 row-major matrices given as one plane (lo == NULL) or split
 * hi/lo planes.  NT: C[M,N] = A[M,K] B[N,K]^T;  NN: C = A[M,K] B[K,N];  TN: C = A[K,M]^T B[K,N].
 * C fp32 [M, ldc] (+bias per column); out_hi/out_lo optional split-bf16 copy of C.
 * ksplit > 1 needs partial >= ksplit*M*ldc floats.  reference != 0 runs the CUDA-core fp32 check kernel. */
int drin_gemm(int32_t layout, const void* a_hi, const void* a_lo, int32_t lda, const void* b_hi, const void* b_lo,
              int32_t ldb, int64_t M, int32_t N, int64_t K, float* C, int32_t ldc, const float* bias, void* out_hi,
              void* out_lo, int32_t ld_planes, int32_t ksplit, float* partial, int32_t reference, void* stream);
void drin_gemm_debug_mn_desc(int32_t lbo_bytes, int32_t sbo_bytes);

/* Front end (Avg.avg ghmfc.py:55-60, EntityEncoder pooling ghmfc.py:237-249, region mean model.py:41,
 * EdgeEncoder model.py:60-94, edge list model.py:201-204).  Outputs, all fp32:
 *   span [B, D], mimean [B, R], epool [B*C, D], edges [4, B*C] (tt, ti, it, ii; enable mask applied). */
int drin_frontend(const drin_config* cfg, const drin_inputs* in, float* span, float* mimean, float* epool,
                  float* edges, void* stream);

/* Measurement hooks (bench.py): number of kernels this library has launched so far, and per-stage device
 * time from CUDA events recorded on the launching stream.  Stage ids: 0 GEMM (flops = 2MNK per launch),
 * 1 front end, 2 GCN forward, 3 GCN backward, 4 scoring fwd/bwd, 5 loss, 6 Adam, 7 weight/plane prep.
 * drin_profile_collect fills four arrays of DRIN_PROFILE_STAGES entries and resets the records. */
#define DRIN_PROFILE_STAGES 8
long long drin_launch_count(void);
void drin_profile_enable(int32_t on);
int drin_profile_collect(double* ms, double* flops, double* bytes, long long* count);

/* Test hook: device pointer / shape of a named fp32 intermediate ("edges0", "x0", "h", "xm", "fu", "g",
 * "edges_out", "dz"; vector edges: "xa", "fu", "fv", "q", and for the first layer "edge_a", "edge_bv", "edge_w1") inside a
 * workspace planned for cfg.  Not part of the drop-in surface. */
int drin_debug_buffer(const drin_config* cfg, void* workspace, const char* name, int32_t layer, void** ptr,
                      int64_t* rows, int64_t* cols);

/* Test hook: force a kernel variant that is normally chosen from the problem size (value -1 = automatic).
 * Options: "score_fwd_variant", "score_bwd_variant", "layer_fwd_variant", "layer_bwd_variant" (0 CTA-per-mention / staged kernel,
 * 1 warp-per-mention kernel; "layer_bwd_variant" 2 = 1 plus the column-wise first-layer backward kernel), "defer_reductions"
 * (0: every split-K / column-sum reduction right after its producer), "row_slice_min" (fewest candidates per slice of the sliced
 * row kernels, default 8), "vec_bwd_width" (2 | 4 columns per thread of the vector-edge backward kernels), "vec_ctas_per_sm"
 * (upper bound on the resident-CTA grid of the vector-edge kernels, 0 = all).  Not part of the drop-in surface. */
int drin_debug_option(const char* name, int32_t value);

/* Test hook (the GPU pool has no compute-sanitizer): with drin_debug_option("workspace_guard", n) every buffer of the
 * workspace plan is followed by n bytes (rounded up to 256) no kernel may touch.  Returns the byte offsets of those
 * guard bands for cfg (offsets may be NULL to query *count); a test poisons the workspace, runs a step and checks the
 * bands.  "gemm_sm_cap" (debug option): upper bound on the SMs a persistent GEMM grid uses, 0 = all. */
int drin_debug_guard_regions(const drin_config* cfg, size_t* offsets, int32_t max_regions, int32_t* count,
                             size_t* guard_bytes);

#ifdef __cplusplus
}
#endif
#endif /* DRIN_B200_H_ */
