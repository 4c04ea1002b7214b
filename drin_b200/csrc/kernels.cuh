// Internal kernel-launcher declarations (host side).
#pragma once
#include "../../include/drin_b200.h"
#include "common.cuh"
#include "gemm.cuh"

namespace drin {

const char* last_error();

// prof.cu
namespace prof {
enum Cat { GEMM = 0, FRONTEND = 1, GCN_FWD = 2, GCN_BWD = 3, SCORE = 4, LOSS = 5, ADAM = 6, PREP = 7, NCAT = 8 };
void enable(bool on);
bool enabled();
struct Scope {
  Scope(cudaStream_t s, int cat, double flops = 0, double bytes = 0);
  ~Scope();
  cudaStream_t stream_;
  int idx_;
};
int collect(double* ms, double* flops, double* bytes, long long* count);
}  // namespace prof

// elementwise.cu
int split_planes(cudaStream_t stream, const float* x, bf16* hi, bf16* lo, long long n);
struct SplitJob {
  const float* x;
  bf16* hi;
  bf16* lo;        // null: plain bf16 rounding
  long long n4;    // number of float4 (element count / 4)
};
struct SplitJobs {
  static constexpr int MAX = 4 + 4 * DRIN_MAX_LAYERS;
  SplitJob job[MAX];
  int count = 0;
};
int split_planes_multi(cudaStream_t stream, const SplitJobs& jobs);   // one launch for all weight matrices of a step

// frontend.cu
struct FrontendArgs {
  int B, C, Lm, Le, P, Om, Oe, D, R;
  int slices;                    // candidate slices per mention (set by the launcher)
  const void* mtf; const long long* start; const long long* end;
  const void* mif; const void* mof; const float* mos;
  const void* etf; const long long* emask; const void* eif; const void* eof; const float* eos;
  const float* miet; const float* mtei;
  const long long* mention_index;   // [B] or null: table row of mention b (null: b)
  const long long* entity_index;    // [B, C] or null: entity-table row of candidate (b, c) (null: m * C + c)
  // outputs (any may be null)
  bf16 *span_hi, *span_lo;       // [B, D]
  bf16 *mim_hi, *mim_lo;         // [B, R]
  bf16 *ep_hi, *ep_lo;           // [B*C, D]
  bf16 *ei_hi, *ei_lo;           // [B*C, R]
  float* edges;                  // [4, B*C]
  float *span_f, *mim_f, *ep_f;  // fp32 copies for stage tests
};

int frontend(cudaStream_t stream, const FrontendArgs& a, bool bf16_features);

// gcn_fwd.cu
struct LayerFwdArgs {
  int B, C, D;
  bool full;                     // false: last layer (only the mt / et vertices are updated, no edge update)
  float en[4];                   // gcn_edge_enabled (model.py:122)
  const float* xm;               // [2B, D] activated mention vertices (mt rows, then mi rows)
  const float* x_et;             // [BC, D] candidate rows: activated (ln_gamma == null) or pre-LN h of the previous layer
  const float* x_ei;
  const float* ln_gamma;         // LayerNorm of the previous layer, applied on the fly with GELU
  const float* ln_beta;
  const float* edges_in;         // [4, BC]
  const float* g;                // [2B, D]  fu W_v           (full)
  const float* beta_u;           // [2B]     fu . b_v         (full)
  float* edges_out;              // [4, BC]                   (full)
  bf16* z_hi;                    // [2B+2BC, D] (full: mt, mi, et, ei) or [B+BC, D] (last: mt, et)
  bf16* z_lo;                    // null in bf16 mode
  int slices;                    // candidate slices per mention of the warp-per-(mention, slice) kernel (>= 1)
  float* acc_part;               // [B * slices][2][D] partial messages to the mention vertices (slices > 1)
};
int gcn_layer_fwd(cudaStream_t stream, const LayerFwdArgs& a);
// Work decomposition of the warp-autonomous row kernels: one warp per (mention, candidate slice).  1 for short
// candidate lists (WikiDiverse); for long ones (WikiMEL, C = 101) the number of slices that best fills whole rounds of
// 148 x 8 warps, with at least 8 candidates per slice.
int row_kernel_slices(int B, int C);
int mention_ln(cudaStream_t stream, int D, const float* h, long long rows, const float* gamma, const float* beta,
               float* x, bf16* x_hi, bf16* x_lo);
int rowdot(cudaStream_t stream, int D, const float* x, long long rows, const float* w, float* out);
int score_fwd(cudaStream_t stream, int D, const float* h_mt, const float* h_et, const float* gamma, const float* beta,
              int B, int C, float* scores);

// gcn_bwd.cu
struct ScoreBwdArgs {
  int B, C, D;
  const float* h_mt;       // [B, D]  pre-LN rows of the last layer
  const float* h_et;       // [BC, D]
  const float* gamma; const float* beta;
  const float* dscores;    // [B, C]
  bf16* dh_hi; bf16* dh_lo;   // [B+BC, D] (mt rows, then et rows)
  float* partials;         // [ctas][3][D]: dgamma, dbeta, db_h
  // candidate slices (row_kernel_slices): with slices > 1 the mention rows are finished by a second kernel
  int slices;
  float* slice_part;       // [B * slices][D + 32]: partial dL/da_m of the slice, coefficient at [D]
  float* partials2;        // [backward_ctas()][3][D]: column partials of the mention-finish kernel (slices > 1)
};
// returns via *used_partials2 whether partials2 was written (slices > 1 and the sliced kernel ran)
int score_bwd(cudaStream_t stream, const ScoreBwdArgs& a, bool* used_partials2);
void debug_set_score_bwd_variant(int v);
void debug_set_score_fwd_variant(int v);
void debug_set_layer_fwd_variant(int v);
void debug_set_layer_bwd_variant(int v);
void debug_set_defer_reductions(int v);
void debug_set_row_slice_min(int v);

struct LayerBwdArgs {
  int B, C, D;
  bool full;
  float en[4];
  const float* xm;            // [2B, D]
  const float* x_et; const float* x_ei; const float* ln_gamma; const float* ln_beta;   // as in LayerFwdArgs
  const float* edges_in;      // [4, BC]
  const float* dz;            // [rows, D] gradient w.r.t. z in this layer's row layout
  const float* g;             // [2B, D]   (full)
  const float* edges_out;     // [4, BC]   (full) sigmoid outputs
  const float* dedges_out;    // [4, BC]   (full) gradient w.r.t. edges_out
  bf16* dcand_hi; bf16* dcand_lo;   // candidate-row gradients in the [mt; mi; et; ei] layout (rows 2B..)
  float* dxm;                 // [2B, D] partial mention-row gradient
  float* dedges_in;           // [4, BC] or null (first layer: edges are inputs)
  bf16* dg_hi; bf16* dg_lo;   // [2B, D]   (full)
  float* dbeta;               // [2B]      (full)
  float* partials;            // [ctas][3][D]: ln -> (dgamma, dbeta, db_h) of the previous layer; else (db_et, db_ei, -)
  int slices;                 // candidate slices (row_kernel_slices); > 1: mention-side results by a finish kernel
  float* slice_part;          // [B * slices][4][D]: A_mt, A_mi, G_mt, G_mi partial sums of the slice
  float* slice_dbeta;         // [B * slices][2]
  int* partial_rows;          // HOST out (may be null): rows of `partials` the launch wrote (layer_bwd_ctas() unless the
                              // column-wise first-layer kernel ran, which may use up to backward_ctas() rows)
};
int gcn_layer_bwd(cudaStream_t stream, const LayerBwdArgs& a);

struct MentionBwdArgs {
  int B, D;
  const float* dxm;           // [2B, D]
  const float* dxu;           // [2B, D] or null
  const float* h_prev;        // [2B, D] pre-LN mention rows of the previous layer (ln)
  const float* ln_gamma; const float* ln_beta;
  bf16* out_hi; bf16* out_lo; // [2B, D] dh of the previous layer (ln) or dx0 (first layer)
  float* partials;            // [ctas][3][D]: ln -> (dgamma, dbeta, db_h); else (db_mt, db_mi, -)
};
int mention_bwd_finish(cudaStream_t stream, const MentionBwdArgs& a);
int dfu_finish(cudaStream_t stream, int D, float* dfu, const float* dbeta, const float* b_v, const float* fu,
               long long rows, bf16* out_hi, bf16* out_lo, float* partials);
int colsum_reduce(cudaStream_t stream, const float* src0, int ctas0, const float* src1, int ctas1, int nvec, int D,
                  float* out0, float* out1, float* out2);
// deferred column-sum reductions (same contract as colsum_reduce), all run by ONE launch at the end of backward
struct ColsumJob {
  const float* src0; const float* src1;
  int ctas0, ctas1, nvec;
  float* out0; float* out1; float* out2;
};
struct ColsumJobs {
  static constexpr int MAX = 4 * DRIN_MAX_LAYERS + 4;
  ColsumJob job[MAX];
  int count = 0;
  int add(const float* src0, int ctas0, const float* src1, int ctas1, int nvec, float* out0, float* out1, float* out2) {
    if (count >= MAX) return 1;
    job[count++] = ColsumJob{src0, src1, ctas0, ctas1, nvec, out0, out1, out2};
    return 0;
  }
};
int colsum_reduce_multi(cudaStream_t stream, const ColsumJobs& jobs, int D);
int backward_ctas();      // grid of score_bwd / mention_bwd_finish / dfu_finish (rows of their partial buffers)
int layer_bwd_ctas();     // grid of gcn_layer_bwd

// gcn_vec.cu -- GCN layer with vector edges (gcn_edge_feature == "vector", drin/model.py:97-153)
struct VecLayerArgs {
  int B, C, D;
  bool full;                  // all four vertex types are updated (every layer but the last)
  bool dyn;                   // this layer runs the dynamic edge update (full layers of a dynamic-edge model)
  float en[4];                // gcn_edge_enabled (model.py:122)
  const float* xa;            // [2B+2BC, D] activated vertices entering the layer (mt, mi, et, ei)
  // edge input, one of three forms (gcn_vec.cu: EdgeMode)
  const float* e_scalar;      // [4, BC] scalar input edges: the edges of a FIRST layer (SCALAR; also read by AFFINE)
  const float* q_in;          // [4BC, D] PRE-sigmoid vector edges produced by the previous layer (VECTOR); this and every
                              // other [4BC, D] matrix below is candidate-major: row r * 4 + k (candidate r, edge type k)
  const float* edge_a;        // AFFINE (the previous layer was a first layer): q_k[r] = edge_a[u_k B + b] + edge_bv[v_k BC + r]
  const float* edge_bv;       //   + en_k e_scalar[k, r] edge_w1;  edge_a [2B, D] = fu W_m[:, :H]^T + b_m, edge_bv [2BC, D] =
  const float* edge_w1;       //   fv W_m[:, H:]^T, edge_w1 [D] = W_m 1
  // forward
  const float* fu;            // [2B, D/2]  W_u xm + b_u   (dyn)
  const float* fv;            // [2BC, D/2] W_v xv + b_v   (dyn; et rows, then ei rows)
  bf16* z_hi; bf16* z_lo;     // [2B+2BC, D] (full) or [B+BC, D] (last: mt, et): A operand of the W_h GEMM
  bf16* m_hi; bf16* m_lo;     // [4BC, D] cat[fu, fv] + E_k: A operand of the W_m GEMM (dyn)
  // backward
  const float* dz;            // gradient w.r.t. z, same row layout
  const float* dm;            // [4BC, D] gradient w.r.t. m (dyn)
  float* dxa;                 // [2B+2BC, D] gradient w.r.t. xa (message paths only; W_u / W_v paths are GEMMs)
  bf16* dfu_hi; bf16* dfu_lo; // [2B, D/2]  (dyn)
  bf16* dfv_hi; bf16* dfv_lo; // [2BC, D/2] (dyn)
  bf16* dq_hi; bf16* dq_lo;   // [4BC, D] gradient w.r.t. q_in (VECTOR)
  bf16* da_hi; bf16* da_lo;   // [2B, D]  gradient w.r.t. edge_a  (AFFINE)
  bf16* dbv_hi; bf16* dbv_lo; // [2BC, D] gradient w.r.t. edge_bv (AFFINE)
  float* partials;            // [ctas][3][D]: sum dq (b_m gradient of the previous layer); [b_u | b_v] gradient of this
                              // layer; sum en_k e_k dq_k (edge_w1 gradient, AFFINE)
};
int rowsum(cudaStream_t stream, const float* w, int rows, int cols, float* out);
int wm_fixup(cudaStream_t stream, const float* dwa, const float* dwb, const float* dw1, const float* db_m,
             const float* w_m, int D, float* dw_m, float* db_u, float* db_v);
void debug_set_vec_bwd_width(int v);
void debug_set_vec_ctas_per_sm(int v);
void debug_set_gemm_sm_cap(int v);
int vec_layer_fwd(cudaStream_t stream, const VecLayerArgs& a);
int vec_layer_bwd(cudaStream_t stream, const VecLayerArgs& a, int* partial_rows);
int vec_layer_ctas();
struct VecRowsBwdArgs {
  long long B, BC;
  int D;
  const float* d0;            // [2B+2BC, D] gradient w.r.t. the activated vertices
  const float* d1;            // [2B+2BC, D] second addend (W_u / W_v data gradients) or null
  const float* h_prev;        // [2B+2BC, D] pre-LayerNorm rows of the previous layer (ln)
  const float* ln_gamma; const float* ln_beta;   // null: first layer, the rows are projection outputs
  bf16* out_hi; bf16* out_lo; // [2B+2BC, D]
  float* partials;            // [vec_rows_ctas()][4][D]: ln -> (dgamma, dbeta, db_h, -); else (db_mt, db_mi, db_et, db_ei)
};
int vec_rows_bwd(cudaStream_t stream, const VecRowsBwdArgs& a);
int vec_rows_ctas();
// column sums of the per-CTA partials above; every job of a backward pass is reduced by ONE launch at its end
struct StridedColsumJob {
  const float* src;
  float* out;
  long long stride;           // floats between consecutive partial rows
  int count, n;               // partial rows, columns
};
struct StridedColsumJobs {
  static constexpr int MAX = 7 * DRIN_MAX_LAYERS;
  StridedColsumJob job[MAX];
  int count = 0;
  int add(const float* src, int cnt, long long stride, int n, float* out) {
    if (count >= MAX) return 1;
    job[count++] = StridedColsumJob{src, out, stride, cnt, n};
    return 0;
  }
};
int strided_colsum_multi(cudaStream_t stream, const StridedColsumJobs& jobs);

// loss.cu
size_t triplet_scratch_bytes(int B, int C);
int triplet_loss(cudaStream_t stream, const float* scores, const unsigned char* labels, int B, int C, int row0,
                 int rows, float margin, float* loss, float* dscores, void* scratch);
int topk_hits(cudaStream_t stream, const float* scores, const unsigned char* labels, int B, int C, const int* topk,
              int nk, long long* hits);
// adam.cu
int adam_step(cudaStream_t stream, float* p, const float* g, float* m, float* v, const unsigned char* skip, long long n,
              int step, float lr, float b1, float b2, float eps);

int adam_step_dev(cudaStream_t stream, float* p, const float* g, float* m, float* v, const unsigned char* skip,
                  long long n, const int* step_dev, float lr, float b1, float b2, float eps);

}  // namespace drin
