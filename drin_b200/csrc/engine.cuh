// Workspace plan and forward/backward orchestration of the DRIN hot path.
#pragma once
#include <vector>
#include "../../include/drin_b200.h"
#include "kernels.cuh"

namespace drin {

struct Planes {
  bf16* hi = nullptr;
  bf16* lo = nullptr;   // null in bf16 mode
};

struct LayerWs {
  bool full = false;           // all four vertex types are updated (every layer but the last)
  bool dyn = false;            // dynamic edge update runs in this layer (full && !static_edges)
  long long rows = 0;          // rows of z / h: full 2B+2BC (mt, mi, et, ei), last B+BC (mt, et)
  Planes w_h, w_u, w_v;        // weight planes (w_u / w_v only for full layers)
  float* xm = nullptr;         // [2B, D] activated mention vertices entering this layer (layer 0: alias of x0)
  Planes xm_p;                 // planes of xm (full layers)
  float* fu = nullptr;         // [2B, D] W_u xm + b_u
  Planes fu_p;
  float* g = nullptr;          // [2B, D] fu W_v
  float* beta_u = nullptr;     // [2B]    fu . b_v
  Planes z;                    // [rows, D] A operand of the W_h GEMM
  float* h = nullptr;          // [rows, D] pre-LayerNorm output of W_h
  float* edges_out = nullptr;  // [4, BC] (full)
  // ---- vector edges (gcn_edge_feature "vector" with a dynamic edge update; see gcn_vec.cu) ----
  Planes w_m;                  // [D, D] planes (dyn); w_u / w_v are [D/2, D] here
  float* xa = nullptr;         // [2B+2BC, D] activated vertices entering this layer (layer 0: alias of x0)
  Planes xa_p;                 // planes of xa (dyn: A operand of the W_u / W_v GEMMs)
  float* fv = nullptr;         // [2BC, D/2] W_v xv + b_v (dyn); fu above is [2B, D/2]
  Planes m_p;                  // [4BC, D] cat[fu, fv] + E_k (dyn, not the first layer): A operand of the W_m GEMM
  float* q = nullptr;          // [4BC, D] pre-sigmoid edge outputs W_m m + b_m (dyn, not the first layer)
  // first layer of the vector path (its own edges are scalars): edge outputs in affine form, see gcn_vec.cu
  bool affine = false;
  Planes fv_p;                 // [2BC, D/2] planes of fv (fu_p above: [2B, D/2])
  float* edge_a = nullptr;     // [2B, D]  fu W_m[:, :H]^T + b_m
  float* edge_bv = nullptr;    // [2BC, D] fv W_m[:, H:]^T
  float* edge_w1 = nullptr;    // [D]      W_m 1
};

struct Workspace {
  size_t bytes = 0;
  Planes w_mt, w_et, w_mi, w_ei;
  Planes span, mim, epool, eimg;     // projection A operands
  float* edges0 = nullptr;           // [4, BC]
  int slices = 1;                    // candidate slices per mention of the row kernels (row_kernel_slices)
  float* acc_part = nullptr;         // [B * slices][2][D] partial messages of sliced mentions (slices > 1)
  float* x0 = nullptr;               // [2B+2BC, D] projection outputs (mt, mi, et, ei)
  Planes xm0_p;                      // planes of the first 2B rows of x0
  LayerWs layer[DRIN_MAX_LAYERS];
  // ---- backward scratch ----
  Planes dh;                         // [2B+2BC, D] gradient w.r.t. h (A of dZ GEMM, A of dW_h GEMM)
  float* dz = nullptr;               // [2B+2BC, D]
  float* dxm = nullptr;              // [2B, D] gradient w.r.t. activated mention vertices (partial)
  float* dxu = nullptr;              // [2B, D] dfu W_u
  float* dg = nullptr;               // [2B, D]
  Planes dg_p;
  float* dbeta = nullptr;            // [2B]
  float* dfu = nullptr;              // [2B, D]
  Planes dfu_p;
  float* dedges = nullptr;           // [4, BC] gradient w.r.t. the edges entering the layer above
  Planes dx0;                        // [2B+2BC, D] gradient w.r.t. projection outputs (A of projection dW GEMMs)
  float* slice_part = nullptr;       // [B * slices][4][D] per-slice sums of the sliced backward row kernels (slices > 1)
  float* slice_dbeta = nullptr;      // [B * slices][2]
  float* partial = nullptr;          // split-K partial-sum arena: one region per weight-gradient GEMM of the pass
  size_t partial_floats = 0;
  float* colsum = nullptr;           // per-CTA partial column sums (bias / LayerNorm gradients): one region per producer
  size_t colsum_floats = 0;
  int colsum_ctas = 0;
  int ksplit = 1;
  // ---- vector edges ----
  bool vec = false;                  // the vector-edge path runs (vector_edges && !static_edges && L > 1)
  Planes x0_p;                       // planes of all rows of x0 (xm0_p aliases its first 2B rows)
  float* dxa = nullptr;              // [2B+2BC, D] gradient w.r.t. the activated vertices (message paths)
  float* dxuv = nullptr;             // [2B+2BC, D] W_u / W_v data gradients
  float* dm = nullptr;               // [4BC, D] gradient w.r.t. m
  Planes dq_p;                       // [4BC, D] gradient w.r.t. the pre-sigmoid edge outputs of the layer below
  Planes dfv_p;                      // [2BC, D/2]; dfu_p above is [2B, D/2]
  Planes da_p;                       // [2B, D]  gradient w.r.t. edge_a of the first layer
  Planes dbv_p;                      // [2BC, D] gradient w.r.t. edge_bv of the first layer
  float* dwm_a = nullptr;            // [D, D/2] dW_m[:, :H] of the first layer (assembled by wm_fixup)
  float* dwm_b = nullptr;            // [D, D/2] dW_m[:, H:]
  float* dw1 = nullptr;              // [D]      gradient w.r.t. edge_w1
  float* vec_part = nullptr;         // [L][vec_layer_ctas()][3][D] column partials of vec_layer_bwd (one region per layer:
  float* rows_part = nullptr;        // [L][vec_rows_ctas()][4][D]  ... of vec_rows_bwd       all reduced by one launch)
};

// vector edges with static edges (or one layer) never leave the scalar representation: model.py:135-136 passes the
// masked, expanded input edges through, so the scalar kernels compute exactly the same thing.
inline bool vector_path(const drin_config& c) { return c.vector_edges && !c.static_edges && c.gcn_layers > 1; }

int check_config(const drin_config& c);
// Carve `base` (may be null: size query) into the buffers above.
int plan_workspace(const drin_config& c, const drin_inputs* in, void* base, Workspace& ws);
void debug_set_workspace_guard(int bytes);
size_t debug_workspace_guard_bytes();
const std::vector<size_t>& debug_workspace_guard_offsets();

// Fork / join of ONE side stream per device, so that GEMMs without a data dependency run concurrently: a persistent
// GEMM leaves SMs idle in its last, partially filled round of tiles (528 tile pairs over 74 CTA pairs = 7.13 rounds) and
// the 2B-row GEMMs of the mention side fill a third of the machine; CTAs of an independent GEMM on the other stream take
// those SMs as they free up.  The pattern (event recorded on the origin, waited on by the side stream, joined back) is
// legal under CUDA-graph stream capture.  Off while the per-launch profiler runs (its event pairs assume serial launches).
class SideStream {
 public:
  explicit SideStream(cudaStream_t main);
  bool on() const { return side_ != nullptr; }
  cudaStream_t fork();                 // side stream, ordered after everything enqueued on main so far (main if off)
  int join();                          // main waits for everything enqueued on the side stream
 private:
  cudaStream_t main_;
  cudaStream_t side_ = nullptr;
  cudaEvent_t fork_ev_ = nullptr, join_ev_ = nullptr;
  bool forked_ = false;
};
void debug_set_side_stream(int v);

int forward(const drin_config& c, const drin_inputs& in, const drin_params& p, void* workspace, size_t workspace_bytes,
            float* scores, cudaStream_t stream);
int backward(const drin_config& c, const drin_inputs& in, const drin_params& p, void* workspace,
             size_t workspace_bytes, const float* dscores, const drin_params& grads, cudaStream_t stream,
             cudaEvent_t layers_done = nullptr);

}  // namespace drin
