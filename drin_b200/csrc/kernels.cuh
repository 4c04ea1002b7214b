// Internal kernel-launcher declarations (host side).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace drin {

const char* last_error();

// elementwise.cu
int split_planes(cudaStream_t stream, const float* x, bf16* hi, bf16* lo, long long n);

// frontend.cu
struct FrontendArgs {
  int B, C, Lm, Le, P, Om, Oe, D, R;
  const void* mtf; const long long* start; const long long* end;
  const void* mif; const void* mof; const float* mos;
  const void* etf; const long long* emask; const void* eif; const void* eof; const float* eos;
  const float* miet; const float* mtei;
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
};
int gcn_layer_fwd(cudaStream_t stream, const LayerFwdArgs& a);
int mention_ln(cudaStream_t stream, int D, const float* h, long long rows, const float* gamma, const float* beta,
               float* x, bf16* x_hi, bf16* x_lo);
int rowdot(cudaStream_t stream, int D, const float* x, long long rows, const float* w, float* out);
int score_fwd(cudaStream_t stream, int D, const float* h_mt, const float* h_et, const float* gamma, const float* beta,
              int B, int C, float* scores);

}  // namespace drin
