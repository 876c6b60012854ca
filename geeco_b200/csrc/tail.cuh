#pragma once
#include "common.cuh"

// Graph variants (which *_concatenation of graph.py:123-192 builds the LSTM input, which conv8 maps feed it)
enum { VAR_GEECOF = 0,        // representation_concatenation_v2 [obs | dyn | jnt | tgt], one step        graph.py:386-407
       VAR_SEQ_CONSTANT = 1,  // representation_concatenation    [feat_t | jnt_t | tgt_feat], K steps     graph.py:365-367
       VAR_SEQ_RESIDUAL = 2,  // state_concatenation             [tgt_feat - feat_t | jnt_t], K steps     graph.py:368-370
       VAR_SEQ_DYNDIFF = 3,   // representation_concatenation    [feat_t | jnt_t | diff_feat_t], K steps  graph.py:371-379
       VAR_VMC = 4 };         // state_concatenation             [feat_t | jnt_t], K steps (e2e_vmc)      graph.py:304-309

// conv8 maps (fp32) of the encoder groups and how a step's LSTM input is cut out of them.  Images of a group are
// ordered step-major: image t*N + n is frame t of sample n; the target frame of the constant / residual variants is
// image K*N + n of group 0.
struct StateMap {
  int variant;
  int N, T, K, J;              // batch rows, LSTM steps, window size, dim_jnt_state
  int D0, D1, D2;              // conv8 widths of groups 0..2
  int per, xdim, ld, Hl;       // channels per 2x2 cell, 4*per, xdim + Hl (row stride of a state), dim_h_lstm
  const float* y[3];           // conv8 outputs per group  [imgs][2][2][D]
  float* g[3];                 // dL/d(pre-activation) of conv8 per group (same layout)
  int ring_start;              // physical slot of the oldest frame in jnt_state [N][K][J]
};

// One dense head on fc1 with its loss (graph.py:229-259, :430-500; estimator.py:205-237)
struct HeadSpec {
  int col, width;              // columns [col, col+width) of the heads matrix
  int kind;                    // 0: tf.losses.mean_squared_error, 1: one_hot + tf.losses.softmax_cross_entropy
  int slot;                    // index into the losses vector (include/geeco_b200.h)
  float weight;                // factor of the term in the total loss (lambda_aux for the auxiliary poses, cartesian)
  int aux;                     // 1: summed with the other auxiliary terms before the weight is applied (estimator.py:224-225)
  const float* w; const float* b;      // [Fc][width], [width]
  float* gw; float* gb;
  const float* target; int tstride, toff;   // row n of the target: target + n*tstride + toff  (CE: the class column)
};
struct TailHeads { int nheads, NH; HeadSpec h[5]; };

struct TailDims {
  int N, Hl, Fc, G;
};

int launch_build_states(const StateMap& sm, const float* jnt, const float* m_prev, const unsigned char* reset_mask,
                        float* states, cudaStream_t st);
int launch_scatter_dstates(const StateMap& sm, const float* dstates, cudaStream_t st);
// c_prev / m written for step t; m_next (optional) = the m part of the NEXT step's state row (stride ld_next)
int launch_lstm_cell(int N, int Hl, const float* gates, const float* c_prev, const unsigned char* reset_mask,
                     float* c_out, float* m_out, float* state_out, float* m_next, int ld_next, cudaStream_t st);
int launch_lstm_cell_bwd(int N, int Hl, const float* gates, const float* c_prev, const unsigned char* reset_mask,
                         const float* dm, int ld_dm, const float* dc_in, float* dgates, float* dc_prev, cudaStream_t st);
// heads_out / fc1_out / losses_out / state_out2: the caller's copies (geeco_outputs), NULL = not wanted
int launch_tail_fwd(const TailDims& d, const TailHeads& th, const float* w_fc1, const float* b_fc1, const float* m,
                    float* fc1, float* heads, float* loss_parts, float* dheads, int with_loss, float* heads_out,
                    float* fc1_out, cudaStream_t st);
int launch_tail_fwd_fused(const TailDims& d, const TailHeads& th, const float* w_fc1, const float* b_fc1,
                          const float* partial, int slices, const float* lstm_bias, float* gates, const float* c_prev,
                          const unsigned char* reset_mask, float* c_out, float* m_out, float* state_out, float* fc1,
                          float* heads, float* loss_parts, float* dheads, int with_loss, float* heads_out, float* fc1_out,
                          float* state_out2, cudaStream_t st);
int launch_loss_reduce(const TailDims& d, const TailHeads& th, const float* loss_parts, const float* reg_term,
                       float* losses, float* losses_out, cudaStream_t st);
int launch_lstm_dstate_scatter(const float* dgates, const float* W, const StateMap& sm, __nv_bfloat16* const* g_bf16, int ncols,
                               cudaStream_t st);
int launch_lstm_wgrad(const float* states, const float* dgates, float* dW, float* db, int R, int ld, int ncols, cudaStream_t st);
// dm_out == NULL: one-step graphs, the LSTM cell backward is fused (writes dgates); else writes dL/dm_T [N][Hl]
int launch_tail_bwd(const TailDims& d, const TailHeads& th, const float* w_fc1, float* gw_fc1, float* gb_fc1,
                    const float* m, const float* fc1, const float* dheads, const float* gates, const float* c_prev,
                    const unsigned char* reset_mask, float* dfc1, float* dgates, float* dm_out, cudaStream_t st);
// t = index of this update (1 for the first step after global_step 0)
int launch_adam(float* theta, const float* grad, float* m, float* v, long long n, long long t, double lr, double b1,
                double b2, double eps, float gscale, float l2, cudaStream_t st);
int launch_l2_term(const float* theta, long long n, float l2, float* sc, cudaStream_t st);
long long lstm_gates_partial_floats(int N, int K, int Ncols);
// d(state)[N][ld] (first `rows` columns) = dgates[N][ncols] x W[rows][ncols]^T
int launch_lstm_dstate(const float* dgates, const float* W, float* dstate, int N, int rows, int ncols, int ld,
                       cudaStream_t st);
// up to four device-to-device output copies in one launch (dst[i] == NULL skips one)
int launch_copy_outputs(const float* const* src, float* const* dst, const long long* n, cudaStream_t st);
// slices_out != NULL: only the split-K partials are produced (the caller's next kernel reduces them)
int launch_lstm_gates(const float* x, int ldx, const float* W, const float* bias, float* gates, float* partial, int N,
                      int K, int Ncols, cudaStream_t st, int* slices_out = nullptr);
int launch_ring_push(void* ring, const void* frame, const unsigned char* fresh, int N, int K, long long row_bytes,
                     int slot, cudaStream_t st);

// lstm_persistent.cu: the T-step recurrence / its back-propagation through time as one cluster launch each, W_h resident
// in shared memory.  gates holds x_t @ W_x + bias on entry of the forward and the full pre-activations on exit.
bool lstm_persistent_supported(int Hl);
int launch_lstm_seq_fwd(int T, int N, int Hl, int xdim, const float* kernel, const float* c0, const float* m0,
                        const unsigned char* reset, float* gates, float* c, float* m, float* states, float* state_out,
                        cudaStream_t st);
int launch_lstm_seq_bwd(int T, int N, int Hl, int xdim, const float* kernel, const float* c0, const unsigned char* reset,
                        const float* gates, const float* c, const float* dm_last, float* dgates, cudaStream_t st);
