#pragma once
#include "common.cuh"

struct TailDims {
  int N, K, J;                 // batch, window size, dim_jnt_state
  int D_obs, D_dyn, D_diff;    // conv8 widths of the three encoders
  int Hl, Fc, G;               // dim_h_lstm, dim_h_fc, num_grp_states
  float lambda_aux;
};
struct TailParams {
  const float *w_fc1, *b_fc1, *w_cmd_ee, *b_cmd_ee, *w_grp, *b_grp, *w_aux_ee, *b_aux_ee, *w_aux_obj, *b_aux_obj;
};
struct TailGrads {
  float *w_fc1, *b_fc1, *w_cmd_ee, *b_cmd_ee, *w_grp, *b_grp, *w_aux_ee, *b_aux_ee, *w_aux_obj, *b_aux_obj;
};

int launch_build_state(const TailDims& d, const float* y_obs, const float* y_dyn, const float* y_tgt, const float* jnt,
                       const float* m_prev, float* state, cudaStream_t st);
int launch_scatter_dstate(const TailDims& d, const float* dstate, int ld, const float* y_obs, const float* y_dyn,
                          const float* y_tgt, float* g_obs, float* g_dyn, float* g_tgt, cudaStream_t st);
int launch_lstm_cell(int N, int Hl, const float* gates, const float* c_prev, float* c_out, float* m_out,
                     float* state_out, cudaStream_t st);
int launch_tail_fwd(const TailDims& d, const TailParams& p, const float* m, float* fc1, float* heads, const float* cmd,
                    const float* ee, const float* obj, float* loss_parts, float* dheads, int with_loss, cudaStream_t st);
int launch_loss_reduce(const TailDims& d, const float* loss_parts, const float* reg_term, float* losses, cudaStream_t st);
int launch_tail_bwd(const TailDims& d, const TailParams& p, const TailGrads& g, const float* m, const float* fc1,
                    const float* dheads, const float* gates, const float* c_prev, float* dfc1, float* dgates,
                    cudaStream_t st);
int launch_adam(float* theta, const float* grad, float* m, float* v, long long n, float* sc, double lr, double b1,
                double b2, double eps, float gscale, float l2, cudaStream_t st);
int launch_l2_term(const float* theta, long long n, float l2, float* sc, cudaStream_t st);
long long lstm_gates_partial_floats(int N, int K, int Ncols);
// d(state) of the LSTM gate GEMM for the x part of its input: dstate[N][ld] (first xdim columns) = dgates[N][ncols] x W[.][ncols]^T
int launch_lstm_dstate(const float* dgates, const float* W, float* dstate, int N, int xdim, int ncols, int ld,
                       cudaStream_t st);
// up to four device-to-device output copies in one launch (dst[i] == NULL skips one)
int launch_copy_outputs(const float* const* src, float* const* dst, const long long* n, cudaStream_t st);
int launch_lstm_gates(const float* x, int ldx, const float* W, const float* bias, float* gates, float* partial, int N,
                      int K, int Ncols, cudaStream_t st);
