// K-step LSTM recurrence of `lstm_decoder` (graph.py:212-225) as a stand-alone op, forward and back-propagation
// through time: the part of the `--proc_obs sequence` / `--goal_condition none` graphs that the one-step GEECO-F tail
// does not have.  fp32; built from the tail's own kernels (split-K gate GEMM, cell, d(state)) plus the cell backward
// with an incoming d(c) below, so that a K = 1 sequence reproduces the step's LSTM bit for bit.
//
//   state_t = [x_t | m_{t-1}]                     (m_{-1} = c_{-1} = 0: the reference never carries the state, graph.py:226)
//   gates_t = state_t @ kernel + bias ; i, j, f, o = split(gates_t)
//   c_t = sigmoid(f + 1) * c_{t-1} + sigmoid(i) * tanh(j) ;  m_t = sigmoid(o) * tanh(c_t)
//
// Scratch layout (floats): states [K][N][xdim+Hl] | dgates [K][N][4Hl] | dstate [N][xdim+Hl] | dc [N][Hl] | partial.
#include "common.cuh"
#include "tail.cuh"
#include "plan.cuh"
#include "../../include/geeco_b200.h"

struct SeqScratch {
  float *states, *dgates, *dstate, *dc, *partial;
  long long partial_cap, total;
};

static SeqScratch carve(float* base, int N, int K, int xdim, int Hl) {
  SeqScratch s;
  const long long ld = xdim + Hl;
  auto up = [](long long v) { return (v + 63) & ~63ll; };          // keep every piece 256-byte aligned
  long long off = 0;
  s.states = base ? base + off : nullptr; off += up((long long)K * N * ld);
  s.dgates = base ? base + off : nullptr; off += up((long long)K * N * 4 * Hl);
  s.dstate = base ? base + off : nullptr; off += up((long long)N * ld);
  s.dc = base ? base + off : nullptr;     off += up((long long)N * Hl);
  GatherGeom gw = dense_geom(K * N, (int)ld, 4 * Hl, 4 * Hl, 0);
  const long long p_gates = lstm_gates_partial_floats(N, (int)ld, 4 * Hl);
  const long long p_wgrad = gemm_tn_partial_floats(gw, 1);
  s.partial_cap = p_gates > p_wgrad ? p_gates : p_wgrad;
  s.partial = base ? base + off : nullptr; off += up(s.partial_cap);
  s.total = off;
  return s;
}

static int check_dims(int N, int K, int xdim, int Hl) {
  if (N < 1 || K < 1 || xdim < 4 || (xdim & 3) || Hl < 4 || (Hl & 3) || 4 * Hl > 1024) {
    geeco_set_error("lstm_seq: need N, K >= 1, xdim a multiple of 4 and dim_h_lstm a multiple of 4 up to 256 (got N=%d K=%d xdim=%d Hl=%d)",
                    N, K, xdim, Hl);
    return GEECO_ERR_INVALID;
  }
  return GEECO_OK;
}

// states[t] = [x_t | m_{t-1}] for every step (m_{-1} = 0)
static int build_states(const SeqScratch& s, const float* x, const float* m, int N, int K, int xdim, int Hl, cudaStream_t st) {
  const size_t ld = (size_t)(xdim + Hl) * sizeof(float);
  CUDA_TRY(cudaMemcpy2DAsync(s.states, ld, x, (size_t)xdim * sizeof(float), (size_t)xdim * sizeof(float), (size_t)K * N,
                             cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemset2DAsync(s.states + xdim, ld, 0, (size_t)Hl * sizeof(float), (size_t)N, st));
  if (K > 1 && m)
    CUDA_TRY(cudaMemcpy2DAsync(s.states + (long long)N * (xdim + Hl) + xdim, ld, m, (size_t)Hl * sizeof(float),
                               (size_t)Hl * sizeof(float), (size_t)(K - 1) * N, cudaMemcpyDeviceToDevice, st));
  return GEECO_OK;
}

extern "C" int64_t geeco_lstm_seq_scratch_floats(int32_t N, int32_t K, int32_t xdim, int32_t Hl) {
  if (check_dims(N, K, xdim, Hl)) return -1;
  return carve(nullptr, N, K, xdim, Hl).total;
}

extern "C" int geeco_lstm_seq_fwd(const float* x, const float* kernel, const float* bias, float* gates, float* c, float* m,
                                  float* scratch, int64_t scratch_floats, int32_t N, int32_t K, int32_t xdim, int32_t Hl,
                                  void* stream) {
  int rc = check_dims(N, K, xdim, Hl);
  if (rc) return rc;
  if (!x || !kernel || !bias || !gates || !c || !m || !scratch) { geeco_set_error("lstm_seq_fwd: NULL tensor"); return GEECO_ERR_INVALID; }
  if (((uintptr_t)scratch) & 255) { geeco_set_error("lstm_seq_fwd: scratch must be 256-byte aligned"); return GEECO_ERR_INVALID; }
  SeqScratch s = carve(scratch, N, K, xdim, Hl);
  if (scratch_floats < s.total) { geeco_set_error("lstm_seq_fwd: scratch of %lld floats < required %lld", (long long)scratch_floats, s.total); return GEECO_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int ld = xdim + Hl;
  rc = build_states(s, x, nullptr, N, K, xdim, Hl, st);
  if (rc) return rc;
  for (int t = 0; t < K; ++t) {
    float* state_t = s.states + (long long)t * N * ld;
    if (t > 0)    // m_{t-1} behind x_t
      CUDA_TRY(cudaMemcpy2DAsync(state_t + xdim, (size_t)ld * sizeof(float), m + (long long)(t - 1) * N * Hl,
                                 (size_t)Hl * sizeof(float), (size_t)Hl * sizeof(float), (size_t)N, cudaMemcpyDeviceToDevice, st));
    float* gates_t = gates + (long long)t * N * 4 * Hl;
    rc = launch_lstm_gates(state_t, ld, kernel, bias, gates_t, s.partial, N, ld, 4 * Hl, st);
    if (rc) return rc;
    rc = launch_lstm_cell(N, Hl, gates_t, t ? c + (long long)(t - 1) * N * Hl : nullptr, nullptr, c + (long long)t * N * Hl,
                          m + (long long)t * N * Hl, nullptr, nullptr, 0, st);
    if (rc) return rc;
  }
  return GEECO_OK;
}

extern "C" int geeco_lstm_seq_bwd(const float* x, const float* kernel, const float* gates, const float* c, const float* m,
                                  const float* dm_last, float* dkernel, float* dbias, float* dx, float* scratch,
                                  int64_t scratch_floats, int32_t N, int32_t K, int32_t xdim, int32_t Hl, void* stream) {
  int rc = check_dims(N, K, xdim, Hl);
  if (rc) return rc;
  if (!x || !kernel || !gates || !c || !m || !dm_last || !dkernel || !dbias || !scratch) {
    geeco_set_error("lstm_seq_bwd: NULL tensor");
    return GEECO_ERR_INVALID;
  }
  if (((uintptr_t)scratch) & 255) { geeco_set_error("lstm_seq_bwd: scratch must be 256-byte aligned"); return GEECO_ERR_INVALID; }
  SeqScratch s = carve(scratch, N, K, xdim, Hl);
  if (scratch_floats < s.total) { geeco_set_error("lstm_seq_bwd: scratch of %lld floats < required %lld", (long long)scratch_floats, s.total); return GEECO_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int ld = xdim + Hl;
  rc = build_states(s, x, m, N, K, xdim, Hl, st);
  if (rc) return rc;
  for (int t = K - 1; t >= 0; --t) {
    const bool last = t == K - 1;
    float* dgates_t = s.dgates + (long long)t * N * 4 * Hl;
    // d(m_t): from the decoder for the last step, else the m part of d(state_{t+1}) left in s.dstate
    rc = launch_lstm_cell_bwd(N, Hl, gates + (long long)t * N * 4 * Hl, t ? c + (long long)(t - 1) * N * Hl : nullptr, nullptr,
                              last ? dm_last : s.dstate + xdim, last ? Hl : ld, last ? nullptr : s.dc, dgates_t, s.dc, st);
    if (rc) return rc;
    if (t > 0 || dx) {
      // d(state_t) = d(gates_t) @ kernel^T over all xdim + Hl kernel rows (x part -> dx_t, m part -> step t-1)
      rc = launch_lstm_dstate(dgates_t, kernel, s.dstate, N, ld, 4 * Hl, ld, st);
      if (rc) return rc;
      if (dx)
        CUDA_TRY(cudaMemcpy2DAsync(dx + (long long)t * N * xdim, (size_t)xdim * sizeof(float), s.dstate, (size_t)ld * sizeof(float),
                                   (size_t)xdim * sizeof(float), (size_t)N, cudaMemcpyDeviceToDevice, st));
    }
  }
  // d(kernel) = sum_t state_t^T @ d(gates_t), d(bias) = column sums: one deterministic split GEMM over K*N rows
  GatherGeom gw = dense_geom(K * N, ld, 4 * Hl, 4 * Hl, 0);
  return launch_gemm_tn_f32(gw, s.states, s.dgates, dkernel, dbias, s.partial, s.partial_cap, 1, 0, 0, st);
}
