// Everything after the conv encoders: state concatenation, LSTM cell, FC + heads, the four loss
// terms with their gradients, the backward of that tail, and the fused TF-style Adam update.
//
// Reference: src/models/e2evmc/graph.py:169-192 (representation_concatenation_v2), :198-260
// (lstm_decoder), :452-500 (losses); src/models/e2evmc/estimator.py:205-244 (targets, loss
// composition, AdamOptimizer).  TF-1.15 semantics restated in oracle/geeco_oracle.py.
#include "tail.cuh"

// ---------------------------------------------------------------------------------------
// state[n] = flatten_hwc(concat_c[obs, dyn, jnt, tgt]) ++ m_prev          (graph.py:187-190)
// feature index = cell * per + block_offset + c,  cell = h*2 + w
// ---------------------------------------------------------------------------------------
__global__ void build_state_kernel(TailDims d, const float* __restrict__ y_obs, const float* __restrict__ y_dyn,
                                   const float* __restrict__ y_tgt, const float* __restrict__ jnt_states,
                                   const float* __restrict__ m_prev, float* __restrict__ state) {
  const int n = blockIdx.x;
  const int per = d.D_obs + d.D_dyn + d.J + d.D_diff;
  const int xdim = 4 * per;
  float* row = state + (long long)n * (xdim + d.Hl);
  const float* jn = jnt_states + ((long long)n * d.K + (d.K - 1)) * d.J;   // last frame of the window, graph.py:388
  for (int i = threadIdx.x; i < xdim + d.Hl; i += blockDim.x) {
    float v;
    if (i >= xdim) {
      v = m_prev ? m_prev[(long long)n * d.Hl + (i - xdim)] : 0.f;
    } else {
      const int cell = i / per, c = i - cell * per;
      if (c < d.D_obs) v = y_obs[((long long)n * 4 + cell) * d.D_obs + c];
      else if (c < d.D_obs + d.D_dyn) v = y_dyn[((long long)n * 4 + cell) * d.D_dyn + (c - d.D_obs)];
      else if (c < d.D_obs + d.D_dyn + d.J) v = jn[c - d.D_obs - d.D_dyn];
      else v = y_tgt[((long long)n * 4 + cell) * d.D_diff + (c - d.D_obs - d.D_dyn - d.J)];
    }
    row[i] = v;
  }
}

// dY8(pre-activation)[enc][n][cell][c] = dstate[n][cell*per + off_enc + c] * (Y8 > 0)
__global__ void scatter_dstate_kernel(TailDims d, const float* __restrict__ dstate, int ld_dstate,
                                      const float* __restrict__ y_obs, const float* __restrict__ y_dyn,
                                      const float* __restrict__ y_tgt, float* __restrict__ g_obs,
                                      float* __restrict__ g_dyn, float* __restrict__ g_tgt) {
  const int n = blockIdx.x;
  const int per = d.D_obs + d.D_dyn + d.J + d.D_diff;
  const float* row = dstate + (long long)n * ld_dstate;
  for (int i = threadIdx.x; i < 4 * per; i += blockDim.x) {
    const int cell = i / per, c = i - cell * per;
    const float v = row[i];
    if (c < d.D_obs) {
      const long long o = ((long long)n * 4 + cell) * d.D_obs + c;
      g_obs[o] = y_obs[o] > 0.f ? v : 0.f;
    } else if (c < d.D_obs + d.D_dyn) {
      const long long o = ((long long)n * 4 + cell) * d.D_dyn + (c - d.D_obs);
      g_dyn[o] = y_dyn[o] > 0.f ? v : 0.f;
    } else if (c >= d.D_obs + d.D_dyn + d.J) {
      const long long o = ((long long)n * 4 + cell) * d.D_diff + (c - d.D_obs - d.D_dyn - d.J);
      g_tgt[o] = y_tgt[o] > 0.f ? v : 0.f;
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// LSTMCell (state_is_tuple=False, forget_bias=1): gates already hold [x,m]W + b.
__global__ void lstm_cell_kernel(int N, int Hl, const float* __restrict__ gates, const float* __restrict__ c_prev,
                                 float* __restrict__ c_out, float* __restrict__ m_out, float* __restrict__ state_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Hl) return;
  const int n = idx / Hl, i = idx - n * Hl;
  const float* gr = gates + (long long)n * 4 * Hl;
  const float gi = gr[i], gj = gr[Hl + i], gf = gr[2 * Hl + i], go = gr[3 * Hl + i];
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const float c = sigmoidf_(gf + 1.0f) * cp + sigmoidf_(gi) * tanhf(gj);
  const float m = sigmoidf_(go) * tanhf(c);
  c_out[idx] = c;
  m_out[idx] = m;
  if (state_out) {
    state_out[(long long)n * 2 * Hl + i] = c;
    state_out[(long long)n * 2 * Hl + Hl + i] = m;
  }
}

// ---------------------------------------------------------------------------------------
// fc1 + heads + per-sample loss pieces + d(loss)/d(head outputs).  One CTA per sample.
// head columns: [cmd_ee 0:3 | logits_cmd_grp 3:3+G | aux_ee | aux_obj]
// loss_parts[n] = {se_cmd_ee, ce_cmd_grp, se_pos_ee, se_pos_obj, correct}
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tail_fwd_kernel(TailDims d, TailParams p, const float* __restrict__ m,
                                                       float* __restrict__ fc1, float* __restrict__ heads,
                                                       const float* __restrict__ cmd, const float* __restrict__ ee_state,
                                                       const float* __restrict__ obj_state,
                                                       float* __restrict__ loss_parts, float* __restrict__ dheads,
                                                       int with_loss) {
  extern __shared__ float sm[];
  float* m_s = sm;                 // Hl
  float* fc_s = sm + d.Hl;         // Fc
  float* out_s = fc_s + d.Fc;      // NH
  const int n = blockIdx.x;
  const int NH = 9 + d.G;
  for (int i = threadIdx.x; i < d.Hl; i += blockDim.x) m_s[i] = m[(long long)n * d.Hl + i];
  __syncthreads();
  for (int j = threadIdx.x; j < d.Fc; j += blockDim.x) {
    float s = p.b_fc1[j];
#pragma unroll 16
    for (int i = 0; i < d.Hl; ++i) s = fmaf(m_s[i], __ldg(p.w_fc1 + (long long)i * d.Fc + j), s);   // loads issue ahead of the FMA chain
    s = fmaxf(s, 0.f);
    fc_s[j] = s;
    fc1[(long long)n * d.Fc + j] = s;
  }
  __syncthreads();
  // heads: one warp per output column (lanes stride over fc1, fixed-order lane reduction), 4 warps take turns
  for (int t = threadIdx.x >> 5; t < NH; t += blockDim.x >> 5) {
    const float* w; const float* b; int col, width;
    if (t < 3) { w = p.w_cmd_ee; b = p.b_cmd_ee; col = t; width = 3; }
    else if (t < 3 + d.G) { w = p.w_grp; b = p.b_grp; col = t - 3; width = d.G; }
    else if (t < 6 + d.G) { w = p.w_aux_ee; b = p.b_aux_ee; col = t - 3 - d.G; width = 3; }
    else { w = p.w_aux_obj; b = p.b_aux_obj; col = t - 6 - d.G; width = 3; }
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int j = lane; j < d.Fc; j += 32) s = fmaf(fc_s[j], __ldg(w + j * width + col), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += b[col];
    if (lane == 0) { out_s[t] = s; heads[(long long)n * NH + t] = s; }
  }
  __syncthreads();
  if (with_loss && threadIdx.x == 0) {
    const float* c = cmd + (long long)n * 4;
    const float* ee = ee_state + ((long long)n * d.K + (d.K - 1)) * 7;
    const float* ob = obj_state + ((long long)n * d.K + (d.K - 1)) * 7;
    float* dh = dheads + (long long)n * NH;
    const float inv = 1.f / (3.f * d.N);
    float se0 = 0.f, se2 = 0.f, se3 = 0.f;
    for (int k = 0; k < 3; ++k) {
      const float e0 = out_s[k] - c[k];
      const float e2 = out_s[3 + d.G + k] - ee[k];
      const float e3 = out_s[6 + d.G + k] - ob[k];
      se0 += e0 * e0; se2 += e2 * e2; se3 += e3 * e3;
      dh[k] = 2.f * e0 * inv;
      dh[3 + d.G + k] = d.lambda_aux * 2.f * e2 * inv;
      dh[6 + d.G + k] = d.lambda_aux * 2.f * e3 * inv;
    }
    // estimator.py:213-216: class = int32(rint(cmd[:,3])) + 1 ; one_hot of an out-of-range index is all zeros
    const int cls = (int)rintf(c[3]) + 1;
    float zmax = -3.4e38f; int amax = 0;
    for (int k = 0; k < d.G; ++k) if (out_s[3 + k] > zmax) { zmax = out_s[3 + k]; amax = k; }
    float se = 0.f;
    for (int k = 0; k < d.G; ++k) se += expf(out_s[3 + k] - zmax);
    const float lse = zmax + logf(se);
    const bool ok = cls >= 0 && cls < d.G;
    for (int k = 0; k < d.G; ++k) {
      const float pk = expf(out_s[3 + k] - lse);
      dh[3 + k] = ok ? (pk - (k == cls ? 1.f : 0.f)) / d.N : 0.f;
    }
    float* lp = loss_parts + (long long)n * 5;
    lp[0] = se0; lp[1] = ok ? lse - out_s[3 + cls] : 0.f; lp[2] = se2; lp[3] = se3;
    lp[4] = (amax == cls) ? 1.f : 0.f;
  }
}

// losses[8] = {loss_cmd_ee, loss_cmd_grp, loss_pos_ee, loss_pos_obj, loss_reg, loss, sum_correct, N}
__global__ void loss_reduce_kernel(int N, float lambda_aux, const float* __restrict__ loss_parts,
                                   const float* __restrict__ reg_term, float* __restrict__ losses) {
  __shared__ float s[5][256];
  float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n = threadIdx.x; n < N; n += blockDim.x)
    for (int k = 0; k < 5; ++k) a[k] += loss_parts[(long long)n * 5 + k];
  for (int k = 0; k < 5; ++k) s[k][threadIdx.x] = a[k];
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) for (int k = 0; k < 5; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l0 = s[0][0] / (3.f * N), l1 = s[1][0] / N, l2 = s[2][0] / (3.f * N), l3 = s[3][0] / (3.f * N);
    const float lr = reg_term ? reg_term[0] : 0.f;
    losses[0] = l0; losses[1] = l1; losses[2] = l2; losses[3] = l3; losses[4] = lr;
    losses[5] = (l0 + l1) + lambda_aux * (l2 + l3) + lr;
    losses[6] = s[4][0]; losses[7] = (float)N;
  }
}

// backward through heads, fc1 and the LSTM cell -> d(gates).  One CTA per sample.
__global__ void __launch_bounds__(128) tail_bwd_kernel(TailDims d, TailParams p, const float* __restrict__ fc1,
                                                       const float* __restrict__ dheads, const float* __restrict__ gates,
                                                       const float* __restrict__ c_prev, float* __restrict__ dfc1,
                                                       float* __restrict__ dgates) {
  extern __shared__ float sm[];
  float* dh_s = sm;                // NH (padded to 16)
  float* dfc_s = sm + 16;          // Fc
  const int n = blockIdx.x;
  const int NH = 9 + d.G;
  if (threadIdx.x < NH) dh_s[threadIdx.x] = dheads[(long long)n * NH + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.x; j < d.Fc; j += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < 3; ++k) s = fmaf(dh_s[k], p.w_cmd_ee[j * 3 + k], s);
    for (int k = 0; k < d.G; ++k) s = fmaf(dh_s[3 + k], p.w_grp[j * d.G + k], s);
    for (int k = 0; k < 3; ++k) s = fmaf(dh_s[3 + d.G + k], p.w_aux_ee[j * 3 + k], s);
    for (int k = 0; k < 3; ++k) s = fmaf(dh_s[6 + d.G + k], p.w_aux_obj[j * 3 + k], s);
    s = fc1[(long long)n * d.Fc + j] > 0.f ? s : 0.f;
    dfc_s[j] = s;
    dfc1[(long long)n * d.Fc + j] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.Hl; i += blockDim.x) {
    float dm = 0.f;
    const float* wr = p.w_fc1 + (long long)i * d.Fc;
    for (int j = 0; j < d.Fc; ++j) dm = fmaf(dfc_s[j], wr[j], dm);
    const float* gr = gates + (long long)n * 4 * d.Hl;
    const float gi = gr[i], gj = gr[d.Hl + i], gf = gr[2 * d.Hl + i], go = gr[3 * d.Hl + i];
    const float cp = c_prev ? c_prev[(long long)n * d.Hl + i] : 0.f;
    const float si = sigmoidf_(gi), tj = tanhf(gj), sf = sigmoidf_(gf + 1.0f), so = sigmoidf_(go);
    const float c = sf * cp + si * tj;
    const float tc = tanhf(c);
    const float dso = dm * tc;
    const float dc = dm * so * (1.f - tc * tc);
    float* dg = dgates + (long long)n * 4 * d.Hl;
    dg[i] = dc * tj * si * (1.f - si);
    dg[d.Hl + i] = dc * si * (1.f - tj * tj);
    dg[2 * d.Hl + i] = dc * cp * sf * (1.f - sf);
    dg[3 * d.Hl + i] = dso * so * (1.f - so);
  }
}

// weight / bias gradients of fc1 and the heads: one thread per element, fixed-order sum over n
__global__ void tail_wgrad_kernel(TailDims d, TailGrads g, const float* __restrict__ m, const float* __restrict__ fc1,
                                  const float* __restrict__ dfc1, const float* __restrict__ dheads) {
  const int NH = 9 + d.G;
  const int n_w1 = d.Hl * d.Fc, n_b1 = d.Fc, n_wh = d.Fc * NH, n_bh = NH;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_w1) {
    const int i = idx / d.Fc, j = idx - i * d.Fc;
    float s = 0.f;
    for (int n = 0; n < d.N; ++n) s = fmaf(m[(long long)n * d.Hl + i], dfc1[(long long)n * d.Fc + j], s);
    g.w_fc1[idx] = s;
  } else if (idx < n_w1 + n_b1) {
    const int j = idx - n_w1;
    float s = 0.f;
    for (int n = 0; n < d.N; ++n) s += dfc1[(long long)n * d.Fc + j];
    g.b_fc1[j] = s;
  } else if (idx < n_w1 + n_b1 + n_wh + n_bh) {
    const int e = idx - n_w1 - n_b1;
    const bool is_bias = e >= n_wh;
    const int j = is_bias ? 0 : e / NH, t = is_bias ? e - n_wh : e - (e / NH) * NH;
    float s = 0.f;
    if (is_bias) for (int n = 0; n < d.N; ++n) s += dheads[(long long)n * NH + t];
    else for (int n = 0; n < d.N; ++n) s = fmaf(fc1[(long long)n * d.Fc + j], dheads[(long long)n * NH + t], s);
    float* w; float* b; int col, width;
    if (t < 3) { w = g.w_cmd_ee; b = g.b_cmd_ee; col = t; width = 3; }
    else if (t < 3 + d.G) { w = g.w_grp; b = g.b_grp; col = t - 3; width = d.G; }
    else if (t < 6 + d.G) { w = g.w_aux_ee; b = g.b_aux_ee; col = t - 3 - d.G; width = 3; }
    else { w = g.w_aux_obj; b = g.b_aux_obj; col = t - 6 - d.G; width = 3; }
    if (is_bias) b[col] = s; else w[j * width + col] = s;
  }
}

// ---------------------------------------------------------------------------------------
// TF Adam (training_ops.apply_adam): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1);
// v += (g*g-v)(1-b2); theta -= lr_t*m/(sqrt(v)+eps).   One launch over the flat arena.
// sc[0] = t (as float), sc[1] = lr_t, sc[2] = 0.5*l2*sum(theta^2)
// ---------------------------------------------------------------------------------------
__global__ void adam_prep_kernel(float* sc, double lr, double b1, double b2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double t = (double)sc[0] + 1.0;
    sc[0] = (float)t;
    sc[1] = (float)(lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t)));
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ theta, const float4* __restrict__ grad,
                                                   float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                   const float* __restrict__ sc, float b1, float b2, float eps,
                                                   float gscale, float l2) {
  const float lr_t = sc[1];
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 th = theta[i], g = grad[i], mm = m[i], vv = v[i];
#define ADAM1(c)                                              \
    {                                                         \
      const float gg = g.c * gscale + l2 * th.c;              \
      mm.c += (gg - mm.c) * omb1;                             \
      vv.c += (gg * gg - vv.c) * omb2;                        \
      th.c -= lr_t * mm.c / (sqrtf(vv.c) + eps);              \
    }
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    theta[i] = th; m[i] = mm; v[i] = vv;
  }
}

// 0.5 * l2 * sum(theta^2) -> sc[2]  (single block, fixed order; only launched when l2 > 0)
__global__ void l2_term_kernel(const float* __restrict__ theta, long long n, float l2, float* sc) {
  __shared__ double s[1024];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) a += (double)theta[i] * theta[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sc[2] = (float)(0.5 * l2 * s[0]);
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
int launch_build_state(const TailDims& d, const float* y_obs, const float* y_dyn, const float* y_tgt, const float* jnt,
                       const float* m_prev, float* state, cudaStream_t st) {
  build_state_kernel<<<d.N, 256, 0, st>>>(d, y_obs, y_dyn, y_tgt, jnt, m_prev, state);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_scatter_dstate(const TailDims& d, const float* dstate, int ld, const float* y_obs, const float* y_dyn,
                          const float* y_tgt, float* g_obs, float* g_dyn, float* g_tgt, cudaStream_t st) {
  scatter_dstate_kernel<<<d.N, 256, 0, st>>>(d, dstate, ld, y_obs, y_dyn, y_tgt, g_obs, g_dyn, g_tgt);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_lstm_cell(int N, int Hl, const float* gates, const float* c_prev, float* c_out, float* m_out,
                     float* state_out, cudaStream_t st) {
  lstm_cell_kernel<<<ceil_div((long long)N * Hl, 256), 256, 0, st>>>(N, Hl, gates, c_prev, c_out, m_out, state_out);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_tail_fwd(const TailDims& d, const TailParams& p, const float* m, float* fc1, float* heads, const float* cmd,
                    const float* ee, const float* obj, float* loss_parts, float* dheads, int with_loss,
                    cudaStream_t st) {
  const size_t smem = (size_t)(d.Hl + d.Fc + 16 + d.G) * sizeof(float);
  tail_fwd_kernel<<<d.N, 128, smem, st>>>(d, p, m, fc1, heads, cmd, ee, obj, loss_parts, dheads, with_loss);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_loss_reduce(const TailDims& d, const float* loss_parts, const float* reg_term, float* losses,
                       cudaStream_t st) {
  loss_reduce_kernel<<<1, 256, 0, st>>>(d.N, d.lambda_aux, loss_parts, reg_term, losses);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_tail_bwd(const TailDims& d, const TailParams& p, const TailGrads& g, const float* m, const float* fc1,
                    const float* dheads, const float* gates, const float* c_prev, float* dfc1, float* dgates,
                    cudaStream_t st) {
  const size_t smem = (size_t)(16 + d.Fc) * sizeof(float);
  tail_bwd_kernel<<<d.N, 128, smem, st>>>(d, p, fc1, dheads, gates, c_prev, dfc1, dgates);
  const int NH = 9 + d.G;
  const int total = d.Hl * d.Fc + d.Fc + d.Fc * NH + NH;
  tail_wgrad_kernel<<<ceil_div(total, 128), 128, 0, st>>>(d, g, m, fc1, dfc1, dheads);
  geeco_count_launch(2);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_adam(float* theta, const float* grad, float* m, float* v, long long n, float* sc, double lr, double b1,
                double b2, double eps, float gscale, float l2, cudaStream_t st) {
  adam_prep_kernel<<<1, 32, 0, st>>>(sc, lr, b1, b2);
  const long long n4 = n / 4;
  int blocks = ceil_div(n4, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<float4*>(theta), reinterpret_cast<const float4*>(grad),
                                      reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4, sc, (float)b1,
                                      (float)b2, (float)eps, gscale, l2);
  geeco_count_launch(2);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_l2_term(const float* theta, long long n, float l2, float* sc, cudaStream_t st) {
  l2_term_kernel<<<1, 1024, 0, st>>>(theta, n, l2, sc);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

// ---------------------------------------------------------------------------------------
// LSTM gate GEMM  gates[N, 4h] = [x, m_prev][N, K] @ kernel[K, 4h] + bias   (graph.py:224)
// M = batch is tiny, K = 3228 is long: split K across CTAs (each streams 32 kernel rows once),
// fixed-order reduction of the partials adds the bias.  fp32 CUDA cores; 0.2 GFLOP.
// ---------------------------------------------------------------------------------------
constexpr int GK_SLICE = 32;
template <int CPT>
__global__ void __launch_bounds__(256) gates_splitk_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           float* __restrict__ partial, int N, int K, int ldx,
                                                           int Ncols) {
  __shared__ float xs[32][GK_SLICE + 1];
  const int k0 = blockIdx.x * GK_SLICE, n0 = blockIdx.y * 32;
  for (int e = threadIdx.x; e < 32 * GK_SLICE; e += 256) {
    const int n = e / GK_SLICE, k = e - n * GK_SLICE;
    xs[n][k] = (n0 + n < N && k0 + k < K) ? x[(long long)(n0 + n) * ldx + k0 + k] : 0.f;
  }
  __syncthreads();
  float acc[32][CPT];
#pragma unroll
  for (int n = 0; n < 32; ++n)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[n][c] = 0.f;
  const int kend = (K - k0) < GK_SLICE ? (K - k0) : GK_SLICE;
  for (int kk = 0; kk < kend; ++kk) {
    float w[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = threadIdx.x + c * 256;
      w[c] = col < Ncols ? __ldg(W + (long long)(k0 + kk) * Ncols + col) : 0.f;
    }
#pragma unroll
    for (int n = 0; n < 32; ++n) {
      const float a = xs[n][kk];
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[n][c] = fmaf(a, w[c], acc[n][c]);
    }
  }
#pragma unroll
  for (int n = 0; n < 32; ++n) {
    if (n0 + n >= N) break;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = threadIdx.x + c * 256;
      if (col < Ncols) partial[((long long)blockIdx.x * N + n0 + n) * Ncols + col] = acc[n][c];
    }
  }
}
__global__ void gates_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                    float* __restrict__ gates, int slices, int N, int Ncols) {
  const long long total = (long long)N * Ncols;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s = bias ? bias[i % Ncols] : 0.f;
  for (int k = 0; k < slices; ++k) s += partial[(long long)k * total + i];
  gates[i] = s;
}

// the caller-visible outputs of a step (heads, fc1, LSTM state, losses) in ONE launch instead of four D2D copies
struct CopyList { const float* src[4]; float* dst[4]; long long n[4]; };
__global__ void copy_outputs_kernel(CopyList cl) {
  const float* s = cl.src[blockIdx.y];
  float* d = cl.dst[blockIdx.y];
  if (!d) return;
  const long long n = cl.n[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
}
int launch_copy_outputs(const float* const* src, float* const* dst, const long long* n, cudaStream_t st) {
  CopyList cl;
  bool any = false;
  for (int i = 0; i < 4; ++i) { cl.src[i] = src[i]; cl.dst[i] = dst[i]; cl.n[i] = n[i]; any = any || dst[i]; }
  if (!any) return GEECO_OK;
  copy_outputs_kernel<<<dim3(16, 4), 256, 0, st>>>(cl);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

// ---------------------------------------------------------------------------------------
// d(state)[n][k] = sum_j dgates[n][j] * kernel[k][j] for the first `xdim` kernel rows (the x part of [x, m_prev]):
// the backward of the gate GEMM with respect to its input.  One block = 32 kernel rows x up to 64 samples; the
// reduction runs over 64-column chunks staged in shared memory (kernel rows padded against bank conflicts, the
// dgates values are warp-wide broadcasts).  0.2 GFLOP; the generic gather GEMM took 50 us with 49 CTAs.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_dstate_kernel(const float* __restrict__ dgates, const float* __restrict__ W,
                                                          float* __restrict__ dstate, int N, int xdim, int NC, int ld) {
  __shared__ float Ws[32][65];
  __shared__ float Ds[64][65];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 64;
  const int k = threadIdx.x & 31, ng = threadIdx.x >> 5;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < NC; j0 += 64) {
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      Ws[r][c] = (k0 + r < xdim && j0 + c < NC) ? __ldg(W + (long long)(k0 + r) * NC + j0 + c) : 0.f;
    }
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      Ds[r][c] = (n0 + r < N && j0 + c < NC) ? __ldg(dgates + (long long)(n0 + r) * NC + j0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < 64; ++jj) {
      const float w = Ws[k][jj];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(Ds[ng * 8 + i][jj], w, acc[i]);
    }
    __syncthreads();
  }
  if (k0 + k < xdim) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + ng * 8 + i;
      if (n < N) dstate[(long long)n * ld + k0 + k] = acc[i];
    }
  }
}
int launch_lstm_dstate(const float* dgates, const float* W, float* dstate, int N, int xdim, int ncols, int ld,
                       cudaStream_t st) {
  lstm_dstate_kernel<<<dim3((xdim + 31) / 32, (N + 63) / 64), 256, 0, st>>>(dgates, W, dstate, N, xdim, ncols, ld);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

long long lstm_gates_partial_floats(int N, int K, int Ncols) {
  return (long long)((K + GK_SLICE - 1) / GK_SLICE) * N * Ncols;
}
int launch_lstm_gates(const float* x, int ldx, const float* W, const float* bias, float* gates, float* partial, int N,
                      int K, int Ncols, cudaStream_t st) {
  if (Ncols > 1024) { geeco_set_error("lstm gates: 4*dim_h_lstm = %d > 1024 not supported", Ncols); return GEECO_ERR_INVALID; }
  const int slices = (K + GK_SLICE - 1) / GK_SLICE;
  dim3 grid(slices, (N + 31) / 32);
  const int cpt = (Ncols + 255) / 256;
  if (cpt == 1) gates_splitk_kernel<1><<<grid, 256, 0, st>>>(x, W, partial, N, K, ldx, Ncols);
  else if (cpt == 2) gates_splitk_kernel<2><<<grid, 256, 0, st>>>(x, W, partial, N, K, ldx, Ncols);
  else if (cpt == 3) gates_splitk_kernel<3><<<grid, 256, 0, st>>>(x, W, partial, N, K, ldx, Ncols);
  else gates_splitk_kernel<4><<<grid, 256, 0, st>>>(x, W, partial, N, K, ldx, Ncols);
  gates_reduce_kernel<<<ceil_div((long long)N * Ncols, 256), 256, 0, st>>>(partial, bias, gates, slices, N, Ncols);
  geeco_count_launch(2);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
