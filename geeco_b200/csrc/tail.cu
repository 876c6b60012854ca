// Everything after the conv encoders: state concatenation, LSTM cell, FC + heads, the four loss
// terms with their gradients, the backward of that tail, and the fused TF-style Adam update.
//
// Reference: src/models/e2evmc/graph.py:169-192 (representation_concatenation_v2), :198-260
// (lstm_decoder), :452-500 (losses); src/models/e2evmc/estimator.py:205-244 (targets, loss
// composition, AdamOptimizer).  TF-1.15 semantics restated in oracle/geeco_oracle.py.
#include "tail.cuh"
#include <math.h>

// ---------------------------------------------------------------------------------------
// LSTM inputs of all T steps:  states[t][n] = [ flatten_hwc(concat_c[...]) | m_{t-1} ]       (graph.py:123-192)
// feature index = cell * per + block_offset + c,  cell = h*2 + w.  The m part is written here for t = 0 only
// (carried m or zeros); the cell kernel of step t writes it for step t+1.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float state_value(const StateMap& sm, const float* __restrict__ jnt, int t, int n, int cell, int c) {
  const int N = sm.N;
  switch (sm.variant) {
    case VAR_GEECOF: {
      if (c < sm.D0) return sm.y[0][((long long)n * 4 + cell) * sm.D0 + c];
      c -= sm.D0;
      if (c < sm.D1) return sm.y[1][((long long)n * 4 + cell) * sm.D1 + c];
      c -= sm.D1;
      if (c < sm.J) return jnt[((long long)n * sm.K + (sm.ring_start + sm.K - 1) % sm.K) * sm.J + c];   // last frame, graph.py:388
      c -= sm.J;
      return sm.y[2][((long long)n * 4 + cell) * sm.D2 + c];
    }
    case VAR_SEQ_CONSTANT: {
      if (c < sm.D0) return sm.y[0][(((long long)t * N + n) * 4 + cell) * sm.D0 + c];
      c -= sm.D0;
      if (c < sm.J) return jnt[((long long)n * sm.K + (sm.ring_start + t) % sm.K) * sm.J + c];
      c -= sm.J;
      return sm.y[0][(((long long)sm.K * N + n) * 4 + cell) * sm.D0 + c];
    }
    case VAR_SEQ_RESIDUAL: {
      if (c < sm.D0)
        return sm.y[0][(((long long)sm.K * N + n) * 4 + cell) * sm.D0 + c] - sm.y[0][(((long long)t * N + n) * 4 + cell) * sm.D0 + c];
      c -= sm.D0;
      return jnt[((long long)n * sm.K + (sm.ring_start + t) % sm.K) * sm.J + c];
    }
    case VAR_SEQ_DYNDIFF: {
      if (c < sm.D0) return sm.y[0][(((long long)t * N + n) * 4 + cell) * sm.D0 + c];
      c -= sm.D0;
      if (c < sm.J) return jnt[((long long)n * sm.K + (sm.ring_start + t) % sm.K) * sm.J + c];
      c -= sm.J;
      return sm.y[1][(((long long)t * N + n) * 4 + cell) * sm.D2 + c];
    }
    default: {   // VAR_VMC
      if (c < sm.D0) return sm.y[0][(((long long)t * N + n) * 4 + cell) * sm.D0 + c];
      c -= sm.D0;
      return jnt[((long long)n * sm.K + (sm.ring_start + t) % sm.K) * sm.J + c];
    }
  }
}

__global__ void build_states_kernel(StateMap sm, const float* __restrict__ jnt_states, const float* __restrict__ m_prev,
                                    const unsigned char* __restrict__ reset_mask, float* __restrict__ states) {
  pdl_enter();
  const int n = blockIdx.x, t = blockIdx.y;
  float* row = states + ((long long)t * sm.N + n) * sm.ld;
  const int hi = t == 0 ? sm.ld : sm.xdim;
  const bool keep = m_prev && !(reset_mask && reset_mask[n]);
  for (int i = threadIdx.x; i < hi; i += blockDim.x) {
    float v;
    if (i >= sm.xdim) v = keep ? m_prev[(long long)n * sm.Hl + (i - sm.xdim)] : 0.f;
    else { const int cell = i / sm.per; v = state_value(sm, jnt_states, t, n, cell, i - cell * sm.per); }
    row[i] = v;
  }
}

// dY8(pre-activation)[group][img][cell][c] = (sum of the d(state) entries that read Y8[group][img][cell][c]) * (Y8 > 0).
// One thread per conv8 element gathers its contributions (a target feature is read by all T steps): deterministic.
__global__ void scatter_dstates_kernel(StateMap sm, const float* __restrict__ ds, int imgs0, int imgs1, int imgs2) {
  pdl_enter();
  const long long n0 = (long long)imgs0 * 4 * sm.D0, n1 = (long long)imgs1 * 4 * sm.D1, n2 = (long long)imgs2 * 4 * sm.D2;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0 + n1 + n2) return;
  int grp = 0, D = sm.D0;
  if (i >= n0 + n1) { grp = 2; i -= n0 + n1; D = sm.D2; }
  else if (i >= n0) { grp = 1; i -= n0; D = sm.variant == VAR_SEQ_DYNDIFF ? sm.D2 : sm.D1; }
  const int c = (int)(i % D);
  const int cell = (int)((i / D) & 3);
  const int img = (int)(i / (4ll * D));
  const int N = sm.N;
  auto at = [&](int t, int n, int off) { return ds[((long long)t * N + n) * sm.ld + cell * sm.per + off + c]; };
  float v = 0.f;
  switch (sm.variant) {
    case VAR_GEECOF:
      v = at(0, img, grp == 0 ? 0 : (grp == 1 ? sm.D0 : sm.D0 + sm.D1 + sm.J));
      break;
    case VAR_SEQ_CONSTANT:
      if (img < sm.K * N) v = at(img / N, img % N, 0);
      else for (int t = 0; t < sm.T; ++t) v += at(t, img - sm.K * N, sm.D0 + sm.J);
      break;
    case VAR_SEQ_RESIDUAL:
      if (img < sm.K * N) v = -at(img / N, img % N, 0);
      else for (int t = 0; t < sm.T; ++t) v += at(t, img - sm.K * N, 0);
      break;
    case VAR_SEQ_DYNDIFF:
      v = at(img / N, img % N, grp == 0 ? 0 : sm.D0 + sm.J);
      break;
    default:
      v = at(img / N, img % N, 0);
      break;
  }
  sm.g[grp][i] = sm.y[grp][i] > 0.f ? v : 0.f;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// LSTMCell (state_is_tuple=False, forget_bias=1): gates already hold [x,m]W + b.
__global__ void lstm_cell_kernel(int N, int Hl, const float* __restrict__ gates, const float* __restrict__ c_prev,
                                 const unsigned char* __restrict__ reset_mask, float* __restrict__ c_out,
                                 float* __restrict__ m_out, float* __restrict__ state_out, float* __restrict__ m_next,
                                 int ld_next) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Hl) return;
  const int n = idx / Hl, i = idx - n * Hl;
  const float* gr = gates + (long long)n * 4 * Hl;
  const float gi = gr[i], gj = gr[Hl + i], gf = gr[2 * Hl + i], go = gr[3 * Hl + i];
  const float cp = (c_prev && !(reset_mask && reset_mask[n])) ? c_prev[idx] : 0.f;
  const float c = sigmoidf_(gf + 1.0f) * cp + sigmoidf_(gi) * tanhf(gj);
  const float m = sigmoidf_(go) * tanhf(c);
  c_out[idx] = c;
  m_out[idx] = m;
  if (m_next) m_next[(long long)n * ld_next + i] = m;
  if (state_out) {
    state_out[(long long)n * 2 * Hl + i] = c;
    state_out[(long long)n * 2 * Hl + Hl + i] = m;
  }
}

// d(gates_t) and d(c_{t-1}) from d(m_t) (row stride ld_dm) and the d(c_t) that flows back from step t+1
__global__ void lstm_cell_bwd_kernel(int N, int Hl, const float* __restrict__ gates, const float* __restrict__ c_prev,
                                     const unsigned char* __restrict__ reset_mask, const float* __restrict__ dm, int ld_dm,
                                     const float* __restrict__ dc_in, float* __restrict__ dgates,
                                     float* __restrict__ dc_prev) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Hl) return;
  const int n = idx / Hl, i = idx - n * Hl;
  const float* gr = gates + (long long)n * 4 * Hl;
  const float gi = gr[i], gj = gr[Hl + i], gf = gr[2 * Hl + i], go = gr[3 * Hl + i];
  const float cp = (c_prev && !(reset_mask && reset_mask[n])) ? c_prev[idx] : 0.f;
  const float si = sigmoidf_(gi), tj = tanhf(gj), sf = sigmoidf_(gf + 1.0f), so = sigmoidf_(go);
  const float c = sf * cp + si * tj;
  const float tc = tanhf(c);
  const float dmv = dm[(long long)n * ld_dm + i];
  const float dso = dmv * tc;
  const float dc = dmv * so * (1.f - tc * tc) + (dc_in ? dc_in[idx] : 0.f);
  float* dg = dgates + (long long)n * 4 * Hl;
  dg[i] = dc * tj * si * (1.f - si);
  dg[Hl + i] = dc * si * (1.f - tj * tj);
  dg[2 * Hl + i] = dc * cp * sf * (1.f - sf);
  dg[3 * Hl + i] = dso * so * (1.f - so);
  if (dc_prev) dc_prev[idx] = dc * sf;
}

// ---------------------------------------------------------------------------------------
// fc1 + heads + per-sample loss pieces + d(loss)/d(head outputs).  One CTA per sample.
// Head table: TailHeads (columns, loss kind, target location).  loss_parts[n] = {term of head 0..4, correct}
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int head_of(const TailHeads& th, int t) {
  int h = 0;
#pragma unroll
  for (int i = 1; i < 5; ++i) if (i < th.nheads && t >= th.h[i].col) h = i;
  return h;
}

// One-step graphs fuse what used to be three launches in front of it (split-K reduce of the gate GEMM, LSTM cell):
// `cell.partial != NULL` -> this sample's gates = bias + sum of the split-K partials (same order as gates_reduce_kernel),
// then the cell, then fc1 / heads / losses from the m it just produced.
struct CellFuse {
  const float* partial; int slices;            // [slices][N][4Hl] split-K partials of the gate GEMM (NULL: m is given)
  const float* bias; float* gates;             // LSTM bias; full pre-activations out (kept for the backward)
  const float* c_prev; const unsigned char* reset_mask;
  float *c_out, *m_out, *state_out, *state_out2;   // state_out2: the caller's copy (geeco_outputs.lstm_state), may be NULL
};
struct TailOut { float *heads2, *fc1_2; };       // the caller's copies (geeco_outputs.heads / .fc1), may be NULL

__global__ void __launch_bounds__(512) tail_fwd_kernel(TailDims d, TailHeads th, const float* __restrict__ w_fc1,
                                                       const float* __restrict__ b_fc1, const float* __restrict__ m,
                                                       float* __restrict__ fc1, float* __restrict__ heads,
                                                       float* __restrict__ loss_parts, float* __restrict__ dheads,
                                                       int with_loss, CellFuse cell, TailOut outs) {
  pdl_enter();
  extern __shared__ float sm[];
  float* m_s = sm;                 // Hl
  float* fc_s = sm + d.Hl;         // Fc
  float* out_s = fc_s + d.Fc;      // NH (<= 32)
  float* g_s = out_s + 32;         // 4*Hl (fused cell only)
  const int n = blockIdx.x;
  const int NH = th.NH;
  if (cell.partial) {
    const int G4 = 4 * d.Hl;
    const long long total = (long long)d.N * G4;
    for (int j = threadIdx.x; j < G4; j += blockDim.x) {
      float s = cell.bias[j];
      const float* p = cell.partial + (long long)n * G4 + j;
      int k = 0;
      for (; k + 8 <= cell.slices; k += 8) {          // eight loads in flight, added in slice order
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(p + (long long)(k + q) * total);
#pragma unroll
        for (int q = 0; q < 8; ++q) s += v[q];
      }
      for (; k < cell.slices; ++k) s += __ldg(p + (long long)k * total);
      g_s[j] = s;
      cell.gates[(long long)n * G4 + j] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d.Hl; i += blockDim.x) {
      const long long idx = (long long)n * d.Hl + i;
      const float cp = (cell.c_prev && !(cell.reset_mask && cell.reset_mask[n])) ? cell.c_prev[idx] : 0.f;
      const float c = sigmoidf_(g_s[2 * d.Hl + i] + 1.0f) * cp + sigmoidf_(g_s[i]) * tanhf(g_s[d.Hl + i]);
      const float mv = sigmoidf_(g_s[3 * d.Hl + i]) * tanhf(c);
      cell.c_out[idx] = c; cell.m_out[idx] = mv; m_s[i] = mv;
      cell.state_out[(long long)n * 2 * d.Hl + i] = c; cell.state_out[(long long)n * 2 * d.Hl + d.Hl + i] = mv;
      if (cell.state_out2) { cell.state_out2[(long long)n * 2 * d.Hl + i] = c; cell.state_out2[(long long)n * 2 * d.Hl + d.Hl + i] = mv; }
    }
  } else {
    for (int i = threadIdx.x; i < d.Hl; i += blockDim.x) m_s[i] = m[(long long)n * d.Hl + i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d.Fc; j += blockDim.x) {
    float s = b_fc1[j];
#pragma unroll 16
    for (int i = 0; i < d.Hl; ++i) s = fmaf(m_s[i], __ldg(w_fc1 + (long long)i * d.Fc + j), s);   // loads issue ahead of the FMA chain
    s = fmaxf(s, 0.f);
    fc_s[j] = s;
    fc1[(long long)n * d.Fc + j] = s;
    if (outs.fc1_2) outs.fc1_2[(long long)n * d.Fc + j] = s;
  }
  __syncthreads();
  // heads: one warp per output column (lanes stride over fc1, fixed-order lane reduction), 4 warps take turns
  for (int t = threadIdx.x >> 5; t < NH; t += blockDim.x >> 5) {
    const HeadSpec& hs = th.h[head_of(th, t)];
    const int col = t - hs.col, width = hs.width;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int j = lane; j < d.Fc; j += 32) s = fmaf(fc_s[j], __ldg(hs.w + j * width + col), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += hs.b[col];
    if (lane == 0) { out_s[t] = s; heads[(long long)n * NH + t] = s; if (outs.heads2) outs.heads2[(long long)n * NH + t] = s; }
  }
  __syncthreads();
  if (with_loss && threadIdx.x == 0) {
    float* dh = dheads + (long long)n * NH;
    float* lp = loss_parts + (long long)n * 6;
    lp[5] = 0.f;
    for (int hi = 0; hi < th.nheads; ++hi) {
      const HeadSpec& hs = th.h[hi];
      const float* tg = hs.target + (long long)n * hs.tstride + hs.toff;
      const float* o = out_s + hs.col;
      if (hs.kind == 0) {
        // tf.losses.mean_squared_error: mean over N*width elements (SUM_BY_NONZERO_WEIGHTS)
        const float inv = 1.f / ((float)hs.width * d.N);
        float se = 0.f;
        for (int k = 0; k < hs.width; ++k) {
          const float e = o[k] - tg[k];
          se += e * e;
          dh[hs.col + k] = hs.weight * 2.f * e * inv;
        }
        lp[hi] = se;
      } else {
        // estimator.py:213-216: class = int32(rint(cmd[:,3])) + 1 ; one_hot of an out-of-range index is all zeros
        const int cls = (int)rintf(tg[0]) + 1;
        float zmax = -3.4e38f; int amax = 0;
        for (int k = 0; k < hs.width; ++k) if (o[k] > zmax) { zmax = o[k]; amax = k; }
        float se = 0.f;
        for (int k = 0; k < hs.width; ++k) se += expf(o[k] - zmax);
        const float lse = zmax + logf(se);
        const bool ok = cls >= 0 && cls < hs.width;
        for (int k = 0; k < hs.width; ++k) {
          const float pk = expf(o[k] - lse);
          dh[hs.col + k] = ok ? hs.weight * (pk - (k == cls ? 1.f : 0.f)) / d.N : 0.f;
        }
        lp[hi] = ok ? lse - o[cls] : 0.f;
        lp[5] = (amax == cls) ? 1.f : 0.f;
      }
    }
  }
}

// losses[12]: slot of every head (HeadSpec::slot), [4] loss_reg, [5] loss, [6] sum_correct, [7] N
__global__ void loss_reduce_kernel(int N, TailHeads th, const float* __restrict__ loss_parts,
                                   const float* __restrict__ reg_term, float* __restrict__ losses,
                                   float* __restrict__ losses2) {
  pdl_enter();
  __shared__ float s[6][256];
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n = threadIdx.x; n < N; n += blockDim.x)
    for (int k = 0; k < 6; ++k) a[k] += loss_parts[(long long)n * 6 + k];
  for (int k = 0; k < 6; ++k) s[k][threadIdx.x] = a[k];
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) for (int k = 0; k < 6; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 12; ++k) losses[k] = 0.f;
    // command terms (weight 1) and auxiliary terms (one common weight): estimator.py:224-225 adds
    // add_n(command losses) + lambda_aux * add_n(pose losses); mse_loss (graph.py:449) is one add_n over all five
    float cmd = 0.f, aux = 0.f, wa = 1.f;
    bool any_cmd = false, any_aux = false;
    for (int hi = 0; hi < th.nheads; ++hi) {
      const HeadSpec& hs = th.h[hi];
      const float l = hs.kind == 0 ? s[hi][0] / ((float)hs.width * N) : s[hi][0] / N;
      losses[hs.slot] = l;
      if (!hs.aux) { cmd = any_cmd ? cmd + l : l; any_cmd = true; }
      else { aux = any_aux ? aux + l : l; any_aux = true; wa = hs.weight; }
    }
    const float lr = reg_term ? reg_term[0] : 0.f;
    losses[4] = lr;
    losses[5] = cmd + (any_aux ? wa * aux : 0.f) + lr;
    losses[6] = s[5][0]; losses[7] = (float)N;
    if (losses2) for (int k = 0; k < 12; ++k) losses2[k] = losses[k];
  }
}

// backward through heads and fc1 -> dL/dm; with dm_out == NULL also through the (single) LSTM cell -> d(gates).
// One CTA per sample.
__global__ void __launch_bounds__(128) tail_bwd_kernel(TailDims d, TailHeads th, const float* __restrict__ w_fc1,
                                                       const float* __restrict__ fc1, const float* __restrict__ dheads,
                                                       const float* __restrict__ gates, const float* __restrict__ c_prev,
                                                       const unsigned char* __restrict__ reset_mask,
                                                       float* __restrict__ dfc1, float* __restrict__ dgates,
                                                       float* __restrict__ dm_out) {
  pdl_enter();
  extern __shared__ float sm[];
  float* dh_s = sm;                // NH (padded to 32)
  float* dfc_s = sm + 32;          // Fc
  const int n = blockIdx.x;
  const int NH = th.NH;
  if (threadIdx.x < NH) dh_s[threadIdx.x] = dheads[(long long)n * NH + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.x; j < d.Fc; j += blockDim.x) {
    float s = 0.f;
    for (int hi = 0; hi < th.nheads; ++hi) {
      const HeadSpec& hs = th.h[hi];
      for (int k = 0; k < hs.width; ++k) s = fmaf(dh_s[hs.col + k], hs.w[j * hs.width + k], s);
    }
    s = fc1[(long long)n * d.Fc + j] > 0.f ? s : 0.f;
    dfc_s[j] = s;
    dfc1[(long long)n * d.Fc + j] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.Hl; i += blockDim.x) {
    float dm = 0.f;
    const float* wr = w_fc1 + (long long)i * d.Fc;
    for (int j = 0; j < d.Fc; ++j) dm = fmaf(dfc_s[j], wr[j], dm);
    if (dm_out) { dm_out[(long long)n * d.Hl + i] = dm; continue; }
    const float* gr = gates + (long long)n * 4 * d.Hl;
    const float gi = gr[i], gj = gr[d.Hl + i], gf = gr[2 * d.Hl + i], go = gr[3 * d.Hl + i];
    const float cp = (c_prev && !(reset_mask && reset_mask[n])) ? c_prev[(long long)n * d.Hl + i] : 0.f;
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
__global__ void tail_wgrad_kernel(TailDims d, TailHeads th, float* __restrict__ gw_fc1, float* __restrict__ gb_fc1,
                                  const float* __restrict__ m, const float* __restrict__ fc1,
                                  const float* __restrict__ dfc1, const float* __restrict__ dheads) {
  pdl_enter();
  const int NH = th.NH;
  const int n_w1 = d.Hl * d.Fc, n_b1 = d.Fc, n_wh = d.Fc * NH, n_bh = NH;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_w1) {
    const int i = idx / d.Fc, j = idx - i * d.Fc;
    float s = 0.f;
    for (int n = 0; n < d.N; ++n) s = fmaf(m[(long long)n * d.Hl + i], dfc1[(long long)n * d.Fc + j], s);
    gw_fc1[idx] = s;
  } else if (idx < n_w1 + n_b1) {
    const int j = idx - n_w1;
    float s = 0.f;
    for (int n = 0; n < d.N; ++n) s += dfc1[(long long)n * d.Fc + j];
    gb_fc1[j] = s;
  } else if (idx < n_w1 + n_b1 + n_wh + n_bh) {
    const int e = idx - n_w1 - n_b1;
    const bool is_bias = e >= n_wh;
    const int j = is_bias ? 0 : e / NH, t = is_bias ? e - n_wh : e - (e / NH) * NH;
    float s = 0.f;
    if (is_bias) for (int n = 0; n < d.N; ++n) s += dheads[(long long)n * NH + t];
    else for (int n = 0; n < d.N; ++n) s = fmaf(fc1[(long long)n * d.Fc + j], dheads[(long long)n * NH + t], s);
    const HeadSpec& hs = th.h[head_of(th, t)];
    const int col = t - hs.col;
    if (is_bias) hs.gb[col] = s; else hs.gw[j * hs.width + col] = s;
  }
}

// ---------------------------------------------------------------------------------------
// TF Adam (training_ops.apply_adam): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1);
// v += (g*g-v)(1-b2); theta -= lr_t*m/(sqrt(v)+eps).   One launch over the flat arena.
// sc[0] = t (as float), sc[1] = lr_t, sc[2] = 0.5*l2*sum(theta^2)
// ---------------------------------------------------------------------------------------
// lr_t comes from the host (the step counter lives in the context): one launch, nothing in front of it.
__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ theta, const float4* __restrict__ grad,
                                                   float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                   float lr_t, float b1, float b2, float eps, float gscale, float l2) {
  pdl_enter();
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
  pdl_enter();
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
int launch_build_states(const StateMap& sm, const float* jnt, const float* m_prev, const unsigned char* reset_mask,
                        float* states, cudaStream_t st) {
  GEECO_LAUNCH((build_states_kernel), dim3(sm.N, sm.T), 256, 0, st, sm, jnt, m_prev, reset_mask, states);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
static void group_images(const StateMap& sm, int* imgs) {
  imgs[0] = imgs[1] = imgs[2] = 0;
  switch (sm.variant) {
    case VAR_GEECOF: imgs[0] = imgs[1] = imgs[2] = sm.N; break;
    case VAR_SEQ_CONSTANT: case VAR_SEQ_RESIDUAL: imgs[0] = (sm.K + 1) * sm.N; break;
    case VAR_SEQ_DYNDIFF: imgs[0] = imgs[1] = sm.K * sm.N; break;
    default: imgs[0] = sm.K * sm.N; break;
  }
}
int launch_scatter_dstates(const StateMap& sm, const float* dstates, cudaStream_t st) {
  int imgs[3];
  group_images(sm, imgs);
  const int D1 = sm.variant == VAR_SEQ_DYNDIFF ? sm.D2 : sm.D1;
  const long long total = 4ll * ((long long)imgs[0] * sm.D0 + (long long)imgs[1] * D1 + (long long)imgs[2] * sm.D2);
  // scatter_dstates_kernel indexes group 1 with width D1 (= D2 for sequence/dyndiff, whose second group is the
  // DynDiffEncoder): pass that width in the D1 slot of a copy
  StateMap k = sm;
  k.D1 = D1;
  GEECO_LAUNCH((scatter_dstates_kernel), ceil_div(total, 256), 256, 0, st, k, dstates, imgs[0], imgs[1], imgs[2]);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_lstm_cell(int N, int Hl, const float* gates, const float* c_prev, const unsigned char* reset_mask,
                     float* c_out, float* m_out, float* state_out, float* m_next, int ld_next, cudaStream_t st) {
  GEECO_LAUNCH((lstm_cell_kernel), ceil_div((long long)N * Hl, 256), 256, 0, st, N, Hl, gates, c_prev, reset_mask, c_out, m_out,
                                                                    state_out, m_next, ld_next);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_lstm_cell_bwd(int N, int Hl, const float* gates, const float* c_prev, const unsigned char* reset_mask,
                         const float* dm, int ld_dm, const float* dc_in, float* dgates, float* dc_prev, cudaStream_t st) {
  GEECO_LAUNCH((lstm_cell_bwd_kernel), ceil_div((long long)N * Hl, 256), 256, 0, st, N, Hl, gates, c_prev, reset_mask, dm, ld_dm, dc_in,
                                                                        dgates, dc_prev);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_tail_fwd(const TailDims& d, const TailHeads& th, const float* w_fc1, const float* b_fc1, const float* m,
                    float* fc1, float* heads, float* loss_parts, float* dheads, int with_loss, float* heads_out,
                    float* fc1_out, cudaStream_t st) {
  const size_t smem = (size_t)(d.Hl + d.Fc + 32) * sizeof(float);
  CellFuse cell = {};
  TailOut outs = {heads_out, fc1_out};
  GEECO_LAUNCH((tail_fwd_kernel), d.N, 128, smem, st, d, th, w_fc1, b_fc1, m, fc1, heads, loss_parts, dheads, with_loss, cell, outs);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
// one-step graphs: split-K partials of the gate GEMM -> gates -> cell -> fc1 -> heads -> losses in ONE launch
int launch_tail_fwd_fused(const TailDims& d, const TailHeads& th, const float* w_fc1, const float* b_fc1,
                          const float* partial, int slices, const float* lstm_bias, float* gates, const float* c_prev,
                          const unsigned char* reset_mask, float* c_out, float* m_out, float* state_out, float* fc1,
                          float* heads, float* loss_parts, float* dheads, int with_loss, float* heads_out, float* fc1_out,
                          float* state_out2, cudaStream_t st) {
  const size_t smem = (size_t)(d.Hl + d.Fc + 32 + 4 * d.Hl) * sizeof(float);
  CellFuse cell = {partial, slices, lstm_bias, gates, c_prev, reset_mask, c_out, m_out, state_out, state_out2};
  TailOut outs = {heads_out, fc1_out};
  int threads = 4 * d.Hl; if (threads > 512) threads = 512; if (threads < 128) threads = 128;
  GEECO_LAUNCH((tail_fwd_kernel), d.N, threads, smem, st, d, th, w_fc1, b_fc1, (const float*)nullptr, fc1, heads, loss_parts, dheads,
               with_loss, cell, outs);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_loss_reduce(const TailDims& d, const TailHeads& th, const float* loss_parts, const float* reg_term,
                       float* losses, float* losses_out, cudaStream_t st) {
  GEECO_LAUNCH((loss_reduce_kernel), 1, 256, 0, st, d.N, th, loss_parts, reg_term, losses, losses_out);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_tail_bwd(const TailDims& d, const TailHeads& th, const float* w_fc1, float* gw_fc1, float* gb_fc1,
                    const float* m, const float* fc1, const float* dheads, const float* gates, const float* c_prev,
                    const unsigned char* reset_mask, float* dfc1, float* dgates, float* dm_out, cudaStream_t st) {
  const size_t smem = (size_t)(32 + d.Fc) * sizeof(float);
  GEECO_LAUNCH((tail_bwd_kernel), d.N, 128, smem, st, d, th, w_fc1, fc1, dheads, gates, c_prev, reset_mask, dfc1, dgates, dm_out);
  const int total = d.Hl * d.Fc + d.Fc + d.Fc * th.NH + th.NH;
  GEECO_LAUNCH((tail_wgrad_kernel), ceil_div(total, 128), 128, 0, st, d, th, gw_fc1, gb_fc1, m, fc1, dfc1, dheads);
  geeco_count_launch(2);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_adam(float* theta, const float* grad, float* m, float* v, long long n, long long t, double lr, double b1,
                double b2, double eps, float gscale, float l2, cudaStream_t st) {
  // tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) for the t-th update (t >= 1), in double
  const float lr_t = (float)(lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t)));
  const long long n4 = n / 4;
  int blocks = ceil_div(n4, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  GEECO_LAUNCH((adam_kernel), blocks, 256, 0, st, reinterpret_cast<float4*>(theta), reinterpret_cast<const float4*>(grad),
                                      reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4, lr_t, (float)b1, (float)b2,
                                      (float)eps, gscale, l2);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_l2_term(const float* theta, long long n, float l2, float* sc, cudaStream_t st) {
  GEECO_LAUNCH((l2_term_kernel), 1, 1024, 0, st, theta, n, l2, sc);
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
  pdl_enter();
  __shared__ float xs[32][GK_SLICE + 8 + 1];      // (+8: the last chunk of a short slice reads zero-weighted columns)
  const int k0 = blockIdx.x * GK_SLICE, n0 = blockIdx.y * 32;
  for (int e = threadIdx.x; e < 32 * (GK_SLICE + 8); e += 256) {
    const int n = e / (GK_SLICE + 8), k = e - n * (GK_SLICE + 8);
    xs[n][k] = (k < GK_SLICE && n0 + n < N && k0 + k < K) ? x[(long long)(n0 + n) * ldx + k0 + k] : 0.f;
  }
  __syncthreads();
  float acc[32][CPT];
#pragma unroll
  for (int n = 0; n < 32; ++n)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[n][c] = 0.f;
  const int kend = (K - k0) < GK_SLICE ? (K - k0) : GK_SLICE;
  // in chunks of 8 kernel rows: all loads of a chunk are issued before its FMAs (the first version loaded one row per
  // iteration and waited for it: 32 dependent L2 round trips per CTA)
  for (int kc = 0; kc < kend; kc += 8) {
    float w[8][CPT];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int col = threadIdx.x + c * 256;
        w[kk][c] = (kc + kk < kend && col < Ncols) ? __ldg(W + (long long)(k0 + kc + kk) * Ncols + col) : 0.f;
      }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
      for (int n = 0; n < 32; ++n) {
        const float a = xs[n][kc + kk];
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[n][c] = fmaf(a, w[kk][c], acc[n][c]);
      }
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
  pdl_enter();
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
  pdl_enter();
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
  GEECO_LAUNCH((copy_outputs_kernel), dim3(16, 4), 256, 0, st, cl);
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
// fused GEECO-F scatter: instead of d(state), write dL/d(pre-activation) of conv8 directly (state index -> encoder
// group / cell / channel, ReLU mask from the conv8 output), fp32 or bf16
struct DstateScatter { int on; int D0, D1, D2, J, per; const float* y[3]; float* g[3]; __nv_bfloat16* gb[3]; };

__global__ void __launch_bounds__(256) lstm_dstate_kernel(const float* __restrict__ dgates, const float* __restrict__ W,
                                                          float* __restrict__ dstate, int N, int xdim, int NC, int ld,
                                                          DstateScatter sc) {
  pdl_enter();
  __shared__ float Ws[16][65];
  __shared__ float Ds[64][65];
  const int k0 = blockIdx.x * 16, n0 = blockIdx.y * 64;
  const int k = threadIdx.x & 15, ng = threadIdx.x >> 4;      // 16 kernel rows x 16 groups of 4 samples
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < NC; j0 += 64) {
    // all 20 loads of this thread are issued before the first shared-memory store (a load/store loop serialised them)
    float wv[4], dv[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + q * 256, r = e >> 6, c = e & 63;
      wv[q] = (k0 + r < xdim && j0 + c < NC) ? __ldg(W + (long long)(k0 + r) * NC + j0 + c) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = threadIdx.x + q * 256, r = e >> 6, c = e & 63;
      dv[q] = (n0 + r < N && j0 + c < NC) ? __ldg(dgates + (long long)(n0 + r) * NC + j0 + c) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int e = threadIdx.x + q * 256; Ws[e >> 6][e & 63] = wv[q]; }
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int e = threadIdx.x + q * 256; Ds[e >> 6][e & 63] = dv[q]; }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < 64; ++jj) {
      const float w = Ws[k][jj];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(Ds[ng * 4 + i][jj], w, acc[i]);
    }
    __syncthreads();
  }
  const int col = k0 + k;
  if (col >= xdim) return;
  int grp = -1, D = 0, cell = 0, cc = 0;
  if (sc.on) {
    cell = col / sc.per;
    int c = col - cell * sc.per;
    if (c < sc.D0) { grp = 0; D = sc.D0; cc = c; }
    else if (c < sc.D0 + sc.D1) { grp = 1; D = sc.D1; cc = c - sc.D0; }
    else if (c >= sc.D0 + sc.D1 + sc.J) { grp = 2; D = sc.D2; cc = c - sc.D0 - sc.D1 - sc.J; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ng * 4 + i;
    if (n >= N) continue;
    if (!sc.on) { dstate[(long long)n * ld + col] = acc[i]; continue; }
    if (grp < 0) continue;                                    // joint-state columns: data, no gradient wanted
    const long long o = ((long long)n * 4 + cell) * D + cc;
    const float v = sc.y[grp][o] > 0.f ? acc[i] : 0.f;
    if (sc.gb[grp]) sc.gb[grp][o] = __float2bfloat16_rn(v);
    else sc.g[grp][o] = v;
  }
}
int launch_lstm_dstate(const float* dgates, const float* W, float* dstate, int N, int xdim, int ncols, int ld,
                       cudaStream_t st) {
  DstateScatter sc = {};
  GEECO_LAUNCH((lstm_dstate_kernel), dim3((xdim + 15) / 16, (N + 63) / 64), 256, 0, st, dgates, W, dstate, N, xdim, ncols, ld, sc);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
// GEECO-F (one step): d(state) is never materialised, the conv8 gradients are written directly
int launch_lstm_dstate_scatter(const float* dgates, const float* W, const StateMap& sm, __nv_bfloat16* const* g_bf16, int ncols,
                               cudaStream_t st) {
  DstateScatter sc = {};
  sc.on = 1; sc.D0 = sm.D0; sc.D1 = sm.D1; sc.D2 = sm.D2; sc.J = sm.J; sc.per = sm.per;
  for (int e = 0; e < 3; ++e) { sc.y[e] = sm.y[e]; sc.g[e] = sm.g[e]; sc.gb[e] = g_bf16 ? g_bf16[e] : nullptr; }
  GEECO_LAUNCH((lstm_dstate_kernel), dim3((sm.xdim + 15) / 16, (sm.N + 63) / 64), 256, 0, st, dgates, W, (float*)nullptr, sm.N, sm.xdim,
               ncols, sm.ld, sc);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

// ---------------------------------------------------------------------------------------
// d(kernel)[k][j] = sum_r states[r][k] * dgates[r][j], d(bias)[j] = sum_r dgates[r][j] over the R = T*N state rows.
// R is small (64 at batch 64): no split over the reduction, one pass, fixed order.  Tile = 32 kernel rows x 64 gate
// columns per CTA, the rows of both operands staged 32 at a time.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_wgrad_kernel(const float* __restrict__ states, const float* __restrict__ dgates,
                                                         float* __restrict__ dW, float* __restrict__ db, int R, int ld, int NC) {
  pdl_enter();
  __shared__ float Ss[32][33];
  __shared__ float Dd[32][64];
  const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 64;
  const int jj = threadIdx.x & 63, kg = threadIdx.x >> 6;      // 64 columns x 4 groups of 8 kernel rows
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float bsum = 0.f;
  for (int r0 = 0; r0 < R; r0 += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int r = e >> 5, k = e & 31;
      Ss[r][k] = (r0 + r < R && k0 + k < ld) ? __ldg(states + (long long)(r0 + r) * ld + k0 + k) : 0.f;
    }
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      Dd[r][c] = (r0 + r < R && j0 + c < NC) ? __ldg(dgates + (long long)(r0 + r) * NC + j0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float dv = Dd[r][jj];
      bsum += dv;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(Ss[r][kg * 8 + i], dv, acc[i]);
    }
    __syncthreads();
  }
  if (j0 + jj < NC) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + kg * 8 + i;
      if (k < ld) dW[(long long)k * NC + j0 + jj] = acc[i];
    }
    if (blockIdx.x == 0 && kg == 0) db[j0 + jj] = bsum;
  }
}
int launch_lstm_wgrad(const float* states, const float* dgates, float* dW, float* db, int R, int ld, int ncols, cudaStream_t st) {
  GEECO_LAUNCH((lstm_wgrad_kernel), dim3((ld + 31) / 32, (ncols + 63) / 64), 256, 0, st, states, dgates, dW, db, R, ld, ncols);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

long long lstm_gates_partial_floats(int N, int K, int Ncols) {
  return (long long)((K + GK_SLICE - 1) / GK_SLICE) * N * Ncols;
}
int launch_lstm_gates(const float* x, int ldx, const float* W, const float* bias, float* gates, float* partial, int N,
                      int K, int Ncols, cudaStream_t st, int* slices_out) {
  if (Ncols > 1024) { geeco_set_error("lstm gates: 4*dim_h_lstm = %d > 1024 not supported", Ncols); return GEECO_ERR_INVALID; }
  const int slices = (K + GK_SLICE - 1) / GK_SLICE;
  dim3 grid(slices, (N + 31) / 32);
  const int cpt = (Ncols + 255) / 256;
  if (cpt == 1) GEECO_LAUNCH((gates_splitk_kernel<1>), grid, 256, 0, st, x, W, partial, N, K, ldx, Ncols);
  else if (cpt == 2) GEECO_LAUNCH((gates_splitk_kernel<2>), grid, 256, 0, st, x, W, partial, N, K, ldx, Ncols);
  else if (cpt == 3) GEECO_LAUNCH((gates_splitk_kernel<3>), grid, 256, 0, st, x, W, partial, N, K, ldx, Ncols);
  else GEECO_LAUNCH((gates_splitk_kernel<4>), grid, 256, 0, st, x, W, partial, N, K, ldx, Ncols);
  if (slices_out) *slices_out = slices;         // the caller's next kernel reduces the partials itself
  else GEECO_LAUNCH((gates_reduce_kernel), ceil_div((long long)N * Ncols, 256), 256, 0, st, partial, bias, gates, slices, N, Ncols);
  geeco_count_launch(slices_out ? 1 : 2);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}


// ---------------------------------------------------------------------------------------
// device ring of K-frame histories (include/geeco_b200.h: geeco_ring_push)
// ---------------------------------------------------------------------------------------
template <typename V>
__global__ void ring_push_kernel(V* __restrict__ ring, const V* __restrict__ frame, const unsigned char* __restrict__ fresh,
                                 int K, long long row_v, int slot) {
  pdl_enter();
  const int n = blockIdx.y;
  const bool all = fresh && fresh[n];
  const V* src = frame + (long long)n * row_v;
  V* dst = ring + (long long)n * K * row_v;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_v; i += (long long)gridDim.x * blockDim.x) {
    const V v = src[i];
    if (all) { for (int k = 0; k < K; ++k) dst[(long long)k * row_v + i] = v; }
    else dst[(long long)slot * row_v + i] = v;
  }
}
int launch_ring_push(void* ring, const void* frame, const unsigned char* fresh, int N, int K, long long row_bytes,
                     int slot, cudaStream_t st) {
  if (N <= 0) return GEECO_OK;
  if (N > 65535) { geeco_set_error("ring_push: N=%d > 65535", N); return GEECO_ERR_INVALID; }
  const bool v16 = row_bytes % 16 == 0 && ((uintptr_t)ring % 16 == 0) && ((uintptr_t)frame % 16 == 0);
  const long long row_v = row_bytes / (v16 ? 16 : 4);
  int bx = (int)((row_v + 255) / 256); if (bx > 64) bx = 64; if (bx < 1) bx = 1;
  if (v16) GEECO_LAUNCH((ring_push_kernel<uint4>), dim3(bx, N), 256, 0, st, (uint4*)ring, (const uint4*)frame, fresh, K, row_v, slot);
  else GEECO_LAUNCH((ring_push_kernel<unsigned int>), dim3(bx, N), 256, 0, st, (unsigned int*)ring, (const unsigned int*)frame, fresh, K, row_v, slot);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
