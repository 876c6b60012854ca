// Persistent recurrent LSTM kernels: the K-step loop of `lstm_decoder` (src/models/e2evmc/graph.py:217-225) and its
// back-propagation through time as ONE launch each, with the recurrent weights W_h (the last dim_h_lstm rows of the
// LSTMCell kernel) resident in shared memory for all steps.
//
//   gates_t = xg_t + m_{t-1} @ W_h            xg_t = x_t @ W_x + bias, for all t at once by one split GEMM (tail.cu)
//   c_t = sigmoid(f + 1) * c_{t-1} + sigmoid(i) * tanh(j) ;  m_t = sigmoid(o) * tanh(c_t)       (i, j, f, o = split(gates_t))
//
// Work split.  Samples are independent, hidden units are not: a thread-block CLUSTER of CL CTAs owns a block of RB
// batch rows for all T steps; CTA r of the cluster owns U = Hl / CL hidden units, i.e. the 4*U gate columns
// {g*Hl + r*U + u} of W_h (forward) or the U rows r*U + u of W_h (backward), loaded into shared memory ONCE.  After
// every step a CTA publishes its part of m_t (forward) / d(gates_t) (backward) into the shared memory of all CTAs of
// its cluster (distributed shared memory), one cluster barrier per step.  fp32 throughout (parity mode and bf16 mode
// share this tail); deterministic (fixed summation order).
#include "tail.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

static constexpr int LP_RB = 32;          // batch rows per cluster
static constexpr int LP_THREADS = 256;

__device__ __forceinline__ float lp_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

struct LstmSeqArgs {
  int T, N, Hl, xdim, ld, U;
  const float* Wh;            // [Hl][4Hl]  (kernel + xdim * 4Hl)
  const float* c0; const float* m0; const unsigned char* reset;   // initial state [N][Hl] (NULL = zeros), per-row reset
  float* gates;               // [T][N][4Hl]  in: xg (forward) ; out: full pre-activations.  Backward: input.
  float* c; float* m;         // [T][N][Hl]
  float* states;              // [T][N][ld]: forward writes m_t into the m part of row t+1 (input of the weight-gradient GEMM)
  float* state_out;           // [N][2Hl] final [c | m] (forward, optional)
  const float* dm_last;       // [N][Hl]  dL/dm_{T-1}           (backward)
  float* dgates;              // [T][N][4Hl]                     (backward out)
};

// ---------------------------------------------------------------------------------------
// forward.  shared: Ws [Hl][4U] | mb [2][RB][Hl]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LP_THREADS) lstm_seq_fwd_kernel(LstmSeqArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float lp_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = cluster.num_blocks(), rank = cluster.block_rank();
  const int Hl = a.Hl, U = a.U, G4 = 4 * Hl;
  float* Ws = lp_smem;                         // [k][g*U + u]
  float* mb = lp_smem + (size_t)Hl * 4 * U;    // [2][RB][Hl]
  const int row0 = blockIdx.y * LP_RB;
  const int u0 = rank * U;
  for (int e = threadIdx.x; e < Hl * 4 * U; e += LP_THREADS) {
    const int k = e / (4 * U), c = e - k * 4 * U, g = c / U, u = c - g * U;
    Ws[e] = a.Wh[(long long)k * G4 + g * Hl + u0 + u];
  }
  for (int e = threadIdx.x; e < LP_RB * Hl; e += LP_THREADS) {
    const int r = e / Hl, k = e - r * Hl, n = row0 + r;
    const bool keep = a.m0 && n < a.N && !(a.reset && a.reset[n]);
    mb[e] = keep ? a.m0[(long long)n * Hl + k] : 0.f;
  }
  cluster.sync();
  // thread -> unit u (fastest) and rows r, r + RS, ...
  const int u = threadIdx.x % U, rl = threadIdx.x / U, RS = LP_THREADS / U;
  for (int t = 0; t < a.T; ++t) {
    const float* mcur = mb + (size_t)(t & 1) * LP_RB * Hl;
    float* mnext = mb + (size_t)((t + 1) & 1) * LP_RB * Hl;
    for (int r = rl; r < LP_RB; r += RS) {
      const int n = row0 + r;
      if (n >= a.N) continue;
      float* gr = a.gates + ((long long)t * a.N + n) * G4;
      float acc[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g] = gr[g * Hl + u0 + u];
      const bool has_prev = t > 0 || a.m0 != nullptr;
      if (has_prev) {
        const float* mr = mcur + r * Hl;
#pragma unroll 4
        for (int k = 0; k < Hl; ++k) {
          const float mv = mr[k];
          const float* w = Ws + (size_t)k * 4 * U + u;
#pragma unroll
          for (int g = 0; g < 4; ++g) acc[g] = fmaf(mv, w[g * U], acc[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) gr[g * Hl + u0 + u] = acc[g];
      const long long si = (long long)n * Hl + u0 + u;
      float cp = 0.f;
      if (t > 0) cp = a.c[(long long)(t - 1) * a.N * Hl + si];
      else if (a.c0 && !(a.reset && a.reset[n])) cp = a.c0[si];
      const float cv = lp_sigmoid(acc[2] + 1.0f) * cp + lp_sigmoid(acc[0]) * tanhf(acc[1]);
      const float mv = lp_sigmoid(acc[3]) * tanhf(cv);
      a.c[(long long)t * a.N * Hl + si] = cv;
      a.m[(long long)t * a.N * Hl + si] = mv;
      if (t + 1 < a.T) {
        a.states[((long long)(t + 1) * a.N + n) * a.ld + a.xdim + u0 + u] = mv;
        for (int p = 0; p < CL; ++p) cluster.map_shared_rank(mnext, p)[r * Hl + u0 + u] = mv;   // DSMEM store
      } else if (a.state_out) {
        a.state_out[(long long)n * 2 * Hl + u0 + u] = cv;
        a.state_out[(long long)n * 2 * Hl + Hl + u0 + u] = mv;
      }
    }
    if (t + 1 < a.T) cluster.sync();     // m_t of every unit is in every CTA's buffer
  }
}

// ---------------------------------------------------------------------------------------
// backward.  shared: Wt [4Hl][U] (Wt[col][u] = W_h[u0 + u][col]) | db [2][RB][4Hl] | dcb [RB][U]
// d(m_t) = (t == T-1 ? dm_last : d(gates_{t+1}) @ W_h^T) ; cell backward with the d(c) carried from step t+1
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LP_THREADS) lstm_seq_bwd_kernel(LstmSeqArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float lp_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = cluster.num_blocks(), rank = cluster.block_rank();
  const int Hl = a.Hl, U = a.U, G4 = 4 * Hl;
  float* Wt = lp_smem;                                  // [col][u]
  float* db = Wt + (size_t)G4 * U;                      // [2][RB][4Hl]
  float* dcb = db + (size_t)2 * LP_RB * G4;             // [RB][U]
  const int row0 = blockIdx.y * LP_RB;
  const int u0 = rank * U;
  for (int e = threadIdx.x; e < G4 * U; e += LP_THREADS) {
    const int uu = e / G4, col = e - uu * G4;           // coalesced read of row u0 + uu
    Wt[(size_t)col * U + uu] = a.Wh[(long long)(u0 + uu) * G4 + col];
  }
  for (int e = threadIdx.x; e < LP_RB * U; e += LP_THREADS) dcb[e] = 0.f;
  cluster.sync();
  const int u = threadIdx.x % U, rl = threadIdx.x / U, RS = LP_THREADS / U;
  for (int t = a.T - 1; t >= 0; --t) {
    const float* dcur = db + (size_t)(t & 1) * LP_RB * G4;          // d(gates_{t+1}), all columns
    float* dnext = db + (size_t)((t + 1) & 1) * LP_RB * G4;         // receives d(gates_t)
    for (int r = rl; r < LP_RB; r += RS) {
      const int n = row0 + r;
      if (n >= a.N) continue;
      const long long si = (long long)n * Hl + u0 + u;
      float dmv;
      if (t == a.T - 1) dmv = a.dm_last[si];
      else {
        dmv = 0.f;
        const float* dr = dcur + r * G4;
#pragma unroll 4
        for (int col = 0; col < G4; ++col) dmv = fmaf(dr[col], Wt[(size_t)col * U + u], dmv);
      }
      const float* gr = a.gates + ((long long)t * a.N + n) * G4;
      const float gi = gr[u0 + u], gj = gr[Hl + u0 + u], gf = gr[2 * Hl + u0 + u], go = gr[3 * Hl + u0 + u];
      float cp = 0.f;
      if (t > 0) cp = a.c[(long long)(t - 1) * a.N * Hl + si];
      else if (a.c0 && !(a.reset && a.reset[n])) cp = a.c0[si];
      const float si_ = lp_sigmoid(gi), tj = tanhf(gj), sf = lp_sigmoid(gf + 1.0f), so = lp_sigmoid(go);
      const float cv = sf * cp + si_ * tj;
      const float tc = tanhf(cv);
      const float dso = dmv * tc;
      const float dc = dmv * so * (1.f - tc * tc) + dcb[r * U + u];
      const float d4[4] = {dc * tj * si_ * (1.f - si_), dc * si_ * (1.f - tj * tj), dc * cp * sf * (1.f - sf),
                           dso * so * (1.f - so)};
      dcb[r * U + u] = dc * sf;                                      // own (row, unit): no other thread touches it
      float* dg = a.dgates + ((long long)t * a.N + n) * G4;
#pragma unroll
      for (int g = 0; g < 4; ++g) dg[g * Hl + u0 + u] = d4[g];
      if (t > 0) {
        for (int p = 0; p < CL; ++p) {
          float* peer = cluster.map_shared_rank(dnext, p) + r * G4;
#pragma unroll
          for (int g = 0; g < 4; ++g) peer[g * Hl + u0 + u] = d4[g];
        }
      }
    }
    if (t > 0) cluster.sync();
  }
}

// cluster size and shared-memory need; returns false when the shape does not fit (caller uses the per-step path)
static bool lp_pick(int Hl, bool backward, int* CL, size_t* smem) {
  for (int cl = 4; cl <= 8; cl *= 2) {
    if (Hl % cl) continue;
    const int U = Hl / cl;
    if (U > LP_THREADS || LP_THREADS % U) continue;
    const size_t need = backward ? ((size_t)4 * Hl * U + (size_t)2 * LP_RB * 4 * Hl + (size_t)LP_RB * U) * sizeof(float)
                                 : ((size_t)Hl * 4 * U + (size_t)2 * LP_RB * Hl) * sizeof(float);
    if (need <= 200 * 1024) { *CL = cl; *smem = need; return true; }
  }
  return false;
}

bool lstm_persistent_supported(int Hl) {
  int cl; size_t sm;
  return lp_pick(Hl, false, &cl, &sm) && lp_pick(Hl, true, &cl, &sm);
}

static int lp_launch(const void* kernel, LstmSeqArgs& a, bool backward, cudaStream_t st) {
  int CL; size_t smem;
  if (!lp_pick(a.Hl, backward, &CL, &smem)) { geeco_set_error("lstm persistent: dim_h_lstm %d not supported", a.Hl); return GEECO_ERR_INVALID; }
  a.U = a.Hl / CL;
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, (a.N + LP_RB - 1) / LP_RB); cfg.blockDim = dim3(LP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = geeco_pdl_enabled() ? 2 : 1;
  void* args[] = {(void*)&a};
  CUDA_TRY(cudaLaunchKernelExC(&cfg, kernel, args));
  geeco_count_launch(1);
  return GEECO_OK;
}

int launch_lstm_seq_fwd(int T, int N, int Hl, int xdim, const float* kernel, const float* c0, const float* m0,
                        const unsigned char* reset, float* gates, float* c, float* m, float* states, float* state_out,
                        cudaStream_t st) {
  LstmSeqArgs a = {};
  a.T = T; a.N = N; a.Hl = Hl; a.xdim = xdim; a.ld = xdim + Hl;
  a.Wh = kernel + (long long)xdim * 4 * Hl; a.c0 = c0; a.m0 = m0; a.reset = reset;
  a.gates = gates; a.c = c; a.m = m; a.states = states; a.state_out = state_out;
  return lp_launch((const void*)lstm_seq_fwd_kernel, a, false, st);
}

int launch_lstm_seq_bwd(int T, int N, int Hl, int xdim, const float* kernel, const float* c0, const unsigned char* reset,
                        const float* gates, const float* c, const float* dm_last, float* dgates, cudaStream_t st) {
  LstmSeqArgs a = {};
  a.T = T; a.N = N; a.Hl = Hl; a.xdim = xdim; a.ld = xdim + Hl;
  a.Wh = kernel + (long long)xdim * 4 * Hl; a.c0 = c0; a.reset = reset;
  a.gates = const_cast<float*>(gates); a.c = const_cast<float*>(c); a.dm_last = dm_last; a.dgates = dgates;
  return lp_launch((const void*)lstm_seq_bwd_kernel, a, true, st);
}
