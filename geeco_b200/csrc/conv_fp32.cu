// fp32 gather-GEMM kernels (CUDA cores) -- the full-precision arithmetic mode of the library.
//
// One geometry description (GatherGeom, common.cuh) covers every dense contraction of the
// e2evmc graph:
//   * conv forward  (graph.py:76-115, tf.layers.conv2d 3x3, SAME, stride 1/2, bias, ReLU)
//   * conv data-gradient, one launch per input-pixel parity class (autodiff of the above,
//     estimator.py:243-244), fused with the ReLU mask of the producing layer
//   * conv / dense weight-gradient with the bias gradient as a by-product (deterministic split-K)
//   * the LSTM gate GEMM and its two gradients (graph.py:217-225)
//
// NN:  C[m][n] = sum_k A[m][k] * B[k][n]      A gathered from an NHWC tensor (implicit im2col)
// TN:  C[k][n] = sum_m A[m][k] * G[m][n]      same gather, reduction over pixels
//
// Tiles: 128 x (16*TN) x 16, 256 threads, 8 x TN accumulators per thread, register-prefetched
// double buffering.  fp32 accumulate in a fixed order -> bitwise run-to-run reproducible.
#include "common.cuh"

static constexpr int GM_THREADS = 256;
static constexpr int GM_BM = 128;
static constexpr int GM_BK = 16;

struct RowInfo {
  int valid;
  int ys, xs;          // y*sy, x*sx
  long long pixbase;   // img * Hs * Ws
};

__device__ __forceinline__ RowInfo decode_row(const GatherGeom& g, int group, long long m, long long Mg) {
  RowInfo r;
  r.valid = m < Mg;
  const int hw = g.Hm * g.Wm;
  const long long mm = r.valid ? m : 0;
  const int img = (int)(mm / hw);
  const int rem = (int)(mm - (long long)img * hw);
  const int y = rem / g.Wm, x = rem - y * g.Wm;
  r.ys = y * g.sy; r.xs = x * g.sx;
  r.pixbase = ((long long)group * g.imgs_per_group + img) * g.Hs * g.Ws;
  return r;
}

__device__ __forceinline__ float4 gather4(const GatherGeom& g, const float* __restrict__ src, const RowInfo& r, int k,
                                          int Ktot) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r.valid && k < Ktot) {
    const int tap = k / g.Cs, c = k - tap * g.Cs;
    const int iy = r.ys + g.dy[tap], ix = r.xs + g.dx[tap];
    if (iy >= 0 && iy < g.Hs && ix >= 0 && ix < g.Ws)
      v = __ldg(reinterpret_cast<const float4*>(src + ((r.pixbase + (long long)iy * g.Ws + ix) * g.Cs + c)));
  }
  return v;
}

// ------------------------------------------------------------------------------------------
// NN kernel
// ------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(GM_THREADS) gemm_nn_f32_kernel(const GatherGeom g, const float* __restrict__ src,
                                                                 const float* __restrict__ Ball,
                                                                 const float* __restrict__ bias_all,
                                                                 const float* __restrict__ mask,
                                                                 float* __restrict__ dst, int epi) {
  pdl_enter();
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[2][GM_BK][GM_BM + 4];
  __shared__ __align__(16) float Bs[2][GM_BK][BN + 4];
  const int tid = threadIdx.x;
  const int group = blockIdx.y;
  const int n0 = blockIdx.z * BN;
  const long long m0 = (long long)blockIdx.x * GM_BM;
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  const int Ktot = g.ntaps * g.Cs;
  const int nk = (Ktot + GM_BK - 1) / GM_BK;
  const float* __restrict__ B = Ball + (long long)group * g.b_group_stride;

  // A loader: 2 rows per thread, one float4 (4 consecutive k) each
  const int a_kq = tid & 3, a_r0 = tid >> 2;
  RowInfo ri[2];
  ri[0] = decode_row(g, group, m0 + a_r0, Mg);
  ri[1] = decode_row(g, group, m0 + a_r0 + 64, Mg);
  // B loader
  constexpr int B_TASKS = 4 * BN;   // float4 tasks per tile
  const bool b_active = tid < B_TASKS;
  int b_k, b_n;                     // local k (multiple of 4 for trans) and local n
  if (!g.transB) { b_k = tid / (BN / 4); b_n = (tid % (BN / 4)) * 4; }
  else           { b_n = tid % BN;       b_k = (tid / BN) * 4; }

  float4 ra[2], rb;
  auto load_regs = [&](int kt) {
    const int k = kt * GM_BK + a_kq * 4;
    ra[0] = gather4(g, src, ri[0], k, Ktot);
    ra[1] = gather4(g, src, ri[1], k, Ktot);
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_active) {
      const int kk = kt * GM_BK + b_k;
      if (kk < Ktot) {
        const int tap = kk / g.Cs, c = kk - tap * g.Cs;
        if (!g.transB) {
          if (c < g.Cw && n0 + b_n < g.Nn)
            rb = __ldg(reinterpret_cast<const float4*>(B + (long long)(g.wbase[tap] + c) * g.ldb + n0 + b_n));
        } else {
          if (c < g.Cw && n0 + b_n < g.Nn)
            rb = __ldg(reinterpret_cast<const float4*>(B + g.wbase[tap] + (long long)(n0 + b_n) * g.ldb + c));
        }
      }
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = a_r0 + 64 * j;
      As[buf][a_kq * 4 + 0][row] = ra[j].x; As[buf][a_kq * 4 + 1][row] = ra[j].y;
      As[buf][a_kq * 4 + 2][row] = ra[j].z; As[buf][a_kq * 4 + 3][row] = ra[j].w;
    }
    if (b_active) {
      if (!g.transB) {
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
      } else {
        Bs[buf][b_k + 0][b_n] = rb.x; Bs[buf][b_k + 1][b_n] = rb.y;
        Bs[buf][b_k + 2][b_n] = rb.z; Bs[buf][b_k + 3][b_n] = rb.w;
      }
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_regs(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_regs(kt + 1);
#pragma unroll
    for (int kk = 0; kk < GM_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_smem(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  const float* __restrict__ bias = bias_all ? bias_all + (long long)group * g.bias_group_stride : nullptr;
  const int hw = g.Hm * g.Wm;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= Mg) continue;
    const int img = (int)(m / hw);
    const int rem = (int)(m - (long long)img * hw);
    const int y = rem / g.Wm, x = rem - y * g.Wm;
    const long long pix = (((long long)group * g.imgs_per_group + img) * g.Hd + (y * g.dsy + g.dy0)) * g.Wd +
                          (x * g.dsx + g.dx0);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.Nn) continue;
      const long long off = pix * g.Nn + n;
      float v = acc[i][j];
      if (epi == EPI_BIAS) v += bias[n];
      else if (epi == EPI_BIAS_RELU) v = fmaxf(v + bias[n], 0.f);
      else if (epi == EPI_MASK) v = mask[off] > 0.f ? v : 0.f;
      dst[off] = v;
    }
  }
}

int launch_gemm_nn_f32(const GatherGeom& g, const float* src, const float* B, const float* bias, const float* mask,
                       float* dst, int groups, int epi, cudaStream_t st) {
  if (g.Cs % 4 || g.ldb % 4 || g.Nn % 4 || (g.transB && g.Cw % 4)) {
    geeco_set_error("gemm_nn_f32: Cs=%d ldb=%d Nn=%d Cw=%d must be multiples of 4", g.Cs, g.ldb, g.Nn, g.Cw);
    return GEECO_ERR_INVALID;
  }
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg <= 0) return GEECO_OK;
  const int TN = g.Nn <= 32 ? 2 : (g.Nn <= 48 ? 3 : 4);
  dim3 grid(ceil_div(Mg, GM_BM), groups, ceil_div(g.Nn, 16 * TN));
  if (TN == 2) GEECO_LAUNCH((gemm_nn_f32_kernel<2>), grid, GM_THREADS, 0, st, g, src, B, bias, mask, dst, epi);
  else if (TN == 3) GEECO_LAUNCH((gemm_nn_f32_kernel<3>), grid, GM_THREADS, 0, st, g, src, B, bias, mask, dst, epi);
  else GEECO_LAUNCH((gemm_nn_f32_kernel<4>), grid, GM_THREADS, 0, st, g, src, B, bias, mask, dst, epi);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

// ------------------------------------------------------------------------------------------
// TN kernel (weight gradient, deterministic split over the pixel dimension)
//   grid.x = tile of 128 reduction-index values k=(tap,c), grid.y = group*splits + split, grid.z = n tile
// ------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(GM_THREADS) gemm_tn_f32_kernel(const GatherGeom g, const float* __restrict__ src,
                                                                 const float* __restrict__ G,
                                                                 float* __restrict__ partial,
                                                                 float* __restrict__ bias_partial, int splits,
                                                                 long long Mchunk, int Krows) {
  pdl_enter();
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[2][GM_BK][GM_BM + 4];
  __shared__ __align__(16) float Gs[2][GM_BK][BN + 4];
  const int tid = threadIdx.x;
  const int group = blockIdx.y / splits, split = blockIdx.y - group * splits;
  const int k0 = blockIdx.x * GM_BM;
  const int n0 = blockIdx.z * BN;
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  const long long mlo = (long long)split * Mchunk;
  long long mhi = mlo + Mchunk; if (mhi > Mg) mhi = Mg;
  const int Ktot = g.ntaps * g.Cs;
  const int nit = mhi > mlo ? (int)((mhi - mlo + GM_BK - 1) / GM_BK) : 0;

  // A loader: fixed k (4 consecutive), 2 pixel rows per iteration
  const int a_kq = tid & 31, a_r0 = tid >> 5;
  const int a_k = k0 + a_kq * 4;
  const bool a_kvalid = a_k < Ktot;
  int a_tap = 0, a_c = 0;
  if (a_kvalid) { a_tap = a_k / g.Cs; a_c = a_k - a_tap * g.Cs; }
  const int a_dy = g.dy[a_tap], a_dx = g.dx[a_tap];
  // G loader
  constexpr int G_TASKS = 4 * BN;
  const bool g_active = tid < G_TASKS;
  const int g_r = tid / (BN / 4), g_n = (tid % (BN / 4)) * 4;
  const long long g_rowbase = (long long)group * Mg;

  float4 ra[2], rg;
  auto load_regs = [&](int it) {
    const long long mb = mlo + (long long)it * GM_BK;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const long long m = mb + a_r0 + 8 * j;
      if (a_kvalid && m < mhi) {
        RowInfo r = decode_row(g, group, m, Mg);
        const int iy = r.ys + a_dy, ix = r.xs + a_dx;
        if (iy >= 0 && iy < g.Hs && ix >= 0 && ix < g.Ws)
          ra[j] = __ldg(reinterpret_cast<const float4*>(src + ((r.pixbase + (long long)iy * g.Ws + ix) * g.Cs + a_c)));
      }
    }
    rg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g_active) {
      const long long m = mb + g_r;
      if (m < mhi && n0 + g_n < g.Nn)
        rg = __ldg(reinterpret_cast<const float4*>(G + (g_rowbase + m) * g.Nn + n0 + g_n));
    }
  };
  auto store_smem = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][a_r0][a_kq * 4]) = ra[0];
    *reinterpret_cast<float4*>(&As[buf][a_r0 + 8][a_kq * 4]) = ra[1];
    if (g_active) *reinterpret_cast<float4*>(&Gs[buf][g_r][g_n]) = rg;
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][TN], bacc[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    bacc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][j] = 0.f;
  }
  const bool do_bias = (blockIdx.x == 0) && (ty == 0) && (bias_partial != nullptr);

  if (nit > 0) {
    load_regs(0);
    store_smem(0);
  }
  __syncthreads();
  for (int it = 0; it < nit; ++it) {
    const int buf = it & 1;
    if (it + 1 < nit) load_regs(it + 1);
#pragma unroll
    for (int kk = 0; kk < GM_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Gs[buf][kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < TN; ++j) bacc[j] += b[j];
      }
    }
    if (it + 1 < nit) store_smem(buf ^ 1);
    __syncthreads();
  }

  float* __restrict__ P = partial + (long long)blockIdx.y * Krows * g.Nn;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + ty * 8 + i;
    if (k >= Ktot) continue;
    const int tap = k / g.Cs, c = k - tap * g.Cs;
    if (c >= g.Cw) continue;
    const long long wr = g.wbase[tap] + c;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.Nn) P[wr * g.Nn + n] = acc[i][j];
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.Nn) bias_partial[(long long)blockIdx.y * g.Nn + n] = bacc[j];
    }
  }
}

// out[g][i] = sum_s partial[(g*splits + s)][i]   (fixed order -> deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits,
                                       long long count, long long out_group_stride, int groups) {
  pdl_enter();
  const long long total = count * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int grp = (int)(i / count);
    const long long e = i - (long long)grp * count;
    const float* p = partial + (long long)grp * splits * count + e;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += p[(long long)k * count];
    out[(long long)grp * out_group_stride + e] = s;
  }
}

static long long tn_splits(const GatherGeom& g, int groups) {
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  const int Ktot = g.ntaps * g.Cs;
  const int TN = g.Nn <= 32 ? 2 : (g.Nn <= 48 ? 3 : 4);
  const int ktiles = ceil_div(Ktot, GM_BM), ntiles = ceil_div(g.Nn, 16 * TN);
  const long long tiles = (long long)ktiles * ntiles * groups;
  long long splits = (2 * 148 + tiles - 1) / tiles;
  const long long max_splits = Mg / 64 > 1 ? Mg / 64 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

long long gemm_tn_partial_floats(const GatherGeom& g, int groups) {
  return tn_splits(g, groups) * groups * ((long long)g.ntaps * g.Cw * g.Nn + g.Nn);
}

int launch_gemm_tn_f32(const GatherGeom& g, const float* src, const float* G, float* dW, float* dbias, float* partial,
                       long long partial_cap_floats, int groups, long long dw_group_stride,
                       long long dbias_group_stride, cudaStream_t st) {
  if (g.Cs % 4 || g.Nn % 4) {
    geeco_set_error("gemm_tn_f32: Cs=%d Nn=%d must be multiples of 4", g.Cs, g.Nn);
    return GEECO_ERR_INVALID;
  }
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg <= 0) return GEECO_OK;
  const int Ktot = g.ntaps * g.Cs, Krows = g.ntaps * g.Cw;
  const int TN = g.Nn <= 32 ? 2 : (g.Nn <= 48 ? 3 : 4);
  const int ktiles = ceil_div(Ktot, GM_BM), ntiles = ceil_div(g.Nn, 16 * TN);
  long long splits = tn_splits(g, groups);
  const long long per_split = (long long)Krows * g.Nn + g.Nn;
  while (splits > 1 && splits * groups * per_split > partial_cap_floats) --splits;
  if (!partial || splits * groups * per_split > partial_cap_floats) {
    geeco_set_error("gemm_tn_f32: partial buffer too small (%lld floats needed)", splits * groups * per_split);
    return GEECO_ERR_WORKSPACE;
  }
  long long Mchunk = (Mg + splits - 1) / splits;
  Mchunk = (Mchunk + GM_BK - 1) / GM_BK * GM_BK;
  float* bias_partial = dbias ? partial + splits * groups * (long long)Krows * g.Nn : nullptr;
  dim3 grid(ktiles, (unsigned)(groups * splits), ntiles);
  if (TN == 2) GEECO_LAUNCH((gemm_tn_f32_kernel<2>), grid, GM_THREADS, 0, st, g, src, G, partial, bias_partial, (int)splits, Mchunk, Krows);
  else if (TN == 3) GEECO_LAUNCH((gemm_tn_f32_kernel<3>), grid, GM_THREADS, 0, st, g, src, G, partial, bias_partial, (int)splits, Mchunk, Krows);
  else GEECO_LAUNCH((gemm_tn_f32_kernel<4>), grid, GM_THREADS, 0, st, g, src, G, partial, bias_partial, (int)splits, Mchunk, Krows);
  const long long cnt = (long long)Krows * g.Nn;
  int rb = ceil_div(cnt * groups, 256); if (rb > 1184) rb = 1184;
  GEECO_LAUNCH((reduce_partials_kernel), rb, 256, 0, st, partial, dW, (int)splits, cnt, dw_group_stride, groups);
  geeco_count_launch(2);
  if (dbias) {
    GEECO_LAUNCH((reduce_partials_kernel), ceil_div((long long)g.Nn * groups, 256), 256, 0, st, bias_partial, dbias, (int)splits,
                                                                                  g.Nn, dbias_group_stride, groups);
    geeco_count_launch(1);
  }
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
