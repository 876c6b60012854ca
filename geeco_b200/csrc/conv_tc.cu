// bf16 implicit-GEMM convolution kernels on the Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// Reference ops: tf.layers.conv2d 3x3/SAME of conv_encoder (src/models/e2evmc/graph.py:76-115) and
// the data / weight gradients TensorFlow derives for it (estimator.py:243-244).
//
// tc_nn_kernel   (forward, data-gradient):  D[m][n] = sum_k A[m][k] * Wp[n][k]
//   A  : implicit im2col rows (one output pixel each), gathered from the NHWC bf16 activation with
//        16-byte cp.async into a 128x64 K-major SWIZZLE_128B shared-memory tile (zero-fill = padding)
//   Wp : pre-packed bf16 weights [N][Kpad], streamed by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B)
//   D  : fp32 accumulators in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of
//        tile i+1; epilogue = tcgen05.ld -> bias+ReLU (fwd) or ReLU-mask (dgrad) -> bf16 NHWC store
//   Persistent CTAs, warp roles: 0-3 gather producers, 4-7 epilogue, 8 MMA issuer (+TMEM alloc), 9 TMA.
//
// tc_wgrad_kernel (weight gradient):  dW^T[co][(tap,ci)] = sum_pixels G[p][co] * im2col[p][(tap,ci)]
//   both operands are gathered as MN-major SWIZZLE_128B tiles (64 pixels x 64 values per sub-tile);
//   M = 128 output channels, N <= 256 reduction-index values per CTA, split over pixel ranges with a
//   deterministic second-stage reduction; the bias gradient rides along as a column of ones when the
//   reduction index has padding to spare.
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <string.h>

using namespace tc;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int LOOKAHEAD = 2;
constexpr int NN_THREADS = 320;
constexpr int WG_THREADS = 288;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB

struct RowDec {
  int pix;      // (group*ipg + img) * Hs * Ws
  int ys, xs;   // y*sy, x*sx  (ys = -(1<<20) marks an invalid row -> every tap is out of bounds)
};

__device__ __forceinline__ RowDec decode_row_tc(const TcGeom& g, int group, long long m, long long Mg) {
  RowDec r;
  if (m < Mg) {
    const int hw = g.Hm * g.Wm;
    const int img = (int)(m / hw);
    const int rem = (int)(m - (long long)img * hw);
    const int y = rem / g.Wm, x = rem - y * g.Wm;
    r.pix = (group * g.imgs_per_group + img) * g.Hs * g.Ws;
    r.ys = y * g.sy; r.xs = x * g.sx;
  } else {
    r.pix = 0; r.ys = -(1 << 20); r.xs = 0;
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// forward / data-gradient kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NN_THREADS, 1)
tc_nn_kernel(const TcGeom g, const __grid_constant__ CUtensorMap wmap, const __nv_bfloat16* __restrict__ src,
             const float* __restrict__ bias_all, const __nv_bfloat16* __restrict__ mask,
             __nv_bfloat16* __restrict__ dst, float* __restrict__ dst_f32, int epi, int tiles_per_group,
             int total_tiles, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = g.Nn;
  const int b_stage_bytes = BN * BK * 2;
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + STAGES * b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 128 + 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
  if (warp == 9 && lane == 0) tma_prefetch_desc(&wmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int nkb = g.Kpad / BK;
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;

  if (warp < 4) {
    // ===================== A producers: implicit-im2col gather =====================
    const int chunk = lane & 7, rsub = lane >> 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int group = tile / tiles_per_group;
      const long long m0 = (long long)(tile - group * tiles_per_group) * BM;
      RowDec rd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rd[i] = decode_row_tc(g, group, m0 + warp * 32 + i * 4 + rsub, Mg);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        const int k = kb * BK + chunk * 8;
        const bool kvalid = k < g.Ktot;
        const int tap = kvalid ? k / g.Cs : 0;
        const int ch = k - tap * g.Cs;
        const int dyt = g.dy[tap], dxt = g.dx[tap];
        const uint32_t a_s = smem_u32(a_base + s * A_STAGE_BYTES);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = warp * 32 + i * 4 + rsub;
          const int iy = rd[i].ys + dyt, ix = rd[i].xs + dxt;
          const bool ok = kvalid && (unsigned)iy < (unsigned)g.Hs && (unsigned)ix < (unsigned)g.Ws;
          const __nv_bfloat16* p = ok ? src + ((long long)(rd[i].pix + iy * g.Ws + ix) * g.Cs + ch) : src;
          cp_async16(a_s + row * 128 + ((chunk ^ (row & 7)) << 4), p, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (it >= LOOKAHEAD) {
          cp_async_wait<LOOKAHEAD>();
          fence_proxy_async();
          mbar_arrive(&full[(it - LOOKAHEAD) % STAGES]);
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    const uint32_t first = it >= LOOKAHEAD ? it - LOOKAHEAD : 0;
    for (uint32_t j = first; j < it; ++j) mbar_arrive(&full[j % STAGES]);
  } else if (warp < 8) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hw = g.Hm * g.Wm;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      const int group = tile / tiles_per_group;
      const long long m = (long long)(tile - group * tiles_per_group) * BM + row;
      const int buf = tl & 1;
      mbar_wait(&tmem_full[buf], (tl >> 1) & 1);
      tc_fence_after();
      const bool valid = m < Mg;
      long long off = 0;
      if (valid) {
        const int img = (int)(m / hw);
        const int rem = (int)(m - (long long)img * hw);
        const int y = rem / g.Wm, x = rem - y * g.Wm;
        off = ((((long long)group * g.imgs_per_group + img) * g.Hd + (y * g.dsy + g.dy0)) * g.Wd + (x * g.dsx + g.dx0)) * g.Nn;
      }
      const float* bias = bias_all ? bias_all + (long long)group * g.bias_group_stride : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (epi == TC_EPI_BIAS_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i] + __ldg(bias + c0 + i), 0.f);
          } else if (epi == TC_EPI_BIAS) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += __ldg(bias + c0 + i);
          } else if (epi == TC_EPI_MASK) {
            const uint4 m0v = __ldg(reinterpret_cast<const uint4*>(mask + off + c0));
            const uint4 m1v = __ldg(reinterpret_cast<const uint4*>(mask + off + c0 + 8));
            const uint32_t mw[8] = {m0v.x, m0v.y, m0v.z, m0v.w, m1v.x, m1v.y, m1v.z, m1v.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // post-ReLU activations are >= 0: "y > 0" == magnitude bits non-zero and sign clear
              const uint32_t lo = mw[i] & 0xffffu, hi = mw[i] >> 16;
              if (!((lo & 0x7fffu) != 0 && (lo & 0x8000u) == 0)) f[2 * i] = 0.f;
              if (!((hi & 0x7fffu) != 0 && (hi & 0x8000u) == 0)) f[2 * i + 1] = 0.f;
            }
          }
          if (dst) {
            uint4 o0, o1;
            o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]);
            o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]);
            o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
            *reinterpret_cast<uint4*>(dst + off + c0) = o0;
            *reinterpret_cast<uint4*>(dst + off + c0 + 8) = o1;
          }
          if (dst_f32) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<float4*>(dst_f32 + off + c0 + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const int buf = tl & 1;
        mbar_wait(&tmem_empty[buf], ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_base + s * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(b_base + s * b_stage_bytes);
#pragma unroll
          for (int j = 0; j < BK / 16; ++j)
            tc_mma(d, make_desc_sw128(a_addr + j * 32, 16, 1024), make_desc_sw128(b_addr + j * 32, 16, 1024), idesc,
                   (kb | j) != 0 ? 1u : 0u);
          tc_commit(&empty[s]);
        }
        tc_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===================== weight tiles by TMA (one thread) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int group = tile / tiles_per_group;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], (uint32_t)b_stage_bytes);
          tma_load_2d(smem_u32(b_base + s * b_stage_bytes), &wmap, &full[s], kb * BK, group * g.b_rows_per_group);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// weight-gradient kernel
//   grid.x = mtile * n_chunks + nchunk ; grid.y = split ; grid.z = group
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const TcGeom g, const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ G,
                float* __restrict__ partial, int Cout, int n_chunks, int kb_per_split, int total_kb, int Mrows_pad,
                int ones_col, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int mtile = blockIdx.x / n_chunks, nchunk = blockIdx.x - mtile * n_chunks;
  const int split = blockIdx.y, group = blockIdx.z, groups = gridDim.z;
  int nsub = (g.Kpad - nchunk * 256) / 64;
  if (nsub > 4) nsub = 4;
  constexpr int SUB = 64 * 128;                     // one 64-pixel x 128-byte sub-tile
  const int stage_bytes = (2 + 4) * SUB;            // G: 2 sub-tiles, im2col: up to 4
  uint8_t* st_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  const int kb_lo = split * kb_per_split;
  int kb_hi = kb_lo + kb_per_split;
  if (kb_hi > total_kb) kb_hi = total_kb;
  const int nkb = kb_hi - kb_lo;

  if (warp < 4) {
    const int chunk = lane & 7, rsub = lane >> 3;
    // per-thread constants of the im2col columns it copies
    int kv[4], kdy[4], kdx[4], kch[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = nchunk * 256 + j * 64 + chunk * 8;
      kv[j] = (j < nsub && k < g.Ktot) ? 1 : 0;
      if (j < nsub && k == ones_col) kv[j] = 2;
      const int tap = kv[j] == 1 ? k / g.Cs : 0;
      kch[j] = k - tap * g.Cs;
      kdy[j] = g.dy[tap]; kdx[j] = g.dx[tap];
    }
    const long long grow = (long long)group * Mg;
    for (int it = 0; it < nkb; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
      const uint32_t sb = smem_u32(st_base + s * stage_bytes);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = warp * 16 + i * 4 + rsub;
        const long long m = (long long)(kb_lo + it) * 64 + r;
        const RowDec rd = decode_row_tc(g, group, m, Mg);
        const bool rvalid = m < Mg;
        const uint32_t roff = r * 128 + ((chunk ^ (r & 7)) << 4);
        // G tile: rows = pixels, 128 output channels of this M tile
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = mtile * 128 + j * 64 + chunk * 8;
          const bool ok = rvalid && c < Cout;
          const __nv_bfloat16* p = ok ? G + ((grow + m) * Cout + c) : G;
          cp_async16(sb + j * SUB + roff, p, ok ? 16u : 0u);
        }
        // im2col tile: rows = pixels, 64 reduction-index values per sub-tile
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j >= nsub) break;
          const uint32_t d = sb + (2 + j) * SUB + roff;
          if (kv[j] == 2) {
            // bias-gradient column: 1.0 in the first padding column of valid pixels
            const uint32_t one = rvalid ? 0x00003f80u : 0u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(d), "r"(one), "r"(0u) : "memory");
          } else {
            const int iy = rd.ys + kdy[j], ix = rd.xs + kdx[j];
            const bool ok = kv[j] == 1 && (unsigned)iy < (unsigned)g.Hs && (unsigned)ix < (unsigned)g.Ws;
            const __nv_bfloat16* p = ok ? src + ((long long)(rd.pix + iy * g.Ws + ix) * g.Cs + kch[j]) : src;
            cp_async16(d, p, ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
      if (it >= LOOKAHEAD) {
        cp_async_wait<LOOKAHEAD>();
        fence_proxy_async();
        mbar_arrive(&full[(it - LOOKAHEAD) % STAGES]);
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int j = nkb >= LOOKAHEAD ? nkb - LOOKAHEAD : 0; j < nkb; ++j) mbar_arrive(&full[j % STAGES]);
  } else if (warp < 8) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int co = mtile * 128 + row;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* P = partial + (((long long)split * groups + group) * Mrows_pad + co) * g.Kpad + nchunk * 256;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < nsub * 64; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if (co < Cout) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(P + c0 + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, nsub * 64, 1, 1);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(st_base + s * stage_bytes);
        const uint32_t b_addr = a_addr + 2 * SUB;
#pragma unroll
        for (int j = 0; j < 4; ++j)   // 4 x 16 pixels
          tc_mma(tmem_base, make_desc_sw128(a_addr + j * 2048, SUB, 1024), make_desc_sw128(b_addr + j * 2048, SUB, 1024),
                 idesc, (it | j) != 0 ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// dW[g][(tap*Cw + ch)][co] = sum_splits partial[s][g][co][tap*Cs + ch] ; bias from the ones column
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ dbias,
                                    int splits, int groups, int Mrows_pad, int Kpad, int Cout, int Cs, int Cw,
                                    int Ktot, int ones_col, long long dw_group_stride, long long dbias_group_stride) {
  const long long per_group = (long long)(Ktot + 1) * Cout;
  const long long total = per_group * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int grp = (int)(i / per_group);
    const long long e = i - (long long)grp * per_group;
    const int k = (int)(e / Cout), co = (int)(e - (long long)k * Cout);
    int col;
    float* out;
    if (k < Ktot) {
      const int tap = k / Cs, ch = k - tap * Cs;
      if (ch >= Cw) continue;
      col = k;
      out = dW + (long long)grp * dw_group_stride + (long long)(tap * Cw + ch) * Cout + co;
    } else {
      if (ones_col < 0 || !dbias) continue;
      col = ones_col;
      out = dbias + (long long)grp * dbias_group_stride + co;
    }
    const float* p = partial + ((long long)grp * Mrows_pad + co) * Kpad + col;
    const long long sstride = (long long)groups * Mrows_pad * Kpad;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += p[sp * sstride];
    *out = s;
  }
}

// column sums of a bf16 [rows][C] matrix per group (bias gradient when no padding column exists);
// stage 1: grid (chunks, groups) -> part[g][chunk][C]; stage 2 reduces chunks in order.
__global__ void colsum_bf16_stage1(const __nv_bfloat16* __restrict__ G, float* __restrict__ part, long long rows_per_group,
                                   int C, int chunks) {
  const int grp = blockIdx.y, chunk = blockIdx.x;
  const long long per = (rows_per_group + chunks - 1) / chunks;
  const long long lo = chunk * per;
  long long hi = lo + per; if (hi > rows_per_group) hi = rows_per_group;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    const __nv_bfloat16* p = G + ((long long)grp * rows_per_group) * C + c;
    for (long long r = lo; r < hi; ++r) s += __bfloat162float(p[r * C]);
    part[((long long)grp * chunks + chunk) * C + c] = s;
  }
}
__global__ void colsum_stage2(const float* __restrict__ part, float* __restrict__ out, int C, int chunks,
                              long long out_group_stride) {
  const int grp = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += part[((long long)grp * chunks + k) * C + c];
    out[(long long)grp * out_group_stride + c] = s;
  }
}

__global__ void pack_weights_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int mode, int groups,
                                    long long w_group_stride, int Cin, int Cout, int Cs, int ntaps, int rows, int Kpad,
                                    int t0, int t1, int t2, int t3, int t4, int t5, int t6, int t7, int t8) {
  const int taps[9] = {t0, t1, t2, t3, t4, t5, t6, t7, t8};
  const long long total = (long long)groups * rows * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    const long long rr = i / Kpad;
    const int r = (int)(rr % rows), grp = (int)(rr / rows);
    float v = 0.f;
    const int per_tap = mode == 0 ? Cs : Cout;
    const int t = k / per_tap, c = k - t * per_tap;
    if (t < ntaps) {
      const float* Wg = W + (long long)grp * w_group_stride + (long long)taps[t] * Cin * Cout;
      if (mode == 0) { if (c < Cin && r < Cout) v = Wg[(long long)c * Cout + r]; }      // row = output channel
      else           { if (r < Cin) v = Wg[(long long)r * Cout + c]; }                   // row = input channel
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_weight_tensor_map(CUtensorMap* map, const void* base, long long rows_total, int Kpad, int box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      geeco_set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
      return GEECO_ERR_CUDA;
    }
    fn = (PFN_encodeTiled)p;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows_total};
  cuuint64_t gstr[1] = {(cuuint64_t)Kpad * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    geeco_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld, Kpad %d, box %d)", (int)r, rows_total, Kpad, box_rows);
    return GEECO_ERR_CUDA;
  }
  return GEECO_OK;
}

static void same_pad_tc(int in, int s, int* out, int* before) {
  *out = (in + s - 1) / s;
  int total = (*out - 1) * s + 3 - in;
  if (total < 0) total = 0;
  *before = total / 2;
}

TcGeom tc_fwd_geom(int H, int W, int Cs, int Cout, int stride, int imgs_per_group, int groups) {
  TcGeom g;
  memset(&g, 0, sizeof(g));
  int Ho, Wo, pt, pl;
  same_pad_tc(H, stride, &Ho, &pt);
  same_pad_tc(W, stride, &Wo, &pl);
  g.Hs = H; g.Ws = W; g.Cs = Cs; g.Hm = Ho; g.Wm = Wo; g.sy = stride; g.sx = stride; g.ntaps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) { g.dy[ky * 3 + kx] = ky - pt; g.dx[ky * 3 + kx] = kx - pl; }
  g.Ktot = 9 * Cs; g.Kpad = (g.Ktot + 63) / 64 * 64;
  g.Nn = Cout; g.Hd = Ho; g.Wd = Wo; g.dsy = 1; g.dsx = 1;
  g.imgs_per_group = imgs_per_group; g.groups = groups; g.b_rows_per_group = Cout; g.bias_group_stride = Cout;
  return g;
}

bool tc_dgrad_geom(int H, int W, int Cin, int Cout, int stride, int py, int px, int imgs_per_group, int groups,
                   TcGeom* out, int* taps_out) {
  TcGeom g;
  memset(&g, 0, sizeof(g));
  int Ho, Wo, pt, pl;
  same_pad_tc(H, stride, &Ho, &pt);
  same_pad_tc(W, stride, &Wo, &pl);
  g.Hs = Ho; g.Ws = Wo; g.Cs = Cout;
  g.Hm = (H - py + stride - 1) / stride; g.Wm = (W - px + stride - 1) / stride;
  if (g.Hm <= 0 || g.Wm <= 0) return false;
  g.sy = 1; g.sx = 1;
  int nt = 0;
  for (int ky = 0; ky < 3; ++ky) {
    const int ny = py + pt - ky;
    if (((ny % stride) + stride) % stride) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int nx = px + pl - kx;
      if (((nx % stride) + stride) % stride) continue;
      g.dy[nt] = (ny >= 0) ? ny / stride : -((-ny + stride - 1) / stride);
      g.dx[nt] = (nx >= 0) ? nx / stride : -((-nx + stride - 1) / stride);
      taps_out[nt] = ky * 3 + kx;
      ++nt;
    }
  }
  if (!nt) return false;
  g.ntaps = nt;
  g.Ktot = nt * Cout; g.Kpad = (g.Ktot + 63) / 64 * 64;
  g.Nn = Cin; g.Hd = H; g.Wd = W; g.dsy = stride; g.dsx = stride; g.dy0 = py; g.dx0 = px;
  g.imgs_per_group = imgs_per_group; g.groups = groups; g.b_rows_per_group = Cin; g.bias_group_stride = 0;
  *out = g;
  return true;
}

static int next_pow2_cols(int c) {
  int p = 32;
  while (p < c) p *= 2;
  return p;
}

int launch_tc_nn(const TcGeom& g, const CUtensorMap* wmap, const __nv_bfloat16* src, const float* bias,
                 const __nv_bfloat16* mask, __nv_bfloat16* dst, float* dst_f32, int epi, int max_ctas,
                 cudaStream_t st) {
  if (g.Nn % 16 || g.Nn < 16 || g.Nn > 256 || g.Cs % 8 || g.Kpad % 64) {
    geeco_set_error("tc_nn: unsupported shape N=%d Cs=%d Kpad=%d", g.Nn, g.Cs, g.Kpad);
    return GEECO_ERR_INVALID;
  }
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg <= 0) return GEECO_OK;
  if ((long long)g.imgs_per_group * g.groups * g.Hs * g.Ws >= (1ll << 31)) {
    geeco_set_error("tc_nn: source tensor has too many pixels for 32-bit indexing");
    return GEECO_ERR_INVALID;
  }
  const int tiles_per_group = ceil_div(Mg, BM);
  const int total_tiles = tiles_per_group * g.groups;
  const size_t smem = 1024 + (size_t)STAGES * (A_STAGE_BYTES + g.Nn * BK * 2) + 256;
  const int tmem_cols = next_pow2_cols(2 * g.Nn);
  int per_sm = (int)(227 * 1024 / smem);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  if (per_sm * tmem_cols > 512) per_sm = 1;
  int ctas = num_sms() * per_sm;
  if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
  if (ctas > total_tiles) ctas = total_tiles;
  CUDA_TRY(cudaFuncSetAttribute(tc_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_nn_kernel<<<ctas, NN_THREADS, smem, st>>>(g, *wmap, src, bias, mask, dst, dst_f32, epi, tiles_per_group, total_tiles,
                                               tmem_cols);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

struct WgradPlan { int m_tiles, n_chunks, splits, kb_per_split, total_kb, Mrows_pad, ones_col; };

static WgradPlan wgrad_plan(const TcGeom& g, int Cout) {
  WgradPlan p;
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  p.m_tiles = (Cout + 127) / 128;
  p.n_chunks = (g.Kpad + 255) / 256;
  p.total_kb = (int)((Mg + 63) / 64);
  const int base = p.m_tiles * p.n_chunks * g.groups;
  int want = (num_sms() + base - 1) / base;
  if (want < 1) want = 1;
  if (want > p.total_kb) want = p.total_kb;
  p.kb_per_split = (p.total_kb + want - 1) / want;
  p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.Mrows_pad = p.m_tiles * 128;
  p.ones_col = g.Kpad > g.Ktot ? g.Ktot : -1;
  return p;
}

long long tc_wgrad_partial_floats(const TcGeom& g, int Cout) {
  WgradPlan p = wgrad_plan(g, Cout);
  long long main_part = (long long)p.splits * g.groups * p.Mrows_pad * g.Kpad;
  long long colsum_part = (long long)g.groups * 64 * Cout;
  return main_part + colsum_part + 64;
}

int launch_tc_wgrad(const TcGeom& g, int Cout, int Cw, const __nv_bfloat16* src, const __nv_bfloat16* G, float* dW,
                    float* dbias, float* partial, long long partial_cap, long long dw_group_stride,
                    long long dbias_group_stride, cudaStream_t st) {
  if (g.Cs % 8 || Cout % 8 || g.Kpad % 64) {
    geeco_set_error("tc_wgrad: unsupported shape Cs=%d Cout=%d Kpad=%d", g.Cs, Cout, g.Kpad);
    return GEECO_ERR_INVALID;
  }
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg <= 0) return GEECO_OK;
  WgradPlan p = wgrad_plan(g, Cout);
  const long long main_part = (long long)p.splits * g.groups * p.Mrows_pad * g.Kpad;
  if (!partial || tc_wgrad_partial_floats(g, Cout) > partial_cap) {
    geeco_set_error("tc_wgrad: partial buffer too small (%lld floats needed, %lld given)", tc_wgrad_partial_floats(g, Cout), partial_cap);
    return GEECO_ERR_WORKSPACE;
  }
  const size_t smem = 1024 + (size_t)STAGES * 6 * 64 * 128 + 256;
  CUDA_TRY(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p.m_tiles * p.n_chunks, p.splits, g.groups);
  tc_wgrad_kernel<<<grid, WG_THREADS, smem, st>>>(g, src, G, partial, Cout, p.n_chunks, p.kb_per_split, p.total_kb,
                                                  p.Mrows_pad, dbias ? p.ones_col : -1, 256);
  const long long total = (long long)(g.Ktot + 1) * Cout * g.groups;
  int rb = ceil_div(total, 256); if (rb > 148 * 8) rb = 148 * 8;
  wgrad_reduce_kernel<<<rb, 256, 0, st>>>(partial, dW, dbias, p.splits, g.groups, p.Mrows_pad, g.Kpad, Cout, g.Cs, Cw,
                                          g.Ktot, dbias ? p.ones_col : -1, dw_group_stride, dbias_group_stride);
  geeco_count_launch(2);
  if (dbias && p.ones_col < 0) {
    float* part = partial + main_part;
    const int chunks = Mg >= 64 * 8 ? 64 : 1;
    colsum_bf16_stage1<<<dim3(chunks, g.groups), 256, 0, st>>>(G, part, Mg, Cout, chunks);
    colsum_stage2<<<g.groups, 256, 0, st>>>(part, dbias, Cout, chunks, dbias_group_stride);
    geeco_count_launch(2);
  }
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

int launch_pack_weights(const float* W, __nv_bfloat16* out, int mode, int groups, long long w_group_stride, int Cin,
                        int Cout, int Cs, int ntaps, const int* taps, int rows, int Kpad, cudaStream_t st) {
  int t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < ntaps && i < 9; ++i) t[i] = taps[i];
  const long long total = (long long)groups * rows * Kpad;
  int blocks = ceil_div(total, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  pack_weights_kernel<<<blocks, 256, 0, st>>>(W, out, mode, groups, w_group_stride, Cin, Cout, Cs, ntaps, rows, Kpad,
                                              t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

int launch_f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st) {
  if (n <= 0) return GEECO_OK;
  int blocks = ceil_div(n, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  f32_to_bf16_kernel<<<blocks, 256, 0, st>>>(in, out, n);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_bf16_to_f32(const __nv_bfloat16* in, float* out, long long n, cudaStream_t st) {
  if (n <= 0) return GEECO_OK;
  int blocks = ceil_div(n, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  bf16_to_f32_kernel<<<blocks, 256, 0, st>>>(in, out, n);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
