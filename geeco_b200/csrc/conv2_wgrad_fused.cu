// conv2 weight gradient with conv1's activation RECOMPUTED on chip: y1, the full-resolution 32-channel map (805 MB per
// step at batch 64, the largest tensor of the network), was written by the forward for exactly one reader -- this
// gradient.  Recomputing a y1 row from x0 costs three K = 16 MMAs per 128 pixels, so the rows are rebuilt here the way the
// fused forward (conv12_fused.cu) builds them and consumed from the same shared-memory ring; with the fused backward
// (conv21_bwd_fused.cu) taking the ReLU mask from its 1-bit copy, y1 no longer exists in HBM at all in training.
//
// Reference ops: d(loss)/d(kernel), d(loss)/d(bias) of the second tf.layers.conv2d of conv_encoder
// (src/models/e2evmc/graph.py:76-115; gradients: estimator.py:243-244): 3x3 / stride 2 / SAME, 32 -> 48 channels, input
// 256 x 256 (= conv1's output: 3x3 / stride 1 / SAME / ReLU over the 3(4)-channel network input).
//
// Work unit = one row oy of G2 = dL/d(pre-activation of conv2) of one image (128 pixels x 48 channels); it meets the y1
// rows 2oy, 2oy+1, 2oy+2 (row 256 does not exist: TF SAME pads after).  A CTA walks a contiguous range of units, so
// every y1 row is computed once (plus one per range start).  y1 rows live in the ring as pixel pairs x 64 channels
// (SWIZZLE_128B): seen as an MN-major operand, ring row ox is the 64 values (p, ci) = (kx = 2j + p, ci) of pair ox + j.
//
//   dW2^T[(ky, j)][(p, ci)][co] += ring(ky)[pair ox + j]^T x G2row[ox]      M = 64, N = 48, K = 128 pixels
//   dbias2[co]                  += ones^T x G2row                           (a constant A tile whose row 0 is 1.0)
//
//   warps 0-3    producers: im2col of x0 for conv1 (cp.async, one 128-pixel half row per stage)
//   warps 4-11   conv1 epilogue: TMEM -> ReLU -> bf16 -> the y1 ring (bit-identical to the forward's y1)
//   warps 12-15  idle until the end, then write the CTA's accumulators as its partial (a small kernel sums the partials
//                of an encoder group in a fixed order: deterministic)
//   warp 16/19   conv1 MMA issuers (left / right half rows), warp 17 weight-gradient MMA issuer
//   warp 18      TMA: conv1's packed weights once, one G2 row tile per unit
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

using namespace tc;

namespace {

constexpr int HW = 256;                        // height = width of x0 / y1
constexpr int C1 = 32, C2 = 48;
constexpr int SLOT_PAIRS = 136;                // 128 pixel pairs of a y1 row + zero pairs (the j = 1 window reads pair 128)
constexpr int SLOT_BYTES = SLOT_PAIRS * 128;
constexpr int A1_BYTES = 128 * 128;            // conv1 im2col tile: 128 pixels x 64 K (bf16)
constexpr int B1_BYTES = C1 * 128;
constexpr int G2_BYTES = 128 * 128;            // G2 row tile: 128 pixels x 64 channels (48 real)
constexpr int S1 = 4;                          // im2col stages
constexpr int RING = 6;                        // y1 row slots
constexpr int NG2 = 3;                         // G2 row tiles in flight
constexpr int NB1 = 4;                         // conv1 accumulator buffers (32 columns each): TMEM columns [0, 128)
constexpr int ACC_COL = 128;                   // six weight-gradient accumulators (ky, j), 48 columns each, then the bias one
constexpr int ONES_BYTES = 16 * 128;           // constant tiles: 16 K rows whose element 0 is 1.0, then two tiles of zeros
constexpr int THREADS = 20 * 32;
constexpr int PART_FLOATS = 6 * 64 * C2 + C2;  // per-CTA partial: [ky][j][(p, ci)][co] and the bias gradient
constexpr int SMEM_BYTES = 1024 + S1 * A1_BYTES + B1_BYTES + RING * SLOT_BYTES + NG2 * G2_BYTES + 3 * ONES_BYTES + 1024;

__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

struct WG2Args {
  const __nv_bfloat16* x0;        // [G*M][256][256][4]
  float* partial;                 // [CTAs][PART_FLOATS]
  int M;                          // images per encoder group
  int cpg;                        // CTAs per encoder group
};

// rows of y1 a unit adds to the ring: [r_begin, r_end]
__device__ __forceinline__ void unit_rows(int u, int u_lo, int& r_begin, int& r_end) {
  const int oy = u & 127;
  r_begin = (u == u_lo || oy == 0) ? 2 * oy : 2 * oy + 1;
  r_end = oy < 127 ? 2 * oy + 2 : 2 * oy + 1;
}


// Weight-gradient MMA issuer shared by the two kernels below (one warp runs it, one elected lane issues): unit u meets
// the ring rows q0, q1, q2 and its G2 row tile.  STACK: the windows j = 0 and j = 1 of a ring row are the two 64-row M
// atoms of ONE M = 128 operand (atom stride = one pair = 128 bytes: the atoms overlap), three accumulators [128 x 48];
// otherwise six M = 64 accumulators (ky, j).  The bias gradient is always an M = 64 product with the ones tile.
template <bool STACK>
__device__ __forceinline__ void wg2_issue(int u_lo, int u_hi, uint32_t dacc, uint32_t ring_u32, uint32_t g2_u32, uint32_t ones_u32,
                                          uint32_t zeros_u32, uint64_t* y_full, uint64_t* y_empty, uint64_t* g_full,
                                          uint64_t* g_empty, uint64_t* acc_full, int ng2, int nring) {
  const uint32_t idw = make_idesc_bf16(64, C2, 1, 1), idw128 = make_idesc_bf16(128, C2, 1, 1);
  const uint64_t dtempl = make_desc_sw128(0, 8192, 1024);      // MN-major: one 64-element atom, 8-row K groups 1024 bytes apart
  const uint64_t dstack = make_desc_sw128(0, 128, 1024);       // two M atoms one pixel pair apart
  const uint32_t ring16 = ring_u32 >> 4, g16 = g2_u32 >> 4, ones16 = ones_u32 >> 4, z16 = zeros_u32 >> 4;
  const uint32_t dbias = dacc + 6 * C2;
  const uint32_t uring = (uint32_t)nring;
  uint32_t q = 0, gs = 0, gphase = 0;
  int last_q2 = -1;
  // every accumulator starts as 0 x 0 (a range made of last image rows only never writes the ky = 2 accumulators)
  if (elect_one()) {
    for (int acc = 0; acc < 7; ++acc)
      tc_mma(dacc + (uint32_t)acc * C2, dtempl | (uint64_t)z16, dtempl | (uint64_t)z16, idw, 0u);
    if (STACK)                                                  // lanes 64..127 of the three stacked accumulators
      for (int acc = 0; acc < 3; ++acc)
        tc_mma(dacc + (uint32_t)acc * C2, make_desc_sw128(0, ONES_BYTES, 1024) | (uint64_t)z16, dtempl | (uint64_t)z16, idw128, 0u);
  }
  __syncwarp();
  for (int u = u_lo; u < u_hi; ++u) {
    int r_begin, r_end;
    unit_rows(u, u_lo, r_begin, r_end);
    const bool first = r_begin == 2 * (u & 127);
    const int nrows = r_end - r_begin + 1;
    int q0, q1, q2;
    if (first) { q0 = (int)q; q1 = (int)q + 1; q2 = nrows == 3 ? (int)q + 2 : -1; }
    else { q0 = last_q2; q1 = (int)q; q2 = nrows == 2 ? (int)q + 1 : -1; }
    last_q2 = q2;
    q += (uint32_t)nrows;
    const int nky = q2 >= 0 ? 3 : 2;
    mbar_wait(&g_full[gs], gphase);
    tc_fence_after();
    const uint32_t b16 = g16 + gs * (uint32_t)(G2_BYTES >> 4);
    if (elect_one()) {
#pragma unroll
      for (int j = 0; j < 8; ++j)        // bias gradient: ones^T x G2 row
        tc_mma(dbias, dtempl | (uint64_t)ones16, dtempl | (uint64_t)(b16 + j * 128), idw, 1u);
    }
    __syncwarp();
    for (int ky = 0; ky < nky; ++ky) {
      const uint32_t qq = (uint32_t)(ky == 0 ? q0 : (ky == 1 ? q1 : q2));
      mbar_wait(&y_full[qq % uring], (qq / uring) & 1u);
      tc_fence_after();
      const uint32_t sa16 = ring16 + (qq % uring) * (uint32_t)(SLOT_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {    // 8 x 16 output pixels; window j = 1 starts one pair (128 bytes) later
          if (STACK) {
            tc_mma(dacc + (uint32_t)ky * C2, dstack | (uint64_t)(sa16 + j * 128), dtempl | (uint64_t)(b16 + j * 128), idw128, 1u);
          } else {
            tc_mma(dacc + (uint32_t)(ky * 2) * C2, dtempl | (uint64_t)(sa16 + j * 128), dtempl | (uint64_t)(b16 + j * 128), idw, 1u);
            tc_mma(dacc + (uint32_t)(ky * 2 + 1) * C2, dtempl | (uint64_t)(sa16 + 8 + j * 128), dtempl | (uint64_t)(b16 + j * 128), idw, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      tc_commit(&g_empty[gs]);
      tc_commit(&y_empty[(uint32_t)q0 % uring]);
      tc_commit(&y_empty[(uint32_t)q1 % uring]);
      // the third row is the next unit's first unless the range ends here (a new image starts with its own row 0,
      // and then this unit had no third row)
      if (q2 >= 0 && u + 1 == u_hi) tc_commit(&y_empty[(uint32_t)q2 % uring]);
    }
    __syncwarp();
    if (++gs == (uint32_t)ng2) { gs = 0; gphase ^= 1u; }
  }
  if (elect_one()) tc_commit(acc_full);
  __syncwarp();
}

// The CTA's accumulators -> its partial [ky][j][(p, ci)][co] (+ the bias gradient); called by the four warps whose
// index mod 4 is `quad`.  M = 64 accumulators keep row m in lane (m % 16) + 32 * (m / 16); a stacked M = 128 accumulator
// keeps row j*64 + (p, ci) in lane j*64 + (p, ci).
template <bool STACK>
__device__ __forceinline__ void wg2_drain(float* P, uint32_t tmem_acc, int quad, int lane) {
  const uint32_t lane_addr = tmem_acc + ((uint32_t)(quad * 32) << 16);
  const int m64 = quad * 16 + (lane & 15);
  const int row128 = quad * 32 + lane;
#pragma unroll 1
  for (int acc = 0; acc < (STACK ? 3 : 6); ++acc) {
#pragma unroll
    for (int c0 = 0; c0 < C2; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(lane_addr + (uint32_t)(acc * C2 + c0), v);
      tmem_ld_wait();
      float* p = STACK ? P + ((long long)(acc * 2 + (row128 >> 6)) * 64 + (row128 & 63)) * C2 + c0 : P + ((long long)acc * 64 + m64) * C2 + c0;
      if (STACK || lane < 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(p + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                              __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
  }
#pragma unroll
  for (int c0 = 0; c0 < C2; c0 += 16) {                        // bias gradient: row 0 of the ones-tile product
    uint32_t v[16];
    tmem_ld16(lane_addr + (uint32_t)(6 * C2 + c0), v);
    tmem_ld_wait();
    if (quad == 0 && lane == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) P[6 * 64 * C2 + c0 + i] = __uint_as_float(v[i]);
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
conv2_wgrad_fused_kernel(const WG2Args a, const __grid_constant__ CUtensorMap w1map, const __grid_constant__ CUtensorMap g2map) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a1 = smem;
  uint8_t* b1 = a1 + S1 * A1_BYTES;
  uint8_t* ring = b1 + B1_BYTES;
  uint8_t* g2t = ring + RING * SLOT_BYTES;
  uint8_t* ones = g2t + NG2 * G2_BYTES;
  uint8_t* zeros = ones + ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(zeros + 2 * ONES_BYTES);
  uint64_t* a_full = bars;                 // [S1]
  uint64_t* a_empty = a_full + S1;         // [S1]
  uint64_t* t1_full = a_empty + S1;        // [NB1]
  uint64_t* t1_empty = t1_full + NB1;      // [NB1]
  uint64_t* y_full = t1_empty + NB1;       // [RING]
  uint64_t* y_empty = y_full + RING;       // [RING]
  uint64_t* g_full = y_empty + RING;       // [NG2]
  uint64_t* g_empty = g_full + NG2;        // [NG2]
  uint64_t* w_full = g_empty + NG2;        // [1]
  uint64_t* acc_full = w_full + 1;         // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x / a.cpg, lb = blockIdx.x - group * a.cpg;
  const long long units = (long long)a.M * 128;
  const int u_lo = (int)(units * lb / a.cpg), u_hi = (int)(units * (lb + 1) / a.cpg);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S1; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < NB1; ++b) { mbar_init(&t1_full[b], 1); mbar_init(&t1_empty[b], 128); }
    for (int r = 0; r < RING; ++r) { mbar_init(&y_full[r], 256); mbar_init(&y_empty[r], 1); }
    for (int s = 0; s < NG2; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 1); }
    mbar_init(w_full, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  // im2col columns >= 37 and the pad pairs of every ring slot are never written: zero everything once
  for (int i = threadIdx.x * 16; i < S1 * A1_BYTES; i += THREADS * 16) *reinterpret_cast<uint4*>(a1 + i) = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x * 16; i < RING * SLOT_BYTES; i += THREADS * 16) *reinterpret_cast<uint4*>(ring + i) = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x * 16; i < 3 * ONES_BYTES; i += THREADS * 16) *reinterpret_cast<uint4*>(ones + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // conv1's bias rides in the GEMM: im2col column 36 is a constant 1.0 (the packed weights hold the bias there)
  for (int i = threadIdx.x; i < S1 * 128; i += THREADS) {
    const int st_i = i >> 7, r = i & 127;
    *reinterpret_cast<uint16_t*>(a1 + st_i * A1_BYTES + r * 128 + ((4 ^ (r & 7)) << 4) + 8) = 0x3f80;
  }
  // constant A tile of the bias gradient: element m = 0 of every K row is 1.0
  if (threadIdx.x < 16) *reinterpret_cast<uint16_t*>(ones + threadIdx.x * 128 + ((threadIdx.x & 7) << 4)) = 0x3f80;
  fence_proxy_async();
  if (warp == 16) tmem_alloc(tmem_ptr_s, 512);
  if (warp == 18 && lane == 0) { tma_prefetch_desc(&w1map); tma_prefetch_desc(&g2map); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const long long gimg0 = (long long)group * a.M;               // first image of this encoder group

  if (warp < 4) {
    // ===================== conv1 producers: one half row (128 pixels) per stage =====================
    const int t = threadIdx.x;
    const uint32_t rsw4 = ((uint32_t)t & 7u) << 4;
    const uint32_t a_row0 = smem_u32(a1) + (uint32_t)t * 128u;
    uint32_t s = 0, sphase = 0;
    for (int u = u_lo; u < u_hi; ++u) {
      int r_begin, r_end;
      unit_rows(u, u_lo, r_begin, r_end);
      const long long gi = gimg0 + (u >> 7);
      for (int r = r_begin; r <= r_end; ++r) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int x = h * 128 + t;
          // top-left tap = source pixel (r - 1, x - 1); a tap outside the image copies zero bytes
          const char* sp = reinterpret_cast<const char*>(a.x0) + (((gi * HW + (r - 1)) * HW) + (x - 1)) * 8;
          const bool in0 = x >= 1, in2 = x + 1 < HW;
          mbar_wait(&a_empty[s], sphase ^ 1u);
          const uint32_t a_row = a_row0 + s * (uint32_t)A1_BYTES;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const bool okr = (unsigned)(r - 1 + ky) < (unsigned)HW;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t bb = (uint32_t)(ky * 24 + kx * 8);
              const uint32_t d = a_row + (((bb >> 4) << 4) ^ rsw4) + (bb & 15u);
              const bool ok = okr && (kx == 0 ? in0 : (kx == 1 ? true : in2));
              cp_async8(d, sp + kx * 8, ok ? 8u : 0u);
            }
            sp += HW * 8;
          }
          cp_async_mbar_arrive_noinc(&a_full[s]);
          if (++s == S1) { s = 0; sphase ^= 1u; }
        }
      }
    }
  } else if (warp < 12) {
    // ===================== conv1 epilogue: thread = pixel x of the y1 row =====================
    const int we = warp - 4, h = we >> 2, quad = we & 3;
    const int x = h * 128 + quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t pair = (uint32_t)x >> 1, hp = (uint32_t)x & 1u;
    const uint32_t slot_off = pair * 128u, psw = pair & 7u;
    const uint32_t ring_u32 = smem_u32(ring);
    uint32_t q = 0;                                             // y1 rows produced so far by this CTA
    for (int u = u_lo; u < u_hi; ++u) {
      int r_begin, r_end;
      unit_rows(u, u_lo, r_begin, r_end);
      for (int r = r_begin; r <= r_end; ++r, ++q) {
        const uint32_t slot = q % RING, use = q / RING;
        const uint32_t tile = 2 * q + (uint32_t)h, buf = tile % NB1, tuse = tile / NB1;
        mbar_wait(&t1_full[buf], tuse & 1u);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * 32u, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&t1_empty[buf]);
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        // the slot's previous row has been consumed by the weight-gradient MMAs
        mbar_wait(&y_empty[slot], (use & 1u) ^ 1u);
        const uint32_t srow = ring_u32 + slot * (uint32_t)SLOT_BYTES + slot_off;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(srow + (((hp * 4u + (uint32_t)j) ^ psw) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        fence_proxy_async();
        mbar_arrive(&y_full[slot]);
      }
    }
  } else if (warp < 16) {
    // ===================== drain: the CTA's accumulators -> its partial =====================
    while (!mbar_try(smem_u32(acc_full), 0)) __nanosleep(2000);
    tc_fence_after();
    wg2_drain<false>(a.partial + (long long)blockIdx.x * PART_FLOATS, tmem_base + (uint32_t)ACC_COL, warp & 3, lane);
  } else if (warp == 16 || warp == 19) {
    // ===================== conv1 MMA issuers (whole warp runs the loop, one elected lane issues) =====================
    const uint32_t idesc1 = make_idesc_bf16(128, C1, 0, 0);
    const uint64_t dtempl = make_desc_sw128(0, 16, 1024);
    const uint32_t a1_16 = smem_u32(a1) >> 4, b1_16 = smem_u32(b1) >> 4;
    mbar_wait(w_full, 0);
    tc_fence_after();
    // tile i uses stage i % S1 and buffer i % NB1 (both even counts): this warp's tiles are i = par, par + 2, ...
    const uint32_t par = warp == 16 ? 0u : 1u;
    uint32_t s = par, sphase = 0, buf = par, bphase = 0;
    for (int u = u_lo; u < u_hi; ++u) {
      int r_begin, r_end;
      unit_rows(u, u_lo, r_begin, r_end);
      const int ntiles = r_end - r_begin + 1;                    // of this warp: one per row
      for (int tl = 0; tl < ntiles; ++tl) {
        mbar_wait(&t1_empty[buf], bphase ^ 1u);
        mbar_wait(&a_full[s], sphase);
        tc_fence_after();
        const uint32_t a16 = a1_16 + s * (uint32_t)(A1_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 3; ++j)      // K = 36 taps x channels + the bias column: three K = 16 steps
            tc_mma(tmem_base + buf * 32u, dtempl | (uint64_t)(a16 + 2 * j), dtempl | (uint64_t)(b1_16 + 2 * j), idesc1, j != 0 ? 1u : 0u);
          tc_commit(&a_empty[s]);
          tc_commit(&t1_full[buf]);
        }
        __syncwarp();
        if ((s += 2) >= S1) { s -= S1; sphase ^= 1u; }
        if ((buf += 2) >= NB1) { buf -= NB1; bphase ^= 1u; }
      }
    }
  } else if (warp == 17) {
    wg2_issue<false>(u_lo, u_hi, tmem_base + (uint32_t)ACC_COL, smem_u32(ring), smem_u32(g2t), smem_u32(ones), smem_u32(zeros),
                     y_full, y_empty, g_full, g_empty, acc_full, NG2, RING);
  } else {
    // ===================== TMA warp: conv1's weights once, then one G2 row tile per unit =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (uint32_t)B1_BYTES);
      tma_load_2d(smem_u32(b1), &w1map, w_full, 0, group * C1);
      uint32_t s = 0, sphase = 0;
      for (int u = u_lo; u < u_hi; ++u) {
        mbar_wait(&g_empty[s], sphase ^ 1u);
        mbar_arrive_expect_tx(&g_full[s], (uint32_t)G2_BYTES);
        // G2 as [pixels][48 channels]: box 128 pixels x 64 channels, channels 48..63 zero-filled
        tma_load_2d(smem_u32(g2t + s * G2_BYTES), &g2map, &g_full[s], 0, (int)(((gimg0 + (u >> 7)) * 128 + (u & 127)) * 128));
        if (++s == NG2) { s = 0; sphase ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------------------------------
// The same weight gradient with the y1 rows LOADED (one TMA box per row straight into the ring: a y1 row in NHWC is
// 128 pixel pairs x 128 bytes, the ring's own layout) instead of recomputed.  Measured (B200, batch 64): recomputing costs
// ~2000 shared-memory wavefronts per unit (im2col copies, conv1 operand reads, ring stores) = 301 us, while the stored
// y1 costs the forward 27 us to write and this kernel reads it at HBM speed.  This is the default; the recomputing kernel
// above is kept as GEECO_WG2_RECOMPUTE=1 (it frees the 805 MB of y1 and 1.6 GB of DRAM traffic per step).
//   warp 0 TMA (y1 rows, G2 row tiles), warp 1 MMA issuer (+ TMEM), warps 4-7 drain
// ------------------------------------------------------------------------------------------------------------------------
constexpr int R_RING = 8, R_NG2 = 4, R_THREADS = 8 * 32;
constexpr int R_SMEM_BYTES = 1024 + R_RING * SLOT_BYTES + R_NG2 * G2_BYTES + 3 * ONES_BYTES + 1024;

template <bool STACK>
__global__ void __launch_bounds__(R_THREADS, 1)
conv2_wgrad_rows_kernel(const WG2Args a, const __grid_constant__ CUtensorMap y1map, const __grid_constant__ CUtensorMap g2map) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* g2t = ring + R_RING * SLOT_BYTES;
  uint8_t* ones = g2t + R_NG2 * G2_BYTES;
  uint8_t* zeros = ones + ONES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(zeros + 2 * ONES_BYTES);
  uint64_t* y_full = bars;                 // [R_RING]
  uint64_t* y_empty = y_full + R_RING;     // [R_RING]
  uint64_t* g_full = y_empty + R_RING;     // [R_NG2]
  uint64_t* g_empty = g_full + R_NG2;      // [R_NG2]
  uint64_t* acc_full = g_empty + R_NG2;    // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x / a.cpg, lb = blockIdx.x - group * a.cpg;
  const long long units = (long long)a.M * 128;
  const int u_lo = (int)(units * lb / a.cpg), u_hi = (int)(units * (lb + 1) / a.cpg);
  const long long gimg0 = (long long)group * a.M;

  if (threadIdx.x == 0) {
    for (int r = 0; r < R_RING; ++r) { mbar_init(&y_full[r], 1); mbar_init(&y_empty[r], 1); }
    for (int s = 0; s < R_NG2; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  // the pad pairs of every ring slot are never written (the j = 1 window of the last pixels reads pair 128): zero once
  for (int i = threadIdx.x * 16; i < R_RING * SLOT_BYTES; i += R_THREADS * 16) *reinterpret_cast<uint4*>(ring + i) = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x * 16; i < 3 * ONES_BYTES; i += R_THREADS * 16) *reinterpret_cast<uint4*>(ones + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < 16) *reinterpret_cast<uint16_t*>(ones + threadIdx.x * 128 + ((threadIdx.x & 7) << 4)) = 0x3f80;
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_ptr_s, 512);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&y1map); tma_prefetch_desc(&g2map); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t q = 0, s = 0, sphase = 0;
      for (int u = u_lo; u < u_hi; ++u) {
        int r_begin, r_end;
        unit_rows(u, u_lo, r_begin, r_end);
        const long long gi = gimg0 + (u >> 7);
        mbar_wait(&g_empty[s], sphase ^ 1u);
        mbar_arrive_expect_tx(&g_full[s], (uint32_t)G2_BYTES);
        tma_load_2d(smem_u32(g2t + s * G2_BYTES), &g2map, &g_full[s], 0, (int)((gi * 128 + (u & 127)) * 128));
        if (++s == R_NG2) { s = 0; sphase ^= 1u; }
        for (int r = r_begin; r <= r_end; ++r, ++q) {
          const uint32_t slot = q % R_RING;
          mbar_wait(&y_empty[slot], ((q / R_RING) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&y_full[slot], 128u * 128u);
          tma_load_2d(smem_u32(ring + slot * SLOT_BYTES), &y1map, &y_full[slot], 0, (int)((gi * HW + r) * 128));
        }
      }
    }
  } else if (warp == 1) {
    wg2_issue<STACK>(u_lo, u_hi, tmem_base, smem_u32(ring), smem_u32(g2t), smem_u32(ones), smem_u32(zeros), y_full, y_empty, g_full,
                     g_empty, acc_full, R_NG2, R_RING);
  } else if (warp >= 4) {
    while (!mbar_try(smem_u32(acc_full), 0)) __nanosleep(2000);
    tc_fence_after();
    wg2_drain<STACK>(a.partial + (long long)blockIdx.x * PART_FLOATS, tmem_base, warp & 3, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dW2[g][(ky*3 + kx)*32 + ci][co] = sum over the group's CTAs of P[cta][ky][j][p*32 + ci][co] with kx = 2j + p (the
// (j, p) = (1, 1) half is kx = 3: no such tap); dbias2 = the extra row.  One block per (group, k): thread (co, jj) sums
// the CTAs jj, jj + 4, ..; fixed order throughout.
__global__ void __launch_bounds__(256) conv2_wgrad_reduce_kernel(const float* __restrict__ P, float* __restrict__ dW,
                                                                 float* __restrict__ dbias, int cpg, long long dw_group_stride,
                                                                 long long dbias_group_stride) {
  pdl_enter();
  __shared__ float red[4][64];
  const int nk = 9 * C1 + 1;
  const int g = blockIdx.x / nk, k = blockIdx.x - g * nk;
  const int co = threadIdx.x & 63, jj = threadIdx.x >> 6;
  int off = 6 * 64 * C2;
  if (k < 9 * C1) {
    const int tap = k / C1, ci = k - tap * C1, ky = tap / 3, kx = tap - ky * 3;
    off = ((ky * 2 + (kx >> 1)) * 64 + (kx & 1) * C1 + ci) * C2;
  }
  float s = 0.f;
  if (co < C2)
    for (int cta = g * cpg + jj; cta < (g + 1) * cpg; cta += 4) s += P[(long long)cta * PART_FLOATS + off + co];
  red[jj][co] = s;
  __syncthreads();
  if (jj == 0 && co < C2) {
    const float t = ((red[0][co] + red[1][co]) + red[2][co]) + red[3][co];
    if (k < 9 * C1) dW[(long long)g * dw_group_stride + (long long)k * C2 + co] = t;
    else if (dbias) dbias[(long long)g * dbias_group_stride + co] = t;
  }
}

}  // namespace

bool tc_wgrad2_supported(int H, int W, int Cin_pad, int Cout1, int Cout2, int stride1, int stride2, const TcGeom& g1) {
  if (getenv("GEECO_NO_FUSE_WG2")) return false;
  return H == HW && W == HW && Cin_pad == 4 && Cout1 == C1 && Cout2 == C2 && stride1 == 1 && stride2 == 2 && g1.bias_in_k &&
         g1.Kpad == 64 && g1.Ktot == 36;
}

long long tc_wgrad2_partial_floats() { return (long long)tc_num_sms() * PART_FLOATS; }

int launch_tc_wgrad2(const __nv_bfloat16* x0, const CUtensorMap* w1map, const __nv_bfloat16* y1, const __nv_bfloat16* G2,
                     float* partial, long long partial_cap, float* dW2, float* dbias2, long long dw_group_stride,
                     long long dbias_group_stride, int G, int M, cudaStream_t st) {
  if (G < 1 || M < 1) return GEECO_OK;
  const long long pixels = (long long)G * M * 128 * 128;
  if (pixels >= (1ll << 31)) { geeco_set_error("wgrad2: %lld pixels exceed the 2^31 the load coordinates hold", pixels); return GEECO_ERR_INVALID; }
  int cpg = tc_num_sms() / G;
  if (cpg < 1) cpg = 1;
  if ((long long)cpg > (long long)M * 128) cpg = M * 128;
  if ((long long)cpg * G * PART_FLOATS > partial_cap) { geeco_set_error("wgrad2: partial buffer too small"); return GEECO_ERR_WORKSPACE; }
  CUtensorMap g2map;
  int rc = make_tensor_map_2d_sw128(&g2map, const_cast<__nv_bfloat16*>(G2), C2, pixels, C2 * 2, 64, 128);
  if (rc) return rc;
  WG2Args a;
  a.x0 = x0; a.partial = partial; a.M = M; a.cpg = cpg;
  if (y1) {
    // y1 rows by TMA: [pixel pairs][64] bf16, one box = the 128 pairs of a row
    CUtensorMap y1map;
    rc = make_tensor_map_2d_sw128(&y1map, const_cast<__nv_bfloat16*>(y1), 64, (long long)G * M * HW * 128, 128, 64, 128);
    if (rc) return rc;
    const bool stack = getenv("GEECO_WG2_NO_STACK") == nullptr;
    auto kern = stack ? conv2_wgrad_rows_kernel<true> : conv2_wgrad_rows_kernel<false>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM_BYTES));
    GEECO_LAUNCH((kern), cpg * G, R_THREADS, R_SMEM_BYTES, st, a, y1map, g2map);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(conv2_wgrad_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_TRY(cudaFuncSetAttribute(conv2_wgrad_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GEECO_LAUNCH((conv2_wgrad_fused_kernel), cpg * G, THREADS, SMEM_BYTES, st, a, *w1map, g2map);
  }
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  GEECO_LAUNCH((conv2_wgrad_reduce_kernel), G * (9 * C1 + 1), 256, 0, st, (const float*)partial, dW2, dbias2, cpg, dw_group_stride,
               dbias_group_stride);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
