// conv2 data gradient -> conv1 weight gradient of the GEECO encoders as ONE kernel: dL/d(pre-activation of conv1), the
// largest tensor of the backward pass (805 MB per step at batch 64), goes from the data gradient's accumulators through
// the ReLU mask straight into the shared-memory operand of conv1's weight-gradient GEMM and never exists in HBM.  conv1
// is the first layer: nothing else reads that tensor.
//
// Reference ops: what tf.gradients derives (estimator.py:243-244) for the first two tf.layers.conv2d of conv_encoder
// (src/models/e2evmc/graph.py:76-115): 3x3 / stride 1 / SAME / ReLU, 3(4) -> 32 channels at 256 x 256, then
// 3x3 / stride 2 / SAME / ReLU, 32 -> 48 channels.
//
// Work unit = one row q of G2 = dL/d(pre-activation of conv2) of one image (128 pixels x 48 channels).  It yields the rows
// 2q and 2q+1 of G1 (256 pixels x 32 channels each) as four input-pixel parity classes (py, px): G1[2q+py][2ox+px] for
// ox = 0..127, from the G2 rows q-1 and q (TF SAME of an even-sized stride-2 layer pads only after).  A CTA walks a
// contiguous range of units, so consecutive units share a G2 row and every row is loaded once (ring of 3 rows).
//
//   data gradient   D[ox][(class, ci)] = sum over the 4 shifted windows (dy, dx) in {0,-1}^2 of G2 window x W2 taps.  The
//                   classes that use a window are adjacent accumulator columns (order c2 c0 c1 c3), so a window is ONE
//                   MMA of N = 128 / 64 / 64 / 32 against the concatenated weight k-blocks: 12 MMAs per unit instead of
//                   27 and each window is read from shared memory once, not once per class.
//   epilogue        thread = ox; set py (4 warps) owns the classes (py, 0) and (py, 1): TMEM -> ReLU mask of y1 (1 bit
//                   per value, written by the forward) -> bf16 -> the G1 tile of G1 row 2q+py: row ox = the pixel pair
//                   (2ox, 2ox+1) x 32 channels = 128 bytes, SWIZZLE_128B, i.e. the MN-major A operand of the
//                   weight-gradient GEMM; the tiles of py = 0 and py = 1 are its two 64-row M atoms.
//   producers       im2col of x0 for BOTH G1 rows of the unit: row ox = the 4 x 4 pixel window (x0 rows 2q-1 .. 2q+2,
//                   columns 2ox-1 .. 2ox+2) x 4 channels = 64 values = 128 bytes, 3 cp.async per window row (16 + 8 + 8
//                   bytes): every tap of every class is one of its columns.
//   weight gradient D2[(py, px, co)][(wy, cx, ch)] += [G1 tile py=0 | G1 tile py=1]^T x window tile: M = 128, N = 64,
//                   K = 128 pairs = 8 MMAs per unit; ky = wy - py, kx = cx - px.  The bias gradient rides in the same MMAs:
//                   N = 80, whose second 64-column atom is a constant tile with 1.0 in its first column (a separate
//                   N = 16 MMA re-read the whole A operand: a fifth of all operand reads).  One accumulator for the whole
//                   CTA, written once at the end as a 33 KB partial; a small kernel sums the partials of a group in a
//                   fixed order (deterministic).
//
//   warps 0-3 producers, 4-11 epilogue (two sets), 12 data-gradient MMA issuer (+ TMEM), 13 weight-gradient MMA issuer,
//   14 TMA (weights once, one G2 row per unit).
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

using namespace tc;

namespace {

constexpr int HW = 256;                        // height = width of x0 / y1 / G1
constexpr int C1 = 32, C2 = 48;
constexpr int ROW_PIX = 136;                   // G2 row box: pixels -1 .. 134 (zero-filled outside 0 .. 127)
constexpr int ROW_BYTES = ROW_PIX * 128;       // 64-channel (48 real) rows of 128 bytes
constexpr int RG = 3;                          // G2 row ring
constexpr int W_SLOT = C1 * 128;               // one (class, tap) k-block of the packed data-gradient weights: 32 x 64
constexpr int W_BYTES = 9 * W_SLOT;
constexpr int TILE = 128 * 128;                // G1 tile / window tile: 128 pixel pairs x 128 bytes
constexpr int NG1 = 2;                         // G1 buffers: [py = 0 | py = 1] tile pairs (the two M atoms of one A operand)
constexpr int NX = 3;                          // window tiles in flight
constexpr int NB = 2;                          // data-gradient accumulators (128 columns each)
constexpr int THREADS = 15 * 32;
constexpr int ACC_COL = 256;                   // weight-gradient accumulator: TMEM columns [256, 320), bias gradient [320, 336)
constexpr int ONES_BYTES = 128 * 128;          // constant second N atom of the weight-gradient B operand: column 0 = 1.0
constexpr int PART_FLOATS = 128 * 64 + 128;    // per-CTA partial: D2 and the bias column
constexpr int SMEM_BYTES = 1024 + W_BYTES + RG * ROW_BYTES + NG1 * 2 * TILE + NX * TILE + ONES_BYTES + 1024;

__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

struct B21Args {
  const __nv_bfloat16* x0;        // [G*M][256][256][4]
  const uint2* bits1;             // ReLU mask of y1: one 32-bit word per pixel, read as pixel pairs
  float* partial;                 // [CTAs][PART_FLOATS]
  int M;                          // images per encoder group
  int cpg;                        // CTAs per encoder group
  short slot_cls[9], slot_tap[9]; // shared-memory weight slot -> (class map, k-block) it is loaded from
};
struct B21Maps { CUtensorMap m[4]; };

// G2 rows a unit adds to the ring: [q - 1, q] at the start of a range / image, else [q, q]
__device__ __forceinline__ int unit_new_rows(int u, int u_lo) { return (u == u_lo || (u & 127) == 0) ? 2 : 1; }

__global__ void __launch_bounds__(THREADS, 1)
conv21_bwd_fused_kernel(const B21Args a, const __grid_constant__ B21Maps wmaps, const __grid_constant__ CUtensorMap g2map) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;
  uint8_t* ring = wsm + W_BYTES;
  uint8_t* g1t = ring + RG * ROW_BYTES;
  uint8_t* xt = g1t + NG1 * 2 * TILE;
  uint8_t* ones = xt + NX * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + ONES_BYTES);
  uint64_t* r_full = bars;                 // [RG]
  uint64_t* r_empty = r_full + RG;         // [RG]
  uint64_t* tm_full = r_empty + RG;        // [NB]
  uint64_t* tm_empty = tm_full + NB;       // [NB]
  uint64_t* g1_full = tm_empty + NB;       // [NG1]
  uint64_t* g1_empty = g1_full + NG1;      // [NG1]
  uint64_t* x_full = g1_empty + NG1;       // [NX]
  uint64_t* x_empty = x_full + NX;         // [NX]
  uint64_t* w_full = x_empty + NX;         // [1]
  uint64_t* acc_full = w_full + 1;         // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x / a.cpg, lb = blockIdx.x - group * a.cpg;
  const long long units = (long long)a.M * 128;
  const int u_lo = (int)(units * lb / a.cpg), u_hi = (int)(units * (lb + 1) / a.cpg);
  const long long gimg0 = (long long)group * a.M;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RG; ++s) { mbar_init(&r_full[s], 1); mbar_init(&r_empty[s], 1); }
    for (int b = 0; b < NB; ++b) { mbar_init(&tm_full[b], 1); mbar_init(&tm_empty[b], 256); }
    for (int s = 0; s < NG1; ++s) { mbar_init(&g1_full[s], 256); mbar_init(&g1_empty[s], 1); }
    for (int s = 0; s < NX; ++s) { mbar_init(&x_full[s], 128); mbar_init(&x_empty[s], 1); }
    mbar_init(w_full, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  // constant operand of the bias gradient: column 0 of every K row is 1.0 (bf16 0x3f80), the rest zero
  for (int i = threadIdx.x * 16; i < ONES_BYTES; i += THREADS * 16) *reinterpret_cast<uint4*>(ones + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < 128) *reinterpret_cast<uint16_t*>(ones + threadIdx.x * 128 + ((threadIdx.x & 7) << 4)) = 0x3f80;
  fence_proxy_async();
  if (warp == 12) tmem_alloc(tmem_ptr_s, 512);
  if (warp == 14 && lane == 0) {
    for (int c = 0; c < 4; ++c) tma_prefetch_desc(&wmaps.m[c]);
    tma_prefetch_desc(&g2map);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp < 4) {
    // ===================== producers: 4 x 4 pixel windows of x0 around the pixel pairs of G1 rows 2q, 2q+1 =====================
    const int ox = threadIdx.x;
    const uint32_t sw = (uint32_t)ox & 7u;
    const uint32_t row_u32 = smem_u32(xt) + (uint32_t)ox * 128u;
    const bool okl = ox >= 1, okr = ox <= 126;
    uint32_t s = 0, sphase = 0;
    for (int u = u_lo; u < u_hi; ++u) {
      const long long gi = gimg0 + (u >> 7);
      const int q = u & 127;
      // pixel (2q - 1, 2ox) of the image
      const char* sp = reinterpret_cast<const char*>(a.x0) + (((gi * HW + (2 * q - 1)) * HW) + 2 * ox) * 8;
      mbar_wait(&x_empty[s], sphase ^ 1u);
      const uint32_t d0 = row_u32 + s * (uint32_t)TILE;
#pragma unroll
      for (int wy = 0; wy < 4; ++wy) {
        const bool ok = (unsigned)(2 * q - 1 + wy) < (unsigned)HW;
        const uint32_t ca = d0 + (((uint32_t)(2 * wy) ^ sw) << 4), cb = d0 + (((uint32_t)(2 * wy + 1) ^ sw) << 4);
        const char* p = ok ? sp : reinterpret_cast<const char*>(a.x0) + 8;
        cp_async16(ca, p, ok ? 16u : 0u);                                   // pixels 2ox, 2ox+1
        cp_async8(cb, p - 8, (ok && okl) ? 8u : 0u);                         // pixel 2ox-1
        cp_async8(cb + 8, ok && okr ? p + 16 : p, (ok && okr) ? 8u : 0u);    // pixel 2ox+2
        sp += HW * 8;
      }
      cp_async_mbar_arrive_noinc(&x_full[s]);
      if (++s == NX) { s = 0; sphase ^= 1u; }
    }
  } else if (warp < 12) {
    // ===================== epilogue: set py, thread = ox =====================
    const int py = (warp - 4) >> 2, quad = warp & 3;
    const int ox = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    // accumulator columns [c2 | c0 | c1 | c3]: class (py, px) = c(2*py + px)
    const uint32_t col0 = py == 0 ? 32u : 0u, col1 = py == 0 ? 64u : 96u;
    const uint32_t sw = (uint32_t)ox & 7u;
    const uint32_t trow = smem_u32(g1t) + (uint32_t)py * (uint32_t)TILE + (uint32_t)ox * 128u;
    uint32_t buf = 0, bphase = 0, gb = 0, gphase = 0;
    // mask words of pixels 2ox, 2ox+1 of G1 row 2q+py, fetched two units ahead: the load's latency sat on the critical
    // path of every unit (ncu: 22 % of all stall samples on its first use)
    const uint2* bp = a.bits1 + (((gimg0 + (u_lo >> 7)) * HW + 2 * (u_lo & 127) + py) * HW) / 2 + ox;
    uint2 bits0 = __ldg(bp), bits1 = make_uint2(0u, 0u);
    if (u_lo + 1 < u_hi) bits1 = __ldg(bp + HW);
    for (int u = u_lo; u < u_hi; ++u) {
      const uint2 bits = bits0;
      bits0 = bits1;
      if (u + 2 < u_hi) bits1 = __ldg(bp + 2 * HW);     // consecutive units are two G1 rows apart, also across images
      bp += HW;
      mbar_wait(&tm_full[buf], bphase);
      tc_fence_after();
      uint32_t o[32];
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * 128u + (px ? col1 : col0), v);
        tmem_ld_wait();
        if (px == 1) {
          tc_fence_before();
          mbar_arrive(&tm_empty[buf]);
        }
        const uint32_t bw = px ? bits.y : bits.x;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t bh = h ? bw >> 16 : bw;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // bit i (channel 2i) -> byte 0's sign, bit 8+i (channel 2i+1) -> byte 1's sign; 0x9988 replicates them
            uint32_t keep;
            asm("prmt.b32 %0, %1, 0, 0x9988;" : "=r"(keep) : "r"(bh << (7 - i)));
            o[px * 16 + h * 8 + i] = pack_bf16x2(__uint_as_float(v[h * 16 + 2 * i]), __uint_as_float(v[h * 16 + 2 * i + 1])) & keep;
          }
        }
      }
      if (++buf == NB) { buf = 0; bphase ^= 1u; }
      // the buffer's tile pair has been consumed by the weight-gradient MMAs of NG1 units ago
      mbar_wait(&g1_empty[gb], gphase ^ 1u);
      const uint32_t tr = trow + gb * (uint32_t)(2 * TILE);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        st_shared_v4(tr + (((uint32_t)c ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
      fence_proxy_async();
      mbar_arrive(&g1_full[gb]);
      if (++gb == NG1) { gb = 0; gphase ^= 1u; }
    }
    if (py == 0) {
      // the CTA's weight-gradient accumulator (M = 128: TMEM lane = row (py, px, co)) -> its partial
      mbar_wait(acc_full, 0);
      tc_fence_after();
      float* P = a.partial + (long long)blockIdx.x * PART_FLOATS;
      float* Pr = P + (quad * 32 + lane) * 64;
#pragma unroll
      for (int c0 = 0; c0 < 80; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_addr + (uint32_t)ACC_COL + (uint32_t)c0, v);
        tmem_ld_wait();
        if (c0 < 64) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(Pr + c0 + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                      __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        } else {
          P[128 * 64 + quad * 32 + lane] = __uint_as_float(v[0]);
        }
      }
    }
  } else if (warp == 12) {
    // ===================== data-gradient MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    const uint32_t id128 = make_idesc_bf16(128, 128, 0, 0), id64 = make_idesc_bf16(128, 64, 0, 0), id32 = make_idesc_bf16(128, 32, 0, 0);
    const uint64_t dtempl = make_desc_sw128(0, 16, 1024);
    const uint32_t w16 = smem_u32(wsm) >> 4, ring16 = smem_u32(ring) >> 4;
    constexpr uint32_t SL16 = W_SLOT >> 4, ROW16 = ROW_BYTES >> 4;
    mbar_wait(w_full, 0);
    tc_fence_after();
    uint32_t n = 0;                      // G2 rows loaded so far by this CTA
    uint32_t cur = 0;                    // load index of row q (of the previous unit before the update below)
    uint32_t buf = 0, bphase = 0;
    for (int u = u_lo; u < u_hi; ++u) {
      uint32_t prev;
      if (unit_new_rows(u, u_lo) == 2) { prev = n; cur = n + 1; n += 2; }
      else { prev = cur; cur = n; n += 1; }
      const uint32_t sp = prev % RG, sc = cur % RG;
      mbar_wait(&r_full[sp], (prev / RG) & 1u);
      mbar_wait(&r_full[sc], (cur / RG) & 1u);
      mbar_wait(&tm_empty[buf], bphase ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + buf * 128u;
      const uint32_t ac = ring16 + sc * ROW16, ap = ring16 + sp * ROW16;
      if (elect_one()) {
        // window (dy, dx): pixel ox + dx sits at box pixel ox + dx + 1 -> byte offset (dx + 1) * 128
#pragma unroll
        for (int j = 0; j < 3; ++j)      // (0, 0): all four classes
          tc_mma(d, dtempl | (uint64_t)(ac + 8 + 2 * j), dtempl | (uint64_t)(w16 + 2 * j), id128, j != 0 ? 1u : 0u);
#pragma unroll
        for (int j = 0; j < 3; ++j)      // (0, -1): classes c2, c0
          tc_mma(d, dtempl | (uint64_t)(ac + 2 * j), dtempl | (uint64_t)(w16 + 4 * SL16 + 2 * j), id64, 1u);
#pragma unroll
        for (int j = 0; j < 3; ++j)      // (-1, 0): classes c0, c1
          tc_mma(d + 32, dtempl | (uint64_t)(ap + 8 + 2 * j), dtempl | (uint64_t)(w16 + 6 * SL16 + 2 * j), id64, 1u);
#pragma unroll
        for (int j = 0; j < 3; ++j)      // (-1, -1): class c0
          tc_mma(d + 32, dtempl | (uint64_t)(ap + 2 * j), dtempl | (uint64_t)(w16 + 8 * SL16 + 2 * j), id32, 1u);
        tc_commit(&tm_full[buf]);
        // row q-1 is not needed again; row q is the next unit's q-1 unless the image or the range ends here
        tc_commit(&r_empty[sp]);
        if ((u & 127) == 127 || u + 1 == u_hi) tc_commit(&r_empty[sc]);
      }
      __syncwarp();
      if (++buf == NB) { buf = 0; bphase ^= 1u; }
    }
  } else if (warp == 13) {
    // ===================== weight-gradient MMA issuer: one item per unit =====================
    const uint32_t idw = make_idesc_bf16(128, 80, 1, 1);
    const uint64_t adesc = make_desc_sw128(0, TILE, 1024);        // two M atoms (py = 0, 1), TILE bytes apart
    const uint32_t g16 = smem_u32(g1t) >> 4, x16 = smem_u32(xt) >> 4, ones16 = smem_u32(ones) >> 4;
    const uint32_t dacc = tmem_base + (uint32_t)ACC_COL;
    uint32_t s = 0, sphase = 0, gb = 0, gphase = 0;
    for (int u = u_lo; u < u_hi; ++u) {
      mbar_wait(&x_full[s], sphase);
      mbar_wait(&g1_full[gb], gphase);
      tc_fence_after();
      const uint32_t a16 = g16 + gb * (uint32_t)(2 * TILE >> 4), b16 = x16 + s * (uint32_t)(TILE >> 4);
      const uint32_t acc = u != u_lo ? 1u : 0u;
      // B = [window tile | ones tile]: the second N atom sits (ones - tile) bytes after the first
      const uint64_t bdesc = make_desc_sw128(0, (ones16 - b16) << 4, 1024);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j)        // 8 x 16 pixel pairs
          tc_mma(dacc, adesc | (uint64_t)(a16 + j * 128), bdesc | (uint64_t)(b16 + j * 128), idw, j != 0 ? 1u : acc);
        tc_commit(&g1_empty[gb]);
        tc_commit(&x_empty[s]);
      }
      __syncwarp();
      if (++s == NX) { s = 0; sphase ^= 1u; }
      if (++gb == NG1) { gb = 0; gphase ^= 1u; }
    }
    if (elect_one()) tc_commit(acc_full);
    __syncwarp();
  } else {
    // ===================== TMA warp: the weight k-blocks once, then the G2 rows =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (uint32_t)W_BYTES);
      for (int sl = 0; sl < 9; ++sl)
        tma_load_2d(smem_u32(wsm + sl * W_SLOT), &wmaps.m[a.slot_cls[sl]], w_full, a.slot_tap[sl] * 64, group * C1);
      uint32_t n = 0;
      for (int u = u_lo; u < u_hi; ++u) {
        const int q = u & 127;
        const int img = (int)(gimg0 + (u >> 7));
        for (int row = q + 1 - unit_new_rows(u, u_lo); row <= q; ++row, ++n) {
          const uint32_t s = n % RG;
          mbar_wait(&r_empty[s], ((n / RG) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&r_full[s], (uint32_t)ROW_BYTES);
          tma_load_5d(smem_u32(ring + s * ROW_BYTES), &g2map, &r_full[s], 0, -1, 0, row, img);   // row -1: zero-filled
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 512);
}

// dW1[g][(ky*3 + kx)*Cw + ch][co] = sum over the group's CTAs and the four classes (py, px) of
// P[cta][(py*2 + px)*32 + co][(ky + py)*16 + slot(px + kx)*4 + ch]; the bias gradient is the extra column.  Window slot of
// pair-relative column cx = px + kx (0: pixel 2ox-1 .. 3: pixel 2ox+2): the two pixels of the pair first, then the left and
// the right neighbour.  One block per (group, k): thread (co, j) sums the CTAs j, j + 8, ..; fixed order throughout.
__global__ void __launch_bounds__(256) conv21_bwd_reduce_kernel(const float* __restrict__ P, float* __restrict__ dW,
                                                                float* __restrict__ dbias, int cpg, int Cw,
                                                                long long dw_group_stride, long long dbias_group_stride) {
  pdl_enter();
  __shared__ float red[8][C1];
  const int nk = 9 * Cw + 1;
  const int g = blockIdx.x / nk, k = blockIdx.x - g * nk;
  const int co = threadIdx.x & 31, j = threadIdx.x >> 5;
  int off[4];
  if (k < 9 * Cw) {
    const int tap = k / Cw, ch = k - tap * Cw, ky = tap / 3, kx = tap - ky * 3;
    for (int c = 0; c < 4; ++c) {
      const int py = c >> 1, px = c & 1, cx = px + kx;
      const int slot = cx == 0 ? 2 : (cx == 1 ? 0 : (cx == 2 ? 1 : 3));
      off[c] = (c * C1 + co) * 64 + (ky + py) * 16 + slot * 4 + ch;
    }
  } else {
    for (int c = 0; c < 4; ++c) off[c] = 128 * 64 + c * C1 + co;
  }
  float s = 0.f;
  for (int cta = g * cpg + j; cta < (g + 1) * cpg; cta += 8) {
    const float* p = P + (long long)cta * PART_FLOATS;
    s += (p[off[0]] + p[off[1]]) + (p[off[2]] + p[off[3]]);
  }
  red[j][co] = s;
  __syncthreads();
  if (j == 0) {
    float t = red[0][co];
    for (int i = 1; i < 8; ++i) t += red[i][co];
    if (k < 9 * Cw) dW[(long long)g * dw_group_stride + (long long)k * C1 + co] = t;
    else if (dbias) dbias[(long long)g * dbias_group_stride + co] = t;
  }
}

}  // namespace

// Whether the fused kernel covers the data gradient of layer 2 (classes dg[0..3] in (py, px) order) followed by the
// weight gradient of layer 1; on success fills the weight-slot table.
static bool bwd21_slots(const TcGeom* dg, int ncls, short* slot_cls, short* slot_tap) {
  if (ncls != 4) return false;
  // accumulator column order c2 c0 c1 c3; windows (0,0) | (0,-1) | (-1,0) | (-1,-1)
  const int want[9][3] = {{2, 0, 0}, {0, 0, 0}, {1, 0, 0}, {3, 0, 0}, {2, 0, -1}, {0, 0, -1}, {0, -1, 0}, {1, -1, 0}, {0, -1, -1}};
  int used = 0;
  for (int s = 0; s < 9; ++s) {
    const TcGeom& g = dg[want[s][0]];
    int found = -1;
    for (int t = 0; t < g.ntaps; ++t)
      if (g.dy[t] == want[s][1] && g.dx[t] == want[s][2]) found = t;
    if (found < 0) return false;
    slot_cls[s] = (short)want[s][0]; slot_tap[s] = (short)found;
    ++used;
  }
  int total = 0;
  for (int c = 0; c < 4; ++c) {
    total += dg[c].ntaps;
    if (dg[c].dy0 != (c >> 1) || dg[c].dx0 != (c & 1) || dg[c].Kt != 64 || !dg[c].a_tma) return false;
  }
  return used == 9 && total == 9;
}

bool tc_bwd21_supported(int H, int W, int Cin_pad, int Cout1, int Cout2, int stride1, int stride2, const TcGeom* dg, int ncls) {
  if (getenv("GEECO_NO_FUSE_BWD21")) return false;
  short sc[9], stp[9];
  return H == HW && W == HW && Cin_pad == 4 && Cout1 == C1 && Cout2 == C2 && stride1 == 1 && stride2 == 2 &&
         bwd21_slots(dg, ncls, sc, stp);
}

long long tc_bwd21_partial_floats() { return (long long)tc_num_sms() * PART_FLOATS; }

int launch_tc_bwd21(const __nv_bfloat16* G2, const TcGeom* dg, const CUtensorMap* const* wmaps, const unsigned short* bits1,
                    const __nv_bfloat16* x0, float* partial, long long partial_cap, float* dW1, float* dbias1, int Cw,
                    long long dw_group_stride, long long dbias_group_stride, int G, int M, cudaStream_t st) {
  if (G < 1 || M < 1) return GEECO_OK;
  B21Args a;
  memset(&a, 0, sizeof(a));
  if (!bwd21_slots(dg, 4, a.slot_cls, a.slot_tap)) { geeco_set_error("bwd21: unexpected data-gradient classes"); return GEECO_ERR_INVALID; }
  if (Cw < 1 || Cw > 4) { geeco_set_error("bwd21: %d input channels", Cw); return GEECO_ERR_INVALID; }
  if ((long long)G * M * HW * HW >= (1ll << 31)) { geeco_set_error("bwd21: too many pixels for 32-bit indexing"); return GEECO_ERR_INVALID; }
  int cpg = tc_num_sms() / G;
  if (cpg < 1) cpg = 1;
  if ((long long)cpg > (long long)M * 128) cpg = M * 128;
  if ((long long)cpg * G * PART_FLOATS > partial_cap) { geeco_set_error("bwd21: partial buffer too small"); return GEECO_ERR_WORKSPACE; }
  B21Maps maps;
  for (int c = 0; c < 4; ++c) maps.m[c] = *wmaps[c];
  CUtensorMap g2map;
  int rc = tc_make_row_tensor_map(&g2map, G2, dg[0], ROW_PIX);
  if (rc) return rc;
  a.x0 = x0; a.bits1 = reinterpret_cast<const uint2*>(bits1); a.partial = partial; a.M = M; a.cpg = cpg;
  CUDA_TRY(cudaFuncSetAttribute(conv21_bwd_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CUDA_TRY(cudaFuncSetAttribute(conv21_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  GEECO_LAUNCH((conv21_bwd_fused_kernel), cpg * G, THREADS, SMEM_BYTES, st, a, maps, g2map);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  GEECO_LAUNCH((conv21_bwd_reduce_kernel), G * (9 * Cw + 1), 256, 0, st, (const float*)partial, dW1, dbias1, cpg, Cw, dw_group_stride, dbias_group_stride);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
