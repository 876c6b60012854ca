// conv1 -> conv2 of the GEECO encoders as ONE kernel: the full-resolution 32-channel map y1 (805 MB per step at batch
// 64, the largest tensor of the network) goes from conv1's accumulators straight into the shared-memory operand of
// conv2 and never makes the HBM round trip between the two layers.
//
// Reference ops: the first two tf.layers.conv2d of conv_encoder (src/models/e2evmc/graph.py:76-115): 3x3 / stride 1 /
// SAME / ReLU, 3(4) -> 32 channels at 256 x 256, then 3x3 / stride 2 / SAME / ReLU, 32 -> 48 channels.
//
// Work unit = one y2 row (128 pixels x 48 channels) of one image; it needs the y1 rows 2*oy, 2*oy+1, 2*oy+2 (TF SAME
// for stride 2 pads after: row 256 does not exist).  A CTA walks a contiguous range of units, so consecutive units
// share a y1 row and every y1 row is computed exactly once (plus one per range start).
//
//   warps 0-3    producers: implicit im2col of x0 for conv1 (cp.async, one 128-pixel half row per stage; the row-window
//                gather of tc_nn_kernel<4,4>)
//   warps 4-11   conv1 epilogue, thread = pixel of the y1 row: TMEM -> ReLU -> bf16 -> the y1 ring in shared memory in
//                conv2's own operand layout (pixel pairs x 64 channels, SWIZZLE_128B), + the 1-bit ReLU mask (training)
//   warps 12-15  conv2 epilogue: TMEM -> bias + ReLU -> bf16 -> y2 (+ mask bits)
//   warp 16/19   conv1 MMA issuers (left / right half rows), warp 17 conv2 MMA issuer: independent instruction streams,
//                the tensor pipe interleaves them
//   warp 18      TMA: the packed weights of both layers, once (a CTA never crosses an encoder group); training only:
//                one TMA tensor store per finished y1 row, ring slot -> y1 in HBM (the backward needs it)
//
// Inference never writes y1: per unit 4 KB of x0 in, 12 KB of y2 out instead of 32 KB + 32 KB + 12 KB.
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

using namespace tc;

namespace {

constexpr int HW = 256;                        // height = width of x0 / y1
constexpr int C1 = 32, C2 = 48;
constexpr int SLOT_PAIRS = 136;                // 128 pixel pairs of a y1 row + zero pairs (the kx = 2 window reads pair 128)
constexpr int SLOT_BYTES = SLOT_PAIRS * 128;   // 17408 = 17 swizzle atoms
constexpr int A1_BYTES = 128 * 128;            // conv1 im2col tile: 128 pixels x 64 K (bf16)
constexpr int B1_BYTES = 2 * C1 * 128;         // conv1's packed weights: 32 rows (per pixel) or 64 rows (per pixel pair)
constexpr int B2_SLOT = C2 * 128;              // one (ky, pair) k-block of conv2's packed weights
constexpr int B2_BYTES = 6 * B2_SLOT;
// im2col stages S1 and y1 row slots RING are template arguments of the kernel: training (the y1 store holds ring slots
// longer, the producers are the critical path) runs best with 6 stages / 4 slots (263 vs 275 us), inference with 4 / 6
// (3.42 vs 3.57 ms per 3072 images); both fill the 227 KB
constexpr int NB1 = 8, NB2 = 2;                // accumulator buffers: conv1 (32 columns each), conv2 (64-column stride)
constexpr int NB1P = 4;                        // pixel-pair conv1: 64 columns per buffer (the same TMEM columns [0, 256))
constexpr int THREADS = 20 * 32;                // 4 producer, 8 + 4 epilogue, 3 MMA warps, 1 TMA warp
constexpr int Y2_STAGE = 32 * C2 * 2;           // 3 KB: the 32 pixels x 48 channels one conv2-epilogue warp stores per unit
constexpr int smem_bytes(int S1, int RING) { return 1024 + S1 * A1_BYTES + B1_BYTES + B2_BYTES + RING * SLOT_BYTES + 1024 + 4 * Y2_STAGE; }

__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
// contiguous shared -> global bulk copy (bulk async group of the issuing thread)
__device__ __forceinline__ void bulk_store(void* dst_global, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

struct C12Args {
  const __nv_bfloat16* x0;        // [G*M][256][256][4]
  __nv_bfloat16* y1;              // [G*M][256][256][32] or nullptr (inference)
  unsigned int* bits1;            // one word per y1 pixel or nullptr
  __nv_bfloat16* y2;              // [G*M][128][128][48]
  unsigned short* bits2;          // three halfwords per y2 pixel or nullptr
  const float* bias2;             // conv2 bias of group 0
  long long bias2_group_stride;
  int M;                          // images per encoder group
  int cpg;                        // CTAs per encoder group
};

// rows of y1 a unit adds to the ring: [r_begin, r_end]
__device__ __forceinline__ void unit_rows(int u, int u_lo, int& r_begin, int& r_end) {
  const int oy = u & 127;
  r_begin = (u == u_lo || oy == 0) ? 2 * oy : 2 * oy + 1;
  r_end = oy < 127 ? 2 * oy + 2 : 2 * oy + 1;
}

// PAIR: conv1 runs on pixel pairs.  A GEMM row is the pair (2t, 2t+1) of a y1 row, its K the 3 x 4 pixel window around
// the pair (k = ky*16 + slot*4 + ch, slots = pixels 2t, 2t+1, 2t-1, 2t+2: 16 + 8 + 8 byte copies per window row, half the
// copies and shared-memory wavefronts of the per-pixel im2col), its N the 64 values (p, co) -- which is exactly the ring
// row conv2 consumes, so one thread writes one whole ring row.  One MMA tile = one y1 row (4 K = 16 steps of N = 64)
// instead of two half rows (3 steps of N = 32 each).  The packed weights are the mode-5 copy (pack_value).  Same
// products as the per-pixel form, summed in a different order inside the tensor core: y1 agrees to fp32 rounding, not
// bit for bit (tests/test_gpu_fused12.py holds the per-pixel form to bit-identity with the separate kernels).
// Opt-in (GEECO_CONV12_PAIR=1): the LSU wavefronts drop from 78 % to 43 %, but training is slower (322 vs 266 us: the
// epilogue waits for ring slots the y1 store still holds) and inference gains only 2.4 % (3.38 vs 3.46 ms per 3072 images).
template <bool PAIR, int S1, int RING>
__global__ void __launch_bounds__(THREADS, 1)
conv12_fused_kernel(const C12Args a, const __grid_constant__ CUtensorMap w1map, const __grid_constant__ CUtensorMap w2map,
                    const __grid_constant__ CUtensorMap y1map) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a1 = smem;
  uint8_t* b1 = a1 + S1 * A1_BYTES;
  uint8_t* b2 = b1 + B1_BYTES;
  uint8_t* ring = b2 + B2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + RING * SLOT_BYTES);
  uint64_t* a_full = bars;                 // [S1]
  uint64_t* a_empty = a_full + S1;         // [S1]
  uint64_t* t1_full = a_empty + S1;        // [NB1]
  uint64_t* t1_empty = t1_full + NB1;      // [NB1]
  uint64_t* y_full = t1_empty + NB1;       // [RING]
  uint64_t* y_empty = y_full + RING;       // [RING]
  uint64_t* t2_full = y_empty + RING;      // [NB2]
  uint64_t* t2_empty = t2_full + NB2;      // [NB2]
  uint64_t* w_full = t2_empty + NB2;       // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(w_full + 1);
  float* bias2_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  uint8_t* y2_stage = reinterpret_cast<uint8_t*>(bars) + 1024;     // [4 warps][3 KB]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x / a.cpg, lb = blockIdx.x - group * a.cpg;
  const long long units = (long long)a.M * 128;
  const int u_lo = (int)(units * lb / a.cpg), u_hi = (int)(units * (lb + 1) / a.cpg);
  const bool store_y1 = a.y1 != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S1; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < NB1; ++b) { mbar_init(&t1_full[b], 1); mbar_init(&t1_empty[b], PAIR ? 256 : 128); }
    for (int r = 0; r < RING; ++r) { mbar_init(&y_full[r], 256); mbar_init(&y_empty[r], store_y1 ? 2 : 1); }
    for (int b = 0; b < NB2; ++b) { mbar_init(&t2_full[b], 1); mbar_init(&t2_empty[b], 128); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  // im2col columns >= 37 and the pad pairs of every ring slot are never written: zero everything once
  for (int i = threadIdx.x * 16; i < S1 * A1_BYTES; i += THREADS * 16) *reinterpret_cast<uint4*>(a1 + i) = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x * 16; i < RING * SLOT_BYTES; i += THREADS * 16) *reinterpret_cast<uint4*>(ring + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < C2) bias2_s[threadIdx.x] = a.bias2[(long long)group * a.bias2_group_stride + threadIdx.x];
  __syncthreads();
  // conv1's bias rides in the GEMM: im2col column 36 (pixel pairs: 48) is a constant 1.0 (the packed weights hold the
  // bias there)
  for (int i = threadIdx.x; i < S1 * 128; i += THREADS) {
    const int st_i = i >> 7, r = i & 127;
    if (PAIR) *reinterpret_cast<uint16_t*>(a1 + st_i * A1_BYTES + r * 128 + ((6 ^ (r & 7)) << 4)) = 0x3f80;
    else *reinterpret_cast<uint16_t*>(a1 + st_i * A1_BYTES + r * 128 + ((4 ^ (r & 7)) << 4) + 8) = 0x3f80;
  }
  fence_proxy_async();
  if (warp == 16) tmem_alloc(tmem_ptr_s, 512);
  if (warp == 18 && lane == 0) { tma_prefetch_desc(&w1map); tma_prefetch_desc(&w2map); if (store_y1) tma_prefetch_desc(&y1map); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const long long gimg0 = (long long)group * a.M;               // first image of this encoder group

  if (PAIR && warp < 4) {
    // ===================== conv1 producers on pixel pairs: one y1 row (128 pairs) per stage =====================
    const int t = threadIdx.x;
    const uint32_t sw = (uint32_t)t & 7u;
    const uint32_t a_row0 = smem_u32(a1) + (uint32_t)t * 128u;
    const bool okl = t >= 1, okr = t <= 126;
    uint32_t s = 0, sphase = 0;
    for (int u = u_lo; u < u_hi; ++u) {
      int r_begin, r_end;
      unit_rows(u, u_lo, r_begin, r_end);
      const long long gi = gimg0 + (u >> 7);
      for (int r = r_begin; r <= r_end; ++r) {
        // pixel (r - 1, 2t) of the image; a window row / pixel outside the image copies zero bytes
        const char* sp = reinterpret_cast<const char*>(a.x0) + (((gi * HW + (r - 1)) * HW) + 2 * t) * 8;
        mbar_wait(&a_empty[s], sphase ^ 1u);
        const uint32_t d0 = a_row0 + s * (uint32_t)A1_BYTES;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const bool ok = (unsigned)(r - 1 + ky) < (unsigned)HW;
          const uint32_t ca = d0 + (((uint32_t)(2 * ky) ^ sw) << 4), cb = d0 + (((uint32_t)(2 * ky + 1) ^ sw) << 4);
          const char* p = ok ? sp : reinterpret_cast<const char*>(a.x0) + 8;
          cp_async16(ca, p, ok ? 16u : 0u);                                   // pixels 2t, 2t+1
          cp_async8(cb, p - 8, (ok && okl) ? 8u : 0u);                         // pixel 2t-1
          cp_async8(cb + 8, ok && okr ? p + 16 : p, (ok && okr) ? 8u : 0u);    // pixel 2t+2
          sp += HW * 8;
        }
        cp_async_mbar_arrive_noinc(&a_full[s]);
        if (++s == S1) { s = 0; sphase ^= 1u; }
      }
    }
  } else if (warp < 4) {
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
  } else if (PAIR && warp < 12) {
    // ===================== conv1 epilogue on pixel pairs: thread = (pair, pixel h of the pair) of the y1 row =====================
    // (one set of 4 warps per row with both pixels per thread was the first version: twice the per-row latency, the
    //  ring ran dry)
    const int we = warp - 4, h = we >> 2, quad = we & 3;
    const uint32_t pair = (uint32_t)(quad * 32 + lane);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)h * 32u;
    const uint32_t slot_off = pair * 128u, psw = pair & 7u;
    const uint32_t ring_u32 = smem_u32(ring);
    uint32_t q = 0;                                             // y1 rows produced so far by this CTA
    for (int u = u_lo; u < u_hi; ++u) {
      int r_begin, r_end;
      unit_rows(u, u_lo, r_begin, r_end);
      const long long gi = gimg0 + (u >> 7);
      for (int r = r_begin; r <= r_end; ++r, ++q) {
        const uint32_t slot = q % RING, use = q / RING;
        const uint32_t buf = q % NB1P, tuse = q / NB1P;
        mbar_wait(&t1_full[buf], tuse & 1u);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * 64u, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&t1_empty[buf]);
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        if (a.bits1) {
          uint32_t w = 0;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t acc = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = (acc >> 1) | ((o[hh * 8 + i] + 0x7fff7fffu) & 0x80008000u);
            w |= (((acc >> 8) & 0xffu) | ((acc >> 16) & 0xff00u)) << (16 * hh);
          }
          a.bits1[(gi * HW + r) * HW + 2 * pair + (uint32_t)h] = w;
        }
        // the slot's previous row has been consumed by conv2 (and read by the y1 store)
        mbar_wait(&y_empty[slot], (use & 1u) ^ 1u);
        const uint32_t srow = ring_u32 + slot * (uint32_t)SLOT_BYTES + slot_off;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(srow + ((((uint32_t)h * 4u + (uint32_t)j) ^ psw) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        fence_proxy_async();
        mbar_arrive(&y_full[slot]);
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
      const long long gi = gimg0 + (u >> 7);
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
        if (a.bits1) {
          uint32_t w = 0;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t acc = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = (acc >> 1) | ((o[hh * 8 + i] + 0x7fff7fffu) & 0x80008000u);
            w |= (((acc >> 8) & 0xffu) | ((acc >> 16) & 0xff00u)) << (16 * hh);
          }
          a.bits1[(gi * HW + r) * HW + x] = w;
        }
        // the slot's previous row has been consumed by conv2 (and read by the y1 store)
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
    // ===================== conv2 epilogue: thread = pixel ox of the y2 row =====================
    const int quad = warp & 3;
    const int ox = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + 256u;
    const uint32_t stage_u32 = smem_u32(y2_stage + quad * Y2_STAGE);
    uint32_t i2 = 0;
    for (int u = u_lo; u < u_hi; ++u, ++i2) {
      const uint32_t buf = i2 % NB2, use = i2 / NB2;
      mbar_wait(&t2_full[buf], use & 1u);
      tc_fence_after();
      uint32_t v[48];
      tmem_ld32(lane_addr + buf * 64u, v);
      tmem_ld16(lane_addr + buf * 64u + 32u, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&t2_empty[buf]);
      uint32_t o[24];
#pragma unroll
      for (int i = 0; i < 24; ++i)
        o[i] = pack_bf16x2_relu(__uint_as_float(v[2 * i]) + bias2_s[2 * i], __uint_as_float(v[2 * i + 1]) + bias2_s[2 * i + 1]);
      const long long p2 = ((gimg0 + (u >> 7)) * 128 + (u & 127)) * 128 + ox;
      // the warp's 32 pixels are 3 KB of contiguous y2: staged in shared memory (linear) and written by ONE bulk copy
      // instead of 96 per-lane 32-byte sectors (the LSU wavefronts were the busiest unit of this kernel: 70 %)
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 6; ++j) st_shared_v4(stage_u32 + (uint32_t)lane * 96u + (uint32_t)j * 16u, o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        bulk_store(a.y2 + (p2 - lane) * C2, stage_u32, (uint32_t)Y2_STAGE);
        bulk_commit();
      }
      if (a.bits2) {
#pragma unroll
        for (int ck = 0; ck < 3; ++ck) {
          uint32_t acc = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc = (acc >> 1) | ((o[ck * 8 + i] + 0x7fff7fffu) & 0x80008000u);
          a.bits2[p2 * 3 + ck] = (unsigned short)(((acc >> 8) & 0xffu) | ((acc >> 16) & 0xff00u));
        }
      }
    }
    if (lane == 0) bulk_wait0();                                   // global writes performed before the CTA retires
  } else if (PAIR && (warp == 16 || warp == 19)) {
    // ===================== conv1 MMA issuers on pixel pairs: warp 16 takes the even rows q, warp 19 the odd ones ==========
    const uint32_t idesc1 = make_idesc_bf16(128, 2 * C1, 0, 0);
    const uint64_t dtempl = make_desc_sw128(0, 16, 1024);
    const uint32_t a1_16 = smem_u32(a1) >> 4, b1_16 = smem_u32(b1) >> 4;
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint32_t par = warp == 16 ? 0u : 1u;
    uint32_t q = 0;
    for (int u = u_lo; u < u_hi; ++u) {
      int r_begin, r_end;
      unit_rows(u, u_lo, r_begin, r_end);
      for (int r = r_begin; r <= r_end; ++r, ++q) {
        if ((q & 1u) != par) continue;
        const uint32_t s = q % S1, sphase = (q / S1) & 1u, buf = q % NB1P, bphase = (q / NB1P) & 1u;
        mbar_wait(&t1_empty[buf], bphase ^ 1u);
        mbar_wait(&a_full[s], sphase);
        tc_fence_after();
        const uint32_t a16 = a1_16 + s * (uint32_t)(A1_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j)      // K = 3 window rows x 16 + the bias column: four K = 16 steps
            tc_mma(tmem_base + buf * 64u, dtempl | (uint64_t)(a16 + 2 * j), dtempl | (uint64_t)(b1_16 + 2 * j), idesc1, j != 0 ? 1u : 0u);
          tc_commit(&a_empty[s]);
          tc_commit(&t1_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 16 || warp == 19) {
    // ===================== conv1 MMA issuers (whole warp runs the loop, one elected lane issues) =====================
    // warp 16 issues the left half rows (even tiles), warp 19 the right ones: one warp spent ~2000 of the 2550 cycles
    // of a unit on its four tiles (dependent uniform-datapath instructions, ~9 cycles each)
    // Two issuing warps, one per layer: a single warp issuing both layers was the bound of the first version (ncu: every
    // other role waiting on it, the tensor pipe 38 % busy, ~400 dependent instructions per unit in one warp).
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
    // ===================== conv2 MMA issuer: unit u on the ring rows q0, q1, q2 =====================
    const uint32_t idesc2 = make_idesc_bf16(128, C2, 0, 0);
    const uint64_t dtempl = make_desc_sw128(0, 16, 1024);
    const uint32_t b2_16 = smem_u32(b2) >> 4, ring16 = smem_u32(ring) >> 4;
    mbar_wait(w_full, 0);
    tc_fence_after();
    uint32_t q = 0, buf = 0, bphase = 0;
    int last_q2 = -1;
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
      mbar_wait(&t2_empty[buf], bphase ^ 1u);
      const uint32_t d = tmem_base + 256u + buf * 64u;
      for (int ky = 0; ky < nky; ++ky) {
        const uint32_t qq = (uint32_t)(ky == 0 ? q0 : (ky == 1 ? q1 : q2));
        mbar_wait(&y_full[qq % RING], (qq / RING) & 1u);
        tc_fence_after();
        const uint32_t sa16 = ring16 + (qq % RING) * (uint32_t)(SLOT_BYTES >> 4);
        const uint32_t bb16 = b2_16 + (uint32_t)(ky * 2) * (uint32_t)(B2_SLOT >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j)      // pixel pair ox: kx = 0, 1 (64 channels of the pair)
            tc_mma(d, dtempl | (uint64_t)(sa16 + 2 * j), dtempl | (uint64_t)(bb16 + 2 * j), idesc2, (ky | j) != 0 ? 1u : 0u);
#pragma unroll
          for (int j = 0; j < 2; ++j)      // first pixel of pair ox + 1: kx = 2 (its 32 channels)
            tc_mma(d, dtempl | (uint64_t)(sa16 + 8 + 2 * j), dtempl | (uint64_t)(bb16 + (B2_SLOT >> 4) + 2 * j), idesc2, 1u);
        }
        __syncwarp();
      }
      if (elect_one()) {
        tc_commit(&t2_full[buf]);
        tc_commit(&y_empty[(uint32_t)q0 % RING]);
        tc_commit(&y_empty[(uint32_t)q1 % RING]);
        // the third row is the next unit's first unless the range ends here (a new image starts with its own row 0,
        // and then this unit had no third row)
        if (q2 >= 0 && u + 1 == u_hi) tc_commit(&y_empty[(uint32_t)q2 % RING]);
      }
      __syncwarp();
      if (++buf == NB2) { buf = 0; bphase ^= 1u; }
    }
  } else {
    // ===================== TMA warp: the weights of both layers once, then (training) the y1 stores =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (uint32_t)((PAIR ? 2 : 1) * C1 * 128 + B2_BYTES));
      tma_load_2d(smem_u32(b1), &w1map, w_full, 0, group * (PAIR ? 2 * C1 : C1));
      for (int sl = 0; sl < 6; ++sl) tma_load_2d(smem_u32(b2 + sl * B2_SLOT), &w2map, w_full, sl * 64, group * C2);
    }
    // ring slot -> y1 in HBM, one TMA tensor store per finished row
    if (store_y1 && lane == 0) {
      uint32_t q = 0;
      int prev_slot = -1;
      for (int u = u_lo; u < u_hi; ++u) {
        int r_begin, r_end;
        unit_rows(u, u_lo, r_begin, r_end);
        const long long gi = gimg0 + (u >> 7);
        for (int r = r_begin; r <= r_end; ++r, ++q) {
          const uint32_t slot = q % RING, use = q / RING;
          mbar_wait(&y_full[slot], use & 1u);
          tma_store_2d(&y1map, smem_u32(ring) + slot * (uint32_t)SLOT_BYTES, 0, (int)((gi * HW + r) * 128));
          bulk_commit();
          if (prev_slot >= 0) { bulk_wait_read1(); mbar_arrive(&y_empty[prev_slot]); }
          prev_slot = (int)slot;
        }
      }
      bulk_wait_read0();
      if (prev_slot >= 0) mbar_arrive(&y_empty[prev_slot]);
      bulk_wait0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Whether the fused kernel covers these two layers (everything else runs the separate kernels).
bool tc_conv12_supported(int H, int W, int Cin_pad, int Cout1, int Cout2, int stride1, int stride2, const TcGeom& g1) {
  if (getenv("GEECO_NO_FUSE12")) return false;
  return H == HW && W == HW && Cin_pad == 4 && Cout1 == C1 && Cout2 == C2 && stride1 == 1 && stride2 == 2 && g1.bias_in_k &&
         g1.Kpad == 64 && g1.Ktot == 36;
}

int launch_tc_conv12(const __nv_bfloat16* x0, const CUtensorMap* w1map, const CUtensorMap* w2map, const float* bias2,
                     long long bias2_group_stride, __nv_bfloat16* y1, unsigned short* bits1, __nv_bfloat16* y2,
                     unsigned short* bits2, int G, int M, cudaStream_t st, const CUtensorMap* w1pair_map) {
  if (G < 1 || M < 1) return GEECO_OK;
  const long long pairs = (long long)G * M * HW * 128;
  if (pairs >= (1ll << 31)) { geeco_set_error("conv12: %lld pixel pairs exceed the 2^31 the store coordinates hold", pairs); return GEECO_ERR_INVALID; }
  CUtensorMap y1map;
  memset(&y1map, 0, sizeof(y1map));
  if (y1) {
    int rc = make_tensor_map_2d_sw128(&y1map, y1, 64, pairs, 128, 64, 128);
    if (rc) return rc;
  }
  int cpg = tc_num_sms() / G;
  if (cpg < 1) cpg = 1;
  if ((long long)cpg > (long long)M * 128) cpg = M * 128;
  C12Args a;
  a.x0 = x0; a.y1 = y1; a.bits1 = reinterpret_cast<unsigned int*>(bits1); a.y2 = y2; a.bits2 = bits2; a.bias2 = bias2;
  a.bias2_group_stride = bias2_group_stride; a.M = M; a.cpg = cpg;
  // w1pair_map: the pixel-pair copy of conv1's packed weights (mode 5) -> conv1 on pixel pairs
  const bool deep = y1 != nullptr && !w1pair_map && getenv("GEECO_CONV12_SHALLOW") == nullptr;   // training: 6 stages, 4 slots
  auto kern = w1pair_map ? conv12_fused_kernel<true, 4, 6> : (deep ? conv12_fused_kernel<false, 6, 4> : conv12_fused_kernel<false, 4, 6>);
  const int smem = deep ? smem_bytes(6, 4) : smem_bytes(4, 6);
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  GEECO_LAUNCH((kern), cpg * G, THREADS, smem, st, a, w1pair_map ? *w1pair_map : *w1map, *w2map, y1map);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
