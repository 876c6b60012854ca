// bf16 implicit-GEMM convolution kernels on the Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// Reference ops: tf.layers.conv2d 3x3/SAME of conv_encoder (src/models/e2evmc/graph.py:76-115) and
// the data / weight gradients TensorFlow derives for it (estimator.py:243-244).
//
// tc_nn_kernel   (forward, data-gradient):  D[m][n] = sum_k A[m][k] * Wp[n][k]
//   A  : implicit im2col rows (one output pixel each), gathered from the NHWC bf16 activation with
//        cp.async (16-byte pieces; 8-byte pieces for the 4-channel network input) into a 128x64
//        K-major SWIZZLE_128B shared-memory tile; zero-fill = SAME padding / out-of-range rows
//   Wp : pre-packed bf16 weights [N][Kpad], streamed by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B)
//   D  : fp32 accumulators in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of
//        tile i+1; epilogue = tcgen05.ld -> bias+ReLU (fwd) or ReLU-mask (dgrad) -> bf16 NHWC store
//   Persistent CTAs; warp roles: 0-7 gather producers, 8-11 epilogue, 12 MMA issuer (+TMEM alloc), 13 TMA.
//
// tc_wgrad_kernel (weight gradient):  dW^T[co][(tap,ci)] = sum_pixels G[p][co] * im2col[p][(tap,ci)]
//   both operands are gathered as MN-major SWIZZLE_128B tiles (64 pixels x 64 values per sub-tile);
//   M = 128 output channels, N <= 256 reduction-index values per CTA, split over pixel ranges with a
//   deterministic second-stage reduction; the bias gradient rides along as a column of ones when the
//   reduction index has padding to spare.
//
// Producer cost matters as much as the MMAs here (conv1-conv3 are HBM/L2-bound): row
// decoding uses shifts (all map sizes are powers of two), source pointers and per-tap validity masks
// are computed once per tile, and pieces that are padding for the whole launch are never written
// (shared memory is zeroed once).
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

using namespace tc;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_STAGES = 8;
constexpr int NN_THREADS = 12 * 32 + 64;   // producer + epilogue warps (12), MMA warp, TMA warp

constexpr int A_STAGE_BYTES = BM * BK * 2;            // 16 KB
constexpr int SUB = 64 * 128;                         // one 64-row x 128-byte sub-tile

__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
template <int PIECE>
__device__ __forceinline__ void cp_piece(uint32_t dst_smem, const void* src, bool ok) {
  if (PIECE == 8) cp_async16(dst_smem, src, ok ? 16u : 0u);
  else cp_async8(dst_smem, src, ok ? 8u : 0u);
}

// (img, y, x) of GEMM row m of a group; shifts when the map sizes are powers of two
__device__ __forceinline__ void decode_pixel(const TcGeom& g, uint32_t m, int& img, int& y, int& x) {
  if (g.hw_shift >= 0) {
    img = (int)(m >> g.hw_shift);
    const uint32_t rem = m & ((1u << g.hw_shift) - 1u);
    y = (int)(rem >> g.w_shift);
    x = (int)(rem & ((1u << g.w_shift) - 1u));
  } else {
    const uint32_t hw = (uint32_t)(g.Hm * g.Wm);
    img = (int)(m / hw);
    const uint32_t rem = m - (uint32_t)img * hw;
    y = (int)(rem / (uint32_t)g.Wm);
    x = (int)(rem - (uint32_t)y * (uint32_t)g.Wm);
  }
}
__device__ __forceinline__ void split_k(const TcGeom& g, int k, int& tap, int& ch) {
  if (g.cs_shift >= 0) { tap = k >> g.cs_shift; ch = k & (g.Cs - 1); }
  else { tap = k / g.Cs; ch = k - tap * g.Cs; }
}
__device__ __forceinline__ void zero_smem(uint8_t* base, int bytes) {
  for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(base + i) = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// epilogue shared by the gather/TMA kernel and the row-resident kernel: TMEM -> registers -> global.
//   warp (q, half) reads TMEM lanes [32q, 32q+32) (hardware rule: q = warp index mod 4) and the 16-column chunks
//   c0 = 16*(NHALF*i + half); a set of 4*NHALF warps serves the (tile, class) items tl with tl % nsets == set
// ------------------------------------------------------------------------------------------------
// EPI: compile-time epilogue of the two hot cases (TC_EPI_BIAS_RELU / TC_EPI_MASK with a bf16 destination only);
// EPI_GENERIC keeps every option (runtime `epi`, optional fp32 copy, optional bf16 destination).
constexpr int EPI_GENERIC = 7;
template <int NHALF, int EPI>
__device__ __forceinline__ void nn_epilogue(const TcGeom& g, const TcClasses& cl, const float* bias_s,
                                            const __nv_bfloat16* __restrict__ mask, __nv_bfloat16* __restrict__ dst,
                                            float* __restrict__ dst_f32, int epi, int tiles_per_group, int tiles_flat, int nbuf,
                                            uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t Mg, int q,
                                            int half, int lane, int set, int nsets, unsigned short* __restrict__ bits_out) {
  constexpr bool MASK = EPI == TC_EPI_MASK;
  constexpr bool MBITS = EPI == TC_EPI_MASKBITS;
  const unsigned short* __restrict__ mbits = reinterpret_cast<const unsigned short*>(mask);
  const int chunks_per_pix = g.Ntot >> 4;
  const int nsplit = g.nsplit, ntot = g.Ntot;
  constexpr bool GEN = EPI == EPI_GENERIC;
  const int BN = g.Nn;
  const int ncls = cl.ncls;
  const int row = q * 32 + lane;
  const int c_first = half * 16;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  // ring position / phase / owner set of the (tile, class) item, carried incrementally (no divisions)
  uint32_t buf = 0, bphase = 0;
  int turn = 0, group = 0, group_end = tiles_per_group;
  for (int flat = blockIdx.x; flat < tiles_flat; flat += gridDim.x) {
    if (ncls == 1 && turn != set) {
      // single-class launch whose tile belongs to another set: only the ring position moves (the per-tile decoding
      // below was a quarter of conv1's instructions when every set ran it for every tile)
      if (++buf == (uint32_t)nbuf) { buf = 0; bphase ^= 1u; }
      if (++turn == nsets) turn = 0;
      continue;
    }
    while (flat >= group_end) { ++group; group_end += tiles_per_group; }
    const uint32_t m = (uint32_t)(flat - (group_end - tiles_per_group)) * BM + row;
    const bool valid = m < Mg;
    int img = 0, y = 0, x = 0;
    if (valid) decode_pixel(g, m, img, y, x);
    // destination pixel of class (0, 0); a class adds dy0 rows and dx0 columns
    // a virtual group = (encoder, column tile): images and destination pixels belong to the encoder
    const int enc = nsplit == 1 ? group : group / nsplit;
    const int ncol0 = nsplit == 1 ? 0 : (group - enc * nsplit) * BN;
    const long long pix00 = (((long long)enc * g.imgs_per_group + img) * g.Hd + y * g.dsy) * g.Wd + x * g.dsx;
    const float* bias = bias_s + group * BN;
    for (int c = 0; c < ncls; ++c) {
      const uint32_t my_buf = buf, my_phase = bphase;
      const bool mine = turn == set;
      if (++buf == (uint32_t)nbuf) { buf = 0; bphase ^= 1u; }
      if (++turn == nsets) turn = 0;
      if (!mine) continue;
      const TcCls& kc = cl.c[c];
      const long long pix = pix00 + kc.dy0 * g.Wd + kc.dx0;
      const long long off = pix * ntot + ncol0;
      const long long bidx = pix * chunks_per_pix + (ncol0 >> 4);          // first mask word of this row's columns
      uint32_t pbits = 0, bits_hold = 0;
      if (MBITS && valid && c_first < BN) pbits = __ldg(mbits + bidx + (c_first >> 4));
      // the ReLU mask of this warp's first chunk is fetched before the accumulator wait (hides the DRAM latency)
      uint32_t pm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (MASK && valid && c_first < BN) ldg256_nc(mask + off + c_first, pm);
      mbar_wait(&tmem_full[my_buf], my_phase);
      tc_fence_after();
      const uint32_t taddr = lane_addr + my_buf * (uint32_t)BN;
      for (int c0 = c_first; c0 < BN; c0 += 16 * NHALF) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (EPI == TC_EPI_BIAS_RELU || (GEN && (epi == TC_EPI_BIAS_RELU || epi == TC_EPI_BIAS))) {
            const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bb = b4[i];
              f[4 * i] += bb.x; f[4 * i + 1] += bb.y; f[4 * i + 2] += bb.z; f[4 * i + 3] += bb.w;
            }
            if (GEN && (epi == TC_EPI_BIAS_RELU || epi == TC_EPI_RELU)) {      // (compile-time modes: ReLU inside the bf16 pack)
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
          } else if (MBITS) {
            uint32_t bits = pbits;
            if (c0 != c_first) bits = __ldg(mbits + bidx + (c0 >> 4));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (!((bits >> i) & 1u)) f[2 * i] = 0.f;
              if (!((bits >> (8 + i)) & 1u)) f[2 * i + 1] = 0.f;
            }
          } else if (MASK) {
            uint32_t mw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) mw[i] = pm[i];
            if (c0 != c_first) ldg256_nc(mask + off + c0, mw);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // post-ReLU activations are >= 0 (never -0): "y > 0" == any bit of the bf16 set
              if ((mw[i] & 0xffffu) == 0) f[2 * i] = 0.f;
              if ((mw[i] >> 16) == 0) f[2 * i + 1] = 0.f;
            }
          }
          if (!GEN || dst) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o[i] = (EPI == TC_EPI_BIAS_RELU || EPI == TC_EPI_RELU) ? pack_bf16x2_relu(f[2 * i], f[2 * i + 1])
                                                                      : pack_bf16x2(f[2 * i], f[2 * i + 1]);
            stg256(dst + off + c0, o);
            if ((EPI == TC_EPI_BIAS_RELU || EPI == TC_EPI_RELU) && bits_out) {
              // 1-bit ReLU mask of the STORED values: halfword != 0, flags gathered per halfword lane
              // both halfwords of o[i] are non-negative bf16 (<= 0x7fff): h + 0x7fff has bit 15 set iff h != 0, and the
              // sum cannot carry into the other half.  Shifting the accumulator right once per word leaves word i's two
              // flags at bits 8+i and 24+i: three instructions per word.
              uint32_t acc = 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) acc = (acc >> 1) | ((o[i] + 0x7fff7fffu) & 0x80008000u);
              const uint32_t b16 = ((acc >> 8) & 0xffu) | ((acc >> 16) & 0xff00u);
              const int ck = c0 >> 4;
              if (NHALF == 1 && !((chunks_per_pix | (BN >> 4)) & 1)) {
                // this warp owns every chunk of its rows: two chunks = one aligned 32-bit store (full sectors per warp)
                if (!(ck & 1)) bits_hold = b16;
                else *reinterpret_cast<uint32_t*>(bits_out + bidx + ck - 1) = bits_hold | (b16 << 16);
              } else {
                bits_out[bidx + ck] = (unsigned short)b16;
              }
            }
          }
          if (GEN && dst_f32) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<float4*>(dst_f32 + off + c0 + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[my_buf]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Lean epilogue of the two largest activations of the network (conv1 forward: ReLU + mask bits out; conv2 data
// gradient: mask bits in), selected by TcGeom::fast32: N = Ntot = 32, power-of-two maps, every group a whole number of
// tiles, < 2^31 destination pixels.  The generic epilogue above spent ~390 instructions per warp and item there (ncu:
// conv1 forward at 72 % issue-slot use, 65 % of its instructions in the epilogue; conv2 data gradient stalled a third of
// its samples on the second chunk's mask load).  Here: the warp walks ITS items only, a row is linear in
// (flat tile, lane), the row's 32 mask bits are ONE 32-bit load issued before the accumulator wait, the accumulator is
// read with one 32-column tcgen05.ld and released BEFORE the arithmetic and the stores, and the mask is applied to the
// packed words (shift + byte-permute with sign replication + and: no predicates).
// ------------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_n32(const TcGeom& g, const TcClasses& cl, const unsigned short* __restrict__ mbits,
                                             __nv_bfloat16* __restrict__ dst, int tiles_flat, int nbuf, uint32_t tmem_base,
                                             uint64_t* tmem_full, uint64_t* tmem_empty, int q, int lane, int set, int nsets,
                                             unsigned short* __restrict__ bits_out, const CUtensorMap* omap, uint8_t* stage) {
  static_assert(EPI == TC_EPI_RELU || EPI == TC_EPI_MASKBITS, "epilogue_n32: ReLU (+ mask bits out) or mask bits in");
  // stage != nullptr: the warp's 32 rows x 64 bytes go through a private 2 KB staging tile (SWIZZLE_64B: conflict-free
  // 128-bit shared stores with a fixed register group per instruction) and ONE TMA tensor store instead of 64 per-lane
  // 32-byte sectors: the LSU wavefronts of the row-per-lane stores were the busiest unit of these kernels (ncu:
  // l1tex__data_pipe_lsu_wavefronts 75-81 %).  No CTA-level barrier is involved (round 1's staged epilogue lost to two
  // named barriers per tile): lane 0 owns the bulk group, the warp synchronises with __syncwarp.
  const uint32_t stage_u32 = stage ? smem_u32(stage) : 0u;
  const uint32_t dshift = g.dsx == 2 ? 1u : 0u;
  const int ncls = cl.ncls;
  const uint32_t row = (uint32_t)(q * 32 + lane);
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t hw_shift = (uint32_t)g.hw_shift, w_shift = (uint32_t)g.w_shift;
  const uint32_t hw_mask = (1u << hw_shift) - 1u, w_mask = (1u << w_shift) - 1u;
  const uint32_t Hd = (uint32_t)g.Hd, Wd = (uint32_t)g.Wd, dsy = (uint32_t)g.dsy, dsx = (uint32_t)g.dsx;
  const uint32_t* __restrict__ mb32 = reinterpret_cast<const uint32_t*>(mbits);
  uint32_t* __restrict__ bo32 = reinterpret_cast<uint32_t*>(bits_out);
  uint4* __restrict__ dst4 = reinterpret_cast<uint4*>(dst);
  // item it = j * ncls + c (j-th tile of this CTA, class c); this warp's set serves it % nsets == set
  int c = set, flat = blockIdx.x;
  uint32_t buf = (uint32_t)set, bphase = 0;
  while (c >= ncls) { c -= ncls; flat += gridDim.x; }
  while (buf >= (uint32_t)nbuf) { buf -= (uint32_t)nbuf; bphase ^= 1u; }
  while (flat < tiles_flat) {
    const uint32_t L = (uint32_t)flat * BM + row;                     // row index over all groups
    const uint32_t gimg = L >> hw_shift, rem = L & hw_mask;
    const uint32_t y = rem >> w_shift, x = rem & w_mask;
    const TcCls& kc = cl.c[c];
    const uint32_t pix = (gimg * Hd + y * dsy + (uint32_t)kc.dy0) * Wd + x * dsx + (uint32_t)kc.dx0;
    uint32_t bits = 0;
    if (EPI == TC_EPI_MASKBITS) bits = __ldg(mb32 + pix);             // two 16-channel chunks = one word per pixel
    mbar_wait(&tmem_full[buf], bphase);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(lane_addr + buf * 32u, v);
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(&tmem_empty[buf]);                                    // the MMA warp may refill while we pack and store
    uint32_t o[16];
    if (EPI == TC_EPI_MASKBITS) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t bh = h ? bits >> 16 : bits;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // bit i (channel 2i) -> byte 0's sign, bit 8+i (channel 2i+1) -> byte 1's sign; 0x9988 replicates them
          uint32_t keep;
          asm("prmt.b32 %0, %1, 0, 0x9988;" : "=r"(keep) : "r"(bh << (7 - i)));
          o[h * 8 + i] = pack_bf16x2(__uint_as_float(v[h * 16 + 2 * i]), __uint_as_float(v[h * 16 + 2 * i + 1])) & keep;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    }
    if (stage) {
      if (lane == 0) bulk_wait_read0();                               // the previous store has read the staging tile
      __syncwarp();
      const uint32_t sw = ((uint32_t)lane >> 1) & 3u, srow = stage_u32 + (uint32_t)lane * 64u;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        st_shared_v4(srow + (((uint32_t)ch ^ sw) << 4), o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {                                                // lane 0's row is the first of the warp's 32
        tma_store_3d(omap, stage_u32, 0, (int)(pix & ((1u << dshift) - 1u)), (int)(pix >> dshift));
        bulk_commit();
      }
    } else {
      uint4* d = dst4 + (size_t)pix * 4;                              // 32 channels = 64 bytes per pixel
      stg256(d, o);
      stg256(d + 2, o + 8);
    }
    if (EPI == TC_EPI_RELU && bits_out) {
      // 1-bit ReLU mask of the stored values (layout of TC_EPI_MASKBITS): halfword h >= 0, h + 0x7fff has bit 15 set iff
      // h != 0; shifting the accumulator right once per word leaves word i's flags at bits 8+i and 24+i
      uint32_t w = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = (acc >> 1) | ((o[h * 8 + i] + 0x7fff7fffu) & 0x80008000u);
        const uint32_t b16 = ((acc >> 8) & 0xffu) | ((acc >> 16) & 0xff00u);
        w |= b16 << (16 * h);
      }
      bo32[pix] = w;
    }
    c += nsets;
    while (c >= ncls) { c -= ncls; flat += gridDim.x; }
    buf += (uint32_t)nsets;
    while (buf >= (uint32_t)nbuf) { buf -= (uint32_t)nbuf; bphase ^= 1u; }
  }
  if (stage && lane == 0) bulk_wait0();                               // global writes performed before the CTA retires
}

// MMA issue loop of tc_nn_kernel.  The whole warp runs it (warp-uniform control flow keeps the descriptors in uniform
// registers), one elected lane issues; ring positions and phases are carried incrementally.  KSTEPS = K = 16 steps
// per 64-wide k-block that hold data.
template <int KSTEPS>
__device__ __forceinline__ void nn_mma_loop(const TcClasses& cl, int ncls, int tiles_flat, int stages, int nbuf, int BN,
                                            uint32_t tmem_base, uint32_t a_base_u32, uint32_t b_base_u32, int b_stage_bytes,
                                            uint64_t* full, uint64_t* empty, uint64_t* tmem_full, uint64_t* tmem_empty) {
  const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
  const uint64_t dtempl = make_desc_sw128(0, 16, 1024);
  const uint32_t a_base16 = a_base_u32 >> 4, b_base16 = b_base_u32 >> 4;
  const uint32_t bstage16 = (uint32_t)b_stage_bytes >> 4;
  uint32_t s = 0, sphase = 0, buf = 0, bphase = 0;
  for (int flat = blockIdx.x; flat < tiles_flat; flat += gridDim.x) {
    for (int c = 0; c < ncls; ++c) {
      const int nkb = cl.c[c].Kpad / BK;
      mbar_wait(&tmem_empty[buf], bphase ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + buf * (uint32_t)BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[s], sphase);
        tc_fence_after();
        const uint32_t a16 = a_base16 + s * (uint32_t)(A_STAGE_BYTES >> 4);
        const uint32_t b16 = b_base16 + s * bstage16;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < KSTEPS; ++j)
            tc_mma(d, dtempl | (uint64_t)(a16 + 2 * j), dtempl | (uint64_t)(b16 + 2 * j), idesc, (kb | j) != 0 ? 1u : 0u);
          tc_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
      }
      if (elect_one()) tc_commit(&tmem_full[buf]);
      __syncwarp();
      if (++buf == (uint32_t)nbuf) { buf = 0; bphase ^= 1u; }
    }
  }
}

// MMA issue loop of tc_rows_kernel (see there): per tile one stage of source rows, per class `nsteps` shifted-window
// MMAs of KSTEPS K = 16 steps against resident weight k-blocks.
// PAIRS (forward stride-2 layer on pixel pairs): steps alternate between a full pair (kx = 0, 1: four K = 16 steps)
// and the first pixel of the next pair (kx = 2: its 32 channels = two steps, the other half of the k-block meets zero
// weights) -> 18 instead of 24 MMAs per tile.
template <int KSTEPS, bool PAIRS>
__device__ __forceinline__ void rows_mma_loop(const TcRowProg& rp, int ncls, int tiles_per_group, int tiles_flat, int stages,
                                              int nbuf, int BN, uint32_t tmem_base, uint32_t a_base_u32, uint32_t b_base_u32,
                                              int stage_bytes, int b_slot_bytes, uint64_t* full, uint64_t* empty,
                                              uint64_t* tmem_full, uint64_t* tmem_empty, uint64_t* wfull) {
  const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
  const uint64_t dtempl = make_desc_sw128(0, 16, 1024);
  const uint32_t a_base16 = a_base_u32 >> 4, b_base16 = b_base_u32 >> 4;
  const uint32_t stage16 = (uint32_t)stage_bytes >> 4, slot16 = (uint32_t)b_slot_bytes >> 4;
  uint32_t wphase = 0, s = 0, sphase = 0, buf = 0, bphase = 0;
  int group = 0, group_end = tiles_per_group, cur_group = -1;
  for (int flat = blockIdx.x; flat < tiles_flat; flat += gridDim.x) {
    while (flat >= group_end) { ++group; group_end += tiles_per_group; }
    if (group != cur_group) { mbar_wait(wfull, wphase); wphase ^= 1u; cur_group = group; }
    mbar_wait(&full[s], sphase);
    tc_fence_after();
    const uint32_t a_stage16 = a_base16 + s * stage16;
    for (int c = 0; c < ncls; ++c) {
      mbar_wait(&tmem_empty[buf], bphase ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + buf * (uint32_t)BN;
      const int ns = rp.nsteps[c];
      if constexpr (PAIRS) {
        for (int st = 0; st < ns; st += 2) {
          const uint32_t a16 = a_stage16 + ((uint32_t)rp.a_off[c][st] >> 4);
          const uint32_t b16 = b_base16 + (uint32_t)rp.b_slot[c][st] * slot16;
          const uint32_t a16n = a_stage16 + ((uint32_t)rp.a_off[c][st + 1] >> 4);
          const uint32_t b16n = b_base16 + (uint32_t)rp.b_slot[c][st + 1] * slot16;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              tc_mma(d, dtempl | (uint64_t)(a16 + 2 * j), dtempl | (uint64_t)(b16 + 2 * j), idesc, (st | j) != 0 ? 1u : 0u);
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tc_mma(d, dtempl | (uint64_t)(a16n + 2 * j), dtempl | (uint64_t)(b16n + 2 * j), idesc, 1u);
          }
          __syncwarp();
        }
      } else {
        for (int st = 0; st < ns; ++st) {
          const uint32_t a16 = a_stage16 + ((uint32_t)rp.a_off[c][st] >> 4);
          const uint32_t b16 = b_base16 + (uint32_t)rp.b_slot[c][st] * slot16;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < KSTEPS; ++j)
              tc_mma(d, dtempl | (uint64_t)(a16 + 2 * j), dtempl | (uint64_t)(b16 + 2 * j), idesc, (st | j) != 0 ? 1u : 0u);
          }
          __syncwarp();
        }
      }
      if (elect_one()) tc_commit(&tmem_full[buf]);
      __syncwarp();
      if (++buf == (uint32_t)nbuf) { buf = 0; bphase ^= 1u; }
    }
    if (elect_one()) tc_commit(&empty[s]);
    __syncwarp();
    if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
  }
}

// ------------------------------------------------------------------------------------------------
// forward / data-gradient kernel.  PIECE = bf16 elements per cp.async (8 -> 16 B, 4 -> 8 B)
// ------------------------------------------------------------------------------------------------
// NPW gather-producer warps and 12 - NPW epilogue warps (8/4 by default; 4/8 when the epilogue is the bottleneck:
// one k-block per tile as in conv1), then the MMA warp and the TMA warp
template <int PIECE, int NPW, int EPI>
__global__ void __launch_bounds__(NN_THREADS, 2)
tc_nn_kernel(const TcGeom g, const TcClasses cl, const __grid_constant__ TcMaps maps,
             const __grid_constant__ CUtensorMap amap, const __nv_bfloat16* __restrict__ src,
             const float* __restrict__ bias_all, const __nv_bfloat16* __restrict__ mask,
             __nv_bfloat16* __restrict__ dst, float* __restrict__ dst_f32, int epi, int tiles_per_group,
             int tiles_flat, int tmem_cols, int stages, int nbuf, unsigned short* __restrict__ bits_out,
             const __grid_constant__ CUtensorMap omap) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = g.Nn;
  const int b_stage_bytes = BN * BK * 2;
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + stages * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + stages * b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 8;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_empty + 8);

  constexpr int PROD_THREADS = NPW * 32;
  constexpr int EPI_THREADS = (12 - NPW) * 32;
  // epilogue warps: NPW == 4 (one k-block per tile, conv1): two sets of 4 warps that alternate tiles, each warp
  // owning whole rows (per-tile overhead amortised over both column chunks); else one set, NHALF warps per quadrant
  constexpr int ESETS = NPW == 4 ? 2 : 1;
  constexpr int NHALF = (12 - NPW) / 4 / ESETS;    // epilogue warps per TMEM lane quadrant (column interleave)
  constexpr bool A_TMA = NPW == 0;                 // no gather producers: the TMA warp fetches the A tiles as well
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], PROD_THREADS + 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < nbuf; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_THREADS / ESETS); }
    fence_barrier_init();
  }
  // bias of every group staged once in shared memory (epilogue reads it per tile)
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  if (bias_all)
    for (int i = threadIdx.x; i < g.groups * BN; i += blockDim.x)
      bias_s[i] = bias_all[(long long)((i / BN) / g.nsplit) * g.bias_group_stride + ((i / BN) % g.nsplit) * BN + (i % BN)];
  // padding columns (k >= Ktot) are never written by the producers: start from zeros so that whatever
  // they hold later is finite data (multiplied by zero weights)
  if (!A_TMA) zero_smem(a_base, stages * A_STAGE_BYTES);
  if (!A_TMA && g.bias_in_k) {
    // bias folded into the GEMM: column Ktot of every im2col row is a constant 1.0 (bf16 0x3f80), written once;
    // the producers only ever write the columns below Ktot
    __syncthreads();
    const uint32_t obyte = (uint32_t)g.Ktot * 2;
    for (int i = threadIdx.x; i < stages * BM; i += blockDim.x) {
      const int st_i = i / BM, r = i - st_i * BM;
      uint8_t* p = a_base + st_i * A_STAGE_BYTES + r * 128 + ((((obyte >> 4) ^ (uint32_t)(r & 7))) << 4) + (obyte & 15u);
      *reinterpret_cast<uint16_t*>(p) = 0x3f80;
    }
  }
  fence_proxy_async();
  if (warp == 12) tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
  if (warp == 13 && lane == 0) {
    for (int c = 0; c < cl.ncls; ++c) tma_prefetch_desc(&maps.m[c]);
    if (A_TMA) tma_prefetch_desc(&amap);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const uint32_t Mg = (uint32_t)g.imgs_per_group * (uint32_t)(g.Hm * g.Wm);
  const int ncls = cl.ncls;
  // work item j of this CTA: flat tile (j / ncls) * gridDim.x + blockIdx.x, class j % ncls  -> the classes
  // of one tile run back to back (their source rows and destination lines are shared)

  if (warp < NPW) {
   if constexpr (NPW > 0) {
    // ===================== A producers: implicit-im2col gather =====================
    constexpr int PPR = BK / PIECE;                 // pieces per 128-byte row
    constexpr int ROWS_PER_PASS = PROD_THREADS / PPR;
    constexpr int PASSES = BM / ROWS_PER_PASS;
    const int piece = threadIdx.x % PPR, rsub = threadIdx.x / PPR;
    const uint32_t pbyte = (uint32_t)piece * PIECE * 2;
    uint32_t s = 0, sphase = 0;
    int group = 0, group_end = tiles_per_group;
    if (PIECE == 4 && NPW == 4 && g.rowwin) {
      // conv1 fast path: the tile is 128 consecutive pixels of one image row and K = 36 fits one k-block.
      // Producer thread t owns GEMM row t; a (row, ky) pair is 24 contiguous source bytes = three 8-byte copies.
      // Source and map have the same size (stride 1) and every group is a whole number of tiles, so the source pixel of
      // a row is LINEAR in (flat tile, thread): one 64-bit base per tile, +Ws pixels per ky, immediates per kx.  A tap
      // outside the image keeps its (unread) address and copies zero bytes (cp.async zero-fill).
      const TcCls& k0 = cl.c[0];
      const int dx0 = k0.dx[0], dy0 = k0.dy[0];                   // taps are dy0 + ky, dx0 + kx
      const int Hs = g.Hs, Ws = g.Ws;
      const uint32_t hw_mask = (1u << g.hw_shift) - 1u, w_mask = (1u << g.w_shift) - 1u, w_shift = (uint32_t)g.w_shift;
      const uint32_t row = threadIdx.x, rsw4 = (row & 7u) << 4;
      const uint32_t a_row0 = smem_u32(a_base) + row * 128u;
      const char* const srcb = reinterpret_cast<const char*>(src) + ((long long)dy0 * Ws + dx0 + (long long)row) * 8;
      const long long row_bytes = (long long)Ws * 8;
      (void)group; (void)group_end;
      for (int flat = blockIdx.x; flat < tiles_flat; flat += gridDim.x) {
        const uint32_t L = (uint32_t)flat * BM + row;
        const uint32_t rem = L & hw_mask;
        const int yt = (int)(rem >> w_shift) + dy0, xl = (int)(rem & w_mask) + dx0;   // top-left source pixel
        const bool in0 = xl >= 0, in2 = xl + 2 < Ws;                // kx = 1 is always inside (Ws >= 2)
        const char* sp = srcb + (long long)flat * (BM * 8);
        mbar_wait(&empty[s], sphase ^ 1u);
        const uint32_t a_row = a_row0 + s * (uint32_t)A_STAGE_BYTES;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const bool okr = (unsigned)(yt + ky) < (unsigned)Hs;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t bb = (uint32_t)(ky * 24 + kx * 8);       // byte of (ky, kx) in the 128-byte im2col row
            const uint32_t d = a_row + (((bb >> 4) << 4) ^ rsw4) + (bb & 15u);
            const bool ok = okr && (kx == 0 ? in0 : (kx == 1 ? true : in2));
            cp_async8(d, sp + kx * 8, ok ? 8u : 0u);
          }
          sp += row_bytes;
        }
        cp_async_mbar_arrive_noinc(&full[s]);
        if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
      }
    } else if constexpr (!(PIECE == 4 && NPW == 4)) {     // (4, 4) is only ever launched on the row-window geometry
      for (int flat = blockIdx.x; flat < tiles_flat; flat += gridDim.x) {
        while (flat >= group_end) { ++group; group_end += tiles_per_group; }
        const uint32_t m0 = (uint32_t)(flat - (group_end - tiles_per_group)) * BM;
        // per row: pointer to source pixel (ys, xs) and the coordinates themselves; an out-of-range row
        // gets ys = 1<<20 so that every tap fails the bounds test and zero-fills
        const __nv_bfloat16* rptr[PASSES];
        int rys[PASSES], rxs[PASSES];
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const uint32_t m = m0 + i * ROWS_PER_PASS + rsub;
          rptr[i] = src; rys[i] = 1 << 20; rxs[i] = 0;
          if (m < Mg) {
            int img, y, x;
            decode_pixel(g, m, img, y, x);
            rys[i] = y * g.sy; rxs[i] = x * g.sx;
            rptr[i] = src + ((long long)((group * g.imgs_per_group + img) * g.Hs + rys[i]) * g.Ws + rxs[i]) * g.Cs;
          }
        }
        for (int c = 0; c < ncls; ++c) {
          const TcCls& kc = cl.c[c];
          const int nkb = kc.Kpad / BK;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty[s], sphase ^ 1u);
            const int k = kb * BK + piece * PIECE;
            if (k < kc.Ktot) {
              int tap, ch;
              split_k(g, k, tap, ch);
              const int dyt = kc.dy[tap], dxt = kc.dx[tap];
              const int toff = (dyt * g.Ws + dxt) * g.Cs + ch;
              const uint32_t a_s = smem_u32(a_base + s * A_STAGE_BYTES);
#pragma unroll
              for (int i = 0; i < PASSES; ++i) {
                const int row = i * ROWS_PER_PASS + rsub;
                const bool ok = (unsigned)(rys[i] + dyt) < (unsigned)g.Hs && (unsigned)(rxs[i] + dxt) < (unsigned)g.Ws;
                const uint32_t d = a_s + row * 128 + ((((pbyte >> 4) ^ (uint32_t)(row & 7))) << 4) + (pbyte & 15u);
                cp_piece<PIECE>(d, ok ? (const void*)(rptr[i] + toff) : (const void*)src, ok);
              }
            }
            // asynchronous arrival: fires when this thread's copies have landed; the producer never waits
            cp_async_mbar_arrive_noinc(&full[s]);
            if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
          }
        }
      }
    }
   }
  } else if (warp < 12) {
    if constexpr (NHALF == 1 && (EPI == TC_EPI_RELU || EPI == TC_EPI_MASKBITS)) {
      if (g.fast32) {
        // fast32 == 2: staged TMA-store epilogue, 2 KB per epilogue warp behind the barriers / bias (1024-aligned)
        uint8_t* epi_stage = nullptr;
        if (g.fast32 == 2)
          epi_stage = smem + ((stages * (A_STAGE_BYTES + b_stage_bytes) + 512 + g.groups * BN * 4 + 1023) & ~1023) + (warp - NPW) * 2048;
        epilogue_n32<EPI>(g, cl, reinterpret_cast<const unsigned short*>(mask), dst, tiles_flat, nbuf, tmem_base, tmem_full,
                          tmem_empty, warp & 3, lane, (warp - NPW) >> 2, ESETS, bits_out, &omap, epi_stage);
      } else {
        nn_epilogue<NHALF, EPI>(g, cl, bias_s, mask, dst, dst_f32, epi, tiles_per_group, tiles_flat, nbuf, tmem_base, tmem_full,
                                 tmem_empty, Mg, warp & 3, 0, lane, (warp - NPW) >> 2, ESETS, bits_out);
      }
    } else {
      nn_epilogue<NHALF, EPI>(g, cl, bias_s, mask, dst, dst_f32, epi, tiles_per_group, tiles_flat, nbuf, tmem_base, tmem_full,
                               tmem_empty, Mg, warp & 3, ((warp - NPW) >> 2) % NHALF, lane, ((warp - NPW) >> 2) / NHALF, ESETS,
                               bits_out);
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    // TMA-fed A with one 64-channel chunk per tap: the chunk holds only the channels the source really has (48 ->
    // three K = 16 steps; the fourth would multiply zero-filled columns by zero weights).  The step count is a
    // template argument of the loop: a runtime bound inside the issue loop cost conv2's forward 20 % (measured).
    const int ksteps = (A_TMA && g.Kt == BK) ? (((g.Cs - 1) & 63) >> 4) + 1 : BK / 16;
    if (ksteps == 3)
      nn_mma_loop<3>(cl, ncls, tiles_flat, stages, nbuf, BN, tmem_base, smem_u32(a_base), smem_u32(b_base), b_stage_bytes, full,
                     empty, tmem_full, tmem_empty);
    else
      nn_mma_loop<BK / 16>(cl, ncls, tiles_flat, stages, nbuf, BN, tmem_base, smem_u32(a_base), smem_u32(b_base), b_stage_bytes,
                           full, empty, tmem_full, tmem_empty);
  } else {
    // ===================== weight tiles (and, A_TMA, activation tiles) by TMA (one thread) =====================
    if (lane == 0) {
      uint32_t s = 0, sphase = 0;
      const int cpt = g.Kt / BK;                     // 64-channel chunks per tap (A_TMA)
      int group = 0, group_end = tiles_per_group;
      for (int flat = blockIdx.x; flat < tiles_flat; flat += gridDim.x) {
        while (flat >= group_end) { ++group; group_end += tiles_per_group; }
        int img0 = 0, y0 = 0, x0 = 0;
        if (A_TMA) {
          const uint32_t m0 = (uint32_t)(flat - (group_end - tiles_per_group)) * BM;
          decode_pixel(g, m0, img0, y0, x0);
          img0 += (group / g.nsplit) * g.imgs_per_group;
        }
        for (int c = 0; c < ncls; ++c) {
          const TcCls& kc = cl.c[c];
          const int nkb = kc.Kpad / BK;
          int tap = 0, chunk = 0;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty[s], sphase ^ 1u);
            if (A_TMA) {
              mbar_arrive_expect_tx(&full[s], (uint32_t)(b_stage_bytes + A_STAGE_BYTES));
              // box = 64 channels x the tile's pixel block; out-of-image coordinates zero-fill (= SAME padding)
              tma_load_5d(smem_u32(a_base + s * A_STAGE_BYTES), &amap, &full[s], kc.tc0[tap] + chunk * BK, x0 + kc.twq[tap],
                          kc.thp[tap], y0 + kc.thq[tap], img0);
              if (++chunk == cpt) { chunk = 0; ++tap; }
            } else {
              mbar_arrive_expect_tx(&full[s], (uint32_t)b_stage_bytes);
            }
            tma_load_2d(smem_u32(b_base + s * b_stage_bytes), &maps.m[c], &full[s], kb * BK, group * g.b_rows_per_group);
            if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// row-resident kernel (layers whose GEMM-row grid is >= 128 pixels wide: conv2 forward / data gradient).
// The per-tap kernels above re-read every source pixel once per tap from L2 (2.25x - 9x its size) and the
// weights once per tile; at ~10 TB/s that L2->SM stream, not HBM, is what bounds them.  Here a tile is 128
// consecutive pixels of ONE row: the TMA warp loads the few source rows the tile touches once (hardware
// zero-fill = SAME padding), every tap is an MMA on a window of those rows shifted by whole pixels (128-byte
// rows of the swizzle atom; base_offset in the descriptor), and the packed weights of all taps / classes stay
// resident in shared memory (reloaded only when the encoder changes).
//   warps: EPW*SETS epilogue (set s serves every SETS-th (tile, class) item; EPW = 8 when N <= 32, else 12), then the
//   MMA issuer, then the TMA warp.
// ------------------------------------------------------------------------------------------------
template <int SETS, int EPW, int EPI>
__global__ void __launch_bounds__(SETS * EPW * 32 + 64, SETS == 1 ? 2 : 1)
tc_rows_kernel(const TcGeom g, const TcClasses cl, const TcRowProg rp, const __grid_constant__ TcMaps maps,
               const __grid_constant__ CUtensorMap amap, const float* __restrict__ bias_all,
               const __nv_bfloat16* __restrict__ mask, __nv_bfloat16* __restrict__ dst, float* __restrict__ dst_f32, int epi,
               int tiles_per_group, int tiles_flat, int tmem_cols, int stages, int nbuf,
               unsigned short* __restrict__ bits_out, const __grid_constant__ CUtensorMap omap) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = g.Nn;
  const int b_slot_bytes = BN * BK * 2;
  const int stage_bytes = rp.nrows * rp.pitch;
  uint8_t* b_base = smem;
  uint8_t* a_base = smem + rp.b_slots * b_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 8;
  uint64_t* wfull = tmem_empty + 8;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(wfull + 1);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

  constexpr int W_MMA = EPW * SETS, W_TMA = EPW * SETS + 1;
  constexpr int SUBSETS = EPW == 8 ? 2 : 1;        // N <= 32: sets of 4 warps, each warp owns whole rows
  constexpr int SETW = EPW / SUBSETS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < nbuf; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPW * 32 / SUBSETS); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (bias_all)
    for (int i = threadIdx.x; i < g.groups * BN; i += blockDim.x)
      bias_s[i] = bias_all[(long long)((i / BN) / g.nsplit) * g.bias_group_stride + ((i / BN) % g.nsplit) * BN + (i % BN)];
  if (warp == W_MMA) tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
  if (warp == W_TMA && lane == 0) {
    for (int c = 0; c < cl.ncls; ++c) tma_prefetch_desc(&maps.m[c]);
    tma_prefetch_desc(&amap);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const uint32_t Mg = (uint32_t)g.imgs_per_group * (uint32_t)(g.Hm * g.Wm);
  const int ncls = cl.ncls;

  if (warp < W_MMA) {
    const int set = warp / SETW, ws = warp - set * SETW;
    if constexpr (SETW == 4 && (EPI == TC_EPI_RELU || EPI == TC_EPI_MASKBITS)) {
      if (g.fast32) {
        uint8_t* epi_stage = nullptr;
        if (g.fast32 == 2)
          epi_stage = smem + ((rp.b_slots * b_slot_bytes + stages * stage_bytes + 512 + g.groups * BN * 4 + 1023) & ~1023) + warp * 2048;
        epilogue_n32<EPI>(g, cl, reinterpret_cast<const unsigned short*>(mask), dst, tiles_flat, nbuf, tmem_base, tmem_full,
                          tmem_empty, ws & 3, lane, set, SETS * SUBSETS, bits_out, &omap, epi_stage);
      } else {
        nn_epilogue<1, EPI>(g, cl, bias_s, mask, dst, dst_f32, epi, tiles_per_group, tiles_flat, nbuf, tmem_base, tmem_full,
                            tmem_empty, Mg, ws & 3, 0, lane, set, SETS * SUBSETS, bits_out);
      }
    } else {
      nn_epilogue<SETW / 4, EPI>(g, cl, bias_s, mask, dst, dst_f32, epi, tiles_per_group, tiles_flat, nbuf, tmem_base, tmem_full,
                                 tmem_empty, Mg, ws & 3, ws >> 2, lane, set, SETS * SUBSETS, bits_out);
    }
  } else if (warp == W_MMA) {
    // The whole warp runs the issue loop (warp-uniform control flow -> descriptors live in uniform registers); one
    // elected lane issues.  No divisions: ring positions and phases are carried incrementally.  A single thread
    // issuing ~25 instructions per MMA was what bounded the earlier kernels (profiles/r01_ncu_notes.md).
    // Unit-stride source: a window row holds the Cs real channels of one pixel, zero-filled to 64 (48 -> three K = 16
    // steps); the step count is a template argument of the loop (a runtime bound inside it cost 20 %).
    const int ksteps = g.rows == 1 ? (((g.Cs - 1) & 63) >> 4) + 1 : BK / 16;
    if (g.rows == 2 && g.Cs == 32 && !(rp.nsteps[0] & 1))
      rows_mma_loop<BK / 16, true>(rp, ncls, tiles_per_group, tiles_flat, stages, nbuf, BN, tmem_base, smem_u32(a_base),
                                   smem_u32(b_base), stage_bytes, b_slot_bytes, full, empty, tmem_full, tmem_empty, wfull);
    else if (ksteps == 3)
      rows_mma_loop<3, false>(rp, ncls, tiles_per_group, tiles_flat, stages, nbuf, BN, tmem_base, smem_u32(a_base),
                              smem_u32(b_base), stage_bytes, b_slot_bytes, full, empty, tmem_full, tmem_empty, wfull);
    else
      rows_mma_loop<BK / 16, false>(rp, ncls, tiles_per_group, tiles_flat, stages, nbuf, BN, tmem_base, smem_u32(a_base),
                                    smem_u32(b_base), stage_bytes, b_slot_bytes, full, empty, tmem_full, tmem_empty, wfull);
  } else {
    if (lane == 0) {
      uint32_t it = 0;
      int cur_group = -1;
      for (int flat = blockIdx.x; flat < tiles_flat; flat += gridDim.x, ++it) {
        const int group = flat / tiles_per_group;
        if (group != cur_group) {
          // every MMA that reads the resident weights of the previous encoder has completed once the stage of the
          // previous tile was released
          if (it > 0) mbar_wait(&empty[(it - 1) % stages], ((it - 1) / stages) & 1);
          mbar_arrive_expect_tx(wfull, (uint32_t)(rp.b_slots * b_slot_bytes));
          for (int sl = 0; sl < rp.b_slots; ++sl)
            tma_load_2d(smem_u32(b_base + sl * b_slot_bytes), &maps.m[rp.slot_cls[sl]], wfull, rp.slot_kb[sl] * BK,
                        group * g.b_rows_per_group);
          cur_group = group;
        }
        const uint32_t m0 = (uint32_t)(flat - group * tiles_per_group) * BM;
        int img0, y0, x0;
        decode_pixel(g, m0, img0, y0, x0);
        img0 += group * g.imgs_per_group;
        const int s = it % stages;
        mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
        for (int r = 0; r < rp.nrows; ++r)
          tma_load_5d(smem_u32(a_base + s * stage_bytes + r * rp.pitch), &amap, &full[s], rp.r_c0[r], x0 + rp.w0, rp.r_hp[r],
                      y0 + rp.r_hq[r], img0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// weight-gradient kernel
//   grid.x = mtile * n_chunks + nchunk ; grid.y = split ; grid.z = group
// ------------------------------------------------------------------------------------------------
__device__ int g_wgrad_m128 = 0;       // experiment switch (GEECO_TC_WGRAD_M128=1): keep M = 128 for Cout <= 64

// Single-split launches (the pixel range of a group fits one CTA column: conv7, conv8) write the gradient itself instead
// of a partial: lane = output channel, so for a fixed reduction index the warp's 32 stores are consecutive floats of
// dW[k][co] -- better coalesced than the row-per-lane partial, and the reduce launch disappears.
struct WgDirect { float* dW; float* dbias; long long dw_group_stride, dbias_group_stride; };

// NPROD producer threads: 512 (one CTA per SM, big stages) or 256 (two CTAs per SM when the stage is small)
template <int PIECE, int NPROD>
__global__ void __launch_bounds__(NPROD + 160, NPROD == 256 ? 2 : 1)
tc_wgrad_kernel(const TcGeom g, const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ G,
                float* __restrict__ partial, int Cout, int n_chunks, int kb_per_split, int total_kb, int Mrows_pad,
                int ones_col, int tmem_cols, int stages, int gsub, int nsub_chunk, const WgDirect direct,
                const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap amap, const TcCls kc, int cpt) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int mtile = blockIdx.x / n_chunks, nchunk = blockIdx.x - mtile * n_chunks;
  const int split = blockIdx.y, group = blockIdx.z, groups = gridDim.z;
  const int col0 = nchunk * nsub_chunk * 64;          // first reduction-index column of this CTA
  int nsub = (g.Kpad - col0) / 64;
  if (nsub > nsub_chunk) nsub = nsub_chunk;
  // stage = [2 G sub-tiles][nsub_chunk im2col sub-tiles]; the MMA always reads both G sub-tiles (M = 128)
  const int stage_bytes = (2 + nsub_chunk) * SUB;
  uint8_t* st_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int EPI_WARP0 = NPROD / 32, MMA_WARP = EPI_WARP0 + 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], NPROD == 32 ? 1 : NPROD); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  zero_smem(st_base, stages * stage_bytes);       // pieces that are padding for this CTA are never written
  __syncthreads();
  // bias gradient = sum_p G[p][co] * 1: the first padding column holds 1.0 (bf16 0x3f80) for EVERY row of EVERY
  // stage, written once; rows past the end of the tensor contribute nothing because their G rows are zero-filled
  if (ones_col >= col0 && ones_col < col0 + nsub * 64) {
    const int oc = ones_col - col0, oj = oc >> 6;
    const uint32_t obyte = (uint32_t)(oc & 63) * 2;
    for (int i = threadIdx.x; i < stages * 64; i += blockDim.x) {
      const int s = i >> 6, r = i & 63;
      uint8_t* p = st_base + s * stage_bytes + (2 + oj) * SUB + r * 128 + ((((obyte >> 4) ^ (uint32_t)(r & 7))) << 4) + (obyte & 15u);
      *reinterpret_cast<uint16_t*>(p) = 0x3f80;
    }
  }
  fence_proxy_async();
  if (warp == MMA_WARP) tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const uint32_t Mg = (uint32_t)g.imgs_per_group * (uint32_t)(g.Hm * g.Wm);
  const int kb_lo = split * kb_per_split;
  int kb_hi = kb_lo + kb_per_split;
  if (kb_hi > total_kb) kb_hi = total_kb;
  const int nkb = kb_hi - kb_lo;

  if (NPROD == 32 && warp < EPI_WARP0) {
    // ---- TMA-fed operands (stride-2 layers with >= 48 stored channels: conv3-conv8): no gather arithmetic at all.  A
    // k-block is 64 consecutive GEMM rows = a (bw x bh x bn) block of output pixels; the G sub-tiles are 2-D boxes of
    // G [pixels][Cout], an im2col sub-tile (tap, 64-channel chunk) is ONE 5-D box of the source (hardware zero-fill =
    // SAME padding), exactly what tc_nn_kernel loads as its K-major A tile -- the same bytes are the MN-major B operand
    // here.  The software gather spent ~290 instructions per thread and k-block and was issue-bound (r02 notes).
    if (lane == 0) {
      tma_prefetch_desc(&gmap);
      tma_prefetch_desc(&amap);
      const int sub0 = col0 >> 6;
      int ng = 0;
      for (int j = 0; j < gsub; ++j) ng += (mtile * 128 + j * 64) < Cout ? 1 : 0;
      int nreal = 0;
      for (int j = 0; j < nsub; ++j) nreal += (sub0 + j) * 64 < g.Ktot ? 1 : 0;
      uint32_t s = 0, sphase = 0;
      for (int it = 0; it < nkb; ++it) {
        const uint32_t m0 = (uint32_t)(kb_lo + it) * 64;
        int img0, y0, x0;
        decode_pixel(g, m0, img0, y0, x0);
        img0 += group * g.imgs_per_group;
        mbar_wait(&empty[s], sphase ^ 1u);
        uint8_t* st_p = st_base + s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], (uint32_t)((ng + nreal) * SUB));
        for (int j = 0; j < ng; ++j)
          tma_load_2d(smem_u32(st_p + j * SUB), &gmap, &full[s], mtile * 128 + j * 64, (int)((long long)group * Mg + m0));
        int tap = sub0 / cpt, chunk = sub0 - tap * cpt;
        for (int j = 0; j < nreal; ++j) {
          tma_load_5d(smem_u32(st_p + (2 + j) * SUB), &amap, &full[s], kc.tc0[tap] + chunk * 64, x0 + kc.twq[tap], kc.thp[tap],
                      y0 + kc.thq[tap], img0);
          if (++chunk == cpt) { chunk = 0; ++tap; }
        }
        if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
      }
    }
  } else if (warp < EPI_WARP0 && PIECE == 4 && NPROD == 256 && g.rowwin && Cout == 32 && (Mg & 63u) == 0 && g.dsx == 1 && g.dsy == 1) {
    // ---- conv1 (3/4 input channels, 32 filters): a k-block is 64 consecutive pixels of one image row.  Everything
    // that does not change from k-block to k-block is hoisted; per k-block a thread issues one 16-byte G copy
    // (thread = (row, 16-byte chunk of the 64-byte G row)) and, threads 0..191, the three 8-byte copies of one
    // (row, ky) window.  ~25 instructions per k-block instead of ~290 in the generic producer below.
    const uint32_t grow_t = threadIdx.x >> 2, gchunk = threadIdx.x & 3;
    const uint32_t g_dst = grow_t * 128 + ((gchunk ^ (grow_t & 7u)) << 4);
    const __nv_bfloat16* gsrc = G + ((long long)group * Mg + (long long)kb_lo * 64 + grow_t) * 32 + gchunk * 8;
    const bool xthread = threadIdx.x < 192;
    const uint32_t xrow = threadIdx.x & 63, xky = (threadIdx.x >> 6) % 3;
    const int xdy = g.dy[xky * 3], xdx0 = g.dx[0];
    uint32_t xd[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const uint32_t bb = xky * 24 + kx * 8;
      xd[kx] = 2 * SUB + xrow * 128 + (((bb >> 4) ^ (xrow & 7u)) << 4) + (bb & 15u);
    }
    const int Hs = g.Hs, Ws = g.Ws;
    const uint32_t st_u32 = smem_u32(st_base);
    uint32_t s = 0, sphase = 0;
    uint32_t mb = (uint32_t)kb_lo * 64;
    for (int it = 0; it < nkb; ++it, mb += 64, gsrc += 64 * 32) {
      mbar_wait(&empty[s], sphase ^ 1u);
      const uint32_t sb = st_u32 + s * (uint32_t)stage_bytes;
      cp_async16(sb + g_dst, gsrc, 16u);
      if (xthread) {
        const int img = (int)(mb >> g.hw_shift);
        const uint32_t rem = mb & ((1u << g.hw_shift) - 1u);
        const int y = (int)(rem >> g.w_shift) + xdy, xl = (int)(rem & ((1u << g.w_shift) - 1u)) + (int)xrow + xdx0;
        const bool okr = (unsigned)y < (unsigned)Hs, ok0 = okr && xl >= 0, ok2 = okr && xl + 2 < Ws;
        const __nv_bfloat16* sp = src + ((long long)((group * g.imgs_per_group + img) * Hs + y) * Ws + xl) * 4;
        cp_async8(sb + xd[0], ok0 ? (const void*)sp : (const void*)src, ok0 ? 8u : 0u);
        cp_async8(sb + xd[1], okr ? (const void*)(sp + 4) : (const void*)src, okr ? 8u : 0u);
        cp_async8(sb + xd[2], ok2 ? (const void*)(sp + 8) : (const void*)src, ok2 ? 8u : 0u);
      }
      cp_async_mbar_arrive_noinc(&full[s]);
      if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
    }
  } else if (warp < EPI_WARP0) {
    // ---- G tile mapping: 16-byte chunks, NPROD/8 rows per pass
    constexpr int G_ROWS = NPROD / 8, G_PASSES = 64 / G_ROWS;
    const int g_chunk = threadIdx.x & 7, g_row0 = threadIdx.x >> 3;
    bool g_ok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) g_ok[j] = j < gsub && (mtile * 128 + j * 64 + g_chunk * 8) < Cout;
    const long long g_col = (long long)mtile * 128 + g_chunk * 8;
    // ---- im2col mapping
    constexpr int PPR = 64 / PIECE;
    constexpr int ROWS_PER_PASS = NPROD / PPR;
    constexpr int PASSES = 64 / ROWS_PER_PASS;
    const int piece = threadIdx.x % PPR, rsub = threadIdx.x / PPR;
    const uint32_t pbyte = (uint32_t)piece * PIECE * 2;
    int kind[8], kdy[8], kdx[8], ktoff[8];      // kind 0 = padding / ones column (never written), 1 = gather
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = col0 + j * 64 + piece * PIECE;
      kind[j] = 0; kdy[j] = 0; kdx[j] = 0; ktoff[j] = 0;
      if (j < nsub) {
        if (k < g.Ktot) {
          int tap, ch;
          split_k(g, k, tap, ch);
          kind[j] = 1; kdy[j] = g.dy[tap]; kdx[j] = g.dx[tap];
          ktoff[j] = (g.dy[tap] * g.Ws + g.dx[tap]) * g.Cs + ch;
        }
      }
    }
    // conv1 fast path constants (thread t < 192: row = t & 63, ky = t >> 6)
    const int rw_row = threadIdx.x & 63, rw_ky = (threadIdx.x >> 6) % 3;
    const int rw_dy = g.dy[rw_ky * 3], rw_dx0 = g.dx[0];
    const long long rw_off = ((long long)rw_dy * g.Ws + rw_row + rw_dx0) * 4;
    uint32_t rw_d[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const uint32_t bb = (uint32_t)(rw_ky * 24 + kx * 8);
      rw_d[kx] = rw_row * 128 + (((bb >> 4) ^ (uint32_t)(rw_row & 7)) << 4) + (bb & 15u);
    }
    const long long grow = (long long)group * Mg;
    uint32_t s = 0, sphase = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(&empty[s], sphase ^ 1u);
      const uint32_t sb = smem_u32(st_base + s * stage_bytes);
      const uint32_t mb = (uint32_t)(kb_lo + it) * 64;
      // G tile: rows = pixels, up to 128 output channels of this M tile
#pragma unroll
      for (int gp_i = 0; gp_i < G_PASSES; ++gp_i) {
        const int g_row = gp_i * G_ROWS + g_row0;
        const uint32_t m = mb + g_row;
        const bool rvalid = m < Mg;
        const uint32_t roff = g_row * 128 + ((g_chunk ^ (g_row & 7)) << 4);
        // the G row of GEMM row m is the destination pixel of m (contiguous for forward geometries,
        // every other column for the conv1 pixel-pair classes)
        long long gpix = grow + m;
        if (g.dsx != 1 || g.dsy != 1) {
          int gi, gy, gx;
          decode_pixel(g, rvalid ? m : 0u, gi, gy, gx);
          gpix = ((long long)(group * g.imgs_per_group + gi) * g.Hd + (gy * g.dsy + g.dy0)) * g.Wd + (gx * g.dsx + g.dx0);
        }
        const __nv_bfloat16* gp = G + (gpix * Cout + g_col);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (g_ok[j]) cp_async16(sb + j * SUB + roff, rvalid ? (const void*)(gp + j * 64) : (const void*)G, rvalid ? 16u : 0u);
      }
      if (PIECE == 4 && g.rowwin) {
        // conv1 fast path: 64 consecutive pixels of one image row; threads 0..191 copy one (row, ky) window of
        // 24 contiguous bytes (all tile-invariant quantities were hoisted into rw_*)
        if (threadIdx.x < 192) {
          const int img = (int)(mb >> g.hw_shift);
          const uint32_t rem = mb & ((1u << g.hw_shift) - 1u);
          const int y = (int)(rem >> g.w_shift), x0 = (int)(rem & ((1u << g.w_shift) - 1u));
          const __nv_bfloat16* sp = src + ((long long)((group * g.imgs_per_group + img) * g.Hs + y) * g.Ws + x0) * 4 + rw_off;
          const int xl = x0 + rw_row + rw_dx0;
          const bool okr = mb + rw_row < Mg && (unsigned)(y + rw_dy) < (unsigned)g.Hs;
          const bool ok0 = okr && xl >= 0, ok2 = okr && xl + 2 < g.Ws;
          const uint32_t base = sb + 2 * SUB;
          cp_async8(base + rw_d[0], ok0 ? (const void*)sp : (const void*)src, ok0 ? 8u : 0u);
          cp_async8(base + rw_d[1], okr ? (const void*)(sp + 4) : (const void*)src, okr ? 8u : 0u);
          cp_async8(base + rw_d[2], ok2 ? (const void*)(sp + 8) : (const void*)src, ok2 ? 8u : 0u);
        }
      } else {
        // im2col tile: rows = pixels, 64 reduction-index values per sub-tile
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const int r = i * ROWS_PER_PASS + rsub;
          const uint32_t m = mb + r;
          const bool rvalid = m < Mg;
          int ys = 1 << 20, xs = 0;          // out-of-range row: every tap fails the bounds test
          const __nv_bfloat16* rp = src;
          if (rvalid) {
            int img, y, x;
            decode_pixel(g, m, img, y, x);
            ys = y * g.sy; xs = x * g.sx;
            rp = src + ((long long)((group * g.imgs_per_group + img) * g.Hs + ys) * g.Ws + xs) * g.Cs;
          }
          const uint32_t roff = r * 128 + ((((pbyte >> 4) ^ (uint32_t)(r & 7))) << 4) + (pbyte & 15u);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (kind[j] == 0) continue;
            const uint32_t d = sb + (2 + j) * SUB + roff;
            const bool ok = (unsigned)(ys + kdy[j]) < (unsigned)g.Hs && (unsigned)(xs + kdx[j]) < (unsigned)g.Ws;
            cp_piece<PIECE>(d, ok ? (const void*)(rp + ktoff[j]) : (const void*)src, ok);
          }
        }
      }
      cp_async_mbar_arrive_noinc(&full[s]);
      if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
    }
  } else if (warp < MMA_WARP) {
    const int q = warp & 3;
    // M = 128: TMEM lane = output channel.  M = 64 (Cout <= 64): the accumulator occupies half of every lane quadrant,
    // row r sits in lane (r % 16) + 32 * (r / 16)
    const bool m64 = Cout <= 64 && gridDim.x == n_chunks && !g_wgrad_m128;
    const int row = m64 ? q * 16 + (lane & 15) : q * 32 + lane;
    const int co = (m64 && lane >= 16) ? (1 << 30) : mtile * 128 + row;
    // these warps have nothing to do until the whole pixel range is accumulated: poll rarely (a tight spin from
    // 4-8 idle warps took 20 % of the issue slots of an issue-bound kernel)
    while (!mbar_try(smem_u32(tmem_full), 0)) __nanosleep(2000);
    tc_fence_after();
    float* P = partial + (((long long)split * groups + group) * Mrows_pad + co) * g.Kpad + col0;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < nsub * 64; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if (co < Cout) {
        if (direct.dW) {
          float* dw = direct.dW + (long long)group * direct.dw_group_stride + co;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = col0 + c0 + i;
            if (k < g.Ktot) dw[(long long)k * Cout] = __uint_as_float(v[i]);
            else if (k == ones_col && direct.dbias) direct.dbias[(long long)group * direct.dbias_group_stride + co] = __uint_as_float(v[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(P + c0 + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                     __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      }
    }
  } else {
    // whole warp runs the loop, one elected lane issues (see tc_nn_kernel)
    {
      // N <= 256 per instruction: up to two MMAs per 16-pixel step (columns [0,256) and [256, nsub*64))
      const int n_lo = nsub > 4 ? 256 : nsub * 64, n_hi = nsub > 4 ? (nsub - 4) * 64 : 0;
      // Cout <= 64: M = 64 (one G sub-tile), half the operand reads and tensor work of the zero-padded M = 128
      const int Mm = (Cout <= 64 && gridDim.x == n_chunks && !g_wgrad_m128) ? 64 : 128;
      const uint32_t idesc_lo = make_idesc_bf16(Mm, n_lo, 1, 1);
      const uint32_t idesc_hi = make_idesc_bf16(Mm, n_hi > 0 ? n_hi : 64, 1, 1);
      const uint64_t dtempl = make_desc_sw128(0, SUB, 1024);
      const uint32_t base16 = smem_u32(st_base) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
      uint32_t s = 0, sphase = 0;
      for (int it = 0; it < nkb; ++it) {
        mbar_wait(&full[s], sphase);
        tc_fence_after();
        const uint32_t a16 = base16 + s * stage16;
        const uint32_t b16 = a16 + (2 * SUB >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // 4 x 16 pixels
            const uint64_t ad = dtempl | (uint64_t)(a16 + j * 128);
            tc_mma(tmem_base, ad, dtempl | (uint64_t)(b16 + j * 128), idesc_lo, (it | j) != 0 ? 1u : 0u);
            if (n_hi > 0)
              tc_mma(tmem_base + 256, ad, dtempl | (uint64_t)(b16 + (4 * SUB >> 4) + j * 128), idesc_hi, (it | j) != 0 ? 1u : 0u);
          }
          tc_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == (uint32_t)stages) { s = 0; sphase ^= 1u; }
      }
      if (elect_one()) tc_commit(tmem_full);
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// dW[g][(tap*Cw + ch)][co] = sum_splits partial[s][g][co][tap*Cs + ch] ; bias from the ones column
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ dbias,
                                    int splits, int groups, int Mrows_pad, int Kpad, int Cout, int Cs, int Cw,
                                    int Ktot, int ones_col, long long dw_group_stride, long long dbias_group_stride) {
  pdl_enter();
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

// Transposing variant: grid (k tiles of 32, co tiles of 32, groups), block (32, 32 / CPT).  Partials are read along k
// (their fastest index), dW is written along co (its fastest index) through a 32x32 shared-memory tile; splits are
// added in order.  CPT = output channels per thread: 4 when there are few splits, 1 when there are many.
template <int CPT>
__global__ void __launch_bounds__(1024 / CPT) wgrad_reduce_t_kernel(const float* __restrict__ partial, float* __restrict__ dW,
                                                                    float* __restrict__ dbias, int splits, int groups,
                                                                    int Mrows_pad, int Kpad, int Cout, int Cs, int Cw, int Ktot,
                                                                    int ones_col, long long dw_group_stride,
                                                                    long long dbias_group_stride) {
  pdl_enter();
  constexpr int RY = 32 / CPT;                     // blockDim.y
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, co0 = blockIdx.y * 32, grp = blockIdx.z;
  const long long sstride = (long long)groups * Mrows_pad * Kpad;
  {
    const int k = k0 + threadIdx.x;
    const int col = k < Ktot ? k : (k == Ktot ? ones_col : -1);
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int co = co0 + threadIdx.y + RY * j;
      float sum = 0.f;
      if (co < Cout && col >= 0) {
        const float* p = partial + ((long long)grp * Mrows_pad + co) * Kpad + col;
#pragma unroll 4
        for (int sp = 0; sp < splits; ++sp) sum += p[sp * sstride];
      }
      tile[threadIdx.y + RY * j][threadIdx.x] = sum;
    }
  }
  __syncthreads();
  const int co = co0 + threadIdx.x;
  if (co >= Cout) return;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int k = k0 + threadIdx.y + RY * j;
    const float v = tile[threadIdx.x][threadIdx.y + RY * j];
    if (k < Ktot) {
      const int tap = k / Cs, ch = k - tap * Cs;
      if (ch < Cw) dW[(long long)grp * dw_group_stride + (long long)(tap * Cw + ch) * Cout + co] = v;
    } else if (k == Ktot && dbias && ones_col >= 0) {
      dbias[(long long)grp * dbias_group_stride + co] = v;
    }
  }
}

// conv1 on pixel pairs: dW[ky][kx][c][co] = sum over the two parity classes and splits of the partial column that
// holds (ky, kx, c) in that class (see pack_value modes 2/3); bias gradient from the ones column k = 48.
__global__ void conv1pair_reduce_kernel(const float* __restrict__ part_even, const float* __restrict__ part_odd,
                                        float* __restrict__ dW, float* __restrict__ dbias, int splits, int groups,
                                        int Mrows_pad, int Kpad, int Cin, int Cout, long long dw_group_stride,
                                        long long dbias_group_stride) {
  pdl_enter();
  const int per_group = (9 * Cin + 1) * Cout;
  const int total = per_group * groups;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int grp = i / per_group, e = i - grp * per_group;
    const int kk = e / Cout, co = e - kk * Cout;
    const long long sstride = (long long)groups * Mrows_pad * Kpad;
    float s = 0.f;
    for (int par = 0; par < 2; ++par) {
      const float* P = (par == 0 ? part_even : part_odd) + ((long long)grp * Mrows_pad + co) * Kpad;
      int col;
      if (kk < 9 * Cin) {
        const int tap = kk / Cin, c = kk - tap * Cin, ky = tap / 3, kx = tap - ky * 3;
        // kx = 2*dx + p + 1 - par  with dx = j - 1 (even) / j (odd)  ->  2*j + p = kx + 1 (even) / kx (odd)
        const int q = par == 0 ? kx + 1 : kx;
        const int j = q >> 1, pp = q & 1;
        col = (ky * 2 + j) * 8 + pp * 4 + c;
      } else {
        col = 48;
      }
      for (int sp = 0; sp < splits; ++sp) s += P[sp * sstride + col];
    }
    if (kk < 9 * Cin) dW[(long long)grp * dw_group_stride + (long long)kk * Cout + co] = s;
    else if (dbias) dbias[(long long)grp * dbias_group_stride + co] = s;
  }
}

// value of packed weight element (row r, column k) of a pack job.
//   mode 0 (forward):       out[n][t*Cs + ch]   = W[tap_t][ch][n]
//   mode 1 (data gradient): out[ci][t*Cout + co] = W[tap_t][ci][co]
//   mode 2/3 (conv1 on pixel pairs, even / odd output columns): k = (ky*2 + j)*8 + p*4 + c  ->  W[ky][kx][c][n]
//            with kx = 2*dx + p + 1 - par, dx = j - 1 (even) or j (odd); columns outside the 3x3 window are zero
//   mode 4 (forward stride-2 layer on pixel pairs, 2*Cs == 64): k = (ky*2 + j)*64 + p*Cs + c -> W[ky][2j+p][c][n]
__device__ __forceinline__ float pack_value(const float* __restrict__ Wg0, int mode, int Cin, int Cout, int Cs, int ntaps,
                                            const int* taps, int Kt, int r, int k) {
  if (mode == 2 || mode == 3) {
    const int par = mode - 2;
    if (k >= 48 || r >= Cout) return 0.f;
    const int t = k >> 3, e = k & 7, ky = t >> 1, j = t & 1, pp = e >> 2, c = e & 3;
    const int dx = par == 0 ? j - 1 : j;
    const int kx = 2 * dx + pp + 1 - par;
    if (kx < 0 || kx > 2 || c >= Cin) return 0.f;
    return Wg0[((long long)(ky * 3 + kx) * Cin + c) * Cout + r];
  }
  if (mode == 5) {
    // conv1 on pixel pairs inside the fused conv1 -> conv2 kernel (conv12_fused.cu): row r = p*Cout + co, column
    // k = ky*16 + slot*4 + ch with the window slots = pixels 2t, 2t+1, 2t-1, 2t+2 of the pair's row (pair-relative
    // column cx = 1, 2, 0, 3); output pixel p meets window column cx through tap kx = cx - p
    if (k >= 48 || r >= 2 * Cout) return 0.f;
    const int pp = r / Cout, co = r - pp * Cout;
    const int ky = k >> 4, slot = (k >> 2) & 3, ch = k & 3;
    const int cx = slot == 0 ? 1 : (slot == 1 ? 2 : (slot == 2 ? 0 : 3));
    const int kx = cx - pp;
    if (kx < 0 || kx > 2 || ch >= Cin) return 0.f;
    return Wg0[((long long)(ky * 3 + kx) * Cin + ch) * Cout + co];
  }
  if (mode == 4) {
    // forward stride-2 layer on pixel pairs (2*Cs == 64): k-block (ky, j) holds columns kx = 2j, 2j+1 of kernel row ky
    const int blk = k >> 6, e = k & 63, ky = blk >> 1, j = blk & 1, pp = e / Cs, c = e - pp * Cs;
    const int kx = 2 * j + pp;
    if (ky > 2 || kx > 2 || c >= Cin || r >= Cout) return 0.f;
    return Wg0[((long long)(ky * 3 + kx) * Cin + c) * Cout + r];
  }
  const int per_tap = Kt > 0 ? Kt : (mode == 0 ? Cs : Cout);
  const int t = k / per_tap, c = k - t * per_tap;
  if (t >= ntaps) return 0.f;
  const float* Wg = Wg0 + (long long)taps[t] * Cin * Cout;
  if (mode == 0) return (c < Cin && r < Cout) ? Wg[(long long)c * Cout + r] : 0.f;
  return (r < Cin && c < Cout) ? Wg[(long long)r * Cout + c] : 0.f;
}

__global__ void pack_weights_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int mode, int groups,
                                    long long w_group_stride, int Cin, int Cout, int Cs, int ntaps, int rows, int Kpad,
                                    int Kt, int t0, int t1, int t2, int t3, int t4, int t5, int t6, int t7, int t8,
                                    const float* __restrict__ bias, int bias_col) {
  pdl_enter();
  const int taps[9] = {t0, t1, t2, t3, t4, t5, t6, t7, t8};
  const long long total = (long long)groups * rows * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    const long long rr = i / Kpad;
    const int r = (int)(rr % rows), grp = (int)(rr / rows);
    float v = pack_value(W + (long long)grp * w_group_stride, mode, Cin, Cout, Cs, ntaps, taps, Kt, r, k);
    if (k == bias_col && r < (mode == 5 ? 2 * Cout : Cout)) v = bias[mode == 5 ? r % Cout : r];
    out[i] = __float2bfloat16_rn(v);
  }
}

// One launch for all repack jobs of a step.  Jobs start on PACK_CHUNK boundaries of a virtual index space, so a
// block belongs to exactly one job: the job is looked up once per block and its constants stay in registers.
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackJob* __restrict__ jobs, int njobs,
                                                                   long long grand_total) {
  pdl_enter();
  __shared__ int s_job;
  // grid-stride over PACK_CHUNK-sized chunks: a launch with few blocks (see launch_pack_weights_batched) keeps at
  // most a block or two per SM resident, so the step's own CTAs always find room next to it
  for (long long v0 = (long long)blockIdx.x * PACK_CHUNK; v0 < grand_total; v0 += (long long)gridDim.x * PACK_CHUNK) {
    __syncthreads();
    if (threadIdx.x == 0) {
      int lo = 0, hi = njobs - 1;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (jobs[mid].start <= v0) lo = mid; else hi = mid - 1; }
      s_job = lo;
    }
    __syncthreads();
    const PackJob& jb = jobs[s_job];
    const uint32_t total = (uint32_t)jb.total, Kp = (uint32_t)jb.Kpad, rows = (uint32_t)jb.rows;
    const int mode = jb.mode, Cin = jb.Cin, Cout = jb.Cout, Cs = jb.Cs, ntaps = jb.ntaps, Kt = jb.Kt;
    const float* W = jb.W;
    const float* bias = jb.bias;
    const int bias_col = jb.bias_col;
    const long long wgs = jb.w_group_stride, bgs = jb.b_group_stride;
    __nv_bfloat16* out = jb.out;
    uint32_t i = (uint32_t)(v0 - jb.start) + threadIdx.x;
#pragma unroll 1
    for (int e = 0; e < PACK_CHUNK / 256; ++e, i += 256) {
      if (i >= total) break;
      const uint32_t rr = i / Kp, k = i - rr * Kp;
      const uint32_t grp = rr / rows, r = rr - grp * rows;
      float v = pack_value(W + (long long)grp * wgs, mode, Cin, Cout, Cs, ntaps, jb.taps, Kt, (int)r, (int)k);
      if ((int)k == bias_col && (int)r < (mode == 5 ? 2 * Cout : Cout)) v = bias[(long long)grp * bgs + (mode == 5 ? (int)r % Cout : (int)r)];
      out[i] = __float2bfloat16_rn(v);
    }
  }
}

// One 32 x 32 tile of one tap per block (see PackTile): 256 threads = 32 x 8.  The element-wise kernel above spends
// ~60 instructions per element on index arithmetic (three divisions by runtime values) and, in mode 0, reads W with a
// stride of Cout floats: 55 + 36 us per step for the conv4-conv8 forward and the data-gradient operands when run alone,
// and its thousands of short blocks kept the tail's small kernels waiting for SM slots (r02 timelines).
__global__ void __launch_bounds__(256) pack_tiles_kernel(const PackTile* __restrict__ tiles) {
  pdl_enter();
  __shared__ float t[32][33];
  const PackTile pt = tiles[blockIdx.x];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (pt.transpose) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      t[r][tx] = (r < pt.rows && tx < pt.cols) ? pt.S[(long long)r * pt.s_ld + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = ty + 8 * i;                 // source column = destination row
      if (c < pt.cols && tx < pt.rows) pt.D[(long long)c * pt.d_ld + tx] = __float2bfloat16_rn(t[tx][c]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      if (r < pt.rows && tx < pt.cols) pt.D[(long long)r * pt.d_ld + tx] = __float2bfloat16_rn(pt.S[(long long)r * pt.s_ld + tx]);
    }
  }
}

// one thread per 16-value chunk: bit j = value 2j != 0, bit 8+j = value 2j+1 != 0 (same layout the forward epilogue writes)
__global__ void relu_mask_bits_kernel(const __nv_bfloat16* __restrict__ y, unsigned short* __restrict__ bits, long long chunks) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(y + i * 16)), b = __ldg(reinterpret_cast<const uint4*>(y + i * 16) + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // "y > 0": post-ReLU values are >= 0, so any set magnitude bit; a (never produced) negative value counts as masked
      const uint32_t lo = w[j] & 0xffffu, hi = w[j] >> 16;
      if ((lo & 0x7fffu) && !(lo & 0x8000u)) r |= 1u << j;
      if ((hi & 0x7fffu) && !(hi & 0x8000u)) r |= 1u << (8 + j);
    }
    bits[i] = (unsigned short)r;
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

int g_num_sms = 0;
// SMs the persistent / one-wave grids are sized for.  GEECO_NUM_SMS=<n> sizes them for fewer than the device has:
// under data-parallel training NCCL's all-reduce CTAs share the SMs with the backward kernels, and a grid sized for
// "exactly two CTAs on every SM" then runs its last CTAs as a second wave (bench.py sets it together with
// NCCL_MAX_CTAS when world > 1).
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
    if (const char* e = getenv("GEECO_NUM_SMS")) { const int v = atoi(e); if (v >= 8 && v <= g_num_sms) g_num_sms = v; }
  }
  return g_num_sms;
}

int ilog2_exact(int v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

void finish_geom(TcGeom* g) {
  const int ws = ilog2_exact(g->Wm), hs = ilog2_exact(g->Hm);
  g->w_shift = ws;
  g->hw_shift = (ws >= 0 && hs >= 0) ? ws + hs : -1;
  if (g->hw_shift < 0) g->w_shift = -1;
  g->cs_shift = ilog2_exact(g->Cs);
}

// dynamic shared memory per SM we plan with: 227 KB is the per-CTA maximum; two co-resident CTAs each
// also pay 1 KB of system-reserved shared memory out of the 228 KB, so plan with 222 KB in total
constexpr size_t SMEM_BUDGET = 222 * 1024;

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

static PFN_encodeTiled encode_tiled_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// Can the A operand of this geometry be fetched by TMA?  One tile = 128 consecutive GEMM rows = a
// (bw x bh x bn) block of pixels (bw pixels of a row, bh rows, bn images), so that every k-block is ONE
// 5-D box of 64 channels; out-of-image coordinates are zero-filled by the hardware (= SAME padding).
// A stride-2 source is viewed as [img][H/2][2][W/2][2*Cs] (row pairs, pixel pairs), which needs taps >= 0
// (TensorFlow's SAME padding of an even-sized stride-2 layer pads only after).
static bool tc_use_tma(const TcGeom& g) {
  static const bool disabled = getenv("GEECO_TC_NO_TMA") != nullptr;
  if (disabled) return false;
  if (g.Cs % 8 || g.Cs < 16 || g.hw_shift < 0) return false;
  if (((g.Cs + 63) / 64 * 64 - g.Cs) * 3 > (g.Cs + 63) / 64 * 64) return false;   // > 1/3 of every k-block would be padding
  if (g.sx != g.sy || (g.sx != 1 && g.sx != 2)) return false;
  if (g.Wm > 128 && g.Wm % 128) return false;
  if (g.sx == 2) {
    if ((g.Hs | g.Ws) & 1) return false;
    for (int t = 0; t < g.ntaps; ++t) if (g.dy[t] < 0 || g.dx[t] < 0) return false;
  }
  return true;
}
static bool tc_rows_enabled() {
  static const bool off = getenv("GEECO_TC_NO_ROWS") != nullptr || getenv("GEECO_TC_NO_TMA") != nullptr;
  return !off;
}
static void tc_tile_block(const TcGeom& g, int* bw, int* bh, int* bn, int tile_px = 128) {
  *bw = g.Wm < tile_px ? g.Wm : tile_px;
  *bh = g.Hm < tile_px / *bw ? g.Hm : tile_px / *bw;
  *bn = tile_px / (*bw * *bh);
}
// K layout of the packed weights for this geometry: per-tap extent Kt (Cs, or Cs rounded up to 64 for TMA)
static void finish_k(TcGeom* g) {
  g->a_tma = tc_use_tma(*g) ? 1 : 0;
  g->Kt = g->a_tma ? (g->Cs + 63) / 64 * 64 : g->Cs;
  g->Ktot = g->ntaps * g->Kt;
  g->Kpad = (g->Ktot + 63) / 64 * 64;
}
// the same geometry as the software-gather kernels (weight gradient) see it: dense K = ntaps * Cs
static TcGeom gather_view(const TcGeom& g) {
  TcGeom v = g;
  // always at least one padding column: the bias gradient rides in it as a column of ones (one more 64-wide
  // sub-tile when ntaps * Cs is a multiple of 64) instead of two extra column-sum launches per layer
  v.a_tma = 0; v.rows = 0; v.wpack = 0; v.bias_in_k = 0; v.Kt = v.Cs; v.Ktot = v.ntaps * v.Cs; v.Kpad = (v.Ktot + 64) / 64 * 64;
  return v;
}

// 5-D tensor map over a bf16 NHWC activation for the tile blocks of geometry g (see tc_use_tma)
static int make_act_tensor_map(CUtensorMap* map, const void* base, const TcGeom& g, int row_box_pw = 0, int tile_px = 128) {
  PFN_encodeTiled fn = encode_tiled_fn();
  if (!fn) { geeco_set_error("cuTensorMapEncodeTiled not available from the driver"); return GEECO_ERR_CUDA; }
  int bw, bh, bn;
  tc_tile_block(g, &bw, &bh, &bn, tile_px);
  if (row_box_pw > 0) { bw = row_box_pw; bh = 1; bn = 1; }       // row-resident kernel: one row of pw pixels per box
  const cuuint64_t imgs = (cuuint64_t)g.imgs_per_group * g.groups;
  const cuuint64_t row = (cuuint64_t)g.Ws * g.Cs * 2, img = row * g.Hs;
  cuuint64_t gdim[5], gstr[4];
  if (g.sx == 1) {
    gdim[0] = g.Cs; gdim[1] = g.Ws; gdim[2] = 1; gdim[3] = g.Hs; gdim[4] = imgs;
    gstr[0] = (cuuint64_t)g.Cs * 2; gstr[1] = row; gstr[2] = row; gstr[3] = img;
  } else {
    gdim[0] = 2 * g.Cs; gdim[1] = g.Ws / 2; gdim[2] = 2; gdim[3] = g.Hs / 2; gdim[4] = imgs;
    gstr[0] = (cuuint64_t)g.Cs * 4; gstr[1] = row; gstr[2] = 2 * row; gstr[3] = img;
  }
  cuuint32_t box[5] = {64u, (cuuint32_t)bw, 1u, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    geeco_set_error("cuTensorMapEncodeTiled (activation %dx%dx%d, stride %d, box %dx%dx%d) failed with CUresult %d", g.Hs, g.Ws,
                    g.Cs, g.sx, bw, bh, bn, (int)r);
    return GEECO_ERR_CUDA;
  }
  return GEECO_OK;
}

int tc_make_row_tensor_map(CUtensorMap* map, const void* base, const TcGeom& g, int pw) { return make_act_tensor_map(map, base, g, pw); }

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
  g.Nn = Cout; g.nsplit = 1; g.Ntot = Cout; g.Hd = Ho; g.Wd = Wo; g.dsy = 1; g.dsx = 1;
  g.imgs_per_group = imgs_per_group; g.groups = groups; g.b_rows_per_group = Cout; g.bias_group_stride = Cout;
  finish_geom(&g);
  finish_k(&g);
  if (tc_rows_enabled() && stride == 2 && Cs == 32 && Wo % 128 == 0 && !((H | W) & 1) && g.hw_shift >= 0 && pt == 0 && pl == 0) {
    g.rows = 2; g.wpack = 4; g.a_tma = 0; g.Kt = 64; g.Ktot = g.Kpad = 6 * 64;
  }
  g.rowwin = (Cs == 4 && stride == 1 && Wo % 128 == 0 && g.hw_shift >= 0) ? 1 : 0;
  g.bias_in_k = (g.rowwin && g.Kpad > g.Ktot && !getenv("GEECO_TC_NO_BIAS_IN_K")) ? 1 : 0;
  return g;
}

// conv1 (stride 1, <= 4 input channels stored as 4) on pixel pairs: the source is viewed as [imgs, H, W/2, 8]
// and the outputs of even / odd columns form two classes with 6 taps (3 dy x 2 pairs) each, K = 48.
TcGeom tc_conv1pair_geom(int H, int W, int Cout, int imgs_per_group, int groups, int par) {
  TcGeom g;
  memset(&g, 0, sizeof(g));
  g.Hs = H; g.Ws = W / 2; g.Cs = 8; g.Hm = H; g.Wm = W / 2; g.sy = 1; g.sx = 1; g.ntaps = 6;
  for (int ky = 0; ky < 3; ++ky)
    for (int j = 0; j < 2; ++j) { g.dy[ky * 2 + j] = ky - 1; g.dx[ky * 2 + j] = par == 0 ? j - 1 : j; }
  g.Ktot = 48; g.Kpad = 64; g.Kt = 8; g.a_tma = 0;
  g.Nn = Cout; g.nsplit = 1; g.Ntot = Cout; g.Hd = H; g.Wd = W; g.dsy = 1; g.dsx = 2; g.dy0 = 0; g.dx0 = par;
  g.imgs_per_group = imgs_per_group; g.groups = groups; g.b_rows_per_group = Cout; g.bias_group_stride = Cout;
  finish_geom(&g);
  return g;
}

int launch_conv1pair_reduce(const float* part_even, const float* part_odd, float* dW, float* dbias, int splits,
                            int groups, int Mrows_pad, int Kpad, int Cin, int Cout, long long dw_group_stride,
                            long long dbias_group_stride, cudaStream_t st) {
  const int total = (9 * Cin + 1) * Cout * groups;
  GEECO_LAUNCH((conv1pair_reduce_kernel), ceil_div(total, 256), 256, 0, st, part_even, part_odd, dW, dbias, splits, groups, Mrows_pad,
                                                               Kpad, Cin, Cout, dw_group_stride, dbias_group_stride);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
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
  g.Nn = Cin; g.nsplit = 1; g.Ntot = Cin; g.Hd = H; g.Wd = W; g.dsy = stride; g.dsx = stride; g.dy0 = py; g.dx0 = px;
  g.imgs_per_group = imgs_per_group; g.groups = groups; g.b_rows_per_group = Cin; g.bias_group_stride = 0;
  finish_geom(&g);
  finish_k(&g);
  if (tc_rows_enabled() && g.a_tma && g.Kt == 64 && g.Wm % 128 == 0) g.rows = 1;
  *out = g;
  return true;
}

static int next_pow2_cols(int c) {
  int p = 32;
  while (p < c) p *= 2;
  return p;
}

static int check_geom(const TcGeom& g, const char* who) {
  if ((g.Cs % 8) && g.Cs != 4) { geeco_set_error("%s: Cs=%d must be 4 or a multiple of 8", who, g.Cs); return GEECO_ERR_INVALID; }
  if (g.Kpad % 64) { geeco_set_error("%s: Kpad=%d must be a multiple of 64", who, g.Kpad); return GEECO_ERR_INVALID; }
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg >= (1ll << 31) || (long long)g.imgs_per_group * g.groups * g.Hs * g.Ws >= (1ll << 31)) {
    geeco_set_error("%s: tensor has too many pixels for 32-bit indexing", who);
    return GEECO_ERR_INVALID;
  }
  return GEECO_OK;
}

static void fill_class(TcCls* c, const TcGeom& g) {
  c->ntaps = g.ntaps; c->Ktot = g.Ktot; c->Kpad = g.Kpad; c->dy0 = g.dy0; c->dx0 = g.dx0;
  for (int t = 0; t < GEECO_MAX_TAPS; ++t) {
    c->dy[t] = g.dy[t]; c->dx[t] = g.dx[t];
    const int sh = g.sx == 2 ? 1 : 0;       // taps are >= 0 when sh == 1 (tc_use_tma)
    c->tc0[t] = (short)(sh ? (g.dx[t] & 1) * g.Cs : 0);
    c->twq[t] = (short)(sh ? g.dx[t] >> 1 : g.dx[t]);
    c->thp[t] = (short)(sh ? g.dy[t] & 1 : 0);
    c->thq[t] = (short)(sh ? g.dy[t] >> 1 : g.dy[t]);
  }
}

// compile-time epilogue for the hot cases (bf16 destination only), the generic one otherwise
static int tc_epi_template(int epi, const __nv_bfloat16* dst, const float* dst_f32) {
  if (dst && !dst_f32 && (epi == TC_EPI_MASK || epi == TC_EPI_BIAS_RELU || epi == TC_EPI_MASKBITS || epi == TC_EPI_RELU)) return epi;
  return EPI_GENERIC;
}

int tc_num_sms() { return num_sms(); }

int make_tensor_map_2d_sw128(CUtensorMap* map, void* base, int inner, long long rows, long long row_stride_bytes, int box_inner,
                             int box_rows) {
  PFN_encodeTiled fn = encode_tiled_fn();
  if (!fn) { geeco_set_error("cuTensorMapEncodeTiled not available from the driver"); return GEECO_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { geeco_set_error("cuTensorMapEncodeTiled (2-D, %d x %lld) failed with CUresult %d", inner, rows, (int)r); return GEECO_ERR_CUDA; }
  return GEECO_OK;
}

// 3-D tensor map over a 32-channel bf16 destination viewed as [pixels / dsx][dsx][32]: one box = 32 pixels of one
// column parity x 64 bytes, SWIZZLE_64B staging tile (epilogue_n32's TMA store)
static int make_out32_tensor_map(CUtensorMap* map, void* base, long long pixels, int dsx) {
  PFN_encodeTiled fn = encode_tiled_fn();
  if (!fn) { geeco_set_error("cuTensorMapEncodeTiled not available from the driver"); return GEECO_ERR_CUDA; }
  cuuint64_t gdim[3] = {32u, (cuuint64_t)dsx, (cuuint64_t)(pixels / dsx)};
  cuuint64_t gstr[2] = {64u, (cuuint64_t)64 * dsx};
  cuuint32_t box[3] = {32u, 1u, 32u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { geeco_set_error("cuTensorMapEncodeTiled (32-channel destination) failed with CUresult %d", (int)r); return GEECO_ERR_CUDA; }
  return GEECO_OK;
}

// lean N = 32 epilogue (epilogue_n32): every condition it relies on, checked on the host
static int tc_fast32(const TcGeom& g, int epi_t) {
  static const bool off = getenv("GEECO_TC_NO_FAST32") != nullptr;
  if (off || (epi_t != TC_EPI_RELU && epi_t != TC_EPI_MASKBITS)) return 0;
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  const long long dst_pixels = (long long)g.groups * g.imgs_per_group * g.Hd * g.Wd;
  return g.Nn == 32 && g.Ntot == 32 && g.nsplit == 1 && g.hw_shift >= 0 && g.w_shift >= 0 && Mg % BM == 0 &&
         dst_pixels < (1ll << 31) && (long long)g.groups * Mg < (1ll << 31);
}

// builds the row program of a row-resident launch (see tc_rows_kernel) and runs it
static int launch_tc_rows(const TcGeom* gs, int ncls, const TcClasses& cl, const TcMaps& maps, const __nv_bfloat16* src,
                          const float* bias, const __nv_bfloat16* mask, __nv_bfloat16* dst, float* dst_f32, int epi,
                          cudaStream_t st, unsigned short* bits_out) {
  const TcGeom& g = gs[0];
  TcRowProg rp;
  memset(&rp, 0, sizeof(rp));
  if (g.rows == 2) {
    // forward stride-2 layer, source viewed as pixel pairs: rows 2*oy + ky, windows at pair ox (kx = 0, 1) and ox+1 (kx = 2)
    if (ncls != 1) { geeco_set_error("tc_rows: the pixel-pair forward has one class"); return GEECO_ERR_INVALID; }
    rp.nrows = 3; rp.w0 = 0; rp.pw = 136;
    for (int ky = 0; ky < 3; ++ky) { rp.r_c0[ky] = 0; rp.r_hp[ky] = (short)(ky & 1); rp.r_hq[ky] = (short)(ky >> 1); }
    rp.pitch = rp.pw * 128;
    rp.nsteps[0] = 6;
    for (int ky = 0; ky < 3; ++ky)
      for (int j = 0; j < 2; ++j) {
        rp.a_off[0][ky * 2 + j] = ky * rp.pitch + j * 128;
        rp.b_slot[0][ky * 2 + j] = (short)(ky * 2 + j);
        rp.slot_cls[ky * 2 + j] = 0; rp.slot_kb[ky * 2 + j] = (short)(ky * 2 + j);
      }
    rp.b_slots = 6;
  } else {
    // unit-stride source (data gradient of a stride-2 layer): rows = the distinct dy of all classes, windows = dx shifts
    int dys[GEECO_MAX_TAPS], ndy = 0, mindx = 1 << 20, maxdx = -(1 << 20);
    for (int c = 0; c < ncls; ++c)
      for (int t = 0; t < gs[c].ntaps; ++t) {
        bool seen = false;
        for (int i = 0; i < ndy; ++i) seen = seen || dys[i] == gs[c].dy[t];
        if (!seen) dys[ndy++] = gs[c].dy[t];
        if (gs[c].dx[t] < mindx) mindx = gs[c].dx[t];
        if (gs[c].dx[t] > maxdx) maxdx = gs[c].dx[t];
      }
    if (ndy > 4) { geeco_set_error("tc_rows: %d source rows per tile > 4", ndy); return GEECO_ERR_INVALID; }
    rp.nrows = ndy; rp.w0 = mindx; rp.pw = (128 + maxdx - mindx + 7) / 8 * 8;
    rp.pitch = rp.pw * 128;
    for (int i = 0; i < ndy; ++i) { rp.r_c0[i] = 0; rp.r_hp[i] = 0; rp.r_hq[i] = (short)dys[i]; }
    int slot = 0;
    for (int c = 0; c < ncls; ++c) {
      rp.nsteps[c] = gs[c].ntaps;
      for (int t = 0; t < gs[c].ntaps; ++t, ++slot) {
        int r = 0;
        while (dys[r] != gs[c].dy[t]) ++r;
        rp.a_off[c][t] = r * rp.pitch + (gs[c].dx[t] - mindx) * 128;
        rp.b_slot[c][t] = (short)slot;
        rp.slot_cls[slot] = (short)c; rp.slot_kb[slot] = (short)t;
      }
    }
    rp.b_slots = slot;
  }
  if (rp.pw > 256) { geeco_set_error("tc_rows: box of %d pixels > 256", rp.pw); return GEECO_ERR_INVALID; }
  CUtensorMap amap;
  int rc = make_act_tensor_map(&amap, src, g, rp.pw);
  if (rc) return rc;
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  const int tiles_per_group = ceil_div(Mg, BM);
  const int tiles_flat = tiles_per_group * g.groups;
  const int stage_bytes = rp.nrows * rp.pitch;
  const int b_bytes = rp.b_slots * g.Nn * BK * 2;
  const int tail_bytes = 512 + g.groups * g.Nn * 4;
  const int epi_t = tc_epi_template(epi, dst, dst_f32);
  TcGeom gk = g;
  gk.fast32 = tc_fast32(g, epi_t);
  // staged TMA-store epilogue (epilogue_n32): 2 KB per epilogue warp; experiment switch, off unless asked for
  // (the conv2 data gradient then no longer fits two CTAs per SM)
  static const bool rows_tmastore = getenv("GEECO_TC_ROWS_TMASTORE") != nullptr;
  const bool want_stage = gk.fast32 && rows_tmastore;
  // two co-resident CTAs with one epilogue set each when two stages fit twice, else one CTA with two sets
  int per_sm = 2 * (1024 + b_bytes + 2 * stage_bytes + tail_bytes + (want_stage ? 1024 + 8 * 2048 : 0)) <= (int)SMEM_BUDGET ? 2 : 1;
  if (const char* e = getenv("GEECO_TC_ROWS_PERSM")) { const int v = atoi(e); if (v == 1 || (v == 2 && per_sm == 2)) per_sm = v; }
  const int epi_stage_bytes = want_stage ? 1024 + (per_sm == 2 ? 8 : 16) * 2048 : 0;
  int stages = (int)((SMEM_BUDGET / per_sm - 1024 - b_bytes - tail_bytes - epi_stage_bytes) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { geeco_set_error("tc_rows: stage of %d bytes does not fit twice", stage_bytes); return GEECO_ERR_INVALID; }
  int nbuf = 2;
  while (nbuf < 8 && 2 * nbuf * g.Nn <= 512 / per_sm) nbuf *= 2;
  const int tmem_cols = next_pow2_cols(nbuf * g.Nn);
  const size_t smem = 1024 + (size_t)b_bytes + (size_t)stages * stage_bytes + tail_bytes + epi_stage_bytes;
  CUtensorMap omap;
  memset(&omap, 0, sizeof(omap));
  if (want_stage) {
    rc = make_out32_tensor_map(&omap, dst, (long long)g.groups * g.imgs_per_group * g.Hd * g.Wd, g.dsx);
    if (rc) return rc;
    gk.fast32 = 2;
  }
  int ctas = num_sms() * per_sm;
  if (ctas > tiles_flat) ctas = tiles_flat;
#define ROWS_LAUNCH(SETS_, EPW_, MASK_)                                                                                 \
  do {                                                                                                                  \
    auto kern = tc_rows_kernel<SETS_, EPW_, MASK_>;                                                                     \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                       \
    GEECO_LAUNCH((kern), ctas, SETS_ * EPW_ * 32 + 64, smem, st, gk, cl, rp, maps, amap, bias, mask, dst, dst_f32, epi,               \
                                                     tiles_per_group, tiles_flat, tmem_cols, stages, nbuf, bits_out, omap); \
  } while (0)
#define ROWS_LAUNCH_E(SETS_, MASK_) do { if (g.Nn <= 32) ROWS_LAUNCH(SETS_, 8, MASK_); else ROWS_LAUNCH(SETS_, 12, MASK_); } while (0)
#define ROWS_LAUNCH_S(MASK_) do { if (per_sm == 2) ROWS_LAUNCH_E(1, MASK_); else ROWS_LAUNCH_E(2, MASK_); } while (0)
  if (epi_t == TC_EPI_MASK) ROWS_LAUNCH_S(TC_EPI_MASK);
  else if (epi_t == TC_EPI_MASKBITS) ROWS_LAUNCH_S(TC_EPI_MASKBITS);
  else if (epi_t == TC_EPI_BIAS_RELU) ROWS_LAUNCH_S(TC_EPI_BIAS_RELU);
  else ROWS_LAUNCH_S(EPI_GENERIC);
#undef ROWS_LAUNCH_S
#undef ROWS_LAUNCH_E
#undef ROWS_LAUNCH
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}


// ncls geometries that differ only in taps / K extent / destination offset (the parity classes of a
// strided data-gradient) run as ONE launch; ncls == 1 is the plain forward / single-class case
int launch_tc_nn_multi(const TcGeom* gs, const CUtensorMap* const* wmaps, int ncls, const __nv_bfloat16* src,
                       const float* bias, const __nv_bfloat16* mask, __nv_bfloat16* dst, float* dst_f32, int epi,
                       int max_ctas, cudaStream_t st, unsigned short* bits_out, const __nv_bfloat16* const* wptrs) {
  if (gs[0].bias_in_k) {
    // the packed weights carry the bias in column Ktot and the im2col rows a constant 1.0: nothing left to add
    if (epi == TC_EPI_BIAS_RELU) epi = TC_EPI_RELU;
    else if (epi == TC_EPI_BIAS) epi = TC_EPI_STORE;
    bias = nullptr;
  }
  if (epi == TC_EPI_MASKBITS && (!mask || !dst || dst_f32)) {
    geeco_set_error("tc_nn: TC_EPI_MASKBITS needs the bit mask and a bf16 destination only");
    return GEECO_ERR_INVALID;
  }
  if (ncls < 1 || ncls > 4) { geeco_set_error("tc_nn: ncls=%d outside [1,4]", ncls); return GEECO_ERR_INVALID; }
  TcGeom g = gs[0];
  if (g.Nn % 16 || g.Nn < 16 || g.Nn > 256) { geeco_set_error("tc_nn: unsupported N=%d", g.Nn); return GEECO_ERR_INVALID; }
  TcClasses cl;
  memset(&cl, 0, sizeof(cl));
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  cl.ncls = ncls;
  for (int c = 0; c < ncls; ++c) {
    int rc = check_geom(gs[c], "tc_nn");
    if (rc) return rc;
    if (gs[c].Hm != g.Hm || gs[c].Wm != g.Wm || gs[c].Nn != g.Nn || gs[c].Cs != g.Cs || gs[c].Hs != g.Hs || gs[c].Ws != g.Ws ||
        gs[c].a_tma != g.a_tma || gs[c].Kt != g.Kt || gs[c].rows != g.rows) {
      geeco_set_error("tc_nn: classes of one launch must share the row grid, source and N");
      return GEECO_ERR_INVALID;
    }
    fill_class(&cl.c[c], gs[c]);
    maps.m[c] = *wmaps[c];
  }
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg <= 0) return GEECO_OK;
  if (getenv("GEECO_TC_NOSTORE")) { dst = nullptr; dst_f32 = nullptr; }     // experiment: time without output stores
  if (g.rows) {
    return launch_tc_rows(gs, ncls, cl, maps, src, bias, mask, dst, dst_f32, epi, st, bits_out);
  }
  const int tiles_per_group = ceil_div(Mg, BM);
  // Small layers (conv7, conv8 and their data gradients): fewer 128-row tiles than half the SMs (measured: splitting
  // conv6's 96 tiles costs more in repeated A tiles than it gains).  Cut N into 64-column tiles,
  // each a virtual group with its own slice of the packed weights / bias / destination columns: 2-4x the CTAs.
  static const bool no_nsplit = getenv("GEECO_TC_NO_NSPLIT") != nullptr;
  if (!no_nsplit && g.a_tma && wptrs && g.Nn >= 128 && g.Nn % 64 == 0 && 2 * tiles_per_group * g.groups <= num_sms() &&
      g.b_rows_per_group == g.Nn) {
    const int S = g.Nn / 64;
    bool ok = true;
    for (int c = 0; c < ncls && ok; ++c) {
      ok = wptrs[c] != nullptr &&
           make_weight_tensor_map(&maps.m[c], wptrs[c], (long long)g.groups * g.Nn, gs[c].Kpad, 64) == GEECO_OK;
    }
    if (!ok) return GEECO_ERR_CUDA;
    g.Ntot = g.Nn; g.nsplit = S; g.Nn = 64; g.groups *= S; g.b_rows_per_group = 64;
  }
  const int tiles_flat = tiles_per_group * g.groups;
  const int stage_bytes = A_STAGE_BYTES + g.Nn * BK * 2;
  // small-N layers are latency/bandwidth-bound: two CTAs per SM; wide layers: one CTA, deeper ring
  const int tail_bytes = 512 + g.groups * g.Nn * 4;     // barriers + tmem pointer, bias staging
  int per_sm = (g.Nn <= 64 && 2 * (1024 + 4 * stage_bytes + tail_bytes) <= (int)SMEM_BUDGET) ? 2 : 1;
  // accumulator ring in TMEM: as many buffers (power of two, <= 8) as fit this CTA's share of the 512 columns
  int nbuf = 2;
  while (nbuf < 8 && 2 * nbuf * g.Nn <= 512 / per_sm) nbuf *= 2;
  // tuning overrides for experiments (tools/prof_conv.py): GEECO_TC_PERSM / GEECO_TC_NBUF / GEECO_TC_STAGES
  if (const char* e = getenv("GEECO_TC_PERSM")) { const int v = atoi(e); if (v == 1 || v == 2) per_sm = v; }
  if (const char* e = getenv("GEECO_TC_NBUF")) { const int v = atoi(e); if ((v == 2 || v == 4 || v == 8) && v * g.Nn <= 512 / per_sm) nbuf = v; }
  const int tmem_cols = next_pow2_cols(nbuf * g.Nn);
  const int epi_t = tc_epi_template(epi, dst, dst_f32);
  g.fast32 = tc_fast32(g, epi_t);
  // staged TMA-store epilogue (epilogue_n32): 2 KB per epilogue warp behind the barriers
  static const bool no_tmastore = getenv("GEECO_TC_NO_TMASTORE") != nullptr;
  CUtensorMap omap;
  memset(&omap, 0, sizeof(omap));
  int epi_stage_bytes = 0;
  if (g.fast32 && !no_tmastore) {
    int rc = make_out32_tensor_map(&omap, dst, (long long)g.groups * g.imgs_per_group * g.Hd * g.Wd, g.dsx);
    if (rc) return rc;
    g.fast32 = 2;
    epi_stage_bytes = 1024 + 8 * 2048;
  }
  int stages = (int)((SMEM_BUDGET / per_sm - 1024 - tail_bytes - epi_stage_bytes) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (const char* e = getenv("GEECO_TC_STAGES")) { const int v = atoi(e); if (v >= 3 && v <= stages) stages = v; }
  if (stages < 3) { geeco_set_error("tc_nn: stage of %d bytes does not fit 3 times", stage_bytes); return GEECO_ERR_INVALID; }
  const size_t smem = 1024 + (size_t)stages * stage_bytes + tail_bytes + epi_stage_bytes;
  int ctas = num_sms() * per_sm;
  if (const char* e = getenv("GEECO_TC_CTAS")) { const int v = atoi(e); if (v > 0) ctas = v; }
  if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
  if (ctas > tiles_flat) ctas = tiles_flat;
  // warp split: epilogue-heavy (4 producer / 8 epilogue warps) when a tile has a single k-block
  int npw = (cl.c[0].Kpad == BK && ncls == 1) ? 4 : 8;
  if (const char* e = getenv("GEECO_TC_NPW")) { const int v = atoi(e); if (v == 4 || v == 8) npw = v; }
  if (g.Cs == 4 && !g.rowwin) npw = 8;      // the 4-warp variant of the 8-byte gather exists for the row-window path only
  CUtensorMap amap;
  memset(&amap, 0, sizeof(amap));
  if (g.a_tma) {
    TcGeom ga = g;
    ga.groups = gs[0].groups;               // the source holds the real groups (a column tile is not a new image set)
    int rc = make_act_tensor_map(&amap, src, ga);
    if (rc) return rc;
    npw = 0;
  }
#define NN_LAUNCH_M(PIECE_, NPW_, MASK_)                                                                               \
  do {                                                                                                                 \
    auto kern = tc_nn_kernel<PIECE_, NPW_, MASK_>;                                                                     \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
    GEECO_LAUNCH((kern), ctas, NN_THREADS, smem, st, g, cl, maps, amap, src, bias, mask, dst, dst_f32, epi, tiles_per_group,        \
                                         tiles_flat, tmem_cols, stages, nbuf, bits_out, omap);                         \
  } while (0)
#define NN_LAUNCH(PIECE_, NPW_)                                                                                        \
  do {                                                                                                                 \
    if (epi_t == TC_EPI_MASK) NN_LAUNCH_M(PIECE_, NPW_, TC_EPI_MASK);                                                  \
    else if (epi_t == TC_EPI_MASKBITS) NN_LAUNCH_M(PIECE_, NPW_, TC_EPI_MASKBITS);                                     \
    else if (epi_t == TC_EPI_BIAS_RELU) NN_LAUNCH_M(PIECE_, NPW_, TC_EPI_BIAS_RELU);                                   \
    else if (epi_t == TC_EPI_RELU) NN_LAUNCH_M(PIECE_, NPW_, TC_EPI_RELU);                                             \
    else NN_LAUNCH_M(PIECE_, NPW_, EPI_GENERIC);                                                                       \
  } while (0)
  if (npw == 0) NN_LAUNCH(8, 0);
  else if (g.Cs == 4) { if (npw == 4) NN_LAUNCH(4, 4); else NN_LAUNCH(4, 8); }
  else { if (npw == 4) NN_LAUNCH(8, 4); else NN_LAUNCH(8, 8); }
#undef NN_LAUNCH
#undef NN_LAUNCH_M
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

int launch_tc_nn(const TcGeom& g, const CUtensorMap* wmap, const __nv_bfloat16* src, const float* bias,
                 const __nv_bfloat16* mask, __nv_bfloat16* dst, float* dst_f32, int epi, int max_ctas,
                 cudaStream_t st, unsigned short* bits_out, const __nv_bfloat16* wptr) {
  const CUtensorMap* maps[1] = {wmap};
  const __nv_bfloat16* wp[1] = {wptr};
  return launch_tc_nn_multi(&g, maps, 1, src, bias, mask, dst, dst_f32, epi, max_ctas, st, bits_out, wptr ? wp : nullptr);
}

// TMA-fed weight gradient (see tc_wgrad_kernel, NPROD == 32): stride-2 layers whose source tc_nn_kernel can box, whole
// 64-channel blocks of G, k-blocks of 64 pixels that never straddle an encoder group.  OPT-IN (GEECO_TC_WGRAD_TMA=1):
// correct (tests pass with it) but measured slower in situ -- conv3 132 vs 89 us, conv5 44 vs 39 us, conv4 / conv6-8
// unchanged: every tap's box comes from L2 (9 x the source per k-block, ~11 TB/s aggregate = the L2 -> SM limit), while
// the software gather's cp.async.ca copies find the overlapping taps of a k-block in L1.
static bool wgrad_use_tma(const TcGeom& g_in, int Cout) {
  static const bool on = getenv("GEECO_TC_WGRAD_TMA") != nullptr;
  if (!on || g_in.sx != 2 || Cout % 64) return false;
  TcGeom t = g_in;
  t.rows = 0;
  if (!tc_use_tma(t)) return false;
  const long long Mg = (long long)g_in.imgs_per_group * g_in.Hm * g_in.Wm;
  if (Mg % 64 || g_in.hw_shift < 0) return false;
  return true;
}
// the geometry as the TMA-fed kernel sees it: every tap padded to whole 64-channel chunks, one more sub-tile for the
// bias gradient's ones column
static TcGeom wgrad_tma_view(const TcGeom& g) {
  TcGeom v = g;
  v.a_tma = 1; v.rows = 0; v.wpack = 0; v.bias_in_k = 0; v.Kt = (v.Cs + 63) / 64 * 64; v.Ktot = v.ntaps * v.Kt; v.Kpad = v.Ktot + 64;
  return v;
}
static TcGeom wgrad_view(const TcGeom& g, int Cout) { return wgrad_use_tma(g, Cout) ? wgrad_tma_view(g) : gather_view(g); }

struct WgradPlan { int m_tiles, n_chunks, nsub_chunk, splits, kb_per_split, total_kb, Mrows_pad, ones_col, gsub, stages, tmem_cols, per_sm; };

static WgradPlan wgrad_plan(const TcGeom& g, int Cout) {
  WgradPlan p;
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  p.m_tiles = (Cout + 127) / 128;
  const int total_sub = g.Kpad / 64;
  // up to 8 sub-tiles (N = 512 TMEM columns) per CTA; the ring must hold at least 3 stages
  int cap = 8;
  while (cap > 1 && 3 * (2 + cap) * SUB + 1536 > (int)SMEM_BUDGET) --cap;
  static const int env_cap = getenv("GEECO_TC_WGRAD_NSUB") ? atoi(getenv("GEECO_TC_WGRAD_NSUB")) : 0;
  static const int env_minkb = getenv("GEECO_TC_WGRAD_MINKB") ? atoi(getenv("GEECO_TC_WGRAD_MINKB")) : 8;
  if (env_cap > 0 && cap > env_cap) cap = env_cap;
  // few pixels per group (conv6-conv8: 4-64 k-blocks): narrow column chunks = more CTAs, shorter epilogues and fewer
  // (or no) splits; measured in situ (tools/sweep_wgrad.sh): conv8 30 -> 20 us, conv7 35 -> 28 us, conv6 47 -> 40 us
  if (env_cap == 0) {
    const long long kb = (Mg + 63) / 64;
    const int small_cap = kb <= 16 ? 2 : (kb <= 64 ? 4 : cap);
    if (cap > small_cap) cap = small_cap;
  }
  p.n_chunks = (total_sub + cap - 1) / cap;
  p.nsub_chunk = (total_sub + p.n_chunks - 1) / p.n_chunks;
  p.total_kb = (int)((Mg + 63) / 64);
  p.gsub = Cout > 64 ? 2 : 1;
  const int stage_bytes = (2 + p.nsub_chunk) * SUB;
  p.tmem_cols = p.nsub_chunk > 4 ? 512 : (p.nsub_chunk > 2 ? 256 : 128);
  // small stages: two CTAs per SM with 256 producer threads each and a 4-deep ring (measured better than one
  // CTA with a deeper ring, as for the forward kernel)
  p.per_sm = (p.tmem_cols <= 256 && 2 * (1024 + 4 * stage_bytes + 512) <= (int)SMEM_BUDGET) ? 2 : 1;
  p.stages = (int)((SMEM_BUDGET / p.per_sm - 1024 - 512) / stage_bytes);
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  const int base = p.m_tiles * p.n_chunks * g.groups;
  // one wave: never more CTAs than can be resident at once (a 149th CTA on 148 SMs doubles the kernel time)
  int want = (num_sms() * p.per_sm) / base;
  if (want < 1) want = 1;
  if (want > p.total_kb / env_minkb) want = p.total_kb / env_minkb > 0 ? p.total_kb / env_minkb : 1;     // >= 8 k-blocks per split
  p.kb_per_split = (p.total_kb + want - 1) / want;
  p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.Mrows_pad = p.m_tiles * 128;
  p.ones_col = g.Kpad > g.Ktot ? g.Ktot : -1;
  return p;
}

long long tc_wgrad_partial_floats(const TcGeom& g_in, int Cout) {
  const TcGeom g = wgrad_view(g_in, Cout);
  WgradPlan p = wgrad_plan(g, Cout);
  return (long long)p.splits * g.groups * p.Mrows_pad * g.Kpad + 64;
}

// runs the weight-gradient GEMM into `partial` ([splits][groups][Mrows_pad][Kpad] fp32)
static int wgrad_gemm(const TcGeom& g, int Cout, const __nv_bfloat16* src, const __nv_bfloat16* G, float* partial,
                      long long partial_cap, bool want_ones, WgradPlan* plan_out, cudaStream_t st,
                      const WgDirect* direct_in = nullptr) {
  if (Cout % 8 || Cout > 256 * 8) { geeco_set_error("tc_wgrad: unsupported Cout=%d", Cout); return GEECO_ERR_INVALID; }
  int rc = check_geom(g, "tc_wgrad");
  if (rc) return rc;
  WgradPlan p = wgrad_plan(g, Cout);
  *plan_out = p;
  const long long need_floats = (long long)p.splits * g.groups * p.Mrows_pad * g.Kpad + 64;
  if (!partial || need_floats > partial_cap) {
    geeco_set_error("tc_wgrad: partial buffer too small (%lld floats needed, %lld given)", need_floats, partial_cap);
    return GEECO_ERR_WORKSPACE;
  }
  if (p.stages < 3) { geeco_set_error("tc_wgrad: stage does not fit 3 times"); return GEECO_ERR_INVALID; }
  const size_t smem = 1024 + (size_t)p.stages * (2 + p.nsub_chunk) * SUB + 512;
  const int ones = want_ones ? p.ones_col : -1;
  WgDirect direct = {nullptr, nullptr, 0, 0};
  if (direct_in && p.splits == 1) direct = *direct_in;
  dim3 grid(p.m_tiles * p.n_chunks, p.splits, g.groups);
  {
    static int flag_set = -1;
    if (flag_set < 0) {
      const int v = getenv("GEECO_TC_WGRAD_M128") ? 1 : 0;
      CUDA_TRY(cudaMemcpyToSymbol(g_wgrad_m128, &v, sizeof(int)));
      flag_set = v;
    }
  }
  CUtensorMap gmap, amap;
  memset(&gmap, 0, sizeof(gmap));
  memset(&amap, 0, sizeof(amap));
  TcCls kc;
  memset(&kc, 0, sizeof(kc));
  const bool tma = g.a_tma != 0;
  const int cpt = tma ? g.Kt / 64 : 1;
  if (tma) {
    const long long Mg_ = (long long)g.imgs_per_group * g.Hm * g.Wm;
    PFN_encodeTiled fn = encode_tiled_fn();
    if (!fn) { geeco_set_error("cuTensorMapEncodeTiled not available from the driver"); return GEECO_ERR_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)Cout, (cuuint64_t)(Mg_ * g.groups)};
    cuuint64_t gstr[1] = {(cuuint64_t)Cout * 2};
    cuuint32_t box[2] = {64u, 64u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(&gmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(G), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { geeco_set_error("cuTensorMapEncodeTiled (weight-gradient G, %d channels) failed with CUresult %d", Cout, (int)r); return GEECO_ERR_CUDA; }
    rc = make_act_tensor_map(&amap, src, g, 0, 64);
    if (rc) return rc;
    fill_class(&kc, g);
  }
#define WG_LAUNCH(PIECE_, NPROD_)                                                                                   \
  do {                                                                                                              \
    CUDA_TRY(cudaFuncSetAttribute(tc_wgrad_kernel<PIECE_, NPROD_>, cudaFuncAttributePreferredSharedMemoryCarveout,  \
                                  cudaSharedmemCarveoutMaxShared));                                                 \
    CUDA_TRY(cudaFuncSetAttribute(tc_wgrad_kernel<PIECE_, NPROD_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)smem));                                                                      \
    GEECO_LAUNCH((tc_wgrad_kernel<PIECE_, NPROD_>), grid, NPROD_ + 160, smem, st, g, src, G, partial, Cout, p.n_chunks,          \
                                                                      p.kb_per_split, p.total_kb, p.Mrows_pad, ones, \
                                                                      p.tmem_cols, p.stages, p.gsub, p.nsub_chunk, direct, gmap, amap, kc, cpt);  \
  } while (0)
  if (tma) WG_LAUNCH(8, 32);
  else if (g.Cs == 4) { if (p.per_sm == 2) WG_LAUNCH(4, 256); else WG_LAUNCH(4, 512); }
  else { if (p.per_sm == 2) WG_LAUNCH(8, 256); else WG_LAUNCH(8, 512); }
#undef WG_LAUNCH
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

// partial-only variant: the caller performs the reduction (conv1 pixel-pair classes)
int launch_tc_wgrad_partial(const TcGeom& g_in, int Cout, const __nv_bfloat16* src, const __nv_bfloat16* G, float* partial,
                            long long partial_cap, int want_ones, int* splits_out, int* mrows_out, cudaStream_t st) {
  const TcGeom g = gather_view(g_in);       // (the pixel-pair classes of conv1: software gather)
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg <= 0) { geeco_set_error("tc_wgrad_partial: empty problem"); return GEECO_ERR_INVALID; }
  WgradPlan p;
  int rc = wgrad_gemm(g, Cout, src, G, partial, partial_cap, want_ones != 0, &p, st);
  if (rc) return rc;
  *splits_out = p.splits; *mrows_out = p.Mrows_pad;
  return GEECO_OK;
}

int launch_tc_wgrad(const TcGeom& g_in, int Cout, int Cw, const __nv_bfloat16* src, const __nv_bfloat16* G, float* dW,
                    float* dbias, float* partial, long long partial_cap, long long dw_group_stride,
                    long long dbias_group_stride, cudaStream_t st) {
  // software gather (dense K) or TMA-fed operands (every tap padded to whole 64-channel chunks)
  const TcGeom g = wgrad_view(g_in, Cout);
  const long long Mg = (long long)g.imgs_per_group * g.Hm * g.Wm;
  if (Mg <= 0) return GEECO_OK;
  WgradPlan p;
  static const bool no_direct = getenv("GEECO_TC_WGRAD_NO_DIRECT") != nullptr;
  // a single split needs no reduction when the stored channels are the real ones (reduction index = dW row)
  WgDirect direct = {dW, dbias, dw_group_stride, dbias_group_stride};
  const bool can_direct = !no_direct && g.Kt == Cw;
  int rc = wgrad_gemm(g, Cout, src, G, partial, partial_cap, dbias != nullptr, &p, st, can_direct ? &direct : nullptr);
  if (rc) return rc;
  if (dbias && p.ones_col < 0) { geeco_set_error("tc_wgrad: no padding column for the bias gradient"); return GEECO_ERR_INVALID; }
  if (can_direct && p.splits == 1) return GEECO_OK;
  const int ones = dbias ? p.ones_col : -1;
  const dim3 tgrid(ceil_div(g.Ktot + 1, 32), ceil_div(Cout, 32), g.groups);
  const int tiles = (int)(tgrid.x * tgrid.y * tgrid.z);
  if (p.splits <= 12 && tiles >= 2 * num_sms()) {
    // few splits, many outputs (conv5-conv8)
    GEECO_LAUNCH((wgrad_reduce_t_kernel<4>), tgrid, dim3(32, 8), 0, st, partial, dW, dbias, p.splits, g.groups, p.Mrows_pad, g.Kpad, Cout,
                                                           g.Kt, Cw, g.Ktot, ones, dw_group_stride, dbias_group_stride);
  } else if (tiles >= 64) {
    // more splits (conv3, conv4): one output channel per thread, 1024 threads per tile
    GEECO_LAUNCH((wgrad_reduce_t_kernel<1>), tgrid, dim3(32, 32), 0, st, partial, dW, dbias, p.splits, g.groups, p.Mrows_pad, g.Kpad, Cout,
                                                            g.Kt, Cw, g.Ktot, ones, dw_group_stride, dbias_group_stride);
  } else {
    // many splits, few outputs (conv1, conv2): one thread per output, no transpose
    const long long total = (long long)(g.Ktot + 1) * Cout * g.groups;
    int rb = ceil_div(total, 256); if (rb > 148 * 8) rb = 148 * 8;
    GEECO_LAUNCH((wgrad_reduce_kernel), rb, 256, 0, st, partial, dW, dbias, p.splits, g.groups, p.Mrows_pad, g.Kpad, Cout, g.Kt, Cw,
                                            g.Ktot, ones, dw_group_stride, dbias_group_stride);
  }
  geeco_count_launch(1);
  if (dbias && p.ones_col < 0) { geeco_set_error("tc_wgrad: no padding column for the bias gradient"); return GEECO_ERR_INVALID; }
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

int launch_pack_weights(const float* W, __nv_bfloat16* out, int mode, int groups, long long w_group_stride, int Cin,
                        int Cout, int Cs, int ntaps, const int* taps, int rows, int Kpad, int Kt, cudaStream_t st,
                        const float* bias, int bias_col) {
  if (groups != 1 && bias_col >= 0) { geeco_set_error("pack_weights: a bias column needs groups == 1 here"); return GEECO_ERR_INVALID; }
  int t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < ntaps && i < 9; ++i) t[i] = taps[i];
  const long long total = (long long)groups * rows * Kpad;
  int blocks = ceil_div(total, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  GEECO_LAUNCH((pack_weights_kernel), blocks, 256, 0, st, W, out, mode, groups, w_group_stride, Cin, Cout, Cs, ntaps, rows, Kpad, Kt,
                                              t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], bias,
                                              bias ? bias_col : -1);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

// max_blocks > 0: a persistent launch of at most that many blocks (the repack that runs next to the step's first
// kernels on the side stream: thousands of short blocks kept every SM's registers fragmented and conv1's CTAs, which
// need 29 K registers each, could not start until the repack grid was exhausted -- r02 timeline: conv1 began 25 us
// after its input was ready, exactly when the repack ended)
int launch_pack_weights_batched(const PackJob* jobs_dev, int njobs, long long grand_total, cudaStream_t st, int max_blocks) {
  if (njobs <= 0 || grand_total <= 0) return GEECO_OK;
  if (njobs > 64) { geeco_set_error("pack_weights_batched: %d jobs > 64", njobs); return GEECO_ERR_INVALID; }
  if (grand_total % PACK_CHUNK) { geeco_set_error("pack_weights_batched: jobs must start on PACK_CHUNK boundaries"); return GEECO_ERR_INVALID; }
  long long blocks = grand_total / PACK_CHUNK;
  if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
  // same shared-memory carve-out as the tcgen05 kernels it runs next to: an SM switches its L1 / shared split only
  // when it is empty, and with repack blocks always resident it never was -- conv1's CTAs waited for the whole repack
  static bool carve_set = false;
  if (!carve_set) {
    CUDA_TRY(cudaFuncSetAttribute(pack_weights_batched_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve_set = true;
  }
  GEECO_LAUNCH((pack_weights_batched_kernel), (unsigned)blocks, 256, 0, st, jobs_dev, njobs, grand_total);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

bool pack_job_is_plain(const PackJob& j) { return (j.mode == 0 || j.mode == 1) && j.bias_col < 0; }

// tiles of a plain job: per group and tap the [Cin][Cout] matrix of that tap, cut into 32 x 32 tiles
void pack_job_tiles(const PackJob& j, std::vector<PackTile>* out) {
  const int per_tap = j.Kt > 0 ? j.Kt : (j.mode == 0 ? j.Cs : j.Cout);
  for (int g = 0; g < j.groups; ++g)
    for (int t = 0; t < j.ntaps; ++t) {
      const float* S = j.W + (long long)g * j.w_group_stride + (long long)j.taps[t] * j.Cin * j.Cout;
      __nv_bfloat16* D = j.out + (long long)g * j.rows * j.Kpad + (long long)t * per_tap;
      for (int r0 = 0; r0 < j.Cin; r0 += 32)
        for (int c0 = 0; c0 < j.Cout; c0 += 32) {
          PackTile pt;
          pt.S = S + (long long)r0 * j.Cout + c0;
          pt.D = j.mode == 0 ? D + (long long)c0 * j.Kpad + r0 : D + (long long)r0 * j.Kpad + c0;
          pt.s_ld = j.Cout; pt.d_ld = j.Kpad;
          pt.rows = (short)(j.Cin - r0 < 32 ? j.Cin - r0 : 32);
          pt.cols = (short)(j.Cout - c0 < 32 ? j.Cout - c0 : 32);
          pt.transpose = j.mode == 0 ? 1 : 0; pt.pad_ = 0;
          out->push_back(pt);
        }
    }
}

int launch_pack_tiles(const PackTile* tiles_dev, int ntiles, cudaStream_t st) {
  if (ntiles <= 0) return GEECO_OK;
  static bool carve_set = false;
  if (!carve_set) {
    CUDA_TRY(cudaFuncSetAttribute(pack_tiles_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve_set = true;
  }
  GEECO_LAUNCH((pack_tiles_kernel), (unsigned)ntiles, 256, 0, st, tiles_dev);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

int launch_relu_mask_bits(const __nv_bfloat16* y, unsigned short* bits, long long chunks, cudaStream_t st) {
  if (chunks <= 0) return GEECO_OK;
  int blocks = ceil_div(chunks, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  GEECO_LAUNCH((relu_mask_bits_kernel), blocks, 256, 0, st, y, bits, chunks);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st) {
  if (n <= 0) return GEECO_OK;
  int blocks = ceil_div(n, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  GEECO_LAUNCH((f32_to_bf16_kernel), blocks, 256, 0, st, in, out, n);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
int launch_bf16_to_f32(const __nv_bfloat16* in, float* out, long long n, cudaStream_t st) {
  if (n <= 0) return GEECO_OK;
  int blocks = ceil_div(n, 256); if (blocks > 148 * 8) blocks = 148 * 8;
  GEECO_LAUNCH((bf16_to_f32_kernel), blocks, 256, 0, st, in, out, n);
  geeco_count_launch(1);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}
