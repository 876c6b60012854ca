// Rank pooling (dynamic image / dynamic difference) for sm_100a.
//
// Reference: src/models/e2evmc/graph.py:17-55 (`_H`, `_alpha`, `dynimg`) and :392-401 (the two
// call sites of GEECO-F: the K-frame buffer and the [current, target] pair).
//
//   d   = sum_k alpha_k * x_k                       (products rounded, then summed in k order)
//   out = (d - min_hwc d) / (max_hwc d - min_hwc d + 1e-6)      per sample
//
// The op is pure HBM streaming with one per-sample reduction in the middle.  To read the K-frame
// buffer ONCE and write the normalised image ONCE, a thread-block CLUSTER owns one sample: every
// CTA keeps its slice of the un-normalised d in shared memory, CTAs exchange (min, max) through
// distributed shared memory, and the normalised slice is written from shared memory.  Loads are
// 128-bit, read-only, no-L1-allocate; K is a template parameter so all K loads of a position are
// in flight together.
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>
#include <float.h>

namespace cg = cooperative_groups;

struct AlphaTab { float a[16]; };

static constexpr int RP_THREADS = 512;
static constexpr int RP_MAX_CLUSTER = 16;

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ float4 f4_scale(float a, const float4& v) {
  return make_float4(__fmul_rn(a, v.x), __fmul_rn(a, v.y), __fmul_rn(a, v.z), __fmul_rn(a, v.w));
}
__device__ __forceinline__ float4 f4_axpy(const float4& acc, float a, const float4& v) {
  return make_float4(__fadd_rn(acc.x, __fmul_rn(a, v.x)), __fadd_rn(acc.y, __fmul_rn(a, v.y)),
                     __fadd_rn(acc.z, __fmul_rn(a, v.z)), __fadd_rn(acc.w, __fmul_rn(a, v.w)));
}
__device__ __forceinline__ void f4_minmax(const float4& v, float& mn, float& mx) {
  mn = fminf(mn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
  mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
}
__device__ __forceinline__ float4 f4_norm(const float4& d, float mn, float rng) {
  return make_float4(__fdiv_rn(d.x - mn, rng), __fdiv_rn(d.y - mn, rng), __fdiv_rn(d.z - mn, rng),
                     __fdiv_rn(d.w - mn, rng));
}

// Block-wide then cluster-wide (min, max).  `s_red` holds 2*32 floats, `s_mm` 2*RP_MAX_CLUSTER per pair.
__device__ __forceinline__ void cluster_minmax(float& mn, float& mx, float* s_red, float* s_mm, int pair,
                                               int npairs) {
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { s_red[warp * 2] = mn; s_red[warp * 2 + 1] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = lane < nwarps ? s_red[lane * 2] : FLT_MAX;
    mx = lane < nwarps ? s_red[lane * 2 + 1] : -FLT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const unsigned cl = cluster.num_blocks(), me = cluster.block_rank();
    // lane r publishes this CTA's pair into CTA r's shared memory (DSMEM store)
    if ((unsigned)lane < cl) {
      float* remote = cluster.map_shared_rank(s_mm, lane);
      remote[(pair * RP_MAX_CLUSTER + me) * 2] = mn;
      remote[(pair * RP_MAX_CLUSTER + me) * 2 + 1] = mx;
    }
  }
  __syncthreads();
  if (pair == npairs - 1) cluster.sync();   // release/acquire: all CTAs' pairs are visible after this
}
__device__ __forceinline__ void cluster_minmax_read(float& mn, float& mx, const float* s_mm, int pair) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cl = cluster.num_blocks();
  mn = FLT_MAX; mx = -FLT_MAX;
  for (unsigned r = 0; r < cl; ++r) {
    mn = fminf(mn, s_mm[(pair * RP_MAX_CLUSTER + r) * 2]);
    mx = fmaxf(mx, s_mm[(pair * RP_MAX_CLUSTER + r) * 2 + 1]);
  }
}

// ---------------------------------------------------------------------------------------
// generic dynimg: in [N,K,HWC] f32 -> out [N,HWC] f32 ; grid = (cluster, N), cluster dims (CL,1,1)
// ---------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(RP_THREADS) dynimg_cluster_kernel(const float* __restrict__ in,
                                                                    float* __restrict__ out, long long total4,
                                                                    long long per4, AlphaTab al) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_red[64];
  __shared__ float s_mm[2 * RP_MAX_CLUSTER];
  float4* sd = reinterpret_cast<float4*>(smem_raw);
  const long long n = blockIdx.y;
  const long long lo = (long long)blockIdx.x * per4;
  long long hi = lo + per4; if (hi > total4) hi = total4;
  const float4* base = reinterpret_cast<const float4*>(in) + n * K * total4;
  float mn = FLT_MAX, mx = -FLT_MAX;
  constexpr int U = (K >= 8) ? 1 : (K >= 4 ? 2 : 4);
  for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += (long long)U * RP_THREADS) {
    float4 v[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * RP_THREADS;
      if (i < hi) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[u][k] = ldg_stream(base + k * total4 + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * RP_THREADS;
      if (i < hi) {
        float4 acc = f4_scale(al.a[0], v[u][0]);
#pragma unroll
        for (int k = 1; k < K; ++k) acc = f4_axpy(acc, al.a[k], v[u][k]);
        sd[i - lo] = acc;
        f4_minmax(acc, mn, mx);
      }
    }
  }
  cluster_minmax(mn, mx, s_red, s_mm, 0, 1);
  cluster_minmax_read(mn, mx, s_mm, 0);
  const float rng = __fadd_rn(__fsub_rn(mx, mn), 1e-6f);
  float4* o = reinterpret_cast<float4*>(out) + n * total4;
  for (long long i = lo + threadIdx.x; i < hi; i += RP_THREADS) stg_stream(o + i, f4_norm(sd[i - lo], mn, rng));
}

// two-pass fallback (samples too large for a 16-CTA cluster's shared memory)
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int* mm, int n) {
  pdl_enter();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { mm[2 * i] = f2ord(FLT_MAX); mm[2 * i + 1] = f2ord(-FLT_MAX); }
}
template <int K>
__global__ void __launch_bounds__(256) dynimg_pass1_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int* __restrict__ mm, long long total4, AlphaTab al) {
  pdl_enter();
  const long long n = blockIdx.y;
  const float4* base = reinterpret_cast<const float4*>(in) + n * K * total4;
  float4* o = reinterpret_cast<float4*>(out) + n * total4;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    float4 v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = ldg_stream(base + k * total4 + i);
    float4 acc = f4_scale(al.a[0], v[0]);
#pragma unroll
    for (int k = 1; k < K; ++k) acc = f4_axpy(acc, al.a[k], v[k]);
    o[i] = acc;
    f4_minmax(acc, mn, mx);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(mm + 2 * n, f2ord(mn)); atomicMax(mm + 2 * n + 1, f2ord(mx)); }
}
__global__ void __launch_bounds__(256) dynimg_pass2_kernel(float* __restrict__ out, const int* __restrict__ mm,
                                                           long long total4) {
  pdl_enter();
  const long long n = blockIdx.y;
  const float mn = ord2f(mm[2 * n]), mx = ord2f(mm[2 * n + 1]);
  const float rng = __fadd_rn(__fsub_rn(mx, mn), 1e-6f);
  float4* o = reinterpret_cast<float4*>(out) + n * total4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x)
    o[i] = f4_norm(o[i], mn, rng);
}

// ---------------------------------------------------------------------------------------
// fused GEECO-F pre-process (graph.py:387-401): one pass over the K-frame buffer + target frame
//   enc 0: current frame (= frame K-1)             -> x0[0][n]  (channel-padded, OutT)
//   enc 1: dynimg(buffer)                          -> x0[1][n]
//   enc 2: dynimg([current, target])  (alpha -1/2, 1/2)  -> x0[2][n]
// optional fp32 un-padded copies of enc 1 / enc 2 for the 'dynbuff' / 'dyndiff' endpoints.
// Work unit = 4 pixels = C float4 loads per frame.
// ---------------------------------------------------------------------------------------
template <typename OutT, int CP> struct PixPack;
template <> struct PixPack<float, 4> {
  static __device__ __forceinline__ void store(float* dst, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
  }
};
template <> struct PixPack<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void store(__nv_bfloat16* dst, float a, float b, float c, float d) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(dst) = u;
  }
};
template <> struct PixPack<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void store(__nv_bfloat16* dst, float a, float b, float c, float d) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
    uint4 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1); u.z = 0u; u.w = 0u;
    *reinterpret_cast<uint4*>(dst) = u;
  }
};

// values e[0..4*C) of 4 consecutive pixels -> padded pixels (channels >= C are zero)
template <typename OutT, int CP, int C>
__device__ __forceinline__ void store_unit(OutT* dst, const float* e) {
  if constexpr (sizeof(OutT) == 2 && CP == 4) {
    // bf16, 4 channels: the four pixels of a unit are 32 contiguous bytes -> ONE 256-bit store (lanes hold consecutive
    // units: 1 KB contiguous per warp instruction instead of four instructions of 8 bytes at a 32-byte stride)
    uint32_t w[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(e[p * C], e[p * C + 1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(e[p * C + 2], C == 4 ? e[p * C + (C - 1)] : 0.f);
      w[2 * p] = *reinterpret_cast<uint32_t*>(&p0); w[2 * p + 1] = *reinterpret_cast<uint32_t*>(&p1);
    }
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p)
      PixPack<OutT, CP>::store(dst + p * CP, e[p * C], e[p * C + 1], e[p * C + 2], C == 4 ? e[p * C + (C - 1)] : 0.f);
  }
}

// One "float4" of frame data (4 consecutive values of the NHWC stream) in either input format.  uint8 records are
// divided by 255.0f in fp32, as the reference's input pipeline does on the host (src/data/geeco_gym.py:310).
// `lut` (uint8 only): 256 floats in shared memory, lut[b] = float(b) / 255.0f computed once per CTA with the IEEE
// division; a lookup per byte replaces ~60 divisions per 4-pixel unit (the uint8 kernel was compute-bound on them).
template <typename InT> struct FrameLoad;
template <> struct FrameLoad<float> {
  typedef float4 raw_t;
  static __device__ __forceinline__ raw_t raw(const float* base, long long i4) { return ldg_stream(reinterpret_cast<const float4*>(base) + i4); }
  static __device__ __forceinline__ float4 cvt(const raw_t& r) { return r; }
  static __device__ __forceinline__ float4 at(const float* base, long long i4, const float*) {
    return ldg_stream(reinterpret_cast<const float4*>(base) + i4);
  }
  // second pass of the fused kernel: the C float4 of a 4-pixel unit share cache lines with each other and with the
  // neighbouring lanes' units, so these loads allocate in L1 (ld.global.nc)
  static __device__ __forceinline__ float4 at_l1(const float* base, long long i4, const float*) {
    return __ldg(reinterpret_cast<const float4*>(base) + i4);
  }
};
// float(b) / 255.0f, correctly rounded, without a division or a table: q = b * fl(1/255), then one exact-residual
// correction step q + (b - 255 q) * fl(1/255).  Checked exhaustively against the IEEE division for b = 0..255 (all equal;
// the bare product differs for 126 of them).  The 256-entry shared-memory table this replaces made the uint8 kernel
// LDS-bound: 1.4 M random lookups per sample, 3-4-way bank conflicts (1.07 ms per 1024-environment policy chunk).
__device__ __forceinline__ float u8_unit(unsigned int b) {
  const float f = (float)b, rc = 0x1.010102p-8f;
  const float q = __fmul_rn(f, rc);
  return __fmaf_rn(__fmaf_rn(-q, 255.f, f), rc, q);
}
template <> struct FrameLoad<unsigned char> {
  typedef unsigned int raw_t;
  static __device__ __forceinline__ raw_t raw(const unsigned char* base, long long i4) { return __ldg(reinterpret_cast<const unsigned int*>(base) + i4); }
  static __device__ __forceinline__ float4 cvt(const raw_t& w) {
    return make_float4(u8_unit(w & 255u), u8_unit((w >> 8) & 255u), u8_unit((w >> 16) & 255u), u8_unit(w >> 24));
  }
  static __device__ __forceinline__ float4 at(const unsigned char* base, long long i4, const float*) {
    const unsigned int w = __ldg(reinterpret_cast<const unsigned int*>(base) + i4);
    return make_float4(u8_unit(w & 255u), u8_unit((w >> 8) & 255u), u8_unit((w >> 16) & 255u), u8_unit(w >> 24));
  }
  static __device__ __forceinline__ float4 at_l1(const unsigned char* base, long long i4, const float* lut) { return at(base, i4, lut); }
};

// Two CTAs per SM (64 registers per thread).  The kernel is latency-bound; at 80 registers only one 512-thread CTA
// fitted (ncu, uint8 frames: 25 % of the warp slots, 15 clusters of 8 resident, 1.6 TB/s).
template <typename InT, typename OutT, int CP, int C, int K>
__global__ void __launch_bounds__(RP_THREADS, 2) preprocess_geecof_kernel(
    const InT* __restrict__ rgb, const InT* __restrict__ tgt, OutT* __restrict__ x0,
    float* __restrict__ dynbuff_f32, float* __restrict__ dyndiff_f32, int N, long long units, long long per_units,
    AlphaTab al, int ring_start, const int* __restrict__ frame_index, const int* __restrict__ target_index, int dbg) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_red[64];
  __shared__ float s_mm[2 * 2 * RP_MAX_CLUSTER];
  __shared__ float s_lut[256];
  if (sizeof(InT) == 1) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) s_lut[b] = __fdiv_rn((float)b, 255.f);
    __syncthreads();
  }
  // smem: [per_units][C] float4 = the un-normalised dynamic image d.  The goal difference 0.5*(tgt - cur) is NOT staged:
  // pass 2 recomputes it from cur / tgt (two L2-resident re-reads), which halves the shared memory per CTA and lets
  // twice as many clusters (samples) run at once (ncu: 14 clusters of 16 x 98 KB were the occupancy limit)
  float4* sd0 = reinterpret_cast<float4*>(smem_raw);
  const long long n = blockIdx.y;
  const long long lo = (long long)blockIdx.x * per_units;
  long long hi = lo + per_units; if (hi > units) hi = units;
  const long long img4 = units * C;                       // float4 per image
  // frame k of this sample: dense [N,K,...] layout, or a pool of frames addressed through frame_index (windows of
  // one episode share frames: the pool holds each once)
  const InT* fk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int kp = (k + ring_start) % K;
    fk[k] = rgb + (frame_index ? (long long)frame_index[n * K + kp] : n * K + kp) * img4 * 4;
  }
  const InT* tbase = tgt + (target_index ? (long long)target_index[n] : n) * img4 * 4;
  const long long img_out = units * 4 * CP;               // OutT elements per padded image
  OutT* x_cur = x0 + n * img_out;
  OutT* x_dyn = x0 + ((long long)N + n) * img_out;
  OutT* x_dif = x0 + (2ll * N + n) * img_out;
  float mn0 = FLT_MAX, mx0 = -FLT_MAX, mn1 = FLT_MAX, mx1 = -FLT_MAX;
  // Pass 1 walks this CTA's slice as a FLAT array of float4 (d and the goal difference are element-wise): lane i of a
  // warp reads float4 i of every stream, 512 contiguous bytes per load instruction.  (The first version mapped a thread
  // to a 4-pixel unit = C float4: 48-byte lane stride, every 32-byte sector requested by two different instructions,
  // 12 partial lines per instruction -- ncu r02: 48 % of the samples on the load scoreboard at 31 % DRAM.)  The
  // channel-padded copy of the current frame needs whole pixels and is written in pass 2, which re-reads cur anyway.
  {
    const long long f_lo = lo * C, f_hi = hi * C;
    // float4 positions per thread in flight (six raw uint8 words per stream instead of two were measured slower:
    // 1.26 vs 0.97 ms per 1024-environment chunk)
    constexpr int U = K >= 6 ? 1 : 2;
    typedef typename FrameLoad<InT>::raw_t raw_t;
    for (long long f0 = f_lo + threadIdx.x; f0 < f_hi; f0 += (long long)U * RP_THREADS) {
      raw_t v[U][K], t[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const long long f = f0 + (long long)q * RP_THREADS;
        if (f < f_hi) {
#pragma unroll
          for (int k = 0; k < K; ++k) v[q][k] = FrameLoad<InT>::raw(fk[k], f);
          t[q] = FrameLoad<InT>::raw(tbase, f);
        }
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const long long f = f0 + (long long)q * RP_THREADS;
        if (f < f_hi) {
          float4 acc = f4_scale(al.a[0], FrameLoad<InT>::cvt(v[q][0]));
          float4 last = FrameLoad<InT>::cvt(v[q][K - 1]);
#pragma unroll
          for (int k = 1; k < K; ++k) acc = f4_axpy(acc, al.a[k], k == K - 1 ? last : FrameLoad<InT>::cvt(v[q][k]));
          const float4 dd = f4_axpy(f4_scale(-0.5f, last), 0.5f, FrameLoad<InT>::cvt(t[q]));
          sd0[f - f_lo] = acc;
          f4_minmax(acc, mn0, mx0);
          f4_minmax(dd, mn1, mx1);
        }
      }
    }
  }
  cluster_minmax(mn0, mx0, s_red, s_mm, 0, 2);
  cluster_minmax(mn1, mx1, s_red, s_mm, 1, 2);
  cluster_minmax_read(mn0, mx0, s_mm, 0);
  cluster_minmax_read(mn1, mx1, s_mm, 1);
  const float rng0 = __fadd_rn(__fsub_rn(mx0, mn0), 1e-6f), rng1 = __fadd_rn(__fsub_rn(mx1, mn1), 1e-6f);
  float4* o0 = dynbuff_f32 ? reinterpret_cast<float4*>(dynbuff_f32) + n * img4 : nullptr;
  float4* o1 = dyndiff_f32 ? reinterpret_cast<float4*>(dyndiff_f32) + n * img4 : nullptr;
  if (dbg & 1) return;                                     // timing experiment: first pass only
  for (long long u = lo + threadIdx.x; u < hi; u += RP_THREADS) {
    float e0[4 * C], e1[4 * C], ec[4 * C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float4 a = f4_norm(sd0[(u - lo) * C + j], mn0, rng0);
      // same two operations on the same operands as in pass 1: bit-identical difference image
      const float4 c4 = (dbg & 2) ? FrameLoad<InT>::at(fk[K - 1], u * C + j, s_lut) : FrameLoad<InT>::at_l1(fk[K - 1], u * C + j, s_lut);
      const float4 tj = (dbg & 2) ? FrameLoad<InT>::at(tbase, u * C + j, s_lut) : FrameLoad<InT>::at_l1(tbase, u * C + j, s_lut);
      const float4 b = f4_norm(f4_axpy(f4_scale(-0.5f, c4), 0.5f, tj), mn1, rng1);
      e0[j * 4] = a.x; e0[j * 4 + 1] = a.y; e0[j * 4 + 2] = a.z; e0[j * 4 + 3] = a.w;
      e1[j * 4] = b.x; e1[j * 4 + 1] = b.y; e1[j * 4 + 2] = b.z; e1[j * 4 + 3] = b.w;
      ec[j * 4] = c4.x; ec[j * 4 + 1] = c4.y; ec[j * 4 + 2] = c4.z; ec[j * 4 + 3] = c4.w;
      if (o0) stg_stream(o0 + u * C + j, a);
      if (o1) stg_stream(o1 + u * C + j, b);
    }
    store_unit<OutT, CP, C>(x_cur + u * 4 * CP, ec);
    store_unit<OutT, CP, C>(x_dyn + u * 4 * CP, e0);
    store_unit<OutT, CP, C>(x_dif + u * 4 * CP, e1);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int pick_cluster(long long bytes_per_sample, int hint) {
  if (hint > 0) return hint;
  int cl = 1;
  while (cl < RP_MAX_CLUSTER && bytes_per_sample / cl > 100 * 1024) cl *= 2;
  return cl;
}

template <typename KernelT>
static int launch_clustered(KernelT kernel, dim3 grid, int cl, size_t smem, cudaStream_t st, void** args) {
  CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cl > 8) CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(RP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = geeco_pdl_enabled() ? 2 : 1;
  CUDA_TRY(cudaLaunchKernelExC(&cfg, (const void*)kernel, args));
  geeco_count_launch(1);
  return GEECO_OK;
}

#define RP_SWITCH_K(K, STMT)                                                                     \
  switch (K) {                                                                                   \
    case 2: { constexpr int KK = 2; STMT; } break;   case 3: { constexpr int KK = 3; STMT; } break;   \
    case 4: { constexpr int KK = 4; STMT; } break;   case 5: { constexpr int KK = 5; STMT; } break;   \
    case 6: { constexpr int KK = 6; STMT; } break;   case 7: { constexpr int KK = 7; STMT; } break;   \
    case 8: { constexpr int KK = 8; STMT; } break;   case 9: { constexpr int KK = 9; STMT; } break;   \
    case 10: { constexpr int KK = 10; STMT; } break; case 11: { constexpr int KK = 11; STMT; } break; \
    case 12: { constexpr int KK = 12; STMT; } break; case 13: { constexpr int KK = 13; STMT; } break; \
    case 14: { constexpr int KK = 14; STMT; } break; case 15: { constexpr int KK = 15; STMT; } break; \
    case 16: { constexpr int KK = 16; STMT; } break;                                              \
    default: geeco_set_error("dynimg: window size K=%d outside [2,16]", K); return GEECO_ERR_INVALID; \
  }

#define RP_SWITCH_K_LO(K, STMT)                                                                  \
  switch (K) {                                                                                   \
    case 2: { constexpr int KK = 2; STMT; } break;   case 3: { constexpr int KK = 3; STMT; } break;   \
    case 4: { constexpr int KK = 4; STMT; } break;   case 5: { constexpr int KK = 5; STMT; } break;   \
    case 6: { constexpr int KK = 6; STMT; } break;   case 7: { constexpr int KK = 7; STMT; } break;   \
    case 8: { constexpr int KK = 8; STMT; } break;                                                \
    default: geeco_set_error("preprocess: window size K=%d outside [2,8]", K); return GEECO_ERR_INVALID; \
  }

static const size_t RP_SMEM_CAP = 200 * 1024;

int launch_dynimg(const float* in, float* out, int N, int K, long long HWC, const float* alpha_host,
                  int cluster_hint, cudaStream_t st) {
  if (HWC % 4) { geeco_set_error("dynimg: H*W*C=%lld must be a multiple of 4", HWC); return GEECO_ERR_INVALID; }
  if (N <= 0) return GEECO_OK;
  if (N > 65535) { geeco_set_error("dynimg: N=%d > 65535", N); return GEECO_ERR_INVALID; }
  AlphaTab al = {};
  for (int k = 0; k < K && k < 16; ++k) al.a[k] = alpha_host[k];
  long long total4 = HWC / 4;
  int cl = pick_cluster(HWC * 4, cluster_hint);
  long long per4 = (total4 + cl - 1) / cl;
  size_t smem = (size_t)per4 * 16;
  if (smem > RP_SMEM_CAP) return GEECO_ERR_WORKSPACE;   // caller falls back to the two-pass path
  void* args[] = {(void*)&in, (void*)&out, (void*)&total4, (void*)&per4, (void*)&al};
  RP_SWITCH_K(K, return launch_clustered(dynimg_cluster_kernel<KK>, dim3(cl, N), cl, smem, st, args));
  return GEECO_OK;
}

int launch_dynimg_twopass(const float* in, float* out, float* minmax_scratch, int N, int K, long long HWC,
                          const float* alpha_host, cudaStream_t st) {
  if (HWC % 4) { geeco_set_error("dynimg: H*W*C=%lld must be a multiple of 4", HWC); return GEECO_ERR_INVALID; }
  if (N <= 0) return GEECO_OK;
  AlphaTab al = {};
  for (int k = 0; k < K && k < 16; ++k) al.a[k] = alpha_host[k];
  long long total4 = HWC / 4;
  int* mm = reinterpret_cast<int*>(minmax_scratch);
  GEECO_LAUNCH((minmax_init_kernel), ceil_div(N, 256), 256, 0, st, mm, N);
  int bps = (int)((total4 + 256 * 8 - 1) / (256 * 8)); if (bps < 1) bps = 1; if (bps > 1024) bps = 1024;
  dim3 grid(bps, N);
  RP_SWITCH_K(K, (GEECO_LAUNCH((dynimg_pass1_kernel<KK>), grid, 256, 0, st, in, out, mm, total4, al)));
  GEECO_LAUNCH((dynimg_pass2_kernel), grid, 256, 0, st, out, mm, total4);
  geeco_count_launch(3);
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

template <typename InT, typename OutT, int CP, int C>
static int launch_pre_t(const InT* rgb, const InT* tgt, void* x0, float* db, float* dd, int N, int K, int H, int W,
                        const AlphaTab& al_in, int cluster_hint, int ring_start, const int* frame_index,
                        const int* target_index, cudaStream_t st) {
  long long units = (long long)H * W / 4;
  if (cluster_hint <= 0) { if (const char* e = getenv("GEECO_PRE_CLUSTER")) cluster_hint = atoi(e); }
  int cl = pick_cluster((long long)H * W * C * 4, cluster_hint);
  // two co-resident CTAs of a 98 KB slice each (the kernel is compiled for 64 registers): float32 frames 84 us at
  // batch 64 against 92-94 us for one CTA per SM with a slice twice as large (GEECO_PRE_CLUSTER=4), uint8 frames 65 us
  // against 104 us when 80 registers kept the second CTA out
  long long per_units = (units + cl - 1) / cl;
  size_t smem = (size_t)per_units * C * 16;
  if (smem > RP_SMEM_CAP) {
    geeco_set_error("preprocess: %dx%dx%d needs %zu B of shared memory per CTA at cluster %d", H, W, C, smem, cl);
    return GEECO_ERR_INVALID;
  }
  OutT* x = reinterpret_cast<OutT*>(x0);
  AlphaTab al = al_in;
  static const int dbg_env = getenv("GEECO_PRE_DBG") ? atoi(getenv("GEECO_PRE_DBG")) : 0;
  int dbg = dbg_env;
  void* args[] = {(void*)&rgb, (void*)&tgt, (void*)&x, (void*)&db, (void*)&dd, (void*)&N, (void*)&units,
                  (void*)&per_units, (void*)&al, (void*)&ring_start, (void*)&frame_index, (void*)&target_index, (void*)&dbg};
  RP_SWITCH_K_LO(K, return launch_clustered(preprocess_geecof_kernel<InT, OutT, CP, C, KK>, dim3(cl, N), cl, smem, st, args));
  return GEECO_OK;
}

int launch_preprocess_geecof(const void* rgb, const void* tgt, int frames_u8, void* x0, int out_bf16, int CP,
                             float* dynbuff_f32, float* dyndiff_f32, int N, int K, int H, int W, int C,
                             const float* alpha_host, int cluster_hint, int ring_start, const int* frame_index,
                             const int* target_index, cudaStream_t st) {
  if ((H * (long long)W) % 4) { geeco_set_error("preprocess: H*W must be a multiple of 4"); return GEECO_ERR_INVALID; }
  if (ring_start < 0 || ring_start >= K) { geeco_set_error("preprocess: ring_start %d outside [0,%d)", ring_start, K); return GEECO_ERR_INVALID; }
  if (N <= 0) return GEECO_OK;
  if (N > 65535) { geeco_set_error("preprocess: N=%d > 65535", N); return GEECO_ERR_INVALID; }
  if (K > 8) { geeco_set_error("preprocess: fused path supports window_size <= 8 (got %d)", K); return GEECO_ERR_INVALID; }
  AlphaTab al = {};
  for (int k = 0; k < K; ++k) al.a[k] = alpha_host[k];
#define PRE_CASE(T, cp, c)                                                                                        \
  return frames_u8 ? launch_pre_t<unsigned char, T, cp, c>((const unsigned char*)rgb, (const unsigned char*)tgt, x0,  \
                                                           dynbuff_f32, dyndiff_f32, N, K, H, W, al, cluster_hint, ring_start, frame_index, \
                                                           target_index, st) \
                   : launch_pre_t<float, T, cp, c>((const float*)rgb, (const float*)tgt, x0, dynbuff_f32,             \
                                                   dyndiff_f32, N, K, H, W, al, cluster_hint, ring_start, frame_index, target_index, st)
  if (!out_bf16 && CP == 4 && C == 3) PRE_CASE(float, 4, 3);
  if (!out_bf16 && CP == 4 && C == 4) PRE_CASE(float, 4, 4);
  if (out_bf16 && CP == 8 && C == 3) PRE_CASE(__nv_bfloat16, 8, 3);
  if (out_bf16 && CP == 8 && C == 4) PRE_CASE(__nv_bfloat16, 8, 4);
  if (out_bf16 && CP == 4 && C == 3) PRE_CASE(__nv_bfloat16, 4, 3);
  if (out_bf16 && CP == 4 && C == 4) PRE_CASE(__nv_bfloat16, 4, 4);
#undef PRE_CASE
  geeco_set_error("preprocess: unsupported (bf16=%d, CP=%d, C=%d)", out_bf16, CP, C);
  return GEECO_ERR_INVALID;
}


// ---------------------------------------------------------------------------------------
// pre-process of the graphs that feed every frame through the encoder (`--proc_obs sequence`, graph.py:360-385, and the
// unconditional e2e_vmc, :304-309).  Group 0 of x0 = the K frames (+ the target frame for --proc_tgt constant / residual,
// :352-355), step-major: image t*N + n = frame t of sample n, channel-padded, OutT.  With --proc_tgt dyndiff group 1 =
// dynimg([frame_t, target]) (:371-376): two passes (per-image min / max through ordered-int atomics, then the
// normalised write); both evaluate 0.5*tgt - 0.5*frame with the same two roundings, so they see the same values.
// ---------------------------------------------------------------------------------------
template <typename InT, typename OutT, int CP, int C>
__global__ void __launch_bounds__(256) seq_frames_kernel(const InT* __restrict__ rgb, const InT* __restrict__ tgt,
                                                         OutT* __restrict__ x0, int* __restrict__ mm, int N, int K,
                                                         int with_diff, long long units, int ring_start,
                                                         const int* __restrict__ frame_index,
                                                         const int* __restrict__ target_index) {
  pdl_enter();
  __shared__ float s_lut[256];
  if (sizeof(InT) == 1) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) s_lut[b] = __fdiv_rn((float)b, 255.f);
    __syncthreads();
  }
  const int img = blockIdx.y, t = img / N, n = img - t * N;
  const long long img4 = units * C;
  const long long fsel = (long long)n * K + (t + ring_start) % K;
  const long long tsel = target_index ? (long long)target_index[n] : n;
  const InT* fbase = t < K ? rgb + (frame_index ? (long long)frame_index[fsel] : fsel) * img4 * 4 : tgt + tsel * img4 * 4;
  const InT* tbase = tgt ? tgt + tsel * img4 * 4 : nullptr;
  OutT* xo = x0 + (long long)img * units * 4 * CP;
  const bool diff = with_diff && t < K;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
    float e[4 * C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float4 v = FrameLoad<InT>::at(fbase, u * C + j, s_lut);
      e[j * 4] = v.x; e[j * 4 + 1] = v.y; e[j * 4 + 2] = v.z; e[j * 4 + 3] = v.w;
      if (diff) f4_minmax(f4_axpy(f4_scale(-0.5f, v), 0.5f, FrameLoad<InT>::at(tbase, u * C + j, s_lut)), mn, mx);
    }
    store_unit<OutT, CP, C>(xo + u * 4 * CP, e);
  }
  if (diff) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm + 2 * img, f2ord(mn)); atomicMax(mm + 2 * img + 1, f2ord(mx)); }
  }
}

template <typename InT, typename OutT, int CP, int C>
__global__ void __launch_bounds__(256) seq_dyndiff_kernel(const InT* __restrict__ rgb, const InT* __restrict__ tgt,
                                                          OutT* __restrict__ x1, const int* __restrict__ mm,
                                                          float* __restrict__ dyndiff_f32, int N, int K, long long units,
                                                          int ring_start, const int* __restrict__ frame_index,
                                                          const int* __restrict__ target_index) {
  pdl_enter();
  __shared__ float s_lut[256];
  if (sizeof(InT) == 1) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) s_lut[b] = __fdiv_rn((float)b, 255.f);
    __syncthreads();
  }
  const int img = blockIdx.y, t = img / N, n = img - t * N;
  const long long img4 = units * C;
  const long long fsel = (long long)n * K + (t + ring_start) % K;
  const InT* fbase = rgb + (frame_index ? (long long)frame_index[fsel] : fsel) * img4 * 4;
  const InT* tbase = tgt + (target_index ? (long long)target_index[n] : n) * img4 * 4;
  OutT* xo = x1 + (long long)img * units * 4 * CP;
  const float mn = ord2f(mm[2 * img]), mx = ord2f(mm[2 * img + 1]);
  const float rng = __fadd_rn(__fsub_rn(mx, mn), 1e-6f);
  float4* o = (dyndiff_f32 && t == K - 1) ? reinterpret_cast<float4*>(dyndiff_f32) + (long long)n * img4 : nullptr;   // endpoints['dyndiff'] keeps the last frame's image
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
    float e[4 * C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float4 v = FrameLoad<InT>::at(fbase, u * C + j, s_lut);
      const float4 b = f4_norm(f4_axpy(f4_scale(-0.5f, v), 0.5f, FrameLoad<InT>::at(tbase, u * C + j, s_lut)), mn, rng);
      e[j * 4] = b.x; e[j * 4 + 1] = b.y; e[j * 4 + 2] = b.z; e[j * 4 + 3] = b.w;
      if (o) stg_stream(o + u * C + j, b);
    }
    store_unit<OutT, CP, C>(xo + u * 4 * CP, e);
  }
}

template <typename InT, typename OutT, int CP, int C>
static int launch_seq_t(const InT* rgb, const InT* tgt, void* x0, int* mm, float* dd, int N, int K, int H, int W,
                        int with_tgt, int with_diff, int ring_start, const int* frame_index, const int* target_index,
                        cudaStream_t st) {
  const long long units = (long long)H * W / 4;
  const int imgs = K * N + (with_tgt ? N : 0);
  int bx = (int)((units + 256 * 4 - 1) / (256 * 4)); if (bx < 1) bx = 1;
  OutT* x = reinterpret_cast<OutT*>(x0);
  if (with_diff) GEECO_LAUNCH((minmax_init_kernel), ceil_div(K * N, 256), 256, 0, st, mm, K * N);
  GEECO_LAUNCH((seq_frames_kernel<InT, OutT, CP, C>), dim3(bx, imgs), 256, 0, st, rgb, tgt, x, mm, N, K, with_diff, units, ring_start, frame_index, target_index);
  geeco_count_launch(with_diff ? 2 : 1);
  if (with_diff) {
    GEECO_LAUNCH((seq_dyndiff_kernel<InT, OutT, CP, C>), dim3(bx, K * N), 256, 0, st, rgb, tgt, x + (long long)K * N * units * 4 * CP, mm, dd,
                                                                         N, K, units, ring_start, frame_index, target_index);
    geeco_count_launch(1);
  }
  CUDA_TRY(cudaGetLastError());
  return GEECO_OK;
}

int launch_preprocess_seq(const void* rgb, const void* tgt, int frames_u8, void* x0, int out_bf16, int CP, int* minmax_scratch,
                          float* dyndiff_f32, int N, int K, int H, int W, int C, int with_tgt, int with_diff,
                          int ring_start, const int* frame_index, const int* target_index, cudaStream_t st) {
  if ((H * (long long)W) % 4) { geeco_set_error("preprocess: H*W must be a multiple of 4"); return GEECO_ERR_INVALID; }
  if (N <= 0) return GEECO_OK;
  if ((long long)(K + 1) * N > 65535) { geeco_set_error("preprocess: (K+1)*N = %lld images > 65535", (long long)(K + 1) * N); return GEECO_ERR_INVALID; }
  if (ring_start < 0 || ring_start >= K) { geeco_set_error("preprocess: ring_start %d outside [0,%d)", ring_start, K); return GEECO_ERR_INVALID; }
  if ((with_tgt || with_diff) && !tgt) { geeco_set_error("preprocess: target frame needed"); return GEECO_ERR_INVALID; }
  if (with_diff && !minmax_scratch) { geeco_set_error("preprocess: min/max scratch missing"); return GEECO_ERR_INVALID; }
#define SEQ_CASE(T, cp, c)                                                                                          \
  return frames_u8 ? launch_seq_t<unsigned char, T, cp, c>((const unsigned char*)rgb, (const unsigned char*)tgt, x0,    \
                                                           minmax_scratch, dyndiff_f32, N, K, H, W, with_tgt, with_diff, \
                                                           ring_start, frame_index, target_index, st)                   \
                   : launch_seq_t<float, T, cp, c>((const float*)rgb, (const float*)tgt, x0, minmax_scratch, dyndiff_f32, \
                                                   N, K, H, W, with_tgt, with_diff, ring_start, frame_index, target_index, st)
  if (!out_bf16 && CP == 4 && C == 3) SEQ_CASE(float, 4, 3);
  if (!out_bf16 && CP == 4 && C == 4) SEQ_CASE(float, 4, 4);
  if (out_bf16 && CP == 4 && C == 3) SEQ_CASE(__nv_bfloat16, 4, 3);
  if (out_bf16 && CP == 4 && C == 4) SEQ_CASE(__nv_bfloat16, 4, 4);
#undef SEQ_CASE
  geeco_set_error("preprocess: unsupported (bf16=%d, CP=%d, C=%d)", out_bf16, CP, C);
  return GEECO_ERR_INVALID;
}
