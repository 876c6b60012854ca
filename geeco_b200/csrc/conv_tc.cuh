// bf16 tensor-core (tcgen05) gather-GEMM kernels: declarations shared with step_bf16.cu / geeco_api.cu
#pragma once
#include <cuda.h>
#include <vector>
#include "common.cuh"

// Geometry of one tensor-core gather GEMM (bf16 NHWC source, implicit im2col).
struct TcGeom {
  int rowwin;                 // 1: 4-channel stride-1 3x3 forward geometry with Wm % 128 == 0 (conv1 fast producers)
  int hw_shift, w_shift, cs_shift;   // log2(Hm*Wm), log2(Wm), log2(Cs) when powers of two, else -1
  int Hs, Ws, Cs;             // source tensor [imgs, Hs, Ws, Cs] (bf16), Cs == 4 or Cs % 8 == 0
  int Hm, Wm;                 // GEMM-row pixel grid per image
  int sy, sx;
  int ntaps;
  int dy[GEECO_MAX_TAPS], dx[GEECO_MAX_TAPS];
  int Ktot, Kpad;             // ntaps*Kt and its round-up to 64
  int Kt;                     // K extent of one tap in the packed weight matrix: Cs, or Cs rounded up to 64 when
                              // the A operand is fetched by TMA (one 64-channel box per k-block, a_tma == 1)
  int a_tma;                  // 1: A tiles come from a 5-D tiled tensor map over the source (see tc_use_tma)
  int rows;                   // row-resident kernel (tc_rows_kernel): 0 no, 1 unit-stride source (data gradient of a
                              // stride-2 layer), 2 forward stride-2 layer on pixel pairs (Cs == 32)
  int wpack;                  // weight packing mode of the forward operand (0 tap-major, 4 pixel pairs; see pack_value)
  int bias_in_k;              // 1: the bias is folded into the GEMM (row-window conv1): packed-weight column Ktot holds the
                              // bias and the im2col rows carry a constant 1.0 there, so the epilogue neither loads nor adds it
  int Nn;                     // GEMM N of one tile (multiple of 16, <= 256)
  int nsplit, Ntot;           // small layers: the N = Ntot output channels of a group are cut into nsplit column tiles
                              // of Nn, each a "virtual group" (more CTAs); nsplit == 1, Ntot == Nn otherwise
  int Hd, Wd;                 // destination tensor [imgs, Hd, Wd, Nn]
  int dsy, dsx, dy0, dx0;
  int imgs_per_group, groups;
  int fast32;                 // set by the launchers: the lean N = 32 epilogue applies (see epilogue_n32 in conv_tc.cu)
  int b_rows_per_group;       // rows of the packed weight matrix per group (tensor-map row offset)
  long long bias_group_stride;
};

// per-class part of a geometry (taps, K extent, destination offset) and the weight maps of a multi-class launch
struct TcCls {
  int ntaps; int dy[GEECO_MAX_TAPS], dx[GEECO_MAX_TAPS]; int Ktot, Kpad, dy0, dx0;
  // TMA coordinates of a tap relative to the tile origin: channel base, column, row parity, row (see make_act_tensor_map)
  short tc0[GEECO_MAX_TAPS], twq[GEECO_MAX_TAPS], thp[GEECO_MAX_TAPS], thq[GEECO_MAX_TAPS];
};
struct TcClasses { int ncls; TcCls c[4]; };
// Program of a row-resident launch: per tile the TMA warp loads `nrows` source rows (one box of `pw` pixels x 64
// channels each) and the MMA warp runs, per class, `nsteps` shifted-window MMAs against resident weight k-blocks.
struct TcRowProg {
  int nrows, pw, pitch, w0;                       // pitch = pw * 128 bytes (pw % 8 == 0)
  short r_c0[4], r_hp[4], r_hq[4];                // TMA coordinates of a row: channel base, row parity, row offset
  int nsteps[4];
  int a_off[4][GEECO_MAX_TAPS];                   // byte offset of the 128-pixel window inside the stage
  short b_slot[4][GEECO_MAX_TAPS];                // resident weight k-block used by the step
  int b_slots;
  short slot_cls[4 * GEECO_MAX_TAPS], slot_kb[4 * GEECO_MAX_TAPS];   // where a slot comes from: class map, k-block
};
struct TcMaps { CUtensorMap m[4]; };

// TC_EPI_MASKBITS: like TC_EPI_MASK, but `mask` points at the 1-bit-per-element mask a TC_EPI_BIAS_RELU launch wrote
// through `bits_out` (uint16 per (pixel, 16-channel chunk); bit j = channel 2j, bit 8+j = channel 2j+1 of the chunk)
enum { TC_EPI_BIAS_RELU = 0, TC_EPI_MASK = 1, TC_EPI_STORE = 2, TC_EPI_BIAS = 3, TC_EPI_MASKBITS = 4,
       TC_EPI_RELU = 5 };   // TC_EPI_RELU: ReLU only (bias already inside the accumulator, see TcGeom::bias_in_k)

// fwd / dgrad: dst = epi(A(src) x Wp^T)
int launch_tc_nn(const TcGeom& g, const CUtensorMap* wmap, const __nv_bfloat16* src, const float* bias,
                 const __nv_bfloat16* mask, __nv_bfloat16* dst, float* dst_f32, int epi, int max_ctas,
                 cudaStream_t st, unsigned short* bits_out = nullptr, const __nv_bfloat16* wptr = nullptr);
//   wptrs (optional): base of the packed weight matrix behind each wmaps[c]; lets the launcher re-tile N for small layers
int launch_tc_nn_multi(const TcGeom* gs, const CUtensorMap* const* wmaps, int ncls, const __nv_bfloat16* src,
                       const float* bias, const __nv_bfloat16* mask, __nv_bfloat16* dst, float* dst_f32, int epi,
                       int max_ctas, cudaStream_t st, unsigned short* bits_out = nullptr,
                       const __nv_bfloat16* const* wptrs = nullptr);
// wgrad: dW[(tap,ci)][co] (+ optional bias gradient) from im2col(src)^T x G, deterministic split reduction.
//   g describes the FORWARD geometry (rows = output pixels); G is [imgs, Hm, Wm, Cout] bf16.
//   Cw = channels per tap present in the weight tensor (Cin_real); dW fp32 [ntaps*Cw, Cout] per group.
long long tc_wgrad_partial_floats(const TcGeom& g, int Cout);
int launch_tc_wgrad(const TcGeom& g, int Cout, int Cw, const __nv_bfloat16* src, const __nv_bfloat16* G, float* dW,
                    float* dbias, float* partial, long long partial_cap, long long dw_group_stride,
                    long long dbias_group_stride, cudaStream_t st);
// conv1 on pixel pairs (see conv_tc.cu): geometry of one output-column parity class, partial-only wgrad, its reduce
TcGeom tc_conv1pair_geom(int H, int W, int Cout, int imgs_per_group, int groups, int par);
int launch_tc_wgrad_partial(const TcGeom& g, int Cout, const __nv_bfloat16* src, const __nv_bfloat16* G, float* partial,
                            long long partial_cap, int want_ones, int* splits_out, int* mrows_out, cudaStream_t st);
int launch_conv1pair_reduce(const float* part_even, const float* part_odd, float* dW, float* dbias, int splits,
                            int groups, int Mrows_pad, int Kpad, int Cin, int Cout, long long dw_group_stride,
                            long long dbias_group_stride, cudaStream_t st);
// packed bf16 weights.  mode 0 (fwd): out[g][n][t*Cs + ch] = W[g][tap_t][ch][n] ; rows = Nn
//                       mode 1 (dgrad): out[g][ci][t*Cout + co] = W[g][tap_t][ci][co] ; rows = Cin
int launch_pack_weights(const float* W, __nv_bfloat16* out, int mode, int groups, long long w_group_stride, int Cin,
                        int Cout, int Cs, int ntaps, const int* taps, int rows, int Kpad, int Kt, cudaStream_t st,
                        const float* bias = nullptr, int bias_col = -1);
// all weight repacks of a step as ONE launch: a table of jobs in device memory
struct PackJob {
  const float* W; __nv_bfloat16* out; long long w_group_stride; long long start; long long total;
  int mode, groups, Cin, Cout, Cs, ntaps, rows, Kpad; int taps[9]; int Kt;   // Kt: K extent per tap (0 = Cs / Cout)
  const float* bias; long long b_group_stride; int bias_col; int pad_;       // bias_col >= 0: column that holds bias[row]
};
// jobs occupy [start, start + total) of a virtual index space; every start is a multiple of PACK_CHUNK and
// grand_total is the end of the last job rounded up to PACK_CHUNK (total < 2^31 per job)
constexpr int PACK_CHUNK = 2048;
int launch_pack_weights_batched(const PackJob* jobs_dev, int njobs, long long grand_total, cudaStream_t st, int max_blocks = 0);
// Tile form of the plain repack jobs (mode 0 / 1, no bias column): one 32 x 32 tile of one tap's [Cin][Cout] matrix per
// block, every index computed once per block.  Writes only real elements: the padding of the packed matrices must have
// been written once by launch_pack_weights_batched (it is never touched again).
//   transpose (mode 0): D[c * d_ld + r] = S[r * s_ld + c]      copy (mode 1): D[r * d_ld + c] = S[r * s_ld + c]
struct PackTile { const float* S; __nv_bfloat16* D; int s_ld, d_ld; short rows, cols, transpose, pad_; };
bool pack_job_is_plain(const PackJob& j);
void pack_job_tiles(const PackJob& j, std::vector<PackTile>* out);
int launch_pack_tiles(const PackTile* tiles_dev, int ntiles, cudaStream_t st);
// 1-bit ReLU mask (layout of TC_EPI_MASKBITS) of a bf16 tensor: one uint16 per 16 consecutive values
int launch_relu_mask_bits(const __nv_bfloat16* y, unsigned short* bits, long long chunks, cudaStream_t st);
int launch_f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st);
int launch_bf16_to_f32(const __nv_bfloat16* in, float* out, long long n, cudaStream_t st);
// 2-D tensor map over a packed weight matrix [rows_total][Kpad] bf16, box = 64 x box_rows, SWIZZLE_128B
int make_weight_tensor_map(CUtensorMap* map, const void* base, long long rows_total, int Kpad, int box_rows);

// 2-D bf16 tensor map [rows][inner] with SWIZZLE_128B (inner box of 64 elements = 128 bytes)
int make_tensor_map_2d_sw128(CUtensorMap* map, void* base, int inner, long long rows, long long row_stride_bytes, int box_inner,
                             int box_rows);
int tc_num_sms();             // SM count the persistent grids are sized for (honours GEECO_NUM_SMS)
// conv1 -> conv2 forward as one kernel (conv12_fused.cu): y1 stays on chip between the layers; y1 / bits1 / bits2 may be
// nullptr (inference writes neither y1 nor the ReLU masks)
bool tc_conv12_supported(int H, int W, int Cin_pad, int Cout1, int Cout2, int stride1, int stride2, const TcGeom& g1);
int launch_tc_conv12(const __nv_bfloat16* x0, const CUtensorMap* w1map, const CUtensorMap* w2map, const float* bias2,
                     long long bias2_group_stride, __nv_bfloat16* y1, unsigned short* bits1, __nv_bfloat16* y2,
                     unsigned short* bits2, int G, int M, cudaStream_t st, const CUtensorMap* w1pair_map = nullptr);

// conv2 data gradient -> conv1 weight gradient as one kernel (conv21_bwd_fused.cu): dL/d(pre-activation of conv1) stays
// on chip.  dg = the four parity classes of the stride-2 layer in (py, px) order, wmaps their packed-weight maps.
bool tc_bwd21_supported(int H, int W, int Cin_pad, int Cout1, int Cout2, int stride1, int stride2, const TcGeom* dg, int ncls);
long long tc_bwd21_partial_floats();
int launch_tc_bwd21(const __nv_bfloat16* G2, const TcGeom* dg, const CUtensorMap* const* wmaps, const unsigned short* bits1,
                    const __nv_bfloat16* x0, float* partial, long long partial_cap, float* dW1, float* dbias1, int Cw,
                    long long dw_group_stride, long long dbias_group_stride, int G, int M, cudaStream_t st);
// conv2 weight gradient on whole y1 rows (conv2_wgrad_fused.cu): y1 != nullptr loads the rows by TMA, y1 == nullptr
// recomputes them from x0 on chip (the forward then need not store y1)
bool tc_wgrad2_supported(int H, int W, int Cin_pad, int Cout1, int Cout2, int stride1, int stride2, const TcGeom& g1);
long long tc_wgrad2_partial_floats();
int launch_tc_wgrad2(const __nv_bfloat16* x0, const CUtensorMap* w1map, const __nv_bfloat16* y1, const __nv_bfloat16* G2,
                     float* partial, long long partial_cap, float* dW2, float* dbias2, long long dw_group_stride,
                     long long dbias_group_stride, int G, int M, cudaStream_t st);
// 5-D tensor map over the source of geometry g whose box is one row of `pw` pixels x 64 channels (row-resident kernels)
int tc_make_row_tensor_map(CUtensorMap* map, const void* base, const TcGeom& g, int pw);

TcGeom tc_fwd_geom(int H, int W, int Cs, int Cout, int stride, int imgs_per_group, int groups);
bool tc_dgrad_geom(int H, int W, int Cin, int Cout, int stride, int py, int px, int imgs_per_group, int groups,
                   TcGeom* out, int* taps_out);
