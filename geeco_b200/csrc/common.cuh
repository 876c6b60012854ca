// Shared declarations for the geeco_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define GEECO_MAX_TAPS 9

// Status codes returned across the C-ABI (include/geeco_b200.h).
enum {
  GEECO_OK = 0,
  GEECO_ERR_INVALID = 1,      // bad argument / unsupported configuration  -> ValueError
  GEECO_ERR_CUDA = 2,         // CUDA runtime error                         -> RuntimeError
  GEECO_ERR_WORKSPACE = 3,    // workspace too small / not bound            -> RuntimeError
  GEECO_ERR_STATE = 4,        // call sequence error                        -> RuntimeError
};

// Geometry of one "gather GEMM": rows are pixels of a [imgs, Hm, Wm] grid, the reduction
// index k = (tap, c) addresses source pixel (y*sy + dy[tap], x*sx + dx[tap]) channel c.
// Covers: conv forward (stride 1/2, TF SAME padding), conv data-gradient (one launch per
// input-pixel parity class), dense layers (Hm = Wm = 1, one tap).
struct GatherGeom {
  int Hs, Ws, Cs;             // source tensor [imgs, Hs, Ws, Cs]
  int Hm, Wm;                 // GEMM row grid per image
  int sy, sx;                 // row (y, x) -> source (y*sy + dy, x*sx + dx)
  int ntaps;
  int dy[GEECO_MAX_TAPS], dx[GEECO_MAX_TAPS];
  int wbase[GEECO_MAX_TAPS];  // B element (tap, c, n): transB ? B[wbase + n*ldb + c] : B[(wbase + c)*ldb + n]
  int Cw;                     // channels per tap that exist in B (c >= Cw contributes zero)
  int ldb;
  int transB;
  int Nn;                     // GEMM N (output channels)
  int Hd, Wd;                 // destination tensor [imgs, Hd, Wd, Nn]
  int dsy, dsx, dy0, dx0;     // row (y, x) -> destination pixel (y*dsy + dy0, x*dsx + dx0)
  int imgs_per_group;         // images per weight group (encoder); grid.y = groups
  long long b_group_stride;   // elements between the B matrices of consecutive groups
  long long bias_group_stride;
};

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      geeco_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return GEECO_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)

void geeco_set_error(const char* fmt, ...);
void geeco_count_launch(int n);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// A step is ~50 dependent launches of 2-300 us each; between two of them the GPU drains, the next grid is launched and
// its CTAs are scheduled: 1-3 us per edge.  Every kernel of this library starts with pdl_enter(): it lets the NEXT
// kernel of the stream be launched as soon as all CTAs of this one are resident (griddepcontrol.launch_dependents) and
// then waits until the PREVIOUS kernel has completed and flushed its memory (griddepcontrol.wait) -- before the first
// global access, so the data dependencies are exactly those of a plain stream.  The host side launches with
// cudaLaunchAttributeProgrammaticStreamSerialization (GEECO_LAUNCH); GEECO_NO_PDL=1 turns the attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
#ifdef GEECO_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;");
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool geeco_pdl_enabled();
void geeco_pdl_suspend(int on);      // launches made while suspended are plain stream-ordered launches
template <typename... KArgs, typename... Args>
static inline void geeco_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = geeco_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface in the cudaGetLastError() that follows every launch
}
#define GEECO_LAUNCH(kernel, grid, block, smem, st, ...) \
  geeco_launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
#endif

// ---- launchers implemented in the .cu files --------------------------------------------
// rankpool.cu
int launch_dynimg(const float* in, float* out, int N, int K, long long HWC, const float* alpha_host,
                  int cluster_hint, cudaStream_t st);
int launch_dynimg_twopass(const float* in, float* out, float* minmax_scratch, int N, int K, long long HWC,
                          const float* alpha_host, cudaStream_t st);
int launch_preprocess_geecof(const void* rgb, const void* tgt, int frames_u8, void* x0, int out_bf16, int CP,
                             float* dynbuff_f32, float* dyndiff_f32, int N, int K, int H, int W, int C,
                             const float* alpha_host, int cluster_hint, int ring_start, const int* frame_index,
                             const int* target_index, cudaStream_t st);
int launch_preprocess_seq(const void* rgb, const void* tgt, int frames_u8, void* x0, int out_bf16, int CP, int* minmax_scratch,
                          float* dyndiff_f32, int N, int K, int H, int W, int C, int with_tgt, int with_diff,
                          int ring_start, const int* frame_index, const int* target_index, cudaStream_t st);
// conv_fp32.cu
int launch_gemm_nn_f32(const GatherGeom& g, const float* src, const float* B, const float* bias,
                       const float* mask, float* dst, int groups, int epi, cudaStream_t st);
int launch_gemm_tn_f32(const GatherGeom& g, const float* src, const float* G, float* dW, float* dbias,
                       float* partial, long long partial_cap_floats, int groups,
                       long long dw_group_stride, long long dbias_group_stride, cudaStream_t st);
enum { EPI_STORE = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2, EPI_MASK = 3 };
