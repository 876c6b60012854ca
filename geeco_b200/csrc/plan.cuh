// Context layout shared by geeco_api.cu and the bf16 step (step_bf16.cu).
#pragma once
#include "common.cuh"
#include "../../include/geeco_b200.h"
#include <vector>

struct LayerPlan {
  int Hin, Hout, stride;
  int Cin_real, Cin_pad;
  int Cout[3];
  bool grouped;               // all encoder groups share shapes -> one grouped launch
  int p_w[3], p_b[3];         // parameter-table indices
  long long act_off[3];       // element offset of encoder e inside y / g
  long long act_elems;
  void* y;                    // post-ReLU output  [G][M][Hout][Hout][Cout]
  void* g;                    // dL/d(pre-activation), same layout (training only)
  void* mbits;                // bf16 training, conv1-conv7: 1-bit ReLU mask of y, uint16 per (pixel, 16 channels)
  // bf16 mode: packed weight copies (see step_bf16.cu)
  void* w_fwd;                // [3][Cout][Kpad]  bf16, K-major
  void* w_dgrad;              // [3][9][Cin][Cout] bf16 (tap-major, Cout contiguous)
  int Kpad;
};

struct HeadPlan { const char* name; int width, kind, slot, aux; int p_w, p_b; };

struct geeco_ctx {
  geeco_config cfg;
  std::vector<geeco_param_desc> params;
  long long arena_floats = 0;
  long long bucket_end[4] = {0, 0, 0, 0};
  size_t workspace_bytes = 0;
  bool bound = false, weights_dirty = true, fwd_done = false, uniform8 = true;
  // graph variant (tail.cuh: VAR_*): G encoder weight sets ("groups") of M images each, T LSTM steps
  int variant = 0, G = 3, M = 0, T = 1;
  int dim8[3] = {0, 0, 0};          // conv8 width per group
  const char* enc_scope[3] = {nullptr, nullptr, nullptr};
  const char* dec_scope = nullptr;
  int nheads = 0;
  HeadPlan heads_plan[5];
  int CP = 4, NH = 12, xdim = 0;
  float alpha[16];
  float host_sc[8];
  LayerPlan layers[8];
  int p_lstm_w, p_lstm_b, p_fc1_w, p_fc1_b;
  float *theta = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr;
  void* x0 = nullptr;
  // tail buffers: states [T][N][xdim+Hl], gates [T][N][4Hl], c / m [T][N][Hl]
  float *states, *gates, *c_seq, *m_seq, *state_out, *c_carry, *m_carry, *fc1, *heads, *loss_parts, *dheads, *losses, *sc;
  float *y8_f32 = nullptr, *g8_f32 = nullptr, *gates_partial = nullptr;
  float *dfc1 = nullptr, *dgates = nullptr, *dstates = nullptr, *dm_last = nullptr, *dc = nullptr, *partial = nullptr;
  int* mm_scratch = nullptr;
  long long partial_cap = 0;
  const unsigned char* reset_mask = nullptr;   // of the batch of the last forward (carry_state)
  int ring_start = 0;
  long long host_step = 0;                     // Adam updates applied so far (global_step of the reference's checkpoints)
  bool y1_stale = false;                       // bf16 training: the last forward did not store y1 (conv2's weight gradient
                                               // recomputes it on chip); geeco_debug_buffer("y1") rebuilds it on demand
  bool g8_bf16_ready = false;                  // bf16 mode: the tail already wrote layers[7].g as bf16
  // bf16 extras
  void* bf16_ws = nullptr;
};

// group stride helpers: distance (floats) between the kernels / biases of consecutive encoder groups of layer L
static inline long long w_group_stride(const geeco_ctx* c, const LayerPlan& L) {
  return c->G > 1 ? c->params[L.p_w[1]].offset - c->params[L.p_w[0]].offset : 0;
}
static inline long long b_group_stride(const geeco_ctx* c, const LayerPlan& L) {
  return c->G > 1 ? c->params[L.p_b[1]].offset - c->params[L.p_b[0]].offset : 0;
}

GatherGeom conv_fwd_geom(int H, int W, int Cs, int Cw, int Cout, int stride, int imgs_per_group);
bool conv_dgrad_geom(int H, int W, int Cin, int Cout, int stride, int py, int px, int imgs_per_group, GatherGeom* out);
GatherGeom dense_geom(int rows, int K, int Nn, int ldb, int transB);
long long gemm_tn_partial_floats(const GatherGeom& g, int groups);
void geeco_count_launch(int n);

// step_bf16.cu
int plan_bf16(geeco_ctx* c, size_t* ws_off, char* ws_base);
void free_bf16(geeco_ctx* c);
int repack_fork_bf16(geeco_ctx* c, cudaStream_t st);     // weight repack on a side stream, ordered after `st` so far
int repack_join_bf16(geeco_ctx* c, cudaStream_t st);     // `st` waits for it
int encoders_fwd_bf16(geeco_ctx* c, cudaStream_t st);
int profile_kernel_bf16(geeco_ctx* c, const char* name, cudaStream_t st);
int recompute_y1_bf16(geeco_ctx* c, cudaStream_t st);      // debug: conv1's activation of the last forward into layers[0].y
int encoders_bwd_bf16(geeco_ctx* c, int lhi, int llo, cudaStream_t st);
