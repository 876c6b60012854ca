/* geeco_b200 -- C-ABI of the B200-native GEECO e2evmc hot path.
 *
 * The reference (ogroth/geeco) is pure Python on TensorFlow 1.15 and has NO native interface; the
 * boundary this library replaces is the set of TF graph functions and wrappers listed below.  Each
 * entry point cites the reference code it stands in for (paths relative to the reference root).
 * The Python host (geeco_b200/*.py) binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns an int status (GEECO_OK == 0); geeco_last_error() gives the message of
 *     the last failure on the calling thread.  No C++ exceptions cross this boundary.
 *   - all tensor pointers are CALLER-OWNED DEVICE memory (e.g. torch tensors' data_ptr()),
 *     contiguous, float32 unless stated; layouts are TensorFlow's: activations NHWC, conv kernels
 *     HWIO, dense kernels [in,out], LSTM kernel [in+h,4h] (gates i,j,f,o), LSTM state [c | m].
 *   - the library never allocates device memory: parameters / gradients / Adam moments live in four
 *     caller-provided flat float arenas, activations in one caller-provided workspace
 *     (sizes from geeco_query_sizes).
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - a context is not thread-safe (the reference drives its session from one Python thread).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GEECO_B200_H_
#define GEECO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEECO_OK 0
#define GEECO_ERR_INVALID 1   /* -> ValueError  (graph.py:250-252,357-359,382-384,408-410; estimator.py:173-175) */
#define GEECO_ERR_CUDA 2      /* -> RuntimeError */
#define GEECO_ERR_WORKSPACE 3 /* -> RuntimeError */
#define GEECO_ERR_STATE 4     /* -> RuntimeError */

#define GEECO_FP32 0          /* fp32 storage + fp32 CUDA-core math (parity mode, <= 1e-4 rel) */
#define GEECO_BF16 1          /* bf16 activations/weights on tcgen05 tensor cores, fp32 accumulate + master weights */

typedef struct geeco_ctx geeco_ctx;

/* Graph switches: the values of train_e2evmc.py's --goal_condition / --proc_obs / --proc_tgt / --control_mode
 * (scripts/train_e2evmc.py:34-75).  Zero everywhere = GEECO-F (goal_e2evmc, dynimg, dyndiff, cartesian). */
enum { GEECO_GOAL_TARGET = 0,   /* goal_e2evmc, graph.py:321-416 (scope GoalVMC) */
       GEECO_GOAL_NONE = 1 };   /* e2e_vmc, graph.py:268-319 (scope VMC); proc_obs / proc_tgt are ignored */
enum { GEECO_OBS_DYNIMG = 0,    /* graph.py:386-407: current frame + rank-pooled buffer + rank-pooled goal difference */
       GEECO_OBS_SEQUENCE = 1 };/* graph.py:360-385: every frame through the encoder, K LSTM steps */
enum { GEECO_TGT_DYNDIFF = 0,   /* graph.py:371-379 / :396-402 */
       GEECO_TGT_CONSTANT = 1,  /* graph.py:352-355,365-367: target frame through ConvEncoder, concatenated */
       GEECO_TGT_RESIDUAL = 2 };/* graph.py:368-370: tgt_feat - feat */
enum { GEECO_CTRL_CARTESIAN = 0,/* heads pred_cmd_ee, logits_cmd_grp (graph.py:233-239) */
       GEECO_CTRL_VELOCITY = 1 };/* heads pred_cmd_vel, pred_cmd_ee, pred_cmd_grp (graph.py:240-249), mse_loss (:430-450) */

/* Subset of E2EVMCConfig (src/models/e2evmc/params.py:7-28) that shapes the graph, plus the execution
 * switches of this library. */
typedef struct geeco_config {
  int32_t img_height, img_width, img_channels;
  int32_t dim_jnt_state, window_size;
  int32_t dim_s_obs, dim_s_dyn, dim_s_diff, dim_h_lstm, dim_h_fc, num_grp_states;
  int32_t batch_size;     /* rows per call on this device (lstm_memory shape, graph.py:212,218) */
  int32_t precision;      /* GEECO_FP32 | GEECO_BF16 */
  int32_t carry_state;    /* 0: zero LSTM state at every call, as the reference executes (dead assign,
                             graph.py:226); 1: carry [c|m] across calls (the intended semantics) */
  int32_t training;       /* 1: reserve backward + optimizer workspace */
  int32_t goal_condition; /* GEECO_GOAL_*  */
  int32_t proc_obs;       /* GEECO_OBS_*   */
  int32_t proc_tgt;       /* GEECO_TGT_*   */
  int32_t control_mode;   /* GEECO_CTRL_*  */
  int32_t dim_grp_command;/* width of pred_cmd_grp in velocity mode (params.py: dim_grp_command = 2) */
  float lr, lambda_aux, l2_regularizer;          /* params.py:24-27 */
  float adam_beta1, adam_beta2, adam_eps;        /* tf.train.AdamOptimizer defaults 0.9 / 0.999 / 1e-8 */
} geeco_config;

typedef struct geeco_sizes {
  int64_t arena_floats;     /* length of each of the theta / grad / m / v arenas */
  int64_t workspace_bytes;
  int32_t num_params;       /* number of named variables */
  int32_t num_buckets;      /* gradient buckets for overlapped all-reduce */
} geeco_sizes;

/* One trainable variable, named as TensorFlow names it (graph.py:73,214,342; predictor.py:87). */
typedef struct geeco_param_desc {
  char name[96];
  int64_t offset;           /* in floats, into each arena */
  int64_t numel;
  int32_t ndim;
  int32_t reserved0;
  int64_t shape[4];
} geeco_param_desc;

/* Input features/labels of one step; layout of `_prepare_v4` (src/data/geeco_gym.py:373-399). */
enum { GEECO_FRAMES_F32 = 0,   /* float32 in [0,1]: what model_fn receives (estimator.py:160-176) */
       GEECO_FRAMES_U8 = 1 }; /* uint8 [0..255] as recorded; divided by 255.0f on the device, bit-identical to the
                                 `parsed_example['rgb'] /= 255.0` of the input pipeline (geeco_gym.py:310) */

typedef struct geeco_batch {
  const void* rgb;          /* [N,K,H,W,C]               features['rgb'] (+depth as 4th channel for rgbd) */
  const void* target_rgb;   /* [N,H,W,C]                 features['target_rgb']; unused (may be NULL) with GEECO_GOAL_NONE */
  const float* jnt_state;   /* [N,K,J]                   features['jnt_state'] */
  const float* ee_state;    /* [N,K,7] or NULL           features['ee_state']   (losses only) */
  const float* obj_state;   /* [N,K,7] or NULL           features['obj_state']  (losses only) */
  const float* cmd;         /* [N,4]   or NULL           labels['cmd']          (losses, cartesian control) */
  const float* vel_target;  /* [N,J]   or NULL           labels['vel_target']   (losses, velocity control; estimator.py:230-236) */
  const float* ee_target;   /* [N,7]   or NULL           labels['ee_target']    (first three columns are used) */
  const float* grp_target;  /* [N,dim_grp_command] / NULL labels['grp_target'] */
  const uint8_t* reset_mask;/* [N] or NULL; with carry_state = 1 the rows whose byte is non-zero start from the zero LSTM
                               state (an environment that was reset); ignored with carry_state = 0 */
  const int32_t* frame_index; /* [N,K] or NULL.  Non-NULL: rgb is a POOL of frames [F,H,W,C] and frame k of row n is pool frame
                               frame_index[n*K + k].  Consecutive windows of an episode share K-1 of their K frames
                               (_window_v3, geeco_gym.py:615-631), so a batch of B windows is B+K-1 pool frames instead of
                               B*K: the layout the input pipeline emits with layout='pool' (4x fewer bytes to upload) */
  const int32_t* target_index;/* [N] or NULL.  Non-NULL: target_rgb is a pool [Ft,H,W,C] (one goal frame per episode,
                               geeco_gym.py:313,624-625) and row n uses pool frame target_index[n] */
  int32_t frame_format;     /* GEECO_FRAMES_F32 | GEECO_FRAMES_U8: element type of rgb AND target_rgb */
  int32_t ring_start;       /* rgb / jnt_state as ring buffers over the K axis: physical slot of the OLDEST frame; logical
                               frame k lives in slot (ring_start + k) % K.  0 = plain layout */
} geeco_batch;

/* Optional outputs (NULL = not wanted).  heads, cartesian control = [pred_cmd_ee 0:3 | logits_cmd_grp 3:3+G |
 * pred_aux_ee | pred_aux_obj] (graph.py:233-259); velocity control = [pred_cmd_vel 0:J | pred_cmd_ee | pred_cmd_grp
 * (dim_grp_command) | pred_aux_ee | pred_aux_obj] (graph.py:240-259): geeco_head_columns(). */
typedef struct geeco_outputs {
  float* heads;             /* [N, NH] */
  float* fc1;               /* [N, dim_h_fc]             endpoints['fc1'] */
  float* dynbuff;           /* [N,H,W,C]                 endpoints['dynbuff']  (graph.py:393; dynimg graphs only) */
  float* dyndiff;           /* [N,H,W,C]                 endpoints['dyndiff']  (graph.py:376,401; with sequence/dyndiff the
                                                         image of the LAST frame, as the endpoint dict keeps it) */
  float* lstm_state;        /* [N, 2*dim_h_lstm]         [c | m] after the step */
  float* losses;            /* [12]: loss_cmd_ee, loss_cmd_grp, loss_pos_ee, loss_pos_obj, loss_reg, loss,
                                    #correct gripper classes (cartesian), N, loss_cmd_vel (velocity), 0, 0, 0
                                    (estimator.py:218-239,246-254; graph.py:430-450) */
} geeco_outputs;
#define GEECO_NUM_LOSS_SLOTS 12

const char* geeco_last_error(void);
int geeco_version(void);

/* ---- context -------------------------------------------------------------------------------- */
int geeco_query_sizes(const geeco_config* cfg, geeco_sizes* out);
int geeco_create(const geeco_config* cfg, geeco_ctx** out);
int geeco_destroy(geeco_ctx* ctx);
/* grad/m/v may be NULL when cfg.training == 0 */
int geeco_bind(geeco_ctx* ctx, float* theta, float* grad, float* m, float* v, void* workspace,
               int64_t workspace_bytes);
int geeco_param_info(const geeco_ctx* ctx, int32_t index, geeco_param_desc* out);
/* number of columns of geeco_outputs.heads for this configuration */
int geeco_head_columns(const geeco_config* cfg);
/* gradient bucket `b` covers arena floats [offset, offset+numel); bucket b is complete after
 * geeco_step_backward(ctx, b, ...) returns (work enqueued on the stream). */
int geeco_grad_bucket(const geeco_ctx* ctx, int32_t bucket, int64_t* offset, int64_t* numel);
/* call after writing theta from the host side (refreshes derived bf16 weight copies) */
int geeco_params_changed(geeco_ctx* ctx, void* stream);
/* Adam step counter t (global_step of the reference's checkpoints) */
int geeco_set_step(geeco_ctx* ctx, int64_t t, void* stream);
/* LSTM carry state [N, 2*dim_h_lstm]; only meaningful with carry_state = 1 */
int geeco_set_lstm_state(geeco_ctx* ctx, const float* state_cm, void* stream);

/* ---- functional ops (parity-testable pieces of graph.py) ------------------------------------ */
/* dynimg(rgb_frames) graph.py:30-55.  in [N,K,H,W,C] -> out [N,H,W,C].  alpha (host, K floats) may be
 * NULL -> the fp32 coefficients of graph.py:17-28.  cluster = 0 picks the cluster size automatically;
 * -1 forces the two-pass path (scratch: 2*N floats of device memory, may be NULL otherwise). */
int geeco_dynimg(const float* in, float* out, int32_t N, int32_t K, int32_t H, int32_t W, int32_t C,
                 const float* alpha, int32_t cluster, float* scratch, void* stream);
int geeco_alpha_table(int32_t K, float* out_host);
/* tf.layers.conv2d(3x3, SAME, stride, bias, optional ReLU) graph.py:76-115.  x [N,H,W,Cin] with Cin % 4 == 0,
 * w HWIO [3,3,Cin,Cout], y [N,ceil(H/s),ceil(W/s),Cout].  fp32 path. */
int geeco_conv2d_same(const float* x, const float* w, const float* b, float* y, int32_t N, int32_t H,
                      int32_t W, int32_t Cin, int32_t Cout, int32_t stride, int32_t relu, void* stream);
/* gradients of the above given dy_pre = dL/d(pre-activation) [N,Ho,Wo,Cout]:
 * dw [3,3,Cin,Cout], db [Cout], dx [N,H,W,Cin] (NULL = skip; relu_mask_x: if non-NULL, dx is multiplied by
 * (relu_mask_x > 0) i.e. the result is the pre-activation gradient of the producing layer).
 * scratch: at least geeco_conv2d_bwd_scratch_floats(...) floats. */
int64_t geeco_conv2d_bwd_scratch_floats(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t stride);
int geeco_conv2d_same_bwd(const float* x, const float* w, const float* dy_pre, const float* relu_mask_x, float* dw,
                          float* db, float* dx, float* scratch, int64_t scratch_floats, int32_t N, int32_t H,
                          int32_t W, int32_t Cin, int32_t Cout, int32_t stride, void* stream);

/* bf16 tensor-core (tcgen05) versions of the two ops above.  x / y / dy_pre / relu_mask_x / dx are bf16 NHWC,
 * Cin % 8 == 0, Cout % 16 == 0, Cout <= 256; w, b, dw, db stay fp32 in TF layout with Cw <= Cin real input
 * channels (w is [3,3,Cw,Cout]; x channels >= Cw must be zero).  y_f32 (optional) receives an fp32 copy. */
int64_t geeco_conv2d_bf16_scratch_bytes(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t stride);
int geeco_conv2d_same_bf16(const void* x, const float* w, const float* b, void* y, float* y_f32, void* scratch,
                           int64_t scratch_bytes, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cw,
                           int32_t Cout, int32_t stride, int32_t relu, void* stream);
int geeco_conv2d_same_bwd_bf16(const void* x, const float* w, const void* dy_pre, const void* relu_mask_x, float* dw,
                               float* db, void* dx, void* scratch, int64_t scratch_bytes, int32_t N, int32_t H,
                               int32_t W, int32_t Cin, int32_t Cw, int32_t Cout, int32_t stride, void* stream);
/* The model step keeps the ReLU mask of an activation as ONE BIT per element (written by the forward epilogue): uint16
 * per (pixel, 16-channel chunk), bit j = channel 2j, bit 8+j = channel 2j+1 of the chunk.  geeco_relu_mask_bits builds
 * that mask from a bf16 activation [pixels][C] (C % 16 == 0); geeco_conv2d_same_bwd_bf16_bits is
 * geeco_conv2d_same_bwd_bf16 with the mask of x given in that form (same results bit for bit, 16x fewer mask bytes). */
int geeco_relu_mask_bits(const void* y, void* bits, int64_t pixels, int32_t C, void* stream);
int geeco_conv2d_same_bwd_bf16_bits(const void* x, const float* w, const void* dy_pre, const void* relu_mask_bits,
                                    float* dw, float* db, void* dx, void* scratch, int64_t scratch_bytes, int32_t N,
                                    int32_t H, int32_t W, int32_t Cin, int32_t Cw, int32_t Cout, int32_t stride,
                                    void* stream);

/* LSTMCell over K steps from the zero state, and its back-propagation through time: the recurrence of `lstm_decoder`
 * (src/models/e2evmc/graph.py:212-225) for the graphs that feed one feature vector per frame (`--proc_obs sequence`,
 * :360-385, and the unconditional e2e_vmc, :304-313).  fp32.  x [K][N][xdim] (feat_list, xdim % 4 == 0), kernel
 * [xdim+Hl][4*Hl] with gate order i, j, f, o and forget_bias 1, bias [4*Hl]; outputs gates [K][N][4*Hl] (kept for the
 * backward), c and m [K][N][Hl]; m[K-1] is `outputs[-1]`, [c[K-1] | m[K-1]] the final state.  The backward takes
 * dm_last = dL/d(outputs[-1]) [N][Hl] and returns d(kernel), d(bias) and (optional) dx [K][N][xdim].  Both need the
 * same 256-byte aligned scratch of geeco_lstm_seq_scratch_floats floats; deterministic (fixed-order split sums). */
int64_t geeco_lstm_seq_scratch_floats(int32_t N, int32_t K, int32_t xdim, int32_t Hl);
int geeco_lstm_seq_fwd(const float* x, const float* kernel, const float* bias, float* gates, float* c, float* m,
                       float* scratch, int64_t scratch_floats, int32_t N, int32_t K, int32_t xdim, int32_t Hl,
                       void* stream);
int geeco_lstm_seq_bwd(const float* x, const float* kernel, const float* gates, const float* c, const float* m,
                       const float* dm_last, float* dkernel, float* dbias, float* dx, float* scratch,
                       int64_t scratch_floats, int32_t N, int32_t K, int32_t xdim, int32_t Hl, void* stream);

/* ---- model step ------------------------------------------------------------------------------ */
/* goal_e2evmc forward (graph.py:321-416) [+ losses when batch->cmd and out->losses are given];
 * the predictor hook (predictor.py:148-190) and EVAL mode (estimator.py:246-258) use this. */
int geeco_forward(geeco_ctx* ctx, const geeco_batch* batch, const geeco_outputs* out, void* stream);
/* model_fn in TRAIN mode (estimator.py:144-244): forward, losses, backward, Adam.
 * grad_scale multiplies the gradients inside the optimizer (1/world_size after a summing all-reduce). */
int geeco_train_step(geeco_ctx* ctx, const geeco_batch* batch, const geeco_outputs* out, float grad_scale,
                     void* stream);
/* the same step in phases, so the host can all-reduce gradient buckets between them */
int geeco_step_forward(geeco_ctx* ctx, const geeco_batch* batch, const geeco_outputs* out, void* stream);
int geeco_step_backward(geeco_ctx* ctx, int32_t bucket, void* stream);
int geeco_step_update(geeco_ctx* ctx, float grad_scale, void* stream);
/* the optimizer step over gradient buckets [first, last] only (0 <= first <= last < num_buckets): lets the host update
 * the buckets whose all-reduce has finished while a later bucket's is still in flight.  Every bucket exactly once per
 * step; the call that includes the last bucket closes the step (same arithmetic as geeco_step_update: the reference's
 * single AdamOptimizer.minimize, estimator.py:243-244, is element-wise). */
int geeco_step_update_buckets(geeco_ctx* ctx, float grad_scale, int32_t first, int32_t last, void* stream);

/* ---- K-frame history of many environments as a device ring (BASELINE config 4) ------------------------------------
 * The FIFO of predictor.py:140-146 without shifting: writes frame[n] ([N][row_bytes]) into slot `slot` of
 * ring [N][K][row_bytes]; rows with fresh[n] != 0 get it in ALL K slots (a new episode's buffer is padded by repeating its
 * first frame, predictor.py:196-198).  row_bytes % 4 == 0; everything is device memory; no host synchronisation.  The step
 * then reads the ring through geeco_batch.ring_start = (slot + 1) % K. */
int geeco_ring_push(void* ring, const void* frame, const uint8_t* fresh, int32_t N, int32_t K, int64_t row_bytes,
                    int32_t slot, void* stream);

/* ---- introspection for tests / profiling ------------------------------------------------------ */
/* internal activation buffers by name ("x0", "y1".."y8", "g1".."g8", "state", "gates", "dstate", ...);
 * dtype: 0 = f32, 1 = bf16 */
int geeco_debug_buffer(const geeco_ctx* ctx, const char* name, void** ptr, int64_t* numel, int32_t* dtype);
/* runs ONE kernel of the bf16 step again on the buffers the last geeco_forward / geeco_train_step left behind, so that a
 * benchmark can time it alone with events on `stream`.  name: "conv12" = the fused conv1 -> conv2 forward
 * (graph.py:76-115, first two layers of conv_encoder). */
int geeco_profile_kernel(geeco_ctx* ctx, const char* name, void* stream);
/* number of kernel launches enqueued by this library since the last call with reset != 0 */
int64_t geeco_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* GEECO_B200_H_ */
