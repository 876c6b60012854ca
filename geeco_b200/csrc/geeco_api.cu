// C-ABI entry points, context / workspace planning and step orchestration.
// Public contract and reference citations: include/geeco_b200.h.
#include "common.cuh"
#include "tail.cuh"
#include "plan.cuh"
#include "../../include/geeco_b200.h"

#include <math.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------
// error string + launch counter
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static long long g_launches = 0;

void geeco_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void geeco_count_launch(int n) { g_launches += n; }
static int g_pdl_suspend = 0;
void geeco_pdl_suspend(int on) { g_pdl_suspend = on; }
bool geeco_pdl_enabled() {
  static const bool on = getenv("GEECO_NO_PDL") == nullptr;
  return on && !g_pdl_suspend;
}

extern "C" const char* geeco_last_error(void) { return g_err; }
extern "C" int geeco_version(void) { return 100; }
extern "C" int64_t geeco_launch_count(int32_t reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// fp32 emulation of graph.py:17-28 (every operand is a float32 tensor in the reference graph)
static float harmonic_f32(int t) {
  volatile float acc = 0.f;
  for (int i = 1; i <= t; ++i) acc = acc + 1.0f / (float)i;
  return acc;
}
extern "C" int geeco_alpha_table(int32_t K, float* out) {
  if (K < 1 || !out) { geeco_set_error("alpha_table: bad arguments"); return GEECO_ERR_INVALID; }
  const float T = (float)K, HT = harmonic_f32(K);
  for (int t = 1; t <= K; ++t) {
    volatile float lhs = 2.0f * ((T - (float)t) + 1.0f);
    volatile float diff = HT - harmonic_f32(t - 1);
    volatile float rhs = (T + 1.0f) * diff;
    out[t - 1] = lhs - rhs;
  }
  return GEECO_OK;
}

// ------------------------------------------------------------------------------------------
// geometry builders
// ------------------------------------------------------------------------------------------
static void same_pad(int in, int k, int s, int* out, int* before) {
  *out = (in + s - 1) / s;
  int total = (*out - 1) * s + k - in;
  if (total < 0) total = 0;
  *before = total / 2;
}

GatherGeom conv_fwd_geom(int H, int W, int Cs, int Cw, int Cout, int stride, int imgs_per_group) {
  GatherGeom g;
  memset(&g, 0, sizeof(g));
  int Ho, Wo, pt, pl;
  same_pad(H, 3, stride, &Ho, &pt);
  same_pad(W, 3, stride, &Wo, &pl);
  g.Hs = H; g.Ws = W; g.Cs = Cs; g.Hm = Ho; g.Wm = Wo; g.sy = stride; g.sx = stride;
  g.ntaps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int t = ky * 3 + kx;
      g.dy[t] = ky - pt; g.dx[t] = kx - pl; g.wbase[t] = t * Cw;
    }
  g.Cw = Cw; g.ldb = Cout; g.transB = 0; g.Nn = Cout;
  g.Hd = Ho; g.Wd = Wo; g.dsy = 1; g.dsx = 1; g.dy0 = 0; g.dx0 = 0;
  g.imgs_per_group = imgs_per_group;
  g.b_group_stride = 9ll * Cw * Cout;
  g.bias_group_stride = Cout;
  return g;
}

// data-gradient geometry of input-pixel class (py, px); returns false when the class has no pixels/taps
bool conv_dgrad_geom(int H, int W, int Cin, int Cout, int stride, int py, int px, int imgs_per_group, GatherGeom* out) {
  GatherGeom g;
  memset(&g, 0, sizeof(g));
  int Ho, Wo, pt, pl;
  same_pad(H, 3, stride, &Ho, &pt);
  same_pad(W, 3, stride, &Wo, &pl);
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
      // floor division (ny may be negative)
      g.dy[nt] = (ny >= 0) ? ny / stride : -((-ny + stride - 1) / stride);
      g.dx[nt] = (nx >= 0) ? nx / stride : -((-nx + stride - 1) / stride);
      g.wbase[nt] = (ky * 3 + kx) * Cin * Cout;
      ++nt;
    }
  }
  if (!nt) return false;
  g.ntaps = nt;
  g.Cw = Cout; g.ldb = Cout; g.transB = 1; g.Nn = Cin;
  g.Hd = H; g.Wd = W; g.dsy = stride; g.dsx = stride; g.dy0 = py; g.dx0 = px;
  g.imgs_per_group = imgs_per_group;
  g.b_group_stride = 9ll * Cin * Cout;
  g.bias_group_stride = 0;
  *out = g;
  return true;
}

GatherGeom dense_geom(int rows, int K, int Nn, int ldb, int transB) {
  GatherGeom g;
  memset(&g, 0, sizeof(g));
  g.Hs = 1; g.Ws = 1; g.Cs = K; g.Hm = 1; g.Wm = 1; g.sy = 1; g.sx = 1; g.ntaps = 1;
  g.Cw = K; g.ldb = ldb; g.transB = transB; g.Nn = Nn;
  g.Hd = 1; g.Wd = 1; g.dsy = 1; g.dsx = 1;
  g.imgs_per_group = rows;
  return g;
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
static const int kEncChannels[7] = {32, 48, 64, 128, 192, 256, 256};   // graph.py:76-110
static const int kEncStrides[8] = {1, 2, 2, 2, 2, 2, 2, 2};            // graph.py:78-113

struct Carver {
  char* base; size_t off;
  void* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

static int add_param(geeco_ctx* c, const std::string& name, std::initializer_list<int64_t> shape) {
  geeco_param_desc d;
  memset(&d, 0, sizeof(d));
  snprintf(d.name, sizeof(d.name), "%s", name.c_str());
  d.ndim = (int)shape.size();
  d.numel = 1;
  int i = 0;
  for (auto s : shape) { d.shape[i++] = s; d.numel *= s; }
  c->arena_floats = (c->arena_floats + 3) & ~3ll;
  d.offset = c->arena_floats;
  c->arena_floats += d.numel;
  c->params.push_back(d);
  return (int)c->params.size() - 1;
}

static int validate(const geeco_config* cfg) {
  if (!cfg) { geeco_set_error("config is NULL"); return GEECO_ERR_INVALID; }
  if (cfg->img_channels != 3 && cfg->img_channels != 4) {
    geeco_set_error("Unsupported number of channels for input frame: %d!", cfg->img_channels);   // estimator.py:173-175
    return GEECO_ERR_INVALID;
  }
  if (cfg->img_height != 256 || cfg->img_width != 256) {
    // graph.py:139,163,188 hard-code the 2x2 tiling of the joint state: 8 layers, 7 stride-2 -> 256 px only
    geeco_set_error("the e2evmc graph is only defined for 256x256 inputs (2x2 conv8 map); got %dx%d",
                    cfg->img_height, cfg->img_width);
    return GEECO_ERR_INVALID;
  }
  if (cfg->goal_condition != GEECO_GOAL_TARGET && cfg->goal_condition != GEECO_GOAL_NONE) { geeco_set_error("unknown goal_condition %d", cfg->goal_condition); return GEECO_ERR_INVALID; }
  if (cfg->proc_obs != GEECO_OBS_DYNIMG && cfg->proc_obs != GEECO_OBS_SEQUENCE) {
    geeco_set_error("Unknown processing mode for frame buffer: %d!", cfg->proc_obs);              // graph.py:408-410
    return GEECO_ERR_INVALID;
  }
  if (cfg->proc_tgt < GEECO_TGT_DYNDIFF || cfg->proc_tgt > GEECO_TGT_RESIDUAL) {
    geeco_set_error("Unknown processing mode for target image: %d!", cfg->proc_tgt);              // graph.py:357-359
    return GEECO_ERR_INVALID;
  }
  if (cfg->control_mode != GEECO_CTRL_CARTESIAN && cfg->control_mode != GEECO_CTRL_VELOCITY) {
    geeco_set_error("Unknown control mode '%d'", cfg->control_mode);                               // graph.py:250-252
    return GEECO_ERR_INVALID;
  }
  if (cfg->window_size < 2 || cfg->window_size > 8) { geeco_set_error("window_size %d outside [2,8]", cfg->window_size); return GEECO_ERR_INVALID; }
  if (cfg->batch_size < 1 || cfg->batch_size > 7000) { geeco_set_error("batch_size %d outside [1,7000]", cfg->batch_size); return GEECO_ERR_INVALID; }
  if (cfg->dim_s_obs % 4 || cfg->dim_s_dyn % 4 || cfg->dim_s_diff % 4 || cfg->dim_h_lstm % 4 || cfg->dim_s_obs < 4 ||
      cfg->dim_s_dyn < 4 || cfg->dim_s_diff < 4 || cfg->dim_h_lstm < 4 || cfg->dim_h_fc < 1) {
    geeco_set_error("dim_s_obs/dim_s_dyn/dim_s_diff/dim_h_lstm must be positive multiples of 4");
    return GEECO_ERR_INVALID;
  }
  if (cfg->num_grp_states < 1 || cfg->num_grp_states > 6) { geeco_set_error("num_grp_states %d outside [1,6]", cfg->num_grp_states); return GEECO_ERR_INVALID; }
  if (cfg->dim_jnt_state < 1 || cfg->dim_jnt_state > 12) { geeco_set_error("dim_jnt_state %d outside [1,12]", cfg->dim_jnt_state); return GEECO_ERR_INVALID; }
  if (cfg->control_mode == GEECO_CTRL_VELOCITY && (cfg->dim_grp_command < 1 || cfg->dim_grp_command > 6)) {
    geeco_set_error("dim_grp_command %d outside [1,6]", cfg->dim_grp_command); return GEECO_ERR_INVALID;
  }
  if (cfg->precision != GEECO_FP32 && cfg->precision != GEECO_BF16) { geeco_set_error("unknown precision %d", cfg->precision); return GEECO_ERR_INVALID; }
  if (cfg->precision == GEECO_BF16 && (cfg->dim_s_obs % 16 || cfg->dim_s_dyn % 16 || cfg->dim_s_diff % 16 ||
                                        cfg->dim_s_obs > 256 || cfg->dim_s_dyn > 256 || cfg->dim_s_diff > 256)) {
    geeco_set_error("bf16 mode: dim_s_obs/dim_s_dyn/dim_s_diff must be multiples of 16 and <= 256");
    return GEECO_ERR_INVALID;
  }
  return GEECO_OK;
}

// graph variant -> encoder groups, images per group, LSTM steps, scopes, conv8 widths, LSTM input width
static void select_variant(geeco_ctx* c) {
  const geeco_config& cfg = c->cfg;
  const int N = cfg.batch_size, K = cfg.window_size, J = cfg.dim_jnt_state;
  c->dim8[0] = c->dim8[1] = c->dim8[2] = 0;
  if (cfg.goal_condition == GEECO_GOAL_NONE) {                          // e2e_vmc, graph.py:268-319
    c->variant = VAR_VMC; c->G = 1; c->M = K * N; c->T = K;
    c->enc_scope[0] = "VMC/ConvEncoder"; c->dec_scope = "VMC/LSTMDecoder/";
    c->dim8[0] = 256;                                                   // conv_encoder's default dim_out (graph.py:61)
    c->xdim = 4 * (256 + J);
    return;
  }
  c->dec_scope = "GoalVMC/LSTMDecoder/";
  c->enc_scope[0] = "GoalVMC/ConvEncoder";
  c->dim8[0] = cfg.dim_s_obs;
  if (cfg.proc_obs == GEECO_OBS_DYNIMG) {                               // graph.py:386-407 (proc_tgt is not consulted there)
    c->variant = VAR_GEECOF; c->G = 3; c->M = N; c->T = 1;
    c->enc_scope[1] = "GoalVMC/DynBuffEncoder"; c->enc_scope[2] = "GoalVMC/DynDiffEncoder";
    c->dim8[1] = cfg.dim_s_dyn; c->dim8[2] = cfg.dim_s_diff;
    c->xdim = 4 * (cfg.dim_s_obs + cfg.dim_s_dyn + J + cfg.dim_s_diff);
  } else if (cfg.proc_tgt == GEECO_TGT_CONSTANT) {                      // graph.py:352-355, :365-367
    c->variant = VAR_SEQ_CONSTANT; c->G = 1; c->M = (K + 1) * N; c->T = K;
    c->xdim = 4 * (2 * cfg.dim_s_obs + J);
  } else if (cfg.proc_tgt == GEECO_TGT_RESIDUAL) {                      // graph.py:368-370
    c->variant = VAR_SEQ_RESIDUAL; c->G = 1; c->M = (K + 1) * N; c->T = K;
    c->xdim = 4 * (cfg.dim_s_obs + J);
  } else {                                                              // graph.py:371-379
    c->variant = VAR_SEQ_DYNDIFF; c->G = 2; c->M = K * N; c->T = K;
    c->enc_scope[1] = "GoalVMC/DynDiffEncoder";
    c->dim8[1] = cfg.dim_s_diff;
    c->xdim = 4 * (cfg.dim_s_obs + J + cfg.dim_s_diff);
  }
}

extern "C" int geeco_head_columns(const geeco_config* cfg) {
  if (!cfg) return -1;
  return cfg->control_mode == GEECO_CTRL_VELOCITY ? cfg->dim_jnt_state + 3 + cfg->dim_grp_command + 6
                                                  : 9 + cfg->num_grp_states;
}

// builds layer table, parameter table and (when base != NULL) the workspace pointers
static int plan(geeco_ctx* c, char* ws_base) {
  const geeco_config& cfg = c->cfg;
  const int N = cfg.batch_size;
  c->params.clear();
  c->arena_floats = 0;
  const bool bf16 = cfg.precision == GEECO_BF16;
  c->CP = 4;   // network input is channel-padded 3 -> 4 (fp32: 16 B / pixel, bf16: 8 B / pixel)
  select_variant(c);
  const int G = c->G, M = c->M, T = c->T;
  // ---- layer geometry
  c->uniform8 = true;
  for (int e = 1; e < G; ++e) c->uniform8 = c->uniform8 && c->dim8[e] == c->dim8[0];
  int H = cfg.img_height, Cin_real = cfg.img_channels, Cin_pad = c->CP;
  for (int l = 0; l < 8; ++l) {
    LayerPlan& L = c->layers[l];
    L.Hin = H; L.stride = kEncStrides[l]; L.Hout = (H + L.stride - 1) / L.stride;
    L.Cin_real = Cin_real; L.Cin_pad = Cin_pad;
    for (int e = 0; e < 3; ++e) L.Cout[e] = l < 7 ? kEncChannels[l] : c->dim8[e < G ? e : 0];
    L.grouped = l < 7 || c->uniform8;
    H = L.Hout; Cin_real = L.Cout[0]; Cin_pad = L.Cout[0];
  }
  // ---- parameters, arena order = gradient-bucket order (late layers first)
  const int xdim = c->xdim;
  const int Hl = cfg.dim_h_lstm, Fc = cfg.dim_h_fc;
  const std::string dsc = c->dec_scope;
  c->p_lstm_w = add_param(c, dsc + "lstm_cell/kernel", {xdim + Hl, 4 * Hl});
  c->p_lstm_b = add_param(c, dsc + "lstm_cell/bias", {4 * Hl});
  c->p_fc1_w = add_param(c, dsc + "fc1/kernel", {Hl, Fc});
  c->p_fc1_b = add_param(c, dsc + "fc1/bias", {Fc});
  // heads in the order the graph creates them (graph.py:233-259); loss slots: include/geeco_b200.h
  if (cfg.control_mode == GEECO_CTRL_CARTESIAN) {
    c->nheads = 4;
    c->heads_plan[0] = {"pred_cmd_ee", 3, 0, 0, 0, 0, 0};
    c->heads_plan[1] = {"logits_cmd_grp", cfg.num_grp_states, 1, 1, 0, 0, 0};
    c->heads_plan[2] = {"pred_aux_ee", 3, 0, 2, 1, 0, 0};
    c->heads_plan[3] = {"pred_aux_obj", 3, 0, 3, 1, 0, 0};
  } else {
    c->nheads = 5;
    c->heads_plan[0] = {"pred_cmd_vel", cfg.dim_jnt_state, 0, 8, 0, 0, 0};
    c->heads_plan[1] = {"pred_cmd_ee", 3, 0, 0, 0, 0, 0};
    c->heads_plan[2] = {"pred_cmd_grp", cfg.dim_grp_command, 0, 1, 0, 0, 0};
    c->heads_plan[3] = {"pred_aux_ee", 3, 0, 2, 0, 0, 0};     // mse_loss adds all five terms with weight 1 (graph.py:446-449)
    c->heads_plan[4] = {"pred_aux_obj", 3, 0, 3, 0, 0, 0};
  }
  c->NH = 0;
  for (int h = 0; h < c->nheads; ++h) {
    HeadPlan& hp = c->heads_plan[h];
    hp.p_w = add_param(c, dsc + hp.name + "/kernel", {Fc, hp.width});
    hp.p_b = add_param(c, dsc + hp.name + "/bias", {hp.width});
    c->NH += hp.width;
  }
  // bucket 0 = LSTM + fc1 + heads (22 % of the bytes): complete as soon as the tail's backward is, before any encoder
  // gradient -- its all-reduce starts 0.2 ms earlier than when it shared a bucket with conv8..conv5 (73 %)
  c->arena_floats = (c->arena_floats + 3) & ~3ll;
  c->bucket_end[0] = c->arena_floats;
  for (int l = 7; l >= 0; --l) {
    LayerPlan& L = c->layers[l];
    char nm[128];
    for (int e = 0; e < G; ++e) {
      snprintf(nm, sizeof(nm), "%s/conv%d/kernel", c->enc_scope[e], l + 1);
      L.p_w[e] = add_param(c, nm, {3, 3, L.Cin_real, L.Cout[e]});
    }
    for (int e = 0; e < G; ++e) {
      snprintf(nm, sizeof(nm), "%s/conv%d/bias", c->enc_scope[e], l + 1);
      L.p_b[e] = add_param(c, nm, {L.Cout[e]});
    }
    if (l == 4) { c->arena_floats = (c->arena_floats + 3) & ~3ll; c->bucket_end[1] = c->arena_floats; }
    if (l == 2) { c->arena_floats = (c->arena_floats + 3) & ~3ll; c->bucket_end[2] = c->arena_floats; }
  }
  c->arena_floats = (c->arena_floats + 3) & ~3ll;
  c->bucket_end[3] = c->arena_floats;

  // ---- workspace
  Carver cv{ws_base, 0};
  const size_t esz = bf16 ? 2 : 4;
  const long long HW = (long long)cfg.img_height * cfg.img_width;
  c->x0 = cv.take((size_t)G * M * HW * c->CP * esz);
  for (int l = 0; l < 8; ++l) {
    LayerPlan& L = c->layers[l];
    long long off = 0;
    for (int e = 0; e < G; ++e) { L.act_off[e] = off; off += (long long)M * L.Hout * L.Hout * L.Cout[e]; }
    L.act_elems = off;
    L.y = cv.take((size_t)off * esz);
    L.g = cfg.training ? cv.take((size_t)off * esz) : nullptr;
    // the data gradient of layer l+1 needs only the SIGN of y_l: the forward epilogue also writes one bit per element
    L.mbits = (bf16 && cfg.training && l < 7) ? cv.take((size_t)off / 16 * 2) : nullptr;
    if (l == 0 && getenv("GEECO_NOBITS0")) L.mbits = nullptr;      // experiment: conv2 data gradient from the bf16 mask
  }
  const int ld = xdim + Hl;
  c->states = (float*)cv.take(sizeof(float) * T * N * ld);
  c->gates = (float*)cv.take(sizeof(float) * T * N * 4 * Hl);
  c->c_seq = (float*)cv.take(sizeof(float) * T * N * Hl);
  c->m_seq = (float*)cv.take(sizeof(float) * T * N * Hl);
  c->state_out = (float*)cv.take(sizeof(float) * N * 2 * Hl);
  c->c_carry = (float*)cv.take(sizeof(float) * N * Hl);
  c->m_carry = (float*)cv.take(sizeof(float) * N * Hl);
  c->fc1 = (float*)cv.take(sizeof(float) * N * Fc);
  c->heads = (float*)cv.take(sizeof(float) * N * c->NH);
  c->loss_parts = (float*)cv.take(sizeof(float) * N * 6);
  c->dheads = (float*)cv.take(sizeof(float) * N * c->NH);
  c->losses = (float*)cv.take(sizeof(float) * GEECO_NUM_LOSS_SLOTS);
  c->sc = (float*)cv.take(sizeof(float) * 8);
  c->mm_scratch = (int*)cv.take(sizeof(int) * 2 * cfg.window_size * N);
  c->gates_partial = (float*)cv.take(sizeof(float) * lstm_gates_partial_floats(T * N, ld, 4 * Hl));
  // fp32 staging of the last conv maps for the tail (bf16 mode converts conv8 output)
  c->y8_f32 = bf16 ? (float*)cv.take(sizeof(float) * c->layers[7].act_elems) : nullptr;
  c->g8_f32 = (bf16 && cfg.training) ? (float*)cv.take(sizeof(float) * c->layers[7].act_elems) : nullptr;
  if (cfg.training) {
    c->dfc1 = (float*)cv.take(sizeof(float) * N * Fc);
    c->dgates = (float*)cv.take(sizeof(float) * T * N * 4 * Hl);
    c->dstates = (float*)cv.take(sizeof(float) * T * N * ld);
    c->dm_last = (float*)cv.take(sizeof(float) * N * Hl);
    c->dc = (float*)cv.take(sizeof(float) * N * Hl);
    // split-K partial buffer: worst case over all TN launches
    long long cap = 0;
    for (int l = 0; l < 8; ++l) {
      LayerPlan& L = c->layers[l];
      const int groups = L.grouped ? G : 1;
      for (int e = 0; e < (L.grouped ? 1 : G); ++e) {
        GatherGeom g = conv_fwd_geom(L.Hin, L.Hin, L.Cin_pad, L.Cin_real, L.Cout[e], L.stride, M);
        long long need = gemm_tn_partial_floats(g, groups);
        if (need > cap) cap = need;
      }
    }
    {
      GatherGeom g = dense_geom(T * N, ld, 4 * Hl, 4 * Hl, 0);
      long long need = gemm_tn_partial_floats(g, 1);
      if (need > cap) cap = need;
    }
    c->partial_cap = cap;
    c->partial = (float*)cv.take(sizeof(float) * cap);
  }
  if (bf16) {
    int rc = plan_bf16(c, &cv.off, ws_base);
    if (rc) return rc;
  }
  c->workspace_bytes = (cv.off + 255) & ~(size_t)255;
  return GEECO_OK;
}

extern "C" int geeco_query_sizes(const geeco_config* cfg, geeco_sizes* out) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!out) { geeco_set_error("out is NULL"); return GEECO_ERR_INVALID; }
  geeco_ctx tmp;
  tmp.cfg = *cfg;
  rc = plan(&tmp, nullptr);
  free_bf16(&tmp);
  if (rc) return rc;
  out->arena_floats = tmp.arena_floats;
  out->workspace_bytes = (int64_t)tmp.workspace_bytes;
  out->num_params = (int32_t)tmp.params.size();
  out->num_buckets = 4;
  return GEECO_OK;
}

extern "C" int geeco_create(const geeco_config* cfg, geeco_ctx** out) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!out) { geeco_set_error("out is NULL"); return GEECO_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    geeco_set_error("no CUDA device: geeco_b200 has no CPU fallback");
    return GEECO_ERR_CUDA;
  }
  geeco_ctx* c = new geeco_ctx();
  c->cfg = *cfg;
  rc = plan(c, nullptr);
  if (rc) { delete c; return rc; }
  geeco_alpha_table(cfg->window_size, c->alpha);
  *out = c;
  return GEECO_OK;
}

extern "C" int geeco_destroy(geeco_ctx* ctx) {
  if (ctx) free_bf16(ctx);
  delete ctx;
  return GEECO_OK;
}

extern "C" int geeco_bind(geeco_ctx* c, float* theta, float* grad, float* m, float* v, void* workspace,
                          int64_t workspace_bytes) {
  if (!c || !theta || !workspace) { geeco_set_error("bind: NULL argument"); return GEECO_ERR_INVALID; }
  if (c->cfg.training && (!grad || !m || !v)) { geeco_set_error("bind: training context needs grad/m/v arenas"); return GEECO_ERR_INVALID; }
  if ((size_t)workspace_bytes < c->workspace_bytes) {
    geeco_set_error("bind: workspace of %lld bytes < required %zu", (long long)workspace_bytes, c->workspace_bytes);
    return GEECO_ERR_WORKSPACE;
  }
  if (((uintptr_t)workspace & 255) || ((uintptr_t)theta & 15) || ((uintptr_t)grad & 15) || ((uintptr_t)m & 15) || ((uintptr_t)v & 15)) {
    geeco_set_error("bind: workspace must be 256-byte aligned, arenas 16-byte aligned");
    return GEECO_ERR_INVALID;
  }
  c->theta = theta; c->grad = grad; c->m = m; c->v = v;
  int rc = plan(c, (char*)workspace);
  if (rc) return rc;
  CUDA_TRY(cudaMemset(c->sc, 0, sizeof(float) * 8));
  CUDA_TRY(cudaMemset(c->c_carry, 0, sizeof(float) * c->cfg.batch_size * c->cfg.dim_h_lstm));
  CUDA_TRY(cudaMemset(c->m_carry, 0, sizeof(float) * c->cfg.batch_size * c->cfg.dim_h_lstm));
  c->bound = true;
  c->weights_dirty = true;
  c->fwd_done = false;
  c->host_step = 0;
  return GEECO_OK;
}

extern "C" int geeco_param_info(const geeco_ctx* c, int32_t index, geeco_param_desc* out) {
  if (!c || !out || index < 0 || index >= (int)c->params.size()) { geeco_set_error("param_info: bad index %d", index); return GEECO_ERR_INVALID; }
  *out = c->params[index];
  return GEECO_OK;
}

extern "C" int geeco_grad_bucket(const geeco_ctx* c, int32_t b, int64_t* offset, int64_t* numel) {
  if (!c || b < 0 || b > 3 || !offset || !numel) { geeco_set_error("grad_bucket: bad arguments"); return GEECO_ERR_INVALID; }
  const long long lo = b == 0 ? 0 : c->bucket_end[b - 1];
  *offset = lo; *numel = c->bucket_end[b] - lo;
  return GEECO_OK;
}

extern "C" int geeco_params_changed(geeco_ctx* c, void* stream) {
  if (!c || !c->bound) { geeco_set_error("params_changed: context not bound"); return GEECO_ERR_STATE; }
  c->weights_dirty = true;
  (void)stream;
  return GEECO_OK;
}

extern "C" int geeco_set_step(geeco_ctx* c, int64_t t, void* stream) {
  if (!c || !c->bound) { geeco_set_error("set_step: context not bound"); return GEECO_ERR_STATE; }
  c->host_step = (long long)t;
  c->host_sc[0] = (float)t; c->host_sc[1] = 0.f; c->host_sc[2] = 0.f; c->host_sc[3] = 0.f;
  CUDA_TRY(cudaMemcpyAsync(c->sc, c->host_sc, 4 * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return GEECO_OK;
}

extern "C" int geeco_set_lstm_state(geeco_ctx* c, const float* state_cm, void* stream) {
  if (!c || !c->bound) { geeco_set_error("set_lstm_state: context not bound"); return GEECO_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = c->cfg.batch_size, Hl = c->cfg.dim_h_lstm;
  if (!state_cm) {
    CUDA_TRY(cudaMemsetAsync(c->c_carry, 0, sizeof(float) * N * Hl, st));
    CUDA_TRY(cudaMemsetAsync(c->m_carry, 0, sizeof(float) * N * Hl, st));
  } else {
    CUDA_TRY(cudaMemcpy2DAsync(c->c_carry, sizeof(float) * Hl, state_cm, sizeof(float) * 2 * Hl, sizeof(float) * Hl, N, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpy2DAsync(c->m_carry, sizeof(float) * Hl, state_cm + Hl, sizeof(float) * 2 * Hl, sizeof(float) * Hl, N, cudaMemcpyDeviceToDevice, st));
  }
  return GEECO_OK;
}

// ------------------------------------------------------------------------------------------
// functional ops
// ------------------------------------------------------------------------------------------
extern "C" int geeco_dynimg(const float* in, float* out, int32_t N, int32_t K, int32_t H, int32_t W, int32_t C,
                            const float* alpha, int32_t cluster, float* scratch, void* stream) {
  if (N == 0) return GEECO_OK;
  if (!in || !out || N < 0) { geeco_set_error("dynimg: NULL tensor"); return GEECO_ERR_INVALID; }
  if (K < 2 || K > 16) { geeco_set_error("dynimg: window size K=%d outside [2,16]", K); return GEECO_ERR_INVALID; }
  float tab[16];
  if (!alpha) { geeco_alpha_table(K, tab); alpha = tab; }
  const long long HWC = (long long)H * W * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (cluster == 0) {
    // measured choice (tools/sweep_dynimg.py on B200, profiles/r02_rankpool_sweep.txt): the one-read cluster kernel
    // wins while a sample's slice and the K-deep load batch fit comfortably; large samples with deep windows do
    // better as two streaming passes (d written un-normalised, normalised in place), and K = 3 / 5..7 at 256 px
    // prefer 4 CTAs with the whole SM's shared memory over 8 with half
    const long long bytes = HWC * 4;
    if (bytes >= (2ll << 20)) { if (K >= 8 && scratch) cluster = -1; }
    else if (bytes >= (512ll << 10)) {
      if (K == 8 && scratch) cluster = -1;
      else if (K == 3 || (K >= 5 && K <= 7)) cluster = 4;
    }
  }
  if (cluster >= 0) {
    int rc = launch_dynimg(in, out, N, K, HWC, alpha, cluster, st);
    if (rc != GEECO_ERR_WORKSPACE) return rc;
  }
  if (!scratch) { geeco_set_error("dynimg: two-pass path needs a scratch buffer of 2*N floats"); return GEECO_ERR_INVALID; }
  return launch_dynimg_twopass(in, out, scratch, N, K, HWC, alpha, st);
}

extern "C" int geeco_conv2d_same(const float* x, const float* w, const float* b, float* y, int32_t N, int32_t H,
                                 int32_t W, int32_t Cin, int32_t Cout, int32_t stride, int32_t relu, void* stream) {
  if (!x || !w || !y) { geeco_set_error("conv2d: NULL tensor"); return GEECO_ERR_INVALID; }
  GatherGeom g = conv_fwd_geom(H, W, Cin, Cin, Cout, stride, N);
  if (H != W) { int Wo, pl; same_pad(W, 3, stride, &Wo, &pl); g.Ws = W; g.Wm = Wo; g.Wd = Wo; }
  return launch_gemm_nn_f32(g, x, w, b, nullptr, y, 1, b ? (relu ? EPI_BIAS_RELU : EPI_BIAS) : EPI_STORE,
                            (cudaStream_t)stream);
}

extern "C" int64_t geeco_conv2d_bwd_scratch_floats(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                                   int32_t stride) {
  GatherGeom g = conv_fwd_geom(H, W, Cin, Cin, Cout, stride, N);
  return gemm_tn_partial_floats(g, 1);
}

extern "C" int geeco_conv2d_same_bwd(const float* x, const float* w, const float* dy_pre, const float* relu_mask_x,
                                     float* dw, float* db, float* dx, float* scratch, int64_t scratch_floats,
                                     int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t stride,
                                     void* stream) {
  if (H != W) { geeco_set_error("conv2d_bwd: square inputs only"); return GEECO_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (dw) {
    GatherGeom g = conv_fwd_geom(H, W, Cin, Cin, Cout, stride, N);
    int rc = launch_gemm_tn_f32(g, x, dy_pre, dw, db, scratch, scratch_floats, 1, 0, 0, st);
    if (rc) return rc;
  }
  if (dx) {
    for (int py = 0; py < stride; ++py)
      for (int px = 0; px < stride; ++px) {
        GatherGeom g;
        if (!conv_dgrad_geom(H, W, Cin, Cout, stride, py, px, N, &g)) continue;
        int rc = launch_gemm_nn_f32(g, dy_pre, w, nullptr, relu_mask_x, dx, 1, relu_mask_x ? EPI_MASK : EPI_STORE, st);
        if (rc) return rc;
      }
  }
  return GEECO_OK;
}

// ------------------------------------------------------------------------------------------
// model step (fp32 path here; bf16 path in step_bf16.cu)
// ------------------------------------------------------------------------------------------
static inline float* P(geeco_ctx* c, int idx) { return c->theta + c->params[idx].offset; }
static inline float* GR(geeco_ctx* c, int idx) { return c->grad + c->params[idx].offset; }

static TailDims tail_dims(const geeco_ctx* c) {
  TailDims d;
  d.N = c->cfg.batch_size; d.Hl = c->cfg.dim_h_lstm; d.Fc = c->cfg.dim_h_fc; d.G = c->cfg.num_grp_states;
  return d;
}

// head table of this step: parameter / gradient pointers and where each head finds its target in the batch
// (estimator.py:206-216 cartesian, :229-236 velocity; the auxiliary poses come from the LAST frame of the window)
static TailHeads tail_heads(geeco_ctx* c, const geeco_batch* b) {
  const geeco_config& cfg = c->cfg;
  const int K = cfg.window_size;
  const int last = (c->ring_start + K - 1) % K;
  TailHeads th;
  memset(&th, 0, sizeof(th));
  th.nheads = c->nheads; th.NH = c->NH;
  int col = 0;
  for (int h = 0; h < c->nheads; ++h) {
    const HeadPlan& hp = c->heads_plan[h];
    HeadSpec& hs = th.h[h];
    hs.col = col; hs.width = hp.width; hs.kind = hp.kind; hs.slot = hp.slot; hs.aux = hp.aux;
    hs.weight = hp.aux ? cfg.lambda_aux : 1.f;
    hs.w = P(c, hp.p_w); hs.b = P(c, hp.p_b);
    hs.gw = c->grad ? GR(c, hp.p_w) : nullptr; hs.gb = c->grad ? GR(c, hp.p_b) : nullptr;
    col += hp.width;
    if (!b) continue;
    const std::string nm = hp.name;
    if (nm == "pred_aux_ee") { hs.target = b->ee_state; hs.tstride = K * 7; hs.toff = last * 7; }
    else if (nm == "pred_aux_obj") { hs.target = b->obj_state; hs.tstride = K * 7; hs.toff = last * 7; }
    else if (nm == "pred_cmd_vel") { hs.target = b->vel_target; hs.tstride = cfg.dim_jnt_state; hs.toff = 0; }
    else if (nm == "pred_cmd_grp") { hs.target = b->grp_target; hs.tstride = cfg.dim_grp_command; hs.toff = 0; }
    else if (nm == "logits_cmd_grp") { hs.target = b->cmd; hs.tstride = 4; hs.toff = 3; }
    else if (cfg.control_mode == GEECO_CTRL_VELOCITY) { hs.target = b->ee_target; hs.tstride = 7; hs.toff = 0; }   // pred_cmd_ee
    else { hs.target = b->cmd; hs.tstride = 4; hs.toff = 0; }                                                       // pred_cmd_ee
  }
  return th;
}

static StateMap state_map(geeco_ctx* c, const float* y8, float* g8) {
  const geeco_config& cfg = c->cfg;
  StateMap sm;
  memset(&sm, 0, sizeof(sm));
  sm.variant = c->variant; sm.N = cfg.batch_size; sm.T = c->T; sm.K = cfg.window_size; sm.J = cfg.dim_jnt_state;
  sm.D0 = c->dim8[0]; sm.D1 = c->dim8[1]; sm.D2 = c->variant == VAR_SEQ_DYNDIFF ? c->dim8[1] : c->dim8[2];
  sm.xdim = c->xdim; sm.per = c->xdim / 4; sm.Hl = cfg.dim_h_lstm; sm.ld = c->xdim + cfg.dim_h_lstm;
  const LayerPlan& L8 = c->layers[7];
  for (int e = 0; e < c->G; ++e) { sm.y[e] = y8 + L8.act_off[e]; sm.g[e] = g8 ? g8 + L8.act_off[e] : nullptr; }
  sm.ring_start = c->ring_start;
  return sm;
}

static int check_batch(const geeco_ctx* c, const geeco_batch* b, bool need_labels) {
  if (!c || !c->bound) { geeco_set_error("context not bound (call geeco_bind first)"); return GEECO_ERR_STATE; }
  const bool goal = c->cfg.goal_condition == GEECO_GOAL_TARGET;
  if (!b || !b->rgb || !b->jnt_state || (goal && !b->target_rgb)) { geeco_set_error("batch: rgb / target_rgb / jnt_state must be given"); return GEECO_ERR_INVALID; }
  if (b->ring_start < 0 || b->ring_start >= c->cfg.window_size) { geeco_set_error("batch: ring_start %d outside [0,%d)", b->ring_start, c->cfg.window_size); return GEECO_ERR_INVALID; }
  if (need_labels) {
    const bool vel = c->cfg.control_mode == GEECO_CTRL_VELOCITY;
    if (!b->ee_state || !b->obj_state || (!vel && !b->cmd) || (vel && (!b->vel_target || !b->ee_target || !b->grp_target))) {
      geeco_set_error(vel ? "batch: vel_target / ee_target / grp_target / ee_state / obj_state needed for the losses"
                          : "batch: cmd / ee_state / obj_state needed for the losses");
      return GEECO_ERR_INVALID;
    }
  }
  return GEECO_OK;
}

// conv stack forward, fp32
static int encoders_fwd_f32(geeco_ctx* c, cudaStream_t st) {
  const int M = c->M, G = c->G;
  const float* src = (const float*)c->x0;
  for (int l = 0; l < 8; ++l) {
    LayerPlan& L = c->layers[l];
    if (L.grouped) {
      GatherGeom g = conv_fwd_geom(L.Hin, L.Hin, L.Cin_pad, L.Cin_real, L.Cout[0], L.stride, M);
      g.b_group_stride = w_group_stride(c, L);
      g.bias_group_stride = b_group_stride(c, L);
      int rc = launch_gemm_nn_f32(g, src, P(c, L.p_w[0]), P(c, L.p_b[0]), nullptr, (float*)L.y, G, EPI_BIAS_RELU, st);
      if (rc) return rc;
    } else {
      for (int e = 0; e < G; ++e) {
        GatherGeom g = conv_fwd_geom(L.Hin, L.Hin, L.Cin_pad, L.Cin_real, L.Cout[e], L.stride, M);
        const float* s = src + (long long)e * M * L.Hin * L.Hin * L.Cin_pad;
        int rc = launch_gemm_nn_f32(g, s, P(c, L.p_w[e]), P(c, L.p_b[e]), nullptr, (float*)L.y + L.act_off[e], 1, EPI_BIAS_RELU, st);
        if (rc) return rc;
      }
    }
    src = (const float*)L.y;
  }
  return GEECO_OK;
}

static bool use_persistent_lstm(const geeco_ctx* c) {
  static const bool off = getenv("GEECO_LSTM_STEPWISE") != nullptr;      // per-step launches (kept as the fallback)
  return c->T > 1 && !off && lstm_persistent_supported(c->cfg.dim_h_lstm);
}

static int tail_forward(geeco_ctx* c, const geeco_batch* b, const geeco_outputs* out, const float* y8, bool with_loss,
                        cudaStream_t st) {
  const geeco_config& cfg = c->cfg;
  const int N = cfg.batch_size, Hl = cfg.dim_h_lstm, T = c->T, ld = c->xdim + Hl;
  TailDims d = tail_dims(c);
  const bool carry = cfg.carry_state != 0;
  const unsigned char* rmask = carry ? b->reset_mask : nullptr;
  c->reset_mask = rmask;
  StateMap sm = state_map(c, y8, nullptr);
  int rc = launch_build_states(sm, b->jnt_state, carry ? c->m_carry : nullptr, rmask, c->states, st);
  if (rc) return rc;
  // lstm_decoder's loop over feat_list (graph.py:223-225): T = 1 for the dynimg graph, K for the sequence graphs.
  // T > 1: the x part of all T steps is ONE split GEMM over T*N rows, the recurrence ONE persistent cluster launch
  // with W_h resident in shared memory (lstm_persistent.cu)
  if (with_loss && cfg.l2_regularizer > 0.f) {
    rc = launch_l2_term(c->theta, c->arena_floats, cfg.l2_regularizer, c->sc, st);
    if (rc) return rc;
  }
  TailHeads th = tail_heads(c, with_loss ? b : nullptr);
  float* o_heads = out ? out->heads : nullptr;
  float* o_fc1 = out ? out->fc1 : nullptr;
  float* o_state = out ? out->lstm_state : nullptr;
  if (T == 1) {
    // one step: split-K gate GEMM (partials only), then ONE launch for reduce + cell + fc1 + heads + per-sample losses
    int slices = 0;
    rc = launch_lstm_gates(c->states, ld, P(c, c->p_lstm_w), P(c, c->p_lstm_b), c->gates, c->gates_partial, N,
                           carry ? ld : c->xdim, 4 * Hl, st, &slices);
    if (rc) return rc;
    rc = launch_tail_fwd_fused(d, th, P(c, c->p_fc1_w), P(c, c->p_fc1_b), c->gates_partial, slices, P(c, c->p_lstm_b), c->gates,
                               carry ? c->c_carry : nullptr, rmask, c->c_seq, c->m_seq, c->state_out, c->fc1, c->heads,
                               c->loss_parts, c->dheads, with_loss ? 1 : 0, o_heads, o_fc1, o_state, st);
    if (rc) return rc;
  } else {
    if (use_persistent_lstm(c)) {
      rc = launch_lstm_gates(c->states, ld, P(c, c->p_lstm_w), P(c, c->p_lstm_b), c->gates, c->gates_partial, T * N, c->xdim,
                             4 * Hl, st);
      if (rc) return rc;
      rc = launch_lstm_seq_fwd(T, N, Hl, c->xdim, P(c, c->p_lstm_w), carry ? c->c_carry : nullptr, carry ? c->m_carry : nullptr,
                               rmask, c->gates, c->c_seq, c->m_seq, c->states, c->state_out, st);
      if (rc) return rc;
    } else
    for (int t = 0; t < T; ++t) {
      const bool has_prev = t > 0 || carry;
      float* gates_t = c->gates + (long long)t * N * 4 * Hl;
      // without a previous state m_prev == 0 and the h-rows of the kernel contribute nothing: K = xdim
      rc = launch_lstm_gates(c->states + (long long)t * N * ld, ld, P(c, c->p_lstm_w), P(c, c->p_lstm_b), gates_t,
                             c->gates_partial, N, has_prev ? ld : c->xdim, 4 * Hl, st);
      if (rc) return rc;
      const float* c_prev = t > 0 ? c->c_seq + (long long)(t - 1) * N * Hl : (carry ? c->c_carry : nullptr);
      rc = launch_lstm_cell(N, Hl, gates_t, c_prev, t == 0 ? rmask : nullptr, c->c_seq + (long long)t * N * Hl,
                            c->m_seq + (long long)t * N * Hl, t == T - 1 ? c->state_out : nullptr,
                            t + 1 < T ? c->states + (long long)(t + 1) * N * ld + c->xdim : nullptr, ld, st);
      if (rc) return rc;
    }
    rc = launch_tail_fwd(d, th, P(c, c->p_fc1_w), P(c, c->p_fc1_b), c->m_seq + (long long)(T - 1) * N * Hl, c->fc1, c->heads,
                         c->loss_parts, c->dheads, with_loss ? 1 : 0, o_heads, o_fc1, st);
    if (rc) return rc;
    if (o_state) {
      const float* src[4] = {c->state_out, nullptr, nullptr, nullptr};
      float* dst[4] = {o_state, nullptr, nullptr, nullptr};
      const long long n[4] = {(long long)N * 2 * Hl, 0, 0, 0};
      rc = launch_copy_outputs(src, dst, n, st);
      if (rc) return rc;
    }
  }
  if (with_loss) {
    rc = launch_loss_reduce(d, th, c->loss_parts, cfg.l2_regularizer > 0.f ? c->sc + 2 : nullptr, c->losses,
                            out ? out->losses : nullptr, st);
    if (rc) return rc;
  }
  return GEECO_OK;
}

static int carry_update(geeco_ctx* c, cudaStream_t st) {
  if (!c->cfg.carry_state) return GEECO_OK;
  const long long last = (long long)(c->T - 1) * c->cfg.batch_size * c->cfg.dim_h_lstm;
  const size_t bytes = sizeof(float) * c->cfg.batch_size * c->cfg.dim_h_lstm;
  CUDA_TRY(cudaMemcpyAsync(c->c_carry, c->c_seq + last, bytes, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c->m_carry, c->m_seq + last, bytes, cudaMemcpyDeviceToDevice, st));
  return GEECO_OK;
}

static int forward_impl(geeco_ctx* c, const geeco_batch* b, const geeco_outputs* out, bool with_loss, cudaStream_t st) {
  const geeco_config& cfg = c->cfg;
  const bool bf16 = cfg.precision == GEECO_BF16;
  if (b->frame_format != GEECO_FRAMES_F32 && b->frame_format != GEECO_FRAMES_U8) {
    geeco_set_error("batch: frame_format %d is neither GEECO_FRAMES_F32 nor GEECO_FRAMES_U8", b->frame_format);
    return GEECO_ERR_INVALID;
  }
  c->ring_start = b->ring_start;
  int rc = bf16 ? repack_fork_bf16(c, st) : GEECO_OK;
  if (rc) return rc;
  const int u8 = b->frame_format == GEECO_FRAMES_U8;
  if (c->variant == VAR_GEECOF) {
    rc = launch_preprocess_geecof(b->rgb, b->target_rgb, u8, c->x0, bf16 ? 1 : 0, c->CP, out ? out->dynbuff : nullptr,
                                  out ? out->dyndiff : nullptr, cfg.batch_size, cfg.window_size, cfg.img_height,
                                  cfg.img_width, cfg.img_channels, c->alpha, 0, b->ring_start, b->frame_index, b->target_index, st);
  } else {
    const int with_tgt = c->variant == VAR_SEQ_CONSTANT || c->variant == VAR_SEQ_RESIDUAL;
    const int with_diff = c->variant == VAR_SEQ_DYNDIFF;
    rc = launch_preprocess_seq(b->rgb, b->target_rgb, u8, c->x0, bf16 ? 1 : 0, c->CP, c->mm_scratch,
                               out ? out->dyndiff : nullptr, cfg.batch_size, cfg.window_size, cfg.img_height,
                               cfg.img_width, cfg.img_channels, with_tgt, with_diff, b->ring_start, b->frame_index, b->target_index, st);
  }
  if (rc) return rc;
  const float* y8;
  if (bf16) {
    rc = repack_join_bf16(c, st);
    if (rc) return rc;
    rc = encoders_fwd_bf16(c, st);
    if (rc) return rc;
    y8 = c->y8_f32;
  } else {
    rc = encoders_fwd_f32(c, st);
    if (rc) return rc;
    y8 = (const float*)c->layers[7].y;
  }
  return tail_forward(c, b, out, y8, with_loss, st);
}

extern "C" int geeco_forward(geeco_ctx* c, const geeco_batch* b, const geeco_outputs* out, void* stream) {
  const bool with_loss = b && (b->cmd || b->vel_target) && out && out->losses;
  int rc = check_batch(c, b, with_loss);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = forward_impl(c, b, out, with_loss, st);
  if (rc) return rc;
  return carry_update(c, st);
}

extern "C" int geeco_step_forward(geeco_ctx* c, const geeco_batch* b, const geeco_outputs* out, void* stream) {
  int rc = check_batch(c, b, true);
  if (rc) return rc;
  if (!c->cfg.training) { geeco_set_error("context was created with training = 0"); return GEECO_ERR_STATE; }
  rc = forward_impl(c, b, out, true, (cudaStream_t)stream);
  if (rc) return rc;
  c->fwd_done = true;
  return GEECO_OK;
}

static int conv_layer_bwd_f32(geeco_ctx* c, int l, cudaStream_t st) {
  const int M = c->M, G = c->G;
  LayerPlan& L = c->layers[l];
  const float* xin = l == 0 ? (const float*)c->x0 : (const float*)c->layers[l - 1].y;
  const int ngroups = L.grouped ? G : 1;
  for (int e = 0; e < (L.grouped ? 1 : G); ++e) {
    const long long in_off = L.grouped ? 0 : (long long)e * M * L.Hin * L.Hin * L.Cin_pad;
    const float* gy = (const float*)L.g + (L.grouped ? 0 : L.act_off[e]);
    GatherGeom g = conv_fwd_geom(L.Hin, L.Hin, L.Cin_pad, L.Cin_real, L.Cout[e], L.stride, M);
    const long long wstride = w_group_stride(c, L);
    const long long bstride = b_group_stride(c, L);
    int rc = launch_gemm_tn_f32(g, xin + in_off, gy, GR(c, L.p_w[e]), GR(c, L.p_b[e]), c->partial, c->partial_cap,
                                ngroups, wstride, bstride, st);
    if (rc) return rc;
    if (l == 0) continue;
    for (int py = 0; py < L.stride; ++py)
      for (int px = 0; px < L.stride; ++px) {
        GatherGeom dg;
        if (!conv_dgrad_geom(L.Hin, L.Hin, L.Cin_real, L.Cout[e], L.stride, py, px, M, &dg)) continue;
        dg.b_group_stride = wstride;
        rc = launch_gemm_nn_f32(dg, gy, P(c, L.p_w[e]), nullptr, xin + in_off, (float*)c->layers[l - 1].g + in_off,
                                ngroups, EPI_MASK, st);
        if (rc) return rc;
      }
  }
  return GEECO_OK;
}

// backward of the tail: heads + fc1 -> dL/dm_T, back-propagation through the T LSTM steps, d(kernel) / d(bias) as ONE
// split GEMM over the T*N state rows, d(states) scattered into the conv8 gradients
static int tail_backward(geeco_ctx* c, cudaStream_t st) {
  const geeco_config& cfg = c->cfg;
  const bool bf16 = cfg.precision == GEECO_BF16;
  const int N = cfg.batch_size, Hl = cfg.dim_h_lstm, T = c->T, ld = c->xdim + Hl;
  const bool carry = cfg.carry_state != 0;
  TailDims d = tail_dims(c);
  TailHeads th = tail_heads(c, nullptr);
  const float* m_last = c->m_seq + (long long)(T - 1) * N * Hl;
  int rc;
  LayerPlan& L8 = c->layers[7];
  StateMap sm = state_map(c, bf16 ? c->y8_f32 : (const float*)L8.y, bf16 ? c->g8_f32 : (float*)L8.g);
  bool scattered = false;
  if (T == 1) {
    rc = launch_tail_bwd(d, th, P(c, c->p_fc1_w), GR(c, c->p_fc1_w), GR(c, c->p_fc1_b), m_last, c->fc1, c->dheads, c->gates,
                         carry ? c->c_carry : nullptr, c->reset_mask, c->dfc1, c->dgates, nullptr, st);
    if (rc) return rc;
    if (c->variant == VAR_GEECOF) {
      // d(state) is never materialised: the GEMM's epilogue writes the masked conv8 gradients directly (bf16 mode: as
      // bf16, which also saves the fp32 -> bf16 conversion launch)
      __nv_bfloat16* gb[3] = {nullptr, nullptr, nullptr};
      if (bf16) for (int e = 0; e < c->G; ++e) gb[e] = (__nv_bfloat16*)L8.g + L8.act_off[e];
      rc = launch_lstm_dstate_scatter(c->dgates, P(c, c->p_lstm_w), sm, bf16 ? gb : nullptr, 4 * Hl, st);
      if (rc) return rc;
      scattered = true;
      c->g8_bf16_ready = bf16;
    } else {
      rc = launch_lstm_dstate(c->dgates, P(c, c->p_lstm_w), c->dstates, N, c->xdim, 4 * Hl, ld, st);
      if (rc) return rc;
    }
  } else {
    rc = launch_tail_bwd(d, th, P(c, c->p_fc1_w), GR(c, c->p_fc1_w), GR(c, c->p_fc1_b), m_last, c->fc1, c->dheads, nullptr,
                         nullptr, nullptr, c->dfc1, nullptr, c->dm_last, st);
    if (rc) return rc;
    if (use_persistent_lstm(c)) {
      rc = launch_lstm_seq_bwd(T, N, Hl, c->xdim, P(c, c->p_lstm_w), carry ? c->c_carry : nullptr, c->reset_mask, c->gates,
                               c->c_seq, c->dm_last, c->dgates, st);
      if (rc) return rc;
      // d(x_t) of all steps: d(gates) [T*N, 4Hl] @ W_x^T in one launch
      rc = launch_lstm_dstate(c->dgates, P(c, c->p_lstm_w), c->dstates, T * N, c->xdim, 4 * Hl, ld, st);
      if (rc) return rc;
    } else
    for (int t = T - 1; t >= 0; --t) {
      const bool last = t == T - 1;
      float* dgates_t = c->dgates + (long long)t * N * 4 * Hl;
      float* dstate_t = c->dstates + (long long)t * N * ld;
      const float* c_prev = t > 0 ? c->c_seq + (long long)(t - 1) * N * Hl : (carry ? c->c_carry : nullptr);
      // d(m_t): from the decoder for the last step, else the m part of d(state_{t+1})
      rc = launch_lstm_cell_bwd(N, Hl, c->gates + (long long)t * N * 4 * Hl, c_prev, t == 0 ? c->reset_mask : nullptr,
                                last ? c->dm_last : dstate_t + (long long)N * ld + c->xdim, last ? Hl : ld,
                                last ? nullptr : c->dc, dgates_t, c->dc, st);
      if (rc) return rc;
      // d(state_t) = d(gates_t) @ kernel^T: the x part for the encoders; for t > 0 also the m part, which feeds step t-1
      rc = launch_lstm_dstate(dgates_t, P(c, c->p_lstm_w), dstate_t, N, t > 0 ? ld : c->xdim, 4 * Hl, ld, st);
      if (rc) return rc;
    }
  }
  rc = launch_lstm_wgrad(c->states, c->dgates, GR(c, c->p_lstm_w), GR(c, c->p_lstm_b), T * N, ld, 4 * Hl, st);
  if (rc) return rc;
  if (scattered) return GEECO_OK;
  c->g8_bf16_ready = false;
  return launch_scatter_dstates(sm, c->dstates, st);
}

extern "C" int geeco_step_backward(geeco_ctx* c, int32_t bucket, void* stream) {
  if (!c || !c->bound || !c->fwd_done) { geeco_set_error("step_backward: call geeco_step_forward first"); return GEECO_ERR_STATE; }
  if (bucket < 0 || bucket > 3) { geeco_set_error("step_backward: bucket %d outside [0,3]", bucket); return GEECO_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf16 = c->cfg.precision == GEECO_BF16;
  int rc;
  // bucket 0: LSTM + fc1 + heads; 1: conv8..conv5; 2: conv4, conv3; 3: conv2, conv1
  if (bucket == 0) return tail_backward(c, st);
  const int lhi = bucket == 1 ? 7 : (bucket == 2 ? 3 : 1);
  const int llo = bucket == 1 ? 4 : (bucket == 2 ? 2 : 0);
  if (bf16) return encoders_bwd_bf16(c, lhi, llo, st);
  for (int l = lhi; l >= llo; --l) {
    rc = conv_layer_bwd_f32(c, l, st);
    if (rc) return rc;
  }
  return GEECO_OK;
}

extern "C" int geeco_step_update(geeco_ctx* c, float grad_scale, void* stream) {
  if (!c || !c->bound || !c->fwd_done) { geeco_set_error("step_update: call geeco_step_forward/backward first"); return GEECO_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  const geeco_config& cfg = c->cfg;
  c->host_step += 1;
  int rc = launch_adam(c->theta, c->grad, c->m, c->v, c->arena_floats, c->host_step, cfg.lr, cfg.adam_beta1, cfg.adam_beta2,
                       cfg.adam_eps, grad_scale, cfg.l2_regularizer, st);
  if (rc) return rc;
  c->weights_dirty = true;
  c->fwd_done = false;
  return carry_update(c, st);
}

// The optimizer step over the gradient buckets [first, last] only, so that the host can update the buckets whose
// all-reduce has finished while a later one is still in flight (the last bucket -- conv2 + conv1 -- can only be reduced
// after the last backward kernel: its small all-reduce used to sit exposed in front of the whole Adam launch).  Every
// bucket must be updated exactly once per step; the call whose `last` is the final bucket closes the step.
extern "C" int geeco_step_update_buckets(geeco_ctx* c, float grad_scale, int32_t first, int32_t last, void* stream) {
  if (!c || !c->bound || !c->fwd_done) { geeco_set_error("step_update_buckets: call geeco_step_forward/backward first"); return GEECO_ERR_STATE; }
  if (first < 0 || last > 3 || first > last) { geeco_set_error("step_update_buckets: buckets [%d, %d] outside [0, 3]", first, last); return GEECO_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const geeco_config& cfg = c->cfg;
  const long long lo = first == 0 ? 0 : c->bucket_end[first - 1], hi = c->bucket_end[last];
  int rc = launch_adam(c->theta + lo, c->grad + lo, c->m + lo, c->v + lo, hi - lo, c->host_step + 1, cfg.lr, cfg.adam_beta1,
                       cfg.adam_beta2, cfg.adam_eps, grad_scale, cfg.l2_regularizer, st);
  if (rc) return rc;
  if (last < 3) return GEECO_OK;
  c->host_step += 1;
  c->weights_dirty = true;
  c->fwd_done = false;
  return carry_update(c, st);
}

extern "C" int geeco_train_step(geeco_ctx* c, const geeco_batch* b, const geeco_outputs* out, float grad_scale,
                                void* stream) {
  // (Replaying the step as a CUDA graph was measured: 2.657 vs 2.659 ms. The ~2 us between dependent kernels are
  // kernel drain + launch latency, not host enqueue, so the step stays a plain launch sequence.)
  int rc = geeco_step_forward(c, b, out, stream);
  if (rc) return rc;
  for (int bucket = 0; bucket < 4; ++bucket) {
    rc = geeco_step_backward(c, bucket, stream);
    if (rc) return rc;
  }
  return geeco_step_update(c, grad_scale, stream);
}

extern "C" int geeco_ring_push(void* ring, const void* frame, const uint8_t* fresh, int32_t N, int32_t K, int64_t row_bytes,
                               int32_t slot, void* stream) {
  if (!ring || !frame || K < 1 || slot < 0 || slot >= K || row_bytes <= 0 || (row_bytes & 3)) {
    geeco_set_error("ring_push: bad arguments (K=%d slot=%d row_bytes=%lld)", K, slot, (long long)row_bytes);
    return GEECO_ERR_INVALID;
  }
  return launch_ring_push(ring, frame, fresh, N, K, row_bytes, slot, (cudaStream_t)stream);
}

extern "C" int geeco_profile_kernel(geeco_ctx* c, const char* name, void* stream) {
  if (!c || !c->bound || !name) { geeco_set_error("profile_kernel: bad arguments"); return GEECO_ERR_INVALID; }
  if (c->cfg.precision != GEECO_BF16) { geeco_set_error("profile_kernel: bf16 contexts only"); return GEECO_ERR_INVALID; }
  return profile_kernel_bf16(c, name, (cudaStream_t)stream);
}

extern "C" int geeco_debug_buffer(const geeco_ctx* c, const char* name, void** ptr, int64_t* numel, int32_t* dtype) {
  if (!c || !c->bound || !name || !ptr || !numel || !dtype) { geeco_set_error("debug_buffer: bad arguments"); return GEECO_ERR_INVALID; }
  const geeco_config& cfg = c->cfg;
  const int N = cfg.batch_size, Hl = cfg.dim_h_lstm;
  const int act_dt = cfg.precision == GEECO_BF16 ? 1 : 0;
  *dtype = 0;
  std::string s(name);
  if (s == "x0") { *ptr = c->x0; *numel = (long long)c->G * c->M * cfg.img_height * cfg.img_width * c->CP; *dtype = act_dt; return GEECO_OK; }
  if (s.size() == 2 && (s[0] == 'y' || s[0] == 'g') && s[1] >= '1' && s[1] <= '8') {
    const LayerPlan& L = c->layers[s[1] - '1'];
    if (s == "y1" && c->y1_stale) {
      // the last training forward kept y1 on chip: rebuild it (debug only; default stream, synchronous)
      geeco_ctx* cm = const_cast<geeco_ctx*>(c);
      if (cudaDeviceSynchronize() != cudaSuccess) { geeco_set_error("debug_buffer: device error before rebuilding y1"); return GEECO_ERR_CUDA; }
      int rc = recompute_y1_bf16(cm, nullptr);
      if (rc) return rc;
      if (cudaDeviceSynchronize() != cudaSuccess) { geeco_set_error("debug_buffer: rebuilding y1 failed"); return GEECO_ERR_CUDA; }
    }
    *ptr = s[0] == 'y' ? L.y : L.g; *numel = L.act_elems; *dtype = act_dt;
    if (!*ptr) { geeco_set_error("debug_buffer: %s not allocated", name); return GEECO_ERR_INVALID; }
    return GEECO_OK;
  }
  const int T = c->T;
  const long long last = (long long)(T - 1) * N * Hl;
  struct { const char* n; float* p; long long cnt; } tab[] = {
      {"state", c->states, (long long)T * N * (c->xdim + Hl)}, {"gates", c->gates, (long long)T * N * 4 * Hl},
      {"c", c->c_seq + last, (long long)N * Hl}, {"m", c->m_seq + last, (long long)N * Hl}, {"fc1", c->fc1, (long long)N * cfg.dim_h_fc},
      {"heads", c->heads, (long long)N * c->NH}, {"dheads", c->dheads, (long long)N * c->NH},
      {"dfc1", c->dfc1, (long long)N * cfg.dim_h_fc}, {"dgates", c->dgates, (long long)T * N * 4 * Hl},
      {"dstate", c->dstates, (long long)T * N * (c->xdim + Hl)}, {"losses", c->losses, GEECO_NUM_LOSS_SLOTS}, {"sc", c->sc, 8},
      {"y8_f32", c->y8_f32, c->layers[7].act_elems}, {"g8_f32", c->g8_f32, c->layers[7].act_elems}};
  for (auto& t : tab)
    if (s == t.n) {
      if (!t.p) { geeco_set_error("debug_buffer: %s not allocated", name); return GEECO_ERR_INVALID; }
      *ptr = t.p; *numel = t.cnt; return GEECO_OK;
    }
  geeco_set_error("debug_buffer: unknown buffer '%s'", name);
  return GEECO_ERR_INVALID;
}
