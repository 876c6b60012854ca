// bf16 / tcgen05 path of the three conv encoders inside the model step, plus the stand-alone
// bf16 conv ops exported for parity tests and kernel benchmarks.
// Reference: conv_encoder, src/models/e2evmc/graph.py:61-117 (three scopes, :390-402).
#include "plan.cuh"
#include "conv_tc.cuh"

#include <stdlib.h>
#include <string.h>
#include <vector>

struct Bf16Layer {
  TcGeom fwd;                          // forward geometry (also the wgrad geometry)
  __nv_bfloat16* w_fwd[3];             // packed [Cout][Kpad] per encoder (contiguous when grouped)
  CUtensorMap fwd_map[3];
  int n_classes;
  TcGeom dg[4];
  int dg_taps[4][9];
  __nv_bfloat16* w_dg[4][3];
  CUtensorMap dg_map[4][3];
  // conv1 on pixel pairs: two output-column parity classes (see conv_tc.cu)
  bool pair;
  TcGeom pg[2];
  __nv_bfloat16* w_pair[2];
  CUtensorMap pair_map[2];
  // conv1 on pixel pairs inside the fused forward kernel: [G][64 rows (p, co)][64] (pack mode 5)
  __nv_bfloat16* w_c12pair = nullptr;
  CUtensorMap c12pair_map;
};

struct Bf16Plan {
  Bf16Layer L[8];
  float* partial;
  long long partial_cap;
  // weight-repack jobs in three tables (uploaded at the first repack after bind), launched where the step has room:
  //   0 EARLY  forward operands of conv1-conv3 (a few hundred KB): next to the pre-process kernel, joined before conv1
  //   1 LATE   forward operands of conv4-conv8: same place, joined before conv4 (~0.6 ms later)
  //   2 DGRAD  all data-gradient operands (60 % of the elements): forked after conv8's forward, runs next to the
  //            LSTM / heads tail (a dozen small kernels that leave most SMs idle), joined before the first data gradient
  // (a single table next to the pre-process kernel kept conv1 from starting: its CTAs need 29 K registers each and the
  //  repack's blocks refill every slot they free -- r02 timelines)
  PackJob* jobs_dev[3];
  std::vector<PackJob> jobs[3];
  long long jobs_total[3];
  bool jobs_uploaded;
  // tables 1 and 2 again as 32 x 32 tiles (pack_tiles_kernel): used from a table's second repack on, after the
  // element-wise kernel has written the zero padding of the packed matrices once
  PackTile* tiles_dev[3] = {nullptr, nullptr, nullptr};
  int tiles_cap[3] = {0, 0, 0}, ntiles[3] = {0, 0, 0};
  bool table_init[3] = {false, false, false};
  // the weight repack of a step does not depend on the batch: it runs on a side stream next to the rank-pooling
  // kernel (whose 16-CTA clusters leave a quarter of the SMs idle) and is joined before the first convolution
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join_late = nullptr, ev_fork_dg = nullptr, ev_join_dg = nullptr;
  bool forked = false, late_pending = false, dg_dirty = true, dg_pending = false;
};

// packed-K extent of a data-gradient class when the layer's output has `cout` channels (the planned geometry was
// built with Cout[0]; encoders of a non-grouped layer differ only in that number)
static int dgrad_kpad(const TcGeom& dg, int cout) {
  const int kt = dg.a_tma ? (cout + 63) / 64 * 64 : cout;
  return (dg.ntaps * kt + 63) / 64 * 64;
}
static int dgrad_kt(const TcGeom& dg, int cout) { return dg.a_tma ? (cout + 63) / 64 * 64 : cout; }

static void* carve(size_t* off, char* base, size_t bytes) {
  *off = (*off + 255) & ~(size_t)255;
  void* p = base ? base + *off : nullptr;
  *off += bytes;
  return p;
}

int plan_bf16(geeco_ctx* c, size_t* ws_off, char* ws_base) {
  if (!c->bf16_ws) c->bf16_ws = new Bf16Plan();
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  const int N = c->M, G = c->G;      // images per encoder group, encoder groups
  long long cap = 0;
  for (int l = 0; l < 8; ++l) {
    LayerPlan& L = c->layers[l];
    Bf16Layer& B = bp->L[l];
    const int groups = L.grouped ? G : 1;
    B.fwd = tc_fwd_geom(L.Hin, L.Hin, L.Cin_pad, L.Cout[0], L.stride, N, groups);
    // pixel-pair formulation of conv1 (16-byte copies, two parity classes): correct but measured slower than the
    // row-window producer on B200, kept as an opt-in experiment (GEECO_TC_CONV1_PAIR=1)
    B.pair = (l == 0 && L.Cin_pad == 4 && L.stride == 1 && L.Hin % 2 == 0 && L.grouped && getenv("GEECO_TC_CONV1_PAIR"));
    if (B.pair) {
      for (int par = 0; par < 2; ++par) B.pg[par] = tc_conv1pair_geom(L.Hin, L.Hin, L.Cout[0], N, groups, par);
      const long long need = 2 * tc_wgrad_partial_floats(B.pg[0], L.Cout[0]);
      if (need > cap) cap = need;
    }
    B.n_classes = 0;
    if (l > 0) {
      for (int py = 0; py < L.stride; ++py)
        for (int px = 0; px < L.stride; ++px) {
          TcGeom dg;
          int taps[9];
          if (!tc_dgrad_geom(L.Hin, L.Hin, L.Cin_real, L.Cout[0], L.stride, py, px, N, groups, &dg, taps)) continue;
          const int ci = B.n_classes++;
          B.dg[ci] = dg;
          memcpy(B.dg_taps[ci], taps, sizeof(taps));
          for (int e = 0; e < 3; ++e) B.w_dg[ci][e] = nullptr;
        }
    }
    for (int e = 0; e < (L.grouped ? 1 : G); ++e) {
      TcGeom g = B.fwd;
      const long long need = tc_wgrad_partial_floats(g, L.Cout[e]);
      if (need > cap) cap = need;
    }
  }
  // contiguous packed-weight storage per layer: [enc][rows][Kpad]
  for (int l = 0; l < 8; ++l) {
    LayerPlan& L = c->layers[l];
    Bf16Layer& B = bp->L[l];
    size_t tot = 0;
    for (int e = 0; e < G; ++e) tot += (size_t)L.Cout[e] * B.fwd.Kpad;
    __nv_bfloat16* base = (__nv_bfloat16*)carve(ws_off, ws_base, tot * 2);
    size_t o = 0;
    for (int e = 0; e < G; ++e) { B.w_fwd[e] = base ? base + o : nullptr; o += (size_t)L.Cout[e] * B.fwd.Kpad; }
    for (int par = 0; par < 2; ++par)
      B.w_pair[par] = B.pair ? (__nv_bfloat16*)carve(ws_off, ws_base, (size_t)G * L.Cout[0] * 64 * 2) : nullptr;
    B.w_c12pair = (l == 0 && L.grouped && L.Cin_pad == 4 && L.Cout[0] == 32) ? (__nv_bfloat16*)carve(ws_off, ws_base, (size_t)G * 64 * 64 * 2) : nullptr;
    for (int ci = 0; ci < B.n_classes; ++ci) {
      size_t t2 = 0;
      for (int e = 0; e < G; ++e) t2 += (size_t)L.Cin_real * dgrad_kpad(B.dg[ci], L.Cout[e]);
      __nv_bfloat16* b2 = (__nv_bfloat16*)carve(ws_off, ws_base, t2 * 2);
      size_t o2 = 0;
      for (int e = 0; e < G; ++e) {
        B.w_dg[ci][e] = b2 ? b2 + o2 : nullptr;
        o2 += (size_t)L.Cin_real * dgrad_kpad(B.dg[ci], L.Cout[e]);
      }
    }
  }
  if (tc_bwd21_partial_floats() > cap) cap = tc_bwd21_partial_floats();
  if (tc_wgrad2_partial_floats() > cap) cap = tc_wgrad2_partial_floats();
  bp->partial_cap = cap;
  bp->partial = (float*)carve(ws_off, ws_base, (size_t)cap * sizeof(float));
  bp->jobs_dev[0] = (PackJob*)carve(ws_off, ws_base, 64 * sizeof(PackJob));
  bp->jobs_dev[1] = (PackJob*)carve(ws_off, ws_base, 64 * sizeof(PackJob));
  bp->jobs_dev[2] = (PackJob*)carve(ws_off, ws_base, 64 * sizeof(PackJob));
  bp->jobs_uploaded = false;
  for (int t = 1; t < 3; ++t) {
    long long n = 0;
    for (int l = (t == 1 ? 3 : 1); l < 8; ++l) {
      const LayerPlan& L = c->layers[l];
      for (int e = 0; e < G; ++e) n += 9ll * ((L.Cin_real + 31) / 32) * ((L.Cout[L.grouped ? 0 : e] + 31) / 32);
    }
    bp->tiles_cap[t] = (int)n;
    bp->tiles_dev[t] = (PackTile*)carve(ws_off, ws_base, (size_t)n * sizeof(PackTile));
    bp->ntiles[t] = 0;
    bp->table_init[t] = false;
  }
  bp->table_init[0] = false;
  if (!ws_base) return GEECO_OK;
  if (!bp->side) {
    // lowest priority: the repack's many short blocks must not hold up the CTAs of the step's own kernels (conv1
    // started 25 us late behind the tail of the late repack, r02 timeline)
    int prio_lo = 0, prio_hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&bp->side, cudaStreamNonBlocking, prio_lo));
    CUDA_TRY(cudaEventCreateWithFlags(&bp->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&bp->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&bp->ev_join_late, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&bp->ev_fork_dg, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&bp->ev_join_dg, cudaEventDisableTiming));
  }
  // tensor maps (need the real addresses)
  for (int l = 0; l < 8; ++l) {
    LayerPlan& L = c->layers[l];
    Bf16Layer& B = bp->L[l];
    if (B.pair) {
      for (int par = 0; par < 2; ++par) {
        int rc = make_weight_tensor_map(&B.pair_map[par], B.w_pair[par], (long long)G * L.Cout[0], 64, L.Cout[0]);
        if (rc) return rc;
      }
    }
    if (B.w_c12pair) {
      int rc = make_weight_tensor_map(&B.c12pair_map, B.w_c12pair, (long long)G * 64, 64, 64);
      if (rc) return rc;
    }
    if (L.grouped) {
      int rc = make_weight_tensor_map(&B.fwd_map[0], B.w_fwd[0], (long long)G * L.Cout[0], B.fwd.Kpad, L.Cout[0]);
      if (rc) return rc;
      for (int ci = 0; ci < B.n_classes; ++ci) {
        rc = make_weight_tensor_map(&B.dg_map[ci][0], B.w_dg[ci][0], (long long)G * L.Cin_real, B.dg[ci].Kpad, L.Cin_real);
        if (rc) return rc;
      }
    } else {
      for (int e = 0; e < G; ++e) {
        int rc = make_weight_tensor_map(&B.fwd_map[e], B.w_fwd[e], L.Cout[e], B.fwd.Kpad, L.Cout[e]);
        if (rc) return rc;
        for (int ci = 0; ci < B.n_classes; ++ci) {
          const int Kp = dgrad_kpad(B.dg[ci], L.Cout[e]);
          rc = make_weight_tensor_map(&B.dg_map[ci][e], B.w_dg[ci][e], L.Cin_real, Kp, L.Cin_real);
          if (rc) return rc;
        }
      }
    }
  }
  return GEECO_OK;
}

void free_bf16(geeco_ctx* c) {
  if (c->bf16_ws) {
    Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
    if (bp->side) { cudaStreamSynchronize(bp->side); cudaStreamDestroy(bp->side); }
    if (bp->ev_fork) cudaEventDestroy(bp->ev_fork);
    if (bp->ev_join) cudaEventDestroy(bp->ev_join);
    if (bp->ev_join_late) cudaEventDestroy(bp->ev_join_late);
    if (bp->ev_fork_dg) cudaEventDestroy(bp->ev_fork_dg);
    if (bp->ev_join_dg) cudaEventDestroy(bp->ev_join_dg);
    delete bp;
    c->bf16_ws = nullptr;
  }
}

static const int kAllTaps[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};

static void add_job(Bf16Plan* bp, int table, const float* W, __nv_bfloat16* out, int mode, int groups, long long wstride, int Cin,
                    int Cout, int Cs, int ntaps, const int* taps, int rows, int Kpad, int Kt, const float* bias = nullptr,
                    long long bstride = 0, int bias_col = -1) {
  PackJob j;
  memset(&j, 0, sizeof(j));
  j.W = W; j.out = out; j.w_group_stride = wstride; j.mode = mode; j.groups = groups; j.Cin = Cin; j.Cout = Cout;
  j.Cs = Cs; j.ntaps = ntaps; j.rows = rows; j.Kpad = Kpad; j.Kt = Kt;
  j.bias = bias; j.b_group_stride = bstride; j.bias_col = bias ? bias_col : -1;
  for (int i = 0; i < ntaps && i < 9; ++i) j.taps[i] = taps[i];
  j.start = bp->jobs_total[table];
  j.total = (long long)groups * rows * Kpad;
  bp->jobs_total[table] = (j.start + j.total + PACK_CHUNK - 1) / PACK_CHUNK * PACK_CHUNK;
  bp->jobs[table].push_back(j);
}

// fp32 master weights -> packed bf16 operands of every conv layer (forward + the data-gradient classes): one launch
// per job table.  which: 0 / 1 / 2 = one table, 3 = the two forward tables, 4 = all three
static int repack_weights(geeco_ctx* c, cudaStream_t st, int which = 4, int max_blocks = 0) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  const int G = c->G;
  if (!bp->jobs_uploaded) {
    for (int t = 0; t < 3; ++t) { bp->jobs[t].clear(); bp->jobs_total[t] = 0; }
    for (int l = 0; l < 8; ++l) {
      LayerPlan& L = c->layers[l];
      Bf16Layer& B = bp->L[l];
      const long long wstride = w_group_stride(c, L);
      const int ne = L.grouped ? 1 : G;
      const int fwd_table = l < 3 ? 0 : 1;
      for (int e = 0; e < ne; ++e) {
        const int groups = L.grouped ? G : 1;
        const float* W = c->theta + c->params[L.p_w[e]].offset;
        if (B.pair) {
          for (int par = 0; par < 2; ++par)
            add_job(bp, fwd_table, W, B.w_pair[par], 2 + par, groups, wstride, L.Cin_real, L.Cout[e], 8, 6, kAllTaps, L.Cout[e], 64, 0);
        } else {
          const long long bstride = b_group_stride(c, L);
          add_job(bp, fwd_table, W, B.w_fwd[e], B.fwd.wpack, groups, wstride, L.Cin_real, L.Cout[e], L.Cin_pad, 9, kAllTaps, L.Cout[e],
                  B.fwd.Kpad, B.fwd.Kt, B.fwd.bias_in_k ? c->theta + c->params[L.p_b[e]].offset : nullptr, bstride, B.fwd.Ktot);
          if (B.w_c12pair && e == 0)
            add_job(bp, fwd_table, W, B.w_c12pair, 5, groups, wstride, L.Cin_real, L.Cout[e], L.Cin_pad, 9, kAllTaps, 64, 64, 0,
                    c->theta + c->params[L.p_b[e]].offset, bstride, 48);
        }
        for (int ci = 0; ci < B.n_classes; ++ci) {
          const int Kp = dgrad_kpad(B.dg[ci], L.Cout[e]);
          add_job(bp, 2, W, B.w_dg[ci][e], 1, groups, wstride, L.Cin_real, L.Cout[e], L.Cout[e], B.dg[ci].ntaps, B.dg_taps[ci],
                  L.Cin_real, Kp, dgrad_kt(B.dg[ci], L.Cout[e]));
        }
      }
    }
    for (int t = 0; t < 3; ++t) {
      if (bp->jobs[t].size() > 64) { geeco_set_error("repack: %zu jobs > 64", bp->jobs[t].size()); return GEECO_ERR_INVALID; }
      for (const PackJob& pj : bp->jobs[t])
        if (pj.total >= (1ll << 31)) { geeco_set_error("repack: a job of %lld elements needs 64-bit indexing", pj.total); return GEECO_ERR_INVALID; }
      if (!bp->jobs[t].empty())
        CUDA_TRY(cudaMemcpyAsync(bp->jobs_dev[t], bp->jobs[t].data(), bp->jobs[t].size() * sizeof(PackJob), cudaMemcpyHostToDevice, st));
    }
    const bool no_tiles = getenv("GEECO_PACK_NO_TILES") != nullptr;     // read per context: the parity test toggles it
    std::vector<PackTile> tiles;
    for (int t = 1; t < 3; ++t) {
      bp->ntiles[t] = 0;
      bool plain = !no_tiles && bp->tiles_dev[t] != nullptr;
      for (const PackJob& pj : bp->jobs[t]) plain = plain && pack_job_is_plain(pj);
      if (!plain) continue;
      tiles.clear();
      for (const PackJob& pj : bp->jobs[t]) pack_job_tiles(pj, &tiles);
      if ((int)tiles.size() > bp->tiles_cap[t]) continue;         // keeps the element-wise kernel
      CUDA_TRY(cudaMemcpyAsync(bp->tiles_dev[t], tiles.data(), tiles.size() * sizeof(PackTile), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaStreamSynchronize(st));                         // `tiles` is reused / freed
      bp->ntiles[t] = (int)tiles.size();
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    bp->jobs_uploaded = true;
  }
  for (int t = 0; t < 3; ++t) {
    const bool sel = which == t || (which == 3 && t < 2) || which == 4;
    if (!sel || bp->jobs[t].empty()) continue;
    int rc;
    if (bp->table_init[t] && bp->ntiles[t] > 0) {
      rc = launch_pack_tiles(bp->tiles_dev[t], bp->ntiles[t], st);
    } else {
      rc = launch_pack_weights_batched(bp->jobs_dev[t], (int)bp->jobs[t].size(), bp->jobs_total[t], st, max_blocks);
      bp->table_init[t] = true;
    }
    if (rc) return rc;
  }
  if (which == 1 || which == 3 || which == 4) c->weights_dirty = false;     // forward operands are current
  if (which == 2 || which == 4) bp->dg_dirty = false;                        // data-gradient operands are current
  return GEECO_OK;
}

// Starts the weight repack (if the weights changed) on the side stream, ordered after everything already enqueued on
// `st` (the previous Adam update); repack_join makes `st` wait for it.  Between the two calls the caller enqueues work
// that does not read the packed weights (rank pooling).
int repack_fork_bf16(geeco_ctx* c, cudaStream_t st) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  static const bool off = getenv("GEECO_NO_OVERLAP") != nullptr;
  if (!bp || !c->weights_dirty || !bp->side || !bp->jobs_uploaded || off) return GEECO_OK;   // first step: inline repack
  bp->dg_dirty = true;                                  // the weights changed: every table is stale
  CUDA_TRY(cudaEventRecord(bp->ev_fork, st));
  CUDA_TRY(cudaStreamWaitEvent(bp->side, bp->ev_fork, 0));
  geeco_pdl_suspend(1);                                 // plain launches on the side stream (events between them)
  int rc = repack_weights(c, bp->side, 0);
  if (rc) { geeco_pdl_suspend(0); return rc; }
  CUDA_TRY(cudaEventRecord(bp->ev_join, bp->side));
  rc = repack_weights(c, bp->side, 1);
  geeco_pdl_suspend(0);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(bp->ev_join_late, bp->side));
  bp->forked = true;
  bp->late_pending = true;
  return GEECO_OK;
}
// `st` waits for the EARLY table (conv1-conv3 forward operands); the LATE table is joined inside encoders_fwd_bf16
// right before conv4, ~0.6 ms of forward work later
int repack_join_bf16(geeco_ctx* c, cudaStream_t st) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  if (!bp || !bp->forked) return GEECO_OK;
  CUDA_TRY(cudaStreamWaitEvent(st, bp->ev_join, 0));
  bp->forked = false;
  return GEECO_OK;
}
// data-gradient operands: forked behind the forward of conv8 (training contexts), joined before the first data gradient
static int repack_dgrad_fork(geeco_ctx* c, cudaStream_t st) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  if (!c->cfg.training || !bp->dg_dirty || bp->dg_pending) return GEECO_OK;
  static const bool off = getenv("GEECO_NO_OVERLAP") != nullptr;
  if (off || !bp->side) return repack_weights(c, st, 2);
  CUDA_TRY(cudaEventRecord(bp->ev_fork_dg, st));
  CUDA_TRY(cudaStreamWaitEvent(bp->side, bp->ev_fork_dg, 0));
  geeco_pdl_suspend(1);
  // few resident blocks (grid-stride): the tail's small kernels must find free SM slots while this runs next to them
  static const int dg_blocks = getenv("GEECO_PACK_DG_BLOCKS") ? atoi(getenv("GEECO_PACK_DG_BLOCKS")) : 0;
  int rc = repack_weights(c, bp->side, 2, dg_blocks);
  geeco_pdl_suspend(0);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(bp->ev_join_dg, bp->side));
  bp->dg_pending = true;
  return GEECO_OK;
}

static bool conv12_applies(geeco_ctx* c, Bf16Plan* bp) {
  LayerPlan& L = c->layers[0];
  LayerPlan& L1 = c->layers[1];
  Bf16Layer& B = bp->L[0];
  return !(B.pair || !L.grouped || !L1.grouped || bp->L[1].fwd.rows != 2 || bp->L[1].fwd.wpack != 4 ||
           !tc_conv12_supported(L.Hin, L.Hin, L.Cin_pad, L.Cout[0], L1.Cout[0], L.stride, L1.stride, B.fwd));
}

// conv2's weight gradient on whole y1 rows (conv2_wgrad_fused.cu) when the two layers have the shape it is written for
static bool wgrad2_active(geeco_ctx* c, Bf16Plan* bp) {
  LayerPlan& L = c->layers[0];
  LayerPlan& L1 = c->layers[1];
  return c->cfg.training && conv12_applies(c, bp) && L.mbits != nullptr && L1.Cin_real == L.Cout[0] &&
         tc_wgrad2_supported(L.Hin, L.Hin, L.Cin_pad, L.Cout[0], L1.Cout[0], L.stride, L1.stride, bp->L[0].fwd);
}

// GEECO_WG2_RECOMPUTE=1: that kernel rebuilds the y1 rows from x0 instead of loading them; the data gradient takes its
// ReLU mask from the 1-bit copy, so no kernel reads y1 from HBM then and the fused forward does not store it (measured
// slower than loading: 248 + 301 us vs 275 + ~200 us for the forward + this gradient; it frees 805 MB and 1.6 GB of DRAM
// traffic per step)
static bool wgrad2_recompute(geeco_ctx* c, Bf16Plan* bp) {
  return wgrad2_active(c, bp) && getenv("GEECO_WG2_RECOMPUTE") != nullptr && getenv("GEECO_KEEP_Y1") == nullptr;
}

// conv1 -> conv2 as one kernel when the two layers have the shape it is written for; returns 1 if it ran, 0 if the
// separate kernels have to, < 0 on error
static int try_conv12(geeco_ctx* c, Bf16Plan* bp, cudaStream_t st) {
  LayerPlan& L = c->layers[0];
  LayerPlan& L1 = c->layers[1];
  Bf16Layer& B = bp->L[0];
  if (!conv12_applies(c, bp)) return 0;
  // y1 stays on chip between the layers; inference does not write it at all, training only when a kernel will read it
  const bool store_y1 = c->cfg.training && !wgrad2_recompute(c, bp);
  // conv1 on pixel pairs: opt-in experiment (GEECO_CONV12_PAIR=1).  Half the im2col copies and a third fewer operand
  // reads, but measured SLOWER in training on B200 (322 vs 266 us at batch 64; inference 2.4 % faster;
  // profiles/r02_ncu_notes.md).  Never together with the y1-rebuilding weight
  // gradient: the two forms sum the same products in a different order, and the rebuilt rows must be the forward's.
  const bool pair = B.w_c12pair != nullptr && !wgrad2_recompute(c, bp) && getenv("GEECO_CONV12_PAIR") != nullptr;
  int rc = launch_tc_conv12((const __nv_bfloat16*)c->x0, &B.fwd_map[0], &bp->L[1].fwd_map[0],
                            c->theta + c->params[L1.p_b[0]].offset, b_group_stride(c, L1),
                            store_y1 ? (__nv_bfloat16*)L.y : nullptr, (unsigned short*)L.mbits, (__nv_bfloat16*)L1.y,
                            (unsigned short*)L1.mbits, c->G, c->M, st, pair ? &B.c12pair_map : nullptr);
  c->y1_stale = c->cfg.training && !store_y1;
  return rc ? -rc : 1;
}

// debug (geeco_debug_buffer("y1")): conv1's activation of the last forward, rebuilt by the stand-alone conv1 kernel
// (bit-identical to what the fused kernels compute, tests/test_gpu_fused12.py)
int recompute_y1_bf16(geeco_ctx* c, cudaStream_t st) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  if (!bp || !c->y1_stale) return GEECO_OK;
  LayerPlan& L = c->layers[0];
  TcGeom g = bp->L[0].fwd;
  g.bias_group_stride = b_group_stride(c, L);
  geeco_pdl_suspend(1);
  int rc = launch_tc_nn(g, &bp->L[0].fwd_map[0], (const __nv_bfloat16*)c->x0, c->theta + c->params[L.p_b[0]].offset, nullptr,
                        (__nv_bfloat16*)L.y, nullptr, TC_EPI_BIAS_RELU, 0, st, nullptr, bp->L[0].w_fwd[0]);
  geeco_pdl_suspend(0);
  if (rc) return rc;
  c->y1_stale = false;
  return GEECO_OK;
}

static int try_wgrad2(geeco_ctx* c, Bf16Plan* bp, cudaStream_t st) {
  if (!wgrad2_active(c, bp)) return 0;
  LayerPlan& L1 = c->layers[1];
  // y1 rows are loaded unless the forward kept them on chip (then they are rebuilt from x0)
  const __nv_bfloat16* y1 = c->y1_stale ? nullptr : (const __nv_bfloat16*)c->layers[0].y;
  int rc = launch_tc_wgrad2((const __nv_bfloat16*)c->x0, &bp->L[0].fwd_map[0], y1, (const __nv_bfloat16*)L1.g, bp->partial,
                            bp->partial_cap, c->grad + c->params[L1.p_w[0]].offset, c->grad + c->params[L1.p_b[0]].offset,
                            w_group_stride(c, L1), b_group_stride(c, L1), c->G, c->M, st);
  return rc ? -rc : 1;
}

// conv2 data gradient -> conv1 weight gradient as one kernel when the two layers have the shape it is written for (G1 is
// then never written); returns 1 if it ran, 0 if the separate kernels have to, < 0 on error
static int try_bwd21(geeco_ctx* c, Bf16Plan* bp, cudaStream_t st) {
  LayerPlan& L0 = c->layers[0];
  LayerPlan& L1 = c->layers[1];
  Bf16Layer& B0 = bp->L[0];
  Bf16Layer& B1 = bp->L[1];
  if (B0.pair || !L0.grouped || !L1.grouped || !L0.mbits || L0.Cin_real > 4 ||
      !tc_bwd21_supported(L0.Hin, L0.Hin, L0.Cin_pad, L0.Cout[0], L1.Cout[0], L0.stride, L1.stride, B1.dg, B1.n_classes))
    return 0;
  const CUtensorMap* dmaps[4];
  for (int ci = 0; ci < 4; ++ci) dmaps[ci] = &B1.dg_map[ci][0];
  int rc = launch_tc_bwd21((const __nv_bfloat16*)L1.g, B1.dg, dmaps, (const unsigned short*)L0.mbits, (const __nv_bfloat16*)c->x0,
                           bp->partial, bp->partial_cap, c->grad + c->params[L0.p_w[0]].offset,
                           c->grad + c->params[L0.p_b[0]].offset, L0.Cin_real, w_group_stride(c, L0), b_group_stride(c, L0),
                           c->G, c->M, st);
  return rc ? -rc : 1;
}

// profiling entry (geeco_profile_kernel): one kernel of the bf16 step on the buffers the last step left behind
int profile_kernel_bf16(geeco_ctx* c, const char* name, cudaStream_t st) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  if (!bp) { geeco_set_error("profile_kernel: bf16 plan missing"); return GEECO_ERR_STATE; }
  if (!strcmp(name, "conv12")) {
    const int r = try_conv12(c, bp, st);
    if (r == 0) { geeco_set_error("profile_kernel: the fused conv1->conv2 kernel does not cover this configuration"); return GEECO_ERR_INVALID; }
    return r < 0 ? -r : GEECO_OK;
  }
  if (!strcmp(name, "wgrad2")) {
    const int r = try_wgrad2(c, bp, st);
    if (r == 0) { geeco_set_error("profile_kernel: the fused conv2 weight gradient does not cover this configuration"); return GEECO_ERR_INVALID; }
    return r < 0 ? -r : GEECO_OK;
  }
  if (!strcmp(name, "bwd21")) {
    const int r = try_bwd21(c, bp, st);
    if (r == 0) { geeco_set_error("profile_kernel: the fused conv2-dgrad -> conv1-wgrad kernel does not cover this configuration"); return GEECO_ERR_INVALID; }
    return r < 0 ? -r : GEECO_OK;
  }
  geeco_set_error("profile_kernel: unknown kernel '%s'", name);
  return GEECO_ERR_INVALID;
}

int encoders_fwd_bf16(geeco_ctx* c, cudaStream_t st) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  if (!bp) { geeco_set_error("bf16 plan missing"); return GEECO_ERR_STATE; }
  if (c->weights_dirty) {
    bp->dg_dirty = true;
    int rc = repack_weights(c, st, 3);
    if (rc) return rc;
  }
  const int N = c->M, G = c->G;
  const __nv_bfloat16* src = (const __nv_bfloat16*)c->x0;
  for (int l = 0; l < 8; ++l) {
    LayerPlan& L = c->layers[l];
    Bf16Layer& B = bp->L[l];
    if (l == 3 && bp->late_pending) {
      CUDA_TRY(cudaStreamWaitEvent(st, bp->ev_join_late, 0));
      bp->late_pending = false;
    }
    if (l == 0) {
      const int r = try_conv12(c, bp, st);
      if (r < 0) return -r;
      if (r > 0) { src = (const __nv_bfloat16*)c->layers[1].y; l = 1; continue; }
    }
    if (B.pair) {
      TcGeom pg[2] = {B.pg[0], B.pg[1]};
      pg[0].bias_group_stride = pg[1].bias_group_stride = b_group_stride(c, L);
      const CUtensorMap* pm[2] = {&B.pair_map[0], &B.pair_map[1]};
      int rc = launch_tc_nn_multi(pg, pm, 2, src, c->theta + c->params[L.p_b[0]].offset, nullptr, (__nv_bfloat16*)L.y, nullptr,
                                  TC_EPI_BIAS_RELU, 0, st, (unsigned short*)L.mbits);
      if (rc) return rc;
    } else if (L.grouped) {
      TcGeom g = B.fwd;
      g.bias_group_stride = b_group_stride(c, L);
      int rc = launch_tc_nn(g, &B.fwd_map[0], src, c->theta + c->params[L.p_b[0]].offset, nullptr,
                            (__nv_bfloat16*)L.y, l == 7 ? c->y8_f32 : nullptr, TC_EPI_BIAS_RELU, 0, st,
                            (unsigned short*)L.mbits, B.w_fwd[0]);
      if (rc) return rc;
    } else {
      for (int e = 0; e < G; ++e) {
        TcGeom g = tc_fwd_geom(L.Hin, L.Hin, L.Cin_pad, L.Cout[e], L.stride, N, 1);
        const __nv_bfloat16* s = src + (long long)e * N * L.Hin * L.Hin * L.Cin_pad;
        int rc = launch_tc_nn(g, &B.fwd_map[e], s, c->theta + c->params[L.p_b[e]].offset, nullptr,
                              (__nv_bfloat16*)L.y + L.act_off[e], l == 7 ? c->y8_f32 + L.act_off[e] : nullptr,
                              TC_EPI_BIAS_RELU, 0, st);
        if (rc) return rc;
      }
    }
    src = (const __nv_bfloat16*)L.y;
  }
  return repack_dgrad_fork(c, st);
}

int encoders_bwd_bf16(geeco_ctx* c, int lhi, int llo, cudaStream_t st) {
  Bf16Plan* bp = (Bf16Plan*)c->bf16_ws;
  if (!bp) { geeco_set_error("bf16 plan missing"); return GEECO_ERR_STATE; }
  if (bp->dg_pending) {
    CUDA_TRY(cudaStreamWaitEvent(st, bp->ev_join_dg, 0));
    bp->dg_pending = false;
  } else if (bp->dg_dirty) {
    int rc = repack_weights(c, st, 2);
    if (rc) return rc;
  }
  const int N = c->M, G = c->G;
  if (lhi == 7 && !c->g8_bf16_ready) {
    int rc = launch_f32_to_bf16(c->g8_f32, (__nv_bfloat16*)c->layers[7].g, c->layers[7].act_elems, st);
    if (rc) return rc;
  }
  for (int l = lhi; l >= llo; --l) {
    LayerPlan& L = c->layers[l];
    Bf16Layer& B = bp->L[l];
    const __nv_bfloat16* xin = l == 0 ? (const __nv_bfloat16*)c->x0 : (const __nv_bfloat16*)c->layers[l - 1].y;
    const long long wstride = w_group_stride(c, L);
    const long long bstride = b_group_stride(c, L);
    if (B.pair) {
      // conv1 weight gradient on pixel pairs: one partial GEMM per output-column parity, one joint reduction
      const long long half = bp->partial_cap / 2;
      int splits = 0, mrows = 0;
      int rc = launch_tc_wgrad_partial(B.pg[0], L.Cout[0], xin, (const __nv_bfloat16*)L.g, bp->partial, half, 1, &splits, &mrows, st);
      if (rc) return rc;
      rc = launch_tc_wgrad_partial(B.pg[1], L.Cout[0], xin, (const __nv_bfloat16*)L.g, bp->partial + half, half, 1, &splits, &mrows, st);
      if (rc) return rc;
      rc = launch_conv1pair_reduce(bp->partial, bp->partial + half, c->grad + c->params[L.p_w[0]].offset,
                                   c->grad + c->params[L.p_b[0]].offset, splits, G, mrows, 64, L.Cin_real, L.Cout[0], wstride,
                                   bstride, st);
      if (rc) return rc;
      continue;
    }
    const int ne = L.grouped ? 1 : G;
    for (int e = 0; e < ne; ++e) {
      const int groups = L.grouped ? G : 1;
      const long long in_off = L.grouped ? 0 : (long long)e * N * L.Hin * L.Hin * L.Cin_pad;
      const __nv_bfloat16* gy = (const __nv_bfloat16*)L.g + (L.grouped ? 0 : L.act_off[e]);
      TcGeom g = L.grouped ? B.fwd : tc_fwd_geom(L.Hin, L.Hin, L.Cin_pad, L.Cout[e], L.stride, N, 1);
      int rc = 0;
      // conv2: y1 is recomputed on chip instead of read (it was not stored)
      const int fused_wg = l == 1 ? try_wgrad2(c, bp, st) : 0;
      if (fused_wg < 0) return -fused_wg;
      if (fused_wg == 0)
        rc = launch_tc_wgrad(g, L.Cout[e], L.Cin_real, xin + in_off, gy, c->grad + c->params[L.p_w[e]].offset,
                             c->grad + c->params[L.p_b[e]].offset, bp->partial, bp->partial_cap, wstride, bstride, st);
      if (rc) return rc;
      if (l == 0) continue;
      if (l == 1 && llo == 0) {
        // conv2's data gradient feeds conv1's weight gradient on chip; layer 0 is then done
        const int r = try_bwd21(c, bp, st);
        if (r < 0) return -r;
        if (r > 0) { llo = 1; continue; }
      }
      {
        TcGeom dgs[4];
        const CUtensorMap* dmaps[4];
        const __nv_bfloat16* dptrs[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int ci = 0; ci < B.n_classes; ++ci) {
          dgs[ci] = B.dg[ci];
          dptrs[ci] = B.w_dg[ci][e];
          if (!L.grouped) {
            int taps[9];
            tc_dgrad_geom(L.Hin, L.Hin, L.Cin_real, L.Cout[e], L.stride, B.dg[ci].dy0, B.dg[ci].dx0, N, 1, &dgs[ci], taps);
          }
          dmaps[ci] = &B.dg_map[ci][e];
        }
        (void)groups;
        // ReLU mask of y_{l-1}: the bit mask its forward wrote (grouped layers), else the activation itself
        const LayerPlan& Lp = c->layers[l - 1];
        const bool bits = L.grouped && Lp.mbits != nullptr;
        rc = launch_tc_nn_multi(dgs, dmaps, B.n_classes, gy, nullptr,
                                bits ? (const __nv_bfloat16*)Lp.mbits : xin + in_off,
                                (__nv_bfloat16*)Lp.g + in_off, nullptr, bits ? TC_EPI_MASKBITS : TC_EPI_MASK, 0, st, nullptr,
                                L.grouped ? dptrs : nullptr);
        if (rc) return rc;
      }
    }
  }
  return GEECO_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone bf16 conv ops (C-ABI, see include/geeco_b200.h)
// ------------------------------------------------------------------------------------------------
static size_t conv_bf16_scratch(int N, int H, int W, int Cin, int Cout, int stride, size_t* o_fwd, size_t o_dg[4],
                                size_t* o_part, long long* part_floats) {
  size_t off = 0;
  TcGeom g = tc_fwd_geom(H, W, Cin, Cout, stride, N, 1);
  auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; size_t r = off; off += bytes; return r; };
  *o_fwd = take((size_t)Cout * g.Kpad * 2);
  int ci = 0;
  for (int py = 0; py < stride; ++py)
    for (int px = 0; px < stride; ++px) {
      TcGeom dg; int taps[9];
      if (!tc_dgrad_geom(H, W, Cin, Cout, stride, py, px, N, 1, &dg, taps)) continue;
      o_dg[ci++] = take((size_t)Cin * dg.Kpad * 2);
    }
  *part_floats = tc_wgrad_partial_floats(g, Cout);
  if (Cin == 4 && stride == 1 && W % 2 == 0) {
    TcGeom pg = tc_conv1pair_geom(H, W, Cout, N, 1, 0);
    const long long need = 2 * tc_wgrad_partial_floats(pg, Cout);
    if (need > *part_floats) *part_floats = need;
  }
  *o_part = take((size_t)*part_floats * 4);
  take((size_t)2 * Cout * 64 * 2 + 512);      // conv1 pixel-pair packed weights (two classes)
  return off + 256;
}

extern "C" int64_t geeco_conv2d_bf16_scratch_bytes(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                                   int32_t stride) {
  size_t a, b[4], p; long long pf;
  return (int64_t)conv_bf16_scratch(N, H, W, Cin, Cout, stride, &a, b, &p, &pf);
}

extern "C" int geeco_conv2d_same_bf16(const void* x, const float* w, const float* b, void* y, float* y_f32,
                                      void* scratch, int64_t scratch_bytes, int32_t N, int32_t H, int32_t W,
                                      int32_t Cin, int32_t Cw, int32_t Cout, int32_t stride, int32_t relu, void* stream) {
  if (!x || !w || !scratch || (!y && !y_f32)) { geeco_set_error("conv2d_bf16: NULL tensor"); return GEECO_ERR_INVALID; }
  if (H != W) { geeco_set_error("conv2d_bf16: square inputs only"); return GEECO_ERR_INVALID; }
  size_t o_fwd, o_dg[4], o_part; long long pf;
  if ((int64_t)conv_bf16_scratch(N, H, W, Cin, Cout, stride, &o_fwd, o_dg, &o_part, &pf) > scratch_bytes) {
    geeco_set_error("conv2d_bf16: scratch too small"); return GEECO_ERR_WORKSPACE;
  }
  char* base = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 4 && stride == 1 && W % 2 == 0 && getenv("GEECO_TC_CONV1_PAIR")) {
    // conv1 on pixel pairs: two parity classes in one launch
    __nv_bfloat16* wpp = (__nv_bfloat16*)(base + ((o_part + (size_t)pf * 4 + 255) & ~(size_t)255));
    TcGeom pg[2]; CUtensorMap pm[2]; const CUtensorMap* pmp[2];
    for (int par = 0; par < 2; ++par) {
      pg[par] = tc_conv1pair_geom(H, W, Cout, N, 1, par);
      int rc = launch_pack_weights(w, wpp + (size_t)par * Cout * 64, 2 + par, 1, 0, Cw, Cout, 8, 6, kAllTaps, Cout, 64, 0, st);
      if (rc) return rc;
      rc = make_weight_tensor_map(&pm[par], wpp + (size_t)par * Cout * 64, Cout, 64, Cout);
      if (rc) return rc;
      pmp[par] = &pm[par];
    }
    return launch_tc_nn_multi(pg, pmp, 2, (const __nv_bfloat16*)x, b, nullptr, (__nv_bfloat16*)y, y_f32,
                              b ? (relu ? TC_EPI_BIAS_RELU : TC_EPI_BIAS) : TC_EPI_STORE, 0, st);
  }
  TcGeom g = tc_fwd_geom(H, W, Cin, Cout, stride, N, 1);
  __nv_bfloat16* wp = (__nv_bfloat16*)(base + o_fwd);
  int rc = launch_pack_weights(w, wp, g.wpack, 1, 0, Cw, Cout, Cin, 9, kAllTaps, Cout, g.Kpad, g.Kt, st,
                               g.bias_in_k ? b : nullptr, g.Ktot);
  if (rc) return rc;
  CUtensorMap map;
  rc = make_weight_tensor_map(&map, wp, Cout, g.Kpad, Cout);
  if (rc) return rc;
  return launch_tc_nn(g, &map, (const __nv_bfloat16*)x, b, nullptr, (__nv_bfloat16*)y, y_f32,
                      b ? (relu ? TC_EPI_BIAS_RELU : TC_EPI_BIAS) : TC_EPI_STORE, 0, st);
}

static int conv2d_bwd_bf16_impl(const void* x, const float* w, const void* dy_pre, const void* relu_mask_x, bool mask_is_bits,
                                float* dw, float* db, void* dx, void* scratch, int64_t scratch_bytes, int32_t N, int32_t H,
                                int32_t W, int32_t Cin, int32_t Cw, int32_t Cout, int32_t stride, void* stream);

extern "C" int geeco_conv2d_same_bwd_bf16(const void* x, const float* w, const void* dy_pre, const void* relu_mask_x,
                                          float* dw, float* db, void* dx, void* scratch, int64_t scratch_bytes,
                                          int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cw, int32_t Cout,
                                          int32_t stride, void* stream) {
  return conv2d_bwd_bf16_impl(x, w, dy_pre, relu_mask_x, false, dw, db, dx, scratch, scratch_bytes, N, H, W, Cin, Cw, Cout,
                              stride, stream);
}

extern "C" int geeco_conv2d_same_bwd_bf16_bits(const void* x, const float* w, const void* dy_pre, const void* relu_mask_bits,
                                               float* dw, float* db, void* dx, void* scratch, int64_t scratch_bytes,
                                               int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cw, int32_t Cout,
                                               int32_t stride, void* stream) {
  if (!relu_mask_bits) { geeco_set_error("conv2d_bwd_bf16_bits: NULL mask"); return GEECO_ERR_INVALID; }
  if (Cin % 16) { geeco_set_error("conv2d_bwd_bf16_bits: Cin=%d must be a multiple of 16", Cin); return GEECO_ERR_INVALID; }
  return conv2d_bwd_bf16_impl(x, w, dy_pre, relu_mask_bits, true, dw, db, dx, scratch, scratch_bytes, N, H, W, Cin, Cw, Cout,
                              stride, stream);
}

extern "C" int geeco_relu_mask_bits(const void* y, void* bits, int64_t pixels, int32_t C, void* stream) {
  if (!y || !bits || pixels < 0 || C <= 0 || C % 16) { geeco_set_error("relu_mask_bits: bad arguments"); return GEECO_ERR_INVALID; }
  return launch_relu_mask_bits((const __nv_bfloat16*)y, (unsigned short*)bits, (long long)pixels * (C / 16), (cudaStream_t)stream);
}

static int conv2d_bwd_bf16_impl(const void* x, const float* w, const void* dy_pre, const void* relu_mask_x, bool mask_is_bits,
                                float* dw, float* db, void* dx, void* scratch, int64_t scratch_bytes, int32_t N, int32_t H,
                                int32_t W, int32_t Cin, int32_t Cw, int32_t Cout, int32_t stride, void* stream) {
  if (!x || !w || !dy_pre || !scratch) { geeco_set_error("conv2d_bwd_bf16: NULL tensor"); return GEECO_ERR_INVALID; }
  if (H != W) { geeco_set_error("conv2d_bwd_bf16: square inputs only"); return GEECO_ERR_INVALID; }
  size_t o_fwd, o_dg[4], o_part; long long pf;
  if ((int64_t)conv_bf16_scratch(N, H, W, Cin, Cout, stride, &o_fwd, o_dg, &o_part, &pf) > scratch_bytes) {
    geeco_set_error("conv2d_bwd_bf16: scratch too small"); return GEECO_ERR_WORKSPACE;
  }
  char* base = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  cudaStream_t st = (cudaStream_t)stream;
  TcGeom g = tc_fwd_geom(H, W, Cin, Cout, stride, N, 1);
  int rc;
  if (dw && Cin == 4 && stride == 1 && W % 2 == 0 && getenv("GEECO_TC_CONV1_PAIR")) {
    float* part = (float*)(base + o_part);
    const long long half = pf / 2;
    int splits = 0, mrows = 0;
    for (int par = 0; par < 2; ++par) {
      TcGeom pg = tc_conv1pair_geom(H, W, Cout, N, 1, par);
      rc = launch_tc_wgrad_partial(pg, Cout, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_pre, part + par * half, half, 1,
                                   &splits, &mrows, st);
      if (rc) return rc;
    }
    rc = launch_conv1pair_reduce(part, part + half, dw, db, splits, 1, mrows, 64, Cw, Cout, 0, 0, st);
    if (rc) return rc;
  } else if (dw) {
    rc = launch_tc_wgrad(g, Cout, Cw, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_pre, dw, db,
                         (float*)(base + o_part), pf, 0, 0, st);
    if (rc) return rc;
  }
  if (dx) {
    if (Cw != Cin) { geeco_set_error("conv2d_bwd_bf16: dx needs Cw == Cin"); return GEECO_ERR_INVALID; }
    int ci = 0;
    TcGeom dgs[4];
    CUtensorMap dm[4];
    const CUtensorMap* dmaps[4];
    for (int py = 0; py < stride; ++py)
      for (int px = 0; px < stride; ++px) {
        TcGeom dg; int taps[9];
        if (!tc_dgrad_geom(H, W, Cin, Cout, stride, py, px, N, 1, &dg, taps)) continue;
        __nv_bfloat16* wp = (__nv_bfloat16*)(base + o_dg[ci]);
        rc = launch_pack_weights(w, wp, 1, 1, 0, Cin, Cout, Cout, dg.ntaps, taps, Cin, dg.Kpad, dg.Kt, st);
        if (rc) return rc;
        rc = make_weight_tensor_map(&dm[ci], wp, Cin, dg.Kpad, Cin);
        if (rc) return rc;
        dgs[ci] = dg; dmaps[ci] = &dm[ci];
        ++ci;
      }
    bool same_grid = true;
    for (int i = 1; i < ci; ++i) same_grid = same_grid && dgs[i].Hm == dgs[0].Hm && dgs[i].Wm == dgs[0].Wm;
    if (same_grid) {
      rc = launch_tc_nn_multi(dgs, dmaps, ci, (const __nv_bfloat16*)dy_pre, nullptr, (const __nv_bfloat16*)relu_mask_x,
                              (__nv_bfloat16*)dx, nullptr, relu_mask_x ? (mask_is_bits ? TC_EPI_MASKBITS : TC_EPI_MASK) : TC_EPI_STORE,
                              0, st);
      if (rc) return rc;
    } else {
      for (int i = 0; i < ci; ++i) {
        rc = launch_tc_nn(dgs[i], dmaps[i], (const __nv_bfloat16*)dy_pre, nullptr, (const __nv_bfloat16*)relu_mask_x,
                          (__nv_bfloat16*)dx, nullptr, relu_mask_x ? (mask_is_bits ? TC_EPI_MASKBITS : TC_EPI_MASK) : TC_EPI_STORE,
                          0, st);
        if (rc) return rc;
      }
    }
  }
  return GEECO_OK;
}
