// libmmae_b200.so -- engine and C ABI (include/mmae_b200.h).
//
// One engine = one MMAE graph instance (multimodal_autoencoder.py:344-452): parameters, two Adam
// states (opt_step :411 and classification_opt_step :443), workspaces and a CUDA stream.
// Parameter layout in the flat buffer P:  [decoder-only vars | encoder + variance vars | head vars]
// so that optimizer 0 owns the prefix [0, end_enc) and optimizer 1 the suffix [begin_enc, end).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mmae_b200.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "chain_tc.cuh"
#include "wgrad_group.cuh"
#include "kernels.cuh"

using namespace mmae;

namespace {

thread_local std::string g_create_error;

struct Var {
  std::string name;
  int64_t rows, cols;   // cols == 0 for vectors (biases); count = rows*max(cols,1)
  int64_t off;
  float l2[2];          // L2 coefficient under optimizer 0 / 1
  int group;            // 0 decoder-only, 1 encoder+variance, 2 head
  int64_t count() const { return rows * (cols > 0 ? cols : 1); }
};

// minimal NCCL binding (dlopen; no link-time dependency so single-GPU users never load it)
struct Id128 { char b[128]; };
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;

bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    err = "libnccl is missing a required symbol"; g_nccl.lib = nullptr; return false;
  }
  return true;
}

int64_t align4(int64_t x) { return (x + 3) & ~int64_t(3); }

}  // namespace

struct mmae_engine {
  // ---- configuration (owned copies)
  mmae_config cfg;
  std::vector<int32_t> starts, layers, head;
  std::vector<uint32_t> type_masks, thresholds;
  int F, M, L, H, E, C;
  int device = 0, num_sms = 148;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  std::string err;
  int sticky = 0;

  // ---- parameters
  std::vector<Var> vars;
  int64_t nP = 0, enc_begin = 0, enc_end = 0;     // group boundaries in the flat buffer
  float *P = nullptr, *G = nullptr;               // G has 8 extra floats: the loss sums (allreduced with it)
  float* PT = nullptr;                            // K-major (transposed) shadows of the 2-D weights, same offsets
  std::vector<char> pt_dirty;                     // per variable
  float* colpart = nullptr; int64_t colpart_cap = 0;   // fused column-sum partials [ceil(M/32), N]
  float *M0 = nullptr, *V0 = nullptr, *M1 = nullptr, *V1 = nullptr;
  int64_t t_opt[2] = {0, 0};
  AdamSeg* d_segs[2] = {nullptr, nullptr};
  int nsegs[2] = {0, 0};
  double* d_scalars = nullptr;                    // MMAE_NUM_SCALARS doubles
  double* d_sums = nullptr;                       // 8 doubles: recon, kl, head loss, head correct

  // ---- small device tables
  uint8_t* d_col_mod = nullptr;
  int32_t* d_starts = nullptr;

  // ---- workspaces (capacity `cap` rows)
  int64_t cap = 0;
  uint32_t *zero_bits = nullptr, *mod_bits = nullptr, *miss_bits = nullptr;
  int64_t noise_rows = 0;
  float* xin[2] = {nullptr, nullptr};             // host-fed batches (double buffered, filled on copy_stream)
  float* yin[2] = {nullptr, nullptr};
  float *gxb = nullptr, *gyb = nullptr;           // batches gathered from a resident dataset (engine stream only: never
  int64_t cap_res = 0;                            // shared with the host staging above, whose copies run on copy_stream)
  cudaEvent_t xin_free[2] = {nullptr, nullptr}, xin_ready[2] = {nullptr, nullptr};
  int xin_turn = 0;
  float* noisy = nullptr;
  std::vector<float*> ea, da, ha;                 // saved activations (encoder / decoder / head)
  float *mu = nullptr, *lv = nullptr, *eps = nullptr, *emb = nullptr, *glv = nullptr;
  float* out = nullptr;                           // [cap, F]: delta_L in training, decoded_X otherwise
  float *dA = nullptr, *dB = nullptr;             // delta ping-pong [cap, maxw]
  std::vector<float*> dch, cpch;                  // backward chain: delta of every dgrad op [cap, N_k] + its per-32-row column sums
  int64_t cap_bch = 0;
  float *hlogits = nullptr, *hdelta = nullptr, *hprobs = nullptr;
  int32_t* hpreds = nullptr;
  float* partials = nullptr; int64_t partials_cap = 0;
  float* colsum_ws = nullptr; int64_t colsum_cap = 0;
  float* splitk_ws = nullptr; int64_t splitk_cap = 0;
  int64_t* d_idx = nullptr;
  int maxw = 0;

  // ---- per-step state in device memory (kernels.cuh StepState): Philox step, Adam step counts and rates
  StepState* d_state = nullptr;

  // ---- CUDA graphs of the train steps (small batches are launch-bound: ~60 launches per step)
  struct GraphKey { int kind; const void* X; const void* Y; const void* T; int64_t B; int noise; float keep; int64_t gb, fr; void* stream;
    bool operator==(const GraphKey& o) const { return kind == o.kind && X == o.X && Y == o.Y && T == o.T && B == o.B && noise == o.noise && keep == o.keep && gb == o.gb && fr == o.fr && stream == o.stream; } };
  struct GraphEntry { GraphKey key; int seen = 0; cudaGraphExec_t exec = nullptr; int64_t n_launches = 0, n_chain = 0, n_bchain = 0, n_wgroup = 0; int opt = 0;
    std::vector<char> dirty_after; bool d_fused_after = false, last_tc_after = false; };
  std::vector<GraphEntry> graphs;
  cudaStream_t gstream = nullptr;   // graphs cannot be captured on the legacy default stream: a blocking stream of our own
                                    // keeps the default stream's implicit ordering with the caller's work
  int graph_mode = -1;
  int64_t graph_replays = 0;

  // ---- pinned staging of host-drawn noise descriptors (rng_mode='numpy'): no stream synchronisation per step
  uint32_t* h_noise[2] = {nullptr, nullptr}; int64_t h_noise_cap = 0; int h_noise_turn = 0;
  cudaEvent_t h_noise_done[2] = {nullptr, nullptr};

  // ---- RNG / sharding
  uint64_t rng_step = 0;
  uint32_t cur_step = 0;                          // step used by the in-flight forward/backward pair
  int64_t global_batch = 0, first_row = 0;        // data-parallel shard description (0 = local batch)

  // ---- resident datasets
  float* ds_X[2] = {nullptr, nullptr}; float* ds_Y[2] = {nullptr, nullptr};
  int64_t ds_rows[2] = {0, 0}; int ds_ycols[2] = {0, 0};
  int64_t* ds_view[2] = {nullptr, nullptr}; int64_t ds_view_rows[2] = {0, 0};     // training view (cross-validation fold) as a row list
  int64_t* d_idx_in = nullptr; int64_t idx_in_cap = 0;

  // ---- NCCL: gradient buckets are all-reduced on comm_stream while backward keeps running on `stream`
  void* comm = nullptr; int world = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  std::vector<cudaEvent_t> comm_events; size_t comm_ev_used = 0;
  cudaEvent_t comm_done = nullptr;
  int64_t buckets_issued = 0;

  int64_t launches = 0;

  // ---- optional per-GEMM device timing (bench.py roofline): CUDA events around every tcgen05 launch
  bool profiling = false;
  struct ProfRec { cudaEvent_t a, b; double flops; int64_t m, n, k; int ta, tb, splits; };
  std::vector<ProfRec> prof_recs; size_t prof_used = 0;
  double prof_ms = 0.0, prof_flops = 0.0; int64_t prof_count = 0;
  int prof_begin(double flops) {
    if (!profiling) return -1;
    if (prof_used == prof_recs.size()) {
      ProfRec r; cudaEventCreate(&r.a); cudaEventCreate(&r.b); r.flops = 0; prof_recs.push_back(r);
    }
    prof_recs[prof_used].flops = flops;
    cudaEventRecord(prof_recs[prof_used].a, stream);
    return (int)prof_used++;
  }
  void prof_end(int i) { if (i >= 0) cudaEventRecord(prof_recs[i].b, stream); }
  void prof_collect() {
    cudaStreamSynchronize(stream);
    for (size_t i = 0; i < prof_used; ++i) {
      float ms = 0.f; cudaEventElapsedTime(&ms, prof_recs[i].a, prof_recs[i].b);
      prof_ms += ms; prof_flops += prof_recs[i].flops; ++prof_count;
      if (getenv("MMAE_PROFILE_DUMP"))
        fprintf(stderr, "[mmae gemm] M=%lld N=%lld K=%lld ta=%d tb=%d splits=%d  %.3f ms  %.1f TFLOP/s\n",
                (long long)prof_recs[i].m, (long long)prof_recs[i].n, (long long)prof_recs[i].k, prof_recs[i].ta,
                prof_recs[i].tb, prof_recs[i].splits, ms, prof_recs[i].flops / (ms * 1e9));
    }
    prof_used = 0;
  }

  // ================================================================= helpers
  int fail(int code, const std::string& m) { err = m; if (code == MMAE_ERR_CUDA) sticky = code; return code; }
  int cuda_fail(cudaError_t e, const char* what) {
    return fail(MMAE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
#define CK(call)                                                           \
  do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail(_e, #call); } while (0)
#define CKL(what)                                                          \
  do { ++launches; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return cuda_fail(_e, what); } while (0)
#define RET(x) do { int _r = (x); if (_r != 0) return _r; } while (0)

  Var* find(const char* name) {
    for (auto& v : vars) if (v.name == name) return &v;
    return nullptr;
  }
  float* pvar(const std::string& n) { Var* v = find(n.c_str()); return v ? P + v->off : nullptr; }
  float* gvar(const std::string& n) { Var* v = find(n.c_str()); return v ? G + v->off : nullptr; }
  int enc_in(int i) const { return i == 0 ? F : layers[i - 1]; }
  int grid_for(int64_t n, int block) const {
    int64_t b = (n + block - 1) / block;
    int64_t cap_b = (int64_t)num_sms * 8;
    return (int)std::max<int64_t>(1, std::min(b, cap_b));
  }
  NoiseView noise_view(bool on) const {
    NoiseView nv; nv.zero_bits = zero_bits; nv.mod_bits = mod_bits; nv.col_mod = d_col_mod;
    nv.zw = (F + 31) / 32; nv.mask_with = cfg.mask_with; nv.enabled = on ? 1 : 0; return nv;
  }
  Epilogue epi(int mode) const {
    Epilogue e; memset(&e, 0, sizeof(e));
    e.mode = mode; e.keep = 1.f; e.seed = cfg.seed; e.step = &d_state->step; e.row0 = first_row; return e;
  }
  void set_dropout(Epilogue& e, float keep, uint32_t slot, int64_t width) const {
    e.keep = keep;
    double t = ceil((double)keep * 16777216.0);
    e.keep_thr = (uint32_t)std::min(t, 16777216.0);
    e.drop_stream = kStreamDrop + slot; e.drop_width = width;
  }

  // ================================================================= construction
  void add_var(const std::string& n, int64_t r, int64_t c, int group, float l2a, float l2b) {
    Var v; v.name = n; v.rows = r; v.cols = c; v.group = group; v.off = 0; v.l2[0] = l2a; v.l2[1] = l2b;
    vars.push_back(v);
  }

  int build(const mmae_config* c) {
    cfg = *c;
    F = c->num_feats; M = c->num_modalities; L = c->num_layers; H = c->num_head_layers;
    if (F <= 0 || L <= 0 || M <= 0 || M > 32) return fail(MMAE_ERR_INVALID, "need F > 0, L > 0, 0 < M <= 32");
    if (!c->modality_starts || !c->layer_sizes) return fail(MMAE_ERR_INVALID, "null modality_starts / layer_sizes");
    starts.assign(c->modality_starts, c->modality_starts + M + 1);
    layers.assign(c->layer_sizes, c->layer_sizes + L);
    if (starts[0] != 0 || starts[M] != F) return fail(MMAE_ERR_INVALID, "modality_starts must run from 0 to num_feats");
    for (int m = 0; m < M; ++m) if (starts[m + 1] < starts[m]) return fail(MMAE_ERR_INVALID, "modality_starts not sorted");
    for (int l : layers) if (l <= 0) return fail(MMAE_ERR_INVALID, "layer size <= 0");
    if (H > 0) { if (!c->head_sizes) return fail(MMAE_ERR_INVALID, "null head_sizes"); head.assign(c->head_sizes, c->head_sizes + H); }
    if (c->variational && L < 2) return fail(MMAE_ERR_INVALID, "variational needs >= 2 layers (multimodal_autoencoder.py:299)");
    if (c->variational && c->tie_weights) return fail(MMAE_ERR_INVALID, "variational forces untied weights (:177)");
    if (c->num_noise_types > 8 || c->num_modalities_to_drop > 4 || c->num_modalities_to_drop < 0)
      return fail(MMAE_ERR_INVALID, "at most 8 noise types and 4 dropped modalities");
    if (c->noise_mode == MMAE_NOISE_INTELLIGENT) {
      if (c->num_noise_types < 1 || !c->noise_type_masks || (c->num_noise_types > 1 && !c->noise_thresholds))
        return fail(MMAE_ERR_INVALID, "intelligent noise needs type masks and thresholds");
      type_masks.assign(c->noise_type_masks, c->noise_type_masks + c->num_noise_types);
      if (c->num_noise_types > 1) thresholds.assign(c->noise_thresholds, c->noise_thresholds + c->num_noise_types - 1);
    }
    E = layers[L - 1]; C = H > 0 ? head[H - 1] : 0;
    if (c->classifier_only && (H != 1 || c->variational)) return fail(MMAE_ERR_INVALID, "classifier_only needs exactly one (logits) head layer and no VAE");
    cfg.modality_starts = nullptr; cfg.layer_sizes = nullptr; cfg.head_sizes = nullptr;
    cfg.noise_type_masks = nullptr; cfg.noise_thresholds = nullptr;

    CK(cudaGetDevice(&device));
    CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    {      // collectives on the highest-priority stream: their CTAs are placed as soon as a persistent GEMM grid retires
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&comm_stream, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&comm_done, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&xin_free[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&xin_ready[i], cudaEventDisableTiming));
    }

    // ---- variables, in reference naming (:279-334)
    const float lam = c->weight_penalty, lamc = c->head_weight_penalty;
    const bool tied = c->tie_weights != 0;
    char buf[64];
    for (int i = 0; i < L; ++i) {
      if (!tied) { snprintf(buf, 64, "decode_weights%d", i); add_var(buf, layers[i], enc_in(i), 0, lam, 0.f); }
      snprintf(buf, 64, "decode_biases%d", i); add_var(buf, enc_in(i), 0, 0, 0.f, 0.f);
    }
    for (int i = 0; i < L; ++i) {
      snprintf(buf, 64, "weights%d", i); add_var(buf, enc_in(i), layers[i], 1, tied ? 2.f * lam : lam, c->classifier_only ? lamc : 0.f);
      snprintf(buf, 64, "encode_biases%d", i); add_var(buf, layers[i], 0, 1, 0.f, 0.f);
    }
    if (c->variational) {
      add_var("variance_weights", layers[L - 2], E, 1, lam, 0.f);
      add_var("variance_bias", E, 0, 1, 0.f, 0.f);
    }
    for (int i = 0; i < H; ++i) {
      int din = i == 0 ? E : head[i - 1];
      snprintf(buf, 64, "classification_weights%d", i); add_var(buf, din, head[i], 2, 0.f, lamc);
      snprintf(buf, 64, "classification_biases%d", i); add_var(buf, head[i], 0, 2, 0.f, 0.f);
    }
    int64_t off = 0; int prev_group = 0;
    enc_begin = -1;
    for (auto& v : vars) {
      if (v.group >= 1 && enc_begin < 0) enc_begin = off;
      if (v.group == 2 && prev_group < 2) enc_end = off;
      prev_group = v.group;
      v.off = off; off += align4(v.count());
    }
    nP = off;
    if (H == 0) enc_end = nP;
    if (enc_begin < 0) enc_begin = 0;

    CK(cudaMalloc(&P, nP * 4)); CK(cudaMemset(P, 0, nP * 4));
    CK(cudaMalloc(&G, (nP + 8) * 4)); CK(cudaMemset(G, 0, (nP + 8) * 4));
    CK(cudaMalloc(&PT, nP * 4)); CK(cudaMemset(PT, 0, nP * 4));
    pt_dirty.assign(vars.size(), 1);
    CK(cudaMalloc(&M0, enc_end * 4)); CK(cudaMemset(M0, 0, enc_end * 4));
    CK(cudaMalloc(&V0, enc_end * 4)); CK(cudaMemset(V0, 0, enc_end * 4));
    if (H > 0) {
      int64_t n1 = nP - enc_begin;
      CK(cudaMalloc(&M1, n1 * 4)); CK(cudaMemset(M1, 0, n1 * 4));
      CK(cudaMalloc(&V1, n1 * 4)); CK(cudaMemset(V1, 0, n1 * 4));
    }
    CK(cudaMalloc(&d_scalars, MMAE_NUM_SCALARS * 8)); CK(cudaMemset(d_scalars, 0, MMAE_NUM_SCALARS * 8));
    CK(cudaMalloc(&d_sums, 8 * 8)); CK(cudaMemset(d_sums, 0, 64));
    CK(cudaMalloc(&d_state, sizeof(StepState))); CK(cudaMemset(d_state, 0, sizeof(StepState)));
    for (int o = 0; o < 2; ++o) {
      std::vector<AdamSeg> segs;
      for (auto& v : vars) { AdamSeg s; s.begin = v.off; s.l2 = v.l2[o]; s.pad = 0.f; s.rows = (int)v.rows; s.cols = (int)v.cols; segs.push_back(s); }
      nsegs[o] = (int)segs.size();
      CK(cudaMalloc(&d_segs[o], segs.size() * sizeof(AdamSeg)));
      CK(cudaMemcpy(d_segs[o], segs.data(), segs.size() * sizeof(AdamSeg), cudaMemcpyHostToDevice));
    }
    std::vector<uint8_t> cm(F);
    for (int m = 0; m < M; ++m) for (int cidx = starts[m]; cidx < starts[m + 1]; ++cidx) cm[cidx] = (uint8_t)m;
    CK(cudaMalloc(&d_col_mod, F)); CK(cudaMemcpy(d_col_mod, cm.data(), F, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_starts, (M + 1) * 4)); CK(cudaMemcpy(d_starts, starts.data(), (M + 1) * 4, cudaMemcpyHostToDevice));

    maxw = E;
    for (int l : layers) maxw = std::max(maxw, l);
    for (int h : head) maxw = std::max(maxw, h);
    ea.assign(L, nullptr); da.assign(L, nullptr); ha.assign(std::max(H, 1), nullptr);
    return ensure_cap(std::max<int64_t>(c->max_batch, 1));
  }

  template <class T> int realloc_dev(T*& p, int64_t count) {
    if (p) { CK(cudaFree(p)); p = nullptr; }
    CK(cudaMalloc(&p, std::max<int64_t>(count, 1) * sizeof(T)));
    return 0;
  }

  // Workspaces grow on demand in three groups so that a 10 M-row fill-in pass does not allocate training buffers:
  //   base  (ensure_cap)  : noise descriptor, embedding, reconstruction, head outputs, reduction scratch
  //   acts  (ensure_acts) : materialised noisy X, saved activations, delta ping-pong, bias-gradient partials
  //   host  (ensure_host) : double-buffered staging of host-fed batches
  int ensure_cap(int64_t B) {
    if (B <= cap) return 0;
    CK(cudaStreamSynchronize(stream));
    clear_graphs();
    int64_t nc = std::max<int64_t>(B, cap + cap / 2);
    const int zw = (F + 31) / 32;
    RET(realloc_dev(zero_bits, nc * zw)); RET(realloc_dev(mod_bits, nc)); RET(realloc_dev(miss_bits, nc));
    noise_rows = 0;
    RET(realloc_dev(mu, nc * E));
    if (cfg.variational) { RET(realloc_dev(lv, nc * E)); RET(realloc_dev(eps, nc * E)); RET(realloc_dev(emb, nc * E)); RET(realloc_dev(glv, nc * E)); }
    RET(realloc_dev(out, nc * F));
    if (H > 0) {
      for (int i = 0; i < H; ++i) RET(realloc_dev(ha[i], nc * head[i]));
      RET(realloc_dev(hlogits, nc * C)); RET(realloc_dev(hdelta, nc * C));
      RET(realloc_dev(hprobs, nc * C)); RET(realloc_dev(hpreds, nc * C));
    }
    RET(realloc_dev(d_idx, nc));
    // loss partials: one per CTA of the largest loss-carrying GEMM (SIMT worst case) or reduction grid
    int64_t pc = std::max<int64_t>(gemm_simt_num_ctas(nc, F), (int64_t)num_sms * 16) + 16;
    if (pc > partials_cap) { RET(realloc_dev(partials, pc)); partials_cap = pc; }
    int64_t cs = (int64_t)64 * std::max(F, maxw);
    if (cs > colsum_cap) { RET(realloc_dev(colsum_ws, cs)); colsum_cap = cs; }
    cap = nc;
    return 0;
  }
  int64_t cap_acts = 0, cap_host = 0;
  // pipelined host inference (mmae_forward_host on large batches): per-slot device staging of the outputs
  cudaStream_t d2h_stream = nullptr;
  float* pipe_out[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // [slot][recon, filled, embedding]
  int64_t pipe_cap = 0;
  cudaEvent_t pipe_ready[2] = {nullptr, nullptr}, pipe_free[2] = {nullptr, nullptr};
  int ensure_acts(int64_t B) {
    RET(ensure_cap(B));
    if (B <= cap_acts) return 0;
    CK(cudaStreamSynchronize(stream));
    clear_graphs();
    int64_t nc = std::max<int64_t>(B, cap_acts + cap_acts / 2);
    RET(realloc_dev(noisy, nc * F));
    for (int i = 0; i + 1 < L; ++i) RET(realloc_dev(ea[i], nc * layers[i]));
    for (int j = 0; j + 1 < L; ++j) RET(realloc_dev(da[j], nc * layers[L - 2 - j]));
    RET(realloc_dev(dA, nc * maxw)); RET(realloc_dev(dB, nc * maxw));
    int64_t cpn = ((nc + 31) / 32) * (int64_t)std::max(F, maxw);
    if (cpn > colpart_cap) { RET(realloc_dev(colpart, cpn)); colpart_cap = cpn; }
    cap_acts = nc;
    return 0;
  }
  int ensure_host(int64_t B) {
    RET(ensure_cap(B));
    if (B <= cap_host) return 0;
    CK(cudaStreamSynchronize(stream)); CK(cudaStreamSynchronize(copy_stream));
    clear_graphs();
    int64_t nc = std::max<int64_t>(B, cap_host + cap_host / 2);
    for (int i = 0; i < 2; ++i) { RET(realloc_dev(xin[i], nc * F)); RET(realloc_dev(yin[i], nc * std::max(C, 1))); }
    cap_host = nc;
    return 0;
  }

  int ensure_resident(int64_t B) {
    RET(ensure_cap(B));
    if (B <= cap_res) return 0;
    CK(cudaStreamSynchronize(stream));
    clear_graphs();
    int64_t nc = std::max<int64_t>(B, cap_res + cap_res / 2);
    RET(realloc_dev(gxb, nc * F)); RET(realloc_dev(gyb, nc * std::max(C, 1)));
    cap_res = nc;
    return 0;
  }

  // width of the delta produced by dgrad op k of the backward chain (see backward_chain)
  int bch_width(int k) const { return k < L ? layers[k] : layers[2 * L - 2 - k]; }
  int ensure_bchain(int64_t B) {
    RET(ensure_acts(B));
    if (B <= cap_bch) return 0;
    CK(cudaStreamSynchronize(stream));
    clear_graphs();
    const int nops = 2 * L - 1;
    dch.resize(nops, nullptr); cpch.resize(nops, nullptr);
    int64_t nc = std::max<int64_t>(B, cap_bch + cap_bch / 2);
    for (int k = 0; k < nops; ++k) {
      RET(realloc_dev(dch[k], nc * bch_width(k)));
      RET(realloc_dev(cpch[k], ((nc + 31) / 32) * (int64_t)bch_width(k)));
    }
    cap_bch = nc;
    return 0;
  }

  int* tile_counters = nullptr; int64_t counters_cap = 0;
  int ensure_counters(int64_t count) {
    if (count <= counters_cap) return 0;
    CK(cudaStreamSynchronize(stream));
    clear_graphs();
    const int64_t nc = std::max<int64_t>(count, 1024);
    RET(realloc_dev(tile_counters, nc));
    CK(cudaMemset(tile_counters, 0, nc * sizeof(int)));
    counters_cap = nc;
    return 0;
  }
  int ensure_splitk(int64_t count) {
    if (count <= splitk_cap) return 0;
    CK(cudaStreamSynchronize(stream));
    clear_graphs();
    RET(realloc_dev(splitk_ws, count)); splitk_cap = count; return 0;
  }

  void release() {
    auto fr = [](void* p) { if (p) cudaFree(p); };
    clear_graphs();
    fr(d_state); fr(PT); fr(colpart); fr(P); fr(G); fr(M0); fr(V0); fr(M1); fr(V1); fr(d_scalars); fr(d_sums); fr(d_segs[0]); fr(d_segs[1]);
    fr(d_col_mod); fr(d_starts); fr(zero_bits); fr(mod_bits); fr(miss_bits);
    for (int i = 0; i < 2; ++i) { fr(xin[i]); fr(yin[i]); fr(ds_X[i]); fr(ds_Y[i]); fr(ds_view[i]); }
    fr(d_idx_in);
    fr(noisy); fr(gxb); fr(gyb); fr(wg_ws);
    for (auto q : dch) fr(q); for (auto q : cpch) fr(q);
    for (auto p : ea) fr(p); for (auto p : da) fr(p); for (auto p : ha) fr(p);
    fr(mu); fr(lv); fr(eps); fr(emb); fr(glv); fr(out); fr(dA); fr(dB);
    fr(hlogits); fr(hdelta); fr(hprobs); fr(hpreds); fr(partials); fr(colsum_ws); fr(splitk_ws); fr(tile_counters); fr(d_idx);
    for (int i = 0; i < 2; ++i) { if (xin_free[i]) cudaEventDestroy(xin_free[i]); if (xin_ready[i]) cudaEventDestroy(xin_ready[i]); }
    for (auto& r : prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (gstream) cudaStreamDestroy(gstream);
    if (d2h_stream) cudaStreamDestroy(d2h_stream);
    for (int i = 0; i < 2; ++i) { if (h_noise[i]) cudaFreeHost(h_noise[i]); if (h_noise_done[i]) cudaEventDestroy(h_noise_done[i]); }
    for (int i = 0; i < 2; ++i) {
      for (int j = 0; j < 3; ++j) if (pipe_out[i][j]) cudaFree(pipe_out[i][j]);
      if (pipe_ready[i]) cudaEventDestroy(pipe_ready[i]);
      if (pipe_free[i]) cudaEventDestroy(pipe_free[i]);
    }
    if (comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(comm); comm = nullptr; }
    for (auto ev : comm_events) cudaEventDestroy(ev);
    if (comm_done) cudaEventDestroy(comm_done);
    if (adam_gate) cudaEventDestroy(adam_gate);
    if (comm_stream) cudaStreamDestroy(comm_stream);
  }

  // Refreshes every stale K-major weight shadow in one launch (called at the top of a tf32 forward pass).
  int refresh_shadows() {
    TransposeGroup g; g.n = 0; int tiles = 0;
    for (size_t i = 0; i < vars.size() && g.n < TransposeGroup::kMax; ++i) {
      Var& v = vars[i];
      if (v.cols == 0 || !pt_dirty[i]) continue;
      g.off[g.n] = v.off; g.rows[g.n] = (int)v.rows; g.cols[g.n] = (int)v.cols; g.tile0[g.n] = tiles;
      tiles += (int)(((v.rows + 31) / 32) * ((v.cols + 31) / 32));
      pt_dirty[i] = 0; ++g.n;
    }
    if (g.n == 0) return 0;
    g.tile0[g.n] = tiles;
    transpose_group_kernel<<<tiles, dim3(32, 8), 0, stream>>>(P, PT, g);
    CKL("transpose_group");
    return 0;
  }
  // K-major shadow of a weight variable (refreshed lazily after set_variable / Adam); null if w is not a variable
  const float* shadowT(const float* w, int64_t rows, int64_t cols) {
    for (size_t i = 0; i < vars.size(); ++i) {
      Var& v = vars[i];
      if (P + v.off != w || v.cols == 0 || v.rows != rows || v.cols != cols) continue;
      if (pt_dirty[i]) {
        dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
        transpose_kernel<<<grid, block, 0, stream>>>(w, PT + v.off, (int)rows, (int)cols);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) return nullptr;
        pt_dirty[i] = 0;
      }
      return PT + v.off;
    }
    return nullptr;
  }
  void mark_dirty(int64_t begin, int64_t end) {
    for (size_t i = 0; i < vars.size(); ++i) if (vars[i].off >= begin && vars[i].off < end) pt_dirty[i] = 1;
  }
  int fuse_noise_mode = -1;      // per engine, read at the first use: MMAE_FUSE_NOISE=1 opts in (measured slower, see gemm_tc2.cu)
  bool noise_fusion_on() {
    if (fuse_noise_mode < 0) { const char* ev = getenv("MMAE_FUSE_NOISE"); fuse_noise_mode = (ev && ev[0] == '1') ? 1 : 0; }
    return fuse_noise_mode == 1;
  }
  bool starts_aligned32() const { for (int v : starts) if (v & 31) return false; return true; }
  bool last_gemm_tc = false;
  bool d_fused = false;      // the current delta's column-sum partials are valid in colpart
  // Small models: Adam rewrites the K-major weight shadows in the same pass (scattered 4-byte stores beat a launch);
  // the shadows are then always current outside a step, and captured graphs need no transpose launch.
  bool shadow_in_adam() const { return cfg.precision == MMAE_PREC_TF32 && nP <= ((int64_t)1 << 21); }
  // ... large models get the same guarantee from the tiled Adam pass (adam_tiled_kernel) when every optimizer's variables
  // fit its table: either way the shadows are current outside a step and captured graphs carry no transpose launch
  bool adam_keeps_shadows() const {
    if (cfg.precision != MMAE_PREC_TF32) return false;
    if (shadow_in_adam()) return true;
    static const bool tiled_off = getenv("MMAE_ADAM_TILED") && getenv("MMAE_ADAM_TILED")[0] == '0';
    if (tiled_off) return false;
    int n0 = 0, n1 = 0;
    for (const Var& v : vars) { if (v.off < enc_end) ++n0; if (v.off >= enc_begin) ++n1; }
    return n0 <= AdamTiles::kMax && (H == 0 || n1 <= AdamTiles::kMax);
  }
  // Fast small-config train step (train_core only, no data parallelism): the loss partials of the forward chain are
  // summed, and the per-step scalars finalised, inside grad_assemble_kernel instead of in launches of their own.
  bool fast_step = false;
  int64_t pending_loss_partials = 0;     // > 0: partials[0..n) still have to be summed into d_sums[0]
  bool step_finalized = false;           // grad_assemble_kernel already did finalize_scalars' work for this step
  int flush_pending_loss() {
    if (pending_loss_partials > 0) { int64_t n = pending_loss_partials; pending_loss_partials = 0; return reduce_partials(n, 0); }
    return 0;
  }
  // The reconstruction-loss GEMM of a resident-dataset step can read its clean target rows through the sampled index
  // list when it runs on the two-SM kernel with the row-layout epilogue (wide models): no clean copy of the batch.
  bool final_gemm_gathers_target(int64_t B, float keep) const {
    static const bool off = (getenv("MMAE_TMA_EPI") && getenv("MMAE_TMA_EPI")[0] == '0') || (getenv("MMAE_TC2") && getenv("MMAE_TC2")[0] == '0') ||
                            (getenv("MMAE_GATHER_TARGET") && getenv("MMAE_GATHER_TARGET")[0] == '0');
    const int k_last = layers[0];
    // (F > 512 also rules out the whole-network kernel, whose TMA-loaded target tile cannot gather rows)
    return !off && cfg.precision == MMAE_PREC_TF32 && !cfg.variational && keep >= 1.f && B >= 256 && F > 512 && (F & 3) == 0 &&
           k_last >= 32 && (k_last & 3) == 0;
  }
  float* wg_ws = nullptr; int64_t wg_ws_cap = 0;
  int64_t wgroup_launches = 0;

  // ================================================================= GEMM dispatch
  // C = opA(A) opB(B) with epilogue.  n_partials receives the number of loss partials written.
  int gemm(bool ta, bool tb, int64_t m, int64_t n, int64_t k, const float* A, int64_t lda, const float* B,
           int64_t ldb, float* Cp, int64_t ldc, const NoiseView& nv, const Epilogue& ep, int64_t* n_partials,
           bool allow_splitk) {
    GemmArgs g; g.splits = 1; g.k_per_split = 0; g.M = m; g.N = n; g.K = k; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = Cp; g.ldc = ldc;
    g.noise = nv; g.ep = ep; g.noise_aligned32 = starts_aligned32() ? 1 : 0;
    last_gemm_tc = false;
    bool tbk = tb;      // forward GEMMs read the K-major shadow of the weight: decide two-SM eligibility with that operand
    if (!tb && (k & 3) == 0 && nv.enabled) { if (shadowT(B, k, n)) tbk = true; }
    if (cfg.precision == MMAE_PREC_TF32 && (tc_gemm_eligible(ta, tb, g) || (nv.enabled && tc2_eligible(ta, tbk, g)))) {
      last_gemm_tc = true;
      if (!tb && (k & 3) == 0) {      // y = x.W with W stored [K,N]: use its K-major shadow W^T [N,K] when W is a variable
        const float* wt = shadowT(B, k, n);
        if (wt) { g.B = wt; g.ldb = k; tb = true; }
      }
      const bool two_sm = tc2_eligible(ta, tb, g);       // 256 x 256 tiles on CTA pairs (cta_group::2)
      if (ep.aux_rows && !two_sm) return fail(MMAE_ERR_STATE, "internal: gathered loss target outside the two-SM row-layout epilogue");
      if (nv.enabled && !two_sm) return fail(MMAE_ERR_STATE, "internal: noisy operand reached the one-SM tcgen05 GEMM");
      if (nv.enabled) ++fused_noise_launches;
      const int max_s = allow_splitk && ep.mode == EPI_PLAIN ? 64 : 1;
      // While gradient buckets are being all-reduced, NCCL's CTAs hold SMs of their own (they cannot share one with a
      // 227 KB GEMM CTA); a persistent grid sized for all 148 SMs would then need a second wave for its last CTAs.
      // The reservation lasts for the launches that can overlap the bucket just issued (reserve_credits), not for the
      // whole backward pass: the collectives are busy for about a third of it.
      int sms_eff = num_sms;
      if (comm_busy() && reserve_credits > 0) { sms_eff = std::max(2, (num_sms - comm_reserve) & ~1); --reserve_credits; }
      TcPlan pl = two_sm ? tc2_plan(g, sms_eff, max_s) : tc_plan(g, sms_eff, max_s);
      if (pl.splits > 1) RET(ensure_splitk((int64_t)pl.splits * m * n));
      int pr = prof_begin(2.0 * (double)m * (double)n * (double)k);
      if (pr >= 0) { auto& R = prof_recs[pr]; R.m = m; R.n = n; R.k = k; R.ta = ta; R.tb = tb; R.splits = pl.splits; }
      cudaError_t e = two_sm ? launch_gemm_tc2(ta, tb, g, pl, splitk_ws, stream) : launch_gemm_tc(ta, tb, g, pl, splitk_ws, stream);
      prof_end(pr);
      ++launches;
      if (e != cudaSuccess) return cuda_fail(e, "tcgen05 gemm launch");
      if (pl.splits > 1) {
        splitk_reduce_kernel<<<grid_for(m * n, 256), 256, 0, stream>>>(splitk_ws, m * n, pl.splits, Cp, n, ldc, ep.beta);
        CKL("splitk_reduce");
      }
      if (n_partials) *n_partials = pl.grid;
      return 0;
    }
    if (ep.aux_rows) return fail(MMAE_ERR_STATE, "internal: gathered loss target reached the CUDA-core GEMM");
    g.ep.colsum_partials = nullptr;
    g.splits = 1; g.k_per_split = 0; g.ws = nullptr; g.tile_counters = nullptr;
    {      // in-kernel split-K (any epilogue): the last CTA of a tile reduces the slices and finishes the tile
      const int s = simt_pick_splits(m, n, k, num_sms);
      if (s > 1) {
        const int64_t tiles = gemm_simt_num_ctas(m, n);
        RET(ensure_splitk((int64_t)s * tiles * SG_BM * SG_BN));
        RET(ensure_counters(tiles));
        g.k_per_split = (((k + s - 1) / s + SG_BK - 1) / SG_BK) * SG_BK;
        g.splits = (int)((k + g.k_per_split - 1) / g.k_per_split);
        g.ws = splitk_ws; g.tile_counters = tile_counters;
      }
    }
    (void)allow_splitk;
    cudaError_t e = launch_gemm_simt(ta, tb, g, stream);
    ++launches;
    if (e != cudaSuccess) return cuda_fail(e, "simt gemm launch");
    if (n_partials) *n_partials = gemm_simt_num_ctas(m, n);
    return 0;
  }

  int colsum(const float* D, int64_t rows, int n, int64_t ld, float* outp) {
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>(64, rows / 256));
    dim3 grid((n + 31) / 32, splits), block(32, 8);
    if (splits == 1) {       // small batches: one pass straight into the gradient (one launch instead of two)
      colsum_partial_kernel<<<grid, block, 0, stream>>>(D, rows, n, ld, outp, 1);
      CKL("colsum");
      return 0;
    }
    colsum_partial_kernel<<<grid, block, 0, stream>>>(D, rows, n, ld, colsum_ws, splits);
    CKL("colsum_partial");
    colsum_final_kernel<<<(n + 127) / 128, 128, 0, stream>>>(colsum_ws, n, splits, outp);
    CKL("colsum_final");
    return 0;
  }

  // db = column sums of delta [rows, n].  `fused` says the GEMM that produced delta already left per-32-row
  // partial sums in colpart (tcgen05 epilogue), so only ceil(rows/32) partial rows are read instead of delta.
  int bias_grad(const float* D, int64_t rows, int n, int64_t ld, float* outp, bool fused, const float* part = nullptr) {
    if (fused) return colsum(part ? part : colpart, (rows + 31) / 32, n, n, outp);
    return colsum(D, rows, n, ld, outp);
  }

  int reduce_partials(int64_t n, int slot, bool accumulate = false) {
    reduce_partials_kernel<<<1, 256, 0, stream>>>(partials, n, d_sums + slot, accumulate ? 1 : 0);
    CKL("reduce_partials");
    return 0;
  }

  // ================================================================= forward (:366-378, :428)
  struct FwdOpts {
    const float* X; const float* target; const float* labels;
    int64_t B; bool noise; float keep; bool train_recon;   // train_recon: last layer emits delta_L
    bool decoder; bool headp; float* recon_out;
    float* fill_out = nullptr;     // whole-network kernel: write the filled matrix (A15) instead of decoded_X
    bool need_mu = true;           // the embedding is wanted in global memory (it is not for plain predict / fill-in)
    bool noisy_ready = false;      // `noisy` already holds the noisy batch of X (sample_noise_kernel drew and applied it)
    const int64_t* target_rows = nullptr;   // the loss target is target[target_rows[r]] (rows of a resident dataset; no clean copy of the batch)
  };

  int begin_step(int64_t B, bool noise) {
    if (sticky) return fail(MMAE_ERR_CUDA, "engine is in a sticky CUDA error state: " + err);
    if (B <= 0) return fail(MMAE_ERR_INVALID, "batch must be > 0");
    RET(ensure_cap(B));
    if (noise && noise_rows < B) return fail(MMAE_ERR_STATE, "noise descriptor covers fewer rows than the batch; call mmae_set_noise / mmae_gen_noise first");
    cur_step = (uint32_t)rng_step;
    return 0;
  }

  bool step_synced = true;       // device StepState.step == (uint32_t)rng_step
  int advance_step(bool on_device = true) {
    rng_step += 1;
    if (!on_device) return 0;       // the caller folds the device-side increment into finalize_scalars
    advance_step_kernel<<<1, 1, 0, stream>>>(d_state); CKL("advance_step");
    return 0;
  }
  int set_step(uint64_t step) {
    if (step == rng_step && step_synced) return 0;      // the device copy already holds it (advance_step keeps both in sync)
    rng_step = step; step_synced = true;
    const uint32_t v = (uint32_t)step;
    CK(cudaMemcpyAsync(&d_state->step, &v, 4, cudaMemcpyHostToDevice, stream));    // pageable source: staged before return
    return 0;
  }
  int set_t(int opt, int64_t t) {
    t_opt[opt] = t;
    const long long v = (long long)t;
    CK(cudaMemcpyAsync(&d_state->t[opt], &v, 8, cudaMemcpyHostToDevice, stream));
    return 0;
  }

  // ----------------------------------------------------------------- CUDA graphs
  // A train step with fixed (input pointer, batch, flags) is the same launch sequence every time once the per-step
  // scalars live in device memory: the first call runs eagerly (sizes every workspace), the second is captured,
  // later ones replay one graph launch instead of ~30-70 kernel launches + tensor-map encodes.
  void clear_graphs() {
    for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }
  bool graphs_allowed() {
    if (graph_mode < 0) { const char* ev = getenv("MMAE_GRAPHS"); graph_mode = (ev && ev[0] == '0') ? 0 : 1; }
    static const bool dp_graphs = !(getenv("MMAE_DP_GRAPHS") && getenv("MMAE_DP_GRAPHS")[0] == '0');
    return graph_mode == 1 && !profiling && (!dp_on() || dp_graphs) && !sticky;
  }
  template <class Body> int run_graphed(const GraphKey& key, int opt, int64_t B, Body body) {
    if (!graphs_allowed() || B * (int64_t)F > ((int64_t)1 << 26)) return body();      // large batches are not launch-bound
    GraphEntry* ge = nullptr;
    for (auto& g : graphs) if (g.key == key) { ge = &g; break; }
    if (!ge) {
      if (graphs.size() >= 16) clear_graphs();
      GraphEntry n; n.key = key; n.seen = 1; n.opt = opt; graphs.push_back(n);
      return body();                                             // eager: grows workspaces, configures kernels
    }
    if (!stream && !gstream) CK(cudaStreamCreate(&gstream));
    cudaStream_t cs = stream ? stream : gstream;
    if (ge->exec) {                                              // replay + the host-side effects of one step
      rng_step += 1; t_opt[ge->opt] += 1; launches += ge->n_launches; chain_launches += ge->n_chain; bchain_launches += ge->n_bchain; wgroup_launches += ge->n_wgroup; last_B = B;
      pt_dirty = ge->dirty_after; d_fused = ge->d_fused_after; last_gemm_tc = ge->last_tc_after;
      ++graph_replays;
      CK(cudaGraphLaunch(ge->exec, cs));
      return 0;
    }
    // capture.  All K-major weight shadows are forced stale so that the graph always refreshes the ones it reads.
    const size_t idx = (size_t)(ge - graphs.data());
    if (!adam_keeps_shadows()) std::fill(pt_dirty.begin(), pt_dirty.end(), 1);
    else { int rr = refresh_shadows(); if (rr) return rr; }     // (current outside steps: set_variable refreshes eagerly, Adam rewrites them)
    const int64_t l0 = launches, c0 = chain_launches, bc0 = bchain_launches, wg0 = wgroup_launches, cap0 = cap, capa0 = cap_acts, caph0 = cap_host, sk0 = splitk_cap;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) { (void)cudaGetLastError(); graph_mode = 0; return body(); }
    cudaStream_t user_stream = stream;
    stream = cs;
    int r = body();
    stream = user_stream;
    ce = cudaStreamEndCapture(cs, &graph);
    GraphEntry& g = graphs[idx];
    if (r != 0 || ce != cudaSuccess || !graph || cap != cap0 || cap_acts != capa0 || cap_host != caph0 || splitk_cap != sk0) {
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      graph_mode = 0; clear_graphs();                            // something in the step is not capturable: stay eager
      if (r != 0) return r;
      return fail(MMAE_ERR_CUDA, "graph capture of the train step failed; rerun with MMAE_GRAPHS=0");
    }
    ce = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { g.exec = nullptr; graph_mode = 0; return cuda_fail(ce, "cudaGraphInstantiate"); }
    g.n_launches = launches - l0; g.n_chain = chain_launches - c0; g.n_bchain = bchain_launches - bc0; g.n_wgroup = wgroup_launches - wg0;
    g.dirty_after = pt_dirty; g.d_fused_after = d_fused; g.last_tc_after = last_gemm_tc;
    CK(cudaGraphLaunch(g.exec, cs));
    return 0;
  }

  // Whole-network launch (chain_tc.cuh): encoder + decoder + loss in one persistent tcgen05 kernel, activations
  // resident in TMEM.  Returns 1 when it ran, 0 when the configuration does not fit (caller runs the per-layer
  // GEMMs), or a (negative) error code.
  // debug timeline of CTA 0 (MMAE_CHAIN_TRACE=1): clock64 deltas per tile and op
  long long* trace_begin(ChainParams& cp) {
    static const bool want_trace = getenv("MMAE_CHAIN_TRACE") != nullptr;
    if (!want_trace) return nullptr;
    long long* trace = nullptr;
    cudaMalloc(&trace, 2 * 64 * CH_MAX_OPS * 4 * 8); cudaMemsetAsync(trace, 0, 2 * 64 * CH_MAX_OPS * 4 * 8, stream); cp.trace = trace;
    return trace;
  }
  void trace_end(const ChainParams& cp, long long* trace, const char* tag) {
    if (!trace) return;
    std::vector<long long> h(2 * 64 * CH_MAX_OPS * 4);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost); cudaFree(trace);
    const long long t0 = h[0];
    for (int it = 0; it < 8 && it * num_sms < cp.m_tiles; ++it)
      for (int i = 0; i < cp.nops; ++i) {
        const long long* q = &h[(it * CH_MAX_OPS + i) * 4];
        const long long* w = &h[64 * CH_MAX_OPS * 4 + (it * CH_MAX_OPS + i) * 4];
        fprintf(stderr, "[chain trace %s] tile %d op %d: mma start %8lld issued %8lld | epi start %8lld end %8lld | issuer waited: A-chunks %6lld W %6lld X %6lld\n",
                tag, it, i, q[0] - t0, q[1] - t0, q[2] - t0, q[3] - t0, w[0], w[1], w[2]);
      }
  }
  int64_t chain_launches = 0;
  int64_t fused_noise_launches = 0;      // two-SM GEMM launches that applied the mask + noise to their A tile in shared memory
  int chain_mode = -1;
  bool fill_fused = false;       // the last forward wrote the filled matrix from the whole-network kernel
  int forward_chain(const FwdOpts& o, const float* a) {
    if (chain_mode < 0) { const char* ev = getenv("MMAE_CHAIN"); chain_mode = (ev && ev[0] == '0') ? 0 : 1; }
    if (!chain_mode || cfg.precision != MMAE_PREC_TF32 || cfg.variational || o.B < 32) return 0;
    const int64_t B = o.B;
    const bool save = o.train_recon;
    if (o.fill_out && (o.train_recon || o.noise)) return 0;
    std::vector<ChainLayer> ls;
    double flops = 0.0;
    for (int i = 0; i < L; ++i) {
      const bool last = i == L - 1;
      const int din = enc_in(i), dout = layers[i];
      char wn[32], bn[32]; snprintf(wn, 32, "weights%d", i); snprintf(bn, 32, "encode_biases%d", i);
      if ((din & 3)) return 0;
      ChainLayer l; l.K = din; l.N = dout;
      l.Wkm = shadowT(pvar(wn), din, dout); l.ldw = din;
      if (!l.Wkm) return 0;
      l.ep = epi(EPI_BIAS_ACT); l.ep.bias = pvar(bn);
      if (!last) { l.ep.act = cfg.activation; if (o.keep < 1.f) set_dropout(l.ep, o.keep, (uint32_t)i, dout); l.out = save ? ea[i] : nullptr; }
      else { l.ep.act = MMAE_ACT_LINEAR; l.out = (save || o.need_mu) ? mu : nullptr; }
      l.ldo = dout;
      ls.push_back(l); flops += 2.0 * din * dout;
    }
    float* dst = o.recon_out ? o.recon_out : out;
    for (int j = 0; j < L; ++j) {
      const int i = L - 1 - j;
      const int din = layers[i], dout = enc_in(i);
      const bool last = j == L - 1;
      char wn[32], bn[32]; snprintf(bn, 32, "decode_biases%d", i);
      ChainLayer l; l.K = din; l.N = dout; l.ldw = din;
      if ((din & 3)) return 0;
      if (cfg.tie_weights) { snprintf(wn, 32, "weights%d", i); l.Wkm = pvar(wn); }          // W_i [dout, din] is K-major for W_i^T (:284)
      else { snprintf(wn, 32, "decode_weights%d", i); l.Wkm = shadowT(pvar(wn), din, dout); }
      if (!l.Wkm) return 0;
      if (!last) {
        l.ep = epi(EPI_BIAS_ACT); l.ep.bias = pvar(bn); l.ep.act = cfg.activation;
        if (o.keep < 1.f) set_dropout(l.ep, o.keep, 32u + (uint32_t)j, dout);
        l.out = save ? da[j] : nullptr;
      } else {
        l.ep = epi(o.train_recon ? EPI_LOSS_TRAIN : EPI_LOSS_PRED); l.ep.bias = pvar(bn); l.ep.loss = cfg.loss_func;
        l.ep.target = o.target; l.ep.ldt = F; l.ep.loss_partials = o.target ? partials : nullptr;
        if (o.train_recon) l.ep.colsum_partials = colpart;
        l.out = dst;
        if (o.fill_out) {          // fill-in fused into the last epilogue: missing blocks <- decoded_X, the rest <- X
          if (!o.target) l.ep.target = o.X;
          l.ep.loss_partials = o.target ? partials : nullptr;
          l.ep.fill_bits = miss_bits; l.ep.fill_col_mod = d_col_mod; l.out = o.fill_out;
        }
      }
      l.ldo = dout;
      ls.push_back(l); flops += 2.0 * din * dout;
    }
    ChainParams cp;
    if (!chain_build(cp, a, B, F, ls)) return 0;
    if (o.fill_out && M <= 32) { cp.scan_miss = 1; cp.num_mod = M; cp.starts = d_starts; }
    const int grid = std::min(cp.m_tiles, num_sms);
    static const int chain_stagger = getenv("MMAE_CHAIN_STAGGER") ? atoi(getenv("MMAE_CHAIN_STAGGER")) : 60;
    cp.stagger_ns = cp.m_tiles >= 8 * grid ? (unsigned)chain_stagger : 0u;
    long long* trace = trace_begin(cp);
    int pr = prof_begin(flops * (double)B);
    if (pr >= 0) { auto& R = prof_recs[pr]; R.m = B; R.n = -1; R.k = -1; R.ta = 0; R.tb = 1; R.splits = 1; }
    cudaError_t e = chain_launch(cp, grid, stream);
    prof_end(pr);
    ++launches; ++chain_launches;
    if (e != cudaSuccess) return cuda_fail(e, "chain launch");
    trace_end(cp, trace, "fwd");
    cur_emb = mu;
    d_fused = o.train_recon;
    last_gemm_tc = true;
    fill_fused = o.fill_out != nullptr;
    if (o.target) {
      if (fast_step && o.train_recon) pending_loss_partials = grid;        // summed by grad_assemble_kernel
      else { int r = reduce_partials(grid, 0); if (r) return r; }
    }
    return 1;
  }

  int forward(const FwdOpts& o) {
    const int64_t B = o.B;
    const int act = cfg.activation;
    const float* a = o.X;
    NoiseView nv = noise_view(o.noise);
    if (cfg.precision == MMAE_PREC_TF32 && B >= 32) RET(refresh_shadows());
    if (o.noise || o.train_recon || o.labels) RET(ensure_acts(B));
    // Wide first layers can go through the two-SM GEMM variant that applies the mask + noise to the A tile in shared
    // memory (forward and the layer's wgrad), so that no noisy copy of X is ever written.  Measured slower than one
    // materialising pass on this ring depth (see gemm_tc2.cu), hence opt-in (MMAE_FUSE_NOISE=1); default: materialise.
    const bool fuse_noise = o.noise && cfg.precision == MMAE_PREC_TF32 && noise_fusion_on() && B >= 256 && F >= 256 &&
                            (F & 31) == 0 && layers[0] > 256 && (layers[0] & 3) == 0;
    if (o.noise && cfg.precision == MMAE_PREC_TF32 && B >= 32 && (F & 3) == 0 && !fuse_noise) {
      // the tcgen05 family reads its operands through TMA: materialise noisy_X once (first GEMM + its wgrad)
      if (!o.noisy_ready) {
        noise_apply_kernel<<<grid_for(B * F, 256), 256, 0, stream>>>(o.X, noisy, B, F, nv);
        CKL("noise_apply");
      }
      a = noisy; nv.enabled = 0;
    }
    x_eff = a; x_noise = nv;
    int64_t lda = F;
    bool chained = false;
    if (o.decoder && !nv.enabled) { int cr = forward_chain(o, a); if (cr < 0) return cr; chained = cr == 1; }
    if (!chained) RET(ensure_acts(B));
    for (int i = 0; i < L && !chained; ++i) {
      const bool last = i == L - 1;
      const int din = enc_in(i), dout = layers[i];
      char wn[32], bn[32]; snprintf(wn, 32, "weights%d", i); snprintf(bn, 32, "encode_biases%d", i);
      NoiseView lnv = i == 0 ? nv : noise_view(false);
      if (cfg.variational && last) {
        Epilogue e = epi(EPI_BIAS_ACT); e.bias = pvar("variance_bias"); e.act = MMAE_ACT_LINEAR;
        RET(gemm(false, false, B, E, din, a, lda, pvar("variance_weights"), E, lv, E, lnv, e, nullptr, false));
      }
      Epilogue e = epi(EPI_BIAS_ACT); e.bias = pvar(bn);
      float* dst;
      if (!last) { e.act = act; if (o.keep < 1.f) set_dropout(e, o.keep, (uint32_t)i, dout); dst = ea[i]; }
      else if (cfg.classifier_only) { e.act = act; if (o.keep < 1.f) set_dropout(e, o.keep, (uint32_t)i, dout); dst = mu; }   // neural_net.py:157-167
      else { e.act = MMAE_ACT_LINEAR; dst = mu; }
      RET(gemm(false, false, B, dout, din, a, lda, pvar(wn), dout, dst, dout, lnv, e, nullptr, false));
      a = dst; lda = dout;
    }
    cur_emb = mu;
    if (cfg.variational) {
      VaeArgs va; va.mu = mu; va.lv = lv; va.eps = eps; va.emb = emb; va.kl_partials = partials;
      va.batch = B; va.row0 = first_row; va.E = E; va.step = &d_state->step; va.seed = cfg.seed; va.gen_eps = eps_injected ? 0 : 1;
      int g = grid_for(B * E, 256);
      vae_sample_kernel<<<g, 256, 0, stream>>>(va);
      CKL("vae_sample");
      RET(reduce_partials(g, 1));
      cur_emb = emb;
    }
    if (o.decoder && !chained) {
      const float* u = cur_emb; int64_t ldu = E;
      for (int j = 0; j < L; ++j) {
        const int i = L - 1 - j;
        const int din = layers[i], dout = enc_in(i);
        const bool last = j == L - 1;
        char wn[32], bn[32];
        snprintf(bn, 32, "decode_biases%d", i);
        const float* W; bool tb; int64_t ldw;
        if (cfg.tie_weights) { snprintf(wn, 32, "weights%d", i); W = pvar(wn); tb = true; ldw = din; }   // W_i^T (:284)
        else { snprintf(wn, 32, "decode_weights%d", i); W = pvar(wn); tb = false; ldw = dout; }
        Epilogue e; float* dst;
        int64_t np = 0;
        if (!last) {
          e = epi(EPI_BIAS_ACT); e.bias = pvar(bn); e.act = act;
          if (o.keep < 1.f) set_dropout(e, o.keep, 32u + (uint32_t)j, dout);
          dst = da[j];
          RET(gemm(false, tb, B, dout, din, u, ldu, W, ldw, dst, dout, noise_view(false), e, nullptr, false));
        } else {
          e = epi(o.train_recon ? EPI_LOSS_TRAIN : EPI_LOSS_PRED); e.bias = pvar(bn); e.loss = cfg.loss_func;
          e.target = o.target; e.ldt = F; e.loss_partials = o.target ? partials : nullptr;
          e.aux_rows = o.target_rows;
          if (o.train_recon) e.colsum_partials = colpart;
          dst = o.recon_out ? o.recon_out : out;
          RET(gemm(false, tb, B, dout, din, u, ldu, W, ldw, dst, dout, noise_view(false), e, &np, false));
          d_fused = o.train_recon && last_gemm_tc;
          if (o.target) RET(reduce_partials(np, 0));
        }
        u = dst; ldu = dout;
      }
    }
    if (o.headp) {
      const float* h = cur_emb; int64_t ldh = E;
      for (int i = 0; i < H; ++i) {
        const int din = i == 0 ? E : head[i - 1], dout = head[i];
        char wn[40], bn[40]; snprintf(wn, 40, "classification_weights%d", i); snprintf(bn, 40, "classification_biases%d", i);
        Epilogue e = epi(EPI_BIAS_ACT); e.bias = pvar(bn);
        const bool activated = !cfg.classifier_only && i < L - 1;   // reference quirk: bound is the AE depth (:533); the plain MLP's logits are linear
        if (activated) { e.act = cfg.head_activation; if (o.keep < 1.f) set_dropout(e, o.keep, 64u + (uint32_t)i, dout); }
        else e.act = MMAE_ACT_LINEAR;
        float* dst = (i == H - 1) ? hlogits : ha[i];
        RET(gemm(false, false, B, dout, din, h, ldh, pvar(wn), dout, dst, dout, noise_view(false), e, nullptr, false));
        h = dst; ldh = dout;
      }
    }
    return 0;
  }
  const float* x_eff = nullptr; NoiseView x_noise; const float* cur_emb = nullptr;
  bool eps_injected = false;

  int64_t gbatch(int64_t B) const { return global_batch > 0 ? global_batch : B; }

  int head_loss(const float* labels, int64_t B, bool want_delta, float* probs, int32_t* preds) {
    HeadLossArgs a; a.logits = hlogits; a.labels = labels; a.delta = want_delta ? hdelta : nullptr;
    a.probs = probs; a.preds = preds; a.partials = partials; a.batch = B; a.C = C; a.loss = cfg.head_loss;
    const double cnt = cfg.head_loss == MMAE_HEAD_SIGMOID_CE ? (double)gbatch(B) * C : (double)gbatch(B);
    a.inv_count = (float)(1.0 / cnt);
    int g = (int)std::max<int64_t>(1, std::min<int64_t>((B + 127) / 128, num_sms * 4));
    head_loss_kernel<<<g, 128, 0, stream>>>(a);
    CKL("head_loss");
    if (labels) {
      reduce_partials_kernel<<<1, 256, 0, stream>>>(partials, g, d_sums + 2, 0); CKL("reduce_partials");
      reduce_partials_kernel<<<1, 256, 0, stream>>>(partials + g, g, d_sums + 3, 0); CKL("reduce_partials");
    }
    return 0;
  }

  // ================================================================= backward
  // Encoder part shared by both optimizers.  d = dL/dmu [B,E] in `d`; g_lv in glv when variational.
  int backward_encoder(int64_t B, float* d, float keep) {
    float* other = (d == dA) ? dB : dA;
    for (int i = L - 1; i >= 0; --i) {
      const int din = enc_in(i), dout = layers[i];
      char wn[32], bn[32]; snprintf(wn, 32, "weights%d", i); snprintf(bn, 32, "encode_biases%d", i);
      const float* a_in = i == 0 ? x_eff : ea[i - 1];
      NoiseView nv = i == 0 ? x_noise : noise_view(false);
      RET(bias_grad(d, B, dout, dout, gvar(bn), d_fused));
      Epilogue ew = epi(EPI_PLAIN); ew.beta = (cfg.tie_weights && !cls_pass) ? 1.f : 0.f;   // tied: decoder part already there
      // (Measured and dropped: computing this layer's gradient in four row blocks so that the largest bucket's all-reduce
      // overlaps it -- 2.42 -> 2.54 ms per step at 8 GPUs: the narrower GEMMs quantise worse over 74 CTA pairs and
      // the NCCL kernels running beside them take SMs away.)
      RET(gemm(true, false, din, dout, B, a_in, din, d, dout, gvar(wn), dout, nv, ew, nullptr, true));
      RET(bucket_vars(wn, bn));
      if (i == 0) { RET(release_adam()); break; }
      const bool var_here = cfg.variational && i == L - 1;
      Epilogue ed = epi(var_here ? EPI_PLAIN : EPI_DGRAD);
      if (!var_here) { ed.saved = ea[i - 1]; ed.lds = din; ed.act = cfg.activation; if (keep < 1.f) set_dropout(ed, keep, (uint32_t)(i - 1), din); ed.colsum_partials = colpart; }
      RET(gemm(false, true, B, din, dout, d, dout, pvar(wn), dout, other, din, noise_view(false), ed, nullptr, false));
      d_fused = !var_here && last_gemm_tc;
      if (var_here) {
        RET(colsum(glv, B, E, E, gvar("variance_bias")));
        Epilogue ev = epi(EPI_PLAIN);
        RET(gemm(true, false, din, E, B, a_in, din, glv, E, gvar("variance_weights"), E, noise_view(false), ev, nullptr, true));
        RET(bucket_vars("variance_weights", "variance_bias"));
        Epilogue e2 = epi(EPI_DGRAD); e2.beta = 1.f; e2.saved = ea[i - 1]; e2.lds = din; e2.act = cfg.activation;
        if (keep < 1.f) set_dropout(e2, keep, (uint32_t)(i - 1), din);
        e2.colsum_partials = colpart;
        RET(gemm(false, true, B, din, E, glv, E, pvar("variance_weights"), E, other, din, noise_view(false), e2, nullptr, false));
        d_fused = last_gemm_tc;
      }
      RET(release_adam());          // the dgrads that read this layer's weights are enqueued: their update may follow the all-reduce
      std::swap(d, other);
    }
    return 0;
  }
  bool cls_pass = false;

  // Backward dgrad chain (chain_tc.cuh): every delta of the step in ONE launch, deltas resident in TMEM from layer to
  // layer.  Op k = 0..L-1 goes down the decoder (delta . D_j^T, j = L-1-k, times act'(da[j-1]); the last of them is
  // the plain g_e = dL/d(embedding)), op k = L..2L-2 down the encoder (delta . W_i^T, i = 2L-1-k, times act'(ea[i-1])).
  // Returns 1 when it ran (deltas in dch[k], bias-gradient partials in cpch[k]), 0 when the configuration does not fit.
  int64_t bchain_launches = 0;
  int backward_chain(int64_t B, float keep) {
    if (chain_mode < 0) { const char* ev = getenv("MMAE_CHAIN"); chain_mode = (ev && ev[0] == '0') ? 0 : 1; }
    static const bool bwd_off = getenv("MMAE_CHAIN_BWD") && getenv("MMAE_CHAIN_BWD")[0] == '0';
    if (!chain_mode || bwd_off || cfg.precision != MMAE_PREC_TF32 || cfg.variational || B < 32 || L < 2 || !d_fused) return 0;
    RET(ensure_bchain(B));
    const int nops = 2 * L - 1;
    std::vector<ChainLayer> ls;
    double flops = 0.0;
    for (int k = 0; k < nops; ++k) {
      ChainLayer l; memset(&l, 0, sizeof(l));
      l.ep = epi(EPI_DGRAD); l.ep.act = MMAE_ACT_LINEAR;
      char wn[32];
      if (k < L) {                       // decoder layer j = L-1-k, encoder index i = k
        const int j = L - 1 - k;
        l.K = enc_in(k); l.N = layers[k]; l.ldw = l.K;
        if (cfg.tie_weights) { snprintf(wn, 32, "weights%d", k); l.Wkm = shadowT(pvar(wn), l.K, l.N); }     // D_j = W_k^T: K-major [N, K] = shadow of W_k
        else { snprintf(wn, 32, "decode_weights%d", k); l.Wkm = pvar(wn); }                                  // D_j stored [N, K] already
        if (j > 0) {
          l.ep.saved = da[j - 1]; l.ep.lds = l.N; l.ep.act = cfg.activation;
          if (keep < 1.f) set_dropout(l.ep, keep, 32u + (uint32_t)(j - 1), l.N);
        }
      } else {                           // encoder layer i = 2L-1-k
        const int i = 2 * L - 1 - k;
        l.K = layers[i]; l.N = enc_in(i); l.ldw = l.K;
        snprintf(wn, 32, "weights%d", i); l.Wkm = pvar(wn);                                                  // W_i stored [N, K]
        l.ep.saved = ea[i - 1]; l.ep.lds = l.N; l.ep.act = cfg.activation;
        if (keep < 1.f) set_dropout(l.ep, keep, (uint32_t)(i - 1), l.N);
      }
      if (!l.Wkm || (l.K & 3) || (l.N & 3)) return 0;
      if (k == nops - 1) { l.ep.target = l.ep.saved; l.ep.ldt = l.ep.lds; }      // final op: the aux tile travels by TMA
      l.ep.colsum_partials = cpch[k];
      l.out = dch[k]; l.ldo = l.N;
      ls.push_back(l); flops += 2.0 * l.K * l.N;
    }
    ChainParams cp;
    if (!chain_build(cp, out, B, F, ls)) return 0;
    const int grid = std::min(cp.m_tiles, num_sms);
    cp.stagger_ns = 0u;
    long long* trace = trace_begin(cp);
    int pr = prof_begin(flops * (double)B);
    if (pr >= 0) { auto& R = prof_recs[pr]; R.m = B; R.n = -2; R.k = -2; R.ta = 0; R.tb = 1; R.splits = 1; }
    cudaError_t e = chain_launch(cp, grid, stream);
    prof_end(pr);
    ++launches; ++chain_launches; ++bchain_launches;
    if (e != cudaSuccess) return cuda_fail(e, "backward chain launch");
    trace_end(cp, trace, "bwd");
    return 1;
  }

  // Weight / bias gradients from the deltas of a backward-chain launch (same GEMMs, same order and the same buckets as
  // the per-layer path below).
  // All weight gradients of the step in ONE grouped tcgen05 launch (wgrad_group.cuh) + ONE assembly launch that sums
  // the split-K slices and the bias-gradient partials into G (and, on the fast train path, the loss partials and the
  // per-step scalars).  Returns 1 when it ran, 0 when a shape does not fit (the per-layer GEMMs run instead).
  int backward_group(int64_t B) {
    static const bool group_off = getenv("MMAE_WGRAD_GROUP") && getenv("MMAE_WGRAD_GROUP")[0] == '0';
    if (group_off || dp_on() || x_noise.enabled) return 0;
    if (nP > ((int64_t)1 << 21)) return 0;          // large models: their weight gradients fill the two-SM GEMM on their own
    if (!dL_fused) return 0;                        // the assembly reads every bias gradient from fused per-32-row partials
    for (char f : dch_fused) if (!f) return 0;
    const bool tied = cfg.tie_weights != 0;
    struct Item { WgDesc d; Var* v; };
    std::vector<Item> items;
    char wn[32];
    auto dec_item = [&](int i) {            // decoder layer j = L-1-i
      const int j = L - 1 - i;
      const int din = layers[i], dout = enc_in(i);
      const float* d = j == L - 1 ? out : dch[L - 2 - j];
      const float* u_in = j == 0 ? cur_emb : da[j - 1];
      Item it; memset(&it.d, 0, sizeof(it.d));
      if (tied) { snprintf(wn, 32, "weights%d", i); it.d.A = d; it.d.lda = dout; it.d.M = dout; it.d.B = u_in; it.d.ldb = din; it.d.N = din; }
      else { snprintf(wn, 32, "decode_weights%d", i); it.d.A = u_in; it.d.lda = din; it.d.M = din; it.d.B = d; it.d.ldb = dout; it.d.N = dout; }
      it.v = find(wn); items.push_back(it);
    };
    auto enc_item = [&](int i) {
      const int din = enc_in(i), dout = layers[i];
      const int k = 2 * L - 2 - i;
      Item it; memset(&it.d, 0, sizeof(it.d));
      snprintf(wn, 32, "weights%d", i);
      it.d.A = i == 0 ? x_eff : ea[i - 1]; it.d.lda = din; it.d.M = din; it.d.B = dch[k]; it.d.ldb = dout; it.d.N = dout;
      it.v = find(wn); items.push_back(it);
    };
    if (tied) { for (int i = 0; i < L; ++i) { dec_item(i); enc_item(i); } }
    else { for (int i = 0; i < L; ++i) dec_item(i); for (int i = 0; i < L; ++i) enc_item(i); }
    if ((int)items.size() > WG_MAX_PROBLEMS || (int)items.size() + 2 * L > GA_MAX_SEGS) return 0;
    std::vector<WgDesc> descs;
    for (auto& it : items) { if (!it.v || !wg_eligible(it.d)) return 0; descs.push_back(it.d); }
    const WgPlan pl = wg_plan(descs, B, num_sms);
    int64_t need = 0;
    for (auto& it : items) need += (int64_t)pl.splits * it.d.M * it.d.N;
    if (need > wg_ws_cap) {
      CK(cudaStreamSynchronize(stream)); clear_graphs();
      RET(realloc_dev(wg_ws, need)); wg_ws_cap = need;
    }
    WgParams wp; memset(&wp, 0, sizeof(wp));
    wp.nprob = (int)items.size(); wp.splits = pl.splits; wp.K = B; wp.k_per_split = pl.k_per_split;
    GaArgs ga; memset(&ga, 0, sizeof(ga));
    ga.G = G;
    int tile0 = 0, nblk = 1; int64_t off = 0;
    for (size_t q = 0; q < items.size(); ++q) {
      const WgDesc& d = items[q].d;
      WgProblem& P = wp.pr[q];
      P.M = d.M; P.N = d.N; P.m_blocks = (d.M + TC_BM - 1) / TC_BM; P.n_blocks = (d.N + WG_BN - 1) / WG_BN;
      P.tile0 = tile0; tile0 += P.m_blocks * P.n_blocks;
      P.a3d = (d.M % 32) == 0; P.b3d = (d.N % 32) == 0;
      P.ws = wg_ws + off;
      bool ok = P.a3d ? make_tmap_mn3d(&P.tmA, d.A, B, d.M, d.lda, TC_BK, TC_BM / 32) : make_tmap(&P.tmA, d.A, B, d.M, d.lda, 32, TC_BK, true);
      ok = ok && (P.b3d ? make_tmap_mn3d(&P.tmB, d.B, B, d.N, d.ldb, TC_BK, WG_BN / 32) : make_tmap(&P.tmB, d.B, B, d.N, d.ldb, 32, TC_BK, true));
      if (!ok) return 0;
      const bool first_of_var = q == 0 || items[q - 1].v != items[q].v;
      if (first_of_var) {
        GaSeg& sg = ga.seg[ga.nseg++];
        sg.g_off = items[q].v->off; sg.count = (int64_t)d.M * d.N; sg.src = P.ws; sg.nslices = pl.splits; sg.kind = 0;
        sg.block0 = nblk; sg.nblocks = (int)std::min<int64_t>(64, (sg.count + GA_THREADS - 1) / GA_THREADS); nblk += sg.nblocks;
      } else {
        ga.seg[ga.nseg - 1].nslices += pl.splits;          // tied: decoder part + encoder part, adjacent slice lists
      }
      off += (int64_t)pl.splits * d.M * d.N;
    }
    wp.tiles = tile0;
    const int groups = (int)((B + 31) / 32);
    auto bias_seg = [&](const char* name, const float* part, int n) {
      Var* v = find(name);
      GaSeg& sg = ga.seg[ga.nseg++];
      sg.g_off = v->off; sg.count = n; sg.src = part; sg.nslices = groups; sg.kind = 1;
      sg.block0 = nblk; sg.nblocks = (n + 31) / 32; nblk += sg.nblocks;
    };
    char bn[32];
    for (int i = 0; i < L; ++i) {
      const int j = L - 1 - i;
      snprintf(bn, 32, "decode_biases%d", i); bias_seg(bn, j == L - 1 ? colpart : cpch[L - 2 - j], enc_in(i));
      snprintf(bn, 32, "encode_biases%d", i); bias_seg(bn, cpch[2 * L - 2 - i], layers[i]);
    }
    double flops = 0.0;
    for (auto& it : items) flops += 2.0 * it.d.M * it.d.N * (double)B;
    int pr = prof_begin(flops);
    if (pr >= 0) { auto& R = prof_recs[pr]; R.m = -3; R.n = wp.tiles; R.k = B; R.ta = 1; R.tb = 0; R.splits = pl.splits; }
    cudaError_t e = wgrad_group_launch(wp, pl.grid, stream);
    prof_end(pr);
    ++launches; ++wgroup_launches;
    if (e != cudaSuccess) return cuda_fail(e, "grouped wgrad launch");
    if (pending_loss_partials > 0) {
      ga.loss_partials = partials; ga.n_loss_partials = (int)pending_loss_partials; ga.loss_sum_out = d_sums + 0;
      pending_loss_partials = 0;
    }
    ga.do_finalize = fast_step ? 1 : 0;
    FinalizeArgs fin = finalize_args(B, true, false, fast_step ? 0 : -1, fast_step);
    grad_assemble_kernel<<<nblk, GA_THREADS, 0, stream>>>(ga, fin);
    CKL("grad_assemble");
    if (fast_step) step_finalized = true;
    d_fused = false;
    return 1;
  }

  // One decoder (dec = true, encoder index i) or encoder layer's bias gradient, weight gradient and bucket, from the deltas a
  // backward chain launch or backward_dgrads_layerwise left in dch[] (bias-gradient partials in cpch[] where fused).
  std::vector<char> dch_fused;      // per dgrad op: its per-32-row column sums are valid in cpch[k]
  bool dL_fused = false;            // ... and delta_L's in colpart
  int wgrad_unit(bool dec, int i, int64_t B) {
    char wn[32], bn[32];
    if (dec) {
      const int j = L - 1 - i;
      const int din = layers[i], dout = enc_in(i);
      const float* d = j == L - 1 ? out : dch[L - 2 - j];
      const bool fused = j == L - 1 ? dL_fused : dch_fused[L - 2 - j] != 0;
      const float* part = j == L - 1 ? colpart : cpch[L - 2 - j];
      snprintf(bn, 32, "decode_biases%d", i);
      const float* u_in = j == 0 ? cur_emb : da[j - 1];
      RET(bias_grad(d, B, dout, dout, gvar(bn), fused, part));
      Epilogue ew = epi(EPI_PLAIN);
      if (cfg.tie_weights) {
        snprintf(wn, 32, "weights%d", i);
        RET(gemm(true, false, dout, din, B, d, dout, u_in, din, gvar(wn), din, noise_view(false), ew, nullptr, true));
        RET(bucket_vars(bn, bn));
      } else {
        snprintf(wn, 32, "decode_weights%d", i);
        RET(gemm(true, false, din, dout, B, u_in, din, d, dout, gvar(wn), dout, noise_view(false), ew, nullptr, true));
        RET(bucket_vars(wn, bn));
      }
    } else {
      const int din = enc_in(i), dout = layers[i];
      const int k = 2 * L - 2 - i;
      snprintf(wn, 32, "weights%d", i); snprintf(bn, 32, "encode_biases%d", i);
      const float* a_in = i == 0 ? x_eff : ea[i - 1];
      NoiseView nv = i == 0 ? x_noise : noise_view(false);
      RET(bias_grad(dch[k], B, dout, dout, gvar(bn), dch_fused[k] != 0, cpch[k]));
      Epilogue ew = epi(EPI_PLAIN); ew.beta = cfg.tie_weights ? 1.f : 0.f;
      RET(gemm(true, false, din, dout, B, a_in, din, dch[k], dout, gvar(wn), dout, nv, ew, nullptr, true));
      RET(bucket_vars(wn, bn));
    }
    return release_adam();          // (every dgrad of the step is already enqueued)
  }

  // big_first (data parallel): the largest buckets are computed -- and their all-reduce started -- first, the smallest
  // last, so that what is still in flight when the last weight gradient retires is a 1 MB bucket instead of weights0's
  // 33.5 MB (39 % of all gradient bytes of the wide model; 160 us in NCCL on 8 x B200).
  int backward_recon_from_chain(int64_t B, bool big_first = false) {
    { int gr = backward_group(B); if (gr < 0) return gr; if (gr == 1) return 0; }
    struct Unit { bool dec; int i; int64_t key; };
    std::vector<Unit> units;
    for (int i = 0; i < L; ++i) units.push_back({true, i, (int64_t)enc_in(i) * layers[i]});
    for (int i = L - 1; i >= 0; --i) units.push_back({false, i, (int64_t)enc_in(i) * layers[i]});
    if (big_first)      // stable: a tied layer's decoder part (same key, earlier in the list) stays ahead of its encoder part
      std::stable_sort(units.begin(), units.end(), [](const Unit& a, const Unit& b) { return a.key > b.key; });
    for (const Unit& u : units) RET(wgrad_unit(u.dec, u.i, B));
    d_fused = false;
    return 0;
  }

  // Data-parallel steps of models the chain kernel does not take: every dgrad first (the critical path to the last delta),
  // each delta kept in its own buffer, then the weight gradients in bucket-size order (backward_recon_from_chain).
  int backward_dgrads_layerwise(int64_t B, float keep) {
    RET(ensure_bchain(B));
    dch_fused.assign(2 * L - 1, 0);
    dL_fused = d_fused;
    const float* d = out; int64_t ldd = F;
    for (int k = 0; k < L; ++k) {                 // down the decoder: layer j = L-1-k, encoder index i = k
      const int j = L - 1 - k, i = k;
      const int din = layers[i], dout = enc_in(i);
      char wn[32]; snprintf(wn, 32, cfg.tie_weights ? "weights%d" : "decode_weights%d", i);
      Epilogue ed = epi(j > 0 ? EPI_DGRAD : EPI_PLAIN);
      if (j > 0) { ed.saved = da[j - 1]; ed.lds = din; ed.act = cfg.activation; if (keep < 1.f) set_dropout(ed, keep, 32u + (uint32_t)(j - 1), din); }
      ed.colsum_partials = cpch[k];
      RET(gemm(false, !cfg.tie_weights, B, din, dout, d, ldd, pvar(wn), cfg.tie_weights ? din : dout, dch[k], din, noise_view(false), ed, nullptr, false));
      dch_fused[k] = last_gemm_tc ? 1 : 0;
      d = dch[k]; ldd = din;
    }
    for (int k = L; k < 2 * L - 1; ++k) {         // down the encoder: layer i = 2L-1-k
      const int i = 2 * L - 1 - k;
      const int din = enc_in(i), dout = layers[i];
      char wn[32]; snprintf(wn, 32, "weights%d", i);
      Epilogue ed = epi(EPI_DGRAD); ed.saved = ea[i - 1]; ed.lds = din; ed.act = cfg.activation;
      if (keep < 1.f) set_dropout(ed, keep, (uint32_t)(i - 1), din);
      ed.colsum_partials = cpch[k];
      RET(gemm(false, true, B, din, dout, d, dout, pvar(wn), dout, dch[k], din, noise_view(false), ed, nullptr, false));
      dch_fused[k] = last_gemm_tc ? 1 : 0;
      d = dch[k];
    }
    return 0;
  }

  int backward_recon(int64_t B, float keep) {
    cls_pass = false;
    { int cr = backward_chain(B, keep); if (cr < 0) return cr;
      if (cr == 1) { dch_fused.assign(2 * L - 1, 1); dL_fused = true; return backward_recon_from_chain(B); } }
    static const bool reorder_off = getenv("MMAE_DP_REORDER") && getenv("MMAE_DP_REORDER")[0] == '0';
    // Per-layer dgrads first, each delta in its own buffer; then the weight gradients: bucket-size order under data
    // parallelism, the grouped launch + one assembly launch for small models whose chain does not fit TMEM (e.g. the
    // classification config [200, 100]: 12 launches per reconstruction step instead of 27).
    const bool small_tc = cfg.precision == MMAE_PREC_TF32 && B >= 32 && nP <= ((int64_t)1 << 21);
    if ((dp_on() || small_tc) && !reorder_off && !cfg.variational && L >= 2) {
      RET(backward_dgrads_layerwise(B, keep));
      return backward_recon_from_chain(B, dp_on());
    }
    float* d = out; int64_t ldd = F;       // delta_L from the EPI_LOSS_TRAIN epilogue
    float* nxt = dA;
    for (int j = L - 1; j >= 0; --j) {
      const int i = L - 1 - j;
      const int din = layers[i], dout = enc_in(i);     // decoder layer maps din -> dout
      char wn[32], bn[32]; snprintf(bn, 32, "decode_biases%d", i);
      const float* u_in = j == 0 ? cur_emb : da[j - 1];
      RET(bias_grad(d, B, dout, ldd, gvar(bn), d_fused));
      Epilogue ew = epi(EPI_PLAIN);
      if (cfg.tie_weights) {   // (dD)^T = delta^T . u accumulates into the tied encoder variable
        snprintf(wn, 32, "weights%d", i);
        RET(gemm(true, false, dout, din, B, d, ldd, u_in, din, gvar(wn), din, noise_view(false), ew, nullptr, true));
      } else {
        snprintf(wn, 32, "decode_weights%d", i);
        RET(gemm(true, false, din, dout, B, u_in, din, d, ldd, gvar(wn), dout, noise_view(false), ew, nullptr, true));
      }
      // this decoder layer's gradients are final: their all-reduce starts behind the wgrad, beside the dgrad below
      if (cfg.tie_weights) RET(bucket_vars(bn, bn)); else RET(bucket_vars(wn, bn));
      Epilogue ed = epi(j > 0 ? EPI_DGRAD : EPI_PLAIN);
      if (j > 0) { ed.saved = da[j - 1]; ed.lds = din; ed.act = cfg.activation; if (keep < 1.f) set_dropout(ed, keep, 32u + (uint32_t)(j - 1), din); }
      ed.colsum_partials = colpart;
      // g_u = delta . D^T : untied D stored [din, dout] = [N, K] -> transposed B operand; tied D^T = W_i stored [dout, din] = [K, N]
      RET(gemm(false, !cfg.tie_weights, B, din, dout, d, ldd, pvar(wn), cfg.tie_weights ? din : dout, nxt, din,
               noise_view(false), ed, nullptr, false));
      d_fused = last_gemm_tc;
      RET(release_adam());          // the dgrad that reads D_j is enqueued: the bucket's update may follow its all-reduce
      d = nxt; ldd = din; nxt = (nxt == dA) ? dB : dA;
    }
    if (cfg.variational) {
      vae_grad_kernel<<<grid_for(B * E, 256), 256, 0, stream>>>(d, glv, emb, lv, eps, B * E, (float)(1.0 / (double)gbatch(B)));
      CKL("vae_grad");
      d_fused = false;      // g_mu differs from the stored g_e
    }
    return backward_encoder(B, d, keep);
  }

  int backward_cls(int64_t B, float keep) {
    cls_pass = true;
    d_fused = false;
    float* d = hdelta; int64_t ldd = C;
    float* nxt = dA;
    if (!cfg.classifier_only && H - 1 < L - 1) {     // quirk (:533): the logits themselves went through act + dropout
      // d <- d * act'(logits) * dropmask/keep, done by a 1-column-block "GEMM-free" pass: reuse the dgrad epilogue
      Epilogue ed = epi(EPI_DGRAD); ed.saved = hlogits; ed.lds = C; ed.act = cfg.head_activation;
      if (keep < 1.f) set_dropout(ed, keep, 64u + (uint32_t)(H - 1), C);
      elementwise_epilogue_kernel<<<grid_for(B * C, 256), 256, 0, stream>>>(d, B, C, ed);
      CKL("head_logit_act_grad");
    }
    for (int i = H - 1; i >= 0; --i) {
      const int din = i == 0 ? E : head[i - 1], dout = head[i];
      char wn[40], bn[40]; snprintf(wn, 40, "classification_weights%d", i); snprintf(bn, 40, "classification_biases%d", i);
      const float* u_in = i == 0 ? cur_emb : ha[i - 1];
      RET(bias_grad(d, B, dout, ldd, gvar(bn), d_fused));
      Epilogue ew = epi(EPI_PLAIN);
      RET(gemm(true, false, din, dout, B, u_in, din, d, ldd, gvar(wn), dout, noise_view(false), ew, nullptr, true));
      RET(bucket_vars(wn, bn));
      const bool act_prev = !cfg.classifier_only && i > 0 && (i - 1) < L - 1;
      Epilogue ed = epi(act_prev ? EPI_DGRAD : EPI_PLAIN);
      if (act_prev) { ed.saved = ha[i - 1]; ed.lds = din; ed.act = cfg.head_activation; if (keep < 1.f) set_dropout(ed, keep, 64u + (uint32_t)(i - 1), din); }
      ed.colsum_partials = colpart;
      RET(gemm(false, true, B, din, dout, d, ldd, pvar(wn), dout, nxt, din, noise_view(false), ed, nullptr, false));
      d_fused = last_gemm_tc;
      RET(release_adam());
      d = nxt; ldd = din; nxt = (nxt == dA) ? dB : dA;
    }
    if (cfg.variational) {
      vae_grad_kernel<<<grid_for(B * E, 256), 256, 0, stream>>>(d, glv, emb, lv, eps, B * E, 0.f);
      CKL("vae_grad");
      d_fused = false;
    }
    if (cfg.classifier_only) {      // the last hidden layer is activated too: back through act' (and its dropout mask)
      Epilogue ed = epi(EPI_DGRAD); ed.saved = mu; ed.lds = E; ed.act = cfg.activation;
      if (keep < 1.f) set_dropout(ed, keep, (uint32_t)(L - 1), E);
      elementwise_epilogue_kernel<<<grid_for(B * E, 256), 256, 0, stream>>>(d, B, E, ed);
      CKL("embedding_act_grad");
      d_fused = false;
    }
    return backward_encoder(B, d, keep);
  }

  // ================================================================= sums -> G tail -> allreduce -> scalars
  int pack_sums() {
    pack_sums_kernel<<<1, 32, 0, stream>>>(d_sums, G + nP, 1); CKL("pack_sums"); return 0;
  }
  int unpack_sums() {
    pack_sums_kernel<<<1, 32, 0, stream>>>(d_sums, G + nP, 0); CKL("unpack_sums"); return 0;
  }
  bool dp_on() const { return comm != nullptr && world > 1; }

  // ---- data-parallel pipeline: per-bucket all-reduce, then that bucket's Adam + shadow refresh, all on comm_stream
  // while backward keeps running on `stream`.  Only the last bucket's all-reduce and update remain exposed.
  int comm_reserve = 32;          // SMs left to NCCL while buckets are in flight (= NCCL_MAX_CTAS set at comm init)
  bool dp_pipeline = false;       // this step updates each bucket right behind its all-reduce (train_core / cls_core)
  int dp_opt = 0; int64_t dp_B = 0;
  int64_t dp_covered = 0;         // parameters updated by the pipeline so far (must equal the optimizer's range at the join)
  struct PendingAdam { int64_t b, e; };
  std::vector<PendingAdam> pend_adam;
  cudaEvent_t adam_gate = nullptr;
  bool comm_busy() const { return dp_on() && buckets_in_step > 0; }
  int64_t buckets_in_step = 0;
  int reserve_credits = 0;        // GEMM launches that still run beside the bucket issued last

  // All-reduce G[begin, end) on the communication stream once everything enqueued so far on `stream` is done.
  int bucket_allreduce(int64_t begin, int64_t end) {
    if (!dp_on() || end <= begin) return 0;
    if (comm_ev_used == comm_events.size()) {
      cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); comm_events.push_back(ev);
    }
    cudaEvent_t ev = comm_events[comm_ev_used++];
    CK(cudaEventRecord(ev, stream));
    CK(cudaStreamWaitEvent(comm_stream, ev, 0));
    static const bool skip_ar = getenv("MMAE_DP_SKIP_AR") != nullptr;      // measurement only: the step without its collectives
    int r = skip_ar ? 0 : g_nccl.AllReduce(G + begin, G + begin, (size_t)(end - begin), /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, comm_stream);
    if (r != 0) return fail(MMAE_ERR_COMM, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    ++buckets_issued;
    if (end <= nP) {
      ++buckets_in_step;
      reserve_credits = (end - begin) > ((int64_t)4 << 20) ? 2 : 1;        // a > 16 MB bucket outlasts one GEMM
      if (dp_pipeline) pend_adam.push_back({begin, end});
    }
    return 0;
  }
  // bucket = the gradient range of the named variables (adjacent in the flat layout)
  int bucket_vars(const char* first, const char* last) {
    if (!dp_on()) return 0;
    Var* a = find(first); Var* b = find(last);
    if (!a || !b) return 0;
    return bucket_allreduce(a->off, b->off + align4(b->count()));
  }
  // The buckets all-reduced so far may be updated once every kernel enqueued on `stream` up to here (the dgrads that
  // still read their weights) has finished.
  int release_adam() {
    if (!dp_pipeline || pend_adam.empty()) return 0;
    if (!adam_gate) CK(cudaEventCreateWithFlags(&adam_gate, cudaEventDisableTiming));
    CK(cudaEventRecord(adam_gate, stream));
    CK(cudaStreamWaitEvent(comm_stream, adam_gate, 0));
    for (const PendingAdam& pa : pend_adam) {
      RET(adam_range(dp_opt, dp_B, pa.b, pa.e, comm_stream));
      dp_covered += pa.e - pa.b;
    }
    pend_adam.clear();
    return 0;
  }
  // loss partial sums travel as the 8-float tail of G; unpacked on the communication stream so that the pipelined
  // Adam kernels (RMSE gradient scale) see the global sums
  int sums_allreduce() {
    if (!dp_on()) return 0;
    RET(pack_sums());
    RET(bucket_allreduce(nP, nP + 8));
    if (dp_pipeline) { pack_sums_kernel<<<1, 32, 0, comm_stream>>>(d_sums, G + nP, 0); CKL("unpack_sums"); }
    return 0;
  }
  // the update (and the scalars) must see fully reduced gradients
  int join_comm() {
    if (!dp_on()) return 0;
    RET(release_adam());
    CK(cudaEventRecord(comm_done, comm_stream));
    CK(cudaStreamWaitEvent(stream, comm_done, 0));
    comm_ev_used = 0; buckets_in_step = 0;
    if (!dp_pipeline) RET(unpack_sums());
    return 0;
  }
  int allreduce_grads() { return join_comm(); }
  // fuse_opt >= 0: the kernel also advances that optimizer's step count / alpha (adam_prep) and, with fuse_advance, the
  // Philox step -- two launches fewer per train step.
  FinalizeArgs finalize_args(int64_t B, bool recon, bool headl, int fuse_opt, bool fuse_advance) const {
    FinalizeArgs a; a.sums = d_sums; a.scalars = d_scalars; a.loss = cfg.loss_func; a.variational = cfg.variational;
    a.state = (fuse_opt >= 0 || fuse_advance) ? d_state : nullptr; a.prep_opt = fuse_opt; a.advance = fuse_advance ? 1 : 0;
    a.lr = fuse_opt == 1 ? cfg.head_learning_rate : cfg.learning_rate; a.b1 = cfg.beta1; a.b2 = cfg.beta2;
    a.n_elems = (double)gbatch(B) * F; a.batch = (double)gbatch(B);
    a.head_count = cfg.head_loss == MMAE_HEAD_SIGMOID_CE ? (double)gbatch(B) * std::max(C, 1) : (double)gbatch(B);
    a.do_recon = recon ? 1 : 0; a.do_head = headl ? 1 : 0;
    return a;
  }
  int finalize_scalars(int64_t B, bool recon, bool headl, int fuse_opt = -1, bool fuse_advance = false) {
    FinalizeArgs a = finalize_args(B, recon, headl, fuse_opt, fuse_advance);
    finalize_scalars_kernel<<<1, 1, 0, stream>>>(a); CKL("finalize_scalars"); return 0;
  }

  // Adam over G / P [b, e) of optimizer `opt` on stream `st` (alpha already prepared in StepState), then the K-major
  // shadows of the 2-D variables in the range.
  int adam_range(int opt, int64_t B, int64_t b, int64_t e, cudaStream_t st) {
    AdamArgs a; a.P = P; a.G = G;
    const int64_t base = opt == 0 ? 0 : enc_begin;
    a.M = (opt == 0 ? M0 : M1) + (b - base); a.V = (opt == 0 ? V0 : V1) + (b - base);
    a.begin = b; a.end = e;
    a.segs = d_segs[opt]; a.nsegs = nsegs[opt]; a.sums = d_sums;
    a.scale_mode = (opt == 0 && cfg.loss_func == MMAE_LOSS_RMSE) ? 1 : 0;
    a.clip_norm = cfg.clip_norm;
    if (opt == 1 && cfg.clip_norm > 0.f) {        // tf.clip_by_global_norm over every gradient of the step (incl. the L2 term)
      const int g = (int)std::min<int64_t>(num_sms * 2, (e - b + 255) / 256);
      grad_sqnorm_kernel<<<g, 256, 0, st>>>(P, G, b, e, d_segs[opt], nsegs[opt], partials);
      CKL("grad_sqnorm");
      reduce_partials_kernel<<<1, 256, 0, st>>>(partials, g, d_sums + 6, 0); CKL("reduce_partials");
      a.scale_mode = 2;
    }
    a.n_elems = (double)gbatch(B) * F;
    a.alpha = &d_state->alpha[opt];
    a.b1 = cfg.beta1; a.b2 = cfg.beta2; a.eps = cfg.adam_eps; a.scalars_out = d_scalars;
    a.PT = shadow_in_adam() ? PT : nullptr;
    // large tf32 models: tiled pass that rewrites the K-major shadows too (no transpose launches at the next forward)
    static const bool tiled_off = getenv("MMAE_ADAM_TILED") && getenv("MMAE_ADAM_TILED")[0] == '0';
    if (!a.PT && !tiled_off && cfg.precision == MMAE_PREC_TF32) {
      AdamTiles g; g.n = 0; int tiles = 0; bool fits = true;
      for (size_t i = 0; i < vars.size(); ++i) {
        Var& v = vars[i];
        if (v.off < b || v.off >= e) continue;
        if (g.n == AdamTiles::kMax) { fits = false; break; }
        const int rows = v.cols > 0 ? (int)v.rows : 1, cols = v.cols > 0 ? (int)v.cols : (int)v.rows;
        g.off[g.n] = v.off; g.rows[g.n] = rows; g.cols[g.n] = cols; g.l2[g.n] = v.l2[opt]; g.shadow[g.n] = v.cols > 0 ? 1 : 0;
        g.tile0[g.n] = tiles; tiles += ((rows + 31) / 32) * ((cols + 31) / 32);
        ++g.n;
      }
      if (fits && g.n > 0) {
        g.tile0[g.n] = tiles;
        a.PT = PT;
        adam_tiled_kernel<<<tiles, dim3(32, 8), 0, st>>>(a, g);
        CKL("adam_tiled");
        for (size_t i = 0; i < vars.size(); ++i) if (vars[i].off >= b && vars[i].off < e) pt_dirty[i] = 0;
        return 0;
      }
      a.PT = nullptr;
    }
    adam_kernel<<<grid_for(e - b, 256), 256, 0, st>>>(a);
    CKL("adam");
    for (size_t i = 0; i < vars.size(); ++i) {
      Var& v = vars[i];
      if (v.off < b || v.off >= e) continue;
      if (a.PT || v.cols == 0) { pt_dirty[i] = 0; continue; }
      if (st != stream && cfg.precision == MMAE_PREC_TF32) {       // pipelined update: refresh the shadow right behind it, off the critical path
        dim3 grid((unsigned)((v.cols + 31) / 32), (unsigned)((v.rows + 31) / 32)), block(32, 8);
        transpose_kernel<<<grid, block, 0, st>>>(P + v.off, PT + v.off, (int)v.rows, (int)v.cols);
        CKL("transpose");
        pt_dirty[i] = 0;
      } else pt_dirty[i] = 1;
    }
    return 0;
  }
  int apply_update(int opt, int64_t B, bool prep_done = false) {
    if (opt == 1 && H == 0) return fail(MMAE_ERR_STATE, "no classification head");
    t_opt[opt] += 1;
    const double lr = opt == 0 ? cfg.learning_rate : cfg.head_learning_rate;
    if (!prep_done) {
      adam_prep_kernel<<<1, 1, 0, stream>>>(d_state, opt, lr, (double)cfg.beta1, (double)cfg.beta2);
      CKL("adam_prep");
    }
    return adam_range(opt, B, opt == 0 ? 0 : enc_begin, opt == 0 ? enc_end : nP, stream);
  }
  // data-parallel step: alpha is prepared up front, every bucket is updated behind its all-reduce (release_adam)
  int begin_dp_pipeline(int opt, int64_t B) {
    dp_pipeline = false;
    // Measured on 8 x B200 (wide workload, 8192 rows per rank): 2.06 ms per step with the update after the join against
    // 2.10 ms with per-bucket updates on the communication stream -- the Adam and transpose kernels queue between the
    // all-reduces and delay the buckets behind them by more than they take off the tail.  Hence opt-in.
    static const bool on = getenv("MMAE_DP_PIPELINE") && getenv("MMAE_DP_PIPELINE")[0] == '1';
    if (!dp_on() || !on) return 0;
    const double lr = opt == 0 ? cfg.learning_rate : cfg.head_learning_rate;
    adam_prep_kernel<<<1, 1, 0, stream>>>(d_state, opt, lr, (double)cfg.beta1, (double)cfg.beta2);
    CKL("adam_prep");
    dp_pipeline = true; dp_opt = opt; dp_B = B; dp_covered = 0; pend_adam.clear();
    return 0;
  }
  int end_dp_pipeline(int opt) {
    dp_pipeline = false;
    t_opt[opt] += 1;
    const int64_t want = opt == 0 ? enc_end : nP - enc_begin;
    if (dp_covered != want) return fail(MMAE_ERR_STATE, "internal: the gradient buckets of the step do not cover the optimizer's variables");
    return 0;
  }
  int64_t last_B = 0;
  int64_t pending_global = 0;
};


// =====================================================================================================
//                                             C ABI
// =====================================================================================================
namespace {
struct DeviceGuard {   // engines are bound to the device that was current at create time
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ENTER(e)                                                        \
  if (!(e)) return MMAE_ERR_INVALID;                                    \
  DeviceGuard _guard((e)->device)

int launch_noise_gen(mmae_engine* e, int64_t batch, int64_t first_row) {
  NoiseGenArgs a; memset(&a, 0, sizeof(a));
  a.zero_bits = e->zero_bits; a.mod_bits = e->mod_bits; a.batch = batch; a.row0 = first_row;
  a.num_feats = e->F; a.zw = (e->F + 31) / 32; a.n_zero = e->cfg.n_zero; a.num_mod = e->M;
  a.mode = e->cfg.noise_mode; a.num_types = (int)e->type_masks.size(); a.num_drop = e->cfg.num_modalities_to_drop;
  for (size_t i = 0; i < e->thresholds.size(); ++i) a.thresholds[i] = e->thresholds[i];
  for (size_t i = 0; i < e->type_masks.size(); ++i) a.type_masks[i] = e->type_masks[i];
  a.step = &e->d_state->step; a.seed = e->cfg.seed;
  const int wpb = 8;
  noise_gen_kernel<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, e->stream>>>(a);
  ++e->launches;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return e->cuda_fail(err, "noise_gen");
  e->noise_rows = batch;
  return 0;
}

// use_noise == 3: draw this step's descriptor inside the call.  On the tcgen05 path one kernel draws it and writes the
// noisy batch (sample_noise_kernel); elsewhere it is noise_gen_kernel followed by the operand-load application.
bool noise_materialises(const mmae_engine* e, int64_t B) {
  return e->cfg.precision == MMAE_PREC_TF32 && B >= 32 && (e->F & 3) == 0 && e->F <= SN_MAX_ZW * 32 &&
         !const_cast<mmae_engine*>(e)->noise_fusion_on();
}
int launch_sample_noise(mmae_engine* e, const float* src, uint32_t n_rows, const int64_t* idx_in, int64_t batch, int64_t first_row,
                        float* clean_out, const int64_t* view = nullptr) {
  int r = e->ensure_acts(batch); if (r) return r;
  SampleNoiseArgs a; memset(&a, 0, sizeof(a));
  a.g.zero_bits = e->zero_bits; a.g.mod_bits = e->mod_bits; a.g.batch = batch; a.g.row0 = first_row;
  a.g.num_feats = e->F; a.g.zw = (e->F + 31) / 32; a.g.n_zero = e->cfg.n_zero; a.g.num_mod = e->M;
  a.g.mode = e->cfg.noise_mode; a.g.num_types = (int)e->type_masks.size(); a.g.num_drop = e->cfg.num_modalities_to_drop;
  for (size_t i = 0; i < e->thresholds.size(); ++i) a.g.thresholds[i] = e->thresholds[i];
  for (size_t i = 0; i < e->type_masks.size(); ++i) a.g.type_masks[i] = e->type_masks[i];
  a.g.step = &e->d_state->step; a.g.seed = e->cfg.seed;
  a.src = src; a.n_rows = n_rows; a.idx_in = idx_in; a.view = view; a.idx_out = n_rows ? e->d_idx : nullptr;
  a.clean_out = clean_out; a.noisy_out = e->noisy; a.col_mod = e->d_col_mod; a.mask_with = e->cfg.mask_with;
  const int64_t blocks = std::min<int64_t>((batch + SN_WARPS - 1) / SN_WARPS, (int64_t)e->num_sms * 16);
  sample_noise_kernel<<<(unsigned)blocks, SN_WARPS * 32, 0, e->stream>>>(a);
  ++e->launches;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return e->cuda_fail(err, "sample_noise");
  e->noise_rows = batch;
  return 0;
}

int do_train(mmae_engine* e, const float* X, int64_t B, int use_noise, float keep, const float* target = nullptr, bool defer_advance = false,
             bool noisy_ready = false, const int64_t* target_rows = nullptr) {
  if (use_noise == 3 && !noisy_ready) {
    if (e->sticky) return e->fail(MMAE_ERR_CUDA, "engine is in a sticky CUDA error state: " + e->err);
    int r0 = e->ensure_cap(B); if (r0) return r0;
    if (noise_materialises(e, B)) { r0 = launch_sample_noise(e, X, 0, nullptr, B, e->first_row, nullptr); noisy_ready = true; }
    else r0 = launch_noise_gen(e, B, e->first_row);
    if (r0) return r0;
  }
  int r = e->begin_step(B, use_noise != 0); if (r) return r;
  mmae_engine::FwdOpts o; o.X = X; o.target = target ? target : X; o.labels = nullptr; o.B = B; o.noise = use_noise != 0; o.keep = keep;
  o.noisy_ready = noisy_ready;
  o.train_recon = true; o.decoder = true; o.headp = false; o.recon_out = nullptr;
  o.target_rows = target_rows;
  r = e->forward(o); if (r) return r;
  r = e->sums_allreduce(); if (r) return r;          // overlaps the whole backward pass
  r = e->backward_recon(B, keep); if (r) return r;
  e->last_B = B;
  return e->advance_step(!defer_advance);
}

int do_cls(mmae_engine* e, const float* X, const float* Y, int64_t B, int use_noise, float keep, bool defer_advance = false,
           bool noisy_ready = false) {
  if (e->H == 0) return e->fail(MMAE_ERR_STATE, "engine was created without a classification head");
  if (use_noise == 3 && !noisy_ready) {
    if (e->sticky) return e->fail(MMAE_ERR_CUDA, "engine is in a sticky CUDA error state: " + e->err);
    int r0 = e->ensure_cap(B); if (r0) return r0;
    if (noise_materialises(e, B)) { r0 = launch_sample_noise(e, X, 0, nullptr, B, e->first_row, nullptr); noisy_ready = true; }
    else r0 = launch_noise_gen(e, B, e->first_row);
    if (r0) return r0;
  }
  int r = e->begin_step(B, use_noise != 0); if (r) return r;
  mmae_engine::FwdOpts o; o.X = X; o.target = nullptr; o.labels = Y; o.B = B; o.noise = use_noise != 0; o.keep = keep;
  o.noisy_ready = noisy_ready;
  o.train_recon = false; o.decoder = false; o.headp = true; o.recon_out = nullptr;
  r = e->forward(o); if (r) return r;
  r = e->head_loss(Y, B, true, nullptr, nullptr); if (r) return r;
  r = e->sums_allreduce(); if (r) return r;
  r = e->backward_cls(B, keep); if (r) return r;
  e->last_B = B;
  return e->advance_step(!defer_advance);
}

// stage a host batch into the double-buffered device input; returns the device pointers
int stage_host(mmae_engine* e, const float* X_host, const float* Y_host, int64_t B, int ycols, float** Xd, float** Yd) {
  int r = e->ensure_host(B); if (r) return r;
  const int t = e->xin_turn; e->xin_turn ^= 1;
  cudaError_t ce;
  // the copy may start once the compute that last read this buffer has finished
  if ((ce = cudaStreamWaitEvent(e->copy_stream, e->xin_free[t], 0)) != cudaSuccess) return e->cuda_fail(ce, "wait xin_free");
  if ((ce = cudaMemcpyAsync(e->xin[t], X_host, (size_t)B * e->F * 4, cudaMemcpyHostToDevice, e->copy_stream)) != cudaSuccess)
    return e->cuda_fail(ce, "H2D X");
  if (Y_host && ycols > 0)
    if ((ce = cudaMemcpyAsync(e->yin[t], Y_host, (size_t)B * ycols * 4, cudaMemcpyHostToDevice, e->copy_stream)) != cudaSuccess)
      return e->cuda_fail(ce, "H2D Y");
  if ((ce = cudaEventRecord(e->xin_ready[t], e->copy_stream)) != cudaSuccess) return e->cuda_fail(ce, "record xin_ready");
  if ((ce = cudaStreamWaitEvent(e->stream, e->xin_ready[t], 0)) != cudaSuccess) return e->cuda_fail(ce, "wait xin_ready");
  *Xd = e->xin[t]; if (Yd) *Yd = e->yin[t];
  return t;
}
int release_stage(mmae_engine* e, int t) {
  cudaError_t ce = cudaEventRecord(e->xin_free[t], e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "record xin_free");
  return 0;
}
}  // namespace

extern "C" {

int mmae_create(const mmae_config* cfg, mmae_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return MMAE_ERR_INVALID; }
  mmae_engine* e = new mmae_engine();
  int r = e->build(cfg);
  if (r != 0) { g_create_error = e->err; e->release(); delete e; *out = nullptr; return r; }
  *out = e;
  return MMAE_OK;
}

void mmae_destroy(mmae_engine* e) {
  if (!e) return;
  DeviceGuard g(e->device);
  cudaStreamSynchronize(e->stream);
  cudaStreamSynchronize(e->copy_stream);
  e->release();
  delete e;
}

const char* mmae_last_error(const mmae_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int mmae_set_stream(mmae_engine* e, void* s) { ENTER(e); e->clear_graphs(); e->stream = (cudaStream_t)s; return 0; }

int mmae_synchronize(mmae_engine* e) {
  ENTER(e);
  cudaError_t ce = cudaStreamSynchronize(e->copy_stream);
  if (ce == cudaSuccess && e->gstream) ce = cudaStreamSynchronize(e->gstream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "synchronize");
  return 0;
}

int mmae_num_variables(const mmae_engine* e) { return e ? (int)e->vars.size() : 0; }

int mmae_variable_info(const mmae_engine* e, int index, char* name_out, int name_cap, int64_t* rows, int64_t* cols) {
  if (!e || index < 0 || index >= (int)e->vars.size()) return MMAE_ERR_INVALID;
  const Var& v = e->vars[index];
  if (name_out && name_cap > 0) { strncpy(name_out, v.name.c_str(), name_cap - 1); name_out[name_cap - 1] = 0; }
  if (rows) *rows = v.rows;
  if (cols) *cols = v.cols;
  return 0;
}

static int var_copy(mmae_engine* e, const char* name, float* base, int64_t base_off, float* host, int64_t count, bool to_dev) {
  Var* v = e->find(name);
  if (!v) return e->fail(MMAE_ERR_NOTFOUND, std::string("unknown variable ") + name);
  if (count != v->count()) return e->fail(MMAE_ERR_INVALID, std::string("size mismatch for ") + name);
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce == cudaSuccess)
    ce = to_dev ? cudaMemcpy(base + v->off - base_off, host, count * 4, cudaMemcpyHostToDevice)
                : cudaMemcpy(host, base + v->off - base_off, count * 4, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "variable copy");
  return 0;
}

int mmae_set_variable(mmae_engine* e, const char* name, const float* host, int64_t count) {
  ENTER(e);
  int r = var_copy(e, name, e->P, 0, const_cast<float*>(host), count, true);
  if (r == 0) {
    Var* v = e->find(name);
    e->mark_dirty(v->off, v->off + 1);
    if (e->adam_keeps_shadows() && v->cols > 0) r = e->refresh_shadows();      // keep "shadows are current outside a step" true
  }
  return r;
}
int mmae_get_variable(mmae_engine* e, const char* name, float* host, int64_t count) {
  ENTER(e); return var_copy(e, name, e->P, 0, host, count, false);
}
int mmae_get_gradient(mmae_engine* e, const char* name, float* host, int64_t count) {
  ENTER(e); return var_copy(e, name, e->G, 0, host, count, false);
}

static int opt_check(mmae_engine* e, int opt, const char* name) {
  Var* v = e->find(name);
  if (!v) return e->fail(MMAE_ERR_NOTFOUND, std::string("unknown variable ") + name);
  if (opt == 0 && v->group == 2) return e->fail(MMAE_ERR_INVALID, "optimizer 0 does not own head variables");
  if (opt == 1 && (v->group == 0 || e->H == 0)) return e->fail(MMAE_ERR_INVALID, "optimizer 1 does not own decoder variables");
  if (opt != 0 && opt != 1) return e->fail(MMAE_ERR_INVALID, "optimizer must be 0 or 1");
  return 0;
}
int mmae_get_opt_state(mmae_engine* e, int opt, const char* name, float* m_host, float* v_host, int64_t count, int64_t* t) {
  ENTER(e);
  int r = opt_check(e, opt, name); if (r) return r;
  const int64_t base = opt == 0 ? 0 : e->enc_begin;
  if (m_host) { r = var_copy(e, name, opt == 0 ? e->M0 : e->M1, base, m_host, count, false); if (r) return r; }
  if (v_host) { r = var_copy(e, name, opt == 0 ? e->V0 : e->V1, base, v_host, count, false); if (r) return r; }
  if (t) *t = e->t_opt[opt];
  return 0;
}
int mmae_set_opt_state(mmae_engine* e, int opt, const char* name, const float* m_host, const float* v_host, int64_t count, int64_t t) {
  ENTER(e);
  int r = opt_check(e, opt, name); if (r) return r;
  const int64_t base = opt == 0 ? 0 : e->enc_begin;
  if (m_host) { r = var_copy(e, name, opt == 0 ? e->M0 : e->M1, base, const_cast<float*>(m_host), count, true); if (r) return r; }
  if (v_host) { r = var_copy(e, name, opt == 0 ? e->V0 : e->V1, base, const_cast<float*>(v_host), count, true); if (r) return r; }
  if (t >= 0) return e->set_t(opt, t);
  return 0;
}

int mmae_set_rng_step(mmae_engine* e, uint64_t step) { ENTER(e); return e->set_step(step); }

int mmae_set_noise(mmae_engine* e, const uint32_t* zero_bits_host, const uint32_t* mod_bits_host, int64_t batch) {
  ENTER(e);
  if (!zero_bits_host || !mod_bits_host || batch <= 0) return e->fail(MMAE_ERR_INVALID, "null descriptor");
  int r = e->ensure_cap(batch); if (r) return r;
  const int zw = (e->F + 31) / 32;
  // The caller's arrays may be pageable and reused right away: copy them into one of two pinned staging buffers and
  // let the H2D run asynchronously on the engine's stream (a per-step stream synchronisation here would serialise
  // the host-side noise drawing of step s+1 with the device work of step s).
  const int64_t words = batch * (int64_t)(zw + 1);
  cudaError_t ce = cudaSuccess;
  if (words > e->h_noise_cap) {
    ce = cudaStreamSynchronize(e->stream);
    for (int i = 0; i < 2 && ce == cudaSuccess; ++i) {
      if (e->h_noise[i]) cudaFreeHost(e->h_noise[i]);
      ce = cudaMallocHost(&e->h_noise[i], (size_t)words * 4 * 2);
      if (ce == cudaSuccess && !e->h_noise_done[i]) ce = cudaEventCreateWithFlags(&e->h_noise_done[i], cudaEventDisableTiming);
    }
    if (ce != cudaSuccess) return e->cuda_fail(ce, "set_noise staging");
    e->h_noise_cap = words * 2;
  }
  const int t = e->h_noise_turn; e->h_noise_turn ^= 1;
  ce = cudaEventSynchronize(e->h_noise_done[t]);            // the copy that last used this staging buffer is done
  if (ce != cudaSuccess) return e->cuda_fail(ce, "set_noise wait");
  uint32_t* hz = e->h_noise[t]; uint32_t* hm = hz + batch * zw;
  memcpy(hz, zero_bits_host, (size_t)batch * zw * 4);
  memcpy(hm, mod_bits_host, (size_t)batch * 4);
  ce = cudaMemcpyAsync(e->zero_bits, hz, (size_t)batch * zw * 4, cudaMemcpyHostToDevice, e->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->mod_bits, hm, (size_t)batch * 4, cudaMemcpyHostToDevice, e->stream);
  if (ce == cudaSuccess) ce = cudaEventRecord(e->h_noise_done[t], e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "set_noise");
  e->noise_rows = batch;
  return 0;
}

int mmae_gen_noise(mmae_engine* e, int64_t batch, int64_t first_row) {
  ENTER(e);
  if (batch <= 0) return e->fail(MMAE_ERR_INVALID, "batch must be > 0");
  int r = e->ensure_cap(batch); if (r) return r;
  return launch_noise_gen(e, batch, first_row);
}

int mmae_get_noise(mmae_engine* e, uint32_t* zero_bits_host, uint32_t* mod_bits_host, int64_t batch) {
  ENTER(e);
  if (batch > e->noise_rows) return e->fail(MMAE_ERR_STATE, "descriptor has fewer rows");
  const int zw = (e->F + 31) / 32;
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce == cudaSuccess && zero_bits_host) ce = cudaMemcpy(zero_bits_host, e->zero_bits, (size_t)batch * zw * 4, cudaMemcpyDeviceToHost);
  if (ce == cudaSuccess && mod_bits_host) ce = cudaMemcpy(mod_bits_host, e->mod_bits, (size_t)batch * 4, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "get_noise");
  return 0;
}

int mmae_apply_noise(mmae_engine* e, const float* X_dev, int64_t batch, float* out_dev) {
  ENTER(e);
  if (batch > e->noise_rows) return e->fail(MMAE_ERR_STATE, "descriptor has fewer rows than the batch");
  noise_apply_kernel<<<e->grid_for(batch * e->F, 256), 256, 0, e->stream>>>(X_dev, out_dev, batch, e->F, e->noise_view(true));
  ++e->launches;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return e->cuda_fail(ce, "noise_apply");
  return 0;
}

int mmae_forward(mmae_engine* e, const float* X_dev, const float* target_dev, const float* labels_dev, int64_t batch,
                 int use_noise, float keep, uint32_t want, const mmae_outputs* out) {
  ENTER(e);
  if (!X_dev) return e->fail(MMAE_ERR_INVALID, "null X");
  const bool need_head = (want & (MMAE_WANT_HEAD | MMAE_WANT_HEAD_LOSS)) != 0;
  if (need_head && e->H == 0) return e->fail(MMAE_ERR_STATE, "engine was created without a classification head");
  if ((want & MMAE_WANT_HEAD_LOSS) && !labels_dev) return e->fail(MMAE_ERR_INVALID, "head loss needs labels");
  const bool need_dec = (want & (MMAE_WANT_RECON | MMAE_WANT_LOSS | MMAE_WANT_FILLED)) != 0;
  int r = e->begin_step(batch, use_noise != 0); if (r) return r;
  mmae_engine::FwdOpts o; o.X = X_dev; o.B = batch; o.noise = use_noise != 0; o.keep = keep;
  o.target = (want & MMAE_WANT_LOSS) ? (target_dev ? target_dev : X_dev) : nullptr;
  o.labels = labels_dev; o.train_recon = false; o.decoder = need_dec; o.headp = need_head;
  o.recon_out = (out && (want & MMAE_WANT_RECON)) ? out->recon : nullptr;
  o.need_mu = need_head || (want & MMAE_WANT_EMBEDDING) != 0;
  e->fill_fused = false;
  const bool want_fill = (want & MMAE_WANT_FILLED) && out && out->filled;
  const bool fuse_fill = want_fill && !(want & MMAE_WANT_RECON) && !use_noise;
  if (fuse_fill) o.fill_out = out->filled;      // detection + select both inside the whole-network kernel when it applies
  if (want_fill && !fuse_fill) {       // missing-block detection (data_funcs.py:366-381) as its own pass
    const int wpb = 8;
    const unsigned mb_grid = (unsigned)std::min<int64_t>((batch + wpb - 1) / wpb, (int64_t)e->num_sms * 8);
    missing_bits_kernel<<<mb_grid, wpb * 32, 0, e->stream>>>(X_dev, batch, e->F, e->d_starts, e->M, e->miss_bits);
    ++e->launches;
  }
  r = e->forward(o); if (r) return r;
  cudaError_t ce = cudaSuccess;
  if ((want & MMAE_WANT_EMBEDDING) && out && out->embedding)
    ce = cudaMemcpyAsync(out->embedding, e->cur_emb, (size_t)batch * e->E * 4, cudaMemcpyDeviceToDevice, e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "embedding copy");
  if (want_fill && !e->fill_fused) {
    const float* rec = o.recon_out ? o.recon_out : e->out;
    if (fuse_fill) {        // the whole-network kernel did not apply: detect the missing blocks now
      const int wpb = 8;
      const unsigned mb_grid = (unsigned)std::min<int64_t>((batch + wpb - 1) / wpb, (int64_t)e->num_sms * 8);
      missing_bits_kernel<<<mb_grid, wpb * 32, 0, e->stream>>>(X_dev, batch, e->F, e->d_starts, e->M, e->miss_bits);
      ++e->launches;
    }
    fill_select_kernel<<<e->grid_for(batch * e->F, 256), 256, 0, e->stream>>>(X_dev, rec, e->miss_bits, e->d_col_mod, out->filled, batch, e->F);
    ++e->launches;
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return e->cuda_fail(ce, "fill-in");
  }
  if (need_head) {
    r = e->head_loss((want & MMAE_WANT_HEAD_LOSS) ? labels_dev : nullptr, batch, false,
                     out ? out->probs : nullptr, out ? out->preds : nullptr);
    if (r) return r;
    if (out && out->logits) {
      ce = cudaMemcpyAsync(out->logits, e->hlogits, (size_t)batch * e->C * 4, cudaMemcpyDeviceToDevice, e->stream);
      if (ce != cudaSuccess) return e->cuda_fail(ce, "logits copy");
    }
  }
  r = e->finalize_scalars(batch, (want & MMAE_WANT_LOSS) != 0, (want & MMAE_WANT_HEAD_LOSS) != 0); if (r) return r;
  if (use_noise || keep < 1.f || e->cfg.variational) { r = e->advance_step(); if (r) return r; }
  e->last_B = batch;
  return 0;
}

int mmae_backward(mmae_engine* e, const float* X_dev, int64_t batch, int64_t global_batch, int use_noise, float keep) {
  ENTER(e);
  const int64_t saved = e->global_batch;
  if (global_batch > 0) e->global_batch = global_batch;
  int r = do_train(e, X_dev, batch, use_noise, keep);
  if (r == 0) r = e->pack_sums();
  e->global_batch = saved;
  if (global_batch > 0) e->pending_global = global_batch; else e->pending_global = 0;
  return r;
}

int mmae_grad_buffer(mmae_engine* e, float** dev_ptr, int64_t* count) {
  ENTER(e);
  if (dev_ptr) *dev_ptr = e->G;
  if (count) *count = e->nP + 8;
  return 0;
}

int mmae_apply_update(mmae_engine* e, int optimizer) {
  ENTER(e);
  const int64_t saved = e->global_batch;
  if (e->pending_global > 0) e->global_batch = e->pending_global;
  int r = e->unpack_sums();
  if (r == 0) r = e->finalize_scalars(e->last_B, optimizer == 0, optimizer == 1);
  if (r == 0) r = e->apply_update(optimizer, e->last_B);
  e->global_batch = saved;
  return r;
}

namespace {
int train_core(mmae_engine* e, const float* Xd, const float* target, int64_t batch, int use_noise, float keep, bool noisy_ready = false,
               const int64_t* target_rows = nullptr) {
  e->fast_step = !e->dp_on(); e->step_finalized = false; e->pending_loss_partials = 0;
  int r = e->begin_dp_pipeline(0, batch); if (r) return r;
  const bool piped = e->dp_pipeline;
  r = do_train(e, Xd, batch, use_noise, keep, target, true, noisy_ready, target_rows);
  e->fast_step = false;
  if (r) { e->dp_pipeline = false; return r; }
  r = e->flush_pending_loss(); if (r) return r;
  r = e->allreduce_grads(); if (r) return r;
  if (piped) {        // alpha was prepared up front and every bucket is already updated: only the scalars and the step counter remain
    r = e->finalize_scalars(batch, true, false, -1, true); if (r) return r;
    return e->end_dp_pipeline(0);
  }
  if (!e->step_finalized) { r = e->finalize_scalars(batch, true, false, 0, true); if (r) return r; }
  return e->apply_update(0, batch, true);
}
int cls_core(mmae_engine* e, const float* Xd, const float* Yd, int64_t batch, int use_noise, float keep, bool noisy_ready = false) {
  int r = e->begin_dp_pipeline(1, batch); if (r) return r;
  const bool piped = e->dp_pipeline;
  r = do_cls(e, Xd, Yd, batch, use_noise, keep, true, noisy_ready);
  if (r) { e->dp_pipeline = false; return r; }
  r = e->allreduce_grads(); if (r) return r;
  if (piped) {
    r = e->finalize_scalars(batch, false, true, -1, true); if (r) return r;
    return e->end_dp_pipeline(1);
  }
  r = e->finalize_scalars(batch, false, true, 1, true); if (r) return r;
  return e->apply_update(1, batch, true);
}
mmae_engine::GraphKey graph_key(mmae_engine* e, int kind, const void* X, const void* Y, const void* T, int64_t B, int noise, float keep) {
  mmae_engine::GraphKey k; memset(&k, 0, sizeof(k));
  k.kind = kind; k.X = X; k.Y = Y; k.T = T; k.B = B; k.noise = noise | (e->eps_injected ? 2 : 0); k.keep = keep; k.gb = e->global_batch; k.fr = e->first_row;
  k.stream = (void*)e->stream;
  return k;
}
int train_graphed(mmae_engine* e, const float* Xd, const float* target, int64_t batch, int use_noise, float keep) {
  if (e->sticky) return e->fail(MMAE_ERR_CUDA, "engine is in a sticky CUDA error state: " + e->err);
  if (use_noise && use_noise != 3 && e->noise_rows < batch) return train_core(e, Xd, target, batch, use_noise, keep);   // reports the state error
  return e->run_graphed(graph_key(e, 0, Xd, nullptr, target, batch, use_noise, keep), 0, batch,
                        [&] { return train_core(e, Xd, target, batch, use_noise, keep); });
}
int cls_graphed(mmae_engine* e, const float* Xd, const float* Yd, int64_t batch, int use_noise, float keep) {
  if (e->H == 0) return e->fail(MMAE_ERR_STATE, "engine was created without a classification head");
  if (e->sticky) return e->fail(MMAE_ERR_CUDA, "engine is in a sticky CUDA error state: " + e->err);
  if (use_noise && use_noise != 3 && e->noise_rows < batch) return cls_core(e, Xd, Yd, batch, use_noise, keep);
  return e->run_graphed(graph_key(e, 1, Xd, Yd, nullptr, batch, use_noise, keep), 1, batch,
                        [&] { return cls_core(e, Xd, Yd, batch, use_noise, keep); });
}
}  // namespace

int mmae_train_step(mmae_engine* e, const float* X_dev, int64_t batch, int use_noise, float keep) {
  ENTER(e);
  return train_graphed(e, X_dev, nullptr, batch, use_noise, keep);
}

int mmae_train_step_pair(mmae_engine* e, const float* X_in_dev, const float* target_dev, int64_t batch, int use_noise, float keep) {
  ENTER(e);
  return train_graphed(e, X_in_dev, target_dev, batch, use_noise, keep);
}

int mmae_cls_train_step(mmae_engine* e, const float* X_dev, const float* labels_dev, int64_t batch, int use_noise, float keep) {
  ENTER(e);
  return cls_graphed(e, X_dev, labels_dev, batch, use_noise, keep);
}

int mmae_train_step_host(mmae_engine* e, const float* X_host, int64_t batch, int gen_noise, float keep) {
  ENTER(e);
  float* Xd = nullptr;
  int t = stage_host(e, X_host, nullptr, batch, 0, &Xd, nullptr); if (t < 0) return t;
  // 1: this step's descriptor is drawn inside the step;  2: descriptor from mmae_set_noise
  int r = train_graphed(e, Xd, nullptr, batch, gen_noise == 1 ? 3 : (gen_noise ? 1 : 0), keep); if (r) return r;
  return release_stage(e, t);
}

int mmae_cls_train_step_host(mmae_engine* e, const float* X_host, const float* labels_host, int64_t batch, int gen_noise, float keep) {
  ENTER(e);
  if (e->H == 0) return e->fail(MMAE_ERR_STATE, "engine was created without a classification head");
  float *Xd = nullptr, *Yd = nullptr;
  const int ycols = e->cfg.head_loss == MMAE_HEAD_SIGMOID_CE ? e->C : 1;
  int t = stage_host(e, X_host, labels_host, batch, ycols, &Xd, &Yd); if (t < 0) return t;
  int r = cls_graphed(e, Xd, Yd, batch, gen_noise == 1 ? 3 : (gen_noise ? 1 : 0), keep); if (r) return r;
  return release_stage(e, t);
}

namespace {
// predict() / fill-in over a large host matrix: rows travel in chunks through a three-stage pipeline -- H2D on the copy
// stream, the forward pass on the engine's stream, D2H on a third stream -- so that with pinned host buffers the pass
// runs at PCIe speed in both directions at once instead of copy, compute, copy back to back.
constexpr int64_t kPipeChunk = 131072;
int forward_host_pipelined(mmae_engine* e, const float* X_host, int64_t batch, float keep, uint32_t want, const mmae_outputs* oh) {
  const int64_t chunk = kPipeChunk;
  int r = e->ensure_host(chunk); if (r) return r;
  cudaError_t ce = cudaSuccess;
  if (!e->d2h_stream) {
    ce = cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && ce == cudaSuccess; ++i) {
      ce = cudaEventCreateWithFlags(&e->pipe_ready[i], cudaEventDisableTiming);
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->pipe_free[i], cudaEventDisableTiming);
    }
    if (ce != cudaSuccess) return e->cuda_fail(ce, "pipeline setup");
  }
  if (e->pipe_cap < chunk) {
    for (int i = 0; i < 2; ++i) {
      const int64_t widths[3] = {e->F, e->F, e->E};
      for (int j = 0; j < 3; ++j) {
        if (e->pipe_out[i][j]) { cudaFree(e->pipe_out[i][j]); e->pipe_out[i][j] = nullptr; }
        ce = cudaMalloc(&e->pipe_out[i][j], (size_t)chunk * widths[j] * 4);
        if (ce != cudaSuccess) return e->cuda_fail(ce, "pipeline staging");
      }
    }
    e->pipe_cap = chunk;
  }
  int n = 0;
  for (int64_t r0 = 0; r0 < batch; r0 += chunk, ++n) {
    const int64_t rows = std::min(chunk, batch - r0);
    const int t = n & 1;
    // H2D of this chunk as soon as the compute that last read the slot is done
    if ((ce = cudaStreamWaitEvent(e->copy_stream, e->xin_free[t], 0)) != cudaSuccess) return e->cuda_fail(ce, "pipe wait");
    if ((ce = cudaMemcpyAsync(e->xin[t], X_host + r0 * e->F, (size_t)rows * e->F * 4, cudaMemcpyHostToDevice, e->copy_stream)) != cudaSuccess)
      return e->cuda_fail(ce, "pipe H2D");
    cudaEventRecord(e->xin_ready[t], e->copy_stream);
    cudaStreamWaitEvent(e->stream, e->xin_ready[t], 0);
    cudaStreamWaitEvent(e->stream, e->pipe_free[t], 0);        // the D2H that last read this slot's outputs is done
    mmae_outputs od; memset(&od, 0, sizeof(od));
    if (want & MMAE_WANT_RECON) od.recon = e->pipe_out[t][0];
    if (want & MMAE_WANT_FILLED) od.filled = e->pipe_out[t][1];
    if (want & MMAE_WANT_EMBEDDING) od.embedding = e->pipe_out[t][2];
    r = mmae_forward(e, e->xin[t], nullptr, nullptr, rows, 0, keep, want, &od); if (r) return r;
    cudaEventRecord(e->xin_free[t], e->stream);
    cudaEventRecord(e->pipe_ready[t], e->stream);
    cudaStreamWaitEvent(e->d2h_stream, e->pipe_ready[t], 0);
    if ((want & MMAE_WANT_RECON) && oh->recon)
      ce = cudaMemcpyAsync((float*)oh->recon + r0 * e->F, od.recon, (size_t)rows * e->F * 4, cudaMemcpyDeviceToHost, e->d2h_stream);
    if (ce == cudaSuccess && (want & MMAE_WANT_FILLED) && oh->filled)
      ce = cudaMemcpyAsync((float*)oh->filled + r0 * e->F, od.filled, (size_t)rows * e->F * 4, cudaMemcpyDeviceToHost, e->d2h_stream);
    if (ce == cudaSuccess && (want & MMAE_WANT_EMBEDDING) && oh->embedding)
      ce = cudaMemcpyAsync((float*)oh->embedding + r0 * e->E, od.embedding, (size_t)rows * e->E * 4, cudaMemcpyDeviceToHost, e->d2h_stream);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "pipe D2H");
    cudaEventRecord(e->pipe_free[t], e->d2h_stream);
  }
  ce = cudaStreamSynchronize(e->d2h_stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "pipelined forward");
  return 0;
}
}  // namespace

int mmae_forward_host(mmae_engine* e, const float* X_host, const float* target_host, const float* labels_host,
                      int64_t batch, int use_noise, float keep, uint32_t want, const mmae_outputs* oh) {
  ENTER(e);
  if (!X_host) return e->fail(MMAE_ERR_INVALID, "null X");
  const uint32_t rowwise = MMAE_WANT_RECON | MMAE_WANT_FILLED | MMAE_WANT_EMBEDDING;
  if (batch > kPipeChunk && oh && want != 0 && (want & ~rowwise) == 0 && !use_noise && !target_host && !labels_host &&
      !e->cfg.variational)
    return forward_host_pipelined(e, X_host, batch, keep, want, oh);
  int r = e->ensure_host(batch); if (r) return r;
  r = e->ensure_acts(batch); if (r) return r;      // output staging borrows the delta / noisy workspaces
  const int ycols = e->H ? (e->cfg.head_loss == MMAE_HEAD_SIGMOID_CE ? e->C : 1) : 0;
  float *Xd = nullptr, *Yd = nullptr;
  int t = stage_host(e, X_host, labels_host, batch, ycols, &Xd, &Yd); if (t < 0) return t;
  const float* Td = nullptr;
  cudaError_t ce;
  if (target_host && target_host != X_host) {      // separate true_X feed: stage it in the other input buffer
    ce = cudaMemcpyAsync(e->xin[t ^ 1], target_host, (size_t)batch * e->F * 4, cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "H2D target");
    Td = e->xin[t ^ 1];
  }
  // device-side outputs live in engine workspaces, then travel back
  mmae_outputs od; memset(&od, 0, sizeof(od));
  if (oh) {
    if (oh->recon) od.recon = e->out;
    if (oh->embedding) od.embedding = e->dA;
    if (oh->logits) od.logits = e->hdelta;
    if (oh->probs) od.probs = e->hprobs;
    if (oh->preds) od.preds = e->hpreds;
    if (oh->filled) od.filled = e->noisy;
  }
  r = mmae_forward(e, Xd, Td, labels_host ? Yd : nullptr, batch, use_noise, keep, want, &od); if (r) return r;
  r = release_stage(e, t); if (r) return r;
  auto back = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
    return (h && d) ? cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, e->stream) : cudaSuccess;
  };
  ce = cudaSuccess;
  if (oh) {
    const size_t pred_count = e->cfg.head_loss == MMAE_HEAD_SIGMOID_CE ? (size_t)batch * e->C : (size_t)batch;
    if (ce == cudaSuccess && (want & MMAE_WANT_RECON)) ce = back(oh->recon, od.recon, (size_t)batch * e->F * 4);
    if (ce == cudaSuccess && (want & MMAE_WANT_EMBEDDING)) ce = back(oh->embedding, od.embedding, (size_t)batch * e->E * 4);
    if (ce == cudaSuccess && (want & MMAE_WANT_FILLED)) ce = back(oh->filled, od.filled, (size_t)batch * e->F * 4);
    if (ce == cudaSuccess && e->H) ce = back(oh->logits, od.logits, (size_t)batch * e->C * 4);
    if (ce == cudaSuccess && e->H) ce = back(oh->probs, od.probs, (size_t)batch * e->C * 4);
    if (ce == cudaSuccess && e->H) ce = back(oh->preds, od.preds, pred_count * 4);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "D2H outputs");
  return 0;
}

int mmae_set_dataset(mmae_engine* e, int slot, const float* X_host, const float* Y_host, int64_t rows, int32_t label_cols) {
  ENTER(e);
  if (slot < 0 || slot > 1 || !X_host || rows <= 0) return e->fail(MMAE_ERR_INVALID, "bad dataset arguments");
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "sync");
  if (e->ds_X[slot]) { cudaFree(e->ds_X[slot]); e->ds_X[slot] = nullptr; }
  if (e->ds_Y[slot]) { cudaFree(e->ds_Y[slot]); e->ds_Y[slot] = nullptr; }
  if (e->ds_view[slot]) { cudaFree(e->ds_view[slot]); e->ds_view[slot] = nullptr; }
  e->ds_view_rows[slot] = 0;
  e->clear_graphs();
  ce = cudaMalloc(&e->ds_X[slot], (size_t)rows * e->F * 4);
  if (ce == cudaSuccess) ce = cudaMemcpy(e->ds_X[slot], X_host, (size_t)rows * e->F * 4, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && Y_host && label_cols > 0) {
    ce = cudaMalloc(&e->ds_Y[slot], (size_t)rows * label_cols * 4);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->ds_Y[slot], Y_host, (size_t)rows * label_cols * 4, cudaMemcpyHostToDevice);
  }
  if (ce != cudaSuccess) return e->cuda_fail(ce, "set_dataset");
  e->ds_rows[slot] = rows; e->ds_ycols[slot] = (Y_host && label_cols > 0) ? label_cols : 0;
  return 0;
}

int mmae_set_dataset_device(mmae_engine* e, int slot, const float* X_dev, const float* Y_dev, int64_t rows, int32_t label_cols) {
  ENTER(e);
  if (slot < 0 || slot > 1 || !X_dev || rows <= 0) return e->fail(MMAE_ERR_INVALID, "bad dataset arguments");
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "sync");
  if (e->ds_X[slot]) { cudaFree(e->ds_X[slot]); e->ds_X[slot] = nullptr; }
  if (e->ds_Y[slot]) { cudaFree(e->ds_Y[slot]); e->ds_Y[slot] = nullptr; }
  if (e->ds_view[slot]) { cudaFree(e->ds_view[slot]); e->ds_view[slot] = nullptr; }
  e->ds_view_rows[slot] = 0;
  e->clear_graphs();
  ce = cudaMalloc(&e->ds_X[slot], (size_t)rows * e->F * 4);
  if (ce == cudaSuccess) ce = cudaMemcpy(e->ds_X[slot], X_dev, (size_t)rows * e->F * 4, cudaMemcpyDeviceToDevice);
  if (ce == cudaSuccess && Y_dev && label_cols > 0) {
    ce = cudaMalloc(&e->ds_Y[slot], (size_t)rows * label_cols * 4);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->ds_Y[slot], Y_dev, (size_t)rows * label_cols * 4, cudaMemcpyDeviceToDevice);
  }
  if (ce != cudaSuccess) return e->cuda_fail(ce, "set_dataset_device");
  e->ds_rows[slot] = rows; e->ds_ycols[slot] = (Y_dev && label_cols > 0) ? label_cols : 0;
  return 0;
}

int mmae_set_dataset_view(mmae_engine* e, int slot, const int64_t* rows_host, int64_t count) {
  ENTER(e);
  if (slot < 0 || slot > 1 || !e->ds_X[slot]) return e->fail(MMAE_ERR_STATE, "dataset slot is empty");
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "sync");
  e->clear_graphs();
  if (e->ds_view[slot]) { cudaFree(e->ds_view[slot]); e->ds_view[slot] = nullptr; }
  e->ds_view_rows[slot] = 0;
  if (!rows_host || count <= 0) return 0;                      // back to "every row of the dataset"
  for (int64_t i = 0; i < count; ++i)
    if (rows_host[i] < 0 || rows_host[i] >= e->ds_rows[slot]) return e->fail(MMAE_ERR_INVALID, "view row outside the dataset");
  ce = cudaMalloc(&e->ds_view[slot], (size_t)count * 8);
  if (ce == cudaSuccess) ce = cudaMemcpy(e->ds_view[slot], rows_host, (size_t)count * 8, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "set_dataset_view");
  e->ds_view_rows[slot] = count;
  return 0;
}

int mmae_train_step_resident(mmae_engine* e, int slot, const int64_t* idx_host, int64_t batch, int gen_noise, float keep,
                             int classification) {
  ENTER(e);
  if (slot < 0 || slot > 1 || !e->ds_X[slot]) return e->fail(MMAE_ERR_STATE, "dataset slot is empty");
  if (classification && !e->ds_Y[slot]) return e->fail(MMAE_ERR_STATE, "dataset slot has no labels");
  if (classification && e->H == 0) return e->fail(MMAE_ERR_STATE, "engine was created without a classification head");
  if (classification) {
    const int want_cols = e->cfg.head_loss == MMAE_HEAD_SIGMOID_CE ? e->C : 1;
    if (e->ds_ycols[slot] != want_cols) return e->fail(MMAE_ERR_INVALID, "dataset label columns do not match the head (need C for a sigmoid head, 1 for softmax)");
  }
  int r = e->ensure_resident(batch); if (r) return r;
  // device-side sampling (Philox row indices, gather, Philox noise) + the optimizer step: every per-step value comes
  // from StepState in device memory, so the whole sequence replays as one graph
  const int64_t* view = e->ds_view[slot];
  const uint32_t n_rows = (uint32_t)(view ? e->ds_view_rows[slot] : e->ds_rows[slot]);
  if (idx_host) {
    for (int64_t i = 0; i < batch; ++i)
      if (idx_host[i] < 0 || idx_host[i] >= (int64_t)n_rows) return e->fail(MMAE_ERR_INVALID, "row index outside the dataset (view)");
    if (batch > e->idx_in_cap) {
      cudaError_t ce0 = cudaStreamSynchronize(e->stream);
      if (ce0 != cudaSuccess) return e->cuda_fail(ce0, "sync");
      e->clear_graphs();
      r = e->realloc_dev(e->d_idx_in, batch); if (r) return r;
      e->idx_in_cap = batch;
    }
  }
  auto body = [&]() -> int {
    cudaError_t ce;
    const int wpb = 8;
    bool noisy_ready = false;
    if (idx_host) {
      ce = cudaMemcpyAsync(e->d_idx_in, idx_host, (size_t)batch * 8, cudaMemcpyHostToDevice, e->stream);
      if (ce != cudaSuccess) return e->cuda_fail(ce, "H2D indices");
    }
    bool gather_target = false;
    if (gen_noise && noise_materialises(e, batch)) {
      // one kernel: Philox row indices (or the given ones), fold view, gather of the clean batch, descriptor, noisy batch.
      // Wide models skip the clean copy: their loss GEMM reads the target rows of the dataset through the index list.
      gather_target = !classification && e->final_gemm_gathers_target(batch, keep);
      int rr = launch_sample_noise(e, e->ds_X[slot], n_rows, idx_host ? e->d_idx_in : nullptr, batch, e->first_row,
                                   gather_target ? nullptr : e->gxb, view);
      if (rr) return rr;
      noisy_ready = true;
    } else {
      philox_indices_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, e->stream>>>(e->d_idx, batch, e->first_row, n_rows, &e->d_state->step,
                                                                                  e->cfg.seed, idx_host ? e->d_idx_in : nullptr, view);
      ++e->launches;
      gather_rows_kernel<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, e->stream>>>(e->ds_X[slot], e->d_idx, e->gxb, batch, e->F);
      ++e->launches;
      if (gen_noise) { int rr = launch_noise_gen(e, batch, e->first_row); if (rr) return rr; }
    }
    if (classification) {
      gather_rows_kernel<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, e->stream>>>(e->ds_Y[slot], e->d_idx, e->gyb, batch, e->ds_ycols[slot]);
      ++e->launches;
    }
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return e->cuda_fail(ce, "gather");
    if (gather_target) return train_core(e, e->ds_X[slot], nullptr, batch, 1, keep, true, e->d_idx);
    return classification ? cls_core(e, e->gxb, e->gyb, batch, gen_noise ? 1 : 0, keep, noisy_ready)
                          : train_core(e, e->gxb, nullptr, batch, gen_noise ? 1 : 0, keep, noisy_ready);
  };
  if (idx_host || e->sticky) return body();            // host-supplied indices: pageable copy, stay eager
  return e->run_graphed(graph_key(e, 2 + slot * 2 + (classification ? 1 : 0), e->ds_X[slot], e->ds_Y[slot], view, batch, gen_noise ? 1 : 0, keep),
                        classification ? 1 : 0, batch, body);
}

int mmae_modality_rmse(mmae_engine* e, const float* X_host, int64_t rows, double* rmse_host) {
  ENTER(e);
  if (!X_host || rows <= 0 || !rmse_host) return e->fail(MMAE_ERR_INVALID, "bad arguments");
  if (e->sticky) return e->fail(MMAE_ERR_CUDA, "engine is in a sticky CUDA error state: " + e->err);
  const int M = e->M, F = e->F;
  // rows travel in chunks of n; each chunk is one forward over its M masked copies (M * n rows)
  const int64_t n = std::min<int64_t>(rows, std::max<int64_t>(1024, ((int64_t)1 << 20) / M));
  int r = e->ensure_host(n); if (r) return r;
  r = e->ensure_acts(n * M); if (r) return r;
  const int nblk = e->num_sms * 2;
  double* d_part = nullptr; double* d_sse = nullptr;
  cudaError_t ce = cudaMalloc(&d_part, (size_t)M * nblk * 8);
  if (ce == cudaSuccess) ce = cudaMalloc(&d_sse, (size_t)M * 8);
  if (ce != cudaSuccess) { cudaFree(d_part); return e->cuda_fail(ce, "modality_rmse scratch"); }
  int rc = 0;
  for (int64_t r0 = 0; r0 < rows && rc == 0; r0 += n) {
    const int64_t nr = std::min(n, rows - r0);
    float* Xd = nullptr;
    int t = stage_host(e, X_host + r0 * F, nullptr, nr, 0, &Xd, nullptr);
    if (t < 0) { rc = t; break; }
    modality_mask_batch_kernel<<<e->grid_for((int64_t)M * nr * F, 256), 256, 0, e->stream>>>(Xd, e->noisy, nr, F, M, e->d_col_mod);
    ++e->launches;
    rc = e->begin_step(nr * M, false); if (rc) break;
    mmae_engine::FwdOpts o; o.X = e->noisy; o.target = nullptr; o.labels = nullptr; o.B = nr * M; o.noise = false; o.keep = 1.f;
    o.train_recon = false; o.decoder = true; o.headp = false; o.recon_out = nullptr; o.need_mu = false;
    rc = e->forward(o); if (rc) break;
    modality_sse_kernel<<<dim3(nblk, M), 256, 0, e->stream>>>(Xd, e->out, nr, F, e->d_starts, M, d_part);
    modality_sse_reduce_kernel<<<M, 32, 0, e->stream>>>(d_part, nblk, d_sse, e->d_starts, rows, r0 + nr >= rows ? 1 : 0, r0 == 0 ? 1 : 0);
    e->launches += 2;
    if (e->cfg.variational) { rc = e->advance_step(); if (rc) break; }
    rc = release_stage(e, t);
  }
  if (rc == 0) {
    ce = cudaMemcpyAsync(rmse_host, d_sse, (size_t)M * 8, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) rc = e->cuda_fail(ce, "modality_rmse");
  } else cudaStreamSynchronize(e->stream);
  cudaFree(d_part); cudaFree(d_sse);
  return rc;
}

int mmae_eval_resident(mmae_engine* e, int slot, int64_t batch, int gen_noise, float keep) {
  ENTER(e);
  if (slot < 0 || slot > 1 || !e->ds_X[slot]) return e->fail(MMAE_ERR_STATE, "dataset slot is empty");
  if (e->sticky) return e->fail(MMAE_ERR_CUDA, "engine is in a sticky CUDA error state: " + e->err);
  int r = e->ensure_resident(batch); if (r) return r;
  const int64_t* view = e->ds_view[slot];
  const uint32_t n_rows = (uint32_t)(view ? e->ds_view_rows[slot] : e->ds_rows[slot]);
  bool noisy_ready = false;
  if (gen_noise && noise_materialises(e, batch)) {
    r = launch_sample_noise(e, e->ds_X[slot], n_rows, nullptr, batch, e->first_row, e->gxb, view); if (r) return r;
    noisy_ready = true;
  } else {
    philox_indices_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, e->stream>>>(e->d_idx, batch, e->first_row, n_rows, &e->d_state->step,
                                                                                e->cfg.seed, nullptr, view);
    gather_rows_kernel<<<(unsigned)((batch + 7) / 8), 256, 0, e->stream>>>(e->ds_X[slot], e->d_idx, e->gxb, batch, e->F);
    e->launches += 2;
    if (gen_noise) { r = launch_noise_gen(e, batch, e->first_row); if (r) return r; }
  }
  r = e->begin_step(batch, gen_noise != 0); if (r) return r;
  mmae_engine::FwdOpts o; o.X = e->gxb; o.target = e->gxb; o.labels = nullptr; o.B = batch; o.noise = gen_noise != 0; o.keep = keep;
  o.train_recon = false; o.decoder = true; o.headp = false; o.recon_out = nullptr; o.need_mu = false; o.noisy_ready = noisy_ready;
  r = e->forward(o); if (r) return r;
  r = e->finalize_scalars(batch, true, false); if (r) return r;
  e->last_B = batch;
  return e->advance_step();
}

int mmae_read_scalars(mmae_engine* e, double* out, int count) {
  ENTER(e);
  if (!out || count <= 0 || count > MMAE_NUM_SCALARS) return e->fail(MMAE_ERR_INVALID, "bad scalar count");
  cudaError_t ce = cudaMemcpyAsync(out, e->d_scalars, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "read_scalars");
  return 0;
}

int mmae_comm_unique_id(void* id_out_128) {
  std::string err;
  if (!id_out_128) return MMAE_ERR_INVALID;
  if (!load_nccl(err)) { g_create_error = err; return MMAE_ERR_COMM; }
  int r = g_nccl.GetUniqueId(id_out_128);
  if (r != 0) { g_create_error = "ncclGetUniqueId failed"; return MMAE_ERR_COMM; }
  return 0;
}

int mmae_comm_init(mmae_engine* e, const void* id_128, int rank, int world_size) {
  ENTER(e);
  std::string err;
  if (!load_nccl(err)) return e->fail(MMAE_ERR_COMM, err);
  Id128 id; memcpy(&id, id_128, 128);
  // NCCL's CTAs cannot share an SM with a 227 KB GEMM CTA: give the collectives a fixed, small number of SMs and size
  // the persistent GEMM grids around them (mmae_engine::gemm).  MMAE_NCCL_CTAS overrides; an NCCL_MAX_CTAS already in
  // the environment wins.
  {
    const char* ev = getenv("MMAE_NCCL_CTAS");
    int ctas = ev ? atoi(ev) : 32;      // measured: 33.5 MB bucket 160 us at 32 CTAs, 185 us at 16, 282 us at 8 (profiles/r02_allreduce_8gpu.txt)
    if (ctas < 1) ctas = 1; if (ctas > 32) ctas = 32;
    char buf[16]; snprintf(buf, 16, "%d", ctas);
    setenv("NCCL_MAX_CTAS", buf, 0);
    const char* got = getenv("NCCL_MAX_CTAS");
    e->comm_reserve = got ? std::max(1, atoi(got)) : ctas;
  }
  int r = g_nccl.CommInitRank(&e->comm, world_size, id, rank);
  if (r != 0) return e->fail(MMAE_ERR_COMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  e->world = world_size; e->rank = rank;
  return 0;
}

int mmae_set_shard(mmae_engine* e, int64_t global_batch, int64_t first_row) {
  ENTER(e); e->global_batch = global_batch; e->first_row = first_row; return 0;
}

int64_t mmae_kernel_launches(const mmae_engine* e) { return e ? e->launches : 0; }
int64_t mmae_chain_launches(const mmae_engine* e) { return e ? e->chain_launches : 0; }
int64_t mmae_backward_chain_launches(const mmae_engine* e) { return e ? e->bchain_launches : 0; }
int64_t mmae_wgrad_group_launches(const mmae_engine* e) { return e ? e->wgroup_launches : 0; }
int64_t mmae_graph_replays(const mmae_engine* e) { return e ? e->graph_replays : 0; }
int64_t mmae_fused_noise_launches(const mmae_engine* e) { return e ? e->fused_noise_launches : 0; }

int mmae_set_profiling(mmae_engine* e, int on) {
  ENTER(e);
  if (e->profiling) e->prof_collect();
  e->profiling = on != 0;
  if (on) { e->prof_ms = 0.0; e->prof_flops = 0.0; e->prof_count = 0; }
  return 0;
}

int mmae_read_profile(mmae_engine* e, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches) {
  ENTER(e);
  e->prof_collect();
  if (gemm_ms) *gemm_ms = e->prof_ms;
  if (gemm_flops) *gemm_flops = e->prof_flops;
  if (gemm_launches) *gemm_launches = e->prof_count;
  return 0;
}

int mmae_read_scalars_async(mmae_engine* e, double* pinned_host, int count) {
  ENTER(e);
  if (!pinned_host || count <= 0 || count > MMAE_NUM_SCALARS) return e->fail(MMAE_ERR_INVALID, "bad scalar count");
  cudaError_t ce = cudaMemcpyAsync(pinned_host, e->d_scalars, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "read_scalars_async");
  return 0;
}

int mmae_get_buffer(mmae_engine* e, const char* name, float* host, int64_t count) {
  ENTER(e);
  const float* src = nullptr;
  std::string n = name ? name : "";
  if (n == "eps") src = e->eps; else if (n == "mu") src = e->mu; else if (n == "lv") src = e->lv;
  else if (n == "emb") src = e->cur_emb; else if (n == "out") src = e->out; else if (n == "logits") src = e->hlogits;
  if (!src) return e->fail(MMAE_ERR_NOTFOUND, "unknown or unallocated buffer " + n);
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce == cudaSuccess) ce = cudaMemcpy(host, src, (size_t)count * 4, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "get_buffer");
  return 0;
}

int mmae_set_eps(mmae_engine* e, const float* eps_host, int64_t count) {
  ENTER(e);
  if (!e->cfg.variational) return e->fail(MMAE_ERR_STATE, "not variational");
  if (!eps_host) { e->eps_injected = false; return 0; }
  int r = e->ensure_cap((count + e->E - 1) / e->E); if (r) return r;
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce == cudaSuccess) ce = cudaMemcpy(e->eps, eps_host, (size_t)count * 4, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "set_eps");
  e->eps_injected = true;
  return 0;
}

int mmae_debug_gemm(int precision, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                    const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias, int activation, float beta,
                    void* stream) {
  GemmArgs g; memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.splits = 1; g.k_per_split = 0;
  g.ep.mode = bias || activation ? EPI_BIAS_ACT : EPI_PLAIN; g.ep.bias = bias; g.ep.act = activation; g.ep.beta = beta; g.ep.keep = 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == MMAE_PREC_TF32) {
    if (!tc_gemm_eligible(transA != 0, transB != 0, g)) return MMAE_ERR_INVALID;
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool two_sm = tc2_eligible(transA != 0, transB != 0, g);
    TcPlan pl = two_sm ? tc2_plan(g, sms, 1) : tc_plan(g, sms, 1);
    cudaError_t e = two_sm ? launch_gemm_tc2(transA != 0, transB != 0, g, pl, nullptr, st)
                           : launch_gemm_tc(transA != 0, transB != 0, g, pl, nullptr, st);
    return e == cudaSuccess ? 0 : MMAE_ERR_CUDA;
  }
  cudaError_t e = launch_gemm_simt(transA != 0, transB != 0, g, st);
  return e == cudaSuccess ? 0 : MMAE_ERR_CUDA;
}

}  // extern "C"
