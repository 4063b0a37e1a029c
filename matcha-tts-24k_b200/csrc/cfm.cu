// libcfm_b200: C-ABI implementation (include/cfm_b200.h).  Host side of the CFM decode: weight packing, the
// pad-aware packed row tables, the per-NFE kernel schedule of the U-Net estimator (reference decoder.py:359-426),
// the fixed-grid ODE stages (reference flow_matching.py:60-63 -> torchdiffeq) and the CUDA graph around them.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/cfm_b200.h"
#include "attn.cuh"
#include "attn_tc.cuh"
#include "attn_persist.cuh"
#include "ff_fused.cuh"
#include "rowln.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

using namespace cfm;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::string g_create_error;

struct GemmW {  // one packed GEMM weight: [n_taps * n_stride, Kp] in the activation type, K-major
  void* w = nullptr;
  void* w3 = nullptr;  // fp32 mode: bf16 hi / lo split of w, [2 * n_taps * n_stride, Kp] (tensor-core fp32 mode, "fp32_tc")
  float* bias = nullptr;
  int N = 0, K = 0, Kp = 0, n_taps = 1, n_stride = 0;
};
struct NormW {
  float* gamma = nullptr;
  float* beta = nullptr;
  double* bias_gsum = nullptr;  // GroupNorm sites only: per-group (sum b, sum b^2) of the preceding conv bias
};
struct ResnetW {
  GemmW conv1, conv2, res;
  NormW gn1, gn2;
  float *mlp_w = nullptr, *mlp_b = nullptr;
};
struct BlockW {
  NormW ln1, ln3;
  GemmW qkv, out, ff1, ff2;
  float *ea = nullptr, *ib = nullptr;
  std::vector<float> h_b1, h_ea, h_ib;  // host copies: the fused feed-forward kernel takes them as a by-value kernel parameter
};
struct StageW {
  ResnetW res;
  std::vector<BlockW> blocks;
};
struct Model {
  float *t1_w = nullptr, *t1_b = nullptr, *t2_w = nullptr, *t2_b = nullptr;
  std::vector<StageW> stages;  // down0, down1, mid..., up0, up1
  GemmW down_s2, down_tail, up_even, up_odd, up_tail, final_conv, final_proj;
  NormW final_gn;
};

struct OdeStage {
  float t;
  float c_v;
  float c_k[3];
  int kin[3];  // index of stored k buffer or -1
  int kout;    // -1: not stored
  bool write_state;
};

struct LaneDef {  // a contiguous group of utterances that runs as its own branch of the CUDA graph
  int b0 = 0, nb = 0;      // utterance range
  int r2 = 0, m2 = 0;      // first half-resolution row, half-resolution rows (full resolution: 2x both)
  int w0[2] = {0, 0};      // first attention work item per resolution
  int nw[2] = {0, 0};      // attention work items per resolution
};

struct Plan {
  int B = 0, T = 0;
  int solver = 0;
  std::vector<float> t_span;   // with L, T, solver: the cache key (cfm_plan)
  unsigned long long last_use = 0;
  char* ws_base = nullptr;     // this plan's workspace block (from the handle's block pool)
  size_t ws_cap = 0, ws_off = 0;
  int n_trows = 0;             // rows of the time-embedding buffers: max(NFE, B) + 1 (per-utterance t of cfm_estimator_t)
  std::vector<LaneDef> lanes;
  std::vector<int> L;
  int M1 = 0, M2 = 0;
  std::vector<OdeStage> stages;
  std::vector<void*> allocs;
  size_t bytes = 0;
  // tables
  UttTable *utt1 = nullptr, *utt2 = nullptr;
  int *info1 = nullptr, *info2 = nullptr;
  int4 *work1 = nullptr, *work2 = nullptr;
  int n_work1 = 0, n_work2 = 0;
  // time embedding
  float *tvals = nullptr, *sinemb = nullptr, *temb_a = nullptr, *temb = nullptr, *tproj = nullptr;
  // state
  float *xstate = nullptr, *vout = nullptr, *kbuf[3] = {nullptr, nullptr, nullptr};
  double* stats = nullptr;
  size_t stats_bytes = 0;
  float *stage_mu = nullptr, *stage_z = nullptr, *stage_out = nullptr, *stage_spk = nullptr;  // cfm_solve_host device staging
  // activations per resolution (index 0 = full, 1 = half)
  void* xin = nullptr;
  int xin_ld = 0;
  float *hraw[2], *rres[2], *X[2];
  void *hact[2], *Xn[2], *qkv[2], *ao[2], *ffh[2], *sin_[2], *cat[2];
  void* split3 = nullptr;  // fp32_tc: bf16 [hi | lo] scratch of the current GEMM's activation operand
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  long long launches_per_solve = 0;
  int solves = 0;  // decodes run with this plan (the graph is captured lazily, see cfm_plan)
};

}  // namespace

struct cfm_handle {
  cfm_config cfg;
  bool bf = true;
  int es = 2;
  std::string err;
  std::vector<void*> wallocs;
  Model model;
  bool weights_loaded = false;
  Plan* plan = nullptr;           // current plan (one of `plans`)
  std::vector<Plan*> plans;       // LRU cache of plans keyed by (lengths, T, t_span, solver): a server alternates between shapes
  int plan_cache = 8;             // plans kept ("plan_cache" option)
  unsigned long long use_clock = 0;
  std::vector<std::pair<char*, size_t>> free_blocks;  // workspace blocks of evicted plans, reused by later plans
  cudaEvent_t busy_event = nullptr;  // recorded after every enqueue; the next call's stream waits on it (one workspace per plan,
  bool busy_valid = false;           // shared state in the handle: calls on different streams are ordered, never concurrent)
  // debug timeline (cfm_debug_timeline): one event before every launch of a direct-launch decode
  struct TlEntry { const char* tag; int M, N, K; double flops; cudaEvent_t ev; };
  bool tl_on = false;
  std::vector<TlEntry> tl;
  size_t tl_n = 0;
  const char* tag = "gemm";       // label of the next launch_gemm (set by the schedule code)
  EncodeTiledFn encode = nullptr;
  int sm_count = 148;
  int max_clusters[5] = {0, 148, 74, 0, 37};  // co-resident clusters of size 1, 2, 4 (queried at create)
  int cluster = 1;                              // 1-CTA kernel: CTAs sharing one weight tile via TMA multicast (no gain measured)
  int pdl = -1;                                 // programmatic dependent launch between the kernels of a decode: -1 auto (on for
                                                // plans of <= 2048 packed rows: -8 % on cfg1, where launch latency dominates; off
                                                // above: +0.5-2 % on cfg2 / cfg4), 0 off, 1 on (CFM_B200_PDL, cfm_set_option "pdl")
  int pdl_now = 0;                              // resolved per plan
  int direct_epi = (1 << EPI_STORE) | (1 << EPI_MASK);  // (-0.1 ms on cfg2; SNAKE is slower this way) bit m: direct (256-bit store, no smem) epilogue for bf16-output EpiMode m; "direct_epi"
  int pair_n256 = 0;                            // also use the pair kernel for short-K GEMMs with 256-column tiles (FF1); "pair_n256"
  int l2_persist_mb = 32;                       // persisting-L2 carve-out (MB) for the fp32 residual stream ("l2_persist_mb",
                                                // CFM_B200_L2_PERSIST_MB; cfg2 sustained: 0 -> 30.65, 32 -> 30.03, 48 -> 30.4, 64 -> 31.8, 79 -> 34.4 ms)
  void* win_ptr = nullptr;
  size_t win_bytes = 0, win_max = 0;
  int snake_warps = 12;                         // epilogue warps of the SnakeBeta (FF1) GEMM: 8 or 12; "snake_warps"
  int graph_after = 1;                          // decodes of a plan that use direct launches before its CUDA graph is built
                                                // (0: capture inside cfm_plan); "graph_after" option
  int small_tiles = 1024;                       // GEMMs with M <= this many rows use 64-column tiles (0: never); "small_tiles" option
  int bf16_mid = 1;                             // tensor-core mode: the conv output between a conv and its GroupNorm-apply and the
                                                // res_conv output are stored as bf16 (statistics still come from the fp32
                                                // accumulators), written by row-per-thread epilogues without smem staging; "bf16_mid"
  int fp32_tc = 0;                              // fp32 handles: GEMMs on the bf16 tensor pipe with bf16 x 3 split operands (hi*hi + lo*hi +
                                                // hi*lo, fp32 accumulate) instead of the fp32-FMA kernels; "fp32_tc".
  int attn_persist = 1;                         // persistent attention kernel (attn_persist.cuh): items streamed through one pipeline; "attn_persist".
  int rowln = 0;                                // Linear + residual + LayerNorm in one kernel (rowln.cuh) for out-proj -> norm3 and FF2 -> next
                                                // block's norm1: 0 off, 1 for plans above `small_tiles` rows, 2 always; "rowln".
  int rowln_ff2 = 1;                            // also fuse FF2 -> next norm1 (K = 4C); "rowln_ff2".
  unsigned long long* rowln_prof = nullptr;     // debug: device buffer for gemm_rowln_kernel's CTA-0 cycle counters
  int ff_fused = 0;                             // FF1 -> SnakeBeta -> FF2 in one kernel (ff_fused.cuh) for plans above `small_tiles` rows; "ff_fused".
                                                // Off by default: correct, removes the [rows, 4C] round trip through HBM, but measured
                                                // 108 / 63 us per full / half resolution block against 96 / 58 us for the two GEMMs
  int bn_full = 0, bn_half = 0;                 // experiment: tile width for N == C GEMMs with M > / <= 16384 rows (0 = pick_bn); "bn_full", "bn_half"
  int pair_min_k = 1024;                        // the pair kernel is used when taps * K >= this; "pair_min_k"
  int pair_mode = 1;                            // use the CTA-pair (cta_group::2) GEMM kernel
  int tma_epi = 1 << EPI_RESID;                 // bit m: TMA-store epilogue for EpiMode m.  Measured on cfg2: the in-place
                                                // residual add through cp.reduce.async.bulk saves 1.2 ms per decode; STORE / SNAKE /
                                                // MASK are 0-2 ms slower than the transposing epilogue, so they stay off
  long long launch_counter = 0;
  const float* spks = nullptr;  // device (B, S) speaker vectors for the next pack (cfm_set_speakers); S = in_channels - 2 F
  unsigned long long* ff_prof = nullptr;    // debug: device buffer for ff_fused_kernel's CTA-0 cycle counters
  unsigned long long* attn_prof = nullptr;  // debug: device buffer for attn_tc_kernel's CTA-0 cycle counters
  long long stop_after = -1;  // debug: skip every launch after this many (cfm_debug_stop_after)
  bool stopped() const { return stop_after >= 0 && launch_counter >= stop_after; }
  cudaStream_t own_stream = nullptr;
  // Utterances never interact inside the solve, so a batch is cut into `lanes_req` contiguous groups ("lanes") whose kernel
  // chains are independent branches of the graph: one lane's kernels fill the partial last wave / launch gaps of another's.
  int lanes_req = 1;  // measured on cfg2-cfg4: 2 lanes +2 %, 3 lanes +12 % decode time: off by default
  int lane_min_rows = 4096;  // do not cut below this many full-resolution rows per lane
  std::vector<cudaStream_t> lane_streams;
  std::vector<cudaEvent_t> lane_events;
  cudaEvent_t fork_event = nullptr;
  cudaStream_t side_stream = nullptr;            // side branch of the schedule (res_conv), "res_side"
  cudaEvent_t side_fork = nullptr, side_join = nullptr;
  int res_side = -1;                             // -1: automatic (plans of <= 2048 packed rows, where every launch is latency-bound: B = 1
                                                 // 6.65 -> 6.55 ms; neutral on cfg2 / cfg3, where the GPU is saturated), 0 off, 1 on
  int C() const { return cfg.channels; }
  int inner() const { return cfg.n_heads * cfg.head_dim; }
};

namespace {

int fail(cfm_handle* h, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                                  \
  do {                                                                                                            \
    cudaError_t e_ = (call);                                                                                      \
    if (e_ != cudaSuccess) return fail(h, CFM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKR(expr)          \
  do {                     \
    int r_ = (expr);       \
    if (r_ != 0) return r_; \
  } while (0)

inline int roundup(int x, int m) { return (x + m - 1) / m * m; }
inline void* act_off(void* p, long long elems, int es) { return static_cast<char*>(p) + elems * es; }
inline const void* act_off(const void* p, long long elems, int es) { return static_cast<const char*>(p) + elems * es; }

int dev_alloc(cfm_handle* h, std::vector<void*>& arena, void** out, size_t bytes, size_t* total = nullptr) {
  bytes = (bytes + 1023) / 1024 * 1024;
  if (bytes == 0) bytes = 1024;
  CK(cudaMalloc(out, bytes));
  CK(cudaMemset(*out, 0, bytes));
  arena.push_back(*out);
  if (total) *total += bytes;
  return 0;
}
template <typename P>
int dev_alloc_t(cfm_handle* h, std::vector<void*>& arena, P** out, size_t count, size_t* total = nullptr) {
  return dev_alloc(h, arena, reinterpret_cast<void**>(out), count * sizeof(P), total);
}
void free_arena(std::vector<void*>& arena) {
  for (void* p : arena) cudaFree(p);
  arena.clear();
}

// Bump allocation from the plan's workspace block; falls back to cudaMalloc if the size estimate was short.
int plan_alloc(cfm_handle* h, Plan* pl, void** out, size_t bytes) {
  bytes = (bytes + 1023) / 1024 * 1024;
  if (bytes == 0) bytes = 1024;
  if (pl->ws_off + bytes <= pl->ws_cap) {
    *out = pl->ws_base + pl->ws_off;
    pl->ws_off += bytes;
    pl->bytes += bytes;
    return 0;
  }
  return dev_alloc(h, pl->allocs, out, bytes, &pl->bytes);
}
template <typename P>
int plan_alloc_t(cfm_handle* h, Plan* pl, P** out, size_t count) {
  return plan_alloc(h, pl, reinterpret_cast<void**>(out), count * sizeof(P));
}
// Workspace block for a new plan: a block of an evicted plan if one fits (a server sees a new length with every request and
// must not pay cudaMalloc / cudaFree, which synchronise the device), else a fresh allocation.  Cleared on `s`.
int acquire_workspace(cfm_handle* h, Plan* pl, size_t need, cudaStream_t s) {
  int best = -1;
  for (int i = 0; i < (int)h->free_blocks.size(); ++i)
    if (h->free_blocks[i].second >= need && (best < 0 || h->free_blocks[i].second < h->free_blocks[best].second)) best = i;
  if (best >= 0 && h->free_blocks[best].second <= 4 * need + (64u << 20)) {
    pl->ws_base = h->free_blocks[best].first, pl->ws_cap = h->free_blocks[best].second;
    h->free_blocks.erase(h->free_blocks.begin() + best);
  } else {
    const size_t cap = need + need / 8;
    CK(cudaMalloc(reinterpret_cast<void**>(&pl->ws_base), cap));
    pl->ws_cap = cap;
  }
  // Guard rows and padding columns are never written by the kernels and must read as zero (stale bytes of a previous
  // plan, reinterpreted as bf16, can be Inf/NaN and 0 * NaN poisons an MMA): clear what this plan will use (~0.1 ms/GB).
  CK(cudaMemsetAsync(pl->ws_base, 0, std::min(need, pl->ws_cap), s));
  pl->ws_off = 0;
  return 0;
}

void free_plan(cfm_handle* h, Plan* pl, bool keep_block) {
  if (!pl) return;
  if (h->plan == pl) h->plan = nullptr, h->win_bytes = 0;  // the L2 access-policy window points into the plan's workspace
  if (pl->exec) cudaGraphExecDestroy(pl->exec);
  if (pl->graph) cudaGraphDestroy(pl->graph);
  free_arena(pl->allocs);
  if (pl->ws_base) {
    if (keep_block && h->free_blocks.size() < 4) h->free_blocks.push_back({pl->ws_base, pl->ws_cap});
    else cudaFree(pl->ws_base);
  }
  delete pl;
}
// Drops every cached plan (weights or kernel-selection options changed, handle destroyed).  Synchronises the device first:
// the plans' buffers may still be in use.
void free_all_plans(cfm_handle* h) {
  if (!h->plans.empty() || !h->free_blocks.empty()) cudaDeviceSynchronize();
  for (Plan* pl : h->plans) free_plan(h, pl, false);
  h->plans.clear();
  h->plan = nullptr;
  for (auto& b : h->free_blocks) cudaFree(b.first);
  h->free_blocks.clear();
}

// Orders this call behind whatever the handle enqueued last, on any stream (the workspace is shared state).
int enter_stream(cfm_handle* h, cudaStream_t s) {
  if (h->busy_valid) CK(cudaStreamWaitEvent(s, h->busy_event, 0));
  return 0;
}
int leave_stream(cfm_handle* h, cudaStream_t s) {
  CK(cudaEventRecord(h->busy_event, s));
  h->busy_valid = true;
  return 0;
}

// Timeline mark: an event before the launch that follows (cfm_debug_timeline).
int tl_mark(cfm_handle* h, cudaStream_t s, const char* tag, int M, int N, int K, double flops) {
  if (!h->tl_on) return 0;
  if (h->tl_n == h->tl.size()) {
    cfm_handle::TlEntry e{};
    CK(cudaEventCreate(&e.ev));
    h->tl.push_back(e);
  }
  cfm_handle::TlEntry& e = h->tl[h->tl_n++];
  e.tag = tag, e.M = M, e.N = N, e.K = K, e.flops = flops;
  CK(cudaEventRecord(e.ev, s));
  return 0;
}

// ------------------------------------------------------------------------------------------------ weights
struct DescMap {
  std::map<std::string, const cfm_weight_desc*> by_name;
  std::map<std::string, bool> used;
};

int want(cfm_handle* h, DescMap& dm, const std::string& name, std::initializer_list<long long> shape, const float** out) {
  auto it = dm.by_name.find(name);
  if (it == dm.by_name.end()) return fail(h, CFM_ERR_WEIGHTS, "missing parameter '%s'", name.c_str());
  const cfm_weight_desc* d = it->second;
  bool ok = d->ndim == (int)shape.size();
  int i = 0;
  for (long long s : shape) ok = ok && d->shape[i++] == s;
  if (!ok) return fail(h, CFM_ERR_WEIGHTS, "parameter '%s' has the wrong shape", name.c_str());
  if (d->data == nullptr) return fail(h, CFM_ERR_WEIGHTS, "parameter '%s' has a null data pointer", name.c_str());
  dm.used[name] = true;
  *out = d->data;
  return 0;
}

int copy_f32(cfm_handle* h, const float* src, size_t n, float** out) {
  CKR(dev_alloc_t(h, h->wallocs, out, n));
  CK(cudaMemcpy(*out, src, n * sizeof(float), cudaMemcpyDeviceToDevice));
  return 0;
}

// Pack src (fp32) into rows [row0, row0 + N) of each tap block of an already allocated GemmW.
int pack_into(cfm_handle* h, GemmW& g, const float* src, long long s_n, long long s_k, long long s_t, int4 tap_k, int N,
              int row0) {
  long long total = (long long)N * g.Kp * g.n_taps;
  int threads = 256;
  int blocks = (int)((total + threads - 1) / threads);
  if (h->bf)
    pack_weight_kernel<bf16><<<blocks, threads>>>(src, s_n, s_k, s_t, g.n_taps, tap_k, N, g.K,
                                                  static_cast<bf16*>(g.w) + (long long)row0 * g.Kp, g.Kp, g.n_stride);
  else
    pack_weight_kernel<float><<<blocks, threads>>>(src, s_n, s_k, s_t, g.n_taps, tap_k, N, g.K,
                                                   static_cast<float*>(g.w) + (long long)row0 * g.Kp, g.Kp, g.n_stride);
  CK(cudaGetLastError());
  if (!h->bf && g.w3) {  // keep the bf16 hi / lo split of the whole weight current (idempotent; weights are packed once per checkpoint)
    const long long n = (long long)g.n_taps * g.n_stride * g.Kp;
    split3_weight_kernel<<<(int)((n + 255) / 256), 256>>>(static_cast<const float*>(g.w), n, static_cast<bf16*>(g.w3));
    CK(cudaGetLastError());
  }
  return 0;
}

int alloc_gemm(cfm_handle* h, GemmW& g, int N, int K, int n_taps) {
  g.N = N, g.K = K, g.Kp = roundup(K, 64), g.n_taps = n_taps, g.n_stride = roundup(N, 64);
  CKR(dev_alloc(h, h->wallocs, &g.w, (size_t)g.n_taps * g.n_stride * g.Kp * h->es));
  if (!h->bf) CKR(dev_alloc(h, h->wallocs, &g.w3, (size_t)2 * g.n_taps * g.n_stride * g.Kp * sizeof(bf16)));
  return 0;
}

int load_linear(cfm_handle* h, DescMap& dm, const std::string& name, int N, int K, bool has_bias, GemmW& g) {
  const float* w;
  CKR(want(h, dm, name + ".weight", {N, K}, &w));
  CKR(alloc_gemm(h, g, N, K, 1));
  CKR(pack_into(h, g, w, K, 1, 0, make_int4(0, 0, 0, 0), N, 0));
  if (has_bias) {
    const float* b;
    CKR(want(h, dm, name + ".bias", {N}, &b));
    CKR(copy_f32(h, b, N, &g.bias));
  }
  return 0;
}

int load_conv(cfm_handle* h, DescMap& dm, const std::string& name, int N, int K, int kw, GemmW& g) {
  const float *w, *b;
  CKR(want(h, dm, name + ".weight", {N, K, kw}, &w));
  CKR(want(h, dm, name + ".bias", {N}, &b));
  CKR(alloc_gemm(h, g, N, K, kw));
  CKR(pack_into(h, g, w, (long long)K * kw, kw, 1, make_int4(0, 1, 2, 3), N, 0));
  CKR(copy_f32(h, b, N, &g.bias));
  return 0;
}

int load_norm(cfm_handle* h, DescMap& dm, const std::string& name, int C, NormW& n, const float* conv_bias) {
  const float *g, *b;
  CKR(want(h, dm, name + ".weight", {C}, &g));
  CKR(want(h, dm, name + ".bias", {C}, &b));
  CKR(copy_f32(h, g, C, &n.gamma));
  CKR(copy_f32(h, b, C, &n.beta));
  if (conv_bias) {
    CKR(dev_alloc_t(h, h->wallocs, &n.bias_gsum, 16));
    bias_group_sums_kernel<<<1, 32>>>(conv_bias, C, C / 8, n.bias_gsum);
    CK(cudaGetLastError());
  }
  return 0;
}

int load_resnet(cfm_handle* h, DescMap& dm, const std::string& p, int cin, ResnetW& r) {
  const int C = h->C(), T4 = 4 * C;
  const float *mw, *mb;
  CKR(want(h, dm, p + ".mlp.1.weight", {C, T4}, &mw));
  CKR(want(h, dm, p + ".mlp.1.bias", {C}, &mb));
  CKR(copy_f32(h, mw, (size_t)C * T4, &r.mlp_w));
  CKR(copy_f32(h, mb, C, &r.mlp_b));
  CKR(load_conv(h, dm, p + ".block1.block.0", C, cin, 3, r.conv1));
  CKR(load_norm(h, dm, p + ".block1.block.1", C, r.gn1, r.conv1.bias));
  CKR(load_conv(h, dm, p + ".block2.block.0", C, C, 3, r.conv2));
  CKR(load_norm(h, dm, p + ".block2.block.1", C, r.gn2, r.conv2.bias));
  CKR(load_conv(h, dm, p + ".res_conv", C, cin, 1, r.res));
  return 0;
}

int load_block(cfm_handle* h, DescMap& dm, const std::string& p, BlockW& b) {
  const int C = h->C(), I = h->inner();
  CKR(load_norm(h, dm, p + ".norm1", C, b.ln1, nullptr));
  CKR(load_norm(h, dm, p + ".norm3", C, b.ln3, nullptr));
  const float *wq, *wk, *wv;
  CKR(want(h, dm, p + ".attn1.to_q.weight", {I, C}, &wq));
  CKR(want(h, dm, p + ".attn1.to_k.weight", {I, C}, &wk));
  CKR(want(h, dm, p + ".attn1.to_v.weight", {I, C}, &wv));
  CKR(alloc_gemm(h, b.qkv, 3 * I, C, 1));
  CKR(pack_into(h, b.qkv, wq, C, 1, 0, make_int4(0, 0, 0, 0), I, 0));
  CKR(pack_into(h, b.qkv, wk, C, 1, 0, make_int4(0, 0, 0, 0), I, I));
  CKR(pack_into(h, b.qkv, wv, C, 1, 0, make_int4(0, 0, 0, 0), I, 2 * I));
  CKR(load_linear(h, dm, p + ".attn1.to_out.0", C, I, true, b.out));
  const std::string ff = p + ".ff._orig_mod.net";
  CKR(load_linear(h, dm, ff + ".0.proj", 4 * C, C, true, b.ff1));
  CKR(load_linear(h, dm, ff + ".2", C, 4 * C, true, b.ff2));
  const float *al, *be;
  CKR(want(h, dm, ff + ".0.alpha", {4 * C}, &al));
  CKR(want(h, dm, ff + ".0.beta", {4 * C}, &be));
  CKR(dev_alloc_t(h, h->wallocs, &b.ea, 4 * C));
  CKR(dev_alloc_t(h, h->wallocs, &b.ib, 4 * C));
  snake_consts_kernel<<<(4 * C + 255) / 256, 256>>>(al, be, 4 * C, b.ea, b.ib);
  CK(cudaGetLastError());
  b.h_b1.resize(4 * C), b.h_ea.resize(4 * C), b.h_ib.resize(4 * C);
  CK(cudaMemcpy(b.h_b1.data(), b.ff1.bias, sizeof(float) * 4 * C, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.h_ea.data(), b.ea, sizeof(float) * 4 * C, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.h_ib.data(), b.ib, sizeof(float) * 4 * C, cudaMemcpyDeviceToHost));
  return 0;
}

int load_stage(cfm_handle* h, DescMap& dm, const std::string& p, int cin, StageW& s) {
  CKR(load_resnet(h, dm, p + ".0", cin, s.res));
  s.blocks.resize(h->cfg.n_blocks);
  for (int j = 0; j < h->cfg.n_blocks; ++j) CKR(load_block(h, dm, p + ".1." + std::to_string(j), s.blocks[j]));
  return 0;
}

// ------------------------------------------------------------------------------------------------ launches
int make_tmap(cfm_handle* h, CUtensorMap* tm, const void* base, long long inner_elems, long long outer_rows,
              long long row_stride_bytes, int box_inner, int box_outer) {
  cuuint64_t dims[2] = {(cuuint64_t)inner_elems, (cuuint64_t)outer_rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, CFM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld stride=%lld box=%dx%d", (int)r,
                inner_elems, outer_rows, row_stride_bytes, box_inner, box_outer);
  return 0;
}

// Output tensor map of the TMA-store epilogue: 32 x 32 element boxes, bf16 (SWIZZLE_64B rows of 64 B) or fp32 (SWIZZLE_128B).
int make_out_tmap(cfm_handle* h, CUtensorMap* tm, const void* base, bool f32, long long cols, long long rows, long long ld_elems) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, CFM_ERR_CUDA, "cuTensorMapEncodeTiled (output map) failed (%d) cols=%lld rows=%lld ld=%lld", (int)r, cols, rows, ld_elems);
  return 0;
}

// Small batches (the server's B = 1 requests) leave most SMs idle and every kernel is latency-bound: narrow tiles spread a
// GEMM over 3-4x more CTAs and shorten each CTA's epilogue (cfg1: measured in DESIGN.md section 7).
int pick_bn_small(int N) {
  if (N % 64 == 0) return 64;
  return N <= 64 ? 64 : 128;
}

int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0) return 128;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 192;
}

// Kernel launch with optional cluster dimension and the programmatic-dependent-launch attribute (PDL): the kernel may be
// scheduled while its predecessor in the stream is still draining; it calls griddepcontrol.wait before touching memory.
template <typename... KArgs, typename... Args>
int launch_ex(cfm_handle* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster,
              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[3];
  int n = 0;
  if (h->win_bytes > 0) {  // keep the fp32 residual stream of the current resolution resident in the persisting part of L2
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = h->win_ptr;
    attr[n].val.accessPolicyWindow.num_bytes = h->win_bytes;
    attr[n].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (h->pdl_now) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr, cfg.numAttrs = n;
  CK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
  return 0;
}

// Launch information of the tensor-core GEMM kernels (instantiated in gemm_inst.cu, one object each).
KernelInfo tc_info(int bn, int epi_warps) {
  if (bn == 256 && epi_warps == 12) return kinfo_tc_256_12();
  switch (bn) {
    case 64: return kinfo_tc_64_8();
    case 128: return kinfo_tc_128_8();
    case 160: return kinfo_tc_160_8();
    case 192: return kinfo_tc_192_8();
    default: return kinfo_tc_256_8();
  }
}
KernelInfo tc2_info(int bn) {
  switch (bn) {
    case 128: return kinfo_tc2_128();
    case 160: return kinfo_tc2_160();
    case 192: return kinfo_tc2_192();
    default: return kinfo_tc2_256();
  }
}

// Same as launch_ex for a kernel known by address (the GEMM kernels all take (tmA0, tmA1, tmW, tmOut, params)).
int launch_gemm_kernel(cfm_handle* h, const KernelInfo& k, int grid, cudaStream_t s, int cluster, const CUtensorMap& a0,
                       const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o, const GemmParams& p) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(k.threads), cfg.dynamicSmemBytes = k.smem, cfg.stream = s;
  cudaLaunchAttribute attr[3];
  int n = 0;
  if (h->win_bytes > 0) {
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = h->win_ptr;
    attr[n].val.accessPolicyWindow.num_bytes = h->win_bytes;
    attr[n].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (h->pdl_now) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr, cfg.numAttrs = n;
  void* args[] = {const_cast<CUtensorMap*>(&a0), const_cast<CUtensorMap*>(&a1), const_cast<CUtensorMap*>(&w),
                  const_cast<CUtensorMap*>(&o), const_cast<GemmParams*>(&p)};
  CK(cudaLaunchKernelExC(&cfg, k.fn, args));
  return 0;
}

int launch_tc_bn(cfm_handle* h, int BN, int epi_warps, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w,
                 const CUtensorMap& o, const GemmParams& p, cudaStream_t s) {
  const int CL = p.cluster;
  const int m_super = ((p.M + 127) / 128 + CL - 1) / CL;
  const int super_tiles = m_super * ((p.N + BN - 1) / BN);
  const int clusters = std::min(super_tiles, h->max_clusters[CL]);
  return launch_gemm_kernel(h, tc_info(BN, epi_warps), clusters * CL, s, CL, a0, a1, w, o, p);
}

int launch_tc2_bn(cfm_handle* h, int BN, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o,
                  const GemmParams& p, cudaStream_t s) {
  const int m_pairs = (p.M + 255) / 256;
  const int pair_tiles = m_pairs * ((p.N + BN - 1) / BN);
  const int pairs = std::min(pair_tiles, h->max_clusters[2]);
  return launch_gemm_kernel(h, tc2_info(BN), pairs * 2, s, 2, a0, a1, w, o, p);
}

int set_gemm_attrs(cfm_handle* h) {
  const int bns[] = {64, 128, 160, 192, 256};
  for (int bn : bns) {
    const KernelInfo k = tc_info(bn, 8);
    CK(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem));
  }
  {
    const KernelInfo k = tc_info(256, 12);
    CK(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem));
  }
  const int bns2[] = {128, 160, 192, 256};
  for (int bn : bns2) {
    const KernelInfo k = tc2_info(bn);
    CK(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem));
  }
  {
    const KernelInfo k = kinfo_ff_fused();
    CK(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem));
  }
  {
    const KernelInfo k = kinfo_rowln();
    CK(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem));
  }
  const KernelInfo k = tc_info(192, 8);  // co-resident cluster capacity (1 CTA per SM): bounds the persistent grid
  for (int CL = 1; CL <= 4; CL *= 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(h->sm_count / CL * CL), cfg.blockDim = dim3(k.threads), cfg.dynamicSmemBytes = k.smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    int n = 0;
    CK(cudaOccupancyMaxActiveClusters(&n, k.fn, &cfg));
    h->max_clusters[CL] = std::max(1, n);
  }
  return 0;
}

// fp32 handle with "fp32_tc": can this GEMM run on the bf16 tensor pipe with split operands (launch_gemm)?
bool x3_ok(const cfm_handle* h, const GemmParams& p) {
  return !h->bf && h->fp32_tc && !(h->cfg.flags & CFM_FLAG_SIMT_GEMM) && h->plan && h->plan->split3 && p.W3 && p.n_taps * 3 <= MAX_TAPS &&
         p.K % 64 == 0 && p.lda[0] % 4 == 0 && p.lda[0] <= 4 * h->C() && !p.A[1];
}

// A sources: p.A / p.lda / p.a_rows already describe element-addressed matrices.
int launch_gemm(cfm_handle* h, GemmParams& p, bool allow_tc, cudaStream_t s) {
  if (h->stopped()) return 0;
  h->launch_counter++;
  CKR(tl_mark(h, s, h->tag, p.M, p.N, p.K * p.n_taps, 2.0 * p.M * p.N * p.K * p.n_taps));
  bool tc = h->bf && allow_tc && !(h->cfg.flags & CFM_FLAG_SIMT_GEMM);
  const bool x3 = allow_tc && x3_ok(h, p);
  if (x3) {
    // fp32 mode on the tensor pipe: the activation operand is split into bf16 [hi | lo] (one pass), every tap becomes three
    // (activation range, weight range) pairs hi*hi + lo*hi + hi*lo accumulated in fp32 by the bf16 kernels; fp32 results.
    const long long lda = p.lda[0], rows = p.a_rows[0];
    const long long quads = rows * (lda / 4);
    h->launch_counter++;
    CKR(launch_ex(h, split3_rows_kernel, dim3((unsigned)((quads + 255) / 256)), dim3(256), 0, s, 1, static_cast<const float*>(p.A[0]), lda, rows,
                  static_cast<bf16*>(h->plan->split3)));
    const int nt = p.n_taps;
    GemmTap t3[MAX_TAPS];
    for (int t = 0; t < nt; ++t) {
      const GemmTap o = p.taps[t];
      t3[3 * t] = o;
      t3[3 * t + 1] = o, t3[3 * t + 1].a_col = o.a_col + (int)lda;
      t3[3 * t + 2] = o, t3[3 * t + 2].w_row = o.w_row + p.w_rows;
    }
    for (int t = 0; t < 3 * nt; ++t) p.taps[t] = t3[t];
    p.n_taps = 3 * nt;
    p.A[0] = h->plan->split3, p.lda[0] = 2 * lda;
    p.W = p.W3, p.w_rows = 2 * p.w_rows;
    p.act_f32 = 1;
    tc = true;
  }
  if (!tc) {
    dim3 grid((p.N + 63) / 64, (p.M + 63) / 64);
    if (h->bf)
      gemm_simt_kernel<bf16><<<grid, 256, 0, s>>>(p);
    else
      gemm_simt_kernel<float><<<grid, 256, 0, s>>>(p);
    CK(cudaGetLastError());
    return 0;
  }
  if (p.K % 64 != 0) return fail(h, CFM_ERR_INVALID, "tensor-core GEMM needs K %% 64 == 0 (K=%d)", p.K);
  int bn = (h->small_tiles && p.M <= h->small_tiles) ? pick_bn_small(p.N) : pick_bn(p.N);
  if (p.N == h->C() && p.M > 1024) {
    const int o = p.M > 16384 ? h->bn_full : h->bn_half;
    if (o == 64 || o == 128 || o == 192) bn = o;
  }
  CUtensorMap tmA[2], tmW;
  for (int i = 0; i < 2; ++i) {
    const int src = p.A[i] ? i : 0;
    CKR(make_tmap(h, &tmA[i], p.A[src], p.lda[src], p.a_rows[src], p.lda[src] * 2, 64, 128));
  }
  // TMA-store epilogue (gemm.cuh epilogue_tile_tma): bf16-output modes, and the in-place fp32 residual add when no
  // activation copy is wanted (L2 performs x += acc + bias through cp.reduce.async.bulk).
  const bool bf_mode = p.mode == EPI_STORE || p.mode == EPI_SNAKE || p.mode == EPI_MASK;
  const bool red_mode = p.mode == EPI_RESID && p.out_act == nullptr && p.resid != nullptr && p.resid == p.out_f32 && p.ld_resid == p.ld_f32;
  p.tma_epi = (((h->tma_epi >> p.mode) & 1) && ((bf_mode && !p.act_f32) || red_mode) && p.N % 8 == 0 && bn % 32 == 0) ? 1 : 0;
  p.direct_epi = (((h->direct_epi >> p.mode) & 1) && bf_mode && !p.act_f32 && !p.tma_epi && p.N % 32 == 0 && p.ld_act % 16 == 0 &&
                  (reinterpret_cast<uintptr_t>(p.out_act) & 31) == 0) ? 1 : 0;
  CUtensorMap tmO = tmA[0];
  if (p.tma_epi) {
    if (red_mode) CKR(make_out_tmap(h, &tmO, p.out_f32, true, p.N, p.M, p.ld_f32));
    else CKR(make_out_tmap(h, &tmO, p.out_act, false, p.N, p.M, p.ld_act));
  }
  // CTA-pair kernel where it measures faster: long reductions (k=3 convs, FF2); short-K GEMMs are epilogue-bound there.
  const bool pair_ok = h->pair_mode == 2 || (h->pair_mode == 1 && (p.n_taps * p.K >= h->pair_min_k || (h->pair_n256 && bn == 256)));
  p.pair = pair_ok && bn >= 128 ? 1 : 0;
  p.cluster = p.pair ? 1 : h->cluster;
  CKR(make_tmap(h, &tmW, p.W, p.ldw, p.w_rows, p.ldw * 2, 64, p.pair ? bn / 2 : bn / p.cluster));
  if (p.pair) return launch_tc2_bn(h, bn, tmA[0], tmA[1], tmW, tmO, p, s);
  const bool snake12 = bn == 256 && p.mode == EPI_SNAKE && h->snake_warps == 12 && p.cluster == 1 && !p.tma_epi && !p.direct_epi;
  return launch_tc_bn(h, bn, snake12 ? 12 : 8, tmA[0], tmA[1], tmW, tmO, p, s);
}

// Common part of a GEMM call: one A matrix, `w.n_taps` taps with the given row shifts / A column offsets.
GemmParams gemm_base(int M, const void* A, long long lda, int a_rows, const GemmW& w, const int* shifts, const int* acols) {
  GemmParams p;
  memset(&p, 0, sizeof p);
  p.M = M, p.N = w.N, p.K = w.Kp, p.n_taps = w.n_taps;
  for (int t = 0; t < w.n_taps; ++t) {
    p.taps[t].a_src = 0;
    p.taps[t].row_shift = shifts ? shifts[t] : 0;
    p.taps[t].a_col = acols ? acols[t] : 0;
    p.taps[t].w_row = t * w.n_stride;
  }
  p.A[0] = A, p.lda[0] = lda, p.a_rows[0] = a_rows;
  p.W = w.w, p.ldw = w.Kp, p.w_rows = w.n_taps * w.n_stride;
  p.W3 = w.w3;
  p.bias = w.bias;
  p.row_mul = 1, p.row_add = 0;
  return p;
}

struct Res {  // per-resolution view of one lane of the plan (all pointers already offset to the lane's first row)
  int M;
  int b0, nb;      // utterance range of the lane
  UttTable* utt;   // whole-batch table (indexed by the global utterance id stored in row_info / work items)
  int* info;
  int4* work;
  int n_work;
  float *hraw, *rres, *X;
  void *hact, *Xn, *qkv, *ao, *ffh, *sin_, *cat;
  void *qkv_all, *ao_all;  // un-offset buffers + total rows: the attention kernel addresses rows through UttTable.start
  int M_all;
};
Res res_of(cfm_handle* h, Plan* pl, int r, const LaneDef& ln) {
  const long long C = h->C(), I = h->inner();
  const int es = h->es;
  const long long row0 = r == 0 ? 2LL * ln.r2 : ln.r2;
  Res v;
  v.M = r == 0 ? 2 * ln.m2 : ln.m2;
  v.b0 = ln.b0, v.nb = ln.nb;
  v.utt = r == 0 ? pl->utt1 : pl->utt2;
  v.info = (r == 0 ? pl->info1 : pl->info2) + row0;
  v.work = (r == 0 ? pl->work1 : pl->work2) + ln.w0[r];
  v.n_work = ln.nw[r];
  v.hraw = pl->hraw[r] + row0 * C, v.rres = pl->rres[r] + row0 * C, v.X = pl->X[r] + row0 * C;
  v.hact = act_off(pl->hact[r], row0 * C, es), v.Xn = act_off(pl->Xn[r], row0 * C, es);
  v.qkv = act_off(pl->qkv[r], row0 * 3 * I, es), v.ao = act_off(pl->ao[r], row0 * I, es);
  v.ffh = act_off(pl->ffh[r], row0 * 4 * C, es);
  v.qkv_all = pl->qkv[r], v.ao_all = pl->ao[r], v.M_all = r == 0 ? pl->M1 : pl->M2;
  v.sin_ = act_off(pl->sin_[r], row0 * C, es), v.cat = act_off(pl->cat[r], row0 * 2 * C, es);
  return v;
}

// bf16 conv / res_conv outputs (cfm_handle::bf16_mid): tensor-core schedule with fused statistics only
bool mid16(const cfm_handle* h) {  // (the row-per-thread statistics assume at most one group boundary per 32 columns: C >= 256)
  return h->bf && h->bf16_mid && h->cfg.channels % 32 == 0 && h->cfg.channels / 8 >= 32 &&
         !(h->cfg.flags & (CFM_FLAG_SIMT_GEMM | CFM_FLAG_UNFUSED_STATS));
}

template <typename T, bool PRECISE, typename HT>
int launch_gn_ln(cfm_handle* h, const Res& R, const NormW& gn, const double* stats, const void* resid, const NormW& ln, cudaStream_t s) {
  const int C = h->C();
  const int blocks = (R.M * 32 + 255) / 256;
#define CFM_GN_LN(NCH)                                                                                                     \
  return launch_ex(h, gn_apply_ln_kernel<T, PRECISE, NCH, HT>, dim3(blocks), dim3(256), 0, s, 1, (const HT*)R.hraw, (long long)C, \
                   R.M, C / 8, (const int*)R.info, stats, (const double*)gn.bias_gsum, (const UttTable*)R.utt, (const float*)gn.gamma, (const float*)gn.beta, (const HT*)resid, (long long)C, R.X,   \
                   (long long)C, (const float*)ln.gamma, (const float*)ln.beta, static_cast<T*>(R.Xn), (long long)C)
  switch (C / 128) {
    case 1: CFM_GN_LN(1);
    case 2: CFM_GN_LN(2);
    case 3: CFM_GN_LN(3);
    default: CFM_GN_LN(4);
  }
#undef CFM_GN_LN
}

// `fuse_ln` != nullptr (block2 of a resnet followed by a transformer stack, C % 128 == 0): also emits LayerNorm(out) -> R.Xn.
int run_gn_apply(cfm_handle* h, Plan* pl, const Res& R, const NormW& gn, int site, const float* addvec, const void* resid,
                 float* out_f32, void* out_act, long long ld_act, cudaStream_t s, const NormW* fuse_ln = nullptr,
                 long long addvec_stride = 0) {
  if (h->stopped()) return 0;
  h->launch_counter++;
  const int C = h->C();
  CKR(tl_mark(h, s, fuse_ln ? "gn_apply_ln" : "gn_apply", R.M, C, 0, 0.0));
  const double* stats = pl->stats + (long long)site * pl->B * 16;
  if (fuse_ln) {
    if (mid16(h)) return launch_gn_ln<bf16, false, bf16>(h, R, gn, stats, resid, *fuse_ln, s);
    if (h->bf) return launch_gn_ln<bf16, false, float>(h, R, gn, stats, resid, *fuse_ln, s);
    return launch_gn_ln<float, true, float>(h, R, gn, stats, resid, *fuse_ln, s);
  }
  const long long items = (long long)R.M * (C / 8);
  const int blocks = (int)((items + 255) / 256);
  if (mid16(h))
    return launch_ex(h, gn_apply_kernel<bf16, false, bf16>, dim3(blocks), dim3(256), 0, s, 1, (const bf16*)R.hraw, (long long)C, R.M, C, C / 8,
                     (const int*)R.info, stats, (const double*)gn.bias_gsum, (const UttTable*)R.utt, (const float*)gn.gamma, (const float*)gn.beta, addvec, addvec_stride, (const bf16*)resid, (long long)C,
                     out_f32, (long long)C, static_cast<bf16*>(out_act), ld_act);
  if (h->bf)
    return launch_ex(h, gn_apply_kernel<bf16, false, float>, dim3(blocks), dim3(256), 0, s, 1, (const float*)R.hraw, (long long)C, R.M, C, C / 8,
                     (const int*)R.info, stats, (const double*)gn.bias_gsum, (const UttTable*)R.utt, (const float*)gn.gamma, (const float*)gn.beta, addvec, addvec_stride, (const float*)resid, (long long)C,
                     out_f32, (long long)C, static_cast<bf16*>(out_act), ld_act);
  return launch_ex(h, gn_apply_kernel<float, true, float>, dim3(blocks), dim3(256), 0, s, 1, (const float*)R.hraw, (long long)C, R.M, C, C / 8,
                   (const int*)R.info, stats, (const double*)gn.bias_gsum, (const UttTable*)R.utt, (const float*)gn.gamma, (const float*)gn.beta, addvec, addvec_stride, (const float*)resid, (long long)C,
                   out_f32, (long long)C, static_cast<float*>(out_act), ld_act);
}

// Conv k=3 (+bias) -> fp32 raw output + GroupNorm statistics for site `site`.
int run_conv_stats(cfm_handle* h, Plan* pl, const Res& R, const void* A, long long lda, const GemmW& w, int site,
                   cudaStream_t s) {
  static const int shifts[3] = {-1, 0, 1};
  GemmParams p = gemm_base(R.M, A, lda, R.M, w, shifts, nullptr);
  p.mode = EPI_STATS;
  h->tag = w.K >= 2 * h->C() ? "conv3_stats_2C" : (w.K == h->C() ? "conv3_stats_C" : "conv3_stats_in");
  p.out_f32 = R.hraw, p.ld_f32 = h->C();
  if (mid16(h)) p.mode = EPI_STATS16, p.out_act = R.hraw, p.ld_act = h->C();  // bf16 result in the same buffer
  p.row_info = R.info;
  p.stats = pl->stats + (long long)site * pl->B * 16;
  p.group_ch = h->C() / 8;
  const bool tc = (h->bf && !(h->cfg.flags & CFM_FLAG_SIMT_GEMM)) || x3_ok(h, p);  // x3: sums from the fp32 accumulators, like bf16
  p.fused_stats = (tc && !(h->cfg.flags & CFM_FLAG_UNFUSED_STATS)) ? 1 : 0;
  CKR(launch_gemm(h, p, true, s));
  if (!p.fused_stats && !h->stopped()) {
    h->launch_counter++;
    CKR(tl_mark(h, s, "gn_stats", R.M, h->C(), 0, 0.0));
    const int items = R.M * 8;
    gn_stats_kernel<float><<<(items + 255) / 256, 256, 0, s>>>(R.hraw, h->C(), R.M, h->C(), h->C() / 8, R.info, p.stats);
    CK(cudaGetLastError());
  }
  return 0;
}

template <typename T>
int launch_ln_vec(cfm_handle* h, const Res& R, const NormW& ln, cudaStream_t s) {
  const int C = h->C();
  const int blocks = (R.M * 32 + 255) / 256;
#define CFM_LN_VEC(NCH)                                                                                                        \
  return launch_ex(h, layernorm_vec_kernel<T, NCH>, dim3(blocks), dim3(256), 0, s, 1, (const float*)R.X, (long long)C, R.M,    \
                   (const float*)ln.gamma, (const float*)ln.beta, static_cast<T*>(R.Xn), (long long)C)
  switch (C / 128) {
    case 1: CFM_LN_VEC(1);
    case 2: CFM_LN_VEC(2);
    case 3: CFM_LN_VEC(3);
    default: CFM_LN_VEC(4);
  }
#undef CFM_LN_VEC
}

int run_layernorm(cfm_handle* h, const Res& R, const NormW& ln, cudaStream_t s) {
  if (h->stopped()) return 0;
  h->launch_counter++;
  const int C = h->C();
  CKR(tl_mark(h, s, "layernorm", R.M, C, 0, 0.0));
  if (C % 128 == 0 && C <= 512) {
    if (h->bf) return launch_ln_vec<bf16>(h, R, ln, s);
    return launch_ln_vec<float>(h, R, ln, s);
  }
  const int blocks = (R.M * 32 + 255) / 256;
  if (h->bf)
    return launch_ex(h, layernorm_kernel<bf16, 16>, dim3(blocks), dim3(256), 0, s, 1, (const float*)R.X, (long long)C, R.M, C,
                     (const float*)ln.gamma, (const float*)ln.beta, static_cast<bf16*>(R.Xn), (long long)C);
  return launch_ex(h, layernorm_kernel<float, 16>, dim3(blocks), dim3(256), 0, s, 1, (const float*)R.X, (long long)C, R.M, C,
                   (const float*)ln.gamma, (const float*)ln.beta, static_cast<float*>(R.Xn), (long long)C);
}

int run_attention(cfm_handle* h, const Res& R, cudaStream_t s) {
  if (h->stopped()) return 0;
  h->launch_counter++;
  const int I = h->inner(), D = h->cfg.head_dim;
  if (h->tl_on) {  // 4 (L + 1)^2 H d per utterance of the lane (QK^T and PV, pad token included)
    double fl = 0.0;
    const bool full = R.M_all == h->plan->M1;
    for (int b = R.b0; b < R.b0 + R.nb; ++b) {
      const double nk = (full ? h->plan->L[b] : (h->plan->L[b] + 1) / 2) + 1;
      fl += 4.0 * nk * nk * I;
    }
    CKR(tl_mark(h, s, "attention", R.M, I, 0, fl));
  }
  const float scale = 1.0f / sqrtf((float)D);
  if (!h->bf && h->fp32_tc && D == 64 && !(h->cfg.flags & CFM_FLAG_SIMT_ATTN) && h->plan && h->plan->split3 && 6LL * I <= 8LL * h->C()) {
    // fp32 mode on the tensor pipe: bf16 [hi | lo] copy of the QKV rows, then the split-operand attention kernel (fp32 output)
    const long long quads = (long long)R.M_all * (3 * I / 4);
    h->launch_counter++;
    CKR(launch_ex(h, split3_rows_kernel, dim3((unsigned)((quads + 255) / 256)), dim3(256), 0, s, 1, static_cast<const float*>(R.qkv_all), 3LL * I,
                  (long long)R.M_all, static_cast<bf16*>(h->plan->split3)));
    return launch_attn_tc(h->encode, h->plan->split3, 3LL * I, I, R.M_all, R.utt, R.work, R.n_work, R.ao_all, I, scale, s, &h->err, nullptr,
                          h->pdl_now != 0, true);
  }
  const bool tc = h->bf && D == 64 && !(h->cfg.flags & CFM_FLAG_SIMT_ATTN);
  if (tc && h->attn_persist && !h->attn_prof)
    return launch_attn_persist(h->encode, R.qkv_all, 3LL * I, I, R.M_all, R.utt, R.work, R.n_work, R.ao_all, I, scale, 2 * h->sm_count, s, &h->err,
                               h->pdl_now != 0);
  if (tc) return launch_attn_tc(h->encode, R.qkv_all, 3LL * I, I, R.M_all, R.utt, R.work, R.n_work, R.ao_all, I, scale, s, &h->err, h->attn_prof, h->pdl_now != 0);
  {
    const KernelInfo k = h->bf ? (D == 64 ? kinfo_attn_simt_bf16_64() : kinfo_attn_simt_bf16_32())
                               : (D == 64 ? kinfo_attn_simt_f32_64() : kinfo_attn_simt_f32_32());
    const void* qkv = R.qkv_all;
    long long ld = 3LL * I, ldo = I;
    int inner = I;
    const UttTable* utt = R.utt;
    const int4* work = R.work;
    void* out = R.ao_all;
    float sc = scale;
    void* args[] = {&qkv, &ld, &inner, &utt, &work, &out, &ldo, &sc};
    CK(cudaLaunchKernel(k.fn, dim3(R.n_work), dim3(k.threads), args, 0, s));
  }
  CK(cudaGetLastError());
  return 0;
}

// ResnetBlock1D (reference decoder.py:48-63) on masked input A (K columns) -> fp32 residual stream R.X.
int run_resnet(cfm_handle* h, Plan* pl, const Res& R, const ResnetW& w, const void* A, long long lda, int& site,
               const float* tproj, long long tproj_stride, cudaStream_t s, const NormW* fuse_ln) {
  const int C = h->C();
  // res_conv only meets the main chain again at the second GroupNorm-apply: as a side branch (its own stream, fork / join events =
  // graph edges) it can fill the partly idle last round of conv1 / conv2 and run under the bandwidth-bound apply pass
  const bool side = (h->res_side < 0 ? pl->M1 <= 2048 : h->res_side != 0) && h->side_stream && pl->lanes.size() == 1 && !h->tl_on &&
                    h->stop_after < 0;
  cudaStream_t sr = side ? h->side_stream : s;
  if (side) {
    CK(cudaEventRecord(h->side_fork, s));
    CK(cudaStreamWaitEvent(sr, h->side_fork, 0));
  }
  CKR(run_conv_stats(h, pl, R, A, lda, w.conv1, site, s));
  {  // res_conv (1x1) on the same masked input -> fp32
    GemmParams p = gemm_base(R.M, A, lda, R.M, w.res, nullptr, nullptr);
    p.mode = EPI_STATS, p.fused_stats = 0;
    p.out_f32 = R.rres, p.ld_f32 = C;
    if (mid16(h)) p.mode = EPI_STORE, p.out_act = R.rres, p.ld_act = C;  // bf16 result, direct-store epilogue
    h->tag = "res_conv";
    CKR(launch_gemm(h, p, true, sr));
  }
  if (side) CK(cudaEventRecord(h->side_join, sr));
  CKR(run_gn_apply(h, pl, R, w.gn1, site, tproj, nullptr, nullptr, R.hact, C, s, nullptr, tproj_stride));
  site++;
  CKR(run_conv_stats(h, pl, R, R.hact, C, w.conv2, site, s));
  if (side) CK(cudaStreamWaitEvent(s, h->side_join, 0));
  CKR(run_gn_apply(h, pl, R, w.gn2, site, nullptr, R.rres, R.X, nullptr, 0, s, fuse_ln));
  site++;
  return 0;
}

// FeedForward (Linear -> SnakeBeta -> Linear, + residual) as one kernel: ff_fused.cuh.
int run_ff_fused(cfm_handle* h, const Res& R, const BlockW& w, void* copy_dst, long long copy_ld, cudaStream_t s) {
  if (h->stopped()) return 0;
  h->launch_counter++;
  const int C = h->C();
  CKR(tl_mark(h, s, copy_dst ? "ff_fused_copy" : "ff_fused", R.M, C, 4 * C, 2.0 * 2.0 * R.M * C * 4.0 * C));
  CUtensorMap tmA, tmW1, tmW2, tmX;
  CKR(make_tmap(h, &tmA, R.Xn, C, R.M, (long long)C * 2, 64, 128));
  CKR(make_out_tmap(h, &tmX, R.X, true, C, R.M, C));
  CKR(make_tmap(h, &tmW1, w.ff1.w, w.ff1.Kp, w.ff1.n_stride, (long long)w.ff1.Kp * 2, 64, FfCfg::HC / 2));
  CKR(make_tmap(h, &tmW2, w.ff2.w, w.ff2.Kp, w.ff2.n_stride, (long long)w.ff2.Kp * 2, 64, C / 4));
  FfParams p;
  memset(&p, 0, sizeof p);
  p.M = R.M, p.C = C;
  p.b1 = w.ff1.bias, p.ea = w.ea, p.ib = w.ib, p.b2 = w.ff2.bias;
  p.X = R.X, p.ldx = C;
  p.copy = static_cast<bf16*>(copy_dst), p.ld_copy = copy_ld, p.row_info = R.info;
  p.prof = h->ff_prof;
  const KernelInfo k = kinfo_ff_fused();
  const int n_tiles = (R.M + 255) / 256;
  const int pairs = std::min(n_tiles, h->max_clusters[2]);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(pairs * 2), cfg.blockDim = dim3(k.threads), cfg.dynamicSmemBytes = FfCfg::smem_bytes(C), cfg.stream = s;
  cudaLaunchAttribute attr[3];
  int n = 0;
  if (h->win_bytes > 0) {
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = h->win_ptr;
    attr[n].val.accessPolicyWindow.num_bytes = h->win_bytes;
    attr[n].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  attr[n].id = cudaLaunchAttributeClusterDimension;
  attr[n].val.clusterDim.x = 2, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
  ++n;
  if (h->pdl_now) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr, cfg.numAttrs = n;
  static thread_local FfConsts cst;  // 18 KB by-value kernel parameter (copied by the launch / into the graph node)
  memcpy(cst.b1, w.h_b1.data(), sizeof(float) * 4 * C);
  memcpy(cst.ea, w.h_ea.data(), sizeof(float) * 4 * C);
  memcpy(cst.ib, w.h_ib.data(), sizeof(float) * 4 * C);
  void* args[] = {&tmA, &tmW1, &tmW2, &tmX, &p, &cst};
  CK(cudaLaunchKernelExC(&cfg, k.fn, args));
  return 0;
}

// Linear + residual add + LayerNorm as one kernel (rowln.cuh): X += A . W^T + b, Xn = LayerNorm(X) gamma + beta.
bool rowln_ok(cfm_handle* h, const Res& R) {
  const int C = h->C();
  const bool tc = h->bf && !(h->cfg.flags & CFM_FLAG_SIMT_GEMM);
  return tc && C % 128 == 0 && C <= RowLnCfg::MAX_N && (h->rowln == 2 || (h->rowln == 1 && R.M > h->small_tiles));
}
int run_rowln(cfm_handle* h, const Res& R, const void* A, long long lda, const GemmW& w, const NormW& ln, const char* tag, cudaStream_t s) {
  if (h->stopped()) return 0;
  h->launch_counter++;
  const int C = h->C();
  CKR(tl_mark(h, s, tag, R.M, C, w.Kp, 2.0 * R.M * C * w.Kp));
  CUtensorMap tmA, tmW, tmX, tmN;
  CKR(make_tmap(h, &tmA, A, w.Kp, R.M, lda * 2, 64, 128));
  CKR(make_tmap(h, &tmW, w.w, w.Kp, w.n_stride, (long long)w.Kp * 2, 64, C / 2));
  CKR(make_out_tmap(h, &tmX, R.X, true, C, R.M, C));
  CKR(make_out_tmap(h, &tmN, R.Xn, false, C, R.M, C));
  RowLnParams p;
  memset(&p, 0, sizeof p);
  p.M = R.M, p.N = C, p.K = w.Kp;
  p.bias = w.bias, p.gamma = ln.gamma, p.beta = ln.beta;
  p.X = R.X, p.ldx = C, p.Xn = static_cast<bf16*>(R.Xn), p.ldn = C, p.eps = 1e-5f;
  p.prof = h->rowln_prof;
  const KernelInfo k = kinfo_rowln();
  const int m_tiles = (R.M + RowLnCfg::BM - 1) / RowLnCfg::BM;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(std::min(m_tiles, h->max_clusters[1])), cfg.blockDim = dim3(k.threads), cfg.dynamicSmemBytes = RowLnCfg::smem_bytes(C), cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (h->win_bytes > 0) {
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = h->win_ptr;
    attr[n].val.accessPolicyWindow.num_bytes = h->win_bytes;
    attr[n].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  if (h->pdl_now) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr, cfg.numAttrs = n;
  void* args[] = {&tmA, &tmW, &tmX, &tmN, &p};
  CK(cudaLaunchKernelExC(&cfg, k.fn, args));
  return 0;
}

// BasicTransformerBlock (reference transformer.py:230-303).  If copy_dst != nullptr the FF2 epilogue also writes the
// masked activation-type copy of the block output there (skip connection / next conv input).
// next_ln1: norm1 of the transformer block that follows in the same stack (nullptr for the last one); *next_ln1_done is set when
// this block's FF2 launch has already produced it.
int run_block(cfm_handle* h, const Res& R, const BlockW& w, void* copy_dst, long long copy_ld, cudaStream_t s, bool ln1_done,
              const NormW* next_ln1 = nullptr, bool* next_ln1_done = nullptr) {
  const int C = h->C(), I = h->inner();
  const bool fuse_ln = rowln_ok(h, R) && I % 64 == 0;
  if (next_ln1_done) *next_ln1_done = false;
  if (!ln1_done) CKR(run_layernorm(h, R, w.ln1, s));
  {
    GemmParams p = gemm_base(R.M, R.Xn, C, R.M, w.qkv, nullptr, nullptr);
    p.mode = EPI_STORE, p.out_act = R.qkv, p.ld_act = 3 * I;
    h->tag = "qkv";
    CKR(launch_gemm(h, p, true, s));
  }
  CKR(run_attention(h, R, s));
  if (fuse_ln) {
    CKR(run_rowln(h, R, R.ao, I, w.out, w.ln3, "out_proj_ln", s));
  } else {
    GemmParams p = gemm_base(R.M, R.ao, I, R.M, w.out, nullptr, nullptr);
    p.mode = EPI_RESID, p.resid = R.X, p.ld_resid = C, p.out_f32 = R.X, p.ld_f32 = C;
    h->tag = "out_proj";
    CKR(launch_gemm(h, p, true, s));
    CKR(run_layernorm(h, R, w.ln3, s));
  }
  const bool tc = h->bf && !(h->cfg.flags & CFM_FLAG_SIMT_GEMM);
  if (tc && h->ff_fused && C % 64 == 0 && C <= FfCfg::MAX_C && R.M > h->small_tiles) return run_ff_fused(h, R, w, copy_dst, copy_ld, s);
  {
    GemmParams p = gemm_base(R.M, R.Xn, C, R.M, w.ff1, nullptr, nullptr);
    p.mode = EPI_SNAKE, p.ea = w.ea, p.ib = w.ib, p.out_act = R.ffh, p.ld_act = 4 * C;
    h->tag = "ff1_snake";
    CKR(launch_gemm(h, p, true, s));
  }
  if (fuse_ln && next_ln1 && !copy_dst && (h->rowln_ff2 != 0)) {
    CKR(run_rowln(h, R, R.ffh, 4 * C, w.ff2, *next_ln1, "ff2_ln", s));
    if (next_ln1_done) *next_ln1_done = true;
    return 0;
  }
  {
    GemmParams p = gemm_base(R.M, R.ffh, 4 * C, R.M, w.ff2, nullptr, nullptr);
    p.mode = EPI_RESID, p.resid = R.X, p.ld_resid = C, p.out_f32 = R.X, p.ld_f32 = C;
    p.out_act = copy_dst, p.ld_act = copy_ld, p.row_info = R.info;
    h->tag = copy_dst ? "ff2_copy" : "ff2";
    CKR(launch_gemm(h, p, true, s));
  }
  return 0;
}

int run_stage(cfm_handle* h, Plan* pl, const Res& R, const StageW& w, const void* A, long long lda, int& site,
              const float* tproj, long long tproj_stride, void* copy_dst, long long copy_ld, cudaStream_t s) {
  if (h->l2_persist_mb > 0) {  // L2 access-policy window = this stage's fp32 residual stream (read / updated 6x per block)
    h->win_ptr = R.X;
    h->win_bytes = std::min((size_t)R.M * h->C() * sizeof(float), h->win_max);
  }
  // the resnet's last GroupNorm-apply also produces LayerNorm1 of the first transformer block when the width allows
  const bool fuse = h->C() % 128 == 0 && h->C() <= 512 && !w.blocks.empty() && !(h->cfg.flags & CFM_FLAG_UNFUSED_STATS);
  CKR(run_resnet(h, pl, R, w.res, A, lda, site, tproj, tproj_stride, s, fuse ? &w.blocks[0].ln1 : nullptr));
  bool ln1_done = fuse;
  for (size_t j = 0; j < w.blocks.size(); ++j) {
    const bool last = j + 1 == w.blocks.size();
    bool next_done = false;
    CKR(run_block(h, R, w.blocks[j], last ? copy_dst : nullptr, copy_ld, s, ln1_done, last ? nullptr : &w.blocks[j + 1].ln1, &next_done));
    ln1_done = next_done;
  }
  return 0;
}

// One estimator evaluation of one lane + the ODE stage update fused in final_proj's epilogue.
// per_utt_t: the time-embedding rows are indexed by utterance (one t per sample) instead of by NFE.
int emit_nfe(cfm_handle* h, Plan* pl, const LaneDef& ln, int nfe_index, const OdeStage& st, float* out_f32_override, cudaStream_t s,
             bool per_utt_t = false) {
  const int C = h->C(), F = h->cfg.out_channels, es = h->es;
  const Model& m = h->model;
  Res R1 = res_of(h, pl, 0, ln), R2 = res_of(h, pl, 1, ln);
  const long long row0 = 2LL * ln.r2;  // first full-resolution row of the lane
  h->launch_counter++;
  CKR(tl_mark(h, s, "memset_stats", 0, 0, 0, 0.0));
  {  // zero the lane's GroupNorm sums of every site: stats is [site][B][16] doubles
    const int n_sites = 2 * (4 + h->cfg.n_mid_blocks) + 1;
    CK(cudaMemset2DAsync(pl->stats + (long long)ln.b0 * 16, (size_t)pl->B * 16 * sizeof(double), 0, (size_t)ln.nb * 16 * sizeof(double),
                         n_sites, s));
  }
  int site = 0;
  int stage_i = 0;
  const int n_res = 4 + h->cfg.n_mid_blocks;
  auto tproj = [&](int r) { return pl->tproj + ((long long)(per_utt_t ? 0 : nfe_index) * n_res + r) * C; };
  const long long tps = per_utt_t ? (long long)n_res * C : 0;
  void* xin = act_off(pl->xin, row0 * pl->xin_ld, es);

  // down 0 (full resolution): [x | mu] -> X1; masked copy of the stage output -> right half of cat1 (skip h0)
  CKR(run_stage(h, pl, R1, m.stages[stage_i], xin, pl->xin_ld, site, tproj(stage_i), tps, act_off(R1.cat, C, es), 2 * C, s));
  stage_i++;
  {  // Downsample1D: Conv1d k3 s2 p1 on the masked skip, rows viewed in pairs [M2, 4C]
    const int shifts[3] = {-1, 0, 0};
    const int acols[3] = {3 * C, C, 3 * C};
    GemmParams p = gemm_base(R2.M, R1.cat, 4LL * C, R2.M, m.down_s2, shifts, acols);
    p.mode = EPI_MASK, p.row_info = R2.info, p.out_act = R2.sin_, p.ld_act = C;
    h->tag = "conv_s2";
    CKR(launch_gemm(h, p, true, s));
  }
  // down 1 (half resolution); masked copy -> right half of cat2 (skip h1)
  CKR(run_stage(h, pl, R2, m.stages[stage_i], R2.sin_, C, site, tproj(stage_i), tps, act_off(R2.cat, C, es), 2 * C, s));
  stage_i++;
  {  // stage-final Conv1d k3 on the masked skip
    const int shifts[3] = {-1, 0, 1};
    const int acols[3] = {C, C, C};
    GemmParams p = gemm_base(R2.M, R2.cat, 2LL * C, R2.M, m.down_tail, shifts, acols);
    p.mode = EPI_MASK, p.row_info = R2.info, p.out_act = R2.sin_, p.ld_act = C;
    h->tag = "conv3_tail";
    CKR(launch_gemm(h, p, true, s));
  }
  for (int i = 0; i < h->cfg.n_mid_blocks; ++i) {
    const bool last = i + 1 == h->cfg.n_mid_blocks;
    void* dst = last ? R2.cat : R2.sin_;  // last mid block feeds the left half of cat2
    CKR(run_stage(h, pl, R2, m.stages[stage_i], R2.sin_, C, site, tproj(stage_i), tps, dst, last ? 2 * C : C, s));
    stage_i++;
  }
  // up 0 (half resolution) on [x | h1]
  CKR(run_stage(h, pl, R2, m.stages[stage_i], R2.cat, 2 * C, site, tproj(stage_i), tps, R2.sin_, C, s));
  stage_i++;
  for (int phase = 0; phase < 2; ++phase) {  // ConvTranspose1d k4 s2 p1 as two interleaved 2-tap GEMMs
    const int sh_even[2] = {-1, 0}, sh_odd[2] = {0, 1};
    GemmParams p = gemm_base(R2.M, R2.sin_, C, R2.M, phase == 0 ? m.up_even : m.up_odd, phase == 0 ? sh_even : sh_odd, nullptr);
    p.mode = EPI_MASK, p.row_info = R1.info, p.row_mul = 2, p.row_add = phase;
    p.out_act = act_off(R1.cat, (long long)phase * 2 * C, es), p.ld_act = 4 * C;
    h->tag = "conv_transpose";
    CKR(launch_gemm(h, p, true, s));
  }
  // up 1 (full resolution) on [x | h0]
  CKR(run_stage(h, pl, R1, m.stages[stage_i], R1.cat, 2 * C, site, tproj(stage_i), tps, R1.sin_, C, s));
  stage_i++;
  {  // stage-final Conv1d k3
    const int shifts[3] = {-1, 0, 1};
    GemmParams p = gemm_base(R1.M, R1.sin_, C, R1.M, m.up_tail, shifts, nullptr);
    p.mode = EPI_MASK, p.row_info = R1.info, p.out_act = R1.hact, p.ld_act = C;
    h->tag = "conv3_tail";
    CKR(launch_gemm(h, p, true, s));
  }
  // final Block1D + 1x1 projection with the ODE update in the epilogue
  CKR(run_conv_stats(h, pl, R1, R1.hact, C, m.final_conv, site, s));
  CKR(run_gn_apply(h, pl, R1, m.final_gn, site, nullptr, nullptr, nullptr, R1.sin_, C, s));
  site++;
  {
    GemmParams p = gemm_base(R1.M, R1.sin_, C, R1.M, m.final_proj, nullptr, nullptr);
    p.mode = EPI_ODE, p.row_info = R1.info;
    p.c_v = st.c_v;
    p.ld_k = F;
    for (int j = 0; j < 3; ++j) {
      p.c_k[j] = st.c_k[j];
      p.kin[j] = st.kin[j] >= 0 ? pl->kbuf[st.kin[j]] + row0 * F : nullptr;
    }
    p.kout = st.kout >= 0 ? pl->kbuf[st.kout] + row0 * F : nullptr;
    if (out_f32_override) {  // bare estimator call: v itself
      p.resid = nullptr, p.out_f32 = out_f32_override + row0 * F, p.ld_f32 = F, p.out_act = nullptr;
    } else {
      p.resid = pl->xstate + row0 * F, p.ld_resid = F;
      p.out_f32 = st.write_state ? pl->xstate + row0 * F : nullptr, p.ld_f32 = F;
      p.out_act = xin, p.ld_act = pl->xin_ld;
    }
    h->tag = "final_proj_ode";
    CKR(launch_gemm(h, p, true, s));
  }
  return 0;
}

// Time-embedding MLP for all NFE time points + the per-resnet projections (reference decoder.py:368-369, :51).
int emit_time_embedding(cfm_handle* h, Plan* pl, int n_t, const float* t_dev, cudaStream_t s) {
  const int C = h->C(), T4 = 4 * C, IC = h->cfg.in_channels;
  const Model& m = h->model;
  auto gemv = [&](const float* x, long long ldx, const float* W, int N, int K, const float* b, int ai, int ao, float* y,
                  long long ldy) -> int {
    h->launch_counter++;
    CKR(tl_mark(h, s, "time_mlp", n_t, N, K, 2.0 * n_t * N * K));
    const long long warps = (long long)n_t * N;
    gemv_rows_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, s>>>(x, ldx, n_t, W, N, K, b, ai, ao, y, ldy);
    CK(cudaGetLastError());
    return 0;
  };
  h->launch_counter++;
  sinusoid_kernel<<<n_t, roundup(IC / 2, 32), 0, s>>>(t_dev, n_t, IC, pl->sinemb);
  CK(cudaGetLastError());
  CKR(gemv(pl->sinemb, IC, m.t1_w, T4, IC, m.t1_b, ACT_NONE, ACT_SILU, pl->temb_a, T4));
  CKR(gemv(pl->temb_a, T4, m.t2_w, T4, T4, m.t2_b, ACT_NONE, ACT_NONE, pl->temb, T4));
  const int n_res = (int)m.stages.size();
  for (int r = 0; r < n_res; ++r)
    CKR(gemv(pl->temb, T4, m.stages[r].res.mlp_w, C, T4, m.stages[r].res.mlp_b, ACT_MISH, ACT_NONE, pl->tproj + (long long)r * C,
             (long long)n_res * C));
  return 0;
}

int emit_pack(cfm_handle* h, Plan* pl, const float* x, const float* mu, cudaStream_t s) {
  const int F = h->cfg.out_channels;
  int max_rows = 0;
  for (int L : pl->L) max_rows = std::max(max_rows, 2 * ((L + 1) / 2 + 2));
  dim3 grid((max_rows + 31) / 32, (F + 31) / 32, pl->B), block(32, 8);
  h->launch_counter += 2;
  if (h->bf) {
    pack_rows_kernel<bf16><<<grid, block, 0, s>>>(x, F, pl->T, pl->utt1, static_cast<bf16*>(pl->xin), pl->xin_ld, 0, pl->xstate, F, 1.f);
    pack_rows_kernel<bf16><<<grid, block, 0, s>>>(mu, F, pl->T, pl->utt1, static_cast<bf16*>(pl->xin), pl->xin_ld, F, nullptr, 0, 1.f);
  } else {
    pack_rows_kernel<float><<<grid, block, 0, s>>>(x, F, pl->T, pl->utt1, static_cast<float*>(pl->xin), pl->xin_ld, 0, pl->xstate, F, 1.f);
    pack_rows_kernel<float><<<grid, block, 0, s>>>(mu, F, pl->T, pl->utt1, static_cast<float*>(pl->xin), pl->xin_ld, F, nullptr, 0, 1.f);
  }
  CK(cudaGetLastError());
  const int S = h->cfg.in_channels - 2 * F;
  if (S > 0) {
    if (!h->spks) return fail(h, CFM_ERR_STATE, "this estimator has %d speaker channels: call cfm_set_speakers before the decode", S);
    h->launch_counter++;
    dim3 g2((max_rows * S + 255) / 256, pl->B);
    if (h->bf) pack_speaker_kernel<bf16><<<g2, 256, 0, s>>>(h->spks, S, pl->utt1, static_cast<bf16*>(pl->xin), pl->xin_ld, 2 * F);
    else pack_speaker_kernel<float><<<g2, 256, 0, s>>>(h->spks, S, pl->utt1, static_cast<float*>(pl->xin), pl->xin_ld, 2 * F);
    CK(cudaGetLastError());
  }
  return 0;
}

int emit_unpack(cfm_handle* h, Plan* pl, const float* state, const float* fill, float* out, cudaStream_t s) {
  const int F = h->cfg.out_channels;
  dim3 grid((pl->T + 31) / 32, (F + 31) / 32, pl->B), block(32, 8);
  h->launch_counter++;
  unpack_rows_kernel<<<grid, block, 0, s>>>(state, F, pl->utt1, F, pl->T, fill, out);
  CK(cudaGetLastError());
  return 0;
}

// Runs fn(lane, stream) for every lane: lane 0 on `s`, the others on the handle's lane streams, forked from and joined
// back into `s` with events (inside a stream capture this makes the lanes parallel branches of the graph).
template <typename Fn>
int for_each_lane(cfm_handle* h, Plan* pl, cudaStream_t s, Fn&& fn) {
  const int n = (int)pl->lanes.size();
  if (n > 1) {
    CK(cudaEventRecord(h->fork_event, s));
    for (int i = 1; i < n; ++i) CK(cudaStreamWaitEvent(h->lane_streams[i - 1], h->fork_event, 0));
  }
  int r = 0;
  for (int i = 0; i < n && r == 0; ++i) r = fn(pl->lanes[i], i == 0 ? s : h->lane_streams[i - 1]);
  for (int i = 1; i < n; ++i) {  // always join, so that a failed capture does not leave unjoined streams behind
    cudaError_t e1 = cudaEventRecord(h->lane_events[i - 1], h->lane_streams[i - 1]);
    cudaError_t e2 = cudaStreamWaitEvent(s, h->lane_events[i - 1], 0);
    if (r == 0 && (e1 != cudaSuccess || e2 != cudaSuccess)) r = fail(h, CFM_ERR_CUDA, "lane join failed");
  }
  return r;
}

int emit_ode_loop(cfm_handle* h, Plan* pl, cudaStream_t s) {
  CKR(emit_time_embedding(h, pl, (int)pl->stages.size(), pl->tvals, s));
  return for_each_lane(h, pl, s, [&](const LaneDef& ln, cudaStream_t ls) -> int {
    for (size_t j = 0; j < pl->stages.size(); ++j) CKR(emit_nfe(h, pl, ln, (int)j, pl->stages[j], nullptr, ls));
    return 0;
  });
}

int build_stages(cfm_handle* h, Plan* pl, const float* t_span, int n_points, int solver) {
  pl->stages.clear();
  for (int i = 0; i + 1 < n_points; ++i) {
    const float t0 = t_span[i], t1 = t_span[i + 1];
    const float dt = t1 - t0;
    auto mk = [&](float t, float cv) {
      OdeStage st;
      st.t = t, st.c_v = cv, st.kout = -1, st.write_state = false;
      for (int j = 0; j < 3; ++j) st.c_k[j] = 0.f, st.kin[j] = -1;
      return st;
    };
    if (solver == CFM_SOLVER_EULER) {
      OdeStage a = mk(t0, dt);
      a.write_state = true;
      pl->stages.push_back(a);
    } else if (solver == CFM_SOLVER_MIDPOINT) {
      const float half = 0.5f * dt;
      pl->stages.push_back(mk(t0, half));
      OdeStage b = mk(t0 + half, dt);
      b.write_state = true;
      pl->stages.push_back(b);
    } else if (solver == CFM_SOLVER_HEUN3) {
      OdeStage a = mk(t0, dt / 3.f);
      a.kout = 0;
      OdeStage b = mk(t0 + dt / 3.f, dt * 2.f / 3.f);
      OdeStage c = mk(t0 + dt * 2.f / 3.f, dt * 0.75f);
      c.kin[0] = 0, c.c_k[0] = dt * 0.25f, c.write_state = true;
      pl->stages.push_back(a), pl->stages.push_back(b), pl->stages.push_back(c);
    } else if (solver == CFM_SOLVER_RK4) {  // torchdiffeq's 3/8-rule
      OdeStage a = mk(t0, dt / 3.f);
      a.kout = 0;
      OdeStage b = mk(t0 + dt / 3.f, dt);
      b.kin[0] = 0, b.c_k[0] = -dt / 3.f, b.kout = 1;
      OdeStage c = mk(t0 + dt * 2.f / 3.f, dt);
      c.kin[0] = 0, c.c_k[0] = dt, c.kin[1] = 1, c.c_k[1] = -dt, c.kout = 2;
      OdeStage d = mk(t1, dt * 0.125f);
      d.kin[0] = 0, d.c_k[0] = dt * 0.125f, d.kin[1] = 1, d.c_k[1] = dt * 0.375f, d.kin[2] = 2, d.c_k[2] = dt * 0.375f;
      d.write_state = true;
      pl->stages.push_back(a), pl->stages.push_back(b), pl->stages.push_back(c), pl->stages.push_back(d);
    } else {
      return fail(h, CFM_ERR_INVALID, "unknown solver id %d (expected euler=0, midpoint=1, heun3=2, rk4=3)", solver);
    }
  }
  return 0;
}

// Persisting-L2 carve-out for the fp32 residual stream (launch_ex attaches the access-policy window).  The device limit is
// process-wide state: it is only ever raised here, never lowered below what somebody else configured.
int apply_l2_persist(cfm_handle* h, int mb) {
  h->l2_persist_mb = mb, h->win_bytes = 0, h->win_max = 0;
  if (mb <= 0) return 0;
  int max_persist = 0, max_win = 0;
  CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->cfg.device));
  CK(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, h->cfg.device));
  const size_t want = std::min((size_t)mb << 20, (size_t)max_persist);
  size_t cur = 0;
  CK(cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize));
  if (cur < want) CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
  h->win_max = std::min(want, (size_t)max_win);
  return 0;
}

int capture_graph(cfm_handle* h, Plan* pl) {
  if (h->cfg.flags & CFM_FLAG_NO_GRAPH) return 0;
  const long long saved = h->launch_counter;
  h->launch_counter = 0;
  CK(cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeThreadLocal));
  int r = emit_ode_loop(h, pl, h->own_stream);
  cudaError_t e = cudaStreamEndCapture(h->own_stream, &pl->graph);
  if (r) return r;
  if (e != cudaSuccess) return fail(h, CFM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  CK(cudaGraphInstantiate(&pl->exec, pl->graph, 0));
  pl->launches_per_solve = h->launch_counter + 3;  // + pack x, pack mu, unpack
  h->launch_counter = saved;
  return 0;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

const char* cfm_last_error(const cfm_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cfm_create(const cfm_config* cfg, cfm_handle** out) {
  cfm_handle* h = nullptr;
  if (!cfg || !out) return fail(h, CFM_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->channels <= 0 || cfg->channels % 64 != 0 || cfg->channels > 512)
    return fail(h, CFM_ERR_INVALID, "channels must be a multiple of 64 in [64, 512], got %d", cfg->channels);
  if (cfg->head_dim != 64 && cfg->head_dim != 32) return fail(h, CFM_ERR_INVALID, "head_dim must be 64 or 32, got %d", cfg->head_dim);
  if (cfg->n_heads <= 0 || cfg->n_blocks < 1 || cfg->n_mid_blocks < 1 || cfg->out_channels <= 0 ||
      cfg->in_channels < 2 * cfg->out_channels || cfg->in_channels % 2 != 0)
    return fail(h, CFM_ERR_INVALID, "invalid estimator configuration");
  if ((cfg->n_heads * cfg->head_dim) % 64 != 0) return fail(h, CFM_ERR_INVALID, "heads*head_dim must be a multiple of 64");
  if (cfg->out_channels % 4 != 0) return fail(h, CFM_ERR_INVALID, "out_channels must be a multiple of 4");
  if (cfg->precision != CFM_PREC_BF16 && cfg->precision != CFM_PREC_FP32 && cfg->precision != CFM_PREC_FP32_TC)
    return fail(h, CFM_ERR_INVALID, "unknown precision");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, CFM_ERR_CUDA, "no CUDA device: libcfm_b200 has no CPU path");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(h, CFM_ERR_INVALID, "device %d out of range", cfg->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(h, CFM_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(h, CFM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major,
                prop.minor);
  h = new cfm_handle();
  h->cfg = *cfg;
  h->bf = cfg->precision == CFM_PREC_BF16;
  h->fp32_tc = cfg->precision == CFM_PREC_FP32_TC ? 1 : 0;
  h->es = h->bf ? 2 : 4;
  h->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("CFM_B200_PDL")) h->pdl = atoi(e) < 0 ? -1 : atoi(e) != 0;
  if (const char* e = getenv("CFM_B200_PAIR")) h->pair_mode = atoi(e);  // 0 never, 1 long-K GEMMs (default), 2 always
  if (const char* e = getenv("CFM_B200_TMA_EPI")) h->tma_epi = atoi(e);
  if (const char* e = getenv("CFM_B200_LANES")) h->lanes_req = std::max(1, std::min(16, atoi(e)));
  if (const char* e = getenv("CFM_B200_LANE_MIN_ROWS")) h->lane_min_rows = std::max(1, atoi(e));
  if (const char* e = getenv("CFM_B200_CLUSTER")) {
    const int c = atoi(e);
    if (c == 1 || c == 2 || c == 4) h->cluster = c;
  }
  auto bail = [&](int code) {
    g_create_error = h->err;
    delete h;
    return code;
  };
  if (cudaSetDevice(cfg->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return bail(CFM_ERR_CUDA); }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
    h->err = "cuTensorMapEncodeTiled not available from the driver";
    return bail(CFM_ERR_CUDA);
  }
  h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  int r = set_gemm_attrs(h);
  r = r ? r : attn_tc_set_attr(&h->err);
  if (r) return bail(r);
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "cudaStreamCreate failed"; return bail(CFM_ERR_CUDA); }
  if (cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->side_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->side_join, cudaEventDisableTiming) != cudaSuccess) { h->err = "side stream / event creation failed"; return bail(CFM_ERR_CUDA); }
  if (cudaEventCreateWithFlags(&h->busy_event, cudaEventDisableTiming) != cudaSuccess) { h->err = "cudaEventCreate failed"; return bail(CFM_ERR_CUDA); }
  if (const char* e = getenv("CFM_B200_ATTN_PERSIST")) h->attn_persist = atoi(e) != 0;
  if (const char* e = getenv("CFM_B200_PLAN_CACHE")) h->plan_cache = std::max(1, atoi(e));
  if (const char* e = getenv("CFM_B200_L2_PERSIST_MB")) h->l2_persist_mb = std::max(0, atoi(e));
  if (apply_l2_persist(h, h->l2_persist_mb) != 0) return bail(CFM_ERR_CUDA);
  *out = h;
  return 0;
}

void cfm_destroy(cfm_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  free_all_plans(h);
  free_arena(h->wallocs);
  for (auto& e : h->tl) cudaEventDestroy(e.ev);
  if (h->busy_event) cudaEventDestroy(h->busy_event);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  for (cudaStream_t st : h->lane_streams) cudaStreamDestroy(st);
  for (cudaEvent_t ev : h->lane_events) cudaEventDestroy(ev);
  if (h->fork_event) cudaEventDestroy(h->fork_event);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->side_fork) cudaEventDestroy(h->side_fork);
  if (h->side_join) cudaEventDestroy(h->side_join);
  delete h;
}

int cfm_load_weights(cfm_handle* h, const cfm_weight_desc* descs, int32_t n) {
  if (!h || !descs || n <= 0) return fail(h, CFM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  free_all_plans(h);  // the captured graphs reference the old weight buffers
  free_arena(h->wallocs);
  h->model = Model();
  h->weights_loaded = false;
  DescMap dm;
  for (int i = 0; i < n; ++i) {
    if (!descs[i].name) return fail(h, CFM_ERR_WEIGHTS, "descriptor %d has no name", i);
    dm.by_name[descs[i].name] = &descs[i];
    dm.used[descs[i].name] = false;
  }
  const int C = h->C(), T4 = 4 * C, IC = h->cfg.in_channels, F = h->cfg.out_channels;
  Model& m = h->model;
  const float* p;
  CKR(want(h, dm, "time_mlp.linear_1.weight", {T4, IC}, &p));
  CKR(copy_f32(h, p, (size_t)T4 * IC, &m.t1_w));
  CKR(want(h, dm, "time_mlp.linear_1.bias", {T4}, &p));
  CKR(copy_f32(h, p, T4, &m.t1_b));
  CKR(want(h, dm, "time_mlp.linear_2.weight", {T4, T4}, &p));
  CKR(copy_f32(h, p, (size_t)T4 * T4, &m.t2_w));
  CKR(want(h, dm, "time_mlp.linear_2.bias", {T4}, &p));
  CKR(copy_f32(h, p, T4, &m.t2_b));
  m.stages.resize(4 + h->cfg.n_mid_blocks);
  int si = 0;
  CKR(load_stage(h, dm, "down_blocks.0", IC, m.stages[si++]));
  CKR(load_stage(h, dm, "down_blocks.1", C, m.stages[si++]));
  for (int i = 0; i < h->cfg.n_mid_blocks; ++i) CKR(load_stage(h, dm, "mid_blocks." + std::to_string(i), C, m.stages[si++]));
  CKR(load_stage(h, dm, "up_blocks.0", 2 * C, m.stages[si++]));
  CKR(load_stage(h, dm, "up_blocks.1", 2 * C, m.stages[si++]));
  CKR(load_conv(h, dm, "down_blocks.0.2.conv", C, C, 3, m.down_s2));
  CKR(load_conv(h, dm, "down_blocks.1.2", C, C, 3, m.down_tail));
  {  // ConvTranspose1d weight [Cin, Cout, 4]: out[2i] = W1 x[i] + W3 x[i-1];  out[2i+1] = W2 x[i] + W0 x[i+1]
    const float *w, *b;
    CKR(want(h, dm, "up_blocks.0.2.conv.weight", {C, C, 4}, &w));
    CKR(want(h, dm, "up_blocks.0.2.conv.bias", {C}, &b));
    CKR(alloc_gemm(h, m.up_even, C, C, 2));
    CKR(pack_into(h, m.up_even, w, 4, (long long)C * 4, 1, make_int4(3, 1, 0, 0), C, 0));
    CKR(copy_f32(h, b, C, &m.up_even.bias));
    CKR(alloc_gemm(h, m.up_odd, C, C, 2));
    CKR(pack_into(h, m.up_odd, w, 4, (long long)C * 4, 1, make_int4(2, 0, 0, 0), C, 0));
    CKR(copy_f32(h, b, C, &m.up_odd.bias));
  }
  CKR(load_conv(h, dm, "up_blocks.1.2", C, C, 3, m.up_tail));
  CKR(load_conv(h, dm, "final_block.block.0", C, C, 3, m.final_conv));
  CKR(load_norm(h, dm, "final_block.block.1", C, m.final_gn, m.final_conv.bias));
  CKR(load_conv(h, dm, "final_proj", F, C, 1, m.final_proj));
  for (auto& kv : dm.used)
    if (!kv.second) return fail(h, CFM_ERR_WEIGHTS, "unexpected parameter '%s'", kv.first.c_str());
  CK(cudaDeviceSynchronize());
  h->weights_loaded = true;
  return 0;
}

int cfm_plan(cfm_handle* h, const int32_t* lengths, int32_t batch, int32_t t_pad, const float* t_span, int32_t n_points,
             int32_t solver) {
  if (!h || !lengths || !t_span) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->weights_loaded) return fail(h, CFM_ERR_STATE, "cfm_plan before cfm_load_weights");
  if (batch <= 0 || batch >= (1 << 24)) return fail(h, CFM_ERR_INVALID, "batch out of range");
  if (t_pad <= 0 || (t_pad & 1)) return fail(h, CFM_ERR_INVALID, "t_pad must be positive and even (reference fix_len_compatibility), got %d", t_pad);
  if (n_points < 2) return fail(h, CFM_ERR_INVALID, "t_span needs at least 2 points");
  for (int b = 0; b < batch; ++b)
    if (lengths[b] < 1 || lengths[b] > t_pad) return fail(h, CFM_ERR_INVALID, "lengths[%d]=%d outside [1, t_pad=%d]", b, lengths[b], t_pad);
  if (solver < CFM_SOLVER_EULER || solver > CFM_SOLVER_RK4)
    return fail(h, CFM_ERR_INVALID, "unknown solver id %d (expected euler=0, midpoint=1, heun3=2, rk4=3)", solver);
  CK(cudaSetDevice(h->cfg.device));
  // ---- plan cache: a TTS server alternates between a handful of shapes; re-planning (tables, workspace, graph) for a shape
  // seen before is wasted work (reference server.py:93-119 decodes one request after the other through the same model).
  h->use_clock++;
  for (Plan* c : h->plans) {
    if (c->B == batch && c->T == t_pad && c->solver == solver && (int)c->t_span.size() == n_points &&
        memcmp(c->L.data(), lengths, sizeof(int) * batch) == 0 && memcmp(c->t_span.data(), t_span, sizeof(float) * n_points) == 0) {
      c->last_use = h->use_clock;
      h->plan = c;
      h->pdl_now = h->pdl < 0 ? (c->M1 <= 2048 ? 1 : 0) : h->pdl;
      return 0;
    }
  }
  while ((int)h->plans.size() >= std::max(1, h->plan_cache)) {  // evict the least recently used plan; its block is reused below
    size_t lru = 0;
    for (size_t i = 1; i < h->plans.size(); ++i)
      if (h->plans[i]->last_use < h->plans[lru]->last_use) lru = i;
    Plan* victim = h->plans[lru];
    h->plans.erase(h->plans.begin() + lru);
    // Nothing is freed here (no device synchronisation): the block goes to the pool, and whoever takes it next is ordered
    // behind the victim's last kernels through the busy event.  A block that does not fit the pool is released by cudaFree,
    // which waits for the device by itself.
    free_plan(h, victim, true);
  }
  Plan* pl = new Plan();
  pl->B = batch, pl->T = t_pad, pl->solver = solver, pl->last_use = h->use_clock;
  pl->L.assign(lengths, lengths + batch);
  pl->t_span.assign(t_span, t_span + n_points);
  struct Guard {  // a failed plan is neither cached nor current
    cfm_handle* h; Plan* pl; bool ok = false;
    ~Guard() { if (!ok) free_plan(h, pl, false); }
  } guard{h, pl};
  CKR(build_stages(h, pl, t_span, n_points, solver));
  cudaStream_t ps = h->own_stream;  // set-up stream: ordered behind earlier work of the handle, later calls wait for it
  CKR(enter_stream(h, ps));

  // ---- packed row tables (DESIGN.md "data layout"): a half-res segment is L2 valid rows + the halo / pad-token row +
  // one zero guard row (the halo row's k=3 conv reads row L2+1, which must not be the next utterance); full-res = 2x that.
  std::vector<UttTable> u1(batch), u2(batch);
  int s2 = 0;
  for (int b = 0; b < batch; ++b) {
    const int L = lengths[b], P = t_pad - L, L2 = (L + 1) / 2, T2 = t_pad / 2, P2 = T2 - L2;
    u2[b].start = s2, u2[b].len = L2, u2[b].rows = L2 + 2, u2[b].bias_rows = std::max(P2 - 1, 0), u2[b].t_res = T2;
    u2[b].pad_key_bias = P2 >= 1 ? logf((float)P2) - 1.f : -INFINITY;
    u2[b].inv_gn_count = 1.0 / ((double)(h->C() / 8) * T2);
    u1[b].start = 2 * s2, u1[b].len = L, u1[b].rows = 2 * (L2 + 2), u1[b].bias_rows = std::max(P - 1, 0), u1[b].t_res = t_pad;
    u1[b].pad_key_bias = P >= 1 ? logf((float)P) - 1.f : -INFINITY;
    u1[b].inv_gn_count = 1.0 / ((double)(h->C() / 8) * t_pad);
    s2 += L2 + 2;
  }
  pl->M2 = s2, pl->M1 = 2 * s2;
  h->pdl_now = h->pdl < 0 ? (pl->M1 <= 2048 ? 1 : 0) : h->pdl;
  std::vector<int> i1(pl->M1, 0), i2(pl->M2, 0);
  std::vector<int4> w1, w2;
  std::vector<int> wfirst1(batch + 1, 0), wfirst2(batch + 1, 0);
  const int QT = 128;
  for (int b = 0; b < batch; ++b) {
    wfirst1[b] = (int)w1.size(), wfirst2[b] = (int)w2.size();
    for (int r = 0; r < 2; ++r) {
      const UttTable& u = r == 0 ? u1[b] : u2[b];
      std::vector<int>& info = r == 0 ? i1 : i2;
      const int P = u.t_res - u.len;
      for (int t = 0; t < u.rows; ++t) {
        int v = b;
        if (t < u.len) v |= ROW_VALID | ROW_INSTAT;
        else if (t == u.len && P >= 1) v |= ROW_INSTAT;
        info[u.start + t] = v;
      }
      std::vector<int4>& w = r == 0 ? w1 : w2;
      for (int hd = 0; hd < h->cfg.n_heads; ++hd)
        for (int q0 = 0; q0 < u.len + 1; q0 += QT) w.push_back(make_int4(b, hd, q0, 0));
    }
  }
  pl->n_work1 = (int)w1.size(), pl->n_work2 = (int)w2.size();
  wfirst1[batch] = pl->n_work1, wfirst2[batch] = pl->n_work2;
  {  // ---- lanes: contiguous utterance groups of equal estimated cost  L (274 C^2 + 1800 C) + 24 C L^2  
    int n_lanes = std::max(1, std::min(h->lanes_req, batch));
    if (!h->bf && h->fp32_tc) n_lanes = 1;  // the split-operand scratch buffer is shared by every launch of the chain
    n_lanes = std::max(1, std::min(n_lanes, pl->M1 / std::max(1, h->lane_min_rows)));
    if (h->stop_after >= 0) n_lanes = 1;  // the debug launch counter is meaningful for a single chain only
    const double Cd = h->C();
    std::vector<double> cum(batch + 1, 0.0);
    for (int b = 0; b < batch; ++b) {
      const double L = lengths[b];
      cum[b + 1] = cum[b] + L * (274.0 * Cd * Cd + 1800.0 * Cd) + 24.0 * Cd * L * L;
    }
    pl->lanes.clear();
    int b0 = 0;
    for (int i = 0; i < n_lanes; ++i) {
      int b1 = batch;
      if (i + 1 < n_lanes) {
        const double target = cum[batch] * (i + 1) / n_lanes;
        b1 = b0 + 1;
        while (b1 < batch - (n_lanes - 1 - i) && fabs(cum[b1 + 1] - target) <= fabs(cum[b1] - target)) ++b1;
      }
      LaneDef ln;
      ln.b0 = b0, ln.nb = b1 - b0;
      ln.r2 = u2[b0].start;
      ln.m2 = (b1 < batch ? u2[b1].start : pl->M2) - ln.r2;
      ln.w0[0] = wfirst1[b0], ln.nw[0] = wfirst1[b1] - wfirst1[b0];
      ln.w0[1] = wfirst2[b0], ln.nw[1] = wfirst2[b1] - wfirst2[b0];
      pl->lanes.push_back(ln);
      b0 = b1;
    }
    // Attention CTAs cost ~ the utterance's key tiles: longest first within a lane, so that the last, partly filled wave of the grid
    // holds the shortest items (ragged batches: cfg3)
    for (const LaneDef& ln : pl->lanes) {
      std::stable_sort(w1.begin() + ln.w0[0], w1.begin() + ln.w0[0] + ln.nw[0], [&](const int4& a, const int4& b) { return u1[a.x].len > u1[b.x].len; });
      std::stable_sort(w2.begin() + ln.w0[1], w2.begin() + ln.w0[1] + ln.nw[1], [&](const int4& a, const int4& b) { return u2[a.x].len > u2[b.x].len; });
    }
    while ((int)h->lane_streams.size() + 1 < n_lanes) {
      cudaStream_t st;
      cudaEvent_t ev;
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      h->lane_streams.push_back(st), h->lane_events.push_back(ev);
    }
    if (!h->fork_event) CK(cudaEventCreateWithFlags(&h->fork_event, cudaEventDisableTiming));
  }
  const int n_t = (int)pl->stages.size(), n_res = 4 + h->cfg.n_mid_blocks;
  pl->n_trows = std::max(n_t, (int)batch) + 1;
  {  // size this plan's workspace block (tables + state + activations + host-path staging)
    const size_t C_ = h->C(), I_ = h->inner(), F_ = h->cfg.out_channels, es_ = h->es, M1_ = pl->M1, M2_ = pl->M2;
    const size_t nt = pl->n_trows, nres = n_res;
    size_t need = (size_t)batch * 2 * sizeof(UttTable) + (M1_ + M2_) * 4 + (w1.size() + w2.size()) * sizeof(int4);
    need += nt * (4 + (size_t)h->cfg.in_channels * 4 + 2 * 4 * C_ * 4 + nres * C_ * 4);
    need += 5 * M1_ * F_ * 4 + (2 * nres + 1) * (size_t)batch * (16 * 8 + 8 * 8);
    need += M1_ * (size_t)roundup(h->cfg.in_channels, 64) * es_;
    need += (M1_ + M2_) * (3 * C_ * 4 + (5 * C_ + 4 * I_ + 4 * C_) * es_);
    need += 3 * (size_t)batch * F_ * t_pad * 4 + (size_t)batch * std::max(h->cfg.in_channels - 2 * h->cfg.out_channels, 1) * 4;
    need += 64 * 1024 + (1 << 20);  // per-allocation 1 KB rounding, slack
    CKR(acquire_workspace(h, pl, need, ps));
  }
  auto up = [&](auto** dst, const auto& vec) -> int {
    using E = typename std::remove_reference<decltype(vec[0])>::type;
    CKR(plan_alloc(h, pl, reinterpret_cast<void**>(dst), vec.size() * sizeof(E)));
    // pageable source: the runtime stages the bytes before returning, so the vectors may go out of scope
    CK(cudaMemcpyAsync(*dst, vec.data(), vec.size() * sizeof(E), cudaMemcpyHostToDevice, ps));
    return 0;
  };
  CKR(up(&pl->utt1, u1));
  CKR(up(&pl->utt2, u2));
  CKR(up(&pl->info1, i1));
  CKR(up(&pl->info2, i2));
  CKR(up(&pl->work1, w1));
  CKR(up(&pl->work2, w2));

  // ---- workspace
  const int C = h->C(), I = h->inner(), F = h->cfg.out_channels, es = h->es, T4 = 4 * C;
  std::vector<float> tv(pl->n_trows, 0.f);
  for (int j = 0; j < n_t; ++j) tv[j] = pl->stages[j].t;
  CKR(up(&pl->tvals, tv));
  CKR(plan_alloc_t(h, pl, &pl->sinemb, (size_t)pl->n_trows * h->cfg.in_channels));
  CKR(plan_alloc_t(h, pl, &pl->temb_a, (size_t)pl->n_trows * T4));
  CKR(plan_alloc_t(h, pl, &pl->temb, (size_t)pl->n_trows * T4));
  CKR(plan_alloc_t(h, pl, &pl->tproj, (size_t)pl->n_trows * n_res * C));
  CKR(plan_alloc_t(h, pl, &pl->xstate, (size_t)pl->M1 * F));
  CKR(plan_alloc_t(h, pl, &pl->vout, (size_t)pl->M1 * F));
  int n_k = 0;
  for (auto& st : pl->stages) n_k = std::max(n_k, st.kout + 1);
  for (int j = 0; j < n_k; ++j) CKR(plan_alloc_t(h, pl, &pl->kbuf[j], (size_t)pl->M1 * F));
  const int n_sites = 2 * n_res + 1;
  pl->stats_bytes = (size_t)n_sites * batch * 16 * sizeof(double);
  CKR(plan_alloc(h, pl, reinterpret_cast<void**>(&pl->stats), pl->stats_bytes));
  pl->xin_ld = roundup(h->cfg.in_channels, 64);
  CKR(plan_alloc(h, pl, &pl->xin, (size_t)pl->M1 * pl->xin_ld * es));
  for (int r = 0; r < 2; ++r) {
    const size_t M = r == 0 ? pl->M1 : pl->M2;
    CKR(plan_alloc_t(h, pl, &pl->hraw[r], M * C));
    CKR(plan_alloc_t(h, pl, &pl->rres[r], M * C));
    CKR(plan_alloc_t(h, pl, &pl->X[r], M * C));
    CKR(plan_alloc(h, pl, &pl->hact[r], M * C * es));
    CKR(plan_alloc(h, pl, &pl->Xn[r], M * C * es));
    CKR(plan_alloc(h, pl, &pl->qkv[r], M * 3 * I * es));
    CKR(plan_alloc(h, pl, &pl->ao[r], M * I * es));
    CKR(plan_alloc(h, pl, &pl->ffh[r], M * 4 * C * es));
    CKR(plan_alloc(h, pl, &pl->sin_[r], M * C * es));
    CKR(plan_alloc(h, pl, &pl->cat[r], M * 2 * C * es));
  }
  pl->split3 = nullptr;
  if (!h->bf && h->fp32_tc)  // [hi | lo] bf16 copy of the widest GEMM operand (ffh, [M1, 4C]): tensor-core fp32 mode
    CKR(plan_alloc(h, pl, &pl->split3, (size_t)pl->M1 * 8 * C * sizeof(bf16)));
  CKR(leave_stream(h, ps));  // tables and the cleared workspace are ready before any later call touches them

  // ---- the whole ODE loop as one CUDA graph: captured here, or lazily before the plan's (graph_after + 1)-th decode.
  // Capturing and instantiating ~1200 nodes costs more than a B = 1 decode itself, and a server sees a new length with
  // almost every request, so a plan's first decode uses direct launches and the graph is built only when the plan is reused.
  h->launch_counter = 0;
  pl->solves = 0;
  h->plan = pl;
  h->plans.push_back(pl);
  guard.ok = true;
  if (h->graph_after == 0) CKR(capture_graph(h, pl));
  return 0;
}

int cfm_solve(cfm_handle* h, const float* mu, const float* z, float* out, void* stream) {
  if (!h || !mu || !z || !out) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->plan) return fail(h, CFM_ERR_STATE, "cfm_solve before cfm_plan");
  Plan* pl = h->plan;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(h->cfg.device));
  h->launch_counter = 0;
  if (!pl->exec && pl->solves >= h->graph_after) CKR(capture_graph(h, pl));
  pl->solves++;
  CKR(enter_stream(h, s));
  CKR(emit_pack(h, pl, z, mu, s));
  if (pl->exec) {
    CK(cudaGraphLaunch(pl->exec, s));
  } else {
    CKR(emit_ode_loop(h, pl, s));
    pl->launches_per_solve = h->launch_counter + 1;
  }
  CKR(emit_unpack(h, pl, pl->xstate, z, out, s));
  CKR(leave_stream(h, s));
  return 0;
}

static int solve_host_impl(cfm_handle* h, const float* mu, const float* z, const float* spks, float* out) {
  if (!h || !mu || !z || !out) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->plan) return fail(h, CFM_ERR_STATE, "cfm_solve_host before cfm_plan");
  Plan* pl = h->plan;
  CK(cudaSetDevice(h->cfg.device));
  const int S = h->cfg.in_channels - 2 * h->cfg.out_channels;
  if ((S > 0) != (spks != nullptr))
    return fail(h, CFM_ERR_INVALID, "this estimator has %d speaker channels: %s", S,
                S > 0 ? "use cfm_solve_host_spks with a host (batch, S) matrix" : "spks must be NULL");
  const size_t n = (size_t)pl->B * h->cfg.out_channels * pl->T;
  cudaStream_t s = h->own_stream;
  CKR(enter_stream(h, s));
  if (!pl->stage_mu) {  // device staging buffers live with the plan: no allocation on the per-call path
    CKR(plan_alloc_t(h, pl, &pl->stage_mu, n));
    CKR(plan_alloc_t(h, pl, &pl->stage_z, n));
    CKR(plan_alloc_t(h, pl, &pl->stage_out, n));
    if (S > 0) CKR(plan_alloc_t(h, pl, &pl->stage_spk, (size_t)pl->B * S));
  }
  CK(cudaMemcpyAsync(pl->stage_mu, mu, n * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(pl->stage_z, z, n * 4, cudaMemcpyHostToDevice, s));
  if (S > 0) CK(cudaMemcpyAsync(pl->stage_spk, spks, (size_t)pl->B * S * 4, cudaMemcpyHostToDevice, s));
  h->spks = S > 0 ? pl->stage_spk : nullptr;  // never a pointer left over from an earlier device-side call
  CKR(cfm_solve(h, pl->stage_mu, pl->stage_z, pl->stage_out, s));
  h->spks = nullptr;
  CK(cudaMemcpyAsync(out, pl->stage_out, n * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

int cfm_solve_host(cfm_handle* h, const float* mu, const float* z, float* out) { return solve_host_impl(h, mu, z, nullptr, out); }

// Sharded host entry: this handle decodes the utterances index[0 .. B) of a HOST batch whose tensors are
// (n_total, F, t_pad): per-utterance H2D from their positions in the caller's buffers, the decode, per-utterance D2H of the
// results to the same positions of `out`.  Everything is enqueued on the handle's own stream and the call returns without
// waiting (cfm_synchronize), so a host thread can start every GPU of the box before it waits for any of them.
int cfm_solve_host_indexed(cfm_handle* h, const float* mu, const float* z, const float* spks, float* out, const int32_t* index,
                           int32_t n_total) {
  if (!h || !mu || !z || !out || !index) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->plan) return fail(h, CFM_ERR_STATE, "cfm_solve_host_indexed before cfm_plan");
  Plan* pl = h->plan;
  CK(cudaSetDevice(h->cfg.device));
  const int S = h->cfg.in_channels - 2 * h->cfg.out_channels;
  if ((S > 0) != (spks != nullptr)) return fail(h, CFM_ERR_INVALID, "this estimator has %d speaker channels: spks must%s be given", S, S > 0 ? "" : " not");
  for (int b = 0; b < pl->B; ++b)
    if (index[b] < 0 || index[b] >= n_total) return fail(h, CFM_ERR_INVALID, "index[%d]=%d outside [0, %d)", b, index[b], n_total);
  const size_t per = (size_t)h->cfg.out_channels * pl->T, n = (size_t)pl->B * per;
  cudaStream_t s = h->own_stream;
  CKR(enter_stream(h, s));
  if (!pl->stage_mu) {
    CKR(plan_alloc_t(h, pl, &pl->stage_mu, n));
    CKR(plan_alloc_t(h, pl, &pl->stage_z, n));
    CKR(plan_alloc_t(h, pl, &pl->stage_out, n));
    if (S > 0) CKR(plan_alloc_t(h, pl, &pl->stage_spk, (size_t)pl->B * S));
  }
  for (int b = 0; b < pl->B; ++b) {
    CK(cudaMemcpyAsync(pl->stage_mu + b * per, mu + (size_t)index[b] * per, per * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(pl->stage_z + b * per, z + (size_t)index[b] * per, per * 4, cudaMemcpyHostToDevice, s));
    if (S > 0) CK(cudaMemcpyAsync(pl->stage_spk + (size_t)b * S, spks + (size_t)index[b] * S, (size_t)S * 4, cudaMemcpyHostToDevice, s));
  }
  h->spks = S > 0 ? pl->stage_spk : nullptr;
  CKR(cfm_solve(h, pl->stage_mu, pl->stage_z, pl->stage_out, s));
  h->spks = nullptr;
  for (int b = 0; b < pl->B; ++b)
    CK(cudaMemcpyAsync(out + (size_t)index[b] * per, pl->stage_out + b * per, per * 4, cudaMemcpyDeviceToHost, s));
  CKR(leave_stream(h, s));
  return 0;
}

// ---- front / back of the decode (reference matcha/inference.py:146-172); kernels in kernels.cuh
int cfm_front_durations(cfm_handle* h, const float* durations, int32_t batch, int32_t t_x, int32_t* cum, int32_t* fine_lengths, void* stream) {
  if (!h || !durations || !cum || !fine_lengths || batch <= 0 || t_x <= 0) return fail(h, CFM_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(h->cfg.device));
  front_cumsum_kernel<<<batch, 32, 0, static_cast<cudaStream_t>(stream)>>>(durations, batch, t_x, cum, fine_lengths);
  CK(cudaGetLastError());
  return 0;
}

int cfm_front_expand(cfm_handle* h, const float* mu_x, const int32_t* cum, const int32_t* fine_lengths, int32_t batch, int32_t t_x,
                     int32_t t_pad, float* mu_y, float* y_mask, void* stream) {
  if (!h || !mu_x || !cum || !fine_lengths || !mu_y || batch <= 0 || t_x <= 0 || t_pad <= 0) return fail(h, CFM_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(h->cfg.device));
  dim3 grid((t_pad + 127) / 128, batch);
  front_expand_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(mu_x, cum, fine_lengths, batch, h->cfg.out_channels, t_x, t_pad, mu_y, y_mask);
  CK(cudaGetLastError());
  return 0;
}

int cfm_denormalize(cfm_handle* h, const float* x, int32_t batch, int32_t t_pad, int32_t t_out, float mean, float stdv, float* out, void* stream) {
  if (!h || !x || !out || batch <= 0 || t_pad <= 0 || t_out <= 0 || t_out > t_pad) return fail(h, CFM_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(h->cfg.device));
  const long long rows = (long long)batch * h->cfg.out_channels, n = rows * t_out;
  back_denormalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, t_pad, t_out, mean, stdv, out);
  CK(cudaGetLastError());
  return 0;
}

// Waits for everything this handle has enqueued (on any stream).
int cfm_synchronize(cfm_handle* h) {
  if (!h) return CFM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  if (h->busy_valid) CK(cudaEventSynchronize(h->busy_event));
  return 0;
}

int cfm_solve_host_spks(cfm_handle* h, const float* mu, const float* z, const float* spks, float* out) {
  if (!spks) return fail(h, CFM_ERR_INVALID, "null argument");
  return solve_host_impl(h, mu, z, spks, out);
}

int cfm_estimator_t(cfm_handle* h, const float* x, const float* mu, const float* t_host, int32_t n_t, float* v, void* stream) {
  if (!h || !x || !mu || !v || !t_host) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->plan) return fail(h, CFM_ERR_STATE, "cfm_estimator before cfm_plan");
  Plan* pl = h->plan;
  if (n_t != 1 && n_t != pl->B) return fail(h, CFM_ERR_INVALID, "n_t must be 1 (one time for the batch) or the planned batch %d, got %d", pl->B, n_t);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(h->cfg.device));
  CKR(enter_stream(h, s));
  CKR(emit_pack(h, pl, x, mu, s));
  // The time values go to the rows behind the planned grid (which stays intact for later solves).  Pageable source: staged by
  // the runtime before the call returns, so the caller's array may be a temporary and no stream synchronisation is needed.
  const int n_grid = (int)pl->stages.size();
  float* t_dev = pl->tvals + (n_t == 1 ? pl->n_trows - 1 : 0);
  std::vector<float> keep;
  if (n_t != 1) {  // B time points overwrite the planned grid's slots: restore them afterwards
    keep.resize(pl->n_trows, 0.f);
    for (int j = 0; j < n_grid; ++j) keep[j] = pl->stages[j].t;
  }
  CK(cudaMemcpyAsync(t_dev, t_host, sizeof(float) * n_t, cudaMemcpyHostToDevice, s));
  CKR(emit_time_embedding(h, pl, n_t, t_dev, s));
  OdeStage st;
  st.t = t_host[0], st.c_v = 1.f, st.kout = -1, st.write_state = false;
  for (int j = 0; j < 3; ++j) st.c_k[j] = 0.f, st.kin[j] = -1;
  h->launch_counter = 0;
  const bool per_utt = n_t != 1;
  CKR(for_each_lane(h, pl, s, [&](const LaneDef& ln, cudaStream_t ls) -> int { return emit_nfe(h, pl, ln, 0, st, pl->vout, ls, per_utt); }));
  CKR(emit_unpack(h, pl, pl->vout, nullptr, v, s));
  if (n_t != 1) CK(cudaMemcpyAsync(pl->tvals, keep.data(), sizeof(float) * pl->n_trows, cudaMemcpyHostToDevice, s));
  CKR(leave_stream(h, s));
  return 0;
}

int cfm_estimator(cfm_handle* h, const float* x, const float* mu, float t, float* v, void* stream) {
  return cfm_estimator_t(h, x, mu, &t, 1, v, stream);
}

// Debug / measurement: one decode with direct launches and a CUDA event before every launch; writes "tag,M,N,K,flops,us" lines
// (one per launch, in order; us = time to the next mark) into buf.  In-situ (warm-cache, back-to-back) per-kernel times: the
// ncu launch lists in profiles/ are cold-cache and serialised.
int cfm_debug_timeline(cfm_handle* h, const float* mu, const float* z, float* out, char* buf, int64_t cap, void* stream) {
  if (!h || !mu || !z || !out || !buf || cap < 64) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->plan) return fail(h, CFM_ERR_STATE, "cfm_debug_timeline before cfm_plan");
  Plan* pl = h->plan;
  if (pl->lanes.size() != 1) return fail(h, CFM_ERR_STATE, "cfm_debug_timeline needs a single-lane plan");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(h->cfg.device));
  CKR(enter_stream(h, s));
  CKR(emit_pack(h, pl, z, mu, s));
  h->tl_on = true, h->tl_n = 0;
  int r = emit_ode_loop(h, pl, s);
  if (r == 0) r = tl_mark(h, s, "end", 0, 0, 0, 0.0);
  h->tl_on = false;
  CKR(r);
  CKR(emit_unpack(h, pl, pl->xstate, z, out, s));
  CKR(leave_stream(h, s));
  CK(cudaStreamSynchronize(s));
  int64_t off = 0;
  for (size_t i = 0; i + 1 < h->tl_n; ++i) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->tl[i].ev, h->tl[i + 1].ev));
    const int w = snprintf(buf + off, (size_t)(cap - off), "%s,%d,%d,%d,%.6g,%.3f\n", h->tl[i].tag, h->tl[i].M, h->tl[i].N, h->tl[i].K,
                           h->tl[i].flops, ms * 1e3);
    if (w < 0 || off + w >= cap) return fail(h, CFM_ERR_INVALID, "timeline buffer too small (%lld bytes)", (long long)cap);
    off += w;
  }
  buf[off] = 0;
  return 0;
}

int cfm_plan_info(const cfm_handle* h, int64_t* rows_full, int64_t* rows_half, int64_t* n_nfe, int64_t* kernels_per_solve,
                  int64_t* workspace_bytes) {
  if (!h || !h->plan) return CFM_ERR_STATE;
  if (rows_full) *rows_full = h->plan->M1;
  if (rows_half) *rows_half = h->plan->M2;
  if (n_nfe) *n_nfe = (int64_t)h->plan->stages.size();
  if (kernels_per_solve) *kernels_per_solve = h->plan->launches_per_solve;
  if (workspace_bytes) *workspace_bytes = (int64_t)h->plan->bytes;
  return 0;
}

int cfm_debug_read(cfm_handle* h, const char* name, float* host_dst, int64_t max_elems, int64_t* rows, int64_t* cols) {
  if (!h || !h->plan || !name) return fail(h, CFM_ERR_STATE, "cfm_debug_read needs a plan");
  Plan* pl = h->plan;
  const int C = h->C(), I = h->inner(), F = h->cfg.out_channels;
  struct Ent { const char* n; const void* p; long long r, c; bool act; };
  const Ent table[] = {
      {"xin", pl->xin, pl->M1, pl->xin_ld, true},   {"xstate", pl->xstate, pl->M1, F, false},
      {"vout", pl->vout, pl->M1, F, false},         {"hraw1", pl->hraw[0], pl->M1, C, mid16(h)},
      {"hraw2", pl->hraw[1], pl->M2, C, mid16(h)},  {"rres1", pl->rres[0], pl->M1, C, mid16(h)},
      {"X1", pl->X[0], pl->M1, C, false},           {"X2", pl->X[1], pl->M2, C, false},
      {"hact1", pl->hact[0], pl->M1, C, true},      {"hact2", pl->hact[1], pl->M2, C, true},
      {"Xn1", pl->Xn[0], pl->M1, C, true},          {"qkv1", pl->qkv[0], pl->M1, 3 * I, true},
      {"qkv2", pl->qkv[1], pl->M2, 3 * I, true},    {"ao1", pl->ao[0], pl->M1, I, true},
      {"ao2", pl->ao[1], pl->M2, I, true},          {"ffh1", pl->ffh[0], pl->M1, 4 * C, true},
      {"sin1", pl->sin_[0], pl->M1, C, true},       {"sin2", pl->sin_[1], pl->M2, C, true},
      {"cat1", pl->cat[0], pl->M1, 2 * C, true},    {"cat2", pl->cat[1], pl->M2, 2 * C, true},
      {"sinemb", pl->sinemb, (long long)pl->stages.size(), h->cfg.in_channels, false},
      {"temb", pl->temb, (long long)pl->stages.size(), 4 * C, false},
      {"tproj", pl->tproj, (long long)pl->stages.size(), (4 + h->cfg.n_mid_blocks) * C, false},
  };
  for (const Ent& e : table) {
    if (strcmp(e.n, name) != 0) continue;
    const long long n = e.r * e.c;
    if (rows) *rows = e.r;
    if (cols) *cols = e.c;
    if (!host_dst) return 0;
    if (n > max_elems) return fail(h, CFM_ERR_INVALID, "buffer '%s' needs %lld elements", name, n);
    CK(cudaDeviceSynchronize());
    if (!e.act || !h->bf) {
      CK(cudaMemcpy(host_dst, e.p, n * 4, cudaMemcpyDeviceToHost));
    } else {
      std::vector<uint16_t> tmp(n);
      CK(cudaMemcpy(tmp.data(), e.p, n * 2, cudaMemcpyDeviceToHost));
      for (long long i = 0; i < n; ++i) {
        uint32_t u = (uint32_t)tmp[i] << 16;
        memcpy(&host_dst[i], &u, 4);
      }
    }
    return 0;
  }
  if (strcmp(name, "stats") == 0) {  // [sites][B][8][2] doubles, returned as fp32
    const long long n = (long long)(pl->stats_bytes / sizeof(double));
    if (rows) *rows = n / 16;
    if (cols) *cols = 16;
    if (!host_dst) return 0;
    if (n > max_elems) return fail(h, CFM_ERR_INVALID, "buffer 'stats' needs %lld elements", n);
    CK(cudaDeviceSynchronize());
    std::vector<double> tmp(n);
    CK(cudaMemcpy(tmp.data(), pl->stats, n * sizeof(double), cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n; ++i) host_dst[i] = (float)tmp[i];
    return 0;
  }
  return fail(h, CFM_ERR_INVALID, "unknown debug buffer '%s'", name);
}

int cfm_set_speakers(cfm_handle* h, const float* spks_dev) {
  if (!h) return CFM_ERR_INVALID;
  h->spks = spks_dev;
  return 0;
}

int cfm_set_lanes(cfm_handle* h, int32_t lanes, int32_t min_rows) {
  if (!h || lanes < 1 || lanes > 16) return fail(h, CFM_ERR_INVALID, "lanes must be in [1, 16]");
  h->lanes_req = lanes;
  if (min_rows > 0) h->lane_min_rows = min_rows;
  free_all_plans(h);  // cached plans were cut with the old lane count
  return 0;
}

int cfm_set_option(cfm_handle* h, const char* key, int32_t value) {
  if (!h || !key) return fail(h, CFM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  if (strcmp(key, "plan_cache") == 0 && value >= 1) {
    h->plan_cache = value;
    return 0;
  }
  free_all_plans(h);  // kernel selection is baked into the cached plans' graphs
  if (strcmp(key, "tma_epi") == 0) h->tma_epi = value == 1 ? 0x3f : value;
  else if (strcmp(key, "pair_mode") == 0 && value >= 0 && value <= 2) h->pair_mode = value;
  else if (strcmp(key, "small_tiles") == 0 && value >= 0) h->small_tiles = value;
  else if (strcmp(key, "graph_after") == 0 && value >= 0) h->graph_after = value;
  else if (strcmp(key, "snake_warps") == 0 && (value == 8 || value == 12)) h->snake_warps = value;
  else if (strcmp(key, "l2_persist_mb") == 0 && value >= 0) return apply_l2_persist(h, value);
  else if (strcmp(key, "pair_n256") == 0) h->pair_n256 = value != 0;
  else if (strcmp(key, "ff_fused") == 0) h->ff_fused = value != 0;
  else if (strcmp(key, "rowln") == 0 && value >= 0 && value <= 2) h->rowln = value;
  else if (strcmp(key, "res_side") == 0) h->res_side = value < 0 ? -1 : value != 0;
  else if (strcmp(key, "attn_persist") == 0) h->attn_persist = value != 0;
  else if (strcmp(key, "fp32_tc") == 0) h->fp32_tc = value != 0;
  else if (strcmp(key, "rowln_ff2") == 0) h->rowln_ff2 = value != 0;
  else if (strcmp(key, "bf16_mid") == 0) h->bf16_mid = value != 0;
  else if (strcmp(key, "bn_full") == 0) h->bn_full = value;
  else if (strcmp(key, "bn_half") == 0) h->bn_half = value;
  else if (strcmp(key, "pair_min_k") == 0 && value >= 0) h->pair_min_k = value;
  else if (strcmp(key, "direct_epi") == 0 && value >= 0) h->direct_epi = value;
  else if (strcmp(key, "pdl") == 0) h->pdl = value < 0 ? -1 : value != 0;
  else if (strcmp(key, "cluster") == 0 && (value == 1 || value == 2 || value == 4)) h->cluster = value;
  else return fail(h, CFM_ERR_INVALID, "unknown option '%s' or value %d out of range", key, (int)value);
  return 0;
}

int cfm_debug_attn_profile(cfm_handle* h, unsigned long long* prof_dev) {
  if (!h) return CFM_ERR_INVALID;
  h->attn_prof = prof_dev;
  return 0;
}

int cfm_debug_ff_profile(cfm_handle* h, unsigned long long* prof_dev) {
  if (!h) return CFM_ERR_INVALID;
  h->ff_prof = prof_dev;
  return 0;
}

int cfm_debug_rowln_profile(cfm_handle* h, unsigned long long* prof_dev) {
  if (!h) return CFM_ERR_INVALID;
  h->rowln_prof = prof_dev;
  return 0;
}

int cfm_debug_stop_after(cfm_handle* h, int64_t n_launches) {
  if (!h) return CFM_ERR_INVALID;
  h->stop_after = n_launches;
  return 0;
}

// Debug: tensor-core GEMM in a chosen epilogue mode with the CTA-0 role profile.  mode: 0 = bf16 store, 1 = fp32 store,
// 2 = fp32 residual add in place (d_f32 is both residual and output).  prof (device, >= 16 u64) receives cycle counters:
// [0] producer total [1] producer waiting for free stages [2] MMA total [3] MMA waiting for operands [4] MMA waiting for a
// free accumulator [5] epilogue warp total [6] epilogue waiting for an accumulator [8] tiles done by CTA 0.
int cfm_debug_gemm_profile(cfm_handle* h, const void* a, const void* w, float* d_f32, void* d_bf16, int32_t M, int32_t N,
                           int32_t K, int32_t n_taps, const int32_t* shifts, int32_t mode, unsigned long long* prof,
                           void* stream) {
  if (!h || !a || !w) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->bf) return fail(h, CFM_ERR_INVALID, "needs a bf16 handle");
  CK(cudaSetDevice(h->cfg.device));
  GemmW gw;
  gw.w = const_cast<void*>(w), gw.N = N, gw.K = K, gw.Kp = K, gw.n_taps = n_taps, gw.n_stride = N;
  GemmParams p = gemm_base(M, a, K, M, gw, shifts, nullptr);
  p.prof = prof;
  if (mode == 0) p.mode = EPI_STORE, p.out_act = d_bf16, p.ld_act = N;
  else if (mode == 1) p.mode = EPI_STATS, p.fused_stats = 0, p.out_f32 = d_f32, p.ld_f32 = N;
  else if (mode == 3) {  // fp32 store + fused GroupNorm statistics on synthetic 942-row utterances
    static int* info = nullptr;
    static double* stats = nullptr;
    static float* bias = nullptr;
    static int info_rows = 0;
    if (info_rows < M) {
      std::vector<int> hinfo(M);
      for (int i = 0; i < M; ++i) hinfo[i] = (i / 942) | ROW_VALID | ROW_INSTAT;
      CK(cudaMalloc(&info, (size_t)M * 4));
      CK(cudaMemcpy(info, hinfo.data(), (size_t)M * 4, cudaMemcpyHostToDevice));
      CK(cudaMalloc(&stats, (size_t)(M / 942 + 1) * 16 * 8 * 64));
      CK(cudaMemset(stats, 0, (size_t)(M / 942 + 1) * 16 * 8 * 64));
      CK(cudaMalloc(&bias, 4096 * 4));
      CK(cudaMemset(bias, 0, 4096 * 4));
      info_rows = M;
    }
    p.mode = EPI_STATS, p.fused_stats = 1, p.out_f32 = d_f32, p.ld_f32 = N, p.row_info = info, p.stats = stats;
    p.group_ch = N / 8, p.bias = bias;
  } else p.mode = EPI_RESID, p.resid = d_f32, p.ld_resid = N, p.out_f32 = d_f32, p.ld_f32 = N;
  const int saved = h->cfg.flags;
  h->cfg.flags &= ~CFM_FLAG_SIMT_GEMM;
  int r = launch_gemm(h, p, true, static_cast<cudaStream_t>(stream));
  h->cfg.flags = saved;
  return r;
}

int cfm_debug_gemm(cfm_handle* h, const void* a, const void* w, float* d, int32_t M, int32_t N, int32_t K, int32_t n_taps,
                   const int32_t* shifts, int32_t use_tc, void* stream) {
  if (!h || !a || !w || !d) return fail(h, CFM_ERR_INVALID, "null argument");
  if (!h->bf) return fail(h, CFM_ERR_INVALID, "cfm_debug_gemm needs a bf16 handle");
  if (n_taps < 1 || n_taps > MAX_TAPS) return fail(h, CFM_ERR_INVALID, "n_taps out of range");
  CK(cudaSetDevice(h->cfg.device));
  GemmW gw;
  gw.w = const_cast<void*>(w), gw.N = N, gw.K = K, gw.Kp = K, gw.n_taps = n_taps, gw.n_stride = N;
  GemmParams p = gemm_base(M, a, K, M, gw, shifts, nullptr);
  p.mode = EPI_STATS, p.fused_stats = 0, p.out_f32 = d, p.ld_f32 = N;
  const int saved = h->cfg.flags;
  if (!use_tc) h->cfg.flags |= CFM_FLAG_SIMT_GEMM; else h->cfg.flags &= ~CFM_FLAG_SIMT_GEMM;
  int r = launch_gemm(h, p, true, static_cast<cudaStream_t>(stream));
  h->cfg.flags = saved;
  return r;
}

}  // extern "C"
