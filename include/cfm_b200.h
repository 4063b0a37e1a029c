/*
 * cfm_b200.h - C ABI of the B200-native CFM decode library (libcfm_b200.so).
 *
 * One hot path of faltiska/Matcha-TTS-24k: the conditional-flow-matching decode, i.e. the fixed-grid ODE
 * loop around the 1-D U-Net estimator.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference repository).  Plain C types only: device pointers are raw CUDA device
 * addresses, `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success or a
 * negative cfm_status; cfm_last_error() returns the message.  No exception or allocation crosses the ABI.
 * Nothing here falls back to the CPU: without a CUDA device of compute capability 10.x every call fails.
 */
#ifndef CFM_B200_H_
#define CFM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfm_handle cfm_handle;

typedef enum {
  CFM_OK = 0,
  CFM_ERR_INVALID = -1,   /* bad argument / unsupported configuration (host wrapper raises ValueError) */
  CFM_ERR_CUDA = -2,      /* CUDA runtime / driver failure */
  CFM_ERR_STATE = -3,     /* call order violated (e.g. solve before load_weights / plan) */
  CFM_ERR_WEIGHTS = -4    /* missing, unexpected or mis-shaped parameter */
} cfm_status;

/* BF16: bf16 operands on the tensor pipe, fp32 accumulation / residual stream / statistics / ODE state (the speed mode).
 * FP32: fp32 storage and fp32-FMA kernels everywhere (the parity reference, ~1e-6 of the oracle).
 * FP32_TC: fp32 storage, GEMMs and attention on the bf16 tensor pipe with bf16 x 3 split operands (a = hi + lo; hi*hi + lo*hi +
 *          hi*lo, fp32 accumulation): ~1-2e-5 relative L2 of the oracle, ~9x faster than FP32 (cfm_set_option "fp32_tc" toggles
 *          it on an FP32 handle). */
typedef enum { CFM_PREC_BF16 = 0, CFM_PREC_FP32 = 1, CFM_PREC_FP32_TC = 2 } cfm_precision;
typedef enum { CFM_SOLVER_EULER = 0, CFM_SOLVER_MIDPOINT = 1, CFM_SOLVER_HEUN3 = 2, CFM_SOLVER_RK4 = 3 } cfm_solver;

/* Mirrors the keyword arguments of the reference estimator constructor
 * Decoder(in_channels, out_channels, channels, dropout, attention_head_dim, n_blocks, num_mid_blocks, num_heads)
 * (matcha/models/components/decoder.py:203-216), as reached through
 * CFM(in_channels, out_channel, cfm_params, decoder_params) (matcha/models/components/flow_matching.py:110-117). */
typedef struct {
  int32_t in_channels;   /* 2 * n_feats */
  int32_t out_channels;  /* n_feats */
  int32_t channels;      /* C: width of both U-Net levels */
  int32_t n_heads;
  int32_t head_dim;
  int32_t n_blocks;      /* transformer blocks per stage */
  int32_t n_mid_blocks;
  int32_t precision;     /* cfm_precision: storage/MMA operand type; statistics, residual stream and ODE state are fp32 */
  int32_t device;        /* CUDA device ordinal */
  int32_t flags;         /* CFM_FLAG_* */
} cfm_config;

#define CFM_FLAG_NO_GRAPH 1      /* launch kernels directly instead of replaying the captured CUDA graph */
#define CFM_FLAG_SIMT_GEMM 2     /* debug: run every GEMM on the fp32-FMA kernel instead of tcgen05 */
#define CFM_FLAG_UNFUSED_STATS 4 /* debug: GroupNorm statistics in a separate pass instead of the GEMM epilogue */
#define CFM_FLAG_SIMT_ATTN 8     /* debug: fp32-FMA attention kernel instead of the tensor-core kernel */

/* One parameter of the estimator, named exactly as in the reference state_dict below `decoder.estimator.`
 * (e.g. "down_blocks.0.0.block1.block.0.weight"); data = device pointer to contiguous fp32. */
typedef struct {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} cfm_weight_desc;

/* Replaces: CFM.__init__ -> Decoder.__init__ (flow_matching.py:110-117, decoder.py:202-310). */
int cfm_create(const cfm_config* cfg, cfm_handle** out);
void cfm_destroy(cfm_handle* h);
const char* cfm_last_error(const cfm_handle* h); /* h may be NULL: error of the last failed cfm_create */

/* Replaces: nn.Module.load_state_dict on decoder.estimator.* (matcha/inference.py:194).  Packs / casts every
 * parameter into the library's own device layouts; must be called again whenever the parameters change.
 * Fails with CFM_ERR_WEIGHTS if any parameter is missing, unexpected or mis-shaped (the reference loads with
 * strict=False and would hide that). */
int cfm_load_weights(cfm_handle* h, const cfm_weight_desc* descs, int32_t n);

/* Prepares a decode of `batch` utterances padded to `t_pad` frames (t_pad even, reference
 * matcha/utils/model.py:15-21) with valid lengths `lengths[b]` (prefix masks, reference
 * matcha/utils/model.py:7-9) over the time grid `t_span[0..n_points)` with a torchdiffeq fixed-grid solver
 * (call site flow_matching.py:60-63).  Builds the packed row tables and the workspace; the CUDA graph of the
 * whole ODE loop is captured when the plan is reused (option "graph_after").  Plans are cached per handle by
 * (lengths, t_pad, t_span, solver): planning a shape seen before only makes that plan current.  No device synchronisation:
 * set-up work runs on the handle's own stream and later calls are ordered behind it.  Replaces the per-call setup of
 * BASECFM.solve + OdeSolverWrapper. */
int cfm_plan(cfm_handle* h, const int32_t* lengths, int32_t batch, int32_t t_pad, const float* t_span, int32_t n_points,
             int32_t solver);

/* Replaces: BASECFM.solve(x, t_span, mu, mask) (flow_matching.py:60-63) for the planned shapes.
 * mu, z, out: device fp32, contiguous (batch, out_channels, t_pad).  `z` is the initial state (the injected
 * noise); padded frames of `out` equal `z` there, exactly as in the reference where the masked velocity never
 * moves them.  Asynchronous on `stream`.  A handle owns one workspace per plan: calls issued on different streams are
 * ordered one after the other by the library (an event recorded after each enqueue), never run concurrently. */
int cfm_solve(cfm_handle* h, const float* mu, const float* z, float* out, void* stream);

/* Upstream Matcha-TTS speaker conditioning (the fork removed it: documentation/PROBLEMS.md:41-46; BASELINE config 5 names
 * it): when the estimator was created with in_channels = 2 * out_channels + S, `spks` is a device fp32 (batch, S) matrix
 * that the next cfm_solve / cfm_estimator broadcasts over time and concatenates after [x, mu] (upstream
 * Decoder.forward(x, mask, mu, t, spks)).  The pointer must stay valid until that call has been enqueued. */
int cfm_set_speakers(cfm_handle* h, const float* spks);

/* Utterances never interact inside the solve (GroupNorm and attention are per utterance; reference decoder.py:35-45,
 * transformer.py:253-258), so the next cfm_plan cuts the batch into up to `lanes` contiguous utterance groups of equal
 * estimated cost whose kernel chains are parallel branches of the CUDA graph (one lane's kernels fill the partial last
 * wave of another's).  A lane is never cut below `min_rows` packed full-resolution rows (<= 0: keep the current value).
 * Results do not depend on the lane count.  Default 1 lane (measured slower on cfg2-4) / 4096 rows (env CFM_B200_LANES, CFM_B200_LANE_MIN_ROWS). */
int cfm_set_lanes(cfm_handle* h, int32_t lanes, int32_t min_rows);

/* Kernel-selection switches for A/B measurements (take effect at the next cfm_plan / debug GEMM).  Results stay within the
 * precision mode's tolerance for every setting (tests/test_gpu_parity.py).
 *   "tma_epi"     bit m: TMA-store GEMM epilogue for epilogue mode m (1 = all modes; default: the in-place residual add only)
 *   "direct_epi"  bit m: 256-bit-store epilogue without shared-memory staging for bf16-output mode m (default: STORE, MASK)
 *   "pair_mode"   0/1/2: CTA-pair (cta_group::2) GEMM never / for long reductions (default) / always
 *   "pair_n256"   0/1: also run short-K GEMMs with 256-column tiles (FF1) on the pair kernel
 *   "cluster"     1/2/4: CTAs sharing one weight tile through TMA multicast in the 1-CTA kernel
 *   "pdl"         -1/0/1: programmatic dependent launch: automatic (plans of <= 2048 packed rows) / off / on
 *   "small_tiles" M: GEMMs of at most M rows use 64-column tiles (0 = never; default 1024)
 *   "l2_persist_mb" MB: persisting-L2 carve-out whose access-policy window follows the fp32 residual stream of the stage being
 *                 computed (default 32; 0 = off).  Raises the process-wide cudaLimitPersistingL2CacheSize if it is lower.
 *   "snake_warps" 8/12: epilogue warps of the SnakeBeta (FF1) GEMM (default 12)
 *   "graph_after" n: a plan's first n decodes use direct launches, its CUDA graph is captured before decode n + 1
 *                 (0 = capture inside cfm_plan; default 1)
 *   "ff_fused"    0/1: FeedForward (Linear -> SnakeBeta -> Linear + residual) as ONE kernel with the 4C hidden kept in tensor memory
 *                 (csrc/ff_fused.cuh) instead of two GEMMs through a [rows, 4C] HBM buffer; plans above "small_tiles" rows only
 *                 (default 0: measured slower)
 *   "attn_persist" 0/1: persistent attention kernel (csrc/attn_persist.cuh: 2 CTAs per SM stream the (utterance, head, query tile)
 *                 items through one software pipeline, Q and O double-buffered, deferred epilogue) instead of one CTA per item
 *                 (default 1: cfg2 attention 106 -> 99 us per full-resolution call, cfg3 -8 %; bitwise identical results)
 *   "res_side"    -1/0/1: res_conv of every resnet as a side branch of the CUDA graph (its own stream, fork / join events): automatic
 *                 (plans of <= 2048 packed rows) / off / on
 *   "rowln"       0/1/2: Linear + residual add + LayerNorm as ONE kernel (csrc/rowln.cuh: a CTA owns whole 128 x C row tiles) for
 *                 out-proj -> norm3 and FF2 -> next block's norm1: off / plans above "small_tiles" rows / always; C % 128 == 0 only
 *                 (default 0: measured equal or slower, header of rowln.cuh); "rowln_ff2" 0/1 excludes / includes the FF2 site
 *   "fp32_tc"     0/1 (fp32 handles): GEMMs and attention on the tensor pipe with bf16 x 3 split operands (CFM_PREC_FP32_TC sets it)
 *   "bf16_mid"    0/1: conv / res_conv outputs stored as bf16, statistics from the fp32 accumulators (default 1)
 *   "bn_full" / "bn_half" / "pair_min_k": tile-shape experiments for the N = C GEMMs (0 = automatic)
 *   "plan_cache"  n: plans (row tables + workspace + CUDA graph) kept per handle, least recently used evicted (default 8);
 *                 the only option that does not drop the cached plans */
int cfm_set_option(cfm_handle* h, const char* key, int32_t value);

/* Same with HOST buffers (pinned or pageable): H2D of mu and z, solve, D2H of out, then synchronises.
 * This is the call a non-PyTorch host (cgo / JNI / N-API) binds.  Fails with CFM_ERR_INVALID for an estimator with speaker
 * channels (use cfm_solve_host_spks); never reuses a pointer left by cfm_set_speakers. */
int cfm_solve_host(cfm_handle* h, const float* mu, const float* z, float* out);
/* cfm_solve_host for an estimator with S = in_channels - 2 * out_channels > 0 speaker channels: spks = host fp32 (batch, S). */
int cfm_solve_host_spks(cfm_handle* h, const float* mu, const float* z, const float* spks, float* out);

/* Multi-GPU (utterance-sharded, no collective: SURVEY.md 8(e); BASELINE config 3).  One handle per GPU, each planned with the
 * lengths of ITS utterances; this call makes the handle decode utterances index[0 .. planned batch) of a host batch of n_total
 * utterances: per-utterance H2D from mu / z (n_total, out_channels, t_pad) [spks (n_total, S) or NULL], the decode, per-utterance
 * D2H into the same positions of out.  Returns after ENQUEUEING (the handle's own stream): the host starts every GPU, then calls
 * cfm_synchronize on each.  Host buffers should be pinned (pageable memory makes the copies synchronous).  The host-side dealing of
 * utterances to GPUs (longest processing time first over the cost L (274 C^2 + 1800 C) + 24 C L^2) is sharding.py. */
int cfm_solve_host_indexed(cfm_handle* h, const float* mu, const float* z, const float* spks, float* out, const int32_t* index,
                           int32_t n_total);
/* Blocks until everything the handle has enqueued, on any stream, is complete. */
int cfm_synchronize(cfm_handle* h);

/* Front of the decode - replaces matcha/inference.py:146-167 (sequence_mask + generate_path, utils/model.py:7-40; the fp32
 * mu_x @ path matmul, :155-162; downsample = avg_pool1d(k 3, s 2, p 1), utils/model.py:57-68) without materialising the path:
 *   cfm_front_durations: durations (batch, t_x) device fp32, already rounded / clamped / masked as at inference.py:143 ->
 *                        cum (batch, t_x) int32 inclusive prefix sums, fine_lengths (batch) = max(sum, 1)      [y_fine_lengths]
 *   (the host reads fine_lengths: it needs max() for the output shape exactly like fix_len_compatibility's .item(), :148;
 *    t_pad = 2 ceil(max / 2) and lengths[b] = max((fine + 1) / 2, 1) then go straight to cfm_plan - no mask round trip)
 *   cfm_front_expand:    mu_x (batch, out_channels, t_x) -> mu_y (batch, out_channels, t_pad), y_mask (batch, t_pad) or NULL
 * Back - replaces decoder_outputs[:, :, :t_out] and denormalize (inference.py:170-172, utils/model.py:52-54):
 *   cfm_denormalize:     x (batch, out_channels, t_pad) -> out (batch, out_channels, t_out) = x * std + mean
 * All asynchronous on `stream`. */
int cfm_front_durations(cfm_handle* h, const float* durations, int32_t batch, int32_t t_x, int32_t* cum, int32_t* fine_lengths, void* stream);
int cfm_front_expand(cfm_handle* h, const float* mu_x, const int32_t* cum, const int32_t* fine_lengths, int32_t batch, int32_t t_x,
                     int32_t t_pad, float* mu_y, float* y_mask, void* stream);
int cfm_denormalize(cfm_handle* h, const float* x, int32_t batch, int32_t t_pad, int32_t t_out, float mean, float stdv, float* out, void* stream);

/* Replaces: one call of Decoder.forward(x, mask, mu, t) (decoder.py:359-426) for the planned shapes:
 * v = estimator(x, mask, mu, t); padded frames of v are zero.  Asynchronous on `stream`. */
int cfm_estimator(cfm_handle* h, const float* x, const float* mu, float t, float* v, void* stream);
/* Same with the time given as a HOST array of n_t values: n_t = 1 (one t for the batch, as the ODE solver passes it) or
 * n_t = batch (one t per utterance: the estimator call of the training forward, BASECFM.compute_loss,
 * flow_matching.py:84-97, where t has shape (B,)).  The array is consumed before the call returns. */
int cfm_estimator_t(cfm_handle* h, const float* x, const float* mu, const float* t_host, int32_t n_t, float* v, void* stream);

/* Introspection for tests and the benchmark. */
int cfm_plan_info(const cfm_handle* h, int64_t* rows_full, int64_t* rows_half, int64_t* n_nfe, int64_t* kernels_per_solve,
                  int64_t* workspace_bytes);
/* Copies a named intermediate buffer of the last estimator evaluation to the host as fp32 (debug; see cfm.cu). */
int cfm_debug_read(cfm_handle* h, const char* name, float* host_dst, int64_t max_elems, int64_t* rows, int64_t* cols);
/* Debug: within cfm_estimator (or an ungraphed solve) skip every kernel launch after the first n (n < 0: off), so that
 * cfm_debug_read sees the intermediate state at that point of the schedule. */
int cfm_debug_stop_after(cfm_handle* h, int64_t n_launches);
/* Stand-alone tapped GEMM on caller buffers (tests of the tcgen05 kernel against the fp32-FMA kernel):
 * D[m,n] = sum_t sum_k A[m + shift[t], k] * W[t*N + n, k]; A (M x K), W (n_taps*N x K) bf16 device, D fp32. */
int cfm_debug_gemm(cfm_handle* h, const void* a_bf16, const void* w_bf16, float* d_f32, int32_t M, int32_t N, int32_t K,
                   int32_t n_taps, const int32_t* shifts, int32_t use_tc, void* stream);
/* Measurement: one decode of the current plan with direct launches and a CUDA event before every launch.  buf receives one
 * text line per launch, in order: "tag,M,N,K,flops,us" (us = time to the next mark).  In-situ per-kernel times for bench.py's
 * roofline.kernel (the ncu launch lists are cold-cache and serialised).  Synchronises `stream`. */
int cfm_debug_timeline(cfm_handle* h, const float* mu, const float* z, float* out, char* buf, int64_t cap, void* stream);
/* Debug: every later attn_tc_kernel launch (direct launches only) writes CTA 0's cycle counters to prof_dev[0..16). */
int cfm_debug_attn_profile(cfm_handle* h, unsigned long long* prof_dev);
/* Debug: every later ff_fused_kernel launch (direct launches only) writes CTA 0's cycle counters to prof_dev[0..17): producer
 * [0] total [1-3] waiting for free Xn / W1 / W2 slots; MMA1 warp [4] total [5] waiting for Xn [6] W1 [14] free H columns; MMA2 warp
 * [16] total [7] waiting for P (epilogue) [8] W2 [9] Y drained; epilogue warp [10] total [11] waiting for H [12] waiting for Y
 * [13] residual update; [15] tiles of CTA 0. */
int cfm_debug_ff_profile(cfm_handle* h, unsigned long long* prof_dev);
/* Debug: every later gemm_rowln_kernel launch (direct launches only) writes CTA 0's cycle counters to prof_dev[0..16): producer
 * [0] total [1] waiting for free stages; MMA warp [2] total [3] waiting for operands [4] waiting for the drained accumulator;
 * epilogue warp [5] total [6] waiting for the accumulator [7] pass 1 (residual add + statistics) [8] exchange [9] pass 2 (normalise)
 * [10] tiles of CTA 0. */
int cfm_debug_rowln_profile(cfm_handle* h, unsigned long long* prof_dev);
/* Debug: tensor-core GEMM in one epilogue mode (0 bf16 store, 1 fp32 store, 2 fp32 in-place residual add) with per-role
 * cycle counters of CTA 0 written to prof[0..16) (device memory); see csrc/cfm.cu for the slot meanings. */
int cfm_debug_gemm_profile(cfm_handle* h, const void* a_bf16, const void* w_bf16, float* d_f32, void* d_bf16, int32_t M,
                           int32_t N, int32_t K, int32_t n_taps, const int32_t* shifts, int32_t mode,
                           unsigned long long* prof, void* stream);
#ifdef __cplusplus
}
#endif
#endif /* CFM_B200_H_ */
