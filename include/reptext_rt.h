/* reptext_rt.h — C-ABI of the B200-native RepText denoising-step runtime (librt_reptext.so).
 *
 * The reference has no FFI layer: its boundary is Python attribute calls made by the two pipelines on
 * three registered modules (RepText/pipeline_flux_controlnet.py:209-218).  The Python shims in
 * reptext_b200/ keep those call signatures and bind the entry points below with ctypes; every entry
 * point cites the reference interface it replaces.  All pointers are DEVICE pointers unless marked
 * host; no torch types cross this boundary.  Every function returns RT_OK (0) or a negative error
 * code; rt_last_error() gives the message (the shim raises ValueError / RuntimeError like the
 * reference does, e.g. RepText/controlnet_flux.py:297).  `stream` is a cudaStream_t.
 */
#ifndef REPTEXT_RT_H
#define REPTEXT_RT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 2
#if defined(__GNUC__)
#define RT_API __attribute__((visibility("default")))
#else
#define RT_API
#endif

enum { RT_OK = 0, RT_ERR_INVALID = -1, RT_ERR_CUDA = -2, RT_ERR_UNSUPPORTED = -3, RT_ERR_INTERNAL = -4 };
enum { RT_F32 = 0, RT_BF16 = 1 };             /* storage dtype of weights and activations */
enum { RT_TRANSFORMER = 0, RT_CONTROLNET = 1 };

typedef struct rt_model rt_model;

/* Fields are the register_to_config names of RepText/controlnet_flux.py:44-60. */
typedef struct rt_model_config {
  int kind;  /* RT_TRANSFORMER | RT_CONTROLNET */
  int dtype; /* RT_F32 | RT_BF16 */
  int in_channels;
  int cond_channels; /* controlnet: in_channels + extra_condition_channels */
  int out_channels;  /* transformer: proj_out features */
  int num_layers;
  int num_single_layers;
  int num_attention_heads;
  int attention_head_dim;
  int joint_attention_dim;
  int pooled_projection_dim;
  int guidance_embeds;
  int axes_dims_rope[3];
} rt_model_config;

/* ------------------------------------------------------------------------------------------------ */
RT_API const char* rt_last_error(void); /* host string, valid until the next failing call on this thread */
RT_API int rt_abi_version(void);
/* sizeof() of a public struct, so that a binding can check its mirror: 0 rt_model_config, 1 rt_forward_args,
 * 2 rt_sp_group, 3 rt_controlnet_call, 4 rt_transformer_call, 5 rt_gemm_segment, 6 rt_gemm_problem,
 * 7 rt_gemm_launch, 8 rt_attention_args, 9 rt_lnmod_group; -1 for an unknown index. */
RT_API int rt_struct_size(int which);
RT_API long long rt_launch_count(void); /* number of this library's kernels launched so far (process-wide) */
/* options: "force_simt" (0/1), "gemm_cta_group" (0 auto, 1, 2), "attn_variant" (0 auto, ...), "profile" (0/1);
 * A/B and timing aids: "ln_impl" (0 = warp-per-row-pair LayerNorm, 1 = the CTA form), "ln_warp_rows" (with ln_impl 1: 1 = the first
 * warp-per-row form), "gemv_single_row" (1 = one row per warp), "mod_inline" (1 = AdaLN GEMV on the caller's stream instead of the
 * side stream), "sp_replicate_mod" (1 = sequence-parallel ranks each compute all AdaLN rows instead of a row shard),
 * "no_pdl" (1 = plain stream-ordered launches instead of programmatic dependent launch),
 * "gemm_band" (tile order of the tcgen05 GEMM: 0 = auto - row tiles fastest unless A is too large to stay in L2 between
 * waves, then bands of row tiles swept over all column tiles; -1 = always row tiles fastest; n > 0 = bands of n row tiles),
 * "gemm_dyn_bn" (tile width of single-segment tcgen05 GEMM launches: 0 = auto - 256 columns unless a narrower tile fills
 * the last wave of the persistent grid better, as on the 1216-row shards of the sequence-parallel mode; -1 = always 256;
 * n = tiles of n columns where eligible; bit-identical results),
 * "sp_sync_kernels" (1 = the sequence-parallel phase barriers as stand-alone kernels instead of at the head of the
 * attention / output-projection kernels; same protocol, same results),
 * "gemm_epi_warps" (0 = auto: eight epilogue warps on the CTA-pair tcgen05 GEMM kernels, four on the others; 4 = four
 * everywhere - A/B, same results),
 * "mod_debug_skip" (1 = forwards skip the timestep / AdaLN chain and run on whatever vectors the workspace holds - a
 * timing experiment with WRONG results: what that chain costs inside a step),
 * "gemm_debug" (bit 1: no epilogue, bit 2: k-block 0 only - timing experiments with WRONG results;
 * bit 4: direct row-per-thread epilogue stores instead of the staged coalesced ones; bits 8 / 32 / 64: L2 eviction hints
 * on the TMA loads - A evict_last + W evict_first / W evict_last / W evict_last + A evict_first - same results, all
 * measured neutral or worse, off by default) */
RT_API int rt_set_option(const char* name, int value);
RT_API int rt_get_option(const char* name, int* value);

/* Per-class device timing for the roofline report (bench.py): with option "profile" = 1 every launch of a class
 * is bracketed by CUDA events on its own stream.  class: 0 tcgen05 GEMM, 1 SIMT GEMM, 2 tcgen05 attention,
 * 3 SIMT attention, 4 LayerNorm-modulate, 5 grouped GEMV, 6 elementwise.  work = algorithmic FLOPs (0-3) or bytes. */
RT_API int rt_profile_reset(void);
RT_API int rt_profile_read(int cls, double* ms, double* work, long long* count);

/* ---- model life cycle: replaces FluxControlNetModel.__init__ / from_pretrained
 *      (RepText/controlnet_flux.py:44-116, RepText/infer.py:30-33) and diffusers' FluxTransformer2DModel. */
RT_API int rt_model_create(const rt_model_config* cfg, rt_model** out);
/* Borrow a device tensor as parameter `name` (diffusers state-dict key).  The caller
 * keeps it alive for the life of the model.  shape is a host array. */
RT_API int rt_model_set_weight(rt_model* m, const char* name, const void* dev_ptr, const int64_t* shape, int ndim);
RT_API int rt_model_finalize(rt_model* m, void* stream); /* checks every parameter is present, builds tables */
RT_API int rt_model_destroy(rt_model* m);
RT_API int64_t rt_model_workspace_bytes(const rt_model* m, int batch, int n_img, int n_txt);

typedef struct rt_forward_args {
  int batch;     /* effective batch of the embeddings (2 under true-CFG)                       */
  int lat_batch; /* batch of hidden_states (1 or batch; 1 broadcasts — inpaint pipeline :1145) */
  int t_batch;   /* batch of timestep / guidance (1 or batch)                                  */
  int n_img, n_txt;
  const void* hidden_states;         /* [lat_batch, n_img, in_channels]     */
  const void* encoder_hidden_states; /* [batch, n_txt, joint_attention_dim] */
  const void* pooled_projections;    /* [batch, pooled_projection_dim]      */
  const void* timestep;              /* [t_batch] model dtype, sigma in [0,1] (the pipelines pass t/1000) */
  const void* guidance;              /* [t_batch] model dtype, or NULL      */
  const float* img_ids;              /* [n_img, 3] fp32 */
  const float* txt_ids;              /* [n_txt, 3] fp32 */
  void* workspace;
  int64_t workspace_bytes;
  void* stream;
  const struct rt_sp_group* sp; /* NULL: the whole sequence is on this GPU.  Otherwise n_img / n_txt / the row
                                   dimension of every tensor are THIS rank's token shard (see rt_sp_group). */
} rt_forward_args;

/* ---- sequence-parallel execution of ONE sample over `world` GPUs (BASELINE.json configs[4]: 1536x1536, 9216
 *      image + 512 text tokens).  The reference has no such mode; the layout is Ulysses-style: every rank owns
 *      n_txt text rows + n_img image rows for all token-wise work (embeds, AdaLN, LayerNorm, every GEMM) and
 *      heads [rank*H/world, (rank+1)*H/world) over the WHOLE sequence for attention.  The two exchanges per
 *      block are not separate collectives: the QKV GEMM epilogue stores each head's q|k|v columns straight into
 *      the owning rank's workspace and the attention epilogue stores each output row straight into the owning
 *      rank's workspace (peer stores over NVLink); a flag barrier between the phases orders them.
 *      Requirements: bf16, head_dim 128, H % world == 0, identical (batch, n_img, n_txt) on every rank, and
 *      every rank's workspace mapped in every process (rt_ipc_*).  RoPE follows the ids each rank passes, so
 *      img_ids / txt_ids are simply the shard's rows. */
#define RT_SP_MAX_RANKS 8
#define RT_SP_FLAG_WORDS 16 /* [0..7] arrival epochs written by rank i, [8] this rank's epoch, [9] time-out flag,
                               [11], [12] the epochs announced for the barriers that run inside the attention /
                               output-projection kernels (csrc/sp_sync.cuh) */
typedef struct rt_sp_group {
  int world, rank;
  void* peer_workspace[RT_SP_MAX_RANKS];               /* rank i's workspace as mapped in this process ([rank] = own) */
  unsigned long long* peer_flags[RT_SP_MAX_RANKS];     /* rank i's RT_SP_FLAG_WORDS words, zeroed once at set-up     */
  int lockstep; /* 1: one process drives every rank in phase order on ONE stream (rt_*_forward_lockstep; tests and
                   single-GPU validation of the exchange indexing): stream order replaces the flag barriers */
} rt_sp_group;

/* Device memory that other processes on the node can map (cudaMalloc + CUDA IPC).  handle64: 64 host bytes. */
RT_API int rt_ipc_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64);
RT_API int rt_ipc_open(const unsigned char* handle64, void** dev_ptr);
RT_API int rt_ipc_close(void* dev_ptr);
RT_API int rt_ipc_free(void* dev_ptr);
/* All-ranks barrier on `stream` (one tiny kernel: release-store of this rank's epoch into every peer's flag
 * block, acquire-spin on its own).  A rank that waits longer than ~10 s sets the ABORT word of EVERY rank's flag
 * block and moves on; the abort is sticky - every later barrier of the group returns at once - so a failed group
 * costs one time-out, not one per barrier.  rt_sp_status synchronises the stream and reports it (the pipelines
 * call it once per denoising step); rt_sp_reset zeroes this rank's flag block (epochs + abort word) and must be
 * bracketed by host-side barriers over all ranks (parallel.SequenceParallelGroup.reset). */
RT_API int rt_sp_barrier(const rt_sp_group* g, void* stream);
RT_API int rt_sp_status(const rt_sp_group* g, void* stream, int* timed_out);
RT_API int rt_sp_reset(const rt_sp_group* g, void* stream);

/* FluxControlNetModel.forward — RepText/controlnet_flux.py:216-413, called at
 * RepText/pipeline_flux_controlnet.py:1043-1056 and pipeline_flux_controlnet_inpaint.py:1167, :1214.
 * block_samples: [num_layers, batch, n_img, D]; single_block_samples: [num_single_layers, batch, n_img, D]
 * (NULL when num_single_layers == 0).  The conditioning scale (:395-396) is fused into the zero-linear
 * epilogue.  mask ([n_img], model dtype, or NULL) fuses the pipelines' regional-mask multiply
 * (pipeline_flux_controlnet.py:1060-1069) and accumulate != 0 fuses the multi-line sum (:1072-1087):
 * out = (zero_linear(h) * scale * mask) + (accumulate ? out : 0). */
RT_API int rt_controlnet_forward(rt_model* m, const rt_forward_args* a, const void* controlnet_cond, int cond_batch,
                          float conditioning_scale, const void* mask, int accumulate, void* block_samples,
                          void* single_block_samples);

/* Tell a ControlNet how many of its samples have a consumer.  diffusers' FluxTransformer2DModel.forward adds
 * controlnet_block_samples[i // ceil(L / n)] after block i (L = 19, n = 6 -> interval 4 -> samples 0..4; sample 5 is
 * never read), so the pair FLUX.1-dev + RepText pays for a sixth ControlNet block and zero-linear
 * (RepText/controlnet_flux.py:320-349, :385-388) whose output nobody uses: 1.38 TFLOP of an 82.7 TFLOP step.
 * With live_layers / live_single_layers set (>= 0) rt_controlnet_forward only runs the blocks that feed a consumed sample
 * and leaves the other samples of the output stack UNTOUCHED (the caller zero-fills them once).  -1 = run everything
 * (the default: a C-ABI caller that reads every sample gets every sample).  The Python pipelines, which only ever hand
 * the lists to `transformer(...)`, set it from the transformer's layer counts. */
RT_API int rt_controlnet_set_live(rt_model* m, int live_layers, int live_single_layers);

/* Step-invariant inputs.  Of what a forward computes, three things depend only on the prompt
 * embeddings, the pooled / guidance vectors and the position ids - which are the same in all 28 steps of an image -
 * and not on the latents or the timestep: context_embedder(encoder_hidden_states) (RepText/controlnet_flux.py:292), the
 * FluxPosEmbed table (:316-317) and the first linears of the guidance and pooled-text embedders (:282-291; the reference
 * even rebuilds the guidance tensor every step, pipeline_flux_controlnet.py:1029).  mode 1: the first forward after this
 * call computes them into model-owned device memory, later forwards reuse them for as long as the
 * encoder_hidden_states / pooled_projections / guidance / txt_ids / img_ids POINTERS and the shapes are the ones of
 * that first forward (the text rows are copied into the residual stream: same bits as the GEMM that produced them, so
 * results are bit-identical).  The caller owns validity: call again (any mode) whenever the CONTENTS behind those
 * pointers change - every call invalidates.  mode 0 (the default) computes everything in every forward.  Not used by
 * the lock-step entry points.  The first forward after an invalidation may allocate: keep it out of a stream capture. */
RT_API int rt_model_set_step_invariant_cache(rt_model* m, int mode);

/* The AdaLN vectors of ALL the steps of an image in one pass.  What a forward derives from (timestep, guidance,
 * pooled_projections) alone - CombinedTimestepGuidanceTextProjEmbeddings (RepText/controlnet_flux.py:282-291) and the
 * AdaLayerNorm linears of every block, 6.5 GB of weights streamed per FLUX.1-dev forward - does not depend on the
 * latents, and a denoising loop knows its timesteps before it starts (pipeline_flux_controlnet.py:1017: `for i, t in
 * enumerate(timesteps)`).  rt_model_build_modulation_table computes that chain for `steps` x `batch` rows at once into
 * `table` - caller-owned device memory of rt_model_modulation_table_bytes(m, steps, batch) bytes, 256-byte aligned, that
 * must stay alive and untouched while a row is selected (timesteps / guidance: [steps * batch] in the model dtype, row s * batch + b = what step s passes
 * for batch element b, i.e. timestep / 1000; pooled_projections: [steps * batch, pooled_projection_dim]); the weights
 * are read from DRAM once instead of once per step.  rt_model_select_modulation(m, s) makes the following forwards of
 * this model (batch must match) use row s and skip the chain (their `timestep` / `guidance` / `pooled_projections`
 * arguments are then ignored - the caller owns the correspondence); -1 returns to computing per forward (the C-ABI
 * default).  Every row holds the bits the forward's own chain produces: results are bit-identical. */
RT_API int64_t rt_model_modulation_table_bytes(const rt_model* m, int steps, int batch);
RT_API int rt_model_build_modulation_table(rt_model* m, const void* timesteps, const void* guidance,
                                           const void* pooled_projections, int steps, int batch, void* table,
                                           int64_t table_bytes, void* stream);
RT_API int rt_model_select_modulation(rt_model* m, int step);

/* FluxTransformer2DModel.forward (diffusers 0.36.0) as called at
 * RepText/pipeline_flux_controlnet.py:1092-1104.  controlnet_*_samples are host arrays of device
 * pointers to [batch, n_img, D] tensors (or NULL / 0); sample i//ceil(L/n) is added after block i,
 * fused into that block's last GEMM epilogue.  out: [batch, n_img, out_channels]. */
RT_API int rt_transformer_forward(rt_model* m, const rt_forward_args* a, const void* const* controlnet_block_samples,
                           int n_block_samples, const void* const* controlnet_single_block_samples,
                           int n_single_block_samples, void* out);

/* Lock-step forms of the two forwards: calls[i] carries rank i's arguments (calls[i].a.sp->rank == i,
 * lockstep == 1, every workspace on the current device); the phases of all ranks are issued in order on
 * calls[0].a.stream.  Same arithmetic and the same peer-store indexing as the multi-process mode. */
typedef struct rt_controlnet_call {
  rt_forward_args a;
  const void* controlnet_cond;
  int cond_batch;
  float conditioning_scale;
  const void* mask;
  int accumulate;
  void* block_samples;
  void* single_block_samples;
} rt_controlnet_call;
typedef struct rt_transformer_call {
  rt_forward_args a;
  const void* const* controlnet_block_samples;
  int n_block_samples;
  const void* const* controlnet_single_block_samples;
  int n_single_block_samples;
  void* out;
} rt_transformer_call;
RT_API int rt_controlnet_forward_lockstep(rt_model* m, int world, const rt_controlnet_call* calls);
RT_API int rt_transformer_forward_lockstep(rt_model* m, int world, const rt_transformer_call* calls);

/* FlowMatchEulerDiscreteScheduler.step — called at RepText/pipeline_flux_controlnet.py:1109:
 * out = dtype( float(sample) + dtype((sigma_next - sigma) * model_output) ); n elements. */
RT_API int rt_euler_step(int dtype, const void* model_output, const void* sample, void* out, int64_t n, float sigma,
                  float sigma_next, void* stream);
/* true-CFG combine, RepText/pipeline_flux_controlnet_inpaint.py:1264-1270: v2 = [uncond; text] (2 x n);
 * zero_pred != 0 is the i == 0 branch (noise_pred = text * 0). */
RT_API int rt_cfg_combine(int dtype, const void* v2, void* out, int64_t n, float true_guidance_scale, int zero_pred,
                   void* stream);
/* CFG combine fused with the Euler step (same arithmetic, one pass) */
RT_API int rt_cfg_euler_step(int dtype, const void* v2, const void* sample, void* out, int64_t n,
                      float true_guidance_scale, int zero_pred, float sigma, float sigma_next, void* stream);
/* Unfused regional-mask multiply + multi-line sum for ControlNet outputs that did not come from this
 * library (pipeline_flux_controlnet.py:1060-1087): out = mask[r] * scale * x + (acc_in ? acc_in : 0). */
RT_API int rt_mask_scale_add(int dtype, const void* x, const void* mask, const void* acc_in, void* out, int batch,
                      int rows, int D, float scale, void* stream);
/* Glyph-latent init blend, RepText/pipeline_flux_controlnet_inpaint.py:643-647:
 * out = mask ? w_glyph * glyph_latents + w_noise * noise : noise  (mask: uint8 per element). */
RT_API int rt_glyph_init_blend(int dtype, const void* noise, const void* glyph_latents, const unsigned char* mask,
                        void* out, int64_t n, float w_glyph, float w_noise, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Operator-level entry points (used by the parity tests and by bench.py's roofline leg).
 * ------------------------------------------------------------------------------------------------ */
enum {
  RT_EPI_BIAS = 0,        /* acc + bias                                                          */
  RT_EPI_GELU = 1,        /* gelu_tanh(acc + bias)                                               */
  RT_EPI_QKNORM_ROPE = 2, /* per head: rmsnorm(acc + bias) * w, then interleaved-pair RoPE       */
  RT_EPI_GATE_RESID = 3,  /* out = out + gate[b, n] * (acc + bias) (+ extra[b, m, n]), in place  */
  RT_EPI_SCALE_MASK = 4   /* out = (acc + bias) * scale * mask[m] (+ out if accumulate)          */
};

typedef struct rt_gemm_segment {
  const void* W;    /* [n_end - n_begin, K] row-major (nn.Linear weight) */
  const void* bias; /* [n_end - n_begin] or NULL */
  int n_begin, n_end;
  int mode;
  void* out; /* [batch, rows, out_ld] */
  int64_t out_batch_stride;
  int out_ld;
  int out_col0;
  const void* norm_w; /* RT_EPI_QKNORM_ROPE: [head_dim] */
  int scatter; /* 1 (tcgen05 path, BIAS / QKNORM_ROPE only): column block c = (n - n_begin) / sp_cols of this segment
                  goes to launch.sp_out[c] at row sp_row0 + out_row0 + m, column out_col0 + (n - n_begin) % sp_cols;
                  `out` is ignored, out_batch_stride / out_ld describe the destination buffers */
  int out_f32; /* 1 (tcgen05 path; RT_EPI_BIAS, or RT_EPI_SCALE_MASK without mask / accumulate): `out` is FP32 - the
                  accumulator leaves without a bf16 rounding (out_batch_stride / out_ld / out_col0 count fp32 elements).
                  Used for the VAE mid-block's attention scores, whose softmax is taken over fp32 logits like the
                  reference's (diffusers AutoencoderKL -> SDPA) */
} rt_gemm_segment;

typedef struct rt_gemm_problem {
  const void* A; /* [batch, a_rows_total, a_ld] */
  int64_t a_batch_stride; /* 0 broadcasts one batch */
  int a_ld;
  int a_row0;
  int a_rows_total;
  int m_rows;
  int out_row0;
  int K;
  int nseg;
  rt_gemm_segment seg[4];
  const float* gate; /* [batch, gate_ld] fp32 or NULL (= 1) */
  int gate_ld;
  const void* extra; /* optional addend; problem rows >= extra_row0 map to extra rows 0.. */
  int64_t extra_batch_stride;
  int extra_ld;
  int extra_row0;
  float scale;
  const void* mask; /* [m_rows] or NULL */
  int accumulate;
  /* 3x3 convolution, stride 1, zero padding 1, as an implicit GEMM (tcgen05 path only; the VAE of SURVEY.md 8f):
   * A is an NHWC image [batch, conv_h, conv_w, a_ld >= conv_c], m_rows = conv_h * conv_w output pixels per image, and
   * the weights are [n, 9 * ceil(conv_c / 64) * 64] with K index tap * (ceil(conv_c / 64) * 64) + channel
   * (tap = ky * 3 + kx, zero padded).  conv_w = 0: plain GEMM. */
  int conv_h, conv_w, conv_c;
} rt_gemm_problem;

typedef struct rt_gemm_launch {
  int dtype;
  int batch;
  int nprob;
  rt_gemm_problem prob[2];
  const float* rope; /* [rows, head_dim/2, 2] (cos, sin) fp32, indexed by out_row0 + m; NULL = no rope */
  int head_dim;
  int sp_cols; /* scatter segments: columns per destination (a multiple of 128); 0 = no scatter segment */
  int sp_row0; /* row offset of this rank's rows in the destination buffers */
  void* sp_out[8]; /* destination buffers (peer-mapped device pointers) */
} rt_gemm_launch;

/* impl: 0 auto, 1 SIMT, 2 tcgen05 cta_group::1, 3 tcgen05 cta_group::2 */
RT_API int rt_gemm(const rt_gemm_launch* g, int impl, void* stream);

typedef struct rt_attention_args {
  int dtype;
  const void* qkv; /* [batch, S, ld]; head h of q at column q_col0 + h*hd, etc. */
  int64_t batch_stride;
  int ld;
  int q_col0, k_col0, v_col0;
  void* out; /* [batch, S, out_ld]; head h at column out_col0 + h*hd */
  int64_t out_batch_stride;
  int out_ld, out_col0;
  int batch, S, heads, hd;
  int sp_rows; /* > 0 (tcgen05 path): output row r goes to sp_out[r / sp_rows] at row r % sp_rows (same out_ld,
                  out_col0, out_batch_stride); `out` is ignored */
  int sp_txt_rows; /* with sp_rows > 0: the first sp_txt_rows rows of every rank's sp_rows-row shard are its TEXT rows.
                  When sp_txt_rows % 64 == 0, (S / sp_rows * sp_txt_rows) % 128 == 0 and (sp_rows - sp_txt_rows) % 128
                  == 0 the keys are walked in the UNSHARDED order (every rank's text rows, then every rank's image
                  rows): same key blocks, same summation order, bit-identical rows to the single-GPU launch.
                  0 = walk the buffer as it lies (rank-major) */
  void* sp_out[8];
} rt_attention_args;
/* impl: 0 auto, 1 SIMT, >= 2 tcgen05 variant (impl - 2) */
RT_API int rt_attention(const rt_attention_args* a, int impl, void* stream);

typedef struct rt_lnmod_group {
  int row_begin, row_end;
  const float* shift; /* [batch, ld] */
  const float* scale;
  int ld;
} rt_lnmod_group;
RT_API int rt_layernorm_modulate(int dtype, const void* x, int64_t x_batch_stride, int x_ld, void* out,
                          int64_t out_batch_stride, int out_ld, int batch, int D, int ngroups,
                          const rt_lnmod_group* groups /* host */, void* stream);
/* FluxPosEmbed (RepText/controlnet_flux.py:65, :316-317): ids [S,3] fp32 -> out [S, sum(axes)/2, 2] fp32 */
RT_API int rt_rope_table(const float* ids, int S, const int* axes_dims /* host, 3 */, float* out, void* stream);
/* In-place per-head RMSNorm * w + RoPE on `heads` heads starting at column col0 (SIMT path's unfused form) */
RT_API int rt_qknorm_rope(int dtype, void* buf, int64_t batch_stride, int ld, int col0, int batch, int row0, int rows,
                   int heads, int hd, const void* norm_w, const float* rope, int rope_row0, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VAE path: diffusers AutoencoderKL as the pipelines call it
 * (RepText/pipeline_flux_controlnet.py:705-715 encode, :1136-1140 decode).  Activations are NHWC bf16; the
 * convolutions run on rt_gemm (conv_h / conv_w / conv_c, or plain GEMMs over pixels / im2col rows).
 * ------------------------------------------------------------------------------------------------ */
/* GroupNorm (+ SiLU when silu != 0) over x [batch, hw, C]; stats_ws: batch * groups * 24 bytes of scratch, 8-byte aligned. */
RT_API int rt_groupnorm_nhwc(const void* x, void* out, int batch, int64_t hw, int C, int groups, const void* gamma,
                             const void* beta, float eps, int silu, void* stats_ws, void* stream);
/* Upsample2D's nearest x2: [batch, H, W, C] -> [batch, 2H, 2W, C] */
RT_API int rt_upsample_nearest2x_nhwc(const void* in, void* out, int batch, int H, int W, int C, void* stream);
/* In-place softmax of every row of a [rows, cols] matrix (the single-head mid-block attention) */
RT_API int rt_softmax_rows(void* x, int64_t rows, int cols, int64_t ld, void* stream);
/* softmax over each row of an FP32 [rows, cols] matrix (row stride ld_in), written as bf16 to `out` (row stride
 * ld_out); the row is held in registers: one read, one write.  cols % 4 == 0, cols <= 65536 */
RT_API int rt_softmax_rows_f32(const void* x, int64_t rows, int cols, int64_t ld_in, void* out, int64_t ld_out,
                               void* stream);
/* im2col of a 3x3 convolution (stride, leading padding pad_lo) for the convolutions TMA cannot address (stride 2,
 * 3 input channels): in [batch, H, W, c_ld] -> out [batch, Ho * Wo, Kp], K index tap * C + c, zero beyond 9 * C */
RT_API int rt_im2col3x3_nhwc(const void* in, void* out, int batch, int H, int W, int C, int c_ld, int Ho, int Wo,
                             int stride, int pad_lo, int Kp, void* stream);

/* Layout changes at the VAE's edges: NCHW (RT_F32 or RT_BF16) -> NHWC bf16 [batch, hw, c_pad] (channels >= C zero), and
 * NHWC bf16 [batch, hw, ld] (first C channels) -> NCHW (RT_F32 or RT_BF16) */
RT_API int rt_nchw_to_nhwc(int src_dtype, const void* in, void* out, int batch, int C, int64_t hw, int c_pad, void* stream);
RT_API int rt_nhwc_to_nchw(const void* in, int ld, void* out, int dst_dtype, int batch, int C, int64_t hw, void* stream);
/* DiagonalGaussianDistribution.sample (encode(x).latent_dist.sample(), RepText/pipeline_flux_controlnet.py:705-708) on
 * NHWC bf16 moments [batch, hw, ld] (mean = channels 0..L, logvar = L..2L): out NCHW [batch, L, hw] =
 * mean + exp(0.5 * clamp(logvar, -30, 20)) * noise; noise NCHW in `dtype` like out, NULL = the mode */
RT_API int rt_vae_posterior_sample(const void* moments, int ld, int latent_channels, int batch, int64_t hw,
                                   const void* noise, void* out, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Prompt-encoder path: transformers' T5EncoderModel and CLIPTextModel as the pipelines call
 * them (RepText/pipeline_flux_controlnet.py:232-347).  bf16 only; the projections / MLPs run on rt_gemm.
 * ------------------------------------------------------------------------------------------------ */
/* out[r, :] = (x[r, :] - mean?) * rsqrt(var + eps) * weight (+ bias): subtract_mean = 0 is T5LayerNorm (RMS, bias NULL),
 * 1 is nn.LayerNorm.  D <= 4096, multiple of 8. */
RT_API int rt_norm_rows(const void* x, int64_t x_ld, void* out, int64_t out_ld, int64_t rows, int D, const void* weight,
                        const void* bias, float eps, int subtract_mean, void* stream);
/* softmax(scale * q k^T + rel_bias[h][key - query + S - 1], keys <= query if causal) v for head_dim 64;
 * q / k / v at column offsets of one [batch, S, ld] buffer, head-major; rel_bias [heads, 2 S - 1] fp32 or NULL */
RT_API int rt_text_attention(const void* qkv, int64_t batch_stride, int ld, int q_col0, int k_col0, int v_col0, void* out,
                             int64_t out_batch_stride, int out_ld, int out_col0, int batch, int S, int heads, int hd,
                             float scale, const float* rel_bias, int causal, void* stream);
/* kind 0: out[r, :F] = in[r, :F] * in[r, F:2F] (T5's gated-GELU; the GELU is the GEMM epilogue's);
 * kind 1: out[r, :F] = quick_gelu(in[r, :F]) (CLIP) */
RT_API int rt_glu_act(int kind, const void* in, int64_t in_ld, void* out, int64_t out_ld, int64_t rows, int F, void* stream);
/* out[i, :] = table[ids[i], :] (+ pos_table[i % S, :]); ids: int64 on the device; *bad_flag (device int, zeroed by the
 * caller) is set when an id is outside [0, vocab) */
RT_API int rt_embedding(const void* table, int64_t vocab, int D, const int64_t* ids, int64_t n, const void* pos_table, int S,
                        void* out, int* bad_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REPTEXT_RT_H */
