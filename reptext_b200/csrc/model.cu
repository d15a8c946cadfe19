// Model-level runtime: FluxControlNetModel.forward (RepText/controlnet_flux.py:216-413) and the
// FluxTransformer2DModel.forward it conditions (diffusers 0.36.0, called at
// RepText/pipeline_flux_controlnet.py:1092-1104), driven from C++ so that one forward is a straight
// sequence of kernel launches on the caller's stream (graph-capturable: no allocation, no sync).
//
// HBM layout (all in the caller-provided workspace; T = text tokens, N = image tokens, S = T + N):
//   x    [B, S, D]   residual stream, TEXT ROWS FIRST then image rows (the order attention uses)
//   xn   [B, S, D]   LayerNorm + AdaLN-modulated copy feeding the projections
//   qkv  [B, S, 3D]  q | k | v, head-major; q and k already RMS-normed and rotated by the GEMM epilogue
//   cat  [B, S, 5D]  attention output (cols 0..D) | MLP hidden (cols D..5D): the single-stream block's
//                    concat (proj_out reads it with K = 5D), and the double block's two scratch areas
//   mod  [B, M] fp32 every block's AdaLN vectors, produced by ONE grouped GEMV per forward
// A double block is 7 launches (2 LN-modulate, 4 grouped GEMMs over the text and image problems,
// 1 attention); a single block is 4.
#include <map>
#include <string>
#include <vector>

#include "dtype_utils.cuh"
#include "rt_internal.h"
#include "sp_sync.cuh"

namespace rt {
namespace {

struct Lin {
  const void* W = nullptr;
  const void* b = nullptr;
  int n = 0, k = 0;
};
struct DoubleBlk {
  Lin norm1, norm1c, q, k, v, aq, ak, av, o, ao, ff1, ff2, cff1, cff2;
  const void *nq = nullptr, *nk = nullptr, *naq = nullptr, *nak = nullptr;
  int mod_img = 0, mod_ctx = 0;
};
struct SingleBlk {
  Lin norm, q, k, v, mlp, out;
  const void *nq = nullptr, *nk = nullptr;
  int mod = 0;
};

struct Workspace {
  float *sin_t, *sin_g, *pooled, *hid, *temb, *temb_s, *mod;
  float2* rope;
  char *x, *xn, *qkv, *cat;
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace rt

using namespace rt;

struct rt_model {
  rt_model_config cfg{};
  int D = 0, hd = 0, H = 0;
  struct Wt {
    const void* p;
    std::vector<int64_t> shape;
  };
  std::map<std::string, Wt> w;
  bool finalized = false;
  Lin x_emb, ctx_emb, cnx_emb, t1, t2, g1, g2, p1, p2, norm_out, proj_out;
  std::vector<DoubleBlk> dbl;
  std::vector<SingleBlk> sgl;
  std::vector<Lin> cn_blk, cn_sgl;
  int mod_total = 0, mod_out = 0;
  GemvJob* jobs_dev = nullptr;  // [0..5] = t1 g1 p1 t2 g2 p2, [6..] = AdaLN linears of every block
  int* prefix_dev = nullptr;    // [0] = 0 (single-job launches), [1..] = row prefix of the AdaLN jobs
  int n_mod_jobs = 0, mod_rows = 0;
  int first_block_jobs = 0, first_block_rows = 0;  // the AdaLN jobs the first block needs
  // ControlNet only: how many of the double / single blocks have a CONSUMER (rt_controlnet_set_live; -1 = all of them)
  int live_layers = -1, live_single = -1;
  // The AdaLN vectors of blocks 1.. are weight-streaming work (6.5 GB per forward) that nothing needs until the
  // second block: they are computed on a side stream, under the first block's tensor-core kernels.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // Step-invariant inputs (rt_model_set_step_invariant_cache): what a forward derives from the prompt
  // embeddings, the pooled / guidance vectors and the position ids alone - context_embedder(enc)
  // (controlnet_flux.py:292), the FluxPosEmbed table (:316-317), the first linears of the guidance and pooled-text MLPs
  // (:282-291) - is the same in all 28 steps of an image: computed on the first forward after an invalidation, kept in
  // model-owned memory, reused while the input pointers and shapes stay the same.
  struct StepInvariants {
    int mode = 0;        // 0 = off (the C-ABI default), 1 = on
    bool valid = false;
    const void *enc = nullptr, *pooled = nullptr, *guidance = nullptr, *txt_ids = nullptr, *img_ids = nullptr;
    int B = 0, T = 0, N = 0;
    char* buf = nullptr;
    size_t bytes = 0;
    char* ctx = nullptr;     // [B, T, D] model dtype
    float2* rope = nullptr;  // [T + N, head_dim / 2]
    float* hid = nullptr;    // [B, 3 D]: h_t (every step) | h_g | h_p (kept)
  } inv;

  // AdaLN vectors of ALL the steps of an image (rt_model_build_modulation_table): the timesteps of a denoising loop are
  // known before it starts, the guidance and pooled vectors do not change - so the time / guidance / pooled-text MLPs and
  // the AdaLN linears of every block (6.5 GB of weights per forward, a GEMV) can run for all steps in ONE pass that
  // reads the weights once.  A forward with a selected row (rt_model_select_modulation) points its `mod` at the table
  // and skips that whole chain.
  struct ModTable {
    float* mod = nullptr;  // [rows, mod_total], inside the caller's buffer
    int steps = 0, batch = 0;
    int sel = -1;          // the step the next forwards use; -1: compute per forward
  } modtab;

  ~rt_model() {
    if (inv.buf) cudaFree(inv.buf);
    if (jobs_dev) cudaFree(jobs_dev);
    if (prefix_dev) cudaFree(prefix_dev);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
  }
};

namespace rt {
namespace {

Lin get_linear(const rt_model& m, const std::string& name, int n, int k) {
  auto iw = m.w.find(name + ".weight");
  if (iw == m.w.end()) throw Error(RT_ERR_INVALID, "missing parameter " + name + ".weight");
  const auto& sw = iw->second.shape;
  if (sw.size() != 2 || sw[0] != n || sw[1] != k)
    throw Error(RT_ERR_INVALID, "parameter " + name + ".weight has the wrong shape (expected [" + std::to_string(n) +
                                    ", " + std::to_string(k) + "])");
  Lin l;
  l.W = iw->second.p;
  l.n = n;
  l.k = k;
  auto ib = m.w.find(name + ".bias");
  if (ib != m.w.end()) {
    if (ib->second.shape.size() != 1 || ib->second.shape[0] != n)
      throw Error(RT_ERR_INVALID, "parameter " + name + ".bias has the wrong shape");
    l.b = ib->second.p;
  }
  return l;
}
const void* get_vec(const rt_model& m, const std::string& name, int n) {
  auto it = m.w.find(name);
  if (it == m.w.end()) throw Error(RT_ERR_INVALID, "missing parameter " + name);
  if (it->second.shape.size() != 1 || it->second.shape[0] != n)
    throw Error(RT_ERR_INVALID, "parameter " + name + " has the wrong shape");
  return it->second.p;
}

size_t carve(const rt_model& m, int B, int N, int T, char* base, Workspace* ws) {
  const size_t es = dtype_size(m.cfg.dtype);
  const size_t S = (size_t)T + N, D = m.D;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes);
    return base ? base + o : nullptr;
  };
  Workspace w{};
  w.sin_t = (float*)take((size_t)B * 256 * 4);
  w.sin_g = (float*)take((size_t)B * 256 * 4);
  w.pooled = (float*)take((size_t)B * align_up(m.cfg.pooled_projection_dim, 4) * 4);
  w.hid = (float*)take((size_t)B * 3 * D * 4);
  w.temb = (float*)take((size_t)B * D * 4);
  w.temb_s = (float*)take((size_t)B * D * 4);
  w.mod = (float*)take((size_t)B * m.mod_total * 4);
  w.rope = (float2*)take(S * (m.hd / 2) * sizeof(float2));
  w.x = take((size_t)B * S * D * es);
  w.xn = take((size_t)B * S * D * es);
  w.qkv = take((size_t)B * S * 3 * D * es);
  w.cat = take((size_t)B * S * 5 * D * es);
  if (ws) *ws = w;
  return off;
}

GemmSegment make_seg(const Lin& l, int n_begin, int mode, void* out, long long obs, int old, int ocol0,
                     const void* norm_w = nullptr) {
  GemmSegment s{};
  s.W = l.W;
  s.bias = l.b;
  s.n_begin = n_begin;
  s.n_end = n_begin + l.n;
  s.mode = mode;
  s.out = out;
  s.out_batch_stride = obs;
  s.out_ld = old;
  s.out_col0 = ocol0;
  s.norm_w = norm_w;
  return s;
}
GemmProblem make_prob(const void* A, long long a_bs, int a_ld, int a_row0, int a_rows_total, int m_rows, int out_row0,
                      int K) {
  GemmProblem p{};
  p.A = A;
  p.a_batch_stride = a_bs;
  p.a_ld = a_ld;
  p.a_row0 = a_row0;
  p.a_rows_total = a_rows_total;
  p.m_rows = m_rows;
  p.out_row0 = out_row0;
  p.K = K;
  p.scale = 1.f;
  return p;
}

struct Ctx {
  const rt_model& m;
  Workspace ws;
  int B, T, N, S, D, dt;
  size_t es;
  cudaStream_t st;
  // sequence-parallel mode (rt_sp_group): P ranks, this is rank r; T / N / S above are THIS rank's rows, the
  // attention runs over Sg = P * S rows and Dl = D / P of the head columns.  peer_qkv / peer_cat: every rank's
  // exchange buffers as mapped here.
  const rt_sp_group* sp = nullptr;
  int P = 1, r = 0, Sg = 0, Dl = 0;
  char* peer_qkv[RT_SP_MAX_RANKS] = {};
  char* peer_cat[RT_SP_MAX_RANKS] = {};
  float* peer_mod[RT_SP_MAX_RANKS] = {};
  bool inv_on = false, inv_hit = false;  // step-invariant cache in use / its contents are current (StepInvariants)
  bool mod_sharded = false;       // this forward computed only its row shard of the AdaLN vectors (peer stores)
  bool mod_join_pending = false;  // the AdaLN vectors of blocks 1.. are still being computed on the side stream
  // Real multi-process group: the two phase barriers of a block run INSIDE the consuming kernels (sp_sync.cuh): the QKV
  // GEMM announces the pre-attention barrier's epoch, attention runs that barrier at its head and announces the next,
  // the output projection runs the post-attention barrier at its head - unless option "sp_sync_kernels" asks for the
  // stand-alone barrier kernels (A/B).
  bool fused_sync = false;
  SpSyncParams sync_announce{}, sync_attn{}, sync_post{};
  const SpSyncParams* sync(int which /*0 QKV GEMM, 1 attention, 2 output projection*/) const {
    if (!fused_sync) return nullptr;
    return which == 0 ? &sync_announce : which == 1 ? &sync_attn : &sync_post;
  }
  long long sD() const { return (long long)S * D; }
};

// The q | k | v segments of a projection launch: local [B, S, 3D] layout, or (sequence-parallel) scattered by head
// block into every rank's [B, Sg, 3 Dl] exchange buffer at this rank's row range.
void set_qkv_targets(const Ctx& c, GemmLaunch& L) {
  if (c.P == 1) return;
  L.sp_cols = c.Dl;
  L.sp_row0 = c.r * c.S;
  for (int i = 0; i < c.P; ++i) L.sp_out[i] = c.peer_qkv[i];
}
GemmSegment make_qkv_seg(const Ctx& c, const Lin& l, int which /*0 q, 1 k, 2 v*/, const void* norm_w) {
  const int mode = norm_w ? EPI_QKNORM_ROPE : EPI_BIAS;
  if (c.P == 1)
    return make_seg(l, which * c.D, mode, c.ws.qkv, 3 * c.sD(), 3 * c.D, which * c.D, norm_w);
  GemmSegment s = make_seg(l, which * c.D, mode, nullptr, (long long)c.Sg * 3 * c.Dl, 3 * c.Dl, which * c.Dl, norm_w);
  s.scatter = 1;
  return s;
}

// temb = MLP_t(sin(1000 t)) + MLP_g(sin(1000 g)) + MLP_p(pooled)   (controlnet_flux.py:282-291), then the AdaLN
// vectors of every block: mod = Linear_i(SiLU(temb)).
void time_text_and_modulation(Ctx& c, const rt_forward_args& a) {
  const rt_model& m = c.m;
  const Workspace& w = c.ws;
  const int D = c.D, B = c.B, P = m.cfg.pooled_projection_dim;
  const bool has_g = m.cfg.guidance_embeds != 0;
  launch_time_sinusoid(c.dt, a.timestep, a.t_batch, B, w.sin_t, c.st);
  // first linears (+ SiLU) into hid = [h_t | h_g | h_p]; h_g and h_p do not depend on the step (kept when cached)
  launch_gemv_grouped(c.dt, w.sin_t, 256, B, 256, m.jobs_dev + 0, m.prefix_dev, 1, D, w.hid, 3 * D, 1, 0, c.st);
  if (!c.inv_hit) {
    if (has_g) {
      launch_time_sinusoid(c.dt, a.guidance, a.t_batch, B, w.sin_g, c.st);
      launch_gemv_grouped(c.dt, w.sin_g, 256, B, 256, m.jobs_dev + 1, m.prefix_dev, 1, D, w.hid, 3 * D, 1, 0, c.st);
    }
    launch_cast_to_f32(c.dt, a.pooled_projections, w.pooled, (long long)B * P, c.st);
    launch_gemv_grouped(c.dt, w.pooled, P, B, P, m.jobs_dev + 2, m.prefix_dev, 1, D, w.hid, 3 * D, 1, 0, c.st);
  }
  // second linears accumulate into temb
  launch_gemv_grouped(c.dt, w.hid, 3 * D, B, D, m.jobs_dev + 3, m.prefix_dev, 1, D, w.temb, D, 0, 0, c.st);
  if (has_g)
    launch_gemv_grouped(c.dt, w.hid + D, 3 * D, B, D, m.jobs_dev + 4, m.prefix_dev, 1, D, w.temb, D, 0, 1, c.st);
  launch_gemv_grouped(c.dt, w.hid + 2 * D, 3 * D, B, D, m.jobs_dev + 5, m.prefix_dev, 1, D, w.temb, D, 0, 1, c.st);
  launch_silu_f32(w.temb, w.temb_s, (long long)B * D, c.st);
  const bool r4 = D % 4 == 0;  // every AdaLN job has k * D rows
  const int n0 = m.first_block_jobs, rows0 = m.first_block_rows;
  if (m.side && n0 > 0 && n0 < m.n_mod_jobs && !get_option("mod_inline")) {
    // first block's vectors on the caller's stream, the rest forked onto the side stream (joined before block 1)
    int rest = m.mod_rows - rows0, base = rows0;
    GemvPeers peers{};
    const bool shard = c.P > 1 && r4 && rest % (4 * c.P) == 0 && !get_option("sp_replicate_mod");
    if (shard) {
      // Sequence-parallel: every rank streams only 1 / P of the AdaLN weights and stores its rows into all ranks'
      // `mod` buffers.  The peers may still be reading the previous forward's vectors (ControlNet and transformer
      // share the workspace): a barrier first; the matching barrier after the join is in run_block.
      if (!c.sp->lockstep) launch_sp_barrier(*c.sp, c.st);
      rest /= c.P;
      base += c.r * rest;
      peers.n = c.P;
      for (int i = 0; i < c.P; ++i) peers.p[i] = c.peer_mod[i];
      c.mod_sharded = true;
    }
    RT_CHECK_CUDA(cudaEventRecord(m.ev_fork, c.st));
    RT_CHECK_CUDA(cudaStreamWaitEvent(m.side, m.ev_fork, 0));
    launch_gemv_grouped(c.dt, w.temb_s, D, B, D, m.jobs_dev + 6, m.prefix_dev + 1, m.n_mod_jobs, rest, w.mod,
                        m.mod_total, 0, 0, m.side, r4, base, shard ? &peers : nullptr);
    RT_CHECK_CUDA(cudaEventRecord(m.ev_join, m.side));
    launch_gemv_grouped(c.dt, w.temb_s, D, B, D, m.jobs_dev + 6, m.prefix_dev + 1, m.n_mod_jobs, rows0, w.mod,
                        m.mod_total, 0, 0, c.st, r4, 0);
    c.mod_join_pending = true;
  } else {
    launch_gemv_grouped(c.dt, w.temb_s, D, B, D, m.jobs_dev + 6, m.prefix_dev + 1, m.n_mod_jobs, m.mod_rows, w.mod,
                        m.mod_total, 0, 0, c.st, r4);
  }
}

// every AdaLN vector is needed from here on: join the side stream (no-op when nothing was forked)
void join_modulation(Ctx& c) {
  if (!c.mod_join_pending) return;
  RT_CHECK_CUDA(cudaStreamWaitEvent(c.st, c.m.ev_join, 0));
  c.mod_join_pending = false;
}

// The chain of time_text_and_modulation for `rows` = steps x batch (timestep, guidance, pooled) triples at once, into
// model-owned memory.  Same kernels and the same arithmetic per row as a forward's own chain (the small MLPs in batch
// passes, the AdaLN linears through the table GEMV that keeps every accumulator's operation order): a forward that uses
// row i is bit-identical to one that computes it.
size_t modulation_table_layout(const rt_model& m, size_t rows, size_t* o /*[7]*/) {
  const size_t D = m.D, P = m.cfg.pooled_projection_dim;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t at = off; off = align_up(off + bytes); return at; };
  const size_t at[7] = {take(rows * 256 * 4), take(rows * 256 * 4), take(rows * align_up(P, 4) * 4), take(rows * 3 * D * 4),
                        take(rows * D * 4), take(rows * D * 4), take(rows * m.mod_total * 4)};
  if (o) for (int i = 0; i < 7; ++i) o[i] = at[i];
  return off;
}

void build_modulation_table(rt_model& m, const void* timesteps, const void* guidance, const void* pooled, int steps,
                            int batch, char* buf, size_t buf_bytes, cudaStream_t st) {
  const int D = m.D, P = m.cfg.pooled_projection_dim, dt = m.cfg.dtype;
  const bool has_g = m.cfg.guidance_embeds != 0;
  const size_t rows = (size_t)steps * batch;
  rt_model::ModTable& t = m.modtab;
  t.sel = -1;
  t.mod = nullptr;
  t.steps = 0;
  size_t o[7];
  RT_REQUIRE(buf && (reinterpret_cast<uintptr_t>(buf) & 255) == 0 && buf_bytes >= modulation_table_layout(m, rows, o),
             "build_modulation_table: buffer missing, not 256-byte aligned or too small (rt_model_modulation_table_bytes)");
  const size_t o_sin_t = o[0], o_sin_g = o[1], o_pool = o[2], o_hid = o[3], o_temb = o[4], o_temb_s = o[5], o_mod = o[6];
  float* sin_t = (float*)(buf + o_sin_t); float* sin_g = (float*)(buf + o_sin_g);
  float* poolf = (float*)(buf + o_pool);  float* hid = (float*)(buf + o_hid);
  float* temb = (float*)(buf + o_temb);   float* temb_s = (float*)(buf + o_temb_s);
  t.mod = (float*)(buf + o_mod);
  const int R = (int)rows;
  launch_time_sinusoid(dt, timesteps, R, R, sin_t, st);
  launch_gemv_grouped(dt, sin_t, 256, R, 256, m.jobs_dev + 0, m.prefix_dev, 1, D, hid, 3 * D, 1, 0, st);
  if (has_g) {
    launch_time_sinusoid(dt, guidance, R, R, sin_g, st);
    launch_gemv_grouped(dt, sin_g, 256, R, 256, m.jobs_dev + 1, m.prefix_dev, 1, D, hid, 3 * D, 1, 0, st);
  }
  launch_cast_to_f32(dt, pooled, poolf, (long long)rows * P, st);
  launch_gemv_grouped(dt, poolf, P, R, P, m.jobs_dev + 2, m.prefix_dev, 1, D, hid, 3 * D, 1, 0, st);
  launch_gemv_grouped(dt, hid, 3 * D, R, D, m.jobs_dev + 3, m.prefix_dev, 1, D, temb, D, 0, 0, st);
  if (has_g) launch_gemv_grouped(dt, hid + D, 3 * D, R, D, m.jobs_dev + 4, m.prefix_dev, 1, D, temb, D, 0, 1, st);
  launch_gemv_grouped(dt, hid + 2 * D, 3 * D, R, D, m.jobs_dev + 5, m.prefix_dev, 1, D, temb, D, 0, 1, st);
  launch_silu_f32(temb, temb_s, (long long)rows * D, st);
  if (D % 4 == 0 && m.mod_rows % 4 == 0)
    launch_gemv_grouped_table(dt, temb_s, D, R, D, m.jobs_dev + 6, m.prefix_dev + 1, m.n_mod_jobs, m.mod_rows, t.mod,
                              m.mod_total, st);
  else
    launch_gemv_grouped(dt, temb_s, D, R, D, m.jobs_dev + 6, m.prefix_dev + 1, m.n_mod_jobs, m.mod_rows, t.mod,
                        m.mod_total, 0, 0, st);
  t.steps = steps;
  t.batch = batch;
}

void embed_inputs(const Ctx& c, rt_model* mm, const rt_forward_args& a, const void* cond, int cond_batch) {
  const rt_model& m = c.m;
  const int D = c.D, T = c.T, N = c.N, J = m.cfg.joint_attention_dim, Cin = m.cfg.in_channels;
  // context_embedder (controlnet_flux.py:292) -> text rows of x; with the step-invariant cache the GEMM runs once per
  // image into model-owned memory and every forward copies its rows (the same bits) into the residual stream
  if (!c.inv_hit) {
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 1;
    L.prob[0] = make_prob(a.encoder_hidden_states, (long long)T * J, J, 0, T, T, 0, J);
    L.prob[0].nseg = 1;
    L.prob[0].seg[0] = c.inv_on ? make_seg(m.ctx_emb, 0, EPI_BIAS, mm->inv.ctx, (long long)T * D, D, 0)
                                : make_seg(m.ctx_emb, 0, EPI_BIAS, c.ws.x, c.sD(), D, 0);
    launch_gemm(L, c.st);
  }
  if (c.inv_on) {
    RT_CHECK_CUDA(cudaMemcpy2DAsync(c.ws.x, (size_t)c.sD() * c.es, mm->inv.ctx, (size_t)T * D * c.es, (size_t)T * D * c.es,
                                    (size_t)c.B, cudaMemcpyDeviceToDevice, c.st));
    mm->inv.valid = true;  // rope, h_g | h_p (begin_forward) and ctx are all on their way
  }
  // x_embedder (:277) -> image rows of x; batch-1 latents broadcast against batch-2 embeddings
  {
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 1;
    const long long bs = (a.lat_batch == 1 && c.B > 1) ? 0 : (long long)N * Cin;
    L.prob[0] = make_prob(a.hidden_states, bs, Cin, 0, N, N, T, Cin);
    L.prob[0].nseg = 1;
    L.prob[0].seg[0] = make_seg(m.x_emb, 0, EPI_BIAS, c.ws.x, c.sD(), D, 0);
    launch_gemm(L, c.st);
  }
  if (cond) {  // + controlnet_x_embedder(controlnet_cond) (:280)
    const int Cc = m.cfg.cond_channels;
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 1;
    const long long bs = (cond_batch == 1 && c.B > 1) ? 0 : (long long)N * Cc;
    L.prob[0] = make_prob(cond, bs, Cc, 0, N, N, T, Cc);
    L.prob[0].nseg = 1;
    L.prob[0].seg[0] = make_seg(m.cnx_emb, 0, EPI_GATE_RESID, c.ws.x, c.sD(), D, 0);
    launch_gemm(L, c.st);
  }
}

void run_attention(const Ctx& c) {
  AttnArgs t{};
  t.dtype = c.dt;
  t.out_batch_stride = (long long)c.S * 5 * c.D; t.out_ld = 5 * c.D;
  t.batch = c.B; t.hd = c.m.hd;
  if (c.P == 1) {
    t.qkv = c.ws.qkv; t.batch_stride = (long long)c.S * 3 * c.D; t.ld = 3 * c.D;
    t.q_col0 = 0; t.k_col0 = c.D; t.v_col0 = 2 * c.D;
    t.out = c.ws.cat; t.out_col0 = 0;
    t.S = c.S; t.heads = c.m.H;
  } else {
    // this rank's heads over the whole sequence; output rows go back to the rank that owns the token
    t.qkv = c.ws.qkv; t.batch_stride = (long long)c.Sg * 3 * c.Dl; t.ld = 3 * c.Dl;
    t.q_col0 = 0; t.k_col0 = c.Dl; t.v_col0 = 2 * c.Dl;
    t.out = nullptr; t.out_col0 = c.r * c.Dl;
    t.S = c.Sg; t.heads = c.m.H / c.P;
    t.sp_rows = c.S;
    t.sp_txt_rows = c.T;  // keys in the unsharded order when the shard sizes allow it: same bits as on one GPU
    for (int i = 0; i < c.P; ++i) t.sp_out[i] = c.peer_cat[i];
  }
  launch_attention(t, c.st, c.sync(1));  // sequence-parallel: the pre-attention barrier runs at its head
}

// FluxTransformerBlock (diffusers).  `extra`: ControlNet residual added to the image rows
// after the block ([B, N, D], or null) - fused into the last GEMM's epilogue.
void double_block_pre(const Ctx& c, const DoubleBlk& k) {
  const Workspace& w = c.ws;
  const int D = c.D, T = c.T, N = c.N, S = c.S, ld = c.m.mod_total;
  const float* mi = w.mod + k.mod_img;
  const float* mc = w.mod + k.mod_ctx;
  const long long sD = c.sD();
  {
    LnModGroup g[2] = {{0, T, mc, mc + D, ld}, {T, S, mi, mi + D, ld}};
    launch_ln_mod(c.dt, w.x, sD, D, w.xn, sD, D, c.B, D, 2, g, c.st);
  }
  {
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 2; L.rope = reinterpret_cast<const float*>(w.rope); L.head_dim = c.m.hd;
    set_qkv_targets(c, L);
    GemmProblem& pt = L.prob[0];
    pt = make_prob(w.xn, sD, D, 0, S, T, 0, D);
    pt.nseg = 3;
    pt.seg[0] = make_qkv_seg(c, k.aq, 0, k.naq);
    pt.seg[1] = make_qkv_seg(c, k.ak, 1, k.nak);
    pt.seg[2] = make_qkv_seg(c, k.av, 2, nullptr);
    GemmProblem& pi = L.prob[1];
    pi = make_prob(w.xn, sD, D, T, S, N, T, D);
    pi.nseg = 3;
    pi.seg[0] = make_qkv_seg(c, k.q, 0, k.nq);
    pi.seg[1] = make_qkv_seg(c, k.k, 1, k.nk);
    pi.seg[2] = make_qkv_seg(c, k.v, 2, nullptr);
    launch_gemm(L, c.st, c.sync(0));  // sequence-parallel: announces the pre-attention barrier's epoch
  }
}

void double_block_post(const Ctx& c, const DoubleBlk& k, const void* extra) {
  const Workspace& w = c.ws;
  const int D = c.D, T = c.T, N = c.N, S = c.S, ld = c.m.mod_total;
  const float* mi = w.mod + k.mod_img;
  const float* mc = w.mod + k.mod_ctx;
  const long long sD = c.sD(), s5D = 5 * sD;
  {  // x += gate_msa * to_out(attn)
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 2;
    L.prob[0] = make_prob(w.cat, s5D, 5 * D, 0, S, T, 0, D);
    L.prob[0].nseg = 1;
    L.prob[0].seg[0] = make_seg(k.ao, 0, EPI_GATE_RESID, w.x, sD, D, 0);
    L.prob[0].gate = mc + 2 * D; L.prob[0].gate_ld = ld;
    L.prob[1] = make_prob(w.cat, s5D, 5 * D, T, S, N, T, D);
    L.prob[1].nseg = 1;
    L.prob[1].seg[0] = make_seg(k.o, 0, EPI_GATE_RESID, w.x, sD, D, 0);
    L.prob[1].gate = mi + 2 * D; L.prob[1].gate_ld = ld;
    launch_gemm(L, c.st, c.sync(2));  // sequence-parallel: the post-attention barrier runs at its head
  }
  {
    LnModGroup g[2] = {{0, T, mc + 3 * D, mc + 4 * D, ld}, {T, S, mi + 3 * D, mi + 4 * D, ld}};
    launch_ln_mod(c.dt, w.x, sD, D, w.xn, sD, D, c.B, D, 2, g, c.st);
  }
  {  // MLP hidden -> cat[:, :, D:5D]
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 2;
    L.prob[0] = make_prob(w.xn, sD, D, 0, S, T, 0, D);
    L.prob[0].nseg = 1;
    L.prob[0].seg[0] = make_seg(k.cff1, 0, EPI_GELU, w.cat, s5D, 5 * D, D);
    L.prob[1] = make_prob(w.xn, sD, D, T, S, N, T, D);
    L.prob[1].nseg = 1;
    L.prob[1].seg[0] = make_seg(k.ff1, 0, EPI_GELU, w.cat, s5D, 5 * D, D);
    launch_gemm(L, c.st);
  }
  {  // x += gate_mlp * ff2(hidden) (+ ControlNet residual on the image rows)
    const char* h = w.cat + (size_t)D * c.es;
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 2;
    L.prob[0] = make_prob(h, s5D, 5 * D, 0, S, T, 0, 4 * D);
    L.prob[0].nseg = 1;
    L.prob[0].seg[0] = make_seg(k.cff2, 0, EPI_GATE_RESID, w.x, sD, D, 0);
    L.prob[0].gate = mc + 5 * D; L.prob[0].gate_ld = ld;
    L.prob[1] = make_prob(h, s5D, 5 * D, T, S, N, T, 4 * D);
    L.prob[1].nseg = 1;
    L.prob[1].seg[0] = make_seg(k.ff2, 0, EPI_GATE_RESID, w.x, sD, D, 0);
    L.prob[1].gate = mi + 5 * D; L.prob[1].gate_ld = ld;
    if (extra) {
      L.prob[1].extra = extra; L.prob[1].extra_batch_stride = (long long)N * D; L.prob[1].extra_ld = D;
      L.prob[1].extra_row0 = 0;
    }
    launch_gemm(L, c.st);
  }
}

// FluxSingleTransformerBlock on the joint sequence.  `extra` ([B, N, D]) is added to the
// image rows (problem rows >= T).
void single_block_pre(const Ctx& c, const SingleBlk& k) {
  const Workspace& w = c.ws;
  const int D = c.D, S = c.S, ld = c.m.mod_total;
  const float* md = w.mod + k.mod;
  const long long sD = c.sD(), s5D = 5 * sD;
  {
    LnModGroup g[1] = {{0, S, md, md + D, ld}};
    launch_ln_mod(c.dt, w.x, sD, D, w.xn, sD, D, c.B, D, 1, g, c.st);
  }
  {  // q | k | v | mlp in one launch
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 1; L.rope = reinterpret_cast<const float*>(w.rope); L.head_dim = c.m.hd;
    set_qkv_targets(c, L);
    GemmProblem& p = L.prob[0];
    p = make_prob(w.xn, sD, D, 0, S, S, 0, D);
    p.nseg = 4;
    p.seg[0] = make_qkv_seg(c, k.q, 0, k.nq);
    p.seg[1] = make_qkv_seg(c, k.k, 1, k.nk);
    p.seg[2] = make_qkv_seg(c, k.v, 2, nullptr);
    p.seg[3] = make_seg(k.mlp, 3 * D, EPI_GELU, w.cat, s5D, 5 * D, D);
    launch_gemm(L, c.st, c.sync(0));  // sequence-parallel: announces the pre-attention barrier's epoch
  }
}

void single_block_post(const Ctx& c, const SingleBlk& k, const void* extra) {
  const Workspace& w = c.ws;
  const int D = c.D, T = c.T, N = c.N, S = c.S, ld = c.m.mod_total;
  const float* md = w.mod + k.mod;
  const long long sD = c.sD(), s5D = 5 * sD;
  {  // x += gate * proj_out([attn | mlp])
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 1;
    GemmProblem& p = L.prob[0];
    p = make_prob(w.cat, s5D, 5 * D, 0, S, S, 0, 5 * D);
    p.nseg = 1;
    p.seg[0] = make_seg(k.out, 0, EPI_GATE_RESID, w.x, sD, D, 0);
    p.gate = md + 2 * D; p.gate_ld = ld;
    if (extra) {
      p.extra = extra; p.extra_batch_stride = (long long)N * D; p.extra_ld = D; p.extra_row0 = T;
    }
    launch_gemm(L, c.st, c.sync(2));  // sequence-parallel: the post-attention barrier runs at its head
  }
}

Ctx begin_forward(rt_model* m, const rt_forward_args* a) {
  RT_REQUIRE(m && a, "null model / args");
  RT_REQUIRE(m->finalized, "rt_model_finalize has not been called");
  RT_REQUIRE(a->batch >= 1 && a->n_img >= 1 && a->n_txt >= 1, "batch, n_img and n_txt must be positive");
  RT_REQUIRE(a->lat_batch == 1 || a->lat_batch == a->batch, "hidden_states batch must be 1 or the embedding batch");
  RT_REQUIRE(a->t_batch == 1 || a->t_batch == a->batch, "timestep batch must be 1 or the embedding batch");
  RT_REQUIRE(a->hidden_states && a->encoder_hidden_states && a->pooled_projections && a->timestep && a->img_ids &&
                 a->txt_ids,
             "null input tensor");
  RT_REQUIRE(!m->cfg.guidance_embeds || a->guidance, "this model has guidance_embeds: `guidance` is required");
  RT_REQUIRE(a->workspace, "null workspace");
  Workspace ws;
  size_t need = carve(*m, a->batch, a->n_img, a->n_txt, (char*)a->workspace, &ws);
  RT_REQUIRE((size_t)a->workspace_bytes >= need, "workspace too small (see rt_model_workspace_bytes)");
  RT_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "workspace must be 256-byte aligned");
  Ctx c{*m, ws, a->batch, a->n_txt, a->n_img, a->n_txt + a->n_img, m->D, m->cfg.dtype, dtype_size(m->cfg.dtype),
        (cudaStream_t)a->stream};
  if (a->sp && a->sp->world > 1) {
    const rt_sp_group& g = *a->sp;
    RT_REQUIRE(g.world <= RT_SP_MAX_RANKS && g.rank >= 0 && g.rank < g.world, "sp group: world / rank");
    RT_REQUIRE(m->cfg.dtype == RT_BF16 && m->hd == 128, "sequence-parallel mode needs bf16 and head_dim 128");
    RT_REQUIRE(m->H % g.world == 0, "sequence-parallel mode: the head count must be a multiple of the world size");
    RT_REQUIRE(g.peer_workspace[g.rank] == a->workspace, "sp group: peer_workspace[rank] must be this call's workspace");
    c.sp = &g;
    c.P = g.world;
    c.r = g.rank;
    c.Sg = c.S * g.world;
    c.Dl = m->D / g.world;
    for (int i = 0; i < g.world; ++i) {
      RT_REQUIRE(g.peer_workspace[i] && (reinterpret_cast<uintptr_t>(g.peer_workspace[i]) & 255) == 0,
                 "sp group: peer workspace missing or not 256-byte aligned");
      RT_REQUIRE(g.lockstep || g.peer_flags[i], "sp group: null flag pointer");
      Workspace pw;
      carve(*m, a->batch, a->n_img, a->n_txt, (char*)g.peer_workspace[i], &pw);  // every rank carves identically
      c.peer_qkv[i] = pw.qkv;
      c.peer_cat[i] = pw.cat;
      c.peer_mod[i] = pw.mod;
    }
    if (!g.lockstep && !get_option("sp_sync_kernels")) {
      c.fused_sync = true;
      SpSyncParams d{};
      d.world = g.world;
      d.rank = g.rank;
      for (int i = 0; i < g.world; ++i) d.flags[i] = g.peer_flags[i];
      c.sync_announce = c.sync_attn = c.sync_post = d;
      c.sync_announce.announce_word = 11;
      c.sync_attn.barrier_word = 11;
      c.sync_attn.announce_word = 12;
      c.sync_post.barrier_word = 12;
    }
  }
  // Step-invariant cache (one model per call only: lock-step ranks share the model object)
  rt_model::StepInvariants& inv = m->inv;
  if (inv.mode && !(a->sp && a->sp->lockstep)) {
    c.inv_on = true;
    c.inv_hit = inv.valid && inv.enc == a->encoder_hidden_states && inv.pooled == a->pooled_projections &&
                inv.guidance == a->guidance && inv.txt_ids == a->txt_ids && inv.img_ids == a->img_ids &&
                inv.B == c.B && inv.T == c.T && inv.N == c.N;
    if (!c.inv_hit) {
      inv.valid = false;
      const size_t ctx_b = align_up((size_t)c.B * c.T * c.D * c.es), rope_b = align_up((size_t)c.S * (m->hd / 2) * 8),
                   hid_b = align_up((size_t)c.B * 3 * c.D * 4);
      if (inv.bytes < ctx_b + rope_b + hid_b) {
        if (inv.buf) RT_CHECK_CUDA(cudaFree(inv.buf));  // (synchronises: only when an image is larger than any before)
        inv.buf = nullptr; inv.bytes = 0;
        RT_CHECK_CUDA(cudaMalloc(&inv.buf, ctx_b + rope_b + hid_b));
        inv.bytes = ctx_b + rope_b + hid_b;
      }
      inv.ctx = inv.buf;
      inv.rope = reinterpret_cast<float2*>(inv.buf + ctx_b);
      inv.hid = reinterpret_cast<float*>(inv.buf + ctx_b + rope_b);
      inv.enc = a->encoder_hidden_states; inv.pooled = a->pooled_projections; inv.guidance = a->guidance;
      inv.txt_ids = a->txt_ids; inv.img_ids = a->img_ids;
      inv.B = c.B; inv.T = c.T; inv.N = c.N;
    }
    c.ws.rope = inv.rope;
    c.ws.hid = inv.hid;
  }
  // FluxPosEmbed over cat(txt_ids, img_ids) (controlnet_flux.py:316-317)
  if (!c.inv_hit) {
    launch_rope_table(a->txt_ids, c.T, m->cfg.axes_dims_rope, c.ws.rope, c.st);
    launch_rope_table(a->img_ids, c.N, m->cfg.axes_dims_rope, c.ws.rope + (size_t)c.T * (m->hd / 2), c.st);
  }
  if (m->modtab.sel >= 0) {
    // this step's AdaLN vectors were computed with the whole table: nothing of the timestep chain runs
    RT_REQUIRE(m->modtab.batch == c.B, "modulation table: built for another batch size");
    c.ws.mod = m->modtab.mod + (size_t)m->modtab.sel * c.B * m->mod_total;
  } else if (!get_option("mod_debug_skip")) {  // (timing experiment: stale AdaLN vectors)
    time_text_and_modulation(c, *a);
  }
  return c;
}

// ---- phased execution -------------------------------------------------------------------------------------
// One block = pre (LayerNorm-modulate + QKV(+MLP) projection) | attention | post (output projections, MLP).
// `cs` holds ONE context (single GPU, or this process's rank of a sequence-parallel group) or, in lock-step
// mode, the contexts of every rank: each phase is issued for all of them before the next one starts.  Between
// the phases of a real multi-process group sits the flag barrier that orders the peer stores.
void phase_sync(const std::vector<Ctx>& cs) {
  if (cs.size() == 1 && cs[0].P > 1 && !cs[0].sp->lockstep) launch_sp_barrier(*cs[0].sp, cs[0].st);
}
// between pre | attention | post: inside the neighbouring kernels when the context says so (Ctx::fused_sync)
void block_phase_sync(const std::vector<Ctx>& cs) {
  if (!(cs.size() == 1 && cs[0].fused_sync)) phase_sync(cs);
}
template <class Pre, class Post>
void run_block(std::vector<Ctx>& cs, Pre pre, Post post) {
  for (size_t i = 0; i < cs.size(); ++i) pre(cs[i], i);
  block_phase_sync(cs);
  for (size_t i = 0; i < cs.size(); ++i) run_attention(cs[i]);
  block_phase_sync(cs);
  for (size_t i = 0; i < cs.size(); ++i) post(cs[i], i);
  // after the FIRST block (no-op later): the AdaLN vectors of blocks 1.. are needed from here on
  bool sharded = false;
  for (size_t i = 0; i < cs.size(); ++i) {
    sharded = sharded || (cs[i].mod_join_pending && cs[i].mod_sharded);
    join_modulation(cs[i]);
  }
  if (sharded) phase_sync(cs);  // every rank's row shard has landed in every rank's buffer
}

void check_lockstep(int world, size_t ncalls, const rt_forward_args* const* args) {
  RT_REQUIRE(world >= 1 && world <= RT_SP_MAX_RANKS && ncalls == (size_t)world, "lockstep: world");
  if (world == 1) return;
  for (int i = 0; i < world; ++i) {
    const rt_forward_args* a = args[i];
    RT_REQUIRE(a->sp && a->sp->lockstep && a->sp->world == world && a->sp->rank == i,
               "lockstep: calls[i].a.sp must be a lock-step group with rank == i");
    RT_REQUIRE(a->stream == args[0]->stream, "lockstep: every rank must use the same stream");
    RT_REQUIRE(a->batch == args[0]->batch && a->n_img == args[0]->n_img && a->n_txt == args[0]->n_txt,
               "lockstep: every rank must hold the same shard shape");
  }
}

void controlnet_forward_impl(rt_model* m, const std::vector<const rt_controlnet_call*>& calls) {
  RT_REQUIRE(m && m->cfg.kind == RT_CONTROLNET, "not a ControlNet model");
  std::vector<Ctx> cs;
  for (const rt_controlnet_call* k : calls) {
    RT_REQUIRE(k->controlnet_cond && (k->cond_batch == 1 || k->cond_batch == k->a.batch), "controlnet_cond batch");
    RT_REQUIRE(m->cfg.num_layers == 0 || k->block_samples, "block_samples is null");
    RT_REQUIRE(m->cfg.num_single_layers == 0 || k->single_block_samples, "single_block_samples is null");
    cs.push_back(begin_forward(m, &k->a));
    embed_inputs(cs.back(), m, k->a, k->controlnet_cond, k->cond_batch);
  }
  auto zero_linear = [&](const Ctx& c, const rt_controlnet_call& k, const Lin& zl, void* base, int idx) {
    // controlnet_flux.py:385-396 (+ the pipelines' regional mask and multi-line sum)
    const long long sample_elems = (long long)c.B * c.N * c.D;
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 1;
    GemmProblem& p = L.prob[0];
    p = make_prob(c.ws.x, c.sD(), c.D, c.T, c.S, c.N, 0, c.D);
    p.nseg = 1;
    p.seg[0] = make_seg(zl, 0, EPI_SCALE_MASK, (char*)base + (size_t)idx * sample_elems * c.es, (long long)c.N * c.D,
                        c.D, 0);
    p.scale = k.conditioning_scale;
    p.mask = k.mask;
    p.accumulate = k.accumulate;
    launch_gemm(L, c.st);
  };
  // Blocks whose sample nobody consumes are not run (rt_controlnet_set_live): a single block's sample only feeds later
  // single blocks, a double block's sample feeds every later block - so trailing singles can always be dropped, trailing
  // doubles only when no single block runs after them.  The skipped samples are left untouched.
  const int ns_live = (m->live_single >= 0 && m->live_single < m->cfg.num_single_layers) ? m->live_single
                                                                                        : m->cfg.num_single_layers;
  const int nl_live = (ns_live == 0 && m->live_layers >= 0 && m->live_layers < m->cfg.num_layers) ? m->live_layers
                                                                                                  : m->cfg.num_layers;
  for (int i = 0; i < nl_live; ++i)
    run_block(cs, [&](const Ctx& c, size_t) { double_block_pre(c, m->dbl[i]); },
              [&](const Ctx& c, size_t r) {
                double_block_post(c, m->dbl[i], nullptr);
                zero_linear(c, *calls[r], m->cn_blk[i], calls[r]->block_samples, i);
              });
  for (int j = 0; j < ns_live; ++j)
    run_block(cs, [&](const Ctx& c, size_t) { single_block_pre(c, m->sgl[j]); },
              [&](const Ctx& c, size_t r) {
                single_block_post(c, m->sgl[j], nullptr);
                zero_linear(c, *calls[r], m->cn_sgl[j], calls[r]->single_block_samples, j);
              });
}

void transformer_forward_impl(rt_model* m, const std::vector<const rt_transformer_call*>& calls) {
  RT_REQUIRE(m && m->cfg.kind == RT_TRANSFORMER, "not a transformer model");
  std::vector<Ctx> cs;
  for (const rt_transformer_call* k : calls) {
    RT_REQUIRE(k->out, "null output");
    RT_REQUIRE(k->n_block_samples >= 0 && k->n_single_block_samples >= 0, "negative sample count");
    RT_REQUIRE(k->n_block_samples == 0 || k->controlnet_block_samples, "controlnet_block_samples is null");
    RT_REQUIRE(k->n_single_block_samples == 0 || k->controlnet_single_block_samples, "controlnet_single_block_samples");
    RT_REQUIRE(k->n_block_samples == calls[0]->n_block_samples &&
                   k->n_single_block_samples == calls[0]->n_single_block_samples, "ranks disagree on the sample counts");
    cs.push_back(begin_forward(m, &k->a));
    embed_inputs(cs.back(), m, k->a, nullptr, 0);
  }
  const int nl = m->cfg.num_layers, ns = m->cfg.num_single_layers;
  const int nbs = calls[0]->n_block_samples, nss = calls[0]->n_single_block_samples;
  // residual injection: sample[i // ceil(L / n)] after block i (diffusers FluxTransformer2DModel.forward)
  const int iv_d = nbs ? (nl + nbs - 1) / nbs : 1;
  const int iv_s = nss ? (ns + nss - 1) / nss : 1;
  for (int i = 0; i < nl; ++i)
    run_block(cs, [&](const Ctx& c, size_t) { double_block_pre(c, m->dbl[i]); },
              [&](const Ctx& c, size_t r) {
                double_block_post(c, m->dbl[i], nbs ? calls[r]->controlnet_block_samples[i / iv_d] : nullptr);
              });
  for (int j = 0; j < ns; ++j)
    run_block(cs, [&](const Ctx& c, size_t) { single_block_pre(c, m->sgl[j]); },
              [&](const Ctx& c, size_t r) {
                single_block_post(c, m->sgl[j], nss ? calls[r]->controlnet_single_block_samples[j / iv_s] : nullptr);
              });
  // norm_out (AdaLayerNormContinuous: chunk order scale, shift) + proj_out on the image rows
  {
    bool sharded = false;  // (a model without blocks: the join has not happened yet)
    for (size_t r = 0; r < cs.size(); ++r) {
      sharded = sharded || (cs[r].mod_join_pending && cs[r].mod_sharded);
      join_modulation(cs[r]);
    }
    if (sharded) phase_sync(cs);
  }
  for (size_t r = 0; r < cs.size(); ++r) {
    const Ctx& c = cs[r];
    const int D = c.D, T = c.T, N = c.N, S = c.S;
    const float* mo = c.ws.mod + m->mod_out;
    LnModGroup g[1] = {{T, S, mo + D, mo, m->mod_total}};
    launch_ln_mod(c.dt, c.ws.x, c.sD(), D, c.ws.xn, c.sD(), D, c.B, D, 1, g, c.st);
    const int Co = m->cfg.out_channels;
    GemmLaunch L{};
    L.dtype = c.dt; L.batch = c.B; L.nprob = 1;
    L.prob[0] = make_prob(c.ws.xn, c.sD(), D, T, S, N, 0, D);
    L.prob[0].nseg = 1;
    L.prob[0].seg[0] = make_seg(m->proj_out, 0, EPI_BIAS, calls[r]->out, (long long)N * Co, Co, 0);
    launch_gemm(L, c.st);
  }
}

}  // namespace
}  // namespace rt

extern "C" {

int rt_model_create(const rt_model_config* cfg, rt_model** out) {
  return guarded([&] {
    RT_REQUIRE(cfg && out, "null argument");
    RT_REQUIRE(cfg->kind == RT_TRANSFORMER || cfg->kind == RT_CONTROLNET, "kind");
    RT_REQUIRE(cfg->dtype == RT_F32 || cfg->dtype == RT_BF16, "dtype");
    RT_REQUIRE(cfg->num_layers >= 0 && cfg->num_single_layers >= 0, "layer counts");
    RT_REQUIRE(cfg->num_attention_heads > 0 && (cfg->attention_head_dim == 64 || cfg->attention_head_dim == 128),
               "attention_head_dim must be 64 or 128");
    RT_REQUIRE(cfg->axes_dims_rope[0] + cfg->axes_dims_rope[1] + cfg->axes_dims_rope[2] == cfg->attention_head_dim,
               "sum(axes_dims_rope) must equal attention_head_dim");
    RT_REQUIRE(cfg->in_channels > 0 && cfg->joint_attention_dim > 0 && cfg->pooled_projection_dim > 0, "dims");
    RT_REQUIRE(cfg->pooled_projection_dim % 4 == 0, "pooled_projection_dim must be a multiple of 4");
    if (cfg->kind == RT_CONTROLNET) RT_REQUIRE(cfg->cond_channels > 0, "cond_channels");
    else RT_REQUIRE(cfg->out_channels > 0, "out_channels");
    auto* m = new rt_model();
    m->cfg = *cfg;
    m->hd = cfg->attention_head_dim;
    m->H = cfg->num_attention_heads;
    m->D = m->hd * m->H;
    *out = m;
  });
}

int rt_model_set_weight(rt_model* m, const char* name, const void* dev_ptr, const int64_t* shape, int ndim) {
  return guarded([&] {
    RT_REQUIRE(m && name && dev_ptr && shape && ndim >= 1 && ndim <= 2, "set_weight: bad argument");
    RT_REQUIRE(!m->finalized, "set_weight after finalize");
    RT_REQUIRE((reinterpret_cast<uintptr_t>(dev_ptr) & 15) == 0, "parameters must be 16-byte aligned");
    rt_model::Wt w{dev_ptr, std::vector<int64_t>(shape, shape + ndim)};
    m->w[name] = w;
  });
}

int rt_model_finalize(rt_model* m, void* stream) {
  return guarded([&] {
    RT_REQUIRE(m && !m->finalized, "finalize: bad model");
    const rt_model_config& c = m->cfg;
    const int D = m->D, hd = m->hd;
    const std::string tte = "time_text_embed.";
    m->x_emb = get_linear(*m, "x_embedder", D, c.in_channels);
    m->ctx_emb = get_linear(*m, "context_embedder", D, c.joint_attention_dim);
    m->t1 = get_linear(*m, tte + "timestep_embedder.linear_1", D, 256);
    m->t2 = get_linear(*m, tte + "timestep_embedder.linear_2", D, D);
    if (c.guidance_embeds) {
      m->g1 = get_linear(*m, tte + "guidance_embedder.linear_1", D, 256);
      m->g2 = get_linear(*m, tte + "guidance_embedder.linear_2", D, D);
    }
    m->p1 = get_linear(*m, tte + "text_embedder.linear_1", D, c.pooled_projection_dim);
    m->p2 = get_linear(*m, tte + "text_embedder.linear_2", D, D);

    std::vector<GemvJob> jobs(6);
    auto job = [](const Lin& l, int off) { return GemvJob{l.W, l.b, l.n, off}; };
    jobs[0] = job(m->t1, 0);
    jobs[1] = c.guidance_embeds ? job(m->g1, D) : GemvJob{nullptr, nullptr, 0, 0};
    jobs[2] = job(m->p1, 2 * D);
    jobs[3] = job(m->t2, 0);
    jobs[4] = c.guidance_embeds ? job(m->g2, 0) : GemvJob{nullptr, nullptr, 0, 0};
    jobs[5] = job(m->p2, 0);
    std::vector<int> prefix = {0};
    int off = 0, rows = 0;
    auto add_mod = [&](const Lin& l) {
      int o = off;
      jobs.push_back(job(l, off));
      prefix.push_back(rows);
      off += l.n;
      rows += l.n;
      return o;
    };

    m->dbl.resize(c.num_layers);
    for (int i = 0; i < c.num_layers; ++i) {
      const std::string p = "transformer_blocks." + std::to_string(i) + ".";
      DoubleBlk& b = m->dbl[i];
      b.norm1 = get_linear(*m, p + "norm1.linear", 6 * D, D);
      b.norm1c = get_linear(*m, p + "norm1_context.linear", 6 * D, D);
      b.q = get_linear(*m, p + "attn.to_q", D, D);
      b.k = get_linear(*m, p + "attn.to_k", D, D);
      b.v = get_linear(*m, p + "attn.to_v", D, D);
      b.aq = get_linear(*m, p + "attn.add_q_proj", D, D);
      b.ak = get_linear(*m, p + "attn.add_k_proj", D, D);
      b.av = get_linear(*m, p + "attn.add_v_proj", D, D);
      b.o = get_linear(*m, p + "attn.to_out.0", D, D);
      b.ao = get_linear(*m, p + "attn.to_add_out", D, D);
      b.nq = get_vec(*m, p + "attn.norm_q.weight", hd);
      b.nk = get_vec(*m, p + "attn.norm_k.weight", hd);
      b.naq = get_vec(*m, p + "attn.norm_added_q.weight", hd);
      b.nak = get_vec(*m, p + "attn.norm_added_k.weight", hd);
      b.ff1 = get_linear(*m, p + "ff.net.0.proj", 4 * D, D);
      b.ff2 = get_linear(*m, p + "ff.net.2", D, 4 * D);
      b.cff1 = get_linear(*m, p + "ff_context.net.0.proj", 4 * D, D);
      b.cff2 = get_linear(*m, p + "ff_context.net.2", D, 4 * D);
      b.mod_img = add_mod(b.norm1);
      b.mod_ctx = add_mod(b.norm1c);
    }
    m->sgl.resize(c.num_single_layers);
    for (int j = 0; j < c.num_single_layers; ++j) {
      const std::string p = "single_transformer_blocks." + std::to_string(j) + ".";
      SingleBlk& b = m->sgl[j];
      b.norm = get_linear(*m, p + "norm.linear", 3 * D, D);
      b.q = get_linear(*m, p + "attn.to_q", D, D);
      b.k = get_linear(*m, p + "attn.to_k", D, D);
      b.v = get_linear(*m, p + "attn.to_v", D, D);
      b.nq = get_vec(*m, p + "attn.norm_q.weight", hd);
      b.nk = get_vec(*m, p + "attn.norm_k.weight", hd);
      b.mlp = get_linear(*m, p + "proj_mlp", 4 * D, D);
      b.out = get_linear(*m, p + "proj_out", D, 5 * D);
      b.mod = add_mod(b.norm);
    }
    if (c.kind == RT_TRANSFORMER) {
      m->norm_out = get_linear(*m, "norm_out.linear", 2 * D, D);
      m->proj_out = get_linear(*m, "proj_out", c.out_channels, D);
      m->mod_out = add_mod(m->norm_out);
    } else {
      m->cnx_emb = get_linear(*m, "controlnet_x_embedder", D, c.cond_channels);
      for (int i = 0; i < c.num_layers; ++i)
        m->cn_blk.push_back(get_linear(*m, "controlnet_blocks." + std::to_string(i), D, D));
      for (int j = 0; j < c.num_single_layers; ++j)
        m->cn_sgl.push_back(get_linear(*m, "controlnet_single_blocks." + std::to_string(j), D, D));
    }
    m->mod_total = off > 0 ? off : 4;
    m->mod_rows = rows;
    m->n_mod_jobs = (int)jobs.size() - 6;
    m->first_block_jobs = c.num_layers > 0 ? 2 : (c.num_single_layers > 0 ? 1 : 0);
    m->first_block_rows = m->first_block_jobs < m->n_mod_jobs ? prefix[1 + m->first_block_jobs] : rows;
    RT_CHECK_CUDA(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
    RT_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    RT_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    RT_CHECK_CUDA(cudaMalloc(&m->jobs_dev, jobs.size() * sizeof(GemvJob)));
    RT_CHECK_CUDA(cudaMalloc(&m->prefix_dev, prefix.size() * sizeof(int)));
    RT_CHECK_CUDA(cudaMemcpyAsync(m->jobs_dev, jobs.data(), jobs.size() * sizeof(GemvJob), cudaMemcpyHostToDevice,
                                  (cudaStream_t)stream));
    RT_CHECK_CUDA(cudaMemcpyAsync(m->prefix_dev, prefix.data(), prefix.size() * sizeof(int), cudaMemcpyHostToDevice,
                                  (cudaStream_t)stream));
    RT_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // the host vectors die here
    m->finalized = true;
  });
}

int rt_controlnet_set_live(rt_model* m, int live_layers, int live_single_layers) {
  return guarded([&] {
    RT_REQUIRE(m && m->cfg.kind == RT_CONTROLNET, "set_live: not a ControlNet model");
    m->live_layers = live_layers;
    m->live_single = live_single_layers;
  });
}

int rt_model_set_step_invariant_cache(rt_model* m, int mode) {
  return guarded([&] {
    RT_REQUIRE(m, "set_step_invariant_cache: null model");
    RT_REQUIRE(mode == 0 || mode == 1, "set_step_invariant_cache: mode must be 0 (off) or 1 (on)");
    m->inv.mode = mode;
    m->inv.valid = false;  // every call invalidates: the next forward recomputes
  });
}

int64_t rt_model_modulation_table_bytes(const rt_model* m, int steps, int batch) {
  if (!m || !m->finalized || steps < 1 || batch < 1) return -1;
  return (int64_t)modulation_table_layout(*m, (size_t)steps * batch, nullptr);
}

int rt_model_build_modulation_table(rt_model* m, const void* timesteps, const void* guidance, const void* pooled_projections,
                                    int steps, int batch, void* table, int64_t table_bytes, void* stream) {
  return guarded([&] {
    RT_REQUIRE(m && m->finalized, "build_modulation_table: model is null or not finalized");
    RT_REQUIRE(steps >= 1 && batch >= 1 && timesteps && pooled_projections, "build_modulation_table: bad argument");
    RT_REQUIRE(!m->cfg.guidance_embeds || guidance, "build_modulation_table: this model has guidance_embeds");
    build_modulation_table(*m, timesteps, guidance, pooled_projections, steps, batch, (char*)table, (size_t)table_bytes,
                           (cudaStream_t)stream);
  });
}

int rt_model_select_modulation(rt_model* m, int step) {
  return guarded([&] {
    RT_REQUIRE(m, "select_modulation: null model");
    RT_REQUIRE(step == -1 || (step >= 0 && step < m->modtab.steps && m->modtab.mod != nullptr),
               "select_modulation: no table, or the step is outside it");
    m->modtab.sel = step;
  });
}

int rt_model_destroy(rt_model* m) {
  return guarded([&] { delete m; });
}

int64_t rt_model_workspace_bytes(const rt_model* m, int batch, int n_img, int n_txt) {
  if (!m || !m->finalized || batch < 1 || n_img < 1 || n_txt < 1) return -1;
  return (int64_t)carve(*m, batch, n_img, n_txt, nullptr, nullptr);
}

int rt_controlnet_forward(rt_model* m, const rt_forward_args* a, const void* controlnet_cond, int cond_batch,
                          float conditioning_scale, const void* mask, int accumulate, void* block_samples,
                          void* single_block_samples) {
  return guarded([&] {
    RT_REQUIRE(a, "null args");
    RT_REQUIRE(!a->sp || !a->sp->lockstep, "a lock-step group runs through rt_controlnet_forward_lockstep");
    rt_controlnet_call k{*a, controlnet_cond, cond_batch, conditioning_scale, mask, accumulate, block_samples,
                         single_block_samples};
    controlnet_forward_impl(m, {&k});
  });
}

int rt_transformer_forward(rt_model* m, const rt_forward_args* a, const void* const* controlnet_block_samples,
                           int n_block_samples, const void* const* controlnet_single_block_samples,
                           int n_single_block_samples, void* out) {
  return guarded([&] {
    RT_REQUIRE(a, "null args");
    RT_REQUIRE(!a->sp || !a->sp->lockstep, "a lock-step group runs through rt_transformer_forward_lockstep");
    rt_transformer_call k{*a, controlnet_block_samples, n_block_samples, controlnet_single_block_samples,
                          n_single_block_samples, out};
    transformer_forward_impl(m, {&k});
  });
}

int rt_controlnet_forward_lockstep(rt_model* m, int world, const rt_controlnet_call* calls) {
  return guarded([&] {
    RT_REQUIRE(calls && world >= 1 && world <= RT_SP_MAX_RANKS, "lockstep: bad argument");
    std::vector<const rt_controlnet_call*> v;
    std::vector<const rt_forward_args*> args;
    for (int i = 0; i < world; ++i) { v.push_back(calls + i); args.push_back(&calls[i].a); }
    check_lockstep(world, v.size(), args.data());
    controlnet_forward_impl(m, v);
  });
}

int rt_transformer_forward_lockstep(rt_model* m, int world, const rt_transformer_call* calls) {
  return guarded([&] {
    RT_REQUIRE(calls && world >= 1 && world <= RT_SP_MAX_RANKS, "lockstep: bad argument");
    std::vector<const rt_transformer_call*> v;
    std::vector<const rt_forward_args*> args;
    for (int i = 0; i < world; ++i) { v.push_back(calls + i); args.push_back(&calls[i].a); }
    check_lockstep(world, v.size(), args.data());
    transformer_forward_impl(m, v);
  });
}

}  // extern "C"
