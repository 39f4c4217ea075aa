// jf_engine.cu — host runtime behind the C ABI (include/jacobi_b200.h): device memory, weight re-tiling, plan upload,
// cooperative launch of the persistent kernel.  Mirrors what the reference's ModelRunner owns per rank
// (inference_engine/engine/model_runner.py:155-291 init, :523-622 KV cache, :1134-1418 Jacobi forward step).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "jf_kernel.cuh"
#include "../../include/jacobi_b200.h"

using namespace jf;

static_assert(sizeof(jf_policy_knobs) == sizeof(PolicyKnobs), "jf_policy_knobs must mirror jf::PolicyKnobs");
static_assert(sizeof(DevCtx) % 16 == 0 && sizeof(DevCtx) <= 5120, "DevCtx is staged in shared memory");

static thread_local std::string g_create_error;

struct jf_engine {
  int device = 0;
  int sm_count = 0;
  int num_ctas = 0;                  // CTAs of the persistent kernel (= sm_count unless jf_runtime_cfg.num_ctas asks for fewer)
  int smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ModelDims md{};
  jf_runtime_cfg rt{};
  DevCtx hctx{};
  DevCtx* dctx = nullptr;
  std::vector<void*> allocs;
  std::vector<void*> ipc_maps;       // peer arenas opened with cudaIpcOpenMemHandle
  std::string err;
  int64_t bytes_weights = 0, bytes_kv = 0, bytes_act = 0;
  // arenas: one cudaMalloc per class (weights / KV cache / activations+state) so each class sits in few, densely
  // used 2 MB pages; mode 0 = direct cudaMalloc, 1 = measuring, 2 = carving
  int arena_mode = 0;
  uint8_t* arena_base[3] = {nullptr, nullptr, nullptr};
  size_t arena_off[3] = {0, 0, 0};
  int64_t launches = 0;
  int* dbg_host = nullptr;
  // weight bookkeeping
  std::vector<uint32_t> seen_layer;  // W_* / B_* bits per layer (jf_finalize_weights refuses a partial checkpoint)
  int seen_global = 0;
  __nv_bfloat16* stage = nullptr;    // staging for row-major -> tile re-layout
  size_t stage_elems = 0;
  float* bias_stage = nullptr;
  int max_splits = 1;
  bool finalized = false;
  // host mirrors
  std::vector<SeqState> h_states;
  // per-call draft sources of jf_jacobi_run (first-call tokens, MT19937 states, pad streams): grown on demand, kept
  int* run_scratch[3] = {nullptr, nullptr, nullptr};
  size_t run_scratch_words[3] = {0, 0, 0};
};

#define JF_CUDA(h, call)                                                                         \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
      return JF_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

template <typename T>
static int dev_alloc(jf_engine* h, T** out, size_t count, int64_t* tally, bool zero = true) {
  const size_t bytes = ((count * sizeof(T) > 0 ? count * sizeof(T) : 16) + 1023) & ~(size_t)1023;
  if (h->arena_mode != 0) {
    const int a = (tally == &h->bytes_weights) ? 0 : (tally == &h->bytes_kv) ? 1 : 2;
    if (h->arena_mode == 2) {
      *out = reinterpret_cast<T*>(h->arena_base[a] + h->arena_off[a]);
      if (tally) *tally += (int64_t)bytes;
    } else {
      *out = nullptr;
    }
    h->arena_off[a] += bytes;
    return JF_OK;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return JF_ERR_CUDA; }
  if (zero) cudaMemsetAsync(p, 0, bytes, h->stream);
  h->allocs.push_back(p);
  if (tally) *tally += (int64_t)bytes;
  *out = reinterpret_cast<T*>(p);
  return JF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// small device helpers: dtype conversion, weight re-tiling, rope tables
// ---------------------------------------------------------------------------------------------------------------
__global__ void convert_to_bf16_kernel(const void* src, __nv_bfloat16* dst, size_t n, int dtype) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v;
    if (dtype == JF_DTYPE_F32) v = reinterpret_cast<const float*>(src)[i];
    else if (dtype == JF_DTYPE_F16) v = __half2float(reinterpret_cast<const __half*>(src)[i]);
    else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
    dst[i] = __float2bfloat16_rn(v);
  }
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* src, float* dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}
// rows [0,R) of a row-major [R][K] bf16 matrix -> rows [dest_row0, dest_row0+R) of the tiled sw128 layout
__global__ void pack_rows_into_tiles_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                            int R, int K, int k_chunks, int dest_row0) {
  const int kv = K >> 3;
  const long long total = (long long)R * kv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / kv), k8 = (int)(i - (long long)n * kv);
    const int k = k8 << 3;
    const int dr = dest_row0 + n;
    const int tile = dr >> 7, r = dr & 127;
    const int kc = k >> 6, chunk_log = (k & 63) >> 3;
    const int phys = chunk_log ^ (r & 7);
    const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)n * K + k);
    reinterpret_cast<uint4*>(dst)[(((size_t)tile * k_chunks + kc) * 128 + r) * 8 + phys] = v;
  }
}
__global__ void rope_table_kernel(__nv_bfloat16* cosb, __nv_bfloat16* sinb, int max_pos, int half, float theta, int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * half) return;
  const int pos = i / half, e = i - pos * half;
  // transformers Qwen2RotaryEmbedding: inv_freq = 1 / theta^(2e/d) in fp32; freqs = pos * inv_freq in fp32; cos/sin fp32
  const float inv_freq = 1.0f / powf(theta, (float)(2 * e) / (float)d);
  const float f = (float)pos * inv_freq;
  cosb[i] = __float2bfloat16_rn(cosf(f));
  sinb[i] = __float2bfloat16_rn(sinf(f));
}

// ---------------------------------------------------------------------------------------------------------------
// creation
// ---------------------------------------------------------------------------------------------------------------
static int pick_splits(int n_tiles, int k_chunks, int sms) {
  int s = sms / (n_tiles > 0 ? n_tiles : 1);
  const int by_k = k_chunks / 4;
  if (s > by_k) s = by_k;
  if (s < 1) s = 1;
  if (s > 8) s = 8;
  return s;
}

extern "C" const char* jf_last_error(jf_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int jf_create(const jf_model_cfg* m, const jf_runtime_cfg* r, jf_handle* out) {
  if (!m || !r || !out) { g_create_error = "null argument"; return JF_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device: libjacobi_b200 has no CPU path";
    return JF_ERR_CUDA;
  }
  jf_engine* h = new jf_engine();
  auto fail = [&](int code, const std::string& msg) { g_create_error = msg; for (void* p : h->allocs) cudaFree(p); delete h; return code; };
  h->device = r->device;
  if (cudaSetDevice(h->device) != cudaSuccess) return fail(JF_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, h->device);
  if (prop.major != 10) return fail(JF_ERR_UNSUPPORTED, "libjacobi_b200 is built for sm_100a (B200) only");
  h->sm_count = prop.multiProcessorCount;
  cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
  if (!coop) return fail(JF_ERR_UNSUPPORTED, "device lacks cooperative launch");

  ModelDims& md = h->md;
  const int tp = r->tp_size > 1 ? r->tp_size : 1;
  const int tpr = r->tp_size > 1 ? r->tp_rank : 0;
  if (tp != 1 && tp != 2 && tp != 4 && tp != 8) { g_create_error = "tp_size must be 1, 2, 4 or 8"; delete h; return JF_ERR_UNSUPPORTED; }
  if (tpr < 0 || tpr >= tp) { g_create_error = "tp_rank outside [0, tp_size)"; delete h; return JF_ERR_INVALID; }
  // tp_attn: ranks the attention block is sharded over.  0 / 1 = replicated on every rank (MLP-only tensor parallelism:
  // one exchange per transformer block, no head-divisibility constraint); tp_size = sharded like the reference.
  const int tpa = (tp > 1 && r->tp_attn == tp) ? tp : 1;
  if (r->tp_attn != 0 && r->tp_attn != 1 && r->tp_attn != tp) { g_create_error = "tp_attn must be 0, 1 or tp_size"; delete h; return JF_ERR_INVALID; }
  if (m->num_attention_heads % tpa || m->num_key_value_heads % tpa) {
    g_create_error = "heads / kv heads must be divisible by tp_attn (reference: inference_engine/models/qwen3.py:31,34); "
                     "use tp_attn = 1 (attention replicated) otherwise";
    delete h; return JF_ERR_INVALID;
  }
  if (m->intermediate_size % tp || m->vocab_size % tp) {
    g_create_error = "intermediate_size / vocab_size must be divisible by tp_size";
    delete h; return JF_ERR_INVALID;
  }
  h->num_ctas = (r->num_ctas > 0 && r->num_ctas < h->sm_count) ? r->num_ctas : h->sm_count;
  md.h = m->hidden_size; md.layers = m->num_hidden_layers; md.n_q = m->num_attention_heads / tpa;
  md.n_kv = m->num_key_value_heads / tpa; md.d = m->head_dim; md.inter = m->intermediate_size / tp;
  md.vocab = m->vocab_size / tp; md.vocab_embed = m->vocab_size; md.vocab_offset = tpr * (m->vocab_size / tp);
  md.q_dim = md.n_q * md.d; md.kv_dim = md.n_kv * md.d; md.qkv_dim = md.q_dim + 2 * md.kv_dim;
  md.n_rep = md.n_kv > 0 ? md.n_q / md.n_kv : 0;
  md.tie = m->tie_word_embeddings; md.eps = m->rms_norm_eps; md.rope_theta = m->rope_theta;
  md.max_pos = m->max_position_embeddings;
  if (md.layers < 1 || md.layers > kMaxLayers) return fail(JF_ERR_INVALID, "num_hidden_layers outside [1, 64]");
  if (md.d != 64 && md.d != 128) return fail(JF_ERR_UNSUPPORTED, "head_dim must be 64 or 128");
  if (md.h % 64 || md.h > 8192) return fail(JF_ERR_UNSUPPORTED, "hidden_size must be a multiple of 64 and <= 8192");
  if (md.inter % 64) return fail(JF_ERR_UNSUPPORTED, "intermediate_size (per rank) must be a multiple of 64");
  if (md.n_kv < 1 || md.n_q % md.n_kv) return fail(JF_ERR_INVALID, "num_attention_heads must be a multiple of num_key_value_heads");
  if (md.vocab < 1 || md.max_pos < 1) return fail(JF_ERR_INVALID, "vocab_size / max_position_embeddings must be positive");

  h->rt = *r;
  jf_runtime_cfg& rt = h->rt;
  if (rt.max_rows < 16) rt.max_rows = 16;
  rt.max_rows = (rt.max_rows + 15) & ~15;
  if (rt.max_seqs < 1) rt.max_seqs = 1;
  if (rt.attn_chunk <= 0) rt.attn_chunk = 64;
  if (rt.attn_chunk % kAttnSub) return fail(JF_ERR_INVALID, "attn_chunk must be a multiple of 64");
  if (rt.num_kv_pages < 1 || rt.max_pages_per_seq < 1) return fail(JF_ERR_INVALID, "num_kv_pages / max_pages_per_seq must be positive");
  if (rt.gen_capacity < 1) rt.gen_capacity = 4096;
  if (rt.max_calls < 4) rt.max_calls = 4;
  if (rt.attn_slots <= 0) {
    // enough for max_rows tokens each seeing max_pages_per_seq pages of keys, capped to 2 GiB of fp32 partials
    long long want = (long long)rt.max_rows * (((long long)rt.max_pages_per_seq * kPage) / rt.attn_chunk + 1);
    const long long cap = (2ll << 30) / ((long long)md.n_q * md.d * 4);
    rt.attn_slots = (int)(want < cap ? want : cap);
  }

  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(JF_ERR_CUDA, "stream");
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);

  DevCtx& c = h->hctx;
  memset(&c, 0, sizeof(c));
  c.md = md;
  const int kc_h = md.h / 64, kc_q = md.q_dim / 64, kc_i = md.inter / 64;
  const int nt_qkv = (md.qkv_dim + 127) / 128, nt_h = (md.h + 127) / 128, nt_i = (md.inter + 127) / 128;
  const int nt_v = (md.vocab + 127) / 128;
  c.splits_qkv = pick_splits(nt_qkv, kc_h, h->num_ctas);
  c.splits_o = pick_splits(nt_h, kc_q, h->num_ctas);
  c.splits_down = pick_splits(nt_h, kc_i, h->num_ctas);
  {
    const int nt_i = (md.inter + 127) / 128;
    c.splits_gu = std::min(tp > 1 ? 8 : 4, pick_splits(nt_i, kc_h, h->num_ctas));
    if (nt_i * c.splits_gu > h->num_ctas) c.splits_gu = 1;          // the SwiGLU tail needs one unit per CTA at most
    if (const char* e = getenv("JF_GU_SPLITS")) c.splits_gu = std::max(1, std::min(c.splits_gu, atoi(e)));
    if (const char* e = getenv("JF_DOWN_SPLITS")) c.splits_down = std::max(1, std::min(c.splits_down, atoi(e)));
  }
  c.tp_rank = tpr; c.tp_size = tp; c.tp_attn = tpa;
  c.fuse_rows = std::min(rt.max_rows, tp > 1 ? kFuseRowsTP : kFuseRows);
  h->max_splits = std::max(c.splits_qkv, std::max(c.splits_o, c.splits_down));

  int rc = JF_OK;
#define ALLOC(ptr, count, tally) if ((rc = dev_alloc(h, &(ptr), (size_t)(count), &(tally))) != JF_OK) return rc
  auto layout = [&]() -> int {
  // ---- weights
  __nv_bfloat16 *embed, *lm_head, *final_norm, *cosb, *sinb;
  ALLOC(embed, (size_t)md.vocab_embed * md.h, h->bytes_weights);
  ALLOC(lm_head, (size_t)nt_v * kc_h * 8192, h->bytes_weights);
  ALLOC(final_norm, md.h, h->bytes_weights);
  ALLOC(cosb, (size_t)md.max_pos * (md.d / 2), h->bytes_weights);
  ALLOC(sinb, (size_t)md.max_pos * (md.d / 2), h->bytes_weights);
  c.embed = embed; c.lm_head = lm_head; c.final_norm = final_norm; c.rope_cos = cosb; c.rope_sin = sinb;
  for (int l = 0; l < md.layers; ++l) {
    __nv_bfloat16 *wqkv, *wo, *wgu, *wd, *ln1, *ln2; float* bqkv;
    ALLOC(wqkv, (size_t)nt_qkv * kc_h * 8192, h->bytes_weights);
    ALLOC(bqkv, nt_qkv * 128, h->bytes_weights);
    ALLOC(wo, (size_t)nt_h * kc_q * 8192, h->bytes_weights);
    ALLOC(wgu, (size_t)2 * nt_i * kc_h * 8192, h->bytes_weights);
    ALLOC(wd, (size_t)nt_h * kc_i * 8192, h->bytes_weights);
    ALLOC(ln1, md.h, h->bytes_weights);
    ALLOC(ln2, md.h, h->bytes_weights);
    c.layer[l] = LayerW{wqkv, bqkv, wo, wgu, wd, ln1, ln2};
  }
  h->seen_layer.assign(md.layers, 0);
  {
    const int n = md.max_pos * (md.d / 2);
    if (h->arena_mode == 2)
      rope_table_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(cosb, sinb, md.max_pos, md.d / 2, md.rope_theta, md.d);
  }
  // ---- KV
  c.num_pages = rt.num_kv_pages; c.max_pages_per_seq = rt.max_pages_per_seq;
  ALLOC(c.kv_cache, (size_t)md.layers * 2 * rt.num_kv_pages * md.n_kv * kPage * md.d, h->bytes_kv);
  ALLOC(c.block_tables, (size_t)rt.max_seqs * rt.max_pages_per_seq, h->bytes_act);
  // ---- activations
  c.T_cap = rt.max_rows; c.xrows_cap = rt.max_rows; c.S_cap = rt.max_seqs; c.G_cap = rt.max_seqs * kRowsMax;
  if (c.G_cap < 64) c.G_cap = 64;
  ALLOC(c.hidden, (size_t)c.T_cap * md.h, h->bytes_act);
  ALLOC(c.xn, (size_t)kc_h * c.xrows_cap * 64, h->bytes_act);
  ALLOC(c.xa, (size_t)kc_q * c.xrows_cap * 64, h->bytes_act);
  ALLOC(c.qbuf, (size_t)c.T_cap * md.q_dim, h->bytes_act);
  c.attn_slots_cap = rt.attn_slots; c.attn_chunk = rt.attn_chunk;
  c.attn_hsplit = 0;                                  // per forward (pick_attn_hsplit); JF_ATTN_HSPLIT fixes it
  if (const char* eh = getenv("JF_ATTN_HSPLIT")) c.attn_hsplit = std::max(0, std::min(md.n_rep, atoi(eh)));
  ALLOC(c.argmax_keys, (size_t)2 * c.T_cap, h->bytes_act);
  ALLOC(c.xflag, kMaxTP, h->bytes_act);
  ALLOC(c.row_ss, (size_t)c.T_cap * kMaxHTiles, h->bytes_act);
  ALLOC(c.hdr, 1, h->bytes_act);
  ALLOC(c.tok, c.T_cap, h->bytes_act);
  ALLOC(c.row_group, c.T_cap, h->bytes_act);
  ALLOC(c.row_pos, c.T_cap, h->bytes_act);
  ALLOC(c.row_attn, c.T_cap, h->bytes_act);
  ALLOC(c.groups, c.G_cap, h->bytes_act);
  ALLOC(c.commits, c.T_cap, h->bytes_act);
  ALLOC(c.greedy, c.T_cap, h->bytes_act);
  c.logits = nullptr;
  {
    const char* e1 = getenv("JF_DEBUG_SKIP");
    c.debug_skip = e1 ? atoi(e1) : 0;
    const char* e3 = getenv("JF_ATTN_PREFETCH");
    c.attn_prefetch = e3 ? atoi(e3) : 1;
    const char* e4 = getenv("JF_W_EVICT_FIRST");
    c.w_evict_first = e4 ? atoi(e4) : 1;
    const char* e6 = getenv("JF_FUSE_ROPE");
    c.fuse_rope = e6 ? atoi(e6) : 1;
    const char* e8 = getenv("JF_DATAFLOW_GD");
    c.dataflow_gd = e8 ? atoi(e8) : 1;
    // the hand-over announces ONE gate/up tile per CTA (tile index == CTA index): models whose intermediate size needs more
    // tiles than there are CTAs (> 148 x 128 features on one GPU) keep the grid barrier between gate/up and down
    if (nt_i > h->num_ctas) c.dataflow_gd = 0;
    if (getenv("JF_DEBUG_HOSTBUF") && h->arena_mode == 2) {
      int* hb = nullptr;
      cudaHostAlloc(&hb, 4096 * sizeof(int), cudaHostAllocMapped);
      memset(hb, 0, 4096 * sizeof(int));
      int* db = nullptr;
      cudaHostGetDevicePointer(&db, hb, 0);
      c.dbg = db;
      h->dbg_host = hb;
    }
    const char* e10 = getenv("JF_ATTN_TC");
    c.attn_tc = e10 ? atoi(e10) : 0;
    const char* e5 = getenv("JF_FUSE_NORM");
    c.fuse_norm = e5 ? atoi(e5) : 1;
    const char* e15 = getenv("JF_EARLY_FENCE");
    c.early_fence = e15 ? atoi(e15) : 1;
    const char* e13 = getenv("JF_X_FENCE");
    c.x_fence = e13 ? atoi(e13) : 0;
    const char* e16 = getenv("JF_EPI_BULK");
    c.epi_bulk = e16 ? atoi(e16) : 1;
    const char* e12 = getenv("JF_SWIGLU_BULK");
    c.swiglu_bulk = e12 ? atoi(e12) : 1;
  }
  ALLOC(c.ng_prop, (size_t)2 * c.T_cap, h->bytes_act);
  ALLOC(c.ng_gumbel, (size_t)2 * c.T_cap, h->bytes_act);
  ALLOC(c.ng_px, c.T_cap, h->bytes_act);
  ALLOC(c.ng_bonus, c.T_cap, h->bytes_act);
  c.ng_parts = std::min(h->num_ctas, nt_v);
  ALLOC(c.ng_ms, (size_t)2 * c.ng_parts * c.T_cap, h->bytes_act);
  ALLOC(c.seq_state, c.S_cap, h->bytes_act);
  ALLOC(c.seq_slot, c.S_cap, h->bytes_act);
  ALLOC(c.knobs, c.S_cap, h->bytes_act);
  ALLOC(c.draft_src, c.S_cap, h->bytes_act);
  c.gen_cap = rt.gen_capacity; c.iters_cap = rt.max_calls;
  ALLOC(c.gen, (size_t)c.S_cap * c.gen_cap, h->bytes_act);
  ALLOC(c.iters_log, (size_t)c.S_cap * c.iters_cap, h->bytes_act);
  c.trace_cap = rt.trace_capacity > 0 ? rt.trace_capacity : 0;
  if (c.trace_cap > 0) { ALLOC(c.trace, c.trace_cap, h->bytes_act); } else c.trace = nullptr;
  ALLOC(c.trace_len, 4, h->bytes_act);
  ALLOC(c.bar_counter, 4 + 3 * kMaxTilesPerKind, h->bytes_act);
  c.tile_ctr = c.bar_counter + 4;
  ALLOC(h->dctx, 1, h->bytes_act);
  // large, sparsely touched buffers last: the small hot ones above share the arena's first pages
  c.part_splits = h->max_splits; c.part_n = md.qkv_dim;
  ALLOC(c.part, (size_t)c.T_cap * c.part_splits * c.part_n, h->bytes_act);
  ALLOC(c.part_rs[0], (size_t)c.T_cap * c.part_splits * md.h, h->bytes_act);
  ALLOC(c.part_rs[1], (size_t)c.T_cap * c.part_splits * md.h, h->bytes_act);
  c.part_rs[2] = c.part_rs[0];
  if (tp > 1 && tpa == 1) ALLOC(c.part_rs[2], (size_t)c.T_cap * c.part_splits * md.h, h->bytes_act);
  if (tp > 1) {
    unsigned long long *pb = nullptr, *sb = nullptr;
    c.push_rows_cap = (c.fuse_rows + tp - 1) / tp + 8;
    ALLOC(pb, (size_t)c.push_rows_cap * tp * md.h, h->bytes_act);
    ALLOC(sb, (size_t)c.fuse_rows * md.h, h->bytes_act);              // [rows][h / 8][8] words
    c.push_buf[tpr] = pb; c.stage_buf[tpr] = sb;
  }
  ALLOC(c.tpart, (size_t)kFuseRopeRows * c.part_splits * c.part_n, h->bytes_act);
  c.tgu = nullptr;
  if (c.splits_gu > 1) ALLOC(c.tgu, (size_t)c.T_cap * c.splits_gu * 2 * md.inter, h->bytes_act);
  ALLOC(c.tpart_rs[0], (size_t)c.fuse_rows * c.part_splits * md.h, h->bytes_act);
  ALLOC(c.tpart_rs[1], (size_t)c.fuse_rows * c.part_splits * md.h, h->bytes_act);
  ALLOC(c.attn_ml, (size_t)c.attn_slots_cap * md.n_q * 2, h->bytes_act);
  ALLOC(c.attn_o, (size_t)c.attn_slots_cap * md.n_q * md.d, h->bytes_act);
  ALLOC(c.xact, (size_t)kc_i * c.xrows_cap * 64, h->bytes_act);
  ALLOC(c.kv_scratch, (size_t)md.layers * 2 * rt.max_rows * md.n_kv * md.d, h->bytes_act);
  // staging: the largest single HF tensor (embedding / lm_head or gate/up/down)
  h->stage_elems = std::max((size_t)md.vocab_embed * md.h, (size_t)md.inter * md.h);
    return JF_OK;
  };
  h->arena_mode = 1;
  if ((rc = layout()) != JF_OK) return fail(rc, h->err);
  for (int a2 = 0; a2 < 3; ++a2) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, h->arena_off[a2] + 4096);
    if (e != cudaSuccess) return fail(JF_ERR_CUDA, std::string("cudaMalloc (arena): ") + cudaGetErrorString(e));
    cudaMemsetAsync(p, 0, h->arena_off[a2] + 4096, h->stream);
    h->allocs.push_back(p);
    h->arena_base[a2] = reinterpret_cast<uint8_t*>(p);
    h->arena_off[a2] = 0;
  }
  h->arena_mode = 2;
  if ((rc = layout()) != JF_OK) return fail(rc, h->err);
  h->arena_mode = 0;
  c.arena_self = h->arena_base[2];
  for (int r2 = 0; r2 < kMaxTP; ++r2) {            // until jf_tp_import: every 'peer' is this rank itself
    c.peer_arena[r2] = h->arena_base[2];
    for (int b = 0; b < 3; ++b) c.peer_part_rs[r2][b] = c.part_rs[b];
    c.push_buf[r2] = c.push_buf[tpr]; c.stage_buf[r2] = c.stage_buf[tpr];
    c.peer_keys[r2] = c.argmax_keys; c.peer_xflag[r2] = c.xflag;
  }
  if ((rc = dev_alloc(h, &h->stage, h->stage_elems, (int64_t*)nullptr)) != JF_OK) return fail(rc, h->err);
  if ((rc = dev_alloc(h, &h->bias_stage, (size_t)md.qkv_dim + 128, (int64_t*)nullptr)) != JF_OK) return fail(rc, h->err);
#undef ALLOC
  if (cudaFuncSetAttribute(jf_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin) != cudaSuccess)
    return fail(JF_ERR_CUDA, "cannot opt in to the full shared memory carve-out");
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jf_persistent_kernel, kNumThreads, h->smem_optin);
  if (occ < 1) return fail(JF_ERR_CUDA, "persistent kernel does not fit on an SM");
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(JF_ERR_CUDA, "initialisation failed");
  *out = h;
  return JF_OK;
}

extern "C" int jf_destroy(jf_handle h) {
  if (!h) return JF_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->ipc_maps) cudaIpcCloseMemHandle(p);
  for (void* p : h->allocs) cudaFree(p);
  for (int* p : h->run_scratch) if (p) cudaFree(p);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
  return JF_OK;
}

extern "C" int jf_device_info(jf_handle h, int32_t* sm_count, int64_t* bw, int64_t* bkv, int64_t* bact) {
  if (!h) return JF_ERR_INVALID;
  if (sm_count) *sm_count = h->num_ctas;
  if (bw) *bw = h->bytes_weights;
  if (bkv) *bkv = h->bytes_kv;
  if (bact) *bact = h->bytes_act;
  return JF_OK;
}

extern "C" int64_t jf_launch_count(jf_handle h) { return h ? h->launches : 0; }

extern "C" int jf_tp_export(jf_handle h, void* handle64) {
  if (!h || !handle64) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t ipc;
  JF_CUDA(h, cudaIpcGetMemHandle(&ipc, h->arena_base[2]));
  memcpy(handle64, &ipc, 64);
  return JF_OK;
}

// every rank lays its activation arena out identically, so a peer's buffer sits at the same offset from its arena base
static void map_peer(jf_engine* h, int r, uint8_t* base) {
  DevCtx& c = h->hctx;
  const uint8_t* mine = h->arena_base[2];
  auto at = [&](const void* own) { return base + (reinterpret_cast<const uint8_t*>(own) - mine); };
  c.peer_arena[r] = base;
  for (int b = 0; b < 3; ++b) c.peer_part_rs[r][b] = reinterpret_cast<const float*>(at(c.part_rs[b]));
  c.push_buf[r] = reinterpret_cast<unsigned long long*>(at(c.push_buf[c.tp_rank]));
  c.stage_buf[r] = reinterpret_cast<unsigned long long*>(at(c.stage_buf[c.tp_rank]));
  c.peer_keys[r] = reinterpret_cast<const unsigned long long*>(at(c.argmax_keys));
  c.peer_xflag[r] = reinterpret_cast<unsigned long long*>(at(c.xflag));
}

extern "C" int jf_tp_import(jf_handle h, const void* handles) {
  if (!h || !handles) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  DevCtx& c = h->hctx;
  for (int r = 0; r < c.tp_size; ++r) {
    if (r == c.tp_rank) continue;
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, reinterpret_cast<const uint8_t*>(handles) + (size_t)r * 64, 64);
    void* p = nullptr;
    JF_CUDA(h, cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_maps.push_back(p);
    map_peer(h, r, reinterpret_cast<uint8_t*>(p));
  }
  if (h->finalized) JF_CUDA(h, cudaMemcpy(h->dctx, &h->hctx, sizeof(DevCtx), cudaMemcpyHostToDevice));
  return JF_OK;
}

extern "C" int jf_tp_attach_local(jf_handle h, const jf_handle* peers) {
  if (!h || !peers) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  DevCtx& c = h->hctx;
  for (int r = 0; r < c.tp_size; ++r) {
    if (r == c.tp_rank) continue;
    const jf_engine* p = peers[r];
    if (!p || p->hctx.tp_size != c.tp_size || p->hctx.tp_rank != r || p->arena_off[2] != h->arena_off[2]) {
      h->err = "jf_tp_attach_local: peer " + std::to_string(r) + " is not a rank of the same tensor-parallel group";
      return JF_ERR_INVALID;
    }
    if (p->device != h->device) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, h->device, p->device);
      if (!can) { h->err = "jf_tp_attach_local: no peer access between the ranks' devices"; return JF_ERR_UNSUPPORTED; }
      cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->err = "cudaDeviceEnablePeerAccess failed"; return JF_ERR_CUDA; }
      cudaGetLastError();
    }
    map_peer(h, r, p->arena_base[2]);
  }
  if (h->finalized) JF_CUDA(h, cudaMemcpy(h->dctx, &h->hctx, sizeof(DevCtx), cudaMemcpyHostToDevice));
  return JF_OK;
}

extern "C" int jf_profile_enable(jf_handle h, int32_t forward_index) {
  if (!h) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  DevCtx& c = h->hctx;
  if (forward_index >= 0 && c.prof == nullptr) {
    c.prof_cap = 4096;
    int rc = dev_alloc(h, &c.prof, (size_t)c.prof_cap, &h->bytes_act);
    if (rc != JF_OK) return rc;
    rc = dev_alloc(h, &c.prof_count, 4, &h->bytes_act);
    if (rc != JF_OK) return rc;
    if (getenv("JF_PROF_MARKS")) {                       // fine-grained marks only on request: they perturb the stages
      rc = dev_alloc(h, &c.prof_marks, 8192, &h->bytes_act);
      if (rc != JF_OK) return rc;
      rc = dev_alloc(h, &c.prof_marks_n, 4, &h->bytes_act);
      if (rc != JF_OK) return rc;
    }
  }
  c.prof_forward = forward_index;
  long long* keep = c.prof;
  if (forward_index < 0) c.prof = nullptr;
  if (h->finalized) JF_CUDA(h, cudaMemcpy(h->dctx, &h->hctx, sizeof(DevCtx), cudaMemcpyHostToDevice));
  c.prof = keep;
  return JF_OK;
}

extern "C" int jf_profile_read(jf_handle h, int64_t* stamps, int32_t capacity, int32_t* n_stamps, int32_t* sm_clock_khz) {
  if (!h || !stamps || !n_stamps) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  DevCtx& c = h->hctx;
  *n_stamps = 0;
  if (c.prof == nullptr) return JF_OK;
  int n = 0;
  JF_CUDA(h, cudaMemcpy(&n, c.prof_count, 4, cudaMemcpyDeviceToHost));
  if (n > capacity) n = capacity;
  if (n > 0) JF_CUDA(h, cudaMemcpy(stamps, c.prof, (size_t)n * 8, cudaMemcpyDeviceToHost));
  *n_stamps = n;
  if (const char* path = c.prof_marks_n ? getenv("JF_PROF_MARKS") : nullptr) {          // developer aid: dump the fine-grained marks as text
    int nm = 0;
    cudaMemcpy(&nm, c.prof_marks_n, 4, cudaMemcpyDeviceToHost);
    std::vector<long long> mk(nm > 0 ? nm : 1);
    if (nm > 0) cudaMemcpy(mk.data(), c.prof_marks, (size_t)nm * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(path, "w")) {
      for (int i = 0; i + 1 < nm; i += 2) fprintf(f, "%lld %lld\n", mk[i], mk[i + 1]);
      fclose(f);
    }
  }
  if (sm_clock_khz) cudaDeviceGetAttribute(sm_clock_khz, cudaDevAttrClockRate, h->device);
  return JF_OK;
}

extern "C" int jf_kv_info(jf_handle h, int32_t* num_pages, int32_t* page_size, int64_t* bytes_per_page, void** base) {
  if (!h) return JF_ERR_INVALID;
  if (num_pages) *num_pages = h->hctx.num_pages;
  if (page_size) *page_size = kPage;
  if (bytes_per_page) *bytes_per_page = (int64_t)h->md.layers * 2 * h->md.n_kv * kPage * h->md.d * 2;
  if (base) *base = h->hctx.kv_cache;
  return JF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------------------------
enum { W_Q = 1, W_K = 2, W_V = 4, W_O = 8, W_G = 16, W_U = 32, W_D = 64, W_LN1 = 128, W_LN2 = 256, B_Q = 512, B_K = 1024, B_V = 2048 };
enum { G_EMBED = 1, G_HEAD = 2, G_NORM = 4 };

static int stage_tensor(jf_engine* h, const void* data, size_t n, int dtype) {
  if (n > h->stage_elems) { h->err = "tensor larger than the staging buffer"; return JF_ERR_INVALID; }
  if (dtype == JF_DTYPE_BF16) {
    JF_CUDA(h, cudaMemcpyAsync(h->stage, data, n * 2, cudaMemcpyDefault, h->stream));
    return JF_OK;
  }
  // convert through a temporary device copy of the source
  void* tmp = nullptr;
  const size_t esz = dtype == JF_DTYPE_F32 ? 4 : 2;
  JF_CUDA(h, cudaMalloc(&tmp, n * esz));
  cudaError_t e = cudaMemcpyAsync(tmp, data, n * esz, cudaMemcpyDefault, h->stream);
  if (e == cudaSuccess) {
    convert_to_bf16_kernel<<<1024, 256, 0, h->stream>>>(tmp, h->stage, n, dtype);
    ++h->launches;
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(tmp);
  if (e != cudaSuccess) { h->err = std::string("dtype conversion: ") + cudaGetErrorString(e); return JF_ERR_CUDA; }
  return JF_OK;
}

static int pack_staged_from(jf_engine* h, __nv_bfloat16* dst, int R, int K, int dest_row0, size_t stage_off) {
  pack_rows_into_tiles_kernel<<<2048, 256, 0, h->stream>>>(h->stage + stage_off, dst, R, K, K / 64, dest_row0);
  ++h->launches;
  JF_CUDA(h, cudaStreamSynchronize(h->stream));
  return JF_OK;
}
static int pack_staged(jf_engine* h, __nv_bfloat16* dst, int R, int K, int dest_row0) {
  pack_rows_into_tiles_kernel<<<2048, 256, 0, h->stream>>>(h->stage, dst, R, K, K / 64, dest_row0);
  ++h->launches;
  JF_CUDA(h, cudaStreamSynchronize(h->stream));   // the staging buffer is reused by the next tensor
  return JF_OK;
}

extern "C" int jf_load_tensor(jf_handle h, const char* name, const void* data, const int64_t* shape, int32_t ndim,
                              int32_t dtype) {
  if (!h || !name || !data || !shape) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  const ModelDims& md = h->md;
  DevCtx& c = h->hctx;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
  auto want = [&](int64_t r, int64_t k) -> bool {
    if (ndim == 2 && shape[0] == r && shape[1] == k) return true;
    h->err = std::string(name) + ": unexpected shape";
    return false;
  };
  auto want1 = [&](int64_t r) -> bool {
    if (ndim == 1 && shape[0] == r) return true;
    h->err = std::string(name) + ": unexpected shape";
    return false;
  };
  int rc;
  std::string s(name);
  if (s == "model.embed_tokens.weight") {
    if (!want(md.vocab_embed, md.h)) return JF_ERR_INVALID;
    if ((rc = stage_tensor(h, data, n, dtype)) != JF_OK) return rc;
    JF_CUDA(h, cudaMemcpyAsync(const_cast<__nv_bfloat16*>(c.embed), h->stage, n * 2, cudaMemcpyDeviceToDevice, h->stream));
    h->seen_global |= G_EMBED;
    if (md.tie) {
      if ((rc = pack_staged_from(h, const_cast<__nv_bfloat16*>(c.lm_head), md.vocab, md.h, 0, (size_t)md.vocab_offset * md.h)) != JF_OK) return rc;
      h->seen_global |= G_HEAD;
    }
    JF_CUDA(h, cudaStreamSynchronize(h->stream));
    return JF_OK;
  }
  if (s == "lm_head.weight") {
    if (!want(md.vocab, md.h)) return JF_ERR_INVALID;
    if (md.tie) return JF_OK;                     // tied checkpoints may still list it; the embedding wins
    if ((rc = stage_tensor(h, data, n, dtype)) != JF_OK) return rc;
    h->seen_global |= G_HEAD;
    return pack_staged(h, const_cast<__nv_bfloat16*>(c.lm_head), md.vocab, md.h, 0);
  }
  if (s == "model.norm.weight") {
    if (!want1(md.h)) return JF_ERR_INVALID;
    if ((rc = stage_tensor(h, data, n, dtype)) != JF_OK) return rc;
    JF_CUDA(h, cudaMemcpyAsync(const_cast<__nv_bfloat16*>(c.final_norm), h->stage, n * 2, cudaMemcpyDeviceToDevice, h->stream));
    JF_CUDA(h, cudaStreamSynchronize(h->stream));
    h->seen_global |= G_NORM;
    return JF_OK;
  }
  const std::string pre = "model.layers.";
  if (s.compare(0, pre.size(), pre) != 0) { h->err = "unknown tensor name: " + s; return JF_ERR_INVALID; }
  size_t dot = s.find('.', pre.size());
  if (dot == std::string::npos) { h->err = "unknown tensor name: " + s; return JF_ERR_INVALID; }
  const int l = atoi(s.substr(pre.size(), dot - pre.size()).c_str());
  if (l < 0 || l >= md.layers) { h->err = "layer index out of range: " + s; return JF_ERR_INVALID; }
  const std::string rest = s.substr(dot + 1);
  const LayerW& lw = c.layer[l];
  auto vec_copy = [&](const __nv_bfloat16* dst, int64_t len, int bit) -> int {
    if (!want1(len)) return JF_ERR_INVALID;
    int r2 = stage_tensor(h, data, n, dtype);
    if (r2 != JF_OK) return r2;
    JF_CUDA(h, cudaMemcpyAsync(const_cast<__nv_bfloat16*>(dst), h->stage, n * 2, cudaMemcpyDeviceToDevice, h->stream));
    JF_CUDA(h, cudaStreamSynchronize(h->stream));
    h->seen_layer[l] |= (uint32_t)bit;
    return JF_OK;
  };
  auto mat_pack = [&](const __nv_bfloat16* dst, int64_t R, int64_t K, int dest_row0, int bit) -> int {
    if (!want(R, K)) return JF_ERR_INVALID;
    int r2 = stage_tensor(h, data, n, dtype);
    if (r2 != JF_OK) return r2;
    r2 = pack_staged(h, const_cast<__nv_bfloat16*>(dst), (int)R, (int)K, dest_row0);
    if (r2 == JF_OK) h->seen_layer[l] |= (uint32_t)bit;
    return r2;
  };
  auto bias_copy = [&](int off, int64_t len, int bit) -> int {
    if (!want1(len)) return JF_ERR_INVALID;
    int r2 = stage_tensor(h, data, n, dtype);
    if (r2 != JF_OK) return r2;
    bf16_to_f32_kernel<<<64, 256, 0, h->stream>>>(h->stage, const_cast<float*>(lw.bqkv) + off, n);
    ++h->launches;
    JF_CUDA(h, cudaStreamSynchronize(h->stream));
    h->seen_layer[l] |= (uint32_t)bit;
    return JF_OK;
  };
  const int nt_i = (md.inter + 127) / 128;
  if (rest == "self_attn.q_proj.weight") return mat_pack(lw.wqkv, md.q_dim, md.h, 0, W_Q);
  if (rest == "self_attn.k_proj.weight") return mat_pack(lw.wqkv, md.kv_dim, md.h, md.q_dim, W_K);
  if (rest == "self_attn.v_proj.weight") return mat_pack(lw.wqkv, md.kv_dim, md.h, md.q_dim + md.kv_dim, W_V);
  if (rest == "self_attn.q_proj.bias") return bias_copy(0, md.q_dim, B_Q);
  if (rest == "self_attn.k_proj.bias") return bias_copy(md.q_dim, md.kv_dim, B_K);
  if (rest == "self_attn.v_proj.bias") return bias_copy(md.q_dim + md.kv_dim, md.kv_dim, B_V);
  if (rest == "self_attn.o_proj.weight") return mat_pack(lw.wo, md.h, md.q_dim, 0, W_O);
  if (rest == "mlp.gate_proj.weight") return mat_pack(lw.wgu, md.inter, md.h, 0, W_G);
  if (rest == "mlp.up_proj.weight") return mat_pack(lw.wgu, md.inter, md.h, nt_i * 128, W_U);
  if (rest == "mlp.down_proj.weight") return mat_pack(lw.wd, md.h, md.inter, 0, W_D);
  if (rest == "input_layernorm.weight") return vec_copy(lw.ln1, md.h, W_LN1);
  if (rest == "post_attention_layernorm.weight") return vec_copy(lw.ln2, md.h, W_LN2);
  if (rest == "self_attn.rotary_emb.inv_freq") return JF_OK;
  h->err = "unknown tensor name: " + s;
  return JF_ERR_INVALID;
}

extern "C" int jf_finalize_weights(jf_handle h) {
  if (!h) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  if (!(h->seen_global & G_EMBED)) { h->err = "model.embed_tokens.weight was never loaded"; return JF_ERR_INVALID; }
  if (!(h->seen_global & G_HEAD)) { h->err = "lm_head.weight was never loaded (and tie_word_embeddings is off)"; return JF_ERR_INVALID; }
  if (!(h->seen_global & G_NORM)) { h->err = "model.norm.weight was never loaded"; return JF_ERR_INVALID; }
  {   // every layer must have received all of its matrices and norms (the q/k/v biases are optional: zero when absent)
    static const struct { int bit; const char* name; } kNeeded[] = {
        {W_Q, "self_attn.q_proj.weight"}, {W_K, "self_attn.k_proj.weight"}, {W_V, "self_attn.v_proj.weight"},
        {W_O, "self_attn.o_proj.weight"}, {W_G, "mlp.gate_proj.weight"}, {W_U, "mlp.up_proj.weight"},
        {W_D, "mlp.down_proj.weight"}, {W_LN1, "input_layernorm.weight"}, {W_LN2, "post_attention_layernorm.weight"}};
    for (int l = 0; l < h->md.layers; ++l)
      for (const auto& nd : kNeeded)
        if (!(h->seen_layer[l] & (uint32_t)nd.bit)) {
          h->err = "model.layers." + std::to_string(l) + "." + nd.name + " was never loaded";
          return JF_ERR_INVALID;
        }
  }
  JF_CUDA(h, cudaMemcpyAsync(h->dctx, &h->hctx, sizeof(DevCtx), cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaStreamSynchronize(h->stream));
  h->finalized = true;
  return JF_OK;
}

extern "C" int jf_set_rope_tables(jf_handle h, const float* cosv, const float* sinv, int32_t max_pos) {
  if (!h || !cosv || !sinv) return JF_ERR_INVALID;
  cudaSetDevice(h->device);
  const int half = h->md.d / 2;
  if (max_pos > h->md.max_pos) max_pos = h->md.max_pos;
  const size_t n = (size_t)max_pos * half;
  float* tmp = nullptr;
  JF_CUDA(h, cudaMalloc(&tmp, n * 4));
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemcpyAsync(tmp, pass ? sinv : cosv, n * 4, cudaMemcpyDefault, h->stream);
    convert_to_bf16_kernel<<<256, 256, 0, h->stream>>>(tmp, const_cast<__nv_bfloat16*>(pass ? h->hctx.rope_sin : h->hctx.rope_cos), n, JF_DTYPE_F32);
    ++h->launches;
    cudaStreamSynchronize(h->stream);
  }
  cudaFree(tmp);
  JF_CUDA(h, cudaGetLastError());
  return JF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------------------------
static int launch_persistent(jf_engine* h, int run_mode, int max_forwards, float* elapsed_ms) {
  JF_CUDA(h, cudaMemsetAsync(h->hctx.bar_counter, 0, (4 + 3 * kMaxTilesPerKind) * sizeof(unsigned int), h->stream));
  if (h->hctx.tag_base > 0xE0000000u) {          // 32-bit stage tags about to wrap: forget every stored tag (0 is never used)
    DevCtx& c = h->hctx;
    JF_CUDA(h, cudaMemsetAsync(c.row_ss, 0, (size_t)c.T_cap * kMaxHTiles * 8, h->stream));
    JF_CUDA(h, cudaMemsetAsync(c.tpart, 0, (size_t)kFuseRopeRows * c.part_splits * c.part_n * 8, h->stream));
    JF_CUDA(h, cudaMemsetAsync(c.tpart_rs[0], 0, (size_t)c.fuse_rows * c.part_splits * h->md.h * 8, h->stream));
    JF_CUDA(h, cudaMemsetAsync(c.tpart_rs[1], 0, (size_t)c.fuse_rows * c.part_splits * h->md.h * 8, h->stream));
    if (c.tgu) JF_CUDA(h, cudaMemsetAsync(c.tgu, 0, (size_t)c.T_cap * c.splits_gu * 2 * h->md.inter * 8, h->stream));
    if (c.tp_size > 1) {
      JF_CUDA(h, cudaMemsetAsync(c.push_buf[c.tp_rank], 0, (size_t)c.push_rows_cap * c.tp_size * h->md.h * 8, h->stream));
      JF_CUDA(h, cudaMemsetAsync(c.stage_buf[c.tp_rank], 0, (size_t)c.fuse_rows * h->md.h * 8, h->stream));
    }
    c.tag_base = 0u;
  }
  h->hctx.xflag_base += (1ull << 32);
  JF_CUDA(h, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->dctx) + offsetof(DevCtx, xflag_base), &h->hctx.xflag_base, 8,
                             cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->dctx) + offsetof(DevCtx, tag_base), &h->hctx.tag_base, 4,
                             cudaMemcpyHostToDevice, h->stream));
  const DevCtx* dctx = h->dctx;
  int smem_total = h->smem_optin;
  void* args[] = {(void*)&dctx, (void*)&smem_total, (void*)&run_mode, (void*)&max_forwards};
  JF_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  JF_CUDA(h, cudaLaunchCooperativeKernel((void*)jf_persistent_kernel, dim3(h->num_ctas), dim3(kNumThreads), args,
                                         (size_t)smem_total, h->stream));
  ++h->launches;
  JF_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  {
    cudaError_t e_ = cudaStreamSynchronize(h->stream);
    if (e_ != cudaSuccess) {
      if (h->dbg_host) {
        fprintf(stderr, "[jf debug] progress words:");
        for (int i = 0; i < 64; ++i) fprintf(stderr, " %d", h->dbg_host[i]);
        fprintf(stderr, "\n");
      }
      h->err = std::string("cudaStreamSynchronize(h->stream): ") + cudaGetErrorString(e_);
      return JF_ERR_CUDA;
    }
  }
  if (elapsed_ms) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    *elapsed_ms = ms;
  }
  {   // advance the stage-tag base past every barrier this launch can have passed (forwards executed is in the header)
    int iters = 0;
    JF_CUDA(h, cudaMemcpy(&iters, reinterpret_cast<const uint8_t*>(h->hctx.hdr) + offsetof(PlanHdr, iter), 4,
                          cudaMemcpyDeviceToHost));
    if (iters < 0) iters = 0;
    h->hctx.tag_base += (unsigned int)(iters + 2) * (unsigned int)(9 * h->md.layers + 8) + 16u;
  }
  return JF_OK;
}

extern "C" int jf_forward(jf_handle h, int32_t T, const int32_t* tokens, int32_t G, const int32_t* groups,
                          int32_t n_seqs, const int32_t* block_tables, int32_t lm_row0, int32_t lm_rows,
                          int32_t n_commit, const int32_t* commits, int32_t* greedy, float* logits, void* hidden,
                          float* elapsed_ms) {
  if (!h || !tokens || !groups || !block_tables) return JF_ERR_INVALID;
  if (!h->finalized) { h->err = "jf_finalize_weights has not been called"; return JF_ERR_INVALID; }
  cudaSetDevice(h->device);
  DevCtx& c = h->hctx;
  const ModelDims& md = h->md;
  if (T < 1 || T > c.T_cap) { h->err = "T outside [1, max_rows]"; return JF_ERR_CAPACITY; }
  if (G < 1 || G > c.G_cap || n_seqs < 1 || n_seqs > c.S_cap) { h->err = "groups / sequences exceed capacity"; return JF_ERR_CAPACITY; }
  if (n_commit < 0 || n_commit > c.T_cap) { h->err = "n_commit outside [0, max_rows]"; return JF_ERR_CAPACITY; }
  if (lm_rows < 0 || lm_row0 < 0 || lm_row0 + lm_rows > T || (lm_row0 & 7)) { h->err = "lm rows must lie inside [0,T) and start at a multiple of 8"; return JF_ERR_INVALID; }
  if (logits && lm_rows > kArgmaxRows) { h->err = "logits are available for at most 256 rows per forward"; return JF_ERR_CAPACITY; }
  std::vector<Group> hg(G);
  std::vector<int> row_group(T, -1), row_pos(T, 0);
  std::vector<int2> row_attn(T);
  int units = 0, slots = 0;
  int base_units = 0, max_len = 1;
  for (int g = 0; g < G; ++g) {
    const int len = groups[4 * g + 1], ctx = groups[4 * g + 3];
    if (len >= 1 && ctx >= 0) {
      base_units += md.n_kv * ((ctx + len + c.attn_chunk - 1) / c.attn_chunk) * ((len + kQTokPerUnit - 1) / kQTokPerUnit);
      max_len = std::max(max_len, len);
    }
  }
  const int hsplit = pick_attn_hsplit(base_units, h->num_ctas, md.n_rep, max_len, c.attn_hsplit);
  for (int g = 0; g < G; ++g) {
    Group& gr = hg[g];
    gr.row0 = groups[4 * g]; gr.len = groups[4 * g + 1]; gr.seq = groups[4 * g + 2]; gr.ctx = groups[4 * g + 3];
    if (gr.len < 1 || gr.row0 < 0 || gr.row0 + gr.len > T || gr.seq < 0 || gr.seq >= n_seqs || gr.ctx < 0) { h->err = "malformed group"; return JF_ERR_INVALID; }
    if (gr.ctx + gr.len > c.max_pages_per_seq * kPage || gr.ctx + gr.len > md.max_pos) { h->err = "sequence longer than the KV / position capacity"; return JF_ERR_CAPACITY; }
    gr.nchunks = (gr.ctx + gr.len + c.attn_chunk - 1) / c.attn_chunk;
    gr.qsplit = (gr.len + kQTokPerUnit - 1) / kQTokPerUnit;
    gr.unit0 = units; gr.slot0 = slots;
    units += md.n_kv * gr.nchunks * gr.qsplit * hsplit;
    slots += gr.len * gr.nchunks;
    for (int i = 0; i < gr.len; ++i) {
      if (row_group[gr.row0 + i] != -1) { h->err = "groups overlap"; return JF_ERR_INVALID; }
      row_group[gr.row0 + i] = g;
      row_pos[gr.row0 + i] = gr.ctx + i;
      row_attn[gr.row0 + i] = make_int2(gr.slot0 + i * gr.nchunks, (gr.ctx + i) / c.attn_chunk);
    }
  }
  for (int t = 0; t < T; ++t) if (row_group[t] < 0) { h->err = "a row belongs to no group"; return JF_ERR_INVALID; }
  if (slots > c.attn_slots_cap) { h->err = "attention partial slots exceed capacity (raise attn_slots or attn_chunk)"; return JF_ERR_CAPACITY; }
  PlanHdr hdr;
  memset(&hdr, 0, sizeof(hdr));
  hdr.T = T; hdr.G = G; hdr.done = 0; hdr.n_commit = 0; hdr.n_commit_after = n_commit; hdr.attn_units = units;
  hdr.attn_slots = slots;
  hdr.attn_hsplit = hsplit;
  hdr.n_lm_tasks = 0;
  for (int r = 0; r < lm_rows; r += kArgmaxRows) {
    if (hdr.n_lm_tasks >= kMaxLmTasks) { h->err = "too many LM rows"; return JF_ERR_CAPACITY; }
    hdr.lm_row0[hdr.n_lm_tasks] = lm_row0 + r;
    hdr.lm_rows[hdr.n_lm_tasks] = std::min(kArgmaxRows, lm_rows - r);
    ++hdr.n_lm_tasks;
  }
  hdr.want_logits = (logits != nullptr && lm_rows > 0) ? 1 : 0;
  if (hdr.want_logits && c.logits == nullptr) {
    int rc = dev_alloc(h, &c.logits, (size_t)kArgmaxRows * md.vocab, &h->bytes_act);
    if (rc != JF_OK) return rc;
    JF_CUDA(h, cudaMemcpyAsync(h->dctx, &h->hctx, sizeof(DevCtx), cudaMemcpyHostToDevice, h->stream));
  }
  std::vector<CommitEntry> hc(n_commit > 0 ? n_commit : 1);
  for (int i = 0; i < n_commit; ++i) {
    hc[i].src_row = commits[3 * i]; hc[i].seq = commits[3 * i + 1]; hc[i].pos = commits[3 * i + 2];
    if (hc[i].src_row < 0 || hc[i].src_row >= T || hc[i].seq < 0 || hc[i].seq >= n_seqs || hc[i].pos < 0 ||
        hc[i].pos >= c.max_pages_per_seq * kPage) { h->err = "malformed commit entry"; return JF_ERR_INVALID; }
  }
  JF_CUDA(h, cudaMemcpyAsync(c.hdr, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.tok, tokens, (size_t)T * 4, cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.row_group, row_group.data(), (size_t)T * 4, cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.row_pos, row_pos.data(), (size_t)T * 4, cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.row_attn, row_attn.data(), (size_t)T * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.groups, hg.data(), (size_t)G * sizeof(Group), cudaMemcpyHostToDevice, h->stream));
  if (n_commit > 0) JF_CUDA(h, cudaMemcpyAsync(c.commits, hc.data(), (size_t)n_commit * sizeof(CommitEntry), cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.block_tables, block_tables, (size_t)n_seqs * c.max_pages_per_seq * 4, cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaStreamSynchronize(h->stream));     // the host vectors above go out of scope
  int rc = launch_persistent(h, RUN_PLAN, 1, elapsed_ms);
  if (rc != JF_OK) return rc;
  {
    int dev_err = 0;
    JF_CUDA(h, cudaMemcpy(&dev_err, reinterpret_cast<const uint8_t*>(c.hdr) + offsetof(PlanHdr, error), 4, cudaMemcpyDeviceToHost));
    if (dev_err != 0) { h->err = "device-side error " + std::to_string(dev_err) + " during the forward"; return JF_ERR_CUDA; }
  }
  if (greedy) JF_CUDA(h, cudaMemcpy(greedy, c.greedy, (size_t)T * 4, cudaMemcpyDeviceToHost));
  if (logits && lm_rows > 0) JF_CUDA(h, cudaMemcpy(logits, c.logits, (size_t)lm_rows * md.vocab * 4, cudaMemcpyDeviceToHost));
  if (hidden) JF_CUDA(h, cudaMemcpy(hidden, c.hidden, (size_t)T * md.h * 2, cudaMemcpyDeviceToHost));
  return JF_OK;
}

extern "C" int jf_commit_rows(jf_handle h, int32_t n_seqs, const int32_t* block_tables, int32_t n_commit,
                              const int32_t* commits) {
  if (!h || !block_tables || (n_commit > 0 && !commits)) return JF_ERR_INVALID;
  if (!h->finalized) { h->err = "jf_finalize_weights has not been called"; return JF_ERR_INVALID; }
  cudaSetDevice(h->device);
  DevCtx& c = h->hctx;
  if (n_seqs < 1 || n_seqs > c.S_cap || n_commit < 0 || n_commit > c.T_cap) { h->err = "commit request exceeds capacity"; return JF_ERR_CAPACITY; }
  if (n_commit == 0) return JF_OK;
  std::vector<CommitEntry> hc(n_commit);
  for (int i = 0; i < n_commit; ++i) {
    hc[i].src_row = commits[3 * i]; hc[i].seq = commits[3 * i + 1]; hc[i].pos = commits[3 * i + 2];
    if (hc[i].src_row < 0 || hc[i].src_row >= c.T_cap || hc[i].seq < 0 || hc[i].seq >= n_seqs || hc[i].pos < 0 ||
        hc[i].pos >= c.max_pages_per_seq * kPage) { h->err = "malformed commit entry"; return JF_ERR_INVALID; }
  }
  PlanHdr hdr;
  memset(&hdr, 0, sizeof(hdr));
  hdr.done = 1; hdr.n_commit = n_commit;
  JF_CUDA(h, cudaMemcpyAsync(c.hdr, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.commits, hc.data(), (size_t)n_commit * sizeof(CommitEntry), cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaMemcpyAsync(c.block_tables, block_tables, (size_t)n_seqs * c.max_pages_per_seq * 4, cudaMemcpyHostToDevice, h->stream));
  JF_CUDA(h, cudaStreamSynchronize(h->stream));
  return launch_persistent(h, RUN_PLAN, 1, nullptr);
}

extern "C" int jf_jacobi_run(jf_handle h, int32_t n_seqs, const jf_policy_knobs* knobs, const int32_t* prompts,
                             const int32_t* prompt_lens, const int32_t* ctx_lens, const int32_t* first_correct,
                             const int32_t* first_call_tokens, const uint32_t* rng_state, int32_t rng_words,
                             const int32_t* pad_stream, int32_t pad_stream_len, const int32_t* block_tables,
                             int32_t max_forwards, int32_t* out_tokens, int32_t out_capacity, int32_t* out_summary,
                             int32_t* iters_log, float* run_stats, int32_t* trace, int32_t* trace_len) {
  if (!h || !knobs || !prompts || !prompt_lens || !ctx_lens || !block_tables || !out_tokens || !out_summary) return JF_ERR_INVALID;
  if (!h->finalized) { h->err = "jf_finalize_weights has not been called"; return JF_ERR_INVALID; }
  cudaSetDevice(h->device);
  DevCtx& c = h->hctx;
  if (n_seqs < 1 || n_seqs > c.S_cap) { h->err = "n_seqs outside [1, max_seqs]"; return JF_ERR_CAPACITY; }
  PolicyKnobs kn;
  memcpy(&kn, knobs, sizeof(kn));
  if (kn.n < 1 || kn.n > kNMax) { h->err = "block length n outside [1, 128]"; return JF_ERR_INVALID; }
  if ((kn.mode == MODE_ENGINE || kn.mode == MODE_ENGINE_NG) && kn.n < 2) { h->err = "Jacobi draft length L must be >= 2"; return JF_ERR_INVALID; }
  float ng_temp = 0.f;
  if (kn.mode == MODE_ENGINE_NG) {
    memcpy(&ng_temp, &kn.temp_bits, 4);
    if (!(ng_temp > 0.f) || !std::isfinite(ng_temp)) { h->err = "non-greedy Jacobi needs a finite temperature > 0"; return JF_ERR_INVALID; }
  }
  if (kn.mode == MODE_HF_MULTI && (kn.K < 1 || kn.K > 2)) { h->err = "device multiblock policy supports K in {1, 2}"; return JF_ERR_UNSUPPORTED; }
  if (kn.pool_size < 0 || kn.pool_size > kPoolMax) { h->err = "n_gram_pool_size outside [0, 16]"; return JF_ERR_INVALID; }
  if (kn.mode < 0 || kn.mode > 3) { h->err = "unknown policy mode"; return JF_ERR_INVALID; }
  if ((kn.mode == MODE_HF_SINGLE || kn.mode == MODE_HF_MULTI) && !first_call_tokens) { h->err = "first_call_tokens required"; return JF_ERR_INVALID; }
  if (kn.max_calls > c.iters_cap) kn.max_calls = c.iters_cap;
  kn.vocab = h->md.vocab;
  const int n = kn.n;
  const int cr = (rng_state && rng_words >= 625) ? rng_words : 0;
  // device copies of the draft sources
  int *d_first = nullptr, *d_choice = nullptr, *d_pad = nullptr;
  // engine-owned scratch, grown on demand and kept between calls (no cudaMalloc / cudaFree on the per-call path)
  auto tmp_alloc = [&](int** p, size_t count, int which) -> int {
    if (count > h->run_scratch_words[which]) {
      void* q = nullptr;
      const size_t words = count + count / 2 + 256;
      cudaError_t e = cudaMalloc(&q, words * 4);
      if (e != cudaSuccess) { h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return JF_ERR_CUDA; }
      if (h->run_scratch[which]) cudaFree(h->run_scratch[which]);
      h->run_scratch[which] = reinterpret_cast<int*>(q);
      h->run_scratch_words[which] = words;
    }
    *p = h->run_scratch[which];
    return JF_OK;
  };
  auto cleanup = [&]() {};
  int rc = JF_OK;
#define TRY(x) do { rc = (x); if (rc != JF_OK) { cleanup(); return rc; } } while (0)
#define TRYC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { h->err = std::string(#x) + ": " + cudaGetErrorString(e_); cleanup(); return JF_ERR_CUDA; } } while (0)
  if (first_call_tokens) {
    TRY(tmp_alloc(&d_first, (size_t)n_seqs * n, 0));
    TRYC(cudaMemcpyAsync(d_first, first_call_tokens, (size_t)n_seqs * n * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if (cr > 0) {
    TRY(tmp_alloc(&d_choice, (size_t)n_seqs * cr, 1));
    TRYC(cudaMemcpyAsync(d_choice, rng_state, (size_t)n_seqs * cr * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if (pad_stream && pad_stream_len > 0) {
    TRY(tmp_alloc(&d_pad, (size_t)n_seqs * pad_stream_len, 2));
    TRYC(cudaMemcpyAsync(d_pad, pad_stream, (size_t)n_seqs * pad_stream_len * 4, cudaMemcpyHostToDevice, h->stream));
  }
  std::vector<PolicyKnobs> hk(n_seqs, kn);
  std::vector<DraftSource> hs(n_seqs);
  std::vector<SeqSlot> hslot(n_seqs);
  memset(hslot.data(), 0, sizeof(SeqSlot) * n_seqs);
  std::vector<int> hgen((size_t)n_seqs * c.gen_cap, 0);
  h->h_states.resize(n_seqs);
  size_t poff = 0;
  for (int s = 0; s < n_seqs; ++s) {
    const int pl = prompt_lens[s];
    if (pl < 1 || pl + kn.max_new_tokens + 2 * n > c.gen_cap) { h->err = "prompt + max_new_tokens exceed gen_capacity"; cleanup(); return JF_ERR_CAPACITY; }
    if (ctx_lens[s] < 0 || ctx_lens[s] > pl) { h->err = "ctx_len must lie in [0, prompt_len]"; cleanup(); return JF_ERR_INVALID; }
    memcpy(&hgen[(size_t)s * c.gen_cap], prompts + poff, (size_t)pl * 4);
    poff += pl;
    SeqState& st = h->h_states[s];
    memset(&st, 0, offsetof(SeqState, ret));
    seq_state_init(st, pl, ctx_lens[s], first_correct ? first_correct[s] : 0);
    hs[s].first_call_tokens = d_first ? d_first + (size_t)s * n : nullptr;
    hs[s].mt = d_choice ? reinterpret_cast<unsigned int*>(d_choice) + (size_t)s * cr : nullptr;
    hs[s].mt_words = d_choice ? cr : 0;
    hs[s].pad_stream = d_pad ? d_pad + (size_t)s * pad_stream_len : nullptr;
    hs[s].pad_stream_len = d_pad ? pad_stream_len : 0;
  }
  for (int s = 0; s < n_seqs; ++s)   // only the scalar head of SeqState needs initialising
    TRYC(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(c.seq_state + s), &h->h_states[s], offsetof(SeqState, ret), cudaMemcpyHostToDevice, h->stream));
  TRYC(cudaMemcpyAsync(c.seq_slot, hslot.data(), sizeof(SeqSlot) * n_seqs, cudaMemcpyHostToDevice, h->stream));
  TRYC(cudaMemcpyAsync(c.knobs, hk.data(), sizeof(PolicyKnobs) * n_seqs, cudaMemcpyHostToDevice, h->stream));
  TRYC(cudaMemcpyAsync(c.draft_src, hs.data(), sizeof(DraftSource) * n_seqs, cudaMemcpyHostToDevice, h->stream));
  TRYC(cudaMemcpyAsync(c.gen, hgen.data(), hgen.size() * 4, cudaMemcpyHostToDevice, h->stream));
  TRYC(cudaMemcpyAsync(c.block_tables, block_tables, (size_t)n_seqs * c.max_pages_per_seq * 4, cudaMemcpyHostToDevice, h->stream));
  TRYC(cudaMemsetAsync(c.iters_log, 0, (size_t)c.S_cap * c.iters_cap * 4, h->stream));
  TRYC(cudaMemsetAsync(c.hdr, 0, sizeof(PlanHdr), h->stream));
  TRYC(cudaMemsetAsync(c.trace_len, 0, 16, h->stream));
  {
    const int ng = kn.mode == MODE_ENGINE_NG ? 1 : 0;
    const float inv_t = ng ? 1.0f / ng_temp : 1.0f;
    if (c.n_seqs != n_seqs || c.ng_mode != ng || (ng && (c.ng_inv_temp != inv_t || c.ng_seed_lo != kn.seed_lo ||
                                                         c.ng_seed_hi != kn.seed_hi || c.ng_rng_offset != kn.rng_offset))) {
      c.n_seqs = n_seqs;
      c.ng_mode = ng; c.ng_inv_temp = inv_t; c.ng_seed_lo = kn.seed_lo; c.ng_seed_hi = kn.seed_hi; c.ng_rng_offset = kn.rng_offset;
      TRYC(cudaMemcpyAsync(h->dctx, &h->hctx, sizeof(DevCtx), cudaMemcpyHostToDevice, h->stream));
    }
    if (ng) {           // accumulators of the sampling epilogue start clean (the policy stage re-arms them after every forward)
      std::vector<float> ninf((size_t)2 * c.T_cap, -INFINITY);
      TRYC(cudaMemcpyAsync(c.ng_prop, ninf.data(), (size_t)2 * c.T_cap * 4, cudaMemcpyHostToDevice, h->stream));
      TRYC(cudaMemsetAsync(c.ng_gumbel, 0, (size_t)2 * c.T_cap * 8, h->stream));
      TRYC(cudaStreamSynchronize(h->stream));
    }
  }
  TRYC(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  TRY(launch_persistent(h, RUN_JACOBI, max_forwards > 0 ? max_forwards : (1 << 30), &ms));
  // results
  PlanHdr hdr;
  TRYC(cudaMemcpy(&hdr, c.hdr, sizeof(hdr), cudaMemcpyDeviceToHost));
  TRYC(cudaMemcpy(hgen.data(), c.gen, hgen.size() * 4, cudaMemcpyDeviceToHost));
  int policy_err = 0;
  for (int s = 0; s < n_seqs; ++s) {
    SeqState& st = h->h_states[s];
    TRYC(cudaMemcpy(&st, reinterpret_cast<uint8_t*>(c.seq_state + s), offsetof(SeqState, ret), cudaMemcpyDeviceToHost));
    const int ngen = st.gen_len - st.prompt_len0;
    const int ncopy = std::min(ngen, (int)out_capacity);
    memcpy(out_tokens + (size_t)s * out_capacity, &hgen[(size_t)s * c.gen_cap + st.prompt_len0], (size_t)ncopy * 4);
    int32_t* sm = out_summary + (size_t)s * 8;
    sm[0] = ngen; sm[1] = st.calls; sm[2] = st.stop_reason; sm[3] = st.total_new; sm[4] = st.error;
    sm[5] = st.total_iters; sm[6] = st.ctx; sm[7] = st.first_correct;
    if (st.error) policy_err = st.error;
  }
  if (iters_log) TRYC(cudaMemcpy(iters_log, c.iters_log, (size_t)n_seqs * c.iters_cap * 4, cudaMemcpyDeviceToHost));
  if (run_stats) { run_stats[0] = (float)hdr.iter; run_stats[1] = ms; run_stats[2] = (float)hdr.rows_total; run_stats[3] = 0.f; }
  if (trace && trace_len) {
    int tl = 0;
    TRYC(cudaMemcpy(&tl, c.trace_len, 4, cudaMemcpyDeviceToHost));
    if (tl > 0 && c.trace) TRYC(cudaMemcpy(trace, c.trace, (size_t)tl * 4, cudaMemcpyDeviceToHost));
    *trace_len = tl;
  } else if (trace_len) {
    *trace_len = 0;
  }
  cleanup();
#undef TRY
#undef TRYC
  if (policy_err) { h->err = "device policy error " + std::to_string(policy_err); return JF_ERR_POLICY; }
  return JF_OK;
}
