// jf_types.h — device-visible structures shared by the host runtime (jf_engine.cu) and the persistent kernel.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include "jf_policy.h"

namespace jf {

constexpr int kPage        = 256;    // KV page size in tokens (reference: kvcache_block_size % 256 == 0, config.py:46)
constexpr int kMaxLayers   = 64;     // the per-layer weight table lives in shared memory (56 B per layer)
constexpr int kAttnSub     = 64;     // keys per mma.sync sub-tile
constexpr int kMaxLmTasks  = 16;     // LM-head row tiles of kArgmaxRows per forward
constexpr int kQTokPerUnit = 64;     // query tokens one attention unit covers
constexpr int kMaxHTiles   = 64;     // 128-column tiles of a hidden row (h <= 8192)
constexpr int kMaxTilesPerKind = 128; // arrival counters per fused GEMM kind (O, down, QKV)
constexpr int kFuseRopeRows = 64;    // forwards of at most this many rows run the fused QKV + rotary tail
constexpr int kFuseRows    = 320;    // default row capacity of the tagged partial buffers of the fused O / down tails (one GPU)
constexpr int kFuseRowsTP  = 2048;   // under tensor parallelism: the fused tails exchange ~1/ranks of what the stand-alone
                                     // stage pulls over NVLink, so batched / prefill-sized forwards stay fused too
constexpr int kMaxTP       = 8;      // tensor-parallel ranks whose buffers one kernel may address (NVLink peers)

struct ModelDims {
  int h, layers, n_q, n_kv, d, inter, vocab;
  int q_dim, kv_dim, qkv_dim;
  int n_rep;                 // n_q / n_kv
  int tie;                   // lm_head shares the embedding matrix
  float eps;
  float rope_theta;
  int max_pos;               // rows of the rope tables
  int vocab_embed;           // rows of the (replicated) embedding table; `vocab` is the LM-head slice of this rank
  int vocab_offset;          // first vocabulary id of this rank's LM-head slice
};

struct LayerW {
  const __nv_bfloat16* wqkv;   // packed tiles, rows = [q | k | v] (HF q_proj / k_proj / v_proj stacked)
  const float* bqkv;           // [qkv_dim] fp32 (bf16 values)
  const __nv_bfloat16* wo;     // packed [h x q_dim]
  const __nv_bfloat16* wgu;    // packed gate tiles [inter x h] then up tiles
  const __nv_bfloat16* wd;     // packed [h x inter]
  const __nv_bfloat16* ln1;    // input_layernorm.weight [h]
  const __nv_bfloat16* ln2;    // post_attention_layernorm.weight [h]
};

// One attention group = one candidate row of one sequence: `len` in-flight tokens at positions ctx .. ctx+len-1,
// causal among themselves, all seeing the sequence's committed KV [0, ctx).
struct Group {
  int row0, len, seq, ctx;
  int nchunks;               // ceil((ctx + len) / attn_chunk)
  int qsplit;                // ceil(len / kQTokPerUnit)
  int unit0;                 // first attention unit of this group (units = n_kv * nchunks * qsplit * attn_hsplit)
  int slot0;                 // first partial slot (slots = len * nchunks)
};

struct CommitEntry { int src_row, seq, pos; };

struct PlanHdr {
  int T;                     // in-flight token rows of this forward
  int G;                     // groups
  int done;                  // 1: no forward; apply the commit list and exit
  int n_commit;
  int attn_units;
  int attn_slots;
  int n_lm_tasks;
  int lm_row0[kMaxLmTasks];
  int lm_rows[kMaxLmTasks];
  int iter;                  // forwards executed so far in this launch
  int error;
  int want_logits;           // plan mode: also write fp32 logits of LM task 0
  int n_commit_after;        // plan mode: commit entries to apply once the forward is done
  int rows_total;            // sum of T over the forwards of this launch
  int attn_hsplit;           // attention units per kv head along its q heads, chosen per forward (pick_attn_hsplit)
};

struct SeqSlot {             // per-sequence bookkeeping owned by the policy stage
  int row0;                  // first plan row of the sequence's rows in the forward in flight
  int B, L;                  // rows x columns in flight
  int g0;                    // first group
  int ctx_at_forward;        // committed KV length when the forward in flight was planned
  int active;                // has rows in the forward in flight
  int had_forward;
  int commit_src_row, commit_n, commit_pos;
};

struct alignas(16) DevCtx {
  ModelDims md;
  // ---- weights
  const __nv_bfloat16* embed;        // [vocab][h] row-major
  const __nv_bfloat16* lm_head;      // packed tiles [ceil(vocab/128)][h/64]
  const __nv_bfloat16* final_norm;   // [h]
  const __nv_bfloat16* rope_cos;     // [max_pos][d/2]
  const __nv_bfloat16* rope_sin;
  LayerW layer[kMaxLayers];
  // ---- paged KV cache [layer][2][num_pages][n_kv][kPage][d] and the in-flight scratch [layer][2][T_cap][n_kv][d]
  __nv_bfloat16* kv_cache;
  __nv_bfloat16* kv_scratch;
  int num_pages;
  int max_pages_per_seq;
  int* block_tables;                 // [S_cap][max_pages_per_seq]
  // ---- activations
  int T_cap, xrows_cap, S_cap, G_cap;
  __nv_bfloat16* hidden;             // [T_cap][h] residual stream
  __nv_bfloat16* xn;                 // packed [h/64][xrows_cap x 64]   normed hidden (GEMM X operand)
  __nv_bfloat16* xa;                 // packed [q_dim/64][...]          attention output
  __nv_bfloat16* xact;               // packed [inter/64][...]          silu(gate)*up
  float* part;                       // [T_cap][part_splits][part_n] split-K partial sums, row-major by ROW so the live
                                     // rows of a forward share a few 2 MB pages (TLB reach matters: see DESIGN.md)
  int part_splits, part_n;
  unsigned long long* tpart;         // [kFuseRopeRows][part_splits][part_n] tagged {stage tag, fp32} partial sums of the fused QKV stage
  unsigned long long* tpart_rs[2];   // [fuse_rows][part_splits][h] the same for the fused O / down stages
  unsigned long long* tgu;           // [T_cap][splits_gu][2][inter] tagged gate / up partial sums (split-K gate/up)
  float* part_rs[3];                 // [T_cap][part_splits][h]: partials of the row-split GEMMs (0 = O, 1 = down); with
                                     // tensor parallelism the peers read them over NVLink, hence double-buffered
                                     // (MLP-only tensor parallelism: down alternates 0 / 1 by layer, the local O uses 2)
  __nv_bfloat16* qbuf;               // [T_cap][q_dim] rotated queries
  float* attn_o;                     // [attn_slots_cap][n_q][d] unnormalised partial outputs
  float* attn_ml;                    // [attn_slots_cap][n_q][2] (running max, running sum)
  int attn_slots_cap;
  int attn_chunk;                    // keys per attention unit (multiple of kAttnSub)
  int attn_hsplit;                   // > 0: fixed attention units per kv head along its q heads; 0: chosen per forward
  unsigned long long* argmax_keys;   // [2][T_cap] packed (value, ~index) keys, double-buffered by forward parity
  unsigned long long* row_ss;        // [T_cap][kMaxHTiles] per-tile {stage tag, sum of squares} words of the row being normalised
  unsigned int tag_base;             // stage tag = tag_base + barriers passed + 1: a 32-bit count of all stages this engine
                                     // ever ran, so a word's stale tag (its LAST writer's) never equals a later stage's
  int splits_qkv, splits_o, splits_down;
  int splits_gu;                     // > 1: the gate/up projection is split along K too (its tiles alone would idle SMs)
  // ---- plan (written by the host in plan mode, by the policy stage in Jacobi mode)
  PlanHdr* hdr;
  int* tok;                          // [T_cap]
  int* row_group;                    // [T_cap]
  int* row_pos;                      // [T_cap] absolute position of the row's token
  int2* row_attn;                    // [T_cap] {first partial slot of the row, last key chunk it can see}
  Group* groups;                     // [G_cap]
  CommitEntry* commits;              // [T_cap]
  // ---- outputs / policy
  int* greedy;                       // [T_cap] argmax token per in-flight row of the last forward
  float* logits;                     // plan mode debug: [kArgmaxRows][vocab] fp32 (may be null)
  SeqState* seq_state;               // [S_cap]
  SeqSlot* seq_slot;                 // [S_cap]
  PolicyKnobs* knobs;                // [S_cap]
  DraftSource* draft_src;            // [S_cap]
  int* gen;                          // [S_cap][gen_cap] prompt + generated tokens
  int gen_cap;
  int* iters_log;                    // [S_cap][iters_cap]
  int iters_cap;
  int n_seqs;
  // ---- non-greedy verify (MODE_ENGINE_NG; reference: engine/jacobi_decoding_nongreedy.py:300-355)
  int ng_mode;                       // 1: the LM head runs EPI_SAMPLE and the policy stage the rejection-sampling verify
  float ng_inv_temp;                 // 1 / temperature
  unsigned int ng_seed_lo, ng_seed_hi;
  int ng_rng_offset;
  // the three accumulators below are double-buffered by forward parity, like argmax_keys: under tensor parallelism the
  // peers read them (vocabulary-parallel LM head) while this rank may already be clearing the other half
  float* ng_prop;                    // [2][T_cap] logit / T of the drafted token of every row (-inf: not in this rank's slice)
  unsigned long long* ng_gumbel;     // [2][T_cap] best Gumbel-perturbed key among the other tokens (of this rank's slice)
  float2* ng_ms;                     // [2][nctas][T_cap] per-CTA (max, sum exp) of logit / T
  int ng_parts;                      // CTAs that hold LM-head tiles (identical on every rank)
  float* ng_px;                      // [T_cap] merged by the policy stage: probability of the drafted token
  int* ng_bonus;                     // [T_cap] and the residual sample
  int* trace;                        // optional forward trace: {iter, seq, ctx, B, L, tokens...}*
  int trace_cap;
  int* trace_len;
  // ---- tensor parallelism: every rank runs the same kernel on its shard and talks to its peers through mapped peer
  // memory (CUDA IPC over NVLink; plain pointers when several ranks share one device).  Fused (Jacobi-sized) forwards
  // exchange nothing but TAGGED DATA - no cross-GPU flag, fence or atomic: a row-split projection reduce-scatters by
  // push (rank r adds its own k-splits and stores the sum of row t into the push buffer of the row's OWNER, t mod
  // ranks), the owner adds the ranks in rank order, applies residual + RMSNorm and pushes the finished row into every
  // peer's staging buffer (all-gather by push), from where each rank copies it into its own hidden / xn before its
  // LOCAL grid barrier.  Prefill-sized forwards: partial sums stay local, peers read them after a flag barrier.
  int tp_rank, tp_size;
  int tp_attn;                                // ranks the attention block (QKV / O) is sharded over: tp_size, or 1 when
                                              // attention is replicated on every rank (MLP-only tensor parallelism)
  int fuse_rows;                              // forwards of at most this many rows run the fused O / down tails
  int push_rows_cap;                          // owned rows a push buffer holds (>= ceil(fuse_rows / ranks))
  unsigned long long* push_buf[kMaxTP];       // rank r's push buffer [push_rows_cap][tp_size][h] of tagged partial sums
  unsigned long long* stage_buf[kMaxTP];      // rank r's staging buffer [fuse_rows][h / 8][8] of tagged finished rows:
                                              // per 8-column group, words 0-3 = hidden, 4-7 = xn (two bf16 per word)
  const uint8_t* arena_self;                  // this rank's activation arena and the peers' mappings of theirs: every rank
  const uint8_t* peer_arena[kMaxTP];          // lays its arena out identically, so peer_ptr() reaches any peer buffer
  const float* peer_part_rs[kMaxTP][3];       // the peers' fp32 partial buffers (stand-alone reduce stage, prefill)
  const unsigned long long* peer_keys[kMaxTP];
  unsigned long long* peer_xflag[kMaxTP];     // rank r's flag array (this rank writes element [tp_rank])
  unsigned long long* xflag;                  // own flag array [kMaxTP]
  unsigned long long xflag_base;              // launch sequence << 32 (flags are monotonic, never reset)
  // ---- sync
  unsigned int* bar_counter;
  unsigned int* tile_ctr;            // [3][kMaxTilesPerKind] arrival counters, zeroed per launch ([0..splits_down): gate/up tiles per down k-split)
  int fuse_rope;                     // 1: the QKV projection runs bias + rotary + q/k/v scatter as its epilogue tail
  int dataflow_gd;                   // 1: no grid barrier between gate/up and down (per-k-split arrival counters instead)
  int fuse_norm;                     // 1: O / down projections run reduce + residual + RMSNorm as their epilogue tail
  // ---- optional stage profile: SM-clock stamps taken by CTA 0 after every grid barrier of forward `prof_forward`
  long long* prof;
  int prof_cap;
  int prof_forward;
  int* prof_count;
  long long* prof_marks;             // fine-grained (id, clock) pairs recorded by CTA 0 / worker thread 0
  int* prof_marks_n;
  int debug_skip;                    // timing experiments only: bit s = skip the body of SIMT stage s (results invalid)
  int w_evict_first;                 // 1: weight tiles are streamed with the L2 evict_first hint (see jf_ptx.cuh)
  int attn_prefetch;                 // 1: committed K/V of each CTA's first attention unit is fetched during the QKV GEMM
  volatile int* dbg;                 // developer aid: host-mapped progress words (survive a trapped launch)
  int early_fence;                   // 1: the grid barrier's acquire fence is issued behind the arrival, not behind the release
  int x_fence;                       // 1: the activation producer fences (acquire + proxy) after every barrier it waits on
  int epi_bulk;                      // 1: the O / down epilogues store their tagged partial sums row by row with bulk copies
  int swiglu_bulk;                   // 1: the SwiGLU epilogue stages its tile in shared memory and stores it with bulk copies
  int attn_tc;                       // 1: attention on the tcgen05 tensor cores (TMEM-staged Q / P, MN-major V); 0: mma.sync
};

enum RunMode : int { RUN_PLAN = 0, RUN_JACOBI = 1 };

// rank r's copy of one of this rank's activation-arena buffers
template <typename T>
__host__ __device__ inline const T* peer_ptr(const DevCtx& c, int r, const T* own) {
  if (r == c.tp_rank) return own;
  return reinterpret_cast<const T*>(c.peer_arena[r] + (reinterpret_cast<const uint8_t*>(own) - c.arena_self));
}

// Attention units per kv head along its query heads for a forward whose units number `base_units` without any head split
// and whose longest group has `max_len` rows.  The stage is latency-bound: a CTA pays a fixed cost per unit (K / V fill,
// query loads, partial stores: ~3 us) plus ~3 us per round of four 16-row tiles (one per worker warp), and the stage lasts
// as long as its busiest CTA.  The split that minimises waves x (fixed + rounds) wins; measured (profiles/
// r02_attention_hsplit.txt): one sequence of 32 rows at 9 key chunks -> 4 (144 single-round units), at 13 chunks -> 2,
// eight sequences -> 1.  Results do not depend on the split (a row's partial results are computed identically).
__host__ __device__ inline int pick_attn_hsplit(int base_units, int nctas, int n_rep, int max_len, int fixed) {
  if (fixed > 0) return fixed < n_rep ? fixed : (n_rep > 0 ? n_rep : 1);
  const int row_tiles = ((max_len < kQTokPerUnit ? max_len : kQTokPerUnit) + 15) >> 4;
  int best = 1, best_cost = 1 << 30;
  for (int hs = 1; hs <= 4 && hs <= n_rep; hs <<= 1) {
    const int hp = (n_rep + hs - 1) / hs;
    const int waves = (base_units * hs + nctas - 1) / nctas;
    const int rounds = (hp * row_tiles + 3) >> 2;
    const int cost = waves * (3 + 3 * rounds);
    if (cost < best_cost) { best_cost = cost; best = hs; }
  }
  return best;
}

__host__ __device__ inline size_t kv_cache_offset(const ModelDims& md, int num_pages, int layer, int kv, int page,
                                                  int kvh, int slot) {
  return ((((size_t)(layer * 2 + kv) * num_pages + page) * md.n_kv + kvh) * kPage + slot) * (size_t)md.d;
}
__host__ __device__ inline size_t kv_scratch_offset(const ModelDims& md, int T_cap, int layer, int kv, int row,
                                                    int kvh) {
  return (((size_t)(layer * 2 + kv) * T_cap + row) * md.n_kv + kvh) * (size_t)md.d;
}

}  // namespace jf
