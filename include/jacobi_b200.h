/* jacobi_b200.h — C ABI of libjacobi_b200.so: the B200-native greedy Jacobi decoding hot path of a Qwen2 model.
 *
 * This is the drop-in boundary for the reference's per-rank model runner (all paths relative to the reference tree
 * hao-ai-lab/JacobiForcing): the Python `ModelRunner` mirror in jacobiforcing_b200/ binds these entry points with
 * ctypes exactly where the reference calls into torch / flash-attn / NCCL.
 *
 * Conventions: every function returns 0 on success or a negative jf_status; jf_last_error(h) gives the message.
 * No exceptions, no torch / Python types: plain pointers and sizes.  Host buffers are caller-owned; device memory is
 * owned by the handle.  One host thread per handle; all work is enqueued on the handle's stream and the calls that
 * return results synchronise that stream.  There is no CPU fallback: without a CUDA device jf_create fails.
 */
#ifndef JACOBI_B200_H_
#define JACOBI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jf_engine* jf_handle;

typedef enum {
  JF_OK = 0,
  JF_ERR_INVALID = -1,        /* ValueError in the Python mirror */
  JF_ERR_CUDA = -2,           /* RuntimeError */
  JF_ERR_CAPACITY = -3,       /* RuntimeError: rows / KV pages / slots exceed what jf_create reserved */
  JF_ERR_UNSUPPORTED = -4,    /* NotImplementedError */
  JF_ERR_POLICY = -5          /* the device policy hit a condition on which the reference itself raises */
} jf_status;

typedef enum { JF_DTYPE_BF16 = 0, JF_DTYPE_F32 = 1, JF_DTYPE_F16 = 2 } jf_dtype;

/* Qwen2 architecture, the fields the reference reads from hf_config (inference_engine/models/qwen3.py:121-140). */
typedef struct {
  int32_t hidden_size, num_hidden_layers, num_attention_heads, num_key_value_heads, head_dim;
  int32_t intermediate_size, vocab_size, tie_word_embeddings, max_position_embeddings;
  float rms_norm_eps, rope_theta;
} jf_model_cfg;

/* Capacities reserved at creation (reference: Config.max_num_seqs / max_model_len / kvcache sizing,
 * inference_engine/config.py:7-52, engine/model_runner.py:523-622). */
typedef struct {
  int32_t device;                 /* CUDA device ordinal */
  int32_t max_rows;               /* in-flight token rows per forward (B * rows * columns), also the prefill chunk */
  int32_t max_seqs;
  int32_t num_kv_pages;           /* pages of 256 tokens shared by all sequences */
  int32_t max_pages_per_seq;
  int32_t attn_chunk;             /* keys per attention work unit, multiple of 64 (0 = 64) */
  int32_t attn_slots;             /* capacity of split-KV partial slots (0 = derived from max_rows) */
  int32_t gen_capacity;           /* tokens (prompt + generated) kept per sequence on the device */
  int32_t max_calls;              /* rows of the per-call iteration log */
  int32_t trace_capacity;         /* int32 words of forward trace kept per run (0 = off) */
  int32_t tp_rank, tp_size;       /* tensor parallelism: this rank / ranks (1, 2, 4 or 8; 0 = 1) */
  int32_t tp_attn;                /* ranks the attention block is sharded over: 0 / 1 = replicated on every rank (MLP-only
                                     tensor parallelism, one exchange per transformer block); tp_size = sharded */
  int32_t num_ctas;               /* CTAs of the persistent kernel (0 = one per SM); fewer lets several ranks share a GPU */
} jf_runtime_cfg;

/* Greedy Jacobi policy knobs (field order = struct jf::PolicyKnobs, csrc/jf_policy.h).  mode: 0 = HF single block
 * (modeling/cllm2_qwen2_modeling_kv_terminate_on_eos_improved.py:140-276), 1 = HF multiblock + rejection recycling
 * (modeling/..._multiblock_lookahead_unified.py:230-740), 2 = engine fixed-length policy
 * (inference_engine/engine/jacobi_decoding.py:303-445), 3 = engine non-greedy policy
 * (inference_engine/engine/jacobi_decoding_nongreedy.py:300-470). */
typedef struct {
  int32_t mode, n, K, spawn_threshold, pool_size, eos, pad, max_iteration_count, max_new_tokens, max_calls;
  int32_t lookahead_min_acc, vocab;
  /* mode 3 only (non-greedy Jacobi, rejection-sampling verify: engine/jacobi_decoding_nongreedy.py:300-355): float bits of
   * the temperature (> 0), the 64-bit key of the counter-based generator behind every random draw, and an offset added
   * to the forward index in its counters (consecutive runs may share a key). */
  int32_t temp_bits;
  uint32_t seed_lo, seed_hi;
  int32_t rng_offset;
} jf_policy_knobs;

/* replaces ModelRunner.__init__ (engine/model_runner.py:155-291): device, buffers, KV cache */
int jf_create(const jf_model_cfg* model, const jf_runtime_cfg* runtime, jf_handle* out);
int jf_destroy(jf_handle h);
const char* jf_last_error(jf_handle h);     /* h may be NULL: error of the last failed jf_create */
int jf_device_info(jf_handle h, int32_t* sm_count, int64_t* bytes_weights, int64_t* bytes_kv, int64_t* bytes_act);

/* replaces load_model's param.weight_loader calls (inference_engine/utils/loader.py:27-47, layers/linear.py:65-128):
 * called once per HF tensor name (model.layers.N.self_attn.q_proj.weight, ... lm_head.weight); q/k/v and gate/up are
 * merged and re-tiled for the tensor cores here.  `data` may be a host or a device pointer. */
int jf_load_tensor(jf_handle h, const char* hf_name, const void* data, const int64_t* shape, int32_t ndim,
                   int32_t dtype);
int jf_finalize_weights(jf_handle h);       /* checks every tensor arrived; ties lm_head to embed_tokens if asked */
/* optional: rotary tables computed by the caller, fp32 [max_pos][head_dim/2]; default = computed natively */
int jf_set_rope_tables(jf_handle h, const float* cos, const float* sin, int32_t max_pos);

/* KV cache geometry for the host block manager (reference: ModelRunner.kv_cache, engine/llm_engine.py:54-55) */
int jf_kv_info(jf_handle h, int32_t* num_pages, int32_t* page_size, int64_t* bytes_per_page, void** device_base);

/* One forward over caller-planned rows: the reference operator `forward_step_batch` / a prefill
 * (engine/model_runner.py:1134-1418, :777-963; HF: unified.py:175-225, :429-463).
 *   tokens[T]; groups[G][4] = {row0, len, seq, ctx}: `len` consecutive rows at positions ctx..ctx+len-1 of sequence
 *   `seq`, causal among themselves, attending to that sequence's committed KV [0, ctx).
 *   block_tables[n_seqs][max_pages_per_seq]; lm_row0/lm_rows: rows whose argmax (and logits) are wanted (row0 % 8 == 0);
 *   commits[n_commit][3] = {src_row, seq, pos}: in-flight K/V rows to append to the cache after the forward.
 * Outputs (may be NULL): greedy[T] (valid for the lm rows), logits[lm_rows][vocab] fp32 (lm_rows <= 256),
 * hidden[T][hidden_size] bf16 (residual stream after the last layer, before the final norm). */
int jf_forward(jf_handle h, int32_t T, const int32_t* tokens, int32_t G, const int32_t* groups, int32_t n_seqs,
               const int32_t* block_tables, int32_t lm_row0, int32_t lm_rows, int32_t n_commit,
               const int32_t* commits, int32_t* greedy, float* logits, void* hidden, float* elapsed_ms);

/* Append in-flight K/V rows of the LAST jf_forward to the cache without another forward (the reference's logical
 * KV roll-back in reverse: BlockManager.trim_kv_only_fast keeps what was accepted, engine/block_manager.py:534-564). */
int jf_commit_rows(jf_handle h, int32_t n_seqs, const int32_t* block_tables, int32_t n_commit, const int32_t* commits);

/* The device-resident Jacobi fixed-point loop: replaces JacobiDecoder.generate_chunk_batch + the per-iteration
 * forward_step_batch round trips (engine/jacobi_decoding.py:448-724) and, for modes 0/1, the per-call loop of
 * JacobiForcing/jacobi_forcing_inference_MR_humaneval.py:152-240 around jacobi_forward_greedy[_multiblock].
 * Per sequence s (all arrays host memory):
 *   prompts (concatenated) / prompt_lens[s]      tokens already committed to the KV cache (ctx_lens[s] of them)
 *   first_correct[s]                              HF: first_correct_token of the prefill; engine: unused
 *   first_call_tokens[s][n]                       HF: prefill_drafted_n_gram; engine: the prefill draft (n-1 used)
 *   rng_state[s][rng_words]                       HF: state of the random.Random (MT19937: 624 words + index = 625)
 *                                                 whose random.choice(generated_ids) calls seed every later call's
 *                                                 draft; the device continues that exact stream
 *   pad_stream[s][pad_stream_len]                 engine: the random padding tokens, consumed in order
 *   block_tables[s][max_pages_per_seq]
 * Outputs: out_tokens[s][out_capacity] generated ids, out_summary[s][8] = {n_generated, calls, stop_reason,
 * total_new_tokens, policy_error, total_iterations, ctx_len, next_token (first_correct_token for a following call)}, iters_log[s][max_calls] (may be NULL),
 * run_stats[4] = {forwards, elapsed_ms (device), rows_total, reserved}, trace (may be NULL) int32[trace_capacity]
 * with trace_len words used: {iter, seq, ctx, B, L, B*L tokens}*. */
int jf_jacobi_run(jf_handle h, int32_t n_seqs, const jf_policy_knobs* knobs, const int32_t* prompts,
                  const int32_t* prompt_lens, const int32_t* ctx_lens, const int32_t* first_correct,
                  const int32_t* first_call_tokens, const uint32_t* rng_state, int32_t rng_words,
                  const int32_t* pad_stream, int32_t pad_stream_len, const int32_t* block_tables,
                  int32_t max_forwards, int32_t* out_tokens, int32_t out_capacity, int32_t* out_summary,
                  int32_t* iters_log, float* run_stats, int32_t* trace, int32_t* trace_len);

/* Stage profile (the reference's PROFILE=1 ProfileTimer, engine/model_runner.py:26-152, without its synchronising
 * side effects): record SM-clock stamps (arrival of CTA 0, barrier release) at every grid barrier of forward number
 * `forward_index` of the following launches (-1 = off); read them back after a run.  sm_clock_khz converts to time. */
int jf_profile_enable(jf_handle h, int32_t forward_index);
int jf_profile_read(jf_handle h, int64_t* stamps, int32_t capacity, int32_t* n_stamps, int32_t* sm_clock_khz);

/* Tensor parallelism over NVLink peer memory (reference: column/row-parallel linears + dist.all_reduce,
 * inference_engine/layers/linear.py:76-153, vocab-parallel LM head + gather, layers/embed_head.py:56-66).  One rank per
 * GPU; jf_model_cfg describes the FULL model, jf_load_tensor receives this rank's shard of each sharded tensor (gate/up
 * rows, down columns, lm_head rows always; q/k/v rows and o columns when tp_attn = tp_size; embeddings and norms
 * whole).  After jf_create every rank exports a 64-byte IPC handle, the host exchanges them (any transport) and imports
 * all of them in rank order; the kernel then stores partial sums / finished rows straight into the peers' memory and
 * reads their argmax keys there (no NCCL on the data path).  Calls are collective: every rank issues the same sequence
 * of jf_forward / jf_commit_rows / jf_jacobi_run calls. */
int jf_tp_export(jf_handle h, void* handle64);
int jf_tp_import(jf_handle h, const void* handles /* tp_size x 64 bytes, rank order */);
/* The same wiring for ranks that live in ONE process (one host thread per rank): peers[tp_size] = the ranks' handles in
 * rank order.  The ranks may sit on different devices with peer access, or share one device (num_ctas splits its SMs). */
int jf_tp_attach_local(jf_handle h, const jf_handle* peers);

/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches claim). */
int64_t jf_launch_count(jf_handle h);

#ifdef __cplusplus
}
#endif
#endif /* JACOBI_B200_H_ */
