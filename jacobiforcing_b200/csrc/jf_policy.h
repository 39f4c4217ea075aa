// jf_policy.h — the greedy Jacobi accept / shift / re-draft state machines, written once as plain C++ that compiles
// both into the persistent CUDA kernel (device side of the fixed-point loop: no host round trip between iterations)
// and into a host test harness (tests/ replays the reference's golden forwards through exactly this code).
//
// Restates, with fixed-capacity arrays instead of Python lists/tensors:
//   single block   modeling/cllm2_qwen2_modeling_kv_terminate_on_eos_improved.py:140-276
//   multiblock     modeling/cllm2_qwen2_modeling_kv_terminate_on_eos_improved_multiblock_lookahead_unified.py:230-740
//                  (candidates :62-91, committed_len :264-271, build_out_and_spans :317-377, pool concat :387-411)
//   driver         JacobiForcing/jacobi_forcing_inference_MR_humaneval.py:152-240
//   engine policy  inference_engine/engine/jacobi_decoding.py:303-445 (fixed length L, draft[0] = cached seed)
//   non-greedy     inference_engine/engine/jacobi_decoding_nongreedy.py:300-355 (accept draft token x with probability
//                  p(x); on the first rejection commit a token drawn from p conditioned on != x), :385-470 (loop, next draft)
// `num_blocks`, `active_blocks` and the block-list length are three separate things, exactly as in the reference
// (SURVEY.md finding 5 / Appendix A).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define JF_HD __host__ __device__
#else
#define JF_HD
#endif

namespace jf {

constexpr int kNMax         = 128;   // max n_token_seq_len / jacobi_block_len
constexpr int kRowsMax      = 16;    // base draft + candidates (n_gram_pool_size <= 16)
constexpr int kPoolMax      = 16;
constexpr int kBlocksMax    = 4;     // block-list capacity (K <= 2 needs 2; head-room for the pad-block quirk)
constexpr int kPoolEntryMax = kBlocksMax * kNMax;
constexpr int kOutColsMax   = kBlocksMax * kNMax;

enum PolicyMode : int { MODE_HF_SINGLE = 0, MODE_HF_MULTI = 1, MODE_ENGINE = 2, MODE_ENGINE_NG = 3 };
enum StopReason : int { STOP_NONE = 0, STOP_EOS = 1, STOP_MAX_NEW = 2, STOP_MAX_CALLS = 3, STOP_MAX_ITERS = 4,
                        STOP_ERROR = 5 };
enum PolicyError : int { PERR_NONE = 0, PERR_BROADCAST = 1, PERR_NO_PAD = 2, PERR_BLOCKS = 3, PERR_PROMOTE_LEN = 4,
                         PERR_CAPACITY = 5, PERR_DRAFT_SOURCE = 6 };

struct PolicyKnobs {
  int mode;
  int n;                       // n_token_seq_len (HF) / jacobi_block_len L (engine)
  int K;                       // max concurrent blocks (1 RA + K-1 pseudo)
  int spawn_threshold;         // ceil(r * n), computed on the host in double precision like math.ceil(r * n)
  int pool_size;               // n_gram_pool_size (deque maxlen)
  int eos;                     // -1: EOS handling disabled
  int pad;                     // -1: no pad token (spawning then raises in the reference)
  int max_iteration_count;     // per call (HF) / per run (engine jacobi_max_iterations)
  int max_new_tokens;
  int max_calls;
  int lookahead_min_acc;         // smallest total_acc with total_acc / n >= lookahead_start_ratio (host, double precision)
  int vocab;
  // ---- non-greedy engine mode (MODE_ENGINE_NG): rejection-sampling verify, jacobi_decoding_nongreedy.py:300-355
  int temp_bits;               // float bits of the sampling temperature (> 0)
  unsigned int seed_lo, seed_hi;   // key of the counter-based generator (Philox4x32-10) behind every random draw
  int rng_offset;              // added to the forward index in the counters (lets consecutive runs share a key)
};

struct MBlock {
  int nrows;                   // rows of q_draft[b]; an emptied draft is one row of length 0 (q_draft = [[]])
  int len;                     // draft length (same for every row)
  int acc_n;                   // len(out_acc[b])
  int total_acc;
  int need_reverify;
  int acc[kNMax];
  int draft[kRowsMax][kNMax];
};

struct SeqState {
  // ---- driver level (MR_humaneval.py:140-240)
  int calls;                   // the prefill call counts as call 0, so decode starts with calls == 1
  int total_new;
  int gen_len;                 // len(generated_ids) = prompt + accepted so far (tokens live in the caller's gen buffer)
  int prompt_len0;             // len(prompt)
  int finished;
  int stop_reason;
  int error;
  int first_correct;           // next_token carried between calls
  int in_call;
  int eos_seen;                // EOS appeared in the generated part
  int total_iters;
  // ---- call level
  int ctx;                     // committed KV length (prompt_len of the call + what the call has committed)
  int call_ctx0;               // prompt_len at call start (:261)
  int iters;
  int num_blocks, active_blocks, list_len, RA;
  int last_next_token;         // -1 = not set
  int ret_n;                   // tokens returned by the call that just ended
  int pool_count;
  int pool_len[kPoolMax];
  // ---- per-forward bookkeeping
  int B_out, L_out;
  int nspans;
  int span_b[kBlocksMax], span_start[kBlocksMax], span_L[kBlocksMax];
  int commit_row, commit_n;    // which in-flight row's first commit_n tokens become committed KV
  // ---- engine mode
  int eng_have_draft;
  int pad_cursor;              // next unread entry of the caller's random-token stream
  // ---- bulk
  int ret[kOutColsMax];
  int out_rows[kRowsMax][kOutColsMax];
  int pool_tok[kPoolMax][kPoolEntryMax];
  MBlock blk[kBlocksMax];
};

// Where a call's first draft comes from: explicit tokens for the first call (prefill_drafted_n_gram), then
// random.choice(generated_ids) for the n-1 guesses of every later call (MR_humaneval.py:209-215).  The reference uses
// Python's `random` (MT19937); the host hands over the generator's state (random.Random.getstate()[1]: 624 words + index)
// and the device continues the very same stream, so drafts are reproducible whatever number of tokens a call returns.
struct DraftSource {
  const int* first_call_tokens;   // [n]
  unsigned int* mt;               // [625] MT19937 state + index, advanced in place; null when only one call can happen
  int mt_words;
  const int* pad_stream;          // engine mode: torch.randint stand-in, consumed in order
  int pad_stream_len;
};

JF_HD inline unsigned int mt_u32(unsigned int* mt) {                    // CPython _randommodule.c genrand_uint32
  unsigned int idx = mt[624];
  if (idx >= 624) {
    for (int k = 0; k < 624; ++k) {
      const unsigned int y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    idx = 0;
  }
  unsigned int y = mt[idx];
  mt[624] = idx + 1;
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
// random.Random._randbelow_with_getrandbits(n) for 1 <= n < 2^32: k = n.bit_length(); r = getrandbits(k) until r < n
JF_HD inline int mt_randbelow(unsigned int* mt, int n) {
  int k = 0;
  for (unsigned int v = (unsigned int)n; v; v >>= 1) ++k;
  unsigned int r = mt_u32(mt) >> (32 - k);
  while (r >= (unsigned int)n) r = mt_u32(mt) >> (32 - k);
  return (int)r;
}

JF_HD inline int leading_matches(const int* draft_tail, const int* greedy_head, int count) {
  int k = 0;
  while (k < count && draft_tail[k] == greedy_head[k]) ++k;
  return k;
}

// ---------------------------------------------------------------------------------------------------------------
// n-gram pool (collections.deque(maxlen=pool_size))
// ---------------------------------------------------------------------------------------------------------------
JF_HD inline void pool_push(SeqState& s, const PolicyKnobs& kn, const int* tok, int len) {
  if (kn.pool_size <= 0) return;                     // deque(maxlen=0) stays empty
  if (len > kPoolEntryMax) len = kPoolEntryMax;
  if (s.pool_count == kn.pool_size) {                // drop the oldest
    for (int i = 1; i < s.pool_count; ++i) {
      s.pool_len[i - 1] = s.pool_len[i];
      for (int j = 0; j < s.pool_len[i]; ++j) s.pool_tok[i - 1][j] = s.pool_tok[i][j];
    }
    s.pool_count -= 1;
  }
  const int e = s.pool_count++;
  s.pool_len[e] = len;
  for (int j = 0; j < len; ++j) s.pool_tok[e][j] = tok[j];
}

// ---------------------------------------------------------------------------------------------------------------
// multiblock (unified.py:230-740)
// ---------------------------------------------------------------------------------------------------------------
JF_HD inline int mb_committed_len(const SeqState& s, int cur_RA) {       // :264-271
  int c = s.call_ctx0;
  for (int b = 0; b < s.num_blocks; ++b)
    if (b != cur_RA && !s.blk[b].need_reverify) c += s.blk[b].acc_n;
  return c + s.blk[cur_RA].acc_n;
}

JF_HD inline void mb_final_ret(SeqState& s) {                            // :533-537, :602-606, :725-730
  int n = 0;
  for (int bb = 0; bb < s.num_blocks; ++bb)
    if (bb != s.RA && !s.blk[bb].need_reverify && s.blk[bb].acc_n > 0)
      for (int j = 0; j < s.blk[bb].acc_n; ++j) s.ret[n++] = s.blk[bb].acc[j];
  for (int j = 0; j < s.blk[s.RA].acc_n; ++j) s.ret[n++] = s.blk[s.RA].acc[j];
  s.ret_n = n;
}

JF_HD inline void mb_begin_call(SeqState& s, const PolicyKnobs& kn, const int* input_ids) {   // :238-262
  s.pool_count = 0;
  MBlock& b0 = s.blk[0];
  b0.nrows = 1; b0.len = kn.n; b0.acc_n = 0; b0.total_acc = 0; b0.need_reverify = 0;
  for (int j = 0; j < kn.n; ++j) b0.draft[0][j] = input_ids[j];
  s.num_blocks = 1; s.active_blocks = 1; s.list_len = 1; s.RA = 0;
  s.last_next_token = -1;
  s.call_ctx0 = s.ctx;
  s.iters = 0;
  s.ret_n = 0;
  s.in_call = 1;
}

// build_out_and_spans (:317-377).  Returns false when there is nothing to forward (:418-419).
JF_HD inline bool mb_build_rows(SeqState& s, const PolicyKnobs& kn) {
  const MBlock& ra = s.blk[s.RA];
  const int B_ra = ra.nrows;
  int cols = 0;
  int cursor = 1;
  s.nspans = 0;
  const int L_ra = B_ra > 0 ? ra.len : 0;
  if (L_ra > 0) {
    for (int i = 0; i < B_ra; ++i)
      for (int j = 0; j < L_ra; ++j) s.out_rows[i][j] = ra.draft[i][j];
    s.span_b[s.nspans] = s.RA; s.span_start[s.nspans] = cursor; s.span_L[s.nspans] = L_ra; s.nspans++;
    cursor += L_ra; cols += L_ra;
  }
  for (int b = 0; b < s.num_blocks; ++b) {
    if (b == s.RA || !s.blk[b].need_reverify) continue;
    const MBlock& pb = s.blk[b];
    if (cols + pb.acc_n + pb.len > kOutColsMax) { s.error = PERR_CAPACITY; return false; }
    if (pb.acc_n > 0) {
      for (int i = 0; i < B_ra; ++i)
        for (int j = 0; j < pb.acc_n; ++j) s.out_rows[i][cols + j] = pb.acc[j];
      cursor += pb.acc_n; cols += pb.acc_n;
    }
    const int L_tail = pb.nrows > 0 ? pb.len : 0;
    if (L_tail > 0) {
      for (int i = 0; i < B_ra; ++i) {                 // _ensure_batch_like_ra (:288-312): broadcast / tile-and-slice
        const int src = i % pb.nrows;
        for (int j = 0; j < L_tail; ++j) s.out_rows[i][cols + j] = pb.draft[src][j];
      }
      s.span_b[s.nspans] = b; s.span_start[s.nspans] = cursor; s.span_L[s.nspans] = L_tail; s.nspans++;
      cursor += L_tail; cols += L_tail;
    }
  }
  s.B_out = B_ra;
  s.L_out = cols;
  return cols > 0 && B_ra > 0;
}

// Everything between the forward and the next build (:467-721).  greedy[i * gstride + j] = argmax at column j of row i.
// Returns true when the call has ended (ret / ret_n / first_correct valid).
JF_HD inline bool mb_post_forward(SeqState& s, const PolicyKnobs& kn, const int* greedy, int gstride) {
  const int n = kn.n;
  s.commit_row = 0;
  s.commit_n = 0;
  const int B_out = s.B_out;
  int accepted[kRowsMax];
  int tmp[kPoolEntryMax];
  for (int sp = 0; sp < s.nspans; ++sp) {
    const int b = s.span_b[sp], start = s.span_start[sp], L = s.span_L[sp];
    MBlock& blk = s.blk[b];
    const int B_b = blk.nrows;
    const int nrows = B_b > B_out ? B_b : B_out;
    if (!((B_b == 1 || B_b == nrows) && (B_out == 1 || B_out == nrows))) {
      s.error = PERR_BROADCAST;                                   // reference would fail to broadcast (:482)
      return true;
    }
    for (int i = 0; i < nrows; ++i) {                             // :482-486
      const int* d = blk.draft[B_b > 1 ? i : 0];
      const int* g = greedy + (size_t)(B_out > 1 ? i : 0) * gstride + (start - 1);
      accepted[i] = leading_matches(d + 1, g, L - 1) + 1;
    }
    int best_idx = 0;                                             // :487-491 (first maximum)
    if (b == s.RA)
      for (int i = 1; i < nrows; ++i)
        if (accepted[i] > accepted[best_idx]) best_idx = i;
    const int acc_len_raw = accepted[best_idx];
    if (best_idx >= B_b) continue;                                // draft_rows[best_idx:best_idx+1] empty (:496, :505)
    const int* greedy_row = greedy + (size_t)(B_out > 1 ? best_idx : 0) * gstride + (start - 1);
    if (b == s.RA) s.commit_row = best_idx;                       // KV keeps that row (:500-502)
    const int L_eff = blk.len;
    if (L_eff == 0) continue;
    int draft[kNMax];
    for (int j = 0; j < L_eff; ++j) draft[j] = blk.draft[best_idx][j];
    int acc_len = acc_len_raw;
    bool eos_reached = false;
    if (kn.eos >= 0 && b == s.RA && acc_len > 0) {                // :514-521
      for (int j = 0; j < acc_len; ++j)
        if (draft[j] == kn.eos) { acc_len = j + 1; eos_reached = true; break; }
    }
    const bool has_rejected = acc_len < L_eff;                    // :523
    if (acc_len > 0) {                                            // :526-528
      if (blk.acc_n + acc_len > kNMax) { s.error = PERR_CAPACITY; return true; }
      for (int j = 0; j < acc_len; ++j) blk.acc[blk.acc_n + j] = draft[j];
      blk.acc_n += acc_len;
      blk.total_acc += acc_len;
      if (b == s.RA) s.commit_n = acc_len;
    }
    if (eos_reached && b == s.RA) {                               // :531-547
      mb_final_ret(s);
      s.first_correct = draft[acc_len - 1];
      s.ctx = s.call_ctx0 + s.ret_n;
      return true;
    }
    int nxt;
    if (has_rejected) {                                           // :550-588
      nxt = greedy_row[acc_len - 1 > 0 ? acc_len - 1 : 0];
      blk.nrows = 1;
      blk.len = L_eff - acc_len;
      blk.draft[0][0] = nxt;
      for (int j = acc_len; j < L_eff - 1; ++j) blk.draft[0][1 + j - acc_len] = greedy_row[j];
      if (b == s.RA) {
        int m = 0;                                                // _concat_all_blocks_seq (:387-411), pads stripped
        for (int bb = 0; bb < s.num_blocks; ++bb) {
          const MBlock& cb = s.blk[bb];
          for (int j = 0; j < cb.acc_n; ++j)
            if (kn.pad < 0 || cb.acc[j] != kn.pad) { if (m < kPoolEntryMax) tmp[m++] = cb.acc[j]; }
          if (cb.nrows > 0 && cb.len > 0)
            for (int j = 0; j < cb.len; ++j)
              if (kn.pad < 0 || cb.draft[0][j] != kn.pad) { if (m < kPoolEntryMax) tmp[m++] = cb.draft[0][j]; }
        }
        if (m > 0) pool_push(s, kn, tmp, m);
        const int tail_n = L_eff - 1 - acc_len;
        if (tail_n > 0) pool_push(s, kn, greedy_row + acc_len, tail_n);
      }
      if (b == s.RA && blk.total_acc >= kn.lookahead_min_acc) {
        // _build_candidates (:62-91), nearest=False: newest -> oldest, skipping the very last pool entry
        const int L_out = blk.len;
        int ncand = 0;
        for (int e = s.pool_count - 2; e >= 0; --e) {
          const int* seq = s.pool_tok[e];
          const int sl = s.pool_len[e];
          int pos = -1;
          for (int j = 0; j < sl; ++j) if (seq[j] == nxt) { pos = j; break; }
          if (pos < 0) continue;
          if (1 + ncand >= kRowsMax) { s.error = PERR_CAPACITY; return true; }
          int* cand = blk.draft[1 + ncand];
          int cl = sl - pos;
          if (cl > L_out) cl = L_out;
          for (int j = 0; j < cl; ++j) cand[j] = seq[pos + j];
          for (int j = cl; j < L_out; ++j) cand[j] = blk.draft[0][j];
          ++ncand;
        }
        if (ncand > 1) blk.nrows = 1 + ncand;                     // :579 only if more than one candidate
      }
    } else {                                                      // :590-593
      blk.nrows = 1;
      blk.len = 0;
      nxt = greedy_row[L_eff - 1];
    }
    if (b == s.RA) s.last_next_token = nxt;
    if (kn.eos >= 0 && b == s.RA && s.last_next_token == kn.eos) {   // :599-614
      // out_acc[b] may momentarily hold n+1 tokens here; ret is assembled straight away
      mb_final_ret(s);
      s.ret[s.ret_n++] = s.last_next_token;
      s.first_correct = s.last_next_token;
      s.ctx = s.call_ctx0 + s.ret_n - 1;                          // the appended EOS has no KV
      return true;
    }
  }

  s.ctx = mb_committed_len(s, s.RA);                              // :617-626

  const int newest = s.num_blocks - 1;                            // :629-653 spawn
  if (s.blk[newest].total_acc >= kn.spawn_threshold && s.active_blocks < kn.K) {
    if (kn.pad < 0) { s.error = PERR_NO_PAD; return true; }
    if (s.list_len >= kBlocksMax) { s.error = PERR_BLOCKS; return true; }
    const MBlock& ra = s.blk[s.RA];
    MBlock& nb = s.blk[s.list_len];
    const int L_cur = ra.nrows > 0 ? ra.len : 0;
    nb.nrows = ra.nrows;
    nb.len = L_cur > n ? L_cur : n;
    for (int i = 0; i < ra.nrows; ++i) {
      for (int j = 0; j < L_cur; ++j) nb.draft[i][j] = ra.draft[i][j];
      for (int j = L_cur; j < n; ++j) nb.draft[i][j] = kn.pad;
    }
    nb.acc_n = 0; nb.total_acc = 0; nb.need_reverify = 1;
    s.list_len += 1;
    s.num_blocks += 1;
    s.active_blocks += 1;
  }

  if (s.blk[s.RA].total_acc >= n) {                               // :656-716 promote
    for (int b = 0; b < s.num_blocks; ++b) {
      MBlock& pb = s.blk[b];
      if (pb.need_reverify && pb.total_acc > 0) {
        const int full = pb.acc_n + pb.len;
        if (full != n) { s.error = PERR_PROMOTE_LEN; return true; }   // assert at :667
        int q_full[kNMax];
        for (int j = 0; j < pb.acc_n; ++j) q_full[j] = pb.acc[j];
        for (int j = 0; j < pb.len; ++j) q_full[pb.acc_n + j] = pb.draft[0][j];
        pb.acc_n = 0; pb.total_acc = 0;
        pb.nrows = 1; pb.len = n;
        pb.draft[0][0] = s.last_next_token;
        for (int j = 1; j < n; ++j) pb.draft[0][j] = q_full[j];
        pb.need_reverify = 0;
        s.RA = b;
        s.ctx = mb_committed_len(s, s.RA);
        s.active_blocks -= 1;
        s.num_blocks -= 1;
        break;
      }
      s.active_blocks -= 1;
    }
  }

  bool all_done = true;                                           // :719-721
  for (int b = 0; b < s.num_blocks; ++b) if (s.blk[b].total_acc < n) { all_done = false; break; }
  if (all_done) {
    mb_final_ret(s);
    s.first_correct = s.last_next_token;
    s.ctx = s.call_ctx0 + s.ret_n;
    return true;
  }
  return false;
}

// Loop head (:414-419): returns true when the call ended without a forward.
JF_HD inline bool mb_pre_forward(SeqState& s, const PolicyKnobs& kn) {
  if (s.iters >= kn.max_iteration_count) {
    mb_final_ret(s);
    s.first_correct = s.last_next_token;
    s.ctx = s.call_ctx0 + s.ret_n;
    return true;
  }
  s.iters += 1;
  if (!mb_build_rows(s, kn)) {
    mb_final_ret(s);
    s.first_correct = s.last_next_token;
    s.ctx = s.call_ctx0 + s.ret_n;
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
// single block (improved.py:140-276); state lives in blk[0]: draft[0][:len] = `out`, acc = accepted_n_gram
// ---------------------------------------------------------------------------------------------------------------
JF_HD inline void sb_begin_call(SeqState& s, const PolicyKnobs& kn, const int* input_ids) {
  MBlock& b0 = s.blk[0];
  b0.nrows = 1; b0.len = kn.n; b0.acc_n = 0; b0.total_acc = 0; b0.need_reverify = 0;
  for (int j = 0; j < kn.n; ++j) b0.draft[0][j] = input_ids[j];
  s.num_blocks = 1; s.active_blocks = 1; s.list_len = 1; s.RA = 0;
  s.call_ctx0 = s.ctx;
  s.iters = 0;
  s.ret_n = 0;
  s.in_call = 1;
  s.pool_count = 0;
}

JF_HD inline bool sb_pre_forward(SeqState& s, const PolicyKnobs& kn) {
  MBlock& b0 = s.blk[0];
  if (b0.total_acc >= kn.n) {                                      // while total_accepted < n (:150)
    int m = b0.acc_n < kn.n ? b0.acc_n : kn.n;
    for (int j = 0; j < m; ++j) s.ret[j] = b0.acc[j];
    s.ret_n = m;
    return true;
  }
  s.iters += 1;
  for (int j = 0; j < b0.len; ++j) s.out_rows[0][j] = b0.draft[0][j];
  s.B_out = 1;
  s.L_out = b0.len;
  s.nspans = 0;
  return false;
}

JF_HD inline bool sb_post_forward(SeqState& s, const PolicyKnobs& kn, const int* greedy, int gstride) {
  (void)gstride;
  MBlock& b0 = s.blk[0];
  const int L = b0.len;
  const int* out = b0.draft[0];
  const int raw = leading_matches(out + 1, greedy, L - 1) + 1;      // :197-202
  int num_accepted = raw;
  bool eos_in = false;
  if (kn.eos >= 0)                                                  // :206-211
    for (int j = 0; j < raw; ++j) if (out[j] == kn.eos) { num_accepted = j + 1; eos_in = true; break; }
  for (int j = 0; j < num_accepted; ++j) if (b0.acc_n < kNMax) b0.acc[b0.acc_n++] = out[j];   // :214-216
  b0.total_acc += num_accepted;
  s.commit_row = 0;
  if (eos_in) {                                                     // :219-227
    s.commit_n = num_accepted;
    s.ctx += num_accepted;
    for (int j = 0; j < b0.total_acc; ++j) s.ret[j] = b0.acc[j];
    s.ret_n = b0.total_acc;
    s.first_correct = kn.eos;
    return true;
  }
  s.commit_n = raw;                                                 // rejected tail deleted (:233) / all kept
  s.ctx += raw;
  if (raw < L) {                                                    // :231-255
    const int next_token = greedy[raw - 1];
    if (kn.eos >= 0 && next_token == kn.eos) {                      // :238-247
      if (b0.acc_n < kNMax) b0.acc[b0.acc_n++] = next_token;
      b0.total_acc += 1;
      for (int j = 0; j < b0.total_acc && j < kNMax; ++j) s.ret[j] = b0.acc[j];
      s.ret_n = b0.total_acc;
      s.first_correct = next_token;
      return true;
    }
    int nd[kNMax];
    nd[0] = next_token;
    for (int j = raw; j < L - 1; ++j) nd[1 + j - raw] = greedy[j];
    b0.len = L - raw;
    for (int j = 0; j < b0.len; ++j) b0.draft[0][j] = nd[j];
    s.first_correct = next_token;
  } else {                                                          // :258-273
    const int next_token = greedy[L - 1];
    b0.total_acc += 1;                       // the write at index n is a no-op on the width-n accepted_n_gram
    b0.len = 0;
    s.first_correct = next_token;
    if (kn.eos >= 0 && next_token == kn.eos) {
      int m = b0.acc_n < kn.n ? b0.acc_n : kn.n;
      for (int j = 0; j < m; ++j) s.ret[j] = b0.acc[j];
      s.ret_n = m;
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
// driver (MR_humaneval.py:152-240) + dispatcher.  `gen` is the caller's token buffer [prompt | generated] of
// capacity gen_cap.  advance() consumes the greedy table of the forward just done (null on the first entry), then
// runs the state machine until it either needs another forward (returns true; rows are in s.out_rows, B_out x L_out,
// to be placed at positions ctx + column) or the sequence is finished (returns false).
// On return s.commit_row / s.commit_n describe which in-flight tokens of the forward JUST DONE become committed KV.
// ---------------------------------------------------------------------------------------------------------------
JF_HD inline void driver_finish_call(SeqState& s, const PolicyKnobs& kn, int* gen, int gen_cap, int* iters_log,
                                     int iters_cap) {
  const int prev_len = s.gen_len;
  for (int j = 0; j < s.ret_n && s.gen_len < gen_cap; ++j) {
    if (kn.eos >= 0 && s.ret[j] == kn.eos) s.eos_seen = 1;
    gen[s.gen_len++] = s.ret[j];
  }
  if (iters_log && s.calls < iters_cap) iters_log[s.calls] = s.iters;
  s.total_iters += s.iters;
  s.calls += 1;
  const int added = s.gen_len - prev_len;
  if (added > 0) s.total_new += added;
  s.in_call = 0;
  if (s.error != PERR_NONE) { s.finished = 1; s.stop_reason = STOP_ERROR; }
}

JF_HD inline bool hf_advance(SeqState& s, const PolicyKnobs& kn, const DraftSource& src, const int* greedy,
                             int gstride, int* gen, int gen_cap, int* iters_log, int iters_cap) {
  const bool multi = kn.mode == MODE_HF_MULTI;
  bool have_greedy = greedy != nullptr && s.in_call;
  s.commit_n = 0;
  s.commit_row = 0;
  int saved_commit_row = 0, saved_commit_n = 0;
  while (!s.finished) {
    if (s.in_call) {
      bool ended;
      if (have_greedy) {
        ended = multi ? mb_post_forward(s, kn, greedy, gstride) : sb_post_forward(s, kn, greedy, gstride);
        saved_commit_row = s.commit_row;
        saved_commit_n = s.commit_n;
        have_greedy = false;
        if (!ended) continue;
      } else {
        ended = multi ? mb_pre_forward(s, kn) : sb_pre_forward(s, kn);
        if (!ended) { s.commit_row = saved_commit_row; s.commit_n = saved_commit_n; return true; }
      }
      driver_finish_call(s, kn, gen, gen_cap, iters_log, iters_cap);
      continue;
    }
    // ---- loop head of the driver (:152-170)
    if (s.eos_seen) { s.finished = 1; s.stop_reason = STOP_EOS; break; }
    if (s.total_new >= kn.max_new_tokens) { s.finished = 1; s.stop_reason = STOP_MAX_NEW; break; }
    if (s.calls >= kn.max_calls) { s.finished = 1; s.stop_reason = STOP_MAX_CALLS; break; }
    int input_ids[kNMax];
    if (s.calls == 1) {                                             // :206-208
      for (int j = 0; j < kn.n; ++j) input_ids[j] = src.first_call_tokens[j];
    } else {                                                        // :209-215
      if (kn.n > 1 && (src.mt == nullptr || src.mt_words < 625)) {
        s.error = PERR_DRAFT_SOURCE; s.finished = 1; s.stop_reason = STOP_ERROR; break;
      }
      input_ids[0] = s.first_correct;
      for (int j = 0; j < kn.n - 1; ++j) input_ids[1 + j] = gen[mt_randbelow(src.mt, s.gen_len)];
    }
    if (multi) mb_begin_call(s, kn, input_ids); else sb_begin_call(s, kn, input_ids);
  }
  s.commit_row = saved_commit_row;
  s.commit_n = saved_commit_n;
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
// engine convention (inference_engine/engine/jacobi_decoding.py:303-445): fixed length L, draft[0] = last committed
// token whose KV is recomputed every iteration; the forward covers positions [S-1, S+L-2].  State: blk[0].draft[0].
// `ctx` here is S-1 (tokens whose KV is final), gen holds token_ids.
// ---------------------------------------------------------------------------------------------------------------
JF_HD inline int eng_next_pad(SeqState& s, const PolicyKnobs& kn, const DraftSource& src) {
  if (src.pad_stream == nullptr || s.pad_cursor >= src.pad_stream_len) { s.error = PERR_DRAFT_SOURCE; return 0; }
  (void)kn;
  return src.pad_stream[s.pad_cursor++];
}

JF_HD inline bool eng_advance(SeqState& s, const PolicyKnobs& kn, const DraftSource& src, const int* greedy,
                              int* gen, int gen_cap) {
  const int L = kn.n;
  MBlock& b0 = s.blk[0];
  int* q = b0.draft[0];
  s.commit_n = 0;
  s.commit_row = 0;
  if (L <= 1) { s.finished = 1; s.stop_reason = STOP_ERROR; return false; }
  if (greedy != nullptr && s.in_call) {
    // verify (:359-369): greedy has L-1 useful entries (argmax at positions 0..L-2 verifies draft[1:])
    int acc_len = leading_matches(q + 1, greedy, L - 1) + 1;
    if (acc_len > L) acc_len = L;
    bool eos_reached = false;
    if (kn.eos >= 0 && acc_len > 1)
      for (int j = 1; j < acc_len; ++j) if (q[j] == kn.eos) { acc_len = j + 1; eos_reached = true; break; }
    int num_spec = acc_len - 1;
    for (int j = 1; j < acc_len && s.gen_len < gen_cap; ++j) gen[s.gen_len++] = q[j];      // :376-381
    if (acc_len == 1) {                                                                     // :383-392
      const int nt = greedy[0];
      if (s.gen_len < gen_cap) gen[s.gen_len++] = nt;
      num_spec = 1;
      if (kn.eos >= 0 && nt == kn.eos) eos_reached = true;
    }
    s.total_new += num_spec;
    // Tokens whose KV is final after this iteration = len(token_ids) - 1 (the new seed is recomputed next iteration,
    // invariant at :404-407); in-flight token j sits at position ctx + j, so the first (new_ctx - ctx) tokens commit.
    const int new_ctx = s.gen_len - 1;
    s.commit_n = new_ctx - s.ctx;
    s.ctx = new_ctx;
    if (eos_reached) { s.finished = 1; s.stop_reason = STOP_EOS; s.eos_seen = 1; return false; }
    if (s.total_new >= kn.max_new_tokens) { s.finished = 1; s.stop_reason = STOP_MAX_NEW; return false; }
    // next draft (:415-436)
    int d[kNMax];
    int dl = 0;
    d[dl++] = gen[s.gen_len - 1];
    if (acc_len < L) {
      const int from = (acc_len == 1) ? 1 : acc_len - 1;
      for (int j = from; j < L - 1 && dl < L; ++j) d[dl++] = greedy[j];
    } else {
      d[dl++] = greedy[L - 2];
    }
    while (dl < L) d[dl++] = eng_next_pad(s, kn, src);
    for (int j = 0; j < L; ++j) q[j] = d[j];
  } else if (!s.in_call) {
    // first draft (:332-347 with a prefill draft, else _init_draft :142-171)
    q[0] = gen[s.gen_len - 1];
    int dl = 1;
    if (src.first_call_tokens != nullptr)
      for (int j = 0; j < L - 1; ++j) q[dl++] = src.first_call_tokens[j];
    while (dl < L) q[dl++] = eng_next_pad(s, kn, src);
    s.in_call = 1;
    s.ctx = s.gen_len - 1;
  }
  if (s.error != PERR_NONE) { s.finished = 1; s.stop_reason = STOP_ERROR; return false; }
  if (s.iters >= kn.max_iteration_count) { s.finished = 1; s.stop_reason = STOP_MAX_ITERS; return false; }
  s.iters += 1;
  s.total_iters += 1;
  q[0] = gen[s.gen_len - 1];                                        // :197-199
  for (int j = 0; j < L; ++j) s.out_rows[0][j] = q[j];
  s.B_out = 1;
  s.L_out = L;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Counter-based random numbers (Philox4x32-10, Salmon et al. 2011): every draw of the non-greedy path is a pure function
// of (key, counter), so the device, the host harness and the Python oracle produce the same stream without shared state.
// The reference draws from torch's global CUDA generator (jacobi_decoding_nongreedy.py:322, :128-131), which is neither
// seeded nor reproducible across batch compositions; only the DISTRIBUTION of its draws is part of its behaviour.
// ---------------------------------------------------------------------------------------------------------------------
struct Philox4 { unsigned int x, y, z, w; };
JF_HD inline Philox4 philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3, unsigned int k0,
                                   unsigned int k1) {
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const unsigned int n0 = (unsigned int)(p1 >> 32) ^ c1 ^ k0;
    const unsigned int n1 = (unsigned int)p1;
    const unsigned int n2 = (unsigned int)(p0 >> 32) ^ c3 ^ k1;
    const unsigned int n3 = (unsigned int)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
enum RngStream : unsigned int { RNG_GUMBEL = 1u, RNG_ACCEPT = 2u };
// 24-bit uniforms: [0, 1) for the accept test (u < p), (0, 1) for the Gumbel noise -log(-log(u))
JF_HD inline float u01_closed_open(unsigned int x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
JF_HD inline float u01_open(unsigned int x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// uniform of the accept test of in-flight row `row` (global row index of the forward) of forward `fwd`
JF_HD inline float ng_accept_uniform(const PolicyKnobs& kn, int row, int fwd) {
  return u01_closed_open(philox4x32_10((unsigned int)row, 0u, (unsigned int)(fwd + kn.rng_offset), RNG_ACCEPT, kn.seed_lo,
                                       kn.seed_hi).x);
}

// Non-greedy engine policy: one call = one Jacobi iteration of JacobiDecoderNonGreedy._generate_single (:385-470).
// Per in-flight row t (0 <= t < L-1, verifying draft[t+1]) the forward supplies, besides the argmax `greedy[t]`:
//   px[t]     target probability of the drafted token, softmax(logits[t] / T)[draft[t+1]]
//   bonus[t]  a token drawn from softmax(logits[t] / T) conditioned on != draft[t+1]  (:134-153; the residual of the
//             delta proposal).  Both come out of the LM-head epilogue; only the accept uniforms are drawn here.
JF_HD inline bool eng_ng_advance(SeqState& s, const PolicyKnobs& kn, const DraftSource& src, const int* greedy,
                                 const float* px, const int* bonus, int row0, int fwd, int* gen, int gen_cap) {
  const int L = kn.n;
  MBlock& b0 = s.blk[0];
  int* q = b0.draft[0];
  s.commit_n = 0;
  s.commit_row = 0;
  if (L <= 1) { s.finished = 1; s.stop_reason = STOP_ERROR; return false; }
  if (greedy != nullptr && s.in_call) {
    // verify (:316-347): sequential accept / reject along the block
    int committed = 0;
    bool eos_reached = false;
    for (int t = 0; t < L - 1; ++t) {
      const int proposed = q[t + 1];
      const float u = ng_accept_uniform(kn, row0 + t, fwd);
      int tok = proposed;
      const bool accept = u < px[t];
      if (!accept) tok = bonus[t];
      if (s.gen_len < gen_cap) gen[s.gen_len++] = tok;
      ++committed;
      if (kn.eos >= 0 && tok == kn.eos) eos_reached = true;
      if (!accept || eos_reached) break;
    }
    s.total_new += committed;
    const int new_ctx = s.gen_len - 1;                       // the last committed token is the next seed (recomputed)
    s.commit_n = new_ctx - s.ctx;
    s.ctx = new_ctx;
    if (eos_reached) { s.finished = 1; s.stop_reason = STOP_EOS; s.eos_seen = 1; return false; }
    if (s.total_new >= kn.max_new_tokens) { s.finished = 1; s.stop_reason = STOP_MAX_NEW; return false; }
    // next draft (:446-466): seed + greedy fill from the current logits + random padding
    const int acc_len = 1 + committed;
    int d[kNMax];
    int dl = 0;
    d[dl++] = gen[s.gen_len - 1];
    if (acc_len < L) {
      for (int j = acc_len - 1; j < L - 1 && dl < L; ++j) d[dl++] = greedy[j];
    } else {
      d[dl++] = greedy[L - 2];
    }
    while (dl < L) d[dl++] = eng_next_pad(s, kn, src);
    for (int j = 0; j < L; ++j) q[j] = d[j];
  } else if (!s.in_call) {
    q[0] = gen[s.gen_len - 1];                                // _init_draft (:216-236): seed + random tokens
    int dl = 1;
    if (src.first_call_tokens != nullptr)
      for (int j = 0; j < L - 1; ++j) q[dl++] = src.first_call_tokens[j];
    while (dl < L) q[dl++] = eng_next_pad(s, kn, src);
    s.in_call = 1;
    s.ctx = s.gen_len - 1;
  }
  if (s.error != PERR_NONE) { s.finished = 1; s.stop_reason = STOP_ERROR; return false; }
  if (s.iters >= kn.max_iteration_count) { s.finished = 1; s.stop_reason = STOP_MAX_ITERS; return false; }
  s.iters += 1;
  s.total_iters += 1;
  q[0] = gen[s.gen_len - 1];
  for (int j = 0; j < L; ++j) s.out_rows[0][j] = q[j];
  s.B_out = 1;
  s.L_out = L;
  return true;
}

JF_HD inline void seq_state_init(SeqState& s, int prompt_len, int ctx, int first_correct) {
  s.calls = 1; s.total_new = 0; s.gen_len = prompt_len; s.prompt_len0 = prompt_len; s.finished = 0;
  s.stop_reason = STOP_NONE; s.error = PERR_NONE; s.first_correct = first_correct; s.in_call = 0; s.eos_seen = 0;
  s.total_iters = 0; s.ctx = ctx; s.call_ctx0 = ctx; s.iters = 0; s.num_blocks = 1; s.active_blocks = 1;
  s.list_len = 1; s.RA = 0; s.last_next_token = -1; s.ret_n = 0; s.pool_count = 0; s.B_out = 0; s.L_out = 0;
  s.nspans = 0; s.commit_row = 0; s.commit_n = 0; s.eng_have_draft = 0; s.pad_cursor = 0;
}

}  // namespace jf
