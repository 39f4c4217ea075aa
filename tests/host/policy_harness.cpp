// policy_harness.cpp — host build of the PRODUCT policy code (jacobiforcing_b200/csrc/jf_policy.h, the same header
// the persistent CUDA kernel compiles) behind a tiny C entry point, so tests can replay the reference's golden
// forwards (tests/golden/*.json) through it on a CPU-only box.  Test infrastructure; not part of the shipped library.
// Build: g++ -O2 -shared -fPIC -I jacobiforcing_b200/csrc tests/host/policy_harness.cpp -o build/libjf_policy_host.so
#include <cstring>
#include <cstdlib>
#include "jf_policy.h"
#include "jf_sched.h"

using namespace jf;

extern "C" {

// Forward table: forward f has B[f] rows of L[f] tokens starting at rows_off[f] in rows_flat / greedy_flat.
// Returns 0 on success, otherwise 100 + index of the first forward whose (kv, rows) differ from the table,
// or 1 when the policy asks for more forwards than recorded.
int jf_policy_replay(const int* knobs_i /*12 ints in PolicyKnobs order*/, const int* prompt, int prompt_len, int ctx0,
                     int first_correct, const int* first_call_tokens, unsigned int* mt_state, int mt_words,
                     const int* pad_stream, int pad_stream_len, int nfw, const int* kv_before, const int* fB,
                     const int* fL, const long long* rows_off, const int* rows_flat, const int* greedy_flat,
                     int* gen_out, int gen_cap, int* iters_log, int iters_cap, int* result /*[8]*/) {
  PolicyKnobs kn;
  std::memcpy(&kn, knobs_i, sizeof(PolicyKnobs));
  SeqState* s = (SeqState*)std::calloc(1, sizeof(SeqState));
  seq_state_init(*s, prompt_len, ctx0, first_correct);
  for (int i = 0; i < prompt_len && i < gen_cap; ++i) gen_out[i] = prompt[i];
  DraftSource src{first_call_tokens, mt_state, mt_words, pad_stream, pad_stream_len};
  int f = 0, rc = 0;
  const int* greedy = nullptr;
  int gstride = 0;
  int* gbuf = (int*)std::malloc(sizeof(int) * kRowsMax * kOutColsMax);
  while (true) {
    bool more;
    if (kn.mode == MODE_ENGINE) more = eng_advance(*s, kn, src, greedy, gen_out, gen_cap);
    else more = hf_advance(*s, kn, src, greedy, gstride, gen_out, gen_cap, iters_log, iters_cap);
    if (!more) break;
    if (f >= nfw) { rc = 1; break; }
    bool same = (kv_before[f] == s->ctx) && (fB[f] == s->B_out) && (fL[f] == s->L_out);
    if (same)
      for (int i = 0; i < s->B_out && same; ++i)
        for (int j = 0; j < s->L_out; ++j)
          if (rows_flat[rows_off[f] + (long long)i * fL[f] + j] != s->out_rows[i][j]) { same = false; break; }
    if (!same) { rc = 100 + f; break; }
    gstride = fL[f];
    std::memcpy(gbuf, greedy_flat + rows_off[f], sizeof(int) * (size_t)fB[f] * fL[f]);
    greedy = gbuf;
    ++f;
  }
  result[0] = s->gen_len; result[1] = s->calls; result[2] = s->stop_reason; result[3] = s->total_new;
  result[4] = s->error; result[5] = f; result[6] = s->total_iters; result[7] = s->ctx;
  std::free(gbuf);
  std::free(s);
  return rc;
}

// Non-greedy engine policy (MODE_ENGINE_NG): forward f supplies px / bonus / greedy for its L-1 verify rows (row0 = 0);
// rows_out[f][L] receives the draft the policy asked to forward.  Returns the number of forwards consumed (<= nfw).
int jf_policy_replay_ng(const int* knobs_i, const int* prompt, int prompt_len, int ctx0, const int* pad_stream,
                        int pad_stream_len, int nfw, const float* px_flat, const int* bonus_flat, const int* greedy_flat,
                        int* rows_out, int* ctx_out, int* gen_out, int gen_cap, int* result /*[8]*/) {
  PolicyKnobs kn;
  std::memcpy(&kn, knobs_i, sizeof(PolicyKnobs));
  SeqState* s = (SeqState*)std::calloc(1, sizeof(SeqState));
  seq_state_init(*s, prompt_len, ctx0, 0);
  for (int i = 0; i < prompt_len && i < gen_cap; ++i) gen_out[i] = prompt[i];
  DraftSource src{nullptr, nullptr, 0, pad_stream, pad_stream_len};
  const int L = kn.n;
  int f = 0;
  while (true) {
    const bool have = f > 0;
    const bool more = eng_ng_advance(*s, kn, src, have ? greedy_flat + (size_t)(f - 1) * (L - 1) : nullptr,
                                     have ? px_flat + (size_t)(f - 1) * (L - 1) : nullptr,
                                     have ? bonus_flat + (size_t)(f - 1) * (L - 1) : nullptr, 0, f - 1, gen_out, gen_cap);
    if (!more || f >= nfw) break;
    for (int j = 0; j < L; ++j) rows_out[(size_t)f * L + j] = s->out_rows[0][j];
    ctx_out[f] = s->ctx;
    ++f;
  }
  result[0] = s->gen_len; result[1] = s->calls; result[2] = s->stop_reason; result[3] = s->total_new;
  result[4] = s->error; result[5] = f; result[6] = s->total_iters; result[7] = s->ctx;
  std::free(s);
  return f;
}

void jf_philox(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3, unsigned int k0, unsigned int k1,
               unsigned int* out4) {
  const Philox4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
float jf_ng_accept_uniform(const int* knobs_i, int row, int fwd) {
  PolicyKnobs kn;
  std::memcpy(&kn, knobs_i, sizeof(PolicyKnobs));
  return ng_accept_uniform(kn, row, fwd);
}

int jf_policy_state_bytes() { return (int)sizeof(SeqState); }
int jf_policy_knob_ints() { return (int)(sizeof(PolicyKnobs) / sizeof(int)); }

// scheduling arithmetic shared with the kernel (jf_sched.h)
int jf_sched_split_k0(int split, int splits, int k_chunks) { return split_k0(split, splits, k_chunks); }
int jf_sched_split_k1(int split, int splits, int k_chunks) { return split_k1(split, splits, k_chunks); }
int jf_sched_tile_feeds_split(int tile, int split, int splits, int k_chunks) {
  return df_tile_feeds_split(tile, split, splits, k_chunks) ? 1 : 0;
}
int jf_sched_tiles_needed(int k0, int k1) { return df_tiles_needed(k0, k1); }

}  // extern "C"
