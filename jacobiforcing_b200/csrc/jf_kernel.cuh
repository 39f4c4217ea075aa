// jf_kernel.cuh — the persistent cooperative kernel: one launch runs Jacobi iterations (Qwen2 forward over the
// in-flight draft rows -> fused LM-head argmax -> accept / shift / re-draft policy -> KV commit) until every sequence
// reached its fixed point / EOS / token budget.  No host round trip between iterations.
//
// One CTA per SM, 224 threads, roles per warp (see jf_gemm.cuh).  Stages are separated by a grid barrier built on a
// monotonic global counter: barrier k is complete when counter >= k * gridDim.x.  Only the worker warps arrive; the
// weight producer never waits on it inside a forward (weights do not depend on activations), the X producer waits
// for the barrier that publishes its operand, the MMA issuer is driven by the shared-memory ring alone.
#pragma once
#include "jf_stages.cuh"

namespace jf {

// fence = false: the caller reads what the counter publishes ONLY through the async proxy (bulk copies served by L2, never
// by this SM's L1) and issues those copies after the successful probe, so it needs neither the L1 invalidation nor the
// ordering of the acquire fence
template <bool kSleep = true>
__device__ __forceinline__ void wait_epoch(const unsigned int* counter, unsigned int target, bool sys = false, bool fence = true) {
  if (ld_relaxed_gpu_u32(counter) < target) {
    const long long t0 = clock64();
    while (ld_relaxed_gpu_u32(counter) < target) {
      if (kSleep) __nanosleep(40);
      if (clock64() - t0 > JF_SPIN_LIMIT_CYCLES) jf_trap_timeout();
    }
  }
  // one acquire (one L1 invalidation) per barrier, not one per probe; system scope when the data published before the
  // barrier may have been written by a tensor-parallel peer (finished rows copied out of the staging buffer are local
  // writes, but the flag barriers of the prefill path order peer reads)
  if (!fence) return;
  if (sys) fence_acq_rel_sys(); else fence_acq_rel_gpu();
}

// Grid barrier of the worker warps.  Writer side: every worker's global stores are ordered before the CTA barrier,
// and thread 0's release-add after it is cumulative over them (the cooperative-groups grid.sync pattern); the
// per-thread proxy fence makes those stores visible to bulk-copy (async proxy) reads issued by other CTAs.
// Barrier k is complete when counter >= k * nctas + (cross-GPU barriers so far): a cross-GPU barrier adds ONE extra
// arrival, made by CTA 0 after it has exchanged flags with every tensor-parallel peer through NVLink peer memory.
struct BarState {
  unsigned int epoch = 0;      // barriers passed
  unsigned int xcount = 0;     // of which cross-GPU
  __device__ __forceinline__ unsigned int target(unsigned int nctas) const { return epoch * nctas + xcount; }
};

__device__ __forceinline__ void grid_barrier(const DevCtx& c, const WorkerView& w, BarState& bs, bool profiled = true,
                                             bool tma_consumer = true, bool xgpu = false, int stage_id = 0) {
  const bool prof_act = profiled && w.prof_on() != 0;
  if (tma_consumer) fence_proxy_async_all();   // only stages whose output a bulk copy (async proxy) will read need it
  workers_sync();
  const unsigned int local_target = (bs.epoch + 1) * (unsigned int)w.nctas() + bs.xcount;
  bs.epoch += 1;
  if (xgpu) bs.xcount += 1;
  if (w.wtid() == 0) {
    long long t_arrive = 0;
    if (prof_act) t_arrive = clock64();
    prof_mark(c, w, 900);
    red_release_gpu_add_u32(c.bar_counter, 1u);
    if (xgpu && w.cta() == 0) {
      wait_epoch<false>(c.bar_counter, local_target);          // every CTA of this GPU has arrived
      fence_acq_rel_sys();
      const unsigned long long val = c.xflag_base + bs.xcount;
      for (int r = 0; r < c.tp_size; ++r)
        if (r != c.tp_rank) st_relaxed_sys_u64(c.peer_xflag[r] + c.tp_rank, val);
      const long long t0 = clock64();
      for (int r = 0; r < c.tp_size; ++r) {
        if (r == c.tp_rank) continue;
        while (ld_relaxed_sys_u64(c.xflag + r) < val)
          if (clock64() - t0 > JF_XSPIN_LIMIT_CYCLES) jf_trap_timeout();
      }
      fence_acq_rel_sys();
      red_release_gpu_add_u32(c.bar_counter, 1u);              // the extra arrival releases this GPU's CTAs
    }
    if (c.early_fence && !xgpu) {
      // The acquire fence's job here is to drop this SM's L1 copies of what other CTAs wrote during the stage.  Issued right
      // behind the arrival it costs its L2 round trip while the other CTAs are still arriving, instead of after the last
      // of them: no worker thread of this CTA loads anything between here and the release (they are parked at the named
      // barrier below, the probes bypass L1), so L1 cannot pick the stale lines up again before the counter completes
      fence_acq_rel_gpu();
      wait_epoch<false>(c.bar_counter, bs.target(w.nctas()), false, false);
    } else {
      wait_epoch<false>(c.bar_counter, bs.target(w.nctas()), xgpu);    // critical path: spin, do not sleep
    }
    if (prof_act) {                                       // (stage id, arrival of CTA 0, release of the barrier)
      const int prof_n = w.shared()->prof_n;
      if (prof_n + 3 <= c.prof_cap) {
        c.prof[prof_n] = stage_id;
        c.prof[prof_n + 1] = t_arrive;
        c.prof[prof_n + 2] = clock64();
        w.shared()->prof_n = prof_n + 3;
      }
    }
    prof_mark(c, w, 901);
  }
  workers_sync();
}

// ---------------------------------------------------------------------------------------------------------------
// policy stages (CTA 0, worker warps)
// ---------------------------------------------------------------------------------------------------------------
// best key over the tensor-parallel ranks (each holds the argmax of its vocabulary slice; the packed key orders by
// value, then by LOWEST global index, so the merge keeps torch.argmax's first-maximum rule)
__device__ __forceinline__ unsigned long long merged_key(const DevCtx& c, int kbuf, int t) {
  const size_t off = (size_t)kbuf * c.T_cap + t;
  if (c.tp_size == 1) return ld_weak(&c.argmax_keys[off]);
  unsigned long long best = 0ull;
  for (int r = 0; r < c.tp_size; ++r) {
    const unsigned long long k = (r == c.tp_rank) ? ld_weak(&c.argmax_keys[off]) : ld_relaxed_sys_u64(c.peer_keys[r] + off);
    best = k > best ? k : best;
  }
  return best;
}

__device__ __noinline__ void policy_stage_plan(const DevCtx& c, const WorkerShared* ws) {
  const WorkerCtx w = load_wctx(ws);
  PlanHdr* hdr = c.hdr;
  const int kbuf = ld_weak(&hdr->iter) & 1;
  const int n_lm = ld_weak(&hdr->n_lm_tasks);
  for (int i = 0; i < n_lm; ++i) {
    const int r0 = ld_weak(&hdr->lm_row0[i]), rn = ld_weak(&hdr->lm_rows[i]);
    for (int t = r0 + w.wtid(); t < r0 + rn; t += kWorkerThreads) {
      c.greedy[t] = argmax_key_index(merged_key(c, kbuf, t));
    }
  }
  if (w.wtid() == 0) {
    hdr->n_commit = ld_weak(&hdr->n_commit_after);
    hdr->iter = ld_weak(&hdr->iter) + 1;
    hdr->rows_total = ld_weak(&hdr->rows_total) + ld_weak(&hdr->T);
    hdr->done = 1;
  }
}

__device__ __noinline__ void policy_stage_jacobi(const DevCtx& c, const WorkerShared* ws, bool first, int max_forwards) {
  const WorkerCtx w = load_wctx(ws);
  PlanHdr* hdr = c.hdr;
  const int Tprev = first ? 0 : ld_weak(&hdr->T);
  const int kbuf = first ? 0 : (ld_weak(&hdr->iter) & 1);
  if (!first) {
    for (int t = w.wtid(); t < Tprev; t += kWorkerThreads) c.greedy[t] = argmax_key_index(merged_key(c, kbuf, t));
    // the OTHER buffer is cleared for the next forward; its last readers (the previous policy stage of every rank)
    // are separated from this point by that forward's cross-GPU barriers
    for (int t = w.wtid(); t < c.T_cap; t += kWorkerThreads) c.argmax_keys[(size_t)(kbuf ^ 1) * c.T_cap + t] = 0ull;
  }
  if (!first && c.ng_mode) {
    // non-greedy verify inputs: merge the (max, sum exp) pairs of every CTA of every rank (rank order, then CTA order: the
    // same on every rank), turn the drafted token's scaled logit into its probability, decode the residual sample (best
    // Gumbel key over the ranks' vocabulary slices), and re-arm the OTHER half of the accumulators for the next forward
    const int parts = c.ng_parts;
    const size_t half = (size_t)kbuf * c.T_cap;
    for (int t = w.wtid(); t < Tprev; t += kWorkerThreads) {
      float M = -INFINITY, S = 0.f, lp = -INFINITY;
      unsigned long long gk = 0ull;
      for (int r = 0; r < c.tp_size; ++r) {
        const float2* ms = peer_ptr(c, r, c.ng_ms) + (size_t)kbuf * parts * c.T_cap;
        for (int p0 = 0; p0 < parts; p0 += 8) {
          float2 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = (p0 + i < parts) ? ld_weak(&ms[(size_t)(p0 + i) * c.T_cap + t]) : make_float2(-INFINITY, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float nm = fmaxf(M, v[i].x);
            if (nm > -INFINITY) { S = S * __expf(M - nm) + v[i].y * __expf(v[i].x - nm); M = nm; }
          }
        }
        lp = fmaxf(lp, ld_weak(peer_ptr(c, r, c.ng_prop) + half + t));          // exactly one rank holds the drafted token
        const unsigned long long k = ld_weak(peer_ptr(c, r, c.ng_gumbel) + half + t);
        gk = k > gk ? k : gk;
      }
      c.ng_px[t] = (S > 0.f && lp > -INFINITY) ? __expf(lp - M) / S : 0.f;
      c.ng_bonus[t] = argmax_key_index(gk);
    }
    for (int t = w.wtid(); t < c.T_cap; t += kWorkerThreads) {
      c.ng_prop[(size_t)(kbuf ^ 1) * c.T_cap + t] = -INFINITY;
      c.ng_gumbel[(size_t)(kbuf ^ 1) * c.T_cap + t] = 0ull;
    }
  }
  __threadfence_block();
  workers_sync();
  if (w.lane() == 0) {
    for (int s = w.wwarp(); s < c.n_seqs; s += 4) {
      SeqState& st = c.seq_state[s];
      SeqSlot& sl = c.seq_slot[s];
      sl.had_forward = (sl.active && !first) ? 1 : 0;
      sl.commit_n = 0;
      if (st.finished) { sl.active = 0; continue; }
      const PolicyKnobs& kn = c.knobs[s];
      const DraftSource& src = c.draft_src[s];
      const int* greedy = sl.had_forward ? (c.greedy + sl.row0) : nullptr;
      bool more;
      if (kn.mode == MODE_ENGINE)
        more = eng_advance(st, kn, src, greedy, c.gen + (size_t)s * c.gen_cap, c.gen_cap);
      else if (kn.mode == MODE_ENGINE_NG)
        more = eng_ng_advance(st, kn, src, greedy, c.ng_px + sl.row0, c.ng_bonus + sl.row0, sl.row0,
                              first ? 0 : ld_weak(&hdr->iter), c.gen + (size_t)s * c.gen_cap, c.gen_cap);
      else
        more = hf_advance(st, kn, src, greedy, sl.L, c.gen + (size_t)s * c.gen_cap, c.gen_cap,
                          c.iters_log + (size_t)s * c.iters_cap, c.iters_cap);
      if (sl.had_forward) {
        sl.commit_src_row = sl.row0 + st.commit_row * sl.L;
        sl.commit_n = st.commit_n;
        sl.commit_pos = sl.ctx_at_forward;
      }
      sl.active = more ? 1 : 0;
    }
  }
  __threadfence_block();
  workers_sync();
  if (w.wtid() == 0) {
    const int iter_done = first ? 0 : ld_weak(&hdr->iter) + 1;
    int T = 0, G = 0, nc = 0, units = 0, slots = 0, err = 0;
    const int CH = c.attn_chunk;
    int base_units = 0, max_len = 1;
    for (int s = 0; s < c.n_seqs; ++s) {
      const SeqState& st = c.seq_state[s];
      if (!c.seq_slot[s].active || iter_done >= max_forwards) continue;
      base_units += st.B_out * c.md.n_kv * ((st.ctx + st.L_out + CH - 1) / CH) * ((st.L_out + kQTokPerUnit - 1) / kQTokPerUnit);
      max_len = max(max_len, st.L_out);
    }
    const int hsplit = pick_attn_hsplit(base_units, (int)gridDim.x, c.md.n_rep, max_len, c.attn_hsplit);
    for (int s = 0; s < c.n_seqs; ++s) {
      SeqState& st = c.seq_state[s];
      SeqSlot& sl = c.seq_slot[s];
      for (int j = 0; j < sl.commit_n && nc < c.T_cap; ++j) {
        c.commits[nc].src_row = sl.commit_src_row + j;
        c.commits[nc].seq = s;
        c.commits[nc].pos = sl.commit_pos + j;
        ++nc;
      }
      if (!sl.active) continue;
      if (iter_done >= max_forwards) {
        st.finished = 1; st.stop_reason = STOP_MAX_ITERS; sl.active = 0;
        continue;
      }
      const int B = st.B_out, L = st.L_out;
      const int nchunks = (st.ctx + L + CH - 1) / CH;
      const int qsplit = (L + kQTokPerUnit - 1) / kQTokPerUnit;
      if (T + B * L > c.T_cap || G + B > c.G_cap || slots + B * L * nchunks > c.attn_slots_cap ||
          st.ctx + L > c.max_pages_per_seq * kPage || st.ctx + L > c.md.max_pos) {
        err = 1; st.finished = 1; st.stop_reason = STOP_ERROR; st.error = PERR_CAPACITY; sl.active = 0;
        continue;
      }
      sl.row0 = T; sl.B = B; sl.L = L; sl.g0 = G; sl.ctx_at_forward = st.ctx;
      for (int i = 0; i < B; ++i) {
        Group& g = c.groups[G];
        g.row0 = T + i * L; g.len = L; g.seq = s; g.ctx = st.ctx;
        g.nchunks = nchunks; g.qsplit = qsplit; g.unit0 = units; g.slot0 = slots;
        units += c.md.n_kv * nchunks * qsplit * hsplit;
        slots += L * nchunks;
        ++G;
      }
      if (c.trace != nullptr) {
        int tl = *c.trace_len;
        if (tl + 5 + B * L <= c.trace_cap) {
          c.trace[tl] = iter_done; c.trace[tl + 1] = s; c.trace[tl + 2] = st.ctx; c.trace[tl + 3] = B; c.trace[tl + 4] = L;
          for (int i = 0; i < B; ++i)
            for (int j = 0; j < L; ++j) c.trace[tl + 5 + i * L + j] = st.out_rows[i][j];
          *c.trace_len = tl + 5 + B * L;
        }
      }
      T += B * L;
    }
    hdr->rows_total = (first ? 0 : ld_weak(&hdr->rows_total)) + T;
    hdr->T = T; hdr->G = G; hdr->n_commit = nc; hdr->attn_units = units; hdr->attn_slots = slots; hdr->attn_hsplit = hsplit;
    hdr->iter = iter_done;
    hdr->want_logits = 0;
    if (err) hdr->error = err;
    int nl = 0;
    const int lm_step = c.ng_mode ? kNgRows : kArgmaxRows;
    for (int r = 0; r < T && nl < kMaxLmTasks; r += lm_step) {
      hdr->lm_row0[nl] = r;
      hdr->lm_rows[nl] = min(lm_step, T - r);
      ++nl;
    }
    hdr->n_lm_tasks = nl;
    hdr->done = (T == 0) ? 1 : 0;
  }
  __threadfence_block();
  workers_sync();
  for (int s = 0; s < c.n_seqs; ++s) {
    const SeqSlot& sl = c.seq_slot[s];
    if (!sl.active) continue;
    const SeqState& st = c.seq_state[s];
    const int n = sl.B * sl.L;
    for (int idx = w.wtid(); idx < n; idx += kWorkerThreads) {
      const int i = idx / sl.L, j = idx - i * sl.L;
      c.tok[sl.row0 + idx] = st.out_rows[i][j];
      c.row_group[sl.row0 + idx] = sl.g0 + i;
      c.row_pos[sl.row0 + idx] = sl.ctx_at_forward + j;
      const Group& g = c.groups[sl.g0 + i];
      c.row_attn[sl.row0 + idx] = make_int2(g.slot0 + j * g.nchunks, (sl.ctx_at_forward + j) / c.attn_chunk);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
enum Role : int { ROLE_W = 0, ROLE_MMA = 1, ROLE_WORKER = 2, ROLE_X = 3 };
// stage ids recorded with every profiled grid barrier (jacobiforcing_b200/runtime.py maps them to names)
enum StageId : int { SID_NONE = 0, SID_EMBED = 1, SID_QKV = 2, SID_ROPE = 3, SID_ATTN = 4, SID_COMBINE = 5, SID_O = 6,
                     SID_O_NORM = 7, SID_GU = 8, SID_DOWN = 9, SID_DOWN_NORM = 10, SID_LM = 11, SID_POLICY = 12,
                     SID_O_FUSED = 13, SID_DOWN_FUSED = 14, SID_QKV_FUSED = 15, SID_GU_DOWN_FUSED = 16 };

struct IterPlan {
  int T, G, done, n_commit, attn_units, n_lm, want_logits, hsplit;
};

template <int ROLE>
__device__ void run_role(uint8_t* tiles, int ring_bytes, int run_mode, int max_forwards) {
  // the fixed shared-memory blocks are addressed by constant offsets: no pointer register (or stack slot) is spent on them
  const DevCtx& c = *reinterpret_cast<const DevCtx*>(jf_smem_base + kSmemCtxOff);
  GemmSmem* const gs = reinterpret_cast<GemmSmem*>(jf_smem_base);
  WorkerShared* const ws = reinterpret_cast<WorkerShared*>(jf_smem_base + kSmemWsOff);
  const WorkerView w{};
  const int cta = blockIdx.x, nctas = gridDim.x;
  BarState bs;
  PipeState pipe;
  AccState acc;
  const ModelDims& md = c.md;
  const int kc_h = md.h >> 6, kc_q = md.q_dim >> 6, kc_i = md.inter >> 6;
  const int nt_qkv = (md.qkv_dim + 127) >> 7, nt_h = (md.h + 127) >> 7, nt_i = (md.inter + 127) >> 7;
  const int nt_v = (md.vocab + 127) >> 7;
  const bool tp = c.tp_size > 1;
  int nstages = 1, slot_bytes = 0;
  const uint64_t w_policy = (ROLE == ROLE_W) ? l2_policy_evict_first() : 0ull;
  unsigned int df_uses = 0u;                   // gate/up -> down dataflow hand-overs so far in this launch

  auto skip_barrier = [&](bool xgpu) { bs.epoch += 1; if (xgpu) bs.xcount += 1; };
  // fuse_kind: 0 = plain GEMM stage; 3 = QKV projection with bias + rotary + scatter as its tail; 1 / 2 = O / down
  // projection whose split-K reduce + residual + RMSNorm runs as the tail of the epilogue.  The CTAs that hold an
  // n-tile's k-splits share the tile's rows and poll the tagged partial sums (EPI_PARTIAL_TAGGED) of the other splits
  // (needs one unit per CTA at most: no CTA may owe a second unit while it polls)
  // df_wait >= 0: instead of a grid barrier, this stage's X producer waits until the df_wait producer tiles of the
  // previous stage that its k-range reads have arrived on the range counter of its k-split (see fuse_kind 5)
  auto gemm_stage = [&](const GemmTask& t, bool barrier_after, int stage_id, bool xgpu = false, int fuse_kind = 0,
                        const __nv_bfloat16* normw = nullptr, int layer = 0, bool df_wait = false) {
    if (ROLE == ROLE_W) {
      if (c.debug_skip & 16) wait_epoch(c.bar_counter, bs.target(nctas));   // experiment: no run-ahead
      gemm_produce_w(t, gs, tiles, nstages, slot_bytes, cta, nctas, pipe, c.w_evict_first != 0, w_policy);
    } else if (ROLE == ROLE_X) {
      // the X producer reads its operand only with bulk copies: no acquire fence, no proxy fence on this side (the writers
      // fenced generic -> async before they released the counter); JF_X_FENCE=1 restores both
      const bool xf = c.x_fence != 0;
      if (!df_wait) {
        wait_epoch(c.bar_counter, bs.target(nctas), false, xf);
      } else if (cta < t.n_tiles * t.splits) {
        // dataflow: the gate/up tiles whose features are this unit's k-range (tile j holds k-chunks 2j, 2j+1)
        const UnitRange r = unit_range(t, cta);
        const unsigned int need = (unsigned int)df_tiles_needed(r.k0, r.k1);
        wait_epoch(c.tile_ctr + r.split, df_uses * need, false, xf);
      }
      if (xf) fence_proxy_async_all();
      gemm_produce_x(t, gs, tiles, nstages, slot_bytes, cta, nctas, pipe);
    } else if (ROLE == ROLE_MMA) {
      gemm_issue_mma(t, gs, tiles, nstages, slot_bytes, cta, nctas, pipe, acc);
    } else {
      ws->tag = c.tag_base + bs.epoch + 1u;
      prof_mark(c, w, 600);
      gemm_epilogue(t, gs, cta, nctas, acc);
      if (w.prof_on() && w.wtid() == 0 && c.prof_marks != nullptr) {
        int* cnt = reinterpret_cast<int*>(w.red()) + 4;
        const int n = *cnt;
        // developer aid: 600 = epilogue entered, 601 = first accumulator handed over, 602 = epilogue done
        if (n + 4 <= 8192) { c.prof_marks[n] = 601; c.prof_marks[n + 1] = gs->t_acc_full; c.prof_marks[n + 2] = 602; c.prof_marks[n + 3] = clock64(); *cnt = n + 4; }
      }
      if (fuse_kind == 5) {
        // gate/up -> down without a grid barrier: publish this tile to the k-splits of the down projection that read it
        const bool bulk = swiglu_bulk_ok(t, gs->stage_rows);
        if (!bulk) {
          fence_proxy_async_all();                      // the tile is read by bulk copies (async proxy) of other CTAs
          workers_sync();
        }
        // bulk: the tile left through the async proxy and thread 0 has waited for the copy to complete (gemm_epilogue)
        if (w.wtid() == 0 && cta < t.n_tiles) {
          const unsigned int df_mask = ws->df_mask;
          if (bulk) {
            prof_mark(c, w, 605);
            fence_acq_rel_gpu();
            prof_mark(c, w, 607);
            for (int s2 = 0; s2 < c.splits_down; ++s2)
              if ((df_mask >> s2) & 1u) red_relaxed_gpu_add_u32(c.tile_ctr + s2, 1u);
          } else {
            for (int s2 = 0; s2 < c.splits_down; ++s2)
              if ((df_mask >> s2) & 1u) red_release_gpu_add_u32(c.tile_ctr + s2, 1u);
          }
          prof_mark(c, w, 606);
        }
      } else if (fuse_kind != 0 && cta < t.n_tiles * t.splits) {
        // no flag, fence or barrier: the tails poll the tagged partial sums themselves
        const int n_tile = cta / t.splits, split = cta - n_tile * t.splits;
        if (fuse_kind == 3) qkv_rope_tail(c, ws, t.M, layer, n_tile, split, t.splits);
        else if (fuse_kind == 4) gu_swiglu_tail(c, ws, t.M, n_tile, split, t.splits);
        else gemm_norm_tail(c, ws, t.M, n_tile, split, t.splits, normw, fuse_kind - 1,
                            fuse_kind == 1 ? c.tp_attn > 1 : c.tp_size > 1);
      }
    }
    if (ROLE == ROLE_WORKER && t.bulk_rows) bulk_wait_read_all();   // the staging tile (attention buffer) may be rewritten
    if (barrier_after) {
      if (ROLE == ROLE_WORKER)
        grid_barrier(c, w, bs, true, t.epilogue == EPI_SWIGLU_BF16 || fuse_kind == 1 || fuse_kind == 2 || fuse_kind == 4, xgpu, stage_id);
      else skip_barrier(xgpu);
    }
  };
#define JF_SIMT_STAGE_X(tma_out, sid, body)                    \
  do {                                                        \
    if (ROLE == ROLE_WORKER) { ws->tag = c.tag_base + bs.epoch + 1u; body; grid_barrier(c, w, bs, true, tma_out, false, sid); } else { skip_barrier(false); } \
  } while (0)

#define JF_SIMT_STAGE(sid, body) JF_SIMT_STAGE_X(true, sid, body)
  auto launch_mark = [&](int id) {     // developer aid (JF_DEBUG_SKIP & 2048): coarse marks of the launch itself
    if (ROLE == ROLE_WORKER && (c.debug_skip & 2048) && cta == 0 && w.wtid() == 0 && c.prof_marks != nullptr) {
      int* cnt = reinterpret_cast<int*>(w.red()) + 5;
      if (id == 990) *cnt = 0;
      const int n = *cnt;
      if (n + 2 <= 8192) { c.prof_marks[n] = id; c.prof_marks[n + 1] = clock64(); *cnt = n + 2; *c.prof_marks_n = n + 2; }
    }
  };
  if (ROLE == ROLE_WORKER && w.wtid() == 0) {
    unsigned int m = 0u;
    for (int s2 = 0; s2 < c.splits_down && s2 < 32; ++s2)
      if (df_tile_feeds_split(cta, s2, c.splits_down, kc_i)) m |= 1u << s2;
    ws->df_mask = m;                     // read by this thread only (gate/up -> down hand-over)
  }
  launch_mark(990);
  if (run_mode == RUN_JACOBI) {
    JF_SIMT_STAGE(SID_POLICY, if (cta == 0) policy_stage_jacobi(c, ws, true, max_forwards));
  }
  launch_mark(991);

  for (;;) {
    if (ROLE != ROLE_WORKER) wait_epoch(c.bar_counter, bs.target(nctas));
    const PlanHdr* hdr = c.hdr;
    IterPlan ip;
    ip.T = ld_weak(&hdr->T); ip.G = ld_weak(&hdr->G); ip.done = ld_weak(&hdr->done);
    ip.n_commit = ld_weak(&hdr->n_commit); ip.attn_units = ld_weak(&hdr->attn_units);
    ip.n_lm = ld_weak(&hdr->n_lm_tasks); ip.want_logits = ld_weak(&hdr->want_logits);
    ip.hsplit = max(1, ld_weak(&hdr->attn_hsplit));
    if (ip.done) {
      if (ROLE == ROLE_WORKER) {
        if (w.prof_on() && w.wtid() == 0) {
          int prof_n = ws->prof_n;
          if (prof_n + 3 <= c.prof_cap) {
            c.prof[prof_n] = -(long long)globaltimer_ns(); c.prof[prof_n + 1] = clock64(); c.prof[prof_n + 2] = c.prof[prof_n + 1];
            prof_n += 3;
          }
          *c.prof_count = prof_n;
          if (c.prof_marks_n) *c.prof_marks_n = reinterpret_cast<int*>(w.red())[4];
        }
        if (tp) grid_barrier(c, w, bs, false, false, true);      // every rank has read every rank's keys
        if (cta == 0)
          for (int t = w.wtid(); t < 2 * c.T_cap; t += kWorkerThreads) c.argmax_keys[t] = 0ull;   // clean for the next launch
        stage_commit(c, ws, ip.n_commit);
      }
      launch_mark(992);
      break;
    }
    const int T = ip.T;
    if (ROLE == ROLE_WORKER && (c.debug_skip & 2048) && cta == 0 && w.wtid() == 0 && c.prof_marks != nullptr) {
      // developer aid: one (1000 + rows, clock) mark per forward of the launch (per-forward durations)
      int* cnt = reinterpret_cast<int*>(w.red()) + 5;
      const int n = *cnt;
      if (n + 2 <= 8192) { c.prof_marks[n] = 1000 + T; c.prof_marks[n + 1] = clock64(); *cnt = n + 2; *c.prof_marks_n = n + 2; }
    }
    if (ROLE == ROLE_WORKER) {
      if (w.prof_on() && w.wtid() == 0) {                                    // forward profiled: publish and stop
        int prof_n = ws->prof_n;
        if (prof_n + 3 <= c.prof_cap) {        // closing record: (-wall clock ns, SM clock) calibrates cycles -> time
          c.prof[prof_n] = -(long long)globaltimer_ns(); c.prof[prof_n + 1] = clock64(); c.prof[prof_n + 2] = c.prof[prof_n + 1];
          prof_n += 3;
        }
        *c.prof_count = prof_n;
        if (c.prof_marks_n) *c.prof_marks_n = reinterpret_cast<int*>(w.red())[4];
      }
      workers_sync();                          // thread 0 has read the flag the others are about to rewrite
      ws->prof_on = 0;
      if (cta == 0 && c.prof != nullptr && ld_weak(&hdr->iter) == c.prof_forward) {
        ws->prof_on = 1;
        if (w.wtid() == 0) {
          reinterpret_cast<int*>(w.red())[4] = 0;
          ws->prof_n = 0;
          if (c.prof_cap >= 3) { c.prof[0] = -(long long)globaltimer_ns(); c.prof[1] = clock64(); c.prof[2] = c.prof[1]; ws->prof_n = 3; }
        }
      }
    }
    if (ROLE == ROLE_WORKER) {
      // plan arrays of this forward -> shared memory (readers are at least one grid barrier away)
      const int ok = (T <= kPlanRows && ip.G <= kPlanGroups) ? 1 : 0;
      if (ok) {
        PlanCache* pc = ws->plan;
        for (int t = w.wtid(); t < T; t += kWorkerThreads) { pc->row_pos[t] = ld_weak(&c.row_pos[t]); pc->row_attn[t] = ld_weak(&c.row_attn[t]); }
        for (int g = w.wtid(); g < ip.G; g += kWorkerThreads) pc->groups[g] = load_group(c, g);
      }
      if (w.wtid() == 0) {
        ws->plan_ok = ok;
        PlanCache* pc = ws->plan;
        pc->first_unit_ok = 0;
        if (ok && cta < ip.attn_units) {       // decoded from the global plan: the shared copy is still being filled
          WorkerCtx wc = load_wctx(ws);
          wc.plan_ = nullptr;
          pc->first_unit = decode_attn_unit(c, wc, cta, ip.G, ip.hsplit);
          pc->first_unit_ok = 1;
        }
      }
    }
    slot_bytes = gemm_slot_bytes(min(kMaxRowsMMA, round_up16(T)));
    nstages = min(kMaxStages, ring_bytes / slot_bytes);
    pipe.stage = 0;

    JF_SIMT_STAGE(SID_EMBED, stage_commit(c, ws, ip.n_commit); stage_rows_norm(c, ws, T, true, 0, c.layer[0].ln1, 0, 1));
    // the fused tails apply to Jacobi-sized forwards on one GPU (see gemm_stage); prefill-sized ones keep the stages
    const bool fuse_ok = c.fuse_norm != 0 && T <= c.fuse_rows;
    const bool fuse_o = fuse_ok && nt_h * c.splits_o <= nctas;
    const bool fuse_d = fuse_ok && nt_h * c.splits_down <= nctas;
    // split-K gate/up is a property of the engine (set when it is created), never of the forward: every forward must
    // add the same partial sums in the same order (Jacobi == AR == chunked prefill, bit for bit)
    const bool fuse_g = c.splits_gu > 1;
    const bool dflow = c.dataflow_gd != 0 && fuse_d && !fuse_g && !tp && 2 * nt_i >= kc_i && c.splits_down <= 32;
    const bool fuse_q = c.fuse_rope != 0 && T <= kFuseRopeRows && nt_qkv * c.splits_qkv <= nctas && nt_qkv <= kMaxTilesPerKind;

    auto base_task = [&]() {
      GemmTask t;
      t.tag = 0u; t.bulk_rows = 0;
      t.M = T; t.x_row0 = 0; t.xrows_cap = c.xrows_cap; t.out_rows_cap = c.xrows_cap; t.index_offset = 0;
      t.out_row_stride = c.part_splits * c.part_n; t.out_split_stride = c.part_n;
      return t;
    };
    auto task_qkv = [&](int l, bool fused = false) {      // QKV projection (split-K partial sums)
      GemmTask t = base_task();
      t.W = c.layer[l].wqkv; t.X = c.xn; t.out = c.part; t.n_tiles = nt_qkv; t.k_chunks = kc_h; t.splits = c.splits_qkv;
      t.N = md.qkv_dim; t.epilogue = EPI_PARTIAL_F32;
      if (fused) { t.out = c.tpart; t.epilogue = EPI_PARTIAL_TAGGED; t.tag = c.tag_base + bs.epoch + 1u; }
      return t;
    };
    // stand-alone (prefill-sized) row-split stages: which fp32 partial buffer a projection uses.  Peers read the buffers of
    // a sharded projection, so consecutive sharded projections alternate between two of them (O = 0, down = 1); with
    // MLP-only tensor parallelism the down projections of consecutive layers alternate and the local O uses a third.
    const bool mlp_only = tp && c.tp_attn == 1;
    auto which_o = [&]() { return mlp_only ? 2 : 0; };
    auto which_d = [&](int l) { return mlp_only ? (l & 1) : 1; };
    auto task_o = [&](int l, bool fused = false) {
      GemmTask t = base_task();
      t.W = c.layer[l].wo; t.X = c.xa; t.out = c.part_rs[which_o()]; t.n_tiles = nt_h; t.k_chunks = kc_q; t.splits = c.splits_o;
      t.N = md.h; t.epilogue = EPI_PARTIAL_F32; t.out_row_stride = c.part_splits * md.h; t.out_split_stride = md.h;
      if (fused) { t.out = c.tpart_rs[0]; t.epilogue = EPI_PARTIAL_TAGGED; t.tag = c.tag_base + bs.epoch + 1u; t.bulk_rows = c.epi_bulk; }
      return t;
    };
    auto task_gu = [&](int l, bool fused = false) {       // gate/up projection with the SwiGLU epilogue
      GemmTask t = base_task();
      t.W = c.layer[l].wgu; t.X = c.xn; t.out = c.xact; t.n_tiles = nt_i; t.k_chunks = kc_h; t.splits = 1;
      t.N = md.inter; t.epilogue = EPI_SWIGLU_BF16;
      if (fused) {                    // split-K: tagged gate / up partial sums, SwiGLU in the tail
        t.out = c.tgu; t.splits = c.splits_gu; t.epilogue = EPI_SWIGLU_TAGGED; t.tag = c.tag_base + bs.epoch + 1u;
        t.out_row_stride = c.splits_gu * 2 * md.inter; t.out_split_stride = 2 * md.inter;
      }
      return t;
    };
    auto task_down = [&](int l, bool fused = false) {
      GemmTask t = base_task();
      t.W = c.layer[l].wd; t.X = c.xact; t.out = c.part_rs[which_d(l)]; t.n_tiles = nt_h; t.k_chunks = kc_i; t.splits = c.splits_down;
      t.N = md.h; t.epilogue = EPI_PARTIAL_F32; t.out_row_stride = c.part_splits * md.h; t.out_split_stride = md.h;
      if (fused) { t.out = c.tpart_rs[1]; t.epilogue = EPI_PARTIAL_TAGGED; t.tag = c.tag_base + bs.epoch + 1u; t.bulk_rows = c.epi_bulk; }
      return t;
    };
    auto task_lm = [&](int i) {
      GemmTask t = base_task();
      t.W = c.lm_head; t.X = c.xn; t.out = c.argmax_keys + (size_t)(ld_weak(&hdr->iter) & 1) * c.T_cap;
      t.n_tiles = nt_v; t.k_chunks = kc_h; t.splits = 1; t.index_offset = md.vocab_offset;
      t.M = ld_weak(&hdr->lm_rows[i]); t.x_row0 = ld_weak(&hdr->lm_row0[i]);
      t.N = md.vocab; t.epilogue = EPI_ARGMAX;
      if (c.ng_mode && run_mode == RUN_JACOBI) {
        t.epilogue = EPI_SAMPLE;
        t.sample_ctx = &c;
        t.fwd = ld_weak(&hdr->iter) + c.ng_rng_offset;
        t.fwd_parity = ld_weak(&hdr->iter) & 1;
        t.rows_total = T;
      }
      return t;
    };

    for (int l = 0; l < md.layers; ++l) {
      const LayerW& lw = c.layer[l];
      if (ROLE == ROLE_WORKER) attention_prefetch(c, ws, l, ip.G, ip.attn_units, ip.hsplit);
      if (fuse_q) {
        gemm_stage(task_qkv(l, true), true, SID_QKV_FUSED, false, 3, nullptr, l);
      } else {
        gemm_stage(task_qkv(l), true, SID_QKV);
        JF_SIMT_STAGE_X(false, SID_ROPE, if (!(c.debug_skip & 1)) stage_qkv_rope(c, ws, T, l));
      }
      JF_SIMT_STAGE_X(false, SID_ATTN, if (!(c.debug_skip & 2)) stage_attention(c, ws, l, ip.G, ip.attn_units, ip.hsplit));
      JF_SIMT_STAGE(SID_COMBINE, if (!(c.debug_skip & 4)) stage_attn_combine(c, ws, T));
      if (fuse_o) {
        gemm_stage(task_o(l, true), true, SID_O_FUSED, false, 1, lw.ln2);
      } else {
        gemm_stage(task_o(l), true, SID_O, c.tp_attn > 1);
        JF_SIMT_STAGE(SID_O_NORM, if (!(c.debug_skip & 8)) stage_rows_norm(c, ws, T, false, c.splits_o, lw.ln2, which_o(), c.tp_attn));
      }
      const __nv_bfloat16* next_norm = (l + 1 < md.layers) ? c.layer[l + 1].ln1 : c.final_norm;
      if (dflow) {
        // no grid barrier between gate/up and down: the down projection's X producers wait for exactly the gate/up
        // tiles they read (every down tile's tail still needs all of them, so nothing of gate/up's input is overwritten early)
        df_uses += 1u;
        gemm_stage(task_gu(l), false, SID_GU, false, 5);
        gemm_stage(task_down(l, true), true, SID_GU_DOWN_FUSED, false, 2, next_norm, 0, true);
        continue;
      }
      if (fuse_g) gemm_stage(task_gu(l, true), true, SID_GU, false, 4);
      else gemm_stage(task_gu(l), true, SID_GU);
      if (fuse_d) {
        gemm_stage(task_down(l, true), true, SID_DOWN_FUSED, false, 2, next_norm);
      } else {
        gemm_stage(task_down(l), true, SID_DOWN, tp);
        JF_SIMT_STAGE(SID_DOWN_NORM, stage_rows_norm(c, ws, T, false, c.splits_down, next_norm, which_d(l), c.tp_size));
      }
    }

    // LM head: fused bf16-rounded argmax (no logits in HBM); optional fp32 logits for the tolerance checks
    for (int i = 0; i < ip.n_lm; ++i) {
      GemmTask t = task_lm(i);
      gemm_stage(t, false, SID_LM);
      if (i == 0 && ip.want_logits) {
        t.out = c.logits - (size_t)t.x_row0 * md.vocab;
        t.epilogue = EPI_PARTIAL_F32; t.out_row_stride = md.vocab; t.out_split_stride = 0;
        gemm_stage(t, false, SID_LM);
      }
    }
    if (ROLE == ROLE_WORKER) grid_barrier(c, w, bs, true, true, tp, SID_LM); else skip_barrier(tp);

    JF_SIMT_STAGE(SID_POLICY, if (cta == 0) {
      if (run_mode == RUN_JACOBI) policy_stage_jacobi(c, ws, false, max_forwards); else policy_stage_plan(c, ws);
    });
  }
#undef JF_SIMT_STAGE
#undef JF_SIMT_STAGE_X
}

__global__ void __launch_bounds__(kNumThreads, 1)
jf_persistent_kernel(const DevCtx* __restrict__ ctxp, int smem_total, int run_mode, int max_forwards) {
  extern __shared__ __align__(1024) uint8_t smem[];
  GemmSmem* gs = reinterpret_cast<GemmSmem*>(smem);
  float* red = reinterpret_cast<float*>(smem + ((sizeof(GemmSmem) + 15) & ~(size_t)15));
  // The context (pointers, dims, per-layer weight table) is copied into shared memory once: every stage reads a dozen
  // of its fields, and from global memory each of them costs an L2 round trip after every grid barrier (the barrier's
  // acquire invalidates L1).
  DevCtx* sctx = reinterpret_cast<DevCtx*>(red + 32);        // red[0..7] scratch words, red[8..31] WorkerShared
  {
    const uint4* src = reinterpret_cast<const uint4*>(ctxp);
    uint4* dst = reinterpret_cast<uint4*>(sctx);
    for (int i = threadIdx.x; i < (int)(sizeof(DevCtx) / 16); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const DevCtx& c = *sctx;
  PlanCache* plan = reinterpret_cast<PlanCache*>(reinterpret_cast<uint8_t*>(sctx) + ((sizeof(DevCtx) + 15) & ~(size_t)15));
  uint8_t* attn = reinterpret_cast<uint8_t*>(plan) + ((sizeof(PlanCache) + 15) & ~(size_t)15);
  const int attn_need = 2 * kAttnSub * (c.md.d + 8) * 2;
  const int best_need = 4 * kArgmaxRows * 8;
  const int attn_bytes = attn_need > best_need ? attn_need : best_need;
  if (threadIdx.x == 0) {
    gs->best = reinterpret_cast<unsigned long long*>(attn);
    gs->stage_rows = c.swiglu_bulk ? ((attn_bytes / 256) & ~15) : 0;
    gs->tag_stage_rows = c.epi_bulk ? ((attn_bytes / (kTileN * 8)) & ~15) : 0;
  }
  const uint32_t base_addr = smem_u32(smem);     // the ring must be 1024-byte aligned in the shared window (sw128)
  const size_t tiles_off = (((size_t)base_addr + (size_t)(attn - smem) + attn_bytes + 1023) & ~(size_t)1023) - base_addr;
  uint8_t* tiles = smem + tiles_off;
  const int ring_bytes = smem_total - (int)tiles_off;

  gemm_setup(gs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(sizeof(WorkerShared) <= 24 * sizeof(float), "WorkerShared must fit red[8..31]");
  WorkerShared* ws = reinterpret_cast<WorkerShared*>(red + 8);      // red[8..31]
  if (threadIdx.x == 0) {
    ws->red = red; ws->attn_smem = attn; ws->prof_on = 0; ws->tag = 0u; ws->plan = plan; ws->plan_ok = 0; ws->prof_n = 0;
    ws->tmem_base = gs->tmem_base; ws->attn_phase = 0u; ws->attn_bar = &gs->attn_bar;
  }
  __syncthreads();

  if (warp == kWarpW) {
    if (lane == 0) run_role<ROLE_W>(tiles, ring_bytes, run_mode, max_forwards);
  } else if (warp == kWarpMMA) {
    if (lane == 0) run_role<ROLE_MMA>(tiles, ring_bytes, run_mode, max_forwards);
  } else if (warp == kWarpX) {
    if (lane == 0) run_role<ROLE_X>(tiles, ring_bytes, run_mode, max_forwards);
  } else {
    run_role<ROLE_WORKER>(tiles, ring_bytes, run_mode, max_forwards);
  }
  __syncwarp();
  gemm_teardown(gs);
}

}  // namespace jf
