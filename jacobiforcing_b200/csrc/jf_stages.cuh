// jf_stages.cuh — the SIMT / mma.sync stages of one Qwen2 forward over the in-flight draft rows, executed by the four
// worker warps (128 threads) of every CTA of the persistent kernel, between the tcgen05 weight GEMMs.
//
// What each stage re-expresses (reference call sites; the arithmetic itself is transformers' Qwen2):
//   stage_rows_norm   embed_tokens / residual add + Qwen2RMSNorm            unified.py:429, :452-462; layers/layernorm.py:16-40
//   stage_qkv_rope    q/k/v bias add, apply_rotary_pos_emb, KV append       layers/rotary_embedding.py:6-48, attention.py:10-40
//   stage_attention   causal GQA attention over [committed pages | in-flight rows]   layers/attention.py:133-136
//   stage_attn_combine  split-KV merge, output in the O-projection's operand layout
//   stage_commit      accepted in-flight K/V -> paged cache (replaces DynamicCache trim / trim_kv_only_fast)
// Rounding points follow the HF bf16 modules: linear outputs, RMSNorm (fp32 statistics, bf16
// normalised value times bf16 weight), rotary products and sums, residual adds are each rounded to bf16.
#pragma once
#include <type_traits>
#include "jf_types.h"
#include "jf_gemm.cuh"

namespace jf {

struct PlanCache;
// Worker context of the non-inlined stage functions: a register copy of the shared-memory block below, taken once per call.
// What derives from special registers (CTA, thread, lane) is never stored - it is re-read where it is used.
struct WorkerCtx {
  float* red_;      // [16] scratch words
  uint8_t* attn_smem_;
  int prof_on_;     // CTA 0, forward being profiled
  unsigned int tag_; // unique id of the stage instance being executed (launch id, barrier count)
  const PlanCache* plan_;   // non-null: plan arrays of the forward in flight, in shared memory
  __device__ __forceinline__ int cta() const { return (int)blockIdx.x; }
  __device__ __forceinline__ int nctas() const { return (int)gridDim.x; }
  __device__ __forceinline__ int wtid() const { return (int)threadIdx.x - 2 * 32; }       // 0..127 among the worker warps
  __device__ __forceinline__ int wwarp() const { return (int)(threadIdx.x >> 5) - 2; }    // 0..3 (kWarpWorker0 = 2)
  __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
  __device__ __forceinline__ float* red() const { return red_; }
  __device__ __forceinline__ uint8_t* attn_smem() const { return attn_smem_; }
  __device__ __forceinline__ int prof_on() const { return prof_on_; }
  __device__ __forceinline__ unsigned int tag() const { return tag_; }
  __device__ __forceinline__ const PlanCache* plan() const { return plan_; }
};
// The part of the worker context that is not derivable from special registers lives in shared memory.  Non-inlined
// stage functions receive a pointer to it and rebuild their WorkerCtx in registers: a struct passed by reference to a
// non-inlined function lives on the caller's stack, and every field access becomes a local-memory load that misses L1
// after each grid barrier (ncu: long-scoreboard stalls on w.lane() / grp.slot0 inside the attention stage).
// Per-forward copy of the small plan arrays every layer's stages index (row positions, attention slots, groups): read
// from global memory they cost an L2 round trip per stage and layer; filled once per forward by the worker warps.
constexpr int kPlanRows = 64, kPlanGroups = 8;
struct AttnUnitId { Group grp; int kvh, hs, chunk, qs; };
struct PlanCache {
  int row_pos[kPlanRows];
  int2 row_attn[kPlanRows];
  Group groups[kPlanGroups];
  AttnUnitId first_unit;     // this CTA's first attention unit, decoded once per forward (every layer runs the same units)
  int first_unit_ok;
};
struct WorkerShared {
  float* red;
  uint8_t* attn_smem;
  int prof_on;
  unsigned int tag;
  PlanCache* plan;     // shared-memory plan cache
  int plan_ok;         // 1: the forward in flight fits the cache and it is filled
  uint32_t tmem_base;  // tensor memory of the CTA (free between the GEMM stages: the attention stage borrows it)
  uint32_t attn_phase; // parity of the attention MMA barrier
  uint64_t* attn_bar;
  int prof_n;          // stage-profile words written so far (CTA 0, worker thread 0)
  unsigned int df_mask; // bit s: this CTA's gate/up tile feeds k-split s of the down projection (fixed per launch)
};
// Shared-memory map of the kernel's fixed blocks (jf_kernel.cuh lays them out): GemmSmem, 8 scratch words, WorkerShared, the
// kernel context.  Fixed offsets from the dynamic shared-memory base, so code that needs them holds no pointer register.
extern __shared__ __align__(1024) uint8_t jf_smem_base[];
constexpr size_t kSmemRedOff = (sizeof(GemmSmem) + 15) & ~(size_t)15;
constexpr size_t kSmemWsOff = kSmemRedOff + 8 * sizeof(float);
constexpr size_t kSmemCtxOff = kSmemRedOff + 32 * sizeof(float);
__device__ __forceinline__ WorkerCtx load_wctx(const WorkerShared* ws) {
  WorkerCtx w;
  w.red_ = ws->red; w.attn_smem_ = ws->attn_smem; w.prof_on_ = ws->prof_on; w.tag_ = ws->tag;
  w.plan_ = ws->plan_ok ? ws->plan : nullptr;
  return w;
}
// The persistent kernel's own loop keeps NO copy: every field is read from the shared-memory block where it is used (a
// register copy held across the whole forward ends up in local memory - 255 registers are not enough for that loop - and
// local memory misses L1 after every grid barrier's acquire fence; shared memory does not).
struct WorkerView {
  __device__ __forceinline__ WorkerShared* shared() const { return reinterpret_cast<WorkerShared*>(jf_smem_base + kSmemWsOff); }
  __device__ __forceinline__ int cta() const { return (int)blockIdx.x; }
  __device__ __forceinline__ int nctas() const { return (int)gridDim.x; }
  __device__ __forceinline__ int wtid() const { return (int)threadIdx.x - 2 * 32; }
  __device__ __forceinline__ int wwarp() const { return (int)(threadIdx.x >> 5) - 2; }
  __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
  __device__ __forceinline__ float* red() const { return reinterpret_cast<float*>(jf_smem_base + kSmemRedOff); }
  __device__ __forceinline__ uint8_t* attn_smem() const { return shared()->attn_smem; }
  __device__ __forceinline__ int prof_on() const { return shared()->prof_on; }
  __device__ __forceinline__ unsigned int tag() const { return shared()->tag; }
};

template <class W>
__device__ __forceinline__ void prof_mark(const DevCtx& c, const W& w, int id) {
  if (w.prof_on() && w.wtid() == 0 && c.prof_marks != nullptr) {
    // the running count lives in shared memory (a global counter costs an L2 round trip per mark: ~0.45 us measured);
    // it is published with the stage stamps when the profiled forward ends
    int* cnt = reinterpret_cast<int*>(w.red()) + 4;
    const int n = *cnt;
    if (n + 2 <= 8192) { c.prof_marks[n] = id; c.prof_marks[n + 1] = clock64(); *cnt = n + 2; }
  }
}

constexpr uint32_t kBarWorkers = 1;   // named barrier ids (0 is __syncthreads)
constexpr uint32_t kBarReduce  = 2;

__device__ __forceinline__ void workers_sync() { named_bar_sync(kBarWorkers, kWorkerThreads); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ---------------------------------------------------------------------------------------------------------------
// rows: (embedding | split-K reduce + residual add) -> hidden, then RMSNorm -> packed X operand of the next GEMM.
// Stand-alone stage: a row is cut into slices of 1024 columns (one CTA each, one 8-column group per thread, every
// load in flight at once); every slice publishes the sums of squares of its 128-column tiles, polls the row's other
// tiles and normalises its own columns.
// ---------------------------------------------------------------------------------------------------------------
// Tagged partial sums (EPI_PARTIAL_TAGGED): a 64-bit word carries {stage tag, fp32 bits}, so the consumer polls the data
// itself with L2 (.cg) loads until every word it needs has the tag of the stage in flight - one L2 round trip between the
// producers' stores and the consumer, instead of store -> fence -> flag -> poll -> fence -> load.
__device__ __forceinline__ bool tagged_ok(unsigned long long wd, unsigned int tag) { return (unsigned int)(wd >> 32) == tag; }
__device__ __forceinline__ float tagged_val(unsigned long long wd) { return __uint_as_float((unsigned int)(wd & 0xffffffffull)); }

// Canonical row statistic (identical in the stand-alone stage and in the fused GEMM tail, so a row's RMSNorm does not
// depend on which of the two produced it): every 128-column tile publishes its sum of squares - 8 columns per lane
// added in column order, then a 16-lane xor tree - as one tagged 64-bit word {stage tag, value}; a single relaxed
// store / polling load carries flag and payload together, so no fences, atomics or resets are needed.  The row total
// is formed by 16 lanes: lane l adds tiles l, l+16, l+32, ... in order, then the same xor tree.
__device__ __forceinline__ float halfwarp_tree(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float sumsq8(const float (&x)[8]) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) ss += x[j] * x[j];
  return ss;
}
__device__ __forceinline__ void publish_tile_ss(const DevCtx& c, int t, int tile, unsigned int tag, float ss) {
  st_relaxed_gpu_u64(c.row_ss + (size_t)t * kMaxHTiles + tile,
                     ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(ss));
}
// called by whole warps; each 16-lane half works on its own row (valid = the half has a row)
__device__ __forceinline__ float row_total_canonical(const DevCtx& c, int t, int nt_h, unsigned int tag, int l16, bool valid) {
  float p = 0.f;
  if (valid) {
    const unsigned long long* slots = c.row_ss + (size_t)t * kMaxHTiles;
    for (int s2 = l16; s2 < nt_h; s2 += 16) {
      unsigned long long word = ld_relaxed_gpu_u64(slots + s2);
      const long long t0 = clock64();
      while ((unsigned int)(word >> 32) != tag) {
        word = ld_relaxed_gpu_u64(slots + s2);
        if (clock64() - t0 > JF_SPIN_LIMIT_CYCLES) jf_trap_timeout();
      }
      p += __uint_as_float((unsigned int)(word & 0xffffffffull));
    }
  }
  return halfwarp_tree(p);
}

// SPLITS == 0: embedding lookup instead of a reduce.  `ranks` = tensor-parallel ranks whose partial sums are added (rank r's
// slices live in peer memory reached through c.peer_part_rs).  Order of the sum, identical on every rank and in the fused
// tail: each rank's k-splits in split order, then the ranks in rank order.
template <int SPLITS>
__device__ __forceinline__ void rows_norm_impl(const DevCtx& c, const WorkerCtx& w, int T, const __nv_bfloat16* normw,
                                               int which, int ranks) {
  const int h = c.md.h;
  const int nvec = h >> 3;
  const int nt_h = (h + 127) >> 7;
  const int NS = (nvec + kWorkerThreads - 1) / kWorkerThreads;     // slices per row (<= 8), one CTA each
  const int items = T * NS;
  const int l16 = w.lane() & 15;
  prof_mark(c, w, 200);
  for (int item = w.cta(); item < items; item += w.nctas()) {
    const int t = item / NS, sl = item - t * NS;
    const int v = sl * kWorkerThreads + w.wtid();
    __nv_bfloat16* hrow = c.hidden + (size_t)t * h;
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, wt[8];
    if (v < nvec) {
      const uint4 wraw = __ldg(reinterpret_cast<const uint4*>(normw) + v);       // needed last: fetch first
      if (SPLITS == 0) {
        int tk = ld_weak(&c.tok[t]);
        tk = tk < 0 ? 0 : (tk >= c.md.vocab_embed ? c.md.vocab_embed - 1 : tk);
        unpack8(__ldg(reinterpret_cast<const uint4*>(c.embed + (size_t)tk * h) + v), x);
      } else {
        // every k-split slice of this 8-column group is in flight before the first add
        const uint4 rr = ld_weak(reinterpret_cast<const uint4*>(hrow) + v);
        const size_t roff = (size_t)t * c.part_splits * h;
        const size_t sstride4 = (size_t)h >> 2;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < ranks; ++r) {
          const float* src = (ranks == 1) ? c.part_rs[which] : c.peer_part_rs[r][which];
          const float4* p = reinterpret_cast<const float4*>(src + roff) + v * 2;
          float4 pa[SPLITS > 0 ? SPLITS : 1], pb[SPLITS > 0 ? SPLITS : 1];
#pragma unroll
          for (int s2 = 0; s2 < SPLITS; ++s2) {
            pa[s2] = ld_weak(p + s2 * sstride4);
            pb[s2] = ld_weak(p + s2 * sstride4 + 1);
          }
          float sr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int s2 = 0; s2 < SPLITS; ++s2) {
            sr[0] += pa[s2].x; sr[1] += pa[s2].y; sr[2] += pa[s2].z; sr[3] += pa[s2].w;
            sr[4] += pb[s2].x; sr[5] += pb[s2].y; sr[6] += pb[s2].z; sr[7] += pb[s2].w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += sr[j];
        }
        float r[8];
        unpack8(rr, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = bf16_round(r[j] + bf16_round(acc[j]));
      }
      unpack8(wraw, wt);
      reinterpret_cast<uint4*>(hrow)[v] = pack8(x);
    }
    prof_mark(c, w, 201);
    const float tile_ss = halfwarp_tree(sumsq8(x));                 // lanes past the row end contribute zeros
    if (l16 == 0 && (v >> 4) < nt_h) publish_tile_ss(c, t, v >> 4, w.tag(), tile_ss);
    prof_mark(c, w, 202);
    const float total = row_total_canonical(c, t, nt_h, w.tag(), l16, true);
    prof_mark(c, w, 203);
    if (v < nvec) {
      const float rstd = 1.0f / sqrtf(total / (float)h + c.md.eps);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = wt[j] * bf16_round(x[j] * rstd);
      uint8_t* dst = reinterpret_cast<uint8_t*>(c.xn) + ((size_t)(v >> 3) * c.xrows_cap + t) * 128 +
                     (((v & 7) ^ (t & 7)) << 4);
      *reinterpret_cast<uint4*>(dst) = pack8(o);
    }
    prof_mark(c, w, 204);
  }
}

// Poll N tagged slices (`stride` words apart) of one 8-column group until every word carries `tag`, then add them to acc
// in slice order.  All 4 * N 16-byte loads are in flight before the first check.  SYS: the words are written by another
// GPU (into THIS GPU's memory), so the probes are system-scope relaxed loads.
template <int N, bool SYS>
__device__ __forceinline__ void poll_slices_add(const unsigned long long* p, size_t stride, unsigned int tag, float (&acc)[8]) {
  ulonglong2 wd[N][4];
  const long long t0 = clock64();
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int s2 = 0; s2 < N; ++s2)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        wd[s2][k] = SYS ? ld_sys_poll_u64x2(p + s2 * stride + 2 * k) : ld_cg_poll_u64x2(p + s2 * stride + 2 * k);
#pragma unroll
    for (int s2 = 0; s2 < N; ++s2)
#pragma unroll
      for (int k = 0; k < 4; ++k) ok = ok && tagged_ok(wd[s2][k].x, tag) && tagged_ok(wd[s2][k].y, tag);
    if (ok) break;
    if (clock64() - t0 > (SYS ? JF_XSPIN_LIMIT_CYCLES : JF_SPIN_LIMIT_CYCLES)) jf_trap_timeout();
  }
#pragma unroll
  for (int s2 = 0; s2 < N; ++s2)
#pragma unroll
    for (int k = 0; k < 4; ++k) { acc[2 * k] += tagged_val(wd[s2][k].x); acc[2 * k + 1] += tagged_val(wd[s2][k].y); }
}
__device__ __forceinline__ ulonglong2 tag2(unsigned long long tg, uint32_t a, uint32_t b) {
  return make_ulonglong2(tg | (unsigned long long)a, tg | (unsigned long long)b);
}

// The same reduce + residual + RMSNorm executed as the TAIL of a row-split GEMM (O / down projection): once every
// k-split of n-tile `n_tile` has stored its partial sums, the CTAs of those splits share the tile's rows - row step
// `it` (8 rows: one per half-warp) belongs to split it % SPLITS - and each owns the tile's 128 columns of its rows
// (8 columns per lane).  The arithmetic (split order, rounding points, the canonical row statistic) is that of
// rows_norm_impl, bit for bit.
// PUSH (tensor parallelism over R ranks, this GEMM's K dimension is sharded): four phases, all of them driven by tagged
// data, none by a cross-GPU flag, fence or atomic:
//   0. every row: add this rank's k-splits (split order) and store the sum, tagged, into the push buffer of the row's
//      owner (row t belongs to rank t mod R; slice = this rank)                                      [reduce-scatter]
//   1. owned rows: poll the R slices (own memory), add them in rank order, add the residual, publish the tile statistic
//   2. owned rows: row statistic -> normalise; the finished hidden / xn columns go to the own buffers and, tagged, into
//      every peer's staging buffer                                                                   [all-gather]
//   3. rows owned by peers: poll the staging buffer (own memory) and copy hidden / xn into place
// so every rank ends with bit-identical rows and the stage closes with the ordinary LOCAL grid barrier.
template <int SPLITS, bool PUSH>
__device__ __noinline__ void gemm_norm_tail_impl(const DevCtx& c, const WorkerShared* ws, int T, int n_tile, int split,
                                                 const __nv_bfloat16* normw, int which) {
  const WorkerCtx w = load_wctx(ws);
  constexpr int MW = kMaxHTiles / 16;                      // statistic words one lane may poll per row
  const int h = c.md.h;
  const int nt_h = (h + 127) >> 7;
  const int l16 = w.lane() & 15, hw = w.lane() >> 4;
  const int v = n_tile * 16 + l16;                       // 16-byte chunk of the row
  const bool cvalid = v < (h >> 3);
  const int nit = (T + 7) >> 3;
  prof_mark(c, w, 500);
  if (split >= nit) return;
  const int R = PUSH ? c.tp_size : 1, me = PUSH ? c.tp_rank : 0;
  const int cnt = PUSH ? ((T > me) ? (T - me + R - 1) / R : 0) : T;     // rows this rank reduces and normalises
  const int njt = (cnt + 7) >> 3;
  const size_t sstride = (size_t)h;                        // words between the k-split slices of a row
  const unsigned long long* tsrc = c.tpart_rs[which];
  const int trow = w.wwarp() * 2 + hw;                       // this half-warp's row inside a step
  const unsigned long long tg = (unsigned long long)w.tag() << 32;
  uint4 wraw = make_uint4(0u, 0u, 0u, 0u);
  if (cvalid) wraw = __ldg(reinterpret_cast<const uint4*>(normw) + v);
  if (PUSH) {
    // phase 0: this rank's sum over its k-splits of every row -> the owner's push buffer
    for (int it = split; it < nit; it += SPLITS) {
      const int t = it * 8 + trow;
      if (t < T && cvalid) {
        float sr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        poll_slices_add<SPLITS, false>(tsrc + ((size_t)t * c.part_splits) * h + (size_t)v * 8, sstride, w.tag(), sr);
        const int owner = t % R, j = t / R;
        unsigned long long* dst = c.push_buf[owner] + ((size_t)j * R + me) * h + (size_t)v * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<ulonglong2*>(dst + 2 * k) = tag2(tg, __float_as_uint(sr[2 * k]), __float_as_uint(sr[2 * k + 1]));
      }
    }
    prof_mark(c, w, 505);
  }
  // phase 1, per row step: poll the tagged partial sums, add the residual, store the row, publish the tile statistic
  for (int it = split; it < njt; it += SPLITS) {
    const int j = it * 8 + trow;
    const int t = me + j * R;
    const bool live = j < cnt && cvalid;
    uint4 rr = make_uint4(0u, 0u, 0u, 0u);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live) {
      rr = ld_weak(reinterpret_cast<const uint4*>(c.hidden + (size_t)t * h) + v);
      if (!PUSH) {
        poll_slices_add<SPLITS, false>(tsrc + ((size_t)t * c.part_splits) * h + (size_t)v * 8, sstride, w.tag(), acc);
      } else {                                              // rank order: deterministic, identical to rows_norm_impl
        const unsigned long long* p = c.push_buf[me] + ((size_t)j * R) * h + (size_t)v * 8;
        if (R == 2) poll_slices_add<2, true>(p, sstride, w.tag(), acc);
        else {
          poll_slices_add<4, true>(p, sstride, w.tag(), acc);
          if (R == 8) poll_slices_add<4, true>(p + 4 * sstride, sstride, w.tag(), acc);
        }
      }
    }
    prof_mark(c, w, 501);
    float r[8], x[8];
    unpack8(rr, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = bf16_round(r[e] + bf16_round(acc[e]));
    if (live) reinterpret_cast<uint4*>(c.hidden + (size_t)t * h)[v] = pack8(x);
    const float tile_ss = halfwarp_tree(sumsq8(x));
    if (l16 == 0 && j < cnt) publish_tile_ss(c, t, n_tile, w.tag(), tile_ss);
    prof_mark(c, w, 502);
  }
  // phase 2, NB row steps at a time: the statistic words of all NB rows are polled together (one round trip when
  // they are there) and added per row in the canonical order (lane l: tiles l, l+16, ...; 16-lane xor tree); the row
  // itself is read back from `hidden` (this thread's own store) in the same round trip.  NB = 4 for forwards whose CTAs
  // own several row steps; Jacobi-sized forwards (at most one row step per CTA) run the NB = 1 body - a quarter of the code,
  // and this code runs cold every time
  float wt[8];
  unpack8(wraw, wt);
  auto phase2 = [&](auto nbtag) {
  constexpr int NB = decltype(nbtag)::value;
  for (int it0 = split; it0 < njt; it0 += NB * SPLITS) {
    uint4 xr[NB];
    unsigned long long sw[NB][MW];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int j = (it0 + q * SPLITS) * 8 + trow;
      const int t = me + j * R;
      xr[q] = make_uint4(0u, 0u, 0u, 0u);
      if (j < cnt && cvalid) xr[q] = ld_weak(reinterpret_cast<const uint4*>(c.hidden + (size_t)t * h) + v);
    }
    const long long t0 = clock64();
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int j = (it0 + q * SPLITS) * 8 + trow;
        if (j < cnt) {
          const unsigned long long* slots = c.row_ss + (size_t)(me + j * R) * kMaxHTiles;
#pragma unroll
          for (int m = 0; m < MW; ++m)
            if (l16 + 16 * m < nt_h) sw[q][m] = ld_relaxed_gpu_u64(slots + l16 + 16 * m);
        }
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int j = (it0 + q * SPLITS) * 8 + trow;
        if (j < cnt) {
#pragma unroll
          for (int m = 0; m < MW; ++m)
            if (l16 + 16 * m < nt_h) ok = ok && tagged_ok(sw[q][m], w.tag());
        }
      }
      if (ok) break;
      if (clock64() - t0 > JF_SPIN_LIMIT_CYCLES) jf_trap_timeout();
    }
    prof_mark(c, w, 503);
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int it = it0 + q * SPLITS;
      const int j = it * 8 + trow;
      const int t = me + j * R;
      if (it < njt) {                                       // whole warps take part in the reduction of their rows
        float pj = 0.f;
        if (j < cnt) {
#pragma unroll
          for (int m = 0; m < MW; ++m)
            if (l16 + 16 * m < nt_h) pj += tagged_val(sw[q][m]);
        }
        const float total = halfwarp_tree(pj);
        if (j < cnt && cvalid) {
          const float rstd = 1.0f / sqrtf(total / (float)h + c.md.eps);
          float x[8], o[8];
          unpack8(xr[q], x);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = wt[e] * bf16_round(x[e] * rstd);
          uint8_t* dst = reinterpret_cast<uint8_t*>(c.xn) + ((size_t)(v >> 3) * c.xrows_cap + t) * 128 +
                         (((v & 7) ^ (t & 7)) << 4);
          const uint4 on = pack8(o);
          *reinterpret_cast<uint4*>(dst) = on;
          if (PUSH) {                                       // the finished row -> every peer's staging buffer
            const size_t soff = ((size_t)t * (h >> 3) + v) * 8;
            for (int r2 = 0; r2 < R; ++r2) {
              if (r2 == me) continue;
              unsigned long long* sd = c.stage_buf[r2] + soff;
              *reinterpret_cast<ulonglong2*>(sd) = tag2(tg, xr[q].x, xr[q].y);
              *reinterpret_cast<ulonglong2*>(sd + 2) = tag2(tg, xr[q].z, xr[q].w);
              *reinterpret_cast<ulonglong2*>(sd + 4) = tag2(tg, on.x, on.y);
              *reinterpret_cast<ulonglong2*>(sd + 6) = tag2(tg, on.z, on.w);
            }
          }
        }
      }
    }
  }
  };
  if (njt <= SPLITS) phase2(std::integral_constant<int, 1>{}); else phase2(std::integral_constant<int, 4>{});
  prof_mark(c, w, 504);
  if (PUSH) {
    // phase 3: the rows the peers own arrive, finished, in this rank's staging buffer
    for (int it = split; it < nit; it += SPLITS) {
      const int t = it * 8 + trow;
      if (t < T && cvalid && (t % R) != me) {
        const unsigned long long* sp = c.stage_buf[me] + ((size_t)t * (h >> 3) + v) * 8;
        ulonglong2 wd[4];
        const long long t0 = clock64();
        for (;;) {
          bool ok = true;
#pragma unroll
          for (int k = 0; k < 4; ++k) wd[k] = ld_sys_poll_u64x2(sp + 2 * k);
#pragma unroll
          for (int k = 0; k < 4; ++k) ok = ok && tagged_ok(wd[k].x, w.tag()) && tagged_ok(wd[k].y, w.tag());
          if (ok) break;
          if (clock64() - t0 > JF_XSPIN_LIMIT_CYCLES) jf_trap_timeout();
        }
        const uint4 hv = make_uint4((uint32_t)wd[0].x, (uint32_t)wd[0].y, (uint32_t)wd[1].x, (uint32_t)wd[1].y);
        const uint4 xv = make_uint4((uint32_t)wd[2].x, (uint32_t)wd[2].y, (uint32_t)wd[3].x, (uint32_t)wd[3].y);
        reinterpret_cast<uint4*>(c.hidden + (size_t)t * h)[v] = hv;
        uint8_t* dst = reinterpret_cast<uint8_t*>(c.xn) + ((size_t)(v >> 3) * c.xrows_cap + t) * 128 +
                       (((v & 7) ^ (t & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = xv;
      }
    }
    prof_mark(c, w, 506);
  }
}

// push = this projection's K dimension is sharded over the tensor-parallel ranks
__device__ __forceinline__ void gemm_norm_tail(const DevCtx& c, const WorkerShared* w, int T, int n_tile, int split,
                                            int splits, const __nv_bfloat16* normw, int which, bool push) {
#define JF_TAIL_CASE(S) case S: if (push) gemm_norm_tail_impl<S, true>(c, w, T, n_tile, split, normw, which); \
                                else gemm_norm_tail_impl<S, false>(c, w, T, n_tile, split, normw, which); break;
  switch (splits) {
    JF_TAIL_CASE(1) JF_TAIL_CASE(2) JF_TAIL_CASE(3) JF_TAIL_CASE(4) JF_TAIL_CASE(5) JF_TAIL_CASE(6) JF_TAIL_CASE(7)
    default: if (push) gemm_norm_tail_impl<8, true>(c, w, T, n_tile, split, normw, which);
             else gemm_norm_tail_impl<8, false>(c, w, T, n_tile, split, normw, which);
             break;
  }
#undef JF_TAIL_CASE
}

// SwiGLU as the TAIL of the split-K gate/up projection (EPI_SWIGLU_TAGGED): the CTAs that hold an n-tile's k-splits
// share the tile's rows (row step `it` belongs to split it % SPLITS; one row per half-warp, 8 features per lane), add the
// splits' gate and up partial sums in split order and write silu(gate) * up into the down projection's operand.
// Rounding points as in the single-accumulator epilogue (HF Qwen2MLP: every op in bf16).
template <int SPLITS>
__device__ __noinline__ void gu_swiglu_tail_impl(const DevCtx& c, const WorkerShared* ws, int T, int n_tile, int split) {
  const WorkerCtx w = load_wctx(ws);
  const int N = c.md.inter;
  const int l16 = w.lane() & 15, hw = w.lane() >> 4;
  const int n0 = n_tile * kTileN + l16 * 8;
  const int nit = (T + 7) >> 3;
  const int trow = w.wwarp() * 2 + hw;
  if (n0 >= N) return;
  for (int it = split; it < nit; it += SPLITS) {
    const int t = it * 8 + trow;
    if (t >= T) continue;
    const unsigned long long* p = c.tgu + ((size_t)t * SPLITS) * 2 * N + n0;
    float gs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, us[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      ulonglong2 wd[SPLITS][4];
      const long long t0 = clock64();
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int s2 = 0; s2 < SPLITS; ++s2)
#pragma unroll
          for (int k = 0; k < 4; ++k) wd[s2][k] = ld_cg_poll_u64x2(p + ((size_t)s2 * 2 + g) * N + 2 * k);
#pragma unroll
        for (int s2 = 0; s2 < SPLITS; ++s2)
#pragma unroll
          for (int k = 0; k < 4; ++k) ok = ok && tagged_ok(wd[s2][k].x, w.tag()) && tagged_ok(wd[s2][k].y, w.tag());
        if (ok) break;
        if (clock64() - t0 > JF_SPIN_LIMIT_CYCLES) jf_trap_timeout();
      }
#pragma unroll
      for (int s2 = 0; s2 < SPLITS; ++s2)                   // split order: deterministic
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (g == 0) { gs[2 * k] += tagged_val(wd[s2][k].x); gs[2 * k + 1] += tagged_val(wd[s2][k].y); }
          else { us[2 * k] += tagged_val(wd[s2][k].x); us[2 * k + 1] += tagged_val(wd[s2][k].y); }
        }
    }
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gf = bf16_round(gs[e]);
      const float uf = bf16_round(us[e]);
      const float sl = bf16_round(gf / (1.f + __expf(-gf)));
      o[e] = sl * uf;
    }
    uint8_t* dst = reinterpret_cast<uint8_t*>(c.xact) + ((size_t)(n0 >> 6) * c.xrows_cap) * 128 + sw128_offset(t, n0 & 63);
    *reinterpret_cast<uint4*>(dst) = pack8(o);
  }
}
__device__ __forceinline__ void gu_swiglu_tail(const DevCtx& c, const WorkerShared* ws, int T, int n_tile, int split, int splits) {
  switch (splits) {
    case 2: gu_swiglu_tail_impl<2>(c, ws, T, n_tile, split); break;
    case 3: gu_swiglu_tail_impl<3>(c, ws, T, n_tile, split); break;
    case 4: gu_swiglu_tail_impl<4>(c, ws, T, n_tile, split); break;
    case 5: gu_swiglu_tail_impl<5>(c, ws, T, n_tile, split); break;
    case 6: gu_swiglu_tail_impl<6>(c, ws, T, n_tile, split); break;
    case 7: gu_swiglu_tail_impl<7>(c, ws, T, n_tile, split); break;
    default: gu_swiglu_tail_impl<8>(c, ws, T, n_tile, split); break;
  }
}

// which: the row-split partial buffer to reduce (0 / 1 / 2, see DevCtx::part_rs); ranks: tensor-parallel ranks that hold
// partial sums of this projection (1 = local)
__device__ __noinline__ void stage_rows_norm(const DevCtx& c, const WorkerShared* ws, int T, bool embed, int splits,
                                             const __nv_bfloat16* normw, int which, int ranks) {
  const WorkerCtx w = load_wctx(ws);
  if (embed) { rows_norm_impl<0>(c, w, T, normw, which, 1); return; }
#define JF_NORM_CASE(S) case S: rows_norm_impl<S>(c, w, T, normw, which, ranks); break;
  switch (splits) {
    JF_NORM_CASE(1) JF_NORM_CASE(2) JF_NORM_CASE(3) JF_NORM_CASE(4) JF_NORM_CASE(5) JF_NORM_CASE(6) JF_NORM_CASE(7)
    default: rows_norm_impl<8>(c, w, T, normw, which, ranks); break;
  }
#undef JF_NORM_CASE
}

// ---------------------------------------------------------------------------------------------------------------
// q/k/v: split-K reduce + bias -> bf16, rotary embedding in bf16 arithmetic, scatter to qbuf / in-flight KV scratch.
// One warp per (row, head slot); two items are processed together so both items' loads are in flight at once.
// ---------------------------------------------------------------------------------------------------------------
// Items are (row, head slot) pairs of the slot range [slot_base, slot_base + nslots); this warp takes items first,
// first + stride, ...  (stand-alone stage: every slot, warps of the whole grid; fused tail: the slots of one n-tile,
// the warps of the CTAs that hold the tile's k-splits).
template <int SPLITS, int PPL, bool TAGGED>
__device__ __forceinline__ void qkv_rope_impl(const DevCtx& c, const WorkerCtx& w, int T, int layer, int first, int stride,
                                              int slot_base, int nslots) {
  const ModelDims& md = c.md;
  const int d = md.d, half = d >> 1;
  const float* bias = c.layer[layer].bqkv;
  const int items = T * nslots;
  const size_t sstride = (size_t)c.part_n;
  prof_mark(c, w, 100);
  for (int item0 = first; item0 < items; item0 += 2 * stride) {
    float lo[2][2], hi[2][2];
    int tt[2], slot[2], pos[2];
    const float* p0[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int item = item0 + q * stride;
      tt[q] = item < items ? 1 : -1;
      if (item >= items) item = item0;                    // keep the second lane of work well-formed; result unused
      const int t = item / nslots;
      if (tt[q] > 0) tt[q] = t;
      slot[q] = slot_base + item - t * nslots;
      pos[q] = w.plan() ? w.plan()->row_pos[t] : ld_weak(&c.row_pos[t]);
      p0[q] = c.part + (size_t)t * c.part_splits * c.part_n + slot[q] * d + w.lane() * PPL;
    }
    // bias and rotary factors do not depend on the projection: fetched BEFORE the partial sums are polled (one L2 round trip less)
    float bl[2][PPL], bh[2][PPL], rc[2][PPL], rs[2][PPL];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int col0 = slot[q] * d;
      int p = pos[q];
      p = p < 0 ? 0 : (p >= md.max_pos ? md.max_pos - 1 : p);
#pragma unroll
      for (int j = 0; j < PPL; ++j) {
        const int e = w.lane() * PPL + j;
        bl[q][j] = __ldg(bias + col0 + e);
        bh[q][j] = __ldg(bias + col0 + half + e);
        rc[q][j] = 1.f; rs[q][j] = 0.f;
        if (slot[q] < md.n_q + md.n_kv) {
          rc[q][j] = __bfloat162float(__ldg(c.rope_cos + (size_t)p * half + e));
          rs[q][j] = __bfloat162float(__ldg(c.rope_sin + (size_t)p * half + e));
        }
      }
    }
    if (TAGGED) {
      // fused tail: poll the tagged partial sums (2 * PPL words per item and split) until all of them are this stage's
      const unsigned long long* tp[2];
#pragma unroll
      for (int q = 0; q < 2; ++q)
        tp[q] = c.tpart + (size_t)(tt[q] < 0 ? 0 : tt[q]) * c.part_splits * c.part_n + slot[q] * d + w.lane() * PPL;
      unsigned long long wl[2][SPLITS][PPL], wh[2][SPLITS][PPL];
      const long long t0 = clock64();
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int s2 = 0; s2 < SPLITS; ++s2) {
            if constexpr (PPL == 2) {
              const ulonglong2 a = ld_cg_poll_u64x2(tp[q] + s2 * sstride);
              const ulonglong2 b = ld_cg_poll_u64x2(tp[q] + s2 * sstride + half);
              wl[q][s2][0] = a.x; wl[q][s2][PPL - 1] = a.y; wh[q][s2][0] = b.x; wh[q][s2][PPL - 1] = b.y;
            } else {
              wl[q][s2][0] = ld_cg_poll_u64(tp[q] + s2 * sstride);
              wh[q][s2][0] = ld_cg_poll_u64(tp[q] + s2 * sstride + half);
            }
          }
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int s2 = 0; s2 < SPLITS; ++s2)
#pragma unroll
            for (int j = 0; j < PPL; ++j)
              ok = ok && (tt[q] < 0 || (tagged_ok(wl[q][s2][j], w.tag()) && tagged_ok(wh[q][s2][j], w.tag())));
        if (ok) break;
        if (clock64() - t0 > JF_SPIN_LIMIT_CYCLES) jf_trap_timeout();
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        lo[q][0] = lo[q][1] = hi[q][0] = hi[q][1] = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < SPLITS; ++s2)               // slice order: deterministic
#pragma unroll
          for (int j = 0; j < PPL; ++j) { lo[q][j] += tagged_val(wl[q][s2][j]); hi[q][j] += tagged_val(wh[q][s2][j]); }
      }
    } else if (PPL == 2) {
      float2 vl[2][SPLITS], vh[2][SPLITS];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int s2 = 0; s2 < SPLITS; ++s2) {
          vl[q][s2] = ld_weak(reinterpret_cast<const float2*>(p0[q] + s2 * sstride));
          vh[q][s2] = ld_weak(reinterpret_cast<const float2*>(p0[q] + s2 * sstride + half));
        }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        lo[q][0] = lo[q][1] = hi[q][0] = hi[q][1] = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < SPLITS; ++s2) {             // slice order: deterministic
          lo[q][0] += vl[q][s2].x; lo[q][1] += vl[q][s2].y; hi[q][0] += vh[q][s2].x; hi[q][1] += vh[q][s2].y;
        }
      }
    } else {
      float vl[2][SPLITS], vh[2][SPLITS];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int s2 = 0; s2 < SPLITS; ++s2) {
          vl[q][s2] = ld_weak(p0[q] + s2 * sstride);
          vh[q][s2] = ld_weak(p0[q] + s2 * sstride + half);
        }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        lo[q][0] = lo[q][1] = hi[q][0] = hi[q][1] = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < SPLITS; ++s2) { lo[q][0] += vl[q][s2]; hi[q][0] += vh[q][s2]; }
      }
    }
    prof_mark(c, w, 101);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (tt[q] < 0) continue;
      const int t = tt[q], col0 = slot[q] * d;
      __nv_bfloat16* dst;
      if (slot[q] < md.n_q) dst = c.qbuf + (size_t)t * md.q_dim + col0;
      else if (slot[q] < md.n_q + md.n_kv) dst = c.kv_scratch + kv_scratch_offset(md, c.T_cap, layer, 0, t, slot[q] - md.n_q);
      else dst = c.kv_scratch + kv_scratch_offset(md, c.T_cap, layer, 1, t, slot[q] - md.n_q - md.n_kv);
#pragma unroll
      for (int j = 0; j < PPL; ++j) {
        const int e = w.lane() * PPL + j;
        float l = bf16_round(lo[q][j] + bl[q][j]);
        float hv = bf16_round(hi[q][j] + bh[q][j]);
        if (slot[q] < md.n_q + md.n_kv) {
          const float cs = rc[q][j], sn = rs[q][j];
          const float nlo = bf16_round(bf16_round(l * cs) - bf16_round(hv * sn));
          const float nhi = bf16_round(bf16_round(hv * cs) + bf16_round(l * sn));
          l = nlo; hv = nhi;
        }
        dst[e] = __float2bfloat16_rn(l);
        dst[half + e] = __float2bfloat16_rn(hv);
      }
    }
  }
  prof_mark(c, w, 102);
}

__device__ __noinline__ void stage_qkv_rope(const DevCtx& c, const WorkerShared* ws, int T, int layer) {
  const WorkerCtx w = load_wctx(ws);
  const int splits = c.splits_qkv;
  const int first = w.cta() * 4 + w.wwarp(), stride = w.nctas() * 4, nslots = c.md.n_q + 2 * c.md.n_kv;
#define JF_ROPE_CASE(S)                                                                   \
  case S:                                                                                 \
    if (c.md.d == 128) qkv_rope_impl<S, 2, false>(c, w, T, layer, first, stride, 0, nslots); \
    else qkv_rope_impl<S, 1, false>(c, w, T, layer, first, stride, 0, nslots);                    \
    break;
  switch (splits) {
    JF_ROPE_CASE(1) JF_ROPE_CASE(2) JF_ROPE_CASE(3) JF_ROPE_CASE(4) JF_ROPE_CASE(5) JF_ROPE_CASE(6) JF_ROPE_CASE(7)
    default: JF_ROPE_CASE(8)
  }
#undef JF_ROPE_CASE
}

// The same bias + rotary + scatter as the TAIL of the QKV GEMM: once every k-split of n-tile `n_tile` has stored its
// partial sums, the 4 * splits worker warps of those CTAs share the tile's (row, head slot) items.
__device__ __noinline__ void qkv_rope_tail(const DevCtx& c, const WorkerShared* ws, int T, int layer, int n_tile, int split,
                                           int splits) {
  const WorkerCtx w = load_wctx(ws);
  const int hpt = kTileN / c.md.d;                         // head slots per 128-feature tile (1 or 2)
  const int slot_base = n_tile * hpt;
  const int nslots = min(hpt, c.md.n_q + 2 * c.md.n_kv - slot_base);
  if (nslots <= 0) return;
  const int first = split * 4 + w.wwarp(), stride = splits * 4;
#define JF_ROPE_CASE(S)                                                                   \
  case S:                                                                                 \
    if (c.md.d == 128) qkv_rope_impl<S, 2, true>(c, w, T, layer, first, stride, slot_base, nslots); \
    else qkv_rope_impl<S, 1, true>(c, w, T, layer, first, stride, slot_base, nslots);            \
    break;
  switch (splits) {
    JF_ROPE_CASE(1) JF_ROPE_CASE(2) JF_ROPE_CASE(3) JF_ROPE_CASE(4) JF_ROPE_CASE(5) JF_ROPE_CASE(6) JF_ROPE_CASE(7)
    default: JF_ROPE_CASE(8)
  }
#undef JF_ROPE_CASE
}

// ---------------------------------------------------------------------------------------------------------------
// attention: unit = (group, kv head, key chunk, query split); mma.sync m16n8k16, online softmax over 64-key sub-tiles
// Chunks are aligned to ABSOLUTE key positions, so a query's result does not depend on how its keys are split between
// committed pages and in-flight scratch rows (AR decoding and Jacobi decoding of the same prefix agree bit for bit).
// ---------------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void attention_unit(const DevCtx& c, const WorkerCtx& w, int layer, const Group& grp, int g_kvh, int chunk,
                               int qs, int hs, bool prefetched, int hsplit) {
  constexpr int LDS = D + 8;                  // padded smem row (bf16 elements) -> conflict-free fragment loads
  constexpr int VPR = D / 8;                  // 16-byte chunks per row
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(w.attn_smem());
  __nv_bfloat16* Vs = Ks + kAttnSub * LDS;
  const uint32_t ks_addr = smem_u32(Ks), vs_addr = smem_u32(Vs);
  const ModelDims& md = c.md;
  const int ctx = grp.ctx, len = grp.len, row0 = grp.row0;
  const int nkeys = ctx + len;
  const int key_lo = chunk * c.attn_chunk;
  const int key_hi = min(key_lo + c.attn_chunk, nkeys);
  const int q0 = qs * kQTokPerUnit;
  const int qn = min(kQTokPerUnit, len - q0);
  const int ntt = (qn + 15) >> 4;
  const int hp = (md.n_rep + hsplit - 1) / hsplit;      // q heads of this kv head handled by one unit
  const int j_lo = hs * hp, j_hi = min(md.n_rep, j_lo + hp);
  const int tiles = max(0, j_hi - j_lo) * ntt;
  const int nsub = (key_hi - key_lo + kAttnSub - 1) / kAttnSub;
  const int* bt = c.block_tables + (size_t)grp.seq * c.max_pages_per_seq;
  const float scale = rsqrtf((float)D);
  const int r0 = w.lane() >> 2, c2 = (w.lane() & 3) * 2;

  // K/V sub-tile fill: asynchronous 16-byte copies, one page-table lookup per 64-key sub-tile (kPage % kAttnSub == 0)
  // skip_committed: the committed keys of this sub-tile are already in flight / in place (attention_prefetch)
  auto fill_issue = [&](int k0, bool skip_committed) {
    if (skip_committed && k0 + kAttnSub <= ctx) return;
    const __nv_bfloat16* cbase[2] = {nullptr, nullptr};
    if (k0 < ctx && !skip_committed) {
      const int page = ld_weak(bt + (k0 / kPage));
      cbase[0] = c.kv_cache + kv_cache_offset(md, c.num_pages, layer, 0, page, g_kvh, k0 % kPage);
      cbase[1] = c.kv_cache + kv_cache_offset(md, c.num_pages, layer, 1, page, g_kvh, k0 % kPage);
    }
    const __nv_bfloat16* sbase[2] = {c.kv_scratch + kv_scratch_offset(md, c.T_cap, layer, 0, row0, g_kvh),
                                     c.kv_scratch + kv_scratch_offset(md, c.T_cap, layer, 1, row0, g_kvh)};
    const size_t srow = (size_t)md.n_kv * D;             // scratch stride between consecutive tokens
#pragma unroll 4
    for (int idx = w.wtid(); idx < kAttnSub * VPR * 2; idx += kWorkerThreads) {
      const int kv = idx / (kAttnSub * VPR);
      const int rem = idx - kv * (kAttnSub * VPR);
      const int jr = rem / VPR, v = rem - jr * VPR;
      const int pos = k0 + jr;
      const __nv_bfloat16* src = sbase[kv];
      uint32_t nbytes = 0;
      if (pos < ctx) {
        if (skip_committed) continue;
        src = cbase[kv] + (size_t)jr * D + v * 8; nbytes = 16;
      } else if (pos < nkeys) { src = sbase[kv] + (size_t)(pos - ctx) * srow + v * 8; nbytes = 16; }
      cp_async_16((kv ? Vs : Ks) + jr * LDS + v * 8, src, nbytes);
    }
    cp_async_commit();
  };

  if (tiles <= 0) return;
  prof_mark(c, w, 400);
  if (nsub == 1) {                                         // the common case: issue the fill first, overlap it with Q loads
    if (!prefetched) workers_sync();                       // previous readers of Ks / Vs are done
    fill_issue(key_lo, prefetched);
  }
  // Q fragments of one tile (head j, 16 tokens): straight from qbuf (written by the rope stage, L2-resident)
  auto load_q = [&](int tile, uint32_t (&q)[D / 16][4]) {
    if (tile >= tiles) return;
    const int j = j_lo + tile / ntt;
    const int i0 = q0 + (tile % ntt) * 16;
    if (key_lo > ctx + min(i0 + 15, len - 1)) return;      // inactive tile (no visible key in this chunk)
    const int head = g_kvh * md.n_rep + j;
    const __nv_bfloat16* qa = c.qbuf + (size_t)(row0 + min(i0 + r0, len - 1)) * md.q_dim + head * D;
    const __nv_bfloat16* qb = c.qbuf + (size_t)(row0 + min(i0 + r0 + 8, len - 1)) * md.q_dim + head * D;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      q[ks][0] = ld_weak(reinterpret_cast<const unsigned int*>(qa + ks * 16 + c2));
      q[ks][1] = ld_weak(reinterpret_cast<const unsigned int*>(qb + ks * 16 + c2));
      q[ks][2] = ld_weak(reinterpret_cast<const unsigned int*>(qa + ks * 16 + 8 + c2));
      q[ks][3] = ld_weak(reinterpret_cast<const unsigned int*>(qb + ks * 16 + 8 + c2));
    }
  };
  uint32_t qnx[D / 16][4];
  for (int tb = 0; tb * 4 < tiles; ++tb) {
    const int tile = tb * 4 + w.wwarp();
    const bool tvalid = tile < tiles;
    const int j = j_lo + (tvalid ? tile / ntt : 0);
    const int tt = tvalid ? tile % ntt : 0;
    const int i0 = q0 + tt * 16;
    const int head = g_kvh * md.n_rep + j;
    const int ia = i0 + r0, ib = i0 + r0 + 8;          // token indices of this thread's two fragment rows
    // the tile takes part only if some key of the chunk is visible to its last token
    const bool tactive = tvalid && (key_lo <= ctx + min(i0 + 15, len - 1));

    uint32_t qf[D / 16][4];
    if (tb == 0) load_q(tile, qf);
    else {
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) { qf[ks][0] = qnx[ks][0]; qf[ks][1] = qnx[ks][1]; qf[ks][2] = qnx[ks][2]; qf[ks][3] = qnx[ks][3]; }
    }
    if ((tb + 1) * 4 < tiles) load_q(tile + 4, qnx);       // next round's queries are in flight while this round computes
    if (nsub == 1 && tb == 0) {
      cp_async_wait_all();
      workers_sync();
      prof_mark(c, w, 401);
    }
    float o[D / 8][4];
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) { o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f; }
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

    for (int sub = 0; sub < nsub; ++sub) {
      const int k0 = key_lo + sub * kAttnSub;
      if (nsub > 1) {
        workers_sync();
        fill_issue(k0, prefetched && sub == 0 && tb == 0);
        cp_async_wait_all();
        workers_sync();
      }
      if (!tactive || k0 > ctx + min(i0 + 15, len - 1)) continue;

      float s[8][4];
      // K fragments with ldmatrix: matrix m of one x4 = keys nt*8..+7 (rows) x dims (ks2*32 + m*8)..+7, i.e. (b0, b1) of
      // k-steps 2*ks2 and 2*ks2+1 for n-tile nt
      const uint32_t k_lane = ks_addr + (uint32_t)(((w.lane() & 7) * LDS + (w.lane() >> 3) * 8) * 2);
      // k-step outer, key tile inner: eight independent accumulator chains in flight (the mma latency, not its
      // issue rate, bounds a lone warp per scheduler)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
      for (int ks2 = 0; ks2 < D / 32; ++ks2) {
        uint32_t kb[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          ldmatrix_x4(k_lane + (uint32_t)((nt * 8 * LDS + ks2 * 32) * 2), kb[nt][0], kb[nt][1], kb[nt][2], kb[nt][3]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mma_bf16_16816(s[nt], qf[2 * ks2], kb[nt][0], kb[nt][1]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mma_bf16_16816(s[nt], qf[2 * ks2 + 1], kb[nt][2], kb[nt][3]);
      }
      prof_mark(c, w, 402);
      // scale, causal mask, running max
      const int qpa = ctx + ia, qpb = ctx + ib;
      float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int kp = k0 + nt * 8 + c2;
        s[nt][0] = (kp <= qpa) ? s[nt][0] * scale : -INFINITY;
        s[nt][1] = (kp + 1 <= qpa) ? s[nt][1] * scale : -INFINITY;
        s[nt][2] = (kp <= qpb) ? s[nt][2] * scale : -INFINITY;
        s[nt][3] = (kp + 1 <= qpb) ? s[nt][3] * scale : -INFINITY;
        mxa = fmaxf(mxa, fmaxf(s[nt][0], s[nt][1]));
        mxb = fmaxf(mxb, fmaxf(s[nt][2], s[nt][3]));
      }
      mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 1));
      mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 2));
      mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 1));
      mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 2));
      const float mna = fmaxf(m_a, mxa), mnb = fmaxf(m_b, mxb);
      const float sca = (mna == -INFINITY) ? 1.f : __expf(m_a - mna);
      const float scb = (mnb == -INFINITY) ? 1.f : __expf(m_b - mnb);
      const float mua = (mna == -INFINITY) ? 0.f : mna;
      const float mub = (mnb == -INFINITY) ? 0.f : mnb;
      m_a = mna; m_b = mnb;
      float suma = 0.f, sumb = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = __expf(s[nt][0] - mua); s[nt][1] = __expf(s[nt][1] - mua);
        s[nt][2] = __expf(s[nt][2] - mub); s[nt][3] = __expf(s[nt][3] - mub);
        suma += s[nt][0] + s[nt][1];
        sumb += s[nt][2] + s[nt][3];
      }
      l_a = l_a * sca + suma;
      l_b = l_b * scb + sumb;
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd) { o[nd][0] *= sca; o[nd][1] *= sca; o[nd][2] *= scb; o[nd][3] *= scb; }
      // O += P V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        // V fragments with ldmatrix.trans: matrix m = keys kk*16 + (m&1)*8..+7 (rows) x dims (nd2*16 + (m>>1)*8)..+7,
        // transposed on load -> (b0, b1) of dim tiles 2*nd2 and 2*nd2+1 for key step kk
        const uint32_t v_lane = vs_addr + (uint32_t)((((kk * 16 + (w.lane() & 7) + ((w.lane() >> 3) & 1) * 8) * LDS) +
                                                     (w.lane() >> 4) * 8) * 2);
        uint32_t vb[D / 16][4];
#pragma unroll
        for (int nd2 = 0; nd2 < D / 16; ++nd2)
          ldmatrix_x4_trans(v_lane + (uint32_t)(nd2 * 16 * 2), vb[nd2][0], vb[nd2][1], vb[nd2][2], vb[nd2][3]);
#pragma unroll
        for (int nd2 = 0; nd2 < D / 16; ++nd2) {
          mma_bf16_16816(o[2 * nd2], pa, vb[nd2][0], vb[nd2][1]);
          mma_bf16_16816(o[2 * nd2 + 1], pa, vb[nd2][2], vb[nd2][3]);
        }
      }
    }

    prof_mark(c, w, 403);
    if (tactive) {
      l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
      l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
      l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
      l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
      if (ia < len) {
        const size_t sh = ((size_t)(grp.slot0 + ia * grp.nchunks + chunk) * md.n_q + head);
        float* op = c.attn_o + sh * D + c2;
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) *reinterpret_cast<float2*>(op + nd * 8) = make_float2(o[nd][0], o[nd][1]);
        if ((w.lane() & 3) == 0) *reinterpret_cast<float2*>(c.attn_ml + sh * 2) = make_float2(m_a, l_a);
      }
      if (ib < len) {
        const size_t sh = ((size_t)(grp.slot0 + ib * grp.nchunks + chunk) * md.n_q + head);
        float* op = c.attn_o + sh * D + c2;
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) *reinterpret_cast<float2*>(op + nd * 8) = make_float2(o[nd][2], o[nd][3]);
        if ((w.lane() & 3) == 0) *reinterpret_cast<float2*>(c.attn_ml + sh * 2) = make_float2(m_b, l_b);
      }
    }
  }
}

__device__ __forceinline__ Group load_group(const DevCtx& c, int g) {
  Group grp;
  const int4* p = reinterpret_cast<const int4*>(c.groups + g);
  const int4 a = ld_weak(p), b = ld_weak(p + 1);
  grp.row0 = a.x; grp.len = a.y; grp.seq = a.z; grp.ctx = a.w;
  grp.nchunks = b.x; grp.qsplit = b.y; grp.unit0 = b.z; grp.slot0 = b.w;
  return grp;
}

__device__ __forceinline__ AttnUnitId decode_attn_unit(const DevCtx& c, const WorkerCtx& w, int u, int G, int hsplit) {
  if (u == w.cta() && w.plan() && w.plan()->first_unit_ok) return w.plan()->first_unit;
  AttnUnitId id;
  int g = 0;
  id.grp = w.plan() ? w.plan()->groups[0] : load_group(c, 0);
  while (g + 1 < G) {                                   // groups are few; linear scan over unit ranges
    const int per = c.md.n_kv * id.grp.nchunks * id.grp.qsplit * hsplit;
    if (u < id.grp.unit0 + per) break;
    ++g;
    id.grp = w.plan() ? w.plan()->groups[g] : load_group(c, g);
  }
  int local = u - id.grp.unit0;
  id.kvh = local % c.md.n_kv; local /= c.md.n_kv;
  id.hs = local % hsplit; local /= hsplit;
  id.chunk = local % id.grp.nchunks;
  id.qs = local / id.grp.nchunks;
  return id;
}

// ---------------------------------------------------------------------------------------------------------------
// tcgen05 attention: the same unit (group, kv head, 64-key chunk, query split, head split) on the 5th-gen tensor cores.
//   S = Q K^T   A = the unit's (head, query row) pairs, up to 128 per tile, staged in TENSOR MEMORY straight from registers
//               (lane = pair, 32-bit column j = dims 2j, 2j+1); B = the K chunk in shared memory, K-major SWIZZLE_128B
//   softmax     one thread per pair reads its 64 scores with tcgen05.ld: scale, causal mask, max, exp, sum - no shuffles
//   O = P V     A = P (bf16) written back to tensor memory with tcgen05.st; B = the V chunk, the SAME [key][dim] rows in
//               shared memory read as an MN-major operand (no transposed copy of V anywhere)
// The accumulators live in the tensor memory the GEMM stages use, idle while this stage runs.  K and V tiles are two
// 64-dim blocks of [64 keys][128 B] each, 16-byte chunk c of key r at chunk position c ^ (r & 7); they fill the attention
// buffer of the mma.sync path (1024-byte aligned inside it).  Results are written exactly like the mma.sync path writes them
// (unnormalised fp32 partial outputs + (max, sum) per pair and chunk), so the combine stage is shared.
// Rows of a result depend only on that row's query and the chunk's keys, and chunks are aligned to absolute key positions:
// Jacobi, autoregressive and prefill forwards still agree bit for bit.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kTcQ = 0, kTcS = 64, kTcP = 128, kTcO = 256;     // tensor-memory columns: Q (D/2), S (64), P (32), O (D)

template <int D>
__device__ __forceinline__ uint32_t tc_kv_base(const WorkerCtx& w) {
  return (smem_u32(w.attn_smem()) + 1023u) & ~1023u;
}
// K (kv = 0) / V (kv = 1) tile address of 16-byte chunk `v` (8 dims) of key row `jr`
template <int D>
__device__ __forceinline__ uint32_t tc_kv_addr(uint32_t base, int kv, int jr, int v) {
  constexpr uint32_t kTile = 64u * D * 2u;                               // bytes of one K or V tile
  return base + (uint32_t)kv * kTile + (uint32_t)(v >> 3) * 8192u + (uint32_t)jr * 128u + (uint32_t)(((v & 7) ^ (jr & 7)) << 4);
}
__device__ __forceinline__ void cp_async_16_saddr(uint32_t saddr, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gmem_src), "r"(src_bytes) : "memory");
}

#define JF_DBG(slot, val) do { if (c.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { c.dbg[(int)(threadIdx.x >> 5) * 8 + (slot)] = (val); __threadfence_system(); } } while (0)
// bounded wait on the attention MMA barrier: a timeout is recorded in the plan header (the host turns it into an error)
// instead of trapping, so a wrong descriptor shows up as a diagnosable failure rather than a dead context
__device__ __forceinline__ bool attn_bar_wait(const DevCtx& c, uint64_t* bar, uint32_t phase, int code) {
  if (mbar_try_wait(bar, phase)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, phase)) {
    if (clock64() - t0 > 400000000LL) {                      // ~0.2 s
      if ((threadIdx.x & 31) == 0) atomicMax(&c.hdr->error, code);
      return false;
    }
  }
  return true;
}

template <int D>
__device__ __forceinline__ void attention_unit_tc(const DevCtx& c, const WorkerCtx& w, const WorkerShared* ws, uint32_t& phase,
                                                  int layer, const Group& grp, int g_kvh, int chunk, int qs, int hs,
                                                  bool prefetched, int hsplit) {
  constexpr int VPR = D / 8;                  // 16-byte chunks per K / V row
  const ModelDims& md = c.md;
  const int ctx = grp.ctx, len = grp.len, row0 = grp.row0;
  const int nkeys = ctx + len;
  const int key_lo = chunk * kAttnSub;
  const int q0 = qs * kQTokPerUnit;
  const int qn = min(kQTokPerUnit, len - q0);
  const int hp = (md.n_rep + hsplit - 1) / hsplit;
  const int j_lo = hs * hp, j_hi = min(md.n_rep, j_lo + hp);
  const int nheads = j_hi - j_lo;
  if (nheads <= 0 || qn <= 0) return;
  if (key_lo > ctx + q0 + qn - 1) return;                     // no key of this chunk is visible to any query of the unit
  const int pairs = nheads * qn;
  const int ntiles = (pairs + 127) >> 7;
  const uint32_t base = tc_kv_base<D>(w);
  const uint32_t tm = ws->tmem_base;
  const int q = (int)(threadIdx.x >> 5) & 3;                 // TMEM lane quarter of this warp
  const int lane_tm = q * 32 + w.lane();                       // TMEM lane = pair index inside a tile
  const uint32_t tq = tm + ((uint32_t)(q * 32) << 16);
  const float scale = rsqrtf((float)D);
  const int* bt = c.block_tables + (size_t)grp.seq * c.max_pages_per_seq;
  // K / V chunk -> shared memory (16-byte asynchronous copies; rows past the last key are zero-filled: 0 * P = 0)
  if (!prefetched) workers_sync();                           // previous readers (MMAs of the previous unit) are done
  {
    const __nv_bfloat16* cbase[2] = {nullptr, nullptr};
    const bool need_committed = key_lo < ctx && !prefetched;
    if (need_committed) {
      const int page = ld_weak(bt + (key_lo / kPage));
      cbase[0] = c.kv_cache + kv_cache_offset(md, c.num_pages, layer, 0, page, g_kvh, key_lo % kPage);
      cbase[1] = c.kv_cache + kv_cache_offset(md, c.num_pages, layer, 1, page, g_kvh, key_lo % kPage);
    }
    const __nv_bfloat16* sbase[2] = {c.kv_scratch + kv_scratch_offset(md, c.T_cap, layer, 0, row0, g_kvh),
                                     c.kv_scratch + kv_scratch_offset(md, c.T_cap, layer, 1, row0, g_kvh)};
    const size_t srow = (size_t)md.n_kv * D;
#pragma unroll 4
    for (int idx = w.wtid(); idx < kAttnSub * VPR * 2; idx += kWorkerThreads) {
      const int kv = idx / (kAttnSub * VPR);
      const int rem = idx - kv * (kAttnSub * VPR);
      const int jr = rem / VPR, v = rem - jr * VPR;
      const int pos = key_lo + jr;
      const __nv_bfloat16* src = sbase[kv];
      uint32_t nbytes = 0;
      if (pos < ctx) {
        if (prefetched) continue;                            // committed rows were fetched during the QKV stage
        src = cbase[kv] + (size_t)jr * D + v * 8; nbytes = 16;
      } else if (pos < nkeys) { src = sbase[kv] + (size_t)(pos - ctx) * srow + v * 8; nbytes = 16; }
      cp_async_16_saddr(tc_kv_addr<D>(base, kv, jr, v), src, nbytes);
    }
    cp_async_commit();
  }

  for (int tile = 0; tile < ntiles; ++tile) {
    const int p = tile * 128 + lane_tm;                      // this thread's pair
    const bool pvalid = p < pairs;
    const int hl = pvalid ? p / qn : 0;
    const int i = q0 + (pvalid ? p - hl * qn : 0);           // query row inside the group
    const int head = g_kvh * md.n_rep + j_lo + hl;
    // queries: registers -> tensor memory (invalid pairs: zeros)
    {
      const uint4* qsrc = reinterpret_cast<const uint4*>(c.qbuf + (size_t)(row0 + i) * md.q_dim + (size_t)head * D);
#pragma unroll
      for (int c16 = 0; c16 < D / 32; ++c16) {               // 16 columns = 32 dims = 4 x 16 bytes
        uint32_t v[16];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          uint4 u = make_uint4(0u, 0u, 0u, 0u);
          if (pvalid) u = ld_weak(qsrc + c16 * 4 + k4);
          v[4 * k4] = u.x; v[4 * k4 + 1] = u.y; v[4 * k4 + 2] = u.z; v[4 * k4 + 3] = u.w;
        }
        tmem_st16(tq + kTcQ + c16 * 16, v);
      }
      tmem_st_wait();
    }
    if (tile == 0) cp_async_wait_all();
    fence_proxy_async_smem();                                // K / V written by this thread's copies -> async proxy (MMA)
    tc_fence_before();
    workers_sync();
    if (w.wtid() == 0) {                                       // S = Q K^T : M = 128 pairs, N = 64 keys, K = D dims
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, kAttnSub);
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
        if (!(c.debug_skip & 65536)) umma_bf16_ts(tm + kTcS, tm + kTcQ + kk * 8, umma_desc_sw128(base + (uint32_t)(kk >> 2) * 8192u + (uint32_t)(kk & 3) * 32u),
                     idesc, kk > 0 ? 1u : 0u);
      umma_commit(ws->attn_bar);
    }
    attn_bar_wait(c, ws->attn_bar, phase, 1001);
    phase ^= 1u;
    tc_fence_after();
    // softmax of this pair's row: 64 scores in registers
    float m_run = -INFINITY, l_run = 0.f;
    {
      uint32_t sv[64];
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) tmem_ld16(tq + kTcS + c16 * 16, *reinterpret_cast<uint32_t(*)[16]>(&sv[c16 * 16]));
      tmem_ld_wait();
      const int qpos = ctx + i;
      const int kvis = min(qpos - key_lo + 1, nkeys - key_lo);      // visible keys of the chunk: columns [0, kvis)
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const float sc = (k < kvis) ? __uint_as_float(sv[k]) * scale : -INFINITY;
        sv[k] = __float_as_uint(sc);
        m_run = fmaxf(m_run, sc);
      }
      const float mu = (m_run == -INFINITY) ? 0.f : m_run;
      uint32_t pk[32];
#pragma unroll
      for (int k = 0; k < 64; k += 2) {
        const float e0 = __expf(__uint_as_float(sv[k]) - mu), e1 = __expf(__uint_as_float(sv[k + 1]) - mu);
        l_run += e0 + e1;
        pk[k >> 1] = pack_bf16x2(e0, e1);
      }
      tmem_st16(tq + kTcP, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
      tmem_st16(tq + kTcP + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
      tmem_st_wait();
    }
    tc_fence_before();
    workers_sync();
    if (w.wtid() == 0) {                                       // O = P V : M = 128 pairs, N = D dims, K = 64 keys
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16_bmn(128, D);
      constexpr uint32_t kTile = 64u * D * 2u;
#pragma unroll
      for (int kk = 0; kk < kAttnSub / 16; ++kk)
        if (!(c.debug_skip & 131072)) umma_bf16_ts(tm + kTcO, tm + kTcP + kk * 8, umma_desc_sw128_mn(base + kTile + (uint32_t)kk * 2048u, 8192u, 1024u),
                     idesc, kk > 0 ? 1u : 0u);
      umma_commit(ws->attn_bar);
    }
    attn_bar_wait(c, ws->attn_bar, phase, 1002);
    phase ^= 1u;
    tc_fence_after();
    // unnormalised partial output + (max, sum) of this pair and chunk (tcgen05.ld is warp-collective: every lane loads,
    // only the lanes that hold a live pair store)
    {
      const bool live = pvalid && key_lo <= ctx + i;
      const size_t sh = live ? ((size_t)(grp.slot0 + i * grp.nchunks + chunk) * md.n_q + head) : 0;
      float* op = c.attn_o + sh * D;
#pragma unroll
      for (int c16 = 0; c16 < D / 16; ++c16) {
        uint32_t ov[16];
        tmem_ld16(tq + kTcO + c16 * 16, ov);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            *reinterpret_cast<float4*>(op + c16 * 16 + k4 * 4) =
                make_float4(__uint_as_float(ov[4 * k4]), __uint_as_float(ov[4 * k4 + 1]), __uint_as_float(ov[4 * k4 + 2]),
                            __uint_as_float(ov[4 * k4 + 3]));
        }
      }
      if (live) *reinterpret_cast<float2*>(c.attn_ml + sh * 2) = make_float2(m_run, l_run);
    }
    tc_fence_before();                                       // the next tile / unit overwrites Q, S, P, O
    if (tile + 1 < ntiles) workers_sync();
  }
}

// Issued by the workers at the START of the layer's QKV GEMM stage: the committed K/V rows of the first 64-key
// sub-tile of the unit this CTA will run in the attention stage do not depend on this layer's projections, so their
// HBM latency hides behind the QKV GEMM and the rope stage.  attention_unit(prefetched = true) then only adds the
// in-flight rows.  (The tile buffer is free: the previous attention stage and LM head ended grid barriers ago.)
__device__ __noinline__ void attention_prefetch(const DevCtx& c, const WorkerShared* ws, int layer, int G, int units, int hsplit) {
  const WorkerCtx w = load_wctx(ws);
  if (w.cta() >= units || !c.attn_prefetch) return;
  const AttnUnitId id = decode_attn_unit(c, w, w.cta(), G, hsplit);
  const ModelDims& md = c.md;
  const int D = md.d, LDS = D + 8, VPR = D >> 3;
  const int ctx = id.grp.ctx;
  const int k0 = id.chunk * c.attn_chunk;
  if (k0 >= ctx) return;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(w.attn_smem());
  __nv_bfloat16* Vs = Ks + kAttnSub * LDS;
  const int page = ld_weak(c.block_tables + (size_t)id.grp.seq * c.max_pages_per_seq + (k0 / kPage));
  const __nv_bfloat16* cbase[2] = {c.kv_cache + kv_cache_offset(md, c.num_pages, layer, 0, page, id.kvh, k0 % kPage),
                                   c.kv_cache + kv_cache_offset(md, c.num_pages, layer, 1, page, id.kvh, k0 % kPage)};
  const int nrow = min(kAttnSub, ctx - k0);
  const bool tc = c.attn_tc != 0 && c.attn_chunk == kAttnSub;
  const uint32_t tbase = (smem_u32(w.attn_smem()) + 1023u) & ~1023u;
  for (int idx = w.wtid(); idx < kAttnSub * VPR * 2; idx += kWorkerThreads) {
    const int kv = idx / (kAttnSub * VPR);
    const int rem = idx - kv * (kAttnSub * VPR);
    const int jr = rem / VPR, v = rem - jr * VPR;
    if (jr >= nrow) continue;
    if (tc) {
      const uint32_t dst = (D == 128) ? tc_kv_addr<128>(tbase, kv, jr, v) : tc_kv_addr<64>(tbase, kv, jr, v);
      cp_async_16_saddr(dst, cbase[kv] + (size_t)jr * D + v * 8, 16);
    } else {
      cp_async_16((kv ? Vs : Ks) + jr * LDS + v * 8, cbase[kv] + (size_t)jr * D + v * 8, 16);
    }
  }
  cp_async_commit();
}

__device__ __noinline__ void stage_attention_tc(const DevCtx& c, WorkerShared* ws, int layer, int G, int units, int hsplit) {
  const WorkerCtx w = load_wctx(ws);
  uint32_t phase = ws->attn_phase;
  for (int u = w.cta(); u < units; u += w.nctas()) {
    prof_mark(c, w, 398);
    const AttnUnitId id = decode_attn_unit(c, w, u, G, hsplit);
    const bool pf = c.attn_prefetch && u == w.cta();
    prof_mark(c, w, 399);
    if (c.md.d == 128) attention_unit_tc<128>(c, w, ws, phase, layer, id.grp, id.kvh, id.chunk, id.qs, id.hs, pf, hsplit);
    else attention_unit_tc<64>(c, w, ws, phase, layer, id.grp, id.kvh, id.chunk, id.qs, id.hs, pf, hsplit);
    prof_mark(c, w, 403);
  }
  workers_sync();
  if (w.wtid() == 0) ws->attn_phase = phase;
}

__device__ __noinline__ void stage_attention(const DevCtx& c, WorkerShared* ws, int layer, int G, int units, int hsplit) {
  if (c.attn_tc != 0 && c.attn_chunk == kAttnSub) { stage_attention_tc(c, ws, layer, G, units, hsplit); return; }
  const WorkerCtx w = load_wctx(ws);
  for (int u = w.cta(); u < units; u += w.nctas()) {
    prof_mark(c, w, 398);
    const AttnUnitId id = decode_attn_unit(c, w, u, G, hsplit);
    const bool pf = c.attn_prefetch && u == w.cta();
    prof_mark(c, w, 399);
    if (c.md.d == 128) attention_unit<128>(c, w, layer, id.grp, id.kvh, id.chunk, id.qs, id.hs, pf, hsplit);
    else attention_unit<64>(c, w, layer, id.grp, id.kvh, id.chunk, id.qs, id.hs, pf, hsplit);
  }
}

// merge the per-chunk partials of every (token, head) in chunk order; write bf16 into the O-projection operand.
// One warp per (token, head), two items in flight per warp; lane ch holds the statistics of chunk ch, the chunk
// vectors are fetched 8 at a time per item.  (Rows with more than 32 visible chunks take the generic two-pass path.)
__device__ __forceinline__ void combine_store(const DevCtx& c, int t, int head, int lane, int vpl, int D,
                                              const float (&acc)[4], float l) {
  const float inv = 1.0f / l;
  const int f = head * D + lane * vpl;                      // feature index = k index of the O projection
  uint8_t* dst = reinterpret_cast<uint8_t*>(c.xa) + ((size_t)(f >> 6) * c.xrows_cap) * 128 + sw128_offset(t, f & 63);
  if (vpl == 4) {
    uint2 pk;
    pk.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
    pk.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<uint2*>(dst) = pk;
  } else {
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(acc[0] * inv, acc[1] * inv);
  }
}

__device__ __noinline__ void stage_attn_combine(const DevCtx& c, const WorkerShared* ws, int T) {
  const WorkerCtx w = load_wctx(ws);
  const ModelDims& md = c.md;
  const int D = md.d;
  const int vpl = D >> 5;                                  // floats per lane: 4 (d=128) or 2 (d=64)
  const int items = T * md.n_q;
  const size_t cstride = md.n_q;                            // slots of consecutive chunks are adjacent
  const int stride = w.nctas() * 4;
  prof_mark(c, w, 300);
  for (int item0 = w.cta() * 4 + w.wwarp(); item0 < items; item0 += 2 * stride) {
    int tt[2], hh[2], cmax[2];
    size_t base[2];
    float wl[2], ll[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int item = item0 + q * stride;
      tt[q] = -1; cmax[q] = -1; base[q] = 0; hh[q] = 0;
      if (item < items) {
        tt[q] = item / md.n_q; hh[q] = item - tt[q] * md.n_q;
        const int2 ra = w.plan() ? w.plan()->row_attn[tt[q]] : ld_weak(&c.row_attn[tt[q]]);
        cmax[q] = ra.y;
        base[q] = (size_t)ra.x * md.n_q + hh[q];
      }
    }
    prof_mark(c, w, 301);
#pragma unroll
    for (int q = 0; q < 2; ++q) {                            // statistics of the first 32 chunks of both items
      float2 ml = make_float2(-INFINITY, 0.f);
      if (tt[q] >= 0 && w.lane() <= cmax[q]) ml = ld_weak(reinterpret_cast<const float2*>(c.attn_ml + (base[q] + w.lane() * cstride) * 2));
      float mx = ml.x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      wl[q] = (ml.x == -INFINITY) ? 0.f : __expf(ml.x - mx);
      ll[q] = wl[q] * ml.y;
    }
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float l[2] = {0.f, 0.f};
    const int cm0 = (tt[0] >= 0 && cmax[0] < 32) ? cmax[0] : -1;     // -1: nothing to do on the fast path
    const int cm1 = (tt[1] >= 0 && cmax[1] < 32) ? cmax[1] : -1;
    const int nmax = max(cm0, cm1) + 1;
    const float* ob[2] = {c.attn_o + base[0] * D + w.lane() * vpl, c.attn_o + base[1] * D + w.lane() * vpl};
    const size_t ostride = cstride * D;
    for (int cc = 0; cc < nmax; cc += 8) {
      // Loads are unconditional (chunk index clamped, weight zeroed) so the batch stays in registers and in flight.
      float4 v4[2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int cm = q ? cm1 : cm0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int ch = max(0, min(cc + u, cm));
          const float* op = ob[q] + (size_t)ch * ostride;
          if (vpl == 4) v4[q][u] = ld_weak(reinterpret_cast<const float4*>(op));
          else { const float2 t2 = ld_weak(reinterpret_cast<const float2*>(op)); v4[q][u] = make_float4(t2.x, t2.y, 0.f, 0.f); }
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int cm = q ? cm1 : cm0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {                        // chunk order: deterministic
          float wgt = __shfl_sync(0xffffffffu, wl[q], (cc + u) & 31);
          float lc = __shfl_sync(0xffffffffu, ll[q], (cc + u) & 31);
          if (cc + u > cm) { wgt = 0.f; lc = 0.f; }
          l[q] += lc;
          acc[q][0] += wgt * v4[q][u].x; acc[q][1] += wgt * v4[q][u].y;
          acc[q][2] += wgt * v4[q][u].z; acc[q][3] += wgt * v4[q][u].w;
        }
      }
    }
    prof_mark(c, w, 302);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (tt[q] < 0) continue;
      if (cmax[q] >= 32) {                                   // long contexts: generic two-pass merge over all chunks
        float mg = -INFINITY;
        for (int c0 = 0; c0 <= cmax[q]; c0 += 32) {
          const int ch = c0 + w.lane();
          const float mc = (ch <= cmax[q]) ? ld_weak(c.attn_ml + (base[q] + ch * cstride) * 2) : -INFINITY;
          mg = fmaxf(mg, mc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
        float lg = 0.f;
        float ag[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = 0; c0 <= cmax[q]; c0 += 32) {
          const int chl = c0 + w.lane();
          float2 ml = make_float2(-INFINITY, 0.f);
          if (chl <= cmax[q]) ml = ld_weak(reinterpret_cast<const float2*>(c.attn_ml + (base[q] + chl * cstride) * 2));
          const float wq = (ml.x == -INFINITY) ? 0.f : __expf(ml.x - mg);
          const float lq = wq * ml.y;
          const int nch = min(32, cmax[q] + 1 - c0);
          for (int u = 0; u < nch; ++u) {
            const float wgt = __shfl_sync(0xffffffffu, wq, u);
            lg += __shfl_sync(0xffffffffu, lq, u);
            const float* op = c.attn_o + (base[q] + (size_t)(c0 + u) * cstride) * D + w.lane() * vpl;
            if (vpl == 4) {
              const float4 v = ld_weak(reinterpret_cast<const float4*>(op));
              ag[0] += wgt * v.x; ag[1] += wgt * v.y; ag[2] += wgt * v.z; ag[3] += wgt * v.w;
            } else {
              const float2 v = ld_weak(reinterpret_cast<const float2*>(op));
              ag[0] += wgt * v.x; ag[1] += wgt * v.y;
            }
          }
        }
        combine_store(c, tt[q], hh[q], w.lane(), vpl, D, ag, lg);
      } else {
        combine_store(c, tt[q], hh[q], w.lane(), vpl, D, acc[q], l[q]);
      }
    }
  }
}

// accepted in-flight K/V rows -> their pages (all layers)
__device__ __noinline__ void stage_commit(const DevCtx& c, const WorkerShared* ws, int n_commit) {
  const WorkerCtx w = load_wctx(ws);
  const ModelDims& md = c.md;
  const int per = md.layers * 2;
  const int items = n_commit * per;
  const int vpr = md.d >> 3;
  for (int item = w.cta() * 4 + w.wwarp(); item < items; item += w.nctas() * 4) {
    const int e = item / per, lk = item - e * per;
    const int layer = lk >> 1, kv = lk & 1;
    const int src_row = ld_weak(&c.commits[e].src_row);
    const int seq = ld_weak(&c.commits[e].seq);
    const int pos = ld_weak(&c.commits[e].pos);
    const int page = ld_weak(c.block_tables + (size_t)seq * c.max_pages_per_seq + pos / kPage);
    for (int kvh = 0; kvh < md.n_kv; ++kvh) {
      const uint4* src = reinterpret_cast<const uint4*>(c.kv_scratch + kv_scratch_offset(md, c.T_cap, layer, kv, src_row, kvh));
      uint4* dst = reinterpret_cast<uint4*>(c.kv_cache + kv_cache_offset(md, c.num_pages, layer, kv, page, kvh, pos % kPage));
      for (int v = w.lane(); v < vpr; v += 32) dst[v] = ld_weak(src + v);
    }
  }
}

}  // namespace jf
