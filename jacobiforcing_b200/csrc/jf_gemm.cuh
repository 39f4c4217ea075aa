// jf_gemm.cuh — the weight-streaming "skinny" GEMM at the heart of one Jacobi iteration.
//
//   Y[m, n] = sum_k X[m, k] * W[n, k]          m = in-flight draft rows (16..256 per MMA), n = output features
//
// computed swap-AB on the 5th-gen tensor cores: the WEIGHT tile is the UMMA A operand (M=128 weight rows), the
// draft rows are the B operand (N = rows padded to 16), so D^T[128 features x rows] lives in TMEM with one
// feature per TMEM lane.  Replaces the reference's cuBLAS calls behind F.linear
// (inference_engine/layers/linear.py:72-73, 149-153; HF Qwen2 q/k/v/o/gate/up/down/lm_head Linear modules
// called from modeling/cllm2_..._multiblock_lookahead_unified.py:452-463).
//
// Warp roles inside the persistent kernel (224 threads):
//   warp 0  W producer   streams 16 KiB pre-swizzled weight tiles with one bulk copy each; depends on NOTHING but
//                        free ring slots, so it runs ahead of the grid barriers and keeps HBM busy while the other
//                        warps are still finishing the previous stage
//   warp 1  MMA issuer   owns TMEM, one elected lane issues tcgen05.mma
//   warps 2-5  workers   GEMM epilogues (TMEM -> registers -> global) and every SIMT stage of the forward
//   warp 6  X producer   loads the activation k-chunks; waits for the grid barrier that publishes them
// Three pipelines: smem ring (full/empty mbarriers, two arrivals per slot: W and X), TMEM accumulator double buffer
// (acc_full/acc_empty) and the static unit loop (unit = (n_tile, k_split)), walked identically by all roles.
#pragma once
#include "jf_ptx.cuh"
#include "jf_sched.h"
#include "jf_policy.h"

namespace jf {

constexpr int kTileN       = 128;          // weight rows per tile (UMMA M)
constexpr int kTileK       = 64;           // bf16 per 128-byte swizzle row
constexpr int kWTileBytes  = kTileN * kTileK * 2;   // 16 KiB, contiguous in HBM
constexpr int kMaxStages   = 12;
constexpr int kMaxRowsMMA  = 256;          // UMMA N limit
constexpr int kArgmaxRows  = 256;          // rows one EPI_ARGMAX task may carry (running best lives in smem)

constexpr int kWarpW = 0, kWarpMMA = 1, kWarpWorker0 = 2, kWarpX = 6;
constexpr int kNumThreads    = 224;
constexpr int kWorkerThreads = 128;

enum GemmEpilogue : int {
  EPI_PARTIAL_F32 = 0,   // out[slot][m][n] fp32 partial sums (slot = k_split); the consumer stage reduces in slot order
  EPI_SWIGLU_BF16 = 1,   // two accumulators (gate tile, up tile) -> silu(g)*u -> bf16 in packed-X layout
  EPI_PARTIAL_TAGGED = 3, // the same partial sums as 64-bit {stage tag, fp32 bits} words: the fused tails poll the DATA,
                          // so no flag / fence / barrier separates the k-splits' stores from their consumers
  EPI_SWIGLU_TAGGED = 4, // split-K gate/up: both accumulators as tagged fp32 partial sums; the fused tail adds the splits
                         // and applies SwiGLU (used when the gate/up tiles alone would leave SMs idle: tensor parallelism)
  EPI_ARGMAX      = 2,   // LM head: bf16-rounded logits, running (max, lowest index) per row -> atomicMax of packed keys
  EPI_SAMPLE      = 5,   // LM head of the non-greedy path: the argmax keys plus, per row, everything the rejection-sampling
                         // verify needs from softmax(logits / T) - its normaliser (running max, sum of exponentials), the
                         // scaled logit of the drafted token and a Gumbel-max sample that excludes it - in the same pass
};
constexpr int kNgRows = 128;   // rows one EPI_SAMPLE task may carry (four running statistics per row live in smem)

struct GemmTask {
  const __nv_bfloat16* W;    // packed weight tiles [n_tiles (x2 for SWIGLU: gate tiles then up tiles)][k_chunks][128x64 sw128]
  const __nv_bfloat16* X;    // packed activations  [k_chunks][xrows_cap x 64 sw128]
  void* out;                 // PARTIAL: float*; SWIGLU: packed bf16 X buffer of the next GEMM; ARGMAX: u64 keys [rows]
  int n_tiles;               // N / 128 (SWIGLU: gate tiles; up tiles follow at +n_tiles)
  int k_chunks;              // K / 64
  int splits;                // split-K factor (1 for SWIGLU / ARGMAX)
  int M;                     // live rows of this task
  int x_row0;                // first row inside X / first output row (multiple of 8)
  int xrows_cap;             // row capacity of the packed X buffer
  int N;                     // real feature count (PARTIAL) / vocab size (ARGMAX) / intermediate size (SWIGLU)
  int out_row_stride;        // PARTIAL: floats between consecutive rows of out
  int out_split_stride;      // PARTIAL: floats between the k-split slices of one row
  int epilogue;
  int out_rows_cap;          // SWIGLU: xrows_cap of the destination packed buffer
  int index_offset;          // ARGMAX: added to the local feature index (vocabulary slice of this rank)
  unsigned int tag;          // PARTIAL_TAGGED: stage tag stored with every partial sum
  int bulk_rows;             // PARTIAL_TAGGED: 1 = the tile's rows leave through shared memory, one 1 KiB bulk store per row
  // ---- SAMPLE: everything else the sampling epilogue needs (temperature, generator key, row tables, accumulators) is read
  // from the kernel context, so the task stays small (it is rebuilt by every role at every stage)
  const DevCtx* sample_ctx;
  int fwd;                   // forward index that enters the random counters
  int fwd_parity;            // which half of the double-buffered accumulators this forward fills
  int rows_total;            // T of the forward
};

struct GemmSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad_;
  uint64_t attn_bar;          // completion of the attention stage's MMAs (tcgen05 attention)
  long long t_acc_full;       // developer aid: clock of the first accumulator hand-over of the last GEMM stage
  unsigned long long* best;   // [4][kArgmaxRows] EPI_ARGMAX running best key per epilogue warp and row (aliases
                              // the attention tile buffer: the LM head never overlaps an attention stage)
  int tag_stage_rows;         // rows of a [rows][128] tile of tagged 64-bit words that fit the same buffer (0 = disabled)
  int stage_rows;             // rows of a [2][rows][128 B] staging tile that fit the same buffer (EPI_SWIGLU_BF16 bulk
                              // stores; 0 = disabled): the MLP never overlaps an attention stage either
};

struct PipeState {      // smem ring position.  Phase parity is tracked PER SLOT (bit i = parity of slot i) so the ring
  uint32_t stage = 0;   // geometry (slot size / slot count) can change between forwards once the ring has drained.
  uint32_t phase_bits = 0;
  __device__ __forceinline__ uint32_t phase() const { return (phase_bits >> stage) & 1u; }
  __device__ __forceinline__ void advance(uint32_t nstages) {
    phase_bits ^= (1u << stage);
    if (++stage == nstages) stage = 0;
  }
};
struct AccState {       // TMEM accumulator ring position
  uint32_t buf = 0, phase = 0;
  __device__ __forceinline__ void advance() {
    if (++buf == 2) { buf = 0; phase ^= 1; }
  }
};

__host__ __device__ __forceinline__ int round_up16(int x) { return (x + 15) & ~15; }

// Packed argmax key: larger value wins; on equal values the LOWER index wins (torch.argmax on CPU returns the first
// maximum; reference: modeling/..._unified.py:476).  Values are bf16-rounded logits, never NaN for finite weights.
__host__ __device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(v);
#else
  uint32_t u; memcpy(&u, &v, 4);
#endif
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}
__host__ __device__ __forceinline__ int argmax_key_index(unsigned long long key) {
  return (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
}

// One-time setup by the whole CTA.
__device__ __forceinline__ void gemm_setup(GemmSmem* gs) {
  const int nstages = kMaxStages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&gs->full[i], 2); mbar_init(&gs->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&gs->acc_full[i], 1); mbar_init(&gs->acc_empty[i], kWorkerThreads); }
    mbar_init(&gs->attn_bar, 1);
    fence_mbar_init();
  }
  if (warp == kWarpMMA) tmem_alloc(&gs->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
__device__ __forceinline__ void gemm_teardown(GemmSmem* gs) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == kWarpMMA) tmem_dealloc(gs->tmem_base, 512);
}

// Bytes of one ring slot: one weight tile + the X k-chunk for up to mpad_max rows.
__host__ __device__ __forceinline__ int gemm_slot_bytes(int mpad_max) {
  return kWTileBytes + ((mpad_max * 128 + 1023) & ~1023);
}

struct UnitRange { int n_tile, split, k0, k1; };
__device__ __forceinline__ UnitRange unit_range(const GemmTask& t, int u) {
  UnitRange r;
  r.n_tile = u / t.splits;
  r.split = u - r.n_tile * t.splits;
  r.k0 = split_k0(r.split, t.splits, t.k_chunks);
  r.k1 = split_k1(r.split, t.splits, t.k_chunks);
  return r;
}
// Rows one accumulator use covers.  SwiGLU needs two accumulators (gate, up): up to 128 rows they share one 256-column
// TMEM buffer (double-buffered across units); 129..256 rows take the whole TMEM (gate in columns [0,256), up in
// [256,512)) so the weight tiles are still streamed once - at the price of no overlap between units.
__device__ __forceinline__ bool task_dual(const GemmTask& t) { return t.epilogue == EPI_SWIGLU_BF16 || t.epilogue == EPI_SWIGLU_TAGGED; }
__device__ __forceinline__ bool task_wide_dual(const GemmTask& t) { return task_dual(t) && t.M > 128; }
__device__ __forceinline__ int task_rows_mma(const GemmTask& t) {
  return task_dual(t) ? (t.M > 128 ? 256 : 128) : kMaxRowsMMA;
}

// ---- warp 0, lane 0 ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_produce_w(const GemmTask& t, GemmSmem* gs, uint8_t* tiles, int nstages,
                                               int slot_bytes, int cta, int nctas, PipeState& pipe, bool hint,
                                               uint64_t policy) {
  const bool dual = task_dual(t);
  const int units = t.n_tiles * t.splits;
  const int rows_mma = task_rows_mma(t);
  const int m_tiles = (t.M + rows_mma - 1) / rows_mma;
  auto copy_w = [&](void* dst, const void* src, uint64_t* bar) {
    if (hint) bulk_g2s_hint(dst, src, kWTileBytes, bar, policy); else bulk_g2s(dst, src, kWTileBytes, bar);
  };
  for (int u = cta; u < units; u += nctas) {
    const UnitRange r = unit_range(t, u);
    for (int mt = 0; mt < m_tiles; ++mt) {
      for (int k = r.k0; k < r.k1; ++k) {
        mbar_wait(&gs->empty[pipe.stage], pipe.phase() ^ 1);
        mbar_expect_tx(&gs->full[pipe.stage], kWTileBytes);
        copy_w(tiles + (size_t)pipe.stage * slot_bytes,
               t.W + ((size_t)r.n_tile * t.k_chunks + k) * (kTileN * kTileK), &gs->full[pipe.stage]);
        pipe.advance(nstages);
        if (dual) {
          mbar_wait(&gs->empty[pipe.stage], pipe.phase() ^ 1);
          mbar_expect_tx(&gs->full[pipe.stage], kWTileBytes);
          copy_w(tiles + (size_t)pipe.stage * slot_bytes,
                 t.W + ((size_t)(r.n_tile + t.n_tiles) * t.k_chunks + k) * (kTileN * kTileK), &gs->full[pipe.stage]);
          pipe.advance(nstages);
        }
      }
    }
  }
}

// ---- warp 6, lane 0 (after it has seen the grid barrier that publishes X) ----------------------------------------
__device__ __forceinline__ void gemm_produce_x(const GemmTask& t, GemmSmem* gs, uint8_t* tiles, int nstages,
                                               int slot_bytes, int cta, int nctas, PipeState& pipe) {
  const bool dual = task_dual(t);
  const int units = t.n_tiles * t.splits;
  const int rows_mma = task_rows_mma(t);
  const int m_tiles = (t.M + rows_mma - 1) / rows_mma;
  for (int u = cta; u < units; u += nctas) {
    const UnitRange r = unit_range(t, u);
    for (int mt = 0; mt < m_tiles; ++mt) {
      const int m0 = mt * rows_mma;
      const uint32_t xbytes = (uint32_t)round_up16(min(t.M - m0, rows_mma)) * 128u;
      for (int k = r.k0; k < r.k1; ++k) {
        mbar_wait(&gs->empty[pipe.stage], pipe.phase() ^ 1);
        mbar_expect_tx(&gs->full[pipe.stage], xbytes);
        bulk_g2s(tiles + (size_t)pipe.stage * slot_bytes + kWTileBytes,
                 t.X + ((size_t)k * t.xrows_cap + t.x_row0 + m0) * kTileK, xbytes, &gs->full[pipe.stage]);
        pipe.advance(nstages);
        if (dual) {   // the "up" slot reuses the gate slot's X chunk: plain arrival keeps the count at two
          mbar_wait(&gs->empty[pipe.stage], pipe.phase() ^ 1);
          mbar_arrive(&gs->full[pipe.stage]);
          pipe.advance(nstages);
        }
      }
    }
  }
}

// ---- warp 1, lane 0 -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_issue_mma(const GemmTask& t, GemmSmem* gs, uint8_t* tiles, int nstages,
                                               int slot_bytes, int cta, int nctas, PipeState& pipe, AccState& acc) {
  const bool dual = task_dual(t);
  const bool wide = task_wide_dual(t);
  const int units = t.n_tiles * t.splits;
  const int rows_mma = task_rows_mma(t);
  const int m_tiles = (t.M + rows_mma - 1) / rows_mma;
  const uint32_t tmem = gs->tmem_base;
  for (int u = cta; u < units; u += nctas) {
    const UnitRange r = unit_range(t, u);
    for (int mt = 0; mt < m_tiles; ++mt) {
      const int m0 = mt * rows_mma;
      const int mpad = round_up16(min(t.M - m0, rows_mma));
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)mpad);
      mbar_wait(&gs->acc_empty[acc.buf], acc.phase ^ 1);
      AccState acc2 = acc;
      if (wide) { acc2.advance(); mbar_wait(&gs->acc_empty[acc2.buf], acc2.phase ^ 1); }   // both buffers must be free
      tc_fence_after();
      const uint32_t d0 = wide ? tmem : tmem + acc.buf * 256u;   // buffer b = columns [b*256, b*256+256); dual: gate | up
      const uint32_t up_off = wide ? 256u : 128u;
      for (int k = r.k0; k < r.k1; ++k) {
        mbar_wait(&gs->full[pipe.stage], pipe.phase());
        const uint32_t sa = smem_u32(tiles + (size_t)pipe.stage * slot_bytes);
        uint64_t* empty_a = &gs->empty[pipe.stage];
        pipe.advance(nstages);
        uint32_t sb = 0;
        uint64_t* empty_b = nullptr;
        if (dual) {
          mbar_wait(&gs->full[pipe.stage], pipe.phase());
          sb = smem_u32(tiles + (size_t)pipe.stage * slot_bytes);
          empty_b = &gs->empty[pipe.stage];
          pipe.advance(nstages);
        }
        tc_fence_after();
        const uint32_t xbase = sa + kWTileBytes;
        const uint32_t accum0 = (k > r.k0) ? 1u : 0u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {   // 4 x UMMA_K(16) = one 64-wide k-chunk; +32 B inside the swizzle atom
          const uint64_t db = umma_desc_sw128(xbase + kk * 32);
          umma_bf16(d0, umma_desc_sw128(sa + kk * 32), db, idesc, (accum0 | (kk > 0)) ? 1u : 0u);
          if (dual) umma_bf16(d0 + up_off, umma_desc_sw128(sb + kk * 32), db, idesc, (accum0 | (kk > 0)) ? 1u : 0u);
        }
        umma_commit(empty_a);               // frees the slot(s) when these MMAs retire
        if (dual) umma_commit(empty_b);
      }
      umma_commit(&gs->acc_full[acc.buf]);
      acc.advance();
      if (wide) { umma_commit(&gs->acc_full[acc.buf]); acc.advance(); }
    }
  }
}

// EPI_SWIGLU_BF16 through shared memory: the tile's silu(g) * u values are staged in the operand layout of the down
// projection ([k-chunk][row][128 B], 128-byte swizzle) and leave as ONE bulk copy per k-chunk.  Against 32 scattered 2-byte
// stores per thread this (a) needs no generic-proxy -> async-proxy fence and no drain of hundreds of small stores before the
// hand-over to the down projection (the bulk group's completion IS the drain), (b) is a 4x smaller, rolled epilogue body.
__device__ __forceinline__ bool swiglu_bulk_ok(const GemmTask& t, int stage_rows) {
  return t.epilogue == EPI_SWIGLU_BF16 && t.M <= 128 && round_up16(t.M) <= stage_rows && ((t.x_row0 & 7) == 0);
}

// ---- warps 2..5: TMEM -> registers -> global ----------------------------------------------------------------------
__device__ __forceinline__ void gemm_epilogue(const GemmTask& t, GemmSmem* gs, int cta, int nctas, AccState& acc) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q = warp & 3;                       // TMEM lane quarter this warp may access (warp id % 4)
  const int feat_in_tile = q * 32 + lane;
  const int units = t.n_tiles * t.splits;
  const bool wide = task_wide_dual(t);
  const int rows_mma = task_rows_mma(t);
  const int m_tiles = (t.M + rows_mma - 1) / rows_mma;
  if (t.epilogue == EPI_ARGMAX) {
    for (int m = lane; m < t.M; m += 32) gs->best[q * kArgmaxRows + m] = 0ull;
    __syncwarp();
  }
  // EPI_SAMPLE shared-memory state (aliases the attention tile buffer like `best`): per epilogue warp and row the running
  // argmax key, Gumbel key and (max, sum) pair; per row the drafted token the row verifies (-1: none)
  unsigned long long* ng_best = gs->best;                                     // [4][kNgRows]
  unsigned long long* ng_gbest = gs->best + 4 * kNgRows;                      // [4][kNgRows]
  float2* ng_msw = reinterpret_cast<float2*>(gs->best + 8 * kNgRows);         // [4][kNgRows]
  int* ng_prop_tok = reinterpret_cast<int*>(gs->best + 12 * kNgRows);         // [kNgRows]
  if (t.epilogue == EPI_SAMPLE) {
    for (int m = lane; m < t.M; m += 32) {
      ng_best[q * kNgRows + m] = 0ull; ng_gbest[q * kNgRows + m] = 0ull;
      ng_msw[q * kNgRows + m] = make_float2(-INFINITY, 0.f);
      if (q == 0) {
        const int r = t.x_row0 + m;
        ng_prop_tok[m] = (r + 1 < t.rows_total && t.sample_ctx->row_group[r + 1] == t.sample_ctx->row_group[r])
                             ? t.sample_ctx->tok[r + 1] : -1;
      }
    }
    named_bar_sync(1, kWorkerThreads);           // kBarWorkers: the four epilogue warps read warp 0's drafted-token table
  }
  for (int u = cta; u < units; u += nctas) {
    const UnitRange r = unit_range(t, u);
    const int n = r.n_tile * kTileN + feat_in_tile;
    for (int mt = 0; mt < m_tiles; ++mt) {
      const int m0 = mt * rows_mma;
      const int mlive = min(t.M - m0, rows_mma);
      const int mpad = round_up16(mlive);
      mbar_wait(&gs->acc_full[acc.buf], acc.phase);
      if (threadIdx.x == 64 && u == cta && mt == 0) gs->t_acc_full = clock64();
      AccState acc2 = acc;
      if (wide) { acc2.advance(); mbar_wait(&gs->acc_full[acc2.buf], acc2.phase); }
      tc_fence_after();
      const uint32_t taddr = gs->tmem_base + (wide ? 0u : acc.buf * 256u) + ((uint32_t)(q * 32) << 16);
      const uint32_t up_off = wide ? 256u : 128u;
      if (t.epilogue == EPI_PARTIAL_F32) {
        float* out = reinterpret_cast<float*>(t.out) + (size_t)r.split * t.out_split_stride;
        const bool nvalid = n < t.N;
        for (int c0 = 0; c0 < mpad; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = t.x_row0 + m0 + c0 + j;
            st_pred_f32(out + (size_t)m * t.out_row_stride + n, __uint_as_float(v[j]), nvalid && c0 + j < mlive);
          }
        }
      } else if (t.epilogue == EPI_PARTIAL_TAGGED && t.bulk_rows && mpad <= gs->tag_stage_rows && (t.N & 127) == 0) {
        // The tagged partial sums of the tile go through shared memory ([row][128 features] 64-bit words) and leave as one
        // 1 KiB bulk store per row, issued by 32 different threads: the epilogue's issue time shrinks from 32 scattered
        // 8-byte stores per thread to 8 shared-memory stores per 8 rows, and the words land together.  The consumers poll
        // the tags as before, so nobody waits for the copies here (the caller waits for their READ side before the buffer
        // is reused).
        unsigned long long* stg = gs->best;
        const unsigned long long tg = (unsigned long long)t.tag << 32;
        named_bar_sync(1, kWorkerThreads);                 // earlier users of the buffer (this CTA) are done with it
#pragma unroll 1
        for (int c0 = 0; c0 < mpad; c0 += 8) {
          uint32_t v[8];
          tmem_ld8(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) stg[(size_t)(c0 + j) * kTileN + feat_in_tile] = tg | (unsigned long long)v[j];
        }
        tc_fence_before();
        mbar_arrive(&gs->acc_empty[acc.buf]);
        fence_proxy_async_smem();
        named_bar_sync(1, kWorkerThreads);
        const int i = (int)threadIdx.x - 2 * 32;
        if (i < mlive) {
          unsigned long long* out = reinterpret_cast<unsigned long long*>(t.out) + (size_t)r.split * t.out_split_stride +
                                    (size_t)(t.x_row0 + m0 + i) * t.out_row_stride + (size_t)r.n_tile * kTileN;
          bulk_s2g(out, stg + (size_t)i * kTileN, kTileN * 8);
          bulk_commit();
        }
        if (u + nctas < units || mt + 1 < m_tiles) { bulk_wait_read_all(); }   // (the sync at the top of the next pass follows)
        acc.advance();
        continue;
      } else if (t.epilogue == EPI_PARTIAL_TAGGED) {
        unsigned long long* out = reinterpret_cast<unsigned long long*>(t.out) + (size_t)r.split * t.out_split_stride;
        const unsigned long long tg = (unsigned long long)t.tag << 32;
        const bool nvalid = n < t.N;
#pragma unroll 1
        for (int c0 = 0; c0 < mpad; c0 += 8) {             // a short rolled body: this code runs once per stage, cold
          uint32_t v[8];
          tmem_ld8(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int m = t.x_row0 + m0 + c0 + j;
            st_pred_u64(out + (size_t)m * t.out_row_stride + n, tg | (unsigned long long)v[j], nvalid && c0 + j < mlive);
          }
        }
      } else if (t.epilogue == EPI_SWIGLU_TAGGED) {
        // out[((m * splits + split) * 2 + {gate, up}) * N + n]
        unsigned long long* out = reinterpret_cast<unsigned long long*>(t.out) + (size_t)r.split * t.out_split_stride;
        const unsigned long long tg = (unsigned long long)t.tag << 32;
        const bool nvalid = n < t.N;
        for (int c0 = 0; c0 < mpad; c0 += 16) {
          uint32_t g[16], up[16];
          tmem_ld16(taddr + c0, g);
          tmem_ld16(taddr + up_off + c0, up);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = t.x_row0 + m0 + c0 + j;
            const bool live = nvalid && c0 + j < mlive;
            st_pred_u64(out + (size_t)m * t.out_row_stride + n, tg | (unsigned long long)g[j], live);
            st_pred_u64(out + (size_t)m * t.out_row_stride + t.N + n, tg | (unsigned long long)up[j], live);
          }
        }
      } else if (t.epilogue == EPI_SWIGLU_BF16 && swiglu_bulk_ok(t, gs->stage_rows)) {
        // gate in columns [0,mpad), up in columns [128,128+mpad); rounding points as in the direct path below
        uint8_t* stg = reinterpret_cast<uint8_t*>(gs->best);
        const int kc2 = feat_in_tile >> 6, cc = feat_in_tile & 63;
        const bool nvalid = n < t.N;
        uint8_t* srow = stg + (size_t)kc2 * mpad * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < mpad; c0 += 8) {
          uint32_t g[8], up[8];
          tmem_ld8(taddr + c0, g);
          tmem_ld8(taddr + up_off + c0, up);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gf = bf16_round(__uint_as_float(g[j]));
            const float uf = bf16_round(__uint_as_float(up[j]));
            const float s = bf16_round(gf / (1.f + __expf(-gf)));
            const __nv_bfloat16 ob = __float2bfloat16_rn(s * uf);
            if (nvalid && c0 + j < mlive)
              *reinterpret_cast<unsigned short*>(srow + sw128_offset((uint32_t)(c0 + j), (uint32_t)cc)) = *reinterpret_cast<const unsigned short*>(&ob);
          }
        }
        tc_fence_before();
        mbar_arrive(&gs->acc_empty[acc.buf]);              // the accumulator is free while the tile is still on its way out
        fence_proxy_async_smem();
        named_bar_sync(1, kWorkerThreads);
        if (threadIdx.x == 64) {
          uint8_t* dst = reinterpret_cast<uint8_t*>(t.out);
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const int nk = r.n_tile * kTileN + k2 * 64;
            if (nk < t.N)
              bulk_s2g(dst + ((size_t)(nk >> 6) * t.out_rows_cap + t.x_row0 + m0) * 128, stg + (size_t)k2 * mpad * 128, (uint32_t)mlive * 128u);
          }
          bulk_commit();
          bulk_wait_all();
        }
        if (u + nctas < units || mt + 1 < m_tiles) named_bar_sync(1, kWorkerThreads);    // the staging tile is reused
        acc.advance();
        continue;
      } else if (t.epilogue == EPI_SWIGLU_BF16) {
        // gate in columns [0,mpad), up in columns [128,128+mpad).  HF Qwen2MLP rounding points
        // (transformers Qwen2MLP.forward: down(act(gate(x)) * up(x)), every op in bf16): g,u -> bf16; silu(g) -> bf16;
        // product -> bf16.  Destination: packed-X layout of the down projection, feature n is its k index.
        uint8_t* dst = reinterpret_cast<uint8_t*>(t.out);
        const int kc = n >> 6, c = n & 63;
        const bool nvalid = n < t.N;
        for (int c0 = 0; c0 < mpad; c0 += 16) {
          uint32_t g[16], up[16];
          tmem_ld16(taddr + c0, g);
          tmem_ld16(taddr + up_off + c0, up);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = t.x_row0 + m0 + c0 + j;
            const float gf = bf16_round(__uint_as_float(g[j]));
            const float uf = bf16_round(__uint_as_float(up[j]));
            const float s = bf16_round(gf / (1.f + __expf(-gf)));
            const __nv_bfloat16 ob = __float2bfloat16_rn(s * uf);
            st_pred_u16(dst + ((size_t)kc * t.out_rows_cap) * 128 + sw128_offset(m, c), *reinterpret_cast<const unsigned short*>(&ob),
                        nvalid && c0 + j < mlive);
          }
        }
      } else if (t.epilogue == EPI_SAMPLE) {
        // Non-greedy verify (reference: engine/jacobi_decoding_nongreedy.py:300-355 over _build_target_probs :110-123):
        // p = softmax(bf16 logits / T).  One pass over the vocabulary tile gives, per row, the tile's argmax key, its
        // (max, sum exp) pair, the scaled logit of the drafted token if this thread owns it, and the tile's best
        // Gumbel-perturbed score among the OTHER tokens: argmax_v (logit_v / T + G_v), v != drafted, is a draw from p
        // conditioned on != drafted - the residual distribution of the delta proposal (:134-153) - without logits in HBM.
        const bool valid = n < t.N;
        const int vglob = n + t.index_offset;
        const DevCtx& sc = *t.sample_ctx;
        const float inv_temp = sc.ng_inv_temp;
        const unsigned int seed_lo = sc.ng_seed_lo, seed_hi = sc.ng_seed_hi;
        float* ng_prop = sc.ng_prop + (size_t)(t.fwd_parity) * sc.T_cap;
        for (int c0 = 0; c0 < mpad; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          unsigned long long mine_a = 0ull, mine_g = 0ull;
          float mine_m = -INFINITY, mine_s = 0.f;
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const int row4 = t.x_row0 + m0 + c0 + j4;                        // multiple of 4
            const Philox4 rnd = philox4x32_10((unsigned int)vglob, (unsigned int)(row4 >> 2), (unsigned int)t.fwd, RNG_GUMBEL,
                                              seed_lo, seed_hi);
            const unsigned int rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = j4 + jj;
              const int mloc = m0 + c0 + j;
              const float val = bf16_round(__uint_as_float(v[j]));
              const float scv = val * inv_temp;
              const int prop = (c0 + j < mlive) ? ng_prop_tok[mloc] : -1;
              if (valid && c0 + j < mlive && vglob == prop) ng_prop[t.x_row0 + mloc] = scv;
              unsigned long long ka = valid ? argmax_key(val, vglob) : 0ull;
              const float gum = -__logf(-__logf(u01_open(rw[jj])));
              unsigned long long kg = (valid && vglob != prop) ? argmax_key(scv + gum, vglob) : 0ull;
              float mx = valid ? scv : -INFINITY;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const unsigned long long oa = __shfl_xor_sync(0xffffffffu, ka, o);
                const unsigned long long og = __shfl_xor_sync(0xffffffffu, kg, o);
                ka = oa > ka ? oa : ka;
                kg = og > kg ? og : kg;
              }
              float ex = valid ? __expf(scv - mx) : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) ex += __shfl_xor_sync(0xffffffffu, ex, o);
              if (lane == j) { mine_a = ka; mine_g = kg; mine_m = mx; mine_s = ex; }
            }
          }
          const int m = m0 + c0 + lane;
          if (lane < 16 && c0 + lane < mlive) {
            if (mine_a > ng_best[q * kNgRows + m]) ng_best[q * kNgRows + m] = mine_a;
            if (mine_g > ng_gbest[q * kNgRows + m]) ng_gbest[q * kNgRows + m] = mine_g;
            const float2 old = ng_msw[q * kNgRows + m];
            const float nm = fmaxf(old.x, mine_m);
            if (nm > -INFINITY)
              ng_msw[q * kNgRows + m] = make_float2(nm, old.y * __expf(old.x - nm) + mine_s * __expf(mine_m - nm));
          }
          __syncwarp();
        }
      } else {  // EPI_ARGMAX
        // bf16-rounded logits (reference: lm_head output is bf16, then .float(); modeling/..._unified.py:463).
        // Each warp reduces its 32 vocab rows with shuffles and folds the winner into its running best in smem.
        const bool valid = n < t.N;
        for (int c0 = 0; c0 < mpad; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          unsigned long long mine = 0ull;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            unsigned long long key = valid ? argmax_key(bf16_round(__uint_as_float(v[j])), n + t.index_offset) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
              key = other > key ? other : key;
            }
            if (lane == j) mine = key;
          }
          const int m = m0 + c0 + lane;
          if (lane < 16 && c0 + lane < mlive && mine > gs->best[q * kArgmaxRows + m]) gs->best[q * kArgmaxRows + m] = mine;
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(&gs->acc_empty[acc.buf]);
      acc.advance();
      if (wide) { mbar_arrive(&gs->acc_empty[acc.buf]); acc.advance(); }
    }
  }
  if (t.epilogue == EPI_ARGMAX) {
    __syncwarp();
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(t.out);
    for (int m = lane; m < t.M; m += 32) {
      const unsigned long long k = gs->best[q * kArgmaxRows + m];
      if (k != 0ull) atomicMax(&keys[t.x_row0 + m], k);
    }
  }
  if (t.epilogue == EPI_SAMPLE) {
    __syncwarp();
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(t.out);
    for (int m = lane; m < t.M; m += 32) {
      const unsigned long long ka = ng_best[q * kNgRows + m], kg = ng_gbest[q * kNgRows + m];
      if (ka != 0ull) atomicMax(&keys[t.x_row0 + m], ka);
      if (kg != 0ull) atomicMax(&t.sample_ctx->ng_gumbel[(size_t)t.fwd_parity * t.sample_ctx->T_cap + t.x_row0 + m], kg);
    }
    named_bar_sync(1, kWorkerThreads);           // every epilogue warp's (max, sum) pairs are final
    if (cta < units)
      for (int m = (q * 32 + lane); m < t.M; m += kWorkerThreads) {
        float2 acc = ng_msw[m];
#pragma unroll
        for (int w2 = 1; w2 < 4; ++w2) {
          const float2 o = ng_msw[w2 * kNgRows + m];
          const float nm = fmaxf(acc.x, o.x);
          if (nm > -INFINITY) acc = make_float2(nm, acc.y * __expf(acc.x - nm) + o.y * __expf(o.x - nm));
        }
        t.sample_ctx->ng_ms[((size_t)t.fwd_parity * t.sample_ctx->ng_parts + cta) * t.sample_ctx->T_cap + t.x_row0 + m] = acc;
      }
    named_bar_sync(1, kWorkerThreads);           // before a following task re-initialises the tables
  }
}

}  // namespace jf
