// jf_gemm.cuh — the weight-streaming "skinny" GEMM at the heart of one Jacobi iteration.
//
//   Y[m, n] = sum_k X[m, k] * W[n, k]          m = draft rows (16..256 per MMA), n = output features
//
// computed swap-AB on the 5th-gen tensor cores: the WEIGHT tile is the UMMA A operand (M=128 weight rows), the
// draft rows are the B operand (N = rows padded to 16), so D^T[128 features x rows] lives in TMEM with one
// feature per TMEM lane.  Replaces the reference's cuBLAS calls behind F.linear
// (inference_engine/layers/linear.py:72-73, 149-153; HF Qwen2 q/k/v/o/gate/up/down/lm_head Linear modules
// called from modeling/cllm2_..._multiblock_lookahead_unified.py:452-463).
//
// Warp roles (192 threads): warp 0 = bulk-copy producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
// Three pipelines: smem ring (full/empty mbarriers), TMEM accumulator double buffer (acc_full/acc_empty), and the
// persistent unit loop (unit = (n_tile, k_split)), identical in all roles.
#pragma once
#include "jf_ptx.cuh"

namespace jf {

constexpr int kTileN      = 128;          // weight rows per tile (UMMA M)
constexpr int kTileK      = 64;           // bf16 per 128-byte swizzle row
constexpr int kWTileBytes = kTileN * kTileK * 2;   // 16 KiB, contiguous in HBM
constexpr int kMaxStages  = 16;
constexpr int kMaxRowsMMA = 256;          // UMMA N limit
constexpr int kGemmThreads = 192;

enum GemmEpilogue : int {
  EPI_PARTIAL_F32 = 0,   // out[slot][m][n] fp32 partial sums (split-K slot = k_split), deterministic consumer-side reduce
  EPI_SWIGLU_BF16 = 1,   // two accumulators (gate tile, up tile) -> silu(g)*u -> bf16 in packed-X layout
  EPI_ARGMAX      = 2,   // LM head: bf16-rounded logits, running (max, lowest index) per row
};

struct GemmTask {
  const __nv_bfloat16* W;    // packed weight tiles [n_tiles][k_chunks][128x64 swizzled]
  const __nv_bfloat16* X;    // packed activations  [k_chunks][xrows_cap x 64 swizzled]
  void* out;                 // epilogue-specific
  void* out2;                // epilogue-specific (argmax: indices)
  const float* bias;         // optional per-feature bias (EPI_PARTIAL: added by slot 0 only)
  int n_tiles;               // N / 128   (for SWIGLU: number of gate tiles; up tiles follow at +n_tiles)
  int k_chunks;              // K / 64
  int splits;                // split-K factor (1 for SWIGLU / ARGMAX)
  int M;                     // live rows
  int xrows_cap;             // row capacity of the packed X buffer
  int N;                     // out row stride in elements (EPI_PARTIAL) / vocab size (ARGMAX)
  int M_cap;                 // row capacity of out (slot stride = M_cap * N)
  int epilogue;
  int out_kchunks_rows;      // SWIGLU: xrows_cap of the destination packed buffer
};

struct GemmSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad_;
  float* best_val;           // [4 epilogue warps][argmax_cap] running best bf16-rounded logit (EPI_ARGMAX)
  int*   best_idx;           // [4][argmax_cap] its vocab index (lowest index wins ties)
  int    argmax_cap;
  int    pad2_;
};

struct PipeState {      // smem ring position; persists across tasks inside a persistent kernel
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t nstages) {
    if (++stage == nstages) { stage = 0; phase ^= 1; }
  }
};
struct AccState {       // TMEM accumulator ring position
  uint32_t buf = 0, phase = 0;
  __device__ __forceinline__ void advance() {
    if (++buf == 2) { buf = 0; phase ^= 1; }
  }
};

__device__ __forceinline__ int round_up16(int x) { return (x + 15) & ~15; }

// One-time setup by the whole CTA.  smem_tiles must be 1024-byte aligned.
__device__ __forceinline__ void gemm_setup(GemmSmem* gs, int nstages) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&gs->full[i], 1); mbar_init(&gs->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&gs->acc_full[i], 1); mbar_init(&gs->acc_empty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&gs->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
__device__ __forceinline__ void gemm_teardown(GemmSmem* gs) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) tmem_dealloc(gs->tmem_base, 512);
}

// Bytes of one ring slot for a given padded row count.
__host__ __device__ __forceinline__ int gemm_slot_bytes(int mpad, bool dual) {
  int x = (mpad * 128 + 1023) & ~1023;
  return (dual ? 2 : 1) * kWTileBytes + x;
}

// Runs one GEMM task cooperatively over `nctas` CTAs (this CTA has index `cta`).  All threads of the CTA call it.
// smem_tiles: ring storage (nstages * slot_bytes, 1024-aligned).
__device__ void gemm_run(const GemmTask& t, GemmSmem* gs, uint8_t* smem_tiles, int nstages, int slot_bytes,
                         int cta, int nctas, PipeState& pipe, AccState& acc) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool dual = (t.epilogue == EPI_SWIGLU_BF16);
  const int units = t.n_tiles * t.splits;
  const int rows_mma = dual ? 128 : kMaxRowsMMA;   // dual mode splits a 256-column accumulator buffer in two
  const int m_tiles = (t.M + rows_mma - 1) / rows_mma;
  const uint32_t wbytes = (dual ? 2u : 1u) * kWTileBytes;

  if (warp == 0) {
    // ===================== producer: one elected lane streams W tiles (+ the matching X k-chunk) =============
    if (lane == 0) {
      for (int u = cta; u < units; u += nctas) {
        const int n_tile = u / t.splits, split = u % t.splits;
        const int k0 = (int)(((long long)split * t.k_chunks) / t.splits);
        const int k1 = (int)(((long long)(split + 1) * t.k_chunks) / t.splits);
        for (int mt = 0; mt < m_tiles; ++mt) {
          const int m0 = mt * rows_mma;
          const int mpad = round_up16(min(t.M - m0, rows_mma));
          const uint32_t xbytes = (uint32_t)mpad * 128u;
          for (int k = k0; k < k1; ++k) {
            mbar_wait(&gs->empty[pipe.stage], pipe.phase ^ 1);
            uint8_t* slot = smem_tiles + (size_t)pipe.stage * slot_bytes;
            mbar_expect_tx(&gs->full[pipe.stage], wbytes + xbytes);
            bulk_g2s(slot, t.W + ((size_t)n_tile * t.k_chunks + k) * (kTileN * kTileK), kWTileBytes,
                     &gs->full[pipe.stage]);
            if (dual)
              bulk_g2s(slot + kWTileBytes,
                       t.W + ((size_t)(n_tile + t.n_tiles) * t.k_chunks + k) * (kTileN * kTileK), kWTileBytes,
                       &gs->full[pipe.stage]);
            bulk_g2s(slot + wbytes, t.X + ((size_t)k * t.xrows_cap + m0) * kTileK, xbytes, &gs->full[pipe.stage]);
            pipe.advance(nstages);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected lane =========================================================
    if (lane == 0) {
      const uint32_t tmem = gs->tmem_base;
      for (int u = cta; u < units; u += nctas) {
        const int split = u % t.splits;
        const int k0 = (int)(((long long)split * t.k_chunks) / t.splits);
        const int k1 = (int)(((long long)(split + 1) * t.k_chunks) / t.splits);
        for (int mt = 0; mt < m_tiles; ++mt) {
          const int m0 = mt * rows_mma;
          const int mpad = round_up16(min(t.M - m0, rows_mma));
          const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)mpad);
          mbar_wait(&gs->acc_empty[acc.buf], acc.phase ^ 1);
          tc_fence_after();
          // accumulator buffer b occupies columns [b*256, b*256+256); dual mode splits it 128 + 128
          const uint32_t d0 = tmem + acc.buf * 256u;
          for (int k = k0; k < k1; ++k) {
            mbar_wait(&gs->full[pipe.stage], pipe.phase);
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem_tiles + (size_t)pipe.stage * slot_bytes);
            const uint32_t xbase = sbase + wbytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {   // 4 x UMMA_K(16) = one 64-wide k-chunk; +32 B inside the swizzle atom
              const uint64_t db = umma_desc_sw128(xbase + kk * 32);
              umma_bf16(d0, umma_desc_sw128(sbase + kk * 32), db, idesc, (k > k0 || kk > 0) ? 1u : 0u);
              if (dual)
                umma_bf16(d0 + 128u, umma_desc_sw128(sbase + kWTileBytes + kk * 32), db, idesc,
                          (k > k0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&gs->empty[pipe.stage]);   // frees the smem slot when these MMAs retire
            pipe.advance(nstages);
          }
          umma_commit(&gs->acc_full[acc.buf]);
          acc.advance();
        }
      }
    }
  } else {
    // ===================== epilogue warps: TMEM -> registers -> global ==========================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int feat_in_tile = q * 32 + lane;
    if (t.epilogue == EPI_ARGMAX) {
      for (int m = lane; m < t.M; m += 32) {
        gs->best_val[q * gs->argmax_cap + m] = -INFINITY;
        gs->best_idx[q * gs->argmax_cap + m] = 0x7fffffff;
      }
      __syncwarp();
    }
    for (int u = cta; u < units; u += nctas) {
      const int n_tile = u / t.splits, split = u % t.splits;
      const int n = n_tile * kTileN + feat_in_tile;
      for (int mt = 0; mt < m_tiles; ++mt) {
        const int m0 = mt * rows_mma;
        const int mlive = min(t.M - m0, rows_mma);
        const int mpad = round_up16(mlive);
        mbar_wait(&gs->acc_full[acc.buf], acc.phase);
        tc_fence_after();
        const uint32_t taddr = gs->tmem_base + acc.buf * 256u + ((uint32_t)(q * 32) << 16);
        if (t.epilogue == EPI_PARTIAL_F32) {
          float* out = reinterpret_cast<float*>(t.out) + (size_t)split * t.M_cap * t.N;
          const float b = (t.bias != nullptr && split == 0) ? t.bias[n] : 0.f;
          for (int c0 = 0; c0 < mpad; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int m = m0 + c0 + j;
              if (c0 + j < mlive) out[(size_t)m * t.N + n] = __uint_as_float(v[j]) + b;
            }
          }
        } else if (t.epilogue == EPI_SWIGLU_BF16) {
          // gate in columns [0,mpad), up in columns [128,128+mpad).  HF Qwen2MLP rounding points
          // (transformers Qwen2MLP.forward: down(act(gate(x)) * up(x)), every op in bf16): g,u -> bf16; silu(g) -> bf16;
          // product -> bf16.  Destination: packed-X layout of the down projection, feature n is its k index.
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(t.out);
          const int kc = n >> 6, c = n & 63;
          for (int c0 = 0; c0 < mpad; c0 += 16) {
            uint32_t g[16], up[16];
            tmem_ld16(taddr + c0, g);
            tmem_ld16(taddr + 128u + c0, up);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int m = m0 + c0 + j;
              if (c0 + j < mlive) {
                const float gf = bf16_round(__uint_as_float(g[j]));
                const float uf = bf16_round(__uint_as_float(up[j]));
                const float s = bf16_round(gf / (1.f + __expf(-gf)));
                const __nv_bfloat16 r = __float2bfloat16_rn(s * uf);
                *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(dst) +
                                                  ((size_t)kc * t.out_kchunks_rows) * 128 + sw128_offset(m, c)) = r;
              }
            }
          }
        } else {  // EPI_ARGMAX
          // bf16-rounded logits (reference: lm_head output is bf16, then .float(); modeling/..._unified.py:463),
          // argmax with FIRST index winning ties (torch.argmax on CPU, :476).  Each warp reduces its 32 vocab rows
          // with shuffles and folds the result into its private running best in shared memory.
          float* bv = gs->best_val + q * gs->argmax_cap;
          int* bi = gs->best_idx + q * gs->argmax_cap;
          const bool valid = n < t.N;
          for (int c0 = 0; c0 < mpad; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
            float rx = -INFINITY;
            int ridx = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = valid ? bf16_round(__uint_as_float(v[j])) : -INFINITY;
              int idx = valid ? n : 0x7fffffff;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                const float ox = __shfl_xor_sync(0xffffffffu, x, o);
                const int oidx = __shfl_xor_sync(0xffffffffu, idx, o);
                if (ox > x || (ox == x && oidx < idx)) { x = ox; idx = oidx; }
              }
              if (lane == j) { rx = x; ridx = idx; }
            }
            const int m = m0 + c0 + lane;
            if (lane < 16 && c0 + lane < mlive) {
              const float cur = bv[m];
              const int curi = bi[m];
              if (rx > cur || (rx == cur && ridx < curi)) { bv[m] = rx; bi[m] = ridx; }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&gs->acc_empty[acc.buf]);
        acc.advance();
      }
    }
    if (t.epilogue == EPI_ARGMAX) {
      // per-CTA, per-warp candidates: out [nctas][4][M_cap] (value), out2 same shape (index); the policy step
      // finishes the reduction over nctas*4 candidates per row with the same tie rule.
      __syncwarp();
      float* ov = reinterpret_cast<float*>(t.out) + ((size_t)cta * 4 + q) * t.M_cap;
      int* oi = reinterpret_cast<int*>(t.out2) + ((size_t)cta * 4 + q) * t.M_cap;
      for (int m = lane; m < t.M; m += 32) {
        ov[m] = gs->best_val[q * gs->argmax_cap + m];
        oi[m] = gs->best_idx[q * gs->argmax_cap + m];
      }
    }
  }
}

}  // namespace jf
