// jf_pack.cuh — HBM data layout helpers: re-tile row-major bf16 matrices into the UMMA-ready layout the
// streaming GEMM consumes (see jf_ptx.cuh for the swizzle definition).
//
//   weights  W[N][K] row-major  ->  tiles [N/128][K/64] of 128 rows x 64 cols, each tile 16 KiB contiguous, SW128
//   rows     X[M][K] row-major  ->  chunks [K/64] of rows_cap rows x 64 cols, SW128 (row r of chunk kc at
//                                   kc*rows_cap*128 + sw128_offset(r, c))
//
// The weight re-tiling happens once at load time (jf_load_tensor), replacing the reference loader's plain
// param.data.copy_ (inference_engine/utils/loader.py:27-47, layers/linear.py:65-128).
#pragma once
#include "jf_ptx.cuh"

namespace jf {

// one thread per 16-byte chunk (8 bf16) of the destination
__global__ void pack_weight_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                   int N, int K, int n_tiles, int k_chunks, long long src_row_stride) {
  const long long total = (long long)n_tiles * k_chunks * 128 * 8;   // 16-byte chunks
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int chunk_phys = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long tile = i >> 10;
    const int kc = (int)(tile % k_chunks);
    const int nt = (int)(tile / k_chunks);
    const int chunk_log = chunk_phys ^ (r & 7);
    const int n = nt * 128 + r;
    const int k = kc * 64 + chunk_log * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < N && k + 8 <= K) {
      v = *reinterpret_cast<const uint4*>(src + (long long)n * src_row_stride + k);
    } else if (n < N && k < K) {
      __nv_bfloat16 tmp[8];
      for (int j = 0; j < 8; ++j) tmp[j] = (k + j < K) ? src[(long long)n * src_row_stride + k + j] : __float2bfloat16(0.f);
      v = *reinterpret_cast<uint4*>(tmp);
    }
    reinterpret_cast<uint4*>(dst)[i] = v;
  }
}

// X[M][K] row-major -> packed chunks; one thread per 16-byte chunk.  Rows M..Mpad16 are zero-filled so the padded
// MMA columns stay finite.
__global__ void pack_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int M,
                                 int K, int rows_cap, int k_chunks) {
  const int mpad = (M + 15) & ~15;
  const long long total = (long long)k_chunks * mpad * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int chunk_log = (int)(i & 7);
    const int r = (int)((i >> 3) % mpad);
    const int kc = (int)((i >> 3) / mpad);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < M) v = *reinterpret_cast<const uint4*>(src + (long long)r * K + kc * 64 + chunk_log * 8);
    uint8_t* d = reinterpret_cast<uint8_t*>(dst) + ((long long)kc * rows_cap) * 128 + sw128_offset(r, chunk_log * 8);
    *reinterpret_cast<uint4*>(d) = v;
  }
}

}  // namespace jf
