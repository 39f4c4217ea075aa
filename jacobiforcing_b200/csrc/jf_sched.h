// jf_sched.h — the integer scheduling arithmetic of the persistent kernel that two sides must agree on (a mismatch is a
// deadlock, not a wrong number): how a GEMM's (n-tile, k-split) units cut the k-chunks, and the gate/up -> down
// hand-over (which down k-splits a gate/up tile arrives on, and how many arrivals a down unit waits for).
// Pure host/device header: the kernel includes it, and tests/host/policy_harness.cpp exposes it to the CPU tests.
#pragma once
#ifndef JF_HD
#ifdef __CUDACC__
#define JF_HD __host__ __device__
#else
#define JF_HD
#endif
#endif

namespace jf {

// k-chunks [k0, k1) of k-split `split` out of `splits` over `k_chunks` chunks (balanced, contiguous)
// (32-bit unsigned arithmetic: split <= splits <= 64 and k_chunks < 2^20, so the product cannot overflow - and a 64-bit
// division is a ~100-instruction library call on the device, paid by every role at the start of every GEMM stage)
JF_HD inline int split_k0(int split, int splits, int k_chunks) { return (int)(((unsigned)split * (unsigned)k_chunks) / (unsigned)splits); }
JF_HD inline int split_k1(int split, int splits, int k_chunks) { return (int)(((unsigned)(split + 1) * (unsigned)k_chunks) / (unsigned)splits); }

// gate/up tile `tile` (128 features) holds k-chunks 2*tile and 2*tile+1 (the last tile may hold one) of the down
// projection's K dimension (k_chunks = intermediate / 64).  Does the k-range of down k-split `split` read it?
JF_HD inline bool df_tile_feeds_split(int tile, int split, int splits, int k_chunks) {
  const int ka = 2 * tile;
  const int kb = (2 * tile + 1 < k_chunks) ? 2 * tile + 1 : k_chunks - 1;
  return ka < split_k1(split, splits, k_chunks) && kb >= split_k0(split, splits, k_chunks);
}
// number of gate/up tiles a down unit with k-range [k0, k1) waits for
JF_HD inline int df_tiles_needed(int k0, int k1) { return ((k1 - 1) >> 1) - (k0 >> 1) + 1; }

}  // namespace jf
