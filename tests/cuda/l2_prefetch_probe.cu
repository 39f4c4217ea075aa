// l2_prefetch_probe.cu — does cp.async.bulk.prefetch.L2 make a later bulk (TMA-engine) load L2-resident, and how fast do
// 148 SMs stream L2-resident 16 KiB tiles into shared memory compared with streaming them from HBM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I jacobiforcing_b200/csrc tests/cuda/l2_prefetch_probe.cu -o build/l2_prefetch_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "jf_ptx.cuh"
using namespace jf;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

constexpr int kTile = 16384;
constexpr int kStages = 8;

// mode 0: stream tiles [first, first+n) of this CTA's region with bulk loads (ring of kStages), report cycles
// mode 1: issue L2 prefetches for the region, then wait `delay_ns`, then stream as mode 0
__global__ void __launch_bounds__(64, 1) probe(const uint8_t* base, size_t region_bytes, int tiles_per_cta, int mode,
                                               unsigned delay_ns, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  const uint8_t* mine = base + (size_t)blockIdx.x * region_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (mode == 2 && threadIdx.x >= 32) {          // warp 1: per-line prefetch.global.L2 (128 B per lane per instruction)
    const int lane = threadIdx.x - 32;
    for (size_t o = (size_t)lane * 128; o < (size_t)tiles_per_cta * kTile; o += 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(mine + o));
  }
  if (mode == 3 && threadIdx.x >= 32) {          // warp 1: discard loads (ld.global.cg, 16 B per lane)
    const int lane = threadIdx.x - 32;
    unsigned acc = 0;
    for (size_t o = (size_t)lane * 16; o < (size_t)tiles_per_cta * kTile; o += 32 * 16 * 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) if (o + (size_t)u * 512 < (size_t)tiles_per_cta * kTile) v[u] = __ldcg(reinterpret_cast<const uint4*>(mine + o + (size_t)u * 512)); else v[u] = make_uint4(0,0,0,0);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc ^= v[u].x;
    }
    if (acc == 0x12345678u) cycles[0] = 0;
  }
  if (mode == 4 && threadIdx.x >= 32) {          // warp 1: fire-and-forget 16-byte cp.async into a 512-byte scratch
    const int lane = threadIdx.x - 32;
    __shared__ __align__(16) uint8_t scratch[512];
    for (size_t o = (size_t)lane * 16; o < (size_t)tiles_per_cta * kTile; o += 512) cp_async_16(scratch + lane * 16, mine + o, 16);
    cp_async_commit();
    cp_async_wait_all();
  }
  if (mode >= 2) __syncthreads();
  if (threadIdx.x == 0) {
    if (mode >= 1) {
      if (mode == 1)
        for (int t = 0; t < tiles_per_cta; ++t) bulk_prefetch_l2(mine + (size_t)t * kTile, kTile);
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)delay_ns * 2) { __nanosleep(200); }
    }
    const long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t phase_bits = 0;
    while (done < tiles_per_cta) {
      while (issued < tiles_per_cta && issued - done < kStages) {
        const int s = issued % kStages;
        mbar_expect_tx(&full[s], kTile);
        bulk_g2s(smem + (size_t)s * kTile, mine + (size_t)issued * kTile, kTile, &full[s]);
        ++issued;
      }
      const int s = done % kStages;
      mbar_wait(&full[s], (phase_bits >> s) & 1u);
      phase_bits ^= (1u << s);
      ++done;
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const size_t total = (size_t)4 << 30;                    // 4 GiB of "weights"
  uint8_t* buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 1, total));
  long long* cyc; CK(cudaMalloc(&cyc, sms * 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kTile));
  std::vector<long long> h(sms);
  auto run = [&](const char* name, size_t offset, int tiles, int mode, unsigned delay_ns) {
    const size_t region = (size_t)tiles * kTile;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    probe<<<sms, 64, kStages * kTile>>>(buf + offset, region, tiles, mode, delay_ns, cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; double avg = 0; for (auto c : h) { mx = c > mx ? c : mx; avg += c; } avg /= sms;
    const double bytes = (double)region * sms;
    const double us = mx / 1.965e3;
    printf("%-44s tiles/SM=%3d (%.1f MB)  stream %.2f us (max CTA)  -> %.2f TB/s\n", name, tiles, bytes / 1e6, us, bytes / us * 1e-6);
  };
  size_t off = 0;
  auto next = [&](int tiles) { size_t o = off; off += (size_t)tiles * kTile * sms; return o; };
  for (int tiles : {16, 32, 48}) {
    size_t o = next(tiles);
    run("cold (HBM)", o, tiles, 0, 0);
    run("same region again (L2-warm by real loads)", o, tiles, 0, 0);
    size_t o2 = next(tiles);
    run("prefetch.L2 then wait 20 us then stream", o2, tiles, 1, 20000);
    size_t o3 = next(tiles);
    run("prefetch.L2 then wait 2 us then stream", o3, tiles, 1, 2000);
    size_t o4 = next(tiles);
    run("per-line prefetch.global.L2 by one warp, wait 20 us", o4, tiles, 2, 20000);
    size_t o6 = next(tiles);
    run("cp.async(16B) discard by one warp, wait 1 us", o6, tiles, 4, 1000);
    size_t o5 = next(tiles);
    run("discard loads by one warp (timed separately), wait 1 us", o5, tiles, 3, 1000);
  }
  printf("PROBE DONE\n");
  return 0;
}
