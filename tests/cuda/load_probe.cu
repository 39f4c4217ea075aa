// load_probe.cu — developer aid: cost of a batch of 20 independent 16-byte loads per thread from L2-resident data, by
// load flavour (the fused GEMM tails poll tagged partial sums with such batches).  148 CTAs x 128 threads; every thread
// reads 20 distinct 16-byte words written by ANOTHER CTA in the previous phase; cycles of CTA 0 / warp 0 per batch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o load_probe load_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int F> __device__ __forceinline__ ulonglong2 ld(const unsigned long long* p) {
  ulonglong2 v;
  if (F == 0) asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  if (F == 1) asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  if (F == 2) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  if (F == 3) asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  if (F == 4) asm volatile("ld.acquire.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}

template <int F>
__global__ void __launch_bounds__(128, 1) probe(unsigned long long* buf, long long* out, int reps) {
  const int cta = blockIdx.x, n = gridDim.x, tid = threadIdx.x;
  const size_t per_cta = 128 * 20 * 2;                       // words
  unsigned long long acc = 0;
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    unsigned long long* mine = buf + (size_t)cta * per_cta;   // phase 1: write my region
    for (int i = tid; i < (int)per_cta; i += 128) mine[i] = (unsigned long long)(r + 1) * 1000003ull + i;
    __threadfence();
    __syncthreads();
    // crude grid barrier through the host is avoided: regions of the NEXT cta were written a phase earlier at worst;
    // we only time the loads, values are not checked
    const unsigned long long* src = buf + (size_t)((cta + 1) % n) * per_cta + (size_t)tid * 2;
    const long long t0 = clock64();
    ulonglong2 v[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) v[k] = ld<F>(src + (size_t)k * 256);
#pragma unroll
    for (int k = 0; k < 20; ++k) acc += v[k].x ^ v[k].y;
    const long long t1 = clock64();
    if (r > 0) total += t1 - t0;
    __syncthreads();
  }
  if (cta == 0 && tid == 0) out[0] = total;
  if (acc == 0x1234567ull) out[1] = (long long)acc;
}

template <int F> static void run(const char* name, int sms, int khz) {
  unsigned long long* buf; long long* out;
  cudaMalloc(&buf, (size_t)sms * 128 * 20 * 2 * 8); cudaMalloc(&out, 64);
  cudaMemset(buf, 0, (size_t)sms * 128 * 20 * 2 * 8);
  const int reps = 200;
  probe<F><<<sms, 128>>>(buf, out, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s FAILED %s\n", name, cudaGetErrorString(e)); return; }
  long long cyc = 0; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
  printf("%-22s %7.3f us per batch of 20 x 16 B loads per thread (attribute clock)\n", name, (double)cyc / (reps - 1) / (khz / 1e3));
  cudaFree(buf); cudaFree(out);
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  for (int rep = 0; rep < 2; ++rep) {
    run<0>("ld.global (weak)", sms, khz);
    run<1>("ld.global.cg", sms, khz);
    run<2>("ld.relaxed.gpu", sms, khz);
    run<3>("ld.volatile", sms, khz);
    run<4>("ld.acquire.gpu", sms, khz);
  }
  printf("PROBE DONE\n");
  return 0;
}
