// barrier_probe.cu — stand-alone microbenchmark of grid-barrier variants for the persistent Jacobi kernel (developer
// aid, not part of the product or the tests).  148 CTAs x 128 threads, cooperative launch, ITER barriers back to back;
// prints microseconds per barrier.
//   A  one counter: red.release.gpu + relaxed polling + one acquire fence (what jf_kernel.cuh does today)
//   B  flag array: every CTA stores its epoch in flags[cta] (st.release), every CTA reads ALL flags with wide relaxed
//      loads until all reached the epoch (no atomics, no single hot address)
//   C  two-level counters: 4 groups arrive on group counters, the last arriver of a group (atom) arrives on the top
//      counter, everybody polls the top counter
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_probe barrier_probe.cu ; run: ./barrier_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define SPIN_LIMIT 2000000000LL
__device__ __forceinline__ void trap_timeout() { asm volatile("trap;"); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_acq_rel(unsigned* p, unsigned v) {
  unsigned o; asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

constexpr int kGroups = 4;

template <int VARIANT>
__global__ void __launch_bounds__(128, 1) probe(unsigned* ctr, unsigned* flags, unsigned* grp, int iters, long long* out,
                                               float* payload) {
  const int cta = blockIdx.x, n = gridDim.x, tid = threadIdx.x;
  __shared__ int ok_s;
  long long t0 = 0;
  for (int it = 0; it <= iters; ++it) {
    if (it == 1 && cta == 0 && tid == 0) t0 = clock64();        // iteration 0 is a warm-up
    payload[(size_t)cta * 128 + tid] += 1.0f;                     // a store the barrier has to publish
    __syncthreads();
    const unsigned epoch = (unsigned)it + 1u;
    if (VARIANT == 0) {
      if (tid == 0) {
        red_release(ctr, 1u);
        const long long s = clock64();
        while (ld_relaxed(ctr) < epoch * (unsigned)n) if (clock64() - s > SPIN_LIMIT) trap_timeout();
        fence_acq_rel();
      }
    } else if (VARIANT == 1) {
      if (tid == 0) st_release(flags + cta, epoch);
      const long long s = clock64();
      for (;;) {                                                   // 128 threads read up to 256 flags per probe
        bool ok = true;
        for (int i = tid; i < n; i += 128) ok = ok && (ld_relaxed(flags + i) >= epoch);
        const int all = __syncthreads_and(ok ? 1 : 0);
        if (all) break;
        if (clock64() - s > SPIN_LIMIT) trap_timeout();
      }
      if (tid == 0) fence_acq_rel();
    } else {
      if (tid == 0) {
        const int g = cta % kGroups;
        const unsigned members = (unsigned)((n - g + kGroups - 1) / kGroups);
        const unsigned old = atom_acq_rel(grp + g * 32, 1u);      // one 128-byte line per group counter
        if (old + 1u == epoch * members) red_release(ctr, 1u);
        const long long s = clock64();
        while (ld_relaxed(ctr) < epoch * (unsigned)kGroups) if (clock64() - s > SPIN_LIMIT) trap_timeout();
        fence_acq_rel();
      }
    }
    __syncthreads();
  }
  if (cta == 0 && tid == 0) out[0] = clock64() - t0;
  (void)ok_s;
}

template <int VARIANT>
static void run(const char* name, int sms, int iters, int khz) {
  unsigned *ctr, *flags, *grp; long long* out; float* payload;
  cudaMalloc(&ctr, 256); cudaMalloc(&flags, 4096); cudaMalloc(&grp, 4096); cudaMalloc(&out, 64);
  cudaMalloc(&payload, (size_t)sms * 128 * 4);
  cudaMemset(ctr, 0, 256); cudaMemset(flags, 0, 4096); cudaMemset(grp, 0, 4096); cudaMemset(payload, 0, (size_t)sms * 128 * 4);
  void* args[] = {&ctr, &flags, &grp, &iters, &out, &payload};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaError_t err = cudaLaunchCooperativeKernel((void*)probe<VARIANT>, dim3(sms), dim3(128), args, 0, 0);
  cudaEventRecord(e1);
  err = err == cudaSuccess ? cudaDeviceSynchronize() : err;
  if (err != cudaSuccess) { printf("%-28s FAILED: %s\n", name, cudaGetErrorString(err)); exit(1); }
  long long cyc = 0; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s %7.3f us / barrier by clock64 at the attribute clock, %7.3f us by events (%d barriers)\n", name,
         (double)cyc / iters / (khz / 1e3), ms * 1e3 / (iters + 1), iters);
  cudaFree(ctr); cudaFree(flags); cudaFree(grp); cudaFree(out); cudaFree(payload);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    run<0>("A counter + relaxed poll", sms, iters, khz);
    run<1>("B flag array, wide reads", sms, iters, khz);
    run<2>("C two-level counters", sms, iters, khz);
  }
  printf("PROBE DONE\n");
  return 0;
}
