// barrier_bench.cu -- microbenchmark of grid-barrier variants for the persistent kernels (recur.cu).
// One cooperative launch, 148 CTAs x 256 threads x `groups` independent barrier domains per CTA is NOT
// modelled: one domain, every CTA does a little work (a global store per thread) between barriers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/barrier_bench tools/barrier_bench.cu
//   tools/barrier_bench [iters]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// v0: the barrier of recur.cu: release-add on one counter, acquire-spin on the same counter
__device__ __forceinline__ void bar_v0(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// v1: per-CTA epoch flags (one 4-byte slot per CTA), no read-modify-write: arrive = st.release of the epoch
// into the CTA's slot; wait = warp 0 reads all slots (relaxed), loops until all >= epoch, then one acquire fence
__device__ __forceinline__ void bar_v1(unsigned* flags, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    const int n = gridDim.x;
    bool ok;
    do {
      ok = true;
      for (int i = threadIdx.x; i < n; i += 32) {
        unsigned v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        ok = ok && (int)(v - epoch) >= 0;
      }
      ok = __all_sync(0xffffffffu, ok);
    } while (!ok);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

// v2: like v1 but every polling load is an acquire (no trailing fence)
__device__ __forceinline__ void bar_v2(unsigned* flags, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    const int n = gridDim.x;
    bool ok;
    do {
      ok = true;
      for (int i = threadIdx.x; i < n; i += 32) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        ok = ok && (int)(v - epoch) >= 0;
      }
      ok = __all_sync(0xffffffffu, ok);
    } while (!ok);
  }
  __syncthreads();
}

// v3: arrival counter + separate release flag: the last arriver (atom returns gridDim-1 mod) publishes the epoch
// on another cache line; waiters poll only that line
__device__ __forceinline__ void bar_v3(unsigned* ctr, unsigned* flag, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(ctr), "r"(1u) : "memory");
    if (old + 1 == epoch * gridDim.x) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    } else {
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      } while ((int)(v - epoch) < 0);
    }
  }
  __syncthreads();
}

// v4: v0 with the counter polled through a relaxed load and a single acquire fence at the end
__device__ __forceinline__ void bar_v4(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

// v5: per-CTA flags in separate 32-byte sectors (less false sharing between arrivals), polled like v1
__device__ __forceinline__ void bar_v5(unsigned* flags, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 8), "r"(epoch) : "memory");
    const int n = gridDim.x;
    bool ok;
    do {
      ok = true;
      for (int i = threadIdx.x; i < n; i += 32) {
        unsigned v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i * 8) : "memory");
        ok = ok && (int)(v - epoch) >= 0;
      }
      ok = __all_sync(0xffffffffu, ok);
    } while (!ok);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}


// v6: NO ordering at all (incorrect as a barrier for data; lower bound of the signalling itself)
__device__ __forceinline__ void bar_v6(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
// v7: release side only
__device__ __forceinline__ void bar_v7(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
// v8: acquire side only
__device__ __forceinline__ void bar_v8(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
// v9: v0 with a short sleep between polls
__device__ __forceinline__ void bar_v9(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    unsigned v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      __nanosleep(40);
    }
  }
  __syncthreads();
}
// v10: four counters on four lines (CTA i arrives on counter i & 3), lanes 0..3 poll one each
__device__ __forceinline__ void bar_v10(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr + (blockIdx.x & 3) * 32), "r"(1u) : "memory");
    const unsigned g = gridDim.x;
    const unsigned lane = threadIdx.x & 3;
    const unsigned want = epoch * ((g + 3 - lane) / 4);
    bool ok;
    do {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr + lane * 32) : "memory");
      ok = __all_sync(0xffffffffu, v >= want);
    } while (!ok);
  }
  __syncthreads();
}
// v11: release by a full fence executed by EVERY thread before the CTA barrier, relaxed signalling, acquire fence by
// every thread after the CTA barrier
__device__ __forceinline__ void bar_v11(unsigned* ctr, unsigned& target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    target += gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
  __threadfence();
}

template <int V>
__global__ void __launch_bounds__(256, 1) bench_kernel(unsigned* sync, float* data, int iters, int work, long long* cycles,
                                                        unsigned* check) {
  unsigned state = 0;
  unsigned* ctr = sync;            // line 0
  unsigned* flag = sync + 64;      // another line
  unsigned* flags = sync + 128;    // flag array
  const long long t0 = clock64();
  unsigned bad = 0;
  for (int it = 0; it < iters; ++it) {
    // "work": every thread writes a value the next iteration's reader (another CTA) checks
    if (work) data[(size_t)blockIdx.x * 256 + threadIdx.x] = (float)(it + 1);
    if (V == 0) bar_v0(ctr, state);
    if (V == 1) bar_v1(flags, state);
    if (V == 2) bar_v2(flags, state);
    if (V == 3) bar_v3(ctr, flag, state);
    if (V == 4) bar_v4(ctr, state);
    if (V == 5) bar_v5(flags, state);
    if (V == 6) bar_v6(ctr, state);
    if (V == 7) bar_v7(ctr, state);
    if (V == 8) bar_v8(ctr, state);
    if (V == 9) bar_v9(ctr, state);
    if (V == 10) bar_v10(ctr, state);
    if (V == 11) bar_v11(ctr, state);
    if (work) {
      const int other = (blockIdx.x + 1 + it % (gridDim.x - 1)) % gridDim.x;
      const float v = __ldcg(data + (size_t)other * 256 + threadIdx.x);
      if (v < (float)(it + 1)) ++bad;      // stale: the barrier leaked
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (bad) atomicAdd(check, bad);
}

template <int V>
void run(const char* name, int sms, int iters, int work) {
  unsigned* sync; float* data; long long* cyc; unsigned* check;
  CK(cudaMalloc(&sync, 65536)); CK(cudaMemset(sync, 0, 65536));
  CK(cudaMalloc(&data, (size_t)sms * 256 * 4)); CK(cudaMemset(data, 0, (size_t)sms * 256 * 4));
  CK(cudaMalloc(&cyc, sms * 8)); CK(cudaMalloc(&check, 4)); CK(cudaMemset(check, 0, 4));
  void* args[] = {&sync, &data, &iters, &work, &cyc, &check};
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemset(sync, 0, 65536));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((void*)bench_kernel<V>, dim3(sms), dim3(256), args, 0, 0));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[256]; CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
  unsigned bad; CK(cudaMemcpy(&bad, check, 4, cudaMemcpyDeviceToHost));
  printf("%-44s work=%d  %8.1f cycles/barrier (CTA 0)  %7.3f us/barrier (events)  stale=%u\n", name, work,
         (double)h[0] / iters, ms * 1000.0 / iters, bad);
  cudaFree(sync); cudaFree(data); cudaFree(cyc); cudaFree(check);
}


// ---- latency of one TMA bulk copy L2 -> shared memory, every CTA copying from its own L2-resident region ----
__global__ void __launch_bounds__(256, 1) tma_kernel(const uint8_t* src, int bytes, int iters, int ncopies, long long* cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* mine = src + (size_t)blockIdx.x * 262144;
  long long tot = 0;
  uint32_t parity = 0;
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes * ncopies) : "memory");
      for (int c = 0; c < ncopies; ++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + c * bytes),
                     "l"(mine + (size_t)((it * ncopies + c) % 4) * bytes), "r"(bytes), "r"(bar_a) : "memory");
    }
    uint32_t done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done) : "r"(bar_a), "r"(parity) : "memory");
    } while (!done);
    parity ^= 1;
    tot += clock64() - t0;
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = tot;
}

void run_tma(int sms, int ctas, int bytes, int ncopies) {
  uint8_t* src; long long* cyc;
  CK(cudaMalloc(&src, (size_t)sms * 262144)); CK(cudaMemset(src, 1, (size_t)sms * 262144));
  CK(cudaMalloc(&cyc, sms * 8));
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  const int iters = 200;
  for (int rep = 0; rep < 2; ++rep) {
    tma_kernel<<<ctas, 256, 131072>>>(src, bytes, iters, ncopies, cyc);
    CK(cudaDeviceSynchronize());
  }
  long long h[256]; CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < ctas; ++i) mean += (double)h[i] / iters; mean /= ctas;
  printf("bulk copy L2->smem: %3d CTAs, %d x %6d B: %8.1f cycles (CTA 0), %8.1f mean\n", ctas, ncopies, bytes, (double)h[0] / iters, mean);
  cudaFree(src); cudaFree(cyc);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d, iters %d\n", sms, iters);
  for (int work = 0; work < 2; ++work) {
    run<0>("v0 red.release + ld.acquire spin (current)", sms, iters, work);
    run<4>("v4 red.release + relaxed spin + fence", sms, iters, work);
    run<3>("v3 atom counter + separate flag line", sms, iters, work);
    run<1>("v1 per-CTA flags, relaxed poll + fence", sms, iters, work);
    run<2>("v2 per-CTA flags, acquire poll", sms, iters, work);
    run<5>("v5 per-CTA flags in own sectors", sms, iters, work);
    run<6>("v6 relaxed red + relaxed spin (NO ordering)", sms, iters, work);
    run<7>("v7 release red + relaxed spin", sms, iters, work);
    run<8>("v8 relaxed red + acquire spin", sms, iters, work);
    run<9>("v9 v0 + nanosleep(40) between polls", sms, iters, work);
    run<10>("v10 four counters, four polling lanes", sms, iters, work);
    run<11>("v11 all-thread fences + relaxed signalling", sms, iters, work);
  }
  if (argc > 2)
  for (int ctas : {1, sms})
    for (int bytes : {1024, 4096, 16384, 32768})
      for (int nc : {1, 2}) run_tma(sms, ctas, bytes, nc);
  return 0;
}
