// stream_bench.cu -- how fast can one row group of the persistent kernels stream its (row, chunk) feature slab
// L2 -> shared memory?  Emulates the weighted-sum phase of recur_fwd_kernel: every CTA owns `groups` slabs of
// `item_bytes` (196 pixels x 512 B) in its own L2-resident region and, `steps` times, pulls each slab through a ring
// of S stages of b bytes with cp.async.bulk + mbarriers, consuming every stage with 16-byte shared-memory loads + FMAs.
// Variants: ring shape (S x b), who re-issues (mode 0: group barrier + elected thread, as in the kernel; mode 1:
// per-stage "empty" mbarrier, thread 0 re-issues as soon as all 8 warps have released the stage).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/stream_bench tools/stream_bench.cu && tools/stream_bench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n)); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void fill(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
// the same fill issued as `parts` sub-copies by lane 0 of `parts` different warps (mbarrier count = parts)
__device__ __forceinline__ void fill_part(uint32_t dst, const uint8_t* src, uint32_t bytes, uint32_t bar, int part, int parts) {
  const uint32_t per = ((bytes / parts) + 15u) & ~15u;
  const uint32_t lo = min(bytes, per * part), hi = min(bytes, lo + per);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(hi - lo) : "memory");
  if (hi > lo)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + lo), "l"(src + lo),
                 "r"(hi - lo), "r"(bar) : "memory");
}

constexpr int GT = 256;

template <int S, int MODE>
__global__ void __launch_bounds__(512, 1) stream_kernel(const uint8_t* src, int item_bytes, int stage_bytes, int steps, int active,
                                                       int skew, long long* cycles, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bars[2][2 * 8];
  const int g = threadIdx.x / GT, tid = threadIdx.x % GT, warp = tid >> 5, lane = tid & 31;
  uint8_t* stg = sm + (size_t)g * S * stage_bytes;
  const uint32_t stg_a = smem_u32(stg);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bars[g][s]), 1);
      mbar_init(smem_u32(&bars[g][8 + s]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((int)blockIdx.x >= active) return;
  if (g == 1 && skew > 0) { const long long t0 = clock64(); while (clock64() - t0 < skew) __nanosleep(200); }
  const uint8_t* mine = src + ((size_t)blockIdx.x * 2 + g) * 131072;
  const int nfill = (item_bytes + stage_bytes - 1) / stage_bytes;
  uint32_t full_par = 0, empty_par = 0;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int vec_per_stage = stage_bytes / 16;
  long long tot = 0;
  for (int st = 0; st < steps; ++st) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory");
    const long long t0 = clock64();
    if (tid == 0)
      for (int s = 0; s < S && s < nfill; ++s)
        fill(stg_a + s * stage_bytes, mine + (size_t)s * stage_bytes, min(stage_bytes, item_bytes - s * stage_bytes), smem_u32(&bars[g][s]));
    for (int fi = 0; fi < nfill; ++fi) {
      const int s = fi % S;
      mbar_wait(smem_u32(&bars[g][s]), (full_par >> s) & 1u);
      full_par ^= 1u << s;
      const int nb = min(stage_bytes, item_bytes - fi * stage_bytes) / 16;
      for (int v = tid; v < nb && v < vec_per_stage; v += GT) {
        const uint4 raw = *reinterpret_cast<const uint4*>(stg + (size_t)s * stage_bytes + (size_t)v * 16);
        const float w = (float)(v & 7);
        acc[0] = fmaf(w, __uint_as_float(raw.x << 16), acc[0]); acc[1] = fmaf(w, __uint_as_float(raw.x & 0xffff0000u), acc[1]);
        acc[2] = fmaf(w, __uint_as_float(raw.y << 16), acc[2]); acc[3] = fmaf(w, __uint_as_float(raw.y & 0xffff0000u), acc[3]);
        acc[4] = fmaf(w, __uint_as_float(raw.z << 16), acc[4]); acc[5] = fmaf(w, __uint_as_float(raw.z & 0xffff0000u), acc[5]);
        acc[6] = fmaf(w, __uint_as_float(raw.w << 16), acc[6]); acc[7] = fmaf(w, __uint_as_float(raw.w & 0xffff0000u), acc[7]);
      }
      if (fi + S < nfill) {
        const int nx = fi + S;
        if (MODE == 0) {
          asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory");
          if (tid == 0)
            fill(stg_a + s * stage_bytes, mine + (size_t)nx * stage_bytes, min(stage_bytes, item_bytes - nx * stage_bytes), smem_u32(&bars[g][s]));
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars[g][8 + s]));
          if (tid == 0) {
            mbar_wait(smem_u32(&bars[g][8 + s]), (empty_par >> s) & 1u);
            empty_par ^= 1u << s;
            fill(stg_a + s * stage_bytes, mine + (size_t)nx * stage_bytes, min(stage_bytes, item_bytes - nx * stage_bytes), smem_u32(&bars[g][s]));
          }
        }
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory");
    tot += clock64() - t0;
  }
  float sum = 0.f;
  for (int k = 0; k < 8; ++k) sum += acc[k];
  if (sum == 123.456f) sink[0] = sum;
  if (tid == 0) cycles[blockIdx.x * 2 + g] = tot;
}

// whole slab requested up front (S stages >= slab), every fill issued as `parts` sub-copies by different warps
template <int S>
__global__ void __launch_bounds__(512, 1) upfront_kernel(const uint8_t* src, int item_bytes, int stage_bytes, int steps, int active,
                                                        int parts, int consume, long long* cycles, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bars[2][8];
  const int g = threadIdx.x / GT, tid = threadIdx.x % GT, warp = tid >> 5, lane = tid & 31;
  uint8_t* stg = sm + (size_t)g * S * stage_bytes;
  const uint32_t stg_a = smem_u32(stg);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&bars[g][s]), parts);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((int)blockIdx.x >= active) return;
  const uint8_t* mine = src + ((size_t)blockIdx.x * 2 + g) * 131072;
  const int nfill = (item_bytes + stage_bytes - 1) / stage_bytes;
  uint32_t par = 0;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  long long tot = 0;
  for (int st = 0; st < steps; ++st) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory");
    const long long t0 = clock64();
    if (lane == 0 && warp < parts)
      for (int s = 0; s < nfill; ++s)
        fill_part(stg_a + s * stage_bytes, mine + (size_t)s * stage_bytes, min(stage_bytes, item_bytes - s * stage_bytes),
                  smem_u32(&bars[g][s]), warp, parts);
    for (int fi = 0; fi < nfill; ++fi) {
      mbar_wait(smem_u32(&bars[g][fi]), par);
      if (consume) {
        const int nb = min(stage_bytes, item_bytes - fi * stage_bytes) / 16;
#pragma unroll 4
        for (int v = tid; v < nb; v += GT) {
          const uint4 raw = *reinterpret_cast<const uint4*>(stg + (size_t)fi * stage_bytes + (size_t)v * 16);
          const float w = (float)(v & 7);
          acc[0] = fmaf(w, __uint_as_float(raw.x << 16), acc[0]); acc[1] = fmaf(w, __uint_as_float(raw.x & 0xffff0000u), acc[1]);
          acc[2] = fmaf(w, __uint_as_float(raw.y << 16), acc[2]); acc[3] = fmaf(w, __uint_as_float(raw.y & 0xffff0000u), acc[3]);
          acc[4] = fmaf(w, __uint_as_float(raw.z << 16), acc[4]); acc[5] = fmaf(w, __uint_as_float(raw.z & 0xffff0000u), acc[5]);
          acc[6] = fmaf(w, __uint_as_float(raw.w << 16), acc[6]); acc[7] = fmaf(w, __uint_as_float(raw.w & 0xffff0000u), acc[7]);
        }
      }
    }
    par ^= 1u;
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory");
    tot += clock64() - t0;
  }
  float sum = 0.f;
  for (int k = 0; k < 8; ++k) sum += acc[k];
  if (sum == 123.456f) sink[0] = sum;
  if (tid == 0) cycles[blockIdx.x * 2 + g] = tot;
}

template <int S>
void run_upfront(int sms, int active, int threads, int stage_bytes, int steps, int parts, int consume, int item_bytes = 196 * 512) {
  uint8_t* src; long long* cyc; float* sink;
  CK(cudaMalloc(&src, (size_t)sms * 2 * 131072)); CK(cudaMemset(src, 1, (size_t)sms * 2 * 131072));
  CK(cudaMalloc(&cyc, sms * 2 * 8)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(cyc, 0, sms * 2 * 8));
  const size_t smem = (size_t)2 * S * stage_bytes;
  CK(cudaFuncSetAttribute(upfront_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) {
    upfront_kernel<S><<<sms, threads, smem>>>(src, item_bytes, stage_bytes, steps, active, parts, consume, cyc, sink);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(sms * 2);
  CK(cudaMemcpy(h.data(), cyc, sms * 2 * 8, cudaMemcpyDeviceToHost));
  double mean = 0; int cnt = 0;
  for (int i = 0; i < active * 2; ++i) if (h[i]) { mean += (double)h[i] / steps; ++cnt; }
  mean /= cnt;
  const int groups = threads / GT;
  printf("upfront %d x %5d B (%6d B), %d issuing warp(s), consume %d, %d group(s)/CTA, %3d CTAs: %7.0f cycles = %5.1f B/clk/group, %6.0f B/clk chip\n",
         S, stage_bytes, item_bytes, parts, consume, groups, active, mean, item_bytes / mean, (double)item_bytes * active * groups / mean);
  cudaFree(src); cudaFree(cyc); cudaFree(sink);
}

template <int S, int MODE>
void run(int sms, int active, int threads, int stage_bytes, int steps, int skew) {
  const int item_bytes = 196 * 512;
  uint8_t* src; long long* cyc; float* sink;
  CK(cudaMalloc(&src, (size_t)sms * 2 * 131072)); CK(cudaMemset(src, 1, (size_t)sms * 2 * 131072));
  CK(cudaMalloc(&cyc, sms * 2 * 8)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(cyc, 0, sms * 2 * 8));
  const size_t smem = (size_t)2 * S * stage_bytes;
  CK(cudaFuncSetAttribute(stream_kernel<S, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) {
    stream_kernel<S, MODE><<<sms, threads, smem>>>(src, item_bytes, stage_bytes, steps, active, skew, cyc, sink);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(sms * 2);
  CK(cudaMemcpy(h.data(), cyc, sms * 2 * 8, cudaMemcpyDeviceToHost));
  double mean = 0; int cnt = 0;
  for (int i = 0; i < active * 2; ++i) if (h[i]) { mean += (double)h[i] / steps; ++cnt; }
  mean /= cnt;
  const int groups = threads / GT;
  printf("ring %d x %5d B, reissue %s, %d group(s)/CTA, %3d CTAs, skew %5d: %7.0f cycles per 100 KB slab = %5.1f B/clk/group, %6.0f B/clk chip\n",
         S, stage_bytes, MODE ? "empty-mbarrier" : "group barrier ", groups, active, skew, mean, item_bytes / mean,
         (double)item_bytes * active * groups / mean);
  cudaFree(src); cudaFree(cyc); cudaFree(sink);
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int steps = 50;
  for (int bytes : {16384, 32768, 49152, 65536, 81920, 98304, 100352, 114688})
    run_upfront<7>(sms, 128, 256, 16384, steps, 1, 0, bytes);
  for (int bytes : {65536, 98304, 100352, 131072}) run_upfront<4>(sms, 128, 256, 32768, steps, 1, 0, bytes);
  for (int bytes : {65536, 98304, 100352, 131072}) run_upfront<2>(sms, 128, 256, 65536, steps, 1, 0, bytes);
  for (int bytes : {65536, 98304, 100352}) run_upfront<1>(sms, 128, 256, 100352, steps, 1, 0, bytes);
  for (int bytes : {50176, 100352}) run_upfront<7>(sms, 128, 512, 16384, steps, 1, 0, bytes);
  for (int bytes : {50176, 100352}) run_upfront<7>(sms, 128, 512, 16384, steps, 1, 1, bytes);
  run_upfront<7>(sms, 128, 256, 16384, steps, 1, 1, 100352);
  run<2, 0>(sms, 128, 512, 16384, steps, 3000);
  run<4, 1>(sms, 128, 512, 8192, steps, 3000);
  return 0;
}
