// umma_rate_probe.cu -- how fast does tcgen05.mma.kind::f16 M128 x N x K16 run back to back, as a function of the
// shared-memory operand layout?  (developer probe behind the K4 tensor-core kernel, selfsim_tc.cuh)
//
//   layout 0: no-swizzle K-major, canonical [k chunk][row][16 B]   (LBO = 2048 B, SBO = 128 B)
//   layout 1: no-swizzle K-major "in place" record array, decim 1  (LBO = 128 B,  SBO = 128 B: overlapping core matrices)
//   layout 2: the same for decim 4                                  (LBO = 32 B,   SBO = 128 B)
//   layout 3: SWIZZLE_128B K-major canonical                        (SBO = 1024 B)
// accumulators: 1 = every MMA into the same TMEM region, 2 / 3 = round robin over 2 / 3 regions.
// Every SM runs one CTA; one elected lane issues `n` MMAs, commits, waits; cycles per MMA = (t1 - t0) / n.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate_probe tools/umma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128, 1) k_rate(int layout, int nAcc, int N, int n, int ksteps, long long *out) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  unsigned char *base = smemRaw + ((1024 - (smem_u32(smemRaw) & 1023)) & 1023);
  __shared__ uint32_t tmemSlot;
  __shared__ __align__(8) uint64_t bar, bar2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x3c003c00u;   // FP16 ones
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmemSlot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmemSlot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(base), sb = smem_u32(base + 48 * 1024);
    uint32_t lbo, sbo, lay, inc;
    if (layout == 0) { lbo = 2048; sbo = 128; lay = 0; inc = 256; }        // next K step: 2 chunks x 2048 B
    else if (layout == 1) { lbo = 128; sbo = 128; lay = 0; inc = 16; }     // 2 chunks x 8 records x 16 B
    else if (layout == 2) { lbo = 32; sbo = 128; lay = 0; inc = 4; }       // 2 chunks x 2 records x 16 B
    else { lbo = 16; sbo = 1024; lay = 2; inc = 2; }                       // SWIZZLE_128B: +32 B per K step
    const uint64_t da0 = desc(sa, lbo, sbo, lay), db0 = desc(sb, lbo, sbo, lay);
    // descriptors in registers before the timed region; the loop body is 6 x nAcc MMAs with compile-time indexing, so the
    // single issuing thread spends ~2 instructions per MMA (a dependent instruction costs >= 4 cycles on one thread)
    uint64_t da[6], db[6];
#pragma unroll
    for (int ks = 0; ks < 6; ks++) { da[ks] = da0 + (uint64_t)inc * (ksteps == 1 ? 0 : ks); db[ks] = db0 + (uint64_t)inc * (ksteps == 1 ? 0 : ks); }
    const uint32_t d0 = tmem, d1 = tmem + (nAcc > 1 ? N : 0), d2 = tmem + (nAcc > 2 ? 2 * N : 0);
#define MMA(D, A, B)                                                                                                    \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(D), \
               "l"(A), "l"(B), "r"(idesc)                                                                               \
               : "memory")
    const long long t0 = clock64();
    if (ksteps == 7) {       // the K4 kernel's channel: 6 main products, then 12 corrections (two operand parts), one commit
      uint64_t da2[6], db2[6];
#pragma unroll
      for (int k = 0; k < 6; k++) { da2[k] = da[k] + 140; db2[k] = db[k] + 140; }   // second parts 2240 B further
      for (int i = 0; i < n; i += 18) {
#pragma unroll
        for (int k = 0; k < 6; k++) MMA(d0, da[k], db[k]);
#pragma unroll
        for (int k = 0; k < 6; k++) { MMA(d1, da2[k], db[k]); MMA(d1, da[k], db2[k]); }
        if (nAcc == 3) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
    } else if (ksteps == 1) {       // change the accumulator after every MMA
      for (int i = 0; i < n; i += 18) {
#pragma unroll
        for (int k = 0; k < 6; k++) { MMA(d0, da[k], db[k]); MMA(d1, da[k], db[k]); MMA(d2, da[k], db[k]); }
      }
    } else {                 // chains of 6 MMAs per accumulator
      for (int i = 0; i < n; i += 18) {
#pragma unroll
        for (int k = 0; k < 6; k++) MMA(d0, da[k], db[k]);
#pragma unroll
        for (int k = 0; k < 6; k++) MMA(d1, da[k], db[k]);
#pragma unroll
        for (int k = 0; k < 6; k++) MMA(d2, da[k], db[k]);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long *d;
  cudaMalloc(&d, sms * sizeof(long long));
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 3600;
  const char *names[4] = {"no-swizzle canonical (LBO 2048)", "no-swizzle in place decim 1 (LBO 128)", "no-swizzle in place decim 4 (LBO 32)",
                          "SWIZZLE_128B canonical"};
  for (int N : {128, 64, 256})
    for (int layout = 0; layout < 4; layout++)
      for (int nAcc : {1, 2, 3}) {
        if (nAcc * N > 512) continue;
        for (int ksteps : {1, 6, 7}) {
          if (ksteps == 1 && nAcc > 1 && N != 128) continue;
          if (ksteps == 7 && (N != 128 || nAcc < 2 || layout > 1)) continue;
          k_rate<<<sms, 128, smem>>>(layout, nAcc, N, n, ksteps, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
          long long h[256];
          cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < sms; i++) s += (double)h[i];
          printf("{\"shape\": \"M128 N%d K16 f16\", \"layout\": \"%s\", \"accumulators\": %d, \"chain\": \"%s\", \"cycles_per_mma\": %.1f}\n", N,
                 names[layout], nAcc, ksteps == 7 ? (nAcc == 3 ? "K4 channel pattern + commit per 18" : "K4 channel pattern") : (ksteps == 1 ? "1" : "6"), s / sms / n);
        }
      }
  return 0;
}
