// umma_probe.cu -- feasibility probe for a tensor-core K1 ("Next step for K1").
//
// Question: can tcgen05.mma read the sliding windows of a planar signal row DIRECTLY, without im2col?
//   A (M = 128 rows, K-major, SWIZZLE_128B):  A[r][k] = b[32 r + k]   -- a Hankel view: row pitch 128 B is the
//       swizzle-atom row pitch, a K step of 8 tf32 is +32 B, and K blocks beyond 128 B simply run on into the
//       next row (start address = base + 4 k0, not 1024-B aligned for most steps).
//   B (N = 32, K-major, SWIZZLE_32B): banded Toeplitz taps, column order reversed so that the 8x8 blocks of
//       consecutive K steps alias each other with a POSITIVE stride: block (step s, group g') = atom s + g'.
//   D[r][c'] = sum_k A[r][k] B[k][c'] = out[32 r + 31 - c'],  out[t] = sum_i q[i] b[t + i].
// Prints the max error against a CPU reference on TF32-truncated inputs for base_offset = 0 and
// base_offset = (start >> 7) & 7.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

constexpr int W = 172, P = 32, M = 128, KS = (P + W + 7) / 8, NATOM = KS + 3;
constexpr int ROWS = M + (KS * 8 + 31) / 32 + 1;   // signal rows of 32 floats that the A operand touches

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, int layout, int baseOff) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major), as CUTLASS sets it
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)(baseOff & 7) << 49;
  d |= (uint64_t)layout << 61;            // 2 = SWIZZLE_128B, 6 = SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(128, 1) k_probe(const float *__restrict__ sig, const float *__restrict__ taps,
                                                  float *__restrict__ out, int baseOffMode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);   // ROWS x 128 B, SWIZZLE_128B, 1024-B aligned
  unsigned char *sB = sA + ((ROWS * 128 + 1023) / 1024) * 1024;            // NATOM x 256 B, SWIZZLE_32B
  __shared__ uint32_t tmemBase;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // signal: linear float index L -> row L/32, 16-byte chunk (L%32)/4, swizzled chunk ^ (row & 7)
  for (int L = tid; L < ROWS * 32; L += blockDim.x) {
    const int row = L >> 5, ch = (L & 31) >> 2, w = L & 3;
    *reinterpret_cast<float *>(sA + row * 128 + ((ch ^ (row & 7)) << 4) + w * 4) = sig[L];
  }
  // taps atoms: atom a, row cc' (0..7), k column kk (0..7): q[8a + kk + cc' - 31]; SWIZZLE_32B: chunk ^ (row>>2 & 1)
  for (int e = tid; e < NATOM * 64; e += blockDim.x) {
    const int a = e >> 6, cc = (e >> 3) & 7, kk = e & 7;
    // c' = 8 g' + cc, k = 8 s + kk, tap index = k - c = k - (31 - c') = 8 s + kk - 31 + 8 g' + cc = 8 a + kk + cc - 31
    const int q = 8 * a + kk + cc - 31;
    const float v = (q >= 0 && q < W) ? taps[q] : 0.f;
    const int ch = kk >> 2, w = kk & 3;
    *reinterpret_cast<float *>(sB + a * 256 + cc * 32 + ((ch ^ ((cc >> 2) & 1)) << 4) + w * 4) = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmemBase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmemBase;

  if (tid == 0) {
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int s = 0; s < KS; s++) {
      const uint32_t aAddr = smem_u32(sA) + 32u * s;
      const uint32_t bAddr = smem_u32(sB) + 256u * s;
      const uint64_t da = make_desc(aAddr, 1024, 2, baseOffMode ? (int)((aAddr >> 7) & 7) : 0);
      const uint64_t db = make_desc(bAddr, 256, 6, 0);
      const uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  // wait for the MMAs (bounded spin: a wrong descriptor must not hang the box)
  {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; it++) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)));
    }
    if (!done && tid == 0) printf("probe: MMA completion never arrived\n");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int r = warp * 32 + lane;
  for (int c = 0; c < 32; c++) out[32 * r + 31 - c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

// throughput of back-to-back MMAs of this shape: REPS x KS steps into one accumulator, cycles per MMA
template <int N>
__global__ void __launch_bounds__(128, 1) k_rate(long long *cyclesOut, int reps, int nAcc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);
  unsigned char *sB = sA + ((ROWS * 128 + 1023) / 1024) * 1024;
  __shared__ uint32_t tmemBase;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (ROWS * 128 + NATOM * 256 * 4) / 4; i += blockDim.x) reinterpret_cast<float *>(sA)[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmemBase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmemBase;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da0 = make_desc(smem_u32(sA), 1024, 2, 0), db0 = make_desc(smem_u32(sB), 256, 6, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      const uint32_t d = tmem + (uint32_t)((r % nAcc) * N);
      for (int s = 0; s < KS; s++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
            "l"(da0 + 2u * s), "l"(db0 + 16u * s), "r"(idesc), "r"((uint32_t)(s > 0)));
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t done = 0;
    while (!done)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)));
    const long long t2 = clock64();
    cyclesOut[0] = t1 - t0;   // issue
    cyclesOut[1] = t2 - t0;   // issue + drain
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---- the same question for FP16 operands (kind::f16, K = 16): A = Hankel view with SWIZZLE_64B (row pitch 64 B =
// 32 halves), taps atoms of 8 rows x 16 k (SWIZZLE_32B), atom index a = 2 s + g'
#include <cuda_fp16.h>
constexpr int KS16 = (P + W + 15) / 16, NATOM16 = 2 * (KS16 - 1) + 4;
constexpr int ROWS16 = M + (KS16 * 16 + 31) / 32 + 1;

__global__ void __launch_bounds__(128, 1) k_probe16(const float *__restrict__ sig, const float *__restrict__ taps,
                                                    float *__restrict__ out, long long *cyc, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);   // ROWS16 x 64 B, SWIZZLE_64B
  unsigned char *sB = sA + ((ROWS16 * 64 + 1023) / 1024) * 1024;          // NATOM16 x 256 B, SWIZZLE_32B
  __shared__ uint32_t tmemBase;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // linear half index L -> row L/32 (64 B), 16-byte chunk (L%32)/8, swizzled chunk ^ ((row >> 1) & 3)
  for (int L = tid; L < ROWS16 * 32; L += blockDim.x) {
    const int row = L >> 5, ch = (L & 31) >> 3, w = L & 7;
    *reinterpret_cast<__half *>(sA + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4) + w * 2) = __float2half_rn(sig[L]);
  }
  // atom a, row cc (0..7), k column kk (0..15): tap index 8 a + kk + cc - 31; SWIZZLE_32B: chunk ^ (cc >> 2 & 1)
  for (int e = tid; e < NATOM16 * 128; e += blockDim.x) {
    const int a = e >> 7, cc = (e >> 4) & 7, kk = e & 15;
    const int q = 8 * a + kk + cc - 31;
    const float v = (q >= 0 && q < W) ? taps[q] : 0.f;
    const int ch = kk >> 3, w = kk & 7;
    *reinterpret_cast<__half *>(sB + a * 256 + cc * 32 + ((ch ^ ((cc >> 2) & 1)) << 4) + w * 2) = __float2half_rn(v);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmemBase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmemBase;
  if (tid == 0) {
    // D = F32, A = B = F16 (format 0), K-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da0 = make_desc(smem_u32(sA), 512, 4, 0), db0 = make_desc(smem_u32(sB), 256, 6, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++)
      for (int s = 0; s < KS16; s++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
            "l"(da0 + 2u * s), "l"(db0 + 32u * s), "r"(idesc), "r"((uint32_t)(s > 0)));
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; it++)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)));
    cyc[0] = clock64() - t0;
  } else {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; it++)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)));
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int r = warp * 32 + lane;
  for (int c = 0; c < 32; c++) out[32 * r + 31 - c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

static float tf32_trunc(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b &= ~0x1FFFu;
  memcpy(&x, &b, 4);
  return x;
}

int main() {
  const int nSig = ROWS * 32, nOut = M * P;
  std::vector<float> sig(nSig), taps(W), out(nOut);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 65536.0f; };
  for (auto &x : sig) x = 0.2f + 0.6f * rnd();
  for (auto &x : taps) x = rnd() - 0.5f;
  float *dSig, *dTaps, *dOut;
  cudaMalloc(&dSig, nSig * 4); cudaMalloc(&dTaps, W * 4); cudaMalloc(&dOut, nOut * 4);
  cudaMemcpy(dSig, sig.data(), nSig * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dTaps, taps.data(), W * 4, cudaMemcpyHostToDevice);
  const size_t smem = ((ROWS * 128 + 1023) / 1024) * 1024 + NATOM * 256 + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 2; mode++) {
    cudaMemset(dOut, 0xFF, nOut * 4);
    k_probe<<<1, 128, smem>>>(dSig, dTaps, dOut, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"mode\": %d, \"cuda_error\": \"%s\"}\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dOut, nOut * 4, cudaMemcpyDeviceToHost);
    double maxErr = 0, maxRef = 0;
    int bad = 0, firstBad = -1;
    for (int t = 0; t < nOut; t++) {
      double ref = 0;
      for (int i = 0; i < W; i++) ref += (double)tf32_trunc(taps[i]) * (double)tf32_trunc(sig[t + i]);
      double err = fabs((double)out[t] - ref);
      if (!(err <= 1e-3)) { bad++; if (firstBad < 0) firstBad = t; }
      if (err > maxErr || err != err) maxErr = err;
      if (fabs(ref) > maxRef) maxRef = fabs(ref);
    }
    printf("{\"base_offset_mode\": %d, \"max_abs_err\": %.3e, \"max_abs_ref\": %.3f, \"bad\": %d, \"first_bad\": %d, "
           "\"out0\": %.6f, \"out1\": %.6f, \"out4095\": %.6f}\n",
           mode, maxErr, maxRef, bad, firstBad, out[0], out[1], out[nOut - 1]);
  }
  long long *dCyc, hCyc[2];
  cudaMalloc(&dCyc, 16);
  {
    const size_t smem16 = ((ROWS16 * 64 + 1023) / 1024) * 1024 + NATOM16 * 256 + 2048;
    cudaFuncSetAttribute(k_probe16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16);
    for (int reps = 1; reps <= 400; reps *= 400) {
      cudaMemset(dOut, 0xFF, nOut * 4);
      k_probe16<<<1, 128, smem16>>>(dSig, dTaps, dOut, dCyc, reps);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("{\"f16_error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(out.data(), dOut, nOut * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(hCyc, dCyc, 8, cudaMemcpyDeviceToHost);
      double maxErr = 0;
      for (int t = 0; t < nOut; t++) {
        double ref = 0;
        for (int i = 0; i < W; i++) ref += (double)__half2float(__float2half_rn(taps[i])) * (double)__half2float(__float2half_rn(sig[t + i]));
        double err = fabs((double)out[t] - ref);
        if (err > maxErr || err != err) maxErr = err;
      }
      printf("{\"shape\": \"M128 N32 K16 f16, Hankel A (SWIZZLE_64B)\", \"reps\": %d, \"max_abs_err_vs_fp16_inputs\": %.3e, "
             "\"cycles_per_mma\": %.1f}\n", reps, maxErr, (double)hCyc[0] / ((double)reps * KS16));
    }
  }
  const size_t smemR = ((ROWS * 128 + 1023) / 1024) * 1024 + NATOM * 256 * 4 + 2048;
  cudaFuncSetAttribute(k_rate<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemR);
  cudaFuncSetAttribute(k_rate<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemR);
  cudaFuncSetAttribute(k_rate<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemR);
  const int reps = 400;
  for (int n = 32; n <= 128; n *= 2)
    for (int nAcc = 1; nAcc <= 2; nAcc++) {
      if (n == 32) k_rate<32><<<1, 128, smemR>>>(dCyc, reps, nAcc);
      else if (n == 64) k_rate<64><<<1, 128, smemR>>>(dCyc, reps, nAcc);
      else k_rate<128><<<1, 128, smemR>>>(dCyc, reps, nAcc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("{\"rate_error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hCyc, dCyc, 16, cudaMemcpyDeviceToHost);
      printf("{\"shape\": \"M128 N%d K8 tf32\", \"accumulators\": %d, \"mmas\": %d, \"issue_cycles_per_mma\": %.1f, "
             "\"total_cycles_per_mma\": %.1f}\n", n, nAcc, reps * KS, (double)hCyc[0] / (reps * KS),
             (double)hCyc[1] / (reps * KS));
    }
  return 0;
}
