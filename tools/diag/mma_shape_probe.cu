// Probe of tcgen05.mma issue cost on B200 by shape and A-operand source (kind::f16, bf16 operands, K = 16 per instruction):
// what a convolution with few output channels could gain from a different MMA formulation.
//   SS : A and B from shared memory (K-major, 128-byte swizzle) - what conv_tc.cu / block_tc.cu issue today (M = 128 pixels,
//        N = output channels)
//   TS : A from tensor memory, B from shared memory - the "operand swap": weights stationary in TMEM as A (M = output
//        channels, or filter taps x output channels), pixels as a wide B operand
// For every case one thread issues R back-to-back MMAs into one accumulator, commits, and waits for the commit; the
// cycles per instruction are (commit observed - start) / R.  Operand contents are irrelevant (zeros).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_shape_probe.bin mma_shape_probe.cu && ./mma_shape_probe.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t kmajor_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ inline uint32_t idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct Case { int m, n, a_in_tmem; };

__global__ void __launch_bounds__(128) probe(const Case* cases, int n_cases, int reps, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_smem = base;                 // 128 rows x 128 B = 16 KB
  const uint32_t b_smem = base + 16384;         // 256 rows x 128 B = 32 KB
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // zero the TMEM columns used as the A operand (columns 256..271 of every lane)
  {
    const uint32_t t_row = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 256u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(t_row), "r"(0u)
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    uint32_t parity = 0;
    for (int c = 0; c < n_cases; ++c) {
      const Case cs = cases[c];
      const uint32_t idesc = idesc_bf16((uint32_t)cs.m, (uint32_t)cs.n);
      const uint64_t da = kmajor_desc_sw128(a_smem), db = kmajor_desc_sw128(b_smem);
      const uint32_t a_t = tmem + 256u;
      for (int pass = 0; pass < 2; ++pass) {     // pass 0 warms up, pass 1 is timed
        const long long t0 = clock64();
        if (cs.a_in_tmem) {
          for (int i = 0; i < reps; ++i)
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(tmem),
                         "r"(a_t), "l"(db), "r"(idesc), "r"(1u)
                         : "memory");
        } else {
          for (int i = 0; i < reps; ++i)
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(1u)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        uint32_t done = 0;
        while (!done) {
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                       : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
          if (clock64() - t1 > (1ll << 28)) __trap();   // a protocol mistake becomes a launch failure, not a hung GPU
        }
        parity ^= 1u;
        const long long t2 = clock64();
        if (pass == 1) { cyc[2 * c] = t1 - t0; cyc[2 * c + 1] = t2 - t0; }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// Several issuing threads (one per warp), each with its own accumulator and barrier, all SS with M = 128: the per-SM
// instruction rate when the per-thread issue floor is taken out of the picture.
__global__ void __launch_bounds__(128) probe_multi(int n, int issuers, int reps, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar[4];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 16384;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int w = threadIdx.x >> 5;
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && w < issuers) {
      const uint32_t idesc = idesc_bf16(128u, (uint32_t)n);
      // each issuer reads its own 16 KB of A rows would need more shared memory: they share the operand tiles but use
      // different K offsets (32 B apart) and their own accumulator columns
      const uint64_t da = kmajor_desc_sw128(a_smem + (uint32_t)w * 32u), db = kmajor_desc_sw128(b_smem + (uint32_t)w * 32u);
      const uint32_t d = tmem + (uint32_t)(w * 128);
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i)
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d),
                     "l"(da), "l"(db), "r"(idesc), "r"(1u)
                     : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
      const long long t1 = clock64();
      uint32_t done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&bar[w])), "r"((uint32_t)pass) : "memory");
        if (clock64() - t1 > (1ll << 28)) __trap();
      }
      if (pass == 1) cyc[w] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  const Case h_cases[] = {{128, 32, 0}, {128, 64, 0}, {128, 96, 0}, {128, 128, 0}, {128, 256, 0}, {64, 64, 0}, {64, 256, 0},
                          {128, 64, 1}, {128, 128, 1}, {128, 256, 1}, {64, 128, 1}, {64, 256, 1}};
  const int n = sizeof(h_cases) / sizeof(h_cases[0]), reps = 512;
  Case* d_cases;
  long long* d_cyc;
  cudaMalloc(&d_cases, sizeof(h_cases));
  cudaMalloc(&d_cyc, 2 * n * sizeof(long long));
  cudaMemcpy(d_cases, h_cases, sizeof(h_cases), cudaMemcpyHostToDevice);
  const size_t smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(d_cases, n, reps, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[64];
  cudaMemcpy(h, d_cyc, 2 * n * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("tcgen05.mma kind::f16, K = 16, %d back-to-back instructions by one thread (one CTA, otherwise idle SM)\n", reps);
  printf("%-4s %5s %5s %14s %14s %12s %22s\n", "A", "M", "N", "issue cyc/mma", "total cyc/mma", "math floor", "MACs/cycle (of 4096)");
  for (int c = 0; c < n; ++c) {
    const double per = (double)h[2 * c + 1] / reps;
    const double floor_c = (double)(h_cases[c].m > 128 ? h_cases[c].m : 128) * h_cases[c].n / 256.0;   // guide: max(M,128)*N/256
    printf("%-4s %5d %5d %14.1f %14.1f %12.0f %22.0f\n", h_cases[c].a_in_tmem ? "TMEM" : "smem", h_cases[c].m, h_cases[c].n,
           (double)h[2 * c] / reps, per, floor_c, (double)h_cases[c].m * h_cases[c].n * 16 / per);
  }
  // per-SM rate with 1, 2 and 4 issuing threads (N <= 128: four accumulators of 128 columns fit the 512)
  printf("\nseveral issuing threads (one per warp, own accumulator each), A and B from shared memory, M = 128\n");
  printf("%5s %8s %22s %22s\n", "N", "issuers", "cycles/mma per thread", "cycles/mma per SM");
  const int ns[] = {32, 64, 96, 128};
  for (int ni = 0; ni < 4; ++ni)
    for (int issuers = 1; issuers <= 4; issuers *= 2) {
      cudaFuncSetAttribute(probe_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaMemset(d_cyc, 0, 4 * sizeof(long long));
      probe_multi<<<1, 128, smem>>>(ns[ni], issuers, reps, d_cyc);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("probe_multi failed: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_cyc, 4 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < issuers; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%5d %8d %22.1f %22.1f\n", ns[ni], issuers, (double)mx / reps, (double)mx / (reps * issuers));
    }
  return 0;
}
