// Probe of tcgen05.shift on B200: what moves where, which columns one instruction covers, what it costs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_shift_probe.bin tmem_shift_probe.cu && ./tmem_shift_probe.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(float* out, long long* cyc, int n_shift_chunks, int first_col) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  const uint32_t t_row = base + ((uint32_t)(warp * 32) << 16);
  // write value row*1000 + col into 64 columns, 16 at a time
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint((float)(row * 1000 + c0 + i));
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(t_row + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < n_shift_chunks; ++i)
      asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(base + (uint32_t)(first_col + 8 * i)) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    const long long t2 = clock64();
    cyc[0] = t1 - t0; cyc[1] = t2 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(t_row + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[row * 64 + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 128 * 64 * 4); cudaMalloc(&cyc, 16);
  static float h[128 * 64]; long long hc[2];
  const int cfgs[3][2] = {{1, 8}, {4, 32}, {8, 0}};
  for (auto& c : cfgs) {
    probe<<<1, 128>>>(out, cyc, c[0], c[1]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    printf("shift x%d from column %d: issue %lld cycles, until commit observed %lld cycles\n", c[0], c[1], hc[0], hc[1]);
    const int rows[] = {0, 1, 2, 30, 31, 32, 33, 63, 64, 65, 126, 127};
    for (int r : rows) {
      printf("  row %3d:", r);
      for (int col : {0, 7, 8, 15, 16, 31, 32, 40, 63}) printf(" c%-2d=%8.0f", col, h[r * 64 + col]);
      printf("\n");
    }
  }
  return 0;
}
