// Throughput of fp32 reductions into global memory (the wgrad epilogue's flush), three access patterns:
//   rows   : lane l adds 16 B at row (l) * pitch + j * 16   (what wgrad_tc_kernel does: a thread owns an accumulator row)
//   coal   : lane l adds 16 B at base + l * 16               (a warp instruction covers 512 contiguous bytes)
//   bulk   : one thread issues cp.reduce.async.bulk.global.shared::cta.add.f32 of 4 KB contiguous from shared memory
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/red_probe tools/red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// every CTA owns a [128 rows][512 cols] fp32 block (256 KB), as one wgrad work item; 128 threads
__global__ void k_rows(float* out, int items, int pitch) {
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    float* base = out + (size_t)it * 128 * pitch + (size_t)threadIdx.x * pitch;
    for (int c = 0; c < 512; c += 32)
#pragma unroll
      for (int j = 0; j < 32; j += 4) red4(base + c + j, 1.f, 2.f, 3.f, 4.f);
  }
}
__global__ void k_coal(float* out, int items, int pitch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    float* base = out + (size_t)it * 128 * pitch;
    for (int r = warp; r < 128; r += 4)
#pragma unroll
      for (int c = 0; c < 512; c += 128) red4(base + (size_t)r * pitch + c + lane * 4, 1.f, 2.f, 3.f, 4.f);
  }
}
__global__ void k_bulk(float* out, int items, int pitch) {
  extern __shared__ __align__(128) float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.f;       // 32 KB
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      float* base = out + (size_t)it * 128 * pitch;
      for (int r = 0; r < 128; ++r) {          // one 2 KB row (512 floats) per bulk op
        uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + (r & 15) * 512);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(base + (size_t)r * pitch), "r"(s), "r"(2048) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main() {
  const int items = 256, pitch = 512;
  const size_t n = (size_t)items * 128 * pitch;
  float* out; cudaMalloc(&out, n * 4); cudaMemset(out, 0, n * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int mode = 0; mode < 3; ++mode) {
    for (int ctas : {148, 296}) {
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_rows<<<ctas, 128>>>(out, items, pitch);
        else if (mode == 1) k_coal<<<ctas, 128>>>(out, items, pitch);
        else k_bulk<<<ctas, 128, 32768>>>(out, items, pitch);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("%s ctas %d: %.3f ms for %.1f MB = %.0f GB/s (%s)\n", mode == 0 ? "rows" : mode == 1 ? "coal" : "bulk", ctas, best, n * 4 / 1e6,
             n * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  float h[4]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("check %g %g %g %g\n", h[0], h[1], h[2], h[3]);
  return 0;
}
