// Probe (GPU dev tool, not product): per-SM TMA throughput as a function of box size and boxes per stage.
// Every CTA (one per SM) streams `iters` stages; a stage is `nb` 2-D boxes of (64 bf16 x R rows) = R x 128 B
// (SWIZZLE_128B, rows 512 B apart in global memory like a 256-channel channels-last tensor) out of an
// L2-resident region.  A consumer thread frees each stage as soon as it lands (no MMA), so the measured
// clocks per stage are the TMA/L2 service time alone.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/tma_probe tools/tma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mra_gan_b200/csrc/conv_tc.cuh"

namespace mra { thread_local std::string g_last_error; std::atomic<long long> g_launch_count{0}; }
using namespace mra;
using namespace mra::tc;

constexpr int kStages = 4;

__global__ void __launch_bounds__(64, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int R, int nb, int iters, int total_rows, long long* clocks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = (uint32_t)R * 128u * nb;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
  uint64_t* empty_bar = full_bar + kStages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    int row = (blockIdx.x * 997) % (total_rows - R);
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u, nullptr, 0);
      mbar_expect_tx(&full_bar[s], stage_bytes);
      for (int b = 0; b < nb; ++b) {
        tma_load_2d(smem + (size_t)s * stage_bytes + (size_t)b * R * 128, &tm, &full_bar[s], (b & 3) * 64, row);
        row += R;
        if (row + R > total_rows) row = 0;
      }
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      mbar_wait(&full_bar[s], ph, nullptr, 0);
      mbar_arrive(&empty_bar[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
}

int main() {
  const int C = 256, rows = 65536;                 // 32 MB region (L2 resident)
  bf16* d;
  cudaMalloc(&d, (size_t)rows * C * 2);
  cudaMemset(d, 0, (size_t)rows * C * 2);
  long long* dclk;
  cudaMalloc(&dclk, 148 * sizeof(long long));
  cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int iters = 2000;
  printf("%6s %4s %10s %12s %12s %12s\n", "rows", "nb", "stage KB", "clk/stage", "clk/box", "chip TB/s");
  for (int R : {32, 64, 128, 256})
    for (int nb : {1, 2, 3, 6}) {
      const size_t stage = (size_t)R * 128 * nb;
      if (stage * kStages + 2048 > 227 * 1024) continue;
      CUtensorMap tm;
      if (make_weight_map(&tm, d, rows, C, R)) { printf("tmap failed %s\n", g_last_error.c_str()); return 1; }
      for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        tma_probe_kernel<<<148, 64, stage * kStages + 2048, 0>>>(tm, R, nb, iters, rows, dclk);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 2; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 0) continue;
        std::vector<long long> clk(148);
        cudaMemcpy(clk.data(), dclk, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (long long c : clk) avg += (double)c / 148;
        printf("%6d %4d %10.1f %12.1f %12.1f %12.2f\n", R, nb, stage / 1024.0, avg / iters, avg / iters / nb,
               148.0 * iters * stage / (ms * 1e-3) / 1e12);
      }
    }
  return 0;
}
