// Micro-benchmark: how many bytes per clock can one SM pull from L2 through TMA, as a function of
//   * how many SMs pull at the same time (grid = 18 .. 148 CTAs, one per SM),
//   * whether every CTA reads the SAME tiles (conv weights) or its own (activations),
//   * whether the tiles are multicast inside a thread-block cluster of 2 / 4 / 8 CTAs (each CTA loads 1/csz of a
//     tile and the hardware delivers it to every CTA of the cluster).
// The conv kernels of stages 3-4 are bound by exactly this (DESIGN.md section 5b item f): a 128x128 tile with
// K = 768 streams 272 KB of operands for 3072 cycles of UMMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dynmm_b200/csrc -o tools/bin/tma_bw_bench tools/tma_bw_bench.cu
#include <cstdio>
#include <vector>
#include "tma_host.cuh"
namespace dynmm { void set_error(const char* fmt, ...) { fprintf(stderr, "%s\n", fmt); } int num_sms() { return 148; } }
using namespace dynmm;

constexpr int kStages = 8;
constexpr int kTileBytes = 16384;            // [128 rows][64 bf16]

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void remote_arrive(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;\n" ::
          "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
                   "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// csz = cluster size (1: plain loads).  rows_total: rows of the source region one cluster cycles through.
template <int kCsz>
__global__ void __launch_bounds__(64, 1)
bw_kernel(const __grid_constant__ CUtensorMap map, int iters, int rows_region, int shared_src, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kStages], empty[kStages];
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = kCsz > 1 ? cluster_rank() : 0;
  const int cluster_id = blockIdx.x / kCsz;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kCsz); }
    fence_mbar_init();
  }
  __syncthreads();
  if (kCsz > 1) cluster_sync_all();
  const int rows_per = 128 / kCsz;
  const int base_row = shared_src ? 0 : cluster_id * rows_region;
  const int tiles_region = rows_region / 128;
  if (warp == 0) {
    long long t0 = clock64();
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[stage], kTileBytes);
        const int row = base_row + (it % tiles_region) * 128 + rank * rows_per;
        uint8_t* dst = smem + stage * kTileBytes + rank * rows_per * 128;
        if (kCsz > 1) tma_load_2d_mc(dst, &map, &full[stage], 0, row, (uint16_t)((1u << kCsz) - 1));
        else tma_load_2d(dst, &map, &full[stage], 0, row);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    // drain: wait until the consumers released every stage
    for (int s = 0; s < kStages; ++s) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[stage], phase);
      if (kCsz > 1) {
        for (uint32_t c = 0; c < (uint32_t)kCsz; ++c) remote_arrive(&empty[stage], c);
      } else {
        mbar_arrive(&empty[stage]);
      }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (kCsz > 1) cluster_sync_all();
}

template <int kCsz>
void run(const void* src, size_t src_rows, int grid, int shared_src, int rows_region, long long* d_cycles, const char* label) {
  CUtensorMap map;
  const uint64_t dims[2] = {64, src_rows};
  const uint64_t strides[1] = {128};
  const uint32_t box[2] = {64, (uint32_t)(128 / kCsz)};
  if (encode_map(&map, src, 2, dims, strides, box)) { printf("encode failed\n"); return; }
  const int smem = kStages * kTileBytes + 2048;
  cudaFuncSetAttribute(bw_kernel<kCsz>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bw_kernel<kCsz>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 4000;
  grid = grid / kCsz * kCsz;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = 200 * 1024;       // one CTA per SM, like the conv kernels
  (void)smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCsz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t le = cudaLaunchKernelEx(&cfg, bw_kernel<kCsz>, map, iters, rows_region, shared_src, d_cycles);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (le != cudaSuccess || e != cudaSuccess) { printf("%s grid %d cluster %d: error %s / %s\n", label, grid, kCsz, cudaGetErrorString(le), cudaGetErrorString(e)); return; }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; long long mx = 0;
  for (long long v : h) { avg += v; if (v > mx) mx = v; }
  avg /= grid;
  const double bytes = (double)iters * kTileBytes;
  printf("%-22s cluster %d grid %3d: %6.1f B/clk/SM received (slowest CTA %6.1f), chip %7.0f B/clk, %6.2f TB/s received\n", label,
         kCsz, grid, bytes / avg, bytes / mx, bytes * grid / avg, bytes * grid / (ms * 1e-3) / 1e12);
}

int main() {
  const size_t rows = 148ull * 3072;             // 148 regions of 384 KB (= 3 taps x 128 ch x 512 ... a weight slice)
  void* src;
  cudaMalloc(&src, rows * 128);
  cudaMemset(src, 0, rows * 128);
  long long* d_cycles;
  cudaMalloc(&d_cycles, 256 * sizeof(long long));
  for (int shared_src = 1; shared_src >= 0; --shared_src) {
    const char* label = shared_src ? "same 384 KB (weights)" : "own 384 KB per cluster";
    for (int grid : {16, 32, 72, 112, 144, 148}) {
      run<1>(src, rows, grid, shared_src, 3072, d_cycles, label);
      if (grid % 2 == 0) run<2>(src, rows, grid, shared_src, 3072, d_cycles, label);
      if (grid % 4 == 0 && grid <= 144) run<4>(src, rows, grid, shared_src, 3072, d_cycles, label);
      if (grid % 8 == 0 && grid <= 128) run<8>(src, rows, grid, shared_src, 3072, d_cycles, label);
    }
  }
  // a larger per-CTA footprint (activations that do not stay hot): 2 MB per cluster
  return 0;
}
