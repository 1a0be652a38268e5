// Micro-benchmark: how fast can ONE thread issue tcgen05.mma (kind::f16, M=128, K=16) for N = 32..256?
// One CTA per SM; operands are zero-filled shared-memory tiles (values do not matter); the loop mirrors the
// conv kernel's inner loop (4 k-steps per descriptor pair, then a commit).  Prints cycles per UMMA and the
// implied fraction of the dense bf16 peak (4096 MAC/clk/SM nominal).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dynmm_b200/csrc -o tools/bin/umma_issue_bench tools/umma_issue_bench.cu
#include <cstdio>
#include <vector>
#include "common.cuh"
namespace dynmm { void set_error(const char*, ...) {} int num_sms() { return 148; } }
using namespace dynmm;

template <bool kElect>
__global__ void __launch_bounds__(128, 1) bench_kernel(int n, int iters, int commit_every, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // kElect = false: one thread runs the whole loop (`if (lane == 0)`), the pattern of conv_igemm.cu before this
  // measurement; kElect = true: the whole warp runs the loop converged and only the tcgen05 instructions sit under
  // elect.sync (the CUTLASS pattern) -- the compiler then needs no ELECT/BRA.U.ANY loop around every UTCHMMA.
  if (kElect ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    uint32_t phase = 0;
    long long t0 = clock64();
    int since = 0;
    for (int it = 0; it < iters; ++it) {
      const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
      if (!kElect || elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + (it & 1) * 256, da + k * 2, db + k * 2, idesc, 1);
      }
      if (kElect) __syncwarp();
      if (++since == commit_every) {
        if (!kElect || elect_one()) umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
        since = 0;
      }
    }
    if (!kElect || elect_one()) umma_commit(&bar);
    mbar_wait(&bar, phase);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(bench_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int elect = 0; elect < 2; ++elect)
  for (int commit_every : {iters + 1, 8, 1}) {
    for (int n : {32, 64, 128, 256}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (elect) bench_kernel<true><<<148, 128, 16384 + 32768 + 2048>>>(n, iters, commit_every, d);
        else bench_kernel<false><<<148, 128, 16384 + 32768 + 2048>>>(n, iters, commit_every, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (long long v : h) avg += v;
      avg /= 148;
      const double per = avg / (iters * 4.0);
      printf("%s N=%3d commit every %4d k-chunks: %.1f cycles per UMMA (128x%dx16), %.0f MAC/clk/SM = %.0f%% of 4096\n", elect ? "elect.sync" : "lane0     ", n,
             commit_every, per, n, 128.0 * n * 16 / per, 100.0 * 128.0 * n * 16 / per / 4096);
    }
  }
  return 0;
}
