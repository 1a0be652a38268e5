// Small training-path kernels around the tensor-core convolutions:
//   * weight packing: fp32 master weight -> the bf16 operand layouts of the forward and of the
//     data-gradient convolution, in one pass (once per optimizer step and layer);
//   * per-channel sums of an NHWC bf16 tensor (the bias gradient), deterministic two-stage.
#include "common.cuh"

namespace dynmm {

namespace {

// w [co][ci][kh][kw] fp32  ->  fwd [tap][co_pad][ci] bf16,  dgr [tap'][ci_pad][co] bf16 with tap' mirrored
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int co, int ci, int kh, int kw, int co_pad,
                                        int ci_pad, __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgr) {
  const int taps = kh * kw;
  const size_t n_fwd = fwd ? static_cast<size_t>(taps) * co_pad * ci : 0;
  const size_t n_dgr = dgr ? static_cast<size_t>(taps) * ci_pad * co : 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_fwd + n_dgr;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (i < n_fwd) {
      const int c = static_cast<int>(i % ci);
      const size_t r = i / ci;
      const int o = static_cast<int>(r % co_pad);
      const int t = static_cast<int>(r / co_pad);
      fwd[i] = __float2bfloat16(o < co ? w[(static_cast<size_t>(o) * ci + c) * taps + t] : 0.f);
    } else {
      const size_t j = i - n_fwd;
      const int o = static_cast<int>(j % co);
      const size_t r = j / co;
      const int c = static_cast<int>(r % ci_pad);
      const int t = static_cast<int>(r / ci_pad);
      dgr[j] = __float2bfloat16(c < ci ? w[(static_cast<size_t>(o) * ci + c) * taps + (taps - 1 - t)] : 0.f);
    }
  }
}

constexpr int kSumThreads = 256;

// stage 1: block b sums rows [b*rows_per, ...) of x [rows][ld] (first c channels, c % 8 == 0)
__global__ void __launch_bounds__(kSumThreads)
channel_sum_partial_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int c, int ld, long long rows_per,
                           float* __restrict__ partial) {
  __shared__ float red[kSumThreads][9];
  const int groups = c >> 3;                       // 8-channel groups per row
  const int lanes = kSumThreads / groups;          // rows processed concurrently (groups <= 256)
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long r0 = blockIdx.x * rows_per;
  const long long r1 = min(r0 + rows_per, rows);
  if (rl < lanes) {
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld + g * 8));
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < groups) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < lanes; ++l) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += red[l * groups + threadIdx.x][e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) partial[static_cast<size_t>(blockIdx.x) * c + threadIdx.x * 8 + e] = s[e];
  }
}

__global__ void channel_sum_final_kernel(const float* __restrict__ partial, int blocks, int c, float* __restrict__ out,
                                         int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[static_cast<size_t>(b) * c + ch];
  out[ch] = accumulate ? out[ch] + s : s;
}

int sum_blocks(long long rows) {
  long long b = (rows + 255) / 256;                // at least 256 rows per block
  const long long cap = 4LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

}  // namespace dynmm

using namespace dynmm;

extern "C" int dynmm_pack_conv_weight(const float* w, int c_out, int c_in, int kh, int kw, void* fwd, void* dgrad,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(w && (fwd || dgrad), "pack_conv_weight: null pointer");
  DYNMM_CHECK_ARG(c_out >= 1 && c_in >= 1 && kh >= 1 && kw >= 1, "pack_conv_weight: bad shape");
  const int co_pad = (c_out + 15) / 16 * 16, ci_pad = (c_in + 15) / 16 * 16;
  const long long n = 1LL * kh * kw * ((fwd ? 1LL * co_pad * c_in : 0) + (dgrad ? 1LL * ci_pad * c_out : 0));
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  pack_conv_weight_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      w, c_out, c_in, kh, kw, co_pad, ci_pad, static_cast<__nv_bfloat16*>(fwd), static_cast<__nv_bfloat16*>(dgrad));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" long long dynmm_channel_sum_workspace(long long rows, int c) {
  if (rows < 1 || c < 8 || c % 8 || c > 2048) return -1;
  return 4LL * sum_blocks(rows) * c;
}

extern "C" int dynmm_channel_sum(const void* x, long long rows, int c, int ld, float* out, void* workspace,
                                 long long workspace_bytes, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(x && out && workspace, "channel_sum: null pointer");
  DYNMM_CHECK_ARG(rows >= 1 && c >= 8 && c % 8 == 0 && c <= 2048 && ld % 8 == 0 && ld >= c, "channel_sum: c/ld %% 8");
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "channel_sum: x must be 16-byte aligned");
  const int blocks = sum_blocks(rows);
  DYNMM_CHECK_ARG(workspace_bytes >= 4LL * blocks * c, "channel_sum: workspace of %lld bytes needed", 4LL * blocks * c);
  const long long rows_per = (rows + blocks - 1) / blocks;
  float* partial = static_cast<float*>(workspace);
  channel_sum_partial_kernel<<<blocks, kSumThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), rows, c, ld,
                                                                 rows_per, partial);
  DYNMM_LAUNCH_CHECK();
  channel_sum_final_kernel<<<(c + 127) / 128, 128, 0, stream>>>(partial, blocks, c, out, accumulate);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
