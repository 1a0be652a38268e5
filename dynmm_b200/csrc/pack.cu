// Engine build without library kernels: eval-mode BatchNorm folding and bf16 operand packing of one convolution
// in ONE launch (FusionEngine.__init__ used ~8 ATen element-wise launches per convolution for the same result).
//
//   scale[o] = bn_w[o] / sqrt(bn_var[o] + eps)
//   shift[o] = bn_b[o] - bn_mean[o] * scale[o]  (+ bias[o] * scale[o])
//   packed[tap][o][c] = bf16_rn(w[o][c][tap] * scale[o])          o < c_out, zero rows up to c_out_pad16
//
// Every product / sum is rounded on its own (__fmul_rn / __fadd_rn / __fsub_rn: no FMA contraction), so the result is
// bit-identical to the fp32 PyTorch expression the oracle evaluates (conv + BatchNorm of the reference,
// FusionDynMM/src/models/resnet.py:124-147, model_utils.py:11-23, folded the same way).
#include "common.cuh"

namespace dynmm {

namespace {

__device__ __forceinline__ float bn_scale(const float* bn_w, const float* bn_var, float eps, int o) {
  return __fdiv_rn(bn_w[o], __fsqrt_rn(__fadd_rn(bn_var[o], eps)));
}

__device__ __forceinline__ float bn_shift(const float* bn_b, const float* bn_mean, const float* bias, float scale, int o) {
  float s = __fsub_rn(bn_b[o], __fmul_rn(bn_mean[o], scale));
  if (bias) s = __fadd_rn(s, __fmul_rn(bias[o], scale));
  return s;
}

__global__ void fold_pack_conv_kernel(const float* __restrict__ w, int co, int ci, int taps, int co_pad,
                                      const float* __restrict__ bias, const float* __restrict__ bn_w,
                                      const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                      const float* __restrict__ bn_var, float eps, __nv_bfloat16* __restrict__ packed,
                                      float* __restrict__ shift) {
  const size_t total = static_cast<size_t>(taps) * co_pad * ci;
  const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (size_t i = tid; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ci);
    const size_t r = i / ci;
    const int o = static_cast<int>(r % co_pad);
    const int t = static_cast<int>(r / co_pad);
    float v = 0.f;
    if (o < co) {
      v = w[(static_cast<size_t>(o) * ci + c) * taps + t];
      if (bn_w) v = __fmul_rn(v, bn_scale(bn_w, bn_var, eps, o));
    }
    packed[i] = __float2bfloat16(v);
  }
  if (shift) {
    for (size_t o = tid; o < static_cast<size_t>(co); o += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const int oi = static_cast<int>(o);
      shift[o] = bn_w ? bn_shift(bn_b, bn_mean, bias, bn_scale(bn_w, bn_var, eps, oi), oi) : (bias ? bias[o] : 0.f);
    }
  }
}

// the same folding with the fp32 result split into bf16 halves: packed[tap][o][3 * ci] = [hi | lo | hi]
__global__ void fold_pack_conv_split_kernel(const float* __restrict__ w, int co, int ci, int taps, int co_pad,
                                            const float* __restrict__ bias, const float* __restrict__ bn_w,
                                            const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                            const float* __restrict__ bn_var, float eps, __nv_bfloat16* __restrict__ packed,
                                            float* __restrict__ shift) {
  const size_t total = static_cast<size_t>(taps) * co_pad * ci;
  const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (size_t i = tid; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ci);
    const size_t r = i / ci;
    const int o = static_cast<int>(r % co_pad);
    const int t = static_cast<int>(r / co_pad);
    float v = 0.f;
    if (o < co) {
      v = w[(static_cast<size_t>(o) * ci + c) * taps + t];
      if (bn_w) v = __fmul_rn(v, bn_scale(bn_w, bn_var, eps, o));
    }
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(__fsub_rn(v, __bfloat162float(hi)));
    __nv_bfloat16* row = packed + r * (3 * static_cast<size_t>(ci));
    row[c] = hi;
    row[ci + c] = lo;
    row[2 * ci + c] = hi;
  }
  if (shift) {
    for (size_t o = tid; o < static_cast<size_t>(co); o += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const int oi = static_cast<int>(o);
      shift[o] = bn_w ? bn_shift(bn_b, bn_mean, bias, bn_scale(bn_w, bn_var, eps, oi), oi) : (bias ? bias[o] : 0.f);
    }
  }
}

__global__ void fold_bn_kernel(int c, const float* __restrict__ bias, const float* __restrict__ bn_w,
                               const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                               const float* __restrict__ bn_var, float eps, float* __restrict__ scale,
                               float* __restrict__ shift) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= c) return;
  const float s = bn_scale(bn_w, bn_var, eps, o);
  scale[o] = s;
  shift[o] = bn_shift(bn_b, bn_mean, bias, s, o);
}

// out = in.permute(p0, p1, p2).contiguous() for a 3-D fp32 tensor [d0][d1][d2] (stem / gate / depthwise-stencil weight
// layouts: [o][c][tap] -> [tap][c][o] or [o][tap][c])
__global__ void permute3d_f32_kernel(const float* __restrict__ in, int d0, int d1, int d2, int p0, int p1, int p2,
                                     float* __restrict__ out) {
  const int d[3] = {d0, d1, d2};
  const long long st[3] = {1LL * d1 * d2, d2, 1};
  const int e0 = d[p0], e1 = d[p1], e2 = d[p2];
  const long long total = 1LL * d0 * d1 * d2;
  for (long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < total;
       k += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i2 = static_cast<int>(k % e2);
    const long long r = k / e2;
    const int i1 = static_cast<int>(r % e1);
    const int i0 = static_cast<int>(r / e1);
    (void)e0;
    out[k] = in[i0 * st[p0] + i1 * st[p1] + i2 * st[p2]];
  }
}

}  // namespace

}  // namespace dynmm

using namespace dynmm;

extern "C" int dynmm_fold_pack_conv(const float* w, int c_out, int c_in, int kh, int kw, const float* bias,
                                    const float* bn_weight, const float* bn_bias, const float* bn_mean,
                                    const float* bn_var, float eps, void* packed, float* shift, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(w && packed, "fold_pack_conv: null pointer");
  DYNMM_CHECK_ARG(c_out >= 1 && c_in >= 1 && kh >= 1 && kw >= 1, "fold_pack_conv: bad shape");
  const bool bn = bn_weight != nullptr;
  DYNMM_CHECK_ARG(!bn || (bn_bias && bn_mean && bn_var), "fold_pack_conv: BatchNorm needs weight, bias, mean and var");
  DYNMM_CHECK_ARG(bn || !(bn_bias || bn_mean || bn_var), "fold_pack_conv: BatchNorm needs weight, bias, mean and var");
  DYNMM_CHECK_ARG(shift || !(bn || bias), "fold_pack_conv: a bias / BatchNorm needs the shift output");
  const int co_pad = (c_out + 15) / 16 * 16;
  const long long total = 1LL * kh * kw * co_pad * c_in;
  long long blocks = (total + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  fold_pack_conv_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(w, c_out, c_in, kh * kw, co_pad, bias, bn_weight, bn_bias,
                                                                    bn_mean, bn_var, eps,
                                                                    static_cast<__nv_bfloat16*>(packed), shift);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_fold_pack_conv_split(const float* w, int c_out, int c_in, int kh, int kw, const float* bias,
                                          const float* bn_weight, const float* bn_bias, const float* bn_mean,
                                          const float* bn_var, float eps, void* packed, float* shift, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(w && packed, "fold_pack_conv_split: null pointer");
  DYNMM_CHECK_ARG(c_out >= 1 && c_in >= 1 && kh >= 1 && kw >= 1, "fold_pack_conv_split: bad shape");
  const bool bn = bn_weight != nullptr;
  DYNMM_CHECK_ARG(!bn || (bn_bias && bn_mean && bn_var), "fold_pack_conv_split: BatchNorm needs weight, bias, mean and var");
  DYNMM_CHECK_ARG(shift || !(bn || bias), "fold_pack_conv_split: a bias / BatchNorm needs the shift output");
  const int co_pad = (c_out + 15) / 16 * 16;
  const long long total = 1LL * kh * kw * co_pad * c_in;
  long long blocks = (total + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  fold_pack_conv_split_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(w, c_out, c_in, kh * kw, co_pad, bias, bn_weight,
                                                                          bn_bias, bn_mean, bn_var, eps,
                                                                          static_cast<__nv_bfloat16*>(packed), shift);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_fold_bn(int c, const float* bias, const float* bn_weight, const float* bn_bias, const float* bn_mean,
                             const float* bn_var, float eps, float* scale, float* shift, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(c >= 1 && bn_weight && bn_bias && bn_mean && bn_var && scale && shift, "fold_bn: null pointer");
  fold_bn_kernel<<<(c + 127) / 128, 128, 0, stream>>>(c, bias, bn_weight, bn_bias, bn_mean, bn_var, eps, scale, shift);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_permute3d_f32(const float* in, int d0, int d1, int d2, int p0, int p1, int p2, float* out,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(in && out && d0 >= 1 && d1 >= 1 && d2 >= 1, "permute3d_f32: bad arguments");
  DYNMM_CHECK_ARG(p0 >= 0 && p0 < 3 && p1 >= 0 && p1 < 3 && p2 >= 0 && p2 < 3 && p0 != p1 && p0 != p2 && p1 != p2,
                  "permute3d_f32: (p0, p1, p2) must be a permutation of (0, 1, 2)");
  long long blocks = (1LL * d0 * d1 * d2 + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  permute3d_f32_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, d0, d1, d2, p0, p1, p2, out);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
