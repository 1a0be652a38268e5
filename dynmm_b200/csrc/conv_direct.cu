// One-thread-per-output direct convolution with the same contract as
// dynmm_conv_igemm_fwd.  TEST COMPARATOR ONLY (independent arithmetic path on
// CUDA cores, fp32 accumulate) -- the product path never calls it.
#include "common.cuh"

namespace dynmm {
namespace {

__global__ void conv_direct_kernel(dynmm_conv_params p) {
  const int active = p.count ? min(*p.count, p.n) : p.n;
  const long long total = 1LL * active * p.h_out * p.w_out * p.c_out;
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(p.in);
  const __nv_bfloat16* wgt = static_cast<const __nv_bfloat16*>(p.weight);
  const __nv_bfloat16* res = static_cast<const __nv_bfloat16*>(p.residual);
  const __nv_bfloat16* gated = static_cast<const __nv_bfloat16*>(p.gated);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
  const int c_out_pad = (p.c_out + 15) / 16 * 16;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = i % p.c_out;
    long long r = i / p.c_out;
    const int w = r % p.w_out;
    r /= p.w_out;
    const int h = r % p.h_out;
    const int n = r / p.h_out;
    const int n_in = p.in_map ? p.in_map[n] : n;
    float acc = 0.f;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int y = h * p.stride_h + ky - p.pad_h;
      if (y < 0 || y >= p.h_in) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int x = w * p.stride_w + kx - p.pad_w;
        if (x < 0 || x >= p.w_in) continue;
        const __nv_bfloat16* ip = in + ((1LL * n_in * p.h_in + y) * p.w_in + x) * p.in_ld;
        const __nv_bfloat16* wp = wgt + (1LL * (ky * p.kw + kx) * c_out_pad + c) * p.c_in;
        for (int k = 0; k < p.c_in; ++k) acc += __bfloat162float(ip[k]) * __bfloat162float(wp[k]);
      }
    }
    float v = acc;
    if (p.scale) v *= p.scale[c];
    if (p.shift) v += p.shift[c];
    const long long pix = (1LL * n * p.h_out + h) * p.w_out + w;
    if (res) {
      const int rn = p.res_map ? p.res_map[n] : n;
      v += __bfloat162float(res[((1LL * rn * p.h_out + h) * p.w_out + w) * p.res_ld + c]);
    }
    if (p.relu == 1) v = fmaxf(v, 0.f);
    else if (p.relu == 2) v = v / (1.f + __expf(-v));
    else if (p.relu == 3) v = v * fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
    if (gated) {
      const float g = p.gate[n];
      if (g != 0.f) {
        const int slot = p.gated_slot ? p.gated_slot[n] : n;
        v += g * __bfloat162float(gated[((1LL * slot * p.h_out + h) * p.w_out + w) * p.gated_ld + c]);
      }
    }
    out[pix * p.out_ld + c] = __float2bfloat16_rn(v);
  }
}

}  // namespace
}  // namespace dynmm

extern "C" int dynmm_conv_direct_fwd(const dynmm_conv_params* p, void* stream) {
  DYNMM_CHECK_ARG(p && p->in && p->weight && p->out, "conv_direct: null pointer");
  const long long total = 1LL * p->n * p->h_out * p->w_out * p->c_out;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  dynmm::conv_direct_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
