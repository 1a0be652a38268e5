// Squeeze-and-Excitation fusion (fuse_depth_in_rgb_encoder='SE-add'):
//   SqueezeAndExcitation.forward  model_utils.py:47-51   GAP -> 1x1 (C -> C/16) -> ReLU -> 1x1 -> sigmoid -> x * s
//   SqueezeAndExciteFusionAdd     rgb_depth_fusion.py:22-26   se_rgb(rgb) + se_depth(depth)
// and its gated blend (model_skip_mod_globalgate.py:280-283):
//   fuse = w*rgb + (1-w)*(rgb*s_r + depth*s_d) = rgb*(1 - g + g*s_r) + g*s_d*depth ,  g = 1 - w
// All reductions have a fixed order (per-chunk partial sums, then a sequential finish).
#include "common.cuh"

namespace dynmm {
namespace {

constexpr int kGapChunks = 64;

// partial[n][chunk][c] = sum over the chunk's pixels of x[n][pixel][c]   (threads stride channels)
// lo_off > 0: x holds [hi | lo] halves of fp32-grade values (DYNMM_CONV_SPLIT layout), lo at channel offset lo_off
__global__ void gap_partial_kernel(const __nv_bfloat16* __restrict__ x, long long hw, int c, int ld, int lo_off,
                                   const int32_t* __restrict__ count, float* __restrict__ partial) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  if (count && n >= *count) return;          // gated-off depth slots hold no data
  const long long p0 = hw * chunk / kGapChunks, p1 = hw * (chunk + 1) / kGapChunks;
  for (int c2 = threadIdx.x * 2; c2 < c; c2 += blockDim.x * 2) {
    float a0 = 0.f, a1 = 0.f;
    const __nv_bfloat16* base = x + (static_cast<long long>(n) * hw) * ld + c2;
    for (long long p = p0; p < p1; ++p) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + p * ld));
      if (lo_off > 0) {
        const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(base + p * ld + lo_off));
        a0 += bf16_lo(v) + bf16_lo(q);
        a1 += bf16_hi(v) + bf16_hi(q);
      } else {
        a0 += bf16_lo(v);
        a1 += bf16_hi(v);
      }
    }
    float* o = partial + (static_cast<long long>(n) * kGapChunks + chunk) * c + c2;
    o[0] = a0;
    o[1] = a1;
  }
}

// mean over chunks (fixed order) -> hidden = relu(W1 mean + b1) -> sigma = sigmoid(W2 hidden + b2)
// one block per row; `chunks` partial rows of `c` floats each are summed first.
__global__ void se_mlp_kernel(const float* __restrict__ partial, int chunks, int ld, int c_off, float inv_area,
                              int c, int hidden,
                              const float* __restrict__ w1, const float* __restrict__ b1,
                              const float* __restrict__ w2, const float* __restrict__ b2,
                              const int32_t* __restrict__ count, float* __restrict__ sigma) {
  extern __shared__ float s[];           // [c] mean, [hidden]
  float* s_mean = s;
  float* s_hid = s + c;
  const int n = blockIdx.x;
  if (count && n >= *count) return;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < chunks; ++k) acc += partial[(static_cast<long long>(n) * chunks + k) * ld + c_off + ch];
    s_mean[ch] = acc * inv_area;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
    float acc = b1[j];
    for (int ch = 0; ch < c; ++ch) acc = fmaf(w1[j * c + ch], s_mean[ch], acc);
    s_hid[j] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float acc = b2[ch];
    for (int j = 0; j < hidden; ++j) acc = fmaf(w2[ch * hidden + j], s_hid[j], acc);
    sigma[static_cast<long long>(n) * c + ch] = 1.f / (1.f + expf(-acc));
  }
}

// out[n,p,c] = rgb*(1 - g + g*s_r[n,c]) + g*s_d[slot,c]*depth[slot,p,c]   (8 channels per thread)
__global__ void se_gated_fuse_kernel(const uint4* __restrict__ rgb, const uint4* __restrict__ depth,
                                     const float* __restrict__ sig_r, const float* __restrict__ sig_d,
                                     const float* __restrict__ gate, const int32_t* __restrict__ slot, int n,
                                     long long hw, int c, int out_ld, __nv_bfloat16* __restrict__ out) {
  const int cv = c >> 3;
  const long long total = 1LL * n * hw * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv) * 8;
    const long long pix = i / cv;                 // n*hw + p
    const int s = (int)(pix / hw);
    const long long p = pix - 1LL * s * hw;
    const float g = gate[s];
    const uint4 vr = rgb[i];
    float f[8] = {bf16_lo(vr.x), bf16_hi(vr.x), bf16_lo(vr.y), bf16_hi(vr.y),
                  bf16_lo(vr.z), bf16_hi(vr.z), bf16_lo(vr.w), bf16_hi(vr.w)};
    if (g != 0.f) {       // gated-off samples keep the RGB features and never touch depth / its sigma
      const int ds = slot ? slot[s] : s;
      const uint4 vd = __ldg(&depth[(1LL * ds * hw + p) * cv + (c8 >> 3)]);
      const float d[8] = {bf16_lo(vd.x), bf16_hi(vd.x), bf16_lo(vd.y), bf16_hi(vd.y),
                          bf16_lo(vd.z), bf16_hi(vd.z), bf16_lo(vd.w), bf16_hi(vd.w)};
      const float* sr = sig_r + 1LL * s * c + c8;
      const float* sd = sig_d + 1LL * ds * c + c8;
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] * (1.f - g + g * __ldg(sr + e)) + g * __ldg(sd + e) * d[e];
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + pix * out_ld + c8) = o;
  }
}

// The same blend on [hi | lo] tensors (fp32-grade engine mode): rgb [n, hw, 2c], depth [slots, hw, 2c], out with pitch
// out_ld and its lo half at channel offset out_ld / 2; values are reconstructed as hi + lo, blended in fp32, split again.
__global__ void se_gated_fuse_split_kernel(const __nv_bfloat16* __restrict__ rgb, const __nv_bfloat16* __restrict__ depth,
                                           const float* __restrict__ sig_r, const float* __restrict__ sig_d,
                                           const float* __restrict__ gate, const int32_t* __restrict__ slot, int n,
                                           long long hw, int c, int out_ld, __nv_bfloat16* __restrict__ out) {
  const int cv = c >> 3;
  const long long total = 1LL * n * hw * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv) * 8;
    const long long pix = i / cv;                 // n*hw + p
    const int s = (int)(pix / hw);
    const long long p = pix - 1LL * s * hw;
    const float g = gate[s];
    const uint4 vr = __ldg(reinterpret_cast<const uint4*>(rgb + pix * 2 * c + c8));
    const uint4 qr = __ldg(reinterpret_cast<const uint4*>(rgb + pix * 2 * c + c + c8));
    float f[8] = {bf16_lo(vr.x) + bf16_lo(qr.x), bf16_hi(vr.x) + bf16_hi(qr.x), bf16_lo(vr.y) + bf16_lo(qr.y),
                  bf16_hi(vr.y) + bf16_hi(qr.y), bf16_lo(vr.z) + bf16_lo(qr.z), bf16_hi(vr.z) + bf16_hi(qr.z),
                  bf16_lo(vr.w) + bf16_lo(qr.w), bf16_hi(vr.w) + bf16_hi(qr.w)};
    if (g != 0.f) {
      const int ds = slot ? slot[s] : s;
      const __nv_bfloat16* dp = depth + (1LL * ds * hw + p) * 2 * c + c8;
      const uint4 vd = __ldg(reinterpret_cast<const uint4*>(dp));
      const uint4 qd = __ldg(reinterpret_cast<const uint4*>(dp + c));
      const float d[8] = {bf16_lo(vd.x) + bf16_lo(qd.x), bf16_hi(vd.x) + bf16_hi(qd.x), bf16_lo(vd.y) + bf16_lo(qd.y),
                          bf16_hi(vd.y) + bf16_hi(qd.y), bf16_lo(vd.z) + bf16_lo(qd.z), bf16_hi(vd.z) + bf16_hi(qd.z),
                          bf16_lo(vd.w) + bf16_lo(qd.w), bf16_hi(vd.w) + bf16_hi(qd.w)};
      const float* sr = sig_r + 1LL * s * c + c8;
      const float* sd = sig_d + 1LL * ds * c + c8;
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] * (1.f - g + g * __ldg(sr + e)) + g * __ldg(sd + e) * d[e];
    }
    uint4 o, l;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    l.x = pack_bf16(f[0] - bf16_lo(o.x), f[1] - bf16_hi(o.x));
    l.y = pack_bf16(f[2] - bf16_lo(o.y), f[3] - bf16_hi(o.y));
    l.z = pack_bf16(f[4] - bf16_lo(o.z), f[5] - bf16_hi(o.z));
    l.w = pack_bf16(f[6] - bf16_lo(o.w), f[7] - bf16_hi(o.w));
    *reinterpret_cast<uint4*>(out + pix * out_ld + c8) = o;
    *reinterpret_cast<uint4*>(out + pix * out_ld + (out_ld >> 1) + c8) = l;
  }
}

}  // namespace
}  // namespace dynmm

using namespace dynmm;

extern "C" long long dynmm_gap_workspace(int n, int c) { return 1LL * n * kGapChunks * c * sizeof(float); }

extern "C" int dynmm_gap_partial(const void* x, int n, long long hw, int c, int ld, const int32_t* count,
                                 float* partial, void* stream) {
  DYNMM_CHECK_ARG(x && partial && n >= 1 && hw >= 1 && c >= 2 && c % 2 == 0 && ld >= c && ld % 2 == 0, "gap: bad args");
  gap_partial_kernel<<<dim3(kGapChunks, n), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), hw, c, ld, 0, count, partial);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_gap_partial_split(const void* x, int n, long long hw, int c, int ld, const int32_t* count,
                                       float* partial, void* stream) {
  DYNMM_CHECK_ARG(x && partial && n >= 1 && hw >= 1 && c >= 2 && c % 2 == 0 && ld >= 2 * c && ld % 4 == 0,
                  "gap_split: bad args");
  gap_partial_kernel<<<dim3(kGapChunks, n), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), hw, c, ld, ld / 2, count, partial);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_se_mlp(const float* partial, int rows, int chunks, int ld, int c_off, float inv_area, int c,
                            int hidden, const float* w1, const float* b1, const float* w2, const float* b2,
                            const int32_t* count, float* sigma, void* stream) {
  DYNMM_CHECK_ARG(partial && w1 && b1 && w2 && b2 && sigma && rows >= 1 && chunks >= 1 && c >= 1 && hidden >= 1 && c_off >= 0 &&
                      ld >= c_off + c &&
                      (c + hidden) * 4 <= 48 * 1024,
                  "se_mlp: bad args");
  se_mlp_kernel<<<rows, 256, (c + hidden) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      partial, chunks, ld, c_off, inv_area, c, hidden, w1, b1, w2, b2, count, sigma);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_se_gated_fuse(const void* rgb, const void* depth, const float* sig_r, const float* sig_d,
                                   const float* gate, const int32_t* slot, int n, long long hw, int c, int out_ld,
                                   void* out, void* stream) {
  DYNMM_CHECK_ARG(rgb && depth && sig_r && sig_d && gate && out && n >= 1 && hw >= 1 && c % 8 == 0 && out_ld % 8 == 0 &&
                      out_ld >= c,
                  "se_gated_fuse: bad args");
  const long long total = 1LL * n * hw * (c / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  se_gated_fuse_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(rgb), static_cast<const uint4*>(depth), sig_r, sig_d, gate, slot, n, hw, c, out_ld,
      static_cast<__nv_bfloat16*>(out));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_se_gated_fuse_split(const void* rgb, const void* depth, const float* sig_r, const float* sig_d,
                                         const float* gate, const int32_t* slot, int n, long long hw, int c,
                                         int out_ld, void* out, void* stream) {
  DYNMM_CHECK_ARG(rgb && depth && sig_r && sig_d && gate && out && n >= 1 && hw >= 1 && c % 8 == 0 && out_ld % 16 == 0 &&
                      out_ld >= 2 * c,
                  "se_gated_fuse_split: bad args");
  const long long total = 1LL * n * hw * (c / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  se_gated_fuse_split_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(rgb), static_cast<const __nv_bfloat16*>(depth), sig_r, sig_d, gate, slot, n, hw, c,
      out_ld, static_cast<__nv_bfloat16*>(out));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
