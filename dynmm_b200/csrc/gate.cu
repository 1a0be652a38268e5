// Global gate (model_skip_mod_globalgate.py:375-394), DiffSoftmax (:20-30) and the
// device-side plan that turns gate weights into skip lists for the encoder kernels.
//
// All fp32 with a FIXED summation order (no floating-point atomics): the hard
// decision is an argmax of these numbers and has to be reproducible and equal
// to the reference's wherever the top-2 margin exceeds rounding noise.
#include "common.cuh"

namespace dynmm {
namespace {

constexpr int kGateC = 8;       // hidden channels
constexpr int kTaps = 25;       // 5x5
constexpr int kPx = 4;          // output pixels per warp (along W) in conv1

// ---------------------------------------------------------------- conv1: 128 -> 8, 5x5 s2, +BN +tanh
// A warp owns 4 adjacent outputs; lane l owns input channels 4l..4l+3 (lanes 0-15
// read the rgb map, 16-31 the depth map: the torch.concat of :389 never exists).
__global__ void __launch_bounds__(256, 2)
gate_conv1_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int b, int h, int w, int h1, int w1,
                  const float* __restrict__ wgt, const float* __restrict__ scale, const float* __restrict__ shift,
                  float* __restrict__ out) {
  extern __shared__ __align__(16) float s_w[];   // [8][25][128]
  for (int i = threadIdx.x; i < kGateC * kTaps * 128; i += blockDim.x) s_w[i] = wgt[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gx = (w1 + kPx - 1) / kPx;
  const long long items = 1LL * b * h1 * gx;
  const float* src = lane < 16 ? rgb : depth;
  const int coff = (lane & 15) * 4;
  for (long long item = blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < items;
       item += 1LL * gridDim.x * warps_per_block) {
    const int xg = item % gx;
    const int oy = (item / gx) % h1;
    const int n = item / (1LL * gx * h1);
    const int ox0 = xg * kPx;
    float acc[kPx * kGateC];
#pragma unroll
    for (int i = 0; i < kPx * kGateC; ++i) acc[i] = 0.f;
#pragma unroll 1
    for (int ky = 0; ky < 5; ++ky) {
      const int y = 2 * oy + ky;
      float4 in[2 * (kPx - 1) + 5];
#pragma unroll
      for (int j = 0; j < 2 * (kPx - 1) + 5; ++j) {
        const int x = 2 * ox0 + j;
        in[j] = x < w ? __ldg(reinterpret_cast<const float4*>(src + ((1LL * n * h + y) * w + x) * 64 + coff))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
        for (int co = 0; co < kGateC; ++co) {
          const float4 wv = *reinterpret_cast<const float4*>(&s_w[((co * 5 + ky) * 5 + kx) * 128 + lane * 4]);
#pragma unroll
          for (int p = 0; p < kPx; ++p) {
            const float4 x = in[2 * p + kx];
            float a = acc[p * kGateC + co];
            a = fmaf(x.x, wv.x, a);
            a = fmaf(x.y, wv.y, a);
            a = fmaf(x.z, wv.z, a);
            a = fmaf(x.w, wv.w, a);
            acc[p * kGateC + co] = a;
          }
        }
      }
    }
    // transposing butterfly: lane l ends with the warp total of acc[l]
#pragma unroll
    for (int s = 16, cnt = 16; s >= 1; s >>= 1, cnt >>= 1) {
      const bool upper = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < cnt; ++i) {
        const float send = upper ? acc[i] : acc[i + cnt];
        const float keep = upper ? acc[i + cnt] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    const int p = lane >> 3, co = lane & 7;
    if (ox0 + p < w1) {
      const float v = tanhf(fmaf(acc[0], scale[co], shift[co]));
      out[((1LL * n * h1 + oy) * w1 + ox0 + p) * kGateC + co] = v;
    }
  }
}

// ---------------------------------------------------------------- conv2: 8 -> 8, 5x5 s2, +BN +tanh, + partial GAP
// one thread per output pixel, fixed-order block reduction, per-block partial sums
__global__ void __launch_bounds__(128)
gate_conv2_gap_kernel(const float* __restrict__ in, int h1, int w1, int h2, int w2, const float* __restrict__ wgt,
                      const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ partial) {
  __shared__ __align__(16) float s_w[kGateC * kTaps * kGateC];   // [co][ky][kx][ci]
  __shared__ float s_red[4][kGateC];
  for (int i = threadIdx.x; i < kGateC * kTaps * kGateC; i += blockDim.x) s_w[i] = wgt[i];
  __syncthreads();
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  float v[kGateC];
#pragma unroll
  for (int co = 0; co < kGateC; ++co) v[co] = 0.f;
  if (pix < h2 * w2) {
    const int oy = pix / w2, ox = pix % w2;
    float acc[kGateC];
#pragma unroll
    for (int co = 0; co < kGateC; ++co) acc[co] = 0.f;
    for (int ky = 0; ky < 5; ++ky) {
      for (int kx = 0; kx < 5; ++kx) {
        const float* ip = in + ((1LL * n * h1 + 2 * oy + ky) * w1 + 2 * ox + kx) * kGateC;
        const float4 a = __ldg(reinterpret_cast<const float4*>(ip));
        const float4 c = __ldg(reinterpret_cast<const float4*>(ip + 4));
        const float x[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int co = 0; co < kGateC; ++co) {
          const float* wp = &s_w[((co * 5 + ky) * 5 + kx) * kGateC];
#pragma unroll
          for (int ci = 0; ci < kGateC; ++ci) acc[co] = fmaf(x[ci], wp[ci], acc[co]);
        }
      }
    }
#pragma unroll
    for (int co = 0; co < kGateC; ++co) v[co] = tanhf(fmaf(acc[co], scale[co], shift[co]));
  }
  // deterministic reduction: xor-butterfly inside the warp, then warps in index order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int co = 0; co < kGateC; ++co) {
    float s = v[co];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_red[warp][co] = s;
  }
  __syncthreads();
  if (threadIdx.x < kGateC) {
    float s = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += s_red[wv][threadIdx.x];
    partial[(1LL * n * gridDim.x + blockIdx.x) * kGateC + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------- GAP finish + fc (8 -> 5, no bias)
__global__ void gate_head_kernel(const float* __restrict__ partial, int blocks, float area,
                                 const float* __restrict__ wfc, int branches, float* __restrict__ logits) {
  __shared__ float s_mean[kGateC];
  const int n = blockIdx.x;
  if (threadIdx.x < kGateC) {
    float s = 0.f;
    for (int i = 0; i < blocks; ++i) s += partial[(1LL * n * blocks + i) * kGateC + threadIdx.x];
    s_mean[threadIdx.x] = s / area;   // adaptive_avg_pool2d divides by the window size
  }
  __syncthreads();
  if (threadIdx.x < branches) {
    float s = 0.f;
#pragma unroll
    for (int ci = 0; ci < kGateC; ++ci) s = fmaf(s_mean[ci], wfc[threadIdx.x * kGateC + ci], s);
    logits[n * branches + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------- DiffSoftmax
__global__ void diffsoftmax_fwd_kernel(const float* __restrict__ logits, int rows, int n, float tau, int hard,
                                       float* __restrict__ y, float* __restrict__ y_soft,
                                       int32_t* __restrict__ index) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float z[32];
  float m = -INFINITY;
  for (int i = 0; i < n; ++i) {
    z[i] = logits[r * n + i] / tau;        // same division as the reference (logits / tau)
    m = fmaxf(m, z[i]);
  }
  float sum = 0.f;
  for (int i = 0; i < n; ++i) {
    z[i] = expf(z[i] - m);
    sum += z[i];
  }
  int best = 0;
  float bestv = -1.f;
  for (int i = 0; i < n; ++i) {
    z[i] = z[i] / sum;
    if (z[i] > bestv) {                    // strict '>' keeps the FIRST maximum of y_soft
      bestv = z[i];
      best = i;
    }
  }
  for (int i = 0; i < n; ++i) {
    if (y_soft) y_soft[r * n + i] = z[i];
    y[r * n + i] = hard ? (i == best ? 1.f : 0.f) : z[i];
  }
  if (index) index[r] = best;
}

__global__ void diffsoftmax_bwd_kernel(const float* __restrict__ grad_y, const float* __restrict__ y_soft, int rows,
                                       int n, float tau, float* __restrict__ grad_logits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float dot = 0.f;
  for (int i = 0; i < n; ++i) dot = fmaf(grad_y[r * n + i], y_soft[r * n + i], dot);
  for (int i = 0; i < n; ++i) grad_logits[r * n + i] = y_soft[r * n + i] * (grad_y[r * n + i] - dot) / tau;
}

// ---------------------------------------------------------------- gate plan
// single block; b is small (a batch).  Stable counting sort by the number of
// depth stages a sample needs, so every stage's active set is a prefix.
__global__ void gate_plan_kernel(const float* __restrict__ weight, int b, float* __restrict__ g,
                                 int32_t* __restrict__ perm, int32_t* __restrict__ slot, int32_t* __restrict__ count,
                                 long long* __restrict__ hist) {
  extern __shared__ int s_need[];   // [b] number of leading stages with g != 0
  for (int i = threadIdx.x; i < b; i += blockDim.x) {
    const float w0 = weight[i * 5 + 0], w1 = weight[i * 5 + 1], w2 = weight[i * 5 + 2], w3 = weight[i * 5 + 3],
                w4 = weight[i * 5 + 4];
    // same expression order as the reference: w = w0 (+ w1 (+ w2)); fuse = w*b0 + (1-w)*b1
    const float g1 = 1.f - w0;
    const float g2 = 1.f - (w0 + w1);
    const float g3 = 1.f - ((w0 + w1) + w2);
    const float g4 = w4;
    g[0 * b + i] = g1;
    g[1 * b + i] = g2;
    g[2 * b + i] = g3;
    g[3 * b + i] = g4;
    // a stage's depth features are needed if this or any LATER stage mixes depth in
    int need = 0;
    if (g1 != 0.f) need = 1;
    if (g2 != 0.f) need = 2;
    if (g3 != 0.f) need = 3;
    if (g4 != 0.f) need = 4;
    s_need[i] = need;
    if (hist) {
      int best = 0;
      float bv = w0;
      if (w1 > bv) { bv = w1; best = 1; }
      if (w2 > bv) { bv = w2; best = 2; }
      if (w3 > bv) { bv = w3; best = 3; }
      if (w4 > bv) { bv = w4; best = 4; }
      atomicAdd(reinterpret_cast<unsigned long long*>(hist + best), 1ULL);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int pos = 0;
    int cnt[4] = {0, 0, 0, 0};
    for (int need = 4; need >= 0; --need) {
      for (int i = 0; i < b; ++i) {
        if (s_need[i] == need) {
          perm[pos] = i;
          slot[i] = pos;
          ++pos;
        }
      }
      if (need >= 1) cnt[need - 1] = pos;   // samples needing >= `need` stages
    }
    for (int s = 0; s < 4; ++s) count[s] = cnt[s];
  }
}

// ---------------------------------------------------------------- GAP finish + fc + DiffSoftmax + plan in one launch
// The tail of the gate (gate_head_kernel -> diffsoftmax_fwd_kernel -> gate_plan_kernel, three launches of a few
// microseconds each on the critical path in front of the depth encoder) as ONE single-block kernel: thread n does the
// arithmetic of sample n with exactly the operations and the order of the three kernels, then thread 0 sorts.
__global__ void gate_decide_kernel(const float* __restrict__ partial, int blocks, float area, const float* __restrict__ wfc,
                                   int b, float tau, int hard, float* __restrict__ logits, float* __restrict__ weight,
                                   float* __restrict__ g, int32_t* __restrict__ perm, int32_t* __restrict__ slot,
                                   int32_t* __restrict__ count, long long* __restrict__ hist) {
  extern __shared__ int s_need[];   // [b]
  for (int n = threadIdx.x; n < b; n += blockDim.x) {
    float mean[kGateC];
    for (int c = 0; c < kGateC; ++c) {
      float s = 0.f;
      for (int i = 0; i < blocks; ++i) s += partial[(1LL * n * blocks + i) * kGateC + c];
      mean[c] = s / area;
    }
    float z[5];
    float m = -INFINITY;
    for (int k = 0; k < 5; ++k) {
      float s = 0.f;
#pragma unroll
      for (int ci = 0; ci < kGateC; ++ci) s = fmaf(mean[ci], wfc[k * kGateC + ci], s);
      logits[n * 5 + k] = s;
      z[k] = s / tau;
      m = fmaxf(m, z[k]);
    }
    float sum = 0.f;
    for (int k = 0; k < 5; ++k) {
      z[k] = expf(z[k] - m);
      sum += z[k];
    }
    int best = 0;
    float bestv = -1.f;
    for (int k = 0; k < 5; ++k) {
      z[k] = z[k] / sum;
      if (z[k] > bestv) {
        bestv = z[k];
        best = k;
      }
    }
    float w[5];
    for (int k = 0; k < 5; ++k) {
      w[k] = hard ? (k == best ? 1.f : 0.f) : z[k];
      weight[n * 5 + k] = w[k];
    }
    const float g1 = 1.f - w[0];
    const float g2 = 1.f - (w[0] + w[1]);
    const float g3 = 1.f - ((w[0] + w[1]) + w[2]);
    const float g4 = w[4];
    g[0 * b + n] = g1;
    g[1 * b + n] = g2;
    g[2 * b + n] = g3;
    g[3 * b + n] = g4;
    int need = 0;
    if (g1 != 0.f) need = 1;
    if (g2 != 0.f) need = 2;
    if (g3 != 0.f) need = 3;
    if (g4 != 0.f) need = 4;
    s_need[n] = need;
    if (hist) {
      int hb = 0;
      float bv = w[0];
      if (w[1] > bv) { bv = w[1]; hb = 1; }
      if (w[2] > bv) { bv = w[2]; hb = 2; }
      if (w[3] > bv) { bv = w[3]; hb = 3; }
      if (w[4] > bv) { bv = w[4]; hb = 4; }
      atomicAdd(reinterpret_cast<unsigned long long*>(hist + hb), 1ULL);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int pos = 0;
    int cnt[4] = {0, 0, 0, 0};
    for (int need = 4; need >= 0; --need) {
      for (int i = 0; i < b; ++i) {
        if (s_need[i] == need) {
          perm[pos] = i;
          slot[i] = pos;
          ++pos;
        }
      }
      if (need >= 1) cnt[need - 1] = pos;
    }
    for (int s = 0; s < 4; ++s) count[s] = cnt[s];
  }
}

}  // namespace
}  // namespace dynmm

using namespace dynmm;

extern "C" int dynmm_diffsoftmax_fwd(const float* logits, int rows, int n, float tau, int hard, float* y,
                                     float* y_soft, int32_t* index, void* stream) {
  DYNMM_CHECK_ARG(logits && y && rows >= 0 && n >= 1 && n <= 32 && tau > 0.f, "diffsoftmax_fwd: bad arguments");
  if (rows == 0) return DYNMM_OK;
  diffsoftmax_fwd_kernel<<<ceil_div(rows, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(logits, rows, n, tau,
                                                                                             hard, y, y_soft, index);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_diffsoftmax_bwd(const float* grad_y, const float* y_soft, int rows, int n, float tau,
                                     float* grad_logits, void* stream) {
  DYNMM_CHECK_ARG(grad_y && y_soft && grad_logits && rows >= 0 && n >= 1 && n <= 32 && tau > 0.f,
                  "diffsoftmax_bwd: bad arguments");
  if (rows == 0) return DYNMM_OK;
  diffsoftmax_bwd_kernel<<<ceil_div(rows, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(grad_y, y_soft, rows, n,
                                                                                             tau, grad_logits);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_gate_plan(const float* weight, int b, float* g, int32_t* perm, int32_t* slot, int32_t* count,
                               long long* hist, void* stream) {
  DYNMM_CHECK_ARG(weight && g && perm && slot && count && b >= 1 && b <= 8192, "gate_plan: bad arguments");
  gate_plan_kernel<<<1, 128, b * sizeof(int), static_cast<cudaStream_t>(stream)>>>(weight, b, g, perm, slot, count,
                                                                                   hist);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

static void gate_dims(int h, int w, int* h1, int* w1, int* h2, int* w2) {
  *h1 = (h - 5) / 2 + 1;
  *w1 = (w - 5) / 2 + 1;
  *h2 = (*h1 - 5) / 2 + 1;
  *w2 = (*w1 - 5) / 2 + 1;
}

extern "C" long long dynmm_global_gate_workspace(int b, int h, int w) {
  int h1, w1, h2, w2;
  gate_dims(h, w, &h1, &w1, &h2, &w2);
  if (h1 < 5 || w1 < 5) return -1;
  const long long conv1 = 1LL * b * h1 * w1 * kGateC * sizeof(float);
  const long long partial = 1LL * b * ceil_div(h2 * w2, 128) * kGateC * sizeof(float);
  return conv1 + partial + 256;
}

namespace {
// conv1 + conv2/GAP launches shared by the two entry points; *partial_out / *blocks_out: the GAP partial sums
int gate_convs(const float* rgb, const float* depth, int b, int h, int w, const float* w1p, const float* scale1,
               const float* shift1, const float* w2p, const float* scale2, const float* shift2, void* work,
               float** partial_out, int* blocks_out, float* area_out, cudaStream_t stream) {
  int h1, w1, h2, w2;
  gate_dims(h, w, &h1, &w1, &h2, &w2);
  DYNMM_CHECK_ARG(b >= 1 && h1 >= 5 && w1 >= 5, "global_gate: feature map %dx%d too small", h, w);
  float* conv1 = static_cast<float*>(work);
  const long long conv1_bytes = (1LL * b * h1 * w1 * kGateC * sizeof(float) + 255) / 256 * 256;
  float* partial = reinterpret_cast<float*>(static_cast<char*>(work) + conv1_bytes);
  const int smem1 = kGateC * kTaps * 128 * sizeof(float);
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run([] {
    return cudaFuncSetAttribute(gate_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGateC * kTaps * 128 * 4);
  }));
  const long long items = 1LL * b * h1 * ceil_div(w1, kPx);
  int grid1 = 2 * num_sms();
  if (grid1 > ceil_div_ll(items, 8)) grid1 = (int)ceil_div_ll(items, 8);
  gate_conv1_kernel<<<grid1, 256, smem1, stream>>>(rgb, depth, b, h, w, h1, w1, w1p, scale1, shift1, conv1);
  DYNMM_LAUNCH_CHECK();
  const int blocks2 = ceil_div(h2 * w2, 128);
  gate_conv2_gap_kernel<<<dim3(blocks2, b), 128, 0, stream>>>(conv1, h1, w1, h2, w2, w2p, scale2, shift2, partial);
  DYNMM_LAUNCH_CHECK();
  *partial_out = partial;
  *blocks_out = blocks2;
  *area_out = (float)(h2 * w2);
  return DYNMM_OK;
}
}  // namespace

extern "C" int dynmm_global_gate_decide(const float* rgb, const float* depth, int b, int h, int w, const float* w1p,
                                        const float* scale1, const float* shift1, const float* w2p,
                                        const float* scale2, const float* shift2, const float* wfc, void* work,
                                        float tau, int hard, float* logits, float* weight, float* g, int32_t* perm,
                                        int32_t* slot, int32_t* count, long long* hist, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(rgb && depth && w1p && w2p && wfc && work && logits && weight && g && perm && slot && count,
                  "global_gate_decide: null pointer");
  DYNMM_CHECK_ARG(b <= 8192, "global_gate_decide: batch too large");
  float* partial = nullptr;
  int blocks2 = 0;
  float area = 0.f;
  int rc = gate_convs(rgb, depth, b, h, w, w1p, scale1, shift1, w2p, scale2, shift2, work, &partial, &blocks2, &area, stream);
  if (rc) return rc;
  gate_decide_kernel<<<1, 128, b * sizeof(int), stream>>>(partial, blocks2, area, wfc, b, tau, hard, logits, weight, g, perm,
                                                         slot, count, hist);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_global_gate_logits(const float* rgb, const float* depth, int b, int h, int w, const float* w1p,
                                        const float* scale1, const float* shift1, const float* w2p,
                                        const float* scale2, const float* shift2, const float* wfc, void* work,
                                        float* logits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(rgb && depth && w1p && w2p && wfc && work && logits, "global_gate: null pointer");
  int h1, w1, h2, w2;
  gate_dims(h, w, &h1, &w1, &h2, &w2);
  DYNMM_CHECK_ARG(b >= 1 && h1 >= 5 && w1 >= 5, "global_gate: feature map %dx%d too small", h, w);
  float* conv1 = static_cast<float*>(work);
  const long long conv1_bytes = (1LL * b * h1 * w1 * kGateC * sizeof(float) + 255) / 256 * 256;
  float* partial = reinterpret_cast<float*>(static_cast<char*>(work) + conv1_bytes);
  const int smem1 = kGateC * kTaps * 128 * sizeof(float);
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run([] {
    return cudaFuncSetAttribute(gate_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGateC * kTaps * 128 * 4);
  }));
  const long long items = 1LL * b * h1 * ceil_div(w1, kPx);
  int grid1 = 2 * num_sms();
  if (grid1 > ceil_div_ll(items, 8)) grid1 = (int)ceil_div_ll(items, 8);
  gate_conv1_kernel<<<grid1, 256, smem1, stream>>>(rgb, depth, b, h, w, h1, w1, w1p, scale1, shift1, conv1);
  DYNMM_LAUNCH_CHECK();
  const int blocks2 = ceil_div(h2 * w2, 128);
  gate_conv2_gap_kernel<<<dim3(blocks2, b), 128, 0, stream>>>(conv1, h1, w1, h2, w2, w2p, scale2, shift2, partial);
  DYNMM_LAUNCH_CHECK();
  gate_head_kernel<<<b, 32, 0, stream>>>(partial, blocks2, (float)(h2 * w2), wfc, 5, logits);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
