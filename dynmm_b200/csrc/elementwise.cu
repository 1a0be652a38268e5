// Memory-bound glue of the gated path: gated adds, the modality-level expert mix,
// row compaction, layout conversion, learned 2x upsampling and pyramid pooling.
// All kernels are coalesced, 16-byte vectorised where the layout allows, and sized
// as a multiple of the SM count with grid-stride loops.
#include <stdlib.h>

#include "common.cuh"

namespace dynmm {
namespace {

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = 1LL * num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------- gated add (bf16, inference)
__global__ void gated_add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                      const float* __restrict__ gate, const int32_t* __restrict__ slot, int n,
                                      long long vec_per_sample, uint4* __restrict__ out) {
  const long long total = 1LL * n * vec_per_sample;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int s = (int)(i / vec_per_sample);
    const long long r = i - 1LL * s * vec_per_sample;
    uint4 va = a[i];
    const float g = gate[s];
    if (g != 0.f) {   // gated-off samples never touch b
      const int bs = slot ? slot[s] : s;
      const uint4 vb = __ldg(&b[1LL * bs * vec_per_sample + r]);
      va.x = pack_bf16(bf16_lo(va.x) + g * bf16_lo(vb.x), bf16_hi(va.x) + g * bf16_hi(vb.x));
      va.y = pack_bf16(bf16_lo(va.y) + g * bf16_lo(vb.y), bf16_hi(va.y) + g * bf16_hi(vb.y));
      va.z = pack_bf16(bf16_lo(va.z) + g * bf16_lo(vb.z), bf16_hi(va.z) + g * bf16_hi(vb.z));
      va.w = pack_bf16(bf16_lo(va.w) + g * bf16_lo(vb.w), bf16_hi(va.w) + g * bf16_hi(vb.w));
    }
    out[i] = va;
  }
}

// ---------------------------------------------------------------- gated add (fp32, training path)
__global__ void gated_add_f32_fwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                         const float* __restrict__ gate, int n, long long vec_per_sample,
                                         float4* __restrict__ out) {
  const long long total = 1LL * n * vec_per_sample;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const float g = gate[i / vec_per_sample];
    float4 va = a[i];
    const float4 vb = b[i];
    va.x = fmaf(g, vb.x, va.x);
    va.y = fmaf(g, vb.y, va.y);
    va.z = fmaf(g, vb.z, va.z);
    va.w = fmaf(g, vb.w, va.w);
    out[i] = va;
  }
}

// grad_b = g * grad ; grad_gate[n] = sum(grad * b) (fixed-order two-level reduction)
__global__ void gated_add_f32_bwd_kernel(const float4* __restrict__ grad, const float4* __restrict__ b,
                                         const float* __restrict__ gate, long long vec_per_sample,
                                         float4* __restrict__ grad_b, float* __restrict__ partial) {
  __shared__ float s_red[8];
  const int s = blockIdx.y;
  const float g = gate[s];
  float acc = 0.f;
  const long long base = 1LL * s * vec_per_sample;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < vec_per_sample;
       i += 1LL * gridDim.x * blockDim.x) {
    const float4 gr = grad[base + i];
    const float4 vb = b[base + i];
    acc += gr.x * vb.x + gr.y * vb.y + gr.z * vb.z + gr.w * vb.w;
    if (grad_b) grad_b[base + i] = make_float4(g * gr.x, g * gr.y, g * gr.z, g * gr.w);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    partial[1LL * s * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int per_row, float* __restrict__ out) {
  // one warp per row, lanes stride the partials, fixed butterfly order
  const int row = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < per_row; i += 32) acc += partial[1LL * row * per_row + i];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) out[row] = acc;
}

// ---------------------------------------------------------------- modality-level expert mix
struct MixArgs {
  const float* pred[4];
  const int32_t* rows[4];
  float* grad_pred[4];
};
__global__ void softgate_mix_fwd_kernel(MixArgs a, const float* __restrict__ w, int b, int c, int ne,
                                        float* __restrict__ out) {
  const long long total = 1LL * b * c;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int r = (int)(i / c), col = (int)(i - 1LL * r * c);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e >= ne) break;
      const float we = w[r * ne + e];
      if (we != 0.f) {   // an expert with zero weight may not have been evaluated for this row at all
        const int src = a.rows[e] ? a.rows[e][r] : r;
        acc = fmaf(we, a.pred[e][1LL * src * c + col], acc);
      }
    }
    out[i] = acc;
  }
}
// one warp per row: grad_pred[e] = w[e]*grad ; grad_w[e] = <grad, pred[e]>
__global__ void softgate_mix_bwd_kernel(MixArgs a, const float* __restrict__ grad_out, const float* __restrict__ w,
                                        int b, int c, int ne, float* __restrict__ grad_w) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= b) return;
  float dots[4] = {0.f, 0.f, 0.f, 0.f};
  for (int col = lane; col < c; col += 32) {
    const float g = grad_out[1LL * r * c + col];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e >= ne) break;
      dots[e] = fmaf(g, a.pred[e][1LL * r * c + col], dots[e]);
      if (a.grad_pred[e]) a.grad_pred[e][1LL * r * c + col] = w[r * ne + e] * g;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dots[e] += __shfl_xor_sync(0xffffffffu, dots[e], o);
    if (lane == 0 && e < ne && grad_w) grad_w[r * ne + e] = dots[e];
  }
}

// stable compaction of the rows that route to `expert` (single block; b is a batch)
__global__ void compact_rows_kernel(const float* __restrict__ w, int b, int ne, int expert, int32_t* __restrict__ idx,
                                    int32_t* __restrict__ inv, int32_t* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int start = 0; start < b; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const bool keep = i < b && w[i * ne + expert] != 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int k = 0; k < warp; ++k) off += s_warp[k];
    if (keep) {
      const int pos = off + __popc(m & ((1u << lane) - 1));
      idx[pos] = i;
      inv[i] = pos;
    } else if (i < b) {
      inv[i] = -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < nwarps; ++k) t += s_warp[k];
      s_base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_base;
}

// ---------------------------------------------------------------- layout conversion (32x32 smem transpose)
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ in, int c, long long hw,
                                             __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = blockIdx.x * 32LL;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ch = c0 + j;
    const long long p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (ch < c && p < hw) ? in[(1LL * n * c + ch) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int ch = c0 + threadIdx.x;
    if (ch < c && p < hw) out[(1LL * n * hw + p) * c + ch] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ in, int c, long long hw, int ld,
                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = blockIdx.x * 32LL;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int ch = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (ch < c && p < hw) ? __bfloat162float(in[(1LL * n * hw + p) * ld + ch]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ch = c0 + j;
    const long long p = p0 + threadIdx.x;
    if (ch < c && p < hw) out[(1LL * n * c + ch) * hw + p] = tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------- nearest x2 + depthwise 3x3 (zero pad) + bias (+ skip)
// Each output pixel (Y,X) reads the 3x3 neighbourhood of the nearest-upsampled map, i.e. input
// pixels ((Y+dy)>>1, (X+dx)>>1); the 4x larger intermediate of the reference never exists.
// NHWC variant: a thread owns 8 channels of one output pixel (16-byte loads/stores).
// `clamp`: replication padding of the up-sampled map (Upsample 'learned-3x3' / the fixed bilinear stencil, model.py:
// 372-399) instead of zero padding -- out-of-range taps read the border pixel.
__global__ void upsample2x_dw_nhwc_kernel(const __nv_bfloat16* __restrict__ in, int n, int h, int w, int c,
                                          const float* __restrict__ wgt, const float* __restrict__ bias,
                                          const __nv_bfloat16* __restrict__ skip, __nv_bfloat16* __restrict__ out,
                                          int clamp) {
  const int cv = c >> 3;
  const long long total = 1LL * n * 4 * h * w * cv;
  const int H = 2 * h, W = 2 * w;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv) * 8;
    long long r = i / cv;
    const int X = (int)(r % W);
    r /= W;
    const int Y = (int)(r % H);
    const int s = (int)(r / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias ? bias[c8 + e] : 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      int yy = Y + dy;
      if (clamp) yy = min(max(yy, 0), H - 1);
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        int xx = X + dx;
        if (clamp) xx = min(max(xx, 0), W - 1);
        if (xx < 0 || xx >= W) continue;
        const uint4 v =
            __ldg(reinterpret_cast<const uint4*>(in + ((1LL * s * h + (yy >> 1)) * w + (xx >> 1)) * c + c8));
        const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                            bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
        const int k = (dy + 1) * 3 + dx + 1;
        const float4 wa = __ldg(reinterpret_cast<const float4*>(wgt + k * c + c8));       // weights are [9][c]
        const float4 wb = __ldg(reinterpret_cast<const float4*>(wgt + k * c + c8 + 4));
        acc[0] = fmaf(f[0], wa.x, acc[0]); acc[1] = fmaf(f[1], wa.y, acc[1]);
        acc[2] = fmaf(f[2], wa.z, acc[2]); acc[3] = fmaf(f[3], wa.w, acc[3]);
        acc[4] = fmaf(f[4], wb.x, acc[4]); acc[5] = fmaf(f[5], wb.y, acc[5]);
        acc[6] = fmaf(f[6], wb.z, acc[6]); acc[7] = fmaf(f[7], wb.w, acc[7]);
      }
    }
    const long long o = ((1LL * s * H + Y) * W + X) * c + c8;
    if (skip) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(skip + o));
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
    uint4 ov;
    ov.x = pack_bf16(acc[0], acc[1]);
    ov.y = pack_bf16(acc[2], acc[3]);
    ov.z = pack_bf16(acc[4], acc[5]);
    ov.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + o) = ov;
  }
}
// Strip variant of the NHWC kernel (used when c % 4 == 0): a thread owns 4 channels of one INPUT column and walks
// down a strip of input rows with a 3x3 sliding window in registers.  Per input pixel it loads 3 x 8 bytes (the
// new row), produces the 2x2 output block with the 16 parity stencils precombined from the 3x3 weights once per
// thread (16 FMAs per channel instead of 36) and stores 4 x 8 bytes; the per-pixel kernel above re-loads 9 inputs
// and 18 weight vectors for every 16-byte store.
// Strip length: the kernel holds 64 stencil + 36 window registers per thread, so two 256-thread CTAs fit an SM and a
// launch runs in waves of 2 x SMs CTAs; every CTA loads rows + 2 input rows.  strip_rows() minimises
// waves x (rows + 2) -- 8-row strips with "at least 4 CTAs per SM" gave 640 CTAs = 2.16 waves (three, the last one at
// 16 %) for the 60x80 -> 120x160 module of the decoder.
template <bool kSkip, bool kSplit = false>       // kSplit: in / skip / out are [hi | lo] halves with pitch 2c (f32x3 mode)
__global__ void __launch_bounds__(256)
upsample2x_dw_nhwc_strip_kernel(const __nv_bfloat16* __restrict__ in, int h, int w, int c,
                                const float* __restrict__ wgt, const float* __restrict__ bias,
                                const __nv_bfloat16* __restrict__ skip, __nv_bfloat16* __restrict__ out, int strip_rows,
                                int clamp) {
  const int cg = c >> 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w * cg) return;
  const int c4 = (idx % cg) * 4, x = idx / cg;
  const int s = blockIdx.z;
  const int y0 = blockIdx.y * strip_rows, y1 = min(y0 + strip_rows, h);
  const int H = 2 * h, W = 2 * w;
  // parity stencils (see upsample2x_dw_to_nchw_kernel): st[p][t][ch], p = 2*(Y&1) + (X&1), t = the 4 inputs of the block
  float st[4][4][4], b4[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = __ldg(wgt + t * c + c4 + ch);
    b4[ch] = bias ? __ldg(bias + c4 + ch) : 0.f;
    st[0][0][ch] = k[0];               st[0][1][ch] = k[1] + k[2];        st[0][2][ch] = k[3] + k[6];        st[0][3][ch] = (k[4] + k[5]) + (k[7] + k[8]);
    st[1][0][ch] = k[0] + k[1];        st[1][1][ch] = k[2];               st[1][2][ch] = (k[3] + k[4]) + (k[6] + k[7]); st[1][3][ch] = k[5] + k[8];
    st[2][0][ch] = k[0] + k[3];        st[2][1][ch] = (k[1] + k[2]) + (k[4] + k[5]); st[2][2][ch] = k[6];    st[2][3][ch] = k[7] + k[8];
    st[3][0][ch] = (k[0] + k[1]) + (k[3] + k[4]); st[3][1][ch] = k[2] + k[5]; st[3][2][ch] = k[6] + k[7];    st[3][3][ch] = k[8];
  }
  const int ld = kSplit ? 2 * c : c;
  const __nv_bfloat16* base = in + static_cast<size_t>(s) * h * w * ld + c4;
  auto load_row = [&](int y, float (&r)[3][4]) {        // columns x-1, x, x+1 of input row y (zeros / border outside)
    if (clamp) y = min(max(y, 0), h - 1);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int xx = x + j - 1;
      if (clamp) xx = min(max(xx, 0), w - 1);
      uint2 v = make_uint2(0u, 0u), q = make_uint2(0u, 0u);
      if (y >= 0 && y < h && xx >= 0 && xx < w) {
        v = __ldg(reinterpret_cast<const uint2*>(base + (static_cast<size_t>(y) * w + xx) * ld));
        if (kSplit) q = __ldg(reinterpret_cast<const uint2*>(base + (static_cast<size_t>(y) * w + xx) * ld + c));
      }
      r[j][0] = bf16_lo(v.x) + bf16_lo(q.x); r[j][1] = bf16_hi(v.x) + bf16_hi(q.x);
      r[j][2] = bf16_lo(v.y) + bf16_lo(q.y); r[j][3] = bf16_hi(v.y) + bf16_hi(q.y);
    }
  };
  float up[3][4], mid[3][4], dn[3][4];
  load_row(y0 - 1, up);
  load_row(y0, mid);
  for (int y = y0; y < y1; ++y) {
    load_row(y + 1, dn);
    float o[4][4];      // [parity][ch]
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      // (2y, 2x): (y-1,x-1) (y-1,x) (y,x-1) (y,x)      (2y, 2x+1): (y-1,x) (y-1,x+1) (y,x) (y,x+1)
      o[0][ch] = fmaf(st[0][3][ch], mid[1][ch], fmaf(st[0][2][ch], mid[0][ch], fmaf(st[0][1][ch], up[1][ch], fmaf(st[0][0][ch], up[0][ch], b4[ch]))));
      o[1][ch] = fmaf(st[1][3][ch], mid[2][ch], fmaf(st[1][2][ch], mid[1][ch], fmaf(st[1][1][ch], up[2][ch], fmaf(st[1][0][ch], up[1][ch], b4[ch]))));
      // (2y+1, 2x): (y,x-1) (y,x) (y+1,x-1) (y+1,x)    (2y+1, 2x+1): (y,x) (y,x+1) (y+1,x) (y+1,x+1)
      o[2][ch] = fmaf(st[2][3][ch], dn[1][ch], fmaf(st[2][2][ch], dn[0][ch], fmaf(st[2][1][ch], mid[1][ch], fmaf(st[2][0][ch], mid[0][ch], b4[ch]))));
      o[3][ch] = fmaf(st[3][3][ch], dn[2][ch], fmaf(st[3][2][ch], dn[1][ch], fmaf(st[3][1][ch], mid[2][ch], fmaf(st[3][0][ch], mid[1][ch], b4[ch]))));
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const size_t oo = ((static_cast<size_t>(s) * H + 2 * y + (p >> 1)) * W + 2 * x + (p & 1)) * ld + c4;
      if (kSkip) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(skip + oo));
        o[p][0] += bf16_lo(v.x); o[p][1] += bf16_hi(v.x); o[p][2] += bf16_lo(v.y); o[p][3] += bf16_hi(v.y);
        if (kSplit) {
          const uint2 q = __ldg(reinterpret_cast<const uint2*>(skip + oo + c));
          o[p][0] += bf16_lo(q.x); o[p][1] += bf16_hi(q.x); o[p][2] += bf16_lo(q.y); o[p][3] += bf16_hi(q.y);
        }
      }
      uint2 ov;
      ov.x = pack_bf16(o[p][0], o[p][1]);
      ov.y = pack_bf16(o[p][2], o[p][3]);
      *reinterpret_cast<uint2*>(out + oo) = ov;
      if (kSplit) {
        uint2 lv;
        lv.x = pack_bf16(o[p][0] - bf16_lo(ov.x), o[p][1] - bf16_hi(ov.x));
        lv.y = pack_bf16(o[p][2] - bf16_lo(ov.y), o[p][3] - bf16_hi(ov.y));
        *reinterpret_cast<uint2*>(out + oo + c) = lv;
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        up[j][ch] = mid[j][ch];
        mid[j][ch] = dn[j][ch];
      }
  }
}
// Final upsample: NHWC bf16 in -> NCHW fp32 out (the module's return layout; 393 MB of stores at
// 8x40x480x640, the HBM-bound tail of the forward).  A CTA stages an 8x32 input tile (+1 halo) for
// all channels in shared memory as fp32 [c][row][col]; then one warp per (channel, output row)
// produces 64 consecutive output columns: lane l owns columns 2l, 2l+1 (one 8-byte store, 256 B per
// warp, fully coalesced), reading 2x3 staged inputs without bank conflicts.
constexpr int kUpTy = 8, kUpTx = 32;
template <bool kSplit>      // kSplit: `in` is [n,h,w,2c] = [hi | lo] halves of fp32-grade values (DYNMM_CONV_SPLIT layout)
__global__ void __launch_bounds__(256)
upsample2x_dw_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, int h, int w, int c,
                             const float* __restrict__ wgt, const float* __restrict__ bias,
                             float* __restrict__ out, uint8_t* __restrict__ labels, int clamp, int c_valid) {
  // c_valid <= c: channels that exist in the NCHW output / take part in the arg-max (the NHWC input may be padded to a
  // multiple of 8 channels, e.g. 37 classes carried as 40)
  extern __shared__ float s_in[];                 // [c][kUpTy + 2][kUpTx + 2 (+1 pad)], then [c][16] stencil weights
  constexpr int RS = kUpTx + 3;                   // row stride (35): odd -> conflict-free transposed fill
  constexpr int CS = (kUpTy + 2) * RS;            // channel stride
  float* s_w = s_in + c * CS;
  // per-channel 2x2-parity stencils on the input, combined once per CTA from the 3x3 weights:
  //   [0..3]  output (2y  ,2x  ): inputs (y-1,x-1) (y-1,x) (y,x-1) (y,x)
  //   [4..7]  output (2y  ,2x+1): inputs (y-1,x) (y-1,x+1) (y,x) (y,x+1)
  //   [8..11] output (2y+1,2x  ): inputs (y,x-1) (y,x) (y+1,x-1) (y+1,x)
  //   [12..15]output (2y+1,2x+1): inputs (y,x) (y,x+1) (y+1,x) (y+1,x+1)
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = __ldg(wgt + t * c + ch);
    float* o = s_w + ch * 16;
    o[0] = k[0];                 o[1] = k[1] + k[2];               o[2] = k[3] + k[6];           o[3] = (k[4] + k[5]) + (k[7] + k[8]);
    o[4] = k[0] + k[1];          o[5] = k[2];                      o[6] = (k[3] + k[4]) + (k[6] + k[7]); o[7] = k[5] + k[8];
    o[8] = k[0] + k[3];          o[9] = (k[1] + k[2]) + (k[4] + k[5]); o[10] = k[6];             o[11] = k[7] + k[8];
    o[12] = (k[0] + k[1]) + (k[3] + k[4]); o[13] = (k[2] + k[5]);  o[14] = k[6] + k[7];          o[15] = k[8];
  }
  const int H = 2 * h, W = 2 * w;
  const int x0 = blockIdx.x * kUpTx, y0 = blockIdx.y * kUpTy, s = blockIdx.z;
  const int cv = c >> 3;
  // fill: one thread per (pixel, 8-channel chunk); zero outside the image (the conv's zero padding
  // applies to the UPSAMPLED map, handled below by tap masks; here out-of-image inputs are never used)
  for (int i = threadIdx.x; i < (kUpTy + 2) * (kUpTx + 2) * cv; i += blockDim.x) {
    const int ch8 = (i % cv) * 8;
    const int p = i / cv;
    const int lx = p % (kUpTx + 2), ly = p / (kUpTx + 2);
    int y = y0 + ly - 1, x = x0 + lx - 1;
    if (clamp) {
      y = min(max(y, 0), h - 1);
      x = min(max(x, 0), w - 1);
    }
    uint4 v = make_uint4(0, 0, 0, 0), q = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < h && x >= 0 && x < w) {
      const __nv_bfloat16* px = in + ((1LL * s * h + y) * w + x) * (kSplit ? 2 * c : c) + ch8;
      v = __ldg(reinterpret_cast<const uint4*>(px));
      if (kSplit) q = __ldg(reinterpret_cast<const uint4*>(px + c));
    }
    float* d = s_in + ch8 * CS + ly * RS + lx;
    d[0 * CS] = bf16_lo(v.x) + bf16_lo(q.x); d[1 * CS] = bf16_hi(v.x) + bf16_hi(q.x);
    d[2 * CS] = bf16_lo(v.y) + bf16_lo(q.y); d[3 * CS] = bf16_hi(v.y) + bf16_hi(q.y);
    d[4 * CS] = bf16_lo(v.z) + bf16_lo(q.z); d[5 * CS] = bf16_hi(v.z) + bf16_hi(q.z);
    d[6 * CS] = bf16_lo(v.w) + bf16_lo(q.w); d[7 * CS] = bf16_hi(v.w) + bf16_hi(q.w);
  }
  __syncthreads();
  // A lane owns input column x = x0 + lane and produces the 2x2 output block (2y..2y+1, 2x..2x+1).
  // Nearest-x2 followed by a 3x3 conv collapses to a 2x2 stencil on the INPUT per output parity:
  //   rows: parity 0 uses inputs (y-1: w[-1]) and (y: w[0]+w[+1]); parity 1 uses (y: w[-1]+w[0]) and (y+1: w[+1])
  // (same for columns).  Inputs outside the image were staged as zeros, which is exactly the conv's
  // zero padding of the upsampled map, so no border masks are needed.
  // A warp owns one input row of the tile and walks the channels, so the per-pixel arg-max over
  // classes (eval.py:120) can be kept in registers and emitted without re-reading the logits.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool x_ok = x0 + lane < w;
  for (int ly = warp; ly < kUpTy; ly += nwarps) {
    const int y = y0 + ly;
    if (y >= h) continue;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {0, 0, 0, 0};
#pragma unroll 2
    for (int ch = 0; ch < c_valid; ++ch) {
      const float b0 = bias ? __ldg(bias + ch) : 0.f;
      const float* base = s_in + ch * CS + ly * RS + lane;          // staged rows ly, ly+1, ly+2 = y-1, y, y+1
      const float al = base[0], am = base[1], ar = base[2];
      const float bl = base[RS], bm = base[RS + 1], br = base[RS + 2];
      const float dl = base[2 * RS], dm = base[2 * RS + 1], dr = base[2 * RS + 2];
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + ch * 16);        // warp-wide broadcasts
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + ch * 16 + 4);
      const float4 w2 = *reinterpret_cast<const float4*>(s_w + ch * 16 + 8);
      const float4 w3 = *reinterpret_cast<const float4*>(s_w + ch * 16 + 12);
      const float o00 = b0 + w0.x * al + w0.y * am + w0.z * bl + w0.w * bm;
      const float o01 = b0 + w1.x * am + w1.y * ar + w1.z * bm + w1.w * br;
      const float o10 = b0 + w2.x * bl + w2.y * bm + w2.z * dl + w2.w * dm;
      const float o11 = b0 + w3.x * bm + w3.y * br + w3.z * dm + w3.w * dr;
      if (out && x_ok) {
        float* o = out + ((1LL * s * c_valid + ch) * H + 2LL * y) * W + 2 * (x0 + lane);
        *reinterpret_cast<float2*>(o) = make_float2(o00, o01);
        *reinterpret_cast<float2*>(o + W) = make_float2(o10, o11);
      }
      if (o00 > best[0]) { best[0] = o00; arg[0] = ch; }      // strict '>': first maximum, like torch.argmax
      if (o01 > best[1]) { best[1] = o01; arg[1] = ch; }
      if (o10 > best[2]) { best[2] = o10; arg[2] = ch; }
      if (o11 > best[3]) { best[3] = o11; arg[3] = ch; }
    }
    if (labels && x_ok) {
      uint8_t* l = labels + (1LL * s * H + 2LL * y) * W + 2 * (x0 + lane);
      *reinterpret_cast<uchar2*>(l) = make_uchar2((unsigned char)arg[0], (unsigned char)arg[1]);
      *reinterpret_cast<uchar2*>(l + W) = make_uchar2((unsigned char)arg[2], (unsigned char)arg[3]);
    }
  }
}

// ---------------------------------------------------------------- pyramid pooling helpers
__global__ void adaptive_avgpool_kernel(const __nv_bfloat16* __restrict__ in, int h, int w, int c, int ld, int bins,
                                        __nv_bfloat16* __restrict__ out) {
  // grid (bins*bins, n, c/64); a block owns 64 channels of one window: 4 pixel lanes x 64 channels,
  // fixed-order reduction.  window = [floor(i*h/bins), ceil((i+1)*h/bins))  (adaptive_avg_pool2d)
  __shared__ float s_part[4][64];
  const int by = blockIdx.x / bins, bx = blockIdx.x % bins, s = blockIdx.y;
  const int y0 = (by * h) / bins, y1 = ((by + 1) * h + bins - 1) / bins;
  const int x0 = (bx * w) / bins, x1 = ((bx + 1) * w + bins - 1) / bins;
  const int ww = x1 - x0, npix = (y1 - y0) * ww;
  const int ch = blockIdx.z * 64 + (threadIdx.x & 63), lane_p = threadIdx.x >> 6;
  float acc = 0.f;
  if (ch < c) {
    for (int p = lane_p; p < npix; p += 4) {
      const int y = y0 + p / ww, x = x0 + p % ww;
      acc += __bfloat162float(in[((1LL * s * h + y) * w + x) * ld + ch]);
    }
  }
  s_part[lane_p][threadIdx.x & 63] = acc;
  __syncthreads();
  if (lane_p == 0 && ch < c) {
    const float t = (s_part[0][threadIdx.x] + s_part[1][threadIdx.x]) + (s_part[2][threadIdx.x] + s_part[3][threadIdx.x]);
    out[((1LL * s * bins + by) * bins + bx) * c + ch] = __float2bfloat16_rn(t / (float)npix);
  }
}
// 16-byte variant (c, ld multiples of 8): 32 pixel lanes x 8 channel octets per block, ~10 independent loads per
// thread for the 15x20 global-pool window instead of 75 dependent 2-byte ones; fixed-order reduction.
__global__ void __launch_bounds__(256)
adaptive_avgpool_v8_kernel(const __nv_bfloat16* __restrict__ in, int h, int w, int c, int ld, int bins,
                           __nv_bfloat16* __restrict__ out) {
  __shared__ float s_part[32][64 + 4];
  const int by = blockIdx.x / bins, bx = blockIdx.x % bins, s = blockIdx.y;
  const int y0 = (by * h) / bins, y1 = ((by + 1) * h + bins - 1) / bins;
  const int x0 = (bx * w) / bins, x1 = ((bx + 1) * w + bins - 1) / bins;
  const int ww = x1 - x0, npix = (y1 - y0) * ww;
  const int oct = threadIdx.x & 7, lane_p = threadIdx.x >> 3;
  const int ch0 = blockIdx.z * 64 + oct * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ch0 < c) {
    for (int p = lane_p; p < npix; p += 32) {
      const int y = y0 + p / ww, x = x0 + p % ww;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((1LL * s * h + y) * w + x) * ld + ch0));
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_part[lane_p][oct * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int ch = blockIdx.z * 64 + threadIdx.x;
    if (ch < c) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 32; ++l) t += s_part[l][threadIdx.x];
      out[((1LL * s * bins + by) * bins + bx) * c + ch] = __float2bfloat16_rn(t / (float)npix);
    }
  }
}
__global__ void nearest_resize_into_kernel(const __nv_bfloat16* __restrict__ src, int n, int hs, int ws, int c,
                                           __nv_bfloat16* __restrict__ dst, int h, int w, int ld, int c_off) {
  const int cv = c >> 3;
  const long long total = 1LL * n * h * w * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv) * 8;
    long long r = i / cv;
    const int x = (int)(r % w);
    r /= w;
    const int y = (int)(r % h);
    const int s = (int)(r / h);
    // F.interpolate(mode='nearest'): src = floor(dst * in/out)
    const int sy = min((int)((long long)y * hs / h), hs - 1), sx = min((int)((long long)x * ws / w), ws - 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + ((1LL * s * hs + sy) * ws + sx) * c + c8));
    *reinterpret_cast<uint4*>(dst + ((1LL * s * h + y) * w + x) * ld + c_off + c8) = v;
  }
}


// ---------------------------------------------------------------- fp32-grade ("f32x3") variants: [hi | lo] bf16 halves
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): ~16 mantissa bits in two bf16 tensors (DYNMM_CONV_SPLIT layout:
// the hi halves in channels [0, ld/2), the lo halves in [ld/2, ld)).  The kernels below read both halves, work in fp32
// and write both halves.
__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
  hi.x = pack_bf16(f[0], f[1]); hi.y = pack_bf16(f[2], f[3]); hi.z = pack_bf16(f[4], f[5]); hi.w = pack_bf16(f[6], f[7]);
  lo.x = pack_bf16(f[0] - bf16_lo(hi.x), f[1] - bf16_hi(hi.x));
  lo.y = pack_bf16(f[2] - bf16_lo(hi.y), f[3] - bf16_hi(hi.y));
  lo.z = pack_bf16(f[4] - bf16_lo(hi.z), f[5] - bf16_hi(hi.z));
  lo.w = pack_bf16(f[6] - bf16_lo(hi.w), f[7] - bf16_hi(hi.w));
}
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float (&f)[8]) {
  f[0] = bf16_lo(hi.x) + bf16_lo(lo.x); f[1] = bf16_hi(hi.x) + bf16_hi(lo.x);
  f[2] = bf16_lo(hi.y) + bf16_lo(lo.y); f[3] = bf16_hi(hi.y) + bf16_hi(lo.y);
  f[4] = bf16_lo(hi.z) + bf16_lo(lo.z); f[5] = bf16_hi(hi.z) + bf16_hi(lo.z);
  f[6] = bf16_lo(hi.w) + bf16_lo(lo.w); f[7] = bf16_hi(hi.w) + bf16_hi(lo.w);
}

// fp32 [rows][c] -> [rows][2c]
__global__ void split_from_f32_kernel(const float* __restrict__ x, long long rows, int c, __nv_bfloat16* __restrict__ out) {
  const int cv = c >> 3;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv) * 8;
    const long long r = i / cv;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * c + c8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + r * c + c8 + 4));
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    split8(f, hi, lo);
    *reinterpret_cast<uint4*>(out + r * 2 * c + c8) = hi;
    *reinterpret_cast<uint4*>(out + r * 2 * c + c + c8) = lo;
  }
}

// nearest x2 + depthwise 3x3 + bias (+ skip), NHWC split in / out: a thread owns 8 channels of one output pixel
__global__ void upsample2x_dw_nhwc_split_kernel(const __nv_bfloat16* __restrict__ in, int n, int h, int w, int c,
                                                const float* __restrict__ wgt, const float* __restrict__ bias,
                                                const __nv_bfloat16* __restrict__ skip, __nv_bfloat16* __restrict__ out,
                                                int clamp) {
  const int cv = c >> 3;
  const long long total = 1LL * n * 4 * h * w * cv;
  const int H = 2 * h, W = 2 * w;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv) * 8;
    long long r = i / cv;
    const int X = (int)(r % W);
    r /= W;
    const int Y = (int)(r % H);
    const int s = (int)(r / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias ? bias[c8 + e] : 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      int yy = Y + dy;
      if (clamp) yy = min(max(yy, 0), H - 1);
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        int xx = X + dx;
        if (clamp) xx = min(max(xx, 0), W - 1);
        if (xx < 0 || xx >= W) continue;
        const __nv_bfloat16* px = in + ((1LL * s * h + (yy >> 1)) * w + (xx >> 1)) * 2 * c + c8;
        float f[8];
        join8(__ldg(reinterpret_cast<const uint4*>(px)), __ldg(reinterpret_cast<const uint4*>(px + c)), f);
        const int k = (dy + 1) * 3 + dx + 1;
        const float4 wa = __ldg(reinterpret_cast<const float4*>(wgt + k * c + c8));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(wgt + k * c + c8 + 4));
        acc[0] = fmaf(f[0], wa.x, acc[0]); acc[1] = fmaf(f[1], wa.y, acc[1]);
        acc[2] = fmaf(f[2], wa.z, acc[2]); acc[3] = fmaf(f[3], wa.w, acc[3]);
        acc[4] = fmaf(f[4], wb.x, acc[4]); acc[5] = fmaf(f[5], wb.y, acc[5]);
        acc[6] = fmaf(f[6], wb.z, acc[6]); acc[7] = fmaf(f[7], wb.w, acc[7]);
      }
    }
    const long long o = ((1LL * s * H + Y) * W + X) * 2 * c + c8;
    if (skip) {
      float f[8];
      join8(__ldg(reinterpret_cast<const uint4*>(skip + o)), __ldg(reinterpret_cast<const uint4*>(skip + o + c)), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    uint4 hi, lo;
    split8(acc, hi, lo);
    *reinterpret_cast<uint4*>(out + o) = hi;
    *reinterpret_cast<uint4*>(out + o + c) = lo;
  }
}

// adaptive average pooling of the first c channels of a split tensor with pitch ld -> [n, bins, bins, 2c]
__global__ void __launch_bounds__(256)
adaptive_avgpool_split_kernel(const __nv_bfloat16* __restrict__ in, int h, int w, int c, int ld, int bins,
                              __nv_bfloat16* __restrict__ out) {
  __shared__ float s_part[32][64 + 4];
  const int by = blockIdx.x / bins, bx = blockIdx.x % bins, s = blockIdx.y;
  const int y0 = (by * h) / bins, y1 = ((by + 1) * h + bins - 1) / bins;
  const int x0 = (bx * w) / bins, x1 = ((bx + 1) * w + bins - 1) / bins;
  const int ww = x1 - x0, npix = (y1 - y0) * ww;
  const int oct = threadIdx.x & 7, lane_p = threadIdx.x >> 3;
  const int ch0 = blockIdx.z * 64 + oct * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ch0 < c) {
    for (int p = lane_p; p < npix; p += 32) {
      const int y = y0 + p / ww, x = x0 + p % ww;
      const __nv_bfloat16* px = in + ((1LL * s * h + y) * w + x) * ld + ch0;
      float f[8];
      join8(__ldg(reinterpret_cast<const uint4*>(px)), __ldg(reinterpret_cast<const uint4*>(px + (ld >> 1))), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_part[lane_p][oct * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int ch = blockIdx.z * 64 + threadIdx.x;
    if (ch < c) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 32; ++l) t += s_part[l][threadIdx.x];
      t = t / (float)npix;
      const __nv_bfloat16 hi = __float2bfloat16_rn(t);
      __nv_bfloat16* o = out + ((1LL * s * bins + by) * bins + bx) * 2 * c;
      o[ch] = hi;
      o[c + ch] = __float2bfloat16_rn(t - __bfloat162float(hi));
    }
  }
}

// nearest resize of a split tensor [n,hs,ws,2c] into channels [c_off, c_off + c) of both halves of dst (pitch ld)
__global__ void nearest_resize_into_split_kernel(const __nv_bfloat16* __restrict__ src, int n, int hs, int ws, int c,
                                                 __nv_bfloat16* __restrict__ dst, int h, int w, int ld, int c_off) {
  const int cv = c >> 3;
  const long long total = 1LL * n * h * w * cv * 2;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int half = (int)(i & 1);
    long long r = i >> 1;
    const int c8 = (int)(r % cv) * 8;
    r /= cv;
    const int x = (int)(r % w);
    r /= w;
    const int y = (int)(r % h);
    const int s = (int)(r / h);
    const int sy = min((int)((long long)y * hs / h), hs - 1), sx = min((int)((long long)x * ws / w), ws - 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + ((1LL * s * hs + sy) * ws + sx) * 2 * c + half * c + c8));
    *reinterpret_cast<uint4*>(dst + ((1LL * s * h + y) * w + x) * ld + half * (ld >> 1) + c_off + c8) = v;
  }
}

// F.interpolate(mode='bilinear', align_corners=False) of src [n,hs,ws,(2)c] into channels [c_off, c_off + c) of dst
// (pitch ld); split: both tensors are [hi | lo] halves, the interpolation runs on the reconstructed fp32 values
template <bool kSplit>
__global__ void bilinear_resize_into_kernel(const __nv_bfloat16* __restrict__ src, int n, int hs, int ws, int c,
                                            __nv_bfloat16* __restrict__ dst, int h, int w, int ld, int c_off) {
  const int cv = c >> 3;
  const int sld = kSplit ? 2 * c : c;
  const long long total = 1LL * n * h * w * cv;
  const float ry = (float)hs / (float)h, rx = (float)ws / (float)w;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv) * 8;
    long long r = i / cv;
    const int x = (int)(r % w);
    r /= w;
    const int y = (int)(r % h);
    const int s = (int)(r / h);
    // area_pixel_compute_source_index(scale, dst, align_corners=false): max(0, (dst + 0.5) * scale - 0.5)
    const float fy = fmaxf((y + 0.5f) * ry - 0.5f, 0.f), fx = fmaxf((x + 0.5f) * rx - 0.5f, 0.f);
    const int y0 = min((int)fy, hs - 1), x0 = min((int)fx, ws - 1);
    const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    float v[4][8];
    const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const __nv_bfloat16* px = src + ((1LL * s * hs + ys[a]) * ws + xs[b]) * sld + c8;
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(px));
        const uint4 lo = kSplit ? __ldg(reinterpret_cast<const uint4*>(px + c)) : make_uint4(0, 0, 0, 0);
        join8(hi, lo, v[a * 2 + b]);
      }
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = hy * (hx * v[0][e] + lx * v[1][e]) + ly * (hx * v[2][e] + lx * v[3][e]);
    uint4 hi, lo;
    split8(o, hi, lo);
    __nv_bfloat16* d = dst + ((1LL * s * h + y) * w + x) * ld + c_off + c8;
    *reinterpret_cast<uint4*>(d) = hi;
    if (kSplit) *reinterpret_cast<uint4*>(d + (ld >> 1)) = lo;
  }
}

}  // namespace
}  // namespace dynmm

using namespace dynmm;

extern "C" int dynmm_gated_add_fwd(const void* a, const void* b, const float* gate, const int32_t* slot, int n,
                                   long long per_sample, void* out, void* stream) {
  DYNMM_CHECK_ARG(a && b && gate && out && n >= 1 && per_sample >= 8 && per_sample % 8 == 0,
                  "gated_add: per_sample must be a positive multiple of 8");
  const long long vec = per_sample / 8;
  gated_add_bf16_kernel<<<grid_for(n * vec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), gate, slot, n, vec, static_cast<uint4*>(out));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_gated_add_f32_fwd(const float* a, const float* b, const float* gate, int n, long long per_sample,
                                       float* out, void* stream) {
  DYNMM_CHECK_ARG(a && b && gate && out && n >= 1 && per_sample >= 4 && per_sample % 4 == 0,
                  "gated_add_f32: per_sample must be a positive multiple of 4");
  const long long vec = per_sample / 4;
  gated_add_f32_fwd_kernel<<<grid_for(n * vec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), gate, n, vec,
      reinterpret_cast<float4*>(out));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

// grad_gate needs n * kBwdBlocks floats of scratch: it is carved from grad_gate's tail, so
// callers allocate grad_gate with n * (1 + 64) floats.
static const int kBwdBlocks = 64;
extern "C" int dynmm_gated_add_f32_bwd(const float* grad, const float* b, const float* gate, int n,
                                       long long per_sample, float* grad_b, float* grad_gate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(grad && b && gate && grad_gate && n >= 1 && per_sample >= 4 && per_sample % 4 == 0,
                  "gated_add_f32_bwd: per_sample must be a positive multiple of 4");
  const long long vec = per_sample / 4;
  float* partial = grad_gate + n;
  gated_add_f32_bwd_kernel<<<dim3(kBwdBlocks, n), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(grad), reinterpret_cast<const float4*>(b), gate, vec,
      reinterpret_cast<float4*>(grad_b), partial);
  DYNMM_LAUNCH_CHECK();
  reduce_partials_kernel<<<n, 32, 0, stream>>>(partial, kBwdBlocks, grad_gate);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_softgate_mix_fwd(const float* const* preds, const int32_t* const* rows, const float* w, int b,
                                      int c, int n_experts, float* out, void* stream) {
  DYNMM_CHECK_ARG(preds && w && out && b >= 1 && c >= 1 && n_experts >= 1 && n_experts <= 4, "softgate_mix: bad args");
  MixArgs a{};
  for (int e = 0; e < n_experts; ++e) {
    DYNMM_CHECK_ARG(preds[e], "softgate_mix: null expert output");
    a.pred[e] = preds[e];
    a.rows[e] = rows ? rows[e] : nullptr;
  }
  softgate_mix_fwd_kernel<<<grid_for(1LL * b * c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, w, b, c,
                                                                                                     n_experts, out);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_softgate_mix_bwd(const float* grad_out, const float* const* preds, const float* w, int b, int c,
                                      int n_experts, float* const* grad_preds, float* grad_w, void* stream) {
  DYNMM_CHECK_ARG(grad_out && preds && w && b >= 1 && c >= 1 && n_experts >= 1 && n_experts <= 4,
                  "softgate_mix_bwd: bad args");
  MixArgs a{};
  for (int e = 0; e < n_experts; ++e) {
    DYNMM_CHECK_ARG(preds[e], "softgate_mix_bwd: null expert output");
    a.pred[e] = preds[e];
    a.grad_pred[e] = grad_preds ? grad_preds[e] : nullptr;
  }
  softgate_mix_bwd_kernel<<<ceil_div(b, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(a, grad_out, w, b, c,
                                                                                         n_experts, grad_w);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_compact_rows(const float* w, int b, int n_experts, int expert, int32_t* idx, int32_t* inv,
                                  int32_t* count, void* stream) {
  DYNMM_CHECK_ARG(w && idx && inv && count && b >= 1 && expert >= 0 && expert < n_experts, "compact_rows: bad args");
  compact_rows_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, b, n_experts, expert, idx, inv, count);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_nchw_f32_to_nhwc_bf16(const float* in, int n, int c, int h, int w, void* out, void* stream) {
  DYNMM_CHECK_ARG(in && out && n >= 1 && c >= 1 && h >= 1 && w >= 1, "nchw_to_nhwc: bad args");
  const long long hw = 1LL * h * w;
  dim3 grid((unsigned)ceil_div_ll(hw, 32), ceil_div(c, 32), n);
  nchw_f32_to_nhwc_bf16_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      in, c, hw, static_cast<__nv_bfloat16*>(out));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_nhwc_bf16_to_nchw_f32(const void* in, int n, int c, int h, int w, int ld, float* out,
                                           void* stream) {
  DYNMM_CHECK_ARG(in && out && n >= 1 && c >= 1 && h >= 1 && w >= 1 && ld >= c, "nhwc_to_nchw: bad args");
  const long long hw = 1LL * h * w;
  dim3 grid((unsigned)ceil_div_ll(hw, 32), ceil_div(c, 32), n);
  nhwc_bf16_to_nchw_f32_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), c, hw, ld, out);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

namespace {
template <typename K>
int strip_rows(K kernel, int ctas_x, int h, int n) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  const long long slots = 1LL * per_sm * num_sms();
  int best = 1;
  long long best_cost = -1;
  for (int rows = 1; rows <= h && rows <= 32; ++rows) {
    const long long ctas = 1LL * ctas_x * ceil_div(h, rows) * n;
    const long long cost = ceil_div_ll(ctas, slots) * (rows + 2);
    if (best_cost < 0 || cost < best_cost) {      // ties: the shorter strip (more CTAs in the last wave)
      best_cost = cost;
      best = rows;
    }
  }
  return best;
}

int upsample2x_impl(const void* in, int n, int h, int w, int c, const float* weight, const float* bias, const void* skip,
                    void* out_nhwc_bf16, float* out_nchw_f32, uint8_t* labels, bool split, int clamp, int c_valid,
                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(in && weight && n >= 1 && h >= 1 && w >= 1 && c >= 8 && c % 8 == 0, "upsample2x: c %% 8");
  DYNMM_CHECK_ARG(!(out_nhwc_bf16 && (out_nchw_f32 || labels)) && (out_nhwc_bf16 || out_nchw_f32 || labels),
                  "upsample2x: either the NHWC output, or the NCHW logits and/or the arg-max labels");
  DYNMM_CHECK_ARG(!labels || c <= 256, "upsample2x: labels are uint8");
  if (c_valid <= 0) c_valid = c;
  DYNMM_CHECK_ARG(c_valid <= c && (c_valid == c || !out_nhwc_bf16), "upsample2x: c_valid applies to the NCHW / label outputs");
  if (out_nhwc_bf16 && split && n <= 65535 && getenv("DYNMM_UPSAMPLE") == nullptr) {
    const int ctas_x = ceil_div(w * (c / 4), 256);
    const int rows = skip ? strip_rows(upsample2x_dw_nhwc_strip_kernel<true, true>, ctas_x, h, n)
                          : strip_rows(upsample2x_dw_nhwc_strip_kernel<false, true>, ctas_x, h, n);
    dim3 grid(ctas_x, ceil_div(h, rows), n);
    if (skip) {
      upsample2x_dw_nhwc_strip_kernel<true, true><<<grid, 256, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(in), h, w, c, weight, bias, static_cast<const __nv_bfloat16*>(skip),
          static_cast<__nv_bfloat16*>(out_nhwc_bf16), rows, clamp);
    } else {
      upsample2x_dw_nhwc_strip_kernel<false, true><<<grid, 256, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(in), h, w, c, weight, bias, nullptr, static_cast<__nv_bfloat16*>(out_nhwc_bf16),
          rows, clamp);
    }
  } else if (out_nhwc_bf16 && split) {
    const long long total = 1LL * n * 4 * h * w * (c / 8);
    upsample2x_dw_nhwc_split_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), n, h, w, c, weight, bias, static_cast<const __nv_bfloat16*>(skip),
        static_cast<__nv_bfloat16*>(out_nhwc_bf16), clamp);
  } else if (out_nhwc_bf16) {
    static const bool use_strip = [] {
      const char* e = getenv("DYNMM_UPSAMPLE");
      return !(e && e[0] == 'p');           // DYNMM_UPSAMPLE=pixel: the one-thread-per-output-pixel kernel
    }();
    if (use_strip && n <= 65535) {
      const int ctas_x = ceil_div(w * (c / 4), 256);
      const int rows = skip ? strip_rows(upsample2x_dw_nhwc_strip_kernel<true, false>, ctas_x, h, n)
                            : strip_rows(upsample2x_dw_nhwc_strip_kernel<false, false>, ctas_x, h, n);
      dim3 grid(ctas_x, ceil_div(h, rows), n);
      if (skip) {
        upsample2x_dw_nhwc_strip_kernel<true><<<grid, 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(in), h, w, c, weight, bias, static_cast<const __nv_bfloat16*>(skip),
            static_cast<__nv_bfloat16*>(out_nhwc_bf16), rows, clamp);
      } else {
        upsample2x_dw_nhwc_strip_kernel<false><<<grid, 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(in), h, w, c, weight, bias, nullptr, static_cast<__nv_bfloat16*>(out_nhwc_bf16),
            rows, clamp);
      }
    } else {
      const long long total = 1LL * n * 4 * h * w * (c / 8);
      upsample2x_dw_nhwc_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(in), n, h, w, c, weight, bias, static_cast<const __nv_bfloat16*>(skip),
          static_cast<__nv_bfloat16*>(out_nhwc_bf16), clamp);
    }
  } else {
    DYNMM_CHECK_ARG(!skip, "upsample2x: skip is only supported for the NHWC output");
    const int smem = (c * (kUpTy + 2) * (kUpTx + 3) + c * 16) * (int)sizeof(float);
    DYNMM_CHECK_ARG(smem <= 200 * 1024, "upsample2x: too many channels for the NCHW output");
    static PerDeviceOnce attr_once;      // the limit checked above, so one setting serves every channel count
    DYNMM_CUDA(attr_once.run([] {
      cudaError_t e = cudaFuncSetAttribute(upsample2x_dw_to_nchw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(upsample2x_dw_to_nchw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      return e;
    }));
    dim3 grid(ceil_div(w, kUpTx), ceil_div(h, kUpTy), n);
    if (split)
      upsample2x_dw_to_nchw_kernel<true><<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(in), h, w, c, weight,
                                                                      bias, out_nchw_f32, labels, clamp, c_valid);
    else
      upsample2x_dw_to_nchw_kernel<false><<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(in), h, w, c,
                                                                       weight, bias, out_nchw_f32, labels, clamp, c_valid);
  }
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
}  // namespace

extern "C" int dynmm_upsample2x_dw3x3(const void* in, int n, int h, int w, int c, const float* weight,
                                      const float* bias, const void* skip, void* out_nhwc_bf16, float* out_nchw_f32,
                                      uint8_t* labels, void* stream) {
  return upsample2x_impl(in, n, h, w, c, weight, bias, skip, out_nhwc_bf16, out_nchw_f32, labels, false, 0, 0, stream);
}
extern "C" int dynmm_upsample2x_dw3x3_split(const void* in, int n, int h, int w, int c, const float* weight,
                                            const float* bias, const void* skip, void* out_nhwc_bf16, float* out_nchw_f32,
                                            uint8_t* labels, void* stream) {
  return upsample2x_impl(in, n, h, w, c, weight, bias, skip, out_nhwc_bf16, out_nchw_f32, labels, true, 0, 0, stream);
}
extern "C" int dynmm_upsample2x_dw3x3_ex(const void* in, int n, int h, int w, int c, const float* weight,
                                         const float* bias, const void* skip, void* out_nhwc_bf16, float* out_nchw_f32,
                                         uint8_t* labels, int flags, int c_valid, void* stream) {
  return upsample2x_impl(in, n, h, w, c, weight, bias, skip, out_nhwc_bf16, out_nchw_f32, labels,
                         (flags & DYNMM_UPSAMPLE_SPLIT) != 0, (flags & DYNMM_UPSAMPLE_REPLICATE) ? 1 : 0, c_valid, stream);
}
extern "C" int dynmm_bilinear_resize_into(const void* src, int n, int hs, int ws, int c, void* dst, int h, int w, int ld,
                                          int c_off, int split, void* stream) {
  DYNMM_CHECK_ARG(src && dst && c >= 8 && c % 8 == 0 && ld % (split ? 16 : 8) == 0 && c_off % 8 == 0 &&
                      c_off + c <= (split ? ld / 2 : ld),
                  "bilinear_resize: c/ld/c_off");
  const long long total = 1LL * n * h * w * (c / 8);
  if (split)
    bilinear_resize_into_kernel<true><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(src), n, hs, ws, c, static_cast<__nv_bfloat16*>(dst), h, w, ld, c_off);
  else
    bilinear_resize_into_kernel<false><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(src), n, hs, ws, c, static_cast<__nv_bfloat16*>(dst), h, w, ld, c_off);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
extern "C" int dynmm_split_from_f32(const float* x, long long rows, int c, void* out, void* stream) {
  DYNMM_CHECK_ARG(x && out && rows >= 1 && c >= 8 && c % 8 == 0, "split_from_f32: c %% 8");
  split_from_f32_kernel<<<grid_for(rows * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, c, static_cast<__nv_bfloat16*>(out));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
extern "C" int dynmm_adaptive_avgpool_split(const void* in, int n, int h, int w, int c, int ld, int bins, void* out,
                                            void* stream) {
  DYNMM_CHECK_ARG(in && out && n >= 1 && h >= 1 && w >= 1 && c >= 8 && c % 8 == 0 && ld % 16 == 0 && ld >= 2 * c &&
                      bins >= 1 && bins <= 64,
                  "avgpool_split: bad args");
  adaptive_avgpool_split_kernel<<<dim3(bins * bins, n, ceil_div(c, 64)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), h, w, c, ld, bins, static_cast<__nv_bfloat16*>(out));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
extern "C" int dynmm_nearest_resize_into_split(const void* src, int n, int hs, int ws, int c, void* dst, int h, int w,
                                               int ld, int c_off, void* stream) {
  DYNMM_CHECK_ARG(src && dst && c % 8 == 0 && ld % 16 == 0 && c_off % 8 == 0 && c_off + c <= ld / 2,
                  "resize_split: c/ld/c_off");
  const long long total = 1LL * n * h * w * (c / 8) * 2;
  nearest_resize_into_split_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), n, hs, ws, c, static_cast<__nv_bfloat16*>(dst), h, w, ld, c_off);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_adaptive_avgpool(const void* in, int n, int h, int w, int c, int ld, int bins, void* out,
                                      void* stream) {
  DYNMM_CHECK_ARG(in && out && n >= 1 && h >= 1 && w >= 1 && c >= 1 && ld >= c && bins >= 1 && bins <= 64,
                  "avgpool: bad args");
  if (c % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    adaptive_avgpool_v8_kernel<<<dim3(bins * bins, n, ceil_div(c, 64)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), h, w, c, ld, bins, static_cast<__nv_bfloat16*>(out));
  } else {
    adaptive_avgpool_kernel<<<dim3(bins * bins, n, ceil_div(c, 64)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), h, w, c, ld, bins, static_cast<__nv_bfloat16*>(out));
  }
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_nearest_resize_into(const void* src, int n, int hs, int ws, int c, void* dst, int h, int w, int ld,
                                         int c_off, void* stream) {
  DYNMM_CHECK_ARG(src && dst && c % 8 == 0 && ld % 8 == 0 && c_off % 8 == 0 && c_off + c <= ld,
                  "resize: c/ld/c_off %% 8");
  const long long total = 1LL * n * h * w * (c / 8);
  nearest_resize_into_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), n, hs, ws, c, static_cast<__nv_bfloat16*>(dst), h, w, ld, c_off);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
