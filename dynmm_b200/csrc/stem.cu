// Fused stem: conv7x7/s2 + BN + ReLU on RGB and on depth, rgb+depth add, and both
// 3x3/s2 max-pools in ONE kernel (resnet.py:352-358 + model_skip_mod_globalgate.py:256-261).
//
// Everything is fp32 on CUDA cores: K is only 147 / 49 (HBM/issue bound, not a
// tensor-core shape) and the pooled maps feed the gate, whose hard decisions
// must match the fp32 reference.  The two 64x240x320 stem maps never reach HBM:
// a CTA produces an 8x8 tile of POOLED outputs from a 17x17 tile of stem outputs
// held in shared memory (13 % halo recompute instead of a 2 x 157 MB round trip
// per 8 images).
#include <stdlib.h>

#include "common.cuh"

namespace dynmm {
namespace {

constexpr int kPT = 8;                    // pooled tile edge
constexpr int kST = 2 * kPT + 1;          // stem tile edge (17)
constexpr int kPatch = 2 * (kST - 1) + 7; // input patch edge (39)
constexpr int kPos = kST * kST;           // 289 stem positions
constexpr int kThreads = 256;
constexpr int kPosThreads = 64;           // threads along positions; 4 channel groups of 16
constexpr int kPosPerThread = (kPos + kPosThreads - 1) / kPosThreads;  // 5
constexpr int kKRgb = 7 * 7 * 3, kKDepth = 7 * 7;

constexpr size_t kSmemFloats = 2 * kPos * 64 + (kKRgb + kKDepth) * 64 + 4 * kPatch * (kPatch + 1) + 4 * 64;

// swizzled index into a [pos][64] fp32 tile: float4 column XOR (pos & 7)
__device__ __forceinline__ int tile_idx(int pos, int c) {
  return pos * 64 + ((((c >> 2) ^ (pos & 7)) << 2) | (c & 3));
}

// packed fp32x2 FMA (FFMA2): plain 3-register FFMA issues at half rate on sm_100, the packed
// form restores the full fp32 rate.  acc.xy += x * w.xy
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long xx, float wa, float wb) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(pack2(wa, wb)));
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// The input patch is stored de-interleaved by column parity: [ch][row][parity][kHalf], so that
// the stride-2 reads of a 7x7/s2 convolution (col = 2*sx + kx) are unit-stride across lanes.
constexpr int kHalf = (kPatch + 1) / 2;          // 20
constexpr int kRowStride = 2 * kHalf;            // 40 floats per patch row
constexpr int kChStride = kPatch * kRowStride;   // per input channel

template <int CIN>
__device__ __forceinline__ void conv_accumulate(const float* __restrict__ patch, const float* __restrict__ wsm,
                                                const int (&poff)[kPosPerThread], int cg,
                                                unsigned long long (&acc)[kPosPerThread][8]) {
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float4* w4 = reinterpret_cast<const float4*>(wsm + ((ky * 7 + kx) * CIN + ci) * 64 + cg * 16);
        const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
        const float* pp = patch + ci * kChStride + ky * kRowStride + (kx & 1) * kHalf + (kx >> 1);
#pragma unroll
        for (int i = 0; i < kPosPerThread; ++i) {
          const float x = pp[poff[i]];
          const unsigned long long xx = pack2(x, x);
          ffma2(acc[i][0], xx, w0.x, w0.y);
          ffma2(acc[i][1], xx, w0.z, w0.w);
          ffma2(acc[i][2], xx, w1.x, w1.y);
          ffma2(acc[i][3], xx, w1.z, w1.w);
          ffma2(acc[i][4], xx, w2.x, w2.y);
          ffma2(acc[i][5], xx, w2.z, w2.w);
          ffma2(acc[i][6], xx, w3.x, w3.y);
          ffma2(acc[i][7], xx, w3.z, w3.w);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
stem_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int H, int W,
            const float* __restrict__ w_rgb, const float* __restrict__ scale_rgb, const float* __restrict__ shift_rgb,
            const float* __restrict__ w_d, const float* __restrict__ scale_d, const float* __restrict__ shift_d,
            float* __restrict__ rgb_f32, float* __restrict__ depth_f32, __nv_bfloat16* __restrict__ rgb_bf16,
            __nv_bfloat16* __restrict__ depth_bf16, int tiles_x, int tiles_y, int batch,
            const float* __restrict__ se_rgb, const float* __restrict__ se_depth, float* __restrict__ gap_partial) {
  extern __shared__ __align__(16) float sm[];
  float* s_fuse = sm;                          // [289][64] swizzled: rgb stem, then rgb+depth
  float* s_dep = s_fuse + kPos * 64;           // [289][64] swizzled: depth stem
  float* s_wr = s_dep + kPos * 64;             // [147][64]
  float* s_wd = s_wr + kKRgb * 64;             // [49][64]
  float* s_patch = s_wd + kKDepth * 64;        // [4][39][2][20]  (r,g,b,depth), columns split by parity
  float* s_bn = s_patch + 4 * kPatch * (kPatch + 1);   // scale_rgb, shift_rgb, scale_d, shift_d

  const int Hs = (H + 2 * 3 - 7) / 2 + 1, Ws = (W + 2 * 3 - 7) / 2 + 1;   // stem map
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;         // pooled map
  const int tid = threadIdx.x;
  const int total_tiles = tiles_x * tiles_y * batch;

  // persistent CTA: the 50 KB of weights and the BN constants are staged ONCE
  for (int i = tid; i < kKRgb * 16; i += kThreads)
    reinterpret_cast<float4*>(s_wr)[i] = __ldg(reinterpret_cast<const float4*>(w_rgb) + i);
  for (int i = tid; i < kKDepth * 16; i += kThreads)
    reinterpret_cast<float4*>(s_wd)[i] = __ldg(reinterpret_cast<const float4*>(w_d) + i);
  if (tid < 64) {
    s_bn[tid] = scale_rgb[tid];
    s_bn[64 + tid] = shift_rgb[tid];
    s_bn[128 + tid] = scale_d[tid];
    s_bn[192 + tid] = shift_d[tid];
  }

  // input patch: element e = tid + 256*j of the [4][39][39] patch; fetched into registers (so the next
  // tile's global loads overlap this tile's pooling) and scattered into the parity-split layout
  constexpr int kPatchElems = 4 * kPatch * kPatch;
  constexpr int kPerThread = (kPatchElems + kThreads - 1) / kThreads;     // 24
  float pre[kPerThread];
  auto fetch_patch = [&](int tile) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, nn = tile / (tiles_x * tiles_y);
    const int iy0 = 2 * (2 * ty * kPT - 1) - 3, ix0 = 2 * (2 * tx * kPT - 1) - 3;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int e = tid + j * kThreads;
      float v = 0.f;
      if (e < kPatchElems) {
        const int ch = e / (kPatch * kPatch);
        const int r = e - ch * (kPatch * kPatch);
        const int py = r / kPatch, px = r - py * kPatch;
        const int y = iy0 + py, x = ix0 + px;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          v = ch < 3 ? __ldg(rgb + ((static_cast<size_t>(nn) * 3 + ch) * H + y) * W + x)
                     : __ldg(depth + (static_cast<size_t>(nn) * H + y) * W + x);
        }
      }
      pre[j] = v;
    }
  };
  auto store_patch = [&]() {
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int e = tid + j * kThreads;
      if (e < kPatchElems) {
        const int ch = e / (kPatch * kPatch);
        const int r = e - ch * (kPatch * kPatch);
        const int py = r / kPatch, px = r - py * kPatch;
        s_patch[ch * kChStride + py * kRowStride + (px & 1) * kHalf + (px >> 1)] = pre[j];
      }
    }
  };
  if ((int)blockIdx.x < total_tiles) {
    fetch_patch(blockIdx.x);
    store_patch();
  }
  __syncthreads();

  const int cg = tid / kPosThreads;       // warp-uniform channel group: weight loads broadcast
  const int tpos = tid % kPosThreads;
  int poff[kPosPerThread];
  int pos[kPosPerThread];
#pragma unroll
  for (int i = 0; i < kPosPerThread; ++i) {
    int p = tpos + i * kPosThreads;
    pos[i] = p;
    if (p >= kPos) p = kPos - 1;          // clamp: computed but never stored
    poff[i] = (2 * (p / kST)) * kRowStride + (p % kST);
  }

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int py0 = ((tile / tiles_x) % tiles_y) * kPT, px0 = (tile % tiles_x) * kPT;
    const int sy0 = 2 * py0 - 1, sx0 = 2 * px0 - 1;     // stem-tile origin
    unsigned long long acc2[kPosPerThread][8];
    float acc[kPosPerThread][16];
    // ---- RGB stem conv
  #pragma unroll
    for (int i = 0; i < kPosPerThread; ++i)
  #pragma unroll
      for (int c = 0; c < 8; ++c) acc2[i][c] = 0ull;
    conv_accumulate<3>(s_patch, s_wr, poff, cg, acc2);
  #pragma unroll
    for (int i = 0; i < kPosPerThread; ++i)
  #pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 t = unpack2(acc2[i][c]);
        acc[i][2 * c] = t.x;
        acc[i][2 * c + 1] = t.y;
      }
  #pragma unroll
    for (int i = 0; i < kPosPerThread; ++i) {
      if (pos[i] < kPos) {
  #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = cg * 16 + q * 4;
          float4 v;
          v.x = fmaxf(fmaf(acc[i][q * 4 + 0], s_bn[c + 0], s_bn[64 + c + 0]), 0.f);
          v.y = fmaxf(fmaf(acc[i][q * 4 + 1], s_bn[c + 1], s_bn[64 + c + 1]), 0.f);
          v.z = fmaxf(fmaf(acc[i][q * 4 + 2], s_bn[c + 2], s_bn[64 + c + 2]), 0.f);
          v.w = fmaxf(fmaf(acc[i][q * 4 + 3], s_bn[c + 3], s_bn[64 + c + 3]), 0.f);
          *reinterpret_cast<float4*>(&s_fuse[tile_idx(pos[i], c)]) = v;
        }
      }
    }
    // ---- depth stem conv, then fuse = rgb + depth (same thread owns the same elements)
  #pragma unroll
    for (int i = 0; i < kPosPerThread; ++i)
  #pragma unroll
      for (int c = 0; c < 8; ++c) acc2[i][c] = 0ull;
    conv_accumulate<1>(s_patch + 3 * kChStride, s_wd, poff, cg, acc2);
  #pragma unroll
    for (int i = 0; i < kPosPerThread; ++i)
  #pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 t = unpack2(acc2[i][c]);
        acc[i][2 * c] = t.x;
        acc[i][2 * c + 1] = t.y;
      }
  #pragma unroll
    for (int i = 0; i < kPosPerThread; ++i) {
      if (pos[i] < kPos) {
  #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = cg * 16 + q * 4;
          float4 d;
          d.x = fmaxf(fmaf(acc[i][q * 4 + 0], s_bn[128 + c + 0], s_bn[192 + c + 0]), 0.f);
          d.y = fmaxf(fmaf(acc[i][q * 4 + 1], s_bn[128 + c + 1], s_bn[192 + c + 1]), 0.f);
          d.z = fmaxf(fmaf(acc[i][q * 4 + 2], s_bn[128 + c + 2], s_bn[192 + c + 2]), 0.f);
          d.w = fmaxf(fmaf(acc[i][q * 4 + 3], s_bn[128 + c + 3], s_bn[192 + c + 3]), 0.f);
          const int idx = tile_idx(pos[i], c);
          float4 r = *reinterpret_cast<float4*>(&s_fuse[idx]);
          if (se_rgb) {
            // SqueezeAndExciteFusionAdd (rgb_depth_fusion.py:22-26): rgb*sigma_r + depth*sigma_d
            const float4 sr = __ldg(reinterpret_cast<const float4*>(se_rgb + n * 64 + c));
            const float4 sd = __ldg(reinterpret_cast<const float4*>(se_depth + n * 64 + c));
            r.x = r.x * sr.x + d.x * sd.x; r.y = r.y * sr.y + d.y * sd.y;
            r.z = r.z * sr.z + d.z * sd.z; r.w = r.w * sr.w + d.w * sd.w;
          } else if (!gap_partial) {
            r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
          }
          *reinterpret_cast<float4*>(&s_fuse[idx]) = r;
          *reinterpret_cast<float4*>(&s_dep[idx]) = d;
        }
      }
    }
    __syncthreads();

    // the patch is free again: start fetching the next tile's pixels while we pool this one
    const int next_tile = tile + gridDim.x;
    if (next_tile < total_tiles) fetch_patch(next_tile);

    const int c = tid & 63;
    if (gap_partial) {
      // squeeze pass: per-tile channel sums of the UNFUSED stem maps over the positions this tile owns
      // (row/column 0 of the 17x17 tile belong to the neighbouring tile); fixed order -> deterministic
      float sr = 0.f, sd = 0.f;
      for (int p = tid >> 6; p < kPos; p += kThreads >> 6) {
        const int ly = p / kST, lx = p % kST;
        const int gy = sy0 + ly, gx = sx0 + lx;
        if (ly >= 1 && lx >= 1 && gy < Hs && gx < Ws) {
          sr += s_fuse[tile_idx(p, c)];
          sd += s_dep[tile_idx(p, c)];
        }
      }
      __syncthreads();                      // tiles fully read; reuse the head of s_fuse as scratch
      s_fuse[(tid >> 6) * 128 + c] = sr;
      s_fuse[(tid >> 6) * 128 + 64 + c] = sd;
      __syncthreads();
      if (tid < 128) {
        const float t = (s_fuse[tid] + s_fuse[128 + tid]) + (s_fuse[256 + tid] + s_fuse[384 + tid]);
        gap_partial[static_cast<size_t>(tile) * 128 + tid] = t;      // [tile][rgb 64 | depth 64]
      }
      if (next_tile < total_tiles) store_patch();
      __syncthreads();
      continue;
    }
    // ---- 3x3 / stride 2 / pad 1 max-pool of both tiles, NHWC stores (64 consecutive channels per pixel)
    for (int pp = tid >> 6; pp < kPT * kPT; pp += kThreads >> 6) {
      const int ly = pp / kPT, lx = pp % kPT;
      const int py = py0 + ly, px = px0 + lx;
      if (py >= Hp || px >= Wp) continue;
      float mf = -INFINITY, md = -INFINITY;
  #pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int gy = sy0 + 2 * ly + dy;
        if (gy < 0 || gy >= Hs) continue;
  #pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int gx = sx0 + 2 * lx + dx;
          if (gx < 0 || gx >= Ws) continue;
          const int idx = tile_idx((2 * ly + dy) * kST + 2 * lx + dx, c);
          mf = fmaxf(mf, s_fuse[idx]);
          md = fmaxf(md, s_dep[idx]);
        }
      }
      const size_t o = ((static_cast<size_t>(n) * Hp + py) * Wp + px) * 64 + c;
      if (rgb_f32) rgb_f32[o] = mf;
      if (depth_f32) depth_f32[o] = md;
      if (rgb_bf16) rgb_bf16[o] = __float2bfloat16_rn(mf);
      if (depth_bf16) depth_bf16[o] = __float2bfloat16_rn(md);
    }
    if (next_tile < total_tiles) store_patch();
    __syncthreads();     // tiles of the next iteration may be overwritten; its patch is in place
  }
}

}  // namespace
}  // namespace dynmm

namespace dynmm {
// tensor-core implementation (stem_tc.cu)
long long stem_tc_tiles(int b, int h, int w);
int stem_tc_launch(const float* rgb, const float* depth, int b, int h, int w, const float* w_rgb, const float* scale_rgb,
                   const float* shift_rgb, const float* w_d, const float* scale_d, const float* shift_d,
                   float* rgb_f32, float* depth_f32, void* rgb_bf16, void* depth_bf16, const float* se_rgb,
                   const float* se_depth, float* gap_partial, cudaStream_t stream);
// DYNMM_STEM=fp32 selects the CUDA-core FFMA2 kernel of this file (kept as the exact-fp32 variant and as
// an independent cross-check of the tensor-core kernel); default is the tensor-core split-bf16 kernel.
static bool stem_use_tc() {
  static const bool v = [] {
    const char* e = getenv("DYNMM_STEM");
    return !(e && (e[0] == 'f' || e[0] == 'F'));
  }();
  return v;
}
}  // namespace dynmm

extern "C" long long dynmm_stem_gap_tiles(int b, int h, int w) {
  if (dynmm::stem_use_tc()) return dynmm::stem_tc_tiles(b, h, w);
  const int Hs = (h + 6 - 7) / 2 + 1, Ws = (w + 6 - 7) / 2 + 1;
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;
  return 1LL * dynmm::ceil_div(Wp, dynmm::kPT) * dynmm::ceil_div(Hp, dynmm::kPT) * b;
}

extern "C" int dynmm_stem_fwd(const float* rgb, const float* depth, int b, int h, int w, const float* w_rgb,
                              const float* scale_rgb, const float* shift_rgb, const float* w_d, const float* scale_d,
                              const float* shift_d, float* rgb_f32, float* depth_f32, void* rgb_bf16,
                              void* depth_bf16, const float* se_rgb, const float* se_depth, float* gap_partial,
                              void* stream) {
  using namespace dynmm;
  DYNMM_CHECK_ARG(rgb && depth && w_rgb && w_d && scale_rgb && shift_rgb && scale_d && shift_d, "stem: null pointer");
  DYNMM_CHECK_ARG(b >= 1 && h >= 7 && w >= 7, "stem: bad shape");
  DYNMM_CHECK_ARG((se_rgb == nullptr) == (se_depth == nullptr), "stem: SE scales come in pairs");
  if (stem_use_tc()) {
    return stem_tc_launch(rgb, depth, b, h, w, w_rgb, scale_rgb, shift_rgb, w_d, scale_d, shift_d, rgb_f32, depth_f32,
                          rgb_bf16, depth_bf16, se_rgb, se_depth, gap_partial, static_cast<cudaStream_t>(stream));
  }
  const int Hs = (h + 6 - 7) / 2 + 1, Ws = (w + 6 - 7) / 2 + 1;
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;
  const size_t smem = kSmemFloats * sizeof(float);
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run([] {
    return cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemFloats * sizeof(float)));
  }));
  const int tiles_x = ceil_div(Wp, kPT), tiles_y = ceil_div(Hp, kPT);
  const long long total = 1LL * tiles_x * tiles_y * b;
  DYNMM_CHECK_ARG(total < (1LL << 30), "stem: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  stem_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      rgb, depth, h, w, w_rgb, scale_rgb, shift_rgb, w_d, scale_d, shift_d, rgb_f32, depth_f32,
      static_cast<__nv_bfloat16*>(rgb_bf16), static_cast<__nv_bfloat16*>(depth_bf16), tiles_x, tiles_y, b, se_rgb,
      se_depth, gap_partial);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
