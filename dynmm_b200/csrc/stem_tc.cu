// Tensor-core stem: the same fused op as stem.cu (conv7x7/s2 + BN + ReLU on RGB and depth, add or
// SE-weighted add, two 3x3/s2 max-pools; resnet.py:352-358 + model_skip_mod_globalgate.py:256-261)
// with the two convolutions on tcgen05.
//
// fp32-grade accuracy from bf16 tensor cores: every fp32 operand is split x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi), and  x*w ~= hi*w_hi + hi*w_lo + lo*w_hi  (three UMMAs into one
// fp32 TMEM accumulator; the dropped lo*lo term is 2^-16 relative).  The stem feeds the gate, whose
// hard decisions must equal the fp32 reference's: measured error of the pooled maps ~1e-5.
//
// A 7x7x3 stride-2 convolution has no TMA-friendly im2col (3 channels = 6 bytes), so the CTA builds
// the im2col tile itself: 121 stem positions (11x11, for a 5x5 tile of pooled outputs) x K=147 (+49
// for depth) are gathered from a staged fp32 input patch, split, and written straight into the UMMA
// K-major 128B-swizzled shared-memory layout (lanes run along K, so the stores are conflict-free).
// One thread then issues 42 UMMAs (128 x 64 x 16); the accumulators come back through tcgen05.ld,
// get BN + ReLU (+ SE scales), are fused, and land in shared memory for the max-pool -- over the A
// operand, which is dead by then.  Persistent CTAs: weights are split and swizzled once.
#include "common.cuh"

namespace dynmm {
namespace stemtc {

constexpr int kPT = 5;                      // pooled tile edge
constexpr int kST = 2 * kPT + 1;            // stem tile edge (11)
constexpr int kPos = kST * kST;             // 121 stem positions = GEMM rows (of 128)
constexpr int kPatch = 2 * (kST - 1) + 7;   // 27
constexpr int kPRow = kPatch + 1;           // 28: patch row stride
constexpr int kPCh = kPatch * kPRow;        // per input channel
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kKRgb = 147, kKDep = 49;
constexpr int kChunk = 128 * 128;           // one [128 rows][64 bf16] operand chunk (16 KiB)
constexpr int kWChunk = 64 * 128;           // one [64 cout][64 bf16] weight chunk (8 KiB)
// A region: rgb_hi[3] rgb_lo[3] dep_hi dep_lo ; W region: same order
constexpr int kARgbHi = 0, kARgbLo = 3 * kChunk, kADepHi = 6 * kChunk, kADepLo = 7 * kChunk, kABytes = 8 * kChunk;
constexpr int kWRgbHi = 0, kWRgbLo = 3 * kWChunk, kWDepHi = 6 * kWChunk, kWDepLo = 7 * kWChunk, kWBytes = 8 * kWChunk;
constexpr int kPatchBytes = 4 * kPCh * 4;
constexpr int kSmemBytes = 1024 + kABytes + kWBytes + kPatchBytes + 256 * 4 + 64;

__device__ __forceinline__ int tile_idx(int pos, int c) {      // swizzled [pos][64] fp32 tile (as in stem.cu)
  return pos * 64 + ((((c >> 2) ^ (pos & 7)) << 2) | (c & 3));
}
// byte offset of element (row, k) inside a K-major 128B-swizzled chunk sequence
__device__ __forceinline__ uint32_t sw_off(int row, int k, int chunk_bytes) {
  const int chunk = k >> 6, e = k & 63;
  return chunk * chunk_bytes + row * 128 + ((((e >> 3) ^ (row & 7)) << 4) | ((e & 7) << 1));
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
  const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

__global__ void __launch_bounds__(kThreads, 1)
stem_tc_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int H, int W,
               const float* __restrict__ w_rgb, const float* __restrict__ scale_rgb,
               const float* __restrict__ shift_rgb, const float* __restrict__ w_d,
               const float* __restrict__ scale_d, const float* __restrict__ shift_d, float* __restrict__ rgb_f32,
               float* __restrict__ depth_f32, __nv_bfloat16* __restrict__ rgb_bf16,
               __nv_bfloat16* __restrict__ depth_bf16, int tiles_x, int tiles_y, int batch,
               const float* __restrict__ se_rgb, const float* __restrict__ se_depth, float* __restrict__ gap_partial) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                                   // operand A chunks; later the fp32 output tiles
  uint8_t* s_w = s_a + kABytes;                          // split weights
  float* s_patch = reinterpret_cast<float*>(s_w + kWBytes);   // [4][27][28]
  float* s_bn = s_patch + 4 * kPCh;                      // scale_rgb, shift_rgb, scale_d, shift_d
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bn + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);
  float* s_fuse = reinterpret_cast<float*>(s_a);         // [121][64] swizzled (aliases A after the MMAs)
  float* s_dep = s_fuse + kPos * 64;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Hs = (H + 2 * 3 - 7) / 2 + 1, Ws = (W + 2 * 3 - 7) / 2 + 1;
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;
  const int total_tiles = tiles_x * tiles_y * batch;

  // ---- once per CTA: split + swizzle the weights.  w_rgb is [k][64] fp32 (k = (ky*7+kx)*3+ci); the B
  // operand is K-major [cout][k], zero padded to 192 / 64.
  for (int i = tid; i < 64 * 96; i += kThreads) {          // (cout, k-pair) of the RGB weights
    const int n = i / 96, k0 = (i % 96) * 2;
    const float x0 = k0 < kKRgb ? w_rgb[k0 * 64 + n] : 0.f, x1 = k0 + 1 < kKRgb ? w_rgb[(k0 + 1) * 64 + n] : 0.f;
    uint32_t hi, lo;
    split2(x0, x1, hi, lo);
    const uint32_t off = sw_off(n, k0, kWChunk);
    *reinterpret_cast<uint32_t*>(s_w + kWRgbHi + off) = hi;
    *reinterpret_cast<uint32_t*>(s_w + kWRgbLo + off) = lo;
  }
  for (int i = tid; i < 64 * 32; i += kThreads) {
    const int n = i / 32, k0 = (i % 32) * 2;
    const float x0 = k0 < kKDep ? w_d[k0 * 64 + n] : 0.f, x1 = k0 + 1 < kKDep ? w_d[(k0 + 1) * 64 + n] : 0.f;
    uint32_t hi, lo;
    split2(x0, x1, hi, lo);
    const uint32_t off = sw_off(n, k0, kWChunk);
    *reinterpret_cast<uint32_t*>(s_w + kWDepHi + off) = hi;
    *reinterpret_cast<uint32_t*>(s_w + kWDepLo + off) = lo;
  }
  if (tid < 64) {
    s_bn[tid] = scale_rgb[tid];
    s_bn[64 + tid] = shift_rgb[tid];
    s_bn[128 + tid] = scale_d[tid];
    s_bn[192 + tid] = shift_d[tid];
  }
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem, 128);          // columns [0,64): RGB accumulator, [64,128): depth accumulator
    tmem_relinquish();
  }
  // rows 121..127 of every A chunk are never written: clear them once so the unused MMA rows stay finite
  for (int i = tid; i < 8 * 7 * 32; i += kThreads) {
    const int chunk = i / (7 * 32), r = kPos + (i / 32) % 7, wd = i % 32;
    reinterpret_cast<uint32_t*>(s_a + chunk * kChunk + r * 128)[wd] = 0u;
  }

  // ---- per-lane im2col schedule (lanes run along K): RGB k-pairs q = lane + 32 j (j < 3, k = 2q, 2q+1 < 160),
  // depth k-pair q = lane (k < 64).  poff = offset of the tap inside the patch, < 0 for zero padding of K.
  int rgb_off[3][2], dep_off[2];
  uint32_t rgb_sw[3], dep_sw;           // byte offset of the pair inside a chunk row, before the row swizzle
  int rgb_chunk[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int k0 = 2 * (lane + 32 * j);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = k0 + e;
      rgb_off[j][e] = k < kKRgb ? (k % 3) * kPCh + (k / 21) * kPRow + (k % 21) / 3 : -1;
    }
    rgb_chunk[j] = k0 >> 6;
    rgb_sw[j] = k0 & 63;
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = 2 * lane + e;
    dep_off[e] = k < kKDep ? 3 * kPCh + (k / 7) * kPRow + (k % 7) : -1;
  }
  dep_sw = (2 * lane) & 63;

  // ---- input patch prefetch (registers), as in stem.cu
  constexpr int kPatchElems = 4 * kPatch * kPatch;
  constexpr int kPerThread = (kPatchElems + kThreads - 1) / kThreads;     // 6
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
        s_patch[ch * kPCh + py * kPRow + px] = pre[j];
      }
    }
  };
  if ((int)blockIdx.x < total_tiles) {
    fetch_patch(blockIdx.x);
    store_patch();
  }
  fence_async_smem();            // weights + cleared rows -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
  const uint32_t idesc = umma_idesc_bf16(128, 64);
  uint32_t mma_phase = 0;

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int py0 = ((tile / tiles_x) % tiles_y) * kPT, px0 = (tile % tiles_x) * kPT;
    const int sy0 = 2 * py0 - 1, sx0 = 2 * px0 - 1;     // stem-tile origin

    // ---- im2col + split into the swizzled A chunks: one warp per GEMM row, lanes along K
    for (int r = warp; r < kPos; r += kWarps) {
      const int pbase = (2 * (r / kST)) * kPRow + 2 * (r % kST);
      const uint32_t row_base = r * 128;
      const uint32_t rsw = (r & 7) << 4;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j == 2 && lane >= 16) break;               // k >= 160 is not part of any MMA k-step
        const float x0 = rgb_off[j][0] >= 0 ? s_patch[rgb_off[j][0] + pbase] : 0.f;
        const float x1 = rgb_off[j][1] >= 0 ? s_patch[rgb_off[j][1] + pbase] : 0.f;
        uint32_t hi, lo;
        split2(x0, x1, hi, lo);
        const uint32_t off = rgb_chunk[j] * kChunk + row_base + ((((rgb_sw[j] >> 3) << 4) ^ rsw) | ((rgb_sw[j] & 7) << 1));
        *reinterpret_cast<uint32_t*>(s_a + kARgbHi + off) = hi;
        *reinterpret_cast<uint32_t*>(s_a + kARgbLo + off) = lo;
      }
      {
        const float x0 = dep_off[0] >= 0 ? s_patch[dep_off[0] + pbase] : 0.f;
        const float x1 = dep_off[1] >= 0 ? s_patch[dep_off[1] + pbase] : 0.f;
        uint32_t hi, lo;
        split2(x0, x1, hi, lo);
        const uint32_t off = row_base + ((((dep_sw >> 3) << 4) ^ rsw) | ((dep_sw & 7) << 1));
        *reinterpret_cast<uint32_t*>(s_a + kADepHi + off) = hi;
        *reinterpret_cast<uint32_t*>(s_a + kADepLo + off) = lo;
      }
    }
    fence_async_smem();
    __syncthreads();

    // the patch is consumed: prefetch the next tile's pixels while the tensor cores and the epilogue run
    const int next_tile = tile + gridDim.x;
    if (next_tile < total_tiles) fetch_patch(next_tile);

    // ---- 42 UMMAs: D_rgb = A_rgb W_rgb^T (10 k-steps), D_dep = A_dep W_dep^T (4 k-steps), 3 split products each
    if (tid == 32) {
      tc_fence_after();
      bool first = true;
      for (int c = 0; c < 3; ++c) {
        const int ksteps = c < 2 ? 4 : 2;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ah = umma_desc_sw128(a_base + kARgbHi + c * kChunk) + 2 * k;
          const uint64_t al = umma_desc_sw128(a_base + kARgbLo + c * kChunk) + 2 * k;
          const uint64_t wh = umma_desc_sw128(w_base + kWRgbHi + c * kWChunk) + 2 * k;
          const uint64_t wl = umma_desc_sw128(w_base + kWRgbLo + c * kWChunk) + 2 * k;
          umma_bf16(tmem_base, ah, wh, idesc, first ? 0u : 1u);
          umma_bf16(tmem_base, ah, wl, idesc, 1u);
          umma_bf16(tmem_base, al, wh, idesc, 1u);
          first = false;
        }
      }
      for (int k = 0; k < 4; ++k) {
        const uint64_t ah = umma_desc_sw128(a_base + kADepHi) + 2 * k;
        const uint64_t al = umma_desc_sw128(a_base + kADepLo) + 2 * k;
        const uint64_t wh = umma_desc_sw128(w_base + kWDepHi) + 2 * k;
        const uint64_t wl = umma_desc_sw128(w_base + kWDepLo) + 2 * k;
        umma_bf16(tmem_base + 64, ah, wh, idesc, k ? 1u : 0u);
        umma_bf16(tmem_base + 64, ah, wl, idesc, 1u);
        umma_bf16(tmem_base + 64, al, wh, idesc, 1u);
      }
      umma_commit(s_bar);
      // only this thread polls the mbarrier; everybody else parks at the hardware barrier below (no
      // diverged sibling lanes spinning next to the single MMA-issuing lane)
      mbar_wait(s_bar, mma_phase);
    }
    mma_phase ^= 1;
    __syncthreads();
    tc_fence_after();

    // ---- epilogue: 16 warps = 4 lane quarters x 4 column groups of 16 channels.  BN + ReLU, fuse, -> smem tiles
    {
      const int quarter = warp & 3, cgp = warp >> 2;          // TMEM lanes 32*quarter.., channels 16*cgp..
      const int row = quarter * 32 + lane;
      uint32_t vr[16], vd[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + cgp * 16;
      tmem_ld16(taddr, vr);
      tmem_ld16(taddr + 64, vd);
      tmem_ld_wait();
      tc_fence_before();
      __syncthreads();                 // every warp has its accumulators: the A region may now be overwritten
      if (row < kPos) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = cgp * 16 + q * 4;
          float4 r, d;
          r.x = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 0]), s_bn[c + 0], s_bn[64 + c + 0]), 0.f);
          r.y = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 1]), s_bn[c + 1], s_bn[64 + c + 1]), 0.f);
          r.z = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 2]), s_bn[c + 2], s_bn[64 + c + 2]), 0.f);
          r.w = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 3]), s_bn[c + 3], s_bn[64 + c + 3]), 0.f);
          d.x = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 0]), s_bn[128 + c + 0], s_bn[192 + c + 0]), 0.f);
          d.y = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 1]), s_bn[128 + c + 1], s_bn[192 + c + 1]), 0.f);
          d.z = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 2]), s_bn[128 + c + 2], s_bn[192 + c + 2]), 0.f);
          d.w = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 3]), s_bn[128 + c + 3], s_bn[192 + c + 3]), 0.f);
          if (se_rgb) {
            // SqueezeAndExciteFusionAdd (rgb_depth_fusion.py:22-26): rgb*sigma_r + depth*sigma_d
            const float4 sr = __ldg(reinterpret_cast<const float4*>(se_rgb + n * 64 + c));
            const float4 sd = __ldg(reinterpret_cast<const float4*>(se_depth + n * 64 + c));
            r.x = r.x * sr.x + d.x * sd.x; r.y = r.y * sr.y + d.y * sd.y;
            r.z = r.z * sr.z + d.z * sd.z; r.w = r.w * sr.w + d.w * sd.w;
          } else if (!gap_partial) {
            r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
          }
          const int idx = tile_idx(row, c);
          *reinterpret_cast<float4*>(&s_fuse[idx]) = r;
          *reinterpret_cast<float4*>(&s_dep[idx]) = d;
        }
      }
    }
    __syncthreads();

    const int c = tid & 63;
    if (gap_partial) {
      // squeeze pass (SE-add): per-tile channel sums of the UNFUSED stem maps over the positions this tile
      // owns (row/column 0 of the 11x11 tile belong to the neighbouring tile); fixed order -> deterministic
      float sr = 0.f, sd = 0.f;
      for (int p = tid >> 6; p < kPos; p += kThreads >> 6) {
        const int ly = p / kST, lx = p % kST;
        const int gy = sy0 + ly, gx = sx0 + lx;
        if (ly >= 1 && lx >= 1 && gy < Hs && gx < Ws) {
          sr += s_fuse[tile_idx(p, c)];
          sd += s_dep[tile_idx(p, c)];
        }
      }
      __syncthreads();
      float* scratch = s_patch;               // 8 x 128 floats; the patch was consumed by the im2col
      scratch[(tid >> 6) * 128 + c] = sr;
      scratch[(tid >> 6) * 128 + 64 + c] = sd;
      __syncthreads();
      if (tid < 128) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < kThreads / 64; ++g) t += scratch[g * 128 + tid];
        gap_partial[static_cast<size_t>(tile) * 128 + tid] = t;      // [tile][rgb 64 | depth 64]
      }
      __syncthreads();
    } else {
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
    }
    if (next_tile < total_tiles) store_patch();
    __syncthreads();     // output tiles fully consumed (they alias A); the next patch is in place
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace stemtc
}  // namespace dynmm

namespace dynmm {

long long stem_tc_tiles(int b, int h, int w) {
  const int Hs = (h + 6 - 7) / 2 + 1, Ws = (w + 6 - 7) / 2 + 1;
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;
  return 1LL * ceil_div(Wp, stemtc::kPT) * ceil_div(Hp, stemtc::kPT) * b;
}

int stem_tc_launch(const float* rgb, const float* depth, int b, int h, int w, const float* w_rgb, const float* scale_rgb,
                   const float* shift_rgb, const float* w_d, const float* scale_d, const float* shift_d,
                   float* rgb_f32, float* depth_f32, void* rgb_bf16, void* depth_bf16, const float* se_rgb,
                   const float* se_depth, float* gap_partial, cudaStream_t stream) {
  using namespace stemtc;
  const int Hs = (h + 6 - 7) / 2 + 1, Ws = (w + 6 - 7) / 2 + 1;
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run(
      [] { return cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes); }));
  const int tiles_x = ceil_div(Wp, kPT), tiles_y = ceil_div(Hp, kPT);
  const long long total = 1LL * tiles_x * tiles_y * b;
  DYNMM_CHECK_ARG(total < (1LL << 30), "stem: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  stem_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(
      rgb, depth, h, w, w_rgb, scale_rgb, shift_rgb, w_d, scale_d, shift_d, rgb_f32, depth_f32,
      static_cast<__nv_bfloat16*>(rgb_bf16), static_cast<__nv_bfloat16*>(depth_bf16), tiles_x, tiles_y, b, se_rgb,
      se_depth, gap_partial);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

}  // namespace dynmm
