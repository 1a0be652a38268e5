// Stem without an in-SM im2col: conv7x7/s2 + BN + ReLU on RGB and depth, add, two 3x3/s2 max-pools
// (resnet.py:352-358 + model_skip_mod_globalgate.py:256-261), the two convolutions as ONE tcgen05 GEMM whose
// A operand is gathered by TMA.
//
// stem_tc.cu builds the im2col tile with ordinary instructions (147 + 49 taps per stem pixel, split into
// bf16 hi/lo halves, 8 x 128 rows x 128 B of swizzled stores per tile) and is bound by their issue rate
// (446 us at batch 8).  Here a light pre-pass rewrites the 4-channel input as a space-to-depth image
//
//     P[n][y2 + 2][x2 + 2][16] = { rgb(c, 2 y2 + dy, 2 x2 + dx) : (dy, dx, c) } (12)  ++  { depth(2 y2 + dy, 2 x2 + dx) } (4)
//
// (bf16 hi and lo planes, zero border of 2 / 1 pixels), in which the 7x7 stride-2 convolution is a 4x4
// unit-stride convolution over 16 channels: for a kernel row qy the K slice of output pixel (y, x) is the 64
// CONTIGUOUS bf16 of pixels x-2 .. x+1 in row y + qy.  A tensor map whose pixel stride (32 B) is smaller than
// its innermost extent (128 B) lets TMA deliver exactly that overlapping window per GEMM row, already in the
// K-major SWIZZLE_128B layout tcgen05 wants: the im2col costs no instructions at all.
//
// GEMM per tile (11 x 11 stem pixels = 5 x 5 pooled outputs, M = 121 of 128 rows):  K = 4 kernel rows x 64,
// N = 128 = [64 RGB | 64 depth] output channels (the RGB rows of the weight matrix are zero on the depth
// channels and vice versa), fp32-grade through three bf16 products  hi*hi + hi*lo + lo*hi  into one TMEM
// accumulator (the stem feeds the gate, whose hard decisions must equal the fp32 reference's).
//
// Persistent warp-specialised CTAs: warp 0 TMA producer (16 KiB slabs, ring of 5), warp 1 MMA issuer (48 UMMAs
// 128x128x16 per tile, all operands in uniform registers: the CTA owns the whole TMEM, base 0), 8 epilogue warps
// (TMEM -> BN + ReLU -> fuse -> shared memory -> max-pool -> NHWC stores, 16 channels at a time, double-buffered
// accumulators so the epilogue of tile i overlaps the MMAs of tile i+1).  The split weights (128 KiB) stay in
// shared memory for the kernel's lifetime.
#include "common.cuh"
#include "tma_host.cuh"

namespace dynmm {
namespace stems2d {

constexpr int kPT = 5;                      // pooled tile edge
constexpr int kST = 2 * kPT + 1;            // stem tile edge (11)
constexpr int kPos = kST * kST;             // 121 GEMM rows of 128
constexpr int kSlab = 128 * 128;            // one [128 rows][64 bf16] operand slab (16 KiB)
constexpr int kRing = 5;                    // A slabs in flight
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kPassCh = 16;                 // channels per epilogue pass
// shared memory layout (after 1024-byte alignment)
constexpr int kOffW = 0;                                   // [4 qy][hi, lo][128 n][128 B]
constexpr int kOffA = kOffW + 8 * kSlab;                   // [kRing][128 rows][128 B]
constexpr int kOffTile = kOffA + kRing * kSlab;            // s_fuse, s_dep: [121][16] fp32 each (swizzled)
constexpr int kTileBytes = kPos * kPassCh * 4;             // 7744
constexpr int kOffBn = kOffTile + 2 * kTileBytes;          // scale_rgb, shift_rgb, scale_d, shift_d (64 each)
constexpr int kOffCtl = kOffBn + 256 * 4;
constexpr int kSmemBytes = 1024 + kOffCtl + 256;
static_assert(kSmemBytes <= 227 * 1024, "stem_s2d shared memory");

struct __align__(8) Ctl {
  uint64_t full[kRing];
  uint64_t empty[kRing];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t w_full;
  uint32_t tmem_base;
};
static_assert(sizeof(Ctl) <= 256, "Ctl");

__device__ __forceinline__ void split1(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ---- pre-pass: NCHW fp32 -> padded space-to-depth bf16 hi / lo planes [b][Hs2+3][Ws2+3][16]
__global__ void s2d_pack_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int b, int H, int W,
                                int Hp2, int Wp2, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const long long total = 1LL * b * Hp2 * Wp2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % Wp2);
    const int py = (int)((i / Wp2) % Hp2);
    const int n = (int)(i / ((long long)Wp2 * Hp2));
    const int y0 = 2 * (py - 2), x0 = 2 * (px - 2);
    __align__(16) __nv_bfloat16 hi[16], lo[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = y0 + dy, x = x0 + dx;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = in ? __ldg(rgb + ((static_cast<size_t>(n) * 3 + c) * H + y) * W + x) : 0.f;
          split1(v, hi[(dy * 2 + dx) * 3 + c], lo[(dy * 2 + dx) * 3 + c]);
        }
        const float d = in ? __ldg(depth + (static_cast<size_t>(n) * H + y) * W + x) : 0.f;
        split1(d, hi[12 + dy * 2 + dx], lo[12 + dy * 2 + dx]);
      }
    }
    uint4* oh = reinterpret_cast<uint4*>(out_hi + i * 16);
    uint4* ol = reinterpret_cast<uint4*>(out_lo + i * 16);
    oh[0] = reinterpret_cast<const uint4*>(hi)[0];
    oh[1] = reinterpret_cast<const uint4*>(hi)[1];
    ol[0] = reinterpret_cast<const uint4*>(lo)[0];
    ol[1] = reinterpret_cast<const uint4*>(lo)[1];
  }
}

// ---- weights: [7][7][cin][64] fp32 (the layout dynmm_stem_fwd takes) -> bf16 [hi, lo][128 n][256 k],
// k = (qy+2) * 64 + (qx+2) * 16 + ch, ky = 2 qy + dy + 3, kx = 2 qx + dx + 3
__global__ void s2d_pack_weights_kernel(const float* __restrict__ w_rgb, const float* __restrict__ w_d,
                                        __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 256) return;
  const int n = i / 256, k = i % 256;
  const int qy = k / 64 - 2, qx = (k % 64) / 16 - 2, ch = k % 16;
  float v = 0.f;
  if (n < 64 && ch < 12) {
    const int dd = ch / 3, c = ch % 3;
    const int ky = 2 * qy + (dd >> 1) + 3, kx = 2 * qx + (dd & 1) + 3;
    if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w_rgb[((ky * 7 + kx) * 3 + c) * 64 + n];
  } else if (n >= 64 && ch >= 12) {
    const int dd = ch - 12;
    const int ky = 2 * qy + (dd >> 1) + 3, kx = 2 * qx + (dd & 1) + 3;
    if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w_d[(ky * 7 + kx) * 64 + (n - 64)];
  }
  __nv_bfloat16 hi, lo;
  split1(v, hi, lo);
  out[i] = hi;
  out[128 * 256 + i] = lo;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// byte offset of (row, ch) in a [121][16] fp32 tile; 16-byte chunks XOR-swizzled so that 8 consecutive rows
// writing the same chunk hit 8 different bank groups
__device__ __forceinline__ uint32_t tile_off(int row, int ch) {
  return row * 64 + ((((ch >> 2) ^ ((row >> 1) & 3)) << 4) | ((ch & 3) << 2));
}

__global__ void __launch_bounds__(kThreads, 1)
stem_s2d_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                const __grid_constant__ CUtensorMap map_w, int Hs, int Ws, const float* __restrict__ scale_rgb,
                const float* __restrict__ shift_rgb, const float* __restrict__ scale_d, const float* __restrict__ shift_d,
                float* __restrict__ rgb_f32, float* __restrict__ depth_f32, __nv_bfloat16* __restrict__ rgb_bf16,
                __nv_bfloat16* __restrict__ depth_bf16, int tiles_x, int tiles_y, int batch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem + kOffW;
  uint8_t* s_a = smem + kOffA;
  uint8_t* s_fuse = smem + kOffTile;
  uint8_t* s_dep = s_fuse + kTileBytes;
  float* s_bn = reinterpret_cast<float*>(smem + kOffBn);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + kOffCtl);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;
  const int total_tiles = tiles_x * tiles_y * batch;

  if (tid == 0) {
    tma_prefetch_desc(&map_hi);
    tma_prefetch_desc(&map_lo);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], kEpiWarps);
    }
    mbar_init(&ctl->w_full, 1);
    fence_mbar_init();
    // split weights: 8 slabs [128 n][64 k] = (qy, hi / lo), once per CTA
    mbar_expect_tx(&ctl->w_full, 8 * kSlab);
    for (int qy = 0; qy < 4; ++qy)
      for (int hl = 0; hl < 2; ++hl) tma_load_3d(s_w + (qy * 2 + hl) * kSlab, &map_w, &ctl->w_full, qy * 64, 0, hl);
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);      // the whole TMEM: base 0, MMA operands stay in uniform registers
    tmem_relinquish();
  }
  if (tid >= 64 && tid < 128) {
    const int c = tid - 64;
    s_bn[c] = scale_rgb[c];
    s_bn[64 + c] = shift_rgb[c];
    s_bn[128 + c] = scale_d[c];
    s_bn[192 + c] = shift_d[c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (ctl->tmem_base != 0) __trap();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / (tiles_x * tiles_y);
        const int sy0 = 2 * (((tile / tiles_x) % tiles_y) * kPT) - 1, sx0 = 2 * ((tile % tiles_x) * kPT) - 1;
        for (int qy = 0; qy < 4; ++qy) {
          for (int hl = 0; hl < 2; ++hl) {
            mbar_wait(&ctl->empty[slot], phase ^ 1);
            mbar_expect_tx(&ctl->full[slot], kPos * 128);
            // GEMM row (ly, lx) <- P[sy0 + ly + qy (+2 border, -2 kernel offset)][sx0 + lx .. + 3][16]
            tma_load_4d(s_a + slot * kSlab, hl ? &map_lo : &map_hi, &ctl->full[slot], 0, sx0, sy0 + qy, n);
            if (++slot == kRing) {
              slot = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
      mbar_wait(&ctl->w_full, 0);
      tc_fence_after();
      int slot = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const uint32_t acc = local & 1;
        mbar_wait(&ctl->acc_empty[acc], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = acc * 128;
        for (int qy = 0; qy < 4; ++qy) {
          const uint32_t wh = w_base + (qy * 2) * kSlab, wl = wh + kSlab;
          // hi slab: A_hi x W_hi and A_hi x W_lo
          mbar_wait(&ctl->full[slot], phase);
          tc_fence_after();
          {
            const uint32_t sa = a_base + slot * kSlab;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(d_tmem, umma_desc_sw128(sa) + 2 * k, umma_desc_sw128(wh) + 2 * k, idesc, (qy | k) != 0);
              umma_bf16(d_tmem, umma_desc_sw128(sa) + 2 * k, umma_desc_sw128(wl) + 2 * k, idesc, 1u);
            }
          }
          umma_commit(&ctl->empty[slot]);
          if (++slot == kRing) {
            slot = 0;
            phase ^= 1;
          }
          // lo slab: A_lo x W_hi
          mbar_wait(&ctl->full[slot], phase);
          tc_fence_after();
          {
            const uint32_t sa = a_base + slot * kSlab;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, umma_desc_sw128(sa) + 2 * k, umma_desc_sw128(wh) + 2 * k, idesc, 1u);
          }
          umma_commit(&ctl->empty[slot]);
          if (qy == 3) umma_commit(&ctl->acc_full[acc]);
          if (++slot == kRing) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (8 warps)
    const int et = tid - 64;                      // 0..255
    const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31
    const int hh = (warp - 2) >> 2;               // which 8 of the 16 channels of a pass
    const int row = quarter * 32 + lane;          // GEMM row = stem position (ly * 11 + lx)
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const uint32_t acc = local & 1;
      const int n = tile / (tiles_x * tiles_y);
      const int py0 = ((tile / tiles_x) % tiles_y) * kPT, px0 = (tile % tiles_x) * kPT;
      const int sy0 = 2 * py0 - 1, sx0 = 2 * px0 - 1;
      mbar_wait(&ctl->acc_full[acc], (local >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = (static_cast<uint32_t>(quarter * 32) << 16) + acc * 128;
      for (int pass = 0; pass < 64 / kPassCh; ++pass) {
        const int c = pass * kPassCh + hh * 8;    // first of this thread's 8 channels
        uint32_t vr[8], vd[8];
        tmem_ld8(t_row + c, vr);
        tmem_ld8(t_row + 64 + c, vd);
        tmem_ld_wait();
        if (pass == 64 / kPassCh - 1) {
          // accumulator fully read: the MMAs of the tile after next may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->acc_empty[acc]);
        }
        if (row < kPos) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int cc = c + q * 4;
            float4 r, d;
            r.x = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 0]), s_bn[cc + 0], s_bn[64 + cc + 0]), 0.f);
            r.y = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 1]), s_bn[cc + 1], s_bn[64 + cc + 1]), 0.f);
            r.z = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 2]), s_bn[cc + 2], s_bn[64 + cc + 2]), 0.f);
            r.w = fmaxf(fmaf(__uint_as_float(vr[q * 4 + 3]), s_bn[cc + 3], s_bn[64 + cc + 3]), 0.f);
            d.x = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 0]), s_bn[128 + cc + 0], s_bn[192 + cc + 0]), 0.f);
            d.y = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 1]), s_bn[128 + cc + 1], s_bn[192 + cc + 1]), 0.f);
            d.z = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 2]), s_bn[128 + cc + 2], s_bn[192 + cc + 2]), 0.f);
            d.w = fmaxf(fmaf(__uint_as_float(vd[q * 4 + 3]), s_bn[128 + cc + 3], s_bn[192 + cc + 3]), 0.f);
            r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;      // rgb + depth (model_skip_mod_globalgate.py:258)
            const uint32_t off = tile_off(row, hh * 8 + q * 4);
            *reinterpret_cast<float4*>(s_fuse + off) = r;
            *reinterpret_cast<float4*>(s_dep + off) = d;
          }
        }
        named_barrier(1, 32 * kEpiWarps);
        // 3x3 / stride 2 / pad 1 max-pool of both tiles, 16 channels of 25 pooled pixels
        for (int i = et; i < kPT * kPT * kPassCh; i += 32 * kEpiWarps) {
          const int pp = i >> 4, ch = i & 15;
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
              const uint32_t off = tile_off((2 * ly + dy) * kST + 2 * lx + dx, ch);
              mf = fmaxf(mf, *reinterpret_cast<const float*>(s_fuse + off));
              md = fmaxf(md, *reinterpret_cast<const float*>(s_dep + off));
            }
          }
          const size_t o = ((static_cast<size_t>(n) * Hp + py) * Wp + px) * 64 + pass * kPassCh + ch;
          if (rgb_f32) rgb_f32[o] = mf;
          if (depth_f32) depth_f32[o] = md;
          if (rgb_bf16) rgb_bf16[o] = __float2bfloat16_rn(mf);
          if (depth_bf16) depth_bf16[o] = __float2bfloat16_rn(md);
        }
        named_barrier(1, 32 * kEpiWarps);       // the tiles are consumed before the next pass overwrites them
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(0, 512);
  }
}

}  // namespace stems2d
}  // namespace dynmm

using namespace dynmm;

static void s2d_geometry(int h, int w, int* Hs, int* Ws, int* Hp2, int* Wp2) {
  *Hs = (h + 6 - 7) / 2 + 1;
  *Ws = (w + 6 - 7) / 2 + 1;
  *Hp2 = *Hs + 3;      // s2d rows -2 .. Hs
  *Wp2 = *Ws + 3;
}

extern "C" long long dynmm_stem_s2d_workspace(int b, int h, int w) {
  if (b < 1 || h < 7 || w < 7) return -1;
  int Hs, Ws, Hp2, Wp2;
  s2d_geometry(h, w, &Hs, &Ws, &Hp2, &Wp2);
  return 2LL * b * Hp2 * Wp2 * 16 * 2 + 256;
}

extern "C" int dynmm_stem_s2d_pack_weights(const float* w_rgb, const float* w_d, void* packed, void* stream) {
  DYNMM_CHECK_ARG(w_rgb && w_d && packed, "stem_s2d_pack_weights: null pointer");
  stems2d::s2d_pack_weights_kernel<<<128, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_rgb, w_d, static_cast<__nv_bfloat16*>(packed));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_stem_s2d_fwd(const float* rgb, const float* depth, int b, int h, int w, const void* w_packed,
                                  const float* scale_rgb, const float* shift_rgb, const float* scale_d,
                                  const float* shift_d, void* workspace, long long workspace_bytes, float* rgb_f32,
                                  float* depth_f32, void* rgb_bf16, void* depth_bf16, void* stream_) {
  using namespace stems2d;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(rgb && depth && w_packed && scale_rgb && shift_rgb && scale_d && shift_d && workspace,
                  "stem_s2d: null pointer");
  DYNMM_CHECK_ARG(b >= 1 && h >= 7 && w >= 7, "stem_s2d: bad shape");
  DYNMM_CHECK_ARG(workspace_bytes >= dynmm_stem_s2d_workspace(b, h, w), "stem_s2d: workspace too small");
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, "stem_s2d: packed weights must be 16-byte aligned");
  int Hs, Ws, Hp2, Wp2;
  s2d_geometry(h, w, &Hs, &Ws, &Hp2, &Wp2);
  const int Hp = (Hs + 2 - 3) / 2 + 1, Wp = (Ws + 2 - 3) / 2 + 1;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 127) & ~uintptr_t(127));
  const size_t plane = static_cast<size_t>(b) * Hp2 * Wp2 * 16 * 2;
  __nv_bfloat16* p_hi = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* p_lo = reinterpret_cast<__nv_bfloat16*>(ws + plane);

  // tensor maps: A = overlapping 4-pixel windows of the padded s2d image; coordinate x is the window start in
  // padded pixels (= stem x, because the border is 2 and the kernel offset -2), coordinate y the padded row
  CUtensorMap map_hi, map_lo, map_w;
  {
    const uint64_t dims[4] = {64, (uint64_t)Ws, (uint64_t)Hp2, (uint64_t)b};
    const uint64_t strides[3] = {32, (uint64_t)Wp2 * 32, (uint64_t)Hp2 * Wp2 * 32};
    const uint32_t box[4] = {64, (uint32_t)kST, (uint32_t)kST, 1};
    int rc = encode_map(&map_hi, p_hi, 4, dims, strides, box);
    if (rc) return rc;
    rc = encode_map(&map_lo, p_lo, 4, dims, strides, box);
    if (rc) return rc;
    const uint64_t wdims[3] = {256, 128, 2};
    const uint64_t wstrides[2] = {256 * 2, 128 * 256 * 2};
    const uint32_t wbox[3] = {64, 128, 1};
    rc = encode_map(&map_w, w_packed, 3, wdims, wstrides, wbox);
    if (rc) return rc;
  }
  const long long pix = 1LL * b * Hp2 * Wp2;
  const int pack_grid = (int)((pix + 255) / 256 < 8LL * num_sms() ? (pix + 255) / 256 : 8LL * num_sms());
  s2d_pack_kernel<<<pack_grid, 256, 0, stream>>>(rgb, depth, b, h, w, Hp2, Wp2, p_hi, p_lo);
  DYNMM_LAUNCH_CHECK();

  static cudaError_t attr_err =
      cudaFuncSetAttribute(stem_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  DYNMM_CUDA(attr_err);
  const int tiles_x = ceil_div(Wp, kPT), tiles_y = ceil_div(Hp, kPT);
  const long long total = 1LL * tiles_x * tiles_y * b;
  DYNMM_CHECK_ARG(total < (1LL << 30), "stem_s2d: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  stem_s2d_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_hi, map_lo, map_w, Hs, Ws, scale_rgb, shift_rgb, scale_d,
                                                         shift_d, rgb_f32, depth_f32,
                                                         static_cast<__nv_bfloat16*>(rgb_bf16),
                                                         static_cast<__nv_bfloat16*>(depth_bf16), tiles_x, tiles_y, b);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
