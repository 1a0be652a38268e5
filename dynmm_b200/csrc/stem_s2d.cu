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
// One box load per plane brings the 10 s2d rows a tile touches; kernel row qy is the view 16 windows (2 KiB,
// swizzle phase preserved) further into the same tile, so each s2d pixel is fetched once per plane, not 4 times.
//
// GEMM per tile (16 x 7 stem pixels, 15 used = 7 x 3 pooled outputs, M = 112 of 128 rows):  K = 4 kernel rows x 64,
// N = 128 = [64 RGB | 64 depth] output channels (the RGB rows of the weight matrix are zero on the depth
// channels and vice versa), fp32-grade through three bf16 products  hi*hi + hi*lo + lo*hi  into one TMEM
// accumulator (the stem feeds the gate, whose hard decisions must equal the fp32 reference's).
//
// Persistent warp-specialised CTAs: warp 0 TMA producer (2 x 20 KiB per tile, 2 tiles in flight), warp 1 MMA issuer (48 UMMAs
// 128x128x16 per tile, the CTA owns the whole TMEM, base 0), two groups of 8 epilogue warps, one per accumulator buffer
// (TMEM -> BN + ReLU -> fuse -> horizontal 3-max by warp shuffles -> 7x7 row maxima in shared memory -> vertical 3-max ->
// NHWC stores, 16 channels at a time), so the epilogues of tiles i and i+1 and the MMAs of tile i+2 overlap.  The split weights (128 KiB) stay in
// shared memory for the kernel's lifetime.
#include <cstring>
#include "common.cuh"
#include "tma_host.cuh"

namespace dynmm {
namespace stems2d {

#ifndef DYNMM_STEM_TRACE
#define DYNMM_STEM_TRACE 0   // experiments only: clock stamps of CTA 0 into the depth bf16 output (tools/stem_trace.py)
#endif
constexpr int kPW = 7, kPH = 3;             // pooled tile: 7 wide, 3 tall
constexpr int kSW = 16;                     // stem columns fetched per tile (2*kPW + 1 = 15 used, +1 keeps views aligned)
constexpr int kSH = 2 * kPH + 1;            // stem rows per tile (7)
constexpr int kRows = kSW * kSH;            // 112 GEMM rows of 128, row = ly * 16 + lx
constexpr int kFetchRows = kSH + 3;         // s2d rows a tile needs: kernel rows qy = 0..3 are views 16 rows apart
constexpr int kPlane = kSW * kFetchRows * 128;   // 20 KiB: one plane (hi or lo) of a tile, [10][16] windows x 128 B
constexpr int kStage = 2 * kPlane;          // hi + lo
constexpr int kWSlab = 128 * 128;           // one [128 n][64 k] weight slab (16 KiB)
constexpr int kRing = 2;                    // tiles in flight
constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kPassCh = 16;                 // channels per epilogue pass
constexpr int kGroupWarps = kEpiWarps / 2;  // two epilogue groups, one per accumulator buffer, work on alternate tiles
// shared memory layout (after 1024-byte alignment)
constexpr int kOffW = 0;                                   // [4 qy][hi, lo][128 n][128 B]
constexpr int kOffA = kOffW + 8 * kWSlab;                  // [kRing][hi, lo][160 rows][128 B]
constexpr int kOffTile = kOffA + kRing * kStage;           // per group: row-maxima of the fused / depth maps,
constexpr int kTileBytes = (3 * 52 + 2 + kSH * kPW) * 16;     // channel-quad major, see h_off: 207 float4 = 3312 B each
constexpr int kOffBn = kOffTile + 4 * kTileBytes;          // scale_rgb, shift_rgb, scale_d, shift_d (64 each)
constexpr int kOffCtl = kOffBn + 256 * 4;
constexpr int kSmemBytes = 1024 + kOffCtl + 256;
static_assert(kSmemBytes <= 227 * 1024, "stem_s2d shared memory");
static_assert(kPlane % 1024 == 0 && (kSW * 128) % 1024 == 0, "tap views must keep the 128B-swizzle phase");

// BN scale / shift of both stems as a kernel parameter (constant bank): [scale_rgb | shift_rgb | scale_d | shift_d]
struct StemBn {
  float v[256];
};

struct __align__(8) Ctl {
  uint64_t full[kRing];
  uint64_t empty[kRing];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
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
  // the stem kernel is launched with programmatic stream serialization: its prologue (TMEM allocation, 128 KiB of
  // weights per CTA) runs under this kernel, its first load of the planes waits for this grid to complete
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
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

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
// byte offset of (stem row ly, pooled column plx, channel ch) in a tile of horizontal maxima.  Channel-quad major
// [4 quads][7 x 7 positions] float4 with the quads at 16-byte units 0, 52, 106, 158 (= 0, 4, 2, 6 mod 8): the writers
// of a warp (even lanes quad 2 h, odd lanes quad 2 h + 1, consecutive positions) and a quarter-warp of pooling
// threads (2 positions x 4 quads) both touch every bank once -- the pixel-major [7][7][16] tile cost 7 wavefronts per
// store
__device__ __forceinline__ uint32_t h_off(int ly, int plx, int ch) {
  const int q = ch >> 2;
  return ((q * 52 + (q >> 1) * 2 + ly * kPW + plx) * 4 + (ch & 3)) * 4;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// kConstBn: the BN vectors come from the parameter `bn` (host copy, constant-bank loads) instead of shared memory --
// eight LDS.128 per pass less in an epilogue that is bound by the shared-memory / shuffle instruction queue
template <bool kConstBn>
__global__ void __launch_bounds__(kThreads, 1)
stem_s2d_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ StemBn bn, int Hs, int Ws,
                const float* __restrict__ scale_rgb,
                const float* __restrict__ shift_rgb, const float* __restrict__ scale_d, const float* __restrict__ shift_d,
                float* __restrict__ rgb_f32, float* __restrict__ depth_f32, __nv_bfloat16* __restrict__ rgb_bf16,
                __nv_bfloat16* __restrict__ depth_bf16, int split, int tiles_x, int tiles_y, int batch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem + kOffW;
  uint8_t* s_a = smem + kOffA;
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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], kGroupWarps);
    }
    mbar_init(&ctl->w_full, 1);
    fence_mbar_init();
    // split weights: 8 slabs [128 n][64 k] = (qy, hi / lo), once per CTA
    mbar_expect_tx(&ctl->w_full, 8 * kWSlab);
    for (int qy = 0; qy < 4; ++qy)
      for (int hl = 0; hl < 2; ++hl) tma_load_3d(s_w + (qy * 2 + hl) * kWSlab, &map_w, &ctl->w_full, qy * 64, 0, hl);
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);      // the whole TMEM: base 0, MMA operands stay in uniform registers
    tmem_relinquish();
  }
  if (!kConstBn && tid >= 64 && tid < 128) {
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
    // ------------------------------------------------------------ TMA producer: two box loads per tile
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;\n" ::: "memory");     // the planes are written by the pre-pass
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / (tiles_x * tiles_y);
        const int sy0 = 2 * (((tile / tiles_x) % tiles_y) * kPH) - 1, sx0 = 2 * ((tile % tiles_x) * kPW) - 1;
        mbar_wait(&ctl->empty[slot], phase ^ 1);
        mbar_expect_tx(&ctl->full[slot], kStage);
        // window (wy, lx) of the box <- P[sy0 + wy][sx0 + lx .. + 3][16]  (border 2, kernel offset -2 cancel)
        tma_load_4d(s_a + slot * kStage, &map_hi, &ctl->full[slot], 0, sx0, sy0, n);
        tma_load_4d(s_a + slot * kStage + kPlane, &map_lo, &ctl->full[slot], 0, sx0, sy0, n);
        if (++slot == kRing) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: the whole warp runs the loop
    // converged, only the tcgen05 instructions sit under elect.sync (one lane looping inside `if (lane == 0)` cannot
    // issue more than one UMMA per ~90 cycles, tools/umma_issue_bench.cu)
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
      mbar_wait(&ctl->w_full, 0);
      tc_fence_after();
      int slot = 0;
      uint32_t phase = 0;
      int local = 0;
#if DYNMM_STEM_TRACE
      long long* trm = reinterpret_cast<long long*>(depth_bf16) + (static_cast<size_t>(blockIdx.x) * 20 + 16) * 1024;
      int tm = 0;
#define MSTAMP() do { if (blockIdx.x < 2 && lane == 0 && tm < 1024) trm[tm++] = clock64(); } while (0)
#else
#define MSTAMP() do { } while (0)
#endif
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const uint32_t acc = local & 3;       // four 128-column accumulators: the MMAs run up to 3 tiles ahead
        MSTAMP();
        mbar_wait(&ctl->acc_empty[acc], ((local >> 2) & 1) ^ 1);
        MSTAMP();
        mbar_wait(&ctl->full[slot], phase);
        MSTAMP();
        tc_fence_after();
        const uint32_t d_tmem = acc * 128;
        const uint32_t stage = a_base + slot * kStage;
        if (elect_one()) {
#pragma unroll
          for (int qy = 0; qy < 4; ++qy) {
            // kernel row qy: GEMM row (ly, lx) reads window (ly + qy, lx) = the same planes, 16 rows (2 KiB) further
            const uint32_t ah = stage + qy * (kSW * 128), al = ah + kPlane;
            const uint32_t wh = w_base + (qy * 2) * kWSlab, wl = wh + kWSlab;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(d_tmem, umma_desc_sw128(ah) + 2 * k, umma_desc_sw128(wh) + 2 * k, idesc, (qy | k) != 0);
              umma_bf16(d_tmem, umma_desc_sw128(ah) + 2 * k, umma_desc_sw128(wl) + 2 * k, idesc, 1u);
              umma_bf16(d_tmem, umma_desc_sw128(al) + 2 * k, umma_desc_sw128(wh) + 2 * k, idesc, 1u);
            }
          }
          umma_commit(&ctl->empty[slot]);
          umma_commit(&ctl->acc_full[acc]);
        }
        __syncwarp();
        if (++slot == kRing) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: 2 groups x 8 warps.  Group g owns the
    // tiles with local index = g (mod 2) (accumulators g and g + 2), with its own staging tiles and named barrier.
    // The 3x3 / stride-2 max-pool is separable: a warp holds two stem rows of 16 columns (TMEM lane = ly * 16 + lx),
    // so the horizontal 3-max is two shuffles; only the 7 x 7 row-maxima per channel go through shared memory and the
    // pooling threads read 3 values per output instead of 9.
    const int ewarp = warp - 2;
    const int g = ewarp / kGroupWarps;
    const int gt = tid - 64 - g * 32 * kGroupWarps;   // 0..255 inside the group
    const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31
    const int half = (ewarp % kGroupWarps) >> 2;  // which 8 of the 16 channels of a pass
    const int row = quarter * 32 + lane;          // GEMM row = stem position ly * 16 + lx
    const int ly = row >> 4, lx = row & 15;
    const bool odd = lane & 1;
    const bool writer = ly < kSH && lx <= 2 * (kPW - 1) + 1;     // lane pair (2 p, 2 p + 1) owns pooled column p
    uint8_t* s_hf = smem + kOffTile + g * 2 * kTileBytes;     // horizontal maxima of rgb + depth
    uint8_t* s_hd = s_hf + kTileBytes;                        // ... of depth
    // pooling: the first 84 threads of the group own (pooled pixel pp = gt >> 2 < 21) x (channel quad c4 of the pass)
    const int pp = gt >> 2, c4 = (gt & 3) * 4;
    const int ply = pp / kPW, plx = pp - ply * kPW;
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
#if DYNMM_STEM_TRACE
    long long* trc = reinterpret_cast<long long*>(depth_bf16) + (static_cast<size_t>(blockIdx.x) * 20 + ewarp) * 1024;
    int tn = 0;
#define STAMP() do { if (blockIdx.x < 2 && lane == 0 && tn < 1024) trc[tn++] = clock64(); } while (0)
#else
#define STAMP() do { } while (0)
#endif
    for (int local = g; blockIdx.x + static_cast<long long>(local) * gridDim.x < total_tiles; local += 2) {
      const int tile = blockIdx.x + local * gridDim.x;
      const int n = tile / (tiles_x * tiles_y);
      const int py0 = ((tile / tiles_x) % tiles_y) * kPH, px0 = (tile % tiles_x) * kPW;
      const int sy0 = 2 * py0 - 1, sx0 = 2 * px0 - 1;
      const int py = py0 + ply, px = px0 + plx;
      const bool item = pp < kPW * kPH && py < Hp && px < Wp;
      // stem positions outside the map (or the unused 8th row / 16th column) never win a maximum
      const int gy = sy0 + ly, gx = sx0 + lx;
      const bool inside = ly < kSH && gy >= 0 && gy < Hs && gx >= 0 && gx < Ws;
      const uint32_t acc = local & 3;             // group g reads buffers g and g + 2
      STAMP();
      mbar_wait(&ctl->acc_full[acc], (local >> 2) & 1);
      tc_fence_after();
      STAMP();
      const uint32_t t_row = t_lane + acc * 128;
#pragma unroll 1
      for (int pass = 0; pass < 64 / kPassCh; ++pass) {
        const int c = pass * kPassCh + half * 8;  // first of this thread's 8 channels
        uint32_t vr[8], vd[8];
        tmem_ld8(t_row + c, vr);
        tmem_ld8(t_row + 64 + c, vd);
        tmem_ld_wait();
        STAMP();
        if (pass == 64 / kPassCh - 1) {
          // accumulator fully read: the MMAs of the tile after next may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->acc_empty[acc]);
        }
        float f[8], d[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float4 sr, br, sd, bd;
          if (kConstBn) {
            const int k = c + 4 * q;
            sr = make_float4(bn.v[k], bn.v[k + 1], bn.v[k + 2], bn.v[k + 3]);
            br = make_float4(bn.v[64 + k], bn.v[64 + k + 1], bn.v[64 + k + 2], bn.v[64 + k + 3]);
            sd = make_float4(bn.v[128 + k], bn.v[128 + k + 1], bn.v[128 + k + 2], bn.v[128 + k + 3]);
            bd = make_float4(bn.v[192 + k], bn.v[192 + k + 1], bn.v[192 + k + 2], bn.v[192 + k + 3]);
          } else {
            sr = *reinterpret_cast<const float4*>(s_bn + c + 4 * q), br = *reinterpret_cast<const float4*>(s_bn + 64 + c + 4 * q);
            sd = *reinterpret_cast<const float4*>(s_bn + 128 + c + 4 * q), bd = *reinterpret_cast<const float4*>(s_bn + 192 + c + 4 * q);
          }
          d[4 * q + 0] = fmaxf(fmaf(__uint_as_float(vd[4 * q + 0]), sd.x, bd.x), 0.f);
          d[4 * q + 1] = fmaxf(fmaf(__uint_as_float(vd[4 * q + 1]), sd.y, bd.y), 0.f);
          d[4 * q + 2] = fmaxf(fmaf(__uint_as_float(vd[4 * q + 2]), sd.z, bd.z), 0.f);
          d[4 * q + 3] = fmaxf(fmaf(__uint_as_float(vd[4 * q + 3]), sd.w, bd.w), 0.f);
          // rgb + depth (model_skip_mod_globalgate.py:258)
          f[4 * q + 0] = fmaxf(fmaf(__uint_as_float(vr[4 * q + 0]), sr.x, br.x), 0.f) + d[4 * q + 0];
          f[4 * q + 1] = fmaxf(fmaf(__uint_as_float(vr[4 * q + 1]), sr.y, br.y), 0.f) + d[4 * q + 1];
          f[4 * q + 2] = fmaxf(fmaf(__uint_as_float(vr[4 * q + 2]), sr.z, br.z), 0.f) + d[4 * q + 2];
          f[4 * q + 3] = fmaxf(fmaf(__uint_as_float(vr[4 * q + 3]), sr.w, br.w), 0.f) + d[4 * q + 3];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (!inside) {
            f[e] = -INFINITY;
            d[e] = -INFINITY;
          }
        }
        // horizontal 3-max over columns 2 p, 2 p + 1, 2 p + 2 for pooled column p, split over the lane pair (2 p, 2 p + 1):
        // the even lane finishes channels 0..3, the odd lane channels 4..7.  Exchange A (partner lane): the even lane
        // hands over its channels 4..7 and receives the odd lane's 0..3; exchange B (two lanes up, inside the 16-lane
        // stem row): the even lane receives column 2 p + 2's channels 0..3 from their owner, the odd lane the channels
        // 4..7 of that column from lane 2 p + 3, which got them in A.  8 shuffles per map instead of 16 -- the epilogue
        // is bound by the shuffle / shared-memory instruction queue (tools/stem_trace.py: 1300 of a pass's 2700 cycles).
        float hf[4], hd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float own_f = odd ? f[e + 4] : f[e], own_d = odd ? d[e + 4] : d[e];
          const float a_f = __shfl_xor_sync(0xffffffffu, odd ? f[e] : f[e + 4], 1);
          const float a_d = __shfl_xor_sync(0xffffffffu, odd ? d[e] : d[e + 4], 1);
          const float b_f = __shfl_down_sync(0xffffffffu, odd ? a_f : own_f, 2, 16);
          const float b_d = __shfl_down_sync(0xffffffffu, odd ? a_d : own_d, 2, 16);
          hf[e] = fmaxf(own_f, fmaxf(a_f, b_f));
          hd[e] = fmaxf(own_d, fmaxf(a_d, b_d));
        }
        STAMP();
        if (writer) {
          const uint32_t off = h_off(ly, lx >> 1, half * 8 + (odd ? 4 : 0));
          *reinterpret_cast<float4*>(s_hf + off) = make_float4(hf[0], hf[1], hf[2], hf[3]);
          *reinterpret_cast<float4*>(s_hd + off) = make_float4(hd[0], hd[1], hd[2], hd[3]);
        }
        STAMP();
        named_barrier(1 + g, 32 * kGroupWarps);
        STAMP();
        // vertical 3-max over stem rows 2*ply .. 2*ply+2: 21 pooled pixels x 4 channel quads
        if (item) {
          float4 mf = *reinterpret_cast<const float4*>(s_hf + h_off(2 * ply, plx, c4));
          float4 md = *reinterpret_cast<const float4*>(s_hd + h_off(2 * ply, plx, c4));
#pragma unroll
          for (int dy = 1; dy < 3; ++dy) {
            const float4 a = *reinterpret_cast<const float4*>(s_hf + h_off(2 * ply + dy, plx, c4));
            const float4 b = *reinterpret_cast<const float4*>(s_hd + h_off(2 * ply + dy, plx, c4));
            mf.x = fmaxf(mf.x, a.x); mf.y = fmaxf(mf.y, a.y); mf.z = fmaxf(mf.z, a.z); mf.w = fmaxf(mf.w, a.w);
            md.x = fmaxf(md.x, b.x); md.y = fmaxf(md.y, b.y); md.z = fmaxf(md.z, b.z); md.w = fmaxf(md.w, b.w);
          }
          const size_t o = ((static_cast<size_t>(n) * Hp + py) * Wp + px) * 64 + pass * kPassCh + c4;
          if (rgb_f32) *reinterpret_cast<float4*>(rgb_f32 + o) = mf;
          if (depth_f32) *reinterpret_cast<float4*>(depth_f32 + o) = md;
          // bf16 maps; split: [hi | lo] halves of the fp32 value (128 channels per pixel), the activation format of the
          // fp32-grade engine mode -- what dynmm_split_from_f32 would make of the fp32 maps
          const size_t o16 = split ? o + (o & ~static_cast<size_t>(63)) : o;      // pixel * 128 + channel
          if (rgb_bf16) {
            uint2 v;
            v.x = pack_bf16(mf.x, mf.y);
            v.y = pack_bf16(mf.z, mf.w);
            *reinterpret_cast<uint2*>(rgb_bf16 + o16) = v;
            if (split) {
              uint2 l;
              l.x = pack_bf16(mf.x - bf16_lo(v.x), mf.y - bf16_hi(v.x));
              l.y = pack_bf16(mf.z - bf16_lo(v.y), mf.w - bf16_hi(v.y));
              *reinterpret_cast<uint2*>(rgb_bf16 + o16 + 64) = l;
            }
          }
          if (depth_bf16 && !DYNMM_STEM_TRACE) {
            uint2 v;
            v.x = pack_bf16(md.x, md.y);
            v.y = pack_bf16(md.z, md.w);
            *reinterpret_cast<uint2*>(depth_bf16 + o16) = v;
            if (split) {
              uint2 l;
              l.x = pack_bf16(md.x - bf16_lo(v.x), md.y - bf16_hi(v.x));
              l.y = pack_bf16(md.z - bf16_lo(v.y), md.w - bf16_hi(v.y));
              *reinterpret_cast<uint2*>(depth_bf16 + o16 + 64) = l;
            }
          }
        }
        STAMP();
        named_barrier(1 + g, 32 * kGroupWarps);   // the row maxima are consumed before the next pass overwrites them
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
                                  float* depth_f32, void* rgb_bf16, void* depth_bf16, int split,
                                  const float* bn_host, void* stream_) {
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
    const uint32_t box[4] = {64, (uint32_t)kSW, (uint32_t)kFetchRows, 1};
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

  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(stem_s2d_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    return e != cudaSuccess ? e
                            : cudaFuncSetAttribute(stem_s2d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }));
  const int tiles_x = ceil_div(Wp, kPW), tiles_y = ceil_div(Hp, kPH);
  const long long total = 1LL * tiles_x * tiles_y * b;
  DYNMM_CHECK_ARG(total < (1LL << 30), "stem_s2d: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  StemBn bn{};
  if (bn_host) memcpy(bn.v, bn_host, sizeof(bn.v));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DYNMM_CUDA(cudaLaunchKernelEx(&cfg, bn_host ? stem_s2d_kernel<true> : stem_s2d_kernel<false>, map_hi, map_lo, map_w, bn,
                                Hs, Ws, scale_rgb, shift_rgb, scale_d, shift_d, rgb_f32, depth_f32,
                                static_cast<__nv_bfloat16*>(rgb_bf16), static_cast<__nv_bfloat16*>(depth_bf16), split ? 1 : 0,
                                tiles_x, tiles_y, b));
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
