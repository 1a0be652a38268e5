// Fused 3x1 -> ReLU -> 1x3 (+BN shift, +residual, ReLU) pair of a NonBottleneck1D block (resnet.py:124-147) for
// the 64-channel stage: ONE kernel, the intermediate never leaves shared memory.
//
// At batch 8 a C = 64 layer is bound by activation traffic (19.7 MB in + 19.7 MB out per conv, 3.4 TB/s reached)
// and by its fixed launch cost.  Here a tile of 8 (H) x 14 (W) output pixels is produced from a 10 x 16 input tile:
//
//   TMA      X tile   [hh 0..9][ww 0..15] x 64 ch     (H halo for the 3x1 taps, W halo for the later 1x3 taps)
//   MMA 1    Y[hh 0..7][ww 0..15] = sum_t  X[hh + t] * W1[t]         three views 16 rows (2 KiB) apart, 12 UMMAs
//   epi 1    TMEM -> +bias, ReLU, zero outside the image (the 1x3 conv pads Y with zeros), bf16 -> shared memory
//            TRANSPOSED to the row order (ww, hh): that is the K-major SWIZZLE_128B A operand of the 1x3 conv,
//            whose taps are views 8 rows (1 KiB) apart
//   MMA 2    Z[wo 0..13][h 0..7] = sum_t  Y[(wo + t), h] * W2[t]     12 UMMAs (rows of wo = 14, 15 are discarded)
//   epi 2    TMEM -> +shift (+ residual via TMA) -> ReLU -> bf16 -> staging -> TMA store of the 8 x 14 box
//
// Same arithmetic, same accumulation order and the same bf16 rounding of the intermediate as the two
// dynmm_conv_igemm_fwd launches it replaces: results are bit-identical (tests/test_gpu_pair.py).
// Persistent CTA per SM: warp 0 TMA producer, warp 1 MMA issuer, 8 warps for epilogue 1 and 8 warps for epilogue 2
// (the epilogues, not the MMAs, bound the kernel); X ring of 3, double-buffered Y / accumulators / staging; both
// weight sets (2 x 24 KiB) resident.  MMA 1 of tile i+1 is issued before MMA 2 of tile i.
#include <stdlib.h>

#include "common.cuh"
#include "tma_host.cuh"

namespace dynmm {
namespace pairk {

constexpr int kC = 64;
constexpr int kTH = 8, kTW = 14;                 // output tile
constexpr int kXW = kTW + 2, kXH = kTH + 2;      // input tile 16 x 10
constexpr int kXBytes = kXW * kXH * 128;         // 20 KiB
constexpr int kYBytes = (kXW + 2) * kTH * 128;   // 18 KiB: 16 written columns + 2 the last tap's view runs into
constexpr int kOutRows = kTW * kTH;              // 112
constexpr int kOutBytes = kOutRows * 128;        // 14 KiB
constexpr int kWBytes = 3 * 64 * 128;            // one conv's weights, [3 taps][64 n][64 k]
constexpr int kXStages = 3;
constexpr int kEpiWarps = 8;                     // per epilogue group; group A runs epilogue 1, group B epilogue 2
constexpr int kThreads = 64 + 2 * 32 * kEpiWarps;
// shared memory (after 1024-byte alignment); every operand region is 1024-byte aligned
constexpr int kOffW1 = 0, kOffW2 = kWBytes;
constexpr int kOffX = 2 * kWBytes;
constexpr int kOffY = kOffX + kXStages * kXBytes;
constexpr int kOffStage = kOffY + 2 * kYBytes;
constexpr int kOffRes = kOffStage + 2 * ((kOutBytes + 1023) / 1024 * 1024);
constexpr int kOutPitch = (kOutBytes + 1023) / 1024 * 1024;
constexpr int kOffShift = kOffRes + 2 * kOutPitch;
constexpr int kOffCtl = kOffShift + 2 * kC * 4;
constexpr int kSmemBytes = 1024 + kOffCtl + 512;
static_assert(kSmemBytes <= 227 * 1024, "conv_pair shared memory");
static_assert(kXBytes % 1024 == 0 && kYBytes % 1024 == 0 && kWBytes % 1024 == 0, "operand alignment");

struct __align__(8) Ctl {
  uint64_t x_full[kXStages], x_empty[kXStages];
  uint64_t acc1_full[2], acc1_empty[2], y_full[2], y_empty[2], acc2_full[2], acc2_empty[2];
  uint64_t res_full[2], res_empty[2];
  uint64_t w_full;
  uint32_t tmem_base;
};
static_assert(sizeof(Ctl) <= 512, "Ctl");

struct Args {
  int n, h, w, tiles_h, tiles_w;
  int relu2, has_res;
  const float* shift1;
  const float* shift2;
  const int32_t* count;
  const int32_t* in_map;
  const int32_t* res_map;
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <bool kRot>      // kRot: conflict-avoiding chunk order in epilogue 1 (default; DYNMM_PAIR_ROT=0 turns it off)
__global__ void __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ Args args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + kOffCtl);
  float* s_shift = reinterpret_cast<float*>(smem + kOffShift);      // [shift1 64 | shift2 64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_out);
    for (int s = 0; s < kXStages; ++s) {
      mbar_init(&ctl->x_full[s], 1);
      mbar_init(&ctl->x_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc1_full[i], 1);
      mbar_init(&ctl->acc1_empty[i], kEpiWarps);
      mbar_init(&ctl->y_full[i], kEpiWarps);
      mbar_init(&ctl->y_empty[i], 1);
      mbar_init(&ctl->acc2_full[i], 1);
      mbar_init(&ctl->acc2_empty[i], kEpiWarps);
      mbar_init(&ctl->res_full[i], 1);
      mbar_init(&ctl->res_empty[i], kEpiWarps);
    }
    mbar_init(&ctl->w_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    if (elect_one()) {        // weights are constants: fetched before the dependency wait
      mbar_expect_tx(&ctl->w_full, 2 * kWBytes);
      tma_load_3d(smem + kOffW1, &map_w1, &ctl->w_full, 0, 0, 0);
      tma_load_3d(smem + kOffW2, &map_w2, &ctl->w_full, 0, 0, 0);
    }
    __syncwarp();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int c = threadIdx.x - 64; c < 2 * kC; c += 2 * 32 * kEpiWarps) {
      const float* src = c < kC ? args.shift1 : args.shift2;
      s_shift[c] = src ? src[c & (kC - 1)] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (ctl->tmem_base != 0) __trap();
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");

  const int active = args.count ? min(*args.count, args.n) : args.n;
  const int tiles_per_sample = args.tiles_h * args.tiles_w;
  const int total_tiles = active * tiles_per_sample;
  const int n_my = (int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // TMEM columns: acc1[b] = 64 b, acc2[b] = 128 + 64 b
  auto tile_coord = [&](int k, int& n, int& h0, int& w0) {
    const int tile = blockIdx.x + k * gridDim.x;
    n = tile / tiles_per_sample;
    const int r = tile - n * tiles_per_sample;
    h0 = (r / args.tiles_w) * kTH;
    w0 = (r % args.tiles_w) * kTW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int xs = 0;
    uint32_t xph = 0;
    for (int k = 0; k < n_my; ++k) {
      int n, h0, w0;
      tile_coord(k, n, h0, w0);
      const int n_in = args.in_map ? args.in_map[n] : n;
      mbar_wait(&ctl->x_empty[xs], xph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&ctl->x_full[xs], kXBytes);
        tma_load_4d(smem + kOffX + xs * kXBytes, &map_x, &ctl->x_full[xs], 0, w0 - 1, h0 - 1, n_in);
      }
      __syncwarp();
      if (++xs == kXStages) {
        xs = 0;
        xph ^= 1;
      }
      if (args.has_res) {
        const int rb = k & 1;
        const int n_res = args.res_map ? args.res_map[n] : n;
        mbar_wait(&ctl->res_empty[rb], ((k >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&ctl->res_full[rb], kOutBytes);
          tma_load_4d(smem + kOffRes + rb * kOutPitch, &map_res, &ctl->res_full[rb], 0, h0, w0, n_res);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: MMA1(0); then MMA1(k+1), MMA2(k)
    constexpr uint32_t idesc = umma_idesc_bf16(128, kC);
    const uint32_t s_base = smem_u32(smem);
    mbar_wait(&ctl->w_full, 0);
    tc_fence_after();
    int xs = 0;
    uint32_t xph = 0;
    auto mma1 = [&](int k) {
      const uint32_t b = k & 1, ph = (k >> 1) & 1;
      mbar_wait(&ctl->acc1_empty[b], ph ^ 1);
      mbar_wait(&ctl->x_full[xs], xph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t x = s_base + kOffX + xs * kXBytes;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint64_t da = umma_desc_sw128(x + t * (kXW * 128));             // rows (hh + t, ww)
          const uint64_t db = umma_desc_sw128(s_base + kOffW1 + t * (64 * 128));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(b * 64, da + 2 * ks, db + 2 * ks, idesc, (t | ks) != 0);
        }
        umma_commit(&ctl->x_empty[xs]);
        umma_commit(&ctl->acc1_full[b]);
      }
      __syncwarp();
      if (++xs == kXStages) {
        xs = 0;
        xph ^= 1;
      }
    };
    auto mma2 = [&](int k) {
      const uint32_t b = k & 1, ph = (k >> 1) & 1;
      mbar_wait(&ctl->acc2_empty[b], ph ^ 1);
      mbar_wait(&ctl->y_full[b], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t y = s_base + kOffY + b * kYBytes;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint64_t da = umma_desc_sw128(y + t * (kTH * 128));              // rows (ww + t, hh)
          const uint64_t db = umma_desc_sw128(s_base + kOffW2 + t * (64 * 128));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(128 + b * 64, da + 2 * ks, db + 2 * ks, idesc, (t | ks) != 0);
        }
        umma_commit(&ctl->y_empty[b]);
        umma_commit(&ctl->acc2_full[b]);
      }
      __syncwarp();
    };
    if (n_my > 0) mma1(0);
    for (int k = 0; k < n_my; ++k) {
      if (k + 1 < n_my) mma1(k + 1);
      mma2(k);
    }
  } else {
    // ------------------------------------------------------------ epilogues: warps 2..9 run epilogue 1 of every tile,
    // warps 10..17 epilogue 2
    const int ewarp = (warp - 2) & (kEpiWarps - 1);
    const int quarter = warp & 3;
    const int half = ewarp >> 2;                 // which 32 of the 64 channels
    const int row = quarter * 32 + lane;         // TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_base = smem_u32(smem);
    const float* sh1 = s_shift + half * 32;
    const float* sh2 = s_shift + kC + half * 32;
    auto epi1 = [&](int k) {
      const uint32_t b = k & 1, ph = (k >> 1) & 1;
      int n, h0, w0;
      tile_coord(k, n, h0, w0);
      // accumulator row = (hh, ww) with ww fastest; Y row = (ww, hh) with hh fastest
      const int hh = row >> 4, ww = row & 15;
      const int wg = w0 - 1 + ww;
      const bool zero = wg < 0 || wg >= args.w;          // the 1x3 conv sees zero padding there
      mbar_wait(&ctl->acc1_full[b], ph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(t_lane + b * 64 + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc1_empty[b]);
      mbar_wait(&ctl->y_empty[b], ph ^ 1);               // MMA2 of tile k-2 has read this Y buffer
      const int yrow = ww * kTH + hh;
      const uint32_t dst = s_base + kOffY + b * kYBytes + yrow * 128;
      const uint32_t swz = yrow & 7;
      if constexpr (!kRot) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[j + e]) + sh1[j + e], 0.f);
          uint4 o;
          o.x = pack_bf16(f[0], f[1]);
          o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]);
          o.w = pack_bf16(f[6], f[7]);
          if (zero) o = make_uint4(0u, 0u, 0u, 0u);
          sts128(dst + ((((half * 4 + (j >> 3)) ^ swz)) << 4), o);
        }
      } else {
        // Default since round 2 (measured 20.3 -> 17.8 us per pair, bit-identical).  The 8 lanes of a quarter-warp write Y rows 8 apart
        // (ww consecutive, same hh): 1024-byte stride and the same swizzle term, i.e. the same 4 banks for a given
        // chunk -> 8-way conflicts on every 16-byte store above.  Here lane ww writes its four chunks in the order
        // (i + ww) & 3, so a quarter-warp covers all four chunk positions in every store (2-way instead of 8-way).
        // Same bytes at the same addresses: the result is unchanged.
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[j + e]) + sh1[j + e], 0.f);
          o[j >> 3].x = pack_bf16(f[0], f[1]);
          o[j >> 3].y = pack_bf16(f[2], f[3]);
          o[j >> 3].z = pack_bf16(f[4], f[5]);
          o[j >> 3].w = pack_bf16(f[6], f[7]);
          if (zero) o[j >> 3] = make_uint4(0u, 0u, 0u, 0u);
        }
        const uint32_t r = ww & 3;
        const bool r1 = r & 1, r2 = r & 2;
        uint4 p[4], q[4];
        auto sel = [](bool c, const uint4& a, const uint4& b) {
          return make_uint4(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z, c ? a.w : b.w);
        };
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = sel(r1, o[(k + 1) & 3], o[k]);        // rotate by (r & 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = sel(r2, p[(k + 2) & 3], p[k]);        // then by (r & 2): q[k] = o[(k + r) & 3]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t c = (i + r) & 3;
          sts128(dst + ((((half * 4 + c) ^ swz)) << 4), q[i]);
        }
      }
      fence_async_smem();                                // generic writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->y_full[b]);
    };
    int sbuf = 0;
    auto epi2 = [&](int k) {
      const uint32_t b = k & 1, ph = (k >> 1) & 1;
      int n, h0, w0;
      tile_coord(k, n, h0, w0);
      mbar_wait(&ctl->acc2_full[b], ph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(t_lane + 128 + b * 64 + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc2_empty[b]);
      if (args.has_res) mbar_wait(&ctl->res_full[b], ph);
      const uint32_t swz = row & 7;
      const uint32_t res = s_base + kOffRes + b * kOutPitch + row * 128;
      const uint32_t dst = s_base + kOffStage + sbuf * kOutPitch + row * 128;
      if (row < kOutRows) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]) + sh2[j + e];
          const uint32_t chunk = ((half * 4 + (j >> 3)) ^ swz) << 4;
          if (args.has_res) {
            const uint4 r = lds128(res + chunk);
            f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x); f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
            f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z); f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
          }
          if (args.relu2) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          uint4 o;
          o.x = pack_bf16(f[0], f[1]);
          o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]);
          o.w = pack_bf16(f[6], f[7]);
          sts128(dst + chunk, o);
        }
      }
      if (args.has_res) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->res_empty[b]);
      }
      fence_async_smem();
      if (ewarp == 0 && elect_one()) bulk_wait_read<0>();        // the other staging buffer's store has drained
      named_barrier(1, 32 * kEpiWarps);
      if (ewarp == 0 && elect_one()) {
        tma_store_4d(&map_out, smem + kOffStage + sbuf * kOutPitch, 0, h0, w0, n);
        bulk_commit();
      }
      sbuf ^= 1;
    };
    if (warp < 2 + kEpiWarps) {
      for (int k = 0; k < n_my; ++k) epi1(k);
    } else {
      for (int k = 0; k < n_my; ++k) epi2(k);
      if (ewarp == 0 && elect_one()) bulk_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(0, 512);
  }
}

}  // namespace pairk
}  // namespace dynmm

using namespace dynmm;

extern "C" int dynmm_conv_pair_fwd(const dynmm_conv_pair_params* p, void* stream_) {
  using namespace pairk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(p && p->in && p->w1 && p->w2 && p->out, "conv_pair: null pointer");
  DYNMM_CHECK_ARG(p->n >= 1 && p->h >= 1 && p->w >= 1, "conv_pair: empty tensor");
  DYNMM_CHECK_ARG(p->in_ld >= kC && p->out_ld >= kC && p->in_ld % 8 == 0 && p->out_ld % 8 == 0, "conv_pair: 64 channels, ld %% 8");
  DYNMM_CHECK_ARG(!p->residual || (p->res_ld >= kC && p->res_ld % 8 == 0), "conv_pair: res_ld");
  DYNMM_CHECK_ARG(((reinterpret_cast<uintptr_t>(p->in) | reinterpret_cast<uintptr_t>(p->out) |
                    reinterpret_cast<uintptr_t>(p->w1) | reinterpret_cast<uintptr_t>(p->w2) |
                    reinterpret_cast<uintptr_t>(p->residual)) & 15) == 0, "conv_pair: pointers must be 16-byte aligned");
  const uint64_t es = 2;
  CUtensorMap map_x, map_w1, map_w2, map_res, map_out;
  {
    // input: (c, W, H, N), box 64 x 16 x 10: rows of the tile = (hh, ww), ww fastest
    const uint64_t dims[4] = {(uint64_t)kC, (uint64_t)p->w, (uint64_t)p->h, (uint64_t)p->n_in};
    const uint64_t str[3] = {(uint64_t)p->in_ld * es, (uint64_t)p->in_ld * p->w * es, (uint64_t)p->in_ld * p->w * p->h * es};
    const uint32_t box[4] = {(uint32_t)kC, (uint32_t)kXW, (uint32_t)kXH, 1};
    int rc = encode_map(&map_x, p->in, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)kC, (uint64_t)kC, 3};
    const uint64_t str[2] = {(uint64_t)kC * es, (uint64_t)kC * kC * es};
    const uint32_t box[3] = {(uint32_t)kC, (uint32_t)kC, 3};
    int rc = encode_map(&map_w1, p->w1, 3, dims, str, box);
    if (rc) return rc;
    rc = encode_map(&map_w2, p->w2, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    // output / residual: (c, H, W, N), box 64 x 8 x 14: rows of the tile = (wo, h), h fastest
    const uint32_t box[4] = {(uint32_t)kC, (uint32_t)kTH, (uint32_t)kTW, 1};
    const uint64_t dims[4] = {(uint64_t)kC, (uint64_t)p->h, (uint64_t)p->w, (uint64_t)p->n};
    const uint64_t str[3] = {(uint64_t)p->out_ld * p->w * es, (uint64_t)p->out_ld * es, (uint64_t)p->out_ld * p->w * p->h * es};
    int rc = encode_map(&map_out, p->out, 4, dims, str, box);
    if (rc) return rc;
    map_res = map_out;
    if (p->residual) {
      const uint64_t rdims[4] = {(uint64_t)kC, (uint64_t)p->h, (uint64_t)p->w, (uint64_t)(p->res_map ? 65536 : p->n)};
      const uint64_t rstr[3] = {(uint64_t)p->res_ld * p->w * es, (uint64_t)p->res_ld * es, (uint64_t)p->res_ld * p->w * p->h * es};
      rc = encode_map(&map_res, p->residual, 4, rdims, rstr, box);
      if (rc) return rc;
    }
  }
  Args a{};
  a.n = p->n;
  a.h = p->h;
  a.w = p->w;
  a.tiles_h = ceil_div(p->h, kTH);
  a.tiles_w = ceil_div(p->w, kTW);
  a.relu2 = p->relu2;
  a.has_res = p->residual ? 1 : 0;
  static const bool rot = [] {
    // rotated chunk order in epilogue 1: 20.3 -> 17.8 us per pair at B = 8 (gpurun_out/r2a, round 2); DYNMM_PAIR_ROT=0
    // keeps the straight order for comparison
    const char* e = getenv("DYNMM_PAIR_ROT");
    return !(e && e[0] == '0');
  }();
  a.shift1 = p->shift1;
  a.shift2 = p->shift2;
  a.count = p->count;
  a.in_map = p->in_map;
  a.res_map = p->res_map;
  const long long max_tiles = 1LL * p->n * a.tiles_h * a.tiles_w;
  DYNMM_CHECK_ARG(max_tiles < (1LL << 30), "conv_pair: too many tiles");
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(conv_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    return e != cudaSuccess ? e : cudaFuncSetAttribute(conv_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }));
  int grid = num_sms();
  if (grid > max_tiles) grid = (int)max_tiles;
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
  if (rot) {
    DYNMM_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<true>, map_x, map_w1, map_w2, map_res, map_out, a));
  } else {
    DYNMM_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<false>, map_x, map_w1, map_w2, map_res, map_out, a));
  }
  return DYNMM_OK;
}
