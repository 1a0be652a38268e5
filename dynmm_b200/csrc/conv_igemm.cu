// Implicit-GEMM convolution for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
// GEMM view (NHWC bf16):  M = 128 output pixels of a (d1, d2, sample) box, N = output
// channels, K = taps x input channels.  No im2col buffer exists: the A operand of a tap
// is the activation tensor itself, shifted.  Two ways a shift is realised:
//
//   * halo mode (stride 1; 1x3, 3x1, 3x3, 1x1): ONE TMA box load brings a tile with a
//     2-pixel halo along d2 into shared memory and the three taps along d2 are three
//     UMMA descriptors into the same tile, b1 rows (a multiple of 8 rows = 1024 B, so the
//     128B-swizzle phase is preserved) apart.  d1 is the fastest pixel index in shared
//     memory, so a 1x3 conv uses (d1,d2) = (H,W) and a 3x1 conv (W,H).  The three
//     kernel rows of a 3x3 conv are three such loads.  L2->SM operand traffic drops from
//     3x (9x) the tile to (b2+2)/b2.
//   * generic mode (strided convs, tiny maps): one TMA load per tap at shifted coordinates;
//     strided convolutions read a parity sub-lattice through a strided tensor map.
//   In both, out-of-image rows/columns are zero-filled by TMA: padding costs nothing.
//
//   Weights are either streamed with the A tiles or, when one channel tile covers c_out and
//   all taps fit (<= ~100 KB: C <= 128), loaded into shared memory ONCE per CTA -- before
//   the programmatic-dependent-launch wait, i.e. overlapped with the previous kernel.
//
// One persistent CTA per SM, warp-specialised (320 threads):
//   warp 0      TMA producer  (A tiles, streamed weights, residual sub-tiles one tile ahead)
//   warp 1      MMA issuer    (one elected lane; UMMA 128 x tile_n x 16)
//   warps 2..9  epilogue      (TMEM -> registers -> shift, residual, ReLU, gated depth add ->
//                              bf16 -> swizzled smem staging -> TMA store)
// Two TMEM accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
// The tile list is derived on the device from `count` (number of active sample slots), so
// samples the gate switched off generate no TMA traffic and no MMA work, and the launch is
// CUDA-graph capturable.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tma_host.cuh"

namespace dynmm {

namespace {

constexpr int kBlockM = 128;       // UMMA M (TMEM lanes)
constexpr int kBlockK = 64;        // bf16 elements per 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kMaxGroups = 9;
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // 320
constexpr int kSmemBudget = 227 * 1024;
constexpr int kSubBytes = kBlockM * kBlockK * 2; // 16 KiB: one [128 rows][64 ch] bf16 epilogue sub-tile
constexpr int kAuxSlots = 3;                     // residual sub-tiles in flight
constexpr int kResidentBudget = 100 * 1024;      // weights kept in shared memory when they fit

// one A-tile load: `tpg` taps (along d2) share it
struct Group {
  int8_t map;   // which A tensor map (parity sub-lattice)
  int8_t o1;    // coordinate offset along d1
  int8_t o2;    // coordinate offset along d2 (start of the halo in halo mode)
  int8_t pad;
};

struct KernelArgs {
  // tiling.  (d1, d2) = (W, H) or (H, W) when `swap` -- d1 is the fastest pixel index of a tile.
  int b1, b2, bn;                   // pixels per tile = b1*b2*bn <= 128
  int tiles1, tiles2;               // tiles per sample group along d1 / d2
  int c_tiles, tile_n;              // output channel tiles
  int num_groups, tpg, k_chunks;    // A loads per K chunk; taps per load (1 or 3)
  int stages, stage_bytes, a_bytes; // pipeline; a_bytes = A part of a stage (1024-aligned)
  int a_rows;                       // rows TMA writes per A load (b1 * (b2 + tpg - 1) * bn)
  int acc_stride, tmem_cols;
  uint32_t m_c, m_1, m_2;           // magic multipliers for dividing by c_tiles / tiles1 / tiles2
  int tma_epi;                      // 1: TMA residual loads + TMA stores (tile_n % 64 == 0)
  int aux_slots;                    // residual ring slots (0 without residual)
  int b_resident;                   // 1: all weights live in smem for the kernel's lifetime
  int swap;                         // 1: d1 = H, d2 = W
  Group groups[kMaxGroups];
  // problem
  int n, h_out, w_out, c_out;
  int out_ld, res_ld, gated_ld;
  const float* shift;
  const float* scale;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  const __nv_bfloat16* gated;
  const float* gate;
  const int32_t* gated_slot;
  const int32_t* in_map;
  const int32_t* res_map;
  const int32_t* count;
  unsigned long long* trace;        // debug: 16 cycle stamps per CTA, or NULL
};

struct __align__(8) SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t aux_full[kAuxSlots];
  uint64_t aux_empty[kAuxSlots];
  uint64_t b_full;
  uint32_t tmem_base;
};

struct TileCoord {
  int c0, x1, x2, n0;               // channel, d1, d2, sample origin of a tile
};

// x / d for x*d < 2^32 with m = ceil(2^32 / d) (host-computed); d == 1 has m == 0
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t m) { return m ? __umulhi(x, m) : x; }

__device__ __forceinline__ TileCoord decode_tile(const KernelArgs& a, int tile) {
  TileCoord t;
  uint32_t r = fast_div(tile, a.m_c);
  const int ct = tile - r * a.c_tiles;
  uint32_t q = fast_div(r, a.m_1);
  const int t1 = r - q * a.tiles1;
  r = fast_div(q, a.m_2);
  const int t2 = q - r * a.tiles2;
  t.c0 = ct * a.tile_n;
  t.x1 = t1 * a.b1;
  t.x2 = t2 * a.b2;
  t.n0 = r * a.bn;
  return t;
}

enum : int { kFlagRes = 1, kFlagGated = 2, kFlagRelu = 4, kFlagScale = 8 };

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// One thread's share of an epilogue sub-tile: 32 consecutive output channels of one pixel.
//   kTma: residual comes from the swizzled smem tile `res_smem`, result goes to the swizzled
//         staging tile `out_smem` (both 32-bit shared addresses of this thread's row);
//   else: direct global loads / stores (narrow channel tiles, partially active sample boxes).
template <int kFlags, bool kTma>
__device__ __forceinline__ void epilogue_chunk(const KernelArgs& args, const uint32_t (&v)[32], int c_first,
                                               int cols_left, bool valid, uint32_t res_smem, uint32_t out_smem,
                                               uint32_t chunk0, uint32_t swz, size_t pix, size_t rpix, size_t gpix,
                                               float g, const float* shift_smem) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c_first + j;
    if (j < cols_left && c < args.c_out) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
      if (kFlags & kFlagScale) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(args.scale + c));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(args.scale + c + 4));
        f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
        f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
      }
      {
        const float4 b0 = *reinterpret_cast<const float4*>(shift_smem + c);      // warp-wide broadcast
        const float4 b1 = *reinterpret_cast<const float4*>(shift_smem + c + 4);
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      const uint32_t chunk = ((chunk0 + (j >> 3)) ^ swz) << 4;
      if (kFlags & kFlagRes) {
        uint4 r;
        if (kTma) {
          r = lds128(res_smem + chunk);
        } else {
          r = valid ? __ldg(reinterpret_cast<const uint4*>(args.residual + rpix * args.res_ld + c))
                    : make_uint4(0, 0, 0, 0);
        }
        f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x); f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
        f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z); f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
      }
      if (kFlags & kFlagRelu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
      }
      if (kFlags & kFlagGated) {
        if (g != 0.f) {      // gated-off samples never touch the depth features
          const uint4 r = __ldg(reinterpret_cast<const uint4*>(args.gated + gpix * args.gated_ld + c));
          f[0] += g * bf16_lo(r.x); f[1] += g * bf16_hi(r.x); f[2] += g * bf16_lo(r.y); f[3] += g * bf16_hi(r.y);
          f[4] += g * bf16_lo(r.z); f[5] += g * bf16_hi(r.z); f[6] += g * bf16_lo(r.w); f[7] += g * bf16_hi(r.w);
        }
      }
      uint4 o;
      o.x = pack_bf16(f[0], f[1]);
      o.y = pack_bf16(f[2], f[3]);
      o.z = pack_bf16(f[4], f[5]);
      o.w = pack_bf16(f[6], f[7]);
      if (kTma) {
        sts128(out_smem + chunk, o);
      } else if (valid) {
        *reinterpret_cast<uint4*>(args.out + pix * args.out_ld + c) = o;
      }
    }
  }
}

#define DYNMM_TRACE(slot)                                                              \
  do {                                                                                 \
    if (args.trace) args.trace[blockIdx.x * 16 + (slot)] = (unsigned long long)clock64(); \
  } while (0)

template <int kFlags>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_res,
                  const __grid_constant__ CUtensorMap map_out, const __grid_constant__ KernelArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int k_iters = args.num_groups * args.k_chunks;             // A loads per tile
  const int b_tile_bytes = args.tile_n * kBlockK * 2;              // one tap's [tile_n][64] weight tile
  const int b_iter_bytes = args.tpg * b_tile_bytes;                // weights consumed per A load
  uint8_t* smem_bres = smem + args.stages * args.stage_bytes;      // resident weights [k_iters][tpg][tile_n][128 B]
  uint8_t* smem_aux = smem_bres + (args.b_resident ? k_iters * b_iter_bytes : 0);   // [aux_slots][16 KiB]
  uint8_t* smem_stage_out = smem_aux + args.aux_slots * kSubBytes;                  // [2][16 KiB] (tma_epi only)
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_stage_out + (args.tma_epi ? 2 * kSubBytes : 0));
  // [c_out]: per-channel shift (zeros if absent), 16-byte aligned for float4 broadcast loads
  float* smem_shift = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ctl + 1) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) DYNMM_TRACE(0);
  const int n_sub = (args.tile_n + 63) >> 6;
  const bool aux_on = (kFlags & kFlagRes) && args.tma_epi && args.aux_slots > 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_b);
    if (args.tma_epi) tma_prefetch_desc(&map_out);
    if (aux_on) tma_prefetch_desc(&map_res);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], kEpiWarps);   // one arrive per epilogue warp
    }
    for (int i = 0; i < kAuxSlots; ++i) {
      mbar_init(&ctl->aux_full[i], 1);
      mbar_init(&ctl->aux_empty[i], kEpiWarps);
    }
    mbar_init(&ctl->b_full, 1);
    fence_mbar_init();
    if (args.b_resident) {
      // weights are constants: fetch them before waiting on the previous kernel (c_tiles == 1)
      mbar_expect_tx(&ctl->b_full, k_iters * b_iter_bytes);
      for (int g = 0; g < args.num_groups; ++g)
        for (int kc = 0; kc < args.k_chunks; ++kc)
          tma_load_3d(smem_bres + (g * args.k_chunks + kc) * b_iter_bytes, &map_b, &ctl->b_full, kc * kBlockK, 0,
                      g * args.tpg);
    }
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, args.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2) {   // epilogue warps stage the shift vector once: no global loads inside the tile loop
    for (int c = threadIdx.x - 64; c < args.c_out; c += 32 * kEpiWarps) smem_shift[c] = args.shift ? args.shift[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch, shift
  // staging, resident weights -- constants only) overlapped the tail of the previous kernel in the
  // stream.  From here on we touch tensors it produced, so wait for it; then let OUR dependent
  // start its prologue.
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  const int active = args.count ? min(*args.count, args.n) : args.n;
  const int n_groups = (active + args.bn - 1) / args.bn;
  const int total_tiles = n_groups * args.tiles2 * args.tiles1 * args.c_tiles;
  if (threadIdx.x == 0) DYNMM_TRACE(1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const CUtensorMap* maps[4] = {&map_a0, &map_a1, &map_a2, &map_a3};
      // TMA always delivers the full box (out-of-bounds elements arrive as zeros)
      const uint32_t a_tx = args.a_rows * kBlockK * 2;
      const uint32_t tx_bytes = a_tx + (args.b_resident ? 0 : b_iter_bytes);
      const uint32_t sub_tx = args.b1 * args.b2 * args.bn * kBlockK * 2;
      int stage = 0;
      uint32_t phase = 0;
      int aux = 0;
      uint32_t aux_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(args, tile);
        const int n_in = args.in_map ? args.in_map[t.n0] : t.n0;
        for (int g = 0; g < args.num_groups; ++g) {
          const Group gp = args.groups[g];
          for (int kc = 0; kc < args.k_chunks; ++kc) {
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * args.stage_bytes;
            mbar_expect_tx(&ctl->full[stage], tx_bytes);
            tma_load_4d(sa, maps[gp.map], &ctl->full[stage], kc * kBlockK, t.x1 + gp.o1, t.x2 + gp.o2, n_in);
            if (!args.b_resident)
              tma_load_3d(sa + args.a_bytes, &map_b, &ctl->full[stage], kc * kBlockK, t.c0, g * args.tpg);
            if (tile == (int)blockIdx.x && g == 0 && kc == 0) DYNMM_TRACE(2);
            if (++stage == args.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (tile == (int)blockIdx.x) DYNMM_TRACE(11);
        // residual sub-tiles of this tile, consumed by the epilogue while the next tile's MMAs run
        if (aux_on && t.n0 + args.bn <= active) {
          const int n_res = args.res_map ? args.res_map[t.n0] : t.n0;
          for (int sub = 0; sub < n_sub; ++sub) {
            mbar_wait(&ctl->aux_empty[aux], aux_phase ^ 1);
            mbar_expect_tx(&ctl->aux_full[aux], sub_tx);
            tma_load_4d(smem_aux + aux * kSubBytes, &map_res, &ctl->aux_full[aux], t.c0 + sub * 64, t.x1, t.x2,
                        n_res);
            if (++aux == args.aux_slots) {
              aux = 0;
              aux_phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, args.tile_n);
      const uint32_t tap_step = args.b1 * kBlockK * 2;      // bytes between the A views of consecutive taps
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      if (args.b_resident) {
        // also when this CTA has no tiles: the bulk loads must land before the CTA may exit
        mbar_wait(&ctl->b_full, 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * args.acc_stride;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          if (local == 0 && it == 0) DYNMM_TRACE(3);
          const uint32_t sa = smem_u32(smem + stage * args.stage_bytes);
          const uint32_t sb = args.b_resident ? smem_u32(smem_bres + it * b_iter_bytes) : sa + args.a_bytes;
          for (int tp = 0; tp < args.tpg; ++tp) {
            const uint64_t da = umma_desc_sw128(sa + tp * tap_step);
            const uint64_t db = umma_desc_sw128(sb + tp * b_tile_bytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advancing K inside the swizzle atom = +32 bytes on the (16-byte unit) start address
              umma_bf16(d_tmem, da + (k * 2), db + (k * 2), idesc, (it | tp | k) != 0);
            }
          }
          if (local == 0 && it == 0) DYNMM_TRACE(12);
          umma_commit(&ctl->empty[stage]);          // frees the smem stage when these MMAs retire
          if (it == k_iters - 1) umma_commit(&ctl->acc_full[acc]);
          if (++stage == args.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (local == 0) DYNMM_TRACE(4);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (8 warps)
    const int ewarp = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31 belong to this warp
    const int half = ewarp >> 2;                  // which 32 of the 64 columns of a sub-tile
    const int row = quarter * 32 + lane;          // GEMM row == pixel inside the box, d1 fastest
    const int i1 = row % args.b1;
    const int i2 = (row / args.b1) % args.b2;
    const int nl = row / (args.b1 * args.b2);
    const bool leader = (threadIdx.x == 64);
    const uint32_t row_off = row * 128;           // byte offset of this row inside a [128][128 B] tile
    const uint32_t swz = row & 7;                 // 128B swizzle: 16-byte chunk index ^= row % 8
    const uint32_t aux_base = smem_u32(smem_aux) + row_off;
    const uint32_t out_base = smem_u32(smem_stage_out) + row_off;
    int local = 0;
    int aux = 0;
    uint32_t aux_phase = 0;
    int sbuf = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const TileCoord t = decode_tile(args, tile);
      const int n = t.n0 + nl;
      const int p1 = t.x1 + i1, p2 = t.x2 + i2;
      const int h = args.swap ? p1 : p2, w = args.swap ? p2 : p1;
      const bool valid = nl < args.bn && n < active && h < args.h_out && w < args.w_out;
      const bool tile_tma = args.tma_epi && (t.n0 + args.bn <= active);   // uniform over the CTA
      const size_t pix = valid ? (static_cast<size_t>(n) * args.h_out + h) * args.w_out + w : 0;
      size_t rpix = pix;
      if ((kFlags & kFlagRes) && valid && args.res_map) {
        rpix = (static_cast<size_t>(args.res_map[n]) * args.h_out + h) * args.w_out + w;
      }
      float g = 0.f;
      size_t gpix = 0;
      if ((kFlags & kFlagGated) && valid) {
        g = args.gate[n];
        const int slot = args.gated_slot ? args.gated_slot[n] : n;
        gpix = (static_cast<size_t>(slot) * args.h_out + h) * args.w_out + w;
      }
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after();
      if (leader && local == 0) DYNMM_TRACE(5);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * args.acc_stride;
      for (int sub = 0; sub < n_sub; ++sub) {
        const int cb = sub * 64 + half * 32;       // first of this thread's 32 columns inside the tile
        const bool cols_live = cb < args.acc_stride;
        uint32_t v[32];
        if (cols_live) {                            // warp-uniform
          tmem_ld32(t_row + cb, v);
          tmem_ld_wait();
        }
        if (leader && local == 0 && sub == 0) DYNMM_TRACE(13);
        if (tile_tma) {
          uint32_t res_smem = 0;
          if (aux_on) {
            mbar_wait(&ctl->aux_full[aux], aux_phase);
            res_smem = aux_base + aux * kSubBytes;
          }
          const uint32_t out_smem = out_base + sbuf * kSubBytes;
          if (cols_live) {
            epilogue_chunk<kFlags, true>(args, v, t.c0 + cb, args.tile_n - cb, valid, res_smem, out_smem, half * 4,
                                         swz, pix, rpix, gpix, g, smem_shift);
          }
          if (aux_on) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->aux_empty[aux]);
            if (++aux == args.aux_slots) {
              aux = 0;
              aux_phase ^= 1;
            }
          }
          if (leader && local == 0 && sub == 0) DYNMM_TRACE(14);
          fence_async_smem();                   // staging writes -> visible to the TMA engine
          if (leader) bulk_wait_read<0>();      // the previous store (other buffer) has drained its smem
          named_barrier(1, 32 * kEpiWarps);
          if (leader && local == 0 && sub == 0) DYNMM_TRACE(15);
          if (leader) {
            tma_store_4d(&map_out, smem_stage_out + sbuf * kSubBytes, t.c0 + sub * 64, t.x1, t.x2, t.n0);
            bulk_commit();
          }
          sbuf ^= 1;
        } else if (cols_live) {
          epilogue_chunk<kFlags, false>(args, v, t.c0 + cb, args.tile_n - cb, valid, 0, 0, 0, 0, pix, rpix, gpix, g,
                                        smem_shift);
        }
      }
      // this warp is done reading the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc_empty[acc]);
      if (leader && local == 0) DYNMM_TRACE(6);
    }
    if (leader) {
      DYNMM_TRACE(7);
      bulk_wait<0>();
      DYNMM_TRACE(8);
      if (args.trace) args.trace[blockIdx.x * 16 + 10] = local;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, args.tmem_cols);
  }
  if (threadIdx.x == 0) DYNMM_TRACE(9);
}

// ------------------------------------------------------------------ host side

// floor division for possibly negative tap offsets
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// generic mode: the (w,h,n) pixel box of a tile, maximising useful rows per 128-row MMA
void choose_box(int w, int h, int n, bool single_sample, int* bw, int* bh, int* bn) {
  double best = -1;
  for (int cw = 1; cw <= w && cw <= kBlockM; ++cw) {
    for (int ch = 1; ch <= h && cw * ch <= kBlockM; ++ch) {
      int cn = single_sample ? 1 : kBlockM / (cw * ch);
      if (cn > n) cn = n;
      if (cn < 1) cn = 1;
      long long tiles = 1LL * ceil_div(w, cw) * ceil_div(h, ch) * ceil_div(n, cn);
      double eff = (double)w * h * n / (double)(tiles * kBlockM);
      double score = eff + 1e-6 * cw;      // prefer wide boxes (contiguous NHWC rows) on ties
      if (score > best) {
        best = score;
        *bw = cw;
        *bh = ch;
        *bn = cn;
      }
    }
  }
}
double box_eff(int w, int h, int n, int bw, int bh, int bn) {
  long long tiles = 1LL * ceil_div(w, bw) * ceil_div(h, bh) * ceil_div(n, bn);
  return (double)w * h * n / (double)(tiles * kBlockM);
}

// halo mode: b1 in {8,16,32} rows along d1 (multiple of 8 keeps tap views 1024-byte aligned), b2 = 128 / b1
void choose_halo_box(int d1, int d2, bool tapped, int* b1, int* b2, double* eff_out) {
  double best = -1;
  for (int c1 = 8; c1 <= (tapped ? 32 : 128); c1 *= 2) {
    const int c2 = kBlockM / c1;
    const double eff = (double)d1 * d2 / ((double)ceil_div(d1, c1) * c1 * ceil_div(d2, c2) * c2);
    const double halo = tapped ? (double)(c2 + 2) / c2 : 1.0;    // operand bytes per useful row
    const double score = eff / halo;
    if (score > best) {
      best = score;
      *b1 = c1;
      *b2 = c2;
      *eff_out = eff;
    }
  }
}

}  // namespace

}  // namespace dynmm

using namespace dynmm;

extern "C" int dynmm_conv_igemm_fwd(const dynmm_conv_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(p && p->in && p->weight && p->out, "conv_igemm: null pointer");
  DYNMM_CHECK_ARG(p->kh >= 1 && p->kw >= 1 && p->kh * p->kw <= kMaxGroups, "conv_igemm: at most %d taps", kMaxGroups);
  DYNMM_CHECK_ARG(p->stride_h >= 1 && p->stride_h <= 2 && p->stride_w >= 1 && p->stride_w <= 2,
                  "conv_igemm: stride must be 1 or 2");
  DYNMM_CHECK_ARG(p->c_in % 8 == 0 && p->in_ld % 8 == 0 && p->in_ld >= p->c_in, "conv_igemm: c_in/in_ld %% 8");
  DYNMM_CHECK_ARG(p->c_out % 8 == 0 && p->out_ld % 8 == 0 && p->out_ld >= p->c_out, "conv_igemm: c_out/out_ld %% 8");
  DYNMM_CHECK_ARG(p->c_out <= 4096, "conv_igemm: c_out too large");
  DYNMM_CHECK_ARG(!p->residual || p->res_ld % 8 == 0, "conv_igemm: res_ld %% 8");
  DYNMM_CHECK_ARG(!p->gated || (p->gated_ld % 8 == 0 && p->gate), "conv_igemm: gated needs gate[] and gated_ld %% 8");
  DYNMM_CHECK_ARG(p->n >= 1 && p->n_in >= 1, "conv_igemm: empty batch");
  const int h_exp = (p->h_in + 2 * p->pad_h - p->kh) / p->stride_h + 1;
  const int w_exp = (p->w_in + 2 * p->pad_w - p->kw) / p->stride_w + 1;
  DYNMM_CHECK_ARG(h_exp == p->h_out && w_exp == p->w_out, "conv_igemm: output size %dx%d does not match %dx%d",
                  p->h_out, p->w_out, h_exp, w_exp);
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(p->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p->weight) & 15) == 0,
                  "conv_igemm: pointers must be 16-byte aligned");

  KernelArgs a{};
  const uint64_t es = 2;
  const int c_out_pad = (p->c_out + 15) / 16 * 16;
  const int sms = num_sms();
  const int num_taps = p->kh * p->kw;
  const bool single_sample = p->in_map != nullptr || p->res_map != nullptr;

  // ---- generic box and its efficiency
  int gw = 1, gh = 1, gn = 1;
  choose_box(p->w_out, p->h_out, p->n, single_sample, &gw, &gh, &gn);
  const double generic_eff = box_eff(p->w_out, p->h_out, p->n, gw, gh, gn);

  // ---- halo mode: unit stride, "same" padding, at most one tapped direction per load
  static const bool allow_halo = [] {
    const char* e = getenv("DYNMM_CONV_HALO");
    return !(e && e[0] == '0');
  }();
  bool halo = allow_halo && p->stride_h == 1 && p->stride_w == 1 && (p->kh == 1 || p->kh == 3) &&
              (p->kw == 1 || p->kw == 3) && p->pad_h == p->kh / 2 && p->pad_w == p->kw / 2;
  int hb1 = 0, hb2 = 0;
  if (halo) {
    // d2 is the direction whose taps share one load: W for 1x3 and 3x3 (swap: d1 = H), H for 3x1
    a.swap = (p->kw == 3) ? 1 : 0;
    const int d1 = a.swap ? p->h_out : p->w_out, d2 = a.swap ? p->w_out : p->h_out;
    double eff = 0;
    choose_halo_box(d1, d2, num_taps > 1, &hb1, &hb2, &eff);
    if (eff < 0.8 * generic_eff) halo = false;      // tiny maps: multi-sample generic boxes fill the MMA better
  }
  int tile_n = 0, m_tiles = 0, b_tile_bytes = 0, b_total = 0, shift_bytes = 0, epi_bytes = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    // second attempt: the halo layout did not leave room for two pipeline stages (wide channel tiles
    // with streamed 3-tap weight tiles) -> per-tap loads
    if (attempt == 1) halo = false;
    if (halo) {
      a.b1 = hb1;
      a.b2 = hb2;
      a.bn = 1;
      a.tpg = (num_taps > 1) ? 3 : 1;
      a.num_groups = num_taps / a.tpg;                 // 1 (1x3, 3x1, 1x1) or 3 (3x3: one load per kernel row)
      for (int g = 0; g < a.num_groups; ++g) {
        a.groups[g].map = 0;
        a.groups[g].o1 = (a.num_groups == 3) ? static_cast<int8_t>(g - 1) : 0;   // 3x3: kernel row ky -> H offset
        a.groups[g].o2 = (a.tpg == 3) ? -1 : 0;                                  // start of the halo
      }
    } else {
      a.swap = 0;
      a.b1 = gw;
      a.b2 = gh;
      a.bn = gn;
      a.tpg = 1;
      a.num_groups = num_taps;
    }
    const int D1 = a.swap ? p->h_out : p->w_out, D2 = a.swap ? p->w_out : p->h_out;
    a.tiles1 = ceil_div(D1, a.b1);
    a.tiles2 = ceil_div(D2, a.b2);
    m_tiles = a.tiles1 * a.tiles2 * ceil_div(p->n, a.bn);

    tile_n = p->tile_n;
    if (tile_n == 0) {
      // multiples of 64 channels (TMA epilogue); widest tile that still gives every SM a tile
      const int c64 = (c_out_pad + 63) / 64 * 64;
      tile_n = c64 < 256 ? c64 : 256;
      if (tile_n == 192) tile_n = 64;
      while (tile_n > 64 && m_tiles * ceil_div(c_out_pad, tile_n) < sms) tile_n /= 2;
    }
    DYNMM_CHECK_ARG(tile_n >= 16 && tile_n <= 256 && tile_n % 16 == 0, "conv_igemm: tile_n %d", tile_n);
    a.tile_n = tile_n;
    a.c_tiles = ceil_div(c_out_pad, tile_n);
    a.k_chunks = ceil_div(p->c_in, kBlockK);
    a.a_rows = a.b1 * (a.b2 + a.tpg - 1) * a.bn;
    a.a_bytes = (a.a_rows * kBlockK * 2 + 1023) / 1024 * 1024;
    if (a.a_bytes < kBlockM * kBlockK * 2) a.a_bytes = kBlockM * kBlockK * 2;   // UMMA reads 128 rows
    if (a.tpg == 3) {
      // the last tap's view spans rows [2*b1, 2*b1 + 128)
      const int need = (2 * a.b1 + kBlockM) * kBlockK * 2;
      if (a.a_bytes < need) a.a_bytes = (need + 1023) / 1024 * 1024;
    }
    b_tile_bytes = tile_n * kBlockK * 2;
    b_total = num_taps * a.k_chunks * b_tile_bytes;
    a.tma_epi = (tile_n % 64 == 0) ? 1 : 0;
    a.aux_slots = (a.tma_epi && p->residual) ? kAuxSlots : 0;
    a.b_resident = (a.c_tiles == 1 && b_total <= kResidentBudget) ? 1 : 0;
    if (a.b_resident && a.aux_slots) a.aux_slots = 2;
    shift_bytes = (p->c_out + 8) * 4 + 16;
    epi_bytes = (a.tma_epi ? 2 * kSubBytes : 0) + a.aux_slots * kSubBytes;
    a.stage_bytes = a.a_bytes + (a.b_resident ? 0 : a.tpg * b_tile_bytes);
    a.stages = (kSmemBudget - 2048 - epi_bytes - shift_bytes - (a.b_resident ? b_total : 0)) / a.stage_bytes;
    if (a.stages < 2 && a.b_resident) {     // not enough room next to the resident weights: stream them instead
      a.b_resident = 0;
      a.aux_slots = (a.tma_epi && p->residual) ? kAuxSlots : 0;
      epi_bytes = (a.tma_epi ? 2 * kSubBytes : 0) + a.aux_slots * kSubBytes;
      a.stage_bytes = a.a_bytes + a.tpg * b_tile_bytes;
      a.stages = (kSmemBudget - 2048 - epi_bytes - shift_bytes) / a.stage_bytes;
    }
    if (a.stages > kMaxStages) a.stages = kMaxStages;
    if (a.stages >= 2 || !halo) break;
  }
  DYNMM_CHECK_ARG(a.stages >= 2, "conv_igemm: not enough shared memory for 2 stages");
  a.acc_stride = (tile_n + 31) / 32 * 32;
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.acc_stride) a.tmem_cols *= 2;
  a.n = p->n;
  a.h_out = p->h_out;
  a.w_out = p->w_out;
  a.c_out = p->c_out;
  a.out_ld = p->out_ld;
  a.res_ld = p->res_ld;
  a.gated_ld = p->gated_ld;
  a.scale = p->scale;
  a.shift = p->shift;
  a.residual = static_cast<const __nv_bfloat16*>(p->residual);
  a.out = static_cast<__nv_bfloat16*>(p->out);
  a.gated = static_cast<const __nv_bfloat16*>(p->gated);
  a.gate = p->gate;
  a.gated_slot = p->gated_slot;
  a.in_map = p->in_map;
  a.res_map = p->res_map;
  a.count = p->count;
  a.trace = static_cast<unsigned long long*>(p->trace);
  auto magic = [](int d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)(((1ULL << 32) + d - 1) / d); };
  a.m_c = magic(a.c_tiles);
  a.m_1 = magic(a.tiles1);
  a.m_2 = magic(a.tiles2);

  // pixel tensor map over an NHWC buffer, dims ordered (c, d1, d2, n)
  auto pixel_map = [&](CUtensorMap* m, const void* base, int c, int w, int h, int n, uint64_t st_w, uint64_t st_h,
                       uint64_t st_n, int box2) -> int {
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)(a.swap ? h : w), (uint64_t)(a.swap ? w : h), (uint64_t)n};
    const uint64_t strides[3] = {a.swap ? st_h : st_w, a.swap ? st_w : st_h, st_n};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)a.b1, (uint32_t)box2, (uint32_t)a.bn};
    return encode_map(m, base, 4, dims, strides, box);
  };

  // A maps
  CUtensorMap maps[4];
  bool used[4] = {false, false, false, false};
  if (halo) {
    used[0] = true;
    int rc = pixel_map(&maps[0], p->in, p->c_in, p->w_in, p->h_in, p->n_in, (uint64_t)p->in_ld * es,
                       (uint64_t)p->in_ld * p->w_in * es, (uint64_t)p->in_ld * p->w_in * p->h_in * es,
                       a.b2 + a.tpg - 1);
    if (rc) return rc;
  } else {
    // one map per (parity_h, parity_w) sub-lattice of the input
    for (int ky = 0; ky < p->kh; ++ky) {
      for (int kx = 0; kx < p->kw; ++kx) {
        const int dy = ky - p->pad_h, dx = kx - p->pad_w;
        const int qy = floordiv(dy, p->stride_h), py = dy - qy * p->stride_h;
        const int qx = floordiv(dx, p->stride_w), px = dx - qx * p->stride_w;
        Group& t = a.groups[ky * p->kw + kx];
        t.map = static_cast<int8_t>(py * p->stride_w + px);
        t.o1 = static_cast<int8_t>(qx);
        t.o2 = static_cast<int8_t>(qy);
        used[t.map] = true;
      }
    }
    for (int m = 0; m < 4; ++m) {
      if (!used[m]) continue;
      const int py = m / p->stride_w, px = m % p->stride_w;
      const int sub_w = (p->w_in - px + p->stride_w - 1) / p->stride_w;
      const int sub_h = (p->h_in - py + p->stride_h - 1) / p->stride_h;
      DYNMM_CHECK_ARG(sub_w >= 1 && sub_h >= 1, "conv_igemm: input too small for stride");
      const __nv_bfloat16* base =
          static_cast<const __nv_bfloat16*>(p->in) + (static_cast<size_t>(py) * p->w_in + px) * p->in_ld;
      int rc = pixel_map(&maps[m], base, p->c_in, sub_w, sub_h, p->n_in, (uint64_t)p->in_ld * p->stride_w * es,
                         (uint64_t)p->in_ld * p->w_in * p->stride_h * es,
                         (uint64_t)p->in_ld * p->w_in * p->h_in * es, a.b2);
      if (rc) return rc;
    }
  }
  int first_used = 0;
  while (!used[first_used]) ++first_used;
  for (int m = 0; m < 4; ++m)
    if (!used[m]) maps[m] = maps[first_used];
  CUtensorMap map_b;
  {
    const uint64_t dims[3] = {(uint64_t)p->c_in, (uint64_t)c_out_pad, (uint64_t)num_taps};
    const uint64_t strides[2] = {(uint64_t)p->c_in * es, (uint64_t)p->c_in * c_out_pad * es};
    const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)tile_n, (uint32_t)a.tpg};
    int rc = encode_map(&map_b, p->weight, 3, dims, strides, box);
    if (rc) return rc;
  }
  // epilogue maps: residual (load) and output (store), one [box pixels][64 channels] sub-tile per transfer
  CUtensorMap map_res = map_b, map_out = map_b;
  if (a.tma_epi) {
    int rc = pixel_map(&map_out, p->out, p->c_out, p->w_out, p->h_out, p->n, (uint64_t)p->out_ld * es,
                       (uint64_t)p->out_ld * p->w_out * es, (uint64_t)p->out_ld * p->w_out * p->h_out * es, a.b2);
    if (rc) return rc;
    if (a.aux_slots) {
      DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(p->residual) & 15) == 0,
                      "conv_igemm: residual must be 16-byte aligned");
      // the residual may hold more samples than n (res_map gathers); the sample extent only bounds the box
      rc = pixel_map(&map_res, p->residual, p->c_out, p->w_out, p->h_out, p->res_map ? 65536 : p->n,
                     (uint64_t)p->res_ld * es, (uint64_t)p->res_ld * p->w_out * es,
                     (uint64_t)p->res_ld * p->w_out * p->h_out * es, a.b2);
      if (rc) return rc;
    }
  }

  const int smem_bytes = a.stages * a.stage_bytes + (a.b_resident ? b_total : 0) + epi_bytes + 1024 /*align*/ +
                         (int)sizeof(SmemCtl) + shift_bytes;
  DYNMM_CHECK_ARG(smem_bytes <= kSmemBudget, "conv_igemm: internal smem accounting error (%d bytes)", smem_bytes);
  const int max_tiles = m_tiles * a.c_tiles;
  DYNMM_CHECK_ARG((long long)max_tiles * a.c_tiles < (1LL << 31) && max_tiles < (1 << 20), "conv_igemm: too many tiles");
  int grid = p->max_ctas > 0 ? p->max_ctas : sms;
  if (grid > max_tiles) grid = max_tiles;
  const int flags = (p->residual ? kFlagRes : 0) | (p->gated ? kFlagGated : 0) | (p->relu ? kFlagRelu : 0) |
                    (p->scale ? kFlagScale : 0);
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                           KernelArgs);
  static const KernelFn table[16] = {
      conv_igemm_kernel<0>,  conv_igemm_kernel<1>,  conv_igemm_kernel<2>,  conv_igemm_kernel<3>,
      conv_igemm_kernel<4>,  conv_igemm_kernel<5>,  conv_igemm_kernel<6>,  conv_igemm_kernel<7>,
      conv_igemm_kernel<8>,  conv_igemm_kernel<9>,  conv_igemm_kernel<10>, conv_igemm_kernel<11>,
      conv_igemm_kernel<12>, conv_igemm_kernel<13>, conv_igemm_kernel<14>, conv_igemm_kernel<15>};
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int i = 0; i < 16 && attr_err == cudaSuccess; ++i)
      attr_err = cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  });
  DYNMM_CUDA(attr_err);
  static const bool use_pdl = [] {
    const char* e = getenv("DYNMM_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  DYNMM_CUDA(cudaLaunchKernelEx(&cfg, table[flags], maps[0], maps[1], maps[2], maps[3], map_b, map_res, map_out, a));
  return DYNMM_OK;
}
