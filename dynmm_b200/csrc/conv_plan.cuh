// Shared pieces of the tcgen05 implicit-GEMM convolution: tiling structs, the per-tile epilogue math and the
// host-side planner (box selection, tensor maps, shared-memory budget).  Used by the one-launch-per-conv
// kernel (conv_igemm.cu) and by the multi-phase persistent program kernel (conv_program.cu).
#pragma once

#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tma_host.cuh"

namespace dynmm {
namespace convk {

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
  int two_per_sm;                   // 1: shared memory / TMEM sized so that two CTAs share an SM (C = 64 layers)
  int mt;                           // M tiles per work unit: 2 = two pixel tiles share every streamed weight tile
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
  int count_settled;                // `count` was written before the previous kernel of the stream started
  unsigned long long* trace;        // debug: 16 cycle stamps per CTA, or NULL
  int flags;                        // kFlag* (program kernel: epilogue variant chosen at run time)
  int cost;                         // relative cost of one tile (program kernel: CTA split between jobs)
  // tile-completion flags (dynmm_tile_flags in the header): layer-to-layer overlap without a kernel boundary
  dynmm_tile_flags in_f, res_f, out_f;
  int kh, kw, stride_h, stride_w, pad_h, pad_w, h_in, w_in;   // input window of an output tile (flag waits)
  int split;                        // DYNMM_CONV_SPLIT: [hi | lo] activation halves, 3-product contraction
  int act;                          // dynmm_conv_params.relu: 1 ReLU, 2 swish, 3 h-swish (kFlagRelu set for all three)
  int kc_c;                         // split: K chunks of one half (c_in / 64), 0 otherwise
  int in_lo_off;                    // split: first channel of the input's lo half (in_ld / 2)
  int out_lo_off;                   // split: first channel of the output's lo half (out_ld / 2)
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

// ---- tile-completion flags
constexpr unsigned kFlagSpinLimit = 1u << 24;     // watchdog: a broken dependency traps instead of hanging the GPU

__device__ __forceinline__ int ld_acquire_gpu_s32(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int32_t* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// flag index of the output tile with origin (n0, h0, w0)
__device__ __forceinline__ int flag_index(const dynmm_tile_flags& f, int n0, int h0, int w0) {
  return ((n0 / f.box_n) * f.tiles_h + h0 / f.box_h) * f.tiles_w + w0 / f.box_w;
}

// Whole warp: have all producer tiles overlapping samples [n_lo, n_hi) x rows [h_lo, h_hi) x columns [w_lo, w_hi)
// been published?  One flag per lane and pass.
__device__ __forceinline__ bool region_ready(const dynmm_tile_flags& f, int n_lo, int n_hi, int h_lo, int h_hi, int w_lo,
                                             int w_hi, int lane) {
  const int g0 = n_lo / f.box_n, th0 = h_lo / f.box_h, tw0 = w_lo / f.box_w;
  const int ng = (n_hi - 1) / f.box_n - g0 + 1, nh = (h_hi - 1) / f.box_h - th0 + 1, nw = (w_hi - 1) / f.box_w - tw0 + 1;
  const int total = ng * nh * nw;
  bool ok = true;
  for (int i = lane; i < total; i += 32) {
    const int tw = i % nw, r = i / nw;
    const int th = r % nh, g = r / nh;
    ok = ok && ld_acquire_gpu_s32(f.flags + ((g0 + g) * f.tiles_h + th0 + th) * f.tiles_w + tw0 + tw) >= f.need;
  }
  return __all_sync(0xffffffffu, ok);
}

// Whole warp: block until the input window (and the residual tile) of output tile `t` of this launch is complete.
__device__ __forceinline__ void wait_tile_inputs(const KernelArgs& a, const TileCoord& t, int active, int lane) {
  const int h0 = a.swap ? t.x1 : t.x2, w0 = a.swap ? t.x2 : t.x1;
  const int bh = a.swap ? a.b1 : a.b2, bw = a.swap ? a.b2 : a.b1;
  const int h1 = min(h0 + bh, a.h_out), w1 = min(w0 + bw, a.w_out);
  const int n1 = min(t.n0 + a.bn, active);
  if (h0 >= a.h_out || w0 >= a.w_out || t.n0 >= n1) return;
  const int ih0 = max(h0 * a.stride_h - a.pad_h, 0), ih1 = min((h1 - 1) * a.stride_h - a.pad_h + a.kh, a.h_in);
  const int iw0 = max(w0 * a.stride_w - a.pad_w, 0), iw1 = min((w1 - 1) * a.stride_w - a.pad_w + a.kw, a.w_in);
  unsigned spins = 0;
  while (!(region_ready(a.in_f, t.n0, n1, ih0, ih1, iw0, iw1, lane) &&
           (a.res_f.flags == nullptr || region_ready(a.res_f, t.n0, n1, h0, h1, w0, w1, lane)))) {
    if (++spins > kFlagSpinLimit) __trap();
    __nanosleep(100);
  }
  __syncwarp();
  fence_proxy_async_global();       // the TMA loads issued next (async proxy) must observe what the flags guard
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <bool kCg>
__device__ __forceinline__ uint4 ld_feat(const __nv_bfloat16* p) {
  return kCg ? __ldcg(reinterpret_cast<const uint4*>(p)) : __ldg(reinterpret_cast<const uint4*>(p));
}

// One thread's share of an epilogue sub-tile: 32 consecutive output channels of one pixel.
//   kTma: residual comes from the swizzled smem tile `res_smem`, result goes to the swizzled
//         staging tile `out_smem` (both 32-bit shared addresses of this thread's row);
//   else: direct global loads / stores (narrow channel tiles, partially active sample boxes).
//   kCg : residual / gated features were written earlier in the SAME kernel by other CTAs (program kernel):
//         read them through L2 (ld.global.cg), never through the non-coherent path.
// kFullCols (TMA epilogue of conv_igemm.cu): the sub-tile is a full 64-column one and the shift vector is staged (zero
// padded) for every column of it, so the per-group bounds tests fall away; columns >= c_out are clipped by the TMA store
// [kJ0, kJ1): which of the thread's 32 columns this call converts (the TMA epilogue converts the first 16 while the
// TMEM load of the other 16 is in flight)
template <bool kTma, bool kCg, bool kSplit = false, bool kAct = false, bool kFullCols = false, int kJ0 = 0, int kJ1 = 32>
__device__ __forceinline__ void epilogue_chunk(const int kFlags, const KernelArgs& args, const uint32_t (&v)[32], int c_first,
                                               int cols_left, bool valid, uint32_t res_smem, uint32_t out_smem,
                                               uint32_t chunk0, uint32_t swz, size_t pix, size_t rpix, size_t gpix,
                                               float g, const float* shift_smem, const uint4* pre_hi = nullptr,
                                               const uint4* pre_lo = nullptr) {
#pragma unroll
  for (int j = kJ0; j < kJ1; j += 8) {
    const int c = c_first + j;
    if (kFullCols || (j < cols_left && c < args.c_out)) {
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
        if (kTma && res_smem != 0) {
          r = lds128(res_smem + chunk);
        } else {
          r = (valid && (!kFullCols || c < args.c_out)) ? ld_feat<kCg>(args.residual + rpix * args.res_ld + c)
                                                        : make_uint4(0, 0, 0, 0);
        }
        if (kSplit) {
          // fp32-grade residual = hi + lo, reconstructed before it is added; both halves come from global memory --
          // prefetched into registers by the caller (pre_hi / pre_lo: this thread's four 8-channel groups) or loaded here
          uint4 q;
          if (pre_hi != nullptr) {
            r = pre_hi[j >> 3];
            q = pre_lo[j >> 3];
          } else {
            q = (valid && (!kFullCols || c < args.c_out))
                    ? ld_feat<kCg>(args.residual + rpix * args.res_ld + (args.res_ld >> 1) + c)
                    : make_uint4(0, 0, 0, 0);
          }
          f[0] += bf16_lo(r.x) + bf16_lo(q.x); f[1] += bf16_hi(r.x) + bf16_hi(q.x);
          f[2] += bf16_lo(r.y) + bf16_lo(q.y); f[3] += bf16_hi(r.y) + bf16_hi(q.y);
          f[4] += bf16_lo(r.z) + bf16_lo(q.z); f[5] += bf16_hi(r.z) + bf16_hi(q.z);
          f[6] += bf16_lo(r.w) + bf16_lo(q.w); f[7] += bf16_hi(r.w) + bf16_hi(q.w);
        } else {
          f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x); f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
          f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z); f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
        }
      }
      if (kFlags & kFlagRelu) {
        // kAct: the kernel variant that knows swish / h-swish.  A run-time test in the ReLU variants cost 6.5 % of the
        // bf16 step (measured, gpurun_out/r2ao), so they do not contain it.
        if (!kAct || args.act <= 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
        } else {
          // swish x * sigmoid(x) (2) / h-swish x * relu6(x + 3) / 6 (3), model_utils.py:100-115
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = args.act == 2 ? f[e] / (1.f + __expf(-f[e])) : f[e] * fminf(fmaxf(f[e] + 3.f, 0.f), 6.f) / 6.f;
        }
      }
      if (kFlags & kFlagGated) {
        if (g != 0.f && (!kFullCols || c < args.c_out)) {      // gated-off samples never touch the depth features
          const uint4 r = ld_feat<kCg>(args.gated + gpix * args.gated_ld + c);
          if (kSplit) {
            const uint4 q = ld_feat<kCg>(args.gated + gpix * args.gated_ld + (args.gated_ld >> 1) + c);
            f[0] += g * (bf16_lo(r.x) + bf16_lo(q.x)); f[1] += g * (bf16_hi(r.x) + bf16_hi(q.x));
            f[2] += g * (bf16_lo(r.y) + bf16_lo(q.y)); f[3] += g * (bf16_hi(r.y) + bf16_hi(q.y));
            f[4] += g * (bf16_lo(r.z) + bf16_lo(q.z)); f[5] += g * (bf16_hi(r.z) + bf16_hi(q.z));
            f[6] += g * (bf16_lo(r.w) + bf16_lo(q.w)); f[7] += g * (bf16_hi(r.w) + bf16_hi(q.w));
          } else {
            f[0] += g * bf16_lo(r.x); f[1] += g * bf16_hi(r.x); f[2] += g * bf16_lo(r.y); f[3] += g * bf16_hi(r.y);
            f[4] += g * bf16_lo(r.z); f[5] += g * bf16_hi(r.z); f[6] += g * bf16_lo(r.w); f[7] += g * bf16_hi(r.w);
          }
        }
      }
      uint4 o;
      o.x = pack_bf16(f[0], f[1]);
      o.y = pack_bf16(f[2], f[3]);
      o.z = pack_bf16(f[4], f[5]);
      o.w = pack_bf16(f[6], f[7]);
      uint4 l = make_uint4(0, 0, 0, 0);
      if (kSplit) {
        // lo half: what the bf16 rounding of the hi half dropped
        l.x = pack_bf16(f[0] - bf16_lo(o.x), f[1] - bf16_hi(o.x));
        l.y = pack_bf16(f[2] - bf16_lo(o.y), f[3] - bf16_hi(o.y));
        l.z = pack_bf16(f[4] - bf16_lo(o.z), f[5] - bf16_hi(o.z));
        l.w = pack_bf16(f[6] - bf16_lo(o.w), f[7] - bf16_hi(o.w));
      }
      if (kTma) {
        sts128(out_smem + chunk, o);
        if (kSplit) sts128(out_smem + kSubBytes + chunk, l);      // staging buffer = [hi sub-tile][lo sub-tile]
      } else if (valid) {
        *reinterpret_cast<uint4*>(args.out + pix * args.out_ld + c) = o;
        if (kSplit) {
          *reinterpret_cast<uint4*>(args.out + pix * args.out_ld + (args.out_ld >> 1) + c) = l;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ host side

// floor division for possibly negative tap offsets
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// generic mode: the (w,h,n) pixel box of a tile, maximising useful rows per 128-row MMA
inline void choose_box(int w, int h, int n, bool single_sample, int* bw, int* bh, int* bn) {
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
inline double box_eff(int w, int h, int n, int bw, int bh, int bn) {
  long long tiles = 1LL * ceil_div(w, bw) * ceil_div(h, bh) * ceil_div(n, bn);
  return (double)w * h * n / (double)(tiles * kBlockM);
}

// halo mode: b1 in {8,16,32} rows along d1 (multiple of 8 keeps tap views 1024-byte aligned), b2 = 128 / b1
inline void choose_halo_box(int d1, int d2, bool tapped, int* b1, int* b2, double* eff_out) {
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

// Everything a launch needs, derived from the C-ABI parameters on the host.
struct ConvPlan {
  CUtensorMap maps[4];      // A operand (one per stride parity in generic mode)
  CUtensorMap map_b, map_res, map_out;
  KernelArgs a;
  int smem_bytes;
  int max_tiles;            // work units (mt pixel tiles x one channel tile) when every sample slot is active
};

// `sms`: CTAs this convolution can expect to run on (the tile width is chosen to give each of them a tile);
// `smem_budget`: dynamic shared memory the kernel may use for this convolution
inline int plan_conv(const dynmm_conv_params* p, ConvPlan* plan, int sms, int smem_budget = kSmemBudget,
                     bool allow_two_per_sm = false, bool allow_dual = false, int partner_slots = 0) {
  // partner_slots > 0: this convolution shares its launch with a second one of identical geometry that has
  // `partner_slots` sample slots (dynmm_conv_igemm_fwd2): weights are streamed (two resident sets do not fit), two
  // shift vectors are staged, and the tile width / dual-unit choice looks at the tiles of BOTH jobs
  const bool merged = partner_slots > 0;
  if (merged) allow_two_per_sm = false;
  const bool split = (p->flags & DYNMM_CONV_SPLIT) != 0;
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
  DYNMM_CHECK_ARG(p->relu >= 0 && p->relu <= 3, "conv_igemm: relu must be 0 (none), 1 (ReLU), 2 (swish) or 3 (h-swish)");
  DYNMM_CHECK_ARG(!(split && p->relu > 1), "conv_igemm: DYNMM_CONV_SPLIT launches fuse ReLU only");
  if (p->relu > 1) allow_two_per_sm = false;          // the swish / h-swish kernel variants run one CTA per SM
  DYNMM_CHECK_ARG(!split || (p->c_in % kBlockK == 0 && p->in_ld % 16 == 0 && p->in_ld >= 2 * p->c_in && p->out_ld % 16 == 0 &&
                             p->out_ld >= 2 * p->c_out && p->res_ld % 16 == 0 && p->gated_ld % 16 == 0 &&
                             !p->in_flags.flags && !p->out_flags.flags),
                  "conv_igemm: DYNMM_CONV_SPLIT needs c_in %% 64 == 0 and [hi | lo] tensors (ld >= 2 * c, ld %% 16 == 0)");
  const int h_exp = (p->h_in + 2 * p->pad_h - p->kh) / p->stride_h + 1;
  const int w_exp = (p->w_in + 2 * p->pad_w - p->kw) / p->stride_w + 1;
  DYNMM_CHECK_ARG(h_exp == p->h_out && w_exp == p->w_out, "conv_igemm: output size %dx%d does not match %dx%d",
                  p->h_out, p->w_out, h_exp, w_exp);
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(p->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p->weight) & 15) == 0,
                  "conv_igemm: pointers must be 16-byte aligned");
  // EXPERIMENT (DYNMM_CONV_SMALL=1): every convolution planned into <= 113 KiB of shared memory and <= 256 TMEM columns,
  // so that two CTAs -- of consecutive launches (programmatic dependent launch then really overlaps the next
  // prologue with this tail) or of the RGB and the depth stream -- share an SM.
  static const bool small_mode = [] {
    const char* e = getenv("DYNMM_CONV_SMALL");
    return e && e[0] == '1';
  }();
  const bool small = small_mode && allow_two_per_sm && !p->trace;
  if (small) smem_budget = 113 * 1024;
  KernelArgs& a = plan->a;
  a = KernelArgs{};
  CUtensorMap* maps = plan->maps;
  CUtensorMap& map_b = plan->map_b;
  const uint64_t es = 2;
  const int c_out_pad = (p->c_out + 15) / 16 * 16;
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
      // channel tile of 64, 128 or 256: fewest MMA cycles on the critical CTA.  A 128 x N x 16 UMMA takes ~64
      // cycles for every N <= 128 and 128 cycles for N = 256 (tools/umma_issue_bench.cu), so the estimate is
      // rounds x (1 or 2); ties go to the narrower tile (more CTAs busy, shorter epilogues).
      const int c64 = (c_out_pad + 63) / 64 * 64;
      tile_n = 64;
      long long best = -1;
      for (int cand = 64; cand <= (small ? 128 : 256) && cand <= c64 && c64 != 192; cand *= 2) {
        const long long tiles = 1LL * m_tiles * (p->n + partner_slots) / p->n * ceil_div(c_out_pad, cand);
        const long long est = ceil_div_ll(tiles, sms) * (cand > 128 ? 2 : 1);
        // DYNMM_CONV_WIDE=1 (experiment, default off): ties go to the WIDER tile (one round of 256 instead of two of 128)
        static const bool prefer_wide = [] {
          const char* e = getenv("DYNMM_CONV_WIDE");
          return e && e[0] == '1';
        }();
        if (best < 0 || est < best || (prefer_wide && est == best)) {
          best = est;
          tile_n = cand;
        }
      }
    }
    DYNMM_CHECK_ARG(tile_n >= 16 && tile_n <= 256 && tile_n % 16 == 0, "conv_igemm: tile_n %d", tile_n);
    a.tile_n = tile_n;
    a.c_tiles = ceil_div(c_out_pad, tile_n);
    a.k_chunks = ceil_div(p->c_in, kBlockK) * (split ? 3 : 1);
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
    // small / split: residual read from global by the epilogue threads (split: both halves, no room for 2 x 16 KiB slots)
    a.aux_slots = (a.tma_epi && p->residual && !small && !split) ? kAuxSlots : 0;
    a.b_resident = (a.c_tiles == 1 && b_total <= kResidentBudget && !merged) ? 1 : 0;
    if (a.b_resident && a.aux_slots) a.aux_slots = 2;
    shift_bytes = (((p->c_out + 63) / 64 * 64 + 12) * 4) * (merged ? 2 : 1) + 16;     // staged per full 64-column sub-tile
    epi_bytes = (a.tma_epi ? (split ? 4 : 2) * kSubBytes : 0) + a.aux_slots * kSubBytes;
    a.stage_bytes = a.a_bytes + (a.b_resident ? 0 : a.tpg * b_tile_bytes);
    a.stages = (smem_budget - 2048 - epi_bytes - shift_bytes - (a.b_resident ? b_total : 0)) / a.stage_bytes;
    if (a.stages < 2 && a.b_resident) {     // not enough room next to the resident weights: stream them instead
      a.b_resident = 0;
      a.aux_slots = (a.tma_epi && p->residual && !small && !split) ? kAuxSlots : 0;
      epi_bytes = (a.tma_epi ? (split ? 4 : 2) * kSubBytes : 0) + a.aux_slots * kSubBytes;
      a.stage_bytes = a.a_bytes + a.tpg * b_tile_bytes;
      a.stages = (smem_budget - 2048 - epi_bytes - shift_bytes) / a.stage_bytes;
    }
    if (a.stages > kMaxStages) a.stages = kMaxStages;
    // Large C = 64 layers (24 KiB of resident weights, 18 KiB stages): a single CTA per SM leaves the tensor pipe, the
    // TMA engine and the epilogue warps idle in turn; with <= 113 KiB per CTA two CTAs share the SM and fill each
    // other's bubbles.  One residual slot and at most 4 stages per CTA.
    a.two_per_sm = 0;
    if (allow_two_per_sm && !split && halo && a.b_resident && tile_n == 64 && a.c_tiles == 1 && a.tma_epi && !p->trace &&
        m_tiles >= 4 * sms) {
      const int aux2 = a.aux_slots ? 1 : 0;
      const int epi2 = 2 * kSubBytes + aux2 * kSubBytes;
      int st2 = (113 * 1024 - 2048 - epi2 - shift_bytes - b_total) / a.stage_bytes;
      if (st2 > 4) st2 = 4;
      if (st2 >= 2) {
        a.two_per_sm = 1;
        a.aux_slots = aux2;
        epi_bytes = epi2;
        a.stages = st2;
      }
    }
    if (a.stages >= 2 || !halo) break;
  }
  DYNMM_CHECK_ARG(a.stages >= 2, "conv_igemm: not enough shared memory for 2 stages");
  // Dual-M work units (streamed weights, C >= 256): the stage-3/4 layers are bound by operand streaming from L2
  // (a 128 x 128 tile with K = 768 pulls 272 KB, 72 % of it weights, for 3072 cycles of UMMAs -- DESIGN.md 5b(f)).
  // Two pixel tiles of the same channel tile share every weight tile: a stage holds A0, A1 and B, the MMA warp issues
  // both tiles' UMMAs against the same B descriptors into two accumulators.  Half the weight traffic per output, half
  // the CTAs per layer (the RGB and the depth launch of a stage then fit on the chip side by side).
  a.mt = 1;
  static const int dual_mode = [] {       // DYNMM_CONV_DUAL=0 off, 1 (default) heuristic, 2 whenever possible
    const char* e = getenv("DYNMM_CONV_DUAL");
    return e ? atoi(e) : 1;
  }();
  const bool force_dual = (p->flags & DYNMM_CONV_FORCE_DUAL) || dual_mode > 1;
  if (allow_dual && !small && !split && dual_mode > 0 && !(p->flags & DYNMM_CONV_NO_DUAL) && !a.b_resident && a.tma_epi && tile_n <= 128 &&
      m_tiles >= 2 && !a.two_per_sm &&
      (force_dual || 4LL * m_tiles * (p->n + partner_slots) / p->n * a.c_tiles > 3LL * sms)) {
    const int stage2 = 2 * a.a_bytes + a.tpg * b_tile_bytes;
    for (int aux2 = a.aux_slots > 2 ? 2 : a.aux_slots; aux2 >= 0; --aux2) {
      // a residual without a free aux slot is read by the epilogue threads straight from global memory
      const int epi2 = 2 * kSubBytes + aux2 * kSubBytes;
      int st2 = (smem_budget - 1024 - 256 - epi2 - shift_bytes) / stage2;
      if (st2 > kMaxStages) st2 = kMaxStages;
      if (st2 >= 2) {
        a.mt = 2;
        a.aux_slots = aux2;
        epi_bytes = epi2;
        a.stage_bytes = stage2;
        a.stages = st2;
        break;
      }
    }
  }
  a.acc_stride = (tile_n + 31) / 32 * 32;
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.mt * a.acc_stride) a.tmem_cols *= 2;
  if (small && a.tmem_cols <= 256) a.two_per_sm = 1;
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
  a.count_settled = (p->flags & DYNMM_CONV_COUNT_SETTLED) ? 1 : 0;
  a.trace = static_cast<unsigned long long*>(p->trace);
  a.kh = p->kh; a.kw = p->kw; a.stride_h = p->stride_h; a.stride_w = p->stride_w; a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.h_in = p->h_in; a.w_in = p->w_in;
  a.split = split ? 1 : 0;
  a.act = p->relu;
  a.kc_c = split ? p->c_in / kBlockK : 0;
  a.in_lo_off = split ? p->in_ld / 2 : 0;
  a.out_lo_off = split ? p->out_ld / 2 : 0;
  a.in_f = p->in_flags;
  a.res_f = p->res_flags;
  if (a.in_f.flags == nullptr) a.res_f.flags = nullptr;        // ordinary stream order covers the residual too
  DYNMM_CHECK_ARG(!merged || a.in_f.flags == nullptr, "conv_igemm: tile flags are not supported in merged launches");
  DYNMM_CHECK_ARG(a.in_f.flags == nullptr || (p->in_map == nullptr && p->res_map == nullptr && p->gated == nullptr),
                  "conv_igemm: in_flags cannot be combined with in_map / res_map / gated");
  DYNMM_CHECK_ARG(a.in_f.flags == nullptr || p->residual == nullptr || a.res_f.flags != nullptr ||
                      (p->flags & DYNMM_CONV_RESIDUAL_SETTLED),
                  "conv_igemm: in_flags with a residual needs res_flags (or DYNMM_CONV_RESIDUAL_SETTLED)");
  DYNMM_CHECK_ARG((a.in_f.flags == nullptr || (a.in_f.box_n >= 1 && a.in_f.box_h >= 1 && a.in_f.box_w >= 1 && a.in_f.need >= 1)) &&
                      (a.res_f.flags == nullptr || (a.res_f.box_n >= 1 && a.res_f.box_h >= 1 && a.res_f.box_w >= 1 && a.res_f.need >= 1)),
                  "conv_igemm: malformed tile flags");
  // flags this launch publishes: one per pixel tile, complete at c_tiles
  a.out_f.flags = p->out_flags.flags;
  a.out_f.box_n = a.bn;
  a.out_f.box_h = a.swap ? a.b1 : a.b2;
  a.out_f.box_w = a.swap ? a.b2 : a.b1;
  a.out_f.tiles_h = a.swap ? a.tiles1 : a.tiles2;
  a.out_f.tiles_w = a.swap ? a.tiles2 : a.tiles1;
  a.out_f.need = a.c_tiles;
  DYNMM_CHECK_ARG(a.out_f.flags == nullptr ||
                      (p->out_flags.box_n == a.out_f.box_n && p->out_flags.box_h == a.out_f.box_h &&
                       p->out_flags.box_w == a.out_f.box_w && p->out_flags.tiles_h == a.out_f.tiles_h &&
                       p->out_flags.tiles_w == a.out_f.tiles_w && p->out_flags.need == a.out_f.need),
                  "conv_igemm: out_flags geometry does not match dynmm_conv_tile_grid() for this launch");
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
  bool used[4] = {false, false, false, false};
  if (halo) {
    used[0] = true;
    int rc = pixel_map(&maps[0], p->in, split ? p->in_ld : p->c_in, p->w_in, p->h_in, p->n_in, (uint64_t)p->in_ld * es,
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
      int rc = pixel_map(&maps[m], base, split ? p->in_ld : p->c_in, sub_w, sub_h, p->n_in, (uint64_t)p->in_ld * p->stride_w * es,
                         (uint64_t)p->in_ld * p->w_in * p->stride_h * es,
                         (uint64_t)p->in_ld * p->w_in * p->h_in * es, a.b2);
      if (rc) return rc;
    }
  }
  int first_used = 0;
  while (!used[first_used]) ++first_used;
  for (int m = 0; m < 4; ++m)
    if (!used[m]) maps[m] = maps[first_used];
  {
    const uint64_t c_in_w = (uint64_t)p->c_in * (split ? 3 : 1);          // split: [W_hi | W_lo | W_hi] along K
    const uint64_t dims[3] = {c_in_w, (uint64_t)c_out_pad, (uint64_t)num_taps};
    const uint64_t strides[2] = {c_in_w * es, c_in_w * c_out_pad * es};
    const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)tile_n, (uint32_t)a.tpg};
    int rc = encode_map(&map_b, p->weight, 3, dims, strides, box);
    if (rc) return rc;
  }
  // epilogue maps: residual (load) and output (store), one [box pixels][64 channels] sub-tile per transfer
  CUtensorMap& map_res = plan->map_res;
  CUtensorMap& map_out = plan->map_out;
  map_res = map_b;
  map_out = map_b;
  if (a.tma_epi) {
    int rc = pixel_map(&map_out, p->out, p->c_out, p->w_out, p->h_out, p->n, (uint64_t)p->out_ld * es,
                       (uint64_t)p->out_ld * p->w_out * es, (uint64_t)p->out_ld * p->w_out * p->h_out * es, a.b2);
    if (rc) return rc;
    if (split) {
      // no residual ring in split mode: the second epilogue map is the store map of the output's lo half
      rc = pixel_map(&map_res, static_cast<const __nv_bfloat16*>(p->out) + p->out_ld / 2, p->c_out, p->w_out, p->h_out, p->n,
                     (uint64_t)p->out_ld * es, (uint64_t)p->out_ld * p->w_out * es,
                     (uint64_t)p->out_ld * p->w_out * p->h_out * es, a.b2);
      if (rc) return rc;
    }
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
  DYNMM_CHECK_ARG(smem_bytes <= smem_budget, "conv_igemm: internal smem accounting error (%d bytes)", smem_bytes);
  const int max_tiles = ceil_div(m_tiles, a.mt) * a.c_tiles;       // work units when every sample slot is active
  DYNMM_CHECK_ARG((long long)m_tiles * a.c_tiles * a.c_tiles < (1LL << 31) && m_tiles * a.c_tiles < (1 << 20),
                  "conv_igemm: too many tiles");
  plan->smem_bytes = smem_bytes;
  plan->max_tiles = max_tiles;
  a.flags = (p->residual ? kFlagRes : 0) | (p->gated ? kFlagGated : 0) | (p->relu ? kFlagRelu : 0) |
            (p->scale ? kFlagScale : 0);
  a.cost = (num_taps * a.k_chunks + 4) * tile_n;
  return DYNMM_OK;
}

}  // namespace convk
}  // namespace dynmm
