// Convolution weight gradient for sm_100a: a tcgen05 GEMM whose K dimension is the pixel index.
//
//   dw[tap][co][ci] = sum_p dy[p][co] * x[p + tap][ci]          (NHWC bf16 in, fp32 out)
//
// Both operands are consumed exactly as they lie in memory: a TMA box of [pixels][64 channels]
// lands in shared memory as rows of 128 B with the 128-byte swizzle, which is the canonical
// MN-major operand layout of the UMMA (channels = M or N contiguous, pixels = K, 8-pixel groups
// 1024 B apart), so no transpose is ever materialised.  The tap shift is a coordinate offset of the
// TMA load (strided convolutions read a parity sub-lattice through a strided tensor map), and
// out-of-image pixels are zero-filled by TMA -- padding contributes nothing to the sum.
//
// Work decomposition: one CTA owns a (128 output channels) x (<=128 input channels) x (<=3 taps)
// block of dw and a contiguous slice of the pixel tiles (split-K).  Accumulators live in TMEM
// (taps x input-channel columns, <= 384 of 512); at the end the CTA writes its fp32 partial block
// to the workspace and a second kernel sums the slices in a fixed order and writes the
// [c_out][c_in][kh][kw] gradient: deterministic, no atomics.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue.
// Shared-memory stages are zeroed once: rows a TMA box never writes (boxes shorter than 64 pixels,
// the second 64-channel atom of a 64-channel layer) therefore stay zero and add nothing.
#include <stdlib.h>

#include "common.cuh"
#include "tma_host.cuh"

namespace dynmm {

namespace {

constexpr int kPix = 64;                       // pixels (K) per pipeline stage
constexpr int kAtomBytes = kPix * 128;         // one [64 pixels][64 channels] bf16 tile: 8 KiB
constexpr int kMaxStagesW = 6;
constexpr int kThreadsW = 192;
constexpr int kSmemBudgetW = 227 * 1024;
constexpr int kMaxTaps = 9;

struct TapW {
  int8_t map, o1, o2, pad;
};

struct WgradArgs {
  int bw, bh, bn;                  // pixel box of a K tile (bw*bh*bn <= 64)
  int tiles_w, tiles_h, tiles_n;   // K tiles per direction
  int k_tiles, k_per_split, splits;
  int rows, k_steps;               // rows per box, UMMA k-steps (16 pixels) per tile
  int co_tiles, ci_tiles, tap_groups, tpu;   // units = co_tiles * ci_tiles * tap_groups; tpu taps per unit
  int n_atoms, n_tile;             // input-channel atoms (64 ch) per unit, n_tile = 64 * n_atoms
  int stages, stage_bytes, tmem_cols;
  int c_out, c_in, taps;
  TapW tap[kMaxTaps];
  float* partial;                  // [splits][taps][c_out][c_in]
};

struct __align__(8) SmemCtlW {
  uint64_t full[kMaxStagesW];
  uint64_t empty[kMaxStagesW];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// MN-major, 128-byte-swizzled operand: 64-channel atoms `lbo` bytes apart, 8-pixel groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16, bf16 x bf16 -> fp32, A and B both MN-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__global__ void __launch_bounds__(kThreadsW, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x0,
                  const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                  const __grid_constant__ CUtensorMap map_x3, const __grid_constant__ WgradArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SmemCtlW* ctl = reinterpret_cast<SmemCtlW*>(smem + args.stages * args.stage_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // unit / split of this CTA
  const int unit = blockIdx.x / args.splits;
  const int split = blockIdx.x - unit * args.splits;
  const int tg = unit % args.tap_groups;
  const int ci_t = (unit / args.tap_groups) % args.ci_tiles;
  const int co_t = unit / (args.tap_groups * args.ci_tiles);
  const int co0 = co_t * 128, ci0 = ci_t * args.n_tile;
  const int m_atoms = min(2, (args.c_out - co0 + 63) >> 6);
  const int n_atoms = min(args.n_atoms, (args.c_in - ci0 + 63) >> 6);
  const int k0 = split * args.k_per_split;
  const int k1 = min(k0 + args.k_per_split, args.k_tiles);

  // zero the operand stages (see header comment), then make the zeros visible to the async proxy
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = args.stages * args.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kThreadsW) z[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x0);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    mbar_init(&ctl->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, args.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const int dy_bytes = 2 * kAtomBytes;           // the A view always spans two atoms (M = 128)

  // Producer and MMA issuer: the whole warp runs the loop converged, the TMA / tcgen05 instructions sit under
  // elect.sync (a lone lane inside `if (lane == 0)` pays an ELECT / BRA.U.ANY loop per instruction: >= 90 cycles per
  // UMMA, tools/umma_issue_bench.cu).
  if (warp == 0) {
    {
      const CUtensorMap* maps[4] = {&map_x0, &map_x1, &map_x2, &map_x3};
      const uint32_t box_bytes = args.rows * 128;
      const uint32_t tx = (m_atoms + args.tpu * n_atoms) * box_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k0; kt < k1; ++kt) {
        const int tw = kt % args.tiles_w;
        const int r = kt / args.tiles_w;
        const int th = r % args.tiles_h;
        const int tn = r / args.tiles_h;
        const int x1 = tw * args.bw, x2 = th * args.bh, n0 = tn * args.bn;
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * args.stage_bytes;
        if (elect_one()) {
          mbar_expect_tx(&ctl->full[stage], tx);
          for (int a = 0; a < m_atoms; ++a) tma_load_4d(sa + a * kAtomBytes, &map_dy, &ctl->full[stage], co0 + a * 64, x1, x2, n0);
          for (int t = 0; t < args.tpu; ++t) {
            const TapW tp = args.tap[tg * args.tpu + t];
            for (int b = 0; b < n_atoms; ++b)
              tma_load_4d(sa + dy_bytes + (t * args.n_atoms + b) * kAtomBytes, maps[tp.map], &ctl->full[stage],
                          ci0 + b * 64, x1 + tp.o1, x2 + tp.o2, n0);
          }
        }
        __syncwarp();
        if (++stage == args.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = umma_idesc_bf16_mn(128, args.n_tile);
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k0; kt < k1; ++kt) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * args.stage_bytes);
        if (elect_one()) {
          for (int t = 0; t < args.tpu; ++t) {
            const uint32_t sb = sa + dy_bytes + t * args.n_atoms * kAtomBytes;
            for (int ks = 0; ks < args.k_steps; ++ks) {
              // 16 pixels = two 8-pixel groups = 2048 bytes further along K
              const uint64_t da = umma_desc_sw128_mn(sa + ks * 2048, kAtomBytes);
              const uint64_t db = umma_desc_sw128_mn(sb + ks * 2048, kAtomBytes);
              umma_bf16(tmem_base + t * args.n_tile, da, db, idesc, (kt > k0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(&ctl->empty[stage]);
        }
        __syncwarp();
        if (++stage == args.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&ctl->acc_full);
      __syncwarp();
    }
  } else {
    // epilogue: TMEM lane = output channel, column = (tap, input channel)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int co = co0 + row;
    mbar_wait(&ctl->acc_full, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    for (int t = 0; t < args.tpu; ++t) {
      const int tap = tg * args.tpu + t;
      float* dst = args.partial + ((static_cast<size_t>(split) * args.taps + tap) * args.c_out + co) * args.c_in + ci0;
      for (int cb = 0; cb < args.n_tile; cb += 32) {
        uint32_t v[32];
        tmem_ld32(t_row + t * args.n_tile + cb, v);
        tmem_ld_wait();
        if (co < args.c_out) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (ci0 + cb + j < args.c_in) {
              *reinterpret_cast<float4*>(dst + cb + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, args.tmem_cols);
  }
}

// dw[co][ci][tap] (=|+=) sum_s partial[s][tap][co][ci]: one thread per (co, ci) -- reads are coalesced over ci
// for every (split, tap), the `taps` results of a thread are adjacent in dw
template <int kTaps>
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int c_out, int c_in,
                                    float* __restrict__ dw, int accumulate) {
  const size_t plane = static_cast<size_t>(c_out) * c_in;
  const size_t per = plane * taps;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float s[kTaps];
#pragma unroll
    for (int t = 0; t < kTaps; ++t) s[t] = 0.f;
    // four splits' loads in flight per step (the adds stay in split order: same rounding as the plain loop)
    int k = 0;
    for (; k + 4 <= splits; k += 4) {
      float v[4][kTaps];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < kTaps; ++t) v[u][t] = (t < taps) ? __ldg(partial + (k + u) * per + t * plane + i) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < kTaps; ++t) s[t] += v[u][t];
    }
    for (; k < splits; ++k) {
#pragma unroll
      for (int t = 0; t < kTaps; ++t) {
        if (t < taps) s[t] += __ldg(partial + k * per + t * plane + i);
      }
    }
    float* o = dw + i * taps;
#pragma unroll
    for (int t = 0; t < kTaps; ++t) {
      if (t < taps) o[t] = accumulate ? o[t] + s[t] : s[t];
    }
  }
}

// comparator: one block per (tap, co), threads over ci; fp32 accumulation in pixel order
__global__ void wgrad_direct_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    float* __restrict__ dw, dynmm_wgrad_params p) {
  const int tap = blockIdx.x / p.c_out, co = blockIdx.x % p.c_out;
  const int ky = tap / p.kw, kx = tap % p.kw;
  for (int ci = threadIdx.x; ci < p.c_in; ci += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < p.n; ++n) {
      for (int ho = 0; ho < p.h_out; ++ho) {
        const int hi = ho * p.stride_h + ky - p.pad_h;
        if (hi < 0 || hi >= p.h_in) continue;
        for (int wo = 0; wo < p.w_out; ++wo) {
          const int wi = wo * p.stride_w + kx - p.pad_w;
          if (wi < 0 || wi >= p.w_in) continue;
          const float g = __bfloat162float(dy[((static_cast<size_t>(n) * p.h_out + ho) * p.w_out + wo) * p.dy_ld + co]);
          const float v = __bfloat162float(x[((static_cast<size_t>(n) * p.h_in + hi) * p.w_in + wi) * p.x_ld + ci]);
          s = fmaf(g, v, s);
        }
      }
    }
    const size_t o = (static_cast<size_t>(co) * p.c_in + ci) * (p.kh * p.kw) + tap;
    dw[o] = p.accumulate ? dw[o] + s : s;
  }
}

inline int floordiv_w(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// the (w,h,n) pixel box of a K tile: as many useful pixels per 64-row stage as possible
void choose_kbox(int w, int h, int n, int* bw, int* bh, int* bn) {
  double best = -1;
  for (int cw = 1; cw <= w && cw <= kPix; ++cw) {
    for (int ch = 1; ch <= h && cw * ch <= kPix; ++ch) {
      int cn = kPix / (cw * ch);
      if (cn > n) cn = n;
      if (cn < 1) cn = 1;
      if (cn > 1 && (cw != w || ch != h)) cn = 1;   // several samples per box only for whole maps
      const long long tiles = 1LL * ceil_div(w, cw) * ceil_div(h, ch) * ceil_div(n, cn);
      const double eff = (double)w * h * n / (double)(tiles * kPix);
      const double score = eff + 1e-6 * cw;
      if (score > best) {
        best = score;
        *bw = cw;
        *bh = ch;
        *bn = cn;
      }
    }
  }
}

int check_wgrad(const dynmm_wgrad_params* p) {
  DYNMM_CHECK_ARG(p && p->x && p->dy && p->dw, "conv_wgrad: null pointer");
  DYNMM_CHECK_ARG(p->kh >= 1 && p->kw >= 1 && p->kh * p->kw <= kMaxTaps, "conv_wgrad: at most %d taps", kMaxTaps);
  DYNMM_CHECK_ARG(p->stride_h >= 1 && p->stride_h <= 2 && p->stride_w >= 1 && p->stride_w <= 2,
                  "conv_wgrad: stride must be 1 or 2");
  DYNMM_CHECK_ARG(p->c_in % 8 == 0 && p->x_ld % 8 == 0 && p->x_ld >= p->c_in, "conv_wgrad: c_in/x_ld %% 8");
  DYNMM_CHECK_ARG(p->c_out % 8 == 0 && p->dy_ld % 8 == 0 && p->dy_ld >= p->c_out, "conv_wgrad: c_out/dy_ld %% 8");
  DYNMM_CHECK_ARG(p->n >= 1, "conv_wgrad: empty batch");
  const int h_exp = (p->h_in + 2 * p->pad_h - p->kh) / p->stride_h + 1;
  const int w_exp = (p->w_in + 2 * p->pad_w - p->kw) / p->stride_w + 1;
  DYNMM_CHECK_ARG(h_exp == p->h_out && w_exp == p->w_out, "conv_wgrad: output size %dx%d does not match %dx%d",
                  p->h_out, p->w_out, h_exp, w_exp);
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->dy) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p->dw) & 15) == 0,
                  "conv_wgrad: pointers must be 16-byte aligned");
  return DYNMM_OK;
}

// everything that depends on shapes only
int plan_wgrad(const dynmm_wgrad_params* p, WgradArgs* a) {
  const int taps = p->kh * p->kw;
  a->taps = taps;
  a->c_out = p->c_out;
  a->c_in = p->c_in;
  choose_kbox(p->w_out, p->h_out, p->n, &a->bw, &a->bh, &a->bn);
  a->tiles_w = ceil_div(p->w_out, a->bw);
  a->tiles_h = ceil_div(p->h_out, a->bh);
  a->tiles_n = ceil_div(p->n, a->bn);
  a->k_tiles = a->tiles_w * a->tiles_h * a->tiles_n;
  a->rows = a->bw * a->bh * a->bn;
  a->k_steps = ceil_div(a->rows, 16);
  a->tpu = (taps % 3 == 0) ? 3 : 1;
  a->tap_groups = taps / a->tpu;
  const int atoms_in = ceil_div(p->c_in, 64);
  a->n_atoms = atoms_in >= 2 ? 2 : 1;
  a->n_tile = 64 * a->n_atoms;
  a->ci_tiles = ceil_div(atoms_in, a->n_atoms);
  a->co_tiles = ceil_div(p->c_out, 128);
  a->stage_bytes = (2 + a->tpu * a->n_atoms) * kAtomBytes;
  a->stages = (kSmemBudgetW - 2048) / a->stage_bytes;
  if (a->stages > kMaxStagesW) a->stages = kMaxStagesW;
  a->tmem_cols = 32;
  while (a->tmem_cols < a->tpu * a->n_tile) a->tmem_cols *= 2;
  const int units = a->co_tiles * a->ci_tiles * a->tap_groups;
  int ctas = p->max_ctas > 0 ? p->max_ctas : num_sms();
  int splits = ctas / units;
  if (splits < 1) splits = 1;
  if (splits > a->k_tiles) splits = a->k_tiles;
  a->k_per_split = ceil_div(a->k_tiles, splits);
  a->splits = ceil_div(a->k_tiles, a->k_per_split);
  return units;
}

}  // namespace

}  // namespace dynmm

using namespace dynmm;

extern "C" long long dynmm_conv_wgrad_workspace(const dynmm_wgrad_params* p) {
  if (check_wgrad(p) != DYNMM_OK) return -1;
  WgradArgs a{};
  plan_wgrad(p, &a);
  return 4LL * a.splits * a.taps * p->c_out * p->c_in;
}

extern "C" int dynmm_conv_wgrad(const dynmm_wgrad_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_wgrad(p);
  if (rc) return rc;
  WgradArgs a{};
  const int units = plan_wgrad(p, &a);
  const long long need = 4LL * a.splits * a.taps * p->c_out * p->c_in;
  DYNMM_CHECK_ARG(p->workspace && p->workspace_bytes >= need, "conv_wgrad: workspace of %lld bytes needed", need);
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0, "conv_wgrad: workspace must be 16-byte aligned");
  DYNMM_CHECK_ARG(a.stages >= 2, "conv_wgrad: not enough shared memory for 2 stages");
  a.partial = static_cast<float*>(p->workspace);
  const uint64_t es = 2;

  auto pixel_map = [&](CUtensorMap* m, const void* base, int c, int w, int h, int n, uint64_t st_w, uint64_t st_h,
                       uint64_t st_n) -> int {
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {st_w, st_h, st_n};
    const uint32_t box[4] = {64u, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bn};
    return encode_map(m, base, 4, dims, strides, box);
  };
  CUtensorMap map_dy;
  rc = pixel_map(&map_dy, p->dy, p->c_out, p->w_out, p->h_out, p->n, (uint64_t)p->dy_ld * es,
                 (uint64_t)p->dy_ld * p->w_out * es, (uint64_t)p->dy_ld * p->w_out * p->h_out * es);
  if (rc) return rc;
  CUtensorMap maps[4];
  bool used[4] = {false, false, false, false};
  for (int ky = 0; ky < p->kh; ++ky) {
    for (int kx = 0; kx < p->kw; ++kx) {
      const int dy = ky - p->pad_h, dx = kx - p->pad_w;
      const int qy = floordiv_w(dy, p->stride_h), py = dy - qy * p->stride_h;
      const int qx = floordiv_w(dx, p->stride_w), px = dx - qx * p->stride_w;
      TapW& t = a.tap[ky * p->kw + kx];
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
    DYNMM_CHECK_ARG(sub_w >= 1 && sub_h >= 1, "conv_wgrad: input too small for stride");
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(p->x) + (static_cast<size_t>(py) * p->w_in + px) * p->x_ld;
    rc = pixel_map(&maps[m], base, p->c_in, sub_w, sub_h, p->n, (uint64_t)p->x_ld * p->stride_w * es,
                   (uint64_t)p->x_ld * p->w_in * p->stride_h * es, (uint64_t)p->x_ld * p->w_in * p->h_in * es);
    if (rc) return rc;
  }
  int first_used = 0;
  while (!used[first_used]) ++first_used;
  for (int m = 0; m < 4; ++m)
    if (!used[m]) maps[m] = maps[first_used];

  const int smem_bytes = a.stages * a.stage_bytes + 1024 + (int)sizeof(SmemCtlW);
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run(
      [] { return cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudgetW); }));
  conv_wgrad_kernel<<<units * a.splits, kThreadsW, smem_bytes, stream>>>(map_dy, maps[0], maps[1], maps[2], maps[3], a);
  DYNMM_LAUNCH_CHECK();
  const long long plane = 1LL * p->c_out * p->c_in;
  int blocks = static_cast<int>((plane + 127) / 128);
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  if (a.taps <= 3) {
    wgrad_reduce_kernel<3><<<blocks, 128, 0, stream>>>(a.partial, a.splits, a.taps, p->c_out, p->c_in, p->dw,
                                                        p->accumulate);
  } else {
    wgrad_reduce_kernel<9><<<blocks, 128, 0, stream>>>(a.partial, a.splits, a.taps, p->c_out, p->c_in, p->dw,
                                                        p->accumulate);
  }
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}

extern "C" int dynmm_conv_wgrad_direct(const dynmm_wgrad_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_wgrad(p);
  if (rc) return rc;
  wgrad_direct_kernel<<<p->kh * p->kw * p->c_out, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(p->x),
                                                                   static_cast<const __nv_bfloat16*>(p->dy), p->dw, *p);
  DYNMM_LAUNCH_CHECK();
  return DYNMM_OK;
}
