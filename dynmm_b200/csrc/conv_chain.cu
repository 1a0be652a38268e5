// Chains of 3-tap convolutions (NonBottleneck1D blocks, resnet.py:124-147) as ONE kernel for sm_100a.
//
// At batch 8 a layer of the deep encoder stages / the decoder modules is a few microseconds of tensor work; launched
// layer by layer, prologue + first-load latency + epilogue drain cost twice that (profiles/r1b_plan_report.txt).  Here
// a CTA owns a STRIP of `rows` image rows of one sample and runs every layer of the chain on it:
//
//   * activations live in shared memory from the first layer to the last, in a pixel-major layout made of 16-byte
//     planes:  byte(slot, ch) = (ch / 8) * plane_bytes + slot * 16 + (ch % 8) * 2,  slot = (row + 1) * (W + 1) + col.
//     This is the canonical K-major NO-SWIZZLE UMMA operand (core matrix = 8 consecutive slots x 16 B = 128
//     contiguous bytes, SBO = 128, LBO = plane_bytes), so the A operand of a tap is the same buffer with the start
//     address moved by one slot (1x3) or one padded row (3x1): no im2col, no halo copies, any strip shape.  Column W
//     of every row is a permanent zero slot (it is both the right padding of its row and the left padding of the
//     next), the rows above / below the image are written as zeros by the input load and never touched again;
//     the input load itself is plain coalesced ld.global (lanes along the channels of a pixel) + st.shared, conflict
//     free because plane_bytes = 16 (mod 128) -- a TMA box with a 16-byte inner extent was 3x slower;
//   * the M tile is 128 consecutive slots (a strip is `tiles` of them), N = all c output channels in one UMMA
//     (128 x c x 16), accumulators in TMEM (tiles * c <= 512 columns);
//   * weights stream through a TMA pipeline ([c][64] SWIZZLE_128B tiles, one per (k chunk, tap)), prefetched across
//     layer boundaries;
//   * the epilogue (TMEM -> +shift, +residual, ReLU -> bf16) writes the next layer's operand IN PLACE (every MMA of
//     the layer has retired), plus global copies where the caller wants them (block outputs: the residual of the
//     next block, the result);
//   * a 3x1 layer needs the last row of the strip above and the first row of the strip below: the producer layer
//     writes its edge rows to a global scratch buffer and every epilogue warp adds 1 to the strip's counter (release,
//     gpu scope); the neighbours poll it (acquire) until it reaches 8 or 16 per exchange, copy the rows into their halo
//     slots and arrive on an mbarrier only the MMA issuer waits on.  Strips of other samples never wait for each other;
//     the last strip of a sample to finish clears the counters (the launch is idempotent, CUDA-graph replayable).
//
// Accumulation order per output element is (k chunk, tap, k16) and the epilogue arithmetic is that of
// conv_igemm.cu: results are bit-identical to the per-layer launches (tests/test_gpu_chain.py).
#include "conv_plan.cuh"

namespace dynmm {

namespace {

using namespace convk;

// warp 0: weight producer, warp 1: MMA issuer, warps 2.. : epilogue (kEW = 8 or 16 of them: TMEM -> registers is
// 64 B / clock per SM and a warp cannot convert while it waits for its load, so more warps keep that pipe busy)
constexpr int kChainMaxStages = 8;
constexpr int kChainMaxLayers = 64;
constexpr unsigned kChainSpinLimit = 1u << 22;

// one layer of the device-resident image (dynmm_conv_chain_build)
struct __align__(128) ChainLayerImg {
  CUtensorMap wmap;          // bf16 [3][c][c], box {64, c, 1}, SWIZZLE_128B
  const float* shift;
  int32_t taps_h, relu, residual, store;
  int32_t pad_[26];
};
static_assert(sizeof(ChainLayerImg) == 256, "image layout");

struct ChainJobArgs {
  const ChainLayerImg* layers;
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  __nv_bfloat16* out_last;
  const int32_t* count;
  int n, n_layers, count_settled, pad_;
};

struct ChainArgs {
  ChainJobArgs job[2];
  int units0;                       // units of job 0
  int h, w, c;
  int w1, rows, strips, tiles;      // slot pitch of a row, rows per strip, strips per sample, M tiles per strip
  int s_buf, plane_bytes, planes;
  int stages, stage_bytes;
  uint32_t m_strips, m_w1, m_w;
  int desc_swap;                    // experiment: LBO / SBO exchanged
  int trace_layer;                  // debug: layers trace_layer and trace_layer + 1 are stamped
  int no_exchange;                  // experiment (wrong results): edge rows are not exchanged
  int32_t* flags;                   // [units] exchange counters, then [samples] finished-strip counters
  uint4* scratch;                   // [units][2 parities][2 sides][planes][w] 16-byte chunks
  unsigned long long* trace;
};

struct __align__(8) ChainCtl {
  uint64_t full[kChainMaxStages];
  uint64_t empty[kChainMaxStages];
  uint64_t acc_full;
  uint64_t a_ready;
  uint64_t in_full;
  uint64_t halo_full;
  uint32_t tmem_base;
};

struct ChainGeom {
  int w1, rows, strips, tiles, s_buf, plane_bytes, planes, stages, stage_bytes, smem_bytes;
};

#define CHAIN_TRACE(slot)                                                                  \
  do {                                                                                     \
    if (args.trace) args.trace[blockIdx.x * 16 + (slot)] = (unsigned long long)clock64();  \
  } while (0)

// K-major operand WITHOUT swizzle: 8-row core matrices of 128 contiguous bytes; `lbo` = byte distance between the
// two core matrices of a 16-element K step, `sbo` = byte distance between consecutive 8-row groups
__device__ __forceinline__ uint64_t umma_desc_plain(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ void st_release_gpu(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// mbarrier wait with a watchdog: a broken dependency traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void chain_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

template <int kC, int kEW>
__global__ void __launch_bounds__(64 + 32 * kEW, 1)
conv_chain_kernel(const __grid_constant__ ChainArgs args) {
  constexpr int kChainEpi = 32 * kEW;           // epilogue threads
  constexpr int kColGroups = kEW / 4;           // the c columns are split between the warps of a TMEM lane quarter
  constexpr int kPlanes = kC / 8;
  constexpr int kChunks = kC / (32 * kColGroups);   // 32-column chunks per epilogue thread and tile
  constexpr int kKChunks = kC / 64;             // 64-channel K chunks
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;                                              // [stages][kC rows][128 B]
  uint8_t* smem_a = smem_b + args.stages * args.stage_bytes;           // [planes][s_buf][16 B]
  float* smem_shift = reinterpret_cast<float*>(smem_a + kPlanes * args.plane_bytes);   // [2][kC]
  ChainCtl* ctl = reinterpret_cast<ChainCtl*>(smem_shift + 2 * kC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x;
  const bool jb = u >= args.units0;
  const int uj = u - (jb ? args.units0 : 0);
  const int slot = fast_div(uj, args.m_strips);
  const int strip = uj - slot * args.strips;
  const ChainLayerImg* layers = jb ? args.job[1].layers : args.job[0].layers;
  const int n_layers = jb ? args.job[1].n_layers : args.job[0].n_layers;
  const int32_t* count = jb ? args.job[1].count : args.job[0].count;
  const int settled = jb ? args.job[1].count_settled : args.job[0].count_settled;
  const int n_slots = jb ? args.job[1].n : args.job[0].n;
  if (threadIdx.x == 0) CHAIN_TRACE(0);

  bool waited = false;
  if (count != nullptr && !settled) {
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    waited = true;
  }
  const int active = count ? min(*reinterpret_cast<const volatile int32_t*>(count), n_slots) : n_slots;
  if (slot >= active) return;                    // the gate switched this sample's depth stage off (whole CTA)

  const int r0 = strip * args.rows;
  const int rows_valid = min(args.rows, args.h - r0);
  const bool has_up = strip > 0;
  const bool has_down = r0 + args.rows < args.h;
  const bool exchange_on = args.strips > 1 && !args.no_exchange;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    mbar_init(&ctl->acc_full, 1);
    mbar_init(&ctl->a_ready, kEW);
    mbar_init(&ctl->in_full, kEW);
    mbar_init(&ctl->halo_full, kEW);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (ctl->tmem_base != 0) __trap();             // one CTA per SM owns all of TMEM: the base is column 0 / lane 0
  if (threadIdx.x == 0) CHAIN_TRACE(1);

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer: constants, runs ahead of the
    // previous kernel's completion and across layer boundaries
    int stage = 0;
    uint32_t phase = 0;
    for (int l = 0; l < n_layers; ++l) {
      const CUtensorMap* wm = &layers[l].wmap;
      for (int kc = 0; kc < kKChunks; ++kc) {
        for (int tap = 0; tap < 3; ++tap) {
          chain_wait(&ctl->empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&ctl->full[stage], kC * 128);
            tma_load_3d(smem_b + stage * args.stage_bytes, wm, &ctl->full[stage], kc * kBlockK, 0, tap);
          }
          __syncwarp();
          if (++stage == args.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ input load + MMA issuer
    if (!waited) asm volatile("griddepcontrol.wait;\n" ::: "memory");
    // (triggering the dependent launch only at the end of the chain was measured: 1.5 % slower -- the next kernel's
    // launch latency no longer hides behind this one)
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    const uint32_t idesc = umma_idesc_bf16(kBlockM, kC);
    const uint32_t a_base = smem_u32(smem_a);
    const uint32_t lbo = args.desc_swap ? 128u : static_cast<uint32_t>(args.plane_bytes);
    const uint32_t sbo = args.desc_swap ? static_cast<uint32_t>(args.plane_bytes) : 128u;
    int stage = 0;
    uint32_t phase = 0;
    int halos = 0;                                 // halo exchanges consumed so far
    int taps_h = layers[0].taps_h;
    for (int l = 0; l < n_layers; ++l) {
      const int taps_next = (l + 1 < n_layers) ? layers[l + 1].taps_h : 0;     // in flight during this layer
      if (l == 0) {
        chain_wait(&ctl->in_full, 0);
      } else {
        chain_wait(&ctl->a_ready, (l - 1) & 1);
        if (taps_h && exchange_on && (has_up || has_down)) {
          chain_wait(&ctl->halo_full, halos & 1);   // the neighbours' edge rows have landed in the halo slots
          ++halos;
        }
      }
      tc_fence_after();
      if (lane == 0 && (l == args.trace_layer || l == args.trace_layer + 1)) CHAIN_TRACE(2 + 7 * (l - args.trace_layer));
      const int tap_step = (taps_h ? args.w1 : 1) * 16;
      for (int kc = 0; kc < kKChunks; ++kc) {
        for (int tap = 0; tap < 3; ++tap) {
          chain_wait(&ctl->full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * args.stage_bytes));
            for (int t = 0; t < args.tiles; ++t) {
              const uint32_t a0 = a_base + (kc * 8) * args.plane_bytes + (args.w1 + t * kBlockM) * 16 + (tap - 1) * tap_step;
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                const uint64_t da = umma_desc_plain(a0 + k * 2 * args.plane_bytes, lbo, sbo);
                umma_bf16(t * kC, da, db + (k * 2), idesc, (kc | tap | k) != 0);
              }
            }
            umma_commit(&ctl->empty[stage]);
            if (kc == kKChunks - 1 && tap == 2) umma_commit(&ctl->acc_full);
          }
          __syncwarp();
          if (++stage == args.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (lane == 0 && (l == args.trace_layer || l == args.trace_layer + 1)) CHAIN_TRACE(3 + 7 * (l - args.trace_layer));
      taps_h = taps_next;
    }
  } else {
    // ------------------------------------------------------------ epilogue (8 warps)
    if (!waited) asm volatile("griddepcontrol.wait;\n" ::: "memory");    // residual rows / scratch of earlier launches
    const int te = threadIdx.x - 64;              // 0..kChainEpi-1
    const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31
    const int half = (warp - 2) >> 2;             // which column group
    const int row = quarter * 32 + lane;
    const bool leader = te == 0;
    const __nv_bfloat16* in_g = jb ? args.job[1].in : args.job[0].in;
    __nv_bfloat16* out_g = jb ? args.job[1].out : args.job[0].out;
    __nv_bfloat16* last_g = jb ? args.job[1].out_last : args.job[0].out_last;
    const uint32_t a_base = smem_u32(smem_a);
    const size_t pix_base = (static_cast<size_t>(slot) * args.h + r0) * args.w;
    const size_t side_chunks = static_cast<size_t>(kPlanes) * args.w;   // 16-byte chunks of one edge row
    const int items = args.tiles * kChunks;        // (tile, 32-column chunk) pairs of this thread
    {
      // chain input -> 16-byte planes.  Lanes run along the channel groups of one pixel (512 / 256 contiguous bytes of
      // NHWC memory); plane_bytes = 16 (mod 128) spreads their 32 / 16 shared-memory stores over all banks.  Column w
      // of every row and the rows outside the image become zeros (the padding of every later layer).
      const int n_slot = (args.rows + 2) * args.w1;
      const int total = n_slot * kPlanes;
      constexpr int kBatch = kEW == 8 ? 8 : 4;
      for (int i0 = te; i0 < total; i0 += kBatch * kChainEpi) {
        uint4 q[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int i = i0 + b * kChainEpi;
          q[b] = make_uint4(0, 0, 0, 0);
          if (i < total) {
            const int sl = i / kPlanes, pl = i - sl * kPlanes;
            const int rr = fast_div(sl, args.m_w1);
            const int wc = sl - rr * args.w1;
            const int gr = r0 - 1 + rr;
            if (gr >= 0 && gr < args.h && wc < args.w)
              q[b] = __ldg(reinterpret_cast<const uint4*>(in_g + ((static_cast<size_t>(slot) * args.h + gr) * args.w + wc) * kC + pl * 8));
          }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int i = i0 + b * kChainEpi;
          if (i < total) {
            const int sl = i / kPlanes, pl = i - sl * kPlanes;
            sts128(a_base + pl * args.plane_bytes + sl * 16, q[b]);
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->in_full);
    }
    int exch = 0;                                  // exchanges published so far
    for (int l = 0; l < n_layers; ++l) {
      const ChainLayerImg& L = layers[l];
      const float* shift_g = L.shift;
      const int relu = L.relu, residual = L.residual, store = L.store;
      const bool publish = (l + 1 < n_layers) && layers[l + 1].taps_h != 0 && exchange_on;
      const float sh = (te < kC && shift_g != nullptr) ? __ldg(shift_g + te) : 0.f;
      const __nv_bfloat16* res_g = residual == 1 ? in_g : out_g;
      __nv_bfloat16* st_g = store == 1 ? out_g : last_g;
      float* shift_s = smem_shift + (l & 1) * kC;
      if (te < kC) shift_s[te] = sh;
      const uint32_t shift_a = smem_u32(shift_s);
      named_barrier(1, kChainEpi);
      chain_wait(&ctl->acc_full, l & 1);
      tc_fence_after();
      const bool tr = leader && (l == args.trace_layer || l == args.trace_layer + 1);
      const int ts = 7 * (l - args.trace_layer);
      if (tr) CHAIN_TRACE(4 + ts);
      uint4* scr = args.scratch + (static_cast<size_t>(u) * 2 + (exch & 1)) * 2 * side_chunks;

      // item -> (tile, chunk); the TMEM load (and the residual rows) of item i + 1 are in flight while item i is
      // converted: two register sets
      auto item_pix = [&](int it, int& p, int& r, int& wc) {
        const int t = it / kChunks;
        p = t * kBlockM + row;
        r = fast_div(p, args.m_w1);
        wc = p - r * args.w1;
      };
      auto issue = [&](int it, uint32_t (&v)[32]) {
        const int t = it / kChunks, ch = it - t * kChunks;
        const int c0 = half * (kC / kColGroups) + ch * 32;
        tmem_ld32((static_cast<uint32_t>(quarter * 32) << 16) + t * kC + c0, v);
      };
      auto process = [&](int it, const uint32_t (&v)[32]) {
        const int t = it / kChunks, ch = it - t * kChunks;
        const int c0 = half * (kC / kColGroups) + ch * 32;
        int p, r, wc;
        item_pix(it, p, r, wc);
        if (!(r < rows_valid && wc < args.w)) return;
        const size_t pix = pix_base + static_cast<size_t>(r) * args.w + wc;
        const uint32_t a_row = a_base + (args.w1 + p) * 16;
        const bool edge_up = publish && r == 0 && has_up;
        const bool edge_dn = publish && r == rows_valid - 1 && has_down;
        uint4 rr[4];
        if (residual) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = __ldcg(reinterpret_cast<const uint4*>(res_g + pix * kC + c0 + j * 8));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
          const uint4 b0 = lds128(shift_a + (c0 + j * 8) * 4);
          const uint4 b1 = lds128(shift_a + (c0 + j * 8 + 4) * 4);
          f[0] += __uint_as_float(b0.x); f[1] += __uint_as_float(b0.y); f[2] += __uint_as_float(b0.z); f[3] += __uint_as_float(b0.w);
          f[4] += __uint_as_float(b1.x); f[5] += __uint_as_float(b1.y); f[6] += __uint_as_float(b1.z); f[7] += __uint_as_float(b1.w);
          if (residual) {
            const uint4 q = rr[j];
            f[0] += bf16_lo(q.x); f[1] += bf16_hi(q.x); f[2] += bf16_lo(q.y); f[3] += bf16_hi(q.y);
            f[4] += bf16_lo(q.z); f[5] += bf16_hi(q.z); f[6] += bf16_lo(q.w); f[7] += bf16_hi(q.w);
          }
          if (relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          uint4 o;
          o.x = pack_bf16(f[0], f[1]);
          o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]);
          o.w = pack_bf16(f[6], f[7]);
          const int plane = (c0 >> 3) + j;
          sts128(a_row + plane * args.plane_bytes, o);
          if (store) *reinterpret_cast<uint4*>(st_g + pix * kC + c0 + j * 8) = o;
          if (edge_up) scr[static_cast<size_t>(plane) * args.w + wc] = o;
          if (edge_dn) scr[side_chunks + static_cast<size_t>(plane) * args.w + wc] = o;
        }
      };
      {
        if constexpr (kEW == 8) {
          // two TMEM loads in flight per wait (tcgen05.wait::ld covers every outstanding load of the thread, so a
          // load issued ahead of the conversion of the previous one would be waited for anyway)
          uint32_t va[32], vb[32];
#pragma unroll 1
          for (int it = 0; it < items; it += 2) {
            issue(it, va);
            if (it + 1 < items) issue(it + 1, vb);
            tmem_ld_wait();
            process(it, va);
            if (it + 1 < items) process(it + 1, vb);
          }
        } else {
          // 16 warps (96 registers each): one load at a time, the other warps of the scheduler cover its latency
          uint32_t va[32];
#pragma unroll 1
          for (int it = 0; it < items; ++it) {
            issue(it, va);
            tmem_ld_wait();
            process(it, va);
          }
        }
      }
      if (tr) CHAIN_TRACE(5 + ts);
      // this warp's share of the next layer's operand is written (generic proxy -> visible to the tensor core) and its
      // accumulator columns are read
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->a_ready);
      if (publish) {
        // edge rows -> neighbours, warp by warp (no CTA-wide barrier on the way): a warp's scratch stores are ordered
        // before its +1 on the strip's counter (release, gpu scope); a neighbour's rows are complete when its counter
        // reaches 8 per exchange.  Every warp polls for itself, then copies its share of the two rows into the halo
        // slots and arrives on `halo_ready`, which only the MMA issuer waits for.
        ++exch;
        if (lane == 0) red_release_gpu_add(args.flags + u, 1);   // release: orders the warp's stores (bar.warp) before it
        if (tr) CHAIN_TRACE(6 + ts);
        if (has_up || has_down) {
          const bool need = (lane == 0 && has_up) || (lane == 1 && has_down);
          const int32_t* f = args.flags + (lane == 0 ? u - 1 : u + 1);
          unsigned spins = 0;
          while (need && ld_acquire_gpu_s32(f) < exch * kEW) {
            if (++spins > kChainSpinLimit) __trap();
            __nanosleep(20);
          }
          __syncwarp();
          if (tr) CHAIN_TRACE(7 + ts);
          const int par = (exch - 1) & 1;
          const uint4* up = args.scratch + ((static_cast<size_t>(u - 1) * 2 + par) * 2 + 1) * side_chunks;
          const uint4* dn = args.scratch + ((static_cast<size_t>(u + 1) * 2 + par) * 2 + 0) * side_chunks;
          const uint32_t top = a_base;                                          // halo row above the strip
          const uint32_t bot = a_base + (args.w1 + rows_valid * args.w1) * 16;  // halo row below it
          // 2 * side_chunks 16-byte chunks over 256 threads: all loads of a thread in flight before its stores
          constexpr int kBatch = kEW == 8 ? 4 : 3;   // 16-byte loads of a thread in flight before its stores
          for (int i0 = te; i0 < 2 * (int)side_chunks; i0 += kBatch * kChainEpi) {
            uint4 q[kBatch];
            uint32_t dst[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
              const int i = i0 + b * kChainEpi;
              dst[b] = 0;
              if (i < 2 * (int)side_chunks) {
                const bool lower = i >= (int)side_chunks;
                const int k = lower ? i - (int)side_chunks : i;
                const int plane = fast_div(k, args.m_w);
                const int wc = k - plane * args.w;
                if (lower ? has_down : has_up) {
                  q[b] = __ldcg((lower ? dn : up) + k);
                  dst[b] = (lower ? bot : top) + plane * args.plane_bytes + wc * 16;
                }
              }
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b)
              if (dst[b]) sts128(dst[b], q[b]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->halo_full);
        }
      }
      if (tr) CHAIN_TRACE(8 + ts);
    }
    // leave the flags zero for the next launch: the last strip of a sample to finish clears them
    if (args.strips > 1) {
      __threadfence();
      named_barrier(1, kChainEpi);
      if (leader) {
        const int sample = fast_div(u, args.m_strips);
        int32_t* done = args.flags + gridDim.x + sample;
        if (atomicAdd(done, 1) == args.strips - 1) {
          __threadfence();
          for (int s = 0; s < args.strips; ++s) args.flags[sample * args.strips + s] = 0;
          *done = 0;
        }
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

// strip geometry for an h x w map with c channels and `total_slots` sample slots over all jobs
int plan_chain(int h, int w, int c, int total_slots, int sms, ChainGeom* g) {
  if (c != 128 && c != 256) {
    set_error("conv_chain: c must be 128 or 256 (got %d)", c);
    return DYNMM_EUNSUPPORTED;
  }
  if (h < 1 || w < 1 || w + 1 > 256 || total_slots < 1) {
    set_error("conv_chain: unsupported map %dx%d", h, w);
    return DYNMM_EUNSUPPORTED;
  }
  const int w1 = w + 1;
  const int tmax = 512 / c;
  bool found = false;
  for (int t = 1; t <= tmax; ++t) {
    int rows = t * kBlockM / w1;
    if (rows > h) rows = h;
    if (rows > 254) rows = 254;
    if (rows < 1) continue;
    ChainGeom cand;
    cand.w1 = w1;
    cand.rows = rows;
    cand.strips = ceil_div(h, rows);
    cand.tiles = ceil_div(rows * w1, kBlockM);
    cand.s_buf = (cand.tiles * kBlockM + 2 * w1 + 7) / 8 * 8 + 1;     // plane_bytes = 16 (mod 128): see the input load
    cand.plane_bytes = cand.s_buf * 16;
    cand.planes = c / 8;
    cand.stage_bytes = c * 128;
    const int fixed = 1024 + cand.planes * cand.plane_bytes + 2 * c * 4 + (int)sizeof(ChainCtl) + 64;
    cand.stages = (kSmemBudget - fixed) / cand.stage_bytes;
    if (cand.stages > kChainMaxStages) cand.stages = kChainMaxStages;
    if (cand.stages < 2) continue;
    cand.smem_bytes = fixed + cand.stages * cand.stage_bytes;
    if (!found || (long long)total_slots * g->strips > sms) {
      // keep the first geometry whose units fit the GPU at once; until then, the one with the fewest units
      if (!found || cand.strips < g->strips) *g = cand;
      found = true;
    }
    if ((long long)total_slots * g->strips <= sms) break;
  }
  if (!found) {
    set_error("conv_chain: %dx%dx%d does not fit shared memory", h, w, c);
    return DYNMM_EUNSUPPORTED;
  }
  // whole rows per strip: a row pitch that fills the 128-slot M tiles badly (e.g. w + 1 = 71: one row per tile) wastes
  // more tensor work than the chain saves -- leave such maps to the per-layer kernel and its 2-D boxes
  if (10 * g->rows * w1 < 7 * g->tiles * kBlockM && g->strips > 1) {
    set_error("conv_chain: %d-slot rows fill %d-slot tiles to less than 70 %%", w1, g->tiles * kBlockM);
    return DYNMM_EUNSUPPORTED;
  }
  // strips of one sample wait for each other: all of a sample's CTAs must be resident together.  CTAs are dispatched
  // in order, so a launch larger than the GPU still completes, but it runs in dependent waves -- refuse beyond 2x.
  if ((long long)total_slots * g->strips > 2LL * sms) {
    set_error("conv_chain: %lld units exceed twice the SM count", (long long)total_slots * g->strips);
    return DYNMM_EUNSUPPORTED;
  }
  return DYNMM_OK;
}

}  // namespace

}  // namespace dynmm

using namespace dynmm;

extern "C" long long dynmm_conv_chain_image_bytes(int n_layers) {
  if (n_layers < 1 || n_layers > kChainMaxLayers) return -1;
  return static_cast<long long>(n_layers) * sizeof(ChainLayerImg);
}

extern "C" int dynmm_conv_chain_build(const dynmm_chain_layer* layers, int n_layers, int c, void* host_image) {
  DYNMM_CHECK_ARG(layers && host_image, "conv_chain_build: null pointer");
  DYNMM_CHECK_ARG(n_layers >= 1 && n_layers <= kChainMaxLayers, "conv_chain_build: 1..%d layers", kChainMaxLayers);
  DYNMM_CHECK_ARG(c == 128 || c == 256, "conv_chain_build: c must be 128 or 256");
  ChainLayerImg* img = static_cast<ChainLayerImg*>(host_image);
  memset(img, 0, sizeof(ChainLayerImg) * n_layers);
  for (int l = 0; l < n_layers; ++l) {
    const dynmm_chain_layer& s = layers[l];
    DYNMM_CHECK_ARG(s.weight && (reinterpret_cast<uintptr_t>(s.weight) & 15) == 0, "conv_chain_build: layer %d weight", l);
    DYNMM_CHECK_ARG(s.residual >= 0 && s.residual <= 2 && s.store >= 0 && s.store <= 2, "conv_chain_build: layer %d flags", l);
    const uint64_t dims[3] = {(uint64_t)c, (uint64_t)c, 3};
    const uint64_t strides[2] = {(uint64_t)c * 2, (uint64_t)c * c * 2};
    const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)c, 1};
    int rc = encode_map(&img[l].wmap, s.weight, 3, dims, strides, box);
    if (rc) return rc;
    img[l].shift = s.shift;
    img[l].taps_h = s.taps_h ? 1 : 0;
    img[l].relu = s.relu ? 1 : 0;
    img[l].residual = s.residual;
    img[l].store = s.store;
  }
  return DYNMM_OK;
}

extern "C" int dynmm_conv_chain_plan(int h, int w, int c, int total_slots, int32_t* units, long long* scratch_bytes) {
  ChainGeom g;
  int rc = plan_chain(h, w, c, total_slots, num_sms(), &g);
  if (rc) return rc;
  const long long u = (long long)total_slots * g.strips;
  if (units) *units = (int32_t)u;
  if (scratch_bytes) *scratch_bytes = u * 2 * 2 * g.planes * w * 16;
  return DYNMM_OK;
}

extern "C" int dynmm_conv_chain_fwd(const dynmm_chain_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(p, "conv_chain_fwd: null pointer");
  DYNMM_CHECK_ARG(p->n_jobs == 1 || p->n_jobs == 2, "conv_chain_fwd: 1 or 2 jobs");
  int total = 0;
  for (int j = 0; j < p->n_jobs; ++j) {
    const dynmm_chain_job& jb = p->jobs[j];
    DYNMM_CHECK_ARG(jb.image && jb.in && jb.n >= 1 && jb.n_layers >= 1 && jb.n_layers <= kChainMaxLayers,
                    "conv_chain_fwd: job %d incomplete", j);
    DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(jb.image) & 127) == 0, "conv_chain_fwd: image must be 128-byte aligned");
    DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(jb.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(jb.out) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(jb.out_last) & 15) == 0,
                    "conv_chain_fwd: tensors must be 16-byte aligned");
    total += jb.n;
  }
  ChainGeom g;
  int rc = plan_chain(p->h, p->w, p->c, total, num_sms(), &g);
  if (rc) return rc;
  const long long units = (long long)total * g.strips;
  DYNMM_CHECK_ARG(g.strips == 1 || (p->flags && p->scratch), "conv_chain_fwd: flags / scratch missing");
  DYNMM_CHECK_ARG(g.strips == 1 || p->scratch_bytes >= units * 2 * 2 * g.planes * p->w * 16,
                  "conv_chain_fwd: scratch too small");
  ChainArgs a{};
  for (int j = 0; j < p->n_jobs; ++j) {
    const dynmm_chain_job& jb = p->jobs[j];
    a.job[j].layers = static_cast<const ChainLayerImg*>(jb.image);
    a.job[j].in = static_cast<const __nv_bfloat16*>(jb.in);
    a.job[j].out = static_cast<__nv_bfloat16*>(jb.out);
    a.job[j].out_last = static_cast<__nv_bfloat16*>(jb.out_last);
    a.job[j].count = jb.count;
    a.job[j].n = jb.n;
    a.job[j].n_layers = jb.n_layers;
    a.job[j].count_settled = jb.count_settled;
  }
  a.units0 = p->jobs[0].n * g.strips;
  a.h = p->h;
  a.w = p->w;
  a.c = p->c;
  a.w1 = g.w1;
  a.rows = g.rows;
  a.strips = g.strips;
  a.tiles = g.tiles;
  a.s_buf = g.s_buf;
  a.plane_bytes = g.plane_bytes;
  a.planes = g.planes;
  a.stages = g.stages;
  a.stage_bytes = g.stage_bytes;
  auto magic = [](int d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)(((1ULL << 32) + d - 1) / d); };
  a.m_strips = magic(g.strips);
  a.m_w1 = magic(g.w1);
  a.m_w = magic(p->w);
  static const int desc_swap = [] {
    const char* e = getenv("DYNMM_CHAIN_DESC_SWAP");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  a.desc_swap = desc_swap;
  {
    const char* e = getenv("DYNMM_CHAIN_TRACE_LAYER");
    a.trace_layer = e ? atoi(e) : 1;
    const char* x = getenv("DYNMM_CHAIN_NOEXCH");
    a.no_exchange = (x && x[0] == '1') ? 1 : 0;
  }
  a.flags = p->flags;
  a.scratch = static_cast<uint4*>(p->scratch);
  a.trace = static_cast<unsigned long long*>(p->trace);

  typedef void (*KernelFn)(ChainArgs);
  static const KernelFn table[4] = {conv_chain_kernel<128, 8>, conv_chain_kernel<256, 8>, conv_chain_kernel<128, 16>,
                                    conv_chain_kernel<256, 16>};
  static const int epi_warps = [] {            // DYNMM_CHAIN_EPI=8: the 8-warp epilogue (experiments)
    const char* e = getenv("DYNMM_CHAIN_EPI");
    return (e && atoi(e) == 8) ? 8 : 16;
  }();
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run([] {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    return e;
  }));
  static const bool use_pdl = [] {
    const char* e = getenv("DYNMM_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)units);
  cfg.blockDim = dim3(64 + 32 * epi_warps);
  cfg.dynamicSmemBytes = g.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  DYNMM_CUDA(cudaLaunchKernelEx(&cfg, table[(p->c == 256 ? 1 : 0) + (epi_warps == 16 ? 2 : 0)], a));
  return DYNMM_OK;
}
