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
#include "conv_plan.cuh"

namespace dynmm {

namespace {

using namespace convk;

#define DYNMM_TRACE(slot)                                                              \
  do {                                                                                 \
    if (args.trace) args.trace[blockIdx.x * 16 + (slot)] = (unsigned long long)clock64(); \
  } while (0)

// Experiments only (-DDYNMM_TRACE_TILES=1, tools/conv_trace.py with TILES=1): per-tile stamps of the first 16 tiles of a
// CTA in a trace buffer of 64 slots per CTA -- [16 + l] MMAs of tile l issued, [32 + l] accumulator l seen full by the
// epilogue, [48 + l] epilogue of tile l done.
#ifndef DYNMM_TRACE_TILES
#define DYNMM_TRACE_TILES 0
#endif
#if DYNMM_TRACE_TILES
#define DYNMM_TRACE_T(base, l)                                                                          \
  do {                                                                                                  \
    if (args.trace && (l) < 16) args.trace[blockIdx.x * 64 + (base) + (l)] = (unsigned long long)clock64(); \
  } while (0)
#undef DYNMM_TRACE
#define DYNMM_TRACE(slot)                                                              \
  do {                                                                                 \
    if (args.trace) args.trace[blockIdx.x * 64 + (slot)] = (unsigned long long)clock64(); \
  } while (0)
#else
#define DYNMM_TRACE_T(base, l) do { } while (0)
#endif

// (The per-job code below is written as lambdas that take the job's KernelArgs / tensor maps BY REFERENCE and are
// called once per job with the kernel parameters themselves: after inlining every field access is a direct
// constant-bank operand, as in a single-job kernel.  Selecting the job through a run-time pointer instead costs a
// local-memory pointer load plus a generic load per field: +13 % kernel time, measured in round 2.)

// Two convolutions of IDENTICAL geometry (the same layer of the RGB and of the depth encoder) may share a launch:
// job 1 has its own tensor maps and KernelArgs (pointers, sample count), the tiling fields of both KernelArgs are
// equal (checked on the host).  The CTAs walk one combined unit list -- job 0's units, then job 1's -- so at batch 8 a
// CTA typically runs one RGB and one depth tile back to back: the second tile's loads and UMMAs hide the first one's
// epilogue, and the fixed cost of a launch (prologue, first-load latency, drain, launch gap) is paid once per LAYER
// instead of once per layer and encoder.  args2.n == 0: single convolution.
// kMode: 0 = one CTA per SM, 1 = two CTAs per SM (large C = 64 layers), 2 = split operands (DYNMM_CONV_SPLIT; one per SM),
// 3 = mode 0 with the swish / h-swish activations compiled in (dynmm_conv_params.relu = 2 / 3)
template <int kFlags, int kMode>
__global__ void __launch_bounds__(kThreads, kMode == 1 ? 2 : 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_res,
                  const __grid_constant__ CUtensorMap map_out, const __grid_constant__ KernelArgs args,
                  const __grid_constant__ CUtensorMap map2_a0, const __grid_constant__ CUtensorMap map2_a1,
                  const __grid_constant__ CUtensorMap map2_a2, const __grid_constant__ CUtensorMap map2_a3,
                  const __grid_constant__ CUtensorMap map2_b, const __grid_constant__ CUtensorMap map2_res,
                  const __grid_constant__ CUtensorMap map2_out, const __grid_constant__ KernelArgs args2) {
  constexpr int kPerSm = kMode == 1 ? 2 : 1;
  constexpr bool kSplit = kMode == 2;
  constexpr bool kAct = kMode == 3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int k_iters = args.num_groups * args.k_chunks;             // A loads per tile
  const int b_tile_bytes = args.tile_n * kBlockK * 2;              // one tap's [tile_n][64] weight tile
  const int b_iter_bytes = args.tpg * b_tile_bytes;                // weights consumed per A load
  uint8_t* smem_bres = smem + args.stages * args.stage_bytes;      // resident weights [k_iters][tpg][tile_n][128 B]
  uint8_t* smem_aux = smem_bres + (args.b_resident ? k_iters * b_iter_bytes : 0);   // [aux_slots][16 KiB]
  uint8_t* smem_stage_out = smem_aux + args.aux_slots * kSubBytes;                  // [2][16 KiB] (tma_epi only)
  // split: a staging buffer holds the hi and the lo sub-tile
  constexpr int stage_out_bytes = kSplit ? 2 * kSubBytes : kSubBytes;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_stage_out + (args.tma_epi ? 2 * stage_out_bytes : 0));
  // [jobs][c_out + 8]: per-channel shift (zeros if absent), 16-byte aligned for float4 broadcast loads
  float* smem_shift = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ctl + 1) + 15) & ~uintptr_t(15));
  const int c_pad64 = (args.c_out + 63) & ~63;            // shift staged (zero padded) for whole 64-column sub-tiles
  const int shift_stride = c_pad64 + 8;
  const bool two_jobs = kPerSm == 1 && args2.n > 0;      // two CTAs per SM: single-job launches only (registers)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) DYNMM_TRACE(0);
  const int n_sub = (args.tile_n + 63) >> 6;
  const bool aux_on = (kFlags & kFlagRes) && args.tma_epi && args.aux_slots > 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_b);
    if (args.tma_epi) tma_prefetch_desc(&map_out);
    if (aux_on || kSplit) tma_prefetch_desc(&map_res);
    if (two_jobs) {
      tma_prefetch_desc(&map2_a0);
      tma_prefetch_desc(&map2_b);
      if (args.tma_epi) tma_prefetch_desc(&map2_out);
      if (aux_on) tma_prefetch_desc(&map2_res);
    }
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
  }
  if (warp == 0) {
    __syncwarp();
    if (args.b_resident && elect_one()) {
      // weights are constants: fetch them before waiting on the previous kernel (c_tiles == 1; single job only)
      mbar_expect_tx(&ctl->b_full, k_iters * b_iter_bytes);
      for (int g = 0; g < args.num_groups; ++g)
        for (int kc = 0; kc < args.k_chunks; ++kc)
          tma_load_3d(smem_bres + (g * args.k_chunks + kc) * b_iter_bytes, &map_b, &ctl->b_full, kc * kBlockK, 0,
                      g * args.tpg);
    }
    __syncwarp();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, kPerSm == 1 ? 512 : args.tmem_cols);
    tmem_relinquish();
  }
  // `count` of a depth-encoder launch: when the gate plan that wrote it is known to be complete (every launch but the
  // first after dynmm_gate_plan), the load is issued here and overlaps the programmatic-dependent-launch wait
  int active_early0 = args.n, active_early1 = args2.n;
  if (args.count && args.count_settled) active_early0 = min(__ldg(args.count), args.n);
  if (two_jobs && args2.count && args2.count_settled) active_early1 = min(__ldg(args2.count), args2.n);
  tc_fence_before();
  __syncthreads();                 // barriers initialised, TMEM allocated -- nothing slower than that in front of it
  tc_fence_after();
  if (warp >= 2) {
    // epilogue warps stage the shift vectors once (no global loads inside the tile loop); the global-load latency
    // overlaps the first TMA loads and UMMAs instead of delaying them: only the epilogue warps wait for it
    for (int c = threadIdx.x - 64; c < c_pad64; c += 32 * kEpiWarps) {
      smem_shift[c] = (args.shift && c < args.c_out) ? args.shift[c] : 0.f;
      if (two_jobs) smem_shift[shift_stride + c] = (args2.shift && c < args.c_out) ? args2.shift[c] : 0.f;
    }
    named_barrier(3, 32 * kEpiWarps);
  }
  // One CTA per SM owns all 512 TMEM columns, so the allocation starts at column 0 / lane 0: using the CONSTANT
  // keeps every tcgen05.mma operand in uniform registers.  Two CTAs per SM allocate what they need (2 x 64 columns)
  // and carry the base in a register.
  if (kPerSm == 1 && ctl->tmem_base != 0) __trap();
  const uint32_t tmem_base = kPerSm == 1 ? 0u : ctl->tmem_base;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch, resident weights --
  // constants only) overlapped the tail of the previous kernel in the stream.  From here on we touch tensors it
  // produced, so wait for it; then let OUR dependent start its prologue.
  // With tile-completion flags on the input (in_f) there is NO wait for the previous kernel as a whole: the loader
  // waits per work unit for the producer tiles its window overlaps, so this CTA -- scheduled on an SM one of the
  // previous layer's early finishers freed -- already works while that layer's last tiles are in flight.
  if (args.in_f.flags == nullptr) asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  // A work unit = `mt` consecutive pixel tiles x one channel tile (mt = 2: both tiles share every streamed weight
  // tile).  unit -> channel tile ct = unit % c_tiles, pixel tiles m = (unit / c_tiles) * mt + w.
  const int mt = args.mt;
  const int active0 = (args.count && !args.count_settled) ? min(*args.count, args.n) : active_early0;
  const int active1 =
      !two_jobs ? 0 : ((args2.count && !args2.count_settled) ? min(*args2.count, args2.n) : active_early1);
  const int m_tiles0 = ((active0 + args.bn - 1) / args.bn) * args.tiles2 * args.tiles1;   // pixel tiles of the active slots
  const int m_tiles1 = ((active1 + args.bn - 1) / args.bn) * args.tiles2 * args.tiles1;
  const int units0 = ((m_tiles0 + mt - 1) / mt) * args.c_tiles;
  const int total_tiles = units0 + ((m_tiles1 + mt - 1) / mt) * args.c_tiles;            // work units of the launch
  if (threadIdx.x == 0) DYNMM_TRACE(1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp converged, the TMA /
    // mbarrier-arrive instructions under elect.sync: like the MMA issuer, a lone lane looping inside `if (lane == 0)`
    // pays an ELECT / BRA.U.ANY loop around every UTMALDG)
    {
      // TMA always delivers the full box (out-of-bounds elements arrive as zeros)
      const uint32_t a_tx = args.a_rows * kBlockK * 2;
      const uint32_t sub_tx = args.b1 * args.b2 * args.bn * kBlockK * 2;
      int stage = 0;
      uint32_t phase = 0;
      int aux = 0;
      uint32_t aux_phase = 0;
      auto produce_unit = [&](const KernelArgs& ja, const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& ma2,
                              const CUtensorMap& ma3, const CUtensorMap& mb, const CUtensorMap& mres, int active_j,
                              int m_tiles_j, int unit, int tile) {
        const CUtensorMap* maps_j[4] = {&ma0, &ma1, &ma2, &ma3};
        const uint32_t q = fast_div(unit, args.m_c);
        const int ct = unit - q * args.c_tiles;
        const int nact = min(mt, m_tiles_j - (int)q * mt);          // pixel tiles of this unit (the last may be alone)
        TileCoord t[2];
        int n_in[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          t[w] = decode_tile(args, (q * mt + (w < nact ? w : 0)) * args.c_tiles + ct);
          n_in[w] = ja.in_map ? ja.in_map[t[w].n0] : t[w].n0;
        }
        const uint32_t tx_bytes = nact * a_tx + (args.b_resident ? 0 : b_iter_bytes);
        if (ja.in_f.flags != nullptr) {
          for (int w = 0; w < nact; ++w) wait_tile_inputs(ja, t[w], active_j, lane);
        }
        for (int g = 0; g < args.num_groups; ++g) {
          const Group gp = args.groups[g];
          for (int kc = 0; kc < args.k_chunks; ++kc) {
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * args.stage_bytes;
            // split activations: chunks [0, kc_c) and [kc_c, 2 kc_c) both read the hi half, [2 kc_c, 3 kc_c) the lo
            // half, which starts at channel in_ld / 2
            int ka = kc * kBlockK;
            if (args.kc_c && kc >= args.kc_c)
              ka = kc < 2 * args.kc_c ? (kc - args.kc_c) * kBlockK : ja.in_lo_off + (kc - 2 * args.kc_c) * kBlockK;
            if (elect_one()) {
              mbar_expect_tx(&ctl->full[stage], tx_bytes);
              tma_load_4d(sa, maps_j[gp.map], &ctl->full[stage], ka, t[0].x1 + gp.o1, t[0].x2 + gp.o2, n_in[0]);
              if (nact > 1)
                tma_load_4d(sa + args.a_bytes, maps_j[gp.map], &ctl->full[stage], ka, t[1].x1 + gp.o1,
                            t[1].x2 + gp.o2, n_in[1]);
              if (!args.b_resident)
                tma_load_3d(sa + mt * args.a_bytes, &mb, &ctl->full[stage], kc * kBlockK, t[0].c0, g * args.tpg);
            }
            __syncwarp();
            if (lane == 0 && tile == (int)blockIdx.x && g == 0 && kc == 0) DYNMM_TRACE(2);
            if (++stage == args.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (lane == 0 && tile == (int)blockIdx.x) DYNMM_TRACE(11);
        // residual sub-tiles of this unit's tiles, consumed by the epilogue while the next unit's MMAs run
        for (int w = 0; w < nact; ++w) {
          if (aux_on && t[w].n0 + args.bn <= active_j) {
            const int n_res = ja.res_map ? ja.res_map[t[w].n0] : t[w].n0;
            for (int sub = 0; sub < n_sub; ++sub) {
              mbar_wait(&ctl->aux_empty[aux], aux_phase ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&ctl->aux_full[aux], sub_tx);
                tma_load_4d(smem_aux + aux * kSubBytes, &mres, &ctl->aux_full[aux], t[w].c0 + sub * 64, t[w].x1,
                            t[w].x2, n_res);
              }
              __syncwarp();
              if (++aux == args.aux_slots) {
                aux = 0;
                aux_phase ^= 1;
              }
            }
          }
        }
      };
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (kPerSm == 2 || tile < units0) {
          produce_unit(args, map_a0, map_a1, map_a2, map_a3, map_b, map_res, active0, m_tiles0, tile, tile);
        } else {
          produce_unit(args2, map2_a0, map2_a1, map2_a2, map2_a3, map2_b, map2_res, active1, m_tiles1, tile - units0, tile);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The WHOLE warp runs the loops converged; only the tcgen05 instructions sit under elect.sync.  With one lane
    // running the loop inside `if (lane == 0)` the compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY loop and a
    // thread cannot issue more than one MMA per ~90 cycles (tools/umma_issue_bench.cu: 90 vs 61.5 / 64 cycles per
    // 128x64x16 / 128x128x16 UMMA) -- more than the MMAs themselves take.
    {
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
        // accumulator buffer `acc` holds the unit's mt tiles side by side: columns [acc * mt + w] * acc_stride
        const uint32_t d_tmem = tmem_base + acc * mt * args.acc_stride;
        const bool jb = kPerSm == 1 && tile >= units0;
        const int unit = tile - (jb ? units0 : 0);
        const int nact = min(mt, (jb ? m_tiles1 : m_tiles0) - (unit / args.c_tiles) * mt);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          if (lane == 0 && local == 0 && it == 0) DYNMM_TRACE(3);
          const uint32_t sa = smem_u32(smem + stage * args.stage_bytes);
          const uint32_t sb = args.b_resident ? smem_u32(smem_bres + it * b_iter_bytes) : sa + mt * args.a_bytes;
          if (elect_one()) {
            for (int tp = 0; tp < args.tpg; ++tp) {
              const uint64_t db = umma_desc_sw128(sb + tp * b_tile_bytes);
              for (int w = 0; w < nact; ++w) {        // both pixel tiles against the same weight tile
                const uint64_t da = umma_desc_sw128(sa + w * args.a_bytes + tp * tap_step);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  // advancing K inside the swizzle atom = +32 bytes on the (16-byte unit) start address
                  umma_bf16(d_tmem + w * args.acc_stride, da + (k * 2), db + (k * 2), idesc, (it | tp | k) != 0);
                }
              }
            }
            umma_commit(&ctl->empty[stage]);          // frees the smem stage when these MMAs retire
            if (it == k_iters - 1) umma_commit(&ctl->acc_full[acc]);
          }
          __syncwarp();
          if (lane == 0 && local == 0 && it == 0) DYNMM_TRACE(12);
          if (++stage == args.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (lane == 0 && local == 0) DYNMM_TRACE(4);
        if (lane == 0) DYNMM_TRACE_T(16, local);
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
    // tile-completion flags: a tile whose last TMA store has been committed is published one commit later (then
    // `cp.async.bulk.wait_group 1` proves its stores are in memory without stalling on the store just issued)
    int32_t* pending = nullptr;
    auto publish_flag = [&](int32_t* f) {
      fence_proxy_async_global();
      __threadfence();
      red_release_gpu_add(f, 1);
    };
    auto epilogue_unit = [&](const KernelArgs& ja, const CUtensorMap& mout, const CUtensorMap& mlo, const float* shift_j,
                             int active, int m_tiles_j, int unit) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const bool publish = ja.out_f.flags != nullptr;
      const uint32_t q = fast_div(unit, args.m_c);
      const int ct = unit - q * args.c_tiles;
      const int nact = min(mt, m_tiles_j - (int)q * mt);
      // split operands: the residual halves are read from global memory (no room for a TMA ring next to the hi + lo
      // staging tiles).  The loads of a sub-tile are issued one step ahead -- sub-tile 0 before the wait for the
      // accumulator, sub-tile s + 1 right after sub-tile s is converted -- so their latency hides behind the UMMAs /
      // the TMEM load instead of sitting in the middle of the conversion (measured: 5000 -> 750 cycles per tile).
      uint4 pre_hi[4], pre_lo[4];
      auto prefetch_res = [&](const TileCoord& t, int sub) {
        const int n = t.n0 + nl;
        const int p1 = t.x1 + i1, p2 = t.x2 + i2;
        const int h = args.swap ? p1 : p2, w = args.swap ? p2 : p1;
        const bool valid = nl < args.bn && n < active && h < args.h_out && w < args.w_out;
        const size_t rn = (valid && ja.res_map) ? ja.res_map[n] : n;
        const size_t rpix = (rn * args.h_out + h) * args.w_out + w;
        const int c = t.c0 + sub * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = valid && sub * 64 + half * 32 + j * 8 < args.tile_n && c + j * 8 < args.c_out;
          pre_hi[j] = ok ? __ldg(reinterpret_cast<const uint4*>(ja.residual + rpix * args.res_ld + c + j * 8))
                         : make_uint4(0, 0, 0, 0);
          pre_lo[j] = ok ? __ldg(reinterpret_cast<const uint4*>(ja.residual + rpix * args.res_ld + (args.res_ld >> 1) + c + j * 8))
                         : make_uint4(0, 0, 0, 0);
        }
      };
      constexpr bool kPreRes = kSplit && (kFlags & kFlagRes) != 0;
      if (kPreRes) prefetch_res(decode_tile(args, (q * mt) * args.c_tiles + ct), 0);
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after();
      if (leader && local == 0) DYNMM_TRACE(5);
      if (leader) DYNMM_TRACE_T(32, local);
     for (int wt = 0; wt < nact; ++wt) {             // the unit's pixel tiles, one after the other
      const TileCoord t = decode_tile(args, (q * mt + wt) * args.c_tiles + ct);
      const int n = t.n0 + nl;
      const int p1 = t.x1 + i1, p2 = t.x2 + i2;
      const int h = args.swap ? p1 : p2, w = args.swap ? p2 : p1;
      const bool valid = nl < args.bn && n < active && h < args.h_out && w < args.w_out;
      const bool tile_tma = args.tma_epi && (t.n0 + args.bn <= active);   // uniform over the CTA
      const int h0t = args.swap ? t.x1 : t.x2, w0t = args.swap ? t.x2 : t.x1;    // tile origin in (h, w)
      const size_t pix = valid ? (static_cast<size_t>(n) * args.h_out + h) * args.w_out + w : 0;
      size_t rpix = pix;
      if ((kFlags & kFlagRes) && valid && ja.res_map) {
        rpix = (static_cast<size_t>(ja.res_map[n]) * args.h_out + h) * args.w_out + w;
      }
      float g = 0.f;
      size_t gpix = 0;
      if ((kFlags & kFlagGated) && valid) {
        g = ja.gate[n];
        const int slot = ja.gated_slot ? ja.gated_slot[n] : n;
        gpix = (static_cast<size_t>(slot) * args.h_out + h) * args.w_out + w;
      }
      const uint32_t t_row =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (acc * mt + wt) * args.acc_stride;
      for (int sub = 0; sub < n_sub; ++sub) {
        const int cb = sub * 64 + half * 32;       // first of this thread's 32 columns inside the tile
        const bool cols_live = cb < args.acc_stride;
        uint32_t v[32];
        if (cols_live) {                            // warp-uniform
          if (tile_tma) {                           // second half in flight while the first is converted
            tmem_ld16_half<0>(t_row + cb, v);
            tmem_ld_wait_half<0>(v);
            tmem_ld16_half<16>(t_row + cb + 16, v);
          } else {
            tmem_ld32(t_row + cb, v);
            tmem_ld_wait();
          }
        }
        if (leader && local == 0 && sub == 0) DYNMM_TRACE(13);
        if (tile_tma) {
          uint32_t res_smem = 0;
          if (aux_on) {
            mbar_wait(&ctl->aux_full[aux], aux_phase);
            res_smem = aux_base + aux * kSubBytes;
          }
          const uint32_t out_smem = out_base + sbuf * stage_out_bytes;
          if (cols_live) {
            epilogue_chunk<true, false, kSplit, kAct, true, 0, 16>(kFlags, ja, v, t.c0 + cb, args.tile_n - cb, valid, res_smem,
                                                                   out_smem, half * 4, swz, pix, rpix, gpix, g, shift_j,
                                                                   kPreRes ? pre_hi : nullptr, pre_lo);
            tmem_ld_wait_half<16>(v);
            epilogue_chunk<true, false, kSplit, kAct, true, 16, 32>(kFlags, ja, v, t.c0 + cb, args.tile_n - cb, valid, res_smem,
                                                                    out_smem, half * 4, swz, pix, rpix, gpix, g, shift_j,
                                                                    kPreRes ? pre_hi : nullptr, pre_lo);
          }
          if (kPreRes && sub + 1 < n_sub) prefetch_res(t, sub + 1);
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
          // bulk async-groups belong to the committing thread: the first epilogue warp's elected lane issues,
          // commits and waits for every store (elect.sync under a converged warp: no ELECT loop around UTMASTG)
          if (ewarp == 0 && elect_one()) bulk_wait_read<0>();   // the previous store (other buffer) has drained its smem
          named_barrier(1, 32 * kEpiWarps);
          if (leader && local == 0 && sub == 0) DYNMM_TRACE(15);
          if (ewarp == 0 && elect_one()) {
            tma_store_4d(&mout, smem_stage_out + sbuf * stage_out_bytes, t.c0 + sub * 64, t.x1, t.x2, t.n0);
            if (kSplit)          // lo half through its own map (the residual map slot: split launches have no residual ring)
              tma_store_4d(&mlo, smem_stage_out + sbuf * stage_out_bytes + kSubBytes, t.c0 + sub * 64, t.x1, t.x2, t.n0);
            bulk_commit();
            if (pending != nullptr) {
              bulk_wait<1>();                   // every store but the one just committed is complete
              publish_flag(pending);
            }
          }
          pending = (publish && sub == n_sub - 1) ? ja.out_f.flags + flag_index(ja.out_f, t.n0, h0t, w0t) : nullptr;
          sbuf ^= 1;
        } else if (cols_live) {
          epilogue_chunk<false, false, kSplit, kAct>(kFlags, ja, v, t.c0 + cb, args.tile_n - cb, valid, 0, 0, 0, 0, pix, rpix, gpix, g,
                                        shift_j, kPreRes ? pre_hi : nullptr, pre_lo);
          if (kPreRes && sub + 1 < n_sub) prefetch_res(t, sub + 1);
        }
      }
      if (publish && !tile_tma) {
        // direct stores by every epilogue thread: all of them must be ordered before the flag
        __threadfence();
        named_barrier(2, 32 * kEpiWarps);
        if (leader) red_release_gpu_add(ja.out_f.flags + flag_index(ja.out_f, t.n0, h0t, w0t), 1);
      }
     }
      // this warp is done reading the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc_empty[acc]);
      if (leader && local == 0) DYNMM_TRACE(6);
      if (leader) DYNMM_TRACE_T(48, local);
    };
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      if (kPerSm == 2 || tile < units0) {
        epilogue_unit(args, map_out, map_res, smem_shift, active0, m_tiles0, tile);
      } else {
        epilogue_unit(args2, map2_out, map2_res, smem_shift + shift_stride, active1, m_tiles1, tile - units0);
      }
    }
    if (leader) DYNMM_TRACE(7);
    if (ewarp == 0 && elect_one()) {
      bulk_wait<0>();
      if (pending != nullptr) publish_flag(pending);
    }
    if (leader) {
      DYNMM_TRACE(8);
      if (args.trace) args.trace[blockIdx.x * (DYNMM_TRACE_TILES ? 64 : 16) + 10] = local;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kPerSm == 1 ? 512 : args.tmem_cols);
  }
  if (threadIdx.x == 0) DYNMM_TRACE(9);
}

}  // namespace

}  // namespace dynmm

using namespace dynmm;

namespace {

bool allow_two_per_sm_env() {
  static const bool allow_two = [] {
    const char* e = getenv("DYNMM_CONV_2CTA");
    return !(e && e[0] == '0');
  }();
  return allow_two;
}

// one launch for plan `p0` (and, merged, `p1`: same geometry, own tensors)
int launch_conv(const ConvPlan& p0, const ConvPlan* p1, int max_ctas, bool pdl, cudaStream_t stream) {
  const int sms = num_sms();
  const KernelArgs& a = p0.a;
  int grid = max_ctas > 0 ? max_ctas : (a.two_per_sm ? 2 * sms : sms);
  const int units = p0.max_tiles + (p1 ? p1->max_tiles : 0);
  if (grid > units) grid = units;
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                           KernelArgs, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                           CUtensorMap, KernelArgs);
  static const KernelFn table[64] = {
      conv_igemm_kernel<0, 0>,  conv_igemm_kernel<1, 0>,  conv_igemm_kernel<2, 0>,  conv_igemm_kernel<3, 0>,
      conv_igemm_kernel<4, 0>,  conv_igemm_kernel<5, 0>,  conv_igemm_kernel<6, 0>,  conv_igemm_kernel<7, 0>,
      conv_igemm_kernel<8, 0>,  conv_igemm_kernel<9, 0>,  conv_igemm_kernel<10, 0>, conv_igemm_kernel<11, 0>,
      conv_igemm_kernel<12, 0>, conv_igemm_kernel<13, 0>, conv_igemm_kernel<14, 0>, conv_igemm_kernel<15, 0>,
      conv_igemm_kernel<0, 1>,  conv_igemm_kernel<1, 1>,  conv_igemm_kernel<2, 1>,  conv_igemm_kernel<3, 1>,
      conv_igemm_kernel<4, 1>,  conv_igemm_kernel<5, 1>,  conv_igemm_kernel<6, 1>,  conv_igemm_kernel<7, 1>,
      conv_igemm_kernel<8, 1>,  conv_igemm_kernel<9, 1>,  conv_igemm_kernel<10, 1>, conv_igemm_kernel<11, 1>,
      conv_igemm_kernel<12, 1>, conv_igemm_kernel<13, 1>, conv_igemm_kernel<14, 1>, conv_igemm_kernel<15, 1>,
      conv_igemm_kernel<0, 2>,  conv_igemm_kernel<1, 2>,  conv_igemm_kernel<2, 2>,  conv_igemm_kernel<3, 2>,
      conv_igemm_kernel<4, 2>,  conv_igemm_kernel<5, 2>,  conv_igemm_kernel<6, 2>,  conv_igemm_kernel<7, 2>,
      conv_igemm_kernel<8, 2>,  conv_igemm_kernel<9, 2>,  conv_igemm_kernel<10, 2>, conv_igemm_kernel<11, 2>,
      conv_igemm_kernel<12, 2>, conv_igemm_kernel<13, 2>, conv_igemm_kernel<14, 2>, conv_igemm_kernel<15, 2>,
      conv_igemm_kernel<0, 3>,  conv_igemm_kernel<1, 3>,  conv_igemm_kernel<2, 3>,  conv_igemm_kernel<3, 3>,
      conv_igemm_kernel<4, 3>,  conv_igemm_kernel<5, 3>,  conv_igemm_kernel<6, 3>,  conv_igemm_kernel<7, 3>,
      conv_igemm_kernel<8, 3>,  conv_igemm_kernel<9, 3>,  conv_igemm_kernel<10, 3>, conv_igemm_kernel<11, 3>,
      conv_igemm_kernel<12, 3>, conv_igemm_kernel<13, 3>, conv_igemm_kernel<14, 3>, conv_igemm_kernel<15, 3>};
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run([] {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 64 && e == cudaSuccess; ++i) {
      const bool two = i >= 16 && i < 32;
      e = cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, two ? 113 * 1024 : kSmemBudget);
      if (e == cudaSuccess && two) e = cudaFuncSetAttribute(table[i], cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    return e;
  }));
  static const bool use_pdl = [] {
    const char* e = getenv("DYNMM_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = p0.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (use_pdl && pdl) ? 1 : 0;
  // single job: the second set of maps repeats the first, args2.n == 0 switches the second job off
  const ConvPlan& q = p1 ? *p1 : p0;
  KernelArgs a2 = q.a;
  if (!p1) a2.n = 0;
  DYNMM_CUDA(cudaLaunchKernelEx(&cfg, table[a.flags + (a.split ? 32 : (a.act > 1 ? 48 : (a.two_per_sm ? 16 : 0)))], p0.maps[0], p0.maps[1], p0.maps[2],
                                p0.maps[3], p0.map_b, p0.map_res, p0.map_out, a, q.maps[0], q.maps[1], q.maps[2], q.maps[3],
                                q.map_b, q.map_res, q.map_out, a2));
  return DYNMM_OK;
}

// the tiling fields two merged jobs must agree on
bool same_tiling(const KernelArgs& x, const KernelArgs& y) {
  if (x.b1 != y.b1 || x.b2 != y.b2 || x.bn != y.bn || x.tiles1 != y.tiles1 || x.tiles2 != y.tiles2 ||
      x.c_tiles != y.c_tiles || x.tile_n != y.tile_n || x.num_groups != y.num_groups || x.tpg != y.tpg ||
      x.k_chunks != y.k_chunks || x.stages != y.stages || x.stage_bytes != y.stage_bytes || x.a_bytes != y.a_bytes ||
      x.a_rows != y.a_rows || x.acc_stride != y.acc_stride || x.tma_epi != y.tma_epi || x.aux_slots != y.aux_slots ||
      x.b_resident != y.b_resident || x.two_per_sm != y.two_per_sm || x.mt != y.mt || x.swap != y.swap ||
      x.flags != y.flags || x.h_out != y.h_out || x.w_out != y.w_out || x.c_out != y.c_out || x.split != y.split || x.act != y.act ||
      x.kc_c != y.kc_c || x.in_lo_off != y.in_lo_off || x.out_lo_off != y.out_lo_off)
    return false;
  for (int g = 0; g < x.num_groups; ++g)
    if (x.groups[g].map != y.groups[g].map || x.groups[g].o1 != y.groups[g].o1 || x.groups[g].o2 != y.groups[g].o2)
      return false;
  return true;
}

}  // namespace

extern "C" int dynmm_conv_igemm_fwd(const dynmm_conv_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ConvPlan plan;
  int rc = plan_conv(p, &plan, num_sms(), kSmemBudget, allow_two_per_sm_env(), /*allow_dual=*/true);
  if (rc) return rc;
  return launch_conv(plan, nullptr, p->max_ctas, !(p->flags & DYNMM_CONV_VOLATILE_WEIGHTS), stream);
}

extern "C" int dynmm_conv_igemm_fwd2(const dynmm_conv_params* pa, const dynmm_conv_params* pb, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(pa && pb, "conv_igemm_fwd2: null pointer");
  if (pa->trace || pb->trace || pa->max_ctas || pb->max_ctas || pa->in_flags.flags || pb->in_flags.flags ||
      pa->out_flags.flags || pb->out_flags.flags || pa->n < 1 || pb->n < 1) {
    set_error("conv_igemm_fwd2: trace / max_ctas / tile flags are single-launch options");
    return DYNMM_EUNSUPPORTED;
  }
  // static so that two ~1.3 KB plans (14 tensor maps) do not live on a small thread stack twice over
  ConvPlan plan_a, plan_b;
  int rc = plan_conv(pa, &plan_a, num_sms(), kSmemBudget, false, /*allow_dual=*/true, /*partner_slots=*/pb->n);
  if (rc) return rc;
  rc = plan_conv(pb, &plan_b, num_sms(), kSmemBudget, false, /*allow_dual=*/true, /*partner_slots=*/pa->n);
  if (rc) return rc;
  if (!same_tiling(plan_a.a, plan_b.a)) {
    set_error("conv_igemm_fwd2: the two convolutions do not plan to the same tiling (launch them separately)");
    return DYNMM_EUNSUPPORTED;
  }
  const bool pdl = !((pa->flags | pb->flags) & DYNMM_CONV_VOLATILE_WEIGHTS);
  return launch_conv(plan_a, &plan_b, 0, pdl, stream);
}

extern "C" int dynmm_conv_tile_grid(const dynmm_conv_params* p, dynmm_tile_flags* grid) {
  DYNMM_CHECK_ARG(p && grid, "conv_tile_grid: null pointer");
  ConvPlan plan;
  dynmm_conv_params q = *p;
  q.out_flags.flags = nullptr;            // geometry only
  int rc = plan_conv(&q, &plan, num_sms(), kSmemBudget, allow_two_per_sm_env(), /*allow_dual=*/true);
  if (rc) return rc;
  int32_t* keep = grid->flags;
  *grid = plan.a.out_f;
  grid->flags = keep;
  return DYNMM_OK;
}
