// Convolution PROGRAM: many dependent tcgen05 implicit-GEMM convolutions in ONE persistent launch.
//
// At batch 8 every encoder / decoder convolution of the gated RGB-D network is 2-10 us of work, and a
// launch boundary costs as much again (tail drain, launch gap, barrier / TMEM / descriptor prologue,
// pipeline fill: in-kernel cycle traces in profiles/).  A program is a list of PHASES; a phase holds up to
// four independent convolutions (the RGB and the depth encoder run in lock step, the depth encoder one
// layer ahead so that its stage output exists when the RGB stage's last conv adds g_s * depth_s; a block's
// 1x1 down-sampling conv rides with its first 3x1 conv).  One cooperative grid of persistent CTAs walks the
// phases; between phases a grid-wide barrier (one L2 atomic per CTA) replaces the kernel boundary:
//
//   * mbarriers, the 512 TMEM columns and the warp roles live for the whole program;
//   * the CTAs of a phase are split between its convolutions in proportion to their tile counts, which are
//     derived ON THE DEVICE from the gate's `count` (samples the gate switched off produce no tiles);
//   * a CTA fetches the next phase's resident weights / shift vector right after it has arrived at the
//     barrier, i.e. while it waits for the slowest CTA of the current phase;
//   * tensors written in one phase and read in the next stay in L2 (ld.global.cg / TMA, never the
//     non-coherent path).
//
// The per-tile work (TMA producer, MMA issue, 8 epilogue warps) is the code of conv_igemm.cu, reading its KernelArgs
// from shared memory and its tensor maps from the device-resident program image; two MMA-issuing warps work on
// alternate tiles when a CTA has several tiles and enough stages (dual mode).  A convolution run through a program
// is bit-identical to the same convolution launched on its own.
#include <string.h>

#include "conv_plan.cuh"

namespace dynmm {

namespace {

using namespace convk;

constexpr int kProgThreads = kThreads + 32;           // producer, MMA issuer A, 8 epilogue warps, MMA issuer B
constexpr int kMaxPhases = 96;
constexpr int kMaxPhaseJobs = 4;
constexpr int kCtlBytes = 4096;                       // control block in front of a job's shared-memory layout
constexpr int kProgBudget = kSmemBudget - kCtlBytes;  // what plan_conv may use per job
constexpr int kHeaderBytes = 1024;
constexpr unsigned kSpinLimit = 1u << 22;             // watchdog: a broken program traps instead of hanging the GPU

struct JobBrief {                // what the CTA split of a phase needs to know about a job (32 bytes)
  const int32_t* count;
  int n, bn, tiles_per_group, cost;
  int pad[2];
};
struct ProgramHeader {
  int n_phases, n_jobs, grid, smem_bytes;
  int phase_begin[kMaxPhases + 1];
};
static_assert(sizeof(ProgramHeader) <= kHeaderBytes, "program header");
static_assert(sizeof(JobBrief) == 32, "JobBrief");
static_assert(sizeof(ConvPlan) % 64 == 0, "tensor maps need 64-byte alignment");

// What the MMA-issuing warps need to know about a job, as KERNEL PARAMETERS (constant bank) next to the device
// image.  (The job index of a CTA is data dependent, so these values still arrive in ordinary registers; what keeps
// the issue rate up is the converged-warp / elect.sync pattern of the issuers, see conv_igemm.cu.)
constexpr int kMaxJobs = 160;
struct MmaJob {
  uint32_t idesc;
  int tap_step, b_tile_bytes, b_iter_bytes, tpg, k_iters, stages, stage_bytes, a_bytes, b_resident, bres_off, pad;
};
struct ProgramParams {
  int n_phases, n_jobs;
  int phase_begin[kMaxPhases + 1];
  MmaJob jobs[kMaxJobs];
};
static_assert(sizeof(ProgramParams) <= 16 * 1024, "kernel parameter space");
constexpr uint32_t kAccCols = 256;   // accumulator buffer b lives at TMEM column b * 256 (the CTA owns all 512)

struct PhaseCtl {                // this CTA's share of one phase
  int job;                       // index into the program's jobs, -1: idle in this phase
  int cta_local, ctas_job;       // position among the CTAs working on `job`
  int active;                    // active sample slots of the job
};
// control block layout
constexpr int kOffArgs = 256, kOffPhase = 1024;
static_assert(sizeof(SmemCtl) <= kOffArgs, "SmemCtl");
static_assert(kOffArgs + sizeof(KernelArgs) <= kOffPhase, "KernelArgs copy");
static_assert(kOffPhase + kMaxPhases * sizeof(PhaseCtl) <= kCtlBytes, "phase table");

inline size_t briefs_offset() { return kHeaderBytes; }
inline size_t plans_offset(int n_jobs) { return kHeaderBytes + ((size_t)n_jobs * sizeof(JobBrief) + 127) / 128 * 128; }

// epilogue_chunk with the flag combinations the network uses folded at compile time (as in the one-launch kernel,
// where they are a template parameter); anything else takes the run-time path
template <bool kTma>
__device__ __forceinline__ void epilogue_chunk_sw(int flags, const KernelArgs& args, const uint32_t (&v)[32], int c_first,
                                                  int cols_left, bool valid, uint32_t res_smem, uint32_t out_smem,
                                                  uint32_t chunk0, uint32_t swz, size_t pix, size_t rpix, size_t gpix,
                                                  float g, const float* shift_smem) {
#define DYNMM_EPI_CASE(F)                                                                                             \
  case (F):                                                                                                           \
    epilogue_chunk<kTma, true>((F), args, v, c_first, cols_left, valid, res_smem, out_smem, chunk0, swz, pix, rpix, gpix, \
                               g, shift_smem);                                                                        \
    break;
  switch (flags) {
    DYNMM_EPI_CASE(kFlagRelu)
    DYNMM_EPI_CASE(kFlagRelu | kFlagRes)
    DYNMM_EPI_CASE(kFlagRelu | kFlagRes | kFlagGated)
    DYNMM_EPI_CASE(0)
    default:
      epilogue_chunk<kTma, true>(flags, args, v, c_first, cols_left, valid, res_smem, out_smem, chunk0, swz, pix, rpix,
                                 gpix, g, shift_smem);
  }
#undef DYNMM_EPI_CASE
}

__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// CTA split of one phase, computed redundantly (and identically) by every CTA: `c[j]` CTAs for job j in
// proportion to tiles x cost, at least one per non-empty job, never more than the job has tiles.
__device__ void split_phase(const ProgramHeader* hdr, const JobBrief* briefs, int phase, int grid, int cta,
                            PhaseCtl* out) {
  const int pb = hdr->phase_begin[phase], pe = hdr->phase_begin[phase + 1];
  const int nj = min(pe - pb, kMaxPhaseJobs);
  int tiles[kMaxPhaseJobs], act[kMaxPhaseJobs], c[kMaxPhaseJobs];
  long long w[kMaxPhaseJobs];
  long long wsum = 0;
#pragma unroll
  for (int j = 0; j < kMaxPhaseJobs; ++j) {
    tiles[j] = act[j] = c[j] = 0;
    w[j] = 0;
    if (j < nj) {
      const JobBrief b = briefs[pb + j];
      int a = b.n;
      if (b.count) a = min(__ldcg(b.count), b.n);
      a = max(a, 0);
      act[j] = a;
      tiles[j] = ((a + b.bn - 1) / b.bn) * b.tiles_per_group;
      w[j] = (long long)tiles[j] * b.cost;
      wsum += w[j];
    }
  }
  out->job = -1;
  out->cta_local = 0;
  out->ctas_job = 1;
  out->active = 0;
  if (wsum == 0) return;
  int sum = 0;
#pragma unroll
  for (int j = 0; j < kMaxPhaseJobs; ++j) {
    if (tiles[j] > 0) {
      int cj = (int)((long long)grid * w[j] / wsum);
      cj = max(1, min(cj, tiles[j]));
      c[j] = cj;
      sum += cj;
    }
  }
  while (sum > grid) {                     // only when the forced minimum of one CTA per job overshoots
    int best = 0;
#pragma unroll
    for (int j = 1; j < kMaxPhaseJobs; ++j)
      if (c[j] > c[best]) best = j;
    if (c[best] <= 1) break;
    --c[best];
    --sum;
  }
  for (int it = 0; it < 8 && sum < grid; ++it) {   // hand rounding leftovers to the most loaded job
    int best = -1;
    long long best_num = 0, best_den = 1;
#pragma unroll
    for (int j = 0; j < kMaxPhaseJobs; ++j) {
      if (c[j] > 0 && c[j] < tiles[j]) {
        if (best < 0 || w[j] * best_den > best_num * c[j]) {
          best = j;
          best_num = w[j];
          best_den = c[j];
        }
      }
    }
    if (best < 0) break;
    ++c[best];
    ++sum;
  }
  int cum = 0;
#pragma unroll
  for (int j = 0; j < kMaxPhaseJobs; ++j) {
    if (c[j] > 0 && cta >= cum && cta < cum + c[j]) {
      out->job = pb + j;
      out->cta_local = cta - cum;
      out->ctas_job = c[j];
      out->active = act[j];
    }
    cum += c[j];
  }
}

// the UMMAs of one A-tile load: `tpg` taps x 4 k-steps, accumulator at the compile-time TMEM column kAcc
template <uint32_t kAcc>
__device__ __forceinline__ void issue_kiter(uint32_t sa, uint32_t sb, const MmaJob& mj, int it) {
  for (int tp = 0; tp < mj.tpg; ++tp) {
    const uint64_t da = umma_desc_sw128(sa + tp * mj.tap_step);
    const uint64_t db = umma_desc_sw128(sb + tp * mj.b_tile_bytes);
#pragma unroll
    for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_bf16(kAcc, da + (k * 2), db + (k * 2), mj.idesc, (it | tp | k) != 0);
  }
}

__global__ void __launch_bounds__(kProgThreads, 1)
conv_program_kernel(const __grid_constant__ ProgramParams params, const uint8_t* __restrict__ image, unsigned* barrier,
                    unsigned long long* trace) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(base);
  KernelArgs* sargs = reinterpret_cast<KernelArgs*>(base + kOffArgs);
  PhaseCtl* phases = reinterpret_cast<PhaseCtl*>(base + kOffPhase);
  uint8_t* smem = base + kCtlBytes;

  const ProgramHeader* hdr = reinterpret_cast<const ProgramHeader*>(image);
  const int n_phases = params.n_phases;
  const JobBrief* briefs = reinterpret_cast<const JobBrief*>(image + kHeaderBytes);
  const ConvPlan* plans = reinterpret_cast<const ConvPlan*>(
      image + kHeaderBytes + ((size_t)hdr->n_jobs * sizeof(JobBrief) + 127) / 128 * 128);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grid = gridDim.x;

  // ---------------------------------------------------------------- once per program
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], kEpiWarps);
    }
    for (int i = 0; i < kAuxSlots; ++i) {
      mbar_init(&ctl->aux_full[i], 1);
      mbar_init(&ctl->aux_empty[i], kEpiWarps);
    }
    mbar_init(&ctl->b_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  // the CTA split of every phase depends only on the gate's counts, which were final before the launch
  for (int p = threadIdx.x; p < n_phases; p += kProgThreads) split_phase(hdr, briefs, p, grid, blockIdx.x, &phases[p]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (ctl->tmem_base != 0) __trap();      // all 512 columns are ours: the allocation starts at 0
  constexpr uint32_t tmem_base = 0;
  const uint32_t smem_u = smem_u32(smem);

  // pipeline state that outlives a phase (per-barrier parity bits; the stage ring restarts at 0 every phase)
  uint32_t empty_bits = 0, full_bits = 0;           // producer / MMA view of the stage ring
  uint32_t aux_empty_bits = 0, aux_full_bits = 0;   // producer / epilogue view of the residual ring
  uint32_t bres_parity = 0;                         // MMA: resident-weight barrier
  uint32_t acc_cnt0 = 0, acc_cnt1 = 0;              // uses of accumulator buffer 0 / 1 so far (an issuer: of ITS buffer)
  int sbuf = 0;                                     // epilogue: staging buffer of the next TMA store

  for (int phase = 0; phase < n_phases; ++phase) {
    const PhaseCtl pc = phases[phase];
    const bool busy = pc.job >= 0;
    const ConvPlan* jd = plans + (busy ? pc.job : 0);
    // ---- this phase's KernelArgs -> shared memory (warp 2), then shift vector + resident weights
    if (busy && warp == 2) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&jd->a);
      uint32_t* dst = reinterpret_cast<uint32_t*>(sargs);
      for (int i = lane; i < (int)(sizeof(KernelArgs) / 4); i += 32) dst[i] = __ldcg(src + i);
    }
    __syncthreads();
    // a per-thread REGISTER copy (only the fields a role uses survive): every PTX asm below clobbers "memory",
    // so fields read through the shared-memory pointer would be re-loaded after each tcgen05 / mbarrier op
    const KernelArgs args = *sargs;
    int k_iters = 0, b_tile_bytes = 0, b_iter_bytes = 0, n_sub = 0, total_tiles = 0;
    uint8_t *smem_bres = smem, *smem_aux = smem, *smem_stage_out = smem;
    float* smem_shift = reinterpret_cast<float*>(smem);
    bool aux_on = false, my_tiles = false;
    if (busy) {
      k_iters = args.num_groups * args.k_chunks;
      b_tile_bytes = args.tile_n * kBlockK * 2;
      b_iter_bytes = args.tpg * b_tile_bytes;
      smem_bres = smem + args.stages * args.stage_bytes;
      smem_aux = smem_bres + (args.b_resident ? k_iters * b_iter_bytes : 0);
      smem_stage_out = smem_aux + args.aux_slots * kSubBytes;
      smem_shift = reinterpret_cast<float*>(
          (reinterpret_cast<uintptr_t>(smem_stage_out + (args.tma_epi ? 2 * kSubBytes : 0)) + 15) & ~uintptr_t(15));
      n_sub = (args.tile_n + 63) >> 6;
      aux_on = (args.flags & kFlagRes) && args.tma_epi && args.aux_slots > 0;
      const int n_groups = (pc.active + args.bn - 1) / args.bn;
      total_tiles = n_groups * args.tiles2 * args.tiles1 * args.c_tiles;
      my_tiles = pc.cta_local < total_tiles;
      if (warp >= 2 && warp < 2 + kEpiWarps) {
        for (int c = threadIdx.x - 64; c < args.c_out; c += 32 * kEpiWarps)
          smem_shift[c] = args.shift ? __ldg(args.shift + c) : 0.f;
      }
      if (warp == 0 && lane == 0 && my_tiles) {
        tma_prefetch_desc(&jd->maps[0]);
        tma_prefetch_desc(&jd->map_b);
        if (args.tma_epi) tma_prefetch_desc(&jd->map_out);
        if (aux_on) tma_prefetch_desc(&jd->map_res);
        if (args.b_resident) {
          // weights are constants: fetch them while the slowest CTA of the previous phase is still working
          mbar_expect_tx(&ctl->b_full, k_iters * b_iter_bytes);
          for (int g = 0; g < args.num_groups; ++g)
            for (int kc = 0; kc < args.k_chunks; ++kc)
              tma_load_3d(smem_bres + (g * args.k_chunks + kc) * b_iter_bytes, &jd->map_b, &ctl->b_full, kc * kBlockK,
                          0, g * args.tpg);
        }
      }
    }
    // ---- everything the previous phase wrote is visible from here on
    if (threadIdx.x == 0 && phase > 0) {
      const unsigned target = (unsigned)phase * (unsigned)grid;
      unsigned spins = 0;
      while (ld_acquire_gpu(barrier) < target) {
        if (++spins > kSpinLimit) {
          barrier[1] = 0xDEADu;
          __trap();
        }
      }
      __threadfence();
      fence_proxy_async_all();
    }
    if (trace && threadIdx.x == 0) {      // debug timeline: [cta][phase] = globaltimer when the CTA may start the phase
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      trace[(size_t)blockIdx.x * (kMaxPhases + 1) + phase] = t;
    }
    __syncthreads();

    // Accumulator buffer of the CTA's k-th tile in the phase: k & 1.  DUAL mode (>= 2 tiles and >= 4 stages): tile k
    // belongs to issuer k & 1 and the stage ring is split in two halves, ring r = slots [r * half, (r + 1) * half),
    // filled by the producer for issuer r's tiles only.  SINGLE mode (one tile, or too few stages to split -- the
    // deep layers with K = 1536 need the whole ring as prefetch depth): issuer 0 works through all tiles on the whole
    // ring.  Either way every mbarrier has exactly one producer and one consumer per phase.
    const int n_my = my_tiles ? (total_tiles - pc.cta_local + pc.ctas_job - 1) / pc.ctas_job : 0;
    const bool dual = busy && args.stages >= 4 && n_my >= 2;
    auto mma_role = [&](const uint32_t me) {
      const MmaJob& mj = params.jobs[pc.job];
      const int ring = dual ? (mj.stages >> 1) : mj.stages;
      const int slot0 = dual ? me * ring : 0;
      if (mj.b_resident) {
        mbar_wait_wd(&ctl->b_full, bres_parity);
        bres_parity ^= 1u;
        tc_fence_after();
      }
      if (dual || me == 0) {
        int slot = 0;
        for (int k = dual ? me : 0; k < n_my; k += dual ? 2 : 1) {
          const uint32_t acc = k & 1;
          const uint32_t uses_before = (acc ? acc_cnt1 : acc_cnt0) + (k >> 1);    // of this accumulator buffer
          mbar_wait_wd(&ctl->acc_empty[acc], (uses_before & 1u) ^ 1u);
          tc_fence_after();
          for (int it = 0; it < mj.k_iters; ++it) {
            const int stage = slot0 + slot;
            mbar_wait_wd(&ctl->full[stage], (full_bits >> stage) & 1u);
            full_bits ^= 1u << stage;
            tc_fence_after();
            const uint32_t sa = smem_u + stage * mj.stage_bytes;
            const uint32_t sb = mj.b_resident ? smem_u + mj.bres_off + it * mj.b_iter_bytes : sa + mj.a_bytes;
            if (elect_one()) {       // the warp runs the loops converged; only the tcgen05 instructions are single-lane
              if (acc == 0) {
                issue_kiter<0>(sa, sb, mj, it);
              } else {
                issue_kiter<kAccCols>(sa, sb, mj, it);
              }
              umma_commit(&ctl->empty[stage]);
              if (it == mj.k_iters - 1) umma_commit(&ctl->acc_full[acc]);
            }
            __syncwarp();
            if (++slot == ring) slot = 0;
          }
        }
      }
      // both issuers keep complete books: uses of the two accumulators, and the parity of the stage barriers the OTHER
      // issuer consumed in this phase (the ring split changes from phase to phase)
      acc_cnt0 += (n_my + 1) >> 1;
      acc_cnt1 += n_my >> 1;
      if (dual || me == 1) {
        const uint32_t other = me ^ 1u;
        const int other_tiles = dual ? ((n_my + (other == 0 ? 1 : 0)) >> 1) : n_my;
        const int other_slot0 = dual ? other * ring : 0;
        const int uses = other_tiles * mj.k_iters;
        for (int j = 0; j < ring; ++j) {
          const int cnt = uses / ring + (j < uses % ring ? 1 : 0);
          if (cnt & 1) full_bits ^= 1u << (other_slot0 + j);
        }
      }
    };
    if (busy && my_tiles) {
      if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (whole warp converged, TMA and
        // mbarrier-arrive instructions under elect.sync)
        {
          const CUtensorMap* maps[4] = {&jd->maps[0], &jd->maps[1], &jd->maps[2], &jd->maps[3]};
          const uint32_t a_tx = args.a_rows * kBlockK * 2;
          const uint32_t tx_bytes = a_tx + (args.b_resident ? 0 : b_iter_bytes);
          const uint32_t sub_tx = args.b1 * args.b2 * args.bn * kBlockK * 2;
          const int half = dual ? (args.stages >> 1) : args.stages;    // slots per ring (single mode: one ring)
          int slot[2] = {0, 0};
          int aux = 0;
          for (int k = 0; k < n_my; k += dual ? 2 : 1) {
            // dual mode: the loads of a pair of tiles are interleaved k-iteration by k-iteration, so both issuers advance
            const int pair = dual ? min(2, n_my - k) : 1;
            TileCoord t[2];
            int n_in[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              t[r] = decode_tile(args, pc.cta_local + (k + (r < pair ? r : 0)) * pc.ctas_job);
              n_in[r] = args.in_map ? __ldg(args.in_map + t[r].n0) : t[r].n0;
            }
            for (int g = 0; g < args.num_groups; ++g) {
              const Group gp = sargs->groups[g];     // dynamic index: stays in shared memory
              for (int kc = 0; kc < args.k_chunks; ++kc) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                  if (r < pair) {
                    const int stage = r * half + slot[r];
                    mbar_wait_wd(&ctl->empty[stage], ((empty_bits >> stage) & 1u) ^ 1u);
                    empty_bits ^= 1u << stage;
                    uint8_t* sa = smem + stage * args.stage_bytes;
                    if (elect_one()) {
                      mbar_expect_tx(&ctl->full[stage], tx_bytes);
                      tma_load_4d(sa, maps[gp.map], &ctl->full[stage], kc * kBlockK, t[r].x1 + gp.o1, t[r].x2 + gp.o2,
                                  n_in[r]);
                      if (!args.b_resident)
                        tma_load_3d(sa + args.a_bytes, &jd->map_b, &ctl->full[stage], kc * kBlockK, t[r].c0, g * args.tpg);
                    }
                    __syncwarp();
                    if (++slot[r] == half) slot[r] = 0;
                  }
                }
              }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              if (r < pair && aux_on && t[r].n0 + args.bn <= pc.active) {
                const int n_res = args.res_map ? __ldg(args.res_map + t[r].n0) : t[r].n0;
                for (int sub = 0; sub < n_sub; ++sub) {
                  mbar_wait_wd(&ctl->aux_empty[aux], ((aux_empty_bits >> aux) & 1u) ^ 1u);
                  aux_empty_bits ^= 1u << aux;
                  if (elect_one()) {
                    mbar_expect_tx(&ctl->aux_full[aux], sub_tx);
                    tma_load_4d(smem_aux + aux * kSubBytes, &jd->map_res, &ctl->aux_full[aux], t[r].c0 + sub * 64, t[r].x1,
                                t[r].x2, n_res);
                  }
                  __syncwarp();
                  if (++aux == args.aux_slots) aux = 0;
                }
              }
            }
          }
        }
      } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuers: warp 1 -> accumulator buffer 0,
        // warp 10 -> buffer 1 (alternate tiles).  One thread needs ~70-100 issue cycles per tcgen05.mma here (its
        // operands travel from ordinary to uniform registers behind an ELECT / R2UR / branch sequence), more than
        // the 32 / 64 cycles a 128 x 64 / 128 x 128 x 16 MMA executes: two issuers keep the tensor pipe fed.
        mma_role(0);
      } else if (warp == kEpiWarps + 2) {
        mma_role(1);
      } else {
        // ------------------------------------------------------------ epilogue (8 warps)
        const int ewarp = warp - 2;
        const int quarter = warp & 3;
        const int half = ewarp >> 2;
        const int row = quarter * 32 + lane;
        const int i1 = row % args.b1;
        const int i2 = (row / args.b1) % args.b2;
        const int nl = row / (args.b1 * args.b2);
        const uint32_t row_off = row * 128;
        const uint32_t swz = row & 7;
        const uint32_t aux_base = smem_u32(smem_aux) + row_off;
        const uint32_t out_base = smem_u32(smem_stage_out) + row_off;
        int aux = 0;
        const int flags = args.flags;
        for (int k = 0; k < n_my; ++k) {
          const int tile = pc.cta_local + k * pc.ctas_job;
          const int acc = k & 1;
          const uint32_t acc_phase = (acc ? acc_cnt1 : acc_cnt0) & 1u;
          if (acc) {
            ++acc_cnt1;
          } else {
            ++acc_cnt0;
          }
          const TileCoord t = decode_tile(args, tile);
          const int n = t.n0 + nl;
          const int p1 = t.x1 + i1, p2 = t.x2 + i2;
          const int h = args.swap ? p1 : p2, w = args.swap ? p2 : p1;
          const bool valid = nl < args.bn && n < pc.active && h < args.h_out && w < args.w_out;
          const bool tile_tma = args.tma_epi && (t.n0 + args.bn <= pc.active);
          const size_t pix = valid ? (static_cast<size_t>(n) * args.h_out + h) * args.w_out + w : 0;
          size_t rpix = pix;
          if ((flags & kFlagRes) && valid && args.res_map) {
            rpix = (static_cast<size_t>(__ldg(args.res_map + n)) * args.h_out + h) * args.w_out + w;
          }
          float g = 0.f;
          size_t gpix = 0;
          if ((flags & kFlagGated) && valid) {
            g = __ldg(args.gate + n);
            const int slot = args.gated_slot ? __ldg(args.gated_slot + n) : n;
            gpix = (static_cast<size_t>(slot) * args.h_out + h) * args.w_out + w;
          }
          mbar_wait_wd(&ctl->acc_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccCols;
          for (int sub = 0; sub < n_sub; ++sub) {
            const int cb = sub * 64 + half * 32;
            const bool cols_live = cb < args.acc_stride;
            uint32_t v[32];
            if (cols_live) {
              tmem_ld32(t_row + cb, v);
              tmem_ld_wait();
            }
            if (tile_tma) {
              uint32_t res_smem = 0;
              if (aux_on) {
                mbar_wait_wd(&ctl->aux_full[aux], (aux_full_bits >> aux) & 1u);
                aux_full_bits ^= 1u << aux;
                res_smem = aux_base + aux * kSubBytes;
              }
              const uint32_t out_smem = out_base + sbuf * kSubBytes;
              if (cols_live) {
                epilogue_chunk_sw<true>(flags, args, v, t.c0 + cb, args.tile_n - cb, valid, res_smem, out_smem,
                                           half * 4, swz, pix, rpix, gpix, g, smem_shift);
              }
              if (aux_on) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->aux_empty[aux]);
                if (++aux == args.aux_slots) aux = 0;
              }
              fence_async_smem();
              if (ewarp == 0 && elect_one()) bulk_wait_read<0>();
              named_barrier(1, 32 * kEpiWarps);
              if (ewarp == 0 && elect_one()) {
                tma_store_4d(&jd->map_out, smem_stage_out + sbuf * kSubBytes, t.c0 + sub * 64, t.x1, t.x2, t.n0);
                bulk_commit();
              }
              sbuf ^= 1;
            } else if (cols_live) {
              epilogue_chunk_sw<false>(flags, args, v, t.c0 + cb, args.tile_n - cb, valid, 0, 0, 0, 0, pix, rpix,
                                          gpix, g, smem_shift);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->acc_empty[acc]);
        }
        // this phase's outputs must be complete before the CTA reports at the grid barrier:
        // TMA stores fully written (not just read out of shared memory), direct stores ordered before later TMA reads
        if (ewarp == 0 && elect_one()) bulk_wait<0>();
        fence_proxy_async_all();
      }
    }

    // ---- end of phase: the roles of this CTA are done -> arrive at the grid barrier
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0 && phase + 1 < n_phases) {
      __threadfence();
      atomicAdd(barrier, 1u);
    }
  }

  if (trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    trace[(size_t)blockIdx.x * (kMaxPhases + 1) + n_phases] = t;
  }
  if (warp == 1) {
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

}  // namespace dynmm

using namespace dynmm;

extern "C" long long dynmm_conv_program_bytes(int n_jobs) {
  if (n_jobs < 1) return -1;
  return (long long)(plans_offset(n_jobs) + (size_t)n_jobs * sizeof(ConvPlan));
}

extern "C" int dynmm_conv_program_build(const dynmm_conv_params* jobs, const int32_t* phase_of_job, int n_jobs,
                                        void* image_, long long image_bytes, int32_t* launch_cfg) {
  DYNMM_CHECK_ARG(jobs && phase_of_job && image_ && launch_cfg, "conv_program_build: null pointer");
  for (int j = 0; j < n_jobs; ++j)
    DYNMM_CHECK_ARG(!(jobs[j].flags & DYNMM_CONV_SPLIT) && jobs[j].relu <= 1,
                    "conv_program_build: DYNMM_CONV_SPLIT / swish / h-swish jobs are per-launch only");
  DYNMM_CHECK_ARG(n_jobs >= 1 && n_jobs <= kMaxJobs, "conv_program_build: 1..%d convolutions per program", kMaxJobs);
  DYNMM_CHECK_ARG(image_bytes >= dynmm_conv_program_bytes(n_jobs), "conv_program_build: image buffer too small");
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(image_) & 127) == 0, "conv_program_build: image must be 128-byte aligned");
  uint8_t* image = static_cast<uint8_t*>(image_);
  memset(image, 0, (size_t)dynmm_conv_program_bytes(n_jobs));
  ProgramHeader* hdr = reinterpret_cast<ProgramHeader*>(image);
  JobBrief* briefs = reinterpret_cast<JobBrief*>(image + briefs_offset());
  ConvPlan* plans = reinterpret_cast<ConvPlan*>(image + plans_offset(n_jobs));
  const int sms = num_sms();
  int phase = -1, n_phases = 0, max_smem = 0, max_phase_tiles = 0, phase_tiles = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const int ph = phase_of_job[j];
    DYNMM_CHECK_ARG(ph >= phase && ph <= phase + 1 && (ph >= 0), "conv_program_build: phases must be 0,1,2,... in job order "
                    "(job %d has phase %d after %d)", j, ph, phase);
    if (ph != phase) {
      DYNMM_CHECK_ARG(n_phases < kMaxPhases, "conv_program_build: more than %d phases", kMaxPhases);
      hdr->phase_begin[n_phases++] = j;
      phase = ph;
      phase_tiles = 0;
    }
    DYNMM_CHECK_ARG(j - hdr->phase_begin[n_phases - 1] < kMaxPhaseJobs, "conv_program_build: more than %d convolutions in "
                    "phase %d", kMaxPhaseJobs, ph);
    DYNMM_CHECK_ARG(jobs[j].trace == nullptr && jobs[j].max_ctas == 0, "conv_program_build: trace / max_ctas are per-launch "
                    "options");
    int rc = plan_conv(&jobs[j], &plans[j], sms, kProgBudget);
    if (rc) return rc;
    const KernelArgs& a = plans[j].a;
    briefs[j].count = a.count;
    briefs[j].n = a.n;
    briefs[j].bn = a.bn;
    briefs[j].tiles_per_group = a.tiles1 * a.tiles2 * a.c_tiles;
    briefs[j].cost = a.cost;
    if (plans[j].smem_bytes > max_smem) max_smem = plans[j].smem_bytes;
    phase_tiles += plans[j].max_tiles;
    if (phase_tiles > max_phase_tiles) max_phase_tiles = phase_tiles;
  }
  hdr->phase_begin[n_phases] = n_jobs;
  hdr->n_phases = n_phases;
  hdr->n_jobs = n_jobs;
  hdr->grid = max_phase_tiles < sms ? max_phase_tiles : sms;
  hdr->smem_bytes = max_smem + kCtlBytes;
  DYNMM_CHECK_ARG(hdr->smem_bytes <= kSmemBudget, "conv_program_build: internal smem accounting error");
  launch_cfg[0] = hdr->grid;
  launch_cfg[1] = hdr->smem_bytes;
  launch_cfg[2] = n_phases;
  return DYNMM_OK;
}

extern "C" int dynmm_conv_program_launch(const void* image_dev, const void* image_host_, void* barrier, void* trace,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYNMM_CHECK_ARG(image_dev && image_host_ && barrier, "conv_program_launch: null pointer");
  DYNMM_CHECK_ARG((reinterpret_cast<uintptr_t>(image_dev) & 127) == 0, "conv_program_launch: image must be 128-byte aligned");
  const uint8_t* image_host = static_cast<const uint8_t*>(image_host_);
  const ProgramHeader* hdr = reinterpret_cast<const ProgramHeader*>(image_host);
  const int grid = hdr->grid, smem_bytes = hdr->smem_bytes, n_jobs = hdr->n_jobs;
  DYNMM_CHECK_ARG(grid >= 1 && grid <= num_sms() && smem_bytes > 0 && smem_bytes <= kSmemBudget && n_jobs >= 1 &&
                      n_jobs <= kMaxJobs && hdr->n_phases >= 1 && hdr->n_phases <= kMaxPhases,
                  "conv_program_launch: not a program image (grid %d, smem %d, jobs %d)", grid, smem_bytes, n_jobs);
  static PerDeviceOnce attr_once;
  DYNMM_CUDA(attr_once.run(
      [] { return cudaFuncSetAttribute(conv_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); }));
  // the MMA thread's view of every job goes into the kernel parameters (constant bank -> uniform registers)
  static thread_local ProgramParams params;
  params.n_phases = hdr->n_phases;
  params.n_jobs = n_jobs;
  for (int i = 0; i <= hdr->n_phases; ++i) params.phase_begin[i] = hdr->phase_begin[i];
  const ConvPlan* plans = reinterpret_cast<const ConvPlan*>(image_host + plans_offset(n_jobs));
  for (int j = 0; j < n_jobs; ++j) {
    const KernelArgs& a = plans[j].a;
    MmaJob& m = params.jobs[j];
    m.idesc = umma_idesc_bf16(kBlockM, a.tile_n);
    m.tap_step = a.b1 * kBlockK * 2;
    m.b_tile_bytes = a.tile_n * kBlockK * 2;
    m.tpg = a.tpg;
    m.b_iter_bytes = a.tpg * m.b_tile_bytes;
    m.k_iters = a.num_groups * a.k_chunks;
    m.stages = a.stages;
    m.stage_bytes = a.stage_bytes;
    m.a_bytes = a.a_bytes;
    m.b_resident = a.b_resident;
    m.bres_off = a.stages * a.stage_bytes;
    m.pad = 0;
  }
  DYNMM_CUDA(cudaMemsetAsync(barrier, 0, 8, stream));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kProgThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  // cooperative: every CTA must be resident, they wait for each other at the phase barriers
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DYNMM_CUDA(cudaLaunchKernelEx(&cfg, conv_program_kernel, params, static_cast<const uint8_t*>(image_dev),
                                static_cast<unsigned*>(barrier), static_cast<unsigned long long*>(trace)));
  return DYNMM_OK;
}
