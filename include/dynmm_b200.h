/*
 * dynmm_b200 -- C ABI of the B200 (sm_100a) kernels behind the DynMM gated hot path.
 *
 * The reference (zihuixue/DynMM) has no native/FFI layer: every op on its hot
 * path is an ATen call made from Python.  The entry
 * points below are therefore the operator boundary a maintainer would bind
 * with ctypes from the reference's own modules; each one cites the reference
 * code it replaces (paths relative to /root/reference).  INTEGRATION.md shows
 * the binding stubs.
 *
 * Conventions
 *   - plain C: raw device pointers, ints, floats; no torch types.
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*),
 *     never synchronises the host, never allocates, and is CUDA-graph
 *     capturable.  Data-dependent work (gate decisions) stays on the device:
 *     kernels read sample counts / slot maps from device memory.
 *   - returns 0 on success, a negative DYNMM_E* code otherwise;
 *     dynmm_last_error() gives a thread-local message.
 *   - activations are NHWC ("channels last"); `ld` arguments are the pixel
 *     pitch in elements (>= channels) so tensors can be channel slices of a
 *     wider buffer.  bf16 unless stated.
 */
#ifndef DYNMM_B200_H_
#define DYNMM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNMM_ABI_VERSION 1

enum {
  DYNMM_OK = 0,
  DYNMM_EINVAL = -1,     /* bad argument / unsupported shape */
  DYNMM_ECUDA = -2,      /* CUDA runtime / driver error */
  DYNMM_EUNSUPPORTED = -3, /* valid arguments, but this entry point cannot run them (the caller has an alternative) */
  DYNMM_ENODEV = -3      /* no sm_100 device */
};

int dynmm_abi_version(void);
const char* dynmm_last_error(void);
/* 1 if the current device is compute capability 10.x */
int dynmm_device_ok(void);

/* ------------------------------------------------------------------ gate */

/* DiffSoftmax forward (model_skip_mod_globalgate.py:20-30, imdb_dyn.py:16-26,
 * affect_dyn.py:18-28).  logits [rows, n] fp32 (n <= 32).  y = softmax(logits/tau);
 * hard: one-hot of the FIRST maximum of y (the value the straight-through
 * expression y_hard - y.detach() + y evaluates to).  y_soft (may be NULL) receives
 * the soft probabilities needed by the backward; index (may be NULL) the argmax. */
int dynmm_diffsoftmax_fwd(const float* logits, int rows, int n, float tau, int hard,
                          float* y, float* y_soft, int32_t* index, void* stream);
/* Backward of the above: hard or soft, the Jacobian is that of the tempered
 * softmax (straight-through).  grad_logits = (y_soft * (g - sum(g*y_soft))) / tau. */
int dynmm_diffsoftmax_bwd(const float* grad_y, const float* y_soft, int rows, int n, float tau,
                          float* grad_logits, void* stream);

/* Turn gate weights [b,5] (fp32, one-hot or soft) into what the encoder kernels
 * consume (model_skip_mod_globalgate.py:282,291,300,309):
 *   g[s*b + i]      = depth mixing coefficient of sample i at stage s+1 (s=0..3):
 *                     1-w0, 1-w0-w1, 1-w0-w1-w2, w4
 *   perm[i]         = sample index held by depth slot i (samples sorted by
 *                     decreasing number of depth stages they need; stable)
 *   slot[i]         = inverse of perm
 *   count[s]        = number of samples with g[s] != 0 (a prefix of perm)
 *   hist[k]        += number of samples whose arg-max branch is k (int64[5]);
 *                     replaces the per-forward weight.cpu() of :273-274. */
int dynmm_gate_plan(const float* weight, int b, float* g, int32_t* perm, int32_t* slot,
                    int32_t* count, long long* hist, void* stream);

/* GlobalGate.forward (model_skip_mod_globalgate.py:388-394) in fp32:
 *   rgb, depth : NHWC fp32 [b,h,w,64] pooled stem features
 *   w1 [8][5][5][128] (kh,kw,cin order, cin = 64 rgb then 64 depth), scale1/shift1 [8] = conv bias + BN folded
 *   w2 [8][5][5][8], scale2/shift2 [8]; wfc [5][8]
 *   work: >= dynmm_global_gate_workspace(b,h,w) bytes
 *   logits [b,5] out.  DiffSoftmax is a separate call. */
long long dynmm_global_gate_workspace(int b, int h, int w);
int dynmm_global_gate_logits(const float* rgb, const float* depth, int b, int h, int w,
                             const float* w1, const float* scale1, const float* shift1,
                             const float* w2, const float* scale2, const float* shift2,
                             const float* wfc, void* work, float* logits, void* stream);

/* The whole gate of an eval forward in three launches: the two convolutions of dynmm_global_gate_logits, then ONE
 * kernel for GAP finish + fc + DiffSoftmax (tau, hard) + the plan of dynmm_gate_plan -- the arithmetic and its order
 * are those of the separate calls (same logits, same decisions), two launches fewer on the path in front of the depth
 * encoder.  logits / weight [b,5]; g, perm, slot, count, hist as in dynmm_gate_plan. */
int dynmm_global_gate_decide(const float* rgb, const float* depth, int b, int h, int w,
                             const float* w1, const float* scale1, const float* shift1,
                             const float* w2, const float* scale2, const float* shift2,
                             const float* wfc, void* work, float tau, int hard, float* logits, float* weight,
                             float* g, int32_t* perm, int32_t* slot, int32_t* count, long long* hist, void* stream);

/* ------------------------------------------------------------------ stem */

/* ResNet.forward_first_conv x2 + add + max_pool2d x2
 * (resnet.py:352-358, model_skip_mod_globalgate.py:256-261), fp32 arithmetic:
 *   rgb [b,3,h,w], depth [b,1,h,w] NCHW fp32 (the module's input layout)
 *   w_rgb [7][7][3][64], w_d [7][7][1][64] fp32; scale/shift [64] = folded BN
 *   outputs, pooled to (h/4, w/4), NHWC, 64 channels:
 *     rgb_f32/depth_f32 : fp32 copies for the gate (NULL to skip)
 *     rgb_bf16/depth_bf16 : bf16 copies for the encoders
 *   rgb_* holds maxpool(relu(bn(conv(rgb))) + relu(bn(conv(depth)))), depth_* holds
 *   maxpool(relu(bn(conv(depth)))). */
int dynmm_stem_fwd(const float* rgb, const float* depth, int b, int h, int w,
                   const float* w_rgb, const float* scale_rgb, const float* shift_rgb,
                   const float* w_d, const float* scale_d, const float* shift_d,
                   float* rgb_f32, float* depth_f32, void* rgb_bf16, void* depth_bf16,
                   const float* se_rgb, const float* se_depth, float* gap_partial, void* stream);
/* SE-add fusion at stage 0 (se_layer0; rgb_depth_fusion.py:22-26) is a two-pass use of the call above:
 *   pass 1: gap_partial != NULL (float[dynmm_stem_gap_tiles() * 128]): only per-tile channel sums of the two
 *           UNFUSED stem maps are produced (no pooled outputs); finish them with dynmm_se_mlp
 *           (chunks = tiles per sample, inv_area = 1 / (h/2 * w/2), rows of 128 = [rgb 64 | depth 64]);
 *   pass 2: se_rgb / se_depth (float [b][64]) scale the two streams before the add. */
long long dynmm_stem_gap_tiles(int b, int h, int w);

/* The same fused stem (plain `add` fusion: no SE scales, no squeeze pass) with the im2col done by TMA:
 * a pre-pass rewrites the 4-channel fp32 NCHW input as a zero-bordered space-to-depth image (bf16 hi + lo
 * planes, 16 channels = 2x2 positions x (3 RGB + 1 depth)) in `workspace`; the 7x7/s2 convolutions of both
 * encoders (resnet.py:352-358) then are one 4x4 unit-stride convolution = one tcgen05 GEMM with N = 128
 * whose A rows TMA gathers as overlapping 128-byte windows.  Same arithmetic class as dynmm_stem_fwd's
 * tensor-core path (three bf16 split products, fp32 accumulation).
 *   w_packed: bf16 [2 (hi, lo)][128][256] from dynmm_stem_s2d_pack_weights (inputs: the [7][7][cin][64] fp32
 *   weights dynmm_stem_fwd takes).  workspace: dynmm_stem_s2d_workspace(b, h, w) bytes of device scratch.
 *   split != 0: rgb_bf16 / depth_bf16 are [b][hp][wp][128] = [hi | lo] halves of the fp32 values (what
 *   dynmm_split_from_f32 makes of rgb_f32 / depth_f32), the input format of DYNMM_CONV_SPLIT convolutions.
 *   bn_host (optional, HOST memory, 256 floats [scale_rgb | shift_rgb | scale_d | shift_d]): a host copy of the four
 *   device vectors; when given it travels as a kernel parameter and the epilogue reads it from the constant bank
 *   instead of shared memory (same values, same arithmetic; the device vectors are then not read). */
long long dynmm_stem_s2d_workspace(int b, int h, int w);
int dynmm_stem_s2d_pack_weights(const float* w_rgb, const float* w_d, void* w_packed, void* stream);
int dynmm_stem_s2d_fwd(const float* rgb, const float* depth, int b, int h, int w, const void* w_packed,
                       const float* scale_rgb, const float* shift_rgb, const float* scale_d, const float* shift_d,
                       void* workspace, long long workspace_bytes, float* rgb_f32, float* depth_f32,
                       void* rgb_bf16, void* depth_bf16, int split, const float* bn_host, void* stream);

/* ------------------------------------------------------ SE-add fusion */

/* Squeeze: partial[n][64 chunks][c] channel sums of NHWC bf16 x [n,hw,ld] (model_utils.py:48).  `count`
 * (device int32 or NULL): only the first *count samples are reduced. */
long long dynmm_gap_workspace(int n, int c);
int dynmm_gap_partial(const void* x, int n, long long hw, int c, int ld, const int32_t* count, float* partial,
                      void* stream);
/* ... of [hi | lo] tensors (DYNMM_CONV_SPLIT layout: pitch ld >= 2c, lo half at channel ld / 2): sums of hi + lo */
int dynmm_gap_partial_split(const void* x, int n, long long hw, int c, int ld, const int32_t* count, float* partial,
                            void* stream);
/* Excite: sigma[row] = sigmoid(W2 relu(W1 mean + b1) + b2) with mean = inv_area * sum of `chunks` partial rows
 * (model_utils.py:40-45,49); partial rows have pitch ld and the c channels start at column c_off.
 * w1 [hidden][c], w2 [c][hidden] (the 1x1 conv weights). */
int dynmm_se_mlp(const float* partial, int rows, int chunks, int ld, int c_off, float inv_area, int c, int hidden,
                 const float* w1, const float* b1, const float* w2, const float* b2, const int32_t* count,
                 float* sigma, void* stream);
/* Gated SE fusion (model_skip_mod_globalgate.py:280-283 with se_layer{s}):
 *   out[n] = rgb[n]*(1 - g + g*sig_r[n]) + g*sig_d[slot(n)]*depth[slot(n)],  g = gate[n];
 * samples with g == 0 never read depth.  rgb/depth NHWC bf16 [*,hw,c], out pitch out_ld. */
int dynmm_se_gated_fuse(const void* rgb, const void* depth, const float* sig_r, const float* sig_d,
                        const float* gate, const int32_t* slot, int n, long long hw, int c, int out_ld, void* out,
                        void* stream);
/* ... on [hi | lo] tensors: rgb / depth [*,hw,2c], out pitch out_ld >= 2c with its lo half at channel out_ld / 2;
 * the blend runs in fp32 on hi + lo and is split again (the fp32-grade engine mode with 'SE-add' fusion). */
int dynmm_se_gated_fuse_split(const void* rgb, const void* depth, const float* sig_r, const float* sig_d,
                              const float* gate, const int32_t* slot, int n, long long hw, int c, int out_ld,
                              void* out, void* stream);

/* --------------------------------------------------------- encoder convs */

/* Implicit-GEMM convolution on tcgen05 tensor cores (TMA-fed, TMEM
 * accumulators) with the whole post-conv chain of the reference fused into the
 * epilogue.  Replaces F.conv2d + BatchNorm2d(eval) + ReLU + residual add of
 * NonBottleneck1D.forward / BasicBlock.forward / Bottleneck.forward /
 * ConvBNAct (resnet.py:66-84,124-147,173-192, model_utils.py:11-23) and the
 * gated fusion `w*rgb + (1-w)*(rgb+depth)` (model_skip_mod_globalgate.py:279-310):
 *
 *   v   = conv(in)[n,h,w,c] * scale[c] + shift[c]
 *   v  += residual[res_map(n),h,w,c]              (if residual)
 *   v   = max(v, 0)                               (if relu)
 *   v  += gate[n] * gated[slot(n),h,w,c]          (if gated and gate[n] != 0;
 *                                                  gated-off samples never touch `gated`)
 *   out[n,h,w,c] = bf16(v)
 *
 * Sample indirection (real skipping of gated-off depth stages): `count` (device
 * int32, may be NULL = n) is the number of leading sample slots that are
 * computed at all; `in_map` (device int32[n], may be NULL) gives the input
 * sample read for output slot i. */

/* Tile-completion flags of a convolution's OUTPUT tensor (layer-to-layer overlap without a kernel boundary).
 * A launch with `out_flags` set adds 1 to flags[tile] (release, gpu scope) once every byte of output pixel tile
 * `tile` x one channel tile is in memory; the tensor's pixel tiles form a grid of box_n x box_h x box_w pixels,
 *   tile = ((n / box_n) * tiles_h + h / box_h) * tiles_w + w / box_w,
 * and a tile is complete when its flag reaches `need` (the producer's number of channel tiles).  A consumer launch
 * that gets this description as `in_flags` / `res_flags` does NOT wait for the previous kernel of the stream as a
 * whole (no griddepcontrol.wait): its loader waits, per work unit, for exactly the producer tiles the unit's input
 * window (and residual tile) overlaps -- so the CTAs of layer k+1 start on the SMs that layer k's early finishers
 * free, run their prologue and every tile whose neighbourhood is complete while layer k's last tiles are still in
 * flight.  flags must be zero before the producer starts (one memset per forward); dynmm_conv_tile_grid() fills the
 * geometry for a given launch.  flags == NULL: not used. */
typedef struct dynmm_tile_flags {
  int32_t* flags;
  int32_t box_n, box_h, box_w;
  int32_t tiles_h, tiles_w;
  int32_t need;
} dynmm_tile_flags;

typedef struct dynmm_conv_params {
  const void* in;         /* bf16 NHWC [n_in, h_in, w_in, in_ld] */
  const void* weight;     /* bf16 [kh*kw][c_out_pad][c_in], c_out_pad = c_out rounded up to 16 */
  const float* scale;     /* [c_out] or NULL (1) */
  const float* shift;     /* [c_out] or NULL (0) */
  const void* residual;   /* bf16 NHWC [*, h_out, w_out, res_ld] or NULL */
  const int32_t* res_map; /* device int32 [n]: sample of `residual` added to output slot i, or NULL (identity) */
  void* out;              /* bf16 NHWC [n, h_out, w_out, out_ld] */
  const void* gated;      /* bf16 NHWC [*, h_out, w_out, gated_ld] or NULL */
  const float* gate;      /* device fp32 [n] */
  const int32_t* gated_slot; /* device int32 [n] or NULL (identity) */
  const int32_t* in_map;  /* device int32 [n] or NULL */
  const int32_t* count;   /* device int32 or NULL */
  int32_t n, n_in;        /* output sample slots; samples in `in` */
  int32_t h_in, w_in, c_in, in_ld;
  int32_t h_out, w_out, c_out, out_ld;
  int32_t res_ld, gated_ld;
  int32_t kh, kw, stride_h, stride_w, pad_h, pad_w;
  int32_t relu;           /* activation after the residual add: 0 none, 1 ReLU, 2 swish x*sigmoid(x), 3 h-swish x*relu6(x+3)/6
                             (model_utils.py:100-115) */
  int32_t tile_n;         /* 0 = choose; else 16..256 output channels per CTA tile */
  int32_t max_ctas;       /* 0 = one per SM */
  int32_t flags;          /* DYNMM_CONV_* bits (fills the padding in front of `trace`: the layout is unchanged) */
  void* trace;            /* debug: device uint64[16 * ctas] in-kernel cycle stamps, or NULL */
  dynmm_tile_flags in_flags;   /* completion flags of `in`'s producer (flags == NULL: ordinary stream order) */
  dynmm_tile_flags res_flags;  /* completion flags of `residual`'s producer; required with in_flags when the
                                  residual was written by a launch that may still be running */
  dynmm_tile_flags out_flags;  /* flags this launch publishes (geometry from dynmm_conv_tile_grid), or NULL */
} dynmm_conv_params;

/* dynmm_conv_params.flags: `weight` / `scale` / `shift` were WRITTEN by earlier work on the same stream (training:
 * the bf16 weights are re-packed every optimizer step right before the convolution).  The kernel normally fetches
 * them in its prologue, before the programmatic-dependent-launch wait, because for inference they are constants;
 * with this flag the launch is ordinary stream-ordered (no programmatic early start), so the prologue cannot run
 * ahead of the producer. */
#define DYNMM_CONV_VOLATILE_WEIGHTS 1
/* Work-unit shape of streamed-weight layers (C >= 256): by default the planner decides whether two pixel tiles share
 * every weight tile ("dual-M" units: half the L2->SM weight traffic, half the CTAs); these bits force the choice
 * (tests, experiments).  Results are bit-identical either way. */
#define DYNMM_CONV_NO_DUAL 2
#define DYNMM_CONV_FORCE_DUAL 4
/* with in_flags: `residual` was complete before the chain of flagged launches began (no res_flags needed) */
#define DYNMM_CONV_RESIDUAL_SETTLED 8
/* `count` was written by a kernel that completed before the PREVIOUS kernel of this stream started (every depth-encoder
 * launch but the first after dynmm_gate_plan): the kernel reads it while it waits for the previous kernel */
#define DYNMM_CONV_COUNT_SETTLED 16
/* fp32-grade arithmetic on the bf16 tensor cores ("f32x3"): every activation tensor of the launch (in, out, residual,
 * gated) holds fp32-grade values as TWO bf16 halves, hi = bf16(x) in channels [0, ld/2) and lo = bf16(x - hi) in
 * [ld/2, ld) of its channel pitch ld; `weight` is packed [taps][c_out_pad][3 * c_in] = [W_hi | W_lo | W_hi]
 * (dynmm_fold_pack_conv_split) and the contraction runs over the three products x_hi*W_hi + x_hi*W_lo + x_lo*W_hi with
 * fp32 accumulation (the dropped x_lo*W_lo term is 2^-16 relative).  c_in / c_out stay the logical channel counts
 * (c_in a multiple of 64); the epilogue works on the reconstructed fp32 values and writes hi / lo again. */
#define DYNMM_CONV_SPLIT 32

int dynmm_conv_igemm_fwd(const dynmm_conv_params* p, void* stream);
/* dynmm_fold_pack_conv with the folded fp32 weights split into bf16 halves for DYNMM_CONV_SPLIT launches:
 * packed [kh*kw][c_out_pad16][3 * c_in] = [hi | lo | hi] along the last axis. */
int dynmm_fold_pack_conv_split(const float* w, int c_out, int c_in, int kh, int kw, const float* bias,
                               const float* bn_weight, const float* bn_bias, const float* bn_mean, const float* bn_var,
                               float eps, void* packed, float* shift, void* stream);
/* Two convolutions of IDENTICAL geometry in ONE launch: the same layer of the RGB and of the depth encoder
 * (FusionDynMM/src/models/model_skip_mod_globalgate.py:276-310 runs the two ResNets in lock step).  Each job keeps its
 * own tensors, sample count (`count`) and epilogue operands; the CTAs walk one combined tile list (job a's tiles,
 * then job b's), so at batch 8 a CTA runs an RGB tile and a depth tile back to back and the fixed cost of a launch is
 * paid once per layer instead of once per layer and encoder.  Weights are streamed (two resident sets do not fit).
 * Results are bit-identical to two dynmm_conv_igemm_fwd calls.  Returns DYNMM_EUNSUPPORTED when the two convolutions
 * do not plan to the same tiling or use single-launch options (trace, max_ctas, tile flags): launch them separately. */
int dynmm_conv_igemm_fwd2(const dynmm_conv_params* a, const dynmm_conv_params* b, void* stream);
/* Geometry of the flags this launch would publish (host only; grid->flags is left untouched): the caller needs
 * grid->tiles_h * grid->tiles_w * ceil(n / grid->box_n) zeroed int32 flags.  Returns DYNMM_EINVAL for launches that
 * cannot publish flags. */
int dynmm_conv_tile_grid(const dynmm_conv_params* p, dynmm_tile_flags* grid);
/* ------------------------------------------------- convolution programs
 *
 * Many dependent convolutions in ONE persistent cooperative launch.  At batch 8 a layer of the gated
 * encoders (resnet.py:124-147 run twice in lock step, model_skip_mod_globalgate.py:276-310) is a few
 * microseconds of work; the launch boundary between two layers costs as much again.  A program is a list
 * of jobs (dynmm_conv_params, exactly as for dynmm_conv_igemm_fwd) grouped into phases: jobs of one phase
 * are independent (at most 4), every job of phase k+1 may read what phase k wrote.  The CTAs of a phase
 * are split between its jobs on the device, in proportion to the tiles that the jobs' `count` leaves.
 *
 *   bytes  = dynmm_conv_program_bytes(n_jobs)
 *   dynmm_conv_program_build(jobs, phase_of_job, n_jobs, host_image, bytes, cfg)   host only, no CUDA work
 *   copy host_image -> device (128-byte aligned), then per forward:
 *   dynmm_conv_program_launch(device_image, host_image, barrier, trace, stream)
 *
 * phase_of_job is non-decreasing, starts at 0 and has no gaps.  The image embeds the tensor addresses of the
 * jobs: it stays valid as long as those buffers do.  `barrier` is 8 bytes of device scratch owned by this
 * launch (zeroed by it).  cfg = {grid, dynamic shared memory, phases} (informational); the launch reads its
 * configuration and the per-job MMA parameters (passed as kernel parameters) from the HOST image.  `trace` (debug, may be NULL): device
 * uint64 [grid][97], receives the %globaltimer value at which each CTA was released into each phase and, in the
 * last used column, at which it finished.  Results are bit-identical to launching the jobs one by one. */
long long dynmm_conv_program_bytes(int n_jobs);
int dynmm_conv_program_build(const dynmm_conv_params* jobs, const int32_t* phase_of_job, int n_jobs,
                             void* host_image, long long image_bytes, int32_t* cfg /* [3] */);
int dynmm_conv_program_launch(const void* device_image, const void* host_image, void* barrier, void* trace,
                              void* stream);

/* Fused pair of a 64-channel NonBottleneck1D block (resnet.py:124-147): conv3x1(+bias)+ReLU followed by
 * conv1x3(+folded BN shift)(+residual)(+ReLU) in ONE kernel; the intermediate stays in shared memory.  Same
 * arithmetic and rounding as two dynmm_conv_igemm_fwd calls (bit-identical).  64 -> 64 -> 64 channels, stride 1.
 * w1 / w2: packed bf16 [3][64][64] as for dynmm_conv_igemm_fwd.  in_map / res_map / count as in dynmm_conv_params. */
typedef struct dynmm_conv_pair_params {
  const void* in;         /* bf16 NHWC [n_in, h, w, in_ld] */
  const void* w1;         /* 3x1 conv */
  const float* shift1;    /* [64] or NULL */
  const void* w2;         /* 1x3 conv */
  const float* shift2;    /* [64] or NULL */
  const void* residual;   /* bf16 NHWC [*, h, w, res_ld] or NULL */
  void* out;              /* bf16 NHWC [n, h, w, out_ld] */
  const int32_t* count;
  const int32_t* in_map;
  const int32_t* res_map;
  int32_t n, n_in, h, w;
  int32_t in_ld, out_ld, res_ld;
  int32_t relu2;
} dynmm_conv_pair_params;
int dynmm_conv_pair_fwd(const dynmm_conv_pair_params* p, void* stream);

/* ------------------------------------------------- convolution chains
 *
 * A run of NonBottleneck1D convolutions (resnet.py:124-147: conv3x1 -> ReLU -> conv1x3 -> BN -> ReLU -> conv3x1 ->
 * ReLU -> conv1x3 -> BN -> +identity -> ReLU, block after block) of ONE encoder stage or decoder module as ONE kernel:
 * c -> c channels (c = 128 or 256), stride 1, 3 taps along H or along W per layer.  A CTA owns a strip of rows of one
 * sample and keeps its activations in shared memory from the first layer to the last; the only per-layer traffic is
 * the streamed weights, the rows a 3x1 layer needs from the strips above and below (exchanged through a global
 * scratch buffer behind per-strip flags -- no grid-wide barrier), and the layer outputs the caller asked for.
 * Arithmetic, accumulation order and rounding are those of dynmm_conv_igemm_fwd layer by layer: the results are
 * bit-identical to the per-layer launches.
 *
 * Up to two jobs of identical geometry (the same stage of the RGB and of the depth encoder,
 * model_skip_mod_globalgate.py:276-310) share a launch; each has its own layers, tensors and sample count.
 *
 *   per layer list (once):   bytes = dynmm_conv_chain_image_bytes(n_layers)
 *                            dynmm_conv_chain_build(layers, n_layers, c, host_image)      host only
 *                            copy host_image -> device (128-byte aligned)
 *   per geometry:            dynmm_conv_chain_plan(h, w, c, total sample slots, &units, &scratch_bytes)
 *   per forward:             dynmm_conv_chain_fwd(&params, stream)
 *
 * `flags`: units + total sample slots int32, ZERO before the first launch that uses them (the kernel leaves them
 * zero again); `scratch`: scratch_bytes of device memory owned by the launch. */
typedef struct dynmm_chain_layer {
  const void* weight;     /* packed bf16 [3][c][c] as for dynmm_conv_igemm_fwd */
  const float* shift;     /* [c] or NULL */
  int32_t taps_h;         /* 1: 3x1 (taps along H), 0: 1x3 (taps along W) */
  int32_t relu;
  int32_t residual;       /* 0: none, 1: += in[n,h,w,:] (the chain input), 2: += out[n,h,w,:] as stored earlier */
  int32_t store;          /* 0: stays in shared memory, 1: also written to `out`, 2: also written to `out_last` */
} dynmm_chain_layer;

typedef struct dynmm_chain_job {
  const void* image;      /* device image of this job's layers (dynmm_conv_chain_build) */
  const void* in;         /* bf16 NHWC [n, h, w, c] */
  void* out;              /* bf16 NHWC [n, h, w, c] (store == 1 layers) or NULL */
  void* out_last;         /* bf16 NHWC [n, h, w, c] (store == 2 layers) or NULL */
  const int32_t* count;   /* device int32 (leading sample slots that are computed) or NULL = n */
  int32_t n;
  int32_t n_layers;
  int32_t count_settled;  /* see DYNMM_CONV_COUNT_SETTLED */
  int32_t reserved;
} dynmm_chain_job;

typedef struct dynmm_chain_params {
  dynmm_chain_job jobs[2];
  int32_t n_jobs;         /* 1 or 2 */
  int32_t h, w, c;
  int32_t* flags;
  void* scratch;
  long long scratch_bytes;
  void* trace;            /* debug: device uint64[16 * units] cycle stamps, or NULL */
} dynmm_chain_params;

long long dynmm_conv_chain_image_bytes(int n_layers);
int dynmm_conv_chain_build(const dynmm_chain_layer* layers, int n_layers, int c, void* host_image);
/* DYNMM_EUNSUPPORTED when the geometry does not fit (channels, shared memory, more units than the GPU holds at once):
 * launch the layers one by one instead. */
int dynmm_conv_chain_plan(int h, int w, int c, int total_slots, int32_t* units, long long* scratch_bytes);
int dynmm_conv_chain_fwd(const dynmm_chain_params* p, void* stream);

/* Same contract, one thread per output element, CUDA cores.  Test comparator
 * for the tensor-core kernel at sizes the CPU oracle cannot reach; never used
 * by the product path. */
int dynmm_conv_direct_fwd(const dynmm_conv_params* p, void* stream);

/* ------------------------------------------------- convolution backward
 *
 * the reference has no explicit backward code; autograd
 * (train.py:323) differentiates every F.conv2d of the encoder/decoder blocks
 * (resnet.py:124-147, 66-84, 173-192; model_utils.py:11-23).
 *
 * Data gradient: a convolution again -- dynmm_conv_igemm_fwd over dy with the
 * taps mirrored and the channel roles swapped (weight packed as
 * [kh*kw][c_in_pad][c_out]); stride-2 layers run it on the zero-interleaved dy.
 *
 * Weight gradient: dw[co][ci][ky][kx] = sum over (n, ho, wo) of
 *   dy[n, ho, wo, co] * x[n, ho*stride_h + ky - pad_h, wo*stride_w + kx - pad_w, ci]
 * as a tcgen05 GEMM whose K dimension is the pixel index: both operands are the
 * NHWC tensors themselves (MN-major shared-memory tiles), shifted per tap by the
 * TMA coordinates, out-of-image pixels zero-filled.  The pixel range is split over
 * CTAs; the fp32 partial sums go to `workspace` and are reduced in a fixed order
 * (deterministic), written in the framework's weight layout [c_out][c_in][kh][kw]. */
typedef struct dynmm_wgrad_params {
  const void* x;          /* bf16 NHWC [n, h_in, w_in, x_ld] */
  const void* dy;         /* bf16 NHWC [n, h_out, w_out, dy_ld] */
  float* dw;              /* fp32 [c_out][c_in][kh][kw] */
  void* workspace;        /* device scratch of dynmm_conv_wgrad_workspace() bytes */
  long long workspace_bytes;
  int32_t n, h_in, w_in, c_in, x_ld;
  int32_t h_out, w_out, c_out, dy_ld;
  int32_t kh, kw, stride_h, stride_w, pad_h, pad_w;
  int32_t accumulate;     /* 0: dw = result, 1: dw += result */
  int32_t max_ctas;       /* 0 = one per SM */
} dynmm_wgrad_params;

/* bytes of scratch the call needs for these shapes (<0: invalid arguments) */
long long dynmm_conv_wgrad_workspace(const dynmm_wgrad_params* p);
int dynmm_conv_wgrad(const dynmm_wgrad_params* p, void* stream);
/* CUDA-core comparator with the same contract (tests only; no workspace needed). */
int dynmm_conv_wgrad_direct(const dynmm_wgrad_params* p, void* stream);

/* Engine build (FusionEngine.__init__), one launch per convolution: eval-mode BatchNorm (and an optional conv bias)
 * folded into the weights, then packed to the bf16 operand layout of dynmm_conv_params.weight:
 *   scale = bn_weight / sqrt(bn_var + eps);  shift = bn_bias - bn_mean * scale (+ bias * scale)
 *   packed [kh*kw][c_out_pad16][c_in] = bf16(w * scale)            (padding rows zero)
 * bn_* all NULL: no BatchNorm (packed = bf16(w), shift = bias).  shift may be NULL only without bias and BatchNorm.
 * Every operation is rounded separately -- bit-identical to the fp32 PyTorch expression (conv + BatchNorm2d of
 * FusionDynMM/src/models/resnet.py:124-147, model_utils.py:11-23 in eval mode). */
int dynmm_fold_pack_conv(const float* w, int c_out, int c_in, int kh, int kw, const float* bias, const float* bn_weight,
                         const float* bn_bias, const float* bn_mean, const float* bn_var, float eps, void* packed,
                         float* shift, void* stream);
/* scale / shift [c] of an eval-mode BatchNorm (+ preceding conv bias), as above (stem, gate: fp32 epilogues). */
int dynmm_fold_bn(int c, const float* bias, const float* bn_weight, const float* bn_bias, const float* bn_mean,
                  const float* bn_var, float eps, float* scale, float* shift, void* stream);
/* out = in.permute(p0, p1, p2).contiguous() of an fp32 tensor [d0][d1][d2] (weight layout changes of the engine build). */
int dynmm_permute3d_f32(const float* in, int d0, int d1, int d2, int p0, int p1, int p2, float* out, void* stream);

/* fp32 master weight [c_out][c_in][kh][kw] -> the bf16 operand layouts, in one pass (either may be NULL):
 *   fwd   [kh*kw][c_out_pad16][c_in]   (dynmm_conv_params.weight of the forward convolution)
 *   dgrad [kh*kw][c_in_pad16][c_out]   taps mirrored: the weight of the data-gradient convolution */
int dynmm_pack_conv_weight(const float* w, int c_out, int c_in, int kh, int kw, void* fwd, void* dgrad,
                           void* stream);
/* Bias gradient: out[ch] (=|+=) sum over rows of x[row][ch], x bf16 [rows][ld] (NHWC flattened), fp32
 * accumulation in a fixed order (deterministic).  workspace: dynmm_channel_sum_workspace(rows, c) bytes. */
long long dynmm_channel_sum_workspace(long long rows, int c);
int dynmm_channel_sum(const void* x, long long rows, int c, int ld, float* out, void* workspace,
                      long long workspace_bytes, int accumulate, void* stream);

/* --------------------------------------------------------- elementwise */

/* out = a + gate[n]*b[slot(n)]  (NHWC bf16; model_skip_mod_globalgate.py:283 etc. as a
 * stand-alone op; b is only read where gate[n] != 0). */
int dynmm_gated_add_fwd(const void* a, const void* b, const float* gate, const int32_t* slot,
                        int n, long long per_sample, void* out, void* stream);
/* fp32 training-path variant with both gradients:
 * fuse = a + g[n]*b; grad_g[n] = sum(grad*b). */
int dynmm_gated_add_f32_fwd(const float* a, const float* b, const float* gate, int n,
                            long long per_sample, float* out, void* stream);
int dynmm_gated_add_f32_bwd(const float* grad, const float* b, const float* gate, int n,
                            long long per_sample, float* grad_b, float* grad_gate, void* stream);

/* Modality-level mix (imdb_dyn.py:100, affect_dyn.py:95,164):
 * out[i,:] = sum_e w[i,e]*pred_e[row_e(i),:], e < n_experts <= 4.  row maps
 * (int32 [b] per expert, may be NULL = identity) let an expert be evaluated on a
 * compacted subset only; experts with w[i,e]==0 are never read. */
int dynmm_softgate_mix_fwd(const float* const* preds, const int32_t* const* rows, const float* w,
                           int b, int c, int n_experts, float* out, void* stream);
int dynmm_softgate_mix_bwd(const float* grad_out, const float* const* preds, const float* w,
                           int b, int c, int n_experts, float* const* grad_preds, float* grad_w,
                           void* stream);
/* Stable compaction for expert skipping: rows with w[i,expert] != 0 first.
 * idx [b] receives the selected row ids (prefix of length *count), inv [b] their
 * position or -1. */
int dynmm_compact_rows(const float* w, int b, int n_experts, int expert, int32_t* idx,
                       int32_t* inv, int32_t* count, void* stream);

/* -------------------------------------------------- eval post-processing */

/* eval.py:120-141 + confusion_matrix.py:122-133 fused: pred = argmax_c logits (first maximum), void pixels
 * (label_orig == 0) ignored, label -= 1, cm[label*c + pred] += 1.  logits NCHW fp32 [n,c,h,w] at label
 * resolution; label_orig uint8 [n,h,w]; cm int64 [c*c] is ACCUMULATED; pred_out uint8 [n,h,w] optional.
 * Pass label_orig = cm = NULL for a pure arg-max. */
int dynmm_argmax_confusion(const float* logits, const uint8_t* label_orig, int n, int c, int h, int w,
                           long long* cm, uint8_t* pred_out, void* stream);
/* iou_pytorch / miou_pytorch (confusion_matrix.py:139-178): iou[k] = diag/(row+col-diag+1e-15) in fp64,
 * miou = mean(iou).  iou (double[c]) optional. */
int dynmm_miou(const long long* cm, int c, double* iou, double* miou, void* stream);

/* ------------------------------------------------------ training loss
 *
 * One scale of CrossEntropyLoss2d (FusionDynMM/src/utils.py:34-50): logits fp32 NCHW [n,c,h,w], targets int32 [n,h,w]
 * with 0 = void and 1..c = class + 1, weight [c].  loss = sum w[t] (logsumexp(x) - x[t]) / sum_c n_c w[c];
 * `lse` [n,h,w] and `divisor` (device scalars / buffers) are kept for the backward, which writes
 * grad_logits = grad_out * w[t] (softmax(x) - onehot(t)) / divisor.  workspace: dynmm_ce2d_workspace() bytes.
 * Used by dynmm_b200.fusion.CrossEntropyLoss2d for CUDA logits (DYNMM_CE_CUDA=0 selects the PyTorch statement). */
long long dynmm_ce2d_workspace(int n, int c, int h, int w);
int dynmm_ce2d_fwd(const float* logits, const int32_t* targets, const float* weight, int n, int c, int h, int w,
                   void* workspace, long long workspace_bytes, float* lse, float* loss, float* divisor, void* stream);
int dynmm_ce2d_bwd(const float* logits, const int32_t* targets, const float* weight, const float* lse,
                   const float* divisor, const float* grad_out, int n, int c, int h, int w, float* grad_logits,
                   void* stream);

/* layout / dtype plumbing */
int dynmm_nchw_f32_to_nhwc_bf16(const float* in, int n, int c, int h, int w, void* out, void* stream);
int dynmm_nhwc_bf16_to_nchw_f32(const void* in, int n, int c, int h, int w, int ld, float* out, void* stream);

/* Upsample.forward for mode 'learned-3x3-zeropad' (model.py:403-410): nearest x2 then
 * depthwise 3x3 (zero pad) + bias, optionally + skip (DecoderModule.forward :353-355).
 * in NHWC bf16 [n,h,w,c]; weight fp32 TAP-MAJOR [3*3][c] (= conv.weight[c,1,3,3] transposed); out NHWC bf16
 * [n,2h,2w,c]; or (final upsampling) out_nchw_f32 fp32 NCHW -- the module's return layout -- and/or
 * labels uint8 [n,2h,2w] = argmax over channels of those logits (first maximum; eval.py:120), produced in the
 * same pass so an evaluation loop never has to write or re-read the logits. */
int dynmm_upsample2x_dw3x3(const void* in, int n, int h, int w, int c, const float* weight,
                           const float* bias, const void* skip, void* out_nhwc_bf16,
                           float* out_nchw_f32, uint8_t* labels, void* stream);

/* fp32-grade ("f32x3") variants of the decoder / context helpers for DYNMM_CONV_SPLIT tensors ([hi | lo] bf16 halves,
 * hi in channels [0, ld/2), lo in [ld/2, ld)): both halves are read, the arithmetic is fp32, both halves are written.
 * `c` is the logical channel count.  dynmm_split_from_f32 turns an fp32 NHWC map (the stem's gate-path copies) into
 * the split form. */
int dynmm_split_from_f32(const float* x, long long rows, int c, void* out /* bf16 [rows][2c] */, void* stream);
int dynmm_upsample2x_dw3x3_split(const void* in, int n, int h, int w, int c, const float* weight, const float* bias,
                                 const void* skip, void* out_nhwc_bf16, float* out_nchw_f32, uint8_t* labels, void* stream);
/* The same up-sampling with options: DYNMM_UPSAMPLE_SPLIT = the _split variant; DYNMM_UPSAMPLE_REPLICATE = replication
 * padding of the up-sampled map instead of zero padding (Upsample 'learned-3x3', model.py:372-384; with the fixed
 * stencil [1 2 1]^T [1 2 1] / 16 and no bias it IS F.interpolate(mode='bilinear', align_corners=False) at scale 2,
 * 'bilinear' of model.py:364-366). */
#define DYNMM_UPSAMPLE_SPLIT 1
#define DYNMM_UPSAMPLE_REPLICATE 2
/* c_valid (0 = c): only the first c_valid channels exist in the NCHW output [n, c_valid, 2h, 2w] and take part in the
 * arg-max -- a class count that is not a multiple of 8 (SUN RGB-D: 37) is carried as c = 40 NHWC channels. */
int dynmm_upsample2x_dw3x3_ex(const void* in, int n, int h, int w, int c, const float* weight, const float* bias,
                              const void* skip, void* out_nhwc_bf16, float* out_nchw_f32, uint8_t* labels, int flags,
                              int c_valid, void* stream);
/* PyramidPoolingModule with upsampling_mode='bilinear' (context_modules.py:79-81): F.interpolate(mode='bilinear',
 * align_corners=False) of src [n,hs,ws,c] into channels [c_off, c_off + c) of dst [n,h,w,ld]; split: [hi | lo] tensors. */
int dynmm_bilinear_resize_into(const void* src, int n, int hs, int ws, int c, void* dst, int h, int w, int ld, int c_off,
                               int split, void* stream);
int dynmm_adaptive_avgpool_split(const void* in, int n, int h, int w, int c, int ld, int bins, void* out, void* stream);
int dynmm_nearest_resize_into_split(const void* src, int n, int hs, int ws, int c, void* dst, int h, int w, int ld,
                                    int c_off, void* stream);
/* PyramidPoolingModule pooling + broadcast (context_modules.py:69-84) for NHWC bf16 input
 * [n,h,w,ld] (first c channels): adaptive average pool to bins x bins (fp32 accumulate) -> out [n,bins,bins,c] bf16. */
int dynmm_adaptive_avgpool(const void* in, int n, int h, int w, int c, int ld, int bins, void* out, void* stream);
/* nearest-neighbour resize of NHWC bf16 [n,hs,ws,c] into a channel slice of dst [n,h,w,ld]. */
int dynmm_nearest_resize_into(const void* src, int n, int hs, int ws, int c, void* dst, int h, int w,
                              int ld, int c_off, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* DYNMM_B200_H_ */
