"""Thin tensor-level wrappers over the C ABI (one function per exported symbol).

Tensors carry device memory only; every function launches on the current
CUDA stream through ``libdynmm_b200.so`` and raises on failure.  Layout
conventions are those of ``include/dynmm_b200.h``: activations NHWC bf16,
gate path fp32.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (ChainJob, ChainLayer, ChainParams, ConvPairParams, ConvParams, TileFlags, WgradParams, check, ptr,
                   stream_ptr)

Tensor = torch.Tensor


def _cuda(*ts):
    """Every launch goes to the CURRENT device's current stream: reject tensors that live elsewhere (a kernel launched
    on cuda:0 with cuda:1 pointers faults asynchronously) -- callers enter ``torch.cuda.device(t.device)`` first."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not (t.is_cuda and t.is_contiguous()):
            raise _lib.DynmmError("dynmm ops need contiguous CUDA tensors")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.DynmmError(f"dynmm ops launch on the current device (cuda:{cur}) but got a tensor on {t.device}; "
                                  f"wrap the call in `with torch.cuda.device(tensor.device):`")


# ------------------------------------------------------------------ weights

def pack_conv_weight(w: Tensor) -> Tensor:
    """[c_out, c_in, kh, kw] fp32 -> bf16 [kh*kw, c_out_pad16, c_in] (K-major B operand)."""
    c_out, c_in, kh, kw = w.shape
    pad = (c_out + 15) // 16 * 16
    out = torch.zeros(kh * kw, pad, c_in, dtype=torch.bfloat16, device=w.device)
    out[:, :c_out] = w.permute(2, 3, 0, 1).reshape(kh * kw, c_out, c_in).to(torch.bfloat16)
    return out.contiguous()


def fold_bn(bn_w, bn_b, bn_m, bn_v, eps, conv_bias=None):
    """Eval-mode BatchNorm (and an optional preceding conv bias) as y = x*scale + shift."""
    scale = bn_w / torch.sqrt(bn_v + eps)
    shift = bn_b - bn_m * scale
    if conv_bias is not None:
        shift = shift + conv_bias * scale
    return scale.float().contiguous(), shift.float().contiguous()


def _f32(t: Optional[Tensor]) -> Optional[Tensor]:
    return None if t is None else t.detach().float().contiguous()


def fold_pack_conv(w: Tensor, bias: Optional[Tensor] = None, bn: Optional[Sequence[Tensor]] = None, eps: float = 1e-5,
                   split: bool = False):
    """One launch (dynmm_fold_pack_conv): eval-mode BatchNorm ``bn = (weight, bias, running_mean, running_var)`` and
    the conv bias folded, weights packed like :func:`pack_conv_weight`.  -> (packed bf16, shift fp32 [c_out] or None);
    bit-identical to ``pack_conv_weight(w * scale.view(-1,1,1,1))`` with :func:`fold_bn`'s scale / shift.
    ``split``: the folded fp32 weights as bf16 halves [taps, c_out_pad, 3 * c_in] = [hi | lo | hi] for ``conv(split=True)``."""
    lib = _lib.load()
    w = _f32(w)
    bias = _f32(bias)
    bn = [_f32(t) for t in bn] if bn is not None else [None] * 4
    _cuda(w, bias, *bn)
    c_out, c_in, kh, kw = w.shape
    packed = torch.empty(kh * kw, (c_out + 15) // 16 * 16, c_in * (3 if split else 1), dtype=torch.bfloat16, device=w.device)
    shift = torch.empty(c_out, dtype=torch.float32, device=w.device) if (bias is not None or bn[0] is not None) else None
    fn = lib.dynmm_fold_pack_conv_split if split else lib.dynmm_fold_pack_conv
    check(fn(ptr(w), c_out, c_in, kh, kw, ptr(bias), ptr(bn[0]), ptr(bn[1]), ptr(bn[2]), ptr(bn[3]),
             float(eps), ptr(packed), ptr(shift), stream_ptr()), "fold_pack_conv")
    return packed, shift


def fold_bn_cuda(bn_w: Tensor, bn_b: Tensor, bn_m: Tensor, bn_v: Tensor, eps: float, conv_bias: Optional[Tensor] = None):
    """:func:`fold_bn` as one launch of our own kernel (same roundings, bit-identical)."""
    lib = _lib.load()
    bn_w, bn_b, bn_m, bn_v, conv_bias = _f32(bn_w), _f32(bn_b), _f32(bn_m), _f32(bn_v), _f32(conv_bias)
    _cuda(bn_w, bn_b, bn_m, bn_v, conv_bias)
    c = bn_w.numel()
    scale = torch.empty(c, dtype=torch.float32, device=bn_w.device)
    shift = torch.empty(c, dtype=torch.float32, device=bn_w.device)
    check(lib.dynmm_fold_bn(c, ptr(conv_bias), ptr(bn_w), ptr(bn_b), ptr(bn_m), ptr(bn_v), float(eps), ptr(scale),
                            ptr(shift), stream_ptr()), "fold_bn")
    return scale, shift


def permute3d(x: Tensor, perm: Sequence[int]) -> Tensor:
    """``x.permute(perm).contiguous()`` of a 3-D fp32 CUDA tensor in one launch of our own kernel."""
    lib = _lib.load()
    x = _f32(x)
    _cuda(x)
    d = x.shape
    out = torch.empty(d[perm[0]], d[perm[1]], d[perm[2]], dtype=torch.float32, device=x.device)
    check(lib.dynmm_permute3d_f32(ptr(x), d[0], d[1], d[2], perm[0], perm[1], perm[2], ptr(out), stream_ptr()), "permute3d")
    return out


# ------------------------------------------------------------------ conv

class TileFlagPool:
    """Device int32 flags for layer-to-layer overlap (``dynmm_tile_flags``): one pool per engine, zeroed at the start
    of every forward (:meth:`reset`), carved up in launch order -- so the addresses are the same in every forward and a
    captured CUDA graph stays valid.  While a pool is installed as ``ops.FLAG_POOL``, :func:`conv` publishes
    completion flags for every output it can and consumes the flags its inputs carry (``tensor._dynmm_flags``)."""

    def __init__(self, device, capacity: int = 1 << 18):
        self.buf = torch.zeros(capacity, dtype=torch.int32, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def alloc(self, n: int) -> Optional[int]:
        if self.off + n > self.buf.numel():
            return None                      # pool exhausted: this output simply carries no flags
        p = self.buf.data_ptr() + 4 * self.off
        self.off += n
        return p


FLAG_POOL: Optional[TileFlagPool] = None
ACTIVATIONS = {"none": 0, "relu": 1, "swish": 2, "silu": 2, "hswish": 3}


def conv(x: Tensor, weight: Tensor, *, c_out: int, kh: int, kw: int, stride=(1, 1), pad=(0, 0),
         scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, residual: Optional[Tensor] = None,
         relu: bool = False, gated: Optional[Tensor] = None, gate: Optional[Tensor] = None,
         gated_slot: Optional[Tensor] = None, in_map: Optional[Tensor] = None, res_map: Optional[Tensor] = None,
         count: Optional[Tensor] = None,
         out: Optional[Tensor] = None, n_out: Optional[int] = None, c_in: Optional[int] = None,
         out_c_off: int = 0, tile_n: int = 0, max_ctas: int = 0, direct: bool = False,
         trace: Optional[Tensor] = None, volatile_weights: bool = False, dual: Optional[bool] = None,
         residual_settled: bool = False, count_settled: bool = False, split: bool = False) -> Tensor:
    """Fused conv + scale/shift + residual + ReLU + gated add (see dynmm_conv_igemm_fwd).

    ``split`` (DYNMM_CONV_SPLIT, fp32-grade arithmetic): x / out / residual / gated are [hi | lo] bf16 halves with
    pitch 2 * channels (hi in [0, ld/2), lo in [ld/2, ld)), ``weight`` comes from ``fold_pack_conv(split=True)``;
    ``c_in`` / ``c_out`` stay the logical channel counts.

    x: NHWC bf16 [n_in, h, w, in_ld] (``c_in`` <= in_ld selects a channel prefix);
    weight: packed by :func:`pack_conv_weight`.  ``out`` may be a wider NHWC buffer, written
    at channel offset ``out_c_off``.  ``direct=True`` runs the CUDA-core comparator."""
    lib = _lib.load()
    _cuda(x, weight, scale, shift, residual, gated, gate, gated_slot, in_map, count, out)
    n_in, h_in, w_in, in_ld = x.shape
    c_in = (in_ld // 2 if split else in_ld) if c_in is None else c_in
    n = n_in if n_out is None else n_out
    h_out = (h_in + 2 * pad[0] - kh) // stride[0] + 1
    w_out = (w_in + 2 * pad[1] - kw) // stride[1] + 1
    if out is None:
        out = torch.empty(n, h_out, w_out, c_out * (2 if split else 1), dtype=torch.bfloat16, device=x.device)
    p = ConvParams()
    p.in_, p.weight, p.scale, p.shift = ptr(x), ptr(weight), ptr(scale), ptr(shift)
    p.residual, p.res_map = ptr(residual), ptr(res_map)
    p.out = out.data_ptr() + 2 * out_c_off
    p.gated, p.gate, p.gated_slot = ptr(gated), ptr(gate), ptr(gated_slot)
    p.in_map, p.count = ptr(in_map), ptr(count)
    p.n, p.n_in = n, n_in
    p.h_in, p.w_in, p.c_in, p.in_ld = h_in, w_in, c_in, in_ld
    p.h_out, p.w_out, p.c_out, p.out_ld = h_out, w_out, c_out, out.shape[3]
    p.res_ld = residual.shape[3] if residual is not None else 0
    p.gated_ld = gated.shape[3] if gated is not None else 0
    p.kh, p.kw, p.stride_h, p.stride_w, p.pad_h, p.pad_w = kh, kw, stride[0], stride[1], pad[0], pad[1]
    # activation code of dynmm_conv_params.relu: False / True (ReLU) or "relu" / "swish" / "hswish"
    p.relu = ACTIVATIONS[relu] if isinstance(relu, str) else int(relu)
    p.tile_n, p.max_ctas = tile_n, max_ctas
    p.flags = 1 if volatile_weights else 0       # DYNMM_CONV_VOLATILE_WEIGHTS: packed on this stream just before
    if dual is not None:                         # DYNMM_CONV_NO_DUAL / DYNMM_CONV_FORCE_DUAL (default: the planner decides)
        p.flags |= 4 if dual else 2
    if count_settled and count is not None:
        p.flags |= 16                            # DYNMM_CONV_COUNT_SETTLED
    if split:
        p.flags |= 32                            # DYNMM_CONV_SPLIT
    pool = FLAG_POOL if not split else None
    if pool is not None and not direct and CONV_RECORDER is None and trace is None:
        # consume: the input's (and the residual's) completion flags replace the wait for the previous kernel.  Only
        # when everything this launch reads from recent launches is covered: no sample indirection, no gated operand
        # (that one comes from the other stream behind an event), residual flagged or known to be settled
        fin = getattr(x, "_dynmm_flags", None)
        fres = getattr(residual, "_dynmm_flags", None) if residual is not None else None
        if fin is not None and in_map is None and res_map is None and gated is None and \
                (residual is None or fres is not None or residual_settled):
            p.in_flags = fin
            if fres is not None:
                p.res_flags = fres
            elif residual is not None:
                p.flags |= 8                     # DYNMM_CONV_RESIDUAL_SETTLED
        # publish: flags for this output (geometry = the tiling the planner picks for exactly this launch)
        grid = TileFlags()
        if lib.dynmm_conv_tile_grid(ctypes.byref(p), ctypes.byref(grid)) == 0:
            nflags = grid.tiles_h * grid.tiles_w * ((n + grid.box_n - 1) // grid.box_n)
            addr = pool.alloc(nflags)
            if addr is not None:
                grid.flags = addr
                p.out_flags = grid
                keep = TileFlags()
                ctypes.memmove(ctypes.byref(keep), ctypes.byref(grid), ctypes.sizeof(TileFlags))
                out._dynmm_flags = keep
    p.trace = ptr(trace)
    if CONV_RECORDER is not None and not direct:
        # inside `with ConvProgram()`: the convolution becomes a job of the program's current phase
        CONV_RECORDER._record(p, (x, weight, scale, shift, residual, res_map, out, gated, gate, gated_slot, in_map, count))
        return out
    # MACs per output sample (three bf16 products per MAC with split operands), sample slots, device count
    job = (h_out * w_out * c_out * c_in * kh * kw * (3 if split else 1), n, count)
    if CONV_MERGE is not None and not direct:
        # inside `with ConvMerge()`: launched when the block ends -- together with its partner if there is one
        CONV_MERGE._record(p, (x, weight, scale, shift, residual, res_map, out, gated, gate, gated_slot, in_map, count), job)
        return out
    fn = lib.dynmm_conv_direct_fwd if direct else lib.dynmm_conv_igemm_fwd
    if CONV_PROFILER is not None and not direct:
        # (launch closure, [(MACs per output sample, output sample slots, device count tensor or None), ...])
        CONV_PROFILER(lambda: check(fn(ctypes.byref(p), stream_ptr()), "conv_igemm"), [job])
    else:
        check(fn(ctypes.byref(p), stream_ptr()), "conv_direct" if direct else "conv_igemm")
    return out


class ConvMerge:
    """Two :func:`conv` calls of identical geometry (the same layer of the RGB and of the depth encoder) as ONE
    launch (``dynmm_conv_igemm_fwd2``)::

        with ops.ConvMerge():
            yd = conv_d(xd, count=cnt)        # outputs are allocated at once, the launch happens at the block's end
            yr = conv_r(xr)

    Anything else recorded in the block (one call, three calls, two calls the planner tiles differently) is launched
    call by call in recording order -- the results are the same bits either way."""

    def __init__(self):
        self.jobs = []

    def _record(self, p, tensors, job):
        self.jobs.append((p, tensors, job))

    def __enter__(self):
        global CONV_MERGE
        if CONV_MERGE is not None or CONV_RECORDER is not None:
            raise _lib.DynmmError("ConvMerge blocks do not nest")
        CONV_MERGE = self
        return self

    def __exit__(self, exc_type, exc, tb):
        global CONV_MERGE
        CONV_MERGE = None
        if exc_type is None:
            self.launch()
        return False

    def launch(self):
        lib = _lib.load()
        jobs, self.jobs = self.jobs, []
        self.merged = False
        if len(jobs) == 2 and MERGE_ENABLED:
            (pa, _, ja), (pb, _, jb) = jobs

            def both():
                rc = lib.dynmm_conv_igemm_fwd2(ctypes.byref(pa), ctypes.byref(pb), stream_ptr())
                if rc not in (0, -3):
                    check(rc, "conv_igemm_fwd2")
                return rc
            rc = both()                           # the real launch (or DYNMM_EUNSUPPORTED: nothing was launched)
            if rc == 0:
                self.merged = True
                if CONV_PROFILER is not None:
                    CONV_PROFILER(both, [ja, jb], launched=True)
                return
        for p, _, job in jobs:
            if CONV_PROFILER is not None:
                CONV_PROFILER(lambda p=p: check(lib.dynmm_conv_igemm_fwd(ctypes.byref(p), stream_ptr()), "conv_igemm"), [job])
            else:
                check(lib.dynmm_conv_igemm_fwd(ctypes.byref(p), stream_ptr()), "conv_igemm")


class ChainImage:
    """Device-resident description of a run of 3-tap convolutions for :func:`conv_chain` (dynmm_conv_chain_build):
    ``layers`` = [(packed weight bf16 [3][c][c], shift fp32 [c] or None, taps_h, relu, residual, store), ...] with
    residual 0 none / 1 chain input / 2 the ``out`` tensor as stored earlier, store 0 none / 1 ``out`` / 2 ``out_last``.
    Built once per weight set; keeps the weight tensors alive."""

    def __init__(self, layers: Sequence[tuple], c: int, device):
        lib = _lib.load()
        n = len(layers)
        arr = (ChainLayer * n)()
        self.keep = []
        for i, (w, shift, taps_h, relu, residual, store) in enumerate(layers):
            _cuda(w, shift)
            if tuple(w.shape) != (3, c, c) or w.dtype != torch.bfloat16:
                raise _lib.DynmmError(f"ChainImage: layer {i} weight must be packed bf16 [3][{c}][{c}]")
            arr[i].weight, arr[i].shift = ptr(w), ptr(shift)
            arr[i].taps_h, arr[i].relu, arr[i].residual, arr[i].store = int(taps_h), int(relu), int(residual), int(store)
            self.keep += [w, shift]
        nbytes = lib.dynmm_conv_chain_image_bytes(n)
        if nbytes < 0:
            raise _lib.DynmmError(f"ChainImage: unsupported number of layers {n}")
        host = torch.zeros(nbytes, dtype=torch.uint8)
        check(lib.dynmm_conv_chain_build(arr, n, c, host.data_ptr()), "conv_chain_build")
        self.dev = host.to(device)
        self.n_layers, self.c = n, c
        self.stores_out = any(l[5] == 1 for l in layers)
        self.stores_last = any(l[5] == 2 for l in layers)


def nbt1d_chain_layers(blocks: Sequence[Sequence[tuple]], drop_last: bool = False):
    """Layer list for :class:`ChainImage` from NonBottleneck1D blocks, each ``[(weight, shift, relu) x 4]`` in the order
    3x1, 1x3, 3x1, 1x3 (resnet.py:124-147): the fourth convolution of a block adds the block input (the chain input for
    the first block, the previous block's stored output afterwards) and stores the block output.  ``drop_last``: the very
    last convolution is left to the caller (a gated epilogue); the layer before it is stored to ``out_last``."""
    layers = []
    for bi, blk in enumerate(blocks):
        for i, (w, shift, relu) in enumerate(blk):
            last = i == 3
            layers.append((w, shift, i % 2 == 0, relu, (1 if bi == 0 else 2) if last else 0, 1 if last else 0))
    if drop_last:
        layers.pop()
        w, shift, taps_h, relu, res, _ = layers[-1]
        layers[-1] = (w, shift, taps_h, relu, res, 2)
    return layers


def chain_plan(h: int, w: int, c: int, total_slots: int):
    """-> (units, scratch bytes) of a :func:`conv_chain` launch, or None when the geometry is not supported."""
    lib = _lib.load()
    units, nbytes = ctypes.c_int32(0), ctypes.c_longlong(0)
    rc = lib.dynmm_conv_chain_plan(h, w, c, total_slots, ctypes.byref(units), ctypes.byref(nbytes))
    if rc == -3:
        return None
    check(rc, "conv_chain_plan")
    return units.value, nbytes.value


def conv_chain(jobs: Sequence[dict], *, flags: Optional[Tensor] = None, trace: Optional[Tensor] = None):
    """One launch for a run of NonBottleneck1D convolutions per job (dynmm_conv_chain_fwd); bit-identical to the
    per-layer :func:`conv` calls.  ``jobs``: 1 or 2 dicts ``x`` (NHWC bf16 [n,h,w,c]), ``image`` (:class:`ChainImage`),
    optional ``count`` / ``count_settled``.  ``flags``: zeroed int32 [units + sample slots] (allocated when None).
    -> [(out, out_last), ...] per job (None where the image stores nothing there)."""
    lib = _lib.load()
    x0 = jobs[0]["x"]
    _, h, w, c = x0.shape
    total = sum(j["x"].shape[0] for j in jobs)
    plan = chain_plan(h, w, c, total)
    if plan is None:
        raise _lib.DynmmError("conv_chain: unsupported geometry: " + lib.dynmm_last_error().decode())
    units, scratch_bytes = plan
    if flags is None:
        flags = torch.zeros(units + total, dtype=torch.int32, device=x0.device)
    elif flags.numel() < units + total or flags.dtype != torch.int32:
        raise _lib.DynmmError("conv_chain: flags too small")
    scratch = torch.empty(max(scratch_bytes, 16), dtype=torch.uint8, device=x0.device)
    p = ChainParams()
    p.n_jobs, p.h, p.w, p.c = len(jobs), h, w, c
    p.flags, p.scratch, p.scratch_bytes, p.trace = ptr(flags), ptr(scratch), scratch.numel(), ptr(trace)
    outs, prof_jobs, keep = [], [], [flags, scratch]
    for i, j in enumerate(jobs):
        x, img = j["x"], j["image"]
        _cuda(x, j.get("count"))
        if tuple(x.shape[1:]) != (h, w, c) or x.dtype != torch.bfloat16 or img.c != c:
            raise _lib.DynmmError("conv_chain: jobs must share one geometry (NHWC bf16 [n,h,w,c])")
        out = torch.empty_like(x) if img.stores_out else None
        last = torch.empty_like(x) if img.stores_last else None
        pj = p.jobs[i]
        pj.image, pj.in_, pj.out, pj.out_last = img.dev.data_ptr(), ptr(x), ptr(out), ptr(last)
        pj.count, pj.n, pj.n_layers = ptr(j.get("count")), x.shape[0], img.n_layers
        pj.count_settled = int(bool(j.get("count_settled", False)) and j.get("count") is not None)
        outs.append((out, last))
        prof_jobs.append((h * w * c * c * 3 * img.n_layers, x.shape[0], j.get("count")))
        keep.append(x)
    if CONV_PROFILER is not None:
        CONV_PROFILER(lambda: check(lib.dynmm_conv_chain_fwd(ctypes.byref(p), stream_ptr()), "conv_chain"), prof_jobs)
    else:
        check(lib.dynmm_conv_chain_fwd(ctypes.byref(p), stream_ptr()), "conv_chain")
    for o in outs:                      # scratch / flags must outlive the launch: tie them (not the results: no cycle) to it
        for t in o:
            if t is not None:
                t._dynmm_keep = keep
    return outs


def conv_pair(x: Tensor, w1: Tensor, shift1: Optional[Tensor], w2: Tensor, shift2: Optional[Tensor], *,
              residual: Optional[Tensor] = None, relu2: bool = True, count: Optional[Tensor] = None,
              in_map: Optional[Tensor] = None, res_map: Optional[Tensor] = None, n_out: Optional[int] = None,
              out: Optional[Tensor] = None) -> Tensor:
    """conv3x1 + bias + ReLU -> conv1x3 + shift (+ residual) (+ ReLU) of a 64-channel NonBottleneck1D block in one
    kernel (dynmm_conv_pair_fwd); bit-identical to the two :func:`conv` calls.  x NHWC bf16 [n_in, h, w, >= 64]."""
    lib = _lib.load()
    _cuda(x, w1, w2, shift1, shift2, residual, count, in_map, res_map, out)
    n_in, h, w, in_ld = x.shape
    n = n_in if n_out is None else n_out
    if out is None:
        out = torch.empty(n, h, w, 64, dtype=torch.bfloat16, device=x.device)
    p = ConvPairParams()
    p.in_, p.w1, p.shift1, p.w2, p.shift2 = ptr(x), ptr(w1), ptr(shift1), ptr(w2), ptr(shift2)
    p.residual, p.out, p.count, p.in_map, p.res_map = ptr(residual), ptr(out), ptr(count), ptr(in_map), ptr(res_map)
    p.n, p.n_in, p.h, p.w = n, n_in, h, w
    p.in_ld, p.out_ld = in_ld, out.shape[3]
    p.res_ld = residual.shape[3] if residual is not None else 0
    p.relu2 = int(relu2)
    if CONV_PROFILER is not None:
        CONV_PROFILER(lambda: check(lib.dynmm_conv_pair_fwd(ctypes.byref(p), stream_ptr()), "conv_pair"),
                      [(2 * h * w * 64 * 64 * 3, n, count)])
    else:
        check(lib.dynmm_conv_pair_fwd(ctypes.byref(p), stream_ptr()), "conv_pair")
    return out


def pack_conv_weight_dgrad(w: Tensor) -> Tensor:
    """[c_out, c_in, kh, kw] fp32 -> the packed weight of the DATA-GRADIENT convolution:
    bf16 [kh*kw, c_in_pad16, c_out] with the taps mirrored (dx = conv(dy, w^T flipped))."""
    return pack_conv_weight(w.flip(2, 3).transpose(0, 1))


def pack_conv_weight_pair(w: Tensor):
    """fp32 [c_out, c_in, kh, kw] (CUDA) -> (forward, data-gradient) packed bf16 weights in one launch."""
    lib = _lib.load()
    w = w.detach().float().contiguous()
    _cuda(w)
    c_out, c_in, kh, kw = w.shape
    fwd = torch.empty(kh * kw, (c_out + 15) // 16 * 16, c_in, dtype=torch.bfloat16, device=w.device)
    dgr = torch.empty(kh * kw, (c_in + 15) // 16 * 16, c_out, dtype=torch.bfloat16, device=w.device)
    check(lib.dynmm_pack_conv_weight(ptr(w), c_out, c_in, kh, kw, ptr(fwd), ptr(dgr), stream_ptr()), "pack_conv_weight")
    return fwd, dgr


_WGRAD_WS = {}


def _workspace(device, need: int) -> Tensor:
    # one scratch buffer per (device, stream): calls on a stream are ordered, so it can be shared
    key = (device.index, stream_ptr())
    ws = _WGRAD_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _WGRAD_WS[key] = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
    return ws


def channel_sum(x: Tensor, c: Optional[int] = None, out: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
    """x NHWC bf16 [..., ld] -> fp32 [c] sums over all leading dimensions (the bias gradient)."""
    lib = _lib.load()
    _cuda(x, out)
    ld = x.shape[-1]
    c = ld if c is None else c
    rows = x.numel() // ld
    if out is None:
        out = torch.empty(c, dtype=torch.float32, device=x.device)
    need = lib.dynmm_channel_sum_workspace(rows, c)
    if need < 0:
        raise _lib.DynmmError("channel_sum: c must be a multiple of 8 (<= 2048)")
    ws = _workspace(x.device, need)
    check(lib.dynmm_channel_sum(ptr(x), rows, c, ld, ptr(out), ws.data_ptr(), ws.numel(), int(accumulate),
                                stream_ptr()), "channel_sum")
    return out


def conv_wgrad(x: Tensor, dy: Tensor, *, kh: int, kw: int, stride=(1, 1), pad=(0, 0), c_in: Optional[int] = None,
               c_out: Optional[int] = None, out: Optional[Tensor] = None, accumulate: bool = False,
               direct: bool = False, max_ctas: int = 0) -> Tensor:
    """Weight gradient of ``conv`` (see dynmm_conv_wgrad): x NHWC bf16 [n,h,w,x_ld], dy NHWC bf16
    [n,ho,wo,dy_ld] -> fp32 [c_out, c_in, kh, kw].  ``direct=True`` runs the CUDA-core comparator."""
    lib = _lib.load()
    _cuda(x, dy, out)
    n, h_in, w_in, x_ld = x.shape
    _, h_out, w_out, dy_ld = dy.shape
    c_in = x_ld if c_in is None else c_in
    c_out = dy_ld if c_out is None else c_out
    if out is None:
        out = torch.empty(c_out, c_in, kh, kw, dtype=torch.float32, device=x.device)
    p = WgradParams()
    p.x, p.dy, p.dw = ptr(x), ptr(dy), ptr(out)
    p.n, p.h_in, p.w_in, p.c_in, p.x_ld = n, h_in, w_in, c_in, x_ld
    p.h_out, p.w_out, p.c_out, p.dy_ld = h_out, w_out, c_out, dy_ld
    p.kh, p.kw, p.stride_h, p.stride_w, p.pad_h, p.pad_w = kh, kw, stride[0], stride[1], pad[0], pad[1]
    p.accumulate, p.max_ctas = int(accumulate), max_ctas
    if direct:
        check(lib.dynmm_conv_wgrad_direct(ctypes.byref(p), stream_ptr()), "conv_wgrad_direct")
        return out
    need = lib.dynmm_conv_wgrad_workspace(ctypes.byref(p))
    if need < 0:
        check(-1, "conv_wgrad_workspace")
    ws = _workspace(x.device, need)
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib.dynmm_conv_wgrad(ctypes.byref(p), stream_ptr()), "conv_wgrad")
    return out


# bench.py installs a callable(launch, jobs, launched=False) here to time every tensor-core conv launch
CONV_PROFILER = None
CONV_RECORDER = None
CONV_MERGE = None
MERGE_ENABLED = True


class ConvProgram:
    """Records :func:`conv` calls and runs them as ONE persistent cooperative launch (dynmm_conv_program_*).

    ::

        with ops.ConvProgram() as prog:
            y = ops.conv(x, w1, ...)          # phase 0
            prog.next_phase()
            z = ops.conv(y, w2, ...)          # phase 1: may read what phase 0 wrote
        # leaving the block builds the program image, uploads it and launches it on the current stream

    Convolutions recorded in one phase must be independent (at most 4).  Tensors are allocated as usual while
    recording, only the launches are deferred.  ``prog.hold`` keeps the pinned host image and the device
    buffers alive: a CUDA graph that captured the launch replays the upload from that pinned memory."""

    def __init__(self):
        self.jobs, self.phases, self.keep = [], [], []
        self.phase = 0
        self.hold = None
        self.cfg = None
        self._prev = None

    def next_phase(self):
        if self.phases and self.phases[-1] == self.phase:     # empty phases are not numbered
            self.phase += 1

    def jobs_in_phase(self) -> int:
        return sum(1 for ph in self.phases if ph == self.phase)

    def _record(self, p, tensors):
        if self.jobs_in_phase() >= 4:
            raise _lib.DynmmError("ConvProgram: at most 4 convolutions per phase")
        self.jobs.append(p)
        self.phases.append(self.phase)
        self.keep.append(tensors)

    def __enter__(self):
        global CONV_RECORDER
        if CONV_RECORDER is not None:
            raise _lib.DynmmError("ConvProgram blocks do not nest")
        CONV_RECORDER = self
        return self

    def __exit__(self, exc_type, exc, tb):
        global CONV_RECORDER
        CONV_RECORDER = None
        if exc_type is None and self.jobs:
            self.launch()
        return False

    def flops(self) -> float:
        """2 x MACs the program executes (samples beyond a job's device-side ``count`` are not computed)."""
        total = 0.0
        for p, tensors in zip(self.jobs, self.keep):
            count = tensors[-1]
            active = min(int(count.item()), p.n) if count is not None else p.n
            total += 2.0 * p.h_out * p.w_out * p.c_out * p.c_in * p.kh * p.kw * active
        return total

    def launch(self, upload: bool = True, trace: Optional[Tensor] = None):
        lib = _lib.load()
        n = len(self.jobs)
        dev = self.keep[0][0].device
        if self.hold is None:
            nbytes = lib.dynmm_conv_program_bytes(n)
            host = torch.empty(nbytes + 128, dtype=torch.uint8).pin_memory()
            off = (-host.data_ptr()) % 128
            host_img = host[off:off + nbytes]
            jobs = (ConvParams * n)(*self.jobs)
            phases = (ctypes.c_int32 * n)(*self.phases)
            cfg = (ctypes.c_int32 * 3)()
            check(lib.dynmm_conv_program_build(jobs, phases, n, host_img.data_ptr(), nbytes, cfg), "conv_program_build")
            dev_buf = torch.empty(nbytes + 128, dtype=torch.uint8, device=dev)
            doff = (-dev_buf.data_ptr()) % 128
            dev_img = dev_buf[doff:doff + nbytes]
            barrier = torch.empty(2, dtype=torch.int32, device=dev)
            self.cfg = cfg
            self.hold = (host, host_img, dev_buf, dev_img, barrier)
        _, host_img, _, dev_img, barrier = self.hold
        if upload:
            dev_img.copy_(host_img, non_blocking=True)
        check(lib.dynmm_conv_program_launch(dev_img.data_ptr(), host_img.data_ptr(), barrier.data_ptr(), ptr(trace),
                                            stream_ptr()),
              "conv_program_launch")

    @property
    def n_phases(self):
        return int(self.cfg[2]) if self.cfg is not None else 0


# ------------------------------------------------------------------ stem / gate

def stem(rgb: Tensor, depth: Tensor, w_rgb: Tensor, scale_rgb: Tensor, shift_rgb: Tensor, w_d: Tensor,
         scale_d: Tensor, shift_d: Tensor, want_f32: bool = True, se_rgb: Optional[Tensor] = None,
         se_depth: Optional[Tensor] = None):
    """rgb [b,3,h,w], depth [b,1,h,w] NCHW fp32 -> pooled NHWC maps
    (rgb_f32, depth_f32, rgb_bf16, depth_bf16); the fp32 pair is None when not wanted.
    se_rgb / se_depth [b,64]: SE scales of the two stem streams (SE-add fusion)."""
    lib = _lib.load()
    _cuda(rgb, depth, w_rgb, w_d)
    b, _, h, w = rgb.shape
    hs, ws = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
    hp, wp = (hs + 2 - 3) // 2 + 1, (ws + 2 - 3) // 2 + 1
    dev = rgb.device
    r32 = torch.empty(b, hp, wp, 64, dtype=torch.float32, device=dev) if want_f32 else None
    d32 = torch.empty(b, hp, wp, 64, dtype=torch.float32, device=dev) if want_f32 else None
    r16 = torch.empty(b, hp, wp, 64, dtype=torch.bfloat16, device=dev)
    d16 = torch.empty(b, hp, wp, 64, dtype=torch.bfloat16, device=dev)
    check(lib.dynmm_stem_fwd(ptr(rgb), ptr(depth), b, h, w, ptr(w_rgb), ptr(scale_rgb), ptr(shift_rgb), ptr(w_d),
                             ptr(scale_d), ptr(shift_d), ptr(r32), ptr(d32), ptr(r16), ptr(d16), ptr(se_rgb),
                             ptr(se_depth), None, stream_ptr()), "stem")
    return r32, d32, r16, d16


def stem_s2d_pack_weights(w_rgb: Tensor, w_d: Tensor) -> Tensor:
    """[7][7][3][64] / [7][7][1][64] fp32 stem weights -> bf16 [2 (hi, lo)][128][256] for :func:`stem_s2d`."""
    lib = _lib.load()
    _cuda(w_rgb, w_d)
    out = torch.empty(2, 128, 256, dtype=torch.bfloat16, device=w_rgb.device)
    check(lib.dynmm_stem_s2d_pack_weights(ptr(w_rgb), ptr(w_d), ptr(out), stream_ptr()), "stem_s2d_pack_weights")
    return out


def stem_s2d_bn_host(scale_rgb: Tensor, shift_rgb: Tensor, scale_d: Tensor, shift_d: Tensor) -> Tensor:
    """Host copy [scale_rgb | shift_rgb | scale_d | shift_d] (256 fp32) of the stem's BN vectors for :func:`stem_s2d`'s
    ``bn_host``: made once per engine (it synchronises), then passed as a kernel parameter on every launch."""
    out = torch.cat([t.detach().float().reshape(64) for t in (scale_rgb, shift_rgb, scale_d, shift_d)]).cpu().contiguous()
    return out


def stem_s2d(rgb: Tensor, depth: Tensor, w_packed: Tensor, scale_rgb: Tensor, shift_rgb: Tensor, scale_d: Tensor,
             shift_d: Tensor, want_f32: bool = True, bn_host: Optional[Tensor] = None, want_bf16: bool = True,
             split: bool = False):
    """:func:`stem` for plain `add` fusion with the im2col done by TMA (dynmm_stem_s2d_fwd); 2 launches.
    ``bn_host``: :func:`stem_s2d_bn_host` of the same four vectors (constant-bank path of the epilogue).
    ``split``: the bf16 outputs are [b, hp, wp, 128] = [hi | lo] halves (:func:`split_from_f32` of the fp32 maps)."""
    lib = _lib.load()
    _cuda(rgb, depth, w_packed)
    b, _, h, w = rgb.shape
    hs, ws = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
    hp, wp = (hs + 2 - 3) // 2 + 1, (ws + 2 - 3) // 2 + 1
    dev = rgb.device
    r32 = torch.empty(b, hp, wp, 64, dtype=torch.float32, device=dev) if want_f32 else None
    d32 = torch.empty(b, hp, wp, 64, dtype=torch.float32, device=dev) if want_f32 else None
    c16 = 128 if split else 64
    r16 = torch.empty(b, hp, wp, c16, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    d16 = torch.empty(b, hp, wp, c16, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    need = lib.dynmm_stem_s2d_workspace(b, h, w)
    ws_buf = torch.empty(need, dtype=torch.uint8, device=dev)
    bn_ptr = None
    if bn_host is not None:
        if bn_host.is_cuda or bn_host.dtype != torch.float32 or bn_host.numel() != 256 or not bn_host.is_contiguous():
            raise _lib.DynmmError("stem_s2d: bn_host must be a contiguous CPU float32 tensor of 256 values")
        bn_ptr = bn_host.data_ptr()
    check(lib.dynmm_stem_s2d_fwd(ptr(rgb), ptr(depth), b, h, w, ptr(w_packed), ptr(scale_rgb), ptr(shift_rgb),
                                 ptr(scale_d), ptr(shift_d), ptr(ws_buf), need, ptr(r32), ptr(d32), ptr(r16), ptr(d16),
                                 1 if split else 0, bn_ptr, stream_ptr()), "stem_s2d")
    return r32, d32, r16, d16


def stem_squeeze(rgb: Tensor, depth: Tensor, w_rgb, scale_rgb, shift_rgb, w_d, scale_d, shift_d):
    """Pass 1 of the SE-add stem: per-tile channel sums of the two unfused stem maps.
    -> (partial [b, tiles_per_sample, 128] fp32 with rows [rgb 64 | depth 64], 1 / stem map area)."""
    lib = _lib.load()
    _cuda(rgb, depth)
    b, _, h, w = rgb.shape
    tiles = lib.dynmm_stem_gap_tiles(b, h, w)
    partial = torch.empty(b, tiles // b, 128, dtype=torch.float32, device=rgb.device)
    check(lib.dynmm_stem_fwd(ptr(rgb), ptr(depth), b, h, w, ptr(w_rgb), ptr(scale_rgb), ptr(shift_rgb), ptr(w_d),
                             ptr(scale_d), ptr(shift_d), None, None, None, None, None, None, ptr(partial),
                             stream_ptr()), "stem_squeeze")
    hs, ws = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
    return partial, 1.0 / (hs * ws)


def gap_partial(x: Tensor, c: Optional[int] = None, count: Optional[Tensor] = None, split: bool = False) -> Tensor:
    """x NHWC bf16 [n,h,w,ld] -> partial channel sums [n, 64, c] fp32 (deterministic).
    ``split``: x is [hi | lo] (lo half at channel ld / 2); the sums are those of hi + lo."""
    lib = _lib.load()
    _cuda(x, count)
    n, h, w, ld = x.shape
    c = (ld // 2 if split else ld) if c is None else c
    out = torch.empty(n, 64, c, dtype=torch.float32, device=x.device)
    fn = lib.dynmm_gap_partial_split if split else lib.dynmm_gap_partial
    check(fn(ptr(x), n, h * w, c, ld, ptr(count), ptr(out), stream_ptr()), "gap_partial")
    return out


def se_mlp(partial: Tensor, inv_area: float, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor,
           count: Optional[Tensor] = None, c_off: int = 0, c: Optional[int] = None) -> Tensor:
    """partial [rows, chunks, ld] fp32 channel sums; the SE block acts on columns [c_off, c_off+c).
    -> sigma [rows, c]."""
    lib = _lib.load()
    _cuda(partial, w1, b1, w2, b2, count)
    rows, chunks, ld = partial.shape
    c = ld - c_off if c is None else c
    hidden = w1.shape[0]
    sigma = torch.empty(rows, c, dtype=torch.float32, device=partial.device)
    check(lib.dynmm_se_mlp(ptr(partial), rows, chunks, ld, c_off, float(inv_area), c, hidden, ptr(w1), ptr(b1),
                           ptr(w2), ptr(b2), ptr(count), ptr(sigma), stream_ptr()), "se_mlp")
    return sigma


def se_gated_fuse(rgb: Tensor, depth: Tensor, sig_r: Tensor, sig_d: Tensor, gate: Tensor,
                  slot: Optional[Tensor] = None, out: Optional[Tensor] = None, split: bool = False) -> Tensor:
    """``split``: rgb / depth / out are [hi | lo] tensors (channels 2c; ``out`` may be wider, lo half at its middle)."""
    lib = _lib.load()
    _cuda(rgb, depth, sig_r, sig_d, gate, slot, out)
    n, h, w, c = rgb.shape
    out = torch.empty_like(rgb) if out is None else out
    if split:
        check(lib.dynmm_se_gated_fuse_split(ptr(rgb), ptr(depth), ptr(sig_r), ptr(sig_d), ptr(gate), ptr(slot), n, h * w,
                                            c // 2, out.shape[3], ptr(out), stream_ptr()), "se_gated_fuse_split")
        return out
    check(lib.dynmm_se_gated_fuse(ptr(rgb), ptr(depth), ptr(sig_r), ptr(sig_d), ptr(gate), ptr(slot), n, h * w, c,
                                  out.shape[3], ptr(out), stream_ptr()), "se_gated_fuse")
    return out


def global_gate_logits(rgb32: Tensor, depth32: Tensor, w1, scale1, shift1, w2, scale2, shift2, wfc,
                       work: Optional[Tensor] = None) -> Tensor:
    lib = _lib.load()
    _cuda(rgb32, depth32)
    b, h, w, _ = rgb32.shape
    need = lib.dynmm_global_gate_workspace(b, h, w)
    if need < 0:
        raise _lib.DynmmError(f"global gate: feature map {h}x{w} too small for two 5x5/s2 convolutions")
    if work is None or work.numel() < need:
        work = torch.empty(need, dtype=torch.uint8, device=rgb32.device)
    logits = torch.empty(b, 5, dtype=torch.float32, device=rgb32.device)
    check(lib.dynmm_global_gate_logits(ptr(rgb32), ptr(depth32), b, h, w, ptr(w1), ptr(scale1), ptr(shift1), ptr(w2),
                                       ptr(scale2), ptr(shift2), ptr(wfc), ptr(work), ptr(logits), stream_ptr()),
          "global_gate_logits")
    return logits


def global_gate_decide(rgb32: Tensor, depth32: Tensor, w1, scale1, shift1, w2, scale2, shift2, wfc, tau: float, hard: bool,
                       hist: Optional[Tensor] = None):
    """GlobalGate.forward + DiffSoftmax + the skip plan in three launches (dynmm_global_gate_decide)
    -> (weight [b,5], GatePlan, logits [b,5])."""
    lib = _lib.load()
    _cuda(rgb32, depth32, hist)
    b, h, w, _ = rgb32.shape
    need = lib.dynmm_global_gate_workspace(b, h, w)
    if need < 0:
        raise _lib.DynmmError(f"global gate: feature map {h}x{w} too small for two 5x5/s2 convolutions")
    work = torch.empty(need, dtype=torch.uint8, device=rgb32.device)
    logits = torch.empty(b, 5, dtype=torch.float32, device=rgb32.device)
    weight = torch.empty(b, 5, dtype=torch.float32, device=rgb32.device)
    plan = GatePlan(b, rgb32.device)
    check(lib.dynmm_global_gate_decide(ptr(rgb32), ptr(depth32), b, h, w, ptr(w1), ptr(scale1), ptr(shift1), ptr(w2),
                                       ptr(scale2), ptr(shift2), ptr(wfc), ptr(work), float(tau), int(hard), ptr(logits),
                                       ptr(weight), ptr(plan.g), ptr(plan.perm), ptr(plan.slot), ptr(plan.count), ptr(hist),
                                       stream_ptr()), "global_gate_decide")
    plan._work = work
    return weight, plan, logits


def diffsoftmax_fwd(logits: Tensor, tau: float, hard: bool):
    """-> (y, y_soft, index) ; logits [rows, n] fp32."""
    lib = _lib.load()
    _cuda(logits)
    rows, n = logits.shape
    y = torch.empty_like(logits)
    ys = torch.empty_like(logits)
    idx = torch.empty(rows, dtype=torch.int32, device=logits.device)
    check(lib.dynmm_diffsoftmax_fwd(ptr(logits), rows, n, float(tau), int(hard), ptr(y), ptr(ys), ptr(idx),
                                    stream_ptr()), "diffsoftmax_fwd")
    return y, ys, idx


def diffsoftmax_bwd(grad_y: Tensor, y_soft: Tensor, tau: float) -> Tensor:
    lib = _lib.load()
    grad_y = grad_y.contiguous()
    _cuda(grad_y, y_soft)
    rows, n = y_soft.shape
    g = torch.empty_like(y_soft)
    check(lib.dynmm_diffsoftmax_bwd(ptr(grad_y), ptr(y_soft), rows, n, float(tau), ptr(g), stream_ptr()),
          "diffsoftmax_bwd")
    return g


class GatePlan:
    """Device-side skip lists derived from gate weights (see dynmm_gate_plan)."""

    def __init__(self, b: int, device):
        self.b = b
        self.g = torch.empty(4, b, dtype=torch.float32, device=device)
        self.perm = torch.empty(b, dtype=torch.int32, device=device)
        self.slot = torch.empty(b, dtype=torch.int32, device=device)
        self.count = torch.empty(4, dtype=torch.int32, device=device)


def gate_plan(weight: Tensor, plan: Optional[GatePlan] = None, hist: Optional[Tensor] = None) -> GatePlan:
    lib = _lib.load()
    _cuda(weight, hist)
    b = weight.shape[0]
    if plan is None:
        plan = GatePlan(b, weight.device)
    check(lib.dynmm_gate_plan(ptr(weight), b, ptr(plan.g), ptr(plan.perm), ptr(plan.slot), ptr(plan.count), ptr(hist),
                              stream_ptr()), "gate_plan")
    return plan


# ------------------------------------------------------------------ elementwise

def gated_add(a: Tensor, b: Tensor, gate: Tensor, slot: Optional[Tensor] = None, out: Optional[Tensor] = None):
    lib = _lib.load()
    _cuda(a, b, gate, slot, out)
    n = a.shape[0]
    per = a.numel() // n
    out = torch.empty_like(a) if out is None else out
    check(lib.dynmm_gated_add_fwd(ptr(a), ptr(b), ptr(gate), ptr(slot), n, per, ptr(out), stream_ptr()), "gated_add")
    return out


def gated_add_f32_fwd(a: Tensor, b: Tensor, gate: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(a, b, gate)
    n = a.shape[0]
    out = torch.empty_like(a)
    check(lib.dynmm_gated_add_f32_fwd(ptr(a), ptr(b), ptr(gate), n, a.numel() // n, ptr(out), stream_ptr()),
          "gated_add_f32_fwd")
    return out


def gated_add_f32_bwd(grad: Tensor, b: Tensor, gate: Tensor, need_grad_b: bool = True):
    lib = _lib.load()
    grad = grad.contiguous()
    _cuda(grad, b, gate)
    n = b.shape[0]
    grad_b = torch.empty_like(b) if need_grad_b else None
    gg = torch.empty(n * 65, dtype=torch.float32, device=b.device)
    check(lib.dynmm_gated_add_f32_bwd(ptr(grad), ptr(b), ptr(gate), n, b.numel() // n, ptr(grad_b), ptr(gg),
                                      stream_ptr()), "gated_add_f32_bwd")
    return grad_b, gg[:n]


def _ptr_array(ts: Sequence[Optional[Tensor]]):
    arr = (ctypes.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = ptr(t)
    return arr


def softgate_mix_fwd(preds: Sequence[Tensor], w: Tensor, rows: Optional[Sequence[Optional[Tensor]]] = None) -> Tensor:
    lib = _lib.load()
    _cuda(w, *preds)
    b, ne = w.shape
    c = preds[0].shape[1]
    out = torch.empty(b, c, dtype=torch.float32, device=w.device)
    rows_arr = _ptr_array(rows) if rows is not None else None
    check(lib.dynmm_softgate_mix_fwd(_ptr_array(preds), rows_arr, ptr(w), b, c, ne, ptr(out), stream_ptr()),
          "softgate_mix_fwd")
    return out


def softgate_mix_bwd(grad_out: Tensor, preds: Sequence[Tensor], w: Tensor, need_pred_grads: Sequence[bool]):
    lib = _lib.load()
    grad_out = grad_out.contiguous()
    _cuda(grad_out, w, *preds)
    b, ne = w.shape
    c = preds[0].shape[1]
    grads = [torch.empty_like(p) if need else None for p, need in zip(preds, need_pred_grads)]
    gw = torch.empty_like(w)
    check(lib.dynmm_softgate_mix_bwd(ptr(grad_out), _ptr_array(preds), ptr(w), b, c, ne, _ptr_array(grads), ptr(gw),
                                     stream_ptr()), "softgate_mix_bwd")
    return grads, gw


def compact_rows(w: Tensor, expert: int):
    """-> (idx [b] int32, inv [b] int32, count [1] int32) for rows with w[:, expert] != 0."""
    lib = _lib.load()
    _cuda(w)
    b, ne = w.shape
    idx = torch.empty(b, dtype=torch.int32, device=w.device)
    inv = torch.empty(b, dtype=torch.int32, device=w.device)
    cnt = torch.empty(1, dtype=torch.int32, device=w.device)
    check(lib.dynmm_compact_rows(ptr(w), b, ne, expert, ptr(idx), ptr(inv), ptr(cnt), stream_ptr()), "compact_rows")
    return idx, inv, cnt


def nchw_f32_to_nhwc_bf16(x: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(x)
    n, c, h, w = x.shape
    out = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=x.device)
    check(lib.dynmm_nchw_f32_to_nhwc_bf16(ptr(x), n, c, h, w, ptr(out), stream_ptr()), "nchw_f32_to_nhwc_bf16")
    return out


def nhwc_bf16_to_nchw_f32(x: Tensor, c: Optional[int] = None) -> Tensor:
    lib = _lib.load()
    _cuda(x)
    n, h, w, ld = x.shape
    c = ld if c is None else c
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=x.device)
    check(lib.dynmm_nhwc_bf16_to_nchw_f32(ptr(x), n, c, h, w, ld, ptr(out), stream_ptr()), "nhwc_bf16_to_nchw_f32")
    return out


def split_from_f32(x: Tensor) -> Tensor:
    """fp32 [..., c] -> bf16 [..., 2c] = [hi | lo] halves (hi = bf16(x), lo = bf16(x - hi)): the activation format of
    the fp32-grade ("f32x3") engine mode."""
    lib = _lib.load()
    x = x.contiguous()
    _cuda(x)
    c = x.shape[-1]
    out = torch.empty(*x.shape[:-1], 2 * c, dtype=torch.bfloat16, device=x.device)
    check(lib.dynmm_split_from_f32(ptr(x), x.numel() // c, c, ptr(out), stream_ptr()), "split_from_f32")
    return out


def upsample2x_dw3x3(x: Tensor, weight: Tensor, bias: Optional[Tensor], skip: Optional[Tensor] = None,
                     to_nchw_f32: bool = False, out: Optional[Tensor] = None, labels: Optional[Tensor] = None,
                     want_logits: bool = True, split: bool = False, replicate: bool = False,
                     c_valid: Optional[int] = None):
    """x NHWC bf16; weight fp32 tap-major [9, c] (``conv.weight.reshape(c, 9).t().contiguous()``).
    Final upsampling (``to_nchw_f32``): returns the NCHW fp32 logits; with ``labels`` (uint8 [n,2h,2w]) the
    arg-max over channels is produced in the same pass, and ``want_logits=False`` skips the logits.
    ``replicate``: replication instead of zero padding of the up-sampled map ('learned-3x3'; with the stencil of
    :func:`bilinear_stencil` and no bias this is bilinear x2 up-sampling, align_corners=False)."""
    lib = _lib.load()
    _cuda(x, weight, bias, skip, out, labels)
    n, h, w, ld = x.shape
    c = ld // 2 if split else ld                 # split: x / skip / NHWC out are [hi | lo] halves
    flags = (1 if split else 0) | (2 if replicate else 0)       # DYNMM_UPSAMPLE_SPLIT | DYNMM_UPSAMPLE_REPLICATE
    fn = lib.dynmm_upsample2x_dw3x3_ex
    if to_nchw_f32 or labels is not None:
        cv = c if c_valid is None else c_valid        # classes that exist (c may be padded to a multiple of 8)
        if want_logits:
            out = torch.empty(n, cv, 2 * h, 2 * w, dtype=torch.float32, device=x.device) if out is None else out
        else:
            out = None
        check(fn(ptr(x), n, h, w, c, ptr(weight), ptr(bias), None, None, ptr(out), ptr(labels), flags, cv, stream_ptr()),
              "upsample2x_dw3x3")
    else:
        out = torch.empty(n, 2 * h, 2 * w, ld, dtype=torch.bfloat16, device=x.device) if out is None else out
        check(fn(ptr(x), n, h, w, c, ptr(weight), ptr(bias), ptr(skip), ptr(out), None, None, flags, 0, stream_ptr()),
              "upsample2x_dw3x3")
    return out


def bilinear_stencil(c: int, device) -> Tensor:
    """Tap-major [9, c] stencil [1 2 1]^T [1 2 1] / 16: nearest x2 + this depthwise conv with replication padding =
    ``F.interpolate(scale_factor=2, mode='bilinear', align_corners=False)`` (the reference initialises its learned
    up-sampling with it, model.py:385-395)."""
    k = torch.tensor([0.0625, 0.125, 0.0625, 0.125, 0.25, 0.125, 0.0625, 0.125, 0.0625], device=device)
    return k.view(9, 1).expand(9, c).contiguous()


def nearest_stencil(c: int, device) -> Tensor:
    """Tap-major [9, c] identity stencil: plain nearest x2 up-sampling through the same kernel."""
    k = torch.zeros(9, c, device=device)
    k[4] = 1.0
    return k


def bilinear_resize_into(src: Tensor, dst: Tensor, c_off: int, split: bool = False) -> None:
    """``F.interpolate(mode='bilinear', align_corners=False)`` of src [n,hs,ws,c] into dst[..., c_off:c_off+c]."""
    lib = _lib.load()
    _cuda(src, dst)
    n, hs, ws, c = src.shape
    _, h, w, ld = dst.shape
    check(lib.dynmm_bilinear_resize_into(ptr(src), n, hs, ws, c // 2 if split else c, ptr(dst), h, w, ld, c_off, int(split),
                                         stream_ptr()), "bilinear_resize_into")


def adaptive_avgpool(x: Tensor, bins: int, c: Optional[int] = None, split: bool = False) -> Tensor:
    lib = _lib.load()
    _cuda(x)
    n, h, w, ld = x.shape
    c = (ld // 2 if split else ld) if c is None else c
    out = torch.empty(n, bins, bins, c * (2 if split else 1), dtype=torch.bfloat16, device=x.device)
    fn = lib.dynmm_adaptive_avgpool_split if split else lib.dynmm_adaptive_avgpool
    check(fn(ptr(x), n, h, w, c, ld, bins, ptr(out), stream_ptr()), "adaptive_avgpool")
    return out


def nearest_resize_into(src: Tensor, dst: Tensor, c_off: int, split: bool = False) -> None:
    lib = _lib.load()
    _cuda(src, dst)
    n, hs, ws, c = src.shape
    _, h, w, ld = dst.shape
    if split:
        check(lib.dynmm_nearest_resize_into_split(ptr(src), n, hs, ws, c // 2, ptr(dst), h, w, ld, c_off, stream_ptr()),
              "nearest_resize_into_split")
        return
    check(lib.dynmm_nearest_resize_into(ptr(src), n, hs, ws, c, ptr(dst), h, w, ld, c_off, stream_ptr()),
          "nearest_resize_into")


# ------------------------------------------------------------------ eval post-processing

def argmax_confusion(logits: Tensor, label_orig: Optional[Tensor] = None, cm: Optional[Tensor] = None,
                     want_pred: bool = False):
    """logits NCHW fp32; label_orig uint8 [n,h,w] (0 = void); cm int64 [c,c] accumulated in place.
    -> pred uint8 [n,h,w] or None."""
    lib = _lib.load()
    _cuda(logits, label_orig, cm)
    n, c, h, w = logits.shape
    pred = torch.empty(n, h, w, dtype=torch.uint8, device=logits.device) if want_pred else None
    check(lib.dynmm_argmax_confusion(ptr(logits), ptr(label_orig), n, c, h, w, ptr(cm), ptr(pred), stream_ptr()),
          "argmax_confusion")
    return pred


def miou(cm: Tensor):
    """-> (miou [1] float64, iou [c] float64) device tensors."""
    lib = _lib.load()
    _cuda(cm)
    c = cm.shape[0]
    iou = torch.empty(c, dtype=torch.float64, device=cm.device)
    m = torch.empty(1, dtype=torch.float64, device=cm.device)
    check(lib.dynmm_miou(ptr(cm), c, ptr(iou), ptr(m), stream_ptr()), "miou")
    return m, iou


# ------------------------------------------------------------------ training loss

def ce2d_fwd(logits: Tensor, targets: Tensor, weight: Tensor):
    """One scale of CrossEntropyLoss2d (dynmm_ce2d_fwd): logits fp32 NCHW, targets int32 [n,h,w] (0 = void).
    -> (loss [1], lse [n,h,w], divisor [1])."""
    lib = _lib.load()
    _cuda(logits, targets, weight)
    n, c, h, w = logits.shape
    need = lib.dynmm_ce2d_workspace(n, c, h, w)
    if need < 0:
        raise _lib.DynmmError("ce2d: 1..255 classes")
    ws = torch.empty(need, dtype=torch.uint8, device=logits.device)
    lse = torch.empty(n, h, w, dtype=torch.float32, device=logits.device)
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    divisor = torch.empty(1, dtype=torch.float32, device=logits.device)
    check(lib.dynmm_ce2d_fwd(ptr(logits), ptr(targets), ptr(weight), n, c, h, w, ptr(ws), need, ptr(lse), ptr(loss),
                             ptr(divisor), stream_ptr()), "ce2d_fwd")
    return loss, lse, divisor


def ce2d_bwd(logits: Tensor, targets: Tensor, weight: Tensor, lse: Tensor, divisor: Tensor, grad_out: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(logits, targets, weight, lse, divisor, grad_out)
    n, c, h, w = logits.shape
    grad = torch.empty_like(logits)
    check(lib.dynmm_ce2d_bwd(ptr(logits), ptr(targets), ptr(weight), ptr(lse), ptr(divisor), ptr(grad_out), n, c, h, w,
                             ptr(grad), stream_ptr()), "ce2d_bwd")
    return grad
