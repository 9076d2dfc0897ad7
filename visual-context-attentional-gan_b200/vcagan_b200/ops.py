"""Autograd operators over the C ABI of libvcagan_b200.so.

Every op is a torch.autograd.Function whose forward AND backward are calls into the hand-written CUDA library
(see include/vcagan.h).  Ops on the discriminator path (conv / LeakyReLU / avg-pool / scale-add / spatial mean /
linear) express their backward through other differentiable ops of this file, so
``torch.autograd.grad(..., create_graph=True)`` (the R1 penalty of train.py:188-194) works to any order.

Internal activation layout is channels-last: (N, H, W, C) or (N, D, H, W, C), contiguous, fp32 or bf16.
torch is used for memory, streams, views/permutes/cat (data movement) only.
"""
import math
import os
import weakref
from typing import Optional, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import ConvGeom, lib

F32, BF16 = 0, 1
ACT_NONE, ACT_LRELU, ACT_PRELU, ACT_RELU = 0, 1, 2, 3


class Config:
    """compute dtype of the activations ('fp32' = exact SIMT path, 'bf16' = tcgen05 path) and switches."""
    dtype = torch.float32
    use_tc = True          # use the tcgen05 kernels when dtype is bf16 and the geometry is supported
    skip_unneeded_wgrad = True
    gru_persistent = True   # one cooperative launch per GRU layer and pass (falls back to per-step kernels)
    fuse_grad_accum = True  # conv weight / bias gradients are accumulated straight into the FlatGroup .grad views
    grad_slab = os.environ.get("VCA_GRAD_SLAB", "1") != "0"   # ... tcgen05 wgrads through a tap-major slab + TMA reduce-add (trainer.FlatGroup.flush_slabs)
    deferred_counters = None  # list collecting BatchNorm.num_batches_tracked tensors to bump in one launch (trainer)
    splitk = True           # small-grid / long-K convs (discriminator heads) run split-K with a lent fp32 workspace
    rowconst = True         # decode.0.conv1: the tiled (row-constant) phoneme channels collapse to one row (conv_rowconst)
    # inference: eval-mode BatchNorm / activation / residual folded into the conv epilogue (conv_epi, vca_conv_fwd_tc_epi).
    # Correct (tests/test_gpu_modules.py::test_eval_epilogue_fusion_matches_unfused) but OFF: measured on B200 (B = 64 + flip
    # TTA, profiles/infer_profile_r02.txt) every layer gets slower by more than the saved bn_act pass costs -- e.g. ResNet
    # layer 3: conv 0.448 -> 0.635 ms for a 0.10 ms BatchNorm pass.  The epilogue of these kernels is not hidden behind the
    # MMAs (one accumulator per CTA, or a mainloop of only ~2k cycles per tile), while the separate pass runs at HBM speed.
    fuse_eval_epilogue = False
    # inference: eval-mode BatchNorm folded into conv weights + bias, activation in the epilogue (conv_folded).  Correct
    # (tests/test_gpu_modules.py) but a net LOSS: it removes 1.1 ms of bn_act passes and 1.1 ms of the stem tail for 64 clips, and
    # the convs that carry the bias + activation get 1.5-1.7x slower (their epilogue, not the MMAs, is the critical path of
    # these HBM-bound layers: 41.8 vs 40.1 ms forward, profiles/infer_profile_r02b.txt).  Off by default.
    fold_eval_bn = os.environ.get("VCA_FOLD_EVAL_BN", "0") == "1"
    fuse_stem_pool = True   # stem BatchNorm3d + PReLU + MaxPool3d as one pass over the raw conv output (bn_prelu_maxpool)
    fuse_bn_stats = True    # train-mode BatchNorm statistics come out of the producing conv's epilogue (vca_conv_fwd_tc_stats)
    pair_merge = True       # 32-channel 5x5 convs run as 64-channel 5x3 convs over pixel pairs (_conv5_via_pairs)
    pair_merge_channels = (32,)   # 64 -> 64 as 128 -> 128 over pairs works too but measured no faster (62.2 vs 61.8 ms/step)
    param_grad_streams = ()  # side streams for those accumulations (installed by the Trainer; () = current stream)
    _pg_next = 0
    _pg_used = set()         # side streams with work queued since the last join


cfg = Config()
# A/B switches from the environment, e.g. VCA_CFG="fuse_bn_stats=0,pair_merge=0" (booleans of Config only)
import os as _os
for _kv in filter(None, _os.environ.get("VCA_CFG", "").split(",")):
    _k, _v = _kv.split("=")
    if not isinstance(getattr(Config, _k.strip(), None), bool):
        raise ValueError(f"VCA_CFG: {_k} is not a boolean switch of ops.Config")
    setattr(cfg, _k.strip(), bool(int(_v)))


def set_precision(p: str):
    cfg.dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[p]


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _needed(ctx, i: int, edge: Optional[int] = None) -> bool:
    """True when the engine will actually consume the gradient of input i in this backward pass.  `edge` = index of
    that input in ctx.next_functions, which lists the TENSOR arguments only (None / non-tensor arguments have no edge);
    it defaults to i, which is right whenever no such argument precedes input i."""
    if not ctx.needs_input_grad[i]:
        return False
    if not cfg.skip_unneeded_wgrad:
        return True
    fn = ctx.next_functions[i if edge is None else edge][0]
    if fn is None:
        return False
    try:
        return bool(torch._C._will_engine_execute_node(fn))
    except RuntimeError:
        return True


def _grad_sink(p):
    """The parameter's own fp32 .grad view inside its FlatGroup (trainer.py) when this backward pass may add into it
    directly -- a plain backward (no create_graph) over a leaf that the trainer re-homed.  The wgrad / column-sum
    kernels accumulate with fp32 atomics anyway, so this is exactly AccumulateGrad's `grad += g`, minus the zero-fill
    of a temporary and the add kernel (two launches per parameter per backward)."""
    if p is None or not cfg.fuse_grad_accum or torch.is_grad_enabled() or not getattr(p, "_vca_flat", False):
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
        return None
    return g


class _param_grad_stream:
    """Context: run the enclosed launches on the next of cfg.param_grad_streams (round robin), ordered after everything
    already queued on the current stream; `tensors` are kept alive for that stream (record_stream).  No-op when the
    trainer has not installed side streams."""

    def __init__(self, *tensors):
        self.tensors = tensors
        self.ctx = None

    def __enter__(self):
        streams = cfg.param_grad_streams
        if not streams:
            return self
        cfg._pg_next = (cfg._pg_next + 1) % len(streams)
        st = streams[cfg._pg_next]
        cfg._pg_used.add(st)
        st.wait_stream(torch.cuda.current_stream())
        for t in self.tensors:
            t.record_stream(st)
        self.ctx = torch.cuda.stream(st)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vcagan_b200 ops need CUDA tensors: there is no CPU fallback")


# ------------------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------------------
_pack_cache = {}


def _packed(w: torch.Tensor, dtype: torch.dtype):
    """(wf [taps][Cin][Cout], wd [taps][Cout][Cin]) in `dtype`, cached per parameter version."""
    base = w._base if w._base is not None else w
    cacheable = isinstance(base, torch.nn.Parameter)
    key = (w.data_ptr(), tuple(w.shape), dtype)
    if cacheable:
        ep = getattr(base, "_vca_epoch", None)
        tag = (base._version, ep[0] if ep is not None else 0)
        hit = _pack_cache.get(key)
        if hit is not None and hit[0] == tag and hit[3]() is base:   # same live Parameter object, not a recycled address
            return hit[1], hit[2]
    else:
        # derived weights (space-to-depth / pixel-pair / zero-padded forms: fresh tensors every forward): the forward's pack is
        # remembered on the tensor object itself, which autograd hands back to the dgrad of the same convolution
        pk = getattr(w, "_vca_pk", None)
        if pk is not None and pk[0] == (w._version, dtype):
            return pk[1], pk[2]
    wc = _c(w.detach())
    if wc.dtype != torch.float32:
        wc = wc.float()
    cout, cin = wc.shape[0], wc.shape[1]
    taps = wc.numel() // (cout * cin)
    wf = torch.empty((taps, cin, cout), dtype=dtype, device=w.device)
    wd = torch.empty((taps, cout, cin), dtype=dtype, device=w.device)
    lib().call("vca_pack_conv_weight", BF16 if dtype == torch.bfloat16 else F32, wc, wf, wd, cout, cin, taps)
    if cacheable:
        # the 5th field says whether a PackPlan may refresh this entry in place: the kernel must be able to read the
        # parameter's own storage in [Cout][Cin][taps] order
        direct = wc.data_ptr() == w.data_ptr() and w.dtype == torch.float32
        _pack_cache[key] = (tag, wf, wd, weakref.ref(base), direct, (cout, cin, taps))
    else:
        w._vca_pk = ((w._version, dtype), wf, wd)
    return wf, wd


def clear_pack_cache(keep=()):
    """Forget every packed weight except the entries owned by the given PackPlans."""
    owned = {key for plan in keep if plan is not None for key, _, _ in plan.items}
    for k in [k for k in _pack_cache if k not in owned]:
        del _pack_cache[k]


class PackPlan:
    """One launch that re-packs EVERY cached conv / linear weight of a set of parameters (an optimizer group) into its
    persistent bf16 slabs -- run right after the fused Adam step that changed them, instead of ~75 separate
    vca_pack_conv_weight launches scattered over the next forward pass.  Built from the pack cache after a warm-up step
    (only then is it known which views of which parameters the step actually packs)."""

    def __init__(self, params):
        ids = {id(p) for p in params}
        self.items = []          # (cache key, base parameter)
        rows, off = [], 0
        self.dtype = None
        for key, hit in _pack_cache.items():
            base = hit[3]()
            if base is None or id(base) not in ids or not hit[4]:
                continue
            if self.dtype is None:
                self.dtype = key[2]
            if key[2] != self.dtype:
                continue
            cout, cin, taps = hit[5]
            rows.append([key[0], hit[1].data_ptr(), hit[2].data_ptr(), cout, cin, taps, off, 0])
            off += lib().query("vca_pack_job_ctas", cout, cin, taps)
            self.items.append((key, base, hit[1].data_ptr()))
        self.total = off
        self.table = torch.tensor(rows, dtype=torch.int64, device=params[0].device) if rows else None

    def run(self):
        """Re-pack now (on the current stream) and mark the cache entries as up to date with the parameters."""
        if self.table is None:
            return
        lib().call("vca_pack_conv_weights_batched", BF16 if self.dtype == torch.bfloat16 else F32, self.table, len(self.items), self.total)
        for key, base, slab in self.items:
            hit = _pack_cache.get(key)
            if hit is None or hit[1].data_ptr() != slab:      # re-packed into fresh slabs since (e.g. load_state_dict): not ours
                continue
            ep = getattr(base, "_vca_epoch", None)
            _pack_cache[key] = ((base._version, ep[0] if ep is not None else 0),) + tuple(hit[1:])


def _geom(xshape, wshape, stride, pad) -> Tuple[ConvGeom, tuple]:
    """xshape: (N,[D],H,W,Cin) channels-last; wshape (Cout,Cin,[KD],KH,KW) (param layout)."""
    if len(xshape) == 4:
        N, IH, IW, Cin = xshape; ID = 1
    else:
        N, ID, IH, IW, Cin = xshape
    Cout = wshape[0]
    ks = tuple(wshape[2:])
    if len(ks) == 1:
        ks = (1, 1, ks[0])
    elif len(ks) == 2:
        ks = (1,) + ks
    st = (1,) * (3 - len(stride)) + tuple(stride)
    pd = (0,) * (3 - len(pad)) + tuple(pad)
    assert wshape[1] == Cin, (wshape, xshape)
    OD = (ID + 2 * pd[0] - ks[0]) // st[0] + 1
    OH = (IH + 2 * pd[1] - ks[1]) // st[1] + 1
    OW = (IW + 2 * pd[2] - ks[2]) // st[2] + 1
    g = ConvGeom(N, ID, IH, IW, Cin, OD, OH, OW, Cout, ks[0], ks[1], ks[2], st[0], st[1], st[2], pd[0], pd[1], pd[2])
    oshape = (N, OH, OW, Cout) if len(xshape) == 4 else (N, OD, OH, OW, Cout)
    return g, oshape


def _tc_ok(g: ConvGeom, kind: int, dtype) -> bool:
    return cfg.use_tc and dtype == torch.bfloat16 and lib().query("vca_conv_tc_supported", g, kind) == 1


def _splitk_workspace(g: ConvGeom, kind: int, device):
    """fp32 scratch the library asks for when it wants to run this geometry split-K (small grid, long K), else (None, 0)."""
    nb = lib().query("vca_conv_tc_workspace", g, kind) if cfg.splitk else 0
    if nb <= 0:
        return None, 0
    return torch.empty(nb // 4, dtype=torch.float32, device=device), nb


def _conv_fwd_raw(x, w, bias, stride, pad, stats=None):
    """stats: fp64 [2 * Cout] accumulators of the BatchNorm that follows (only passed when _bn_stats_possible)."""
    g, oshape = _geom(x.shape, w.shape, stride, pad)
    wf, wd = _packed(w, x.dtype)
    y = torch.empty(oshape, dtype=x.dtype, device=x.device)
    b = None if bias is None else _c(bias.detach().float())
    if stats is not None:
        lib().call("vca_conv_fwd_tc_stats", g, x, wd, b, y, stats)
    elif _tc_ok(g, 0, x.dtype):
        ws, nb = _splitk_workspace(g, 0, x.device)
        lib().call("vca_conv_fwd_tc_ws", g, x, wd, b, y, ws, nb)
    else:
        lib().call("vca_conv_fwd_simt", _dt(x), g, x, wf, b, y)
    return y


def _bn_stats_possible(xshape, wshape, stride, pad, dtype) -> bool:
    """True when the forward conv of this geometry runs on the tcgen05 kernel whose epilogue emits BatchNorm statistics
    for free (the weights-stationary persistent kernel; see vca_conv_fwd_tc_stats_supported)."""
    if not cfg.fuse_bn_stats or dtype != torch.bfloat16:
        return False
    g, _ = _geom(xshape, wshape, stride, pad)
    return cfg.use_tc and lib().query("vca_conv_fwd_tc_stats_supported", g) == 1


def _conv_dgrad_raw(dy, w, stride, pad, xshape):
    g, oshape = _geom(xshape, w.shape, stride, pad)
    assert tuple(dy.shape) == tuple(oshape), (dy.shape, oshape)
    wf, wd = _packed(w, dy.dtype)
    dx = torch.empty(xshape, dtype=dy.dtype, device=dy.device)
    if _tc_ok(g, 1, dy.dtype):
        ws, nb = _splitk_workspace(g, 1, dy.device)
        lib().call("vca_conv_dgrad_tc_ws", g, dy, wf, dx, ws, nb)
    else:
        lib().call("vca_conv_dgrad_simt", _dt(dy), g, dy, wd, dx)
    return dx


def _conv_wgrad_raw(x, dy, stride, pad, wshape, out=None, param=None):
    """dw (fp32, parameter layout); every wgrad kernel ADDS into its destination, so `out` may be a live .grad view.
    param: the re-homed parameter `out` is the .grad of.  Its tcgen05 wgrad then goes through the TMA reduce-add epilogue
    (vca_conv_wgrad_tc_tm) into the parameter's TAP-MAJOR slab (FlatGroup.flush_slabs folds it into .grad before the
    gradients are consumed), or straight into .grad for a pointwise conv / linear layer, where the two layouts coincide."""
    g, oshape = _geom(x.shape, wshape, stride, pad)
    assert tuple(dy.shape) == tuple(oshape), (dy.shape, oshape)
    tc = _tc_ok(g, 2, x.dtype)
    taps = g.KD * g.KH * g.KW
    if tc and out is None and cfg.grad_slab and g.Cin % 4 == 0:
        # a fresh gradient tensor: accumulate it TAP-MAJOR (TMA reduce-add epilogue) and hand back the parameter-layout view
        dw_tm = torch.zeros((taps, wshape[0], wshape[1]), dtype=torch.float32, device=x.device)
        lib().call("vca_conv_wgrad_tc_tm", g, dy, x, dw_tm)
        return dw_tm.view(wshape) if taps == 1 else dw_tm.permute(1, 2, 0).reshape(wshape)
    dw = torch.zeros(wshape, dtype=torch.float32, device=x.device) if out is None else out
    if tc:
        slab = None
        if cfg.grad_slab and g.Cin % 4 == 0 and out is not None and param is not None:
            slab = out if taps == 1 else getattr(param, "_vca_slab", None)
        if slab is not None and slab.data_ptr() % 16 == 0:
            lib().call("vca_conv_wgrad_tc_tm", g, dy, x, slab)
            return dw
        lib().call("vca_conv_wgrad_tc", g, dy, x, dw)
    else:
        lib().call("vca_conv_wgrad_simt", _dt(x), g, dy, x, dw)
    return dw


class ConvFn(Function):
    """y = conv(x, w) + bias on channels-last x; w/bias are fp32 parameters in the reference layout."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, zero_bias_grad=False, bn_sums=None):
        _require_cuda(x, w)
        x = _c(x)
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.pad, ctx.has_bias = stride, pad, bias is not None
        # zero_bias_grad: the caller feeds y straight into a train-mode BatchNorm.  BN subtracts the batch mean, so the
        # loss does not depend on this bias and BN's dx sums to zero over the rows: the bias gradient is identically 0
        # (the reference computes ~1e-15 of rounding noise there).  The column-sum pass over dy is not run.
        ctx.zero_bias_grad = bool(zero_bias_grad) and bias is not None
        ctx.bias_shape = None if bias is None else (tuple(bias.shape), bias.device)
        ctx.bias_leaf = bias if (bias is not None and bias.is_leaf) else None   # only consulted by _grad_sink
        return _conv_fwd_raw(x, w, bias, stride, pad, bn_sums)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _c(dy)
        dx = dw = db = None
        w_sink = _grad_sink(w) if _needed(ctx, 1) else None
        want_b = ctx.has_bias and _needed(ctx, 2)
        if want_b and ctx.zero_bias_grad and not torch.is_grad_enabled():
            want_b = False
            if _grad_sink(ctx.bias_leaf) is None:        # plain autograd leaf: hand over an explicit zero (Adam still decays it)
                db = torch.zeros(ctx.bias_shape[0], dtype=torch.float32, device=ctx.bias_shape[1])
        b_sink = _grad_sink(ctx.bias_leaf) if want_b else None
        if w_sink is not None or b_sink is not None:
            # Parameter gradients that land straight in the flat .grad buffer have no consumer in the autograd graph:
            # they run on a side stream and overlap the dgrad chain (joined by Trainer._join_branches before Adam).
            with _param_grad_stream(x, dy):
                if w_sink is not None:
                    _conv_wgrad_raw(x, dy, ctx.stride, ctx.pad, tuple(w.shape), out=w_sink, param=w)   # dw stays None
                if b_sink is not None:
                    _colsum_raw(dy, out=b_sink, accumulate=True)
        if _needed(ctx, 0):
            dx = ConvDgradFn.apply(dy, w, ctx.stride, ctx.pad, tuple(x.shape))
        if _needed(ctx, 1) and w_sink is None:
            dw = ConvWgradFn.apply(x, dy, ctx.stride, ctx.pad, tuple(w.shape))
        if want_b and b_sink is None:
            db = ColSumFn.apply(dy)
        return dx, dw, db, None, None, None, None


class ConvDgradFn(Function):
    """dx = conv_transpose(dy, w); linear in both arguments."""

    @staticmethod
    def forward(ctx, dy, w, stride, pad, xshape):
        dy = _c(dy)
        ctx.save_for_backward(dy, w)
        ctx.stride, ctx.pad, ctx.xshape = stride, pad, xshape
        return _conv_dgrad_raw(dy, w, stride, pad, xshape)

    @staticmethod
    def backward(ctx, ggx):
        dy, w = ctx.saved_tensors
        ggx = _c(ggx)
        g_dy = g_w = None
        if _needed(ctx, 0):
            g_dy = ConvFn.apply(ggx, w, None, ctx.stride, ctx.pad)
        if _needed(ctx, 1):
            # R1 double-backward weight gradient.  When the parameter has a flat .grad sink it must go there through the
            # atomic wgrad kernels as well: handing it to AccumulateGrad would make torch do a plain `grad += g_w`
            # read-modify-write that can interleave with the atomic adds still queued on the side streams.
            sink = _grad_sink(w)
            if sink is not None:
                with _param_grad_stream(ggx, dy):
                    _conv_wgrad_raw(ggx, dy, ctx.stride, ctx.pad, tuple(w.shape), out=sink, param=w)
            else:
                g_w = ConvWgradFn.apply(ggx, dy, ctx.stride, ctx.pad, tuple(w.shape))
        return g_dy, g_w, None, None, None


class ConvWgradFn(Function):
    """dw[co,ci,tap] = sum_pixels dy * shifted x (fp32, parameter layout); linear in both arguments."""

    @staticmethod
    def forward(ctx, x, dy, stride, pad, wshape):
        x, dy = _c(x), _c(dy)
        ctx.save_for_backward(x, dy)
        ctx.stride, ctx.pad, ctx.wshape = stride, pad, wshape
        return _conv_wgrad_raw(x, dy, stride, pad, wshape)

    @staticmethod
    def backward(ctx, ggw):
        x, dy = ctx.saved_tensors
        g_x = g_dy = None
        if _needed(ctx, 0):
            g_x = ConvDgradFn.apply(dy, ggw, ctx.stride, ctx.pad, tuple(x.shape))
        if _needed(ctx, 1):
            g_dy = ConvFn.apply(x, ggw, None, ctx.stride, ctx.pad)
        return g_x, g_dy, None, None, None


def _colsum_raw(x, out=None, accumulate=False):
    x = _c(x)
    C = x.shape[-1]
    if out is None:
        out = torch.empty(C, dtype=torch.float32, device=x.device)
    scratch = torch.empty(C, dtype=torch.float64, device=x.device)
    lib().call("vca_colsum", _dt(x), x, x.numel() // C, C, scratch, out, 1 if accumulate else 0)
    return out


class ColSumFn(Function):
    """fp32 per-channel sum over all leading dims (bias gradient)."""

    @staticmethod
    def forward(ctx, x):
        ctx.xshape, ctx.xdtype = tuple(x.shape), x.dtype
        return _colsum_raw(x)

    @staticmethod
    def backward(ctx, g):
        rows = 1
        for s in ctx.xshape[:-1]:
            rows *= s
        return SpatialBcastFn.apply(g.to(ctx.xdtype).view(1, -1), rows, 1.0).view(ctx.xshape)


class S2DFn(Function):
    """(N,H,W,C) -> (N,H2,W2,4C): zero-bordered space-to-depth, y[n,i,j,(pa*2+pb)*C+c] = x[n,2i+pa-1,2j+pb-1,c]."""

    @staticmethod
    def forward(ctx, x, H2, W2):
        x = _c(x)
        N, H, W, C = x.shape
        ctx.hw = (H, W)
        y = torch.empty((N, H2, W2, 4 * C), dtype=x.dtype, device=x.device)
        lib().call("vca_s2d", _dt(x), x, y, N, H, W, C, H2, W2)
        return y

    @staticmethod
    def backward(ctx, g):
        return D2SFn.apply(g, ctx.hw[0], ctx.hw[1]), None, None


class D2SFn(Function):
    """transpose of S2DFn: (N,H2,W2,4C) -> (N,H,W,C)."""

    @staticmethod
    def forward(ctx, y, H, W):
        y = _c(y)
        N, H2, W2, C4 = y.shape
        ctx.hw2 = (H2, W2)
        x = torch.empty((N, H, W, C4 // 4), dtype=y.dtype, device=y.device)
        lib().call("vca_d2s", _dt(y), y, x, N, H, W, C4 // 4, H2, W2)
        return x

    @staticmethod
    def backward(ctx, g):
        return S2DFn.apply(g, ctx.hw2[0], ctx.hw2[1]), None, None


def _conv_bn(x, w, bias, stride, pad, zero_bias_grad, bn, fold=1):
    """ConvFn, handing the BatchNorm `bn` that consumes the output its batch statistics from the conv epilogue when this
    geometry allows it.  The result then carries `_vca_bn_sums` = (fp64 accumulators, fold); bn_act picks it up."""
    sums = None
    if bn is not None and bn.training and _bn_stats_possible(x.shape, w.shape, stride, pad, x.dtype):
        sums = _bn_sums(bn.running_mean, 2 * w.shape[0], 0)
    y = ConvFn.apply(x, w, bias, stride, pad, zero_bias_grad, sums)
    if sums is not None:
        y._vca_bn_sums = (sums, fold)
    return y


def _conv3x3_s2_via_s2d(x, w, bias, zero_bias_grad=False, bn=None):
    """3x3 / stride 2 / pad 1 conv as a 2x2 stride-1 conv over the space-to-depth input (4C channels), so that the
    forward, dgrad and wgrad all run on the stride-1 tcgen05 kernels (resnet.py:33 with stride 2, generator.py:326).
    W'[co, (pa*2+pb)*C + c, a, b] = w[co, c, 2a+pa, 2b+pb] (zero where the index is 3)."""
    N, H, W, C = x.shape
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    xs = S2DFn.apply(x, OH + 1, OW + 1)
    Cout = w.shape[0]
    wp = torch.nn.functional.pad(w, (0, 1, 0, 1))                      # (Cout, C, 4, 4)   [data movement]
    w2 = wp.view(Cout, C, 2, 2, 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(Cout, 4 * C, 2, 2)
    return _conv_bn(xs, w2, bias, (1, 1), (0, 0), zero_bias_grad, bn)


class PairExpandFn(Function):
    """w (Cout,Cin,KH,5) -> Toeplitz-expanded w2 (2Cout,2Cin,KH,3) of the pixel-pair merged convolution (linear)."""

    @staticmethod
    def forward(ctx, w):
        ctx.wshape = tuple(w.shape)
        ctx.w_leaf = w if w.is_leaf else None
        Cout, Cin, KH, KW = w.shape
        assert KW == 5
        w2 = torch.empty((2 * Cout, 2 * Cin, KH, 3), dtype=torch.float32, device=w.device)
        lib().call("vca_pair_expand_weight", _c(w.detach().float()), w2, Cout, Cin, KH)
        return w2

    @staticmethod
    @once_differentiable
    def backward(ctx, dw2):
        Cout, Cin, KH, _ = ctx.wshape
        sink = _grad_sink(ctx.w_leaf)
        if sink is not None:
            lib().call("vca_pair_contract_wgrad", _c(dw2), sink, Cout, Cin, KH, 1)
            return None
        dw = torch.empty(ctx.wshape, dtype=torch.float32, device=dw2.device)
        lib().call("vca_pair_contract_wgrad", _c(dw2), dw, Cout, Cin, KH, 0)
        return dw


def _conv5_via_pairs(x, w, bias, zero_bias_grad=False, bn=None):
    """(32 -> 32, KH x 5, pad 2) convolution as a (64 -> 64, KH x 3, pad 1) convolution over pixel pairs: the view
    [N,H,W/2,64] of x costs nothing, the weight is Toeplitz-expanded (1.2x the MACs), and the tcgen05 kernels run with
    N = 64 instead of 32 -- their tensor pipe is bound by the A-operand shared-memory read, i.e. proportional to N."""
    N, H, W, C = x.shape
    w2 = PairExpandFn.apply(w)
    b2 = None if bias is None else torch.cat([bias, bias])
    y2 = _conv_bn(x.view(N, H, W // 2, 2 * C), w2, b2, (1, 1), (w.shape[2] // 2, 1), zero_bias_grad, bn, fold=2)
    y = y2.view(N, H, W, w.shape[0])
    if hasattr(y2, "_vca_bn_sums"):
        y._vca_bn_sums = y2._vca_bn_sums        # statistics columns are [pair position][channel]: folded by the finalize kernel
    return y


def conv(x, w, bias=None, stride=(1, 1), pad=(0, 0), zero_bias_grad=False, bn=None):
    """zero_bias_grad: see ConvFn.forward (the output goes straight into a train-mode BatchNorm).  bn: that BatchNorm
    (module holding running_mean / training), so that its batch statistics can come out of the conv epilogue."""
    stride, pad = tuple(stride), tuple(pad)
    zb = bool(zero_bias_grad) and bias is not None
    if (cfg.pair_merge and x.dim() == 4 and x.dtype == torch.bfloat16 and cfg.use_tc and stride == (1, 1)
            and w.dim() == 4 and tuple(w.shape[2:]) == (5, 5) and w.shape[0] == w.shape[1] and w.shape[1] in cfg.pair_merge_channels
            and pad == (2, 2) and x.shape[-1] == w.shape[1] and x.shape[2] % 2 == 0 and x.is_contiguous()):
        return _conv5_via_pairs(x, w, bias, zb, bn)
    if (x.dim() == 4 and x.dtype == torch.bfloat16 and cfg.use_tc and all(s == 1 for s in stride) and w.shape[0] % 8 != 0
            and w.shape[0] >= 64 and x.shape[-1] % 8 == 0):
        # Postnet's 256 -> 321 projection (generator.py:185): the tcgen05 kernels want Cout % 8 == 0 (16-byte rows for the
        # TMA maps of dgrad / wgrad), so run it with 7 zero output channels appended and hand back the first 321
        padc = (-w.shape[0]) % 8
        w2 = torch.cat([w, w.new_zeros((padc,) + tuple(w.shape[1:]))], 0)
        b2 = None if bias is None else torch.cat([bias, bias.new_zeros(padc)])
        return ConvFn.apply(x, w2, b2, stride, pad)[..., :w.shape[0]]
    if x.dim() == 4 and stride == (2, 2) and x.dtype == torch.bfloat16 and cfg.use_tc and x.shape[-1] % 8 == 0:
        k = tuple(w.shape[2:])
        if k == (3, 3) and pad == (1, 1) and x.shape[-1] >= 8:
            return _conv3x3_s2_via_s2d(x, w, bias, zb, bn)
        if k == (1, 1) and pad == (0, 0):
            return _conv_bn(x[:, ::2, ::2, :].contiguous(), w, bias, (1, 1), (0, 0), zb, bn)   # slicing = data movement
    return _conv_bn(x, w, bias, stride, pad, zb, bn)


# ---- inference: eval-mode BatchNorm / activation / residual folded into the conv epilogue (no autograd) -------------------
def epi_ok(x) -> bool:
    """The fused inference epilogues apply: bf16 tcgen05 path, and nothing asks for gradients."""
    return cfg.fuse_eval_epilogue and cfg.use_tc and x.dtype == torch.bfloat16 and not torch.is_grad_enabled()


def _fold(bn, bias, C, device, out_scale=1.0):
    """(scale, shift) fp32 [C] of `y * scale + shift` == out_scale * BN_eval(y + bias)   (bn may be None)."""
    if bn is None:
        if bias is None and out_scale == 1.0:
            return None, None
        scale = torch.full((C,), float(out_scale), dtype=torch.float32, device=device) if out_scale != 1.0 else None
        shift = None if bias is None else (bias.detach().float() * out_scale if out_scale != 1.0 else bias.detach().float())
        return scale, shift
    scale = torch.empty(C, dtype=torch.float32, device=device)
    shift = torch.empty(C, dtype=torch.float32, device=device)
    lib().call("vca_bn_fold", bn.running_mean, bn.running_var, bn.weight.detach(), bn.bias.detach(),
               None if bias is None else bias.detach().float(), C, float(bn.eps), scale, shift)
    if out_scale != 1.0:
        scale, shift = scale * out_scale, shift * out_scale
    return scale, shift


def _conv_epi_raw(x, w, scale, shift, res, res_scale, act, slope, prelu_w, pad):
    g, oshape = _geom(x.shape, w.shape, (1, 1), pad)
    if not _tc_ok(g, 0, x.dtype):
        return None
    _, wd = _packed(w, x.dtype)
    y = torch.empty(oshape, dtype=x.dtype, device=x.device)
    lib().call("vca_conv_fwd_tc_epi", g, _c(x), wd, scale, shift, None if res is None else _c(res), float(res_scale), int(act), float(slope),
               None if prelu_w is None else prelu_w.detach().float(), y)
    return y


def conv_epi(x, w, bias=None, stride=(1, 1), pad=(0, 0), bn=None, act=ACT_NONE, slope=0.0, prelu_w=None, res=None, res_scale=1.0,
             out_scale=1.0):
    """Inference only:  act(out_scale * BN_eval(conv(x, w) + bias) + res_scale * res)  in ONE kernel (the conv's epilogue);
    returns None when the geometry has no tcgen05 route (the caller then runs the separate kernels).  Same routes as
    conv(): pixel-pair merge for 32 channels, space-to-depth / slicing for stride 2."""
    stride, pad = tuple(stride), tuple(pad)
    Cout = w.shape[0]
    scale, shift = _fold(bn, bias, Cout, x.device, out_scale)
    if (cfg.pair_merge and x.dim() == 4 and stride == (1, 1) and w.dim() == 4 and tuple(w.shape[2:]) == (5, 5) and w.shape[0] == w.shape[1]
            and w.shape[1] in cfg.pair_merge_channels and pad == (2, 2) and x.shape[-1] == w.shape[1] and x.shape[2] % 2 == 0
            and x.is_contiguous()):
        N, H, W, C = x.shape
        w2 = PairExpandFn.apply(w)
        dup = lambda t: None if t is None else torch.cat([t, t])      # noqa: E731  (output channels are [pair position][channel])
        y2 = _conv_epi_raw(x.view(N, H, W // 2, 2 * C), w2, dup(scale), dup(shift), None if res is None else _c(res).view(N, H, W // 2, 2 * Cout),
                           res_scale, act, slope, dup(None if prelu_w is None else prelu_w.detach().float()), (w.shape[2] // 2, 1))
        return None if y2 is None else y2.view(N, H, W, Cout)
    if stride == (2, 2) and x.dim() == 4 and x.shape[-1] % 8 == 0:
        k = tuple(w.shape[2:])
        if k == (3, 3) and pad == (1, 1):
            N, H, W, C = x.shape
            OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            xs = S2DFn.apply(x, OH + 1, OW + 1)
            wp = torch.nn.functional.pad(w, (0, 1, 0, 1))
            w2 = wp.view(Cout, C, 2, 2, 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(Cout, 4 * C, 2, 2)
            return _conv_epi_raw(xs, w2, scale, shift, res, res_scale, act, slope, prelu_w, (0, 0))
        if k == (1, 1) and pad == (0, 0):
            return _conv_epi_raw(x[:, ::2, ::2, :].contiguous(), w, scale, shift, res, res_scale, act, slope, prelu_w, (0, 0))
        return None
    if any(s_ != 1 for s_ in stride) or w.shape[0] % 8 != 0:
        return None
    return _conv_epi_raw(x, w, scale, shift, res, res_scale, act, slope, prelu_w, pad)


# ---- inference: eval-mode BatchNorm folded into the conv's WEIGHTS and bias, activation in the epilogue --------------------
# (SURVEY appendix A #20.)  BN_eval(conv(x, w) + b) = conv(x, w * s) + (b - mean) * s + beta with s = gamma / sqrt(var + eps): the
# scale goes into the packed bf16 weights, the shift is the conv's bias, and the LeakyReLU / PReLU that follows is one select
# per value in the weights-stationary kernel's epilogue (vca_conv_fwd_tc_act) -- unlike the scale/shift/residual epilogue
# of vca_conv_fwd_tc_epi (measured slower than the pass it removed) it adds no loads and keeps the stacked MMAs and the
# TMA-store epilogue.  The folded, packed weights are cached per (weight, BatchNorm) version.
_fold_cache = {}
# Bumped whenever something rewrites parameters or BatchNorm running statistics THROUGH RAW POINTERS (train-mode BatchNorm
# kernels, the fused Adam kernel): torch's tensor versions do not see those writes, so the cache of folded weights keys on it.
_train_touch = [0]


def fold_ok(x) -> bool:
    return cfg.fold_eval_bn and cfg.use_tc and x.dtype == torch.bfloat16 and not torch.is_grad_enabled()


def _folded_weights(w, bias, bn, act, slope, prelu_w, expand_pairs):
    vers = (_train_touch[0], w._version, bn.running_mean._version, bn.running_var._version, bn.weight._version, bn.bias._version,
            -1 if bias is None else bias._version, -1 if prelu_w is None else prelu_w._version, act, float(slope), expand_pairs)
    key = (w.data_ptr(), bn.running_mean.data_ptr())
    hit = _fold_cache.get(key)
    if hit is not None and hit[0] == vers:
        return hit[1:]
    Cout = w.shape[0]
    scale, shift = _fold(bn, bias, Cout, w.device)
    wf32 = (w.detach().float() * scale.view(-1, *([1] * (w.dim() - 1)))).contiguous()
    if act == ACT_PRELU:
        sl = prelu_w.detach().float().clone()
    else:
        sl = torch.full((Cout,), float(slope) if act == ACT_LRELU else (0.0 if act == ACT_RELU else 1.0), dtype=torch.float32, device=w.device)
    if expand_pairs:             # pixel-pair merged 32-channel 5x5 layers: Toeplitz-expanded weight, channels = [pair position][c]
        KH = wf32.shape[2]
        w2 = torch.empty((2 * Cout, 2 * wf32.shape[1], KH, 3), dtype=torch.float32, device=w.device)
        lib().call("vca_pair_expand_weight", wf32, w2, Cout, wf32.shape[1], KH)
        wf32, shift, sl = w2, torch.cat([shift, shift]), torch.cat([sl, sl])
    co, ci = wf32.shape[0], wf32.shape[1]
    taps = wf32.numel() // (co * ci)
    wf = torch.empty((taps, ci, co), dtype=torch.bfloat16, device=w.device)
    wd = torch.empty((taps, co, ci), dtype=torch.bfloat16, device=w.device)
    lib().call("vca_pack_conv_weight", BF16, wf32, wf, wd, co, ci, taps)
    val = (wd, shift.contiguous(), sl.contiguous(), tuple(wf32.shape))
    _fold_cache[key] = (vers,) + val
    return val


def conv_folded(x, w, bias, pad, bn, act=ACT_NONE, slope=0.0, prelu_w=None):
    """Inference only: act(BN_eval(conv(x, w) + bias)) with the BatchNorm folded into weights + bias and the activation in the
    conv's epilogue; stride-1 2-D convs on the weights-stationary tcgen05 kernel (incl. the pixel-pair merged 32-channel
    5x5 layers).  Returns None where that kernel does not take the geometry (the caller keeps the separate pass)."""
    pad = tuple(pad)
    if x.dim() != 4 or w.dim() != 4 or not x.is_contiguous():
        return None
    pairs = (cfg.pair_merge and tuple(w.shape[2:]) == (5, 5) and w.shape[0] == w.shape[1] and w.shape[1] in cfg.pair_merge_channels
             and pad == (2, 2) and x.shape[-1] == w.shape[1] and x.shape[2] % 2 == 0)
    N, H, W, C = x.shape
    xs = x.view(N, H, W // 2, 2 * C) if pairs else x
    wshape = (2 * w.shape[0], 2 * w.shape[1], w.shape[2], 3) if pairs else tuple(w.shape)
    g, oshape = _geom(xs.shape, wshape, (1, 1), (w.shape[2] // 2, 1) if pairs else pad)
    if wshape[0] % 8 or not _tc_ok(g, 0, x.dtype) or lib().query("vca_conv_fwd_tc_act_supported", g) != 1:
        return None
    wd, shift, sl, _ = _folded_weights(w, bias, bn, act, slope, prelu_w, pairs)
    y = torch.empty(oshape, dtype=x.dtype, device=x.device)
    lib().call("vca_conv_fwd_tc_act", g, xs, wd, shift, sl, y)
    return y.view(N, oshape[1], W, w.shape[0]) if pairs else y


_stem_fold_w = {}


def stem_conv(vid, w, bn=None, fold=None):
    """Visual front-end stem Conv3d(1,64,(5,7,7),(1,2,2),(2,3,3)) (visual_front.py:11) on the tcgen05 path:
    a gather kernel unrolls the 7x7 spatial taps of the single input channel into 64 channels (49 used), the
    remaining 5-tap temporal convolution is a (5,1) stride-1 conv over (T, 56*56).  vid: (B,1,T,H,W) fp32 -> (B,T,OH,OW,64)."""
    B, _, T, H, W = vid.shape
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    vid = _c(vid)
    cols = torch.empty((B * T, OH, OW, 64), dtype=cfg.dtype, device=vid.device)
    lib().call("vca_stem_im2col", _dt(vid), BF16 if cfg.dtype == torch.bfloat16 else F32, vid, cols, B * T, H, W)
    Cout = w.shape[0]
    w2 = torch.nn.functional.pad(w.view(Cout, 5, 49).permute(0, 2, 1), (0, 0, 0, 15)).unsqueeze(-1)   # (Cout,64,5,1)
    if fold is not None and fold_ok(cols):
        # inference: BatchNorm3d + PReLU ride in the (5,1) conv (weights scaled, shift as bias, slope in the epilogue)
        fbn, fprelu = fold
        key_w = _stem_fold_w.get(w.data_ptr())
        if key_w is None or key_w[0] != w._version:
            key_w = (w._version, w2.contiguous())
            _stem_fold_w[w.data_ptr()] = key_w                    # a stable tensor, so that conv_folded's cache can key on it
        yf = conv_folded(cols.view(B, T, OH * OW, 64), key_w[1], None, (2, 0), fbn, ACT_PRELU, 0.0, fprelu)
        if yf is not None:
            y = yf.view(B, T, OH, OW, Cout)
            y._vca_activated = True
            return y
    y4 = _conv_bn(cols.view(B, T, OH * OW, 64), w2, None, (1, 1), (2, 0), False, bn)
    y = y4.view(B, T, OH, OW, Cout)
    if hasattr(y4, "_vca_bn_sums"):
        y._vca_bn_sums = y4._vca_bn_sums
    return y


def linear(x, w, bias=None):
    """x (..., K) -> (..., N) with w (N, K) (nn.Linear); runs as a 1x1 convolution over rows."""
    lead = x.shape[:-1]
    x4 = x.reshape(-1, 1, 1, x.shape[-1])
    y = ConvFn.apply(x4, w.view(w.shape[0], w.shape[1], 1, 1), bias, (1, 1), (0, 0))
    return y.view(*lead, w.shape[0])


# ------------------------------------------------------------------------------------------------------------
# BatchNorm (+residual) + activation
# ------------------------------------------------------------------------------------------------------------
_bn_scratch = {}


def _bn_sums(key_tensor, n, slot):
    """Persistent fp64 scratch of one BatchNorm (keyed by its running_mean storage), zero when a kernel sequence starts
    and left zeroed by its last kernel -- so the per-call memset launches disappear (56 x 2 per step)."""
    key = (key_tensor.data_ptr(), n, slot)
    t = _bn_scratch.get(key)
    if t is None or t.device != key_tensor.device:
        t = torch.zeros(n, dtype=torch.float64, device=key_tensor.device)
        _bn_scratch[key] = t
    return t


class BNActFn(Function):
    """y = act(BN(x) [+ res]).  Training mode uses batch statistics over all leading dims and updates the running
    buffers in place (momentum 0.1, unbiased variance), eval mode uses the running statistics."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, running_mean, running_var, prelu_w, training, act, slope, eps, momentum,
                pre_sums=None, fold=1):
        _require_cuda(x)
        if training:
            _train_touch[0] += 1
        x = _c(x)
        res = None if res is None else _c(res)
        C = x.shape[-1]
        R = x.numel() // C
        dev = x.device
        mean = torch.empty(C, dtype=torch.float32, device=dev)
        invstd = torch.empty(C, dtype=torch.float32, device=dev)
        if training and pre_sums is not None:
            # the producing convolution's epilogue already accumulated sum / sum of squares: no pass over x
            lib().call("vca_bn_finalize_stats", pre_sums, R, C, fold, eps, momentum, mean, invstd, running_mean, running_var)
        elif training:
            lib().call("vca_bn_stats", _dt(x), x, R, C, eps, momentum, _bn_sums(running_mean, 2 * C, 0), 1, mean, invstd,
                       running_mean, running_var)
        else:
            lib().call("vca_bn_eval_stats", running_mean, running_var, C, eps, mean, invstd)
        y = torch.empty_like(x)
        g32, b32 = gamma.detach(), beta.detach()
        pw = None if prelu_w is None else prelu_w.detach()
        lib().call("vca_bn_act_fwd", _dt(x), x, res, y, R, C, mean, invstd, g32, b32, act, slope, pw)
        ctx.save_for_backward(x, res, gamma, beta, prelu_w, mean, invstd)
        ctx.cfg = (training, act, slope)
        ctx.key = running_mean
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, res, gamma, beta, prelu_w, mean, invstd = ctx.saved_tensors
        training, act, slope = ctx.cfg
        dy = _c(dy)
        C = x.shape[-1]
        R = x.numel() // C
        dev = x.device
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if res is not None else None
        # parameter gradients straight into the flat .grad views when all of them are trainer-owned leaves
        e0 = 1 + (res is not None)       # edge index of gamma: x [, res], gamma, beta, running_mean, running_var [, prelu_w]
        wanted = [_needed(ctx, 2, e0), _needed(ctx, 3, e0 + 1)] + ([_needed(ctx, 6, e0 + 4)] if prelu_w is not None else [])
        sinks = [_grad_sink(p) for p in (gamma, beta) + ((prelu_w,) if prelu_w is not None else ())]
        fused = all(wanted) and all(t is not None for t in sinks)
        if not any(wanted):          # e.g. the sync discriminator's BatchNorms in the G phase: no parameter gradients at all
            dgamma = dbeta = dprelu = None
        elif fused:
            dgamma, dbeta = sinks[0], sinks[1]
            dprelu = sinks[2] if prelu_w is not None else None
        else:
            dgamma = torch.empty(C, dtype=torch.float32, device=dev)
            dbeta = torch.empty(C, dtype=torch.float32, device=dev)
            dprelu = torch.empty(C, dtype=torch.float32, device=dev) if prelu_w is not None else None
        lib().call("vca_bn_act_bwd", _dt(x), dy, x, res, dx, dres, R, C, mean, invstd, gamma.detach(), beta.detach(), act, slope,
                   None if prelu_w is None else prelu_w.detach(), 1 if training else 0, _bn_sums(ctx.key, 3 * C, 1), dgamma, dbeta,
                   dprelu, 1 | (2 if fused else 0))
        if fused:
            return (dx, dres) + (None,) * 12
        return (dx, dres, dgamma, dbeta, None, None, dprelu) + (None,) * 7


def bn_act(x, bn: torch.nn.Module, act=ACT_NONE, slope=0.0, prelu_w=None, res=None):
    """`bn` is any module holding weight/bias/running_mean/running_var/num_batches_tracked/eps/momentum/training."""
    training = bn.training
    if training and bn.num_batches_tracked is not None:
        if cfg.deferred_counters is not None:       # the trainer bumps all of them with one foreach add per step
            cfg.deferred_counters.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked += 1
    pre = getattr(x, "_vca_bn_sums", None) if training else None
    if pre is not None and pre[0].data_ptr() != _bn_sums(bn.running_mean, pre[0].numel(), 0).data_ptr():
        raise RuntimeError("BatchNorm statistics were accumulated for a different BatchNorm than the one consuming the tensor")
    return BNActFn.apply(x, res, bn.weight, bn.bias, bn.running_mean, bn.running_var, prelu_w, training, act, float(slope),
                         float(bn.eps), float(bn.momentum), None if pre is None else pre[0], 1 if pre is None else pre[1])


class BNPReluMaxPoolFn(Function):
    """maxpool3x3s2(prelu(BN(x))) for the visual front-end stem (visual_front.py:12-14) in ONE pass over the raw conv
    output (csrc/bn.cu: bn_prelu_maxpool_*): the 963 MB activated tensor is never written.  x (NF,H,W,C) bf16."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, prelu_w, training, eps, momentum, pre_sums=None, fold=1):
        if training:
            _train_touch[0] += 1
        _require_cuda(x)
        x = _c(x)
        NF, H, W, C = x.shape
        R = NF * H * W
        dev = x.device
        mean = torch.empty(C, dtype=torch.float32, device=dev)
        invstd = torch.empty(C, dtype=torch.float32, device=dev)
        if training and pre_sums is not None:
            lib().call("vca_bn_finalize_stats", pre_sums, R, C, fold, eps, momentum, mean, invstd, running_mean, running_var)
        elif training:
            lib().call("vca_bn_stats", _dt(x), x, R, C, eps, momentum, _bn_sums(running_mean, 2 * C, 0), 1, mean, invstd,
                       running_mean, running_var)
        else:
            lib().call("vca_bn_eval_stats", running_mean, running_var, C, eps, mean, invstd)
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((NF, OH, OW, C), dtype=x.dtype, device=dev)
        xmax = torch.empty_like(y)
        idx = torch.empty((NF, OH, OW, C), dtype=torch.uint8, device=dev)
        lib().call("vca_bn_prelu_maxpool_fwd", x, y, idx, xmax, NF, H, W, C, mean, invstd, gamma.detach(), beta.detach(), prelu_w.detach())
        ctx.save_for_backward(x, idx, xmax, gamma, beta, prelu_w, mean, invstd)
        ctx.training, ctx.key = training, running_mean
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, idx, xmax, gamma, beta, prelu_w, mean, invstd = ctx.saved_tensors
        dy = _c(dy)
        NF, H, W, C = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        wanted = [_needed(ctx, 1), _needed(ctx, 2), _needed(ctx, 5, 5)]
        sinks = [_grad_sink(p) for p in (gamma, beta, prelu_w)]
        fused = all(wanted) and all(t is not None for t in sinks)
        if fused:
            dgamma, dbeta, dprelu = sinks
        elif any(wanted):
            dgamma, dbeta, dprelu = (torch.empty(C, dtype=torch.float32, device=x.device) for _ in range(3))
        else:
            dgamma = dbeta = dprelu = None
        if dx is None:
            dx = torch.empty_like(x)          # the kernel always writes it (the stem conv's wgrad is its only consumer)
        lib().call("vca_bn_prelu_maxpool_bwd", dy, idx, xmax, x, dx, NF, H, W, C, mean, invstd, gamma.detach(), beta.detach(),
                   prelu_w.detach(), 1 if ctx.training else 0, _bn_sums(ctx.key, 3 * C, 1), dgamma, dbeta, dprelu, 1 | (2 if fused else 0))
        if fused or not any(wanted):
            return (dx,) + (None,) * 10
        return (dx, dgamma, dbeta, None, None, dprelu) + (None,) * 5


def bn_prelu_maxpool_supported(x, C) -> bool:
    cv = C // 8
    return cfg.fuse_stem_pool and x.dtype == torch.bfloat16 and C % 8 == 0 and cv <= 256 and (cv & (cv - 1)) == 0


def bn_prelu_maxpool(x, bn: torch.nn.Module, prelu_w):
    """x (NF,H,W,C) raw stem-conv output -> (NF,OH,OW,C); `bn` as for bn_act."""
    training = bn.training
    if training and bn.num_batches_tracked is not None:
        if cfg.deferred_counters is not None:
            cfg.deferred_counters.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked += 1
    pre = getattr(x, "_vca_bn_sums", None) if training else None
    return BNPReluMaxPoolFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, prelu_w, training, float(bn.eps),
                                  float(bn.momentum), None if pre is None else pre[0], 1 if pre is None else pre[1])


def flush_deferred_counters():
    """num_batches_tracked += (number of train-mode forwards since the last flush), one fused launch per distinct count."""
    lst, cfg.deferred_counters = cfg.deferred_counters, None
    if not lst:
        return
    count, by_ptr = {}, {}
    for t in lst:
        count[t.data_ptr()] = count.get(t.data_ptr(), 0) + 1
        by_ptr[t.data_ptr()] = t
    for k in sorted(set(count.values())):
        torch._foreach_add_([by_ptr[p] for p, c in count.items() if c == k], k)


# ------------------------------------------------------------------------------------------------------------
# element-wise (double-differentiable where the discriminator path needs it)
# ------------------------------------------------------------------------------------------------------------
class LReluFn(Function):
    @staticmethod
    def forward(ctx, x, slope):
        x = _c(x)
        ctx.save_for_backward(x)
        ctx.slope = slope
        y = torch.empty_like(x)
        lib().call("vca_lrelu_fwd", _dt(x), x, y, x.numel(), slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return LReluBwdFn.apply(dy, x, ctx.slope), None


class LReluBwdFn(Function):
    """dx = dy * (x > 0 ? 1 : slope): linear in dy, zero derivative in x almost everywhere."""

    @staticmethod
    def forward(ctx, dy, x, slope):
        dy = _c(dy)
        ctx.save_for_backward(x)
        ctx.slope = slope
        dx = torch.empty_like(dy)
        lib().call("vca_lrelu_bwd", _dt(dy), dy, x, dx, dy.numel(), slope)
        return dx

    @staticmethod
    def backward(ctx, gg):
        (x,) = ctx.saved_tensors
        return LReluBwdFn.apply(gg, x, ctx.slope), None, None


def lrelu(x, slope=0.2):
    return LReluFn.apply(x, float(slope))


class TanhFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        lib().call("vca_tanh_fwd", _dt(x), x, y, x.numel())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        lib().call("vca_tanh_bwd", _dt(dy), dy, y, dx, dy.numel())
        return dx


def tanh(x):
    return TanhFn.apply(x)


class AxpbyFn(Function):
    """out = alpha*a + beta*b  (b optional)."""

    @staticmethod
    def forward(ctx, a, b, alpha, beta):
        a = _c(a)
        b = None if b is None else _c(b)
        ctx.alpha, ctx.beta, ctx.has_b = alpha, beta, b is not None
        out = torch.empty_like(a)
        lib().call("vca_axpby", _dt(a), a, b, out, a.numel(), alpha, beta)
        return out

    @staticmethod
    def backward(ctx, g):
        ga = AxpbyFn.apply(g, None, ctx.alpha, 0.0) if ctx.needs_input_grad[0] else None
        if ctx.has_b and ctx.needs_input_grad[1]:
            # (r + s) / sqrt(2): both branches receive the SAME scaled gradient -- one kernel, one tensor
            gb = ga if (ga is not None and ctx.alpha == ctx.beta) else AxpbyFn.apply(g, None, ctx.beta, 0.0)
        else:
            gb = None
        return ga, gb, None, None


def add_scale(a, b, s):
    return AxpbyFn.apply(a, b, float(s), float(s))


class RowTapsFn(Function):
    """y[b,f,t,c] = yn[b,f,t,c] + sum_{kh: 0 <= f+kh-ph < F} R[b,0,t,kh*C+c]: adds the collapsed form of a KH x KW
    convolution over channels that are constant along F (see csrc/row_taps.cu) to the convolution of the others."""

    @staticmethod
    def forward(ctx, yn, R, KH, ph):
        yn, R = _c(yn), _c(R)
        B, F_, T, C = yn.shape
        assert tuple(R.shape) == (B, 1, T, KH * C) and R.dtype == yn.dtype, (R.shape, yn.shape)
        ctx.dims = (B, F_, T, C, KH, ph)
        y = torch.empty_like(yn)
        lib().call("vca_row_taps_fwd", _dt(yn), yn, R, y, B, F_, T, C, KH, ph)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        B, F_, T, C, KH, ph = ctx.dims
        dy = _c(dy)
        dR = None
        if ctx.needs_input_grad[1]:
            dR = torch.empty((B, 1, T, KH * C), dtype=dy.dtype, device=dy.device)
            lib().call("vca_row_taps_bwd", _dt(dy), dy, dR, B, F_, T, C, KH, ph)
        return (dy if ctx.needs_input_grad[0] else None), dR, None, None


def conv_rowconst(x, nc, w, bias, pad, zero_bias_grad=False):
    """conv(x, w, bias, stride 1, pad) for x (B,F,T,C) whose first `nc` channels are constant along F: those channels
    go through ONE row (kh folded into the output channels of a 1 x KW conv) + the row-tap combine, the remaining
    channels through the ordinary conv.  Exact up to summation order; F/KH-fold fewer MACs on the constant part."""
    Cout, C, KH, KW = w.shape
    rc = x[:, :1, :, :nc].contiguous()                                             # the single distinct row
    rn = x[..., nc:].contiguous()
    w_rows = w[:, :nc].permute(2, 0, 1, 3).reshape(KH * Cout, nc, 1, KW)            # R[kh] = conv1d(row, w[:, :nc, kh, :])
    R = conv(rc, w_rows, None, (1, 1), (0, pad[1]))
    yn = conv(rn, w[:, nc:].contiguous(), bias, (1, 1), pad, zero_bias_grad)
    return RowTapsFn.apply(yn, R, KH, pad[0])


def scale(a, s):
    return AxpbyFn.apply(a, None, float(s), 0.0)


class CastFn(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        x = _c(x)
        ctx.src = x.dtype
        if x.dtype == dtype:
            return x
        y = torch.empty(x.shape, dtype=dtype, device=x.device)
        lib().call("vca_cast", _dt(x), BF16 if dtype == torch.bfloat16 else F32, x, y, x.numel())
        return y

    @staticmethod
    def backward(ctx, g):
        return CastFn.apply(g, ctx.src), None


def cast(x, dtype):
    return x if x.dtype == dtype else CastFn.apply(x, dtype)


class MulFn(Function):
    """y = x * mask (mask is a constant: dropout)."""

    @staticmethod
    def forward(ctx, x, mask):
        x = _c(x)
        ctx.save_for_backward(mask)
        y = torch.empty_like(x)
        lib().call("vca_mul", _dt(x), x, mask, y, x.numel())
        return y

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return MulFn.apply(g, mask), None


_rng_state = {"seed": 0x5EED, "rank": 0, "ctr": {}}


def manual_seed(seed: int):
    """Seed of the device Philox streams; the stream position is a device-resident counter per GPU, so draws stay
    fresh when the step is replayed from a CUDA graph."""
    _rng_state["seed"] = int(seed)
    for c in _rng_state["ctr"].values():
        c.zero_()


def set_rng_rank(rank: int):
    """Data-parallel runs: every rank draws from its own Philox key (seed + rank * odd 64-bit constant), so that the
    shards of a global batch get different dropout masks / generator noise under the same manual_seed()."""
    _rng_state["rank"] = int(rank)


def _rng_key() -> int:
    return (_rng_state["seed"] + _rng_state["rank"] * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF


def _rng(shape, dtype, device, mode, param=0.0):
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    ctr = _rng_state["ctr"].get(key)
    if ctr is None:
        ctr = _rng_state["ctr"][key] = torch.zeros(1, dtype=torch.int64, device=device)
    out = torch.empty(shape, dtype=dtype, device=device)
    lib().call("vca_rng_dev", BF16 if dtype == torch.bfloat16 else F32, out, out.numel(), _rng_key(), ctr, mode, float(param))
    return out


def randn(shape, dtype, device):
    """Device-side Philox N(0,1) (replaces the host torch.randn + H2D of generator.py:248)."""
    return _rng(shape, dtype, device, 0)


def dropout(x, p: float, training: bool, mask: Optional[torch.Tensor] = None):
    """nn.Dropout semantics; `mask` (already scaled by 1/(1-p)) may be injected for parity tests."""
    if mask is None:
        if not training or p <= 0.0:
            return x
        mask = _rng(x.shape, x.dtype, x.device, 1, p)
    return MulFn.apply(x, _c(mask.to(x.dtype)))


# ------------------------------------------------------------------------------------------------------------
# pooling / resampling
# ------------------------------------------------------------------------------------------------------------
class Pool2x2Fn(Function):
    """(N,H,W,C) -> (N,H//2,W//2,C): scale * 2x2 block sums (scale 0.25 = F.avg_pool2d(x, 2))."""

    @staticmethod
    def forward(ctx, x, scale_):
        x = _c(x)
        N, H, W, C = x.shape
        ctx.hw, ctx.scale = (H, W), scale_
        y = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
        lib().call("vca_pool2x2_sum", _dt(x), x, y, N, H, W, C, scale_)
        return y

    @staticmethod
    def backward(ctx, g):
        return Expand2x2Fn.apply(g, ctx.hw, ctx.scale), None


class Expand2x2Fn(Function):
    """(N,h,w,C) -> (N,H,W,C) with y[i,j] = scale * x[i//2, j//2] (zero on an odd trailing row/col).
    scale 1 with (H,W) = (2h,2w) is F.interpolate(scale_factor=2, mode='nearest')."""

    @staticmethod
    def forward(ctx, x, hw, scale_):
        x = _c(x)
        N, h, w, C = x.shape
        H, W = hw
        ctx.scale = scale_
        y = torch.empty((N, H, W, C), dtype=x.dtype, device=x.device)
        lib().call("vca_expand2x2", _dt(x), x, y, N, h, w, C, H, W, scale_)
        return y

    @staticmethod
    def backward(ctx, g):
        return Pool2x2Fn.apply(g, ctx.scale), None, None


def avg_pool2(x):
    return Pool2x2Fn.apply(x, 0.25)


def upsample2(x):
    return Expand2x2Fn.apply(x, (2 * x.shape[1], 2 * x.shape[2]), 1.0)


class MaxPool3x3s2Fn(Function):
    """Per-frame 3x3/stride 2/pad 1 max pool on (NF,H,W,C)  (MaxPool3d((1,3,3),(1,2,2),(0,1,1)), visual_front.py:14)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        NF, H, W, C = x.shape
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((NF, OH, OW, C), dtype=x.dtype, device=x.device)
        idx = torch.empty((NF, OH, OW, C), dtype=torch.uint8, device=x.device)
        lib().call("vca_maxpool3x3s2_fwd", _dt(x), x, y, idx, NF, H, W, C)
        ctx.save_for_backward(idx)
        ctx.xshape = (NF, H, W, C)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        NF, H, W, C = ctx.xshape
        dy = _c(dy)
        dx = torch.empty(ctx.xshape, dtype=dy.dtype, device=dy.device)
        lib().call("vca_maxpool3x3s2_bwd", _dt(dy), dy, idx, dx, NF, H, W, C)
        return dx


def maxpool3x3s2(x):
    return MaxPool3x3s2Fn.apply(x)


class SpatialSumFn(Function):
    """(N,P,C) -> (N,C): scale * sum over P."""

    @staticmethod
    def forward(ctx, x, scale_):
        x = _c(x)
        N, P, C = x.shape
        ctx.P, ctx.scale = P, scale_
        y = torch.empty((N, C), dtype=x.dtype, device=x.device)
        lib().call("vca_spatial_sum", _dt(x), x, y, N, P, C, scale_)
        return y

    @staticmethod
    def backward(ctx, g):
        return SpatialBcastFn.apply(g, ctx.P, ctx.scale), None


class SpatialBcastFn(Function):
    """(N,C) -> (N,P,C): scale * x broadcast over P."""

    @staticmethod
    def forward(ctx, x, P, scale_):
        x = _c(x)
        N, C = x.shape
        ctx.scale = scale_
        y = torch.empty((N, P, C), dtype=x.dtype, device=x.device)
        lib().call("vca_spatial_bcast", _dt(x), x, y, N, P, C, scale_)
        return y

    @staticmethod
    def backward(ctx, g):
        return SpatialSumFn.apply(g, ctx.scale), None, None


def spatial_mean(x):
    """(N,H,W,C) -> (N,C) mean over H,W  (Avgpool of generator.py:137-140, AvgPool2d(4) of resnet.py:82)."""
    N, C = x.shape[0], x.shape[-1]
    P = x.numel() // (N * C)
    return SpatialSumFn.apply(x.reshape(N, P, C), 1.0 / P)


def spatial_tile(x, P):
    """(N,C) -> (N,P,C)."""
    return SpatialBcastFn.apply(x, P, 1.0)


# ------------------------------------------------------------------------------------------------------------
# batched GEMM, softmax, losses (fp32)
# ------------------------------------------------------------------------------------------------------------
def _gemm_raw(A, B, C, bias=None, alpha=1.0, beta=0.0):
    """C[z] = alpha*A[z]@B[z] + bias + beta*C[z]; A (Z,M,K), B (Z,K,N), C (Z,M,N) arbitrary-stride views
    (2-D inputs are treated as Z = 1; a Z-stride of 0 broadcasts)."""
    if A.dim() == 2:
        A, B, C = A.unsqueeze(0), B.unsqueeze(0), C.unsqueeze(0)
    Z, M, K = A.shape
    N = B.shape[2]
    assert B.shape[1] == K and C.shape[1] == M and C.shape[2] == N, (A.shape, B.shape, C.shape)
    sa, sb, sc = A.stride(), B.stride(), C.stride()
    lib().call("vca_gemm_simt", _dt(A), _dt(B), _dt(C), A, B, C, bias, Z, M, N, K, sa[0] if A.shape[0] > 1 else 0, sa[1], sa[2],
               sb[0] if B.shape[0] > 1 else 0, sb[1], sb[2], sc[0], sc[1], sc[2], float(alpha), float(beta))
    return C


def bmm_tc_raw(a, b, M, N, K, a_mn=False, b_mn=False, out_dtype=torch.bfloat16, alpha=1.0, out=None):
    """out[z] (M x N) = alpha * op(a[z]) op(b[z]) on the tcgen05 path (csrc/bmm_tc.cu).  a, b: bf16, 3-D, last dim
    contiguous; a is (Z, M, K) or -- a_mn -- (Z, K, M); b is (Z, N, K) or -- b_mn -- (Z, K, N).  Row pitches and batch
    strides are taken from the tensors (views with padded rows are fine), so ragged K / M / N just pass smaller extents."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 3 and b.dim() == 3
    assert a.stride(2) == 1 and b.stride(2) == 1
    Z = a.shape[0]
    if out is None:
        out = torch.empty((Z, M, N), dtype=out_dtype, device=a.device)
    assert out.stride(2) == 1
    lib().call("vca_bmm_tc", a, b, out, Z, M, N, K, 1 if a_mn else 0, 1 if b_mn else 0, a.stride(1), a.stride(0) if Z > 1 else 0,
               b.stride(1), b.stride(0) if Z > 1 else 0, out.stride(1), out.stride(0), 1 if out.dtype == torch.float32 else 0, float(alpha))
    return out


class BmmFn(Function):
    """out = alpha * a @ b for a (Z,M,K), b (Z,K,N) (views with any strides are fine)."""

    @staticmethod
    def forward(ctx, a, b, alpha):
        ctx.save_for_backward(a, b)
        ctx.alpha = alpha
        out = torch.empty((a.shape[0], a.shape[1], b.shape[2]), dtype=a.dtype, device=a.device)
        return _gemm_raw(a, b, out, None, alpha, 0.0)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = BmmFn.apply(g, b.transpose(1, 2), ctx.alpha) if ctx.needs_input_grad[0] else None
        gb = BmmFn.apply(a.transpose(1, 2), g, ctx.alpha) if ctx.needs_input_grad[1] else None
        return ga, gb, None


def bmm(a, b, alpha=1.0):
    return BmmFn.apply(a, b, float(alpha))


class AttentionFn(Function):
    """O = softmax_keys(Q K^T * scale, keys >= lens masked) V on the tcgen05 path (csrc/att_tc.cu): q (B,Tq,256),
    k, v (B,S,256) bf16, lens int32 (B,) -> (B,Tq,256) bf16.  Forward is ONE kernel (scores and O never leave TMEM / the
    probabilities go to shared memory as the second MMA's operand); the backward's four contractions run on the batched
    tcgen05 GEMM (csrc/bmm_tc.cu) around one row-wise softmax-gradient kernel."""

    @staticmethod
    def forward(ctx, q, k, v, lens, scale):
        q, k, v = _c(q), _c(k), _c(v)
        B, Tq, Dm = q.shape
        S = k.shape[1]
        SP = (S + 15) // 16 * 16
        o = torch.empty((B, Tq, Dm), dtype=torch.bfloat16, device=q.device)
        p = torch.empty((B, Tq, SP), dtype=torch.bfloat16, device=q.device)
        lib().call("vca_att_fwd_tc", q, k, v, lens, o, p, B, Tq, S, float(scale))
        ctx.save_for_backward(q, k, v, p)
        ctx.scale = float(scale)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, do):
        q, k, v, p = ctx.saved_tensors
        do = _c(do)
        B, Tq, Dm = q.shape
        S, SP = k.shape[1], p.shape[2]
        dp = torch.empty((B, Tq, SP), dtype=torch.float32, device=q.device)
        bmm_tc_raw(do, v, Tq, S, Dm, False, False, out=dp)                         # dP = dO V^T
        ds = torch.empty((B, Tq, SP), dtype=torch.bfloat16, device=q.device)
        lib().call("vca_att_softmax_bwd", p, dp, ds, B * Tq, S, SP, ctx.scale)      # dS = scale * P o (dP - <dP, P>)
        dq = bmm_tc_raw(ds, k, Tq, Dm, S, False, True) if ctx.needs_input_grad[0] else None      # dS K
        dk = bmm_tc_raw(ds, q, S, Dm, Tq, True, True) if ctx.needs_input_grad[1] else None       # dS^T Q
        dv = bmm_tc_raw(p, do, S, Dm, Tq, True, True) if ctx.needs_input_grad[2] else None       # P^T dO
        return dq, dk, dv, None, None


def attention_supported(q, k) -> bool:
    return (cfg.use_tc and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16
            and lib().query("vca_att_tc_supported", int(k.shape[1]), int(q.shape[2])) == 1)


def attention(q, k, v, lens, scale):
    return AttentionFn.apply(q, k, v, lens, float(scale))


class MaskedSoftmaxFn(Function):
    """softmax over the last dim of (Z,R,S) with keys >= lens[z] masked out (generator.py:161-164)."""

    @staticmethod
    def forward(ctx, x, lens):
        x = _c(x)
        Z, R, S = x.shape
        p = torch.empty_like(x)
        lib().call("vca_masked_softmax_fwd", x, p, lens, Z, R, S)
        ctx.save_for_backward(p)
        return p

    @staticmethod
    @once_differentiable
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        dp = _c(dp)
        dx = torch.empty_like(p)
        lib().call("vca_softmax_bwd", dp, p, dx, p.shape[0] * p.shape[1], p.shape[2])
        return dx, None


def masked_softmax(x, lens):
    return MaskedSoftmaxFn.apply(x, lens)


class L2NormFn(Function):
    """F.normalize(x, dim=-1) (eps 1e-12)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        D = x.shape[-1]
        rows = x.numel() // D
        y = torch.empty_like(x)
        norms = torch.empty(rows, dtype=torch.float32, device=x.device)
        lib().call("vca_l2norm_fwd", x, y, norms, rows, D, 1e-12)
        ctx.save_for_backward(y, norms)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, norms = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(y)
        lib().call("vca_l2norm_bwd", dy, y, norms, dx, norms.numel(), y.shape[-1], 1e-12)
        return dx


class NceDiagFn(Function):
    """(B,S,S) similarities -> (B,) symmetric InfoNCE of generator.py:354-359."""

    @staticmethod
    def forward(ctx, sim):
        sim = _c(sim)
        B, S, _ = sim.shape
        loss = torch.empty(B, dtype=torch.float32, device=sim.device)
        dsim = torch.empty_like(sim)
        lib().call("vca_nce_diag", sim, loss, dsim, B, S)
        ctx.save_for_backward(dsim)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (dsim,) = ctx.saved_tensors
        return dsim * g.view(-1, 1, 1)


class CosAbsMeanFn(Function):
    """(B,S,D) x (B,S,D) -> (B,): 5 - mean_t |cos(v_t, a_t)|  (generator.py:347-349)."""

    @staticmethod
    def forward(ctx, v, a):
        v, a = _c(v), _c(a)
        B, S, D = v.shape
        loss = torch.empty(B, dtype=torch.float32, device=v.device)
        saved = torch.empty((B, S, 3), dtype=torch.float32, device=v.device)
        lib().call("vca_cos_abs_mean_fwd", v, a, loss, saved, B, S, D)
        ctx.save_for_backward(v, a, saved)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        v, a, saved = ctx.saved_tensors
        B, S, D = v.shape
        g = _c(g)
        dv = torch.empty_like(v) if ctx.needs_input_grad[0] else None
        da = torch.empty_like(a) if ctx.needs_input_grad[1] else None
        if dv is None and da is None:
            return None, None
        lib().call("vca_cos_abs_mean_bwd", g, v, a, saved, da, dv, B, S, D)
        return dv, da


class SoftplusMeanFn(Function):
    """mean(softplus(sign * x)) -> scalar  (gan_loss, generator.py:363-366)."""

    @staticmethod
    def forward(ctx, x, sign):
        x = _c(x)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        lib().call("vca_softplus_mean", x, out, dx, x.numel(), sign)
        ctx.save_for_backward(dx)
        return out.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None


class L1MeanFn(Function):
    """scale * mean|a - b| (nn.L1Loss, train.py:150,226-229); gradient flows to a only."""

    @staticmethod
    def forward(ctx, a, b, scale_):
        a, b = _c(a), _c(b)
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        s = scale_ / a.numel()
        lib().call("vca_reduce_l1_sq", _dt(a), a, b, a.numel(), s, 0, out)
        ctx.save_for_backward(a, b)
        ctx.s = s
        return out.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a)
        lib().call("vca_l1_bwd", _dt(a), a, b, _c(g.float().view(1)), a.numel(), ctx.s, da)
        return da, None, None


def l1_mean(a, b, scale_=1.0):
    return L1MeanFn.apply(a, b.detach(), float(scale_))


class SumSqFn(Function):
    """scale * sum(x^2) -> scalar; backward 2*scale*x*g (differentiable: it is AxpbyFn)."""

    @staticmethod
    def forward(ctx, x, scale_):
        x = _c(x)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        lib().call("vca_reduce_l1_sq", _dt(x), x, None, x.numel(), scale_, 1, out)
        ctx.save_for_backward(x)
        ctx.scale = scale_
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return scale(x, 2.0 * ctx.scale) * g.to(x.dtype), None


def sum_sq(x, scale_=1.0):
    return SumSqFn.apply(x, float(scale_))


# ------------------------------------------------------------------------------------------------------------
# GRU layer (bidirectional), fp32
# ------------------------------------------------------------------------------------------------------------
class GRURecurrenceFn(Function):
    """Recurrent part of one bidirectional GRU layer (fp32, exact): gi (2,T,B,3H) = input projections incl. b_ih for
    the forward and reverse direction -> out (T,B,2H).  Gate order r,z,n, b_hn inside the r-product (torch.nn.GRU,
    visual_front.py:20).  Per time step: one skinny GEMM h @ W_hh^T (both directions in one launch) + one gate kernel."""

    @staticmethod
    def forward(ctx, gi, w_hh_f, b_hh_f, w_hh_r, b_hh_r):
        _require_cuda(gi)
        gi = _c(gi)
        _, T, B, H3 = gi.shape
        H = H3 // 3
        dev = gi.device
        whh = torch.stack([w_hh_f.detach(), w_hh_r.detach()], 0).contiguous()   # (2,3H,H)
        bhh = torch.stack([b_hh_f.detach(), b_hh_r.detach()], 0).contiguous()   # (2,3H)
        out = torch.empty((T, B, 2 * H), dtype=torch.float32, device=dev)
        gates = torch.empty((2, T, B, 4 * H), dtype=torch.float32, device=dev)
        h = torch.zeros((2, 2, B, H), dtype=torch.float32, device=dev)          # ping-pong
        gh = torch.empty((2, B, 3 * H), dtype=torch.float32, device=dev)
        L = lib()
        bar = torch.empty(1, dtype=torch.int32, device=dev)
        if cfg.gru_persistent and L.try_call("vca_gru_seq_fwd", gi, whh, bhh, h, out, gates, bar, 2, T, B, H):
            pass   # whole sequence in one cooperative launch
        else:
            for s in range(T):
                hp, hn = h[s & 1], h[(s + 1) & 1]
                L.call("vca_skinny_gemm", hp, whh, gh, 2, B, 3 * H, H, 0.0)
                L.call("vca_gru_gate_fwd", gi, gh, bhh, hp, hn, out, gates, 2, T, B, H, s)
        ctx.save_for_backward(whh, out, gates)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        whh, out, gates = ctx.saved_tensors
        dout = _c(dout.float())
        T, B, H2 = out.shape
        H = H2 // 2
        dev = out.device
        dgi = torch.empty((2, T, B, 3 * H), dtype=torch.float32, device=dev)
        dgh = torch.empty((2, T, B, 3 * H), dtype=torch.float32, device=dev)
        dh = torch.zeros((2, B, H), dtype=torch.float32, device=dev)
        whh_t = whh.transpose(1, 2).contiguous()                                 # (2,H,3H): dh += dgh @ W_hh
        cur = torch.empty((2, B, 3 * H), dtype=torch.float32, device=dev)
        L = lib()
        bar = torch.empty(1, dtype=torch.int32, device=dev)
        dhc = torch.empty((2, 2, B, H), dtype=torch.float32, device=dev)
        dhz = torch.empty((2, B, H), dtype=torch.float32, device=dev)
        if cfg.gru_persistent and L.try_call("vca_gru_seq_bwd", dout, whh, gates, out, dgi, dgh, dhc, cur, dhz, bar, 2, T, B, H):
            pass
        else:
            for s in range(T):
                L.call("vca_gru_gate_bwd", dout, dh, gates, out, dgi, dgh, cur, 2, T, B, H, s)
                L.call("vca_skinny_gemm", cur, whh_t, dh, 2, B, H, 3 * H, 1.0)
        grads = []
        for d in range(2):
            dw_hh = torch.zeros((3 * H, H), dtype=torch.float32, device=dev)
            if T > 1:  # h_{t-1} of direction 0 is out[t-1,:, :H]; of direction 1 it is out[t+1,:, H:]
                if d == 0:
                    a, hprev = dgh[0, 1:].reshape((T - 1) * B, 3 * H), out[:T - 1, :, :H].reshape((T - 1) * B, H)
                else:
                    a, hprev = dgh[1, :T - 1].reshape((T - 1) * B, 3 * H), out[1:, :, H:].reshape((T - 1) * B, H)
                rows = (T - 1) * B
                g1, _ = _geom((rows, 1, 1, H), (3 * H, H, 1, 1), (1, 1), (0, 0))
                if cfg.dtype == torch.bfloat16 and _tc_ok(g1, 2, torch.bfloat16):
                    # bf16 mode: dW_hh = dgh^T @ h_prev as a 1x1 wgrad on the tcgen05 path (K = (T-1)*B "pixels")
                    lib().call("vca_conv_wgrad_tc", g1, a.to(torch.bfloat16).contiguous(), hprev.to(torch.bfloat16).contiguous(), dw_hh)
                else:
                    _gemm_raw(a.t(), hprev, dw_hh)
            grads += [dw_hh, ColSumFn.apply(dgh[d].view(T * B, 3 * H))]
        return (dgi, *grads)


def gru_layer(x, params):
    """One bidirectional GRU layer: x (T,B,I) fp32 -> (T,B,2H) fp32.  params = (w_ih, w_hh, b_ih, b_hh) forward then
    reverse.  The input projections are ordinary linears (tcgen05 path in bf16 mode); the recurrence stays fp32."""
    w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r = params
    xc = cast(x, cfg.dtype)
    gi = torch.stack([cast(linear(xc, w_ih_f, b_ih_f), torch.float32), cast(linear(xc, w_ih_r, b_ih_r), torch.float32)], 0)
    return GRURecurrenceFn.apply(gi, w_hh_f, b_hh_f, w_hh_r, b_hh_r)
