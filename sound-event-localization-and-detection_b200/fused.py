"""Fused training-mode paths through the glue between the convolutions.

`cnn_stack` runs the whole CNN front of ConvTC_Block (model.py:261-285: per block Q/DQ conv2d ->
BatchNorm2d -> ReLU -> MaxPool2d([p, 1]) -> Dropout) as one autograd Function over the C ABI:

    forward, per block    conv (tcgen05, fp16 NCHW output)  ->  bn_stats  ->  bn_finalize  ->  cnn_tail_fwd
                          (the pooled activation leaves cnn_tail_fwd directly as the next convolution's
                          channels-last bf16 operand; only the last block also emits fp32 NCHW)
    backward, per block   cnn_tail_bwd (both BatchNorm reductions + d(conv out) in the two bf16 operand
                          layouts)  ->  wgrad  ->  dgrad

The arithmetic is the reference's (batch statistics, biased variance for normalisation, unbiased for the
running estimate, first-maximum pooling, inverted dropout); the full-resolution fp32 intermediates are
what is gone.  The unfused layer-by-layer modules remain the definition of the semantics and are what
eval mode, fp32 mode and every configuration outside `cnn_stack_supported` run.
"""
import ctypes
import os

import torch

from . import _lib
from . import functional as F
from ._lib import PASS_DGRAD, PASS_FWD, PASS_WGRAD, PREC_BF16

_P = ctypes.c_void_p
ENABLED = os.environ.get("SELDQ_FUSED", "1") != "0"      # SELDQ_FUSED=0: always the layer-by-layer modules
# SELDQ_FIRST_FUSED=0: the first CNN block's backward writes d(conv out) and runs the stand-alone wgrad kernel
FIRST_FUSED = os.environ.get("SELDQ_FIRST_FUSED", "1") != "0"
# SELDQ_SIDE_WGRAD=0: the TCN weight gradients stay on the main stream
SIDE_WGRAD = os.environ.get("SELDQ_SIDE_WGRAD", "1") != "0"
# SELDQ_TCN_PAIR=0: one launch per convolution of a residual block and stand-alone statistics / residual kernels
# (the round-1 schedule); default: sibling convolutions share a launch and the glue rides in their epilogues
TCN_PAIR = os.environ.get("SELDQ_TCN_PAIR", "1") != "0"
# SELDQ_TCN_EPI=1: the statistics / residual / skip-sum glue rides in the sibling launches' epilogues instead of
# stand-alone kernels (seldq_conv_epilogue_t)
TCN_EPI = os.environ.get("SELDQ_TCN_EPI", "0") != "0"
# SELDQ_TCN_FUSED_GLUE=0: every reduce / apply pair of the residual-block glue stays two launches (default: one launch
# with a grid barrier where all tiles of the tensor can be resident at once -- seldq_tcn_glue_fused_supported)
TCN_FUSED_GLUE = os.environ.get("SELDQ_TCN_FUSED_GLUE", "0") != "0"
# SELDQ_TCN_SKIP_FORK=0: the running sum of the skip outputs stays in the residual kernel of the main stream
TCN_SKIP_FORK = os.environ.get("SELDQ_TCN_SKIP_FORK", "1") != "0"
_SIDE = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def next_drop_seed(module, device):
    """Seed of this step's dropout masks on the fused paths of `module` (a ConvTC_Block or TC_Block, the mirror's or
    the reference's own): an int64 device scalar kept as the non-persistent buffer `_drop_seed`.  Its base value is
    drawn from the torch RNG the first time the module takes a fused path -- so torch.manual_seed governs it, every
    replica (per-rank seed) and every module (branch_A / branch_B, CNN / TCN) gets its own, and a resumed run that
    restores the RNG state continues with fresh masks -- and it advances by one per step (a device-side add, so a
    captured CUDA graph keeps stepping it).  The kernels hash (seed, per-layer salt, element index)."""
    seed = module._buffers.get("_drop_seed")
    if seed is None or seed.device != device or not getattr(module, "_drop_seed_ready", False):
        base = torch.randint(1, 2 ** 62, (1,), dtype=torch.int64)          # default CPU generator
        if seed is None or seed.device != device:
            module.register_buffer("_drop_seed", base.to(device), persistent=False)
        else:
            seed.copy_(base)
        module._drop_seed_ready = True
    module._drop_seed.add_(1)
    return module._drop_seed


_COUNTERS = {}


def _bump_counters(bns):
    """nn.BatchNorm's step counters (num_batches_tracked += 1) of a fused stack in ONE element-wise launch: the
    counters become 0-dim views of one flat int64 tensor (as the trainer does with parameters and gradients); a
    multi-tensor launch over 30 one-element tensors cost 13 us.  Rebuilt whenever a counter no longer aliases the flat
    tensor (model.to(...), load of a different module, ...)."""
    key = tuple(id(b) for b in bns)
    flat = _COUNTERS.get(key)
    ok = flat is not None and flat.numel() == len(bns) and all(
        b.num_batches_tracked is not None and b.num_batches_tracked.device == flat.device
        and b.num_batches_tracked.data_ptr() == flat[i].data_ptr() for i, b in enumerate(bns))
    if not ok:
        flat = torch.stack([b.num_batches_tracked.detach().reshape(()).to(torch.int64) for b in bns]).contiguous()
        for i, b in enumerate(bns):
            b._buffers["num_batches_tracked"] = flat[i]
        if len(_COUNTERS) > 64:
            _COUNTERS.clear()
        _COUNTERS[key] = flat
    flat.add_(1)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _CnnStack(torch.autograd.Function):
    """spec: list (one entry per block) of dicts with keys
         algebra, ksize, padding, pool, drop_p, eps, momentum, nw, running_mean, running_var, salt
       tensors: per block nw compact weights, then gamma, beta."""

    @staticmethod
    def forward(ctx, x, spec, seed, *tensors):
        L = _lib.lib()
        dev = x.device
        x = x.contiguous()
        N = x.shape[0]
        saved, metas = [], []
        pos = 0
        blocks = []
        for s in spec:
            ws = tuple(t.contiguous() for t in tensors[pos:pos + s["nw"]])
            gamma, beta = tensors[pos + s["nw"]], tensors[pos + s["nw"] + 1]
            pos += s["nw"] + 2
            blocks.append((s, ws, gamma, beta))
        # descriptors first: every block's tail needs its consumer's descriptor
        descs = []
        cin, H, W = x.shape[1], x.shape[2], x.shape[3]
        for s, ws, _, _ in blocks:
            nc = F._NCOMP[s["algebra"]]
            cout = ws[0].shape[0] * nc
            kh, kw = s["ksize"]
            d = _lib.ConvDesc(s["algebra"], PREC_BF16, 2, N, cin, cout, H, W, kh, kw, 1, 1, s["padding"], s["padding"],
                              1, 1)
            oh, ow = ctypes.c_int32(), ctypes.c_int32()
            _lib.check(L.seldq_conv_out_shape(ctypes.byref(d), ctypes.byref(oh), ctypes.byref(ow)))
            descs.append((d, oh.value, ow.value))
            cin, H, W = cout, oh.value // s["pool"], ow.value
        with torch.cuda.device(dev):
            x_cl, _ = F.stage_operand(x, descs[0][0], 0)
            x_dense = F._operand_info(descs[0][0], 0)[1]
            cur_cl = x_cl
            z32 = None
            for k, (s, ws, gamma, beta) in enumerate(blocks):
                d, oh, ow = descs[k]
                last = k == len(blocks) - 1
                C = d.cout
                wp = _lib.ptr_array([w.data_ptr() for w in ws])
                pk = F.packed_weights(ws, d, PASS_FWD)
                y16 = torch.empty((N, C, oh, ow), dtype=torch.float16, device=dev)
                F._timed("qconv_cl_fprop_kernel", F._conv_flop(d, oh, ow), 1, lambda: _lib.check(
                    L.seldq_conv_fwd_bf16(ctypes.byref(d), None, cur_cl.data_ptr(), wp, _ptr(pk), None, y16.data_ptr(),
                                          None, 0, _stream())))
                sums = torch.zeros((C, 2), dtype=torch.float64, device=dev)
                coef = torch.empty((C, 4), dtype=torch.float32, device=dev)
                F._timed("bn_stats_kernel", 0.0, 1, lambda: _lib.check(
                    L.seldq_bn_stats(y16.data_ptr(), 1, N, C, oh * ow, sums.data_ptr(), _stream())))
                _lib.check(L.seldq_bn_finalize(sums.data_ptr(), _ptr(gamma), _ptr(beta), C, float(N * oh * ow), s["eps"],
                                               s["momentum"], _ptr(s["running_mean"]), _ptr(s["running_var"]),
                                               coef.data_ptr(), _stream()))
                hp = oh // s["pool"]
                td = _lib.CnnTailDesc(N, C, oh, ow, s["pool"], s["drop_p"], s["salt"])
                idx = torch.empty((N, C, hp, ow), dtype=torch.uint8, device=dev)
                ymax = torch.empty((N, C, hp, ow), dtype=torch.float16, device=dev)
                if last:
                    z32 = torch.empty((N, C, hp, ow), dtype=torch.float32, device=dev)
                    nxt_cl, consumer = None, None
                else:
                    consumer = descs[k + 1][0]
                    nxt_cl = torch.empty(F._operand_info(consumer, 0)[2], dtype=torch.uint8, device=dev)
                F._timed("cnn_tail_fwd_kernel", 0.0, 1, lambda: _lib.check(
                    L.seldq_cnn_tail_fwd(ctypes.byref(td), None if consumer is None else ctypes.byref(consumer),
                                         y16.data_ptr(), coef.data_ptr(), _ptr(seed) if s["drop_p"] > 0 else None,
                                         _ptr(nxt_cl), _ptr(z32), idx.data_ptr(), ymax.data_ptr(), _stream())))
                # block 0 of a narrow first layer keeps the fp32 input for its weight gradient (seldq.h)
                keep_in = x if (k == 0 and x_dense) else cur_cl
                saved += [keep_in, y16, coef, idx, ymax]
                metas.append((s, d, oh, ow, td))
                cur_cl = nxt_cl
        ctx.metas = metas
        ctx.x_dense = x_dense
        ctx.nblocks = len(blocks)
        ctx.block_params = [(ws, gamma, beta) for _, ws, gamma, beta in blocks]
        ctx.save_for_backward(*saved)
        return z32

    @staticmethod
    def backward(ctx, gz):
        L = _lib.lib()
        saved = ctx.saved_tensors
        dev = gz.device
        gz = gz.contiguous()
        grads = []
        gx = None
        with torch.cuda.device(dev):
            for k in reversed(range(ctx.nblocks)):
                s, d, oh, ow, td = ctx.metas[k]
                xin, y16, coef, idx, ymax = saved[5 * k:5 * k + 5]
                ws, gamma, beta = ctx.block_params[k]
                N, C = d.batch, d.cout
                need_gx = k > 0 or ctx.needs_input_grad[0]
                if (FIRST_FUSED and k == 0 and ctx.x_dense and not need_gx
                        and L.seldq_cnn_first_bwd_supported(ctypes.byref(td), ctypes.byref(d))):
                    # no input gradient: d(conv out) is formed inside the weight-gradient kernel (wgrad_first.cu)
                    dsums = torch.zeros((C * 3,), dtype=torch.float64, device=dev)
                    gws, direct = F._grad_targets(ws, [True] * len(ws))
                    gp = _lib.ptr_array([g.data_ptr() for g in gws])
                    work = torch.empty(L.seldq_cnn_first_bwd_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8,
                                       device=dev)
                    F._timed("qconv_cl_wgrad_kernel", F._conv_flop(d, oh, ow), 4, lambda: _lib.check(
                        L.seldq_cnn_first_bwd(ctypes.byref(td), ctypes.byref(d), xin.data_ptr(), y16.data_ptr(),
                                              coef.data_ptr(), idx.data_ptr(), ymax.data_ptr(), gz.data_ptr(),
                                              dsums.data_ptr(), gp, 1 if direct else 0, work.data_ptr(),
                                              work.numel(), _stream())))
                    dsf = dsums[:2 * C].view(C, 2).float()
                    if direct:
                        gws = [None] * len(ws)
                    grads = list(gws) + [dsf[:, 1].contiguous() if gamma is not None else None,
                                         dsf[:, 0].contiguous() if beta is not None else None] + grads
                    continue
                _, _, clb, t16b = F._operand_info(d, 1)
                d_t16 = torch.empty(t16b, dtype=torch.uint8, device=dev)
                d_cl = torch.empty(clb, dtype=torch.uint8, device=dev) if need_gx else None
                dsums = torch.zeros((C * 3,), dtype=torch.float64, device=dev)
                F._timed("cnn_tail_bwd_kernels", 0.0, 3, lambda: _lib.check(
                    L.seldq_cnn_tail_bwd(ctypes.byref(td), ctypes.byref(d), y16.data_ptr(), coef.data_ptr(),
                                         idx.data_ptr(), ymax.data_ptr(), gz.data_ptr(), dsums.data_ptr(),
                                         d_t16.data_ptr(), _ptr(d_cl),
                                         _stream())))
                dsf = dsums[:2 * C].view(C, 2).float()
                g_gamma, g_beta = dsf[:, 1].contiguous(), dsf[:, 0].contiguous()
                # weight gradient
                gws, direct = F._grad_targets(ws, [True] * len(ws))
                gp = _lib.ptr_array([g.data_ptr() for g in gws])
                dense0 = k == 0 and ctx.x_dense
                work = None
                if dense0:
                    work = torch.empty(L.seldq_conv_workspace_bytes(ctypes.byref(d), PASS_WGRAD), dtype=torch.uint8,
                                       device=dev)
                F._timed("qconv_cl_wgrad_kernel", F._conv_flop(d, oh, ow), 1, lambda: _lib.check(
                    L.seldq_conv_wgrad(ctypes.byref(d), xin.data_ptr() if dense0 else None,
                                       None if dense0 else xin.data_ptr(), None, d_t16.data_ptr(), gp, None,
                                       1 if direct else 0, _ptr(work), 0 if work is None else work.numel(), _stream())))
                if direct:
                    gws = [None] * len(ws)
                # input gradient = pooled-output gradient of the previous block
                if need_gx:
                    wp = _lib.ptr_array([w.data_ptr() for w in ws])
                    pk = F.packed_weights(ws, d, PASS_DGRAD)
                    gprev = torch.empty((N, d.cin, d.in_h, d.in_w), dtype=torch.float32, device=dev)
                    F._timed("qconv_cl_fprop_kernel", F._conv_flop(d, oh, ow), 1, lambda: _lib.check(
                        L.seldq_conv_dgrad(ctypes.byref(d), None, d_cl.data_ptr(), wp, _ptr(pk), gprev.data_ptr(), None,
                                           0, _stream())))
                    if k > 0:
                        gz = gprev
                    else:
                        gx = gprev
                grads = list(gws) + [g_gamma if gamma is not None else None,
                                     g_beta if beta is not None else None] + grads
        return (gx, None, None) + tuple(grads)


def cnn_stack_supported(convs, bns, pools, drops, x, training):
    """The fused path serves the shipped training configuration: CUDA, bf16 precision, training mode, Q / DQ
    conv2d without bias, stride 1, BatchNorm2d with a float momentum, pooling over frequency only."""
    if not ENABLED:
        return False
    if not (training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and F.get_precision() == "bf16"):
        return False
    if torch.is_grad_enabled() is False:
        return False
    for k, (conv, bn, pool) in enumerate(zip(convs, bns, pools)):
        if not hasattr(conv, "_algebra") or conv.bias is not None or conv.rotation or conv.groups != 1:
            return False
        if k > 0 and conv.in_channels < 8:       # narrow inner layer: its wgrad would need the fp32 activation
            return False
        if F._pair(conv.stride) != (1, 1) or F._pair(conv.dilatation) != (1, 1) or not isinstance(conv.padding, int):
            return False
        if bn is None or not isinstance(bn, torch.nn.BatchNorm2d) or bn.momentum is None or not bn.track_running_stats:
            return False
        ks = pool.kernel_size if isinstance(pool.kernel_size, (tuple, list)) else (pool.kernel_size, pool.kernel_size)
        st = pool.stride if isinstance(pool.stride, (tuple, list)) else (pool.stride, pool.stride)
        if tuple(ks)[1] != 1 or tuple(st) != tuple(ks) or not 1 <= tuple(ks)[0] <= 8 or pool.padding not in (0, (0, 0)):
            return False
    return True


def cnn_stack(x, convs, bns, pools, drops, seed):
    """x (N, Cin, F, T) fp32 -> pooled activation of the last CNN block (N, C, F', T) fp32."""
    spec, tensors = [], []
    for k, (conv, bn, pool, drop) in enumerate(zip(convs, bns, pools, drops)):
        ws = conv._weights()
        ks = pool.kernel_size if isinstance(pool.kernel_size, (tuple, list)) else (pool.kernel_size, pool.kernel_size)
        spec.append(dict(algebra=conv._algebra, ksize=tuple(ws[0].shape[2:]), padding=int(conv.padding),
                         pool=int(tuple(ks)[0]), drop_p=float(drop.p) if drop is not None else 0.0, eps=float(bn.eps),
                         momentum=float(bn.momentum), nw=len(ws), running_mean=bn.running_mean,
                         running_var=bn.running_var, salt=k + 1))
        tensors += list(ws) + [bn.weight, bn.bias]
    _bump_counters(bns)
    return _CnnStack.apply(x, spec, seed, *tensors)


# ---- TCN residual blocks (model.py:109-132) -------------------------------------------------------------------
def _glue(op, layout_of, which, N, C, T, c2=0, eps=1e-5, momentum=0.1, drop_p=0.0, salt=0, bn=(), inp=(), out32=None,
          out_cl=(), out_t16=(), dsums=None, stats_out=(), accum=None, seed=None, flag=0, sync=None, out32b=None):
    """One seldq_tcn_glue call (include/seldq.h).  bn: up to two (sums, gamma, beta, running_mean, running_var)."""
    a = _lib.TcnGlue()
    a.n, a.c, a.t, a.c2 = N, C, T, c2
    a.eps, a.momentum, a.drop_p, a.salt = eps, momentum, drop_p, salt
    a.count = float(N * T)
    for i, b in enumerate(bn):
        a.bn[i].sums, a.bn[i].gamma, a.bn[i].beta, a.bn[i].running_mean, a.bn[i].running_var = [_ptr(t) for t in b]
    for i, t in enumerate(inp):
        a.inp[i] = _ptr(t)
    a.out32 = _ptr(out32)
    for i, t in enumerate(out_cl):
        a.out_cl[i] = _ptr(t)
    for i, t in enumerate(out_t16):
        a.out_t16[i] = _ptr(t)
    a.dsums = _ptr(dsums)
    for i, t in enumerate(stats_out):
        a.stats_out[i] = _ptr(t)
    a.accum = _ptr(accum)
    a.seed = _ptr(seed) if drop_p > 0 else None
    a.flag = flag
    a.sync, a.out32b = _ptr(sync), _ptr(out32b)
    F._timed("tcn_glue_kernels", 0.0, 1, lambda: _lib.check(_lib.lib().seldq_tcn_glue(
        op, ctypes.byref(a), None if layout_of is None else ctypes.byref(layout_of), which, _stream())))


def _epi(mode=0, addend=None, stats=None):
    e = _lib.ConvEpilogue()
    e.mode, e.addend, e.stats = mode, _ptr(addend), _ptr(stats)
    return e


def _conv_epi(d, pass_, in_cl, pk, out, epi, T):
    """One convolution pass from pre-staged operands with a fused epilogue (seldq_conv_epi)."""
    F._timed("qconv_cl_fprop_kernel", F._conv_flop(d, 1, T), 1, lambda: _lib.check(_lib.lib().seldq_conv_epi(
        ctypes.byref(d), pass_, in_cl.data_ptr(), pk.data_ptr(), out.data_ptr(), ctypes.byref(epi), _stream())))


def _conv_pair(d, pass_, in_a, in_b, pk_a, pk_b, out_a, out_b, epi_a, epi_b, T):
    """Two sibling convolutions of a residual block in one launch (seldq_conv_pair)."""
    F._timed("qconv_cl_fprop_kernel", 2 * F._conv_flop(d, 1, T), 1, lambda: _lib.check(_lib.lib().seldq_conv_pair(
        ctypes.byref(d), pass_, in_a.data_ptr(), in_b.data_ptr(), pk_a.data_ptr(), pk_b.data_ptr(), out_a.data_ptr(),
        out_b.data_ptr(), ctypes.byref(epi_a), ctypes.byref(epi_b), _stream())))


def _same_geometry(a, b):
    return bytes(a) == bytes(b)


class _TcnStack(torch.autograd.Function):
    """All residual blocks of TC_Block in one autograd node.
       spec: per block dict(algebra, nw, k, dil, pad, drop_p, salt, has_res, bn1, bnf, bng) with bn* = (eps, momentum,
             running_mean, running_var)
       tensors: per block  conv1_filter (nw), conv1_gate (nw), conv2_skip (nw), conv2_residual (nw) compact
             weights, then gamma / beta of batch_filter1, batch_filter2, batch_gate2."""

    @staticmethod
    def forward(ctx, r, spec, seed, *tensors):
        L_ = _lib.lib()
        dev = r.device
        r = r.contiguous()
        N, Lc, T = r.shape
        f32 = dict(dtype=torch.float32, device=dev)
        blocks, pos = [], 0
        for s in spec:
            nw = s["nw"]
            wf, wg, wsk, wr = (tuple(t.contiguous() for t in tensors[pos + i * nw:pos + (i + 1) * nw]) for i in range(4))
            bnp = tensors[pos + 4 * nw:pos + 4 * nw + 6]
            pos += 4 * nw + 6
            blocks.append((s, wf, wg, wsk, wr, bnp))
        saved, metas = [], []
        skip_sum = None
        with torch.cuda.device(dev):
            # batch statistics of every BatchNorm of the stack: one zeroed buffer, [block](sums1 (L,2), sums2 (2,G,2))
            gs_ = [w[1][0].shape[0] * F._NCOMP[w[0]["algebra"]] for w in blocks]
            stats = torch.zeros(sum(2 * Lc + 4 * g for g in gs_), dtype=torch.float64, device=dev)
            offs, o = [], 0
            for g in gs_:
                offs.append(o)
                o += 2 * Lc + 4 * g
            sums1 = stats[:2 * Lc].view(Lc, 2)
            _glue(_lib.TCN_ROW_STATS, None, 0, N, Lc, T, inp=(r,), stats_out=(sums1,), flag=1)
            # single-launch glue steps (grid barrier): one zeroed counter per step
            fuse_l = TCN_FUSED_GLUE and not TCN_EPI and bool(L_.seldq_tcn_glue_fused_supported(N, Lc, T))
            syncs = torch.zeros(2 * len(blocks), dtype=torch.int32, device=dev)
            xa = xa_cl = None
            skip_stream = None
            for k, (s, wf, wg, wsk, wr, bnp) in enumerate(blocks):
                nc = F._NCOMP[s["algebra"]]
                G, U = wf[0].shape[0] * nc, wsk[0].shape[0] * nc
                d1 = _lib.ConvDesc(s["algebra"], PREC_BF16, 1, N, Lc, G, 1, T, 1, s["k"], 1, 1, 0, s["pad"], 1, s["dil"])
                dsk = _lib.ConvDesc(s["algebra"], PREC_BF16, 1, N, G, U, 1, T, 1, 1, 1, 1, 0, 0, 1, 1)
                dre = _lib.ConvDesc(s["algebra"], PREC_BF16, 1, N, G, Lc, 1, T, 1, 1, 1, 1, 0, 0, 1, 1)
                g1, b1, gf, bf, gg, bg = bnp
                e1, m1, rm1, rv1 = s["bn1"]
                ef, mf, rmf, rvf = s["bnf"]
                eg, mg, rmg, rvg = s["bng"]
                # x = tanh(BN1(r)) (already there when the block before produced it together with r)
                if xa is None:
                    xa = torch.empty((N, Lc, T), **f32)
                    xa_cl = torch.empty(F._operand_info(d1, 0)[2], dtype=torch.uint8, device=dev)
                    _glue(_lib.TCN_PREACT_FWD, d1, 0, N, Lc, T, eps=e1, momentum=m1, bn=((sums1, g1, b1, rm1, rv1),),
                          inp=(r,), out32=xa, out_cl=(xa_cl,))
                # y_f, y_g: one sibling launch; the batch statistics of batch_filter2 / batch_gate2 come from the
                # convolutions' epilogues (TCN_EPI) or from one row-statistics kernel
                sums2 = stats[offs[k] + 2 * Lc:offs[k] + 2 * Lc + 4 * G].view(2, G, 2)
                pair1 = TCN_PAIR and bool(L_.seldq_conv_pair_supported(ctypes.byref(d1), PASS_FWD))
                yf, yg = torch.empty((N, G, T), **f32), torch.empty((N, G, T), **f32)
                if pair1:
                    _conv_pair(d1, PASS_FWD, xa_cl, xa_cl, F.packed_weights(wf, d1, PASS_FWD),
                               F.packed_weights(wg, d1, PASS_FWD), yf, yg, _epi(stats=sums2[0] if TCN_EPI else None),
                               _epi(stats=sums2[1] if TCN_EPI else None), T)
                else:
                    for w, y in ((wf, yf), (wg, yg)):
                        wp = _lib.ptr_array([t.data_ptr() for t in w])
                        pk = F.packed_weights(w, d1, PASS_FWD)
                        F._timed("qconv_cl_fprop_kernel", F._conv_flop(d1, 1, T), 1, lambda: _lib.check(
                            L_.seldq_conv_fwd(ctypes.byref(d1), None, xa_cl.data_ptr(), wp, _ptr(pk), None, y.data_ptr(),
                                              None, 0, _stream())))
                fuse_g = fuse_l and bool(L_.seldq_tcn_glue_fused_supported(N, G, T))
                # y = dropout1d(tanh(BN_f y_f) * sigmoid(BN_g y_g)), only as conv2's operand
                y_cl = torch.empty(F._operand_info(dsk, 0)[2], dtype=torch.uint8, device=dev)
                if fuse_g:
                    _glue(_lib.TCN_GATE_FWD_STATS, dsk, 0, N, G, T, eps=ef, momentum=mf, drop_p=s["drop_p"], salt=s["salt"],
                          bn=((sums2[0], gf, bf, rmf, rvf), (sums2[1], gg, bg, rmg, rvg)), inp=(yf, yg),
                          stats_out=(sums2[0], sums2[1]), out_cl=(y_cl,), seed=seed, sync=syncs[2 * k:])
                else:
                    if not (pair1 and TCN_EPI):
                        _glue(_lib.TCN_ROW_STATS, None, 0, N, G, T, inp=(yf, yg), stats_out=(sums2[0], sums2[1]), flag=2)
                    _glue(_lib.TCN_GATE_FWD, dsk, 0, N, G, T, eps=ef, momentum=mf, drop_p=s["drop_p"], salt=s["salt"],
                          bn=((sums2[0], gf, bf, rmf, rvf), (sums2[1], gg, bg, rmg, rvg)), inp=(yf, yg), out_cl=(y_cl,),
                          seed=seed)
                if skip_sum is None:
                    skip_sum = torch.empty((N, U, T), **f32)
                r_next = sums1_next = None
                if s["has_res"]:
                    r_next = torch.empty((N, Lc, T), **f32)
                    sums1_next = stats[offs[k + 1]:offs[k + 1] + 2 * Lc].view(Lc, 2)
                pair2 = TCN_PAIR and bool(L_.seldq_conv_pair_supported(ctypes.byref(dsk), PASS_FWD))
                if pair2 and TCN_EPI:
                    # skips (+)= conv2_skip(y) and r' = x + conv2_residual(y) with the statistics of r' for the next
                    # block's batch_filter1, all in the convolutions' epilogues: no stand-alone residual kernel
                    e_skip = _epi(_lib.EPI_STORE if k == 0 else _lib.EPI_ACCUMULATE)
                    if s["has_res"]:
                        e_res = _epi(_lib.EPI_ADD, addend=xa, stats=sums1_next)
                        if _same_geometry(dsk, dre):
                            _conv_pair(dsk, PASS_FWD, y_cl, y_cl, F.packed_weights(wsk, dsk, PASS_FWD),
                                       F.packed_weights(wr, dre, PASS_FWD), skip_sum, r_next, e_skip, e_res, T)
                        else:
                            _conv_epi(dsk, PASS_FWD, y_cl, F.packed_weights(wsk, dsk, PASS_FWD), skip_sum, e_skip, T)
                            _conv_epi(dre, PASS_FWD, y_cl, F.packed_weights(wr, dre, PASS_FWD), r_next, e_res, T)
                    else:
                        _conv_epi(dsk, PASS_FWD, y_cl, F.packed_weights(wsk, dsk, PASS_FWD), skip_sum, e_skip, T)
                else:
                    skip = torch.empty((N, U, T), **f32)
                    res = torch.empty((N, Lc, T), **f32) if s["has_res"] else None
                    if pair2 and s["has_res"] and _same_geometry(dsk, dre):
                        _conv_pair(dsk, PASS_FWD, y_cl, y_cl, F.packed_weights(wsk, dsk, PASS_FWD),
                                   F.packed_weights(wr, dre, PASS_FWD), skip, res, _epi(), _epi(), T)
                    else:
                        for w, d, o in ((wsk, dsk, skip), (wr, dre, res)):
                            if o is None:
                                continue
                            wp = _lib.ptr_array([t.data_ptr() for t in w])
                            pk = F.packed_weights(w, d, PASS_FWD)
                            F._timed("qconv_cl_fprop_kernel", F._conv_flop(d, 1, T), 1, lambda: _lib.check(
                                L_.seldq_conv_fwd(ctypes.byref(d), None, y_cl.data_ptr(), wp, _ptr(pk), None, o.data_ptr(),
                                                  None, 0, _stream())))
                    xa_next = xa_cl_next = None
                    if fuse_l and s["has_res"] and U == Lc:
                        # r' and the skip sum, then -- behind the grid barrier -- the next block's x = tanh(BN1(r'))
                        sn = blocks[k + 1][0]
                        d1n = _lib.ConvDesc(sn["algebra"], PREC_BF16, 1, N, Lc, blocks[k + 1][2][0].shape[0] *
                                            F._NCOMP[sn["algebra"]], 1, T, 1, sn["k"], 1, 1, 0, sn["pad"], 1, sn["dil"])
                        g1n, b1n = blocks[k + 1][5][0], blocks[k + 1][5][1]
                        e1n, m1n, rm1n, rv1n = sn["bn1"]
                        xa_next = torch.empty((N, Lc, T), **f32)
                        xa_cl_next = torch.empty(F._operand_info(d1n, 0)[2], dtype=torch.uint8, device=dev)
                        _glue(_lib.TCN_RESIDUAL_PREACT_FWD, d1n, 0, N, Lc, T, c2=U, eps=e1n, momentum=m1n,
                              bn=((sums1_next, g1n, b1n, rm1n, rv1n),), inp=(xa, res, skip), out32=r_next,
                              out32b=xa_next, out_cl=(xa_cl_next,), dsums=sums1_next, accum=skip_sum,
                              flag=1 if k == 0 else 0, sync=syncs[2 * k + 1:])
                    elif TCN_SKIP_FORK:
                        # the running sum of the skip outputs is a chain of its own (nothing in the residual stream
                        # reads it): it accumulates on a forked stream, joined behind the last block, and the residual
                        # kernel of the main stream moves three tensors instead of six
                        if skip_stream is None:
                            skip_stream = _side_stream(dev)
                        skip_stream.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(skip_stream):
                            _glue(_lib.TCN_RESIDUAL_FWD, None, 0, N, Lc, T, c2=U, inp=(None, None, skip), accum=skip_sum,
                                  flag=1 if k == 0 else 0)
                        skip.record_stream(skip_stream)
                        if s["has_res"]:
                            _glue(_lib.TCN_RESIDUAL_FWD, None, 0, N, Lc, T, c2=U, inp=(xa, res, None), out32=r_next,
                                  dsums=sums1_next)
                    else:
                        _glue(_lib.TCN_RESIDUAL_FWD, None, 0, N, Lc, T, c2=U, inp=(xa, res, skip), out32=r_next,
                              dsums=sums1_next, accum=skip_sum, flag=1 if k == 0 else 0)
                saved += [r, xa, xa_cl, yf, yg, y_cl, sums1, sums2]
                metas.append((s, d1, dsk, dre, G, U))
                r, sums1 = r_next, sums1_next
                if TCN_EPI and pair2:
                    xa = xa_cl = None
                else:
                    xa, xa_cl = xa_next, xa_cl_next
            if skip_stream is not None:
                torch.cuda.current_stream().wait_stream(skip_stream)
        ctx.metas = metas
        ctx.shape = (N, Lc, T)
        ctx.block_params = [(wf, wg, wsk, wr, bnp) for _, wf, wg, wsk, wr, bnp in blocks]
        ctx.seed = seed
        ctx.save_for_backward(*saved)
        return skip_sum

    @staticmethod
    def backward(ctx, gs):
        L_ = _lib.lib()
        saved = ctx.saved_tensors
        dev = gs.device
        gs = gs.contiguous()
        N, Lc, T = ctx.shape
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        nblocks = len(ctx.metas)
        grads = [None] * nblocks

        def dgrad(d, w, gy_cl):
            gx = torch.empty((N, d.cin, T), **f32)
            wp = _lib.ptr_array([t.data_ptr() for t in w])
            pk = F.packed_weights(w, d, PASS_DGRAD)
            F._timed("qconv_cl_fprop_kernel", F._conv_flop(d, 1, T), 1, lambda: _lib.check(
                L_.seldq_conv_dgrad(ctypes.byref(d), None, gy_cl.data_ptr(), wp, _ptr(pk), gx.data_ptr(), None, 0,
                                    _stream())))
            return gx

        # The weight gradients are off the critical path (nothing in this backward pass reads them): they run on a
        # side stream, forked behind the producers of their operands and joined at the end, so that they fill the
        # SMs the latency-bound glue kernels of the dgrad chain leave idle.
        side = _side_stream(dev) if SIDE_WGRAD else None
        keep = []                       # operands of side-stream work stay referenced until the join

        def on_side(fn, *tensors):
            if side is None:
                return fn()
            keep.extend(tensors)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                return fn()

        def wgrad(d, w, x_cl, gy_t16):
            gws, direct = F._grad_targets(w, [True] * len(w))
            gp = _lib.ptr_array([g.data_ptr() for g in gws])
            on_side(lambda: F._timed("qconv_cl_wgrad_kernel", F._conv_flop(d, 1, T), 1, lambda: _lib.check(
                L_.seldq_conv_wgrad(ctypes.byref(d), None, x_cl.data_ptr(), None, gy_t16.data_ptr(), gp, None,
                                    1 if direct else 0, None, 0, _stream()))), x_cl, gy_t16, *gws)
            return [None] * len(w) if direct else list(gws)

        def wgrad_pair(d, wa, wb, x_cl, gya_t16, gyb_t16):
            gwa, da = F._grad_targets(wa, [True] * len(wa))
            gwb, db = F._grad_targets(wb, [True] * len(wb))
            if da != db:                                   # mixed targets: keep it simple
                return wgrad(d, wa, x_cl, gya_t16), wgrad(d, wb, x_cl, gyb_t16)
            pa = _lib.ptr_array([g.data_ptr() for g in gwa])
            pb = _lib.ptr_array([g.data_ptr() for g in gwb])
            on_side(lambda: F._timed("qconv_cl_wgrad_kernel", 2 * F._conv_flop(d, 1, T), 1, lambda: _lib.check(
                L_.seldq_conv_wgrad_pair(ctypes.byref(d), x_cl.data_ptr(), gya_t16.data_ptr(), gyb_t16.data_ptr(), pa, pb,
                                         1 if da else 0, _stream()))), x_cl, gya_t16, gyb_t16, *gwa, *gwb)
            return ([None] * len(wa) if da else list(gwa)), ([None] * len(wb) if db else list(gwb))

        with torch.cuda.device(dev):
            g_rn = g_rn_cl = g_rn_t16 = None
            gs_cl = gs_t16 = None
            # BatchNorm backward reductions of the whole stack (= d beta / d gamma): one zeroed double buffer,
            # [block](pre-activation (2, L), gate (4, G)), converted to fp32 once at the end
            sizes = [2 * Lc + 4 * m[4] for m in ctx.metas]
            red = torch.zeros(sum(sizes), dtype=torch.float64, device=dev)
            fuse_l = TCN_FUSED_GLUE and bool(L_.seldq_tcn_glue_fused_supported(N, Lc, T))
            syncs = torch.zeros(2 * nblocks, dtype=torch.int32, device=dev)
            roff = [sum(sizes[:i]) for i in range(nblocks)]
            for k in reversed(range(nblocks)):
                s, d1, dsk, dre, G, U = ctx.metas[k]
                r, xa, xa_cl, yf, yg, y_cl, sums1, sums2 = saved[8 * k:8 * k + 8]
                wf, wg, wsk, wr, (g1, b1, gf, bf, gg, bg) = ctx.block_params[k]
                e1 = s["bn1"][0]
                ef = s["bnf"][0]
                if gs_cl is None:
                    gs_cl, gs_t16 = F.stage_operand(gs, dsk, 1, want_cl=True, want_t16=True)
                # conv2: gradient w.r.t. y and the weights
                gy2, gw_r = None, [None] * len(wr)
                if (TCN_PAIR and s["has_res"] and _same_geometry(dsk, dre)
                        and L_.seldq_conv_pair_supported(ctypes.byref(dsk), PASS_DGRAD)):
                    gy1, gy2 = torch.empty((N, dsk.cin, T), **f32), torch.empty((N, dsk.cin, T), **f32)
                    _conv_pair(dsk, PASS_DGRAD, gs_cl, g_rn_cl, F.packed_weights(wsk, dsk, PASS_DGRAD),
                               F.packed_weights(wr, dre, PASS_DGRAD), gy1, gy2, _epi(), _epi(), T)
                else:
                    gy1 = dgrad(dsk, wsk, gs_cl)
                    if s["has_res"]:
                        gy2 = dgrad(dre, wr, g_rn_cl)
                if s["has_res"]:
                    if dsk.cout == dre.cout:               # same geometry: both weight gradients in one launch
                        gw_sk, gw_r = wgrad_pair(dsk, wsk, wr, y_cl, gs_t16, g_rn_t16)
                    else:
                        gw_sk = wgrad(dsk, wsk, y_cl, gs_t16)
                        gw_r = wgrad(dre, wr, y_cl, g_rn_t16)
                else:
                    gw_sk = wgrad(dsk, wsk, y_cl, gs_t16)
                # gate
                bn2 = ((sums2[0], gf, bf, None, None), (sums2[1], gg, bg, None, None))
                dsg = red[roff[k] + 2 * Lc:roff[k] + 2 * Lc + 4 * G]
                _, _, clb, t16b = F._operand_info(d1, 1)
                df_cl, dg_cl = torch.empty(clb, **u8), torch.empty(clb, **u8)
                df_t16, dg_t16 = torch.empty(t16b, **u8), torch.empty(t16b, **u8)
                if fuse_l and L_.seldq_tcn_glue_fused_supported(N, G, T):
                    _glue(_lib.TCN_GATE_BWD, d1, 1, N, G, T, eps=ef, drop_p=s["drop_p"], salt=s["salt"], bn=bn2,
                          inp=(yf, yg, gy1, gy2), dsums=dsg, out_cl=(df_cl, dg_cl), out_t16=(df_t16, dg_t16),
                          seed=ctx.seed, sync=syncs[2 * k:])
                else:
                    _glue(_lib.TCN_GATE_BWD_REDUCE, None, 0, N, G, T, eps=ef, drop_p=s["drop_p"], salt=s["salt"], bn=bn2,
                          inp=(yf, yg, gy1, gy2), dsums=dsg, seed=ctx.seed)
                    _glue(_lib.TCN_GATE_BWD_APPLY, d1, 1, N, G, T, eps=ef, drop_p=s["drop_p"], salt=s["salt"], bn=bn2,
                          inp=(yf, yg, gy1, gy2), dsums=dsg, out_cl=(df_cl, dg_cl), out_t16=(df_t16, dg_t16),
                          seed=ctx.seed)
                # conv1
                if TCN_PAIR and L_.seldq_conv_pair_supported(ctypes.byref(d1), PASS_DGRAD):
                    gx1, gx2 = torch.empty((N, d1.cin, T), **f32), torch.empty((N, d1.cin, T), **f32)
                    _conv_pair(d1, PASS_DGRAD, df_cl, dg_cl, F.packed_weights(wf, d1, PASS_DGRAD),
                               F.packed_weights(wg, d1, PASS_DGRAD), gx1, gx2, _epi(), _epi(), T)
                else:
                    gx1 = dgrad(d1, wf, df_cl)
                    gx2 = dgrad(d1, wg, dg_cl)
                gw_f, gw_g = wgrad_pair(d1, wf, wg, xa_cl, df_t16, dg_t16)
                # pre-activation
                bn1 = ((sums1, g1, b1, None, None),)
                ds1 = red[roff[k]:roff[k] + 2 * Lc]
                apply_op = _lib.TCN_PREACT_BWD if fuse_l else _lib.TCN_PREACT_BWD_APPLY
                sync = syncs[2 * k + 1:] if fuse_l else None
                if not fuse_l:
                    _glue(_lib.TCN_PREACT_BWD_REDUCE, None, 0, N, Lc, T, eps=e1, bn=bn1, inp=(g_rn, gx1, gx2, xa, r),
                          dsums=ds1)
                g_r = torch.empty((N, Lc, T), **f32)
                if k > 0:
                    dprev = ctx.metas[k - 1][3]
                    _, _, clb, t16b = F._operand_info(dprev, 1)
                    g_rn_cl, g_rn_t16 = torch.empty(clb, **u8), torch.empty(t16b, **u8)
                    _glue(apply_op, dprev, 1, N, Lc, T, eps=e1, bn=bn1, inp=(g_rn, gx1, gx2, xa, r),
                          dsums=ds1, out32=g_r, out_cl=(g_rn_cl,), out_t16=(g_rn_t16,), sync=sync)
                else:
                    _glue(apply_op, None, 0, N, Lc, T, eps=e1, bn=bn1, inp=(g_rn, gx1, gx2, xa, r),
                          dsums=ds1, out32=g_r, sync=sync)
                g_rn = g_r
                grads[k] = list(gw_f) + list(gw_g) + list(gw_sk) + list(gw_r)
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)      # join: the weight gradients are complete
                keep.clear()
            redf = red.float()
            bn_grads = []
            for k in range(nblocks):
                G = ctx.metas[k][4]
                p1 = redf[roff[k]:roff[k] + 2 * Lc].view(2, Lc)
                pg = redf[roff[k] + 2 * Lc:roff[k] + 2 * Lc + 4 * G].view(4, G)
                # (gamma, beta) of batch_filter1, batch_filter2, batch_gate2
                bn_grads.append([p1[1], p1[0], pg[1], pg[0], pg[3], pg[2]])
            # with the trainer's gradient bucket in place the 6 x nblocks BatchNorm gradients are added by one
            # multi-tensor launch instead of one accumulation kernel per parameter
            bn_params = [t for k in range(nblocks) for t in ctx.block_params[k][4]]
            direct_bn = F._ACCUMULATE and all(t.grad is not None and t.grad.dtype == torch.float32 for t in bn_params)
            if direct_bn:
                torch._foreach_add_([t.grad for t in bn_params], [g for gs_k in bn_grads for g in gs_k])
                bn_grads = [[None] * 6 for _ in range(nblocks)]
        flat = []
        for k, g in enumerate(grads):
            flat += g + bn_grads[k]
        return (g_rn, None, None) + tuple(flat)


def tcn_stack_supported(blocks, x, training):
    """Fused residual-block path: CUDA, bf16 precision, training mode, Q / DQ conv1d without bias, BatchNorm on,
    stride 1 and 'same' padding, time extent a multiple of 8."""
    if not ENABLED or not (training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3):
        return False
    if F.get_precision() != "bf16" or not torch.is_grad_enabled() or x.shape[2] % 8:
        return False
    for b in blocks:
        if not hasattr(b, "batch_filter1"):
            return False
        convs = (b.conv1_filter, b.conv1_gate, b.conv2_skip, b.conv2_residual)
        alg = getattr(convs[0], "_algebra", None)
        for c in convs:
            if getattr(c, "_algebra", None) is None or c._algebra != alg or c.bias is not None or c.rotation or c.groups != 1:
                return False
            if F._pair(c.stride) != (1, 1):
                return False
            if c.in_channels < 8:                      # narrow layers take the dense path, which reads fp32 x
                return False
        k = b.conv1_filter.kernel_size
        k = int(k[-1] if isinstance(k, (tuple, list)) else k)
        dil = F._pair(b.conv1_filter.dilatation)[1]
        if not isinstance(b.conv1_filter.padding, int) or 2 * b.conv1_filter.padding != dil * (k - 1):
            return False
        if (F._pair(b.conv1_gate.dilatation)[1] != dil or b.conv1_gate.padding != b.conv1_filter.padding
                or b.conv1_gate.kernel_size != b.conv1_filter.kernel_size):
            return False
        for c in (b.conv2_skip, b.conv2_residual):
            kk = c.kernel_size
            if int(kk[-1] if isinstance(kk, (tuple, list)) else kk) != 1 or c.padding != 0:
                return False
        if b.conv1_filter.out_channels != b.conv1_gate.out_channels:
            return False
        if b.conv2_skip.out_channels > max(b.conv2_residual.out_channels, 1) * 64:
            return False
        for bn in (b.batch_filter1, b.batch_filter2, b.batch_gate2):
            if bn.momentum is None or not bn.track_running_stats or not bn.affine:
                return False
        if b.batch_filter2.eps != b.batch_gate2.eps or b.batch_filter2.momentum != b.batch_gate2.momentum:
            return False
    return True


def tcn_stack(x, blocks, seed):
    """x (N, L, T) fp32 -> sum of the blocks' skip outputs (N, U, T) fp32 (TC_Block.forward, model.py:210-216)."""
    spec, tensors, counters = [], [], []
    for k, b in enumerate(blocks):
        ws = [c._weights() for c in (b.conv1_filter, b.conv1_gate, b.conv2_skip, b.conv2_residual)]
        kk = b.conv1_filter.kernel_size
        bns = (b.batch_filter1, b.batch_filter2, b.batch_gate2)
        spec.append(dict(algebra=b.conv1_filter._algebra, nw=len(ws[0]),
                         k=int(kk[-1] if isinstance(kk, (tuple, list)) else kk),
                         dil=F._pair(b.conv1_filter.dilatation)[1], pad=int(b.conv1_filter.padding),
                         drop_p=float(b.spatial_dropout_rate), salt=101 + k, has_res=k < len(blocks) - 1,
                         bn1=(float(bns[0].eps), float(bns[0].momentum), bns[0].running_mean, bns[0].running_var),
                         bnf=(float(bns[1].eps), float(bns[1].momentum), bns[1].running_mean, bns[1].running_var),
                         bng=(float(bns[2].eps), float(bns[2].momentum), bns[2].running_mean, bns[2].running_var)))
        for w in ws:
            tensors += list(w)
        for bn in bns:
            tensors += [bn.weight, bn.bias]
            counters.append(bn)
    _bump_counters(counters)
    return _TcnStack.apply(x, spec, seed, *tensors)
