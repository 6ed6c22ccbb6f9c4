"""Fused training-mode paths through the glue between the convolutions.

`cnn_stack` runs the whole CNN front of ConvTC_Block (model.py:261-285: per block Q/DQ conv2d ->
BatchNorm2d -> ReLU -> MaxPool2d([p, 1]) -> Dropout) as one autograd Function over the C ABI:

    forward, per block    conv (tcgen05, bf16 NCHW output)  ->  bn_stats  ->  bn_finalize  ->  cnn_tail_fwd
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
                y16 = torch.empty((N, C, oh, ow), dtype=torch.bfloat16, device=dev)
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
                if last:
                    z32 = torch.empty((N, C, hp, ow), dtype=torch.float32, device=dev)
                    nxt_cl, consumer = None, None
                else:
                    consumer = descs[k + 1][0]
                    nxt_cl = torch.empty(F._operand_info(consumer, 0)[2], dtype=torch.uint8, device=dev)
                F._timed("cnn_tail_fwd_kernel", 0.0, 1, lambda: _lib.check(
                    L.seldq_cnn_tail_fwd(ctypes.byref(td), None if consumer is None else ctypes.byref(consumer),
                                         y16.data_ptr(), coef.data_ptr(), _ptr(seed) if s["drop_p"] > 0 else None,
                                         _ptr(nxt_cl), _ptr(z32), idx.data_ptr(), _stream())))
                # block 0 of a narrow first layer keeps the fp32 input for its weight gradient (seldq.h)
                keep_in = x if (k == 0 and x_dense) else cur_cl
                saved += [keep_in, y16, coef, idx]
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
                xin, y16, coef, idx = saved[4 * k:4 * k + 4]
                ws, gamma, beta = ctx.block_params[k]
                N, C = d.batch, d.cout
                need_gx = k > 0 or ctx.needs_input_grad[0]
                _, _, clb, t16b = F._operand_info(d, 1)
                d_t16 = torch.empty(t16b, dtype=torch.uint8, device=dev)
                d_cl = torch.empty(clb, dtype=torch.uint8, device=dev) if need_gx else None
                dsums = torch.zeros((C * 3,), dtype=torch.float64, device=dev)
                F._timed("cnn_tail_bwd_kernels", 0.0, 3, lambda: _lib.check(
                    L.seldq_cnn_tail_bwd(ctypes.byref(td), ctypes.byref(d), y16.data_ptr(), coef.data_ptr(),
                                         idx.data_ptr(), gz.data_ptr(), dsums.data_ptr(), d_t16.data_ptr(), _ptr(d_cl),
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
        bn.num_batches_tracked.add_(1)
    return _CnnStack.apply(x, spec, seed, *tensors)
