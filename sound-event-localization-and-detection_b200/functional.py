"""Differentiable ops over the C ABI (libseldq.so): quaternion / dual-quaternion convolution and
linear, and the STFT magnitude/phase front end.  This is the host-side mirror of the reference's
functional layer (quaternion/quaternion_ops.py, dual_quaternion/dual_quaternion_ops.py): the
public names with the reference's exact signatures live in the drop-in modules next to this file;
they all funnel into the autograd Functions below.

PyTorch supplies device memory, the current stream and autograd plumbing; every FLOP of the ops
themselves runs in the hand-written sm_100a kernels.  There is no CPU path: CPU tensors raise.
"""
import contextlib
import os

import torch

from . import _lib
from ._lib import ALG_DQ, ALG_DQ_LINEAR, ALG_DQ_LINEAR_IO, ALG_Q, ALG_Q_LINEAR_IO, ALG_REAL, PASS_DGRAD, PASS_FWD, PASS_WGRAD, PREC_BF16, PREC_FP32

_PRECISION = {"fp32": PREC_FP32, "bf16": PREC_BF16}[os.environ.get("SELDQ_PRECISION", "bf16").lower()]
_NCOMP = {ALG_REAL: 1, ALG_Q: 4, ALG_DQ: 8, ALG_DQ_LINEAR: 8, ALG_Q_LINEAR_IO: 4, ALG_DQ_LINEAR_IO: 8}
_IO_LAYOUT = (ALG_Q_LINEAR_IO, ALG_DQ_LINEAR_IO)      # compact tensors stored (in/nc, out/nc): the linear layers' own


def _apply_tf32_policy():
    """The real-valued layers around the hot path (MultiHeadAttention projections / attention products, the final
    nn.Linear heads) are PyTorch ops.  In 'fp32' mode they keep the reference's true-fp32 arithmetic (the reference
    disables cuDNN, model.py:10, so no TF32 is involved); in 'bf16' mode -- the tensor-core mode, gated at rel 2e-2 --
    they may use TF32 tensor cores (rel ~5e-4), which replaces their SIMT sgemm kernels."""
    on = _PRECISION == PREC_BF16
    torch.backends.cuda.matmul.allow_tf32 = on
    torch.backends.cudnn.allow_tf32 = on


def set_precision(name):
    """'fp32': FFMA kernels, parity gate rel 1e-4.  'bf16': tcgen05 tensor-core kernels (bf16
    operands, fp32 accumulation), parity gate rel 2e-2.  Returns the previous setting."""
    global _PRECISION
    prev = get_precision()
    _PRECISION = {"fp32": PREC_FP32, "bf16": PREC_BF16}[name]
    _apply_tf32_policy()
    return prev


def get_precision():
    return "fp32" if _PRECISION == PREC_FP32 else "bf16"


@contextlib.contextmanager
def precision(name):
    prev = set_precision(name)
    try:
        yield
    finally:
        set_precision(prev)


# ---- gradient accumulation straight into preallocated .grad buffers (trainer.FlatGradBucket) ----------
_ACCUMULATE = False


def set_grad_accumulation(flag):
    """True: the weight-gradient kernels ADD their result into `param.grad` when that buffer already exists
    (the trainer's flat bucket) and autograd receives None for the parameter, which removes one zero-fill and
    one add kernel per parameter and step.  Returns the previous setting."""
    global _ACCUMULATE
    prev, _ACCUMULATE = _ACCUMULATE, bool(flag)
    return prev


def _grad_targets(params, needs):
    """(buffers, direct): param.grad buffers to accumulate into when every parameter that needs a gradient has
    a contiguous fp32 one; otherwise fresh buffers that autograd accumulates as usual."""
    if _ACCUMULATE and all((not n) or (p.is_leaf and p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32
                                       and p.grad.device == p.device) for p, n in zip(params, needs)) and all(needs):
        return [p.grad for p in params], True
    return [torch.empty_like(p) for p in params], False


# ---- optional per-kernel timing (bench.py): CUDA events around each C-ABI call on the launching stream
_PROF = None


def profile_reset(enable=True):
    global _PROF
    _PROF = {"calls": [], "launches": 0} if enable else None


def _timed(kernel, flop, launches, fn):
    """Runs fn() (one C-ABI call); when profiling, brackets it with events and books `launches`
    kernel launches and `flop` algorithmic FLOPs under `kernel`."""
    if _PROF is None:
        return fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    _PROF["calls"].append((kernel, flop, a, b))
    _PROF["launches"] += launches
    return out


def profile_collect():
    """{'launches': total kernel launches, 'kernels': {name: {ms, flop, launches}}} since profile_reset."""
    torch.cuda.synchronize()
    out = {"launches": 0 if _PROF is None else _PROF["launches"], "kernels": {}}
    for kernel, flop, a, b in ([] if _PROF is None else _PROF["calls"]):
        k = out["kernels"].setdefault(kernel, {"name": kernel, "ms": 0.0, "flop": 0.0, "launches": 0})
        k["ms"] += a.elapsed_time(b)
        k["flop"] += flop
        k["launches"] += 1
    return out


def _conv_flop(desc, oh, ow):
    """Algorithmic FLOPs of one pass: only the non-zero blocks of the expanded weight count
    (DQ: 0.75 of dense)."""
    nz = 0.75 if desc.algebra in (ALG_DQ, ALG_DQ_LINEAR, ALG_DQ_LINEAR_IO) else 1.0
    return 2.0 * nz * desc.cout * desc.cin * desc.k_h * desc.k_w * desc.batch * oh * ow


def _require_cuda_f32(t, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % what)
    if not t.is_cuda:
        raise RuntimeError("seldq: %s is on %s; the sm_100a kernels have no CPU fallback" % (what, t.device))
    if t.dtype != torch.float32:
        raise TypeError("seldq: %s must be float32 (got %s)" % (what, t.dtype))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _pair(v):
    if isinstance(v, (tuple, list)):
        if len(v) == 1:
            return int(v[0]), int(v[0])
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _operand_info(desc, which):
    """(padded channels, dense?, CL operand bytes, T16 operand bytes) of x (which=0) or gy (which=1)."""
    import ctypes
    cp, dense = ctypes.c_int32(), ctypes.c_int32()
    clb, t16b = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.check(_lib.lib().seldq_conv_operand_info(ctypes.byref(desc), which, ctypes.byref(cp), ctypes.byref(dense),
                                                  ctypes.byref(clb), ctypes.byref(t16b)))
    return cp.value, bool(dense.value), clb.value, t16b.value


def stage_operand(t, desc, which, want_cl=True, want_t16=False):
    """bf16 operand copies of an NCW / NCHW fp32 tensor for the tensor-core kernels (include/seldq.h):
    the channels-last operand (forward / dgrad / wgrad's x) and / or the pitched NCHW copy (wgrad's gy).
    Returns (cl, t16) uint8 buffers (None where not requested)."""
    import ctypes
    _, _, clb, t16b = _operand_info(desc, which)
    cl = torch.empty(clb, dtype=torch.uint8, device=t.device) if want_cl else None
    t16 = torch.empty(t16b, dtype=torch.uint8, device=t.device) if want_t16 else None
    _timed("stage_operand_kernel", 0.0, 1, lambda: _lib.check(
        _lib.lib().seldq_stage_operand(ctypes.byref(desc), which, t.data_ptr(), _ptr(cl), _ptr(t16), _stream())))
    return cl, t16


# packed bf16 weight tiles, one set per (layer, pass); re-packed when a weight's version counter moves
# (optimizer.step updates in place) and always while a CUDA graph is being captured, so that a replayed
# step packs the weights it is about to use
# Entries hold their weights through weak references (a model that is dropped frees its tiles, and repack_all()
# never re-packs dead models).  Validity = (version counters, epoch): writes through `.data` (dist.broadcast of
# p.data, EMA swaps, clipping via p.data.copy_) do NOT move a version counter -- such callers must call
# invalidate_packed_weights(), as trainer.broadcast_parameters / FlatGradBucket do.
_PACKED = {}
_PACK_EPOCH = 0


def _live_weights(info):
    """The weight tensors of a cache entry, or None when any of them has been freed."""
    ws = tuple(r() for r in info[0])
    return None if any(w is None for w in ws) else ws


def _evict_dead():
    global _PACK_TABLE
    dead = [k for k, ent in _PACKED.items() if _live_weights(ent[2]) is None]
    for k in dead:
        del _PACKED[k]
    if dead:
        _PACK_TABLE = None


def invalidate_packed_weights():
    """Forget every packed weight set.  For callers that update weights through storage the parameters' version
    counters do not see (trainer.FlatGradBucket steps ONE flat tensor the parameters are views of)."""
    global _PACK_EPOCH
    _PACK_EPOCH += 1


def packed_weights(weights, desc, pass_, cache=True):
    """cache=False: weights that are temporaries (their address says nothing about their content) are packed into
    a fresh buffer on every call."""
    import ctypes
    L = _lib.lib()
    nbytes = L.seldq_conv_packed_bytes(ctypes.byref(desc), pass_)
    if nbytes == 0:
        return None
    if not cache:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=weights[0].device)
        wp = _lib.ptr_array([w.data_ptr() for w in weights])
        _timed("pack_weights_kernel", 0.0, 1, lambda: _lib.check(
            L.seldq_conv_pack_weights(ctypes.byref(desc), pass_, wp, buf.data_ptr(), _stream())))
        return buf
    key = (tuple(w.data_ptr() for w in weights), pass_, desc.algebra, desc.cin, desc.cout, desc.k_h, desc.k_w)
    versions = tuple(w._version for w in weights)
    ent = _PACKED.get(key)
    if ent is not None:
        if _live_weights(ent[2]) is None:
            ent = None                  # the entry's tensors were freed: the address now belongs to something else
    capturing = torch.cuda.is_current_stream_capturing()
    if ent is not None and ent[1].numel() == nbytes and ent[0] == versions and ent[3] == _PACK_EPOCH:
        # valid as long as the version counters stand still; while a graph is being captured only sets that
        # repack_all() refreshes (epoch > 0) may be reused, because the replayed step must see current weights
        if not capturing or ent[3] > 0:
            return ent[1]
    buf = ent[1] if ent is not None and ent[1].numel() == nbytes else torch.empty(
        nbytes, dtype=torch.uint8, device=weights[0].device)
    wp = _lib.ptr_array([w.data_ptr() for w in weights])
    _timed("pack_weights_kernel", 0.0, 1, lambda: _lib.check(
        L.seldq_conv_pack_weights(ctypes.byref(desc), pass_, wp, buf.data_ptr(), _stream())))
    if ent is None or ent[1] is not buf:
        global _PACK_TABLE
        _PACK_TABLE = None                              # a new buffer: the multi-pack table must be rebuilt
    import weakref
    _PACKED[key] = (versions, buf, (tuple(weakref.ref(w) for w in weights), _lib.ConvDesc.from_buffer_copy(desc), pass_),
                    _PACK_EPOCH)
    return buf


_PACK_TABLE = None


def repack_all():
    """Re-packs every weight set packed so far in ONE launch (seldq_conv_pack_table_run) and marks them fresh: the
    call a trainer makes right after its optimiser step, so that the next step's layers find their bf16 tiles
    ready (and a captured step holds one pack launch instead of one per layer and pass)."""
    import ctypes
    global _PACK_TABLE, _PACK_EPOCH
    _evict_dead()
    if not _PACKED:
        return
    L = _lib.lib()
    if _PACK_TABLE is None:
        esz = L.seldq_conv_pack_table_entry_bytes()
        keys = list(_PACKED)
        host = torch.zeros(len(keys) * esz, dtype=torch.uint8)
        max_items = 0
        for i, k in enumerate(keys):
            _, buf, info, _ = _PACKED[k]
            ws, desc, pass_ = _live_weights(info), info[1], info[2]
            wp = _lib.ptr_array([w.data_ptr() for w in ws])
            items = ctypes.c_int32()
            _lib.check(L.seldq_conv_pack_table_fill(ctypes.byref(desc), pass_, wp, buf.data_ptr(),
                                                    host.data_ptr() + i * esz, ctypes.byref(items)))
            max_items = max(max_items, items.value)
        dev = next(iter(_PACKED.values()))[1].device
        _PACK_TABLE = (host.to(dev), len(keys), max_items)
    table, count, max_items = _PACK_TABLE
    _timed("pack_weights_kernel", 0.0, 1, lambda: _lib.check(
        L.seldq_conv_pack_table_run(table.data_ptr(), count, max_items, _stream())))
    _PACK_EPOCH += 1
    for k, (ver, buf, info, _) in list(_PACKED.items()):
        _PACKED[k] = (tuple(w._version for w in _live_weights(info)), buf, info, _PACK_EPOCH)


def _conv_desc(algebra, prec, x_shape, cout, ksize, stride, padding, dilation):
    if len(x_shape) == 3:
        n, c, w = x_shape
        return _lib.ConvDesc(algebra, prec, 1, n, c, cout, 1, w, 1, ksize[-1], 1, stride[-1], 0, padding[-1], 1,
                             dilation[-1])
    n, c, h, w = x_shape
    return _lib.ConvDesc(algebra, prec, 2, n, c, cout, h, w, ksize[0], ksize[1], stride[0], stride[1],
                         padding[0], padding[1], dilation[0], dilation[1])


_PACK_CACHE = True      # False while _linear_as_conv builds its node: its weights are per-call temporaries
_CONV_BWD_FORK = os.environ.get("SELDQ_CONV_BWD_FORK", "1") != "0"
_BWD_SIDE = {}


def _bwd_side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _BWD_SIDE:
        _BWD_SIDE[key] = torch.cuda.Stream(device=dev)
    return _BWD_SIDE[key]


class _BlockConv(torch.autograd.Function):
    """y = conv(x, expand(weights)) + bias  with the expansion fused into the kernels
    (reference: quaternion_ops.py:125-147, dual_quaternion_ops.py:111-153)."""

    @staticmethod
    def forward(ctx, x, bias, stride, padding, dilation, algebra, prec, *weights):
        L = _lib.lib()
        nc = _NCOMP[algebra]
        _require_cuda_f32(x, "input")
        for w in weights:
            _require_cuda_f32(w, "weight")
        if bias is not None:
            _require_cuda_f32(bias, "bias")
        x = x.contiguous()
        weights = tuple(w.contiguous() for w in weights)
        w0 = weights[0]
        if x.dim() == 3:
            stride, padding, dilation = (1, _pair(stride)[1]), (0, _pair(padding)[1]), (1, _pair(dilation)[1])
        else:
            stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
        io = algebra in _IO_LAYOUT
        ksize = (1,) * (x.dim() - 2) if io else tuple(w0.shape[2:])
        cout = (w0.shape[1] if io else w0.shape[0]) * nc
        if (w0.shape[0] if io else w0.shape[1]) * nc != x.shape[1]:
            raise RuntimeError("Given groups=1, weight of size %s (x%d components), expected input%s to have %d "
                               "channels, but got %d channels instead"
                               % (list(w0.shape), nc, list(x.shape), w0.shape[1] * nc, x.shape[1]))
        desc = _conv_desc(algebra, prec, tuple(x.shape), cout, ksize, stride, padding, dilation)
        import ctypes
        oh, ow = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(L.seldq_conv_out_shape(ctypes.byref(desc), ctypes.byref(oh), ctypes.byref(ow)))
        out_shape = (x.shape[0], cout, ow.value) if x.dim() == 3 else (x.shape[0], cout, oh.value, ow.value)
        y = torch.empty(out_shape, dtype=torch.float32, device=x.device)
        wp = _lib.ptr_array([w.data_ptr() for w in weights])
        bf16 = prec == PREC_BF16
        with torch.cuda.device(x.device):
            x_cl = pk = None
            if bf16:
                x_cl, _ = stage_operand(x, desc, 0)
                pk = packed_weights(weights, desc, PASS_FWD, cache=_PACK_CACHE)
            kern = "qconv_cl_fprop_kernel" if bf16 else "conv_simt_kernel"
            _timed(kern, _conv_flop(desc, oh.value, ow.value), 1, lambda: _lib.check(
                L.seldq_conv_fwd(ctypes.byref(desc), x.data_ptr(), _ptr(x_cl), wp, _ptr(pk), _ptr(bias),
                                 y.data_ptr(), None, 0, _stream())))
        ctx.desc = desc
        ctx.pack_cache = _PACK_CACHE
        ctx.out_hw = (oh.value, ow.value)
        ctx.has_bias = bias is not None
        # the tensor-core wgrad reads the channels-last bf16 operand only: keep that instead of the fp32
        # input (except for the narrow first layer, whose wgrad stages its own copy of the fp32 input)
        ctx.x_is_cl = bf16 and not _operand_info(desc, 0)[1]
        ctx.bias_ref = bias
        ctx.save_for_backward(x_cl if ctx.x_is_cl else x, *weights)
        return y

    @staticmethod
    def backward(ctx, gy):
        import ctypes
        L = _lib.lib()
        desc = ctx.desc
        xs, *weights = ctx.saved_tensors
        bf16 = desc.precision == PREC_BF16
        gy = gy.contiguous()
        if gy.dtype != torch.float32:
            raise TypeError("seldq: grad_output must be float32")
        dev = gy.device
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[7:])
        need_b = ctx.has_bias and ctx.needs_input_grad[1]
        gx = gb = None
        gws = [None] * len(weights)
        wp = _lib.ptr_array([w.data_ptr() for w in weights])
        with torch.cuda.device(dev):
            need_w_any = need_w or need_b
            gy_cl = gy_t16 = None
            if bf16 and (need_x or need_w_any):
                gy_cl, gy_t16 = stage_operand(gy, desc, 1, want_cl=need_x, want_t16=need_w_any)
            fork_ev = None
            if bf16 and need_x and need_w_any and _CONV_BWD_FORK and _PROF is None:
                fork_ev = torch.cuda.current_stream().record_event()       # operands staged: the weight gradient may start here
            if need_x:
                if desc.ndim == 1:
                    gx = torch.empty((desc.batch, desc.cin, desc.in_w), dtype=torch.float32, device=dev)
                else:
                    gx = torch.empty((desc.batch, desc.cin, desc.in_h, desc.in_w), dtype=torch.float32, device=dev)
                pk = packed_weights(weights, desc, PASS_DGRAD, cache=ctx.pack_cache) if bf16 else None
                kern = "qconv_cl_fprop_kernel" if bf16 else "conv_simt_kernel"
                _timed(kern, _conv_flop(desc, *ctx.out_hw), 1, lambda: _lib.check(
                    L.seldq_conv_dgrad(ctypes.byref(desc), gy.data_ptr(), _ptr(gy_cl), wp, _ptr(pk), gx.data_ptr(),
                                       None, 0, _stream())))
            if need_w_any:
                gws, direct = _grad_targets(weights, ctx.needs_input_grad[7:])
                gb = None
                if need_b:
                    gb = ctx.bias_ref.grad if direct and ctx.bias_ref.grad is not None else torch.empty(
                        desc.cout, dtype=torch.float32, device=dev)
                    direct_b = direct and gb is ctx.bias_ref.grad
                    if direct and not direct_b:       # mixed case: keep it simple, overwrite semantics for all
                        gws, direct = [torch.empty_like(w) for w in weights], False
                gp = _lib.ptr_array([g.data_ptr() for g in gws])
                x_cl = xs if ctx.x_is_cl else None
                x32 = None if ctx.x_is_cl else xs
                work = None
                if bf16 and not ctx.x_is_cl:      # narrow first layer: the library stages x itself
                    work = torch.empty(L.seldq_conv_workspace_bytes(ctypes.byref(desc), PASS_WGRAD), dtype=torch.uint8,
                                       device=dev)
                kern = "qconv_cl_wgrad_kernel" if bf16 else "wgrad_simt_kernel"

                def run_wgrad():
                    _timed(kern, _conv_flop(desc, *ctx.out_hw), 1 + (1 if need_b else 0), lambda: _lib.check(
                        L.seldq_conv_wgrad(ctypes.byref(desc), _ptr(x32), _ptr(x_cl), gy.data_ptr(), _ptr(gy_t16), gp,
                                           _ptr(gb), 1 if direct else 0, _ptr(work),
                                           0 if work is None else work.numel(), _stream())))

                # The input-gradient and weight-gradient passes are independent; on the small layers of the TC tail
                # neither fills the SMs (76 and 40 work units on 148 SMs), so with both wanted the weight gradient
                # runs on a forked stream next to the dgrad launch above and is joined before this node returns
                # (every tensor it touches outlives the join).  SELDQ_CONV_BWD_FORK=0: one after the other.
                if bf16 and need_x and _CONV_BWD_FORK and _PROF is None:
                    cur = torch.cuda.current_stream()
                    side = _bwd_side_stream(dev)
                    side.wait_event(fork_ev)
                    with torch.cuda.stream(side):
                        run_wgrad()
                    cur.wait_stream(side)
                else:
                    run_wgrad()
                if direct:                            # already added into param.grad
                    gws, gb = [None] * len(weights), None
        return (gx, gb, None, None, None, None, None) + tuple(gws)


class _BlockConvTranspose(torch.autograd.Function):
    """y = conv_transpose(x, expand(weights)) + bias (quaternion_ops.py:149-172).
    The expanded (in, out, k...) weight of the transposed convolution has the block table of the convolution's, so
    the operator IS the input-gradient pass of the convolution -- same stride, padding, dilation -- whose compact weights
    are these same tensors read as (out', in') = (in, out): forward = seldq_conv_dgrad, gradient w.r.t. x =
    seldq_conv_fwd, weight gradient = seldq_conv_wgrad with the roles of the two activations swapped.  No new kernel.
    output_padding only picks the output size among those the strided convolution maps onto x's."""

    @staticmethod
    def forward(ctx, x, bias, stride, padding, output_padding, dilation, algebra, prec, *weights):
        import ctypes
        L = _lib.lib()
        nc = _NCOMP[algebra]
        _require_cuda_f32(x, "input")
        for w in weights:
            _require_cuda_f32(w, "weight")
        x = x.contiguous()
        weights = tuple(w.contiguous() for w in weights)
        w0 = weights[0]
        nd = x.dim() - 2
        if w0.shape[0] * nc != x.shape[1]:
            raise RuntimeError("Given transposed=1, weight of size %s (x%d components), expected input%s to have %d channels, "
                               "but got %d channels instead" % (list(w0.shape), nc, list(x.shape), w0.shape[0] * nc, x.shape[1]))
        pad = (0, _pair(padding)[1]) if nd == 1 else _pair(padding)
        dil = (1, _pair(dilation)[1]) if nd == 1 else _pair(dilation)
        std = (1, _pair(stride)[1]) if nd == 1 else _pair(stride)
        opad = (0, _pair(output_padding)[1]) if nd == 1 else _pair(output_padding)
        ks = (1, w0.shape[2]) if nd == 1 else tuple(w0.shape[2:])
        cout = w0.shape[1] * nc
        in_sp = (1, x.shape[2]) if nd == 1 else tuple(x.shape[2:])
        if any(opad[i] >= max(std[i], dil[i]) for i in range(2) if opad[i]):
            raise RuntimeError("output padding must be smaller than either stride or dilation, but got output_padding=%s, "
                               "stride=%s, dilation=%s" % (list(opad[2 - nd:]), list(std[2 - nd:]), list(dil[2 - nd:])))
        out_sp = tuple((in_sp[i] - 1) * std[i] + (ks[i] - 1) * dil[i] - 2 * pad[i] + opad[i] + 1 for i in range(2))
        if min(out_sp) < 1:
            raise RuntimeError("transposed convolution: output size is too small")
        # the convolution this operator is the input gradient of: (N, cout, out_sp) -> (N, cin, in_sp)
        desc = _lib.ConvDesc(algebra, prec, nd, x.shape[0], cout, x.shape[1], out_sp[0], out_sp[1], ks[0], ks[1], std[0], std[1],
                             pad[0], pad[1], dil[0], dil[1])
        oh, ow = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(L.seldq_conv_out_shape(ctypes.byref(desc), ctypes.byref(oh), ctypes.byref(ow)))
        if (oh.value, ow.value) != tuple(in_sp):
            raise RuntimeError("transposed convolution: inconsistent output_padding %s for stride %s"
                               % (list(opad[2 - nd:]), list(std[2 - nd:])))
        y = torch.empty((x.shape[0], cout) + (out_sp[1:] if nd == 1 else out_sp), dtype=torch.float32, device=x.device)
        wp = _lib.ptr_array([w.data_ptr() for w in weights])
        with torch.cuda.device(x.device):
            work = torch.empty(max(1, L.seldq_conv_workspace_bytes(ctypes.byref(desc), PASS_DGRAD)), dtype=torch.uint8,
                               device=x.device)
            _timed("qconv_cl_fprop_kernel" if prec == PREC_BF16 else "conv_simt_kernel", 0.0, 1, lambda: _lib.check(
                L.seldq_conv_dgrad(ctypes.byref(desc), x.data_ptr(), None, wp, None, y.data_ptr(), work.data_ptr(),
                                   work.numel(), _stream())))
            if bias is not None:
                y += bias.view((1, -1) + (1,) * nd)
        ctx.desc, ctx.nd, ctx.has_bias = desc, nd, bias is not None
        ctx.save_for_backward(x, *weights)
        return y

    @staticmethod
    def backward(ctx, gy):
        import ctypes
        L = _lib.lib()
        desc = ctx.desc
        x, *weights = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gb = None
        gws = [None] * len(weights)
        wp = _lib.ptr_array([w.data_ptr() for w in weights])
        with torch.cuda.device(gy.device):
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                work = torch.empty(max(1, L.seldq_conv_workspace_bytes(ctypes.byref(desc), PASS_FWD)), dtype=torch.uint8,
                                   device=gy.device)
                _lib.check(L.seldq_conv_fwd(ctypes.byref(desc), gy.data_ptr(), None, wp, None, None, gx.data_ptr(),
                                            work.data_ptr(), work.numel(), _stream()))
            if any(ctx.needs_input_grad[8:]):
                gws = [torch.empty_like(w) for w in weights]
                gp = _lib.ptr_array([g.data_ptr() for g in gws])
                work = torch.empty(max(1, L.seldq_conv_workspace_bytes(ctypes.byref(desc), PASS_WGRAD)), dtype=torch.uint8,
                                   device=gy.device)
                # <conv_transpose(x; W), gy> = <x, conv(gy; W)>: the convolution's "input" is gy, its "output gradient" x
                _lib.check(L.seldq_conv_wgrad(ctypes.byref(desc), gy.data_ptr(), None, x.data_ptr(), None, gp, None, 0,
                                              work.data_ptr(), work.numel(), _stream()))
            if ctx.has_bias and ctx.needs_input_grad[1]:
                gb = gy.sum(dim=[0] + list(range(2, gy.dim())))
        return (gx, gb, None, None, None, None, None, None) + tuple(gws)


def block_conv_transpose(x, weights, bias, stride, padding, output_padding, groups, dilation, algebra, prec=None):
    """quaternion_transpose_conv (quaternion_ops.py:149-172) on the convolution kernels: groups 1; stride 1 on either
    path, stride > 1 on the fp32 kernels (the tensor-core path implements stride 1)."""
    if x.dim() not in (3, 4):
        if x.dim() == 5:
            raise NotImplementedError("seldq: 3-d transposed convolution (5-d input) is not implemented")
        raise Exception("The convolutional input is either 3, 4 or 5 dimensions. input.dim = " + str(x.dim()))
    if groups != 1:
        raise NotImplementedError("seldq: groups != 1 is not implemented")
    prec = _PRECISION if prec is None else prec
    if prec == PREC_BF16 and (min(weights[0].shape[0], weights[0].shape[1]) < 8 or _pair(stride) != (1, 1)):
        prec = PREC_FP32          # narrow layers: the tensor-core path's dense mode serves forward convolutions only
    return _BlockConvTranspose.apply(x, bias, stride, padding, output_padding, dilation, algebra, prec, *weights)


class _BlockLinear(torch.autograd.Function):
    """y = x @ expand(weights) + bias  (quaternion_ops.py:299-327 / :392-464,
    dual_quaternion_ops.py:156-203).  x is (rows, in)."""

    @staticmethod
    def forward(ctx, x, bias, algebra, prec, *weights):
        import ctypes
        L = _lib.lib()
        nc = _NCOMP[algebra]
        _require_cuda_f32(x, "input")
        for w in weights:
            _require_cuda_f32(w, "weight")
        if bias is not None:
            _require_cuda_f32(bias, "bias")
        if x.dim() != 2:
            raise RuntimeError("seldq linear expects a 2-d (rows, features) input")
        x = x.contiguous()
        weights = tuple(w.contiguous() for w in weights)
        fin, fout = weights[0].shape[0] * nc, weights[0].shape[1] * nc
        if x.shape[1] != fin:
            raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)"
                               % (x.shape[0], x.shape[1], fin, fout))
        desc = _lib.LinearDesc(algebra, prec, x.shape[0], fin, fout)
        y = torch.empty((x.shape[0], fout), dtype=torch.float32, device=x.device)
        wp = _lib.ptr_array([w.data_ptr() for w in weights])
        with torch.cuda.device(x.device):
            _timed("linear_simt", 0.0, 1, lambda: _lib.check(
                L.seldq_linear_fwd(ctypes.byref(desc), x.data_ptr(), wp, _ptr(bias), y.data_ptr(), None, 0,
                                   _stream())))
        ctx.desc = desc
        ctx.has_bias = bias is not None
        ctx.bias_ref = bias
        ctx.save_for_backward(x, *weights)
        return y

    @staticmethod
    def backward(ctx, gy):
        import ctypes
        L = _lib.lib()
        desc = ctx.desc
        x, *weights = ctx.saved_tensors
        gy = gy.contiguous()
        dev = gy.device
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[4:])
        need_b = ctx.has_bias and ctx.needs_input_grad[1]
        gx = gb = None
        gws = [None] * len(weights)
        wp = _lib.ptr_array([w.data_ptr() for w in weights])
        with torch.cuda.device(dev):
            if need_x:
                gx = torch.empty_like(x)
                _timed("linear_simt", 0.0, 1, lambda: _lib.check(
                    L.seldq_linear_dgrad(ctypes.byref(desc), gy.data_ptr(), wp, gx.data_ptr(), None, 0,
                                         _stream())))
            if need_w or need_b:
                gws, direct = _grad_targets(weights, ctx.needs_input_grad[4:])
                gb = None
                if need_b:
                    gb = ctx.bias_ref.grad if direct and ctx.bias_ref.grad is not None else torch.empty(
                        desc.out_features, dtype=torch.float32, device=dev)
                    if direct and gb is not ctx.bias_ref.grad:
                        gws, direct = [torch.empty_like(w) for w in weights], False
                gp = _lib.ptr_array([g.data_ptr() for g in gws])
                _timed("linear_simt", 0.0, 1 + (1 if need_b else 0), lambda: _lib.check(
                    L.seldq_linear_wgrad(ctypes.byref(desc), x.data_ptr(), gy.data_ptr(), gp, _ptr(gb),
                                         1 if direct else 0, None, 0, _stream())))
                if direct:
                    gws, gb = [None] * len(weights), None
        return (gx, gb, None, None) + tuple(gws)


def block_conv(x, weights, bias, stride, padding, dilation, algebra, prec=None):
    if x.dim() not in (3, 4):
        if x.dim() == 5:
            raise NotImplementedError("seldq: 3-d convolution (5-d input) is not on the SELD hot path")
        # same complaint as quaternion_ops.py:143-145
        raise Exception("The convolutional input is either 3, 4 or 5 dimensions. input.dim = " + str(x.dim()))
    prec = _PRECISION if prec is None else prec
    return _BlockConv.apply(x, bias, stride, padding, dilation, algebra, prec, *weights)


def _linear_as_conv(x, weights, bias, algebra):
    """(rows, in) @ expand(weights) on the tensor-core path: x is transposed into a (1, in, rows) NCW tensor and the
    layer runs as a 1x1 convolution that reads the layer's OWN compact tensors in their (in/nc, out/nc) layout
    (SELDQ_ALG_Q_LINEAR_IO / SELDQ_ALG_DQ_LINEAR_IO: the block table of quaternion_linear / dual_quaternion_linear):
    no transposed weight copies, the packed tiles live in the cache the trainer re-packs in one launch, and the
    weight gradient is accumulated straight into the parameters' gradient buffers."""
    conv_alg = ALG_DQ_LINEAR_IO if algebra == ALG_DQ else ALG_Q_LINEAR_IO
    y = _BlockConv.apply(x.t().contiguous().unsqueeze(0), bias, 1, 0, 1, conv_alg, PREC_BF16, *weights)
    return y[0].t()


def block_linear(x, weights, bias, algebra, prec=None):
    prec = _PRECISION if prec is None else prec
    if prec == PREC_BF16 and x.is_cuda and weights[0].shape[0] >= 8 and weights[0].shape[1] >= 8:
        lead = x.shape[:-1]
        y = _linear_as_conv(x.reshape(-1, x.shape[-1]), weights, bias, algebra)
        return y.reshape(lead + (y.shape[-1],))
    if x.dim() == 2:
        return _BlockLinear.apply(x, bias, algebra, prec, *weights)
    lead = x.shape[:-1]
    y = _BlockLinear.apply(x.reshape(-1, x.shape[-1]), bias, algebra, prec, *weights)
    return y.reshape(lead + (y.shape[-1],))


# ---- rotation variants and quaternion point-wise operators -------------------------------------------
# The real-algebra contraction of the rotation variants follows the global precision where the tensor-core path serves
# it (_real_prec below: stride 1, 8 ... 256 channels on both sides -- the kernels' dense mode, which builds the expanded
# bf16 tile from the fp32 weight in its prologue) and runs the fp32 kernels otherwise; SELDQ_ROTATION_BF16=0: always the
# fp32 kernels.
_ROTATION_BF16 = os.environ.get("SELDQ_ROTATION_BF16", "1") != "0"


class _RotationWeight(torch.autograd.Function):
    """(r, i, j, k) of shape (d0, d1, k...) -> the real weight (nc d0, nc d1, k...) the reference's rotation variants
    build (quaternion_ops.py:188-220, :249-281, :344-376), or its (nc d1, nc d0, k...) transpose (csrc/rotation.cu);
    nc = 4 with quaternion_format, else 3."""

    @staticmethod
    def forward(ctx, quaternion_format, transpose_out, *weights):
        if len(weights) != 4:
            raise ValueError("rotation weight: expected the four tensors r, i, j, k")
        for w in weights:
            _require_cuda_f32(w, "weight")
            if w.shape != weights[0].shape or w.dim() < 2:
                raise RuntimeError("rotation weight: r, i, j, k must share one shape of at least 2 dimensions")
        weights = tuple(w.contiguous() for w in weights)
        w0 = weights[0]
        d0, d1 = w0.shape[0], w0.shape[1]
        taps = 1
        for k in w0.shape[2:]:
            taps *= k
        nc = 4 if quaternion_format else 3
        lead = (nc * d1, nc * d0) if transpose_out else (nc * d0, nc * d1)
        out = torch.empty(lead + tuple(w0.shape[2:]), dtype=torch.float32, device=w0.device)
        if out.numel():
            with torch.cuda.device(w0.device):
                _lib.check(_lib.lib().seldq_rotation_weight(_lib.ptr_array([w.data_ptr() for w in weights]), d0, d1, taps,
                                                            int(bool(quaternion_format)), int(bool(transpose_out)),
                                                            out.data_ptr(), _stream()))
        ctx.geom = (d0, d1, taps, int(bool(quaternion_format)), int(bool(transpose_out)))
        ctx.save_for_backward(*weights)
        return out

    @staticmethod
    def backward(ctx, gout):
        weights = ctx.saved_tensors
        d0, d1, taps, qf, tr = ctx.geom
        gout = gout.contiguous()
        if gout.dtype != torch.float32:
            raise TypeError("seldq: grad_output must be float32")
        gws = [torch.empty_like(w) for w in weights]
        if gws[0].numel():
            with torch.cuda.device(gout.device):
                _lib.check(_lib.lib().seldq_rotation_weight_bwd(_lib.ptr_array([w.data_ptr() for w in weights]),
                                                                gout.data_ptr(), d0, d1, taps, qf, tr,
                                                                _lib.ptr_array([g.data_ptr() for g in gws]), _stream()))
        return (None, None) + tuple(gws)


def rotation_weight(weights, quaternion_format=False, transpose_out=False):
    return _RotationWeight.apply(bool(quaternion_format), bool(transpose_out), *weights)


def _real_prec(prec, stride, channels):
    """Precision of a real-algebra contraction: bf16 only where all three tensor-core passes serve the layer -- stride 1,
    8 ... 256 channels on both sides (forward / input gradient: dense mode, csrc/conv_cl_plan.h) and channel counts that
    are both multiples of 8 or both at most 64 (weight gradient: csrc/wgrad_umma.cu) -- else the fp32 kernels."""
    prec = _PRECISION if prec is None else prec
    if prec == PREC_BF16:
        lo, hi = min(channels), max(channels)
        ok = (_ROTATION_BF16 and _pair(stride) == (1, 1) and 8 <= lo and hi <= 256
              and (all(c % 8 == 0 for c in channels) or hi <= 64))
        if not ok:
            prec = PREC_FP32
    return prec


def quaternion_conv_rotation(x, weights, bias, stride, padding, groups, dilation, quaternion_format, prec=None):
    """quaternion_ops.py:174-232: a real convolution with the rotation weight of (r, i, j, k)."""
    if groups != 1:
        raise NotImplementedError("seldq: groups != 1 is not implemented")
    w = rotation_weight(weights, quaternion_format)
    return block_conv(x, (w,), bias, stride, padding, dilation, ALG_REAL, _real_prec(prec, stride, w.shape[:2]))


def quaternion_transpose_conv_rotation(x, weights, bias, stride, padding, output_padding, groups, dilation,
                                       quaternion_format, prec=None):
    """quaternion_ops.py:235-295: a real transposed convolution (functional.block_conv_transpose) with the rotation weight
    of the (in, out, k...) tensors."""
    w = rotation_weight(weights, quaternion_format)
    return block_conv_transpose(x, (w,), bias, stride, padding, output_padding, groups, dilation, ALG_REAL,
                                _real_prec(prec, stride, w.shape[:2]))


def quaternion_linear_rotation(x, weights, bias=None, quaternion_format=False, prec=None):
    """quaternion_ops.py:330-388: x @ R(r, i, j, k) (+ bias), R of shape (nc in, nc out).  Runs as a 1 x 1 real
    convolution over the flattened rows, whose (out, in, 1) weight the rotation kernel writes directly."""
    w = rotation_weight(tuple(t.unsqueeze(-1) for t in weights), quaternion_format, transpose_out=True)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    y = block_conv(x2.t().contiguous().unsqueeze(0), (w,), bias, 1, 0, 1, ALG_REAL, _real_prec(prec, 1, w.shape[:2]))
    return y[0].t().reshape(lead + (w.shape[0],))


class _QPointwise(torch.autograd.Function):
    """hamilton_product / q_normalize / quaternion_exp on (outer, 4, m) views (csrc/rotation.cu)."""

    @staticmethod
    def forward(ctx, op, a, b):
        _require_cuda_f32(a, "input")
        a = a.contiguous()
        if b is not None:
            _require_cuda_f32(b, "input")
            if b.shape != a.shape:
                raise RuntimeError("hamilton_product: operands must have the same shape, got %s and %s"
                                   % (list(a.shape), list(b.shape)))
            b = b.contiguous()
        if a.dim() < 2 or a.shape[1] % 4:
            raise RuntimeError("Quaternion Tensors must be divisible by 4. input.size()[1] = "
                               + str(a.shape[1] if a.dim() > 1 else a.numel()))
        outer = a.shape[0]
        m = a.numel() // (4 * outer) if outer else 0
        out = torch.empty_like(a)
        if a.numel():
            with torch.cuda.device(a.device):
                _lib.check(_lib.lib().seldq_quaternion_pointwise(op, a.data_ptr(), _ptr(b), out.data_ptr(), outer, m,
                                                                 _stream()))
        ctx.op, ctx.outer, ctx.m = op, outer, m
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous()
        L = _lib.lib()

        def run(op, p, q):
            out = torch.empty_like(a)
            if a.numel():
                with torch.cuda.device(a.device):
                    _lib.check(L.seldq_quaternion_pointwise(op, p.data_ptr(), q.data_ptr(), out.data_ptr(), ctx.outer, ctx.m,
                                                            _stream()))
            return out

        if ctx.op == _lib.QOP_HAMILTON:
            ga = run(_lib.QOP_HAMILTON_CONJ_B, g, b) if ctx.needs_input_grad[1] else None
            gb = run(_lib.QOP_HAMILTON_CONJ_A, a, g) if ctx.needs_input_grad[2] else None
            return None, ga, gb
        bwd = _lib.QOP_NORMALIZE_BWD if ctx.op == _lib.QOP_NORMALIZE else _lib.QOP_EXP_BWD
        return None, run(bwd, a, g), None


def hamilton_product(q0, q1):
    """quaternion_ops.py:467-507 / dual_quaternion_ops.py:374-414: component slices and concatenation along dim 1
    (2-d and >= 4-d inputs; the drop-in modules reject the 3-d inputs whose slices run along the last dimension)."""
    return _QPointwise.apply(_lib.QOP_HAMILTON, q0, q1)


def q_normalize(x):
    """dual_quaternion_ops.py:206-223 with channel = 1."""
    return _QPointwise.apply(_lib.QOP_NORMALIZE, x, None)


def quaternion_exp(x):
    """dual_quaternion_ops.py:227-246."""
    return _QPointwise.apply(_lib.QOP_EXP, x, None)


class _ActPool1d(torch.autograd.Function):
    """y = MaxPool1d(pool)(act(x)), act = ReLU | tanh: the activation + pooling pairs of the TC_Block tail
    (model.py:214-231) as one kernel per direction (csrc/tail.cu)."""

    @staticmethod
    def forward(ctx, x, act, pool):
        _require_cuda_f32(x, "input")
        x = x.contiguous()
        n, c, t = x.shape
        y = torch.empty((n, c, t // pool), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _timed("act_pool_kernel", 0.0, 1, lambda: _lib.check(_lib.lib().seldq_act_pool1d_fwd(
                x.data_ptr(), n * c, t, pool, act, y.data_ptr(), _stream())))
        ctx.act, ctx.pool = act, pool
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y = ctx.saved_tensors
        gy = gy.contiguous()
        n, c, t = x.shape
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _timed("act_pool_kernel", 0.0, 1, lambda: _lib.check(_lib.lib().seldq_act_pool1d_bwd(
                x.data_ptr(), y.data_ptr(), gy.data_ptr(), n * c, t, ctx.pool, ctx.act, gx.data_ptr(), _stream())))
        return gx, None, None


def act_pool1d(x, act_module, pool_module):
    """act_module (nn.ReLU | nn.Tanh) followed by pool_module (nn.MaxPool1d | None) on the fused kernel where it
    applies -- CUDA float32 (N, C, T), pooling with stride = kernel, no padding / dilation, floor mode -- and through
    the modules themselves otherwise.  Same arithmetic in both precision modes."""
    import torch.nn as nn
    act = _lib.ACT_RELU if isinstance(act_module, nn.ReLU) else _lib.ACT_TANH if isinstance(act_module, nn.Tanh) else None
    ok = (act is not None and isinstance(pool_module, nn.MaxPool1d) and isinstance(x, torch.Tensor) and x.is_cuda
          and x.dtype == torch.float32 and x.dim() == 3 and os.environ.get("SELDQ_TAIL", "1") != "0")
    if ok:
        k, s = pool_module.kernel_size, pool_module.stride
        k = k[0] if isinstance(k, (tuple, list)) else k
        s = s[0] if isinstance(s, (tuple, list)) else s
        ok = (isinstance(k, int) and s == k and pool_module.padding in (0, (0,)) and pool_module.dilation in (1, (1,))
              and not pool_module.ceil_mode and not pool_module.return_indices and 1 <= k <= x.shape[2])
    if ok:
        return _ActPool1d.apply(x, act, k)
    y = act_module(x)
    return y if pool_module is None else pool_module(y)


class _Attention(torch.autograd.Function):
    """out = softmax(q k^T / sqrt(d)) v per (sample, head) (model.py:40-48) on the fused tcgen05 kernels
    (csrc/attention.cu): q, k, v (N, E, S) float32 -- the 1x1 projections' output layout -- -> out (N, S, E)."""

    @staticmethod
    def forward(ctx, q, k, v, heads):
        import ctypes
        L = _lib.lib()
        for t, name in ((q, "q"), (k, "k"), (v, "v")):
            _require_cuda_f32(t, name)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        n, e, s = q.shape
        desc = _lib.AttentionDesc(n, heads, s, e // heads)
        out = torch.empty((n, s, e), dtype=torch.float32, device=q.device)
        lse = torch.empty((n * heads * s,), dtype=torch.float32, device=q.device)
        saved = torch.empty(L.seldq_attention_saved_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=q.device)
        with torch.cuda.device(q.device):
            _timed("attn_kernel", 4.0 * n * heads * s * s * (e // heads), 4, lambda: _lib.check(
                L.seldq_attention_fwd(ctypes.byref(desc), q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                      lse.data_ptr(), saved.data_ptr(), _stream())))
        ctx.desc = desc
        ctx.save_for_backward(saved, out, lse)
        return out

    @staticmethod
    def backward(ctx, gout):
        import ctypes
        L = _lib.lib()
        saved, out, lse = ctx.saved_tensors
        desc = ctx.desc
        gout = gout.contiguous()
        n, s, e = out.shape
        dq, dk, dv = (torch.empty((n, e, s), dtype=torch.float32, device=out.device) for _ in range(3))
        work = torch.empty(L.seldq_attention_bwd_workspace_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=out.device)
        with torch.cuda.device(out.device):
            _timed("attn_kernel", 10.0 * n * desc.heads * s * s * desc.head_dim, 3, lambda: _lib.check(
                L.seldq_attention_bwd(ctypes.byref(desc), saved.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                      gout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), work.data_ptr(),
                                      work.numel(), _stream())))
        return dq, dk, dv, None


def attention_supported(q, heads):
    """The fused attention kernels serve CUDA float32 (N, E, S) projections with E / heads in {16, 32, 48} and S a
    multiple of 8, in the tensor-core ('bf16') precision mode."""
    import ctypes
    if not (isinstance(q, torch.Tensor) and q.is_cuda and q.dtype == torch.float32 and q.dim() == 3):
        return False
    if get_precision() != "bf16" or os.environ.get("SELDQ_ATTN", "own") != "own" or q.shape[1] % heads:
        return False
    desc = _lib.AttentionDesc(q.shape[0], heads, q.shape[2], q.shape[1] // heads)
    return bool(_lib.lib().seldq_attention_supported(ctypes.byref(desc)))


def attention(q, k, v, heads):
    return _Attention.apply(q, k, v, heads)


def stft_magphase(x, nperseg=512, noverlap=128, cut_dc=True, output_phase=True, cut_last_timeframe=True):
    """x: (C, n) or (B, C, n) float32 CUDA tensor -> ((1+phase)*C, F, T) or (B, (1+phase)*C, F, T)
    (utility_functions.py:129-155)."""
    import ctypes
    L = _lib.lib()
    _require_cuda_f32(x, "signal")
    if x.dim() not in (2, 3):
        raise ValueError("stft_magphase expects (channels, samples) or (batch, channels, samples)")
    batched = x.dim() == 3
    xb = (x if batched else x[None]).contiguous()
    B, C, n = xb.shape
    nb, nf = ctypes.c_int32(), ctypes.c_int32()
    _lib.check(L.seldq_stft_shape(n, nperseg, noverlap, int(cut_dc), int(cut_last_timeframe), ctypes.byref(nb),
                                  ctypes.byref(nf)))
    planes = 2 if output_phase else 1
    out = torch.empty((B, planes * C, nb.value, nf.value), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _timed("stft_magphase_kernel", 0.0, 1, lambda: _lib.check(
            L.seldq_stft_magphase(xb.data_ptr(), B, C, n, nperseg, noverlap, int(cut_dc), int(output_phase),
                                  int(cut_last_timeframe), out.data_ptr(), _stream())))
    return out if batched else out[0]


def stft_features(x, nperseg=512, noverlap=128, cut_dc=True, output_phase=True, cut_last_timeframe=True, mean_std=None,
                  stats_only=False):
    """The front end with train.py's surrounding steps fused in (include/seldq.h, seldq_stft_features).
    x: (B, C, n) CUDA tensor, float32 or int16 PCM.  mean_std: ((mean_mag, std_mag), (mean_phase, std_phase)) or
    ((mean_mag, std_mag),): the data-set normalisation of train.py:374-408 applied on the way out.
    stats_only=True: nothing is stored; returns a (2, 2) float64 tensor [plane][sum, sum of squares] of the features
    (feature_mean_std turns it into the mean / population std train.py computes with np.mean / np.std)."""
    import ctypes
    L = _lib.lib()
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 3 and x.dtype in (torch.float32, torch.int16)):
        raise TypeError("stft_features expects a (batch, channels, samples) CUDA tensor of float32 or int16")
    x = x.contiguous()
    B, C, n = x.shape
    nb, nf = ctypes.c_int32(), ctypes.c_int32()
    _lib.check(L.seldq_stft_shape(n, nperseg, noverlap, int(cut_dc), int(cut_last_timeframe), ctypes.byref(nb),
                                  ctypes.byref(nf)))
    planes = 2 if output_phase else 1
    opt = _lib.StftOptions()
    opt.input_int16 = 1 if x.dtype == torch.int16 else 0
    for k in range(2):
        opt.mean[k], opt.inv_std[k] = 0.0, 1.0
    if mean_std is not None:
        for k, (mu, sd) in enumerate(mean_std):
            opt.mean[k], opt.inv_std[k] = float(mu), 1.0 / float(sd)
    stats = out = None
    if stats_only:
        stats = torch.zeros((2, 2), dtype=torch.float64, device=x.device)
        opt.stats = stats.data_ptr()
    else:
        out = torch.empty((B, planes * C, nb.value, nf.value), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _timed("stft_magphase_kernel", 0.0, 1, lambda: _lib.check(
            L.seldq_stft_features(x.data_ptr(), B, C, n, nperseg, noverlap, int(cut_dc), int(output_phase),
                                  int(cut_last_timeframe), ctypes.byref(opt), _ptr(out), _stream())))
    return stats if stats_only else out


def feature_mean_std(stats, count_per_plane):
    """(sum, sum of squares) per plane -> ((mean, std), ...) with the population std (np.std default, train.py:379-380)."""
    out = []
    for s1, s2 in stats.cpu().tolist():
        mu = s1 / count_per_plane
        out.append((mu, max(s2 / count_per_plane - mu * mu, 0.0) ** 0.5))
    return tuple(out)
