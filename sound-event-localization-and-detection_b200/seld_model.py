"""DQSELD-TCN assembly: CNN (3 conv2d blocks) -> TCN (gated dilated residual blocks) ->
multi-head self-attention -> SED / DOA heads.

Host-side mirror of the reference's model.py (MultiHeadAttention :12-51, ResBlock :53-132,
TC_Block :134-232, ConvTC_Block :234-322, SELD_Model :324-480): same constructor arguments,
same sub-module names and construction order, hence the same state_dict keys and -- given the
same numpy / torch seeds -- bit-identical initial weights, so reference checkpoints load here and
vice versa.  The reference model.py itself also runs unchanged on top of the drop-in layer
modules; this mirror exists because the GPU box has no reference tree, and
because it is where the fused epilogue kernels are wired in.

`layer_lib` lets a caller substitute the Q / DQ layer classes (the CPU oracle does that to time
the reference's arithmetic on host cores); the default is the sm_100a implementation.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as tF

from . import functional as _F
from . import fused as _fused
from . import layers as _default_layers

# The reference runs its real-valued nn.Conv* / nn.Linear layers (MultiHeadAttention projections, the 'R' domain)
# in true fp32: it disables cuDNN altogether (model.py:10), so ATen's native kernels are used and no TF32 is
# involved.  cuDNN stays enabled here (it is far faster); TF32 follows the precision mode (functional.py): off in
# 'fp32' mode so that those layers keep the reference's arithmetic, on in the tensor-core 'bf16' mode.
_F._apply_tf32_policy()

import os as _os
# In 'bf16' mode MultiHeadAttention (model.py:12-51; real-valued, not a Q / DQ layer -- SURVEY 8f N1) may feed 16-bit
# operands to the library's fused attention kernel, so the S x S energy tensor never exists in HBM.
#   SELDQ_ATTN=fp16  IEEE half operands: 10-bit mantissa = the TF32 rounding the fp32 path applies anyway
#   SELDQ_ATTN=bf16  bf16 operands (7-bit mantissa: costs 0.2e-2 of the 2e-2 output tolerance)
#   SELDQ_ATTN=fp32  fp32 / TF32 scaled_dot_product_attention
#   SELDQ_ATTN=own   (default) this repository's fused tcgen05 attention kernels where they apply (functional.attention)
ATTN_DTYPE = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": None, "own": None}[
    "bf16" if _os.environ.get("SELDQ_ATTN_BF16", "0") == "1" else _os.environ.get("SELDQ_ATTN", "own")]
_BN_TCN = {'BN', 'BN_on_TCN', 'BNonTCN'}
_BN_CNN = {'BN', 'BN_on_CNN', 'BNonCNN'}
_TWO_BRANCH = {'2Parallel', '2BParallel', '2ParallelBranches', '2PB'}


def _make_conv(layer_lib, domain, ndim, cin, cout, k, padding, dilation, bias):
    op = 'convolution1d' if ndim == 1 else 'convolution2d'
    if domain == 'Q':
        return layer_lib.QuaternionConv(cin, cout, kernel_size=k, stride=1, padding=padding,
                                        dilatation=dilation, bias=bias, operation=op)
    if domain == 'DQ':
        return layer_lib.DualQuaternionConv(cin, cout, kernel_size=k, stride=1, padding=padding,
                                            dilatation=dilation, bias=bias, operation=op)
    cls = nn.Conv1d if ndim == 1 else nn.Conv2d
    return cls(cin, cout, kernel_size=k, stride=1, padding=padding, dilation=dilation, bias=bias)


class MultiHeadAttention(nn.Module):
    """model.py:12-51.  Real-valued; 1x1 conv projections without bias, softmax(QK^T / sqrt(d)) V,
    output Linear.  The S x S energy tensor of the reference is not materialised: the same
    expression goes through scaled_dot_product_attention."""

    def __init__(self, embed_size, num_heads):
        super().__init__()
        assert embed_size % num_heads == 0, "Embedding size must be divisible by number of heads"
        self.num_heads = num_heads
        self.head_dim = embed_size // num_heads
        self.values = nn.Conv1d(embed_size, embed_size, kernel_size=1, bias=False)
        self.keys = nn.Conv1d(embed_size, embed_size, kernel_size=1, bias=False)
        self.queries = nn.Conv1d(embed_size, embed_size, kernel_size=1, bias=False)
        self.fc_out = nn.Linear(embed_size, embed_size)

    def forward(self, v, k, q, mask=None):
        return mha_forward(self, v, k, q, mask)


def _split_heads(t, num_heads, head_dim):       # (N, E, L) -> (N, heads, L, head_dim); e = h*head_dim + d
    n, _, length = t.shape
    return t.view(n, num_heads, head_dim, length).transpose(2, 3)


def mha_forward(self, v, k, q, mask=None):
    """MultiHeadAttention.forward (model.py:25-51) for any module with the reference's attributes
    (values / keys / queries / fc_out, num_heads, head_dim).  Inputs are (N, L, E)."""
    n, length = q.shape[0], q.shape[1]
    convs = (self.values, self.keys, self.queries)
    if (v is k and k is q and n == 1 and q.is_cuda
            and all(isinstance(c, nn.Conv1d) and c.bias is None and c.kernel_size == (1,) and c.stride == (1,)
                    and c.padding == (0,) and c.dilation == (1,) and c.groups == 1 for c in convs)
            and self.values.weight.shape == self.keys.weight.shape == self.queries.weight.shape):
        # self-attention (model.py:218 passes the same tensor three times): the three 1x1 projections of model.py:33-35
        # as ONE convolution over the stacked weights -- one GEMM forward, one dgrad and one wgrad GEMM backward instead
        # of three each (the per-sample slices of the (1, 3E, L) result are contiguous)
        e = self.values.weight.shape[0]
        vkq = tF.conv1d(q.permute(0, 2, 1), torch.cat([c.weight for c in convs], 0))
        vc, kc, qc = vkq.split(e, dim=1)              # one node: its backward is a single concatenation
    else:
        vc, kc, qc = self.values(v.permute(0, 2, 1)), self.keys(k.permute(0, 2, 1)), self.queries(q.permute(0, 2, 1))
    if (mask is None and kc.shape == qc.shape == vc.shape and self.head_dim * self.num_heads == qc.shape[1]
            and _F.attention_supported(qc, self.num_heads)):
        # the repository's fused attention kernels (csrc/attention.cu): the (N, heads, S, S) energy / attention
        # tensors of model.py:40-46 never exist in HBM, forward or backward
        return self.fc_out(_F.attention(qc, kc, vc, self.num_heads))
    vp = _split_heads(vc, self.num_heads, self.head_dim)
    kp = _split_heads(kc, self.num_heads, self.head_dim)
    qp = _split_heads(qc, self.num_heads, self.head_dim)
    attn_mask = None if mask is None else (mask != 0)
    if ATTN_DTYPE is not None and q.is_cuda and _F.get_precision() == "bf16" and attn_mask is None:
        # tensor-core mode: 16-bit operands into the library's fused attention kernel (no S x S tensor in HBM)
        out = tF.scaled_dot_product_attention(qp.to(ATTN_DTYPE), kp.to(ATTN_DTYPE), vp.to(ATTN_DTYPE)).float()
    elif attn_mask is None and q.is_cuda:
        # the reference's three steps (model.py:40-47) as two batched GEMMs around one softmax; the scale is
        # folded into q.  (scaled_dot_product_attention's fp32 math backend adds a masking pass, an -inf scan
        # and a scaling pass over the S x S tensor: 0.25 ms of a 5 ms step.)
        attn = torch.softmax(torch.matmul(qp * (1.0 / self.head_dim ** 0.5), kp.transpose(-1, -2)), dim=-1)
        out = torch.matmul(attn, vp)
    else:
        out = tF.scaled_dot_product_attention(qp, kp, vp, attn_mask=attn_mask)    # scale = 1/sqrt(head_dim)
    out = out.transpose(1, 2).reshape(n, length, self.num_heads * self.head_dim)
    return self.fc_out(out)


class ResBlock(nn.Module):
    """Pre-activation gated residual block (model.py:53-132)."""

    def __init__(self, in_channels, domain='DQ', G=128, U=128, kernel_size_dilated_conv=3, dilation=1, stride=1,
                 spatial_dropout_rate=0.5, use_bias_conv=True, batch_norm='BN', verbose=False, layer_lib=None):
        super().__init__()
        lib = layer_lib or _default_layers
        self.verbose = verbose
        self.batch_norm = batch_norm
        self.spatial_dropout_rate = spatial_dropout_rate
        self.domain = domain
        padding = int(((kernel_size_dilated_conv - 1) * dilation) / 2)
        L = in_channels
        self.conv1_filter = _make_conv(lib, domain, 1, L, G, kernel_size_dilated_conv, padding, dilation, use_bias_conv)
        self.conv1_gate = _make_conv(lib, domain, 1, L, G, kernel_size_dilated_conv, padding, dilation, use_bias_conv)
        if batch_norm in _BN_TCN:
            self.batch_filter1 = nn.BatchNorm1d(L)
            self.batch_gate1 = nn.BatchNorm1d(L)      # allocated but unused, as in the reference
            self.batch_filter2 = nn.BatchNorm1d(G)
            self.batch_gate2 = nn.BatchNorm1d(G)
        self.tanh = nn.Tanh()
        self.sigmoid = nn.Sigmoid()
        if not spatial_dropout_rate == 0:
            self.dropout = nn.Dropout1d(p=spatial_dropout_rate)
        self.conv2_skip = _make_conv(lib, domain, 1, G, U, 1, 0, 1, use_bias_conv)
        self.conv2_residual = _make_conv(lib, domain, 1, G, L, 1, 0, 1, use_bias_conv)

    def forward(self, x):
        bn = self.batch_norm in _BN_TCN
        if bn:
            x = self.tanh(self.batch_filter1(x))
        y_f = self.conv1_filter(x)
        y_g = self.conv1_gate(x)
        if bn:
            y_f = self.batch_filter2(y_f)
            y_g = self.batch_gate2(y_g)
        y = self.tanh(y_f) * self.sigmoid(y_g)
        if not self.spatial_dropout_rate == 0:
            y = self.dropout(y)
        return x + self.conv2_residual(y), self.conv2_skip(y)


def _dilations(D, dilation_mode):
    """Per-stack dilation lists (model.py:150-176): fibonacci 1,1,2,3,5,... or powers of two."""
    out = []
    for n_resblock in D:
        if type(n_resblock) == list:
            out.extend(n_resblock)
            continue
        prev1, prev2 = 1, 0
        for d in range(n_resblock):
            if dilation_mode == 'fibonacci':
                if d == 0:
                    dil = 1
                else:
                    dil = prev1 + prev2
                    prev2, prev1 = prev1, dil
            else:
                dil = 2 ** d
            out.append(dil)
    return out


class TC_Block(nn.Module):
    """model.py:134-232."""

    def __init__(self, in_channels, domain='DQ', G=128, U=128, V=[128, 128], V_kernel_size=3,
                 pool_size=[[8, 2], [8, 2], [2, 2]], D=[10], spatial_dropout_rate=0.5, use_bias_conv=True,
                 dilation_mode='fibonacci', pool_time='TCN', batch_norm='BN', kernel_size_dilated_conv=3,
                 verbose=False, attention_type=None, key_size=None, value_size=None, layer_lib=None):
        super().__init__()
        lib = layer_lib or _default_layers
        self.verbose = verbose
        self.ResBlocks = nn.ModuleList()
        self.D = D
        self.pool_time = pool_time
        self.domain = domain
        for dil in _dilations(D, dilation_mode):
            self.ResBlocks.append(ResBlock(in_channels=in_channels, domain=domain, G=G, U=U,
                                           kernel_size_dilated_conv=kernel_size_dilated_conv, dilation=dil,
                                           spatial_dropout_rate=spatial_dropout_rate, use_bias_conv=use_bias_conv,
                                           batch_norm=batch_norm, verbose=verbose, layer_lib=lib))
        self.relu1 = nn.ReLU()
        if self.pool_time == 'TCN':
            self.maxpool1 = nn.MaxPool1d(pool_size[0][1])
        self.conv1 = _make_conv(lib, domain, 1, in_channels, V[0], V_kernel_size, 1, 1, use_bias_conv)
        self.attention = MultiHeadAttention(embed_size=V[0], num_heads=8)
        self.relu2 = nn.ReLU()
        if self.pool_time == 'TCN':
            self.maxpool2 = nn.MaxPool1d(pool_size[1][1])
        self.conv2 = _make_conv(lib, domain, 1, V[0], V[1], V_kernel_size, 1, 1, use_bias_conv)
        self.tanh = nn.Tanh()
        if self.pool_time == 'TCN':
            self.maxpool3 = nn.MaxPool1d(pool_size[2][1])
        # seed of the channel-dropout masks of the fused residual-block path (fused.next_drop_seed; not in the state_dict)
        self.register_buffer("_drop_seed", torch.zeros(1, dtype=torch.int64), persistent=False)

    def forward(self, residual):
        return tc_block_forward(self, residual)


def tc_block_forward(self, residual):
    """TC_Block.forward (model.py:204-232) for any module with the reference's attributes: the residual blocks run
    through the fused kernels where fused.tcn_stack_supported says so, layer by layer otherwise."""
    if _fused.tcn_stack_supported(self.ResBlocks, residual, self.training):
        sum_skip = _fused.tcn_stack(residual, self.ResBlocks, _fused.next_drop_seed(self, residual.device))
    else:
        sum_skip = None
        for block in self.ResBlocks:
            residual, skip = block(residual)
            sum_skip = skip if sum_skip is None else sum_skip + skip
    # activation + MaxPool1d pairs of the tail (model.py:214-231): one kernel per direction each (csrc/tail.cu)
    pooled = self.pool_time == 'TCN'
    out = _F.act_pool1d(sum_skip, self.relu1, self.maxpool1) if pooled else self.relu1(sum_skip)
    out = self.conv1(out)
    out = out.permute(0, 2, 1)
    out = self.attention(out, out, out, mask=None)
    out = out.permute(0, 2, 1)
    out = _F.act_pool1d(out, self.relu2, self.maxpool2) if pooled else self.relu2(out)
    out = self.conv2(out)
    out = _F.act_pool1d(out, self.tanh, self.maxpool3) if pooled else self.tanh(out)
    return out


class ConvTC_Block(nn.Module):
    """model.py:234-322."""

    def __init__(self, time_dim, freq_dim=256, input_channels=4, domain='DQ', cnn_filters=[64, 64, 64],
                 kernel_size_cnn_blocks=3, pool_size=[[8, 2], [8, 2], [2, 2]], pool_time='TCN', D=[10],
                 dilation_mode='fibonacci', G=128, U=128, kernel_size_dilated_conv=3, spatial_dropout_rate=0.5,
                 V=[128, 128], V_kernel_size=3, dropout_perc=0.3, use_bias_conv=True, batch_norm='noBN',
                 attention_type=None, key_size=None, value_size=None, verbose=False, layer_lib=None):
        super().__init__()
        lib = layer_lib or _default_layers
        self.time_dim = time_dim
        self.freq_dim = freq_dim
        self.domain = domain
        self.verbose = verbose
        self.D = D
        self.kernel_size_dilated_conv = kernel_size_dilated_conv
        self.dilation_mode = dilation_mode
        if pool_time == 'CNN':
            self.time_pooled_size = int(time_dim / np.prod(np.array(pool_size), axis=0)[-1])
        else:
            self.time_pooled_size = time_dim
        blocks = []
        in_chans = input_channels
        for p, c in zip(pool_size, np.array(cnn_filters)):
            pool = [p[0], p[1]] if pool_time == 'CNN' else [p[0], 1]
            mods = [_make_conv(lib, domain, 2, in_chans, c, kernel_size_cnn_blocks, 1, 1, use_bias_conv)]
            if batch_norm in _BN_CNN:
                mods.append(nn.BatchNorm2d(c))
            mods += [nn.ReLU(), nn.MaxPool2d(pool), nn.Dropout(dropout_perc)]
            blocks.append(nn.Sequential(*mods))
            in_chans = c
        self.cnn = nn.Sequential(*blocks)
        # seed of the dropout masks of the fused CNN path (fused.next_drop_seed; not part of the state_dict)
        self.register_buffer("_drop_seed", torch.zeros(1, dtype=torch.int64), persistent=False)
        L = int(freq_dim / np.prod(np.array(pool_size), axis=0)[0] * cnn_filters[-1])
        self.tcn = TC_Block(in_channels=L, domain=domain, G=G, U=U, V=V, V_kernel_size=V_kernel_size,
                            pool_size=pool_size, D=D, spatial_dropout_rate=spatial_dropout_rate,
                            use_bias_conv=use_bias_conv, dilation_mode=dilation_mode, pool_time=pool_time,
                            batch_norm=batch_norm, kernel_size_dilated_conv=kernel_size_dilated_conv,
                            verbose=verbose, attention_type=attention_type, key_size=key_size,
                            value_size=value_size, layer_lib=lib)

    def forward(self, x):
        return convtc_block_forward(self, x)

    def _cnn_forward(self, x):
        return _cnn_forward(self, x)


def _cnn_forward(self, x):
    """The CNN front: fused conv -> BN -> ReLU -> pool -> dropout kernels (fused.cnn_stack) in the training
    configuration they serve, the layer-by-layer modules otherwise."""
    convs, bns, pools, drops = [], [], [], []
    for blk in self.cnn:
        mods = list(blk)
        convs.append(mods[0])
        bns.append(next((m for m in mods if isinstance(m, nn.BatchNorm2d)), None))
        pools.append(next((m for m in mods if isinstance(m, nn.MaxPool2d)), None))
        drops.append(next((m for m in mods if isinstance(m, nn.Dropout)), None))
    if all(p is not None for p in pools) and _fused.cnn_stack_supported(convs, bns, pools, drops, x, self.training):
        return _fused.cnn_stack(x, convs, bns, pools, drops, _fused.next_drop_seed(self, x.device))
    return self.cnn(x)


def convtc_block_forward(self, x):
    """ConvTC_Block.forward (model.py:297-322) for any module with the reference's attributes."""
    x = _cnn_forward(self, x)                          # (B, C, F', T)
    # model.py:301-310 permutes to (B, T, C, F'), flattens (C, F') and permutes back: channel c*F' + f, time last,
    # which is this view of the contiguous CNN output (no copy)
    assert x.shape[3] == self.time_pooled_size
    x = x.reshape(x.shape[0], x.shape[1] * x.shape[2], x.shape[3])
    x = self.tcn(x)
    return x.permute(0, 2, 1)                         # (B, T/8, V)


_HEAD_FORK = _os.environ.get("SELDQ_HEAD_FORK", "1") != "0"
_FORK_SIDE = {}


def _forked(fn_main, fn_side, ref):
    """(fn_main(), fn_side()) with fn_side on a forked CUDA stream, joined before the results are used.  Autograd runs
    each node's backward on the stream of its forward, so the backward passes overlap the same way.  Tensors that
    cross the join are registered with the stream that consumes them.  Capturable into a CUDA graph."""
    if not (_HEAD_FORK and isinstance(ref, torch.Tensor) and ref.is_cuda):
        return fn_main(), fn_side()
    key = (ref.device.type, ref.device.index)
    side = _FORK_SIDE.get(key)
    if side is None:
        side = _FORK_SIDE[key] = torch.cuda.Stream(device=ref.device)
    cur = torch.cuda.current_stream(ref.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        out_side = fn_side()
    out_main = fn_main()
    cur.wait_stream(side)
    out_side.record_stream(cur)
    return out_main, out_side


def seld_heads(self, x):
    """(self.sed(x), self.doa(x)) of model.py:473-474.  The two heads are independent chains of small kernels (Q / DQ
    linear, dropout, nn.Linear, activation): the DOA head runs on a forked stream next to the SED head
    (SELDQ_HEAD_FORK=0: one after the other)."""
    sed, doa = _forked(lambda: self.sed(x), lambda: self.doa(x), x)
    if x.is_cuda and _HEAD_FORK:
        x.record_stream(_FORK_SIDE[(x.device.type, x.device.index)])
    return sed, doa


def seld_model_forward(self, x):
    """SELD_Model.forward (model.py:461-480) for any module with the reference's attributes: same tensors, same order
    of the concatenation; the two ConvTC branches of the parallel configurations and the two heads each run as a
    forked pair."""
    if self.parallel_ConvTC_block in _TWO_BRANCH:
        if self.parallel_magphase:
            x_a = torch.cat((x[:, :4], x[:, 8:12]), 1)      # mic A magnitude + phase
            x_b = torch.cat((x[:, 4:8], x[:, 12:]), 1)      # mic B magnitude + phase
        else:
            half = self.input_channels // 2
            x_a, x_b = x[:, :half].contiguous(), x[:, half:].contiguous()
        a, b = _forked(lambda: self.branch_A(x_a), lambda: self.branch_B(x_b), x)
        if x.is_cuda and _HEAD_FORK:
            x_b.record_stream(_FORK_SIDE[(x.device.type, x.device.index)])
        x = torch.cat((a, b), 2)
    else:
        x = self.seld_block(x)
    sed, doa = seld_heads(self, x)
    if getattr(self, "verbose", False):
        print('sed prediction:  ', sed.shape)
        print('doa prediction: ', doa.shape)
    return sed, doa


_PATCHED = set()


def patch_reference_model(mod):
    """`mod` = the reference's own model.py, imported UNMODIFIED on top of the drop-in layer modules (model.py:7-8).
    Rebinds the forward methods of its TC_Block / ConvTC_Block / MultiHeadAttention / SELD_Model classes to the functions above:
    the same module attributes, parameters and semantics (every branch falls back to the layer-by-layer modules where
    the fused kernels do not apply), but the glue between the convolutions now runs in this repository's kernels.
    The originals stay reachable as `<class>._reference_forward`.  Idempotent."""
    if id(mod) in _PATCHED:
        return False
    targets = (("TC_Block", tc_block_forward), ("ConvTC_Block", convtc_block_forward),
               ("MultiHeadAttention", mha_forward), ("SELD_Model", seld_model_forward))
    for cls_name, fn in targets:
        cls = getattr(mod, cls_name, None)
        if cls is None or not isinstance(cls, type) or cls.__module__ == __name__:
            return False
    for cls_name, fn in targets:
        cls = getattr(mod, cls_name)
        cls._reference_forward = cls.forward
        cls.forward = fn
    _PATCHED.add(id(mod))
    return True


class SELD_Model(nn.Module):
    """model.py:324-480."""

    def __init__(self, time_dim, freq_dim=256, input_channels=4, output_classes=14, domain='DQ',
                 domain_classifier='same', cnn_filters=[64, 64, 64], kernel_size_cnn_blocks=3,
                 pool_size=[[8, 2], [8, 2], [2, 2]], pool_time='TCN', D=[10], dilation_mode='fibonacci',
                 G=128, U=128, kernel_size_dilated_conv=3, spatial_dropout_rate=0.5, V=[128, 128],
                 V_kernel_size=3, fc_layers=[128], fc_activations='Linear', fc_dropout='all', dropout_perc=0.3,
                 class_overlaps=3., use_bias_conv=False, use_bias_linear=True, batch_norm='BN',
                 parallel_ConvTC_block='False', parallel_magphase=False, extra_name='', attention_type=None,
                 key_size=None, value_size=None, verbose=False, layer_lib=None):
        super().__init__()
        lib = layer_lib or _default_layers
        self.input_channels = input_channels
        self.time_dim = time_dim
        self.freq_dim = freq_dim
        self.domain = domain
        self.verbose = verbose
        self.D = D
        self.kernel_size_dilated_conv = kernel_size_dilated_conv
        self.dilation_mode = dilation_mode
        self.parallel_magphase = parallel_magphase
        self.domain_classifier = domain if domain_classifier == 'same' else domain_classifier
        self.receptive_field, self.total_n_resblocks = self.calculate_receptive_field()
        self.parallel_ConvTC_block = parallel_ConvTC_block
        self.model_name = self._name(domain, dilation_mode, D, parallel_ConvTC_block, batch_norm, pool_time,
                                     extra_name)
        sed_output_size = int(output_classes * class_overlaps)
        doa_output_size = sed_output_size * 3
        block_kw = dict(time_dim=time_dim, freq_dim=freq_dim, domain=domain, cnn_filters=cnn_filters,
                        kernel_size_cnn_blocks=kernel_size_cnn_blocks, pool_size=pool_size, pool_time=pool_time,
                        D=D, dilation_mode=dilation_mode, G=G, U=U,
                        kernel_size_dilated_conv=kernel_size_dilated_conv,
                        spatial_dropout_rate=spatial_dropout_rate, V=V, V_kernel_size=V_kernel_size,
                        dropout_perc=dropout_perc, use_bias_conv=use_bias_conv, batch_norm=batch_norm,
                        verbose=False, layer_lib=lib)
        if parallel_ConvTC_block in _TWO_BRANCH:
            self.branch_A = ConvTC_Block(input_channels=input_channels // 2, **block_kw)
            self.branch_B = ConvTC_Block(input_channels=input_channels // 2, **block_kw)
            fc_input_size = V[-1] * 2
        else:
            self.seld_block = ConvTC_Block(input_channels=input_channels, attention_type=attention_type,
                                           key_size=key_size, value_size=value_size, **block_kw)
            fc_input_size = V[-1]

        def fc(n_in, n_out):
            if self.domain_classifier == 'Q':
                return lib.QuaternionLinear(n_in, n_out, bias=use_bias_linear)
            if self.domain_classifier == 'DQ':
                return lib.DualQuaternionLinear(n_in, n_out, bias=use_bias_linear)
            return nn.Linear(n_in, n_out, bias=use_bias_linear)

        sed_list, doa_list = [], []
        for width in fc_layers:
            sed_list.append(fc(fc_input_size, width))       # sed before doa: keeps the RNG stream of the reference
            doa_list.append(fc(fc_input_size, width))
            if fc_activations in {'relu', 'ReLU', 'RELU'}:
                sed_list.append(nn.ReLU())
                doa_list.append(nn.ReLU())
            if fc_dropout in {'all', 'ALL', 'True'}:
                sed_list.append(nn.Dropout(dropout_perc))
                doa_list.append(nn.Dropout(dropout_perc))
            fc_input_size = width
        if fc_dropout in {'last', 'Last', 'LAST'}:
            sed_list.append(nn.Dropout(dropout_perc))
            doa_list.append(nn.Dropout(dropout_perc))
        self.sed = nn.Sequential(*sed_list, nn.Linear(fc_layers[-1], sed_output_size, bias=use_bias_linear),
                                 nn.Sigmoid())
        self.doa = nn.Sequential(*doa_list, nn.Linear(fc_layers[-1], doa_output_size, bias=use_bias_linear),
                                 nn.Tanh())

    def _name(self, domain, dilation_mode, D, parallel, batch_norm, pool_time, extra_name):
        # model.py:347-372 -- the name doubles as the output directory of train.py:461-468
        if domain in {'q', 'Q', 'quaternion', 'Quaternion'}:
            name = 'Q'
        elif domain in {'dq', 'dQ', 'DQ', 'dual_quaternion', 'Dual_Quaternion'}:
            name = 'DualQ'
        else:
            name = ''
        name += 'SELD-TCN'
        if dilation_mode == 'fibonacci':
            name += '-PHI'
        name += '-'
        if len(D) > 1 and D[0] < D[1]:
            name += 'I'
        name += 'S' + str(len(D))
        if parallel not in {'False', 'false', 'None', 'none'}:
            name += '_' + parallel
        name += '_' + batch_norm
        if pool_time == 'CNN':
            name += '_pooltCNN'
        name += '_RF{}_{}RB'.format(self.receptive_field, self.total_n_resblocks)
        return name + extra_name

    def forward(self, x):
        return seld_model_forward(self, x)

    def calculate_receptive_field(self, verbose=0):
        dils = _dilations(self.D, self.dilation_mode)
        rf = 1 + sum((self.kernel_size_dilated_conv - 1) * d for d in dils)
        return rf, len(dils)
