"""Batch-sharded data-parallel training step: one process per GPU, weights replicated, each rank
takes its own samples, one exchange per step.

Reproduces the step body of the reference's train.py:546-561
    zero_grad -> seld_loss (train.py:186-204: BCE(sed) + 5 * MSE(doa)) -> backward -> Adam.step
and adds what the reference does not have: all gradients live in ONE flat fp32
bucket (every p.grad is a view into it, so the wgrad kernels' results are accumulated straight
into the bucket by autograd), and the bucket is summed across ranks with a single NCCL all-reduce
over NVLink and divided by the world size.  BatchNorm stays per replica, as N independent
batches would in the reference.  Parameters that never receive a gradient (the unused
batch_gate1 of every ResBlock and the last block's conv2_residual) simply keep
zeros in their slice.
"""
import os

import torch
import torch.distributed as dist
import torch.nn.functional as tF

# SELDQ_AR_SUM=1: exchange with ncclSum and scale by 1 / world afterwards instead of ncclAvg.  NCCL implements the average
# as a pre-multiplied sum, which its in-switch reduction (NVLS) does not take for float32 (ncclNvlsSupported: ncclDevSum
# only); the plain sum is eligible.  Not yet measured at 8 GPUs -- the default stays the measured schedule.
_AR_SUM = os.environ.get("SELDQ_AR_SUM", "0") != "0"


def _nccl_mean_(t, group):
    if _AR_SUM:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / dist.get_world_size(group))
    else:
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)      # ncclAvg: no separate division pass


def seld_loss(sed, doa, target, n_sed, sed_weight=1.0, doa_weight=5.0):
    """train.py:186-204; target is (B, frames, n_sed + 3*n_sed) with SED first (train.py:191-192)."""
    def sed_term():
        return tF.binary_cross_entropy(torch.flatten(sed, start_dim=1), torch.flatten(target[:, :, :n_sed], start_dim=1)) * sed_weight

    def doa_term():
        return tF.mse_loss(torch.flatten(doa, start_dim=1), torch.flatten(target[:, :, n_sed:], start_dim=1)) * doa_weight

    if sed.is_cuda:
        # the two terms are independent chains of small kernels (forward and backward): a forked stream pair
        from .seld_model import _forked
        loss_sed, loss_doa = _forked(sed_term, doa_term, sed)
    else:
        loss_sed, loss_doa = sed_term(), doa_term()
    return loss_sed + loss_doa


class FlatGradBucket(object):
    """One contiguous buffer (the parameters' dtype: fp32 in training) holding every parameter's gradient, and a
    second one holding the parameters themselves (every p.data / p.grad is a view), so that the all-reduce and the
    optimiser each touch ONE tensor instead of a few hundred."""

    def __init__(self, params, late=None):
        """late(p) -> True for parameters whose gradient completes LAST in the backward pass (the CNN front): they are
        laid out first, so that flat[n_late:] -- everything whose gradient is complete once the TCN backward has
        finished -- is one contiguous slice that can be all-reduced while the CNN backward still runs."""
        params = [p for p in params if p.requires_grad]
        self.model_order = list(params)                       # model.parameters() order (checkpoint layout)
        if late is not None:
            params = [p for p in params if late(p)] + [p for p in params if not late(p)]
        self.params = params
        self.n_late = sum(p.numel() for p in params if late is not None and late(p))
        # 16-byte alignment of the boundary between the late and the early slice (the optimiser updates them separately)
        self.pad = (-self.n_late) % 4 if self.n_late else 0
        total = sum(p.numel() for p in self.params) + self.pad
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=self.params[0].dtype, device=dev)
        self.flat_param = torch.nn.Parameter(torch.empty(total, dtype=self.params[0].dtype, device=dev))
        off = 0
        self.offsets = {}
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat_param.data[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param.data[off:off + n].view_as(p)
                p.grad = self.flat[off:off + n].view_as(p)
                self.offsets[id(p)] = (off, n)
                off += n
                if off == self.n_late and self.pad:
                    off += self.pad                         # padding elements: zero gradient, never read back
        self.flat_param.grad = self.flat
        if dev.type == "cuda":
            # p.data was re-pointed: packed bf16 tiles keyed on the old addresses are stale
            from . import functional
            functional.invalidate_packed_weights()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.flat.is_cuda:
                _nccl_mean_(self.flat, group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))

    # ---- the exchange in two parts: flat[n_late:] as soon as the TCN backward is done, flat[:n_late] at the end ----
    def all_reduce_early(self, group, stream):
        """Sum flat[n_late:] across ranks on `stream` (forked behind the current stream; None on a CPU run, where
        the call is simply made at this point of the backward pass)."""
        if stream is None:
            dist.all_reduce(self.flat[self.n_late:], op=dist.ReduceOp.SUM, group=group)
            return
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            _nccl_mean_(self.flat[self.n_late:], group)

    def all_reduce_rest_mean(self, group, stream):
        """Sum flat[:n_late], join `stream`, divide everything by the world size."""
        if stream is not None:
            if self.n_late:
                _nccl_mean_(self.flat[:self.n_late], group)
            torch.cuda.current_stream().wait_stream(stream)
            return
        if self.n_late:
            dist.all_reduce(self.flat[:self.n_late], op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(dist.get_world_size(group))


class FlatAdam(object):
    """torch.optim.Adam(lr, betas, eps) -- the reference's optimiser, train.py:502-504 -- over the trainer's one flat
    CUDA parameter: the update is ONE kernel of this library (seldq_adam_step, csrc/tail.cu) with the step counter on
    the device, so it can be captured in the step's CUDA graph.  `state` / `param_groups` keep torch.optim's layout
    (Trainer.optimizer_state_dict splits it per parameter for the reference's checkpoints)."""

    def __init__(self, flat_param, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.flat_param = flat_param
        self.param_groups = [dict(params=[flat_param], lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False,
                                  capturable=True)]
        self.state = {}

    def _ensure_state(self):
        st = self.state.setdefault(self.flat_param, {})
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(self.flat_param.data)
            st["exp_avg_sq"] = torch.zeros_like(self.flat_param.data)
        if not torch.is_tensor(st.get("step")) or st["step"].device != self.flat_param.device:
            st["step"] = torch.as_tensor(float(st.get("step", 0.0)), dtype=torch.float32).to(self.flat_param.device).reshape(())
        return st

    def step(self):
        self.step_part(0, self.flat_param.numel(), True)

    def step_part(self, lo, hi, advance):
        """The update of elements [lo, hi) on the current stream.  All parts of one step see the same step count; the
        last one (advance=True) increments it.  lo must be a multiple of 4 (16-byte alignment)."""
        from . import _lib
        st = self._ensure_state()
        g = self.param_groups[0]
        fp = self.flat_param
        if hi <= lo:
            if not advance:
                return
            hi = lo                                           # nothing to update, but the counter must move
        assert lo % 4 == 0
        with torch.cuda.device(fp.device):
            _lib.check(_lib.lib().seldq_adam_step_part(
                fp.data.data_ptr() + 4 * lo, fp.grad.data_ptr() + 4 * lo, st["exp_avg"].data_ptr() + 4 * lo,
                st["exp_avg_sq"].data_ptr() + 4 * lo, max(hi - lo, 0), float(g["lr"]), float(g["betas"][0]),
                float(g["betas"][1]), float(g["eps"]), st["step"].data_ptr(), 1 if advance else 0,
                torch.cuda.current_stream().cuda_stream))


class Trainer(object):
    def __init__(self, model, lr=1e-4, n_sed=42, group=None, overlap_all_reduce=True):
        self.model = model
        self.n_sed = n_sed
        self.group = group
        # gradients of the CNN front complete last (backward runs heads -> TCN -> CNN): they go first in the bucket
        late_ids = {id(p) for n, p in model.named_parameters() if ".cnn." in n or n.startswith("cnn.")}
        self.bucket = FlatGradBucket(model.parameters(), late=lambda p: id(p) in late_ids)
        on_cuda = self.bucket.flat.is_cuda
        self._functional = None
        self._prev_accumulate = None
        if on_cuda:
            # the weight-gradient kernels add straight into the bucket (functional.set_grad_accumulation)
            from . import functional
            self._prev_accumulate = functional.set_grad_accumulation(True)
            self._functional = functional
        # the exchange of everything but the CNN front's gradients (flat[n_late:], ~80 % of the bucket)
        # is launched on a side stream the moment the TCN backward has finished and overlaps the CNN backward: a
        # gradient hook on the input of every TC_Block (the mirror's or the reference's own) marks that moment
        # SELDQ_AR_OVERLAP=0: one all-reduce of the whole bucket behind the backward pass (the round-1 schedule)
        self._overlap = bool(overlap_all_reduce) and __import__("os").environ.get("SELDQ_AR_OVERLAP", "1") != "0"
        self._on_cuda = on_cuda
        self._ar_stream = None
        self._tc_blocks = [m for m in model.modules() if hasattr(m, "ResBlocks") and hasattr(m, "attention")]
        self._pending = 0
        self._early_launched = False
        self._early_adam_done = False
        self._early_adam_ok = False                   # set below, once the optimiser is known
        if self._overlap:
            for m in self._tc_blocks:
                m.register_forward_pre_hook(self._tcn_pre_hook)
        # Adam with the reference's hyper-parameters (train.py:502-504); fused = one kernel per step,
        # capturable so that the whole step can live in a CUDA graph
        # (element-wise update: one flat parameter is the same arithmetic as one tensor per parameter)
        if on_cuda:
            self.optimizer = FlatAdam(self.bucket.flat_param, lr=lr)
            self.optimizer._ensure_state()                 # allocated outside any graph capture
            # SELDQ_EARLY_ADAM=0: one update of the whole bucket behind the backward pass
            self._early_adam_ok = (self._overlap and bool(self._tc_blocks) and self.bucket.n_late > 0
                                   and __import__("os").environ.get("SELDQ_EARLY_ADAM", "1") != "0")
        else:
            self.optimizer = torch.optim.Adam([self.bucket.flat_param], lr=lr)
        self._graph = None

    def close(self):
        """Restores the process-wide gradient-accumulation switch this trainer flipped."""
        if self._functional is not None and self._prev_accumulate is not None:
            self._functional.set_grad_accumulation(self._prev_accumulate)
            self._prev_accumulate = None

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _tcn_pre_hook(self, module, inputs):
        x = inputs[0]
        if self._overlap and (self._distributed() or self._early_adam_ok) and torch.is_grad_enabled() and x.requires_grad:
            self._pending += 1
            x.register_hook(self._tcn_backward_done)

    def _tcn_backward_done(self, grad):
        self._pending -= 1
        if self._pending == 0 and not self._early_launched:
            if self._ar_stream is None and self._on_cuda:
                self._ar_stream = torch.cuda.Stream()
            if self._distributed():
                self.bucket.all_reduce_early(self.group, self._ar_stream)
                self._early_launched = True
            elif self._ar_stream is not None:
                self._ar_stream.wait_stream(torch.cuda.current_stream())
            if self._early_adam_ok and self._ar_stream is not None:
                # everything but the CNN front has its final gradient: its Adam update runs on the side stream, behind
                # the early all-reduce, while the CNN backward still runs (the step counter moves with the last part)
                with torch.cuda.stream(self._ar_stream):
                    self.optimizer.step_part(self.bucket.n_late + self.bucket.pad, self.bucket.flat.numel(), False)
                self._early_adam_done = True
        return None

    def reduce_gradients(self):
        """The step's one exchange: mean of the flat bucket over the ranks (second half of it when the first was
        launched from the backward pass)."""
        if not self._distributed():
            return
        if self._early_launched:
            self.bucket.all_reduce_rest_mean(self.group, self._ar_stream)
        else:
            self.bucket.all_reduce_mean(self.group)

    # ---- optimiser state in the reference's checkpoint layout (utility_functions.save_model / load_model) ----------
    def optimizer_state_dict(self):
        """Adam's state as `torch.optim.Adam(model.parameters())` would hold it: one {step, exp_avg, exp_avg_sq} entry
        per parameter in model.parameters() order (the flat optimiser here keeps ONE entry for the whole bucket)."""
        st = self.optimizer.state.get(self.bucket.flat_param, {})
        group = {k: v for k, v in self.optimizer.param_groups[0].items() if k != "params"}
        out = {"state": {}, "param_groups": [dict(group, params=list(range(len(self.bucket.model_order))))]}
        if st:
            for i, p in enumerate(self.bucket.model_order):
                off, n = self.bucket.offsets[id(p)]
                out["state"][i] = {"step": st["step"].detach().clone().cpu() if torch.is_tensor(st["step"]) else st["step"],
                                   "exp_avg": st["exp_avg"][off:off + n].view_as(p).clone(),
                                   "exp_avg_sq": st["exp_avg_sq"][off:off + n].view_as(p).clone()}
        return out

    def load_optimizer_state_dict(self, sd):
        """Inverse of optimizer_state_dict: accepts the reference's per-parameter 'optimizer_state_dict'."""
        fp = self.bucket.flat_param
        if not sd.get("state"):
            return
        st = self.optimizer.state.setdefault(fp, {}) if isinstance(self.optimizer, FlatAdam) else self.optimizer.state[fp]
        step = None
        if "exp_avg" not in st:
            st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(fp.data), torch.zeros_like(fp.data)
        for i, p in enumerate(self.bucket.model_order):
            e = sd["state"].get(i, sd["state"].get(str(i)))
            if e is None:
                continue
            off, n = self.bucket.offsets[id(p)]
            st["exp_avg"][off:off + n].copy_(e["exp_avg"].reshape(-1))
            st["exp_avg_sq"][off:off + n].copy_(e["exp_avg_sq"].reshape(-1))
            step = e["step"]
        if step is not None:
            step = torch.as_tensor(step, dtype=torch.float32)
            st["step"] = step.to(fp.device) if self.optimizer.param_groups[0].get("capturable") else step.cpu()
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.optimizer.param_groups[0][k] = v

    def broadcast_parameters(self, src=0):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            # the dropout seeds of the fused paths (seld_model._drop_seed) stay per replica: every rank draws its own
            # from its torch RNG, so the Dropout / Dropout1d masks differ across the global batch
            bufs = [b for n, b in self.model.named_buffers() if not n.endswith("_drop_seed")]
            for t in list(self.model.parameters()) + bufs:
                dist.broadcast(t.data, src=src, group=self.group)
            if self._functional is not None:
                # writes through .data do not move the version counters the packed-weight cache watches
                self._functional.invalidate_packed_weights()

    def capture(self, x, target, warmup=3):
        """Captures zero_grad -> forward -> loss -> backward -> all-reduce -> Adam into ONE CUDA graph
        (static shapes: the step at per-GPU batch 1 is ~2000 launches of a few microseconds each, i.e.
        launch-bound when driven from Python).  `warmup` eager steps run first on a side stream, as
        graph capture requires; they are real optimisation steps."""
        self._sx, self._st = x.clone(), target.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self._sx, self._st)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._sloss = self.step(self._sx, self._st)
        self._graph = graph

    def step_graph(self, x, target):
        """Replays the captured step on new data (copied into the graph's static input buffers)."""
        self._sx.copy_(x, non_blocking=True)
        self._st.copy_(target, non_blocking=True)
        self._graph.replay()
        return self._sloss

    # ---- input pipeline of the captured step: the host -> device copy of batch i+1 overlaps step i ---------
    def stage(self, x_host, target_host):
        """Starts the asynchronous copy of a (pinned) host batch into the device staging buffers on a copy stream."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._gx, self._gt = torch.empty_like(self._sx), torch.empty_like(self._st)
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
        self._copy_stream.wait_event(self._consumed)          # the previous staged batch has been taken over
        with torch.cuda.stream(self._copy_stream):
            self._gx.copy_(x_host, non_blocking=True)
            self._gt.copy_(target_host, non_blocking=True)
            self._staged.record(self._copy_stream)

    def step_graph_staged(self):
        """Replays the captured step on the batch `stage` copied last (device -> device into the graph's static
        input buffers, ordered after the copy stream's transfer)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self._sx.copy_(self._gx, non_blocking=True)
        self._st.copy_(self._gt, non_blocking=True)
        self._consumed.record(cur)
        self._graph.replay()
        return self._sloss

    def step(self, x, target):
        """One optimisation step on this rank's shard; returns the (local) loss tensor."""
        self.bucket.zero()
        self._pending, self._early_launched, self._early_adam_done = 0, False, False
        sed, doa = self.model(x)
        loss = seld_loss(sed, doa, target, self.n_sed)
        loss.backward()
        self.reduce_gradients()
        if self._early_adam_done:
            torch.cuda.current_stream().wait_stream(self._ar_stream)       # (a no-op after the distributed join)
            self.optimizer.step_part(0, self.bucket.n_late, True)
        else:
            self.optimizer.step()
        if self._functional is not None:
            # the flat update bypasses the parameters' version counters: refresh every packed bf16 weight set now,
            # in one launch, for the next step
            self._functional.repack_all()
        return loss
