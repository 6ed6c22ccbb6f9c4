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
import torch
import torch.distributed as dist
import torch.nn.functional as tF


def seld_loss(sed, doa, target, n_sed, sed_weight=1.0, doa_weight=5.0):
    """train.py:186-204; target is (B, frames, n_sed + 3*n_sed) with SED first (train.py:191-192)."""
    t_sed = torch.flatten(target[:, :, :n_sed], start_dim=1)
    t_doa = torch.flatten(target[:, :, n_sed:], start_dim=1)
    loss_sed = tF.binary_cross_entropy(torch.flatten(sed, start_dim=1), t_sed) * sed_weight
    loss_doa = tF.mse_loss(torch.flatten(doa, start_dim=1), t_doa) * doa_weight
    return loss_sed + loss_doa


class FlatGradBucket(object):
    """One contiguous buffer (the parameters' dtype: fp32 in training) holding every parameter's gradient, and a
    second one holding the parameters themselves (every p.data / p.grad is a view), so that the all-reduce and the
    optimiser each touch ONE tensor instead of a few hundred."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=self.params[0].dtype, device=dev)
        self.flat_param = torch.nn.Parameter(torch.empty(total, dtype=self.params[0].dtype, device=dev))
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat_param.data[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param.data[off:off + n].view_as(p)
                p.grad = self.flat[off:off + n].view_as(p)
                off += n
        self.flat_param.grad = self.flat
        if dev.type == "cuda":
            # p.data was re-pointed: packed bf16 tiles keyed on the old addresses are stale
            from . import functional
            functional.invalidate_packed_weights()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


class Trainer(object):
    def __init__(self, model, lr=1e-4, n_sed=42, group=None):
        self.model = model
        self.n_sed = n_sed
        self.group = group
        self.bucket = FlatGradBucket(model.parameters())
        on_cuda = self.bucket.flat.is_cuda
        self._functional = None
        if on_cuda:
            # the weight-gradient kernels add straight into the bucket (functional.set_grad_accumulation)
            from . import functional
            functional.set_grad_accumulation(True)
            self._functional = functional
        # Adam with the reference's hyper-parameters (train.py:502-504); fused = one kernel per step,
        # capturable so that the whole step can live in a CUDA graph
        # (element-wise update: one flat parameter is the same arithmetic as one tensor per parameter)
        self.optimizer = torch.optim.Adam([self.bucket.flat_param], lr=lr, fused=on_cuda, capturable=on_cuda)
        self._graph = None

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
        sed, doa = self.model(x)
        loss = seld_loss(sed, doa, target, self.n_sed)
        loss.backward()
        self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()
        if self._functional is not None:
            # the flat update bypasses the parameters' version counters: refresh every packed bf16 weight set now,
            # in one launch, for the next step
            self._functional.repack_all()
        return loss
