import torch, time
from torch.nn.attention import sdpa_kernel, SDPBackend
import torch.nn.functional as F
torch.backends.cuda.matmul.allow_tf32 = True
q = torch.randn(1, 8, 2400, 48, device="cuda", requires_grad=True)
k = torch.randn_like(q, requires_grad=True); v = torch.randn_like(q, requires_grad=True)
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
ref = None
for name, be, dt in [("math fp32", SDPBackend.MATH, torch.float32), ("efficient fp32", SDPBackend.EFFICIENT_ATTENTION, torch.float32),
                     ("cudnn bf16", SDPBackend.CUDNN_ATTENTION, torch.bfloat16), ("flash bf16", SDPBackend.FLASH_ATTENTION, torch.bfloat16),
                     ("efficient bf16", SDPBackend.EFFICIENT_ATTENTION, torch.bfloat16)]:
    try:
        def fn():
            with sdpa_kernel(be):
                o = F.scaled_dot_product_attention(q.to(dt), k.to(dt), v.to(dt))
            o.float().sum().backward()
            return o
        o = fn().float()
        if ref is None: ref = o.detach()
        err = ((o.detach() - ref).abs().max() / ref.abs().max()).item()
        print("%-16s %8.1f us fwd+bwd   max-abs-normalised err vs math %.2e" % (name, bench(fn), err), flush=True)
    except Exception as e:
        print("%-16s unavailable: %s" % (name, str(e)[:120]), flush=True)
