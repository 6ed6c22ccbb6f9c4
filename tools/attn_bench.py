"""Kernel-level timing of the fused attention (forward, backward) at the model's shape against the PyTorch ops."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
n, heads, d, s = int(os.environ.get("N", 1)), 8, 48, 2400
e = heads * d
q, k, v = (torch.randn(n, e, s, device="cuda", requires_grad=True) for _ in range(3))
go = torch.randn(n, s, e, device="cuda")


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / iters


def own_fwd():
    return pkg.attention(q, k, v, heads)


def own_fwdbwd():
    pkg.attention(q, k, v, heads).backward(go)


def torch_fwdbwd():
    sp = lambda t: t.view(n, heads, d, s).transpose(2, 3)
    att = torch.softmax(torch.matmul(sp(q) * (1.0 / d ** 0.5), sp(k).transpose(-1, -2)), dim=-1)
    torch.matmul(att, sp(v)).transpose(1, 2).reshape(n, s, e).backward(go)


torch.backends.cuda.matmul.allow_tf32 = True
print("own forward        %8.1f us" % timeit(own_fwd))
print("own forward+bwd    %8.1f us" % timeit(own_fwdbwd))
print("torch forward+bwd  %8.1f us" % timeit(torch_fwdbwd))
