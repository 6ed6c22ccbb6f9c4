import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seldq = importlib.import_module("sound-event-localization-and-detection_b200")
model_mod = importlib.import_module(seldq.__name__ + ".seld_model")
torch.manual_seed(5); np.random.seed(5)
blk = model_mod.TC_Block(in_channels=128, domain="DQ", G=128, U=128, V=[128, 128], D=[2],
                         spatial_dropout_rate=0.5, use_bias_conv=False, batch_norm='BN').cuda().train()
x = torch.randn(2, 128, 128, device="cuda")
def run(seed_value):
    blk._drop_seed.fill_(seed_value)
    blk.zero_grad(set_to_none=True)
    xi = x.clone().requires_grad_(True)
    with seldq.precision("bf16"):
        y = seldq.fused.tcn_stack(xi, blk.ResBlocks, blk._drop_seed)
        y.square().mean().backward()
    torch.cuda.synchronize()
    return y.detach().clone(), xi.grad.clone()
y0, g0 = run(7)
for i in range(12):
    y, g = run(7)
    print(i, "y maxabs diff %.3e (max |y| %.3e)  g maxabs diff %.3e (max |g| %.3e)  y ok %s g ok %s" % (
        (y - y0).abs().max().item(), y0.abs().max().item(), (g - g0).abs().max().item(), g0.abs().max().item(),
        torch.allclose(y, y0, rtol=1e-4, atol=1e-5), torch.allclose(g, g0, rtol=1e-3, atol=1e-6)), flush=True)
print("== poisoned torch.empty")
orig = torch.empty
def poisoned(*a, **k):
    t = orig(*a, **k)
    if t.is_cuda:
        if t.dtype == torch.uint8: t.fill_(0xFF)
        elif t.is_floating_point(): t.fill_(float('nan'))
    return t
torch.empty = poisoned
y, g = run(7)
print("poisoned: y nan %d of %d, g nan %d of %d, y maxabs diff vs y0 %.3e" % (
    torch.isnan(y).sum().item(), y.numel(), torch.isnan(g).sum().item(), g.numel(), (y - y0).abs().nan_to_num(0).max().item()))
