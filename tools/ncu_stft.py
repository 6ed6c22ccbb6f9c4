import importlib, sys, os, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
B = int(os.environ.get("B", "16"))
x = torch.randn(B, 8, 60 * 32000, device="cuda")
for _ in range(2):
    y = pkg.stft_magphase(x, 512, 112, True, False, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
y = pkg.stft_magphase(x, 512, 112, True, False, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
