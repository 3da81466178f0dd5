"""One HiFi-GAN V1 forward (for ncu / AB_TC_DEBUG_TIMING): python scripts/profile_forward.py [B] [T] [precision] [iters]"""
import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
prec = sys.argv[3] if len(sys.argv) > 3 else "tc_f16"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
from amphion_b200.vocoders.hifigan import HiFiGAN
torch.manual_seed(1234)
m = HiFiGAN(bench.make_cfg()).cuda().eval()
m.precision = prec
mel = torch.randn(B, 80, T, device="cuda")
with torch.no_grad():
    for _ in range(iters):
        w = m(mel)
torch.cuda.synchronize()
print("ok", tuple(w.shape), float(w.abs().max()))
