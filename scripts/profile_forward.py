"""One generator forward (for ncu / AB_*_DEBUG_TIMING):
   python scripts/profile_forward.py [B] [T] [precision] [iters] [workload=hifigan_v1|bigvgan_base|bigvgan_large|mel]"""
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
workload = sys.argv[5] if len(sys.argv) > 5 else "hifigan_v1"
if workload == "mel":
    from amphion_b200.stft import TacotronSTFT
    y = ((torch.rand(B, T, generator=torch.Generator().manual_seed(0)) * 2 - 1) * 0.9).cuda()
    taco = TacotronSTFT(1024, 256, 1024, 80, 22050, 0, 8000)
    from amphion_b200 import mel as M
    for _ in range(iters):
        out = M.native_stft_mel(y, 1024, 256, 1024, taco.stft_fn.fft_window, taco.mel_basis.cuda(), 512, 0.0,
                                want_energy=True, fused=os.environ.get("AMPHION_B200_MEL", "fused") != "cufft")
    torch.cuda.synchronize()
    print("ok", tuple(out[1].shape))
    sys.exit(0)
from amphion_b200.vocoders import _vocoders
w = bench.WORKLOADS[workload]
torch.manual_seed(1234)
m = _vocoders[w["kind"]](bench.make_cfg(workload)).cuda().eval()
if w["kind"] == "bigvgan":
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith(".alpha") or n.endswith(".beta"):
                p.copy_((torch.randn(p.shape, generator=g) * 0.3).cuda())
m.precision = prec
mel = torch.randn(B, w["n_mel"], T, device="cuda")
with torch.no_grad():
    for _ in range(iters):
        wv = m(mel)
torch.cuda.synchronize()
print("ok", tuple(wv.shape), float(wv.abs().max()))
