"""Time APNet.forward (egs/vocoder/gan/apnet/exp_config.json widths) on cuda:0: python scripts/bench_apnet.py [B] [frames]"""
import os
import sys
from types import SimpleNamespace as NS

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from amphion_b200.vocoders import APNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 861
hp = dict(ASP_channel=512, ASP_resblock_kernel_sizes=[3, 7, 11], ASP_resblock_dilation_sizes=[[1, 3, 5]] * 3,
          ASP_input_conv_kernel_size=7, ASP_output_conv_kernel_size=7,
          PSP_channel=512, PSP_resblock_kernel_sizes=[3, 7, 11], PSP_resblock_dilation_sizes=[[1, 3, 5]] * 3,
          PSP_input_conv_kernel_size=7, PSP_output_R_conv_kernel_size=7, PSP_output_I_conv_kernel_size=7)
pre = dict(n_mel=80, n_fft=1024, hop_size=256, win_size=1024, extract_amplitude_phase=True, sample_rate=22050)
torch.manual_seed(0)
model = APNet(NS(preprocess=NS(**pre), model=NS(generator="apnet", apnet=NS(**hp)))).eval().cuda()
mel = torch.randn(B, 80, T, device="cuda")
for prec in ("tc_f16", "fp32"):
    model.precision = prec
    for _ in range(3):
        model(mel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        out = model(mel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"apnet {prec}: B={B} frames={T}: {ms:.3f} ms/forward, {B * T * 256 / ms / 1e3:.1f} M samples/s, launches {model.last_launches}")
