#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nsfhifigan or vits or wide" 2>&1 | tail -4 | tee gpurun_out/nsf_pytest.log
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tail -4 | tee gpurun_out/nsf_bench.log
import sys, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import HP_NSF_EXP, build_model
m = build_model("nsfhifigan", HP_NSF_EXP, 100, seed=3).cuda()
B, T = 16, 1024
mel = torch.randn(B, 100, T, device="cuda"); f0 = torch.rand(B, T, device="cuda") * 300 + 80
for prec in ("tc_f16",):
    m.precision = prec
    for _ in range(3): m(mel, f0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.set_profiling(True); m.get_profile()
    e0.record()
    for _ in range(5): w = m(mel, f0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("nsfhifigan exp_config B=16 T=1024", prec, "ms/step %.2f" % ms, "samples/s %.1fM" % (B * T * 256 / ms / 1e3),
          {k: round(v["ms"] / 5, 2) for k, v in m.get_profile().items()})
PY
