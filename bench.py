#!/usr/bin/env python
"""bench.py — vocoder audio samples/s on B200 (BASELINE.json metric, config 2).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (config.workload): HiFi-GAN V1 22.05 kHz generator forward, batch 64
per GPU, 80x1024 synthetic mel (random-init weights, torch.manual_seed(1234)).
A "step" is one generator forward over the per-GPU batch (+ one NCCL all-gather
of the wav shards when N > 1; weak scaling: the utterance batch grows with N).
`value` times the device-resident path with CUDA events; `e2e` times the
reference-facing call (vocoder_inference) with pinned HOST buffers, H2D and D2H
inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "vocoder_audio_samples_per_sec_22.05kHz"
UNIT = "samples/s"
HP_V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
             upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
             resblock_dilation_sizes=[[1, 3, 5]] * 3)
N_MEL, HOP = 80, 256
FLOP_PER_SAMPLE = 2398848          # conv FLOPs per output sample, HiFi-GAN V1 / BigVGAN-base
FLOP_PER_SAMPLE_LARGE = 7047456    # BigVGAN-large


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--precision", default=os.environ.get("AMPHION_B200_PRECISION", "tc_f16"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="hifigan_v1", choices=["hifigan_v1", "bigvgan_base", "bigvgan_large"],
                    help="hifigan_v1 = BASELINE config 2 (the headline, default); bigvgan_base = config 3 "
                         "(batch 32, 100x1024 mel, 24 kHz) for the results table")
    return ap.parse_args()


def make_cfg(workload="hifigan_v1"):
    from types import SimpleNamespace as NS
    if workload == "bigvgan_base":
        pre = NS(n_mel=100, hop_size=HOP, extract_amplitude_phase=False)
        hp = dict(HP_V1, activation="snakebeta", snake_logscale=True)
        return NS(preprocess=pre, model=NS(generator="bigvgan", bigvgan=NS(**hp)))
    if workload == "bigvgan_large":   # egs/vocoder/gan/bigvgan_large/exp_config.json:14-57
        pre = NS(n_mel=100, hop_size=HOP, extract_amplitude_phase=False)
        hp = dict(resblock="1", upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
                  upsample_initial_channel=1536, resblock_kernel_sizes=[3, 7, 11],
                  resblock_dilation_sizes=[[1, 3, 5]] * 3, activation="snakebeta", snake_logscale=True)
        return NS(preprocess=pre, model=NS(generator="bigvgan", bigvgan=NS(**hp)))
    pre = NS(n_mel=N_MEL, hop_size=HOP, extract_amplitude_phase=False)
    return NS(preprocess=pre, model=NS(generator="hifigan", hifigan=NS(**HP_V1)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------
# CPU baseline: the oracle port (oracle/generator.py follows the reference's
# hifigan.py:203-219 op for op with the same torch CPU primitives)
# --------------------------------------------------------------------------
_CPU_THREADS = None


def _pick_cpu_threads(params, og, torch):
    """The reference's CPU path runs torch's CPU convolutions; on many-core hosts they are FASTER with fewer
    threads than cores (oneDNN on small channel counts).  Probe a short forward at a few thread counts and
    keep the fastest, so the baseline is the strongest the host offers; `cores` reports the count used."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    ncpu = os.cpu_count() or 1
    cands = sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True)
    probe = torch.randn(1, N_MEL, 96, generator=torch.Generator().manual_seed(1))
    best = (float("inf"), ncpu)
    for n in cands:
        torch.set_num_threads(n)
        og.hifigan_forward(params, HP_V1, probe[:, :, :32])
        t0 = time.perf_counter()
        og.hifigan_forward(params, HP_V1, probe)
        dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, n)
    _CPU_THREADS = best[1]
    return _CPU_THREADS


def cpu_oracle_samples_per_sec(frames, batch=2, repeats=2):
    import torch
    from oracle import generator as og
    from amphion_b200.vocoders.hifigan import HiFiGAN
    torch.manual_seed(1234)
    model = HiFiGAN(make_cfg())
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    params = {}
    for k in list(sd):  # fold weight norm once, like remove_weight_norm(), outside the timed region
        if k.endswith(".weight_v"):
            params[k[:-2]] = og.fold_weight_norm(sd[k], sd[k[:-2] + "_g"])
        elif not k.endswith(".weight_g"):
            params[k] = sd[k]
    cores = _pick_cpu_threads(params, og, torch)
    torch.set_num_threads(cores)
    mel = torch.randn(batch, N_MEL, frames, generator=torch.Generator().manual_seed(0))
    og.hifigan_forward(params, HP_V1, mel[:1, :, : min(frames, 64)])      # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        og.hifigan_forward(params, HP_V1, mel)
        best = min(best, time.perf_counter() - t0)
    return (batch * frames * HOP / best, cores, best,
            f"B={batch} slice at full T={frames}, best of {repeats}, {cores} of {os.cpu_count()} host threads (fastest probed)")


def run_reference(args, rank):
    """The reference's own CPU path for the same workload (the oracle port: identical op sequence on the same
    torch CPU primitives), all host threads.  Each step is a bounded sample of the batch: ONE utterance at the
    full frame count (the generator has no cross-batch op, so cost is linear in B)."""
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    per, total_t, cores = [], 0.0, os.cpu_count() or 1
    for i in range(steps):
        _, cores, dt, _ = cpu_oracle_samples_per_sec(args.frames, batch=1, repeats=1)
        per.append(dt)
        total_t += dt
        if total_t > 150:      # keep the whole run within a few minutes
            steps = i + 1
            break
    ms = statistics.mean(per) * 1e3
    value = args.frames * HOP / (ms / 1e3)
    sample = "each step = 1 of the %d utterances of the batch at full T=%d, %d of %d host threads (fastest probed)" % (args.batch, args.frames, cores, os.cpu_count() or 1)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="HiFi-GAN V1 22.05kHz generator forward, batch=%d per GPU, 80x%d synthetic mel"
                                     % (args.batch, args.frames), global_batch=args.gpus * args.batch, frames=args.frames,
                            hop=HOP, precision="fp32 (CPU)", sample=sample),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, f"/tmp/ab_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in open(self.path):
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                    power_w_max=max(pw), samples=len(sm))


def run_native(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from amphion_b200.vocoders import vocoder_inference
    from amphion_b200.vocoders.hifigan import HiFiGAN

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_cfg(args.workload)
    torch.manual_seed(1234)
    if args.workload.startswith("bigvgan"):
        from amphion_b200.vocoders.bigvgan import BigVGAN
        model = BigVGAN(cfg).to(dev).eval()
        gsn = torch.Generator().manual_seed(1)
        with torch.no_grad():       # alpha/beta ~ N(0, 0.3): the default 0 is too benign
            for n, prm in model.named_parameters():
                if n.endswith(".alpha") or n.endswith(".beta"):
                    prm.copy_((torch.randn(prm.shape, generator=gsn) * 0.3).to(dev))
        if args.batch == 64:
            args.batch = 32
        if args.workload == "bigvgan_large" and args.frames == 1024:
            args.frames = 2048
    else:
        model = HiFiGAN(cfg).to(dev).eval()
    model.precision = args.precision
    B, T = args.batch, args.frames
    n_mel = cfg.preprocess.n_mel
    mel = torch.randn(B, n_mel, T, generator=torch.Generator().manual_seed(rank)).to(dev)
    samples_step = world * B * T * HOP
    gathered = torch.empty(world * B, 1, T * HOP, device=dev) if world > 1 else None

    def step():
        wav = model(mel)
        if world > 1:   # the path's one collective: final gather of the wav shards
            dist.all_gather_into_tensor(gathered, wav)
        return wav

    def barrier():
        if world > 1:
            dist.barrier()

    with torch.no_grad():
        nwarm = args.warmup if os.environ.get("AB_BENCH_PROFILE") else max(args.warmup, 3)
        for _ in range(nwarm):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        model.set_profiling(True)
        model.get_profile()
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize(); barrier()
        ms_total = e0.elapsed_time(e1)
        prof = model.get_profile()
        model.set_profiling(False)
        clocks = sampler.stop() if rank == 0 else None
        launches = model.last_launches * args.steps * world      # whole job (every rank runs the same pipeline)

        # ---- end to end through the reference-facing call, host buffers ----
        mel_host = mel.cpu().pin_memory()
        for _ in range(0 if os.environ.get("AB_BENCH_PROFILE") else 2):
            vocoder_inference(cfg, model, mel_host, device=dev)
        barrier(); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out_pinned = torch.empty(B, T * HOP, dtype=torch.float32, pin_memory=True) if world > 1 else None
        for _ in range(args.steps):
            if world == 1:
                out_host = vocoder_inference(cfg, model, mel_host, device=dev)   # H2D + forward + D2H (+sync)
            else:
                # same call sequence as vocoder_inference() plus the path's one collective before the D2H
                wav = model(mel_host.to(dev, non_blocking=True))
                dist.all_gather_into_tensor(gathered, wav)
                out_pinned.copy_(wav.squeeze(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
        t1.record()
        torch.cuda.synchronize(); barrier()
        e2e_ms_total = t0.elapsed_time(t1)

    tms = torch.tensor([ms_total, e2e_ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = tms.tolist()
    ms_step = ms_total / args.steps
    value = samples_step / (ms_step / 1e3)
    e2e_value = samples_step / (e2e_ms_total / args.steps / 1e3)

    if rank == 0:
        pk = peaks()
        dom = max(prof, key=lambda k: prof[k]["ms"])
        d = prof[dom]
        tf = d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0
        gbs = d["bytes"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else 0.0
        fps = FLOP_PER_SAMPLE_LARGE if args.workload == "bigvgan_large" else FLOP_PER_SAMPLE
        hbm_bound = dom == "activation1d"     # the anti-aliased Snake is an 8 B/element streaming kernel
        roofline = dict(bound="hbm" if hbm_bound else "tensor", kernel=dom + "_kernel",
                        achieved=gbs if hbm_bound else tf, peak=pk["hbm"] if hbm_bound else pk["tensor"],
                        unit="GB/s" if hbm_bound else "TFLOP/s",
                        frac=(gbs / pk["hbm"]) if hbm_bound else (tf / pk["tensor"]), traffic=None,
                        peak_source=pk["source"] + (", copy" if hbm_bound else ", bf16 sustained"),
                        launches=d["launches"], avg_launch_ms=d["ms"] / max(d["launches"], 1),
                        share_of_step=d["ms"] / ms_total,
                        tensor=dict(achieved=tf, peak=pk["tensor"], unit="TFLOP/s", frac=tf / pk["tensor"]),
                        hbm=dict(achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                                 note="algorithmic bytes: fp32 x read + y write (+ branch sum) + weights once"),
                        whole_step=dict(tflops=samples_step / world * fps / (ms_step / 1e3) / 1e12,
                                        frac_of_tensor_peak=samples_step / world * fps / (ms_step / 1e3) / 1e12 / pk["tensor"]),
                        classes={k: dict(launches=v["launches"], ms=round(v["ms"], 3)) for k, v in prof.items()})
        # DRAM traffic of the dominant kernel: measured once under ncu (profiles/r1_tc_traffic.json, same workload,
        # batch and precision), per launch like `achieved`; null when no capture matches this configuration
        roofline["algorithmic_bytes_per_launch"] = d["bytes"] / max(d["launches"], 1)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_tc_traffic.json")))
            if dom == "tc_conv" and args.precision == "tc_f16" and (tr["workload"], tr["batch"], tr["frames"]) == (args.workload, B, T):
                roofline["traffic"] = tr["bytes_per_launch"]
                roofline["traffic_source"] = "profiles/r1_tc_traffic.json (ncu dram__bytes_read+write, mean of the %d tc_conv launches of one forward)" % tr["launches"]
        except (OSError, KeyError, ValueError):
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.workload == "hifigan_v1":
            v, cores, dt, sample = cpu_oracle_samples_per_sec(T, batch=2, repeats=2)
            cpu = dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample, seconds=dt)
        dt = {"fp32": "f32", "tc_f16": "f16 operands, f32 accumulate (tcgen05); f32 elsewhere",
              "tc_bf16": "bf16 operands, f32 accumulate (tcgen05); f32 elsewhere"}[args.precision]
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=dt,
                    data="synthetic",
                    config=dict(workload={"hifigan_v1": "HiFi-GAN V1 22.05kHz", "bigvgan_base": "BigVGAN-base 24kHz", "bigvgan_large": "BigVGAN-large 24kHz"}[args.workload]
                                + " generator forward, batch=%d per GPU, %dx%d synthetic mel" % (B, n_mel, T),
                                global_batch=world * B, frames=T, hop=HOP, precision=args.precision,
                                parallelism="utterance-batch sharding dp%d, NCCL all-gather of wav" % world,
                                l2="no explicit flush: each step streams >8 GB of stage tensors (>> 126 MB L2)"),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=world * B * n_mel * T * 4,
                             d2h_bytes_per_step=world * B * T * HOP * 4, ms_per_step=e2e_ms_total / args.steps),
                    gpu_launches=launches, clocks=clocks, roofline=roofline, cpu_baseline=cpu, impl="native")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun exactly as the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference(args, rank)
    return run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
