"""Oracle helper: random generator parameters in the reference's state-dict layout,
already weight-norm-folded, built with plain numpy — no product code involved.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by `bench.py --impl reference`, the `cpu_baseline` legs
and the eager-GPU comparison leg, where only the SHAPES matter (timing), so the values are i.i.d. normal with the
standard deviation the reference's default init produces (`weight_v.std()` = 0.02)."""
from __future__ import annotations

import numpy as np


def random_generator_params(kind: str, hp: dict, n_mel: int, seed: int = 1234) -> dict:
    """kind: "hifigan" | "bigvgan".  Keys follow models/vocoders/gan/generator/hifigan.py:151-201 and
    bigvgan.py:231-311 (transposed convs under ``ups.{i}.0`` for BigVGAN, six ``activations`` per AMPBlock1)."""
    rng = np.random.default_rng(seed)
    p = {}

    def conv(name, cout, cin, k):
        p[name + ".weight"] = (rng.standard_normal((cout, cin, k)) * 0.02).astype(np.float32)
        p[name + ".bias"] = (rng.standard_normal(cout) * 0.02).astype(np.float32)

    def act(name, ch):
        p[name + ".act.alpha"] = (rng.standard_normal(ch) * 0.3).astype(np.float32)
        if hp.get("activation") == "snakebeta":
            p[name + ".act.beta"] = (rng.standard_normal(ch) * 0.3).astype(np.float32)

    big = kind == "bigvgan"
    c0 = hp["upsample_initial_channel"]
    conv("conv_pre", c0, n_mel, 7)
    nk = len(hp["resblock_kernel_sizes"])
    ch = c0
    for i, (u, k) in enumerate(zip(hp["upsample_rates"], hp["upsample_kernel_sizes"])):
        cin, ch = c0 >> i, c0 >> (i + 1)
        name = f"ups.{i}" + (".0" if big else "")
        p[name + ".weight"] = (rng.standard_normal((cin, ch, k)) * 0.02).astype(np.float32)   # [C_in, C_out, k]
        p[name + ".bias"] = (rng.standard_normal(ch) * 0.02).astype(np.float32)
        for j in range(nk):
            kk, ds = hp["resblock_kernel_sizes"][j], hp["resblock_dilation_sizes"][j]
            pre = f"resblocks.{i * nk + j}"
            for q in range(len(ds)):
                if str(hp["resblock"]) == "1":
                    conv(f"{pre}.convs1.{q}", ch, ch, kk)
                    conv(f"{pre}.convs2.{q}", ch, ch, kk)
                else:
                    conv(f"{pre}.convs.{q}", ch, ch, kk)
            if big:
                for a in range(len(ds) * (2 if str(hp["resblock"]) == "1" else 1)):
                    act(f"{pre}.activations.{a}", ch)
    if big:
        act("activation_post", ch)
    conv("conv_post", 1, ch, 7)
    return p
