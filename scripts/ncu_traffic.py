"""Turn an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` log of ONE forward
into profiles/r2_traffic.json entries: per kernel class the launches, DRAM bytes per launch and per forward.

    python scripts/ncu_traffic.py <workload> <batch> <frames> <precision> <log.csv> [--skip N] [--out profiles/r2_traffic.json]

Kernel classes follow `ab_generator_get_profile`: tc_conv = rb_kernel + tc_conv_kernel, tc_gemmconv = gemmconv*,
activation1d, conv1d_fp32 (conv_post), mel = frame_window + cuFFT + mag_mel."""
import csv
import json
import os
import sys


def classify(name):
    if "rb_kernel" in name or "tc_conv_kernel" in name:
        return "tc_conv"
    if "gemmconv" in name:
        return "tc_gemmconv"
    if "activation1d" in name:
        return "activation1d"
    if "conv_post" in name or "conv1d_f" in name:
        return "conv1d_fp32"
    if "convT_fp32" in name:
        return "conv_transpose1d_fp32"
    if "mel_fused" in name or "frame_window" in name or "mag_mel" in name or "fft" in name.lower() or "mel_span" in name:
        return "mel"
    return None


def main():
    a = sys.argv[1:]
    out = "profiles/r2_traffic.json"
    skip = 0
    if "--out" in a:
        i = a.index("--out"); out = a[i + 1]; del a[i:i + 2]
    if "--skip" in a:
        i = a.index("--skip"); skip = int(a[i + 1]); del a[i:i + 2]
    workload, batch, frames, precision, log = a[0], int(a[1]), int(a[2]), a[3], a[4]
    rows = {}
    for r in csv.reader(open(log, errors="replace")):
        if len(r) < 15 or not r[0].isdigit():
            continue
        rid = int(r[0])
        rows.setdefault(rid, {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))
        rows[rid]["unit_" + r[12]] = r[13]
    ids = sorted(rows)[skip:]
    kern, detail = {}, []
    for rid in ids:
        e = rows[rid]
        cls = classify(e["name"])
        if cls is None:
            continue
        def val(m):
            v, u = e.get(m, 0.0), e.get("unit_" + m, "byte")
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1}.get(u, 1)
        b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        k = kern.setdefault(cls, dict(launches=0, bytes=0.0, ns=0.0))
        k["launches"] += 1; k["bytes"] += b; k["ns"] += val("gpu__time_duration.sum")
        detail.append(dict(kernel=e["name"].split("(")[0].split("::")[-1][:48], cls=cls, dram_bytes=b,
                           us=val("gpu__time_duration.sum") / 1e3))
    entry = dict(batch=batch, frames=frames, precision=precision,
                 how="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over one forward (cold, serialised)",
                 kernels={c: dict(launches=k["launches"], bytes_per_launch=k["bytes"] / k["launches"],
                                  bytes_per_forward=k["bytes"], ms_under_ncu=k["ns"] / 1e6) for c, k in kern.items()},
                 total_bytes_per_forward=sum(k["bytes"] for k in kern.values()), launches=detail)
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[workload] = entry
    json.dump(data, open(out, "w"), indent=1)
    print(workload, {c: (k["launches"], round(k["bytes"] / 1e9, 2)) for c, k in kern.items()}, "GB per forward")


if __name__ == "__main__":
    main()
