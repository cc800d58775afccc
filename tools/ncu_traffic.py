"""Summarise an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` launch list of
tools/profile_step.py into per-kernel-family DRAM traffic per launch -> profiles/r01_ncu_traffic.json (read by bench.py
for `roofline.traffic`) and a text table.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python tools/profile_step.py --warmup 1      # on the GPU box
    python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r01_ncu_traffic.json  # here
"""
import csv
import json
import sys

FAMILY = {"conv_gemm_kernel": "tap_gemm", "conv3x3_halo_kernel": "tap_gemm", "conv3x3_rows_kernel": "tap_gemm", "wgrad3x3_kernel": "wgrad", "wgrad_kernel": "wgrad",
          "wgrad_reduce_kernel": "wgrad", "wgrad_reduce_sliced_kernel": "wgrad", "wgrad_reduce_tiled_kernel": "wgrad",
          "wgrad_up_kernel": "wgrad", "wgrad3x3_2sm_kernel": "wgrad", "wgrad_reduce_upfold_kernel": "wgrad", "wgrad_reduce_upfold_sliced_kernel": "wgrad"}
TENSOR = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}


def main(src, dst):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        key = (r["ID"], r["Kernel Name"])
        per.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    # the profiled program runs 1 warm-up step + 1 measured step: keep the last half of the launches
    keys = list(per)
    keys = keys[len(keys) // 2:]
    fam, kern = {}, {}
    for key in keys:
        name = key[1].split("(")[0].split("::")[-1].split("<")[0]
        m = per[key]
        b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        t = m.get("gpu__time_duration.sum", 0.0)
        k = kern.setdefault(name, {"launches": 0, "bytes": 0.0, "seconds": 0.0, "tensor": 0.0})
        k["launches"] += 1; k["bytes"] += b; k["seconds"] += t; k["tensor"] += t * m.get(TENSOR, 0.0)
        if name in FAMILY:
            f_ = fam.setdefault(FAMILY[name], {"launches": 0, "bytes": 0.0, "seconds": 0.0})
            if "reduce" not in name:
                f_["launches"] += 1
            f_["bytes"] += b; f_["seconds"] += t
    out = {k: {"launches": v["launches"], "bytes_per_launch": v["bytes"] / max(1, v["launches"]),
               "ms_under_ncu": 1e3 * v["seconds"]} for k, v in fam.items()}
    out["_source"] = src
    out["_kernels"] = {k: {"launches": v["launches"], "MB_per_launch": v["bytes"] / max(1, v["launches"]) / 1e6,
                           "ms_under_ncu": 1e3 * v["seconds"]} for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["seconds"])}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    total = sum(v["seconds"] for v in kern.values())
    print(f"{'kernel':36s} {'launches':>8s} {'ms (ncu)':>9s} {'share':>6s} {'MB/launch':>10s} {'tensor pipe %':>14s}")
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["seconds"]):
        tp = f"{v['tensor'] / v['seconds']:14.1f}" if v["tensor"] > 0 else f"{'':14s}"
        print(f"{k:36s} {v['launches']:8d} {1e3 * v['seconds']:9.3f} {v['seconds'] / total:6.1%} {v['bytes'] / max(1, v['launches']) / 1e6:10.1f} {tp}")
    tt = sum(v["seconds"] for v in kern.values() if v["tensor"] > 0)
    if tt > 0:
        print(f"tensor-core launches: {1e3 * tt:.3f} ms, time-weighted tensor-pipe activity "
              f"{sum(v['tensor'] for v in kern.values()) / tt:.1f} % of peak sustained active")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
