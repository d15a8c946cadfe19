#!/usr/bin/env python
"""Condense ncu output into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof_gemm.ncu-rep     > profiles/rNN_ncu_gemm.txt
    python tools/ncu_summary.py traffic  gpurun_out/traffic.csv           > profiles/rNN_gemm_dram_traffic.json
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "")[:70]
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}.get(r[ui], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): SHARES, not absolutes")
    print(f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {v[1] / tot:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none --import-source on; one block per captured launch ({path})")
    for r in rows[2:]:
        print(f"\n== {r[ki][:100]}")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:100s} {r[i]:>16s} {units[i]}")


def traffic(path):
    """CSV of `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` -> one JSON record:
    DRAM bytes per launch of every kernel family captured (the bench line's `roofline.traffic` reads the GEMM's)."""
    import json
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ii, ki, mi, vi, ui = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
             "ms": 1e3, "msecond": 1e3}
    per = collections.OrderedDict()   # launch id -> (family, {metric: value})
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        fam = r[ki].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "").replace("rt::", "")
        fam = fam.split("(")[0].split("<")[0].strip()
        d = per.setdefault(r[ii], (fam, {}))[1]
        d[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    out = collections.OrderedDict()
    for fam, d in per.values():
        a = out.setdefault(fam, dict(launches=0, dram_read_gb=0.0, dram_write_gb=0.0, time_ms_under_ncu=0.0))
        a["launches"] += 1
        a["dram_read_gb"] += d.get("dram__bytes_read.sum", 0.0) / 1e9
        a["dram_write_gb"] += d.get("dram__bytes_write.sum", 0.0) / 1e9
        a["time_ms_under_ncu"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    for a in out.values():
        a["bytes_per_launch"] = (a["dram_read_gb"] + a["dram_write_gb"]) * 1e9 / a["launches"]
    print(json.dumps(dict(what="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                               "--clock-control none over the launches of ONE cfg2 step of bench.py (the first timed step; "
                               "warm-up launches skipped with --launch-skip)", families=out), indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
