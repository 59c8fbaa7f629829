#!/usr/bin/env python
"""Turns the scratch ncu outputs in gpurun_out/ into the committed summaries under profiles/.

  python scripts/summarise_profiles.py r01      # reads gpurun_out/{launches.csv,prof_*.ncu-rep,bench.json}
writes profiles/<tag>_launches.md, profiles/<tag>_ncu_<name>.md, profiles/traffic.json, profiles/<tag>_bench.json
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__cycles_active.avg", "smsp__inst_executed.sum"]


def raw_rows(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def ncu_summary(tag, name, rep):
    hdr, units, rows = raw_rows(rep)
    ki = hdr.index("Kernel Name")
    # also keep every tensor-pipe metric present in the report
    extra = [h for h in hdr if "pipe_tensor" in h and h not in KEYS]
    lines = [f"# ncu --set full: {name} ({tag})", "",
             "`ncu --set full --clock-control none --import-source on` under gpurun; one block per captured launch.", ""]
    out = []
    for r in rows:
        d = {}
        for k in KEYS + extra:
            if k in hdr and r[hdr.index(k)] != "":
                d[k] = (r[hdr.index(k)], units[hdr.index(k)])
        out.append((r[ki], d))
        lines.append(f"## {r[ki][:110]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k, (v, u) in d.items():
            lines.append(f"| {k} | {v} | {u} |")
        if "dram__bytes_read.sum" in d and "gpu__time_duration.sum" in d:
            def to_bytes(v, u):
                f = float(v.replace(",", ""))
                return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            def to_s(v, u):
                f = float(v.replace(",", ""))
                return f * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1e-9)
            rd = to_bytes(*d["dram__bytes_read.sum"]); wr = to_bytes(*d["dram__bytes_write.sum"])
            t = to_s(*d["gpu__time_duration.sum"])
            lines.append("")
            lines.append(f"DRAM traffic {(rd + wr) / 1e6:.2f} MB in {t * 1e6:.1f} us -> {(rd + wr) / t / 1e9:.0f} GB/s (cold-cache, serialised replay)")
        lines.append("")
    open(os.path.join(PROF, f"{tag}_ncu_{name}.md"), "w").write("\n".join(lines))
    return out


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[start + 1:]:
        try:
            v = float(r[vi].replace(",", "")) / 1e3
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
        total += v
    lines = [f"# kernel launch list ({tag})", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 1` "
             "(set-up launches included; cold-cache, serialised: compare SHARES, not absolutes).", "",
             "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:70]}` | {n} | {v:.1f} | {v / n:.1f} | {100 * v / total:.1f}% |")
    lines.append("")
    lines.append(f"total: {sum(n for n, _ in agg.values())} launches, {total / 1e3:.2f} ms")
    open(os.path.join(PROF, f"{tag}_launches.md"), "w").write("\n".join(lines))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    traffic = {}

    def bytes_of(d):
        def f(k):
            v, u = d[k]
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return f("dram__bytes_read.sum") + f("dram__bytes_write.sum")

    for name, key in (("dbscan", "db_scan_dram_bytes_per_launch"), ("conv", "conv1_fused_dram_bytes_per_launch"),
                      ("conv_mid", "conv_n128_dram_bytes_per_launch"), ("conv_heads", "conv_heads_dram_bytes_per_launch"),
                      ("conv_box64", "conv2_box64_dram_bytes_per_launch")):
        rep = os.path.join(OUT, f"prof_{name}.ncu-rep")
        if os.path.exists(rep):
            out = ncu_summary(tag, name, rep)
            best = max(out, key=lambda o: bytes_of(o[1]) if "dram__bytes_read.sum" in o[1] else 0)
            traffic[key] = bytes_of(best[1])
            traffic[key + "_kernel"] = best[0][:80]
    rep = os.path.join(OUT, "prof_solve.ncu-rep")
    if os.path.exists(rep):
        ncu_summary(tag, "solve", rep)
    traffic["source"] = f"profiles/{tag}_ncu_*.md (ncu --set full, largest captured launch)"
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    b = os.path.join(OUT, "bench.json")
    if os.path.exists(b) and os.path.getsize(b):
        json.dump(json.loads(open(b).read().strip().splitlines()[-1]), open(os.path.join(PROF, f"{tag}_bench.json"), "w"), indent=1)
    print("profiles written:", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
