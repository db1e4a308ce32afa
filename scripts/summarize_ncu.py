"""Condense ncu outputs into the small text files committed under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/r1_launches_bench.csv profiles/r1_launches_bench.md
  python scripts/summarize_ncu.py full gpurun_out/r1_prof_step.ncu-rep profiles/r1_ncu_full_step.md
  python scripts/summarize_ncu.py traffic gpurun_out/r2_prof_1024.ncu-rep 1024 [more.ncu-rep grid ...]
      -> profiles/ncu_traffic.json: dram bytes per launch by kernel class and grid (read by bench.py's roofline.traffic)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, items = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] == "gpu__time_duration.sum":
                items.append((d["Kernel Name"], d["Grid Size"], d["Block Size"], float(d["Metric Value"].replace(",", "")), d["Metric Unit"]))
    agg = collections.OrderedDict()
    for name, grid, block, val, unit in items:
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        k = (name.split("(")[0], grid, block)
        agg.setdefault(k, []).append(val * scale)
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).\n\n")
        f.write("| kernel | grid | block | launches | avg us | share of listed time |\n|---|---|---|---:|---:|---:|\n")
        for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{name}` | {grid} | {block} | {len(v)} | {sum(v)/len(v):.1f} | {100*sum(v)/total:.1f}% |\n")
        f.write("\nFirst 40 launches in order:\n\n```\n")
        for name, grid, block, val, unit in items[:40]:
            f.write(f"{name.split('(')[0][:60]:60s} {grid:>14s} {val:>12.1f} {unit}\n")
        f.write("```\n")


def raw_page(src):
    """The raw page of a capture as CSV text: exported here from a .ncu-rep, or read as is when the GPU-side script
    already exported it (the reports of the 1024^3 kernels exceed what gpurun copies back)."""
    if src.endswith(".csv"):
        return open(src).read()
    return subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def full(src, dst):
    out = raw_page(src)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n`--clock-control none`; one row block per captured launch.\n")
        for r in rows[2:]:
            f.write(f"\n## {r[idx['Kernel Name']].split('(')[0]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")


def kernel_class(name, order):
    """bench.py's kernel classes (Solver::tick names) from the demangled kernel name and launch order."""
    if "k_fused_kspace" in name:
        return "fused_kspace"
    if "k_fused_real" in name:
        return "fused_real"
    if "k_pass_strided" in name:
        # inside the fused step the middle passes alternate inverse, forward
        return "pass_inverse_mid" if order % 2 == 0 else "pass_forward_mid"
    return None


def traffic(args):
    import json
    import os
    entries = []
    for src, grid in zip(args[0::2], args[1::2]):
        out = raw_page(src)
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        acc = collections.OrderedDict()
        n_strided = 0
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            cls = kernel_class(name, n_strided)
            if "k_pass_strided" in name:
                n_strided += 1
            if cls is None:
                continue
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[idx[k]].replace(",", "")) * scale.get(units[idx[k]], 1.0)
            acc.setdefault((cls, name.split("(")[0]), []).append(tot)
        for (cls, name), v in acc.items():
            entries.append({"kernel": cls, "grid": int(grid), "dram_bytes": sum(v) / len(v), "launches": len(v),
                            "kernel_name": name, "source": os.path.basename(src)})
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full --clock-control none)",
               "entries": entries}, open(dst, "w"), indent=1)
    print(dst, len(entries), "entries")


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
