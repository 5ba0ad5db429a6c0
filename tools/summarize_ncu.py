"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_ncu_full.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = name.replace("void ", "").replace("ebk::<unnamed>::", "").replace("(anonymous namespace)::", "")
    return name.split("(")[0][:90]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows[1:]:
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        k = short(r[i_name])
        agg[k][0] += 1
        agg[k][1] += v
        total += v
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "ns"
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), source: {src}\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"| kernel | launches | total ({unit}) | share |\n|---|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {v:,.0f} | {100 * v / total:.1f}% |\n")
        f.write(f"| **total** | {sum(n for n, _ in agg.values())} | {total:,.0f} | 100% |\n")
    print("wrote", dst)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary, source: {src}\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in WANT:
                if m in hdr:
                    f.write(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
