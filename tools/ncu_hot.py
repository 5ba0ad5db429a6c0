"""Stall-sample summary of one ncu --set full --import-source on capture:

  python tools/ncu_hot.py gpurun_out/r02_x.ncu-rep [top]

prints the kernel name, duration / DRAM / issue figures, samples per 100-instruction SASS region and the hottest SASS
lines with their two leading stall reasons (what the optimisation notes in profiles/ quote)."""
import csv
import io
import subprocess
import sys


def page(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = list(csv.reader(io.StringIO(page(rep, "--page", "raw", "--csv"))))
    hdr, row = raw[0], raw[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "lts__t_sector_hit_rate.pct"]
    for w in want:
        if w in hdr:
            print(f"{w}: {row[hdr.index(w)][:120]} {raw[1][hdr.index(w)]}")
    rows = list(csv.reader(io.StringIO(page(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h, data = rows[1], rows[2:]
    iS, iI = h.index("# Samples"), h.index("Instructions Executed")
    tot = sum(int(r[iS]) for r in data)
    print("samples", tot, "warp-instructions", sum(int(r[iI]) for r in data), "sass lines", len(data))
    print("region: samples", " ".join(f"{b}:{sum(int(r[iS]) for r in data[b:b + 100])}" for b in range(0, len(data), 100)))
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    agg = {h[c]: sum(int(r[c]) for r in data) for c in stall}
    print("stalls:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:top]
    for i in sorted(idx):
        r = data[i]
        st = sorted(((int(r[c]), h[c][6:]) for c in stall), reverse=True)[:2]
        print(i, r[1].strip()[:64].ljust(64), r[iS], r[iI], st)


main()
