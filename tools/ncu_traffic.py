"""Per-kernel DRAM traffic and key metrics from an `ncu --set full` report of ONE train step (tools/profile_step.py):

  python tools/ncu_traffic.py gpurun_out/r02_step.ncu-rep profiles/r02_ncu_step.md profiles/r02_traffic.json

Writes a table (one line per launch, in launch order) and the bench.py traffic table {kernel group: dram bytes per
launch}; kernel groups are matched by launch ORDER inside the step (DESIGN.md section 5 lists the step's launches).
"""
import csv
import io
import json
import subprocess
import sys

WANT = {"gpu__time_duration.sum": "ns", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor%",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(src, dst_md, dst_json):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        rec = {"name": r[col["Kernel Name"]]}
        for m, short in WANT.items():
            if m in col:
                raw, u = r[col[m]], units[col[m]]
                try:
                    rec[short] = to_bytes(raw, u) if short in ("rd", "wr") else float(raw.replace(",", ""))
                except ValueError:
                    rec[short] = None
        if rec.get("ns") is not None and units[col["gpu__time_duration.sum"]] in ("usecond", "us"):
            rec["ns"] *= 1e3
        if rec.get("ns") is not None and units[col["gpu__time_duration.sum"]] in ("msecond", "ms"):
            rec["ns"] *= 1e6
        recs.append(rec)
    total = sum(r["ns"] or 0 for r in recs)
    with open(dst_md, "w") as f:
        f.write(f"# ncu --set full, one eager train step (tools/profile_step.py), source: {src}\n\n")
        f.write("Launch order; times are cold-cache and serialised under ncu (compare shares); dram = dram__bytes_read + write.\n\n")
        f.write("| # | kernel | us | share | dram MB | dram % | tensor % | sm % | warps % | regs | grid |\n|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for i, r in enumerate(recs):
            nm = r["name"].replace("void ", "").replace("ebk::<unnamed>::", "").split("(")[0][:70]
            dram = ((r.get("rd") or 0) + (r.get("wr") or 0)) / 1e6
            f.write(f"| {i} | `{nm}` | {(r['ns'] or 0) / 1e3:.1f} | {100 * (r['ns'] or 0) / total:.1f}% | {dram:.1f} | "
                    f"{r.get('dram%') or 0:.0f} | {r.get('tensor%') or 0:.0f} | {r.get('sm%') or 0:.0f} | {r.get('warps%') or 0:.0f} | "
                    f"{int(r.get('regs') or 0)} | {int(r.get('grid') or 0)} |\n")
        f.write(f"\ntotal {total / 1e6:.3f} ms over {len(recs)} launches; DRAM total {sum((r.get('rd') or 0) + (r.get('wr') or 0) for r in recs) / 1e9:.2f} GB\n")
    # traffic table: match kernel groups by name pattern + order
    def dram(r):
        return (r.get("rd") or 0) + (r.get("wr") or 0)
    big = [r for r in recs if (r["ns"] or 0) > 60e3]           # the news-encoder-sized launches
    table = {}
    gem = [r for r in big if "gemm_tma_kernel" in r["name"]]
    names = {"embed_rows": "news.embed_gather", "attn_fwd_pre": "news.attn_core_fwd", "attn_bwd_pre": "news.attn_core_bwd",
             "attpool_fwd": "news.attpool_fwd", "attpool_bwd_fused": "news.attpool_bwd", "embed_adam_kernel": "news.adam"}
    for r in big:
        for pat, key in names.items():
            if pat in r["name"] and key not in table:
                table[key] = dram(r)
    # big GEMMs in launch order: [qkv fwd (fused: ADH>0)], att fwd, att wgrad, att dgrad, qkv dgrad, qkv wgrad
    order = ["news.qkv_gemm_fwd", "news.att_gemm_fwd", "news.att_wgrad_gemm", "news.att_dgrad_gemm", "news.qkv_dgrad_gemm", "news.qkv_wgrad_gemm"]
    for key, r in zip(order, gem):
        table[key] = dram(r)
    with open(dst_json, "w") as f:
        json.dump(table, f, indent=1)
    print("wrote", dst_md, dst_json, table)


if __name__ == "__main__":
    main(*sys.argv[1:4])
