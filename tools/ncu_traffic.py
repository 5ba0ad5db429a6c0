"""Per-kernel DRAM traffic and key metrics of ONE eager train step (tools/profile_step.py) from the ncu metric list

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... \\
      --clock-control none --csv --log-file gpurun_out/r02_step_metrics.csv python tools/profile_step.py --no-defer
  python tools/ncu_traffic.py gpurun_out/r02_step_metrics.csv profiles/r02_ncu_step.md profiles/r02_traffic.json

Writes a table (one line per launch, in launch order) and the bench.py traffic table {kernel group: dram bytes per
launch}; kernel groups are matched by kernel name and launch ORDER inside the step.
"""
import csv
import json
import sys
from collections import OrderedDict

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6}


def main(src, dst_md, dst_json):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    ks = OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        ks.setdefault((int(r[ix["ID"]]), r[ix["Kernel Name"]]), {})[r[ix["Metric Name"]]] = (r[ix["Metric Value"]], r[ix["Metric Unit"]])

    def val(m, k):
        if k not in m:
            return 0.0
        v, u = m[k]
        return float(v.replace(",", "")) * SCALE.get(u, 1)

    recs = []
    for (i, name), m in ks.items():
        recs.append(dict(i=i, name=name, ns=val(m, "gpu__time_duration.sum"),
                         dram=val(m, "dram__bytes_read.sum") + val(m, "dram__bytes_write.sum"),
                         tensor=val(m, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                         drampct=val(m, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         warps=val(m, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                         sm=val(m, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                         regs=val(m, "launch__registers_per_thread"), grid=val(m, "launch__grid_size")))
    total = sum(r["ns"] for r in recs)
    with open(dst_md, "w") as f:
        f.write(f"# ncu metric list of ONE eager train step (tools/profile_step.py --no-defer), source: {src}\n\n")
        f.write("NRMS ebnerd_small shape, B=256, E=768 (bench default workload), CUDA-graph replay off so that every kernel is a "
                "separate named launch.  Launch order; times are cold-cache and serialised under ncu (compare SHARES); "
                "dram = dram__bytes_read.sum + dram__bytes_write.sum.\n\n")
        f.write("| # | kernel | us | share | dram MB | dram % | tensor % | sm % | warps % | regs | grid |\n|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in recs:
            nm = r["name"].replace("void ", "").replace("ebk::<unnamed>::", "").split("(")[0][:64]
            f.write(f"| {r['i']} | `{nm}` | {r['ns'] / 1e3:.1f} | {100 * r['ns'] / total:.1f}% | {r['dram'] / 1e6:.1f} | {r['drampct']:.0f} | "
                    f"{r['tensor']:.0f} | {r['sm']:.0f} | {r['warps']:.0f} | {int(r['regs'])} | {int(r['grid'])} |\n")
        f.write(f"\n**total {total / 1e6:.3f} ms over {len(recs)} launches; DRAM {sum(r['dram'] for r in recs) / 1e9:.2f} GB per step** "
                f"(round 1: 16.8 GB).\n")
    big = [r for r in recs if r["ns"] > 60e3]   # the news-encoder-sized launches
    table = {}
    names = {"embed_rows": "news.embed_gather", "attn_fwd_pre": "news.attn_core_fwd", "attn_bwd_pre": "news.attn_core_bwd",
             "attpool_fwd": "news.attpool_fwd", "attpool_bwd_fused": "news.attpool_bwd", "embed_adam_kernel": "news.adam"}
    for r in big:
        for pat, key in names.items():
            if pat in r["name"] and key not in table:
                table[key] = r["dram"]
    # news-sized GEMMs in launch order of the step: fused QKV+attention fwd, AttLayer2 fwd, AttLayer2 dgrad, QKV dgrad,
    # AttLayer2 wgrad (behind the scatter), QKV wgrad
    gem = [r for r in big if "gemm_tma_kernel" in r["name"]]
    order = ["news.qkv_gemm_fwd", "news.att_gemm_fwd", "news.att_dgrad_gemm", "news.qkv_dgrad_gemm", "news.att_wgrad_gemm", "news.qkv_wgrad_gemm"]
    for key, r in zip(order, gem):
        table[key] = r["dram"]
    with open(dst_json, "w") as f:
        json.dump(table, f, indent=1)
    print("wrote", dst_md, dst_json, {k: round(v / 1e6) for k, v in table.items()})


if __name__ == "__main__":
    main(*sys.argv[1:4])
