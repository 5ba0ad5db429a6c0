"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` output.
  ncu -i rep --page source --csv --kernel-name regex:K --launch-skip n --launch-count 1 > src.csv
  python tools/ncu_source_hot.py src.csv [topN]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # n-th kernel section of the file
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = starts[which]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
print("sections:", len(starts), "kernel:", rows[hi - 1][1][:110])
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0] != "Address"]
c = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[c["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {s: sum(int(r[c[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
print("inst executed (warp):", sum(int(r[c["Instructions Executed"]] or 0) for r in body))
for r in sorted(body, key=lambda r: -int(r[c["# Samples"]] or 0))[:top]:
    st = {s[6:]: int(r[c[s]] or 0) for s in stalls if int(r[c[s]] or 0)}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{int(r[c["# Samples"]]):7d} {100*int(r[c["# Samples"]])/tot:5.1f}%  ex={r[c["Instructions Executed"]]:>9}  {r[c["Source"]][:70]:70s} {st}')
