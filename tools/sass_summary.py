"""SASS evidence for the Blackwell-native kernels: per-kernel counts of the mnemonics that prove tcgen05 / TMEM / TMA
(B200_PROFILING.md "What proves a Blackwell-native kernel") from `cuobjdump -sass` of the in-tree libebk.so.

  python tools/sass_summary.py [profiles/r02_sass_summary.md]
"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "ebnerd-benchmark_b200" / "csrc" / "libebk.so"
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "REDG", "RED."]


def main(dst):
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    chunks = re.split(r"\n\s*Function : \S+\n", "\n" + sass)[1:]
    rows = OrderedDict()
    for name, body in zip(names, chunks):
        short = re.sub(r"^void |ebk::\(anonymous namespace\)::|ebk::", "", name).split("(")[0]
        counts = [len(re.findall(r"\b" + re.escape(m), body)) for m in MNEMONICS]
        n_inst = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", body, flags=re.M))
        r = rows.setdefault(short, [0] * (len(MNEMONICS) + 2))
        r[0] += 1
        r[1] += n_inst
        for i, c in enumerate(counts):
            r[2 + i] += c
    total = [sum(r[i] for r in rows.values()) for i in range(len(MNEMONICS) + 2)]
    with open(dst, "w") as f:
        f.write(f"# SASS mnemonic counts per kernel (`cuobjdump -sass {LIB.relative_to(ROOT)}`, sm_100a)\n\n")
        f.write("`UTCHMMA` = tcgen05.mma (`.2CTA` variants counted too), `LDTM`/`STTM` = tcgen05.ld/st (TMEM), `UTMALDG`/`UTMASTG`/"
                "`UTMAREDG` = TMA tensor load / store / reduce-add, `UTCBAR` = tcgen05.commit, `HMMA` = mma.sync (warp-level tensor "
                "path of the per-head 30x30x20 attention products), `LDGSTS` = cp.async.  Template instantiations of one kernel are summed.\n\n")
        f.write("| kernel | instantiations | SASS instructions | " + " | ".join(MNEMONICS) + " |\n")
        f.write("|---|---:|---:|" + "---:|" * len(MNEMONICS) + "\n")
        for k, r in sorted(rows.items(), key=lambda kv: -kv[1][2]):
            f.write(f"| `{k[:70]}` | " + " | ".join(str(x) for x in r) + " |\n")
        f.write("| **total** | " + " | ".join(str(x) for x in total) + " |\n")
    print("wrote", dst, "kernels:", len(rows), "UTCHMMA:", total[2])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "profiles" / "r02_sass_summary.md"))
