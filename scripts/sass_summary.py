"""Counts the Blackwell-specific SASS mnemonics per kernel of libnrrt.so (cuobjdump -sass) -> profiles/sass_r2.md.
Evidence that the conv kernels are tcgen05 (UTCHMMA, .2CTA = cta_group::2), TMA (UTMALDG / UTMASTG / UBLKCP) and
tensor-memory (LDTM) code, independent of any launch list."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mgr_sampling_cnn_b200", "libnrrt.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "SYNCS", "DFMA", "DMUL", "DADD", "HMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::", "", re.sub(r"\(CUtensorMap_st.*", "", n)) for n in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            base = op.split(".")[0]
            counts[cur][base] += 1
            if op.startswith("UTCHMMA.2CTA"):
                counts[cur]["UTCHMMA.2CTA"] += 1
    names = demangle(order)
    rows = []
    for raw, name in zip(order, names):
        c = counts[raw]
        rows.append((name, c["_total"], [c[k] for k in KEYS]))
    rows.sort(key=lambda r: r[0])
    out = ["# SASS summary of libnrrt.so (round 2)", "",
           "`python scripts/sass_summary.py` (cuobjdump -sass of the in-tree library, sm_100a).  Counts of static instructions per kernel.",
           "UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA), UTMAPF = TMA",
           "prefetch, UBLKCP = cp.async.bulk, LDTM = tcgen05.ld (tensor memory -> registers), SYNCS = mbarrier ops, D* = fp64.", "",
           "| kernel | instr | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    tot = collections.Counter()
    for name, total, vals in rows:
        out.append(f"| `{name}` | {total} | " + " | ".join(str(v) if v else "" for v in vals) + " |")
        for k, v in zip(KEYS, vals):
            tot[k] += v
    out.append("| **all kernels** | | " + " | ".join(str(tot[k]) for k in KEYS) + " |")
    path = os.path.join(ROOT, "profiles", "sass_r2.md")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(f"wrote {path}: {len(rows)} kernels; totals " + ", ".join(f"{k} {tot[k]}" for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
