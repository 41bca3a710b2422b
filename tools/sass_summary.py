"""Opcode counts per kernel of the shipped library (cuobjdump -sass), for profiles/rNN_sass.txt: the evidence that the
contraction and the tile movers are Blackwell-native (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA).

    python tools/sass_summary.py > profiles/r02_sass.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mergerec_b200", "csrc", "libmergerec_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA",
         "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "BAR", "SHFL", "FADD", "FMUL", "FFMA", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    names = list(per)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: static opcode counts per kernel (sm_100a)")
    print("# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS")
    total = collections.Counter()
    rows = []
    for n, d in zip(names, dem):
        c = per[n]
        total.update(c)
        short = re.sub(r"\(.*", "", d).replace("void ", "").replace("mr::", "")
        rows.append((short, c))
    agg = collections.OrderedDict()
    for short, c in rows:
        base = re.sub(r"<.*", "", short)
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1
        a[1].update(c)
    print(f"# {len(rows)} kernels (template instantiations) in {len(agg)} families\n")
    print("## library totals")
    print("  " + ", ".join(f"{k} {total[k]}" for k in WATCH if total[k]))
    print("\n## per kernel family (all instantiations summed)")
    for base, (n, c) in agg.items():
        keys = [k for k in WATCH if c[k]]
        print(f"{base}  [{n} instantiation(s), {c['_total']} SASS instructions]")
        print("    " + ", ".join(f"{k} {c[k]}" for k in keys))
    print("\n## the tensor-core / TMA kernels, per instantiation")
    for short, c in rows:
        if any(c[k] for k in ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "UTMALDG", "UBLKCP")):
            print(f"{short}")
            print("    " + ", ".join(f"{k} {c[k]}" for k in WATCH if c[k] and k not in ("FADD", "FMUL", "FFMA")))


if __name__ == "__main__":
    sys.exit(main())
