"""Per-kernel census of the Blackwell-only SASS instructions in libtimewarp_b200.so (cuobjdump -sass):
UTCHMMA(.2CTA) = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCCP = tcgen05.cp, UBLKCP = cp.async.bulk (TMA engine, 1-D),
UTCBAR = tcgen05.commit, SYNCS = mbarrier.  Usage: python tools/sass_census.py > profiles/sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "timewarp_b200", "libtimewarp_b200.so")
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTCCP", "UBLKCP.S.G", "UBLKCP.G.S", "UTMALDG", "UTMASTG",
            "UTCBAR", "SYNCS", "HMMA", "FFMA", "DFMA", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            cur = re.sub(r"\(.*", "", cur).replace("tw::", "").replace("(anonymous namespace)::", "")
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["_total"] += 1
        for p in PATTERNS:
            if op == p or op.startswith(p + ".") or (p in ("UTCHMMA",) and op.startswith("UTCHMMA") and ".2CTA" not in op and p == "UTCHMMA"):
                if p == "UTCHMMA" and ".2CTA" in op:
                    continue
                per[cur][p] += 1
                break
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; arch {sorted(set(arch))}); counts of static instructions per kernel")
    cols = [p for p in PATTERNS if any(c[p] for c in per.values())]
    print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = collections.Counter()
    for k, c in per.items():
        if not any(c[p] for p in cols if p not in ("FFMA", "DFMA", "MUFU", "SYNCS")):
            continue  # CUDA-core-only kernels are listed in the summary line only
        print(f"{k[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[p]:12d}" for p in cols))
        tot.update(c)
    allk = collections.Counter()
    for c in per.values():
        allk.update(c)
    print(f"{'TOTAL (kernels above)':58s} {tot['_total']:7d} " + " ".join(f"{tot[p]:12d}" for p in cols))
    print(f"{'TOTAL (all ' + str(len(per)) + ' kernels)':58s} {allk['_total']:7d} " + " ".join(f"{allk[p]:12d}" for p in cols))


if __name__ == "__main__":
    sys.exit(main())
