"""profiles/sass_summary.txt: which tensor-core / TMA / tensor-memory instructions each kernel of libgg_b200.so holds.

    python scripts/sass_summary.py > profiles/sass_summary.txt

cuobjdump -sass of the in-tree library, one row per kernel that contains at least one of the mnemonics below
(UTCIMMA = tcgen05.mma kind::i8, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
IMMA = legacy mma.sync integer MMA, ATOMS = shared-memory atomics, RED/ATOMG = global reductions)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "genomic_gnn_b200", "libgg_b200.so")
MNEMONICS = ["UTCIMMA.2CTA", "UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "IMMA", "LDSM", "ATOMS", "RED.", "ATOMG", "IDP.4A"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["cu++filt", *names], capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out)) if len(out) == len(names) else {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn), line):
                if mn == "UTCIMMA" and "UTCIMMA.2CTA" in line:
                    continue
                counts[cur][mn] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print("# cuobjdump -sass genomic_gnn_b200/libgg_b200.so  (arch:", ", ".join(sorted(set(arch))) + ")")
    print("# instruction counts per kernel; kernels without any of the mnemonics are omitted")
    print("%-78s %s" % ("kernel", "  ".join("%-12s" % m for m in MNEMONICS)))
    total = collections.Counter()
    for fn, c in counts.items():
        if not sum(c.values()):
            continue
        total.update(c)
        name = demangle.get(fn, fn)
        name = re.sub(r"\(.*", "", name)[:76]
        print("%-78s %s" % (name, "  ".join("%-12d" % c[m] for m in MNEMONICS)))
    print("%-78s %s" % ("TOTAL", "  ".join("%-12d" % total[m] for m in MNEMONICS)))


if __name__ == "__main__":
    sys.exit(main())
