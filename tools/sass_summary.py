#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libpegasus_b200.so (cuobjdump -sass): the instructions that show the
sm_100a features each kernel uses.  Usage: python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt

UBLKCP = 1-D bulk TMA copy (cp.async.bulk), SYNCS = mbarrier arrive/wait, LDGSTS = cp.async gather,
FFMA2/FMUL2/FADD2 = packed f32x2 arithmetic, MUFU.EX2 = ex2.approx (fast numerics only),
ATOMS = shared-memory atomics (POPC.INC for the +1 tickets), REDUX/VOTE/MATCH = warp collectives."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU.EX2", "MUFU.RCP",
         "ATOMS", "ATOMG", "REDG", "RED", "REDUX", "VOTE", "MATCH", "SHFL", "BAR", "LDG", "STG", "LDS", "STS", "STL", "LDL"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pegasus_b200", "libpegasus_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    funcs = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel, %s, arch %s" % (os.path.basename(lib), ",".join(arch)))
    print("# (static instruction counts from `cuobjdump -sass`; tools/sass_summary.py)")
    tot = collections.Counter()
    for (name, c), dn in zip(funcs.items(), demangle):
        short = re.sub(r"\(.*", "", dn)
        short = re.sub(r"^void ", "", short)
        items = ["%s=%d" % (w, c[w]) for w in WATCH if c[w]]
        print("%-72s total=%-6d %s" % (short[:72], c["_total"], " ".join(items)))
        tot.update(c)
    print("%-72s total=%-6d %s" % ("ALL KERNELS", tot["_total"], " ".join("%s=%d" % (w, tot[w]) for w in WATCH if tot[w])))


if __name__ == "__main__":
    main()
