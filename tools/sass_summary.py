"""Per-kernel SASS evidence from the built library: counts of the instructions that show what the kernels are made of
(DMMA = FP64 tensor-core MMA, UTMALDG = TMA tensor load, UBLKCP = bulk async copy, SYNCS = mbarrier, BAR, MEMBAR, ...).

  python tools/sass_summary.py > profiles/r02_sass_summary.txt          (runs cuobjdump -sass on sparseeigen_b200/libspeigen_b200.so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparseeigen_b200", "libspeigen_b200.so")
WATCH = ["DMMA", "DFMA", "UTMALDG", "UBLKCP", "SYNCS", "BAR", "MEMBAR", "ERRBAR", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU",
         "UTCHMMA", "UTCQMMA", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w):
                    kernels[cur][w] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, code objects for {', '.join(arch)}")
    print("# no tcgen05 (UTC*MMA / LDTM) is expected: tcgen05.mma has no FP64 kind; FP64 tensor work is DMMA.8x8x4 fed by TMA + mbarrier")
    tot = collections.Counter()
    print(f"{'kernel':<78} {'instr':>7} " + " ".join(f"{w:>7}" for w in WATCH))
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"speigen::\(anonymous namespace\)::|speigen::", "", dn)
        short = re.sub(r"\(.*$", "", short)[:77]
        print(f"{short:<78} {c['_total']:>7} " + " ".join(f"{c[w]:>7}" for w in WATCH))
        tot.update(c)
    print(f"{'TOTAL':<78} {tot['_total']:>7} " + " ".join(f"{tot[w]:>7}" for w in WATCH))


if __name__ == "__main__":
    sys.exit(main())
