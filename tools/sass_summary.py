"""SASS opcode summary of the shipped library: what proves the kernels are Blackwell-native
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG).

    python tools/sass_summary.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200", "lib", "libdhg_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "HMMA", "HGMMA", "FFMA2", "FADD2", "FMUL2",
         "MUFU.TANH", "MUFU.EX2", "LDGSTS", "SYNCS", "UCGABAR", "ELECT", "FFMA", "REDG", "LDS.128"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    per[cur][w] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("SASS opcode summary of lib/libdhg_b200.so (cuobjdump -sass, sm_100a)")
    print("total: " + "  ".join(f"{w}={total[w]}" for w in WATCH if total[w]))
    print()
    groups = collections.OrderedDict()
    for fn, c in per.items():
        d = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        key = re.sub(r"<.*", "", d.replace("void ", "").replace("dhg::(anonymous namespace)::", "").replace("dhg::", ""))
        g = groups.setdefault(key, [0, collections.Counter()])
        g[0] += 1
        g[1].update(c)
    for key, (n, c) in groups.items():
        print(f"{key} x{n} instance(s): " + ("  ".join(f"{w}={c[w]}" for w in WATCH if c[w]) or "(none of the watched opcodes)"))


if __name__ == "__main__":
    sys.exit(main())
