"""profiles/sass_summary.txt: per kernel of libdlux_b200.so, the count of the SASS mnemonics that prove the
Blackwell-native path (tcgen05 MMA = UTC*MMA, TMA = UTMALDG/UTMASTG, tensor memory = LDTM/STTM, ...).

    python tools/sass_summary.py [out.txt]
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dlux_b200", "lib", "libdlux_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS",
        "SYNCS", "MUFU.SIN", "MUFU.COS", "FFMA", "HMMA", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "SHFL", "ELECT"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            cur = re.sub(r"\(.*", "", name).split("::")[-1]
            kernels[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k] += 1
            if op.startswith("UTCHMMA") or op.startswith("UTMALDG") or op.startswith("UTMASTG") or op.startswith("UTCBAR"):
                kernels[cur]["full:" + op] += 1
    with open(out, "w") as f:
        f.write("# cuobjdump -sass dlux_b200/lib/libdlux_b200.so, instruction counts per kernel (sm_100a)\n")
        for name, c in kernels.items():
            f.write(f"\n{name}: {c['_total']} SASS instructions\n")
            for k in KEYS:
                if c[k]:
                    f.write(f"  {k:14s} {c[k]}\n")
            for k in sorted(c):
                if k.startswith("full:"):
                    f.write(f"    {k[5:]:48s} {c[k]}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
