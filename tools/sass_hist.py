#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of csi-nn2_b200/lib/libb200nn.so: the mnemonics that prove a Blackwell-native
kernel (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit,
IDP.4A = dp4a, HMMA would be the legacy mma.sync path).

    python tools/sass_hist.py [lib] > profiles/sass_histogram_rNN.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "IDP", "HMMA", "IMMA", "LDGSTS"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "csi-nn2_b200", "lib", "libb200nn.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::", "")
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    # merge template instantiations of the same kernel
    merged = collections.OrderedDict()
    for name, c in per.items():
        base = re.sub(r"<.*", "", name)
        m = merged.setdefault(base, [0, collections.Counter()])
        m[0] += 1
        m[1].update(c)
    print(f"SASS opcode counts per kernel family of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; instantiations summed)\n")
    print("| kernel | variants | instructions | " + " | ".join(KEY) + " |")
    print("|---|---|---|" + "---|" * len(KEY))
    tot = collections.Counter()
    for base, (n, c) in merged.items():
        print(f"| {base} | {n} | {c['_total']} | " + " | ".join(str(c.get(k, 0)) for k in KEY) + " |")
        tot.update(c)
    print(f"| **all** | {sum(n for n, _ in merged.values())} | {tot['_total']} | " + " | ".join(str(tot.get(k, 0)) for k in KEY) + " |")


if __name__ == "__main__":
    main()
