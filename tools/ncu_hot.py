#!/usr/bin/env python
"""Top stall sites of one kernel launch from an `ncu --set full --import-source on` report.

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep <kernel-name-regex> [launch-skip] [top]

Prints, per SASS instruction with the most warp-stall samples, its executed count and the three
dominant stall reasons -- the per-instruction view behind the summaries in profiles/.
"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0][1][:110] if rows and len(rows[0]) > 1 else "")
    hdr = rows[h]
    data = []
    for r in rows[h + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break  # next launch
        if len(r) == len(hdr):
            data.append(r)
    idx = {k: i for i, k in enumerate(hdr)}
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    inst = sum(int(r[idx["Instructions Executed"]]) for r in data)
    print(f"samples {tot}  warp instructions {inst}")
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[idx[k]]) for r in data) for k in stalls}
    print("stall totals:", ", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top_n]:
        s = sorted(((int(r[idx[k]]), k[6:]) for k in stalls), reverse=True)[:3]
        print(f"{int(r[idx['# Samples']]):6d} {int(r[idx['Instructions Executed']]):9d}  {r[idx['Source']].strip()[:64]:64s} "
              + " ".join(f"{n}:{c}" for c, n in s if c))


if __name__ == "__main__":
    main()
