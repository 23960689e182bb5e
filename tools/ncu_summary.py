#!/usr/bin/env python
"""Turn ncu captures brought back in gpurun_out/ into the small summaries committed under profiles/.

    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep  profiles/ncu_full_rNN.md
    python tools/ncu_summary.py list gpurun_out/launches.csv  profiles/ncu_launches_rNN.md

`full`  : one row per captured kernel launch of an `ncu --set full` report: duration, DRAM bytes
          (dram__bytes_read/write.sum), DRAM %, issue-active %, pipe utilisation, occupancy, registers,
          top stall reasons.
`list`  : the `--metrics gpu__time_duration.sum` launch list: per-kernel totals and shares of one step.
"""
import csv
import collections
import subprocess
import sys


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full(rep, dst):
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]

    def g(r, k, f=float):
        try:
            return f(r[idx[k]])
        except (KeyError, ValueError):
            return float("nan")

    lines = ["| kernel | grid x block | time us | DRAM rd MB | DRAM wr MB | DRAM % | issue % | alu % | fma % | lsu % | "
             "tensor % | warps % | regs | top stalls |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        name = r[idx["Kernel Name"]].replace("|", "/")[:58]
        st = sorted(((g(r, h), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall), reverse=True)
        st = [f"{n} {v:.2f}" for v, n in st if n not in ("selected",)][:4]
        lines.append("| %s | %s x %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %s |" % (
            name, r[idx["Grid Size"]].split(",")[0].strip("( "), r[idx["Block Size"]].split(",")[0].strip("( "),
            g(r, "gpu__time_duration.sum"), g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum"),
            g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g(r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
            g(r, "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
            g(r, "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"),
            g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), int(g(r, "launch__registers_per_thread")),
            ", ".join(st)))
    open(dst, "w").write(f"ncu --set full --clock-control none, source: {rep}\n"
                         f"(units: time {units[idx['gpu__time_duration.sum']]}, bytes {units[idx['dram__bytes_read.sum']]})\n\n"
                         + "\n".join(lines) + "\n")
    print(f"wrote {dst}: {len(rows)} launches")


def launch_list(path, dst):
    rows = [r for r in csv.reader(open(path)) if r and r[0] != ""]
    # find the header line of the csv part
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    tot = collections.OrderedDict()
    n = 0
    for r in rows[start + 1:]:
        if len(r) <= idx["Metric Value"] or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0]
        val = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        us = val / 1e3 if unit.startswith("ns") else (val if unit.startswith("us") else val * 1e3)
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += us
        n += 1
    total = sum(v[1] for v in tot.values())
    lines = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k[:70]} | {v[0]} | {v[1]:.1f} | {100 * v[1] / total:.1f} % |")
    open(dst, "w").write(f"ncu --metrics gpu__time_duration.sum --clock-control none, source: {path}\n"
                         f"{n} launches, {total:.1f} us in total (cold-cache, serialised: compare shares, not absolutes)\n\n"
                         + "\n".join(lines) + "\n")
    print(f"wrote {dst}: {n} launches")


CLASSES = [("gemm_tc_kernel", "b200_gemm_tcgen05"), ("dw3x3_tma_kernel", "b200_dwconv2d"), ("dw3x3_imma_kernel", "b200_dwconv2d"), ("dwconv_", "b200_dwconv2d"),
           ("conv_direct", "b200_conv2d_direct"), ("conv_stem_tc", "b200_conv2d_stem_tcgen05"), ("im2col", "b200_im2col_gemm_tcgen05"), ("gap_i8", "b200_global_avgpool"),
           ("pool_", "b200_pool"), ("softmax", "b200_softmax")]


def traffic(rep, dst):
    """profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, per kernel
    class as bench.py names them (what the `roofline.traffic` key of the bench line quotes)"""
    import json
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ur, uw = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
    acc = {}
    for r in rows:
        name = r[idx["Kernel Name"]]
        cls = next((c for key, c in CLASSES if key in name), None)
        if cls is None:
            continue
        b = float(r[idx["dram__bytes_read.sum"]]) * scale[ur] + float(r[idx["dram__bytes_write.sum"]]) * scale[uw]
        a = acc.setdefault(cls, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += b
        a[2] += float(r[idx["gpu__time_duration.sum"]])
    out = {k: {"launches_captured": v[0], "dram_bytes_per_launch": v[1] / v[0], "ncu_time_us_per_launch": v[2] / v[0]}
           for k, v in acc.items()}
    out["_source"] = f"ncu --set full --clock-control none, {rep}: one batch-256 MobileNetV1 int8 step"
    json.dump(out, open(dst, "w"), indent=1)
    print(f"wrote {dst}: {sorted(acc)}")


if __name__ == "__main__":
    {"full": full, "list": launch_list, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
