#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/r1_xxx_full.txt
"""
import csv
import io
import subprocess
import sys

KEEP = ["Duration", "Registers Per Thread", "Grid Size", "Block Size", "Waves Per SM", "Theoretical Occupancy", "Achieved Occupancy",
        "Memory Throughput", "DRAM Throughput", "L2 Cache Throughput", "Compute (SM) Throughput", "L1/TEX Hit Rate", "L2 Hit Rate",
        "Executed Ipc Active", "No Eligible", "Eligible Warps Per Scheduler", "Warp Cycles Per Issued Instruction",
        "Avg. Active Threads Per Warp"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__grid_size"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep):
    det = list(csv.DictReader(io.StringIO(run(["-i", rep, "--page", "details", "--csv"]))))
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, body = raw[0], raw[1], raw[2:]
    ids = []
    for r in det:
        if r["ID"] not in ids:
            ids.append(r["ID"])
    for n, kid in enumerate(ids):
        rows = [r for r in det if r["ID"] == kid]
        name = rows[0]["Kernel Name"]
        print("=" * 100)
        print("launch %s: %s" % (kid, name[:90]))
        seen = set()
        for r in rows:
            key = (r["Metric Name"], r["Metric Unit"])
            if r["Metric Name"] in KEEP and key not in seen:
                seen.add(key)
                print("    %-40s %14s %s" % (r["Metric Name"], r["Metric Value"], r["Metric Unit"]))
        rr = body[n]
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                print("    %-40s %14s %s" % (m, rr[i], units[i]))
        short = name.split("(")[0].split("::")[-1].split("<")[0]
        src = run(["-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:^%s$:%d" % (short, 1)])
        lines = src.splitlines()
        try:
            rd = csv.reader(lines[1:]); h = next(rd)
            ia, isamp = h.index("Source"), h.index("# Samples")
            srows = [r for r in rd if len(r) > isamp]
            tot = sum(int(r[isamp] or 0) for r in srows) or 1
            top = sorted(srows, key=lambda r: -int(r[isamp] or 0))[:8]
            print("    top stall samples (of %d):" % tot)
            for r in top:
                print("      %5.1f%%  %s" % (100.0 * int(r[isamp]) / tot, r[ia].strip()[:100]))
        except Exception as e:  # noqa: BLE001
            print("    (no source page: %s)" % e)


if __name__ == "__main__":
    main(sys.argv[1])
