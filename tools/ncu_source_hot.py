#!/usr/bin/env python
"""Hot source lines of one kernel launch in an .ncu-rep (needs -lineinfo + --import-source on):
   python tools/ncu_source_hot.py REPORT.ncu-rep KERNEL_REGEX LAUNCH_SKIP [min_pct]
Prints, per CUDA source line, its share of executed warp instructions and of warp-stall samples, plus the top stall reasons."""
import csv
import subprocess
import sys


def main(rep, regex, skip, min_pct=0.7):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + regex,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, agg = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            print("kernel:", r[1][:120])
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) >= len(hdr) and r[2] == "-":
            ix = {h: i for i, h in enumerate(hdr)}
            ie = float(r[ix["Instructions Executed"]] or 0)
            sm = float(r[ix["# Samples"]] or 0)
            stalls = {h[6:]: float(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            agg.append((cur_file, r[0], r[1].strip(), ie, sm, stalls))
    ti = sum(a[3] for a in agg) or 1
    ts = sum(a[4] for a in agg) or 1
    print("warp instructions %.3g, stall samples %d" % (ti, ts))
    tot_st = {}
    for a in agg:
        for k, v in a[5].items():
            tot_st[k] = tot_st.get(k, 0) + v
    print("stall mix:", ", ".join("%s %.0f%%" % (k, 100 * v / ts) for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))
    for a in agg:
        pi, ps = 100 * a[3] / ti, 100 * a[4] / ts
        if pi >= min_pct or ps >= min_pct:
            top = sorted(a[5].items(), key=lambda kv: -kv[1])[:2]
            print("%s:%4s inst %5.1f%% stall %5.1f%% [%s] | %s" % (a[0][6:], a[1], pi, ps, ",".join("%s %d" % (k, v) for k, v in top if v), a[2][:90]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4]) if len(sys.argv) > 4 else 0.7)
