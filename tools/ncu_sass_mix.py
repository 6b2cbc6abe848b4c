"""Opcode mix (share of executed warp instructions) of one tile_kernel launch in an .ncu-rep.
   python tools/ncu_sass_mix.py REPORT.ncu-rep LAUNCH_SKIP"""
import csv, subprocess, sys, collections, re
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","sass","--csv","--kernel-name","regex:tile_kernel","--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None; mix=collections.Counter(); tot=0
for r in rows:
    if r and r[0]=="Address": hdr={h:i for i,h in enumerate(r)}; continue
    if hdr and len(r)>=len(hdr):
        src=r[hdr["Source"]].strip(); ie=float(r[hdr["Instructions Executed"]] or 0)
        m=re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        if not m: continue
        op=m.group(2).split('.')[0]
        mix[op]+=ie; tot+=ie
print("total %.4g"%tot)
for op,v in mix.most_common(28): print("%-10s %5.1f%%"%(op,100*v/tot))
