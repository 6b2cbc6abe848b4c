"""Hottest basic blocks of one tile_kernel launch in an .ncu-rep (SASS page): contiguous runs of instructions with the same executed
count, with their opcode histogram.   python tools/ncu_sass_blocks.py REPORT.ncu-rep LAUNCH_SKIP [n_blocks]"""
import csv, subprocess, sys, collections, re
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","sass","--csv","--kernel-name","regex:tile_kernel","--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None; seen=set(); L=[]
for r in rows:
    if r and r[0]=="Address": hdr={h:i for i,h in enumerate(r)}; continue
    if hdr and len(r)>=len(hdr):
        if r[0] in seen: continue
        seen.add(r[0])
        L.append((r[0], r[hdr["Source"]].strip(), float(r[hdr["Instructions Executed"]] or 0), float(r[hdr["# Samples"]] or 0)))
tot=sum(x[2] for x in L)
# contiguous runs of instructions with the same executed count = basic blocks of a loop
runs=[]; cur=None
for a,s,ie,sm in L:
    if cur and abs(cur[0]-ie)<1e-9: cur[1].append((a,s)); cur[2]+=sm
    else:
        if cur: runs.append(cur)
        cur=[ie,[(a,s)],sm]
runs.append(cur)
runs.sort(key=lambda r:-r[0]*len(r[1]))
print("total %.4g, %d instructions"%(tot,len(L)))
for ie,ins,sm in runs[:int(sys.argv[3]) if len(sys.argv)>3 else 14]:
    ops=collections.Counter(re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)",s).group(2) for a,s in ins if re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)",s))
    print("%5.2f%%  n=%3d x %.3g  @%s  %s"%(100*ie*len(ins)/tot,len(ins),ie,ins[0][0][-5:],dict(ops.most_common(7))))
