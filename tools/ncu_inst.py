"""Instructions executed per source line of one kernel in an ncu report (warp-level counts):  python tools/ncu_inst.py rep launch_index [n]"""
import csv, subprocess, sys, io, collections
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", kid, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; agg = collections.Counter(); st = collections.Counter(); src = {}
hdr = None; seen = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if r[1] not in seen: print(r[1][:150])
        seen.add(r[1]); continue
    if r[0] == "Line No": hdr = r; ii = r.index("Instructions Executed"); si = r.index("Warp Stall Sampling (All Samples)"); continue
    if hdr and r[0].isdigit() and len(r) > ii and r[ii].isdigit():
        agg[(cur, int(r[0]))] += int(r[ii]); st[(cur, int(r[0]))] += int(r[si]) if r[si].isdigit() else 0; src[(cur, int(r[0]))] = r[1].strip()[:110]
tot = sum(agg.values()); stot = sum(st.values())
print("total warp instructions", tot, "stall samples", stot)
for k, v in agg.most_common(n):
    print(f"{100*v/tot:5.1f}% inst {100*st[k]/max(stot,1):5.1f}% stall  {k[0]}:{k[1]:<5d} {src[k]}")
