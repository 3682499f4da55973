"""Per-source-line stall samples of one kernel: python tools/ncu_lines.py rep launch_index [n]"""
import csv, subprocess, sys, io, collections
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", kid, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; agg = collections.Counter(); src = {}
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": print(r[1][:150]); continue
    if r[0] == "Line No": hdr = r; si = r.index("Warp Stall Sampling (All Samples)"); continue
    if hdr and r[0].isdigit() and len(r) > si and r[si].isdigit():
        agg[(cur, int(r[0]))] += int(r[si]); src[(cur, int(r[0]))] = r[1].strip()[:120]
tot = sum(agg.values())
print("total samples", tot)
for k, v in agg.most_common(n):
    print(f"{100*v/tot:5.1f}%  {k[0]}:{k[1]:<5d} {src[k]}")
