"""Top stall locations of one kernel in an ncu report (SASS view): python tools/ncu_src.py rep.ncu-rep <launch-id> [n]"""
import csv, subprocess, sys, io
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", kid, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
name = [r for r in rows if r and r[0] == "Kernel Name"]
print(name[0][1][:160] if name else "?")
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
si = H.index("Warp Stall Sampling (All Samples)")
ins = H.index("Instructions Executed")
data = [(int(r[si] or 0), i, r) for i, r in enumerate(rows[hi + 1:]) if len(r) > si and r[si].isdigit()]
tot = sum(d[0] for d in data)
print("total samples", tot)
stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") or "Stall" in h and "Sampling" not in h]
for s, i, r in sorted(data, reverse=True)[:n]:
    # context: the instruction text
    extra = ""
    print(f"{100*s/tot:5.1f}%  #{i:5d}  exec={r[ins]:>8s}  {r[1].strip()[:110]}")
