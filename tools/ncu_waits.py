import csv, subprocess, sys, io
rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", kid, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
name = [r for r in rows if r and r[0] == "Kernel Name"]
print(name[0][1][:140])
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]; si = H.index("Warp Stall Sampling (All Samples)"); ins = H.index("Instructions Executed")
data = rows[hi+1:]
tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
print("total", tot)
for i, r in enumerate(data):
    if len(r) > si and ("TRYWAIT" in r[1] or "UTCBAR" in r[1] or "UTMALDG" in r[1] or "UTCHMMA" in r[1] or "LDTM" in r[1] or "UTMASTG" in r[1] or "UTMAREDG" in r[1] or "SYNCS.ARRIVE" in r[1]):
        # include the following BRA (the spin) samples
        nxt = data[i+1] if i + 1 < len(data) else None
        extra = int(nxt[si]) if nxt and len(nxt) > si and nxt[si].isdigit() and "BRA" in nxt[1] else 0
        print(f"{int(r[si]):5d}+{extra:<5d} exec={r[ins]:>8s}  {r[1].strip()[:100]}")
