"""Is the K = 512 GEMM bound chip-wide (L2 -> SM fabric / HBM) or per SM (latency)?  Time the same launch on fewer SMs: a chip-wide bound
shows up as better-than-proportional per-SM throughput when fewer SMs compete."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import epi_probe as E

for form, name in ((0, "fwd 512x512"), (1, "dgrad 512x512")):
    base = None
    for g in (148, 112, 74, 36):
        us, tf = E.run(65536, 512, 512, form, opts={"tc_grid_limit": g})
        base = base or us * 148
        print(f"{name} grid {g:3d} SMs: {us:7.1f} us  {tf:7.1f} TFLOP/s  per-SM {tf / g:6.2f} TFLOP/s  (SM-us {us * g:8.0f}, x{us * g / base:4.2f} of full grid)", flush=True)
us, tf = E.run(65536, 512, 512, 0, store=False, mask=False, opts={"tc_grid_limit": 74})
print(f"fwd 512x512 no store, 74 SMs: {us:7.1f} us {tf:7.1f} TFLOP/s per-SM {tf/74:6.2f}")
us, tf = E.run(65536, 512, 512, 0, store=False, mask=False)
print(f"fwd 512x512 no store, 148 SMs: {us:7.1f} us {tf:7.1f} TFLOP/s per-SM {tf/148:6.2f}")
