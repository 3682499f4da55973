"""Bare generator throughput (normals/s): the ceiling the conditional sampler is measured against."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pseudo_speaker_vae_b200 import _lib as L
n_rows, L_ = 1 << 22, 64
out = torch.empty(n_rows, L_, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    L.check(L.lib().psvae_philox_normal(out.data_ptr(), n_rows, L_, 1, 0, 0, st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    L.check(L.lib().psvae_philox_normal(out.data_ptr(), n_rows, L_, 1, i, 0, st))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"philox_normal: {n_rows * L_ / ms / 1e6:.1f} G normals/s ({ms:.3f} ms for {n_rows * L_ / 1e6:.0f} M; write {n_rows * L_ * 4 / ms / 1e6:.0f} GB/s)")
