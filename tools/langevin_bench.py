"""Conditional sampler alone: N samples x num_steps Langevin steps (single-label linear head), embeddings/s and normals/s."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseudo_speaker_vae_b200 as P

N, STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20, 100
torch.manual_seed(0)
m = P.PseudoSpeakerVAE(model=dict(input_dim=256, latent_dim=64), classifier=dict(input_dim=64, num_classes=2), optimizer={}, scheduler=dict(T_max=1),
                       precision="bf16").to("cuda")
hot = m.hot_path
for _ in range(2):
    hot.langevin(N, 1, 0.01, STEPS, 1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    hot.langevin(N, 1, 0.01, STEPS, 1.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"langevin {N} x {STEPS} steps: {ms:.3f} ms  {N / ms / 1e3:.1f} M samples/s  {N * 64 * (STEPS + 1) / ms / 1e6:.1f} G normals/s")
