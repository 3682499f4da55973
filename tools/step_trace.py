"""Barrier-wait cycles of ONE tcgen05 launch inside the real train step (options tc_trace_ptr + tc_trace_skip = index of the launch among the
step's tcgen05 launches: 0 enc L0, 1 enc L1, 2 fused head, 3 dec L0, 4 dec L1, 5 dec out + MSE, 6.. the dgrads, last the merged wgrad).
    python tools/step_trace.py [launch indices...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseudo_speaker_vae_b200 as P
from pseudo_speaker_vae_b200 import _lib as L
from bench import synth_batch

B = 65536
torch.manual_seed(0)
m = P.PseudoSpeakerVAE(model=dict(input_dim=256, latent_dim=64), classifier=dict(input_dim=64, num_classes=2), optimizer=dict(lr=1e-3),
                       scheduler=dict(T_max=10), precision="bf16").to("cuda")
tr = P.DataParallelTrainer(m)
tr.set_shard(B)
m.hot_path.manual_seed(1, 0)
bs = []
for i in range(3):
    x, y, _ = synth_batch(B, 256, 64, 2, seed=5 + i)
    bs.append((torch.from_numpy(x).cuda().to(torch.bfloat16), torch.from_numpy(y).cuda()))
for i in range(6):
    tr.train_step(*bs[i % 3])
torch.cuda.synchronize()
names = {0: "enc L0", 1: "enc L1", 2: "fused head", 3: "dec L0", 4: "dec L1", 5: "dec out + MSE", 6: "dgrad out->hd1", 7: "dgrad hd1->hd0", 8: "dz", 9: "head dgrad", 10: "enc dgrad", 11: "merged wgrad"}
for k in ([int(a) for a in sys.argv[1:]] or list(range(12))):
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    L.set_option("tc_trace_ptr", buf.data_ptr())
    L.set_option("tc_trace_skip", k)
    tr.train_step(*bs[k % 3])
    torch.cuda.synchronize()
    L.set_option("tc_trace_ptr", 0)
    L.set_option("tc_trace_select", 0)
    t = buf.view(148, 16).double().cpu()
    if t.abs().sum() == 0:
        print(f"launch {k}: no trace written"); continue
    act = t[:, 3] > 0          # CTAs whose MMA warp ran (leaders)
    lead = t[act] if act.any() else t
    f = lambda x: f"{x.mean().item():8.0f}"
    ti = buf.view(148, 16).cpu()
    ent = ti[:, 12]; k0 = int(ent[ent > 0].min()) if (ent > 0).any() else 0
    ext = ti[:, 15]
    print(f"launch {k:2d} {names.get(k, ''):16s} span {(int(ext.max()) - k0) / 1e3:6.1f} us | MMA total {f(lead[:,3])} wait-full {f(lead[:,4])} wait-tempty {f(lead[:,5])} | "
          f"producer total {f(t[:,0])} wait-empty {f(t[:,1])} | epi w0 total {f(t[:,7])} wait-tfull {f(t[:,8])} wait-aux {f(t[:,9])}", flush=True)
