#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: VAE train samples/s (fwd+bwd+Adam) on synthetic 256-d
speaker embeddings, batch 65,536 per GPU (configs[1]; weak scaling for N > 1, configs[2]), plus the sampling
throughputs of configs[3] as secondary lines inside the same JSON object.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|fp32] [--batch B]

One JSON line on stdout (rank 0).  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same metric
through the Lightning-style public API (training_step -> loss.backward() -> optimizer.step()) with the batch coming
from pinned host memory every step and the loss read back; `roofline` = algorithmic FLOPs of the step / measured step
time against the measured bf16 peak (the WHOLE step, all 24 launches: the strict figure), with `roofline.dominant_kernel` =
the hidden-layer GEMM alone (algorithmic FLOPs per launch / its launch time, CUDA events on rotating operand sets) and
`roofline.traffic` = DRAM bytes of one step from the committed ncu pass; `cpu_baseline` = oracle/torch_port.py (the reference's own torch CPU path,
restated) on this box's host cores.  `--impl reference` times only that CPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train samples/s (fwd+bwd+Adam)"
UNIT = "samples/s"
D, LAT, HID, NH, NCLS = 256, 64, 512, 2, 2
FLOPS_PER_SAMPLE = 7_144_192          # 7,143,424 + 768 for the 2-class latent classifier


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (started before the warm-up so that the poller is
    already running; only the rows that arrive between mark_start() and mark_end() are summarised)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, 0.0, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            threading.Thread(target=self._read, daemon=True).start()
            t = time.time()
            while not self.rows and time.time() - t < 3.0:      # wait for the first row: the poller is up
                time.sleep(0.01)
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t0 <= t <= self.t1 + 0.06]
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, power_w_max=max(power) if power else None,
                    samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline (also the --impl reference arm)
# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline(budget_s: float = 20.0, steps: int = 0, warmup: int = 1, batch: int = 65536):
    """oracle/torch_port.py -- the reference's torch CPU path restated -- on all host cores, fp32 'highest'
    (the parity setting; the reference's own 'medium' turns on bf16 AMX where the CPU has it, reported alongside)."""
    import torch

    from oracle import ps_vae_oracle as O
    from oracle import torch_port as T

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {}
    for prec in ("highest", "medium"):
        torch.set_float32_matmul_precision(prec)
        torch.manual_seed(0)
        mod = T.TorchStep(D, LAT, NCLS, HID, NH)
        opt = torch.optim.Adam(mod.parameters(), lr=1e-3)
        B = batch
        x, y, _ = O.synth_batch(B, D, LAT, NCLS, seed=1234)
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        t0 = time.perf_counter()
        T.train_steps(mod, opt, xt, yt, max(1, warmup))
        t_w = (time.perf_counter() - t0) / max(1, warmup)
        n = steps if steps > 0 else max(1, min(20, int(budget_s / 2 / max(t_w, 1e-3))))
        t0 = time.perf_counter()
        T.train_steps(mod, opt, xt, yt, n)
        dt = (time.perf_counter() - t0) / n
        out[prec] = dict(samples_per_s=B / dt, ms_per_step=dt * 1e3, steps=n)
    torch.set_float32_matmul_precision("highest")
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:  # noqa: BLE001
        model = "unknown"
    return dict(value=out["highest"]["samples_per_s"], unit=UNIT, cores=cores, kind="port",
                sample=f"{out['highest']['steps']} steps of batch {batch} (fp32 'highest', torch {torch.__version__} CPU, {model}); "
                       f"'medium' (the reference's own setting): {out['medium']['samples_per_s']:.0f} samples/s",
                ms_per_step=out["highest"]["ms_per_step"], medium_value=out["medium"]["samples_per_s"])


def run_reference_arm(args):
    """--impl reference: the CPU path alone, `--steps K --warmup W` honoured.  Each step is a BOUNDED sample of the
    workload: the per-step batch is cut so that the whole run stays within ~2.5 minutes (CPU throughput is flat in the
    batch size from 4,096 rows up), and samples/s = rows actually processed / time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import ps_vae_oracle as O
    from oracle import torch_port as T

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_float32_matmul_precision("highest")
    torch.manual_seed(0)
    mod = T.TorchStep(D, LAT, NCLS, HID, NH)
    opt = torch.optim.Adam(mod.parameters(), lr=1e-3)
    K, W = max(1, args.steps), max(0, args.warmup)
    xp, yp, _ = O.synth_batch(4096, D, LAT, NCLS, seed=99)
    T.train_steps(mod, opt, torch.from_numpy(xp), torch.from_numpy(yp), 1)
    t0 = time.perf_counter()
    T.train_steps(mod, opt, torch.from_numpy(xp), torch.from_numpy(yp), 2)
    rate = 2 * 4096 / (time.perf_counter() - t0)                     # rows/s probe
    per_step_s = 150.0 / (K + W)
    Bs = int(min(args.batch, max(256, rate * per_step_s)))
    Bs = max(256, Bs // 256 * 256)
    x, y, _ = O.synth_batch(Bs, D, LAT, NCLS, seed=1234)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    if W:
        T.train_steps(mod, opt, xt, yt, W)
    t0 = time.perf_counter()
    T.train_steps(mod, opt, xt, yt, K)
    dt = (time.perf_counter() - t0) / K
    value = Bs / dt
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:  # noqa: BLE001
        model = "unknown"
    sample = (f"{K} steps (+{W} warm-up) of {Bs} rows each (bounded sample of the {args.batch}-row step), fp32 'highest', torch {torch.__version__} CPU, "
              f"{cores} threads, {model}")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"ps_vae conditional VAE train step fwd+bwd+Adam, D={D} L={LAT} hidden {HID}x{NH}, 2-class latent classifier, "
                                     f"batch {args.batch} per GPU -- CPU arm: the reference's torch path restated (oracle/torch_port.py)"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def time_events(fn, torch, dist_on):
    """barrier + sync, CUDA events on the current stream around fn(), sync + barrier; returns ms (max over ranks)."""
    import torch.distributed as dist

    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=65536, help="rows per GPU (weak scaling)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the sampling / e2e / CPU legs")
    ap.add_argument("--sample-n", type=int, default=4 * 1024 * 1024)
    ap.add_argument("--cond-n", type=int, default=1024 * 1024)
    ap.add_argument("--wide-batch", type=int, default=32768, help="rows per GPU of the widened-VAE secondary line (BASELINE configs[4])")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option name=value (psvae_set_option), repeatable")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 400 if args.impl != "reference" else 3
    if args.warmup is None:
        args.warmup = 20 if args.impl != "reference" else 1
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import pseudo_speaker_vae_b200 as P
    from oracle import ps_vae_oracle as O          # synthetic-input generator + the cpu_baseline leg only
    from pseudo_speaker_vae_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if dist_on:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and dist_on:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    n_gpus = world
    for kv in args.opt:
        k, v = kv.split("=")
        L.set_option(k, int(v))
    peaks = measured_peaks()
    B, K, W = args.batch, args.steps, max(3, args.warmup)

    torch.manual_seed(0)
    module = P.PseudoSpeakerVAE(model=dict(input_dim=D, latent_dim=LAT), classifier=dict(input_dim=LAT, num_classes=NCLS),
                                optimizer=dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0), scheduler=dict(T_max=200),
                                precision=args.precision).to(dev)
    trainer = P.DataParallelTrainer(module)
    trainer.set_shard(B * n_gpus)
    hot = module.hot_path
    hot.manual_seed(1236, 0)

    # synthetic inputs: unit-norm N(0,1) rows, labels from the CV gender marginals; NB distinct batches are rotated so
    # the 67 MB input of a step is never L2-resident from the previous one (4 x 67 MB > 126 MB L2)
    NB = 4
    host_batches = []
    for i in range(NB):
        x, y, _ = O.synth_batch(B, D, LAT, NCLS, seed=1234 + 17 * i + 1000 * rank)
        host_batches.append((torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()))
    dev_batches = [(x.to(dev), y.to(dev)) for x, y in host_batches]

    def train_steps(n):
        for i in range(n):
            x, y = dev_batches[i % NB]
            trainer.train_step(x, y)

    sampler = ClockSampler(local).start() if rank == 0 else None
    train_steps(W)
    torch.cuda.synchronize()
    l0 = L.lib().psvae_launch_count()
    if sampler:
        sampler.mark_start()
    ms = time_events(lambda: train_steps(K), torch, dist_on)
    if sampler:
        sampler.mark_end()
    launches = int(L.lib().psvae_launch_count() - l0)
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / K
    value = B * n_gpus * K / (ms * 1e-3)
    tflops = value * FLOPS_PER_SAMPLE / 1e12
    peak = peaks["bf16_sustained"] * n_gpus
    # dram__bytes_read.sum + dram__bytes_write.sum over the 24 launches of one step, from the ncu pass of this same command
    # (profiles/r01_launches_v15_time_dram.csv at B = 65,536, bf16, one GPU); null for any other config
    traffic = 2_397_704_192 if (B == 65536 and args.precision == "bf16" and not args.opt) else None
    roofline = dict(bound="tensor", achieved=tflops, peak=peak, unit="TFLOP/s", frac=tflops / peak, traffic=traffic,
                    note=f"whole fused step (all launches): {FLOPS_PER_SAMPLE} algorithmic FLOP/sample x {B * n_gpus} samples / measured step time; "
                         f"peak = sustained bf16 {peaks['source']}" + ("" if args.precision == "bf16" else " [fp32 parity mode runs on CUDA cores]"))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=K, warmup=W, ms_per_step=ms_per_step, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
                config=dict(workload=f"ps_vae conditional VAE train step fwd+bwd+Adam, D={D} L={LAT} hidden {HID}x{NH}, 2-class latent classifier, "
                                     f"batch {B} per GPU ({B * n_gpus} global), {args.precision}",
                            global_batch=B * n_gpus, parallelism=f"dp{n_gpus}", l2="4 distinct 67 MB input batches rotated (268 MB > 126 MB L2)",
                            eps="in-kernel Philox4x32-10", options={kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}, allreduce="1 bucket, flat fp32 grads 5.13 MB, NCCL" if dist_on else "none (1 GPU)"),
                clocks=clocks, gpu_launches=launches, roofline=roofline)

    if not args.no_secondary and args.precision == "bf16":
        # ---- the dominant kernel alone (hidden Linear 512 x 512 + bias + ReLU + mask: the shape of 6 of the step's GEMMs), timed live with
        #      CUDA events on rotating operand sets (3 x 67 MB in, 3 x 67 MB out: every launch streams from HBM as inside the step)
        hk = HID
        a_sets = [(torch.randn(B, hk, device=dev) * 0.1).to(torch.bfloat16) for _ in range(3)]
        o_sets = [torch.empty(B, hk, device=dev, dtype=torch.bfloat16) for _ in range(3)]
        wk = (torch.randn(hk, hk, device=dev) * 0.05).to(torch.bfloat16)
        bk = torch.zeros(hk, device=dev)
        mk = torch.empty(hk // 32 * B, device=dev, dtype=torch.int32)
        stp = torch.cuda.current_stream().cuda_stream

        def probe(n):
            for i in range(n):
                L.check(L.lib().psvae_gemm_probe(a_sets[i % 3].data_ptr(), wk.data_ptr(), bk.data_ptr(), o_sets[i % 3].data_ptr(), mk.data_ptr(), None,
                                                 B, hk, hk, 0, stp))
        probe(6)
        torch.cuda.synchronize()
        us_k = time_events(lambda: probe(30), torch, False) * 1e3 / 30
        fl_k = 2.0 * B * hk * hk
        roofline["dominant_kernel"] = dict(name="gemm_tc_kernel<256, K-major, K-major, EpiBiasAct<bf16, relu>, cta_group::2> (hidden Linear 512x512)",
                                           flops_per_launch=fl_k, us_per_launch=us_k, achieved=fl_k / us_k / 1e6, unit="TFLOP/s",
                                           frac=fl_k / us_k / 1e6 / peaks["bf16_sustained"], frac_of_burst_peak=fl_k / us_k / 1e6 / peaks["bf16_burst"])
        del a_sets, o_sets

    if not args.no_secondary:
        # ---- e2e: the public Lightning-style API with host inputs every step -----------------------------------------
        opt = trainer.optimizer
        copy_stream = torch.cuda.Stream(dev)
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def e2e_steps(n):
            nxt = None
            with torch.cuda.stream(copy_stream):
                nxt = (host_batches[0][0].to(dev, non_blocking=True), host_batches[0][1].to(dev, non_blocking=True))
                ev = torch.cuda.Event(); ev.record(copy_stream)
            for i in range(n):
                torch.cuda.current_stream().wait_event(ev)
                x, y = nxt
                if i + 1 < n:                      # prefetch the next batch while this one computes
                    hx, hy = host_batches[(i + 1) % NB]
                    with torch.cuda.stream(copy_stream):
                        nxt = (hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True))
                        ev = torch.cuda.Event(); ev.record(copy_stream)
                opt.zero_grad()
                loss = module.training_step((x, y), i)["loss"]
                loss.backward()
                if dist_on:
                    g = hot.arena.flat_grad()
                    P.parallel.all_reduce_flat(g)
                opt.step()
                x.record_stream(torch.cuda.current_stream()); y.record_stream(torch.cuda.current_stream())
                loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e2e_steps(2 * NB + 2)      # touch every pinned host batch and let the copy path warm up (the first H2D copies of a process are slow)
        Ke = max(5, min(K, 100))
        ms_e = time_events(lambda: e2e_steps(Ke), torch, dist_on)
        e2e_val = B * n_gpus * Ke / (ms_e * 1e-3)
        line["e2e"] = dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=int(B * D * 4 + B * 8), d2h_bytes_per_step=4, ms_per_step=ms_e / Ke,
                           api="PseudoSpeakerVAE.training_step -> loss.backward() -> FusedAdam.step(), batch from pinned host memory each step "
                               "(prefetched on a copy stream), loss read back to the host")

        # ---- secondary: sampling (BASELINE configs[3]) ---------------------------------------------------------------
        sec = {}
        try:
            Ns = args.sample_n
            out = torch.empty(Ns, D, dtype=torch.float32, device=dev)
            row0 = rank * Ns
            for _ in range(2):
                P.sample_on_device(module, Ns, out=out, row0=row0)
            reps = 5
            ms_s = time_events(lambda: [P.sample_on_device(module, Ns, out=out, row0=row0) for _ in range(reps)], torch, dist_on)
            sps = Ns * n_gpus * reps / (ms_s * 1e-3)
            sec["unconditional_sampling"] = dict(value=sps, unit="embeddings/s", n_per_gpu=Ns, ms=ms_s / reps,
                                                 tensor_frac=sps * 851968 / 1e12 / (peaks["bf16_sustained"] * n_gpus),
                                                 hbm_frac=sps * D * 4 / 1e9 / (peaks["hbm_gbs"] * n_gpus),
                                                 note="z ~ Philox in-kernel -> decoder chain -> fp32 [N,256] left in HBM; 851,968 FLOP + 1,024 B per sample")
            Nc = args.cond_n
            outc = out[:Nc]
            P.sample_on_device(module, Nc, classifier_target=1, num_steps=100, out=outc, row0=rank * Nc)
            ms_c = time_events(lambda: P.sample_on_device(module, Nc, classifier_target=1, num_steps=100, out=outc, row0=rank * Nc), torch, dist_on)
            cps = Nc * n_gpus / (ms_c * 1e-3)
            sec["conditional_sampling"] = dict(value=cps, unit="embeddings/s", n_per_gpu=Nc, num_steps=100, ms=ms_c,
                                               normals_per_s=cps * (100 * 64 + 64),
                                               note="100 Langevin steps (1 launch) + decode; bound by Philox/Box-Muller on the CUDA cores")
            del out
        except Exception as e:  # noqa: BLE001
            sec["error"] = repr(e)
        # ---- secondary: the widened VAE of BASELINE configs[4] (512-d, 4 x 2048 hidden, latent classifier) ------------------------
        try:
            Dw, Hw, NHw, Bw = 512, 2048, 4, args.wide_batch
            torch.manual_seed(0)
            wide = P.PseudoSpeakerVAE(model=dict(input_dim=Dw, latent_dim=LAT, hidden_dim=Hw, num_hidden_layers=NHw),
                                      classifier=dict(input_dim=LAT, num_classes=NCLS), optimizer=dict(lr=1e-3), scheduler=dict(T_max=200),
                                      precision=args.precision).to(dev)
            wtr = P.DataParallelTrainer(wide)
            wtr.set_shard(Bw * n_gpus)
            wide.hot_path.manual_seed(1236, 0)
            xw, yw, _ = O.synth_batch(Bw, Dw, LAT, NCLS, seed=4321 + rank)
            wb = [(torch.from_numpy(xw).to(dev), torch.from_numpy(yw).to(dev)), (torch.from_numpy(xw[::-1].copy()).to(dev), torch.from_numpy(yw[::-1].copy()).to(dev))]
            for i in range(3):
                wtr.train_step(*wb[i % 2])
            Kw = 10
            ms_w = time_events(lambda: [wtr.train_step(*wb[i % 2]) for i in range(Kw)], torch, dist_on)
            fl = float(wide.hot_path.flops_train)            # psvae_flops_per_sample: 243,532,544 (+768 classifier), SURVEY 8(d)
            wps = Bw * n_gpus * Kw / (ms_w * 1e-3)
            sec["widened_config5"] = dict(value=wps, unit=UNIT, batch_per_gpu=Bw, ms_per_step=ms_w / Kw, flops_per_sample=fl,
                                          tensor_frac=wps * fl / 1e12 / (peaks["bf16_sustained"] * n_gpus),
                                          note="D=512, 4 x 2048 hidden, L=64, 2-class latent classifier, fwd+bwd+Adam, same fused step")
            del wide, wtr, wb
        except Exception as e:  # noqa: BLE001
            sec["widened_config5"] = dict(error=repr(e))
        line["secondary"] = sec

        # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) -----------------------------------------------
        if rank == 0 and n_gpus == 1:
            try:
                base = cpu_baseline(budget_s=20.0, batch=B)
                line["cpu_baseline"] = dict(value=base["value"], unit=UNIT, cores=base["cores"], kind=base["kind"], sample=base["sample"])
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample=f"failed: {e!r}")
    if dist_on:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
