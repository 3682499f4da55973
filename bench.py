#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: VAE train samples/s (fwd+bwd+Adam) on synthetic 256-d
speaker embeddings, batch 65,536 per GPU (configs[1]; weak scaling for N > 1, configs[2]), plus the sampling
throughputs of configs[3] and the widened VAE of configs[4] as secondary lines inside the same JSON object.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|fp32] [--batch B] [--x-dtype bf16|fp32]

One JSON line on stdout (rank 0).
  value      whole-job samples/s with the batches resident in HBM (bf16 rows by default: what a bf16 PackedEmbeddingStore holds;
             `secondary.fp32_input` is the same step fed fp32 rows, which adds the cast pass);
  e2e        the same metric through the public API -- PinnedBatchLoader over a pinned bf16 store -> PseudoSpeakerVAE.training_step ->
             loss.backward() -> FusedAdam.step() -- with every step's batch DMA'd from pinned host memory inside the timed region and
             the loss read back; `secondary.e2e_resident_store` = the store resident in HBM, only the index list crossing PCIe;
  roofline   algorithmic FLOPs of the step / measured step time against the measured bf16 peak: `peak` is the burst figure when the
             timed region is shorter than a second (no power capping yet) and the sustained one otherwise, both fractions are printed;
             `roofline.dominant_kernel` = the hidden-layer GEMM alone (CUDA events on rotating operand sets);
             `roofline.traffic` = DRAM bytes of one step from the committed ncu pass of THIS build (profiles/r02_traffic.json, produced by
             tools/traffic_from_ncu.py; null when the kernel sources changed since);
  cpu_baseline   oracle/torch_port.py (the reference's own torch CPU path, restated) on this box's host cores, B = 65,536 and the
             B = 256 configuration of configs[0];  `secondary.torch_gpu_reference` = the same torch module on this B200 through stock
             torch (cuBLAS + ATen + foreach Adam) in fp32 'highest', 'medium' and bf16 autocast -- the library path this repo replaces;
  dp_parity  (N > 1) pseudo_speaker_vae_b200.parallel.verify_data_parallel_step on the running ranks: bit-identical parameters across ranks
             and <= 1e-5 against a single-process run on the global batch, asserted.
`--impl reference` times only the CPU path.
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


def synth_batch(B: int, Dm: int, Lm: int, num_classes: int = 2, seed: int = 1234):
    """SURVEY 8(d) synthetic inputs: x ~ N(0,1) rows L2-normalised (speaker embeddings are ~unit norm), y from the Common Voice gender
    marginals (plots/dataset_info_train.json:161-176 through utils.py:92-119), eps ~ N(0,1); numpy PCG64."""
    import numpy as np

    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, Dm))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    probs = {2: [0.717, 0.283], 3: [0.694, 0.274, 0.032]}
    y = rng.choice(num_classes, size=B, p=probs.get(num_classes)).astype(np.int64)
    eps = rng.standard_normal((B, Lm))
    return x.astype(np.float32), y, eps.astype(np.float32)


def sources_sha():
    """sha256 over the kernel sources: ties profiles/r02_traffic.json (an ncu pass) to the build that is running."""
    import hashlib

    h = hashlib.sha256()
    base = os.path.join(ROOT, "pseudo_speaker_vae_b200", "csrc")
    for f in sorted(os.listdir(base)):
        with open(os.path.join(base, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def measured_traffic(batch: int, precision: str, x_dtype: str, opts) -> tuple:
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if opts or not os.path.isfile(path):
        return None, "no ncu pass on record for this configuration"
    with open(path) as f:
        rec = json.load(f)
    if rec.get("batch") != batch or rec.get("precision") != precision or rec.get("x_dtype") != x_dtype:
        return None, "the ncu pass on record is for another configuration"
    if rec.get("sources_sha") != sources_sha():
        return None, f"kernel sources changed since the ncu pass on record ({rec.get('sources_sha')})"
    return int(rec["traffic_bytes_per_step"]), f"dram__bytes_read.sum + dram__bytes_write.sum over the {rec.get('launches_per_step')} launches of one step ({rec.get('source')})"


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
def _cpu_model_name():
    try:
        return [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:  # noqa: BLE001
        return "unknown"


def cpu_baseline(budget_s: float = 20.0, steps: int = 0, warmup: int = 1, batch: int = 65536):
    """oracle/torch_port.py -- the reference's torch CPU path restated -- on all host cores, fp32 'highest'
    (the parity setting; the reference's own 'medium' turns on bf16 AMX where the CPU has it, reported alongside), at the bench batch
    and at BASELINE configs[0]'s batch 256."""
    import torch

    from oracle import torch_port as T

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {}
    for prec in ("highest", "medium"):
        torch.set_float32_matmul_precision(prec)
        for B, share in ((batch, 0.4), (256, 0.1)):
            torch.manual_seed(0)
            mod = T.TorchStep(D, LAT, NCLS, HID, NH)
            opt = torch.optim.Adam(mod.parameters(), lr=1e-3)
            x, y, _ = synth_batch(B, D, LAT, NCLS, seed=1234)
            xt, yt = torch.from_numpy(x), torch.from_numpy(y)
            t0 = time.perf_counter()
            T.train_steps(mod, opt, xt, yt, max(1, warmup if B > 256 else 20))
            t_w = (time.perf_counter() - t0) / max(1, warmup if B > 256 else 20)
            n = steps if (steps > 0 and B > 256) else max(1, min(20 if B > 256 else 100, int(budget_s * share / max(t_w, 1e-4))))
            t0 = time.perf_counter()
            T.train_steps(mod, opt, xt, yt, n)
            dt = (time.perf_counter() - t0) / n
            out[(prec, B)] = dict(samples_per_s=B / dt, ms_per_step=dt * 1e3, steps=n)
    torch.set_float32_matmul_precision("highest")
    model = _cpu_model_name()
    hi, med = out[("highest", batch)], out[("medium", batch)]
    return dict(value=hi["samples_per_s"], unit=UNIT, cores=cores, kind="port",
                sample=f"{hi['steps']} steps of batch {batch} (fp32 'highest', torch {torch.__version__} CPU, {model}); "
                       f"'medium' (the reference's own setting): {med['samples_per_s']:.0f} samples/s",
                ms_per_step=hi["ms_per_step"], medium_value=med["samples_per_s"],
                config1_batch256=dict(highest=dict(value=out[("highest", 256)]["samples_per_s"], ms_per_step=out[("highest", 256)]["ms_per_step"],
                                                   steps=out[("highest", 256)]["steps"]),
                                      medium=dict(value=out[("medium", 256)]["samples_per_s"], ms_per_step=out[("medium", 256)]["ms_per_step"],
                                                  steps=out[("medium", 256)]["steps"]),
                                      unit=UNIT, note="BASELINE configs[0]: batch 256, fp32, CPU, all host threads"))


def torch_gpu_baseline(dev, batch: int, peaks):
    """The reference module on THIS B200 through stock torch -- nn.Linear / autograd / torch.optim.Adam(foreach), i.e. cuBLASLt + ATen -- in the
    three numeric settings a user of the reference could pick ('the kernel to beat').  Baseline leg: the torch
    restatement under oracle/ is what is being timed here, never the product."""
    import torch

    from oracle import torch_port as T

    x, y, _ = synth_batch(batch, D, LAT, NCLS, seed=1234)
    xs = [torch.from_numpy(x).to(dev), torch.from_numpy(x[::-1].copy()).to(dev)]
    ys = [torch.from_numpy(y).to(dev), torch.from_numpy(y[::-1].copy()).to(dev)]
    out = {}
    for name in ("fp32_highest", "fp32_medium", "bf16_autocast"):
        torch.set_float32_matmul_precision("medium" if name == "fp32_medium" else "highest")
        torch.manual_seed(0)
        mod = T.TorchStep(D, LAT, NCLS, HID, NH).to(dev)
        opt = torch.optim.Adam(mod.parameters(), lr=1e-3, foreach=True)

        def steps(n):
            for i in range(n):
                opt.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(name == "bf16_autocast")):
                    loss = mod.loss(xs[i % 2], ys[i % 2])
                loss.backward()
                opt.step()

        steps(5)
        torch.cuda.synchronize()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps(n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        v = batch / (ms * 1e-3)
        out[name] = dict(value=v, unit=UNIT, ms_per_step=ms, tensor_frac_of_sustained=v * FLOPS_PER_SAMPLE / 1e12 / peaks["bf16_sustained"])
        del mod, opt
    torch.set_float32_matmul_precision("highest")
    out["note"] = (f"oracle/torch_port.py (the reference's torch path restated) on this GPU, batch {batch}, stock torch {torch.__version__}: cuBLASLt GEMMs, ATen "
                   "element-wise kernels, foreach Adam, eager mode as the reference runs it (ps_vae/training.py:18 sets 'medium')")
    return out


def run_reference_arm(args):
    """--impl reference: the CPU path alone, `--steps K --warmup W` honoured.  Each step is a BOUNDED sample of the
    workload: the per-step batch is cut so that the whole run stays within ~2.5 minutes (CPU throughput is flat in the
    batch size from 4,096 rows up), and samples/s = rows actually processed / time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import torch_port as T

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_float32_matmul_precision("highest")
    torch.manual_seed(0)
    mod = T.TorchStep(D, LAT, NCLS, HID, NH)
    opt = torch.optim.Adam(mod.parameters(), lr=1e-3)
    K, W = max(1, args.steps), max(0, args.warmup)
    xp, yp, _ = synth_batch(4096, D, LAT, NCLS, seed=99)
    T.train_steps(mod, opt, torch.from_numpy(xp), torch.from_numpy(yp), 1)
    t0 = time.perf_counter()
    T.train_steps(mod, opt, torch.from_numpy(xp), torch.from_numpy(yp), 2)
    rate = 2 * 4096 / (time.perf_counter() - t0)                     # rows/s probe
    per_step_s = 150.0 / (K + W)
    Bs = int(min(args.batch, max(256, rate * per_step_s)))
    Bs = max(256, Bs // 256 * 256)
    x, y, _ = synth_batch(Bs, D, LAT, NCLS, seed=1234)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    if W:
        T.train_steps(mod, opt, xt, yt, W)
    t0 = time.perf_counter()
    T.train_steps(mod, opt, xt, yt, K)
    dt = (time.perf_counter() - t0) / K
    value = Bs / dt
    model = _cpu_model_name()
    sample = (f"{K} steps (+{W} warm-up) of {Bs} rows each (bounded sample of the {args.batch}-row step), fp32 'highest', torch {torch.__version__} CPU, "
              f"{cores} threads, {model}")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload_string(args.batch, max(1, int(os.environ.get("WORLD_SIZE", "1")))),
                            arm="CPU arm: the reference's torch path restated (oracle/torch_port.py), rank 0 only"),
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


def workload_string(batch: int, n_gpus: int) -> str:
    return (f"ps_vae conditional VAE train step fwd+bwd+Adam, D={D} L={LAT} hidden {HID}x{NH}, 2-class latent classifier, "
            f"batch {batch} per GPU ({batch * n_gpus} global)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--x-dtype", default=None, choices=["bf16", "fp32"], help="element type of the resident input batches (default: bf16 in bf16 mode)")
    ap.add_argument("--batch", type=int, default=65536, help="rows per GPU (weak scaling)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the sampling / e2e / CPU legs")
    ap.add_argument("--sample-n", type=int, default=12_500_000, help="unconditional samples per GPU (BASELINE configs[3]: 100 M over 8 GPUs)")
    ap.add_argument("--cond-n", type=int, default=12_500_000, help="conditional samples per GPU (100 Langevin steps)")
    ap.add_argument("--wide-batch", type=int, default=65536, help="rows per GPU of the widened-VAE secondary line (BASELINE configs[4])")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option name=value (psvae_set_option), repeatable")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 400 if args.impl != "reference" else 3
    if args.warmup is None:
        args.warmup = 20 if args.impl != "reference" else 1
    if args.impl == "reference":
        return run_reference_arm(args)

    import tempfile

    import numpy as np
    import torch
    import torch.distributed as dist

    import pseudo_speaker_vae_b200 as P
    from pseudo_speaker_vae_b200 import _lib as L
    from pseudo_speaker_vae_b200 import data as PD

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
    x_dtype = args.x_dtype or ("bf16" if args.precision == "bf16" else "fp32")
    if args.precision == "fp32":
        x_dtype = "fp32"
    tdt = torch.bfloat16 if x_dtype == "bf16" else torch.float32

    torch.manual_seed(0)
    module = P.PseudoSpeakerVAE(model=dict(input_dim=D, latent_dim=LAT), classifier=dict(input_dim=LAT, num_classes=NCLS),
                                optimizer=dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0), scheduler=dict(T_max=200),
                                precision=args.precision).to(dev)
    trainer = P.DataParallelTrainer(module)
    trainer.set_shard(B * n_gpus)
    hot = module.hot_path
    hot.manual_seed(1236, 0)

    # synthetic inputs: unit-norm N(0,1) rows, labels from the CV gender marginals; NB distinct batches are rotated so that
    # the input of a step is never L2-resident from the previous one (4 x 33.5 MB bf16 / 4 x 67 MB fp32 > 126 MB L2 together with the
    # 1 GB of activations every step streams through it)
    NB = 4
    xs_np, ys_np = [], []
    for i in range(NB):
        x, y, _ = synth_batch(B, D, LAT, NCLS, seed=1234 + 17 * i + 1000 * rank)
        xs_np.append(x)
        ys_np.append(y)
    dev_batches = [(torch.from_numpy(x).to(dev).to(tdt), torch.from_numpy(y).to(dev)) for x, y in zip(xs_np, ys_np)]

    def train_steps(n, batches=dev_batches):
        for i in range(n):
            x, y = batches[i % NB]
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
    # a timed region shorter than a second runs at boost clocks (no power capping yet): the burst peak is the fair denominator there
    use_burst = ms < 1000.0
    peak = (peaks["bf16_burst"] if use_burst else peaks["bf16_sustained"]) * n_gpus
    traffic, traffic_note = measured_traffic(B, args.precision, x_dtype, args.opt)
    roofline = dict(bound="tensor", achieved=tflops, peak=peak, unit="TFLOP/s", frac=tflops / peak, traffic=traffic,
                    frac_of_burst_peak=tflops / (peaks["bf16_burst"] * n_gpus), frac_of_sustained_peak=tflops / (peaks["bf16_sustained"] * n_gpus),
                    peak_kind="burst" if use_burst else "sustained", traffic_note=traffic_note,
                    note=f"whole fused step (all launches): {FLOPS_PER_SAMPLE} algorithmic FLOP/sample x {B * n_gpus} samples / measured step time; "
                         f"peak = {'burst' if use_burst else 'sustained'} bf16 {peaks['source']} (timed region {ms:.0f} ms)"
                         + ("" if args.precision == "bf16" else " [fp32 parity mode runs on CUDA cores]"))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=K, warmup=W, ms_per_step=ms_per_step, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
                config=dict(workload=workload_string(B, n_gpus), precision=args.precision, x_dtype=x_dtype,
                            global_batch=B * n_gpus, parallelism=f"dp{n_gpus}",
                            l2=f"{NB} distinct input batches rotated; every step streams > 1 GB of activations through the 126 MB L2",
                            eps="in-kernel Philox4x32-10", options={kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt},
                            allreduce=(f"1 bucket, flat fp32 grads 5.13 MB, {trainer.allreduce_kind}"
                                       + (f" (timed at start-up, us: {getattr(trainer, 'allreduce_timings', {})})" if hasattr(trainer, "allreduce_timings") else ""))
                            if dist_on else "none (1 GPU)"),
                clocks=clocks, gpu_launches=launches, roofline=roofline)

    if dist_on:
        # the data-parallel step checked numerically on the ranks that are running (ps_vae/training.py:78 DDPStrategy semantics)
        par = P.parallel.verify_data_parallel_step(dev, steps=3, global_batch=1024 * n_gpus if n_gpus > 8 else 8192, precision="fp32")
        line["dp_parity"] = par
        if not (par["ranks_identical"] and par["param_rel_err"] <= 1e-5 and par["grad_rel_err"] <= 1e-5):
            raise AssertionError(f"data-parallel parity check failed: {par}")

    sec = {}
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
                                           frac=fl_k / us_k / 1e6 / peaks["bf16_burst"], frac_of_burst_peak=fl_k / us_k / 1e6 / peaks["bf16_burst"],
                                           frac_of_sustained_peak=fl_k / us_k / 1e6 / peaks["bf16_sustained"], peak_kind="burst (kernel timed alone)")
        del a_sets, o_sets
        if x_dtype == "bf16":
            # the same step fed fp32 rows (what the reference's DataLoader yields): + the fp32 -> bf16 cast pass and the fp32 MSE target
            f32_batches = [(torch.from_numpy(x).to(dev), yb) for x, (_, yb) in zip(xs_np, dev_batches)]
            train_steps(4, f32_batches)
            Kf = max(20, min(K, 100))
            ms_f = time_events(lambda: train_steps(Kf, f32_batches), torch, dist_on)
            sec["fp32_input"] = dict(value=B * n_gpus * Kf / (ms_f * 1e-3), unit=UNIT, ms_per_step=ms_f / Kf,
                                     note="fp32 [B,256] batches resident in HBM: cast_bf16_kernel + fp32 target tiles in the MSE epilogue")
            del f32_batches

    if not args.no_secondary:
        # ---- e2e: the public API end to end.  A bf16 PackedEmbeddingStore (pinned) -> PinnedBatchLoader -> training_step -> backward -> FusedAdam.step();
        #      every batch is DMA'd from pinned host memory inside the timed region, the loss is read back every step
        opt = trainer.optimizer
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        tmp = tempfile.mkdtemp(prefix=f"psvae_bench_r{rank}_")
        store = PD.PackedEmbeddingStore.from_arrays(tmp, np.concatenate(xs_np), np.concatenate(ys_np), ["gender"], dtype=x_dtype if args.precision == "bf16" else "f32")

        def e2e_steps(n, loader, lightning):
            done = 0
            while done < n:
                for x, y in loader:
                    if lightning:          # what Lightning's automatic optimisation runs per batch
                        opt.zero_grad()
                        loss = module.training_step((x, y), done)["loss"]
                        loss.backward()
                        if dist_on:
                            P.parallel.all_reduce_flat(hot.arena.flat_grad())
                        opt.step()
                        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
                    else:                  # the repo's own fit-loop body: fused step -> all-reduce -> fused Adam, three device-side calls
                        losses = trainer.train_step(x, y)
                        loss_host.copy_(losses[:1], non_blocking=True)
                    done += 1
                    if done >= n:
                        break
            torch.cuda.current_stream().synchronize()

        Ke = max(5, min(K, 100))
        legs = (("e2e", False, False), ("e2e_resident_store", True, False), ("e2e_lightning_api", False, True))
        for name, resident, lightning in legs:
            try:
                if not resident:
                    store.pin()
                loader = PD.PinnedBatchLoader(store, B, device=dev, shuffle=resident, seed=3, resident=resident)
                e2e_steps(2 * NB + 2, loader, lightning)       # touch every pinned page and let the copy path warm up (the first H2D copies of a process are slow)
                ms_e = time_events(lambda: e2e_steps(Ke, loader, lightning), torch, dist_on)
                rec = dict(value=B * n_gpus * Ke / (ms_e * 1e-3), unit=UNIT, h2d_bytes_per_step=int(loader.h2d_bytes_per_batch), d2h_bytes_per_step=4,
                           ms_per_step=ms_e / Ke)
                src = (f"PinnedBatchLoader over a pinned {x_dtype} PackedEmbeddingStore (every step's batch DMA'd from pinned host memory on a copy stream, one batch ahead)"
                       if not resident else
                       "PinnedBatchLoader(resident=True): the packed store lives in HBM (uploaded once), each step's shuffled index list is copied from pinned host "
                       "memory and the batch is assembled on the device (psvae_gather_rows)")
                body = ("PseudoSpeakerVAE.training_step -> loss.backward() -> FusedAdam.step() (the calls Lightning's fit loop makes)" if lightning else
                        "DataParallelTrainer.train_step (fused fwd+bwd -> gradient all-reduce -> FusedAdam)")
                rec["api"] = f"{src} -> {body}, loss read back to the host every step"
                if name == "e2e":
                    line[name] = rec
                else:
                    sec[name] = rec
            except Exception as e:  # noqa: BLE001
                (line if name == "e2e" else sec)[name] = dict(value=None, unit=UNIT, error=repr(e), h2d_bytes_per_step=0, d2h_bytes_per_step=0)
        del store

        # ---- secondary: sampling (BASELINE configs[3]: 100 M embeddings over 8 GPUs = 12.5 M per GPU, fp32 [N,256] left in a 12.8 GB buffer) ----
        try:
            Ns = args.sample_n
            out = torch.empty(Ns, D, dtype=torch.float32, device=dev)
            row0 = rank * Ns
            P.sample_on_device(module, Ns, out=out, row0=row0)
            reps = 3
            ms_s = time_events(lambda: [P.sample_on_device(module, Ns, out=out, row0=row0) for _ in range(reps)], torch, dist_on)
            sps = Ns * n_gpus * reps / (ms_s * 1e-3)
            burst_s = ms_s < 1000.0
            pk = peaks["bf16_burst"] if burst_s else peaks["bf16_sustained"]
            sec["unconditional_sampling"] = dict(value=sps, unit="embeddings/s", n_per_gpu=Ns, ms=ms_s / reps,
                                                 tensor_frac=sps * 851968 / 1e12 / (pk * n_gpus), peak_kind="burst" if burst_s else "sustained",
                                                 tensor_frac_of_sustained=sps * 851968 / 1e12 / (peaks["bf16_sustained"] * n_gpus),
                                                 tensor_frac_of_burst=sps * 851968 / 1e12 / (peaks["bf16_burst"] * n_gpus),
                                                 hbm_frac=sps * D * 4 / 1e9 / (peaks["hbm_gbs"] * n_gpus),
                                                 note="z ~ Philox in-kernel -> decoder chain -> fp32 [N,256] left in HBM; 851,968 FLOP + 1,024 B per sample")
            Nc = min(args.cond_n, Ns)
            outc = out[:Nc]
            P.sample_on_device(module, min(Nc, 1 << 20), classifier_target=1, num_steps=100, out=outc[:min(Nc, 1 << 20)], row0=rank * Nc)
            ms_c = time_events(lambda: P.sample_on_device(module, Nc, classifier_target=1, num_steps=100, out=outc, row0=rank * Nc), torch, dist_on)
            cps = Nc * n_gpus / (ms_c * 1e-3)
            # the ceiling of the Langevin loop: the bare counter-based generator (Philox4x32-10 + Box-Muller) writing fp32 normals
            zb = torch.empty(1 << 22, LAT, device=dev)
            stq = torch.cuda.current_stream().cuda_stream
            for i in range(2):
                L.check(L.lib().psvae_philox_normal(zb.data_ptr(), zb.shape[0], LAT, 1, i, 0, stq))
            ms_g = time_events(lambda: [L.check(L.lib().psvae_philox_normal(zb.data_ptr(), zb.shape[0], LAT, 1, i, 0, stq)) for i in range(10)], torch, False) / 10
            bare = zb.numel() / (ms_g * 1e-3)
            sec["conditional_sampling"] = dict(value=cps, unit="embeddings/s", n_per_gpu=Nc, num_steps=100, ms=ms_c,
                                               normals_per_s=cps * (100 * 64 + 64), bare_generator_normals_per_s=bare * n_gpus,
                                               generator_frac=cps * (100 * 64 + 64) / (bare * n_gpus),
                                               note="100 Langevin steps (1 launch) + decode; bound by Philox/Box-Muller on the CUDA cores: normals/s against "
                                                    "psvae_philox_normal alone (268 M fp32 normals written to HBM per launch)")
            del out, zb
        except Exception as e:  # noqa: BLE001
            sec["sampling_error"] = repr(e)
        # ---- secondary: the widened VAE of BASELINE configs[4] (512-d, 4 x 2048 hidden, latent classifier), batch 65,536 per GPU ------------------
        try:
            Dw, Hw, NHw, Bw = 512, 2048, 4, args.wide_batch
            torch.manual_seed(0)
            wide = P.PseudoSpeakerVAE(model=dict(input_dim=Dw, latent_dim=LAT, hidden_dim=Hw, num_hidden_layers=NHw),
                                      classifier=dict(input_dim=LAT, num_classes=NCLS), optimizer=dict(lr=1e-3), scheduler=dict(T_max=200),
                                      precision=args.precision).to(dev)
            wtr = P.DataParallelTrainer(wide)
            wtr.set_shard(Bw * n_gpus)
            wide.hot_path.manual_seed(1236, 0)
            xw, yw, _ = synth_batch(Bw, Dw, LAT, NCLS, seed=4321 + rank)
            wb = [(torch.from_numpy(xw).to(dev).to(tdt), torch.from_numpy(yw).to(dev)),
                  (torch.from_numpy(xw[::-1].copy()).to(dev).to(tdt), torch.from_numpy(yw[::-1].copy()).to(dev))]
            for i in range(3):
                wtr.train_step(*wb[i % 2])
            Kw = 10
            ms_w = time_events(lambda: [wtr.train_step(*wb[i % 2]) for i in range(Kw)], torch, dist_on)
            fl = float(wide.hot_path.flops_train)            # psvae_flops_per_sample: 243,532,544 (+768 classifier), SURVEY 8(d)
            wps = Bw * n_gpus * Kw / (ms_w * 1e-3)
            sec["widened_config5"] = dict(value=wps, unit=UNIT, batch_per_gpu=Bw, ms_per_step=ms_w / Kw, flops_per_sample=fl,
                                          tensor_frac=wps * fl / 1e12 / (peaks["bf16_burst"] * n_gpus), peak_kind="burst (timed region < 1 s)",
                                          tensor_frac_of_sustained=wps * fl / 1e12 / (peaks["bf16_sustained"] * n_gpus),
                                          note="D=512, 4 x 2048 hidden, L=64, 2-class latent classifier, fwd+bwd+Adam, same fused step")
            del wide, wtr, wb
        except Exception as e:  # noqa: BLE001
            sec["widened_config5"] = dict(error=repr(e))

        # ---- baselines on this box (rank 0, N = 1 only): the reference's torch path on the same GPU through stock torch, and on the host cores ----
        if rank == 0 and n_gpus == 1:
            try:
                sec["torch_gpu_reference"] = torch_gpu_baseline(dev, B, peaks)
                sec["torch_gpu_reference"]["speedup_of_value_vs_bf16_autocast"] = value / sec["torch_gpu_reference"]["bf16_autocast"]["value"]
            except Exception as e:  # noqa: BLE001
                sec["torch_gpu_reference"] = dict(error=repr(e))
            try:
                base = cpu_baseline(budget_s=20.0, batch=B)
                line["cpu_baseline"] = dict(value=base["value"], unit=UNIT, cores=base["cores"], kind=base["kind"], sample=base["sample"])
                sec["config1_cpu_batch256"] = base["config1_batch256"]
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample=f"failed: {e!r}")
    if sec:
        line["secondary"] = sec
    if dist_on:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
