"""Sampling: drop-ins for ``unconditional_synthesis`` / ``conditional_synthesis`` of ps_vae/inference.py:10-110
and the CLI of :113-156.

Same signatures and return values (CPU tensors; ``(x_hat, history)`` with ``return_history=True`` where
``history`` is a list of ``num_steps`` numpy arrays ``[N, latent]``).  Underneath:

  * unconditional: z ~ N(0, I) is drawn by the library's counter-based Philox generator inside the decode call
    (counter = global sample index, so a sharded run reproduces the single-GPU batch bit for bit) and goes through
    the decoder GEMM chain;
  * conditional: the whole Langevin loop (inference.py:77-103) is ONE kernel launch -- closed-form classifier
    gradient, in-kernel noise, no host round trips (the reference does a D2H copy of z plus two ``.item()`` syncs
    every step); the z history is materialised only when asked for.

``sample_on_device`` is the bulk path (BASELINE config 4): it writes into a caller-provided device buffer and
returns without synchronising, and ``shard_rows`` gives each data-parallel rank its slice of the sample index range.
"""
from __future__ import annotations

import argparse
import os
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from .utils import parse_classifier_target, sample_filename


def _hot(vae_model):
    hot = getattr(vae_model, "hot_path", None)
    if hot is None:
        raise TypeError("vae_model must be a pseudo_speaker_vae_b200 PseudoSpeakerVAE / VAEModel")
    return hot() if callable(hot) else hot


def _check_device(vae_model, device) -> None:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"device={device!r}: pseudo_speaker_vae_b200 samples on a CUDA (B200) device only; there is no CPU fallback")
    mdev = _hot(vae_model).arena.device
    if mdev.type != "cuda" or (dev.index is not None and mdev.index != dev.index):
        raise RuntimeError(f"the model lives on {mdev} but device={device!r} was requested; move the model first (model.to(device))")


def shard_rows(num_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [row0, row0 + rows) of the global sample index range owned by ``rank``."""
    base, rem = divmod(int(num_samples), int(world_size))
    rows = base + (1 if rank < rem else 0)
    row0 = rank * base + min(rank, rem)
    return row0, rows


def sample_on_device(vae_model, num_samples: int, classifier_target: Union[int, dict, None] = None, step_size: float = 0.01,
                     num_steps: int = 100, noise_weight: float = 1.0, out: Optional[Tensor] = None, row0: int = 0) -> Tensor:
    """Generate ``num_samples`` embeddings into ``out`` (device tensor) without any host synchronisation.

    ``row0`` is the global index of the first sample: rank r of a sharded run passes ``shard_rows(N, r, W)``."""
    hot = _hot(vae_model)
    if classifier_target is None:
        return hot.decode(None, num_samples=num_samples, out=out, row0=row0)
    z, _, _ = hot.langevin(num_samples, classifier_target, step_size, num_steps, noise_weight, row0=row0)
    return hot.decode(z, out=out)


def unconditional_synthesis(vae_model, num_samples: int, device: str) -> Tensor:
    """z ~ N(0, I) -> decode -> CPU (inference.py:10-27)."""
    _check_device(vae_model, device)
    return _hot(vae_model).decode(None, num_samples=num_samples).detach().cpu()


def conditional_synthesis(vae_model, num_samples: int, classifier_target: Union[int, dict], step_size: float = 0.01, num_steps: int = 100,
                          noise_weight: float = 1.0, return_history: bool = False, device: str = "cpu", *, z0: Optional[Tensor] = None,
                          noise: Optional[Tensor] = None):
    """Classifier-guided Langevin dynamics in latent space, then decode (inference.py:29-110).

    ``z0`` / ``noise`` (keyword-only, optional) inject the initial draw and the per-step noise ``[num_steps, N, latent]``
    for parity runs; by default both come from the in-kernel Philox generator."""
    _check_device(vae_model, device)
    hot = _hot(vae_model)
    z, hist, _ = hot.langevin(num_samples, classifier_target, step_size, num_steps, noise_weight, z0=z0, noise=noise,
                              return_history=return_history)
    x_hat = hot.decode(z).detach().cpu()
    if return_history:
        h = hist.cpu().numpy()
        history: List = [h[i] for i in range(h.shape[0])]
        return x_hat, history
    return x_hat


def latent_transformation(vae_model, embeddings: Tensor, classifier_target: Union[int, dict], step_size: float = 0.02, max_steps: int = 10000,
                          noise_weight: float = 0.0, prior_weight: float = 0.5, threshold: float = 0.95, return_history: bool = False,
                          noise: Optional[Tensor] = None):
    """The attribute-transformation loop of analysis/sample_gender_transformation.py:57-99, batched: every embedding is encoded
    (``_, z, _ = vae_model(embed)``: z = the encoder mean), its latent ascends ``log p(y|z) + prior_weight * log p(z)`` with steps of
    ``0.5 * step_size**2`` (plus ``step_size * noise_weight * N(0, I)``) and is frozen after the update of the first step whose
    ``p(y|z)`` exceeded ``threshold`` -- one kernel launch for all samples and all steps instead of a Python loop with two ``.item()`` syncs per
    step.  Returns ``(x_hat [N, D] cpu, z [N, L] cpu, stop_step [N] cpu int32 (max_steps: never stopped), prob [N] cpu)``, plus the list of
    per-step latents with ``return_history``."""
    hot = _hot(vae_model)
    dev = hot.arena.device
    with torch.no_grad():
        _, mu, _ = hot.forward(embeddings.to(dev))
    z, hist, (_, stop, prob) = hot.langevin(mu.shape[0], classifier_target, step_size, max_steps, noise_weight, z0=mu, noise=noise,
                                             return_history=return_history, prior_weight=prior_weight, threshold=threshold, return_stop=True)
    x_hat = hot.decode(z).detach().cpu()
    out = (x_hat, z.cpu(), stop.cpu(), prob.cpu())
    if return_history:
        h = hist.cpu().numpy()
        return out + ([h[i] for i in range(h.shape[0])],)
    return out


def _write_rows(rows: Tensor, first: int, save_dir: str) -> List[str]:
    paths = []
    for j in range(rows.shape[0]):
        path = os.path.join(save_dir, sample_filename(first + j))
        torch.save(rows[j].clone(), path)
        paths.append(path)
    return paths


def save_samples(x_hat: Tensor, save_dir: str, workers: int = 0, chunk: int = 4096) -> List[str]:
    """Row i -> ``sample_{i}.pt`` (inference.py:154-156), file contents identical to the reference's ``torch.save(x, path)``.

    The reference writes one file per row in a serial loop after the whole batch has been copied to the host.  Here a device batch is
    brought over in chunks of ``chunk`` rows on a copy stream into pinned memory while ``workers`` threads (default: 8, or 0 for the
    caller's thread when the batch is small) serialise the previous chunk -- the D2H copy, the pickling and the file system calls overlap."""
    os.makedirs(save_dir, exist_ok=True)
    n = int(x_hat.shape[0])
    if workers <= 0:
        workers = 8 if n >= 256 else 0
    if workers == 0 and not x_hat.is_cuda:
        return _write_rows(x_hat, 0, save_dir)
    from concurrent.futures import ThreadPoolExecutor

    futures = []
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        if x_hat.is_cuda:
            stream = torch.cuda.Stream(x_hat.device)
            stream.wait_stream(torch.cuda.current_stream(x_hat.device))
            for a in range(0, n, chunk):
                b = min(n, a + chunk)
                host = torch.empty((b - a,) + tuple(x_hat.shape[1:]), dtype=x_hat.dtype, pin_memory=True)
                with torch.cuda.stream(stream):
                    host.copy_(x_hat[a:b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(stream)

                def job(host=host, a=a, ev=ev):
                    ev.synchronize()
                    return _write_rows(host, a, save_dir)

                futures.append(pool.submit(job))
        else:
            per = max(1, -(-n // (4 * max(1, workers))))
            for a in range(0, n, per):
                futures.append(pool.submit(_write_rows, x_hat[a:a + per], a, save_dir))
        paths: List[str] = []
        for f in futures:
            paths += f.result()
    return paths


def main(argv=None) -> None:
    from .lightning import PseudoSpeakerVAE

    parser = argparse.ArgumentParser(description="Generate synthetic embeddings using a VAE model.")
    parser.add_argument("--vae_ckpt_path", type=str, required=True, help="Path to the VAE checkpoint.")
    parser.add_argument("--n_samples", type=int, default=16, help="Number of samples to generate.")
    parser.add_argument("--save_dir", type=str, required=True, help="Directory to save the generated samples.")
    parser.add_argument("--synthesis_type", type=str, choices=["conditional", "unconditional"], required=True)
    parser.add_argument("--classifier_target", type=str, default="1", help="Target class (int, or JSON dict for multi-label).")
    parser.add_argument("--num_steps", type=int, default=5000)
    parser.add_argument("--step_size", type=float, default=0.01)
    parser.add_argument("--noise_weight", type=float, default=1.0)
    parser.add_argument("--device", type=str, default="cuda", help="CUDA device (this implementation has no CPU path).")
    args = parser.parse_args(argv)

    classifier_target = parse_classifier_target(args.classifier_target)
    model = PseudoSpeakerVAE.load_from_checkpoint(args.vae_ckpt_path)
    model = model.to(args.device)
    if args.synthesis_type == "conditional":
        x_hat = conditional_synthesis(model, classifier_target=classifier_target, num_samples=args.n_samples, num_steps=args.num_steps,
                                      step_size=args.step_size, noise_weight=args.noise_weight, device=args.device)
    else:
        x_hat = unconditional_synthesis(model, num_samples=args.n_samples, device=args.device)
    save_samples(x_hat, args.save_dir)


if __name__ == "__main__":
    main()
