"""Data plane in front of the hot path: a packed, memory-mapped embedding store and a pinned,
double-buffered host->device batch loader.

The reference reads ONE ``torch.load`` per sample in ``__getitem__`` (ps_vae/data/cv.py:73-76, ps_vae/data/vctk.py:62-64) and
lets a stock ``DataLoader`` collate them; at the batch sizes the B200 step wants (65,536 rows = 67 MB per step) that is tens
of thousands of file opens per step.  Here a dataset is packed ONCE into three flat files

    <dir>/embeddings.f32   [N][D] float32, row i = sample i (``embed.squeeze()`` of the reference)
    <dir>/labels.i64       [n_label_columns][N] int64 (the reference's metadata transforms, ps_vae/utils.py:82-136)
    <dir>/index.json       {"n", "dim", "label_names", "files"}

and training reads batches out of the memory map: ``PinnedBatchLoader`` gathers a batch's rows into one of two pinned staging
buffers, issues the H2D copy on its own CUDA stream, and hands the compute stream a tensor guarded by an event, so the copy of
batch k+1 runs under the step of batch k.  Sharding over data-parallel ranks follows ``DistributedSampler`` (rank r takes
indices r, r + W, ... of the epoch's permutation, padded by wrapping to equal length).

``PackedEmbeddingStore`` is also a map-style dataset with the reference's item contract -- ``store[i] -> (embedding [D],
label | {name: label})`` -- so the reference's ``get_*_dataloaders`` keep working on it.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .utils import map_cv_age_to_label, map_cv_gender_to_label, map_vctk_gender_to_label

#: metadata transforms of ps_vae/data/cv.py:8-16 and ps_vae/data/vctk.py:9-11, as {name: [(label column, function of the metadata row)]}
CV_TRANSFORMS: Dict[str, List[Tuple[str, Callable[[dict], int]]]] = {
    "gender": [("gender", lambda m: map_cv_gender_to_label(m["gender"]))],
    "age": [("age", lambda m: map_cv_age_to_label(m["age"]))],
    "age_and_gender": [("age", lambda m: map_cv_age_to_label(m["age"])), ("gender", lambda m: map_cv_gender_to_label(m["gender"]))],
}
VCTK_TRANSFORMS: Dict[str, List[Tuple[str, Callable[[dict], int]]]] = {
    "gender": [("gender", lambda m: map_vctk_gender_to_label(m["gender"]))],
}


class PackedEmbeddingStore(torch.utils.data.Dataset):
    """[N][D] fp32 embeddings + int64 label columns, memory-mapped from ``root``."""

    def __init__(self, root: str, mode: str = "r"):
        with open(os.path.join(root, "index.json")) as f:
            idx = json.load(f)
        self.root = root
        self.n, self.dim = int(idx["n"]), int(idx["dim"])
        self.label_names: List[str] = list(idx["label_names"])
        self.files: List[str] = list(idx.get("files", []))
        self.multilabel = bool(idx.get("multilabel", len(self.label_names) > 1))
        self.embeddings = np.memmap(os.path.join(root, "embeddings.f32"), dtype=np.float32, mode=mode, shape=(self.n, self.dim))
        nl = len(self.label_names)
        self.labels = np.memmap(os.path.join(root, "labels.i64"), dtype=np.int64, mode=mode, shape=(nl, self.n)) if nl else np.zeros((0, self.n), np.int64)

    # ---- building -----------------------------------------------------------------------------------------------
    @classmethod
    def build(cls, root: str, samples: Iterable[Tuple[torch.Tensor, Union[int, Dict[str, int], None]]], n: int, dim: int,
              label_names: Sequence[str] = (), files: Optional[Sequence[str]] = None, multilabel: Optional[bool] = None) -> "PackedEmbeddingStore":
        """Pack ``n`` (embedding, label) items -- the reference's ``__getitem__`` contract -- into ``root``."""
        os.makedirs(root, exist_ok=True)
        label_names = list(label_names)
        emb = np.memmap(os.path.join(root, "embeddings.f32"), dtype=np.float32, mode="w+", shape=(n, dim))
        lab = np.memmap(os.path.join(root, "labels.i64"), dtype=np.int64, mode="w+", shape=(max(1, len(label_names)), n))
        count = 0
        for i, (e, y) in enumerate(samples):
            if i >= n:
                raise ValueError(f"more than n={n} samples")
            v = torch.as_tensor(e).detach().to(torch.float32).squeeze().reshape(-1).numpy()
            if v.shape[0] != dim:
                raise ValueError(f"sample {i} has {v.shape[0]} elements, expected dim={dim}")
            emb[i] = v
            if label_names:
                if isinstance(y, dict):
                    for c, name in enumerate(label_names):
                        lab[c, i] = int(y[name])
                else:
                    lab[0, i] = int(y)
            count += 1
        if count != n:
            raise ValueError(f"got {count} samples, expected n={n}")
        emb.flush()
        lab.flush()
        del emb, lab
        if not label_names:
            os.truncate(os.path.join(root, "labels.i64"), 0)
        with open(os.path.join(root, "index.json"), "w") as f:
            json.dump({"n": n, "dim": dim, "label_names": label_names, "files": list(files or []),
                       "multilabel": (len(label_names) > 1) if multilabel is None else bool(multilabel)}, f)
        return cls(root)

    @classmethod
    def from_cv(cls, data_root: str, out_root: str, split: str = "train", se_model: str = "sb", metadata_transform: Optional[str] = None):
        """Pack what ``CVEmbeddingDataset(data_root, split, se_model, metadata_transform)`` would serve (ps_vae/data/cv.py:17-76):
        every ``*.pth`` under ``embeds_<se_model>/<split>/``, labels from ``<split>.tsv`` through the same transforms."""
        if metadata_transform is not None and metadata_transform not in CV_TRANSFORMS:
            raise AssertionError(f"Invalid metadata transform: {metadata_transform}")
        with open(os.path.join(data_root, f"{split}.tsv"), "r", encoding="utf-8") as f:
            lines = f.readlines()
        headers = lines[0].strip().split("\t")
        rows = [dict(zip(headers, line.strip().split("\t"))) for line in lines[1:]]
        meta = {r.pop("path").replace(".mp3", ".pth"): r for r in rows}
        embed_dir = os.path.join(data_root, f"embeds_{se_model}", split)
        files = [f for f in os.listdir(embed_dir) if f.endswith(".pth")]        # the reference keeps os.listdir order (cv.py:52)
        cols = CV_TRANSFORMS[metadata_transform] if metadata_transform else []
        return cls._pack_files(out_root, [os.path.join(embed_dir, f) for f in files], [meta[f] for f in files], files, cols,
                               multilabel=metadata_transform == "age_and_gender")

    @classmethod
    def from_vctk(cls, data_root: str, out_root: str, metadata_transform: Optional[str] = None):
        """Pack what ``VCTKEmbeddingDataset(data_root, metadata_transform=...)`` would serve (ps_vae/data/vctk.py:13-67)."""
        import csv

        if metadata_transform is not None and metadata_transform not in VCTK_TRANSFORMS:
            raise AssertionError(f"Invalid metadata transform: {metadata_transform}")
        with open(os.path.join(data_root, "vctk_metadata.csv"), newline="") as f:
            meta = {r["file_name"]: {k: v for k, v in r.items() if k != "file_name"} for r in csv.DictReader(f)}
        paths, names = [], []
        for speaker in os.listdir(data_root):
            if speaker.startswith("p") and os.path.isdir(os.path.join(data_root, speaker)):
                for fn in os.listdir(os.path.join(data_root, speaker)):
                    if fn.endswith(".pt"):
                        paths.append(os.path.join(data_root, speaker, fn))
                        names.append(fn)
        metas = [meta[fn.replace("_mic1.pt", ".wav")] for fn in names]
        cols = VCTK_TRANSFORMS[metadata_transform] if metadata_transform else []
        return cls._pack_files(out_root, paths, metas, names, cols, multilabel=False)

    @classmethod
    def _pack_files(cls, out_root, paths, metas, names, cols, multilabel):
        if not paths:
            raise ValueError("no embedding files found")
        first = torch.load(paths[0], weights_only=False)
        dim = int(torch.as_tensor(first).squeeze().numel())

        def items():
            for p, m in zip(paths, metas):
                e = torch.load(p, weights_only=False)
                if not cols:
                    yield e, None
                elif multilabel:
                    yield e, {name: fn(m) for name, fn in cols}
                else:
                    yield e, cols[0][1](m)

        return cls.build(out_root, items(), len(paths), dim, [name for name, _ in cols], names, multilabel=multilabel)

    # ---- the reference's item contract ----------------------------------------------------------------------------
    def __len__(self) -> int:
        return self.n

    def __getitem__(self, idx: int):
        if not -self.n <= idx < self.n:
            raise IndexError(idx)
        e = torch.from_numpy(np.array(self.embeddings[idx]))
        if not self.label_names:
            return e, {}
        if self.multilabel:
            return e, {name: int(self.labels[c, idx]) for c, name in enumerate(self.label_names)}
        return e, int(self.labels[0, idx])

    # ---- batched access -------------------------------------------------------------------------------------------
    def gather(self, indices: np.ndarray, out_x: torch.Tensor, out_y: Optional[torch.Tensor] = None) -> None:
        """Rows ``indices`` -> ``out_x[:len]`` (float32 [*, D]) and ``out_y[:, :len]`` (int64 [n_label_columns, *]), host tensors."""
        k = len(indices)
        xs = out_x.numpy()
        srt = np.all(indices[1:] == indices[:-1] + 1) if k > 1 else True
        if srt:                                   # a contiguous run of the store: one memcpy out of the page cache
            xs[:k] = self.embeddings[indices[0]:indices[0] + k]
        else:
            np.take(self.embeddings, indices, axis=0, out=xs[:k])
        if out_y is not None and self.label_names:
            ys = out_y.numpy()
            for c in range(len(self.label_names)):
                if srt:
                    ys[c, :k] = self.labels[c, indices[0]:indices[0] + k]
                else:
                    np.take(self.labels[c], indices, out=ys[c, :k])


def shard_indices(n: int, rank: int, world_size: int, shuffle: bool, seed: int, epoch: int, drop_last: bool = False) -> np.ndarray:
    """The index list ``torch.utils.data.DistributedSampler`` hands rank ``rank`` (same permutation source semantics: a
    generator seeded with seed + epoch; padded by wrapping so that every rank gets the same count, or truncated with drop_last)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).numpy()
    else:
        idx = np.arange(n)
    if world_size <= 1:
        return idx
    if drop_last and n % world_size:
        total = (n // world_size) * world_size
        idx = idx[:total]
    else:
        total = -(-n // world_size) * world_size
        pad = total - n
        if pad:
            reps = -(-pad // max(1, n))
            idx = np.concatenate([idx, np.tile(idx, reps)[:pad]])
    return idx[rank:total:world_size]


class PinnedBatchLoader:
    """Iterates ``(x, y)`` device batches of a ``PackedEmbeddingStore``.

    ``y`` follows the reference's collated batch: an int64 tensor [B] (single label), a dict {name: tensor} (multi-label), or
    the zero tensor the trainer ignores when the store has no labels.  With ``device`` a CUDA device, two pinned staging
    buffers alternate: while the caller trains on batch k, batch k+1 is gathered and copied on ``copy_stream``; the yielded
    tensors are made safe for the CURRENT stream with an event wait (no host synchronisation).  With ``device='cpu'`` (tests,
    tooling) the same batches come back as host tensors."""

    def __init__(self, store: PackedEmbeddingStore, batch_size: int, device: Union[str, torch.device] = "cuda", shuffle: bool = True,
                 seed: int = 0, rank: int = 0, world_size: int = 1, drop_last: bool = False, indices: Optional[np.ndarray] = None):
        self.store, self.batch_size = store, int(batch_size)
        self.device = torch.device(device)
        self.shuffle, self.seed, self.rank, self.world_size, self.drop_last = shuffle, int(seed), int(rank), int(world_size), drop_last
        self.subset = None if indices is None else np.asarray(indices, dtype=np.int64)       # e.g. one side of a train / val split
        self.epoch = 0
        self._cuda = self.device.type == "cuda"
        nl = max(1, len(store.label_names))
        pin = self._cuda and torch.cuda.is_available()
        self._hx = [torch.empty(self.batch_size, store.dim, dtype=torch.float32, pin_memory=pin) for _ in range(2)]
        self._hy = [torch.empty(nl, self.batch_size, dtype=torch.int64, pin_memory=pin) for _ in range(2)]
        self._dx = self._dy = None
        self._copy_stream = torch.cuda.Stream(self.device) if self._cuda else None
        self._free = [None, None]        # event: the compute stream is done with device buffer i (recorded when the NEXT batch is requested)

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def _order(self) -> np.ndarray:
        n = self.store.n if self.subset is None else len(self.subset)
        idx = shard_indices(n, self.rank, self.world_size, self.shuffle, self.seed, self.epoch, self.drop_last)
        return idx if self.subset is None else self.subset[idx]

    def __len__(self) -> int:
        n = len(self._order())
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def _labels(self, y: torch.Tensor, k: int):
        names = self.store.label_names
        if not names:
            return torch.zeros(k, dtype=torch.int64, device=y.device)
        if self.store.multilabel:
            return {name: y[c, :k] for c, name in enumerate(names)}
        return y[0, :k]

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, Union[torch.Tensor, Dict[str, torch.Tensor]]]]:
        order = self._order()
        B = self.batch_size
        starts = list(range(0, len(order) - (B - 1 if self.drop_last else 0), B))
        if not self._cuda:
            for s in starts:
                idx = order[s:s + B]
                x = torch.empty(len(idx), self.store.dim, dtype=torch.float32)
                y = torch.empty(max(1, len(self.store.label_names)), len(idx), dtype=torch.int64)
                self.store.gather(idx, x, y)
                yield x, self._labels(y, len(idx))
            return
        if self._dx is None:
            nl = max(1, len(self.store.label_names))
            self._dx = [torch.empty(B, self.store.dim, dtype=torch.float32, device=self.device) for _ in range(2)]
            self._dy = [torch.empty(nl, B, dtype=torch.int64, device=self.device) for _ in range(2)]
        staged: List[Optional[Tuple[int, torch.cuda.Event]]] = [None, None]

        def stage(slot: int, s: int):
            idx = order[s:s + B]
            if staged[slot] is not None:
                staged[slot][1].synchronize()                       # the previous H2D out of this pinned buffer has finished
            self.store.gather(idx, self._hx[slot], self._hy[slot])
            with torch.cuda.stream(self._copy_stream):
                if self._free[slot] is not None:
                    self._copy_stream.wait_event(self._free[slot])  # the step that used device buffer `slot` is done with it
                self._dx[slot][:len(idx)].copy_(self._hx[slot][:len(idx)], non_blocking=True)
                self._dy[slot][:, :len(idx)].copy_(self._hy[slot][:, :len(idx)], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            staged[slot] = (len(idx), ev)

        if starts:
            stage(0, starts[0])
        for i, s in enumerate(starts):
            slot = i & 1
            if i + 1 < len(starts):
                stage(slot ^ 1, starts[i + 1])
            k, ev = staged[slot]
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            yield self._dx[slot][:k], self._labels(self._dy[slot], k)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self._free[slot] = done


def get_packed_dataloaders(store: PackedEmbeddingStore, batch_size: int = 16, train_frac: float = 1.0, device: Union[str, torch.device] = "cuda",
                           seed: int = 0, **loader_kwargs):
    """Counterpart of ``get_cv_dataloaders`` / ``get_vctk_dataloaders`` (cv.py:79-138, vctk.py:69-109) on a packed store:
    one loader, or ``{"train", "val"}`` loaders over a random split when ``train_frac < 1``."""
    if train_frac >= 1.0:
        return PinnedBatchLoader(store, batch_size, device, seed=seed, **loader_kwargs)
    n = len(store)
    n_train = int(n * train_frac)
    g = torch.Generator()
    g.manual_seed(seed)
    perm = torch.randperm(n, generator=g).numpy()
    return {"train": PinnedBatchLoader(store, batch_size, device, seed=seed, indices=np.sort(perm[:n_train]), **loader_kwargs),
            "val": PinnedBatchLoader(store, batch_size, device, seed=seed, indices=np.sort(perm[n_train:]), **loader_kwargs)}
