"""Data plane in front of the hot path: a packed, memory-mapped embedding store and a pinned,
double-buffered host->device batch loader.

The reference reads ONE ``torch.load`` per sample in ``__getitem__`` (ps_vae/data/cv.py:73-76, ps_vae/data/vctk.py:62-64) and
lets a stock ``DataLoader`` collate them; at the batch sizes the B200 step wants (65,536 rows = 67 MB per step) that is tens
of thousands of file opens per step.  Here a dataset is packed ONCE into three flat files

    <dir>/embeddings.f32 | embeddings.bf16   [N][D] float32 or bfloat16, row i = sample i (``embed.squeeze()`` of the reference)
    <dir>/labels.i64       [n_label_columns][N] int64 (the reference's metadata transforms, ps_vae/utils.py:82-136)
    <dir>/index.json       {"n", "dim", "dtype", "label_names", "files"}

and training reads batches out of it.  Two ways, same batches (sharding over data-parallel ranks follows ``DistributedSampler``:
rank r takes indices r, r + W, ... of the epoch's permutation, padded by wrapping to equal length):

* streaming (``PinnedBatchLoader``, default): a batch's rows are gathered into one of two pinned staging buffers -- or, for a
  contiguous run of a store that was loaded into pinned memory (``store.pin()``), taken from there without a staging copy -- and
  copied H2D on a copy stream; the compute stream gets a tensor guarded by an event, so the copy of batch k+1 runs under the
  step of batch k.  A bf16 store halves the bytes (33.5 MB per 65,536-row step: 0.63 ms of PCIe Gen5 under a 0.75 ms step) and
  the tensor-core step takes the bf16 rows as they are (``PSVAE_X_BF16``: no cast pass).
* resident (``PinnedBatchLoader(resident=True)``): the whole store is uploaded to HBM once (Common Voice train, 674,068 x 256:
  345 MB in bf16 of 180 GB) and every batch is assembled ON the device by ``psvae_gather_rows`` from the epoch's index list --
  the only per-step PCIe traffic is 8 B per row of indices.  At B200 step rates (65,536 rows in 0.75 ms = 45 GB/s of bf16 rows
  per GPU, 8 GPUs on one host) this is the only way the input side keeps up.

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


_DTYPES = {"f32": (np.float32, torch.float32, "embeddings.f32"), "bf16": (np.uint16, torch.bfloat16, "embeddings.bf16")}


def _bf16_bits(v: torch.Tensor) -> np.ndarray:
    """fp32 tensor -> uint16 array of its round-to-nearest-even bf16 bit patterns."""
    return v.to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


class PackedEmbeddingStore(torch.utils.data.Dataset):
    """[N][D] fp32 or bf16 embeddings + int64 label columns, memory-mapped from ``root``."""

    def __init__(self, root: str, mode: str = "r"):
        with open(os.path.join(root, "index.json")) as f:
            idx = json.load(f)
        self.root = root
        self.n, self.dim = int(idx["n"]), int(idx["dim"])
        self.dtype_name = str(idx.get("dtype", "f32"))
        if self.dtype_name not in _DTYPES:
            raise ValueError(f"unknown store dtype {self.dtype_name!r}")
        np_dt, self.torch_dtype, fname = _DTYPES[self.dtype_name]
        self.label_names: List[str] = list(idx["label_names"])
        self.files: List[str] = list(idx.get("files", []))
        self.multilabel = bool(idx.get("multilabel", len(self.label_names) > 1))
        self.embeddings = np.memmap(os.path.join(root, fname), dtype=np_dt, mode=mode, shape=(self.n, self.dim))     # bf16: the raw 16-bit patterns
        nl = len(self.label_names)
        self.labels = np.memmap(os.path.join(root, "labels.i64"), dtype=np.int64, mode=mode, shape=(nl, self.n)) if nl else np.zeros((0, self.n), np.int64)
        self._pinned: Optional[torch.Tensor] = None          # pin(): the whole store in page-locked memory
        self._resident: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}   # to_device(): {device: (embeddings, labels)} in HBM

    # ---- building -----------------------------------------------------------------------------------------------
    @classmethod
    def build(cls, root: str, samples: Iterable[Tuple[torch.Tensor, Union[int, Dict[str, int], None]]], n: int, dim: int,
              label_names: Sequence[str] = (), files: Optional[Sequence[str]] = None, multilabel: Optional[bool] = None,
              dtype: str = "f32") -> "PackedEmbeddingStore":
        """Pack ``n`` (embedding, label) items -- the reference's ``__getitem__`` contract -- into ``root``.  ``dtype='bf16'`` stores the
        rows rounded to bfloat16 (what the tensor-core step computes with anyway; half the file, half the PCIe bytes)."""
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}, got {dtype!r}")
        np_dt, _, fname = _DTYPES[dtype]
        os.makedirs(root, exist_ok=True)
        label_names = list(label_names)
        emb = np.memmap(os.path.join(root, fname), dtype=np_dt, mode="w+", shape=(n, dim))
        lab = np.memmap(os.path.join(root, "labels.i64"), dtype=np.int64, mode="w+", shape=(max(1, len(label_names)), n))
        count = 0
        for i, (e, y) in enumerate(samples):
            if i >= n:
                raise ValueError(f"more than n={n} samples")
            v = torch.as_tensor(e).detach().to(torch.float32).squeeze().reshape(-1)
            if v.shape[0] != dim:
                raise ValueError(f"sample {i} has {v.shape[0]} elements, expected dim={dim}")
            emb[i] = v.numpy() if dtype == "f32" else _bf16_bits(v)
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
            json.dump({"n": n, "dim": dim, "dtype": dtype, "label_names": label_names, "files": list(files or []),
                       "multilabel": (len(label_names) > 1) if multilabel is None else bool(multilabel)}, f)
        return cls(root)

    @classmethod
    def from_arrays(cls, root: str, embeddings, labels=None, label_names: Sequence[str] = (), dtype: str = "f32", multilabel: Optional[bool] = None):
        """Pack an [N][D] array (numpy / torch) and optional int label columns ([N] or [n_columns][N]) in one go."""
        e = torch.as_tensor(embeddings).detach().to(torch.float32)
        n, dim = int(e.shape[0]), int(e.shape[1])
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}, got {dtype!r}")
        np_dt, _, fname = _DTYPES[dtype]
        os.makedirs(root, exist_ok=True)
        label_names = list(label_names)
        emb = np.memmap(os.path.join(root, fname), dtype=np_dt, mode="w+", shape=(n, dim))
        emb[:] = e.numpy() if dtype == "f32" else _bf16_bits(e)
        emb.flush()
        del emb
        with open(os.path.join(root, "labels.i64"), "wb") as f:
            if label_names:
                lab = np.asarray(torch.as_tensor(labels).numpy() if not isinstance(labels, np.ndarray) else labels, dtype=np.int64).reshape(len(label_names), n)
                f.write(np.ascontiguousarray(lab).tobytes())
        with open(os.path.join(root, "index.json"), "w") as f:
            json.dump({"n": n, "dim": dim, "dtype": dtype, "label_names": label_names, "files": [],
                       "multilabel": (len(label_names) > 1) if multilabel is None else bool(multilabel)}, f)
        return cls(root)

    @classmethod
    def from_cv(cls, data_root: str, out_root: str, split: str = "train", se_model: str = "sb", metadata_transform: Optional[str] = None, dtype: str = "f32"):
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
                               multilabel=metadata_transform == "age_and_gender", dtype=dtype)

    @classmethod
    def from_vctk(cls, data_root: str, out_root: str, metadata_transform: Optional[str] = None, dtype: str = "f32"):
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
        return cls._pack_files(out_root, paths, metas, names, cols, multilabel=False, dtype=dtype)

    @classmethod
    def _pack_files(cls, out_root, paths, metas, names, cols, multilabel, dtype="f32"):
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

        return cls.build(out_root, items(), len(paths), dim, [name for name, _ in cols], names, multilabel=multilabel, dtype=dtype)

    # ---- the reference's item contract ----------------------------------------------------------------------------
    def __len__(self) -> int:
        return self.n

    def __getitem__(self, idx: int):
        if not -self.n <= idx < self.n:
            raise IndexError(idx)
        e = torch.from_numpy(np.array(self.embeddings[idx]))
        if self.dtype_name == "bf16":
            e = e.view(torch.bfloat16).to(torch.float32)      # the reference's item contract: a float32 embedding
        if not self.label_names:
            return e, {}
        if self.multilabel:
            return e, {name: int(self.labels[c, idx]) for c, name in enumerate(self.label_names)}
        return e, int(self.labels[0, idx])

    # ---- batched access -------------------------------------------------------------------------------------------
    def labelled_indices(self) -> np.ndarray:
        """Samples whose every label column is >= 0: ``map_cv_*_to_label`` returns -1 for metadata it does not know (utils.py:92-119),
        which ``cross_entropy`` rejects; a loader built with ``indices=store.labelled_indices()`` never serves such a row."""
        if not self.label_names:
            return np.arange(self.n)
        return np.nonzero(np.all(np.asarray(self.labels) >= 0, axis=0))[0]

    def pin(self) -> torch.Tensor:
        """The whole store in ONE page-locked host tensor (loaded once): contiguous runs of it go to the device without a staging copy."""
        if self._pinned is None:
            t = torch.empty(self.n, self.dim, dtype=self.torch_dtype, pin_memory=torch.cuda.is_available())
            (t.view(torch.int16) if self.dtype_name == "bf16" else t).numpy()[:] = self.embeddings.view(np.int16) if self.dtype_name == "bf16" else self.embeddings
            self._pinned = t
        return self._pinned

    def to_device(self, device: Union[str, torch.device]) -> Tuple[torch.Tensor, torch.Tensor]:
        """(embeddings [N][D], labels [n_label_columns][N]) resident in HBM on ``device`` (uploaded once)."""
        key = str(torch.device(device))
        if key not in self._resident:
            e = self.pin().to(device, non_blocking=False)
            y = torch.from_numpy(np.array(self.labels)).to(device) if self.label_names else torch.zeros(1, self.n, dtype=torch.int64, device=device)
            self._resident[key] = (e, y)
        return self._resident[key]

    def gather(self, indices: np.ndarray, out_x: Optional[torch.Tensor], out_y: Optional[torch.Tensor] = None) -> None:
        """Rows ``indices`` -> ``out_x[:len]`` ([*, D] in the store's dtype; None: labels only) and ``out_y[:, :len]`` (int64
        [n_label_columns, *]), host tensors."""
        k = len(indices)
        srt = np.all(indices[1:] == indices[:-1] + 1) if k > 1 else True
        if out_x is not None:
            if out_x.dtype != self.torch_dtype:
                raise TypeError(f"out_x must be {self.torch_dtype} for this store, got {out_x.dtype}")
            xs = out_x.view(torch.int16).numpy().view(np.uint16) if self.dtype_name == "bf16" else out_x.numpy()
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

    ``x`` comes in the store's dtype (float32, or bfloat16 for a bf16 store -- the tensor-core step takes it as it is); ``y`` follows
    the reference's collated batch: an int64 tensor [B] (single label), a dict {name: tensor} (multi-label), or the zero tensor the
    trainer ignores when the store has no labels.  With ``device`` a CUDA device, two device buffers alternate: while the caller
    trains on batch k, batch k+1 is brought in on ``copy_stream``; the yielded tensors are made safe for the CURRENT stream with an
    event wait (no host synchronisation).  Streaming mode gathers the rows into a pinned staging buffer and copies it H2D (a
    contiguous run of a ``store.pin()``-ned store is copied straight out of the pinned store); ``resident=True`` keeps the store in
    HBM and assembles the batch there (``psvae_gather_rows``), so only the index list crosses PCIe.  With ``device='cpu'`` (tests,
    tooling) the same batches come back as host tensors."""

    def __init__(self, store: PackedEmbeddingStore, batch_size: int, device: Union[str, torch.device] = "cuda", shuffle: bool = True,
                 seed: int = 0, rank: int = 0, world_size: int = 1, drop_last: bool = False, indices: Optional[np.ndarray] = None,
                 resident: bool = False):
        self.store, self.batch_size = store, int(batch_size)
        self.device = torch.device(device)
        self.shuffle, self.seed, self.rank, self.world_size, self.drop_last = shuffle, int(seed), int(rank), int(world_size), drop_last
        self.subset = None if indices is None else np.asarray(indices, dtype=np.int64)       # e.g. one side of a train / val split
        self.epoch = 0
        self._cuda = self.device.type == "cuda"
        self.resident = bool(resident) and self._cuda
        nl = max(1, len(store.label_names))
        pin = self._cuda and torch.cuda.is_available()
        if self.resident:
            self._hi = [torch.empty(self.batch_size, dtype=torch.int64, pin_memory=pin) for _ in range(2)]      # the step's index list
            self._hx = self._hy = None
        else:
            self._hx = [torch.empty(self.batch_size, store.dim, dtype=store.torch_dtype, pin_memory=pin) for _ in range(2)]
            self._hy = [torch.empty(nl, self.batch_size, dtype=torch.int64, pin_memory=pin) for _ in range(2)]
        self._dx = self._dy = self._di = None
        self._copy_stream = torch.cuda.Stream(self.device) if self._cuda else None
        self._free = [None, None]        # event: the compute stream is done with device buffer i (recorded when the NEXT batch is requested)
        self.h2d_bytes_per_batch = (8 * self.batch_size) if self.resident else (store.dim * self.batch_size * (2 if store.dtype_name == "bf16" else 4)
                                                                                 + 8 * self.batch_size * len(store.label_names))

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
                x = torch.empty(len(idx), self.store.dim, dtype=self.store.torch_dtype)
                y = torch.empty(max(1, len(self.store.label_names)), len(idx), dtype=torch.int64)
                self.store.gather(idx, x, y)
                yield x, self._labels(y, len(idx))
            return
        nl = max(1, len(self.store.label_names))
        if self._dx is None:
            self._dx = [torch.empty(B, self.store.dim, dtype=self.store.torch_dtype, device=self.device) for _ in range(2)]
            self._dy = [torch.empty(nl, B, dtype=torch.int64, device=self.device) for _ in range(2)]
            if self.resident:
                self._di = [torch.empty(B, dtype=torch.int64, device=self.device) for _ in range(2)]
        staged: List[Optional[Tuple[int, torch.cuda.Event]]] = [None, None]
        if self.resident:
            from . import _lib as L

            ex, ey = self.store.to_device(self.device)
            row_bytes = self.store.dim * ex.element_size()
        pinned = self.store._pinned            # set by store.pin(): contiguous runs skip the staging copy

        def stage(slot: int, s: int):
            idx = order[s:s + B]
            k = len(idx)
            if staged[slot] is not None:
                staged[slot][1].synchronize()                       # the previous H2D out of this pinned buffer has finished
            if self.resident:
                self._hi[slot].numpy()[:k] = idx
            else:
                run = k > 0 and pinned is not None and bool(np.all(idx[1:] == idx[:-1] + 1))
                if run:
                    self.store.gather(idx, None, self._hy[slot])      # labels only (a few hundred KB)
                else:
                    self.store.gather(idx, self._hx[slot], self._hy[slot])
            with torch.cuda.stream(self._copy_stream):
                if self._free[slot] is not None:
                    self._copy_stream.wait_event(self._free[slot])  # the step that used device buffer `slot` is done with it
                if self.resident:
                    self._di[slot][:k].copy_(self._hi[slot][:k], non_blocking=True)
                    with torch.cuda.device(self.device):
                        st = self._copy_stream.cuda_stream
                        L.check(L.lib().psvae_gather_rows(ex.data_ptr(), self.store.n, row_bytes, self._di[slot].data_ptr(), k, self._dx[slot].data_ptr(), st),
                                "psvae_gather_rows")
                        for c in range(len(self.store.label_names)):      # label columns: 8-byte rows do not fit the 16-byte gather; plain index_select
                            torch.index_select(ey[c], 0, self._di[slot][:k], out=self._dy[slot][c, :k])
                else:
                    src = pinned[int(idx[0]):int(idx[0]) + k] if run else self._hx[slot][:k]
                    self._dx[slot][:k].copy_(src, non_blocking=True)
                    self._dy[slot][:, :k].copy_(self._hy[slot][:, :k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            staged[slot] = (k, ev)

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
