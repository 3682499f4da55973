"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based generator the CUDA kernels use.

Philox4x32-10 (Salmon et al., SC'11; the Random123 constants) is integer arithmetic and must match the
kernels bit for bit; it is pinned by the Random123 known-answer vectors in tests/test_oracle_golden.py.
The normal transform is the repo's own definition (the reference draws from torch's global generator --
model.py:57, inference.py:23,73,95 -- which no counter-based kernel can reproduce; parity runs inject
the normals instead):

    element e of a row-major [N, L] tensor whose first row has global index row0:
        g = row0*L + e;  q = g // 4;  lane = g % 4
        (r0,r1,r2,r3) = philox4x32_10(counter=(q_lo, q_hi, offset_lo, offset_hi), key=(seed_lo, seed_hi))
        pair (r0,r1) -> lanes 0,1 ; pair (r2,r3) -> lanes 2,3
        u = fma(float(ra), 2^-32, 2^-33);  s = sqrt(-2 ln u)
        t = fma(float(rb), 2^-31, -1)                 # in [-1, 1]
        lane even -> s * sin(pi*t) ; lane odd -> s * cos(pi*t)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(a, dtype=np.uint32) for a in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * M0
            p1 = c2.astype(np.uint64) * M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_uint32(n_elems: int, seed: int, offset: int, first_elem: int = 0) -> np.ndarray:
    """The raw uint32 stream: element g -> word (g % 4) of the block with counter q = g // 4."""
    g = np.arange(first_elem, first_elem + n_elems, dtype=np.uint64)
    q = g >> np.uint64(2)
    r = philox4x32_10((q & MASK).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                      np.uint32(offset & 0xFFFFFFFF), np.uint32((offset >> 32) & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r = np.stack(r, axis=-1)
    lane = (g & np.uint64(3)).astype(np.int64)
    return r[np.arange(n_elems), lane]


def philox_normal(n_rows: int, n_cols: int, seed: int, offset: int, row0: int = 0, dtype=np.float32) -> np.ndarray:
    """[n_rows, n_cols] standard normals for global rows row0..row0+n_rows (fp64 math, cast at the end)."""
    n = n_rows * n_cols
    first = row0 * n_cols
    g = np.arange(first, first + n, dtype=np.uint64)
    q = g >> np.uint64(2)
    r = philox4x32_10((q & MASK).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                      np.uint32(offset & 0xFFFFFFFF), np.uint32((offset >> 32) & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lane = (g & np.uint64(3)).astype(np.int64)
    pair_hi = lane >= 2
    ra = np.where(pair_hi, r[2], r[0]).astype(np.float32)     # float(r): round-to-nearest like cvt.rn.f32.u32
    rb = np.where(pair_hi, r[3], r[1]).astype(np.float32)
    u = (ra * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float64)   # exact in fp32 up to the final rounding
    u = np.float32(u).astype(np.float64)
    t = np.float32(rb.astype(np.float64) * 2.0 ** -31 - 1.0).astype(np.float64)   # one rounding == fmaf
    s = np.sqrt(-2.0 * np.log(u))
    out = np.where((lane & 1) == 0, s * np.sin(np.pi * t), s * np.cos(np.pi * t))
    return out.reshape(n_rows, n_cols).astype(dtype)
