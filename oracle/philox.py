"""ORACLE (test infrastructure, never on the product path).

CPU restatement (numpy, integer arithmetic exact) of the Exp(1) race-noise stream biom3_decode() draws on the
device when the caller supplies no noise (biom3_b200/csrc/kernels.cuh: philox4x32_10, philox_exp1).

The reference draws this noise inside ``OneHotCategorical.sample()`` (/root/reference/Stage3_source/
sampling_analysis.py:251): ATen's single-draw multinomial computes ``argmax(p / q)`` with ``q ~ Exp(1)`` taken
from the device generator; the stream itself is backend dependent and unseeded in the reference,
so the contract is the DISTRIBUTION, not torch's bits.  This module pins our stream two ways:

  * ``exp1_stream`` reproduces the device values (the 32-bit Philox words exactly; the float32 transform
    ``-log1p(-v)`` to the last ulp or two of libm vs CUDA), so tests can feed the oracle sampler the very noise the
    CUDA path consumes and demand identical tokens;
  * the statistical tests in tests/ check the stream against Exp(1) (KS) and the drawn tokens against softmax (chi^2).

Philox4x32-10 as published (Salmon et al., SC'11): counter (c0, c1, c2, c3), key (k0, k1), ten rounds of
    (c0, c1, c2, c3) <- (hi(M1 * c2) ^ c1 ^ k0, lo(M1 * c2), hi(M0 * c0) ^ c3 ^ k1, lo(M0 * c0)),  k += (W0, W1).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised over numpy uint32 counters; returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def uniform_words(seed: int, step: int, n_pos: int, C: int) -> np.ndarray:
    """The 32-bit word behind every draw: uint32 [n_pos, C]; word for (pos, c) = output (c & 3) of the Philox block
    with counter (pos, step, c >> 2, 0) and key (seed lo, seed hi)."""
    pos = np.arange(n_pos, dtype=np.uint64)
    out = np.empty((n_pos, C), dtype=np.uint32)
    for blk in range((C + 3) // 4):
        w = philox4x32_10(pos, np.full(n_pos, step, np.uint64), np.full(n_pos, blk, np.uint64), np.zeros(n_pos, np.uint64),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for j in range(4):
            c = blk * 4 + j
            if c < C:
                out[:, c] = w[j]
    return out


def exp1_from_words(w: np.ndarray) -> np.ndarray:
    """v = min(fma(float(w), 2^-32, 2^-33), 1 - 2^-24) in float32, q = -log1p(-v)  (kernels.cuh::philox_exp1)."""
    wf = w.astype(np.float32)                                   # round to nearest even, as the device conversion does
    v = (wf.astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(np.float32)     # one rounding, like fmaf
    v = np.minimum(v, np.float32(0.99999994))
    return (-np.log1p(-v.astype(np.float64))).astype(np.float32)


def exp1_stream(seed: int, step: int, n_pos: int, C: int) -> np.ndarray:
    """float32 [n_pos, C]: noise[pos][c] a decode step at time index ``step`` consumes under ``seed``."""
    return exp1_from_words(uniform_words(seed, step, n_pos, C))
