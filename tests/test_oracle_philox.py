"""CPU tests of oracle/philox.py, the restatement of the on-device Exp(1) race-noise stream (csrc/kernels.cuh:
philox4x32_10 + philox_exp1): Random123's published known-answer vectors for Philox4x32-10, the distribution of the
stream (Kolmogorov-Smirnov against Exp(1)), and the statistics of the tokens the reference sampler's race
``argmax(p / q)`` (/root/reference/Stage3_source/sampling_analysis.py:251) draws with it (chi-square against softmax)."""
import numpy as np
import torch
from scipy import stats

from oracle import philox
from oracle import sampler as osamp


def _block(c, k):
    r = philox.philox4x32_10(*[np.array([x], dtype=np.uint64) for x in c], k[0], k[1])
    return tuple(int(x[0]) for x in r)


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors: philox4x32 10 rounds."""
    assert _block((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert _block((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert _block((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_stream_layout_and_range():
    q = philox.exp1_stream(seed=7, step=3, n_pos=512, C=29)
    assert q.shape == (512, 29) and q.dtype == np.float32
    assert np.isfinite(q).all() and (q > 0).all() and q.max() < 17.0       # open interval: no 0, no inf
    # a draw depends on (seed, step, pos, class) only
    again = philox.exp1_stream(seed=7, step=3, n_pos=100, C=29)
    np.testing.assert_array_equal(q[:100], again)
    assert not np.array_equal(q, philox.exp1_stream(seed=8, step=3, n_pos=512, C=29))
    assert not np.array_equal(q, philox.exp1_stream(seed=7, step=4, n_pos=512, C=29))
    # classes 4k..4k+3 of a position share one Philox block
    w = philox.uniform_words(7, 3, 4, 29)
    blk = philox.philox4x32_10(np.arange(4), np.full(4, 3), np.full(4, 5), np.zeros(4), 7, 0)
    np.testing.assert_array_equal(w[:, 20:24], np.stack(blk, 1))
    # extreme words stay inside the open interval
    ext = philox.exp1_from_words(np.array([0, 1, 0xFFFFFFFF, 0xFFFFFF00], dtype=np.uint32))
    assert ext[0] > 0 and np.isfinite(ext).all() and ext[2] == ext[3] and abs(ext[2] - 16.635532) < 1e-4


def test_stream_is_exponential_ks():
    q = np.concatenate([philox.exp1_stream(seed=s, step=s * 3, n_pos=4096, C=29).ravel() for s in range(4)])
    res = stats.kstest(q.astype(np.float64), 'expon')
    assert res.pvalue > 1e-3, res
    assert abs(q.mean() - 1.0) < 5e-3 and abs(q.var() - 1.0) < 2e-2
    # the small draws decide races: the lower tail must be uniform in q (density 1 at 0)
    assert abs((q < 1e-3).mean() / 1e-3 - 1.0) < 0.1


def test_tokens_drawn_with_stream_follow_softmax_chi_square():
    """The race argmax(p / q) with the Philox stream draws class c with probability p_c: peaked and flat rows."""
    C, n = 29, 1 << 16
    g = torch.Generator().manual_seed(0)
    for scale in (0.5, 3.0):
        logit = torch.randn(C, generator=g) * scale
        logits = logit.view(1, C, 1).expand(1, C, n).contiguous()          # the same distribution at n positions
        q = torch.from_numpy(philox.exp1_stream(seed=11, step=2, n_pos=n, C=C))
        tok = osamp.sample_tokens(logits, q).view(-1).numpy()
        p = torch.softmax(logit.double(), 0).numpy()
        counts = np.bincount(tok, minlength=C).astype(np.float64)
        keep = p * n >= 5                                                  # chi-square validity
        chi2 = ((counts[keep] - p[keep] * n) ** 2 / (p[keep] * n)).sum()
        pval = stats.chi2.sf(chi2, int(keep.sum()) - 1)
        assert pval > 1e-3, (scale, chi2, pval)
