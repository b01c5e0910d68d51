"""CPU: the extended-code / exact-rule statement in the oracle (ko_ext_*; OUR definition -- the reference only builds
n = 2^m - 1, src/main.cpp:60, so the (128,64,22) curve is parity-unpinned) is pinned two ways: with ext = 0, rules = 0 it
IS the literal restatement of KanekoKernelProcessor::decode(answer, word, res), and on the small extended codes its
decisions are the exhaustive maximum-likelihood codeword (always with the exact rules, like the reference's own
near-ML behaviour with the reference rules)."""
import numpy as np
import pytest


def _all_codewords(o):
    k = o.k
    allinfo = ((np.arange(1 << k)[:, None] >> np.arange(k)) & 1).astype(np.uint8)
    cws = o.encode(allinfo)
    return np.concatenate([cws, cws.sum(1, keepdims=True) % 2], 1).astype(np.uint8)


@pytest.mark.parametrize("m,t,J,snr", [(4, 3, -1, 1.0), (5, 3, -1, 3.0), (6, 4, 9, 3.0)])
def test_ext0_rules0_is_the_literal_restatement(oracle_mod, m, t, J, snr):
    o = oracle_mod.Oracle(m, t, J)
    o.seed(5)
    _, _, y = o.gen_frames(snr, 400)
    d0, t0, c0, s0 = o.kaneko_decode(y)
    d1, t1, c1, s1, _ = o.ext_kaneko_decode(y, ext=0, rules=0)
    assert np.array_equal(d0, d1) and np.array_equal(t0, t1) and np.array_equal(c0, c1) and np.array_equal(s0, s1)


@pytest.mark.parametrize("m,t,snr,B", [(4, 3, 0.0, 6000), (4, 2, 1.0, 6000), (4, 1, 1.0, 6000), (5, 3, 1.0, 400)])
def test_extended_decisions_against_exhaustive_ml(oracle_mod, m, t, snr, B):
    """eBCH(16,5,8), (16,7,6), (16,11,4), (32,16,8): brute force over all 2^k codewords."""
    o = oracle_mod.Oracle(m, t)
    o.seed(7)
    info, cw, y = o.ext_gen_frames(snr, B)
    assert np.array_equal(cw[:, :-1], o.encode(info)) and np.array_equal(cw[:, -1], cw[:, :-1].sum(1) % 2)
    cws = _all_codewords(o)
    ml = cws[np.argmax(y @ (2.0 * cws - 1).T, 1)]
    d2, _, _, _, l2 = o.ext_kaneko_decode(y, ext=1, rules=2)
    assert np.array_equal(d2, ml), "exact rules must return the ML codeword"
    alpha = np.abs(2 * y / (1 / (10 ** 0.05 * 2 * o.k / (o.n + 1))))
    assert np.allclose(l2, (alpha * ((y > 0).astype(np.uint8) != d2)).sum(1), rtol=1e-12)
    d0, *_ = o.ext_kaneko_decode(y, ext=1, rules=0)
    assert (d0 == ml).all(1).mean() > 0.99   # the reference's rules are near-ML, not ML: same here
