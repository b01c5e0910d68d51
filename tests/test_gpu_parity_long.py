"""GPU parity at the settings that are MEASURED: long searches of the large codes
BCH(127,64,21) / BCH(255,139,31) and of the headline code BCH(63,30,13), all at J = 15, several SNR points,
hundreds to thousands of frames each -- against the COMPILED REFERENCE (oracle/_ref/libkaneko_ref_cap.so, the
reference's own KanekoKernelProcessor with the cap line enabled) fanned over every host core.  Where oracle/_ref is
absent the literal restatement (liboracle.so, byte-identical to the reference on the CPU suite) stands in.

The large codes spend up to 2^31 trials on a frame whose hard decision is far from every codeword (the bound stays
`(1 << n) - 1` until the first success), which no CPU finishes; the GPU decodes ALL frames and the reference re-decodes
the ones whose search fits the CPU budget (selected by the GPU's trial count; a wrong GPU count would make the
reference disagree on it, so the selection cannot hide a mismatch in the frames compared).
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_ENG = {}


def _worker(args):
    """One host core: the compiled reference (or the restatement) on a slice of frames."""
    m, t, J, y = args
    import oracle_py

    key = (m, t, J)
    if key not in _ENG:
        use_ref = oracle_py.ref_available(True)
        _ENG[key] = oracle_py.Reference(m, t, J) if use_ref else oracle_py.Oracle(m, t, J)
    dec, tr, cmp_, sum_ = _ENG[key].kaneko_decode(y)
    return dec, tr, cmp_, sum_


def _cpu_decode_all_cores(m, t, J, y, order):
    """Frames dealt to the cores longest-first (by the GPU's trial counts) so that the pool finishes together."""
    cores = max(1, len(os.sched_getaffinity(0)))
    idx = np.argsort(-order, kind="stable")
    parts = [idx[i::4 * cores] for i in range(4 * cores)]
    parts = [p for p in parts if len(p)]
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_worker, [(m, t, J, np.ascontiguousarray(y[p])) for p in parts], chunksize=1)
    n = y.shape[1]
    dec = np.zeros((len(y), n), np.uint8)
    tr = np.zeros(len(y), np.uint32)
    cmp_ = np.zeros(len(y), np.uint64)
    sum_ = np.zeros(len(y), np.uint64)
    for p, (d, a, c, s) in zip(parts, res):
        dec[p], tr[p], cmp_[p], sum_[p] = d, a, c, s
    return dec, tr, cmp_, sum_


# (m, t, J, Eb/N0 dB, frames drawn, frames compared at least, per-frame trial budget for the CPU, total trial budget)
LONG_CASES = [
    (6, 6, 15, 0.0, 2000, 2000, 1 << 15, 1 << 31),
    (6, 6, 15, 1.0, 2000, 2000, 1 << 15, 1 << 31),
    (6, 6, 15, 2.0, 2000, 2000, 1 << 15, 1 << 31),
    (7, 10, 15, 3.0, 2000, 500, 1 << 18, 16_000_000),
    (7, 10, 15, 4.0, 1500, 500, 1 << 18, 16_000_000),
    (7, 10, 15, 5.0, 1500, 500, 1 << 18, 16_000_000),
    (8, 15, 15, 3.5, 4000, 500, 1 << 16, 14_000_000),
    (8, 15, 15, 4.0, 2500, 500, 1 << 16, 6_000_000),
    (8, 15, 15, 4.5, 2000, 500, 1 << 16, 6_000_000),
]


@pytest.mark.parametrize("m,t,J,snr,B,need,per_frame,total", LONG_CASES)
def test_long_searches_match_compiled_reference(pk, oracle_mod, m, t, J, snr, B, need, per_frame, total):
    code = pk.Code(m, t, device=0)
    kan = pk.Kaneko(code, J=J, max_trials=1 << 22)      # bounds the GPU time of hopeless frames; such frames are not compared
    info, cw, y = kan.generate_frames(snr, int(round(2 * snr)), 77, 0, B)
    g_dec, g_tr, recs, tot = kan.decode(y)
    d, c, s = pk.counters_from_recs(recs, code.n)
    trunc = (recs["flags"] & pk.PK_FLAG_TRUNCATED) != 0
    ok = (~trunc) & (g_tr <= per_frame)
    sel = np.nonzero(ok)[0]
    sel = sel[np.cumsum(g_tr[sel].astype(np.int64)) <= total]
    assert len(sel) >= need, f"only {len(sel)} of {B} frames fit the CPU budget"
    r_dec, r_tr, r_cmp, r_sum = _cpu_decode_all_cores(m, t, J, y[sel], g_tr[sel])
    bad = np.nonzero(g_tr[sel] != r_tr)[0]
    assert not len(bad), f"trial counts differ at frames {sel[bad][:5]}: gpu {g_tr[sel][bad][:5]} ref {r_tr[bad][:5]}"
    assert np.array_equal(g_dec[sel], r_dec), f"decisions differ at frames {sel[np.nonzero((g_dec[sel] != r_dec).any(1))[0][:5]]}"
    assert np.array_equal(c[sel], r_cmp) and np.array_equal(s[sel], r_sum), "operation counters differ"
    # the sample really exercises long searches
    assert int(r_tr.max()) >= 1023
    print(f"\n[{code.n},{code.k}] J={J} {snr} dB: {len(sel)}/{B} frames compared, {int(r_tr.sum())} trials, max {int(r_tr.max())}; "
          f"{int(trunc.sum())} frames beyond 2^22 trials not compared")
