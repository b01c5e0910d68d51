"""GPU, generation mode: FER of the Philox-driven Monte-Carlo against EVERY published curve of the reference
(out/*.csv: 36 files with data, 11 Eb/N0 points each; numbers committed as tests/golden/ref_curves.json by
tests/golden/make_ref_curves.py) -- the `FER within the 95 % CI of the reference's curves` half of the north star.

A reference point is k_ref frame errors in n_ref frames (stop rule `count < p && countErr < e`, dataForPlot.cpp:43:
e = 100 or 1000, p = 10^6), i.e. a wide binomial interval; ours uses up to 3000 errors / 2*10^7 frames.  Per point the
two exact (Clopper-Pearson) intervals must intersect: at 95 % this fails by chance for about one point in twenty even for
identical decoders, so the per-curve test demands the 99.99 % intervals for every point and at most 3 of 11 misses at
95 %, and the summary test bounds the overall 95 % miss rate (396 points) at 9 % (expected 5 % + 3.6 sigma).
The average number of algebraic trials per frame (column 3/4 of the files) is heavy-tailed; it is compared loosely.

FIVE published files cannot be reproduced by the reference itself: for the three BCH(63,16,23) files (J = 9, 10, 11) and the
uncapped BCH(63,51,5) and BCH(63,39,9) files the reference COMPILED HERE (its own generator, its own decoder, 10^5..10^6 frames per point:
tests/golden/make_ref_fer_recomputed.py -> ref_fer_recomputed.json) gives the GPU's FER, not the published one (e.g.
(63,16,23) J=9 at 3.5 dB: published 6.66e-3, compiled reference 4.8e-3, GPU 4.85e-3) -- those files come from another
revision of the program (the uncapped n = 63 searches of this one stop early whenever T >= 32, the bound `(1 << T) - 1`
wraps: its uncapped (63,39,9) curve is WORSE than its J = 9 curve, 1.05e-1 vs 8.6e-2 at 2 dB, where the published uncapped file
says 6.4e-2; make_ref_fer_recomputed_budget.py).  For them the check runs against the recomputed reference points (same 95 % criterion)."""
import json
import os

import numpy as np
import pytest
from scipy.stats import beta

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CURVES = json.load(open(os.path.join(GOLD, "ref_curves.json")))
RECOMPUTED = json.load(open(os.path.join(GOLD, "ref_fer_recomputed.json")))
P_REF = 1_000_000
RESULTS = []   # (curve, ebn0, miss95, miss9999) of every point checked in this session


def cp(k, n, conf):
    """Clopper-Pearson interval of a binomial proportion."""
    a = 1 - conf
    lo = 0.0 if k == 0 else float(beta.ppf(a / 2, k, n - k + 1))
    hi = 1.0 if k == n else float(beta.ppf(1 - a / 2, k + 1, n - k))
    return lo, hi


def ref_point(fer, e):
    if fer <= 0:
        return 0, P_REF
    if e / fer <= P_REF * 1.0005:
        return e, max(e, int(round(e / fer)))
    return int(round(fer * P_REF)), P_REF


def _check_curve(pk, name, kan, points, trials_ref, J):
    """points: (Eb/N0, snr index, k_ref, n_ref); returns the report lines (raises on a 99.99 % miss)"""
    miss95, report = 0, []
    for (snr, si, k_ref, n_ref) in points:
        r = kan.run_point(snr, si, 20261018, 20_000_000, 3000)
        assert not (r["flags_or"] & pk.PK_FLAG_TRUNCATED)
        k, n = r["frame_errors"], r["frames"]
        out = {}
        for conf in (0.95, 0.9999):
            lo_r, hi_r = cp(k_ref, n_ref, conf)
            lo_o, hi_o = cp(k, n, conf)
            out[conf] = not (hi_o < lo_r or hi_r < lo_o)
        tr_ratio = (r["trials"] / n) / trials_ref[si] if trials_ref else 1.0
        tol = (0.6, 1.6) if J >= 0 else (0.2, 5.0)
        ok_tr = tol[0] <= tr_ratio <= tol[1]
        # An uncapped search of an n = 63 code can run to the wrapped bound, 2^31 - 1 trials; ONE such frame among
        # the reference's few hundred is most of its published mean (63_39_9 at 0.5 dB: 1.02e7 trials/frame over 254 frames =
        # one 2^31 frame + 1.7e6), so the mean is not a statistic there: FER only.
        if J < 0 and trials_ref and trials_ref[si] * n_ref > 0.25 * 2 ** 31:
            ok_tr = True
        RESULTS.append((name, snr, not out[0.95], not out[0.9999]))
        miss95 += not out[0.95]
        report.append(f"{snr:.1f} dB: ref {k_ref}/{n_ref} = {k_ref / n_ref:.3e}, ours {k}/{n} = {k / n:.3e}, trials x{tr_ratio:.2f}"
                      + ("" if out[0.95] else "  [outside 95 %]") + ("" if out[0.9999] else "  [OUTSIDE 99.99 %]") + ("" if ok_tr else "  [TRIALS]"))
        assert out[0.9999] and ok_tr, "\n".join(report)
    assert miss95 <= max(3, len(points) // 3), "\n".join(report)
    return report


@pytest.mark.parametrize("name", sorted(CURVES))
def test_fer_curve_within_reference_interval(pk, name):
    c = CURVES[name]
    code = pk.Code(c["m"], c["t"], device=0)
    assert (code.n, code.k, code.d) == (c["n"], c["k"], c["d"])
    kan = pk.Kaneko(code, J=c["J"])   # the reference's own loop bound, no safety cap: uncapped n = 63 searches reach 2^31 - 1 trials
    if name in RECOMPUTED:
        rc = RECOMPUTED[name]
        assert (rc["m"], rc["t"], rc["J"]) == (c["m"], c["t"], c["J"])
        pts = [(p["ebn0_db"], int(round(2 * p["ebn0_db"])), p["frame_errors"], p["frames"]) for p in rc["points"]]
        trials_ref = {int(round(2 * p["ebn0_db"])): p["trials"] / p["frames"] for p in rc["points"]}
        if rc.get("trials_unreliable"):   # recomputed with a time budget: the frames with the longest searches are missing
            trials_ref = None
        report = _check_curve(pk, name + " (recomputed reference)", kan, pts, trials_ref, c["J"])
    else:
        pts = []
        for si, snr in enumerate(c["ebn0_db"]):
            assert abs(snr - 0.5 * si) < 1e-9
            k_ref, n_ref = ref_point(c["fer"][si], c["e"])
            pts.append((snr, si, k_ref, n_ref))
        report = _check_curve(pk, name, kan, pts, c["trials"], c["J"])
    print("\n" + name + "\n  " + "\n  ".join(report))


def test_zz_overall_miss_rate_at_95_percent():
    """runs after the per-curve tests (alphabetical order inside the module)"""
    if len(RESULTS) < 100:
        pytest.skip("needs the per-curve tests of this module in the same session")
    rate = sum(r[2] for r in RESULTS) / len(RESULTS)
    assert rate <= 0.09, f"{rate:.3f} of {len(RESULTS)} points outside the 95 % intervals"
    assert not any(r[3] for r in RESULTS)
    print(f"\n{len(RESULTS)} reference points checked, {100 * rate:.1f} % outside the 95 % intervals (5 % expected by chance)")
