"""GPU, generation mode: FER (and the average number of algebraic trials) of the Philox-driven
Monte-Carlo against the reference's published curves (out/*.csv, values quoted in BASELINE.md).

The reference points carry only e = 100 or 1000 error events (dataForPlot.cpp:43), so the check is:
our FER (>= 10x more frames) lies inside the 99.9 % two-sided binomial interval of the reference
point, widened by our own (much smaller) sampling error.  Full-size property checks (10^6 frames per
point: split invariance, counters) are in test_gpu_fullsize.py."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# out/15_5_7_new.csv (e = 1000), out/31_16_7.csv (e = 100), out/63_30_13_e15.csv (e = 100, J = 15)
REF = {
    (4, 3, -1, 1000): dict(
        fer=[0.170242, 0.130107, 0.0997009, 0.0637389, 0.0433426, 0.0299204, 0.017404, 0.00997079, 0.00511946, 0.00249565, 0.00101912],
        trials=[9.75672, 8.1478, 6.29182, 5.07426, 4.00121, 3.11555, 2.44913, 1.99461, 1.68766, 1.51712, 1.4479]),
    (5, 3, -1, 100): dict(
        fer=[0.2849, 0.262467, 0.154083, 0.0976562, 0.0660939, 0.034002, 0.0154154, 0.00677186, 0.00211341, 0.000681426, 0.000210735],
        trials=[1231.44, 936.034, 614.404, 255.855, 182.939, 97.185, 43.1989, 15.6042, 5.93472, 2.6529, 1.84171]),
    (6, 6, 15, 100): dict(
        fer=[0.37594, 0.248139, 0.132802, 0.0693481, 0.0269179, 0.0103402, 0.00231358, 0.000495233, 7.4e-05, 1.1e-05, 0.0],
        trials=[24465.3, 20655.5, 16494.2, 11572.5, 7504.13, 4161.11, 1644.66, 596.437, 164.899, 36.463, 7.16645]),
}


def _ref_interval(fer, e, z=3.3):
    """Interval for the TRUE error rate given the reference stopped at e errors (n = e / fer frames)."""
    if fer <= 0:
        return 0.0, 1.0
    n = max(e / fer, e)
    # Wilson interval
    c = fer + z * z / (2 * n)
    h = z * math.sqrt(fer * (1 - fer) / n + z * z / (4 * n * n))
    d = 1 + z * z / n
    return max(0.0, (c - h) / d), min(1.0, (c + h) / d)


@pytest.mark.parametrize("m,t,J,e,frames,points", [
    (4, 3, -1, 1000, 400_000, range(11)),
    (5, 3, -1, 100, 100_000, range(11)),
    (6, 6, 15, 100, 30_000, [0, 2, 4, 6, 8]),
])
def test_fer_inside_reference_interval(pk, m, t, J, e, frames, points):
    code = pk.Code(m, t, device=0)
    kan = pk.Kaneko(code, J=J)
    ref = REF[(m, t, J, e)]
    bad = []
    for si in points:
        snr = 0.5 * si
        tot, _ = kan.run_frames(snr, si, 20260101, 0, frames)
        assert tot["frames"] == frames
        fer = tot["frame_errors"] / frames
        lo, hi = _ref_interval(ref["fer"][si], e)
        ours = 3.3 * math.sqrt(max(fer * (1 - fer), 1.0 / frames) / frames)
        ok = (lo - ours) <= fer <= (hi + ours)
        # the average trial count is heavy-tailed; require the right order of magnitude only
        tr = tot["trials"] / frames
        ok_tr = 0.5 * ref["trials"][si] <= tr <= 2.0 * ref["trials"][si]
        if not (ok and ok_tr):
            bad.append((snr, fer, (lo, hi), tr, ref["trials"][si]))
    assert not bad, f"outside the reference interval: {bad}"
