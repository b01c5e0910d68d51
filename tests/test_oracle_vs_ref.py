"""CPU: the oracle against the compiled reference (oracle/_ref) on fresh seeded inputs.
Skipped where oracle/_ref is absent (it is built only where /root/reference exists)."""
import os

import numpy as np
import pytest


def _need_ref(oracle_mod, capped=False):
    if not oracle_mod.ref_available(capped):
        if os.path.isdir("/root/reference/src"):
            oracle_mod.build(ref=True)
        else:
            pytest.skip("oracle/_ref not built")


@pytest.mark.parametrize("m,t,J,snr,B", [(4, 3, -1, 0.0, 3000), (4, 3, -1, 4.0, 3000), (4, 1, -1, 2.0, 3000), (5, 3, -1, 1.0, 150),
                                           (5, 2, -1, 2.0, 600), (6, 6, 9, 2.0, 100), (6, 11, 9, 3.5, 40), (7, 10, 9, 5.0, 20)])
def test_frames_and_decode_identical(oracle_mod, m, t, J, snr, B):
    _need_ref(oracle_mod, J >= 0)
    o, r = oracle_mod.Oracle(m, t, J), oracle_mod.Reference(m, t, J)
    o.seed(77); r.seed(77)
    i1, c1, y1 = o.gen_frames(snr, B)
    i2, c2, y2 = r.gen_frames(snr, B)
    assert np.array_equal(i1, i2) and np.array_equal(c1, c2) and np.array_equal(y1.view(np.uint64), y2.view(np.uint64))
    a, b = o.kaneko_decode(y1), r.kaneko_decode(y2, answer=c2)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_quantised_inputs_with_ties(oracle_mod):
    """std::sort tie order (unstable for n > 16) is reproduced by the restated introsort."""
    for (m, t, J) in [(4, 3, -1), (5, 3, -1), (6, 6, 9)]:
        _need_ref(oracle_mod, J >= 0)
        o, r = oracle_mod.Oracle(m, t, J), oracle_mod.Reference(m, t, J)
        o.seed(5)
        _, cw, y = o.gen_frames(2.0, 800)
        yq = np.round(y, 1)
        yq[yq == 0] = 0.1
        for x, z in zip(o.kaneko_decode(yq), r.kaneko_decode(yq, answer=cw)):
            assert np.array_equal(x, z)


def test_fun_csv_identical(oracle_mod, tmp_path):
    _need_ref(oracle_mod)
    o, r = oracle_mod.Oracle(4, 2), oracle_mod.Reference(4, 2)
    o.seed(9); r.seed(9)
    rows, _ = o.fun(2000, 50)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref_rows = r.fun_csv(str(tmp_path / "f"), 2000, 50)
    finally:
        os.chdir(cwd)
    mine = "".join(",".join("%g" % v for v in row) + "\n" for row in rows)
    assert mine == open(tmp_path / "f.csv").read()
    assert ref_rows.shape == (11, 6)
