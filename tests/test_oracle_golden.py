"""CPU: the oracle (our C restatement) against the committed golden vectors that were produced
by the compiled reference (tests/golden/make_golden.py) -- bit-exact."""
import io
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    z = np.load(os.path.join(G, "kaneko_frames.npz"))
    return [(int(m), int(t), int(J), float(s), int(B), int(seed)) for (m, t, J, s, B, seed) in z["cases"]]


@pytest.mark.parametrize("m,t,J,snr,B,seed", _cases())
def test_kaneko_decode_matches_golden(oracle_mod, m, t, J, snr, B, seed):
    z = np.load(os.path.join(G, "kaneko_frames.npz"))
    key = f"kaneko_m{m}_t{t}_J{J}_snr{snr}"
    o = oracle_mod.Oracle(m, t, J)
    # the frames themselves: restated libstdc++ RNG stream + encoder + AWGN
    o.seed(seed)
    info, cw, y = o.gen_frames(snr, B)
    assert np.array_equal(info, z[key + "_info"]) and np.array_equal(cw, z[key + "_cw"])
    assert np.array_equal(y.view(np.uint64), z[key + "_y"].view(np.uint64))
    assert np.array_equal(o.encode(z[key + "_info"]), z[key + "_cw"])
    dec, tr, cmp_, sum_ = o.kaneko_decode(z[key + "_y"])
    assert np.array_equal(dec, z[key + "_decided"])
    assert np.array_equal(tr, z[key + "_trials"])
    assert np.array_equal(cmp_, z[key + "_cmp"]) and np.array_equal(sum_, z[key + "_sum"])


def test_code_construction_kats(oracle_mod):
    z = np.load(os.path.join(G, "code_kats.npz"))
    for m, t in z["codes"]:
        o = oracle_mod.Oracle(int(m), int(t))
        assert np.array_equal(o.g, z[f"g_m{m}_t{t}"])
        assert [o.n, o.k] == list(z[f"nk_m{m}_t{t}"])
        if t == 1:
            alog, log = o.tables()
            assert np.array_equal(alog, z[f"alog_m{m}"]) and np.array_equal(log, z[f"log_m{m}"])
            assert np.array_equal(o.make_matrix(), z[f"kernel_m{m}"])
    # banner KATs observed from the reference CLI: `kaneko 4 3 s` prints this g and (15, 5, 7)
    o = oracle_mod.Oracle(4, 3)
    assert list(o.g) == [1, 1, 1, 0, 1, 1, 0, 0, 1, 0, 1] and (o.n, o.k) == (15, 5)
    for (m, t, nk) in [(5, 3, (31, 16)), (6, 6, (63, 30)), (7, 10, (127, 64)), (8, 15, (255, 139))]:
        o = oracle_mod.Oracle(m, t)
        assert (o.n, o.k) == nk
    # 15 x 15 nested-BCH kernel (`matrixMain 4`): its rows are rows 1..15 of the 16 x 16 extended kernel
    # without the all-ones first column, whose row weights SURVEY.md 8c lists as 1,2,2,2,2,4,4,4,4,6,6,8,8,8,8,16
    assert [int(w) + 1 for w in oracle_mod.Oracle(4, 1).make_matrix().sum(1)] == [2, 2, 2, 2, 4, 4, 4, 4, 6, 6, 8, 8, 8, 8, 16]


def test_bdd_matches_golden(oracle_mod):
    z = np.load(os.path.join(G, "bdd_vectors.npz"))
    for (m, t) in [(4, 3), (5, 3), (6, 6), (7, 10), (8, 15)]:
        o = oracle_mod.Oracle(m, t)
        w = np.unpackbits(z[f"bdd_m{m}_t{t}_words"], axis=1)[:, : o.n]
        ans, ok, synd, _, _ = o.bdd(w)
        ans[ok == 0] = 0
        assert np.array_equal(ok, z[f"bdd_m{m}_t{t}_ok"])
        assert np.array_equal(ans, np.unpackbits(z[f"bdd_m{m}_t{t}_answers"], axis=1)[:, : o.n])
        assert np.array_equal(synd.astype(np.uint16), z[f"bdd_m{m}_t{t}_synd"])
        # 1..t errors are always corrected, 0 errors report failure (zero-syndrome quirk)
        assert ok.mean() > 0.3


def test_infile_fixture_and_fun_csv(oracle_mod):
    """in/infile.txt (the reference's only input fixture) and the CSV of `kaneko 4 3 f 3000 100`."""
    z = np.load(os.path.join(G, "infile_and_fun.npz"))
    o = oracle_mod.Oracle(6, 4)
    d2, tr2, c2, s2 = o.kaneko_decode(z["infile_y"], two_arg=True)
    assert np.array_equal(d2, z["infile_dec2"]) and np.array_equal(d2, z["infile_cw"])  # prints "Ok"
    assert tr2[0] == z["infile_trials2"][0] and c2[0] == z["infile_cmp2"][0] and s2[0] == z["infile_sum2"][0]
    d3, tr3, c3, s3 = o.kaneko_decode(z["infile_y"])
    assert np.array_equal(d3, z["infile_dec3"]) and tr3[0] == z["infile_trials3"][0]
    o = oracle_mod.Oracle(4, 3)
    o.seed(1)
    rows, raw = o.fun(3000, 100)
    s = io.StringIO()
    for r in rows:
        s.write(",".join("%g" % v for v in r) + "\n")
    assert s.getvalue() == str(z["fun_csv_m4_t3_p3000_e100"])
    # the same stream reproduces the reference's published first FER value (out/15_5_7_e.csv:1 = 100/706)
    assert raw[0][0] == 706 and raw[0][1] == 100 and "%g" % rows[0][1] == "0.141643"


def test_std_sort_restated_with_ties(oracle_mod):
    """ko_std_sort_pairs == a stable sort for n <= 16 and a valid sort (same multiset, ascending) beyond."""
    o = oracle_mod.Oracle(4, 3)
    rng = np.random.default_rng(3)
    for n in (15, 16, 17, 31, 63, 255):
        for _ in range(20):
            keys = rng.integers(0, 8, n).astype(np.float64)
            k, idx = o.std_sort(keys)
            assert np.all(np.diff(k) >= 0) and sorted(idx) == list(range(n)) and np.array_equal(keys[idx], k)
            if n <= 16:
                assert np.array_equal(idx, np.argsort(keys, kind="stable"))
