"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs -- bit-exact decisions, trial counts and operation counters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (m, t, J, Eb/N0 dB, frames)
CASES = [
    (4, 3, -1, 0.0, 4000),
    (4, 3, -1, 5.0, 4000),
    (4, 2, -1, 1.0, 4000),
    (4, 1, -1, 1.0, 4000),
    (5, 3, -1, 0.0, 600),
    (5, 3, -1, 3.0, 3000),
    (5, 3, 9, 0.0, 1000),
    (5, 2, -1, 1.0, 2000),
    (6, 6, 9, 1.0, 200),
    (6, 6, 15, 3.0, 60),
    (6, 4, 9, 2.0, 300),
    (6, 2, 9, 3.0, 500),
    (6, 11, 9, 3.0, 100),
    (5, 5, 12, 2.0, 300),
    (5, 7, 10, 2.0, 200),
    (6, 3, 12, 3.0, 400),
    (6, 5, 12, 2.0, 200),
    (6, 6, 15, 0.0, 24),
    (7, 10, 9, 4.5, 40),
    (8, 15, 9, 5.5, 16),
]


def _oracle_frames(oracle_mod, m, t, J, snr, B, seed=11):
    o = oracle_mod.Oracle(m, t, J)
    o.seed(seed)
    info, cw, y = o.gen_frames(snr, B)
    dec, tr, cmp_, sum_ = o.kaneko_decode(y)
    return o, info, cw, y, dec, tr, cmp_, sum_


@pytest.mark.parametrize("m,t,J,snr,B", CASES)
@pytest.mark.parametrize("lut", [True, False])
def test_kaneko_replay_matches_oracle(pk, oracle_mod, m, t, J, snr, B, lut):
    code = pk.Code(m, t, device=0)
    if lut and not code.uses_lut:
        pytest.skip("no coset table for this code")
    code.set_lut(lut)
    o, info, cw, y, dec, tr, cmp_, sum_ = _oracle_frames(oracle_mod, m, t, J, snr, B)
    kan = pk.Kaneko(code, J=J)
    g_dec, g_tr, recs, tot = kan.decode(y)
    assert not (recs["flags"] & (pk.PK_FLAG_NO_DECISION | pk.PK_FLAG_TRUNCATED | pk.PK_FLAG_SORT_TIE)).any()
    _compare(pk, code, recs, g_dec, g_tr, tot, dec, tr, cmp_, sum_, B)


def _compare(pk, code, recs, g_dec, g_tr, tot, dec, tr, cmp_, sum_, B):
    assert np.array_equal(g_tr, tr), f"trial counts differ at frames {np.nonzero(g_tr != tr)[0][:5]}"
    assert np.array_equal(g_dec, dec), f"decisions differ at frames {np.nonzero((g_dec != dec).any(1))[0][:5]}"
    d, c, s = pk.counters_from_recs(recs, code.n)
    assert np.array_equal(d, tr.astype(np.uint64))
    assert np.array_equal(c, cmp_) and np.array_equal(s, sum_)
    assert tot["frames"] == B and tot["trials"] == int(tr.sum())
    assert tot["cmp"] == int(cmp_.sum()) and tot["sum"] == int(sum_.sum())
    assert tot["max_trials_seen"] == int(tr.max())


@pytest.mark.parametrize("m,t", [(3, 1), (4, 1), (4, 2), (4, 3), (5, 1), (5, 2), (5, 3), (5, 5), (5, 7), (6, 2), (6, 4), (6, 6), (6, 11), (7, 10), (8, 15)])
def test_bch_decode_matches_oracle(pk, oracle_mod, m, t):
    """Algebraic decoder alone (Decoder::findSyndromPoly + decode) on words with 0 .. t+4 errors."""
    code = pk.Code(m, t, device=0)
    o = oracle_mod.Oracle(m, t)
    rng = np.random.default_rng(5)
    B = 20000
    info = rng.integers(0, 2, (B, o.k), dtype=np.uint8)
    cw = o.encode(info)
    assert np.array_equal(code.encode(info), cw)
    ne = rng.integers(0, t + 5, B)
    w = cw.copy()
    for f in range(B):
        w[f, rng.choice(o.n, ne[f], replace=False)] ^= 1
    ans, ok, *_ = o.bdd(w)
    g_ans, g_ok = code.bch_decode(w)
    assert np.array_equal(g_ok, ok)
    good = ok == 1
    assert np.array_equal(g_ans[good], ans[good])
    assert not g_ans[~good].any()  # failed rows left untouched (zero-initialised here)


@pytest.mark.parametrize("m,t", [(4, 3), (5, 3)])
def test_bch_decode_exhaustive_cosets(pk, oracle_mod, m, t):
    """Every coset of the code through the BM + Chien kernel == the oracle's Sugiyama decoder."""
    code = pk.Code(m, t, device=0)
    o = oracle_mod.Oracle(m, t)
    nk = o.n - o.k
    r = np.arange(1 << nk, dtype=np.uint32)
    words = np.zeros((len(r), o.n), np.uint8)
    for p in range(nk):
        words[:, p] = (r >> p) & 1
    ans, ok, *_ = o.bdd(words)
    g_ans, g_ok = code.bch_decode(words)
    assert np.array_equal(g_ok, ok)
    assert np.array_equal(g_ans[ok == 1], ans[ok == 1])


@pytest.mark.parametrize("m,t,J,snr,B", [(4, 3, -1, 1.0, 20000), (5, 3, -1, 2.0, 4000), (6, 6, 9, 2.0, 500), (7, 10, 9, 4.5, 64)])
def test_generation_mode_matches_oracle_on_dumped_frames(pk, oracle_mod, m, t, J, snr, B):
    """The fused generate+encode+noise+decode+compare kernel == oracle on the frames it drew."""
    code = pk.Code(m, t, device=0)
    kan = pk.Kaneko(code, J=J)
    o = oracle_mod.Oracle(m, t, J)
    info, cw, y = kan.generate_frames(snr, 3, 2024, 1000, B)
    assert np.array_equal(o.encode(info), cw)
    # noise statistics: y - (2c-1) ~ N(0, sigma^2)
    sigma = np.sqrt(1 / (10 ** (snr / 10) * 2 * o.k / o.n))
    z = (y - (2.0 * cw - 1.0)) / sigma
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 8 * np.sqrt(2 / z.size)
    dec, tr, cmp_, sum_ = o.kaneko_decode(y)
    tot, recs = kan.run_frames(snr, 3, 2024, 1000, B, want_recs=True)
    assert np.array_equal(recs["trials"], tr)
    be = (dec != cw).sum(1)
    assert np.array_equal(recs["bit_errors"], be.astype(np.uint16))
    assert np.array_equal((recs["flags"] & pk.PK_FLAG_FRAME_ERROR) != 0, be > 0)
    assert tot["frames"] == B and tot["frame_errors"] == int((be > 0).sum()) and tot["bit_errors"] == int(be.sum())
    assert tot["trials"] == int(tr.sum()) and tot["cmp"] == int(cmp_.sum()) and tot["sum"] == int(sum_.sum())
    # split invariance: two half ranges give the same totals (what multi-GPU sharding relies on)
    t1, _ = kan.run_frames(snr, 3, 2024, 1000, B // 2)
    t2, _ = kan.run_frames(snr, 3, 2024, 1000 + B // 2, B - B // 2)
    for key in ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum"):
        assert t1[key] + t2[key] == tot[key]


def test_run_point_stop_rule(pk, oracle_mod):
    """fun()'s `count < p && countErr < e` evaluated in frame order (dataForPlot.cpp:43)."""
    code = pk.Code(4, 3, device=0)
    kan = pk.Kaneko(code)
    p, e = 50000, 100
    res = kan.run_point(0.0, 0, 7, p, e)
    assert res["frame_errors"] == e and res["frames"] < p
    _, recs = kan.run_frames(0.0, 0, 7, 0, res["frames"], want_recs=True)
    errs = (recs["flags"] & pk.PK_FLAG_FRAME_ERROR) != 0
    assert errs.sum() == e and errs[-1]  # the e-th error is the last frame counted
    assert res["trials"] == int(recs["trials"].sum())
    res2 = kan.run_point(5.0, 10, 7, 3000, 100)
    assert res2["frames"] == 3000 and res2["frame_errors"] < 100


@pytest.mark.parametrize("m,t,J,B", [(4, 3, -1, 3000), (5, 3, -1, 1500), (5, 2, -1, 2000), (6, 6, 9, 600), (6, 4, 9, 600), (7, 10, 9, 60)])
@pytest.mark.parametrize("lut", [True, False])
def test_tied_reliabilities_follow_std_sort(pk, oracle_mod, m, t, J, B, lut):
    """Quantised inputs: many equal |alpha|.  std::sort is unstable, the reference's order is whatever libstdc++'s
    introsort produces -- the kernels replay it (pk_stdsort.cuh); decisions, trial counts, counters stay identical."""
    code = pk.Code(m, t, device=0)
    if lut and not code.uses_lut:
        pytest.skip("no coset table for this code")
    code.set_lut(lut)
    o = oracle_mod.Oracle(m, t, J)
    o.seed(21)
    _, _, y = o.gen_frames(3.0, B)
    yq = np.round(y, 1)
    yq[yq == 0] = 0.1
    dec, tr, cmp_, sum_ = o.kaneko_decode(yq)
    kan = pk.Kaneko(code, J=J)
    g_dec, g_tr, recs, tot = kan.decode(yq)
    assert (recs["flags"] & pk.PK_FLAG_SORT_TIE).any()
    _compare(pk, code, recs, g_dec, g_tr, tot, dec, tr, cmp_, sum_, B)


def test_reference_infile_fixture(pk):
    """in/infile.txt of the reference (a (63,39,9) word and 63 six-digit samples), 3-argument decode: the result the
    compiled reference gives (tests/golden/infile_and_fun.npz) -- 15 trials and NOT the transmitted word."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "infile_and_fun.npz"))
    code = pk.Code(6, 4, device=0)
    dec, tr, recs, tot = pk.Kaneko(code).decode(z["infile_y"])
    assert np.array_equal(dec, z["infile_dec3"]) and tr[0] == z["infile_trials3"][0] == 15
    d, c, s = pk.counters_from_recs(recs, code.n)
    assert c[0] == z["infile_cmp3"][0] and s[0] == z["infile_sum3"][0]


@pytest.mark.parametrize("m,t,snr,B", [(4, 3, 1.0, 4000), (4, 2, 2.0, 3000), (5, 3, 2.0, 2000), (5, 2, 3.0, 2000), (6, 4, 4.0, 300), (6, 6, 4.0, 200)])
@pytest.mark.parametrize("lut", [True, False])
def test_two_argument_flavour_matches_oracle(pk, oracle_mod, m, t, snr, B, lut):
    """decode(word, res), the file-mode flavour (KanekoKernelProcessor.cpp:212-276): bound 1 << T, T unbounded until the
    first improvement, no cap J, 2n+1 extra counter units.  Frames where the reference's own unbounded calcT loop runs past
    its array (undefined behaviour there) carry PK_FLAG_REF_UNDEFINED and are left out of the comparison."""
    code = pk.Code(m, t, device=0)
    if lut and not code.uses_lut:
        pytest.skip("no coset table for this code")
    code.set_lut(lut)
    o = oracle_mod.Oracle(m, t)
    o.seed(33)
    _, _, y = o.gen_frames(snr, B)
    kan = pk.Kaneko(code, max_trials=1 << 16)
    kan.set_variant(True)
    g_dec, g_tr, recs, tot = kan.decode(y)
    keep = (recs["flags"] & (pk.PK_FLAG_REF_UNDEFINED | pk.PK_FLAG_TRUNCATED)) == 0
    assert keep.mean() > 0.9
    dec, tr, cmp_, sum_ = o.kaneko_decode(y[keep], two_arg=True)
    assert np.array_equal(g_tr[keep], tr)
    assert np.array_equal(g_dec[keep], dec)
    d, c, s = pk.counters_from_recs(recs, code.n)
    assert np.array_equal(c[keep], cmp_) and np.array_equal(s[keep], sum_)
    # back to the default flavour on the same handle
    kan.set_variant(False)
    dec3, tr3, *_ = o.kaneko_decode(y[:200])
    g3, gt3, *_ = kan.decode(y[:200])
    assert np.array_equal(g3, dec3) and np.array_equal(gt3, tr3)


def test_reference_infile_fixture_two_argument(pk):
    """in/infile.txt through the flavour main.cpp:158 really uses for it: 2048 trials and the transmitted word ("Ok")."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "infile_and_fun.npz"))
    code = pk.Code(6, 4, device=0)
    kan = pk.Kaneko(code)
    kan.set_variant(True)
    dec, tr, recs, tot = kan.decode(z["infile_y"])
    assert np.array_equal(dec, z["infile_dec2"]) and np.array_equal(dec, z["infile_cw"])
    assert tr[0] == z["infile_trials2"][0] == 2048
    d, c, s = pk.counters_from_recs(recs, code.n)
    assert c[0] == z["infile_cmp2"][0] and s[0] == z["infile_sum2"][0]


def test_async_batches_equal_blocking_calls(pk, oracle_mod):
    """pk_kaneko_decode_batch_async x N + pk_kaneko_wait == N blocking pk_kaneko_decode_batch calls (results and totals)."""
    import torch

    code = pk.Code(5, 3, device=0)
    kan = pk.Kaneko(code)
    o = oracle_mod.Oracle(5, 3)
    o.seed(3)
    ys, refs = [], []
    for snr in (1.0, 3.0, 5.0):
        _, _, y = o.gen_frames(snr, 3000)
        ys.append(y)
        refs.append(kan.decode(y))
    hy = [torch.from_numpy(y).pin_memory() for y in ys]
    hd = [torch.zeros((3000, code.n), dtype=torch.uint8).pin_memory() for _ in ys]
    ht = [torch.zeros(3000, dtype=torch.int32).pin_memory() for _ in ys]
    for y, d, t in zip(hy, hd, ht):
        kan.decode_async_ptr(y.data_ptr(), 3000, d.data_ptr(), t.data_ptr())
    with pytest.raises(pk.PkError):   # a blocking call may not be mixed into pending asynchronous batches
        kan.decode(ys[0])
    tot = kan.wait()
    for (dec, tr, recs, rt), d, t in zip(refs, hd, ht):
        assert np.array_equal(d.numpy(), dec) and np.array_equal(t.numpy().astype(np.uint32), tr)
    for key in ("frames", "trials", "cmp", "sum"):
        assert tot[key] == sum(r[3][key] for r in refs)
    assert kan.wait()["frames"] == 0   # nothing pending
    dec, tr, *_ = kan.decode(ys[1])    # blocking calls work again
    assert np.array_equal(dec, refs[1][0])


def test_uncapped_search_with_huge_frames(pk, oracle_mod):
    """Uncapped BCH(31,16,7) at 0 dB: a few frames keep 2^16 .. 2^19 patterns after their last improvement; they travel
    through the separate "huge" list of the parked frames (always searched by a whole CTA) and still match the oracle."""
    code = pk.Code(5, 3, device=0)
    o, info, cw, y, dec, tr, cmp_, sum_ = _oracle_frames(oracle_mod, 5, 3, -1, 0.0, 2500, seed=42)
    assert (tr >= 65536 + 256).any(), "sample has no huge frame: pick another seed"
    kan = pk.Kaneko(code)
    g_dec, g_tr, recs, tot = kan.decode(y)
    _compare(pk, code, recs, g_dec, g_tr, tot, dec, tr, cmp_, sum_, 2500)
