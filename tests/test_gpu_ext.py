"""GPU: extended codes (n+1, k, 2t+2) -- eBCH(128,64,22) of BASELINE configs[4] and its small siblings -- and the exact
stopping rules, through pk_kaneko_create_ext, bit-exact against the oracle's statement of the same definition
(oracle/kaneko_oracle.c ko_ext_*; pinned by tests/test_ext_oracle.py) and against exhaustive ML where 2^k is small."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (m, t, J, rules, Eb/N0 dB, frames)
CASES = [
    (4, 3, -1, 0, 0.0, 4000), (4, 3, -1, 2, 1.0, 4000), (4, 2, -1, 0, 1.0, 3000), (4, 1, -1, 2, 2.0, 3000),
    (5, 3, -1, 0, 1.0, 600), (5, 3, -1, 2, 2.0, 600), (5, 2, -1, 0, 2.0, 1500), (5, 5, 12, 0, 2.0, 300),
    (6, 6, 9, 0, 2.0, 300), (6, 6, 12, 2, 3.5, 300), (6, 4, 9, 0, 2.0, 300), (6, 2, 9, 2, 3.0, 500),
    (7, 10, 9, 0, 4.5, 60), (7, 10, 15, 0, 4.0, 200), (8, 15, 9, 0, 5.5, 24),
]


@pytest.mark.parametrize("m,t,J,rules,snr,B", CASES)
@pytest.mark.parametrize("lut", [True, False])
def test_extended_replay_matches_oracle(pk, oracle_mod, m, t, J, rules, snr, B, lut):
    code = pk.Code(m, t, device=0)
    if lut and not code.uses_lut:
        pytest.skip("no lookup table for this code")
    code.set_lut(lut)
    o = oracle_mod.Oracle(m, t, J)
    o.seed(21)
    info, cw, y = o.ext_gen_frames(snr, B)
    kan = pk.Kaneko(code, J=J, extended=True, rules=rules, max_trials=1 << 20)
    assert kan.n == code.n + 1
    g_dec, g_tr, recs, tot = kan.decode(y)
    keep = (recs["flags"] & pk.PK_FLAG_TRUNCATED) == 0
    if m < 6 or (m == 6 and rules == 0):
        assert keep.all()
    else:   # frames whose search the CPU cannot finish are not compared (see test_gpu_parity_long.py); the exact rules ignore J
        keep &= g_tr <= (1 << 15)
        assert keep.sum() >= B // 2
    dec, tr, cmp_, sum_, lbest = o.ext_kaneko_decode(y[keep], ext=1, rules=rules)
    assert np.array_equal(g_tr[keep], tr), f"trial counts differ at {np.nonzero(g_tr[keep] != tr)[0][:5]}"
    assert np.array_equal(g_dec[keep], dec)
    d, c, s = pk.counters_from_recs(recs[keep], code.n + 1)
    assert np.array_equal(c, cmp_) and np.array_equal(s, sum_)
    # every decision is a codeword of the extended code: BCH part has zero syndrome, last position = its parity
    ans, ok = code.bch_decode(g_dec[keep][:, :-1])
    assert not ok.any()
    assert np.array_equal(g_dec[keep][:, -1], g_dec[keep][:, :-1].sum(1) % 2)


@pytest.mark.parametrize("m,t,snr,B", [(4, 3, 0.0, 20000), (4, 2, 1.0, 20000), (4, 1, 1.0, 20000), (5, 3, 1.0, 3000)])
def test_exact_rules_return_the_ml_codeword(pk, oracle_mod, m, t, snr, B):
    """eBCH(16,5,8) / (16,7,6) / (16,11,4) / (32,16,8): all 2^k codewords by brute force."""
    code = pk.Code(m, t, device=0)
    o = oracle_mod.Oracle(m, t)
    kan = pk.Kaneko(code, extended=True, rules=2)
    info, cw, y = kan.generate_frames(snr, 2, 5, 0, B)
    assert np.array_equal(cw[:, :-1], o.encode(info)) and np.array_equal(cw[:, -1], cw[:, :-1].sum(1) % 2)
    k = o.k
    allinfo = ((np.arange(1 << k)[:, None] >> np.arange(k)) & 1).astype(np.uint8)
    cws = o.encode(allinfo)
    cws = np.concatenate([cws, cws.sum(1, keepdims=True) % 2], 1).astype(np.uint8)
    ml = cws[np.argmax(y @ (2.0 * cws - 1).T, 1)]
    g_dec, g_tr, recs, tot = kan.decode(y)
    assert np.array_equal(g_dec, ml)
    ref_rules = pk.Kaneko(code, extended=True, rules=0)
    r_dec, *_ = ref_rules.decode(y)
    assert (r_dec == ml).all(1).mean() > 0.99
    # noise statistics at the extended code's rate k / (n+1)
    sigma = np.sqrt(1 / (10 ** (snr / 10) * 2 * o.k / (o.n + 1)))
    z = (y - (2.0 * cw - 1.0)) / sigma
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 8 * np.sqrt(2 / z.size)


@pytest.mark.parametrize("m,t,J,snr,B", [(4, 3, -1, 1.0, 30000), (6, 6, 15, 3.0, 4000), (7, 10, 15, 4.5, 2000)])
def test_extended_generation_mode_equals_replay(pk, m, t, J, snr, B):
    """eBCH(128,64,22) included: fused generate + decode + compare == replay of the dumped frames; split invariance."""
    code = pk.Code(m, t, device=0)
    kan = pk.Kaneko(code, J=J, extended=True, max_trials=1 << 22)
    gen, grecs = kan.run_frames(snr, 3, 11, 500, B, want_recs=True)
    info, cw, y = kan.generate_frames(snr, 3, 11, 500, B)
    dec, trials, recs, tot = kan.decode(y)
    assert np.array_equal(trials, grecs["trials"])
    be = (dec != cw).sum(1)
    ok = (recs["flags"] & pk.PK_FLAG_NO_DECISION) == 0
    assert np.array_equal(be[ok].astype(np.uint16), grecs["bit_errors"][ok])
    for k in ("frames", "trials", "cmp", "sum"):
        assert tot[k] == gen[k], k
    t1, _ = kan.run_frames(snr, 3, 11, 500, B // 3)
    t2, _ = kan.run_frames(snr, 3, 11, 500 + B // 3, B - B // 3)
    for k in ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum"):
        assert t1[k] + t2[k] == gen[k]
