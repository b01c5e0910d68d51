"""GPU: polar rows (a15-a19) against the reference's vendored library (oracle/_ref/libpolar_ref.so): encoder
bit-exact, kernel LLRs bit-identical fp32, SC (L = 1) and SC-list (L = 8, 32) lists and path metrics identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(ref, B, snr, seed):
    rng = np.random.default_rng(seed)
    info = rng.integers(0, 2, (B, ref.K), dtype=np.uint8)
    cw = ref.encode(info)
    sigma = np.sqrt(1 / (2 * (ref.K / ref.N) * 10 ** (snr / 10)))
    y = (1 - 2.0 * cw) + sigma * rng.standard_normal(cw.shape)   # bit 0 -> +1 (Modem.h:64)
    return info, cw, (2 * y / sigma ** 2).astype(np.float32)     # LLR > 0 <=> bit 0


@pytest.fixture(scope="module")
def need_ref(oracle_mod):
    if not oracle_mod.polar_ref_available():
        pytest.skip("oracle/_ref/libpolar_ref.so not built")
    return oracle_mod


def test_encoder_matches_reference(pk, need_ref):
    spec = pk.load_spec()
    ref = need_ref.PolarReference(spec, 1)
    p = pk.Polar(spec, L=1, device=0)
    info = np.random.default_rng(1).integers(0, 2, (3000, ref.K), dtype=np.uint8)
    cw = p.encode(info)
    assert np.array_equal(cw, ref.encode(info))
    assert np.array_equal(p.encode(info[:100] ^ info[100:200]), cw[:100] ^ cw[100:200])   # linearity


def test_kernel_llrs_bit_identical(pk, need_ref):
    p = pk.Polar(pk.load_spec(), L=1, device=0)
    rng = np.random.default_rng(2)
    B = 300
    chan = (rng.standard_normal((B, 16)) * 3).astype(np.float32)
    u = rng.integers(0, 2, (B, 16), dtype=np.uint8)
    got = p.kernel_llrs(chan, u, layer=0)
    kfile = os.path.join(pk.SPEC_DIR, "ebch16.kernel")
    for b in range(B):
        want, _ = need_ref.polar_ref_trellis_llrs(kfile, chan[b], u[b])
        assert np.array_equal(got[b].view(np.uint32), want.view(np.uint32)), b


@pytest.mark.parametrize("L,B,snr", [(1, 400, 2.0), (1, 200, 0.5), (8, 150, 1.5), (32, 60, 1.0), (32, 40, 2.5)])
def test_sc_list_decoder_matches_reference(pk, need_ref, L, B, snr):
    spec = pk.load_spec()
    ref = need_ref.PolarReference(spec, L)
    p = pk.Polar(spec, L=L, device=0)
    info, cw, llr = _frames(ref, B, snr, 10 + L)
    r_cnt, r_inf, r_cw, r_met = ref.decode(llr)
    g_cnt, g_inf, g_cw, g_met = p.decode(llr)
    assert np.array_equal(g_cnt, r_cnt)
    assert np.array_equal(g_met.view(np.uint32), r_met.view(np.uint32)), "path metrics differ"
    assert np.array_equal(g_inf, r_inf) and np.array_equal(g_cw, r_cw)
    # the decoder finds the sent word more often as L grows; at least sanity-check it decodes something
    assert (g_inf[:, 0, :] == info).all(1).mean() > 0.2


def test_against_committed_golden_vectors(pk):
    """Same checks against tests/golden/polar_vectors.npz (made from the reference library by
    tests/golden/make_golden_polar.py) -- does not need oracle/_ref at run time."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "polar_vectors.npz"))
    spec = pk.load_spec()
    p1 = pk.Polar(spec, L=1, device=0)
    got = p1.kernel_llrs(z["kernel_chan"], z["kernel_u"], layer=1)
    assert np.array_equal(got.view(np.uint32), z["kernel_llr"].view(np.uint32))
    for L in (1, 8, 32):
        p = pk.Polar(spec, L=L, device=0)
        k = f"L{L}"
        info = np.unpackbits(z[k + "_info"], axis=1)[:, : p.K]
        assert np.array_equal(np.packbits(p.encode(info), axis=1), z[k + "_cw"])
        cnt, inf, cw, met = p.decode(z[k + "_llr"])
        assert np.array_equal(cnt, z[k + "_count"])
        assert np.array_equal(met.view(np.uint32), z[k + "_metric"].view(np.uint32))
        assert np.array_equal(np.packbits(inf, axis=2), z[k + "_inf"])


# ---- the dynamic-frozen and the shortened + punctured specifications (tests/golden/make_polar_specs.py)
SPECS2 = [("dyn", "polar_256_128_ebch16_dyn.spec.in"), ("sp", "polar_240_114_ebch16_sp.spec.in")]


@pytest.mark.parametrize("tag,name", SPECS2)
@pytest.mark.parametrize("L,B,snr", [(1, 300, 2.5), (8, 120, 2.0), (32, 40, 1.5)])
def test_dynamic_frozen_and_shortened_specs_match_reference(pk, need_ref, tag, name, L, B, snr):
    """MixedKernelEncoder.cpp:115-139 (Shorten), :148-159 (dynamic frozen symbols), :179-203 (LoadLLRs) and the
    dynamic-frozen branch of ContinuePathsFrozen, bit-exact against the reference library."""
    spec = pk.load_spec(name)
    ref = need_ref.PolarReference(spec, L)
    p = pk.Polar(spec, L=L, device=0)
    assert (p.N, p.K, p.N0) == (ref.N, ref.K, ref.N0)
    info, cw, llr = _frames(ref, B, snr, 40 + L)
    assert np.array_equal(p.encode(info), cw)
    r_cnt, r_inf, r_cw, r_met = ref.decode(llr)
    g_cnt, g_inf, g_cw, g_met = p.decode(llr)
    assert np.array_equal(g_cnt, r_cnt)
    assert np.array_equal(g_met.view(np.uint32), r_met.view(np.uint32)), "path metrics differ"
    assert np.array_equal(g_inf, r_inf) and np.array_equal(g_cw, r_cw)
    assert (g_inf[:, 0, :] == info).all(1).mean() > 0.5


@pytest.mark.parametrize("tag,name", SPECS2)
def test_extra_specs_against_committed_golden_vectors(pk, tag, name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "polar_vectors2.npz"))
    spec = pk.load_spec(name)
    for L in (1, 8, 32):
        p = pk.Polar(spec, L=L, device=0)
        k = f"{tag}_L{L}"
        info = np.unpackbits(z[k + "_info"], axis=1)[:, : p.K]
        assert np.array_equal(np.packbits(p.encode(info), axis=1), z[k + "_cw"])
        cnt, inf, cw, met = p.decode(z[k + "_llr"])
        assert np.array_equal(cnt, z[k + "_count"])
        assert np.array_equal(met.view(np.uint32), z[k + "_metric"].view(np.uint32))
        assert np.array_equal(np.packbits(inf, axis=2), z[k + "_inf"])
        assert np.array_equal(np.packbits(cw, axis=2), z[k + "_cwl"])


@pytest.mark.parametrize("name,L", [("polar_256_128_ebch16.spec.in", 1), ("polar_256_128_ebch16.spec.in", 8), ("polar_240_114_ebch16_sp.spec.in", 8),
                                    ("polar_256_128_ebch16_dyn.spec.in", 32)])
def test_generation_mode_equals_decode_of_the_dumped_frames(pk, name, L):
    """pk_polar_run_frames (generate + decode + compare on the device) against decode() of the frames it draws; the drawn
    codewords are the encoder's, the noise has the right statistics, and the totals do not depend on the split."""
    p = pk.Polar(pk.load_spec(name), L=L, device=0)
    B, snr, si = 3000, 2.0, 4
    info, cw, llr = p.generate_frames(snr, si, 7, 100, B)
    assert np.array_equal(p.encode(info), cw)
    assert 0.45 < info.mean() < 0.55
    sigma = np.sqrt(1 / (2 * (p.K / p.N) * 10 ** (snr / 10)))
    z = (llr.astype(np.float64) * sigma ** 2 / 2 - (1 - 2.0 * cw)) / sigma
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 8 * np.sqrt(2 / z.size)
    cnt, inf, _, _ = p.decode(llr)
    be = (inf[:, 0, :] != info).sum(1)
    tot = p.run_frames(snr, si, 7, 100, B)
    assert tot["frames"] == B and tot["frame_errors"] == int((be > 0).sum()) and tot["bit_errors"] == int(be.sum())
    t1, t2 = p.run_frames(snr, si, 7, 100, 1234), p.run_frames(snr, si, 7, 100 + 1234, B - 1234)
    for k in ("frames", "frame_errors", "bit_errors"):
        assert t1[k] + t2[k] == tot[k]


@pytest.mark.parametrize("name", ["polar_256_128_ebch16.spec.in", "polar_256_128_ebch16_dyn.spec.in", "polar_240_114_ebch16_sp.spec.in"])
@pytest.mark.parametrize("L,G", [(1, 1), (1, 2), (1, 4), (2, 2), (3, 2), (4, 1), (4, 4), (5, 4), (6, 1), (8, 1), (8, 2), (8, 4), (12, 2), (16, 1), (16, 2), (24, 1), (32, 1)])
def test_lanes_decoder_equals_warp_per_path_decoder(pk, name, L, G, monkeypatch):
    """k_polar_lanes (paths across lanes, G lanes per path) against k_polar_decode (one warp per path, pinned to the
    reference library above) on the same LLRs: list sizes, information vectors, codewords and fp32 metrics identical,
    for every lane split the dispatcher knows and list sizes that are not powers of two (they run in the next power of
    two of slots per frame); the frame count is not a multiple of the frames per warp."""
    spec = pk.load_spec(name)
    monkeypatch.setenv("PK_POLAR_LANES", "0")
    old = pk.Polar(spec, L=L, device=0)
    monkeypatch.setenv("PK_POLAR_LANES", "1")
    monkeypatch.setenv("PK_POLAR_LANES_G", str(G))
    new = pk.Polar(spec, L=L, device=0)
    B = 1500 // L + 7
    _, _, llr = new.generate_frames(1.5, 3, 11, 0, B)
    llr[5] = 0.0                      # all-zero LLRs: every decision is a tie
    llr[6, ::3] = 0.0
    o_cnt, o_inf, o_cw, o_met = old.decode(llr)
    n_cnt, n_inf, n_cw, n_met = new.decode(llr)
    assert np.array_equal(n_cnt, o_cnt)
    assert np.array_equal(n_met.view(np.uint32), o_met.view(np.uint32)), "path metrics differ"
    assert np.array_equal(n_inf, o_inf) and np.array_equal(n_cw, o_cw)


def _make_spec(tmp_path, sizes, K, seed, dyn=0):
    """A specification with the given kernel sizes per layer (eBCH kernels of size 8 / 16), K information symbols at the
    positions of largest index weight, the others frozen to zero (the first `dyn` frozen symbols after an information
    symbol get a dynamic constraint on earlier information symbols)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pkb200
    pk = pkb200.pk
    files = {}
    for n in set(sizes):
        Kn = pk.ebch_kernel({8: 3, 16: 4}[n])
        f = tmp_path / f"ebch{n}.kernel"
        f.write_text(f"{n}\n" + "\n".join(" ".join(str(int(v)) for v in r) for r in Kn) + "\n")
        files[n] = f
    N0 = int(np.prod(sizes))
    rng = np.random.default_rng(seed)
    score = np.array([bin(i).count("1") for i in range(N0)]) + rng.random(N0) * 0.5
    info = set(np.argsort(-score)[:K].tolist())
    lines = [f"{N0} {K} 1 {len(sizes)} 0 0", " ".join(f"-{files[n]}" for n in sizes)]
    seen_info = []
    for i in range(N0):
        if i in info:
            seen_info.append(i)
            continue
        if dyn > 0 and len(seen_info) >= 2:
            terms = sorted(rng.choice(seen_info, size=2, replace=False).tolist()) + [i]
            lines.append(f"{len(terms)} " + " ".join(map(str, terms)))
            dyn -= 1
        else:
            lines.append(f"1 {i}")
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("sizes,K,dyn", [((16,), 8, 0), ((8,), 4, 1), ((8, 16), 64, 0), ((16, 8), 70, 5), ((8, 8, 8), 256, 0), ((8, 8), 32, 6)])
@pytest.mark.parametrize("L", [1, 4, 7, 32])
def test_lanes_decoder_on_other_code_shapes(pk, oracle_mod, tmp_path, monkeypatch, sizes, K, dyn, L):
    """One, two and three layers, different kernels per layer, dynamic frozen symbols: k_polar_lanes against
    k_polar_decode and, where oracle/_ref is present, against the reference library itself."""
    spec = _make_spec(tmp_path, sizes, K, 7, dyn)
    monkeypatch.setenv("PK_POLAR_LANES", "0")
    old = pk.Polar(spec, L=L, device=0)
    monkeypatch.setenv("PK_POLAR_LANES", "1")
    monkeypatch.delenv("PK_POLAR_LANES_G", raising=False)
    new = pk.Polar(spec, L=L, device=0)
    B = 96 // min(L, 8) + 5
    _, _, llr = new.generate_frames(2.0, 4, 3, 0, B)
    llr[1] = 0.0
    o = old.decode(llr)
    n = new.decode(llr)
    assert np.array_equal(n[0], o[0]) and np.array_equal(n[3].view(np.uint32), o[3].view(np.uint32))
    assert np.array_equal(n[1], o[1]) and np.array_equal(n[2], o[2])
    if oracle_mod.polar_ref_available():
        ref = oracle_mod.PolarReference(spec, L)
        r = ref.decode(llr)
        assert np.array_equal(n[0], r[0]) and np.array_equal(n[3].view(np.uint32), r[3].view(np.uint32))
        assert np.array_equal(n[1], r[1]) and np.array_equal(n[2], r[2])


@pytest.mark.parametrize("n,seed", [(20, 1), (20, 2), (24, 2), (24, 5)])
@pytest.mark.parametrize("L", [1, 4, 8, 32])
def test_wide_random_kernels(pk, oracle_mod, tmp_path, monkeypatch, n, seed, L):
    """Random invertible n x n kernels (single layer) whose trellises have 2^9 .. 2^12 states: the lanes decoder takes more
    lanes per slot where the 16-bit row offsets need it (2^9: G = 2, 2^10: G = 4) and equals the warp-per-path decoder and
    the reference library; trellis tables beyond the shared memory of an SM (2^11 and more states) are refused loudly."""
    rng = np.random.default_rng(seed)
    while True:
        K = rng.integers(0, 2, (n, n), dtype=np.uint8)
        if round(abs(np.linalg.det(K.astype(float)))) % 2 == 1:
            break
    kf = tmp_path / "k.kernel"
    kf.write_text(f"{n}\n" + "\n".join(" ".join(str(int(v)) for v in r) for r in K) + "\n")
    spec = f"{n} {n // 2} 1 1 0 0\n-{kf}\n" + "".join("1 %d\n" % j for j in range(n - n // 2))
    bits = int(pk.Polar(spec, L=1, device=None).trellis_profile(0).max())
    monkeypatch.setenv("PK_POLAR_LANES", "0")
    try:
        old = pk.Polar(spec, L=L, device=0)
    except pk.PkError as exc:            # tables + 2^bits states x L paths do not fit the shared memory of an SM
        assert "shared memory" in str(exc) and (bits >= 11 or L == 32)
        pytest.skip(f"{bits} state bits at L = {L}: no decoder fits")
    monkeypatch.setenv("PK_POLAR_LANES", "1")
    new = pk.Polar(spec, L=L, device=0)
    _, _, llr = new.generate_frames(2.0, 4, 3, 0, 40)
    o, nn = old.decode(llr), new.decode(llr)
    assert np.array_equal(nn[0], o[0]) and np.array_equal(nn[3].view(np.uint32), o[3].view(np.uint32))
    assert np.array_equal(nn[1], o[1]) and np.array_equal(nn[2], o[2])
    if oracle_mod.polar_ref_available():
        r = oracle_mod.PolarReference(spec, L).decode(llr)
        assert np.array_equal(nn[0], r[0]) and np.array_equal(nn[3].view(np.uint32), r[3].view(np.uint32))
        assert np.array_equal(nn[1], r[1]) and np.array_equal(nn[2], r[2])
