"""GPU: Kaneko decoding as a kernel processor (pk_kproc_*): the LLR of kernel input `phase` from
maximum-likelihood Kaneko decodings of extended BCH codes equals the trellis processor's pStateMetric0[1] - pStateMetric0[0]
(out/external/TrellisKernelProcessor.cpp:292) -- against this library's Viterbi kernel, the committed golden vectors of the
reference's CTrellisKernelProcessor and, for the 64 x 64 kernel the trellis processor rejects (:71-72), against exhaustive
enumeration of the coset."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-5   # north_star: kernel LLRs within 1e-5 relative in the reference's float precision


def _close(a, b):
    scale = np.maximum(np.abs(a), np.abs(b)).max() + 1e-30
    return np.abs(a - b).max() <= REL * scale


@pytest.mark.parametrize("enum_dim", [-1, 3, 0])
def test_bridge_equals_trellis_processor_16(pk, enum_dim):
    """16 x 16 kernel, all 16 phases; enum_dim 0 sends every phase with a non-trivial tail through the Kaneko search."""
    kp = pk.KanekoKernelProc(4, enum_dim=enum_dim)
    assert kp.size == 16 and kp.mode[0] == 1
    if enum_dim == 0:
        # rows 1..10 leave an extended BCH code with t = 1, 2, 3 in the tail; the last rows (repetition-code tail) are enumerated
        assert (kp.mode[1:11] == 2).all() and set(kp.t[1:11]) == {1, 2, 3} and (kp.mode[11:] == 0).all()
    p = pk.Polar(pk.load_spec(), L=1, device=0)
    rng = np.random.default_rng(3)
    B = 400
    chan = (rng.standard_normal((B, 16)) * 3).astype(np.float32)
    u = rng.integers(0, 2, (B, 16), dtype=np.uint8)
    want = p.kernel_llrs(chan, u, layer=0)
    got, trunc = kp.kernel_llrs(chan, u)
    assert trunc == 0
    assert _close(got, want), np.abs(got - want).max()
    assert (got.view(np.uint32) == want.view(np.uint32)).mean() > 0.99   # in fact the same floats almost everywhere
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "polar_vectors.npz"))
    got, _ = kp.kernel_llrs(z["kernel_chan"], z["kernel_u"])
    assert _close(got, z["kernel_llr"])
    # the CKernProcLLR-shaped call: one phase, stride elements, [l][stride] layout
    ph = 6
    out, _ = kp.get_llrs(ph, np.ascontiguousarray(u.T), np.ascontiguousarray(chan.T))
    assert _close(out, want[:, ph])


def _brute(rows, chan, u, phase):
    """min-sum LLR of input `phase` by enumerating the coset (numpy; tails up to 2^16)."""
    l = len(rows)
    tail = rows[phase + 1:]
    k = len(tail)
    combos = ((np.arange(1 << k)[:, None] >> np.arange(k)) & 1).astype(np.uint8)
    span = (combos @ tail.astype(np.int64) % 2).astype(np.uint8) if k else np.zeros((1, l), np.uint8)
    hard = (chan < 0).astype(np.uint8)
    base = (u[:phase] @ rows[:phase].astype(np.int64) % 2).astype(np.uint8) if phase else np.zeros(l, np.uint8)
    M = []
    for v in (0, 1):
        o = base ^ (rows[phase] if v else 0)
        diff = span ^ o ^ hard
        M.append((diff * np.abs(chan)).sum(1).min())
    return np.float32(M[1] - M[0])


@pytest.mark.parametrize("m,sigma", [(5, 0.7), (6, 0.35)])
def test_bridge_32_and_64_against_enumeration(pk, m, sigma):
    """32 x 32 and 64 x 64 extended-BCH kernels: the last 17 phases (row tails of at most 2^16 words) against brute force,
    by two independent routes: the Kaneko search (enum_dim = 0) and the in-kernel enumeration (default).  A maximum-likelihood
    search over a (64, k) code is exponential in the number of unreliable positions: searches that exhaust the 2^22-trial
    budget are reported (`truncated`) and such phases are not compared."""
    l = 1 << m
    rows = pk.ebch_kernel(m)
    ka = pk.KanekoKernelProc(m, enum_dim=0, max_trials=1 << 22)
    kb = pk.KanekoKernelProc(m)
    assert ka.size == l and (ka.mode == 2).sum() >= l // 2
    rng = np.random.default_rng(m)
    B = 24
    u = rng.integers(0, 2, (B, l), dtype=np.uint8)
    cw = (u @ rows.astype(np.int64) % 2).astype(np.uint8)
    chan = (2 * ((1 - 2.0 * cw) + sigma * rng.standard_normal(cw.shape)) / sigma ** 2).astype(np.float32)
    checked = 0
    for ph in range(l - 17, l):
        known, ch = np.ascontiguousarray(u.T), np.ascontiguousarray(chan.T)
        ga, ta = ka.get_llrs(ph, known, ch)
        gb, tb = kb.get_llrs(ph, known, ch)
        want = np.array([_brute(rows, chan[b], u[b], ph) for b in range(B)], np.float32)
        if tb == 0:
            assert _close(gb, want), ("enumeration route", ph)
        if ta == 0:
            assert _close(ga, want), ("Kaneko route", ph)
            checked += 1
    assert checked >= (12 if m == 5 else 5)
    # the early phases (tails of dimension up to l-2) run through the Kaneko search as well; at this noise level the
    # genie-aided LLRs reproduce the inputs
    ga, ta = ka.kernel_llrs(chan[:8], u[:8])
    assert ((ga < 0).astype(np.uint8) == u[:8])[:, l // 2:].mean() > 0.95


def test_bridge_32_equals_trellis_processor(pk, tmp_path):
    """the 32 x 32 kernel still has a trellis processor: all 32 phases equal."""
    E = pk.ebch_kernel(5)
    kf = tmp_path / "ebch32.kernel"
    kf.write_text("32\n" + "\n".join(" ".join(str(int(v)) for v in r) for r in E) + "\n")
    spec = "32 16 1 1 0 0\n-%s\n" % kf + "".join("1 %d\n" % i for i in range(16))
    try:
        p = pk.Polar(spec, L=1, device=0)
    except pk.PkError as ex:
        pytest.skip(f"trellis processor does not fit: {ex}")
    kp = pk.KanekoKernelProc(5, enum_dim=0)
    rng = np.random.default_rng(9)
    B = 60
    chan = (rng.standard_normal((B, 32)) * 2.5).astype(np.float32)
    u = rng.integers(0, 2, (B, 32), dtype=np.uint8)
    want = p.kernel_llrs(chan, u, layer=0)
    got, trunc = kp.kernel_llrs(chan, u)
    assert trunc == 0 and _close(got, want), np.abs(got - want).max()


def test_trellis_processor_rejects_64_like_the_reference(pk, tmp_path):
    E = pk.ebch_kernel(6)
    kf = tmp_path / "ebch64.kernel"
    kf.write_text("64\n" + "\n".join(" ".join(str(int(v)) for v in r) for r in E) + "\n")
    spec = "64 32 1 1 0 0\n-%s\n" % kf + "".join("1 %d\n" % i for i in range(32))
    with pytest.raises(pk.PkError, match="too big"):
        pk.Polar(spec, L=1, device=None)
