"""CPU: polar host logic -- extended-BCH kernel construction, specification parser and the per-phase kernel
trellises (state-complexity profile) against the survey's table and, where oracle/_ref/libpolar_ref.so exists,
against the reference's own CTrellisKernelProcessor."""
import os

import numpy as np
import pytest

# active-bit profile of the 16 x 16 eBCH kernel per phase (sections 0..16)
PROFILE = {
    0: "0 1 1 1 1 1 1 1 1 1 1 1 1 1 1 1 1", 15: "0 1 1 1 1 1 1 1 1 1 1 1 1 1 1 1 1",
    1: "0 1 2 2 2 2 2 2 2 2 2 2 2 2 2 2 1", 14: "0 1 2 2 2 2 2 2 2 2 2 2 2 2 2 2 1",
    2: "0 1 2 3 3 3 3 3 3 3 3 3 3 2 2 2 1", 3: "0 1 2 3 4 4 4 4 4 4 4 4 4 3 2 2 1",
    4: "0 1 2 3 4 5 5 5 5 5 5 5 5 4 3 2 1", 11: "0 1 2 3 4 5 5 5 5 5 5 5 5 4 3 2 1",
    5: "0 1 2 3 4 5 6 6 6 6 6 6 5 4 3 2 1", 6: "0 1 2 3 4 5 6 7 7 7 6 6 5 4 3 2 1",
    7: "0 1 2 3 4 5 6 7 8 8 7 6 5 4 3 2 1", 8: "0 1 2 3 4 5 6 7 8 8 7 6 5 4 3 2 1",
    9: "0 1 2 3 4 5 6 7 7 7 7 6 5 4 3 2 1", 10: "0 1 2 3 4 5 5 6 6 6 6 6 5 4 3 2 1",
    12: "0 1 2 3 4 4 4 4 4 4 4 4 4 4 3 2 1", 13: "0 1 2 2 3 3 3 3 3 3 3 3 3 3 3 2 1",
}


def test_ebch_kernel_matches_committed_file_and_survey(pk):
    E = pk.ebch_kernel(4)
    rows = ["".join(str(v) for v in r) for r in E]
    assert rows[0] == "1000000000000000" and rows[5] == "1110010000000000" and rows[15] == "1" * 16
    assert list(E.sum(1)) == [1, 2, 2, 2, 2, 4, 4, 4, 4, 6, 6, 8, 8, 8, 8, 16]
    txt = open(os.path.join(pk.SPEC_DIR, "ebch16.kernel")).read().split()
    assert int(txt[0]) == 16 and np.array_equal(np.array(txt[1:], np.uint8).reshape(16, 16), E)
    assert pk.ebch_kernel(3).shape == (8, 8)


def test_spec_parser_and_trellis_profile(pk):
    p = pk.Polar(pk.load_spec(), L=8, device=None)
    assert (p.N, p.K, p.N0, p.layers, p.L) == (256, 128, 256, 2, 8)
    prof = p.trellis_profile(0)
    assert prof.shape == (16, 17) and prof.max() == 8
    for ph, s in PROFILE.items():
        assert list(prof[ph]) == [int(x) for x in s.split()], ph
    # branch evaluations per kernel block: sum over phases of sum_j 2 * 2^ab[j]
    assert int((2 * 2 ** prof[:, :16].astype(np.int64)).sum()) == 11712


def test_spec_errors(pk):
    good = pk.load_spec()
    for bad in ("", "256 300 1 2 0 0", good.replace("256 128", "255 128"), good.replace("-/", "A /", 1)):
        with pytest.raises(pk.PkError):
            pk.Polar(bad, device=None)
    with pytest.raises(pk.PkError):
        pk.Polar(good, L=64, device=None)
    with pytest.raises(pk.PkError):   # no CPU compute path
        pk.Polar(good, device=None).encode(np.zeros((1, 128), np.uint8))


def test_trellis_profile_equals_reference_processor(pk, oracle_mod):
    if not oracle_mod.polar_ref_available():
        pytest.skip("oracle/_ref/libpolar_ref.so not built")
    _, ab = oracle_mod.polar_ref_trellis_llrs(os.path.join(pk.SPEC_DIR, "ebch16.kernel"), np.ones(16, np.float32), np.zeros(16, np.uint8))
    assert np.array_equal(ab, pk.Polar(pk.load_spec(), device=None).trellis_profile(0))


def test_dynamic_frozen_and_shortened_specs_parse(pk):
    """Host handles of the two extra specifications (tests/golden/make_polar_specs.py)."""
    p = pk.Polar(pk.load_spec("polar_256_128_ebch16_dyn.spec.in"), L=8, device=None)
    assert (p.N, p.K, p.N0, p.layers) == (256, 128, 256, 2)
    p = pk.Polar(pk.load_spec("polar_240_114_ebch16_sp.spec.in"), L=1, device=None)
    assert (p.N, p.K, p.N0, p.layers) == (240, 114, 256, 2)
    # a shortened index outside the code and a constraint that is not ascending are rejected like the reference does
    bad = pk.load_spec("polar_240_114_ebch16_sp.spec.in").replace("248 249", "248 300", 1)
    with pytest.raises(pk.PkError):
        pk.Polar(bad, device=None)
    dyn = pk.load_spec("polar_256_128_ebch16_dyn.spec.in").split("\n")
    i = next(k for k, ln in enumerate(dyn) if ln.split() and ln.split()[0] == "3")
    t = dyn[i].split()
    dyn[i] = " ".join([t[0], t[2], t[1], t[3]])
    with pytest.raises(pk.PkError):
        pk.Polar("\n".join(dyn), device=None)


def test_trellis_cost_equals_the_built_trellis(pk, tmp_path):
    """pk_kernel_trellis_cost (rank formula, what the permutation search scores) == the branch count of the trellises
    pk_polar_build_trellis constructs, for the natural order and for random linear column permutations."""
    E = pk.ebch_kernel(4)
    assert pk.kernel_trellis_cost(E) == (11712, 8)
    for trial in range(6):
        P, basis = pk.kernel_permute_columns(4, E, 3, trial)
        assert sorted(P.sum(0)) == sorted(E.sum(0)) and np.array_equal(np.sort(P, axis=1), np.sort(E, axis=1))
        cost, mb = pk.kernel_trellis_cost(P)
        kf = tmp_path / f"k{trial}.kernel"
        kf.write_text("16\n" + "\n".join(" ".join(str(int(v)) for v in r) for r in P) + "\n")
        spec = "16 8 1 1 0 0\n-%s\n" % kf + "".join("1 %d\n" % i for i in range(8))
        prof = pk.Polar(spec, L=1, device=None).trellis_profile(0)
        assert cost == int((2 * 2 ** prof[:, :16].astype(np.int64)).sum()) and mb == prof.max()


def test_swap_columns_restatement(pk):
    """swapColumns (root bchCoder.cpp:478-496): columns 0..2 stay, column i <- column j+1 with fieldElements[j] == i."""
    E = pk.ebch_kernel(4)
    rng = np.random.default_rng(4)
    fe = np.zeros(16, np.uint64)
    fe[2:15] = rng.permutation(np.arange(3, 16))    # fieldElements[j] = value carried by column j+1
    fe[15] = 2
    S = pk.kernel_swap_columns(4, fe, E)
    want = E.copy()
    for i in range(3, 16):
        j = next(j for j in range(2, 16) if fe[j] == i)
        want[:, i] = E[:, j + 1]
    assert np.array_equal(S[:, :3], E[:, :3]) and np.array_equal(S, want)
    with pytest.raises(pk.PkError):   # a value nobody carries
        pk.kernel_swap_columns(4, np.zeros(16, np.uint64), E)


def test_inplace_trellis_tables_equal_the_gather_form(pk, tmp_path):
    """The lanes decoder runs the kernel-trellis Viterbi in place on a renumbered trellis (pk_polar.h: one bit position of
    the state index per generator row); on the host, with integer costs, that recursion gives the gather-form recursion's
    LLR for every phase -- for the 16 x 16 and 8 x 8 eBCH kernels, column-permuted 16 x 16 kernels and
    random invertible kernels (rows that start and end in the same section, single-state phases)."""
    assert pk.Polar(pk.load_spec(), L=1, device=None).trellis_selfcheck(0, seed=3, ntests=200) == 8
    rng = np.random.default_rng(4)
    kernels = [pk.ebch_kernel(3)]   # (the 32 x 32 kernel's trellis has more than 2^14 states: no trellis processor, see pk_bridge.cu)
    E = pk.ebch_kernel(4)
    kernels += [E[:, rng.permutation(16)] for _ in range(6)]
    while len(kernels) < 20:
        n = int(rng.integers(2, 13))
        K = rng.integers(0, 2, (n, n), dtype=np.uint8)
        if round(abs(np.linalg.det(K.astype(float)))) % 2 == 1:   # invertible over GF(2)
            kernels.append(K)
    for i, K in enumerate(kernels):
        n = K.shape[0]
        kf = tmp_path / f"k{i}.kernel"
        kf.write_text(f"{n}\n" + "\n".join(" ".join(str(int(v)) for v in r) for r in K) + "\n")
        spec = f"{n} {n // 2} 1 1 0 0\n-{kf}\n" + "".join("1 %d\n" % j for j in range(n - n // 2))
        p = pk.Polar(spec, L=1, device=None)
        bits = p.trellis_selfcheck(0, seed=10 + i, ntests=40)
        assert bits == int(p.trellis_profile(0).max()), (i, bits)
