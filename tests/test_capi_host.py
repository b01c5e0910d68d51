"""CPU: libpkb200.so loads, exports every symbol include/pk_capi.h declares, its host-side code
construction equals the oracle, and compute calls fail loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(pk):
    hdr = open(os.path.join(ROOT, "include", "pk_capi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(pk.lib, name), f"{name} declared in pk_capi.h but not exported by libpkb200.so"
    assert declared == set(pk.SYMBOLS), declared ^ set(pk.SYMBOLS)


@pytest.mark.parametrize("m,t", [(3, 1), (4, 1), (4, 2), (4, 3), (5, 1), (5, 2), (5, 3), (5, 5), (6, 2), (6, 4), (6, 6), (6, 11), (7, 10), (8, 15)])
def test_host_code_construction_matches_oracle(pk, oracle_mod, m, t):
    c = pk.Code(m, t, device=None)
    o = oracle_mod.Oracle(m, t)
    assert (c.n, c.k, c.d) == (o.n, o.k, 2 * t + 1) and np.array_equal(c.g, o.g)
    a1, l1 = c.tables()
    a2, l2 = o.tables()
    assert np.array_equal(a1, a2) and np.array_equal(l1, l2)
    assert np.array_equal(c.kernel_matrix(), o.make_matrix())


@pytest.mark.parametrize("m,t", [(3, 1), (4, 1), (4, 2), (4, 3), (5, 1), (5, 2), (5, 3), (6, 2)])
def test_coset_table_equals_oracle_decoder_on_every_coset(pk, oracle_mod, m, t):
    """The table the LUT kernels read is built by the host instance of the SAME pk_alg_decode<M,T>
    template the BM+Chien kernels run; here it is checked against the oracle's literal Sugiyama
    decoder for all 2^(n-k) syndromes (verdict and error positions)."""
    c = pk.Code(m, t, device=None)
    o = oracle_mod.Oracle(m, t)
    lut = c.coset_table()
    nk = c.n - c.k
    assert len(lut) == 1 << nk
    r = np.arange(len(lut), dtype=np.uint32)
    words = np.zeros((len(lut), c.n), np.uint8)
    for p in range(nk):
        words[:, p] = (r >> p) & 1
    ans, ok, *_ = o.bdd(words)
    assert np.array_equal(lut != 0xFFFF, ok.astype(bool))
    A = np.zeros((len(lut), c.n + 1), np.uint8)
    for j in range(t):
        A[np.arange(len(lut)), (lut.astype(np.uint32) >> (j * m)) & c.n] = 1
    good = ok == 1
    assert np.array_equal((words ^ A[:, : c.n])[good], ans[good])
    assert lut[0] == 0xFFFF  # zero syndrome is a decoding FAILURE in the reference


def test_invalid_arguments_are_rejected(pk):
    for (m, t) in [(4, 0), (4, 8), (2, 1), (9, 1), (5, 16)]:
        with pytest.raises(pk.PkError) as ei:
            pk.Code(m, t, device=None)
        assert ei.value.status == -1


def test_compute_without_gpu_fails_loudly(pk):
    """No CPU fallback: a host-only handle (or a machine without a GPU) refuses to compute."""
    c = pk.Code(4, 3, device=None)
    with pytest.raises(pk.PkError) as ei:
        c.encode(np.zeros((1, c.k), np.uint8))
    assert ei.value.status == -3
    with pytest.raises(pk.PkError):
        pk.Kaneko(c)
    if pk.lib.pk_device_count() == 0:
        with pytest.raises(pk.PkError) as ei:
            pk.Code(4, 3, device=0)
        assert ei.value.status == -3


def test_unsupported_code_is_reported_not_emulated(pk):
    c = pk.Code(8, 20, device=None)  # valid BCH code, no kernel instantiated
    assert c.n == 255
    with pytest.raises(pk.PkError):
        pk.Kaneko(c)


@pytest.mark.parametrize("m,t,kb", [(5, 5, 15), (5, 7, 20), (6, 3, 12), (6, 4, 18), (6, 5, 21), (6, 6, 27)])
def test_class_table_equals_algebraic_decoder(pk, m, t, kb):
    """The cyclic-class table (bitmap over syndrome classes + position table) the wide search of these codes reads
    gives the verdict and the error positions of pk_alg_decode<M,T> on random patterns of weight 0 .. t+3
    (pk_alg_decode itself is pinned to the oracle's Sugiyama decoder by the tests above and the GPU parity tests)."""
    c = pk.Code(m, t, device=None)
    bad, info = c.class_table_check(seed=7, ntrials=300000)
    assert bad == 0
    assert info["key_bits"] == kb == (c.n - c.k) - m
    # one class per cyclic orbit of the patterns with S_1 != 0, one key per pattern with S_1 = 0
    assert 0 < info["entries"] <= (1 << info["log2_slots"]) * 0.75
    assert info["bitmap_bytes"] >= (1 << (kb + 1)) // 8
