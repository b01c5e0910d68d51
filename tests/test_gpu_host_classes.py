"""GPU: the C++ drop-in class Decoder (host/bch_decoder.cpp; reference interface headers/Decoder.h:67-78) against the
oracle: findSyndromPoly, the incremental alterSyndromPoly (public syndromPoly / syndromPolySize fields) and decode()."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-codes-with-bch-kernel_b200")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("hd") / "libhd.so"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(PKG, "host"), "-I", os.path.join(ROOT, "include"), "-o", str(out),
           os.path.join(ROOT, "tests", "host_decoder_harness.cpp"), os.path.join(PKG, "host", "bch_decoder.cpp"), "-L", PKG, "-lpkb200", "-Wl,-rpath," + PKG]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    if r.returncode != 0:
        pytest.skip("cannot compile the harness here: " + r.stderr[-400:])
    return C.CDLL(str(out))


@pytest.mark.parametrize("m,t", [(4, 3), (5, 3), (6, 6), (7, 10)])
def test_decoder_class_syndromes_and_decode(pk, oracle_mod, harness, m, t):
    o = oracle_mod.Oracle(m, t)
    rng = np.random.default_rng(m * 10 + t)
    B = 300
    cw = o.encode(rng.integers(0, 2, (B, o.k), dtype=np.uint8))
    w = cw.copy()
    for f in range(B):
        w[f, rng.choice(o.n, rng.integers(0, t + 3), replace=False)] ^= 1
    sf = np.zeros((B, 2 * t), np.uint64); sa = np.zeros((B, 2 * t), np.uint64)
    zf = np.zeros(B, np.int64); za = np.zeros(B, np.int64)
    ok = np.zeros(B, np.uint8); ans = np.zeros((B, o.n), np.uint8)
    vp = C.c_void_p
    rc = harness.hd_check(m, t, w.ctypes.data_as(vp), C.c_long(B), sf.ctypes.data_as(vp), zf.ctypes.data_as(vp), sa.ctypes.data_as(vp),
                          za.ctypes.data_as(vp), ok.ctypes.data_as(vp), ans.ctypes.data_as(vp))
    assert rc == 0
    o_ans, o_ok, o_synd, _, _ = o.bdd(w)
    assert np.array_equal(sf, o_synd) and np.array_equal(sa, o_synd)
    size = np.array([max([j + 1 for j in range(2 * t) if o_synd[f, j]], default=0) for f in range(B)])
    assert np.array_equal(zf, size) and np.array_equal(za, size)
    assert np.array_equal(ok, o_ok) and np.array_equal(ans[o_ok == 1], o_ans[o_ok == 1])
