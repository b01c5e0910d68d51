"""Generates tests/golden/polar_vectors.npz from the reference's vendored polar library compiled into
oracle/_ref/libpolar_ref.so (oracle/Makefile, oracle/polar_shim).  Build container only:

    python tests/golden/make_golden_polar.py

Code: specs/polar_256_128_ebch16.spec.in (two layers of the 16 x 16 extended-BCH kernel, our frozen set)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import oracle_py  # noqa: E402

SPEC_DIR = os.path.join(ROOT, "polar-codes-with-bch-kernel_b200", "specs")
spec = open(os.path.join(SPEC_DIR, "polar_256_128_ebch16.spec.in")).read().replace("@KERNEL@", os.path.abspath(os.path.join(SPEC_DIR, "ebch16.kernel")))
out = {}
rng = np.random.default_rng(2026)
for (L, B, snr) in [(1, 60, 2.0), (8, 30, 1.5), (32, 12, 1.0)]:
    ref = oracle_py.PolarReference(spec, L)
    info = rng.integers(0, 2, (B, ref.K), dtype=np.uint8)
    cw = ref.encode(info)
    sigma = np.sqrt(1 / (2 * (ref.K / ref.N) * 10 ** (snr / 10)))
    llr = (2 * ((1 - 2.0 * cw) + sigma * rng.standard_normal(cw.shape)) / sigma ** 2).astype(np.float32)
    cnt, inf, cwl, met = ref.decode(llr)
    k = f"L{L}"
    out[k + "_info"] = np.packbits(info, axis=1)
    out[k + "_cw"] = np.packbits(cw, axis=1)
    out[k + "_llr"] = llr
    out[k + "_count"] = cnt
    out[k + "_inf"] = np.packbits(inf, axis=2)
    out[k + "_metric"] = met
chan = (rng.standard_normal((64, 16)) * 3).astype(np.float32)
u = rng.integers(0, 2, (64, 16), dtype=np.uint8)
kl = np.stack([oracle_py.polar_ref_trellis_llrs(os.path.join(SPEC_DIR, "ebch16.kernel"), chan[b], u[b])[0] for b in range(64)])
np.savez_compressed(os.path.join(HERE, "polar_vectors.npz"), kernel_chan=chan, kernel_u=u, kernel_llr=kl, **out)
print("written", os.path.getsize(os.path.join(HERE, "polar_vectors.npz")), "bytes")
