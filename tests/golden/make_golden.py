"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, built by oracle/Makefile
from /root/reference/src).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array below is an output of the reference's own code (ref_harness.cpp calls
KanekoKernelProcessor::decode, Decoder::decode, multiplyPolynomials, fun(), makeMatrix), on
frames drawn by the reference's own generator (std::default_random_engine, seeds given).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle_py  # noqa: E402

oracle_py.build(ref=True)

# (m, t, J, Eb/N0 dB, frames, seed)
KANEKO_CASES = [
    (4, 3, -1, 0.0, 300, 1), (4, 3, -1, 2.5, 300, 2), (4, 3, -1, 5.0, 300, 3),
    (4, 2, -1, 1.0, 200, 4), (4, 1, -1, 1.0, 200, 5),
    (5, 3, -1, 0.0, 40, 6), (5, 3, -1, 3.0, 200, 7), (5, 2, -1, 1.0, 100, 8), (5, 3, 9, 0.0, 100, 9),
    (6, 6, 9, 1.0, 40, 10), (6, 6, 15, 3.0, 30, 11), (6, 4, 9, 2.0, 60, 12), (6, 2, 9, 3.0, 100, 13),
    (6, 11, 9, 3.0, 30, 14), (7, 10, 9, 4.5, 12, 15), (8, 15, 9, 5.5, 6, 16),
]
out = {}
for (m, t, J, snr, B, seed) in KANEKO_CASES:
    r = oracle_py.Reference(m, t, J)
    r.seed(seed)
    info, cw, y = r.gen_frames(snr, B)
    dec, tr, cmp_, sum_ = r.kaneko_decode(y, answer=cw)
    key = f"kaneko_m{m}_t{t}_J{J}_snr{snr}"
    out[key + "_info"] = info
    out[key + "_cw"] = cw
    out[key + "_y"] = y
    out[key + "_decided"] = dec
    out[key + "_trials"] = tr
    out[key + "_cmp"] = cmp_
    out[key + "_sum"] = sum_
np.savez_compressed(os.path.join(HERE, "kaneko_frames.npz"), cases=np.array(KANEKO_CASES, dtype=np.float64), **out)

# ---- code construction KATs: g(x), (n,k), field tables; kernel matrices (matrixMain)
kat = {}
CODES = [(3, 1), (4, 1), (4, 2), (4, 3), (5, 1), (5, 2), (5, 3), (5, 5), (5, 7), (6, 2), (6, 4), (6, 6), (6, 11), (7, 10), (8, 15)]
for (m, t) in CODES:
    r = oracle_py.Reference(m, t)
    kat[f"g_m{m}_t{t}"] = r.g
    kat[f"nk_m{m}_t{t}"] = np.array([r.n, r.k])
    if t == 1:
        alog, log = r.tables()
        kat[f"alog_m{m}"] = alog
        kat[f"log_m{m}"] = log
        kat[f"kernel_m{m}"] = r.make_matrix()
np.savez_compressed(os.path.join(HERE, "code_kats.npz"), codes=np.array(CODES), **kat)

# ---- algebraic decoder vectors: words with 0 .. t+4 errors
bdd = {}
rng = np.random.default_rng(12345)
for (m, t, B) in [(4, 3, 3000), (5, 3, 3000), (6, 6, 2000), (7, 10, 500), (8, 15, 200)]:
    r = oracle_py.Reference(m, t)
    info = rng.integers(0, 2, (B, r.k), dtype=np.uint8)
    cw = r.encode(info)
    w = cw.copy()
    ne = rng.integers(0, t + 5, B)
    for f in range(B):
        w[f, rng.choice(r.n, ne[f], replace=False)] ^= 1
    ans, ok, synd, lam, lsz = r.bdd(w)
    ans[ok == 0] = 0
    bdd[f"bdd_m{m}_t{t}_words"] = np.packbits(w, axis=1)
    bdd[f"bdd_m{m}_t{t}_answers"] = np.packbits(ans, axis=1)
    bdd[f"bdd_m{m}_t{t}_ok"] = ok
    bdd[f"bdd_m{m}_t{t}_synd"] = synd.astype(np.uint16)
np.savez_compressed(os.path.join(HERE, "bdd_vectors.npz"), **bdd)

# ---- the reference's only input fixture, in/infile.txt: a (63,39,9) word + 63 samples
txt = open("/root/reference/in/infile.txt").read().split()
cw = np.array([int(v) for v in txt[:63]], np.uint8)[None]
y = np.array([float(v) for v in txt[63:126]])[None]
r = oracle_py.Reference(6, 4)
d2, tr2, c2, s2 = r.kaneko_decode(y, two_arg=True)
d3, tr3, c3, s3 = r.kaneko_decode(y, answer=cw)
# ---- fun(): the reference's CSV for `kaneko 4 3 f 3000 100` (default seed)
r = oracle_py.Reference(4, 3)
r.seed(1)
os.chdir("/tmp")
r.fun_csv("/tmp/_golden_fun", 3000, 100)
csv_text = open("/tmp/_golden_fun.csv").read()
np.savez_compressed(os.path.join(HERE, "infile_and_fun.npz"), infile_cw=cw, infile_y=y, infile_dec2=d2, infile_trials2=tr2,
                    infile_cmp2=c2, infile_sum2=s2, infile_dec3=d3, infile_trials3=tr3, infile_cmp3=c3, infile_sum3=s3,
                    fun_csv_m4_t3_p3000_e100=np.array(csv_text))
print("golden fixtures written:", [f for f in os.listdir(HERE) if f.endswith(".npz")])
