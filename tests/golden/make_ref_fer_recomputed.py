"""FER of the reference AS COMPILED HERE (oracle/_ref/libkaneko_ref*.so, the reference's own generator and decoder on all
host cores) at the points where its PUBLISHED curve and the GPU path disagree beyond chance -- the three BCH(63,16,23)
files and the uncapped BCH(63,51,5) file.  Result: the compiled reference reproduces the GPU numbers, not its published
ones (e.g. (63,16,23) J=9 at 3.5 dB: published 6.66e-3 = 100/15021, compiled reference 4.8e-3, GPU 4.85e-3), i.e. those
files come from another revision of the program.  Build container only (needs /root/reference):

    python tests/golden/make_ref_fer_recomputed.py     ->  tests/golden/ref_fer_recomputed.json"""
import json
import multiprocessing as mp
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

# curve -> (m, t, J, [(Eb/N0 dB, frames)])
PLAN = {
    "63_16_23_e": (6, 11, 9, [(1.0, 4000), (2.5, 16000), (3.5, 96000)]),
    "63_16_23_e10": (6, 11, 10, [(1.0, 4000), (2.5, 16000), (3.5, 96000)]),
    "63_16_23_e11": (6, 11, 11, [(1.0, 4000), (2.5, 16000), (3.5, 96000)]),
    "63_51_5": (6, 2, -1, [(3.0, 40000), (4.0, 160000), (4.5, 480000), (5.0, 960000)]),
}


def work(a):
    m, t, J, snr, B, seed = a
    import oracle_py

    r = oracle_py.Reference(m, t, J)
    r.seed(seed)
    errs = fr = tr = 0
    while fr < B:
        nb = min(2000, B - fr)
        _, cw, y = r.gen_frames(snr, nb)
        dec, trials, _, _ = r.kaneko_decode(y, answer=cw)
        errs += int((dec != cw).any(1).sum())
        fr += nb
        tr += int(trials.sum())
    return errs, fr, tr


if __name__ == "__main__":
    cores = len(os.sched_getaffinity(0))
    out = {}
    with mp.get_context("spawn").Pool(cores) as pool:
        for name, (m, t, J, pts) in PLAN.items():
            out[name] = {"m": m, "t": t, "J": J, "points": []}
            for snr, B in pts:
                res = pool.map(work, [(m, t, J, snr, B // cores, 500 + i) for i in range(cores)])
                e, f, tr = (sum(r[i] for r in res) for i in range(3))
                out[name]["points"].append({"ebn0_db": snr, "frame_errors": e, "frames": f, "trials": tr})
                print(name, snr, e, f, e / f, tr / f, flush=True)
    json.dump(out, open(os.path.join(HERE, "ref_fer_recomputed.json"), "w"), indent=1)
