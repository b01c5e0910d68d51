"""Authors the two extra polar specifications and their golden vectors (build container only; needs
oracle/_ref/libpolar_ref.so, i.e. the reference's vendored library compiled by oracle/Makefile):

    python tests/golden/make_polar_specs.py

  specs/polar_256_128_ebch16_dyn.spec.in   -- same frozen POSITIONS as polar_256_128_ebch16.spec.in, but 48 of the 128
      constraints are DYNAMIC (u_f = XOR of 1..3 earlier information symbols): exercises
      MixedKernelEncoder.cpp:148-159 and the cmask parity of ContinuePathsFrozen (KernelListEngine.cpp:8-41).
  specs/polar_240_114_ebch16_sp.spec.in    -- 8 shortened + 8 punctured symbols (MixedKernelEncoder.cpp:115-139,
      179-203): the polarising transform is lower triangular, so the last 8 codeword symbols vanish iff the last 8 input
      symbols are frozen to zero; 6 of the remaining constraints are dynamic.

The reference ships no specification at all: these are ours, the reference library only validates them
(its constructor parses them, its encoder throws on an invalid shortening) and produces the golden vectors
tests/golden/polar_vectors2.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py  # noqa: E402

SPEC_DIR = os.path.join(ROOT, "polar-codes-with-bch-kernel_b200", "specs")
KERNEL = os.path.join(SPEC_DIR, "ebch16.kernel")


def parse_frozen(path):
    tok = open(path).read().split()
    n0, k = 256, int(tok[1])
    pos = 8
    frozen = []
    for _ in range(n0 - k):
        w = int(tok[pos])
        frozen.append(int(tok[pos + w]))
        pos += 1 + w
    return sorted(frozen)


def write_spec(name, N, K, dmin, shortened, punctured, constraints):
    lines = [f"{N} {K} {dmin} 2 {len(shortened)} {len(punctured)}", "-@KERNEL@ -@KERNEL@"]
    if shortened:
        lines.append(" ".join(map(str, shortened)))
    if punctured:
        lines.append(" ".join(map(str, punctured)))
    for c in constraints:
        lines.append(f"{len(c)} " + " ".join(map(str, c)))
    with open(os.path.join(SPEC_DIR, name), "w") as f:
        f.write("\n".join(lines) + "\n")
    return "\n".join(lines).replace("@KERNEL@", KERNEL) + "\n"


def main():
    rng = np.random.default_rng(20261018)
    base_frozen = parse_frozen(os.path.join(SPEC_DIR, "polar_256_128_ebch16.spec.in"))
    fset = set(base_frozen)

    # ---- dynamic-frozen variant
    cons = []
    dyn_candidates = [f for f in base_frozen if sum(1 for i in range(f) if i not in fset) >= 3]
    dyn = set(rng.choice(dyn_candidates, 48, replace=False).tolist())
    for f in base_frozen:
        if f in dyn:
            earlier = [i for i in range(f) if i not in fset]
            w = int(rng.integers(1, 4))
            terms = sorted(rng.choice(earlier, w, replace=False).tolist())
            cons.append(terms + [f])
        else:
            cons.append([f])
    spec_dyn = write_spec("polar_256_128_ebch16_dyn.spec.in", 256, 128, 8, [], [], cons)

    # ---- shortened + punctured variant
    shortened = list(range(248, 256))
    punctured = [0, 1, 2, 3, 16, 32, 48, 64]
    fset2 = sorted(fset | set(shortened))
    unfrozen2 = [i for i in range(256) if i not in set(fset2)]
    # drop the six least reliable (lowest index) information symbols as dynamic frozen symbols of earlier ones
    cons2 = []
    extra_dyn = [i for i in unfrozen2 if i > 40][:6]
    fall = sorted(set(fset2) | set(extra_dyn))
    info_left = [i for i in range(256) if i not in set(fall)]
    for f in fall:
        if f in extra_dyn:
            earlier = [i for i in info_left if i < f]
            terms = sorted(rng.choice(earlier, min(2, len(earlier)), replace=False).tolist()) if earlier else []
            cons2.append(terms + [f])
        else:
            cons2.append([f])
    K2 = 256 - len(fall)
    spec_sp = write_spec("polar_240_%d_ebch16_sp.spec.in" % K2, 240, K2, 8, shortened, punctured, cons2)

    out = {}
    for tag, spec in (("dyn", spec_dyn), ("sp", spec_sp)):
        for (L, B, snr) in [(1, 48, 2.5), (8, 24, 2.0), (32, 8, 1.5)]:
            ref = oracle_py.PolarReference(spec, L)
            info = rng.integers(0, 2, (B, ref.K), dtype=np.uint8)
            cw = ref.encode(info)            # throws "Invalid shortening specification" if a shortened symbol is not zero
            sigma = np.sqrt(1 / (2 * (ref.K / ref.N) * 10 ** (snr / 10)))
            llr = (2 * ((1 - 2.0 * cw) + sigma * rng.standard_normal(cw.shape)) / sigma ** 2).astype(np.float32)
            cnt, inf, cwl, met = ref.decode(llr)
            k = f"{tag}_L{L}"
            out[k + "_info"] = np.packbits(info, axis=1)
            out[k + "_cw"] = np.packbits(cw, axis=1)
            out[k + "_llr"] = llr
            out[k + "_count"] = cnt
            out[k + "_inf"] = np.packbits(inf, axis=2)
            out[k + "_cwl"] = np.packbits(cwl, axis=2)
            out[k + "_metric"] = met
            print(tag, "L", L, "N K", ref.N, ref.K, "block errors", int((inf[:, 0, :] != info).any(1).sum()), "of", B)
    np.savez_compressed(os.path.join(HERE, "polar_vectors2.npz"), **out)
    print("written", os.path.getsize(os.path.join(HERE, "polar_vectors2.npz")), "bytes")


if __name__ == "__main__":
    main()
