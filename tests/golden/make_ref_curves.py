"""Copies the NUMBERS of the reference's published curves (/root/reference/out/*.csv, 36 files with data) into
tests/golden/ref_curves.json so that the GPU tests can check FER against them on a box without /root/reference.
Build container only:   python tests/golden/make_ref_curves.py

Per file: (n, k, d) from the name -> (m, t); cap J from the suffix (none = HEAD / uncapped, _e = 9, _e10, _e11, _e15;
SURVEY.md section 6); error budget e of the run (1000 for *_new.csv, else 100 -- see below); rows as
published: 5 columns `EbN0,FER,trials,cmp,sum` or 6 columns `EbN0,FER,BER*,trials,cmp,sum`."""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CODES = {(15, 11, 3): (4, 1), (15, 7, 5): (4, 2), (15, 5, 7): (4, 3), (31, 21, 5): (5, 2), (31, 16, 7): (5, 3),
         (63, 51, 5): (6, 2), (63, 39, 9): (6, 4), (63, 30, 13): (6, 6), (63, 16, 23): (6, 11)}
SUFFIX = {"": -1, "new": -1, "e": 9, "e10": 10, "e11": 11, "e15": 15}

out = {}
for path in sorted(glob.glob("/root/reference/out/*.csv")):
    if os.path.getsize(path) == 0:
        continue
    name = os.path.basename(path)[:-4]
    parts = name.split("_")
    try:
        nkd = tuple(int(x) for x in parts[:3])
    except ValueError:
        continue
    if nkd not in CODES:
        continue
    suf = parts[3] if len(parts) > 3 else ""
    rows = [[float(v) for v in ln.strip().split(",") if v != ""] for ln in open(path) if ln.strip()]
    six = len(rows[0]) == 6
    m, t = CODES[nkd]
    # the stop rule's error budget: FER of the first row times an integer frame count = e (dataForPlot.cpp:43)
    # 15_7_5.csv (6 columns, no suffix) reads as e = 100 or e = 1000 alike from its six-digit FER column (0.236967 =
    # 100/422 = 1000/4220); its point-to-point scatter (0.130 at 1 dB, 0.127 at 1.5 dB, 0.062 at 2 dB, where the sibling
    # 15_7_5_new.csv has a smooth 0.126 / 0.093 / 0.062) is that of 100-event points, so e = 100.
    e = 1000 if suf == "new" else 100
    out[name] = {"m": m, "t": t, "n": nkd[0], "k": nkd[1], "d": nkd[2], "J": SUFFIX[suf], "e": e, "six_columns": six,
                 "ebn0_db": [r[0] for r in rows], "fer": [r[1] for r in rows], "ber_star": [r[2] for r in rows] if six else None,
                 "trials": [r[3 if six else 2] for r in rows], "cmp": [r[4 if six else 3] for r in rows], "sum": [r[5 if six else 4] for r in rows]}
json.dump(out, open(os.path.join(HERE, "ref_curves.json"), "w"), indent=1)
print(len(out), "curves")
for k, v in out.items():
    print(k, v["m"], v["t"], v["J"], v["e"], "frames@0dB ~", round(v["e"] / v["fer"][0]), "fer5dB", v["fer"][-1])
