"""Profiling driver: a few replay-mode launches of one code at one SNR point.
usage: python profiles/prof_replay.py M T J SNR_DB FRAMES [LUT(0/1)] [REPS] [MAX_TRIALS]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pkb200

pk = pkb200.pk
m, t, J, snr, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
lut = int(sys.argv[6]) if len(sys.argv) > 6 else 1
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
max_trials = int(sys.argv[8]) if len(sys.argv) > 8 else 0
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
code = pk.Code(m, t, device=0)
if not lut and code.uses_lut:
    code.set_lut(False)
kan = pk.Kaneko(code, J=J, max_trials=max_trials)
y = torch.empty((B, code.n), dtype=torch.float64, device="cuda")
dec = torch.zeros((B, code.n), dtype=torch.uint8, device="cuda")
tr = torch.zeros(B, dtype=torch.int32, device="cuda")
tot = torch.zeros(8, dtype=torch.int64, device="cuda")
kan.generate_frames_dev(snr, int(round(snr * 2)), 1, 0, B, y.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
for r in range(reps):
    tot.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    kan.decode_dev(y.data_ptr(), B, dec.data_ptr(), tr.data_ptr(), None, tot.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    trials = int(tot[3].item())
    print(f"BCH({code.n},{code.k}) t={t} J={J} lut={code.uses_lut} {snr} dB B={B}: {ms:.3f} ms, {B / ms * 1e3:.0f} frames/s, "
          f"{trials / B:.1f} trials/frame, {trials / ms * 1e3:.3e} trials/s, max trials {int(tot[6].item())}, geometry {kan.geometry()}")
