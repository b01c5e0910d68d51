"""Profiling driver: the SNR-independent search workload of SURVEY 8d -- pure-noise frames (Eb/N0 = -20 dB never decodes),
every search stopped after exactly MAX_TRIALS patterns.  usage: python profiles/prof_fixed.py M T FRAMES [MAX_TRIALS] [REPS]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pkb200

pk = pkb200.pk
m, t, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mt = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 15
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
code = pk.Code(m, t, device=0)
kan = pk.Kaneko(code, J=15, max_trials=mt)
tot = torch.zeros(8, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
for r in range(reps):
    tot.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    kan.run_frames_dev(-20.0, 0, 1, 0, B, tot.data_ptr(), None, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tr = int(tot[3].item())
    print(f"BCH({code.n},{code.k}) t={t}: {B} frames x {tr / B:.0f} patterns: {ms:.3f} ms, {tr / ms * 1e3:.3e} trials/s")
