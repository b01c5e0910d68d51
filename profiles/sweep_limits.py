"""Tuning helper: replay-mode timing of one code over several SNR points and phase-A limits.
usage: python profiles/sweep_limits.py M T J FRAMES LIMIT[,LIMIT..] SNR[,SNR..]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pkb200

pk = pkb200.pk
m, t, J, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
limits = [int(x) for x in sys.argv[5].split(",")]
snrs = [float(x) for x in sys.argv[6].split(",")]
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
code = pk.Code(m, t, device=0)
kan = pk.Kaneko(code, J=J)
ys = {}
for s in snrs:
    y = torch.empty((B, code.n), dtype=torch.float64, device="cuda")
    kan.generate_frames_dev(s, int(round(s * 2)), 1, 0, B, y.data_ptr(), stream=st.cuda_stream)
    ys[s] = y
dec = torch.zeros((B, code.n), dtype=torch.uint8, device="cuda")
tr = torch.zeros(B, dtype=torch.int32, device="cuda")
tot = torch.zeros(8, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
for lim in limits:
    kan.set_phase_a_limit(lim)
    line = [f"limit {lim:5d}:"]
    total = 0.0
    for s in snrs:
        best = 1e9
        for r in range(3):
            tot.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            kan.decode_dev(ys[s].data_ptr(), B, dec.data_ptr(), tr.data_ptr(), None, tot.data_ptr(), st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        total += best
        line.append(f"{s:.1f}dB {best:7.3f}ms ({int(tot[3].item()) / best * 1e3:.2e} tr/s)")
    print(" ".join(line), f"| sum {total:.3f} ms", flush=True)
