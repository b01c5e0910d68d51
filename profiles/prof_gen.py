"""Generation-mode launches (fun()'s loop on the device: Philox frames -> decode -> count) of one code at one SNR point.
usage: python profiles/prof_gen.py M T J SNR_DB FRAMES [REPS]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pkb200

pk = pkb200.pk
m, t, J, snr, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
code = pk.Code(m, t, device=0)
kan = pk.Kaneko(code, J=J)
tot = torch.zeros(8, dtype=torch.int64, device="cuda")
best = 1e9
for r in range(reps):
    tot.zero_()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    kan.run_frames_dev(snr, int(round(2 * snr)), 1, 0, B, tot.data_ptr(), stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t_ = tot.cpu().numpy()
print(f"BCH({code.n},{code.k}) t={t} J={J} {snr} dB generation mode B={B}: {best:.3f} ms, {B / best * 1e3:.0f} frames/s, FER {t_[1] / t_[0]:.3e}, {t_[3] / t_[0]:.1f} trials/frame")
