"""Tail analysis of a replay launch: time of the whole batch, of the parked frames alone and of the single longest frame.
usage: python profiles/prof_tail.py M T J SNR_DB FRAMES [MAX_TRIALS]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pkb200

pk = pkb200.pk
m, t, J, snr, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
max_trials = int(sys.argv[6]) if len(sys.argv) > 6 else 0
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
code = pk.Code(m, t, device=0)
kan = pk.Kaneko(code, J=J, max_trials=max_trials)
y = torch.empty((B, code.n), dtype=torch.float64, device="cuda")
kan.generate_frames_dev(snr, int(round(snr * 2)), 1, 0, B, y.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()


def run(yy, reps=5):
    n = yy.shape[0]
    dec = torch.zeros((n, code.n), dtype=torch.uint8, device="cuda")
    tr = torch.zeros(n, dtype=torch.int32, device="cuda")
    tot = torch.zeros(8, dtype=torch.int64, device="cuda")
    best = 1e9
    for _ in range(reps):
        tot.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        kan.decode_dev(yy.data_ptr(), n, dec.data_ptr(), tr.data_ptr(), None, tot.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, tr.cpu().numpy()


ms, tr = run(y)
print(f"BCH({code.n},{code.k}) J={J} {snr} dB: all {B} frames {ms:.3f} ms; trials mean {tr.mean():.1f} max {tr.max()}; "
      f"frames > 64 trials: {(tr > 64).sum()}, > 8192: {(tr > 8192).sum()}")
idx = np.nonzero(tr > 64)[0]
if len(idx):
    ms2, _ = run(y[torch.from_numpy(idx).cuda()].contiguous())
    print(f"  parked frames alone ({len(idx)}): {ms2:.3f} ms")
    short = np.nonzero(tr <= 64)[0]
    ms3, _ = run(y[torch.from_numpy(short).cuda()].contiguous())
    print(f"  short frames alone ({len(short)}): {ms3:.3f} ms")
    big = int(np.argmax(tr))
    ms4, _ = run(y[big : big + 1].contiguous())
    print(f"  longest frame alone ({tr[big]} trials): {ms4:.3f} ms")
    for k in (8, 64, 148, 592):
        sel = idx[np.argsort(-tr[idx])][:k]
        ms5, _ = run(y[torch.from_numpy(sel).cuda()].contiguous())
        print(f"  {len(sel)} longest frames: {ms5:.3f} ms ({tr[sel].sum()} trials)")
    sel = idx[np.argsort(-tr[idx])][:12]
    for i in sel:
        msi, _ = run(y[int(i) : int(i) + 1].contiguous(), reps=3)
        print(f"    frame {i}: {tr[i]} trials, alone {msi:.3f} ms")
