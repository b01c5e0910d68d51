"""Decode ONE frame of a generated batch alone (latency of a single long search).
usage: python profiles/prof_one.py M T J SNR_DB FRAMES INDEX"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pkb200

pk = pkb200.pk
m, t, J, snr, B, idx = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
code = pk.Code(m, t, device=0)
kan = pk.Kaneko(code, J=J)
y = torch.empty((B, code.n), dtype=torch.float64, device="cuda")
kan.generate_frames_dev(snr, int(round(snr * 2)), 1, 0, B, y.data_ptr(), stream=st.cuda_stream)
yy = y[idx : idx + 1].contiguous()
dec = torch.zeros((1, code.n), dtype=torch.uint8, device="cuda")
tr = torch.zeros(1, dtype=torch.int32, device="cuda")
tot = torch.zeros(8, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    kan.decode_dev(yy.data_ptr(), 1, dec.data_ptr(), tr.data_ptr(), None, tot.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    print(f"frame {idx}: {int(tr[0])} trials, {e0.elapsed_time(e1):.3f} ms")
