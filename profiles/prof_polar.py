"""Polar SC / SC-list decode throughput on one GPU (device-resident LLRs) next to the reference library on one core.
usage: python profiles/prof_polar.py L FRAMES SNR_DB [REF_FRAMES]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import torch

import pkb200

pk = pkb200.pk
L, B, snr = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
nref = int(sys.argv[4]) if len(sys.argv) > 4 else 0
spec = pk.load_spec()
p = pk.Polar(spec, L=L, device=0)
rng = np.random.default_rng(1)
info = rng.integers(0, 2, (B, p.K), dtype=np.uint8)
cw = p.encode(info)
sigma = np.sqrt(1 / (2 * (p.K / p.N) * 10 ** (snr / 10)))
llr = (2 * ((1 - 2.0 * cw) + sigma * rng.standard_normal(cw.shape)) / sigma ** 2).astype(np.float32)
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
d_llr = torch.from_numpy(llr).cuda()
d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
d_inf = torch.zeros((B, L, p.K), dtype=torch.uint8, device="cuda")
d_met = torch.zeros((B, L), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
best = 1e9
for r in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    p.decode_dev(d_llr.data_ptr(), B, d_cnt.data_ptr(), d_inf.data_ptr(), None, d_met.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
fer = float((d_inf[:, 0, :].cpu().numpy() != info).any(1).mean())
line = f"polar (256,128) 2x eBCH16, L={L}, {snr} dB, B={B}: {best:.3f} ms, {B / best * 1e3:.0f} frames/s, FER {fer:.4f}"
if nref:
    import oracle_py
    ref = oracle_py.PolarReference(spec, L)
    t0 = time.perf_counter()
    r_cnt, r_inf, r_cw, r_met = ref.decode(llr[:nref])
    dt = time.perf_counter() - t0
    same = np.array_equal(r_inf, d_inf[:nref].cpu().numpy()) and np.array_equal(r_met.view(np.uint32), d_met[:nref].cpu().numpy().view(np.uint32))
    line += f" | reference library, 1 core: {nref / dt:.1f} frames/s, identical lists+metrics on {nref} frames: {same}"
print(line)
