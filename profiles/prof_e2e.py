"""Timeline of one end-to-end step of bench.py (async C-ABI batches, pinned host buffers) through torch.profiler (CUPTI)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import pkb200

pk = pkb200.pk
torch.cuda.set_device(0)
B = 16384
CODES = [(5, 3, -1), (6, 6, 15)]
if len(sys.argv) > 1 and sys.argv[1] == "only66":
    CODES = [(6, 6, 15)]
if len(sys.argv) > 1 and sys.argv[1] == "with43":
    CODES = [(4, 3, -1), (6, 6, 15)]
if len(sys.argv) > 1 and sys.argv[1] == "with64":
    CODES = [(6, 4, 12), (6, 6, 15)]
SYNC = len(sys.argv) > 2 and sys.argv[2] == "sync"
SNRS = [0.5 * i for i in range(11)]
codes = [pk.Code(m, t, device=0) for m, t, _ in CODES]
kans = [pk.Kaneko(c, J=J) for c, (_, _, J) in zip(codes, CODES)]
hy, hd, ht = [], [], []
for c, k in zip(codes, kans):
    y = torch.empty((len(SNRS), B, c.n), dtype=torch.float64, device="cuda")
    for si, s in enumerate(SNRS):
        k.generate_frames_dev(s, si, 1, 0, B, y[si].data_ptr())
    torch.cuda.synchronize()
    hy.append(y.cpu().pin_memory())
    hd.append(torch.zeros((len(SNRS), B, c.n), dtype=torch.uint8).pin_memory())
    ht.append(torch.zeros((len(SNRS), B), dtype=torch.int32).pin_memory())


def step():
    for si in range(len(SNRS)):
        for ci, k in enumerate(kans):
            if SYNC:
                k.decode_ptr(hy[ci][si].data_ptr(), B, hd[ci][si].data_ptr(), ht[ci][si].data_ptr())
            else:
                k.decode_async_ptr(hy[ci][si].data_ptr(), B, hd[ci][si].data_ptr(), ht[ci][si].data_ptr())
    for k in kans:
        k.wait()


step()
t0 = time.perf_counter()
for _ in range(3):
    step()
print("ms/step", (time.perf_counter() - t0) / 3 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t_first = ev[0].time_range.start
t_last = max(e.time_range.end for e in ev)
print("span ms", (t_last - t_first) / 1e3, "events", len(ev))
kinds = {}
for e in ev:
    k = e.name.split("<")[0][:40]
    kinds.setdefault(k, [0, 0.0])
    kinds[k][0] += 1
    kinds[k][1] += (e.time_range.end - e.time_range.start) / 1e3
for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} {v[0]:4d} {v[1]:9.3f} ms")
# busy time of kernels (union of intervals)
iv = sorted((e.time_range.start, e.time_range.end) for e in ev if "k_phase" in e.name)
busy, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_s is None:
        cur_s, cur_e = s, e
    elif s <= cur_e:
        cur_e = max(cur_e, e)
    else:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
busy += cur_e - cur_s
print("kernel busy (union) ms", busy / 1e3)
for e in ev:
    if "k_phase_b<6, 6" in e.name:
        print(f"   phase B<6,6> {(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms")
for e in ev[:0]:
    print(f"{(e.time_range.start - t_first) / 1e3:9.3f} {(e.time_range.end - e.time_range.start) / 1e3:8.3f} {e.name[:60]}")
