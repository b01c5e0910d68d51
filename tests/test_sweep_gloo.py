"""CPU: the multi-GPU sweep's host logic (frame dealing, stop rule in global frame order, counter
all-reduce) under torch.distributed gloo with world_size 2, against a 1-rank run of the same
deterministic stand-in engine."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polar-codes-with-bch-kernel_b200"))

REC = np.dtype([("trials", "<u4"), ("extra_cmp", "<u4"), ("extra_sum", "<u4"), ("bit_errors", "<u2"), ("flags", "u1"), ("reserved", "u1")])


class FakeEngine:
    """Deterministic function of the global frame index (like the Philox-indexed device engine)."""

    def run_frames(self, ebn0_db, snr_index, seed, first_frame, nframes, want_recs=False):
        idx = np.arange(first_frame, first_frame + nframes, dtype=np.uint64)
        h = (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed * 1000003 + snr_index * 7919)) >> np.uint64(40)
        recs = np.zeros(nframes, REC)
        recs["trials"] = 1 + (h % np.uint64(97)).astype(np.uint32)
        recs["extra_cmp"] = (h % np.uint64(5)).astype(np.uint32)
        recs["extra_sum"] = (h % np.uint64(3)).astype(np.uint32)
        err = (h % np.uint64(11 + 3 * snr_index)) == 0
        recs["bit_errors"] = np.where(err, 1 + (h % np.uint64(4)), 0).astype(np.uint16)
        recs["flags"] = np.where(err, 0x08, 0).astype(np.uint8) | ((h % np.uint64(2)).astype(np.uint8) & 1)
        n = 15
        tr = recs["trials"].astype(np.int64)
        run = tr - (recs["flags"] & 1)
        tot = dict(frames=nframes, frame_errors=int(err.sum()), bit_errors=int(recs["bit_errors"].sum()), trials=int(tr.sum()),
                   cmp=int((run * (n + 6) + recs["extra_cmp"]).sum()), sum=int((run * (n + 1) + recs["extra_sum"]).sum()))
        return tot, (recs if want_recs else None)


def _sequential(p, e, snr_index, seed):
    """fun()'s loop, frame by frame."""
    eng = FakeEngine()
    _, recs = eng.run_frames(0.0, snr_index, seed, 0, p, want_recs=True)
    import sweep

    cnt = errs = 0
    while cnt < p and (e <= 0 or errs < e):
        errs += int((recs["flags"][cnt] & 0x08) != 0)
        cnt += 1
    return sweep._totals_from_recs(recs[:cnt], 15)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    import sweep

    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = sweep.Comm()
    out = []
    for (p, e, si) in [(5000, 40, 0), (5000, 40, 3), (777, 0, 1), (3000, 10_000, 2), (100, 1, 0), (9000, 333, 4)]:
        out.append(sweep.run_point(FakeEngine(), 15, comm, 0.5 * si, si, 7, p, e, chunk=512))
    rows, raw = sweep.sweep(FakeEngine(), 15, comm, 600, 25, seed=3, max_snr=1.0, log=open(os.devnull, "w"))
    dist.destroy_process_group()
    q.put((rank, [o.tolist() for o in out], rows, raw.tolist()))


def test_world2_equals_sequential():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    res.sort()
    assert res[0][1:] == res[1][1:], "ranks disagree"
    cases = [(5000, 40, 0), (5000, 40, 3), (777, 0, 1), (3000, 10_000, 2), (100, 1, 0), (9000, 333, 4)]
    for got, (p, e, si) in zip(res[0][1], cases):
        assert got == _sequential(p, e, si, 7).tolist(), (p, e, si)
    # world 1 gives the same CSV rows (BER* cumulative across points)
    os.environ.update(RANK="0", WORLD_SIZE="1")
    import sweep

    rows1, raw1 = sweep.sweep(FakeEngine(), 15, sweep.Comm(), 600, 25, seed=3, max_snr=1.0, log=open(os.devnull, "w"))
    assert rows1 == res[0][2] and raw1.tolist() == res[0][3]
    assert len(rows1) == 3 and all(len(r.split(",")) == 6 for r in rows1)
