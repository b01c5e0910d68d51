"""Multi-GPU Monte-Carlo sweep: the reference's fun() (src/dataForPlot.cpp:16-115) with the
frames of every SNR point sharded over ranks (one process per GPU, torchrun + NCCL).

Frames are identified by a global index (the Philox counter), so the result of a point does
not depend on the number of GPUs.  Work is dealt in rounds: round r covers global frames
[r*W*C, (r+1)*W*C) and rank k decodes [r*W*C + k*C, r*W*C + (k+1)*C).

  * e <= 0 (fixed p frames, BASELINE configs[1]): every rank decodes its share and the per-point
    counters are combined by ONE all-reduce (sum of 6 x int64 + max + or).
  * finite e: fun()'s stop rule `count < p && countErr < e` (dataForPlot.cpp:43) is applied in
    global frame order: after each round the ranks exchange their error counts (one small
    all-gather), the rank that holds the e-th error cuts its chunk at that frame, later ranks
    drop theirs.  The totals equal a sequential run over the same frames.

The engine only has to offer run_frames(ebn0_db, snr_index, seed, first_frame, nframes,
want_recs) -> (totals dict, records) -- the GPU engine is pk.Kaneko; the CPU tests plug in a
deterministic stand-in to exercise the host logic under gloo.
"""
import os
import sys
import time

import numpy as np

FIELDS = ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum")
FLAG_EARLY, FLAG_ERR = 0x01, 0x08


def _totals_from_recs(recs, n):
    tr = recs["trials"].astype(np.int64)
    run = tr - (recs["flags"] & FLAG_EARLY).astype(np.int64)
    return np.array([len(recs), int(((recs["flags"] & FLAG_ERR) != 0).sum()), int(recs["bit_errors"].astype(np.int64).sum()),
                     int(tr.sum()), int((run * (n + 6) + recs["extra_cmp"]).sum()), int((run * (n + 1) + recs["extra_sum"]).sum())],
                    dtype=np.int64)


class Comm:
    """torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests); world 1 needs no torch."""

    def __init__(self, device=None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            assert dist.is_initialized(), "init_process_group first"
            self.dist = dist

    def allreduce_sum(self, vec):
        if self.world == 1:
            return vec
        import torch

        t = torch.from_numpy(np.ascontiguousarray(vec, np.int64))
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allgather_int(self, value):
        if self.world == 1:
            return np.array([value], np.int64)
        import torch

        t = torch.tensor([value], dtype=torch.int64, device=self.device if self.device is not None else "cpu")
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.array([int(o.item()) for o in out], np.int64)


class PolarEngine:
    """pk.Polar behind the engine interface: the simulator loop of the reference's vendored library (Simulator.cpp:139-335)
    for a mixed-kernel polar code -- generate, encode, AWGN, SC / SC-list decode, compare, all on the device
    (pk_polar_run_frames).  With want_recs the frames are drawn to the host, decoded through the batch API and compared
    there, which yields the per-frame error flags the stop rule needs (same frames: the Philox counter is the frame index)."""

    REC = np.dtype([("trials", "<u4"), ("bit_errors", "<u2"), ("flags", "<u2"), ("extra_cmp", "<u4"), ("extra_sum", "<u4")])

    def __init__(self, polar):
        self.p = polar

    def run_frames(self, ebn0_db, snr_index, seed, first_frame, nframes, want_recs=False):
        if not want_recs:
            tot = self.p.run_frames(ebn0_db, snr_index, seed, first_frame, nframes)
            return {f: int(tot.get(f, 0)) for f in FIELDS}, None
        info, _, llr = self.p.generate_frames(ebn0_db, snr_index, seed, first_frame, nframes)
        cnt, inf, _, _ = self.p.decode(llr)
        be = (inf[:, 0, :] != info).sum(1)
        be[cnt < 1] = self.p.K
        recs = np.zeros(nframes, self.REC)
        recs["bit_errors"] = be
        recs["flags"] = np.where(be > 0, FLAG_ERR, 0)
        return {"frames": nframes, "frame_errors": int((be > 0).sum()), "bit_errors": int(be.sum()), "trials": 0, "cmp": 0, "sum": 0}, recs


def run_point(engine, n, comm, ebn0_db, snr_index, seed, p, e, chunk=1 << 16):
    """One SNR point; returns int64[6] totals (FIELDS order), identical on every rank."""
    W, k = comm.world, comm.rank
    local = np.zeros(6, np.int64)
    if e <= 0:
        # fixed frame count: contiguous share per rank, one all-reduce per point
        per = (p + W - 1) // W
        first, cnt = k * per, max(0, min(per, p - k * per))
        if cnt > 0:
            tot, _ = engine.run_frames(ebn0_db, snr_index, seed, first, cnt, want_recs=False)
            local += np.array([tot[f] for f in FIELDS], np.int64)
        return comm.allreduce_sum(local)
    done, errs, c = 0, 0, min(chunk, 4096)
    while done < p and errs < e:
        span = min(W * c, p - done)
        first = done + min(k * c, span)
        cnt = max(0, min(c, span - k * c))
        recs = None
        my_err = 0
        if cnt > 0:
            _, recs = engine.run_frames(ebn0_db, snr_index, seed, first, cnt, want_recs=True)
            my_err = int(((recs["flags"] & FLAG_ERR) != 0).sum())
        all_err = comm.allgather_int(my_err)
        before = errs + int(all_err[:k].sum())      # errors in frames that precede this rank's chunk
        if cnt > 0 and before < e:
            if before + my_err < e:
                local += _totals_from_recs(recs, n)
            else:                                    # the e-th error lies in this chunk: cut right after it
                is_err = (recs["flags"] & FLAG_ERR) != 0
                cut = int(np.nonzero(np.cumsum(is_err) == (e - before))[0][0]) + 1
                local += _totals_from_recs(recs[:cut], n)
        errs += int(all_err.sum())
        done += span
        c = min(c * 2, chunk)
    return comm.allreduce_sum(local)


def format_row(stnr, tot, cum_bit_errors, n):
    """One CSV row exactly as dataForPlot.cpp:80-87 prints it (default ostream precision = %g)."""
    frames = float(tot[0])
    vals = [stnr, tot[1] / frames, cum_bit_errors / frames / n, tot[3] / frames, tot[4] / frames, tot[5] / frames]
    return ",".join("%g" % v for v in vals)


def sweep(engine, n, comm, p, e, seed=1, max_snr=5.0, out_path=None, chunk=1 << 16, log=None, cumulative_ber=True, min_snr=0.0):
    """cumulative_ber: the reference's BER* column never resets its bit-error count (dataForPlot.cpp:20,71,95); the polar
    sweep divides the point's own information-bit errors by frames * K (pass n = K)."""
    rows, raw = [], []
    cum_be = 0
    t0 = time.perf_counter()
    idx, stnr = int(round(2 * min_snr)), float(min_snr)
    fout = open(out_path + ".csv", "w") if (out_path and comm.rank == 0) else None
    while stnr <= max_snr:
        tot = run_point(engine, n, comm, stnr, idx, seed, p, e, chunk)
        cum_be = cum_be + int(tot[2]) if cumulative_ber else int(tot[2])   # BER* is cumulative over the sweep (dataForPlot.cpp:20,71,95)
        row = format_row(stnr, tot, cum_be, n)
        rows.append(row)
        raw.append(tot.copy())
        if comm.rank == 0:
            if fout:
                fout.write(row + "\n")
                fout.flush()
            print("%g" % stnr, file=log or sys.stdout, flush=True)
        stnr += 0.5
        idx += 1
    if fout:
        fout.close()
    if comm.rank == 0:
        print("Общее время: %g секунд" % (time.perf_counter() - t0), file=log or sys.stdout)
    return rows, np.array(raw)


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(description="kaneko <m> <t> <file> <p> <e> on 1..8 B200s (torchrun for > 1); with --polar SPEC: "
                                             "the FER sweep of a mixed-kernel polar code (m = list size, t ignored)")
    ap.add_argument("m", type=int)
    ap.add_argument("t", type=int)
    ap.add_argument("file")
    ap.add_argument("p", type=int)
    ap.add_argument("e", type=int, help="frame-error budget per point; <= 0 = run exactly p frames")
    ap.add_argument("--J", type=int, default=-1)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-snr", type=float, default=5.0)
    ap.add_argument("--min-snr", type=float, default=0.0)
    ap.add_argument("--polar", default=None, metavar="SPEC", help="code specification file (MixedKernelEncoder.cpp:10-93 format); m = list size")
    a = ap.parse_args(argv)
    import torch

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pkb200

    pk = pkb200.pk
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    comm = Comm(dev)
    if a.polar:
        if os.path.exists(a.polar):
            txt = open(a.polar).read().replace("@KERNEL@", os.path.join(pk.SPEC_DIR, "ebch16.kernel"))
        else:
            txt = pk.load_spec(a.polar)   # a specification committed under specs/
        pol = pk.Polar(txt, L=a.m, device=local)
        if comm.rank == 0:
            print(f"polar ({pol.N}, {pol.K}), {pol.layers} layers, L = {pol.L}")
        sweep(PolarEngine(pol), pol.K, comm, a.p, a.e, seed=a.seed, max_snr=a.max_snr, min_snr=a.min_snr, out_path=a.file, cumulative_ber=False, chunk=1 << 14)
        if comm.world > 1:
            torch.distributed.destroy_process_group()
        return
    code = pk.Code(a.m, a.t, device=local)
    kan = pk.Kaneko(code, J=a.J)
    if comm.rank == 0:
        print(" ".join(str(int(b)) for b in code.g) + " ")
        print(f"({code.n}, {code.k}, {code.d})")
    sweep(kan, code.n, comm, a.p, a.e, seed=a.seed, max_snr=a.max_snr, out_path=a.file)
    if comm.world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
