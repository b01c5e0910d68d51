#!/usr/bin/env python
"""bench.py -- decoded frames/s of the Kaneko/BCH Monte-Carlo hot path on B200 (BASELINE.json metric).

Default workload = BASELINE.json configs[1] AS WRITTEN: BCH(31,16,7) with the HEAD (uncapped) test-pattern rule and
BCH(63,30,13) with the J = 15 cap, every Eb/N0 point of the reference grid 0..5 dB step 0.5, **10^7 frames per SNR point**,
the frames of every point sharded over the N GPUs (strong scaling).  One "step" = the whole sweep: 2 codes x 11 points =
22 points, each = fun()'s per-frame loop (src/dataForPlot.cpp:43-74: random info -> c = u g -> BPSK + AWGN -> Kaneko
decode -> compare) for this rank's share of the 10^7 frames, fully on the device, followed by ONE NCCL all-reduce of the
point's counters through the C ABI (pk_allreduce_point) -- inside the timed region, on the stream the kernels run on.

  value   : frames/s of that sweep, device-timed (CUDA events on the launching stream, max over ranks)
  replay  : frames/s of the replay path (y resident in HBM -> decisions + trial counts), 65 536 frames per point per GPU
  e2e     : the replay path through the C-ABI batch call a reference-side binding makes, with pinned HOST buffers:
            H2D of y and D2H of decisions + trial counts inside the timed region (pk_kaneko_decode_batch_async per batch
            + pk_kaneko_wait per step)
  fun_e2e : the sweep through pk_comm_run_point, one blocking host call per point (what the drop-in fun() does)
  --impl reference : the compiled reference (oracle/_ref) on all host cores, same codes / grid / J

  --workload large : BASELINE configs[4] -- BCH(127,64,21) and BCH(255,139,31), J = 15 (see run_large)
  --workload polar : BASELINE configs[2-3] -- (256,128) polar code with two 16x16 eBCH kernels, SC / SCL-8 / SCL-32
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODES = [  # (m, t, J, label)
    (5, 3, -1, "BCH(31,16,7) uncapped"),
    (6, 6, 15, "BCH(63,30,13) J=15"),
]
SNRS = [0.5 * i for i in range(11)]
METRIC = "decoded frames/s (BCH(31,16,7) uncapped + BCH(63,30,13) J=15, Eb/N0 grid 0..5 dB)"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="configs1", choices=["configs1", "large", "polar"])
    ap.add_argument("--frames-per-point", type=int, default=10_000_000, help="frames per (code, SNR point), sharded over the GPUs")
    ap.add_argument("--replay-frames", type=int, default=1 << 16, help="frames per (code, SNR point) per GPU of the replay / e2e measurements")
    ap.add_argument("--cpu-sample", type=int, default=12, help="frames per (code, point) for the single-core CPU baseline")
    ap.add_argument("--ref-frames", type=int, default=16, help="--impl reference: frames per (code, point) per process per step")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-side", action="store_true", help="skip the polar / large-code side measurements of the default workload")
    ap.add_argument("--detail", action="store_true", help="print per-code / per-SNR numbers to stderr")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    """One process: the compiled reference decodes its own frames (reference RNG, own seed)."""
    widx, nframes, steps, codes, snrs = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    out = []
    refs = []
    for (m, t, J, _) in codes:
        r = oracle_py.Reference(m, t, J)
        r.seed(1000 + widx)
        refs.append(r)
    for _ in range(steps):
        t0 = time.perf_counter()
        trials = 0
        for r in refs:
            for s in snrs:
                _, cw, y = r.gen_frames(s, nframes)
                _, tr, _, _ = r.kaneko_decode(y, answer=cw)
                trials += int(tr.sum())
        out.append((time.perf_counter() - t0, trials))
    return out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if a.workload == "polar":
        return run_reference_polar(a)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    if not (oracle_py.ref_available(False) and oracle_py.ref_available(True)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return 0
    import multiprocessing as mp

    codes, snrs, metric, wl = CODES, SNRS, METRIC, "BASELINE configs[1]: BCH(31,16,7) uncapped + BCH(63,30,13) J=15 (patched cap), Eb/N0 0..5 dB step 0.5"
    if a.workload == "large":
        codes, snrs, metric, wl = LARGE_CODES, LARGE_SNRS, LARGE_METRIC, LARGE_WORKLOAD
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nf = a.ref_frames if a.workload == "configs1" else max(1, a.ref_frames // 8)
    total_steps = a.warmup + a.steps
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(w, nf, total_steps, codes, snrs) for w in range(cores)])
        wall = time.perf_counter() - t0
    # per step: all processes run concurrently; step time = the slowest process
    step_t = [max(res[w][s][0] for w in range(cores)) for s in range(total_steps)][a.warmup:]
    trials = sum(res[w][s][1] for w in range(cores) for s in range(a.warmup, total_steps))
    frames_per_step = cores * nf * len(snrs) * len(codes)
    tsum = float(sum(step_t))
    value = frames_per_step * a.steps / tsum
    sample = f"{nf} frames x {len(snrs)} SNR points x {len(codes)} codes per process per step, {cores} processes (reference RNG, distinct seeds)"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tsum / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8 GF(2^m) + f64 metrics", "data": "synthetic (reference generator: minstd_rand0 AWGN)",
        "config": {"workload": wl + ", equal frames per point and code (a bounded sample of the 10^7 frames per point)",
                   "frames_per_point_per_process": nf, "processes": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "trials_per_s": trials / tsum, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- B200 arm: shared plumbing
def bind_to_gpu_numa_node(local):
    """Run this rank on the cores of the NUMA node its GPU hangs off (so that page-locked host buffers, placed by first
    touch, and the H2D / D2H copies of the end-to-end legs stay on one socket).  Multi-GPU runs only; returns a note."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa_node unknown"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node}: no allowed cores"
        os.sched_setaffinity(0, cpus)
        return f"node {node}, {len(cpus)} cores"
    except Exception as exc:   # noqa: BLE001  (best effort: a box without NVML / sysfs topology runs unbound)
        return f"unbound ({type(exc).__name__})"


class Env:
    """One rank: torch for device memory and process-group plumbing, the C-ABI communicator for the data path."""

    def __init__(self, a):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.numa = bind_to_gpu_numa_node(self.local) if self.world > 1 else "single rank: not bound"
        import torch
        import torch.distributed as dist

        import pkb200

        self.torch, self.dist, self.pk = torch, dist, pkb200.pk
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        pk = self.pk
        if self.world > 1:
            # NCCL prints its version banner on stdout when a communicator is created (NCCL_DEBUG >= VERSION);
            # rank 0 must print ONE JSON line, so stdout points at stderr until the communicators exist
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                # the data path's own communicator (raw NCCL behind the C ABI): rank 0 draws the id, torch hands it round
                uid = torch.zeros(128, dtype=torch.uint8, device=self.dev)
                if self.rank == 0:
                    uid.copy_(torch.from_numpy(pk.Comm.unique_id()))
                dist.broadcast(uid, 0)
                self.comm = pk.Comm.from_rank(self.world, self.rank, uid.cpu().numpy(), self.local)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
        else:
            self.comm = pk.Comm.from_rank(1, 0, None, self.local)
        # every kernel, collective and CUDA event of the timed regions is issued on the communicator's stream
        self.sp = self.comm.stream(0)
        self.stream = torch.cuda.ExternalStream(self.sp, device=self.dev)
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def close(self):
        """End of the run: every rank has printed what it had to.  Pinned host tensors and CUDA events still reference the
        communicator's stream, and the caching allocators touch it again when they are freed -- at interpreter shutdown,
        in no particular order relative to the destruction of that stream -- so the process leaves through os._exit
        once the process group is down (stdout flushed first)."""
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


class Sweep:
    """fun()'s sweep for a list of codes: per (code, SNR point) this rank's share of the frames in generation mode
    (pk_kaneko_run_frames_dev) + one all-reduce of the counters (pk_allreduce_point), all on env.sp."""

    def __init__(self, env, codes, snrs, frames_per_point, seed, max_trials=0):
        torch, pk = env.torch, env.pk
        self.env, self.codes, self.snrs, self.P, self.seed = env, codes, snrs, frames_per_point, seed
        self.pkcodes, self.kans = [], []
        for (m, t, J, _) in codes:
            c = pk.Code(m, t, device=env.local)
            self.pkcodes.append(c)
            self.kans.append(pk.Kaneko(c, J=J, max_trials=max_trials))
        per = (frames_per_point + env.world - 1) // env.world
        self.first = env.rank * per
        self.count = max(0, min(per, frames_per_point - self.first))
        self.tot = torch.zeros((len(codes), len(snrs), 8), dtype=torch.int64, device=env.dev)

    def snr_index(self, s):
        return int(round(2 * s))

    def step(self, events=None):
        env = self.env
        self.tot.zero_()   # (torch fill on the same stream) the counters of a sweep start at zero, like fun()'s
        for ci, k in enumerate(self.kans):
            for si, s in enumerate(self.snrs):
                if events is not None:
                    e0, e1, e2 = env.event(), env.event(), env.event()
                    e0.record(env.stream)
                t = self.tot[ci, si]
                if self.count > 0:
                    k.run_frames_dev(s, self.snr_index(s), self.seed, self.first, self.count, t.data_ptr(), None, env.sp)
                if events is not None:
                    e1.record(env.stream)
                env.comm.allreduce_point([t.data_ptr()])
                if events is not None:
                    e2.record(env.stream)
                    events.append((ci, si, e0, e1, e2))

    def timed(self, steps, warmup):
        """returns (device ms over `steps` steps (max over ranks), per-point event list, reduced totals of the last
        step as numpy, kernels of this library launched inside the timed region)"""
        env = self.env
        for _ in range(warmup):
            self.step()
        env.barrier()
        env.pk.launch_count_reset()
        events = []
        ev0, ev1 = env.event(), env.event()
        ev0.record(env.stream)
        for _ in range(steps):
            self.step(events)
        ev1.record(env.stream)
        env.barrier()
        launches = env.pk.launch_count()
        ms = env.max_over_ranks(ev0.elapsed_time(ev1))
        return ms, events, self.tot.cpu().numpy().astype(np.int64), launches


def wilson(k, n, z=1.96):
    if n == 0:
        return (0.0, 1.0)
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return (max(0.0, c - h), min(1.0, c + h))


# ----------------------------------------------------------------------------- default workload: configs[1]
def run_b200(a):
    env = Env(a)
    torch, pk = env.torch, env.pk
    world, rank, dev, sp, stream = env.world, env.rank, env.dev, env.sp, env.stream
    P = a.frames_per_point

    sweep = Sweep(env, CODES, SNRS, P, a.seed)
    codes, kans = sweep.pkcodes, sweep.kans
    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()
    # ---- the timed region of `value`: K sweeps, 22 points each, all-reduce per point inside
    dev_ms, events, tot_np, launches = sweep.timed(a.steps, a.warmup)
    frames_per_step = P * len(SNRS) * len(CODES)
    value = frames_per_step * a.steps / (dev_ms * 1e-3)
    assert (tot_np[..., 0] == P).all(), "frame counters do not add up to the frames per point"
    trials_step = int(tot_np[..., 3].sum())
    per_point = {}
    for ci, si, e0, e1, e2 in events:
        d = per_point.setdefault((ci, si), [[], []])
        d[0].append(e0.elapsed_time(e1)); d[1].append(e1.elapsed_time(e2))

    # ---- fun_e2e: the same sweep through the blocking host call of the drop-in fun() (pk_comm_run_point per point)
    cks = [pk.CommKaneko(env.comm, m, t, J=J) for (m, t, J, _) in CODES]

    def fun_step():
        out = []
        for ck in cks:
            for s in SNRS:
                out.append(ck.run_point(s, sweep.snr_index(s), a.seed, P, 0))
        return out

    fun_step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fun_tot = fun_step()
    fun_s = env.max_over_ranks(time.perf_counter() - t0)
    for i, r in enumerate(fun_tot):
        ci, si = divmod(i, len(SNRS))
        assert [r[k] for k in ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum")] == [int(v) for v in tot_np[ci, si, :6]], "pk_comm_run_point != in-stream sweep"
    del cks

    # ---- replay mode: y resident in HBM -> decisions + trial counts (device-timed), B frames per point per GPU
    B = a.replay_frames
    ys, decs, trs, rtots = [], [], [], []
    for c, k in zip(codes, kans):
        y = torch.empty((len(SNRS), B, c.n), dtype=torch.float64, device=dev)
        for si, s in enumerate(SNRS):
            k.generate_frames_dev(s, sweep.snr_index(s), a.seed, rank * B, B, y[si].data_ptr(), stream=sp)
        ys.append(y)
        decs.append(torch.zeros((len(SNRS), B, c.n), dtype=torch.uint8, device=dev))
        trs.append(torch.zeros((len(SNRS), B), dtype=torch.int32, device=dev))
        rtots.append(torch.zeros((len(SNRS), 8), dtype=torch.int64, device=dev))
    torch.cuda.synchronize()

    def replay_step(evs=None):
        for ci, k in enumerate(kans):
            for si in range(len(SNRS)):
                if evs is not None:
                    e0, e1 = env.event(), env.event()
                    e0.record(stream)
                k.decode_dev(ys[ci][si].data_ptr(), B, decs[ci][si].data_ptr(), trs[ci][si].data_ptr(), None, rtots[ci][si].data_ptr(), sp)
                if evs is not None:
                    e1.record(stream)
                    evs.append((ci, si, e0, e1))

    for _ in range(a.warmup):
        replay_step()
    for t_ in rtots:
        t_.zero_()
    env.barrier()
    revents = []
    r0, r1 = env.event(), env.event()
    r0.record(stream)
    for _ in range(a.steps):
        replay_step(revents)
    r1.record(stream)
    env.barrier()
    replay_ms = env.max_over_ranks(r0.elapsed_time(r1))
    replay_frames_step = world * B * len(SNRS) * len(CODES)
    replay_value = replay_frames_step * a.steps / (replay_ms * 1e-3)
    rshare = {}
    for ci, si, e0, e1 in revents:
        rshare.setdefault((ci, si), []).append(e0.elapsed_time(e1))
    rshare = {k_: float(np.mean(v)) for k_, v in rshare.items()}
    rtot_np = torch.stack(rtots).cpu().numpy().astype(np.int64) // a.steps

    # ---- e2e: replay through the C ABI with pinned host buffers
    h_y = [y.cpu().pin_memory() for y in ys]
    h_dec = [torch.zeros(d.shape, dtype=torch.uint8).pin_memory() for d in decs]
    h_tr = [torch.zeros(t_.shape, dtype=torch.int32).pin_memory() for t_ in trs]

    def e2e_step(sync_calls=False):
        # one pk_kaneko handle per code; every (code, SNR point) batch goes host -> device -> host
        for si in range(len(SNRS)):
            for ci, k in enumerate(kans):
                if sync_calls:
                    k.decode_ptr(h_y[ci][si].data_ptr(), B, h_dec[ci][si].data_ptr(), h_tr[ci][si].data_ptr())
                else:
                    k.decode_async_ptr(h_y[ci][si].data_ptr(), B, h_dec[ci][si].data_ptr(), h_tr[ci][si].data_ptr())
        if not sync_calls:
            for k in kans:
                k.wait()   # results of the whole step are in host memory after this

    def time_e2e(sync_calls):
        e2e_step(sync_calls)  # allocates the pipeline workspaces
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step(sync_calls)
        torch.cuda.synchronize()
        return env.max_over_ranks(time.perf_counter() - t0)

    e2e_sync_s = time_e2e(True)
    for h in h_dec + h_tr:
        h.zero_()
    e2e_s = time_e2e(False)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = replay_frames_step * a.steps / e2e_s
    h2d = sum(B * len(SNRS) * c.n * 8 for c in codes)
    d2h = sum(B * len(SNRS) * (c.n + 4) for c in codes)
    for ci in range(len(CODES)):
        assert torch.equal(h_dec[ci], decs[ci].cpu()) and torch.equal(h_tr[ci], trs[ci].cpu()), "e2e != device-resident results"

    if rank != 0:
        env.close()
        return 0

    # ---- roofline of the dominant kernel
    gen_ms = {k_: float(np.mean(v[0])) for k_, v in per_point.items()}
    ar_ms = {k_: float(np.mean(v[1])) for k_, v in per_point.items()}
    by_code = [sum(v for (ci, _), v in gen_ms.items() if ci == c) for c in range(len(CODES))]
    dom = int(np.argmax(by_code))
    dom_code = codes[dom]
    pk_ = peaks()
    sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    mode = {0: "bit-sliced BM+Chien", 1: "coset table", 2: "cyclic-class table"}[dom_code.table_kind]
    # the 0 dB point (throughput regime, 25 000 trials per frame) against the unit that binds the class-table search: the
    # L1TEX -> XBAR request port, ONE L2 sector request per SM per clock (every probed trial gathers one random 32-byte
    # sector of the 32 MB class bitmap from L2).  Requests per trial: committed ncu capture (profiles/), stated there.
    NCU = json.load(open(os.path.join(ROOT, "profiles", "ncu_constants.json")))
    ms0 = gen_ms[(dom, 0)]
    trials0 = int(tot_np[dom, 0, 3]) // world
    tps0 = trials0 / (ms0 * 1e-3)
    spt = NCU["l2_sectors_per_trial"]
    launches_per_point = (sweep.count + (1 << 18) - 1) // (1 << 18)
    # contract view: algorithmic HBM bytes of the replay path (9n + 4 per frame) against the measured copy bandwidth
    rdom_ms = sum(v for (ci, _), v in rshare.items() if ci == dom) / len(SNRS)
    bytes_per_frame = 9 * dom_code.n + 4
    hbm_peak = float(pk_.get("hbm_gbs", 6650.0))
    hbm_ach = bytes_per_frame * B / (rdom_ms * 1e-3) / 1e9
    roofline = {
        "bound": "l2_request_port", "unit": "Gsector/s",
        "achieved": tps0 * spt / 1e9, "peak": 148 * sm_hz / 1e9, "frac": tps0 * spt / (148 * sm_hz),
        "traffic": NCU.get("dram_bytes_per_launch"), "traffic_note": NCU.get("dram_note"), "l2_hit_rate_pct_ncu": NCU.get("l2_hit_rate_pct"),
        "kernel": f"k_phase_b<{dom_code.m},{dom_code.t},{mode},generation> (wide search; {NCU.get('phase_b_share_pct', 92)} % of this code's kernel time)",
        "launch": f"0 dB point of {CODES[dom][3]}: {sweep.count} frames per GPU in {launches_per_point} launch pairs",
        "avg_launch_ms": ms0 / launches_per_point, "trials_per_s": tps0, "l2_sectors_per_trial_ncu": spt,
        "share_of_step": by_code[dom] / sum(by_code),
        "peak_source": "one L2 sector request per SM per clock: 148 SMs x the SM clock sampled during the run",
        "ncu": {k_: NCU[k_] for k_ in NCU if k_.startswith("ncu_")}, "ncu_source": NCU.get("source"),
        "note": ("integer search kernel: neither HBM- nor tensor-bound (SURVEY 8d); generation mode moves no per-frame data through HBM at all. "
                 "The unit that binds is the L1TEX -> XBAR request port (one L2 sector request per SM per clock)."),
        "hbm_view": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                     "algorithmic_bytes_per_launch": bytes_per_frame * B, "avg_launch_ms": rdom_ms,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if pk_ else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "what": f"replay launch of the same code ({B} frames, 9n+4 = {bytes_per_frame} B per frame)"},
    }

    # ---- single-core CPU baseline on a bounded sample of the same inputs (+ parity check)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    cpu = None
    nb = min(a.cpu_sample, B)
    if nb > 0:
        kind = "reference" if (oracle_py.ref_available(False) and oracle_py.ref_available(True)) else "port"
        t_cpu, n_cpu, tr_cpu = 0.0, 0, 0
        for ci, (m, t, J, _) in enumerate(CODES):
            eng = oracle_py.Reference(m, t, J) if kind == "reference" else oracle_py.Oracle(m, t, J)
            for si in range(len(SNRS)):
                ysample = h_y[ci][si, :nb].numpy()
                t0 = time.perf_counter()
                d_cpu, tr, _, _ = eng.kaneko_decode(ysample)
                t_cpu += time.perf_counter() - t0
                n_cpu += nb
                tr_cpu += int(tr.sum())
                assert np.array_equal(d_cpu, h_dec[ci][si, :nb].numpy()), "GPU decisions differ from the CPU reference"
                assert np.array_equal(tr.astype(np.int32), h_tr[ci][si, :nb].numpy()), "GPU trial counts differ"
        cpu = {"value": n_cpu / t_cpu, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"first {nb} frames of every (code, SNR point) of the replay inputs ({n_cpu} frames, {t_cpu:.1f} s, -O2 -fwrapv build); decisions and trial counts checked equal to the GPU's",
               "trials_per_s": tr_cpu / t_cpu}
        cpu["as_shipped"] = as_shipped_cpu()

    # ---- FER of the sweep against the published curves (out/*.csv values quoted in BASELINE.md)
    fer = {}
    for ci, (m, t, J, label) in enumerate(CODES):
        fer[label] = [{"ebn0_db": s, "fer": float(tot_np[ci, si, 1]) / P, "frame_errors": int(tot_np[ci, si, 1]),
                       "ci95": wilson(int(tot_np[ci, si, 1]), P), "trials_per_frame": float(tot_np[ci, si, 3]) / P}
                      for si, s in enumerate(SNRS)]

    side = {}
    if not a.no_side and world == 1:
        side["polar"] = guarded(lambda: polar_measure(env, a, short=True))
        side["large"] = guarded(lambda: large_measure(env, a, short=True))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8 GF(2^m) + f64 metrics", "data": "synthetic (Philox4x32-10 info bits + AWGN drawn on the device, counter = global frame index)",
        "config": {"workload": "BASELINE configs[1] as written: BCH(31,16,7) uncapped + BCH(63,30,13) J=15, Eb/N0 0..5 dB step 0.5 (11 points), "
                               f"{P} frames per SNR point sharded over the GPUs, generation mode (fun()'s loop on the device), one NCCL all-reduce of the counters per point inside the timed region",
                   "frames_per_point": P, "frames_per_step": frames_per_step, "frames_per_point_per_gpu": sweep.count,
                   "l2_policy": "generation mode reads no inputs; the replay / e2e measurements stream inputs larger than L2 "
                                f"({h2d / 1e6:.0f} MB of y per step per GPU, every launch a distinct buffer)"},
        "info_mbit_per_s": sum(P * len(SNRS) * c.k for c in codes) * a.steps / (dev_ms * 1e-3) / 1e6,
        "trials_per_s": trials_step * a.steps / (dev_ms * 1e-3),
        "allreduce_ms_per_step": sum(ar_ms.values()), "collectives_per_step": len(SNRS) * len(CODES),
        "replay": {"value": replay_value, "unit": UNIT, "ms_per_step": replay_ms / a.steps, "frames_per_point_per_gpu": B, "scaling": "weak",
                   "what": "replay mode, y resident in HBM -> decisions + trial counts, device-timed (round 1's `value`)",
                   "trials_per_s_per_gpu": int(rtot_np[..., 3].sum()) * a.steps / (replay_ms * 1e-3)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "frames_per_point_per_gpu": B,
                "api": ("pk_kaneko_decode_batch_async per (code, SNR point) batch + pk_kaneko_wait per step "
                        "(pinned host y -> host decisions + trial counts; copies overlap the kernels of the neighbouring batches)"),
                "sync_call_value": replay_frames_step * a.steps / e2e_sync_s,
                "sync_call_api": "pk_kaneko_decode_batch, one blocking call per (code, SNR point) batch",
                "host_binding_rank0": env.numa},
        "fun_e2e": {"value": frames_per_step * a.steps / fun_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64 * len(SNRS) * len(CODES),
                    "api": "pk_comm_run_point: one blocking host call per (code, SNR point), reduced counters back in host memory (what the drop-in fun() does)"},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "per_code_ms_per_step": {CODES[c][3]: by_code[c] for c in range(len(CODES))},
        "fer": fer, "side": side,
    }
    if a.detail:
        for ci in range(len(CODES)):
            for si, s in enumerate(SNRS):
                ms = gen_ms[(ci, si)]
                trp = int(tot_np[ci, si, 3]) / P
                print(f"# {CODES[ci][3]:24s} {s:3.1f} dB  {ms:9.3f} ms/point  {sweep.count / ms * 1e3:12.0f} frames/s/GPU  {trp:10.1f} trials/frame  "
                      f"{trp * sweep.count / ms * 1e3:14.0f} trials/s/GPU  allreduce {ar_ms[(ci, si)] * 1e3:7.1f} us  FER {tot_np[ci, si, 1] / P:.3e}   "
                      f"replay {rshare[(ci, si)]:8.3f} ms/launch", file=sys.stderr)
    print(json.dumps(line))
    env.close()
    return 0


def guarded(fn):
    try:
        return fn()
    except Exception as ex:   # a side measurement must never break the headline line
        return {"error": f"{type(ex).__name__}: {ex}"}


def as_shipped_cpu():
    """The reference binary as shipped (no -O flag, CMakeLists.txt:7-10) next to the -O2 -fwrapv build, single thread, on the
    one config[1] code HEAD can run (BCH(31,16,7) uncapped; J = 15 needs the patched cap line): `kaneko 5 3 <file> p e`."""
    import tempfile

    out = {}
    for tag, exe in (("O0_as_shipped", "kaneko_ref_O0"), ("O2_fwrapv", "kaneko_ref")):
        path = os.path.join(ROOT, "oracle", "_ref", exe)
        if not os.path.exists(path):
            out[tag] = None
            continue
        p = 40
        with tempfile.TemporaryDirectory() as td:
            t0 = time.perf_counter()
            try:
                subprocess.run([path, "5", "3", os.path.join(td, "x"), str(p), "1000000"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=120, check=True)
            except (subprocess.SubprocessError, OSError) as ex:
                out[tag] = f"failed: {ex}"
                continue
            dt = time.perf_counter() - t0
        out[tag] = {"frames_per_s": 11 * p / dt, "what": f"BCH(31,16,7) uncapped, whole grid, {p} frames per point, 1 thread, {dt:.1f} s"}
    return out


# ----------------------------------------------------------------------------- workload: large codes (configs[4])
LARGE_CODES = [(7, 10, 15, "BCH(127,64,21) J=15"), (8, 15, 15, "BCH(255,139,31) J=15")]
LARGE_SNRS = [3.0, 3.5, 4.0, 4.5, 5.0]
LARGE_METRIC = "decoded frames/s (BCH(127,64,21) + BCH(255,139,31), J=15, Eb/N0 3..5 dB)"
LARGE_WORKLOAD = "BASELINE configs[4]: BCH(127,64,21) and BCH(255,139,31) Kaneko decoding, J=15, Eb/N0 3..5 dB step 0.5"
LARGE_CAP = 1 << 24


def large_measure(env, a, short=False):
    """frames/s per (code, point) of the large codes in generation mode, frames sharded over the GPUs, all-reduce per point,
    searches bounded at 2^24 trials (a frame whose hard decision is far from every codeword keeps the initial bound
    2^31 - 1 until its first decodable pattern; the fraction of frames that hit the safety bound is reported), plus the
    SNR-independent trials/s figure of SURVEY 8d: pure-noise frames, every search stopped after exactly 2^15 patterns."""
    torch, pk = env.torch, env.pk
    P = (1 << 14) if short else (1 << 16)
    steps = 1 if short else max(1, a.steps)
    sweep = Sweep(env, LARGE_CODES, LARGE_SNRS, P, a.seed, max_trials=LARGE_CAP)
    ms, events, tot, _ = sweep.timed(steps, 1 if short else a.warmup)
    per = {}
    for ci, si, e0, e1, e2 in events:
        per.setdefault((ci, si), []).append(e0.elapsed_time(e1))
    out = {"frames_per_point": P, "search_bound": LARGE_CAP, "ms_per_step": ms / steps,
           "value": P * len(LARGE_SNRS) * len(LARGE_CODES) * steps / (ms * 1e-3), "codes": {}}
    for ci, (m, t, J, label) in enumerate(LARGE_CODES):
        pts = []
        for si, s in enumerate(LARGE_SNRS):
            pm = float(np.mean(per[(ci, si)]))
            pts.append({"ebn0_db": s, "frames_per_s": P / pm * 1e3 if env.world == 1 else None, "ms_per_point": pm,
                        "trials_per_frame": float(tot[ci, si, 3]) / P, "fer": float(tot[ci, si, 1]) / P,
                        "max_trials_seen": int(tot[ci, si, 6]), "any_truncated": bool(int(tot[ci, si, 7]) & pk.PK_FLAG_TRUNCATED)})
        out["codes"][label] = pts
    # fixed 2^15 patterns per frame: Eb/N0 = -20 dB frames never decode, max_trials = 2^15 ends every search there
    fixed = {}
    for (m, t, J, label) in LARGE_CODES:
        c = pk.Code(m, t, device=env.local)
        k = pk.Kaneko(c, J=J, max_trials=1 << 15)
        Bf = 16384 if short else 65536
        t_ = torch.zeros(8, dtype=torch.int64, device=env.dev)
        best = None
        for _ in range(3):
            t_.zero_()
            e0, e1 = env.event(), env.event()
            e0.record(env.stream)
            k.run_frames_dev(-20.0, 0, a.seed, env.rank * Bf, Bf, t_.data_ptr(), None, env.sp)
            e1.record(env.stream)
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1)
            best = dt if best is None else min(best, dt)
        tr = int(t_[3].item())
        fixed[label] = {"frames": Bf, "trials_per_frame": tr / Bf, "trials_per_s_per_gpu": tr / (best * 1e-3), "ms": best}
    out["fixed_2^15_patterns"] = fixed
    return out


def run_large(a):
    env = Env(a)
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    env.pk.launch_count_reset()
    res = large_measure(env, a)
    launches = env.pk.launch_count()
    clocks = sampler.stop() if env.rank == 0 else None
    if env.rank == 0:
        NCU = json.load(open(os.path.join(ROOT, "profiles", "ncu_constants.json")))
        sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        lab = LARGE_CODES[0][3]
        tps = res["fixed_2^15_patterns"][lab]["trials_per_s_per_gpu"]
        ipt = NCU.get("m7t10_warp_inst_per_trial")
        roofline = None
        if ipt:
            # only the instructions that go to the INT ALU pipe count against its roof (ncu: ALU pipe 74.2 % busy at 59.5 %
            # issue slots => 62 % of the warp instructions)
            alu = ipt * NCU.get("m7t10_alu_share_of_warp_inst", 1.0)
            roofline = {"bound": "alu_pipe", "unit": "Gwarp-inst/s", "achieved": tps * alu / 1e9, "peak": 148 * 4 * 0.5 * sm_hz / 1e9,
                        "frac": tps * alu / (148 * 4 * 0.5 * sm_hz), "traffic": NCU.get("m7t10_dram_bytes_per_launch"),
                        "kernel": "k_phase_b<7,10,bit-sliced BM+Chien> (fixed 2^15 patterns per frame)", "warp_inst_per_trial_ncu": ipt,
                        "alu_pipe_warp_inst_per_trial_ncu": alu, "ncu_alu_pipe_pct": NCU.get("m7t10_alu_pipe_pct"), "ncu_issue_active_pct": NCU.get("m7t10_issue_active_pct"),
                        "peak_source": "INT ALU pipe: 148 SMs x 4 SMSPs x 1/2 warp instruction per clock (LOP3 / IADD3 / SHF / PRMT)", "ncu_source": NCU.get("m7t10_source")}
        line = {"metric": LARGE_METRIC, "value": res["value"], "unit": UNIT, "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8 GF(2^m) + f64 metrics", "data": "synthetic (Philox4x32-10 info bits + AWGN drawn on the device)",
                "config": {"workload": LARGE_WORKLOAD + f", {res['frames_per_point']} frames per point sharded over the GPUs, searches bounded at 2^24 trials (any_truncated per point in detail)"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "detail": res}
        print(json.dumps(line))
    env.close()
    return 0


# ----------------------------------------------------------------------------- workload: polar (configs[2-3])
def polar_measure(env, a, short=False):
    from polar_bench import polar_measure as pm   # noqa: WPS433 (kept in its own file: bench_polar side)

    return pm(env, a, short)


def run_polar(a):
    from polar_bench import run_polar as rp

    return rp(a, Env, ClockSampler)


def run_reference_polar(a):
    from polar_bench import run_reference_polar as rr

    return rr(a)


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        sys.exit(run_reference(args))
    sys.exit({"configs1": run_b200, "large": run_large, "polar": run_polar}[args.workload](args))
