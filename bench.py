#!/usr/bin/env python
"""bench.py -- decoded frames/s of the Kaneko/BCH hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): BCH(31,16,7) with the HEAD (uncapped) test-pattern rule
and BCH(63,30,13) with the J=15 cap, every Eb/N0 point of the reference grid 0..5 dB step 0.5,
the same number of frames per point and code.  One "step" = one pass of the decoder over that
whole batch (2 codes x 11 points = 22 launch pairs: narrow phase A + wide phase B, 65 536 frames each).  Inputs are channel outputs y (f64)
drawn once by the device-side Philox generator.

  value  : frames/s with y resident in HBM, timed with CUDA events on the launching stream
  e2e    : frames/s through the C-ABI batch call a reference-side binding makes, with pinned HOST
           buffers: H2D of y and D2H of decisions + trial counts inside the timed region
           (pk_kaneko_decode_batch_async per batch + pk_kaneko_wait per step; the blocking
           pk_kaneko_decode_batch number is reported next to it as e2e.sync_call_value)
  --impl reference : the compiled reference (oracle/_ref) on all host cores, same codes/grid

Multi-GPU: frames are independent, so each rank decodes its own frames (weak scaling); the
only collective is one all-reduce of the per-point counters (what the sweep driver does per
SNR point), plus the max-over-ranks of the timings.
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
    ap.add_argument("--frames-per-point", type=int, default=1 << 16, help="frames per (code, SNR point) per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=12, help="frames per (code, point) for the single-core CPU baseline")
    ap.add_argument("--ref-frames", type=int, default=16, help="--impl reference: frames per (code, point) per process per step")
    ap.add_argument("--seed", type=int, default=1)
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
    widx, nframes, steps = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    out = []
    refs = []
    for (m, t, J, _) in CODES:
        r = oracle_py.Reference(m, t, J)
        r.seed(1000 + widx)
        refs.append(r)
    for _ in range(steps):
        t0 = time.perf_counter()
        trials = 0
        for r in refs:
            for s in SNRS:
                _, cw, y = r.gen_frames(s, nframes)
                _, tr, _, _ = r.kaneko_decode(y, answer=cw)
                trials += int(tr.sum())
        out.append((time.perf_counter() - t0, trials))
    return out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    if not (oracle_py.ref_available(False) and oracle_py.ref_available(True)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return 0
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nf = a.ref_frames
    total_steps = a.warmup + a.steps
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(w, nf, total_steps) for w in range(cores)])
        wall = time.perf_counter() - t0
    # per step: all processes run concurrently; step time = the slowest process
    step_t = [max(res[w][s][0] for w in range(cores)) for s in range(total_steps)][a.warmup:]
    trials = sum(res[w][s][1] for w in range(cores) for s in range(a.warmup, total_steps))
    frames_per_step = cores * nf * len(SNRS) * len(CODES)
    tsum = float(sum(step_t))
    value = frames_per_step * a.steps / tsum
    sample = f"{nf} frames x {len(SNRS)} SNR points x {len(CODES)} codes per process per step, {cores} processes (reference RNG, distinct seeds)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tsum / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 GF(2^m) + f64 metrics", "data": "synthetic (reference generator: minstd_rand0 AWGN)",
        "config": {"workload": "BASELINE configs[1]: BCH(31,16,7) uncapped + BCH(63,30,13) J=15 (patched cap), Eb/N0 0..5 dB step 0.5, equal frames per point and code",
                   "frames_per_point_per_process": nf, "processes": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "trials_per_s": trials / tsum, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- B200 arm
def run_b200(a):
    import torch
    import torch.distributed as dist

    import pkb200
    pk = pkb200.pk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG >= VERSION);
        # rank 0 must print ONE JSON line, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    B = a.frames_per_point
    # a real (non-default) stream: its handle goes through the C ABI, and every torch op and
    # CUDA event below is issued on the same stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    # ---- handles + device-resident inputs (drawn once by the Philox generator)
    codes, kans, ys, decs, trs, tots = [], [], [], [], [], []
    for (m, t, J, _) in CODES:
        c = pk.Code(m, t, device=local)
        k = pk.Kaneko(c, J=J)
        codes.append(c); kans.append(k)
        y = torch.empty((len(SNRS), B, c.n), dtype=torch.float64, device=dev)
        for si, s in enumerate(SNRS):
            k.generate_frames_dev(s, si, a.seed, rank * B, B, y[si].data_ptr(), stream=sp)
        ys.append(y)
        decs.append(torch.zeros((len(SNRS), B, c.n), dtype=torch.uint8, device=dev))
        trs.append(torch.zeros((len(SNRS), B), dtype=torch.int32, device=dev))
        tots.append(torch.zeros((len(SNRS), 8), dtype=torch.int64, device=dev))
    torch.cuda.synchronize()

    def device_step(events=None):
        for ci, k in enumerate(kans):
            for si in range(len(SNRS)):
                if events is not None:
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                k.decode_dev(ys[ci][si].data_ptr(), B, decs[ci][si].data_ptr(), trs[ci][si].data_ptr(), None,
                             tots[ci][si].data_ptr(), sp)
                if events is not None:
                    e1.record(stream)
                    events.append((ci, si, e0, e1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps (CUDA events on the launching stream)
    for _ in range(a.warmup):
        device_step()
    for tt in tots:
        tt.zero_()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pk.launch_count_reset()
    events = []
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.steps):
        device_step(events)
    ev1.record(stream)
    barrier()
    launches = pk.launch_count()
    dev_ms = ev0.elapsed_time(ev1)
    per_launch = {}
    for ci, si, e0, e1 in events:
        per_launch.setdefault((ci, si), []).append(e0.elapsed_time(e1))

    # ---- the one collective of the path: per-point counters summed over ranks
    tot_all = torch.stack(tots)[..., :6].contiguous()   # frames, frame errors, bit errors, trials, cmp, sum
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_all, op=dist.ReduceOp.SUM)
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms.item())
    frames_per_step = world * B * len(SNRS) * len(CODES)
    value = frames_per_step * a.steps / (dev_ms * 1e-3)
    tot_np = tot_all.cpu().numpy()
    trials_total = int(tot_np[..., 3].sum())

    # ---- end to end through the C ABI with pinned host buffers
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
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step(sync_calls)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    e2e_sync_s = time_e2e(True)
    for h in h_dec + h_tr:
        h.zero_()
    e2e_s = time_e2e(False)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = frames_per_step * a.steps / e2e_s
    h2d = sum(B * len(SNRS) * c.n * 8 for c in codes)
    d2h = sum(B * len(SNRS) * (c.n + 4) for c in codes)
    # same answers both ways
    for ci in range(len(CODES)):
        assert torch.equal(h_dec[ci], decs[ci].cpu()) and torch.equal(h_tr[ci], trs[ci].cpu()), "e2e != device-resident results"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (largest share of the step)
    share = {key: float(np.mean(v)) for key, v in per_launch.items()}
    step_ms = sum(share.values())
    by_code = [sum(v for (ci, _), v in share.items() if ci == c) for c in range(len(CODES))]
    dom = int(np.argmax(by_code))
    dom_code = codes[dom]
    dom_ms = by_code[dom] / len(SNRS)   # average launch duration of that kernel
    bytes_per_frame = 9 * dom_code.n + 4   # 8n in + n out + 4 (trial count)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_per_frame * B / (dom_ms * 1e-3) / 1e9
    dom_trials = int(tot_np[dom, :, 3].sum()) // max(1, world)
    mode = {0: "bit-sliced BM+Chien", 1: "coset table", 2: "cyclic-class table"}[dom_code.table_kind]
    kname = (f"k_phase_a + k_phase_b<{dom_code.m},{dom_code.t},{mode}> "
             "(one launch pair per SNR point; phase B is ~92 % of this code's kernel time, profiles/r1_launches.md)")
    sm_hz = (clocks["sm_mhz"] or 1965.0) * 1e6
    trials_per_s_dom = dom_trials / (by_code[dom] * a.steps * 1e-3)
    # the 0 dB launch (throughput regime, 25 000 trials per frame) against the unit that binds the class-table search:
    # the L1TEX -> XBAR request port, ONE L2 sector request per SM per clock (every trial gathers one random 32-byte
    # sector of the 32 MB class bitmap from L2; 32 lanes = 32 requests).  Per-trial counts are those of the committed
    # ncu capture of this kernel (profiles/r1_ncu_summary.md, prof_r1g_ct_tex: BCH(63,30,13) J=15 at 0 dB, 65 536 frames).
    NCU = {"l2_sectors_per_trial": 1.0124, "warp_inst_per_trial": 1.495, "xbar_req_cycles_active_pct": 96.2,
           "l1tex_throughput_pct": 96.2, "lts_throughput_pct": 77.4, "tex_data_pipe_pct": 57.2, "lsu_data_pipe_pct": 36.5,
           "dram_bytes_per_launch": 320465664}
    ms0 = share[(dom, 0)]
    trials0 = int(tot_np[dom, 0, 3]) // max(1, world) // a.steps
    tps0 = trials0 / (ms0 * 1e-3)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": NCU["dram_bytes_per_launch"] if (dom_code.n == 63 and B == 65536 and dom_code.table_kind == 2) else None,
        "algorithmic_bytes_per_launch": bytes_per_frame * B,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
        "kernel": kname, "avg_launch_ms": dom_ms, "share_of_step": by_code[dom] / step_ms,
        "note": ("integer search kernel: neither HBM- nor tensor-bound (SURVEY 8d).  `traffic` is the ncu DRAM traffic of the "
                 "0 dB launch (cold L2, ncu flushes it): 37 MB of frames, the first touch of the 32 MB class bitmap and the 32 MB "
                 "position table, and the ~1 % of bitmap gathers that miss L2 while the frames stream through it -- 53 GB/s, "
                 "0.8 % of the HBM roof, not a limiter.  The binding unit is the L1TEX -> XBAR request port (one L2 sector "
                 "request per SM per clock) -- see l2_request_port"),
        "l2_request_port": {
            "launch": "0 dB point of the dominant code", "launch_ms": ms0, "trials_per_s": tps0,
            "l2_sectors_per_trial_ncu": NCU["l2_sectors_per_trial"],
            "achieved_sector_requests_per_s": tps0 * NCU["l2_sectors_per_trial"],
            "peak_sector_requests_per_s": 148 * sm_hz,   # one request per SM per clock
            "frac": tps0 * NCU["l2_sectors_per_trial"] / (148 * sm_hz),
            "ncu_l1tex2xbar_req_cycles_active_pct": NCU["xbar_req_cycles_active_pct"],
            "ncu_l1tex_throughput_pct": NCU["l1tex_throughput_pct"], "ncu_lts_throughput_pct": NCU["lts_throughput_pct"],
            "ncu_tex_data_pipe_pct": NCU["tex_data_pipe_pct"], "ncu_lsu_data_pipe_pct": NCU["lsu_data_pipe_pct"],
            "warp_inst_per_trial_ncu": NCU["warp_inst_per_trial"],
            "whole_code_trials_per_s": trials_per_s_dom,
            "algorithmic_gf_macs_per_trial": 2 * dom_code.t * 2 + 2 * dom_code.t ** 2 + dom_code.n * dom_code.t,
            "source": "profiles/r1_ncu_summary.md (prof_r1g_ct_tex)",
        },
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
               "sample": f"first {nb} frames of every (code, SNR point) of this run's inputs ({n_cpu} frames, {t_cpu:.1f} s); decisions and trial counts checked equal to the GPU's",
               "trials_per_s": tr_cpu / t_cpu}

    # ---- side measurement (not part of `value`): polar SC / SC-list decode of BASELINE configs 3-4
    polar = None
    try:
        spec = pk.load_spec()
        polar = {"code": "(256,128), two layers of the 16x16 extended-BCH kernel, frozen set of this repository", "ebn0_db": 2.0}
        rng = np.random.default_rng(5)
        for Lp, Bp in ((1, 16384), (8, 4096), (32, 2048)):
            pp = pk.Polar(spec, L=Lp, device=local)
            info = rng.integers(0, 2, (Bp, pp.K), dtype=np.uint8)
            sg = np.sqrt(1 / (2 * (pp.K / pp.N) * 10 ** 0.2))
            cwp = pp.encode(info)
            llr = torch.from_numpy((2 * ((1 - 2.0 * cwp) + sg * rng.standard_normal(cwp.shape)) / sg ** 2).astype(np.float32)).to(dev)
            d_cnt = torch.zeros(Bp, dtype=torch.int32, device=dev)
            d_inf = torch.zeros((Bp, Lp, pp.K), dtype=torch.uint8, device=dev)
            best = None
            for _ in range(3):
                q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
                q0.record(stream)
                pp.decode_dev(llr.data_ptr(), Bp, d_cnt.data_ptr(), d_inf.data_ptr(), None, None, sp)
                q1.record(stream)
                torch.cuda.synchronize()
                t_ms = q0.elapsed_time(q1)
                best = t_ms if best is None else min(best, t_ms)
            polar[f"L{Lp}_frames_per_s"] = Bp / best * 1e3
            polar[f"L{Lp}_fer"] = float((d_inf[:, 0, :].cpu().numpy() != info).any(1).mean())
    except Exception as ex:   # the polar side measurement must never break the headline line
        polar = {"error": str(ex)}

    # ---- side measurement (not part of `value`): the large codes of BASELINE configs[4], J = 15
    large = None
    try:
        large = {}
        for (m_, t_, label) in ((7, 10, "BCH(127,64,21)"), (8, 15, "BCH(255,139,31)")):
            c_ = pk.Code(m_, t_, device=local)
            k_ = pk.Kaneko(c_, J=15, max_trials=1 << 16)   # bounds the pre-first-success search of hopeless frames
            Bl = 4096
            yl = torch.empty((Bl, c_.n), dtype=torch.float64, device=dev)
            dl = torch.zeros((Bl, c_.n), dtype=torch.uint8, device=dev)
            tl = torch.zeros(Bl, dtype=torch.int32, device=dev)
            totl = torch.zeros(8, dtype=torch.int64, device=dev)
            entry = {}
            for snr_ in (3.0, 4.0, 5.0):
                k_.generate_frames_dev(snr_, int(round(2 * snr_)), a.seed, 0, Bl, yl.data_ptr(), stream=sp)
                best = None
                for _ in range(2):
                    totl.zero_()
                    q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
                    q0.record(stream)
                    k_.decode_dev(yl.data_ptr(), Bl, dl.data_ptr(), tl.data_ptr(), None, totl.data_ptr(), sp)
                    q1.record(stream)
                    torch.cuda.synchronize()
                    t_ms = q0.elapsed_time(q1)
                    best = t_ms if best is None else min(best, t_ms)
                entry[f"{snr_:.0f}dB"] = {"frames_per_s": Bl / best * 1e3, "trials_per_s": int(totl[3].item()) / best * 1e3,
                                          "trials_per_frame": int(totl[3].item()) / Bl}
            large[label + " J=15, 4096 frames per launch, searches capped at 65536 trials"] = entry
    except Exception as ex:
        large = {"error": str(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 GF(2^m) + f64 metrics", "data": "synthetic (Philox4x32-10 info bits + AWGN drawn on the device)",
        "config": {"workload": "BASELINE configs[1]: BCH(31,16,7) uncapped + BCH(63,30,13) J=15, Eb/N0 0..5 dB step 0.5 (11 points), replay mode",
                   "frames_per_point_per_gpu": B, "frames_per_step": frames_per_step,
                   "l2_policy": f"inputs larger than L2: {h2d / 1e6:.0f} MB of y per step per GPU, every launch reads a distinct buffer"},
        "info_mbit_per_s": sum((B * len(SNRS) * world * c.k) for c in codes) * a.steps / (dev_ms * 1e-3) / 1e6,
        "trials_per_s": trials_total / (dev_ms * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": ("pk_kaneko_decode_batch_async per (code, SNR point) batch + pk_kaneko_wait per step "
                        "(pinned host y -> host decisions + trial counts; copies overlap the kernels of the neighbouring batches)"),
                "sync_call_value": frames_per_step * a.steps / e2e_sync_s,
                "sync_call_api": "pk_kaneko_decode_batch, one blocking call per (code, SNR point) batch"},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "per_code_ms_per_step": {CODES[c][3]: by_code[c] for c in range(len(CODES))},
        "polar_side_measurement": polar,
        "large_code_side_measurement": large,
    }
    if a.detail:
        for ci in range(len(CODES)):
            for si, s in enumerate(SNRS):
                ms = share[(ci, si)]
                trp = int(tot_np[ci, si, 3]) / a.steps / max(1, world) / B
                print(f"# {CODES[ci][3]:24s} {s:3.1f} dB  {ms:9.3f} ms/launch  {B / ms * 1e3:12.0f} frames/s  {trp:10.1f} trials/frame  {trp * B / ms * 1e3:14.0f} trials/s", file=sys.stderr)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse_args()
    sys.exit(run_reference(args) if args.impl == "reference" else run_b200(args))
