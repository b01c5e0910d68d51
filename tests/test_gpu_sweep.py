"""GPU: the three Monte-Carlo drivers agree -- sweep.py (Python, rank 0 of 1), pk_kaneko_run_point (C ABI) and the
C++ drop-in CLI kaneko_b200 (host/monte_carlo.cpp fun()) write the same CSV for the same seed, and a 2-GPU
torchrun sweep (when two devices are visible) reproduces the 1-GPU file byte for byte."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-codes-with-bch-kernel_b200")
sys.path.insert(0, PKG)


def test_python_sweep_equals_c_abi_run_point_and_cpp_cli(pk, tmp_path):
    import sweep

    os.environ.update(RANK="0", WORLD_SIZE="1")
    m, t, p, e = 4, 3, 30000, 200
    code = pk.Code(m, t, device=0)
    kan = pk.Kaneko(code)
    rows, raw = sweep.sweep(kan, code.n, sweep.Comm(), p, e, seed=1, out_path=str(tmp_path / "py"), log=open(os.devnull, "w"))
    assert len(rows) == 11
    cum = 0
    for si in range(11):
        r = kan.run_point(0.5 * si, si, 1, p, e)
        assert [r[k] for k in ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum")] == raw[si].tolist()
        assert r["frames"] == p or r["frame_errors"] == e
        cum += r["bit_errors"]
        assert rows[si] == sweep.format_row(0.5 * si, raw[si], cum, code.n)
    exe = os.path.join(PKG, "kaneko_b200")
    if not os.path.exists(exe):
        pytest.skip("kaneko_b200 not built")
    out = subprocess.run([exe, str(m), str(t), str(tmp_path / "cpp"), str(p), str(e)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0].split() == [str(int(b)) for b in code.g] and lines[1] == f"({code.n}, {code.k}, {code.d})"
    assert open(tmp_path / "cpp.csv").read() == open(tmp_path / "py.csv").read()


def test_cpp_cli_single_word_modes(pk, tmp_path):
    exe = os.path.join(PKG, "kaneko_b200")
    if not os.path.exists(exe):
        pytest.skip("kaneko_b200 not built")
    out = subprocess.run([exe, "4", "3", "4.0"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1] in ("Ok", "Errors were not corrected!")
    bad = subprocess.run([exe, "4", "9", "1.0"], capture_output=True, text=True, timeout=120)
    assert "Invalid values of arguments" in bad.stderr


def test_cpp_cli_file_mode(pk, tmp_path):
    """`kaneko 6 4 <snr> <file>` (main.cpp:125-166) on the reference's own in/infile.txt contents (kept as a golden
    fixture): the 2-argument decode flavour restores the transmitted word -> "Ok"."""
    exe = os.path.join(PKG, "kaneko_b200")
    if not os.path.exists(exe):
        pytest.skip("kaneko_b200 not built")
    z = np.load(os.path.join(ROOT, "tests", "golden", "infile_and_fun.npz"))
    f = tmp_path / "infile.txt"
    f.write_text(" ".join(str(int(b)) for b in z["infile_cw"][0]) + "\n" + " ".join(repr(float(v)) for v in z["infile_y"][0]) + "\n")
    out = subprocess.run([exe, "6", "4", "3.0", str(f)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "Ok"
    assert lines[-2].split() == [str(int(b)) for b in z["infile_cw"][0]]


def test_two_gpu_sweep_reproduces_one_gpu_csv(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(PKG, "sweep.py")
    one = subprocess.run([sys.executable, script, "5", "3", str(tmp_path / "g1"), "40000", "300"], capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stderr
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", script, "5", "3", str(tmp_path / "g2"), "40000", "300"], capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stderr
    assert open(tmp_path / "g1.csv").read() == open(tmp_path / "g2.csv").read()


KEYS6 = ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum")


def test_c_abi_communicator_of_one_rank_equals_run_point(pk):
    """pk_comm_* with a single rank (no NCCL needed): pk_comm_run_point == pk_kaneko_run_point for both stop rules, and
    pk_allreduce_point leaves a one-rank result alone."""
    import torch

    comm = pk.Comm(ndev=1)
    assert (comm.world, comm.rank, comm.local_devices) == (1, 0, 1)
    ck = pk.CommKaneko(comm, 4, 3)
    kan = pk.Kaneko(pk.Code(4, 3, device=0))
    for (snr, si, p, e) in [(0.0, 0, 50000, 100), (2.0, 4, 30000, 0), (5.0, 10, 20000, 1000)]:
        a, b = ck.run_point(snr, si, 9, p, e), kan.run_point(snr, si, 9, p, e)
        assert [a[k] for k in KEYS6] == [b[k] for k in KEYS6], (snr, p, e)
        assert a["max_trials_seen"] == b["max_trials_seen"]
    t = torch.arange(8, dtype=torch.int64, device="cuda") + 5
    torch.cuda.synchronize()
    comm.allreduce_point([t.data_ptr()])
    comm.sync()
    assert t.tolist() == list(range(5, 13))


def test_c_abi_two_gpus_equal_one_gpu(pk, tmp_path):
    """One process, two devices (ncclCommInitAll): the sharded point equals the one-device point for fixed-p and for the
    stop rule in global frame order, and `kaneko_b200 --gpus 2` writes the CSV of `--gpus 1` byte for byte."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    c1, c2 = pk.Comm(ndev=1), pk.Comm(ndev=2)
    assert c2.world == 2 and c2.local_devices == 2
    for (m, t, J, snr, si, p, e) in [(4, 3, -1, 0.0, 0, 100000, 300), (5, 3, -1, 1.0, 2, 200001, 0), (6, 6, 9, 3.0, 6, 50000, 0), (5, 3, -1, 3.0, 6, 300000, 150)]:
        k1, k2 = pk.CommKaneko(c1, m, t, J=J), pk.CommKaneko(c2, m, t, J=J)
        a, b = k1.run_point(snr, si, 3, p, e), k2.run_point(snr, si, 3, p, e)
        assert [a[k] for k in KEYS6] == [b[k] for k in KEYS6], (m, t, snr, p, e)
        assert a["max_trials_seen"] == b["max_trials_seen"] and a["flags_or"] == b["flags_or"]
    # device-resident results: sum / max / or over the two devices
    ts = []
    for d in range(2):
        with torch.cuda.device(d):
            ts.append(torch.tensor([1 + d, 2, 3, 4, 5, 6, 100 * (d + 1), 1 << d], dtype=torch.int64, device=f"cuda:{d}"))
            torch.cuda.synchronize()
    c2.allreduce_point([x.data_ptr() for x in ts])
    c2.sync()
    for x in ts:
        assert x.tolist() == [3, 4, 6, 8, 10, 12, 200, 3]
    exe = os.path.join(PKG, "kaneko_b200")
    if os.path.exists(exe):
        for g in (1, 2):
            out = subprocess.run([exe, "5", "3", str(tmp_path / f"c{g}"), "60000", "250", "--gpus", str(g)], capture_output=True, text=True, timeout=600)
            assert out.returncode == 0, out.stderr
        assert open(tmp_path / "c1.csv").read() == open(tmp_path / "c2.csv").read()


def test_non_ml_flag_and_errwords_dump(pk, tmp_path):
    """The reference's DEBUG build writes the frames whose transmitted word is more likely than the decision
    (calcL(res) < calcL(decoded), dataForPlot.cpp:55-64) to out/errWords.txt.  BCH(15,11,3) has plenty of them (the
    reference's rules are not ML there): the kernels' PK_FLAG_NON_ML equals the definition recomputed from
    the dumped frames, and `kaneko_b200 --errwords` writes exactly those frames in the reference's format."""
    code = pk.Code(4, 1, device=0)
    kan = pk.Kaneko(code)
    B, snr, si, seed = 20000, 1.0, 2, 1
    tot, recs = kan.run_frames(snr, si, seed, 0, B, want_recs=True)
    info, cw, y = kan.generate_frames(snr, si, seed, 0, B)
    dec, *_ = kan.decode(y)
    alpha = np.abs(y)
    yh = (y > 0).astype(np.uint8)
    l_tx = (alpha * (yh != cw)).sum(1)
    l_dec = (alpha * (yh != dec)).sum(1)
    want = (l_tx < l_dec) & (dec != cw).any(1)
    got = (recs["flags"] & pk.PK_FLAG_NON_ML) != 0
    # sums in another order may differ in the last ulp: compare away from exact ties
    clear = np.abs(l_tx - l_dec) > 1e-9 * (l_tx + l_dec + 1e-300)
    assert np.array_equal(got[clear], want[clear]) and got.sum() > 100
    assert bool(tot["flags_or"] & pk.PK_FLAG_NON_ML)
    exe = os.path.join(PKG, "kaneko_b200")
    if not os.path.exists(exe):
        pytest.skip("kaneko_b200 not built")
    ew = tmp_path / "errWords.txt"
    out = subprocess.run([exe, "4", "1", str(tmp_path / "dbg"), "3000", "1000000", "--errwords", str(ew)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = ew.read_text().splitlines()
    blocks = [lines[i:i + 4] for i in range(0, len(lines), 4)]
    assert len(blocks) > 50 and all(len(b) == 4 and b[3] == "" for b in blocks)
    n0 = sum(1 for b in blocks if b[0] == "0")
    _, r0 = kan.run_frames(0.0, 0, 1, 0, 3000, want_recs=True)
    assert n0 == int(((r0["flags"] & pk.PK_FLAG_NON_ML) != 0).sum())
    assert len(blocks[0][1].split()) == 15 and len(blocks[0][2].split()) == 15


def test_random_column_permutation_search(pk):
    """randomSwapColumns (root bchCoder.cpp:541-699) on the GPU: the winner's cost is what the host evaluates for the
    same trial, it does not exceed the natural order's, and every candidate is a column permutation of the kernel."""
    E = pk.ebch_kernel(4)
    few = pk.kernel_random_search(4, E, 500, seed=5, want_costs=True)
    host = np.array([pk.kernel_trellis_cost(pk.kernel_permute_columns(4, E, 5, t)[0])[0] for t in range(500)], np.uint64)
    assert np.array_equal(few["costs"], host), f"device and host disagree on candidates {np.nonzero(few['costs'] != host)[0][:8]}"
    assert (few["cost"], few["trial"]) == (int(host.min()), int(np.argmin(host)))
    r = pk.kernel_random_search(4, E, 300000, seed=5)
    assert r["input_cost"] == 11712 and r["cost"] <= r["input_cost"]
    again, basis = pk.kernel_permute_columns(4, E, 5, r["trial"])
    assert np.array_equal(again, r["matrix"]) and np.array_equal(basis, r["basis"])
    assert pk.kernel_trellis_cost(r["matrix"])[0] == r["cost"]
    assert sorted(map(tuple, r["matrix"].T.tolist())) == sorted(map(tuple, E.T.tolist()))
    r5 = pk.kernel_random_search(5, pk.ebch_kernel(5), 100000, seed=2, max_state_bits=14)
    assert r5["cost"] <= r5["input_cost"]


@pytest.mark.parametrize("L", [1, 8])
def test_polar_sweep_engine_and_stop_rule(pk, tmp_path, L):
    """sweep.py --polar: the polar simulator loop behind the same point driver -- fixed frame count = pk_polar_run_frames,
    the stop rule `count < p && countErr < e` = a sequential pass over the same frames cut right after the e-th error, and the
    CLI writes one CSV row per SNR point."""
    import sweep

    pol = pk.Polar(pk.load_spec(), L=L, device=0)
    eng = sweep.PolarEngine(pol)
    snr, si, seed = 1.5, 3, 9
    tot = sweep.run_point(eng, pol.K, sweep.Comm(), snr, si, seed, 5000, 0)
    ref = pol.run_frames(snr, si, seed, 0, 5000)
    assert (int(tot[0]), int(tot[1]), int(tot[2])) == (ref["frames"], ref["frame_errors"], ref["bit_errors"])
    # stop rule against a sequential pass
    e = 25
    got = sweep.run_point(eng, pol.K, sweep.Comm(), snr, si, seed, 20000, e, chunk=1 << 10)
    info, _, llr = pol.generate_frames(snr, si, seed, 0, int(got[0]) + 500)
    cnt, inf, _, _ = pol.decode(llr)
    be = (inf[:, 0, :] != info).sum(1)
    cut = int(np.nonzero(np.cumsum(be > 0) == e)[0][0]) + 1
    assert (int(got[0]), int(got[1]), int(got[2])) == (cut, e, int(be[:cut].sum()))
    # CLI
    out = tmp_path / f"polar_L{L}"
    sweep.main([str(L), "0", str(out), "3000", "0", "--polar", "polar_256_128_ebch16.spec.in", "--min-snr", "1.0", "--max-snr", "2.0", "--seed", "3"])
    rows = open(str(out) + ".csv").read().split()
    assert len(rows) == 3
    fer = [float(r.split(",")[1]) for r in rows]
    assert [r.split(",")[0] for r in rows] == ["1", "1.5", "2"] and fer[0] > fer[1] > fer[2] > 0
    p2 = pol.run_frames(2.0, 4, 3, 0, 3000)
    assert abs(fer[2] - p2["frame_errors"] / 3000) < 1e-9
