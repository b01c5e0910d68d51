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
