"""GPU, BASELINE-size runs checked through size-independent properties (the oracle cannot run these
sizes): range-split invariance (what multi-GPU sharding relies on), generation mode == replay mode on
the same frames (a checksum of checksums over two different kernels/IO paths), counter identities."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum")


@pytest.mark.parametrize("m,t,J,snr,frames", [
    (5, 3, -1, 4.0, 10_000_000),   # config 2: 10^7 frames per point, BCH(31,16,7) uncapped
    (6, 6, 15, 2.0, 1_000_000),    # config 2: BCH(63,30,13) J=15, 7.4e9 trials
    (4, 3, -1, 0.0, 10_000_000),   # config 1 code at its hardest point
])
def test_split_invariance_and_counter_identities(pk, m, t, J, snr, frames):
    code = pk.Code(m, t, device=0)
    kan = pk.Kaneko(code, J=J)
    si = int(round(snr * 2))
    whole, _ = kan.run_frames(snr, si, 5, 0, frames)
    parts = [kan.run_frames(snr, si, 5, off, n)[0] for off, n in ((0, frames // 8), (frames // 8, frames // 2), (frames // 8 + frames // 2, frames - frames // 8 - frames // 2))]
    for k in KEYS:
        assert sum(p[k] for p in parts) == whole[k], k
    assert max(p["max_trials_seen"] for p in parts) == whole["max_trials_seen"]
    assert whole["frames"] == frames and whole["trials"] >= frames
    n = code.n
    # comparisonCount / summCount identities (KanekoKernelProcessor.cpp:386-404): every completed trial adds
    # n+6 / n+1, improvements add small extras; an early return skips the add of its own trial
    assert whole["sum"] >= (whole["trials"] - frames) * (n + 1)
    assert whole["cmp"] - whole["sum"] >= (whole["trials"] - frames) * 5
    assert whole["frame_errors"] <= whole["bit_errors"] <= whole["frame_errors"] * n
    assert not (whole["flags_or"] & (pk.PK_FLAG_TRUNCATED | pk.PK_FLAG_NO_DECISION))


@pytest.mark.parametrize("m,t,J,snr,frames", [(5, 3, -1, 1.0, 300_000), (6, 6, 15, 1.5, 60_000), (4, 3, -1, 2.0, 2_000_000)])
def test_generation_equals_replay_on_same_frames(pk, m, t, J, snr, frames):
    """k_phase_*<GEN> (fused Philox front end, compare back end) vs k_phase_*<replay> through host buffers."""
    code = pk.Code(m, t, device=0)
    kan = pk.Kaneko(code, J=J)
    si = int(round(snr * 2))
    gen, grecs = kan.run_frames(snr, si, 99, 12345, frames, want_recs=True)
    info, cw, y = kan.generate_frames(snr, si, 99, 12345, frames)
    dec, trials, recs, tot = kan.decode(y)
    assert np.array_equal(trials, grecs["trials"])
    for k in ("frames", "trials", "cmp", "sum"):
        assert tot[k] == gen[k], k
    be = (dec != cw).sum(1)
    assert int(be.sum()) == gen["bit_errors"] and int((be > 0).sum()) == gen["frame_errors"]
    assert np.array_equal(be.astype(np.uint16), grecs["bit_errors"])
    # every decision is a codeword: re-encoding its information part (systematic check via syndromes = 0)
    ans, ok = code.bch_decode(dec[: min(frames, 50_000)])
    assert not ok.any()   # zero syndrome => the algebraic decoder reports failure (reference quirk), i.e. all are codewords
    # linearity of the encoder on the drawn words
    a, b = info[: 1000], info[1000:2000]
    assert np.array_equal(code.encode(a ^ b), code.encode(a) ^ code.encode(b))


def test_empty_and_single_frame_batches(pk):
    code = pk.Code(4, 3, device=0)
    kan = pk.Kaneko(code)
    dec, tr, recs, tot = kan.decode(np.zeros((0, 15)))
    assert dec.shape == (0, 15) and tot["frames"] == 0
    y = np.full((1, 15), -1.0)
    y[0, 3] = 0.2   # one unreliable wrong position around the all-zero codeword
    dec, tr, recs, tot = kan.decode(y)
    assert not dec.any() and tot["frames"] == 1 and tr[0] >= 1
    assert code.encode(np.zeros((0, 5), np.uint8)).shape == (0, 15)


@pytest.mark.parametrize("m,t,J,snr,frames", [(6, 6, 15, 0.5, 65_536), (6, 6, 15, 3.0, 262_144), (6, 4, 12, 1.0, 131_072), (5, 5, 12, 1.0, 131_072),
                                               (5, 3, -1, 1.0, 262_144)])
def test_table_search_equals_algebraic_search_at_bench_size(pk, m, t, J, snr, frames):
    """Two independent device implementations of the algebraic decoder inside the same search -- lookup tables
    (cyclic-class table / coset table) vs Berlekamp-Massey + Chien (table-driven in the narrow phase, bit-sliced in the
    wide phase) -- give identical decisions, trial counts and counters on a bench-size batch (1.6e9 trials for the first case)."""
    import torch

    out = []
    for tables in (True, False):
        code = pk.Code(m, t, device=0)
        assert code.uses_lut
        code.set_lut(tables)
        kan = pk.Kaneko(code, J=J)
        y = torch.empty((frames, code.n), dtype=torch.float64, device="cuda")
        kan.generate_frames_dev(snr, int(round(2 * snr)), 9, 0, frames, y.data_ptr())
        dec = torch.zeros((frames, code.n), dtype=torch.uint8, device="cuda")
        tr = torch.zeros(frames, dtype=torch.int32, device="cuda")
        tot = torch.zeros(8, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()   # the handle launches on its own stream: torch's fills must have landed
        kan.decode_dev(y.data_ptr(), frames, dec.data_ptr(), tr.data_ptr(), None, tot.data_ptr())
        torch.cuda.synchronize()
        out.append((dec.cpu(), tr.cpu(), tot.cpu()))
        del kan, code
    assert torch.equal(out[0][1], out[1][1]), "trial counts differ"
    assert torch.equal(out[0][0], out[1][0]), "decisions differ"
    assert torch.equal(out[0][2][:6], out[1][2][:6]), "totals differ"
    assert int(out[0][2][0]) == frames
