"""ctypes binding of libpkb200.so (include/pk_capi.h).

Thin plumbing for tests, bench.py and the multi-GPU sweep: it passes numpy host buffers or
torch device pointers straight through the C ABI.  There is no Python implementation of any
algorithm here and no fallback: if the shared library is missing the import fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpkb200.so")

PK_FLAG_EARLY_RETURN = 0x01
PK_FLAG_NO_DECISION = 0x02
PK_FLAG_SORT_TIE = 0x04
PK_FLAG_FRAME_ERROR = 0x08
PK_FLAG_TRUNCATED = 0x10
PK_FLAG_REF_UNDEFINED = 0x20
PK_FLAG_NON_ML = 0x40

#: numpy view of pk_frame_rec (16 bytes)
FRAME_REC = np.dtype(
    [("trials", "<u4"), ("extra_cmp", "<u4"), ("extra_sum", "<u4"), ("bit_errors", "<u2"), ("flags", "u1"), ("reserved", "u1")]
)
#: field order of pk_point_result (8 x u64)
POINT_FIELDS = ("frames", "frame_errors", "bit_errors", "trials", "cmp", "sum", "max_trials_seen", "flags_or")

#: every symbol include/pk_capi.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = (
    "pk_last_error pk_device_count pk_code_create pk_code_create_host pk_code_destroy pk_code_info pk_code_tables "
    "pk_code_uses_lut pk_code_set_lut pk_code_class_table_check pk_code_coset_table pk_encode_batch pk_bch_decode_batch pk_kaneko_create pk_kaneko_create_ext "
    "pk_kaneko_destroy pk_kaneko_set_variant pk_kaneko_set_frames_per_grab pk_kaneko_set_phase_a_limit pk_kaneko_launch_geometry pk_kaneko_decode_batch "
    "pk_kaneko_decode_batch_async pk_kaneko_wait pk_kaneko_decode_batch_dev pk_kaneko_run_frames_dev pk_kaneko_run_frames pk_generate_frames pk_generate_frames_dev "
    "pk_kaneko_run_point pk_make_kernel_matrix pk_launch_count pk_launch_count_reset "
    "pk_polar_create pk_polar_destroy pk_polar_info pk_polar_trellis_profile pk_polar_trellis_selfcheck pk_make_ebch_kernel pk_polar_encode_batch "
    "pk_polar_kernel_llrs pk_polar_decode_batch pk_polar_decode_batch_dev "
    "pk_comm_create pk_comm_unique_id pk_comm_create_rank pk_comm_destroy pk_comm_size pk_comm_rank pk_comm_local_devices pk_comm_stream "
    "pk_allreduce_point pk_comm_sync pk_comm_kaneko_create pk_comm_kaneko_destroy pk_comm_kaneko_local pk_comm_run_point "
    "pk_kproc_create pk_kproc_destroy pk_kproc_info pk_kproc_get_llrs pk_kproc_kernel_llrs "
    "pk_polar_run_frames_dev pk_polar_run_frames pk_polar_generate_frames_dev pk_polar_generate_frames "
    "pk_kernel_trellis_cost pk_kernel_swap_columns pk_kernel_permute_columns pk_kernel_random_search"
).split()


class PkError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libpkb200 status {status}: {msg}")
        self.status = status


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C {_HERE}` (or __graft_entry__.build()); "
            "there is no CPU fallback for the CUDA path"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i, l, d, u64 = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_uint64
    pp = C.POINTER(C.c_void_p)
    ip = C.POINTER(C.c_int)
    lib.pk_last_error.restype = C.c_char_p
    lib.pk_device_count.restype = i
    lib.pk_code_create.argtypes = [i, i, i, pp]
    lib.pk_code_create_host.argtypes = [i, i, pp]
    lib.pk_code_destroy.argtypes = [vp]
    lib.pk_code_destroy.restype = None
    lib.pk_code_info.argtypes = [vp, ip, ip, ip, ip, vp]
    lib.pk_code_tables.argtypes = [vp, vp, vp]
    lib.pk_code_uses_lut.argtypes = [vp]
    lib.pk_code_set_lut.argtypes = [vp, i]
    lib.pk_code_class_table_check.argtypes = [vp, u64, l, C.POINTER(l), C.POINTER(l)]
    lib.pk_code_coset_table.argtypes = [vp, vp, C.POINTER(l)]
    lib.pk_encode_batch.argtypes = [vp, vp, l, vp]
    lib.pk_bch_decode_batch.argtypes = [vp, vp, l, vp, vp]
    lib.pk_kaneko_create.argtypes = [vp, d, l, l, pp]
    lib.pk_kaneko_create_ext.argtypes = [vp, d, l, l, i, i, pp]
    lib.pk_kaneko_destroy.argtypes = [vp]
    lib.pk_kaneko_destroy.restype = None
    lib.pk_kaneko_set_frames_per_grab.argtypes = [vp, i]
    lib.pk_kaneko_set_phase_a_limit.argtypes = [vp, l]
    lib.pk_kaneko_set_variant.argtypes = [vp, i]
    lib.pk_kaneko_launch_geometry.argtypes = [vp, ip, ip, C.POINTER(l)]
    lib.pk_kaneko_decode_batch.argtypes = [vp, vp, l, vp, vp, vp, vp]
    lib.pk_kaneko_decode_batch_async.argtypes = [vp, vp, l, vp, vp, vp]
    lib.pk_kaneko_wait.argtypes = [vp, vp]
    lib.pk_kaneko_decode_batch_dev.argtypes = [vp, vp, l, vp, vp, vp, vp, vp]
    lib.pk_kaneko_run_frames_dev.argtypes = [vp, d, i, u64, u64, l, vp, vp, vp]
    lib.pk_kaneko_run_frames.argtypes = [vp, d, i, u64, u64, l, vp, vp]
    lib.pk_generate_frames.argtypes = [vp, d, i, u64, u64, l, vp, vp, vp]
    lib.pk_generate_frames_dev.argtypes = [vp, d, i, u64, u64, l, vp, vp, vp, vp]
    lib.pk_kaneko_run_point.argtypes = [vp, d, i, u64, l, l, vp]
    lib.pk_make_kernel_matrix.argtypes = [vp, vp]
    lib.pk_polar_create.argtypes = [C.c_char_p, i, i, pp]
    lib.pk_polar_destroy.argtypes = [vp]
    lib.pk_polar_destroy.restype = None
    lib.pk_polar_info.argtypes = [vp, ip, ip, ip, ip, ip]
    lib.pk_polar_trellis_profile.argtypes = [vp, i, ip, vp]
    lib.pk_polar_trellis_selfcheck.argtypes = [vp, i, u64, i, ip]
    lib.pk_make_ebch_kernel.argtypes = [i, vp]
    lib.pk_polar_encode_batch.argtypes = [vp, vp, l, vp]
    lib.pk_polar_kernel_llrs.argtypes = [vp, i, vp, vp, l, vp]
    lib.pk_polar_decode_batch.argtypes = [vp, vp, l, vp, vp, vp, vp]
    lib.pk_polar_decode_batch_dev.argtypes = [vp, vp, l, vp, vp, vp, vp, vp]
    lib.pk_comm_create.argtypes = [i, vp, pp]
    lib.pk_comm_unique_id.argtypes = [vp]
    lib.pk_comm_create_rank.argtypes = [i, i, vp, i, pp]
    lib.pk_comm_destroy.argtypes = [vp]
    lib.pk_comm_destroy.restype = None
    lib.pk_comm_size.argtypes = [vp]
    lib.pk_comm_rank.argtypes = [vp]
    lib.pk_comm_local_devices.argtypes = [vp]
    lib.pk_comm_stream.argtypes = [vp, i]
    lib.pk_comm_stream.restype = vp
    lib.pk_allreduce_point.argtypes = [vp, vp]
    lib.pk_comm_sync.argtypes = [vp]
    lib.pk_comm_kaneko_create.argtypes = [vp, i, i, d, l, l, pp]
    lib.pk_comm_kaneko_destroy.argtypes = [vp]
    lib.pk_comm_kaneko_destroy.restype = None
    lib.pk_comm_kaneko_local.argtypes = [vp, i]
    lib.pk_comm_kaneko_local.restype = vp
    lib.pk_comm_run_point.argtypes = [vp, d, i, u64, l, l, vp]
    lib.pk_polar_run_frames_dev.argtypes = [vp, d, i, u64, u64, l, vp, vp]
    lib.pk_polar_run_frames.argtypes = [vp, d, i, u64, u64, l, vp]
    lib.pk_polar_generate_frames_dev.argtypes = [vp, d, i, u64, u64, l, vp, vp, vp, vp]
    lib.pk_polar_generate_frames.argtypes = [vp, d, i, u64, u64, l, vp, vp, vp]
    lib.pk_kernel_trellis_cost.argtypes = [i, vp, vp, ip]
    lib.pk_kernel_swap_columns.argtypes = [i, vp, vp, vp]
    lib.pk_kernel_permute_columns.argtypes = [i, vp, u64, u64, vp, vp]
    lib.pk_kernel_random_search.argtypes = [i, vp, l, u64, i, i, vp, vp, vp, vp, vp, vp]
    lib.pk_kproc_create.argtypes = [i, i, l, i, pp]
    lib.pk_kproc_destroy.argtypes = [vp]
    lib.pk_kproc_destroy.restype = None
    lib.pk_kproc_info.argtypes = [vp, ip, vp, vp, vp]
    lib.pk_kproc_get_llrs.argtypes = [vp, i, i, vp, vp, vp, C.POINTER(l)]
    lib.pk_kproc_kernel_llrs.argtypes = [vp, vp, vp, l, vp, C.POINTER(l)]
    lib.pk_launch_count.restype = u64
    lib.pk_launch_count_reset.restype = None
    return lib


lib = _load()


def _check(rc):
    if rc != 0:
        raise PkError(rc, lib.pk_last_error().decode(errors="replace"))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def launch_count():
    return int(lib.pk_launch_count())


def launch_count_reset():
    lib.pk_launch_count_reset()


class Code:
    """pk_code handle: BCH code (m, t) with its tables; device=None builds host tables only."""

    def __init__(self, m, t, device=0):
        h = C.c_void_p()
        if device is None:
            _check(lib.pk_code_create_host(m, t, C.byref(h)))
        else:
            _check(lib.pk_code_create(m, t, int(device), C.byref(h)))
        self.h = h
        self.m, self.t, self.device = m, t, device
        n, k, d, gs = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        g = np.zeros(1 << m, np.uint8)
        _check(lib.pk_code_info(h, C.byref(n), C.byref(k), C.byref(d), C.byref(gs), _np_ptr(g)))
        self.n, self.k, self.d = n.value, k.value, d.value
        self.g = g[: gs.value].copy()

    def close(self):
        if getattr(self, "h", None) and lib is not None:  # lib is None during interpreter shutdown
            lib.pk_code_destroy(self.h)
            self.h = None

    __del__ = close

    def tables(self):
        alog = np.zeros(self.n, np.uint64)
        log = np.zeros(self.n + 1, np.uint64)
        _check(lib.pk_code_tables(self.h, _np_ptr(alog), _np_ptr(log)))
        return alog, log

    @property
    def uses_lut(self):
        return bool(lib.pk_code_uses_lut(self.h))

    @property
    def table_kind(self):
        """0: algebraic decoding only, 1: coset table, 2: cyclic-class table."""
        return int(lib.pk_code_uses_lut(self.h))

    def class_table_check(self, seed=1, ntrials=100000):
        """(mismatches, info dict) of the host-side class table against the algebraic decoder."""
        bad = C.c_long(-1)
        info = (C.c_long * 4)()
        _check(lib.pk_code_class_table_check(self.h, int(seed), int(ntrials), C.byref(bad), info))
        return bad.value, dict(key_bits=info[0], log2_slots=info[1], entries=info[2], bitmap_bytes=info[3])

    def set_lut(self, enable):
        _check(lib.pk_code_set_lut(self.h, int(bool(enable))))

    def coset_table(self):
        cnt = C.c_long()
        _check(lib.pk_code_coset_table(self.h, None, C.byref(cnt)))
        out = np.zeros(cnt.value, np.uint16)
        if cnt.value:
            _check(lib.pk_code_coset_table(self.h, _np_ptr(out), C.byref(cnt)))
        return out

    def kernel_matrix(self):
        out = np.zeros((self.n, self.n), np.uint8)
        _check(lib.pk_make_kernel_matrix(self.h, _np_ptr(out)))
        return out

    def encode(self, info):
        info = np.ascontiguousarray(info, np.uint8)
        assert info.ndim == 2 and info.shape[1] == self.k
        cw = np.zeros((info.shape[0], self.n), np.uint8)
        _check(lib.pk_encode_batch(self.h, _np_ptr(info), info.shape[0], _np_ptr(cw)))
        return cw

    def bch_decode(self, words, answers=None):
        words = np.ascontiguousarray(words, np.uint8)
        assert words.ndim == 2 and words.shape[1] == self.n
        B = words.shape[0]
        if answers is None:
            answers = np.zeros((B, self.n), np.uint8)
        ok = np.zeros(B, np.uint8)
        _check(lib.pk_bch_decode_batch(self.h, _np_ptr(words), B, _np_ptr(answers), _np_ptr(ok)))
        return answers, ok


class Kaneko:
    """pk_kaneko handle (KanekoKernelProcessor): J < 0 = HEAD semantics, J >= 0 = capped variant."""

    def __init__(self, code, J=-1, llr_snr_db=0.5, max_trials=0, extended=False, rules=0):
        self.code = code
        h = C.c_void_p()
        _check(lib.pk_kaneko_create_ext(code.h, float(llr_snr_db), int(J), int(max_trials), int(bool(extended)), int(rules), C.byref(h)))
        self.h = h
        self.J = J
        self.n = code.n + int(bool(extended))   # frame length (extended code: overall parity at position n)

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.pk_kaneko_destroy(self.h)
            self.h = None

    __del__ = close

    def set_frames_per_grab(self, g):
        _check(lib.pk_kaneko_set_frames_per_grab(self.h, int(g)))

    def set_variant(self, two_argument):
        """False: decode(answer, word, res) (default); True: the file-mode flavour decode(word, res)."""
        _check(lib.pk_kaneko_set_variant(self.h, int(bool(two_argument))))

    def set_phase_a_limit(self, trials):
        _check(lib.pk_kaneko_set_phase_a_limit(self.h, int(trials)))

    def geometry(self):
        g, b, s = C.c_int(), C.c_int(), C.c_long()
        _check(lib.pk_kaneko_launch_geometry(self.h, C.byref(g), C.byref(b), C.byref(s)))
        return g.value, b.value, s.value

    # ---- replay mode, host buffers (H2D / D2H inside the call)
    def decode(self, y, decided=None, want_recs=True):
        y = np.ascontiguousarray(y, np.float64)
        assert y.ndim == 2 and y.shape[1] == self.n
        B = y.shape[0]
        if decided is None:
            decided = np.zeros((B, self.n), np.uint8)
        trials = np.zeros(B, np.uint32)
        recs = np.zeros(B, FRAME_REC) if want_recs else None
        tot = np.zeros(8, np.uint64)
        _check(lib.pk_kaneko_decode_batch(self.h, _np_ptr(y), B, _np_ptr(decided), _np_ptr(trials), _np_ptr(recs), _np_ptr(tot)))
        return decided, trials, recs, dict(zip(POINT_FIELDS, (int(v) for v in tot)))

    def decode_ptr(self, y_ptr, B, decided_ptr, trials_ptr=None, recs_ptr=None, totals_ptr=None):
        """Host pointers given as ints (e.g. pinned torch tensors' data_ptr())."""
        _check(lib.pk_kaneko_decode_batch(self.h, y_ptr, B, decided_ptr, trials_ptr, recs_ptr, totals_ptr))

    def decode_async_ptr(self, y_ptr, B, decided_ptr, trials_ptr=None, recs_ptr=None):
        """Enqueue a batch (page-locked host pointers as ints) and return; wait() collects."""
        _check(lib.pk_kaneko_decode_batch_async(self.h, y_ptr, B, decided_ptr, trials_ptr, recs_ptr))

    def wait(self):
        """Block until every enqueued batch is finished; totals accumulated over them."""
        tot = np.zeros(8, np.uint64)
        _check(lib.pk_kaneko_wait(self.h, _np_ptr(tot)))
        return dict(zip(POINT_FIELDS, (int(v) for v in tot)))

    # ---- replay mode, device pointers (ints), asynchronous on `stream`
    def decode_dev(self, d_y, B, d_decided, d_trials=None, d_recs=None, d_totals=None, stream=None):
        _check(lib.pk_kaneko_decode_batch_dev(self.h, d_y, B, d_decided, d_trials, d_recs, d_totals, stream))

    # ---- generation mode
    def run_frames_dev(self, ebn0_db, snr_index, seed, first_frame, nframes, d_totals, d_recs=None, stream=None):
        _check(lib.pk_kaneko_run_frames_dev(self.h, float(ebn0_db), int(snr_index), int(seed), int(first_frame), int(nframes), d_recs, d_totals, stream))

    def run_frames(self, ebn0_db, snr_index, seed, first_frame, nframes, want_recs=False):
        recs = np.zeros(nframes, FRAME_REC) if want_recs else None
        tot = np.zeros(8, np.uint64)
        _check(lib.pk_kaneko_run_frames(self.h, float(ebn0_db), int(snr_index), int(seed), int(first_frame), int(nframes), _np_ptr(recs), _np_ptr(tot)))
        return dict(zip(POINT_FIELDS, (int(v) for v in tot))), recs

    def generate_frames(self, ebn0_db, snr_index, seed, first_frame, nframes):
        info = np.zeros((nframes, self.code.k), np.uint8)
        cw = np.zeros((nframes, self.n), np.uint8)
        y = np.zeros((nframes, self.n), np.float64)
        _check(lib.pk_generate_frames(self.h, float(ebn0_db), int(snr_index), int(seed), int(first_frame), int(nframes), _np_ptr(info), _np_ptr(cw), _np_ptr(y)))
        return info, cw, y

    def generate_frames_dev(self, ebn0_db, snr_index, seed, first_frame, nframes, d_y, d_cw=None, d_info=None, stream=None):
        _check(lib.pk_generate_frames_dev(self.h, float(ebn0_db), int(snr_index), int(seed), int(first_frame), int(nframes), d_info, d_cw, d_y, stream))

    def run_point(self, ebn0_db, snr_index, seed, p, e):
        tot = np.zeros(8, np.uint64)
        _check(lib.pk_kaneko_run_point(self.h, float(ebn0_db), int(snr_index), int(seed), int(p), int(e), _np_ptr(tot)))
        return dict(zip(POINT_FIELDS, (int(v) for v in tot)))


class Comm:
    """pk_comm: the GPUs of one box.  Comm(ndev=N) drives N devices from this process; Comm.from_rank(world, rank, id,
    device) is one rank of a job with one process per GPU (the id comes from Comm.unique_id() on rank 0)."""

    def __init__(self, ndev=1, devices=None, _h=None):
        if _h is not None:
            self.h = _h
        else:
            h = C.c_void_p()
            arr = None if devices is None else (C.c_int * ndev)(*devices)
            _check(lib.pk_comm_create(int(ndev), arr, C.byref(h)))
            self.h = h
        self.world = lib.pk_comm_size(self.h)
        self.rank = lib.pk_comm_rank(self.h)
        self.local_devices = lib.pk_comm_local_devices(self.h)

    @staticmethod
    def unique_id():
        buf = np.zeros(128, np.uint8)
        _check(lib.pk_comm_unique_id(_np_ptr(buf)))
        return buf

    @classmethod
    def from_rank(cls, world, rank, uid, device):
        h = C.c_void_p()
        uid = None if uid is None else np.ascontiguousarray(uid, np.uint8)
        _check(lib.pk_comm_create_rank(int(world), int(rank), _np_ptr(uid), int(device), C.byref(h)))
        return cls(_h=h)

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.pk_comm_destroy(self.h)
            self.h = None

    __del__ = close

    def stream(self, local=0):
        return lib.pk_comm_stream(self.h, local)

    def allreduce_point(self, d_result_ptrs):
        """d_result_ptrs: one device pointer (int) per local device, each to a pk_point_result."""
        arr = (C.c_void_p * len(d_result_ptrs))(*d_result_ptrs)
        _check(lib.pk_allreduce_point(self.h, arr))

    def sync(self):
        _check(lib.pk_comm_sync(self.h))


class CommKaneko:
    """pk_comm_kaneko: one Kaneko decoder per local device of a Comm; run_point shards an SNR point over all ranks."""

    def __init__(self, comm, m, t, J=-1, llr_snr_db=0.5, max_trials=0):
        self.comm = comm
        h = C.c_void_p()
        _check(lib.pk_comm_kaneko_create(comm.h, int(m), int(t), float(llr_snr_db), int(J), int(max_trials), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.pk_comm_kaneko_destroy(self.h)
            self.h = None

    __del__ = close

    def run_point(self, ebn0_db, snr_index, seed, p, e):
        tot = np.zeros(8, np.uint64)
        _check(lib.pk_comm_run_point(self.h, float(ebn0_db), int(snr_index), int(seed), int(p), int(e), _np_ptr(tot)))
        return dict(zip(POINT_FIELDS, (int(v) for v in tot)))


def kernel_trellis_cost(matrix):
    """(branch evaluations of one pass over all phases, max state bits) of the trellis kernel processor for `matrix`."""
    matrix = np.ascontiguousarray(matrix, np.uint8)
    cost = np.zeros(1, np.uint64)
    mb = C.c_int()
    _check(lib.pk_kernel_trellis_cost(matrix.shape[0], _np_ptr(matrix), _np_ptr(cost), C.byref(mb)))
    return int(cost[0]), mb.value


def kernel_swap_columns(power, field_elements, matrix):
    matrix = np.ascontiguousarray(matrix, np.uint8)
    fe = np.ascontiguousarray(field_elements, np.uint64)
    out = np.zeros_like(matrix)
    _check(lib.pk_kernel_swap_columns(int(power), _np_ptr(fe), _np_ptr(matrix), _np_ptr(out)))
    return out


def kernel_permute_columns(power, matrix, seed, trial):
    matrix = np.ascontiguousarray(matrix, np.uint8)
    out = np.zeros_like(matrix)
    basis = np.zeros(power, np.uint32)
    _check(lib.pk_kernel_permute_columns(int(power), _np_ptr(matrix), int(seed), int(trial), _np_ptr(out), _np_ptr(basis)))
    return out, basis


def kernel_random_search(power, matrix, ntrials, seed=1, device=0, max_state_bits=0, want_costs=False):
    """randomSwapColumns on the GPU -> dict(matrix, basis, cost, trial, input_cost[, costs of all candidates])"""
    matrix = np.ascontiguousarray(matrix, np.uint8)
    out = np.zeros_like(matrix)
    basis = np.zeros(power, np.uint32)
    v = np.zeros(3, np.uint64)
    costs = np.zeros(ntrials, np.uint64) if want_costs else None
    _check(lib.pk_kernel_random_search(int(power), _np_ptr(matrix), int(ntrials), int(seed), int(device), int(max_state_bits), _np_ptr(out), _np_ptr(basis),
                                       v[0:1].ctypes.data_as(C.c_void_p), v[1:2].ctypes.data_as(C.c_void_p), v[2:3].ctypes.data_as(C.c_void_p), _np_ptr(costs)))
    return dict(matrix=out, basis=basis, cost=int(v[0]), trial=int(v[1]), input_cost=int(v[2]), costs=costs)


class KanekoKernelProc:
    """pk_kproc: Kaneko decoding as the kernel processor (CKernProcLLR) of the (2^m) x (2^m) extended-BCH kernel."""

    def __init__(self, m, device=0, max_trials=0, enum_dim=-1):
        h = C.c_void_p()
        _check(lib.pk_kproc_create(int(m), int(device), int(max_trials), int(enum_dim), C.byref(h)))
        self.h = h
        sz = C.c_int()
        _check(lib.pk_kproc_info(h, C.byref(sz), None, None, None))
        self.size = sz.value
        self.mode = np.zeros(self.size, np.int32)
        self.t = np.zeros(self.size, np.int32)
        self.nextra = np.zeros(self.size, np.int32)
        _check(lib.pk_kproc_info(h, C.byref(sz), _np_ptr(self.mode), _np_ptr(self.t), _np_ptr(self.nextra)))

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.pk_kproc_destroy(self.h)
            self.h = None

    __del__ = close

    def get_llrs(self, phase, known, chan):
        """known, chan: [l][stride] -> (out [stride], truncated searches)"""
        known = np.ascontiguousarray(known, np.uint8)
        chan = np.ascontiguousarray(chan, np.float32)
        stride = chan.shape[1]
        out = np.zeros(stride, np.float32)
        tr = C.c_long()
        _check(lib.pk_kproc_get_llrs(self.h, stride, int(phase), _np_ptr(known), _np_ptr(chan), _np_ptr(out), C.byref(tr)))
        return out, tr.value

    def kernel_llrs(self, chan, u):
        """chan, u: [B][l] -> (out [B][l], truncated searches)"""
        chan = np.ascontiguousarray(chan, np.float32)
        u = np.ascontiguousarray(u, np.uint8)
        out = np.zeros(chan.shape, np.float32)
        tr = C.c_long()
        _check(lib.pk_kproc_kernel_llrs(self.h, _np_ptr(chan), _np_ptr(u), chan.shape[0], _np_ptr(out), C.byref(tr)))
        return out, tr.value


def counters_from_recs(recs, n):
    """(decodingCount, comparisonCount, summCount) per frame from pk_frame_rec records."""
    tr = recs["trials"].astype(np.uint64)
    run = tr - (recs["flags"] & PK_FLAG_EARLY_RETURN).astype(np.uint64)
    return tr, run * np.uint64(n + 6) + recs["extra_cmp"], run * np.uint64(n + 1) + recs["extra_sum"]


SPEC_DIR = os.path.join(_HERE, "specs")


def ebch_kernel(m):
    """(2^m) x (2^m) extended-BCH kernel matrix (root bchCoder.cpp:356-389)."""
    n = 1 << m
    out = np.zeros((n, n), np.uint8)
    _check(lib.pk_make_ebch_kernel(int(m), _np_ptr(out)))
    return out


def load_spec(name="polar_256_128_ebch16.spec.in"):
    """A committed specification template with @KERNEL@ replaced by the absolute path of the kernel file."""
    txt = open(os.path.join(SPEC_DIR, name)).read()
    return txt.replace("@KERNEL@", os.path.join(SPEC_DIR, "ebch16.kernel"))


class Polar:
    """pk_polar handle: mixed-kernel polar code + SC (L = 1) / SC-list decoder; device=None: host tables only."""

    def __init__(self, spec_text, L=1, device=0):
        h = C.c_void_p()
        _check(lib.pk_polar_create(spec_text.encode(), int(L), -1 if device is None else int(device), C.byref(h)))
        self.h = h
        v = [C.c_int() for _ in range(5)]
        _check(lib.pk_polar_info(h, *[C.byref(x) for x in v]))
        self.N, self.K, self.N0, self.layers, self.L = (x.value for x in v)

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.pk_polar_destroy(self.h)
            self.h = None

    __del__ = close

    def trellis_profile(self, layer=0):
        sz = C.c_int()
        _check(lib.pk_polar_trellis_profile(self.h, layer, C.byref(sz), None))
        out = np.zeros((sz.value, sz.value + 1), np.uint8)
        _check(lib.pk_polar_trellis_profile(self.h, layer, C.byref(sz), _np_ptr(out)))
        return out

    def trellis_selfcheck(self, layer=0, seed=1, ntests=50):
        """in-place vs gather-form trellis tables of the layer's kernel on the host -> state index bits of the in-place numbering"""
        bits = C.c_int()
        _check(lib.pk_polar_trellis_selfcheck(self.h, layer, int(seed), int(ntests), C.byref(bits)))
        return bits.value

    def encode(self, info):
        info = np.ascontiguousarray(info, np.uint8)
        cw = np.zeros((info.shape[0], self.N), np.uint8)
        _check(lib.pk_polar_encode_batch(self.h, _np_ptr(info), info.shape[0], _np_ptr(cw)))
        return cw

    def kernel_llrs(self, chan, u, layer=0):
        chan = np.ascontiguousarray(chan, np.float32)
        u = np.ascontiguousarray(u, np.uint8)
        out = np.zeros(chan.shape, np.float32)
        _check(lib.pk_polar_kernel_llrs(self.h, layer, _np_ptr(chan), _np_ptr(u), chan.shape[0], _np_ptr(out)))
        return out

    def decode(self, llr):
        llr = np.ascontiguousarray(llr, np.float32)
        B = llr.shape[0]
        cnt = np.zeros(B, np.int32)
        inf = np.zeros((B, self.L, self.K), np.uint8)
        cw = np.zeros((B, self.L, self.N), np.uint8)
        met = np.zeros((B, self.L), np.float32)
        _check(lib.pk_polar_decode_batch(self.h, _np_ptr(llr), B, _np_ptr(cnt), _np_ptr(inf), _np_ptr(cw), _np_ptr(met)))
        return cnt, inf, cw, met

    def decode_dev(self, d_llr, B, d_count, d_inf, d_cw=None, d_metric=None, stream=None):
        _check(lib.pk_polar_decode_batch_dev(self.h, d_llr, B, d_count, d_inf, d_cw, d_metric, stream))

    # ---- generation mode (the simulator loop on the device)
    def run_frames_dev(self, ebn0_db, snr_index, seed, first_frame, nframes, d_totals, stream=None):
        _check(lib.pk_polar_run_frames_dev(self.h, float(ebn0_db), int(snr_index), int(seed), int(first_frame), int(nframes), d_totals, stream))

    def run_frames(self, ebn0_db, snr_index, seed, first_frame, nframes):
        tot = np.zeros(8, np.uint64)
        _check(lib.pk_polar_run_frames(self.h, float(ebn0_db), int(snr_index), int(seed), int(first_frame), int(nframes), _np_ptr(tot)))
        return dict(zip(POINT_FIELDS, (int(v) for v in tot)))

    def generate_frames(self, ebn0_db, snr_index, seed, first_frame, nframes):
        info = np.zeros((nframes, self.K), np.uint8)
        cw = np.zeros((nframes, self.N), np.uint8)
        llr = np.zeros((nframes, self.N), np.float32)
        _check(lib.pk_polar_generate_frames(self.h, float(ebn0_db), int(snr_index), int(seed), int(first_frame), int(nframes), _np_ptr(info), _np_ptr(cw), _np_ptr(llr)))
        return info, cw, llr
