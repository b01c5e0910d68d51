"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle (liboracle.so, our
restatement) and for the compiled reference (oracle/_ref/libkaneko_ref*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref and os.path.isdir("/root/reference/src") else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


class _Base:
    """Common numpy-facing API over either library (same call shapes)."""

    _prefix = ""

    def _fn(self, name):
        return getattr(self.lib, self._prefix + name)

    def info(self):
        n, k, t, gs = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        g = np.zeros(1 << 16, np.uint8)
        self._fn("info")(self.h, C.byref(n), C.byref(k), C.byref(t), C.byref(gs), g)
        return n.value, k.value, t.value, g[: gs.value].copy()

    def tables(self):
        alog = np.zeros(self.n, np.uint64)
        log = np.zeros(self.n + 1, np.uint64)
        self._fn("tables")(self.h, alog, log)
        return alog, log

    def seed(self, s: int):
        self._seed(s)

    def gen_frames(self, ebn0_db: float, B: int):
        info = np.zeros((B, self.k), np.uint8)
        cw = np.zeros((B, self.n), np.uint8)
        y = np.zeros((B, self.n), np.float64)
        self._fn("gen_frames")(self.h, C.c_double(ebn0_db), C.c_long(B), info, cw, y)
        return info, cw, y

    def encode(self, info):
        info = np.ascontiguousarray(info, np.uint8)
        B = info.shape[0]
        cw = np.zeros((B, self.n), np.uint8)
        self._fn("encode")(self.h, info, C.c_long(B), cw)
        return cw

    def bdd(self, words):
        words = np.ascontiguousarray(words, np.uint8)
        B = words.shape[0]
        ans = np.zeros((B, self.n), np.uint8)
        ok = np.zeros(B, np.uint8)
        synd = np.zeros((B, 2 * self.t), np.uint64)
        lam = np.zeros((B, self.t + 1), np.uint64)
        lsz = np.zeros(B, np.int32)
        self._fn("bdd")(self.h, words, C.c_long(B), ans, ok, synd, lam, lsz)
        return ans, ok, synd, lam, lsz

    def make_matrix(self):
        out = np.zeros((self.n, self.n), np.uint8)
        self._fn("make_matrix")(self.h, out)
        return out


class Oracle(_Base):
    """Our CPU restatement (oracle/kaneko_oracle.c)."""

    _prefix = "ko_"

    def __init__(self, m: int, t: int, J: int = -1):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = lib = C.CDLL(path)
        lib.ko_create.restype = C.c_void_p
        lib.ko_create.argtypes = [C.c_int, C.c_int]
        lib.ko_destroy.argtypes = [C.c_void_p]
        lib.ko_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [_u8p]
        lib.ko_tables.argtypes = [C.c_void_p, _u64p, _u64p]
        lib.ko_set_J.argtypes = [C.c_void_p, C.c_long]
        lib.ko_seed.argtypes = [C.c_void_p, C.c_uint64]
        lib.ko_gen_frames.argtypes = [C.c_void_p, C.c_double, C.c_long, _u8p, _u8p, _f64p]
        lib.ko_encode.argtypes = [C.c_void_p, _u8p, C.c_long, _u8p]
        for nm in ("ko_kaneko_decode", "ko_kaneko_decode2"):
            getattr(lib, nm).argtypes = [C.c_void_p, _f64p, C.c_long, _u8p, _u32p, _u64p, _u64p]
        lib.ko_bdd.argtypes = [C.c_void_p, _u8p, C.c_long, _u8p, _u8p, _u64p, _u64p, _i32p]
        lib.ko_fun.restype = C.c_int
        lib.ko_fun.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_double, _f64p, _u64p]
        lib.ko_std_sort_pairs.argtypes = [_f64p, _i32p, C.c_int]
        lib.ko_make_matrix.argtypes = [C.c_void_p, _u8p]
        lib.ko_ext_gen_frames.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_long, _u8p, _u8p, _f64p]
        lib.ko_ext_kaneko_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, _f64p, C.c_long, _u8p, _u32p, _u64p, _u64p, _f64p]
        self.h = lib.ko_create(m, t)
        if not self.h:
            raise ValueError("Invalid values of arguments")
        self.m = m
        self.n, self.k, self.t, self.g = self.info()
        self.J = J
        lib.ko_set_J(self.h, J)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ko_destroy(self.h)
            self.h = None

    def _seed(self, s):
        self.lib.ko_seed(self.h, s)

    def set_J(self, J):
        self.J = J
        self.lib.ko_set_J(self.h, J)

    def kaneko_decode(self, y, decided=None, two_arg=False):
        y = np.ascontiguousarray(y, np.float64)
        B = y.shape[0]
        if decided is None:
            decided = np.zeros((B, self.n), np.uint8)
        trials = np.zeros(B, np.uint32)
        cmp_ = np.zeros(B, np.uint64)
        sum_ = np.zeros(B, np.uint64)
        fn = self.lib.ko_kaneko_decode2 if two_arg else self.lib.ko_kaneko_decode
        fn(self.h, y, C.c_long(B), decided, trials, cmp_, sum_)
        return decided, trials, cmp_, sum_

    # ---- extended codes / exact rules: OUR definition (the reference has neither), see kaneko_oracle.c
    def ext_gen_frames(self, ebn0_db, B, ext=1):
        ne = self.n + ext
        info = np.zeros((B, self.k), np.uint8)
        cw = np.zeros((B, ne), np.uint8)
        y = np.zeros((B, ne), np.float64)
        self.lib.ko_ext_gen_frames(self.h, ext, C.c_double(ebn0_db), C.c_long(B), info, cw, y)
        return info, cw, y

    def ext_kaneko_decode(self, y, ext=1, rules=0, llr_snr_db=0.5):
        y = np.ascontiguousarray(y, np.float64)
        B, ne = y.shape
        assert ne == self.n + ext
        decided = np.zeros((B, ne), np.uint8)
        trials = np.zeros(B, np.uint32)
        cmp_ = np.zeros(B, np.uint64)
        sum_ = np.zeros(B, np.uint64)
        lbest = np.zeros(B, np.float64)
        self.lib.ko_ext_kaneko_decode(self.h, ext, rules, C.c_double(llr_snr_db), y, C.c_long(B), decided, trials, cmp_, sum_, lbest)
        return decided, trials, cmp_, sum_, lbest

    def fun(self, p, e, max_snr=5.0):
        rows = np.zeros((64, 6), np.float64)
        raw = np.zeros((64, 6), np.uint64)
        npts = self.lib.ko_fun(self.h, p, e, C.c_double(max_snr), rows, raw)
        return rows[:npts].copy(), raw[:npts].copy()

    def std_sort(self, keys):
        k = np.ascontiguousarray(keys, np.float64).copy()
        idx = np.arange(len(k), dtype=np.int32)
        self.lib.ko_std_sort_pairs(k, idx, len(k))
        return k, idx


def ref_available(capped: bool = False) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libkaneko_ref_cap.so" if capped else "libkaneko_ref.so"))


class Reference(_Base):
    """The compiled reference (oracle/_ref/libkaneko_ref[_cap].so via ref_harness.cpp).

    NOTE: the reference keeps its RNG in a process-global, and each .so carries its
    own copy, so seeding is per library."""

    _prefix = "ref_"

    def __init__(self, m: int, t: int, J: int = -1):
        capped = J >= 0
        path = os.path.join(REF_DIR, "libkaneko_ref_cap.so" if capped else "libkaneko_ref.so")
        self.lib = lib = C.CDLL(path)
        lib.ref_create.restype = C.c_void_p
        lib.ref_create.argtypes = [C.c_int, C.c_int]
        lib.ref_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [_u8p]
        lib.ref_tables.argtypes = [C.c_void_p, _u64p, _u64p]
        lib.ref_seed.argtypes = [C.c_ulong]
        lib.ref_set_J.argtypes = [C.c_long]
        lib.ref_set_J.restype = C.c_int
        lib.ref_gen_frames.argtypes = [C.c_void_p, C.c_double, C.c_long, _u8p, _u8p, _f64p]
        lib.ref_encode.argtypes = [C.c_void_p, _u8p, C.c_long, _u8p]
        lib.ref_kaneko_decode.argtypes = [C.c_void_p, _f64p, _u8p, C.c_long, _u8p, _u32p, _u64p, _u64p]
        lib.ref_kaneko_decode2.argtypes = [C.c_void_p, _f64p, C.c_long, _u8p, _u32p, _u64p, _u64p]
        lib.ref_bdd.argtypes = [C.c_void_p, _u8p, C.c_long, _u8p, _u8p, _u64p, _u64p, _i32p]
        lib.ref_fun.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_long, C.c_double]
        lib.ref_make_matrix.argtypes = [C.c_void_p, _u8p]
        self.h = lib.ref_create(m, t)
        self.m = m
        self.n, self.k, self.t, self.g = self.info()
        self.J = J
        if capped:
            assert lib.ref_set_J(J) == 0

    def _seed(self, s):
        self.lib.ref_seed(s)

    def kaneko_decode(self, y, decided=None, two_arg=False, answer=None):
        y = np.ascontiguousarray(y, np.float64)
        B = y.shape[0]
        if decided is None:
            decided = np.zeros((B, self.n), np.uint8)
        trials = np.zeros(B, np.uint32)
        cmp_ = np.zeros(B, np.uint64)
        sum_ = np.zeros(B, np.uint64)
        if self.J >= 0:
            self.lib.ref_set_J(self.J)
        if two_arg:
            self.lib.ref_kaneko_decode2(self.h, y, C.c_long(B), decided, trials, cmp_, sum_)
        else:
            if answer is None:
                answer = np.zeros((B, self.n), np.uint8)
            self.lib.ref_kaneko_decode(self.h, y, np.ascontiguousarray(answer, np.uint8), C.c_long(B), decided, trials, cmp_, sum_)
        return decided, trials, cmp_, sum_

    def fun_csv(self, path_noext: str, p: int, e: int, max_snr: float = 5.0):
        """Runs the reference's own fun(); returns the parsed CSV rows."""
        if self.J >= 0:
            self.lib.ref_set_J(self.J)
        self.lib.ref_fun(self.h, path_noext.encode(), p, e, C.c_double(max_snr))
        return np.loadtxt(path_noext + ".csv", delimiter=",", ndmin=2)


class PolarReference:
    """The reference's vendored SC-list polar library (oracle/_ref/libpolar_ref.so via polar_ref_harness.cpp,
    built through oracle/polar_shim).  TEST INFRASTRUCTURE ONLY."""

    def __init__(self, spec_text: str, L: int = 1):
        path = os.path.join(REF_DIR, "libpolar_ref.so")
        self.lib = lib = C.CDLL(path, mode=os.RTLD_NOW)
        lib.pref_create.restype = C.c_void_p
        lib.pref_create.argtypes = [C.c_char_p, C.c_uint]
        lib.pref_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        lib.pref_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        lib.pref_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pref_trellis_llrs.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pref_trellis_llrs.restype = C.c_int
        self.h = lib.pref_create(spec_text.encode(), L)
        if not self.h:
            raise ValueError("reference rejected the specification")
        v = [C.c_int() for _ in range(4)]
        lib.pref_dims(self.h, *[C.byref(x) for x in v])
        self.N, self.K, self.N0, self.layers = (x.value for x in v)
        self.L = L

    def encode(self, info):
        info = np.ascontiguousarray(info, np.uint8)
        cw = np.zeros((info.shape[0], self.N), np.uint8)
        self.lib.pref_encode(self.h, info.ctypes.data, info.shape[0], cw.ctypes.data)
        return cw

    def decode(self, llr):
        llr = np.ascontiguousarray(llr, np.float32)
        B = llr.shape[0]
        cnt = np.zeros(B, np.int32)
        inf = np.zeros((B, self.L, self.K), np.uint8)
        cw = np.zeros((B, self.L, self.N), np.uint8)
        met = np.zeros((B, self.L), np.float32)
        self.lib.pref_decode(self.h, llr.ctypes.data, B, cnt.ctypes.data, inf.ctypes.data, cw.ctypes.data, met.ctypes.data)
        return cnt, inf, cw, met


def polar_ref_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libpolar_ref.so"))


def polar_ref_trellis_llrs(kernel_file: str, chan, u, stride: int = 1):
    """CTrellisKernelProcessor::GetLLRs over phases 0..l-1 for ONE kernel block group: chan/u [l*stride]."""
    lib = C.CDLL(os.path.join(REF_DIR, "libpolar_ref.so"), mode=os.RTLD_NOW)
    lib.pref_trellis_llrs.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pref_trellis_llrs.restype = C.c_int
    chan = np.ascontiguousarray(chan, np.float32)
    u = np.ascontiguousarray(u, np.uint8)
    out = np.zeros(chan.shape, np.float32)
    ab = np.zeros(64 * 65, np.uint32)
    l = lib.pref_trellis_llrs(kernel_file.encode(), stride, chan.ctypes.data, u.ctypes.data, out.ctypes.data, ab.ctypes.data)
    if l < 0:
        raise ValueError("reference trellis processor failed")
    return out, ab[: l * (l + 1)].reshape(l, l + 1)
