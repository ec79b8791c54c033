"""TEST INFRASTRUCTURE -- ctypes loaders for the checkers under oracle/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
The shipped engine (agatha_b200/) never does.

  Oracle      -- this repo's C restatement (oracle/agatha_oracle.c -> libagatha_oracle.so)
  RefHost     -- the reference's own kernel header compiled as single-lane host code
                 (oracle/_ref/libagatha_ref_host.so, built by oracle/Makefile where /root/reference exists)
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libagatha_oracle.so")
REF_HOST_SO = os.path.join(HERE, "_ref", "libagatha_ref_host.so")
REF_GPU_BIN = os.path.join(HERE, "_ref", "agatha_ref_manual")

STOP_END, STOP_ZDROP, STOP_BANDEXIT = 0, 1, 2


class Params(ctypes.Structure):
    """Mirror of gasal_subst_scores (AGAThA/src/gasal.h:165-173)."""
    _fields_ = [("match", ctypes.c_int32), ("mismatch", ctypes.c_int32), ("gap_open", ctypes.c_int32),
                ("gap_extend", ctypes.c_int32), ("slice_width", ctypes.c_int32),
                ("z_threshold", ctypes.c_int32), ("band_width", ctypes.c_int32)]


# AGAThA.sh:44 -- the scoring the reference ships with
DEFAULT_PARAMS = dict(match=1, mismatch=4, gap_open=6, gap_extend=2, slice_width=3, z_threshold=400, band_width=751)

RESULT_DTYPE = np.dtype([("score", "<i4"), ("query_end", "<i4"), ("target_end", "<i4"), ("stop", "<i4"),
                         ("d_stop", "<i4"), ("reserved", "<i4"), ("cells", "<i8")])


def make_params(**kw):
    d = dict(DEFAULT_PARAMS)
    d.update(kw)
    return Params(**d)


def build(ref=True):
    """Compile the checkers (idempotent). ref=True also (re)builds oracle/_ref when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _as_u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    if isinstance(x, str):
        return np.frombuffer(x.encode(), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def _ptr(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


class Oracle:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = ctypes.CDLL(path)
        self.lib.agatha_oracle_align_batch.restype = ctypes.c_int
        self.lib.agatha_oracle_band_cells.restype = ctypes.c_int64
        self.lib.agatha_oracle_band_cells.argtypes = [ctypes.c_int32] * 3

    def set_model(self, phantom=True):
        self.lib.agatha_oracle_set_model(ctypes.c_int32(1 if phantom else 0))

    def align_batch(self, qbuf, qoff, qlen, tbuf, toff, tlen, params, nthreads=0):
        """qbuf/tbuf: uint8 ASCII; offsets in bytes; returns a structured array (RESULT_DTYPE)."""
        qbuf, tbuf = _as_u8(qbuf), _as_u8(tbuf)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint32); toff = np.ascontiguousarray(toff, dtype=np.uint32)
        qlen = np.ascontiguousarray(qlen, dtype=np.uint32); tlen = np.ascontiguousarray(tlen, dtype=np.uint32)
        n = len(qlen)
        out = np.zeros(n, dtype=RESULT_DTYPE)
        rc = self.lib.agatha_oracle_align_batch(
            _ptr(qbuf, ctypes.c_uint8), _ptr(qoff, ctypes.c_uint32), _ptr(qlen, ctypes.c_uint32),
            _ptr(tbuf, ctypes.c_uint8), _ptr(toff, ctypes.c_uint32), _ptr(tlen, ctypes.c_uint32),
            ctypes.c_int32(n), ctypes.byref(params), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(nthreads))
        if rc < 0:
            raise MemoryError("oracle scratch allocation failed")
        self.threads_used = rc
        return out

    def align_pairs(self, pairs, params, nthreads=0):
        """pairs: list of (query, target) as bytes/str/uint8 arrays."""
        batch = concat_pairs(pairs)
        return self.align_batch(*batch, params, nthreads=nthreads)

    def band_cells(self, qlen, tlen, w):
        return int(self.lib.agatha_oracle_band_cells(int(qlen), int(tlen), int(w)))

    def apply_op(self, seq, op):
        """Reverse (bit 0) / complement (bit 1) of one ASCII sequence; returns a new uint8 array."""
        a = np.array(_as_u8(seq), dtype=np.uint8, copy=True)
        self.lib.agatha_oracle_apply_op(_ptr(a, ctypes.c_uint8), ctypes.c_int32(len(a)), ctypes.c_int32(int(op)))
        return a


class RefHost:
    """The reference kernel itself, as host code. Valid only inside the reference's int16 domain."""

    def __init__(self, path=REF_HOST_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        self.lib.ref_host_align_batch.restype = ctypes.c_int

    @staticmethod
    def available():
        return os.path.exists(REF_HOST_SO)

    def align_batch(self, qbuf, qoff, qlen, tbuf, toff, tlen, params, nthreads=0):
        qbuf, tbuf = _as_u8(qbuf), _as_u8(tbuf)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint32); toff = np.ascontiguousarray(toff, dtype=np.uint32)
        qlen = np.ascontiguousarray(qlen, dtype=np.uint32); tlen = np.ascontiguousarray(tlen, dtype=np.uint32)
        n = len(qlen)
        p = np.array([params.match, params.mismatch, params.gap_open, params.gap_extend, params.slice_width,
                      params.z_threshold, params.band_width], dtype=np.int32)
        out = np.zeros((n, 3), dtype=np.int32)
        self.lib.ref_host_align_batch(
            _ptr(qbuf, ctypes.c_uint8), _ptr(qoff, ctypes.c_uint32), _ptr(qlen, ctypes.c_uint32),
            _ptr(tbuf, ctypes.c_uint8), _ptr(toff, ctypes.c_uint32), _ptr(tlen, ctypes.c_uint32),
            ctypes.c_int32(n), _ptr(p, ctypes.c_int32), _ptr(out, ctypes.c_int32), ctypes.c_int32(nthreads))
        return out

    def align_pairs(self, pairs, params, nthreads=0):
        return self.align_batch(*concat_pairs(pairs), params, nthreads=nthreads)


def concat_pairs(pairs):
    """[(q, t), ...] -> (qbuf, qoff, qlen, tbuf, toff, tlen) with byte offsets (no padding)."""
    qs = [_as_u8(q) for q, _ in pairs]
    ts = [_as_u8(t) for _, t in pairs]
    qlen = np.array([len(x) for x in qs], dtype=np.uint32)
    tlen = np.array([len(x) for x in ts], dtype=np.uint32)
    qoff = np.zeros(len(qs), dtype=np.uint32); toff = np.zeros(len(ts), dtype=np.uint32)
    if len(qs):
        qoff[1:] = np.cumsum(qlen[:-1]); toff[1:] = np.cumsum(tlen[:-1])
    qbuf = np.concatenate(qs) if qs else np.zeros(0, np.uint8)
    tbuf = np.concatenate(ts) if ts else np.zeros(0, np.uint8)
    if len(qbuf) == 0:
        qbuf = np.zeros(1, np.uint8)
    if len(tbuf) == 0:
        tbuf = np.zeros(1, np.uint8)
    return qbuf, qoff, qlen, tbuf, toff, tlen
