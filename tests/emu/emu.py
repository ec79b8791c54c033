"""TEST INFRASTRUCTURE -- loader of the SIMT emulation (tests/emu/emu_main.cpp): the product's CUDA kernel headers
compiled for the CPU, one fiber per CUDA thread. Used by the `-m "not gpu"` tests to compare the kernel LOGIC with the
oracle where there is no GPU. Nothing under agatha_b200/ imports this."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "agatha_b200", "csrc")
SO = os.path.join(HERE, "libagatha_emu.so")
STAMP = os.path.join(HERE, "libagatha_emu.stamp")
N_BYTE = 0x4E

RESULT_DTYPE = np.dtype([("score", "<i4"), ("query_end", "<i4"), ("target_end", "<i4"), ("stop", "<i4"), ("dstop", "<i4")])


class Params(ctypes.Structure):
    _fields_ = [("match", ctypes.c_int32), ("mismatch", ctypes.c_int32), ("gap_open", ctypes.c_int32),
                ("gap_extend", ctypes.c_int32), ("slice_width", ctypes.c_int32),
                ("z_threshold", ctypes.c_int32), ("band_width", ctypes.c_int32)]


DEFAULT_PARAMS = dict(match=1, mismatch=4, gap_open=6, gap_extend=2, slice_width=3, z_threshold=400, band_width=751)


def make_params(**kw):
    d = dict(DEFAULT_PARAMS)
    d.update(kw)
    return Params(**d)


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(HERE, f) for f in ("emu_main.cpp", "cuda_runtime.h")]
    files += sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(ROOT, "include", "agatha_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, opt="-O1"):
    dig = _digest() + opt
    if not force and os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return SO
    cmd = ["g++", opt, "-std=c++17", "-shared", "-fPIC", "-w", "-I" + HERE, "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
           "-o", SO, os.path.join(HERE, "emu_main.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stdout[-4000:])
    with open(STAMP, "w") as f:
        f.write(dig)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_last_error.restype = ctypes.c_char_p
    return _lib


def stage_pairs(pairs):
    """gasal_host_batch_fill's layout (host_batch.cpp:79-154): multiples of 8, 'N' padding; offsets in bases."""
    def one(seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        padded = (lens.astype(np.int64) + 7) & ~7
        offs = np.zeros(len(seqs), dtype=np.int64)
        if len(seqs) > 1:
            offs[1:] = np.cumsum(padded[:-1])
        buf = np.full(max(int(padded.sum()), 8), N_BYTE, dtype=np.uint8)
        for s, o in zip(seqs, offs):
            a = s if isinstance(s, np.ndarray) else np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8)
            buf[o:o + len(a)] = a
        return buf, offs.astype(np.uint32), lens
    qbuf, qoff, qlen = one([q for q, _ in pairs])
    tbuf, toff, tlen = one([t for _, t in pairs])
    return qbuf, qoff, qlen, tbuf, toff, tlen


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def align_pairs(pairs, params, s16_mode=-1, ops=None, bucket=True):
    """stage -> pack_kernel -> [apply_ops_kernel] -> extend_kernel, all emulated. Returns a structured array."""
    L = lib()
    L.emu_set_s16_mode(ctypes.c_int(s16_mode))
    qbuf, qoff, qlen, tbuf, toff, tlen = stage_pairs(pairs)
    n = len(pairs)
    qp = np.zeros(len(qbuf) // 8 + 64, dtype=np.uint32)
    tp = np.zeros(len(tbuf) // 8 + 64, dtype=np.uint32)
    L.emu_pack(_p(qbuf), ctypes.c_uint64(len(qbuf)), _p(tbuf), ctypes.c_uint64(len(tbuf)), _p(qp), _p(tp))
    if ops is not None:
        qo = np.ascontiguousarray(ops[0], dtype=np.uint8); to = np.ascontiguousarray(ops[1], dtype=np.uint8)
        L.emu_apply_ops(_p(qbuf), _p(tbuf), _p(qoff), _p(toff), _p(qlen), _p(tlen), _p(qo), _p(to), ctypes.c_uint32(n), _p(qp), _p(tp))
    order = None
    if bucket:
        order = np.argsort(-np.minimum(qlen, tlen).astype(np.int64), kind="stable").astype(np.uint32)
    out = {k: np.zeros(n, dtype=np.int32) for k in RESULT_DTYPE.names}
    p = params if isinstance(params, Params) else make_params(**params)
    rc = L.emu_extend(_p(qp), _p(tp), _p(qoff), _p(toff), _p(qlen), _p(tlen), _p(order), ctypes.c_uint32(n), ctypes.byref(p),
                      _p(out["score"]), _p(out["query_end"]), _p(out["target_end"]), _p(out["stop"]), _p(out["dstop"]))
    if rc != 0:
        raise RuntimeError("emu_extend failed (%d): %s" % (rc, L.emu_last_error().decode()))
    res = np.zeros(n, dtype=RESULT_DTYPE)
    for k in RESULT_DTYPE.names:
        res[k] = out[k]
    return res


def last_redo_count():
    """Pairs of the last align_pairs() call that the packed kernel handed over to the general kernel."""
    L = lib()
    L.emu_last_redo_count.restype = ctypes.c_uint32
    return int(L.emu_last_redo_count())


def last_used_packed():
    """True when the last align_pairs() call ran the packed kernel (extend16_kernel.cuh) before the general one."""
    return bool(lib().emu_last_used_packed())
