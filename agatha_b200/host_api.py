"""Host-level Python mirror of the C ABI (include/agatha_b200.h): whole-job alignment over one or more GPUs,
the stream object that replaces gasal_gpu_storage_t, and the host utilities (bucketing, sharding, cell
accounting, FASTA reader, synthetic workloads). numpy arrays in, numpy arrays out; no torch needed."""
import ctypes

import numpy as np

from ._lib import AgathaError, Params, check, lib, make_params

u8p, u32p, u64p, i32p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32))


def _a(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


def _ptr(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


class JobConfig(ctypes.Structure):
    _fields_ = [("n_devices", ctypes.c_int32), ("devices", i32p), ("batch_alns", ctypes.c_uint32), ("streams_per_device", ctypes.c_int32),
                ("staging_threads", ctypes.c_int32), ("query_ops", u8p), ("target_ops", u8p)]


class JobStats(ctypes.Structure):
    _fields_ = [("seconds_total", ctypes.c_double), ("seconds_kernel_max", ctypes.c_double), ("h2d_bytes", ctypes.c_uint64),
                ("d2h_bytes", ctypes.c_uint64), ("n_batches", ctypes.c_uint32), ("n_devices", ctypes.c_uint32)]


RESULT_DTYPE = np.dtype([("score", "<i4"), ("query_end", "<i4"), ("target_end", "<i4"), ("stop", "<i4"), ("dstop", "<i4")])


def device_count():
    return int(lib().agatha_device_count())


_synth = None


def _synth_lib():
    """libagatha_synth.so: the workload generator is bench / test tooling and lives outside the product library."""
    global _synth
    if _synth is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libagatha_synth.so")
        if not os.path.exists(path):
            raise AgathaError("libagatha_synth.so not found: run `python -m agatha_b200.build`")
        _synth = ctypes.CDLL(path)
    return _synth


def synth_pairs(profile, seed, n_pairs, first_pair=0, n_threads=0):
    """Deterministic synthetic pairs: profile 1=C1, 2=ONT-like, 3=HiFi-like, 4=heavy tail.
    Returns dict(qbuf, qoff, qlen, tbuf, toff, tlen): ASCII bases, byte offsets (uint64), lengths (uint32)."""
    L = _synth_lib()
    n = int(n_pairs)
    qlen = np.zeros(n, np.uint32); tlen = np.zeros(n, np.uint32)
    qoff = np.zeros(n, np.uint64); toff = np.zeros(n, np.uint64)
    args = lambda qb, qc, tb, tc: (ctypes.c_int32(profile), ctypes.c_uint64(seed), ctypes.c_uint64(first_pair), ctypes.c_uint64(n),
                                   _ptr(qlen, u32p), _ptr(tlen, u32p), _ptr(qoff, u64p), _ptr(toff, u64p),
                                   qb, ctypes.c_uint64(qc), tb, ctypes.c_uint64(tc), ctypes.c_int32(n_threads))
    if L.agatha_synth_pairs(*args(None, 0, None, 0)) != 0:
        raise AgathaError("agatha_synth_pairs: bad arguments")
    qtot = int(qlen.sum(dtype=np.uint64)); ttot = int(tlen.sum(dtype=np.uint64))
    qbuf = np.empty(max(qtot, 1), np.uint8); tbuf = np.empty(max(ttot, 1), np.uint8)
    if L.agatha_synth_pairs(*args(_ptr(qbuf, u8p), qtot, _ptr(tbuf, u8p), ttot)) != 0:
        raise AgathaError("agatha_synth_pairs failed")
    return dict(qbuf=qbuf, qoff=qoff, qlen=qlen, tbuf=tbuf, toff=toff, tlen=tlen)


def bucket_order(qlen, tlen, band_width):
    qlen, tlen = _a(qlen, np.uint32), _a(tlen, np.uint32)
    out = np.empty(len(qlen), np.uint32)
    check(lib().agatha_bucket_order(_ptr(qlen, u32p), _ptr(tlen, u32p), ctypes.c_uint32(len(qlen)), ctypes.c_int32(band_width), _ptr(out, u32p)))
    return out


def shard_pairs(qlen, tlen, band_width, n_shards):
    qlen, tlen = _a(qlen, np.uint32), _a(tlen, np.uint32)
    out = np.empty(len(qlen), np.int32)
    check(lib().agatha_shard_pairs(_ptr(qlen, u32p), _ptr(tlen, u32p), ctypes.c_uint64(len(qlen)), ctypes.c_int32(band_width),
                                   ctypes.c_int32(n_shards), _ptr(out, i32p)))
    return out


def count_cells(qlen, tlen, band_width, dstop=None):
    """In-band real cells on anti-diagonals < dstop per pair (GCUPS numerator); returns (per_pair, total)."""
    qlen, tlen = _a(qlen, np.uint32), _a(tlen, np.uint32)
    ds = _a(dstop, np.int32) if dstop is not None else None
    out = np.empty(len(qlen), np.uint64)
    tot = ctypes.c_uint64(0)
    check(lib().agatha_count_cells(_ptr(qlen, u32p), _ptr(tlen, u32p), _ptr(ds, i32p), ctypes.c_uint64(len(qlen)),
                                   ctypes.c_int32(band_width), _ptr(out, u64p), ctypes.byref(tot)))
    return out, int(tot.value)


def fasta_load(query_path, target_path):
    """Lock-step FASTA reader for the reference's '>>> idx' format (test_prog.cpp:94-149)."""
    L = lib()
    L.agatha_fasta_load.restype = ctypes.c_void_p
    h = L.agatha_fasta_load(query_path.encode(), target_path.encode())
    if not h:
        raise AgathaError(L.agatha_last_error().decode())
    h = ctypes.c_void_p(h)
    try:
        L.agatha_fasta_count.restype = ctypes.c_uint64
        L.agatha_fasta_max_len.restype = ctypes.c_uint32
        n = int(L.agatha_fasta_count(h))

        def arr(fn, ty, dt, cnt):
            f = getattr(L, fn); f.restype = ty
            p = f(h)
            return np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True) if cnt else np.zeros(0, dt)
        qlen = arr("agatha_fasta_query_lens", u32p, np.uint32, n); tlen = arr("agatha_fasta_target_lens", u32p, np.uint32, n)
        qoff = arr("agatha_fasta_query_offsets", u64p, np.uint64, n); toff = arr("agatha_fasta_target_offsets", u64p, np.uint64, n)
        qtot = int(qlen.sum(dtype=np.uint64)); ttot = int(tlen.sum(dtype=np.uint64))
        out = dict(qbuf=arr("agatha_fasta_query_bases", u8p, np.uint8, qtot), tbuf=arr("agatha_fasta_target_bases", u8p, np.uint8, ttot),
                   qoff=qoff, toff=toff, qlen=qlen, tlen=tlen,
                   qop=arr("agatha_fasta_query_ops", u8p, np.uint8, n), top=arr("agatha_fasta_target_ops", u8p, np.uint8, n),
                   max_len=int(L.agatha_fasta_max_len(h)))
    finally:
        L.agatha_fasta_free(h)
    return out


def write_fasta(path, buf, off, lens, ops=None):
    """Writes the reference's two-line '>>> idx' records (README.md:41-51). ops (optional, 0..3 per record) selects the
    first header character '>', '<', '/' or '+' = forward, reverse, complement, reverse-complement (test_prog.cpp:83-92)."""
    with open(path, "wb") as f:
        for i in range(len(lens)):
            f.write(b"%c>> %d\n" % (b"></+"[int(ops[i]) & 3] if ops is not None else b">"[0], i + 1))
            f.write(bytes(buf[int(off[i]):int(off[i]) + int(lens[i])]))
            f.write(b"\n")


def align_job(qbuf, qoff, qlen, tbuf, toff, tlen, params, devices=None, n_devices=0, batch_alns=0, streams_per_device=0, staging_threads=0,
              query_ops=None, target_ops=None):
    """agatha_align_job: pairs (unpadded ASCII, byte offsets) -> structured results in input order, plus stats dict.
    query_ops / target_ops: optional per-pair op bytes (bit 0 = reverse, bit 1 = complement; test_prog.cpp:83-92)."""
    qbuf, tbuf = _a(qbuf, np.uint8), _a(tbuf, np.uint8)
    qoff, toff = _a(qoff, np.uint64), _a(toff, np.uint64)
    qlen, tlen = _a(qlen, np.uint32), _a(tlen, np.uint32)
    n = len(qlen)
    p = params if isinstance(params, Params) else make_params(**params)
    cfg = JobConfig()
    dev_arr = None
    if devices is not None:
        dev_arr = _a(devices, np.int32)
        cfg.n_devices = len(dev_arr); cfg.devices = _ptr(dev_arr, i32p)
    else:
        cfg.n_devices = n_devices
    cfg.batch_alns = batch_alns; cfg.streams_per_device = streams_per_device; cfg.staging_threads = staging_threads
    qops = tops = None
    if query_ops is not None or target_ops is not None:
        qops = _a(query_ops if query_ops is not None else np.zeros(n, np.uint8), np.uint8)
        tops = _a(target_ops if target_ops is not None else np.zeros(n, np.uint8), np.uint8)
        if len(qops) != n or len(tops) != n:
            raise ValueError("ops must have one byte per pair")
        cfg.query_ops = _ptr(qops, u8p); cfg.target_ops = _ptr(tops, u8p)
    res = {k: np.zeros(n, np.int32) for k in ("score", "query_end", "target_end", "stop", "dstop")}
    st = JobStats()
    check(lib().agatha_align_job(_ptr(qbuf, u8p), _ptr(qoff, u64p), _ptr(qlen, u32p), _ptr(tbuf, u8p), _ptr(toff, u64p), _ptr(tlen, u32p),
                                 ctypes.c_uint64(n), ctypes.byref(p), ctypes.byref(cfg),
                                 _ptr(res["score"], i32p), _ptr(res["query_end"], i32p), _ptr(res["target_end"], i32p),
                                 _ptr(res["stop"], i32p), _ptr(res["dstop"], i32p), ctypes.byref(st)))
    out = np.zeros(n, RESULT_DTYPE)
    for k in res:
        out[k] = res[k]
    stats = {k: getattr(st, k) for k, _ in JobStats._fields_}
    return out, stats


def align_job_starts(qbuf, qoff, qlen, tbuf, toff, tlen, params, devices=None, batch_alns=0):
    """agatha_align_job_starts: results of align_job plus (query_start, target_start) int32 arrays (GASAL2's WITH_START
    convention: start of the best alignment that ends in the reported end cell)."""
    qbuf, tbuf = _a(qbuf, np.uint8), _a(tbuf, np.uint8)
    qoff, toff = _a(qoff, np.uint64), _a(toff, np.uint64)
    qlen, tlen = _a(qlen, np.uint32), _a(tlen, np.uint32)
    n = len(qlen)
    p = params if isinstance(params, Params) else make_params(**params)
    cfg = JobConfig()
    dev_arr = None
    if devices is not None:
        dev_arr = _a(devices, np.int32)
        cfg.n_devices = len(dev_arr); cfg.devices = _ptr(dev_arr, i32p)
    cfg.batch_alns = batch_alns
    res = {k: np.zeros(n, np.int32) for k in ("score", "query_end", "target_end", "stop", "dstop", "query_start", "target_start")}
    st = JobStats()
    check(lib().agatha_align_job_starts(_ptr(qbuf, u8p), _ptr(qoff, u64p), _ptr(qlen, u32p), _ptr(tbuf, u8p), _ptr(toff, u64p), _ptr(tlen, u32p),
                                        ctypes.c_uint64(n), ctypes.byref(p), ctypes.byref(cfg),
                                        _ptr(res["score"], i32p), _ptr(res["query_end"], i32p), _ptr(res["target_end"], i32p),
                                        _ptr(res["stop"], i32p), _ptr(res["dstop"], i32p),
                                        _ptr(res["query_start"], i32p), _ptr(res["target_start"], i32p), ctypes.byref(st)))
    out = np.zeros(n, RESULT_DTYPE)
    for k in RESULT_DTYPE.names:
        out[k] = res[k]
    return out, res["query_start"], res["target_start"]


def align_pairs(pairs, params, **kw):
    """[(query, target), ...] as str/bytes/uint8 arrays -> structured results (convenience for tests)."""
    qs = [np.frombuffer(q.encode() if isinstance(q, str) else bytes(q), np.uint8) if not isinstance(q, np.ndarray) else q for q, _ in pairs]
    ts = [np.frombuffer(t.encode() if isinstance(t, str) else bytes(t), np.uint8) if not isinstance(t, np.ndarray) else t for _, t in pairs]
    qlen = np.array([len(x) for x in qs], np.uint32); tlen = np.array([len(x) for x in ts], np.uint32)
    qoff = np.zeros(len(qs), np.uint64); toff = np.zeros(len(ts), np.uint64)
    if len(qs) > 1:
        qoff[1:] = np.cumsum(qlen[:-1], dtype=np.uint64); toff[1:] = np.cumsum(tlen[:-1], dtype=np.uint64)
    qbuf = np.concatenate(qs) if len(qs) and qlen.sum() else np.zeros(1, np.uint8)
    tbuf = np.concatenate(ts) if len(ts) and tlen.sum() else np.zeros(1, np.uint8)
    return align_job(qbuf, qoff, qlen, tbuf, toff, tlen, params, **kw)[0]


class Stream:
    """agatha_stream_t: the engine's replacement for one gasal_gpu_storage_t (pinned staging + device buffers + CUDA stream)."""

    def __init__(self, device=0, max_alns=8192, max_query_bytes=1 << 20, max_target_bytes=1 << 20):
        L = lib()
        L.agatha_stream_create.restype = ctypes.c_void_p
        h = L.agatha_stream_create(ctypes.c_int(device), ctypes.c_uint32(max_alns), ctypes.c_uint64(max_query_bytes), ctypes.c_uint64(max_target_bytes))
        if not h:
            raise AgathaError(L.agatha_last_error().decode())
        self.h = ctypes.c_void_p(h)
        self.n = 0

    def close(self):
        if self.h:
            lib().agatha_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, fn, ty, cnt):
        f = getattr(lib(), fn); f.restype = ty
        return np.ctypeslib.as_array(f(self.h), shape=(cnt,))

    def fill(self, qbuf, qoff, qlen, tbuf, toff, tlen):
        """Stage a batch laid out like the reference's host batch (offsets in bases, multiples of 8, 'N' padded)."""
        n = len(qlen)
        check(lib().agatha_stream_reserve(self.h, ctypes.c_uint32(n), ctypes.c_uint64(len(qbuf)), ctypes.c_uint64(len(tbuf))))
        self._view("agatha_stream_query_bases", u8p, len(qbuf))[:] = qbuf
        self._view("agatha_stream_target_bases", u8p, len(tbuf))[:] = tbuf
        self._view("agatha_stream_query_offsets", u32p, n)[:] = qoff
        self._view("agatha_stream_target_offsets", u32p, n)[:] = toff
        self._view("agatha_stream_query_lens", u32p, n)[:] = qlen
        self._view("agatha_stream_target_lens", u32p, n)[:] = tlen
        self.n = n
        self.qbytes, self.tbytes = len(qbuf), len(tbuf)

    def set_ops(self, query_ops, target_ops):
        """Per-sequence op bytes of the staged batch (gasal_op_fill, interfaces.cpp:69-84); used by submit(ops=True)."""
        self._view("agatha_stream_query_ops", u8p, self.n)[:] = query_ops
        self._view("agatha_stream_target_ops", u8p, self.n)[:] = target_ops

    def submit(self, params, qbytes=None, tbytes=None, n=None, ops=False):
        p = params if isinstance(params, Params) else make_params(**params)
        fn = lib().agatha_stream_submit_ops if ops else lib().agatha_stream_submit
        check(fn(self.h, ctypes.c_uint64(self.qbytes if qbytes is None else qbytes),
                                         ctypes.c_uint64(self.tbytes if tbytes is None else tbytes),
                                         ctypes.c_uint32(self.n if n is None else n), ctypes.byref(p)))

    def poll(self):
        return int(lib().agatha_stream_poll(self.h))

    def wait(self):
        check(lib().agatha_stream_wait(self.h))

    def timings(self):
        ms = (ctypes.c_float * 3)()
        check(lib().agatha_stream_timings(self.h, ms))
        return dict(h2d_pack_ms=ms[0], kernel_ms=ms[1], total_ms=ms[2])

    def results(self):
        out = np.zeros(self.n, RESULT_DTYPE)
        for k, fn in (("score", "agatha_stream_scores"), ("query_end", "agatha_stream_query_ends"), ("target_end", "agatha_stream_target_ends"),
                      ("stop", "agatha_stream_stops"), ("dstop", "agatha_stream_dstops")):
            out[k] = self._view(fn, i32p, self.n)
        return out


def stage_batch(buf, off, lens, ids=None, n_threads=0):
    """agatha_stage_batch: unpadded sequences -> the reference's padded host-batch layout.
    Returns (staged uint8 array, offsets uint32 in bases, lens uint32)."""
    buf = _a(buf, np.uint8); off = _a(off, np.uint64); lens = _a(lens, np.uint32)
    idp = _a(ids, np.uint64) if ids is not None else None
    n = len(idp) if idp is not None else len(lens)
    L = lib()
    L.agatha_staged_bytes.restype = ctypes.c_uint64
    need = int(L.agatha_staged_bytes(_ptr(lens, u32p), _ptr(idp, u64p), ctypes.c_uint64(n)))
    dst = np.empty(need, np.uint8)
    doff = np.empty(n, np.uint32); dlen = np.empty(n, np.uint32)
    nbytes = ctypes.c_uint64(0)
    check(L.agatha_stage_batch(_ptr(buf, u8p), _ptr(off, u64p), _ptr(lens, u32p), _ptr(idp, u64p), ctypes.c_uint64(n),
                               _ptr(dst, u8p), ctypes.c_uint64(need), _ptr(doff, u32p), _ptr(dlen, u32p), ctypes.byref(nbytes),
                               ctypes.c_int32(n_threads)))
    return dst[:int(nbytes.value)], doff, dlen


def pack_batch(buf, off, lens, is_target, ids=None, ops=None, n_threads=0):
    """agatha_pack_batch: unpadded ASCII sequences -> the packed device format (8 bases per uint32 word) with the reference's
    batch layout (each sequence at a multiple of 8 bases, 'N' padded), ops applied. Returns (words, offsets in bases, lens)."""
    buf = _a(buf, np.uint8); off = _a(off, np.uint64); lens = _a(lens, np.uint32)
    idp = _a(ids, np.uint64) if ids is not None else None
    opp = _a(ops, np.uint8) if ops is not None else None
    n = len(idp) if idp is not None else len(lens)
    L = lib()
    L.agatha_staged_bytes.restype = ctypes.c_uint64
    need = int(L.agatha_staged_bytes(_ptr(lens, u32p), _ptr(idp, u64p), ctypes.c_uint64(n)))
    dst = np.empty(need // 8, np.uint32)
    doff = np.empty(n, np.uint32); dlen = np.empty(n, np.uint32)
    nb = ctypes.c_uint64(0)
    check(L.agatha_pack_batch(_ptr(buf, u8p), _ptr(off, u64p), _ptr(lens, u32p), _ptr(idp, u64p), _ptr(opp, u8p), ctypes.c_uint64(n),
                              ctypes.c_int32(1 if is_target else 0), _ptr(dst, u32p), ctypes.c_uint64(len(dst)),
                              _ptr(doff, u32p), _ptr(dlen, u32p), ctypes.byref(nb), ctypes.c_int32(n_threads)))
    return dst[:int(nb.value) // 8], doff, dlen
