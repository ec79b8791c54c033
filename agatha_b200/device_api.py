"""Device-level Python mirror of the C ABI: torch tensors own the memory, the library does the work."""
import ctypes

import numpy as np

from ._lib import Params, check, lib, make_params

PACK_SLACK_WORDS = 64
WORKSPACE_BYTES = 256
N_BYTE = 0x4E  # 'N', the reference's padding symbol (host_batch.cpp:143-146)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("agatha_b200 needs a CUDA device (no CPU fallback)")
    return torch


def _p(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


def launch_count():
    return int(lib().agatha_launch_count())


def measure_int_peak(device=0):
    """agatha_measure_int_peak: dict(alu, fma, mixed) in 1e12 lane-operations/s, measured now on `device`."""
    a, f, m = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
    check(lib().agatha_measure_int_peak(ctypes.c_int(device), ctypes.byref(a), ctypes.byref(f), ctypes.byref(m)))
    return dict(alu=a.value, fma=f.value, mixed=m.value)


def stage_pairs(pairs):
    """Host staging exactly like gasal_host_batch_fill (host_batch.cpp:79-154): every sequence is copied to an
    offset that is a multiple of 8 and padded to a multiple of 8 with 'N'. Returns numpy arrays
    (qbuf, qoff, qlen, tbuf, toff, tlen); offsets are in bases as in the reference."""
    def one(seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        padded = (lens.astype(np.int64) + 7) & ~7
        offs = np.zeros(len(seqs), dtype=np.int64)
        if len(seqs) > 1:
            offs[1:] = np.cumsum(padded[:-1])
        total = int(padded.sum())
        buf = np.full(max(total, 8), N_BYTE, dtype=np.uint8)
        for s, o in zip(seqs, offs):
            a = np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray) else s
            buf[o:o + len(a)] = a
        return buf, offs.astype(np.uint32), lens
    qbuf, qoff, qlen = one([q for q, _ in pairs])
    tbuf, toff, tlen = one([t for _, t in pairs])
    return qbuf, qoff, qlen, tbuf, toff, tlen


def pack_device(q_bases, t_bases, stream=None):
    """uint8 CUDA tensors (lengths multiples of 8) -> (query_packed, target_packed) uint32-as-int32 CUDA tensors."""
    torch = _torch()
    assert q_bases.is_cuda and t_bases.is_cuda and q_bases.dtype == torch.uint8 and t_bases.dtype == torch.uint8
    qb, tb = q_bases.numel(), t_bases.numel()
    qp = torch.empty(qb // 8 + PACK_SLACK_WORDS, dtype=torch.int32, device=q_bases.device)
    tp = torch.empty(tb // 8 + PACK_SLACK_WORDS, dtype=torch.int32, device=q_bases.device)
    st = stream if stream is not None else torch.cuda.current_stream(q_bases.device)
    check(lib().agatha_pack_device(_p(q_bases), ctypes.c_uint64(qb), _p(t_bases), ctypes.c_uint64(tb), _p(qp), _p(tp),
                                   ctypes.c_void_p(st.cuda_stream)))
    return qp, tp


def apply_ops_device(q_bases, t_bases, qoff, toff, qlen, tlen, qops, tops, qp, tp, stream=None):
    """agatha_apply_ops_device: re-pack, in place in (qp, tp), the sequences whose op byte (uint8 CUDA tensors qops/tops;
    bit 0 = reverse, bit 1 = complement) is non-zero. Call after pack_device on the same stream."""
    torch = _torch()
    st = stream if stream is not None else torch.cuda.current_stream(qp.device)
    check(lib().agatha_apply_ops_device(_p(q_bases), _p(t_bases), _p(qoff), _p(toff), _p(qlen), _p(tlen), _p(qops), _p(tops),
                                        ctypes.c_uint32(qlen.numel()), _p(qp), _p(tp), ctypes.c_void_p(st.cuda_stream)))


def extend_device(qp, tp, qoff, toff, qlen, tlen, params, order=None, out=None, workspace=None, stream=None):
    """Launch the extension kernel on already packed, device-resident batches. All tensors int32 on CUDA
    (uint32 values); returns dict(score, query_end, target_end, stop, dstop) of int32 CUDA tensors."""
    torch = _torch()
    n = qlen.numel()
    dev = qp.device
    if out is None:
        out = {k: torch.empty(n, dtype=torch.int32, device=dev) for k in ("score", "query_end", "target_end", "stop", "dstop")}
    if workspace is None:
        workspace = torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    p = params if isinstance(params, Params) else make_params(**params)
    check(lib().agatha_extend_device(_p(qp), _p(tp), _p(qoff), _p(toff), _p(qlen), _p(tlen), _p(order),
                                     ctypes.c_uint32(n), ctypes.byref(p),
                                     _p(out["score"]), _p(out["query_end"]), _p(out["target_end"]),
                                     _p(out.get("stop")), _p(out.get("dstop")), _p(workspace),
                                     ctypes.c_void_p(st.cuda_stream)))
    return out


def align_pairs_device(pairs, params, device="cuda:0", bucket=True, ops=None):
    """Convenience for tests: stage -> H2D -> pack [-> ops] -> extend -> D2H. Returns a structured numpy array.
    ops: optional (query_ops, target_ops) uint8 arrays."""
    torch = _torch()
    qbuf, qoff, qlen, tbuf, toff, tlen = stage_pairs(pairs)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        tq = torch.from_numpy(qbuf).to(dev); tt = torch.from_numpy(tbuf).to(dev)
        qp, tp = pack_device(tq, tt)
        d = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
        if ops is not None:
            u8 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8)).to(dev)
            apply_ops_device(tq, tt, d(qoff), d(toff), d(qlen), d(tlen), u8(ops[0]), u8(ops[1]), qp, tp)
        order = None
        if bucket:
            cost = np.minimum(qlen, tlen).astype(np.int64)
            order = d(np.argsort(-cost, kind="stable").astype(np.uint32))
        out = extend_device(qp, tp, d(qoff), d(toff), d(qlen), d(tlen), params, order=order)
        torch.cuda.synchronize(dev)
    res = np.zeros(len(pairs), dtype=[("score", "<i4"), ("query_end", "<i4"), ("target_end", "<i4"), ("stop", "<i4"), ("dstop", "<i4")])
    for k in ("score", "query_end", "target_end", "stop", "dstop"):
        res[k] = out[k].cpu().numpy()
    return res
