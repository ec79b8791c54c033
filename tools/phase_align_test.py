"""Are mixed phases (prologue / steady state / tail loops of different warps on one SM) what short pairs wait for?
Uniform-length pairs keep every warp of the device in the same loop at the same time; compare with mixed lengths of the
same mean."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agatha_b200 as ag
dev = torch.device("cuda:0")
rng = np.random.default_rng(7)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
def make(lens):
    n = len(lens)
    off = np.zeros(n, np.uint64); off[1:] = np.cumsum(lens[:-1])
    tot = int(lens.sum())
    t = acgt[rng.integers(0, 4, tot)]
    q = t.copy()
    sub = rng.random(tot) < 0.08
    q[sub] = acgt[rng.integers(0, 4, int(sub.sum()))]
    return q, off, lens.astype(np.uint32), t, off.copy(), lens.astype(np.uint32)
def run(name, lens, W=751):
    qb, qo, ql, tb, to, tl = make(lens)
    stq, qoff, qlen = ag.stage_batch(qb, qo, ql); stt, toff, tlen = ag.stage_batch(tb, to, tl)
    qp, tp = ag.pack_device(torch.from_numpy(stq).to(dev), torch.from_numpy(stt).to(dev))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
    order = d(ag.bucket_order(qlen, tlen, W)); p = ag.make_params(band_width=W)
    a = (qp, tp, d(qoff), d(toff), d(qlen), d(tlen), p)
    out = ag.extend_device(*a, order=order); torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ag.extend_device(*a, order=order, out=out); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    _, cells = ag.count_cells(qlen, tlen, W, out["dstop"].cpu().numpy())
    print(name, "pairs", len(lens), "ms %.2f" % min(ms), "GCUPS %.0f" % (cells / min(ms) / 1e6), flush=True)
n = 14208   # 8 jobs per warp slot at 4 CTAs x 4 warps x 148 SMs = 2368 slots -> 6 full waves
run("uniform 4500", np.full(n, 4500))
run("mixed 1000-8000", rng.integers(1000, 8001, n))
run("uniform 2000", np.full(n, 2000))
run("uniform 10000", np.full(7104, 10000))
