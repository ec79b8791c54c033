"""CPU tests: the library loads and exports every symbol include/agatha_b200.h declares, the host logic (FASTA reader,
synthetic workloads, bucketing, sharding, staging, cell accounting) is right, and the product path fails loudly
without a CUDA device. No compute calls here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ag():
    from agatha_b200 import build
    build.build()
    import agatha_b200
    return agatha_b200


def test_abi_exports_every_declared_symbol(ag):
    hdr = open(os.path.join(ROOT, "include", "agatha_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)          # declarations only, not the citations in comments
    names = sorted(set(re.findall(r"\b(agatha_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40
    L = ctypes.CDLL(ag.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_shim_exports_reference_api(ag):
    out = subprocess.run(["nm", "-DC", ag.lib_path()], capture_output=True, text=True).stdout
    for fn in ("gasal_init_gpu_storage_v", "gasal_init_streams", "gasal_host_batch_fill", "gasal_op_fill", "gasal_host_alns_resize",
               "gasal_set_device", "gasal_copy_subst_scores", "gasal_aln_async", "gasal_is_aln_async_done", "gasal_destroy_streams",
               "gasal_destroy_gpu_storage_v", "Parameters::parse"):
        assert re.search(r" T %s\(" % re.escape(fn), out), fn


@pytest.mark.skipif(not os.path.isdir("/root/reference/AGAThA/test_prog"), reason="reference not mounted")
def test_reference_driver_compiles_unchanged_against_shim(ag, tmp_path):
    tp = tmp_path / "test_prog"
    tp.mkdir()
    os.symlink(os.path.join(ROOT, "include", "gasal_compat"), tmp_path / "include")
    os.symlink("/root/reference/AGAThA/test_prog/test_prog.cpp", tp / "test_prog.cpp")
    os.symlink("/root/reference/AGAThA/test_prog/Timer.h", tp / "Timer.h")
    r = subprocess.run(["g++", "-O1", "-std=c++11", "-fopenmp", "-I/usr/local/cuda/include", "-c", "test_prog.cpp", "-o", "t.o"], cwd=tp, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run(["g++", "-o", "manual", "t.o", "-L" + os.path.dirname(ag.lib_path()), "-lagatha_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-fopenmp"],
                       cwd=tp, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_no_gpu_fails_loudly(ag):
    if ag.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ag.AgathaError, match="no CUDA device"):
        ag.align_pairs([("ACGT", "ACGT")], ag.make_params())
    with pytest.raises(ag.AgathaError, match="no CUDA device"):
        ag.Stream()


def test_fasta_reader_roundtrip_and_lockstep(ag, tmp_path):
    d = ag.synth_pairs(1, 7, 50)
    qf, tf = str(tmp_path / "q.fa"), str(tmp_path / "t.fa")
    ag.write_fasta(qf, d["qbuf"], d["qoff"], d["qlen"])
    ag.write_fasta(tf, d["tbuf"], d["toff"], d["tlen"])
    f = ag.fasta_load(qf, tf)
    assert (f["qlen"] == d["qlen"]).all() and (f["tlen"] == d["tlen"]).all()
    assert bytes(f["qbuf"]) == bytes(d["qbuf"]) and bytes(f["tbuf"]) == bytes(d["tbuf"])
    assert f["max_len"] == max(d["qlen"].max(), d["tlen"].max())
    assert (f["qop"] == 0).all()
    # multi-line records and the op characters of test_prog.cpp:83-92
    open(qf, "w").write(">>> 1\nACGT\nACGT\n<<< 2\nTTTT\n/ 3\nGG\n+ 4\nC\n")
    open(tf, "w").write(">>> 1\nAC\nGTAC\n>>> 2\nTTAA\n>>> 3\nGGG\n>>> 4\nCA\n")
    f = ag.fasta_load(qf, tf)
    assert list(f["qlen"]) == [8, 4, 2, 1] and list(f["tlen"]) == [6, 4, 3, 2]
    assert list(f["qop"]) == [0, 1, 2, 3] and bytes(f["qbuf"][:8]) == b"ACGTACGT" and bytes(f["tbuf"][:6]) == b"ACGTAC"
    with pytest.raises(ag.AgathaError):
        ag.fasta_load(qf, str(tmp_path / "missing.fa"))


def test_synth_is_deterministic_and_shardable(ag):
    a = ag.synth_pairs(2, 2, 300)
    b = ag.synth_pairs(2, 2, 100, first_pair=200)
    assert (a["qlen"][200:] == b["qlen"]).all() and (a["tlen"][200:] == b["tlen"]).all()
    assert bytes(a["tbuf"][int(a["toff"][200]):]) == bytes(b["tbuf"])
    c = ag.synth_pairs(2, 3, 300)
    assert not (a["tlen"] == c["tlen"]).all()
    assert a["tlen"].min() >= 1000 and a["tlen"].max() <= 30000 and 8000 < a["tlen"].mean() < 12000     # C2
    assert set(np.unique(a["qbuf"])) <= set(b"ACGT")
    h = ag.synth_pairs(3, 3, 300)
    assert h["tlen"].min() >= 5000 and h["tlen"].max() <= 25000 and 14000 < h["tlen"].mean() < 16000      # C3
    t = ag.synth_pairs(4, 4, 3000)
    assert t["tlen"].min() >= 1000 and t["tlen"].max() <= 100000 and np.median(t["tlen"]) < 2500          # C4 heavy tail
    # error rate of the ONT-like profile is about 10 %
    assert 0.97 < a["qlen"].sum() / a["tlen"].sum() < 1.03


def test_bucket_order_is_a_permutation_longest_first(ag):
    d = ag.synth_pairs(4, 4, 5000)
    o = ag.bucket_order(d["qlen"], d["tlen"], 751)
    assert sorted(o.tolist()) == list(range(5000))
    cost = np.minimum(d["qlen"], d["tlen"]).astype(np.int64) * np.minimum(1503, np.maximum(d["qlen"], d["tlen"]))
    assert (np.diff(cost[o]) <= 0).all()


def test_shards_are_balanced(ag):
    d = ag.synth_pairs(4, 4, 20000)
    for k in (2, 4, 8):
        s = ag.shard_pairs(d["qlen"], d["tlen"], 751, k)
        cells, _ = ag.count_cells(d["qlen"], d["tlen"], 751)
        load = np.array([cells[s == i].sum() for i in range(k)], dtype=np.float64)
        assert set(np.unique(s)) == set(range(k))
        assert load.max() / load.mean() < 1.02


def test_stage_batch_matches_reference_layout(ag):
    d = ag.synth_pairs(1, 9, 200)
    ids = np.arange(199, -1, -1, dtype=np.uint64)
    s, off, lens = ag.stage_batch(d["qbuf"], d["qoff"], d["qlen"], ids=ids)
    assert (off % 8 == 0).all() and (lens == d["qlen"][::-1]).all()
    for j in (0, 17, 199):
        i = int(ids[j])
        assert bytes(s[off[j]:off[j] + lens[j]]) == bytes(d["qbuf"][int(d["qoff"][i]):int(d["qoff"][i]) + int(lens[j])])
        end = off[j + 1] if j + 1 < len(off) else len(s)
        assert (s[off[j] + lens[j]:end] == ord("N")).all()       # padded with 'N' (host_batch.cpp:143-146)


def test_count_cells_matches_oracle(ag, oracle):
    from oracle import oracle_py as op
    from pairgen import make_pairs
    pairs = make_pairs(3, 120, 5, 900, mixed=True)
    exp = oracle.align_pairs(pairs, op.make_params(band_width=63, z_threshold=100))
    ql = np.array([len(q) for q, _ in pairs], np.uint32); tl = np.array([len(t) for _, t in pairs], np.uint32)
    cells, tot = ag.count_cells(ql, tl, 63, exp["d_stop"])
    assert (cells == exp["cells"]).all() and tot == int(exp["cells"].sum())
    full, _ = ag.count_cells(ql, tl, 63)
    assert all(int(full[i]) == oracle.band_cells(int(ql[i]), int(tl[i]), 63) for i in range(len(ql)))


def test_two_rank_gloo_shards_cover_the_workload(ag, tmp_path):
    """The multi-GPU bench gives rank r the pairs [r*n, (r+1)*n); there is no data-path collective. Two gloo ranks on
    CPU: each generates its shard, the union must equal the single-process workload, and the LPT shard map agrees."""
    script = tmp_path / "rank.py"
    script.write_text('''
import os, sys, hashlib
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import agatha_b200 as ag
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
n = 500
d = ag.synth_pairs(2, 2, n, first_pair=r * n)
h = int(hashlib.sha1(bytes(d["qbuf"]) + bytes(d["tbuf"])).hexdigest()[:12], 16)
t = torch.tensor([h, int(d["qlen"].sum()), int(d["tlen"].sum())], dtype=torch.int64)
g = [torch.zeros(3, dtype=torch.int64) for _ in range(w)]
dist.all_gather(g, t)
if r == 0:
    whole = ag.synth_pairs(2, 2, n * w)
    for k in range(w):
        sl = slice(k * n, (k + 1) * n)
        qb = whole["qbuf"][int(whole["qoff"][k * n]):int(whole["qoff"][k * n]) + int(whole["qlen"][sl].sum())]
        tb = whole["tbuf"][int(whole["toff"][k * n]):int(whole["toff"][k * n]) + int(whole["tlen"][sl].sum())]
        hk = int(hashlib.sha1(bytes(qb) + bytes(tb)).hexdigest()[:12], 16)
        assert hk == int(g[k][0]), k
    s = ag.shard_pairs(whole["qlen"], whole["tlen"], 751, w)
    assert set(np.unique(s)) == set(range(w))
    print("OK")
dist.barrier()
dist.destroy_process_group()
''' % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_host_packer_matches_the_device_pack_kernels(ag):
    """agatha_pack_batch (host, AVX2 + scalar) writes exactly the words pack_kernel + apply_ops_kernel write on the device --
    checked here against those kernels run by the SIMT emulation (tests/emu), including ops, ragged lengths, IUPAC codes,
    lower case, empty sequences and a sub-selection through ids."""
    from emu import emu
    L = emu.lib()
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTNacgtnRYKMSWBDHVU-", dtype=np.uint8)
    lens = np.array([0, 1, 7, 8, 9, 31, 32, 33, 63, 64, 65, 100, 257, 1000, 4097] + list(rng.integers(0, 300, 40)), dtype=np.uint32)
    seqs = [alphabet[rng.integers(0, 5 if i % 3 else len(alphabet), int(n))] for i, n in enumerate(lens)]
    buf = np.concatenate(seqs) if lens.sum() else np.zeros(1, np.uint8)
    off = np.zeros(len(lens), np.uint64); off[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    ops = rng.integers(0, 4, len(lens)).astype(np.uint8)
    ids = rng.permutation(len(lens))[:40].astype(np.uint64)
    for target in (False, True):
        for use_ops in (False, True):
            words, doff, dlen = ag.host_api.pack_batch(buf, off, lens, target, ids=ids, ops=ops if use_ops else None, n_threads=3)
            # the device path: stage as ASCII, pack_kernel, then apply_ops_kernel
            staged, soff, slen = ag.stage_batch(buf, off, lens, ids=ids)
            assert (soff == doff).all() and (slen == dlen).all() and len(words) == len(staged) // 8
            qp = np.zeros(len(staged) // 8 + 64, np.uint32); tp = np.zeros(len(staged) // 8 + 64, np.uint32)
            p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            L.emu_pack(p(staged), ctypes.c_uint64(len(staged)), p(staged), ctypes.c_uint64(len(staged)), p(qp), p(tp))
            if use_ops:
                o = np.ascontiguousarray(ops[ids.astype(np.int64)])
                L.emu_apply_ops(p(staged), p(staged), p(soff), p(soff), p(slen), p(slen), p(o), p(o), ctypes.c_uint32(len(ids)), p(qp), p(tp))
            ref = (tp if target else qp)[:len(words)]
            bad = np.nonzero(ref != words)[0]
            assert len(bad) == 0, (target, use_ops, bad[:5], [hex(int(x)) for x in ref[bad[:3]]], [hex(int(x)) for x in words[bad[:3]]])
