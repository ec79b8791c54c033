"""GPU tests of the host-level C ABI: whole-job call, stream protocol (the gasal_aln_async / gasal_is_aln_async_done
equivalents), argument checks, growth, multi-batch double buffering -- all checked against the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle_py as op
from pairgen import make_pairs

pytestmark = pytest.mark.gpu


def _exp(oracle, pairs, **pkw):
    return oracle.align_pairs(pairs, op.make_params(**pkw))


def _same(got, exp):
    return all((got[a] == exp[b]).all() for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")))


def test_align_job_many_small_batches(oracle):
    import agatha_b200 as ag
    pairs = make_pairs(31, 700, 5, 900, mixed=True)
    got = ag.align_pairs(pairs, ag.make_params(band_width=63), batch_alns=64, streams_per_device=3)
    assert _same(got, _exp(oracle, pairs, band_width=63))


def test_align_job_synthetic_profiles_vs_oracle(oracle):
    import agatha_b200 as ag
    for prof, n, W in ((1, 400, 751), (2, 60, 751), (4, 300, 751)):
        d = ag.synth_pairs(prof, prof, n)
        res, stats = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], ag.make_params(band_width=W))
        exp = oracle.align_batch(d["qbuf"], d["qoff"].astype(np.uint32), d["qlen"], d["tbuf"], d["toff"].astype(np.uint32), d["tlen"], op.make_params(band_width=W))
        assert _same(res, exp), "profile %d" % prof
        assert stats["h2d_bytes"] > 0 and stats["n_batches"] >= 1


def test_full_size_properties_100k_pairs():
    """BASELINE size (C2, 100k pairs): size-independent properties instead of the oracle -- idempotence, order invariance,
    self-alignment of every target scores len*match and ends on the last base, coordinates inside the sequences."""
    import agatha_b200 as ag
    d = ag.synth_pairs(2, 2, 100000)
    p = ag.make_params()
    a, _ = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p)
    b, _ = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, batch_alns=5000)
    assert (a == b).all()
    assert (a["query_end"] < d["qlen"].astype(np.int64)).all() and (a["target_end"] < d["tlen"].astype(np.int64)).all()
    assert (a["score"] > 0).all() and (a["score"] <= np.minimum(d["qlen"], d["tlen"])).all()
    s, _ = ag.align_job(d["tbuf"], d["toff"], d["tlen"], d["tbuf"], d["toff"], d["tlen"], p)
    assert (s["score"] == d["tlen"]).all() and (s["query_end"] == d["tlen"] - 1).all() and (s["target_end"] == d["tlen"] - 1).all()
    assert (s["stop"] == 0).all()
    # checksum of checksums: a permutation of the pairs permutes the results
    perm = np.random.default_rng(0).permutation(20000)
    sub = lambda k: np.concatenate([d[k + "buf"][int(d[k + "off"][i]):int(d[k + "off"][i]) + int(d[k + "len"][i])] for i in perm])
    ql, tl = d["qlen"][perm], d["tlen"][perm]
    qo = np.concatenate([[0], np.cumsum(ql[:-1], dtype=np.uint64)]).astype(np.uint64); to = np.concatenate([[0], np.cumsum(tl[:-1], dtype=np.uint64)]).astype(np.uint64)
    c, _ = ag.align_job(sub("q"), qo, ql, sub("t"), to, tl, p)
    assert (c == a[perm]).all()


def test_stream_protocol_and_argument_checks(oracle):
    import agatha_b200 as ag
    pairs = make_pairs(77, 300, 20, 1200, mixed=True)
    qbuf, qoff, qlen, tbuf, toff, tlen = ag.stage_pairs(pairs)
    s = ag.Stream(device=0, max_alns=16, max_query_bytes=64, max_target_bytes=64)   # far too small: must grow
    assert s.poll() == -2                                  # nothing submitted (gasal_align.cu:279)
    s.fill(qbuf, qoff, qlen, tbuf, toff, tlen)
    p = ag.make_params(band_width=127)
    s.submit(p)
    rc = s.poll()
    assert rc in (-1, 0)
    s.wait()
    assert s.poll() == -2
    got = s.results()
    assert _same(got, _exp(oracle, pairs, band_width=127))
    t = s.timings()
    assert t["kernel_ms"] > 0 and t["total_ms"] >= t["kernel_ms"]
    # the reference's argument checks (gasal_align.cu:33-53)
    for kw, msg in ((dict(n=0), "actual_n_alns"), (dict(qbytes=0), "actual_query_batch_bytes"), (dict(tbytes=0), "actual_target_batch_bytes"),
                    (dict(qbytes=12), "multiple of 8"), (dict(tbytes=20), "multiple of 8")):
        with pytest.raises(ag.AgathaError, match=msg):
            s.submit(p, **kw)
    with pytest.raises(ag.AgathaError, match="band_width"):
        s.submit(ag.make_params(band_width=20000))
    # reuse after an error, second batch on the same stream
    s.submit(ag.make_params(band_width=63))
    s.wait()
    assert _same(s.results(), _exp(oracle, pairs, band_width=63))
    s.close()


def test_two_streams_overlap(oracle):
    import agatha_b200 as ag
    A = make_pairs(5, 200, 100, 2500, mixed=True)
    B = make_pairs(6, 200, 100, 2500, mixed=True)
    sa, sb = ag.Stream(), ag.Stream()
    sa.fill(*ag.stage_pairs(A)); sb.fill(*ag.stage_pairs(B))
    p = ag.make_params()
    sa.submit(p); sb.submit(p)
    sb.wait(); sa.wait()
    assert _same(sa.results(), _exp(oracle, A)) and _same(sb.results(), _exp(oracle, B))
    sa.close(); sb.close()


def test_job_sharded_over_all_visible_gpus(oracle):
    """agatha_align_job over every GPU of the box (one worker thread + streams per device, LPT sharding, no collective):
    same results as one device and as the oracle. Runs with one GPU too (then it only checks the device list path)."""
    import agatha_b200 as ag
    ndev = ag.device_count()
    d = ag.synth_pairs(4, 11, 1500)                 # heavy tail + early Z-drop: uneven work, exercises the balancing
    p = ag.make_params()
    one, _ = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=[0], batch_alns=256)
    allg, stats = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=list(range(ndev)), batch_alns=256)
    assert (one == allg).all()
    assert stats["n_devices"] == ndev
    sub = np.arange(0, 1500, 7)
    exp = oracle.align_pairs([(d["qbuf"][int(d["qoff"][i]):int(d["qoff"][i]) + int(d["qlen"][i])],
                               d["tbuf"][int(d["toff"][i]):int(d["toff"][i]) + int(d["tlen"][i])]) for i in sub], op.make_params())
    assert _same(allg[sub], exp)
    if ndev > 1:
        rev, _ = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=list(range(ndev))[::-1])
        assert (rev == one).all()


def test_start_positions_reverse_pass_vs_oracle(oracle):
    """agatha_align_job_starts: GASAL2's WITH_START convention (gasal.h:36-39) -- the reference declares the fields but never
    fills them (res.cpp:27-28), so this row is pinned by the oracle only: forward extension, then the oracle again on the
    reversed prefixes that end in the reported cell, Z-drop off."""
    import agatha_b200 as ag
    from pairgen import make_pairs
    pairs = make_pairs(808, 160, 1, 2500, mixed=True) + [("", "ACGT"), ("A", "A"), ("T", "A"), ("ACGTACGT", "ACGTACGT")]
    # junk in front of a good alignment: the start moves away from the origin
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(20):
        core = acgt[rng.integers(0, 4, int(rng.integers(300, 1500)))]
        pairs.append((np.concatenate([acgt[rng.integers(0, 4, 40)], core]), np.concatenate([acgt[rng.integers(0, 4, 40)], core])))
    for pkw in (dict(band_width=751), dict(band_width=127, z_threshold=100), dict(band_width=255, match=2, mismatch=3, gap_open=4, gap_extend=1)):
        qs = [np.frombuffer(q.encode(), np.uint8) if isinstance(q, str) else q for q, _ in pairs]
        ts = [np.frombuffer(t.encode(), np.uint8) if isinstance(t, str) else t for _, t in pairs]
        qlen = np.array([len(x) for x in qs], np.uint32); tlen = np.array([len(x) for x in ts], np.uint32)
        qoff = np.concatenate([[0], np.cumsum(qlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        toff = np.concatenate([[0], np.cumsum(tlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        qbuf = np.concatenate([x for x in qs if len(x)] or [np.zeros(1, np.uint8)]); tbuf = np.concatenate([x for x in ts if len(x)] or [np.zeros(1, np.uint8)])
        res, qstart, tstart = ag.align_job_starts(qbuf, qoff, qlen, tbuf, toff, tlen, ag.make_params(**pkw))
        fwd = oracle.align_pairs(list(zip(qs, ts)), op.make_params(**pkw))
        assert (res["score"] == fwd["score"]).all() and (res["query_end"] == fwd["query_end"]).all() and (res["target_end"] == fwd["target_end"]).all()
        rev_pairs, live = [], []
        for i, (q, t) in enumerate(zip(qs, ts)):
            if len(q) and len(t) and fwd["score"][i] > 0:
                rev_pairs.append((q[:fwd["query_end"][i] + 1][::-1].copy(), t[:fwd["target_end"][i] + 1][::-1].copy()))
                live.append(i)
        kw2 = dict(pkw); kw2["z_threshold"] = -1
        rev = oracle.align_pairs(rev_pairs, op.make_params(**kw2))
        exp_q = np.zeros(len(pairs), np.int32); exp_t = np.zeros(len(pairs), np.int32)
        exp_q[live] = fwd["query_end"][live] - rev["query_end"]; exp_t[live] = fwd["target_end"][live] - rev["target_end"]
        assert (qstart == exp_q).all() and (tstart == exp_t).all(), (pkw, np.nonzero((qstart != exp_q) | (tstart != exp_t))[0][:5])
        assert (qstart >= 0).all() and (qstart <= res["query_end"]).all() and (tstart <= res["target_end"]).all()
        assert (qstart[-20:] > 20).sum() >= 15            # the junk prefixes are trimmed
