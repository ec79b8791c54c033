"""Per-sequence reverse / complement ops: gasal_op_fill + params->isReverseComplement in the
reference (interfaces.cpp:69-84, gasal_align.cu:199-212, kernels/pack_rc_seqs.h:56-212). The engine applies the op to the
real bases; the oracle does the same on the host before aligning. The reference binary is compared where its own kernel
is well defined: the complement bit. Its reverse is broken as compiled (-DN_CODE=0x4E never equals a 4-bit code, so the
padding count is 0 and the word-straddling shifts become shifts by 32; measured on a B200: 0/60 pairs reverse correctly
even when every length is a multiple of 8) -- see oracle/README.md; reverse is pinned by the oracle alone."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as op
import pairgen

pytestmark = pytest.mark.gpu

REF_RC = os.path.join(os.path.dirname(op.REF_GPU_BIN), "agatha_ref_manual_rc")
DROPIN_RC = os.path.join(os.path.dirname(op.REF_GPU_BIN), "agatha_dropin_manual_rc")
FLAGS = ["-m", "1", "-x", "4", "-q", "6", "-r", "2", "-s", "3", "-z", "400", "-w", "751"]
KEYS = ("score", "query_end", "target_end")


def _pairs(rng, n, lo, hi):
    out = []
    for _ in range(n):
        t = pairgen.random_seq(rng, int(rng.integers(lo, hi)))
        out.append((pairgen.mutate_fast(rng, t, 0.05, 0.02, 0.02), t))
    return out


def _expected(oracle, pairs, qops, tops, params):
    """What the ops mean: apply them on the host, then align. A reverse-complemented query still has to match, so the
    generator below transforms the inputs with the inverse op first and lets the engine undo it."""
    done = [(oracle.apply_op(q, a), oracle.apply_op(t, b)) for (q, t), a, b in zip(pairs, qops, tops)]
    return oracle.align_pairs(done, params)


def _pre_transform(oracle, pairs, qops, tops):
    # reverse and complement are involutions and commute, so op(op(x)) == x: feeding op(x) makes the engine align x
    return [(oracle.apply_op(q, a), oracle.apply_op(t, b)) for (q, t), a, b in zip(pairs, qops, tops)]


@pytest.mark.parametrize("w", [31, 751])
def test_job_api_ops_match_oracle(oracle, w):
    import agatha_b200 as ag
    rng = np.random.default_rng(100 + w)
    pairs = _pairs(rng, 300, 20, 1500)
    pairs += [(pairgen.random_seq(rng, k), pairgen.random_seq(rng, k + 1)) for k in range(1, 20)]     # tiny, ragged
    n = len(pairs)
    qops = rng.integers(0, 4, n).astype(np.uint8); tops = rng.integers(0, 4, n).astype(np.uint8)
    fed = _pre_transform(oracle, pairs, qops, tops)
    params = dict(match=1, mismatch=4, gap_open=6, gap_extend=2, slice_width=3, z_threshold=400, band_width=w)
    exp = oracle.align_pairs(pairs, op.make_params(**params))                    # == align(op(fed))
    chk = _expected(oracle, fed, qops, tops, op.make_params(**params))
    for k in KEYS:
        assert (exp[k] == chk[k]).all()
    got = ag.align_pairs(fed, params, query_ops=qops, target_ops=tops, batch_alns=128)
    for k in KEYS + ("stop",):
        assert (got[k] == exp[k]).all(), k
    # without ops the same inputs give something else (the ops are really applied)
    plain = ag.align_pairs(fed, params)
    assert (plain["score"] != exp["score"]).any()
    # all-zero ops == no ops
    zero = ag.align_pairs(fed, params, query_ops=np.zeros(n, np.uint8), target_ops=np.zeros(n, np.uint8))
    for k in KEYS:
        assert (zero[k] == plain[k]).all()


def test_device_api_and_stream_ops_match_oracle(oracle):
    import agatha_b200 as ag
    rng = np.random.default_rng(7)
    pairs = _pairs(rng, 200, 30, 900)
    n = len(pairs)
    qops = rng.integers(0, 4, n).astype(np.uint8); tops = rng.integers(0, 4, n).astype(np.uint8)
    params = dict(match=2, mismatch=4, gap_open=4, gap_extend=2, slice_width=3, z_threshold=200, band_width=127)
    exp = _expected(oracle, pairs, qops, tops, op.make_params(**params))
    got = ag.align_pairs_device(pairs, params, ops=(qops, tops))
    for k in KEYS + ("stop",):
        assert (got[k] == exp[k]).all(), k
    s = ag.Stream(0, max_alns=n)
    s.fill(*ag.stage_pairs(pairs))
    s.set_ops(qops, tops)
    s.submit(params, ops=True); s.wait()
    got = s.results()
    for k in KEYS + ("stop",):
        assert (got[k] == exp[k]).all(), k
    s.submit(params); s.wait()                      # the op bytes stay staged but are ignored by the plain submit
    plain = oracle.align_pairs(pairs, op.make_params(**params))
    assert (s.results()["score"] == plain["score"]).all()
    s.close()


def _run(binary, qf, tf, workdir, tag, extra=()):
    raw = os.path.join(workdir, "raw_%s.log" % tag)
    score = os.path.join(workdir, "score_%s.log" % tag)
    with open(score, "w") as so:
        r = subprocess.run([binary, "-p"] + FLAGS + list(extra) + [qf, tf, raw], stdout=so, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return open(score).read()


def _lines(res):
    return "".join("%d\tquery_batch_end=%d\ttarget_batch_end=%d\n" % (r["score"], r["query_end"], r["target_end"]) for r in res)


@pytest.mark.skipif(not (os.path.exists(REF_RC) and os.path.exists(DROPIN_RC)), reason="oracle/_ref RC binaries not built")
def test_reference_driver_with_ops_enabled_prints_identical_scores(oracle, tmp_path):
    """The reference's driver (one-line edit: isReverseComplement = true) linked with the reference library and with this
    repository's library, on FASTA files whose header characters request the complement op ('/') on either side."""
    import agatha_b200 as ag
    rng = np.random.default_rng(11)
    pairs = _pairs(rng, 400, 64, 3000)
    n = len(pairs)
    qops = (2 * rng.integers(0, 2, n)).astype(np.uint8); tops = (2 * rng.integers(0, 2, n)).astype(np.uint8)
    fed = _pre_transform(oracle, pairs, qops, tops)
    qbuf, qoff, qlen, tbuf, toff, tlen = op.concat_pairs(fed)
    qf, tf = str(tmp_path / "q.fasta"), str(tmp_path / "t.fasta")
    ag.write_fasta(qf, qbuf, qoff, qlen, ops=qops)
    ag.write_fasta(tf, tbuf, toff, tlen, ops=tops)
    exp = _lines(oracle.align_pairs(pairs, op.make_params(match=1, mismatch=4, gap_open=6, gap_extend=2, slice_width=3, z_threshold=400, band_width=751)))
    ref = _run(REF_RC, qf, tf, str(tmp_path), "ref")
    new = _run(DROPIN_RC, qf, tf, str(tmp_path), "new")
    assert new == exp
    assert ref == exp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    drv = _run(os.path.join(root, "agatha_b200", "bin", "agatha_manual"), qf, tf, str(tmp_path), "drv", extra=("-R",))
    assert drv == exp


@pytest.mark.skipif(not os.path.exists(DROPIN_RC), reason="oracle/_ref/agatha_dropin_manual_rc not built")
def test_dropin_driver_ops_on_ragged_lengths_match_oracle(oracle, tmp_path):
    import agatha_b200 as ag
    rng = np.random.default_rng(12)
    pairs = _pairs(rng, 300, 9, 1200)
    n = len(pairs)
    qops = rng.integers(0, 4, n).astype(np.uint8); tops = rng.integers(0, 4, n).astype(np.uint8)
    fed = _pre_transform(oracle, pairs, qops, tops)
    qbuf, qoff, qlen, tbuf, toff, tlen = op.concat_pairs(fed)
    qf, tf = str(tmp_path / "q.fasta"), str(tmp_path / "t.fasta")
    ag.write_fasta(qf, qbuf, qoff, qlen, ops=qops)
    ag.write_fasta(tf, tbuf, toff, tlen, ops=tops)
    exp = _lines(oracle.align_pairs(pairs, op.make_params(match=1, mismatch=4, gap_open=6, gap_extend=2, slice_width=3, z_threshold=400, band_width=751)))
    assert _run(DROPIN_RC, qf, tf, str(tmp_path), "new") == exp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert _run(os.path.join(root, "agatha_b200", "bin", "agatha_manual"), qf, tf, str(tmp_path), "drv", extra=("-R",)) == exp
