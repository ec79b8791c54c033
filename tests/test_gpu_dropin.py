"""Drop-in proof on the GPU box: the reference's own driver program, once linked with the reference library
(oracle/_ref/agatha_ref_manual, built for sm_100) and once -- same unmodified test_prog.cpp -- linked with this
repository's shim + libagatha_b200.so (oracle/_ref/agatha_dropin_manual), must print byte-identical score logs."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as op

pytestmark = pytest.mark.gpu

REF = op.REF_GPU_BIN
DROPIN = os.path.join(os.path.dirname(op.REF_GPU_BIN), "agatha_dropin_manual")
FLAGS = ["-m", "1", "-x", "4", "-q", "6", "-r", "2", "-s", "3", "-z", "400", "-w", "751"]   # AGAThA.sh:44


def _run(binary, qf, tf, workdir, tag, extra=(), flags=None):
    raw = os.path.join(workdir, "raw_%s.log" % tag)
    score = os.path.join(workdir, "score_%s.log" % tag)
    with open(score, "w") as so:
        r = subprocess.run([binary, "-p"] + list(flags or FLAGS) + list(extra) + [qf, tf, raw], stdout=so, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return open(score).read(), [float(x) for x in open(raw).read().split()]


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DROPIN)), reason="oracle/_ref binaries not built (need /root/reference at build time)")
def test_reference_driver_prints_identical_scores_with_either_library(tmp_path):
    import agatha_b200 as ag
    d = ag.synth_pairs(1, 1, 8192 + 700)           # C1 stand-in for the bundled dataset: two batches of the default 8192
    qf, tf = str(tmp_path / "query.fasta"), str(tmp_path / "ref.fasta")
    ag.write_fasta(qf, d["qbuf"], d["qoff"], d["qlen"])
    ag.write_fasta(tf, d["tbuf"], d["toff"], d["tlen"])
    ref_scores, ref_ms = _run(REF, qf, tf, str(tmp_path), "ref")
    new_scores, new_ms = _run(DROPIN, qf, tf, str(tmp_path), "new")
    assert len(ref_scores.splitlines()) == len(d["qlen"])
    assert new_scores == ref_scores
    assert len(new_ms) == len(ref_ms) == 2          # one raw.log line per batch, like the reference (gasal_align.cu:219-236)
    # and the same through the C ABI job call
    res, _ = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], ag.make_params())
    lines = ["%d\tquery_batch_end=%d\ttarget_batch_end=%d" % (r["score"], r["query_end"], r["target_end"]) for r in res]
    assert "\n".join(lines) + "\n" == ref_scores


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/agatha_ref_manual not built")
def test_reference_gpu_kernel_agrees_on_long_ont_like_pairs(tmp_path):
    import agatha_b200 as ag
    d = ag.synth_pairs(2, 2, 3000)                  # C2 lengths, inside the reference's int16 domain
    qf, tf = str(tmp_path / "q.fasta"), str(tmp_path / "t.fasta")
    ag.write_fasta(qf, d["qbuf"], d["qoff"], d["qlen"])
    ag.write_fasta(tf, d["tbuf"], d["toff"], d["tlen"])
    ref_scores, _ = _run(REF, qf, tf, str(tmp_path), "ref")
    res, _ = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], ag.make_params())
    got = np.array([[r["score"], r["query_end"], r["target_end"]] for r in res])
    exp = np.array([[int(a), int(b.split("=")[1]), int(c.split("=")[1])] for a, b, c in (ln.split("\t") for ln in ref_scores.splitlines())])
    assert (got == exp).all()


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/agatha_ref_manual not built")
def test_native_driver_and_runner_script_match_reference_outputs(tmp_path):
    """agatha_manual (this repo's own driver, reference CLI contract) and tools/agatha.sh (AGAThA.sh-compatible runner:
    raw.log, score.log, time.json) against the reference binary on the same FASTA files."""
    import json
    import agatha_b200 as ag
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    driver = os.path.join(root, "agatha_b200", "bin", "agatha_manual")
    assert os.path.exists(driver), "run python -m agatha_b200.build"
    d = ag.synth_pairs(1, 5, 3000)
    ds = tmp_path / "dataset"; ds.mkdir()
    # AGAThA.sh:44 passes ref.fasta first (query batch role) and query.fasta second (target batch role)
    ag.write_fasta(str(ds / "ref.fasta"), d["qbuf"], d["qoff"], d["qlen"])
    ag.write_fasta(str(ds / "query.fasta"), d["tbuf"], d["toff"], d["tlen"])
    ref_scores, _ = _run(REF, str(ds / "ref.fasta"), str(ds / "query.fasta"), str(tmp_path), "ref")
    new_scores, new_ms = _run(driver, str(ds / "ref.fasta"), str(ds / "query.fasta"), str(tmp_path), "drv")
    assert new_scores == ref_scores and len(new_ms) == 1
    out = tmp_path / "output"
    r = subprocess.run(["bash", os.path.join(root, "tools", "agatha.sh"), "-i", "2", "-d", str(ds), "-o", str(out)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(out / "score.log").read() == ref_scores
    assert len(open(out / "raw.log").read().split()) == 2
    tj = json.load(open(out / "time.json"))
    assert tj["AGAThA"]["test"] > 0


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DROPIN)), reason="oracle/_ref binaries not built")
def test_reference_driver_with_several_host_threads(tmp_path):
    """-n 3: three OpenMP threads of the reference driver, each with its own pair of storages/streams (test_prog.cpp:210-245),
    against the library's process-global state (scoring, launch configuration, stream cache). Batches finish in any order, so
    the multiset of lines is compared."""
    import agatha_b200 as ag
    d = ag.synth_pairs(1, 21, 5000)
    qf, tf = str(tmp_path / "q.fasta"), str(tmp_path / "t.fasta")
    ag.write_fasta(qf, d["qbuf"], d["qoff"], d["qlen"])
    ag.write_fasta(tf, d["tbuf"], d["toff"], d["tlen"])
    ref_scores, _ = _run(REF, qf, tf, str(tmp_path), "ref1", extra=("-a", "512"))
    new_scores, new_ms = _run(DROPIN, qf, tf, str(tmp_path), "new3", extra=("-n", "3", "-a", "512"))
    assert sorted(new_scores.splitlines()) == sorted(ref_scores.splitlines())
    assert len(new_ms) == 12          # 3 threads x ceil(1667/512) batches: one whole raw.log line per batch (the shim locks the file)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/agatha_ref_manual not built")
@pytest.mark.parametrize("name,profile,seed,band", [("C3 HiFi-like, wide band", 3, 3, 4095), ("C4 heavy tail, early Z-drop", 4, 4, 751), ("C5 ONT-like", 2, 5, 751)])
def test_reference_gpu_kernel_agrees_on_2048_pair_slices_of_every_long_workload(tmp_path, name, profile, seed, band):
    """Full-size pairs of BASELINE configs 3, 4 and 5 against the UNMODIFIED reference GPU program on the same B200:
    score and both end coordinates of every pair inside the reference's valid domain (lengths < 32768 for
    its rejoin path, scores <= 32767) must be identical."""
    import agatha_b200 as ag
    d = ag.synth_pairs(profile, seed, 2600)
    ok = np.nonzero((d["qlen"] < 32768) & (d["tlen"] < 32768))[0][:2048]
    assert len(ok) == 2048
    sub_q = [d["qbuf"][int(d["qoff"][i]):int(d["qoff"][i]) + int(d["qlen"][i])] for i in ok]
    sub_t = [d["tbuf"][int(d["toff"][i]):int(d["toff"][i]) + int(d["tlen"][i])] for i in ok]
    ql = d["qlen"][ok]; tl = d["tlen"][ok]
    qo = np.concatenate([[0], np.cumsum(ql[:-1], dtype=np.uint64)]).astype(np.uint64)
    to = np.concatenate([[0], np.cumsum(tl[:-1], dtype=np.uint64)]).astype(np.uint64)
    qb, tb = np.concatenate(sub_q), np.concatenate(sub_t)
    qf, tf = str(tmp_path / "q.fasta"), str(tmp_path / "t.fasta")
    ag.write_fasta(qf, qb, qo, ql)
    ag.write_fasta(tf, tb, to, tl)
    flags = FLAGS[:-1] + [str(band)]
    ref_scores, _ = _run(REF, qf, tf, str(tmp_path), "ref", flags=flags)
    res, _ = ag.align_job(qb, qo, ql, tb, to, tl, ag.make_params(band_width=band))
    got = np.array([[r["score"], r["query_end"], r["target_end"]] for r in res])
    exp = np.array([[int(a), int(b.split("=")[1]), int(c.split("=")[1])] for a, b, c in (ln.split("\t") for ln in ref_scores.splitlines())])
    inside = got[:, 0] <= 32767                      # the reference's scores wrap beyond int16
    assert inside.sum() >= 2000, name
    bad = np.nonzero((got[inside] != exp[inside]).any(axis=1))[0]
    assert len(bad) == 0, "%s: %d of %d pairs differ, first: ours %s reference %s" % (name, len(bad), inside.sum(), got[inside][bad[0]], exp[inside][bad[0]])
