"""GPU parity tests: the CUDA path (through the C ABI of libagatha_b200.so) against
 (1) golden vectors produced by the reference's own kernel code (tests/golden/ref_vectors.json.gz),
 (2) the CPU oracle on seeded inputs, bit-exact on score, end coordinates, stop reason and stop diagonal."""
import numpy as np
import pytest

from oracle import oracle_py as op
from pairgen import make_pairs

pytestmark = pytest.mark.gpu


def _engine():
    import agatha_b200
    return agatha_b200


def _cmp_oracle(oracle, pairs, pkw, what=""):
    ag = _engine()
    got = ag.align_pairs_device(pairs, ag._lib.make_params(**pkw))
    exp = oracle.align_pairs(pairs, op.make_params(**pkw))
    for k_got, k_exp in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
        bad = np.nonzero(got[k_got] != exp[k_exp])[0]
        assert len(bad) == 0, (f"{what} {pkw}: {len(bad)}/{len(pairs)} pairs differ in {k_got}; first idx {bad[0]}: "
                               f"gpu {got[bad[0]]} oracle {exp[bad[0]]} qlen {len(pairs[bad[0]][0])} tlen {len(pairs[bad[0]][1])}")


def test_golden_vectors_from_reference_kernel(golden):
    ag = _engine()
    for g in golden["groups"]:
        pairs = list(zip(g["queries"], g["targets"]))
        got = ag.align_pairs_device(pairs, ag._lib.make_params(**g["params"]))
        exp = np.array(g["expected"], dtype=np.int32)
        tri = np.stack([got["score"], got["query_end"], got["target_end"]], axis=1)
        bad = np.nonzero((tri != exp).any(axis=1))[0]
        assert len(bad) == 0, f"group {g['name']}: {len(bad)} mismatches, first {bad[0]}: gpu {tri[bad[0]]} reference {exp[bad[0]]}"


@pytest.mark.parametrize("W,sw,Z,m,go", [(7, 1, 10, 1, 6), (15, 3, 400, 1, 6), (31, 7, 50, 2, 4), (63, 3, -1, 1, 6),
                                          (127, 3, 100, 1, 6), (255, 3, 400, 1, 6), (263, 3, 400, 1, 6),
                                          (511, 3, 200, 1, 6), (751, 3, 400, 1, 6), (759, 3, 400, 2, 4), (1023, 3, 400, 1, 6)])
def test_random_pairs_vs_oracle(oracle, W, sw, Z, m, go):
    hi = 300 if W < 100 else 2500
    pairs = make_pairs(9000 + W + sw, 200 if W > 100 else 500, 1, hi, mixed=True)
    _cmp_oracle(oracle, pairs, dict(band_width=W, slice_width=sw, z_threshold=Z, match=m, gap_open=go), "random")


@pytest.mark.parametrize("W,hi", [(1031, 3000), (2047, 5000), (2055, 5000), (4095, 7000), (4103, 9000), (8191, 9000)])
def test_wide_bands_multi_warp_groups(oracle, W, hi):
    # bands wider than one warp's registers: NW = 2, 4, 8 warps per alignment (pipelined steady state over arrive / wait barriers)
    pairs = make_pairs(9700 + W, 40, 1, hi, mixed=True) + make_pairs(9800 + W, 6, hi, hi + 500, err=0.01)
    _cmp_oracle(oracle, pairs, dict(band_width=W), "wide")
    _cmp_oracle(oracle, pairs[:24], dict(band_width=W, z_threshold=60, slice_width=1), "wide z60")


def test_hifi_like_wide_band_profile(oracle):
    # C3: HiFi-like 15 kb pairs, -w 4095
    import agatha_b200 as ag
    d = ag.synth_pairs(3, 3, 24)
    res, _ = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], ag.make_params(band_width=4095))
    exp = oracle.align_batch(d["qbuf"], d["qoff"].astype(np.uint32), d["qlen"], d["tbuf"], d["toff"].astype(np.uint32), d["tlen"], op.make_params(band_width=4095))
    for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
        assert (res[a] == exp[b]).all(), a


def test_hifi_like_wide_band_profile_512_pairs(oracle):
    # C3 at full pair length through the multi-warp packed kernel (4 warps per alignment), 512 pairs against the oracle on all
    # five result fields. Every batch is two launches: the packed kernel and the general kernel's redo pass -- no pack kernel,
    # the job API packs on the host (and starts with small batches, so there are several).
    import agatha_b200 as ag
    d = ag.synth_pairs(3, 3, 512)
    n0 = ag.launch_count()
    res, stats = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], ag.make_params(band_width=4095), batch_alns=512)
    assert stats["n_batches"] >= 1 and ag.launch_count() - n0 == 2 * stats["n_batches"]
    exp = oracle.align_batch(d["qbuf"], d["qoff"].astype(np.uint32), d["qlen"], d["tbuf"], d["toff"].astype(np.uint32), d["tlen"], op.make_params(band_width=4095))
    for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
        bad = np.nonzero(res[a] != exp[b])[0]
        assert len(bad) == 0, (a, len(bad), res[bad[0]], exp[bad[0]])


def test_even_and_odd_unaligned_band_widths(oracle):
    # strict band semantics for widths the reference does not define exactly: GPU == oracle
    for W in (0, 1, 8, 10, 33, 100, 750):
        pairs = make_pairs(9500 + W, 150, 1, 900, mixed=True)
        _cmp_oracle(oracle, pairs, dict(band_width=W, z_threshold=100), "unaligned W")
    # W = 3 (mod 8) with C = 4 cells per lane: must NOT take the compile-time-position prologue (engine dispatch)
    for W in (67, 75, 83, 91, 99, 107, 115, 123):
        pairs = make_pairs(9600 + W, 120, W + 1, 4 * W, mixed=True)
        _cmp_oracle(oracle, pairs, dict(band_width=W, z_threshold=100), "W = 3 mod 8")


def test_edge_cases(oracle):
    pairs = [("", "ACGT"), ("ACGT", ""), ("A", "A"), ("T", "A"), ("NNNN", "NNNN"), ("ACGTACGTA", "ACG"), ("ACG", "ACGTACGTACGT"),
             ("acgtacgt", "ACGTACGT"), ("ACGTNACGT", "ACGTNACGT"), ("A" * 40, "A" * 40), ("ACGT" * 10, "TGCA" * 10)]
    for W in (7, 15, 751):
        _cmp_oracle(oracle, pairs, dict(band_width=W), "edge")


def test_n_rich_lowercase_iupac(oracle):
    _cmp_oracle(oracle, make_pairs(300, 200, 100, 800, err=0.1, n_rate=0.05), dict(band_width=63), "nrich")
    _cmp_oracle(oracle, make_pairs(301, 200, 100, 800, err=0.1, lower=True), dict(band_width=63), "lower")
    _cmp_oracle(oracle, make_pairs(302, 200, 100, 800, err=0.1, iupac=True), dict(band_width=63), "iupac")


def test_long_pairs_default_scoring(oracle):
    # ONT-like lengths, default AGAThA.sh scoring; includes Z-drop tails
    rng_pairs = make_pairs(4242, 24, 6000, 12000, err=0.1) + make_pairs(4243, 8, 6000, 9000, err=0.1, tail=-1)
    _cmp_oracle(oracle, rng_pairs, dict(), "long")


def test_order_does_not_change_results(oracle):
    ag = _engine()
    pairs = make_pairs(555, 300, 10, 1500, mixed=True)
    p = ag._lib.make_params(band_width=127)
    a = ag.align_pairs_device(pairs, p, bucket=True)
    b = ag.align_pairs_device(pairs, p, bucket=False)
    assert (a == b).all()
