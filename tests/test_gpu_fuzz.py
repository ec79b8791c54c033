"""Randomised GPU-vs-oracle fuzzing over scoring, band, slice and Z-drop parameters (seeded, a few seconds)."""
import numpy as np
import pytest

from oracle import oracle_py as op
from pairgen import make_pairs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_parameters(oracle, seed):
    import agatha_b200 as ag
    rng = np.random.default_rng(1000 + seed)
    W = int(rng.choice([0, 3, 7, 12, 15, 23, 31, 47, 63, 95, 127, 200, 255, 383, 511, 751, 767, 1000, 1023, 1500]))
    pkw = dict(band_width=W, slice_width=int(rng.choice([1, 2, 3, 4, 7, 8, 15])), z_threshold=int(rng.choice([-1, 0, 1, 5, 30, 100, 400, 5000])),
               match=int(rng.choice([1, 2, 3, 5])), mismatch=int(rng.choice([1, 2, 4, 6, 9])), gap_open=int(rng.choice([0, 1, 4, 6, 12])),
               gap_extend=int(rng.choice([1, 2, 3])))
    hi = int(rng.choice([40, 300, 1500, 4000]))
    pairs = make_pairs(2000 + seed, 160 if hi > 1000 else 400, 1, hi, mixed=True)
    got = ag.align_pairs_device(pairs, ag.make_params(**pkw))
    exp = oracle.align_pairs(pairs, op.make_params(**pkw))
    for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
        bad = np.nonzero(got[a] != exp[b])[0]
        assert len(bad) == 0, f"{pkw}: {len(bad)} pairs differ in {a}, first {bad[0]}: gpu {got[bad[0]]} oracle {exp[bad[0]]} lens {len(pairs[bad[0]][0])},{len(pairs[bad[0]][1])}"


def test_extreme_shapes(oracle):
    import agatha_b200 as ag
    rng = np.random.default_rng(7)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = lambda n: acgt[rng.integers(0, 4, n)]
    t = seq(20000)
    pairs = [(seq(1), t), (t, seq(1)), (t[:7], t), (t, t[:9]), (t, t), (t[:15000], t[40:15040]), (seq(8), seq(8)), (seq(33000), seq(33000)),
             (np.concatenate([t[:5000], seq(3000)]), t[:9000])]
    for pkw in (dict(), dict(band_width=63, z_threshold=50), dict(band_width=1023, slice_width=7), dict(band_width=2047, z_threshold=-1)):
        got = ag.align_pairs_device(pairs, ag.make_params(**pkw))
        exp = oracle.align_pairs(pairs, op.make_params(**pkw))
        for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
            assert (got[a] == exp[b]).all(), (pkw, a, got, exp)


def test_scores_outside_the_byte_table_use_generic_scoring(oracle):
    import agatha_b200 as ag
    pairs = make_pairs(4321, 120, 5, 700, mixed=True)
    for pkw in (dict(match=200, mismatch=300, gap_open=500, gap_extend=150, band_width=63, z_threshold=20000),
                dict(match=1, mismatch=0, band_width=31), dict(match=0, mismatch=1, band_width=31, z_threshold=5)):
        got = ag.align_pairs_device(pairs, ag.make_params(**pkw))
        exp = oracle.align_pairs(pairs, op.make_params(**pkw))
        for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
            assert (got[a] == exp[b]).all(), (pkw, a)
