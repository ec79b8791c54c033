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
    W = int(rng.choice([0, 3, 7, 12, 15, 23, 31, 47, 63, 67, 75, 95, 99, 123, 127, 200, 255, 383, 511, 751, 767, 1000, 1023, 1500]))
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


def test_packed_steady_state_rebasing_and_fallback(oracle):
    """The 16-bit packed steady-state loop re-centres its values every time the score has grown by 8192 and leaves to the
    32-bit loop when live values no longer fit: high match scores force many rebases, a disabled Z-drop on diverged pairs
    forces the fallback, long pairs with the default scoring stay packed all the way."""
    import agatha_b200 as ag
    cases = [
        (dict(match=5, mismatch=4, gap_open=6, gap_extend=2, band_width=751), make_pairs(71, 10, 9000, 14000, err=0.05)),        # ~50k scores: 6 rebases
        (dict(match=3, mismatch=1, gap_open=12, gap_extend=1, band_width=383, slice_width=1, z_threshold=100), make_pairs(72, 60, 1500, 4000, mixed=True)),
        (dict(match=1, mismatch=4, gap_open=6, gap_extend=2, band_width=751, z_threshold=-1), make_pairs(73, 8, 9000, 12000, err=0.1, tail=-1)),  # no Z-drop: values sink
        (dict(match=2, mismatch=9, gap_open=12, gap_extend=3, band_width=511, z_threshold=30000), make_pairs(74, 8, 7000, 9000, err=0.3)),
        (dict(band_width=751), make_pairs(75, 6, 28000, 31000, err=0.02)),                                                      # scores near 30k
        (dict(match=100, mismatch=100, gap_open=1000, gap_extend=500, band_width=255), make_pairs(76, 20, 1500, 3000, err=0.1)),
    ]
    for pkw, pairs in cases:
        got = ag.align_pairs_device(pairs, ag.make_params(**pkw))
        exp = oracle.align_pairs(pairs, op.make_params(**pkw))
        for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
            bad = np.nonzero(got[a] != exp[b])[0]
            assert len(bad) == 0, f"{pkw}: {a} differs for {len(bad)} pairs, first {bad[0]}: gpu {got[bad[0]]} oracle {exp[bad[0]]}"


def test_off_band_repeat_does_not_leak_into_the_band(oracle):
    """Dead cells just outside the band see a perfect repeat for tens of thousands of anti-diagonals (they gain `match`
    per matching pair while nothing else feeds them); in the 16-bit packed loop they sit only 30000 below live values, so
    they must be pushed back regularly. Target = unit repeated; query = the same shifted by one period > band width."""
    import agatha_b200 as ag
    rng = np.random.default_rng(99)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    pairs = []
    for period, W in ((800, 751), (300, 255), (1100, 1023)):
        unit = acgt[rng.integers(0, 4, period)]
        t = np.tile(unit, 40000 // period + 2)[:40000]
        noise = acgt[rng.integers(0, 4, 40000)]
        q = t.copy()
        bad = rng.random(40000) < 0.30                     # the in-band alignment is poor but alive ...
        q[bad] = noise[bad]
        q = np.concatenate([q[period:], unit])             # ... while diagonal offset = period matches perfectly
        pairs.append((W, q, t))
    for W, q, t in pairs:
        for pkw in (dict(band_width=W, z_threshold=-1), dict(band_width=W, z_threshold=20000, match=2, mismatch=2)):
            got = ag.align_pairs_device([(q, t)], ag.make_params(**pkw))
            exp = oracle.align_pairs([(q, t)], op.make_params(**pkw))
            for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
                assert got[a][0] == exp[b][0], (pkw, a, got, exp)


@pytest.mark.parametrize("w", [63, 255, 751, 1023, 4095])
def test_zdrop_and_band_shifts_inside_the_prologue(oracle, w, monkeypatch):
    """The first W+1 anti-diagonals run in aligned blocks of 8 with the matrix-edge cells injected at compile-time positions
    (packed 16-bit state for one-warp shapes, 32-bit for the wide ones). Z-drop must be able to fire on every place of a
    block, a low anti-diagonal that does NOT fire (the best cell moved off the diagonal, so l*ge raises the bar) must leave
    the block and come back, and all of it must match the oracle and the plain 32-bit path."""
    import agatha_b200 as ag
    rng = np.random.default_rng(w)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = lambda n: acgt[rng.integers(0, 4, int(n))]
    L = 3 * w + 600
    pairs = []
    for cut in list(range(0, 40)) + [int(x) for x in rng.integers(40, max(41, w // 2 + 20), 60)]:
        t = seq(L)
        pairs.append((np.concatenate([t[:cut], seq(L - cut)]), t))                 # junk after `cut` matching bases
    for _ in range(40):                                                            # a long gap early: the best cell leaves the diagonal
        t = seq(L)
        a, gap = int(rng.integers(5, max(6, w // 3))), int(rng.integers(10, max(11, w // 2)))
        q = np.concatenate([t[:a], t[a + gap:]]) if rng.random() < 0.5 else np.concatenate([t[:a], seq(gap), t[a:]])
        b = int(rng.integers(a + 5, a + 5 + w // 2))
        pairs.append((np.concatenate([q[:b], seq(max(L - b, 1))]), t))
    for pkw in (dict(z_threshold=20), dict(z_threshold=100, gap_extend=2), dict(z_threshold=400), dict(z_threshold=60, match=2, mismatch=6, gap_open=3, gap_extend=1),
                dict(z_threshold=-1)):
        pkw = dict(pkw, band_width=w)
        exp = oracle.align_pairs(pairs, op.make_params(**pkw))
        res = {}
        for mode in ("2", "7", "1", "0"):                                          # default (packed prologue + steady state) / + packed tail / steady state only / 32-bit
            monkeypatch.setenv("AGATHA_S16", mode)
            res[mode] = ag.align_pairs_device(pairs, ag.make_params(**pkw))
        for mode, got in res.items():
            for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
                bad = np.nonzero(got[a] != exp[b])[0]
                assert len(bad) == 0, f"AGATHA_S16={mode} {pkw}: {len(bad)} pairs differ in {a}, first {bad[0]}: gpu {got[bad[0]]} oracle {exp[bad[0]]}"
        if w <= 1023:
            ds = exp["d_stop"][exp["stop"] == 1]
            if pkw["z_threshold"] == 20:
                assert len(set(int(x) % 8 for x in ds if x <= w)) == 8             # fired on every place of a block


@pytest.mark.parametrize("seed", range(4))
def test_packed_tail_opt_in_matches_oracle(oracle, seed, monkeypatch):
    """AGATHA_S16=7 also runs the far-edge part of the alignment on packed state (valid-cell masks, padding-column patches
    and the slice-wise band-exit check inside the packed loop) in builds made with -DAGATHA_TAIL16=1; the default build does
    not contain that loop (it costs the steady state 4 %) and ignores the switch. Either way the results must be bit-exact."""
    import agatha_b200 as ag
    from pairgen import make_pair
    monkeypatch.setenv("AGATHA_S16", "7")
    rng = np.random.default_rng(7000 + seed)
    W = int(rng.choice([127, 255, 511, 751, 1023]))
    pkw = dict(band_width=W, slice_width=int(rng.choice([1, 3, 7])), z_threshold=int(rng.choice([-1, 100, 400, 5000])),
               match=int(rng.choice([1, 2])), mismatch=int(rng.choice([2, 4])), gap_open=int(rng.choice([2, 6])), gap_extend=int(rng.choice([1, 2])))
    pairs = make_pairs(7100 + seed, 120, 2 * W, 6 * W + 2000, mixed=True)
    pairs += [make_pair(rng, int(rng.integers(3 * W, 5 * W + 1000)), err=0.06, skew=int(rng.integers(-2 * W, 3 * W))) for _ in range(40)]   # band exit
    pairs += [make_pair(rng, int(rng.integers(3 * W, 5 * W + 1000)), err=0.06, tail=-1) for _ in range(20)]                               # junk tails
    got = ag.align_pairs_device(pairs, ag.make_params(**pkw))
    exp = oracle.align_pairs(pairs, op.make_params(**pkw))
    for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
        bad = np.nonzero(got[a] != exp[b])[0]
        assert len(bad) == 0, f"{pkw}: {len(bad)} pairs differ in {a}, first {bad[0]}: gpu {got[bad[0]]} oracle {exp[bad[0]]} lens {len(pairs[bad[0]][0])},{len(pairs[bad[0]][1])}"
    assert len(set(exp["stop"])) >= 2
