"""Randomized GPU campaign: random parameter sets (band 7 ... 4095, scoring, slice width, Z-drop) x random pair sets,
agatha_extend_device against the oracle on all five result fields.   python tools/campaign.py [configs]"""
import sys, time
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import agatha_b200 as ag
from oracle import oracle_py as op
from pairgen import make_pairs, make_pair
orc = op.Oracle()
t0 = time.time(); total = bad = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    rng = np.random.default_rng(50000 + seed)
    W = int(rng.choice([7, 15, 31, 63, 127, 255, 383, 511, 751, 759, 767, 1023, 250, 500, 752, 1031, 2047, 4095, 1543]))
    pkw = dict(band_width=W, slice_width=int(rng.choice([1, 3, 7])), z_threshold=int(rng.choice([-1, 50, 400, 400, 2000, 20000])),
               match=int(rng.choice([1, 1, 2, 4, 7])), mismatch=int(rng.choice([1, 3, 4, 4, 8])), gap_open=int(rng.choice([0, 2, 6, 6, 15])),
               gap_extend=int(rng.choice([1, 2, 2, 4])))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        pairs = make_pairs(60000 + seed, 30, 1, 6000, mixed=True)
    elif kind == 1:
        pairs = [make_pair(rng, int(rng.integers(3000, 16000)), err=float(rng.choice([0.0, 0.02, 0.1, 0.25]))) for _ in range(10)]
    elif kind == 2:
        pairs = [make_pair(rng, int(rng.integers(2000, 9000)), err=0.08, tail=-1) for _ in range(12)]
    else:
        pairs = [make_pair(rng, int(rng.integers(2000, 9000)), err=0.05, skew=int(rng.integers(-1500, 3000))) for _ in range(12)]
    got = ag.align_pairs_device(pairs, ag.make_params(**pkw))
    exp = orc.align_pairs(pairs, op.make_params(**pkw))
    ok = all((got[a] == exp[b]).all() for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")))
    total += len(pairs)
    if not ok:
        bad += 1
        i = int(np.nonzero((got['score'] != exp['score']) | (got['query_end'] != exp['query_end']) | (got['target_end'] != exp['target_end']) | (got['stop'] != exp['stop']) | (got['dstop'] != exp['d_stop']))[0][0])
        print('MISMATCH seed', seed, pkw, 'kind', kind, 'pair', i, 'gpu', got[i], 'oracle', exp[i], len(pairs[i][0]), len(pairs[i][1]), flush=True)
print('campaign: configs', seed + 1, 'pairs', total, 'bad configs', bad, 'sec', round(time.time() - t0, 1))
