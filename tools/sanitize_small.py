"""A handful of pairs of one band width through agatha_extend_device, checked against the oracle -- small enough to run
under compute-sanitizer (profiles/sanitizer_r02.txt):   compute-sanitizer --tool memcheck python tools/sanitize_small.py 1031 2400"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import agatha_b200 as ag
from oracle import oracle_py as op
from pairgen import make_pairs
orc = op.Oracle()
W = int(sys.argv[1]); hi = int(sys.argv[2])
pairs = make_pairs(9700 + W, 4, W + 50, hi, mixed=True) + make_pairs(9800 + W, 2, hi, hi + 200, err=0.01)
got = ag.align_pairs_device(pairs, ag.make_params(band_width=W))
exp = orc.align_pairs(pairs, op.make_params(band_width=W))
print("W", W, "ok", all((got[a] == exp[b]).all() for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop"))))
