"""CPU tests of the oracle (oracle/agatha_oracle.c): golden vectors produced by the reference's own kernel
code, live cross-check against oracle/_ref when it is present, and hand-checkable known answers."""
import numpy as np
import pytest

from oracle import oracle_py as op
from pairgen import make_pairs


def _triples(res):
    return np.stack([res["score"], res["query_end"], res["target_end"]], axis=1)


def test_oracle_matches_reference_golden_vectors(oracle, golden):
    total = 0
    for g in golden["groups"]:
        pairs = list(zip(g["queries"], g["targets"]))
        res = oracle.align_pairs(pairs, op.make_params(**g["params"]))
        exp = np.array(g["expected"], dtype=np.int32)
        bad = np.nonzero((_triples(res) != exp).any(axis=1))[0]
        assert len(bad) == 0, f"group {g['name']}: first mismatch at {bad[0]}: oracle {_triples(res)[bad[0]]} reference {exp[bad[0]]}"
        total += len(pairs)
    assert total >= 800


@pytest.mark.skipif(not op.RefHost.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("W,sw,Z,m,go", [(7, 1, 10, 1, 6), (15, 3, 400, 1, 6), (31, 7, 50, 2, 4), (63, 3, -1, 1, 6),
                                          (127, 3, 100, 1, 6), (751, 3, 400, 1, 6)])
def test_oracle_matches_reference_host_build_live(oracle, W, sw, Z, m, go):
    ref = op.RefHost()
    hi = 300 if W < 100 else 1500
    pairs = make_pairs(7000 + W + sw, 150 if W > 100 else 600, 1, hi, mixed=True)
    p = op.make_params(band_width=W, slice_width=sw, z_threshold=Z, match=m, gap_open=go)
    a = _triples(oracle.align_pairs(pairs, p))
    b = ref.align_pairs(pairs, p)
    assert (a == b).all()


def test_known_answers(oracle):
    p = op.make_params(band_width=7, z_threshold=400)
    # identical sequences: score = len * match, ends at the last base
    r = oracle.align_pairs([("ACGTACGTAC", "ACGTACGTAC")], p)[0]
    assert (r["score"], r["query_end"], r["target_end"]) == (10, 9, 9)
    assert r["stop"] == op.STOP_END
    # first base mismatches: the best prefix alignment is empty -> reference reports 0 at (0,0) (max starts at 0)
    r = oracle.align_pairs([("T", "A")], p)[0]
    assert (r["score"], r["query_end"], r["target_end"]) == (0, 0, 0)
    # N never matches, not even N (N_PENALTY=1)
    r = oracle.align_pairs([("NNNN", "NNNN")], p)[0]
    assert r["score"] == 0
    # one mismatch in the middle (5 + 6 - 4 = 7 beats stopping at 5)
    r = oracle.align_pairs([("ACGTAGACGTAC", "ACGTACACGTAC")], p)[0]
    assert (r["score"], r["query_end"], r["target_end"]) == (7, 11, 11)
    # lower case is the same as upper case (ascii & 15)
    r = oracle.align_pairs([("acgtacgt", "ACGTACGT")], p)[0]
    assert r["score"] == 8
    # a deletion in the query: 8 + 8 matches, gap of length 1 costs q + r = 8
    r = oracle.align_pairs([("ACGTTGCA" "GGATCCAA", "ACGTTGCA" "T" "GGATCCAA")], op.make_params(band_width=7, gap_open=6, gap_extend=2))[0]
    assert (r["score"], r["query_end"], r["target_end"]) == (8, 7, 7) or r["score"] == 8


def test_zdrop_stops_on_random_tail(oracle):
    rng = np.random.default_rng(5)
    core = rng.integers(0, 4, 300)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    q = np.concatenate([acgt[core], acgt[rng.integers(0, 4, 2000)]])
    t = np.concatenate([acgt[core], acgt[rng.integers(0, 4, 2000)]])
    r = oracle.align_pairs([(q, t)], op.make_params(band_width=63, z_threshold=100))[0]
    assert r["stop"] == op.STOP_ZDROP
    assert r["score"] >= 300 and r["query_end"] >= 299
    assert r["d_stop"] < 2 * 300 + 600
    r2 = oracle.align_pairs([(q, t)], op.make_params(band_width=63, z_threshold=-1))[0]
    assert r2["stop"] == op.STOP_END and r2["score"] >= r["score"]


def test_band_exit_when_lengths_are_skewed(oracle):
    rng = np.random.default_rng(6)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    t = acgt[rng.integers(0, 4, 200)]
    q = np.concatenate([t, acgt[rng.integers(0, 4, 1500)]])
    r = oracle.align_pairs([(q, t)], op.make_params(band_width=15, z_threshold=-1))[0]
    assert r["stop"] == op.STOP_BANDEXIT
    assert r["score"] >= 190


def test_empty_and_ragged_inputs(oracle):
    p = op.make_params(band_width=7)
    res = oracle.align_pairs([("", "ACGT"), ("ACGT", ""), ("A", "A"), ("ACGTACGTA", "ACG")], p)
    assert tuple(res[0][["score", "query_end", "target_end"]]) == (0, 0, 0)
    assert tuple(res[1][["score", "query_end", "target_end"]]) == (0, 0, 0)
    assert tuple(res[2][["score", "query_end", "target_end"]]) == (1, 0, 0)
    assert res[3]["score"] == 3


def test_cells_accounting(oracle):
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    t = acgt[rng.integers(0, 4, 700)]
    p = op.make_params(band_width=63, z_threshold=-1)
    r = oracle.align_pairs([(t, t)], p)[0]
    assert r["stop"] == op.STOP_END
    assert r["cells"] == oracle.band_cells(700, 700, 63)
    # closed form from SURVEY 8d for |qlen - tlen| <= w
    exp = sum(min(699, q + 63) - max(0, q - 63) + 1 for q in range(700))
    assert r["cells"] == exp


def test_batch_is_thread_count_invariant(oracle):
    pairs = make_pairs(77, 64, 50, 600, mixed=True)
    p = op.make_params(band_width=63)
    a = oracle.align_pairs(pairs, p, nthreads=1)
    b = oracle.align_pairs(pairs, p, nthreads=4)
    assert (a == b).all()


def test_apply_op_known_answers(oracle):
    """Reverse / complement ops on the 4-bit code (kernels/pack_rc_seqs.h:111-205; header characters test_prog.cpp:83-92)."""
    s = b"ACGTNacgtn"
    assert bytes(oracle.apply_op(s, 0)) == s
    assert bytes(oracle.apply_op(s, 1)) == s[::-1]
    comp = bytes(oracle.apply_op(s, 2))
    assert [c & 15 for c in comp] == [4, 7, 3, 1, 14, 4, 7, 3, 1, 14]          # T G C A N (by low nibble)
    assert [c & 0xF0 for c in comp] == [c & 0xF0 for c in s]                   # high nibble (case) untouched
    assert bytes(oracle.apply_op(s, 3)) == comp[::-1]
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, 1001).astype(np.uint8)
    for k in range(4):
        assert (oracle.apply_op(oracle.apply_op(x, k), k) == x).all()          # involutions
    assert len(oracle.apply_op(b"", 3)) == 0
