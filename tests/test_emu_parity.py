"""CPU tests of the CUDA kernel LOGIC: the product's kernel headers (agatha_b200/csrc/extend_kernel.cuh, pack_kernel.cuh)
compiled for the host by the SIMT emulation under tests/emu (one fiber per CUDA thread, collectives as rendezvous) and
compared bit-exactly with the oracle. This is test infrastructure: it proves nothing about the GPU build except that the
same source, run with CUDA semantics, gives the oracle's results -- the `-m gpu` tests are the parity tests proper."""
import numpy as np
import pytest

from oracle import oracle_py as op
from pairgen import make_pairs


@pytest.fixture(scope="module")
def emu():
    from emu import emu as e
    e.build()
    return e


def _cmp(emu, oracle, pairs, pkw, what="", s16=-1):
    got = emu.align_pairs(pairs, emu.make_params(**pkw), s16_mode=s16)
    exp = oracle.align_pairs(pairs, op.make_params(**pkw))
    for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")):
        bad = np.nonzero(got[a] != exp[b])[0]
        assert len(bad) == 0, (f"{what} {pkw} s16={s16}: {len(bad)}/{len(pairs)} pairs differ in {a}; first idx {bad[0]}: "
                               f"emu {got[bad[0]]} oracle {exp[bad[0]]} qlen {len(pairs[bad[0]][0])} tlen {len(pairs[bad[0]][1])}")


@pytest.mark.parametrize("W", [0, 1, 7, 8, 15, 33, 63, 67, 75, 99, 100, 123, 127, 255, 263, 511])
def test_emulated_kernel_vs_oracle_every_single_warp_shape(emu, oracle, W):
    # includes the band widths 3 (mod 8) that once took the static-injection prologue by mistake
    hi = 300 if W < 100 else 1400
    _cmp(emu, oracle, make_pairs(9000 + W, 50, 1, hi, mixed=True), dict(band_width=W, z_threshold=100), "rand")
    _cmp(emu, oracle, make_pairs(9100 + W, 20, max(W, 1), max(3 * W, 10), mixed=True), dict(band_width=W, slice_width=1, z_threshold=30), "rand sw1")


@pytest.mark.parametrize("s16", [-1, 0, 1, 2])
def test_emulated_default_band_all_packed_modes(emu, oracle, s16):
    # -1: the packed kernel (extend16_kernel.cuh) first, general kernel for what it marks; 0/1/2: general kernel only, with
    # its 16-bit loops off / steady state / prologue + steady state
    pairs = make_pairs(77, 10, 1600, 3500, mixed=True) + make_pairs(78, 30, 1, 1500, mixed=True)
    _cmp(emu, oracle, pairs, dict(), "w751", s16)
    _cmp(emu, oracle, pairs[:16], dict(match=2, mismatch=5, gap_open=4, gap_extend=1, z_threshold=50, band_width=759), "w759", s16)


@pytest.mark.parametrize("W,hi", [(1031, 2600), (2047, 4500), (4095, 6000)])
def test_emulated_wide_bands_multi_warp_groups(emu, oracle, W, hi):
    pairs = make_pairs(9700 + W, 6, 1, hi, mixed=True) + make_pairs(9800 + W, 2, hi, hi + 300, err=0.01)
    _cmp(emu, oracle, pairs, dict(band_width=W), "wide")
    assert emu.last_used_packed() == (W % 8 == 7)            # the packed kernel covers W = 7 (mod 8); the general kernel the rest
    _cmp(emu, oracle, pairs, dict(band_width=W), "wide, general kernel only", s16=2)
    assert not emu.last_used_packed()
    # the long clean pairs (target lengths not multiples of 8: padding-column patches in the tail, warps that skip part of the
    # prologue) must be finished by the packed kernel itself: the range monitor once mistook the stored floor of a patched
    # input for a live value on its way down and handed such pairs to the general kernel -- same results, ten times the cost
    long_pairs = [pr for pr in make_pairs(9800 + W, 2, hi, hi + 300, err=0.01)]
    long_pairs = [(q, t[:len(t) - (len(t) % 8 == 0)]) for q, t in long_pairs]
    _cmp(emu, oracle, long_pairs, dict(band_width=W), "wide, long pairs")
    assert emu.last_used_packed() and emu.last_redo_count() == 0


@pytest.mark.parametrize("sched", [1, 2, 5])
def test_emulated_pipelined_groups_under_other_fibre_schedules(emu, oracle, sched, monkeypatch):
    """The pipelined steady state of multi-warp groups (arrive / wait barriers, double-buffered hand-over slots, maxima tested
    two anti-diagonals late) must not depend on the order in which warps get their turn: backward order (1) and reshuffled
    orders with warps held back for dozens of rounds (2, 5). The warp boundary of W = 2047 lies on the main diagonal, so a
    wrong hand-over value changes scores."""
    monkeypatch.setenv("AGATHA_EMU_SCHED", str(sched))
    pairs = make_pairs(9900 + sched, 3, 2100, 3200, mixed=True) + make_pairs(9950 + sched, 1, 4200, 4400, err=0.02)
    _cmp(emu, oracle, pairs, dict(band_width=2047, z_threshold=200), "wide, schedule %d" % sched)
    assert emu.last_used_packed()


def test_emulated_edge_cases_and_rare_symbols(emu, oracle):
    pairs = [("", "ACGT"), ("ACGT", ""), ("A", "A"), ("T", "A"), ("NNNN", "NNNN"), ("ACGTACGTA", "ACG"), ("ACG", "ACGTACGTACGT"),
             ("acgtacgt", "ACGTACGT"), ("ACGTNACGT", "ACGTNACGT"), ("A" * 40, "A" * 40), ("ACGT" * 10, "TGCA" * 10)]
    for W in (7, 15, 751):
        _cmp(emu, oracle, pairs, dict(band_width=W), "edge")
    _cmp(emu, oracle, make_pairs(300, 40, 100, 800, err=0.1, n_rate=0.05), dict(band_width=63), "nrich")
    _cmp(emu, oracle, make_pairs(302, 40, 100, 800, err=0.1, iupac=True), dict(band_width=63), "iupac")
    _cmp(emu, oracle, make_pairs(303, 20, 100, 800, err=0.1), dict(band_width=63, match=200, mismatch=300), "generic scoring")


@pytest.mark.parametrize("W", [135, 255, 511, 751, 1023])
def test_emulated_packed_kernel_stops_and_hand_over(emu, oracle, W):
    """The packed kernel on every stop reason (end, Z-drop in prologue / steady state / tail, band exit at slice
    granularity, wrap-up) and its hand-over: pairs it cannot finish are marked and redone by the general kernel."""
    from pairgen import make_pair
    rng = np.random.default_rng(40 + W)
    pairs = [make_pair(rng, int(rng.integers(2 * W, 4 * W + 600)), err=0.06, skew=int(rng.integers(-2 * W, 3 * W))) for _ in range(10)]   # band exit
    pairs += [make_pair(rng, int(rng.integers(2 * W, 4 * W + 600)), err=0.08, tail=-1) for _ in range(8)]                               # junk tails
    pairs += make_pairs(50 + W, 10, W + 40, 3 * W + 500, mixed=True)
    n_long = len(pairs)
    pairs += make_pairs(60 + W, 6, 1, W, mixed=True)                     # not longer than the band: general kernel
    pairs += make_pairs(70 + W, 3, 2 * W, 3 * W, err=0.1, iupac=True)    # symbols outside {A,C,G,T,N}: general kernel
    for sw, Z in ((1, 30), (3, 400), (7, -1)):
        _cmp(emu, oracle, pairs, dict(band_width=W, slice_width=sw, z_threshold=Z), "packed")
        assert emu.last_used_packed()
        # handed over: the 9 short / IUPAC pairs for certain, plus the mixed pairs that happen to carry an N in the read
        n_query_n = sum(1 for q, _ in pairs[:n_long] if (np.asarray(q) == ord("N")).any())
        assert 6 <= emu.last_redo_count() <= len(pairs) - n_long + n_query_n, emu.last_redo_count()


def test_emulated_packed_kernel_rebase_and_range_bailout(emu, oracle):
    # long identical pairs: the score climbs past the re-centring threshold (8192) several times; match = 5 makes it fast.
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    t = acgt[rng.integers(0, 4, 9000)]
    pairs = [(t.copy(), t.copy()), (t[:7000].copy(), t[:7000].copy())]
    _cmp(emu, oracle, pairs, dict(band_width=135, match=5, mismatch=4), "rebase")
    assert emu.last_redo_count() == 0
    # the same with a multi-warp group: the lane-edge slots in shared memory have to follow the re-basing
    _cmp(emu, oracle, pairs[1:], dict(band_width=1031, match=5, mismatch=4), "rebase, 2 warps")
    assert emu.last_redo_count() == 0
    # the failure of the first drifting build: a maximum early in a long diverged pair (Z-drop practically off), its snapshot
    # must survive the many re-basings that follow
    _cmp(emu, oracle, make_pairs(74, 3, 7000, 9000, err=0.3), dict(match=2, mismatch=9, gap_open=12, gap_extend=3, band_width=511, z_threshold=30000), "old snapshot")
    assert emu.last_used_packed() and emu.last_redo_count() == 0
    # scoring the packed kernel does not take (biased table entries must fit a byte): general kernel
    _cmp(emu, oracle, pairs[:1], dict(band_width=135, match=100, mismatch=100, z_threshold=-1), "not eligible")
    assert not emu.last_used_packed()
    # Z-drop off on junk: true scores sink towards MINUS_INF2 and the packed kernel hands the pair over before the exact
    # value of the sentinel could matter
    junk = [(acgt[rng.integers(0, 4, 12000)], acgt[rng.integers(0, 4, 12000)])]
    _cmp(emu, oracle, junk, dict(band_width=135, z_threshold=-1), "sinking scores")
    assert emu.last_used_packed() and emu.last_redo_count() == 1
