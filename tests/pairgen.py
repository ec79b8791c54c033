"""Seeded random read/reference pair generator for the parity tests (small sizes; numpy only)."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_seq(rng, n, alphabet=ACGT):
    return alphabet[rng.integers(0, len(alphabet), size=int(n))]


def mutate(rng, t, sub=0.04, ins=0.03, dele=0.03):
    """Query = target with i.i.d. per-base substitutions / insertions / deletions."""
    n = len(t)
    u = rng.random(n)
    keep = u >= dele
    out = []
    subm = (u >= dele) & (u < dele + sub)
    t2 = t.copy()
    if subm.any():
        t2[subm] = ACGT[(np.searchsorted(ACGT, np.clip(t2[subm], 65, 84)) + rng.integers(1, 4, size=int(subm.sum()))) % 4]
    insm = rng.random(n) < ins
    # build with insertions after kept bases
    for i in range(n):
        if keep[i]:
            out.append(t2[i])
        if insm[i]:
            out.append(ACGT[rng.integers(0, 4)])
    return np.array(out, dtype=np.uint8) if out else np.zeros(0, np.uint8)


def mutate_fast(rng, t, sub=0.04, ins=0.03, dele=0.03):
    """Vectorised version of mutate() for longer sequences."""
    n = len(t)
    u = rng.random(n)
    t2 = t.copy()
    subm = (u >= dele) & (u < dele + sub)
    k = int(subm.sum())
    if k:
        idx = np.searchsorted(ACGT, t2[subm])
        idx = np.where(ACGT[np.clip(idx, 0, 3)] == t2[subm], idx, 0)
        t2[subm] = ACGT[(idx + rng.integers(1, 4, size=k)) % 4]
    keep = u >= dele
    insm = rng.random(n) < ins
    counts = keep.astype(np.int64) + insm.astype(np.int64)
    total = int(counts.sum())
    out = np.empty(total, dtype=np.uint8)
    pos = np.cumsum(counts) - counts
    out[pos[keep]] = t2[keep]
    ins_pos = pos[insm] + keep[insm].astype(np.int64)
    out[ins_pos] = random_seq(rng, int(insm.sum()))
    return out


def make_pair(rng, tlen, err=0.1, tail=0, n_rate=0.0, skew=0, lower=False, iupac=False):
    """One (query, target) pair.
    tail>0 appends `tail` random bases to the query (forces Z-drop); tail<0 cuts the query and replaces the rest
    with random sequence from that fraction point; skew adds/removes bases to make |qlen-tlen| large."""
    t = random_seq(rng, tlen)
    q = mutate_fast(rng, t, sub=err * 0.4, ins=err * 0.3, dele=err * 0.3)
    if tail > 0:
        q = np.concatenate([q, random_seq(rng, tail)])
    elif tail < 0:
        cut = int(len(q) * rng.random())
        q = np.concatenate([q[:cut], random_seq(rng, len(q) - cut)])
    if skew > 0:
        q = np.concatenate([q, mutate_fast(rng, random_seq(rng, skew), 0, 0, 0)])
    elif skew < 0:
        q = q[:max(1, len(q) + skew)]
    if n_rate > 0:
        for s in (q, t):
            m = rng.random(len(s)) < n_rate
            s[m] = ord('N')
    if iupac:
        for s in (q, t):
            m = rng.random(len(s)) < 0.01
            s[m] = np.frombuffer(b"RYKMSWBDHV", dtype=np.uint8)[rng.integers(0, 10, size=int(m.sum()))]
    if lower:
        q = q | 0x20
    if len(q) == 0:
        q = random_seq(rng, 1)
    return q, t


def make_pairs(seed, n, len_lo, len_hi, **kw):
    rng = np.random.default_rng(seed)
    pairs = []
    for _ in range(n):
        tlen = int(rng.integers(len_lo, len_hi + 1))
        k = dict(kw)
        if k.get("mixed"):
            k.pop("mixed")
            r = rng.random()
            k["err"] = float(rng.choice([0.01, 0.05, 0.1, 0.2, 0.35]))
            k["tail"] = int(rng.integers(50, 600)) if r < 0.25 else (-1 if r < 0.45 else 0)
            k["skew"] = int(rng.integers(-tlen // 2, tlen)) if rng.random() < 0.2 else 0
            k["n_rate"] = 0.02 if rng.random() < 0.15 else 0.0
        pairs.append(make_pair(rng, tlen, **k))
    return pairs
