"""Generate tests/golden/ref_vectors.json.gz from the REFERENCE kernel itself.

Runs only in the authoring container: it needs oracle/_ref/libagatha_ref_host.so, which oracle/Makefile
builds from /root/reference/AGAThA/src/kernels/agatha_kernel.h (compiled as single-lane host code,
SURVEY.md Appendix B). The reference ships no golden vectors of its own, so these
are the pin: inputs + the (score, query_end, target_end) the reference's code produces for them.

    python tests/golden/make_golden.py
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle_py as op  # noqa: E402
from pairgen import make_pairs      # noqa: E402

# (name, seed, n, len_lo, len_hi, generator kwargs, scoring overrides)
GROUPS = [
    ("default_w751", 11, 40, 600, 2600, dict(mixed=True), dict()),
    ("w63_sw3", 12, 120, 20, 500, dict(mixed=True), dict(band_width=63)),
    ("w15_sw1_z20", 13, 120, 5, 200, dict(mixed=True), dict(band_width=15, slice_width=1, z_threshold=20)),
    ("w31_sw7_m2", 14, 120, 5, 300, dict(mixed=True), dict(band_width=31, slice_width=7, match=2, gap_open=4)),
    ("w127_nozdrop", 15, 60, 100, 900, dict(mixed=True), dict(band_width=127, z_threshold=-1)),
    ("tiny", 16, 200, 1, 24, dict(mixed=True), dict(band_width=7, z_threshold=10)),
    ("n_rich_lower", 17, 60, 50, 400, dict(err=0.1, n_rate=0.05, lower=True), dict(band_width=63)),
    ("iupac", 18, 60, 50, 400, dict(err=0.1, iupac=True), dict(band_width=63)),
    ("skewed_bandexit", 19, 80, 100, 700, dict(err=0.05, skew=600), dict(band_width=63, z_threshold=-1)),
]


def main():
    ref = op.RefHost()
    groups = []
    for name, seed, n, lo, hi, gkw, pkw in GROUPS:
        pairs = make_pairs(seed, n, lo, hi, **gkw)
        params = dict(op.DEFAULT_PARAMS)
        params.update(pkw)
        res = ref.align_pairs(pairs, op.make_params(**params))
        groups.append(dict(name=name, params=params,
                           queries=[bytes(q).decode() for q, _ in pairs],
                           targets=[bytes(t).decode() for _, t in pairs],
                           expected=res.tolist()))
        print(name, n, "pairs")
    out = os.path.join(HERE, "ref_vectors.json.gz")
    with gzip.open(out, "wt", compresslevel=9) as f:
        json.dump(dict(source="reference agatha_kernel.h compiled as host code (oracle/_ref/libagatha_ref_host.so)",
                       groups=groups), f)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
