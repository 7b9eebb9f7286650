"""CPU tests of the order-independent digest bench.py uses to compare the edge-filter verdicts of a (multi-)GPU run
with the reference's (bench.py: filters_sum_digest / check_filters_parity; goldens in tests/golden/sum_digests.json,
recorded by tests/golden/make_sum_digests.py --filters from the reference's own jobs)."""
import os
import sys

import numpy as np
import pytest

import util
from muchsalsa_b200 import synth

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import filters_sum_digest  # noqa: E402

FIELDS = ("contraction_order", "deletable", "n_kept_orders", "weight", "consensus")


def _random_verdicts(n, seed=0):
    rng = np.random.default_rng(seed)
    return dict(contraction_order=rng.integers(-1, 4, n).astype(np.int32), deletable=rng.integers(0, 2, n).astype(np.uint8),
                n_kept_orders=rng.integers(0, 5, n).astype(np.uint32),
                weight=rng.integers(0, 1 << 62, n).astype(np.uint64), consensus=rng.integers(97, 100, n).astype(np.uint8))


def test_digests_of_slices_add_up_to_the_digest_of_the_table():
    f = _random_verdicts(10_000)
    whole = filters_sum_digest(f)
    cuts = [0, 1, 3333, 3333, 9000, 10_000]  # uneven slices, one of them empty
    parts = [filters_sum_digest({k: f[k][a:b] for k in FIELDS}, base=a) for a, b in zip(cuts[:-1], cuts[1:])]
    assert sum(p["sum"] for p in parts) % (1 << 64) == whole["sum"]
    for k in ("n_edges", "contraction_edges", "deletable_edges"):
        assert sum(p[k] for p in parts) == whole[k]


@pytest.mark.parametrize("field", FIELDS)
def test_digest_sees_every_field_and_the_position(field):
    f = _random_verdicts(1000, seed=3)
    base = filters_sum_digest(f)["sum"]
    g = {k: v.copy() for k, v in f.items()}
    g[field][517] = g[field][517] + 1
    assert filters_sum_digest(g)["sum"] != base
    h = {k: v.copy() for k, v in f.items()}
    for k in FIELDS:  # two edges exchanged: the same multiset of verdicts at other positions
        h[k][[10, 20]] = h[k][[20, 10]]
    if any(f[k][10] != f[k][20] for k in FIELDS):
        assert filters_sum_digest(h)["sum"] != base


@pytest.mark.parametrize("n", [1, 2, 3])
def test_recorded_filter_digests_are_the_references(reference, n):
    """The goldens bench.py compares with are what the reference's own jobs give on the same PAF today."""
    paf = b"".join(synth.rank_chunk("small", r, n) for r in range(n))
    got = filters_sum_digest(reference.filters(paf, 300))
    gold = util.load_golden("sum_digests.json")[f"filters_smallx{n}"]
    assert all(got[k] == gold[k] for k in ("n_edges", "sum", "contraction_edges", "deletable_edges"))
