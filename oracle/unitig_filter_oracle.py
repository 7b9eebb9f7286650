"""CPU restatement of the pipeline's unitig filter (pipeline/unitig_filter.py) — TEST INFRASTRUCTURE, never imported by
the product.  tests/golden/make_prefilter_golden.py checks it against the UNMODIFIED script (run whole, with a stand-in
for Bio.SeqIO.index_db backed by a plain FASTA parser) and records golden vectors.

What the script computes, per BLOCK of consecutive PAF lines naming the same unitig:
  unitig_filter.py:34-75    pass 1: the number of distinct reads of the block, the unitig length of its first line,
                            and the maximum of a per-position coverage profile in which every read counts ONCE
                            (its first line inside the block; illuminaStart .. illuminaEnd = c2 .. c3 - 1 inclusive);
                            mapping_freq[unitig] = the values of the LAST block of that unitig
  :83-97                    cutoff = Q3 + 1.5 IQR of the maximum coverages, q3 = the 75th percentile
  :102-127, 132-176         pass 2: a block whose unitig's maximum is above the cutoff gets a profile of ALL its lines
                            (no per-read dedup here) and is cut into the maximal runs of positions with coverage <= q3
                            that are >= 500 long; every other block's unitig is written whole (once per block)
"""
from __future__ import annotations

import numpy as np


def _lines(paf: bytes):
    text = paf.decode("latin-1")
    start = 0
    while start < len(text):
        nl = text.find("\n", start)
        yield text[start:] if nl < 0 else text[start:nl + 1]
        if nl < 0:
            return
        start = nl + 1


def blocks_pass1(paf: bytes):
    """[(unitig, n_reads, length, max_coverage)] per block in file order (unitig_filter.py:34-75)."""
    out, cur, length, reads, profile = [], None, 0, [], []
    for line in _lines(paf):
        f = line.rstrip().split("\t")
        illu, ilen, a, b, nano = f[0], int(f[1]), int(f[2]), int(f[3]) - 1, f[5]
        int(f[6])
        if cur == illu:
            if nano not in reads:
                reads.append(nano)
                for pos in range(a, b + 1):
                    profile[pos] += 1
        else:
            if cur is not None:
                out.append((cur, len(reads), length, max(profile)))
            profile = [0] * ilen
            for pos in range(a, b + 1):
                profile[pos] += 1
            cur, length, reads = illu, ilen, [nano]
    if cur is not None:
        out.append((cur, len(reads), length, max(profile)))
    return out


def cutoffs(blocks):
    """(upper whisker, q3) over mapping_freq's values = the LAST block of every unitig, in first-seen order
    (unitig_filter.py:83-97)."""
    last = {}
    for u, _, _, mx in blocks:
        last[u] = mx
    vals = list(last.values())
    q1, q3 = np.percentile(vals, 25), np.percentile(vals, 75)
    return float(q3 + 1.5 * (q3 - q1)), float(q3), last


def blocks_pass2(paf: bytes, last_max: dict, cutoff: float, q3: float):
    """[(unitig, None)] for a block written whole, [(unitig, [(start, end), ...])] for an outlier block
    (unitig_filter.py:102-176)."""
    out, cur, profile = [], None, None

    def cut(profile):
        frags, sub, below = [], (0, 0), False
        for i, cov in enumerate(profile):
            if cov <= q3:
                sub = (sub[0], i) if below else (i, i)
                below = True
            else:
                if below and sub[1] - sub[0] + 1 >= 500:
                    frags.append(sub)
                    sub = (0, 0)
                below = False
        if below and sub[1] - sub[0] + 1 >= 500:
            frags.append(sub)
        return frags

    for line in _lines(paf):
        f = line.rstrip().split("\t")
        illu, ilen, a, b = f[0], int(f[1]), int(f[2]), int(f[3]) - 1
        if cur == illu:
            if last_max[illu] > cutoff:
                for pos in range(a, b + 1):
                    profile[pos] += 1
        else:
            if cur is not None and last_max[cur] > cutoff:
                out.append((cur, cut(profile)))
            if last_max[illu] > cutoff:
                profile = [0] * ilen
                for pos in range(a, b + 1):
                    profile[pos] += 1
            else:
                out.append((illu, None))
            cur = illu
    if cur is not None and last_max[cur] > cutoff:
        out.append((cur, cut(profile)))
    return out
