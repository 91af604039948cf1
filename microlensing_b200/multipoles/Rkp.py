"""Q_(p-n,n) factors entering the expansion of z(zeta) of the lens equation (Cassan 2017).

Python-3 counterpart of the reference's offline formula generator
(``microlensing/multipoles/Rkp.py``, Python 2 only: ``Rkp.py:16-18``).  ``Q(p)`` prints, for every
order m = 3..p and every n = 0..m, the factor Q[m-n, n] as a sum over W_j of products of a_kl.
Instead of the reference's string rewriting (``Rkp.py:37-47,75-97``) the terms are enumerated
as set partitions of the m derivative labels (j-1 blocks <-> W_j) with identical monomials
collected, which is the form pasted into ``multipoles.py:98-133``.  Not numeric, not on the
GPU path; ``tools/gen_qkl.py`` uses the same enumeration to emit the CUDA code.
"""
from __future__ import annotations

from collections import Counter

__all__ = ["Q", "q_terms", "_akl"]


def _set_partitions(items):
    if len(items) == 1:
        yield [items]
        return
    first, rest = items[0], items[1:]
    for part in _set_partitions(rest):
        for i in range(len(part)):
            yield part[:i] + [[first] + part[i]] + part[i + 1:]
        yield [[first]] + part


def q_terms(k, l):
    """{j: {((k1,l1),(k2,l2),...): multiplicity}}: Q[k,l] = sum_j W_j * sum mult * prod a_{ki li},
    j = 3 .. k+l+1 (the j = 2 term, W2 * a_kl, is what ``_akl`` solves for)."""
    labels = "a" * k + "b" * l
    out: dict = {}
    for part in _set_partitions(list(range(k + l))):
        j = len(part) + 1
        if j < 3:
            continue
        mono = tuple(sorted(((sum(labels[i] == "a" for i in blk), sum(labels[i] == "b" for i in blk))
                             for blk in part), reverse=True))
        out.setdefault(j, Counter())[mono] += 1
    return {j: dict(c) for j, c in sorted(out.items())}


def Q(p):
    """Compute factors Q(p-n, n) and print them (reference ``Rkp.py:13``; returns None).

    Parameter
    ---------
    p : int
        Order of expansion (p >= 3; cf. publication).
    """
    for m in range(3, p + 1):
        print(" Order p=" + str(m))
        for n in range(m + 1):
            print("\n   Q[" + str(m - n) + "," + str(n) + "] :")
            for j, mons in q_terms(m - n, n).items():
                terms = []
                for mono, c in sorted(mons.items(), reverse=True):
                    prod = " * ".join("a" + str(a) + str(b) for a, b in mono)
                    terms.append((str(c) + " * " if c != 1 else "") + prod)
                print("\n     W" + str(j) + " * ( " + " + ".join(terms) + " )")
        print("\n")


def _akl(mu, W2, Qkl):
    """Compute a(p-n, n) factors from Q(p-n, n) (reference ``Rkp.py:99-103``); GPU-backed."""
    from .multipoles import _akl as _gpu_akl
    return _gpu_akl(mu, W2, Qkl)
