"""Parity rules of SURVEY.md §8(c), shared by the CPU (host-emulated logic) and GPU tests.

1. image count must match exactly;
2. A0, A2, A4 relative difference <= RTOL = 1e-9 (BASELINE.json north_star tolerance);
3. exclusion mask, always REPORTED: near-critical (min_i |1-|W2_i|^2| < 1e-3), multipole series not
   converging (|A4-A2| > |A2-A0|), or a reference root residual within a decade of the 1e-6
   selection threshold (image-count ambiguity);
4. disputed points (masked or not) are adjudicated against the 40-digit evaluation of the same
   formulae: pass iff image count == truth count and |got-truth| <= max(|ref-truth|, RTOL |truth|);
5. image-count disputes are ACCOUNTED FOR, not just adjudicated: each one must either carry MLM_FLAG_EXTRA with
   nimg - ((flags >> 6) & 3) == the reference's count (the kernel kept an image that the plain 1e-6 test of
   multipoles.py:183 rejects on its own roots), or lie in rule 3's mask (the REFERENCE's root residual is within a
   decade of its threshold: np.roots' unpolished root fails the test that the kernel's polished root passes);
   the number of disputes OUTSIDE rule 3's mask is capped (default 1 % of the sample); those inside it are positions
   the rules exclude (the reference's own answer is unreliable there) and are only reported and adjudicated.
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-9


def mask_from_reference(ref_A, residuals, minJ):
    """Rule 3 from reference-side data: ref_A (n,3), residuals (n,5) (NaN padded), minJ (n,)."""
    with np.errstate(invalid="ignore"):
        amb = ((residuals > 1e-7) & (residuals < 1e-5)).any(axis=1)
    series = np.abs(ref_A[:, 2] - ref_A[:, 1]) > np.abs(ref_A[:, 1] - ref_A[:, 0])
    return (minJ < 1e-3) | series | amb


FLAG_EXTRA, FLAG_EXTRA_SHIFT = 2, 6


def compare(got_n, got_A, ref_n, ref_A, mask=None, truth_fn=None, max_adjudicate=None, got_flags=None,
            max_dispute_frac=0.01, strict=True):
    """Returns a report dict; raises AssertionError on a parity failure (strict=False, used by bench.py: never
    raises, the failures are in the report -- `failed` lists which rules broke; more disputes than max_adjudicate
    are then adjudicated on a seeded subset of that size).

    got_A/ref_A: (n,3) arrays.  truth_fn(i) -> (n_truth, A0, A2, A4) for rule 4 (optional).
    got_flags: the kernel's flag bytes (rule 5; None skips the accounting of count disputes).
    max_adjudicate: how many disputes are adjudicated at most (default max_dispute_frac of the sample, at least 8; in
    strict mode more disputes than that fail); max_dispute_frac: cap on the disputes outside the mask.
    """
    got_n = np.asarray(got_n).astype(np.int64).ravel()
    ref_n = np.asarray(ref_n).astype(np.int64).ravel()
    n = len(ref_n)
    mask = np.zeros(n, bool) if mask is None else mask
    with np.errstate(all="ignore"):
        rel = np.max(np.abs(got_A / ref_A - 1.0), axis=1)
    rel = np.where((got_A == ref_A).all(axis=1), 0.0, rel)        # 0/0 when both are zero-image
    bad_n = got_n != ref_n
    bad_v = ~(rel <= RTOL)
    disputed = np.where(bad_n | bad_v)[0]
    unmasked_disputes = int(np.sum((bad_n | bad_v) & ~mask))
    if max_adjudicate is None:
        max_adjudicate = max(8, int(np.ceil(max_dispute_frac * n)))
    failed = []

    def require(cond, what, msg):
        if not cond:
            if strict:
                raise AssertionError(msg)
            failed.append(what)

    report = dict(n=n, masked=int(mask.sum()), masked_frac=float(mask.mean()) if n else 0.0,
                  nimg_mismatch=int(bad_n.sum()), rel_gt_tol=int((bad_v & ~bad_n).sum()),
                  unmasked_disputes=unmasked_disputes,
                  max_rel_clean=float(np.nanmax(rel[~(bad_n | bad_v)])) if (~(bad_n | bad_v)).any() else 0.0,
                  adjudicated=0, adjudication_failures=[])
    if got_flags is not None:
        fl = np.asarray(got_flags).astype(np.int64).ravel()
        extra = (fl >> FLAG_EXTRA_SHIFT) & 3
        by_flag = bad_n & ((fl & FLAG_EXTRA) != 0) & (got_n - extra == ref_n)
        by_mask = bad_n & ~by_flag & mask
        report.update(nimg_mismatch_flagged_extra=int(by_flag.sum()), nimg_mismatch_ref_ambiguous=int(by_mask.sum()),
                      nimg_mismatch_unaccounted=int((bad_n & ~by_flag & ~by_mask).sum()),
                      flagged_extra=int(((fl & FLAG_EXTRA) != 0).sum()))
        clean = (fl == 0) & ~mask
        report.update(rel_gt_tol_unflagged_unmasked=int((bad_v & ~bad_n & clean).sum()))
        require(report["nimg_mismatch_unaccounted"] == 0, "nimg_unaccounted",
                f"image-count disputes without flag or mask: {report}")
    if not strict:
        report["failed"] = failed
    if len(disputed) == 0:
        return report
    if truth_fn is None:
        # without a truth evaluator only masked disputes are tolerated
        require(unmasked_disputes == 0, "unmasked_disputes_without_adjudicator",
                f"{unmasked_disputes} unmasked parity disputes, no adjudicator: {report}")
        return report
    cap = max(8, int(np.ceil(max_dispute_frac * n)))
    require(unmasked_disputes <= cap, "too_many_disputes", f"{unmasked_disputes} disputes outside the mask (cap {cap}): {report}")
    if strict:
        require(len(disputed) <= max_adjudicate, "too_many_to_adjudicate", f"too many disputes to adjudicate ({len(disputed)}): {report}")
    if len(disputed) > max_adjudicate:
        disputed = np.sort(np.random.default_rng(1).choice(disputed, max_adjudicate, replace=False))
        report["adjudicated_subset_of"] = int(np.sum(bad_n | bad_v))
    for i in disputed:
        t = truth_fn(int(i))
        tn, tA = int(t[0]), np.array(t[1:4])
        g, r = got_A[i], ref_A[i]
        ok_n = got_n[i] == tn
        with np.errstate(all="ignore"):
            ok_v = bool(np.all(np.abs(g - tA) <= np.maximum(np.abs(r - tA), RTOL * np.abs(tA)) * (1 + 1e-12)))
        report["adjudicated"] += 1
        if not (ok_n and ok_v):
            report["adjudication_failures"].append((int(i), int(got_n[i]), int(ref_n[i]), tn, g.tolist(), r.tolist(),
                                                    tA.tolist()))
    require(not report["adjudication_failures"], "adjudication", str(report))
    return report
