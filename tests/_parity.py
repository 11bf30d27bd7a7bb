"""Parity metrics shared by the GPU parity tests, smoke() and bench.py's checker leg.

north_star's bar for floating point is `|got - want| <= 1e-6 + 1e-5 * |want|` per element.  SURVEY.md Appendix C shows
the reference's OWN fp32 evaluation violates that form against an fp64 evaluation on 1e-4..1e-3 of the elements at
BASELINE shapes (long fp32 sums at outputs of magnitude O(1-10)), so every comparison reports three numbers:

  viol      fraction of elements outside  1e-6 + 1e-5 |want|          (the strict form, never softened)
  maxrel    max |got - want| / max |want|                              (max-norm relative error)
  and, when an fp64 evaluation is available, the same two for the reference-order fp32 oracle -- the yardstick.
"""
import torch

RTOL, ATOL = 1e-5, 1e-6


def violation_fraction(got, want, rtol=RTOL, atol=ATOL):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    if want.numel() == 0:
        return 0.0
    bad = (got - want).abs() > atol + rtol * want.abs()
    return float(bad.double().mean())


def maxnorm_rel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    if want.numel() == 0:
        return 0.0
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-300))


def compare(ours, ref32, ref64=None):
    """{name: {viol_vs_ref32, maxrel_vs_ref32[, viol_vs_fp64, maxrel_vs_fp64, ref32_viol_vs_fp64, ref32_maxrel_vs_fp64]}}"""
    rep = {}
    for k, w32 in ref32.items():
        g = ours[k]
        e = {"viol_vs_ref32": violation_fraction(g, w32), "maxrel_vs_ref32": maxnorm_rel(g, w32)}
        if ref64 is not None:
            w64 = ref64[k]
            e.update(viol_vs_fp64=violation_fraction(g, w64), maxrel_vs_fp64=maxnorm_rel(g, w64),
                     ref32_viol_vs_fp64=violation_fraction(w32, w64), ref32_maxrel_vs_fp64=maxnorm_rel(w32, w64))
        rep[k] = e
    return rep


def summarise(rep):
    """Worst case over tensors of each metric (what bench.py prints as `parity`)."""
    keys = sorted({m for e in rep.values() for m in e})
    return {m: max(e.get(m, 0.0) for e in rep.values()) for m in keys}
