"""Shared helpers of the GPU gradient-parity tests (TEST INFRASTRUCTURE).

Gradient parity against fp64 has one structural caveat that is a property of the REFERENCE FUNCTION, not of an implementation: the
coupling MLPs use leaky-ReLU, whose derivative jumps from 0.01 to 1 at zero.  A hidden pre-activation that lies closer to zero than an
arithmetic's forward error takes the other branch, and that one unit then moves the gradient of its row's image by ~1e-3..1e-2
relative.  The reference's own fp32 arithmetic shows it at its ~1e-7 threshold (measured: fp32 oracle vs fp64, B=64 x S=10, seed 77:
dfeat 1.3e-3 max-relative, 4 parameter tensors above 1e-3 - gpurun_out/grad_error_budget.json, tools/grad_error_budget.py); the
tensor-core path (forward error ~1e-5 absolute) at its threshold.  Away from such crossings the tensor-core gradients agree with
fp64 to 4e-5 (flat) / 1.1e-4 (dfeat) - exactly what a CPU emulation of the ideal 3-product scheme predicts
(profiles/r2_precision_schemes.json).  The tests therefore assert the 1e-3 bar on the Frobenius norm of every gradient (crossings
included) and, per image, allow only a small fraction of crossing-affected rows.
"""
import torch


def fro(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def flat_error(grads: dict, ref: dict):
    """(Frobenius error of the flat gradient, number of tensors above 1e-3, worst tensor name, worst tensor error)."""
    num = sum(float((grads[k].detach().double().cpu() - ref[k].double()).pow(2).sum()) for k in ref)
    den = sum(float(ref[k].double().pow(2).sum()) for k in ref)
    per = sorted(((fro(grads[k], ref[k]), k) for k in ref), reverse=True)
    return (num / den) ** 0.5, sum(1 for e, _ in per if e > 1e-3), per[0][1], per[0][0]


def per_image_errors(got, ref):
    """Relative error of every image's row of a (B, C) gradient."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return (got - ref).norm(dim=1) / (ref.norm(dim=1) + 1e-300)


def kink_aware_ok(got, ref, bar=1e-3, max_fraction=0.1, hard=0.1):
    """(ok, median error, images above the bar): the typical image meets `bar`; at most `max_fraction` of the images (at least one)
    may sit above it because of a leaky-ReLU crossing, none above `hard`."""
    e = per_image_errors(got, ref)
    n_over = int((e > bar).sum())
    allowed = max(1, int(max_fraction * e.numel()))
    return bool(e.median() < bar and n_over <= allowed and e.max() < hard), float(e.median()), n_over
