"""Shared helpers of the post-processing tests (CPU host-core tests and -m gpu parity tests).

A thresholded / arg-maxed output can legitimately differ from the oracle at pixels that sit on a decision boundary
(the reference's own CPU and CUDA interpolation kernels round differently there), so comparisons are exact everywhere
except where the oracle's own margin is below `tol`; the tests additionally bound how many such pixels there are.
"""
import torch


def assert_masks_match(ours, ref_masks, ref_logits, tol=1e-4, max_boundary_frac=2e-3):
    """ours, ref_masks bool (n, T, H, W); ref_logits the oracle's resized logits.  Mismatches only where |logit| < tol."""
    ours, ref_masks = ours.bool().cpu(), ref_masks.bool().cpu()
    assert ours.shape == ref_masks.shape, (ours.shape, ref_masks.shape)
    bad = ours != ref_masks
    if bad.any():
        worst = ref_logits.cpu()[bad].abs().max().item()
        assert worst < tol, f"{int(bad.sum())} mismatching pixels, largest |logit| among them {worst:.3e}"
        assert bad.float().mean().item() <= max_boundary_frac, f"boundary mismatches {bad.float().mean().item():.2e}"
    return int(bad.sum())


def assert_labels_match(ours, ref_labels, ref_scores, tol=1e-4, max_boundary_frac=2e-3):
    """Arg-max labels over dim 0 of ref_scores (C, ...).  A differing label must be a near-tie in the oracle's scores."""
    ours, ref_labels, ref_scores = ours.cpu().long(), ref_labels.cpu().long(), ref_scores.cpu()
    assert ours.shape == ref_labels.shape, (ours.shape, ref_labels.shape)
    bad = ours != ref_labels
    if bad.any():
        top = ref_scores.gather(0, ref_labels[None])[0]
        alt = ref_scores.gather(0, ours.clamp(0, ref_scores.shape[0] - 1)[None])[0]
        gap = (top - alt)[bad].abs().max().item()
        assert gap < tol, f"{int(bad.sum())} differing labels, largest score gap {gap:.3e}"
        assert bad.float().mean().item() <= max_boundary_frac
    return int(bad.sum())


def sort_instances(scores, labels, ids, masks=None):
    """torch.topk(sorted=False) leaves the order of the selected instances unspecified: compare in canonical order
    (score descending, then label, then id)."""
    scores, labels, ids = (torch.as_tensor(x) for x in (scores, labels, ids))
    key = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), int(labels[i]), int(ids[i])))
    key = torch.as_tensor(key, dtype=torch.long)
    out = (scores[key], labels[key], ids[key])
    return out + ((masks[key],) if masks is not None else ())
