"""Shared helpers for the parity tests."""
import numpy as np
import torch


def bf16_round(a):
    """Round-to-nearest-even to bfloat16 and back (what the device stores between layers)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16().float().numpy()


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def softmax_probs(rng, shape, scale=2.0):
    """softmax(N(0, scale^2)) float32 - the synthetic expert output of SURVEY.md 8(d)."""
    s = rng.normal(0, scale, size=shape).astype(np.float32)
    e = np.exp(s - s.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


def top2_margin(score):
    """Gap between the best and second-best class score per pixel."""
    part = np.partition(score, -2, axis=-1)
    return part[..., -1] - part[..., -2]


def assert_labels_match(got, score_ref, atol_margin, max_flip_frac=0.0):
    """Device labels must equal the oracle argmax wherever the oracle's own decision margin
    exceeds `atol_margin` (a few ulps of the score); elsewhere either of the tied classes is
    acceptable.  Returns the fraction of near-tie pixels."""
    ref = np.argmax(score_ref, axis=-1)
    margin = top2_margin(score_ref)
    differs = got != ref
    decisive = margin > atol_margin
    assert not (differs & decisive).any(), (
        '%d labels differ on pixels with a decisive margin' % int((differs & decisive).sum()))
    # on near-ties the device must still have picked a class within the margin of the best
    if differs.any():
        picked = np.take_along_axis(score_ref, got[..., None].astype(np.int64), -1)[..., 0]
        best = score_ref.max(-1)
        assert (best[differs] - picked[differs] <= atol_margin).all()
    assert differs.mean() <= max(max_flip_frac, (~decisive).mean())
    return float((~decisive).mean())
