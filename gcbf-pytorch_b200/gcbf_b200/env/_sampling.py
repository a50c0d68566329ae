"""Host-side rejection sampling used by reset() (scaffolding, not on the hot path)."""
import torch


def sample_separated(count: int, dim: int, side: float, min_dist: float, avoid=None, avoid_dist: float = 0.0,
                     max_tries: int = 200000) -> torch.Tensor:
    """`count` points in [0, side]^dim, pairwise farther than min_dist (and farther than avoid_dist from `avoid`)."""
    pts = torch.zeros(count, dim)
    i = tries = 0
    while i < count:
        tries += 1
        if tries > max_tries:
            raise RuntimeError(f'could not place {count} points with spacing {min_dist} in a box of side {side}; '
                               'increase area_size (the reference has the same limitation, SURVEY section 0)')
        cand = torch.rand(dim) * side
        if i > 0 and torch.norm(pts[:i] - cand, dim=1).min() <= min_dist:
            continue
        if avoid is not None and avoid.numel() and torch.norm(avoid - cand, dim=1).min() <= avoid_dist:
            continue
        pts[i] = cand
        i += 1
    return pts
