"""Deterministic synthetic head maps for tests and benchmarks.

Two regimes: (R) "random-init" -- all logits ~ N(0,1): nearly every row clears valid_thresh, the worst
case for candidate selection; (T) "trained-like" -- low objectness/class logits with a few boosted,
spatially clustered cells, so that a handful of real, overlapping detections exist per frame.
"""
from __future__ import annotations

import numpy as np


def grid_sizes(size: int):
    """network order: strides 32, 16, 8 (yolo3.py:416-417)"""
    return [size // 32, size // 16, size // 8]


def random_heads_np(rng, B, C, size, std=1.0, A=3):
    return [rng.normal(0, std, size=(B, A * (5 + C), g, g)).astype(np.float32) for g in grid_sizes(size)]


def trained_like_heads(rng, B, C, size, boost_frac=0.004, A=3):
    heads = []
    for g in grid_sizes(size):
        h = np.empty((B, A, 5 + C, g, g), dtype=np.float32)
        h[:, :, 0:2] = rng.normal(0, 1, size=h[:, :, 0:2].shape)
        h[:, :, 2:4] = rng.normal(0, 0.5, size=h[:, :, 2:4].shape)
        h[:, :, 4] = rng.normal(-6, 1.5, size=h[:, :, 4].shape)
        h[:, :, 5:] = rng.normal(-3, 1.5, size=h[:, :, 5:].shape)
        n = max(1, int(boost_frac * g * g))
        for b in range(B):
            for _ in range(n):
                y, x, a = rng.randint(g), rng.randint(g), rng.randint(A)
                c = rng.randint(C)
                for dy in (0, 1):
                    for dx in (0, 1, 2):
                        yy, xx = min(g - 1, y + dy), min(g - 1, x + dx)
                        h[b, a, 4, yy, xx] += 10
                        h[b, a, 5 + c, yy, xx] += 6
                        h[b, :, 4, yy, xx] += 4        # neighbouring anchors fire too -> overlaps
        heads.append(h.reshape(B, A * (5 + C), g, g))
    return heads


def random_heads_cuda(B, C, size, seed, device, A=3, regime="R"):
    """Device-side generation for the large benchmark shapes (no host copy)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    heads = []
    for s in grid_sizes(size):
        h = torch.randn((B, A, 5 + C, s, s), generator=g, device=device, dtype=torch.float32)
        if regime == "T":
            h[:, :, 2:4] *= 0.5
            h[:, :, 4] = h[:, :, 4] * 1.5 - 6
            h[:, :, 5:] = h[:, :, 5:] * 1.5 - 3
            boost = torch.rand((B, 1, 1, s, s), generator=g, device=device) < 0.004
            h[:, :, 4:5] += boost * 12.0
            h[:, :, 5:6] += boost * 6.0
        elif regime == "S":
            # sparse: what a trained detector gives on a frame with one or two objects -- objectness ~ 1e-4 everywhere but
            # in a few cells (all anchors of the cell confident), so tens to a few hundred scores pass valid_thresh
            h[:, :, 2:4] *= 0.5
            h[:, :, 4] = h[:, :, 4] - 9
            h[:, :, 5:] = h[:, :, 5:] * 1.5 - 3
            boost = torch.rand((B, 1, 1, s, s), generator=g, device=device) < 0.0005
            h[:, :, 4:5] += boost * 15.0
            h[:, :, 5:6] += boost * 6.0
        heads.append(h.reshape(B, A * (5 + C), s, s).contiguous())
    return heads
