"""Locality schedules for the MSDA gather kernel.

When the queries of a deformable-attention call are the pixels of the feature pyramid itself (encoder
self-attention: Lq == sum(H_l*W_l), P/mask2former/modeling/pixel_decoder/msdeformattn.py:124), walking the
(query, head) items as 2-D tiles per head keeps the value rows a CTA gathers inside a small window that
stays resident in that SM's L1.  The schedule is only a permutation of work: results do not depend on it.
"""
import functools

import numpy as np
import torch


@functools.lru_cache(maxsize=64)
def _tiled_order_np(shapes, num_heads, tile_h, tile_w):
    chunks = []
    start = 0
    for (H, W) in shapes:
        ys, xs = np.arange(H), np.arange(W)
        q = (start + ys[:, None] * W + xs[None, :]).astype(np.int64)
        for ty in range(0, H, tile_h):
            for tx in range(0, W, tile_w):
                tile = q[ty:ty + tile_h, tx:tx + tile_w].reshape(-1)
                for m in range(num_heads):
                    chunks.append(tile * num_heads + m)
        start += H * W
    order = np.concatenate(chunks).astype(np.int32)
    assert order.size == start * num_heads
    return order


_device_cache = {}


def tiled_item_order(shapes, num_heads, device, tile=(16, 16)):
    """int32 device tensor of length sum(H*W)*num_heads: item ids q*num_heads+m in tile-major order."""
    key = (tuple((int(h), int(w)) for h, w in shapes), int(num_heads), tuple(tile), str(device))
    t = _device_cache.get(key)
    if t is None:
        t = torch.from_numpy(_tiled_order_np(key[0], key[1], tile[0], tile[1])).to(device)
        _device_cache[key] = t
    return t
