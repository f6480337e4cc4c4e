"""FeatureNet sharing across reference views.

`predict.py` runs FeatureNet on all V images of every reference view (`adamvs.py:570-574`, `cas_mvsnet.py:189-192`,
`msrednet.py:286-289`), although neighbouring reference views of a scene block use the same source images
(`viewpair.txt`): with V = 5 every image is pushed through the network about five times.  `FeatureCache` keeps the
feature pyramids of the most recently used images resident in HBM (286 MB per 1856 x 2752 image: 32 ch @ 1/4 +
16 ch @ 1/2 + 8 ch @ full resolution; 64 images = 18 GB of the B200's 180 GB) and hands them back instead of
recomputing them.

It attaches to the reference's own module WITHOUT wrapping it -- `feature_net.forward` is rebound on the instance,
so `state_dict` keys, `feature.out_channels` and checkpoint loading are untouched:

    cache = FeatureCache(capacity=64).attach(model.feature)
    with cache.views(image_ids_of_this_sample):       # ids in the order the model walks imgs[:, i]
        outputs = model(imgs, proj_matrices, depth_values)

Outside a `views(...)` block, with a batch larger than one, or once the ids of the block are used up, calls pass
straight through.  Inference only (cached tensors are detached).  A cached pyramid is the very tensor object the
first computation returned: results are bit-identical to recomputing with deterministic convolutions.
"""
from __future__ import annotations

import contextlib
from collections import OrderedDict
from typing import Hashable, Iterable, Optional


def _detach(tree):
    import torch
    if isinstance(tree, torch.Tensor):
        return tree.detach()
    if isinstance(tree, dict):
        return {k: _detach(v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_detach(v) for v in tree)
    return tree


def _nbytes(tree) -> int:
    import torch
    if isinstance(tree, torch.Tensor):
        return tree.numel() * tree.element_size()
    if isinstance(tree, dict):
        return sum(_nbytes(v) for v in tree.values())
    if isinstance(tree, (list, tuple)):
        return sum(_nbytes(v) for v in tree)
    return 0


class FeatureCache:
    """LRU cache of per-image feature pyramids, keyed by (image id, input shape, device)."""

    def __init__(self, capacity: int = 64, max_bytes: Optional[int] = None):
        if capacity < 1:
            raise ValueError("capacity must be >= 1")
        self.capacity = capacity
        self.max_bytes = max_bytes
        self.store: "OrderedDict[Hashable, object]" = OrderedDict()
        self.bytes = 0
        self.hits = 0
        self.misses = 0
        self._pending: Optional[list] = None
        self._net = None

    # ------------------------------------------------------------------ attachment
    def attach(self, feature_net):
        """Route `feature_net(img)` through the cache (rebinds `forward` on the instance only)."""
        if self._net is not None:
            raise RuntimeError("this cache is already attached to a network")
        inner = feature_net.forward                      # the class's bound method
        cache = self

        def forward(img, *args, **kwargs):
            key = cache._next_key(img) if not args and not kwargs else None
            if key is None:
                return inner(img, *args, **kwargs)
            hit = cache.store.get(key)
            if hit is not None:
                cache.store.move_to_end(key)
                cache.hits += 1
                return hit
            cache.misses += 1
            out = _detach(inner(img))
            cache._insert(key, out)
            return out

        feature_net.forward = forward
        self._net = feature_net
        return self

    def detach(self):
        if self._net is not None:
            del self._net.forward                        # the class's forward shows through again
            self._net = None

    # ------------------------------------------------------------------ per-sample ids
    @contextlib.contextmanager
    def views(self, image_ids: Iterable[Hashable]):
        """The next `len(image_ids)` FeatureNet calls are for these images, in this order."""
        self._pending = list(image_ids)
        try:
            yield self
        finally:
            self._pending = None

    def _next_key(self, img):
        if not self._pending:
            return None
        image_id = self._pending.pop(0)                  # consumed even when this call is not cacheable
        if img.dim() != 4 or img.shape[0] != 1:
            return None
        return (image_id, tuple(img.shape), str(img.device), str(img.dtype))

    # ------------------------------------------------------------------ storage
    def _insert(self, key, value):
        size = _nbytes(value)
        self.store[key] = value
        self.bytes += size
        while len(self.store) > self.capacity or (self.max_bytes is not None and self.bytes > self.max_bytes
                                                  and len(self.store) > 1):
            _, old = self.store.popitem(last=False)
            self.bytes -= _nbytes(old)

    def clear(self):
        self.store.clear()
        self.bytes = 0

    def stats(self) -> dict:
        n = self.hits + self.misses
        return {"hits": self.hits, "misses": self.misses, "hit_rate": self.hits / n if n else 0.0,
                "images": len(self.store), "bytes": self.bytes}
