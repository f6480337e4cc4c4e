"""Per-image channels-last texels for callers that hand the SAME feature-map tensors to many reference views.

With `FeatureCache` attached (feature_cache.py) the feature pyramid of an image is one set of tensor
objects that every reference view using the image receives again; the reference views of a scene block share 4 of their 5
images.  `TexelPool` keeps the channels-last layout of such maps resident -- one pool [S,H,W,C] per (device, C, H, W), an
LRU over its slots keyed by the identity of the map tensor -- so an image is laid out once and the sweeps name their views by
pool slot (`sweep.cost_volume(pool, ..., view_slots=...)`, D3dCostVolumeArgs.texel_slots): per reference view and stage one
relayout launch (the image new to the walk) instead of V.

An entry is valid only while its tensor is alive and unmodified: it holds a weak reference (a dead tensor's id can be reused)
and the tensor's version counter.  Everything runs on the current stream, like the sweeps that read the pool.  Inference only.
"""
from __future__ import annotations

import collections
import weakref
from typing import List, Optional, Sequence, Tuple

import torch

from . import sweep


class TexelPool:
    _cuda_only = True      # (the CPU test of the slot bookkeeping clears it and substitutes the relayout)

    def __init__(self, capacity: int = 16):
        if capacity < 2:
            raise ValueError("capacity must be >= 2")
        self.capacity = capacity
        self.pools = {}                                   # (device, C, H, W) -> state
        self.hits = self.misses = 0

    def _state(self, t: torch.Tensor):
        c, h, w = t.shape[-3:]
        key = (t.device, c, h, w)
        st = self.pools.get(key)
        if st is None:
            slots = min(self.capacity, (2 ** 31 - 1) // (h * w))
            st = {"texels": torch.empty((slots, h, w, c), device=t.device, dtype=torch.float32),
                  "entries": collections.OrderedDict(),   # id(tensor) -> (slot, weakref, version), oldest first
                  "free": list(range(slots))}
            self.pools[key] = st
        return st

    def lookup(self, maps: Sequence[torch.Tensor]) -> Optional[Tuple[torch.Tensor, List[int]]]:
        """maps: the V feature maps of one reference view, [C,H,W] or [1,C,H,W] fp32 CUDA tensors of one shape -> (pool
        [S,H,W,C], slot of every map).  None when the view does not fit the pool (the caller lays out a dense block)."""
        first = maps[0]
        if any((self._cuda_only and not m.is_cuda) or m.dtype != torch.float32 or m.shape[-3:] != first.shape[-3:] or m.device != first.device
               or m.numel() != first.shape[-3] * first.shape[-2] * first.shape[-1] for m in maps):
            return None
        st = self._state(first)
        if len({id(m) for m in maps}) > st["texels"].shape[0]:
            return None
        entries, slots, mine = st["entries"], [], {id(m) for m in maps}
        for m in maps:
            ent = entries.get(id(m))
            if ent is not None and ent[1]() is m and ent[2] == m._version:
                entries.move_to_end(id(m))
                self.hits += 1
                slots.append(ent[0])
                continue
            self.misses += 1
            if ent is not None:                           # a dead or modified tensor's entry: its slot is reused
                slot = entries.pop(id(m))[0]
            elif st["free"]:
                slot = st["free"].pop()
            else:
                victim = next(k for k in entries if k not in mine)
                slot = entries.pop(victim)[0]
            sweep.to_texels([m], out=st["texels"][slot:slot + 1])
            entries[id(m)] = (slot, weakref.ref(m), m._version)
            slots.append(slot)
        return st["texels"], slots

    def clear(self) -> None:
        self.pools.clear()

    def stats(self) -> dict:
        return {"hits": self.hits, "misses": self.misses,
                "bytes": sum(s["texels"].numel() * 4 for s in self.pools.values())}
