"""Host-to-host view pipeline: reference views arrive as pinned HOST tensors, depth + confidence maps
leave as pinned HOST tensors, and the host<->device copies of neighbouring views overlap the sweep.

    pipe = ViewPipeline(num_views=5, channels=32, height=688, width=464, num_depth=384, device="cuda:0")
    for i, view in enumerate(views):
        pipe.submit(view.features, view.proj, view.hyps, logits_fn, image_ids=view.image_ids)   # returns at once
        if i: depth, conf = pipe.collect()                              # result of view i-1
    depth, conf = pipe.collect()

This mirrors what the reference's inference loop does per reference view (mvs/mvs_cas/predict.py:126-133:
`tocuda(sample)`, `model(...)`, `tensor2numpy(outputs)`), but with two input slots, a copy stream and
events instead of a synchronous round trip.  The CNN regulariser between the two kernels is the caller's
(`logits_fn(volume) -> [D,H,W] logits`, run on the compute stream).  CUDA only: there is no CPU path.

Image-keyed residency (`image_ids`).  The reference views of a scene block share their source images: with
V = 5 and a viewpair list of nearest neighbours, consecutive reference views have 4 of their 5 images in common
(mvs/mvs_cas/predict.py:126-133 walks the block in order; adamvs.py:570-574 re-encodes every image of every
view).  When the caller names the images of a view, their feature maps stay on the device in an LRU keyed by
image id and a view uploads only the images that are not there yet -- 41 MB instead of 204 MB at the WHU-OMVS
shape, which is what makes the end-to-end number scale across the 8 GPUs of a box (the host->device copies of
8 ranks share the host's memory system: SCALE_r01 measured 22 GB/s per rank at N = 8 against 34 GB/s at N = 1).
The resident images are kept as channels-last TEXELS (`ImageFeatureLRU(texels=True)`): an image is laid out once, when
it arrives, on the copy stream, and the sweep names its five views by pool slot (`sweep.cost_volume(..., view_slots=)`,
D3dCostVolumeArgs.texel_slots) -- no relayout and no gather copy per reference view.
"""
from __future__ import annotations

import collections
from typing import Callable, Optional, Sequence

import torch

from . import sweep


class ImageFeatureLRU:
    """Per-image feature maps [C,H,W] resident on one device, least recently used first out.  `fetch` returns the
    device tensor of an image, copying it from the host (on the current stream) only on a miss; a buffer is not
    refilled before the stream that last read it has passed the event the reader recorded (`mark_read`)."""

    def __init__(self, capacity: int, channels: int, height: int, width: int, device, texels: bool = False):
        """texels: keep the images as channels-last TEXELS [capacity,H,W,C] -- the layout the sweep kernels gather from --
        laid out once, when an image arrives (`fetch_slot`); a sweep then names its views by pool slot
        (`sweep.cost_volume(lru.texels, ..., view_slots=...)`) and nothing is relaid out or copied per reference view."""
        if capacity < 1:
            raise ValueError("capacity must be >= 1")
        self.capacity = capacity
        self.texels = None
        if texels:
            if capacity * height * width >= 2 ** 31:
                raise ValueError("texel pool of %d x %d x %d exceeds 2^31 texels" % (capacity, height, width))
            self.texels = torch.empty((capacity, height, width, channels), device=device)
            self.pool = None
            self._staging = torch.empty((channels, height, width), device=device)     # one upload at a time, in stream order
        else:
            self.pool = torch.empty((capacity, channels, height, width), device=device)
        self.free = list(range(capacity))
        self.slot_of = collections.OrderedDict()          # image id -> pool slot, oldest first
        self.read_done = [None] * capacity                # event after the last reader of each slot
        self.hits = self.misses = 0
        self.bytes_copied = 0

    def fetch(self, image_id, host_map: torch.Tensor, pinned: Sequence = ()) -> torch.Tensor:
        """`pinned`: ids that must survive this call (the other images of the view being assembled)."""
        slot = self.fetch_slot(image_id, host_map, pinned)
        return self.pool[slot] if self.texels is None else self.texels[slot]

    def fetch_slot(self, image_id, host_map: torch.Tensor, pinned: Sequence = ()) -> int:
        """The pool slot of an image, uploaded (and, in texel mode, laid out channels-last) on the current stream on a miss."""
        slot = self.slot_of.get(image_id)
        if slot is not None:
            self.slot_of.move_to_end(image_id)
            self.hits += 1
            return slot
        self.misses += 1
        if self.free:
            slot = self.free.pop()
        else:
            victim = next((k for k in self.slot_of if k not in pinned), None)
            if victim is None:
                raise RuntimeError("ImageFeatureLRU: capacity %d is smaller than one view's image list" % self.capacity)
            slot = self.slot_of.pop(victim)
            if self.read_done[slot] is not None:          # the sweep that last read this buffer must be past it
                torch.cuda.current_stream().wait_event(self.read_done[slot])
        if self.texels is None:
            self.pool[slot].copy_(host_map, non_blocking=True)
        else:                                             # upload, then the image's ONE relayout, both in stream order
            src = host_map
            if not host_map.is_cuda:
                self._staging.copy_(host_map, non_blocking=True)
                src = self._staging
            sweep.to_texels([src], out=self.texels[slot:slot + 1])
        self.bytes_copied += host_map.numel() * host_map.element_size()
        self.slot_of[image_id] = slot
        return slot

    def mark_read(self, image_ids, event) -> None:
        for k in image_ids:
            slot = self.slot_of.get(k)
            if slot is not None:
                self.read_done[slot] = event


class ViewPipeline:
    def __init__(self, num_views: int, channels: int, height: int, width: int, num_depth: int, device,
                 mode: int = sweep.AGG_VARIANCE, groups: int = 0, conf_mode: int = sweep.CONF_MAX_PROB,
                 per_pixel_hyps: bool = False, variant: int = 0, slots: int = 2, resident_images: int = 0,
                 resident_texels: bool = True):
        """resident_images: capacity of the image-keyed feature LRU (0 = 3 * num_views when a submit names its
        images).  It must hold the images of every view in flight: >= slots * num_views is always enough.
        resident_texels: the LRU keeps the images as channels-last texels, laid out once per image; the sweep names its
        views by pool slot (False: per-image [C,H,W] maps, relaid out into a dense block for every reference view)."""
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("ViewPipeline runs on CUDA only (no CPU fallback)")
        v, c, h, w, d = num_views, channels, height, width, num_depth
        self.shape = (v, c, h, w, d)
        self.mode, self.groups, self.conf_mode, self.variant = mode, groups, conf_mode, variant
        cout = groups if mode == sweep.AGG_GROUP_CORR else (v - 1 if mode == sweep.AGG_PAIR_MEAN else c)
        hyp_shape = (d, h, w) if per_pixel_hyps else (d,)
        self._hyp_elems = d * h * w if per_pixel_hyps else d
        self._resident = resident_images
        self._resident_texels = resident_texels
        self.lru: Optional[ImageFeatureLRU] = None
        with torch.cuda.device(self.dev):
            self.copy_stream = torch.cuda.Stream()
            self.compute_stream = torch.cuda.Stream()
            self.slots = [{
                "feats": None,                                # [V,C,H,W] staging, allocated on first use without image ids
                "proj": torch.empty((v, 4, 4), device=self.dev),
                "hyps": torch.empty(hyp_shape, device=self.dev),
                "out": torch.empty((2, h, w), dtype=torch.float32).pin_memory(),
                "copied": torch.cuda.Event(), "free": torch.cuda.Event(), "done": torch.cuda.Event(),
            } for _ in range(slots)]
            self.texels = None                                # dense [V,H,W,C] block: only views submitted without image ids need it
            self.volume = torch.empty((cout, d, h, w), device=self.dev)
        self._next = 0
        self._pending = collections.deque()
        self.h2d_bytes = 4 * (v * c * h * w + 16 * v + self._hyp_elems)   # per view when every image is uploaded
        self.h2d_bytes_total = 0                                          # what submit() actually copied so far
        self.views_submitted = 0
        self.d2h_bytes = 8 * h * w

    def submit(self, feats, proj: torch.Tensor, hyps: torch.Tensor,
               logits_fn: Callable[[torch.Tensor], torch.Tensor], image_ids: Optional[Sequence] = None) -> None:
        """Enqueue one reference view.  feats [V,C,H,W] (or a list of V [C,H,W] maps), proj [V,4,4], hyps [D] or
        [D,H,W]: pinned host tensors (device tensors are accepted too and copied device-to-device, after the
        stream that produced them).  image_ids: V hashable ids (view 0 = the reference image) -- the feature maps
        of images already resident on the device are not copied again."""
        if len(self._pending) == len(self.slots):
            raise RuntimeError("all slots are in flight: collect() a result first")
        v, c, h, w, _ = self.shape
        if image_ids is not None and len(image_ids) != v:
            raise ValueError("image_ids must name the %d images of the view" % v)
        s = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        producer = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            if any(isinstance(t, torch.Tensor) and t.is_cuda for t in (proj, hyps)) or \
                    any(t.is_cuda for t in (feats if isinstance(feats, (list, tuple)) else [feats])):
                self.copy_stream.wait_stream(producer)        # device inputs: not before whoever wrote them is done
            self.copy_stream.wait_event(s["free"])            # the previous user of this slot has read its inputs
            copied = 4 * (16 * v + self._hyp_elems)
            if image_ids is None:
                if s["feats"] is None:
                    s["feats"] = torch.empty((v, c, h, w), device=self.dev)
                if isinstance(feats, (list, tuple)):
                    for i in range(v):
                        s["feats"][i].copy_(feats[i], non_blocking=True)
                else:
                    s["feats"].copy_(feats, non_blocking=True)
                s["maps"] = list(s["feats"].unbind(0))
                s["ids"] = None
                copied += 4 * v * c * h * w
            else:
                if self.lru is None:
                    cap = self._resident or 3 * v
                    if cap < len(self.slots) * v:
                        raise ValueError("resident_images must be >= slots * num_views = %d" % (len(self.slots) * v))
                    self.lru = ImageFeatureLRU(cap, c, h, w, self.dev, texels=self._resident_texels)
                before = self.lru.bytes_copied
                ids = list(image_ids)
                # the previous views in flight keep their images: an LRU of >= slots * V entries cannot reach them
                if self.lru.texels is not None:
                    s["view_slots"] = [self.lru.fetch_slot(ids[i], feats[i], pinned=ids) for i in range(v)]
                    s["maps"] = None
                else:
                    s["maps"] = [self.lru.fetch(ids[i], feats[i], pinned=ids) for i in range(v)]
                s["ids"] = ids
                copied += self.lru.bytes_copied - before
            s["proj"].copy_(proj, non_blocking=True)
            s["hyps"].copy_(hyps, non_blocking=True)
            # camera geometry of this view (the reference's own torch calls, module.py:528 and :538) also runs on
            # the copy stream, i.e. behind the previous view's sweep; the slot keeps the tensors alive
            s["pose"] = sweep.relative_poses(s["proj"])
            s["rays"] = sweep.rays_for(s["pose"], h, w)
            s["copied"].record()
        self.h2d_bytes_total += copied
        self.views_submitted += 1
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(s["copied"])
            if s["maps"] is None:                             # the view's images are texel-pool slots: nothing to lay out
                sweep.cost_volume(self.lru.texels, s["pose"], s["hyps"], self.mode, groups=self.groups, out=self.volume,
                                  variant=self.variant, rays=s["rays"], view_slots=s["view_slots"])
            else:
                if self.texels is None:
                    self.texels = torch.empty((v, h, w, c), device=self.dev)
                sweep.to_texels(s["maps"], out=self.texels)
            if s["ids"] is not None:                          # (the sweep itself is the pool's last reader)
                read = torch.cuda.Event()
                read.record()
                self.lru.mark_read(s["ids"], read)
            if s["maps"] is not None:
                sweep.cost_volume(self.texels, s["pose"], s["hyps"], self.mode, groups=self.groups, out=self.volume,
                                  variant=self.variant, rays=s["rays"])
            logits = logits_fn(self.volume)
            r = sweep.depth_regress(logits, s["hyps"], conf_mode=self.conf_mode, want_index=False)
            s["free"].record()                                # inputs consumed: the slot may be refilled
            s["out"][0].copy_(r["depth"], non_blocking=True)
            s["out"][1].copy_(r["conf"], non_blocking=True)
            s["done"].record()
        self._pending.append(s)

    def collect(self):
        """Block until the oldest submitted view is finished; returns (depth [H,W], conf [H,W]) as views of
        a pinned host buffer that stays valid until the slot is reused (`slots` submissions later)."""
        if not self._pending:
            raise RuntimeError("nothing submitted")
        s = self._pending.popleft()
        s["done"].synchronize()
        return s["out"][0], s["out"][1]

    def drain(self):
        out = []
        while self._pending:
            out.append(self.collect())
        return out
