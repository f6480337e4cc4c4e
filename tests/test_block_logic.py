"""Host logic of the scene-block workloads (no GPU): the viewpair windows bench.py walks (block_image_ids) and the deal of
a block to ranks (shard.partition, 'contiguous' -- what keeps neighbouring views, and so their shared images, on one GPU)."""
import importlib.util
import os

import pytest

from deep3d_aerial_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


@pytest.mark.parametrize("n_block,v", [(64, 5), (16, 5), (5, 5), (9, 3)])
def test_viewpair_windows(n_block, v):
    prev = None
    for i in range(n_block):
        ids = bench.block_image_ids(i, n_block, v)
        assert ids[0] == i and len(ids) == v and len(set(ids)) == v          # reference image first, V distinct images
        assert all(0 <= j < n_block for j in ids)
        assert max(ids) - min(ids) == v - 1                                  # a window of V consecutive images
        if prev is not None:
            assert len(set(ids) & set(prev)) >= v - 1                        # at most one new image per view
        prev = ids
    # every image of the block is some view's reference image and appears in at most V windows' worth of uploads
    seen = set()
    for i in range(n_block):
        seen |= set(bench.block_image_ids(i, n_block, v))
    assert seen == set(range(n_block))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_contiguous_deal_uploads_few_images_per_rank(world):
    """A rank that owns a contiguous share of a 64-view block touches (share + V - 1) images at most: with image residency
    that is ~1.5 uploads per view at 8 ranks, against 5 per view without."""
    n_block, v = 64, 5
    owned = []
    for rank in range(world):
        mine = shard.partition(list(range(n_block)), world, rank, "contiguous")
        owned += mine
        images = set()
        for i in mine:
            images |= set(bench.block_image_ids(i, n_block, v))
        assert len(images) <= len(mine) + v - 1
    assert sorted(owned) == list(range(n_block))
