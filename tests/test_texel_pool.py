"""Host logic of texel_pool.TexelPool (slot bookkeeping, identity / version checks, eviction order) on CPU tensors with the
relayout kernel substituted; the GPU behaviour is tests/test_sweep_gpu.py::test_depthnet_texel_pool_lays_each_image_out_once."""
import gc

import pytest
import torch

from deep3d_aerial_b200 import sweep
from deep3d_aerial_b200.texel_pool import TexelPool


@pytest.fixture
def pool(monkeypatch):
    laid_out = []

    def to_texels(maps, out):
        laid_out.append(maps[0])
        out[0].copy_((maps[0][0] if maps[0].dim() == 4 else maps[0]).permute(1, 2, 0))
        return out

    monkeypatch.setattr(sweep, "to_texels", to_texels)
    monkeypatch.setattr(TexelPool, "_cuda_only", False)
    p = TexelPool(4)
    p.laid_out = laid_out
    return p


def _maps(n, c=3, h=4, w=5):
    g = torch.Generator().manual_seed(n)
    return [torch.randn(1, c, h, w, generator=g) for _ in range(n)]


def test_slots_hold_the_texels_of_their_maps_and_hits_do_not_lay_out_again(pool):
    maps = _maps(3)
    texels, slots = pool.lookup(maps)
    assert texels.shape == (4, 4, 5, 3) and sorted(slots) == sorted(set(slots)) and len(pool.laid_out) == 3
    for m, s in zip(maps, slots):
        assert torch.equal(texels[s], m[0].permute(1, 2, 0))
    again, slots2 = pool.lookup([maps[2], maps[0], maps[1]])
    assert again is texels and slots2 == [slots[2], slots[0], slots[1]] and len(pool.laid_out) == 3
    assert pool.stats()["hits"] == 3 and pool.stats()["misses"] == 3


def test_least_recently_used_map_leaves_first_and_the_views_own_maps_never_evict_each_other(pool):
    maps = _maps(6)
    _, s01 = pool.lookup(maps[0:2])                      # slots: m0 m1
    _, s23 = pool.lookup(maps[2:4])                      # full: m0 m1 m2 m3
    pool.lookup([maps[0]])                               # m0 is now the most recent
    _, s45 = pool.lookup([maps[4], maps[5]])             # evicts m1 and m2 (oldest first), never m4 for m5
    assert sorted(s45) == sorted([s01[1], s23[0]])
    n = len(pool.laid_out)
    _, s = pool.lookup([maps[0], maps[3]])               # both survived
    assert s == [s01[0], s23[1]] and len(pool.laid_out) == n
    assert pool.lookup(_maps(5)) is None                 # five distinct maps do not fit four slots: the caller goes dense


def test_a_modified_or_replaced_tensor_is_laid_out_again(pool):
    maps = _maps(2)
    texels, slots = pool.lookup(maps)
    maps[0].add_(1.0)                                    # in place: the version counter moves
    _, slots2 = pool.lookup(maps)
    assert slots2 == slots and len(pool.laid_out) == 3
    assert torch.equal(texels[slots[0]], maps[0][0].permute(1, 2, 0))
    # a dead tensor's id may be handed to a new one: the weak reference, not the id, decides
    dead_id = id(maps[1])
    del maps[1]
    pool.laid_out.clear()            # (the recorder's own reference)
    gc.collect()
    for _ in range(64):                                  # provoke id reuse
        t = torch.randn(1, 3, 4, 5)
        if id(t) == dead_id:
            break
    before = pool.stats()["misses"]
    got = pool.lookup([t])
    assert got is not None and pool.stats()["misses"] == before + 1
    assert torch.equal(got[0][got[1][0]], t[0].permute(1, 2, 0))


def test_shapes_get_their_own_pool_and_mixed_views_go_dense(pool):
    a, b = _maps(2), _maps(2, c=2, h=8, w=10)
    ta, _ = pool.lookup(a)
    tb, _ = pool.lookup(b)
    assert ta.shape == (4, 4, 5, 3) and tb.shape == (4, 8, 10, 2) and len(pool.pools) == 2
    assert pool.lookup([a[0], b[0]]) is None
    assert pool.lookup([a[0], a[1].double()]) is None
