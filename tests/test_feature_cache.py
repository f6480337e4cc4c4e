"""FeatureNet sharing across reference views.  Host-side logic, runs on CPU."""
import torch
import torch.nn as nn

from deep3d_aerial_b200.feature_cache import FeatureCache


class ToyFeatureNet(nn.Module):
    out_channels = [8, 4]

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 8, 3, padding=1)
        self.calls = 0

    def forward(self, x):
        self.calls += 1
        y = self.conv(x)
        return {"stage1": torch.nn.functional.avg_pool2d(y, 2), "stage2": y[:, :4]}


class ToyModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.feature = ToyFeatureNet()

    def forward(self, imgs):
        feats = [self.feature(imgs[:, v]) for v in range(imgs.size(1))]          # adamvs.py:570-574
        return sum(f["stage1"].mean() + f["stage2"].mean() for f in feats)


def _images(n):
    g = torch.Generator().manual_seed(0)
    return {i: torch.randn(1, 3, 16, 20, generator=g) for i in range(n)}


def test_cached_pyramids_are_reused_and_identical():
    torch.manual_seed(0)
    model = ToyModel().eval()
    keys_before = sorted(model.state_dict())
    imgs = _images(6)
    view_lists = [[0, 1, 2], [1, 0, 2], [2, 1, 3], [3, 2, 4], [4, 3, 5]]
    with torch.no_grad():
        plain = [model(torch.stack([imgs[i][0] for i in ids], 0).unsqueeze(0)) for ids in view_lists]
    assert model.feature.calls == 15
    cache = FeatureCache(capacity=4).attach(model.feature)
    assert sorted(model.state_dict()) == keys_before and model.feature.out_channels == [8, 4]
    model.feature.calls = 0
    with torch.no_grad():
        cached = []
        for ids in view_lists:
            with cache.views(ids):
                cached.append(model(torch.stack([imgs[i][0] for i in ids], 0).unsqueeze(0)))
    assert model.feature.calls == 6 and cache.stats()["hits"] == 9 and cache.stats()["misses"] == 6
    assert all(torch.equal(a, b) for a, b in zip(plain, cached))                 # bit-identical outputs
    assert cache.stats()["images"] == 4 and cache.stats()["bytes"] == 4 * (8 * 8 * 10 + 4 * 16 * 20) * 4
    cache.detach()
    with torch.no_grad():
        model(torch.stack([imgs[i][0] for i in (0, 1, 2)], 0).unsqueeze(0))
    assert model.feature.calls == 9                                              # pass-through again


def test_lru_eviction_and_bypass():
    model = ToyModel().eval()
    cache = FeatureCache(capacity=2).attach(model.feature)
    imgs = _images(3)
    with torch.no_grad():
        for i in (0, 1, 2, 0):                      # 0 is evicted by 2, then recomputed
            with cache.views([i]):
                model.feature(imgs[i])
        assert model.feature.calls == 4 and cache.stats()["hits"] == 0
        with cache.views([0]):
            model.feature(imgs[0])
        assert cache.stats()["hits"] == 1
        model.feature(imgs[0])                      # outside views(): not cached, not counted
        with cache.views([7]):
            model.feature(torch.cat([imgs[0], imgs[1]], 0))      # batch of two: bypass
        with cache.views([0]):
            model.feature(imgs[0][:, :, :8])        # another shape under the same id: its own entry
    assert model.feature.calls == 7 and cache.stats()["hits"] == 1
    capped = FeatureCache(capacity=8, max_bytes=1)
    capped.attach(ToyModel().feature)
    with torch.no_grad():
        for i in range(3):
            with capped.views([i]):
                capped._net(imgs[i])
    assert capped.stats()["images"] == 1
