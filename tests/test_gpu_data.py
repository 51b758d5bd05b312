import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_prefetcher_yields_identical_batches_in_order():
    from visuelle2_multimodal_fusion_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [((torch.randn(4, 3, generator=g), torch.randint(0, 9, (4,), generator=g)), torch.randn(4, 3, 8, 8, generator=g))
            for _ in range(5)]
    seen = []
    for (a, b), im in DevicePrefetcher(host, "cuda"):
        assert a.is_cuda and b.is_cuda and im.is_cuda
        torch.cuda._sleep(2_000_000)              # consumer busy while the next copy is in flight
        seen.append((a.cpu(), b.cpu(), im.cpu()))
    assert len(seen) == 5
    for ((a, b), im), (a2, b2, im2) in zip(host, seen):
        assert torch.equal(a, a2) and torch.equal(b, b2) and torch.equal(im, im2)
