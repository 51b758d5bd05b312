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


@pytest.mark.parametrize("shape", [(3, 299, 299, 3), (2, 17, 5, 3), (1, 1, 1, 3), (4, 8, 8, 1)])
def test_device_image_transform_equals_totensor_normalize(shape):
    """csrc/image_prep.cu against the reference's per-item transform (dataset_fusion.py:50-65) restated with
    torchvision: Normalize(mean, std)(ToTensor(PIL image)); fp32 output bit-identical, bf16 = one rounding of it."""
    import numpy as np
    from PIL import Image
    from torchvision.transforms import Compose, Normalize, ToTensor
    from visuelle2_multimodal_fusion_b200.data import IMAGENET_MEAN, IMAGENET_STD, normalize_uint8_images
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g)
    u8.view(-1)[:2] = torch.tensor([0, 255], dtype=torch.uint8)[: u8.numel()]
    C = shape[-1]
    mean, std = IMAGENET_MEAN[:C], IMAGENET_STD[:C]
    tf = Compose([ToTensor(), Normalize(mean=list(mean), std=list(std))])
    ref = torch.stack([tf(Image.fromarray(im.numpy() if C == 3 else im.numpy()[:, :, 0])) for im in u8])
    out32 = normalize_uint8_images(u8.cuda(), torch.float32, mean, std)
    assert out32.shape == ref.shape
    if C > 1 and shape[1] * shape[2] > 1:
        assert out32.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out32.cpu(), ref)
    out16 = normalize_uint8_images(u8.cuda(), torch.bfloat16, mean, std)
    assert torch.equal(out16.cpu(), ref.to(torch.bfloat16))


def test_model_accepts_device_normalized_images():
    """The drop-in forward takes the bf16 channels_last batch the device transform produces and gives the same forecast (to the noise of the bf16 trunk)
    as with the fp32 batch of the reference's DataLoader (same bf16 trunk arithmetic from the first convolution on)."""
    import bench
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.data import IMAGENET_MEAN, IMAGENET_STD, normalize_uint8_images
    model = bench._build_model("rnn210", "cuda:0", "bf16").eval()
    data, _ = synth.make_batch(2, out_len=10, seed=4, feat_hw=10)
    data = tuple(t.cuda() for t in data)
    u8 = torch.randint(0, 256, (2, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).cuda()
    mean = torch.tensor(IMAGENET_MEAN, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device="cuda").view(1, 3, 1, 1)
    ref_images = (u8.permute(0, 3, 1, 2).float().div(255) - mean) / std          # what the reference's loader yields
    with torch.no_grad():
        a = model(*data, ref_images)[0]
        b = model(*data, normalize_uint8_images(u8))[0]
    # same arithmetic from the first convolution on; cuDNN may still pick another algorithm for the differently
    # provenanced input, so the forecasts agree to bf16-trunk noise rather than bit for bit
    assert torch.allclose(a, b, rtol=2e-3, atol=1e-5), float((a - b).abs().max())
