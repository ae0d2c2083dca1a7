"""-m gpu: the steps either side of the network through the C ABI.

input edge       uint8 BGR image -> z-normalised fp32 CHW (test_mc3serousv5.py:100-127, DataLoader.py:661-671): against the
                 golden outputs of the reference's own `preprocess` (oracle/make_golden_edge.py) and the numpy oracle.
inference heads  fused OutConv + softmax + argmax + uint8 (test_mc3serousv5.py:879-887) and OutConv + relu + /200 + counts
                 (:961-974): bit-exact against the unfused kernels of this library and the oracle on the same logits.
"""
import numpy as np
import pytest
import torch

from gpu_util import to_nhwc_bf16
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    """float32 equality where nan == nan (a constant channel has std 0 -> nan, like the reference)."""
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)) or bool(
        ((a == b) | (a.isnan() & b.isnan())).all())


def test_preprocess_matches_reference_golden(golden):
    import unet_torch_b200 as U

    g = golden("ref_edge.pt")
    for name, c in g.items():
        if name.startswith("crop_"):   # preprocessCrop (test.py:91-126): padded with 255 to a multiple of the crop size
            got = U.preprocess_crop(c["img"].numpy(), c["crop"]).cpu()
            assert got.shape == c["out"].shape and _same_bits(got, c["out"]), name
            continue
        got = U.preprocess(c["img"].numpy(), c["img"].shape[:2]).cpu()
        assert got.shape == c["out"].shape and got.dtype == torch.float32, name
        assert _same_bits(got, c["out"]), f"{name}: max diff {float((got - c['out']).abs().nan_to_num().max()):.3e}"


@pytest.mark.parametrize("n,h,w,c", [(16, 512, 512, 3), (3, 250, 130, 3), (2, 1024, 1024, 1), (4, 64, 64, 4), (1, 16, 16, 2)])
def test_znorm_batch_against_oracle(n, h, w, c):
    """Full-size batches (config 2's 16 x 512^2 x 3): every image and channel normalised on its own."""
    import unet_torch_b200 as U
    from unet_torch_b200 import ops

    rng = np.random.default_rng(n * 1000 + h)
    imgs = np.clip(rng.normal(120, 40, size=(n, h, w, c)) + rng.normal(0, 25, size=(n, 1, 1, c)), 0, 255).astype(np.uint8)
    got = ops.znorm_to_chw(torch.from_numpy(imgs).cuda(), reverse_channels=True).cpu()
    for i in range(n):
        img = imgs[i] if c > 1 else imgs[i, :, :, 0]
        want = O.preprocess(img)[0]
        # same formula in fp64, exact integer moments here vs numpy's rounded two-pass variance: the float32 results can
        # differ by at most one ulp where the fp64 quotient sits on a rounding boundary (never observed)
        assert torch.allclose(got[i], want, rtol=2e-7, atol=0), (i, float((got[i] - want).abs().max()))
        frac_equal = float((got[i] == want).float().mean())
        assert frac_equal > 0.9999, frac_equal
    # property at full size: each output plane has mean 0 and population std 1
    m = got.double().mean(dim=(2, 3))
    s = got.double().var(dim=(2, 3), unbiased=False).sqrt()
    assert float(m.abs().max()) < 1e-6 and float((s - 1).abs().max()) < 1e-6
    # grey images keep their single channel; U.preprocess takes [H,W] too
    if c == 1:
        one = U.preprocess(imgs[0, :, :, 0]).cpu()
        assert torch.equal(one[0], got[0])


def test_znorm_rejects_bad_input():
    from unet_torch_b200 import ops
    import unet_torch_b200 as U

    with pytest.raises(TypeError):
        ops.znorm_to_chw(torch.zeros(1, 8, 8, 3, device="cuda"))          # not uint8
    with pytest.raises(RuntimeError):
        ops.znorm_to_chw(torch.zeros(1, 8, 8, 5, dtype=torch.uint8, device="cuda"))  # C > 4
    with pytest.raises(ValueError):
        U.preprocess(np.zeros((8, 8, 3), np.uint8), (16, 16))            # resize is the caller's job


@pytest.mark.parametrize("ncls,n,h,w", [(5, 2, 64, 96), (2, 1, 48, 48), (3, 3, 32, 80), (8, 1, 16, 16)])
def test_fused_heads_bit_exact(ncls, n, h, w):
    from unet_torch_b200 import ops

    g = torch.Generator().manual_seed(ncls * 7 + n)
    a = to_nhwc_bf16(torch.randn(n, 64, h, w, generator=g))
    wt = (torch.randn(ncls, 64, 1, 1, generator=g) * 0.2).cuda()
    wt[1] = wt[0]  # an exact tie between classes 0 and 1 wherever they win: the first must be reported
    b = (torch.randn(ncls, generator=g) * 0.1).cuda()
    b[1] = b[0]
    logits = torch.empty((n, ncls, h, w), dtype=torch.float32, device="cuda")
    ops.head_fprop(a, wt, b, logits)
    mask = ops.head_mask(a, wt, b)
    assert mask.dtype == torch.uint8 and mask.shape == (n, h, w)
    assert torch.equal(mask.long(), ops.softmax_argmax(logits))              # == the two-kernel path
    assert torch.equal(mask.cpu(), O.mask_uint8(logits.cpu()))               # == the oracle on the same logits
    assert not bool((mask == 1).any())
    dens, counts = ops.head_density(a, wt, b, 200.0)
    want, want_counts = O.density_maps(logits.cpu(), 200.0)
    assert torch.equal(dens.cpu(), want)
    assert torch.allclose(counts.cpu(), want_counts, rtol=1e-12, atol=1e-12)
    sm = ops.head_sigmoid_mask(a, wt, b, 0.5)
    assert torch.equal(sm, (torch.sigmoid(logits)[:, 0] >= 0.5).to(torch.uint8))   # torch's own CUDA sigmoid on the same logits
    assert torch.equal(sm.cpu(), O.sigmoid_mask(logits.cpu()))
    assert 0.2 < float(sm.float().mean()) < 0.8
    relu_only, none = ops.head_density(a, wt, b, 1.0, with_counts=False)
    assert none is None and torch.equal(relu_only.cpu(), torch.relu(logits.cpu()))


def test_model_predict_equals_unfused_path():
    """net.predict(x) == np.uint8(argmax(softmax(net(x)))) and net.predict_density(x) == relu(net(x)) / 200, eval mode."""
    import unet_torch_b200 as U

    torch.manual_seed(5)
    net = U.UNet(3, 5).cuda().eval()
    x = torch.randn(2, 3, 64, 96, device="cuda")
    with torch.no_grad():
        logits = net(x)
        mask = net.predict(x)
        dens, counts = net.predict_density(x)
    assert torch.equal(mask.long(), U.predict_mask(logits))
    assert torch.equal(net.predict_binary(x), (torch.sigmoid(logits)[:, 0] >= 0.5).to(torch.uint8))
    # the reference divides on the CPU in numpy (true fp32 division; torch's CUDA div-by-scalar multiplies by 1/200)
    assert np.array_equal(dens.cpu().numpy(), torch.relu(logits).cpu().numpy() / 200)
    assert torch.allclose(counts, dens.double().sum(dim=(2, 3)), rtol=1e-12)
    # end to end from the uint8 image, the way test_mc3serousv5.py:876-887 chains the calls
    img = torch.randint(0, 256, (64, 96, 3), dtype=torch.uint8)
    m2 = net.predict(U.preprocess(img.numpy(), (64, 96)))
    with torch.no_grad():
        ref = O.mask_uint8(net(O.preprocess(img.numpy()).cuda()).cpu())
    assert torch.equal(m2.cpu(), ref)


def test_predict_tiled_equals_the_reference_crop_loop():
    """test_single_crop (test.py:420-447): every crop through the network one at a time, sigmoid >= 0.5, stitched - against
    the batched, fused-head `predict_tiled`; the per-crop logits do not depend on which batch a crop is in."""
    import unet_torch_b200 as U

    torch.manual_seed(9)
    net = U.UNet(3, 1).cuda().eval()
    img = torch.randint(0, 256, (100, 150, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    crop = 64
    got = U.predict_tiled(net, img.numpy(), crop, head="sigmoid", max_batch=4)
    assert got.shape == (128, 192) and got.dtype == torch.uint8
    x = O.preprocess_crop(img.numpy(), crop)
    with torch.no_grad():
        want = O.tiled_masks(lambda t: net(t.cuda()).cpu(), x, crop, O.sigmoid_mask)
    agree = float((got.cpu() == want).float().mean())
    print(f"predict_tiled vs crop-by-crop loop: agreement {agree:.6f}")
    assert agree > 0.9999
    assert 0.05 < float(got.float().mean()) < 0.95
    multi = U.UNet(3, 3).cuda().eval()
    m = U.predict_tiled(multi, img.numpy(), crop, head="mask")
    with torch.no_grad():
        want_m = O.tiled_masks(lambda t: multi(t.cuda()).cpu(), x, crop, O.mask_uint8)
    assert float((m.cpu() == want_m).float().mean()) > 0.9999
    with pytest.raises(ValueError):
        U.predict_tiled(net, img.numpy(), 50)
