"""Hand crop: the oracle's loop restatement vs the host function the data-set shim uses (CPU), and
the on-device kernel vs both (GPU).  The resampling parity is UNPINNED against the reference itself (kornia is not installed, see
oracle/crop_restated.py); box arithmetic is restated from plain torch code."""
import numpy as np
import pytest
import torch

from helpers import rel


def _case(n, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    frames = torch.rand(n, 3, H, W, generator=g)
    cx, cy = torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H
    hw, hh = 5 + torch.rand(n, generator=g) * W / 3, 5 + torch.rand(n, generator=g) * H / 3
    tight = torch.stack([cx - hw, cy - hh, cx + hw, cy + hh], dim=-1)          # some boxes stick out of the frame
    return frames, tight


def test_host_crop_matches_oracle_loops():
    from cs_vit.utils.img import crop_tensor_with_square_box, expand_bbox_square
    from oracle import crop_restated as oc
    frames, tight = _case(3, 40, 56, 1)
    got, scales, sq = crop_tensor_with_square_box(frames, tight, 1.25, 16)
    want, wscales, wsq = oc.crop_tensor_with_square_box(frames.numpy(), tight.numpy(), 1.25, 16)
    assert np.allclose(sq.numpy(), wsq, atol=1e-4) and np.allclose(scales.numpy(), wscales, rtol=1e-6)
    assert rel(got, torch.from_numpy(want)) < 1e-5
    assert torch.allclose(expand_bbox_square(tight, 1.25), sq, atol=1e-4)      # the two box helpers of the reference agree
    # a box that covers the frame exactly with size == frame size is the identity
    ident, _, _ = crop_tensor_with_square_box(frames[:, :, :40, :40], torch.tensor([[0.0, 0.0, 39.0, 39.0]] * 3), 1.0, 40)
    assert torch.allclose(ident, frames[:, :, :40, :40], atol=1e-4)      # (grid_sample normalises and de-normalises the coordinates in fp32)


@pytest.mark.gpu
@pytest.mark.parametrize("u8", [False, True])
def test_device_crop_matches_oracle_and_host(u8):
    from cs_vit import ops
    from cs_vit.utils.img import crop_tensor_with_square_box
    from oracle import crop_restated as oc
    frames, tight = _case(4, 48, 64, 2)
    if u8:
        bytes_ = (frames * 255).round().to(torch.uint8)
        frames = bytes_.float() / 255
        dev_frames = bytes_.permute(0, 2, 3, 1).contiguous().cuda()
    else:
        dev_frames = frames.cuda()
    got, sq = ops.crop_resize(dev_frames, tight.cuda(), 20, expansion_ratio=1.25)
    want, _, wsq = oc.crop_tensor_with_square_box(frames.numpy(), tight.numpy(), 1.25, 20)
    assert np.allclose(sq.cpu().numpy(), wsq, atol=1e-4)
    assert rel(got, torch.from_numpy(want)) < 1e-5, rel(got, torch.from_numpy(want))
    # full size, against the host function (grid_sample), DexYCB-sized frames; boxes given as final boxes (no expansion)
    frames, tight = _case(6, 480, 640, 3)
    host, _, hsq = crop_tensor_with_square_box(frames, tight, 1.25, 224)
    dev, dsq = ops.crop_resize(frames.cuda(), tight.cuda(), 224, expansion_ratio=1.25)
    assert torch.allclose(dsq.cpu(), hsq, atol=1e-3)
    assert rel(dev, host) < 1e-4, rel(dev, host)
    again, same = ops.crop_resize(frames.cuda(), hsq.cuda(), 224)
    assert torch.allclose(same.cpu(), hsq) and rel(again, host) < 1e-4
    with pytest.raises(ValueError):
        ops.crop_resize(frames.cuda(), tight[:2].cuda(), 224)


@pytest.mark.gpu
def test_crops_feed_predict_batch():
    """frames + tight boxes -> crop_resize -> predict_batch equals predict_batch on the host-cropped patches."""
    from helpers import build_product
    from cs_vit import ops
    from cs_vit.utils.img import crop_tensor_with_square_box
    model, inputs, _, _ = build_product("swint_encoder_patch_spatial", "fp32")
    model = model.cuda()
    frames, tight = _case(2, 480, 640, 4)
    host, _, hsq = crop_tensor_with_square_box(frames, tight, 1.25, 224)
    dev, dsq = ops.crop_resize(frames.cuda(), tight.cuda(), 224, expansion_ratio=1.25)
    kw = {k: inputs[k][:2].cuda() for k in ("timestamp", "focal", "princpt")}
    with torch.no_grad():
        a = model.predict_batch(dev[:, None], dsq[:, None], kw["timestamp"], kw["focal"], kw["princpt"])["joint_cam"]
        b = model.predict_batch(host[:, None].cuda(), hsq[:, None].cuda(), kw["timestamp"], kw["focal"], kw["princpt"])["joint_cam"]
    assert rel(a, b) < 1e-3
