"""GPU: device-side frame ingest (cv2.resize of the uint8 frame, Testing/dataloader.py:63) and the quarter-size label
preview (Testing/test.py:61-64), bit-exact against the resampling oracle (itself pinned on cv2, tests/test_ingest.py)."""
import numpy as np
import pytest
import torch

from common import make_weights
from oracle import cv2_resize_oracle as O
from tdnet_b200.synth import synth_clip

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("src,dst,n", [((1024, 2048), (769, 1537), 1), ((37, 53), (97, 161), 2), ((97, 161), (37, 53), 1),
                                       ((480, 640), (1024, 2048), 1), ((5, 7), (64, 96), 3), ((128, 256), (128, 256), 1)])
def test_resize_linear_u8_is_bit_exact(src, dst, n):
    from tdnet_b200.ingest import FrameResizer
    rng = np.random.default_rng(src[0] + dst[1] + n)
    frames = rng.integers(0, 256, (n, src[0], src[1], 3), dtype=np.uint8)
    resize = FrameResizer(src, dst, "cuda:0")
    got = resize(torch.from_numpy(frames).cuda()).cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], O.resize_linear_u8(frames[i], dst[1], dst[0])), i
    try:
        import cv2
        assert np.array_equal(got[0], cv2.resize(frames[0], (dst[1], dst[0])))
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="uint8 HWC"):
        resize(torch.zeros(n, src[0] + 1, src[1], 3, dtype=torch.uint8, device="cuda"))


def test_device_ingest_chain_equals_the_dataloader_pipeline():
    """uint8 camera frame -> resize -> (x/255 - mean)/std -> NCHW -> model (Testing/dataloader.py:63-71, test.py:53) on the
    host with fp64 normalisation exactly as the dataloader, against FrameResizer -> forward_u8 on the device."""
    from tdnet_b200.ingest import FrameResizer
    from tdnet_b200.model import td4_psp18
    src, (H, W) = (120, 200), (97, 161)
    sd = make_weights("td4_psp18", "resnet18", 13, 21)
    nets = []
    for _ in range(2):
        net = td4_psp18.td4_psp18(nclass=19, path_num=4, backbone="resnet18", ln_shape=(13, 21))
        net.load_state_dict(sd, strict=True)
        nets.append(net.eval().to("cuda:0"))
    resize = FrameResizer(src, (H, W), "cuda:0")
    mean, std = np.array([.485, .456, .406]), np.array([.229, .224, .225])
    rng = np.random.default_rng(3)
    for i in range(6):
        frame = rng.integers(0, 256, (src[0], src[1], 3), dtype=np.uint8)
        img = O.resize_linear_u8(frame, W, H) / 255.0
        img = ((img - mean) / std).transpose(2, 0, 1)[np.newaxis, :]
        ref = nets[0](torch.from_numpy(img).float().cuda(), pos_id=i % 4)
        out = nets[1].forward_u8(resize(torch.from_numpy(frame[None]).cuda()), pos_id=i % 4)
        assert torch.equal(ref, out), i


@pytest.mark.parametrize("arch", ["td4_psp18", "td2_fa", "pspnet"])
def test_forward_preview_equals_nearest_resize_of_the_label_map(arch):
    """forward_preview == cv2.resize(output.max(1)[1].astype(int8), (W//4, H//4), INTER_NEAREST) (test.py:61-64)."""
    H, W = 128, 200                                               # W//4 = 50: nearest picks columns 0, 4, 8, ...
    if arch == "td2_fa":
        from common import make_fanet_oracle
        from tdnet_b200.model import td2_fa
        _, sd = make_fanet_oracle("resnet18", H, W)
        build = lambda: td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=2, ln_shape=(16, 25))  # noqa: E731
    elif arch == "pspnet":
        from tdnet_b200.model import pspnet
        from tdnet_b200.synth import synth_state_dict
        build = lambda: pspnet.pspnet(nclass=19, backbone="resnet18")                                # noqa: E731
        sd = synth_state_dict(build().state_dict(), seed=0)
    else:
        from tdnet_b200.model import td4_psp18
        sd = make_weights("td4_psp18", "resnet18", 16, 25)
        build = lambda: td4_psp18.td4_psp18(nclass=19, path_num=4, backbone="resnet18", ln_shape=(16, 25))  # noqa: E731
    a, b = build(), build()
    for net in (a, b):
        net.load_state_dict(sd, strict=True)
        net.eval().to("cuda:0")
    frames = [f.cuda() for f in synth_clip(7, H, W, clip_id=4)]
    paths = a.PATHS
    for i in range(6):
        x = [frames[i], frames[i + 1]] if arch == "td2_fa" else frames[i]
        full = a.forward_labels(x, pos_id=i % paths).cpu().numpy()
        prev = b.forward_preview(x, pos_id=i % paths).cpu().numpy()
        assert prev.shape == (1, H // 4, W // 4) and prev.dtype == np.uint8
        want = O.resize_nearest(full[0].astype(np.int8), W // 4, H // 4)
        assert np.array_equal(prev[0].astype(np.int8), want), i
    odd = b.forward_preview([frames[0], frames[1]] if arch == "td2_fa" else frames[0], pos_id=0, out_hw=(37, 53))
    assert odd.shape == (1, 37, 53)
