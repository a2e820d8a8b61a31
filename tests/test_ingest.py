"""CPU: the resampling oracle (oracle/cv2_resize_oracle.py) against OpenCV itself -- the cv2 installed in the image and
the stored outputs of tests/golden/make_golden_resize.py -- and the product's host-side tap tables against the oracle."""
import os

import numpy as np
import pytest

from oracle import cv2_resize_oracle as O
from tdnet_b200 import ingest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_cases.npz")
SHAPES = [((1024, 2048), (769, 1537)), ((1024, 2048), (512, 1024)), ((720, 960), (769, 1537)), ((37, 53), (97, 161)),
          ((97, 161), (37, 53)), ((480, 640), (1024, 2048)), ((5, 7), (64, 96)), ((333, 777), (111, 259)),
          ((64, 64), (64, 64)), ((2, 2), (9, 9))]


def test_oracle_matches_stored_cv2_outputs():
    g = np.load(GOLDEN)
    for i in range(5):
        out = g[f"lin{i}_out"]
        assert np.array_equal(O.resize_linear_u8(g[f"lin{i}_in"], out.shape[1], out.shape[0]), out), i
    big = np.random.default_rng(int(g["big_seed"])).integers(0, 256, (1024, 2048, 3), dtype=np.uint8)
    out = O.resize_linear_u8(big, 1537, 769)                      # the reference's deployment shape (test.py:36)
    assert np.array_equal(out[[0, 1, 384, 767, 768]], g["big_rows"])
    assert np.array_equal(out[:, [0, 1, 768, 1535, 1536]], g["big_cols"])
    for i in range(3):
        out = g[f"nn{i}_out"]
        assert np.array_equal(O.resize_nearest(g[f"nn{i}_in"], out.shape[1], out.shape[0]), out), i


@pytest.mark.parametrize("src,dst", SHAPES)
def test_oracle_matches_installed_cv2(src, dst):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(src[0] * 7 + dst[1])
    img = rng.integers(0, 256, (src[0], src[1], 3), dtype=np.uint8)
    assert np.array_equal(O.resize_linear_u8(img, dst[1], dst[0]), cv2.resize(img, (dst[1], dst[0])))
    lab = rng.integers(0, 19, src, dtype=np.int8)                  # test.py:63-64: int8 labels, INTER_NEAREST
    assert np.array_equal(O.resize_nearest(lab, dst[1], dst[0]),
                          cv2.resize(lab, (dst[1], dst[0]), interpolation=cv2.INTER_NEAREST))


@pytest.mark.parametrize("src,dst", [(1024, 769), (2048, 1537), (37, 97), (97, 37), (5, 64), (100, 100), (2, 9), (1080, 1024)])
def test_product_tap_tables_equal_the_oracle(src, dst):
    tx, ty = ingest.linear_taps(src, dst, "x"), ingest.linear_taps(src, dst, "y")
    for i, (a, b) in enumerate(zip(O.linear_tables_x(src, dst), O.linear_tables_y(src, dst))):
        assert np.array_equal(tx[:, i], a) and np.array_equal(ty[:, i], b), i
    assert tx.dtype == np.int32 and tx.shape == (dst, 4)
    assert (tx[:, 2] + tx[:, 3] == 2048).all() and (ty[:, 2] + ty[:, 3] == 2048).all()
    assert np.array_equal(ingest.nearest_coords(src, dst), O.nearest_offsets(src, dst))
    with pytest.raises(ValueError):
        ingest.linear_taps(0, 4, "x")


def test_vertical_border_rows_are_not_plain_copies():
    """The quirk the oracle pins: OpenCV resets the fraction for out-of-range COLUMNS but not for ROWS, so the first /
    last output rows of an up-scaled image blend the border row with itself using both weights (two >> 16 roundings)."""
    ty = ingest.linear_taps(37, 97, "y")
    assert ty[0, 0] == ty[0, 1] == 0 and ty[0, 2] != 2048          # both taps read row 0, weights (w0, w1) kept
    tx = ingest.linear_taps(37, 97, "x")
    assert tx[0, 0] == 0 and tx[0, 2] == 2048 and tx[0, 3] == 0     # columns: border pixel copied


def test_oracle_matches_installed_cv2_on_random_shapes():
    """Sixty random (source, destination) shape pairs between 2 and 90 pixels per side, up- and down-scaling mixed per
    axis: the restatement must agree with cv2 bit for bit on every one of them."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(99)
    for _ in range(60):
        h, w, H, W = (int(v) for v in rng.integers(2, 91, 4))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(O.resize_linear_u8(img, W, H), cv2.resize(img, (W, H))), (h, w, H, W)
        lab = rng.integers(0, 19, (h, w), dtype=np.int8)
        assert np.array_equal(O.resize_nearest(lab, W, H), cv2.resize(lab, (W, H), interpolation=cv2.INTER_NEAREST))


def test_header_is_valid_c(tmp_path):
    """include/tdnet_b200.h is a C header (the boundary is a C ABI): it must compile as C99, not only as C++."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "tdnet_b200.h"\n'
                   "int main(void) { tdn_tensor t; tdn_attention_desc d; (void)t; (void)d; return TDN_ABI_VERSION - 2; }\n")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           "-c", str(src), "-o", str(tmp_path / "abi.o")])
