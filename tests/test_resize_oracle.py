"""CPU: the INTER_AREA restatement against the installed OpenCV (third-party dependency of the reference)."""
import numpy as np
import pytest

from oracle import resize_oracle as R

cv2 = pytest.importorskip("cv2")

SIZES = [(448, 448), (672, 448), (896, 224), (224, 224), (300, 300), (500, 333), (640, 905), (230, 225),
         (159, 159), (100, 180), (180, 300), (300, 180), (223, 223), (225, 225), (1, 1), (2, 500), (449, 449)]


@pytest.mark.parametrize("hw", SIZES)
def test_resize_matches_cv2(hw):
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    src = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    ref = cv2.resize(src, (224, 224), interpolation=cv2.INTER_AREA)
    got = R.resize_area_u8(src)
    d = np.abs(ref.astype(int) - got.astype(int))
    assert d.max() <= 1
    assert (d > 0).mean() < 1e-3


def test_crop_to_model_input_swaps_channels():
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (300, 260, 3), dtype=np.uint8)
    ref = cv2.cvtColor(cv2.resize(src, (224, 224), interpolation=cv2.INTER_AREA), cv2.COLOR_RGB2BGR)
    np.testing.assert_array_equal(R.crop_to_model_input(src), ref)
