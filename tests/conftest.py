import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def synth_image(shape, seed=1234, kind="image"):
    """Synthetic inputs: 'image' = integer-valued 0..255 data like scipy's ascent/lena,
    'smooth' = gaussian noise + low-frequency term."""
    rng = np.random.default_rng(seed)
    if kind == "image":
        return rng.integers(0, 256, size=shape).astype(np.float32)
    shape = tuple(shape)
    x = rng.standard_normal(shape).astype(np.float32) * 50 + 128
    if len(shape) >= 2:
        i = np.arange(shape[-2], dtype=np.float32)[:, None]
        j = np.arange(shape[-1], dtype=np.float32)[None, :]
        x = x + 64 * np.sin(2 * np.pi * i / shape[-2] * 3) * np.cos(2 * np.pi * j / shape[-1] * 5)
    return x.astype(np.float32)


@pytest.fixture(scope="session")
def have_gpu():
    import pypwt_b200
    return pypwt_b200.device_count() > 0
