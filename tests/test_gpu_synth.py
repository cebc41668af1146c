"""The benchmark's pinned counter-based frame generator: the device kernel (ef_synth_frames_async, product arm of
bench.py) and the oracle's synth_frame (CPU arm) produce the same pixels for the same (seed, frame)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h", [(3840, 2160), (641, 77), (33, 5)])
def test_device_generator_equals_oracle(oracle, w, h):
    import torch
    import efb200
    seed = 0xEFB20004
    g = efb200.synth_frames(3, h, w, seed, first_frame=5).cpu().numpy()
    for i in range(3):
        assert np.array_equal(g[i], oracle.synth_frame(seed, 5 + i, w, h)), f"frame {5 + i} differs"
    # pitched destination (a column slice): only the addressed pixels are written
    big = torch.full((2, h, w + 7), 9, dtype=torch.uint8, device="cuda")
    efb200.synth_frames(2, h, w, seed, first_frame=0, out=big[:, :, 3:3 + w])
    b = big.cpu().numpy()
    assert np.array_equal(b[1, :, 3:3 + w], oracle.synth_frame(seed, 1, w, h)) and (b[:, :, :3] == 9).all() and (b[:, :, 3 + w:] == 9).all()
