"""GPU parity tests of the frame-ingest kernel (row a21) through the C-ABI against the CPU oracle: depth and colour at the integration
resolution bit-identical, for the reference's default pipeline (2 erosions + range-gated Gaussian), each stage switched off, image
sizes that are not multiples of the 32-pixel tile, and down- / up-sampling to the integration resolution."""
import numpy as np
import pytest

from bundlefusion_b200 import synth
from bundlefusion_b200.image_manager import CUDAImageManager
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,wi,hi,erode,filt", [(640, 480, 640, 480, True, True), (640, 480, 320, 240, True, True), (200, 150, 200, 150, True, False),
                                                   (200, 150, 200, 150, False, True), (97, 61, 130, 77, True, True), (640, 480, 640, 480, False, False)])
def test_ingest_matches_oracle_bit_for_bit(cuda_device, W, H, wi, hi, erode, filt):
    import torch
    d, c, _ = synth.make_frame(77, W, H)
    d[H // 2: H // 2 + 7, W // 3: W // 3 + 25] = -np.inf
    d[3, 5] = 0.0                                                      # the reference treats 0 as invalid in the erosion only
    mgr = CUDAImageManager(wi, hi, cuda_device, erodeSIFTdepth=erode, depthFilter=filt)
    gd, gc = mgr.process(torch.from_numpy(d).to(cuda_device), torch.from_numpy(c).to(cuda_device))
    torch.cuda.synchronize()
    od, oc = orc.ingest_frame(d, c, wi, hi, erode=erode, depth_filter=filt)
    np.testing.assert_array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32))
    np.testing.assert_array_equal(gc.cpu().numpy(), oc)
    assert np.isfinite(od).mean() > 0.3
