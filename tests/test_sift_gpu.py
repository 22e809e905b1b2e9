"""GPU parity tests of the SIFT descriptor matcher (row a18) through the C-ABI against the CPU oracle.  The dot products are exact
integers on both sides, so the SET of matches (and the counter) must be identical; the stored distance goes through acosf, whose
CUDA and glibc implementations may differ in the last bit (tolerance 2e-7).  The reference appends with an atomicAdd (order, and beyond the
128-slot cap the kept subset, race-dependent); this library stores the matches in ascending image-2 feature and keeps the first 128 -- a
deterministic member of the outcomes the reference can produce.  Stored matches are compared with the oracle's as sets."""
import numpy as np
import pytest

from bundlefusion_b200 import synth
from bundlefusion_b200.sift import ImagePairMatch, SiftMatchGPU
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def gpu_match(dev, d1, d2, distmax=0.7, ratiomax=0.8, offset=(0, 0)):
    import torch
    m = SiftMatchGPU(device=dev)
    ipm = ImagePairMatch(dev)
    t1 = torch.from_numpy(d1).to(dev) if len(d1) else torch.zeros(1, 128, dtype=torch.uint8, device=dev)
    t2 = torch.from_numpy(d2).to(dev) if len(d2) else torch.zeros(1, 128, dtype=torch.uint8, device=dev)
    m.SetDescriptors(0, len(d1), t1); m.SetDescriptors(1, len(d2), t2)
    m.GetSiftMatch(len(d1), ipm, offset, distmax, ratiomax)
    torch.cuda.synchronize()
    return ipm.download()


def assert_same(g, o):
    gi, gd, gc = g
    oi, od, oc = o
    assert gc == oc
    gm = {(int(a), int(b)): float(d) for (a, b), d in zip(gi, gd)}
    om = {(int(a), int(b)): float(d) for (a, b), d in zip(oi, od)}
    assert gm.keys() == om.keys()
    for k in gm:
        assert abs(gm[k] - om[k]) <= 2e-7


@pytest.mark.parametrize("n1,n2,nc,seed", [(1024, 1024, 100, 1), (300, 260, 120, 2), (64, 64, 64, 3), (1000, 37, 30, 4), (5, 700, 5, 5),
                                           (1, 1, 1, 6), (129, 65, 60, 7), (777, 1023, 110, 8)])
def test_pair_matches_oracle(cuda_device, n1, n2, nc, seed):
    d1, d2, _ = synth.make_sift_pair(n1, n2, nc, noise=0.05, seed=seed)
    assert_same(gpu_match(cuda_device, d1, d2, offset=(7, 11)), orc.sift_match(d1, d2, offset=(7, 11)))


def test_ties_follow_the_reference(cuda_device):
    """Duplicated descriptors produce equal dot products: with ratiomax > 1 the tie-break decides which index is reported; it must
    be the reference's (bit-reversed lane order), for rows and for columns."""
    base = synth.make_sift_descriptors(40, seed=11)
    d1 = np.concatenate([base, base[:10], synth.make_sift_descriptors(90, seed=12), base[5:15]])      # duplicates at scattered rows
    d2 = np.concatenate([base[:20], synth.make_sift_descriptors(50, seed=13), base[:20], base[10:20]])
    for ratiomax in (1.5, 0.8):
        assert_same(gpu_match(cuda_device, d1, d2, 0.7, ratiomax), orc.sift_match(d1, d2, 0.7, ratiomax))
    g = gpu_match(cuda_device, d1, d2, 0.7, 1.5)
    assert g[2] > 0


def test_cap_and_empty(cuda_device):
    d = synth.make_sift_descriptors(400, seed=9)
    gi, gd, gc = gpu_match(cuda_device, d, d.copy())
    assert gc == 400 and len(gi) == 128                                       # the counter keeps counting, 128 are stored
    assert np.all(gi[:, 0] == gi[:, 1]) and len(set(gi[:, 0].tolist())) == 128
    assert gi[:, 1].tolist() == list(range(128)), "beyond the cap: the 128 matches with the lowest image-2 feature, in that order"
    for _ in range(3):                                                         # and the same every time
        g2 = gpu_match(cuda_device, d, d.copy())
        assert np.array_equal(g2[0], gi) and np.array_equal(g2[1], gd) and g2[2] == gc
    assert gpu_match(cuda_device, d[:0], d)[2] == 0 and gpu_match(cuda_device, d, d[:0])[2] == 0
    z = np.zeros((10, 128), np.uint8)
    assert gpu_match(cuda_device, z, d[:10])[2] == 0


def test_batch_of_pairs_is_one_call(cuda_device):
    """Bundler::matchAndFilter's loop (one frame against every earlier frame) as a single bfSiftMatchBatch call, with an invalid /
    empty image in the middle (its counter must be zeroed, FL/Bundler.cpp:127-130)."""
    import torch
    dev = cuda_device
    cur = synth.make_sift_descriptors(900, seed=21)
    prevs, jobs, ipms = [], [], []
    for k, n in enumerate([1024, 0, 512, 333, 64]):
        if n:
            rng = np.random.default_rng(100 + k)
            sel = rng.permutation(900)[:min(n, 100)]          # stay under the 128-match cap: above it the stored subset is race-dependent
            obs = synth.quantize_descriptors(cur[sel].astype(np.float64) + rng.normal(0, 0.05 * 512 / np.sqrt(128), (len(sel), 128)))
            dk = np.concatenate([obs, synth.make_sift_descriptors(n - len(sel), seed=40 + k)]) if n > len(sel) else obs
        else:
            dk = np.zeros((0, 128), np.uint8)
        prevs.append(dk)
    tcur = torch.from_numpy(cur).to(dev)
    m = SiftMatchGPU(device=dev)
    keep = []
    for k, dk in enumerate(prevs):
        ipm = ImagePairMatch(dev); ipm.d_numMatches.fill_(12345)
        tk = torch.from_numpy(dk).to(dev) if len(dk) else torch.zeros(1, 128, dtype=torch.uint8, device=dev)
        keep.append(tk); ipms.append(ipm)
        jobs.append((tk, len(dk), tcur, len(cur), ipm, (1000 * k, 5)))
    m.matchBatch(jobs)
    torch.cuda.synchronize()
    for k, dk in enumerate(prevs):
        assert_same(ipms[k].download(), orc.sift_match(dk, cur, offset=(1000 * k, 5)))
    assert ipms[1].download()[2] == 0


def test_sort_matches_matches_oracle(cuda_device):
    import torch
    from bundlefusion_b200 import _capi as capi
    dev = cuda_device
    rng = np.random.default_rng(1)
    P = 40
    nm = rng.integers(0, 160, P).astype(np.int32); nm[3] = 0; nm[7] = 128; nm[9] = 1
    d = rng.random((P, 128)).astype(np.float32); ix = rng.integers(0, 1024, (P, 128, 2)).astype(np.uint32)
    d[5, :60] = np.round(d[5, :60], 1)                                # many equal distances
    L = capi.lib(); L.bfSetStream(None)
    t_nm, t_d, t_ix = torch.from_numpy(nm).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(ix.view(np.int32)).to(dev)
    torch.cuda.synchronize()
    capi.check(L.bfSiftSortKeyPointMatches(11, 2, P, t_nm.data_ptr(), t_d.data_ptr(), t_ix.data_ptr()), "bfSiftSortKeyPointMatches")
    torch.cuda.synchronize()
    od, oi = orc.sift_sort_matches(11, 2, P, nm, d, ix)
    np.testing.assert_array_equal(t_d.cpu().numpy(), od)
    np.testing.assert_array_equal(t_ix.cpu().numpy().view(np.uint32), oi)
