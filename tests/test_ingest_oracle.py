"""Known-answer tests pinning oracle/ingest_oracle.c (row a21): erosion removes isolated / boundary pixels by the reference's
30 %-of-49 rule, the range-gated Gaussian does not mix across a depth step, resampling is the reference's nearest rule."""
import numpy as np

from oracle import oracle as orc

F = np.float32


def test_erode_rule_by_hand():
    d = np.full((40, 40), 1.0, F)
    d[20, 20] = 2.0                                   # an outlier: all 48 neighbours differ by > 0.05 -> 48/49 >= 0.3 -> removed
    d[5:8, 5:8] = -np.inf                             # a 3x3 hole: a pixel 3 away sees 9 bad taps of 49 = 0.18 -> kept
    d[25:35, 0:5] = -np.inf                           # a wide hole: the pixel beside it sees 3 columns x 7 rows = 21 of 49 bad taps = 0.43
    out, _ = orc.ingest_frame(d, np.zeros((40, 40, 4), np.uint8), 40, 40, depth_filter=False)
    assert out[20, 20] == -np.inf
    assert out[6, 10] == 1.0 and out[10, 6] == 1.0
    assert np.all(np.isinf(out[5:8, 5:8]))
    assert out[6, 8] == 1.0                           # beside the small hole: 9 of 49 bad -> kept
    assert out[30, 5] == -np.inf                      # beside the wide hole: eroded in the first pass
    assert out[30, 6] == -np.inf                      # 14 of 49 in the first pass (kept), 21 of 49 in the second (col 5 is gone by then)
    assert out[30, 7] == 1.0
    # a pixel in the image corner counts only in-image taps (16 of them, all good) against the FULL window size 49
    assert out[0, 0] == 1.0


def test_gate_keeps_step_edges_and_resample_rule():
    d = np.full((60, 80), 1.0, F); d[:, 40:] = 1.5
    c = (np.arange(60 * 80 * 4) % 251).astype(np.uint8).reshape(60, 80, 4)
    out, cout = orc.ingest_frame(d, c, 80, 60, erode=False)
    fin = out[np.isfinite(out)]
    assert np.all((np.abs(fin - 1.0) < 1e-6) | (np.abs(fin - 1.5) < 1e-6))
    np.testing.assert_array_equal(cout, c)
    out2, cout2 = orc.ingest_frame(d, c, 37, 23, erode=False, depth_filter=False)
    xs = (np.arange(37, dtype=F) * F(79.0 / 36.0) + F(0.5)).astype(int); ys = (np.arange(23, dtype=F) * F(59.0 / 22.0) + F(0.5)).astype(int)
    np.testing.assert_array_equal(out2, d[np.ix_(ys, xs)])
    np.testing.assert_array_equal(cout2, c[np.ix_(ys, xs)])
