import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _make_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("vy_make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


DECODE_REF_NAMES = ["voc416", "coco608", "vid320", "agnostic", "small", "nonsquare"]


def load_decode_ref(name):
    """A reference-executed decode fixture (tests/golden/decode_ref_*.npz, provenance in make_golden.py)
    plus its head maps, regenerated from the stored seed and verified by checksum.  Reads nothing
    outside the repo (no /root/reference at test time)."""
    import numpy as np
    mg = _make_golden()
    z = dict(np.load(os.path.join(GOLDEN, "decode_ref_%s.npz" % name)))
    heads = mg.decode_ref_heads(name)
    got = np.array([h.astype(np.float64).sum() for h in heads])
    assert np.array_equal(got, z["heads_sum"]), "regenerated head maps differ from the generator's"
    z["heads"] = heads
    z["B"], z["C"], z["R"], z["agnostic"] = int(z["B"]), int(z["C"]), int(z["R"]), bool(z["agnostic"])
    return z


def assert_decode_close(got, ref, score_ulps=None, box_eps=None, rtol=None):
    """Per-element decode comparison of (…, 6) rows [id, score, x1, y1, x2, y2].

    ids exact.  Either an ulp-style bound (oracle vs reference: `score_ulps` ulps on the score,
    `box_eps`·2^-23·(|centre| + |half size|) on a corner, the rounding unit of the subtraction
    that forms it, yolo3.py:176-177) or the north_star's 1e-5 RELATIVE bound applied per element:
    |Δscore| <= rtol·score, |Δcorner| <= rtol·(|centre| + |half size|)."""
    import numpy as np
    got, ref = np.asarray(got, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got[..., 0], ref[..., 0])
    r64 = ref.astype(np.float64)
    cx, hw = (r64[..., 2] + r64[..., 4]) / 2, (r64[..., 4] - r64[..., 2]) / 2
    cy, hh = (r64[..., 3] + r64[..., 5]) / 2, (r64[..., 5] - r64[..., 3]) / 2
    span = np.stack([np.abs(cx) + np.abs(hw), np.abs(cy) + np.abs(hh)] * 2, axis=-1)
    d = np.abs(got.astype(np.float64) - r64)
    eps = float(np.finfo(np.float32).eps)
    if rtol is not None:
        sb, bb = rtol * np.abs(r64[..., 1]), rtol * span
    else:
        sb, bb = score_ulps * eps * np.abs(r64[..., 1]), box_eps * eps * span
    assert (d[..., 1] <= sb + 1e-45).all(), "score: worst error / bound = %g" % (d[..., 1] / (sb + 1e-45)).max()
    assert (d[..., 2:] <= bb).all(), "box: worst error / bound = %g" % (d[..., 2:] / np.maximum(bb, 1e-300)).max()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
