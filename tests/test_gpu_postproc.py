"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(videoyolo_b200.ops -> ctypes -> libvyolo.so) and is checked against the CPU oracle.

Bars (BASELINE.json north_star):
  * box_nms keep-sets, output rows and source-row indices: BIT-EXACT vs the oracle on identical input;
  * fused decode+NMS: BIT-EXACT vs oracle box_nms applied to the GPU-decoded rows;
  * decode: scores/boxes within 1e-5 relative of the fp32 CPU restatement (+ an absolute term for
    the cx - w/2 cancellation, proportional to the coordinate range);
  * bbox_iou: 1e-12 relative (float64) vs the reference's own outputs.
"""
import json
import os

import numpy as np
import pytest

import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

DEC_RTOL = 1e-5          # north_star: "decoded boxes and scores within 1e-5 relative (fp32)"


@pytest.fixture(scope="module")
def vy():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import videoyolo_b200
    videoyolo_b200._lib.lib()          # fails loudly when the CUDA library is missing
    return videoyolo_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def random_heads(rng, B, C, size, std=1.0, quant=None):
    hs = [rng.normal(0, std, size=(B, 3 * (5 + C), g, g)).astype(np.float32) for g in oracle.grid_sizes(size)]
    if quant:
        hs = [np.round(h * quant) / quant for h in hs]      # massive score ties
    return [h.astype(np.float32) for h in hs]


def trained_heads(rng, B, C, size):
    from videoyolo_b200.synth import trained_like_heads
    return trained_like_heads(rng, B, C, size)


AN, ST = oracle.ANCHORS[::-1], oracle.STRIDES[::-1]


# ----------------------------------------------------------------------------------- decode
@pytest.mark.parametrize("B,C,size,agnostic", [(1, 20, 416, False), (2, 80, 608, False), (3, 30, 320, False),
                                               (2, 30, 416, True), (1, 1, 64, False), (2, 7, 96, False)])
def test_decode_matches_oracle(vy, B, C, size, agnostic):
    rng = np.random.RandomState(B * 1000 + C + size)
    heads = random_heads(rng, B, C, size, std=1.5)
    got = vy.yolo3_decode([dev(h) for h in heads], C, AN, ST, agnostic).cpu().numpy()
    ref = oracle.decode_c(heads, C, agnostic=agnostic)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got[..., 0], ref[..., 0])
    np.testing.assert_allclose(got[..., 1], ref[..., 1], rtol=DEC_RTOL, atol=1e-12)
    # corners are differences of O(size) centres and O(anchor*e^t) half-sizes
    scale = np.maximum(np.abs(ref[..., 2:]).max(), size)
    np.testing.assert_allclose(got[..., 2:], ref[..., 2:], rtol=DEC_RTOL, atol=DEC_RTOL * scale)


def test_yolo_output_block_surface(vy):
    rng = np.random.RandomState(3)
    out = vy.YOLOOutputV3(0, 20, AN[0], ST[0], in_channels=32).cuda()
    x = torch.from_numpy(rng.normal(size=(2, 32, 13, 13)).astype(np.float32)).cuda()
    with torch.no_grad():
        dets = out(x)
        pred = out.prediction(x)
    assert dets.shape == (2, 20 * 13 * 13 * 3, 6)
    ref = oracle.decode_numpy(pred.cpu().numpy(), AN[0], ST[0], 20)
    np.testing.assert_allclose(dets.cpu().numpy()[..., 1], ref[..., 1], rtol=DEC_RTOL, atol=1e-12)
    with pytest.raises(ValueError):
        vy.YOLOOutputV3(0, 20, AN[0], ST[0], alloc_size=(8, 8)).decode(torch.zeros(1, 75, 13, 13).cuda())


# ----------------------------------------------------------------------------------- box_nms (generic)
def _check_nms(vy, d, **kw):
    okw = dict(kw)
    out_rows = okw.pop("out_rows", None)
    exp, rec = oracle.box_nms_c(d, return_record=True, **okw)
    got, kept = vy.box_nms(dev(d), return_kept=True, out_rows=out_rows, **okw)
    got, kept = got.cpu().numpy(), kept.cpu().numpy()
    if out_rows is not None:
        exp, rec = exp[..., :out_rows, :], rec[..., :out_rows]
    np.testing.assert_array_equal(kept, rec)
    np.testing.assert_array_equal(got, exp)


def test_box_nms_mxnet_known_answers(vy, golden_dir):
    with open(os.path.join(golden_dir, "box_nms_mxnet_doc.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        d = np.array(c["data"], dtype=np.float32)
        got, kept = vy.box_nms(dev(d), return_kept=True, **c["args"])
        exp = np.array(c["expected"], dtype=np.float32)
        if c.get("approx"):
            np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=1e-5, atol=1e-6, err_msg=c["name"])
        else:
            np.testing.assert_array_equal(got.cpu().numpy(), exp, err_msg=c["name"])
        np.testing.assert_array_equal(kept.cpu().numpy(), np.array(c["kept"]), err_msg=c["name"])


def _rand_dets(rng, B, R, n_cls, quant=None, scale=100.0, W=6, coord_start=2, score_index=1, id_index=0):
    xy = rng.uniform(0, scale, size=(B, R, 2))
    wh = rng.uniform(-0.05 * scale, 0.5 * scale, size=(B, R, 2))
    sc = rng.uniform(-0.1, 1.0, size=(B, R))
    if quant:
        sc = np.round(sc * quant) / quant
    d = rng.uniform(-1, 1, size=(B, R, W))
    d[..., coord_start:coord_start + 2] = xy
    d[..., coord_start + 2:coord_start + 4] = xy + wh
    d[..., score_index] = sc
    if id_index >= 0:
        d[..., id_index] = rng.randint(0, n_cls, size=(B, R))
    return d.astype(np.float32)


@pytest.mark.parametrize("R,topk", [(1, 400), (37, 5), (400, 400), (3000, 400), (3000, 1), (5000, 1024), (70000, 400)])
@pytest.mark.parametrize("force", [False, True])
def test_box_nms_bit_exact_random(vy, R, topk, force):
    rng = np.random.RandomState(R + topk)
    d = _rand_dets(rng, 3, R, 4, quant=(64 if R % 2 else None))
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=topk, id_index=0, force_suppress=force)


def test_box_nms_argument_variants(vy):
    rng = np.random.RandomState(77)
    d = _rand_dets(rng, 2, 900, 3)
    _check_nms(vy, d, overlap_thresh=0.3, valid_thresh=0.0, topk=200, id_index=-1)               # no ids
    _check_nms(vy, d, overlap_thresh=0.5, valid_thresh=-1.0, topk=300, id_index=0)                # negative scores take part
    _check_nms(vy, d, overlap_thresh=0.5, valid_thresh=0.2, topk=300, id_index=0, background_id=1)
    _check_nms(vy, d, overlap_thresh=0.5, valid_thresh=5.0, topk=300, id_index=0)                 # nothing valid
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, out_rows=100)  # fused slice
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=50, id_index=0, out_rows=100)   # topk < out_rows
    # other column layout: [score, x1,y1,x2,y2, pad, id, pad]
    d8 = _rand_dets(rng, 2, 500, 5, W=8, coord_start=1, score_index=0, id_index=6)
    _check_nms(vy, d8, overlap_thresh=0.45, valid_thresh=0.0, topk=100, coord_start=1, score_index=0, id_index=6)
    # center format in and out
    c = d.copy()
    c[..., 2:4] = (d[..., 2:4] + d[..., 4:6]) / 2
    c[..., 4:6] = d[..., 4:6] - d[..., 2:4]
    _check_nms(vy, c, overlap_thresh=0.45, valid_thresh=0.01, topk=300, id_index=0, in_format="center", out_format="center")
    _check_nms(vy, c, overlap_thresh=0.45, valid_thresh=0.01, topk=300, id_index=0, in_format="center", out_format="corner")
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=300, id_index=0, in_format="corner", out_format="center")
    # temporal model: (B, T, R, 6), leading dims are batch (yolo3_temporal.py:545)
    d4 = _rand_dets(rng, 6, 300, 3).reshape(2, 3, 300, 6)
    _check_nms(vy, d4, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)


def test_box_nms_adversarial_orders(vy):
    """ascending scores (every row beats the running threshold), all-equal scores (pure row-order
    ties) and duplicates of one box."""
    R = 40000
    d = _rand_dets(np.random.RandomState(1), 2, R, 3)
    d[..., 1] = np.linspace(0.02, 0.99, R, dtype=np.float32)
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)
    d[..., 1] = 0.5
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)
    d[..., 2:6] = [10, 10, 50, 60]
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)


def test_box_nms_fed_reference_decoded_boxes(vy):
    """The north-star wording: keep-sets bit-exact when fed the reference's decoded boxes."""
    for (B, C, size, regime) in [(2, 20, 416, "R"), (2, 30, 320, "T"), (1, 80, 608, "T")]:
        rng = np.random.RandomState(C)
        heads = random_heads(rng, B, C, size) if regime == "R" else trained_heads(rng, B, C, size)
        dets = oracle.decode_c(heads, C)
        _check_nms(vy, dets, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, out_rows=100)


# ----------------------------------------------------------------------------------- box_nms, > 1024 candidates
def _sparse_dets(rng, B, R, n_cls, extent, box):
    """small boxes scattered over a large extent: most of them survive, so the kept lists grow long"""
    d = _rand_dets(rng, B, R, n_cls)
    xy = rng.uniform(0, extent, size=(B, R, 2))
    wh = rng.uniform(0.2 * box, box, size=(B, R, 2))
    d[..., 2:4] = xy
    d[..., 4:6] = xy + wh
    return d.astype(np.float32)


@pytest.mark.parametrize("R,topk,n_cls", [(1025, -1, 3), (3000, -1, 4), (5000, 2000, 1), (24000, -1, 20), (20000, 15000, 80)])
@pytest.mark.parametrize("force", [False, True])
def test_box_nms_large_bit_exact_random(vy, R, topk, n_cls, force):
    """MXNet's default topk=-1 / topk > 1024 (BASELINE config 4 arguments): sorted-list path."""
    rng = np.random.RandomState(R + n_cls)
    d = _rand_dets(rng, 3, R, n_cls, quant=(64 if R % 2 else None))
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.001, topk=topk, id_index=0, force_suppress=force)


@pytest.mark.parametrize("force", [False, True])
def test_box_nms_large_long_kept_lists(vy, force):
    """thousands of survivors per segment: several staged chunks of the kept list, full 1024-wide tiles"""
    rng = np.random.RandomState(17)
    d = _sparse_dets(rng, 2, 16000, 3, extent=300.0, box=4.0)
    exp, rec = oracle.box_nms_c(d, overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0,
                                force_suppress=force, return_record=True)
    assert (rec >= 0).sum(axis=1).min() > 3000
    got, kept = vy.box_nms(dev(d), overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0,
                           force_suppress=force, return_kept=True)
    np.testing.assert_array_equal(kept.cpu().numpy(), rec)
    np.testing.assert_array_equal(got.cpu().numpy(), exp)


def test_box_nms_large_argument_variants(vy):
    rng = np.random.RandomState(78)
    d = _rand_dets(rng, 2, 6000, 5, scale=200.0)
    _check_nms(vy, d, overlap_thresh=0.3, valid_thresh=0.0, topk=-1, id_index=-1)                  # no ids
    _check_nms(vy, d, overlap_thresh=0.5, valid_thresh=-1.0, topk=-1, id_index=0)                   # negative scores take part
    _check_nms(vy, d, overlap_thresh=0.5, valid_thresh=0.2, topk=-1, id_index=0, background_id=1)
    _check_nms(vy, d, overlap_thresh=0.5, valid_thresh=5.0, topk=-1, id_index=0)                    # nothing valid
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=-1, id_index=0, out_rows=100)    # fused slice
    _check_nms(vy, d, overlap_thresh=0.0, valid_thresh=0.01, topk=-1, id_index=0)                   # thr 0: exact-division path
    _check_nms(vy, d, overlap_thresh=0.999, valid_thresh=0.01, topk=-1, id_index=0)
    d8 = _rand_dets(rng, 2, 3000, 5, W=8, coord_start=1, score_index=0, id_index=6, scale=150.0)
    _check_nms(vy, d8, overlap_thresh=0.45, valid_thresh=0.0, topk=-1, coord_start=1, score_index=0, id_index=6)
    c = d.copy()
    c[..., 2:4] = (d[..., 2:4] + d[..., 4:6]) / 2
    c[..., 4:6] = d[..., 4:6] - d[..., 2:4]
    _check_nms(vy, c, overlap_thresh=0.45, valid_thresh=0.01, topk=-1, id_index=0, in_format="center", out_format="center")
    _check_nms(vy, c, overlap_thresh=0.45, valid_thresh=0.01, topk=-1, id_index=0, in_format="center", out_format="corner")
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.01, topk=-1, id_index=0, in_format="corner", out_format="center")
    d4 = _rand_dets(rng, 6, 1500, 3).reshape(2, 3, 1500, 6)                                        # (B, T, R, 6)
    _check_nms(vy, d4, overlap_thresh=0.45, valid_thresh=0.01, topk=-1, id_index=0)
    # all-equal scores and duplicates of one box (pure row-order ties; a single survivor per class)
    e = _rand_dets(rng, 2, 5000, 3)
    e[..., 1] = 0.5
    _check_nms(vy, e, overlap_thresh=0.45, valid_thresh=0.01, topk=-1, id_index=0)
    e[..., 2:6] = [10, 10, 50, 60]
    _check_nms(vy, e, overlap_thresh=0.45, valid_thresh=0.01, topk=-1, id_index=0)


def test_box_nms_large_grid_index_equals_exhaustive(vy):
    """The spatial index over the kept boxes must not change a single keep decision: indexed kernel ==
    exhaustive kernel (same exact predicate) on awkward geometry, and == the oracle where the oracle is defined."""
    rng = np.random.RandomState(23)
    B, R = 2, 12000
    d = _sparse_dets(rng, B, R, 3, extent=200.0, box=6.0)
    # a mix of scales (tiny .. huge), far-away coordinates, degenerate and inverted boxes
    idx = rng.permutation(R)
    d[:, idx[:1500], 4:6] = d[:, idx[:1500], 2:4] + rng.uniform(20, 400, size=(B, 1500, 2))            # large boxes
    d[:, idx[1500:2500], 4:6] = d[:, idx[1500:2500], 2:4] + rng.uniform(1e-3, 0.2, size=(B, 1000, 2))  # tiny boxes
    d[:, idx[2500:2800], 2:6] += 3.0e6                                                                 # far away: coarse fp32 grid
    d[:, idx[2800:3000], 4] = d[:, idx[2800:3000], 2]                                                  # zero width
    d[:, idx[3000:3200], 5] = d[:, idx[3000:3200], 3] - 1.0                                            # inverted
    d[:, idx[3200:3300], 2:6] *= 1.0e12                                                                # enormous
    d[:, idx[3300:3400], 2:6] *= 1.0e-12                                                               # microscopic
    d = d.astype(np.float32)
    for force in (False, True):
        for fmt in ("corner", "center"):
            kw = dict(overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0, force_suppress=force,
                      in_format=fmt, out_format=fmt, return_kept=True)
            g_out, g_kept = vy.box_nms(dev(d), **kw)
            e_out, e_kept = vy.box_nms(dev(d), _exhaustive=True, **kw)
            assert torch.equal(g_kept, e_kept) and torch.equal(g_out, e_out)
            if fmt == "corner":
                exp, rec = oracle.box_nms_c(d, overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0,
                                            force_suppress=force, return_record=True)
                np.testing.assert_array_equal(g_kept.cpu().numpy(), rec)
    # non-finite coordinates: only GPU against GPU (same predicate code on both sides)
    d2 = d.copy()
    d2[:, idx[4000:4050], 2] = np.inf
    d2[:, idx[4050:4100], 5] = np.nan
    d2[:, idx[4100:4150], 2:6] = -np.inf
    for thr in (0.45, 0.1, 0.9):
        kw = dict(overlap_thresh=thr, valid_thresh=0.001, topk=-1, id_index=0, return_kept=True)
        g_out, g_kept = vy.box_nms(dev(d2), **kw)
        e_out, e_kept = vy.box_nms(dev(d2), _exhaustive=True, **kw)
        assert torch.equal(g_kept, e_kept)
        assert torch.equal(torch.nan_to_num(g_out, nan=-7.0), torch.nan_to_num(e_out, nan=-7.0))


def test_box_nms_large_grid_index_full_size(vy):
    """BASELINE config 4 rows (851 760 per image): indexed == exhaustive, force_suppress off and on."""
    B, C, size = 2, 80, 416
    g = torch.Generator(device="cuda").manual_seed(1239)
    heads = [torch.randn((B, 3 * (5 + C), s, s), generator=g, device="cuda") for s in oracle.grid_sizes(size)]
    dets = vy.yolo3_decode(heads, C, AN, ST)
    for force in (False, True):
        kw = dict(overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0, force_suppress=force, return_kept=True)
        g_out, g_kept = vy.box_nms(dets, **kw)
        e_out, e_kept = vy.box_nms(dets, _exhaustive=True, **kw)
        assert torch.equal(g_kept, e_kept) and torch.equal(g_out, e_out)


def test_box_nms_large_fed_reference_decoded_boxes(vy):
    """BASELINE config 4 arguments (valid_thresh 0.001, topk -1, force on/off) on decoded YOLO rows."""
    rng = np.random.RandomState(41)
    heads = random_heads(rng, 2, 20, 160)
    dets = oracle.decode_c(heads, 20)                                   # (2, 31500, 6)
    for force in (False, True):
        _check_nms(vy, dets, overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0, force_suppress=force)
    # through the block surface: set_nms(nms_topk=-1) routes decode -> box_nms (yolo3.py:523-530)
    net = vy.get_yolov3_postprocess(["c%d" % i for i in range(20)])
    net.set_nms(nms_thresh=0.45, nms_topk=-1, post_nms=100)
    ids, scores, bboxes = net(*[dev(h) for h in heads])
    gd = vy.yolo3_decode([dev(h) for h in heads], 20, AN, ST).cpu().numpy()
    o_ids, o_sc, o_bb = oracle.yolov3_tail(gd, nms_topk=-1)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(scores.cpu().numpy(), o_sc)
    np.testing.assert_array_equal(bboxes.cpu().numpy(), o_bb)


# ----------------------------------------------------------------------------------- fused path
def _fused_vs_oracle(vy, heads, C, agnostic=False, nms_thresh=0.45, topk=400, post_nms=100, valid=0.01, force=False):
    hd = [dev(h) for h in heads]
    out, kept = vy.yolo3_decode_nms(hd, C, AN, ST, nms_thresh, valid, topk, post_nms, force, agnostic)
    dets = vy.yolo3_decode(hd, C, AN, ST, agnostic).cpu().numpy()
    exp, rec = oracle.box_nms_c(dets, overlap_thresh=nms_thresh, valid_thresh=valid, topk=topk, id_index=0,
                                force_suppress=force, return_record=True)
    np.testing.assert_array_equal(kept.cpu().numpy(), rec[:, :post_nms])
    np.testing.assert_array_equal(out.cpu().numpy(), exp[:, :post_nms])
    return out, kept


@pytest.mark.parametrize("B,C,size,regime", [(1, 20, 416, "R"), (4, 20, 416, "T"), (3, 80, 608, "R"),
                                             (2, 80, 608, "T"), (5, 30, 320, "R"), (5, 30, 320, "T"),
                                             (2, 3, 96, "R"), (1, 1, 64, "R")])
def test_fused_equals_decode_then_oracle_nms(vy, B, C, size, regime):
    rng = np.random.RandomState(B + C + size)
    heads = random_heads(rng, B, C, size) if regime == "R" else trained_heads(rng, B, C, size)
    _fused_vs_oracle(vy, heads, C)


def test_fused_variants(vy):
    rng = np.random.RandomState(9)
    heads = random_heads(rng, 2, 20, 416)
    _fused_vs_oracle(vy, heads, 20, agnostic=True)
    _fused_vs_oracle(vy, heads, 20, topk=1)
    _fused_vs_oracle(vy, heads, 20, topk=1024, post_nms=300)
    _fused_vs_oracle(vy, heads, 20, force=True)
    _fused_vs_oracle(vy, heads, 20, valid=0.0)
    _fused_vs_oracle(vy, heads, 20, valid=0.9)                     # almost nothing valid
    _fused_vs_oracle(vy, [np.full_like(h, -30.0) for h in heads], 20)   # nothing valid at all
    _fused_vs_oracle(vy, random_heads(rng, 2, 20, 416, quant=2), 20)    # massive exact ties
    _fused_vs_oracle(vy, [np.zeros_like(h) for h in heads], 20)         # every score == 0.25


def test_finalize_branches(vy):
    """Every branch of the finalize kernel against the oracle: counting regroup vs (class, rank) sort
    (more than 256 classes), short-segment lanes vs 32x32 tiles with chained blocks (2-3 classes: segments
    of >100 slots that straddle many blocks), topk at the CTA size, and lists the bucket front end hands to
    the general path (one bin holding far more keys than the CTA has threads)."""
    rng = np.random.RandomState(21)
    _fused_vs_oracle(vy, random_heads(rng, 2, 300, 96), 300)                       # C > FIN_CMAX: sort regroup
    _fused_vs_oracle(vy, random_heads(rng, 2, 300, 96), 300, topk=1024, post_nms=1024)
    for C in (2, 3):
        heads = random_heads(rng, 3, C, 416)
        _fused_vs_oracle(vy, heads, C)                                             # long segments: tiles + chains
        _fused_vs_oracle(vy, heads, C, topk=1024, post_nms=1024)
        _fused_vs_oracle(vy, heads, C, topk=70, post_nms=100)                      # one or two blocks
        _fused_vs_oracle(vy, heads, C, force=True, topk=64)                        # all pairs, short path
        _fused_vs_oracle(vy, heads, C, force=True, topk=65)                        # all pairs, tiles
    # rows: a fat bin at the K-th key (2000 equal scores, keys differ only in the row) under a few spread ones
    for n_eq, topk in ((2000, 400), (6000, 400), (2000, 1024), (9000, 100)):
        d = _rand_dets(rng, 2, n_eq + 300, 5)
        d[:, :n_eq, 1] = 0.5
        d[:, n_eq:, 1] = rng.uniform(0.5, 1.0, size=(2, 300)).astype(np.float32)
        d[:, n_eq + 250:, 1] = rng.uniform(1e-3, 1e-2, size=(2, 50)).astype(np.float32)
        d = d[:, rng.permutation(d.shape[1])]
        _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.0, topk=topk, id_index=0)
        _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.0, topk=topk, id_index=0, force_suppress=True)
    # scores spanning many binades, negative ones included
    d = _rand_dets(rng, 2, 5000, 7)
    d[..., 1] = (rng.standard_normal((2, 5000)) * np.exp(rng.uniform(-20, 3, size=(2, 5000)))).astype(np.float32)
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=-1e9, topk=700, id_index=0)
    _check_nms(vy, d, overlap_thresh=0.45, valid_thresh=0.0, topk=33, id_index=0)


def test_fused_ascending_plane_order(vy):
    """class logits increasing with the class index: later planes always beat the running threshold."""
    rng = np.random.RandomState(4)
    C = 80
    heads = random_heads(rng, 2, C, 416)
    for h in heads:
        v = h.reshape(h.shape[0], 3, 5 + C, -1)
        v[:, :, 5:, :] = np.sort(v[:, :, 5:, :], axis=2)
    _fused_vs_oracle(vy, heads, C)


def test_fused_close_to_cpu_end_to_end(vy, golden_dir):
    """Against the committed oracle fixtures (CPU decode + CPU NMS): identical detections wherever the
    CPU and GPU scores are not within rounding of each other -- checked as >= 99% identical rows."""
    for name in ["voc416_random", "vid320_trained", "coco_small_trained"]:
        z = np.load(os.path.join(golden_dir, "postproc_regress_%s.npz" % name))
        C = int(z["C"])
        net = vy.get_yolov3_postprocess(["c"] * C)
        ids, scores, bboxes = net(*[dev(z[k]) for k in ("h0", "h1", "h2")])
        kept = net.last_kept_rows.cpu().numpy()
        same = (kept == z["kept_rows"]).mean()
        assert same >= 0.99, (name, same)
        m = kept == z["kept_rows"]
        np.testing.assert_allclose(scores.cpu().numpy()[..., 0][m], z["scores"][..., 0][m], rtol=DEC_RTOL, atol=1e-12)
        np.testing.assert_allclose(bboxes.cpu().numpy()[m], z["bboxes"][m], rtol=DEC_RTOL, atol=DEC_RTOL * 640)


def test_yolov3_block_surface_and_set_nms(vy):
    rng = np.random.RandomState(12)
    C = 20
    heads = trained_heads(rng, 2, C, 416)
    net = vy.get_yolov3_postprocess(["c%d" % i for i in range(C)])
    assert net.classes == ["c%d" % i for i in range(C)]
    net.set_nms(nms_thresh=0.45, nms_topk=400)                      # detect_yolo3.py:200
    ids, scores, bboxes = net(*[dev(h) for h in heads])
    assert ids.shape == (2, 100, 1) and scores.shape == (2, 100, 1) and bboxes.shape == (2, 100, 4)
    o_ids, o_sc, o_bb = oracle.yolov3_tail(vy.yolo3_decode([dev(h) for h in heads], C, AN, ST).cpu().numpy())
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(bboxes.cpu().numpy(), o_bb)
    # nms disabled (yolo3.py:525): the unsorted full tensor is split
    net.set_nms(nms_thresh=1.5)
    ids2, _, _ = net(*[dev(h) for h in heads])
    assert ids2.shape == (2, C * 10647, 1)
    # post_nms = -1: all rows of the operator output are returned
    net.set_nms(nms_thresh=0.45, nms_topk=400, post_nms=-1)
    ids3, sc3, _ = net(*[dev(h) for h in heads])
    assert ids3.shape == (2, C * 10647, 1)
    np.testing.assert_array_equal(ids3.cpu().numpy()[:, :100], o_ids)
    assert (ids3.cpu().numpy()[:, 400:] == -1).all()
    with pytest.raises(RuntimeError):
        net(*[torch.from_numpy(h) for h in heads])                   # CPU tensors: no fallback


# ----------------------------------------------------------------------------------- full-size properties
def test_full_size_coco608_b64_properties(vy):
    """BASELINE config 2 at full size: size-independent properties + oracle spot check on 3 frames."""
    B, C, size = 64, 80, 608
    g = torch.Generator(device="cuda").manual_seed(1236)
    heads = [torch.randn((B, 3 * (5 + C), s, s), generator=g, device="cuda") for s in oracle.grid_sizes(size)]
    out, kept = vy.yolo3_decode_nms(heads, C, AN, ST)
    out2, kept2 = vy.yolo3_decode_nms(heads, C, AN, ST)
    assert torch.equal(out, out2) and torch.equal(kept, kept2)                  # deterministic
    o = out.cpu().numpy()
    k = kept.cpu().numpy()
    valid = k >= 0
    assert ((o[..., 0] >= 0) == valid).all()
    for b in range(B):
        n = valid[b].sum()
        assert (valid[b][:n]).all()                                             # survivors first, padding after
        assert (np.diff(o[b, :n, 1]) <= 0).all()                                # score order
        assert len(set(k[b, :n])) == n                                          # unique source rows
        assert (o[b, n:] == -1).all()
    # survivors of one class never overlap above the threshold
    bx = out[..., 2:6]
    area = (bx[..., 2] - bx[..., 0]).clamp(min=0) * (bx[..., 3] - bx[..., 1]).clamp(min=0)
    iw = (torch.minimum(bx[:, :, None, 2], bx[:, None, :, 2]) - torch.maximum(bx[:, :, None, 0], bx[:, None, :, 0])).clamp(min=0)
    ih = (torch.minimum(bx[:, :, None, 3], bx[:, None, :, 3]) - torch.maximum(bx[:, :, None, 1], bx[:, None, :, 1])).clamp(min=0)
    iou = iw * ih / (area[:, :, None] + area[:, None, :] - iw * ih)
    same = (out[:, :, None, 0] == out[:, None, :, 0]) & (out[:, :, None, 0] >= 0)
    same &= ~torch.eye(out.shape[1], dtype=torch.bool, device="cuda")[None]
    assert not ((iou > 0.45) & same).any()
    # batch independence (what frame sharding relies on): frames 5..7 alone == the same frames in the batch
    sub, ksub = vy.yolo3_decode_nms([h[5:8].contiguous() for h in heads], C, AN, ST)
    assert torch.equal(sub, out[5:8]) and torch.equal(ksub, kept[5:8])
    # oracle on the GPU-decoded rows of those frames
    dets = vy.yolo3_decode([h[5:8].contiguous() for h in heads], C, AN, ST).cpu().numpy()
    exp, rec = oracle.box_nms_c(dets, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, return_record=True)
    np.testing.assert_array_equal(ksub.cpu().numpy(), rec[:, :100])
    np.testing.assert_array_equal(sub.cpu().numpy(), exp[:, :100])


def test_box_nms_idempotent_full_size(vy):
    """NMS of an NMS output changes nothing (valid rows stay, -1 padding is dropped by valid_thresh)."""
    rng = np.random.RandomState(5)
    d = dev(_rand_dets(rng, 4, 200000, 20))
    a = vy.box_nms(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)
    b = vy.box_nms(a, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)
    assert torch.equal(a, b)


def test_full_size_stress_config4_properties(vy):
    """BASELINE config 4 at full row count (80 cls x 10647 boxes = 851760 rows, valid_thresh 0.001, topk -1,
    force_suppress off/on): size-independent properties of the operator output + idempotence."""
    B, C, size = 2, 80, 416
    g = torch.Generator(device="cuda").manual_seed(1238)
    heads = [torch.randn((B, 3 * (5 + C), s, s), generator=g, device="cuda") for s in oracle.grid_sizes(size)]
    dets = vy.yolo3_decode(heads, C, AN, ST)
    R = dets.shape[1]
    assert R == 851760
    n_valid = (dets[..., 1] > 0.001).sum(dim=1)
    for force in (False, True):
        out, kept = vy.box_nms(dets, overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0,
                               force_suppress=force, return_kept=True)
        out2, kept2 = vy.box_nms(dets, overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0,
                                 force_suppress=force, return_kept=True)
        assert torch.equal(out, out2) and torch.equal(kept, kept2)                       # deterministic
        valid = kept >= 0
        n = valid.sum(dim=1)
        assert (n > 0).all() and (n <= n_valid).all()
        for b in range(B):
            nb = int(n[b])
            assert bool(valid[b, :nb].all()) and not bool(valid[b, nb:].any())            # survivors first
            assert bool((out[b, nb:] == -1).all())
            sc = out[b, :nb, 1]
            assert bool((sc[1:] <= sc[:-1]).all())                                       # score order
            rows = kept[b, :nb].long()
            assert int(torch.unique(rows).numel()) == nb                                 # unique source rows
            assert torch.equal(out[b, :nb], dets[b][rows])                               # rows copied verbatim
            assert bool((sc > 0.001).all())
            # the best row always survives; equal-score runs keep ascending source rows
            top = int(torch.argmax(dets[b, :, 1]))
            assert int(rows[0]) == top or float(dets[b, top, 1]) == float(sc[0])
            tie = sc[1:] == sc[:-1]
            assert bool((rows[1:][tie] > rows[:-1][tie]).all())
            # survivors do not suppress each other: checked on the 2000 best
            m = min(nb, 2000)
            bx, ids = out[b, :m, 2:6], out[b, :m, 0]
            area = (bx[:, 2] - bx[:, 0]).clamp(min=0) * (bx[:, 3] - bx[:, 1]).clamp(min=0)
            iw = (torch.minimum(bx[:, None, 2], bx[None, :, 2]) - torch.maximum(bx[:, None, 0], bx[None, :, 0])).clamp(min=0)
            ih = (torch.minimum(bx[:, None, 3], bx[None, :, 3]) - torch.maximum(bx[:, None, 1], bx[None, :, 1])).clamp(min=0)
            iou = iw * ih / (area[:, None] + area[None, :] - iw * ih)
            clash = (iou > 0.45 + 1e-6) & ~torch.eye(m, dtype=torch.bool, device="cuda")   # torch may contract the IoU arithmetic
            if not force:
                clash &= ids[:, None] == ids[None, :]
            assert not bool(clash.any())
        # idempotence: the survivors survive a second pass unchanged
        again = vy.box_nms(out, overlap_thresh=0.45, valid_thresh=0.001, topk=-1, id_index=0, force_suppress=force)
        assert torch.equal(again, out)


def test_cuda_graph_replay_equals_eager(vy):
    """The fused call captured in a CUDA graph (pipeline.GraphedDetector) gives the eager results, replay after replay."""
    from videoyolo_b200.pipeline import GraphedDetector
    rng = np.random.RandomState(8)
    for (B, C, size) in [(1, 20, 416), (4, 80, 320)]:
        shapes = [(B, 3 * (5 + C), g, g) for g in oracle.grid_sizes(size)]
        det = GraphedDetector(C, AN, ST, shapes, "cuda")
        for trial in range(3):
            heads = [dev(h) for h in (random_heads(rng, B, C, size) if trial != 1 else trained_heads(rng, B, C, size))]
            out, kept = det(heads)
            torch.cuda.synchronize()
            e_out, e_kept = vy.yolo3_decode_nms(heads, C, AN, ST)
            assert torch.equal(out, e_out) and torch.equal(kept, e_kept)


def test_detect_consume_matches_reference_consumer(vy):
    """clip / valid rows / normalise on the device == the numpy steps of detect() (detect_yolo3.py:226,254-258)."""
    rng = np.random.RandomState(31)
    heads = trained_heads(rng, 3, 20, 416)
    out, _ = vy.yolo3_decode_nms([dev(h) for h in heads], 20, AN, ST)
    clipped, normed, counts = vy.ops.detect_consume(out, 416)
    o = out.cpu().numpy()
    ref_clip, per_image = oracle.detect_consume(o[..., 0:1], o[..., 2:6], 416)
    np.testing.assert_array_equal(clipped.cpu().numpy(), ref_clip)
    for i, (valid, boxes) in enumerate(per_image):
        assert int(counts[i]) == len(valid)
        np.testing.assert_array_equal(valid, np.arange(len(valid)))              # survivors are the first rows
        np.testing.assert_array_equal(normed[i, : len(valid)].cpu().numpy(), boxes)
        assert bool((normed[i, len(valid):] == -1).all())


# ----------------------------------------------------------------------------------- bbox_iou
def test_bbox_iou_against_reference_outputs(vy, golden_dir):
    z = np.load(os.path.join(golden_dir, "bbox_iou_ref.npz"))
    for n in sorted({k[:-4] for k in z.files if k.endswith("_iou")}):
        got = vy.bbox_iou(dev(z[n + "_a"]), dev(z[n + "_b"]), float(z[n + "_off"])).cpu().numpy()
        np.testing.assert_allclose(got, z[n + "_iou"], rtol=1e-12, atol=0, equal_nan=True, err_msg=n)
        got32 = vy.bbox_iou(dev(z[n + "_a"].astype(np.float32)), dev(z[n + "_b"].astype(np.float32)),
                            float(z[n + "_off"])).cpu().numpy()
        np.testing.assert_allclose(got32, z[n + "_iou"], rtol=2e-4, atol=1e-6, equal_nan=True, err_msg=n)
    with pytest.raises(IndexError):
        vy.bbox_iou(torch.zeros(2, 3).cuda(), torch.zeros(2, 4).cuda())


def test_bbox_batch_iou_matches_oracle(vy):
    """"next" row f4: BBoxBatchIOU + max + ignore mask of the dynamic-target step (yolo_target.py:202-204), bit-exact
    against the fp32 restatement (same operation order, un-contracted arithmetic)."""
    rng = np.random.RandomState(8)
    for B, N, M in [(3, 1000, 7), (2, 10647, 50), (1, 333, 300), (4, 1, 1)]:
        a = rng.uniform(0, 416, size=(B, N, 4)).astype(np.float32)
        a[..., 2:] = a[..., :2] + rng.uniform(0, 200, size=(B, N, 2)).astype(np.float32)
        b = rng.uniform(0, 416, size=(B, M, 4)).astype(np.float32)
        b[..., 2:] = b[..., :2] + rng.uniform(1, 200, size=(B, M, 2)).astype(np.float32)
        if M > 2:
            b[:, M // 2:] = -1.0                                           # the reference pads gt rows with -1
        exp = oracle.bbox_batch_iou(a, b)
        got = vy.bbox_batch_iou(dev(a), dev(b))
        np.testing.assert_array_equal(got.cpu().numpy(), exp)
        ious, imax, obj = vy.bbox_batch_iou(dev(a), dev(b), ignore_iou_thresh=0.7)
        np.testing.assert_array_equal(ious.cpu().numpy(), exp)
        np.testing.assert_array_equal(imax.cpu().numpy(), exp.max(axis=-1, keepdims=True))
        np.testing.assert_array_equal(obj.cpu().numpy(), (exp.max(axis=-1, keepdims=True) > 0.7) * -1.0)
        none, imax2, obj2 = vy.bbox_batch_iou(dev(a), dev(b), ignore_iou_thresh=0.7, return_ious=False)
        assert none is None
        np.testing.assert_array_equal(imax2.cpu().numpy(), imax.cpu().numpy())
        np.testing.assert_array_equal(obj2.cpu().numpy(), obj.cpu().numpy())
    with pytest.raises(ValueError):
        vy.bbox_batch_iou(dev(np.zeros((2, 3, 4), np.float32)), dev(np.zeros((3, 3, 4), np.float32)))


def test_empty_batch(vy):
    """B = 0 gives empty results of the right shape from every entry point (no launch)."""
    heads = [dev(np.zeros((0, 75, g, g), np.float32)) for g in oracle.grid_sizes(416)]
    out, kept = vy.yolo3_decode_nms(heads, 20, AN, ST)
    assert tuple(out.shape) == (0, 100, 6) and tuple(kept.shape) == (0, 100)
    assert tuple(vy.yolo3_decode(heads, 20, AN, ST).shape) == (0, 212940, 6)
    assert tuple(vy.box_nms(dev(np.zeros((0, 50, 6), np.float32))).shape) == (0, 50, 6)
    assert tuple(vy.box_nms(dev(np.zeros((2, 0, 6), np.float32))).shape) == (2, 0, 6)
    net = vy.get_yolov3_postprocess(["c"] * 20)
    ids, scores, bboxes = net(*heads)
    assert tuple(ids.shape) == (0, 100, 1) and tuple(bboxes.shape) == (0, 100, 4)
