"""CPU tests: the oracle against the golden vectors and against independent restatements.

The oracle (oracle/) is the checker for every GPU parity test, so it is pinned here first:
MXNet's published box_nms answers, the reference's own bbox_iou outputs, a python twin,
torchvision's nms, and the closed-form decode identities of SURVEY.md 8(c).
"""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle


# ----------------------------------------------------------------------------- decode vs the reference's own code
from conftest import DECODE_REF_NAMES, assert_decode_close, load_decode_ref  # noqa: E402


@pytest.mark.parametrize("name", DECODE_REF_NAMES)
def test_decode_oracle_matches_reference_execution(name):
    """oracle.decode_c against what the REFERENCE'S OWN YOLOOutputV3.hybrid_forward (yolo3.py:130-199,
    agnostic :184-188) + concat (:523) produced on the same head maps (tests/golden/make_golden.py,
    reference_decode()).  Same rows in the same order; values to a few ulps (libm expf vs the
    correctly rounded exp of the shim, then one rounding per fp32 op in identical order)."""
    z = load_decode_ref(name)
    mine = oracle.decode_c(z["heads"], z["C"], agnostic=z["agnostic"])
    assert mine.shape == (z["B"], z["R"], 6)
    assert_decode_close(mine[:, z["rows"]], z["dets_rows"], score_ulps=4, box_eps=2)
    # every row, not just the stored ones: float64 column sums of the whole tensor
    np.testing.assert_allclose(mine.astype(np.float64).sum(axis=1), z["col_sums"], rtol=1e-7)
    # numpy twin, scale by scale
    if not z["agnostic"]:
        tw = np.concatenate([oracle.decode_numpy(h, a, s, z["C"]) for h, a, s in
                             zip(z["heads"], oracle.ANCHORS[::-1], oracle.STRIDES[::-1])], axis=1)
        assert_decode_close(tw[:, z["rows"]], z["dets_rows"], score_ulps=8, box_eps=4)


@pytest.mark.parametrize("name", DECODE_REF_NAMES)
def test_tail_oracle_matches_reference_execution(name):
    """oracle.yolov3_postprocess against the (ids, scores, bboxes) the reference's own
    YOLOV3.hybrid_forward (yolo3.py:448-534: concat, box_nms call with its arguments, slice_axis
    post_nms, split) returned.  The box_nms step inside that run was oracle.box_nms_c (MXNet's
    operator is not in /root/reference), so this pins the tail's plumbing, the argument values and
    the decode feeding it -- not the NMS arithmetic (that: MXNet's documented vectors below)."""
    z = load_decode_ref(name)
    ids, scores, bboxes = oracle.yolov3_postprocess(z["heads"], z["C"], agnostic=z["agnostic"])
    assert ids.shape == z["ids"].shape == (z["B"], 100, 1)
    same = (ids == z["ids"]).all(axis=(1, 2))
    assert same.all(), "kept classes differ in frames %s" % np.nonzero(~same)[0]
    got = np.concatenate([ids, scores, bboxes], axis=-1)
    ref = np.concatenate([z["ids"], z["scores"], z["bboxes"]], axis=-1)
    live = ref[..., 0] >= 0
    assert_decode_close(got[live], ref[live], score_ulps=4, box_eps=2)
    assert (got[~live] == -1).all()


# ----------------------------------------------------------------------------- box_nms
def _cases(golden_dir=os.path.join(os.path.dirname(__file__), "golden")):
    with open(os.path.join(golden_dir, "box_nms_mxnet_doc.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
@pytest.mark.parametrize("impl", ["c", "py"])
def test_box_nms_mxnet_known_answers(case, impl):
    fn = oracle.box_nms_c if impl == "c" else oracle.box_nms_py
    out, rec = fn(np.array(case["data"], dtype=np.float32), return_record=True, **case["args"])
    exp = np.array(case["expected"], dtype=np.float32)
    if case.get("approx"):
        np.testing.assert_allclose(out, exp, rtol=1e-5, atol=1e-6)
    else:
        np.testing.assert_array_equal(out, exp)
    np.testing.assert_array_equal(rec, np.array(case["kept"]))


def _random_dets(rng, B, R, n_cls, quant=None, scale=100.0):
    xy = rng.uniform(0, scale, size=(B, R, 2))
    wh = rng.uniform(-0.05 * scale, 0.5 * scale, size=(B, R, 2))     # a few inverted boxes -> area 0
    sc = rng.uniform(-0.1, 1.0, size=(B, R, 1))
    if quant:
        sc = np.round(sc * quant) / quant                            # force score ties
    ids = rng.randint(0, n_cls, size=(B, R, 1)).astype(np.float64)
    return np.concatenate([ids, sc, xy, xy + wh], axis=-1).astype(np.float32)


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), R=st.integers(1, 70), n_cls=st.integers(1, 4),
       topk=st.sampled_from([-1, 1, 5, 30, 1000]), force=st.booleans(),
       thr=st.sampled_from([0.1, 0.45, 0.7]), valid=st.sampled_from([-1.0, 0.0, 0.01, 0.5]),
       quant=st.sampled_from([None, 8, 64]), id_index=st.sampled_from([0, -1]),
       fmt=st.sampled_from(["corner", "center"]))
def test_box_nms_c_equals_python_twin(seed, R, n_cls, topk, force, thr, valid, quant, id_index, fmt):
    rng = np.random.RandomState(seed)
    d = _random_dets(rng, 2, R, n_cls, quant)
    if fmt == "center":
        c = d.copy()
        c[..., 2:4] = (d[..., 2:4] + d[..., 4:6]) / 2
        c[..., 4:6] = d[..., 4:6] - d[..., 2:4]
        d = c
    kw = dict(overlap_thresh=thr, valid_thresh=valid, topk=topk, id_index=id_index,
              force_suppress=force, in_format=fmt, out_format=fmt, return_record=True)
    oc, rc = oracle.box_nms_c(d, **kw)
    op, rp = oracle.box_nms_py(d, **kw)
    np.testing.assert_array_equal(rc, rp)
    np.testing.assert_array_equal(oc, op)


def test_box_nms_against_torchvision_per_class():
    import torch
    import torchvision
    rng = np.random.RandomState(7)
    d = _random_dets(rng, 1, 400, 5)
    d[..., 4:6] = np.maximum(d[..., 4:6], d[..., 2:4] + 1.0)         # torchvision needs x2>x1
    d[..., 1] = rng.permutation(400).reshape(1, 400) / 400.0 + 0.001  # distinct scores
    out, rec = oracle.box_nms_c(d, overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=0,
                                return_record=True)
    keep = []
    for c in range(5):
        rows = np.nonzero(d[0, :, 0] == c)[0]
        k = torchvision.ops.nms(torch.from_numpy(d[0, rows, 2:6]), torch.from_numpy(d[0, rows, 1]), 0.45)
        keep += list(rows[k.numpy()])
    keep = sorted(keep, key=lambda r: -d[0, r, 1])
    assert list(rec[0][rec[0] >= 0]) == keep


def test_box_nms_edge_cases():
    # empty candidate set -> all -1
    d = _random_dets(np.random.RandomState(0), 2, 9, 3)
    out, rec = oracle.box_nms_c(d, valid_thresh=5.0, return_record=True)
    assert (out == -1).all() and (rec == -1).all()
    # strict '>' on valid_thresh and stable ties (lower row first)
    d = np.zeros((1, 4, 6), np.float32)
    d[0, :, 1] = [0.5, 0.7, 0.7, 0.01]
    d[0, :, 2:6] = [[0, 0, 1, 1], [10, 10, 11, 11], [20, 20, 21, 21], [30, 30, 31, 31]]
    out, rec = oracle.box_nms_c(d, valid_thresh=0.01, id_index=0, return_record=True)
    assert list(rec[0]) == [1, 2, 0, -1]
    # identical boxes, same class: IoU = 1 > thr -> only the first survives; 0-area pair: 0/0 = NaN keeps
    d = np.zeros((1, 3, 6), np.float32)
    d[0, :, 1] = [0.9, 0.8, 0.7]
    d[0, :2, 2:6] = [5, 5, 9, 9]
    d[0, 2, 2:6] = [3, 3, 3, 3]
    out, rec = oracle.box_nms_c(d, overlap_thresh=0.45, id_index=0, return_record=True)
    assert list(rec[0]) == [0, 2, -1]
    # topk cuts BEFORE suppression, globally across classes
    d = _random_dets(np.random.RandomState(3), 1, 50, 3)
    o1, r1 = oracle.box_nms_c(d, topk=7, id_index=0, return_record=True)
    order = np.argsort(-d[0, :, 1], kind="stable")[:7]
    assert set(r1[0][r1[0] >= 0]) <= set(order)


# ----------------------------------------------------------------------------- decode
@pytest.mark.parametrize("C,size,agnostic", [(20, 416, False), (3, 96, False), (30, 320, True), (1, 64, False)])
def test_decode_c_equals_numpy_graph(C, size, agnostic):
    rng = np.random.RandomState(C + size)
    grids = oracle.grid_sizes(size)
    heads = [rng.normal(0, 1.5, size=(2, 3 * (5 + C), g, g)).astype(np.float32) for g in grids]
    dets = oracle.decode_c(heads, C, agnostic=agnostic)
    ref = np.concatenate([oracle.decode_numpy(h, a, s, C, agnostic) for h, a, s in
                          zip(heads, oracle.ANCHORS[::-1], oracle.STRIDES[::-1])], axis=1)
    assert dets.shape == ref.shape
    np.testing.assert_array_equal(dets[..., 0], ref[..., 0])
    # glibc expf vs numpy's SIMD exp differ by <= a few ulp; the centre/half-size cancellation
    # in x1 = cx - w/2 needs an absolute term scaled by the coordinate magnitude
    np.testing.assert_allclose(dets[..., 1], ref[..., 1], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(dets[..., 2:], ref[..., 2:], rtol=2e-6, atol=2e-6 * 4 * size)


def test_decode_identities_zero_logits():
    """zero logits => sigma=0.5, wh=anchor, centre=(x+0.5)*stride, score=0.25."""
    C, size = 4, 64
    grids = oracle.grid_sizes(size)
    heads = [np.zeros((1, 3 * (5 + C), g, g), np.float32) for g in grids]
    dets = oracle.decode_c(heads, C)
    off = 0
    for g, stride, anc in zip(grids, oracle.STRIDES[::-1], oracle.ANCHORS[::-1]):
        n_s = g * g * 3
        blk = dets[0, off:off + C * n_s].reshape(C, g, g, 3, 6)
        for c in range(C):
            assert (blk[c, ..., 0] == c).all() and (blk[c, ..., 1] == 0.25).all()
        for a in range(3):
            w, h = anc[2 * a], anc[2 * a + 1]
            ys, xs = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
            np.testing.assert_allclose(blk[0, :, :, a, 2], (xs + 0.5) * stride - w / 2.0, rtol=1e-6)
            np.testing.assert_allclose(blk[0, :, :, a, 3], (ys + 0.5) * stride - h / 2.0, rtol=1e-6)
            np.testing.assert_allclose(blk[0, :, :, a, 4], (xs + 0.5) * stride + w / 2.0, rtol=1e-6)
            np.testing.assert_allclose(blk[0, :, :, a, 5], (ys + 0.5) * stride + h / 2.0, rtol=1e-6)
        off += C * n_s
    assert off == dets.shape[1]


def test_row_order_formula():
    """global row r = C*sum_{s'<s} n_s' + c*n_s + pos*A + a  (yolo3.py:191-197, :523)."""
    C, size = 3, 64
    grids = oracle.grid_sizes(size)
    rng = np.random.RandomState(5)
    heads = [rng.normal(size=(1, 3 * (5 + C), g, g)).astype(np.float32) for g in grids]
    dets = oracle.decode_c(heads, C)
    s, c, y, x, a = 1, 2, 3, 1, 2
    g = grids[s]
    r = C * grids[0] ** 2 * 3 + c * g * g * 3 + (y * g + x) * 3 + a
    t = heads[s][0].reshape(3, 5 + C, g, g)[a, :, y, x]
    sig = lambda v: 1 / (1 + np.exp(-v))
    assert dets[0, r, 0] == c
    np.testing.assert_allclose(dets[0, r, 1], sig(t[5 + c]) * sig(t[4]), rtol=1e-5)
    cx = (sig(t[0]) + x) * oracle.STRIDES[::-1][s]
    w = np.exp(t[2]) * oracle.ANCHORS[::-1][s][2 * a]
    np.testing.assert_allclose(dets[0, r, 2], cx - w / 2, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("name", ["voc416_random", "vid320_trained", "coco_small_trained"])
def test_postproc_regression_fixture(name, golden_dir):
    z = np.load(os.path.join(golden_dir, "postproc_regress_%s.npz" % name))
    ids, scores, bboxes, rec = oracle.yolov3_postprocess([z["h0"], z["h1"], z["h2"]], int(z["C"]),
                                                         return_record=True)
    np.testing.assert_array_equal(rec, z["kept_rows"])
    np.testing.assert_array_equal(ids, z["ids"])
    np.testing.assert_array_equal(scores, z["scores"])
    np.testing.assert_array_equal(bboxes, z["bboxes"])


def test_temporal_leading_dims_are_batch():
    """yolo3_temporal.py:545: box_nms on (B,T,N,6) treats B*T as images."""
    d = _random_dets(np.random.RandomState(11), 6, 40, 3).reshape(2, 3, 40, 6)
    a = oracle.box_nms_c(d, overlap_thresh=0.45, valid_thresh=0.01, topk=10, id_index=0)
    b = oracle.box_nms_c(d.reshape(6, 40, 6), overlap_thresh=0.45, valid_thresh=0.01, topk=10, id_index=0)
    np.testing.assert_array_equal(a.reshape(6, 40, 6), b)


# ----------------------------------------------------------------------------- bbox_iou
def test_bbox_iou_against_reference_outputs(golden_dir):
    z = np.load(os.path.join(golden_dir, "bbox_iou_ref.npz"))
    names = sorted({k[:-4] for k in z.files if k.endswith("_iou")})
    assert names
    for n in names:
        got = oracle.bbox_iou(z[n + "_a"], z[n + "_b"], float(z[n + "_off"]))
        np.testing.assert_allclose(got, z[n + "_iou"], rtol=1e-12, atol=0, equal_nan=True)
    with pytest.raises(IndexError):
        oracle.bbox_iou(np.zeros((2, 3)), np.zeros((2, 4)))


# ----------------------------------------------------------------------------- fusion conv
def test_conv_inflation_identity():
    """three_darknet.py:335-347 property: a clip of identical frames through a (kt,3,3) conv whose
    weights are the 2-D weights / kt ... equals the 2-D conv (interior frames; zero temporal pad)."""
    rng = np.random.RandomState(2)
    x2 = rng.normal(size=(1, 8, 6, 6)).astype(np.float32)
    w2 = rng.uniform(-0.07, 0.07, size=(16, 8, 3, 3)).astype(np.float32)
    bn = (np.ones(16), np.zeros(16), np.zeros(16), np.ones(16))
    y2 = oracle.conv_bn_leaky(x2, w2, *bn, padding=1)
    x3 = np.repeat(x2[:, :, None], 5, axis=2)
    w3 = np.repeat(w2[:, :, None], 3, axis=2) / 3.0
    y3 = oracle.conv_bn_leaky(x3, w3, *bn, padding=1)
    np.testing.assert_allclose(y3[:, :, 2], y2, rtol=1e-4, atol=1e-5)
    assert oracle.temporal_pool(np.stack([x2, 2 * x2], 1), "max").shape == x2.shape


def test_conv1d_temporal_merge_closed_form():
    """_conv1d (layers.py:50-60): depthwise (T,1,1) conv over a window == per-channel weighted sum of frames."""
    rng = np.random.RandomState(3)
    x = rng.normal(size=(2, 8, 3, 4, 5)).astype(np.float32)
    w = rng.normal(size=(8, 1, 3, 1, 1)).astype(np.float32)
    g, b, m, v = (rng.uniform(0.5, 1.5, 8).astype(np.float32), rng.normal(size=8).astype(np.float32),
                  rng.normal(size=8).astype(np.float32), rng.uniform(0.5, 2, 8).astype(np.float32))
    y = oracle.conv1d_bn_leaky(x, w, g, b, m, v)
    assert y.shape == (2, 8, 1, 4, 5)
    s = (x * w.reshape(1, 8, 3, 1, 1)).sum(axis=2, keepdims=True)
    sh = (1, 8, 1, 1, 1)
    ref = (s - m.reshape(sh)) / np.sqrt(v.reshape(sh) + 1e-5) * g.reshape(sh) + b.reshape(sh)
    ref = np.where(ref > 0, ref, 0.1 * ref)
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5)


def test_bbox_batch_iou_against_plain_loops():
    """oracle.bbox_batch_iou (gluoncv BBoxBatchIOU restated, yolo_target.py:171,202) against the formula written out
    box by box in fp32, including the -1 padding rows the reference's gt tensors carry."""
    rng = np.random.RandomState(5)
    a = rng.uniform(0, 100, size=(2, 9, 4)).astype(np.float32)
    a[..., 2:] += a[..., :2]
    b = rng.uniform(0, 100, size=(2, 4, 4)).astype(np.float32)
    b[..., 2:] += b[..., :2]
    b[:, -1] = -1.0
    got = oracle.bbox_batch_iou(a, b)
    f = np.float32
    for i in range(2):
        for n in range(9):
            for m in range(4):
                p, q = a[i, n], b[i, m]
                iw = min(max(f(min(p[2], q[2]) - max(p[0], q[0])), f(0)), f(65504.0))
                ih = min(max(f(min(p[3], q[3]) - max(p[1], q[1])), f(0)), f(65504.0))
                inter = f(iw * ih)
                union = f(f(f((p[2] - p[0]) * (p[3] - p[1])) + f((q[2] - q[0]) * (q[3] - q[1]))) - inter)
                assert got[i, n, m] == f(inter / f(union + f(1e-15)))
    assert got.shape == (2, 9, 4) and (got[:, :, -1] == 0).all()          # a padding box overlaps nothing
    same = oracle.bbox_batch_iou(a, a)
    np.testing.assert_allclose(same[0].diagonal(), 1.0, rtol=1e-6)


def test_decode_torch_graph_equals_numpy_graph():
    """the torch-CPU graph restatement (bench.py's graph-faithful CPU figure) against the numpy one: same rows, same order"""
    rng = np.random.RandomState(2)
    C, size = 7, 96
    heads = [rng.normal(size=(2, 3 * (5 + C), g, g)).astype(np.float32) for g in oracle.grid_sizes(size)]
    ref = np.concatenate([oracle.decode_numpy(h, a, s, C) for h, a, s in zip(heads, oracle.ANCHORS[::-1], oracle.STRIDES[::-1])], axis=1)
    got = oracle.decode_torch_graph(heads, C)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got[..., 0], ref[..., 0])
    np.testing.assert_allclose(got[..., 1:], ref[..., 1:], rtol=1e-5, atol=1e-5 * size)      # libm vs torch exp


# ----------------------------------------------------------------------------- device-side consumers vs the reference's own code
def _consumer_ref(golden_dir=os.path.join(os.path.dirname(__file__), "golden")):
    return np.load(os.path.join(golden_dir, "consumer_ref.npz"))


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_hierarchical_nms_oracle_matches_reference_execution(case):
    """oracle.hierarchical_nms against the reference's own `hierarchical_nms` / `iou` (detect_yolo3.py:712-789), executed
    from its source on seeded float32 detections (tests/golden/make_golden.py:reference_consumers)."""
    z = _consumer_ref()
    ov, conf = z["hier_%s_args" % case]
    for i, boxes in enumerate(z["hier_%s_in" % case]):
        got = oracle.hierarchical_nms(boxes, z["hier_%s_lifted" % case], z["hier_%s_branch" % case], ov, conf)
        n = int(z["hier_%s_count" % case][i])
        assert len(got) == n
        np.testing.assert_array_equal(got, z["hier_%s_out" % case][i, :n])


@pytest.mark.parametrize("with_diff", [0, 1])
def test_voc_match_oracle_matches_reference_execution(with_diff):
    """oracle.voc_match against what the reference's own VOCMApMetric.update (metrics/pascalvoc.py:85-184) accumulated."""
    z = _consumer_ref()
    B = z["voc_pb"].shape[0]
    score, match, npos = {}, {}, {}
    for b in range(B):
        l, s, m, n = oracle.voc_match(z["voc_pb"][b], z["voc_pl"][b], z["voc_ps"][b], z["voc_gtb"][b], z["voc_gtl"][b],
                                      z["voc_gtd"][b] if with_diff else None)
        for c in np.unique(l):
            score.setdefault(int(c), []).extend(s[l == c]); match.setdefault(int(c), []).extend(m[l == c])
        for c, v in n.items():
            npos[c] = npos.get(c, 0) + v
            score.setdefault(c, []); match.setdefault(c, [])
    classes = list(z["voc%d_classes" % with_diff])
    assert sorted(set(npos) | set(score)) == classes
    np.testing.assert_array_equal([npos.get(c, 0) for c in classes], z["voc%d_npos" % with_diff])
    np.testing.assert_array_equal(np.concatenate([np.array(score[c], dtype=np.float32) for c in classes]), z["voc%d_score" % with_diff])
    # match values: identical wherever the order is defined.  Among EQUAL scores of one (image, class) the reference's
    # order is whatever numpy's default (unstable, platform-dependent) argsort returns (metrics/pascalvoc.py:141), and
    # the greedy TP assignment follows that order; the oracle and the CUDA kernel define it (later row first), so within
    # a run of equal scores the match values are compared as multisets.  (The fixture quantises scores to force ties.)
    got_m = np.concatenate([np.array(match[c], dtype=np.int32) for c in classes])
    ref_m, ref_s = z["voc%d_match" % with_diff], z["voc%d_score" % with_diff]
    cls_of = np.repeat(np.arange(len(classes)), z["voc%d_len" % with_diff])
    i = 0
    while i < len(ref_m):
        j = i
        while j + 1 < len(ref_m) and ref_s[j + 1] == ref_s[i] and cls_of[j + 1] == cls_of[i]:
            j += 1
        assert sorted(got_m[i:j + 1]) == sorted(ref_m[i:j + 1]), (i, j)
        i = j + 1
    assert (got_m == ref_m).mean() > 0.97


def test_anchor_match_known_answers():
    """yolo_target.py:86-94 restated (oracle.anchor_match): a box of exactly an anchor's size matches it with IoU 1
    wherever it lies; IoU of zero-centred boxes = min-area / max-area when one contains the other; padding rows give 0."""
    an = np.array([[10, 13], [16, 30], [33, 23]], np.float32)
    gt = np.array([[[100, 200, 116, 230], [5, 5, 6, 6], [-1, -1, -1, -1], [0, 0, 66, 46]]], np.float32)
    m, iou = oracle.anchor_match(gt, an)
    assert m.tolist() == [[1, 0, 0, 2]]
    assert iou.shape == (1, 3, 4) and iou[0, 1, 0] == 1.0
    np.testing.assert_allclose(iou[0, :, 1], [1 / 130, 1 / 480, 1 / 759], rtol=1e-6)       # 1x1 box inside every anchor
    assert (iou[0, :, 2] == 0).all()                                                     # the -1 row: zero extent
    np.testing.assert_allclose(iou[0, 2, 3], 0.25, rtol=1e-6)                            # anchor 2 scaled by 2 per side
    np.testing.assert_allclose(oracle.box_iou_mxnet(np.array([[0, 0, 2, 2]], np.float32), np.array([[1, 1, 3, 3]], np.float32)),
                               [[1 / 7]], rtol=1e-6)
