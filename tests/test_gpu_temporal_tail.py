"""GPU parity of the temporal detector tail (BASELINE configs[2]: K=3 fusion conv + late join + decode + NMS)
through the block surface videoyolo_b200.YOLOV3T / layers.Conv / TemporalPooling / TimeDistributed / Conv1D,
against the CPU oracle chain (oracle.conv_bn_leaky -> temporal_pool -> decode -> box_nms).

Tolerances: the joined tip features within the fusion-conv tolerance (bf16 operands, fp32 accumulation:
|d| <= 1e-2 * max|y|, rtol 2e-2); the final (ids, scores, bboxes) BIT-EXACT against the oracle tail applied to the
rows the GPU decoded from its own head maps (the north-star wording: keep-sets bit-exact on identical boxes)."""
import numpy as np
import pytest

import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vy():
    assert torch.cuda.is_available()
    import videoyolo_b200
    videoyolo_b200._lib.lib()
    return videoyolo_b200


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def randomize_bn(net, rng):
    with torch.no_grad():
        for name, p in list(net.named_parameters()) + list(net.named_buffers()):
            if name.endswith("gamma"):
                p.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, p.shape).astype(np.float32)))
            elif name.endswith("beta") or name.endswith("running_mean"):
                p.copy_(torch.from_numpy(rng.normal(0, 0.2, p.shape).astype(np.float32)))
            elif name.endswith("running_var"):
                p.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, p.shape).astype(np.float32)))


def oracle_cell(cell, x_ncdhw):
    w = bf16_round(cell.weight.detach().cpu().numpy())
    bn = [t.detach().cpu().numpy() for t in (cell.gamma, cell.beta, cell.running_mean, cell.running_var)]
    y = oracle.conv_bn_leaky(x_ncdhw, w, *bn, padding=tuple(k // 2 for k in cell.k3))
    return bf16_round(y)                                  # the kernel stores bf16 activations


def check(got, ref, name):
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-2 * scale, (name, np.abs(got - ref).max(), scale)
    np.testing.assert_allclose(got, ref, rtol=2e-2, atol=1e-2 * scale, err_msg=name)


@pytest.mark.parametrize("join,ctype", [("max", "3"), ("mean", "21"), ("cat", "3")])
def test_yolov3t_tail_matches_oracle_chain(vy, join, ctype):
    rng = np.random.RandomState(len(join) * 7 + len(ctype))
    torch.manual_seed(5)
    B, K, C, size, channels = 2, 3, 20, 160, (128, 64, 64)
    net = vy.YOLOV3T(["c%d" % i for i in range(C)], k=K, k_join_type=join, block_conv_type=ctype, channels=channels).cuda().eval()
    randomize_bn(net, rng)
    xs = [bf16_round(rng.normal(0, 1, size=(B, K, c, g, g))) for c, g in zip(channels, oracle.grid_sizes(size))]
    with torch.no_grad():
        feats = net.tip_features(*[torch.from_numpy(x).cuda() for x in xs])
        ids, scores, bboxes = net(*[torch.from_numpy(x).cuda() for x in xs])
    # ---- tip conv + late join vs the oracle
    for i, x in enumerate(xs):
        y = x.transpose(0, 2, 1, 3, 4)                    # (B, C, K, H, W): swapaxes(1, 2), yolo3.py:256
        for cell in net.tips[i].cells:
            y = oracle_cell(cell, y)
        t = y.transpose(0, 2, 1, 3, 4)                    # back to (B, K, C', H, W)
        if join == "cat":
            ref = t.reshape(B, -1, t.shape[3], t.shape[4])                     # yolo3.py:1136
        else:
            ref = bf16_round(oracle.temporal_pool(t, join).astype(np.float32))  # layers.py:201-205
        check(feats[i].cpu().numpy(), ref, "scale %d %s %s" % (i, join, ctype))
    # ---- output layers: the head maps the forward pass used (Prediction: the library's own kernel with split,
    # fp32-grade weights on the exact bf16 tip, for every join type) against the fp32 prediction conv (yolo3.py:62,157)
    # on the oracle-checked features with the FP32 weights: what the split drops is ~2^-16 relative per product
    with torch.no_grad():
        heads = net.head_maps(*[torch.from_numpy(x).cuda() for x in xs])
    for i, o in enumerate(net.tail.yolo_outputs):
        f64 = feats[i].double().cpu()
        ref_h = torch.nn.functional.conv2d(f64, o.prediction.weight.detach().double().cpu(),
                                           o.prediction.bias.detach().double().cpu()).numpy()
        got_h = heads[i].cpu().numpy()
        scale = np.abs(ref_h).max()
        assert np.abs(got_h - ref_h).max() <= 1e-4 * scale, ("head %d %s" % (i, join), np.abs(got_h - ref_h).max(), scale)
    # ---- NMS tail: exact against the oracle on the GPU's own decoded rows
    AN, ST = oracle.ANCHORS[::-1], oracle.STRIDES[::-1]
    dets = vy.yolo3_decode(heads, C, AN, ST).cpu().numpy()
    o_ids, o_sc, o_bb, o_rec = oracle.yolov3_tail(dets, return_record=True)
    np.testing.assert_array_equal(net.last_kept_rows.cpu().numpy(), o_rec)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(scores.cpu().numpy(), o_sc)
    np.testing.assert_array_equal(bboxes.cpu().numpy(), o_bb)
    # and the decode itself within 1e-5 of the CPU restatement on the same head maps
    ref = oracle.decode_c([h.cpu().numpy() for h in heads], C)
    np.testing.assert_allclose(dets[..., 1], ref[..., 1], rtol=1e-5, atol=1e-12)


def test_layer_blocks(vy):
    """Conv('2') under TimeDistributed == the same 2-D conv per frame; Conv1D == oracle _conv1d."""
    rng = np.random.RandomState(3)
    B, K, Cin, Cout, g = 2, 3, 64, 64, 9
    x = bf16_round(rng.normal(size=(B, K, Cin, g, g)))
    conv2 = vy.Conv("2", Cout, 3, 1, 1, in_channels=Cin).cuda().eval()
    randomize_bn(conv2, rng)
    xp = vy.ops.pack_p(torch.from_numpy(x).cuda(), "NTCHW")
    with pytest.raises(ValueError):
        conv2(xp)                                          # a 2-D cell needs T == 1
    with torch.no_grad():
        y = vy.TimeDistributed(conv2)(xp)
    got = vy.ops.unpack_p(y, "NTCHW").cpu().numpy()
    cell = conv2.cells[0]
    for t in range(K):
        ref = oracle_cell(cell, x[:, t][:, :, None])[:, :, 0]          # (B, Cout, H, W)
        check(got[:, t], ref, "frame %d" % t)
    c1 = vy.Conv1D(Cin, K).cuda().eval()
    with torch.no_grad():
        c1.weight.copy_(torch.from_numpy(rng.uniform(-0.5, 0.5, c1.weight.shape).astype(np.float32)))
    randomize_bn(c1, rng)
    with torch.no_grad():
        z = vy.ops.unpack_p(c1(xp), "NCHW").cpu().numpy()
    ref = oracle.conv1d_bn_leaky(x.transpose(0, 2, 1, 3, 4), c1.weight.detach().cpu().numpy(),
                                 *[t.detach().cpu().numpy() for t in (c1.gamma, c1.beta, c1.running_mean, c1.running_var)])
    check(z, ref[:, :, 0], "conv1d")
    with pytest.raises(ValueError):
        vy.Conv("3", 64, 3, 1, 2, in_channels=64)          # strided cells are not on this path


def test_upsample_concat(vy):
    """_upsample x2 + slice_like + channel concat (layers.py:11-20, yolo3.py:1170-1175) on P-layout data: exact."""
    rng = np.random.RandomState(11)
    for (B, K, Cu, Cr, hu, h) in [(2, 3, 64, 128, 5, 10), (1, 2, 8, 8, 7, 13), (2, 1, 128, 64, 13, 26)]:
        up = bf16_round(rng.normal(size=(B, K, Cu, hu, hu)))
        rt = bf16_round(rng.normal(size=(B, K, Cr, h, h)))
        y = vy.ops.upsample_concat(vy.ops.pack_p(torch.from_numpy(up).cuda(), "NTCHW"),
                                   vy.ops.pack_p(torch.from_numpy(rt).cuda(), "NTCHW"))
        assert (y.B, y.T, y.H, y.W, y.C) == (B, K, h, h, Cu + Cr)
        got = vy.ops.unpack_p(y, "NTCHW").cpu().numpy()
        ref_up = up.repeat(2, axis=-1).repeat(2, axis=-2)[..., :h, :h]                 # _upsample + slice_like
        np.testing.assert_array_equal(got, np.concatenate([ref_up, rt], axis=2))
        d = y.data.float().cpu().numpy()                                               # the border stays zero
        assert not d[:, :, 0].any() and not d[:, :, -1].any() and not d[:, :, :, 0].any() and not d[:, :, :, -1].any()
    with pytest.raises(ValueError):
        vy.ops.upsample_concat(vy.ops.pack_p(torch.zeros(1, 1, 8, 3, 3).cuda(), "NTCHW"),
                               vy.ops.pack_p(torch.zeros(1, 1, 8, 9, 9).cuda(), "NTCHW"))


@pytest.mark.parametrize("join,ctype,widths", [("max", "3", "small"), ("mean", "21", "small"), ("max", "3", "darknet53")])
def test_yolov3t_neck_matches_oracle_chain(vy, join, ctype, widths):
    """"next" row f2: the whole post-backbone part of YOLOV3T (detection blocks, transitions, upsample + concat, late
    join, outputs, NMS; yolo3.py:1126-1206) against the CPU oracle chain.  The block-body outputs (``route``) are compared
    scale by scale with a tolerance that grows with the depth of the bf16 chain; the NMS tail is exact on the GPU's own
    head maps, as for the tail alone."""
    rng = np.random.RandomState(31 + len(join))
    torch.manual_seed(7)
    B, K, C, size = 1, 3, 20, 96
    stage_channels, channels = (128, 64, 64), (64, 64, 64)
    if widths == "darknet53":                              # the widths bench.py's temporal_neck leg runs (wrappers.py:91-103)
        stage_channels, channels, size = (1024, 512, 256), (512, 256, 128), 128
    net = vy.YOLOV3TNeck(["c%d" % i for i in range(C)], k=K, k_join_type=join, block_conv_type=ctype,
                         stage_channels=stage_channels, channels=channels).cuda().eval()
    randomize_bn(net, rng)
    rs = [bf16_round(rng.normal(0, 1, size=(B, K, c, g, g))) for c, g in zip(stage_channels, oracle.grid_sizes(size))]
    with torch.no_grad():
        routes = net.routes(*[torch.from_numpy(r).cuda() for r in rs])
        ids, scores, bboxes = net(*[torch.from_numpy(r).cuda() for r in rs])

    def run_conv(conv, y):                                 # a Conv module = one or two cells (layers.py:135-158)
        for cell in conv.cells:
            y = oracle_cell(cell, y)
        return y

    x = rs[0].transpose(0, 2, 1, 3, 4)                     # NCDHW
    for i, block in enumerate(net.blocks):
        for conv in block.body:
            x = run_conv(conv, x)
        got = vy.ops.unpack_p(routes[i], "NCDHW").cpu().numpy()
        scale = np.abs(x).max()
        assert np.abs(got - x).max() <= 2e-2 * (i + 1) * scale, ("route %d" % i, np.abs(got - x).max(), scale)
        if i + 1 < len(net.blocks):
            t = x.transpose(0, 2, 1, 3, 4)                 # (B, K, C, H, W): TimeDistributed folds K into the batch
            t = t.reshape((B * K,) + t.shape[2:])[:, :, None]          # (B*K, C, 1, H, W)
            t = run_conv(net.transitions[i].model, t)[:, :, 0].reshape((B, K, -1) + t.shape[3:])
            g = rs[i + 1].shape[-1]
            up = t.repeat(2, axis=-1).repeat(2, axis=-2)[..., :g, :g]  # _upsample + slice_like
            x = np.concatenate([up, rs[i + 1]], axis=2).transpose(0, 2, 1, 3, 4)
    # the tail: exact against the oracle on the GPU's own head maps
    with torch.no_grad():
        heads = net.head.head_maps(*routes)
    AN, ST = oracle.ANCHORS[::-1], oracle.STRIDES[::-1]
    dets = vy.yolo3_decode(heads, C, AN, ST).cpu().numpy()
    o_ids, o_sc, o_bb, o_rec = oracle.yolov3_tail(dets, return_record=True)
    np.testing.assert_array_equal(net.last_kept_rows.cpu().numpy(), o_rec)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(scores.cpu().numpy(), o_sc)
    np.testing.assert_array_equal(bboxes.cpu().numpy(), o_bb)


@pytest.mark.parametrize("join", ["max", "mean", "cat"])
def test_yolov3t_neck_early_join(vy, join):
    """Early join (yolo3.py:1107-1123): every stage output joined over the window first ('cat' reshape (0,-3,-2) or
    TemporalPooling), then the 2-D neck.  Routes against the oracle chain on the joined inputs, NMS tail exact on the
    GPU's own head maps."""
    rng = np.random.RandomState(77 + len(join))
    torch.manual_seed(9)
    B, K, C, size = 2, 3, 20, 96
    stage_channels, channels = (128, 64, 64), (64, 64, 64)
    net = vy.YOLOV3TNeck(["c%d" % i for i in range(C)], k=K, k_join_type=join, block_conv_type="2", k_join_pos="early",
                         stage_channels=stage_channels, channels=channels).cuda().eval()
    randomize_bn(net, rng)
    rs = [bf16_round(rng.normal(0, 1, size=(B, K, c, g, g))) for c, g in zip(stage_channels, oracle.grid_sizes(size))]
    with torch.no_grad():
        routes = net.routes(*[torch.from_numpy(r).cuda() for r in rs])
        ids, scores, bboxes = net(*[torch.from_numpy(r).cuda() for r in rs])

    def joined(r):                                         # (B, K, C, H, W) -> (B, C', 1, H, W)
        if join == "cat":
            j = r.reshape(r.shape[0], -1, r.shape[3], r.shape[4])                       # yolo3.py:1110
        else:
            j = bf16_round(oracle.temporal_pool(r, join).astype(np.float32))            # yolo3.py:1112
        return j[:, :, None]

    def run_conv(conv, y):
        for cell in conv.cells:
            y = oracle_cell(cell, y)
        return y

    x = joined(rs[0])
    for i, block in enumerate(net.blocks):
        for conv in block.body:
            x = run_conv(conv, x)
        assert routes[i].T == 1
        got = vy.ops.unpack_p(routes[i], "NCDHW").cpu().numpy()
        scale = np.abs(x).max()
        assert np.abs(got - x).max() <= 2e-2 * (i + 1) * scale, ("route %d" % i, np.abs(got - x).max(), scale)
        if i + 1 < len(net.blocks):
            t = run_conv(net.transitions[i].model, x)
            g = rs[i + 1].shape[-1]
            up = t.repeat(2, axis=-1).repeat(2, axis=-2)[..., :g, :g]
            x = np.concatenate([up, joined(rs[i + 1])], axis=1)
    with torch.no_grad():
        heads = net.head.head_maps(*routes)
    AN, ST = oracle.ANCHORS[::-1], oracle.STRIDES[::-1]
    dets = vy.yolo3_decode(heads, C, AN, ST).cpu().numpy()
    o_ids, o_sc, o_bb, o_rec = oracle.yolov3_tail(dets, return_record=True)
    np.testing.assert_array_equal(net.last_kept_rows.cpu().numpy(), o_rec)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(bboxes.cpu().numpy(), o_bb)


def test_yolov3temporal_tail_over_the_window(vy):
    """YOLOV3Temporal with t_out (yolo3_temporal.py:468, 540-552): TimeDistributed output layers, box_nms over
    (B, T, R, 6) with every leading dimension as batch, slice on axis -2: (B, T, post_nms, .) outputs, each frame's
    result bit-exact against the oracle tail on that frame's decoded rows."""
    rng = np.random.RandomState(12)
    B, T, C, size = 2, 3, 20, 160
    net = vy.YOLOV3Temporal(classes=["c%d" % i for i in range(C)])
    heads = [rng.normal(0, 1, size=(B, T, 3 * (5 + C), g, g)).astype(np.float32) for g in oracle.grid_sizes(size)]
    ids, scores, bboxes = net(*[torch.from_numpy(h).cuda() for h in heads])
    assert ids.shape == (B, T, 100, 1) and scores.shape == (B, T, 100, 1) and bboxes.shape == (B, T, 100, 4)
    assert net.last_kept_rows.shape == (B, T, 100)
    AN, ST = oracle.ANCHORS[::-1], oracle.STRIDES[::-1]
    for b in range(B):
        for t in range(T):
            frame = [torch.from_numpy(h[b:b + 1, t]).cuda() for h in heads]
            dets = vy.yolo3_decode(frame, C, AN, ST).cpu().numpy()
            o_ids, o_sc, o_bb, o_rec = oracle.yolov3_tail(dets, return_record=True)
            np.testing.assert_array_equal(net.last_kept_rows[b, t].cpu().numpy(), o_rec[0])
            np.testing.assert_array_equal(ids[b, t].cpu().numpy(), o_ids[0])
            np.testing.assert_array_equal(scores[b, t].cpu().numpy(), o_sc[0])
            np.testing.assert_array_equal(bboxes[b, t].cpu().numpy(), o_bb[0])
    # a 4-D input is the 2-D model
    i2, s2, b2 = net(*[torch.from_numpy(h[:, 0]).cuda() for h in heads])
    assert i2.shape == (B, 100, 1) and torch.equal(i2, ids[:, 0])


def test_graphed_module_equals_eager(vy):
    """The neck captured in a CUDA graph (pipeline.GraphedModule) replays to exactly the eager results, also on new inputs."""
    from videoyolo_b200.pipeline import GraphedModule
    torch.manual_seed(3)
    net = vy.YOLOV3TNeck(["c%d" % i for i in range(20)], k=3, stage_channels=(128, 64, 64), channels=(64, 64, 64)).cuda().eval()
    mk = lambda: [torch.randn((2, 3, c, g, g), device="cuda") for c, g in zip((128, 64, 64), oracle.grid_sizes(96))]
    g = GraphedModule(net, mk())
    for _ in range(2):
        xs = mk()
        with torch.no_grad():
            ref = [t.clone() for t in net(*xs)]
        out = g(*xs)
        torch.cuda.synchronize()
        for a, b in zip(out, ref):
            assert torch.equal(a, b)


@pytest.mark.parametrize("ctype", ["3", "21"])
def test_fused_max_join_equals_conv_then_pool(vy, ctype):
    """YOLOV3T / YOLOV3TNeck with the late 'max' join done in the tip conv's epilogue (the default) give exactly the
    detections, head maps and tip features of the same network with TemporalPooling as its own kernel."""
    torch.manual_seed(11)
    for net, chans in ((vy.YOLOV3T(["c%d" % i for i in range(20)], k=3, k_join_type="max", block_conv_type=ctype,
                                   channels=(128, 64, 64)).cuda().eval(), (128, 64, 64)),
                       (vy.YOLOV3TNeck(["c%d" % i for i in range(20)], k=3, k_join_type="max", block_conv_type=ctype,
                                       stage_channels=(128, 64, 64), channels=(64, 64, 64)).cuda().eval(), (128, 64, 64))):
        xs = [torch.randn((3, 3, c, g, g), device="cuda") for c, g in zip(chans, oracle.grid_sizes(96))]
        head = net.head if hasattr(net, "head") else net
        assert head.fuse_max_join
        with torch.no_grad():
            fused = [t.clone() for t in net(*xs)]
            head.fuse_max_join = False
            plain = [t.clone() for t in net(*xs)]
            head.fuse_max_join = True
        for a, b in zip(fused, plain):
            assert torch.equal(a, b)
    # head maps and tip features of the tail itself
    net = vy.YOLOV3T(["c%d" % i for i in range(20)], k=3, k_join_type="max", block_conv_type=ctype, channels=(128, 64, 64)).cuda().eval()
    xs = [torch.randn((2, 3, c, g, g), device="cuda") for c, g in zip((128, 64, 64), oracle.grid_sizes(96))]
    with torch.no_grad():
        h1, f1 = net.head_maps(*xs), net.tip_features(*xs)
        net.fuse_max_join = False
        h0, f0 = net.head_maps(*xs), net.tip_features(*xs)
    for a, b in zip(h1 + f1, h0 + f0):
        assert torch.equal(a, b)
