"""GPU parity of the temporal fusion conv (tcgen05 implicit GEMM) against the CPU oracle
(oracle.conv_bn_leaky: torch CPU fp32 restatement of layers.py:63-89).

Tolerance: operands are bf16, accumulation fp32; the oracle gets the SAME
bf16-rounded inputs and weights, so the only differences are the summation order and the final
rounding of the output to bf16 (2^-9 relative): |d| <= 1e-2 * max|y| and rtol 2e-2 elementwise.
"""
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
    return torch.from_numpy(a).to(torch.bfloat16).float().numpy()


def make_cell(rng, B, T, H, W, Cin, Cout, k3):
    x = bf16_round(rng.normal(0, 1, size=(B, Cin, T, H, W)).astype(np.float32))
    w = bf16_round(rng.uniform(-0.07, 0.07, size=(Cout, Cin) + k3).astype(np.float32))   # MXNet Uniform(0.07)
    gamma = rng.uniform(0.5, 1.5, Cout).astype(np.float32)
    beta = rng.normal(0, 0.2, Cout).astype(np.float32)
    mean = rng.normal(0, 0.2, Cout).astype(np.float32)
    var = rng.uniform(0.5, 2.0, Cout).astype(np.float32)
    return x, w, (gamma, beta, mean, var)


def run_cell(vy, x, w, bn, out_f32=False):
    ops = vy.ops
    dev = "cuda"
    xp = ops.pack_p(torch.from_numpy(x).to(dev), "NCDHW")
    wt = ops.conv_weight(torch.from_numpy(w).to(dev))
    scale, shift = ops.fold_bn(*[torch.from_numpy(v).to(dev) for v in bn])
    y = ops.fusion_conv(xp, wt, scale, shift, 0.1, out_f32=out_f32)
    # the border of the output must be zero again (cells chain without repacking)
    d = y.data.float()
    assert float(d[:, :, 0].abs().max()) == 0 and float(d[:, :, -1].abs().max()) == 0
    assert float(d[:, :, :, 0].abs().max()) == 0 and float(d[:, :, :, -1].abs().max()) == 0
    return y, ops.unpack_p(y, "NCDHW").cpu().numpy()


def check(got, ref, name):
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    assert err.max() <= 1e-2 * scale, (name, err.max(), scale)
    np.testing.assert_allclose(got, ref, rtol=2e-2, atol=1e-2 * scale, err_msg=name)


@pytest.mark.parametrize("B,T,H,W,Cin,Cout,k3", [
    (2, 3, 13, 13, 64, 64, (3, 3, 3)),        # smallest shape, BN=64
    (2, 3, 13, 13, 128, 256, (3, 3, 3)),      # BN=256
    (1, 3, 26, 26, 64, 128, (1, 3, 3)),       # '21' spatial half (layers.py:86), BN=128
    (1, 3, 26, 26, 128, 128, (3, 1, 1)),      # '21' temporal half (layers.py:87)
    (3, 3, 13, 13, 256, 128, (1, 1, 1)),      # the 1x1x1 cells (yolo3.py:229-230)
    (2, 1, 20, 20, 192, 64, (1, 3, 3)),       # 2-D conv after the 'cat' join: T=1, K*C channels
    (1, 5, 10, 10, 64, 64, (3, 3, 3)),        # window of 5
    # the three shapes bench.py times (BASELINE configs[2], K=3 tip convs at 416^2; GEMM K = 13 824 / 6 912 / 3 456)
    (2, 3, 13, 13, 512, 1024, (3, 3, 3)),
    (2, 3, 26, 26, 256, 512, (3, 3, 3)),
    (2, 3, 52, 52, 128, 256, (3, 3, 3)),
    (2, 3, 13, 13, 1024, 512, (1, 1, 1)),     # first body cell of the stride-32 block at Darknet-53 width
    (2, 1, 13, 13, 3072, 128, (1, 1, 1)),     # 'cat' join head conv: K = 3 * C_tip = 3072 (yolo3.py:1135-1136)
])
def test_fusion_conv_matches_oracle(vy, B, T, H, W, Cin, Cout, k3):
    rng = np.random.RandomState(B * 100 + T + H + Cin)
    x, w, bn = make_cell(rng, B, T, H, W, Cin, Cout, k3)
    _, got = run_cell(vy, x, w, bn)
    ref = oracle.conv_bn_leaky(x, w, *bn, padding=tuple(k // 2 for k in k3))
    check(got, ref, str((B, T, H, W, Cin, Cout, k3)))


_HASH_SCRIPT = r"""
import hashlib, json, sys, torch
import videoyolo_b200 as vy
ops = vy.ops
out = {}
for B, g, Cin, Cout, k3 in [(8, 13, 512, 1024, (3, 3, 3)), (8, 26, 256, 512, (3, 3, 3)), (8, 52, 128, 256, (3, 3, 3)),
                             (32, 13, 512, 1024, (3, 3, 3)), (8, 26, 768, 256, (1, 1, 1)), (5, 13, 64, 128, (3, 3, 3))]:
    gen = torch.Generator(device="cuda").manual_seed(B * 1000 + g + Cout)
    x = torch.randn((B, 3, Cin, g, g), generator=gen, device="cuda")
    w = (torch.rand((Cout, Cin) + k3, generator=gen, device="cuda") * 0.14 - 0.07)
    scale = torch.rand(Cout, generator=gen, device="cuda") + 0.5
    shift = torch.randn(Cout, generator=gen, device="cuda") * 0.2
    y = ops.fusion_conv(ops.pack_p(x, "NTCHW"), ops.conv_weight(w), scale, shift, 0.1)
    torch.cuda.synchronize()
    out["%d_%d_%d_%d_%s" % (B, g, Cin, Cout, k3)] = hashlib.sha256(y.data.view(torch.int16).cpu().numpy().tobytes()).hexdigest()
print(json.dumps(out))
"""


def test_fusion_conv_pair_kernel_is_bit_identical_to_one_cta_kernel(vy):
    """The CTA-pair kernel (tcgen05 cta_group::2, 256-row tiles) and the one-CTA kernel accumulate every output in the
    same k order in fp32 TMEM, so at the benchmarked shapes (batch 8 and 32 windows: the plan picks pairs at 13 x 13 and
    52 x 52) the whole P-layout output must be bit-identical with pairs switched off (VY_CONV_CTA2=0; the library reads
    the switch once, hence two processes)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for env_extra in ({"VY_CONV_CTA2": "0"}, {}):
        env = dict(os.environ, PYTHONPATH=root, **env_extra)
        env.pop("VY_CONV_BN", None)
        if not env_extra:
            env.pop("VY_CONV_CTA2", None)
        p = subprocess.run([sys.executable, "-c", _HASH_SCRIPT], env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        res.append(json.loads(p.stdout.strip().splitlines()[-1]))
    assert res[0] == res[1]


def test_fusion_conv_fp32_output_and_chain(vy):
    """conv21d = two chained cells without repacking (layers.py:82-89); last output in fp32."""
    rng = np.random.RandomState(5)
    B, T, H, W, C = 2, 3, 13, 13, 64
    x, w1, bn1 = make_cell(rng, B, T, H, W, C, 128, (1, 3, 3))
    _, w2, bn2 = make_cell(rng, B, T, H, W, 128, 128, (3, 1, 1))
    ops = vy.ops
    xp = ops.pack_p(torch.from_numpy(x).cuda(), "NCDHW")
    s1, h1 = ops.fold_bn(*[torch.from_numpy(v).cuda() for v in bn1])
    s2, h2 = ops.fold_bn(*[torch.from_numpy(v).cuda() for v in bn2])
    y1 = ops.fusion_conv(xp, ops.conv_weight(torch.from_numpy(w1).cuda()), s1, h1)
    y2 = ops.fusion_conv(y1, ops.conv_weight(torch.from_numpy(w2).cuda()), s2, h2, out_f32=True)
    got = ops.unpack_p(y2, "NCDHW").cpu().numpy()
    # oracle: the intermediate is rounded to bf16 like the GPU's
    r1 = bf16_round(oracle.conv_bn_leaky(x, w1, *bn1, padding=(0, 1, 1)))
    ref = oracle.conv_bn_leaky(r1, w2, *bn2, padding=(1, 0, 0))
    check(got, ref, "conv21d")


def test_fusion_conv_nchw_output_equals_unpacked_p_output(vy):
    """vy_fusion_conv_bf16_nchw (the prediction conv's epilogue writes the reference's NCHW head map itself) == the P-layout
    fp32 output of the same cell unpacked, bit for bit; channel counts that are not a multiple of 64 (all_pred = 75, 105,
    255), odd grids, 1x1 and 3x3 kernels."""
    ops = vy.ops
    rng = np.random.RandomState(5)
    for B, H, W, Cin, Cout, C, k in ((3, 13, 13, 128, 128, 75, 1), (2, 19, 19, 192, 256, 255, 1), (2, 10, 10, 64, 128, 105, 1),
                                     (1, 26, 26, 64, 64, 64, 3), (2, 7, 5, 64, 64, 33, 3)):
        x = torch.from_numpy(bf16_round(rng.normal(0, 1, size=(B, Cin, H, W)).astype(np.float32))).cuda()
        w = torch.from_numpy(bf16_round(rng.uniform(-0.07, 0.07, size=(Cout, Cin, k, k)).astype(np.float32))).cuda()
        scale = torch.from_numpy(rng.uniform(0.5, 1.5, Cout).astype(np.float32)).cuda()
        shift = torch.from_numpy(rng.normal(0, 0.2, Cout).astype(np.float32)).cuda()
        xp = ops.pack_p(x, "NCHW")
        wt = ops.conv_weight(w)
        for slope in (1.0, 0.1):
            ref = ops.unpack_p(ops.fusion_conv(xp, wt, scale, shift, slope, out_f32=True), "NCHW", channels=C)
            got = ops.fusion_conv_nchw(xp, wt, scale, shift, slope, channels=C)
            assert got.shape == (B, C, H, W)
            assert torch.equal(got, ref), (B, H, W, Cin, Cout, C, k, slope)


def test_inflated_weights_identity(vy):
    """three_darknet.py:335-347: a clip of identical frames through a 3x3x3 conv whose weights are the
    2-D weights / kt in every temporal tap equals the 2-D conv (interior frames)."""
    rng = np.random.RandomState(8)
    B, T, H, W, Cin, Cout = 1, 3, 13, 13, 64, 64
    x2, w2, bn = make_cell(rng, B, 1, H, W, Cin, Cout, (1, 3, 3))
    x3 = np.repeat(x2, T, axis=2)
    w3 = bf16_round(np.repeat(w2, 3, axis=2) / 4.0)          # /4 is exact in bf16; 3 taps -> 3/4 of the 2-D sum
    _, got3 = run_cell(vy, x3, w3, bn)
    ref = oracle.conv_bn_leaky(x3, w3, *bn, padding=(1, 1, 1))
    check(got3, ref, "inflated")
    # middle frame sees all three taps: conv3d = 0.75 * conv2d before BN
    g2, b2, m2, v2 = bn
    ref2 = oracle.conv_bn_leaky(x2, bf16_round(w2 * 0.75), g2, b2, m2, v2, padding=(0, 1, 1))
    check(got3[:, :, 1:2], ref2, "inflated-vs-2d")


def test_pack_unpack_roundtrip_and_pool(vy):
    rng = np.random.RandomState(2)
    ops = vy.ops
    x = bf16_round(rng.normal(size=(2, 3, 64, 7, 9)).astype(np.float32))     # (B, K, C, H, W)
    xp = ops.pack_p(torch.from_numpy(x).cuda(), "NTCHW")
    assert xp.shape == (2, 64, 3, 7, 9)
    back = ops.unpack_p(xp, "NTCHW").cpu().numpy()
    np.testing.assert_array_equal(back, x)
    back2 = ops.unpack_p(xp, "NCDHW").cpu().numpy()
    np.testing.assert_array_equal(back2, x.transpose(0, 2, 1, 3, 4))
    for typ in ("max", "mean"):
        got = ops.unpack_p(ops.temporal_pool(xp, typ), "NCHW").cpu().numpy()
        ref = bf16_round(oracle.temporal_pool(x, typ).astype(np.float32))
        np.testing.assert_allclose(got, ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("B,K,C,H,W", [(2, 3, 128, 26, 26), (1, 3, 72, 20, 12), (2, 1, 64, 13, 13), (1, 2, 30, 10, 10),
                                       (3, 3, 256, 52, 52), (2, 3, 128, 13, 13), (3, 2, 192, 11, 15), (1, 3, 64, 7, 9)])
def test_pack_layout_and_zero_border(vy, B, K, C, H, W):
    """The P layout bit for bit -- [T][B][H+2][W+2][C] bf16 with a ZERO one-pixel border -- written into a buffer that
    held garbage: even grids take the 16-byte load path, small odd planes of contiguous channels (13 x 13) the flat
    path, channel counts that are multiples of 8 write the border from the pack kernel itself, the others through the
    border kernel (yolo3.py:256-262 swapaxes view, layers.py:76 padding)."""
    ops = vy.ops
    rng = np.random.RandomState(B * 1000 + C)
    x = bf16_round(rng.normal(size=(B, K, C, H, W)).astype(np.float32))
    junk = [torch.full((K * B * (H + 2) * (W + 2) * C,), float("nan"), dtype=torch.bfloat16, device="cuda") for _ in range(4)]
    del junk                                                    # the allocator hands the same memory back to pack_p
    xt = torch.from_numpy(x).cuda()
    for view, layout in ((xt, "NTCHW"), (xt.transpose(1, 2), "NCDHW-strided")):
        if layout == "NTCHW":
            xp = ops.pack_p(view, "NTCHW")
        else:
            xp = ops.pack_p(view.contiguous(), "NCDHW")
        d = xp.data.float().cpu().numpy()                       # (T, B, H+2, W+2, C)
        assert d.shape == (K, B, H + 2, W + 2, C)
        np.testing.assert_array_equal(d[:, :, 1:-1, 1:-1], x.transpose(1, 0, 3, 4, 2))
        for edge in (d[:, :, 0], d[:, :, -1], d[:, :, :, 0], d[:, :, :, -1]):
            assert np.all(edge == 0.0)


@pytest.mark.parametrize("B,T,H,W,Cin,Cout,k3", [(2, 3, 13, 13, 64, 128, (3, 3, 3)), (8, 3, 13, 13, 128, 1024, (3, 3, 3)),
                                                 (3, 3, 26, 26, 64, 256, (3, 1, 1)), (2, 5, 10, 10, 64, 64, (3, 3, 3)),
                                                 (32, 3, 52, 52, 64, 256, (1, 1, 1))])
def test_fusion_conv_max_join_in_the_epilogue(vy, B, T, H, W, Cin, Cout, k3):
    """Conv + TemporalPooling(k, 'max') (layers.py:201-205; the late join of yolo3.py:1134-1138) in one call: the pooled
    frame comes out of the conv's epilogue (bf16x2 max reductions into a frame that starts at -inf) and must be
    bit-identical to the conv followed by the pool kernel -- border zero, every tile shape (one CTA and CTA pairs)."""
    ops = vy.ops
    gen = torch.Generator(device="cuda").manual_seed(B * 100 + H + Cout)
    x = torch.randn((B, T, Cin, H, W), generator=gen, device="cuda")
    w = ops.conv_weight(torch.rand((Cout, Cin) + k3, generator=gen, device="cuda") * 0.14 - 0.07)
    scale = torch.rand(Cout, generator=gen, device="cuda") + 0.5
    shift = torch.randn(Cout, generator=gen, device="cuda") * 0.2
    xp = ops.pack_p(x, "NTCHW")
    ref = ops.temporal_pool(ops.fusion_conv(xp, w, scale, shift, 0.1), "max")
    for _ in range(2):                                            # (twice: the -inf fill is part of every call)
        got = ops.fusion_conv(xp, w, scale, shift, 0.1, pool_max=True)
        assert got.T == 1 and got.data.shape == ref.data.shape
        assert torch.equal(got.data.view(torch.int16), ref.data.view(torch.int16))
    d = got.data.float()
    assert float(d[:, :, 0].abs().max()) == 0 and float(d[:, :, :, -1].abs().max()) == 0


@pytest.mark.parametrize("B,T,H,W,C,n", [(2, 1, 13, 13, 128, 75), (3, 3, 13, 13, 64, 105), (8, 1, 52, 52, 256, 105),
                                         (2, 3, 26, 26, 512, 255)])
def test_prediction_conv_over_the_joined_window_without_materialising_it(vy, B, T, H, W, C, n):
    """The 1x1 prediction conv on a bf16 tip (yolo3.py:62,157; 'cat' join :1135-1136; weight split hi | lo against the
    activation walked twice): addressing frame t / channel block c of the tip from the conv's k loop gives exactly the
    head map of the conv on the materialised ``cat_repeat(x, 2)``."""
    ops = vy.ops
    gen = torch.Generator(device="cuda").manual_seed(B + T + C + n)
    x = ops.pack_p(torch.randn((B, T, C, H, W), generator=gen, device="cuda"), "NTCHW")
    w = ops.split_weight(torch.rand((n, T * C, 1, 1), generator=gen, device="cuda") * 0.14 - 0.07, 2)
    scale = torch.ones(w.shape[0], device="cuda")
    shift = torch.randn(w.shape[0], generator=gen, device="cuda")
    ref = ops.fusion_conv_nchw(ops.cat_repeat(x, 2), w, scale, shift, slope=1.0, channels=n)
    got = ops.fusion_conv_nchw_joined(x, w, scale, shift, rep=2, slope=1.0, channels=n)
    assert got.shape == (B, n, H, W)
    assert torch.equal(got, ref)


def test_temporal_dwconv_matches_oracle(vy):
    """_conv1d temporal merge (layers.py:50-60, h_darknet.py:97-119): window of 3 frames, C=32."""
    rng = np.random.RandomState(21)
    ops = vy.ops
    for (B, C, T, H, W) in [(2, 32, 3, 20, 20), (3, 64, 5, 7, 9)]:
        x = bf16_round(rng.normal(0, 1, size=(B, C, T, H, W)).astype(np.float32))
        w = rng.uniform(-0.5, 0.5, size=(C, 1, T, 1, 1)).astype(np.float32)
        gamma = rng.uniform(0.5, 1.5, C).astype(np.float32)
        beta = rng.normal(0, 0.2, C).astype(np.float32)
        mean = rng.normal(0, 0.2, C).astype(np.float32)
        var = rng.uniform(0.5, 2.0, C).astype(np.float32)
        xp = ops.pack_p(torch.from_numpy(x).cuda(), "NCDHW")
        scale, shift = ops.fold_bn(*[torch.from_numpy(v).cuda() for v in (gamma, beta, mean, var)])
        y = ops.temporal_dwconv(xp, torch.from_numpy(w).cuda(), scale, shift, 0.1)
        d = y.data.float()
        assert float(d[:, :, 0].abs().max()) == 0 and float(d[:, :, :, -1].abs().max()) == 0     # border stays zero
        got = ops.unpack_p(y, "NCDHW").cpu().numpy()
        ref = oracle.conv1d_bn_leaky(x, w, gamma, beta, mean, var)
        check(got, ref, "conv1d %s" % ((B, C, T, H, W),))


def test_fusion_conv_rejects_bad_shapes(vy):
    ops = vy.ops
    xp = ops.pack_p(torch.zeros(1, 48, 3, 5, 5).cuda(), "NCDHW")
    w = torch.zeros(64, 3, 3, 3, 48, dtype=torch.bfloat16).cuda()
    with pytest.raises(vy._lib.VyoloError):
        ops.fusion_conv(xp, w, torch.ones(64).cuda(), torch.zeros(64).cuda())
