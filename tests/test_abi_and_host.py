"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol
include/vyolo.h declares (no compute calls without a GPU), the Python surface refuses CPU tensors
(no fallback), and the frame-sharding logic works across 2 gloo ranks."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vyolo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vy_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from videoyolo_b200 import _lib, build
    so = build.build()
    assert os.path.exists(so)
    names = _declared()
    assert len(names) >= 10
    L = ctypes.CDLL(so)
    for n in names:
        assert hasattr(L, n), "libvyolo.so does not export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().vy_version() == 1
    assert _lib.lib().vy_last_error() is not None


def test_library_is_sm100a_with_native_kernels():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "videoyolo_b200", "libvyolo.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_argument_validation_without_gpu():
    """Host-side checks run before any CUDA call, so they are testable on the CPU box."""
    from videoyolo_b200 import _lib
    L = _lib.lib()
    rc = L.vy_box_nms_f32(None, 1, 10, 6, 0.5, 0.0, -1, 2, 1, 0, -1, 0, 0, 0, 10, None, None, None, 0, None)
    assert rc == -1 and b"bad data" in L.vy_last_error()
    fake = ctypes.c_void_p(256)
    rc = L.vy_box_nms_f32(fake, 1, 10, 5, 0.5, 0.0, -1, 2, 1, 0, -1, 0, 0, 0, 10, fake, None, None, 0, None)
    assert rc == -1 and b"outside the row" in L.vy_last_error()          # coord_start+4 > W
    rc = L.vy_box_nms_f32(fake, 1, 10, 6, 0.5, 0.0, 5, 2, 1, 0, -1, 0, 0, 0, 10, fake, None, None, 0, None)
    assert rc == -3                                                        # workspace missing
    n = L.vy_box_nms_workspace_bytes(64, 1819440, 6, 400)
    assert 0 < n < (1 << 30)
    H = (ctypes.c_int * 3)(19, 38, 76)
    n2 = L.vy_decode_nms_workspace_bytes(H, H, 3, 64, 3, 80, 0, 400)
    assert 0 < n2 < (1 << 28)
    assert L.vy_decode_nms_workspace_bytes(H, H, 3, 64, 3, 80, 0, 5000) == 0   # K > 1024: unfused path


def test_no_cpu_fallback():
    import videoyolo_b200 as vy
    with pytest.raises(RuntimeError, match="no CPU"):
        vy.box_nms(torch.zeros(1, 4, 6))
    with pytest.raises(RuntimeError, match="no CPU"):
        vy.yolo3_decode([torch.zeros(1, 75, 13, 13)], 20, [vy.ANCHORS[2]], [32])
    with pytest.raises(TypeError):
        vy.bbox_iou(np.zeros((2, 4)), np.zeros((2, 4)))
    # the product never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "videoyolo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_block_surface_constants():
    import videoyolo_b200 as vy
    net = vy.get_yolov3_postprocess(["a", "b"])
    assert (net.nms_thresh, net.nms_topk, net.post_nms) == (0.45, 400, 100)          # yolo3.py:394-396
    assert [o._stride for o in net.yolo_outputs] == [32, 16, 8]                       # reversed, :416-417
    assert net.yolo_outputs[0]._anchors == [116, 90, 156, 198, 373, 326]
    net.set_nms(0.5, 200, 50)
    assert (net.nms_thresh, net.nms_topk, net.post_nms) == (0.5, 200, 50)
    assert net.classes == ["a", "b"] and net.num_class == 2


def test_split_frames_matches_gluon_split_data():
    from videoyolo_b200.parallel import shard, split_frames, window_halo
    assert split_frames(256, 8) == [(i * 32, (i + 1) * 32) for i in range(8)]
    assert split_frames(10, 4) == [(0, 2), (2, 4), (4, 6), (6, 10)]        # last slice takes the remainder
    assert split_frames(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert shard(10, 3, 4) == (6, 10)
    for n in range(0, 40):
        for w in range(1, 9):
            b = split_frames(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    assert window_halo(3, 1) == 1 and window_halo(5, 2) == 4


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import oracle
from videoyolo_b200.parallel import shard, gather_detections
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
N, C = 5, 6
rng = np.random.RandomState(0)
heads = [rng.normal(size=(N, 3 * (5 + C), g, g)).astype(np.float32) for g in (2, 4, 8)]
s, e = shard(N, rank, 2)
# each rank post-processes only its shard (the oracle stands in for the GPU path on the CPU box)
ids, sc, bb = oracle.yolov3_postprocess([h[s:e] for h in heads], C, post_nms=20)
got = gather_detections([torch.from_numpy(np.ascontiguousarray(t)) for t in (ids, sc, bb)], N)
full = oracle.yolov3_postprocess(heads, C, post_nms=20)
for g, f in zip(got, full):
    assert g.shape[0] == N and np.array_equal(g.numpy(), f), "sharded != unsharded"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_frame_sharding_world_size_2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_contract_and_no_cpu_fallback():
    """`bench.py --impl reference` is CPU-only (the oracle port on the host cores) and prints ONE JSON line with the
    contract keys; the native arm refuses to run without a CUDA device instead of falling back."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and d["gpu_launches"] == 0
    assert "workload" in d["config"]
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                           capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode != 0 and "CUDA device" in (r.stderr + r.stdout)
