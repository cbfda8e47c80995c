import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, videoyolo_b200 as vy
from videoyolo_b200.pipeline import GraphedModule
dev = torch.device("cuda:0")
torch.manual_seed(9)
for name, net, chans, B in [("neck", vy.YOLOV3TNeck(["c%d" % i for i in range(30)], k=3).to(dev).eval(), (1024, 512, 256), 8),
                            ("tail", vy.YOLOV3T(["c%d" % i for i in range(30)], k=3).to(dev).eval(), (512, 256, 128), 32)]:
    xs = [torch.randn((B, 3, c, g, g), device=dev) for c, g in zip(chans, (13, 26, 52))]
    with torch.no_grad():
        ref = net(*xs)
    g = GraphedModule(net, xs)
    out = g(*xs)
    torch.cuda.synchronize()
    print(name, "graph == eager:", all(torch.equal(a, b) for a, b in zip(out, ref)))
    def t(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    with torch.no_grad():
        print(name, "eager %.3f ms/call, graph %.3f ms/call (back to back, no L2 flush)" % (t(lambda: net(*xs)), t(lambda: g())))
