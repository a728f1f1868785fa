"""Per-layer timing of the tcgen05 conv kernels on a backbone's real shapes (diagnostic, GPU box only).
usage: python tools/conv_bench.py [arch] [batch]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200 import ops  # noqa: E402
from rspnet_b200.models import get_model_class  # noqa: E402

import os  # noqa: E402
from rspnet_b200 import _lib  # noqa: E402
if os.environ.get("RSP_WGRAD_DEBUG"):   # 1: never the plane-run wgrad, 2: wherever the geometry allows
    _lib.load().rsp_debug_wgrad(int(os.environ["RSP_WGRAD_DEBUG"]))

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
shapes = []
orig = ops.conv3d_fprop


def spy(desc, x, wp, bias=None, **kw):
    shapes.append((desc, tuple(x.shape)))
    return orig(desc, x, wp, bias, **kw)


ops.conv3d_fprop = spy
net = get_model_class(arch=arch)(num_classes=1).cuda()
with torch.no_grad():
    net.feature_ndhwc(torch.zeros(B, 3, 16, 112, 112, device="cuda"))
ops.conv3d_fprop = orig
del net
torch.cuda.empty_cache()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def t(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tot = {"fprop": 0, "dgrad": 0, "wgrad": 0}
print(f"{'layer':>3} {'in':>22} {'Co':>4} {'k':>7} {'s':>7} {'GF':>7} | {'fprop ms':>8} {'TF/s':>6} | {'dgrad ms':>8} {'TF/s':>6} | {'wgrad ms':>8} {'TF/s':>6}")
for li, (d, xs) in enumerate(shapes):
    x = torch.randn(xs, device="cuda").bfloat16()
    ci_l = 3 if d.Ci == 4 else d.Ci
    w = torch.randn(d.Co, ci_l, d.kt, d.kh, d.kw, device="cuda") * 0.05
    wp = ops.conv3d_pack_weight(d, w, 0)
    y = ops.conv3d_fprop(d, x, wp)
    dy = torch.randn_like(y)
    to, ho, wo = d.out_dims()
    gf = 2.0 * d.N * to * ho * wo * d.Co * ci_l * d.kt * d.kh * d.kw / 1e9
    f = t(lambda: ops.conv3d_fprop(d, x, wp))
    dw = torch.empty_like(w)
    g = t(lambda: ops.conv3d_wgrad(d, x, dy, w.shape, out=dw))
    if li > 0:
        wd = ops.conv3d_pack_weight(d, w, 1)
        dg = t(lambda: ops.conv3d_dgrad(d, dy, wd))
    else:
        dg = 0.0
    tot["fprop"] += f
    tot["dgrad"] += dg
    tot["wgrad"] += g
    print(f"{li:>3} {str(xs):>22} {d.Co:>4} {d.kt}x{d.kh}x{d.kw:<3} {d.st}x{d.sh}x{d.sw:<3} {gf:7.1f} | {f:8.3f} {gf / f:6.1f} | "
          f"{dg:8.3f} {(gf / dg if dg else 0):6.1f} | {g:8.3f} {gf / g:6.1f}", flush=True)
    del x, y, dy, w, wp
print("totals ms:", tot, " step conv ms = 3*fprop + dgrad + wgrad =", 3 * tot["fprop"] + tot["dgrad"] + tot["wgrad"])
