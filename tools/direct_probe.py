"""Timing experiments on the direct conv kernel (GPU box): which part of the pipeline bounds it?
The library's diagnostic switches make the kernel skip one part at a time (results are then garbage, only time counts)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def t(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


cases = [("R3D-18 layer1 64->64 @8x28x28", (64, 8, 28, 28, 64), 64), ("C3D conv2 64->128 @16x56x56", (64, 16, 56, 56, 64), 128),
         ("C3D conv3b 256->256 @8x28x28", (64, 8, 28, 28, 256), 256)]
for name, xs, co in cases:
    x = torch.randn(xs, device="cuda").bfloat16()
    w = torch.randn(co, xs[-1], 3, 3, 3, device="cuda") * 0.05
    d = ops.conv_desc(xs, co, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    wp = ops.conv3d_pack_weight(d, w, 0)
    gf = 2.0 * xs[0] * xs[1] * xs[2] * xs[3] * co * xs[-1] * 27 / 1e9
    for flags, what in [(0, "normal"), (1, "MMA lane does not wait for operands"), (2, "epilogue does not store"),
                        (4, "producer loads nothing (barriers only)"), (7, "all three")]:
        assert lib.rsp_debug_direct(flags) == 0
        us = t(lambda: ops.conv3d_fprop(d, x, wp))
        print(f"{name:34s} {what:42s} {us:8.1f} us  {gf / us * 1e-3:7.1f} TF/s", flush=True)
    lib.rsp_debug_direct(0)
