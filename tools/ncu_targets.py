"""One launch of each distinct tcgen05 conv kernel on its R3D-18 (batch 64) shape, for `ncu --set full` captures.
usage (GPU box):  ncu --set full --clock-control none --import-source on --profile-from-start off \
                      -k regex:"^conv_|bn_relu_maxpool" -o gpurun_out/conv_full \
                      python tools/ncu_targets.py [arch] [batch] [layer,layer,...]
Only the launches between cudaProfilerStart/Stop are captured.  Order per selected layer: fprop, wgrad, dgrad (dgrad
skipped for layer 0); then the fused BN+ReLU+MaxPool forward / backward-reduce / backward-apply on the stem output."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200 import ops  # noqa: E402
from rspnet_b200.models import get_model_class  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
layers = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 7]
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
torch.cuda.synchronize()
torch.cuda.profiler.start()
for li in layers:
    d, xs = shapes[li]
    x = torch.randn(xs, device="cuda").bfloat16()
    ci_l = 3 if d.Ci == 4 else d.Ci
    w = torch.randn(d.Co, ci_l, d.kt, d.kh, d.kw, device="cuda") * 0.05
    wp = ops.conv3d_pack_weight(d, w, 0)
    stats = torch.zeros(2, d.Co, device="cuda")
    y = ops.conv3d_fprop(d, x, wp, stats=stats)
    dy = torch.randn_like(y)
    ops.conv3d_wgrad(d, x, dy, w.shape)
    if li > 0:
        ops.conv3d_dgrad(d, dy, ops.conv3d_pack_weight(d, w, 1))
    torch.cuda.synchronize()
    print("layer", li, xs, d.Co, flush=True)
if arch == "resnet18":
    c = 64
    y = torch.randn(B, 16, 56, 56, c, device="cuda").bfloat16()
    pd = ops.pool_desc(y.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    scale, shift = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    mean, invstd, gamma = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda"), torch.ones(c, device="cuda")
    out, idx, xmax = ops.bn_relu_maxpool_fwd(pd, y, scale, shift)
    ops.bn_relu_maxpool_bwd(pd, torch.randn_like(out), idx, xmax, y, scale, shift, mean, invstd, gamma)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
