import sys, torch
sys.path.insert(0, "/root/repo")
from rspnet_b200 import ops
x = torch.randn(2, 3, 4, 12, 12, device="cuda")
w = torch.randn(64, 3, 3, 3, 3, device="cuda") * 0.2
xn = ops.to_ndhwc_bf16(x, 4)
d = ops.conv_desc(xn.shape, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1))
wp = ops.conv3d_pack_weight(d, w, 0)
torch.cuda.synchronize(); print("packed", flush=True)
y = ops.conv3d_fprop(d, xn, wp)
torch.cuda.synchronize(); print("ok", y.float().abs().mean().item(), flush=True)
ref = torch.nn.functional.conv3d(x.bfloat16().float(), w.bfloat16().float(), None, 1, 1)
print("err", (ops.to_ncdhw_f32(y, 64) - ref).abs().max().item(), ref.abs().max().item())
