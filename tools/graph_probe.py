import sys, traceback
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from helpers import build_product_moco
from rspnet_b200.engine import PretrainEngine
from rspnet_b200.moco import Loss
cfg = dict(arch="resnet18", seed=0, K=64); hyper = dict(dim=128, m=0.999, T=0.07, diff_speed=[2])
gen = torch.Generator().manual_seed(21)
ring = [(torch.randn(8, 3, 8, 64, 64, generator=gen).cuda(), torch.randn(8, 3, 8, 64, 64, generator=gen).cuda()) for _ in range(2)]
model = build_product_moco(cfg, hyper, rank=0).cuda()
eng = PretrainEngine(model, Loss(2.0, 1.0, 1.0), lr=0.00625, cuda_graph=True)
import warnings
warnings.simplefilter("always")
from rspnet_b200 import _lib, ops
_raw = _lib.call
def _spy(name, *a):
    if torch.cuda.is_current_stream_capturing() and SPY["on"]:
        SPY["log"].append((name, torch.cuda.current_stream().cuda_stream))
    return _raw(name, *a)
SPY = {"on": False, "log": []}
_lib.call = _spy; ops.call = _spy
import rspnet_b200.moco.exchange as _ex
for i in range(3):
    eng.step(*ring[i % 2])
torch.cuda.synchronize()
for i in range(3, 9):
    out = eng.step(*ring[i % 2])
torch.cuda.synchronize()
print("graph_error:", eng.graph_error, "graphs:", [g is not None for g in eng._graphs], "graph steps", eng._graph_steps,
      "loss", [float(x) for x in out])
