"""cProfile of the host side of PretrainEngine.step (GPU box): where the CPU time of one step goes.
usage: python tools/host_profile.py [arch] [batch] [steps]"""
import cProfile
import pstats
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200.engine import PretrainEngine  # noqa: E402
from rspnet_b200.models import get_model_class  # noqa: E402
from rspnet_b200.moco import Loss, MoCoDiffLossTwoFc, MultiTaskWrapper  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
torch.manual_seed(0)
base = get_model_class(arch=arch)
model = MoCoDiffLossTwoFc(lambda num_classes=128: MultiTaskWrapper(base, num_classes=num_classes), dim=128, K=16384,
                          m=0.999, T=0.07, diff_speed=[2]).cuda()
engine = PretrainEngine(model, Loss(2.0, 1.0, 1.0), 0.1)
q = torch.randn(B, 3, 32, 112, 112, device="cuda")
k = torch.randn(B, 3, 32, 112, 112, device="cuda")
for _ in range(3):
    engine.step(q, k)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    engine.step(q, k)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
