"""Diagnostic (not a test): run every conv kernel variant against torch fp32 on bf16-rounded data and print
error statistics. Used on the GPU box while bringing the tcgen05 kernels up."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def stats(name, got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    bad = (~torch.isfinite(got)).sum().item()
    print(f"  {name}: max_abs_err={err.max().item():.4e} rel_to_max={err.max().item() / denom:.4e} "
          f"mean_abs_err={err.mean().item():.4e} ref_absmax={denom:.3e} nonfinite={bad}", flush=True)
    return err.max().item() / denom


def run_case(n, ci, co, tin, hin, win, k, s, p, do_dgrad=True):
    print(f"case N={n} Ci={ci} Co={co} in={tin}x{hin}x{win} k={k} s={s} p={p}", flush=True)
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(n, ci, tin, hin, win, generator=g).to(dev)
    w = (torch.randn(co, ci, *k, generator=g) * (1.0 / (ci * k[0] * k[1] * k[2]) ** 0.5)).to(dev)
    xb = x.bfloat16().float()
    wb = w.bfloat16().float()
    xr = xb.clone().requires_grad_(True)
    wr = wb.clone().requires_grad_(True)
    yref = F.conv3d(xr, wr, None, s, p)
    cis, cos = ops.pad_channels(ci), ops.pad_channels(co)
    xn = ops.to_ndhwc_bf16(x, cis)
    desc = ops.conv_desc(xn.shape, cos, k, s, p)
    wp = ops.conv3d_pack_weight(desc, w, 0)
    y = ops.conv3d_fprop(desc, xn, wp, None)
    torch.cuda.synchronize()
    r = [stats("fprop", ops.to_ncdhw_f32(y, co), yref.detach())]
    dy = torch.randn(yref.shape, generator=g).to(dev)
    dyb = dy.bfloat16().float()
    yref.backward(dyb)
    dyn = ops.to_ndhwc_bf16(dy, cos)
    dw = ops.conv3d_wgrad(desc, xn, dyn, w.shape)
    torch.cuda.synchronize()
    r.append(stats("wgrad", dw, wr.grad))
    if do_dgrad and cis % 64 == 0:
        wd = ops.conv3d_pack_weight(desc, w, 1)
        dx = ops.conv3d_dgrad(desc, dyn, wd)
        torch.cuda.synchronize()
        r.append(stats("dgrad", ops.to_ncdhw_f32(dx, ci), xr.grad))
    return max(r)


if __name__ == "__main__":
    cases = [
        (2, 64, 64, 4, 8, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
        (1, 64, 128, 3, 9, 7, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
        (2, 128, 128, 4, 10, 10, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
        (2, 64, 128, 4, 8, 8, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
        (2, 128, 256, 2, 6, 6, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
        (1, 3, 64, 6, 20, 20, (7, 7, 7), (1, 2, 2), (3, 3, 3)),
        (2, 3, 64, 4, 12, 12, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
        (4, 64, 64, 8, 28, 28, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ]
    worst = 0.0
    for c in cases:
        try:
            worst = max(worst, run_case(*c))
        except Exception as e:  # keep going: we want every data point from one GPU call
            print(f"  EXCEPTION: {type(e).__name__}: {e}", flush=True)
            worst = float("inf")
            if "CUDA error" in str(e) or "trap" in str(e) or "launch failure" in str(e):
                print("  (context is dead; stopping)")
                break
    print(f"WORST rel_to_max = {worst:.4e}")
