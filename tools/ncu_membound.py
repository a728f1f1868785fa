"""One launch of every HBM-bound kernel of the R3D-18 step at the bench shape (batch 64, K=16384), for
`ncu --set full` captures and for CUDA-event GB/s figures (north_star: "memory-bound kernels evidenced by achieved HBM
GB/s against peak").

    # GB/s table (CUDA events, L2 flushed between launches):
    python tools/ncu_membound.py time > gpurun_out/membound_gbs.txt
    # ncu capture (only the region between cudaProfilerStart/Stop):
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/membound \
        python tools/ncu_membound.py

Algorithmic bytes per launch are the SURVEY.md §8(d) figures (read + written bytes that must cross HBM once).
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "ncu"
dev = "cuda"
B, K, D = 64, 16384, 128
P = 33_335_745 // 4 * 4 + 264     # R3D-18 wrapper parameters in the flat buffer (16-byte aligned segments)
torch.manual_seed(0)

jobs = []   # (name, algorithmic bytes, fn)

# --- EMA / SGD over the flat buffers -------------------------------------------------------------------------
fq, fk, fg, fm = (torch.randn(P, device=dev) for _ in range(4))
jobs.append(("ema_kernel", 12 * P, lambda: ops.ema_update_(fk, fq, 0.999)))
jobs.append(("sgd_kernel", 24 * P, lambda: ops.sgd_step_(fq, fg, fm, 0.1, 0.9, 1e-4, 1.0, False)))
if hasattr(ops, "sgd_ema_step_"):
    jobs.append(("sgd_ema_kernel", 32 * P, lambda: ops.sgd_ema_step_(fq, fg, fm, fk, 0.1, 0.9, 1e-4, 0.999, False)))

# --- speed gather (fp32 NCDHW in, 3 x bf16 NDHWC4 out) -------------------------------------------------------
im_q = torch.randn(B, 3, 32, 112, 112, device=dev)
im_k = torch.randn(B, 3, 32, 112, 112, device=dev)
perm = torch.randperm(B, device=dev)
S = B * 3 * 32 * 112 * 112 * 4                 # bytes of one fp32 input
# reads: q needs 16 of 32 frames per sample, k and k_neg together touch 24 of 32 frames on average (speed-1 rows read
# frames 0..15 for k and 0,2,..30 for k_neg) -> (0.5 + 0.75) * S; writes: 3 outputs of B*16*112*112*4 bf16
sg_bytes = int(1.25 * S) + 3 * B * 16 * 112 * 112 * 4 * 2
jobs.append(("speed_gather_kernel", sg_bytes, lambda: ops.speed_gather(im_q, im_k, perm, B // 2, 2, 1)))

# --- logits + enqueue ----------------------------------------------------------------------------------------
qa, qm, ka, km, na, nm = (torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1) for _ in range(6))
queue = torch.nn.functional.normalize(torch.randn(D, K, device=dev), dim=0)
qptr = torch.zeros(1, dtype=torch.long, device=dev)
state = {}


def logits_fwd():
    state["l"], state["rows"], _ = ops.moco_logits_fwd(qa, qm, ka, km, na, nm, queue, 0.07, True)


def logits_bwd():
    rows = state["rows"]
    g = torch.zeros_like(rows)
    g[2:4] = 1.0 / B
    g[4:6] = -1.0 / B
    ops.moco_logits_bwd(qa, qm, ka, km, na, nm, queue, 0.07, rows, g, None, None)


jobs.append(("neg_logits_kernel(+lse_finalize), logits materialised", D * K * 4 + 2 * B * (K + 1) * 4, logits_fwd))
jobs.append(("neg_logits_bwd_kernel", D * K * 4 + 2 * B * D * 4, logits_bwd))
jobs.append(("enqueue_kernel", 2 * B * D * 4, lambda: ops.queue_enqueue_(queue, na, qptr)))

# --- BN / pool passes on the layer1 activation [64,8,28,28,64] and the stem output [64,16,56,56,64] ----------
C = 64
y1 = torch.randn(B, 8, 28, 28, C, device=dev).bfloat16()
n1 = y1.numel()
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
sums = torch.zeros(2, 2, C, device=dev)
res = torch.randn_like(y1)
jobs.append(("channel_reduce_kernel<0> (bn_stats)", 2 * n1, lambda: ops.bn_stats(y1, out=sums[0].zero_())))
bn = {}


def bn_fwd():
    ops.bn_stats(y1, out=sums[0].zero_())
    bn["out"], bn["rows"] = ops.bn_finalize_act_fwd(y1, sums[0], sums[1], n1 // C, gamma, beta, 1e-5, 0.1, rm, rv, res,
                                                    True)


def bn_bwd():
    _, _, mean, invstd = bn["rows"]
    ops.bn_act_bwd(res, bn["out"], y1, mean, invstd, gamma, True, True)


jobs.append(("bn_finalize_act_fwd (x + residual -> out)", 6 * n1, bn_fwd))
jobs.append(("bn_act_bwd_reduce + bn_act_bwd_apply (dout,out,x -> dx,dres)", 6 * n1 + 10 * n1, bn_bwd))
ys = torch.randn(B, 16, 56, 56, C, device=dev).bfloat16()
ns = ys.numel()
pd = ops.pool_desc(ys.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
mean0, inv0 = torch.zeros(C, device=dev), torch.ones(C, device=dev)
pool = {}


def pool_fwd():
    pool["out"], pool["idx"], pool["xmax"] = ops.bn_relu_maxpool_fwd(pd, ys, scale, shift)


def pool_bwd():
    ops.bn_relu_maxpool_bwd(pd, pool["dout"], pool["idx"], pool["xmax"], ys, scale, shift, mean0, inv0, gamma)


pool_fwd()
pool["dout"] = torch.randn_like(pool["out"])
no = pool["out"].numel()
jobs.append(("bn_relu_maxpool_fwd (+ argmax, x_max)", 2 * ns + 5 * no, pool_fwd))
jobs.append(("bn_relu_maxpool_fwd, no-grad variant", 2 * ns + 2 * no, lambda: ops.bn_relu_maxpool_fwd(pd, ys, scale, shift, aux=False)))
jobs.append(("bn_relu_maxpool_bwd_sums + bn_relu_maxpool_bwd_dx", 4 * no + (5 * no + 2 * ns + 2 * ns), pool_bwd))

torch.cuda.synchronize()
if mode == "time":
    peak = 6550.4
    pk = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text())["hbm_gbs"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print(f"{'kernel(s)':<62} {'alg MB':>8} {'us':>8} {'GB/s':>8} {'frac of %.0f' % peak:>12}")
    for name, nbytes, fn in jobs:
        fn()
        best = 1e9
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        gbs = nbytes / (best * 1e-3) / 1e9
        print(f"{name:<62} {nbytes / 1e6:8.1f} {best * 1e3:8.1f} {gbs:8.0f} {gbs / peak:12.3f}")
else:
    for _, _, fn in jobs:
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _, _, fn in jobs:
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
