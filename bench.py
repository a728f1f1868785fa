#!/usr/bin/env python
"""Benchmark of the RSPNet pretraining step (BASELINE.json: "pretrain clips/sec (R3D-18, 16x112x112)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arch resnet18|c3d|r2plus1d-vcop|s3dg] [--batch B]
                    [--frames F --size S]    # e.g. BASELINE config 4: --arch s3dg --frames 128 --size 224 --batch 8
    python bench.py --impl reference ...      # the UNMODIFIED reference (baseline/_ref) on the host CPU cores

One "step" = one full training iteration on a batch of B synthetic videos per GPU: EMA of the key encoder,
speed re-sampling, two key-encoder forwards (shuffle-BN), query forward, logits + 3-term loss, backward,
gradient all-reduce, SGD, queue update.  One "clip"  = one video = one (clip_q, clip_k) pair.

`value`  : clips/s with the fp32 input clips already resident in HBM (a ring of batches larger than L2).
`e2e`    : clips/s through the public API from PINNED HOST buffers (double-buffered H2D on a copy stream inside the timed
           region, loss read back device->host every step).  Two feeds are measured and both kept in `e2e_feeds`:
           "uint8_frames" — the reference loader's hand-off: uint8 frames cross PCIe, GPUClipSampler (the
           SequentialGPUCollateFn replacement) builds the clips on the device; "fp32_clips" — ready-made fp32 clips, the
           bare model(clip_q, clip_k) contract.  `e2e` is the uint8 loader hand-off at every N.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HYPER = dict(dim=128, K=16384, m=0.999, T=0.07, diff_speed=[2], margin=2.0, A=1.0, M=1.0, lr=0.1, momentum=0.9,
             weight_decay=1e-4)
# forward conv GFLOP per 16-frame clip, first-conv share and the clip size they were probed at;
# other clip sizes scale with frames x height x width (S3D-G 64 x 224^2 = 4.0 x the 16-frame figure)
CONV_GF = {"resnet18": (16.62, 6.61, 112), "c3d": (76.99, 2.08, 112), "r2plus1d-vcop": (42.72, 1.22, 112),
           "s3dg": (34.07, 1.89, 224)}


def conv_gflop(arch, frames, size):
    """(forward GFLOP per clip, first-conv GFLOP) at `frames` loaded frames (clip = frames / 2) and size x size."""
    if arch not in CONV_GF:
        return None, None
    fwd, first, base = CONV_GF[arch]
    scale = (frames / 2 / 16.0) * (size / float(base)) ** 2
    return fwd * scale, first * scale


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--batch", type=int, default=64, help="videos per GPU")
    ap.add_argument("--frames", type=int, default=32, help="loaded frames per clip (2 x 16)")
    ap.add_argument("--size", type=int, default=112)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step from Python instead of replaying CUDA graphs")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_clips_per_s(arch, steps, warmup, batch=4, frames=32, size=112, K=16384, budget_s=150.0):
    """The reference's own CPU implementation of the step on the host cores: the UNMODIFIED reference modules from the
    git-ignored copy baseline/_ref (oracle/install_ref.py; loaded by oracle/ref_loader.py with the three shims of
    SURVEY.md 8c) — MoCoDiffLossTwoFc + MultiTaskWrapper + backbone under DistributedDataParallel (gloo, world 1),
    Loss(margin 2), torch.optim.SGD — on BASELINE config 1 (batch 4, 2x16x112x112, K=16384), all host threads.
    Falls back to the oracle port (oracle/rspnet_oracle.py) only when baseline/_ref is absent.  Returns
    (clips/s, ms per step, threads, batch, kind, steps actually timed)."""
    import torch
    from oracle import ref_loader
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    im_q = torch.randn(batch, 3, frames, size, size, generator=g)
    im_k = torch.randn(batch, 3, frames, size, size, generator=g)
    lr = HYPER["lr"] * batch / 64
    if ref_loader.available():
        import torch.distributed as dist
        kind = "reference"
        ref_loader.FORCE_CPU = True            # Tensor.cuda() becomes a no-op while the reference runs on the host
        orig_cuda = torch.Tensor.cuda
        own_group = not dist.is_initialized()
        if own_group:
            # an in-process store: no rendezvous port (under torchrun MASTER_PORT belongs to the launcher's store)
            dist.init_process_group("gloo", store=dist.HashStore(), rank=0, world_size=1)
        torch.manual_seed(0)
        model = ref_loader.build_reference_moco(arch, dim=HYPER["dim"], K=K, m=HYPER["m"], T=HYPER["T"],
                                                diff_speed=HYPER["diff_speed"])
        ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
        crit = ref_loader.build_reference_loss(HYPER["margin"], HYPER["A"], HYPER["M"])
        opt = torch.optim.SGD(ddp.parameters(), lr=lr, momentum=HYPER["momentum"], dampening=0,
                              weight_decay=HYPER["weight_decay"], nesterov=False)

        def step():
            output, target, rl, rt = ddp(im_q, im_k)
            loss, _, _ = crit(output, target, rl, rt)
            opt.zero_grad()
            loss.backward()
            opt.step()
    else:
        from oracle import rspnet_oracle as oracle
        from rspnet_b200.models import get_model_class
        from rspnet_b200.moco import MoCoDiffLossTwoFc, MultiTaskWrapper
        kind, own_group = "port", False
        torch.manual_seed(0)
        base = get_model_class(arch=arch)
        init = MoCoDiffLossTwoFc(lambda num_classes=128: MultiTaskWrapper(base, num_classes=num_classes),
                                 dim=HYPER["dim"], K=K, m=HYPER["m"], T=HYPER["T"], diff_speed=HYPER["diff_speed"])
        sd = {k: v.clone() for k, v in init.state_dict().items()}
        del init
        mom = {}

        def step():
            oracle.train_step(arch, [sd], [im_q], [im_k], [torch.randperm(batch)],
                              (torch.randperm(batch), torch.randperm(batch)), d=2, m=HYPER["m"], T=HYPER["T"],
                              margin=HYPER["margin"], lr=lr, momentum=HYPER["momentum"],
                              weight_decay=HYPER["weight_decay"], mom_bufs=mom)
    times, t_start = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break    # bounded sample: the whole arm must end within a few minutes whatever --steps asks for
    if kind == "reference":
        torch.Tensor.cuda = orig_cuda
    if own_group:
        import torch.distributed as dist
        dist.destroy_process_group()
    ms = statistics.median(times) * 1e3
    return batch / (ms / 1e3), ms, cores, batch, kind, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arch = args.arch if args.arch in ("resnet18", "c3d", "r2plus1d-vcop", "s3dg") else "resnet18"
    v, ms, cores, batch, kind, timed = cpu_reference_clips_per_s(arch, max(1, args.steps), max(1, args.warmup),
                                                                 frames=args.frames, size=args.size)
    src = ("the unmodified reference modules (baseline/_ref) under stock torch CPU kernels" if kind == "reference"
           else "oracle/rspnet_oracle.py (torch CPU restatement; baseline/_ref absent)")
    sample = (f"{timed} steps of {batch} videos ({arch}, 2x{args.frames // 2}x{args.size}x{args.size} frames, "
              f"K=16384) on {cores} host threads, {src}")
    print(json.dumps({
        "impl": "reference", "metric": "pretrain clips/sec", "value": v, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": timed, "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"RSPNet {arch} pretraining step (MoCoDiffLossTwoFc), per-GPU batch {args.batch} videos, "
                               f"2x{args.frames // 2}x{args.size}x{args.size} clips, K={HYPER['K']}, dim 128",
                   "sample": f"each CPU step is a batch of {batch} videos of that workload (same clip size, K and loss: "
                             "BASELINE config 1)"},
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ B200 arm
KERNELS_PER_CALL = {"rsp_conv3d_wgrad": 2, "rsp_queue_enqueue": 2, "rsp_moco_logits_fwd": 3, "rsp_moco_logits_fwd_ranked": 3,
                    "rsp_moco_logits_bwd": 2}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from rspnet_b200 import _lib
    from rspnet_b200.engine import PretrainEngine, scale_learning_rate
    from rspnet_b200.models import get_model_class
    from rspnet_b200.moco import Loss, MoCoDiffLossTwoFc, MultiTaskWrapper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the ONLY stdout line of this script is the JSON: libraries that print banners on stdout (NCCL's version line) go to
    # stderr for the duration of the run
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT; the only stdout line of this script is the JSON
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # launch accounting: every C-ABI call is counted (most launch exactly one kernel)
    counter = {"n": 0, "on": False}
    raw_call = _lib.call

    def counting_call(name, *a):
        if counter["on"]:
            counter["n"] += KERNELS_PER_CALL.get(name, 1)
        return raw_call(name, *a)

    _lib.call = counting_call
    from rspnet_b200 import ops
    ops.call = counting_call

    torch.manual_seed(0 + rank)
    base = get_model_class(arch=args.arch)
    model = MoCoDiffLossTwoFc(
        lambda num_classes=128: MultiTaskWrapper(base, num_classes=num_classes, fc_type="linear"),
        dim=HYPER["dim"], K=HYPER["K"], m=HYPER["m"], T=HYPER["T"], diff_speed=HYPER["diff_speed"]).to(dev)
    lr = scale_learning_rate(HYPER["lr"], world, args.batch)
    engine = PretrainEngine(model, Loss(HYPER["margin"], HYPER["A"], HYPER["M"]), lr, HYPER["momentum"],
                            HYPER["weight_decay"], cuda_graph=not args.no_graph)

    B = args.batch
    shape = (B, 3, args.frames, args.size, args.size)
    in_bytes = 2 * B * 3 * args.frames * args.size * args.size * 4
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring = [(torch.randn(shape, device=dev, generator=gen), torch.randn(shape, device=dev, generator=gen))
            for _ in range(2)]  # 2 x 617 MB >> 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_t = {}

    def timed(fn, steps, detail=None):
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
            marks[i + 1].record()
        host_t["ms"] = (time.perf_counter() - t0) * 1e3 / steps   # CPU time to enqueue one step (no device sync inside)
        barrier()
        ms = marks[0].elapsed_time(marks[-1])
        per = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        stall = 1.0 if max(per) > 2.5 * statistics.median(per) else 0.0
        if world > 1:
            t = torch.tensor([ms, stall], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, stall = float(t[0]), float(t[1])
        if detail is not None:
            detail["stall"] = stall > 0
            detail["max_step_ms"] = max(per)
            detail["median_step_ms"] = statistics.median(per)
        return ms

    last = {}

    def resident_step(i):
        q, k = ring[i % len(ring)]
        last["loss"] = engine.step(q, k)

    # W warm-up steps as asked, and never fewer than 10: the caching allocator's per-stream pools (main, key-encoder and
    # filter-gradient streams) and NCCL's channels settle during the first steps
    # kernels one step launches: counted on the first (eager) warm-up steps through the C-ABI call counter; CUDA-graph
    # replays launch exactly the recorded kernels without passing through Python
    counter["on"], counter["n"] = True, 0
    resident_step(0)
    resident_step(1)
    launches_per_step = counter["n"] // 2
    counter["on"] = False
    for i in range(2, max(args.warmup, 10)):
        resident_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    counter["on"], counter["n"] = True, 0
    detail = {}
    ms_total = timed(resident_step, args.steps, detail)
    remeasured = None
    if detail["stall"]:
        # one step took > 2.5x the median (a driver / allocator hiccup, seen once in ~10 runs as a 270 ms gap): the
        # region is measured once more and the first figure is kept in the line, as the clocks rule does for throttling
        remeasured = {"first_ms_per_step": ms_total / args.steps, "first_max_step_ms": detail["max_step_ms"],
                      "first_median_step_ms": detail["median_step_ms"]}
        counter["n"] = 0
        ms_total = timed(resident_step, args.steps, detail)
    counter["on"] = False
    graphs_on = engine.cuda_graph and any(g is not None for g in engine._graphs)
    # kernels of the repo's library launched inside the timed region (K steps): counted call by call when the steps
    # are enqueued from Python, K x the kernels recorded per step when they are replayed from CUDA graphs
    launches = launches_per_step * args.steps if graphs_on else counter["n"]
    # host cost of enqueueing one step, measured on an empty launch queue (in the timed loop the CPU runs ahead until the
    # driver's launch queue fills and then advances at the GPU's pace)
    host_ms_step = float("inf")
    for i in range(3):
        barrier()
        t0 = time.perf_counter()
        resident_step(i)
        host_ms_step = min(host_ms_step, (time.perf_counter() - t0) * 1e3)
    barrier()
    graphs_state = graphs_on
    # device time of the phases of a step (CUDA events on the main stream; median of 10 steps, max over ranks); these
    # steps are enqueued eagerly (events cannot be recorded inside a replayed graph)
    engine.phase_log = []
    for i in range(10):
        resident_step(i)
    barrier()
    ph = torch.tensor([[m[j].elapsed_time(m[j + 1]) for j in range(3)] for m in engine.phase_log], device=dev)
    engine.phase_log = None
    ph = ph.median(dim=0).values
    if world > 1:
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    phases = dict(zip(("forward_3_encoder_passes_and_logits", "loss_backward_gradient_allreduce_and_per_bucket_sgd",
                       "after_backward"), [round(v, 3) for v in ph.tolist()]))
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(last["loss"][0])
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)

    # ---- dominant kernel roofline: conv kernels (fprop+dgrad+wgrad) timed over a pass of the same shapes ------
    roofline = None if args.no_roofline else conv_roofline(args, B, ms_step)

    # ---- end to end from pinned host memory --------------------------------------------------------------------
    # Two feeds, both double-buffered on a copy stream inside the timed region, both reading every step's loss back:
    #  e2e            : the reference's loader hand-off (datasets/classification/__init__.py:22-50 +
    #                   SequentialGPUCollateFn, transforms_tensor.py:214-233): uint8 decoded frames cross PCIe, the clip
    #                   sampler kernel (crop / resize / gray / colour jitter / flip / normalise, random decisions drawn on
    #                   the host in the reference's order) builds (clip_q, clip_k) on the device, then the step runs;
    #  e2e_fp32_clips : the model-API contract alone — ready-made fp32 clips [B,3,32,H,W] x 2 cross PCIe (617 MB/step).
    e2e = e2e_feeds = None
    if not args.no_e2e:
        import random as pyrandom
        from rspnet_b200.sampler import GPUClipSampler
        copy_stream = torch.cuda.Stream()
        loss_host = [torch.empty(3, pin_memory=True) for _ in range(2)]
        loss_done = [torch.cuda.Event() for _ in range(2)]
        seen = {"loss": None}

        def run_feed(host, stage, make_inputs, sampled):
            """`sampled`: the inputs of a step are BUILT on the device from the staged bytes (clip sampler).  Then three
            stages are in flight, each on its own stream: H2D of step i+2 | sampler of step i+1, writing straight into
            the input pair the captured step reads (PretrainEngine.next_input_pair) | step i.  Otherwise the staged
            tensors are the inputs themselves and only the copy of step i+1 overlaps step i."""
            ready = [torch.cuda.Event() for _ in range(2)]      # H2D of the slot has landed
            consumed = [torch.cuda.Event() for _ in range(2)]   # the staged bytes of the slot have been read
            built = [torch.cuda.Event() for _ in range(2)]      # the sampler has written the slot's clips
            stepped = [torch.cuda.Event() for _ in range(2)]    # the step that read the slot's clips has finished
            clips = [None, None]
            build_stream = torch.cuda.Stream() if sampled else None

            def prefetch(i):
                s = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[s])
                    for dst, src in zip(stage[s], host[s]):
                        dst.copy_(src, non_blocking=True)
                    ready[s].record(copy_stream)

            def build(i, ahead):
                s = i % 2
                with torch.cuda.stream(build_stream):
                    build_stream.wait_event(ready[s])
                    build_stream.wait_event(stepped[s])      # step i-2 read the buffers this fills
                    clips[s] = make_inputs(stage[s], engine.next_input_pair(ahead))
                    consumed[s].record(build_stream)
                    built[s].record(build_stream)

            tick = {"i": 0}

            def e2e_step(_):
                i = tick["i"]
                tick["i"] += 1
                s = i % 2
                main = torch.cuda.current_stream()
                if sampled:
                    if i == 0:
                        prefetch(0)
                        prefetch(1)
                        build(0, 0)
                    build(i + 1, 1)      # next step's clips are sampled while this step computes
                    prefetch(i + 2)      # and the frames of the step after that cross PCIe
                    main.wait_event(built[s])
                    clip_q, clip_k = clips[s]
                    loss = engine.step(clip_q, clip_k)
                    stepped[s].record(main)
                else:
                    if i == 0:
                        prefetch(0)
                    prefetch(i + 1)      # next step's inputs move while this step computes
                    main.wait_event(ready[s])
                    clip_q, clip_k = make_inputs(stage[s], None)
                    loss = engine.step(clip_q, clip_k)
                    consumed[s].record(main)
                loss_host[s].copy_(torch.stack(loss), non_blocking=True)
                loss_done[s].record()
                # the caller reads every step's loss on the host, one step behind the device (as a logging loop does):
                # wait for the PREVIOUS step's copy, so that the CPU can enqueue step i+1 while the GPU still runs step i
                if i > 0:
                    loss_done[1 - s].synchronize()
                    seen["loss"] = float(loss_host[1 - s][0])

            for s in range(2):
                consumed[s].record()
            for i in range(max(2, args.warmup)):
                e2e_step(i)
            ms = timed(e2e_step, args.steps) / args.steps
            torch.cuda.synchronize()
            return ms

        # (1) uint8 frames -> sampler -> step.  Per video the 64 frames its two 32-frame clips are cut from, 128x171
        # (SURVEY.md 8d "pipeline benchmark" frame size); frame indices / boxes / gray / jitter / flip drawn per step.
        FV = 2 * args.frames
        HS, WS = (128, 171) if args.size <= 112 else (256, 342)
        pyrandom.seed(1234 + rank)
        sampler_k = GPUClipSampler(size=args.size, temporal_size=args.frames, strides=[{"stride": 1, "weight": 1}],
                                   crop_scale=(0.4, 1.0),
                                   color_jitter=dict(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.4))
        host_u8 = [(torch.randint(0, 256, (B * FV, HS, WS, 3), dtype=torch.uint8).pin_memory(),) for _ in range(2)]
        stage_u8 = [(torch.empty((B * FV, HS, WS, 3), dtype=torch.uint8, device=dev),) for _ in range(2)]
        offsets, lengths = [v * FV for v in range(B)], [FV] * B

        def from_frames(st, out):
            (clip_q, clip_k), _ = sampler_k(st[0], offsets, lengths, out=out)
            return clip_q, clip_k

        ms_u8 = run_feed(host_u8, stage_u8, from_frames, True)
        u8_bytes = B * FV * HS * WS * 3
        e2e_u8 = {"value": B * world / (ms_u8 / 1e3), "unit": "clips/s", "ms_per_step": ms_u8, "feed": "uint8_frames",
               "h2d_bytes_per_step": u8_bytes + 2 * B * (args.frames * 4 + 16 + 1 + 20), "d2h_bytes_per_step": 12,
               "note": f"loader hand-off as in the reference: pinned host uint8 frames ({FV} frames of {HS}x{WS} per video) "
                       "-> double-buffered H2D on a copy stream -> clip sampler kernels on a second stream (crop, bilinear "
                       "resize, gray, colour jitter, flip, normalise; decisions drawn on the host per step), written "
                       "into the input pair of the captured step -> PretrainEngine.step; three stages in flight (copy "
                       "of step i+2, sampling of step i+1, step i); "
                       "every step's loss is copied to pinned host memory and read there one step later (the timed "
                       "region ends with a full synchronize)"}
        del host_u8, stage_u8
        # (2) ready-made fp32 clips
        host_f = [(torch.randn(shape).pin_memory(), torch.randn(shape).pin_memory()) for _ in range(2)]
        stage_f = [(torch.empty(shape, device=dev), torch.empty(shape, device=dev)) for _ in range(2)]
        ms_f = run_feed(host_f, stage_f, lambda st, out: (st[0], st[1]), False)
        e2e_fp32 = {"value": B * world / (ms_f / 1e3), "unit": "clips/s", "ms_per_step": ms_f, "feed": "fp32_clips",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 12,
                    "note": "pinned host fp32 clips [B,3,T,H,W] x 2 straight into PretrainEngine.step (PCIe-bound: "
                            f"{in_bytes / 1e6:.0f} MB per step)"}
        del host_f, stage_f
        # ONE declared e2e experiment at every N: the reference loader's hand-off (uint8 frames -> device clip sampler ->
        # step), i.e. the contract of datasets/classification/__init__.py:22-50.  The fp32-clip feed (bare
        # model(clip_q, clip_k) contract) stays in `e2e_feeds` for comparison.
        e2e = dict(e2e_u8)
        e2e_feeds = {"uint8_frames": e2e_u8, "fp32_clips": e2e_fp32}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores, cb, kind, timed = cpu_reference_clips_per_s(
            "resnet18" if args.arch not in ("resnet18", "c3d") else args.arch, steps=8, warmup=2, budget_s=40.0)
        cpu = {"value": v, "unit": "clips/s", "cores": cores, "kind": kind, "ms_per_step": ms,
               "sample": f"{timed} steps of {cb} videos (BASELINE config 1: batch 4, 2x16x112x112, K=16384) on {cores} "
                         "host threads, " + ("the unmodified reference modules (baseline/_ref) under stock torch CPU "
                                             "kernels" if kind == "reference" else "oracle/rspnet_oracle.py")}
    parity = multi_gpu_parity(model, engine, last["loss"], dev) if world > 1 else None
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps({
            "metric": "pretrain clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"RSPNet {args.arch} pretraining step (MoCoDiffLossTwoFc), per-GPU batch {B} videos, "
                                   f"2x{args.frames // 2}x{args.size}x{args.size} clips, K={HYPER['K']}, dim 128",
                       "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": f"inputs alternate between two {in_bytes / 1e6:.0f} MB device batches (> 126 MB L2)"},
            "encoder_clip_passes_per_s": 3 * value,   # q, k and k_neg forwards of 16-frame clips per video
            "loss": loss_val, "remeasured": remeasured, "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms_step, "clocks": clocks, "e2e": e2e, "e2e_feeds": e2e_feeds, "roofline": roofline,
            "cpu_baseline": cpu, "multi_gpu_parity": parity, "phases_ms": phases,
            "cuda_graph": {"replayed": graphs_state, "error": engine.graph_error,
                           "kernels_recorded_per_step": launches_per_step},
        }) + "\n").encode())
    # Teardown: captured graphs hold NCCL kernels, and destroying the communicator underneath them can block for
    # minutes (seen once: the NCCL watchdog aborted the process 8 minutes after the JSON line).  The line is out; release
    # the graphs, let every rank pass a last barrier, and leave without the communicator teardown — with a hard stop
    # in case anything still hangs.
    sys.stdout.flush()
    sys.stderr.flush()
    hard_stop = threading.Timer(60.0, os._exit, args=(0,))
    hard_stop.daemon = True
    hard_stop.start()
    engine._graphs = [None, None]
    engine._graph_pool = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def multi_gpu_parity(model, engine, loss3, dev):
    """Evidence carried by every N > 1 line (raises when a check fails, so a broken exchange cannot produce a number):
    after the timed steps the queue, its pointer and the query-encoder parameters must be BIT-IDENTICAL on all ranks
    (keys gathered in rank order, gradients averaged, same SGD everywhere), every rank's loss finite, and the shuffle-BN
    transport must return exactly ``concat_all_gather(x)[idx_shuffle.view(W, -1)[rank]]`` (builder:361-387) on a
    bench-sized batch — the reference formula evaluated with a plain NCCL all_gather."""
    import torch
    import torch.distributed as dist
    from rspnet_b200 import ops
    rank, world = dist.get_rank(), dist.get_world_size()
    res = {}

    def same_everywhere(t):
        ref = t.clone()
        dist.broadcast(ref, src=0)
        return bool(torch.equal(ref, t))

    res["queue_bit_identical"] = same_everywhere(model.queue)
    res["queue_ptr_identical"] = same_everywhere(model.queue_ptr)
    res["encoder_q_parameters_bit_identical"] = same_everywhere(engine.flat_q)
    res["loss_finite"] = bool(torch.isfinite(torch.stack(loss3)).all())
    k_buf = next(iter(model._exchanges.values()))._bufs[0][0]
    x = (torch.randn(k_buf.shape, device=dev) + rank).to(k_buf.dtype)
    got, idx_unshuffle = model._batch_shuffle_ddp(x)
    idx_shuffle = ops.invert_permutation(idx_unshuffle)
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    want = torch.cat(parts, 0)[idx_shuffle.view(world, -1)[rank]]
    res["shuffle_equals_reference_formula"] = bool(torch.equal(got, want))
    res["shuffle_transport"] = next(iter(model._exchanges.values())).mode
    res["shuffle_bytes_per_rank"] = got.numel() * got.element_size()
    flags = torch.tensor([1 if v else 0 for k, v in res.items() if isinstance(v, bool)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    names = [k for k, v in res.items() if isinstance(v, bool)]
    res.update({k: bool(f) for k, f in zip(names, flags.tolist())})
    if not all(res[k] for k in names):
        raise RuntimeError(f"multi-GPU parity check failed: {res}")
    res["ranks"] = world
    return res


def conv_roofline(args, B, ms_step):
    """Tensor-pipe roofline of the conv kernels: algorithmic FLOPs of one step / time spent in conv kernels.

    The conv time is measured live with CUDA events around fprop / dgrad / wgrad launches of every conv shape of the
    backbone (same batch, same tensors sizes), weighted as in a step (3 fprop passes + 1 dgrad + 1 wgrad)."""
    import torch
    from rspnet_b200 import ops
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    # every conv launch below is timed alone (a synchronize on both sides): the BURST figure is the denominator
    peak, sustained, src = 1590.0, None, "fallback"
    if peaks_file.exists():
        pk = json.loads(peaks_file.read_text())
        peak, src = float(pk.get("bf16_tflops", 1590.0)), "measured (burst: each launch is timed in isolation)"
        sustained = pk.get("bf16_tflops_sustained")
    fwd, first = conv_gflop(args.arch, args.frames, args.size)
    if fwd is None:
        return None
    from rspnet_b200.models import get_model_class
    from rspnet_b200 import nn as rnn
    shapes = []
    orig = ops.conv3d_fprop

    def spy(desc, x, wp, bias=None, **kw):
        shapes.append((desc, tuple(x.shape), x.requires_grad))
        return orig(desc, x, wp, bias, **kw)

    ops.conv3d_fprop = spy
    net = get_model_class(arch=args.arch)(num_classes=1).cuda()
    with torch.no_grad():
        net.feature_ndhwc(torch.zeros(B, 3, args.frames // 2, args.size, args.size, device="cuda"))
    ops.conv3d_fprop = orig
    del net
    tot_ms = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    per_launch = []   # (ms weighted as in a step, pass, layer, GFLOP, ms per launch)
    weight = {"fprop": 3, "dgrad": 1, "wgrad": 1}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for li, (desc, xshape, _) in enumerate(shapes):
        x = torch.randn(xshape, device="cuda").bfloat16()
        ci_l = 3 if desc.Ci == 4 else desc.Ci
        w = torch.randn(desc.Co, ci_l, desc.kt, desc.kh, desc.kw, device="cuda") * 0.05
        wp = ops.conv3d_pack_weight(desc, w, 0)
        y = ops.conv3d_fprop(desc, x, wp)
        dy = torch.randn_like(y)
        jobs = [("fprop", lambda: ops.conv3d_fprop(desc, x, wp))]
        dw = torch.empty_like(w)
        jobs.append(("wgrad", lambda: ops.conv3d_wgrad(desc, x, dy, w.shape, out=dw)))
        if li > 0:
            wd = ops.conv3d_pack_weight(desc, w, 1)
            jobs.append(("dgrad", lambda: ops.conv3d_dgrad(desc, dy, wd)))
        for name, fn in jobs:
            fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            tot_ms[name] += ms
            to, ho, wo = desc.out_dims()
            gf = 2.0 * desc.N * to * ho * wo * desc.Co * ci_l * desc.kt * desc.kh * desc.kw / 1e9
            per_launch.append((ms * weight[name], name, li, gf, ms))
        del x, y, dy, w, wp
    conv_ms = 3 * tot_ms["fprop"] + tot_ms["dgrad"] + tot_ms["wgrad"]
    flops = (3 * fwd + (2 * fwd - first)) * 1e9 * B
    achieved = flops / (conv_ms / 1e3) / 1e12
    per_launch.sort(reverse=True)
    top = [{"pass": n, "conv_layer": li, "gflop_per_launch": round(gf, 1), "ms_per_launch": round(ms, 4),
            "tflops": round(gf / ms, 1), "frac": round(gf / ms / peak, 3), "launches_per_step": weight[n]}
           for _, n, li, gf, ms in per_launch[:4]]
    # DRAM traffic of the dominant launch: only when a capture of THIS round is committed (profiles/r02_traffic.json,
    # written from an `ncu --set full` report by tools/ncu_traffic.py); never a constant carried over
    traffic, traffic_note = None, None
    tfile = ROOT / "profiles" / "r02_traffic.json"
    if tfile.exists():
        t = json.loads(tfile.read_text()).get(f"{args.arch}_b{B}_{args.size}")
        if t:
            traffic, traffic_note = t["dram_bytes"], t["note"]
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "frac_of_sustained_peak": (achieved / float(sustained)) if sustained else None,
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": src, "kernel": "tcgen05 conv kernels (conv_direct / conv_igemm / conv_stem / conv_stem3 / conv_wgrad / conv_wgrad_direct)",
            "top_launches": top,
            "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / ms_step,
            "per_pass_ms": tot_ms, "algorithmic_gflop_per_step": flops / 1e9}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
