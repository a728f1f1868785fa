#!/bin/bash
# One 8-GPU session: W = 2 / 4 / 8 parity logs, N = 1 / 8 bench lines (R3D-18 config 3, S3D-G config 4, R(2+1)D config 5), timeline.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for W in 2 4 8; do
  $TR --nproc-per-node $W --master-port $((29600 + W)) tools/ddp_parity.py > gpurun_out/r02_ddp_parity_w$W.txt 2>&1
  tail -1 gpurun_out/r02_ddp_parity_w$W.txt
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_r18_n1_box8.json 2> gpurun_out/r02_bench_r18_n1_box8.err
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_r18_n8.json 2> gpurun_out/r02_bench_r18_n8.err
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --arch s3dg --frames 128 --size 224 --batch 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_s3dg_cfg4_n8.json 2> gpurun_out/r02_bench_s3dg_cfg4_n8.err
$TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --arch r2plus1d-vcop --batch 32 --steps 10 --warmup 3 > gpurun_out/r02_bench_r2plus1d_cfg5_n8.json 2> gpurun_out/r02_bench_r2plus1d_cfg5_n8.err
$TR --nproc-per-node 8 --master-port 29614 tools/step_timeline.py > gpurun_out/r02_timeline_n8.txt 2>&1
for f in r18_n1_box8 r18_n8 s3dg_cfg4_n8 r2plus1d_cfg5_n8; do
  python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02_bench_{f}.json").read())
    print(f, "clips/s %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % (d["e2e"]["value"] if d.get("e2e") else -1),
          "host_ms %.2f" % d["host_enqueue_ms_per_step"], d.get("phases_ms"), d.get("multi_gpu_parity"))
except Exception as e:
    print(f, "FAILED", e)
PY
done
tail -30 gpurun_out/r02_timeline_n8.txt
