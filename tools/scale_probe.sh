#!/bin/bash
# N-GPU experiments on the data-parallel step (phases + knobs); usage: tools/scale_probe.sh N  -> gpurun_out/r02_scale_probe_nN.txt
N=${1:-2}
out=gpurun_out/r02_scale_probe_n$N.txt
: > $out
run() {  # label, env...
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29560 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-roofline --no-cpu-baseline 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', 'N=%d' % d['n_gpus'], 'clips/s %.0f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'host_ms %.2f' % d['host_enqueue_ms_per_step'], d['phases_ms'])" >> $out
}
python bench.py --steps 20 --warmup 5 --no-e2e --no-roofline --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single', 'clips/s %.0f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'host_ms %.2f' % d['host_enqueue_ms_per_step'], d['phases_ms'])" >> $out
run default X=1
run max_ctas_4 NCCL_MAX_CTAS=4
run max_ctas_8 NCCL_MAX_CTAS=8
run bucket_8MB RSP_DDP_BUCKET_MB=8
run bucket_256MB RSP_DDP_BUCKET_MB=256
run no_wgrad_overlap RSP_WGRAD_OVERLAP=0
run no_key_overlap RSP_KEY_OVERLAP=0
run a2a_exchange RSP_SHUFFLE_EXCHANGE=a2a
cat $out
