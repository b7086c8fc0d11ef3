#!/bin/bash
# scaling evidence on one multi-GPU box: bash tools/gpu_scale.sh "8 4 2"
mkdir -p gpurun_out
for n in ${1:-8}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  echo "N=$n rc=$?"; grep -E '^\{' gpurun_out/bench_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('full_model') or {}).get('frame_pairs_per_s'), (d.get('eval64') or {}).get('pairs_per_s'), d['clocks'])
"
done
