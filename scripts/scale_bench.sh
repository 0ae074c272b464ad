#!/bin/bash
# bench.py at N = 1, 2, 4, 8 GPUs of one box (run under `gpurun --gpus 8`)
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
port=29700
for n in 1 2 4 8; do
  port=$((port + 1))
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/scale_${tag}_g$n.json 2> $out/scale_${tag}_g$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/scale_${tag}_g$n.json 2> $out/scale_${tag}_g$n.err
  fi
  echo "n=$n rc=$? $(tail -1 $out/scale_${tag}_g$n.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["clocks"])')"
done
