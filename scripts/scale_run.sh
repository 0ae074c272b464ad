#!/bin/bash
# 1 -> 8 GPU evidence (run under `gpurun --gpus 8`): the bench line at N = 1, 2, 4, 8, the row-band
# parity check and the C4 row-band benchmark at 1, 2, 4, 8 GPUs.
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
port=29600
for n in 1 2 4 8; do
  port=$((port + 1))
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/scale_${tag}_g$n.json 2> $out/scale_${tag}_g$n.err
    python scripts/bench_rowband.py > $out/rowband_${tag}_g$n.json 2> $out/rowband_${tag}_g$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/scale_${tag}_g$n.json 2> $out/scale_${tag}_g$n.err
    port=$((port + 1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port scripts/bench_rowband.py > $out/rowband_${tag}_g$n.json 2> $out/rowband_${tag}_g$n.err
  fi
  echo "n=$n bench rc=$? $(cut -c1-120 $out/scale_${tag}_g$n.json | head -1)"
  tail -1 $out/rowband_${tag}_g$n.json | cut -c1-400
done
port=$((port + 1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port scripts/multigpu_check.py > $out/mgpu_check8_${tag}.log 2>&1
echo "mgpu check rc=$?"; tail -1 $out/mgpu_check8_${tag}.log
