#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU parity tests, the bench line, the ncu launch
# list of the same bench command and one `ncu --set full` capture per hot kernel.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_evidence.sh r02'
# Two parts (gpurun merges at most 64 MiB back per call; a full capture is ~8 MB):
#   A = tests, bench line, launch list, the three kernels of the C2 step;  B = the other kernels and the IDF kernel
set -u
tag=${1:-r02}
part=${2:-AB}
out=gpurun_out
mkdir -p $out
if [[ $part == *A* ]]; then
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
echo "bench rc=$?"
# the step loop alone (no e2e arm, no per-kernel timing loops), IDF set-up launches filtered out
timeout 300 python bench.py --steps 4 --warmup 3 --step-only > $out/bench_short_$tag.json 2> $out/bench_short_$tag.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!k_idf).*' -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 4 --warmup 3 --step-only > $out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
specs="7:k_vg_partition:part 8:k_obs_fused:obsvor 4:k_obs_fused:obs"
else
specs=""
fi
if [[ $part == *B* ]]; then specs="$specs 2:k_voronoi_window:vwin 3:k_mark_windows:mark"; fi
# one `ncu --set full` capture per hot kernel, 256 environments (PROFILE_ENVS of make_profile_summary.py)
for spec in $specs; do
  which=$(echo $spec | cut -d: -f1); kern=$(echo $spec | cut -d: -f2); name=$(echo $spec | cut -d: -f3)
  timeout 200 python scripts/prof_kernels.py --which $which --envs 256 --iters 3 > $out/prof_plain_${tag}_$name.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$kern" -s 1 -c 1 -f \
      -o $out/prof_${tag}_$name python scripts/prof_kernels.py --which $which --envs 256 --iters 3 \
      > $out/ncu_${tag}_$name.log 2>&1
  echo "ncu $name rc=$?"
done
# the IDF integration kernel at C1 and C4 (FP32 / XU pipe fractions)
if [[ $part == *B* ]]; then
for spec in "c1:--world 1024 --features 32" "c4:--world 8192 --features 1024 --polygons 64"; do
  name=idf_${spec%%:*}; args=${spec#*:}
  timeout 200 python scripts/prof_idf.py $args > $out/prof_plain_${tag}_$name.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_idf_generate -c 1 -f \
      -o $out/prof_${tag}_$name python scripts/prof_idf.py $args > $out/ncu_${tag}_$name.log 2>&1
  echo "ncu $name rc=$?"
done
fi
