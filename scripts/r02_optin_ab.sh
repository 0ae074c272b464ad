#!/bin/bash
# First GPU call of the next round: gate and time the two opt-in paths that were written after round 1's GPU
# minutes were spent: CCB_VG_PREFIX=1 (row prefix sums, k_vg_prefix / k_vg_lookup) and
# CCB_VG_PART_RPT=2|4 (taller CTA bands of k_vg_partition).
#   gpurun --timeout 900 -- 'bash scripts/r02_optin_ab.sh r02a'
set -u
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
# 1. the default suite, then the gated parity test of the prefix path and the whole Voronoi suites under each opt-in
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest default rc=$?"
CCB_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_vglobal.py -m gpu -q -k row_prefix > $out/pytest_prefix_$tag.log 2>&1
echo "pytest prefix rc=$?"; tail -3 $out/pytest_prefix_$tag.log
for rpt in 2 4; do
  CCB_VG_PART_RPT=$rpt python -m pytest tests/test_gpu_vglobal.py tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -q \
      > $out/pytest_rpt${rpt}_$tag.log 2>&1
  echo "pytest rpt=$rpt rc=$?"; tail -1 $out/pytest_rpt${rpt}_$tag.log
done
# the rest of the GPU suite with the prefix path on (1e-9 / 1e-12 comparisons against the oracle and between paths
# show what the prefix differences cost in the last digits)
CCB_VG_PREFIX=1 python -m pytest tests -m gpu -q > $out/pytest_all_prefix_$tag.log 2>&1
echo "pytest all, prefix on rc=$?"; tail -3 $out/pytest_all_prefix_$tag.log
# 2. kernel and chain times on the bench workload, objectives compared with the default
python scripts/ab_vg_reduce.py --envs 1024 --iters 20 \
    --configs "WARPS=1;PREFIX=1;PART_RPT=2;PART_RPT=4;PREFIX=1,PART_RPT=2;PREFIX=1,PART_RPT=4" \
    > $out/ab_optin_$tag.json 2> $out/ab_optin_$tag.err
echo "ab rc=$?"; cat $out/ab_optin_$tag.json
# 3. the bench line with and without the prefix path
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/bench_default_$tag.json 2> $out/bench_default_$tag.err
CCB_VG_PREFIX=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/bench_prefix_$tag.json 2> $out/bench_prefix_$tag.err
echo "bench rc=$?"
python - "$tag" <<'PY'
import json, sys
for name in ("default", "prefix"):
    try:
        line = json.loads(open(f"gpurun_out/bench_{name}_{sys.argv[1]}.json").read().strip().splitlines()[-1])
        print(name, line["value"], line["ms_per_step"])
    except Exception as ex:  # noqa: BLE001
        print(name, "unreadable:", ex)
PY
