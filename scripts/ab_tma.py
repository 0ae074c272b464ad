"""A/B: the streaming whole-map reduce with per-lane 16-byte loads (k_vg_reduce) against the same reduce fed by
bulk-asynchronous copies into a shared-memory ring (k_vg_reduce_bulk: cp.async.bulk + mbarrier, the TMA engine).
    python scripts/ab_tma.py [--envs 1024]
Same state, same row partition, CUDA events around 10 launches of the reduce alone (ccb_bench_kernel 6), and the
moments of both compared.  Prints one JSON line."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import coveragecontrol_b200 as ccb  # noqa: E402
from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1024)
args = ap.parse_args()
lib = _lib.load()
p = bench.toml_params(ccb)
E = args.envs
batch = BatchedCoverageSystem(p, E, bench.N_ROBOTS)
pos = np.zeros((E, bench.N_ROBOTS, 2))
for e in range(E):
    bnd, pos[e] = bench.synth_inputs(e)
    idf = bench.make_idf(ccb, p, bnd)
    idf.GenerateMap()
    batch.set_world(e, idf)
batch.reset(pos)
ctrl = BatchedCVT(batch, 2)
ctrl.rollout(20)
obs = torch.empty((E, bench.N_ROBOTS, 4, 32, 32), device="cuda")
batch.set_option("vg_prefix", 0)
out = {"envs": E, "bytes_per_launch": 4.0 * 1024 * 1024 * E}
cells = {}
for name, bulk in (("k_vg_reduce (ld.global.cs, 16 B per lane)", 0), ("k_vg_reduce_bulk (cp.async.bulk + mbarrier)", 1)):
    batch.set_option("vg_bulk", bulk)
    _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
    _lib.check(lib.ccb_bench_kernel(batch._h, 6, 3, obs.data_ptr()))
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _lib.check(lib.ccb_bench_kernel(batch._h, 6, 10, obs.data_ptr()))
    t.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(t) / 10
    out[name] = {"ms": ms, "gbs": out["bytes_per_launch"] / (ms * 1e-3) / 1e9}
    batch._invalidate()
    cells[bulk] = [np.array(x) for x in batch.cells(3)]
out["max_rel_diff_of_moments"] = float(max(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300) for a, b in zip(cells[0], cells[1])))
sass = subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True).stdout
out["sass"] = {k: sass.count(k) for k in ("UBLKCP", "SYNCS", "UTMALDG", "LDGSTS")}
print(json.dumps(out))
