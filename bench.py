#!/usr/bin/env python
"""Benchmark of the B200-native CoverageControl grid hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): 1024 x 1024 world, 32 robots, 32 Gaussian features,
params/coverage_control_params.toml values, DecentralizedCVT + the LPAC observation (sensor 64,
local 256, CNN 32 x 32) + the objective, i.e. per environment step
    ComputeActions(); a = GetActions(); StepActions(a); GetObjectiveValue(); get_maps(C=32, comm)
`--envs-per-gpu` independent environments (default 1024, the batch of BASELINE config 3) are stepped
together on every GPU (no data-path collective: environments are independent).  One bench "step" = one such pass over
the whole batch; `value` = environment steps per second over all GPUs.

Prints ONE JSON line (rank 0).  See the task contract for the keys.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WORLD, N_ROBOTS, N_FEATURES = 1024, 32, 32
SENSOR, LOCAL, CNN = 64, 256, 32


def toml_params(cc):
    """params/coverage_control_params.toml of the reference, set explicitly (the file is not on the GPU box)."""
    p = cc.Parameters()
    p.pNumRobots = N_ROBOTS
    p.pNumGaussianFeatures = N_FEATURES
    p.pWorldMapSize = p.pRobotMapSize = N_WORLD
    p.pLocalMapSize = LOCAL
    p.pSensorSize = SENSOR
    p.pUpdateExplorationMap = False
    p.pMinSigma, p.pMaxSigma = 40.0, 60.0
    p.pMinPeak, p.pMaxPeak = 6.0, 10.0
    p.pCommunicationRange = 128.0
    p.pMaxRobotSpeed = 5.0
    p.pRobotInitDist = 1024.0
    p.pRobotPosHistorySize = 40
    p.pTimeStep = 0.2
    p.pEpisodeSteps = 2100
    return p


def synth_inputs(seed: int):
    """SURVEY.md 8(d) synthetic inputs: circular Gaussians + robot start positions."""
    rng = np.random.default_rng(seed)
    bnd = np.zeros((N_FEATURES, 6))
    bnd[:, 0:2] = rng.uniform(0, N_WORLD, (N_FEATURES, 2))
    bnd[:, 2] = bnd[:, 3] = rng.uniform(40.0, 60.0, N_FEATURES)
    bnd[:, 4] = rng.uniform(6.0, 10.0, N_FEATURES)
    pos = rng.uniform(1e-4, 1024.0 - 1e-4, (N_ROBOTS, 2))
    return bnd, pos


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path (the reference itself cannot be built here)
# ------------------------------------------------------------------------------------------------
def cpu_c2_step(orc, oenv, torchvision_resize):
    """One C2 environment step as the reference's Python does it (controllers.py:67-88 +
    coverage_env_utils.py:227-277), on the structure-faithful CPU environment."""
    import torch

    _, a, _ = oenv.compute_actions(2)
    oenv.step_actions(a)
    obj = oenv.objective()
    R = oenv.R
    local = torch.from_numpy(np.stack([oenv.local_map(r) for r in range(R)]))
    obst = torch.from_numpy(np.stack([oenv.obstacle_map(r) for r in range(R)]))
    rl, ro = torchvision_resize(local), torchvision_resize(obst)
    comm = np.stack([np.stack(oenv.comm_maps(r, CNN, "py")) for r in range(R)])
    maps = torch.cat([rl.unsqueeze(1), torch.from_numpy(comm), ro.unsqueeze(1)], 1)
    return obj, maps


def use_all_host_threads() -> int:
    """The CPU arm runs with every host core whatever launched it: torch.distributed.run exports OMP_NUM_THREADS=1 to
    its workers, which halved the round-1 reference arm at N >= 2.  Returns the thread count in force."""
    n = int(os.environ.get("CCB_REF_THREADS", os.cpu_count() or 1))
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:  # libgomp may already be loaded (and have read the old value): set it through the runtime as well
        C.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def run_cpu(steps: int, warmup: int, budget_s: float):
    threads = use_all_host_threads()
    import torch
    import torchvision

    from oracle import oracle as orc

    torch.set_num_threads(threads)

    class _P:
        pass

    import coveragecontrol_b200 as ccb

    p = toml_params(ccb)
    op = orc.params_from(p)
    bnd, pos = synth_inputs(0)
    wm, _ = orc.idf_generate(op, bnd, mode=1)
    oenv = orc.Env(op, wm, pos)
    rs = lambda t: torchvision.transforms.functional.resize(  # noqa: E731
        t, (CNN, CNN), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)
    for _ in range(warmup):
        cpu_c2_step(orc, oenv, rs)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        cpu_c2_step(orc, oenv, rs)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    orc.idf_generate(op, bnd, mode=0)
    idf_ms = (time.perf_counter() - t1) * 1e3
    cores = threads
    return {"value": done / dt, "steps_done": done, "seconds": dt, "cores": cores, "idf_cpu_ms": idf_ms,
            "torch_threads": torch.get_num_threads()}



# ------------------------------------------------------------------------------------------------
# extras
# ------------------------------------------------------------------------------------------------
def c4_features(n_world: int, n_gauss: int, n_poly: int, seed: int = 0):
    """BASELINE config 4 inputs: Gaussians (sigma 40-60) + star-shaped polygons of 4-10 vertices, radius <= 64
    (coverage_system.cpp:83-98)."""
    rng = np.random.default_rng(seed)
    bnd = np.zeros((n_gauss, 6))
    bnd[:, 0:2] = rng.uniform(0, n_world, (n_gauss, 2))
    bnd[:, 2] = bnd[:, 3] = rng.uniform(40, 60, n_gauss)
    bnd[:, 4] = rng.uniform(6, 10, n_gauss)
    polys = []
    for _ in range(n_poly):
        c = rng.uniform(64, n_world - 64, 2)
        nv = int(rng.integers(4, 11))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        rad = rng.uniform(16, 64, nv)
        polys.append((np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1), float(5e-5 * rng.uniform(6, 10))))
    return bnd, polys


def make_idf(ccb, p, bnd, polys=(), device=0):
    idf = ccb.WorldIDF(p, device=device)
    for b in bnd:
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    for pts, imp in polys:
        idf.AddUniformDistributionPolygon(ccb.PolygonFeature(ccb.PointVector(pts), imp))
    return idf


def idf_extras(ccb, lib, p, device, clocks):
    """IDF generation at C1 (1024^2, 32 Gaussians) and C4 (8192^2, 1024 Gaussians + 64 polygons): the integration
    kernel alone against a MEASURED FP32 FMA rate, the whole GenerateMap call (upload + kernel + maximum + normalise
    + the map back on the host, which is what the reference's wrapper returns), and the reference's own
    generate_world_map_cuda (oracle/_ref, generate_world_map.cu:175-280: 6 cudaMalloc/Free + thrust max + D2H) on the
    same inputs, same GPU, host clock."""
    import torch

    from coveragecontrol_b200 import _lib

    out = {}
    peak = C.c_double(0.0)
    _lib.check(lib.ccb_bench_fp32_peak(device, C.byref(peak)))
    out["fp32_fma_peak_tflops_measured"] = peak.value
    try:
        from oracle import oracle as orc
        have_ref = orc.ref_idf_available()
    except Exception:
        orc, have_ref = None, False
    for tag, n, bnd, polys in (("c1", N_WORLD, synth_inputs(0)[0], []), ("c4", 8192) + c4_features(8192, 1024, 64)):
        q = p.copy()
        q.pWorldMapSize = q.pRobotMapSize = n
        idf = make_idf(ccb, q, bnd, polys, device)
        idf.GenerateMap()
        ms = C.c_float(0.0)
        _lib.check(lib.ccb_bench_idf_kernel(idf._h, 0, 10, C.byref(ms)))
        # flops = 8 N^2 F + 16 N^2 F p_in + 2 N^2, p_in = share of (cell, feature) pairs inside the
        # truncation radius (pi (T sigma)^2 / N^2 at the mean sigma)
        sig = float(np.mean(bnd[:, 2]))
        p_in = min(1.0, math.pi * (q.pTruncationBND * sig) ** 2 / (n * n))
        flops = n * n * (8.0 * len(bnd) + 16.0 * len(bnd) * p_in + 2.0)
        # what the kernel actually evaluates after the tile cull and the 1-D tabulation: 2 flops (truncation test) per
        # (cell, feature) pair whose feature survives the cull of the cell's 32 x 32 tile, 8 more when inside the
        # truncation ellipse (4 products + 4 FMAs of tabulated factors), 2 x 33 erfcf per (tile, feature)
        tsig = q.pTruncationBND * bnd[:, 2]
        pairs_in = float(np.sum(np.minimum(math.pi * tsig * tsig, n * n)))
        pairs_tested = float(np.sum(np.minimum((2 * tsig + 64.0) ** 2, n * n)))
        useful = 8.0 * pairs_in + 2.0 * pairs_tested + 2.0 * n * n
        hbm, _ = measured_peak_gbs()
        r = {"kernel_ms": ms.value, "map_write_gbs": 4.0 * n * n / (ms.value * 1e-3) / 1e9,
             "frac_of_hbm_write_floor": (4.0 * n * n / (hbm * 1e9)) / (ms.value * 1e-3),
             "useful_flops": useful, "useful_tflops": useful / (ms.value * 1e-3) / 1e12,
             "useful_frac_of_measured_fp32_peak": useful / (ms.value * 1e-3) / 1e12 / peak.value,
             "survey_flops": flops, "survey_formula_tflops": flops / (ms.value * 1e-3) / 1e12,
             "survey_formula_frac_of_fp32_peak": flops / (ms.value * 1e-3) / 1e12 / peak.value,
             "note": ("survey_flops = SURVEY 8(d): 8 N^2 F + ... counts every (cell, feature) pair, which is what the reference's "
                      "kernel evaluates; this kernel culls features per tile, so the formula over-counts its work (a fraction above "
                      "1 measures work avoided); useful_flops / the HBM write floor are the physical bounds")}
        t = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            idf.GenerateMap()
            idf._host = None
            np.asarray(idf.GetWorldMap())
            t.append((time.perf_counter() - t0) * 1e3)
        r["generate_plus_d2h_ms"] = float(np.min(t[1:]))
        if have_ref:
            # the reference kernel wants y-monotone pieces: the triangles this library cuts the polygons into
            pieces = []
            for pts, imp in polys:
                xy = np.ascontiguousarray(pts, np.float64)
                tri = np.zeros((len(xy), 6))
                k = lib.ccb_selftest_triangulate(xy.ctypes.data_as(_lib.c_double_p), len(xy), tri.ctypes.data_as(_lib.c_double_p), len(tri))
                pieces += [(tri[i].reshape(3, 2).copy(), imp) for i in range(max(k, 0))]
            op = orc.params_from(q)
            t = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ref_map, _ = orc.ref_idf_generate(op, bnd, polys=pieces)
                t.append((time.perf_counter() - t0) * 1e3)
            r["reference_generate_world_map_cuda_ms"] = float(np.min(t[1:]))
            r["speedup_vs_reference_call"] = r["reference_generate_world_map_cuda_ms"] / r["generate_plus_d2h_ms"]
            idf.GenerateMap()
            idf._host = None
            r["bit_identical_to_reference"] = bool(np.array_equal(np.asarray(idf.GetWorldMap()), ref_map))
        out[f"idf_{tag}"] = r
    return out


def c2_reference_api(p, device, timed):
    """One environment through the REFERENCE-facing API, call for call what an evaluator does per step
    (python/coverage_control/algorithms/controllers.py:67-88 ControllerCVT.step + coverage_env_utils.py:227-277
    get_maps): ComputeActions, GetActions, StepActions(PointVector), GetObjectiveValue, get_maps(32, comm).  The CPU
    figure beside it is this line's cpu_baseline (the same workload on the oracle port)."""
    import coverage_control as cc
    from coverage_control.algorithms import DecentralizedCVT

    bnd, pos = synth_inputs(0)
    idf = make_idf(cc, p, bnd, device=device)
    idf.GenerateMap()
    env = cc.CoverageSystem(p, idf, cc.PointVector(pos))
    ctrl = DecentralizedCVT(p, env)

    def step():
        ctrl.ComputeActions()
        env.StepActions(ctrl.GetActions())
        env.GetObjectiveValue()
        cc.CoverageEnvUtils.get_maps(env, p, CNN, True)

    ms, _ = timed(step, 200, 20)
    return {"steps_per_s": 200 / (ms * 1e-3), "ms_per_step": ms / 200,
            "calls": "ComputeActions; StepActions(GetActions()); GetObjectiveValue(); CoverageEnvUtils.get_maps(env, params, 32, True)"}


def c4_row_bands(ccb, device, rank, world, dist, args):
    """BASELINE config 4 on the GPUs of this run: one 8192^2 world (1024 Gaussians + 64 polygons), 256 robots, tiled
    in row bands; per step ClairvoyantCVT ComputeActions + StepActions with the per-robot moment partials combined by
    ONE NCCL all-reduce (coveragecontrol_b200/parallel.py).  Device time, max over ranks."""
    import torch

    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT
    from coveragecontrol_b200.parallel import BandedCoverageSystem, BandedWorldIDF

    n, r = 8192, 256
    p = ccb.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = n
    p.pNumRobots = r
    p.pUpdateExplorationMap = False
    p.pTimeStep = 0.2
    bnd, polys = c4_features(n, 1024, 64)
    idf = make_idf(ccb, p, bnd, polys, device)
    pos = np.random.default_rng(1).uniform(1e-4, n - 1e-4, (1, r, 2))
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    BandedWorldIDF(idf)
    t.record()
    torch.cuda.synchronize()
    idf_ms = s.elapsed_time(t)
    stream = torch.cuda.Stream(device=device)
    env = BatchedCoverageSystem(p, 1, r, device=device, stream=stream.cuda_stream)
    env.set_world(0, idf)
    env.reset(pos)
    steps = 50

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    peak, _ = measured_peak_gbs()
    ctrls = None

    def run(banded):
        res = {}
        for kind, ctrl in ctrls.items():
            for _ in range(5):
                banded.step(ctrl)
            barrier()
            with torch.cuda.stream(stream):
                s.record()
                for _ in range(steps):
                    banded.step(ctrl)
                t.record()
            torch.cuda.synchronize()
            ms = torch.tensor([s.elapsed_time(t)], device="cuda")
            if dist is not None:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[kind] = float(ms.item()) / steps
        return res

    # the step's all-reduce: one kernel over peer memory (default where the ranks can map each other's buffers),
    # with NCCL's all_reduce timed beside it on the same state
    banded = BandedCoverageSystem(env, 0, stream=stream, graph=True)
    banded.compute_cells()
    ctrls = {"clairvoyant": BatchedCVT(env, 0), "decentralized": BatchedCVT(env, 2)}
    per_default = run(banded)
    out = {"collective": banded.collective}
    if banded.collective == "peer":
        banded.close()
        banded_nccl = BandedCoverageSystem(env, 0, stream=stream, graph=True, collective="nccl")
        per_nccl = run(banded_nccl)
        banded_nccl.close()
        banded = BandedCoverageSystem(env, 0, stream=stream, graph=True)
        out["ms_per_step_with_nccl_all_reduce"] = per_nccl
    rows = banded.rows[1] - banded.rows[0]
    for kind, per in per_default.items():
        out[kind] = {"ms_per_step": per, "env_steps_per_s": 1e3 / per}
        if kind == "clairvoyant":
            out[kind].update({"rows_per_gpu": rows, "band_map_bytes": 4.0 * n * rows,
                              "band_hbm_floor_ms": 4.0 * n * rows / (peak * 1e9) * 1e3,
                              "frac_of_band_hbm_floor": (4.0 * n * rows / (peak * 1e9) * 1e3) / per})
        else:
            out[kind]["robots_per_gpu"] = banded.robots[1] - banded.robots[0]
    ctrl = ctrls["clairvoyant"]
    obj = banded.step(ctrl, with_objective=True)
    barrier()
    objective = float(obj[0])
    banded.close()  # graphs with a captured all-reduce must go before the process group does
    out.update({"workload": f"C4: {n}x{n} world, {r} robots, 1024 Gaussians + 64 polygons, row bands over {world} GPU(s)",
                "n_gpus": world, "idf_gen_ms_banded": idf_ms, "objective": objective,
                "step": "partials -> all-reduce (peer-memory kernel, or NCCL) -> finalise -> actions -> move + window marks, ONE captured CUDA graph",
                "per_rank_state_bytes": int(banded.per_rank_bytes)})
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1024,
                    help="independent environments stepped together on every GPU (BASELINE config 3 batches 1024)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--episode-steps", type=int, default=600, help="steps of the whole-episode mean (value_episode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-environment / C1 / IDF / C4 extras")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 row-band extra (8192^2 world tiled over the GPUs)")
    ap.add_argument("--step-only", action="store_true",
                    help="only the device-resident step loop (the ncu launch list of the step: no e2e arm, no per-kernel timing)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    E = args.envs_per_gpu
    workload = (f"C2: {N_WORLD}x{N_WORLD} world, {N_ROBOTS} robots, {N_FEATURES} Gaussian features, TOML params, "
                f"DecentralizedCVT step + objective + LPAC observation (S={SENSOR}, L={LOCAL}, C={CNN})")
    config = {"workload": workload, "envs_per_gpu": E, "world": N_WORLD, "robots": N_ROBOTS, "features": N_FEATURES,
              "l2": "state per GPU (world + seen + sys planes = 9 MB, prefix records 17 MB, x envs) is larger than the 126 MB L2",
              "parallelism": f"independent environments sharded over {world} GPU(s), no data-path collective"}
    overrides = {k: v for k, v in sorted(os.environ.items()) if k.startswith("CCB_")}
    if overrides:  # opt-in / diagnostic kernel paths: a line measured with one says so
        config["env_overrides"] = overrides

    if args.impl == "reference":
        # the reference's own CPU path cannot be built here (no Eigen/CGAL/Boost):
        # this arm times the oracle port of it on the host cores; rank 0 only.
        if rank != 0:
            return
        r = run_cpu(10**6, args.warmup, 20.0)
        sample = f"1 environment of the workload, {r['steps_done']} consecutive env steps in {r['seconds']:.1f} s"
        line = {
            "impl": "reference", "metric": "env steps/sec (1024^2 map, 32 robots)", "value": r["value"],
            "unit": "env_steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / r["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 sums over f32 maps", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": "env_steps/s", "cores": r["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": r["value"], "unit": "env_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "idf_gen_ms": r["idf_cpu_ms"],
        }
        print(json.dumps(line))
        return

    import torch

    import coveragecontrol_b200 as ccb
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    p = toml_params(ccb)

    # ---- build the batch: worlds generated on the GPU, one WorldIDF per environment ----------------
    batch = BatchedCoverageSystem(p, E, N_ROBOTS, device=local_rank)
    positions = np.zeros((E, N_ROBOTS, 2))
    idf_evt = []
    for e in range(E):
        bnd, pos = synth_inputs(rank * E + e)
        positions[e] = pos
        idf = ccb.WorldIDF(p, device=local_rank)
        for b in bnd:
            idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
        if e < 8:
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            idf.GenerateMap()
            t.record()
            idf_evt.append((s, t))
        else:
            idf.GenerateMap()
        batch.set_world(e, idf)
    batch.reset(positions)
    ctrl = BatchedCVT(batch, 2)
    obs = torch.empty((E, N_ROBOTS, 4, CNN, CNN), dtype=torch.float32, device=dev)
    obj_dev = torch.empty((1, E), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    idf_ms = float(np.median([s.elapsed_time(t) for s, t in idf_evt[2:]])) if len(idf_evt) > 2 else None

    def device_step():
        # one call: ComputeActions, StepActions, then GetObjectiveValue and get_maps side by side (two streams)
        ctrl.rollout_obs(1, True, obj_dev.data_ptr(), CNN, True, obs.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.ccb_kernel_launch_count()
        s.record()
        for _ in range(steps):
            fn()
        t.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(t)
        launches = lib.ccb_kernel_launch_count() - l0
        if dist is not None:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        barrier()
        return ms, launches

    # ---- device-resident number -------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(device_step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = world * E * args.steps / (ms_total * 1e-3)

    if args.step_only:
        if rank == 0:
            print(json.dumps({"metric": "env steps/sec (1024^2 map, 32 robots)", "value": value, "unit": "env_steps/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                              "config": config, "gpu_launches": int(launches), "note": "--step-only: step loop alone"}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- end to end through the host-buffer API: actions D2H -> host -> H2D, objective D2H ----------
    # page-locked host buffers (the library copies straight from / to a caller's buffer when it is pinned)
    host_actions_t = torch.empty((E, N_ROBOTS, 2), dtype=torch.float64).pin_memory()
    host_actions = host_actions_t.numpy()
    host_obj = np.empty(E)

    def e2e_step():
        ctrl.compute_actions()
        _lib.check(lib.ccb_ctrl_get_actions(ctrl._h, host_actions.ctypes.data_as(C.POINTER(C.c_double))))
        bad = C.c_int(-1)
        _lib.check(lib.ccb_env_step_actions(batch._h, host_actions.ctypes.data_as(C.POINTER(C.c_double)), C.byref(bad)))
        # the observation is launched first: the objective chain runs beside it on the library's second stream
        _lib.check(lib.ccb_env_cnn_input_dev(batch._h, CNN, 1, obs.data_ptr()))
        _lib.check(lib.ccb_env_objective(batch._h, host_obj.ctypes.data_as(C.POINTER(C.c_double))))

    ms_e2e, _ = timed(e2e_step, args.steps, args.warmup)
    e2e_value = world * E * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel durations (same data, CUDA events on the launching stream) ----------------------
    # Three byte counts per kernel:
    #   survey_bytes    SURVEY.md 8(d)'s per-environment formula x E: what the REFERENCE's dense maps would move.
    #   required_bytes  what this design has to move for the same result, counted on the device from the live state
    #                   (ccb_test_count): world values only where a robot-map bit is set, the bit planes themselves,
    #                   the (site, row) look-ups of the block-prefix path.  roofline.achieved / frac use THESE: a
    #                   physical rate that cannot exceed the HBM peak.
    #   traffic         dram__bytes_read + dram__bytes_write of the committed `ncu --set full` capture, scaled to E.
    cells = N_WORLD * N_WORLD
    cnt = C.c_uint64(0)
    _lib.check(lib.ccb_test_count(batch._h, 0, C.byref(cnt)))
    seen_cells = float(cnt.value)        # cells of all robots' local windows their robot maps have seen
    batch.set_option("vg_fused", 0)      # (the fused kernel writes no interval table: count the pairs on the table path)
    _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
    _lib.check(lib.ccb_test_count(batch._h, 1, C.byref(cnt)))
    pairs = float(cnt.value)             # (site, row) pairs with cells in the whole-map partition
    batch.set_option("vg_fused", int(os.environ.get("CCB_VG_FUSED", "1") != "0"))
    ER = float(N_ROBOTS * E)
    win_bits = LOCAL * (LOCAL // 32 + 1) * 4.0          # the mask words of one window
    obs_out = 16.0 * CNN * CNN                          # 4 channels x C^2 floats
    prefix_on = os.environ.get("CCB_VG_PREFIX", "1") != "0"
    fused_on = prefix_on and os.environ.get("CCB_VG_FUSED", "1") != "0"
    rows_all = float(N_WORLD * E)
    # fused partition + moments: one 64-byte prefix entry + one 32-byte map sector per Voronoi boundary of a row, one
    # entry per row for its total, the band moments out; look-up kernel: two entries + two sectors + the table word per pair
    part_bytes = 4.0 * N_ROBOTS * N_WORLD * E
    if fused_on:
        chain_bytes = (pairs - rows_all) * 96.0 + rows_all * 64.0 + 32.0 * N_ROBOTS * (N_WORLD / 128) * E
    elif prefix_on:
        chain_bytes = part_bytes + pairs * (4.0 + 2 * 64.0 + 2 * 32.0)
    else:
        chain_bytes = part_bytes + 4.0 * cells * E
    survey = {
        "voronoi_global(whole chain)": (0, 4.0 * cells * E, chain_bytes),
        ("k_vg_partition<moments>" if fused_on else "k_vg_partition"): (7, 4.0 * cells * E if fused_on else part_bytes,
                                                                         chain_bytes if fused_on else part_bytes),
        ("k_vg_bandsum" if fused_on else "k_vg_lookup" if prefix_on else "k_vg_reduce"):
            (6, 0.0 if fused_on else 4.0 * cells * E, 64.0 * N_ROBOTS * (N_WORLD / 128) * E if fused_on else chain_bytes - part_bytes),
        "k_vg_fixup": (9, 0.0, 0.0),
        "k_voronoi_window": (2, 4.0 * N_ROBOTS * LOCAL * LOCAL * E, 4.0 * seen_cells + win_bits * ER),
        "k_obs_fused(local+obstacle+comm)": (4, (4.0 * N_ROBOTS * LOCAL * LOCAL + 16.0 * N_ROBOTS * CNN * CNN) * E,
                                             4.0 * seen_cells + (win_bits + obs_out) * ER),
        "k_obs_fused<window Voronoi>": (8, (8.0 * N_ROBOTS * LOCAL * LOCAL + 16.0 * N_ROBOTS * CNN * CNN) * E,
                                        4.0 * seen_cells + (win_bits + obs_out) * ER),
        "k_mark_windows": (3, 12.0 * N_ROBOTS * SENSOR * SENSOR * E, 24.0 * SENSOR * (SENSOR // 32 + 2) * ER),
        "k_finalize_cells": (5, 32.0 * N_ROBOTS * E, 32.0 * N_ROBOTS * E),
    }
    kernels = {}

    def time_kernel(which, sbytes, rbytes, iters=10):
        fn = lambda: _lib.check(lib.ccb_bench_kernel(batch._h, which, 1, obs.data_ptr()))  # noqa: E731
        ms, _ = timed(fn, iters, 3)
        t = ms / iters * 1e-3
        return {"ms": ms / iters, "survey_bytes": sbytes, "survey_gbs": sbytes / t / 1e9,
                "required_bytes": rbytes, "required_gbs": rbytes / t / 1e9}

    for name, (which, sbytes, rbytes) in survey.items():
        kernels[name] = time_kernel(which, sbytes, rbytes)
    # the kernels the default path replaces, timed on the same state (A/B): the interval table + site-ordered look-ups,
    # and the streaming reduce (still the CentralizedCVT / labels path)
    if fused_on:
        batch.set_option("vg_fused", 0)
        _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
        kernels["A/B k_vg_partition(table only)"] = time_kernel(7, part_bytes, part_bytes)
        kernels["A/B k_vg_lookup(site-ordered)"] = time_kernel(6, 4.0 * cells * E, pairs * (4.0 + 2 * 64.0 + 2 * 32.0))
        batch.set_option("vg_fused", 1)
    if prefix_on:
        batch.set_option("vg_prefix", 0)
        _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
        kernels["A/B k_vg_reduce(streaming)"] = time_kernel(6, 4.0 * cells * E, 4.0 * cells * E)
        batch.set_option("vg_prefix", 1)
        _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
    batch.set_option("obs_v1", 1)
    kernels["A/B k_cnn_fused8<window Voronoi>(+k_comm_maps), round 1"] = time_kernel(
        8, (8.0 * N_ROBOTS * LOCAL * LOCAL + 16.0 * N_ROBOTS * CNN * CNN) * E, 4.0 * seen_cells + (win_bits + obs_out) * ER)
    batch.set_option("obs_v1", 0)
    peak, peak_src = measured_peak_gbs()
    # the dominant single kernel of the step
    # (the step itself runs the fused observation + window-Voronoi variant, not the two kernels it replaces)
    part = "k_vg_partition<moments>" if fused_on else "k_vg_partition"
    dom = max((part, "k_obs_fused<window Voronoi>"), key=lambda k: kernels[k]["ms"])
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture
        with open(tpath) as f:
            t = json.load(f)
        for k, v in t["kernels"].items():
            if k in kernels:
                kernels[k]["traffic"] = v["dram_bytes_per_env"] * E
                kernels[k]["traffic_gbs"] = v["dram_bytes_per_env"] * E / (kernels[k]["ms"] * 1e-3) / 1e9
        if dom in t["kernels"]:
            traffic = t["kernels"][dom]["dram_bytes_per_env"] * E
            traffic_src = t["source"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["required_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["required_gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "frac_survey_formula": kernels[dom]["survey_gbs"] / peak,
                "note": ("achieved = bytes this design must move (counted on the device: world values under set robot-map bits, "
                         "bit planes, prefix look-ups) / CUDA-event time: a physical rate; frac_survey_formula divides the "
                         "reference's dense-map byte count (SURVEY 8d) by the same time and may exceed 1 -- it measures work "
                         "avoided, not bandwidth; traffic = ncu DRAM bytes of the committed capture"),
                "seen_cells_per_robot": seen_cells / ER, "site_row_pairs_per_env": pairs / E,
                "kernels": kernels}
    if traffic is not None:
        roofline["achieved_dram"] = traffic / (kernels[dom]["ms"] * 1e-3) / 1e9
        roofline["frac_dram"] = roofline["achieved_dram"] / peak

    # ---- the same step over a whole episode: seen regions grow, the window kernels slow down -------------
    value_episode = None
    if not args.no_extras:
        batch.reset(positions)
        ms_ep, _ = timed(device_step, args.episode_steps, 0)
        value_episode = world * E * args.episode_steps / (ms_ep * 1e-3)

    # ---- extras: IDF generation (ours and the reference's own kernel), single-environment API latency, C1, C4 ---
    extras = {"idf_gen_ms_1024_32f_mode0": idf_ms}
    try:
        if args.no_extras:
            raise RuntimeError("skipped (--no-extras)")
        extras.update(idf_extras(ccb, lib, p, local_rank, clocks))
        one = BatchedCoverageSystem(p, 1, N_ROBOTS, device=local_rank)
        bnd, pos = synth_inputs(0)
        idf = ccb.WorldIDF(p, device=local_rank)
        for b in bnd:
            idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
        idf.GenerateMap()
        one.set_world(0, idf)
        one.reset(pos[None])
        c1 = BatchedCVT(one, 0)
        ha = np.empty((1, N_ROBOTS, 2))

        def c1_api_step():
            c1.compute_actions()
            _lib.check(lib.ccb_ctrl_get_actions(c1._h, ha.ctypes.data_as(C.POINTER(C.c_double))))
            bad = C.c_int(-1)
            _lib.check(lib.ccb_env_step_actions(one._h, ha.ctypes.data_as(C.POINTER(C.c_double)), C.byref(bad)))

        ms, _ = timed(c1_api_step, 200, 20)
        extras["c1_single_env_api_steps_per_s"] = 200 / (ms * 1e-3)
        ms, _ = timed(lambda: c1.rollout(1), 200, 20)
        extras["c1_single_env_device_steps_per_s"] = 200 / (ms * 1e-3)
        c1b = BatchedCVT(batch, 0)
        ms, _ = timed(lambda: c1b.rollout(1), max(5, args.steps // 2), 3)
        extras["c1_batched_device_steps_per_s"] = world * E * max(5, args.steps // 2) / (ms * 1e-3)
        extras["c2_single_env_reference_api"] = c2_reference_api(p, local_rank, timed)
    except Exception as exc:  # extras never invalidate the headline line
        extras["extras_error"] = f"{type(exc).__name__}: {exc}"
    try:
        if not args.no_extras and not args.no_c4:
            extras["c4_row_bands"] = c4_row_bands(ccb, local_rank, rank, world, dist, args)
    except Exception as exc:
        extras["c4_error"] = f"{type(exc).__name__}: {exc}"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = run_cpu(10**6, 2, args.cpu_seconds)
        cpu = {"value": r["value"], "unit": "env_steps/s", "cores": r["cores"], "kind": "port",
               "sample": f"1 environment of the workload, {r['steps_done']} consecutive env steps in {r['seconds']:.1f} s",
               "idf_gen_ms": r["idf_cpu_ms"]}

    line = {
        "metric": "env steps/sec (1024^2 map, 32 robots)", "value": value, "unit": "env_steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 sums over f32 maps (u32 seen masks)", "data": "synthetic", "config": config,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "env_steps/s", "ms_per_step": ms_e2e / args.steps,
                "scope": ("the batched C ABI with HOST buffers (page-locked): actions D2H, StepActions H2D, objective D2H per step for all "
                          "environments of the batch; the [E, R, 4, 32, 32] observation stays on the device by design "
                          "(zero-copy hand-off to LPAC); extras.c2_single_env_reference_api is the per-environment API"),
                "h2d_bytes_per_step": int(host_actions.nbytes),
                "d2h_bytes_per_step": int(host_actions.nbytes + host_obj.nbytes)},
        "gpu_launches": int(launches),
        "value_episode": value_episode, "episode_steps": args.episode_steps if value_episode else None,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "extras": extras,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
