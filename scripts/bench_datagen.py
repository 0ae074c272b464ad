"""BASELINE config 3 end to end: batched LPAC dataset generation (coveragecontrol_b200/data_generation.py),
samples per second including the device -> host transfer of every sample.

    python scripts/bench_datagen.py [--envs 1024 --samples 4096 --every 5]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import coveragecontrol_b200 as ccb  # noqa: E402
from coveragecontrol_b200.data_generation import BatchedDataGenerator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1024)
ap.add_argument("--samples", type=int, default=4096)
ap.add_argument("--every", type=int, default=5)
a = ap.parse_args()
p = bench.toml_params(ccb)
ccb.seed(0)
t0 = time.perf_counter()
gen = BatchedDataGenerator(p, a.envs, cnn_map_size=32, every_num_steps=a.every, converged_data_ratio=0.02)
torch.cuda.synchronize()
t1 = time.perf_counter()
data = gen.generate(a.samples)
torch.cuda.synchronize()
t2 = time.perf_counter()
nbytes = sum(t.numel() * t.element_size() for t in data.values())
print(json.dumps({"workload": f"C3 data generation: {a.envs} environments (1024^2, 32 robots, 32 features, ClairvoyantCVT oracle), "
                              f"a sample every {a.every} steps", "samples": a.samples, "setup_s": t1 - t0, "generate_s": t2 - t1,
                  "samples_per_s": a.samples / (t2 - t1), "dataset_bytes": nbytes, "environments_used": gen.env_count}))
