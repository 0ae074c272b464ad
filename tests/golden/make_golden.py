"""Generate the small golden fixtures under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

The reference's own golden data (python/tests/data/**, a GitHub release tarball, pyproject.toml:79)
is not in the reference tree and the reference cannot be built in this container,
so these vectors come from the oracle, which is itself pinned on the reference's three in-tree
known answers (tests/test_oracle_kat.py).  The torchvision resize fixture comes from the real
torchvision installed in the image.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402

KAT_BND = np.array([[182, 438, 55, 55, 8, 0], [950, 97, 47, 47, 7, 0], [714, 920, 46, 46, 7, 0]], np.float64)
KAT_POS = np.array([[433, 154], [139, 319], [92, 112]], np.float64)


def main():
    p = orc.default_params(pNumRobots=3)
    wm, factor = orc.idf_generate(p, KAT_BND, mode=1)
    env = orc.Env(p, wm, KAT_POS)
    cells = env.voronoi_cells()
    np.savez_compressed(
        os.path.join(HERE, "kat_world.npz"), factor=np.float32(factor), world_16=wm[::16, ::16],
        mass=np.array([c.mass for c in cells]), centroid=np.array([list(c.centroid) for c in cells]),
        objective=env.objective())
    # trajectories of the three controllers (positions every 10 steps)
    for kind, name in [(0, "clairvoyant"), (1, "centralized"), (2, "decentralized")]:
        env = orc.Env(p, wm, KAT_POS)
        traj, objs = [], []
        for i in range(60):
            conv, a, g = env.compute_actions(kind)
            env.step_actions(a)
            if i % 10 == 9:
                traj.append(env.positions())
                objs.append(env.objective())
        np.savez_compressed(os.path.join(HERE, f"kat_traj_{name}.npz"), positions=np.array(traj),
                            objective=np.array(objs))
    # antialias bilinear 256 -> 32 weights from the real torchvision
    try:
        import torch
        import torchvision

        eye = torch.eye(256).unsqueeze(0)
        w = torchvision.transforms.functional.resize(
            eye, (32, 256), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)[0]
        np.save(os.path.join(HERE, "resize_256_32_weights.npy"), w.numpy())
    except ImportError:
        print("torchvision missing: resize fixture not regenerated")


if __name__ == "__main__":
    main()
