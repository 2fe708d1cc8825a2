"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root):
    python tests/golden/make_golden.py
Each file holds the converged RGBA8 image, RGBA32F image, page table, visibility counts and the
request lists of one small seeded scene (tests/golden_scenes.py).  The reference renderer itself
cannot run here (no GL), so these vectors pin the ORACLE; the data side of the oracle is pinned
against the compiled reference converter in tests/test_octree_ref.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from golden_scenes import SCENES, make  # noqa: E402


def main():
    for name in SCENES:
        s = make(name)
        r = s.oracle_render()
        reqs = np.concatenate([q.reshape(-1, 4) for q in r["requests"]] + [np.zeros((0, 4), np.uint32)])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rgba8=r["rgba8"], image=r["image"].astype(np.float32),
                            meta=r["meta"], counts=np.array(r["counts"], np.uint32), requests=reqs,
                            samples=np.uint64(r["stats"].samples), subframes=np.uint32(r["subframes"]),
                            minmax=s.octree.minmax)
        print(name, r["rgba8"].shape, "subframes", r["subframes"], "samples", r["stats"].samples)


if __name__ == "__main__":
    main()
