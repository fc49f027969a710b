"""Generate tests/golden/tsdf_golden.npz from the VERBATIM reference binary.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_tsdf.py
For every scene in tests/helpers.ALL_SCENES it writes the scene's input files, runs
oracle/_ref/Integrate_ref on them and records, for the world.pcd it wrote:
  <name>_count   number of voxels,
  <name>_digest  sha256 of the records sorted by integer voxel coordinate,
  <name>_sample  every STRIDE-th record of that sorted list (float32 [m,4]).
The inputs are regenerated deterministically by the tests, so only these small outputs are committed.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import helpers  # noqa: E402
import pyoracle  # noqa: E402
from elasticreconstruction_b200 import synth  # noqa: E402

STRIDE = 997


def main():
    pyoracle.build()
    out = {}
    for mk in helpers.ALL_SCENES:
        sc = mk()
        with tempfile.TemporaryDirectory() as d:
            args = helpers.write_scene(sc, d)
            log, _ = pyoracle.run_reference_integrate(args, cwd=d)
            world = pyoracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "world.pcd")))
        out[sc["name"] + "_count"] = np.int64(len(world))
        out[sc["name"] + "_digest"] = np.array(pyoracle.world_digest(world))
        out[sc["name"] + "_sample"] = world[::STRIDE].copy()
        print(sc["name"], len(world), out[sc["name"] + "_digest"], [l for l in log.splitlines() if "Total" in l])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tsdf_golden.npz"), **out)


if __name__ == "__main__":
    main()
