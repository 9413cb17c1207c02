"""Write one of the synthetic BASELINE scenes as a raw float32 file for julia/bench_reference.jl
(the real RANSAC.jl timed on the same points): 3n vertex floats (x0 y0 z0 x1 ...) followed by 3n normal floats.

  python -m tools.dump_scene c1 scene_c1.bin          # c1 | c2 | c4   (c5 is 2.4 GB: pass it explicitly)
  julia --project=<env with RANSAC v0.6.0> julia/bench_reference.jl scene_c1.bin

Needs no GPU and does not load the CUDA library (the scene generators are NumPy only).
"""
import sys

import numpy as np

from ransac_jl_b200 import scenes

MAKERS = {"c1": scenes.scene_c1, "c2": scenes.scene_c2, "c4": scenes.scene_cad, "c5": scenes.scene_lidar}


def main(argv):
    if len(argv) != 3 or argv[1] not in MAKERS:
        sys.exit(__doc__)
    sc = MAKERS[argv[1]]()
    v = np.ascontiguousarray(sc.vertices, dtype=np.float32)
    n = np.ascontiguousarray(sc.normals, dtype=np.float32)
    with open(argv[2], "wb") as f:
        f.write(v.tobytes())
        f.write(n.tobytes())
    print(f"{argv[2]}: {len(v)} points, {2 * v.nbytes} bytes")


if __name__ == "__main__":
    main(sys.argv)
