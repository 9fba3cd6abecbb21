"""Golden vector for the binvox obstacle reader: a synthetic voxel map is written with
sca_b200.maps.write_binvox, read by the UNMODIFIED reference (mamp/read_map.py:42-85, build container only), and
the synthetic file + the reference's obstacle list are stored in tests/golden/map_synthetic.npz.

    python tests/golden/make_map_golden.py
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness as rh  # noqa: E402
from sca_b200 import maps  # noqa: E402

if __name__ == '__main__':
    rng = np.random.default_rng(3)
    occ = rng.random((48, 30, 40)) < 0.3          # file axis order x, z, y
    occ[:, :3, :] = rng.random((48, 3, 40)) < 0.9  # a dense "floor" (z index < 3 -> below ground_z)
    center = (5.0, 7.0)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'synthetic.binvox')
        maps.write_binvox(path, occ, translate=(-2.0, -3.0, -1.0), scale=0.05)
        rh.import_reference()
        from mamp.read_map import read_obstacle
        with contextlib.redirect_stdout(io.StringIO()):
            obs = read_obstacle(center=center, environ='exp3', obs_path=path)
        raw = np.frombuffer(open(path, 'rb').read(), dtype=np.uint8)
    pos = np.array([o.pos_global_frame for o in obs], 'float64')
    rad = np.array([o.radius for o in obs], 'float64')
    ids = np.array([o.id for o in obs], np.int32)
    np.savez_compressed(os.path.join(HERE, 'map_synthetic.npz'), binvox=raw, center=np.array(center), pos=pos, radius=rad, ids=ids)
    print('obstacles', len(obs), 'below ground', int((pos[:, 2] <= -1).sum()), 'file bytes', raw.size)
