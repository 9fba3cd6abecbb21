"""CPU: the binvox obstacle reader (sca_b200/maps.py) against the reference's own reader
(mamp/read_map.py:42-85) -- on the committed synthetic map (tests/golden/make_map_golden.py) and, where the
reference tree is present (build container), on the map it ships."""
import io
import os

import numpy as np
import pytest

from sca_b200 import maps

HERE = os.path.dirname(os.path.abspath(__file__))


def test_synthetic_map_matches_reference_reader():
    g = np.load(os.path.join(HERE, 'golden', 'map_synthetic.npz'))
    obs = maps.obstacles_from_binvox(io.BytesIO(g['binvox'].tobytes()), center=tuple(g['center']))
    assert obs.shape == (len(g['pos']), 4) and len(obs) > 1000
    assert np.array_equal(obs[:, :3], g['pos'])          # same arithmetic, same order (== obstacle ids)
    assert np.array_equal(obs[:, 3], g['radius'])
    assert np.array_equal(g['ids'], np.arange(len(obs)))
    assert (obs[:, 2] <= -1).sum() == 5                  # the sparse "floor" class is exercised too


def test_every_voxel_mode_and_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    occ = rng.random((20, 12, 16)) < 0.2
    path = str(tmp_path / 'm.binvox')
    maps.write_binvox(path, occ, translate=(1.0, 2.0, 3.0), scale=0.5)
    data, dims, tr, sc = maps.read_binvox(path)
    assert dims == [20, 12, 16] and tr == [1.0, 2.0, 3.0] and sc == 0.5
    assert np.array_equal(data, np.transpose(occ, (0, 2, 1)))
    # NB the reference loops y over dims[2] and z over dims[1] of the TRANSPOSED array (shape dims[0], dims[2], dims[1])
    obs = maps.obstacles_from_binvox(path, center=(0, 0), above_every=1, floor_every=1)
    assert len(obs) == int(occ.sum())
    with pytest.raises(IOError):
        maps.read_binvox(io.BytesIO(b'not a voxel file\n'))


@pytest.mark.skipif(not os.path.exists('/root/reference/visualization/map/map.binvox'), reason='reference tree not present')
def test_shipped_exp3_map_gives_the_reference_obstacle_count():
    obs = maps.obstacles_from_binvox('/root/reference/visualization/map/map.binvox', center=(35, 30))
    assert len(obs) == 1491                               # read_obstacle(center=(35, 30), environ='exp3', ...) in the build container
    assert np.allclose(obs[:, :3].min(0), [20.0, 15.0, -1.4]) and np.allclose(obs[:, :3].max(0), [49.9, 44.9, 4.2])
