"""Obstacle fields from voxel maps: the reference's binvox reader, mamp/read_map.py:15-85,
vectorised.  `obstacles_from_binvox` returns the [m, 4] (x, y, z, r) array `Swarm.set_obstacles` /
`scab_set_obstacles` takes -- the obstacles then live in the library's static uniform grid (built once) instead of
the reference's second kd-tree, so fields of 10^4-10^5 spheres cost the query no more than a few do.

Host-side data-format code (NumPy only); the step itself never runs on the CPU.
"""
import numpy as np


def read_binvox(path_or_file):
    """mamp/read_map.py:15-39: header (dims, translate, scale) + run-length-encoded occupancy, returned with the
    reference's `fix_coords` transpose (x, z, y) -> (x, y, z)."""
    fp = open(path_or_file, 'rb') if isinstance(path_or_file, (str, bytes)) else path_or_file
    try:
        line = fp.readline().strip()
        if not line.startswith(b'#binvox'):
            raise IOError('Not a binvox file')
        dims = list(map(int, fp.readline().strip().split(b' ')[1:]))
        translate = list(map(float, fp.readline().strip().split(b' ')[1:]))
        scale = list(map(float, fp.readline().strip().split(b' ')[1:]))[0]
        fp.readline()                                   # "data"
        raw = np.frombuffer(fp.read(), dtype=np.uint8)
    finally:
        if fp is not path_or_file:
            fp.close()
    values, counts = raw[::2], raw[1::2]
    data = np.repeat(values, counts).astype(bool).reshape(dims)
    return np.transpose(data, (0, 2, 1)), dims, translate, scale


def obstacles_from_binvox(path_or_file, center, resolution=0.1, bias=(-13.5, -13.5, -1.4), radius=0.2,
                          above_every=11, floor_every=1001, ground_z=-1.0):
    """mamp/read_map.py:42-85 (`environ == "exp3"`): every `above_every`-th occupied voxel above `ground_z` and
    every `floor_every`-th below becomes a sphere of `radius`, in the reference's visiting order (x, then the
    transposed array's second axis, then its third) -- which is also the obstacle id order.  The reference's
    counters skip 10 (resp. 1000) voxels between picks: above_every = 11, floor_every = 1001.  Pass 1 / 1 to keep
    every voxel (the 'obstacle fields at scale' case)."""
    data, dims, translate, scale = read_binvox(path_or_file)
    # the reference loops x in range(dims[0]), y in range(dims[2]), z in range(dims[1]) over data[x][y][z]
    sub = data[:dims[0], :dims[2], :dims[1]]
    x, y, z = np.nonzero(sub)                           # row-major == the reference's loop order
    px = (y + translate[1]) * resolution + bias[0] + center[0]
    py = (x + translate[0]) * resolution + bias[1] + center[1]
    pz = z * resolution + bias[2]
    above = pz > ground_z
    rank_above = np.cumsum(above) - 1                   # index of each voxel within its own class
    rank_floor = np.cumsum(~above) - 1
    keep = np.where(above, rank_above % above_every == above_every - 1, rank_floor % floor_every == floor_every - 1)
    out = np.empty((int(keep.sum()), 4), 'float64')
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = px[keep], py[keep], pz[keep], radius
    return out


def write_binvox(path, occupancy_xzy, translate=(0.0, 0.0, 0.0), scale=1.0):
    """Write a (dims[0], dims[1], dims[2]) boolean array (file axis order x, z, y) as binvox: test / tooling helper."""
    occ = np.ascontiguousarray(occupancy_xzy).astype(np.uint8)
    flat = occ.ravel()
    change = np.flatnonzero(np.diff(flat)) + 1
    starts = np.r_[0, change]
    lengths = np.diff(np.r_[starts, flat.size])
    vals = flat[starts]
    pairs = bytearray()
    for v, n in zip(vals.tolist(), lengths.tolist()):
        while n > 0:
            c = min(n, 255)
            pairs += bytes((v, c))
            n -= c
    with open(path, 'wb') as f:
        f.write(b'#binvox 1\n')
        f.write(('dim %d %d %d\n' % occ.shape).encode())
        f.write(('translate %s %s %s\n' % tuple(repr(float(t)) for t in translate)).encode())
        f.write(('scale %s\n' % repr(float(scale))).encode())
        f.write(b'data\n')
        f.write(bytes(pairs))
