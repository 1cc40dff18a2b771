"""The generated marching-cubes table (vdn_nerf_b200/mcubes_table.py) yields watertight, consistently oriented meshes:
checked with the CPU restatement (oracle/mc_oracle.py) on a sphere and on random smooth fields that exercise the
ambiguous faces."""
import numpy as np

from oracle import mc_oracle
from vdn_nerf_b200.mcubes_table import TRI_COUNT, TRI_TABLE


def test_table_shape_and_trivial_cases():
    assert TRI_TABLE.shape == (256, 16) and TRI_COUNT[0] == 0 and TRI_COUNT[255] == 0
    assert TRI_COUNT.max() == 5 and (TRI_TABLE[0] == -1).all()
    for m in range(256):
        n = TRI_COUNT[m]
        assert (TRI_TABLE[m, : 3 * n] >= 0).all() and (TRI_TABLE[m, 3 * n:] == -1).all()


def test_sphere_is_closed_oriented_and_accurate():
    n, r0 = 40, 0.62
    ax = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    u = r0 - np.sqrt(x * x + y * y + z * z)          # = -sdf, inside positive (renderer.py:446)
    v, t = mc_oracle.marching_cubes(u, 0.0)
    bad_edges, dup_directed, euler, vol = mc_oracle.mesh_report(v, t)
    assert bad_edges == 0 and dup_directed == 0 and euler == 2
    h = 2.0 / (n - 1)
    vol_world = vol * h ** 3
    assert vol_world > 0                             # outward orientation
    assert abs(vol_world - 4.0 / 3.0 * np.pi * r0 ** 3) < 0.02 * 4.0 / 3.0 * np.pi * r0 ** 3
    radius = np.linalg.norm(v * h - 1.0, axis=1)
    assert np.abs(radius - r0).max() < 0.5 * h * h / r0 + 1e-6      # linear interpolation error of a curved field


def test_random_fields_are_watertight():
    rng = np.random.default_rng(0)
    n = 20
    for trial in range(6):
        f = rng.standard_normal((n, n, n))
        for _ in range(2):                           # smooth a little, keep plenty of ambiguous configurations
            f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, 1, 2)) / 4.0
        f[0, :, :] = f[-1, :, :] = f[:, 0, :] = f[:, -1, :] = f[:, :, 0] = f[:, :, -1] = -10.0     # closed inside the grid
        v, t = mc_oracle.marching_cubes(f, 0.0)
        assert len(t) > 100
        bad_edges, dup_directed, _, vol = mc_oracle.mesh_report(v, t)
        assert bad_edges == 0, (trial, bad_edges)
        assert dup_directed == 0, (trial, dup_directed)
        assert vol > 0
