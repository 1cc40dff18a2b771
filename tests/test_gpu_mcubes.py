"""GPU marching cubes (csrc/mcubes.cu) against the CPU restatement over the same generated table
(oracle/mc_oracle.py): identical vertex sets and triangles, watertight, outward oriented; and NeuSRenderer.extract_geometry
end to end on the SDF network at its geometric initialisation (a sphere of radius ~0.5)."""
import numpy as np
import pytest
import torch

from oracle import mc_oracle
from tests import util
from tests.test_gpu_parity import DEV, make_renderer
from vdn_nerf_b200 import ops

pytestmark = pytest.mark.gpu


def _canon(v, t, shape):
    """Order-independent form of an indexed mesh: every vertex lies on one grid edge, identified by the integer key
    3 * (linear index of the edge's lower end point) + axis.  Returns (sorted keys, positions in key order, the set of
    triangles as rotation-normalised key triples)."""
    v = v.astype(np.float64)
    frac = np.abs(v - np.round(v))
    axis = frac.argmax(axis=1)
    lo = np.round(v).astype(np.int64)
    lo[np.arange(len(v)), axis] = np.floor(v[np.arange(len(v)), axis] + 1e-9).astype(np.int64)
    key = ((lo[:, 0] * shape[1] + lo[:, 1]) * shape[2] + lo[:, 2]) * 3 + axis
    order = np.argsort(key)
    tk = key[t]
    first = tk.argmin(axis=1)
    tk = np.take_along_axis(tk, (first[:, None] + np.arange(3)[None, :]) % 3, axis=1)      # keeps the orientation
    return key[order], v[order], set(map(tuple, tk.tolist()))


@pytest.mark.parametrize("seed", [0, 1])
def test_gpu_marching_cubes_equals_cpu_restatement(seed):
    rng = np.random.default_rng(seed)
    shape = (23, 19, 21)
    f = rng.standard_normal(shape).astype(np.float32)
    for _ in range(2):
        f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, 1, 2)) / 4.0
    f[0, :, :] = f[-1, :, :] = f[:, 0, :] = f[:, -1, :] = f[:, :, 0] = f[:, :, -1] = -10.0
    want_v, want_t = mc_oracle.marching_cubes(f.astype(np.float64), 0.0)
    got_v, got_t = ops.marching_cubes(torch.from_numpy(f).to(DEV), 0.0)
    got_v, got_t = got_v.cpu().numpy(), got_t.cpu().numpy()
    assert got_t.shape == want_t.shape and got_v.shape == want_v.shape
    gk, gv, gt = _canon(got_v, got_t, shape)
    wk, wv, wt = _canon(want_v, want_t, shape)
    assert np.array_equal(gk, wk) and np.allclose(gv, wv, atol=1e-4) and gt == wt
    bad_edges, dup_directed, _, vol = mc_oracle.mesh_report(got_v.astype(np.float64), got_t)
    assert bad_edges == 0 and dup_directed == 0 and vol > 0


def test_extract_geometry_on_the_initial_sdf_is_a_sphere():
    mods, conf = util.build("womsk_white", device=DEV)
    rend = make_renderer(mods, conf)
    for mode in ("fp32", "tf32"):
        ops.set_precision(mode)
        try:
            res = 64
            v, t = rend.extract_geometry(torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3), res, threshold=0.0)
        finally:
            ops.set_precision("fp32")
        assert v.shape[1] == 3 and t.shape[1] == 3 and len(t) > 1000
        bad_edges, dup_directed, euler, vol = mc_oracle.mesh_report(v.astype(np.float64), t)
        assert bad_edges == 0 and dup_directed == 0 and euler == 2
        r = np.linalg.norm(v, axis=1)
        print(f"[{mode}] extract_geometry: {len(v)} vertices, {len(t)} triangles, radius {r.min():.3f} .. {r.max():.3f}, volume {vol:.4f}")
        # geometric init: sdf ~ |x| - 0.5 up to a mean deviation of 0.12: a blob around the origin
        assert 0.15 < r.min() and r.max() < 0.8 and 0.02 < vol < 1.0
