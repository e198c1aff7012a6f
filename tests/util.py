"""helpers shared by the GPU parity tests"""
import numpy as np

from oracle.mesh import CartesianMesh


def make_meshes(dgx, dim, cells0, nref, periodic, lo=None, hi=None, boundary_ids=None):
    lo = [0.0] * dim if lo is None else lo
    hi = [1.0, 1.5, 0.75][:dim] if hi is None else hi
    gm = dgx.Mesh(dim, cells0, nref, lo, hi, periodic, boundary_ids)
    om = CartesianMesh(dim, cells0, nref, lo, hi, periodic, boundary_ids)
    return gm, om


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def smooth_plus_noise(om, nodes, seed=1234, ncomp=1, noise=1e-3):
    """src of SURVEY 8d C2: product of sines interpolated + small uniform noise."""
    pts = om.dof_points(np.asarray(nodes, dtype=np.float64))
    f = np.prod(np.sin(2 * np.pi * pts), axis=-1)
    rng = np.random.default_rng(seed)
    comps = [f * (c + 1) + noise * rng.uniform(-1, 1, size=f.shape) for c in range(ncomp)]
    return np.stack(comps, axis=1).reshape(-1)
