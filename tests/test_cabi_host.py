"""CPU tests of the C ABI: the library loads, exports every declared symbol, and its host-side
numbering / index tables are bit-exact against the oracle (no compute calls: there is no GPU here)."""
import os
import re

import numpy as np
import pytest

from oracle.mesh import CartesianMesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dgx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dgx_[a-zA-Z0-9_]+)\s*\(", text)))


def test_all_declared_symbols_exported(dgx):
    syms = declared_symbols()
    assert len(syms) > 50
    missing = [s for s in syms if not hasattr(dgx.lib, s)]
    assert not missing, missing


def test_version_and_error_string(dgx):
    assert b"sm_100a" in dgx.lib.dgx_version()
    with pytest.raises(dgx.DgxError):
        dgx.Mesh(4, [1, 1, 1], 1, [0, 0, 0], [1, 1, 1], [0, 0, 0])


CASES = [
    (2, [1, 1], 3, [True, True]),
    (2, [4, 2], 2, [False, False]),
    (3, [1, 1, 1], 3, [True, True, True]),
    (3, [2, 1, 3], 2, [False, True, False]),
    (3, [1, 1, 1], 0, [True, True, True]),
]


@pytest.mark.parametrize("dim,cells0,nref,periodic", CASES)
def test_cell_order_and_neighbors_bit_exact(dgx, dim, cells0, nref, periodic):
    mesh = dgx.Mesh(dim, cells0, nref, [0.0] * dim, [1.0, 2.0, 3.0][:dim], periodic)
    assert mesh.n_levels() == nref + 1
    for level in range(nref + 1):
        om = CartesianMesh(dim, cells0, level, [0.0] * dim, [1.0, 2.0, 3.0][:dim], periodic)
        assert mesh.n_cells(level) == om.n_cells
        assert np.array_equal(mesh.cell_coords(level), om.coords)
        assert np.array_equal(mesh.neighbors(level), om.nbr)


def test_morton_child_order(dgx):
    """children of cell j on level l are cells 2^dim*j + c, c = i + 2j + 4k (refine_global order)."""
    mesh = dgx.Mesh(3, [2, 1, 1], 2, [0, 0, 0], [2, 1, 1], [False] * 3)
    coarse, fine = mesh.cell_coords(1), mesh.cell_coords(2)
    for j in range(len(coarse)):
        for c in range(8):
            off = np.array([c & 1, (c >> 1) & 1, (c >> 2) & 1])
            assert np.array_equal(fine[8 * j + c], 2 * coarse[j] + off)


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_partition_and_ghost_plan(dgx, nranks):
    """host-only contexts (no device here): owned ranges tile the cell order, ghost lists are the
    off-rank face neighbours, and send/recv plans of different ranks match pairwise."""
    dim, cells0, nref, periodic = 3, [1, 1, 1], 2, [True, True, False]
    mesh = dgx.Mesh(dim, cells0, nref, [0.0] * 3, [1.0] * 3, periodic)
    om = CartesianMesh(dim, cells0, nref, [0.0] * 3, [1.0] * 3, periodic)
    bounds = om.partition(nranks)
    covered = 0
    for r in range(nranks):
        ctx = dgx.Context(0, r, nranks)
        mf = dgx.MatrixFree(ctx, mesh)
        first, no, ng = mf.first_owned_cell(), mf.n_owned_cells(), mf.n_ghost_cells()
        assert (first, first + no) == (bounds[r], bounds[r + 1])
        covered += no
        nb = om.nbr[first:first + no]
        expect_ghost = np.unique(nb[(nb >= 0) & ((nb < first) | (nb >= first + no))])
        ghosts = mf.ghost_cells()
        assert np.array_equal(ghosts, expect_ghost)
        loc = mf.local_neighbors()          # [2*dim][n_owned]
        glob = np.where(loc < 0, loc, np.where(loc < no, loc + first, 0))
        isg = loc >= no
        glob[isg] = ghosts[loc[isg] - no]
        assert np.array_equal(glob.T, nb)
        ctx.close()
    assert covered == om.n_cells


def test_dealii_shim_header_and_driver_build(dgx):
    """include/dgx_dealii_shim.h (the C++ mirror of the deal.II surface) compiles with plain g++, every dgx_* function it
    or the example driver calls is declared in dgx.h, and the driver -- built by build() next to the library -- fails
    loudly (exit code 1 like NS.cc:3055-3077) when there is no device instead of falling back to a CPU path."""
    import subprocess
    import torch
    inc = os.path.join(ROOT, "include")
    src = "#include \"dgx_dealii_shim.h\"\nint main() { dgxshim::SolverControl c(10, 1e-8); return (int)c.max_steps() - 10; }\n"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c++", "-"], input=src, text=True,
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    declared = set(declared_symbols())
    for f in (os.path.join(inc, "dgx_dealii_shim.h"), os.path.join(ROOT, "examples", "trbdf2_tgv.cpp")):
        used = set(re.findall(r"\b(dgx_[a-z0-9_]+)\s*\(", open(f).read()))
        assert used and used <= declared, sorted(used - declared)
    exe = os.path.join(os.path.dirname(dgx.LIB_PATH), "trbdf2_tgv")
    assert os.path.exists(exe)
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "2", "1"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm) prints one JSON line with the contract's
    keys; it times the oracle's C port on a bounded sample and needs no GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-refine", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "DoF/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
