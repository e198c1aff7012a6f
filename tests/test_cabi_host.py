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


def test_shim_gridtools_transform_builds_metric_tables(dgx, tmp_path):
    """GridTools::transform of the C++ shim -> dgx_mesh_set_vertices for every level -> MatrixFree carries MappingQ1 metric tables: a
    small C++ program (plain g++ against libdgx.so, host-only context) shears a 2 x 1 mesh and prints two table entries with closed forms"""
    import subprocess
    src = r'''
#include <cstdio>
#include "dgx_dealii_shim.h"
using namespace dgxshim;
int main() {
  Context ctx(0);
  Triangulation<2> tria;
  tria.subdivided_hyper_rectangle({2, 1}, {0.0, 0.0}, {1.0, 1.0 / 3.0}, {false, false}, 1);
  GridTools::transform<2>([](const std::array<double, 2> &X) { return std::array<double, 2>{X[0] + 0.75 * X[1], X[1]}; }, tria);
  if (!dgx_mesh_is_general(tria.h, 0) || !dgx_mesh_is_general(tria.h, 1)) return 2;
  MatrixFree<2> mf(ctx, tria);
  long long nc = 0, nf = 0;
  check(dgx_mf_general_metric_tables(mf.h, 1, nullptr, 0, nullptr, 0, &nc, &nf));
  std::vector<double> cell(nc), face(nf);
  check(dgx_mf_general_metric_tables(mf.h, 1, cell.data(), nc, face.data(), nf, nullptr, nullptr));
  // cell 0, first quadrature point: JxW = det J w = hx hy / 4; G_00 = (1 + s^2) / hx^2 JxW
  std::printf("%lld %lld %.17g %.17g\n", nc, nf, cell[3 * 4], cell[0]);
  return 0;
}
'''
    exe = str(tmp_path / "shim_transform")
    libdir = os.path.dirname(dgx.LIB_PATH)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-", "-L", libdir, "-ldgx",
                        "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined", "-o", exe], input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    nc, nf, jxw, g00 = r.stdout.split()
    hx, hy = 0.25, 1.0 / 6.0                     # 4 x 2 cells on the finest level
    assert int(nc) == 8 * 4 * 4 and int(nf) == 8 * 4 * 5 * 2
    assert abs(float(jxw) - hx * hy / 4) < 1e-15 and abs(float(g00) - (1 + 0.75 ** 2) / hx ** 2 * hx * hy / 4) < 1e-13


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


@pytest.mark.parametrize("dim,cells0,nref,periodic,lg", [
    (3, [1, 1, 1], 4, [True, True, True], 5),      # 4096 cells, 128 bricks of 32 cells, periodic: slabs uploaded from both ends
    (3, [2, 1, 1], 4, [False, True, False], 8),    # boundaries, 2 coarse cells
    (3, [1, 1, 1], 6, [True, True, True], 11),     # the production brick shape (16 x 16 x 8 cells) on 64^3: 8 slabs
    (2, [1, 1], 6, [True, False], 5),              # 2-D: bricks 8 x 4 cells
    (2, [3, 1], 5, [False, False], 3),
])
def test_host_pipeline_schedule_is_valid(dgx, dim, cells0, nref, periodic, lg):
    """The schedule of the pipelined host entry point (pure host logic): uploads and compute groups each tile the owned cells
    exactly once, and no group is computed before the upload segments holding its cells and all their face neighbours are
    in.  With the slab-ordered upload most groups are computable long before the upload ends."""
    mesh = dgx.Mesh(dim, cells0, nref, [0.0] * dim, [1.0] * dim, periodic)
    ctx = dgx.Context(0, 0, 1)                         # host-only context: table queries only
    mf = dgx.MatrixFree(ctx, mesh)
    nc = mf.n_owned_cells()
    ups, grps = mf.host_pipeline_plan(1000, lg)
    assert len(ups) >= 16 and len(grps) >= 8
    for segs in (ups, grps):
        order = np.argsort(segs[:, 0])
        assert segs[order[0], 0] == 0 and segs[order[-1], 1] == nc
        assert np.array_equal(segs[order[1:], 0], segs[order[:-1], 1])       # disjoint and gap-free
    seg_of = np.empty(nc, dtype=np.int64)              # upload position of every cell
    for k, (b, e) in enumerate(ups):
        seg_of[b:e] = k
    nbr = mf.local_neighbors()
    assert np.all(np.diff(grps[:, 2]) >= 0)            # computed in the order they become ready
    for b, e, need in grps:
        assert seg_of[b:e].max() <= need
        nb = nbr[:, b:e]
        nb = nb[(nb >= 0) & (nb < nc)]
        assert seg_of[nb].max() <= need
    # effectiveness: half of the cells are computable when 80 % of the upload is in (unpipelined: none before 100 %)
    ready_at = np.empty(nc, dtype=np.int64)
    for b, e, need in grps:
        ready_at[b:e] = need
    n_slabs = (cells0[dim - 1] << nref) >> ((lg - (dim - 1)) // dim)      # bricks are 2^k cells thick in the last direction
    if n_slabs >= 8:
        assert np.median(ready_at) <= 0.8 * len(ups)
    # automatic brick size: small meshes are not pipelined
    ups0, grps0 = mf.host_pipeline_plan(1000, 0)
    if nc < (1 << 18):
        assert len(ups0) == 0 and len(grps0) == 0


def test_compute_entry_points_fail_loudly_without_a_device(dgx):
    """no CPU fallback anywhere on the product path: on a machine without a GPU operators, multigrid and vectors refuse to be
    created (the oracle is test infrastructure and is never routed to)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    ctx = dgx.Context(0, 0, 1)
    mesh = dgx.Mesh(3, [1, 1, 1], 1, [0, 0, 0], [1, 1, 1], [True] * 3)
    mf = dgx.MatrixFree(ctx, mesh)
    with pytest.raises(dgx.DgxError, match="no CUDA device"):
        dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 2, 100.0, 5e-3)
    with pytest.raises(dgx.DgxError, match="no CUDA device"):
        dgx.Multigrid(ctx, mesh, 2, 100.0, 5e-3)
    with pytest.raises(dgx.DgxError):
        mf.initialize_dof_vector(2)


def test_header_is_plain_c():
    """include/dgx.h is a C header (opaque handles, plain pointers and sizes): it compiles as C11 with -Werror"""
    import subprocess
    src = '#include "dgx.h"\nint main(void) { return dgx_version() == 0; }\n'
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr
