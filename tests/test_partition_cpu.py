"""Multi-rank host logic on CPU (gloo, world_size 2): the contiguous Morton partition, the local neighbour tables and
the ghost-exchange plan (what Utilities::MPI::Partitioner / p4est give the reference) must be
consistent ACROSS ranks: what one rank sends is exactly what its peer expects, in the same order (bit-exact)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

CASES = [
    (3, [2, 1, 1], 2, [True, True, True]),
    (3, [1, 1, 1], 2, [True, False, True]),
    (2, [3, 2], 2, [False, True]),
]


def _worker(rank, world, port, q):
    import dgx
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = dgx.Context(0, rank, world, None)          # host-only context: table queries only
        for dim, cells0, nref, periodic in CASES:
            mesh = dgx.Mesh(dim, cells0, nref, [0.0] * dim, [1.0] * dim, periodic)
            for level in range(nref + 1):
                mf = dgx.MatrixFree(ctx, mesh, level)
                mine = {"first": mf.first_owned_cell(), "n": mf.n_owned_cells(), "ghosts": mf.ghost_cells(),
                        "peers": [(r, (cells.astype(np.int64) + mf.first_owned_cell()), rf, rc) for r, cells, rf, rc in mf.peers()]}
                allv = [None] * world
                dist.all_gather_object(allv, mine)
                N = mesh.n_cells(level)
                # contiguous cover of the cell order
                assert allv[0]["first"] == 0 and sum(v["n"] for v in allv) == N
                for a, b in zip(allv[:-1], allv[1:]):
                    assert a["first"] + a["n"] == b["first"]
                # local neighbour table maps to the global one
                if mine["n"]:
                    glob = mesh.neighbors(level, mine["first"], mine["n"])          # [n][2 dim]
                    loc = mf.local_neighbors()                                      # [2 dim][n]
                    for f in range(2 * dim):
                        l = loc[f].astype(np.int64)
                        owned = (l >= 0) & (l < mine["n"])
                        ghost = l >= mine["n"]
                        back = np.where(owned, l + mine["first"], l)
                        back[ghost] = mine["ghosts"][l[ghost] - mine["n"]]
                        assert np.array_equal(back, glob[:, f])
                # send list of A -> B equals B's ghost range from A, element by element
                for other in range(world):
                    if other == rank:
                        continue
                    theirs = allv[other]
                    sent = [c for r, c, _, _ in theirs["peers"] if r == rank]
                    expect = [(rf, rc) for r, _, rf, rc in mine["peers"] if r == other]
                    if not expect or expect[0][1] == 0:
                        assert not sent or len(sent[0]) == 0
                        continue
                    rf, rc = expect[0]
                    assert len(sent) == 1 and np.array_equal(sent[0], mine["ghosts"][rf:rf + rc])
        # host entry point on several ranks: every rank's upload / compute schedule of its OWNED cells is valid on its own (the cells
        # other ranks need are staged and exchanged before the schedule starts, so only owned neighbours constrain it)
        mesh = dgx.Mesh(3, [1, 1, 1], 5, [0.0] * 3, [1.0] * 3, [True] * 3)
        mf = dgx.MatrixFree(ctx, mesh)
        nc = mf.n_owned_cells()
        ups, grps = mf.host_pipeline_plan(1000, 8)
        assert len(ups) >= 16 and len(grps) >= 8
        for segs in (ups, grps):
            order = np.argsort(segs[:, 0])
            assert segs[order[0], 0] == 0 and segs[order[-1], 1] == nc and np.array_equal(segs[order[1:], 0], segs[order[:-1], 1])
        seg_of = np.empty(nc, dtype=np.int64)
        for k, (b, e) in enumerate(ups):
            seg_of[b:e] = k
        nbr = mf.local_neighbors()
        for b, e, need in grps:
            nb = nbr[:, b:e]
            nb = nb[(nb >= 0) & (nb < nc)]
            assert seg_of[b:e].max() <= need and seg_of[nb].max() <= need
        sent = np.unique(np.concatenate([cells for _r, cells, _rf, _rc in mf.peers()]))
        ghost_nb = np.unique(np.nonzero((nbr >= nc).any(axis=0))[0])
        assert np.array_equal(sent, ghost_nb)     # the staged cells are exactly the owned cells with a neighbour on another rank
        # metric tables of a MappingQ1 level (host logic of MatrixFree::reinit on non-Cartesian cells): a rank's tables are its slice of
        # the single-rank tables, bit for bit -- also for faces whose neighbour cell lives on the other rank (its Jacobian enters A_nbr, sigma)
        def wavy(X):
            x = np.array(X, dtype=np.float64)
            out = x.copy()
            out[..., 0] = x[..., 0] + 0.03 * np.sin(2 * np.pi * x[..., 1])
            out[..., 1] = x[..., 1] + 0.03 * np.sin(2 * np.pi * x[..., 2])
            out[..., 2] = x[..., 2] + 0.03 * np.sin(2 * np.pi * x[..., 0])
            return out
        solo = dgx.Context(0)
        mesh = dgx.Mesh(3, [1, 1, 1], 2, [0.0] * 3, [1.0] * 3, [True, False, True], [0, 1, 1, 1, 4, 5])
        mesh.transform(wavy, [0.0] * 3, [1.0] * 3, [1, 1, 1])
        for level in (1, 2):
            mfm, mfs = dgx.MatrixFree(ctx, mesh, level), dgx.MatrixFree(solo, mesh, level)
            cm, fm = mfm.general_metric_tables(2)
            cs, fs = mfs.general_metric_tables(2)
            lo, n = mfm.first_owned_cell(), mfm.n_owned_cells()
            assert n > 0 and np.array_equal(cm, cs[lo:lo + n]) and np.array_equal(fm, fs[lo:lo + n])
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_and_ghost_plan():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
