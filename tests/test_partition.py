"""Domain decomposition (host side, no GPU): invariants of the partition every rank computes for itself, checked in one
process and -- the way the ranks really run -- across a world_size-2 gloo group."""
import os
import sys

import numpy as np
import pytest

import mgcfd_b200 as M
from conftest import mesh_levels

MESHES = [(0, [[12, 11, 10], [6, 6, 5], [3, 3, 3]]), (1, [[9, 8, 7], [5, 4, 4]]), (2, [[5, 4, 4]])]


@pytest.mark.parametrize("kind,dims", MESHES)
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_partition_invariants(kind, dims, nranks):
    mesh = M.Mesh.generate(kind, dims, mesh_variant=0 if kind == 2 else 2)
    lv = mesh_levels(mesh)
    for l in range(mesh.levels):
        plans = [M.partition_plan(mesh, nranks, r, l) for r in range(nranks)]
        n = lv[l]["nel"]
        owner = np.full(n, -1)
        for r, p in enumerate(plans):
            own = p["gid"][:p["owned"]]
            assert np.all(owner[own] == -1)                       # owned sets are disjoint ...
            owner[own] = r
            assert np.all(np.diff(own) > 0)                       # ... ascending ...
            assert len(np.unique(p["gid"])) == len(p["gid"])      # a node is local at most once
            assert p["global_nodes"] == n
        assert np.all(owner >= 0)                                  # ... and cover the level
        sizes = np.bincount(owner, minlength=nranks)
        assert sizes.max() - sizes.min() <= max(2, nranks)        # balanced bisection
        e = lv[l]["edges"]
        nI = lv[l]["nI"]
        for r, p in enumerate(plans):
            loc = np.zeros(n, bool); loc[p["gid"]] = True
            own = owner == r
            # flux closure: both ends of every edge of an owned node are local; the rank holds exactly those edges
            touch = own[e["a"][:nI]] | own[e["b"][:nI]]
            assert np.all(loc[e["a"][:nI]][touch]) and np.all(loc[e["b"][:nI]][touch])
            assert p["nI"] == int(touch.sum())
            assert p["nB"] + p["nW"] == int(own[e["b"][nI:]].sum())
            # ghosts grouped by owner, ascending id inside a group; receive counts match
            gh = p["gid"][p["owned"]:]
            assert np.all(np.diff(owner[gh]) >= 0)
            assert np.array_equal(np.bincount(owner[gh], minlength=nranks), p["recv_counts"])
            assert p["recv_counts"][r] == 0 and p["send_counts"][r] == 0
            # what r sends to q is exactly what q holds as ghosts of r, in the same order
            off = 0
            for q in range(nranks):
                cnt = p["send_counts"][q]
                sent = p["send_gids"][off:off + cnt]; off += cnt
                ghq = plans[q]["gid"][plans[q]["owned"]:]
                assert np.array_equal(sent, ghq[owner[ghq] == r])
            if l + 1 < mesh.levels:                                # restrict closure: all children of owned coarse nodes are local
                co = M.partition_plan(mesh, nranks, r, l + 1)
                cown = np.zeros(lv[l + 1]["nel"], bool); cown[co["gid"][:co["owned"]]] = True
                cloc = np.zeros(lv[l + 1]["nel"], bool); cloc[co["gid"]] = True
                assert np.all(loc[cown[lv[l]["map"]]])
                # prolong closure: parents of owned fine nodes and of their edge neighbours are local on the coarse level
                nb = np.zeros(n, bool); nb[own] = True
                nb[e["b"][:nI][own[e["a"][:nI]]]] = True; nb[e["a"][:nI][own[e["b"][:nI]]]] = True
                assert np.all(cloc[lv[l]["map"][nb]])


@pytest.mark.parametrize("kind,dims", MESHES + [(1, [[17, 6, 5], [9, 3, 3], [5, 2, 2]])])
@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
@pytest.mark.parametrize("ewt", [False, True])
def test_rank_local_generation_equals_partition_of_the_assembled_mesh(kind, dims, nranks, ewt):
    """mgcfd_generate_partition_plan (no rank assembles the mesh) must hold, bit for bit, what partition_mesh cuts out of the
    assembled mesh: ids, exchange lists and a hash over volumes, coordinates, edges (indices + weights), maps, lists."""
    variant = 0 if kind == 2 else 2
    lengths = (2.0, 1.0, 1.5)
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant, lengths=lengths)
    if ewt:
        mesh.apply_ewt()
    for l in range(mesh.levels):
        for r in range(nranks):
            a = M.partition_plan(mesh, nranks, r, l)
            b = M.generate_partition_plan(kind, dims, nranks, r, l, mesh_variant=variant, lengths=lengths, apply_ewt=ewt)
            for k in ("owned", "ghosts", "sent", "global_nodes", "nI", "nB", "nW", "hash"):
                assert a[k] == b[k], (k, l, r)
            for k in ("gid", "send_counts", "recv_counts", "send_gids"):
                assert np.array_equal(a[k], b[k]), (k, l, r)
    # the hash is sensitive to the edge weights: adjusted vs raw differ on every non-fvcorr mesh
    if variant != 0 and ewt:
        raw = M.generate_partition_plan(kind, dims, nranks, 0, 0, mesh_variant=variant, lengths=lengths, apply_ewt=False)
        assert raw["hash"] != M.partition_plan(mesh, nranks, 0, 0)["hash"]


@pytest.mark.parametrize("copies,nranks", [(4, 2), (4, 4), (6, 3), (3, 2)])
def test_duplicated_meshes_are_dealt_out_copy_by_copy(copies, nranks):
    """-m copies on N ranks with N | m: every rank holds whole copies (no ghosts, nothing to exchange) -- the reference's one mesh
    copy per thread, across GPUs; otherwise the mesh is bisected as usual and the closure rules still hold."""
    mesh = M.Mesh.generate(M.GEN_HEX_BOX, [[7, 6, 5], [4, 3, 3]], mesh_variant=2)
    base = [mesh.dims(l)[0] for l in range(mesh.levels)]
    mesh.duplicate(copies)
    for l in range(mesh.levels):
        plans = [M.partition_plan(mesh, nranks, r, l) for r in range(nranks)]
        assert sum(p["owned"] for p in plans) == copies * base[l]
        if copies % nranks == 0:
            for r, p in enumerate(plans):
                assert p["ghosts"] == 0 and p["sent"] == 0 and p["owned"] == copies // nranks * base[l]
                assert np.array_equal(p["gid"], np.arange(r * p["owned"], (r + 1) * p["owned"]))
        else:
            assert any(p["ghosts"] > 0 for p in plans)


@pytest.mark.parametrize("kind,dims", MESHES + [(0, [[26, 24, 22], [13, 12, 11], [7, 6, 6], [4, 3, 3]]), (0, [[16, 8, 8], [8, 4, 4], [2, 2, 2]])])
@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
@pytest.mark.parametrize("tile_nodes", [0, 128, 256])
def test_in_kernel_delivery_tables(kind, dims, nranks, tile_nodes):
    """The multi-GPU V-cycle has no exchange kernel: the kernel that produces a row stores it into the peers' ghost rows
    (DESIGN.md 5).  The tables behind that -- row -> (peer, remote row) from what the peers publish, the tile order with the
    ghost-reading tiles last, the transfer kernels' per-block wait flags -- are built here for every rank of the partition by the
    code mgcfd_dist_p2p_attach runs, and the delivery is replayed on global node ids: every ghost row of every rank and level is
    written exactly once, with the node it holds; no tile outside the last group reads a ghost row."""
    mesh = M.Mesh.generate(kind, dims, mesh_variant=0 if kind == 2 else 2)
    r = M.delivery_check(mesh, nranks, tile_nodes)
    assert r["errors"] == 0 and r["ghost_reading_tiles_not_last"] == 0, r
    assert r["delivered"] == r["ghost_rows"] > 0, r
    ghosts = sum(M.partition_plan(mesh, nranks, rk, l)["ghosts"] for rk in range(nranks) for l in range(mesh.levels))
    assert r["ghost_rows"] == ghosts
    if mesh.levels > 1:
        assert r["waiting_transfer_blocks"] > 0          # some restrict / prolong blocks read rows of other ranks, most do not
    one = M.delivery_check(mesh, 1, tile_nodes)
    assert one == dict(delivered=0, ghost_rows=0, errors=0, ghost_reading_tiles_not_last=0, waiting_transfer_blocks=0)


def _gloo_worker(rank, world, port, q, rank_local=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims = [[9, 8, 7], [5, 4, 4]]
        mesh = M.Mesh.generate(1, dims, mesh_variant=2)
        ok = True
        for l in range(mesh.levels):
            # rank_local: the way bench.py's ranks build their parts (no rank assembles the mesh)
            p = M.generate_partition_plan(1, dims, world, rank, l, mesh_variant=2) if rank_local else M.partition_plan(mesh, world, rank, l)
            mine = {"gid": p["gid"].tolist(), "owned": p["owned"], "send_counts": p["send_counts"].tolist(), "send_gids": p["send_gids"].tolist(),
                    "recv_counts": p["recv_counts"].tolist()}
            allp = [None] * world
            dist.all_gather_object(allp, mine)
            # the rank-to-rank contract, checked with the peers' own data: my ghosts from q == q's send list to me
            for qr in range(world):
                if qr == rank:
                    continue
                o = allp[qr]
                off = sum(o["send_counts"][:rank])
                their_send = o["send_gids"][off:off + o["send_counts"][rank]]
                start = p["owned"] + int(np.sum(p["recv_counts"][:qr]))
                my_ghosts = p["gid"][start:start + p["recv_counts"][qr]].tolist()
                ok = ok and (their_send == my_ghosts) and (o["recv_counts"][rank] == p["send_counts"][qr])
            total = sum(a["owned"] for a in allp)
            ok = ok and total == p["global_nodes"]
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rank_local", [False, True])
def test_partition_contract_across_two_gloo_ranks(rank_local):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + (7 if rank_local else 0)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q, rank_local)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_distributed_plans_are_unchanged(tmp_path):
    """Regression guard (host only): level plans and transfer operators of every rank of 2 / 3 / 8-rank partitions, hashed by
    tools/plan_hash_harness.cpp, equal the recorded hashes -- serial or threaded.  A deliberate change of the plan format
    updates tests/golden/plan_hashes.txt (see the harness header)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "mg-cfd-app-plain_b200", "csrc")
    exe = str(tmp_path / "plan_hash")
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-ffp-contract=off", "-fno-math-errno", "-I" + src, "-o", exe,
           os.path.join(root, "tools", "plan_hash_harness.cpp")] + [os.path.join(src, f) for f in ("plan.cpp", "partition.cpp", "mesh_gen.cpp", "mesh_io.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    want = open(os.path.join(root, "tests", "golden", "plan_hashes.txt")).read().split()
    for threads in ("1", "8"):
        out = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, MGCFD_PLAN_THREADS=threads))
        assert out.returncode == 0 and out.stdout.split() == want, out.stdout + out.stderr
