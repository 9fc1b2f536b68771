"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank runs its part of the mesh; rank 0 also runs the whole mesh on its own GPU and compares (1e-11, per-variable Linf)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import mgcfd_b200 as M


def bcast_id(rank):
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(M.dist_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(buf, 0)
    return bytes(buf.numpy().tobytes())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    # "coarse8": a coarsest level of 8 nodes (one or two per rank on 4 / 8 ranks); "replicas": the mesh duplicated once per rank (-m), dealt
    # out copy by copy -- NO rank has a halo at any level, yet every rank must still take part in every collective step (the epoch
    # numbers advance alike on all ranks)
    cases = [("hex4", 0, [[26, 24, 22], [13, 12, 11], [7, 6, 6], [4, 3, 3]], 2), ("tet3", 1, [[17, 15, 13], [9, 8, 7], [5, 4, 4]], 3), ("fvcorr", 2, [[9, 8, 7]], 0),
             ("coarse8", 0, [[16, 8, 8], [8, 4, 4], [2, 2, 2]], 2), ("replicas", 0, [[12, 10, 8], [6, 5, 4]], 2)]
    cycles = 8
    if len(sys.argv) > 1 and sys.argv[1] == "c2":       # the bench.py unit: BASELINE config C2 grown `world`-fold along x, rank-local generation
        cases = [("c2-unit", 0, [[67 * world - (world - 1), 67, 67], [55 * world - (world - 1), 55, 55], [48 * world - (world - 1), 48, 48], [43 * world - (world - 1), 43, 43]], 2)]
        cycles = 3
    if len(sys.argv) > 1 and sys.argv[1] == "big":      # one mid-size case: many tiles per persistent CTA, 100 KB-class messages
        cases = [("hex4-big", 0, [[133, 67, 67], [109, 55, 55], [95, 48, 48], [85, 43, 43]], 2)]
        cycles = 4
    if len(sys.argv) > 1 and sys.argv[1] == "tet":      # tets, 275 K nodes per rank: 256-node tiles, thousands of transfer-kernel blocks (two-level ticket)
        g = lambda n: n * world - (world - 1)
        cases = [("tet-mid", 1, [[g(65), 65, 65], [g(33), 33, 33], [g(17), 17, 17]], 2)]
        cycles = 4
    worst = 0.0
    for name, kind, dims, variant in cases:
        uid = bcast_id(rank)
        mesh = M.Mesh.generate(kind, dims, mesh_variant=variant)
        if name == "replicas":
            mesh.duplicate(world)
        if name in ("tet3", "hex4-big", "c2-unit", "tet-mid"):      # rank-local generation (no rank assembles the mesh): the path bench.py takes
            s = M.Solver.generate_distributed(kind, dims, rank, world, uid, mesh_variant=variant, device=local)
        else:
            s = M.Solver.from_mesh_distributed(mesh, rank, world, uid, device=local)
        if os.environ.get("MGCFD_NO_P2P", "0") != "1":      # direct peer-to-peer data path (CUDA IPC) instead of NCCL send/recv
            mine = s.p2p_prepare()
            allp = [None] * world
            dist.all_gather_object(allp, mine)
            s.p2p_attach([a[0] for a in allp], [a[1] for a in allp])
        ra, rv = s.run_cycles(cycles)
        pieces = []
        for l in range(mesh.levels):
            info = s.dist_level_info(l)
            var = s.get_field(l, M.FIELD_VARIABLES)[:info["owned"]]
            pieces.append((s.global_ids(l)[:info["owned"]], var.copy()))
        gathered = [None] * world
        dist.all_gather_object(gathered, pieces)
        ex = s.dist_level_info(0)["exchanges"]
        if os.environ.get("MGCFD_GUARD", "0") == "1":       # guard zones around every device array (mgcfd_guard_check): none may be damaged
            zones = [None] * world
            dist.all_gather_object(zones, M.guard_check())
            if rank == 0:
                nbad = sum(z[0] for z in zones)
                print(f"dist_check {name}: guard zones damaged on all ranks: {nbad}", "".join(z[1] for z in zones), flush=True)
                if nbad:
                    worst = float("inf")
        s.close()
        if rank == 0:
            rmesh = M.Mesh.generate(kind, dims, mesh_variant=variant)
            if name == "replicas":
                rmesh.duplicate(world)
            ref = M.Solver.from_mesh(rmesh, device=local)
            rra, rrv = ref.run_cycles(cycles)
            e_rms = float(np.max(np.abs(ra - rra) / rra))
            e_var = 0.0
            for l in range(mesh.levels):
                want = ref.get_field(l, M.FIELD_VARIABLES)
                got = np.full_like(want, np.nan)
                for p in gathered:
                    got[p[l][0]] = p[l][1]
                scale = np.max(np.abs(want), axis=0)
                e_var = max(e_var, float(np.max(np.max(np.abs(got - want), axis=0) / scale)))
            ref.close()
            print(f"dist_check {name}: {world} ranks, {cycles} cycles, exchanges={ex}, max rel err rms={e_rms:.2e} variables={e_var:.2e}", flush=True)
            worst = max(worst, e_rms, e_var)
    ok = torch.tensor([1 if worst < 1e-11 else 0])
    dist.broadcast(ok, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("dist_check", "PASS" if worst < 1e-11 else "FAIL", f"worst={worst:.2e}", flush=True)
    sys.exit(0 if int(ok) else 1)


if __name__ == "__main__":
    main()
