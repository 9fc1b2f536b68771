"""BASELINE.json config 5: flux-kernel-only sweep over node orderings (in the style of the reference's assess-memory /
assess-compute runs, run-inputs/assess-*.json): the same level-0 mesh numbered four ways -- lexicographic ("original"),
randomly shuffled on input (seeded Fisher-Yates in the generator), reverse Cuthill-McKee, partition + CM (default) -- against
the kernels: fused stage (tiled sorted-segment), its flux-only form, atomic flux (one thread per edge in original edge order)
and indirect_rw (the reference's bandwidth probe, indirect_rw_loop.cpp:11-78).  One JSON line per combination."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mgcfd_b200 as M

PEAK = 6548.5e9
SHAPES = {"c2": (M.GEN_HEX_BOX, [[67] * 3, [34] * 3]), "tet1m": (M.GEN_TET_BOX, [[101] * 3, [51] * 3])}
NAMES = {0: "fused_stage", 1: "flux_only", 2: "indirect_rw", 3: "flux_atomic"}


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "c2"
    kind, dims = SHAPES[shape]
    for mesh_order, mesh_name in ((0, "original"), (1, "random")):
        for ordering, oname in ((M.ORDER_AS_GIVEN, "as_given"), (M.ORDER_RCM, "rcm"), (M.ORDER_PARTITION_RCM, "partition_rcm")):
            mesh = M.Mesh.generate(kind, dims, mesh_variant=M.MESH_M6_WING, ordering=mesh_order, seed=12345)
            out = {"mesh": shape, "input_numbering": mesh_name, "renumbering": oname}
            try:
                s = M.Solver.from_mesh(mesh, ordering=ordering)
            except M.MgcfdError as e:          # a shuffled mesh taken as given cannot be tiled (halo of a tile ~ the whole mesh)
                out["tiled_kernels"] = "not applicable: " + str(e)[:90]
                mesh = M.Mesh.generate(kind, dims, mesh_variant=M.MESH_M6_WING, ordering=mesh_order, seed=12345)
                s = M.Solver.from_mesh(mesh, ordering=ordering, flux_mode=M.FLUX_ATOMIC)
            info = s.level_info(0)
            nI, nel = info["nI"], info["nel"]
            out.update(nodes=nel, edges=nI, max_halo=info["max_halo"], halo_entries=info["halo_entries"])
            for which in (0, 1, 2, 3):
                try:
                    s.time_kernel(0, which, 3)
                    ms = s.time_kernel(0, which, 20) / 20
                except M.MgcfdError:
                    continue
                out[NAMES[which] + "_us"] = round(ms * 1e3, 2)
                out[NAMES[which] + "_Gedges/s"] = round(nI / ms / 1e6, 2)
                if which == 0:
                    out["fused_frac_hbm"] = round((32 * nI + 128 * nel) / (ms * 1e-3) / PEAK, 3)
            print(json.dumps(out), flush=True)
            s.close()


if __name__ == "__main__":
    main()
