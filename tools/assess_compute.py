"""The reference's assess-compute / assess-memory protocols on the GPU: the flux kernel's arithmetic in the
forms the reference's FLUX_* toggles select, the production arithmetic (per-node derived quantities) and the indirect_rw
bandwidth probe, all as one-thread-per-edge kernels with identical memory traffic, on the lexicographic node numbering.
    python tools/assess_compute.py [c2|tet]  > profiles/<round>_assess_compute.jsonl"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mgcfd_b200 as M

NAMES = {2: "indirect_rw (memory ceiling)", 3: "production arithmetic, atomic scatter"}
for bits in range(8):
    NAMES[16 + bits] = "reference arithmetic" + "".join(t for k, t in ((1, " +REUSE_DIV"), (2, " +REUSE_FACTOR/FLUX"), (4, " +PRECOMPUTE_EDGE_WEIGHTS")) if bits & k)
NAMES[24] = "reference arithmetic, all toggles, node state gathered from SoA planes instead of 64-byte records"
what = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, dims = (M.GEN_HEX_BOX, [[67] * 3]) if what == "c2" else (M.GEN_TET_BOX, [[129] * 3])
ORDERINGS = {"as given (lexicographic)": (0, M.ORDER_AS_GIVEN), "partition + RCM": (0, M.ORDER_PARTITION_RCM), "random": (1, M.ORDER_AS_GIVEN)}
for oname, (mesh_ordering, ordering) in ORDERINGS.items():
    mesh = M.Mesh.generate(kind, dims, mesh_variant=M.MESH_M6_WING, ordering=mesh_ordering)
    s = M.Solver.from_mesh(mesh, flux_mode=M.FLUX_ATOMIC, ordering=ordering)
    info = s.level_info(0)
    for which in sorted(NAMES):
        if oname != "as given (lexicographic)" and which not in (2, 3, 23, 24):      # the toggles once; the layout A/B on every numbering
            continue
        s.time_kernel(0, which, 3)
        reps = 20
        ms = s.time_kernel(0, which, reps) / reps
        print(json.dumps({"mesh": what, "nel": info["nel"], "nI": info["nI"], "numbering": oname, "kernel": NAMES[which], "selector": which,
                          "us": round(ms * 1e3, 2), "Gedge/s": round(info["nI"] / ms / 1e6, 2)}), flush=True)
    s.close(); mesh.close()
