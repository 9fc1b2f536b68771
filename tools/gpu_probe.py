"""Exploratory timing of the kernels on a B200 (not the bench contract; see bench.py)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mgcfd_b200 as M

PEAK = 6548.5e9
def probe(name, kind, dims, tn, ordering=2, cycles=20, modes=(0, 1), flux_mode=1, pf=True):
    t = time.time()
    mesh = M.Mesh.generate(kind, dims, mesh_variant=M.MESH_M6_WING)
    tg = time.time() - t
    t = time.time()
    s = M.Solver.from_mesh(mesh, tile_nodes=tn, ordering=ordering, flux_mode=flux_mode, pipeline=pf)
    tu = time.time() - t
    info = s.level_info(0)
    nel, nI, nB, nW = info["nel"], info["nI"], info["nB"], info["nW"]
    alg = 32 * nI + 28 * (nB + nW) + 128 * nel
    out = {"mesh": name, "flux_mode": flux_mode, "pipeline": pf, "tile": tn, "ordering": ordering, "nel": nel, "nI": nI, "gen_s": round(tg, 2), "setup_s": round(tu, 2),
           "rounds": info["max_rounds"], "slot_util": round(info["used_slots"] / max(info["slots"], 1), 3), "smem": info["smem_bytes"]}
    names = {0: "fused_stage", 1: "tile_flux_only", 2: "indirect_rw", 3: "flux_atomic", 5: "stage_without_edges"}
    for which in modes:
        s.time_kernel(0, which, 3)
        reps = 20
        ms = s.time_kernel(0, which, reps) / reps
        out[names[which] + "_us"] = round(ms * 1e3, 2)
        out[names[which] + "_Gedge/s"] = round(nI / ms / 1e6, 2)
        if which == 0:
            out["fused_frac_hbm(32E+128N)"] = round(alg / (ms * 1e-3) / PEAK, 3)
    s.run_cycles(3)
    t = time.time()
    s.run_cycles(cycles)
    dt = time.time() - t
    out["cycles/s"] = round(cycles / dt, 1)
    print(json.dumps(out), flush=True)
    s.close()

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "quick"
    c2 = [[67, 67, 67], [55, 55, 55], [48, 48, 48], [43, 43, 43]]
    tet = [[129, 129, 129], [65, 65, 65], [33, 33, 33], [17, 17, 17]]
    tet8 = [[201, 201, 201], [101, 101, 101], [51, 51, 51], [26, 26, 26]]
    if what == "quick":
        for fm, tn, pf in ((1, 128, True), (1, 128, False), (1, 256, True), (0, 128, True), (0, 256, True)):
            probe("c2-hex", M.GEN_HEX_BOX, c2, tn, flux_mode=fm, pf=pf)
        for fm, tn, pf in ((1, 128, True), (1, 128, False), (1, 256, True), (0, 128, True)):
            probe("tet-2M", M.GEN_TET_BOX, tet, tn, cycles=5, flux_mode=fm, pf=pf)
    elif what == "overhead":
        for tn in (128, 256):
            probe("c2-hex", M.GEN_HEX_BOX, c2, tn, modes=(0, 5))
            probe("tet-2M", M.GEN_TET_BOX, tet, tn, cycles=5, modes=(0, 5))
    elif what == "orderings":
        for o in (0, 1, 2):
            probe("c2-hex-order%d" % o, M.GEN_HEX_BOX, c2, 128, ordering=o, modes=(0, 1, 2, 3))
    elif what == "big":
        for fm in (1, 0):
            probe("c3-tet-8M", M.GEN_TET_BOX, tet8, 128, cycles=3, flux_mode=fm)
