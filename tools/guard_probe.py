"""Probe: where does a tile-size / guard-zone configuration leave the oracle?  (one GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mgcfd_b200 as M
from conftest import mesh_levels
from oracle.loader import Oracle
orc = Oracle()
dims = [[20, 18, 16], [13, 12, 11], [8, 7, 7], [5, 5, 4]]
for guard in (0, 1):
    os.environ["MGCFD_GUARD"] = str(guard)
    for tn in (0, 128, 256):
        for premin in (1, 0):
            os.environ["MGCFD_PREMIN"] = str(premin)
            for cycles in (1, 2, 5):
                mesh = M.Mesh.generate(0, dims, mesh_variant=2)
                lv = mesh_levels(mesh, apply_ewt_with=orc)
                ora, _, st = orc.run_cycles(mesh.mesh_variant, lv, cycles)
                s = M.Solver.from_mesh(mesh, tile_nodes=tn)
                ra, _ = s.run_cycles(cycles)
                out = []
                for l in range(mesh.levels):
                    got = s.get_field(l, M.FIELD_VARIABLES).reshape(-1, 5); want = st[l]["var"].reshape(-1, 5)
                    err = np.abs(got - want).max(axis=1)
                    bad = np.nonzero(err > 1e-15)[0]
                    out.append(f"L{l}: max {err.max():.1e} n>1e-15 {bad.size}/{err.size} first {bad[:4].tolist()}")
                sf = [float(np.abs(s.get_field(l, M.FIELD_STEP_FACTORS) - st[l]["sf"]).max()) if "sf" in st[l] else -1 for l in range(mesh.levels)]
                print(f"guard={guard} tn={tn} tile={s.level_info(0)['tile_nodes']} premin={premin} cycles={cycles} rms_err={np.max(np.abs(ra-ora)/ora):.1e} sf_err={sf} | " + " | ".join(out), flush=True)
                s.close(); mesh.close()
