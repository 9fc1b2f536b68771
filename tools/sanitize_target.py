"""Small target for compute-sanitizer (memcheck / racecheck / synccheck): a few V-cycles on small meshes through every fused path --
the visit kernel (resident and streaming configurations) and the stage kernels (both tile sizes, both scatter modes).
usage: sanitize_target.py [visit|stage|all] [cycles]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mgcfd_b200 as M

what = sys.argv[1] if len(sys.argv) > 1 else "all"
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dims = [[21, 19, 17], [11, 10, 9], [6, 5, 5]]


def run(tag, **kw):
    env = kw.pop("env", {})
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        s = M.Solver.from_mesh(M.Mesh.generate(M.GEN_HEX_BOX, dims, mesh_variant=M.MESH_M6_WING), use_graph=False, **kw)
        ra, _ = s.run_cycles(cycles)
        print(tag, "visit" if s.visit_info(0)["visit"] else "stage", s.visit_info(0), "rms", ra[-1], flush=True)
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return ra


outs = []
if what in ("visit", "all"):
    outs.append(run("visit K=1 resident", visit=True))
    outs.append(run("visit K=1 reload", visit=True, env=dict(MGCFD_VISIT_RESIDENT=0)))
    outs.append(run("visit K=2 8 warps", visit=True, env=dict(MGCFD_VISIT_K=2, MGCFD_VISIT_WARPS=8)))
    outs.append(run("visit K=3", visit=True, env=dict(MGCFD_VISIT_K=3)))
if what in ("stage", "all"):
    outs.append(run("stage TN=128 segment", visit=False, tile_nodes=128))
    outs.append(run("stage TN=256 segment", visit=False, tile_nodes=256))
    outs.append(run("stage TN=128 coloured", visit=False, tile_nodes=128, flux_mode=M.FLUX_TILED_COLOURED))
for o in outs[1:]:
    assert np.max(np.abs(o - outs[0]) / outs[0]) < 1e-11, (o, outs[0])
print("sanitize_target done", flush=True)
