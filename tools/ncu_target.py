"""Small target for ncu: a few launches of one kernel on the C2-shaped (or tet) level-0 mesh.
usage: ncu_target.py <which> <c2|tet> <tile_nodes> <reps> <flux_mode>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mgcfd_b200 as M
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
shape = sys.argv[2] if len(sys.argv) > 2 else "c2"
tn = int(sys.argv[3]) if len(sys.argv) > 3 else 128
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
fm = int(sys.argv[5]) if len(sys.argv) > 5 else 1
if shape == "c2":
    mesh = M.Mesh.generate(M.GEN_HEX_BOX, [[67, 67, 67], [34, 34, 34]], mesh_variant=M.MESH_M6_WING)
else:
    mesh = M.Mesh.generate(M.GEN_TET_BOX, [[101, 101, 101], [51, 51, 51]], mesh_variant=M.MESH_M6_WING)
s = M.Solver.from_mesh(mesh, tile_nodes=tn, flux_mode=fm)
s.time_kernel(0, which, 2)
print("ms per launch", s.time_kernel(0, which, reps) / reps)
s.close()
