"""One short GPU session for A/B work on the stage kernel: C2 (hex, 4 levels) and a 2.1 M-node tet box.
usage: exp.py <tile_nodes for c2> <tile_nodes for tets>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_probe import probe
import mgcfd_b200 as M
c2 = [[67] * 3, [55] * 3, [48] * 3, [43] * 3]
tet = [[129] * 3, [65] * 3, [33] * 3, [17] * 3]
t2 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if t2 >= 0: probe("c2-hex", M.GEN_HEX_BOX, c2, t2, cycles=300, modes=(0,))
if tt >= 0: probe("tet-2M", M.GEN_TET_BOX, tet, tt, cycles=30, modes=(0,))
