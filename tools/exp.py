"""One short GPU session for A/B work on the stage kernel: C2 (hex, 4 levels) and a 2.1 M-node tet box, default settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_probe import probe
import mgcfd_b200 as M
c2 = [[67] * 3, [55] * 3, [48] * 3, [43] * 3]
tet = [[129] * 3, [65] * 3, [33] * 3, [17] * 3]
probe("c2-hex", M.GEN_HEX_BOX, c2, 0, cycles=200, modes=(0, 1, 5))
probe("tet-2M", M.GEN_TET_BOX, tet, 0, cycles=20, modes=(0, 1, 5))
