import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def mesh_levels(mesh, apply_ewt_with=None):
    """numpy copies of every level of a mgcfd_b200.Mesh in the reference layout (dicts for oracle.loader)."""
    out = []
    for l in range(mesh.levels):
        nel, nI, nB, nW, mgc = mesh.dims(l)
        c, mp = mesh.coords(l), mesh.mg_map(l)
        lv = dict(nel=nel, nI=nI, nB=nB, nW=nW, vol=mesh.volumes(l).copy(), edges=mesh.edges(l).copy(),
                  coords=None if c is None else c.copy(), map=None if mp is None else mp.copy())
        if apply_ewt_with is not None:
            apply_ewt_with.adjust_dampen(mesh.mesh_variant, lv["coords"], lv["edges"])
        out.append(lv)
    return out


def linf_rel(test, ref):
    """per-variable max|test-ref| / max|ref| over AoS [n,5] arrays (SURVEY 8c comparison rule)."""
    t, r = np.asarray(test).reshape(-1, 5), np.asarray(ref).reshape(-1, 5)
    scale = np.max(np.abs(r), axis=0)
    scale[scale == 0] = 1.0
    return np.max(np.abs(t - r), axis=0) / scale


def perturbed_state(nel, seed):
    """smooth, physically valid, non-uniform state so that gathers are non-degenerate."""
    import mgcfd_b200 as M
    ffv, _ = M.far_field_conditions()
    rng = np.random.default_rng(seed)
    x = np.arange(nel) / max(nel, 1)
    var = np.tile(ffv, (nel, 1))
    var[:, 0] *= 1.0 + 0.05 * np.sin(7 * x) + 0.01 * rng.standard_normal(nel)
    var[:, 1] *= 1.0 + 0.05 * np.cos(5 * x)
    var[:, 2] = 0.05 * np.sin(11 * x) + 0.01 * rng.standard_normal(nel)
    var[:, 3] = 0.03 * np.cos(3 * x)
    var[:, 4] *= 1.0 + 0.04 * np.sin(2 * x + 1.0)
    return np.ascontiguousarray(var.reshape(-1))


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ("hex3_m6wing", "tet2_cascade", "hex4_nonnested_rotor", "fvcorr_cells", "hex2_random_m6wing")
# generator arguments of every golden case (tests/golden/make_golden.py): kind, dims, ordering
GOLDEN_SPECS = {
    "hex3_m6wing": (0, [[9, 9, 9], [5, 5, 5], [3, 3, 3]], 0),
    "tet2_cascade": (1, [[7, 6, 5], [4, 4, 3]], 0),
    "hex4_nonnested_rotor": (0, [[10, 9, 8], [7, 6, 6], [5, 4, 4], [3, 3, 3]], 0),
    "fvcorr_cells": (2, [[4, 3, 3]], 0),
    "hex2_random_m6wing": (0, [[8, 7, 6], [4, 4, 3]], 1),
}


def load_golden(name):
    """A tests/golden fixture (outputs of the unmodified reference, see make_golden.py) as
    (dict of arrays, raw levels, levels with the reference's adjusted edge weights)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    raw, adj = [], []
    for l in range(int(g["levels"])):
        lv = dict(nel=int(g[f"L{l}_nel"]), nI=int(g[f"L{l}_nI"]), nB=int(g[f"L{l}_nB"]), nW=int(g[f"L{l}_nW"]),
                  vol=g[f"L{l}_vol"], edges=g[f"L{l}_edges"], coords=g.get(f"L{l}_coords"), map=g.get(f"L{l}_map"))
        raw.append(lv)
        a = dict(lv)
        a["edges"] = g[f"L{l}_ewt_edges"]
        adj.append(a)
    return g, raw, adj
