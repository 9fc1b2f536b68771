"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libmgcfd_ref.so = the reference's own
src/Kernels/*.cpp + src/Base/io*.cpp compiled where they lie under /root/reference, driven through oracle/ref_shim.cpp).

Run in the build container (where /root/reference is mounted):
    make -C oracle ref REF=/root/reference && python tests/golden/make_golden.py

The reference ships no golden vectors of its own (its solution.* files live only in the GitHub release datasets,
README.md:66-71), so these fixtures are outputs of the reference itself on small synthetic meshes written in its
own in-memory layout.  Every fixture stores the INPUT mesh (raw, un-adjusted edge weights) next to the outputs, so the
tests that consume it need neither /root/reference nor the mesh generator to agree with anything.

Per case:
  mesh        : per level nel,nI,nB,nW, vol, edges (edge_neighbour AoS, 40 B), coords, map
  ewt_edges   : edges after the reference's adjust_ewt + dampen_ewt (validation.cpp:28-75)
  kat_*       : kernel-level known answers on a perturbed state (flux internal / +boundary / +wall, step factor
                (both variants), time_step for j=0..2, residual + calc_rms, mg_restrict, prolong)
  run_*       : `cycles` iterations of main()'s V-cycle loop (euler3d_cpu_double.cpp:371-694): printed RMS per cycle,
                per-variable RMS, final variables / residuals of every level
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mgcfd_b200 as M  # noqa: E402  (mesh generator only; host code, no GPU)
from conftest import mesh_levels, perturbed_state  # noqa: E402
from oracle.loader import Reference  # noqa: E402

CASES = {
    # name: kind, dims, variant, ordering, cycles
    "hex3_m6wing": (M.GEN_HEX_BOX, [[9, 9, 9], [5, 5, 5], [3, 3, 3]], M.MESH_M6_WING, 0, 8),
    "tet2_cascade": (M.GEN_TET_BOX, [[7, 6, 5], [4, 4, 3]], M.MESH_LA_CASCADE, 0, 6),
    "hex4_nonnested_rotor": (M.GEN_HEX_BOX, [[10, 9, 8], [7, 6, 6], [5, 4, 4], [3, 3, 3]], M.MESH_ROTOR_37, 0, 6),
    "fvcorr_cells": (M.GEN_TET_CELLS, [[4, 3, 3]], M.MESH_FVCORR, 0, 6),
    "hex2_random_m6wing": (M.GEN_HEX_BOX, [[8, 7, 6], [4, 4, 3]], M.MESH_M6_WING, 1, 6),
}


def main():
    ref = Reference()
    for name, (kind, dims, variant, ordering, cycles) in CASES.items():
        mesh = M.Mesh.generate(kind, dims, mesh_variant=variant, ordering=ordering)
        raw = mesh_levels(mesh)
        nl = len(raw)
        out = {"levels": nl, "variant": variant, "cycles": cycles}
        ref.set_globals(nl, variant)
        ffv, ffc = ref.far_field()
        out["ff_variable"], out["ff_flux_contribution"] = ffv, ffc
        adj = []
        for l, lv in enumerate(raw):
            for k in ("nel", "nI", "nB", "nW"):
                out[f"L{l}_{k}"] = lv[k]
            out[f"L{l}_vol"], out[f"L{l}_edges"] = lv["vol"], lv["edges"]
            if lv["coords"] is not None:
                out[f"L{l}_coords"] = lv["coords"]
            if lv["map"] is not None:
                out[f"L{l}_map"] = lv["map"]
            e = lv["edges"].copy()
            ref.adjust_dampen(variant, lv["coords"], e)
            adj.append(e)
            out[f"L{l}_ewt_edges"] = e
        # ---- kernel-level known answers on level 0 (and the 0->1 transfers) ----
        L0, e0 = raw[0], adj[0]
        n = L0["nel"]
        var = perturbed_state(n, seed=1)
        out["kat_var"] = var
        flux = np.zeros(5 * n)
        ref.flux_edge(0, L0["nI"], e0, var, flux)
        out["kat_flux_internal"] = flux.copy()
        ref.boundary_flux_edge(L0["nI"], L0["nB"], e0, var, flux)
        out["kat_flux_boundary"] = flux.copy()
        ref.wall_flux_edge(L0["nI"] + L0["nB"], L0["nW"], e0, var, flux)
        out["kat_flux_all"] = flux.copy()
        out["kat_sf"] = ref.step_factor(var, L0["vol"], False)
        out["kat_sf_legacy"] = ref.step_factor(var, L0["vol"], True)
        old = perturbed_state(n, seed=2)
        out["kat_old"] = old
        for j in range(3):
            f2, v2 = flux.copy(), np.zeros(5 * n)
            ref.time_step(j, out["kat_sf"], f2, old, v2)
            out[f"kat_time_step_{j}"] = v2
            assert not f2.any()
        res = ref.residual(old, var)
        out["kat_residual"] = res
        out["kat_rms"] = ref.calc_rms(res)
        if nl > 1:
            L1 = raw[1]
            vc = perturbed_state(L1["nel"], seed=3)
            out["kat_coarse_var"] = vc.copy()
            ref.mg_restrict(var, vc, L0["map"])
            out["kat_restrict"] = vc
            rng = np.random.default_rng(4)
            r1, r2 = 1e-3 * rng.standard_normal(5 * L1["nel"]), 1e-3 * rng.standard_normal(5 * n)
            out["kat_res1"], out["kat_res2"] = r1, r2
            v2 = var.copy()
            ref.prolong(e0, L0["nI"], r1, r2, v2, L0["map"], L1["coords"], L0["coords"])
            out["kat_prolong"] = v2
        # ---- the V-cycle loop ----
        sess = ref.session(variant, raw)
        sess.prepare()
        ra, rv, _ = sess.run(cycles)
        out["run_rms"], out["run_rms_var"] = ra, rv
        for l in range(nl):
            out[f"run_L{l}_variables"] = sess.field(l, 0)
            out[f"run_L{l}_residuals"] = sess.field(l, 2)
            assert np.array_equal(sess.field(l, 6), adj[l])
        sess.close()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {nl} levels, {n} fine nodes, rms[-1]={ra[-1]:.6e} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
