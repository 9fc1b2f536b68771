"""Pins the CPU oracle (oracle/mgcfd_oracle.c, our plain-C restatement) to the UNMODIFIED reference:
 * against the committed golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
   reference's own sources) -- runs anywhere, including the GPU box where /root/reference does not exist;
 * live against oracle/_ref/libmgcfd_ref.so where that prebuilt checker is present.
Restrict, time_step, residual and the edge-weight adjustment are bit-exact; flux / step factor / prolong / whole
runs agree to ~1e-15 (FMA contraction is the only licence the two compilations differ by)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, linf_rel, load_golden, perturbed_state
from oracle.loader import Oracle, Reference, reference_available

TIGHT = 5e-14   # cancellation in the flux sums amplifies the 1-ulp contraction differences a little


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_far_field_and_ewt_bit_exact(orc, name):
    g, raw, adj = load_golden(name)
    ffv, ffc = orc.far_field()
    assert np.array_equal(ffv, g["ff_variable"]) and np.array_equal(ffc, g["ff_flux_contribution"])
    for r, a in zip(raw, adj):
        e = r["edges"].copy()
        orc.adjust_dampen(int(g["variant"]), r["coords"], e)
        assert e.tobytes() == a["edges"].tobytes()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_kernel_known_answers(orc, name):
    g, raw, adj = load_golden(name)
    L0, n = adj[0], adj[0]["nel"]
    var, old = g["kat_var"], g["kat_old"]
    flux = np.zeros(5 * n)
    orc.flux_edge(0, L0["nI"], L0["edges"], var, flux)
    assert np.all(linf_rel(flux, g["kat_flux_internal"]) < TIGHT)
    orc.boundary_flux_edge(L0["nI"], L0["nB"], L0["edges"], var, flux)
    assert np.all(linf_rel(flux, g["kat_flux_boundary"]) < TIGHT)
    orc.wall_flux_edge(L0["nI"] + L0["nB"], L0["nW"], L0["edges"], var, flux)
    assert np.all(linf_rel(flux, g["kat_flux_all"]) < TIGHT)
    for legacy, key in ((False, "kat_sf"), (True, "kat_sf_legacy")):
        sf = orc.step_factor(var, L0["vol"], legacy)
        assert np.max(np.abs(sf - g[key]) / g[key]) < TIGHT
    for j in range(3):
        f2, v2 = g["kat_flux_all"].copy(), np.zeros(5 * n)
        orc.time_step(j, g["kat_sf"], f2, old, v2)
        assert np.all(linf_rel(v2, g[f"kat_time_step_{j}"]) < TIGHT) and not f2.any()
    res = orc.residual(old, var)
    assert np.array_equal(res, g["kat_residual"])
    assert abs(orc.calc_rms(res) - float(g["kat_rms"])) <= 1e-15 * float(g["kat_rms"])
    # per-variable RMS (not a reference function): its squares must add up to calc_rms's
    pv = orc.rms_per_var(res)
    assert abs(np.sqrt(np.sum(pv ** 2)) - float(g["kat_rms"])) <= 1e-14 * float(g["kat_rms"])
    if len(adj) > 1:
        vc = g["kat_coarse_var"].copy()
        orc.mg_restrict(var, vc, L0["map"])
        assert np.array_equal(vc, g["kat_restrict"])
        v2 = var.copy()
        orc.prolong(L0["edges"], L0["nI"], g["kat_res1"], g["kat_res2"], v2, L0["map"], adj[1]["coords"], L0["coords"])
        assert np.all(linf_rel(v2, g["kat_prolong"]) < TIGHT)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_vcycle_runs(orc, name):
    g, raw, adj = load_golden(name)
    cycles = int(g["cycles"])
    ra, rv, st = orc.run_cycles(int(g["variant"]), adj, cycles)
    assert np.max(np.abs(ra - g["run_rms"]) / g["run_rms"]) < 1e-13
    assert np.max(np.abs(rv - g["run_rms_var"]) / np.maximum(g["run_rms_var"], 1e-300)) < 1e-11
    for l in range(len(adj)):
        assert np.all(linf_rel(st[l]["var"], g[f"run_L{l}_variables"]) < 1e-13), l
        assert np.all(linf_rel(st[l]["res"], g[f"run_L{l}_residuals"]) < 1e-9), l


def test_invalid_variable_scan(orc):
    var = perturbed_state(50, seed=1)
    assert orc.check_invalid(var) is None
    bad = var.copy(); bad[5 * 30 + 0] = -1.0; bad[5 * 12 + 4] = -3.0; bad[5 * 40 + 1] = np.inf
    assert orc.check_invalid(bad) == (12, 3)      # first offending cell, negative energy (validation.cpp:107-138)
    bad[5 * 3 + 2] = np.nan
    assert orc.check_invalid(bad) == (3, 1)


@pytest.mark.skipif(not reference_available(), reason="oracle/_ref/libmgcfd_ref.so not built (needs /root/reference)")
def test_oracle_against_live_reference_larger_mesh(orc):
    """A mesh bigger than the fixtures, 4 levels, 10 cycles: oracle vs the reference's own objects."""
    import mgcfd_b200 as M
    from conftest import mesh_levels
    ref = Reference()
    mesh = M.Mesh.generate(M.GEN_TET_BOX, [[17, 15, 13], [9, 8, 7], [5, 5, 4], [3, 3, 3]], mesh_variant=M.MESH_M6_WING)
    raw = mesh_levels(mesh)
    sess = ref.session(mesh.mesh_variant, raw)
    sess.prepare()
    ra, rv, _ = sess.run(10)
    adj = mesh_levels(mesh, apply_ewt_with=orc)
    oa, ov, st = orc.run_cycles(mesh.mesh_variant, adj, 10)
    assert np.max(np.abs(oa - ra) / ra) < 1e-13
    for l in range(4):
        assert adj[l]["edges"].tobytes() == sess.field(l, 6).tobytes()
        assert np.all(linf_rel(st[l]["var"], sess.field(l, 0)) < 1e-13)
    sess.close()


@pytest.mark.skipif(not reference_available(), reason="oracle/_ref/libmgcfd_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("variant", [4, 0])
def test_oracle_against_live_reference_unstructured_mesh(orc, variant):
    """Pins the oracle on inputs no generator produces: three unstructured levels (k-nearest-neighbour edges, hub nodes, nearest-node
    multigrid maps with childless coarse nodes and exact coincidences) run through the reference's own objects and the oracle."""
    from scipy.spatial import cKDTree
    from test_host_mesh import _random_level
    ref = Reference()
    rng = np.random.default_rng(variant)
    raw = [_random_level(1200, 6, 1), _random_level(500, 5, 2), _random_level(160, 4, 3)]
    for lv in raw:                                         # physical-ish scale so that six cycles stay well inside the valid states
        for f in ("x", "y", "z"):
            lv["edges"][f] *= 1e-2 if variant == 0 else 1.0
    for f, c in zip(raw[:-1], raw[1:]):
        f["coords"][:30] = c["coords"][rng.choice(c["nel"], 30)]
        f["map"] = cKDTree(c["coords"]).query(f["coords"])[1].astype(np.int64)
    sess = ref.session(variant, [dict(lv, edges=lv["edges"].copy()) for lv in raw])
    sess.prepare()
    ra, rv, _ = sess.run(6)
    adj = [dict(lv, edges=lv["edges"].copy()) for lv in raw]
    for lv in adj:
        orc.adjust_dampen(variant, lv["coords"], lv["edges"])
    oa, ov, st = orc.run_cycles(variant, adj, 6)
    assert np.all(np.isfinite(ra)) and np.max(np.abs(oa - ra) / ra) < 1e-13
    for l in range(3):
        assert adj[l]["edges"].tobytes() == sess.field(l, 6).tobytes()
        assert np.all(linf_rel(st[l]["var"], sess.field(l, 0)) < 1e-13)
    sess.close()
