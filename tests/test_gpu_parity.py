"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerance (north star): per variable, max|gpu - oracle| <= 1e-11 * max|oracle| for fp64 fields and RMS histories.
Kernel-level checks are far tighter in practice (~1e-15); the assertions use 1e-12 there."""
import numpy as np
import pytest

from conftest import linf_rel, mesh_levels, perturbed_state

pytestmark = pytest.mark.gpu

TOL = 1e-11
KTOL = 1e-12


@pytest.fixture(scope="module")
def M():
    import mgcfd_b200 as M
    return M


@pytest.fixture(scope="module")
def oracle():
    from oracle.loader import Oracle
    return Oracle()


MESHES = {
    "hex3": dict(kind=0, dims=[[17, 17, 17], [9, 9, 9], [5, 5, 5]], variant=2),
    "tet3": dict(kind=1, dims=[[13, 12, 11], [7, 7, 6], [4, 4, 4]], variant=2),
    "hex_nonnested": dict(kind=0, dims=[[20, 18, 16], [13, 12, 11], [8, 7, 7], [5, 5, 4]], variant=2),
    "fvcorr": dict(kind=2, dims=[[8, 7, 6]], variant=0),
    "hex_random": dict(kind=0, dims=[[15, 14, 13], [8, 8, 7]], variant=3, ordering=1),
}


def make(M, name):
    s = MESHES[name]
    return M.Mesh.generate(s["kind"], s["dims"], mesh_variant=s["variant"], ordering=s.get("ordering", 0))


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("flux_mode", [0, 1, 2])
def test_flux_kernels_match_oracle(M, oracle, name, flux_mode):
    mesh = make(M, name)
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    s = M.Solver.from_mesh(mesh, flux_mode=flux_mode)
    for l in range(mesh.levels):
        L = lv[l]
        var = perturbed_state(L["nel"], seed=100 + l)
        s.set_field(l, M.FIELD_VARIABLES, var)
        s.zero_fluxes(l)
        want = np.zeros(5 * L["nel"])
        s.compute_flux_edge(l)
        oracle.flux_edge(0, L["nI"], L["edges"], var, want)
        got = s.get_field(l, M.FIELD_FLUXES)
        assert np.all(linf_rel(got, want) < KTOL), ("internal", l, linf_rel(got, want))
        s.compute_boundary_flux_edge(l)
        oracle.boundary_flux_edge(L["nI"], L["nB"], L["edges"], var, want)
        s.compute_wall_flux_edge(l)
        oracle.wall_flux_edge(L["nI"] + L["nB"], L["nW"], L["edges"], var, want)
        got = s.get_field(l, M.FIELD_FLUXES)
        assert np.all(linf_rel(got, want) < KTOL), ("all", l, linf_rel(got, want))
    s.close()


@pytest.mark.parametrize("name", ["hex3", "fvcorr"])
def test_node_kernels_match_oracle(M, oracle, name):
    mesh = make(M, name)
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    s = M.Solver.from_mesh(mesh)
    L = lv[0]
    n = L["nel"]
    var = perturbed_state(n, seed=7)
    s.set_field(0, M.FIELD_VARIABLES, var)
    for legacy in (False, True):
        s.compute_step_factor(0, legacy)
        got = s.get_field(0, M.FIELD_STEP_FACTORS)
        want = oracle.step_factor(var, L["vol"], legacy)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-14, legacy
    # time_step
    rng = np.random.default_rng(3)
    flux = rng.standard_normal(5 * n)
    old = perturbed_state(n, seed=8)
    sf = oracle.step_factor(var, L["vol"], False)
    for j in range(3):
        s.set_field(0, M.FIELD_FLUXES, flux)
        s.set_field(0, M.FIELD_OLD_VARIABLES, old)
        s.set_field(0, M.FIELD_STEP_FACTORS, sf)
        s.time_step(0, j)
        f2, v2 = flux.copy(), np.zeros(5 * n)
        oracle.time_step(j, sf, f2, old, v2)
        assert np.all(linf_rel(s.get_field(0, M.FIELD_VARIABLES), v2) < 1e-15)
        assert np.all(s.get_field(0, M.FIELD_FLUXES) == 0.0)
    # residual + rms
    s.set_field(0, M.FIELD_VARIABLES, var)
    s.set_field(0, M.FIELD_OLD_VARIABLES, old)
    s.residual(0)
    res = oracle.residual(old, var)
    assert np.array_equal(s.get_field(0, M.FIELD_RESIDUALS).reshape(-1), res)
    ra, rv = s.calc_rms(0)
    assert abs(ra - oracle.calc_rms(res)) / ra < 1e-13
    assert np.max(np.abs(rv - oracle.rms_per_var(res))) < 1e-13 * ra
    # copy + validity
    s.copy_old_variables(0)
    assert np.array_equal(s.get_field(0, M.FIELD_OLD_VARIABLES).reshape(-1), var)
    assert s.check_for_invalid_variables(0) is None
    bad = var.copy()
    bad[5 * 11 + 4] = -1.0
    bad[5 * 40 + 2] = np.nan
    bad[5 * 90 + 0] = -2.0
    s.set_field(0, M.FIELD_VARIABLES, bad)
    assert s.check_for_invalid_variables(0) == oracle.check_invalid(bad) == (11, 3)
    s.close()


@pytest.mark.parametrize("name", ["hex3", "tet3", "hex_nonnested", "hex_random"])
def test_mg_transfers_match_oracle(M, oracle, name):
    mesh = make(M, name)
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    s = M.Solver.from_mesh(mesh)
    rng = np.random.default_rng(5)
    for l in range(mesh.levels - 1):
        F, Cc = lv[l], lv[l + 1]
        vf = perturbed_state(F["nel"], seed=20 + l)
        vc = perturbed_state(Cc["nel"], seed=30 + l)
        s.set_field(l, M.FIELD_VARIABLES, vf)
        s.set_field(l + 1, M.FIELD_VARIABLES, vc)
        s.mg_restrict(l + 1)
        want = vc.copy()
        oracle.mg_restrict(vf, want, F["map"])
        # restrict sums children in the reference's order: bit-exact
        assert np.array_equal(s.get_field(l + 1, M.FIELD_VARIABLES).reshape(-1), want)
        r1 = 1e-3 * rng.standard_normal(5 * Cc["nel"])
        r2 = 1e-3 * rng.standard_normal(5 * F["nel"])
        s.set_field(l + 1, M.FIELD_RESIDUALS, r1)
        s.set_field(l, M.FIELD_RESIDUALS, r2)
        s.set_field(l, M.FIELD_VARIABLES, vf)
        s.prolong(l)
        want = vf.copy()
        oracle.prolong(F["edges"], F["nI"], r1, r2, want, F["map"], Cc["coords"], F["coords"])
        assert np.all(linf_rel(s.get_field(l, M.FIELD_VARIABLES), want) < 1e-14)
    s.close()


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("graph", [True, False])
def test_cycles_match_oracle(M, oracle, name, graph):
    cycles = 12
    mesh = make(M, name)
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    ora, orv, st = oracle.run_cycles(mesh.mesh_variant, lv, cycles)
    s = M.Solver.from_mesh(mesh, use_graph=graph)
    ra, rv = s.run_cycles(cycles)
    assert np.max(np.abs(ra - ora) / ora) < TOL
    assert np.max(np.abs(rv - orv) / np.maximum(orv, 1e-300)) < TOL
    for l in range(mesh.levels):
        assert np.all(linf_rel(s.get_field(l, M.FIELD_VARIABLES), st[l]["var"]) < TOL), l
    assert np.all(linf_rel(s.get_field(0, M.FIELD_RESIDUALS), st[0]["res"]) < 1e-9)
    # split calls continue the same trajectory (buffer roles / graph cache are consistent)
    s2 = M.Solver.from_mesh(make(M, name), use_graph=graph)
    a1, _ = s2.run_cycles(5)
    a2, _ = s2.run_cycles(cycles - 5)
    assert np.array_equal(np.concatenate([a1, a2]), ra)
    assert np.array_equal(s2.get_field(0, M.FIELD_VARIABLES), s.get_field(0, M.FIELD_VARIABLES))
    s.close(); s2.close()


@pytest.mark.parametrize("flux_mode", [0, 1, 2])
def test_granular_api_matches_oracle(M, oracle, flux_mode):
    from mgcfd_b200 import run_cycles_granular
    cycles = 4
    mesh = make(M, "hex_nonnested")
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    ora, orv, st = oracle.run_cycles(mesh.mesh_variant, lv, cycles)
    s = M.Solver.from_mesh(mesh, flux_mode=flux_mode)
    ra, rv = run_cycles_granular(s, cycles)
    assert np.max(np.abs(ra - ora) / ora) < TOL
    for l in range(mesh.levels):
        assert np.all(linf_rel(s.get_field(l, M.FIELD_VARIABLES), st[l]["var"]) < TOL)
    s.close()


def test_tiled_mode_is_deterministic_and_orderings_agree(M, oracle):
    mesh = make(M, "tet3")
    outs = []
    for ordering in (0, 1, 2, 2):
        s = M.Solver.from_mesh(make(M, "tet3"), ordering=ordering)
        s.run_cycles(6)
        outs.append(s.get_field(0, M.FIELD_VARIABLES).copy())
        s.close()
    assert np.array_equal(outs[2], outs[3])            # bit-reproducible run to run
    for o in outs[:2]:
        assert np.all(linf_rel(o, outs[2]) < TOL)      # renumbering only changes summation order


@pytest.mark.parametrize("tile_nodes", [128, 256, 512])
def test_tile_sizes(M, oracle, tile_nodes):
    mesh = make(M, "hex3")
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    ora, orv, st = oracle.run_cycles(mesh.mesh_variant, lv, 5)
    s = M.Solver.from_mesh(mesh, tile_nodes=tile_nodes)
    assert s.check_colouring(0) == 0
    ra, _ = s.run_cycles(5)
    assert np.max(np.abs(ra - ora) / ora) < TOL
    assert np.all(linf_rel(s.get_field(0, M.FIELD_VARIABLES), st[0]["var"]) < TOL)
    s.close()


def test_invalid_state_is_reported(M):
    mesh = make(M, "hex3")
    s = M.Solver.from_mesh(mesh)
    n = mesh.dims(0)[0]
    var = s.get_field(0, M.FIELD_VARIABLES).copy().reshape(-1)
    var[5 * 17 + 0] = np.nan
    s.set_field(0, M.FIELD_VARIABLES, var)
    with pytest.raises(M.MgcfdError) as e:
        s.run_cycles(1)
    assert e.value.code == 3
    s.close()


@pytest.mark.parametrize("name", ["hex_nonnested", "tet3", "fvcorr"])
@pytest.mark.parametrize("flux_mode", [0, 1])
@pytest.mark.parametrize("tile_nodes", [128, 256])
def test_pipelined_kernel_is_bit_identical_to_simple_kernel(M, name, flux_mode, tile_nodes):
    """the persistent TMA/cp.async pipeline only changes WHEN data moves, never the arithmetic or its order"""
    outs = []
    for pipeline in (True, False):
        s = M.Solver.from_mesh(make(M, name), flux_mode=flux_mode, tile_nodes=tile_nodes, pipeline=pipeline, visit=False)
        assert (s.level_info(0)["pipe_grid"] > 0) == pipeline and s.visit_info(0)["visit"] == 0
        ra, rv = s.run_cycles(7)
        outs.append((ra, rv, [s.get_field(l, M.FIELD_VARIABLES).copy() for l in range(s.levels)]))
        s.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2], outs[1][2]):
        assert np.array_equal(a, b)


def test_pipelined_kernel_many_tiles_per_cta(M, oracle):
    """a mesh with far more tiles than the persistent grid: every CTA walks a long tile sequence (ring wrap-around, parity flips)"""
    mesh = M.Mesh.generate(1, [[96, 80, 64], [48, 40, 32]], mesh_variant=2)
    s = M.Solver.from_mesh(mesh, tile_nodes=128, visit=False)
    info = s.level_info(0)
    assert info["pipe_grid"] > 0 and info["ntiles"] > 4 * info["pipe_grid"]
    s2 = M.Solver.from_mesh(mesh, tile_nodes=128, pipeline=False)
    ra, _ = s.run_cycles(3)
    rb, _ = s2.run_cycles(3)
    assert np.array_equal(ra, rb)
    assert np.array_equal(s.get_field(0, M.FIELD_VARIABLES), s2.get_field(0, M.FIELD_VARIABLES))
    s.close(); s2.close()


@pytest.mark.parametrize("workload,cycles", [("c1", 5), ("c2", 3)])
def test_baseline_configs_at_full_size_match_oracle(M, oracle, workload, cycles):
    """BASELINE.json configs[0] (fvcorr-shaped, 97.5 K cells, the reference's CPU-runnable validation case) and configs[1]
    (Onera-M6-shaped 4-level, 300 K fine nodes) at FULL size against the CPU oracle, 1e-11 per variable."""
    import bench
    kind, dims, variant, _ = bench.WORKLOADS[workload]
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant)
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    ora, orv, st = oracle.run_cycles(variant, lv, cycles)
    s = M.Solver.from_mesh(mesh)
    ra, rv = s.run_cycles(cycles)
    assert np.max(np.abs(ra - ora) / ora) < TOL
    assert np.max(np.abs(rv - orv) / np.maximum(orv, 1e-300)) < TOL
    for l in range(mesh.levels):
        assert np.all(linf_rel(s.get_field(l, M.FIELD_VARIABLES), st[l]["var"]) < TOL), l
    s.close()


def test_large_mesh_properties(M):
    """C3-class size (2.1 M-node tet box, 4 levels): size-independent properties instead of an oracle run --
    bit-reproducibility run to run, agreement of the two tiled flux modes (different summation orders) to 1e-11,
    a finite, non-growing RMS history and conservation of the uniform far-field state away from the walls."""
    dims = [[129] * 3, [65] * 3, [33] * 3, [17] * 3]
    outs = []
    for fm in (1, 1, 0):
        s = M.Solver.from_mesh(M.Mesh.generate(1, dims, mesh_variant=2), flux_mode=fm)
        ra, _ = s.run_cycles(4)
        outs.append((ra, s.get_field(0, M.FIELD_VARIABLES).copy()))
        s.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.max(np.abs(outs[0][0] - outs[2][0]) / outs[0][0]) < TOL
    assert np.all(linf_rel(outs[2][1], outs[0][1]) < TOL)
    ra = outs[0][0]
    assert np.all(np.isfinite(ra)) and np.all(ra > 0) and ra[-1] <= 1.01 * ra[0] and ra[0] < 1e-5
    ffv, _ = M.far_field_conditions()
    assert np.max(np.abs(outs[0][1] - ffv) / np.abs(ffv[0])) < 1e-3


def test_programmatic_dependent_launch_changes_nothing(M):
    """stage kernels launched with / without programmatic stream serialisation (and with / without the CUDA graph) are bit-identical"""
    outs = []
    for pdl, graph in ((True, True), (False, True), (True, False)):
        s = M.Solver.from_mesh(make(M, "hex_nonnested"), pdl=pdl, use_graph=graph)
        ra, _ = s.run_cycles(9)
        outs.append((ra, s.get_field(0, M.FIELD_VARIABLES).copy()))
        s.close()
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])


@pytest.mark.parametrize("name", ["hex3", "tet3", "fvcorr"])
def test_assess_compute_flux_variants_match_oracle(M, oracle, name):
    """The reference's FLUX_* arithmetic toggles as benchmark kernels (csrc/assess_kernels.cuh; their arithmetic is also checked on
    the CPU by tests/test_host_mesh.py): every variant accumulates compute_flux_edge's fluxes."""
    mesh = make(M, name)
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    s = M.Solver.from_mesh(mesh)
    L = lv[0]
    var = perturbed_state(L["nel"], seed=77)
    want = np.zeros(5 * L["nel"])
    oracle.flux_edge(0, L["nI"], L["edges"], var, want)
    s.set_field(0, M.FIELD_VARIABLES, var)
    for bits in range(9):          # 8: all toggles, node state gathered from SoA planes (the layout A/B)
        s.zero_fluxes(0)
        s.flux_variant(0, bits)
        got = s.get_field(0, M.FIELD_FLUXES)
        assert np.all(linf_rel(got, want) < KTOL), (bits, linf_rel(got, want))
    s.close()


@pytest.mark.parametrize("variant,tile_nodes", [(4, 0), (4, 128), (0, 256)])
def test_unstructured_levels_with_high_degree_nodes(M, oracle, variant, tile_nodes):
    """Unstructured input through mgcfd_upload_level: k-nearest-neighbour edges, hub nodes with 40+ neighbours (tiles with many edge
    rounds and several ring chunks), nearest-node multigrid maps -- V-cycles against the oracle."""
    from scipy.spatial import cKDTree
    from test_host_mesh import _random_level
    rng = np.random.default_rng(variant)
    raw = [_random_level(6000, 6, 1), _random_level(2500, 5, 2), _random_level(800, 4, 3)]
    for lv in raw:
        for f in ("x", "y", "z"):
            lv["edges"][f] *= 1e-2 if variant == 0 else 1.0
    for f, c in zip(raw[:-1], raw[1:]):
        f["map"] = cKDTree(c["coords"]).query(f["coords"])[1].astype(np.int64)
    adj = [dict(lv, edges=lv["edges"].copy()) for lv in raw]
    for lv in adj:
        oracle.adjust_dampen(variant, lv["coords"], lv["edges"])
    cycles = 5
    oa, ov, st = oracle.run_cycles(variant, [dict(lv, edges=lv["edges"].copy()) for lv in adj], cycles)
    s = M.Solver(3, variant, tile_nodes=tile_nodes)
    for l, lv in enumerate(adj):
        s.upload_level(l, lv["vol"], lv["coords"], lv["nI"], lv["nB"], lv["nW"], lv["edges"], lv.get("map"))
    s.finalize()
    assert s.level_info(0)["max_rounds"] > 40
    ra, rv = s.run_cycles(cycles)
    assert np.max(np.abs(ra - oa) / oa) < TOL
    for l in range(3):
        assert np.all(linf_rel(s.get_field(l, M.FIELD_VARIABLES), st[l]["var"]) < TOL), l
    s.close()


# ---- the persistent visit kernel (visit_kernel.cuh): one launch per smoothing visit -----------------------------------------------
def _env(**kw):
    import contextlib
    import os

    @contextlib.contextmanager
    def cm():
        old = {k: os.environ.get(k) for k in kw}
        os.environ.update({k: str(v) for k, v in kw.items()})
        try:
            yield
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return cm()


def test_launches_per_cycle_of_both_fused_paths(M):
    """stage path (default): 18 stage kernels + 3 restrictions + 3 prolongations + the RMS reduction = 25 launches per 4-level cycle once
    the transfer kernels supply the minimum dt (the first cycle also runs k_min_dt for level 0); visit path: 6 visits + 6 transfers = 12"""
    for visit, want in ((False, 25), (True, 12)):
        s = M.Solver.from_mesh(make(M, "hex_nonnested"), visit=visit)
        assert all(s.visit_info(l)["visit"] == int(visit) for l in range(4))
        s.run_cycles(1)
        l0 = s.launch_count()
        s.run_cycles(3)
        assert s.launch_count() - l0 == 3 * want, (visit, s.launch_count() - l0)
        s.close()


def test_min_dt_from_the_transfer_kernels_changes_nothing(M):
    """the per-block minima restrict / prolong leave behind give the same global minimum as k_min_dt: bit-identical runs"""
    outs = []
    for premin in (1, 0):
        with _env(MGCFD_PREMIN=premin):
            s = M.Solver.from_mesh(make(M, "hex_nonnested"))
            ra, rv = s.run_cycles(7)
            outs.append((ra, s.get_field(0, M.FIELD_VARIABLES).copy(), s.get_field(1, M.FIELD_STEP_FACTORS).copy()))
            s.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("name", ["hex_nonnested", "tet3", "fvcorr", "hex_random"])
@pytest.mark.parametrize("env", [dict(MGCFD_VISIT_K=1), dict(MGCFD_VISIT_K=1, MGCFD_VISIT_RESIDENT=0), dict(MGCFD_VISIT_K=2), dict(MGCFD_VISIT_K=3),
                                 dict(MGCFD_VISIT_K=1, MGCFD_VISIT_WARPS=8), dict(MGCFD_VISIT_K=2, MGCFD_VISIT_WARPS=8)])
def test_visit_kernel_configurations_match_oracle_and_stage_kernels(M, oracle, name, env):
    """every way the visit kernel can be configured -- own rows resident (K = 1) or re-read per stage, several double-buffered
    super-tiles per CTA (K > 1), 16 warps (one CTA per SM) or 8 (two) -- against the oracle (1e-11) and against the stage-per-launch
    path (different summation order inside a node's edge list only: 1e-13), and bit-reproducible run to run."""
    cycles = 6
    mesh = make(M, name)
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    ora, orv, st = oracle.run_cycles(mesh.mesh_variant, lv, cycles)
    outs = []
    with _env(**env):
        for rep in range(2):
            s = M.Solver.from_mesh(make(M, name), visit=True)
            vi = s.visit_info(0)
            if vi["visit"] == 0:
                s.close()
                pytest.skip("the level has fewer tiles than this configuration needs")
            assert vi["supers_per_cta"] == env["MGCFD_VISIT_K"] and vi["resident"] == (1 if env["MGCFD_VISIT_K"] == 1 and env.get("MGCFD_VISIT_RESIDENT", 1) else 0)
            ra, rv = s.run_cycles(cycles)
            outs.append((ra, rv, [s.get_field(l, M.FIELD_VARIABLES).copy() for l in range(s.levels)], s.get_field(0, M.FIELD_RESIDUALS).copy(),
                         s.get_field(0, M.FIELD_STEP_FACTORS).copy()))
            s.close()
    ra, rv, var, res, sf = outs[0]
    assert np.max(np.abs(ra - ora) / ora) < TOL and np.max(np.abs(rv - orv) / np.maximum(orv, 1e-300)) < TOL
    for l in range(mesh.levels):
        assert np.all(linf_rel(var[l], st[l]["var"]) < TOL), l
    assert np.all(linf_rel(res, st[0]["res"]) < 1e-9)
    assert np.array_equal(outs[1][0], ra) and all(np.array_equal(a, b) for a, b in zip(outs[1][2], var))
    s = M.Solver.from_mesh(make(M, name), visit=False)
    rb, _ = s.run_cycles(cycles)
    assert np.max(np.abs(ra - rb) / rb) < 1e-12
    for l in range(mesh.levels):
        assert np.all(linf_rel(var[l], s.get_field(l, M.FIELD_VARIABLES)) < 1e-12), l
    assert np.max(np.abs(sf - s.get_field(0, M.FIELD_STEP_FACTORS)) / np.abs(sf)) < 1e-13
    s.close()


def test_visit_kernel_reports_invalid_state_like_the_stage_kernels(M):
    mesh = make(M, "hex3")
    keys = []
    for visit in (True, False):
        s = M.Solver.from_mesh(make(M, "hex3"), visit=visit)
        var = s.get_field(0, M.FIELD_VARIABLES).copy().reshape(-1)
        var[5 * 1234 + 4] = -3.0
        s.set_field(0, M.FIELD_VARIABLES, var)
        with pytest.raises(M.MgcfdError) as e:
            s.run_cycles(2)
        assert e.value.code == 3
        keys.append(s.invalid_cell())
        s.close()
    assert keys[0] == keys[1] and keys[0][0] >= 0 and keys[0][1] in (1, 2, 3)     # same first offending cell and reason on both paths


def test_indirect_rw_matches_oracle(M, oracle):
    """a16: the bandwidth probe indirect_rw (src/Kernels/indirect_rw_loop.cpp:11-78, indirect_rw_kernel.elemfunc.c:4-94) through the C ABI
    against the oracle's restatement -- sums of plain copies, so any difference is summation order (atomics): 1e-13."""
    for name in ("hex3", "tet3", "fvcorr"):
        mesh = make(M, name)
        lv = mesh_levels(mesh, apply_ewt_with=oracle)
        s = M.Solver.from_mesh(mesh)
        L = lv[0]
        var = perturbed_state(L["nel"], seed=41)
        s.set_field(0, M.FIELD_VARIABLES, var)
        s.zero_fluxes(0)
        s.indirect_rw(0)
        want = np.zeros(5 * L["nel"])
        oracle.indirect_rw(0, L["nI"], L["edges"], var, want)
        got = s.get_field(0, M.FIELD_FLUXES)
        assert np.all(linf_rel(got, want) < 1e-13), (name, linf_rel(got, want))
        s.zero_fluxes(0)
        assert np.all(s.get_field(0, M.FIELD_FLUXES) == 0.0)
        s.close()


def test_c3_class_mesh_matches_the_serial_reference(M, oracle):
    """BASELINE.json configs[2] (8.1 M-node tet box, 4 levels; the 128-node-tile streaming path) against the UNMODIFIED reference
    compiled from its own sources (oracle/_ref/libmgcfd_ref.so, serial): 2 V-cycles, 1e-11 per variable.  Skipped when the reference
    build is absent."""
    import bench
    from oracle.loader import Reference, reference_available
    if not reference_available(omp=False):
        pytest.skip("oracle/_ref/libmgcfd_ref.so is not built")
    kind, dims, variant, _ = bench.WORKLOADS["c3"]
    cycles = 2
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant)
    ref = Reference(omp=False)
    sess = ref.session(variant, mesh_levels(mesh))      # RAW edge weights, so before Solver.from_mesh (the upload adjusts the mesh's in place, as the reference's loader does)
    s = M.Solver.from_mesh(mesh)
    assert s.level_info(0)["tile_nodes"] == 128 and s.level_info(0)["nel"] > 8000000
    ra, rv = s.run_cycles(cycles)
    got = [s.get_field(l, M.FIELD_VARIABLES).copy() for l in range(mesh.levels)]
    s.close()
    mesh.close()
    sess.prepare()
    rra, _, _ = sess.run(cycles)
    assert np.max(np.abs(ra - np.asarray(rra)) / np.asarray(rra)) < TOL
    for l in range(len(got)):
        assert np.all(linf_rel(got[l], sess.field(l, 0)) < TOL), l
    sess.close()


# ---- guard zones: the stand-in for a memcheck run (compute-sanitizer is not available on this GPU pool) ----------------------------
@pytest.mark.parametrize("name,kw", [("hex_nonnested", dict()), ("hex_nonnested", dict(tile_nodes=256)), ("hex_nonnested", dict(tile_nodes=128, flux_mode=0)),
                                     ("tet3", dict(tile_nodes=256)), ("tet3", dict(flux_mode=2)), ("fvcorr", dict()), ("hex_random", dict(tile_nodes=512)),
                                     ("hex_nonnested", dict(visit=True)), ("tet3", dict(visit=True))])
def test_no_kernel_writes_outside_its_arrays(M, oracle, name, kw):
    """MGCFD_GUARD=1 brackets every device allocation (and every sub-buffer of the node-state slab) with 64 KB of a byte pattern;
    after graph-replayed cycles, granular cycles, the flux variants and field transfers no zone may have changed -- and the results
    still match the oracle (the pattern also fills the allocations themselves: nothing reads memory it has not written).
    Round 2 found a real overrun this way: the per-block minima the transfer kernels leave (one per 128 rows) in an array sized
    one per 256 rows (levels with 256-node tiles)."""
    from mgcfd_b200 import run_cycles_granular
    with _env(MGCFD_GUARD=1):
        mesh = make(M, name)
        lv = mesh_levels(mesh, apply_ewt_with=oracle)
        ora, _, st = oracle.run_cycles(mesh.mesh_variant, lv, 5)
        s = M.Solver.from_mesh(mesh, **kw)
        ra, _ = s.run_cycles(5)
        assert np.max(np.abs(ra - ora) / ora) < TOL
        for l in range(mesh.levels):
            assert np.all(linf_rel(s.get_field(l, M.FIELD_VARIABLES), st[l]["var"]) < TOL)
        run_cycles_granular(s, 2)
        for l in range(mesh.levels):
            s.zero_fluxes(l); s.compute_flux_edge(l); s.compute_boundary_flux_edge(l); s.compute_wall_flux_edge(l)
            s.set_field(l, M.FIELD_VARIABLES, perturbed_state(lv[l]["nel"], seed=5 + l))
        s.run_cycles(2)
        bad, report = M.guard_check()
        s.close()
    assert bad == 0, report


def test_guard_zones_do_report_an_overrun(M):
    """the check itself: mgcfd_guard_selftest writes one byte 100 bytes past the end of the context's 64-byte RMS-sums array"""
    with _env(MGCFD_GUARD=1):
        s = M.Solver.from_mesh(make(M, "hex3"))
        s.run_cycles(1)
        assert M.guard_check()[0] == 0
        assert M.lib().mgcfd_guard_selftest(s._h) == 0
        bad, report = M.guard_check()
        s.close()
    assert bad == 1 and "offset 164" in report and "size 64" in report, report
    s = M.Solver.from_mesh(make(M, "hex3"))          # without MGCFD_GUARD there are no zones: the self-test refuses, the check returns 0
    assert M.lib().mgcfd_guard_selftest(s._h) != 0 and M.guard_check()[0] == 0
    s.close()


def test_two_contexts_keep_their_own_far_field(M, oracle):
    """the far-field state used to live in module-wide __constant__ symbols, so mgcfd_set_farfield on one context
    changed every context of the device.  It now travels in the kernel arguments: a second solver with another far field, driven
    in alternation (graph replays included), leaves the first one's results on the oracle's."""
    mesh = make(M, "hex_nonnested")
    lv = mesh_levels(mesh, apply_ewt_with=oracle)
    ora, _, st = oracle.run_cycles(mesh.mesh_variant, lv, 6)
    s1 = M.Solver.from_mesh(mesh)
    mesh2 = make(M, "hex_nonnested")
    s2 = M.Solver.from_mesh(mesh2)
    ffv, ffc = M.far_field_conditions()
    s2.set_farfield(ffv * np.array([1.0, 0.5, 0.5, 0.5, 1.0]), ffc * 0.5)
    ra1, ra2 = [], []
    for _ in range(3):
        ra1.append(s1.run_cycles(2)[0]); ra2.append(s2.run_cycles(2)[0])
    ra1, ra2 = np.concatenate(ra1), np.concatenate(ra2)
    assert np.max(np.abs(ra1 - ora) / ora) < TOL
    for l in range(mesh.levels):
        assert np.all(linf_rel(s1.get_field(l, M.FIELD_VARIABLES), st[l]["var"]) < TOL)
    assert np.max(np.abs(ra2 - ora) / ora) > 1e-3          # the other far field did take effect on the second solver
    s1.close(); s2.close()
