"""Host-side code (no GPU): the synthetic generators, the reference's mesh interchange formats, and the integer
preprocessing (renumbering, tiling, colouring).  Integer work is checked bit-exactly: edge lists and MG
maps must equal what the reference's read_grid / read_mg_connectivity produce from the same files."""
import os
import subprocess

import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import pytest

import mgcfd_b200 as M
from conftest import GOLDEN_CASES, GOLDEN_SPECS, ROOT, load_golden, mesh_levels
from oracle.loader import HERE as ORACLE_DIR
from oracle.loader import Reference, reference_available


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_generator_is_deterministic_and_matches_fixture_mesh(name):
    g, raw, adj = load_golden(name)
    kind, dims, ordering = GOLDEN_SPECS[name]
    mesh = M.Mesh.generate(kind, dims, mesh_variant=int(g["variant"]), ordering=ordering)
    lv = mesh_levels(mesh)
    assert len(lv) == len(raw)
    for a, b in zip(lv, raw):
        assert (a["nel"], a["nI"], a["nB"], a["nW"]) == (b["nel"], b["nI"], b["nB"], b["nW"])
        assert a["edges"].tobytes() == b["edges"].tobytes()
        assert np.array_equal(a["vol"], b["vol"])
        if b["coords"] is not None:
            assert np.array_equal(a["coords"], b["coords"])
        if b["map"] is not None:
            assert np.array_equal(a["map"], b["map"])
    # adjust_ewt + dampen_ewt on the host mesh == the reference's (bit for bit)
    mesh.apply_ewt()
    for l, a in enumerate(adj):
        assert mesh.edges(l).tobytes() == a["edges"].tobytes()


def test_mesh_structure_invariants():
    """constraints the reference's loader and kernels rely on."""
    for kind, dims in ((0, [[9, 8, 7], [5, 4, 4]]), (1, [[6, 6, 5], [3, 3, 3]]), (2, [[3, 3, 2]])):
        mesh = M.Mesh.generate(kind, dims, mesh_variant=0 if kind == 2 else 2)
        for l in range(mesh.levels):
            nel, nI, nB, nW, mgc = mesh.dims(l)
            e = mesh.edges(l)
            assert np.all(e["a"][:nI] >= 0) and np.all(e["a"][:nI] < e["b"][:nI]) and np.all(e["b"] < nel)
            assert np.all(e["a"][nI:nI + nB] == -1) and np.all(e["a"][nI + nB:] == -2)
            deg = np.bincount(np.concatenate([e["a"][:nI], e["b"][:nI]]), minlength=nel)
            assert deg.min() >= 1                       # prolong divides by the weight sum of internal edges
            assert np.all(mesh.volumes(l) > 0)
            if l + 1 < mesh.levels:
                assert mgc == nel
                assert mesh.mg_map(l).min() >= 0 and mesh.mg_map(l).max() < mesh.dims(l + 1)[0]
        # closed control volumes: the signed face vectors around every node sum to ~0 (raw weights)
        nel, nI, nB, nW, _ = mesh.dims(0)
        e = mesh.edges(0)
        s = np.zeros((nel, 3))
        w = np.stack([e["x"], e["y"], e["z"]], axis=1)
        np.add.at(s, e["a"][:nI], w[:nI])
        np.add.at(s, e["b"][:nI], -w[:nI])
        if kind != 2:                                    # boundary weights are inward for non-fvcorr meshes
            np.add.at(s, e["b"][nI:], -w[nI:])
            tilt = np.abs(s).max()
            assert tilt < 0.06 * np.abs(w).max()         # only the tilted wall patch is not closed


@pytest.mark.parametrize("binary", [False, True])
def test_write_then_load_round_trip(tmp_path, binary):
    mesh = M.Mesh.generate(M.GEN_HEX_BOX, [[8, 7, 6], [4, 4, 3], [2, 2, 2]], mesh_variant=M.MESH_ROTOR_37)
    mesh.write(str(tmp_path), "input.dat", binary=binary)
    if binary:                                           # the .bin cache is preferred: remove the text so it has to be used
        for f in os.listdir(tmp_path):
            if not f.endswith((".bin", ".dat")):
                os.remove(tmp_path / f)
    back = M.Mesh.load("input.dat", str(tmp_path))
    assert back.levels == mesh.levels and back.mesh_variant == mesh.mesh_variant
    for a, b in zip(mesh_levels(mesh), mesh_levels(back)):
        assert (a["nel"], a["nI"], a["nB"], a["nW"]) == (b["nel"], b["nI"], b["nB"], b["nW"])
        assert a["edges"].tobytes() == b["edges"].tobytes()
        assert np.array_equal(a["vol"], b["vol"]) and np.array_equal(a["coords"], b["coords"])
        assert (a["map"] is None and b["map"] is None) or np.array_equal(a["map"], b["map"])


def test_load_errors_are_reported(tmp_path):
    with pytest.raises(M.MgcfdError):
        M.Mesh.load("does-not-exist.dat", str(tmp_path))
    (tmp_path / "input.dat").write_text("size = 1\nnum_levels = 1\nmesh_name = m6wing\n[levels]\n0 = missing.mesh\n")
    with pytest.raises(M.MgcfdError):
        M.Mesh.load("input.dat", str(tmp_path))


def test_corrupt_input_is_an_error_never_a_crash(tmp_path):
    """A damaged .bin cache (header offsets or sizes pointing outside the file), an absurd num_levels and a multigrid map that
    points outside the coarse level all come back as MgcfdError through the C ABI -- no exception, no out-of-bounds access."""
    import subprocess, sys, textwrap
    mesh = M.Mesh.generate(M.GEN_HEX_BOX, [[6, 5, 4], [3, 3, 2]], mesh_variant=M.MESH_M6_WING)
    mesh.write(str(tmp_path), "input.dat", binary=True)
    bins = sorted(f for f in os.listdir(tmp_path) if f.endswith(".bin"))
    good = (tmp_path / bins[0]).read_bytes()
    texts = [f for f in os.listdir(tmp_path) if not f.endswith((".bin", ".dat"))]
    saved = {f: (tmp_path / f).read_bytes() for f in texts}
    for f in texts:                                      # force the .bin cache to be the only source
        os.remove(tmp_path / f)
    assert M.Mesh.load("input.dat", str(tmp_path)).levels == 2
    hdr = np.frombuffer(good[:64], dtype=np.int64).copy()
    # each damaged file is loaded in a child process: a crash would show up as a signal, not as an exception here
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        import mgcfd_b200 as M
        try:
            M.Mesh.load("input.dat", sys.argv[1])
            print("LOADED")
        except M.MgcfdError as e:
            print("ERROR", e)
    """) % ROOT
    for k, v in ((5, 10 ** 9), (6, hdr[1] + 1), (0, 2 ** 40), (1, 2 ** 40), (2, hdr[1] + 5), (7, 2 ** 62)):
        h = hdr.copy()
        h[k] = v
        (tmp_path / bins[0]).write_bytes(h.tobytes() + good[64:])
        r = subprocess.run([sys.executable, "-c", code, str(tmp_path)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.startswith("ERROR"), (k, v, r.returncode, r.stdout, r.stderr[-500:])
    (tmp_path / bins[0]).write_bytes(good[:len(good) // 2])                                   # truncated
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ERROR")
    (tmp_path / bins[0]).write_bytes(good)
    # num_levels out of range
    dat = (tmp_path / "input.dat").read_text()
    import re
    (tmp_path / "input.dat").write_text(re.sub(r"num_levels\s*=\s*\d+", "num_levels = -1", dat))
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ERROR") and "num_levels" in r.stdout, (r.stdout, r.stderr[-500:])
    (tmp_path / "input.dat").write_text(dat)
    # a multigrid map that points outside the coarse level (text files, no cache)
    for f in bins:
        os.remove(tmp_path / f)
    for f, b in saved.items():
        (tmp_path / f).write_bytes(b)
    mgf = [f for f in texts if f.endswith(".mg")][0]
    toks = (tmp_path / mgf).read_text().split()
    toks[3] = "99999"
    (tmp_path / mgf).write_text(" ".join(toks))
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ERROR") and "multigrid map" in r.stdout, (r.stdout, r.stderr[-500:])


def test_text_formats_as_other_tools_write_them(tmp_path):
    """The text mesh is whitespace-separated numbers (the reference reads it with operator>>, io.cpp:56-137): fixed and
    scientific notation, explicit '+' signs, tabs, CRLF and blank lines all parse to the same arrays; a damaged or
    truncated file is reported instead of yielding garbage."""
    mesh = M.Mesh.generate(M.GEN_TET_CELLS, [[2, 2, 1]], mesh_variant=M.MESH_FVCORR)
    mesh.write(str(tmp_path))
    ref = mesh_levels(M.Mesh.load("input.dat", str(tmp_path)))[0]
    name = [f for f in os.listdir(tmp_path) if f.endswith(".dat") and f != "input.dat"][0]
    toks = (tmp_path / name).read_text().split()
    restyled = []
    for k, t in enumerate(toks):
        if "." in t or "e" in t:
            v = float(t)
            restyled.append(("+" if v >= 0 and k % 3 == 0 else "") + ("%.17e" % v if k % 2 else repr(v)))
        else:
            restyled.append(t)
    (tmp_path / name).write_text("\r\n\t ".join(restyled) + "\r\n\r\n")
    again = mesh_levels(M.Mesh.load("input.dat", str(tmp_path)))[0]
    assert again["edges"].tobytes() == ref["edges"].tobytes() and np.array_equal(again["vol"], ref["vol"])
    (tmp_path / name).write_text(" ".join(toks[:len(toks) // 2]))                     # truncated
    with pytest.raises(M.MgcfdError, match="Corruption"):
        M.Mesh.load("input.dat", str(tmp_path))
    broken = list(toks)
    broken[7] = "1.0.0x"
    (tmp_path / name).write_text(" ".join(broken))
    with pytest.raises(M.MgcfdError, match="Corruption"):
        M.Mesh.load("input.dat", str(tmp_path))


@pytest.mark.skipif(not reference_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,dims,variant", [(0, [[7, 6, 5], [4, 3, 3]], 2), (1, [[5, 5, 4], [3, 3, 2]], 3), (2, [[3, 3, 2]], 0)])
def test_files_read_by_the_reference_loader_give_our_arrays(tmp_path, kind, dims, variant):
    """our writer -> the reference's read_input_dat / read_grid / read_mg_connectivity (io.cpp:14-199,
    io_enhanced.cpp:407-650) -> identical in-memory arrays (edge order, orientation, sign flips, classes)."""
    ref = Reference()
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant)
    mesh.write(str(tmp_path), "input.dat")
    info = ref.read_input_dat(str(tmp_path / "input.dat"))
    assert info["levels"] == mesh.levels and info["variant"] == variant
    for l, a in enumerate(mesh_levels(mesh)):
        r = ref.read_grid(str(tmp_path / info["layers"][l]), mesh.levels, variant)
        assert (r["nel"], r["nI"], r["nB"], r["nW"]) == (a["nel"], a["nI"], a["nB"], a["nW"])
        assert r["edges"].tobytes() == a["edges"].tobytes()
        assert np.array_equal(r["vol"], a["vol"])
        if mesh.levels > 1:
            assert np.array_equal(r["coords"], a["coords"])
        if l + 1 < mesh.levels:
            assert np.array_equal(ref.read_mg(str(tmp_path / info["mg"][l])), a["map"])


@pytest.mark.skipif(not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "euler3d_ref.b")), reason="reference binary not built")
def test_reference_binary_runs_our_files_and_matches_oracle(tmp_path):
    """the reference's own main() on files we wrote: its printed RMS history equals the oracle's to print precision."""
    from oracle.loader import Oracle
    mesh = M.Mesh.generate(M.GEN_HEX_BOX, [[9, 9, 9], [5, 5, 5], [3, 3, 3]], mesh_variant=M.MESH_M6_WING)
    mesh.write(str(tmp_path), "input.dat")
    exe = os.path.join(ORACLE_DIR, "_ref", "euler3d_ref.b")
    out = subprocess.run([exe, "-i", "input.dat", "-d", str(tmp_path) + "/", "-g", "4", "-o", str(tmp_path) + "/", "--output-variables"],
                         capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rms = [float(line.split("RMS =")[1].split(")")[0]) for line in out.stdout.splitlines() if "RMS =" in line]
    orc = Oracle()
    oa, _, st = orc.run_cycles(mesh.mesh_variant, mesh_levels(mesh, apply_ewt_with=orc), 4)
    assert len(rms) == 4 and np.allclose(rms, oa, rtol=6e-4)        # printed with %.3e
    dump = [f for f in os.listdir(tmp_path) if f.startswith("variables")]
    if dump:
        vals = np.loadtxt(tmp_path / dump[0], skiprows=1).reshape(-1)
        if vals.size == st[0]["var"].size:
            assert np.max(np.abs(vals - st[0]["var"])) < 1e-13


@pytest.mark.parametrize("kind,dims", [(0, [[21, 20, 19]]), (1, [[14, 13, 12]]), (2, [[6, 5, 5]])])
@pytest.mark.parametrize("ordering", [M.ORDER_AS_GIVEN, M.ORDER_RCM, M.ORDER_PARTITION_RCM])
@pytest.mark.parametrize("tile_nodes", [128, 256])
@pytest.mark.parametrize("flux_mode", [M.FLUX_TILED_COLOURED, M.FLUX_SORTED_SEGMENT])
def test_plan_invariants(kind, dims, ordering, tile_nodes, flux_mode):
    mesh = M.Mesh.generate(kind, dims, mesh_variant=0 if kind == 2 else 2)
    info, perm, conflicts = M.plan_level(mesh, 0, ordering=ordering, tile_nodes=tile_nodes, flux_mode=flux_mode)
    nel, nI = info["nel"], info["nI"]
    assert conflicts == 0               # every edge stored the right number of times; no two edges of a round write one node
    assert len(np.unique(perm)) == nel and perm.min() >= 0 and perm.max() < info["npad"]      # injective renumbering
    assert info["npad"] == info["ntiles"] * tile_nodes and info["npad"] - nel < tile_nodes * max(1, info["ntiles"] // 8 + 1)
    if flux_mode == M.FLUX_TILED_COLOURED:
        assert info["used_slots"] == nI + info["cut_edges"]           # inside edges once, cut edges from both sides
    else:
        assert info["used_slots"] == 2 * nI                           # every edge from both of its ends
    # determinism: the plan is a pure function of the mesh
    info2, perm2, _ = M.plan_level(mesh, 0, ordering=ordering, tile_nodes=tile_nodes, flux_mode=flux_mode)
    assert info2 == info and np.array_equal(perm, perm2)


def test_partition_rcm_improves_locality_on_a_shuffled_mesh():
    shuffled = M.Mesh.generate(M.GEN_HEX_BOX, [[24, 24, 24]], ordering=1, seed=7)
    as_given, _, _ = M.plan_level(shuffled, 0, ordering=M.ORDER_AS_GIVEN, tile_nodes=256)
    rcm, _, _ = M.plan_level(shuffled, 0, ordering=M.ORDER_RCM, tile_nodes=256)
    part, _, _ = M.plan_level(shuffled, 0, ordering=M.ORDER_PARTITION_RCM, tile_nodes=256)
    assert part["cut_edges"] < rcm["cut_edges"] < as_given["cut_edges"]
    assert part["halo_entries"] < 0.5 * as_given["halo_entries"]


PLAN_MESHES = [(0, [[15, 14, 13], [8, 8, 7]], 2, 0), (1, [[11, 10, 9], [6, 5, 5]], 3, 0), (2, [[6, 5, 5]], 0, 0), (0, [[13, 12, 11], [7, 6, 6]], 4, 1)]


@pytest.mark.parametrize("kind,dims,variant,node_order", PLAN_MESHES)
@pytest.mark.parametrize("ordering", [M.ORDER_AS_GIVEN, M.ORDER_RCM, M.ORDER_PARTITION_RCM])
@pytest.mark.parametrize("tile_nodes,flux_mode", [(128, M.FLUX_SORTED_SEGMENT), (256, M.FLUX_SORTED_SEGMENT), (0, M.FLUX_SORTED_SEGMENT),
                                                  (128, M.FLUX_TILED_COLOURED), (256, M.FLUX_TILED_COLOURED)])
def test_plan_byte_streams_reproduce_the_oracle_fluxes(kind, dims, variant, node_order, ordering, tile_nodes, flux_mode):
    """No GPU: the tile headers, edge round blocks (weights, swizzled row codes, halo lists) and boundary blocks the device
    receives, walked on the host the way the stage kernel's threads walk them, give the oracle's compute_flux_edge /
    compute_boundary_flux_edge / compute_wall_flux_edge sums on every level (rounding apart: a different, fixed summation order)."""
    from conftest import linf_rel, mesh_levels, perturbed_state
    from oracle.loader import Oracle
    orc = Oracle()
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant, ordering=node_order)
    if node_order == 1 and ordering != M.ORDER_PARTITION_RCM:
        pytest.skip("a shuffled numbering is only tiled after renumbering")
    for l, L in enumerate(mesh_levels(mesh, apply_ewt_with=orc)):
        var = perturbed_state(L["nel"], seed=300 + l)
        want = np.zeros(5 * L["nel"])
        orc.flux_edge(0, L["nI"], L["edges"], var, want)
        got = M.plan_emulate_flux(L, var, mask=1, ordering=ordering, tile_nodes=tile_nodes, flux_mode=flux_mode)
        assert np.all(linf_rel(got, want) < 1e-13), ("internal", l, linf_rel(got, want))
        orc.boundary_flux_edge(L["nI"], L["nB"], L["edges"], var, want)
        orc.wall_flux_edge(L["nI"] + L["nB"], L["nW"], L["edges"], var, want)
        got = M.plan_emulate_flux(L, var, mask=7, ordering=ordering, tile_nodes=tile_nodes, flux_mode=flux_mode)
        assert np.all(linf_rel(got, want) < 1e-13), ("all", l, linf_rel(got, want))
        only_wall = M.plan_emulate_flux(L, var, mask=4, ordering=ordering, tile_nodes=tile_nodes, flux_mode=flux_mode)
        ref_wall = np.zeros(5 * L["nel"])
        orc.wall_flux_edge(L["nI"] + L["nB"], L["nW"], L["edges"], var, ref_wall)
        assert np.allclose(only_wall, ref_wall, rtol=1e-13, atol=1e-22)


@pytest.mark.parametrize("kind,dims,variant,node_order", [m for m in PLAN_MESHES if len(m[1]) > 1] + [(0, [[12, 11, 10], [9, 8, 7], [5, 5, 4]], 2, 0)])
@pytest.mark.parametrize("tile_nodes", [0, 128, 256])
def test_transfer_operators_reproduce_the_oracle_transfers(kind, dims, variant, node_order, tile_nodes):
    """No GPU: the restrict child lists and the prolong operator (own-parent weights, entries with the b-side quirk of
    mg_loops.cpp:804-810 baked in) applied on the host give mg_restrict bit for bit and prolong_residuals_interpolate_proper to rounding."""
    from conftest import linf_rel, mesh_levels, perturbed_state
    from oracle.loader import Oracle
    orc = Oracle()
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant, ordering=node_order)
    lv = mesh_levels(mesh, apply_ewt_with=orc)
    rng = np.random.default_rng(11)
    for l in range(mesh.levels - 1):
        F, C = lv[l], lv[l + 1]
        vf, vc = perturbed_state(F["nel"], seed=40 + l), perturbed_state(C["nel"], seed=50 + l)
        r1, r2 = 1e-3 * rng.standard_normal(5 * C["nel"]), 1e-3 * rng.standard_normal(5 * F["nel"])
        got_c, got_f = M.plan_emulate_transfers(F, C, vf, r2, r1, vc, tile_nodes=tile_nodes)
        want_c = vc.copy()
        orc.mg_restrict(vf, want_c, F["map"])
        assert np.array_equal(got_c, want_c)
        want_f = vf.copy()
        orc.prolong(F["edges"], F["nI"], r1, r2, want_f, F["map"], C["coords"], F["coords"])
        ok = np.isfinite(want_f)                   # nodes without an internal edge are 0/0 in the reference (and here)
        assert np.array_equal(np.isfinite(got_f), ok)
        assert np.all(linf_rel(np.where(ok, got_f, 0.0), np.where(ok, want_f, 0.0)) < 1e-14)


@pytest.mark.parametrize("kind,dims,variant,node_order", PLAN_MESHES)
@pytest.mark.parametrize("supers", [1, 3, 8, 1000])
def test_visit_plan_byte_streams_reproduce_the_oracle_fluxes(kind, dims, variant, node_order, supers):
    """No GPU: the visit kernel's view of a level -- super-tile descriptors (consecutive 128-node tiles grouped by a two-level
    bisection), their halo lists and the edge rounds re-addressed inside the super-tile (plan.h VisitPlan) -- walked on the host
    the way k_visit's threads walk them, gives the oracle's flux sums; every row belongs to exactly one super-tile."""
    from conftest import linf_rel, mesh_levels, perturbed_state
    from oracle.loader import Oracle
    orc = Oracle()
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant, ordering=node_order)
    for l, L in enumerate(mesh_levels(mesh, apply_ewt_with=orc)):
        var = perturbed_state(L["nel"], seed=500 + l)
        want = np.zeros(5 * L["nel"])
        orc.flux_edge(0, L["nI"], L["edges"], var, want)
        got, info = M.plan_emulate_visit_flux(L, var, supers, mask=1)
        assert np.all(linf_rel(got, want) < 1e-13), ("internal", l, linf_rel(got, want))
        orc.boundary_flux_edge(L["nI"], L["nB"], L["edges"], var, want)
        orc.wall_flux_edge(L["nI"] + L["nB"], L["nW"], L["edges"], var, want)
        got, info = M.plan_emulate_visit_flux(L, var, supers, mask=7)
        assert np.all(linf_rel(got, want) < 1e-13), ("all", l, linf_rel(got, want))
        assert info["supers"] == min(supers, info["tiles"]) and info["rows"] >= info["tiles"] * 128
        assert info["max_tiles"] == -(-info["tiles"] // info["supers"])
        cfg = M.plan_visit_config(L, num_sms=4)
        assert cfg["visit"] == 1 and cfg["smem_bytes"] <= 230000 and cfg["ring_rounds"] >= 1 and cfg["ctas"] * cfg["supers_per_cta"] <= info["tiles"]


def test_visit_plan_on_unstructured_levels_with_hub_nodes():
    """k-nearest-neighbour levels with 40+-neighbour hubs (tiles with many rounds, several ring chunks): visit streams vs oracle."""
    from conftest import linf_rel, perturbed_state
    from oracle.loader import Oracle
    orc = Oracle()
    for n, k, seed in ((6000, 6, 1), (2500, 5, 2)):
        L = _random_level(n, k, seed)
        orc.adjust_dampen(4, L["coords"], L["edges"])
        var = perturbed_state(L["nel"], seed=seed)
        want = np.zeros(5 * L["nel"])
        orc.flux_edge(0, L["nI"], L["edges"], var, want)
        orc.boundary_flux_edge(L["nI"], L["nB"], L["edges"], var, want)
        orc.wall_flux_edge(L["nI"] + L["nB"], L["nW"], L["edges"], var, want)
        for supers in (2, 148):
            got, info = M.plan_emulate_visit_flux(L, var, supers)
            assert np.all(linf_rel(got, want) < 1e-13), (n, supers, linf_rel(got, want))


def test_visit_configuration_of_the_baseline_workload():
    """BASELINE.json configs[1] (C2) on a 148-SM device: every level runs the visit kernel; the two coarsest keep their own rows
    resident in shared memory, the finer ones stream double-buffered super-tiles; the super-tile halo is well below the owned rows."""
    import bench
    from conftest import mesh_levels
    kind, dims, variant, _ = bench.WORKLOADS["c2"]
    mesh = M.Mesh.generate(kind, dims, mesh_variant=variant)
    cfgs = [M.plan_visit_config(L) for L in mesh_levels(mesh)]
    assert all(c["visit"] == 1 and c["ctas"] == 148 and c["smem_bytes"] <= 230000 for c in cfgs)
    assert [c["resident"] for c in cfgs] == [0, 0, 1, 1] and all(c["ring_rounds"] >= 2 for c in cfgs)
    for c, L in zip(cfgs, mesh_levels(mesh)):
        assert c["halo_rows"] < 0.75 * L["nel"]


def test_plan_does_not_depend_on_the_number_of_host_threads():
    """The preprocessing runs tile ranges on a pool of host threads; the bytes handed to the device must not depend on it."""
    import subprocess, sys, json
    code = ("import sys, json; sys.path.insert(0, %r); import mgcfd_b200 as M\n"
            "m = M.Mesh.generate(M.GEN_TET_BOX, [[45, 41, 37]])\n"
            "print(json.dumps([M.plan_level(m, 0, tile_nodes=128, flux_mode=fm)[0]['plan_hash'] for fm in (0, 1)]))\n") % ROOT
    out = []
    for threads in ("1", "3", "64"):
        env = dict(os.environ, MGCFD_PLAN_THREADS=threads)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        out.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert out[0] == out[1] == out[2] and out[0][0] != out[0][1]


def test_assess_compute_arithmetic_on_the_host(tmp_path):
    """The flux arithmetic variants of csrc/assess_kernels.cuh (the reference's FLUX_REUSE_* toggles) are __host__ __device__: their
    host instantiation reproduces the oracle's compute_flux_edge (default variant bit for bit) -- tools/assess_host_check.cu."""
    import shutil
    if shutil.which("nvcc") is None or shutil.which("gcc") is None:
        pytest.skip("needs nvcc and gcc")
    obj, exe = str(tmp_path / "orc.o"), str(tmp_path / "assess_check")
    r = subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-c", "-o", obj, os.path.join(ROOT, "oracle", "mgcfd_oracle.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "mg-cfd-app-plain_b200", "csrc"),
                        "-o", exe, os.path.join(ROOT, "tools", "assess_host_check.cu"), obj], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "PASS" in out.stdout and "default 0.00e+00" in out.stdout, out.stdout + out.stderr


def _write_irregular_mesh(directory, n=600, k=5, hubs=2, hub_degree=45, seed=3):
    """A single-level text mesh as an unstructured mesher writes it: random points, k-nearest-neighbour edges, a couple of hub
    nodes with more than 32 neighbours, boundary and wall entries; every node lists ALL its neighbours (the loader keeps the
    entries with nbr < i, io.cpp:84-137)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 3))
    _, nb = cKDTree(pts).query(pts, k=k + 1)
    adj = [set() for _ in range(n)]
    for i in range(n):
        for j in nb[i, 1:]:
            if i != int(j):
                adj[i].add(int(j)); adj[int(j)].add(i)
    for hub in rng.choice(n, hubs, replace=False):
        for j in rng.choice(n, hub_degree, replace=False):
            if int(hub) != int(j):
                adj[int(hub)].add(int(j)); adj[int(j)].add(int(hub))
    weight = {}
    lines, nedges = [], 0
    for i in range(n):
        entries = []
        for j in sorted(adj[i]):
            key = (min(i, j), max(i, j))
            if key not in weight:
                weight[key] = 1e-3 * (rng.random(3) - 0.5)
            w = weight[key] if i == key[1] else -weight[key]          # outward from node i
            entries.append((j, w))
            nedges += j < i
        if i % 7 == 0:
            entries.append((-1, 1e-3 * (rng.random(3) - 0.5))); nedges += 1
        if i % 11 == 0:
            entries.append((-2, 1e-3 * (rng.random(3) - 0.5))); nedges += 1
        lines.append(f"{float(0.5 + rng.random())!r} {len(entries)}")
        lines += [f"{j} {float(w[0])!r} {float(w[1])!r} {float(w[2])!r}" for j, w in entries]
    (directory / "irregular.dat").write_text(f"{n} {nedges}\n" + "\n".join(lines) + "\n")
    (directory / "irregular.dat.coords").write_text("\n".join(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r}" for p in pts) + "\n")
    (directory / "input.dat").write_text("size = 1\nnum_levels = 1\nmesh_name = rotor37\n[levels]\n0 = irregular.dat\n")
    return max(len(a) for a in adj)


def test_irregular_mesh_with_high_degree_nodes(tmp_path):
    """Nodes with more than 32 neighbours (unstructured vertex-centred meshes have them) load completely -- the arrays equal the
    reference loader's -- survive a write / load round trip, and the plan built from them reproduces the oracle's fluxes."""
    from conftest import linf_rel, perturbed_state
    from oracle.loader import Oracle
    max_degree = _write_irregular_mesh(tmp_path)
    assert max_degree > 40
    mesh = M.Mesh.load("input.dat", str(tmp_path))
    L = mesh_levels(mesh)[0]
    assert L["nel"] == 600 and np.bincount(np.concatenate([L["edges"]["a"][:L["nI"]], L["edges"]["b"][:L["nI"]]])).max() == max_degree
    if reference_available():
        r = Reference().read_grid(str(tmp_path / "irregular.dat"), 2, M.MESH_ROTOR_37)        # levels > 1: the reference reads .coords
        assert (r["nel"], r["nI"], r["nB"], r["nW"]) == (L["nel"], L["nI"], L["nB"], L["nW"])
        assert r["edges"].tobytes() == L["edges"].tobytes() and np.array_equal(r["vol"], L["vol"]) and np.array_equal(r["coords"], L["coords"])
    out = tmp_path / "again"
    out.mkdir()
    mesh.write(str(out))
    back = mesh_levels(M.Mesh.load("input.dat", str(out)))[0]
    assert back["edges"].tobytes() == L["edges"].tobytes() and np.array_equal(back["vol"], L["vol"])
    orc = Oracle()
    orc.adjust_dampen(M.MESH_ROTOR_37, L["coords"], L["edges"])
    var = perturbed_state(L["nel"], seed=9)
    want = np.zeros(5 * L["nel"])
    orc.flux_edge(0, L["nI"], L["edges"], var, want)
    orc.boundary_flux_edge(L["nI"], L["nB"], L["edges"], var, want)
    orc.wall_flux_edge(L["nI"] + L["nB"], L["nW"], L["edges"], var, want)
    for tile_nodes, flux_mode in ((128, M.FLUX_SORTED_SEGMENT), (256, M.FLUX_SORTED_SEGMENT), (128, M.FLUX_TILED_COLOURED)):
        got = M.plan_emulate_flux(L, var, mask=7, tile_nodes=tile_nodes, flux_mode=flux_mode)
        assert np.all(linf_rel(got, want) < 1e-13), (tile_nodes, flux_mode, linf_rel(got, want))
    # the partition closure rules hold on it too, and rank-local text loading is not involved (one process reads the file)
    for r in range(3):
        p = M.partition_plan(mesh, 3, r, 0)
        assert p["owned"] > 0 and p["ghosts"] > 0


def _random_level(n, k, seed, isolated=0):
    """Unstructured level in memory (read_grid's edge order): random points, k-nearest-neighbour edges, three hubs, boundary and wall
    edges; the last `isolated` nodes have no internal edge at all."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 3))
    m = n - isolated
    _, nb = cKDTree(pts[:m]).query(pts[:m], k=k + 1)
    pairs = {(min(i, int(j)), max(i, int(j))) for i in range(m) for j in nb[i, 1:] if i != int(j)}
    for hub in rng.choice(m, 3, replace=False):
        pairs |= {(min(int(hub), int(j)), max(int(hub), int(j))) for j in rng.choice(m, 40, replace=False) if int(hub) != int(j)}
    pairs = sorted(pairs, key=lambda p: (p[1], p[0]))
    nI, nB, nW = len(pairs), n // 20, n // 20
    e = np.zeros(nI + nB + nW, dtype=M.EDGE_DTYPE)
    e["a"][:nI] = [p[0] for p in pairs]; e["b"][:nI] = [p[1] for p in pairs]
    e["a"][nI:nI + nB] = -1; e["b"][nI:nI + nB] = np.sort(rng.choice(n, nB))
    e["a"][nI + nB:] = -2; e["b"][nI + nB:] = np.sort(rng.choice(n, nW))
    for f in ("x", "y", "z"):
        e[f] = 1e-3 * (rng.random(len(e)) - 0.5)
    return dict(nel=n, nI=nI, nB=nB, nW=nW, vol=rng.random(n) + 0.5, edges=e, coords=pts.copy(), map=None)


@pytest.mark.parametrize("nf,nc,k", [(1500, 400, 5), (2400, 1900, 8), (700, 90, 4)])
def test_transfer_operators_on_unstructured_levels(nf, nc, k):
    """Arbitrary (nearest-node) multigrid maps between unstructured levels: coarse nodes without children keep their value
    (mg_loops.cpp:43-90), fine nodes that coincide with their parent take its residual (:741-775), fine nodes without an internal
    edge come out NaN as in the reference (0/0, :844-852) -- restriction bit for bit, prolongation to rounding."""
    from scipy.spatial import cKDTree
    from conftest import linf_rel, perturbed_state
    from oracle.loader import Oracle
    orc = Oracle()
    rng = np.random.default_rng(nf)
    F, C = _random_level(nf, k, 10 + nf, isolated=4), _random_level(nc, k, 20 + nc)
    F["coords"][:40] = C["coords"][rng.choice(nc, 40)]                    # exact coincidences
    F["map"] = cKDTree(C["coords"]).query(F["coords"])[1].astype(np.int64)
    assert len(np.unique(F["map"])) < nc or nc < 200                      # some coarse nodes have no children
    vf, vc = perturbed_state(nf, seed=1), perturbed_state(nc, seed=2)
    r1, r2 = 1e-3 * rng.standard_normal(5 * nc), 1e-3 * rng.standard_normal(5 * nf)
    for tile_nodes in (128, 256):
        got_c, got_f = M.plan_emulate_transfers(F, C, vf, r2, r1, vc, tile_nodes=tile_nodes)
        want_c = vc.copy()
        orc.mg_restrict(vf, want_c, F["map"])
        assert np.array_equal(got_c, want_c)
        want_f = vf.copy()
        orc.prolong(F["edges"], F["nI"], r1, r2, want_f, F["map"], C["coords"], F["coords"])
        ok = np.isfinite(want_f)
        assert (~ok).sum() == 5 * 4 and np.array_equal(np.isfinite(got_f), ok)
        assert np.all(linf_rel(np.where(ok, got_f, 0.0), np.where(ok, want_f, 0.0)) < 1e-14)
