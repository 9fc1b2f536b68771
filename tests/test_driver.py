"""euler3d_b200, the reference's driver on top of the C ABI: same CLI, same files.  CPU part: option handling and the loud
no-device failure; GPU part: run beside the reference's own binary on the same input files."""
import os
import subprocess

import numpy as np
import pytest

import mgcfd_b200 as M
from oracle.loader import HERE as ORACLE_DIR

REF_EXE = os.path.join(ORACLE_DIR, "_ref", "euler3d_ref.b")


def run(exe, args, cwd):
    return subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd, timeout=300)


def rms_lines(stdout):
    return [float(line.split("RMS =")[1].split(")")[0]) for line in stdout.splitlines() if "RMS =" in line]


def test_driver_help_and_argument_errors(tmp_path):
    assert os.path.exists(M.DRIVER_PATH)
    r = run(M.DRIVER_PATH, ["-h"], tmp_path)
    assert r.returncode == 0 and "--mesh-duplicate-count" in r.stderr and "--output-step-factors" in r.stderr
    r = run(M.DRIVER_PATH, [], tmp_path)
    assert r.returncode != 0 and "Input file not specified" in r.stderr
    r = run(M.DRIVER_PATH, ["-i", "missing.dat", "-d", str(tmp_path)], tmp_path)
    assert r.returncode != 0 and "missing.dat" in r.stderr


def test_driver_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    M.Mesh.generate(M.GEN_HEX_BOX, [[5, 5, 5], [3, 3, 3]]).write(str(tmp_path))
    r = run(M.DRIVER_PATH, ["-i", "input.dat", "-d", str(tmp_path), "-g", "1"], tmp_path)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_multi_gpu_driver_fails_loudly_and_does_not_hang_without_gpus(tmp_path):
    """--gpus N forks one process per GPU; when the ranks cannot get their devices every one of them reports it and the parent
    returns a failure -- nobody is left waiting at a barrier."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    M.Mesh.generate(M.GEN_HEX_BOX, [[5, 5, 5], [3, 3, 3]]).write(str(tmp_path))
    r = subprocess.run([M.DRIVER_PATH, "-i", "input.dat", "-d", str(tmp_path), "-g", "1", "--gpus", "3"], capture_output=True, text=True, cwd=tmp_path, timeout=60)
    assert r.returncode != 0 and "a rank process failed" in r.stderr
    assert all(f"ERROR (rank {k})" in r.stderr for k in range(3)) and "no CPU fallback" in r.stderr
    r = run(M.DRIVER_PATH, ["-i", "input.dat", "-d", str(tmp_path), "--gpus", "0"], tmp_path)
    assert r.returncode != 0 and "--gpus must be in 1..64" in r.stderr


def test_config_file_keys_and_relative_input_directory(tmp_path):
    """The key = value config file (config.cpp:81-217): a relative input_file_directory is relative to the config file
    (config.cpp:192-216); unknown keys warn; the reference's full key set is accepted."""
    import torch
    (tmp_path / "meshes").mkdir()
    M.Mesh.generate(M.GEN_HEX_BOX, [[5, 5, 5], [3, 3, 3]]).write(str(tmp_path / "meshes"))
    (tmp_path / "run.conf").write_text("# comment\ninput_file = input.dat\ninput_file_directory = meshes\ncycles = 1\noutput_volumes = Y\n"
                                       "output_old_variables = N\noutput_edge_fluxes = N\nomp_num_threads = 4\nnot_a_key = 1\n")
    elsewhere = tmp_path / "elsewhere"
    elsewhere.mkdir()
    r = run(M.DRIVER_PATH, ["-c", str(tmp_path / "run.conf"), "-o", str(elsewhere) + "/"], elsewhere)
    assert "Unknown key 'not_a_key'" in r.stdout and "output_old_variables" not in r.stdout
    if torch.cuda.is_available():
        assert r.returncode == 0 and os.path.exists(elsewhere / "volumes.size=1x.cycles=1.level=0"), r.stdout + r.stderr
    else:       # the mesh was found and read (through the config-relative directory); what fails is the missing device
        assert r.returncode != 0 and "no CPU fallback" in r.stderr and "input.dat" not in r.stderr


def test_mesh_duplicate_layout():
    """-m: nodes copy-major, each edge class copy-major in its own range, MG map shifted (io_enhanced.cpp:89-201)."""
    from conftest import mesh_levels
    base = M.Mesh.generate(M.GEN_HEX_BOX, [[5, 4, 4], [3, 2, 2]])
    one = mesh_levels(base)
    dup = M.Mesh.generate(M.GEN_HEX_BOX, [[5, 4, 4], [3, 2, 2]])
    dup.duplicate(3)
    three = mesh_levels(dup)
    for l, (a, b) in enumerate(zip(one, three)):
        n, nI, nB, nW = a["nel"], a["nI"], a["nB"], a["nW"]
        assert (b["nel"], b["nI"], b["nB"], b["nW"]) == (3 * n, 3 * nI, 3 * nB, 3 * nW)
        assert np.array_equal(b["vol"], np.tile(a["vol"], 3)) and np.array_equal(b["coords"], np.tile(a["coords"], (3, 1)))
        for c in range(3):
            ei = b["edges"][c * nI:(c + 1) * nI]
            assert np.array_equal(ei["a"], a["edges"]["a"][:nI] + c * n) and np.array_equal(ei["b"], a["edges"]["b"][:nI] + c * n)
            eb = b["edges"][3 * nI + c * nB:3 * nI + (c + 1) * nB]
            assert np.all(eb["a"] == -1) and np.array_equal(eb["b"], a["edges"]["b"][nI:nI + nB] + c * n)
            ew = b["edges"][3 * nI + 3 * nB + c * nW:3 * nI + 3 * nB + (c + 1) * nW]
            assert np.all(ew["a"] == -2) and np.array_equal(ew["x"], a["edges"]["x"][nI + nB:])
        if a["map"] is not None:
            nc = one[l + 1]["nel"]
            assert np.array_equal(b["map"], np.concatenate([a["map"] + c * nc for c in range(3)]))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="reference binary not built")
@pytest.mark.parametrize("kind,dims,variant,dup", [(0, [[13, 12, 11], [7, 6, 6], [4, 3, 3]], 2, 1), (1, [[9, 8, 7], [5, 4, 4]], 3, 2), (2, [[6, 5, 4]], 0, 1)])
def test_driver_beside_the_reference_binary(tmp_path, kind, dims, variant, dup):
    M.Mesh.generate(kind, dims, mesh_variant=variant).write(str(tmp_path))
    d = str(tmp_path) + "/"
    os.makedirs(tmp_path / "ref"); os.makedirs(tmp_path / "gpu")
    common = ["-i", "input.dat", "-d", d, "-g", "6", "-m", str(dup), "--output-variables", "--output-step-factors"]
    a = run(REF_EXE, common + ["-o", d + "ref/"], tmp_path)
    b = run(M.DRIVER_PATH, common + ["-o", d + "gpu/"], tmp_path)
    assert a.returncode == 0, a.stdout + a.stderr
    assert b.returncode == 0, b.stdout + b.stderr
    ra, rb = rms_lines(a.stdout), rms_lines(b.stdout)
    assert len(ra) == len(rb) == 6 and np.allclose(ra, rb, rtol=2e-3)           # both printed with %.3e
    assert "Total runtime = " in b.stdout
    name = f"variables.size={dup}x.cycles=6.level=0"
    va, vb = np.loadtxt(tmp_path / "ref" / name), np.loadtxt(tmp_path / "gpu" / name)
    assert va.shape == vb.shape
    assert np.all(np.max(np.abs(va - vb), axis=0) <= 1e-11 * np.max(np.abs(va), axis=0))
    sa = np.loadtxt(tmp_path / "ref" / f"step_factors.size={dup}x.cycles=6.level=0")
    sb = np.loadtxt(tmp_path / "gpu" / f"step_factors.size={dup}x.cycles=6.level=0")
    assert np.max(np.abs(sa - sb) / sa) < 1e-11
    # CSV schema: same columns as the reference's files
    for csv in ("Times.csv", "LoopNumIters.csv"):
        ha = open(tmp_path / "ref" / csv).readline().strip()
        hb = open(tmp_path / "gpu" / csv).readline().strip()
        assert ha == hb, csv
    it_a = open(tmp_path / "ref" / "LoopNumIters.csv").read().splitlines()[1].split(",")
    it_b = open(tmp_path / "gpu" / "LoopNumIters.csv").read().splitlines()[1].split(",")
    cols = open(tmp_path / "ref" / "LoopNumIters.csv").readline().strip().split(",")
    for k, c in enumerate(cols):
        if c.startswith(("flux", "compute_step", "time_step")):               # work counts must agree exactly
            assert int(float(it_a[k])) == int(float(it_b[k])), c
    # -v: the reference's dump as the golden solution file
    os.rename(tmp_path / "ref" / name, tmp_path / ("solution." + name))
    v = run(M.DRIVER_PATH, ["-i", "input.dat", "-d", d, "-g", "6", "-m", str(dup), "-v", "-o", d + "gpu/"], tmp_path)
    assert v.returncode == 0 and "PASS" in v.stdout, v.stdout + v.stderr
    w = run(M.DRIVER_PATH, ["-i", "input.dat", "-d", d, "-g", "5", "-m", str(dup), "-v", "-o", d + "gpu/"], tmp_path)
    assert w.returncode != 0                                                    # no solution file for 5 cycles



@pytest.mark.gpu
@pytest.mark.parametrize("kind,dims,variant", [(0, [[21, 19, 17], [11, 10, 9], [6, 5, 5]], 2), (2, [[7, 6, 5]], 0)])
def test_multi_gpu_driver_matches_the_single_gpu_driver(tmp_path, kind, dims, variant):
    """euler3d_b200 --gpus 2 (one forked process per GPU, mesh split by mgcfd_dist.h) writes the same files as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    M.Mesh.generate(kind, dims, mesh_variant=variant).write(str(tmp_path))
    d = str(tmp_path) + "/"
    os.makedirs(tmp_path / "one"); os.makedirs(tmp_path / "two")
    common = ["-i", "input.dat", "-d", d, "-g", "5", "--output-variables", "--output-step-factors"]
    a = run(M.DRIVER_PATH, common + ["-o", d + "one/"], tmp_path)
    b = run(M.DRIVER_PATH, common + ["-o", d + "two/", "--gpus", "2"], tmp_path)
    assert a.returncode == 0, a.stdout + a.stderr
    assert b.returncode == 0, b.stdout + b.stderr
    assert np.allclose(rms_lines(a.stdout), rms_lines(b.stdout), rtol=2e-3) and len(rms_lines(b.stdout)) == 5
    for name in ("variables.size=1x.cycles=5.level=0", "step_factors.size=1x.cycles=5.level=0"):
        va, vb = np.loadtxt(tmp_path / "one" / name), np.loadtxt(tmp_path / "two" / name)
        assert va.shape == vb.shape
        assert np.all(np.max(np.abs(va - vb), axis=0) <= 1e-11 * np.max(np.abs(va), axis=0)), name
    rows = open(tmp_path / "two" / "Times.csv").read().splitlines()
    assert len(rows) == 3 and rows[0] == open(tmp_path / "one" / "Times.csv").readline().strip()     # header + one row per GPU
    # the flux work of the two GPUs covers every internal edge (cut edges are evaluated on both sides)
    cols = rows[0].split(",")
    k = cols.index("flux0")
    it = [open(tmp_path / w / "LoopNumIters.csv").read().splitlines() for w in ("one", "two")]
    one = int(float(it[0][1].split(",")[k]))
    two = sum(int(float(r.split(",")[k])) for r in it[1][1:])
    assert one <= two <= 1.5 * one
