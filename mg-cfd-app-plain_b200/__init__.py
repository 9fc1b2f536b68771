"""mgcfd_b200 -- host-side Python mirror of the reference's kernel interface for the MG-CFD solver loop.

Thin ctypes bindings over ``libmgcfd_b200.so`` (C ABI in ``include/mgcfd_b200.h`` / ``include/mgcfd_mesh.h``).
Method names follow the reference's free functions (``src/Kernels/*.h``) so parity tests read like the
reference's own call sequence in ``main()`` (``src/euler3d_cpu_double.cpp:371-694``).

There is no CPU fallback here: if the CUDA library is missing or no device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgcfd_b200.so")
DRIVER_PATH = os.path.join(_HERE, "euler3d_b200")     # the reference's euler3d driver on top of the C ABI

NVAR = 5
RK = 3
MESH_FVCORR, MESH_M6_WING, MESH_LA_CASCADE, MESH_ROTOR_37 = 0, 2, 3, 4
FIELD_VARIABLES, FIELD_OLD_VARIABLES, FIELD_RESIDUALS, FIELD_FLUXES, FIELD_STEP_FACTORS, FIELD_VOLUMES = range(6)
FLUX_TILED_COLOURED, FLUX_SORTED_SEGMENT, FLUX_ATOMIC = 0, 1, 2
ORDER_AS_GIVEN, ORDER_RCM, ORDER_PARTITION_RCM = 0, 1, 2
GEN_HEX_BOX, GEN_TET_BOX, GEN_TET_CELLS = 0, 1, 2
KERNEL_NAMES = ("compute_step", "flux", "update", "indirect_rw", "time_step", "restrict", "prolong")

# layout of the reference's `edge_neighbour` (src/Base/definitions.h:83)
EDGE_DTYPE = np.dtype([("a", np.int64), ("b", np.int64), ("x", np.float64), ("y", np.float64), ("z", np.float64)])
assert EDGE_DTYPE.itemsize == 40


class MgcfdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mgcfd error {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("flux_mode", C.c_int), ("ordering", C.c_int), ("tile_nodes", C.c_int),
                ("use_graph", C.c_int), ("timing", C.c_int), ("no_pipeline", C.c_int), ("no_pdl", C.c_int), ("visit", C.c_int), ("reserved", C.c_int * 7)]


_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA library; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). mgcfd_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, l, dp = C.c_void_p, C.c_int, C.c_long, C.POINTER(C.c_double)
    L.mgcfd_last_error.restype = C.c_char_p
    L.mgcfd_guard_check.restype = C.c_int
    L.mgcfd_guard_check.argtypes = [C.c_char_p, C.c_int]
    L.mgcfd_guard_selftest.restype = C.c_int
    L.mgcfd_guard_selftest.argtypes = [C.c_void_p]
    L.mgcfd_mesh_last_error.restype = C.c_char_p
    L.mgcfd_version.restype = C.c_char_p
    L.mgcfd_default_options.argtypes = [C.POINTER(Options)]
    L.mgcfd_create.argtypes = [i, i, C.POINTER(Options), C.POINTER(vp)]
    L.mgcfd_destroy.argtypes = [vp]
    L.mgcfd_set_farfield.argtypes = [vp, dp, dp]
    L.mgcfd_far_field_conditions.argtypes = [dp, dp]
    L.mgcfd_far_field_conditions.restype = None
    L.mgcfd_upload_level.argtypes = [vp, i, l, vp, vp, l, l, l, vp, vp, l]
    L.mgcfd_finalize.argtypes = [vp]
    L.mgcfd_adjust_dampen_ewt.argtypes = [i, vp, l, vp]
    for name in ("initialize_variables", "copy_old_variables", "compute_flux_edge", "compute_boundary_flux_edge",
                 "compute_wall_flux_edge", "zero_fluxes", "indirect_rw", "residual", "mg_restrict", "prolong"):
        getattr(L, "mgcfd_" + name).argtypes = [vp, i]
    L.mgcfd_flux_variant.argtypes = [vp, i, i]
    L.mgcfd_compute_step_factor.argtypes = [vp, i, i]
    L.mgcfd_time_step.argtypes = [vp, i, i]
    L.mgcfd_calc_rms.argtypes = [vp, i, dp, dp]
    L.mgcfd_check_for_invalid_variables.argtypes = [vp, i, C.POINTER(l), C.POINTER(i)]
    L.mgcfd_run_cycles.argtypes = [vp, i, vp, vp]
    L.mgcfd_enqueue_cycles.argtypes = [vp, i]
    L.mgcfd_collect.argtypes = [vp, vp, vp]
    L.mgcfd_invalid_cell.argtypes = [vp, C.POINTER(l), C.POINTER(i)]
    L.mgcfd_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.mgcfd_set_timing.argtypes = [vp, i]
    L.mgcfd_get_field.argtypes = [vp, i, i, vp]
    L.mgcfd_set_field.argtypes = [vp, i, i, vp]
    L.mgcfd_synchronize.argtypes = [vp]
    L.mgcfd_level_info.argtypes = [vp, i, C.POINTER(l)]
    L.mgcfd_visit_info.argtypes = [vp, i, C.POINTER(l)]
    L.mgcfd_visit_debug.argtypes = [vp, vp, l]
    L.mgcfd_get_permutation.argtypes = [vp, i, vp]
    L.mgcfd_check_colouring.argtypes = [vp, i]
    L.mgcfd_check_colouring.restype = l
    L.mgcfd_get_times.argtypes = [vp, vp, vp]
    L.mgcfd_reset_times.argtypes = [vp]
    L.mgcfd_launch_count.argtypes = [vp]
    L.mgcfd_launch_count.restype = l
    L.mgcfd_time_kernel.argtypes = [vp, i, i, i, dp]
    L.mgcfd_plan_level.argtypes = [l, vp, l, l, l, vp, i, i, i, C.POINTER(l), vp, C.POINTER(l)]
    L.mgcfd_plan_emulate_flux.argtypes = [l, vp, l, l, l, vp, i, i, i, vp, i, vp]
    L.mgcfd_plan_emulate_visit_flux.argtypes = [l, vp, l, l, l, vp, i, vp, i, vp, C.POINTER(l)]
    L.mgcfd_plan_visit_config.argtypes = [l, vp, l, l, l, vp, i, C.POINTER(l)]
    L.mgcfd_plan_emulate_transfers.argtypes = [l, vp, l, l, l, vp, vp, l, vp, l, l, l, vp, i, i, vp, vp, vp, vp, vp]
    L.mgcfd_mesh_generate.argtypes = [i, i, vp, dp, i, i, C.c_ulong, C.c_double, C.POINTER(vp)]
    L.mgcfd_mesh_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.mgcfd_mesh_write.argtypes = [vp, C.c_char_p, C.c_char_p, i]
    L.mgcfd_mesh_levels.argtypes = [vp]
    L.mgcfd_mesh_variant.argtypes = [vp]
    L.mgcfd_mesh_dims.argtypes = [vp, i, C.POINTER(l)]
    L.mgcfd_mesh_ptr.argtypes = [vp, i, i]
    L.mgcfd_mesh_ptr.restype = vp
    L.mgcfd_mesh_apply_ewt.argtypes = [vp]
    L.mgcfd_mesh_upload.argtypes = [vp, vp]
    L.mgcfd_mesh_free.argtypes = [vp]
    L.mgcfd_mesh_upload_partition.argtypes = [vp, vp]
    L.mgcfd_mesh_duplicate.argtypes = [vp, i]
    L.mgcfd_mesh_partition_plan.argtypes = [vp, i, i, i, C.POINTER(l), vp, vp, vp, vp]
    L.mgcfd_mesh_delivery_check.argtypes = [vp, i, i, C.POINTER(l)]
    L.mgcfd_dist_get_unique_id.argtypes = [C.c_char_p]
    L.mgcfd_dist_init.argtypes = [vp, i, i, C.c_char_p]
    L.mgcfd_dist_level_info.argtypes = [vp, i, C.POINTER(l)]
    L.mgcfd_dist_global_ids.argtypes = [vp, i, vp]
    L.mgcfd_dist_p2p_table_len.argtypes = [vp]
    L.mgcfd_dist_p2p_table_len.restype = l
    L.mgcfd_dist_p2p_prepare.argtypes = [vp, C.c_char_p, vp, l]
    L.mgcfd_dist_p2p_attach.argtypes = [vp, C.c_char_p, vp, l]
    L.mgcfd_generate_upload_partition.argtypes = [vp, i, i, vp, dp, i, C.c_double]
    L.mgcfd_generate_partition_plan.argtypes = [i, i, vp, dp, i, C.c_double, i, i, i, i, C.POINTER(l), vp, vp, vp, vp]
    L.mgcfd_mesh_free.restype = None
    _lib = L
    return L


def _visit_dict(info):
    d = dict(zip(("visit", "supers_per_cta", "ctas", "ring_rounds", "resident", "super_rows", "smem_bytes", "halo_rows"), info))
    d["warps"] = d["ring_rounds"] >> 16                # warps per CTA (16: one CTA per SM, 8: two)
    d["ring_entries"] = (d["ring_rounds"] >> 8) & 0xFF  # per-warp ring: entries x rounds per entry
    d["ring_rounds"] &= 0xFF
    return d


def _check(rc: int, mesh: bool = False):
    if rc != 0:
        L = lib()
        msg = (L.mgcfd_mesh_last_error() if mesh else L.mgcfd_last_error()).decode()
        raise MgcfdError(rc, msg)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def far_field_conditions():
    """initialize_far_field_conditions (src/Kernels/cfd_loops.h:85-119) -> (ff_variable[5], ff_flux_contribution[12])."""
    ffv, ffc = np.zeros(5), np.zeros(12)
    lib().mgcfd_far_field_conditions(ffv.ctypes.data_as(C.POINTER(C.c_double)), ffc.ctypes.data_as(C.POINTER(C.c_double)))
    return ffv, ffc


class Mesh:
    """A multigrid mesh as the reference holds it in memory after read_grid / read_mg_connectivity."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def generate(kind: int, dims: Sequence[Sequence[int]], mesh_variant: int = MESH_M6_WING, lengths=(1.0, 1.0, 1.0),
                 ordering: int = 0, seed: int = 12345, tilt: float = 0.05) -> "Mesh":
        d = np.ascontiguousarray(np.asarray(dims, dtype=np.int64).reshape(-1, 3))
        ln = np.asarray(lengths, dtype=np.float64)
        h = C.c_void_p()
        _check(lib().mgcfd_mesh_generate(kind, d.shape[0], _ptr(d), ln.ctypes.data_as(C.POINTER(C.c_double)), mesh_variant,
                                         ordering, seed, tilt, C.byref(h)), mesh=True)
        return Mesh(h)

    @staticmethod
    def load(input_dat: str, directory: str = "") -> "Mesh":
        h = C.c_void_p()
        _check(lib().mgcfd_mesh_load(input_dat.encode(), directory.encode(), C.byref(h)), mesh=True)
        return Mesh(h)

    def write(self, directory: str, input_dat: str = "input.dat", binary: bool = False):
        _check(lib().mgcfd_mesh_write(self._h, directory.encode(), input_dat.encode(), int(binary)), mesh=True)

    @property
    def levels(self) -> int:
        return lib().mgcfd_mesh_levels(self._h)

    @property
    def mesh_variant(self) -> int:
        return lib().mgcfd_mesh_variant(self._h)

    def dims(self, level: int):
        out = (C.c_long * 5)()
        _check(lib().mgcfd_mesh_dims(self._h, level, out), mesh=True)
        return tuple(out)  # nel, nI, nB, nW, mgc

    def _view(self, level, what, dtype, count):
        p = lib().mgcfd_mesh_ptr(self._h, level, what)
        if not p or count == 0:
            return None
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def volumes(self, level):
        return self._view(level, 0, np.float64, self.dims(level)[0])

    def edges(self, level):
        nel, nI, nB, nW, _ = self.dims(level)
        return self._view(level, 1, EDGE_DTYPE, nI + nB + nW)

    def coords(self, level):
        v = self._view(level, 2, np.float64, 3 * self.dims(level)[0])
        return None if v is None else v.reshape(-1, 3)

    def mg_map(self, level):
        return self._view(level, 3, np.int64, self.dims(level)[4])

    def duplicate(self, count: int):
        """-m / --mesh-duplicate-count: `count` independent copies, laid out as duplicate_mesh (io_enhanced.cpp:89-201)."""
        _check(lib().mgcfd_mesh_duplicate(self._h, count), mesh=True)

    def apply_ewt(self):
        """adjust_ewt + dampen_ewt (src/Kernels/validation.cpp:28-75) as main() applies them per mesh variant."""
        _check(lib().mgcfd_mesh_apply_ewt(self._h), mesh=True)

    def close(self):
        if self._h:
            lib().mgcfd_mesh_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Solver:
    """Device-resident multigrid state + the reference's kernel functions, one method per function."""

    def __init__(self, levels: int, mesh_variant: int, device: int = 0, flux_mode: int = FLUX_SORTED_SEGMENT,
                 ordering: int = ORDER_PARTITION_RCM, tile_nodes: int = 0, use_graph: bool = True, timing: bool = False,
                 pipeline: bool = True, pdl: bool = True, visit: bool = False):
        L = lib()
        opt = Options()
        L.mgcfd_default_options(C.byref(opt))
        opt.device, opt.flux_mode, opt.ordering, opt.tile_nodes = device, flux_mode, ordering, tile_nodes
        opt.use_graph, opt.timing = int(use_graph), int(timing)
        opt.no_pipeline = int(not pipeline)
        opt.no_pdl = int(not pdl)
        opt.visit = int(visit)
        self._h = C.c_void_p()
        self.levels, self.mesh_variant = levels, mesh_variant
        _check(L.mgcfd_create(levels, mesh_variant, C.byref(opt), C.byref(self._h)))
        self._nel = {}

    @classmethod
    def from_mesh(cls, mesh: Mesh, **kw) -> "Solver":
        s = cls(mesh.levels, mesh.mesh_variant, **kw)
        _check(lib().mgcfd_mesh_upload(mesh._h, s._h), mesh=True)
        for l in range(mesh.levels):
            s._nel[l] = mesh.dims(l)[0]
        return s

    @classmethod
    def from_mesh_distributed(cls, mesh: Mesh, rank: int, nranks: int, unique_id: bytes, **kw) -> "Solver":
        """One rank of a multi-GPU run (include/mgcfd_dist.h): joins the NCCL communicator identified by `unique_id`
        (from `dist_unique_id()` on rank 0, broadcast by the caller), keeps this rank's part of `mesh` and uploads it."""
        s = cls(mesh.levels, mesh.mesh_variant, **kw)
        _check(lib().mgcfd_dist_init(s._h, rank, nranks, unique_id))
        _check(lib().mgcfd_mesh_upload_partition(mesh._h, s._h), mesh=True)
        for l in range(mesh.levels):
            info = s.dist_level_info(l)
            s._nel[l] = info["owned"] + info["ghosts"]
        return s

    @classmethod
    def generate_distributed(cls, kind: int, dims, rank: int, nranks: int, unique_id: bytes, mesh_variant: int = MESH_M6_WING,
                             lengths=(1.0, 1.0, 1.0), tilt: float = 0.05, **kw) -> "Solver":
        """One rank of a multi-GPU run on a synthetic mesh that no rank ever assembles as a whole: the rank generates its own part
        (mgcfd_generate_upload_partition, include/mgcfd_dist.h) -- the same partition, bit for bit, as
        `from_mesh_distributed(Mesh.generate(...))`."""
        d = np.ascontiguousarray(np.asarray(dims, dtype=np.int64).reshape(-1, 3))
        ln = np.asarray(lengths, dtype=np.float64)
        s = cls(d.shape[0], mesh_variant, **kw)
        _check(lib().mgcfd_dist_init(s._h, rank, nranks, unique_id))
        _check(lib().mgcfd_generate_upload_partition(s._h, kind, d.shape[0], _ptr(d), ln.ctypes.data_as(C.POINTER(C.c_double)),
                                                     mesh_variant, tilt))
        for l in range(d.shape[0]):
            info = s.dist_level_info(l)
            s._nel[l] = info["owned"] + info["ghosts"]
        return s

    def p2p_prepare(self):
        """(64-byte CUDA IPC handle of this rank's window, offset table) for the direct peer-to-peer data path (mgcfd_dist.h)."""
        n = lib().mgcfd_dist_p2p_table_len(self._h)
        handle, table = C.create_string_buffer(64), np.zeros(n, dtype=np.int64)
        _check(lib().mgcfd_dist_p2p_prepare(self._h, handle, _ptr(table), n))
        return handle.raw, table

    def p2p_attach(self, handles, tables):
        """handles / tables: what every rank's p2p_prepare returned, in rank order (all-gathered by the caller)."""
        hb = b"".join(handles)
        tb = np.ascontiguousarray(np.stack(tables), dtype=np.int64)
        _check(lib().mgcfd_dist_p2p_attach(self._h, hb, _ptr(tb), tb.shape[1]))

    def dist_level_info(self, level):
        out = (C.c_long * 8)()
        _check(lib().mgcfd_dist_level_info(self._h, level, out))
        return dict(zip(("owned", "ghosts", "sent", "global_nodes", "rank", "nranks", "exchanges", "global_internal_edges"), out))

    def global_ids(self, level):
        g = np.empty(self._nel[level], dtype=np.int64)
        _check(lib().mgcfd_dist_global_ids(self._h, level, _ptr(g)))
        return g

    def upload_level(self, level, volumes, coords, nI, nB, nW, edges, mg_map=None):
        volumes = np.ascontiguousarray(volumes, dtype=np.float64)
        coords = None if coords is None else np.ascontiguousarray(coords, dtype=np.float64)
        edges = np.ascontiguousarray(edges)
        assert edges.dtype.itemsize == 40
        mg = None if mg_map is None else np.ascontiguousarray(mg_map, dtype=np.int64)
        self._nel[level] = volumes.shape[0]
        _check(lib().mgcfd_upload_level(self._h, level, volumes.shape[0], _ptr(volumes), _ptr(coords), nI, nB, nW, _ptr(edges),
                                        _ptr(mg), 0 if mg is None else mg.shape[0]))

    def finalize(self):
        _check(lib().mgcfd_finalize(self._h))

    # ---- one method per reference function ----
    def initialize_variables(self, level): _check(lib().mgcfd_initialize_variables(self._h, level))
    def copy_old_variables(self, level): _check(lib().mgcfd_copy_old_variables(self._h, level))
    def compute_step_factor(self, level, legacy=False): _check(lib().mgcfd_compute_step_factor(self._h, level, int(legacy)))
    def compute_flux_edge(self, level): _check(lib().mgcfd_compute_flux_edge(self._h, level))
    def compute_boundary_flux_edge(self, level): _check(lib().mgcfd_compute_boundary_flux_edge(self._h, level))
    def compute_wall_flux_edge(self, level): _check(lib().mgcfd_compute_wall_flux_edge(self._h, level))
    def time_step(self, level, j): _check(lib().mgcfd_time_step(self._h, level, j))
    def zero_fluxes(self, level): _check(lib().mgcfd_zero_fluxes(self._h, level))
    def indirect_rw(self, level): _check(lib().mgcfd_indirect_rw(self._h, level))

    def flux_variant(self, level, bits):
        """compute_flux_edge in the arithmetic form the reference's FLUX_* toggles select (assess-compute; include/mgcfd_b200.h)."""
        _check(lib().mgcfd_flux_variant(self._h, level, bits))
    def residual(self, level): _check(lib().mgcfd_residual(self._h, level))
    def mg_restrict(self, coarse_level): _check(lib().mgcfd_mg_restrict(self._h, coarse_level))
    def prolong(self, fine_level): _check(lib().mgcfd_prolong(self._h, fine_level))

    def calc_rms(self, level):
        a, v = C.c_double(), np.zeros(5)
        _check(lib().mgcfd_calc_rms(self._h, level, C.byref(a), v.ctypes.data_as(C.POINTER(C.c_double))))
        return a.value, v

    def check_for_invalid_variables(self, level):
        """Returns None when the state is valid, else (cell, reason) like validation.cpp:107-138 would report."""
        cell, reason = C.c_long(-1), C.c_int(0)
        rc = lib().mgcfd_check_for_invalid_variables(self._h, level, C.byref(cell), C.byref(reason))
        if rc == 3:
            return cell.value, reason.value
        _check(rc)
        return None

    def run_cycles(self, ncycles: int):
        """The fused device loop; returns (rms_all[ncycles], rms_var[ncycles, 5])."""
        ra, rv = np.zeros(ncycles), np.zeros((ncycles, 5))
        _check(lib().mgcfd_run_cycles(self._h, ncycles, _ptr(ra), _ptr(rv)))
        return ra, rv

    def enqueue_cycles(self, ncycles: int):
        """Queues V-cycles on the solver's stream without synchronising the host (at most 4096 outstanding)."""
        _check(lib().mgcfd_enqueue_cycles(self._h, ncycles))
        self._pending = getattr(self, "_pending", 0) + ncycles

    def collect(self):
        """Synchronises and returns (rms_all, rms_var) of the cycles enqueued since the last collect."""
        n = getattr(self, "_pending", 0)
        ra, rv = np.zeros(n), np.zeros((n, 5))
        self._pending = 0
        _check(lib().mgcfd_collect(self._h, _ptr(ra), _ptr(rv)))
        return ra, rv

    def invalid_cell(self):
        cell, reason = C.c_long(-1), C.c_int(0)
        _check(lib().mgcfd_invalid_cell(self._h, C.byref(cell), C.byref(reason)))
        return cell.value, reason.value

    def cuda_stream(self) -> int:
        """The cudaStream_t (as an integer) all kernels of this solver are launched on."""
        p = C.c_void_p()
        _check(lib().mgcfd_get_stream(self._h, C.byref(p)))
        return p.value or 0

    def set_farfield(self, ff_variable, ff_flux_contribution):
        """The far-field state of THIS solver (mgcfd_set_farfield; the reference's globals, globals.h:10-14): 5 + 12 doubles."""
        v = np.ascontiguousarray(ff_variable, dtype=np.float64); c = np.ascontiguousarray(ff_flux_contribution, dtype=np.float64)
        assert v.size == 5 and c.size == 12
        dp = C.POINTER(C.c_double)
        _check(lib().mgcfd_set_farfield(self._h, v.ctypes.data_as(dp), c.ctypes.data_as(dp)))

    def set_timing(self, on: bool): _check(lib().mgcfd_set_timing(self._h, int(on)))

    def get_field(self, level, field, out: Optional[np.ndarray] = None):
        n = self._nel[level]
        ncomp = 1 if field in (FIELD_STEP_FACTORS, FIELD_VOLUMES) else NVAR
        if out is None:
            out = np.empty(n * ncomp)
        _check(lib().mgcfd_get_field(self._h, level, field, _ptr(out)))
        return out.reshape(n, ncomp) if ncomp > 1 else out

    def set_field(self, level, field, values):
        values = np.ascontiguousarray(values, dtype=np.float64)
        _check(lib().mgcfd_set_field(self._h, level, field, _ptr(values)))

    def synchronize(self): _check(lib().mgcfd_synchronize(self._h))

    def level_info(self, level):
        out = (C.c_long * 16)()
        _check(lib().mgcfd_level_info(self._h, level, out))
        keys = ("nel", "nI", "nB", "nW", "npad", "ntiles", "tile_nodes", "max_rounds", "slots", "halo_entries", "cut_edges",
                "used_slots", "max_halo", "bslots", "smem_bytes", "pipe_grid")
        return dict(zip(keys, out))

    def visit_info(self, level):
        """Configuration of the persistent visit kernel on `level` (include/mgcfd_b200.h mgcfd_visit_info)."""
        out = (C.c_long * 8)()
        _check(lib().mgcfd_visit_info(self._h, level, out))
        return _visit_dict(out)

    def visit_debug(self):
        """Clock stamps of the most recent visit-kernel launch, [ctas, 64] (MGCFD_VISIT_DEBUG=1 at construction)."""
        out = np.zeros(64 * 512, dtype=np.int64)
        n = lib().mgcfd_visit_debug(self._h, _ptr(out), out.size)
        if n < 16:           # an error code, not a CTA count
            _check(n or 2)
        return out[:64 * n].reshape(n, 64)

    def permutation(self, level):
        p = np.empty(self._nel[level], dtype=np.int64)
        _check(lib().mgcfd_get_permutation(self._h, level, _ptr(p)))
        return p

    def check_colouring(self, level) -> int:
        return lib().mgcfd_check_colouring(self._h, level)

    def times(self):
        ms = np.zeros((len(KERNEL_NAMES), self.levels))
        it = np.zeros((len(KERNEL_NAMES), self.levels), dtype=np.int64)
        _check(lib().mgcfd_get_times(self._h, _ptr(ms), _ptr(it)))
        return ms, it

    def reset_times(self): _check(lib().mgcfd_reset_times(self._h))

    def launch_count(self) -> int:
        return lib().mgcfd_launch_count(self._h)

    def time_kernel(self, level, which, reps) -> float:
        ms = C.c_double()
        _check(lib().mgcfd_time_kernel(self._h, level, which, reps, C.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "_h", None):
            lib().mgcfd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def guard_check():
    """(number of damaged guard zones, description) over all live solvers; needs MGCFD_GUARD=1 in the environment when the solvers
    were created (mgcfd_guard_check, include/mgcfd_b200.h) -- the library's stand-in for a memcheck run."""
    buf = C.create_string_buffer(4096)
    n = lib().mgcfd_guard_check(buf, 4096)
    return n, buf.value.decode()


def dist_unique_id() -> bytes:
    """128-byte NCCL unique id (call on rank 0, broadcast to the others)."""
    buf = C.create_string_buffer(128)
    _check(lib().mgcfd_dist_get_unique_id(buf))
    return buf.raw


def partition_plan(mesh: Mesh, nranks: int, rank: int, level: int):
    """Host-only: what `rank` of `nranks` holds of `level` (owned/ghost global ids, per-peer send/receive counts, send order)."""
    L = lib()
    info = (C.c_long * 8)()
    _check(L.mgcfd_mesh_partition_plan(mesh._h, nranks, rank, level, info, None, None, None, None), mesh=True)
    owned, ghosts, sent = info[0], info[1], info[2]
    gid = np.empty(owned + ghosts, dtype=np.int64)
    sc, rc = np.zeros(nranks, dtype=np.int64), np.zeros(nranks, dtype=np.int64)
    sg = np.empty(max(sent, 1), dtype=np.int64)
    _check(L.mgcfd_mesh_partition_plan(mesh._h, nranks, rank, level, info, _ptr(gid), _ptr(sc), _ptr(rc), _ptr(sg)), mesh=True)
    return dict(owned=owned, ghosts=ghosts, sent=sent, global_nodes=info[3], nI=info[4], nB=info[5], nW=info[6], hash=info[7], gid=gid,
                send_counts=sc, recv_counts=rc, send_gids=sg[:sent])


def delivery_check(mesh: Mesh, nranks: int, tile_nodes: int = 0):
    """Host-only: the tables of the in-kernel halo exchange for all `nranks` ranks at once, replayed on global node ids
    (mgcfd_mesh_delivery_check, include/mgcfd_mesh.h)."""
    out = (C.c_long * 5)()
    _check(lib().mgcfd_mesh_delivery_check(mesh._h, nranks, tile_nodes, out), mesh=True)
    return dict(delivered=out[0], ghost_rows=out[1], errors=out[2], ghost_reading_tiles_not_last=out[3], waiting_transfer_blocks=out[4])


def generate_partition_plan(kind: int, dims, nranks: int, rank: int, level: int, mesh_variant: int = MESH_M6_WING,
                            lengths=(1.0, 1.0, 1.0), tilt: float = 0.05, apply_ewt: bool = False):
    """Host-only: `partition_plan` of the synthetic mesh `Mesh.generate(kind, dims, ...)` computed by rank-local generation, i.e.
    without assembling the mesh (mgcfd_generate_partition_plan).  `hash` covers everything the rank holds of the level."""
    L = lib()
    d = np.ascontiguousarray(np.asarray(dims, dtype=np.int64).reshape(-1, 3))
    ln = np.asarray(lengths, dtype=np.float64)
    lp = ln.ctypes.data_as(C.POINTER(C.c_double))
    info = (C.c_long * 8)()
    args = (kind, d.shape[0], _ptr(d), lp, mesh_variant, tilt, int(apply_ewt), nranks, rank, level)
    _check(L.mgcfd_generate_partition_plan(*args, info, None, None, None, None))
    owned, ghosts, sent = info[0], info[1], info[2]
    gid = np.empty(owned + ghosts, dtype=np.int64)
    sc, rc = np.zeros(nranks, dtype=np.int64), np.zeros(nranks, dtype=np.int64)
    sg = np.empty(max(sent, 1), dtype=np.int64)
    _check(L.mgcfd_generate_partition_plan(*args, info, _ptr(gid), _ptr(sc), _ptr(rc), _ptr(sg)))
    return dict(owned=owned, ghosts=ghosts, sent=sent, global_nodes=info[3], nI=info[4], nB=info[5], nW=info[6], hash=info[7], gid=gid,
                send_counts=sc, recv_counts=rc, send_gids=sg[:sent])


INFO_KEYS = ("nel", "nI", "nB", "nW", "npad", "ntiles", "tile_nodes", "max_rounds", "slots", "halo_entries", "cut_edges",
             "used_slots", "max_halo", "bslots", "smem_bytes", "pipe_grid")


def plan_level(mesh: Mesh, level: int, ordering: int = ORDER_PARTITION_RCM, tile_nodes: int = 0, flux_mode: int = FLUX_SORTED_SEGMENT):
    """Host-only integer preprocessing of one level: returns (info dict, new_of_old permutation, colouring conflicts)."""
    nel, nI, nB, nW, _ = mesh.dims(level)
    info = (C.c_long * 16)()
    perm = np.empty(nel, dtype=np.int64)
    conflicts = C.c_long(-1)
    c = mesh.coords(level)
    _check(lib().mgcfd_plan_level(nel, _ptr(c), nI, nB, nW, _ptr(mesh.edges(level)), ordering, tile_nodes, flux_mode, info, _ptr(perm),
                                  C.byref(conflicts)))
    d = dict(zip(INFO_KEYS, info))
    d["plan_hash"] = d.pop("pipe_grid")          # mgcfd_plan_level reports the plan's hash in the last slot
    return d, perm, conflicts.value


def plan_emulate_flux(level: dict, variables, mask: int = 7, ordering: int = ORDER_PARTITION_RCM, tile_nodes: int = 0,
                      flux_mode: int = FLUX_SORTED_SEGMENT):
    """Host-only checking aid: the fluxes the stage kernel's threads accumulate from the plan's tile headers / round blocks for
    `variables` (AoS [nel*5]), walked on the host (mgcfd_plan_emulate_flux).  `level` is a dict as conftest.mesh_levels makes them
    (edge weights already adjusted)."""
    var = np.ascontiguousarray(variables, dtype=np.float64).reshape(-1)
    out = np.zeros(5 * level["nel"])
    e = np.ascontiguousarray(level["edges"])
    _check(lib().mgcfd_plan_emulate_flux(level["nel"], _ptr(level.get("coords")), level["nI"], level["nB"], level["nW"], _ptr(e), ordering, tile_nodes,
                                         flux_mode, _ptr(var), mask, _ptr(out)))
    return out


def plan_emulate_visit_flux(level: dict, variables, supers: int, mask: int = 7):
    """Host-only checking aid: the fluxes the visit kernel's threads accumulate from the super-tile descriptors / halo lists /
    re-addressed edge rounds of a plan with `supers` super-tiles (mgcfd_plan_emulate_visit_flux).  Returns (fluxes, info dict)."""
    var = np.ascontiguousarray(variables, dtype=np.float64).reshape(-1)
    out = np.zeros(5 * level["nel"])
    e = np.ascontiguousarray(level["edges"])
    info = (C.c_long * 8)()
    _check(lib().mgcfd_plan_emulate_visit_flux(level["nel"], _ptr(level.get("coords")), level["nI"], level["nB"], level["nW"], _ptr(e), supers,
                                               _ptr(var), mask, _ptr(out), info))
    keys = ("supers", "max_tiles", "max_halo", "halo_total", "max_rounds", "tiles", "rows", "warp_tiles")
    return out, dict(zip(keys, info))


def plan_visit_config(level: dict, num_sms: int = 148):
    """Host-only: the visit-kernel configuration the library would choose for `level` on a device with `num_sms` SMs."""
    e = np.ascontiguousarray(level["edges"])
    info = (C.c_long * 8)()
    _check(lib().mgcfd_plan_visit_config(level["nel"], _ptr(level.get("coords")), level["nI"], level["nB"], level["nW"], _ptr(e), num_sms, info))
    return _visit_dict(info)


def plan_emulate_transfers(fine: dict, coarse: dict, var_f, res_f, res_c, var_c, ordering: int = ORDER_PARTITION_RCM, tile_nodes: int = 0):
    """Host-only checking aid: (restricted coarse variables, prolonged fine variables) computed from the transfer operators the
    device receives (mgcfd_plan_emulate_transfers)."""
    vf = np.ascontiguousarray(var_f, dtype=np.float64).reshape(-1)
    rf = np.ascontiguousarray(res_f, dtype=np.float64).reshape(-1)
    rc = np.ascontiguousarray(res_c, dtype=np.float64).reshape(-1)
    vc = np.array(var_c, dtype=np.float64).reshape(-1).copy()
    out_f = np.zeros_like(vf)
    ef, ec = np.ascontiguousarray(fine["edges"]), np.ascontiguousarray(coarse["edges"])
    mp = np.ascontiguousarray(fine["map"], dtype=np.int64)
    _check(lib().mgcfd_plan_emulate_transfers(fine["nel"], _ptr(fine["coords"]), fine["nI"], fine["nB"], fine["nW"], _ptr(ef), _ptr(mp),
                                              coarse["nel"], _ptr(coarse["coords"]), coarse["nI"], coarse["nB"], coarse["nW"], _ptr(ec),
                                              ordering, tile_nodes, _ptr(vf), _ptr(rf), _ptr(rc), _ptr(vc), _ptr(out_f)))
    return vc, out_f


def smooth_granular(s: Solver, level: int, legacy: bool):
    """One smoothing visit through the per-function API, call for call as main() (euler3d_cpu_double.cpp:383-512)."""
    s.copy_old_variables(level)
    s.compute_step_factor(level, legacy)
    for j in range(RK):
        s.compute_flux_edge(level)
        s.compute_boundary_flux_edge(level)
        s.compute_wall_flux_edge(level)
        s.time_step(level, j)
    s.residual(level)


def run_cycles_granular(s: Solver, cycles: int):
    """main()'s V-cycle loop (euler3d_cpu_double.cpp:371-694) through the per-function API."""
    legacy = s.mesh_variant == MESH_FVCORR
    rms_all, rms_var = [], []
    nl = s.levels
    for _ in range(cycles):
        smooth_granular(s, 0, legacy)
        a, v = s.calc_rms(0)
        rms_all.append(a)
        rms_var.append(v)
        for l in range(1, nl):
            s.mg_restrict(l)
            smooth_granular(s, l, legacy)
        for l in range(nl - 2, -1, -1):
            s.prolong(l)
            if l > 0:
                smooth_granular(s, l, legacy)
    return np.array(rms_all), np.array(rms_var)
