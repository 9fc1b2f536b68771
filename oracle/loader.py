"""ctypes access to the two CPU checkers (test infrastructure; see oracle/mgcfd_oracle.c header):

* ``Oracle``    -- oracle/libmgcfd_oracle.so, our plain-C restatement (always buildable: ``make -C oracle oracle``)
* ``Reference`` -- oracle/_ref/libmgcfd_ref[_omp].so, the UNMODIFIED reference compiled from /root/reference with
                   oracle/ref_shim.cpp (``make -C oracle ref``); present wherever it was prebuilt.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EDGE_DTYPE = np.dtype([("a", np.int64), ("b", np.int64), ("x", np.float64), ("y", np.float64), ("z", np.float64)])
vp, dp, lp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_long)


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def build_reference(ref="/root/reference"):
    if not os.path.isdir(ref):
        raise FileNotFoundError(ref)
    subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={ref}"])


class OrcLevel(C.Structure):
    _fields_ = [("nel", C.c_long), ("nI", C.c_long), ("nB", C.c_long), ("nW", C.c_long),
                ("vol", vp), ("edges", vp), ("coords", vp), ("map", vp),
                ("var", vp), ("old", vp), ("res", vp), ("flux", vp), ("sf", vp)]


class Oracle:
    """Plain-C restatement; all arrays numpy, reference layout (AoS)."""

    def __init__(self):
        path = os.path.join(HERE, "libmgcfd_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.orc_calc_rms.restype = C.c_double
        L.orc_check_invalid.restype = C.c_long
        L.orc_run_cycles.restype = C.c_long
        L.orc_calc_rms.argtypes = [C.c_long, vp]
        L.orc_flux_edge.argtypes = [C.c_long, C.c_long, vp, vp, vp]
        L.orc_boundary_flux_edge.argtypes = [C.c_long, C.c_long, vp, vp, vp]
        L.orc_wall_flux_edge.argtypes = [C.c_long, C.c_long, vp, vp, vp, vp, vp]
        L.orc_indirect_rw.argtypes = [C.c_long, C.c_long, vp, vp, vp]
        L.orc_step_factor.argtypes = [C.c_long, vp, vp, vp]
        L.orc_step_factor_legacy.argtypes = [C.c_long, vp, vp, vp]
        L.orc_time_step.argtypes = [C.c_int, C.c_long, vp, vp, vp, vp]
        L.orc_residual.argtypes = [C.c_long, vp, vp, vp]
        L.orc_rms_per_var.argtypes = [C.c_long, vp, vp]
        L.orc_check_invalid.argtypes = [vp, C.c_long, C.POINTER(C.c_int)]
        L.orc_mg_restrict.argtypes = [vp, vp, C.c_long, vp, vp, C.c_long]
        L.orc_prolong.argtypes = [vp, C.c_long, vp, vp, vp, C.c_long, vp, vp, vp]
        L.orc_adjust_dampen.argtypes = [C.c_int, vp, C.c_long, vp]
        L.orc_far_field.argtypes = [vp, vp]
        L.orc_run_cycles.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, vp]

    def far_field(self):
        v, c = np.zeros(5), np.zeros(12)
        self.L.orc_far_field(_p(v), _p(c))
        return v, c

    def adjust_dampen(self, variant, coords, edges):
        self.L.orc_adjust_dampen(variant, _p(coords), edges.shape[0], _p(edges))

    def step_factor(self, var, vol, legacy=False):
        sf = np.zeros(vol.shape[0])
        (self.L.orc_step_factor_legacy if legacy else self.L.orc_step_factor)(vol.shape[0], _p(var), _p(vol), _p(sf))
        return sf

    def flux_edge(self, first, n, edges, var, flux): self.L.orc_flux_edge(first, n, _p(edges), _p(var), _p(flux))
    def boundary_flux_edge(self, first, n, edges, var, flux): self.L.orc_boundary_flux_edge(first, n, _p(edges), _p(var), _p(flux))

    def wall_flux_edge(self, first, n, edges, var, flux):
        v, c = self.far_field()
        self.L.orc_wall_flux_edge(first, n, _p(edges), _p(var), _p(flux), _p(v), _p(c))

    def indirect_rw(self, first, n, edges, var, flux): self.L.orc_indirect_rw(first, n, _p(edges), _p(var), _p(flux))
    def time_step(self, j, sf, flux, old, var): self.L.orc_time_step(j, sf.shape[0], _p(sf), _p(flux), _p(old), _p(var))

    def residual(self, old, var):
        r = np.zeros_like(var)
        self.L.orc_residual(var.size // 5, _p(old), _p(var), _p(r))
        return r

    def calc_rms(self, res): return self.L.orc_calc_rms(res.size // 5, _p(res))

    def rms_per_var(self, res):
        o = np.zeros(5)
        self.L.orc_rms_per_var(res.size // 5, _p(res), _p(o))
        return o

    def check_invalid(self, var):
        why = C.c_int()
        cell = self.L.orc_check_invalid(_p(var), var.size // 5, C.byref(why))
        return None if cell < 0 else (cell, why.value)

    def mg_restrict(self, var1, var2, mapping):
        scratch = np.zeros(var2.size // 5, dtype=np.int64)
        self.L.orc_mg_restrict(_p(var1), _p(var2), var2.size // 5, _p(mapping), _p(scratch), mapping.shape[0])

    def prolong(self, edges, nI, res1, res2, var2, mapping, coords1, coords2):
        self.L.orc_prolong(_p(edges), nI, _p(res1), _p(res2), _p(var2), var2.size // 5, _p(mapping), _p(coords1), _p(coords2))

    def run_cycles(self, variant, levels, cycles):
        """levels: list of dicts with nel,nI,nB,nW,vol,edges,coords,map (numpy, ewt already applied).
        Returns (rms_all, rms_var, per-level state dicts)."""
        arr = (OrcLevel * len(levels))()
        state = []
        ffv, _ = self.far_field()
        for i, lv in enumerate(levels):
            n = lv["nel"]
            st = {"var": np.tile(ffv, n), "old": np.zeros(5 * n), "res": np.zeros(5 * n), "flux": np.zeros(5 * n), "sf": np.zeros(n)}
            state.append(st)
            arr[i].nel, arr[i].nI, arr[i].nB, arr[i].nW = n, lv["nI"], lv["nB"], lv["nW"]
            arr[i].vol, arr[i].edges = _p(lv["vol"]), _p(lv["edges"])
            arr[i].coords, arr[i].map = _p(lv.get("coords")), _p(lv.get("map"))
            for k in ("var", "old", "res", "flux", "sf"):
                setattr(arr[i], k, _p(st[k]))
        ra, rv = np.zeros(cycles), np.zeros((cycles, 5))
        rc = self.L.orc_run_cycles(len(levels), variant, arr, cycles, _p(ra), _p(rv))
        if rc:
            raise FloatingPointError(f"invalid variables at cell {rc - 1}")
        return ra, rv, state


def reference_available(omp=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libmgcfd_ref_omp.so" if omp else "libmgcfd_ref.so"))


class Reference:
    """The unmodified reference kernels through oracle/ref_shim.cpp. One instance per process per variant
    (the reference keeps `levels`, `mesh_variant`, far-field state in globals)."""

    def __init__(self, omp=False):
        path = os.path.join(HERE, "_ref", "libmgcfd_ref_omp.so" if omp else "libmgcfd_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build with `make -C oracle ref REF=/root/reference` where the reference is mounted")
        L = self.L = C.CDLL(path)
        L.refs_calc_rms.restype = C.c_double
        L.refs_run.restype = C.c_double
        L.refs_create.restype = vp
        L.refs_ptr.restype = vp
        L.refs_set_globals.argtypes = [C.c_int, C.c_int]
        L.refs_far_field.argtypes = [vp, vp]
        L.refs_compute_step_factor.argtypes = [C.c_long, vp, vp, vp, C.c_int]
        for f in ("refs_compute_flux_edge", "refs_compute_boundary_flux_edge", "refs_compute_wall_flux_edge", "refs_indirect_rw"):
            getattr(L, f).argtypes = [C.c_long, C.c_long, vp, vp, vp]
        L.refs_time_step.argtypes = [C.c_int, C.c_long, vp, vp, vp, vp]
        L.refs_residual.argtypes = [C.c_long, vp, vp, vp]
        L.refs_calc_rms.argtypes = [C.c_long, vp]
        L.refs_mg_restrict.argtypes = [vp, vp, C.c_long, vp, vp, C.c_long]
        L.refs_prolong.argtypes = [vp, C.c_long, vp, vp, vp, C.c_long, vp, vp, vp]
        L.refs_adjust_ewt.argtypes = [vp, C.c_long, vp]
        L.refs_dampen_ewt.argtypes = [C.c_long, vp, C.c_double]
        L.refs_read_grid.argtypes = [C.c_char_p, lp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.refs_read_mg.argtypes = [C.c_char_p, C.POINTER(vp), lp]
        L.refs_free.argtypes = [vp]
        L.refs_read_input_dat.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        L.refs_create.argtypes = [C.c_int, C.c_int]
        L.refs_set_mesh.argtypes = [vp, C.c_int, C.c_long, vp, vp, C.c_long, C.c_long, C.c_long, vp, vp, C.c_long]
        L.refs_duplicate.argtypes = [vp, C.c_int]
        L.refs_prepare.argtypes = [vp]
        L.refs_ptr.argtypes = [vp, C.c_int, C.c_int]
        L.refs_dims.argtypes = [vp, C.c_int, lp]
        L.refs_run.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.refs_times.argtypes = [vp, vp]
        L.refs_destroy.argtypes = [vp]
        assert L.refs_sizeof_edge() == 40

    def threads(self): return self.L.refs_omp_threads()
    def set_globals(self, levels, variant): self.L.refs_set_globals(levels, variant)

    def far_field(self):
        v, c = np.zeros(5), np.zeros(12)
        self.L.refs_far_field(_p(v), _p(c))
        return v, c

    def step_factor(self, var, vol, legacy=False):
        sf = np.zeros(vol.shape[0])
        self.L.refs_compute_step_factor(vol.shape[0], _p(var), _p(vol), _p(sf), int(legacy))
        return sf

    def flux_edge(self, first, n, edges, var, flux): self.L.refs_compute_flux_edge(first, n, _p(edges), _p(var), _p(flux))
    def boundary_flux_edge(self, first, n, edges, var, flux): self.L.refs_compute_boundary_flux_edge(first, n, _p(edges), _p(var), _p(flux))

    def wall_flux_edge(self, first, n, edges, var, flux):
        self.far_field()
        self.L.refs_compute_wall_flux_edge(first, n, _p(edges), _p(var), _p(flux))

    def indirect_rw(self, first, n, edges, var, flux): self.L.refs_indirect_rw(first, n, _p(edges), _p(var), _p(flux))
    def time_step(self, j, sf, flux, old, var): self.L.refs_time_step(j, sf.shape[0], _p(sf), _p(flux), _p(old), _p(var))

    def residual(self, old, var):
        r = np.zeros_like(var)
        self.L.refs_residual(var.size // 5, _p(old), _p(var), _p(r))
        return r

    def calc_rms(self, res): return self.L.refs_calc_rms(res.size // 5, _p(res))

    def mg_restrict(self, var1, var2, mapping):
        scratch = np.zeros(max(var2.size // 5, var1.size // 5), dtype=np.int64)
        self.L.refs_mg_restrict(_p(var1), _p(var2), var2.size // 5, _p(mapping), _p(scratch), mapping.shape[0])

    def prolong(self, edges, nI, res1, res2, var2, mapping, coords1, coords2):
        self.L.refs_prolong(_p(edges), nI, _p(res1), _p(res2), _p(var2), var2.size // 5, _p(mapping), _p(coords1), _p(coords2))

    def adjust_dampen(self, variant, coords, edges):
        damp = {2: 5e-8, 3: 1e-7, 4: 2e-7}.get(variant)
        if damp is None:
            return
        self.L.refs_adjust_ewt(_p(coords), edges.shape[0], _p(edges))
        self.L.refs_dampen_ewt(edges.shape[0], _p(edges), damp)

    def read_grid(self, path, levels, variant):
        """read_grid (io.cpp:14-199) -> dict(nel,nI,nB,nW,vol,edges,coords)."""
        self.set_globals(levels, variant)
        hdr = (C.c_long * 8)()
        vol, ed, co = vp(), vp(), vp()
        self.L.refs_read_grid(path.encode(), hdr, C.byref(vol), C.byref(ed), C.byref(co))
        nel, ne, nI, nB, nW = hdr[0], hdr[1], hdr[2], hdr[3], hdr[4]
        out = {"nel": nel, "nI": nI, "nB": nB, "nW": nW, "starts": (hdr[5], hdr[6], hdr[7]),
               "vol": np.ctypeslib.as_array(C.cast(vol, dp), (nel,)).copy(),
               "edges": np.frombuffer((C.c_char * (40 * ne)).from_address(ed.value), dtype=EDGE_DTYPE, count=ne).copy(),
               "coords": np.ctypeslib.as_array(C.cast(co, dp), (nel, 3)).copy()}
        for p in (vol, ed, co):
            self.L.refs_free(p)
        return out

    def read_mg(self, path):
        mg, n = vp(), C.c_long()
        self.L.refs_read_mg(path.encode(), C.byref(mg), C.byref(n))
        out = np.ctypeslib.as_array(C.cast(mg, lp), (n.value,)).copy()
        self.L.refs_free(mg)
        return out

    def read_input_dat(self, path):
        size, nl, var = C.c_int(), C.c_int(), C.c_int()
        buf = C.create_string_buffer(1 << 16)
        rc = self.L.refs_read_input_dat(path.encode(), C.byref(size), C.byref(nl), C.byref(var), buf, len(buf))
        assert rc == 0
        names = buf.value.decode().split("\n")[:-1]
        return {"size": size.value, "levels": nl.value, "variant": var.value, "layers": names[:nl.value], "mg": names[nl.value:]}

    # ---- sessions: multi-level runs sequenced like main() ----
    def session(self, variant, levels):
        """levels: list of dicts nel,nI,nB,nW,vol,edges,coords,map with RAW (un-adjusted) edge weights."""
        s = self.L.refs_create(len(levels), variant)
        for i, lv in enumerate(levels):
            mp = lv.get("map")
            self.L.refs_set_mesh(s, i, lv["nel"], _p(lv["vol"]), _p(lv.get("coords")), lv["nI"], lv["nB"], lv["nW"], _p(lv["edges"]),
                                 _p(mp), 0 if mp is None else mp.shape[0])
        return RefSession(self, s, len(levels))


class RefSession:
    def __init__(self, ref, handle, nl):
        self.ref, self.h, self.nl = ref, handle, nl

    def duplicate(self, m): self.ref.L.refs_duplicate(self.h, m)
    def prepare(self): self.ref.L.refs_prepare(self.h)

    def dims(self, l):
        o = (C.c_long * 5)()
        self.ref.L.refs_dims(self.h, l, o)
        return tuple(o)

    def field(self, l, field):
        """0 variables 1 old 2 residuals 3 fluxes 4 step_factors 5 volumes 6 edges"""
        nel, nI, nB, nW, _ = self.dims(l)
        p = self.ref.L.refs_ptr(self.h, l, field)
        if field == 6:
            ne = nI + nB + nW
            return np.frombuffer((C.c_char * (40 * ne)).from_address(p), dtype=EDGE_DTYPE, count=ne).copy()
        n = nel if field in (4, 5) else 5 * nel
        return np.ctypeslib.as_array(C.cast(p, dp), (n,)).copy()

    def run(self, cycles, probe=False):
        ra, rv = np.zeros(cycles), np.zeros((cycles, 5))
        t = self.ref.L.refs_run(self.h, cycles, int(probe), _p(ra), _p(rv))
        return ra, rv, t

    def times(self):
        o = np.zeros((5, 8))
        self.ref.L.refs_times(self.h, _p(o))
        return dict(zip(("flux", "compute_step", "time_step", "restrict", "prolong"), o))

    def close(self):
        if self.h:
            self.ref.L.refs_destroy(self.h)
            self.h = None
