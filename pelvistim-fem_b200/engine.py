"""ctypes binding of ``libptfem.so`` (``include/ptfem.h``) — the only way host code reaches the
solver.  There is no CPU fallback: if the shared library is missing or no CUDA device is
present, construction fails loudly.

The classes mirror the stages of the ElmerSolver run the library replaces
(``step03_ankle_layers/run_layered_sweep.py:1099``): mesh upload -> pattern -> assembly ->
boundary conditions -> solve -> current recovery -> metric reductions.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_LIB = None
_LIB_PATH = Path(__file__).resolve().parent / "libptfem.so"

PRECOND_JACOBI, PRECOND_CHEBYSHEV, PRECOND_TWOLEVEL, PRECOND_AUTO = 0, 1, 2, -1
RECOVER_L2, RECOVER_LUMPED, RECOVER_AVERAGE = 0, 1, 2
SPMV_AUTO, SPMV_VECTOR, SPMV_STREAM, SPMV_STREAM1 = 0, 1, 2, 3
ERR_NOCONV = -4
_RECOVER = {"l2": RECOVER_L2, "lumped": RECOVER_LUMPED, "average": RECOVER_AVERAGE}


class PtfemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libptfem error {code}: {msg}")
        self.code = code


class SolveOpts(C.Structure):
    _fields_ = [("precond", C.c_int32), ("maxit", C.c_int32), ("check_every", C.c_int32),
                ("cheb_degree", C.c_int32), ("rtol", C.c_double), ("cheb_ratio", C.c_double),
                ("spmv_variant", C.c_int32), ("use_graph", C.c_int32), ("warm_start", C.c_int32), ("sample_spmv", C.c_int32),
                ("coarse_nodes", C.c_int32), ("coarse_levels", C.c_int32)]


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("nsys", C.c_int32),
                ("spmv_calls", C.c_int32), ("rel_residual", C.c_double), ("true_rel_residual", C.c_double),
                ("solve_ms", C.c_double), ("spmv_ms", C.c_double), ("setup_ms", C.c_double),
                ("precond", C.c_int32), ("coarse_unknowns", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Footprint(C.Structure):
    _fields_ = [("cx", C.c_double), ("cy", C.c_double), ("r", C.c_double), ("square", C.c_int32), ("pad_", C.c_int32)]


class MetricReq(C.Structure):
    """``ptfem_metric_req`` (include/ptfem.h): one reduction of a metric batch."""
    _fields_ = [("kind", C.c_int32), ("sys", C.c_int32), ("field", C.c_int32), ("mode", C.c_int32),
                ("zmin", C.c_double), ("zmax", C.c_double), ("scale_r", C.c_double), ("fp", Footprint * 2),
                ("nfp", C.c_int32), ("include_tris", C.c_int32), ("cen", C.c_double * 3), ("r0", C.c_double),
                ("mult", C.c_double * 4), ("nmult", C.c_int32), ("pad_", C.c_int32), ("z0", C.c_double), ("z1", C.c_double)]


METRIC_NODES, METRIC_PAD_CURRENT, METRIC_ROI, METRIC_OUT_STRIDE = 0, 1, 2, 24


def lib_path():
    return Path(os.environ.get("PTFEM_LIB", _LIB_PATH))


def load_library():
    """Load libptfem.so and declare the argument types of every entry point of ptfem.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise PtfemError(-100, f"{p} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
    L = C.CDLL(str(p))
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "ptfem_last_error": (C.c_char_p, []),
        "ptfem_version": (C.c_int, []),
        "ptfem_device_count": (C.c_int, [P(C.c_int)]),
        "ptfem_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
        "ptfem_ctx_create": (C.c_int, [C.c_int, P(vp)]),
        "ptfem_ctx_destroy": (C.c_int, [vp]),
        "ptfem_ctx_sync": (C.c_int, [vp]),
        "ptfem_ctx_launch_count": (C.c_int, [vp, P(i64)]),
        "ptfem_ctx_stream": (C.c_int, [vp, P(vp)]),
        "ptfem_mesh_create": (C.c_int, [vp, i64, vp, i64, vp, vp, i64, vp, vp, P(vp)]),
        "ptfem_mesh_create_async": (C.c_int, [vp, i64, vp, i64, vp, vp, i64, vp, vp, P(vp)]),
        "ptfem_mesh_destroy": (C.c_int, [vp]),
        "ptfem_mesh_set_coords": (C.c_int, [vp, vp]),
        "ptfem_pattern": (C.c_int, [vp, P(i64)]),
        "ptfem_window_plan_info": (C.c_int, [vp, P(i64), P(dbl)]),
        "ptfem_pattern_get": (C.c_int, [vp, vp, vp]),
        "ptfem_e2nnz_get": (C.c_int, [vp, vp]),
        "ptfem_assemble": (C.c_int, [vp, i32, vp, vp, i32]),
        "ptfem_values_get": (C.c_int, [vp, i32, i32, vp]),
        "ptfem_bc_reset": (C.c_int, [vp, i32]),
        "ptfem_bc_dirichlet": (C.c_int, [vp, i32, i32, dbl]),
        "ptfem_bc_neumann": (C.c_int, [vp, i32, i32, dbl]),
        "ptfem_bc_neumann_tris": (C.c_int, [vp, i32, i64, vp, dbl]),
        "ptfem_rhs_get": (C.c_int, [vp, i32, vp]),
        "ptfem_solve_opts_default": (None, [P(SolveOpts)]),
        "ptfem_solve": (C.c_int, [vp, P(SolveOpts), vp, P(SolveStats)]),
        "ptfem_solve_device": (C.c_int, [vp, P(SolveOpts), P(SolveStats)]),
        "ptfem_phi_get": (C.c_int, [vp, i32, vp]),
        "ptfem_phi_set": (C.c_int, [vp, i32, vp]),
        "ptfem_phi_get_all_async": (C.c_int, [vp, vp]),
        "ptfem_spmv": (C.c_int, [vp, i32, i32, i32, vp, vp]),
        "ptfem_spmv_bench": (C.c_int, [vp, i32, i32, P(dbl)]),
        "ptfem_element_fields": (C.c_int, [vp, i32, vp, vp]),
        "ptfem_recover_current": (C.c_int, [vp, i32, i32, vp]),
        "ptfem_recover_current_async": (C.c_int, [vp, i32, i32, vp]),
        "ptfem_current_get": (C.c_int, [vp, vp]),
        "ptfem_recover_current_batch": (C.c_int, [vp, i32, vp, i32]),
        "ptfem_metrics_batch": (C.c_int, [vp, i32, P(MetricReq), P(dbl)]),
        "ptfem_metric_nodes": (C.c_int, [vp, i32, i32, dbl, dbl, i32, vp, i32, dbl, P(dbl)]),
        "ptfem_metric_pad_current": (C.c_int, [vp, i32, dbl, P(Footprint), dbl, P(dbl)]),
        "ptfem_metric_roi": (C.c_int, [vp, i32, P(dbl), dbl, P(dbl), i32, dbl, dbl, i32, P(dbl)]),
        "ptfem_metric_column_fit": (C.c_int, [vp, i32, dbl, dbl, dbl, P(dbl)]),
        "ptfem_metric_jstats": (C.c_int, [vp, i32, dbl, P(dbl)]),
        "ptfem_metric_reaction": (C.c_int, [vp, i32, i32, P(dbl)]),
        "ptfem_sample_polyline": (C.c_int, [vp, i32, i64, vp, vp, vp]),
        "ptfem_dist_unique_id": (C.c_int, [C.c_char_p, vp]),
        "ptfem_dist_init": (C.c_int, [vp, C.c_char_p, vp, i32, i32]),
        "ptfem_dist_finalize": (C.c_int, [vp]),
        "ptfem_dist_system_create": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, P(vp)]),
        "ptfem_dist_solve": (C.c_int, [vp, P(SolveOpts), vp, P(SolveStats), P(dbl), P(dbl), P(dbl)]),
        "ptfem_dist_coarse_attach": (C.c_int, [vp, vp, i64]),
        "ptfem_mesh_set_bbox": (C.c_int, [vp, vp, vp]),
        "ptfem_dist_coarse_partial": (C.c_int, [vp, i64, i64, i32, i32, P(i64), vp, i64]),
        "ptfem_dist_coarse_finish": (C.c_int, [vp, vp, i64]),
        "ptfem_dist_coarse_ranges_get": (C.c_int, [vp, vp]),
        "ptfem_dist_coarse_ranges_set": (C.c_int, [vp, i32, vp]),
        "ptfem_dist_p2p_export": (C.c_int, [vp, vp]),
        "ptfem_dist_p2p_connect": (C.c_int, [vp, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)          # AttributeError here = header and library out of sync
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def exported_symbols():
    """Names declared in include/ptfem.h (parsed from the header)."""
    import re
    hdr = Path(__file__).resolve().parent.parent / "include" / "ptfem.h"
    txt = hdr.read_text()
    return sorted(set(re.findall(r"\b(ptfem_[a-z0-9_]+)\s*\(", txt)))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_gpu(device=0, info=None):
    """Restrict the calling process to the CPUs of the NUMA node the GPU is attached to (``/sys/bus/pci/devices/<bus id>/
    local_cpulist``), so that the host buffers it allocates afterwards - pinned staging for the mesh upload and the phi / J
    read-back above all - are first touched on that node and the copies do not cross the socket interconnect.  One process
    per GPU makes this matter: eight ranks stream 1.3 GB per sweep step each.  Returns the CPU set applied, or None when the
    topology is not exposed (single-node hosts, containers without sysfs) or the set would be empty or change nothing.
    ``info``: optional dict that receives what was found (bus id, local CPUs, allowed CPUs)."""
    import os
    info = info if info is not None else {}
    L = load_library()
    buf = C.create_string_buffer(32)
    if L.ptfem_device_pci_bus_id(int(device), buf, 32) != 0:
        info["why"] = "no PCI bus id"
        return None
    bus = buf.value.decode().lower()
    info["bus"] = bus
    try:
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            local = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
    except (OSError, ValueError, AttributeError) as e:
        info["why"] = f"topology not readable: {type(e).__name__}"
        return None
    info["local_cpus"], info["allowed_cpus"] = len(local), len(allowed)
    cpus = local & allowed
    if not cpus or cpus == allowed:
        info["why"] = "GPU-local CPUs are all the process may use already" if cpus else "no GPU-local CPU allowed"
        return None
    try:
        os.sched_setaffinity(0, cpus)
    except OSError as e:
        info["why"] = f"sched_setaffinity: {e}"
        return None
    return sorted(cpus)


class Context:
    """One per GPU (``ptfem_ctx``)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.device = device
        h = C.c_void_p()
        self._h = None
        self._ck(self.lib.ptfem_ctx_create(device, C.byref(h)))
        self._h = h

    def _ck(self, rc):
        if rc != 0:
            raise PtfemError(rc, self.lib.ptfem_last_error().decode(errors="replace"))

    def sync(self):
        self._ck(self.lib.ptfem_ctx_sync(self._h))

    @property
    def launches(self):
        n = C.c_int64()
        self._ck(self.lib.ptfem_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    @property
    def stream(self):
        s = C.c_void_p()
        self._ck(self.lib.ptfem_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    def mesh(self, nodes, tets, region, tris, bcid, prefetch=False):
        return DeviceMesh(self, nodes, tets, region, tris, bcid, prefetch=prefetch)

    def close(self):
        if self._h is not None:
            self.lib.ptfem_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    n = C.c_int(0)
    load_library().ptfem_device_count(C.byref(n))
    return n.value


class DeviceMesh:
    """A mesh resident on the GPU together with its pattern, matrices, right-hand sides and
    solution (``ptfem_mesh``)."""

    def __init__(self, ctx: Context, nodes, tets, region, tris, bcid, prefetch=False):
        """``prefetch=True``: the upload is only queued (``ptfem_mesh_create_async``) - give pinned arrays and leave them
        alone until ``pattern()``; a sweep uses it to send the mesh of its next point while the current one is being solved."""
        self.ctx = ctx
        self.lib = ctx.lib
        nodes, tets, region = _f64(nodes), _i32(tets), _i32(region)
        tris, bcid = _i32(tris), _i32(bcid)
        self.nn, self.nt, self.nb = nodes.shape[0], tets.shape[0], tris.shape[0]
        h = C.c_void_p()
        self._h = None
        create = self.lib.ptfem_mesh_create_async if prefetch else self.lib.ptfem_mesh_create
        self._ck(create(ctx._h, self.nn, _ptr(nodes), self.nt, _ptr(tets), _ptr(region), self.nb, _ptr(tris), _ptr(bcid), C.byref(h)))
        self._host = (nodes, tets, region, tris, bcid) if prefetch else None     # alive until the upload has been waited for
        self._h = h
        self.nnz = None
        self.nsys = 1
        self.last_stats = None

    _ck = Context._ck

    def close(self):
        if self._h is not None and self.ctx._h is not None:
            self.lib.ptfem_mesh_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- K1 -------------------------------------------------------------------
    def pattern(self):
        n = C.c_int64()
        self._ck(self.lib.ptfem_pattern(self._h, C.byref(n)))
        self._host = None
        self.nnz = n.value
        return self.nnz

    def window_plan(self):
        """Plan of the window SpMM for this pattern: dict(valid, tiles, wmax, capblob, line, plane, rows_per_row)."""
        info = (C.c_int64 * 6)()
        rpr = C.c_double()
        self._ck(self.lib.ptfem_window_plan_info(self._h, info, C.byref(rpr)))
        return dict(valid=bool(info[0]), tiles=int(info[1]), wmax=int(info[2]), capblob=int(info[3]), line=int(info[4]),
                    plane=int(info[5]), rows_per_row=rpr.value)

    def get_pattern(self):
        nnz = self.pattern()
        rowptr = np.empty(self.nn + 1, dtype=np.int32)
        col = np.empty(nnz, dtype=np.int32)
        self._ck(self.lib.ptfem_pattern_get(self._h, _ptr(rowptr), _ptr(col)))
        return rowptr, col

    def get_e2nnz(self):
        self.pattern()
        out = np.empty((self.nt, 16), dtype=np.int32)
        self._ck(self.lib.ptfem_e2nnz_get(self._h, _ptr(out)))
        return out

    def set_bbox(self, lo, hi):
        """Bounding box the coarse grids are laid over (distributed set-up: the WHOLE mesh's box on a rank's local mesh)."""
        lo, hi = _f64(lo), _f64(hi)
        self._ck(self.lib.ptfem_mesh_set_bbox(self._h, _ptr(lo), _ptr(hi)))

    def coarse_partial(self, nrows_owned, nn_global, coarse_nodes=0, coarse_levels=-1):
        """Raw Galerkin sums of the owned rows of this (local) mesh, to be added up over the ranks."""
        n = C.c_int64()
        self._ck(self.lib.ptfem_dist_coarse_partial(self._h, int(nrows_owned), int(nn_global), coarse_nodes, coarse_levels,
                                                    C.byref(n), None, 0))
        out = np.empty(n.value, dtype=np.float64)
        self._ck(self.lib.ptfem_dist_coarse_partial(self._h, int(nrows_owned), int(nn_global), coarse_nodes, coarse_levels,
                                                    C.byref(n), _ptr(out), out.size))
        return out

    def coarse_finish(self, sums):
        sums = _f64(sums)
        self._ck(self.lib.ptfem_dist_coarse_finish(self._h, _ptr(sums), sums.size))

    def set_coords(self, nodes):
        nodes = _f64(nodes)
        assert nodes.shape == (self.nn, 3)
        self._ck(self.lib.ptfem_mesh_set_coords(self._h, _ptr(nodes)))

    # -- K2/K3 -----------------------------------------------------------------
    def assemble(self, sigma_by_body):
        """``sigma_by_body``: dict body id -> conductivity, or a list of such dicts (batched
        matrices on one pattern; all dicts must have the same keys)."""
        batch = sigma_by_body if isinstance(sigma_by_body, (list, tuple)) else [sigma_by_body]
        ids = sorted(batch[0].keys())
        for d in batch:
            if sorted(d.keys()) != ids:
                raise ValueError("all conductivity sets of a batch must name the same bodies")
        reg = _i32(ids)
        sig = _f64([[d[i] for i in ids] for d in batch])
        self._ck(self.lib.ptfem_assemble(self._h, len(ids), _ptr(reg), _ptr(sig), len(batch)))
        self.nmat = len(batch)
        return self

    def get_values(self, sys=0, with_bc=False):
        if self.nnz is None:
            self.pattern()
        out = np.empty(self.nnz, dtype=np.float64)
        self._ck(self.lib.ptfem_values_get(self._h, sys, 1 if with_bc else 0, _ptr(out)))
        return out

    # -- K4 ----------------------------------------------------------------------
    def bc_reset(self, nrhs=1):
        self._ck(self.lib.ptfem_bc_reset(self._h, nrhs))
        self.nrhs = nrhs
        return self

    def dirichlet(self, bcid, value, rhs=-1):
        self._ck(self.lib.ptfem_bc_dirichlet(self._h, rhs, int(bcid), float(value)))
        return self

    def neumann(self, bcid, g, rhs=-1):
        self._ck(self.lib.ptfem_bc_neumann(self._h, rhs, int(bcid), float(g)))
        return self

    def neumann_tris(self, tri_idx, g, rhs=-1):
        idx = _i32(tri_idx)
        self._ck(self.lib.ptfem_bc_neumann_tris(self._h, rhs, idx.shape[0], _ptr(idx), float(g)))
        return self

    def get_rhs(self, rhs=0):
        out = np.empty(self.nn, dtype=np.float64)
        self._ck(self.lib.ptfem_rhs_get(self._h, rhs, _ptr(out)))
        return out

    # -- solve -------------------------------------------------------------------
    def _opts(self, **kw):
        o = SolveOpts()
        self.lib.ptfem_solve_opts_default(C.byref(o))
        names = {k for k, _ in SolveOpts._fields_}
        for k, v in kw.items():
            if k not in names:
                raise TypeError(f"unknown solver option {k!r}")
            setattr(o, k, v)
        return o

    def solve(self, to_host=True, raise_on_noconv=True, out=None, **opts):
        """PCG solve of every assembled system.  Returns phi [nsys, nn] (host) when
        ``to_host`` else ``None`` (solution stays on the device for post-processing).
        ``out``: optional preallocated (e.g. pinned) float64 array [nsys, nn] to receive phi."""
        o = self._opts(**opts)
        st = SolveStats()
        nsys = max(getattr(self, "nmat", 1), getattr(self, "nrhs", 1))
        self.nsys = nsys
        if to_host:
            phi = out if out is not None else np.empty((nsys, self.nn), dtype=np.float64)
            if phi.shape != (nsys, self.nn) or phi.dtype != np.float64 or not phi.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous float64 array of shape [nsys, nn]")
            rc = self.lib.ptfem_solve(self._h, C.byref(o), _ptr(phi), C.byref(st))
        else:
            phi = None
            rc = self.lib.ptfem_solve_device(self._h, C.byref(o), C.byref(st))
        self.last_stats = st.as_dict()
        if rc == ERR_NOCONV and not raise_on_noconv:
            return phi
        self._ck(rc)
        return phi

    def get_phi(self, sys=0):
        out = np.empty(self.nn, dtype=np.float64)
        self._ck(self.lib.ptfem_phi_get(self._h, sys, _ptr(out)))
        return out

    def get_phi_all_async(self, out):
        """All potentials [nsys, nn] into ``out`` (pinned) on the side stream; valid after ``Context.sync()`` / ``close()``."""
        if out.shape != (self.nsys, self.nn) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape [nsys, nn]")
        self._ck(self.lib.ptfem_phi_get_all_async(self._h, _ptr(out)))
        return out

    def set_phi(self, phi, sys=0):
        phi = _f64(phi)
        self._ck(self.lib.ptfem_phi_set(self._h, sys, _ptr(phi)))

    def spmv(self, x, sys=0, with_bc=False, variant=SPMV_AUTO):
        x = _f64(x)
        y = np.empty(self.nn, dtype=np.float64)
        self._ck(self.lib.ptfem_spmv(self._h, sys, 1 if with_bc else 0, variant, _ptr(x), _ptr(y)))
        return y

    def spmv_bench(self, variant=SPMV_AUTO, iters=20):
        ms = C.c_double()
        self._ck(self.lib.ptfem_spmv_bench(self._h, variant, iters, C.byref(ms)))
        return ms.value

    # -- K10/K11 ---------------------------------------------------------------------
    def element_fields(self, sys=0):
        E = np.empty((self.nt, 3), dtype=np.float64)
        J = np.empty((self.nt, 3), dtype=np.float64)
        self._ck(self.lib.ptfem_element_fields(self._h, sys, _ptr(E), _ptr(J)))
        return E, J

    def recover_current(self, sys=0, method="l2", to_host=True, out=None, wait=True):
        """Nodal current of one system.  ``wait=False`` (needs ``out`` in pinned memory): the read-back runs on a side
        stream and ``out`` is valid after ``Context.sync()``; it overlaps the metric calls that follow."""
        J = (out if out is not None else np.empty((self.nn, 3), dtype=np.float64)) if to_host else None
        if to_host and not wait:
            if out is None:
                raise ValueError("wait=False needs a caller-owned (pinned) out array")
            self._ck(self.lib.ptfem_recover_current_async(self._h, sys, _RECOVER[method], _ptr(J)))
        else:
            self._ck(self.lib.ptfem_recover_current(self._h, sys, _RECOVER[method], _ptr(J)))
        return J

    def recover_current_batch(self, method="lumped", to_host=False, out=None, wait=True):
        """Nodal currents of EVERY system of the last solve in two launches (``ptfem_recover_current_batch``).  Returns
        J [nsys, nn, 3] when ``to_host``; ``wait=False`` (``out`` in pinned memory) leaves the read-back running on the side
        stream until ``Context.sync()``.  Later per-system metric calls use these currents."""
        J = None
        if to_host:
            J = out if out is not None else np.empty((self.nsys, self.nn, 3), dtype=np.float64)
            if J.shape != (self.nsys, self.nn, 3) or J.dtype != np.float64 or not J.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous float64 array of shape [nsys, nn, 3]")
            if not wait and out is None:
                raise ValueError("wait=False needs a caller-owned (pinned) out array")
        self._ck(self.lib.ptfem_recover_current_batch(self._h, _RECOVER[method], _ptr(J), 1 if wait else 0))
        return J

    # -- K12 -------------------------------------------------------------------------
    def metrics_batch(self, reqs):
        """Any number of metric reductions in one pass per kind and ONE read-back (``ptfem_metrics_batch``).  ``reqs``: list of
        dicts ``dict(kind="nodes", sys, field, zmin, zmax=nan, mode=0, footprints=(), scale_r=1.0)``,
        ``dict(kind="pad_current", sys, zmin, footprint, scale_r=1.2)`` or
        ``dict(kind="roi", sys, cen, r0, mults=(1, 1.5, 2, 3), z0=0, z1=0, include_tris=True)``; returns one result dict (or
        list of dicts for "roi") per request, shaped like the single-request methods' results."""
        n = len(reqs)
        arr = (MetricReq * n)()
        for q, r in zip(arr, reqs):
            kind = r["kind"]
            q.sys = int(r.get("sys", 0))
            if kind == "nodes":
                q.kind, q.field, q.mode = METRIC_NODES, int(r["field"]), int(r.get("mode", 0))
                q.zmin, q.zmax, q.scale_r = float(r["zmin"]), float(r.get("zmax", float("nan"))), float(r.get("scale_r", 1.0))
                fps = list(r.get("footprints", ()))
                if len(fps) > 2:
                    raise ValueError("at most 2 footprints per batched request")
                for k, (cx, cy, rr, square) in enumerate(fps):
                    q.fp[k] = Footprint(cx, cy, rr, 1 if square else 0, 0)
                q.nfp = len(fps)
            elif kind == "pad_current":
                q.kind, q.zmin, q.scale_r = METRIC_PAD_CURRENT, float(r["zmin"]), float(r.get("scale_r", 1.2))
                cx, cy, rr, square = r["footprint"]
                q.fp[0] = Footprint(cx, cy, rr, 1 if square else 0, 0)
                q.nfp = 1
            elif kind == "roi":
                mults = tuple(r.get("mults", (1.0, 1.5, 2.0, 3.0)))
                q.kind, q.r0, q.nmult = METRIC_ROI, float(r["r0"]), len(mults)
                for k in range(3):
                    q.cen[k] = float(r["cen"][k])
                for k, mval in enumerate(mults):
                    q.mult[k] = float(mval)
                q.z0, q.z1, q.include_tris = float(r.get("z0", 0.0)), float(r.get("z1", 0.0)), 1 if r.get("include_tris", True) else 0
            else:
                raise ValueError(f"unknown metric kind {kind!r}")
        out = (C.c_double * (METRIC_OUT_STRIDE * n))()
        self._ck(self.lib.ptfem_metrics_batch(self._h, n, arr, out))
        res = []
        for i, r in enumerate(reqs):
            o = out[i * METRIC_OUT_STRIDE:(i + 1) * METRIC_OUT_STRIDE]
            if r["kind"] == "nodes":
                res.append(dict(count=int(o[0]), sum=o[1], max=o[2], min=o[3]))
            elif r["kind"] == "pad_current":
                res.append(dict(I_signed=o[0], area=o[1], count=int(o[2])))
            else:
                nm = len(tuple(r.get("mults", (1.0, 1.5, 2.0, 3.0))))
                res.append([dict(n=int(o[6 * k]), sum_J=o[6 * k + 1], sum_E=o[6 * k + 2], n_above=int(o[6 * k + 3]),
                                 n_mid=int(o[6 * k + 4]), n_below=int(o[6 * k + 5])) for k in range(nm)])
        return res

    @staticmethod
    def _fps(fps):
        arr = (Footprint * max(1, len(fps)))()
        for k, (cx, cy, r, square) in enumerate(fps):
            arr[k] = Footprint(cx, cy, r, 1 if square else 0, 0)
        return arr

    def metric_nodes(self, field, zmin, zmax=float("nan"), mode=0, footprints=(), scale_r=1.0, sys=0):
        """{count, sum, max, min} of field (0 |J|, 1 phi, 2 |J_z|, 3 J_z) over nodes with
        zmin < z (< zmax) and inside (mode 1) / outside all (mode 2) footprints."""
        out = (C.c_double * 4)()
        fps = self._fps(footprints)
        self._ck(self.lib.ptfem_metric_nodes(self._h, sys, field, zmin, zmax, mode, C.cast(fps, C.c_void_p),
                                             len(footprints), scale_r, out))
        return dict(count=int(out[0]), sum=out[1], max=out[2], min=out[3])

    def metric_pad_current(self, zmin, footprint, scale_r=1.2, sys=0):
        out = (C.c_double * 3)()
        fp = self._fps([footprint])
        self._ck(self.lib.ptfem_metric_pad_current(self._h, sys, zmin, fp, scale_r, out))
        return dict(I_signed=out[0], area=out[1], count=int(out[2]))

    def metric_roi(self, cen, r0, mults=(1.0, 1.5, 2.0, 3.0), z0=0.0, z1=0.0, include_tris=True, sys=0):
        n = len(mults)
        out = (C.c_double * (6 * n))()
        cen_a = (C.c_double * 3)(*cen)
        mul_a = (C.c_double * n)(*mults)
        self._ck(self.lib.ptfem_metric_roi(self._h, sys, cen_a, r0, mul_a, n, z0, z1, 1 if include_tris else 0, out))
        return [dict(n=int(out[6 * k]), sum_J=out[6 * k + 1], sum_E=out[6 * k + 2], n_above=int(out[6 * k + 3]),
                     n_mid=int(out[6 * k + 4]), n_below=int(out[6 * k + 5])) for k in range(n)]

    def metric_column_fit(self, cx, cy, rad, sys=0):
        out = (C.c_double * 6)()
        self._ck(self.lib.ptfem_metric_column_fit(self._h, sys, cx, cy, rad, out))
        return list(out)

    def metric_jstats(self, shift=0.0, sys=0):
        out = (C.c_double * 3)()
        self._ck(self.lib.ptfem_metric_jstats(self._h, sys, float(shift), out))
        return list(out)

    def metric_reaction(self, bcid, sys=0):
        out = C.c_double()
        self._ck(self.lib.ptfem_metric_reaction(self._h, sys, int(bcid), C.byref(out)))
        return out.value

    # -- K13 ---------------------------------------------------------------------------
    def sample_polyline(self, pts, sys=0):
        pts = _f64(pts)
        n = pts.shape[0]
        phi = np.empty(n, dtype=np.float64)
        af = np.empty(n, dtype=np.float64)
        self._ck(self.lib.ptfem_sample_polyline(self._h, sys, n, _ptr(pts), _ptr(phi), _ptr(af)))
        return phi, af


def solve_case(ctx: Context, mesh, sigma_by_body, dirichlet, neumann, recover="l2", **opts):
    """One full solve step on the GPU with the call shape of ``oracle.fem_oracle.solve_case``
    (used by the drivers and by the parity tests).  Returns dict(phi, J, stats, dmesh)."""
    dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    dm.assemble(sigma_by_body)
    dm.bc_reset(1)
    if not dirichlet:
        raise ValueError("pure Neumann problem (no Potential BC) is singular")
    for bid, g in neumann:
        dm.neumann(bid, g)
    for bid, v in dirichlet:
        dm.dirichlet(bid, v)
    phi = dm.solve(**opts)[0]
    J = dm.recover_current(0, recover) if recover else None
    return dict(phi=phi, J=J, stats=dm.last_stats, dmesh=dm)


# -- multi-GPU: row-partitioned single solve ------------------------------------------------------------
def nccl_library_path():
    """Path of the NCCL shared library torch ships (``nvidia/nccl/lib/libnccl.so.2``), else the soname."""
    try:
        import nvidia.nccl as _n
        base = Path(list(_n.__path__)[0]) / "lib"
        for name in ("libnccl.so.2", "libnccl.so"):
            if (base / name).exists():
                return str(base / name)
    except Exception:
        pass
    return "libnccl.so.2"


def dist_unique_id(libnccl=None):
    """128-byte ncclUniqueId (rank 0 creates it, the launcher broadcasts it)."""
    buf = C.create_string_buffer(128)
    L = load_library()
    rc = L.ptfem_dist_unique_id((libnccl or nccl_library_path()).encode(), buf)
    if rc != 0:
        raise PtfemError(rc, L.ptfem_last_error().decode(errors="replace"))
    return buf.raw


def dist_init(ctx: Context, unique_id, rank, nranks, libnccl=None):
    """``unique_id`` = the 128-byte ncclUniqueId, or ``None`` for the peer-memory-only transport (no NCCL)."""
    if unique_id is None:
        ctx._ck(ctx.lib.ptfem_dist_init(ctx._h, None, None, rank, nranks))
        return
    buf = C.create_string_buffer(unique_id, 128)
    ctx._ck(ctx.lib.ptfem_dist_init(ctx._h, (libnccl or nccl_library_path()).encode(), buf, rank, nranks))


def dist_finalize(ctx: Context):
    ctx._ck(ctx.lib.ptfem_dist_finalize(ctx._h))


class DistSystem:
    """One rank's block of a row-partitioned system on its GPU (``ptfem_dist_system_create``)."""

    def __init__(self, ctx: Context, blk):
        self.ctx, self.lib, self.nloc = ctx, ctx.lib, blk.nloc
        h = C.c_void_p()
        self._h = None
        nnbr = int(blk.nbr_rank.shape[0])
        arrs = [_i32(blk.rowptr), _i32(blk.col), _f64(blk.val), _f64(blk.b), _i32(blk.nbr_rank), _i32(blk.send_ptr),
                _i32(blk.send_idx), _i32(blk.recv_ptr)]
        ctx._ck(self.lib.ptfem_dist_system_create(ctx._h, blk.nloc, blk.nhalo, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                                  _ptr(arrs[3]), nnbr, _ptr(arrs[4]) if nnbr else None,
                                                  _ptr(arrs[5]) if nnbr else None, _ptr(arrs[6]) if nnbr else None,
                                                  _ptr(arrs[7]) if nnbr else None, C.byref(h)))
        self._h = h

    def coarse_attach(self, replica, row0):
        """Attach the coarse-grid preconditioner: ``replica`` is this rank's full :class:`DeviceMesh` with the same matrix
        and boundary conditions; rows ``[row0, row0 + nloc)`` of its coarse spaces are taken (``ptfem_dist_coarse_attach``).
        Call before :meth:`p2p_export`."""
        self.ctx._ck(self.lib.ptfem_dist_coarse_attach(self._h, replica._h, int(row0)))

    def coarse_ranges(self):
        """This rank's {a0, b0, a1, b1} of the sharded coarse exchange (``ptfem_dist_coarse_ranges_get``)."""
        out = np.zeros(4, dtype=np.int64)
        self.ctx._ck(self.lib.ptfem_dist_coarse_ranges_get(self._h, _ptr(out)))
        return out

    def set_coarse_ranges(self, all_ranges):
        """Every rank's ranges, [nranks, 4] (all-gathered by the launcher), before :meth:`p2p_connect`."""
        a = np.ascontiguousarray(all_ranges, dtype=np.int64)
        self.ctx._ck(self.lib.ptfem_dist_coarse_ranges_set(self._h, a.shape[0], _ptr(a)))

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.ctx._ck(self.lib.ptfem_dist_p2p_export(self._h, buf))
        return buf.raw

    def p2p_connect(self, all_handles, halo_src):
        """``all_handles``: list (by rank) of the 128-byte exports; ``halo_src``: owner-local row of every halo slot."""
        blob = b"".join(all_handles)
        src = _i32(halo_src)
        self.ctx._ck(self.lib.ptfem_dist_p2p_connect(self._h, len(all_handles), blob, _ptr(src) if src.size else None))

    def solve(self, **opts):
        o = SolveOpts()
        self.lib.ptfem_solve_opts_default(C.byref(o))
        for k, v in opts.items():
            setattr(o, k, v)
        st = SolveStats()
        x = np.empty(self.nloc, dtype=np.float64)
        t = [C.c_double(), C.c_double(), C.c_double()]
        rc = self.lib.ptfem_dist_solve(self._h, C.byref(o), _ptr(x), C.byref(st), C.byref(t[0]), C.byref(t[1]), C.byref(t[2]))
        self.last_stats = st.as_dict()
        self.timings = dict(spmv_ms=t[0].value, halo_ms=t[1].value, allreduce_ms=t[2].value)
        self.ctx._ck(rc)
        return x

    def close(self):
        if self._h is not None and self.ctx._h is not None:
            self.lib.ptfem_mesh_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
