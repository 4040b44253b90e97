"""CPU ORACLE (test infrastructure only): ctypes wrapper of ``oracle/fem_c.c`` (``liboracle_c.so``).

Same arithmetic as ``fem_oracle.py`` (P1 assembly of the ``StatCurrentSolve`` problem,
``step01_box/case.sif:33-45``), in C + OpenMP so that the 20 M-tet benchmark mesh can be
assembled and iterated on the host: used by ``bench.py``'s ``cpu_baseline`` leg and
``--impl reference`` arm, and cross-checked against ``fem_oracle.py`` in ``tests/``.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = None


def build():
    subprocess.run(["make", "-C", str(_DIR), "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        p = _DIR / "liboracle_c.so"
        if not p.exists():
            build()
        L = C.CDLL(str(p))
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.oc_threads.restype = C.c_int
        L.oc_set_threads.argtypes = [C.c_int]
        L.oc_pattern_build.restype = i64
        L.oc_pattern_build.argtypes = [i64, i64, vp, vp]
        L.oc_pattern_col.argtypes = [vp]
        L.oc_assemble.restype = C.c_int
        L.oc_assemble.argtypes = [i64, i64, vp, vp, vp, vp, vp]
        L.oc_neumann.argtypes = [vp, i64, vp, vp, i32, dbl, vp]
        L.oc_dirichlet.argtypes = [i64, vp, vp, vp, vp, vp]
        L.oc_spmv.argtypes = [i64, vp, vp, vp, vp, vp]
        L.oc_pcg.restype = C.c_int
        L.oc_pcg.argtypes = [i64, vp, vp, vp, vp, vp, dbl, C.c_int, C.POINTER(dbl)]
        L.oc_coarse_locate.argtypes = [i64, vp, vp, vp, vp, vp, vp]
        L.oc_galerkin_dense.restype = C.c_int
        L.oc_galerkin_dense.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp]
        L.oc_galerkin_diag.restype = C.c_int
        L.oc_galerkin_diag.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp]
        L.oc_pcg_coarse.restype = C.c_int
        L.oc_pcg_coarse.argtypes = [i64, vp, vp, vp, vp, vp, dbl, C.c_int, C.POINTER(dbl), C.c_int, vp, vp, vp, vp, vp, vp]
        L.oc_recover_lumped.restype = C.c_int
        L.oc_recover_lumped.argtypes = [i64, vp, vp, vp, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def threads():
    return lib().oc_threads()


def set_threads(n):
    lib().oc_set_threads(int(n))


def host_cores():
    """Cores this process may run on (the affinity mask, not ``os.cpu_count()``: containers are often pinned)."""
    import os
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_cores():
    """Pin the OpenMP thread count to every core the process may use.  ``torch.distributed.run`` exports
    ``OMP_NUM_THREADS=1`` to its children, which would otherwise leave the oracle on one thread."""
    n = host_cores()
    set_threads(n)
    return n


class CSystem:
    """Assembled, BC-eliminated system of one case (all host arrays)."""

    def __init__(self, mesh, sigma_by_body, dirichlet, neumann):
        L = lib()
        nodes = np.ascontiguousarray(mesh.nodes, dtype=np.float64)
        tets = np.ascontiguousarray(mesh.tets, dtype=np.int32)
        tris = np.ascontiguousarray(mesh.tris, dtype=np.int32)
        bcid = np.ascontiguousarray(mesh.bcid, dtype=np.int32)
        nn, nt, nb = nodes.shape[0], tets.shape[0], tris.shape[0]
        self.nn = nn
        self.rowptr = np.empty(nn + 1, dtype=np.int32)
        nnz = L.oc_pattern_build(nn, nt, _p(tets), _p(self.rowptr))
        if nnz < 0:
            raise MemoryError("oc_pattern_build failed")
        self.col = np.empty(nnz, dtype=np.int32)
        L.oc_pattern_col(_p(self.col))
        sig = np.empty(nt, dtype=np.float64)
        seen = np.zeros(nt, dtype=bool)
        for body, s in sigma_by_body.items():
            m = mesh.region == body
            sig[m] = s
            seen |= m
        if not seen.all():
            raise ValueError("element without conductivity")
        self.val_raw = np.empty(nnz, dtype=np.float64)
        if L.oc_assemble(nn, nt, _p(nodes), _p(tets), _p(sig), _p(self.rowptr), _p(self.val_raw)) != 0:
            raise RuntimeError("oc_assemble failed")
        self.b = np.zeros(nn, dtype=np.float64)
        for bid, g in neumann:
            L.oc_neumann(_p(nodes), nb, _p(tris), _p(bcid), int(bid), float(g), _p(self.b))
        self.b_neumann = self.b.copy()
        isdir = np.zeros(nn, dtype=np.uint8)
        dval = np.zeros(nn, dtype=np.float64)
        for bid, v in dirichlet:
            idx = np.unique(tris[bcid == bid].ravel())
            isdir[idx] = 1
            dval[idx] = v
        self.val = self.val_raw.copy()
        L.oc_dirichlet(nn, _p(self.rowptr), _p(isdir), _p(dval), _p(self.val), _p(self.b))
        self.isdir = isdir
        self.nodes, self.tets, self.sigma_e = nodes, tets, sig
        self.coarse = None

    def spmv(self, x, raw=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.nn, dtype=np.float64)
        lib().oc_spmv(self.nn, _p(self.rowptr), _p(self.col), _p(self.val_raw if raw else self.val), _p(x), _p(y))
        return y

    def pcg(self, rtol=1e-12, maxit=100000, x0=None):
        x = np.zeros(self.nn, dtype=np.float64) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        rel = C.c_double()
        it = lib().oc_pcg(self.nn, _p(self.rowptr), _p(self.col), _p(self.val), _p(self.b), _p(x), float(rtol),
                          int(maxit), C.byref(rel))
        if it < 0:
            raise MemoryError("oc_pcg failed")
        return x, it, rel.value

    # -- Jacobi + geometric coarse grids (restates coarse_oracle.py in C; same grid / level / weight rules) --------
    def coarse_setup(self, coarse_nodes=300, extra_levels=-1, level_weight=None):
        """Grids, Galerkin operators and their inverses for :meth:`pcg_coarse`.  Returns the set-up dict."""
        from . import coarse_oracle as cor
        L = lib()
        lo, hi = self.nodes.min(axis=0), self.nodes.max(axis=0)
        base = cor.choose_grid(lo, hi, float(coarse_nodes if coarse_nodes > 0 else 300))
        nlev = cor.level_count(self.nn, base, extra_levels)
        if level_weight is None:
            level_weight = 2.0 / (nlev + 1)
        ext = np.where(hi - lo > 0.0, hi - lo, 1.0)
        ns, bdiag, node0_f, t_f, bdense = [], [], None, None, None
        for l in range(nlev):                                  # 0 = finest ... nlev-1 = coarsest (exact)
            n = (base * (1 << (nlev - 1 - l))).astype(np.int32)
            inv_h = np.ascontiguousarray(n / (ext * (1.0 + 1e-12)), dtype=np.float64)
            node0 = np.empty(self.nn, dtype=np.int32)
            t = np.empty((self.nn, 3), dtype=np.float64)
            lo_c = np.ascontiguousarray(lo, dtype=np.float64)
            L.oc_coarse_locate(self.nn, _p(self.nodes), _p(lo_c), _p(inv_h), _p(n), _p(node0), _p(t))
            k = int(np.prod(n + 1))
            if l == nlev - 1:
                E = np.empty((k, k), dtype=np.float64)
                if L.oc_galerkin_dense(self.nn, _p(self.rowptr), _p(self.col), _p(self.val), _p(self.isdir), _p(node0), _p(t),
                                       _p(n), _p(E)) != 0:
                    raise MemoryError("oc_galerkin_dense failed")
                d = E.diagonal().copy()
                empty = ~(d > 0.0)
                E[empty, empty] = 1.0
                bdense = np.ascontiguousarray(level_weight * np.linalg.inv(E))
            else:
                d = np.empty(k, dtype=np.float64)
                if L.oc_galerkin_diag(self.nn, _p(self.rowptr), _p(self.col), _p(self.val), _p(self.isdir), _p(node0), _p(t),
                                      _p(n), _p(d)) != 0:
                    raise MemoryError("oc_galerkin_diag failed")
                bdiag.append(level_weight * np.where(d > 0.0, 1.0 / np.where(d > 0.0, d, 1.0), 0.0))
            if l == 0:
                node0_f, t_f = node0, t
            ns.append(n)
        self.coarse = dict(nlev=nlev, n=np.ascontiguousarray(np.stack(ns), dtype=np.int32), node0=node0_f, t=t_f,
                           bdiag=np.ascontiguousarray(np.concatenate(bdiag)) if bdiag else np.zeros(1), bdense=bdense,
                           coarse_unknowns=int(bdense.shape[0]), level_weight=level_weight)
        return self.coarse

    def pcg_coarse(self, rtol=1e-12, maxit=100000, x0=None):
        """PCG with the Jacobi + coarse-grid preconditioner (call :meth:`coarse_setup` first)."""
        if self.coarse is None:
            self.coarse_setup()
        c = self.coarse
        x = np.zeros(self.nn, dtype=np.float64) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        rel = C.c_double()
        it = lib().oc_pcg_coarse(self.nn, _p(self.rowptr), _p(self.col), _p(self.val), _p(self.b), _p(x), float(rtol), int(maxit),
                                 C.byref(rel), int(c["nlev"]), _p(c["n"]), _p(c["node0"]), _p(c["t"]), _p(self.isdir),
                                 _p(c["bdiag"]), _p(c["bdense"]))
        if it < 0:
            raise MemoryError("oc_pcg_coarse failed")
        return x, it, rel.value

    def recover_lumped(self, phi):
        """Nodal ``volume current`` by the volume-weighted average (= fem_oracle ``method="lumped"``)."""
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        J = np.empty((self.nn, 3), dtype=np.float64)
        if lib().oc_recover_lumped(self.nn, _p(self.nodes), _p(self.tets), _p(self.sigma_e), _p(phi), _p(J)) != 0:
            raise RuntimeError("oc_recover_lumped needs the pattern of this mesh (build the CSystem last)")
        return J
