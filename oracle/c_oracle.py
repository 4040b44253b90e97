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
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def threads():
    return lib().oc_threads()


def set_threads(n):
    lib().oc_set_threads(int(n))


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
