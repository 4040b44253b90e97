"""CPU prototype (scipy) behind csrc/coarse.cu: PCG iteration counts of Jacobi vs Jacobi + trilinear coarse grids on the
bench workload.  Usage: python scripts/proto_coarse_space.py S|M|NXxNYxNZ   (SKIPJ=1 skips the plain Jacobi solve)."""
import sys, time, importlib, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import fem_oracle as fo
mg = importlib.import_module('pelvistim-fem_b200.meshgen')
import bench
size = sys.argv[1]
mesh = mg.synth_slab(tuple(int(v) for v in size.split('x')) if 'x' in size else size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8)
K = fo.assemble_stiffness(mesh.nodes, mesh.tets, mesh.region, bench.SIGMA).tocsr()
nn = mesh.nn
b = np.zeros(nn)
c = confs[3]
tr = mesh.tris[c['tris']]; p = mesh.nodes[tr]
ar = 0.5*np.linalg.norm(np.cross(p[:,1]-p[:,0],p[:,2]-p[:,0]),axis=1)
g = bench.I_INJECT/c['area']
for a in range(3): np.add.at(b, tr[:,a], g*ar/3)
isd = fo.dirichlet_nodes(mesh.tris, mesh.bcid, [(102,0.0)], nn)
if isinstance(isd, tuple): isd, val = isd
else: val = np.zeros(nn)
K, b = fo.apply_dirichlet_symmetric(K, b, isd, val)
K = K.tocsr()
print(size, 'nn', nn, 'nnz', K.nnz, 'ndir', isd.sum())
d = K.diagonal()

def pcg(K, b, M, rtol=1e-10, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r@z; bn = np.linalg.norm(b); it=0
    while it < maxit and np.linalg.norm(r) > rtol*bn:
        q = K@p; al = rz/(p@q); x += al*p; r -= al*q; z = M(r); rzn = r@z; p = z + (rzn/rz)*p; rz = rzn; it+=1
    return x, it
import os
if os.environ.get('SKIPJ'): x0=None; it0=0
else:
    t=time.time(); x0, it0 = pcg(K, b, lambda r: r/d); print('jacobi its', it0, time.time()-t)

free = ~isd
def coarse_trilinear(ncx, ncy, zlines):
    """Z: trilinear interpolation from a tensor coarse grid (x,y uniform, z given lines) to fine nodes."""
    X = mesh.nodes
    bx = np.linspace(0, 0.08, ncx+1); by = np.linspace(0, 0.06, ncy+1); bz = np.asarray(zlines)
    def w1(c, lines):
        i = np.clip(np.searchsorted(lines, c, side='right')-1, 0, len(lines)-2)
        t = (c-lines[i])/(lines[i+1]-lines[i]); t = np.clip(t,0,1)
        return i, t
    ix, tx = w1(X[:,0], bx); iy, ty = w1(X[:,1], by); iz, tz = w1(X[:,2], bz)
    nxn, nyn, nzn = len(bx), len(by), len(bz)
    rows=[];cols=[];vals=[]
    for dx in (0,1):
        for dy in (0,1):
            for dz in (0,1):
                w = (tx if dx else 1-tx)*(ty if dy else 1-ty)*(tz if dz else 1-tz)
                cid = (ix+dx) + nxn*((iy+dy) + nyn*(iz+dz))
                rows.append(np.arange(nn)); cols.append(cid); vals.append(w)
    Z = sp.csr_matrix((np.concatenate(vals),(np.concatenate(rows),np.concatenate(cols))),shape=(nn,nxn*nyn*nzn))
    Z = sp.diags(free.astype(float)) @ Z
    keep = np.asarray(Z.sum(axis=0)).ravel() > 1e-12
    return Z[:, keep].tocsr()
def coarse_const(ncx, ncy, zlines):
    X = mesh.nodes
    ix = np.minimum((X[:,0]/0.08*ncx).astype(int), ncx-1); iy = np.minimum((X[:,1]/0.06*ncy).astype(int), ncy-1)
    iz = np.clip(np.searchsorted(zlines, X[:,2], side='right')-1, 0, len(zlines)-2)
    cid = ix + ncx*(iy + ncy*iz)
    Z = sp.csr_matrix((free.astype(float),(np.arange(nn),cid)),shape=(nn,ncx*ncy*(len(zlines)-1)))
    keep = np.asarray(Z.sum(axis=0)).ravel() > 0
    return Z[:, keep].tocsr()

Lz = mesh.meta['Lz']; ts = mesh.meta['t_skin']; tf = 0.005
zl_layers = lambda nm: np.concatenate([np.linspace(0, Lz-ts-tf, nm+1), [Lz-ts, Lz]])
for name, Z in [
    ('tri 12x9x(3+2)', coarse_trilinear(12,9,zl_layers(3))),
    ('tri 24x18x(6+2)', coarse_trilinear(24,18,zl_layers(6))),
    ('tri 16x12x(4+2)', coarse_trilinear(16,12,zl_layers(4))),
    ('tri 32x24x(8+2)', coarse_trilinear(32,24,zl_layers(8))),
    ('tri 32x24 uniform z 10', coarse_trilinear(32,24,np.linspace(0,Lz,11))),
]:
    E = (Z.T @ K @ Z).tocsc(); lu = spla.splu(E)
    M = lambda r: r/d + Z @ lu.solve(Z.T @ r)
    t=time.time(); x1, it1 = pcg(K, b, M)
    print(f'{name}: k={Z.shape[1]} its {it1}    {time.time()-t:.1f}s', flush=True)
