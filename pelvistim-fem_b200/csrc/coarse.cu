// coarse.cu — geometric coarse spaces for the PCG preconditioner: grids, Galerkin operators, dense inverse,
// restriction / coarse solve kernels.  See coarse.cuh for the operator; solver.cu adds Z y_c to z in its p-update.
//
// Everything is deterministic: restriction sums per coarse cell in a fixed order, a grid node then adds the
// partials of its (up to) eight cells in a fixed order; the Galerkin matrix is built the same way.  The only
// atomics are the shared-memory accumulations inside one CTA of the Galerkin build (their order changes the
// preconditioner in the last bits, never the solution the iteration converges to) and the rarely taken slow path.
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include "coarse.cuh"

namespace cg = cooperative_groups;

namespace ptfem {
namespace {

constexpr int kGjNb = 32;     // pivot block of the Gauss-Jordan inverse
constexpr int kGjTile = 64;   // update tile (kp is a multiple of it)

// ---- grids ------------------------------------------------------------------------------------------------
// table row = position of the mesh node in the coarsest grid (see CoarseSpace::ctab)
__global__ void coarse_table_kernel(CoarseGrid g, const double* __restrict__ xyz, int64_t nn, float4* __restrict__ ctab) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  int c[3];
  double t[3];
  coarse_locate(g, xyz, i, c, t);
  ctab[i] = coarse_row_pack(c, t, false);
}
// Dirichlet rows carry the sign bit: coarse_row() then reports them as outside every coarse space
__global__ void coarse_table_flag_kernel(const uint8_t* __restrict__ isdir, int64_t nn, float4* __restrict__ ctab) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const unsigned int cell = __float_as_uint(ctab[i].w) & ~kCoarseDirBit;
  ctab[i].w = __uint_as_float(isdir[i] ? (cell | kCoarseDirBit) : cell);
}

// table rows in the order of a row list
__global__ void gather_table_kernel(const float4* __restrict__ ctab, const int32_t* __restrict__ rows, int64_t nn,
                                    float4* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nn) return;
  out[p] = ctab[rows[p]];
}

// (called before the Dirichlet flags are set: every row has a cell)
__global__ void cell_key_kernel(int n0, int n1, int shift, const float4* __restrict__ ctab, int64_t nn, int32_t* __restrict__ key,
                                int32_t* __restrict__ id) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  int c[3];
  double t[3];
  coarse_row(ctab, i, shift, c, t);
  key[i] = c[0] + n0 * (c[1] + n1 * c[2]);
  id[i] = (int32_t)i;
}

// cellptr[c] = first position of the sorted keys holding a key >= c
__global__ void cell_ptr_kernel(const int32_t* __restrict__ key, int64_t nn, int64_t ncell, int32_t* __restrict__ cellptr) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > nn) return;
  const int64_t hi = p < nn ? key[p] : ncell;
  const int64_t lo = p == 0 ? -1 : key[p - 1];
  for (int64_t c = lo + 1; c <= hi; ++c) cellptr[c] = (int32_t)p;
}

// last CTA of the grid (ticket) — same protocol as solver.cu
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// Sum of a task's 8 * NV per-lane accumulators over the row slots and store of the task's partials (shared by the restriction kernels).
template <int S>
__device__ __forceinline__ void restrict_reduce_store(const double (&acc)[8][S >= 2 ? 2 : 1], int lane, int pr, int64_t w,
                                                      double* __restrict__ part) {
  constexpr int NP = S >= 2 ? S / 2 : 1;
  constexpr int NV = S >= 2 ? 2 : 1;
  // Sum over the row slots (lanes with the same system pair).  Butterfly with halving: at every step a lane hands half of
  // its remaining sums to its partner and receives the partner's half of the others, so 8*NV sums over 32/NP slots cost
  // 8*NV - (what is left per lane) shuffles instead of 8*NV per step (14 instead of 48 for S = 8: the shuffles were
  // half of this kernel's LSU wavefronts), and the sums end up spread over the lanes, which then store in parallel.
  constexpr int NVAL = 8 * NV;
  double v[NVAL];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int q = 0; q < NV; ++q) v[a * NV + q] = acc[a][q];
  int base = 0;        // v[i] of this lane is sum number base + i
  bool owner = true;   // false for the duplicates left by steps taken after a lane is down to one sum
  {
    int n = NVAL;
#pragma unroll
    for (int o = NP; o < 32; o <<= 1) {
      const bool up = (lane & o) != 0;
      if (n > 1) {
#pragma unroll
        for (int i = 0; i < NVAL / 2; ++i) {
          if (i < n / 2) {
            const double send = up ? v[i] : v[i + n / 2];
            const double keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        base += up ? n / 2 : 0;
        n /= 2;
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        owner = owner && !up;
      }
    }
    if (owner) {
      if (n >= 2 && NV == 2) {   // pairs (sum 2j, 2j+1) = corner j, both systems of this lane: one 16-byte store each
#pragma unroll
        for (int i = 0; i < NVAL; i += 2)
          if (i < n) *reinterpret_cast<double2*>(part + ((size_t)w * 8 + (base + i) / 2) * S + 2 * pr) = make_double2(v[i], v[i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < NVAL; ++i)
          if (i < n) part[((size_t)w * 8 + (base + i) / NV) * S + NV * pr + (base + i) % NV] = v[i];
      }
    }
  }
}

// ---- restriction: per-cell partials -------------------------------------------------------------------------
// One warp per task = (cell, split index); lane = (row slot, system pair).  part[task][corner][s] = sum over the
// task's rows of w_corner(row) r[row][s]: registers and shuffles only, fixed order, no atomics.
// FUSE: the CG residual update rides along - r[row] -= alpha q[row] is applied to every row as it is gathered (each mesh row
// sits in exactly one task), written back, and r.D^-1 r / r.r are accumulated; the separate pass of cg_update_kernel over
// r, q and the inverse diagonal is gone (one launch and one read of r less per iteration).  fu.* are only read when FUSE.
struct FusedUpdate {
  const double* q = nullptr;        // [nn][S]
  const double* dinv = nullptr;     // [nn] (one shared matrix)
  const double* alpha = nullptr;    // [S] device scalars
  double* r = nullptr;              // [nn][S] updated in place
  double* partial = nullptr;        // [grid][2][S]
  unsigned int* ticket = nullptr;
  double* out_rz = nullptr;         // [S] r.D^-1 r (Jacobi part of rho)
  double* out_rr = nullptr;         // [S]
  int prefetch = 0;                 // the next trip's rows of r, q, dinv are asked into L2 as soon as their indices are known
};
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int S, int OCC, bool FUSE = false>
__global__ void __launch_bounds__(256, OCC) restrict_cell_kernel(int64_t ntask, int split, int shift, const int32_t* __restrict__ cellptr,
                                                            const int32_t* __restrict__ rows, const float4* __restrict__ ctab0,
                                                            const double* __restrict__ r, double* __restrict__ part, FusedUpdate fu) {
  constexpr int NP = S >= 2 ? S / 2 : 1;  // lanes per row
  constexpr int NV = S >= 2 ? 2 : 1;      // systems per lane
  constexpr int RPW = 32 / NP;            // rows per warp trip
  const int lane = threadIdx.x & 31;
  const int pr = lane % NP, slot = lane / NP;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double alpha[NV], dz[NV], dr[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    alpha[v] = FUSE ? fu.alpha[NV * pr + v] : 0.0;
    dz[v] = 0.0;
    dr[v] = 0.0;
  }
  for (int64_t w = warp0; w < ntask; w += nwarp) {
    const int64_t c = w / split;
    const int sp = (int)(w % split);
    double acc[8][NV];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[a][v] = 0.0;
    const int32_t p1 = cellptr[c + 1];
    // the row indices of the NEXT trip are loaded before this trip's values are used: one dependent latency (index -> value) less
    // per trip after the first (the kernel is latency-bound: ncu r03, SM 26 %, DRAM 41 %, occupancy 34 % with the update fused in)
    const int32_t pfirst = cellptr[c] + sp * RPW + slot;
    int32_t nia = pfirst < p1 ? __ldg(rows + pfirst) : -1, nib = pfirst + split * RPW < p1 ? __ldg(rows + pfirst + split * RPW) : -1;
    for (int32_t p = pfirst; p < p1; p += 2 * split * RPW) {
      // two rows per trip so that both index -> value chains are in flight together
      const int32_t pb = p + split * RPW;
      const int32_t ia = nia, ib = nib;
      {
        const int32_t pn = p + 2 * split * RPW, pnb = pn + split * RPW;
        nia = pn < p1 ? __ldg(rows + pn) : -1;
        nib = pnb < p1 ? __ldg(rows + pnb) : -1;
      }
      if constexpr (FUSE) {
        if (fu.prefetch && ia >= 0) {    // (the indices of this trip arrived one trip ago: these are the next trip's)
          if (nia >= 0) {
            prefetch_l2(fu.r + (size_t)nia * S + NV * pr);
            prefetch_l2(fu.q + (size_t)nia * S + NV * pr);
            if (pr == 0) prefetch_l2(fu.dinv + nia);
          }
          if (nib >= 0) {
            prefetch_l2(fu.r + (size_t)nib * S + NV * pr);
            prefetch_l2(fu.q + (size_t)nib * S + NV * pr);
            if (pr == 0) prefetch_l2(fu.dinv + nib);
          }
        }
      }
      const CoarseRaw ra = coarse_row_load(ctab0, p), rb = coarse_row_load(ctab0, pb < p1 ? pb : p);
      double va[NV], vb[NV];
      if constexpr (FUSE) {          // r is rewritten by this kernel: plain loads, through the writable pointer
        if constexpr (NV == 2) {
          const double2 x = *reinterpret_cast<const double2*>(fu.r + (size_t)ia * S + 2 * pr);
          va[0] = x.x;
          va[1] = x.y;
          const double2 y = ib >= 0 ? *reinterpret_cast<const double2*>(fu.r + (size_t)ib * S + 2 * pr) : make_double2(0.0, 0.0);
          vb[0] = y.x;
          vb[1] = y.y;
        } else {
          va[0] = fu.r[ia];
          vb[0] = ib >= 0 ? fu.r[ib] : 0.0;
        }
      } else if constexpr (NV == 2) {
        const double2 x = __ldg(reinterpret_cast<const double2*>(r + (size_t)ia * S + 2 * pr));
        va[0] = x.x;
        va[1] = x.y;
        const double2 y = ib >= 0 ? __ldg(reinterpret_cast<const double2*>(r + (size_t)ib * S + 2 * pr)) : make_double2(0.0, 0.0);
        vb[0] = y.x;
        vb[1] = y.y;
      } else {
        va[0] = __ldg(r + ia);
        vb[0] = ib >= 0 ? __ldg(r + ib) : 0.0;
      }
      if constexpr (FUSE) {
        double qa[NV], qb[NV];
        if constexpr (NV == 2) {
          const double2 x = __ldg(reinterpret_cast<const double2*>(fu.q + (size_t)ia * S + 2 * pr));
          qa[0] = x.x;
          qa[1] = x.y;
          const double2 y = ib >= 0 ? __ldg(reinterpret_cast<const double2*>(fu.q + (size_t)ib * S + 2 * pr)) : make_double2(0.0, 0.0);
          qb[0] = y.x;
          qb[1] = y.y;
        } else {
          qa[0] = __ldg(fu.q + ia);
          qb[0] = ib >= 0 ? __ldg(fu.q + ib) : 0.0;
        }
        const double da = __ldg(fu.dinv + ia), db = ib >= 0 ? __ldg(fu.dinv + ib) : 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          va[v] = fma(-alpha[v], qa[v], va[v]);
          vb[v] = fma(-alpha[v], qb[v], vb[v]);
          dz[v] = fma(va[v] * da, va[v], dz[v]);
          dr[v] = fma(va[v], va[v], dr[v]);
          dz[v] = fma(vb[v] * db, vb[v], dz[v]);
          dr[v] = fma(vb[v], vb[v], dr[v]);
        }
        if constexpr (NV == 2) {
          *reinterpret_cast<double2*>(fu.r + (size_t)ia * S + 2 * pr) = make_double2(va[0], va[1]);
          if (ib >= 0) *reinterpret_cast<double2*>(fu.r + (size_t)ib * S + 2 * pr) = make_double2(vb[0], vb[1]);
        } else {
          fu.r[ia] = va[0];
          if (ib >= 0) fu.r[ib] = vb[0];
        }
      }
      int cc[3];
      double t[3], wgt[8];
      {
        const double live = coarse_row_decode(ra, shift, cc, t) ? 1.0 : 0.0;
        coarse_weights(t, wgt);
#pragma unroll
        for (int v = 0; v < NV; ++v) va[v] *= live;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[a][v] = fma(wgt[a], va[v], acc[a][v]);
      }
      if (ib >= 0) {
        const double live = coarse_row_decode(rb, shift, cc, t) ? 1.0 : 0.0;
        coarse_weights(t, wgt);
#pragma unroll
        for (int v = 0; v < NV; ++v) vb[v] *= live;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[a][v] = fma(wgt[a], vb[v], acc[a][v]);
      }
    }
    if constexpr (FUSE) {
      // the next task's row list and table entries (consecutive in list order) are asked into L2 before the butterfly: its first
      // trip then waits for L2, not DRAM, on the index -> value chain
      if (fu.prefetch >= 2 && w + nwarp < ntask) {
        const int64_t c2 = (w + nwarp) / split;
        const int32_t q0 = __ldg(cellptr + c2) + lane;
        if (q0 < __ldg(cellptr + c2 + 1)) {     // (never past the lists)
          prefetch_l2(rows + q0);
          prefetch_l2(ctab0 + q0);
        }
      }
    }
    restrict_reduce_store<S>(acc, lane, pr, w, part);
  }
  if constexpr (FUSE) {
    // r.D^-1 r and r.r per system: slot 2*tid + v of the block belongs to system (2*tid + v) mod S (NV == 2) / system 0; fixed
    // order inside the block, the last block adds the blocks up in order
    __shared__ double s_red[2][NV * 256];
    const int tid = threadIdx.x;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      s_red[0][NV * tid + v] = dz[v];
      s_red[1][NV * tid + v] = dr[v];
    }
    __syncthreads();
    if (tid < 2 * S) {
      const int qn = tid / S, sys = tid % S;
      double t = 0.0;
      for (int k = sys; k < NV * 256; k += S) t += s_red[qn][k];
      fu.partial[(size_t)blockIdx.x * 2 * S + tid] = t;
    }
    if (last_block(fu.ticket)) {
      // thread t: slot t % 2S of blocks t / 2S, t / 2S + G, ...; then the G group sums in order
      constexpr int G = 256 / (2 * S);
      const int k = tid % (2 * S), g = tid / (2 * S);
      double t = 0.0;
      for (unsigned int b = g; b < gridDim.x; b += G) t += __ldcg(fu.partial + (size_t)b * 2 * S + k);
      __syncthreads();
      s_red[0][tid] = t;
      __syncthreads();
      if (tid < 2 * S) {
        double tot = 0.0;
        for (int gg = 0; gg < G; ++gg) tot += s_red[0][gg * 2 * S + tid];
        if (tid < S) fu.out_rz[tid] = tot; else fu.out_rr[tid - S] = tot;
      }
    }
  }
}

// ---- the fused update + restriction as a software pipeline (PTFEM_FUSE_PIPE) ---------------------------------------------
// The kernel above is latency-bound (ncu r03: 67 % long-scoreboard stalls, SM 26 %, DRAM 41 %, 24 warps per SM at 80 registers):
// a warp loads a trip's indices, waits, loads the rows, waits, computes, and only then starts the next trip.  Here every warp
// owns a CONTIGUOUS run of tasks (its cell pointers sit in shared memory, so where a later trip starts costs no global load),
// the row indices are fetched two trips ahead into registers, and the rows of r, q, the table entries and the inverse
// diagonal of the NEXT trip are on their way into a shared-memory stage (cp.async, one group per trip, two stages) while the
// current trip is decoded, accumulated and - at the end of a task - reduced and stored.  Same per-lane order of sums as above,
// so the partials are bit-identical; r.D^-1 r and r.r group their per-thread sums by the contiguous runs.
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int S>
struct PipeStage {   // one warp, one trip: two rows per lane group
  static constexpr int RPW = 32 / (S / 2);
  double2 ra[32], rb[32], qa[32], qb[32];
  float4 ta[RPW], tb[RPW];
  double da[RPW], db[RPW];
};
constexpr int kPipeTasks = 32;   // tasks per cached run of list ranges (two entries per task and warp)
template <int S>
constexpr size_t pipe_smem_bytes() { return 8 * (2 * sizeof(PipeStage<S>) + 64 * sizeof(int32_t)); }

template <int S>
__global__ void __launch_bounds__(256, 3) restrict_fused_pipe_kernel(int64_t ntask, int split, int shift, const int32_t* __restrict__ cellptr,
                                                                    const int32_t* __restrict__ rows, const float4* __restrict__ ctab0,
                                                                    double* __restrict__ part, FusedUpdate fu, int strided) {
  static_assert(S >= 2, "two systems per lane");
  constexpr int NP = S / 2, NV = 2, RPW = 32 / NP;
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pr = lane % NP, slot = lane / NP;
  PipeStage<S>* stage = reinterpret_cast<PipeStage<S>*>(pipe_smem) + 2 * wid;
  int32_t* cache = reinterpret_cast<int32_t*>(pipe_smem + 8 * 2 * sizeof(PipeStage<S>)) + 64 * wid;
  const int64_t gw = (int64_t)blockIdx.x * 8 + wid, nw = (int64_t)gridDim.x * 8;
  // the warp's k-th task: a contiguous run (strided == 0) or every nw-th task (neighbouring warps in neighbouring cells at a time)
  const int64_t n_mine = strided ? (ntask > gw ? (ntask - gw + nw - 1) / nw : 0) : ntask / nw + (gw < ntask % nw ? 1 : 0);
  const int64_t first = strided ? gw : ntask / nw * gw + min(gw, ntask % nw), step = strided ? nw : 1;
  const int64_t t_begin = 0, t_end = n_mine;
  const int stride = 2 * split * RPW, boff = split * RPW;
  double alpha[NV], dz[NV], dr[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    alpha[v] = fu.alpha[NV * pr + v];
    dz[v] = 0.0;
    dr[v] = 0.0;
  }
  for (int64_t run0 = t_begin; run0 < t_end; run0 += kPipeTasks) {
    const int tend = (int)min((int64_t)kPipeTasks, t_end - run0);   // tasks of this run, numbered from 0
    __syncwarp();
    if (lane < tend) {   // list range of the run's tasks: [cell start + split index * RPW, cell end)
      const int64_t w = first + (run0 + lane) * step, c = w / split;
      cache[2 * lane] = __ldg(cellptr + c) + (int)(w - c * split) * RPW;
      cache[2 * lane + 1] = __ldg(cellptr + c + 1);
    }
    __syncwarp();
    // a trip: task t, first list position `base` of its rows a (rows b follow at base + boff), end of the task's cell p1
    auto task_desc = [&](int t, int32_t& base, int32_t& p1) {
      base = cache[2 * t];
      p1 = cache[2 * t + 1];
    };
    auto advance = [&](int& t, int32_t& base, int32_t& p1) {
      base += stride;
      if (base >= p1 && ++t < tend) task_desc(t, base, p1);
    };
    auto load_idx = [&](int t, int32_t base, int32_t p1, int32_t& ia, int32_t& ib) {
      const int32_t pa = base + slot, pb = pa + boff;
      ia = (t < tend && pa < p1) ? __ldg(rows + pa) : -1;
      ib = (t < tend && pb < p1) ? __ldg(rows + pb) : -1;
    };
    auto issue = [&](PipeStage<S>& st, int32_t base, int32_t ia, int32_t ib) {
      if (ia >= 0) {
        cp_async16(&st.ra[lane], fu.r + (size_t)ia * S + 2 * pr);
        cp_async16(&st.qa[lane], fu.q + (size_t)ia * S + 2 * pr);
        if (pr == 0) {
          cp_async16(&st.ta[slot], ctab0 + base + slot);
          cp_async8(&st.da[slot], fu.dinv + ia);
        }
      }
      if (ib >= 0) {
        cp_async16(&st.rb[lane], fu.r + (size_t)ib * S + 2 * pr);
        cp_async16(&st.qb[lane], fu.q + (size_t)ib * S + 2 * pr);
        if (pr == 0) {
          cp_async16(&st.tb[slot], ctab0 + base + boff + slot);
          cp_async8(&st.db[slot], fu.dinv + ib);
        }
      }
      cp_async_commit();
    };
    int t0 = 0, t1, t2;
    int32_t b0, e0, b1, e1, b2, e2, ia0, ib0, ia1, ib1, ia2, ib2;
    task_desc(t0, b0, e0);
    load_idx(t0, b0, e0, ia0, ib0);
    t1 = t0; b1 = b0; e1 = e0;
    advance(t1, b1, e1);
    load_idx(t1, b1, e1, ia1, ib1);
    t2 = t1; b2 = b1; e2 = e1;
    if (t2 < tend) advance(t2, b2, e2);
    issue(stage[0], b0, ia0, ib0);
    int s = 0;
    double acc[8][NV];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[a][v] = 0.0;
    while (t0 < tend) {
      issue(stage[s ^ 1], b1, ia1, ib1);          // next trip's rows on their way (an empty group past the end)
      load_idx(t2, b2, e2, ia2, ib2);             // indices two trips ahead
      cp_async_wait<1>();                         // this trip's group has landed
      __syncwarp();                               // (table entry and inverse diagonal were fetched by the lane group's first lane)
      const PipeStage<S>& st = stage[s];
      int cc[3];
      double t[3], wgt[8];
      if (ia0 >= 0) {
        const double2 rv = st.ra[lane], qv = st.qa[lane];
        const double d = st.da[slot];
        CoarseRaw raw;
        raw.v = st.ta[slot];
        double va[NV];
        va[0] = fma(-alpha[0], qv.x, rv.x);
        va[1] = fma(-alpha[1], qv.y, rv.y);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          dz[v] = fma(va[v] * d, va[v], dz[v]);
          dr[v] = fma(va[v], va[v], dr[v]);
        }
        *reinterpret_cast<double2*>(fu.r + (size_t)ia0 * S + 2 * pr) = make_double2(va[0], va[1]);
        const double live = coarse_row_decode(raw, shift, cc, t) ? 1.0 : 0.0;
        coarse_weights(t, wgt);
#pragma unroll
        for (int v = 0; v < NV; ++v) va[v] *= live;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[a][v] = fma(wgt[a], va[v], acc[a][v]);
      }
      if (ib0 >= 0) {
        const double2 rv = st.rb[lane], qv = st.qb[lane];
        const double d = st.db[slot];
        CoarseRaw raw;
        raw.v = st.tb[slot];
        double vb[NV];
        vb[0] = fma(-alpha[0], qv.x, rv.x);
        vb[1] = fma(-alpha[1], qv.y, rv.y);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          dz[v] = fma(vb[v] * d, vb[v], dz[v]);
          dr[v] = fma(vb[v], vb[v], dr[v]);
        }
        *reinterpret_cast<double2*>(fu.r + (size_t)ib0 * S + 2 * pr) = make_double2(vb[0], vb[1]);
        const double live = coarse_row_decode(raw, shift, cc, t) ? 1.0 : 0.0;
        coarse_weights(t, wgt);
#pragma unroll
        for (int v = 0; v < NV; ++v) vb[v] *= live;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[a][v] = fma(wgt[a], vb[v], acc[a][v]);
      }
      if (b0 + stride >= e0) {                    // last trip of task t0 (a task without rows still stores its zeros)
        restrict_reduce_store<S>(acc, lane, pr, first + (run0 + t0) * step, part);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[a][v] = 0.0;
      }
      __syncwarp();                               // every lane is done with stage s before the next trip but one refills it
      t0 = t1; b0 = b1; e0 = e1; ia0 = ia1; ib0 = ib1;
      t1 = t2; b1 = b2; e1 = e2; ia1 = ia2; ib1 = ib2;
      if (t2 < tend) advance(t2, b2, e2);
      s ^= 1;
    }
    cp_async_wait<0>();
  }
  {
    // r.D^-1 r and r.r per system, as in restrict_cell_kernel<S, OCC, true>
    __shared__ double s_red[2][NV * 256];
    const int tid = threadIdx.x;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      s_red[0][NV * tid + v] = dz[v];
      s_red[1][NV * tid + v] = dr[v];
    }
    __syncthreads();
    if (tid < 2 * S) {
      const int qn = tid / S, sys = tid % S;
      double t = 0.0;
      for (int k = sys; k < NV * 256; k += S) t += s_red[qn][k];
      fu.partial[(size_t)blockIdx.x * 2 * S + tid] = t;
    }
    if (last_block(fu.ticket)) {
      constexpr int G = 256 / (2 * S);
      const int k = tid % (2 * S), g = tid / (2 * S);
      double t = 0.0;
      for (unsigned int b = g; b < gridDim.x; b += G) t += __ldcg(fu.partial + (size_t)b * 2 * S + k);
      __syncthreads();
      s_red[0][tid] = t;
      __syncthreads();
      if (tid < 2 * S) {
        double tot = 0.0;
        for (int gg = 0; gg < G; ++gg) tot += s_red[0][gg * 2 * S + tid];
        if (tid < S) fu.out_rz[tid] = tot; else fu.out_rr[tid - S] = tot;
      }
    }
  }
}

// CTA sum per system of one value per thread whose system is threadIdx.x % S; result -> dpart[blockIdx.x][S];
// last CTA adds the CTAs up in order -> cdot[S]
template <int S>
__device__ __forceinline__ void dot_by_sys(double v, double* s_buf /*[blockDim]*/, double* __restrict__ dpart,
                                           double* __restrict__ cdot, unsigned int* ticket) {
  s_buf[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x < S) {
    double tot = 0.0;
    for (int k = threadIdx.x; k < (int)blockDim.x; k += S) tot += s_buf[k];
    dpart[(size_t)blockIdx.x * S + threadIdx.x] = tot;
  }
  if (last_block(ticket)) {
    // thread t adds the partials of CTAs t/S, t/S + G, ... for system t % S; then G group sums in order
    const int G = (int)blockDim.x / S, sys = threadIdx.x % S, g = threadIdx.x / S;
    double tot = 0.0;
    for (unsigned int b = g; b < gridDim.x; b += G) tot += __ldcg(dpart + (size_t)b * S + sys);
    __syncthreads();
    s_buf[threadIdx.x] = tot;
    __syncthreads();
    if (threadIdx.x < S) {
      double t2 = 0.0;
      for (int gg = 0; gg < G; ++gg) t2 += s_buf[gg * S + threadIdx.x];
      cdot[threadIdx.x] = t2;
    }
  }
}

// grid node I gathers the partials of the cells around it (fixed order).  DIAG: y = r_c * binv, dot.
template <int S, bool DIAG>
__global__ void __launch_bounds__(256) coarse_node_kernel(CoarseGrid g, int64_t k, int split, const double* __restrict__ part,
                                                          const double* __restrict__ binv, int64_t bstride, double* __restrict__ rc,
                                                          double* __restrict__ yc, double* __restrict__ dpart,
                                                          double* __restrict__ cdot, unsigned int* ticket) {
  __shared__ double s_buf[256];
  const int s = threadIdx.x % S;  // the stride is a multiple of S: a thread stays with one system
  double dot = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < k * S; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = e / S;
    double v = 0.0;
    const int nx1 = g.n[0] + 1, ny1 = g.n[1] + 1;
    const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int cx = ix - (a & 1), cy = iy - ((a >> 1) & 1), cz = iz - (a >> 2);
      if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
      const int64_t c = cx + (int64_t)g.n[0] * (cy + (int64_t)g.n[1] * cz);
      for (int sp = 0; sp < split; ++sp) v += __ldcg(part + (((size_t)c * split + sp) * 8 + a) * S + s);
    }
    rc[I * S + s] = v;
    if (DIAG) {
      const double y = v * __ldg(binv + (size_t)s * bstride + I);   // bstride > 0: every system has its own operator
      yc[I * S + s] = y;
      dot = fma(v, y, dot);
    }
  }
  if (DIAG) dot_by_sys<S>(dot, s_buf, dpart, cdot, ticket);
}

// ---- grid-to-grid transfers (nested grids, cells halve from level l to l-1) -------------------------------------
// Trilinear interpolation from the coarser grid reproduces itself on the finer one, so Z_coarse = Z_fine P: the mesh
// is touched once (finest level) and the other levels are reached through these small kernels.
// r_c[I] = sum_f P[f][I] r_f[f] (27 fine nodes around 2I);  DIAG: y = binv r_c and the level's dot
template <int S, bool DIAG>
__global__ void __launch_bounds__(256) grid_restrict_kernel(CoarseGrid gc, int64_t k, const double* __restrict__ rf,
                                                            const double* __restrict__ binv, int64_t bstride, double* __restrict__ rc,
                                                            double* __restrict__ yc, double* __restrict__ dpart,
                                                            double* __restrict__ cdot, unsigned int* ticket) {
  __shared__ double s_buf[256];
  const int s = threadIdx.x % S;
  double dot = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < k * S; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = e / S;
    double v = 0.0;
    const int nx1 = gc.n[0] + 1, ny1 = gc.n[1] + 1;
    const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
    const int fx1 = 2 * gc.n[0] + 1, fy1 = 2 * gc.n[1] + 1, fz1 = 2 * gc.n[2] + 1;
    for (int dz = -1; dz <= 1; ++dz) {
      const int fz = 2 * iz + dz;
      if (fz < 0 || fz >= fz1) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int fy = 2 * iy + dy;
        if (fy < 0 || fy >= fy1) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int fx = 2 * ix + dx;
          if (fx < 0 || fx >= fx1) continue;
          const double w = (dx ? 0.5 : 1.0) * (dy ? 0.5 : 1.0) * (dz ? 0.5 : 1.0);
          v = fma(w, __ldcg(rf + ((size_t)fx + (size_t)fx1 * (fy + (size_t)fy1 * fz)) * S + s), v);
        }
      }
    }
    rc[I * S + s] = v;
    if (DIAG) {
      const double y = v * __ldg(binv + (size_t)s * bstride + I);
      yc[I * S + s] = y;
      dot = fma(v, y, dot);
    }
  }
  if (DIAG) dot_by_sys<S>(dot, s_buf, dpart, cdot, ticket);
}

// yt_f[f] = y_f[f] + (P yt_c)[f]
template <int S>
__global__ void __launch_bounds__(256) grid_prolong_kernel(CoarseGrid gc, int64_t kf, const double* __restrict__ yf,
                                                           const double* __restrict__ ytc, double* __restrict__ ytf) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t F = e / S;
  const int s = (int)(e % S);
  if (F >= kf) return;
  const int fx1 = 2 * gc.n[0] + 1, fy1 = 2 * gc.n[1] + 1;
  const int nx1 = gc.n[0] + 1, ny1 = gc.n[1] + 1;
  const int fx = (int)(F % fx1), fy = (int)((F / fx1) % fy1), fz = (int)(F / ((int64_t)fx1 * fy1));
  double v = yf[F * S + s];
  for (int az = 0; az <= (fz & 1); ++az)
    for (int ay = 0; ay <= (fy & 1); ++ay)
      for (int ax = 0; ax <= (fx & 1); ++ax) {
        const double w = ((fx & 1) ? 0.5 : 1.0) * ((fy & 1) ? 0.5 : 1.0) * ((fz & 1) ? 0.5 : 1.0);
        const size_t c = (size_t)(fx / 2 + ax) + (size_t)nx1 * ((fy / 2 + ay) + (size_t)ny1 * (fz / 2 + az));
        v = fma(w, __ldcg(ytc + c * S + s), v);
      }
  ytf[F * S + s] = v;
}

// exact level: y_c = B r_c (dense), dot r_c . y_c.  CTA = 4 rows of B; warp w takes the w-th eighth of the columns;
// lane = (column within a group of 32/S, system), so the r_c loads of a warp are one contiguous 256-byte run.
constexpr int kDenseRows = 4;
template <int S>
__global__ void __launch_bounds__(256) coarse_dense_kernel(int kp, const double* __restrict__ binv0, int64_t bstride, const double* __restrict__ rc,
                                                           double* __restrict__ yc, double* __restrict__ dpart,
                                                           double* __restrict__ cdot, unsigned int* ticket) {
  __shared__ double s_buf[256];
  __shared__ double s_part[8][kDenseRows][S];
  constexpr int JPW = 32 / S;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int jj = lane / S, s = lane % S;
  const double* binv = binv0 + (size_t)s * bstride;     // bstride > 0: the inverse of this lane's system
  const int I0 = blockIdx.x * kDenseRows;
  const int chunk = kp / 8;  // kp is a multiple of 64
  const int j0 = wid * chunk, j1 = j0 + chunk;
  double acc[kDenseRows];
#pragma unroll
  for (int r = 0; r < kDenseRows; ++r) acc[r] = 0.0;
#pragma unroll 4
  for (int J = j0 + jj; J < j1; J += JPW) {
    const double v = __ldcg(rc + (size_t)J * S + s);
#pragma unroll
    for (int r = 0; r < kDenseRows; ++r) acc[r] = fma(__ldg(binv + (size_t)(I0 + r) * kp + J), v, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kDenseRows; ++r) {
#pragma unroll
    for (int o = S; o < 32; o <<= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  }
  if (lane < S) {
#pragma unroll
    for (int r = 0; r < kDenseRows; ++r) s_part[wid][r][lane] = acc[r];
  }
  __syncthreads();
  double dot = 0.0;
  if (threadIdx.x < kDenseRows * S) {
    const int r = threadIdx.x / S, ss = threadIdx.x % S;
    double y = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) y += s_part[w][r][ss];
    yc[(size_t)(I0 + r) * S + ss] = y;
    dot = y * __ldcg(rc + (size_t)(I0 + r) * S + ss);
  }
  dot_by_sys<S>(dot, s_buf, dpart, cdot, ticket);
}

// ---- the grid hierarchy in ONE cooperative launch (PTFEM_COARSE_FUSED) ---------------------------------------------
// node gather, grid restrictions, dense coarsest solve, prolongations and the per-level dots are six to eight small
// dependent kernels per CG iteration; between them sit launch gaps and the ticket / last-CTA epilogues of the dots.
// Here they are phases of one kernel separated by grid-wide barriers (cooperative groups).  Per element the arithmetic
// is that of the separate kernels; the dots are summed per CTA and then over the CTAs in index order by CTA 0
// (deterministic, but a different grouping than the separate kernels': the results agree to round-off, not bit for bit).
struct ChainLevel {
  CoarseGrid g;
  int64_t k;
  int kp, exact;
  const double* binv;
  double *rc, *yc, *yt;
};
struct ChainArgs {
  int nlev, split, do_node, scaled0;
  ChainLevel lev[kMaxCoarseLevels];
  const double* part;   // restriction partials of level 0 (do_node)
  double* dpart;        // [nlev][gridDim.x][S]
  double* cdot;         // [nlev][16]
};

template <int S>
__device__ __forceinline__ double chain_node_gather(const CoarseGrid& g, int split, const double* __restrict__ part, int64_t I, int s) {
  double v = 0.0;
  const int nx1 = g.n[0] + 1, ny1 = g.n[1] + 1;
  const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int cx = ix - (a & 1), cy = iy - ((a >> 1) & 1), cz = iz - (a >> 2);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
    const int64_t c = cx + (int64_t)g.n[0] * (cy + (int64_t)g.n[1] * cz);
    for (int sp = 0; sp < split; ++sp) v += __ldcg(part + (((size_t)c * split + sp) * 8 + a) * S + s);
  }
  return v;
}
template <int S>
__device__ __forceinline__ double chain_restrict27(const CoarseGrid& gc, const double* __restrict__ rf, int64_t I, int s) {
  double v = 0.0;
  const int nx1 = gc.n[0] + 1, ny1 = gc.n[1] + 1;
  const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
  const int fx1 = 2 * gc.n[0] + 1, fy1 = 2 * gc.n[1] + 1, fz1 = 2 * gc.n[2] + 1;
  for (int dz = -1; dz <= 1; ++dz) {
    const int fz = 2 * iz + dz;
    if (fz < 0 || fz >= fz1) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int fy = 2 * iy + dy;
      if (fy < 0 || fy >= fy1) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int fx = 2 * ix + dx;
        if (fx < 0 || fx >= fx1) continue;
        const double w = (dx ? 0.5 : 1.0) * (dy ? 0.5 : 1.0) * (dz ? 0.5 : 1.0);
        v = fma(w, __ldcg(rf + ((size_t)fx + (size_t)fx1 * (fy + (size_t)fy1 * fz)) * S + s), v);
      }
    }
  }
  return v;
}
template <int S>
__device__ __forceinline__ double chain_prolong8(const CoarseGrid& gc, const double* __restrict__ ytc, double yf, int64_t F, int s) {
  const int fx1 = 2 * gc.n[0] + 1, fy1 = 2 * gc.n[1] + 1;
  const int nx1 = gc.n[0] + 1, ny1 = gc.n[1] + 1;
  const int fx = (int)(F % fx1), fy = (int)((F / fx1) % fy1), fz = (int)(F / ((int64_t)fx1 * fy1));
  double v = yf;
  for (int az = 0; az <= (fz & 1); ++az)
    for (int ay = 0; ay <= (fy & 1); ++ay)
      for (int ax = 0; ax <= (fx & 1); ++ax) {
        const double w = ((fx & 1) ? 0.5 : 1.0) * ((fy & 1) ? 0.5 : 1.0) * ((fz & 1) ? 0.5 : 1.0);
        const size_t c = (size_t)(fx / 2 + ax) + (size_t)nx1 * ((fy / 2 + ay) + (size_t)ny1 * (fz / 2 + az));
        v = fma(w, __ldcg(ytc + c * S + s), v);
      }
  return v;
}

template <int S>
__global__ void __launch_bounds__(256) coarse_chain_kernel(ChainArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_buf[256];
  __shared__ double s_part[8][kDenseRows][S];
  const int s = threadIdx.x % S;   // 256 and the grid stride are multiples of S: a thread stays with one system
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  double dot[kMaxCoarseLevels];
#pragma unroll
  for (int l = 0; l < kMaxCoarseLevels; ++l) dot[l] = 0.0;

  // level 0: r_c from the restriction partials (or already there, summed over the ranks by the caller)
  if (a.do_node || (!a.scaled0 && !a.lev[0].exact)) {
    const ChainLevel& L = a.lev[0];
    for (int64_t e = tid; e < L.k * S; e += nthr) {
      const int64_t I = e / S;
      const double v = a.do_node ? chain_node_gather<S>(L.g, a.split, a.part, I, s) : L.rc[e];
      if (a.do_node) L.rc[e] = v;
      if (!L.exact) {
        const double y = v * __ldg(L.binv + I);
        L.yc[e] = y;
        dot[0] = fma(v, y, dot[0]);
      }
    }
    grid.sync();
  }
  // coarser levels: r_c by 27-point restriction of the next finer grid
#pragma unroll 1
  for (int l = 1; l < a.nlev; ++l) {
    const ChainLevel& L = a.lev[l];
    const double* rf = a.lev[l - 1].rc;
    double dl = 0.0;
    for (int64_t e = tid; e < L.k * S; e += nthr) {
      const int64_t I = e / S;
      const double v = chain_restrict27<S>(L.g, rf, I, s);
      L.rc[e] = v;
      if (!L.exact) {
        const double y = v * __ldg(L.binv + I);
        L.yc[e] = y;
        dl = fma(v, y, dl);
      }
    }
    dot[l] = dl;
    grid.sync();
  }
  // coarsest level: y_c = B r_c (dense); CTA takes groups of kDenseRows rows, warp w the w-th eighth of the columns
  {
    const ChainLevel& L = a.lev[a.nlev - 1];
    constexpr int JPW = 32 / S;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int jj = lane / S, ls = lane % S;
    const int chunk = L.kp / 8;
    const int j0 = wid * chunk, j1 = j0 + chunk;
    double dx = 0.0;
    for (int I0 = blockIdx.x * kDenseRows; I0 < L.kp; I0 += gridDim.x * kDenseRows) {
      double acc[kDenseRows];
#pragma unroll
      for (int r = 0; r < kDenseRows; ++r) acc[r] = 0.0;
#pragma unroll 4
      for (int J = j0 + jj; J < j1; J += JPW) {
        const double v = __ldcg(L.rc + (size_t)J * S + ls);
#pragma unroll
        for (int r = 0; r < kDenseRows; ++r) acc[r] = fma(__ldg(L.binv + (size_t)(I0 + r) * L.kp + J), v, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kDenseRows; ++r) {
#pragma unroll
        for (int o = S; o < 32; o <<= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
      }
      if (lane < S) {
#pragma unroll
        for (int r = 0; r < kDenseRows; ++r) s_part[wid][r][lane] = acc[r];
      }
      __syncthreads();
      if (threadIdx.x < kDenseRows * S) {
        const int r = threadIdx.x / S, ss = threadIdx.x % S;
        double y = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) y += s_part[w][r][ss];
        L.yc[(size_t)(I0 + r) * S + ss] = y;
        dx = fma(y, __ldcg(L.rc + (size_t)(I0 + r) * S + ss), dx);
      }
      __syncthreads();
    }
    dot[a.nlev - 1] = dx;
  }
  // per-level dots of this CTA (by system) -> dpart
#pragma unroll 1
  for (int l = 0; l < a.nlev; ++l) {
    __syncthreads();
    s_buf[threadIdx.x] = dot[l];
    __syncthreads();
    if (threadIdx.x < S) {
      double tot = 0.0;
      for (int k = threadIdx.x; k < (int)blockDim.x; k += S) tot += s_buf[k];
      a.dpart[((size_t)l * gridDim.x + blockIdx.x) * S + threadIdx.x] = tot;
    }
  }
  grid.sync();
  // CTA 0: the sums over the CTAs, in order (thread = (CTA group g, system); then the groups in order)
  if (blockIdx.x == 0) {
    const int G = (int)blockDim.x / S, sys = threadIdx.x % S, g = threadIdx.x / S;
#pragma unroll 1
    for (int l = 0; l < a.nlev; ++l) {
      double tot = 0.0;
      for (unsigned int b = g; b < gridDim.x; b += G) tot += __ldcg(a.dpart + ((size_t)l * gridDim.x + b) * S + sys);
      __syncthreads();
      s_buf[threadIdx.x] = tot;
      __syncthreads();
      if (threadIdx.x < S) {
        double t2 = 0.0;
        for (int gg = 0; gg < G; ++gg) t2 += s_buf[gg * S + threadIdx.x];
        a.cdot[l * 16 + threadIdx.x] = t2;
      }
    }
  }
  // coarsest -> finest grid: yt_l = y_l + P yt_{l+1}
#pragma unroll 1
  for (int l = a.nlev - 2; l >= 0; --l) {
    const ChainLevel& L = a.lev[l];
    const double* ytc = (l + 1 == a.nlev - 1) ? a.lev[l + 1].yc : a.lev[l + 1].yt;
    for (int64_t e = tid; e < L.k * S; e += nthr) L.yt[e] = chain_prolong8<S>(a.lev[l + 1].g, ytc, L.yc[e], e / S, s);
    if (l > 0) grid.sync();
  }
}

// ---- Galerkin operator of the exact level ----------------------------------------------------------------------
// CTA per cell.  For the rows i of the cell: Y[i][slot] = sum_j K_ij w_j(slot), slot = position of the grid node in
// the 4x4x4 block of nodes around the cell (columns in the 27 neighbouring cells); then the cell's 8 x 64 block
// sum_i w_i(I) Y[i][slot].  Columns further away (mesh edge longer than a coarse cell) take the slow path.
constexpr int kGalRows = 64;
__global__ void __launch_bounds__(256) galerkin_cell_kernel(CoarseGrid g, const int32_t* __restrict__ cellptr,
                                                            const int32_t* __restrict__ rows, const float4* __restrict__ ctab,
                                                            const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ col, const double* __restrict__ val,
                                                            double* __restrict__ blockE /*[ncell][gsplit][8][64]*/,
                                                            double* __restrict__ E, int kp, int32_t* __restrict__ flag, int gsplit) {
  __shared__ double Y[kGalRows][65];
  __shared__ double W[kGalRows][8];
  // gsplit CTAs share a cell (contiguous pieces of its row list, whole chunks of kGalRows rows): the exactly inverted grid
  // is small (a few hundred cells), one CTA per cell would leave most of the GPU idle
  const int64_t c = blockIdx.x / gsplit;
  const int sp = (int)(blockIdx.x % gsplit);
  int cc[3];
  cc[0] = (int)(c % g.n[0]);
  cc[1] = (int)((c / g.n[0]) % g.n[1]);
  cc[2] = (int)(c / ((int64_t)g.n[0] * g.n[1]));
  double acc[2] = {0.0, 0.0};
  const int rloc = threadIdx.x >> 2, l4 = threadIdx.x & 3;
  int32_t p0 = cellptr[c], p1 = cellptr[c + 1];
  {
    const int32_t chunks = (p1 - p0 + kGalRows - 1) / kGalRows, per = (chunks + gsplit - 1) / gsplit;
    const int32_t q0 = p0 + sp * per * kGalRows, q1 = q0 + per * kGalRows;
    p0 = q0 < p1 ? q0 : p1;
    p1 = q1 < p1 ? q1 : p1;
  }
  for (int32_t base = p0; base < p1; base += kGalRows) {
    for (int k = threadIdx.x; k < kGalRows * 65; k += 256) (&Y[0][0])[k] = 0.0;
    __syncthreads();
    const int32_t p = base + rloc;
    bool live = false;
    int32_t i = 0;
    double ti[3] = {0, 0, 0};
    if (p < p1) {
      i = rows[p];
      int ci[3];
      live = coarse_row(ctab, i, 0, ci, ti);
    }
    if (l4 == 0) {
#pragma unroll
      for (int a = 0; a < 8; ++a) W[rloc][a] = live ? coarse_weight(ti, a) : 0.0;
    }
    // The four lanes of a row all walk the whole row, and lane z adds only the corners that fall into plane z of the
    // 4x4x4 slot block: no two lanes ever touch the same slot, so plain read-modify-writes do (four lanes adding into
    // the same 64 slots needed shared-memory atomics - 400 M of them, the whole cost of the first version).  Non-zeros
    // are taken four at a time, all loads first: the col -> table chain is two dependent gathers deep.
    if (live) {
      const int32_t rb = rowptr[i], re = rowptr[i + 1];
      for (int32_t e0 = rb; e0 < re; e0 += 4) {
        double v4[4];
        CoarseRaw r4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int32_t e = e0 + u < re ? e0 + u : re - 1;
          v4[u] = e0 + u < re ? val[e] : 0.0;
          r4[u] = coarse_row_load(ctab, col[e]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double v = v4[u];
          if (v == 0.0) continue;
          int cj[3];
          double tj[3];
          if (!coarse_row_decode(r4[u], 0, cj, tj)) continue;
          const int d0 = cj[0] - cc[0], d1 = cj[1] - cc[1], d2 = cj[2] - cc[2];
          if (d0 >= -1 && d0 <= 1 && d1 >= -1 && d1 <= 1 && d2 >= -1 && d2 <= 1) {
            const int az = l4 - d2 - 1;   // corner layer of column j's cell that lies in this lane's plane
            if (az == 0 || az == 1) {
              const double wz = v * (az ? tj[2] : 1.0 - tj[2]);
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const int slot = (d0 + (a & 1) + 1) + 4 * (d1 + (a >> 1) + 1) + 16 * l4;
                Y[rloc][slot] += wz * ((a & 1) ? tj[0] : 1.0 - tj[0]) * ((a & 2) ? tj[1] : 1.0 - tj[1]);
              }
            }
          } else if (l4 == 0) {
            flag[1] = 1;
            int ci[3] = {cc[0], cc[1], cc[2]};
            for (int a = 0; a < 8; ++a) {
              const double wi = coarse_weight(ti, a);
              const int64_t I = coarse_node(g, ci, a);
              for (int b2 = 0; b2 < 8; ++b2) atomicAdd(E + (size_t)I * kp + coarse_node(g, cj, b2), wi * v * coarse_weight(tj, b2));
            }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = threadIdx.x + 256 * h, I = o >> 6, slot = o & 63;
      double t = acc[h];
      for (int rr = 0; rr < kGalRows; ++rr) t = fma(W[rr][I], Y[rr][slot], t);
      acc[h] = t;
    }
    __syncthreads();
  }
  blockE[(size_t)blockIdx.x * 512 + threadIdx.x] = acc[0];
  blockE[(size_t)blockIdx.x * 512 + 256 + threadIdx.x] = acc[1];
}

// E[I][J] += sum over the cells around I of their block entry for (I, J); padding / empty nodes -> identity
__global__ void __launch_bounds__(256) galerkin_gather_kernel(CoarseGrid g, int64_t k, int kp, const double* __restrict__ blockE,
                                                              double* __restrict__ E, int gsplit, int fix) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)kp * kp) return;
  const int64_t I = e / kp, J = e % kp;
  if (I >= k || J >= k) {
    E[e] = (I == J && fix) ? 1.0 : 0.0;
    return;
  }
  const int nx1 = g.n[0] + 1, ny1 = g.n[1] + 1;
  const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
  const int jx = (int)(J % nx1), jy = (int)((J / nx1) % ny1), jz = (int)(J / ((int64_t)nx1 * ny1));
  double v = E[e];
  if (abs(jx - ix) <= 2 && abs(jy - iy) <= 2 && abs(jz - iz) <= 2) {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int cx = ix - (a & 1), cy = iy - ((a >> 1) & 1), cz = iz - (a >> 2);
      if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
      const int sx = jx - cx + 1, sy = jy - cy + 1, sz = jz - cz + 1;
      if (sx < 0 || sx > 3 || sy < 0 || sy > 3 || sz < 0 || sz > 3) continue;
      const int64_t c = cx + (int64_t)g.n[0] * (cy + (int64_t)g.n[1] * cz);
      for (int sp = 0; sp < gsplit; ++sp) v += blockE[((size_t)c * gsplit + sp) * 512 + a * 64 + (sx + 4 * sy + 16 * sz)];
    }
  }
  if (fix && I == J && !(v > 0.0)) v = 1.0;  // grid node without a free mesh node in its support
  E[e] = v;
}
// the same fix applied to raw sums added up over the ranks (distributed set-up)
__global__ void galerkin_fix_kernel(int64_t k, int kp, double* __restrict__ E) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= kp) return;
  const double v = E[(size_t)I * kp + I];
  if (I >= k || !(v > 0.0)) E[(size_t)I * kp + I] = 1.0;
}
__global__ void diag_invert_kernel(int64_t k, double w, double* __restrict__ binv) {
  const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (I < k) binv[I] = binv[I] > 0.0 ? w / binv[I] : 0.0;
}

// ---- Galerkin diagonals of the finer (BPX) levels ---------------------------------------------------------------
// All diagonal-only levels in ONE pass over the matrix: CTAs walk the cells of the FINEST grid (its row list); for every
// row the levels are taken one after the other - the row's val / col / table entries are read from DRAM for the first
// level and from L1 for the others.  dpartf[l][finest cell][8] = this finest cell's contribution to the eight corners of
// the level-l cell that contains it; galerkin_diag_node_multi_kernel adds the children up.
__global__ void __launch_bounds__(256) galerkin_diag_multi_kernel(int nb, int shift0, int64_t ncell, const int32_t* __restrict__ cellptr,
                                                                  const int32_t* __restrict__ rows, const float4* __restrict__ ctab,
                                                                  const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                  const double* __restrict__ val, double* __restrict__ dpartf) {
  // one WARP per finest cell (a few dozen rows): lane = (row slot, quarter of the row's non-zeros); the eight corner sums
  // are reduced with shuffles - no shared memory, no block barrier (the block-wide reduction of the first version cost
  // more than the arithmetic)
  const int lane = threadIdx.x & 31, slot = lane >> 2, q4 = lane & 3;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp0; c < ncell; c += nwarp) {
    const int32_t p0 = cellptr[c], p1 = cellptr[c + 1];
    for (int l = 0; l < nb; ++l) {
      const int shift = shift0 - l;   // level l's cells are 2^l larger than the finest's
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int32_t p = p0 + slot; p < p1; p += 8) {
        const int32_t i = rows[p];
        int ci[3];
        double ti[3];
        if (!coarse_row(ctab, i, shift, ci, ti)) continue;
        double h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int32_t re = rowptr[i + 1];
        for (int32_t e0 = rowptr[i] + q4; e0 < re; e0 += 16) {   // four non-zeros of this lane at a time, all loads first
          double v4[4];
          CoarseRaw r4[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int32_t e = e0 + 4 * t;
            v4[t] = e < re ? val[e] : 0.0;
            r4[t] = coarse_row_load(ctab, col[e < re ? e : re - 1]);
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const double v = v4[t];
            if (v == 0.0) continue;
            int cj[3];
            double u[3];
            if (!coarse_row_decode(r4[t], shift, cj, u)) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) u[d] += (double)cj[d];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              const double hx = 1.0 - fabs(u[0] - (double)(ci[0] + (a & 1)));
              const double hy = 1.0 - fabs(u[1] - (double)(ci[1] + ((a >> 1) & 1)));
              const double hz = 1.0 - fabs(u[2] - (double)(ci[2] + (a >> 2)));
              if (hx > 0.0 && hy > 0.0 && hz > 0.0) h[a] = fma(v, hx * hy * hz, h[a]);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) acc[a] = fma(coarse_weight(ti, a), h[a], acc[a]);
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) acc[a] = warp_sum(acc[a]);
      if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a) dpartf[((size_t)l * ncell + c) * 8 + a] = acc[a];
      }
    }
  }
}
// one warp per grid node of level l (cells 2^l = f times the finest's): binv[I] = 1 / sum over the (up to) eight cells around
// I of the contributions of their f^3 finest children (lanes stride over the children, fixed shuffle tree: deterministic)
__global__ void galerkin_diag_node_multi_kernel(CoarseGrid g, int64_t k, int f, int nf0, int nf1, const double* __restrict__ dpart_l,
                                                double* __restrict__ binv, int invert) {
  const int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (I >= k) return;
  const int nx1 = g.n[0] + 1, ny1 = g.n[1] + 1;
  const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
  double v = 0.0;
  const int nch = f * f * f;
  for (int a = 0; a < 8; ++a) {
    const int cx = ix - (a & 1), cy = iy - ((a >> 1) & 1), cz = iz - (a >> 2);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
    for (int q = lane; q < nch; q += 32) {
      const int qx = q % f, qy = (q / f) % f, qz = q / (f * f);
      const int64_t cf = (int64_t)(cx * f + qx) + (int64_t)nf0 * ((cy * f + qy) + (int64_t)nf1 * (cz * f + qz));
      v += dpart_l[(size_t)cf * 8 + a];
    }
  }
  v = warp_sum(v);
  if (lane == 0) binv[I] = invert ? (v > 0.0 ? 1.0 / v : 0.0) : v;
}

// ---- dense inverse: blocked Gauss-Jordan without pivoting (the matrix is SPD) ---------------------------------
// step t: P = A[t,t]; R = P^-1 [A[t,:] with block column t replaced by I]; C = A[:,t];
//         A[i,:] = A[i,:](block column t zeroed) - C[i] R  (i != t),  A[t,:] = R
__global__ void gj_diag_kernel(const double* __restrict__ A, int kp, double* __restrict__ d0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kp) d0[i] = A[(size_t)i * kp + i];
}
// a pivot that has lost ten digits against the original diagonal means (numerically) dependent coarse functions
__global__ void __launch_bounds__(1024) gj_pivot_kernel(const double* __restrict__ A, int kp, int t, const double* __restrict__ d0,
                                                        double* __restrict__ Pinv, int32_t* __restrict__ flag) {
  __shared__ double P[kGjNb][kGjNb + 1], Q[kGjNb][kGjNb + 1];
  const int r = threadIdx.x / kGjNb, c = threadIdx.x % kGjNb;
  P[r][c] = A[(size_t)(t * kGjNb + r) * kp + t * kGjNb + c];
  Q[r][c] = r == c ? 1.0 : 0.0;
  __syncthreads();
  for (int s = 0; s < kGjNb; ++s) {
    const double piv = P[s][s], f = P[r][s];
    if (threadIdx.x == 0 && !(piv > 1e-10 * d0[t * kGjNb + s])) flag[0] = 1;
    __syncthreads();
    if (r == s) {
      P[r][c] /= piv;
      Q[r][c] /= piv;
    }
    __syncthreads();
    if (r != s) {
      P[r][c] -= f * P[s][c];
      Q[r][c] -= f * Q[s][c];
    }
    __syncthreads();
  }
  Pinv[r * kGjNb + c] = Q[r][c];
}

__global__ void __launch_bounds__(256) gj_panel_kernel(const double* __restrict__ A, int kp, int t, const double* __restrict__ Pinv,
                                                       double* __restrict__ Rbuf /*[nb][kp]*/, double* __restrict__ Cbuf /*[kp][nb]*/) {
  __shared__ double Ps[kGjNb][kGjNb + 1];
  __shared__ double As[kGjNb][kGjTile];
  const int j0 = blockIdx.x * kGjTile;
  for (int k = threadIdx.x; k < kGjNb * kGjNb; k += 256) Ps[k / kGjNb][k % kGjNb] = Pinv[k];
  for (int k = threadIdx.x; k < kGjNb * kGjTile; k += 256) {
    const int c = k / kGjTile, j = j0 + k % kGjTile;
    const int jt = j - t * kGjNb;
    As[c][k % kGjTile] = (jt >= 0 && jt < kGjNb) ? (jt == c ? 1.0 : 0.0) : A[(size_t)(t * kGjNb + c) * kp + j];
  }
  // column panel copy: rows j0 .. j0+63
  for (int k = threadIdx.x; k < kGjTile * kGjNb; k += 256) {
    const int i = j0 + k / kGjNb, c = k % kGjNb;
    Cbuf[(size_t)i * kGjNb + c] = A[(size_t)i * kp + t * kGjNb + c];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kGjNb * kGjTile; k += 256) {
    const int r = k / kGjTile, jj = k % kGjTile;
    double acc = 0.0;
#pragma unroll 8
    for (int c = 0; c < kGjNb; ++c) acc = fma(Ps[r][c], As[c][jj], acc);
    Rbuf[(size_t)r * kp + j0 + jj] = acc;
  }
}

__global__ void __launch_bounds__(256) gj_update_kernel(double* __restrict__ A, int kp, int t, const double* __restrict__ Rbuf,
                                                        const double* __restrict__ Cbuf) {
  __shared__ double Cs[kGjTile][kGjNb + 1];
  __shared__ double Rs[kGjNb][kGjTile];
  const int i0 = blockIdx.y * kGjTile, j0 = blockIdx.x * kGjTile;
  for (int k = threadIdx.x; k < kGjTile * kGjNb; k += 256) Cs[k / kGjNb][k % kGjNb] = Cbuf[(size_t)(i0 + k / kGjNb) * kGjNb + k % kGjNb];
  for (int k = threadIdx.x; k < kGjNb * kGjTile; k += 256) Rs[k / kGjTile][k % kGjTile] = Rbuf[(size_t)(k / kGjTile) * kp + j0 + k % kGjTile];
  __syncthreads();
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
  for (int k = 0; k < kGjNb; ++k) {
    double cv[4], rv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) cv[a] = Cs[ty * 4 + a][k];
#pragma unroll
    for (int b = 0; b < 4; ++b) rv[b] = Rs[k][tx * 4 + b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(cv[a], rv[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    const int it = i - t * kGjNb;
    const bool pivot_row = it >= 0 && it < kGjNb;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int jj = tx * 4 + b, j = j0 + jj;
      const int jt = j - t * kGjNb;
      double out;
      if (pivot_row) {
        out = Rs[it][jj];
      } else {
        const double old = (jt >= 0 && jt < kGjNb) ? 0.0 : A[(size_t)i * kp + j];
        out = old - acc[a][b];
      }
      A[(size_t)i * kp + j] = out;
    }
  }
}

int dense_inverse(ptfem_ctx* ctx, double* A, int kp, int32_t* flag) {
  DevBuf<double> Pinv, Rbuf, Cbuf, d0;
  PT_TRY(d0.alloc(kp));
  gj_diag_kernel<<<ceil_div(kp, 256), 256, 0, ctx->stream>>>(A, kp, d0.p);
  PT_LAUNCH_CHECK(ctx);
  PT_TRY(Pinv.alloc(kGjNb * kGjNb));
  PT_TRY(Rbuf.alloc((size_t)kGjNb * kp));
  PT_TRY(Cbuf.alloc((size_t)kp * kGjNb));
  const int steps = kp / kGjNb, tiles = kp / kGjTile;
  for (int t = 0; t < steps; ++t) {
    gj_pivot_kernel<<<1, kGjNb * kGjNb, 0, ctx->stream>>>(A, kp, t, d0.p, Pinv.p, flag);
    PT_LAUNCH_CHECK(ctx);
    gj_panel_kernel<<<tiles, 256, 0, ctx->stream>>>(A, kp, t, Pinv.p, Rbuf.p, Cbuf.p);
    PT_LAUNCH_CHECK(ctx);
    gj_update_kernel<<<dim3(tiles, tiles), 256, 0, ctx->stream>>>(A, kp, t, Rbuf.p, Cbuf.p);
    PT_LAUNCH_CHECK(ctx);
  }
  PT_CK(cudaStreamSynchronize(ctx->stream));  // the scratch buffers go back to the allocator here
  return PTFEM_OK;
}

// ---- row-partitioned solve (dist.cu): table of a row block, level-0 scaling after the cross-rank sum -----------
// splits the Dirichlet bit off a copied table (the row lists are built from cells, which Dirichlet rows keep)
__global__ void coarse_table_split_kernel(int64_t nn, float4* __restrict__ ctab, uint8_t* __restrict__ isdir) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const unsigned int cell = __float_as_uint(ctab[i].w);
  isdir[i] = (cell & kCoarseDirBit) ? 1 : 0;
  ctab[i].w = __uint_as_float(cell & ~kCoarseDirBit);
}
// level weight: B_l *= w (the additive levels overlap in what they correct)
// out[k] = val[k][sys] of the interleaved value sets
__global__ void gather_value_set_kernel(const double* __restrict__ val, int64_t nnz, int VS, int sys, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) out[k] = val[k * VS + sys];
}
__global__ void coarse_weight_kernel(double* __restrict__ binv, int64_t n, double w) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) binv[i] *= w;
}
// y_c = binv * r_c on a diagonal-only level whose r_c was summed over the ranks by the caller
__global__ void coarse_scale_kernel(int64_t k, const double* __restrict__ binv, const double* __restrict__ rc,
                                    double* __restrict__ yc) {
  const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (I < k) yc[I] = rc[I] * __ldg(binv + I);
}

// ---- host side ------------------------------------------------------------------------------------------------
int build_level_geometry(ptfem_mesh* m, CoarseLevel& L, bool lists = true) {
  ptfem_ctx* ctx = m->ctx;
  const int64_t nn = (m->coarse && m->coarse->row_limit >= 0) ? m->coarse->row_limit : m->nn;   // rows that enter the lists
  L.ncell = (int64_t)L.g.n[0] * L.g.n[1] * L.g.n[2];
  L.k = (int64_t)(L.g.n[0] + 1) * (L.g.n[1] + 1) * (L.g.n[2] + 1);
  if (!lists) return PTFEM_OK;   // only the finest level (restriction, diagonals) and the exact one (Galerkin matrix) walk row lists
  DevBuf<int32_t> key, key2, id;
  PT_TRY(key.alloc(nn));
  PT_TRY(key2.alloc(nn));
  PT_TRY(id.alloc(nn));
  PT_TRY(L.rows.alloc(nn));
  PT_TRY(L.cellptr.alloc(L.ncell + 1));
  cell_key_kernel<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(L.g.n[0], L.g.n[1], L.shift, m->coarse->ctab.p, nn, key.p, id.p);
  PT_LAUNCH_CHECK(ctx);
  int bits = 1;
  while (((int64_t)1 << bits) < L.ncell) ++bits;
  size_t tmp_bytes = 0;
  PT_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, id.p, L.rows.p, (int)nn, 0, bits, ctx->stream));
  DevBuf<uint8_t> tmp;
  PT_TRY(tmp.alloc(tmp_bytes));
  PT_CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, id.p, L.rows.p, (int)nn, 0, bits, ctx->stream));
  ctx->launches += 2;
  cell_ptr_kernel<<<ceil_div(nn + 1, 256), 256, 0, ctx->stream>>>(key2.p, nn, L.ncell, L.cellptr.p);
  PT_LAUNCH_CHECK(ctx);
  PT_CK(cudaStreamSynchronize(ctx->stream));
  return PTFEM_OK;
}

void choose_grid(const ptfem_mesh* m, double target_nodes, CoarseGrid& g) {
  double ext[3], vol = 1.0;
  for (int d = 0; d < 3; ++d) {
    ext[d] = m->bb_hi[d] - m->bb_lo[d];
    if (!(ext[d] > 0.0)) ext[d] = 1.0;
    vol *= ext[d];
  }
  // about `target_nodes` grid nodes with near-cubic cells, at least 2 cells per axis
  double h = cbrt(vol / target_nodes);
  for (int it = 0; it < 8; ++it) {
    double nodes = 1.0;
    for (int d = 0; d < 3; ++d) nodes *= std::max(2.0, std::round(ext[d] / h)) + 1.0;
    h *= cbrt(nodes / target_nodes);
  }
  for (int d = 0; d < 3; ++d) {
    g.n[d] = (int)std::max(2.0, std::round(ext[d] / h));
    // the box is widened by a hair so that nodes on the upper faces fall inside the last cell
    g.lo[d] = m->bb_lo[d];
    g.inv_h[d] = (double)g.n[d] / (ext[d] * (1.0 + 1e-12));
  }
}

// grid of the cooperative chain kernel (all CTAs co-resident) and its per-CTA dot buffer; called outside graph capture
template <int S>
int chain_setup_t(ptfem_ctx* ctx, CoarseSpace& cs) {
  int nb = 0;
  PT_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, coarse_chain_kernel<S>, 256, 0));
  if (nb > 4) nb = 4;
  cs.chain_grid = nb > 0 ? nb * ctx->sm_count : 0;
  if (cs.chain_grid > 0) PT_TRY(cs.chain_part.alloc((size_t)kMaxCoarseLevels * cs.chain_grid * 16));
  return PTFEM_OK;
}
int chain_setup(ptfem_ctx* ctx, CoarseSpace& cs, int S) {
  cs.chain_grid = 0;
  if (!ctx->tune_coarse_fused || cs.VS > 1) return PTFEM_OK;
  switch (S) {
    case 1: return chain_setup_t<1>(ctx, cs);
    case 2: return chain_setup_t<2>(ctx, cs);
    case 4: return chain_setup_t<4>(ctx, cs);
    case 8: return chain_setup_t<8>(ctx, cs);
    case 16: return chain_setup_t<16>(ctx, cs);
  }
  return PTFEM_OK;
}
template <int S>
int chain_launch(ptfem_ctx* ctx, CoarseSpace& cs, bool do_node, bool scaled0) {
  ChainArgs a;
  a.nlev = cs.nlev;
  a.split = cs.lev[0].split;
  a.do_node = do_node ? 1 : 0;
  a.scaled0 = scaled0 ? 1 : 0;
  for (int l = 0; l < cs.nlev; ++l) {
    CoarseLevel& L = cs.lev[l];
    a.lev[l].g = L.g;
    a.lev[l].k = L.k;
    a.lev[l].kp = L.kp;
    a.lev[l].exact = L.exact ? 1 : 0;
    a.lev[l].binv = L.binv.p;
    a.lev[l].rc = L.rc.p;
    a.lev[l].yc = L.yc.p;
    a.lev[l].yt = L.yt.p;
  }
  a.part = cs.lev[0].part.p;
  a.dpart = cs.chain_part.p;
  a.cdot = cs.cdot.p;
  void* args[] = {&a};
  PT_CK(cudaLaunchCooperativeKernel((const void*)coarse_chain_kernel<S>, dim3(cs.chain_grid), dim3(256), args, 0, ctx->stream));
  ctx->launches++;
  return PTFEM_OK;
}

template <int S>
int apply_t(ptfem_ctx* ctx, CoarseSpace& cs, const double* r, const FusedUpdate* fu = nullptr, cudaEvent_t after_restrict = nullptr) {
  // mesh -> finest grid
  CoarseLevel& L0 = cs.lev[0];
  {
    const int64_t ntask = L0.ncell * L0.split;
    if (fu) {
      // with the CG residual update fused in (r is rewritten): the grid stays within the CG workspace's partial sums
      // ONE wave of the resident CTAs (3 per SM, 80 registers): the grid used to be 8 per SM = 2.67 waves.  Measured on L / 8 RHS
      // (profiles/r03_fused_restrict_knobs.txt, ms per PCG iteration): 3 per SM 0.696, 6: 0.715, 8: 0.729, 4: 0.772; compiled for
      // 4 CTAs per SM (64 registers, spills) 0.739-0.784; L2 prefetch of the next trip's rows / the next task's lists: no gain
      const int occ = ctx->tune_fuse_occ == 4 ? 4 : 3;
      const int per_sm = ctx->tune_fuse_grid > 0 ? std::min(ctx->tune_fuse_grid, 8) : occ;
      const int grid = (int)std::min<int64_t>((ntask + 7) / 8, (int64_t)ctx->sm_count * per_sm);
      bool piped = false;
      if constexpr (S >= 2) {
        if (ctx->tune_fuse_pipe) {
          constexpr size_t smem = pipe_smem_bytes<S>();
          const void* fn = reinterpret_cast<const void*>(&restrict_fused_pipe_kernel<S>);
          auto it = ctx->func_smem.find(fn);
          if (it == ctx->func_smem.end()) {
            PT_CK(cudaFuncSetAttribute(restrict_fused_pipe_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->func_smem[fn] = smem;
          }
          const int pgrid = (int)std::min<int64_t>((ntask + 7) / 8, (int64_t)ctx->sm_count * (ctx->tune_fuse_grid > 0 ? std::min(ctx->tune_fuse_grid, 8) : 3));
          restrict_fused_pipe_kernel<S><<<pgrid, 256, smem, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p,
                                                                          L0.part.p, *fu, ctx->tune_fuse_pipe == 2 ? 1 : 0);
          piped = true;
        }
      }
      if (piped) {
      } else if (occ == 4)
        restrict_cell_kernel<S, 4, true><<<grid, 256, 0, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p, r,
                                                                        L0.part.p, *fu);
      else
        restrict_cell_kernel<S, 3, true><<<grid, 256, 0, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p, r,
                                                                        L0.part.p, *fu);
    } else {
    const int grid = (int)std::min<int64_t>((ntask + 7) / 8, (int64_t)ctx->sm_count * (ctx->tune_restrict_grid > 0 ? ctx->tune_restrict_grid : 32));
    if (ctx->tune_restrict_occ >= 6)
      restrict_cell_kernel<S, 6><<<grid, 256, 0, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p, r,
                                                                L0.part.p, FusedUpdate());
    else if (ctx->tune_restrict_occ >= 3)
      restrict_cell_kernel<S, 4><<<grid, 256, 0, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p, r,
                                                                L0.part.p, FusedUpdate());
    else
      restrict_cell_kernel<S, 2><<<grid, 256, 0, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p,
                                                                r, L0.part.p, FusedUpdate());
    }
    PT_LAUNCH_CHECK(ctx);
  }
  if (after_restrict) PT_CK(cudaEventRecord(after_restrict, ctx->stream));
  if (cs.chain_grid > 0) return chain_launch<S>(ctx, cs, true, false);
  for (int l = 0; l < cs.nlev; ++l) {
    CoarseLevel& L = cs.lev[l];
    double* cdot = cs.cdot.p + (size_t)l * 16;
    // (24 CTAs per SM for the finest level's gather, one element per thread, measured slower: 0.813 vs 0.780 ms per iteration)
    const int ngrid = std::min(ceil_div(L.k * S, 256), 4 * ctx->sm_count);
    // one operator per system (batched matrices): system s reads binv + s * bs
    const int64_t bs = cs.VS > 1 ? (L.exact ? (int64_t)L.kp * L.kp : L.k) : 0;
    if (l == 0) {
      if (L.exact)
        coarse_node_kernel<S, false><<<ngrid, 256, 0, ctx->stream>>>(L.g, L.k, L.split, L.part.p, nullptr, 0, L.rc.p, L.yc.p, cs.dpart.p,
                                                                     cdot, cs.ticket.p);
      else
        coarse_node_kernel<S, true><<<ngrid, 256, 0, ctx->stream>>>(L.g, L.k, L.split, L.part.p, L.binv.p, bs, L.rc.p, L.yc.p, cs.dpart.p,
                                                                    cdot, cs.ticket.p);
    } else {
      if (L.exact)
        grid_restrict_kernel<S, false><<<ngrid, 256, 0, ctx->stream>>>(L.g, L.k, cs.lev[l - 1].rc.p, nullptr, 0, L.rc.p, L.yc.p,
                                                                       cs.dpart.p, cdot, cs.ticket.p);
      else
        grid_restrict_kernel<S, true><<<ngrid, 256, 0, ctx->stream>>>(L.g, L.k, cs.lev[l - 1].rc.p, L.binv.p, bs, L.rc.p, L.yc.p,
                                                                      cs.dpart.p, cdot, cs.ticket.p);
    }
    PT_LAUNCH_CHECK(ctx);
    if (L.exact) {
      coarse_dense_kernel<S><<<L.kp / kDenseRows, 256, 0, ctx->stream>>>(L.kp, L.binv.p, bs, L.rc.p, L.yc.p, cs.dpart.p, cdot,
                                                                        cs.ticket.p);
      PT_LAUNCH_CHECK(ctx);
    }
  }
  // coarsest -> finest grid: yt_l = y_l + P yt_{l+1}
  for (int l = cs.nlev - 2; l >= 0; --l) {
    CoarseLevel& L = cs.lev[l];
    const double* ytc = (l + 1 == cs.nlev - 1) ? cs.lev[l + 1].yc.p : cs.lev[l + 1].yt.p;
    grid_prolong_kernel<S><<<ceil_div(L.k * S, 256), 256, 0, ctx->stream>>>(cs.lev[l + 1].g, L.k, L.yc.p, ytc, L.yt.p);
    PT_LAUNCH_CHECK(ctx);
  }
  return PTFEM_OK;
}

}  // namespace

int coarse_apply(ptfem_ctx* ctx, CoarseSpace& cs, int S, const double* r) {
  switch (S) {
    case 1: return apply_t<1>(ctx, cs, r);
    case 2: return apply_t<2>(ctx, cs, r);
    case 4: return apply_t<4>(ctx, cs, r);
    case 8: return apply_t<8>(ctx, cs, r);
    case 16: return apply_t<16>(ctx, cs, r);
  }
  return set_err(PTFEM_ERR_ARG, "unsupported system count %d", S);
}

int coarse_apply_fused_update(ptfem_ctx* ctx, CoarseSpace& cs, int S, double* r, const double* q, const double* dinv, const double* alpha,
                              double* partial, unsigned int* ticket, double* out_rz, double* out_rr, cudaEvent_t after_restrict) {
  if (cs.VS != 1 || cs.row_limit >= 0) return set_err(PTFEM_ERR_STATE, "the fused residual update takes one shared matrix and every row");
  FusedUpdate fu;
  fu.q = q; fu.dinv = dinv; fu.alpha = alpha; fu.r = r; fu.partial = partial; fu.ticket = ticket; fu.out_rz = out_rz; fu.out_rr = out_rr;
  fu.prefetch = ctx->tune_fuse_prefetch;
  switch (S) {
    case 1: return apply_t<1>(ctx, cs, r, &fu, after_restrict);
    case 2: return apply_t<2>(ctx, cs, r, &fu, after_restrict);
    case 4: return apply_t<4>(ctx, cs, r, &fu, after_restrict);
    case 8: return apply_t<8>(ctx, cs, r, &fu, after_restrict);
    case 16: return apply_t<16>(ctx, cs, r, &fu, after_restrict);
  }
  return set_err(PTFEM_ERR_ARG, "unsupported system count %d", S);
}

CoarseDev coarse_dev(const CoarseSpace& cs) {
  // the CG p-update interpolates from the finest grid only (it carries the sum of all levels)
  CoarseDev d;
  d.nx1 = cs.lev[0].g.n[0] + 1;
  d.ny1 = cs.lev[0].g.n[1] + 1;
  d.shift = cs.lev[0].shift;
  d.y = cs.nlev > 1 ? cs.lev[0].yt.p : cs.lev[0].yc.p;
  d.ctab = cs.ctab.p;
  return d;
}

void coarse_free(CoarseSpace* cs) { delete cs; }

int coarse_prepare(ptfem_mesh* m, int target_nodes, int extra_levels, int S) {
  ptfem_ctx* ctx = m->ctx;
  if (!m->coarse) m->coarse = new CoarseSpace();
  CoarseSpace& cs = *m->coarse;
  // batched matrices (nvalp value sets on one pattern, one per system): the grids, Z and the row lists are shared, every
  // system gets its own Galerkin operators (stored one after the other), built by the same kernels from its value set
  const int VS = m->nvalp > 1 ? m->nvalp : 1;
  if (VS > 1 && VS != S) return set_err(PTFEM_ERR_ARG, "batched matrices: %d value sets but %d systems", VS, S);
  if (VS > 1 && (cs.partial_mode || cs.row_limit >= 0))
    return set_err(PTFEM_ERR_ARG, "the distributed coarse set-up takes one matrix");
  const bool vs_changed = cs.VS != VS;
  cs.VS = VS;
  // The iteration count is set by the FINEST grid; the exactly inverted one only has to be coarse enough for its O(k^3)
  // inverse to cost nothing: 8x6x4 cells (315 unknowns) under three diagonal-only levels needs the same 60 iterations on the
  // 20 M-tet slab as 16x12x8 (1989 unknowns) under two (CPU study, profiles/r02_coarse_grid_size_cpu.txt), and its inverse
  // takes 0.1 ms instead of 6.4.
  if (target_nodes <= 0) target_nodes = kDefaultCoarseNodes;
  if (extra_levels > kMaxCoarseLevels - 1) extra_levels = kMaxCoarseLevels - 1;
  struct Events {  // destroyed on every return path
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
    }
  } ev;
  PT_CK(cudaEventCreate(&ev.e0));
  PT_CK(cudaEventCreate(&ev.e1));
  cudaEvent_t e0 = ev.e0, e1 = ev.e1;
  PT_CK(cudaEventRecord(e0, ctx->stream));
  bool rebuilt = false;
  if (!cs.geom_ok || cs.req_nodes != target_nodes || cs.req_levels != extra_levels) {
    {  // extra_levels < 0: as many finer grids as keep >= 16 mesh nodes per finest cell; a request is honoured down to 8
      CoarseGrid b0;
      choose_grid(m, (double)target_nodes, b0);
      const double cells = (double)b0.n[0] * b0.n[1] * b0.n[2];
      const double min_rows = extra_levels < 0 ? 16.0 : 8.0;
      cs.nlev = extra_levels < 0 ? kMaxCoarseLevels : extra_levels + 1;
      const double nn_rule = (double)(cs.nn_levels > 0 ? cs.nn_levels : m->nn);   // (distributed set-up: the WHOLE mesh's node count)
      while (cs.nlev > 1 && cells * pow(8.0, cs.nlev - 1) * min_rows > nn_rule) --cs.nlev;
    }
    CoarseGrid base;
    choose_grid(m, (double)target_nodes, base);
    if (base.n[0] > kCoarseMaxCells || base.n[1] > kCoarseMaxCells || base.n[2] > kCoarseMaxCells)
      return set_err(PTFEM_ERR_ARG, "coarsest grid of %dx%dx%d cells: at most %d per axis", base.n[0], base.n[1], base.n[2], kCoarseMaxCells);
    PT_TRY(cs.ctab.alloc((size_t)m->nn));
    coarse_table_kernel<<<ceil_div(m->nn, 256), 256, 0, ctx->stream>>>(base, m->xyz.p, m->nn, cs.ctab.p);
    PT_LAUNCH_CHECK(ctx);
    for (int l = 0; l < cs.nlev; ++l) {
      CoarseLevel& L = cs.lev[l];
      L.shift = cs.nlev - 1 - l;
      const int f = 1 << L.shift;
      L.g = base;
      for (int d = 0; d < 3; ++d) {
        L.g.n[d] = base.n[d] * f;
        L.g.inv_h[d] = base.inv_h[d] * f;
      }
      L.exact = (l == cs.nlev - 1);
      PT_TRY(build_level_geometry(m, L, l == 0 || L.exact));
      L.kp = L.exact ? (int)((L.k + kGjTile - 1) / kGjTile) * kGjTile : 0;
      // restriction tasks (one warp each) of at most ~96 rows
      L.split = 1;
      while ((cs.nn_levels > 0 ? cs.nn_levels : m->nn) / (L.ncell * L.split) > 96 && L.split < 64) L.split *= 2;
    }
    cs.geom_ok = true;
    cs.req_nodes = target_nodes;
    cs.req_levels = extra_levels;
    cs.matrix_epoch = -1;
    cs.S = 0;
    rebuilt = true;
    cs.generation++;
  }
  if (cs.S != S) {
    size_t maxgrid = 1;
    for (int l = 0; l < cs.nlev; ++l) {
      CoarseLevel& L = cs.lev[l];
      const size_t kk = L.exact ? (size_t)L.kp : (size_t)L.k;
      if (l == 0) PT_TRY(L.part.alloc((size_t)L.ncell * L.split * 8 * S));
      if (l < cs.nlev - 1) PT_TRY(L.yt.alloc(kk * S));
      PT_TRY(L.rc.alloc(kk * S));
      PT_TRY(L.yc.alloc(kk * S));
      PT_CK(cudaMemsetAsync(L.rc.p, 0, kk * S * sizeof(double), ctx->stream));
      PT_CK(cudaMemsetAsync(L.yc.p, 0, kk * S * sizeof(double), ctx->stream));
      maxgrid = std::max(maxgrid, (size_t)ceil_div(kk * S, 256) + 1);
      maxgrid = std::max(maxgrid, (size_t)kk / kDenseRows + 1);
    }
    PT_TRY(cs.dpart.alloc(maxgrid * 16));
    PT_TRY(cs.cdot.alloc((size_t)kMaxCoarseLevels * 16));
    PT_CK(cudaMemsetAsync(cs.cdot.p, 0, (size_t)kMaxCoarseLevels * 16 * sizeof(double), ctx->stream));
    if (!cs.ticket.p) {
      PT_TRY(cs.ticket.alloc(4));
      PT_CK(cudaMemsetAsync(cs.ticket.p, 0, 4 * sizeof(unsigned int), ctx->stream));
    }
    PT_TRY(chain_setup(ctx, cs, S));
    cs.S = S;
    rebuilt = true;
    cs.generation++;
  }
  if (cs.matrix_epoch != m->matrix_epoch || vs_changed) {
    if (vs_changed) cs.generation++;
    PT_TRY(cs.flag.alloc(2));
    PT_CK(cudaMemsetAsync(cs.flag.p, 0, 2 * sizeof(int32_t), ctx->stream));
    coarse_table_flag_kernel<<<ceil_div(m->nn, 256), 256, 0, ctx->stream>>>(m->isdir.p, m->nn, cs.ctab.p);
    PT_LAUNCH_CHECK(ctx);
    const int64_t nlist = cs.row_limit >= 0 ? cs.row_limit : m->nn;
    PT_TRY(cs.ctab0.alloc((size_t)m->nn));
    gather_table_kernel<<<ceil_div(nlist, 256), 256, 0, ctx->stream>>>(cs.ctab.p, cs.lev[0].rows.p, nlist, cs.ctab0.p);
    PT_LAUNCH_CHECK(ctx);
    // the additive levels overlap in what they correct (every level sees the smooth part of r): each is weighted by
    // 2 / (levels + 1) against the Jacobi term - CPU study profiles/r01_precond_level_weights_L_cpu.txt: 71 -> 60 iterations
    // on the 3-level bench mesh, 59 -> 56 with 2 levels, unchanged with one; PTFEM_COARSE_WEIGHT overrides
    const double level_w = ctx->tune_coarse_weight > 0.0 ? ctx->tune_coarse_weight : 2.0 / (cs.nlev + 1);
    DevBuf<double> valset;          // value set of one system, gathered out of the interleaved [nnz][VS] array
    if (VS > 1) PT_TRY(valset.alloc(m->nnz + 8));
    for (int l = 0; l < cs.nlev; ++l) {
      CoarseLevel& L = cs.lev[l];
      const size_t want = (L.exact ? (size_t)L.kp * L.kp : (size_t)L.k) * VS;
      const double* before = L.binv.p;
      PT_TRY(L.binv.alloc(want));
      if (L.binv.p != before) cs.generation++;
    }
    for (int sys = 0; sys < VS; ++sys) {
    const double* val = m->val_bc.p;
    if (VS > 1) {
      gather_value_set_kernel<<<ceil_div(m->nnz, 256), 256, 0, ctx->stream>>>(m->val_bc.p, m->nnz, VS, sys, valset.p);
      PT_LAUNCH_CHECK(ctx);
      val = valset.p;
    }
    for (int l = 0; l < cs.nlev; ++l) {
      CoarseLevel& L = cs.lev[l];
      if (L.exact) {
        const size_t n2 = (size_t)L.kp * L.kp;
        double* binv = L.binv.p + (size_t)sys * n2;
        PT_CK(cudaMemsetAsync(binv, 0, n2 * sizeof(double), ctx->stream));
        DevBuf<double> blockE;
        int gsplit = 1;   // CTAs per cell: about eight CTAs per SM in all, each with at least a few chunks of rows
        while (L.ncell * gsplit < (int64_t)8 * ctx->sm_count && nlist / (L.ncell * gsplit) > 4 * kGalRows && gsplit < 64) gsplit *= 2;
        PT_TRY(blockE.alloc((size_t)L.ncell * gsplit * 512));
        galerkin_cell_kernel<<<(unsigned)(L.ncell * gsplit), 256, 0, ctx->stream>>>(L.g, L.cellptr.p, L.rows.p, cs.ctab.p,
                                                                                     m->rowptr.p, m->col.p, val, blockE.p,
                                                                                     binv, L.kp, cs.flag.p, gsplit);
        PT_LAUNCH_CHECK(ctx);
        galerkin_gather_kernel<<<ceil_div((int64_t)n2, 256), 256, 0, ctx->stream>>>(L.g, L.k, L.kp, blockE.p, binv, gsplit,
                                                                                    cs.partial_mode ? 0 : 1);
        PT_LAUNCH_CHECK(ctx);
        PT_CK(cudaStreamSynchronize(ctx->stream));   // blockE goes back to the allocator
        if (cs.partial_mode) continue;
        PT_TRY(dense_inverse(ctx, binv, L.kp, cs.flag.p));
        if (level_w != 1.0) {
          coarse_weight_kernel<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(binv, (int64_t)n2, level_w);
          PT_LAUNCH_CHECK(ctx);
        }
      }
    }
    // diagonal-only levels (0 .. nlev-2): one pass over the matrix for all of them
    if (cs.nlev > 1) {
      const int nb = cs.nlev - 1;
      CoarseLevel& L0 = cs.lev[0];
      DevBuf<double> dpartf;
      PT_TRY(dpartf.alloc((size_t)nb * L0.ncell * 8));
      const int grid = (int)std::min<int64_t>((L0.ncell + 7) / 8, (int64_t)ctx->sm_count * 32);
      galerkin_diag_multi_kernel<<<grid, 256, 0, ctx->stream>>>(nb, L0.shift, L0.ncell, L0.cellptr.p, L0.rows.p, cs.ctab.p, m->rowptr.p,
                                                                m->col.p, val, dpartf.p);
      PT_LAUNCH_CHECK(ctx);
      for (int l = 0; l < nb; ++l) {
        CoarseLevel& L = cs.lev[l];
        double* binv = L.binv.p + (size_t)sys * L.k;
        galerkin_diag_node_multi_kernel<<<ceil_div(L.k * 32, 256), 256, 0, ctx->stream>>>(L.g, L.k, 1 << l, L0.g.n[0], L0.g.n[1],
                                                                                         dpartf.p + (size_t)l * L0.ncell * 8, binv,
                                                                                         cs.partial_mode ? 0 : 1);
        PT_LAUNCH_CHECK(ctx);
        if (level_w != 1.0 && !cs.partial_mode) {
          coarse_weight_kernel<<<ceil_div(L.k, 256), 256, 0, ctx->stream>>>(binv, L.k, level_w);
          PT_LAUNCH_CHECK(ctx);
        }
      }
      PT_CK(cudaStreamSynchronize(ctx->stream));   // dpartf goes back to the allocator
    }
    }   // systems
    int32_t hflag[2] = {0, 0};
    PT_CK(cudaMemcpyAsync(hflag, cs.flag.p, sizeof hflag, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    if (hflag[0])
      return set_err(PTFEM_ERR_STATE, "coarse Galerkin matrix is singular or indefinite to working precision (grid %dx%dx%d too fine for this mesh?)",
                     cs.lev[cs.nlev - 1].g.n[0], cs.lev[cs.nlev - 1].g.n[1], cs.lev[cs.nlev - 1].g.n[2]);
    if (cs.partial_mode) {
      cs.partial_ready = true;   // raw sums of the owned rows; coarse_finish_sums completes them
    } else {
      cs.matrix_epoch = m->matrix_epoch;
    }
    rebuilt = true;
  }
  PT_CK(cudaEventRecord(e1, ctx->stream));
  PT_CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (rebuilt) cs.setup_ms = ms;
  return PTFEM_OK;
}


// ---- distributed set-up: Galerkin sums of the owned rows, added up over the ranks by the launcher ------------------------
static int64_t partial_sum_count(const CoarseSpace& cs) {
  int64_t n = 0;
  for (int l = 0; l < cs.nlev; ++l) n += cs.lev[l].exact ? (int64_t)cs.lev[l].kp * cs.lev[l].kp : cs.lev[l].k;
  return n;
}
int coarse_partial_sums(ptfem_mesh* m, int64_t* n_out, double* out_host, int64_t cap) {
  if (!m->coarse || !m->coarse->partial_ready) return set_err(PTFEM_ERR_STATE, "the partial coarse set-up has not been run");
  CoarseSpace& cs = *m->coarse;
  const int64_t n = partial_sum_count(cs);
  if (n_out) *n_out = n;
  if (!out_host) return PTFEM_OK;
  if (cap < n) return set_err(PTFEM_ERR_ARG, "buffer of %lld doubles, %lld needed", (long long)cap, (long long)n);
  int64_t off = 0;
  for (int l = cs.nlev - 1; l >= 0; --l) {   // exact level first, then levels 0 .. nlev-2
    if (!cs.lev[l].exact) continue;
    const int64_t c = (int64_t)cs.lev[l].kp * cs.lev[l].kp;
    PT_CK(cudaMemcpyAsync(out_host + off, cs.lev[l].binv.p, c * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    off += c;
  }
  for (int l = 0; l < cs.nlev - 1; ++l) {
    PT_CK(cudaMemcpyAsync(out_host + off, cs.lev[l].binv.p, cs.lev[l].k * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    off += cs.lev[l].k;
  }
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}
int coarse_finish_sums(ptfem_mesh* m, const double* sums_host, int64_t n) {
  if (!m->coarse || !m->coarse->partial_ready) return set_err(PTFEM_ERR_STATE, "the partial coarse set-up has not been run");
  ptfem_ctx* ctx = m->ctx;
  CoarseSpace& cs = *m->coarse;
  if (n != partial_sum_count(cs)) return set_err(PTFEM_ERR_ARG, "%lld sums given, %lld expected", (long long)n, (long long)partial_sum_count(cs));
  const double level_w = ctx->tune_coarse_weight > 0.0 ? ctx->tune_coarse_weight : 2.0 / (cs.nlev + 1);
  int64_t off = 0;
  CoarseLevel& X = cs.lev[cs.nlev - 1];
  const int64_t n2 = (int64_t)X.kp * X.kp;
  PT_CK(cudaMemcpyAsync(X.binv.p, sums_host + off, n2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  off += n2;
  galerkin_fix_kernel<<<ceil_div(X.kp, 256), 256, 0, ctx->stream>>>(X.k, X.kp, X.binv.p);
  PT_LAUNCH_CHECK(ctx);
  PT_CK(cudaMemsetAsync(cs.flag.p, 0, 2 * sizeof(int32_t), ctx->stream));
  PT_TRY(dense_inverse(ctx, X.binv.p, X.kp, cs.flag.p));
  if (level_w != 1.0) {
    coarse_weight_kernel<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(X.binv.p, n2, level_w);
    PT_LAUNCH_CHECK(ctx);
  }
  for (int l = 0; l < cs.nlev - 1; ++l) {
    CoarseLevel& L = cs.lev[l];
    PT_CK(cudaMemcpyAsync(L.binv.p, sums_host + off, L.k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    off += L.k;
    diag_invert_kernel<<<ceil_div(L.k, 256), 256, 0, ctx->stream>>>(L.k, level_w, L.binv.p);
    PT_LAUNCH_CHECK(ctx);
  }
  int32_t hflag[2] = {0, 0};
  PT_CK(cudaMemcpyAsync(hflag, cs.flag.p, sizeof hflag, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  if (hflag[0]) return set_err(PTFEM_ERR_STATE, "coarse Galerkin matrix is singular or indefinite to working precision");
  cs.matrix_epoch = m->matrix_epoch;
  cs.partial_mode = false;   // the spaces now hold the operators of the whole matrix
  cs.generation++;
  return PTFEM_OK;
}

// ---- row-partitioned solve ------------------------------------------------------------------------------------
// The ranks of a partitioned solve each hold a replica of the mesh (they assembled it, as in a sweep) and a block of
// rows.  The coarse spaces are those of the replica: grids, Dirichlet flags and Galerkin operators are computed on
// every rank from the whole matrix (deterministic, hence identical everywhere, no communication) and copied; what is
// distributed is the work that scales with the mesh - restriction and prolongation touch the owned rows only, the
// finest grid vector is summed over the ranks (dist.cu), the grid hierarchy above it is replicated.
int coarse_attach_rows(ptfem_mesh* sys, ptfem_mesh* full, int64_t row0) {
  ptfem_ctx* ctx = sys->ctx;
  if (!full->coarse || !full->coarse->geom_ok || full->coarse->matrix_epoch != full->matrix_epoch)
    return set_err(PTFEM_ERR_STATE, "the replica's coarse spaces are not prepared for its current matrix");
  const CoarseSpace& F = *full->coarse;
  const int64_t nloc = sys->nn;
  if (row0 < 0 || row0 + nloc > full->nn) return set_err(PTFEM_ERR_ARG, "row block [%lld, %lld) outside the replica", (long long)row0, (long long)(row0 + nloc));
  if (sys->coarse) coarse_free(sys->coarse);
  sys->coarse = new CoarseSpace();
  CoarseSpace& cs = *sys->coarse;
  cs.nlev = F.nlev;
  PT_TRY(cs.ctab.alloc((size_t)nloc));
  PT_CK(cudaMemcpyAsync(cs.ctab.p, F.ctab.p + row0, (size_t)nloc * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  DevBuf<uint8_t> isdir;
  PT_TRY(isdir.alloc(nloc));
  coarse_table_split_kernel<<<ceil_div(nloc, 256), 256, 0, ctx->stream>>>(nloc, cs.ctab.p, isdir.p);
  PT_LAUNCH_CHECK(ctx);
  size_t maxgrid = 1;
  for (int l = 0; l < cs.nlev; ++l) {
    const CoarseLevel& FL = F.lev[l];
    CoarseLevel& L = cs.lev[l];
    L.g = FL.g;
    L.shift = FL.shift;
    L.exact = FL.exact;
    L.kp = FL.kp;
    L.k = FL.k;
    L.ncell = FL.ncell;
    L.split = FL.split;     // populated cells are as dense as on the replica
    if (l == 0) {
      PT_TRY(build_level_geometry(sys, L));   // row list of the owned rows (most cells of the grid are empty here)
      PT_TRY(L.part.alloc((size_t)L.ncell * L.split * 8));
    }
    const size_t kk = L.exact ? (size_t)L.kp : (size_t)L.k;
    const size_t nb = L.exact ? (size_t)L.kp * L.kp : (size_t)L.k;
    PT_TRY(L.binv.alloc(nb));
    PT_CK(cudaMemcpyAsync(L.binv.p, FL.binv.p, nb * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (l < cs.nlev - 1) PT_TRY(L.yt.alloc(kk));
    PT_TRY(L.rc.alloc(kk));
    PT_TRY(L.yc.alloc(kk));
    PT_CK(cudaMemsetAsync(L.rc.p, 0, kk * sizeof(double), ctx->stream));
    PT_CK(cudaMemsetAsync(L.yc.p, 0, kk * sizeof(double), ctx->stream));
    maxgrid = std::max(maxgrid, (size_t)ceil_div(kk, 256) + 1);
    maxgrid = std::max(maxgrid, (size_t)kk / kDenseRows + 1);
  }
  coarse_table_flag_kernel<<<ceil_div(nloc, 256), 256, 0, ctx->stream>>>(isdir.p, nloc, cs.ctab.p);
  PT_LAUNCH_CHECK(ctx);
  PT_TRY(cs.ctab0.alloc((size_t)nloc));
  gather_table_kernel<<<ceil_div(nloc, 256), 256, 0, ctx->stream>>>(cs.ctab.p, cs.lev[0].rows.p, nloc, cs.ctab0.p);
  PT_LAUNCH_CHECK(ctx);
  PT_TRY(cs.dpart.alloc(maxgrid * 16));
  PT_TRY(cs.cdot.alloc((size_t)kMaxCoarseLevels * 16));
  PT_TRY(cs.ticket.alloc(4));
  PT_CK(cudaMemsetAsync(cs.cdot.p, 0, (size_t)kMaxCoarseLevels * 16 * sizeof(double), ctx->stream));
  PT_CK(cudaMemsetAsync(cs.ticket.p, 0, 4 * sizeof(unsigned int), ctx->stream));
  PT_TRY(chain_setup(ctx, cs, 1));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  cs.S = 1;
  cs.geom_ok = true;
  cs.req_nodes = F.req_nodes;
  cs.req_levels = F.req_levels;
  cs.setup_ms = F.setup_ms;
  cs.generation++;
  return PTFEM_OK;
}

// owned rows -> finest grid: rc_out[k_0] = this rank's part of Z_0^T r (the caller sums over the ranks)
int coarse_restrict_rows(ptfem_ctx* ctx, CoarseSpace& cs, const double* r, double* rc_out) {
  CoarseLevel& L0 = cs.lev[0];
  const int64_t ntask = L0.ncell * L0.split;
  const int grid = (int)std::min<int64_t>((ntask + 7) / 8, (int64_t)ctx->sm_count * 32);
  restrict_cell_kernel<1, 4><<<grid, 256, 0, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p, r,
                                                            L0.part.p, FusedUpdate());
  PT_LAUNCH_CHECK(ctx);
  const int ngrid = std::min(ceil_div(L0.k, 256), 4 * ctx->sm_count);
  coarse_node_kernel<1, false><<<ngrid, 256, 0, ctx->stream>>>(L0.g, L0.k, L0.split, L0.part.p, nullptr, 0, rc_out, nullptr, nullptr,
                                                               nullptr, nullptr);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

// ---- sharded form (peer-memory transport, >= 2 levels) -------------------------------------------------------------------
// A contiguous block of mesh rows touches a slab of the finest grid only: nodes [a0, b0) of level 0, and through the 27-point
// grid restriction nodes [a1, b1) of level 1.  The rank computes its partial sums on those ranges only -
// out0[a0..b0) = this rank's part of Z_0^T r and out1[a1..b1) = P^T of that part (the restriction is linear, so the sum over
// the ranks of these equals P^T of the summed level-0 vector) - and the cross-rank exchange carries a few grid planes of
// level 0 (neighbouring slabs) plus the small level-1 vector instead of the whole finest grid.
namespace {
// the touched node range of a row list: min / max finest-grid node over the rows' cells
__global__ void touched_range_kernel(int64_t nn, const float4* __restrict__ ctab, int shift, int nx1, int ny1, int nz1,
                                     unsigned long long* __restrict__ out /*[0] = min, [1] = max*/) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  int c[3];
  double t[3];
  const CoarseRaw raw = coarse_row_load(ctab, i);
  CoarseRaw live = raw;
  live.v.w = __uint_as_float(__float_as_uint(raw.v.w) & ~kCoarseDirBit);   // Dirichlet rows still sit in a cell
  coarse_row_decode(live, shift, c, t);
  (void)nz1;
  const unsigned long long lo = (unsigned long long)c[0] + (unsigned long long)nx1 * (c[1] + (unsigned long long)ny1 * c[2]);
  const unsigned long long hi = (unsigned long long)(c[0] + 1) + (unsigned long long)nx1 * ((c[1] + 1) + (unsigned long long)ny1 * (c[2] + 1));
  atomicMin(out, lo);
  atomicMax(out + 1, hi);
}
// rc[I] for finest nodes I in [I0, I1) from the restriction partials (as coarse_node_kernel, S = 1)
__global__ void __launch_bounds__(256) coarse_node_range_kernel(CoarseGrid g, int64_t I0, int64_t I1, int split,
                                                                const double* __restrict__ part, double* __restrict__ rc) {
  const int64_t I = I0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= I1) return;
  double v = 0.0;
  const int nx1 = g.n[0] + 1, ny1 = g.n[1] + 1;
  const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int cx = ix - (a & 1), cy = iy - ((a >> 1) & 1), cz = iz - (a >> 2);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
    const int64_t c = cx + (int64_t)g.n[0] * (cy + (int64_t)g.n[1] * cz);
    for (int sp = 0; sp < split; ++sp) v += __ldcg(part + (((size_t)c * split + sp) * 8 + a));
  }
  rc[I] = v;
}
// rc1[J] for level-1 nodes J in [J0, J1): 27-point restriction of rf, fine nodes outside [f0, f1) counting as zero
__global__ void __launch_bounds__(256) grid_restrict_range_kernel(CoarseGrid gc, int64_t J0, int64_t J1, int64_t f0, int64_t f1,
                                                                  const double* __restrict__ rf, double* __restrict__ rc,
                                                                  CoarseSignal sig) {
  const int64_t I = J0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (I < J1) {
  double v = 0.0;
  const int nx1 = gc.n[0] + 1, ny1 = gc.n[1] + 1;
  const int ix = (int)(I % nx1), iy = (int)((I / nx1) % ny1), iz = (int)(I / ((int64_t)nx1 * ny1));
  const int fx1 = 2 * gc.n[0] + 1, fy1 = 2 * gc.n[1] + 1, fz1 = 2 * gc.n[2] + 1;
  for (int dz = -1; dz <= 1; ++dz) {
    const int fz = 2 * iz + dz;
    if (fz < 0 || fz >= fz1) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int fy = 2 * iy + dy;
      if (fy < 0 || fy >= fy1) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int fx = 2 * ix + dx;
        if (fx < 0 || fx >= fx1) continue;
        const int64_t F = (int64_t)fx + (int64_t)fx1 * (fy + (int64_t)fy1 * fz);
        if (F < f0 || F >= f1) continue;
        const double w = (dx ? 0.5 : 1.0) * (dy ? 0.5 : 1.0) * (dz ? 0.5 : 1.0);
        v = fma(w, __ldcg(rf + F), v);
      }
    }
  }
  rc[I] = v;
  }
  if (sig.n <= 0 && !sig.seq) return;
  // last CTA: both parts of the exchange buffer are complete -> tell every other rank (release at system scope)
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(sig.ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence_system();
    const unsigned long long seq = *sig.seq + 1;
    *sig.seq = seq;
    for (int q = 0; q < sig.n; ++q) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sig.peer_flag[q]), "l"(seq) : "memory");
  }
}
// yt_f[F] = y_f[F] + (P yt_c)[F] for fine nodes F in [F0, F1)
__global__ void __launch_bounds__(256) grid_prolong_range_kernel(CoarseGrid gc, int64_t F0, int64_t F1, const double* __restrict__ yf,
                                                                 const double* __restrict__ ytc, double* __restrict__ ytf) {
  const int64_t F = F0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (F >= F1) return;
  const int fx1 = 2 * gc.n[0] + 1, fy1 = 2 * gc.n[1] + 1;
  const int nx1 = gc.n[0] + 1, ny1 = gc.n[1] + 1;
  const int fx = (int)(F % fx1), fy = (int)((F / fx1) % fy1), fz = (int)(F / ((int64_t)fx1 * fy1));
  double v = yf[F];
  for (int az = 0; az <= (fz & 1); ++az)
    for (int ay = 0; ay <= (fy & 1); ++ay)
      for (int ax = 0; ax <= (fx & 1); ++ax) {
        const double w = ((fx & 1) ? 0.5 : 1.0) * ((fy & 1) ? 0.5 : 1.0) * ((fz & 1) ? 0.5 : 1.0);
        const size_t c = (size_t)(fx / 2 + ax) + (size_t)nx1 * ((fy / 2 + ay) + (size_t)ny1 * (fz / 2 + az));
        v = fma(w, __ldcg(ytc + c), v);
      }
  ytf[F] = v;
}
}  // namespace

// ranges[4] = {a0, b0, a1, b1}: finest-grid nodes the rows of this system touch, and the level-1 nodes their restriction reaches
int coarse_touched_ranges(ptfem_ctx* ctx, CoarseSpace& cs, int64_t nn, int64_t ranges[4]) {
  CoarseLevel& L0 = cs.lev[0];
  ranges[0] = 0; ranges[1] = L0.k; ranges[2] = 0; ranges[3] = cs.nlev > 1 ? cs.lev[1].k : 0;
  if (cs.nlev < 2 || nn <= 0) return PTFEM_OK;
  DevBuf<unsigned long long> mm;
  PT_TRY(mm.alloc(2));
  const unsigned long long init[2] = {~0ull, 0ull};
  PT_CK(cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  const int nx1 = L0.g.n[0] + 1, ny1 = L0.g.n[1] + 1, nz1 = L0.g.n[2] + 1;
  touched_range_kernel<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(nn, cs.ctab.p, L0.shift, nx1, ny1, nz1, mm.p);
  PT_LAUNCH_CHECK(ctx);
  unsigned long long h[2];
  PT_CK(cudaMemcpyAsync(h, mm.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  if (h[0] > h[1]) return PTFEM_OK;
  // whole z-planes of the finest grid (node index = x + nx1 (y + ny1 z)); level 1: planes floor(zlo/2) .. ceil(zhi/2)
  const int64_t plane0 = (int64_t)nx1 * ny1;
  const int64_t zlo = (int64_t)h[0] / plane0, zhi = (int64_t)h[1] / plane0;
  ranges[0] = zlo * plane0;
  ranges[1] = std::min<int64_t>(L0.k, (zhi + 1) * plane0);
  const CoarseLevel& L1 = cs.lev[1];
  const int64_t plane1 = (int64_t)(L1.g.n[0] + 1) * (L1.g.n[1] + 1);
  ranges[2] = (zlo / 2) * plane1;
  ranges[3] = std::min<int64_t>(L1.k, ((zhi + 1) / 2 + 1) * plane1);
  return PTFEM_OK;
}

// restriction of the owned rows on their ranges: out0[a0..b0) (level 0) and out1[a1..b1) (level 1), see above
int coarse_restrict_rows_sharded(ptfem_ctx* ctx, CoarseSpace& cs, const double* r, const int64_t ranges[4], double* out0,
                                 double* out1, const CoarseSignal& sig) {
  CoarseLevel& L0 = cs.lev[0];
  const int64_t ntask = L0.ncell * L0.split;
  const int grid = (int)std::min<int64_t>((ntask + 7) / 8, (int64_t)ctx->sm_count * 32);
  restrict_cell_kernel<1, 4><<<grid, 256, 0, ctx->stream>>>(ntask, L0.split, L0.shift, L0.cellptr.p, L0.rows.p, cs.ctab0.p, r,
                                                            L0.part.p, FusedUpdate());
  PT_LAUNCH_CHECK(ctx);
  if (ranges[1] > ranges[0]) {
    coarse_node_range_kernel<<<ceil_div(ranges[1] - ranges[0], 256), 256, 0, ctx->stream>>>(L0.g, ranges[0], ranges[1], L0.split,
                                                                                           L0.part.p, out0);
    PT_LAUNCH_CHECK(ctx);
  }
  // (launched even for an empty range: its last CTA raises the "buffer complete" signal)
  const int64_t n1 = ranges[3] > ranges[2] ? ranges[3] - ranges[2] : 0;
  grid_restrict_range_kernel<<<std::max(1, ceil_div(n1, 256)), 256, 0, ctx->stream>>>(cs.lev[1].g, ranges[2], ranges[2] + n1, ranges[0],
                                                                                      ranges[1], out0, out1, sig);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

// The grid levels above `start` (a few thousand nodes in all) as ONE block of 1024 threads (row-partitioned solve, one
// right-hand side): restrictions up from level `start`, the dense coarsest solve, prolongations back down to level `down_to`;
// phases separated by block barriers.  Four dependent kernels of ~6 us launch-to-launch each did a few microseconds of
// work here.  Level `start` itself (18 k nodes on the 20 M-tet slab) is prolonged by the ordinary multi-CTA kernel: one block
// for it as well was measured slower (0.34 vs 0.27 ms per iteration on 2 GPUs).
namespace {
__global__ void __launch_bounds__(1024) coarse_chain_tail_kernel(ChainArgs a, int start, int down_to) {
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int l = start + 1; l < a.nlev; ++l) {
    const ChainLevel& L = a.lev[l];
    const double* rf = a.lev[l - 1].rc;
    for (int64_t I = tid; I < L.k; I += nth) {
      const double v = chain_restrict27<1>(L.g, rf, I, 0);
      L.rc[I] = v;
      if (!L.exact) L.yc[I] = v * L.binv[I];
    }
    __syncthreads();
  }
  {  // exact level: y = B r, one warp per row
    const ChainLevel& L = a.lev[a.nlev - 1];
    const int lane = tid & 31, wid = tid >> 5, nw = nth >> 5;
    for (int I = wid; I < L.kp; I += nw) {
      double acc = 0.0;
      for (int J = lane; J < L.kp; J += 32) acc = fma(L.binv[(size_t)I * L.kp + J], L.rc[J], acc);
      acc = warp_sum(acc);
      if (lane == 0) L.yc[I] = acc;
    }
    __syncthreads();
  }
  for (int l = a.nlev - 2; l >= down_to; --l) {
    const ChainLevel& L = a.lev[l];
    const double* ytc = (l + 1 == a.nlev - 1) ? a.lev[l + 1].yc : a.lev[l + 1].yt;
    for (int64_t F = tid; F < L.k; F += nth) L.yt[F] = chain_prolong8<1>(a.lev[l + 1].g, ytc, L.yc[F], F, 0);
    __syncthreads();
  }
}
}  // namespace

// lev[start].rc holds the summed restriction of level `start` (and, when scaled, lev[start].yc = binv rc on a diagonal level):
// the replicated grid hierarchy from that level up and back down to it (yt of level `start`, or yc when it is the last level)
int coarse_grids_apply(ptfem_ctx* ctx, CoarseSpace& cs, bool scaled0, int start) {
  if (cs.chain_grid > 0 && start == 0) return chain_launch<1>(ctx, cs, false, scaled0);
  if (start >= 1 && start + 1 < cs.nlev && scaled0 && cs.lev[start + 1].k <= 8192 && ctx->tune_chain_tail) {
    ChainArgs a;
    a.nlev = cs.nlev;
    a.split = 1;
    a.do_node = 0;
    a.scaled0 = 1;
    for (int l = 0; l < cs.nlev; ++l) {
      CoarseLevel& L = cs.lev[l];
      a.lev[l].g = L.g;
      a.lev[l].k = L.k;
      a.lev[l].kp = L.kp;
      a.lev[l].exact = L.exact ? 1 : 0;
      a.lev[l].binv = L.binv.p;
      a.lev[l].rc = L.rc.p;
      a.lev[l].yc = L.yc.p;
      a.lev[l].yt = L.yt.p;
    }
    a.part = nullptr;
    a.dpart = nullptr;
    a.cdot = nullptr;
    coarse_chain_tail_kernel<<<1, 1024, 0, ctx->stream>>>(a, start, start + 1);
    PT_LAUNCH_CHECK(ctx);
    CoarseLevel& L = cs.lev[start];
    const double* ytc = (start + 1 == cs.nlev - 1) ? cs.lev[start + 1].yc.p : cs.lev[start + 1].yt.p;
    grid_prolong_kernel<1><<<ceil_div(L.k, 256), 256, 0, ctx->stream>>>(cs.lev[start + 1].g, L.k, L.yc.p, ytc, L.yt.p);
    PT_LAUNCH_CHECK(ctx);
    return PTFEM_OK;
  }
  for (int l = start; l < cs.nlev; ++l) {
    CoarseLevel& L = cs.lev[l];
    double* cdot = cs.cdot.p + (size_t)l * 16;
    const int ngrid = std::min(ceil_div(L.k, 256), 4 * ctx->sm_count);
    if (l == start) {
      if (!L.exact && !scaled0) {
        coarse_scale_kernel<<<ceil_div(L.k, 256), 256, 0, ctx->stream>>>(L.k, L.binv.p, L.rc.p, L.yc.p);
        PT_LAUNCH_CHECK(ctx);
      }
    } else {
      if (L.exact)
        grid_restrict_kernel<1, false><<<ngrid, 256, 0, ctx->stream>>>(L.g, L.k, cs.lev[l - 1].rc.p, nullptr, 0, L.rc.p, L.yc.p,
                                                                       cs.dpart.p, cdot, cs.ticket.p);
      else
        grid_restrict_kernel<1, true><<<ngrid, 256, 0, ctx->stream>>>(L.g, L.k, cs.lev[l - 1].rc.p, L.binv.p, 0, L.rc.p, L.yc.p,
                                                                      cs.dpart.p, cdot, cs.ticket.p);
      PT_LAUNCH_CHECK(ctx);
    }
    if (L.exact) {
      coarse_dense_kernel<1><<<L.kp / kDenseRows, 256, 0, ctx->stream>>>(L.kp, L.binv.p, 0, L.rc.p, L.yc.p, cs.dpart.p, cdot,
                                                                        cs.ticket.p);
      PT_LAUNCH_CHECK(ctx);
    }
  }
  for (int l = cs.nlev - 2; l >= start; --l) {
    CoarseLevel& L = cs.lev[l];
    const double* ytc = (l + 1 == cs.nlev - 1) ? cs.lev[l + 1].yc.p : cs.lev[l + 1].yt.p;
    grid_prolong_kernel<1><<<ceil_div(L.k, 256), 256, 0, ctx->stream>>>(cs.lev[l + 1].g, L.k, L.yc.p, ytc, L.yt.p);
    PT_LAUNCH_CHECK(ctx);
  }
  return PTFEM_OK;
}

// yt_0 = y_0 + P yt_1 on the finest-grid nodes [a0, b0) only (the rows of this rank read nothing else)
int coarse_prolong_finest_range(ptfem_ctx* ctx, CoarseSpace& cs, int64_t a0, int64_t b0) {
  if (cs.nlev < 2 || b0 <= a0) return PTFEM_OK;
  CoarseLevel& L0 = cs.lev[0];
  const double* ytc = (cs.nlev == 2) ? cs.lev[1].yc.p : cs.lev[1].yt.p;
  grid_prolong_range_kernel<<<ceil_div(b0 - a0, 256), 256, 0, ctx->stream>>>(cs.lev[1].g, a0, b0, L0.yc.p, ytc, L0.yt.p);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

}  // namespace ptfem
