// dist.cu — row-partitioned single solve over several GPUs (BASELINE.json config #5: one refined mesh,
// strong-scaled over 2/4/8 B200s).  Not in the reference (its ElmerSolver run is serial,
// step03_ankle_layers/run_layered_sweep.py:1099); this is the multi-GPU form of the same linear solve.
//
// One process per GPU.  Each rank owns a contiguous block of rows; columns are renumbered
// [0,nloc) owned, [nloc,nloc+nhalo) halo.  Per CG iteration there is exactly one halo exchange
// (ncclSend/ncclRecv pairs on a side stream, overlapped with the SpMV of the rows that need no halo)
// and exactly one all-reduce of three scalars (Chronopoulos-Gear single-reduction CG):
//
//   p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = D^-1 r      (one fused kernel)
//   halo(u)  ||  w[interior] = A u        then  w[boundary] = A u
//   (gamma, delta, rr) = (r.u, w.u, r.r)  -> ncclAllReduce(3 doubles)
//   beta = gamma/gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old)        (one tiny kernel)
//
// With the coarse-grid preconditioner attached (ptfem_dist_coarse_attach) u = D^-1 r + sum_l Z_l B_l Z_l^T r: every rank
// restricts its own rows to the finest grid, the grid vector (k_0 doubles, ~1 MB on the 20 M-tet slab) is summed over
// the ranks - ncclAllReduce, or with peer memory one kernel in which every rank adds up all ranks' buffers in rank
// order - the grid hierarchy above it is replicated, and each rank interpolates back to its rows.
//
// The iteration is captured once into a CUDA graph (NCCL calls included) and replayed, so launch
// latency does not bound the strong-scaled case.  NCCL is loaded with dlopen (the path comes from
// the Python host, which knows where torch's bundled libnccl lives).
#include <dlfcn.h>

#include <cmath>

#include "coarse.cuh"
#include "solver.cuh"

using namespace ptfem;

// ---- minimal NCCL ABI (nccl.h 2.27/2.28: stable since 2.7) ---------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclSum = 0, kNcclFloat64 = 8 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

namespace {

NcclApi* g_nccl = nullptr;

int load_nccl(const char* path, NcclApi** out) {
  if (g_nccl) {
    *out = g_nccl;
    return PTFEM_OK;
  }
  const char* cand[3] = {path && path[0] ? path : nullptr, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (int k = 0; k < 3 && !h; ++k)
    if (cand[k]) h = dlopen(cand[k], RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_err(PTFEM_ERR_NCCL, "cannot load NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
  NcclApi* a = new NcclApi();
  a->handle = h;
#define PT_SYM(field, name)                                                              \
  a->field = reinterpret_cast<decltype(a->field)>(dlsym(h, name));                       \
  if (!a->field) {                                                                       \
    delete a;                                                                            \
    return set_err(PTFEM_ERR_NCCL, "NCCL symbol %s not found", name);                    \
  }
  PT_SYM(GetUniqueId, "ncclGetUniqueId");
  PT_SYM(CommInitRank, "ncclCommInitRank");
  PT_SYM(CommDestroy, "ncclCommDestroy");
  PT_SYM(AllReduce, "ncclAllReduce");
  PT_SYM(Send, "ncclSend");
  PT_SYM(Recv, "ncclRecv");
  PT_SYM(GroupStart, "ncclGroupStart");
  PT_SYM(GroupEnd, "ncclGroupEnd");
  PT_SYM(GetErrorString, "ncclGetErrorString");
#undef PT_SYM
  g_nccl = a;
  *out = a;
  return PTFEM_OK;
}

#define PT_NCCL(api, expr)                                                                        \
  do {                                                                                            \
    ncclResult_t r__ = (expr);                                                                    \
    if (r__ != 0) return set_err(PTFEM_ERR_NCCL, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,    \
                                 (api)->GetErrorString(r__));                                     \
  } while (0)

constexpr int kT = 256;
// device scalars of the distributed solve (doubles): see dist_scalar_kernel
enum { D_GAMMA = 0, D_DELTA, D_RR, D_ALPHA, D_BETA, D_GAMMA_OLD, D_BN2, D_LOC_G, D_LOC_RR, D_COUNT };

__global__ void dist_dinv_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 const double* __restrict__ val, int64_t nloc, double* __restrict__ dinv,
                                 uint8_t* __restrict__ needs_halo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nloc) return;
  double d = 0.0;
  uint8_t h = 0;
  for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
    const int32_t c = col[k];
    if (c == (int32_t)i) d = val[k];
    if (c >= nloc) h = 1;
  }
  dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
  needs_halo[i] = h;
}

__global__ void dist_max_row_kernel(const int32_t* __restrict__ rowptr, int64_t n, int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMax(out, rowptr[i + 1] - rowptr[i]);
}

__global__ void dist_pack_kernel(const double* __restrict__ u, const int32_t* __restrict__ idx, int64_t n,
                                 double* __restrict__ buf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = u[idx[i]];
}

// last CTA: tot[c] = sum_b partial[b*3 + c] with all threads (fixed order: strided per thread, then serial over
// the 85 thread groups) - the single-thread version of this loop was ~40 % of an iteration at 592 CTAs
__device__ __forceinline__ void sum3_partials(const double* __restrict__ partial, unsigned nblocks, double* s_part /*[255]*/,
                                              double tot[3]) {
  const int t = threadIdx.x;
  if (t < 255) {
    const int c = t % 3, g = t / 3;
    double acc = 0.0;
    for (unsigned b = g; b < nblocks; b += 85) acc += __ldcg(partial + (size_t)b * 3 + c);
    s_part[t] = acc;
  }
  __syncthreads();
  for (int c = 0; c < 3; ++c) {
    double a = 0.0;
    for (int g = 0; g < 85; ++g) a += s_part[g * 3 + c];
    tot[c] = a;
  }
}

// first = 1: u = dinv*r only (p = s = 0 beforehand, alpha/beta unused)
__global__ void __launch_bounds__(kT) dist_update_kernel(int64_t n, const double* __restrict__ scal,
                                                         const double* __restrict__ dinv, const double* __restrict__ w,
                                                         double* __restrict__ p, double* __restrict__ s,
                                                         double* __restrict__ x, double* __restrict__ r,
                                                         double* __restrict__ u) {
  const double alpha = scal[D_ALPHA], beta = scal[D_BETA];
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) {
    const double pi = fma(beta, p[i], u[i]);
    const double si = fma(beta, s[i], w[i]);
    const double ri = fma(-alpha, si, r[i]);
    p[i] = pi;
    s[i] = si;
    x[i] = fma(alpha, pi, x[i]);
    r[i] = ri;
    u[i] = ri * dinv[i];
  }
}

__global__ void __launch_bounds__(kT) dist_init_kernel(int64_t n, const double* __restrict__ b,
                                                       const double* __restrict__ dinv, double* __restrict__ p,
                                                       double* __restrict__ s, double* __restrict__ x,
                                                       double* __restrict__ r, double* __restrict__ u) {
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) {
    p[i] = 0.0;
    s[i] = 0.0;
    x[i] = 0.0;
    r[i] = b[i];
    u[i] = b[i] * dinv[i];
  }
}

// local (r.u, w.u, r.r) -> scal[D_GAMMA..D_RR]; per-CTA partials summed in order by the last CTA
__global__ void __launch_bounds__(kT) dist_dots_kernel(int64_t n, const double* __restrict__ r,
                                                       const double* __restrict__ u, const double* __restrict__ w,
                                                       double* __restrict__ partial, double* __restrict__ scal,
                                                       unsigned int* __restrict__ ticket) {
  __shared__ double s_red[3 * (kT / 32)];
  __shared__ int s_last;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) {
    const double ri = r[i], ui = u[i];
    a0 = fma(ri, ui, a0);
    a1 = fma(w[i], ui, a1);
    a2 = fma(ri, ri, a2);
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s_red[wid * 3] = a0;
    s_red[wid * 3 + 1] = a1;
    s_red[wid * 3 + 2] = a2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double acc = 0.0;
    for (int k = 0; k < kT / 32; ++k) acc += s_red[k * 3 + threadIdx.x];
    partial[(size_t)blockIdx.x * 3 + threadIdx.x] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __shared__ double s_part[255];
    __threadfence();
    double tot[3];
    sum3_partials(partial, gridDim.x, s_part, tot);
    if (threadIdx.x < 3) scal[D_GAMMA + threadIdx.x] = tot[threadIdx.x];
  }
}

// after the all-reduce: first -> alpha = gamma/delta, beta = 0 ; else Chronopoulos-Gear recurrences
__global__ void dist_scalar_kernel(double* __restrict__ scal, int first) {
  const double gamma = scal[D_GAMMA], delta = scal[D_DELTA];
  double alpha, beta;
  if (first) {
    beta = 0.0;
    alpha = delta > 0.0 ? gamma / delta : 0.0;
  } else {
    const double g_old = scal[D_GAMMA_OLD], a_old = scal[D_ALPHA];
    beta = g_old > 0.0 ? gamma / g_old : 0.0;
    const double den = a_old != 0.0 ? delta - beta * gamma / a_old : 0.0;
    alpha = den > 0.0 ? gamma / den : 0.0;
  }
  scal[D_ALPHA] = alpha;
  scal[D_BETA] = beta;
  scal[D_GAMMA_OLD] = gamma;
}

// ---- peer-memory path (NVLink P2P through CUDA IPC): no NCCL call inside the iteration ----------------------
// Every rank exports its u vector and a small mailbox.  Per iteration:
//   update kernel   -> last CTA bumps the local iteration counter and PUSHES "u of iteration k is ready" into each
//                      neighbour's mailbox (remote store, release at system scope)
//   pull kernel     -> waits (acquire) for the neighbour's ready flag in the LOCAL mailbox, then loads the halo
//                      entries straight out of the neighbour's u (ld.volatile over NVLink) - no pack, no send/recv
//   SpMV            -> one launch over all rows
//   dots kernel     -> last CTA pushes (gamma, delta, rr, seq) into the mailbox of EVERY rank, waits until its own
//                      mailbox holds all nranks contributions of this iteration, sums them in rank order (identical
//                      on all ranks, deterministic) and computes alpha / beta: all-reduce + scalar kernel in one
// The all-reduce of iteration k is also the barrier that protects u: a neighbour contributes to it only after its
// pull of iteration k, and this rank overwrites u (update k+1) only after the all-reduce completed.
// Mailboxes are double-buffered by iteration parity.  Every wait is bounded; a timeout raises err in the mailbox.
constexpr int kMaxRanks = 16;
constexpr unsigned long long kWaitDefaultNs = 20ull * 1000000000ull;   // bounded waits: wall-clock (globaltimer), then err is raised
struct alignas(16) MailBox {
  double v[3];
  unsigned long long seq;
};
struct Mail {
  unsigned long long iter;                    // local: iterations started (written by the update kernel)
  unsigned long long err;                     // local: a bounded wait timed out
  unsigned long long nred;                    // local: reductions started (sequence number of the mailbox protocol)
  unsigned long long xred;                    // local: coarse-grid sums started (sequence number of the exchange buffers)
  unsigned long long wait_ns[4];              // local: time spent in cross-rank waits: [0] halo flags, [1] scalar all-reduce, [2] grid exchange
  unsigned long long ready_from[kMaxRanks];   // ready_from[q] = k : rank q's u of iteration k is complete (pushed by q)
  unsigned long long xready_from[kMaxRanks];  // xready_from[q] = n : rank q's exchange buffer of coarse sum n is complete
  MailBox box[2][kMaxRanks];                  // box[k & 1][q] : rank q's partial sums of iteration k (pushed by q)
};
struct PeerTable {
  Mail* mail[kMaxRanks];       // every rank's mailbox (own entry = local pointer)
  const double* u[kMaxRanks];  // neighbours' u vectors, indexed by rank (null when not a neighbour)
  const double* xch[kMaxRanks];// every rank's coarse exchange area [2][xstride] (behind its mailbox; null without coarse grids)
  int64_t xstride;
  int64_t k0;                      // finest-grid unknowns: a rank's exchange buffer is [level-0 part | level-1 part]
  int64_t range[kMaxRanks][4];     // per rank {a0, b0, a1, b1}: finest / level-1 grid nodes its rows reach (sharded coarse exchange)
  unsigned long long timeout_ns;   // bound of every cross-rank wait (PTFEM_P2P_TIMEOUT_MS, default 20 s)
  int32_t nbr_rank[kMaxRanks];
  int32_t nnbr, rank, nranks;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_f64(double* p, double v) {
  asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Spin until *p >= want.  The bound is wall-clock time (a descheduled peer - two ranks time-sliced on one GPU, a busy
// host - costs milliseconds, not a spin count).  On a timeout err is raised in the local mailbox AND pushed into every
// peer's, and a wait that finds err already raised returns at once: after the first missed wait no rank spins again,
// the host sees err with the next scalar read-back and abandons the solve (ptfem_dist_solve).
__device__ __noinline__ bool wait_ge(const unsigned long long* p, unsigned long long want, const PeerTable& pt, int kind) {
  if (ld_acquire_sys(p) >= want) return true;
  Mail* me = pt.mail[pt.rank];
  const unsigned long long t0 = globaltimer_ns();
  for (unsigned long long n = 1;; ++n) {
    if (ld_acquire_sys(p) >= want) {
      if (kind >= 0) me->wait_ns[kind] += globaltimer_ns() - t0;   // (accounted by one thread per kernel - CTA 0's: no race)
      return true;
    }
    if ((n & 31) == 0) {
      if (ld_volatile_u64(&me->err)) return false;
      if (globaltimer_ns() - t0 > pt.timeout_ns) break;
    }
    __nanosleep(64);
  }
  atomicExch(&me->err, 1ull);
  for (int q = 0; q < pt.nranks; ++q)
    if (q != pt.rank) st_release_sys(&pt.mail[q]->err, 1ull);
  return false;
}

// Thread 0 of one CTA: all-reduce of three local sums through the mailboxes of all ranks, then the
// Chronopoulos-Gear scalars (or ||b||^2 when bnorm).  Every rank sums in rank order: identical results.
__device__ void mailbox_allreduce(const double loc[3], const PeerTable& pt, double* __restrict__ scal, int first, int bnorm) {
  Mail* me = pt.mail[pt.rank];
  // every rank runs the same sequence of reductions, so a local counter numbers them consistently
  const unsigned long long seq = me->nred + 1;
  me->nred = seq;
  const int par = (int)(seq & 1);
  for (int q = 0; q < pt.nranks; ++q) {
    MailBox* bx = &pt.mail[q]->box[par][pt.rank];
    st_volatile_f64(&bx->v[0], loc[0]);
    st_volatile_f64(&bx->v[1], loc[1]);
    st_volatile_f64(&bx->v[2], loc[2]);
    st_release_sys(&bx->seq, seq);
  }
  double tot[3] = {0.0, 0.0, 0.0};
  for (int q = 0; q < pt.nranks; ++q) {
    MailBox* bx = &me->box[par][q];
    if (!wait_ge(&bx->seq, seq, pt, 1)) break;
    for (int c = 0; c < 3; ++c) tot[c] += ld_volatile_f64(&bx->v[c]);
  }
  if (bnorm) {
    scal[D_BN2] = tot[0];
    return;
  }
  const double gamma = tot[0], delta = tot[1];
  double alpha, beta;
  if (first) {
    beta = 0.0;
    alpha = delta > 0.0 ? gamma / delta : 0.0;
  } else {
    const double g_old = scal[D_GAMMA_OLD], a_old = scal[D_ALPHA];
    beta = g_old > 0.0 ? gamma / g_old : 0.0;
    const double den = a_old != 0.0 ? delta - beta * gamma / a_old : 0.0;
    alpha = den > 0.0 ? gamma / den : 0.0;
  }
  scal[D_GAMMA] = gamma; scal[D_DELTA] = delta; scal[D_RR] = tot[2];
  scal[D_ALPHA] = alpha; scal[D_BETA] = beta; scal[D_GAMMA_OLD] = gamma;
}

// one thread: "u of the next iteration is complete" into every neighbour's mailbox
__device__ __forceinline__ void signal_u_ready(const PeerTable& pt) {
  __threadfence_system();
  Mail* me = pt.mail[pt.rank];
  const unsigned long long k = me->iter + 1;
  me->iter = k;
  for (int j = 0; j < pt.nnbr; ++j) st_release_sys(&pt.mail[pt.nbr_rank[j]]->ready_from[pt.rank], k);
}

// update + "u is ready" signal (signal = 0: u is completed by the coarse-grid kernels that follow, which signal).
// first: x = 0, r = b, u = D^-1 b, p = s = 0.
__global__ void __launch_bounds__(kT) p2p_update_kernel(int64_t n, int first, int signal, const double* __restrict__ b,
                                                        const double* scal /* aliases scal_out */, const double* __restrict__ dinv,
                                                        const double* __restrict__ w, double* __restrict__ p,
                                                        double* __restrict__ s, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ u, PeerTable pt,
                                                        unsigned int* __restrict__ ticket, double* __restrict__ partial,
                                                        double* scal_out) {
  __shared__ int s_last;
  __shared__ double s_red[2 * (kT / 32)];
  const double alpha = scal[D_ALPHA], beta = scal[D_BETA];
  const int64_t stride = (int64_t)gridDim.x * kT;
  double g_acc = 0.0, rr_acc = 0.0;   // r.u and r.r of the updated vectors (consumed by the fused path)
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) {
    if (first) {
      p[i] = 0.0; s[i] = 0.0; x[i] = 0.0;
      r[i] = b[i];
      u[i] = b[i] * dinv[i];
      g_acc = fma(b[i], b[i] * dinv[i], g_acc);
      rr_acc = fma(b[i], b[i], rr_acc);
    } else {
      const double pi = fma(beta, p[i], u[i]);
      const double si = fma(beta, s[i], w[i]);
      const double ri = fma(-alpha, si, r[i]);
      p[i] = pi; s[i] = si;
      x[i] = fma(alpha, pi, x[i]);
      r[i] = ri;
      u[i] = ri * dinv[i];
      g_acc = fma(ri, ri * dinv[i], g_acc);
      rr_acc = fma(ri, ri, rr_acc);
    }
  }
  g_acc = warp_sum(g_acc);
  rr_acc = warp_sum(rr_acc);
  if ((threadIdx.x & 31) == 0) {
    s_red[(threadIdx.x >> 5) * 2] = g_acc;
    s_red[(threadIdx.x >> 5) * 2 + 1] = rr_acc;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int k = 0; k < kT / 32; ++k) a += s_red[k * 2 + threadIdx.x];
    partial[(size_t)blockIdx.x * 2 + threadIdx.x] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {  // fixed-order sum of the per-CTA partials (two values), all threads then one
    __shared__ double s_part[kT];
    const int c = threadIdx.x & 1, g = threadIdx.x >> 1;
    double a = 0.0;
    for (unsigned bk = g; bk < gridDim.x; bk += kT / 2) a += __ldcg(partial + (size_t)bk * 2 + c);
    s_part[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
      for (int gg = 0; gg < kT / 2; ++gg) t += s_part[gg * 2 + threadIdx.x];
      scal_out[D_LOC_G + threadIdx.x] = t;
    }
  }
  if (threadIdx.x == 0 && signal) signal_u_ready(pt);
}

// ---- coarse grids in the partitioned solve -------------------------------------------------------------------
// u += Z y on the owned rows (y = sum of all levels on the finest grid).  With `partial`: also the local sums r.u (u complete
// now) and r.r - per-CTA partials added up in fixed order by the last CTA into scal_out[D_LOC_G], [D_LOC_RR] - so that the
// iteration needs no separate pass over r, u, w for its dot products (w.u comes out of the SpMV's epilogue).  signal: the
// last CTA then announces u to the neighbours (peer memory).
__global__ void __launch_bounds__(kT) dist_zadd_kernel(int64_t n, CoarseDev cd, double* __restrict__ u, const double* __restrict__ r,
                                                       PeerTable pt, int signal, unsigned int* __restrict__ ticket,
                                                       double* __restrict__ partial, double* __restrict__ scal_out) {
  __shared__ int s_last;
  __shared__ double s_red[2 * (kT / 32)];
  const int64_t stride = (int64_t)gridDim.x * kT;
  double g_acc = 0.0, rr_acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) {
    double c0[1];
    coarse_prolong<1, 1>(cd, coarse_row_load(cd.ctab, i), 0, c0);
    const double ui = u[i] + c0[0];
    u[i] = ui;
    if (partial) {
      const double ri = r[i];
      g_acc = fma(ri, ui, g_acc);
      rr_acc = fma(ri, ri, rr_acc);
    }
  }
  if (!signal && !partial) return;
  if (partial) {
    g_acc = warp_sum(g_acc);
    rr_acc = warp_sum(rr_acc);
    if ((threadIdx.x & 31) == 0) {
      s_red[(threadIdx.x >> 5) * 2] = g_acc;
      s_red[(threadIdx.x >> 5) * 2 + 1] = rr_acc;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      double a = 0.0;
      for (int k = 0; k < kT / 32; ++k) a += s_red[k * 2 + threadIdx.x];
      partial[(size_t)blockIdx.x * 2 + threadIdx.x] = a;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (partial) {
    __shared__ double s_part[kT];
    const int c = threadIdx.x & 1, g = threadIdx.x >> 1;
    double a = 0.0;
    for (unsigned bk = g; bk < gridDim.x; bk += kT / 2) a += __ldcg(partial + (size_t)bk * 2 + c);
    s_part[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
      for (int gg = 0; gg < kT / 2; ++gg) t += s_part[gg * 2 + threadIdx.x];
      scal_out[D_LOC_G + threadIdx.x] = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && signal) signal_u_ready(pt);
}

// one thread: this rank's exchange buffer of the next coarse sum is complete -> every other rank's mailbox
__global__ void p2p_coarse_signal_kernel(PeerTable pt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  __threadfence_system();
  Mail* me = pt.mail[pt.rank];
  const unsigned long long seq = me->xred + 1;
  me->xred = seq;
  for (int q = 0; q < pt.nranks; ++q)
    if (q != pt.rank) st_release_sys(&pt.mail[q]->xready_from[pt.rank], seq);
}

// r_c = sum over the ranks (in rank order: identical on all ranks) of their exchange buffers, read straight out of
// the peers' memory; binv: y_c = binv r_c (diagonal-only finest level)
__global__ void __launch_bounds__(kT) p2p_coarse_reduce_kernel(int64_t k, int par, const double* __restrict__ binv,
                                                               double* __restrict__ rc, double* __restrict__ yc, PeerTable pt) {
  Mail* me = pt.mail[pt.rank];
  if (threadIdx.x == 0) {
    const unsigned long long seq = me->xred;   // bumped by the signal kernel just before this launch
    for (int q = 0; q < pt.nranks; ++q)
      if (q != pt.rank && !wait_ge(&me->xready_from[q], seq, pt, blockIdx.x == 0 ? 2 : -1)) break;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t I = (int64_t)blockIdx.x * kT + threadIdx.x; I < k; I += stride) {
    // all loads first (a remote load is microseconds; one after the other they would add up), then the sum in rank order
    double part[kMaxRanks];
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) part[q] = q < pt.nranks ? ld_volatile_f64(pt.xch[q] + (size_t)par * pt.xstride + I) : 0.0;
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) v += part[q];   // + 0.0 for the unused slots changes nothing
    rc[I] = v;
    if (binv) yc[I] = v * __ldg(binv + I);
  }
}

// Sharded form: this rank needs the level-0 sums on ITS nodes [a0, b0) only - contributions of the ranks whose slab
// overlaps (in rank order: both sides of an overlap add the same numbers in the same order) - and the whole level-1
// vector, to which every rank contributes the restriction of its own part on its own range.  Per rank the NVLink reads are a
// few grid planes + the level-1 ranges instead of nranks whole finest-grid vectors.
__global__ void __launch_bounds__(kT) p2p_coarse_reduce_sharded_kernel(int64_t k1, int par, const double* __restrict__ binv0,
                                                                       const double* __restrict__ binv1, double* __restrict__ rc0,
                                                                       double* __restrict__ yc0, double* __restrict__ rc1,
                                                                       double* __restrict__ yc1, PeerTable pt) {
  Mail* me = pt.mail[pt.rank];
  if (threadIdx.x == 0) {
    const unsigned long long seq = me->xred;   // bumped by the signal kernel just before this launch
    for (int q = 0; q < pt.nranks; ++q)
      if (q != pt.rank && !wait_ge(&me->xready_from[q], seq, pt, blockIdx.x == 0 ? 2 : -1)) break;
  }
  __syncthreads();
  const int64_t a0 = pt.range[pt.rank][0], n0 = pt.range[pt.rank][1] - a0;
  const size_t base = (size_t)par * pt.xstride;
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t e = (int64_t)blockIdx.x * kT + threadIdx.x; e < n0 + k1; e += stride) {
    const bool lev1 = e >= n0;
    const int64_t I = lev1 ? e - n0 : a0 + e;
    const size_t off = base + (lev1 ? (size_t)pt.k0 : 0) + (size_t)I;
    double part[kMaxRanks];
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) {
      const bool in = q < pt.nranks && I >= pt.range[q][lev1 ? 2 : 0] && I < pt.range[q][lev1 ? 3 : 1];
      part[q] = in ? ld_volatile_f64(pt.xch[q] + off) : 0.0;
    }
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) v += part[q];
    if (lev1) {
      rc1[I] = v;
      if (binv1) yc1[I] = v * __ldg(binv1 + I);
    } else {
      rc0[I] = v;
      yc0[I] = v * __ldg(binv0 + I);
    }
  }
}

// halo pull: slot h of neighbour j (recv_ptr ranges) = peer u[src[h]]
__global__ void __launch_bounds__(kT) p2p_pull_kernel(int64_t nloc, const int32_t* __restrict__ recv_ptr,
                                                      const int32_t* __restrict__ src, double* __restrict__ u, PeerTable pt) {
  Mail* me = pt.mail[pt.rank];
  const unsigned long long k = me->iter;
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int j = 0; j < pt.nnbr; ++j) {
    const int q = pt.nbr_rank[j];
    if (threadIdx.x == 0) wait_ge(&me->ready_from[q], k, pt, blockIdx.x == 0 ? 0 : -1);
    __syncthreads();
    const double* pu = pt.u[q];
    for (int64_t h = recv_ptr[j] + (int64_t)blockIdx.x * kT + threadIdx.x; h < recv_ptr[j + 1]; h += stride)
      u[nloc + h] = ld_volatile_f64(pu + src[h]);
  }
}

// local (r.u, w.u, r.r); last CTA: all-reduce through the mailboxes + Chronopoulos-Gear scalars
__global__ void __launch_bounds__(kT) p2p_dots_kernel(int64_t n, const double* __restrict__ r, const double* __restrict__ u,
                                                      const double* __restrict__ w, double* __restrict__ partial,
                                                      double* __restrict__ scal, unsigned int* __restrict__ ticket,
                                                      PeerTable pt, int first, int bnorm) {
  __shared__ double s_red[3 * (kT / 32)];
  __shared__ int s_last;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) {
    const double ri = r[i], ui = u[i];
    a0 = fma(ri, ui, a0);
    a1 = fma(w[i], ui, a1);
    a2 = fma(ri, ri, a2);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_red[wid * 3] = a0; s_red[wid * 3 + 1] = a1; s_red[wid * 3 + 2] = a2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double acc = 0.0;
    for (int k = 0; k < kT / 32; ++k) acc += s_red[k * 3 + threadIdx.x];
    partial[(size_t)blockIdx.x * 3 + threadIdx.x] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __shared__ double s_part[255];
  __threadfence();
  double loc[3];
  sum3_partials(partial, gridDim.x, s_part, loc);
  if (threadIdx.x != 0) return;
  mailbox_allreduce(loc, pt, scal, first, bnorm);
}

// fused path: the three local sums already exist (update kernel: r.u, r.r; SpMV epilogue: w.u)
__global__ void p2p_reduce_kernel(double* __restrict__ scal, const double* __restrict__ spmv_scal, PeerTable pt, int first) {
  if (threadIdx.x != 0) return;
  const double loc[3] = {scal[D_LOC_G], spmv_scal[kScalPqOffset], scal[D_LOC_RR]};
  mailbox_allreduce(loc, pt, scal, first, 0);
}

int dist_grid(ptfem_ctx* ctx, int64_t n) {
  int64_t g = (n + kT - 1) / kT;
  const int64_t cap = (int64_t)ctx->sm_count * 4;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

struct DistState {
  DevBuf<double> x, r, u, w, p, s, scal, partial;
  DevBuf<unsigned int> ticket, ticket2;
  DevBuf<int32_t> recv_dummy;
  cudaEvent_t ev_u = nullptr, ev_halo = nullptr;
  cudaGraphExec_t graph = nullptr;
  int graph_iters = 0;
  int64_t graph_launches = 0;
  bool warmed = false;
  int64_t i0 = 0, i1 = 0;  // rows [i0, i1) need no halo value
  int32_t stream_rows = 0, stream_cap = 0;  // streaming SpMV tile geometry valid for any row range
  // peer-memory path
  bool p2p = false;
  bool p2p_broken = false;     // a cross-rank wait timed out: sequence counters are out of step until the ranks reconnect
  DevBuf<Mail> mail;
  DevBuf<int32_t> halo_src, recv_ptr_dev;
  PeerTable pt{};
  std::vector<void*> opened;   // cudaIpcOpenMemHandle mappings to close
  // fused peer-memory SpMV: per halo slot the neighbour address, per neighbour the local ready flag
  DevBuf<const double*> halo_ptr;
  DevBuf<const unsigned long long*> flag_ptr;
  DevBuf<double> partial2;
  PcgWork spmv_work;           // partial / scal / ticket of the SpMV's fused dot
  bool fused = false;
  // coarse-grid preconditioner (ptfem_dist_coarse_attach)
  bool coarse = false;         // attached
  bool use_coarse = false;     // this solve applies it
  bool graph_coarse = false;   // what the captured graph applies
  int64_t xk = 0;              // unknowns of the finest grid = length of one exchange buffer
  int64_t xstride = 0;
  double* xch = nullptr;       // own exchange area [2][xstride] behind the mailbox (peer memory)
  unsigned long long xseq = 0; // coarse sums enqueued so far (host copy of Mail::xred)
  bool sharded = false;        // coarse exchange by grid slab (>= 2 levels, peer memory): level-0 overlap planes + level-1 vector
  int64_t ranges[4] = {0, 0, 0, 0};            // own {a0, b0, a1, b1}
  std::vector<int64_t> all_ranges;             // every rank's, [nranks][4] (ptfem_dist_coarse_ranges_set; empty = whole grids)
};

}  // namespace

struct ptfem_dist_state : DistState {};

static int halo_exchange(ptfem_mesh* m, DistState& d, cudaStream_t st) {
  ptfem_ctx* ctx = m->ctx;
  NcclApi* api = ctx->nccl;
  const int64_t nsend = m->send_ptr.empty() ? 0 : m->send_ptr.back();
  if (nsend > 0) {
    dist_pack_kernel<<<ceil_div(nsend, 256), 256, 0, st>>>(d.u.p, m->send_idx.p, nsend, m->send_buf.p);
    PT_LAUNCH_CHECK(ctx);
  }
  if (m->nnbr > 0) {
    PT_NCCL(api, api->GroupStart());
    for (int k = 0; k < m->nnbr; ++k) {
      const int64_t ns = m->send_ptr[k + 1] - m->send_ptr[k], nr = m->recv_ptr[k + 1] - m->recv_ptr[k];
      if (ns > 0)
        PT_NCCL(api, api->Send(m->send_buf.p + m->send_ptr[k], (size_t)ns, kNcclFloat64, m->nbr_rank[k], (ncclComm_t)ctx->comm, st));
      if (nr > 0)
        PT_NCCL(api, api->Recv(d.u.p + m->nloc + m->recv_ptr[k], (size_t)nr, kNcclFloat64, m->nbr_rank[k], (ncclComm_t)ctx->comm, st));
    }
    PT_NCCL(api, api->GroupEnd());
  }
  return PTFEM_OK;
}

static int spmv_rows(ptfem_mesh* m, DistState& d, int64_t r0, int64_t r1, PcgWork* dot_work = nullptr) {
  if (r1 <= r0) return PTFEM_OK;
  LinSys A;
  A.nn = r1 - r0;
  A.row0 = r0;
  A.nnz = m->nnz;
  A.rowptr = m->rowptr.p;
  A.col = m->col.p;
  A.val = m->val_bc.p;
  A.VS = 1;
  A.S = 1;
  A.stream_rows = d.stream_rows;
  A.stream_cap = d.stream_cap;
  // short ranges (the few boundary rows) are not worth a persistent launch
  const int variant = (d.stream_rows > 0 && r1 - r0 >= 16384) ? PTFEM_SPMV_STREAM : PTFEM_SPMV_VECTOR;
  return spmv_launch(m->ctx, A, variant, d.u.p, d.w.p, dot_work, dot_work != nullptr);   // dot_work: local w.u in the epilogue
}

// w = A u with the halo exchange of u overlapped with the interior rows
static int dist_matvec(ptfem_mesh* m, DistState& d) {
  ptfem_ctx* ctx = m->ctx;
  PT_CK(cudaEventRecord(d.ev_u, ctx->stream));
  PT_CK(cudaStreamWaitEvent(ctx->stream2, d.ev_u, 0));
  PT_TRY(halo_exchange(m, d, ctx->stream2));
  PT_CK(cudaEventRecord(d.ev_halo, ctx->stream2));
  PT_TRY(spmv_rows(m, d, d.i0, d.i1));
  PT_CK(cudaStreamWaitEvent(ctx->stream, d.ev_halo, 0));
  PT_TRY(spmv_rows(m, d, 0, d.i0));
  PT_TRY(spmv_rows(m, d, d.i1, m->nloc));
  return PTFEM_OK;
}

static int dist_reduce(ptfem_mesh* m, DistState& d, int first) {
  ptfem_ctx* ctx = m->ctx;
  NcclApi* api = ctx->nccl;
  dist_dots_kernel<<<dist_grid(ctx, m->nloc), kT, 0, ctx->stream>>>(m->nloc, d.r.p, d.u.p, d.w.p, d.partial.p, d.scal.p, d.ticket.p);
  PT_LAUNCH_CHECK(ctx);
  if (ctx->nranks > 1)
    PT_NCCL(api, api->AllReduce(d.scal.p + D_GAMMA, d.scal.p + D_GAMMA, 3, kNcclFloat64, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
  dist_scalar_kernel<<<1, 1, 0, ctx->stream>>>(d.scal.p, first);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

// u (= D^-1 r on entry) += sum_l Z_l B_l Z_l^T r
static int dist_coarse_add(ptfem_mesh* m, DistState& d) {
  ptfem_ctx* ctx = m->ctx;
  CoarseSpace& cs = *m->coarse;
  CoarseLevel& L0 = cs.lev[0];
  if (d.p2p && d.sharded) {
    const int par = (int)((d.xseq + 1) & 1);
    double* mine = d.xch + (size_t)par * d.xstride;
    // the last CTA of the restriction to level 1 raises "buffer complete" in every other rank's mailbox (no separate launch)
    CoarseSignal sig;
    sig.seq = &d.mail.p->xred;
    sig.ticket = d.ticket2.p;
    for (int q = 0; q < d.pt.nranks; ++q)
      if (q != d.pt.rank) sig.peer_flag[sig.n++] = &d.pt.mail[q]->xready_from[d.pt.rank];
    PT_TRY(coarse_restrict_rows_sharded(ctx, cs, d.r.p, d.ranges, mine, mine + L0.k, sig));
    d.xseq++;
    CoarseLevel& L1 = cs.lev[1];
    const int g = std::min(ceil_div(d.ranges[1] - d.ranges[0] + L1.k, kT), 2 * ctx->sm_count);
    p2p_coarse_reduce_sharded_kernel<<<g, kT, 0, ctx->stream>>>(L1.k, par, L0.binv.p, L1.exact ? nullptr : L1.binv.p, L0.rc.p, L0.yc.p,
                                                                L1.rc.p, L1.yc.p, d.pt);
    PT_LAUNCH_CHECK(ctx);
    PT_TRY(coarse_grids_apply(ctx, cs, !L1.exact, 1));
    PT_TRY(coarse_prolong_finest_range(ctx, cs, d.ranges[0], d.ranges[1]));
  } else if (d.p2p) {
    const int par = (int)((d.xseq + 1) & 1);
    PT_TRY(coarse_restrict_rows(ctx, cs, d.r.p, d.xch + (size_t)par * d.xstride));
    p2p_coarse_signal_kernel<<<1, 32, 0, ctx->stream>>>(d.pt);
    PT_LAUNCH_CHECK(ctx);
    d.xseq++;
    const int g = std::min(ceil_div(L0.k, kT), 2 * ctx->sm_count);
    p2p_coarse_reduce_kernel<<<g, kT, 0, ctx->stream>>>(L0.k, par, L0.exact ? nullptr : L0.binv.p, L0.rc.p, L0.yc.p, d.pt);
    PT_LAUNCH_CHECK(ctx);
    PT_TRY(coarse_grids_apply(ctx, cs, !L0.exact));
  } else {
    PT_TRY(coarse_restrict_rows(ctx, cs, d.r.p, L0.rc.p));
    if (ctx->nranks > 1) {
      NcclApi* api = ctx->nccl;
      PT_NCCL(api, api->AllReduce(L0.rc.p, L0.rc.p, (size_t)L0.k, kNcclFloat64, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
    }
    PT_TRY(coarse_grids_apply(ctx, cs, false));
  }
  dist_zadd_kernel<<<dist_grid(ctx, m->nloc), kT, 0, ctx->stream>>>(m->nloc, coarse_dev(cs), d.u.p, d.r.p, d.pt, d.p2p ? 1 : 0,
                                                                    d.ticket.p, d.p2p ? d.partial2.p : nullptr, d.scal.p);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

static int dist_iteration(ptfem_mesh* m, DistState& d) {
  ptfem_ctx* ctx = m->ctx;
  dist_update_kernel<<<dist_grid(ctx, m->nloc), kT, 0, ctx->stream>>>(m->nloc, d.scal.p, m->dinv.p, d.w.p, d.p.p, d.s.p, d.x.p,
                                                                      d.r.p, d.u.p);
  PT_LAUNCH_CHECK(ctx);
  if (d.use_coarse) PT_TRY(dist_coarse_add(m, d));
  PT_TRY(dist_matvec(m, d));
  return dist_reduce(m, d, 0);
}

// ---- peer-memory iteration --------------------------------------------------------------------------------
static int p2p_matvec_reduce(ptfem_mesh* m, DistState& d, int first) {
  ptfem_ctx* ctx = m->ctx;
  if (m->nhalo > 0) {
    int g = ceil_div(m->nhalo / (m->nnbr > 0 ? m->nnbr : 1) + 1, kT);
    if (g > 64) g = 64;
    p2p_pull_kernel<<<g, kT, 0, ctx->stream>>>(m->nloc, d.recv_ptr_dev.p, d.halo_src.p, d.u.p, d.pt);
    PT_LAUNCH_CHECK(ctx);
  }
  // w = A u with the local w.u in the SpMV's epilogue; r.u and r.r were left by the kernel that completed u (update, or
  // the coarse-grid interpolation): the all-reduce + Chronopoulos-Gear scalars are then one single-thread kernel
  PT_TRY(spmv_rows(m, d, 0, m->nloc, &d.spmv_work));
  p2p_reduce_kernel<<<1, 32, 0, ctx->stream>>>(d.scal.p, d.spmv_work.scal.p, d.pt, first);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}
static int p2p_iteration(ptfem_mesh* m, DistState& d, int first) {
  ptfem_ctx* ctx = m->ctx;
  p2p_update_kernel<<<dist_grid(ctx, m->nloc), kT, 0, ctx->stream>>>(m->nloc, first, d.use_coarse ? 0 : 1, m->b.p, d.scal.p,
                                                                     m->dinv.p, d.w.p, d.p.p, d.s.p, d.x.p, d.r.p, d.u.p, d.pt,
                                                                     d.ticket.p, d.partial2.p, d.scal.p);
  PT_LAUNCH_CHECK(ctx);
  if (d.use_coarse) {
    PT_TRY(dist_coarse_add(m, d));
    return p2p_matvec_reduce(m, d, first);
  }
  if (!d.fused) return p2p_matvec_reduce(m, d, first);
  // one kernel: wait for the neighbours' flags, w = A u with halo entries loaded straight from the neighbours'
  // memory, local w.u in the epilogue; then the single-thread mailbox all-reduce
  LinSys A;
  A.nn = m->nloc;
  A.nnz = m->nnz;
  A.rowptr = m->rowptr.p;
  A.col = m->col.p;
  A.val = m->val_bc.p;
  A.VS = 1;
  A.S = 1;
  A.stream_rows = d.stream_rows;
  A.stream_cap = d.stream_cap;
  A.peer.nloc = m->nloc;
  A.peer.halo = d.halo_ptr.p;
  A.peer.flags = d.flag_ptr.p;
  A.peer.nflags = m->nnbr;
  A.peer.wait = &d.mail.p->iter;
  A.peer.err = &d.mail.p->err;
  PT_TRY(spmv_launch(ctx, A, PTFEM_SPMV_STREAM, d.u.p, d.w.p, &d.spmv_work, true));
  p2p_reduce_kernel<<<1, 32, 0, ctx->stream>>>(d.scal.p, d.spmv_work.scal.p, d.pt, first);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

void ptfem_dist_ctx_release(ptfem_ctx* ctx) {
  if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
}

void ptfem_dist_mesh_release(ptfem_mesh* m) {
  if (!m->dist) return;
  DistState* d = m->dist;
  if (d->graph) cudaGraphExecDestroy(d->graph);
  if (d->ev_u) cudaEventDestroy(d->ev_u);
  if (d->ev_halo) cudaEventDestroy(d->ev_halo);
  for (void* p : d->opened) cudaIpcCloseMemHandle(p);
  delete m->dist;
  m->dist = nullptr;
}

extern "C" {

int ptfem_dist_unique_id(const char* libnccl_path, void* id128) {
  PT_ARG(id128, "null pointer");
  NcclApi* api = nullptr;
  PT_TRY(load_nccl(libnccl_path, &api));
  ncclUniqueId id;
  PT_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return PTFEM_OK;
}

int ptfem_dist_init(ptfem_ctx* ctx, const char* libnccl_path, const void* id128, int32_t rank, int32_t nranks) {
  PT_ARG(ctx, "null pointer");
  PT_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  PT_CK(cudaSetDevice(ctx->device));
  if (!id128) {  // peer-memory only: no NCCL communicator
    ctx->rank = rank;
    ctx->nranks = nranks;
    return PTFEM_OK;
  }
  NcclApi* api = nullptr;
  PT_TRY(load_nccl(libnccl_path, &api));
  if (ctx->comm) return set_err(PTFEM_ERR_STATE, "context already has a communicator");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t comm = nullptr;
  PT_NCCL(api, api->CommInitRank(&comm, nranks, id, rank));
  ctx->nccl = api;
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return PTFEM_OK;
}

int ptfem_dist_finalize(ptfem_ctx* ctx) {
  PT_ARG(ctx, "null context");
  PT_CK(cudaSetDevice(ctx->device));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream2));
  ptfem_dist_ctx_release(ctx);
  ctx->rank = 0;
  ctx->nranks = 1;
  return PTFEM_OK;
}

int ptfem_dist_system_create(ptfem_ctx* ctx, int64_t nloc, int64_t nhalo, const int32_t* rowptr, const int32_t* col,
                             const double* val, const double* b, int32_t nnbr, const int32_t* nbr_rank,
                             const int32_t* send_ptr, const int32_t* send_idx, const int32_t* recv_ptr, ptfem_mesh** out) {
  PT_ARG(ctx && out && rowptr && col && val && b, "null pointer");
  PT_ARG(nloc > 0 && nhalo >= 0 && nnbr >= 0, "bad sizes");
  PT_ARG(nnbr == 0 || (nbr_rank && send_ptr && send_idx && recv_ptr), "null neighbour arrays");
  if (nnbr > 0 && ctx->nranks < 2) return set_err(PTFEM_ERR_STATE, "ptfem_dist_init has not been called on this context");
  *out = nullptr;
  const int64_t nnz = rowptr[nloc];
  for (int64_t k = 0; k < nnz; ++k)
    if (col[k] < 0 || col[k] >= nloc + nhalo) return set_err(PTFEM_ERR_ARG, "column %d out of range at entry %lld", col[k], (long long)k);
  if (nnbr > 0) {
    if (recv_ptr[nnbr] != nhalo) return set_err(PTFEM_ERR_ARG, "recv_ptr[nnbr] (%d) != nhalo (%lld)", recv_ptr[nnbr], (long long)nhalo);
    for (int64_t k = 0; k < send_ptr[nnbr]; ++k)
      if (send_idx[k] < 0 || send_idx[k] >= nloc) return set_err(PTFEM_ERR_ARG, "send index out of range");
    for (int k = 0; k < nnbr; ++k)
      if (nbr_rank[k] < 0 || nbr_rank[k] >= ctx->nranks || nbr_rank[k] == ctx->rank)
        return set_err(PTFEM_ERR_ARG, "bad neighbour rank %d", nbr_rank[k]);
  } else if (nhalo != 0) {
    return set_err(PTFEM_ERR_ARG, "halo columns without neighbours");
  }
  PT_CK(cudaSetDevice(ctx->device));
  ptfem_mesh* m = new ptfem_mesh();
  m->ctx = ctx;
  m->is_dist = true;
  m->nn = nloc;
  m->nloc = nloc;
  m->nhalo = nhalo;
  m->nnz = nnz;
  m->nnbr = nnbr;
  m->nbr_rank.assign(nbr_rank, nbr_rank + nnbr);
  m->send_ptr.assign(send_ptr, send_ptr + (nnbr > 0 ? nnbr + 1 : 0));
  m->recv_ptr.assign(recv_ptr, recv_ptr + (nnbr > 0 ? nnbr + 1 : 0));
  m->dist = new ptfem_dist_state();
  DistState& d = *m->dist;
  int rc = PTFEM_OK;
  const int64_t nsend = nnbr > 0 ? send_ptr[nnbr] : 0;
  auto body = [&]() -> int {
    PT_TRY(m->rowptr.alloc(nloc + 1));
    PT_TRY(m->col.alloc(nnz + 8));
    PT_TRY(m->val_bc.alloc(nnz + 8));
    PT_TRY(m->b.alloc(nloc));
    PT_TRY(m->dinv.alloc(nloc));
    PT_TRY(m->send_idx.alloc(nsend));
    PT_TRY(m->send_buf.alloc(nsend));
    PT_CK(cudaMemcpyAsync(m->rowptr.p, rowptr, (nloc + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaMemcpyAsync(m->col.p, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaMemcpyAsync(m->val_bc.p, val, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaMemcpyAsync(m->b.p, b, nloc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (nsend > 0) PT_CK(cudaMemcpyAsync(m->send_idx.p, send_idx, nsend * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t nv = nloc + nhalo;
    PT_TRY(d.x.alloc(nloc));
    PT_TRY(d.r.alloc(nloc));
    PT_TRY(d.u.alloc(nv));
    PT_TRY(d.w.alloc(nloc));
    PT_TRY(d.p.alloc(nloc));
    PT_TRY(d.s.alloc(nloc));
    PT_TRY(d.scal.alloc(D_COUNT));
    PT_TRY(d.partial.alloc((size_t)ctx->sm_count * 4 * 3));
    PT_TRY(d.ticket.alloc(1));
    PT_CK(cudaMemsetAsync(d.ticket.p, 0, sizeof(unsigned int), ctx->stream));
    PT_TRY(d.ticket2.alloc(1));
    PT_CK(cudaMemsetAsync(d.ticket2.p, 0, sizeof(unsigned int), ctx->stream));
    PT_CK(cudaMemsetAsync(d.u.p, 0, nv * sizeof(double), ctx->stream));
    PT_CK(cudaMemsetAsync(d.scal.p, 0, D_COUNT * sizeof(double), ctx->stream));
    PT_CK(cudaEventCreateWithFlags(&d.ev_u, cudaEventDisableTiming));
    PT_CK(cudaEventCreateWithFlags(&d.ev_halo, cudaEventDisableTiming));
    DevBuf<uint8_t> flag;
    PT_TRY(flag.alloc(nloc));
    dist_dinv_kernel<<<ceil_div(nloc, 128), 128, 0, ctx->stream>>>(m->rowptr.p, m->col.p, m->val_bc.p, nloc, m->dinv.p, flag.p);
    PT_LAUNCH_CHECK(ctx);
    std::vector<uint8_t> hflag(nloc);
    PT_CK(cudaMemcpyAsync(hflag.data(), flag.p, nloc, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    // longest run of rows that touch no halo column
    int64_t best0 = 0, best1 = 0, run0 = 0;
    for (int64_t i = 0; i <= nloc; ++i) {
      if (i == nloc || hflag[i]) {
        if (i - run0 > best1 - best0) {
          best0 = run0;
          best1 = i;
        }
        run0 = i + 1;
      }
    }
    d.i0 = best0;
    d.i1 = best1;
    // streaming SpMV on arbitrary row ranges: a stage must hold any 64 consecutive rows
    {
      DevBuf<int32_t> mr;
      PT_TRY(mr.alloc(1));
      PT_CK(cudaMemsetAsync(mr.p, 0, sizeof(int32_t), ctx->stream));
      dist_max_row_kernel<<<ceil_div(nloc, 256), 256, 0, ctx->stream>>>(m->rowptr.p, nloc, mr.p);
      PT_LAUNCH_CHECK(ctx);
      int32_t mx = 0;
      PT_CK(cudaMemcpyAsync(&mx, mr.p, sizeof mx, cudaMemcpyDeviceToHost, ctx->stream));
      PT_CK(cudaStreamSynchronize(ctx->stream));
      const int cap = ((64 * mx + 8) + 31) & ~31;
      if (mx > 0 && cap <= 4096) {
        d.stream_rows = 64;
        d.stream_cap = cap;
      }
    }
    // rows outside [i0,i1) may or may not need the halo; they all wait for it (correct, slightly conservative)
    return PTFEM_OK;
  };
  rc = body();
  if (rc) {
    ptfem_dist_mesh_release(m);
    delete m;
    return rc;
  }
  *out = m;
  return PTFEM_OK;
}

int ptfem_dist_solve(ptfem_mesh* m, const ptfem_solve_opts* opts, double* x_local, ptfem_solve_stats* stats, double* ms_spmv,
                     double* ms_halo, double* ms_allreduce) {
  PT_ARG(m && m->is_dist && m->dist, "not a distributed system");
  ptfem_ctx* ctx = m->ctx;
  NcclApi* api = ctx->nccl;
  PT_CK(cudaSetDevice(ctx->device));
  ptfem_solve_opts o;
  if (opts) o = *opts; else ptfem_solve_opts_default(&o);
  PT_ARG(o.rtol > 0.0 && o.maxit > 0, "rtol and maxit must be positive");
  DistState& d = *m->dist;
  if (o.precond == PTFEM_PRECOND_AUTO) o.precond = d.coarse ? PTFEM_PRECOND_TWOLEVEL : PTFEM_PRECOND_JACOBI;
  if (o.precond != PTFEM_PRECOND_JACOBI && o.precond != PTFEM_PRECOND_TWOLEVEL)
    return set_err(PTFEM_ERR_ARG, "the row-partitioned solve supports the Jacobi and the coarse-grid preconditioner only");
  if (o.precond == PTFEM_PRECOND_TWOLEVEL && !d.coarse)
    return set_err(PTFEM_ERR_STATE, "coarse-grid preconditioner requested but ptfem_dist_coarse_attach has not been called");
  d.use_coarse = o.precond == PTFEM_PRECOND_TWOLEVEL;
  if (d.use_coarse && d.p2p && !d.xch)
    return set_err(PTFEM_ERR_STATE, "peer memory was exported before the coarse grids were attached (no exchange area)");
  if (d.graph && d.graph_coarse != d.use_coarse) {
    cudaGraphExecDestroy(d.graph);
    d.graph = nullptr;
  }
  // the exchange buffers alternate with the parity of the coarse-sum count and the captured graph assumes it starts on
  // an even count after the initial application: check_every is kept even, every solve starts on an even count
  int check = o.check_every > 0 ? o.check_every : (d.use_coarse ? 10 : 50);
  if (d.use_coarse && (check & 1)) ++check;
  if (d.use_coarse && d.p2p && (d.xseq & 1)) {
    p2p_coarse_signal_kernel<<<1, 32, 0, ctx->stream>>>(d.pt);
    PT_LAUNCH_CHECK(ctx);
    d.xseq++;
  }
  const int grid = dist_grid(ctx, m->nloc);
  double* h = ctx->h_pinned;

  const bool p2p = d.p2p;
  if (p2p && d.p2p_broken)
    return set_err(PTFEM_ERR_STATE, "peer-memory connection is out of step after a timed-out wait: build a new system and reconnect, "
                                    "or use the NCCL transport");
  if (ctx->nranks > 1 && !p2p && !ctx->comm)
    return set_err(PTFEM_ERR_STATE, "neither an NCCL communicator (ptfem_dist_init) nor peer memory (ptfem_dist_p2p_connect) is set up");
  // first use of a peer connection / collective sets up NCCL channels (seconds): keep it out of the timing
  if (ctx->nranks > 1 && !p2p && !d.warmed) {
    PT_TRY(halo_exchange(m, d, ctx->stream));
    PT_NCCL(api, api->AllReduce(d.partial.p, d.partial.p, 3, kNcclFloat64, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    d.warmed = true;
  }
  if (p2p) PT_CK(cudaMemsetAsync(d.mail.p->wait_ns, 0, sizeof(d.mail.p->wait_ns), ctx->stream));
  // the timing events go with the scope: every early return below (a failed launch, a timed-out wait, divergence) releases them
  struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } evp;
  PT_CK(cudaEventCreate(&evp.a));
  PT_CK(cudaEventCreate(&evp.b));
  const cudaEvent_t e0 = evp.a, e1 = evp.b;
  PT_CK(cudaEventRecord(e0, ctx->stream));
  if (p2p) {
    // ||b||^2 through the mailboxes, then iteration 0 (x = 0, r = b, u = D^-1 r, w = A u, first scalars)
    p2p_dots_kernel<<<grid, kT, 0, ctx->stream>>>(m->nloc, m->b.p, m->b.p, m->b.p, d.partial.p, d.scal.p, d.ticket.p, d.pt, 0, 1);
    PT_LAUNCH_CHECK(ctx);
    PT_TRY(p2p_iteration(m, d, 1));
  } else {
    // ||b||^2 via the dots kernel (r = u = b): gamma slot
    dist_dots_kernel<<<grid, kT, 0, ctx->stream>>>(m->nloc, m->b.p, m->b.p, m->b.p, d.partial.p, d.scal.p, d.ticket.p);
    PT_LAUNCH_CHECK(ctx);
    if (ctx->nranks > 1)
      PT_NCCL(api, api->AllReduce(d.scal.p + D_GAMMA, d.scal.p + D_BN2, 1, kNcclFloat64, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
    else
      PT_CK(cudaMemcpyAsync(d.scal.p + D_BN2, d.scal.p + D_GAMMA, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    // x = 0, r = b, u = D^-1 r, w = A u, first scalars
    dist_init_kernel<<<grid, kT, 0, ctx->stream>>>(m->nloc, m->b.p, m->dinv.p, d.p.p, d.s.p, d.x.p, d.r.p, d.u.p);
    PT_LAUNCH_CHECK(ctx);
    if (d.use_coarse) PT_TRY(dist_coarse_add(m, d));
    PT_TRY(dist_matvec(m, d));
    PT_TRY(dist_reduce(m, d, 1));
  }

  // scalars, and with them the mailbox's error flag: a timed-out wait ends the solve at the next read-back
  unsigned long long* h_err = reinterpret_cast<unsigned long long*>(h + D_COUNT);
  *h_err = 0;
  auto read_scal = [&]() -> int {
    PT_CK(cudaMemcpyAsync(h, d.scal.p, D_COUNT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (p2p) PT_CK(cudaMemcpyAsync(h_err, &d.mail.p->err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    if (p2p && *h_err) {
      d.p2p_broken = true;
      return set_err(PTFEM_ERR_STATE, "peer-memory solve: a cross-rank wait timed out (on this rank or a peer); the connection is out "
                                      "of step - reconnect (ptfem_dist_p2p_export / _connect on a new system) or use the NCCL transport");
    }
    return PTFEM_OK;
  };
  PT_TRY(read_scal());
  const double bn2 = h[D_BN2];
  const double rtol2 = o.rtol * o.rtol;
  int it = 0;
  bool converged = bn2 <= 0.0 || h[D_RR] <= rtol2 * bn2;
  double rel = bn2 > 0.0 ? sqrt(h[D_RR] / bn2) : 0.0;
  while (!converged && it < o.maxit) {
    const int n_it = (o.maxit - it) < check ? (o.maxit - it) : check;
    if (o.use_graph && n_it == check) {
      if (!d.graph || d.graph_iters != check) {
        if (d.graph) cudaGraphExecDestroy(d.graph);
        d.graph = nullptr;
        cudaGraph_t g = nullptr;
        int rc = PTFEM_OK;
        PT_CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int64_t l0 = ctx->launches;
        const unsigned long long x0 = d.xseq;
        for (int k = 0; k < check && rc == PTFEM_OK; ++k) rc = p2p ? p2p_iteration(m, d, 0) : dist_iteration(m, d);
        d.graph_launches = ctx->launches - l0;
        ctx->launches = l0;
        d.xseq = x0;            // captured, not run
        d.graph_coarse = d.use_coarse;
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);   // always: an error above must not leave the stream capturing
        if (rc) {
          if (g) cudaGraphDestroy(g);
          return rc;
        }
        if (ce != cudaSuccess) return set_err(PTFEM_ERR_CUDA, "graph capture of the distributed iteration failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&d.graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return set_err(PTFEM_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
        d.graph_iters = check;
      }
      PT_CK(cudaGraphLaunch(d.graph, ctx->stream));
      ctx->launches += d.graph_launches;
      if (d.use_coarse && p2p) d.xseq += (unsigned long long)check;
    } else {
      for (int k = 0; k < n_it; ++k) PT_TRY(p2p ? p2p_iteration(m, d, 0) : dist_iteration(m, d));
    }
    it += n_it;
    PT_TRY(read_scal());
    const double rr = h[D_RR];
    if (!(rr == rr)) return set_err(PTFEM_ERR_NOCONV, "distributed PCG diverged after %d iterations", it);
    rel = sqrt(rr / bn2);
    converged = rr <= rtol2 * bn2;
  }
  // the recurrences have advanced r/u one step beyond x: apply the pending update so x matches r
  // (x_{i+1} uses alpha_i p_i, done inside the update kernel of the next iteration)
  dist_update_kernel<<<grid, kT, 0, ctx->stream>>>(m->nloc, d.scal.p, m->dinv.p, d.w.p, d.p.p, d.s.p, d.x.p, d.r.p, d.u.p);
  PT_LAUNCH_CHECK(ctx);
  PT_CK(cudaEventRecord(e1, ctx->stream));
  PT_CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);

  // phase timings, each phase alone (no overlap), averaged over a few repetitions
  const int reps = 20;
  float t_spmv = 0.f, t_halo = 0.f, t_ar = 0.f;
  {
    PT_CK(cudaEventRecord(e0, ctx->stream));
    for (int k = 0; k < reps; ++k) PT_TRY(spmv_rows(m, d, 0, m->nloc));
    PT_CK(cudaEventRecord(e1, ctx->stream));
    PT_CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&t_spmv, e0, e1);
    if (!p2p) {   // (the peer-memory phases are not separate launches; they are part of ms_per_iteration)
      PT_CK(cudaEventRecord(e0, ctx->stream));
      for (int k = 0; k < reps; ++k) PT_TRY(halo_exchange(m, d, ctx->stream));
      PT_CK(cudaEventRecord(e1, ctx->stream));
      PT_CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&t_halo, e0, e1);
      PT_CK(cudaEventRecord(e0, ctx->stream));
      if (ctx->nranks > 1)
        for (int k = 0; k < reps; ++k)
          PT_NCCL(api, api->AllReduce(d.partial.p, d.partial.p, 3, kNcclFloat64, kNcclSum, (ncclComm_t)ctx->comm, ctx->stream));
      PT_CK(cudaEventRecord(e1, ctx->stream));
      PT_CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&t_ar, e0, e1);
    }
  }
  if (p2p) {
    // time the iteration spent waiting for other ranks (globaltimer inside the waiting kernels), per iteration: the peer-memory
    // transport has no separate halo / all-reduce launches to time - what it has are these waits
    unsigned long long wn[4] = {0, 0, 0, 0};
    PT_CK(cudaMemcpyAsync(wn, d.mail.p->wait_ns, sizeof wn, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    const double per = 1e-6 / (double)(it > 0 ? it : 1);
    t_halo = (float)(wn[0] * per) * reps;            // (reported as t / reps below)
    t_ar = (float)((wn[1] + wn[2]) * per) * reps;
  }
  if (p2p) PT_TRY(read_scal());   // a wait of the last chunk may have timed out after the last read-back
  if (ms_spmv) *ms_spmv = t_spmv / reps;
  if (ms_halo) *ms_halo = t_halo / reps;
  if (ms_allreduce) *ms_allreduce = t_ar / reps;
  if (x_local) {
    PT_CK(cudaMemcpyAsync(x_local, d.x.p, m->nloc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
  }
  if (stats) {
    stats->iterations = it;
    stats->converged = converged ? 1 : 0;
    stats->nsys = 1;
    stats->spmv_calls = it + 1;
    stats->rel_residual = rel;
    stats->true_rel_residual = rel;
    stats->solve_ms = ms;
    stats->spmv_ms = t_spmv / reps;
    stats->setup_ms = 0.0;
    stats->precond = o.precond;
    stats->coarse_unknowns = d.use_coarse ? (int32_t)m->coarse->lev[m->coarse->nlev - 1].k : 0;
  }
  if (!converged) return set_err(PTFEM_ERR_NOCONV, "distributed PCG did not reach rtol=%g in %d iterations (rel. residual %.3e)", o.rtol, it, rel);
  return PTFEM_OK;
}

int ptfem_dist_coarse_attach(ptfem_mesh* m, ptfem_mesh* full, int64_t row0) {
  PT_ARG(m && m->is_dist && m->dist, "not a distributed system");
  PT_ARG(full && !full->is_dist && full->ctx == m->ctx, "the replica must be a mesh of the same context");
  PT_CK(cudaSetDevice(m->ctx->device));
  DistState& d = *m->dist;
  if (d.mail.p) return set_err(PTFEM_ERR_STATE, "attach the coarse grids before ptfem_dist_p2p_export (the exchange area is part of the export)");
  PT_TRY(coarse_replica_prepare(full));
  PT_TRY(coarse_attach_rows(m, full, row0));
  d.coarse = true;
  d.xk = m->coarse->lev[0].k;
  // sharded exchange (>= 2 levels): the grid nodes this rank's rows reach; other ranks' ranges arrive through
  // ptfem_dist_coarse_ranges_set (without them every rank is taken to reach the whole grids: correct, just not sharded)
  d.sharded = m->coarse->nlev >= 2 && !m->coarse->lev[0].exact;
  PT_TRY(coarse_touched_ranges(m->ctx, *m->coarse, m->nloc, d.ranges));
  d.all_ranges.clear();
  if (d.graph) cudaGraphExecDestroy(d.graph);
  d.graph = nullptr;
  return PTFEM_OK;
}

int ptfem_dist_coarse_ranges_get(ptfem_mesh* m, int64_t ranges4[4]) {
  PT_ARG(m && m->is_dist && m->dist && ranges4, "not a distributed system");
  DistState& d = *m->dist;
  if (!d.coarse) return set_err(PTFEM_ERR_STATE, "ptfem_dist_coarse_attach has not been called");
  for (int k = 0; k < 4; ++k) ranges4[k] = d.ranges[k];
  return PTFEM_OK;
}

int ptfem_dist_coarse_ranges_set(ptfem_mesh* m, int32_t nranks, const int64_t* all_ranges) {
  PT_ARG(m && m->is_dist && m->dist && all_ranges, "not a distributed system");
  PT_ARG(nranks == m->ctx->nranks && nranks <= kMaxRanks, "rank count mismatch");
  DistState& d = *m->dist;
  if (!d.coarse) return set_err(PTFEM_ERR_STATE, "ptfem_dist_coarse_attach has not been called");
  if (d.p2p) return set_err(PTFEM_ERR_STATE, "set the ranges before ptfem_dist_p2p_connect");
  for (int q = 0; q < nranks; ++q) {
    const int64_t* r = all_ranges + 4 * q;
    if (r[0] < 0 || r[1] < r[0] || r[1] > d.xk || r[2] < 0 || r[3] < r[2]) return set_err(PTFEM_ERR_ARG, "bad range of rank %d", q);
  }
  for (int k = 0; k < 4; ++k)
    if (all_ranges[4 * m->ctx->rank + k] != d.ranges[k]) return set_err(PTFEM_ERR_ARG, "own range differs from ptfem_dist_coarse_ranges_get");
  d.all_ranges.assign(all_ranges, all_ranges + 4 * nranks);
  return PTFEM_OK;
}

int ptfem_dist_p2p_export(ptfem_mesh* m, void* handles128) {
  PT_ARG(m && m->is_dist && m->dist && handles128, "not a distributed system");
  PT_CK(cudaSetDevice(m->ctx->device));
  DistState& d = *m->dist;
  if (!d.mail.p) {
    // one allocation = one IPC handle: the coarse exchange area [2][xstride] lives behind the mailbox
    size_t nmail = 1;
    if (d.coarse) {
      d.xstride = (d.xk + (d.sharded ? m->coarse->lev[1].k : 0) + 15) & ~(int64_t)15;
      nmail += (2 * (size_t)d.xstride * sizeof(double) + sizeof(Mail) - 1) / sizeof(Mail);
    }
    PT_TRY(d.mail.alloc(nmail));
    PT_CK(cudaMemsetAsync(d.mail.p, 0, nmail * sizeof(Mail), m->ctx->stream));
    PT_CK(cudaStreamSynchronize(m->ctx->stream));
    d.xch = d.coarse ? reinterpret_cast<double*>(d.mail.p + 1) : nullptr;
    d.xseq = 0;
  }
  cudaIpcMemHandle_t h[2];
  PT_CK(cudaIpcGetMemHandle(&h[0], d.u.p));
  PT_CK(cudaIpcGetMemHandle(&h[1], d.mail.p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handles128, h, 128);
  return PTFEM_OK;
}

int ptfem_dist_p2p_connect(ptfem_mesh* m, int32_t nranks, const void* all_handles, const int32_t* halo_src) {
  PT_ARG(m && m->is_dist && m->dist && all_handles, "not a distributed system");
  ptfem_ctx* ctx = m->ctx;
  PT_ARG(nranks == ctx->nranks && nranks <= kMaxRanks, "rank count mismatch (at most 16 ranks)");
  PT_ARG(m->nhalo == 0 || halo_src, "null halo source list");
  PT_CK(cudaSetDevice(ctx->device));
  DistState& d = *m->dist;
  if (!d.mail.p) return set_err(PTFEM_ERR_STATE, "ptfem_dist_p2p_export must be called first");
  PeerTable pt;
  memset(&pt, 0, sizeof pt);
  pt.rank = ctx->rank;
  pt.nranks = nranks;
  pt.nnbr = m->nnbr;
  pt.xstride = d.xstride;
  pt.k0 = d.xk;
  for (int q = 0; q < kMaxRanks; ++q) {
    const bool have = (int)d.all_ranges.size() == 4 * nranks && q < nranks;
    pt.range[q][0] = have ? d.all_ranges[4 * q] : 0;
    pt.range[q][1] = have ? d.all_ranges[4 * q + 1] : d.xk;
    pt.range[q][2] = have ? d.all_ranges[4 * q + 2] : 0;
    pt.range[q][3] = have ? d.all_ranges[4 * q + 3] : (d.sharded ? m->coarse->lev[1].k : 0);
  }
  if ((int)d.all_ranges.size() != 4 * nranks && d.sharded) {   // own ranges too: everybody reaches everything
    d.ranges[0] = 0; d.ranges[1] = d.xk; d.ranges[2] = 0; d.ranges[3] = m->coarse->lev[1].k;
  }
  pt.timeout_ns = kWaitDefaultNs;
  if (const char* e = getenv("PTFEM_P2P_TIMEOUT_MS")) {
    const double ms = atof(e);
    if (ms > 0.0) pt.timeout_ns = (unsigned long long)(ms * 1e6);
  }
  if (m->nnbr > kMaxRanks) return set_err(PTFEM_ERR_ARG, "too many neighbours");
  for (int k = 0; k < m->nnbr; ++k) pt.nbr_rank[k] = m->nbr_rank[k];
  const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int q = 0; q < nranks; ++q) {
    if (q == ctx->rank) {
      pt.mail[q] = d.mail.p;
      pt.u[q] = d.u.p;
      pt.xch[q] = d.xch;
      continue;
    }
    void* pm = nullptr;
    PT_CK(cudaIpcOpenMemHandle(&pm, hs[2 * q + 1], cudaIpcMemLazyEnablePeerAccess));
    d.opened.push_back(pm);
    pt.mail[q] = reinterpret_cast<Mail*>(pm);
    pt.xch[q] = d.xch ? reinterpret_cast<const double*>(pt.mail[q] + 1) : nullptr;   // every rank attaches the same grids
    bool is_nbr = false;
    for (int k = 0; k < m->nnbr; ++k) is_nbr = is_nbr || m->nbr_rank[k] == q;
    if (is_nbr) {
      void* pu = nullptr;
      PT_CK(cudaIpcOpenMemHandle(&pu, hs[2 * q], cudaIpcMemLazyEnablePeerAccess));
      d.opened.push_back(pu);
      pt.u[q] = reinterpret_cast<const double*>(pu);
    }
  }
  d.pt = pt;
  PT_TRY(d.halo_src.alloc(m->nhalo));
  PT_TRY(d.recv_ptr_dev.alloc(m->nnbr + 1));
  if (m->nhalo > 0)
    PT_CK(cudaMemcpyAsync(d.halo_src.p, halo_src, m->nhalo * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (m->nnbr > 0)
    PT_CK(cudaMemcpyAsync(d.recv_ptr_dev.p, m->recv_ptr.data(), (m->nnbr + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  else
    PT_CK(cudaMemsetAsync(d.recv_ptr_dev.p, 0, sizeof(int32_t), ctx->stream));
  // tables of the fused SpMV
  PT_TRY(d.partial2.alloc((size_t)ctx->sm_count * 4 * 2));
  PT_TRY(d.spmv_work.partial.alloc((size_t)ctx->sm_count * 32 * 2));
  PT_TRY(d.spmv_work.scal.alloc(8 * 16));
  PT_TRY(d.spmv_work.ticket.alloc(4));
  PT_CK(cudaMemsetAsync(d.spmv_work.ticket.p, 0, 4 * sizeof(unsigned int), ctx->stream));
  PT_CK(cudaMemsetAsync(d.spmv_work.scal.p, 0, 8 * 16 * sizeof(double), ctx->stream));
  {
    std::vector<const double*> hp((size_t)m->nhalo);
    for (int k = 0; k < m->nnbr; ++k)
      for (int32_t hh = m->recv_ptr[k]; hh < m->recv_ptr[k + 1]; ++hh) hp[hh] = pt.u[m->nbr_rank[k]] + halo_src[hh];
    std::vector<const unsigned long long*> fl((size_t)m->nnbr);
    for (int k = 0; k < m->nnbr; ++k) fl[k] = &d.mail.p->ready_from[m->nbr_rank[k]];
    PT_TRY(d.halo_ptr.alloc(hp.size()));
    PT_TRY(d.flag_ptr.alloc(fl.size()));
    if (!hp.empty()) PT_CK(cudaMemcpyAsync(d.halo_ptr.p, hp.data(), hp.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    if (!fl.empty()) PT_CK(cudaMemcpyAsync(d.flag_ptr.p, fl.data(), fl.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
  }
  d.fused = d.stream_rows > 0 && ctx->tune_p2p_fused;
  PT_CK(cudaStreamSynchronize(ctx->stream));
  d.p2p = true;
  if (d.graph) cudaGraphExecDestroy(d.graph);
  d.graph = nullptr;
  return PTFEM_OK;
}

}  // extern "C"
