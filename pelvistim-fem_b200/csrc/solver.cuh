// solver.cuh — internal interface of the SpMV / PCG layer (solver.cu).
#pragma once
#include "common.cuh"
#include "coarse.cuh"

// Arguments of the fused peer-memory gather of the streaming SpMV (see dist.cu).
struct PeerGather {
  int64_t nloc = 0;                                     // 0 = disabled
  const double* const* halo = nullptr;                  // [nhalo] device pointers (neighbour memory)
  const unsigned long long* const* flags = nullptr;     // [nflags] local flags written remotely by the neighbours
  int nflags = 0;
  const unsigned long long* wait = nullptr;             // value the flags must reach (local iteration counter)
  unsigned long long* err = nullptr;                    // raised when a wait times out
};

// A (batch of) linear system(s) on one CSR pattern, all device pointers.
struct LinSys {
  int64_t nn = 0, nnz = 0;
  const int32_t* rowptr = nullptr;
  const int32_t* col = nullptr;
  const double* val = nullptr;   // [nnz][VS]
  int VS = 1;                    // value sets: 1 (shared matrix) or S
  int S = 1;                     // systems (padded to 1,2,4,8,16); vectors are [nn][S]
  const double* dinv = nullptr;  // [nn][VS] inverse diagonal
  const double* b = nullptr;     // [nn][S]
  int32_t stream_rows = 0;       // rows per tile of the streaming SpMV (0: vector kernel only)
  int32_t stream_cap = 0;        // staged entries per tile (shared-memory stage size / 12 bytes)
  // optional processing-order view for the streaming kernel: row j of (prowptr, pcol, pval) is mesh row rowid[j]
  const int32_t* rowid = nullptr;
  const int32_t* prowptr = nullptr;
  const int32_t* pcol = nullptr;
  const double* pval = nullptr;
  // optional even-padded copy for the multi-RHS streaming kernel (natural row order)
  const int32_t* qrowptr = nullptr;
  const int32_t* qcol = nullptr;
  const double* qval = nullptr;
  int32_t q_rows = 0, q_cap = 0;
  // peer-memory gathers (row-partitioned solve, S == 1): columns >= peer.nloc are not local; entry c is read
  // through peer.halo[c - nloc] (a pointer into a neighbour GPU's vector) after the neighbours' ready flags
  // (peer.flags, local memory) have reached *peer.wait
  PeerGather peer;
  // window SpMM plan (values already refreshed from val), or nullptr
  const WindowPlan* win = nullptr;
  // two-level preconditioner (PTFEM_PRECOND_TWOLEVEL): prepared coarse spaces + what their kernels read
  CoarseSpace* coarse = nullptr;
  int64_t row0 = 0;              // the system is rows [row0, row0+nn) of the arrays (row-range SpMV)
};

// where the fused dot of spmv_launch(cg_dot = true) leaves sum_i y_i x_i of system 0 in PcgWork::scal
constexpr int kScalPqOffset = 4 * 16;

namespace ptfem {
// y = A x for all S systems; if dot_with != nullptr also leaves sum_i y_i*dot_with_i per system in
// work.scal (slot pq) and alpha = rho/pq (CG use).  variant: PTFEM_SPMV_*.
int spmv_launch(ptfem_ctx* ctx, const LinSys& A, int variant, const double* x, double* y, PcgWork* work, bool cg_dot);
int pcg_solve(ptfem_ctx* ctx, const LinSys& A, PcgWork& w, const ptfem_solve_opts& o, double* x /*[nn][S] device*/,
              ptfem_solve_stats* stats);
int pcg_work_alloc(ptfem_ctx* ctx, PcgWork& w, int64_t nn, int S, int VS);
void pcg_work_drop_graph(PcgWork& w);
int resolve_variant(const LinSys& A, int variant);
// pval = val gathered into processing order (one matrix): pval[prowptr[j]+t] = val[rowptr[rowid[j]]+t]
int permute_values(ptfem_ctx* ctx, int64_t nn, const int32_t* rowptr, const int32_t* rowid, const int32_t* prowptr,
                   const double* val, double* pval);
// qval = val with every row padded to even length (pad value 0)
int pad_values(ptfem_ctx* ctx, int64_t nn, const int32_t* rowptr, const int32_t* qrowptr, const double* val, double* qval);
}  // namespace ptfem
