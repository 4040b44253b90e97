// pattern.cu — K1: CSR sparsity pattern of the P1 stiffness matrix and the deterministic
// element -> nnz map (plus its inverse, used by the atomics-free gather assembly).
//
// Replaces ElmerSolver's matrix-structure creation for `StatCurrentSolve`
// (step01_box/case.sif:33-45).  Canonical pattern: rows in mesh node order, columns sorted
// ascending, diagonal always present -> bit-exact comparable with oracle/fem_oracle.csr_pattern.
//
// All integer work; atomics are used only for counting / slot claiming and every list is sorted
// afterwards, so the result does not depend on thread scheduling.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

using namespace ptfem;

namespace {

// ---- incidence lists: vertex -> elements (elements have NV vertices) ---------------------------
template <int NV>
__global__ void count_incidence(const int32_t* __restrict__ elems, int64_t ne, int32_t* __restrict__ cnt) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
#pragma unroll
  for (int a = 0; a < NV; ++a) atomicAdd(&cnt[elems[e * NV + a]], 1);
}

template <int NV>
__global__ void fill_incidence(const int32_t* __restrict__ elems, int64_t ne, const int32_t* __restrict__ ptr,
                               int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
#pragma unroll
  for (int a = 0; a < NV; ++a) {
    int32_t v = elems[e * NV + a];
    int32_t pos = atomicAdd(&cursor[v], 1);
    list[ptr[v] + pos] = (int32_t)e;
  }
}

// insertion sort of each (short) list: thread per list.  The lists of a block are one contiguous span of the array: it is
// brought into shared memory with coalesced loads, every thread sorts its own segment there, and the span goes back the
// same way (a span that does not fit is sorted in place in global memory, as before).
constexpr int kSortThreads = 256;
constexpr int kSortSpanMax = 12032;  // entries of shared memory a block may get (47 per list on average)
__global__ void __launch_bounds__(kSortThreads) sort_lists(const int32_t* __restrict__ ptr, int32_t* __restrict__ list, int64_t n, int span) {
  extern __shared__ int32_t s_span[];    // [span]: sized by the caller to about twice the block's average share, so short
                                         // lists leave room for many resident blocks
  __shared__ int s_changed;
  const int64_t i0 = (int64_t)blockIdx.x * kSortThreads;
  const int64_t i1 = min(n, i0 + kSortThreads);
  const int64_t i = i0 + threadIdx.x;
  const int32_t base = ptr[i0], top = ptr[i1];
  const bool fits = top - base <= span;
  int32_t* buf = list;
  int32_t off = 0;
  if (threadIdx.x == 0) s_changed = 0;
  if (fits) {
    for (int32_t k = threadIdx.x; k < top - base; k += kSortThreads) s_span[k] = list[base + k];
    buf = s_span;
    off = base;
  }
  __syncthreads();
  if (i < n) {
    const int32_t b = ptr[i] - off, e = ptr[i + 1] - off;
    bool changed = false;
    for (int32_t k = b + 1; k < e; ++k) {
      const int32_t v = buf[k];
      int32_t j = k - 1;
      while (j >= b && buf[j] > v) {
        buf[j + 1] = buf[j];
        --j;
        changed = true;
      }
      buf[j + 1] = v;
    }
    if (changed) s_changed = 1;
  }
  if (fits) {
    __syncthreads();
    if (s_changed)     // lists that arrived in order (most do) are not written back
      for (int32_t k = threadIdx.x; k < top - base; k += kSortThreads) list[base + k] = s_span[k];
  }
}
// launch with a span of about twice the average share of a block
static int launch_sort_lists(ptfem_ctx* ctx, const int32_t* ptr, int32_t* list, int64_t n, int64_t total) {
  if (n <= 0) return PTFEM_OK;
  int64_t span = 2 * ((total * kSortThreads + n - 1) / n) + 256;
  if (span > kSortSpanMax) span = kSortSpanMax;
  sort_lists<<<ceil_div(n, kSortThreads), kSortThreads, (size_t)span * sizeof(int32_t), ctx->stream>>>(ptr, list, n, (int)span);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

// ---- rows: sorted unique neighbour list of every node ------------------------------------------
// scratch region of node i: [4*n2t_ptr[i] + i, 4*n2t_ptr[i+1] + i + 1)  (room for 4*cnt + 1 entries)
constexpr int kUniqThreads = 64;
constexpr int kUniqPitch = 129;     // odd: the threads' private segments start in different banks; room for 32 tets per node
__global__ void __launch_bounds__(kUniqThreads)
    row_unique(const int32_t* __restrict__ tets, const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ n2t, int64_t nn,
               int32_t* __restrict__ scratch, int32_t* __restrict__ rowlen) {
  __shared__ int32_t s_row[kUniqThreads * kUniqPitch];
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t tb = n2t_ptr[i], te = n2t_ptr[i + 1];
  int32_t* out = scratch + ((int64_t)4 * tb + i);
  // the sorted insertion works in shared memory when the node's 4*cnt + 1 candidates fit there, else in the scratch itself
  const bool in_smem = 4 * (te - tb) + 1 <= kUniqPitch;
  int32_t* row = in_smem ? s_row + threadIdx.x * kUniqPitch : out;
  int32_t u = 1;
  row[0] = (int32_t)i;  // the diagonal is always present
  for (int32_t t = tb; t < te; ++t) {
    const int32_t e = n2t[t];
    const int4 v = *reinterpret_cast<const int4*>(tets + (int64_t)e * 4);
    const int32_t nd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int32_t c = nd[a];
      // lower bound in row[0..u)
      int32_t lo = 0, hi = u;
      while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if (row[mid] < c) lo = mid + 1; else hi = mid;
      }
      if (lo < u && row[lo] == c) continue;
      for (int32_t k = u; k > lo; --k) row[k] = row[k - 1];
      row[lo] = c;
      ++u;
    }
  }
  if (in_smem)
    for (int32_t k = 0; k < u; ++k) out[k] = row[k];
  rowlen[i] = u;
}

__global__ void row_copy(const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ scratch,
                         const int32_t* __restrict__ rowptr, int64_t nn, int32_t* __restrict__ col,
                         int32_t* __restrict__ diag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t* row = scratch + ((int64_t)4 * n2t_ptr[i] + i);
  const int32_t b = rowptr[i], e = rowptr[i + 1];
  for (int32_t k = b; k < e; ++k) {
    int32_t c = row[k - b];
    col[k] = c;
    if (c == (int32_t)i) diag[i] = k;
  }
}

// ---- element -> nnz map ---------------------------------------------------------------------------
__global__ void build_e2nnz(const int32_t* __restrict__ tets, int64_t nt, const int32_t* __restrict__ rowptr,
                            const int32_t* __restrict__ col, int32_t* __restrict__ e2nnz, int32_t* __restrict__ gcount) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (tet, local row)
  if (t >= nt * 4) return;
  const int64_t e = t >> 2;
  const int a = (int)(t & 3);
  const int32_t r = tets[e * 4 + a];
  const int32_t b = rowptr[r], en = rowptr[r + 1];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int32_t target = tets[e * 4 + c];
    int32_t lo = b, hi = en;
    while (lo < hi) {
      int32_t mid = (lo + hi) >> 1;
      if (col[mid] < target) lo = mid + 1; else hi = mid;
    }
    e2nnz[e * 16 + a * 4 + c] = lo;
    atomicAdd(&gcount[lo], 1);      // contributions per non-zero (counting only: the order is fixed by the sort below)
  }
}

__global__ void fill_contrib(const int32_t* __restrict__ e2nnz, int64_t n, const int32_t* __restrict__ gptr,
                             int32_t* __restrict__ cursor, int32_t* __restrict__ gsrc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t k = e2nnz[i];
  const int32_t pos = atomicAdd(&cursor[k], 1);
  gsrc[gptr[k] + pos] = (int32_t)i;  // = tet*16 + local ij
}

// ---- space-filling-curve row order for the streaming SpMV -------------------------------------------
// 30-bit Morton key of a node (10 bits per axis over the bounding box)
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void morton_keys(const double* __restrict__ xyz, int64_t nn, double lx, double ly, double lz, double sx,
                            double sy, double sz, uint32_t* __restrict__ key, int32_t* __restrict__ id) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const uint32_t qx = (uint32_t)fmin(1023.0, fmax(0.0, (xyz[i * 3] - lx) * sx));
  const uint32_t qy = (uint32_t)fmin(1023.0, fmax(0.0, (xyz[i * 3 + 1] - ly) * sy));
  const uint32_t qz = (uint32_t)fmin(1023.0, fmax(0.0, (xyz[i * 3 + 2] - lz) * sz));
  key[i] = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
  id[i] = (int32_t)i;
}
// rows whose column list is the previous row's shifted by one (first and last entries): a numbering in
// which consecutive rows gather consecutive vector entries, i.e. warp-level gathers are coalesced
__global__ void count_coherent_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t nn,
                                    unsigned long long* __restrict__ count) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int ok = 0;
  if (r >= 1 && r < nn) {
    const int32_t b0 = rowptr[r - 1], e0 = rowptr[r], e1 = rowptr[r + 1];
    ok = (e1 - e0 == e0 - b0) && col[e0] == col[b0] + 1 && col[e1 - 1] == col[e0 - 1] + 1;
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}
__global__ void perm_row_len(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rowid, int64_t nn,
                             int32_t* __restrict__ len) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nn) len[j] = rowptr[rowid[j] + 1] - rowptr[rowid[j]];
}
// even-padded copy: row lengths rounded up to even; the pad entry points at the row's own column
__global__ void pad_row_len(const int32_t* __restrict__ rowptr, int64_t nn, int32_t* __restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nn) len[i] = (rowptr[i + 1] - rowptr[i] + 1) & ~1;
}
__global__ void pad_copy_col(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ qrowptr,
                             int64_t nn, int32_t* __restrict__ qcol) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t src = rowptr[i], n = rowptr[i + 1] - src, dst = qrowptr[i], nq = qrowptr[i + 1] - dst;
  for (int32_t t = 0; t < n; ++t) qcol[dst + t] = col[src + t];
  for (int32_t t = n; t < nq; ++t) qcol[dst + t] = (int32_t)i;
}
// pcol[prowptr[j] + t] = col[rowptr[rowid[j]] + t]
__global__ void perm_copy_col(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const int32_t* __restrict__ rowid, const int32_t* __restrict__ prowptr, int64_t nn,
                              int32_t* __restrict__ pcol) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nn) return;
  const int32_t src = rowptr[rowid[j]], dst = prowptr[j], n = prowptr[j + 1] - dst;
  for (int32_t t = 0; t < n; ++t) pcol[dst + t] = col[src + t];
}

// largest number of staged entries over the tiles of R rows (alignment slack included)
__global__ void max_tile_nnz(const int32_t* __restrict__ rowptr, int64_t nn, int32_t R, int32_t* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = t * R;
  if (r0 >= nn) return;
  const int64_t r1 = r0 + R < nn ? r0 + R : nn;
  const int32_t k0 = rowptr[r0] & ~3, k1 = (rowptr[r1] + 3) & ~3;
  atomicMax(out, k1 - k0);
}

__global__ void max_row_len(const int32_t* __restrict__ rowptr, int64_t nn, int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t v = 0;
  if (i < nn) v = rowptr[i + 1] - rowptr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

template <int NV>
int build_incidence(ptfem_ctx* ctx, const int32_t* elems, int64_t ne, int64_t nn, DevBuf<int32_t>& ptr,
                    DevBuf<int32_t>& list) {
  PT_TRY(ptr.alloc(nn + 1));
  DevBuf<int32_t> cursor;
  PT_TRY(cursor.alloc(nn + 1));
  PT_TRY(fill_i32(ctx, ptr.p, 0, nn + 1));
  if (ne > 0) {
    count_incidence<NV><<<ceil_div(ne, 256), 256, 0, ctx->stream>>>(elems, ne, ptr.p);
    PT_LAUNCH_CHECK(ctx);
  }
  int64_t total = 0;
  PT_TRY(exclusive_scan_i32(ctx, ptr.p, ptr.p, nn, &total));
  PT_TRY(list.alloc(total > 0 ? total : 1));
  if (ne > 0) {
    PT_TRY(fill_i32(ctx, cursor.p, 0, nn + 1));
    fill_incidence<NV><<<ceil_div(ne, 256), 256, 0, ctx->stream>>>(elems, ne, ptr.p, cursor.p, list.p);
    PT_LAUNCH_CHECK(ctx);
    PT_TRY(launch_sort_lists(ctx, ptr.p, list.p, nn, total));
  }
  PT_CK(cudaStreamSynchronize(ctx->stream));  // cursor goes out of scope
  return PTFEM_OK;
}

}  // namespace

// geometry of the streaming SpMV (solver.cu)
int ptfem_stream_cap_max();
int ptfem_stream_rows_default();
int ptfem_stream_threads();

int ptfem_build_pattern(ptfem_mesh* m) {
  ptfem_ctx* ctx = m->ctx;
  const int64_t nn = m->nn, nt = m->nt, nb = m->nb;
  PT_TRY(build_incidence<4>(ctx, m->tets.p, nt, nn, m->n2t_ptr, m->n2t));
  PT_TRY(build_incidence<3>(ctx, m->tris.p, nb, nn, m->n2b_ptr, m->n2b));

  // sorted unique neighbour lists into scratch, then compact into CSR
  DevBuf<int32_t> scratch, rowlen;
  PT_TRY(scratch.alloc((size_t)16 * nt + nn + 1));
  PT_TRY(rowlen.alloc(nn + 1));
  PT_TRY(m->rowptr.alloc(nn + 1));
  row_unique<<<ceil_div(nn, kUniqThreads), kUniqThreads, 0, ctx->stream>>>(m->tets.p, m->n2t_ptr.p, m->n2t.p, nn, scratch.p, rowlen.p);
  PT_LAUNCH_CHECK(ctx);
  int64_t nnz = 0;
  PT_TRY(exclusive_scan_i32(ctx, rowlen.p, m->rowptr.p, nn, &nnz));
  m->nnz = nnz;
  PT_TRY(m->col.alloc(nnz + 8));   // +8: streaming SpMV over-reads up to one 16-byte granule
  PT_TRY(m->diag.alloc(nn));
  PT_CK(cudaMemsetAsync(m->col.p, 0, (nnz + 8) * sizeof(int32_t), ctx->stream));
  row_copy<<<ceil_div(nn, 128), 128, 0, ctx->stream>>>(m->n2t_ptr.p, scratch.p, m->rowptr.p, nn, m->col.p, m->diag.p);
  PT_LAUNCH_CHECK(ctx);

  // element -> nnz and its inverse (nnz -> sorted list of tet*16+ij)
  PT_TRY(m->e2nnz.alloc((size_t)16 * nt));
  PT_TRY(m->gptr.alloc(nnz + 1));
  PT_TRY(fill_i32(ctx, m->gptr.p, 0, nnz + 1));
  if (nt > 0) {
    build_e2nnz<<<ceil_div(nt * 4, 256), 256, 0, ctx->stream>>>(m->tets.p, nt, m->rowptr.p, m->col.p, m->e2nnz.p, m->gptr.p);
    PT_LAUNCH_CHECK(ctx);
  }
  int64_t total = 0;
  PT_TRY(exclusive_scan_i32(ctx, m->gptr.p, m->gptr.p, nnz, &total));
  if (total != 16 * nt) return set_err(PTFEM_ERR_STATE, "pattern: contribution count %lld != 16*nt", (long long)total);
  PT_TRY(m->gsrc.alloc((size_t)16 * nt));
  if (nt > 0) {
    // reuse rowlen-sized cursor over nnz
    DevBuf<int32_t> cursor;
    PT_TRY(cursor.alloc(nnz + 1));
    PT_TRY(fill_i32(ctx, cursor.p, 0, nnz + 1));
    fill_contrib<<<ceil_div(nt * 16, 256), 256, 0, ctx->stream>>>(m->e2nnz.p, nt * 16, m->gptr.p, cursor.p, m->gsrc.p);
    PT_LAUNCH_CHECK(ctx);
    PT_TRY(launch_sort_lists(ctx, m->gptr.p, m->gsrc.p, nnz, total));
    PT_CK(cudaStreamSynchronize(ctx->stream));
  }

  // processing order of the streaming SpMV: rows sorted along a Morton curve through the node coordinates,
  // so the 64 rows of a tile are a compact 3-D block sharing most of their columns (x gathers then hit in
  // L1 instead of L2).  Only the kernel's private copy of the matrix is permuted; x, y and the API stay
  // in mesh numbering.
  m->has_rowperm = false;
  bool want_perm = ctx->tune_morton > 0;
  if (ctx->tune_morton < 0 && nn >= 4096) {
    // auto: only numberings without row-to-row coherence profit (measured on size L: natural order 0.116 ms
    // without / 0.159 ms with the Morton order, random order 0.360 ms without / 0.249 ms with)
    DevBuf<unsigned long long> cnt;
    PT_TRY(cnt.alloc(1));
    PT_CK(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), ctx->stream));
    count_coherent_rows<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(m->rowptr.p, m->col.p, nn, cnt.p);
    PT_LAUNCH_CHECK(ctx);
    unsigned long long h = 0;
    PT_CK(cudaMemcpyAsync(&h, cnt.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    m->row_coherence = (double)h / (double)nn;
    want_perm = m->row_coherence < 0.25;
  }
  if (want_perm && nn >= 4096) {
    DevBuf<uint32_t> key, key2;
    DevBuf<int32_t> id;
    PT_TRY(key.alloc(nn));
    PT_TRY(key2.alloc(nn));
    PT_TRY(id.alloc(nn));
    PT_TRY(m->rowid.alloc(nn));
    const double ex = m->bb_hi[0] - m->bb_lo[0], ey = m->bb_hi[1] - m->bb_lo[1], ez = m->bb_hi[2] - m->bb_lo[2];
    morton_keys<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(m->xyz.p, nn, m->bb_lo[0], m->bb_lo[1], m->bb_lo[2],
                                                             ex > 0 ? 1024.0 / ex : 0.0, ey > 0 ? 1024.0 / ey : 0.0,
                                                             ez > 0 ? 1024.0 / ez : 0.0, key.p, id.p);
    PT_LAUNCH_CHECK(ctx);
    size_t tmp_bytes = 0;
    PT_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, id.p, m->rowid.p, (int)nn, 0, 30, ctx->stream));
    DevBuf<uint8_t> tmp;
    PT_TRY(tmp.alloc(tmp_bytes));
    PT_CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, id.p, m->rowid.p, (int)nn, 0, 30, ctx->stream));
    ctx->launches += 2;
    DevBuf<int32_t> plen;
    PT_TRY(plen.alloc(nn + 1));
    PT_TRY(m->prowptr.alloc(nn + 1));
    perm_row_len<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(m->rowptr.p, m->rowid.p, nn, plen.p);
    PT_LAUNCH_CHECK(ctx);
    int64_t tot = 0;
    PT_TRY(exclusive_scan_i32(ctx, plen.p, m->prowptr.p, nn, &tot));
    if (tot != nnz) return set_err(PTFEM_ERR_STATE, "row permutation lost entries");
    PT_TRY(m->pcol.alloc(nnz + 8));
    PT_CK(cudaMemsetAsync(m->pcol.p + nnz, 0, 8 * sizeof(int32_t), ctx->stream));
    perm_copy_col<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(m->rowptr.p, m->col.p, m->rowid.p, m->prowptr.p, nn, m->pcol.p);
    PT_LAUNCH_CHECK(ctx);
    PT_CK(cudaStreamSynchronize(ctx->stream));
    m->has_rowperm = true;
  }
  const int32_t* tile_rowptr = m->has_rowperm ? m->prowptr.p : m->rowptr.p;
  // streaming SpMV tiles: R rows (64 by default, halved while a tile would not fit the largest stage);
  // the stage capacity is the largest tile of this pattern rounded up, so uniform meshes get small stages
  // and therefore more resident CTAs per SM (measured: occupancy, not stage depth, is what pays).
  {
    DevBuf<int32_t> mt;
    PT_TRY(mt.alloc(1));
    m->stream_rows = 0;
    m->stream_cap = 0;
    const int cap_max = ctx->tune_stream_cap > 0 ? (ctx->tune_stream_cap + 31) & ~31 : ptfem_stream_cap_max();
    int rmax = ctx->tune_stream_rows > 0 ? ctx->tune_stream_rows : ptfem_stream_rows_default();
    if (rmax > ptfem_stream_threads()) rmax = ptfem_stream_threads();
    for (int R = rmax & ~31; R >= 32 && m->stream_rows == 0; R -= 32) {
      PT_TRY(fill_i32(ctx, mt.p, 0, 1));
      const int64_t ntile = (nn + R - 1) / R;
      max_tile_nnz<<<ceil_div(ntile, 256), 256, 0, ctx->stream>>>(tile_rowptr, nn, R, mt.p);
      PT_LAUNCH_CHECK(ctx);
      int32_t mx = 0;
      PT_CK(cudaMemcpyAsync(&mx, mt.p, sizeof mx, cudaMemcpyDeviceToHost, ctx->stream));
      PT_CK(cudaStreamSynchronize(ctx->stream));
      if (mx <= cap_max) {
        m->stream_rows = R;
        m->stream_cap = ctx->tune_stream_cap > 0 ? cap_max : ((mx + 31) & ~31);
        if (m->stream_cap < 256) m->stream_cap = 256;
      }
    }
  }
  // even-padded copy for the multi-RHS streaming SpMM (natural row order only; the Morton copy keeps its own path)
  m->has_qcopy = false;
  if (!m->has_rowperm && m->stream_rows > 0 && ctx->tune_pair && nn >= 4096) {
    DevBuf<int32_t> qlen;
    PT_TRY(qlen.alloc(nn + 1));
    PT_TRY(m->qrowptr.alloc(nn + 1));
    pad_row_len<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(m->rowptr.p, nn, qlen.p);
    PT_LAUNCH_CHECK(ctx);
    int64_t tot = 0;
    PT_TRY(exclusive_scan_i32(ctx, qlen.p, m->qrowptr.p, nn, &tot));
    m->qnnz = tot;
    PT_TRY(m->qcol.alloc(tot + 8));
    PT_CK(cudaMemsetAsync(m->qcol.p + tot, 0, 8 * sizeof(int32_t), ctx->stream));
    pad_copy_col<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(m->rowptr.p, m->col.p, m->qrowptr.p, nn, m->qcol.p);
    PT_LAUNCH_CHECK(ctx);
    // tile geometry of the padded copy (same rows per tile, a slightly larger stage)
    DevBuf<int32_t> mt;
    PT_TRY(mt.alloc(1));
    PT_TRY(fill_i32(ctx, mt.p, 0, 1));
    const int R = m->stream_rows;
    max_tile_nnz<<<ceil_div((nn + R - 1) / R, 256), 256, 0, ctx->stream>>>(m->qrowptr.p, nn, R, mt.p);
    PT_LAUNCH_CHECK(ctx);
    int32_t mx = 0;
    PT_CK(cudaMemcpyAsync(&mx, mt.p, sizeof mx, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    if (mx <= ptfem_stream_cap_max()) {
      m->q_rows = R;
      m->q_cap = std::max(256, (mx + 31) & ~31);
      m->has_qcopy = true;
    }
  }
  DevBuf<int32_t> mrl;
  PT_TRY(mrl.alloc(1));
  PT_TRY(fill_i32(ctx, mrl.p, 0, 1));
  max_row_len<<<ceil_div(nn, 256), 256, 0, ctx->stream>>>(m->rowptr.p, nn, mrl.p);
  PT_LAUNCH_CHECK(ctx);
  PT_CK(cudaMemcpyAsync(&m->max_row, mrl.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  PT_TRY(window_plan_build(m));    // multi-RHS SpMM out of shared-memory x windows, where the numbering allows it
  m->has_pattern = true;
  return PTFEM_OK;
}
