// Prototype 2: "window" SpMM (S = 8 right-hand sides, [nn][8] vectors) with a producer warp and bulk copies.
// Rows are processed in bricks.  Per tile ONE blob (values, 16-bit local column indices, row offsets, row ids) and the
// tile's x window - a handful of contiguous row ranges of x - are brought into shared memory by cp.async.bulk, issued by a
// producer warp STAGES-1 tiles ahead and completing on an mbarrier; the 256 consumer threads read shared memory only.
// LSU wavefronts per (row, non-zero): 0.5 (x from shared memory, conflict-free for neighbouring columns) + 0.25 (value and
// index) against 0.82 + 0.27 of the streaming kernel with global gathers.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o spmm_window2 scripts/proto/spmm_window2.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int S = 8;
constexpr int kConsumers = 256;

struct TileInfo {
  long long blob_off;   // bytes into the blob array
  int blob_bytes;       // multiple of 16
  int nrows;            // rows of the tile (<= 128)
  int nnzp;             // padded non-zeros (multiple of 8)
  int rbeg, nranges;    // x ranges
  int tx_bytes;         // blob + all ranges
  int pad;
};
struct Range { int xstart, nrows, woff, pad; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

template <int STAGES, int UNROLL>
__global__ void __launch_bounds__(kConsumers + 32) spmm_window2_kernel(int ntiles, const TileInfo* __restrict__ info, const unsigned char* __restrict__ blob,
                                                                       const Range* __restrict__ ranges, const double* __restrict__ x,
                                                                       double* __restrict__ y, int wmax, int capblob) {
  extern __shared__ __align__(128) unsigned char smem[];
  const size_t stage_bytes = (size_t)capblob + (size_t)wmax * S * 8;
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumers / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nloc = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid >= kConsumers) {
    // ---- producer warp
    const int lane = tid - kConsumers;
    for (int j = 0; j < nloc; ++j) {
      const int st = j % STAGES;
      if (j >= STAGES) mbar_wait(&empty[st], ((j / STAGES) - 1) & 1);
      const TileInfo ti = info[blockIdx.x + (size_t)j * gridDim.x];
      unsigned char* sb = smem + (size_t)st * stage_bytes;
      double* xs = reinterpret_cast<double*>(sb + capblob);
      if (lane == 0) {
        mbar_arrive_tx(&full[st], (unsigned)ti.tx_bytes);
        bulk_g2s(sb, blob + ti.blob_off, (unsigned)ti.blob_bytes, &full[st]);
      }
      __syncwarp();
      for (int i = lane; i < ti.nranges; i += 32) {
        const Range r = ranges[ti.rbeg + i];
        bulk_g2s(xs + (size_t)r.woff * S, x + (size_t)r.xstart * S, (unsigned)r.nrows * S * 8u, &full[st]);
      }
    }
  } else {
    // ---- consumers: 4 lanes per row, 64 rows in flight
    const int lane = tid & 3, rloc = tid >> 2;
    for (int j = 0; j < nloc; ++j) {
      const int st = j % STAGES;
      const TileInfo ti = info[blockIdx.x + (size_t)j * gridDim.x];
      const unsigned char* sb = smem + (size_t)st * stage_bytes;
      const double* vs = reinterpret_cast<const double*>(sb);
      const int rp = (ti.nrows + 3) & ~3;
      const int* rid = reinterpret_cast<const int*>(sb + (size_t)ti.nnzp * 8);
      const uint16_t* roff = reinterpret_cast<const uint16_t*>(sb + (size_t)ti.nnzp * 8 + (size_t)rp * 4);
      const uint16_t* ls = roff + ((ti.nrows + 1 + 7) & ~7);
      const double* xs = reinterpret_cast<const double*>(sb + capblob);
      mbar_wait(&full[st], (j / STAGES) & 1);
      for (int rr = rloc; rr < ti.nrows; rr += kConsumers / 4) {
        const int b = roff[rr], e = roff[rr + 1];
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll UNROLL
        for (int k = b; k < e; ++k) {
          const double a = vs[k];
          const double2 xv = *reinterpret_cast<const double2*>(xs + (int)ls[k] * S + 2 * lane);
          acc.x = fma(a, xv.x, acc.x);
          acc.y = fma(a, xv.y, acc.y);
        }
        *reinterpret_cast<double2*>(y + (size_t)rid[rr] * S + 2 * lane) = acc;
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[st]);
    }
  }
}

int main(int argc, char** argv) {
  const int nx1 = 193, ny1 = 145, nz1 = 121;
  int bx = argc > 1 ? atoi(argv[1]) : 16, by = argc > 2 ? atoi(argv[2]) : 4, bz = argc > 3 ? atoi(argv[3]) : 2;
  int ctas_per_sm = argc > 4 ? atoi(argv[4]) : 2;
  int stages = argc > 5 ? atoi(argv[5]) : 2;
  int gapmax = argc > 6 ? atoi(argv[6]) : 1;      // columns further apart than this start a new x range
  const int64_t nn = (int64_t)nx1 * ny1 * nz1;
  const int d[15][3] = {{0,0,0},{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{1,1,0},{-1,-1,0},{0,0,1},{0,0,-1},{1,0,1},{-1,0,-1},{0,1,1},{0,-1,-1},{1,1,1},{-1,-1,-1}};
  auto id = [&](int x, int y, int z) { return x + nx1 * (y + ny1 * z); };
  std::vector<int> rowid; rowid.reserve(nn);
  std::vector<int> tile_row0{0};
  for (int Z = 0; Z < nz1; Z += bz) for (int Y = 0; Y < ny1; Y += by) for (int X = 0; X < nx1; X += bx) {
    for (int z = Z; z < std::min(Z + bz, nz1); ++z) for (int y = Y; y < std::min(Y + by, ny1); ++y) for (int x = X; x < std::min(X + bx, nx1); ++x) rowid.push_back(id(x, y, z));
    tile_row0.push_back((int)rowid.size());
  }
  const int ntiles = (int)tile_row0.size() - 1;
  std::vector<TileInfo> info(ntiles);
  std::vector<Range> ranges;
  std::vector<unsigned char> blob;
  std::vector<int> pcol_all; std::vector<double> pval_all; std::vector<int> prow_ptr(nn + 1);
  int wmax = 0, capblob = 0, rmax = 0;
  size_t wrows = 0, nnz_total = 0;
  std::vector<int> cols, tcol; std::vector<double> tval; std::vector<uint16_t> toff;
  for (int t = 0; t < ntiles; ++t) {
    cols.clear(); tcol.clear(); tval.clear(); toff.clear();
    const int R = tile_row0[t + 1] - tile_row0[t];
    for (int r = tile_row0[t]; r < tile_row0[t + 1]; ++r) {
      const int g = rowid[r], x = g % nx1, y = (g / nx1) % ny1, z = g / (nx1 * ny1);
      toff.push_back((uint16_t)tcol.size());
      prow_ptr[r] = (int)pcol_all.size();
      int cc[15], n = 0;
      for (auto& o : d) { const int X = x + o[0], Y = y + o[1], Z = z + o[2]; if (X < 0 || Y < 0 || Z < 0 || X >= nx1 || Y >= ny1 || Z >= nz1) continue; cc[n++] = id(X, Y, Z); }
      std::sort(cc, cc + n);
      for (int k = 0; k < n; ++k) {
        const double v = cc[k] == g ? 4.0 : -0.25 - 1e-3 * (cc[k] % 7);
        tcol.push_back(cc[k]); tval.push_back(v); cols.push_back(cc[k]); pcol_all.push_back(cc[k]); pval_all.push_back(v);
      }
    }
    toff.push_back((uint16_t)tcol.size());
    nnz_total += tcol.size();
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    // ranges + local index of every window column
    TileInfo& ti = info[t];
    ti.rbeg = (int)ranges.size();
    std::vector<int> local(cols.size());
    int woff = 0;
    for (size_t i = 0; i < cols.size();) {
      size_t j = i;
      while (j + 1 < cols.size() && cols[j + 1] - cols[j] <= gapmax) ++j;
      Range rg{cols[i], cols[j] - cols[i] + 1, woff, 0};
      for (size_t k = i; k <= j; ++k) local[k] = woff + (cols[k] - cols[i]);
      woff += rg.nrows;
      ranges.push_back(rg);
      i = j + 1;
    }
    ti.nranges = (int)ranges.size() - ti.rbeg;
    wmax = std::max(wmax, woff); wrows += woff; rmax = std::max(rmax, ti.nranges);
    const int nnzp = ((int)tcol.size() + 7) & ~7, rp = (R + 3) & ~3, ro = (R + 1 + 7) & ~7;
    const int bytes = nnzp * 8 + rp * 4 + ro * 2 + nnzp * 2;
    while (blob.size() % 16) blob.push_back(0);
    ti.blob_off = (long long)blob.size(); ti.blob_bytes = bytes; ti.nrows = R; ti.nnzp = nnzp;
    ti.tx_bytes = bytes + woff * S * 8; ti.pad = 0;
    blob.resize(blob.size() + bytes, 0);
    unsigned char* p = blob.data() + ti.blob_off;
    double* bv = reinterpret_cast<double*>(p);
    int* bid = reinterpret_cast<int*>(p + (size_t)nnzp * 8);
    uint16_t* bo = reinterpret_cast<uint16_t*>(p + (size_t)nnzp * 8 + (size_t)rp * 4);
    uint16_t* bl = bo + ro;
    for (size_t k = 0; k < tcol.size(); ++k) { bv[k] = tval[k]; bl[k] = (uint16_t)local[std::lower_bound(cols.begin(), cols.end(), tcol[k]) - cols.begin()]; }
    for (int r = 0; r < R; ++r) bid[r] = rowid[tile_row0[t] + r];
    for (int r = 0; r <= R; ++r) bo[r] = toff[r];
    capblob = std::max(capblob, bytes);
  }
  prow_ptr[nn] = (int)pcol_all.size();
  capblob = (capblob + 127) & ~127;
  const size_t stage_bytes = (size_t)capblob + (size_t)wmax * S * 8;
  printf("brick %dx%dx%d gap %d: %d tiles, wmax %d (%.2f window rows per row), max ranges %d (%.1f avg), capblob %d, stage %zu B, blob %.1f MB (%.2f B/nnz)\n", bx, by,
         bz, gapmax, ntiles, wmax, (double)wrows / nn, rmax, (double)ranges.size() / ntiles, capblob, stage_bytes, blob.size() / 1e6, (double)blob.size() / nnz_total);
  std::vector<double> hx((size_t)nn * S);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = 1.0 + 1e-3 * (double)((i * 2654435761u) % 1000);
  TileInfo* d_info; Range* d_ranges; unsigned char* d_blob; double *d_x, *d_y;
  CK(cudaMalloc(&d_info, info.size() * sizeof(TileInfo))); CK(cudaMalloc(&d_ranges, ranges.size() * sizeof(Range)));
  CK(cudaMalloc(&d_blob, blob.size() + 256)); CK(cudaMalloc(&d_x, hx.size() * 8)); CK(cudaMalloc(&d_y, hx.size() * 8));
  CK(cudaMemcpy(d_info, info.data(), info.size() * sizeof(TileInfo), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ranges, ranges.data(), ranges.size() * sizeof(Range), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_x, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice));
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](auto kern, int st, int unroll) {
    const size_t smem = stage_bytes * st;
    if (smem > 227 * 1024 - 1024) { printf("stages %d: %zu B of shared memory do not fit\n", st, smem); return; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kConsumers + 32, smem));
    const int per = std::min(occ, ctas_per_sm);
    const int grid = std::min(ntiles, nsm * per);
    for (int k = 0; k < 3; ++k) kern<<<grid, kConsumers + 32, smem>>>(ntiles, d_info, d_blob, d_ranges, d_x, d_y, wmax, capblob);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int k = 0; k < 20; ++k) kern<<<grid, kConsumers + 32, smem>>>(ntiles, d_info, d_blob, d_ranges, d_x, d_y, wmax, capblob);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double alg = 12.0 * 49826470 + 4.0 * nn + 16.0 * S * nn;
    printf("stages %d unroll %d, %d CTAs/SM (occupancy %d), grid %d: %.4f ms per launch -> %.0f GB/s algorithmic (frac of 6543.7: %.3f)\n", st, unroll, per, occ, grid,
           ms / 20, alg / (ms / 20 * 1e-3) / 1e9, alg / (ms / 20 * 1e-3) / 1e9 / 6543.7);
  };
  if (stages == 2) { run(spmm_window2_kernel<2, 4>, 2, 4); run(spmm_window2_kernel<2, 8>, 2, 8); }
  if (stages == 3) { run(spmm_window2_kernel<3, 4>, 3, 4); run(spmm_window2_kernel<3, 8>, 3, 8); }
  if (stages == 4) { run(spmm_window2_kernel<4, 4>, 4, 4); run(spmm_window2_kernel<4, 8>, 4, 8); }
  std::vector<double> hy(hx.size());
  CK(cudaMemcpy(hy.data(), d_y, hy.size() * 8, cudaMemcpyDeviceToHost));
  double worst = 0.0;
  for (int64_t r = 0; r < nn; r += 997) {
    const int g = rowid[r];
    for (int s = 0; s < S; ++s) {
      double acc = 0.0;
      for (int k = prow_ptr[r]; k < prow_ptr[r + 1]; ++k) acc = fma(pval_all[k], hx[(size_t)pcol_all[k] * S + s], acc);
      worst = std::max(worst, fabs(acc - hy[(size_t)g * S + s]) / fabs(acc));
    }
  }
  printf("max rel. error on sampled rows: %.3e\n", worst);
  return 0;
}
