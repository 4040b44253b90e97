// util.cu — error string, exclusive scan, fills.
#include <cstdarg>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace ptfem {

thread_local std::string g_err;

int set_err(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// ---- caching device allocator ---------------------------------------------------------------------
// One cache per process, keyed by device.  Blocks are only reused after the owning buffer was released
// on the host, and every kernel that touched a buffer was launched on the releasing thread's context stream
// before that release, so a block handed back to the SAME host thread is reused in stream order.  A block
// released by another host thread (sweep pipelines: one thread and one context each) may still be in flight
// on that thread's stream and is never handed out; it returns to the driver when the cache is flushed
// (allocation failure, context destruction), which synchronises the device.
namespace {
struct CacheBlock {
  void* p;
  size_t bytes;
  int device;
  std::thread::id owner;      // host thread that released it last
};
std::mutex g_cache_mu;
std::multimap<size_t, CacheBlock> g_free;                 // by size
std::unordered_map<void*, CacheBlock> g_live;
}  // namespace

int dev_alloc(void** out, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  bytes = (bytes + 511) & ~(size_t)511;
  const std::thread::id me = std::this_thread::get_id();
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto it = g_free.lower_bound(bytes); it != g_free.end() && it->first <= bytes + bytes / 4 + 4096; ++it) {
      if (it->second.device == dev && it->second.owner == me) {
        CacheBlock b = it->second;
        g_free.erase(it);
        g_live[b.p] = b;
        *out = b.p;
        return PTFEM_OK;
      }
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    dev_cache_flush();
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err(PTFEM_ERR_CUDA, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
  }
  std::lock_guard<std::mutex> lk(g_cache_mu);
  g_live[p] = CacheBlock{p, bytes, dev, std::this_thread::get_id()};
  *out = p;
  return PTFEM_OK;
}

void dev_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  auto it = g_live.find(p);
  if (it == g_live.end()) return;
  it->second.owner = std::this_thread::get_id();
  g_free.emplace(it->second.bytes, it->second);
  g_live.erase(it);
}

void dev_cache_flush() {
  std::vector<CacheBlock> blocks;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto& kv : g_free) blocks.push_back(kv.second);
    g_free.clear();
  }
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& b : blocks) {
    cudaSetDevice(b.device);
    cudaFree(b.p);
  }
  cudaSetDevice(cur);
}

// ---- exclusive scan (int32 in, int32 out, int64 carries) ------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* total, long long* smem) {
  // inclusive warp scan
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    long long w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
    long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[32 + lane] = winc - w;  // exclusive offsets of warps
    if (lane == 31) smem[64] = winc;
  }
  __syncthreads();
  long long excl = inc - v + smem[32 + wid];
  *total = smem[64];
  __syncthreads();
  return excl;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int32_t* __restrict__ in, int64_t n,
                                                              long long* __restrict__ sums) {
  __shared__ long long smem[72];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) s += in[i];
  }
  long long tot;
  block_exclusive_scan(s, &tot, smem);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_inplace(long long* sums, int64_t nb, long long* total) {
  __shared__ long long smem[72];
  long long carry = 0;
  for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    long long v = (i < nb) ? sums[i] : 0;
    long long tot;
    long long ex = block_exclusive_scan(v, &tot, smem);
    if (i < nb) sums[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                          int64_t n, const long long* __restrict__ sums) {
  __shared__ long long smem[72];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    v[k] = (i < n) ? in[i] : 0;
    s += v[k];
  }
  long long tot;
  long long ex = block_exclusive_scan(s, &tot, smem) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = (int32_t)ex;
    ex += v[k];
  }
}

// out has n+1 entries: out[n] = total.  in and out may alias.
int exclusive_scan_i32(ptfem_ctx* ctx, const int32_t* in, int32_t* out, int64_t n, int64_t* total) {
  if (n <= 0) {
    if (total) *total = 0;
    int32_t z = 0;
    PT_CK(cudaMemcpyAsync(out, &z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return PTFEM_OK;
  }
  const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  DevBuf<long long> sums;
  PT_TRY(sums.alloc(nb + 1));
  scan_tile_sums<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.p);
  PT_LAUNCH_CHECK(ctx);
  scan_sums_inplace<<<1, SCAN_THREADS, 0, ctx->stream>>>(sums.p, nb, sums.p + nb);
  PT_LAUNCH_CHECK(ctx);
  scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, sums.p);
  PT_LAUNCH_CHECK(ctx);
  long long tot = 0;
  PT_CK(cudaMemcpyAsync(&tot, sums.p + nb, sizeof tot, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  if (tot > 2147483647LL) return set_err(PTFEM_ERR_ARG, "scan total %lld exceeds int32 indexing", tot);
  int32_t t32 = (int32_t)tot;
  PT_CK(cudaMemcpyAsync(out + n, &t32, sizeof t32, cudaMemcpyHostToDevice, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  if (total) *total = tot;
  return PTFEM_OK;
}

__global__ void fill_i32_k(int32_t* p, int32_t v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
__global__ void fill_f64_k(double* p, double v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

int fill_i32(ptfem_ctx* ctx, int32_t* p, int32_t v, int64_t n) {
  if (n <= 0) return PTFEM_OK;
  int grid = (int)((n + 255) / 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  fill_i32_k<<<grid, 256, 0, ctx->stream>>>(p, v, n);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}
int fill_f64(ptfem_ctx* ctx, double* p, double v, int64_t n) {
  if (n <= 0) return PTFEM_OK;
  int grid = (int)((n + 255) / 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  fill_f64_k<<<grid, 256, 0, ctx->stream>>>(p, v, n);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

}  // namespace ptfem
