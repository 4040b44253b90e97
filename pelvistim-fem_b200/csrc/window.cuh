// window.cuh — plan of the "window" multi-RHS SpMM (window.cu builds it, solver.cu runs it).
//
// Rows are processed in tiles of 64 (bricks of the mesh when the numbering is that of a structured grid).  Per tile the plan
// holds ONE blob - values, 16-bit local column indices, row offsets, row ids - and the tile's x window as a few ranges of
// consecutive vector rows.  The kernel brings blob and window into shared memory with cp.async.bulk from a producer warp
// and multiplies out of shared memory: per (row, non-zero) 0.75 LSU wavefronts instead of 1.1 with global gathers, which
// is what bounds the streaming kernel on 8 right-hand sides.
#pragma once
#include <cstdint>

struct WinTile {          // 32 bytes
  int64_t blob_off;       // bytes into WindowPlan::blob (multiple of 16)
  int32_t blob_bytes;     // multiple of 16
  int32_t nrows;          // rows of the tile (<= kWinRows)
  int32_t nnzp;           // non-zeros, padded to a multiple of 8
  int32_t nranges;        // x ranges (at WindowPlan::ranges + tile * kWinMaxRanges)
  int32_t wrows;          // vector rows of the window
  int32_t pad;
};
struct WinRange {         // 16 bytes
  int32_t xstart, nrows, woff, pad;
};

constexpr int kWinRows = 64;        // rows per tile
constexpr int kWinMaxRanges = 32;   // ranges per tile the producer warp issues in one pass

// blob layout of a tile with R rows and nnzp padded non-zeros (every section a multiple of 16 bytes):
//   double   val  [nnzp]
//   int32    rid  [(R + 3) & ~3]          mesh row of tile row rr
//   uint16   roff [(R + 1 + 7) & ~7]      first non-zero of tile row rr within the tile
//   uint16   ldiag[(R + 7) & ~7]          window index of the row's own vector entry (fused p.Ap)
//   uint16   lcol [nnzp]                  window index of each non-zero's column
__host__ __device__ inline int win_blob_bytes(int R, int nnzp) {
  return nnzp * 8 + ((R + 3) & ~3) * 4 + ((R + 1 + 7) & ~7) * 2 + ((R + 7) & ~7) * 2 + nnzp * 2;
}

struct WindowPlan {
  bool valid = false;
  int64_t ntiles = 0;
  int32_t wmax = 0;        // largest window (vector rows)
  int32_t capblob = 0;     // largest blob (bytes, multiple of 128)
  int32_t grid_a = 0, grid_b = 0;   // detected line length / plane size of the numbering (0: none)
  double window_rows_per_row = 0.0;
  WinTile* tiles = nullptr;
  WinRange* ranges = nullptr;
  unsigned char* blob = nullptr;
  int64_t blob_bytes = 0;
};
