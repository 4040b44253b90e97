// dist.cu — row-partitioned single solve over several GPUs (placeholder until the NCCL path lands).
#include "solver.cuh"
using namespace ptfem;
void ptfem_dist_ctx_release(ptfem_ctx*) {}
