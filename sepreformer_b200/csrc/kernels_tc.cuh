// tcgen05 / TMA kernels of the separator path (GEMM_PATH 1).  [stub: filled in by the next milestone]
#pragma once
#include "common.cuh"

namespace sepref {
namespace tc {

struct GcfnPack {
  const float *w1 = nullptr, *b1 = nullptr, *dw = nullptr, *dwb = nullptr, *w2 = nullptr, *b2 = nullptr;
};
inline const char* last_error() { return "tensor-core path not built yet"; }
inline int init(int) { return 0; }
inline int prepare_gcfn(GcfnPack&, int) { return 0; }
inline int launch_gcfn(const GcfnPack&, const float*, float*, int, int, int, int, cudaStream_t) { return -1; }

}  // namespace tc
}  // namespace sepref
