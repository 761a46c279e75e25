// Microbenchmark: per-SM TMA ingest rate of L2-resident weight slabs, vs ring depth and cluster multicast.
// Every CTA repeatedly streams the same [ROWS x 128] fp32 matrix as [128 x 32] SWIZZLE_128B slabs (16 KB) through a
// ring of NST slots; a consumer thread "uses" a slot for USE cycles (stand-in for 4 MMAs) and releases it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../sepreformer_b200/csrc/kernels_tc.cuh"
using namespace sepref::tc;

template <int CL>
__global__ void __launch_bounds__(64, 1) k_ingest(const __grid_constant__ CUtensorMap map, int nst, int slabs_per_pass, int passes,
                                                  int use_clk, long long* out_clk) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + 12 * 16384);
  uint64_t* empty = full + 16;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
    fence_barrier_init();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  const long long t0 = clock64();
  const int total = slabs_per_pass * passes;
  if (threadIdx.x == 0) {          // producer
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < total; ++i) {
      const int s = i % slabs_per_pass;
      const int row0 = (s / 4) * 128, col0 = (s % 4) * 32;
      mbar_wait(&empty[st], ph ^ 1, 1);
      mbar_arrive_expect_tx(&full[st], 16384);
      if (CL == 1) tma_load_2d(&map, &full[st], sm + st * 16384, col0, row0);
      else tma_load_2d_mc(&map, &full[st], sm + st * 16384 + crank * (128 / CL) * 128, col0, row0 + crank * (128 / CL), (uint16_t)((1u << CL) - 1));
      if (++st == nst) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(&full[st], ph, 2);
      const long long t = clock64();
      while (clock64() - t < use_clk) {}
      if (CL == 1) mbar_arrive(&empty[st]);
      else {
        for (uint32_t r = 0; r < CL; ++r) {     // release the slot in every CTA of the cluster
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[st])), "r"(r));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
      }
      if (++st == nst) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (threadIdx.x == 0) out_clk[blockIdx.x] = clock64() - t0;
}

template <int CL>
double run(const CUtensorMap& map, int nst, int slabs, int passes, int use_clk, int grid, long long* d_clk) {
  const int smem = 1024 + 12 * 16384 + 512;
  cudaFuncSetAttribute(k_ingest<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CL; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_ingest<CL>, map, nst, slabs, passes, use_clk, d_clk);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return -1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d_clk, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  return (double)slabs * passes * 16384.0 / mx;      // bytes per clock per SM
}

int main() {
  init(128);
  const int rows = 1152, cols = 128;      // 590 KB like the GCFN weights
  float* w; cudaMalloc(&w, rows * cols * 4); cudaMemset(w, 0, rows * cols * 4);
  long long* d_clk; cudaMalloc(&d_clk, 1024 * 8);
  CUtensorMap m1, m2, m4;
  make_weight_map(&m1, w, rows, cols, 128); make_weight_map(&m2, w, rows, cols, 64); make_weight_map(&m4, w, rows, cols, 32);
  const int slabs = (rows / 128) * 4;
  printf("use_clk nst  |  B/clk/SM: CL=1 grid148   CL=2 grid148   CL=4 grid132  | CL=1 grid=1\n");
  for (int use : {0, 160}) for (int nst : {2, 4, 6, 8, 12}) {
    double a = run<1>(m1, nst, slabs, 40, use, 148, d_clk);
    double b = run<2>(m2, nst, slabs, 40, use, 148, d_clk);
    double c = run<4>(m4, nst, slabs, 40, use, 132, d_clk);
    double d = run<1>(m1, nst, slabs, 40, use, 1, d_clk);
    printf("%6d %3d  |  %8.1f %14.1f %14.1f  | %8.1f\n", use, nst, a, b, c, d);
  }
  return 0;
}
