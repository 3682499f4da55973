// Microbenchmark: how fast can the epilogue warps of one CTA read TMEM?  (the K <= 256 layers are bound by the accumulator drain.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_ld_bench tools/tmem_ld_bench.cu && build/tmem_ld_bench
// One CTA per SM allocates all 512 TMEM columns (contents irrelevant) and W warps (W = 4, 8, 16; warp w reads lane quarter w % 4) issue
// tcgen05.ld.32x32b.xN back to back over their column range, DEPTH loads in flight before each tcgen05.wait::ld.  Reports bytes per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int N> struct Regs { uint32_t r[N]; };

template <int N> __device__ __forceinline__ void ld_issue(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void ld_issue<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void ld_issue<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                 "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void ld_issue<64>(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
               "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                 "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
                 "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]),
                 "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]),
                 "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
               : "r"(taddr) : "memory");
}
// 16x256b: 16 lanes x 256 bits per repetition; .x8 = 32 registers per thread (a warp reads 16 lanes x 64 columns)
__device__ __forceinline__ void ld_issue_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                 "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: 32x32b.xN; MODE 1: 16x256b.x8 (two instructions cover the warp's 32 lanes x 64 columns)
template <int N, int DEPTH, int MODE>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int warps, int iters, unsigned long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    const int quarter = warp & 3, grp = warp >> 2, groups = warps >> 2 ? warps >> 2 : 1;
    const int cols_per_grp = 512 / groups;             // this warp's column range
    const uint32_t lane_base = base + ((uint32_t)(quarter * 32) << 16);
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      constexpr int STEP = MODE == 0 ? N : 64;
      for (int c = 0; c < cols_per_grp; c += STEP * DEPTH) {
        uint32_t r[DEPTH][MODE == 0 ? N : 64];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
          const uint32_t col = (uint32_t)(grp * cols_per_grp + (c + d * STEP) % cols_per_grp);
          if constexpr (MODE == 0) {
            ld_issue<N>(lane_base + col, r[d]);
          } else {
            ld_issue_16x256b_x8(lane_base + col, *reinterpret_cast<uint32_t(*)[32]>(&r[d][0]));
            ld_issue_16x256b_x8(lane_base + (16u << 16) + col, *reinterpret_cast<uint32_t(*)[32]>(&r[d][32]));
          }
        }
        ld_wait();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
          for (int i = 0; i < (MODE == 0 ? N : 64); ++i) acc ^= r[d][i];
      }
    }
    t1 = clock64();
  }
  if (acc == 0x12345u) sink[0] = acc;
  if (lane == 0 && warp < warps) atomicMax(out_cycles + blockIdx.x, (unsigned long long)(t1 - t0));
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
  }
}

template <int N, int DEPTH, int MODE>
int run(const char* name, int warps) {
  unsigned long long* d;
  uint32_t* sink;
  CK(cudaMalloc(&d, 148 * 8));
  CK(cudaMalloc(&sink, 4));
  const int iters = 200;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemset(d, 0, 148 * 8));
    tmem_read_kernel<N, DEPTH, MODE><<<148, 512>>>(warps, iters, d, sink);
    CK(cudaDeviceSynchronize());
  }
  unsigned long long h[148];
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i];
  mean /= 148;
  // every group of 4 warps reads its 512/groups columns of all 128 lanes per iteration: the CTA reads 128 lanes x 512 columns x 4 B = 256 KB per iteration
  const double bytes = 256.0 * 1024 * iters;
  printf("%-28s warps %2d depth %d: %8.0f cycles / 256 KB  = %6.1f B/clk/SM\n", name, warps, DEPTH, mean / iters, bytes / mean);
  cudaFree(d);
  cudaFree(sink);
  return 0;
}

int main() {
  for (int w : {4, 8, 16}) {
    run<16, 1, 0>("32x32b.x16", w);
    run<32, 1, 0>("32x32b.x32", w);
    run<32, 2, 0>("32x32b.x32", w);
    run<64, 1, 0>("32x32b.x64", w);
    run<32, 1, 1>("16x256b.x8 (x2 per 64 cols)", w);
  }
  return 0;
}
