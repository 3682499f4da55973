// fp32 CUDA-core GEMM engine: C[M,N] = A[M,K] * B[N,K]^T through a fused epilogue.
// This is the parity engine (north_star: fp32 forward/loss/gradients within 1e-5 of the reference at
// 'highest' -- tensor cores round operands to <= 11 mantissa bits), and it also carries the
// latent-classifier layers whose N (2..3 classes) is far below a tensor-core tile.
// Operands are addressed by (row stride, k stride) so K-major and MN-major inputs (dgrad uses W as
// stored, wgrad contracts over the batch) need no transposed copies.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace psvae {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_LD = 68, SG_THREADS = 256;

struct SgemmOperand {
  const float* ptr;
  int64_t s_mn;   // stride between consecutive rows (m or n)
  int64_t s_k;    // stride between consecutive k
};

template <class Epi>
__global__ void __launch_bounds__(SG_THREADS) sgemm_kernel(SgemmOperand A, SgemmOperand B, int M, int N, int64_t K, int64_t k_chunk, int vec_ok, Epi epi) {
  PSVAE_GRID_DEP();
  __shared__ __align__(16) float As[2][SG_BK][SG_LD];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_LD];
  __shared__ float red_scratch[32];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int split = blockIdx.z;
  const int64_t kb = (int64_t)split * k_chunk;
  const int64_t ke = min(K, kb + k_chunk);

  // loader mapping: put the unit-stride direction on consecutive threads
  const bool a_kmajor = (A.s_k == 1), b_kmajor = (B.s_k == 1);
  int a_mn[4], a_k[4], b_mn[4], b_k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (a_kmajor) { a_k[i] = t & 15; a_mn[i] = (t >> 4) + 16 * i; } else { a_mn[i] = t & 63; a_k[i] = (t >> 6) + 4 * i; }
    if (b_kmajor) { b_k[i] = t & 15; b_mn[i] = (t >> 4) + 16 * i; } else { b_mn[i] = t & 63; b_k[i] = (t >> 6) + 4 * i; }
  }
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + a_mn[i], k = k0 + a_k[i];
      ra[i] = (m < M && k < ke) ? __ldg(A.ptr + m * A.s_mn + k * A.s_k) : 0.f;
      const int64_t n = n0 + b_mn[i], kk = k0 + b_k[i];
      rb[i] = (n < N && kk < ke) ? __ldg(B.ptr + n * B.s_mn + kk * B.s_k) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][a_k[i]][a_mn[i]] = ra[i];
      Bs[buf][b_k[i]][b_mn[i]] = rb[i];
    }
  };

  int buf = 0;
  if (kb < ke) {
    gload(kb);
    sstore(0);
  }
  __syncthreads();
  for (int64_t k0 = kb; k0 < ke; k0 += SG_BK) {
    const bool more = (k0 + SG_BK) < ke;
    if (more) gload(k0 + SG_BK);        // global loads in flight while this tile is multiplied
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  float red = 0.f;
  const int col0 = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= M) continue;
    if (vec_ok && col0 + 3 < N) {
      epi.template apply<4>(row, col0, acc[i], red, split);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + j < N) {
          float one[1] = {acc[i][j]};
          epi.template apply<1>(row, col0 + j, one, red, split);
        }
    }
  }
  if constexpr (Epi::kReduce) {
    const float s = block_sum(red, red_scratch);
    if (t == 0) epi.red_out[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
  }
}

// number of reduction slots an Epi::kReduce launch writes
static inline int64_t sgemm_red_slots(int64_t M, int N, int splits = 1) { return ceil_div64(M, SG_BM) * ceil_div64(N, SG_BN) * splits; }

void count_launch();

template <class Epi>
int sgemm_launch(const SgemmOperand& A, const SgemmOperand& B, int64_t M, int N, int64_t K, int splits, bool vec_ok, const Epi& epi, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (splits < 1) splits = 1;
  int64_t k_chunk = align_up64(ceil_div64(K, splits), SG_BK);
  dim3 grid((unsigned)ceil_div64(M, SG_BM), (unsigned)ceil_div64(N, SG_BN), (unsigned)splits);
  launch_dep(sgemm_kernel<Epi>, dim3(grid), dim3(SG_THREADS), 0, st, A, B, (int)M, N, K, k_chunk, vec_ok ? 1 : 0, epi);
  count_launch();
  PSVAE_LAUNCH_CHECK("sgemm_kernel");
  return 0;
}

}  // namespace psvae
