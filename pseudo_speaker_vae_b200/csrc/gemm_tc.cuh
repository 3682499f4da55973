// tcgen05 GEMM engine for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  through a fused epilogue.
//
//   * persistent, one CTA per SM, static round-robin tile schedule;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one thread issues
//     tcgen05.mma), warps 2..9 = epilogue (TMEM -> registers -> fused epilogue -> global);
//   * operands bf16, staged by TMA (cp.async.bulk.tensor, 128B swizzle) through a STAGES-deep mbarrier
//     ring; fp32 accumulators live in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1;
//   * both operands may be K-major (row-major [rows][K]) or MN-major (stored [K][rows]): dgrad consumes W
//     as stored and wgrad contracts over the batch, so no transposed copies exist anywhere;
//   * optional split-K (wgrad: K = batch) -- partials go through EpiStore's split slot.
//
// Tile: 128 x BN x 64 (UMMA 128 x BN x 16, cta_group::1).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace psvae {

constexpr int TC_BM = 128, TC_BK = 64, TC_UMMA_K = 16;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);

template <int BN> struct TcCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = TC_BM * TC_BK * 2;
  static constexpr int kBBytes = BN * TC_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;     // power of two >= 32 for BN in {64,128,256}
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Column sums of a 32(lanes = rows) x 32(registers = columns) block: afterwards lane l holds sum over the 32 lanes of v[l].
// Butterfly that halves the number of live values per lane at every step: 16+8+4+2+1 = 31 shuffles.
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <class Epi, class = void> struct epi_colsum : std::false_type {};
template <class Epi> struct epi_colsum<Epi, std::enable_if_t<Epi::kColSum>> : std::true_type {};

struct TcShape {
  int64_t M;        // rows of C
  int32_t N;        // cols of C
  int64_t K;
  int32_t splits;   // split-K factor (>= 1)
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // UMMA smem-descriptor strides (bytes)
};

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, TcShape s, Epi epi) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* red_smem = reinterpret_cast<float*>(tmem_slot + 2);
  // bias-gradient column sums (Epi::kColSum): one private slice per epilogue warp, [TC_EPI_WARPS][n_tiles * BN/2] floats,
  // accumulated in program order over the CTA's static tile sequence -> deterministic
  constexpr bool kCS = epi_colsum<Epi>::value;
  float* cs_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], TC_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t m_tiles = (s.M + TC_BM - 1) / TC_BM;
  const int64_t n_tiles = (s.N + BN - 1) / BN;
  const int64_t kb_total = (s.K + TC_BK - 1) / TC_BK;
  const int64_t kb_per_split = (kb_total + s.splits - 1) / s.splits;
  const int64_t num_tiles = m_tiles * n_tiles * s.splits;
  const int cs_slice = (int)n_tiles * (BN / 2);
  if constexpr (kCS) {
    for (int i = threadIdx.x; i < TC_EPI_WARPS * cs_slice; i += TC_THREADS) cs_smem[i] = 0.f;
    __syncthreads();
  }

  float red = 0.f;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, sp = tile / (n_tiles * m_tiles);
        const int64_t kb0 = sp * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        for (int64_t kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int32_t k_el = (int32_t)(kb * TC_BK);
          if constexpr (!A_MN) {
            ptx::tma_load_2d(sa, &tma_a, &full_bar[stage], k_el, (int32_t)(m_t * TC_BM));
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / 64; ++j)
              ptx::tma_load_2d(sa + j * (64 * TC_BK * 2), &tma_a, &full_bar[stage], (int32_t)(m_t * TC_BM + j * 64), k_el);
          }
          if constexpr (!B_MN) {
            ptx::tma_load_2d(sb, &tma_b, &full_bar[stage], k_el, (int32_t)(n_t * BN));
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sb + j * (64 * TC_BK * 2), &tma_b, &full_bar[stage], (int32_t)(n_t * BN + j * 64), k_el);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t a_kstep = A_MN ? TC_UMMA_K * 128 : TC_UMMA_K * 2;   // bytes to the next K=16 slice
      constexpr uint32_t b_kstep = B_MN ? TC_UMMA_K * 128 : TC_UMMA_K * 2;
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int64_t sp = tile / (n_tiles * m_tiles);
        const int64_t kb0 = sp * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        if (kb0 >= kb1) {
          // empty K range (more splits than K blocks): the epilogue must still see zeros -- issue nothing,
          // signal immediately; the epilogue treats `kb0 >= kb1` as an all-zero accumulator.
          ptx::umma_commit(&tfull_bar[acc]);
          continue;
        }
        for (int64_t kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase, 3);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int kk = 0; kk < TC_BK / TC_UMMA_K; ++kk) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + kk * a_kstep, s.a_lbo, s.a_sbo);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + kk * b_kstep, s.b_lbo, s.b_sbo);
            ptx::umma_f16(tmem_d, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);      // smem slot free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);          // accumulator complete
      }
    }
  } else {
    // ============================ epilogue ================================
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch (warp id % 4)
    const int half = ew >> 2;                // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int64_t n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, sp = tile / (n_tiles * m_tiles);
      const int64_t kb0 = sp * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase, 4);
      ptx::tc_fence_after();
      const int64_t row = m_t * TC_BM + quarter * 32 + lane;
      const bool zero_acc = kb0 >= kb1;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        const int col_in_tile = half * COLS_PER_WARP + c;
        const int col = (int)(n_t * BN) + col_in_tile;
        float v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + col_in_tile);
        ptx::tmem_ld_32x32(taddr, v);         // warp-collective: executed by all lanes regardless of bounds
        if (zero_acc) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (row < s.M) {
          if (col + 32 <= s.N) {
            epi.template apply<32>(row, col, v, red, (int)sp);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col + j + 8 <= s.N) {
                float w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = v[j + i];
                epi.template apply<8>(row, col + j, w, red, (int)sp);
              }
            }
          }
        }
        if constexpr (kCS) {
          // v now holds what apply<32> stored (the host enables kColSum only when N % 32 == 0); rows past M contribute zeros.
          // Whole-warp shuffle: outside any lane-divergent branch.
          if (row >= s.M) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          const float cs = warp_transpose_reduce32(v, lane);
          cs_smem[ew * cs_slice + (int)n_t * (BN / 2) + c + lane] += cs;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
    }
  }

  // ============================ teardown ==================================
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCS) {
    // fixed-order sum of the four row-quarter slices that saw each column -> one partial row per CTA
    for (int cidx = threadIdx.x; cidx < s.N; cidx += TC_THREADS) {
      const int nt = cidx / BN, within = cidx % BN, hf = within / (BN / 2), idx = nt * (BN / 2) + within % (BN / 2);
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) t += cs_smem[(hf * 4 + q) * cs_slice + idx];
      epi.colsum[(int64_t)blockIdx.x * s.N + cidx] = t;
    }
  }
  if constexpr (Epi::kReduce) {
    const float ws = warp_sum(red);
    if (lane == 0) red_smem[warp] = ws;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < TC_THREADS / 32; ++w) t += red_smem[w];
      epi.red_out[blockIdx.x] = t;
    }
  }
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcOperand {
  const void* ptr;   // bf16
  int64_t rows;      // M (A) or N (B)
  int64_t ld;        // K-major: elements between rows;  MN-major: elements between consecutive k
  bool mn_major;
};

// Encodes (and caches by value) the 2D tensor map of one operand.  Returns 0, or < 0 with the error text set.
int tc_tensor_map(const TcOperand& op, int64_t K, int box_rows, CUtensorMap* out);
int tc_grid_size();       // number of SMs of the current device (persistent grid)
int tc_device_check();    // 0 when the current device is sm_100
void count_launch();
// UMMA smem-descriptor strides of one staged operand tile (see the canonical layouts in ptx.cuh)
static inline void tc_desc_strides(bool mn_major, uint32_t* lbo, uint32_t* sbo) {
  if (!mn_major) { *lbo = 16; *sbo = 8 * 128; }                 // K-major SW128: 8-row groups 1024 B apart; LBO unused
  else { *lbo = TC_BK * 128; *sbo = 8 * 128; }                  // MN-major SW128: next 64-wide MN block / next 8 k
}

static inline int tc_pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  return 64;
}

template <int BN, bool A_MN, bool B_MN, class Epi>
int gemm_tc_launch_bn(const TcOperand& A, const TcOperand& B, int64_t M, int N, int64_t K, int splits, const Epi& epi, int grid_limit, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  CUtensorMap ta, tb;
  PSVAE_TRY(tc_tensor_map(A, K, A_MN ? TC_BK : TC_BM, &ta));
  PSVAE_TRY(tc_tensor_map(B, K, B_MN ? TC_BK : BN, &tb));
  TcShape s;
  s.M = M; s.N = N; s.K = K; s.splits = splits < 1 ? 1 : splits;
  tc_desc_strides(A_MN, &s.a_lbo, &s.a_sbo);
  tc_desc_strides(B_MN, &s.b_lbo, &s.b_sbo);
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, Epi>;
  static unsigned long long attr_mask = 0;   // per instantiation, one bit per device
  int dev = 0;
  PSVAE_CUDA(cudaGetDevice(&dev));
  if (!(attr_mask >> (dev & 63) & 1ull)) {
    PSVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_mask |= 1ull << (dev & 63);
  }
  const int64_t tiles = ceil_div64(M, TC_BM) * ceil_div64(N, BN) * s.splits;
  int smem_bytes = Cfg::kSmemBytes;
  if constexpr (epi_colsum<Epi>::value) {
    const int64_t cs_bytes = (int64_t)TC_EPI_WARPS * ceil_div64(N, BN) * (BN / 2) * (int64_t)sizeof(float);
    if (N % 32 != 0 || smem_bytes + cs_bytes > 227 * 1024) { set_error("gemm_tc: column sums need N %% 32 == 0 and N <= ~2048 (N=%d)", N); return -2; }
    smem_bytes += (int)cs_bytes;
  }
  int grid = tc_grid_size();
  if (grid_limit > 0 && grid_limit < grid) grid = grid_limit;
  if (tiles < grid) grid = (int)tiles;
  if (grid < 1) return 0;
  kern<<<grid, TC_THREADS, smem_bytes, st>>>(ta, tb, s, epi);
  count_launch();
  PSVAE_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

// can the tcgen05 epilogue produce the column sums of an [M, N] output? (see gemm_tc_launch_bn)
static inline bool tc_colsum_ok(int N, int force_bn = 0) {
  const int bn = force_bn ? force_bn : tc_pick_bn(N);
  const int64_t cs_bytes = (int64_t)TC_EPI_WARPS * ceil_div64(N, bn) * (bn / 2) * (int64_t)sizeof(float);
  const int base = bn == 256 ? TcCfg<256>::kSmemBytes : (bn == 128 ? TcCfg<128>::kSmemBytes : TcCfg<64>::kSmemBytes);
  return N % 32 == 0 && base + cs_bytes <= 227 * 1024;
}

// number of reduction slots a kReduce epilogue needs (one per CTA of the persistent grid)
static inline int tc_red_slots() { return tc_grid_size(); }

template <bool A_MN, bool B_MN, class Epi>
int gemm_tc_launch(const TcOperand& A, const TcOperand& B, int64_t M, int N, int64_t K, int splits, const Epi& epi, cudaStream_t st, int force_bn = 0) {
  if (N % 8 != 0) { set_error("gemm_tc: N=%d must be a multiple of 8", N); return -2; }
  const int bn = force_bn ? force_bn : tc_pick_bn(N);
  switch (bn) {
    case 256: return gemm_tc_launch_bn<256, A_MN, B_MN, Epi>(A, B, M, N, K, splits, epi, 0, st);
    case 128: return gemm_tc_launch_bn<128, A_MN, B_MN, Epi>(A, B, M, N, K, splits, epi, 0, st);
    case 64: return gemm_tc_launch_bn<64, A_MN, B_MN, Epi>(A, B, M, N, K, splits, epi, 0, st);
  }
  set_error("gemm_tc: unsupported BN=%d", bn);
  return -2;
}

}  // namespace psvae
