// Chained decoder for sampling: z -> Linear(64,H)+ReLU -> Linear(H,H)+ReLU -> Linear(H,D) (ps_vae/model.py:30-36,65-69; the tail of
// unconditional_synthesis / conditional_synthesis, ps_vae/inference.py:22-25,105) in ONE persistent kernel.  The hidden activations never
// leave the SM: a CTA pair (tcgen05 cta_group::2) owns 256 rows, each CTA keeps its 128 x H bf16 first hidden layer in shared memory as the
// A operand of the second layer, the second hidden layer exists only 128 columns at a time (it is consumed as K-blocks of the last layer,
// whose 128 x D fp32 accumulator stays in TMEM for the whole tile), z comes from the counter-based generator inside the kernel, and the only
// HBM traffic is the fp32 output (1 KB per sample).  Weights stream from L2 by TMA, each CTA of the pair loading half of every block.
//
// Warp roles (640 threads): 0 = weight producer (TMA), 1 = TMEM allocator + MMA issuer (leader CTA only), 2..17 = epilogue (TMEM -> bias / ReLU
// -> bf16 -> swizzled smem operand, or fp32 -> staged block -> TMA store), 18..19 = z (Philox + Box-Muller, or an fp32 z given by the caller).
//
// Per tile the leader issues, in this order (B = one [128 n x 64 k] weight block, half per CTA):
//   L0  chunk c = 0..NC-1      acc[c & 1]  = z (K = 64) x W0[c]                                 1 B each
//   L1  chunk c = 0..NC-1      acc[c & 1]  = hd0 (K = H) x W1[c]                                H/64 B each
//   L2  after L1(c), c >= 1:   out        += hd1 chunk c-1 (2 K-blocks) x W2[:, chunk c-1]      4 B each;   L2(NC-1) after the loop
// so that the epilogue of one chunk always runs under the MMAs of the next.  TMEM: acc[2] = 2 x 128 columns, out = 256 columns.
//
// TRAIN = true is the decoder half of the training forward pass (ps_vae/model.py:58 + the reconstruction term of ps_vae/lightning.py:113) on
// the same schedule: z arrives in bf16 from the fused encoder-head kernel, every hidden K-block is ALSO written to HBM straight out of its
// operand buffer by TMA (the backward pass needs the activations: wgrad operands) together with its 1-bit ReLU mask, and the output phase
// does what EpiMse does -- bias, sum of squared errors, d x_hat = 2 (x_hat - x) / (B D 10) in bf16 by TMA store, its column sums (= the bias
// gradient of the last layer) -- instead of storing x_hat.  The hidden activations are written once and never read back in the forward pass.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace psvae {

constexpr int DC_ROWS = 128;           // rows per CTA (256 per pair)
constexpr int DC_BK = 64;
constexpr int DC_CHUNK = 128;          // columns of a hidden layer per accumulator chunk
constexpr int DC_L = 64;               // latent width the chain is written for
constexpr int DC_STAGES = 6;           // weight ring depth; a stage = this CTA's [64 n x 64 k] bf16 half block
constexpr int DC_STAGE_BYTES = 64 * DC_BK * 2;       // 8 KB
constexpr int DC_KBLOCK_BYTES = DC_ROWS * DC_BK * 2; // 16 KB: one [128 rows x 64 k] A operand block
constexpr int DC_EPI_WARPS = 16;
constexpr int DC_Z_WARPS = 2;          // each thread generates two rows of z (640 threads: 96 registers each, no spills in the epilogue)
constexpr int DC_THREADS = 32 * (2 + DC_EPI_WARPS + DC_Z_WARPS);
constexpr int DC_MAX_H = 512;
constexpr int DC_BAR_BYTES = 1024;
// shared memory: [barriers 1 KB][hd0: H/64 K-blocks][hd1: 2 K-blocks, later the output staging][z: 1 K-block][weight ring]
static inline int dc_smem_bytes(int H) { return DC_BAR_BYTES + (H / DC_BK) * DC_KBLOCK_BYTES + 2 * DC_KBLOCK_BYTES + DC_KBLOCK_BYTES + DC_STAGES * DC_STAGE_BYTES + 1024; }

struct DcArgs {
  int64_t rows;
  int32_t H, D;
  const float* b0;
  const float* b1;
  const float* b2;
  const float* z_in;      // optional fp32 z [rows][64] (conditional sampling: the Langevin result); nullptr: z ~ N(0, I) from Philox
  float* z_out;           // optional: the z used, fp32 [rows][64]
  uint64_t seed, offset;
  int64_t first_row;      // Philox element index = (first_row + r) * 64 + c: independent of how the rows are sharded
  unsigned long long* trace;   // profiling (option "tc_trace_ptr"): per CTA 16 cycle counters, see tools/chain_trace.py; nullptr = off
  // ---- TRAIN
  const bf16* z16;        // z [rows][64] bf16 (written by the fused encoder head)
  uint32_t* mask0;        // ReLU bit masks of the two hidden layers, [H/32][mask_ld]
  uint32_t* mask1;
  int64_t mask_ld;
  const void* x;          // reconstruction target [rows][D], fp32 or (x_bf16) bf16
  int32_t x_bf16;
  float scale;            // 2 / (B * D * 10)
  float* sse_part;        // [CTAs]: sum of squared errors
  float* bias_grad;       // [D]: += column sums of d x_hat (atomics into the zeroed gradient)
};

__device__ __forceinline__ uint32_t dc_swz128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t dc_swz64(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

template <bool TRAIN, bool TRACE>
__global__ void __launch_bounds__(DC_THREADS, 1)
decoder_chain_kernel(const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                     const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_hd0, const __grid_constant__ CUtensorMap tm_hd1,
                     DcArgs a) {
  // tm_out: sampling: x_hat fp32 [rows][D], box 16 x 32;  TRAIN: d x_hat bf16 [rows][D], box 32 x 32.  tm_hd0 / tm_hd1 (TRAIN): [rows][H] bf16, box 64 x 32
  extern __shared__ uint8_t dc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dc_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int KB = a.H / DC_BK;            // K-blocks of a hidden layer
  const int NC = a.H / DC_CHUNK;         // accumulator chunks of a hidden layer
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* b_full = bars;                       // [STAGES] weights landed (leader)
  uint64_t* b_empty = b_full + DC_STAGES;        // [STAGES] stage consumed (both CTAs)
  uint64_t* acc_full = b_empty + DC_STAGES;      // [2] chunk accumulator complete (both CTAs)
  uint64_t* acc_empty = acc_full + 2;            // [2] chunk accumulator drained (leader; 2 x 16 warps)
  uint64_t* z_full = acc_empty + 2;              // z block written (leader; 2 x 4 warps)
  uint64_t* z_empty = z_full + 1;                // L0 MMAs done with z (both CTAs)
  uint64_t* hd0_ready = z_empty + 1;             // [8] K-block of hd0 written (leader; 2 x 8 warps each)
  uint64_t* hd0_free = hd0_ready + 8;            // L1 MMAs done with hd0 (both CTAs)
  uint64_t* hd1_ready = hd0_free + 1;            // [2] K-block of the hd1 chunk written (leader; 2 x 8 warps each)
  uint64_t* hd1_empty = hd1_ready + 2;           // L2 MMAs done with the hd1 chunk (both CTAs)
  uint64_t* out_full = hd1_empty + 1;            // output accumulator complete (both CTAs)
  uint64_t* out_empty = out_full + 1;            // output accumulator drained (leader; 2 x 16 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_empty + 1);
  float* red_smem = reinterpret_cast<float*>(smem + 512);      // [DC_EPI_WARPS]: per-warp sums of squared errors (TRAIN)
  uint8_t* const hd0 = smem + DC_BAR_BYTES;
  uint8_t* const hd1 = hd0 + KB * DC_KBLOCK_BYTES;
  uint8_t* const zbuf = hd1 + 2 * DC_KBLOCK_BYTES;
  uint8_t* const ring = zbuf + DC_KBLOCK_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int64_t pair_id = blockIdx.x >> 1, pairs = gridDim.x >> 1;
  const int64_t tiles = (a.rows + 2 * DC_ROWS - 1) / (2 * DC_ROWS);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_w0);
    ptx::prefetch_tensormap(&tm_w1);
    ptx::prefetch_tensormap(&tm_w2);
    ptx::prefetch_tensormap(&tm_out);
    for (int i = 0; i < DC_STAGES; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 2 * DC_EPI_WARPS); }
    ptx::mbar_init(z_full, 2 * DC_Z_WARPS);
    ptx::mbar_init(z_empty, 1);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&hd0_ready[i], 2 * 8);
    ptx::mbar_init(hd0_free, 1);
    for (int i = 0; i < 2; ++i) ptx::mbar_init(&hd1_ready[i], 2 * 8);
    ptx::mbar_init(hd1_empty, 1);
    ptx::mbar_init(out_full, 1);
    ptx::mbar_init(out_empty, 2 * DC_EPI_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_cg2(tmem_slot, 512);
    ptx::tmem_relinquish_cg2();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // shared::cluster address of a barrier in the leader CTA (remote arrive)
  auto lead = [&](uint64_t* bar) { return ptx::mapa_u32(ptx::smem_u32(bar), 0u); };
  // profiling: cycles a role spent inside each class of barrier wait (slot = class), written by lane 0 of the MMA warp / epilogue warp 2 / z warp
  constexpr bool tracing = TRACE;          // a template parameter: the counters cost 16 registers the product instantiation does not have
  long long tw[TRACE ? 8 : 1] = {};
  const long long t_begin = tracing ? clock64() : 0;
  auto wait = [&](uint64_t* bar, uint32_t parity, int tag, int slot) {
    if constexpr (TRACE) {
      const long long t0 = clock64();
      ptx::mbar_wait(bar, parity, tag);
      tw[slot] += clock64() - t0;
    } else {
      ptx::mbar_wait(bar, parity, tag);
    }
  };

  if (warp == 0) {
    // ============================ weight producer ============================
    int stage = 0;
    uint32_t phase = 0;
    auto item = [&](const CUtensorMap* map, int32_t k0, int32_t n0) {
      wait(&b_empty[stage], phase ^ 1, 1, 0);
      if (ptx::elect_one()) {
        if (leader) ptx::mbar_arrive_expect_tx(&b_full[stage], 2 * DC_STAGE_BYTES);
        ptx::tma_load_2d_cg2(ring + stage * DC_STAGE_BYTES, map, lead(&b_full[stage]), k0, n0 + (int32_t)cta_rank * 64);
      }
      __syncwarp();
      if (++stage == DC_STAGES) { stage = 0; phase ^= 1; }
    };
    auto l2_items = [&](int cc) {
      for (int j = 0; j < 2; ++j)
        for (int h = 0; h < 2; ++h) item(&tm_w2, (cc * 2 + j) * DC_BK, h * DC_CHUNK);
    };
    for (int64_t tile = pair_id; tile < tiles; tile += pairs) {
      for (int c = 0; c < NC; ++c) item(&tm_w0, 0, c * DC_CHUNK);
      for (int c = 0; c < NC; ++c) {
        for (int kb = 0; kb < KB; ++kb) item(&tm_w1, kb * DC_BK, c * DC_CHUNK);
        if (c >= 1) l2_items(c - 1);
      }
      l2_items(NC - 1);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA) ============================
    if (leader) {
      const uint32_t idesc = ptx::make_idesc_bf16(2 * DC_ROWS, DC_CHUNK, 0, 0);
      const uint64_t d_z = ptx::make_smem_desc_sw128(ptx::smem_u32(zbuf), 16, 1024);
      const uint64_t d_hd0 = ptx::make_smem_desc_sw128(ptx::smem_u32(hd0), 16, 1024);
      const uint64_t d_hd1 = ptx::make_smem_desc_sw128(ptx::smem_u32(hd1), 16, 1024);
      const uint64_t d_ring = ptx::make_smem_desc_sw128(ptx::smem_u32(ring), 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t chunk_n = 0;          // running count of accumulator chunks: buffer = chunk_n & 1, use number = chunk_n >> 1
      uint32_t l2_n = 0;             // running count of hd1 chunks consumed
      uint32_t tile_n = 0;
      // one weight block: 4 MMAs (K = 64) of A block `da` against ring stage; first MMA overwrites when !acc_first
      auto block = [&](uint64_t da, uint32_t tmem_d, bool acc_first) {
        wait(&b_full[stage], phase, 3, 0);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t db = d_ring + (uint64_t)((uint32_t)(stage * DC_STAGE_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < DC_BK / 16; ++kk)
            ptx::umma_f16_cg2(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (acc_first || kk > 0) ? 1u : 0u);
          ptx::umma_commit_cg2_mc(&b_empty[stage], (uint16_t)3);
        }
        __syncwarp();
        if (++stage == DC_STAGES) { stage = 0; phase ^= 1; }
      };
      auto commit = [&](uint64_t* bar) {
        if (ptx::elect_one()) ptx::umma_commit_cg2_mc(bar, (uint16_t)3);
        __syncwarp();
      };
      auto l2_chunk = [&](int cc) {
        if (cc == 0) {                   // the output accumulator of the previous tile has been drained
          wait(out_empty, (tile_n & 1) ^ 1, 8, 4);
          ptx::tc_fence_after();
        }
        for (int j = 0; j < 2; ++j) {
          wait(&hd1_ready[j], l2_n & 1, 9, 3);
          ptx::tc_fence_after();
          for (int h = 0; h < 2; ++h)
            block(d_hd1 + (uint64_t)((uint32_t)(j * DC_KBLOCK_BYTES) >> 4), tmem_base + 256u + (uint32_t)(h * DC_CHUNK), !(cc == 0 && j == 0));
        }
        commit(hd1_empty);
        ++l2_n;
      };
      for (int64_t tile = pair_id; tile < tiles; tile += pairs, ++tile_n) {
        // ---- L0
        wait(z_full, tile_n & 1, 4, 5);
        ptx::tc_fence_after();
        for (int c = 0; c < NC; ++c, ++chunk_n) {
          const uint32_t buf = chunk_n & 1;
          wait(&acc_empty[buf], ((chunk_n >> 1) & 1) ^ 1, 5, 1);
          ptx::tc_fence_after();
          block(d_z, tmem_base + buf * DC_CHUNK, false);
          commit(&acc_full[buf]);
        }
        commit(z_empty);
        // ---- L1, with the last layer's partial products one chunk behind
        for (int c = 0; c < NC; ++c, ++chunk_n) {
          const uint32_t buf = chunk_n & 1;
          wait(&acc_empty[buf], ((chunk_n >> 1) & 1) ^ 1, 6, 1);
          ptx::tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            if (c == 0) {
              wait(&hd0_ready[kb], tile_n & 1, 7, 2);
              ptx::tc_fence_after();
            }
            block(d_hd0 + (uint64_t)((uint32_t)(kb * DC_KBLOCK_BYTES) >> 4), tmem_base + buf * DC_CHUNK, kb > 0);
          }
          commit(&acc_full[buf]);
          if (c == NC - 1) commit(hd0_free);
          if (c >= 1) l2_chunk(c - 1);
        }
        l2_chunk(NC - 1);
        commit(out_full);
      }
    }
  } else if (warp < 2 + DC_EPI_WARPS) {
    // ============================ epilogue ============================
    const int ew = warp - 2;
    const int quarter = warp & 3;              // TMEM lane quarter this warp may read
    const int cg = ew >> 2;                    // 32-column group of the 128-column chunk
    const int r_in = quarter * 32 + lane;      // row inside this CTA's 128
    uint32_t chunk_n = 0, l1_n = 0, tile_n = 0;
    float sse = 0.f;                           // TRAIN: this thread's share of the sum of squared errors
    float cs_acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // TRAIN: column sums of d x_hat, this lane's two columns of each 32-column block
    // Output staging: 32 rows x 16 fp32 (64-byte rows, 2 KB).  It lives inside the 4 KB of the hd1 chunk that this warp and its partner
    // (same lane quarter, the other 32-column group of the same K-block) are the only writers of, so the only cross-warp ordering the
    // region needs is between those two: a 64-thread named barrier before the next tile's first hd1 write.
    const int pair = (cg >> 1) * 4 + (ew & 3);
    uint8_t* const stage_out = hd1 + (cg >> 1) * DC_KBLOCK_BYTES + quarter * 4096 + (cg & 1) * 2048;
    const uint32_t stage_out_s = ptx::smem_u32(stage_out);
    // One layer's NC accumulator chunks of one tile: L0 -> hd0, L1 -> the hd1 chunk buffer.  `tn` = this CTA pair's running index of `tile`.
    auto hidden = [&](const int layer, const int64_t tile, const uint32_t tile_n) {
      const int32_t row_base = (int32_t)(tile * 2 * DC_ROWS) + (int32_t)cta_rank * DC_ROWS + quarter * 32;
        for (int c = 0; c < NC; ++c, ++chunk_n) {
          const uint32_t buf = chunk_n & 1;
          wait(&acc_full[buf], (chunk_n >> 1) & 1, 10, 0);
          ptx::tc_fence_after();
          uint32_t r[32];
          ptx::tmem_ld_32x32_issue(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * DC_CHUNK + cg * 32), r);
          float bv[32];
          load_vec<32>((layer == 0 ? a.b0 : a.b1) + c * DC_CHUNK + cg * 32, bv);
          ptx::tmem_ld_wait(r);
          // the accumulator buffer is free as soon as its values are in registers
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(lead(&acc_empty[buf]));
          uint8_t* dst;
          if (layer == 0) {
            if (c == 0) wait(hd0_free, (tile_n & 1) ^ 1, 11, 1);       // the previous tile's L1 MMAs have read hd0
            dst = hd0 + (c * 2 + (cg >> 1)) * DC_KBLOCK_BYTES;
          } else {
            wait(hd1_empty, (l1_n & 1) ^ 1, 12, 2);                    // the previous chunk's L2 MMAs have read hd1
            dst = hd1 + (cg >> 1) * DC_KBLOCK_BYTES;
          }
          // the pair's TMA stores out of the region about to be rewritten (TRAIN: the activation write-out of the K-block's previous
          // use; sampling: the output staging blocks, which alias the hd1 chunk) have finished reading it
#ifdef DC_EXP_NOSTORE
          if (layer == 1 && c == 0) {
#else
          if (TRAIN || (layer == 1 && c == 0)) {
#endif
            if (lane == 0) {
              // TRAIN, first hidden layer: the K-block about to be rewritten was stored one whole tile ago -- 2 NC + 2 bulk groups back in
              // the issuing warp's sequence (NC + NC activation blocks, 2 output blocks per tile); only the hd1 chunk (rewritten every
              // chunk) and the sampling path's staging need the most recent store to have finished reading
              if (TRAIN && layer == 0) ptx::bulk_wait_read_n<4>();
              else ptx::bulk_wait_read0();
            }
            __syncwarp();
            asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
          }
#ifndef DC_EXP_NOMASK
          if constexpr (TRAIN) {
            // 1-bit ReLU mask of this warp's 32 columns, one word per row: what the backward pass multiplies the gradient with
            uint32_t bits = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) bits |= ((__uint_as_float(r[i]) + bv[i]) > 0.f ? 1u : 0u) << i;
            const int64_t grow = (int64_t)row_base + lane;
            if (grow < a.rows) (layer == 0 ? a.mask0 : a.mask1)[(int64_t)(c * 4 + cg) * a.mask_ld + grow] = bits;
          }
#endif
          {
            // bias add as packed fp32 pairs (FADD2), ReLU inside the bf16 pack, 16-byte stores through 32-bit shared addresses
            const uint32_t dst_s = ptx::smem_u32(dst);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                v[i] = __uint_as_float(r[8 * j + i]);
                v[i + 1] = __uint_as_float(r[8 * j + i + 1]);
                ptx::add2(v[i], v[i + 1], bv[8 * j + i], bv[8 * j + i + 1]);
              }
              ptx::sts128(dst_s + dc_swz128(r_in, (cg & 1) * 4 + j), ptx::pack_bf16x2_relu(v[0], v[1]), ptx::pack_bf16x2_relu(v[2], v[3]),
                          ptx::pack_bf16x2_relu(v[4], v[5]), ptx::pack_bf16x2_relu(v[6], v[7]));
            }
          }
          ptx::fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
          __syncwarp();
#ifndef DC_EXP_NOSTORE
          if constexpr (TRAIN) {
            // both halves of the 128-byte rows are in place: one TMA store of the pair's [32 rows x 64 columns] writes the activation out
            asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
            if ((cg & 1) == 0 && lane == 0) {
              ptx::tma_store_2d(layer == 0 ? &tm_hd0 : &tm_hd1, dst + quarter * 4096, (c * 2 + (cg >> 1)) * DC_BK, row_base);
              ptx::bulk_commit();
            }
          }
#endif
          if (lane == 0) ptx::mbar_arrive_cluster(lead(layer == 0 ? &hd0_ready[c * 2 + (cg >> 1)] : &hd1_ready[cg >> 1]));
          if (layer == 1) ++l1_n;
        }
    };
    auto output = [&](const int64_t tile, const uint32_t tile_n) {
      const int32_t row_base = (int32_t)(tile * 2 * DC_ROWS) + (int32_t)cta_rank * DC_ROWS + quarter * 32;
      // ---- output: 256 accumulator columns, this warp's 64 (4 sub-blocks of 16 columns through the 2 KB staging block)
      wait(out_full, tile_n & 1, 13, 3);
      ptx::tc_fence_after();
      const int ocg = ew >> 2;                   // 64-column group of the output
      if constexpr (TRAIN) {
        // the staging block aliases the hd1 chunk region whose write-out the pair's issuer may still be reading
        if (lane == 0) ptx::bulk_wait_read0();
        __syncwarp();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
        const int64_t grow = (int64_t)row_base + lane;
        const bool valid = grow < a.rows;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col0 = ocg * 64 + half * 32;
          if (col0 >= a.D) continue;             // warp-uniform
          uint32_t r[32];
          ptx::tmem_ld_32x32_issue(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(256 + col0), r);
          // the target tile: a bf16 batch is fetched NOW, packed (16 registers), so that its latency hides under the accumulator read; an fp32
          // batch is read 8 columns at a time inside the loop (the slower path: 32 more registers would spill)
          const bool full = valid && col0 + 32 <= a.D;
          uint4 xq[4] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
          if (full && a.x_bf16) {
            const uint4* xp = reinterpret_cast<const uint4*>(static_cast<const bf16*>(a.x) + grow * a.D + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) xq[j] = __ldg(xp + j);
          }
          ptx::tmem_ld_wait(r);
          if (lane == 0) ptx::bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float xv[8];
            if (full && a.x_bf16) {
              const uint32_t wv[4] = {xq[j].x, xq[j].y, xq[j].z, xq[j].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                xv[2 * q] = __uint_as_float(wv[q] << 16);
                xv[2 * q + 1] = __uint_as_float(wv[q] & 0xFFFF0000u);
              }
            } else if (full) {
              load_vec<8>(static_cast<const float*>(a.x) + grow * a.D + col0 + 8 * j, xv);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int cc = col0 + 8 * j + i;
                xv[i] = 0.f;
                if (valid && cc < a.D) xv[i] = a.x_bf16 ? __bfloat162float(static_cast<const bf16*>(a.x)[grow * a.D + cc]) : static_cast<const float*>(a.x)[grow * a.D + cc];
              }
            }
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int cc = col0 + 8 * j + i;
              const bool in = valid && cc < a.D;
              const float dd = in ? (__uint_as_float(r[8 * j + i]) + __ldg(a.b2 + min(cc, a.D - 1)) - xv[i]) : 0.f;
              sse = fmaf(dd, dd, sse);
              o[i] = dd * a.scale;
            }
            uint4 u;
            u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
            *reinterpret_cast<uint4*>(stage_out + dc_swz64(lane, j)) = u;
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          {   // column sums of the ROUNDED gradient (= bias gradient of the last layer): lanes 0-15 walk the even rows, 16-31 the odd rows
            const int hw = lane >> 4, w2 = lane & 15;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
              const int row = 2 * rr + hw;
              const uint32_t u = *reinterpret_cast<const uint32_t*>(stage_out + dc_swz64(row, w2 >> 2) + (w2 & 3) * 4);
              s0 += __uint_as_float(u << 16);
              s1 += __uint_as_float(u & 0xFFFF0000u);
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            cs_acc[half][0] += s0;
            cs_acc[half][1] += s1;
          }
          if (lane == 0) {
            ptx::tma_store_2d(&tm_out, stage_out, col0, row_base);
            ptx::bulk_commit();
          }
        }
      } else {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col0 = ocg * 64 + half * 32;
        if (col0 >= a.D) continue;               // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32_issue(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(256 + col0), r);
        float bv[32];
        if (col0 + 32 <= a.D) {          // arena offsets are multiples of 8 floats: 16-byte loads
          load_vec<32>(a.b2 + col0, bv);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) bv[i] = (col0 + i < a.D) ? __ldg(a.b2 + col0 + i) : 0.f;
        }
        ptx::tmem_ld_wait(r);
#pragma unroll
        for (int sb = 0; sb < 2; ++sb) {
          const int col = col0 + sb * 16;
          if (col >= a.D) continue;
          if (lane == 0) ptx::bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = sb * 16 + j * 4;
            float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]), v2 = __uint_as_float(r[i + 2]), v3 = __uint_as_float(r[i + 3]);
            ptx::add2(v0, v1, bv[i], bv[i + 1]);
            ptx::add2(v2, v3, bv[i + 2], bv[i + 3]);
            ptx::sts128(stage_out_s + dc_swz64(lane, j), __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tm_out, stage_out, col, row_base);
            ptx::bulk_commit();
          }
        }
      }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(lead(out_empty));
    };
    // Order: the NEXT tile's first hidden layer is drained before this tile's output, so that the second layer's MMAs of the next tile can
    // start while the output accumulator (which is not needed again before the next tile's second L2 chunk) is still being stored
    {
      int64_t tile = pair_id;
      if (tile < tiles) hidden(0, tile, 0u);
      for (; tile < tiles; tile += pairs, ++tile_n) {
        hidden(1, tile, tile_n);
        if (tile + pairs < tiles) hidden(0, tile + pairs, tile_n + 1);
        output(tile, tile_n);
      }
    }
    if constexpr (TRAIN) {
      // bias gradient of the last layer: this warp always owned the same 64 output columns
      if (lane < 16) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = (ew >> 2) * 64 + half * 32 + 2 * lane;
          if (col + 1 < a.D) {
            atomicAdd(a.bias_grad + col, cs_acc[half][0]);
            atomicAdd(a.bias_grad + col + 1, cs_acc[half][1]);
          } else if (col < a.D) {
            atomicAdd(a.bias_grad + col, cs_acc[half][0]);
          }
        }
      }
      const float ws = warp_sum(sse);
      if (lane == 0) red_smem[ew] = ws;
    }
    if (lane == 0) ptx::bulk_wait_all();
  } else {
    // ============================ z ============================
    const int zt = (warp - 2 - DC_EPI_WARPS) * 32 + lane;          // 0..63: this thread writes rows zt and zt + 64 of the CTA's 128
    uint32_t tile_n = 0;
    for (int64_t tile = pair_id; tile < tiles; tile += pairs, ++tile_n) {
      wait(z_empty, (tile_n & 1) ^ 1, 14, 0);               // the previous tile's L0 MMAs have read the block
#pragma unroll 1
      for (int rr = 0; rr < DC_ROWS / (32 * DC_Z_WARPS); ++rr) {
        const int zr = zt + rr * 32 * DC_Z_WARPS;
        const int64_t row = tile * 2 * DC_ROWS + (int64_t)cta_rank * DC_ROWS + zr;
        const bool valid = row < a.rows;
        const uint64_t q0 = (uint64_t)(a.first_row + row) * (DC_L / 4);
#pragma unroll 1
        for (int j = 0; j < DC_L / 8; ++j) {
          if constexpr (TRAIN) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (valid) u = __ldg(reinterpret_cast<const uint4*>(a.z16 + row * DC_L + 8 * j));
            *reinterpret_cast<uint4*>(zbuf + dc_swz128(zr, j)) = u;
            continue;
          }
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (valid) {
            if (a.z_in) {
              v0 = __ldg(reinterpret_cast<const float4*>(a.z_in + row * DC_L + 8 * j));
              v1 = __ldg(reinterpret_cast<const float4*>(a.z_in + row * DC_L + 8 * j + 4));
            } else {
              v0 = philox_normal4(q0 + 2 * j, a.seed, a.offset);
              v1 = philox_normal4(q0 + 2 * j + 1, a.seed, a.offset);
            }
            if (a.z_out) {
              *reinterpret_cast<float4*>(a.z_out + row * DC_L + 8 * j) = v0;
              *reinterpret_cast<float4*>(a.z_out + row * DC_L + 8 * j + 4) = v1;
            }
          }
          uint4 u;
          u.x = pack_bf16x2(v0.x, v0.y); u.y = pack_bf16x2(v0.z, v0.w); u.z = pack_bf16x2(v1.x, v1.y); u.w = pack_bf16x2(v1.z, v1.w);
          *reinterpret_cast<uint4*>(zbuf + dc_swz128(zr, j)) = u;
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(lead(z_full));
    }
  }

  // ============================ teardown ============================
  if constexpr (TRACE) if (lane == 0) {
    unsigned long long* t = a.trace + (size_t)blockIdx.x * 24;
    const long long total = clock64() - t_begin;
    // MMA warp: [0] total, [1] weights, [2] acc_empty, [3] hd0_ready, [4] hd1_ready, [5] out_empty, [6] z_full
    if (warp == 1) { t[0] = total; for (int i = 0; i < 6; ++i) t[1 + i] = tw[i]; }
    // epilogue warp 2: [8] total, [9] acc_full, [10] hd0_free, [11] hd1_empty, [12] out_full
    if (warp == 2) { t[8] = total; for (int i = 0; i < 4; ++i) t[9 + i] = tw[i]; }
    // producer: [16] total, [17] b_empty; z warp: [18] total, [19] z_empty
    if (warp == 0) { t[16] = total; t[17] = tw[0]; }
    if (warp == 2 + DC_EPI_WARPS) { t[18] = total; t[19] = tw[0]; }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();               // neither CTA may retire while the pair's MMAs / remote arrives can still touch it
  if constexpr (TRAIN) {
    if (threadIdx.x == 0) {          // fixed order: the loss is reproducible for a given grid
      float t = 0.f;
      for (int w = 0; w < DC_EPI_WARPS; ++w) t += red_smem[w];
      a.sse_part[blockIdx.x] = t;
    }
  }
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_cg2(tmem_base, 512);
  }
}

}  // namespace psvae

namespace psvae {

int tc_tensor_map(const struct TcOperand& op, int64_t K, int box_rows, CUtensorMap* out);
int tc_block_map(const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld, int64_t splits, int64_t split_stride, CUtensorMap* out, int box_cols);
int tc_grid_size();
void count_launch();

// can the chained kernel run this decoder?  (two hidden layers of H <= 512, latent 64, at most 256 outputs)
static inline bool decoder_chain_ok(int D, int L, int H, int num_hidden) {
  return num_hidden == 2 && L == DC_L && H % DC_CHUNK == 0 && H >= DC_CHUNK && H <= DC_MAX_H && D % 16 == 0 && D >= 16 && D <= 256;
}

// W0 [H][64], W1 [H][H], W2 [D][H]: bf16, row-major (the shadow copy of the parameter arena).  Sampling: out = x_hat [rows][D] fp32.
// TRAIN: out = d x_hat [rows][D] bf16, hd0 / hd1 = the hidden activations [rows][H] bf16 (written for the backward pass).  *ctas = CTAs launched
// (= sum-of-squared-error partials written to a.sse_part).
template <bool TRAIN>
static inline int decoder_chain_launch(const bf16* W0, const bf16* W1, const bf16* W2, void* out, bf16* hd0, bf16* hd1, const DcArgs& a, cudaStream_t st,
                                       int* ctas = nullptr) {
  CUtensorMap t0, t1, t2, to, th0, th1;
  TcOperand o0{W0, (int64_t)a.H, (int64_t)DC_L, false};
  TcOperand o1{W1, (int64_t)a.H, (int64_t)a.H, false};
  TcOperand o2{W2, (int64_t)a.D, (int64_t)a.H, false};
  PSVAE_TRY(tc_tensor_map(o0, DC_L, 64, &t0));
  PSVAE_TRY(tc_tensor_map(o1, a.H, 64, &t1));
  PSVAE_TRY(tc_tensor_map(o2, a.H, 64, &t2));
  if constexpr (TRAIN) {
    PSVAE_TRY(tc_block_map(out, 2, a.rows, a.D, a.D, 0, 0, &to, 32));
    PSVAE_TRY(tc_block_map(hd0, 2, a.rows, a.H, a.H, 0, 0, &th0, 64));
    PSVAE_TRY(tc_block_map(hd1, 2, a.rows, a.H, a.H, 0, 0, &th1, 64));
  } else {
    PSVAE_TRY(tc_block_map(out, 4, a.rows, a.D, a.D, 0, 0, &to, 16));
    th0 = to;
    th1 = to;
  }
  const int smem = dc_smem_bytes(a.H);
  auto kern = a.trace ? decoder_chain_kernel<TRAIN, true> : decoder_chain_kernel<TRAIN, false>;
  static unsigned long long attr_mask[2] = {0, 0};
  int dev = 0;
  PSVAE_CUDA(cudaGetDevice(&dev));
  if (!(attr_mask[a.trace ? 1 : 0] >> (dev & 63) & 1ull)) {
    PSVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dc_smem_bytes(DC_MAX_H)));
    attr_mask[a.trace ? 1 : 0] |= 1ull << (dev & 63);
  }
  const int64_t tiles = (a.rows + 2 * DC_ROWS - 1) / (2 * DC_ROWS);
  int64_t pairs = tc_grid_size() / 2;
  if (tiles < pairs) pairs = tiles;
  if (pairs < 1) return 0;
  if (ctas) *ctas = (int)(2 * pairs);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(DC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  at[na].id = cudaLaunchAttributeClusterDimension;
  at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  PSVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, t0, t1, t2, to, th0, th1, a));
  count_launch();
  PSVAE_LAUNCH_CHECK("decoder_chain_kernel");
  return 0;
}

}  // namespace psvae
