// Fused GEMM epilogues shared by the fp32 CUDA-core engine (sgemm.cuh) and the tcgen05 engine
// (gemm_tc.cuh).  An epilogue sees one row fragment at a time: NV consecutive accumulator columns
// of one output row, fully inside the matrix, and may add to a per-thread reduction value that the
// engine sums over the CTA and stores (deterministically, one slot per CTA / tile).
//
// They implement, in the producing kernel, the element-wise tail of each reference op:
//   EpiBiasAct   nn.Linear bias + ReLU (model.py:14-36) / classifier activations (latent_classifier.py:18-23)
//   EpiMse       mse_loss(x_hat, x)/10 and its gradient 2(x_hat-x)/(B*D*10) (lightning.py:113)
//   EpiActGrad   dgrad * act'(.) (autograd of ReLU etc.)
//   EpiStore     plain alpha*acc (+ beta*old): split-K wgrad partials, classifier dgrads
#pragma once
#include "common.cuh"

namespace psvae {

enum : int { ACT_NONE = -1, ACT_RELU = 0, ACT_TANH = 1, ACT_SIGMOID = 2, ACT_LEAKY = 3 };

template <int ACT> __device__ __forceinline__ float act_fwd(float u) {
  if constexpr (ACT == ACT_RELU) return fmaxf(u, 0.f);
  else if constexpr (ACT == ACT_TANH) return tanhf(u);
  else if constexpr (ACT == ACT_SIGMOID) return 1.f / (1.f + expf(-u));
  else if constexpr (ACT == ACT_LEAKY) return u > 0.f ? u : 0.01f * u;
  else return u;
}
// derivative expressed through the POST-activation value a (what the forward pass stored)
template <int ACT> __device__ __forceinline__ float act_grad_from_out(float a) {
  if constexpr (ACT == ACT_RELU) return a > 0.f ? 1.f : 0.f;
  else if constexpr (ACT == ACT_TANH) return 1.f - a * a;
  else if constexpr (ACT == ACT_SIGMOID) return a * (1.f - a);
  else if constexpr (ACT == ACT_LEAKY) return a > 0.f ? 1.f : 0.01f;
  else return 1.f;
}

template <typename TOut, int ACT>
struct EpiBiasAct {
  static constexpr bool kReduce = false;
  static constexpr bool kColSum = false;
  const float* bias;   // [N] or nullptr
  TOut* out;
  int64_t ldo;
  float* red_out;      // unused
  template <int NV>
  __device__ __forceinline__ void apply(int64_t row, int col, float (&v)[NV], float& red, int split) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = act_fwd<ACT>(v[i] + (bias ? __ldg(bias + col + i) : 0.f));
    store_vec<NV>(out + row * ldo + col, v);
  }
};

// CS: the tcgen05 engine also emits the column sums of the stored gradient (= the bias gradient of the last decoder layer)
template <typename TOut, bool CS = false>
struct EpiMse {
  static constexpr bool kReduce = true;
  static constexpr bool kColSum = CS;
  const float* bias;   // [N]
  const float* x;      // [M, ldx] fp32 target
  int64_t ldx;
  float* x_hat;        // optional fp32 output [M, ldxh]
  int64_t ldxh;
  TOut* dxh;           // optional gradient output (x_hat - x) * scale
  int64_t ldd;
  float scale;
  float* red_out;      // one slot per CTA: sum of squared differences
  float* colsum;       // CS: [CTAs][N] partial column sums of dxh
  template <int NV>
  __device__ __forceinline__ void apply(int64_t row, int col, float (&v)[NV], float& red, int split) const {
    float xv[NV];
    load_vec<NV>(x + row * ldx + col, xv);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __ldg(bias + col + i);
    if (x_hat) store_vec<NV>(x_hat + row * ldxh + col, v);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float d = v[i] - xv[i];
      red = fmaf(d, d, red);
      v[i] = d * scale;
    }
    if (dxh) store_vec<NV>(dxh + row * ldd + col, v);
  }
};

// CS: also emit the column sums of the stored gradient (= the bias gradient of the layer below)
template <typename TAct, typename TOut, int ACT, bool CS = false>
struct EpiActGrad {
  static constexpr bool kReduce = false;
  static constexpr bool kColSum = CS;
  const TAct* act;     // forward activation (post-activation) [M, lda]
  int64_t lda;
  TOut* out;
  int64_t ldo;
  float beta;          // out = acc * act'(.) + beta * out   (sum over classifier heads)
  float* red_out;
  float* colsum;       // CS: [CTAs][N] partial column sums of the output
  template <int NV>
  __device__ __forceinline__ void apply(int64_t row, int col, float (&v)[NV], float& red, int split) const {
    float a[NV];
    load_vec<NV>(act + row * lda + col, a);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= act_grad_from_out<ACT>(a[i]);
    if (beta != 0.f) {
      float o[NV];
      load_vec<NV>(out + row * ldo + col, o);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaf(beta, o[i], v[i]);
    }
    store_vec<NV>(out + row * ldo + col, v);
  }
};

struct EpiStore {
  static constexpr bool kReduce = false;
  static constexpr bool kColSum = false;
  float* out;
  int64_t ldo;
  int64_t split_stride;  // elements between split-K partials
  float alpha, beta;
  float* red_out;
  template <int NV>
  __device__ __forceinline__ void apply(int64_t row, int col, float (&v)[NV], float& red, int split) const {
    float* p = out + (int64_t)split * split_stride + row * ldo + col;
    if (beta != 0.f) {
      float o[NV];
      load_vec<NV>(p, o);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaf(alpha, v[i], beta * o[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] *= alpha;
    }
    store_vec<NV>(p, v);
  }
};

}  // namespace psvae
