// Fused optimizer step for the U-Net parameters (reference: train.py:341-347 builds
// torch.optim.SGD, Trainer.py:719-725 steps it every iteration). One pass over a conv / convT weight does
//   g' = g + wd*w ;  buf = mom*buf + (1-damp)*g'  (buf = g' on the first step) ;  w -= lr * (nesterov ? g' + mom*buf : buf)
// in fp32 on the master parameter AND writes the two bf16 GEMM operands the tensor-core kernels consume (fprop and
// dgrad layouts), so the separate re-cast pass (prep_conv3_kernel) and torch's foreach kernels disappear.
// The same pass exists for Adam (train.py:341-343 `optim.Adam(model.parameters(), lr, weight_decay)`, the optimizer of
// configseros.yml:15 and of Trainer.multi_task_uc_train): torch.optim.Adam arithmetic (L2 weight decay added to the gradient,
// bias-corrected first / second moments, no amsgrad) with the bias corrections folded into two host-computed scalars.
// All global accesses are coalesced: the fp32 tensors are walked in their own order, the bf16 tile is transposed
// through shared memory and written as 128-byte rows (operand 1) and 32-byte runs (operand 2).
#include "../../include/b200unet.h"
#include "host_common.h"

#include <cuda_bf16.h>

namespace {

struct SgdHyper {
  float lr, momentum, dampening, weight_decay;
  int nesterov, first_step;
  // Adam (adam != 0): exp_avg lives in `buf`, exp_avg_sq in `buf2`
  int adam;
  float beta1, beta2, eps;
  float omb1, omb2;     // 1 - beta1, 1 - beta2 rounded from the host's doubles (1.f - 0.999f is off by 5e-5 relative)
  float step_size;      // lr / (1 - beta1^t)
  float inv_sqrt_bc2;   // 1 / sqrt(1 - beta2^t)
};

// torch.optim.Adam (single-tensor path): g += wd*w; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
// w -= step_size * m / (sqrt(v) / sqrt(bc2) + eps)
__device__ __forceinline__ float adam_update(float& w, float g, float& m, float& v, const SgdHyper& h) {
  g = fmaf(h.weight_decay, w, g);
  m = fmaf(g - m, h.omb1, m);
  v = fmaf(h.beta2, v, h.omb2 * g * g);
  const float denom = fmaf(sqrtf(v), h.inv_sqrt_bc2, h.eps);
  w = fmaf(-h.step_size, m / denom, w);
  return w;
}

__device__ __forceinline__ float sgd_update(float& w, float g, float& buf, const SgdHyper& h) {
  g = fmaf(h.weight_decay, w, g);
  if (h.momentum != 0.f) {
    buf = h.first_step ? g : fmaf(h.momentum, buf, (1.f - h.dampening) * g);
    g = h.nesterov ? fmaf(h.momentum, buf, g) : buf;
  }
  w = fmaf(-h.lr, g, w);
  return w;
}

// Parameter tensor w[A][B][TAPS] (fp32). Tile = 32 a x 64 b x TAPS, walked as float4's in the tensor's own order.
//   op1[a][tap][b]                                   (conv3: fprop operand [K][rs][C]; convT: dgrad operand [ci][ij][d])
//   ROT ? op2[b][TAPS-1-tap][a] : op2[tap][b][a]     (conv3: dgrad operand [C][8-rs][K]; convT: fprop operand [ij][d][ci])
template <int TAPS, bool ROT>
__device__ __forceinline__ void weight_tile(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ buf,
                                            float* __restrict__ buf2, __nv_bfloat16* __restrict__ op1,
                                            __nv_bfloat16* __restrict__ op2, int A, int B, int a0, int b0, const SgdHyper& h) {
  constexpr int TA = 32, TB = 64;
  constexpr int ROW4 = TB * TAPS / 4;  // float4's per a-row of the tile (contiguous in the parameter tensor)
  __shared__ __align__(16) __nv_bfloat16 sm[TA][TAPS][TB];
  const bool have_buf = (buf != nullptr) && !h.first_step;
  for (int v = threadIdx.x; v < TA * ROW4; v += 256) {
    const int aa = v / ROW4, r4 = v - aa * ROW4;
    const size_t gi = (static_cast<size_t>(a0 + aa) * B + b0) * TAPS + static_cast<size_t>(r4) * 4;
    float4 w4 = *reinterpret_cast<const float4*>(w + gi);
    if (g != nullptr) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + gi));
      float4 b4 = have_buf ? *reinterpret_cast<const float4*>(buf + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (h.adam) {
        float4 v4 = h.first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(buf2 + gi);
        adam_update(w4.x, g4.x, b4.x, v4.x, h);
        adam_update(w4.y, g4.y, b4.y, v4.y, h);
        adam_update(w4.z, g4.z, b4.z, v4.z, h);
        adam_update(w4.w, g4.w, b4.w, v4.w, h);
        *reinterpret_cast<float4*>(buf2 + gi) = v4;
      } else {
        sgd_update(w4.x, g4.x, b4.x, h);
        sgd_update(w4.y, g4.y, b4.y, h);
        sgd_update(w4.z, g4.z, b4.z, h);
        sgd_update(w4.w, g4.w, b4.w, h);
      }
      *reinterpret_cast<float4*>(w + gi) = w4;
      if (buf != nullptr) *reinterpret_cast<float4*>(buf + gi) = b4;
    }
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = r4 * 4 + e;
      const int bb = r / TAPS, tap = r - bb * TAPS;
      sm[aa][tap][bb] = __float2bfloat16_rn(wv[e]);
    }
  }
  __syncthreads();
  // operand 1: rows of 64 b (128 B), 8 threads per row
  for (int idx = threadIdx.x; idx < TA * TAPS * 8; idx += 256) {
    const int row = idx >> 3, ch = idx & 7;
    const int aa = row / TAPS, tap = row - aa * TAPS;
    *reinterpret_cast<uint4*>(op1 + (static_cast<size_t>(a0 + aa) * TAPS + tap) * B + b0 + ch * 8) =
        *reinterpret_cast<const uint4*>(&sm[aa][tap][ch * 8]);
  }
  // operand 2: runs of 32 a (64 B) per (b, tap), four 16-byte stores each
  if (op2 != nullptr) {
    for (int idx = threadIdx.x; idx < TB * TAPS * 4; idx += 256) {
      const int q = idx & 3, pr = idx >> 2;
      const int bb = pr % TB, tap = pr / TB;
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = sm[q * 8 + i][tap][bb];
      const size_t row = ROT ? (static_cast<size_t>(b0 + bb) * TAPS + (TAPS - 1 - tap))
                             : (static_cast<size_t>(tap) * B + b0 + bb);
      *reinterpret_cast<uint4*>(op2 + row * A + a0 + q * 8) = *reinterpret_cast<const uint4*>(v);
    }
  }
}

template <int TAPS, bool ROT>
__global__ void __launch_bounds__(256) sgd_weight_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                         float* __restrict__ buf, float* __restrict__ buf2,
                                                         __nv_bfloat16* __restrict__ op1, __nv_bfloat16* __restrict__ op2,
                                                         int A, int B, SgdHyper h) {
  weight_tile<TAPS, ROT>(w, g, buf, buf2, op1, op2, A, B, blockIdx.y * 32, blockIdx.x * 64, h);
}

// All conv (or all convT) weights of the network in ONE launch: most of the 21 weight tensors are small (64x64x9 = two
// tiles), so one launch per tensor is latency-bound; a flat tile list keeps every SM busy with the few large ones.
constexpr int MAX_MULTI = 24;
struct WeightTable {
  float* w[MAX_MULTI];
  const float* g[MAX_MULTI];
  float* buf[MAX_MULTI];
  float* buf2[MAX_MULTI];
  __nv_bfloat16* op1[MAX_MULTI];
  __nv_bfloat16* op2[MAX_MULTI];
  int A[MAX_MULTI], B[MAX_MULTI];
  int tile_start[MAX_MULTI + 1];
  int count;
};

template <int TAPS, bool ROT>
__global__ void __launch_bounds__(256) sgd_weight_multi_kernel(const __grid_constant__ WeightTable t, SgdHyper h) {
  int ti = 0;
  while (ti + 1 < t.count && static_cast<int>(blockIdx.x) >= t.tile_start[ti + 1]) ++ti;
  const int lt = blockIdx.x - t.tile_start[ti];
  const int tiles_b = t.B[ti] / 64;
  weight_tile<TAPS, ROT>(t.w[ti], t.g[ti], t.buf[ti], t.buf2[ti], t.op1[ti], t.op2[ti], t.A[ti], t.B[ti], (lt / tiles_b) * 32,
                         (lt % tiles_b) * 64, h);
}

// small tensors (BatchNorm affine, biases, head, inc.conv1): up to 48 per launch, one block column per tensor
struct SmallTable {
  float* w[48];
  const float* g[48];
  float* buf[48];
  float* buf2[48];
  int n[48];
  int count;
};

__global__ void __launch_bounds__(256) sgd_small_kernel(SmallTable t, SgdHyper h) {
  const int ti = blockIdx.y;
  if (ti >= t.count) return;
  float* w = t.w[ti];
  const float* g = t.g[ti];
  float* buf = t.buf[ti];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.n[ti]; i += gridDim.x * blockDim.x) {
    float wv = w[i], bv = (buf != nullptr && !h.first_step) ? buf[i] : 0.f;
    if (h.adam) {
      float vv = h.first_step ? 0.f : t.buf2[ti][i];
      adam_update(wv, g[i], bv, vv, h);
      t.buf2[ti][i] = vv;
    } else {
      sgd_update(wv, g[i], bv, h);
    }
    w[i] = wv;
    if (buf != nullptr) buf[i] = bv;
  }
}

SgdHyper sgd_hyper(float lr, float momentum, float dampening, float weight_decay, int nesterov, int first_step) {
  SgdHyper h{};
  h.lr = lr;
  h.momentum = momentum;
  h.dampening = dampening;
  h.weight_decay = weight_decay;
  h.nesterov = nesterov;
  h.first_step = first_step;
  return h;
}

SgdHyper adam_hyper(double beta1, double beta2, float eps, float weight_decay, float step_size, float inv_sqrt_bc2, int first_step) {
  SgdHyper h{};
  h.adam = 1;
  h.beta1 = static_cast<float>(beta1);
  h.beta2 = static_cast<float>(beta2);
  h.omb1 = static_cast<float>(1.0 - beta1);
  h.omb2 = static_cast<float>(1.0 - beta2);
  h.eps = eps;
  h.weight_decay = weight_decay;
  h.step_size = step_size;
  h.inv_sqrt_bc2 = inv_sqrt_bc2;
  h.first_step = first_step;
  h.momentum = 1.f;  // "has a first-moment buffer"
  return h;
}

int launch_small(float* const* w, const float* const* grad, float* const* buf, float* const* buf2, const int* numel, int count,
                 const SgdHyper& h, cudaStream_t st, const char* what) {
  for (int base = 0; base < count; base += 48) {
    SmallTable t;
    t.count = count - base < 48 ? count - base : 48;
    int maxn = 1;
    for (int i = 0; i < t.count; ++i) {
      t.w[i] = w[base + i];
      t.g[i] = grad[base + i];
      t.buf[i] = buf ? buf[base + i] : nullptr;
      t.buf2[i] = buf2 ? buf2[base + i] : nullptr;
      t.n[i] = numel[base + i];
      if (t.n[i] > maxn) maxn = t.n[i];
    }
    int bx = (maxn + 255) / 256;
    if (bx > 296) bx = 296;  // a few multi-million-element tensors (the gates' ConvTranspose2d weights) ride along here
    sgd_small_kernel<<<dim3(bx, t.count), 256, 0, st>>>(t, h);
    if (int e = b2h::check_launch(what)) return e;
  }
  return 0;
}

// kind 0: conv3x3 weights [K][C][3][3] (op1 = fprop operand, op2 = dgrad operand); kind 1: ConvTranspose2d weights
// [Cin][Cup][2][2] (op1 = dgrad operand, op2 = fprop operand)
int launch_multi(int kind, float* const* w, const float* const* grad, float* const* buf, float* const* buf2,
                 void* const* w_fprop, void* const* w_dgrad, const int* dim_a, const int* dim_b, int count, const SgdHyper& h,
                 cudaStream_t st, const char* what) {
  if (count < 0 || (kind != 0 && kind != 1)) {
    b2h::set_error("%s: bad kind / count", what);
    return 1;
  }
  for (int base = 0; base < count; base += MAX_MULTI) {
    WeightTable t;
    t.count = count - base < MAX_MULTI ? count - base : MAX_MULTI;
    int tiles = 0;
    for (int i = 0; i < t.count; ++i) {
      const int j = base + i;
      if (dim_a[j] % 32 != 0 || dim_b[j] % 64 != 0 || w[j] == nullptr || grad[j] == nullptr || w_fprop[j] == nullptr ||
          w_dgrad[j] == nullptr || (h.momentum != 0.f && buf[j] == nullptr) || (h.adam && buf2[j] == nullptr)) {
        b2h::set_error("%s: tensor %d: dims (%d, %d) must be multiples of (32, 64) and no pointer may be null", what, j, dim_a[j],
                       dim_b[j]);
        return 1;
      }
      t.w[i] = w[j];
      t.g[i] = grad[j];
      t.buf[i] = buf ? buf[j] : nullptr;
      t.buf2[i] = buf2 ? buf2[j] : nullptr;
      t.op1[i] = static_cast<__nv_bfloat16*>(kind == 0 ? w_fprop[j] : w_dgrad[j]);
      t.op2[i] = static_cast<__nv_bfloat16*>(kind == 0 ? w_dgrad[j] : w_fprop[j]);
      t.A[i] = dim_a[j];
      t.B[i] = dim_b[j];
      t.tile_start[i] = tiles;
      tiles += (dim_a[j] / 32) * (dim_b[j] / 64);
    }
    t.tile_start[t.count] = tiles;
    if (tiles == 0) continue;
    if (kind == 0)
      sgd_weight_multi_kernel<9, true><<<tiles, 256, 0, st>>>(t, h);
    else
      sgd_weight_multi_kernel<4, false><<<tiles, 256, 0, st>>>(t, h);
    if (int e = b2h::check_launch(what)) return e;
  }
  return 0;
}

}  // namespace

extern "C" {

int b200unet_sgd_weights(int kind, float* const* w, const float* const* grad, float* const* momentum_buf, void* const* w_fprop,
                         void* const* w_dgrad, const int* dim_a, const int* dim_b, int count, float lr, float momentum,
                         float dampening, float weight_decay, int nesterov, int first_step, b200_stream_t stream) {
  return launch_multi(kind, w, grad, momentum_buf, nullptr, w_fprop, w_dgrad, dim_a, dim_b, count,
                      sgd_hyper(lr, momentum, dampening, weight_decay, nesterov, first_step), static_cast<cudaStream_t>(stream),
                      "sgd_weights");
}

int b200unet_adam_weights(int kind, float* const* w, const float* const* grad, float* const* exp_avg, float* const* exp_avg_sq,
                          void* const* w_fprop, void* const* w_dgrad, const int* dim_a, const int* dim_b, int count, double beta1,
                          double beta2, float eps, float weight_decay, float step_size, float inv_sqrt_bc2, int first_step,
                          b200_stream_t stream) {
  return launch_multi(kind, w, grad, exp_avg, exp_avg_sq, w_fprop, w_dgrad, dim_a, dim_b, count,
                      adam_hyper(beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2, first_step),
                      static_cast<cudaStream_t>(stream), "adam_weights");
}

int b200unet_sgd_conv3x3_weight(float* w_oihw, const float* grad, float* momentum_buf, void* w_fprop, void* w_dgrad,
                                int K, int C, float lr, float momentum, float dampening, float weight_decay,
                                int nesterov, int first_step, b200_stream_t stream) {
  B2_REQUIRE(K % 32 == 0 && C % 64 == 0, "sgd_conv3x3_weight: K=%d must be a multiple of 32 and C=%d of 64", K, C);
  B2_REQUIRE(w_oihw != nullptr && w_fprop != nullptr, "sgd_conv3x3_weight: null parameter / operand");
  const SgdHyper h = sgd_hyper(lr, momentum, dampening, weight_decay, nesterov, first_step);
  dim3 grid(C / 64, K / 32);
  sgd_weight_kernel<9, true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, grad, momentum_buf, nullptr, static_cast<__nv_bfloat16*>(w_fprop), static_cast<__nv_bfloat16*>(w_dgrad), K, C, h);
  return b2h::check_launch("sgd_conv3x3_weight");
}

int b200unet_sgd_convt2x2_weight(float* w, const float* grad, float* momentum_buf, void* w_fprop, void* w_dgrad,
                                 int Cin, int Cup, float lr, float momentum, float dampening, float weight_decay,
                                 int nesterov, int first_step, b200_stream_t stream) {
  B2_REQUIRE(Cin % 32 == 0 && Cup % 64 == 0, "sgd_convt2x2_weight: Cin=%d must be a multiple of 32 and Cup=%d of 64", Cin, Cup);
  B2_REQUIRE(w != nullptr && w_dgrad != nullptr, "sgd_convt2x2_weight: null parameter / operand");
  const SgdHyper h = sgd_hyper(lr, momentum, dampening, weight_decay, nesterov, first_step);
  dim3 grid(Cup / 64, Cin / 32);
  // parameter [Cin][Cup][4]: operand 1 = dgrad operand [ci][ij][d], operand 2 = fprop operand [ij][d][ci]
  sgd_weight_kernel<4, false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, grad, momentum_buf, nullptr, static_cast<__nv_bfloat16*>(w_dgrad), static_cast<__nv_bfloat16*>(w_fprop), Cin, Cup, h);
  return b2h::check_launch("sgd_convt2x2_weight");
}

int b200unet_sgd_small(float* const* w, const float* const* grad, float* const* momentum_buf, const int* numel,
                       int count, float lr, float momentum, float dampening, float weight_decay, int nesterov,
                       int first_step, b200_stream_t stream) {
  B2_REQUIRE(count >= 0, "sgd_small: negative count");
  return launch_small(w, grad, momentum_buf, nullptr, numel, count,
                      sgd_hyper(lr, momentum, dampening, weight_decay, nesterov, first_step), static_cast<cudaStream_t>(stream),
                      "sgd_small");
}

int b200unet_adam_conv3x3_weight(float* w_oihw, const float* grad, float* exp_avg, float* exp_avg_sq, void* w_fprop,
                                 void* w_dgrad, int K, int C, double beta1, double beta2, float eps, float weight_decay,
                                 float step_size, float inv_sqrt_bc2, int first_step, b200_stream_t stream) {
  B2_REQUIRE(K % 32 == 0 && C % 64 == 0, "adam_conv3x3_weight: K=%d must be a multiple of 32 and C=%d of 64", K, C);
  B2_REQUIRE(w_oihw && grad && exp_avg && exp_avg_sq && w_fprop, "adam_conv3x3_weight: null argument");
  const SgdHyper h = adam_hyper(beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2, first_step);
  sgd_weight_kernel<9, true><<<dim3(C / 64, K / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(w_fprop), static_cast<__nv_bfloat16*>(w_dgrad), K, C, h);
  return b2h::check_launch("adam_conv3x3_weight");
}

int b200unet_adam_convt2x2_weight(float* w, const float* grad, float* exp_avg, float* exp_avg_sq, void* w_fprop, void* w_dgrad,
                                  int Cin, int Cup, double beta1, double beta2, float eps, float weight_decay, float step_size,
                                  float inv_sqrt_bc2, int first_step, b200_stream_t stream) {
  B2_REQUIRE(Cin % 32 == 0 && Cup % 64 == 0, "adam_convt2x2_weight: Cin=%d must be a multiple of 32 and Cup=%d of 64", Cin, Cup);
  B2_REQUIRE(w && grad && exp_avg && exp_avg_sq && w_dgrad, "adam_convt2x2_weight: null argument");
  const SgdHyper h = adam_hyper(beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2, first_step);
  sgd_weight_kernel<4, false><<<dim3(Cup / 64, Cin / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(w_dgrad), static_cast<__nv_bfloat16*>(w_fprop), Cin, Cup, h);
  return b2h::check_launch("adam_convt2x2_weight");
}

int b200unet_adam_small(float* const* w, const float* const* grad, float* const* exp_avg, float* const* exp_avg_sq,
                        const int* numel, int count, double beta1, double beta2, float eps, float weight_decay, float step_size,
                        float inv_sqrt_bc2, int first_step, b200_stream_t stream) {
  B2_REQUIRE(count >= 0 && exp_avg && exp_avg_sq, "adam_small: bad arguments");
  return launch_small(w, grad, exp_avg, exp_avg_sq, numel, count,
                      adam_hyper(beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2, first_step),
                      static_cast<cudaStream_t>(stream), "adam_small");
}

}  // extern "C"
