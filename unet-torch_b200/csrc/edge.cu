// The steps either side of the network, as bandwidth-bound kernels:
//   input edge      DataLoader.py:661-671 / test_mc3serousv5.py:115-125: per-image, per-channel z-normalisation of the
//                   uint8 HWC (BGR) image, HWC -> CHW, BGR -> RGB, float32.     (numpy on the CPU in the reference)
//   inference heads test_mc3serousv5.py:879-887: OutConv (Model.py:86-92) + softmax(dim=1) + argmax(dim=1) + np.uint8, fused:
//                   the fp32 logits (20 B/pixel at 5 classes) never reach HBM, only the 1-byte class mask does;
//                   test_mc3serousv5.py:961-974: OutConv + F.relu + /200 (density maps) and their per-map sums (counts).
#include "../../include/b200unet.h"
#include "host_common.h"
#include "head_common.cuh"

#include <cuda_bf16.h>

namespace {

// ---------------------------------------------------------------------------------------------- z-normalisation
// Exact integer statistics: sums[(n*C + c)*2 + {0,1}] = sum x, sum x^2 over the H*W pixels (uint8 inputs -> no rounding).
// VEC: a thread takes 16 consecutive pixels = C 16-byte loads (needs H*W % 16 == 0 and a 16-byte aligned base).
template <int C, bool VEC>
__global__ void __launch_bounds__(256) znorm_stats_kernel(const uint8_t* __restrict__ img, unsigned long long* __restrict__ sums,
                                                         long long HW, int blocks_per_image) {
  const int n = blockIdx.x / blocks_per_image, b = blockIdx.x % blocks_per_image;
  const uint8_t* src = img + static_cast<long long>(n) * HW * C;
  unsigned long long s1[C], s2[C];
#pragma unroll
  for (int c = 0; c < C; ++c) s1[c] = s2[c] = 0;
  const long long t0 = static_cast<long long>(b) * blockDim.x + threadIdx.x;
  const long long tstride = static_cast<long long>(blocks_per_image) * blockDim.x;
  if (VEC) {
    for (long long g = t0; g < HW / 16; g += tstride) {
      uint4 v[C];
#pragma unroll
      for (int i = 0; i < C; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(src + g * 16 * C) + i);
      const uint8_t* by = reinterpret_cast<const uint8_t*>(v);
      unsigned a1[C], a2[C];  // 16 pixels: at most 16*255^2 per channel
#pragma unroll
      for (int c = 0; c < C; ++c) a1[c] = a2[c] = 0;
#pragma unroll
      for (int i = 0; i < 16 * C; ++i) {
        const unsigned x = by[i];
        a1[i % C] += x;
        a2[i % C] += x * x;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        s1[c] += a1[c];
        s2[c] += a2[c];
      }
    }
  } else {
    for (long long p = t0; p < HW; p += tstride) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const unsigned v = src[p * C + c];
        s1[c] += v;
        s2[c] += v * v;
      }
    }
  }
  __shared__ unsigned long long sh[8][2 * C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    unsigned long long a = s1[c], q = s2[c];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o);
      q += __shfl_down_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      sh[warp][2 * c] = a;
      sh[warp][2 * c + 1] = q;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    atomicAdd(sums + (static_cast<long long>(n) * C) * 2 + threadIdx.x, t);
  }
}

// A uint8 image has 256 distinct values, so (x - mean) / std is a 256-entry table per (image, channel): every block
// builds the table of its image in fp64 exactly as numpy evaluates it (np.mean = sum/count; np.std = sqrt of the
// population variance, here from the exact integer moments) and the per-pixel work is a lookup + float store.
// out[n][reverse ? C-1-c : c][p] = float32((x - mean) / std).  std == 0 gives inf/nan like the reference.
template <int C, bool VEC>
__global__ void __launch_bounds__(256) znorm_apply_kernel(const uint8_t* __restrict__ img,
                                                         const unsigned long long* __restrict__ sums,
                                                         float* __restrict__ out, long long HW, int blocks_per_image,
                                                         int reverse) {
  __shared__ float lut[C][256];
  const int n = blockIdx.x / blocks_per_image, b = blockIdx.x % blocks_per_image;
  {
    const double cnt = static_cast<double>(HW);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const unsigned long long s1 = sums[(static_cast<long long>(n) * C + c) * 2];
      const unsigned long long s2 = sums[(static_cast<long long>(n) * C + c) * 2 + 1];
      const double mean = static_cast<double>(s1) / cnt;
      // HW*s2 - s1^2 >= 0 exactly (Cauchy-Schwarz on integers); 128-bit keeps it exact for any image size
      const unsigned __int128 num = static_cast<unsigned __int128>(static_cast<unsigned long long>(HW)) * s2 -
                                    static_cast<unsigned __int128>(s1) * s1;
      const double hi = static_cast<double>(static_cast<unsigned long long>(num >> 64)) * 18446744073709551616.0;
      const double var = (hi + static_cast<double>(static_cast<unsigned long long>(num))) / (cnt * cnt);
      const double sd = sqrt(var);
      lut[c][threadIdx.x] = static_cast<float>((static_cast<double>(threadIdx.x) - mean) / sd);
    }
  }
  __syncthreads();
  const uint8_t* src = img + static_cast<long long>(n) * HW * C;
  float* dst = out + static_cast<long long>(n) * C * HW;
  const long long t0 = static_cast<long long>(b) * blockDim.x + threadIdx.x;
  const long long tstride = static_cast<long long>(blocks_per_image) * blockDim.x;
  if (VEC) {
    for (long long g = t0; g < HW / 16; g += tstride) {
      uint4 v[C];
#pragma unroll
      for (int i = 0; i < C; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(src + g * 16 * C) + i);
      const uint8_t* by = reinterpret_cast<const uint8_t*>(v);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int co = reverse ? C - 1 - c : c;
        float4* o = reinterpret_cast<float4*>(dst + co * HW + g * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          o[q] = make_float4(lut[c][by[(4 * q) * C + c]], lut[c][by[(4 * q + 1) * C + c]], lut[c][by[(4 * q + 2) * C + c]],
                             lut[c][by[(4 * q + 3) * C + c]]);
      }
    }
  } else {
    for (long long p = t0; p < HW; p += tstride) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int co = reverse ? C - 1 - c : c;
        dst[co * HW + p] = lut[c][src[p * C + c]];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- fused inference heads
using b2head::MAXC;

// head_fprop_kernel (small.cu) followed by softmax_argmax_kernel (loss.cu), bit-identical to the two-kernel path (the
// logits come from the same warp-cooperative routine): torch.softmax in fp32, then the first maximum (probabilities that
// round to the same float tie), then np.uint8.
// SIGMOID: test.py:393-399 instead - torch.sigmoid (fp32: 1 / (1 + exp(-z))) of channel 0, then >= threshold -> {0, 1}.
template <bool SIGMOID>
__global__ void __launch_bounds__(256) head_mask_kernel(const __nv_bfloat16* __restrict__ a, int a_cs,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       uint8_t* __restrict__ mask, long long P, int Cin, int ncls,
                                                       float threshold) {
  extern __shared__ __align__(16) float wsm[];
  b2head::load_weights(wsm, w, Cin, ncls);
  uint32_t* stage = b2head::warp_stage(wsm, Cin, ncls);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long g = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); g * 32 < P; g += warps) {
    float acc[MAXC];
    b2head::logits_warp32(a, a_cs, wsm, stage, bias, g * 32, P, Cin, ncls, acc);
    const long long p = g * 32 + lane;
    if (p >= P) continue;
    if (SIGMOID) {
      const float sg = 1.f / (1.f + expf(-acc[0]));
      mask[p] = sg >= threshold ? 1 : 0;
      continue;
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXC; ++j)
      if (j < ncls) m = fmaxf(m, acc[j]);
    float pr[MAXC], sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC; ++j)
      if (j < ncls) {
        pr[j] = expf(acc[j] - m);
        sum += pr[j];
      }
    int best = 0;
    float bv = pr[0] / sum;
#pragma unroll
    for (int j = 1; j < MAXC; ++j)
      if (j < ncls) {
        const float pj = pr[j] / sum;
        if (pj > bv || (pj != pj && bv == bv)) {
          bv = pj;
          best = j;
        }
      }
    mask[p] = static_cast<uint8_t>(best);
  }
}

// OutConv + F.relu + division by `divisor` (fp32, like numpy's float32 array / 200) -> fp32 NCHW density maps, and the
// per-(image, class) sums of the stored values (the cell counts) accumulated in fp64.
__global__ void __launch_bounds__(256) head_density_kernel(const __nv_bfloat16* __restrict__ a, int a_cs,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ out, double* __restrict__ counts, long long HW,
                                                          int blocks_per_image, int Cin, int ncls, float divisor) {
  extern __shared__ __align__(16) float wsm[];
  __shared__ double red[8][MAXC];
  b2head::load_weights(wsm, w, Cin, ncls);
  uint32_t* stage = b2head::warp_stage(wsm, Cin, ncls);
  __syncthreads();
  const int n = blockIdx.x / blocks_per_image, b = blockIdx.x % blocks_per_image;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const __nv_bfloat16* src = a + static_cast<long long>(n) * HW * a_cs;
  double part[MAXC];
#pragma unroll
  for (int j = 0; j < MAXC; ++j) part[j] = 0.0;
  for (long long g = static_cast<long long>(b) * 8 + warp; g * 32 < HW; g += static_cast<long long>(blocks_per_image) * 8) {
    float acc[MAXC];
    b2head::logits_warp32(src, a_cs, wsm, stage, bias, g * 32, HW, Cin, ncls, acc);
    const long long hw = g * 32 + lane;
    if (hw >= HW) continue;
#pragma unroll
    for (int j = 0; j < MAXC; ++j)
      if (j < ncls) {
        // F.relu keeps NaN; fmaxf would drop it
        const float r = (acc[j] > 0.f || acc[j] != acc[j]) ? acc[j] : 0.f;
        const float v = r / divisor;
        out[(static_cast<long long>(n) * ncls + j) * HW + hw] = v;
        part[j] += static_cast<double>(v);
      }
  }
  if (counts == nullptr) return;
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    double v = part[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < ncls) {
    double t = 0.0;
    for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
    atomicAdd(counts + static_cast<long long>(n) * ncls + threadIdx.x, t);
  }
}

template <int C>
int znorm_launch(const uint8_t* img, unsigned long long* sums, float* out, int N, long long HW, int reverse, cudaStream_t st) {
  const bool vec = HW % 16 == 0 && reinterpret_cast<uintptr_t>(img) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  const long long work = vec ? HW / 16 : HW;  // thread iterations per image
  // enough blocks per image to fill 148 SMs x 8 resident blocks over the batch, at least 2 iterations per thread
  long long bpi = (148 * 8 + N - 1) / N;
  const long long cap = (work + 2 * 256 - 1) / (2 * 256);
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  const int grid = static_cast<int>(N * bpi), b = static_cast<int>(bpi);
  if (vec) {
    znorm_stats_kernel<C, true><<<grid, 256, 0, st>>>(img, sums, HW, b);
    if (int er = b2h::check_launch("znorm_stats")) return er;
    znorm_apply_kernel<C, true><<<grid, 256, 0, st>>>(img, sums, out, HW, b, reverse);
  } else {
    znorm_stats_kernel<C, false><<<grid, 256, 0, st>>>(img, sums, HW, b);
    if (int er = b2h::check_launch("znorm_stats")) return er;
    znorm_apply_kernel<C, false><<<grid, 256, 0, st>>>(img, sums, out, HW, b, reverse);
  }
  return b2h::check_launch("znorm_apply");
}

}  // namespace

extern "C" {

int64_t b200unet_znorm_workspace_bytes(int N, int C) { return static_cast<int64_t>(N) * C * 2 * 8; }

int b200unet_znorm_to_chw(const uint8_t* img_nhwc, void* workspace, float* out_nchw, int N, int H, int W, int C,
                          int reverse_channels, b200_stream_t stream) {
  B2_REQUIRE(C >= 1 && C <= 4, "znorm_to_chw: C=%d must be in [1,4]", C);
  B2_REQUIRE(N > 0 && H > 0 && W > 0, "znorm_to_chw: empty image batch (N=%d, H=%d, W=%d)", N, H, W);
  B2_REQUIRE(static_cast<long long>(H) * W <= (1ll << 40), "znorm_to_chw: image too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* sums = static_cast<unsigned long long*>(workspace);
  cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(b200unet_znorm_workspace_bytes(N, C)), st);
  if (e != cudaSuccess) {
    b2h::set_error("znorm_to_chw memset: %s", cudaGetErrorString(e));
    return 2;
  }
  const long long HW = static_cast<long long>(H) * W;
  switch (C) {
    case 1: return znorm_launch<1>(img_nhwc, sums, out_nchw, N, HW, reverse_channels, st);
    case 2: return znorm_launch<2>(img_nhwc, sums, out_nchw, N, HW, reverse_channels, st);
    case 3: return znorm_launch<3>(img_nhwc, sums, out_nchw, N, HW, reverse_channels, st);
    default: return znorm_launch<4>(img_nhwc, sums, out_nchw, N, HW, reverse_channels, st);
  }
}

int b200unet_head_mask(const void* a, int a_cs, const float* w, const float* bias, uint8_t* mask, int N, int H, int W,
                       int Cin, int ncls, b200_stream_t stream) {
  B2_REQUIRE(ncls >= 1 && ncls <= MAXC, "head_mask: n_classes=%d must be in [1,%d]", ncls, MAXC);
  B2_REQUIRE(Cin > 0 && Cin % 64 == 0 && a_cs % 8 == 0 && a_cs >= Cin, "head_mask: Cin=%d must be a multiple of 64 (pitch %d)", Cin, a_cs);
  B2_REQUIRE(b2head::smem_bytes(Cin, ncls, 8) <= 48 * 1024, "head_mask: Cin=%d x n_classes=%d weights do not fit shared memory", Cin, ncls);
  const long long P = static_cast<long long>(N) * H * W;
  long long blocks = (P + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  head_mask_kernel<false><<<static_cast<int>(blocks), 256, b2head::smem_bytes(Cin, ncls, 8), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), a_cs, w, bias, mask, P, Cin, ncls, 0.f);
  return b2h::check_launch("head_mask");
}

int b200unet_head_sigmoid_mask(const void* a, int a_cs, const float* w, const float* bias, uint8_t* mask, int N, int H, int W,
                               int Cin, int ncls, float threshold, b200_stream_t stream) {
  B2_REQUIRE(ncls >= 1 && ncls <= MAXC, "head_sigmoid_mask: n_classes=%d must be in [1,%d]", ncls, MAXC);
  B2_REQUIRE(Cin > 0 && Cin % 64 == 0 && a_cs % 8 == 0 && a_cs >= Cin, "head_sigmoid_mask: Cin=%d must be a multiple of 64 (pitch %d)", Cin, a_cs);
  B2_REQUIRE(b2head::smem_bytes(Cin, 1, 8) <= 48 * 1024, "head_sigmoid_mask: Cin=%d weights do not fit shared memory", Cin);
  const long long P = static_cast<long long>(N) * H * W;
  long long blocks = (P + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  // only channel 0 is thresholded (out[0, 0] in the reference): its weight row and bias come first, so ncls = 1 suffices
  head_mask_kernel<true><<<static_cast<int>(blocks), 256, b2head::smem_bytes(Cin, 1, 8), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), a_cs, w, bias, mask, P, Cin, 1, threshold);
  return b2h::check_launch("head_sigmoid_mask");
}

int b200unet_head_density(const void* a, int a_cs, const float* w, const float* bias, float* out_nchw, double* counts,
                          int N, int H, int W, int Cin, int ncls, float divisor, b200_stream_t stream) {
  B2_REQUIRE(ncls >= 1 && ncls <= MAXC, "head_density: n_classes=%d must be in [1,%d]", ncls, MAXC);
  B2_REQUIRE(Cin > 0 && Cin % 64 == 0 && a_cs % 8 == 0 && a_cs >= Cin, "head_density: Cin=%d must be a multiple of 64 (pitch %d)", Cin, a_cs);
  B2_REQUIRE(b2head::smem_bytes(Cin, ncls, 8) <= 48 * 1024, "head_density: Cin=%d x n_classes=%d weights do not fit shared memory", Cin, ncls);
  B2_REQUIRE(N > 0 && H > 0 && W > 0, "head_density: empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long HW = static_cast<long long>(H) * W;
  if (counts != nullptr) {
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(double) * N * ncls, st);
    if (e != cudaSuccess) {
      b2h::set_error("head_density memset: %s", cudaGetErrorString(e));
      return 2;
    }
  }
  long long bpi = (148 * 8 + N - 1) / N;
  const long long cap = (HW + 255) / 256;
  if (bpi > cap) bpi = cap;
  head_density_kernel<<<static_cast<int>(N * bpi), 256, b2head::smem_bytes(Cin, ncls, 8), st>>>(
      static_cast<const __nv_bfloat16*>(a), a_cs, w, bias, out_nchw, counts, HW, static_cast<int>(bpi), Cin, ncls, divisor);
  return b2h::check_launch("head_density");
}

}  // extern "C"
