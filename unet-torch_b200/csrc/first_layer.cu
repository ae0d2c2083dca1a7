// inc.conv1 = nn.Conv2d(n_channels <= 4, 64, 3, padding=1, bias=False) (reference Model.py:15-16 via :111) on the
// tensor cores: its reduction is only K = 9 * n_channels <= 36 long, so the network input (fp32 NCHW, as
// Trainer.py:700-702 hands it over) is expanded ONCE into a 64-column bf16 im2col tensor
//     col[n, h, w, c*9 + r*3 + s] = x[n, c, h + r - 1, w + s - 1]   (zero outside the image; columns >= 9*Cin are 0)
// which is one 128-byte swizzle row per pixel. The convolution is then a 1x1 GEMM over `col` (persistent
// resident-weight kernel, conv3_res.cu, TAPS = 1, same BatchNorm-statistics epilogue), and its weight gradient a
// plain pixel-reduction GEMM (wgrad.cu KIND_PLAIN). Both are bound by the 128 B/pixel they stream, not by math.
// Since round 2 the engine uses the FUSED forms (b200unet_conv3x3_first_igemm / _first_tc_wgrad): the same GEMMs with the
// im2col rows built in shared memory (first_tile.cuh), so `col` never exists in HBM; the materialising kernel below stays as
// an entry point (and as the reference the fused path is tested against).
#include "../../include/b200unet.h"
#include "host_common.h"

#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Block = one segment of SEG pixels of one image row. The 3 x CIN input rows (with the one-pixel halo, zero outside
// the image) are staged in smem with coalesced loads; then one thread-item = one pixel x one 16-byte chunk (8 im2col
// columns), chunk index fastest => coalesced 128-bit stores.
constexpr int SEG = 256;
template <int CIN>
__global__ void __launch_bounds__(256) first_im2col_kernel(const float* __restrict__ x, uint4* __restrict__ col,
                                                           long long col_pitch16, int N, int H, int W, int segs) {
  constexpr int T = CIN * 9;
  __shared__ float xs[CIN * 3][SEG + 2];
  const int seg = blockIdx.x % segs;
  const int row = blockIdx.x / segs;  // n * H + h
  const int h = row % H;
  const long long n = row / H;
  const int w0 = seg * SEG;
  for (int i = threadIdx.x; i < CIN * 3 * (SEG + 2); i += 256) {
    const int cr = i / (SEG + 2), wi = i - cr * (SEG + 2);
    const int c = cr / 3, r = cr - c * 3;
    const int hh = h + r - 1, ww = w0 + wi - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(x + ((n * CIN + c) * H + hh) * static_cast<long long>(W) + ww);
    xs[cr][wi] = v;
  }
  __syncthreads();
  uint4* dst = col + (static_cast<long long>(row) * W + w0) * col_pitch16;
#pragma unroll
  for (int it = 0; it < SEG * 8 / 256; ++it) {
    const int i = it * 256 + threadIdx.x;
    const int px = i >> 3, ch = i & 7;
    if (w0 + px >= W) break;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = ch * 8 + e;
      const int c = j / 9, rs = j - c * 9, r = rs / 3, s = rs - r * 3;
      v[e] = (j < T) ? xs[(j < T) ? c * 3 + r : 0][px + s] : 0.f;
    }
    dst[px * col_pitch16 + ch] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
  }
}

// OIHW fp32 [K][Cin][3][3] -> bf16 [K][64], columns >= 9*Cin zero
__global__ void prep_first_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w1, int K, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * 64) return;
  const int k = i >> 6, j = i & 63;
  w1[i] = __float2bfloat16_rn(j < T ? w[k * T + j] : 0.f);
}

}  // namespace

extern "C" {

int b200unet_first_im2col(const float* x_nchw, void* col, int col_cs, int N, int H, int W, int Cin, b200_stream_t stream) {
  B2_REQUIRE(Cin >= 1 && Cin <= 7, "first_im2col: Cin=%d must be in [1,7] (9*Cin <= 64 columns)", Cin);
  B2_REQUIRE(col_cs >= 64 && col_cs % 8 == 0, "first_im2col: col pitch %d must be >= 64 and a multiple of 8", col_cs);
  B2_REQUIRE(N > 0 && H > 0 && W > 0, "first_im2col: empty tensor");
  const int segs = (W + SEG - 1) / SEG;
  const long long blocks = static_cast<long long>(N) * H * segs;
  B2_REQUIRE(blocks < (1ll << 31), "first_im2col: tensor too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint4* c = static_cast<uint4*>(col);
  const long long p16 = col_cs / 8;
#define B2_IM2COL(CI) \
  case CI: first_im2col_kernel<CI><<<static_cast<int>(blocks), 256, 0, st>>>(x_nchw, c, p16, N, H, W, segs); break;
  switch (Cin) {
    B2_IM2COL(1) B2_IM2COL(2) B2_IM2COL(3) B2_IM2COL(4) B2_IM2COL(5) B2_IM2COL(6) B2_IM2COL(7)
  }
#undef B2_IM2COL
  return b2h::check_launch("first_im2col");
}

int b200unet_prep_first_weight(const float* w_oihw, void* w1, int Cout, int Cin, b200_stream_t stream) {
  B2_REQUIRE(Cin >= 1 && Cin <= 7 && Cout > 0, "prep_first_weight: bad shape Cout=%d Cin=%d", Cout, Cin);
  prep_first_kernel<<<(Cout * 64 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, static_cast<__nv_bfloat16*>(w1), Cout, Cin * 9);
  return b2h::check_launch("prep_first_weight");
}

int b200unet_conv3x3_first_igemm(const float* x_nchw, const void* w1, void* y, int y_cs, float* stats_partial, int N, int H,
                                 int W, int Cin, int Cout, b200_stream_t stream) {
  B2_REQUIRE(Cin >= 1 && Cin <= 7 && Cout % 64 == 0 && Cout > 0, "conv3x3_first_igemm: Cin=%d must be in [1,7], Cout=%d a multiple of 64", Cin, Cout);
  B2_REQUIRE(x_nchw && w1 && y && y_cs >= Cout && y_cs % 8 == 0 && N > 0 && H > 0 && W > 0, "conv3x3_first_igemm: bad arguments");
  return b2h::conv3x3_first_launch(x_nchw, w1, y, y_cs, stats_partial, N, H, W, Cin, Cout, static_cast<cudaStream_t>(stream));
}

int b200unet_conv3x3_first_bn_relu_igemm(const float* x_nchw, const void* w1, const float* scale, const float* shift, void* a,
                                         int a_cs, int N, int H, int W, int Cin, int Cout, b200_stream_t stream) {
  B2_REQUIRE(Cin >= 1 && Cin <= 7 && Cout % 64 == 0 && Cout > 0, "conv3x3_first_bn_relu_igemm: Cin=%d must be in [1,7], Cout=%d a multiple of 64", Cin, Cout);
  B2_REQUIRE(x_nchw && w1 && a && a_cs >= Cout && a_cs % 8 == 0 && N > 0 && H > 0 && W > 0, "conv3x3_first_bn_relu_igemm: bad arguments");
  B2_REQUIRE(scale != nullptr && shift != nullptr && reinterpret_cast<uintptr_t>(scale) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(shift) % 16 == 0,
             "conv3x3_first_bn_relu_igemm: scale and shift are required, 16-byte aligned");
  return b2h::conv3x3_first_launch(x_nchw, w1, a, a_cs, nullptr, N, H, W, Cin, Cout, static_cast<cudaStream_t>(stream), scale, shift);
}

int b200unet_conv1x1_c64_stat_rows(int N, int H, int W, int Cout) { return b2h::conv1x1_c64_stat_rows(N, H, W, Cout); }

int b200unet_conv1x1_c64_igemm(const void* x, int x_cs, const void* w, void* y, int y_cs, float* stats_partial, int N,
                               int H, int W, int Cout, b200_stream_t stream) {
  B2_REQUIRE(Cout % 64 == 0 && Cout > 0, "conv1x1_c64_igemm: Cout=%d must be a multiple of 64", Cout);
  B2_REQUIRE(x_cs >= 64 && y_cs >= Cout && x_cs % 8 == 0 && y_cs % 8 == 0, "conv1x1_c64_igemm: bad pitches %d %d", x_cs, y_cs);
  B2_REQUIRE(N > 0 && H > 0 && W > 0, "conv1x1_c64_igemm: empty tensor");
  return b2h::conv1x1_c64_launch(x, x_cs, w, y, y_cs, stats_partial, N, H, W, Cout, static_cast<cudaStream_t>(stream));
}

int b200unet_conv1x1_c64_bn_relu_igemm(const void* x, int x_cs, const void* w, const float* scale, const float* shift, void* a,
                                       int a_cs, int N, int H, int W, int Cout, b200_stream_t stream) {
  B2_REQUIRE(Cout % 64 == 0 && Cout > 0, "conv1x1_c64_bn_relu_igemm: Cout=%d must be a multiple of 64", Cout);
  B2_REQUIRE(x_cs >= 64 && a_cs >= Cout && x_cs % 8 == 0 && a_cs % 8 == 0, "conv1x1_c64_bn_relu_igemm: bad pitches %d %d", x_cs, a_cs);
  B2_REQUIRE(N > 0 && H > 0 && W > 0, "conv1x1_c64_bn_relu_igemm: empty tensor");
  B2_REQUIRE(scale != nullptr && shift != nullptr && reinterpret_cast<uintptr_t>(scale) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(shift) % 16 == 0,
             "conv1x1_c64_bn_relu_igemm: scale and shift are required, 16-byte aligned");
  return b2h::conv1x1_c64_launch(x, x_cs, w, a, a_cs, nullptr, N, H, W, Cout, static_cast<cudaStream_t>(stream), scale, shift);
}

}  // extern "C"
