// CUDA-core kernels for the two layers whose channel counts are too small for the tensor-core path:
//   inc.conv1  nn.Conv2d(n_channels<=4, 64, 3, padding=1, bias=False)   (reference Model.py:15-16 via :111)
//   outc       nn.Conv2d(64, n_classes, 1) + bias                        (reference Model.py:86-92)
// Both are bound by the 64-channel bf16 activation they write / read (128 B per pixel), not by arithmetic.
#include "../../include/b200unet.h"
#include "host_common.h"
#include "head_common.cuh"

#include <cuda_bf16.h>

namespace {

constexpr int MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// ---------------------------------------------------------------------------------------- inc.conv1 fprop
// Block = 128 consecutive pixels (linear over N*H*W) x one 64-channel group; thread = one pixel.
// The bf16 tile is staged in smem (16-byte chunks XOR-swizzled by row) for a coalesced store and for the
// per-channel sum / sum of squares that feed BatchNorm.
template <int CIN>
__global__ void __launch_bounds__(128) first_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          __nv_bfloat16* __restrict__ y, int y_cs,
                                                          float* __restrict__ stats, int N, int H, int W, int Cout) {
  __shared__ float ws[CIN * 9][64];          // [c*9+rs][k]
  __shared__ __align__(16) uint8_t tile[128 * 128];
  __shared__ float red[2][2][64];
  const int kg = blockIdx.y;                 // 64-channel group
  for (int i = threadIdx.x; i < CIN * 9 * 64; i += 128) {
    const int k = i & 63, j = i >> 6;
    ws[j][k] = w[(static_cast<size_t>(kg * 64 + k) * CIN) * 9 + j];
  }
  __syncthreads();
  const long long P = static_cast<long long>(N) * H * W;
  const long long p = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  const bool valid = p < P;
  float acc[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) acc[k] = 0.f;
  if (valid) {
    const int wq = static_cast<int>(p % W);
    const int hq = static_cast<int>((p / W) % H);
    const long long n = p / (static_cast<long long>(W) * H);
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float* xc = x + (n * CIN + c) * static_cast<long long>(H) * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = hq + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ww = wq + s - 1;
          float xv = 0.f;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) xv = __ldg(xc + static_cast<long long>(hh) * W + ww);
          const float4* wr = reinterpret_cast<const float4*>(ws[c * 9 + r * 3 + s]);
#pragma unroll
          for (int k4 = 0; k4 < 16; ++k4) {
            const float4 wv = wr[k4];
            acc[4 * k4 + 0] = fmaf(xv, wv.x, acc[4 * k4 + 0]);
            acc[4 * k4 + 1] = fmaf(xv, wv.y, acc[4 * k4 + 1]);
            acc[4 * k4 + 2] = fmaf(xv, wv.z, acc[4 * k4 + 2]);
            acc[4 * k4 + 3] = fmaf(xv, wv.w, acc[4 * k4 + 3]);
          }
        }
      }
    }
  }
  const int r = threadIdx.x;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    uint4 pk = make_uint4(pack2(acc[8 * t], acc[8 * t + 1]), pack2(acc[8 * t + 2], acc[8 * t + 3]),
                          pack2(acc[8 * t + 4], acc[8 * t + 5]), pack2(acc[8 * t + 6], acc[8 * t + 7]));
    *reinterpret_cast<uint4*>(tile + r * 128 + ((t ^ (r & 7)) << 4)) = pk;
  }
  __syncthreads();
  // coalesced copy-out: 8 threads per pixel row
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int rr = i >> 3, ch = i & 7;
    const long long pp = static_cast<long long>(blockIdx.x) * 128 + rr;
    if (pp < P)
      *reinterpret_cast<uint4*>(y + pp * y_cs + kg * 64 + ch * 8) =
          *reinterpret_cast<const uint4*>(tile + rr * 128 + ((ch ^ (rr & 7)) << 4));
  }
  if (stats != nullptr) {
    const int c = threadIdx.x & 63, hf = threadIdx.x >> 6;
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < 64; ++i) {
      const int rr = hf * 64 + i;
      const long long pp = static_cast<long long>(blockIdx.x) * 128 + rr;
      const uint16_t u = *reinterpret_cast<const uint16_t*>(tile + rr * 128 + (((c >> 3) ^ (rr & 7)) << 4) + (c & 7) * 2);
      const float v = (pp < P) ? __uint_as_float(static_cast<uint32_t>(u) << 16) : 0.f;
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
    red[hf][0][c] = s1;
    red[hf][1][c] = s2;
    __syncthreads();
    if (threadIdx.x < 64) {
      float* dst = stats + static_cast<size_t>(blockIdx.x) * 2 * Cout + kg * 64;
      dst[threadIdx.x] = red[0][0][threadIdx.x] + red[1][0][threadIdx.x];
      dst[Cout + threadIdx.x] = red[0][1][threadIdx.x] + red[1][1][threadIdx.x];
    }
  }
}

// ---------------------------------------------------------------------------------------- inc.conv1 wgrad
// dW[k][c][r][s] = sum_p dy[p][k] * x[p + (r-1, s-1)][c]: a [64 x pixels] x [pixels x CIN*9] product, FMA-bound.
// Per 64-pixel tile the block stages dy (fp32) and the im2col row of every pixel in smem; thread
// (stream q, 4-channel group kq, tap group jq) keeps a 4 x JG register tile and walks its stream's 16 pixels with
// 3 LDS.128 per 28 FMAs. Streams are reduced through smem at the end; per-block partials go to the workspace.
template <int CIN>
__global__ void __launch_bounds__(256) first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                          int dy_cs, float* __restrict__ partial, int N, int H, int W,
                                                          int Cout) {
  constexpr int T = CIN * 9;
  constexpr int JG = (T + 3) / 4;          // taps per thread
  constexpr int JGP = (JG + 3) / 4 * 4;    // padded to a multiple of 4 floats
  constexpr int TP = 64;                   // pixels per tile
  __shared__ __align__(16) float dys[TP][64];
  __shared__ __align__(16) float xcol[TP][4 * JGP];
  const int kg = blockIdx.y;
  const int q = threadIdx.x >> 6, kq = (threadIdx.x & 63) >> 2, jq = threadIdx.x & 3;
  float acc[4][JG];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < JG; ++j) acc[i][j] = 0.f;
  const long long P = static_cast<long long>(N) * H * W;
  const long long ntiles = (P + TP - 1) / TP;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long p0 = tile * TP;
    __syncthreads();  // previous tile fully consumed
    // stage dy: 64 px x 64 ch bf16 = 512 uint4, two per thread
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int v = threadIdx.x + it * 256;
      const int pr = v >> 3, ch = v & 7;
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (p0 + pr < P) unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (p0 + pr) * dy_cs + kg * 64 + ch * 8)), f);
      *reinterpret_cast<float4*>(&dys[pr][ch * 8]) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(&dys[pr][ch * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
    }
    // stage the im2col rows: element (pixel pr, tap j) with j = c*9 + r*3 + s stored at [pr][(j / JG) * JGP + j % JG];
    // thread (pr = t/4, grp = t%4) fills the JG taps of its group: one coordinate decode, JG predicated loads
    {
      const int pr = threadIdx.x >> 2, grp = threadIdx.x & 3;
      const long long p = p0 + pr;
      const int wq = static_cast<int>(p % W);
      const int hq = static_cast<int>((p / W) % H);
      const long long n = p / (static_cast<long long>(W) * H);
      const float* xn = x + n * CIN * static_cast<long long>(H) * W;
#pragma unroll
      for (int jj = 0; jj < JGP; ++jj) {
        const int j = grp * JG + jj;
        float xv = 0.f;
        if (jj < JG && j < T && p < P) {
          const int c = j / 9, r = (j % 9) / 3, sx = j % 3;
          const int hh = hq + r - 1, ww = wq + sx - 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) xv = __ldg(xn + (static_cast<long long>(c) * H + hh) * W + ww);
        }
        xcol[pr][grp * JGP + jj] = xv;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < TP / 4; ++i) {
      const int pr = q * (TP / 4) + i;
      const float4 d = *reinterpret_cast<const float4*>(&dys[pr][kq * 4]);
      float xv[JGP];
#pragma unroll
      for (int j4 = 0; j4 < JGP / 4; ++j4) {
        const float4 t = *reinterpret_cast<const float4*>(&xcol[pr][jq * JGP + j4 * 4]);
        xv[4 * j4] = t.x; xv[4 * j4 + 1] = t.y; xv[4 * j4 + 2] = t.z; xv[4 * j4 + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < JG; ++j) {
        acc[0][j] = fmaf(d.x, xv[j], acc[0][j]);
        acc[1][j] = fmaf(d.y, xv[j], acc[1][j]);
        acc[2][j] = fmaf(d.z, xv[j], acc[2][j]);
        acc[3][j] = fmaf(d.w, xv[j], acc[3][j]);
      }
    }
  }
  // reduce the 4 pixel streams: reuse dys as [4 streams][64 k][T] would not fit; go tap by tap through xcol-sized scratch
  __syncthreads();
  float* scratch = &dys[0][0];  // 4096 floats >= 4 streams * 64 k * 4 (one tap slot per jq) * ... processed per j
#pragma unroll
  for (int j = 0; j < JG; ++j) {
    // slot layout: [q][k = kq*4+i][jq]
#pragma unroll
    for (int i = 0; i < 4; ++i) scratch[(q * 64 + kq * 4 + i) * 4 + jq] = acc[i][j];
    __syncthreads();
    {
      const int k = threadIdx.x >> 2, g4 = threadIdx.x & 3;  // 64 k x 4 tap groups = 256 outputs per j
      const int jt = g4 * JG + j;
      if (jt < T) {
        const float t = scratch[(0 * 64 + k) * 4 + g4] + scratch[(1 * 64 + k) * 4 + g4] + scratch[(2 * 64 + k) * 4 + g4] +
                        scratch[(3 * 64 + k) * 4 + g4];
        partial[(static_cast<size_t>(blockIdx.x) * Cout + kg * 64 + k) * T + jt] = t;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------- head
// OutConv forward -> fp32 NCHW logits; the dot products are warp-cooperative (head_common.cuh): coalesced 512-byte loads.
__global__ void __launch_bounds__(256) head_fprop_kernel(const __nv_bfloat16* __restrict__ a, int a_cs,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ z, long long P, long long HW, int Cin,
                                                         int ncls) {
  extern __shared__ __align__(16) float wsm[];
  b2head::load_weights(wsm, w, Cin, ncls);
  uint32_t* stage = b2head::warp_stage(wsm, Cin, ncls);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long g = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); g * 32 < P; g += warps) {
    float acc[b2head::MAXC];
    b2head::logits_warp32(a, a_cs, wsm, stage, bias, g * 32, P, Cin, ncls, acc);
    const long long p = g * 32 + lane;
    if (p < P) {
      const long long n = p / HW, hw = p % HW;
#pragma unroll
      for (int j = 0; j < b2head::MAXC; ++j)
        if (j < ncls) z[(n * ncls + j) * HW + hw] = acc[j];
    }
  }
}

// Block = 256 threads walks tiles of 256 pixels: dz of the tile is staged (coalesced) in smem, then thread
// (pixel lane pl, 8-channel group cg) handles the pixels pl, pl+ppb, ... of the tile with all its activation loads
// issued up front. dW/db partial sums stay in registers across tiles and are reduced once per block.
template <int NCLS>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dz, const __nv_bfloat16* __restrict__ a,
                                                       int a_cs, const float* __restrict__ w,
                                                       __nv_bfloat16* __restrict__ da, int da_cs,
                                                       float* __restrict__ partial, long long P, long long HW, int Cin) {
  constexpr int ncls = NCLS;
  __shared__ float dzs[NCLS][256];
  __shared__ float sh[256][9];
  const int cgs = Cin >> 3;
  const int cg = threadIdx.x % cgs, pl = threadIdx.x / cgs, ppb = 256 / cgs;
  float wr[NCLS][8], dwacc[NCLS][8], dbacc[NCLS];
#pragma unroll
  for (int j = 0; j < NCLS; ++j) {
    dbacc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      wr[j][i] = w[j * Cin + cg * 8 + i];
      dwacc[j][i] = 0.f;
    }
  }
  const long long ntiles = (P + 255) / 256;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long p0 = tile * 256;
    __syncthreads();
    {
      const long long p = p0 + threadIdx.x;
      const long long n = p / HW, hw = p - n * HW;
#pragma unroll
      for (int j = 0; j < NCLS; ++j) dzs[j][threadIdx.x] = (p < P) ? __ldg(dz + (n * ncls + j) * HW + hw) : 0.f;
    }
    __syncthreads();
    // cgs <= 32 -> ppb >= 8 pixels per pass, 256/ppb = cgs passes; process 4 passes per batch of loads
    for (int it0 = 0; it0 < cgs; it0 += 4) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long p = p0 + (it0 + u) * ppb + pl;
        raw[u] = (it0 + u < cgs && p < P) ? __ldg(reinterpret_cast<const uint4*>(a + p * a_cs + cg * 8)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int lp = (it0 + u) * ppb + pl;
        const long long p = p0 + lp;
        if (it0 + u < cgs && p < P) {
          float f[8], o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          unpack8(raw[u], f);
#pragma unroll
          for (int j = 0; j < NCLS; ++j) {
            const float g = dzs[j][lp];
            dbacc[j] += g;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i] = fmaf(g, wr[j][i], o[i]);
              dwacc[j][i] = fmaf(g, f[i], dwacc[j][i]);
            }
          }
          *reinterpret_cast<uint4*>(da + p * da_cs + cg * 8) =
              make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
        }
      }
    }
  }
  // partial layout: [block][ncls][Cin + 1], last column = bias gradient
#pragma unroll
  for (int j = 0; j < NCLS; ++j) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) sh[threadIdx.x][i] = dwacc[j][i];
    sh[threadIdx.x][8] = (cg == 0) ? dbacc[j] : 0.f;
    __syncthreads();
    float* dst = partial + (static_cast<size_t>(blockIdx.x) * ncls + j) * (Cin + 1);
    for (int c = threadIdx.x; c <= Cin; c += blockDim.x) {
      float t = 0.f;
      if (c < Cin) {
        const int g = c >> 3, i = c & 7;
        for (int l = 0; l < ppb; ++l) t += sh[l * cgs + g][i];
      } else {
        for (int l = 0; l < 256; ++l) t += sh[l][8];
      }
      dst[c] = t;
    }
  }
}

// column sums of partial [rows][ncols] (fp64 accumulate), optional split into (dw [ncls][Cin], db [ncls])
__global__ void colsum_kernel(const float* __restrict__ partial, int rows, int ncols, float* __restrict__ out,
                              float* __restrict__ out2, int split_stride) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double acc = 0.0;
  for (int r = 0; r < rows; ++r) acc += static_cast<double>(partial[static_cast<size_t>(r) * ncols + col]);
  if (split_stride == 0) {
    out[col] = static_cast<float>(acc);
  } else {
    const int j = col / split_stride, c = col % split_stride;
    if (c < split_stride - 1) out[j * (split_stride - 1) + c] = static_cast<float>(acc);
    else out2[j] = static_cast<float>(acc);
  }
}

int first_wgrad_blocks(long long P) {
  long long b = (P + 63) / 64;
  if (b > 148 * 4) b = 148 * 4;
  return static_cast<int>(b < 1 ? 1 : b);
}
int head_bwd_blocks(long long P, int Cin) {
  (void)Cin;
  long long b = (P + 255) / 256;
  if (b > 148 * 4) b = 148 * 4;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int b200unet_conv3x3_first_fprop(const float* x_nchw, const float* w_oihw, void* y, int y_cs, float* stats_partial,
                                 int N, int H, int W, int Cin, int Cout, b200_stream_t stream) {
  B2_REQUIRE(Cin >= 1 && Cin <= 4, "conv3x3_first_fprop: Cin=%d must be in [1,4]", Cin);
  B2_REQUIRE(Cout % 64 == 0 && y_cs % 8 == 0, "conv3x3_first_fprop: Cout=%d must be a multiple of 64", Cout);
  const long long P = static_cast<long long>(N) * H * W;
  dim3 grid(static_cast<unsigned>((P + 127) / 128), Cout / 64);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* yy = static_cast<__nv_bfloat16*>(y);
  switch (Cin) {
    case 1: first_fprop_kernel<1><<<grid, 128, 0, st>>>(x_nchw, w_oihw, yy, y_cs, stats_partial, N, H, W, Cout); break;
    case 2: first_fprop_kernel<2><<<grid, 128, 0, st>>>(x_nchw, w_oihw, yy, y_cs, stats_partial, N, H, W, Cout); break;
    case 3: first_fprop_kernel<3><<<grid, 128, 0, st>>>(x_nchw, w_oihw, yy, y_cs, stats_partial, N, H, W, Cout); break;
    default: first_fprop_kernel<4><<<grid, 128, 0, st>>>(x_nchw, w_oihw, yy, y_cs, stats_partial, N, H, W, Cout); break;
  }
  return b2h::check_launch("conv3x3_first_fprop");
}

int64_t b200unet_conv3x3_first_wgrad_workspace_floats(int N, int H, int W, int Cin, int Cout) {
  return static_cast<int64_t>(first_wgrad_blocks(static_cast<long long>(N) * H * W)) * Cout * Cin * 9;
}

int b200unet_conv3x3_first_wgrad(const float* x_nchw, const void* dy, int dy_cs, float* partial, float* dw_oihw, int N,
                                 int H, int W, int Cin, int Cout, b200_stream_t stream) {
  B2_REQUIRE(Cin >= 1 && Cin <= 4, "conv3x3_first_wgrad: Cin=%d must be in [1,4]", Cin);
  B2_REQUIRE(Cout % 64 == 0 && dy_cs % 8 == 0, "conv3x3_first_wgrad: Cout=%d must be a multiple of 64", Cout);
  const long long P = static_cast<long long>(N) * H * W;
  const int blocks = first_wgrad_blocks(P);
  dim3 grid(blocks, Cout / 64);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto* dd = static_cast<const __nv_bfloat16*>(dy);
  switch (Cin) {
    case 1: first_wgrad_kernel<1><<<grid, 256, 0, st>>>(x_nchw, dd, dy_cs, partial, N, H, W, Cout); break;
    case 2: first_wgrad_kernel<2><<<grid, 256, 0, st>>>(x_nchw, dd, dy_cs, partial, N, H, W, Cout); break;
    case 3: first_wgrad_kernel<3><<<grid, 256, 0, st>>>(x_nchw, dd, dy_cs, partial, N, H, W, Cout); break;
    default: first_wgrad_kernel<4><<<grid, 256, 0, st>>>(x_nchw, dd, dy_cs, partial, N, H, W, Cout); break;
  }
  if (int e = b2h::check_launch("conv3x3_first_wgrad")) return e;
  const int ncols = Cout * Cin * 9;
  colsum_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(partial, blocks, ncols, dw_oihw, nullptr, 0);
  return b2h::check_launch("conv3x3_first_wgrad_reduce");
}

int b200unet_head_fprop(const void* a, int a_cs, const float* w, const float* bias, float* logits_nchw, int N, int H,
                        int W, int Cin, int ncls, b200_stream_t stream) {
  B2_REQUIRE(ncls >= 1 && ncls <= 8, "head_fprop: n_classes=%d must be in [1,8]", ncls);
  B2_REQUIRE(Cin > 0 && Cin % 64 == 0 && a_cs % 8 == 0 && a_cs >= Cin, "head_fprop: Cin=%d must be a multiple of 64 (pitch %d)", Cin, a_cs);
  B2_REQUIRE(b2head::smem_bytes(Cin, ncls, 8) <= 48 * 1024, "head_fprop: Cin=%d x n_classes=%d weights do not fit shared memory", Cin, ncls);
  const long long P = static_cast<long long>(N) * H * W;
  long long blocks = (P + 255) / 256;
  if (blocks > MAX_BLOCKS) blocks = MAX_BLOCKS;
  head_fprop_kernel<<<static_cast<int>(blocks), 256, b2head::smem_bytes(Cin, ncls, 8), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), a_cs, w, bias, logits_nchw, P, static_cast<long long>(H) * W, Cin, ncls);
  return b2h::check_launch("head_fprop");
}

int64_t b200unet_head_bwd_workspace_floats(int N, int H, int W, int Cin, int ncls) {
  return static_cast<int64_t>(head_bwd_blocks(static_cast<long long>(N) * H * W, Cin)) * ncls * (Cin + 1);
}

int b200unet_head_bwd(const float* dz_nchw, const void* a, int a_cs, const float* w, void* da, int da_cs, float* partial,
                      float* dw, float* db, int N, int H, int W, int Cin, int ncls, b200_stream_t stream) {
  B2_REQUIRE(ncls >= 1 && ncls <= 8, "head_bwd: n_classes=%d must be in [1,8]", ncls);
  B2_REQUIRE(Cin % 8 == 0 && 256 % (Cin / 8) == 0 && a_cs % 8 == 0 && da_cs % 8 == 0, "head_bwd: unsupported Cin=%d", Cin);
  const long long P = static_cast<long long>(N) * H * W;
  const int blocks = head_bwd_blocks(P, Cin);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2_HEAD_BWD(NC)                                                                                          \
  case NC:                                                                                                      \
    head_bwd_kernel<NC><<<blocks, 256, 0, st>>>(dz_nchw, static_cast<const __nv_bfloat16*>(a), a_cs, w,          \
                                                static_cast<__nv_bfloat16*>(da), da_cs, partial, P,             \
                                                static_cast<long long>(H) * W, Cin);                            \
    break;
  switch (ncls) {
    B2_HEAD_BWD(1) B2_HEAD_BWD(2) B2_HEAD_BWD(3) B2_HEAD_BWD(4) B2_HEAD_BWD(5) B2_HEAD_BWD(6) B2_HEAD_BWD(7) B2_HEAD_BWD(8)
  }
#undef B2_HEAD_BWD
  if (int e = b2h::check_launch("head_bwd")) return e;
  const int ncols = ncls * (Cin + 1);
  colsum_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(partial, blocks, ncols, dw, db, Cin + 1);
  return b2h::check_launch("head_bwd_reduce");
}

}  // extern "C"
