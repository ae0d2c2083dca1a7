// Generic fp32 path of the U-Net hot path: every operator of Model.py:7-92 as a plain CUDA-core kernel on NCHW fp32
// tensors (the reference's own layout and precision). It is the CHECK MODE of the library (north_star: logits and loss
// within 1e-4 of the reference in fp32) and the documented slow-but-correct route for shapes the tensor-core path
// does not take: H or W not divisible by 16 (F.pad branch, Model.py:69-73; floor-mode pooling), channel widths that
// are not multiples of 64, and the dropout variants (Model.py:34-39, 81-82). Not tuned: one thread per output
// element, fp32 FMA accumulation, fp64 for every reduction over pixels.
//
// Tensor convention: pointer to element (n=0, c=0, h=0, w=0) of an NCHW view, `*_ns` = batch stride in elements
// (channel stride is H*W, row stride W), so a channel range of a wider tensor (the concat buffer) is a valid view.
#include "../../include/b200unet.h"
#include "host_common.h"

namespace {

constexpr int GT = 256;
inline int gblocks(long long n) {
  long long b = (n + GT - 1) / GT;
  if (b > 148 * 32) b = 148 * 32;
  return static_cast<int>(b < 1 ? 1 : b);
}
#define GEN_LOOP(i, total)                                                                    \
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < (total); \
       i += static_cast<long long>(gridDim.x) * blockDim.x)

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += sh[i];
  return t;  // valid in thread 0
}

// out[n,o,h,w] = sum_{i,r,s} in[n,i,h+r-1,w+s-1] * Wt(o,i,r,s);  transposed = 0: Wt = w[o][i][r][s] (fprop),
// transposed = 1: Wt = w[i][o][2-r][2-s] (backward-data of the same nn.Conv2d, Model.py:15-16)
__global__ void gen_conv3x3_kernel(const float* __restrict__ in, long long in_ns, const float* __restrict__ w,
                                   float* __restrict__ out, long long out_ns, int N, int Ci, int Co, int H, int W,
                                   int transposed) {
  const long long total = static_cast<long long>(N) * Co * H * W;
  GEN_LOOP(idx, total) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int o = static_cast<int>((idx / (static_cast<long long>(W) * H)) % Co);
    const long long n = idx / (static_cast<long long>(W) * H * Co);
    float acc = 0.f;
    for (int i = 0; i < Ci; ++i) {
      const float* ip = in + n * in_ns + static_cast<long long>(i) * H * W;
      const float* wp = transposed ? (w + (static_cast<long long>(i) * Co + o) * 9) : (w + (static_cast<long long>(o) * Ci + i) * 9);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int yy = y + r - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int xx = x + s - 1;
          if (xx < 0 || xx >= W) continue;
          const float wv = transposed ? wp[(2 - r) * 3 + (2 - s)] : wp[r * 3 + s];
          acc = fmaf(ip[static_cast<long long>(yy) * W + xx], wv, acc);
        }
      }
    }
    out[n * out_ns + (static_cast<long long>(o) * H + y) * W + x] = acc;
  }
}

// dw[k][c][r][s] = sum_{n,h,w} dy[n,k,h,w] * x[n,c,h+r-1,w+s-1]; one block per (c, k)
__global__ void gen_conv3x3_wgrad_kernel(const float* __restrict__ x, long long x_ns, const float* __restrict__ dy,
                                         long long dy_ns, float* __restrict__ dw, int N, int Ci, int Co, int H, int W) {
  __shared__ double sh[8];
  const int c = blockIdx.x, k = blockIdx.y;
  double acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.0;
  const long long P = static_cast<long long>(N) * H * W;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const int xw = static_cast<int>(p % W);
    const int yh = static_cast<int>((p / W) % H);
    const long long n = p / (static_cast<long long>(W) * H);
    const float g = dy[n * dy_ns + (static_cast<long long>(k) * H + yh) * W + xw];
    const float* xp = x + n * x_ns + static_cast<long long>(c) * H * W;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = yh + r - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int xx = xw + s - 1;
        if (xx < 0 || xx >= W) continue;
        acc[r * 3 + s] += static_cast<double>(g) * static_cast<double>(xp[static_cast<long long>(yy) * W + xx]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const double s = block_sum(acc[t], sh);
    if (threadIdx.x == 0) dw[(static_cast<long long>(k) * Ci + c) * 9 + t] = static_cast<float>(s);
  }
}

// sums[c] = sum y, sums[C + c] = sum y^2 over (n,h,w); one block per channel (BatchNorm2d, Model.py:17,21)
__global__ void gen_channel_stats_kernel(const float* __restrict__ y, long long y_ns, double* __restrict__ sums, int N,
                                         int C, int HW) {
  __shared__ double sh[8];
  const int c = blockIdx.x;
  double s1 = 0.0, s2 = 0.0;
  const long long P = static_cast<long long>(N) * HW;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const long long n = p / HW, q = p - n * HW;
    const double v = y[n * y_ns + static_cast<long long>(c) * HW + q];
    s1 += v;
    s2 += v * v;
  }
  const double t1 = block_sum(s1, sh);
  const double t2 = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    sums[c] = t1;
    sums[C + c] = t2;
  }
}

__global__ void gen_bn_relu_fwd_kernel(const float* __restrict__ y, long long y_ns, const float* __restrict__ scale,
                                       const float* __restrict__ shift, float* __restrict__ a, long long a_ns, int N, int C,
                                       int HW) {
  const long long total = static_cast<long long>(N) * C * HW;
  GEN_LOOP(idx, total) {
    const long long q = idx % HW;
    const int c = static_cast<int>((idx / HW) % C);
    const long long n = idx / (static_cast<long long>(HW) * C);
    const float v = fmaf(scale[c], y[n * y_ns + static_cast<long long>(c) * HW + q], shift[c]);
    a[n * a_ns + static_cast<long long>(c) * HW + q] = fmaxf(v, 0.f);
  }
}

// nn.MaxPool2d(2) (Model.py:36,42): floor mode, first maximum in row-major window order, NaN wins
__global__ void gen_maxpool_kernel(const float* __restrict__ a, long long a_ns, float* __restrict__ p, uint8_t* __restrict__ idx,
                                   int N, int C, int H, int W) {
  const int Hp = H / 2, Wp = W / 2;
  const long long total = static_cast<long long>(N) * C * Hp * Wp;
  GEN_LOOP(i, total) {
    const int wp = static_cast<int>(i % Wp);
    const int hp = static_cast<int>((i / Wp) % Hp);
    const int c = static_cast<int>((i / (static_cast<long long>(Wp) * Hp)) % C);
    const long long n = i / (static_cast<long long>(Wp) * Hp * C);
    const float* src = a + n * a_ns + (static_cast<long long>(c) * H + 2 * hp) * W + 2 * wp;
    float best = src[0];
    int bi = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = src[(k >> 1) * W + (k & 1)];
      if (v > best || v != v) {  // torch's max_pool2d: `val > maxval || isnan(val)`
        best = v;
        bi = k;
      }
    }
    p[i] = best;
    if (idx != nullptr) idx[i] = static_cast<uint8_t>(bi);
  }
}

// g[n,c,2hp+di,2wp+dj] += gp[n,c,hp,wp] at the recorded window position (MaxPool2d backward into the skip gradient)
__global__ void gen_unpool_add_kernel(const float* __restrict__ gp, const uint8_t* __restrict__ idx, float* __restrict__ g,
                                      long long g_ns, int N, int C, int H, int W) {
  const int Hp = H / 2, Wp = W / 2;
  const long long total = static_cast<long long>(N) * C * Hp * Wp;
  GEN_LOOP(i, total) {
    const int wp = static_cast<int>(i % Wp);
    const int hp = static_cast<int>((i / Wp) % Hp);
    const int c = static_cast<int>((i / (static_cast<long long>(Wp) * Hp)) % C);
    const long long n = i / (static_cast<long long>(Wp) * Hp * C);
    const int k = idx[i];
    g[n * g_ns + (static_cast<long long>(c) * H + 2 * hp + (k >> 1)) * W + 2 * wp + (k & 1)] += gp[i];
  }
}

// sums[c] = sum da, sums[C+c] = sum da*xhat with da = g * [bn(y) > 0]; one block per channel
__global__ void gen_bn_bwd_reduce_kernel(const float* __restrict__ g, long long g_ns, const float* __restrict__ y,
                                         long long y_ns, const float* __restrict__ scale, const float* __restrict__ shift,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         double* __restrict__ sums, int N, int C, int HW, int relu) {
  __shared__ double sh[8];
  const int c = blockIdx.x;
  const float sc = scale[c], sf = shift[c];
  const double m = mean[c], r = rstd[c];
  double s1 = 0.0, s2 = 0.0;
  const long long P = static_cast<long long>(N) * HW;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const long long n = p / HW, q = p - n * HW;
    const float yv = y[n * y_ns + static_cast<long long>(c) * HW + q];
    if (!relu || fmaf(sc, yv, sf) > 0.f) {
      const double da = g[n * g_ns + static_cast<long long>(c) * HW + q];
      s1 += da;
      s2 += da * ((static_cast<double>(yv) - m) * r);
    }
  }
  const double t1 = block_sum(s1, sh);
  const double t2 = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    sums[c] = t1;
    sums[C + c] = t2;
  }
}

// dy = gamma*rstd*(da - sum_da/count - xhat*sum_dax/count); dgamma/dbeta from the LOCAL sums
__global__ void gen_bn_bwd_apply_kernel(const float* __restrict__ g, long long g_ns, const float* __restrict__ y,
                                        long long y_ns, const float* __restrict__ gamma, const float* __restrict__ scale,
                                        const float* __restrict__ shift, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const double* __restrict__ sums, double count,
                                        const double* __restrict__ sums_local, float* __restrict__ dy, long long dy_ns,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int C, int HW, int relu) {
  const long long total = static_cast<long long>(N) * C * HW;
  if (blockIdx.x == 0 && dgamma != nullptr) {
    const double* sl = sums_local != nullptr ? sums_local : sums;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = static_cast<float>(sl[c]);
      dgamma[c] = static_cast<float>(sl[C + c]);
    }
  }
  GEN_LOOP(idx, total) {
    const long long q = idx % HW;
    const int c = static_cast<int>((idx / HW) % C);
    const long long n = idx / (static_cast<long long>(HW) * C);
    const float yv = y[n * y_ns + static_cast<long long>(c) * HW + q];
    const double da = (!relu || fmaf(scale[c], yv, shift[c]) > 0.f) ? static_cast<double>(g[n * g_ns + static_cast<long long>(c) * HW + q]) : 0.0;
    const double r = rstd[c];
    const double xh = (static_cast<double>(yv) - static_cast<double>(mean[c])) * r;
    const double v = static_cast<double>(gamma[c]) * r * (da - sums[c] / count - xh * sums[C + c] / count);
    dy[n * dy_ns + static_cast<long long>(c) * HW + q] = static_cast<float>(v);
  }
}

// nn.ConvTranspose2d(Cin, Cup, 2, 2) + bias (Model.py:56-57,66) into a (H2 x W2) canvas at (pt, pl) (F.pad, :69-73)
__global__ void gen_convt_fprop_kernel(const float* __restrict__ x, long long x_ns, const float* __restrict__ w,
                                       const float* __restrict__ b, float* __restrict__ out, long long out_ns, int N, int Ci,
                                       int Cu, int h, int wd, int H2, int W2, int pt, int pl) {
  const long long total = static_cast<long long>(N) * Cu * (2 * h) * (2 * wd);
  GEN_LOOP(idx, total) {
    const int X = static_cast<int>(idx % (2 * wd));
    const int Y = static_cast<int>((idx / (2 * wd)) % (2 * h));
    const int d = static_cast<int>((idx / (static_cast<long long>(4) * wd * h)) % Cu);
    const long long n = idx / (static_cast<long long>(4) * wd * h * Cu);
    const int hh = Y >> 1, i = Y & 1, ww = X >> 1, j = X & 1;
    float acc = b != nullptr ? b[d] : 0.f;
    for (int c = 0; c < Ci; ++c)
      acc = fmaf(x[n * x_ns + (static_cast<long long>(c) * h + hh) * wd + ww], w[((static_cast<long long>(c) * Cu + d) * 2 + i) * 2 + j], acc);
    out[n * out_ns + (static_cast<long long>(d) * H2 + Y + pt) * W2 + X + pl] = acc;
  }
}

__global__ void gen_convt_dgrad_kernel(const float* __restrict__ du, long long du_ns, const float* __restrict__ w,
                                       float* __restrict__ dx, long long dx_ns, int N, int Ci, int Cu, int h, int wd, int H2,
                                       int W2, int pt, int pl) {
  const long long total = static_cast<long long>(N) * Ci * h * wd;
  GEN_LOOP(idx, total) {
    const int ww = static_cast<int>(idx % wd);
    const int hh = static_cast<int>((idx / wd) % h);
    const int c = static_cast<int>((idx / (static_cast<long long>(wd) * h)) % Ci);
    const long long n = idx / (static_cast<long long>(wd) * h * Ci);
    float acc = 0.f;
    for (int d = 0; d < Cu; ++d) {
      const float* up = du + n * du_ns + (static_cast<long long>(d) * H2 + 2 * hh + pt) * W2 + 2 * ww + pl;
      const float* wp = w + (static_cast<long long>(c) * Cu + d) * 4;
      acc = fmaf(up[0], wp[0], acc);
      acc = fmaf(up[1], wp[1], acc);
      acc = fmaf(up[W2], wp[2], acc);
      acc = fmaf(up[W2 + 1], wp[3], acc);
    }
    dx[n * dx_ns + (static_cast<long long>(c) * h + hh) * wd + ww] = acc;
  }
}

// dW[c][d][i][j] = sum x[n,c,h,w] du[n,d,2h+i,2w+j]; db[d] = sum du over the un-padded region. One block per (d, c).
__global__ void gen_convt_wgrad_kernel(const float* __restrict__ x, long long x_ns, const float* __restrict__ du,
                                       long long du_ns, float* __restrict__ dw, float* __restrict__ db, int N, int Ci, int Cu,
                                       int h, int wd, int H2, int W2, int pt, int pl) {
  __shared__ double sh[8];
  const int d = blockIdx.x, c = blockIdx.y;
  double acc[4] = {0.0, 0.0, 0.0, 0.0}, sb = 0.0;
  const long long P = static_cast<long long>(N) * h * wd;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const int ww = static_cast<int>(p % wd);
    const int hh = static_cast<int>((p / wd) % h);
    const long long n = p / (static_cast<long long>(wd) * h);
    const double xv = x[n * x_ns + (static_cast<long long>(c) * h + hh) * wd + ww];
    const float* up = du + n * du_ns + (static_cast<long long>(d) * H2 + 2 * hh + pt) * W2 + 2 * ww + pl;
    const double u0 = up[0], u1 = up[1], u2 = up[W2], u3 = up[W2 + 1];
    acc[0] += xv * u0;
    acc[1] += xv * u1;
    acc[2] += xv * u2;
    acc[3] += xv * u3;
    sb += u0 + u1 + u2 + u3;
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const double s = block_sum(acc[t], sh);
    if (threadIdx.x == 0) dw[(static_cast<long long>(c) * Cu + d) * 4 + t] = static_cast<float>(s);
  }
  if (c == 0 && db != nullptr) {
    const double s = block_sum(sb, sh);
    if (threadIdx.x == 0) db[d] = static_cast<float>(s);
  }
}

// OutConv (Model.py:86-92): z[n,j,p] = b[j] + sum_c a[n,c,p] w[j,c]
__global__ void gen_conv1x1_fwd_kernel(const float* __restrict__ a, long long a_ns, const float* __restrict__ w,
                                       const float* __restrict__ b, float* __restrict__ z, int N, int C, int J, long long HW) {
  const long long total = static_cast<long long>(N) * J * HW;
  GEN_LOOP(idx, total) {
    const long long q = idx % HW;
    const int j = static_cast<int>((idx / HW) % J);
    const long long n = idx / (HW * J);
    float acc = b != nullptr ? b[j] : 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(a[n * a_ns + static_cast<long long>(c) * HW + q], w[static_cast<long long>(j) * C + c], acc);
    z[idx] = acc;
  }
}
__global__ void gen_conv1x1_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ da,
                                         long long da_ns, int N, int C, int J, long long HW) {
  const long long total = static_cast<long long>(N) * C * HW;
  GEN_LOOP(idx, total) {
    const long long q = idx % HW;
    const int c = static_cast<int>((idx / HW) % C);
    const long long n = idx / (HW * C);
    float acc = 0.f;
    for (int j = 0; j < J; ++j) acc = fmaf(dz[(n * J + j) * HW + q], w[static_cast<long long>(j) * C + c], acc);
    da[n * da_ns + static_cast<long long>(c) * HW + q] = acc;
  }
}
// one block per (c, j): dw[j][c] = sum dz*a, db[j] = sum dz
__global__ void gen_conv1x1_wgrad_kernel(const float* __restrict__ dz, const float* __restrict__ a, long long a_ns,
                                         float* __restrict__ dw, float* __restrict__ db, int N, int C, int J, long long HW) {
  __shared__ double sh[8];
  const int c = blockIdx.x, j = blockIdx.y;
  double s1 = 0.0, s2 = 0.0;
  const long long P = static_cast<long long>(N) * HW;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const long long n = p / HW, q = p - n * HW;
    const double g = dz[(n * J + j) * HW + q];
    s1 += g * static_cast<double>(a[n * a_ns + static_cast<long long>(c) * HW + q]);
    s2 += g;
  }
  const double t1 = block_sum(s1, sh);
  if (threadIdx.x == 0) dw[static_cast<long long>(j) * C + c] = static_cast<float>(t1);
  if (c == 0 && db != nullptr) {
    const double t2 = block_sum(s2, sh);
    if (threadIdx.x == 0) db[j] = static_cast<float>(t2);
  }
}

// x[n,c,q] *= mask[n,c,q] (nn.Dropout forward and backward with the saved keep/scale mask, Model.py:37,81-82)
__global__ void gen_mul_kernel(float* __restrict__ x, long long x_ns, const float* __restrict__ mask, int N, long long CHW) {
  const long long total = static_cast<long long>(N) * CHW;
  GEN_LOOP(idx, total) {
    const long long n = idx / CHW, q = idx - n * CHW;
    x[n * x_ns + q] *= mask[idx];
  }
}

// ---- attention gates of UNet_attention (Attention_block, Model.py:257-296): BatchNorm without ReLU (W_q, W_x) or followed
// by a sigmoid (psi), E = relu(Q1 + X1), out = x * A with A broadcast over the channels, and their backward.
// act: 0 = identity, 1 = relu, 2 = sigmoid
__global__ void gen_bn_act_fwd_kernel(const float* __restrict__ y, long long y_ns, const float* __restrict__ scale,
                                      const float* __restrict__ shift, float* __restrict__ a, long long a_ns, int N, int C,
                                      int HW, int act) {
  const long long total = static_cast<long long>(N) * C * HW;
  GEN_LOOP(idx, total) {
    const long long q = idx % HW;
    const int c = static_cast<int>((idx / HW) % C);
    const long long n = idx / (static_cast<long long>(HW) * C);
    float v = fmaf(scale[c], y[n * y_ns + static_cast<long long>(c) * HW + q], shift[c]);
    if (act == 1) v = fmaxf(v, 0.f);
    if (act == 2) v = 1.f / (1.f + expf(-v));
    a[n * a_ns + static_cast<long long>(c) * HW + q] = v;
  }
}

// e = relu(a + b) over dense tensors
__global__ void gen_add_relu_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ e, long long total) {
  GEN_LOOP(idx, total) e[idx] = fmaxf(a[idx] + b[idx], 0.f);
}
// d = de * [e > 0]
__global__ void gen_relu_bwd_kernel(const float* __restrict__ de, const float* __restrict__ e, float* __restrict__ d, long long total) {
  GEN_LOOP(idx, total) d[idx] = e[idx] > 0.f ? de[idx] : 0.f;
}
// out[n,c,q] = x[n,c,q] * gate[n,q]
__global__ void gen_gate_fwd_kernel(const float* __restrict__ x, long long x_ns, const float* __restrict__ gate,
                                    float* __restrict__ out, long long out_ns, int N, int C, long long HW) {
  const long long total = static_cast<long long>(N) * C * HW;
  GEN_LOOP(idx, total) {
    const long long q = idx % HW;
    const long long c = (idx / HW) % C;
    const long long n = idx / (HW * C);
    out[n * out_ns + c * HW + q] = x[n * x_ns + c * HW + q] * gate[n * HW + q];
  }
}
// dx[n,c,q] = dout[n,c,q] * gate[n,q]
__global__ void gen_gate_bwd_x_kernel(const float* __restrict__ dout, long long dout_ns, const float* __restrict__ gate,
                                      float* __restrict__ dx, long long dx_ns, int N, int C, long long HW) {
  const long long total = static_cast<long long>(N) * C * HW;
  GEN_LOOP(idx, total) {
    const long long q = idx % HW;
    const long long c = (idx / HW) % C;
    const long long n = idx / (HW * C);
    dx[n * dx_ns + c * HW + q] = dout[n * dout_ns + c * HW + q] * gate[n * HW + q];
  }
}
// dpre[n,q] = (sum_c dout[n,c,q] * x[n,c,q]) * A (1 - A): gradient w.r.t. the sigmoid's input (A = gate)
__global__ void gen_gate_bwd_gate_kernel(const float* __restrict__ dout, long long dout_ns, const float* __restrict__ x,
                                         long long x_ns, const float* __restrict__ gate, float* __restrict__ dpre, int N, int C,
                                         long long HW) {
  const long long total = static_cast<long long>(N) * HW;
  GEN_LOOP(idx, total) {
    const long long q = idx % HW, n = idx / HW;
    double s = 0.0;
    for (int c = 0; c < C; ++c)
      s += static_cast<double>(dout[n * dout_ns + c * HW + q]) * static_cast<double>(x[n * x_ns + c * HW + q]);
    const double a = gate[idx];
    dpre[idx] = static_cast<float>(s * a * (1.0 - a));
  }
}
// dst[n, :] += src[n, :] over views with their own batch strides
__global__ void gen_add_inplace_kernel(float* __restrict__ dst, long long dst_ns, const float* __restrict__ src, long long src_ns,
                                       int N, long long CHW) {
  const long long total = static_cast<long long>(N) * CHW;
  GEN_LOOP(idx, total) {
    const long long n = idx / CHW, q = idx - n * CHW;
    dst[n * dst_ns + q] += src[n * src_ns + q];
  }
}

}  // namespace

#define GEN_STREAM static_cast<cudaStream_t>(stream)

extern "C" {

int b200unet_gen_conv3x3(const float* in, int64_t in_ns, const float* w, float* out, int64_t out_ns, int N, int Ci, int Co,
                         int H, int W, int transposed, b200_stream_t stream) {
  B2_REQUIRE(N > 0 && Ci > 0 && Co > 0 && H > 0 && W > 0, "gen_conv3x3: empty tensor");
  gen_conv3x3_kernel<<<gblocks(static_cast<long long>(N) * Co * H * W), GT, 0, GEN_STREAM>>>(in, in_ns, w, out, out_ns, N, Ci, Co, H, W, transposed);
  return b2h::check_launch("gen_conv3x3");
}

int b200unet_gen_conv3x3_wgrad(const float* x, int64_t x_ns, const float* dy, int64_t dy_ns, float* dw, int N, int Ci,
                               int Co, int H, int W, b200_stream_t stream) {
  B2_REQUIRE(N > 0 && Ci > 0 && Co > 0 && Co <= 65535, "gen_conv3x3_wgrad: bad shape");
  gen_conv3x3_wgrad_kernel<<<dim3(Ci, Co), GT, 0, GEN_STREAM>>>(x, x_ns, dy, dy_ns, dw, N, Ci, Co, H, W);
  return b2h::check_launch("gen_conv3x3_wgrad");
}

int b200unet_gen_channel_stats(const float* y, int64_t y_ns, double* sums, int N, int C, int HW, b200_stream_t stream) {
  gen_channel_stats_kernel<<<C, GT, 0, GEN_STREAM>>>(y, y_ns, sums, N, C, HW);
  return b2h::check_launch("gen_channel_stats");
}

int b200unet_gen_bn_relu_fwd(const float* y, int64_t y_ns, const float* scale, const float* shift, float* a, int64_t a_ns,
                             int N, int C, int HW, b200_stream_t stream) {
  gen_bn_relu_fwd_kernel<<<gblocks(static_cast<long long>(N) * C * HW), GT, 0, GEN_STREAM>>>(y, y_ns, scale, shift, a, a_ns, N, C, HW);
  return b2h::check_launch("gen_bn_relu_fwd");
}

int b200unet_gen_maxpool2x2(const float* a, int64_t a_ns, float* pooled, uint8_t* idx, int N, int C, int H, int W,
                            b200_stream_t stream) {
  B2_REQUIRE(H >= 2 && W >= 2, "gen_maxpool2x2: H=%d W=%d too small", H, W);
  gen_maxpool_kernel<<<gblocks(static_cast<long long>(N) * C * (H / 2) * (W / 2)), GT, 0, GEN_STREAM>>>(a, a_ns, pooled, idx, N, C, H, W);
  return b2h::check_launch("gen_maxpool2x2");
}

int b200unet_gen_unpool_add(const float* g_pooled, const uint8_t* idx, float* g, int64_t g_ns, int N, int C, int H, int W,
                            b200_stream_t stream) {
  gen_unpool_add_kernel<<<gblocks(static_cast<long long>(N) * C * (H / 2) * (W / 2)), GT, 0, GEN_STREAM>>>(g_pooled, idx, g, g_ns, N, C, H, W);
  return b2h::check_launch("gen_unpool_add");
}

int b200unet_gen_bn_relu_bwd_reduce(const float* g, int64_t g_ns, const float* y, int64_t y_ns, const float* scale,
                                    const float* shift, const float* mean, const float* rstd, double* sums, int N, int C,
                                    int HW, b200_stream_t stream) {
  gen_bn_bwd_reduce_kernel<<<C, GT, 0, GEN_STREAM>>>(g, g_ns, y, y_ns, scale, shift, mean, rstd, sums, N, C, HW, 1);
  return b2h::check_launch("gen_bn_relu_bwd_reduce");
}

int b200unet_gen_bn_relu_bwd_apply(const float* g, int64_t g_ns, const float* y, int64_t y_ns, const float* gamma,
                                   const float* scale, const float* shift, const float* mean, const float* rstd,
                                   const double* sums, double count, const double* sums_local, float* dy, int64_t dy_ns,
                                   float* dgamma, float* dbeta, int N, int C, int HW, b200_stream_t stream) {
  gen_bn_bwd_apply_kernel<<<gblocks(static_cast<long long>(N) * C * HW), GT, 0, GEN_STREAM>>>(
      g, g_ns, y, y_ns, gamma, scale, shift, mean, rstd, sums, count, sums_local, dy, dy_ns, dgamma, dbeta, N, C, HW, 1);
  return b2h::check_launch("gen_bn_relu_bwd_apply");
}

int b200unet_gen_convt2x2_fprop(const float* x, int64_t x_ns, const float* w, const float* bias, float* out, int64_t out_ns,
                                int N, int Cin, int Cup, int H, int W, int H2, int W2, int pad_top, int pad_left,
                                b200_stream_t stream) {
  B2_REQUIRE(pad_top >= 0 && pad_left >= 0 && 2 * H + pad_top <= H2 && 2 * W + pad_left <= W2, "gen_convt2x2_fprop: bad canvas");
  gen_convt_fprop_kernel<<<gblocks(static_cast<long long>(N) * Cup * 4 * H * W), GT, 0, GEN_STREAM>>>(
      x, x_ns, w, bias, out, out_ns, N, Cin, Cup, H, W, H2, W2, pad_top, pad_left);
  return b2h::check_launch("gen_convt2x2_fprop");
}

int b200unet_gen_convt2x2_dgrad(const float* du, int64_t du_ns, const float* w, float* dx, int64_t dx_ns, int N, int Cin,
                                int Cup, int H, int W, int H2, int W2, int pad_top, int pad_left, b200_stream_t stream) {
  B2_REQUIRE(pad_top >= 0 && pad_left >= 0 && 2 * H + pad_top <= H2 && 2 * W + pad_left <= W2, "gen_convt2x2_dgrad: bad canvas");
  gen_convt_dgrad_kernel<<<gblocks(static_cast<long long>(N) * Cin * H * W), GT, 0, GEN_STREAM>>>(
      du, du_ns, w, dx, dx_ns, N, Cin, Cup, H, W, H2, W2, pad_top, pad_left);
  return b2h::check_launch("gen_convt2x2_dgrad");
}

int b200unet_gen_convt2x2_wgrad(const float* x, int64_t x_ns, const float* du, int64_t du_ns, float* dw, float* db, int N,
                                int Cin, int Cup, int H, int W, int H2, int W2, int pad_top, int pad_left,
                                b200_stream_t stream) {
  B2_REQUIRE(Cin <= 65535 && pad_top >= 0 && pad_left >= 0 && 2 * H + pad_top <= H2 && 2 * W + pad_left <= W2, "gen_convt2x2_wgrad: bad shape");
  gen_convt_wgrad_kernel<<<dim3(Cup, Cin), GT, 0, GEN_STREAM>>>(x, x_ns, du, du_ns, dw, db, N, Cin, Cup, H, W, H2, W2, pad_top, pad_left);
  return b2h::check_launch("gen_convt2x2_wgrad");
}

int b200unet_gen_conv1x1_fwd(const float* a, int64_t a_ns, const float* w, const float* bias, float* z, int N, int C, int J,
                             int64_t HW, b200_stream_t stream) {
  gen_conv1x1_fwd_kernel<<<gblocks(static_cast<long long>(N) * J * HW), GT, 0, GEN_STREAM>>>(a, a_ns, w, bias, z, N, C, J, HW);
  return b2h::check_launch("gen_conv1x1_fwd");
}

int b200unet_gen_conv1x1_bwd(const float* dz, const float* a, int64_t a_ns, const float* w, float* da, int64_t da_ns,
                             float* dw, float* db, int N, int C, int J, int64_t HW, b200_stream_t stream) {
  gen_conv1x1_dgrad_kernel<<<gblocks(static_cast<long long>(N) * C * HW), GT, 0, GEN_STREAM>>>(dz, w, da, da_ns, N, C, J, HW);
  if (int e = b2h::check_launch("gen_conv1x1_dgrad")) return e;
  gen_conv1x1_wgrad_kernel<<<dim3(C, J), GT, 0, GEN_STREAM>>>(dz, a, a_ns, dw, db, N, C, J, HW);
  return b2h::check_launch("gen_conv1x1_wgrad");
}

int b200unet_gen_bn_act_fwd(const float* y, int64_t y_ns, const float* scale, const float* shift, float* a, int64_t a_ns, int N,
                            int C, int HW, int act, b200_stream_t stream) {
  B2_REQUIRE(act >= 0 && act <= 2, "gen_bn_act_fwd: act=%d (0 identity, 1 relu, 2 sigmoid)", act);
  gen_bn_act_fwd_kernel<<<gblocks(static_cast<long long>(N) * C * HW), GT, 0, GEN_STREAM>>>(y, y_ns, scale, shift, a, a_ns, N, C, HW, act);
  return b2h::check_launch("gen_bn_act_fwd");
}

int b200unet_gen_bn_bwd_reduce(const float* g, int64_t g_ns, const float* y, int64_t y_ns, const float* scale, const float* shift,
                               const float* mean, const float* rstd, double* sums, int N, int C, int HW, int relu,
                               b200_stream_t stream) {
  gen_bn_bwd_reduce_kernel<<<C, GT, 0, GEN_STREAM>>>(g, g_ns, y, y_ns, scale, shift, mean, rstd, sums, N, C, HW, relu);
  return b2h::check_launch("gen_bn_bwd_reduce");
}

int b200unet_gen_bn_bwd_apply(const float* g, int64_t g_ns, const float* y, int64_t y_ns, const float* gamma, const float* scale,
                              const float* shift, const float* mean, const float* rstd, const double* sums, double count,
                              const double* sums_local, float* dy, int64_t dy_ns, float* dgamma, float* dbeta, int N, int C,
                              int HW, int relu, b200_stream_t stream) {
  gen_bn_bwd_apply_kernel<<<gblocks(static_cast<long long>(N) * C * HW), GT, 0, GEN_STREAM>>>(
      g, g_ns, y, y_ns, gamma, scale, shift, mean, rstd, sums, count, sums_local, dy, dy_ns, dgamma, dbeta, N, C, HW, relu);
  return b2h::check_launch("gen_bn_bwd_apply");
}

int b200unet_gen_add_relu(const float* a, const float* b, float* e, int64_t total, b200_stream_t stream) {
  gen_add_relu_kernel<<<gblocks(total), GT, 0, GEN_STREAM>>>(a, b, e, total);
  return b2h::check_launch("gen_add_relu");
}

int b200unet_gen_relu_bwd(const float* de, const float* e, float* d, int64_t total, b200_stream_t stream) {
  gen_relu_bwd_kernel<<<gblocks(total), GT, 0, GEN_STREAM>>>(de, e, d, total);
  return b2h::check_launch("gen_relu_bwd");
}

int b200unet_gen_gate_fwd(const float* x, int64_t x_ns, const float* gate, float* out, int64_t out_ns, int N, int C, int64_t HW,
                          b200_stream_t stream) {
  gen_gate_fwd_kernel<<<gblocks(static_cast<long long>(N) * C * HW), GT, 0, GEN_STREAM>>>(x, x_ns, gate, out, out_ns, N, C, HW);
  return b2h::check_launch("gen_gate_fwd");
}

int b200unet_gen_gate_bwd(const float* dout, int64_t dout_ns, const float* x, int64_t x_ns, const float* gate, float* dx,
                          int64_t dx_ns, float* dpre, int N, int C, int64_t HW, b200_stream_t stream) {
  gen_gate_bwd_x_kernel<<<gblocks(static_cast<long long>(N) * C * HW), GT, 0, GEN_STREAM>>>(dout, dout_ns, gate, dx, dx_ns, N, C, HW);
  if (int e = b2h::check_launch("gen_gate_bwd_x")) return e;
  gen_gate_bwd_gate_kernel<<<gblocks(static_cast<long long>(N) * HW), GT, 0, GEN_STREAM>>>(dout, dout_ns, x, x_ns, gate, dpre, N, C, HW);
  return b2h::check_launch("gen_gate_bwd_gate");
}

int b200unet_gen_add_inplace(float* dst, int64_t dst_ns, const float* src, int64_t src_ns, int N, int64_t CHW,
                             b200_stream_t stream) {
  gen_add_inplace_kernel<<<gblocks(static_cast<long long>(N) * CHW), GT, 0, GEN_STREAM>>>(dst, dst_ns, src, src_ns, N, CHW);
  return b2h::check_launch("gen_add_inplace");
}

int b200unet_gen_mul(float* x, int64_t x_ns, const float* mask, int N, int64_t CHW, b200_stream_t stream) {
  gen_mul_kernel<<<gblocks(static_cast<long long>(N) * CHW), GT, 0, GEN_STREAM>>>(x, x_ns, mask, N, CHW);
  return b2h::check_launch("gen_mul");
}

}  // extern "C"
