// Attention gates of UNet_attention (reference Model.py:257-296, Attention_block.forward :286-296) on the tensor-core
// engine's NHWC bf16 activations:
//     Q1 = BN_q(W_q(up(q)))    X1 = BN_x(W_x(x))    E = relu(Q1 + X1)    A = sigmoid(BN_p(psi(E)))    out = x * A
// The two 1x1 convolutions (and the gate's ConvTranspose2d, composed with W_q on the weight side) run as tcgen05 GEMMs
// (igemm.cu); this file holds the bandwidth-bound remainder. E and A are never stored: the forward keeps the two
// pre-BatchNorm maps (bf16, C channels) and the per-pixel psi pre-activation s (fp32), the backward recomputes E from them.
//   gate_psi_fwd     s = b_psi + sum_c w_psi[c] * relu(sq*Q + tq + sx*X + tx)[c]   (+ sum s, sum s^2 rows for BN_p)
//   gate_apply_fwd   out = x * sigmoid(sp*s + tp)                                   (into the concat buffer's skip half)
//   gate_apply_bwd   dA = sum_c g*x, dz = dA*A*(1-A), dx = g*A                     (+ sum dz, sum dz*shat rows)
//   gate_bwd_reduce  ds = BN_p backward of dz; dE = ds*w_psi*[E>0]; per-channel sum dE, sum dE*Qhat, sum dE*Xhat, sum ds*E
//   gate_bwd_apply   dQ = BN_q backward of dE, dX = BN_x backward of dE (bf16, in place over Q and X), parameter gradients
// plus a strided fp32 GEMM used on the weight side (composition of up and W_q and its backward).
// A pixel is handled by C/8 consecutive lanes (one 128-bit load per map and lane); per-pixel channel reductions are
// xor-shuffles inside that lane group.
#include "../../include/b200unet.h"
#include "ew_common.cuh"
#include "host_common.h"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

constexpr int GT = 256;
constexpr int G_MAX_BLOCKS = 148 * 8;

using namespace b2ew;
__device__ __forceinline__ float sigmoidf(float v) { return 1.f / (1.f + __expf(-v)); }

int gate_blocks(long long items, int per_sm) {
  long long b = (items + GT - 1) / GT;
  const long long cap = 148ll * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// block-wide sum of two per-thread values -> dst[0], dst[1] (thread 0 writes)
__device__ __forceinline__ void block_sum2(float a1, float a2, float* dst) {
  __shared__ float sh2[2][GT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh2[0][threadIdx.x >> 5] = a1;
    sh2[1][threadIdx.x >> 5] = a2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < GT / 32; ++i) {
      t1 += sh2[0][i];
      t2 += sh2[1][i];
    }
    dst[0] = t1;
    dst[1] = t2;
  }
}

struct GateMaps {
  const __nv_bfloat16* q;  // pre-BatchNorm W_q(up(q)) map
  int q_cs;
  const __nv_bfloat16* x;  // pre-BatchNorm W_x(x) map
  int x_cs;
  const float *sq, *tq, *sx, *tx;  // BatchNorm affine of the two maps (scale, shift), [C] each
  const float* wpsi;               // psi convolution weight [C]
};

// the per-lane constants of one channel group
struct GateConsts {
  float sq[8], sx[8], t[8], wp[8];
  __device__ __forceinline__ void load(const GateMaps& m, int g) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      sq[j] = m.sq[c];
      sx[j] = m.sx[c];
      t[j] = m.tq[c] + m.tx[c];
      wp[j] = m.wpsi[c];
    }
  }
  __device__ __forceinline__ float e(int j, float qv, float xv) const { return fmaxf(fmaf(sq[j], qv, fmaf(sx[j], xv, t[j])), 0.f); }
};

// ------------------------------------------------------------------------------------------------ forward
template <int TPP>
__global__ void __launch_bounds__(GT) gate_psi_fwd_kernel(GateMaps m, const float* __restrict__ bpsi, float* __restrict__ s_out,
                                                          float* __restrict__ partial, long long pixels) {
  const int g = threadIdx.x % TPP;
  GateConsts k;
  k.load(m, g);
  const float b = bpsi[0];
  float a1 = 0.f, a2 = 0.f;
  const long long total = pixels * TPP;
  const long long stride = static_cast<long long>(gridDim.x) * GT;
  constexpr int U = 2;  // pixels per thread and iteration: all loads are issued before the arithmetic (bytes in flight)
  for (long long i0 = static_cast<long long>(blockIdx.x) * GT + threadIdx.x;; i0 += stride * U) {
    if (!__any_sync(0xffffffffu, i0 < total)) break;
    uint4 qv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        qv[u] = ldg128(m.q + (i / TPP) * m.q_cs + g * 8);
        xv[u] = ldg128(m.x + (i / TPP) * m.x_cs + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      const bool valid = i < total;
      const long long pix = i / TPP;
      float p = 0.f;
      if (valid) {
        float qf[8], xf[8];
        unpack8(qv[u], qf);
        unpack8(xv[u], xf);
#pragma unroll
        for (int j = 0; j < 8; ++j) p = fmaf(k.wp[j], k.e(j, qf[j], xf[j]), p);
      }
#pragma unroll
      for (int o = TPP / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (valid && g == 0) {
        const float s = p + b;
        s_out[pix] = s;
        a1 += s;
        a2 = fmaf(s, s, a2);
      }
    }
  }
  if (partial != nullptr) block_sum2(a1, a2, partial + 2 * static_cast<size_t>(blockIdx.x));
}

__global__ void __launch_bounds__(GT) gate_apply_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_cs,
                                                            const float* __restrict__ s, const float* __restrict__ scale_p,
                                                            const float* __restrict__ shift_p, __nv_bfloat16* __restrict__ out,
                                                            int out_cs, long long pixels, int cgx) {
  const float sp = scale_p[0], tp = shift_p[0];
  const long long total = pixels * cgx;
  const long long stride = static_cast<long long>(gridDim.x) * GT;
  for (long long i = static_cast<long long>(blockIdx.x) * GT + threadIdx.x; i < total; i += stride) {
    const long long pix = i / cgx;
    const int g = static_cast<int>(i - pix * cgx);
    const float a = sigmoidf(fmaf(sp, __ldg(s + pix), tp));
    float f[8];
    unpack8(ldg128(x + pix * x_cs + g * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= a;
    *reinterpret_cast<uint4*>(out + pix * out_cs + g * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// TPP lanes per pixel, each covering GPT channel groups (Cx = 8 * TPP * GPT); two pixels in flight per thread. dx == nullptr:
// only dz and the sums (the engine adds g*A to the W_x backward-data later, gate_dx_kernel).
template <int TPP, int GPT>
__global__ void __launch_bounds__(GT) gate_apply_bwd_kernel(const __nv_bfloat16* __restrict__ gr, int g_cs,
                                                            const __nv_bfloat16* __restrict__ x, int x_cs,
                                                            const float* __restrict__ s, const float* __restrict__ scale_p,
                                                            const float* __restrict__ shift_p, const float* __restrict__ mean_p,
                                                            const float* __restrict__ rstd_p, __nv_bfloat16* __restrict__ dx,
                                                            int dx_cs, float* __restrict__ dz, float* __restrict__ partial,
                                                            long long pixels) {
  constexpr int U = 2;
  const int lg = threadIdx.x % TPP;
  const float sp = scale_p[0], tp = shift_p[0], mp = mean_p[0], rp = rstd_p[0];
  float a1 = 0.f, a2 = 0.f;
  const long long total = pixels * TPP;
  const long long stride = static_cast<long long>(gridDim.x) * GT;
  for (long long i0 = static_cast<long long>(blockIdx.x) * GT + threadIdx.x;; i0 += stride * U) {
    if (!__any_sync(0xffffffffu, i0 < total)) break;
    uint4 gv[U][GPT], xv[U][GPT];
    float sv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        const long long pix = i / TPP;
#pragma unroll
        for (int v = 0; v < GPT; ++v) {
          gv[u][v] = ldg128(gr + pix * g_cs + (lg + v * TPP) * 8);
          xv[u][v] = ldg128(x + pix * x_cs + (lg + v * TPP) * 8);
        }
        sv[u] = __ldg(s + pix);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      const bool valid = i < total;
      const long long pix = i / TPP;
      float d = 0.f;
      if (valid) {
#pragma unroll
        for (int v = 0; v < GPT; ++v) {
          float gf[8], xf[8];
          unpack8(gv[u][v], gf);
          unpack8(xv[u][v], xf);
#pragma unroll
          for (int j = 0; j < 8; ++j) d = fmaf(gf[j], xf[j], d);
        }
      }
#pragma unroll
      for (int o = TPP / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (valid) {
        const float a = sigmoidf(fmaf(sp, sv[u], tp));
        if (dx != nullptr) {
#pragma unroll
          for (int v = 0; v < GPT; ++v) {
            float gf[8];
            unpack8(gv[u][v], gf);
#pragma unroll
            for (int j = 0; j < 8; ++j) gf[j] *= a;
            *reinterpret_cast<uint4*>(dx + pix * dx_cs + (lg + v * TPP) * 8) = pack8(gf);
          }
        }
        if (lg == 0) {
          const float v = d * a * (1.f - a);
          dz[pix] = v;
          a1 += v;
          a2 = fmaf(v, (sv[u] - mp) * rp, a2);
        }
      }
    }
  }
  block_sum2(a1, a2, partial + 2 * static_cast<size_t>(blockIdx.x));
}

// dx <- g * sigmoid(sp*s + tp) + dx: the gradient through the product x * A added to the W_x backward-data already in dx
__global__ void __launch_bounds__(GT) gate_dx_kernel(const __nv_bfloat16* __restrict__ gr, int g_cs, const float* __restrict__ s,
                                                     const float* __restrict__ scale_p, const float* __restrict__ shift_p,
                                                     __nv_bfloat16* dx, int dx_cs, long long pixels, int cgx) {
  const float sp = scale_p[0], tp = shift_p[0];
  const long long total = pixels * cgx;
  const long long stride = static_cast<long long>(gridDim.x) * GT;
  for (long long i = static_cast<long long>(blockIdx.x) * GT + threadIdx.x; i < total; i += stride) {
    const long long pix = i / cgx;
    const int g = static_cast<int>(i - pix * cgx);
    const uint4 gv = ldg128(gr + pix * g_cs + g * 8);
    const uint4 dv = *reinterpret_cast<const uint4*>(dx + pix * dx_cs + g * 8);
    const float a = sigmoidf(fmaf(sp, __ldg(s + pix), tp));
    float gf[8], df[8];
    unpack8(gv, gf);
    unpack8(dv, df);
#pragma unroll
    for (int j = 0; j < 8; ++j) df[j] = fmaf(gf[j], a, df[j]);
    *reinterpret_cast<uint4*>(dx + pix * dx_cs + g * 8) = pack8(df);
  }
}

// BatchNorm2d(1) backward of the psi branch for one pixel: ds = gamma*rstd*(dz - sum(dz)/m - shat*sum(dz*shat)/m)
struct PsiBwd {
  float k1, k2, k3, mp, rp;
  __device__ __forceinline__ void init(const float* gamma_p, const float* mean_p, const float* rstd_p, const double* sums2,
                                       double count) {
    const double g = gamma_p[0], r = rstd_p[0];
    mp = mean_p[0];
    rp = rstd_p[0];
    k1 = static_cast<float>(g * r);
    k2 = static_cast<float>(-g * r * sums2[1] / count);
    k3 = static_cast<float>(-g * r * sums2[0] / count);
  }
  __device__ __forceinline__ float ds(float dzv, float sv) const { return fmaf(k1, dzv, fmaf(k2, (sv - mp) * rp, k3)); }
};

// partial row layout (pitch 4C + 8): [0,C) sum dE; [C,2C) sum dE*Qhat; [2C,3C) sum dE*Xhat; [3C,4C) sum ds*E; [4C] sum ds
template <int TPP>
__global__ void __launch_bounds__(GT) gate_bwd_reduce_kernel(GateMaps m, const float* __restrict__ mean_q,
                                                             const float* __restrict__ rstd_q, const float* __restrict__ mean_x,
                                                             const float* __restrict__ rstd_x, const float* __restrict__ s,
                                                             const float* __restrict__ dz, const float* __restrict__ gamma_p,
                                                             const float* __restrict__ mean_p, const float* __restrict__ rstd_p,
                                                             const double* __restrict__ sums2, double count,
                                                             float* __restrict__ ds_out, float* __restrict__ partial,
                                                             long long pixels) {
  constexpr int C = TPP * 8;
  const int g = threadIdx.x % TPP;
  GateConsts k;
  k.load(m, g);
  PsiBwd pb;
  pb.init(gamma_p, mean_p, rstd_p, sums2, count);
  float s1[8], s2[8], s3[8], s4[8], sds = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = s3[j] = s4[j] = 0.f;
  const long long total = pixels * TPP;
  const long long stride = static_cast<long long>(gridDim.x) * GT;
  constexpr int U = 2;
  for (long long i0 = static_cast<long long>(blockIdx.x) * GT + threadIdx.x; i0 < total; i0 += stride * U) {
    uint4 qv[U], xv[U];
    float dzv[U], sv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        const long long pix = i / TPP;
        qv[u] = ldg128(m.q + pix * m.q_cs + g * 8);
        xv[u] = ldg128(m.x + pix * m.x_cs + g * 8);
        dzv[u] = __ldg(dz + pix);
        sv[u] = __ldg(s + pix);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= total) break;
      const long long pix = i / TPP;
      float qf[8], xf[8];
      unpack8(qv[u], qf);
      unpack8(xv[u], xf);
      const float ds = pb.ds(dzv[u], sv[u]);
      if (g == 0) {
        ds_out[pix] = ds;
        sds += ds;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float e = k.e(j, qf[j], xf[j]);
        const float de = e > 0.f ? ds * k.wp[j] : 0.f;
        s1[j] += de;
        s2[j] = fmaf(de, qf[j], s2[j]);
        s3[j] = fmaf(de, xf[j], s3[j]);
        s4[j] = fmaf(ds, e, s4[j]);
      }
    }
  }
  __shared__ float red[GT][33];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x][j] = s1[j];
    red[threadIdx.x][8 + j] = s2[j];
    red[threadIdx.x][16 + j] = s3[j];
    red[threadIdx.x][24 + j] = s4[j];
  }
  red[threadIdx.x][32] = sds;
  __syncthreads();
  float* row = partial + static_cast<size_t>(blockIdx.x) * (4 * C + 8);
  for (int ch = threadIdx.x; ch < C; ch += GT) {
    const int cg = ch >> 3, j = ch & 7;
    double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
    for (int th = cg; th < GT; th += TPP) {
      a1 += static_cast<double>(red[th][j]);
      a2 += static_cast<double>(red[th][8 + j]);
      a3 += static_cast<double>(red[th][16 + j]);
      a4 += static_cast<double>(red[th][24 + j]);
    }
    row[ch] = static_cast<float>(a1);
    row[C + ch] = static_cast<float>(static_cast<double>(rstd_q[ch]) * (a2 - static_cast<double>(mean_q[ch]) * a1));
    row[2 * C + ch] = static_cast<float>(static_cast<double>(rstd_x[ch]) * (a3 - static_cast<double>(mean_x[ch]) * a1));
    row[3 * C + ch] = static_cast<float>(a4);
  }
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int th = 0; th < GT; th += TPP) t += static_cast<double>(red[th][32]);
    row[4 * C] = static_cast<float>(t);
    for (int i = 1; i < 8; ++i) row[4 * C + i] = 0.f;  // alignment padding of the row: defined values for the row reduction
  }
}

struct GateBwdOut {
  float *dgamma_q, *dbeta_q, *dgamma_x, *dbeta_x, *dwpsi, *dbpsi, *dgamma_p, *dbeta_p;
};

// partial2 row layout (pitch 2C): [0,C) sum dQ; [C,2C) sum dX (of the bf16 values stored: the bias gradients of the two convs)
template <int TPP>
__global__ void __launch_bounds__(GT, 2) gate_bwd_apply_kernel(GateMaps m, const float* __restrict__ gamma_q,
                                                            const float* __restrict__ mean_q, const float* __restrict__ rstd_q,
                                                            const float* __restrict__ gamma_x, const float* __restrict__ mean_x,
                                                            const float* __restrict__ rstd_x, const float* __restrict__ ds,
                                                            const double* __restrict__ sums, const double* __restrict__ sums_local,
                                                            const double* __restrict__ sums2_local, double count, GateBwdOut o,
                                                            __nv_bfloat16* dq, __nv_bfloat16* dxm,
                                                            float* __restrict__ partial2, long long pixels) {
  constexpr int C = TPP * 8;
  const int g = threadIdx.x % TPP;
  GateConsts k;
  k.load(m, g);
  float q1[8], q2[8], q3[8], x1[8], x2[8], x3[8], aq[8], ax[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    const double sde = sums[c] / count;
    {
      const double gm = gamma_q[c], r = rstd_q[c], mu = mean_q[c], sdx = sums[C + c] / count;
      q1[j] = static_cast<float>(gm * r);
      q2[j] = static_cast<float>(-gm * r * r * sdx);
      q3[j] = static_cast<float>(-gm * r * sde + gm * r * r * mu * sdx);
    }
    {
      const double gm = gamma_x[c], r = rstd_x[c], mu = mean_x[c], sdx = sums[2 * C + c] / count;
      x1[j] = static_cast<float>(gm * r);
      x2[j] = static_cast<float>(-gm * r * r * sdx);
      x3[j] = static_cast<float>(-gm * r * sde + gm * r * r * mu * sdx);
    }
    aq[j] = ax[j] = 0.f;
  }
  if (blockIdx.x == 0) {  // parameter gradients keep the LOCAL sums (SyncBN averages them with the other gradients later)
    const double* sl = sums_local != nullptr ? sums_local : sums;
    for (int c = threadIdx.x; c < C; c += GT) {
      o.dbeta_q[c] = static_cast<float>(sl[c]);
      o.dgamma_q[c] = static_cast<float>(sl[C + c]);
      o.dbeta_x[c] = static_cast<float>(sl[c]);
      o.dgamma_x[c] = static_cast<float>(sl[2 * C + c]);
      o.dwpsi[c] = static_cast<float>(sl[3 * C + c]);
    }
    if (threadIdx.x == 0) {
      o.dbpsi[0] = static_cast<float>(sl[4 * C]);
      o.dbeta_p[0] = static_cast<float>(sums2_local[0]);
      o.dgamma_p[0] = static_cast<float>(sums2_local[1]);
    }
  }
  const long long total = pixels * TPP;
  const long long stride = static_cast<long long>(gridDim.x) * GT;
  constexpr int U = 2;
  for (long long i0 = static_cast<long long>(blockIdx.x) * GT + threadIdx.x; i0 < total; i0 += stride * U) {
    uint4 qv[U], xv[U];
    float dsu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        const long long pix = i / TPP;
        qv[u] = ldg128(m.q + pix * m.q_cs + g * 8);
        xv[u] = ldg128(m.x + pix * m.x_cs + g * 8);
        dsu[u] = __ldg(ds + pix);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= total) break;
      const long long pix = i / TPP;
      float qf[8], xf[8], oq[8], ox[8];
      unpack8(qv[u], qf);
      unpack8(xv[u], xf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float de = k.e(j, qf[j], xf[j]) > 0.f ? dsu[u] * k.wp[j] : 0.f;
        oq[j] = fmaf(q1[j], de, fmaf(q2[j], qf[j], q3[j]));
        ox[j] = fmaf(x1[j], de, fmaf(x2[j], xf[j], x3[j]));
      }
      const uint4 pq = pack8(oq), px = pack8(ox);
      *reinterpret_cast<uint4*>(dq + pix * m.q_cs + g * 8) = pq;
      *reinterpret_cast<uint4*>(dxm + pix * m.x_cs + g * 8) = px;
      unpack8(pq, oq);
      unpack8(px, ox);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        aq[j] += oq[j];
        ax[j] += ox[j];
      }
    }
  }
  __shared__ float red[GT][17];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x][j] = aq[j];
    red[threadIdx.x][8 + j] = ax[j];
  }
  __syncthreads();
  float* row = partial2 + static_cast<size_t>(blockIdx.x) * 2 * C;
  for (int ch = threadIdx.x; ch < C; ch += GT) {
    const int cg = ch >> 3, j = ch & 7;
    float t1 = 0.f, t2 = 0.f;
    for (int th = cg; th < GT; th += TPP) {
      t1 += red[th][j];
      t2 += red[th][8 + j];
    }
    row[ch] = t1;
    row[C + ch] = t2;
  }
}

// ------------------------------------------------------------------------------------------------ weight-side GEMM
// C[z][m][n] (+)= sum_k A[z][m][k] * B[z][k][n] (+ bias[m]) over arbitrary element strides, fp32. 128 x 64 tile per CTA,
// 8 x 4 outputs per thread, K in steps of 16 through a double-buffered shared-memory tile (the next step's global loads are
// in flight while the current one is multiplied).
struct SgemmArgs {
  const float *A, *B;
  float* C;
  const float* bias_m;
  int M, N, K;
  long long am, ak, bk, bn, cm, cn, az, bz, cz;
  int accumulate;
};

constexpr int SG_TM = 128, SG_TN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256) sgemm_strided_kernel(SgemmArgs a) {
  __shared__ __align__(16) float As[2][SG_BK][SG_TM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_TN + 4];
  const float* A = a.A + blockIdx.z * a.az;
  const float* B = a.B + blockIdx.z * a.bz;
  float* Cc = a.C + blockIdx.z * a.cz;
  const int m0 = blockIdx.y * SG_TM, n0 = blockIdx.x * SG_TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool a_kfast = a.ak <= a.am, b_nfast = a.bn <= a.bk;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[4];
  auto load_global = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = threadIdx.x + r * 256;
      const int kk = a_kfast ? (e & 15) : (e >> 7), mm = a_kfast ? (e >> 4) : (e & 127);
      const int gm = m0 + mm, gk = k0 + kk;
      ra[r] = (gm < a.M && gk < a.K) ? __ldg(A + gm * a.am + gk * a.ak) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + r * 256;
      const int kk = b_nfast ? (e >> 6) : (e & 15), nn = b_nfast ? (e & 63) : (e >> 4);
      const int gn = n0 + nn, gk = k0 + kk;
      rb[r] = (gn < a.N && gk < a.K) ? __ldg(B + gk * a.bk + gn * a.bn) : 0.f;
    }
  };
  auto store_shared = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = threadIdx.x + r * 256;
      const int kk = a_kfast ? (e & 15) : (e >> 7), mm = a_kfast ? (e >> 4) : (e & 127);
      As[buf][kk][mm] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + r * 256;
      const int kk = b_nfast ? (e >> 6) : (e & 15), nn = b_nfast ? (e & 63) : (e >> 4);
      Bs[buf][kk][nn] = rb[r];
    }
  };
  load_global(0);
  store_shared(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < a.K; k0 += SG_BK) {
    const bool more = k0 + SG_BK < a.K;
    if (more) load_global(k0 + SG_BK);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (more) {
      store_shared(buf ^ 1);  // the other buffer was last read before the barrier that ended the previous step
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= a.M) continue;
    const float bias = a.bias_m != nullptr ? a.bias_m[gm] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float* dst = Cc + gm * a.cm + gn * a.cn;
      float v = acc[i][j] + bias;
      if (a.accumulate) v += *dst;
      *dst = v;
    }
  }
}

// N = 1 (the two bias products of a gate): one warp per output row, lanes stride over k
__global__ void __launch_bounds__(256) sgemv_strided_kernel(SgemmArgs a) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= a.M) return;
  float acc = 0.f;
  for (int k = lane; k < a.K; k += 32) acc = fmaf(__ldg(a.A + m * a.am + k * a.ak), __ldg(a.B + k * a.bk), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    float* dst = a.C + m * a.cm;
    float v = acc + (a.bias_m != nullptr ? a.bias_m[m] : 0.f);
    if (a.accumulate) v += *dst;
    *dst = v;
  }
}

// out[i] = sum_b part[b][i] (fixed order): the four (i,j) partial products of a gate's dW_q
__global__ void sum_batches_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, int batches) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < batches; ++b) acc += part[b * n + i];
    out[i] = acc;
  }
}

bool ok_gate_channels(int C) { return C == 32 || C == 64 || C == 128 || C == 256; }

}  // namespace

extern "C" {

int64_t b200unet_gate_workspace_floats(int C) { return static_cast<int64_t>(G_MAX_BLOCKS) * (4 * C + 8); }

int b200unet_gate_stat_rows(int64_t pixels, int C) { return gate_blocks(pixels * (C / 8), 8); }

#define GATE_MAPS(m)                                                                                              \
  GateMaps m;                                                                                                     \
  m.q = static_cast<const __nv_bfloat16*>(q1);                                                                    \
  m.q_cs = q1_cs;                                                                                                 \
  m.x = static_cast<const __nv_bfloat16*>(x1);                                                                    \
  m.x_cs = x1_cs;                                                                                                 \
  m.sq = scale_q;                                                                                                 \
  m.tq = shift_q;                                                                                                 \
  m.sx = scale_x;                                                                                                 \
  m.tx = shift_x;                                                                                                 \
  m.wpsi = w_psi

#define GATE_CHECK(name)                                                                                                    \
  B2_REQUIRE(ok_gate_channels(C), name ": C=%d must be 32, 64, 128 or 256", C);                                             \
  B2_REQUIRE(q1 && x1 && scale_q && shift_q && scale_x && shift_x && w_psi, name ": null argument");                        \
  B2_REQUIRE(pixels > 0 && q1_cs % 8 == 0 && x1_cs % 8 == 0 && q1_cs >= C && x1_cs >= C, name ": bad pitches / empty tensor")

int b200unet_gate_psi_fwd(const void* q1, int q1_cs, const void* x1, int x1_cs, const float* scale_q, const float* shift_q,
                          const float* scale_x, const float* shift_x, const float* w_psi, const float* b_psi, float* s,
                          float* stats_partial, int64_t pixels, int C, b200_stream_t stream) {
  GATE_CHECK("gate_psi_fwd");
  B2_REQUIRE(b_psi && s, "gate_psi_fwd: null argument");
  GATE_MAPS(m);
  const int blocks = gate_blocks(pixels * (C / 8), 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 32: gate_psi_fwd_kernel<4><<<blocks, GT, 0, st>>>(m, b_psi, s, stats_partial, pixels); break;
    case 64: gate_psi_fwd_kernel<8><<<blocks, GT, 0, st>>>(m, b_psi, s, stats_partial, pixels); break;
    case 128: gate_psi_fwd_kernel<16><<<blocks, GT, 0, st>>>(m, b_psi, s, stats_partial, pixels); break;
    default: gate_psi_fwd_kernel<32><<<blocks, GT, 0, st>>>(m, b_psi, s, stats_partial, pixels); break;
  }
  return b2h::check_launch("gate_psi_fwd");
}

int b200unet_gate_apply_fwd(const void* x, int x_cs, const float* s, const float* scale_p, const float* shift_p, void* out,
                            int out_cs, int64_t pixels, int Cx, b200_stream_t stream) {
  B2_REQUIRE(x && s && scale_p && shift_p && out, "gate_apply_fwd: null argument");
  B2_REQUIRE(pixels > 0 && Cx > 0 && Cx % 8 == 0 && x_cs % 8 == 0 && out_cs % 8 == 0 && x_cs >= Cx && out_cs >= Cx,
             "gate_apply_fwd: Cx=%d and the pitches (%d, %d) must be multiples of 8 with pitch >= Cx", Cx, x_cs, out_cs);
  gate_apply_fwd_kernel<<<gate_blocks(pixels * (Cx / 8), 8), GT, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cs, s, scale_p, shift_p, static_cast<__nv_bfloat16*>(out), out_cs, pixels, Cx / 8);
  return b2h::check_launch("gate_apply_fwd");
}

int b200unet_gate_apply_bwd(const void* g, int g_cs, const void* x, int x_cs, const float* s, const float* scale_p,
                            const float* shift_p, const float* mean_p, const float* rstd_p, void* dx, int dx_cs, float* dz,
                            float* workspace, double* sums2, int64_t pixels, int Cx, b200_stream_t stream) {
  B2_REQUIRE(g && x && s && scale_p && shift_p && mean_p && rstd_p && dz && workspace && sums2, "gate_apply_bwd: null argument");
  B2_REQUIRE(Cx == 64 || Cx == 128 || Cx == 256 || Cx == 512 || Cx == 1024, "gate_apply_bwd: Cx=%d must be a power of two in [64, 1024]", Cx);
  B2_REQUIRE(pixels > 0 && g_cs % 8 == 0 && x_cs % 8 == 0 && g_cs >= Cx && x_cs >= Cx && (dx == nullptr || (dx_cs % 8 == 0 && dx_cs >= Cx)),
             "gate_apply_bwd: bad pitches / empty tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(g);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dp = static_cast<__nv_bfloat16*>(dx);
  const int tpp = Cx / 8 < 32 ? Cx / 8 : 32;
  const int blocks = gate_blocks(pixels * tpp, 4);
#define GAB(T, G) gate_apply_bwd_kernel<T, G><<<blocks, GT, 0, st>>>(gp, g_cs, xp, x_cs, s, scale_p, shift_p, mean_p, rstd_p, dp, dx_cs, dz, workspace, pixels)
  switch (Cx) {
    case 64: GAB(8, 1); break;
    case 128: GAB(16, 1); break;
    case 256: GAB(32, 1); break;
    case 512: GAB(32, 2); break;
    default: GAB(32, 4); break;
  }
#undef GAB
  if (int e = b2h::check_launch("gate_apply_bwd")) return e;
  return b2h::reduce_partials_launch(workspace, blocks, 2, sums2, st);
}

int b200unet_gate_dx(const void* g, int g_cs, const float* s, const float* scale_p, const float* shift_p, void* dx, int dx_cs,
                     int64_t pixels, int Cx, b200_stream_t stream) {
  B2_REQUIRE(g && s && scale_p && shift_p && dx, "gate_dx: null argument");
  B2_REQUIRE(pixels > 0 && Cx > 0 && Cx % 8 == 0 && g_cs % 8 == 0 && dx_cs % 8 == 0 && g_cs >= Cx && dx_cs >= Cx,
             "gate_dx: Cx=%d and the pitches (%d, %d) must be multiples of 8 with pitch >= Cx", Cx, g_cs, dx_cs);
  gate_dx_kernel<<<gate_blocks(pixels * (Cx / 8), 8), GT, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(g), g_cs, s, scale_p, shift_p, static_cast<__nv_bfloat16*>(dx), dx_cs, pixels, Cx / 8);
  return b2h::check_launch("gate_dx");
}

int b200unet_gate_bwd_reduce(const void* q1, int q1_cs, const void* x1, int x1_cs, const float* scale_q, const float* shift_q,
                             const float* scale_x, const float* shift_x, const float* mean_q, const float* rstd_q,
                             const float* mean_x, const float* rstd_x, const float* w_psi, const float* s, const float* dz,
                             const float* gamma_p, const float* mean_p, const float* rstd_p, const double* sums2, double count,
                             float* ds, float* workspace, double* sums, int64_t pixels, int C, b200_stream_t stream) {
  GATE_CHECK("gate_bwd_reduce");
  B2_REQUIRE(mean_q && rstd_q && mean_x && rstd_x && s && dz && gamma_p && mean_p && rstd_p && sums2 && ds && workspace && sums,
             "gate_bwd_reduce: null argument");
  GATE_MAPS(m);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = gate_blocks(pixels * (C / 8), 2);
#define GBR(T) gate_bwd_reduce_kernel<T><<<blocks, GT, 0, st>>>(m, mean_q, rstd_q, mean_x, rstd_x, s, dz, gamma_p, mean_p, rstd_p, sums2, count, ds, workspace, pixels)
  switch (C) {
    case 32: GBR(4); break;
    case 64: GBR(8); break;
    case 128: GBR(16); break;
    default: GBR(32); break;
  }
#undef GBR
  if (int e = b2h::check_launch("gate_bwd_reduce")) return e;
  return b2h::reduce_partials_launch(workspace, blocks, 4 * C + 8, sums, st);
}

int b200unet_gate_bwd_apply(void* q1, int q1_cs, void* x1, int x1_cs, const float* scale_q, const float* shift_q,
                            const float* scale_x, const float* shift_x, const float* gamma_q, const float* mean_q,
                            const float* rstd_q, const float* gamma_x, const float* mean_x, const float* rstd_x,
                            const float* w_psi, const float* ds, const double* sums, const double* sums_local,
                            const double* sums2_local, double count, float* dgamma_q, float* dbeta_q, float* dgamma_x,
                            float* dbeta_x, float* dw_psi, float* db_psi, float* dgamma_p, float* dbeta_p, float* workspace,
                            float* dbias, int64_t pixels, int C, b200_stream_t stream) {
  GATE_CHECK("gate_bwd_apply");
  B2_REQUIRE(gamma_q && mean_q && rstd_q && gamma_x && mean_x && rstd_x && ds && sums && sums2_local && dgamma_q && dbeta_q &&
                 dgamma_x && dbeta_x && dw_psi && db_psi && dgamma_p && dbeta_p && workspace && dbias,
             "gate_bwd_apply: null argument");
  GATE_MAPS(m);
  GateBwdOut o{dgamma_q, dbeta_q, dgamma_x, dbeta_x, dw_psi, db_psi, dgamma_p, dbeta_p};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = gate_blocks(pixels * (C / 8), 2);
  __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(q1);
  __nv_bfloat16* dxm = static_cast<__nv_bfloat16*>(x1);
#define GBA(T) gate_bwd_apply_kernel<T><<<blocks, GT, 0, st>>>(m, gamma_q, mean_q, rstd_q, gamma_x, mean_x, rstd_x, ds, sums, sums_local, sums2_local, count, o, dq, dxm, workspace, pixels)
  switch (C) {
    case 32: GBA(4); break;
    case 64: GBA(8); break;
    case 128: GBA(16); break;
    default: GBA(32); break;
  }
#undef GBA
  if (int e = b2h::check_launch("gate_bwd_apply")) return e;
  return b2h::partial_colsum_launch(workspace, blocks, 2 * C, 0, 2 * C, dbias, st);
}

int b200unet_sgemm_strided(const float* A, const float* B, float* C, const float* bias_m, int M, int N, int K, int64_t am,
                           int64_t ak, int64_t bk, int64_t bn, int64_t cm, int64_t cn, int batch, int64_t az, int64_t bz,
                           int64_t cz, int accumulate, b200_stream_t stream) {
  B2_REQUIRE(A && B && C, "sgemm_strided: null argument");
  B2_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, "sgemm_strided: empty problem (M=%d N=%d K=%d batch=%d)", M, N, K, batch);
  SgemmArgs a{A, B, C, bias_m, M, N, K, am, ak, bk, bn, cm, cn, az, bz, cz, accumulate};
  if (N == 1 && batch == 1) {
    sgemv_strided_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return b2h::check_launch("sgemv_strided");
  }
  dim3 grid((N + SG_TN - 1) / SG_TN, (M + SG_TM - 1) / SG_TM, batch);
  B2_REQUIRE(grid.y <= 65535, "sgemm_strided: M=%d too large", M);
  sgemm_strided_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return b2h::check_launch("sgemm_strided");
}

int b200unet_sum_batches(const float* part, float* out, int64_t n, int batches, b200_stream_t stream) {
  B2_REQUIRE(part && out && n > 0 && batches > 0, "sum_batches: bad arguments");
  const long long blocks = (n + 255) / 256;
  sum_batches_kernel<<<static_cast<int>(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(part, out, n, batches);
  return b2h::check_launch("sum_batches");
}

}  // extern "C"
