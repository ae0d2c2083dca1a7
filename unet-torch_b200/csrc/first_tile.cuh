// inc.conv1 = nn.Conv2d(n_channels <= 7, 64, 3, padding=1, bias=False) (reference Model.py:15-16 via :111): its im2col row
//     col[pixel][c*9 + r*3 + s] = bf16(x[n, c, h + r - 1, w + s - 1])     (zero outside the image, zero for columns >= 9*Cin)
// is built DIRECTLY IN SHARED MEMORY by the GEMM kernels (conv3_res.cu RES_FIRST for the forward pass, wgrad.cu KIND_FIRST
// for the weight gradient) from the fp32 NCHW network input. The round-1 path wrote a 64-column im2col tensor to HBM
// (537 MB at 16 x 512^2) and read it back twice: ~1.6 GB of traffic for a layer whose input is 50 MB.
//
// A row is one pixel = 128 bytes = eight 16-byte chunks in the 128B-swizzle image the tensor core expects (the image
// cp.async.bulk.tensor with CU_TENSOR_MAP_SWIZZLE_128B would have produced): chunk j of row r lives at chunk (j ^ (r & 7)).
// Only the ceil(9*Cin/8) leading chunks carry data; the rest of every stage is zeroed once at kernel start.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace b2first {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int CIN>
struct Row {
  static constexpr int T = 9 * CIN;
  static constexpr int CHUNKS = (T + 7) / 8;
  float v[CHUNKS * 8];
};

// (Measured alternative: staging the (TH+2) x (TW+2) halo of a tile in shared memory with coalesced loads and gathering the
// taps from there was SLOWER - 0.42 ms vs 0.30 ms forward, 0.62 vs 0.25 ms weight gradient at 16 x 3 x 512^2: the builders
// are bound by the latency of their dependent chain (load -> barrier -> gather -> store -> fence -> arrive), not by L1
// wavefronts; what helps is more loads in flight: register double-buffering + L2 prefetch a few tiles ahead.)

// touch the input lines a later tile will need (thread t < 3 * CIN * ROWS: one row segment each), L2 only
template <int CIN, int TH, int TW>
__device__ __forceinline__ void prefetch_tile_l2(const float* __restrict__ x, long long img, int h0, int w0, int H, int W, int t) {
  if (t < CIN * (TH + 2)) {
    const int c = t / (TH + 2), h = h0 - 1 + (t - c * (TH + 2));
    if (h >= 0 && h < H && w0 < W) {
      const float* p = x + ((img * CIN + c) * static_cast<long long>(H) + h) * W + w0;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      if (TW > 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (w0 + TW < W ? TW : 0)));
    }
  }
}

// the 9*CIN taps of pixel (hh, ww) of image `img` (independent loads: issue them all, then wait once)
template <int CIN>
__device__ __forceinline__ void load_row(const float* __restrict__ x, long long img, int hh, int ww, int H, int W, Row<CIN>& row) {
#pragma unroll
  for (int j = Row<CIN>::T; j < Row<CIN>::CHUNKS * 8; ++j) row.v[j] = 0.f;
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    const float* plane = x + (img * CIN + c) * static_cast<long long>(H) * W;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = hh + r - 1;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = ww + s - 1;
        const bool in = (h >= 0) && (h < H) && (w >= 0) && (w < W);
        row.v[c * 9 + r * 3 + s] = in ? __ldg(plane + static_cast<long long>(h) * W + w) : 0.f;
      }
    }
  }
}

// store the row as bf16 into row `r` of a 128B-swizzled tile at shared address `tile` (1024-byte aligned)
template <int CIN>
__device__ __forceinline__ void store_row(uint32_t tile, int r, const Row<CIN>& row) {
#pragma unroll
  for (int j = 0; j < Row<CIN>::CHUNKS; ++j) {
    const uint32_t addr = tile + r * 128 + ((static_cast<uint32_t>(j) ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2(row.v[8 * j], row.v[8 * j + 1])),
                 "r"(pack2(row.v[8 * j + 2], row.v[8 * j + 3])), "r"(pack2(row.v[8 * j + 4], row.v[8 * j + 5])),
                 "r"(pack2(row.v[8 * j + 6], row.v[8 * j + 7]))
                 : "memory");
  }
}

// zero `bytes` of shared memory starting at `base` with all `nthreads` threads of the CTA (kernel prologue)
__device__ __forceinline__ void zero_smem(uint32_t base, int bytes, int tid, int nthreads) {
  for (int o = tid * 16; o < bytes; o += nthreads * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
}

}  // namespace b2first
