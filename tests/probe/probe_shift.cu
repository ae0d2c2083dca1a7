// TEST-ONLY object (tests/probe/libb200probe.so; NOT part of libb200unet.so or include/b200unet.h).
// Hardware probe (bring-up / test aid, not on the product path): does a 128B-swizzled K-major UMMA operand work
// when its start address is offset by a whole number of 128-byte rows that is NOT a multiple of the 8-row
// swizzle atom, and when its 8-row groups are `sbo` bytes apart with sbo NOT a multiple of 1024 (the 16x8-pixel
// tile inside an 18x10 halo tile of conv3_res.cu uses sbo = 1280)?
// D[m = 8g + i][0:64] = A[shift + g * sbo/128 + i][0:64] * B[64 x 64]^T, A tile = 256 rows loaded by TMA.
#include "../../unet-torch_b200/csrc/host_common.h"
#include "../../unet-torch_b200/csrc/tc_common.cuh"

namespace {
using namespace b2;

__global__ void __launch_bounds__(128) probe_shift_kernel(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB, float* out, int shift,
                                                          int use_base_offset, int sbo) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + 256 * 128, bar = sB + 64 * 128, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 256 * 128 + 64 * 128 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = sA + shift * 128 + k * 32;
      uint64_t ad = umma_desc_sw128(a_addr, 16, sbo);
      if (use_base_offset) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      umma_bf16(tmem, ad, umma_desc_sw128(sB + k * 32, 16, 1024), idesc, k > 0);
    }
    umma_commit(bar2);
  }
  __syncwarp();
  mbar_wait(bar2, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + h * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * 64 + h * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}
}  // namespace

extern "C" int b200probe_shift(const void* a_256x64, const void* b_64x64, float* out_128x64, int shift,
                                    int use_base_offset, int sbo_bytes, void* stream) {
  CUtensorMap ta, tb;
  if (int e = b2h::make_tmap_2d(&ta, a_256x64, 64, 256, 256)) return e;
  if (int e = b2h::make_tmap_2d(&tb, b_64x64, 64, 64, 64)) return e;
  const int smem = 256 * 128 + 64 * 128 + 64 + 1024;
  probe_shift_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(ta, tb, out_128x64, shift, use_base_offset, sbo_bytes);
  return b2h::check_launch("probe_shift");
}
