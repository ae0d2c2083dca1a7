// Synchronised BatchNorm over NVLink / NVSwitch peer memory: one-shot all-reduce of the per-channel fp64 statistics
// fused with the BatchNorm finalisation, in ONE kernel per rank (SURVEY.md section 8e ii/iii; reference semantics:
// torch.nn.SyncBatchNorm over nn.BatchNorm2d, Model.py:17,21).
//
// A step issues 36 of these reductions (18 forward, 18 backward), each <= 16 KB and on the critical path, so they are
// latency-bound: a NCCL call costs ~20-30 us of stream time, this kernel a few us. Every rank owns a symmetric buffer
// (torch.distributed._symmetric_memory, mapped into all peers). Protocol for reduction number `seq` (slot = seq & 1):
//   1. write my vector into slot[seq&1][my_rank] of EVERY rank's buffer (plain stores to mapped peer memory),
//      __threadfence_system, then store-release `seq` into flag[slot][my_rank] of every rank;
//   2. spin (acquire loads, bounded) until my own flag[slot][r] == seq for all r;
//   3. sum the `world` vectors in rank order (bit-identical on all ranks), optionally finalise BatchNorm.
// Two slots suffice: a rank can only reach reduction seq+2 after every peer has published seq+1, i.e. after every peer
// has finished reading seq. Sequence numbers never repeat, so flags are never reset.
// seq = 0 in the call means "next": the kernel takes the number from a counter in its own buffer and advances it, so a
// launch captured in a CUDA graph stays valid on every replay (all ranks issue the same reductions in the same order, hence
// their counters agree). Explicit (host-side) and counter-based numbering must not be mixed on one buffer.
// A peer that never arrives (crashed rank; a rank stuck in a data loader for longer than the timeout): the wait is bounded
// by a configurable timeout (b200unet_nvl_set_timeout_ms, default 10 minutes like NCCL's watchdog, 0 = wait forever). On
// expiry the kernel does NOT trap (that would poison the CUDA context of every waiting rank): it records {sequence number,
// mask of missing ranks} in its own buffer, fills its outputs with NaN so that the failure is loud in the very next loss
// value, and returns. The host reads the record with b200unet_nvl_status (DataParallelContext.check_health()).
#include "../../include/b200unet.h"
#include "host_common.h"

namespace {

constexpr int MAX_WORLD = 8;
constexpr int SLOT_DOUBLES = 2048;                                   // 2 * Cmax
constexpr size_t FLAG_OFFSET = size_t(2) * MAX_WORLD * SLOT_DOUBLES * 8;  // bytes: data region first, then flags
constexpr size_t COUNTER_OFFSET = FLAG_OFFSET + 2 * MAX_WORLD * 8;  // this rank's own reduction counter (device-side seq)
constexpr size_t ERROR_OFFSET = COUNTER_OFFSET + 8;  // {failed sequence number, bit mask of ranks that never arrived}
constexpr size_t LEGACY_BYTES = COUNTER_OFFSET + 64;
// ---- second region: the multi-block "rows" reductions (one block per 32 channels, see nvl_rows_kernel)
constexpr int ROWS_MAX_BLOCKS = 64;    // 2048 channels / 32
constexpr int ROWS_BLOCK_DOUBLES = 64;  // {sum, sum of squares} x 32 channels
constexpr size_t ROWS_DATA_OFFSET = (LEGACY_BYTES + 127) / 128 * 128;
constexpr size_t ROWS_DATA_BYTES = size_t(2) * MAX_WORLD * ROWS_MAX_BLOCKS * ROWS_BLOCK_DOUBLES * 8;
constexpr size_t ROWS_FLAG_OFFSET = ROWS_DATA_OFFSET + ROWS_DATA_BYTES;          // [2][MAX_WORLD][ROWS_MAX_BLOCKS] u64
constexpr size_t ROWS_COUNTER_OFFSET = ROWS_FLAG_OFFSET + size_t(2) * MAX_WORLD * ROWS_MAX_BLOCKS * 8;  // [ROWS_MAX_BLOCKS] u64
constexpr size_t BUFFER_BYTES = ROWS_COUNTER_OFFSET + ROWS_MAX_BLOCKS * 8;
unsigned long long g_timeout_ns = 600ull * 1000000000ull;

struct PeerTable {
  unsigned char* buf[MAX_WORLD];
};

struct FinalizeArgs {
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* mean;
  float* rstd;
  float* scale;
  float* shift;
  double count;  // GLOBAL element count per channel
  float eps, momentum;
  int C;         // 0 = plain all-reduce
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) nvl_allreduce_kernel(const double* __restrict__ local, double* __restrict__ out, int n,
                                                           PeerTable peers, int world, int rank,
                                                           unsigned long long seq, FinalizeArgs fin,
                                                           unsigned long long timeout_ns) {
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(peers.buf[rank] + COUNTER_OFFSET);
  if (seq == 0) {  // device-side numbering (graph-replayable)
    __shared__ unsigned long long seq_sh;
    if (threadIdx.x == 0) {
      seq_sh = *counter + 1;
      *counter = seq_sh;  // only this kernel (stream-ordered, one block) touches the counter
    }
    __syncthreads();
    seq = seq_sh;
  }
  const int slot = static_cast<int>(seq & 1ull);
  const size_t my_off = (static_cast<size_t>(slot) * MAX_WORLD + rank) * SLOT_DOUBLES;
  __shared__ unsigned int missing_sh;
  if (threadIdx.x == 0)  // an earlier reduction already failed on this rank: do not wait another timeout per launch
    missing_sh = reinterpret_cast<const unsigned long long*>(peers.buf[rank] + ERROR_OFFSET)[0] != 0ull ? 0x80000000u : 0u;
  __syncthreads();
  const bool already_failed = missing_sh != 0u;
  // 1. publish (also after a failure: peers that are still healthy must not stall on THIS rank's account)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = local[i];
    for (int p = 0; p < world; ++p) reinterpret_cast<double*>(peers.buf[p])[my_off + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world)
    st_release_sys(reinterpret_cast<unsigned long long*>(peers.buf[threadIdx.x] + FLAG_OFFSET) + slot * MAX_WORLD + rank, seq);
  // 2. wait for every rank's vector of this sequence number (bounded, see the header comment)
  if (threadIdx.x < world && !already_failed) {
    const unsigned long long* f =
        reinterpret_cast<const unsigned long long*>(peers.buf[rank] + FLAG_OFFSET) + slot * MAX_WORLD + threadIdx.x;
    unsigned long long t0 = 0;
    unsigned int it = 0;
    while (ld_acquire_sys(f) != seq) {
      if ((++it & 0xfffu) == 0 && timeout_ns != 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) {
          atomicOr(&missing_sh, 1u << threadIdx.x);
          break;
        }
      }
    }
  }
  __syncthreads();
  if (missing_sh != 0u) {  // uniform: a peer never arrived. Record it, poison the outputs, keep the context alive.
    if (threadIdx.x == 0) {
      unsigned long long* err = reinterpret_cast<unsigned long long*>(peers.buf[rank] + ERROR_OFFSET);
      if (err[0] == 0ull) {
        err[0] = seq;
        err[1] = missing_sh;
      }
    }
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = qnan;
    const float fnan = __int_as_float(0x7fc00000);
    for (int c = threadIdx.x; c < fin.C; c += blockDim.x) fin.mean[c] = fin.rstd[c] = fin.scale[c] = fin.shift[c] = fnan;
    return;
  }
  // 3. reduce in rank order
  const double* mine = reinterpret_cast<const double*>(peers.buf[rank]) + static_cast<size_t>(slot) * MAX_WORLD * SLOT_DOUBLES;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_volatile_f64(mine + static_cast<size_t>(r) * SLOT_DOUBLES + i);
    out[i] = s;
  }
  if (fin.C > 0) {
    __syncthreads();  // out[] written by this block is read back below (same block: visible after the barrier)
    const int C = fin.C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double m = out[c] / fin.count;
      double var = out[C + c] / fin.count - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = static_cast<float>(m);
      const float rs = static_cast<float>(1.0 / sqrt(var + static_cast<double>(fin.eps)));
      fin.mean[c] = mf;
      fin.rstd[c] = rs;
      const float sc = fin.gamma[c] * rs;
      fin.scale[c] = sc;
      fin.shift[c] = fin.beta[c] - mf * sc;
      if (fin.running_mean != nullptr) {
        const double unbiased = fin.count > 1.0 ? var * fin.count / (fin.count - 1.0) : var;
        fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mf;
        fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * static_cast<float>(unbiased);
      }
    }
  }
}

// One kernel for "reduce the partial rows of this rank -> exchange with all ranks -> (finalise BatchNorm)":
//   partial [rows][2][C] fp32 (a conv epilogue's statistics rows, or the block partials of bn_bwd_reduce_kernel).
// grid = C / 32 blocks of 1024 threads. Block b owns channels [32b, 32b + 32): it reduces its 64 columns over the rows
// (32 row-lanes, fp64, fixed order), publishes the 64 doubles into slot[seq_b & 1][rank][b] of EVERY rank's buffer, raises
// flag[slot][rank][b] = seq_b everywhere, waits for all ranks' flags of the same (slot, b), sums the `world` vectors in rank
// order and writes local / global sums and - forward - the BatchNorm finalisation of its 32 channels. Every block keeps its
// OWN sequence counter (all ranks issue the same launches with the same grids, so block b's counters agree across ranks), so
// the launch is graph-replayable and no block waits for another block of its own grid. Compared with reduce_partials + the
// single-block nvl_allreduce_kernel this removes a launch from the critical path and spreads the work over C/32 SMs
// (measured at 2 ranks: 36 single-block reductions cost 0.86 ms per step, 24 us each).
struct RowsArgs {
  const float* partial;
  long long rows;
  int C;
  double* sums_local;   // [2C] or null
  double* sums_global;  // [2C] or null
  long long* num_batches;
  int finalize;
};

__global__ void __launch_bounds__(1024) nvl_rows_kernel(RowsArgs a, PeerTable peers, int world, int rank, FinalizeArgs fin,
                                                        unsigned long long timeout_ns) {
  __shared__ double sh[2][32][33];
  __shared__ double tot[ROWS_BLOCK_DOUBLES];
  __shared__ unsigned long long seq_sh;
  __shared__ unsigned int missing_sh;
  const int b = blockIdx.x, C = a.C;
  const int lane_c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = b * 32 + lane_c;
  unsigned char* mine = peers.buf[rank];
  if (threadIdx.x == 0) {
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(mine + ROWS_COUNTER_OFFSET) + b;
    seq_sh = *counter + 1;
    *counter = seq_sh;
    missing_sh = reinterpret_cast<const unsigned long long*>(mine + ERROR_OFFSET)[0] != 0ull ? 0x80000000u : 0u;
  }
  // 1. reduce this rank's rows
  double a1 = 0.0, a2 = 0.0;
  if (c < C) {
#pragma unroll 4
    for (long long r = rl; r < a.rows; r += 32) {
      a1 += static_cast<double>(__ldg(a.partial + r * 2 * C + c));
      a2 += static_cast<double>(__ldg(a.partial + r * 2 * C + C + c));
    }
  }
  sh[0][rl][lane_c] = a1;
  sh[1][rl][lane_c] = a2;
  __syncthreads();
  const unsigned long long seq = seq_sh;
  const bool already_failed = missing_sh != 0u;
  const int slot = static_cast<int>(seq & 1ull);
  const int t = threadIdx.x;  // t < 64: stat = t >> 5, channel lane = t & 31
  if (t < ROWS_BLOCK_DOUBLES) {
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v += sh[t >> 5][i][t & 31];
    const int cc = b * 32 + (t & 31);
    if (a.sums_local != nullptr && cc < C) a.sums_local[(t >> 5) * C + cc] = v;
    // 2. publish to every rank (plain stores to mapped peer memory), then make them visible system-wide
    const size_t off = ((static_cast<size_t>(slot) * MAX_WORLD + rank) * ROWS_MAX_BLOCKS + b) * ROWS_BLOCK_DOUBLES + t;
    for (int p = 0; p < world; ++p) reinterpret_cast<double*>(peers.buf[p] + ROWS_DATA_OFFSET)[off] = v;
    __threadfence_system();
  }
  __syncthreads();
  if (t < world)
    st_release_sys(reinterpret_cast<unsigned long long*>(peers.buf[t] + ROWS_FLAG_OFFSET) +
                       (static_cast<size_t>(slot) * MAX_WORLD + rank) * ROWS_MAX_BLOCKS + b, seq);
  // 3. wait for every rank's vector of this (block, sequence number)
  if (t < world && !already_failed) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + ROWS_FLAG_OFFSET) +
                                  (static_cast<size_t>(slot) * MAX_WORLD + t) * ROWS_MAX_BLOCKS + b;
    unsigned long long t0 = 0;
    unsigned int it = 0;
    while (ld_acquire_sys(f) != seq) {
      if ((++it & 0xfffu) == 0 && timeout_ns != 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) {
          atomicOr(&missing_sh, 1u << t);
          break;
        }
      }
    }
  }
  __syncthreads();
  const bool failed = missing_sh != 0u;
  if (failed && t == 0) {
    unsigned long long* err = reinterpret_cast<unsigned long long*>(mine + ERROR_OFFSET);
    if (atomicCAS(err, 0ull, seq) == 0ull) err[1] = missing_sh;
  }
  // 4. sum in rank order (bit-identical on all ranks); NaN after a failure so that it is loud
  if (t < ROWS_BLOCK_DOUBLES) {
    double v = 0.0;
    const double* d = reinterpret_cast<const double*>(mine + ROWS_DATA_OFFSET) +
                      (static_cast<size_t>(slot) * MAX_WORLD * ROWS_MAX_BLOCKS + b) * ROWS_BLOCK_DOUBLES + t;
    for (int r = 0; r < world; ++r) v += ld_volatile_f64(d + static_cast<size_t>(r) * ROWS_MAX_BLOCKS * ROWS_BLOCK_DOUBLES);
    if (failed) v = __longlong_as_double(0x7ff8000000000000ll);
    tot[t] = v;
    const int cc = b * 32 + (t & 31);
    if (a.sums_global != nullptr && cc < C) a.sums_global[(t >> 5) * C + cc] = v;
  }
  if (!a.finalize) return;
  __syncthreads();
  if (b == 0 && t == 0 && a.num_batches != nullptr) *a.num_batches += 1;
  if (t < 32 && c < C) {
    const double m = tot[t] / fin.count;
    double var = tot[32 + t] / fin.count - m * m;
    if (var < 0.0) var = 0.0;
    const float mf = static_cast<float>(m);
    const float rs = static_cast<float>(1.0 / sqrt(var + static_cast<double>(fin.eps)));
    fin.mean[c] = mf;
    fin.rstd[c] = rs;
    const float sc = fin.gamma[c] * rs;
    fin.scale[c] = sc;
    fin.shift[c] = fin.beta[c] - mf * sc;
    if (fin.running_mean != nullptr) {
      const double unbiased = fin.count > 1.0 ? var * fin.count / (fin.count - 1.0) : var;
      fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mf;
      fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * static_cast<float>(unbiased);
    }
  }
}

int fill_table(PeerTable* t, void* const* peer_bufs, int world, int rank, int n) {
  if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world) {
    b2h::set_error("nvl: world %d / rank %d out of range (max %d ranks)", world, rank, MAX_WORLD);
    return 1;
  }
  if (n < 1 || n > SLOT_DOUBLES) {
    b2h::set_error("nvl: vector length %d out of range (max %d doubles)", n, SLOT_DOUBLES);
    return 1;
  }
  for (int i = 0; i < MAX_WORLD; ++i) t->buf[i] = i < world ? static_cast<unsigned char*>(peer_bufs[i]) : nullptr;
  for (int i = 0; i < world; ++i)
    if (t->buf[i] == nullptr) {
      b2h::set_error("nvl: null peer buffer %d", i);
      return 1;
    }
  return 0;
}

}  // namespace

extern "C" {

int64_t b200unet_nvl_buffer_bytes(void) { return static_cast<int64_t>(BUFFER_BYTES); }

int b200unet_nvl_allreduce_f64(const double* local, double* out, int n, void* const* peer_bufs, int world, int rank,
                               int64_t seq, b200_stream_t stream) {
  PeerTable t;
  if (int e = fill_table(&t, peer_bufs, world, rank, n)) return e;
  B2_REQUIRE(seq >= 0, "nvl_allreduce_f64: sequence numbers start at 1 (0 = device-side counter)");
  FinalizeArgs fin{};
  fin.C = 0;
  nvl_allreduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(local, out, n, t, world, rank,
                                                                        static_cast<unsigned long long>(seq), fin, g_timeout_ns);
  return b2h::check_launch("nvl_allreduce_f64");
}

int b200unet_nvl_bn_sync_finalize(const double* local_sums, double* global_sums, void* const* peer_bufs, int world,
                                  int rank, int64_t seq, double global_count, const float* gamma, const float* beta,
                                  float eps, float momentum, float* running_mean, float* running_var, float* mean,
                                  float* rstd, float* scale, float* shift, int C, b200_stream_t stream) {
  PeerTable t;
  if (int e = fill_table(&t, peer_bufs, world, rank, 2 * C)) return e;
  B2_REQUIRE(seq >= 0 && C > 0, "nvl_bn_sync_finalize: bad arguments");
  FinalizeArgs fin{gamma, beta, running_mean, running_var, mean, rstd, scale, shift, global_count, eps, momentum, C};
  nvl_allreduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(local_sums, global_sums, 2 * C, t, world, rank,
                                                                        static_cast<unsigned long long>(seq), fin, g_timeout_ns);
  return b2h::check_launch("nvl_bn_sync_finalize");
}

int b200unet_nvl_rows_allreduce(const float* partial, int64_t rows, int C, void* const* peer_bufs, int world, int rank,
                                double* sums_local, double* sums_global, b200_stream_t stream) {
  PeerTable t;
  if (int e = fill_table(&t, peer_bufs, world, rank, 2)) return e;
  B2_REQUIRE(partial != nullptr && rows > 0 && C > 0 && C <= 32 * ROWS_MAX_BLOCKS, "nvl_rows_allreduce: C=%d out of range (max %d)",
             C, 32 * ROWS_MAX_BLOCKS);
  RowsArgs a{partial, rows, C, sums_local, sums_global, nullptr, 0};
  FinalizeArgs fin{};
  nvl_rows_kernel<<<(C + 31) / 32, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a, t, world, rank, fin, g_timeout_ns);
  return b2h::check_launch("nvl_rows_allreduce");
}

int b200unet_nvl_bn_rows_sync_finalize(const float* stats_partial, int64_t rows, int C, void* const* peer_bufs, int world,
                                       int rank, double global_count, const float* gamma, const float* beta, float eps,
                                       float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                       float* mean, float* rstd, float* scale, float* shift, b200_stream_t stream) {
  PeerTable t;
  if (int e = fill_table(&t, peer_bufs, world, rank, 2)) return e;
  B2_REQUIRE(stats_partial && gamma && beta && mean && rstd && scale && shift && rows > 0 && C > 0 && C <= 32 * ROWS_MAX_BLOCKS &&
                 global_count > 0, "nvl_bn_rows_sync_finalize: bad arguments (C=%d, max %d)", C, 32 * ROWS_MAX_BLOCKS);
  RowsArgs a{stats_partial, rows, C, nullptr, nullptr, reinterpret_cast<long long*>(num_batches_tracked), 1};
  FinalizeArgs fin{gamma, beta, running_mean, running_var, mean, rstd, scale, shift, global_count, eps, momentum, C};
  nvl_rows_kernel<<<(C + 31) / 32, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a, t, world, rank, fin, g_timeout_ns);
  return b2h::check_launch("nvl_bn_rows_sync_finalize");
}

int b200unet_nvl_set_timeout_ms(int64_t ms) {
  B2_REQUIRE(ms >= 0, "nvl_set_timeout_ms: negative timeout (0 = wait forever)");
  g_timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
  return 0;
}

int b200unet_nvl_status(const void* my_buffer, b200_stream_t stream, int64_t* out3) {
  // {reductions issued with the device-side counter, failed sequence number (0 = none), mask of ranks that never arrived}
  B2_REQUIRE(my_buffer != nullptr && out3 != nullptr, "nvl_status: null argument");
  unsigned long long h[3] = {0, 0, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(h, static_cast<const unsigned char*>(my_buffer) + COUNTER_OFFSET, sizeof(h),
                                  cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    b2h::set_error("nvl_status: %s", cudaGetErrorString(e));
    return 2;
  }
  for (int i = 0; i < 3; ++i) out3[i] = static_cast<int64_t>(h[i]);
  return 0;
}

}  // extern "C"
