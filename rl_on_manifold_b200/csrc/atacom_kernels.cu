// libatacom_b200.so — kernels and C ABI (include/atacom_b200.h).  sm_100a only.
//
// Mapping: one thread owns one environment; a warp owns 32 consecutive environments, whose rows of
// the AoS [B, dim] arrays form one contiguous piece per array.  The step kernels run constraint
// evaluation (one FK pass) + the dual projection (atacom_dual.cuh) per thread, with the two big
// operands of the projection in a per-warp region of shared memory; rows move either directly
// (device-resident arrays) or through that region with the bulk-copy engine (mapped host arrays,
// peer buffers of the fused gather).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>

#include "../../include/atacom_b200.h"
#include "atacom_envs.cuh"

using namespace atacom;

static_assert(sizeof(ParamsT<float>) == sizeof(AtacomParams), "ParamsT<float> must mirror AtacomParams");

namespace {

#ifndef ATACOM_TPB
#define ATACOM_TPB 128   // environments (threads) per block of the auxiliary kernels
#endif
#ifndef ATACOM_STEP_MAX_TPB
#define ATACOM_STEP_MAX_TPB 448   // largest block of the step kernels: 448 x 144 registers fill one SM
#endif
#ifndef ATACOM_PHASE_BARRIERS
#define ATACOM_PHASE_BARRIERS 0   // 1: block barriers between the phases of the projection (measured: no gain)
#endif
#ifndef ATACOM_STEP_DUAL
#define ATACOM_STEP_DUAL 1        // 1: dual projection in FP64 (atacom_dual.cuh); 0: fp32 structured path
#endif
#ifndef ATACOM_STEP_STAGED_IO
// I/O of the step kernels.  0: every thread moves its own rows (best on HBM-resident arrays, by 2 %);
// 1: per-warp bulk copies (cp.async.bulk + mbarrier) through shared memory (best when the arrays are mapped
// host memory: large PCIe requests, +30 % end to end); 2 (default): two instantiations, chosen per launch.
#define ATACOM_STEP_STAGED_IO 2
#endif
#ifndef ATACOM_STEP_DEVICE_IO    // I/O mode of the device entry points: 0 (direct rows) or 3 (s, alpha in the background)
#define ATACOM_STEP_DEVICE_IO 0
#endif
#ifndef ATACOM_X_BULK_LOAD      // experiments: which half of the I/O the bulk instantiation moves with cp.async.bulk
#define ATACOM_X_BULK_LOAD 1
#endif
#ifndef ATACOM_X_BULK_STORE
#define ATACOM_X_BULK_STORE 1
#endif
#ifndef ATACOM_PDL_DEFAULT
#define ATACOM_PDL_DEFAULT 1
#endif
#ifndef ATACOM_STEP_MAXNREG
#define ATACOM_STEP_MAXNREG 128   // 448 threads = 14 warps, 4 on two of the SM sub-partitions: 16384 / (4 x 32) registers each
#endif
#ifndef ATACOM_SPIN_BUDGET
#define ATACOM_SPIN_BUDGET (4000000000ll)   // clock64 ticks (~2 s) a device-side wait may take before it gives up
#endif
constexpr int TPB = ATACOM_TPB;
constexpr int STEP_MAX_TPB = ATACOM_STEP_MAX_TPB;
constexpr int MAX_DEVICES = 64;
std::atomic<int64_t> g_launches{0};
// Device-side waits (ordered admission, in-kernel cross-rank barrier) are bounded: a wait that exceeds
// ATACOM_SPIN_BUDGET gives up, counts itself here and the kernel carries on (admission: unrationed loads, results
// unaffected; barrier: the step is not published to / awaited from the late rank).  atacom_spin_timeouts() reads it.
__device__ unsigned g_spin_timeouts = 0;

// NVTX range around a host-side section (step launch, gather launch, host pipeline); a no-op without a profiler
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
  return dev;
}

// ------------------------------------------------------------------ staging helpers
template <int DIM>
__device__ __forceinline__ void slab_load(const float* __restrict__ g, float* sm, int64_t env0, int nvalid) {
  const float* src = g + env0 * DIM;
  const int total = nvalid * DIM;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int nv = total >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(sm);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = __ldg(s4 + i);
    for (int i = (nv << 2) + threadIdx.x; i < total; i += blockDim.x) sm[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = __ldg(src + i);
  }
}

template <int DIM>
__device__ __forceinline__ void slab_store(float* __restrict__ g, const float* sm, int64_t env0, int nvalid) {
  float* dst = g + env0 * DIM;
  const int total = nvalid * DIM;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int nv = total >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(sm);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = s4[i];
    for (int i = (nv << 2) + threadIdx.x; i < total; i += blockDim.x) dst[i] = sm[i];
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = sm[i];
  }
}

template <int X> struct round4_t {
  static constexpr int value = (X + 3) & ~3;
  static __host__ __device__ constexpr int up(int v) { return (v + 3) & ~3; }
};

struct StepArgs {
  const float* q;
  const float* dq;
  const float* s_in;
  const float* alpha;
  float* ddq;
  float* s_out;
  uint8_t* status;
  float* w_dbg;
  int64_t B;
  // fused gather: every rank's [world * B, n] buffer (peer-mapped over NVLink); this rank's rows start at
  // gather_row0.  n_peers = 0 disables it.
  float* peer[ATACOM_MAX_PEERS];
  int64_t gather_row0;
  int32_t n_peers;
  int32_t aligned16;   // every array pointer (and peer buffer) is 16-byte aligned: bulk copies allowed
  // in-kernel cross-rank barrier of the fused gather (null / unused otherwise)
  uint32_t* peer_flags[ATACOM_MAX_PEERS];   // rank w's flag array [world]
  uint32_t* local_sync;                     // this rank: { blocks done, step sequence number }
  int32_t rank;
  // ordered admission of the bulk loads (IO mode 1 over mapped host memory): { warps loaded, warps finished,
  // next block ticket }, all zero between launches; at most gate_window warps have their loads in flight, in
  // ticket order, so the first blocks compute and store while the last ones are still waiting for PCIe
  int32_t gate_window;
  uint32_t* gate;
  // ATACOM_BASIS_LAPACK: the environments the dual path leaves to the LAPACK-basis routine (Dual::project,
  // band_defer).  Block b appends their indices to fix_list[b * blockDim.x ...] and stores how many in fix_count[b];
  // the fix-up kernel (same grid) then redoes exactly those.  Null: nothing is deferred (canonical basis).
  int32_t* fix_list;
  int32_t* fix_count;
  // ... and their minimum-norm parts ([B, N] doubles, written for the deferred environments only): the same for every
  // null basis, so the fix-up kernel redoes the null part alone
  double* fix_wmn;
  // zero-copy launch on the caller's (mapped host) arrays in that mode: every thread also writes the input rows it has
  // just fetched over PCIe to these device arrays, and the fix-up kernel reads the deferred environments' rows from
  // there instead of crossing PCIe a second time.  Null: no mirror (the arrays are in HBM already).
  float* mirror_q;
  float* mirror_dq;
  float* mirror_s;
  float* mirror_alpha;      // [B, k]
};

// ------------------------------------------------------------------ row access
// One thread moves its own rows.  A warp's 32 rows of a [B, DIM] array are one contiguous piece of
// 32*DIM*4 bytes, so its DIM loads touch exactly the lines of that piece (the first load brings them
// to L1, the others hit): DRAM traffic is the algorithmic minimum without a shared-memory round trip,
// and the loads of all arrays are in flight together.  Rows are 8-byte aligned when DIM is even.
template <int DIM>
__device__ __forceinline__ void row_load(const float* __restrict__ g, int64_t e, float* dst) {
  const float* src = g + e * DIM;
  if (DIM % 2 == 0 && (reinterpret_cast<uintptr_t>(g) & 7) == 0) {
#pragma unroll
    for (int j = 0; j < DIM / 2; ++j) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src) + j);
      dst[2 * j] = v.x;
      dst[2 * j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < DIM; ++j) dst[j] = __ldg(src + j);
  }
}

template <int DIM>
__device__ __forceinline__ void row_store(float* __restrict__ g, int64_t e, const float* src) {
  float* dst = g + e * DIM;
  if (DIM % 2 == 0 && (reinterpret_cast<uintptr_t>(g) & 7) == 0) {
#pragma unroll
    for (int j = 0; j < DIM / 2; ++j) reinterpret_cast<float2*>(dst)[j] = make_float2(src[2 * j], src[2 * j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < DIM; ++j) dst[j] = src[j];
  }
}

// ------------------------------------------------------------------ scratch of the dual projection
// Dynamic shared memory is cut into one region per warp.  A region is
//   * the staging area of the warp's I/O: its 32 rows of each [B, dim] array are one contiguous,
//     16-byte aligned piece per array, moved by the bulk-copy engine (cp.async.bulk, one elected lane,
//     completion on the warp's own mbarrier — no block barrier anywhere in the kernel), and
//   * for the iiwa kernels, the scratch of the projection: Y (m x n) and L (m (m+1) / 2) as
//     [entry][lane] arrays of doubles (conflict-free), 57 x 32 x 8 = 14.6 KB per warp, 204 KB per 448-thread
//     block — together with the register file that holds the working set of all 448 environments of
//     the SM at once.  Small environments keep Y and L in registers.
// The staging area aliases the scratch, which is idle at both ends of the kernel.
template <class Env>
struct StepScratch {
  using ED = typename Env::D;
  using DU = Dual<double, ED, Env::NDIAG>;
  static constexpr bool SHARED = DU::Y_SIZE + DU::L_SIZE > 24;
  static constexpr int MAX_WARPS = ATACOM_STEP_MAX_TPB / 32;
  static constexpr size_t SCRATCH = SHARED ? sizeof(double) * (DU::Y_SIZE + DU::L_SIZE) * 32 : 0;
  static constexpr size_t STAGE = sizeof(float) * 32 * (3 * ED::n + ED::G);   // q, dq, alpha, s
  static constexpr size_t WARP_BYTES = ((SCRATCH > STAGE ? SCRATCH : STAGE) + 127) / 128 * 128;
  // Layout: a fixed header — the block's ticket, (iiwa) the scratch slots of the warp-cooperative general
  // null-space routine (Dual::null_part_general_warp), each followed by its lock, one mbarrier per warp — then the
  // regions of the warps the block actually has: a launch with smaller blocks asks for less and more blocks fit an
  // SM.  As many slots as fit next to a full-size block (warp w uses slot w mod COOP_SLOTS): one per warp for
  // iiwa-6, two for iiwa-7; the 448-thread blocks take the 228 KB carve-out either way.
  static constexpr size_t SMEM_MAX = 227 * 1024;
  static constexpr size_t TICKET_OFFSET = 0;
  static constexpr size_t FIXN_OFFSET = 4;       // number of environments this block left to the fix-up kernel
  static constexpr size_t COOP_OFFSET = 16;
  static constexpr size_t COOP_SLOT_BYTES = (sizeof(double) * DU::COOP_DOUBLES + 16 + 15) / 16 * 16;
  static constexpr size_t COOP_SPARE = SMEM_MAX - 128 - COOP_OFFSET - sizeof(uint64_t) * MAX_WARPS - WARP_BYTES * MAX_WARPS;
#ifndef ATACOM_COOP_SLOTS_MAX
#define ATACOM_COOP_SLOTS_MAX (ATACOM_STEP_MAX_TPB / 32)
#endif
  static constexpr int COOP_SLOTS = !SHARED ? 0 : (COOP_SPARE / COOP_SLOT_BYTES >= size_t(ATACOM_COOP_SLOTS_MAX)
                                                   ? ATACOM_COOP_SLOTS_MAX : static_cast<int>(COOP_SPARE / COOP_SLOT_BYTES));
  static_assert(!SHARED || COOP_SLOTS >= 1, "no room for the cooperative scratch next to a full-size block");
  static constexpr size_t BAR_OFFSET = COOP_OFFSET + COOP_SLOT_BYTES * COOP_SLOTS;
  static constexpr size_t HEAD = (BAR_OFFSET + sizeof(uint64_t) * MAX_WARPS + 127) / 128 * 128;
  static constexpr size_t BYTES = HEAD + WARP_BYTES * MAX_WARPS;
  static_assert(BYTES <= SMEM_MAX, "step kernel scratch exceeds the shared memory of an SM");
  static constexpr size_t bytes(int tpb) { return HEAD + WARP_BYTES * static_cast<size_t>((tpb + 31) / 32); }
  static __device__ __forceinline__ unsigned char* region(unsigned char* smem, int warp) {
    return smem + HEAD + warp * WARP_BYTES;
  }
  static constexpr int COOP_STRIDE = static_cast<int>(COOP_SLOT_BYTES / sizeof(double));
  static __device__ __forceinline__ double* coop(unsigned char* smem) {     // slot 0
    return SHARED ? reinterpret_cast<double*>(smem + COOP_OFFSET) : nullptr;
  }
  static __device__ __forceinline__ void coop_init(unsigned char* smem) {   // one thread, before a block barrier
    for (int i = 0; i < COOP_SLOTS; ++i)
      *reinterpret_cast<volatile unsigned*>(smem + COOP_OFFSET + i * COOP_SLOT_BYTES + sizeof(double) * DU::COOP_DOUBLES) = 0u;
  }
  using YS = typename std::conditional<SHARED, SharedStore<double, 32>, LocalStore<double, DU::Y_SIZE>>::type;
  using LS = typename std::conditional<SHARED, SharedStore<double, 32>, LocalStore<double, DU::L_SIZE>>::type;
  static __device__ __forceinline__ YS y(unsigned char* region, int lane) {
    if constexpr (SHARED) return YS{reinterpret_cast<double*>(region) + lane};
    else return YS{};
  }
  static __device__ __forceinline__ LS l(unsigned char* region, int lane) {
    if constexpr (SHARED) return LS{reinterpret_cast<double*>(region) + DU::Y_SIZE * 32 + lane};
    else return LS{};
  }
};

// ---- bulk-copy engine and mbarrier (PTX ISA 8.x, sm_90+)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ AtacomEnvWrapper.step_action_function
// One thread = one environment; the block size is chosen at launch (<= STEP_MAX_TPB) so that the batch
// spreads evenly over the SMs in as few waves as possible — 65 536 environments are one block of 448 per SM.
// No block barrier anywhere on the main path: warps drift apart on purpose (while one waits for its loads
// another is in the FP64-heavy part of the projection).  Timing model and measurements: DESIGN.md §6.
template <class Env, int IO, bool BAND>   // IO 0: direct rows, 1: bulk loads and stores, 2: direct loads, bulk stores,
                                         // 3: q, dq direct; s, alpha bulk-copied in the background of the kinematics
                                         // BAND: ATACOM_BASIS_LAPACK — environments inside rref's tolerance band are
                                         // left to the fix-up kernel (the slack-pivot phases are not compiled in)
__global__ void __maxnreg__(ATACOM_STEP_MAXNREG) atacom_step_kernel(const __grid_constant__ StepArgs a,
                                                                    const __grid_constant__ ParamsT<float> P,
                                                                    const __grid_constant__ DualConsts<double> Kd) {
  using D = typename Env::D;
  constexpr int n = D::n, G = D::G, k = D::k, N = D::N;
  constexpr int G1 = at_least_1<G>::value, K1 = at_least_1<k>::value;
  extern __shared__ __align__(128) unsigned char atacom_smem[];
  using SC = StepScratch<Env>;
  // Block index.  With ordered admission the blocks number themselves in the order they start (a ticket), so a
  // warp only ever waits for warps of blocks that are already running — no assumption on the dispatch order.
  unsigned bid = blockIdx.x;
  const bool gated = IO == 1 && a.gate != nullptr;
  volatile uint32_t* tk = reinterpret_cast<volatile uint32_t*>(atacom_smem + SC::TICKET_OFFSET);
  // programmatic dependent launch (no-ops otherwise): let the next kernel's blocks queue up behind this one's,
  // and do not touch global memory before the previous kernel has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    SC::coop_init(atacom_smem);
    *reinterpret_cast<volatile uint32_t*>(atacom_smem + SC::FIXN_OFFSET) = 0u;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) {
    if (gated) *tk = atomicAdd(a.gate + 2, 1u);
  }
  if (SC::SHARED || gated || a.fix_list != nullptr) __syncthreads();     // before any work (measured free)
  if (gated) bid = *tk;
  const int64_t e_raw = static_cast<int64_t>(bid) * blockDim.x + threadIdx.x;
  const bool valid = e_raw < a.B;
  // threads past the end recompute the last environment and discard it, so that every thread of the
  // block reaches the phase barriers inside the projection
  const int64_t e = valid ? e_raw : a.B - 1;
  const bool ec = P.variant == VARIANT_EC;

  // Inputs.  A full warp whose slabs are 16-byte aligned has the bulk-copy engine fetch them into its region
  // and every lane picks its own rows up from there; any other warp (tail of the batch, unaligned views) has
  // each thread load its rows directly.
  float q[n], dq[n], s[G1], al[n], ddq[n], so[G1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* region = SC::region(atacom_smem, warp);
  const int64_t wenv0 = e_raw - lane;
  const int na = ec ? n : k;
  uint64_t* bar = reinterpret_cast<uint64_t*>(atacom_smem + SC::BAR_OFFSET) + warp;
  const bool bulk = IO == 1 && a.aligned16 && (wenv0 + 32 <= a.B);
  bool lazy = false;
  float* lz = reinterpret_cast<float*>(region + (SC::SHARED ? sizeof(double) * SC::DU::Y_SIZE * 32 : 0));
  float* sq = reinterpret_cast<float*>(region);
  float* sdq = sq + 32 * n;
  float* ss = sdq + 32 * n;
  float* sa = ss + 32 * G;
  if (bulk && ATACOM_X_BULK_LOAD) {
    if (lane == 0) {
      if (gated) {
        // admission: warp number `order` may load once fewer than gate_window warps ahead of it are outstanding
        const unsigned order = bid * (blockDim.x >> 5) + static_cast<unsigned>(warp);
        const unsigned window = static_cast<unsigned>(a.gate_window);
        if (order >= window) {
          unsigned seen;
          const long long t0 = clock64();
          for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.gate) : "memory");
            if (seen + window > order) break;
            if (clock64() - t0 > ATACOM_SPIN_BUDGET) {   // watchdog: load unrationed rather than hang
              atomicAdd(&g_spin_timeouts, 1u);
              break;
            }
            __nanosleep(64);
          }
        }
      }
      mbar_init(bar, 1);
      mbar_expect_tx(bar, 4u * 32u * static_cast<uint32_t>(2 * n + G + na));
      bulk_g2s(sq, a.q + wenv0 * n, 128u * n, bar);
      bulk_g2s(sdq, a.dq + wenv0 * n, 128u * n, bar);
      if (G > 0) bulk_g2s(ss, a.s_in + wenv0 * G, 128u * G, bar);
      if (na > 0) bulk_g2s(sa, a.alpha + wenv0 * na, 128u * static_cast<uint32_t>(na), bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (gated && lane == 0) atomicAdd(a.gate, 1u);     // this warp's inputs have landed: admit the next one
#pragma unroll
    for (int j = 0; j < n; ++j) {
      q[j] = sq[lane * n + j];
      dq[j] = sdq[lane * n + j];
      al[j] = j < na ? sa[lane * na + j] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < G; ++i) s[i] = ss[lane * G1 + i];
    __syncwarp();     // the region is scratch from here on
  } else {
    if (gated && lane == 0) atomicAdd(a.gate, 1u);     // direct loads (tail, unaligned views) are not rationed
    // s and alpha are not needed before the projection: in mode 3 the bulk-copy engine brings the warp's two
    // slabs into the (still idle) L part of its region while the kinematics run on q and dq
    if (IO == 3 && G > 0) {
      lazy = a.aligned16 && (wenv0 + 32 <= a.B);
      if (lazy) {
        if (lane == 0) {
          mbar_init(bar, 1);
          mbar_expect_tx(bar, 128u * static_cast<uint32_t>(G + na));
          bulk_g2s(lz, a.s_in + wenv0 * G, 128u * G, bar);
          if (na > 0) bulk_g2s(lz + 32 * G, a.alpha + wenv0 * na, 128u * static_cast<uint32_t>(na), bar);
        }
        __syncwarp();
      }
    }
    row_load<n>(a.q, e, q);
    row_load<n>(a.dq, e, dq);
    if (!lazy) {
      if (G > 0) row_load<G1>(a.s_in, e, s);
      if (ec) {
        row_load<n>(a.alpha, e, al);
      } else {
        float ak[K1];
        if (k > 0) row_load<K1>(a.alpha, e, ak);
#pragma unroll
        for (int j = 0; j < n; ++j) al[j] = j < k ? ak[j < k ? j : 0] : 0.f;
      }
    }
  }

  if (BAND && IO == 1 && a.mirror_q != nullptr && valid) {      // (see StepArgs: rows for the fix-up kernel)
    row_store<n>(a.mirror_q, e, q);
    row_store<n>(a.mirror_dq, e, dq);
    if (G > 0) row_store<G1>(a.mirror_s, e, s);
    if (k > 0) row_store<K1>(a.mirror_alpha, e, al);
  }
  float* dbg = (a.w_dbg && valid) ? a.w_dbg + e * (2 * N) : nullptr;
#if ATACOM_STEP_DUAL
  using DU = Dual<double, D, Env::NDIAG>;
  typename SC::YS Ys = SC::y(region, lane);
  typename SC::LS Ls = SC::l(region, lane);
  auto fetch = [&](float* s_row, float* a_row) {
    if (IO == 3 && lazy) {
      mbar_wait(bar, 0);
#pragma unroll
      for (int i = 0; i < G; ++i) s[i] = lz[lane * G1 + i];
#pragma unroll
      for (int j = 0; j < n; ++j) al[j] = j < na ? lz[32 * G + lane * na + j] : 0.f;
      __syncwarp();     // every lane has its rows: the L part of the region may be overwritten
    }
#pragma unroll
    for (int i = 0; i < G; ++i) s_row[i] = s[i];
#pragma unroll
    for (int j = 0; j < n; ++j) a_row[j] = al[j];
  };
  // the row of fix_wmn this thread's environment owns, worked out when (and only if) the environment is deferred
  struct WmnOut {
    double* base;
    const uint32_t* ticket;     // ordered admission: the block's ticket replaces blockIdx.x
    __device__ __forceinline__ double* row() const {
      if (base == nullptr) return nullptr;
      unsigned t;
      asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
      const unsigned b = ticket ? *reinterpret_cast<const volatile uint32_t*>(ticket) : blockIdx.x;
      return base + (static_cast<int64_t>(b) * blockDim.x + t) * N;
    }
  };
  const WmnOut wmn_out{BAND ? a.fix_wmn : nullptr,
                       (IO == 1 && a.gate != nullptr) ? reinterpret_cast<const uint32_t*>(atacom_smem + SC::TICKET_OFFSET) : nullptr};
  const uint8_t st = step_dual_lazy<Env, float, double, BAND>(P, Kd, Ys, Ls, q, dq, fetch, ddq, so, dbg,
                                                              SC::coop(atacom_smem), SC::COOP_SLOTS, SC::COOP_STRIDE,
                                                              wmn_out);
#else
  RawConstraints<float, double, D> R;
  Env::template eval<float, double>(P, q, dq, R);
  const uint8_t st = step_from_raw<float, double, D, Env::NDIAG, ATACOM_PHASE_BARRIERS != 0>(P, R, dq, s, al, ddq, so, dbg);
#endif
  // Outputs go back the same way: rows to the warp's region, slabs to HBM by the bulk-copy engine.  Thread and
  // row indices are recomputed here (a volatile read of %tid.x cannot be merged with the one at the top), so
  // that nothing of the prologue stays live in registers across the projection.
  unsigned tid2;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid2));
  const int lane2 = tid2 & 31;
  unsigned bid2 = blockIdx.x;
  if (IO == 1 && a.gate != nullptr) bid2 = *reinterpret_cast<volatile uint32_t*>(atacom_smem + SC::TICKET_OFFSET);
  const int64_t e2 = static_cast<int64_t>(bid2) * blockDim.x + tid2;
  const int64_t wenv2 = e2 - lane2;
  // left to the fix-up kernel (which writes ddq, s_out and status of these environments): queue the index and
  // leave the outputs alone — s_out may alias s_in, which that kernel still has to read
  const bool fix = (st & ST_LAPACK_PATH) != 0;
  if (fix && e2 < a.B) {
    const unsigned pos = atomicAdd(reinterpret_cast<unsigned*>(atacom_smem + SC::FIXN_OFFSET), 1u);
    a.fix_list[static_cast<int64_t>(bid2) * blockDim.x + pos] = static_cast<int32_t>(e2);
  }
  if (e2 < a.B && a.status && !fix) a.status[e2] = st;
  if ((IO == 1 || IO == 2) && ATACOM_X_BULK_STORE && a.aligned16 && (wenv2 + 32 <= a.B)) {
    if (fix) {      // the slab store below covers this row too: hand the slack row on unchanged, a placeholder for ddq
      if (G > 0) row_load<G1>(a.s_in, e2, so);
#pragma unroll
      for (int j = 0; j < n; ++j) ddq[j] = 0.f;
    }
    float* oq = reinterpret_cast<float*>(SC::region(atacom_smem, static_cast<int>(tid2 >> 5)));
    float* os = oq + 32 * n;
    __syncwarp();     // every lane is done with its scratch
#pragma unroll
    for (int j = 0; j < n; ++j) oq[lane2 * n + j] = ddq[j];
#pragma unroll
    for (int i = 0; i < G; ++i) os[lane2 * G1 + i] = so[i];
    fence_async_smem();
    __syncwarp();
    if (lane2 == 0) {
      if (a.ddq) bulk_s2g(a.ddq + wenv2 * n, oq, 128u * n);
      if (G > 0) bulk_s2g(a.s_out + wenv2 * G, os, 128u * G);
      // fused all-gather: this warp's slab of rows straight into every rank's gather buffer
      for (int w = 0; w < a.n_peers; ++w) bulk_s2g(a.peer[w] + (a.gather_row0 + wenv2) * n, oq, 128u * n);
      if (a.local_sync != nullptr) bulk_commit_wait_all();   // the step is published below: writes must have landed
      else bulk_commit_wait_read();
    }
  } else if (e2 < a.B && !fix) {
    if (a.ddq) row_store<n>(a.ddq, e2, ddq);
    if (G > 0) row_store<G1>(a.s_out, e2, so);
    for (int w = 0; w < a.n_peers; ++w) row_store<n>(a.peer[w], a.gather_row0 + e2, ddq);
  }
  if (a.fix_list != nullptr) {
    __syncthreads();                                   // every thread of the block has queued its environment
    if (tid2 == 0) a.fix_count[bid2] = static_cast<int32_t>(*reinterpret_cast<volatile uint32_t*>(atacom_smem + SC::FIXN_OFFSET));
  }
  // Ordered admission: the last warp of the launch (every ticket taken, every gate passed) zeroes the counters.
  if (IO == 1 && a.gate != nullptr && lane2 == 0) {
    const unsigned done = atomicAdd(a.gate + 1, 1u);
    if (done == gridDim.x * (blockDim.x >> 5) - 1u) {
      atomicExch(a.gate, 0u);
      atomicExch(a.gate + 2, 0u);
      atomicExch(a.gate + 1, 0u);
    }
  }
  // Fused gather, cross-rank barrier: the last block of this launch publishes the step to every rank and waits
  // for every rank's flag of the same step, so kernel completion on a rank means its gather buffer is complete.
  if (a.local_sync != nullptr) {
    __syncthreads();                                   // every thread of the block has issued its peer stores
    if (tid2 == 0) {
      __threadfence_system();
      const unsigned done = atomicAdd(a.local_sync, 1u);
      if (done == gridDim.x - 1) {
        atomicExch(a.local_sync, 0u);
        const unsigned seq = atomicAdd(a.local_sync + 1, 1u) + 1u;
        __threadfence_system();
        for (int w = 0; w < a.n_peers; ++w)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flags[w] + a.rank), "r"(seq) : "memory");
        const long long t0 = clock64();
        bool late = false;
        for (int w = 0; w < a.n_peers && !late; ++w) {
          unsigned seen;
          for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.peer_flags[a.rank] + w) : "memory");
            if (static_cast<int>(seen - seq) >= 0) break;
            if (clock64() - t0 > ATACOM_SPIN_BUDGET) {   // watchdog: a rank that skipped the launch must not hang the others
              atomicAdd(&g_spin_timeouts, 1u);
              late = true;
              break;
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ fused simulator sub-steps
// PyBullet-style environments call the hook once per simulator sub-step (env_base.py:161-165, 4 per agent step)
// while q and dq stay what the last step() set (atacom.py:111-112, 123-126): the K projections of one agent step
// share the kinematics and differ only through the slacks, which each call integrates (atacom.py:135).  One
// launch does all K: the functor runs once, its products (K-scaled dense Jacobian rows, diagonal rows, the
// slack-free part of the right-hand side) are parked in a caller-provided workspace ([WS][B] doubles, L2-resident
// at these sizes) and reloaded for sub-steps 2..K; the slacks are carried in double.  Without a workspace the
// functor is simply re-run.  ddq: [K, B, n], one row per sub-step (each goes to the simulator's inverse dynamics).
struct SubstepArgs {
  const float* q;
  const float* dq;
  const float* s_in;
  const float* alpha;
  float* ddq;
  float* s_out;
  uint8_t* status;
  double* ws;
  int64_t B;
  int32_t K;
};

template <class Env, bool PARK>
__global__ void __maxnreg__(ATACOM_STEP_MAXNREG) atacom_substeps_kernel(const __grid_constant__ SubstepArgs a,
                                                                        const __grid_constant__ ParamsT<float> P,
                                                                        const __grid_constant__ DualConsts<double> Kd) {
  using D = typename Env::D;
  using SC = StepScratch<Env>;
  using DU = typename SC::DU;
  constexpr int n = D::n, G = D::G, k = D::k, C = D::C, F = D::F, NDIAG = Env::NDIAG;
  constexpr int G1 = at_least_1<G>::value, K1 = at_least_1<k>::value, ND1 = at_least_1<NDIAG>::value;
  const int64_t e_raw = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = e_raw < a.B;
  const int64_t e = valid ? e_raw : a.B - 1;
  const bool ec = P.variant == VARIANT_EC;
  extern __shared__ __align__(128) unsigned char atacom_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* region = SC::region(atacom_smem, warp);
  typename SC::YS Ys = SC::y(region, lane);
  typename SC::LS Ls = SC::l(region, lane);
  if (SC::SHARED) {
    if (threadIdx.x == 0) SC::coop_init(atacom_smem);
    __syncthreads();
  }

  // what stays live across the sub-steps is kept small on purpose (slacks, dq, alpha): shared memory takes
  // 204 of the SM's 256 KB, so a register spill inside the loop would go all the way to L2
  float dq[n], al[n];
  double sh[G1], sn[G1];
  row_load<n>(a.dq, e, dq);
  {
    float s[G1];
    if (G > 0) row_load<G1>(a.s_in, e, s);
#pragma unroll
    for (int i = 0; i < G; ++i) sh[i] = s[i];
  }
  if (ec) {
    row_load<n>(a.alpha, e, al);
  } else {
    float ak[K1];
    if (k > 0) row_load<K1>(a.alpha, e, ak);
#pragma unroll
    for (int j = 0; j < n; ++j) al[j] = j < k ? ak[j < k ? j : 0] : 0.f;
  }
  double* wp = a.ws + e;          // this environment's column of the [WS][B] workspace (PARK only)
  uint8_t status = 0;
  for (int step = 0; step < a.K; ++step) {
    double dg[ND1], r[at_least_1<C>::value];
    if (PARK && step > 0) {
#pragma unroll
      for (int i = 0; i < DU::Y_SIZE; ++i) Ys.set(i, wp[static_cast<int64_t>(i) * a.B]);
#pragma unroll
      for (int j = 0; j < NDIAG; ++j) dg[j] = wp[static_cast<int64_t>(DU::Y_SIZE + j) * a.B];
#pragma unroll
      for (int i = 0; i < C; ++i) r[i] = wp[static_cast<int64_t>(DU::Y_SIZE + NDIAG + i) * a.B];
    } else {
      float q[n];
      row_load<n>(a.q, e, q);
      DualSink<float, double, D, NDIAG, typename SC::YS> sink(Kd, Ys);
      Env::template eval<float, double>(P, q, dq, sink);
#pragma unroll
      for (int j = 0; j < NDIAG; ++j) dg[j] = sink.dg[j];
#pragma unroll
      for (int i = 0; i < C; ++i) r[i] = sink.r[i];
      if (PARK && a.K > 1 && valid) {   // only this thread reads these entries back: no fence needed
#pragma unroll
        for (int i = 0; i < DU::Y_SIZE; ++i) wp[static_cast<int64_t>(i) * a.B] = Ys.get(i);
#pragma unroll
        for (int j = 0; j < NDIAG; ++j) wp[static_cast<int64_t>(DU::Y_SIZE + j) * a.B] = dg[j];
#pragma unroll
        for (int i = 0; i < C; ++i) wp[static_cast<int64_t>(DU::Y_SIZE + NDIAG + i) * a.B] = r[i];
      }
    }
#pragma unroll
    for (int i = F; i < C; ++i) r[i] += 0.5 * Kd.K_c[i] * sh[i - F] * sh[i - F];      // atacom.py:195
    float ddq[n];
    const uint8_t st = dual_tail<Env, float, double>(P, Kd, Ys, Ls, dg, r, sh, al, dq, ddq, sn, nullptr,
                                                     SC::coop(atacom_smem), SC::COOP_SLOTS, SC::COOP_STRIDE);
    status |= st;
#pragma unroll
    for (int i = 0; i < G; ++i) sh[i] = sn[i];
    if (valid) row_store<n>(a.ddq + static_cast<int64_t>(step) * a.B * n, e, ddq);
  }
  if (valid) {
    float so[G1];
#pragma unroll
    for (int i = 0; i < G; ++i) so[i] = static_cast<float>(sh[i]);
    if (G > 0) row_store<G1>(a.s_out, e, so);
    if (a.status) a.status[e] = status;
  }
}

// ------------------------------------------------------------------ fix-up: the LAPACK-basis routine on the deferred
// environments (ATACOM_BASIS_LAPACK).  Block b works through segment b of the list the step kernel wrote — ~21 % of
// the batch at the benchmark's mix of active constraints, compacted per block — one thread per environment, the
// C x N working array of atacom_lapack.cuh as a column of a [cell][thread] shared-memory array.
struct FixArgs {
  const float* q;
  const float* dq;
  const float* s_in;
  const float* alpha;
  float* ddq;
  float* s_out;
  uint8_t* status;
  float* w_dbg;
  const int32_t* fix_list;
  const int32_t* fix_count;
  const double* fix_wmn;    // [B, N]: the minimum-norm parts the step kernel handed on
  int32_t seg_stride;       // block size of the step kernel = length of a segment
  float* peer[ATACOM_MAX_PEERS];
  int64_t gather_row0;
  int32_t n_peers;
};

#ifndef ATACOM_FIX_LPE
#define ATACOM_FIX_LPE 2
#endif
#ifndef ATACOM_FIX_NULL_ONLY
#define ATACOM_FIX_NULL_ONLY 1     // 0: the fix-up kernel redoes the minimum-norm part too (A/B builds)
#endif
// Lanes of a warp per deferred environment (2, 3 or 4).  Measured: 2 lanes 51.5 us per step,
// 3 lanes (10 environments per warp, two lanes left over) 54.4 us, 4 lanes 61.5 us — the sweeps get shorter with more
// lanes, but everything that is not a sweep (kinematics, forming the reflectors, the rref) is repeated by every warp
// and the kernel pays for the warp-instructions it issues as well as for the length of one environment's chain.
constexpr int FIX_LPE_WANTED = ATACOM_FIX_LPE;
static_assert(FIX_LPE_WANTED >= 2 && FIX_LPE_WANTED <= 4, "lanes per environment of the fix-up kernel");

// The lanes of a warp that share one environment in the fix-up kernel (Lapack::project, GRP).  The groups of a warp
// synchronise TOGETHER, with the full mask: a group-sized mask makes the compiler wrap every __syncwarp in a
// MATCH / vote loop (~15 instructions at each of the 31 synchronisation points of the routine), and the groups run the
// same instruction stream anyway.  Every lane of a live warp therefore takes part from start to end — slots past the
// end of the list redo the last environment and store nothing — and the routine passes the same synchronisation
// points whatever an environment's data decide.
struct WarpLanes {
  int s;
  __device__ __forceinline__ int sub() const { return s; }
  __device__ __forceinline__ int write_lane() const { return s; }
  // sum over the LPE (2 or 4, adjacent) lanes of the group: a butterfly, the same bits in every lane
  template <int LPE, typename R> __device__ __forceinline__ R sum(R x) const {
#pragma unroll
    for (int o = LPE / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// Block shape of the fix-up kernel for LPE lanes per environment.  A warp holds EPW = 32 / LPE environments (when 32
// is not a multiple of LPE the lanes left over shadow the warp's first environment).
// The [cell][column] array is padded to a stride = TARGET (mod 16) doubles, chosen so that the 16 threads of a
// half-warp — lanes of one environment are 17 cells (rows) or 1 cell (columns) apart — hit 16 different pairs of banks.
template <class Env, int LPE>
struct FixShape {
  using LP = Lapack<double, typename Env::D>;
  static constexpr size_t SMEM_MAX = 225 * 1024;           // (2 KB short of the limit: the block's small static tables)
  static constexpr int EPW = 32 / LPE;
  static constexpr int CPW = EPW;                          // shared-memory columns per warp
  static constexpr int TARGET = LPE == 2 ? 8 : LPE == 3 ? 11 : 4;
  static constexpr int WARPS_WANTED = LPE == 2 ? 8 : LPE == 3 ? 10 : 16;
  static constexpr int FIT = static_cast<int>(SMEM_MAX / (sizeof(double) * LP::SIZE));     // columns that fit
  static constexpr int stride_for(int cols) { return cols + ((TARGET - cols % 16) + 16) % 16; }
  static constexpr int warps_that_fit() {
    int w = WARPS_WANTED;
    while (w > 1 && stride_for(w * CPW) > FIT) --w;
    return w;
  }
  static constexpr int WARPS = warps_that_fit();
  static constexpr int ENVS = WARPS * EPW;                 // environments per block and round
  static constexpr int STRIDE = stride_for(WARPS * CPW);
  static constexpr int TPB = WARPS * 32;
  static constexpr size_t BYTES = sizeof(double) * LP::SIZE * STRIDE;
  static constexpr bool FITS = STRIDE <= FIT;
};
// The wanted number of lanes if the block then still takes the ~92 environments per SM of the benchmark's mix in one
// round, else 2 lanes (iiwa-7: 96 environments per block).
template <class Env>
struct FixCfg : FixShape<Env, (FixShape<Env, FIX_LPE_WANTED>::WARPS == FixShape<Env, FIX_LPE_WANTED>::WARPS_WANTED) ? FIX_LPE_WANTED : 2> {
  static constexpr int LPE = (FixShape<Env, FIX_LPE_WANTED>::WARPS == FixShape<Env, FIX_LPE_WANTED>::WARPS_WANTED) ? FIX_LPE_WANTED : 2;
  using SH = FixShape<Env, LPE>;
  static_assert(SH::FITS && SH::ENVS >= 10, "the LAPACK working arrays do not fit into shared memory");
};

template <class Env>
__global__ void __launch_bounds__(FixCfg<Env>::TPB) atacom_fix_kernel(const __grid_constant__ FixArgs a,
                                                                      const __grid_constant__ ParamsT<float> P,
                                                                      const __grid_constant__ DualConsts<double> Kd) {
  using D = typename Env::D;
  using CFG = FixCfg<Env>;
  constexpr int n = D::n, G = D::G, k = D::k, N = D::N;
  constexpr int G1 = at_least_1<G>::value, K1 = at_least_1<k>::value;
  extern __shared__ __align__(128) unsigned char atacom_smem[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");      // the step kernel has completed: list, counts and its outputs are visible
  // The step kernel's block b queued its deferred environments in segment b.  The segments differ in length (a
  // binomial spread, 70..115 of 448 at the benchmark's mix) and the duration of this kernel is that of its fullest
  // block, so the environments are dealt out evenly instead: the concatenation of all segments is cut into gridDim.x
  // equal ranges.  Warp 0 sums the counts, then walks them once more and notes the segments that overlap this block's
  // range [lo, hi) (a handful) in a small table; every slot then finds its environment there.
  constexpr int TABN = FixCfg<Env>::ENVS + 2;
  constexpr int CNTS = 1024;                             // counts kept in shared memory (more segments: read again)
  __shared__ int tab_seg[TABN], tab_beg[TABN], cnts[CNTS];
  __shared__ int tab_n, range_lo, range_hi;
  const int nseg = static_cast<int>(gridDim.x);
  for (int b = static_cast<int>(threadIdx.x); b < nseg && b < CNTS; b += static_cast<int>(blockDim.x)) cnts[b] = a.fix_count[b];
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = static_cast<int>(threadIdx.x);
    auto count_of = [&](int b) { return b < CNTS ? cnts[b] : a.fix_count[b]; };
    int sum = 0;
    for (int b = lane; b < nseg; b += 32) sum += count_of(b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int per = (sum + nseg - 1) / nseg;
    const int64_t lo64 = static_cast<int64_t>(blockIdx.x) * per;
    const int lo = lo64 < sum ? static_cast<int>(lo64) : sum;
    const int hi = lo + per < sum ? lo + per : sum;
    int run = 0, ntab = 0;
    for (int base = 0; base < nseg && run < hi; base += 32) {
      const int c = base + lane < nseg ? count_of(base + lane) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const int beg = run + incl - c;
      const bool ov = c > 0 && beg < hi && beg + c > lo;
      const unsigned m = __ballot_sync(0xffffffffu, ov);
      if (ov) {
        const int idx = ntab + __popc(m & ((1u << lane) - 1u));
        if (idx < TABN) {
          tab_seg[idx] = base + lane;
          tab_beg[idx] = beg;
        }
      }
      ntab += __popc(m);
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      tab_n = ntab < TABN ? ntab : TABN;
      range_lo = lo;
      range_hi = hi;
    }
  }
  __syncthreads();
  const int cnt = range_hi - range_lo;
  constexpr int LPE = CFG::LPE;
  const int warp = static_cast<int>(threadIdx.x >> 5), lane = static_cast<int>(threadIdx.x & 31u);
  // lanes left over (32 is not a multiple of LPE) shadow the warp's first environment: same column, same loads, same
  // decisions — and a lane number no row, column or vector is dealt to, so they never write
  const bool spare = lane / LPE >= CFG::EPW;
  const int grp_in_warp = spare ? 0 : lane / LPE, sub = spare ? 64 : lane % LPE;
  const int slot = warp * CFG::EPW + grp_in_warp;
  PlainSharedStore<double, CFG::STRIDE> S{reinterpret_cast<double*>(atacom_smem) + slot};
  const bool ec = P.variant == VARIANT_EC;
  const int warp_slot0 = warp * CFG::EPW;              // first slot of this warp
  for (int t0 = 0; t0 < cnt; t0 += CFG::ENVS) {
    if (t0 + warp_slot0 >= cnt) continue;              // (the whole warp)
    const bool in_range = t0 + slot < cnt;             // a slot past the end runs along on the last environment
    const bool live = !spare && in_range;
    const int gi = range_lo + (in_range ? t0 + slot : cnt - 1);
    int ts = tab_seg[0], tb = tab_beg[0];
    for (int u = 1; u < tab_n; ++u) {
      if (tab_beg[u] <= gi) {
        ts = tab_seg[u];
        tb = tab_beg[u];
      }
    }
    const int64_t e = a.fix_list[static_cast<int64_t>(ts) * a.seg_stride + (gi - tb)];
    float q[n], dq[n], s[G1], al[n], ddq[n], so[G1];
    row_load<n>(a.q, e, q);
    row_load<n>(a.dq, e, dq);
    if (G > 0) row_load<G1>(a.s_in, e, s);
    if (ec) {
      row_load<n>(a.alpha, e, al);
    } else {
      float ak[K1];
      if (k > 0) row_load<K1>(a.alpha, e, ak);
#pragma unroll
      for (int j = 0; j < n; ++j) al[j] = j < k ? ak[j < k ? j : 0] : 0.f;
    }
    float* dbg = (a.w_dbg && sub == 0 && live) ? a.w_dbg + e * (2 * N) : nullptr;
    const WarpLanes grp{sub};
#if ATACOM_FIX_NULL_ONLY
    // the minimum-norm part is the step kernel's (basis-free); only the null part is redone here
    const uint8_t st = step_lapack_null<Env, float, double, LPE>(P, Kd, S, q, dq, s, al, a.fix_wmn + e * N, ddq, so,
                                                                     dbg, grp);
#else
    const uint8_t st = step_lapack<Env, float, double, LPE>(P, Kd, S, q, dq, s, al, ddq, so, dbg, grp);
#endif
    if (sub == 0 && live) {
      if (a.status) a.status[e] = st;
      if (a.ddq) row_store<n>(a.ddq, e, ddq);
      if (G > 0) row_store<G1>(a.s_out, e, so);
      for (int w = 0; w < a.n_peers; ++w) row_store<n>(a.peer[w], a.gather_row0 + e, ddq);
    }
    __syncwarp();          // the next round reuses the group's array
  }
}

// ------------------------------------------------------------------ zero-copy step, both passes in one launch
// The host-buffer entry points on page-locked arrays (ATACOM_HOST_ZERO_COPY) with the reference-exact null basis.  Run
// as the two launches above on mapped host memory, the fix-up kernel writes its 20 % of the rows back one by one — 24
// and 44 byte stores across PCIe — and costs 140 us instead of 30 (324 against 185 us per 65 536-environment step).
// Here a block of 128 threads takes 128 environments through BOTH passes: bulk loads into shared memory in ticket
// order (the admission window of the step kernel), the dual path for every environment, then the LAPACK-basis routine
// for the ones it deferred — their inputs are still in the staging area, their minimum-norm parts next to it — and
// only then the output slabs leave, whole, through the bulk-copy engine.  No list, no scratch in global memory, no
// second trip over PCIe.  The launch is PCIe-bound; two such blocks fit an SM (iiwa-6).
struct ZcArgs {
  const float* q;
  const float* dq;
  const float* s_in;
  const float* alpha;
  float* ddq;
  float* s_out;
  uint8_t* status;
  int64_t B;
  int32_t aligned16;
  int32_t gate_window;
  uint32_t* gate;           // { warps loaded, warps finished, next block ticket }, zero between launches
};

template <class Env>
struct ZcFused {
  using D = typename Env::D;
  using SC = StepScratch<Env>;
  using LP = Lapack<double, D>;
#ifndef ATACOM_ZC_LPE
#define ATACOM_ZC_LPE 4
#endif
  // lanes per deferred environment: the block has registers to spare (128 threads), so 4 lanes — the whole block on
  // a round of 32 environments — cost nothing here, unlike in the fix-up kernel, and shorten the tail of the launch
  static constexpr int TPB = 128, WARPS = TPB / 32, LPE = ATACOM_ZC_LPE;
  static constexpr int SLOTS = 32;                      // deferred environments per round
  static constexpr int STRIDE = SLOTS + (LPE == 2 ? 8 : 4);   // = 8 / 4 (mod 16): see FixShape
  static constexpr int NA = D::k;
  static constexpr size_t align128(size_t x) { return (x + 127) / 128 * 128; }
  static constexpr size_t IN_WARP = align128(sizeof(float) * 32 * (2 * D::n + D::G + NA));
  static constexpr size_t OUT_WARP = align128(sizeof(float) * 32 * (D::n + D::G));
  static constexpr size_t HEAD = 128 + align128(sizeof(int32_t) * TPB);       // ticket, count, 4 mbarriers; the list
  static constexpr size_t IN_OFF = HEAD;
  static constexpr size_t OUT_OFF = IN_OFF + IN_WARP * WARPS;
  static constexpr size_t ST_OFF = OUT_OFF + OUT_WARP * WARPS;
  static constexpr size_t WMN_OFF = ST_OFF + 128;
  static constexpr size_t WORK_OFF = WMN_OFF + align128(sizeof(double) * TPB * D::N);
  static constexpr size_t WORK_DUAL = SC::WARP_BYTES * WARPS;
  static constexpr size_t WORK_LAPACK = sizeof(double) * LP::SIZE * STRIDE;
  static constexpr size_t BYTES = WORK_OFF + (WORK_DUAL > WORK_LAPACK ? WORK_DUAL : WORK_LAPACK);
  static_assert(BYTES <= 227 * 1024, "fused zero-copy block does not fit into shared memory");
};

template <class Env>
__global__ void __launch_bounds__(ZcFused<Env>::TPB) atacom_zc_fused_kernel(const __grid_constant__ ZcArgs a,
                                                                           const __grid_constant__ ParamsT<float> P,
                                                                           const __grid_constant__ DualConsts<double> Kd) {
  using D = typename Env::D;
  using Z = ZcFused<Env>;
  using SC = StepScratch<Env>;
  constexpr int n = D::n, G = D::G, k = D::k, N = D::N;
  constexpr int G1 = at_least_1<G>::value, K1 = at_least_1<k>::value;
  extern __shared__ __align__(128) unsigned char atacom_smem[];
  volatile uint32_t* tk = reinterpret_cast<volatile uint32_t*>(atacom_smem);
  uint32_t* fixn = reinterpret_cast<uint32_t*>(atacom_smem + 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(atacom_smem + 64);
  int32_t* list = reinterpret_cast<int32_t*>(atacom_smem + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool gated = a.gate != nullptr;
  if (threadIdx.x == 0) {
    *fixn = 0u;
    *tk = gated ? atomicAdd(a.gate + 2, 1u) : blockIdx.x;
  }
  __syncthreads();
  const unsigned bid = *tk;
  const int64_t e_raw = static_cast<int64_t>(bid) * Z::TPB + threadIdx.x;
  const bool valid = e_raw < a.B;
  const int64_t e = valid ? e_raw : a.B - 1;
  const int64_t wenv0 = e_raw - lane;
  float* sq = reinterpret_cast<float*>(atacom_smem + Z::IN_OFF + warp * Z::IN_WARP);
  float* sdq = sq + 32 * n;
  float* ss = sdq + 32 * n;
  float* sa = ss + 32 * G;
  float* oq = reinterpret_cast<float*>(atacom_smem + Z::OUT_OFF + warp * Z::OUT_WARP);
  float* os = oq + 32 * n;
  uint8_t* ost = atacom_smem + Z::ST_OFF;
  double* wmn = reinterpret_cast<double*>(atacom_smem + Z::WMN_OFF);
  const bool bulk = a.aligned16 && (wenv0 + 32 <= a.B);

  // ---- inputs: the warp's four slabs, admitted in ticket order (see atacom_step_kernel)
  float q[n], dq[n], s[G1], al[n], ddq[n], so[G1];
  if (bulk) {
    uint64_t* bar = bars + warp;
    if (lane == 0) {
      if (gated) {
        const unsigned order = bid * Z::WARPS + static_cast<unsigned>(warp);
        const unsigned window = static_cast<unsigned>(a.gate_window);
        if (order >= window) {
          unsigned seen;
          const long long t0 = clock64();
          for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.gate) : "memory");
            if (seen + window > order) break;
            if (clock64() - t0 > ATACOM_SPIN_BUDGET) {   // watchdog: load unrationed rather than hang
              atomicAdd(&g_spin_timeouts, 1u);
              break;
            }
            __nanosleep(64);
          }
        }
      }
      mbar_init(bar, 1);
      mbar_expect_tx(bar, 4u * 32u * static_cast<uint32_t>(2 * n + G + k));
      bulk_g2s(sq, a.q + wenv0 * n, 128u * n, bar);
      bulk_g2s(sdq, a.dq + wenv0 * n, 128u * n, bar);
      if (G > 0) bulk_g2s(ss, a.s_in + wenv0 * G, 128u * G, bar);
      if (k > 0) bulk_g2s(sa, a.alpha + wenv0 * k, 128u * static_cast<uint32_t>(k), bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (gated && lane == 0) atomicAdd(a.gate, 1u);     // this warp's inputs have landed: admit the next one
  } else {
    if (gated && lane == 0) atomicAdd(a.gate, 1u);     // direct loads (tail, unaligned views) are not rationed
    float ak[K1];
    row_load<n>(a.q, e, q);
    row_load<n>(a.dq, e, dq);
    if (G > 0) row_load<G1>(a.s_in, e, s);
    if (k > 0) row_load<K1>(a.alpha, e, ak);
#pragma unroll
    for (int j = 0; j < n; ++j) {
      sq[lane * n + j] = q[j];
      sdq[lane * n + j] = dq[j];
    }
#pragma unroll
    for (int i = 0; i < G; ++i) ss[lane * G1 + i] = s[i];
#pragma unroll
    for (int j = 0; j < k; ++j) sa[lane * k + j] = ak[j];
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < n; ++j) {
    q[j] = sq[lane * n + j];
    dq[j] = sdq[lane * n + j];
    al[j] = j < k ? sa[lane * k + (j < k ? j : 0)] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < G; ++i) s[i] = ss[lane * G1 + i];

  // ---- first pass: the dual path; a deferred environment leaves its minimum-norm part in wmn[thread]
  {
    unsigned char* region = atacom_smem + Z::WORK_OFF + warp * SC::WARP_BYTES;
    typename SC::YS Ys = SC::y(region, lane);
    typename SC::LS Ls = SC::l(region, lane);
    auto fetch = [&](float* s_row, float* a_row) {
#pragma unroll
      for (int i = 0; i < G; ++i) s_row[i] = s[i];
#pragma unroll
      for (int j = 0; j < n; ++j) a_row[j] = al[j];
    };
    const uint8_t st = step_dual_lazy<Env, float, double, true>(P, Kd, Ys, Ls, q, dq, fetch, ddq, so, nullptr, nullptr, 1, 0,
                                                                WmnRow<double>{wmn + threadIdx.x * N});
    if (st & ST_LAPACK_PATH) {
      if (valid) list[atomicAdd(fixn, 1u)] = static_cast<int32_t>(threadIdx.x);
    } else {
#pragma unroll
      for (int j = 0; j < n; ++j) oq[lane * n + j] = ddq[j];
#pragma unroll
      for (int i = 0; i < G; ++i) os[lane * G1 + i] = so[i];
      ost[threadIdx.x] = st;
    }
  }
  __syncthreads();     // every environment of the block is done or queued; the dual path's scratch is free

  // ---- second pass: the null part of the queued environments with the LAPACK basis, LPE lanes each
  const int cnt = static_cast<int>(*reinterpret_cast<volatile uint32_t*>(fixn));
  if (warp < Z::LPE * Z::SLOTS / 32) {
    const int slot = static_cast<int>(threadIdx.x) / Z::LPE, sub = static_cast<int>(threadIdx.x) % Z::LPE;
    PlainSharedStore<double, Z::STRIDE> S{reinterpret_cast<double*>(atacom_smem + Z::WORK_OFF) + slot};
    const int warp_slot0 = warp * (32 / Z::LPE);
    for (int t0 = 0; t0 < cnt; t0 += Z::SLOTS) {
      if (t0 + warp_slot0 >= cnt) continue;            // (the whole warp)
      const bool in_range = t0 + slot < cnt;           // a slot past the end runs along on the last environment
      const int t = list[in_range ? t0 + slot : cnt - 1];
      const int tw = t >> 5, tl = t & 31;
      const float* iq = reinterpret_cast<const float*>(atacom_smem + Z::IN_OFF + tw * Z::IN_WARP);
      const float* idq = iq + 32 * n;
      const float* is = idq + 32 * n;
      const float* ia = is + 32 * G;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        q[j] = iq[tl * n + j];
        dq[j] = idq[tl * n + j];
        al[j] = j < k ? ia[tl * k + (j < k ? j : 0)] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < G; ++i) s[i] = is[tl * G1 + i];
      const WarpLanes grp{sub};
      const uint8_t st = step_lapack_null<Env, float, double, Z::LPE>(P, Kd, S, q, dq, s, al, wmn + t * N, ddq, so, nullptr,
                                                                      grp);
      if (sub == 0 && in_range) {
        float* tq = reinterpret_cast<float*>(atacom_smem + Z::OUT_OFF + tw * Z::OUT_WARP);
        float* ts = tq + 32 * n;
#pragma unroll
        for (int j = 0; j < n; ++j) tq[tl * n + j] = ddq[j];
#pragma unroll
        for (int i = 0; i < G; ++i) ts[tl * G1 + i] = so[i];
        ost[t] = st;
      }
      __syncwarp();        // the next round reuses the group's array
    }
  }
  __syncthreads();

  // ---- outputs: whole slabs through the bulk-copy engine (rows of the tail / unaligned views one by one)
  if (bulk) {
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (a.ddq) bulk_s2g(a.ddq + wenv0 * n, oq, 128u * n);
      if (G > 0) bulk_s2g(a.s_out + wenv0 * G, os, 128u * G);
      if (a.status) bulk_s2g(a.status + wenv0, ost + warp * 32, 32u);
      bulk_commit_wait_read();
    }
  } else if (valid) {
#pragma unroll
    for (int j = 0; j < n; ++j) ddq[j] = oq[lane * n + j];
#pragma unroll
    for (int i = 0; i < G; ++i) so[i] = os[lane * G1 + i];
    if (a.ddq) row_store<n>(a.ddq, e, ddq);
    if (G > 0) row_store<G1>(a.s_out, e, so);
    if (a.status) a.status[e] = ost[threadIdx.x];
  }
  // ordered admission: the last warp of the launch zeroes the counters
  if (gated && lane == 0) {
    const unsigned done = atomicAdd(a.gate + 1, 1u);
    if (done == gridDim.x * Z::WARPS - 1u) {
      atomicExch(a.gate, 0u);
      atomicExch(a.gate + 2, 0u);
      atomicExch(a.gate + 1, 0u);
    }
  }
}


template <class Env>
__global__ void __launch_bounds__(TPB) atacom_slack_init_kernel(const float* __restrict__ q,
                                                                const float* __restrict__ dq, float* s,
                                                                const uint8_t* __restrict__ mask, int64_t B,
                                                                ParamsT<float> P) {
  using D = typename Env::D;
  constexpr int n = D::n, G = D::G;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TPB + threadIdx.x;
  if (e >= B) return;
  if (mask && !mask[e]) return;
  float qq[n], dd[n], so[at_least_1<G>::value];
#pragma unroll
  for (int j = 0; j < n; ++j) {
    qq[j] = q[e * n + j];
    dd[j] = dq[e * n + j];
  }
  RawConstraints<float, double, D> R;
  Env::template eval<float, double>(P, qq, dd, R);
  slack_from_raw<float, double, D>(P, R, so);
#pragma unroll
  for (int i = 0; i < G; ++i) s[e * G + i] = so[i];
}

// ------------------------------------------------------------------ generic ConstraintsSet
template <int n_, int F_, int G_>
struct GenericDims { using D = Dims<n_, F_, G_>; };

// The user's callbacks were evaluated batched in fp32; everything from there on runs in double (the
// structured -> dense path instantiated for double), so the kernel adds nothing to the rounding of its inputs.
template <int n_, int F_, int G_>
__global__ void __launch_bounds__(TPB) atacom_generic_kernel(const float* __restrict__ c,
                                                             const float* __restrict__ J,
                                                             const float* __restrict__ b, StepArgs a,
                                                             const __grid_constant__ ParamsT<double> P) {
  using D = Dims<n_, F_, G_>;
  constexpr int n = D::n, G = D::G, k = D::k, C = D::C, N = D::N;
  constexpr int G1 = at_least_1<G>::value;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TPB + threadIdx.x;
  if (e >= a.B) return;
  const bool ec = P.variant == VARIANT_EC;
  const int na = ec ? n : k;
  double dq[n], s[G1], al[n], ddq[n], so[G1];
#pragma unroll
  for (int j = 0; j < n; ++j) {
    dq[j] = a.dq[e * n + j];
    al[j] = j < na ? static_cast<double>(a.alpha[e * na + j]) : 0.0;
  }
  RawConstraints<double, double, D> R;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    R.c[i] = c[e * C + i];
    R.b[i] = b[e * C + i];
    double jdq = 0.0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      R.J[i][j] = J[(e * C + i) * n + j];
      jdq += R.J[i][j] * dq[j];
    }
    R.Jdq[i] = jdq;
  }
#pragma unroll
  for (int i = 0; i < G; ++i) s[i] = a.s_in[e * G + i];
  double dbg[2 * N];
  const uint8_t st = step_from_raw<double, double, D, 0>(P, R, dq, s, al, ddq, so, a.w_dbg ? dbg : nullptr);
  if (a.status) a.status[e] = st;
#pragma unroll
  for (int j = 0; j < n; ++j) a.ddq[e * n + j] = static_cast<float>(ddq[j]);
#pragma unroll
  for (int i = 0; i < G; ++i) a.s_out[e * G + i] = static_cast<float>(so[i]);
  if (a.w_dbg) {
#pragma unroll
    for (int i = 0; i < 2 * N; ++i) a.w_dbg[e * (2 * N) + i] = static_cast<float>(dbg[i]);
  }
}

// The same with the reference's LAPACK null basis (ATACOM_BASIS_LAPACK, the default): every environment goes
// through atacom_lapack.cuh, the working array in local memory.
template <int n_, int F_, int G_>
__global__ void __launch_bounds__(TPB) atacom_generic_lapack_kernel(const float* __restrict__ c,
                                                                    const float* __restrict__ J,
                                                                    const float* __restrict__ b, StepArgs a,
                                                                    const __grid_constant__ ParamsT<double> P,
                                                                    const __grid_constant__ DualConsts<double> Kd) {
  using D = Dims<n_, F_, G_>;
  constexpr int n = D::n, G = D::G, k = D::k, C = D::C, N = D::N;
  constexpr int G1 = at_least_1<G>::value;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TPB + threadIdx.x;
  if (e >= a.B) return;
  const bool ec = P.variant == VARIANT_EC;
  const int na = ec ? n : k;
  double dq[n], s[G1], al[n], ddq[n], so[G1];
#pragma unroll
  for (int j = 0; j < n; ++j) {
    dq[j] = a.dq[e * n + j];
    al[j] = j < na ? static_cast<double>(a.alpha[e * na + j]) : 0.0;
  }
  RawConstraints<double, double, D> R;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    R.c[i] = c[e * C + i];
    R.b[i] = b[e * C + i];
    double jdq = 0.0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      R.J[i][j] = J[(e * C + i) * n + j];
      jdq += R.J[i][j] * dq[j];
    }
    R.Jdq[i] = jdq;
  }
#pragma unroll
  for (int i = 0; i < G; ++i) s[i] = a.s_in[e * G + i];
  double dbg[2 * N];
  ArrayStore<double, Lapack<double, D>::SIZE> S;
  const uint8_t st = step_lapack_from_raw<D, double, double>(P, Kd, S, R, dq, s, al, ddq, so, a.w_dbg ? dbg : nullptr);
  if (a.status) a.status[e] = st;
#pragma unroll
  for (int j = 0; j < n; ++j) a.ddq[e * n + j] = static_cast<float>(ddq[j]);
#pragma unroll
  for (int i = 0; i < G; ++i) a.s_out[e * G + i] = static_cast<float>(so[i]);
  if (a.w_dbg) {
#pragma unroll
    for (int i = 0; i < 2 * N; ++i) a.w_dbg[e * (2 * N) + i] = static_cast<float>(dbg[i]);
  }
}

// ------------------------------------------------------------------ PointReachAtacom.step
struct PointArgs {
  const float* q;
  const float* dq;
  const float* p;
  const float* dp;
  const float* s_in;
  const float* action;
  float* w;
  float* s_out;
  uint8_t* status;
  float* w_dbg;
  int64_t B;
};

// fp32 in HBM, double inside (the structured -> dense path instantiated for double)
template <int G_>
__global__ void __launch_bounds__(TPB) point_reach_step_kernel(PointArgs a, const __grid_constant__ ParamsT<double> P) {
  using Env = PointReachEnv<G_>;
  constexpr int G = G_, N = 2 + G_;
  __shared__ __align__(16) float sq[round4_t<TPB * 2>::value];
  __shared__ __align__(16) float sdq[round4_t<TPB * 2>::value];
  __shared__ __align__(16) float sp[round4_t<TPB * 2 * G>::value];
  __shared__ __align__(16) float sdp[round4_t<TPB * 2 * G>::value];
  __shared__ __align__(16) float ss[round4_t<TPB * G>::value];
  __shared__ __align__(16) float sa[round4_t<TPB * 2>::value];
  const int64_t env0 = static_cast<int64_t>(blockIdx.x) * TPB;
  const int64_t left = a.B - env0;
  const int nvalid = left < TPB ? static_cast<int>(left) : TPB;
  slab_load<2>(a.q, sq, env0, nvalid);
  slab_load<2>(a.dq, sdq, env0, nvalid);
  slab_load<2 * G>(a.p, sp, env0, nvalid);
  slab_load<2 * G>(a.dp, sdp, env0, nvalid);
  slab_load<G>(a.s_in, ss, env0, nvalid);
  slab_load<2>(a.action, sa, env0, nvalid);
  __syncthreads();
  const int t = threadIdx.x;
  if (t < nvalid) {
    double q[2], dq[2], p[2 * G], dp[2 * G], s[G], act[2], w[2], so[G];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      q[j] = sq[t * 2 + j];
      dq[j] = sdq[t * 2 + j];
      act[j] = sa[t * 2 + j];
    }
#pragma unroll
    for (int j = 0; j < 2 * G; ++j) {
      p[j] = sp[t * 2 * G + j];
      dp[j] = sdp[t * 2 * G + j];
    }
#pragma unroll
    for (int i = 0; i < G; ++i) s[i] = ss[t * G + i];
    double dbg[2 * N];
    const uint8_t st = Env::template step<double, double>(P, q, dq, p, dp, s, act, w, so, a.w_dbg ? dbg : nullptr);
    if (a.status) a.status[env0 + t] = st;
    if (a.w_dbg) {
#pragma unroll
      for (int i = 0; i < 2 * N; ++i) a.w_dbg[(env0 + t) * (2 * N) + i] = static_cast<float>(dbg[i]);
    }
    sq[t * 2] = static_cast<float>(w[0]);
    sq[t * 2 + 1] = static_cast<float>(w[1]);
#pragma unroll
    for (int i = 0; i < G; ++i) ss[t * G + i] = static_cast<float>(so[i]);
  }
  __syncthreads();
  slab_store<2>(a.w, sq, env0, nvalid);
  slab_store<G>(a.s_out, ss, env0, nvalid);
}

template <int G_>
__global__ void __launch_bounds__(TPB) point_reach_slack_init_kernel(const float* __restrict__ q,
                                                                     const float* __restrict__ p, float* s,
                                                                     const uint8_t* __restrict__ mask,
                                                                     int64_t B, ParamsT<float> P) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TPB + threadIdx.x;
  if (e >= B) return;
  if (mask && !mask[e]) return;
  float qq[2] = {q[e * 2], q[e * 2 + 1]}, pp[2 * G_], so[G_];
#pragma unroll
  for (int j = 0; j < 2 * G_; ++j) pp[j] = p[e * 2 * G_ + j];
  PointReachEnv<G_>::template slack_init<float, double>(P, qq, pp, so);
#pragma unroll
  for (int i = 0; i < G_; ++i) s[e * G_ + i] = so[i];
}

// ------------------------------------------------------------------ constraint statistics and fused roll-outs
// stats[4] (device, double): { sum of c_i, max of c_i, max of c_dq_i, number of samples } with
// c_i = max(|c_f(q)|, c_g(q)) and c_dq_i = max_j(|dq_j| - vel_max_j) per environment and step
// (atacom.py:201-216; circle_base.py:86-115 for the circle's own log).  One warp-reduced atomic update per warp.
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double(static_cast<long long>(old)) < v) {
    const unsigned long long seen = atomicCAS(a, old, static_cast<unsigned long long>(__double_as_longlong(v)));
    if (seen == old) break;
    old = seen;
  }
}

__device__ __forceinline__ void stats_commit(double* stats, double sum_c, double max_c, double max_dq, double count) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum_c += __shfl_xor_sync(0xffffffffu, sum_c, o);
    count += __shfl_xor_sync(0xffffffffu, count, o);
    max_c = fmax(max_c, __shfl_xor_sync(0xffffffffu, max_c, o));
    max_dq = fmax(max_dq, __shfl_xor_sync(0xffffffffu, max_dq, o));
  }
  if ((threadIdx.x & 31) == 0 && count > 0.0) {
    atomicAdd(stats + 0, sum_c);
    atomic_max_double(stats + 1, max_c);
    atomic_max_double(stats + 2, max_dq);
    atomicAdd(stats + 3, count);
  }
}

template <class Env>
__global__ void __launch_bounds__(TPB) atacom_constraint_stats_kernel(const float* __restrict__ q,
                                                                      const float* __restrict__ dq,
                                                                      float* per_env, double* stats, int64_t B,
                                                                      const __grid_constant__ ParamsT<float> P) {
  using D = typename Env::D;
  constexpr int n = D::n;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TPB + threadIdx.x;
  double cm = -1e300, cdq = -1e300;
  if (e < B) {
    float qq[n], dd[n];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      qq[j] = q[e * n + j];
      dd[j] = dq[e * n + j];
      const double v = fabs(static_cast<double>(dd[j])) - static_cast<double>(P.vel_max[j]);
      cdq = v > cdq ? v : cdq;
    }
    StatsSink<float, double, D> sink;
    Env::template eval<float, double>(P, qq, dd, sink);
    cm = sink.cmax;
    if (per_env) {
      per_env[e * 2] = static_cast<float>(cm);
      per_env[e * 2 + 1] = static_cast<float>(cdq);
    }
  }
  if (stats) stats_commit(stats, e < B ? cm : 0.0, cm, cdq, e < B ? 1.0 : 0.0);
}

struct RolloutArgs {
  float* state;          // [B, state_dim], in / out
  float* s;              // [B, G], in / out
  const float* actions;  // [T, B, action_dim]
  const float* draws;    // point reach: [T, B, 2 G] U(-1, 1) draws of the obstacles' random walk, or null
  const float* centers;  // point reach: [B, 2 G] circle centres of the obstacles (when draws is null), or null
  float* rewards;        // [T, B] or null
  double* stats;         // [4] or null
  uint8_t* status;       // [B] or null: OR of the status bits of every step
  int64_t B;
  int32_t T;
  double time0;
};

// CircleEnvAtacom / CircleEnvErrorCorrection, T agent steps per launch with the state in registers:
// AtacomEnvWrapper.step (atacom.py:106-115) -> CircularMotion.step (circle_base.py:53-67) -> hook
// step_action_function (atacom.py:123-139).  env[] = action_scale, base time step, goal x, goal y.
__global__ void __launch_bounds__(TPB) circle_rollout_kernel(const __grid_constant__ RolloutArgs a,
                                                             const __grid_constant__ ParamsT<double> P,
                                                             const __grid_constant__ DualConsts<double> Kd) {
  using Env = CircleEnv;
  using D = Env::D;
  using DU = Dual<double, D, Env::NDIAG>;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TPB + threadIdx.x;
  const bool valid = e < a.B;
  const int64_t ee = valid ? e : a.B - 1;
  const bool ec = P.variant == VARIANT_EC;
  const int na = ec ? 2 : 1;
  double st[4], s[1];
#pragma unroll
  for (int j = 0; j < 4; ++j) st[j] = a.state[ee * 4 + j];
  s[0] = a.s[ee];
  const double scale = P.env[0], dt = P.env[1];
  const double alpha_max = P.acc_max[0] > P.acc_max[1] ? P.acc_max[0] : P.acc_max[1];   // atacom.py:71
  double sum_c = 0.0, max_c = -1e300, max_dq = -1e300;
  uint8_t status = 0;
  for (int t = 0; t < a.T; ++t) {
    double al[2] = {0.0, 0.0};
    for (int j = 0; j < na; ++j) {
      double v = a.actions[(static_cast<int64_t>(t) * a.B + ee) * na + j];
      v = v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v);                                         // atacom.py:107
      al[j] = v * (ec ? P.acc_max[j] : alpha_max);                                       // :108 / ec_wrapper:105-106
    }
    // CircularMotion.check_constraint (circle_base.py:86-92), logged before the step
    {
      const double c1 = fabs(st[0] * st[0] + st[1] * st[1] - 1.0), c2 = -st[1] - 0.5;
      const double cm = c1 > c2 ? c1 : c2;
      const double c3 = fabs(st[2]) - 1.0, c4 = fabs(st[3]) - 1.0;
      sum_c += cm;
      max_c = cm > max_c ? cm : max_c;
      max_dq = fmax(max_dq, c3 > c4 ? c3 : c4);
    }
    double ddq[2], so[1];
    LocalStore<double, DU::Y_SIZE> Ys;
    LocalStore<double, DU::L_SIZE> Ls;
    status |= step_dual<Env, double, double>(P, Kd, Ys, Ls, st, st + 2, s, al, ddq, so, nullptr);
    s[0] = so[0];
    double r2 = 0.0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double u = ddq[j] / P.acc_max[j];                                                  // circle_atacom.py:26-27
      u = (u < -1.0 ? -1.0 : (u > 1.0 ? 1.0 : u)) * scale;                               // circle_base.py:59-60
      st[j] += st[2 + j] * dt + u * dt * dt / 2;                                         // :62
      st[2 + j] += u * dt;                                                               // :63
      const double d = P.env[2 + j] - st[j];
      r2 += d * d;
    }
    if (a.rewards && valid) a.rewards[static_cast<int64_t>(t) * a.B + e] = static_cast<float>(exp(-sqrt(r2)));   // :65
  }
  if (valid) {
#pragma unroll
    for (int j = 0; j < 4; ++j) a.state[e * 4 + j] = static_cast<float>(st[j]);
    a.s[e] = static_cast<float>(s[0]);
    if (a.status) a.status[e] = status;
  }
  if (a.stats) stats_commit(a.stats, valid ? sum_c : 0.0, valid ? max_c : -1e300, valid ? max_dq : -1e300,
                            valid ? static_cast<double>(a.T) : 0.0);
}

// PointReachAtacom.step (collision_avoidance_atacom.py:29-48) -> PointGoalReach.step
// (collision_avoidance_base.py:41-76), T steps per launch.  env[] = radius^2, K, K_c, action_scale, base time
// step, goal x, goal y, wall lo, wall hi, obstacle lo, obstacle hi, obstacle action scale, obstacle speed limit,
// obstacle circle radius.
template <int G_>
__global__ void __launch_bounds__(TPB) point_reach_rollout_kernel(const __grid_constant__ RolloutArgs a,
                                                                  const __grid_constant__ ParamsT<double> P) {
  constexpr int G = G_, SD = 4 + 4 * G_;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TPB + threadIdx.x;
  const bool valid = e < a.B;
  const int64_t ee = valid ? e : a.B - 1;
  double q[2], dq[2], p[2 * G], dp[2 * G], s[G];
  q[0] = a.state[ee * SD]; q[1] = a.state[ee * SD + 1];
  dq[0] = a.state[ee * SD + 2]; dq[1] = a.state[ee * SD + 3];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    p[2 * i] = a.state[ee * SD + 4 + 4 * i];
    p[2 * i + 1] = a.state[ee * SD + 5 + 4 * i];
    dp[2 * i] = a.state[ee * SD + 6 + 4 * i];
    dp[2 * i + 1] = a.state[ee * SD + 7 + 4 * i];
    s[i] = a.s[ee * G + i];
  }
  const double scale = P.env[3], dt = P.env[4];
  double sum_c = 0.0, max_c = -1e300, time = a.time0;
  uint8_t status = 0;
  for (int t = 0; t < a.T; ++t) {
    double act[2], w[2], so[G];
    act[0] = a.actions[(static_cast<int64_t>(t) * a.B + ee) * 2];
    act[1] = a.actions[(static_cast<int64_t>(t) * a.B + ee) * 2 + 1];
    {  // constr_logs.append([max(c_origin), 0]) (collision_avoidance_atacom.py:38-40)
      double cm = -1e300;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const double dx = q[0] - p[2 * i], dy = q[1] - p[2 * i + 1];
        const double c = P.env[0] - (dx * dx + dy * dy);
        cm = c > cm ? c : cm;
      }
      sum_c += cm;
      max_c = cm > max_c ? cm : max_c;
    }
    status |= PointReachEnv<G>::template step<double, double>(P, q, dq, p, dp, s, act, w, so, nullptr);
#pragma unroll
    for (int i = 0; i < G; ++i) s[i] = so[i];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double u = (w[j] < -1.0 ? -1.0 : (w[j] > 1.0 ? 1.0 : w[j])) * scale;          // base.py:44-45
      q[j] += dq[j] * dt;                                                                // :47
      dq[j] += u * dt;                                                                   // :48
      if (q[j] <= P.env[7] || q[j] >= P.env[8]) dq[j] = -dq[j];                          // :50-52
    }
    if (a.draws) {                                                                       // random walk, :57-66
#pragma unroll
      for (int c = 0; c < 2 * G; ++c) {
        p[c] += dp[c] * dt;
        p[c] = p[c] < P.env[9] ? P.env[9] : (p[c] > P.env[10] ? P.env[10] : p[c]);
        const double oa = a.draws[(static_cast<int64_t>(t) * a.B + ee) * (2 * G) + c] * P.env[11];
        if (p[c] <= P.env[9] || p[c] >= P.env[10]) dp[c] = -dp[c];
        dp[c] += oa * dt;
        dp[c] = dp[c] < -P.env[12] ? -P.env[12] : (dp[c] > P.env[12] ? P.env[12] : dp[c]);
      }
    } else if (a.centers) {                                                              // circles, :67-72
      const double ang = time * 6.283185307179586, rad = P.env[13];
      double sn, cs;
      sincos(ang, &sn, &cs);
#pragma unroll
      for (int i = 0; i < G; ++i) {
        p[2 * i] = a.centers[ee * 2 * G + 2 * i] + rad * cs;
        p[2 * i + 1] = a.centers[ee * 2 * G + 2 * i + 1] + rad * sn;
        dp[2 * i] = -2 * rad * 3.141592653589793 * sn;
        dp[2 * i + 1] = 2 * rad * 3.141592653589793 * cs;
      }
    }
    time += dt;                                                                          // :74
    if (a.rewards && valid) {
      const double dx = P.env[5] - q[0], dy = P.env[6] - q[1];
      a.rewards[static_cast<int64_t>(t) * a.B + e] = static_cast<float>(-sqrt(dx * dx + dy * dy) / (8.0 * 1.4142135623730951));   // :75
    }
  }
  if (valid) {
    a.state[e * SD] = static_cast<float>(q[0]); a.state[e * SD + 1] = static_cast<float>(q[1]);
    a.state[e * SD + 2] = static_cast<float>(dq[0]); a.state[e * SD + 3] = static_cast<float>(dq[1]);
#pragma unroll
    for (int i = 0; i < G; ++i) {
      a.state[e * SD + 4 + 4 * i] = static_cast<float>(p[2 * i]);
      a.state[e * SD + 5 + 4 * i] = static_cast<float>(p[2 * i + 1]);
      a.state[e * SD + 6 + 4 * i] = static_cast<float>(dp[2 * i]);
      a.state[e * SD + 7 + 4 * i] = static_cast<float>(dp[2 * i + 1]);
      a.s[e * G + i] = static_cast<float>(s[i]);
    }
    if (a.status) a.status[e] = status;
  }
  if (a.stats) stats_commit(a.stats, valid ? sum_c : 0.0, valid ? max_c : -1e300, valid ? 0.0 : -1e300,
                            valid ? static_cast<double>(a.T) : 0.0);
}

// ------------------------------------------------------------------ host-side helpers
inline const ParamsT<float>& as_params(const AtacomParams* p) {
  return *reinterpret_cast<const ParamsT<float>*>(p);
}

inline int check_launch() {
  const cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? ATACOM_OK : ATACOM_ERR_CUDA;
}

inline unsigned blocks_for(int64_t B) { return static_cast<unsigned>((B + TPB - 1) / TPB); }

// Block size of the step kernels: spread B environments over the SMs in as few equal waves as possible.
int step_block_size(int64_t B) {
  static std::atomic<int> sm_counts[MAX_DEVICES];   // per device: a process may drive several GPUs
  const int dev = current_device();
  int sm_count = sm_counts[dev].load(std::memory_order_relaxed);
  if (sm_count == 0) {
    int v = 0;
    sm_count = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
    sm_counts[dev].store(sm_count, std::memory_order_relaxed);
  }
  if (const char* f = getenv("ATACOM_STEP_TPB")) {   // experiments only
    const int v = atoi(f);
    if (v >= 32 && v <= STEP_MAX_TPB) return v / 32 * 32;
  }
  const int64_t waves = (B + static_cast<int64_t>(sm_count) * STEP_MAX_TPB - 1) / (static_cast<int64_t>(sm_count) * STEP_MAX_TPB);
  const int64_t blocks = waves * sm_count;                       // one block per SM per wave
  int64_t tpb = (B + blocks - 1) / blocks;
  tpb = (tpb + 31) / 32 * 32;
  if (tpb < 64) tpb = 64;
  if (tpb > STEP_MAX_TPB) tpb = STEP_MAX_TPB;
  return static_cast<int>(tpb);
}

// Programmatic dependent launch of the step kernels (ATACOM_PDL = 0 | 1 overrides the default).
bool step_pdl_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* f = getenv("ATACOM_PDL");
    mode = f ? (atoi(f) != 0 ? 1 : 0) : ATACOM_PDL_DEFAULT;
  }
  return mode == 1;
}

// Fused gather: rows to the peers as per-warp bulk stores (large NVLink writes) or as per-thread row stores.
// ATACOM_GATHER_IO = rows | bulk overrides the default (experiments).
bool gather_bulk_stores() {
  static int mode = -1;
  if (mode < 0) {
    const char* f = getenv("ATACOM_GATHER_IO");
    mode = (f && f[0] == 'r') ? 0 : 1;
  }
  return mode == 1;
}

int check_common(int64_t B, const AtacomParams* p) {
  if (!p) return ATACOM_ERR_NULL_POINTER;
  if (B < 0 || B > (int64_t(1) << 31) * TPB) return ATACOM_ERR_BAD_DIMS;
  if (p->variant != ATACOM_VARIANT_ATACOM && p->variant != ATACOM_VARIANT_ERROR_CORRECTION)
    return ATACOM_ERR_BAD_PARAM;
  if (!(p->rref_tol >= 0.f) || !(p->dt == p->dt)) return ATACOM_ERR_BAD_PARAM;
  return ATACOM_OK;
}

// Opt the kernel in to its dynamic shared memory: the attribute belongs to the (function, device) pair, so
// once per instantiation AND device (a process that first runs on cuda:0 and then on cuda:1 needs both).
template <class Env, int IO, bool BAND = false>
bool configure_step_kernel() {
  static std::atomic<int> states[MAX_DEVICES];   // 0: not yet, 1: done, -1: failed
  std::atomic<int>& state = states[current_device()];
  if (state.load(std::memory_order_acquire) == 0) {
    constexpr size_t smem = StepScratch<Env>::BYTES;
    state.store((smem <= 48 * 1024 ||
                 cudaFuncSetAttribute(atacom_step_kernel<Env, IO, BAND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)) == cudaSuccess) ? 1 : -1, std::memory_order_release);
  }
  return state.load(std::memory_order_acquire) == 1;
}

// The per-launch scratch comes from the device's default stream-ordered pool, whose release threshold is zero by
// default: every synchronisation would hand the memory back to the OS and the next launch would pay for a fresh
// mapping.  Raised once per device, so that the pool keeps what it has.
void keep_scratch_pool_warm() {
  static std::atomic<int> done[MAX_DEVICES];
  const int dev = current_device();
  if (done[dev].load(std::memory_order_acquire)) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = 64ull << 20;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[dev].store(1, std::memory_order_release);
}

template <class Env>
bool configure_fix_kernel() {
  static std::atomic<int> states[MAX_DEVICES];
  std::atomic<int>& state = states[current_device()];
  if (state.load(std::memory_order_acquire) == 0) {
    state.store((FixCfg<Env>::BYTES <= 48 * 1024 ||
                 cudaFuncSetAttribute(atacom_fix_kernel<Env>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(FixCfg<Env>::BYTES)) == cudaSuccess) ? 1 : -1,
                std::memory_order_release);
  }
  return state.load(std::memory_order_acquire) == 1;
}

template <class Env>
bool configure_zc_fused_kernel() {
  static std::atomic<int> states[MAX_DEVICES];
  std::atomic<int>& state = states[current_device()];
  if (state.load(std::memory_order_acquire) == 0) {
    state.store((ZcFused<Env>::BYTES <= 48 * 1024 ||
                 cudaFuncSetAttribute(atacom_zc_fused_kernel<Env>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(ZcFused<Env>::BYTES)) == cudaSuccess) ? 1 : -1,
                std::memory_order_release);
  }
  return state.load(std::memory_order_acquire) == 1;
}

// Both passes of a step on the caller's page-locked arrays in one launch (atacom_zc_fused_kernel).
template <class Env>
int launch_zc_fused(const float* q, const float* dq, const float* s_in, const float* alpha, float* ddq, float* s_out,
                    uint8_t* status, int64_t B, const AtacomParams* p, void* stream, uint32_t* gate, int gate_window) {
  using D = typename Env::D;
  int rc = check_common(B, p);
  if (rc) return rc;
  if (B == 0) return ATACOM_OK;
  if (!q || !dq || !ddq || (D::G > 0 && (!s_in || !s_out)) || (D::k > 0 && !alpha)) return ATACOM_ERR_NULL_POINTER;
  if (!configure_zc_fused_kernel<Env>()) return ATACOM_ERR_CUDA;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(s_in) |
                         reinterpret_cast<uintptr_t>(alpha) | reinterpret_cast<uintptr_t>(ddq) |
                         reinterpret_cast<uintptr_t>(s_out) | reinterpret_cast<uintptr_t>(status);
  ZcArgs a{q, dq, s_in, alpha, ddq, s_out, status, B, (bits & 15) == 0 ? 1 : 0, gate_window,
           (gate != nullptr && gate_window > 0) ? gate : nullptr};
  const ParamsT<float> Pk = as_params(p);
  const DualConsts<double> Kd = make_dual_consts<float, double>(Pk, D::F, D::G);
  const unsigned grid = static_cast<unsigned>((B + ZcFused<Env>::TPB - 1) / ZcFused<Env>::TPB);
  NvtxRange range("atacom_step");
  atacom_zc_fused_kernel<Env><<<grid, ZcFused<Env>::TPB, ZcFused<Env>::BYTES, static_cast<cudaStream_t>(stream)>>>(a, Pk, Kd);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

template <class Env, int IO = (ATACOM_STEP_STAGED_IO == 1 ? 1 : ATACOM_STEP_DEVICE_IO)>
int launch_step(const float* q, const float* dq, const float* s_in, const float* alpha, float* ddq, float* s_out,
                uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p, void* stream,
                float* const* peers = nullptr, int n_peers = 0, int64_t gather_row0 = 0,
                uint32_t* const* peer_flags = nullptr, uint32_t* local_sync = nullptr, int rank = 0,
                uint32_t* gate = nullptr, int gate_window = 0, int tpb_override = 0, bool mirror_inputs = false) {
  int rc = check_common(B, p);
  if (rc) return rc;
  using D = typename Env::D;
  if (B == 0) return ATACOM_OK;
  if (n_peers < 0 || n_peers > ATACOM_MAX_PEERS || (n_peers > 0 && !peers) || gather_row0 < 0) return ATACOM_ERR_BAD_DIMS;
  if (!q || !dq || (!ddq && n_peers == 0) || (D::G > 0 && (!s_in || !s_out))) return ATACOM_ERR_NULL_POINTER;
  const bool needs_alpha = p->variant == ATACOM_VARIANT_ERROR_CORRECTION || D::k > 0;
  if (needs_alpha && !alpha) return ATACOM_ERR_NULL_POINTER;
  StepArgs a{q, dq, s_in, alpha, ddq, s_out, status, w_dbg, B, {}, gather_row0, n_peers, 0, {}, nullptr, rank, 0, nullptr,
             nullptr, nullptr};
  if (IO == 1 && gate != nullptr && gate_window > 0) {
    a.gate = gate;
    a.gate_window = gate_window;
  }
  if (local_sync) {
    if (!peer_flags || n_peers < 1 || rank < 0 || rank >= n_peers) return ATACOM_ERR_BAD_PARAM;
    for (int w = 0; w < n_peers; ++w) {
      if (!peer_flags[w]) return ATACOM_ERR_NULL_POINTER;
      a.peer_flags[w] = peer_flags[w];
    }
    a.local_sync = local_sync;
  }
  uintptr_t bits = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(s_in) |
                   reinterpret_cast<uintptr_t>(alpha) | reinterpret_cast<uintptr_t>(ddq) | reinterpret_cast<uintptr_t>(s_out);
  for (int w = 0; w < n_peers; ++w) {
    if (!peers[w]) return ATACOM_ERR_NULL_POINTER;
    a.peer[w] = peers[w];
    bits |= reinterpret_cast<uintptr_t>(peers[w]) | static_cast<uintptr_t>((gather_row0 * D::n * 4) & 15);
  }
  a.aligned16 = (bits & 15) == 0;
  const int tpb = (tpb_override >= 32 && tpb_override <= STEP_MAX_TPB) ? tpb_override / 32 * 32 : step_block_size(B);
  const unsigned grid = static_cast<unsigned>((B + tpb - 1) / tpb);
  const size_t smem = StepScratch<Env>::bytes(tpb);
  NvtxRange range("atacom_step");
  // ATACOM_BASIS_LAPACK (default): the step kernel leaves the environments inside the tolerance band of rref to the
  // fix-up kernel.  The per-block lists live in stream-ordered scratch memory (capturable in a CUDA graph).
  const bool two_pass = p->basis_mode == ATACOM_BASIS_LAPACK && p->variant == ATACOM_VARIANT_ATACOM && D::k > 1;
  constexpr bool CAN_BAND = D::k > 1;          // (k = 1: the null basis is unique up to its sign, nothing to defer)
  if (!(two_pass ? configure_step_kernel<Env, IO, CAN_BAND>() : configure_step_kernel<Env, IO, false>())) return ATACOM_ERR_CUDA;
  const void* kernel = two_pass ? reinterpret_cast<const void*>(&atacom_step_kernel<Env, IO, CAN_BAND>)
                                : reinterpret_cast<const void*>(&atacom_step_kernel<Env, IO, false>);
  if (p->basis_mode != ATACOM_BASIS_LAPACK && p->basis_mode != ATACOM_BASIS_CANONICAL) return ATACOM_ERR_BAD_PARAM;
  if (two_pass && local_sync) return ATACOM_ERR_BAD_PARAM;   // the in-kernel barrier would publish the step before the fix-up
  int32_t* scratch = nullptr;
  if (two_pass) {
    if (!configure_fix_kernel<Env>()) return ATACOM_ERR_CUDA;
    keep_scratch_pool_warm();
    // [grid * tpb] indices, [grid] counts, then (8-byte aligned) [grid * tpb, N] minimum-norm parts
    const size_t n_idx = (static_cast<size_t>(grid) * tpb + grid + 1) / 2 * 2;
    const size_t rows = static_cast<size_t>(grid) * tpb;
    const size_t wmn_bytes = ATACOM_FIX_NULL_ONLY ? sizeof(double) * rows * D::N : 0;
    // ... then, for a launch on mapped host arrays, the mirror of the input rows (sections 16-byte aligned)
    const bool mirror = mirror_inputs && IO == 1;
    auto sec = [&](size_t width) { return (sizeof(float) * rows * width + 15) / 16 * 16; };
    const size_t mirror_bytes = mirror ? 2 * sec(D::n) + sec(D::G) + sec(D::k) : 0;
    const size_t bytes = sizeof(int32_t) * n_idx + wmn_bytes + mirror_bytes;
    if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), bytes, static_cast<cudaStream_t>(stream)) != cudaSuccess) {
      cudaGetLastError();
      return ATACOM_ERR_CUDA;
    }
    a.fix_list = scratch;
    a.fix_count = scratch + static_cast<size_t>(grid) * tpb;
    a.fix_wmn = ATACOM_FIX_NULL_ONLY ? reinterpret_cast<double*>(scratch + n_idx) : nullptr;
    if (mirror) {
      unsigned char* m = reinterpret_cast<unsigned char*>(scratch + n_idx) + wmn_bytes;
      a.mirror_q = reinterpret_cast<float*>(m);
      a.mirror_dq = reinterpret_cast<float*>(m + sec(D::n));
      a.mirror_s = reinterpret_cast<float*>(m + 2 * sec(D::n));
      a.mirror_alpha = reinterpret_cast<float*>(m + 2 * sec(D::n) + sec(D::G));
    }
  }
  auto fix_up = [&]() -> int {      // second launch + release of the scratch, after the step kernel is in the stream
    if (!two_pass) return ATACOM_OK;
    FixArgs f{q, dq, s_in, alpha, ddq, s_out, status, w_dbg, a.fix_list, a.fix_count, a.fix_wmn, tpb, {}, gather_row0,
              n_peers};
    if (a.mirror_q != nullptr) {      // the rows the step kernel has copied into device memory
      f.q = a.mirror_q;
      f.dq = a.mirror_dq;
      f.s_in = a.mirror_s;
      f.alpha = a.mirror_alpha;
    }
    for (int w = 0; w < n_peers; ++w) f.peer[w] = peers[w];
    const ParamsT<float> Pk = as_params(p);
    const DualConsts<double> Kd = make_dual_consts<float, double>(Pk, D::F, D::G);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(static_cast<unsigned>(FixCfg<Env>::TPB));
    cfg.dynamicSmemBytes = FixCfg<Env>::BYTES;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = step_pdl_enabled() ? 1 : 0;
    static const bool skip_fix = getenv("ATACOM_DEBUG_SKIP_FIX") != nullptr;     // debugging aid: step kernel alone
    const bool ok = skip_fix || cudaLaunchKernelEx(&cfg, atacom_fix_kernel<Env>, f, Pk, Kd) == cudaSuccess;
    cudaFreeAsync(scratch, static_cast<cudaStream_t>(stream));
    if (!ok) {
      cudaGetLastError();
      return ATACOM_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ATACOM_OK;
  };
  if (step_pdl_enabled() && IO != 1 && n_peers == 0) {   // (fused gather: measured neutral to slightly slower, left out)
    // Programmatic dependent launch: the blocks of this launch may start — one by one, as the blocks of the
    // previous kernel in the stream leave their SMs — and set up while that kernel drains; they touch global
    // memory only after griddepcontrol.wait, i.e. once it has completed and flushed.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(static_cast<unsigned>(tpb));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const ParamsT<float> Pk = as_params(p);
    const DualConsts<double> Kd = make_dual_consts<float, double>(Pk, D::F, D::G);
    void* kargs[3] = {const_cast<StepArgs*>(&a), const_cast<ParamsT<float>*>(&Pk), const_cast<DualConsts<double>*>(&Kd)};
    if (cudaLaunchKernelExC(&cfg, kernel, kargs) != cudaSuccess) {
      cudaGetLastError();
      if (scratch) cudaFreeAsync(scratch, static_cast<cudaStream_t>(stream));
      return ATACOM_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return fix_up();
  }
  {
    const ParamsT<float> Pk = as_params(p);
    const DualConsts<double> Kd = make_dual_consts<float, double>(Pk, D::F, D::G);
    void* kargs[3] = {&a, const_cast<ParamsT<float>*>(&Pk), const_cast<DualConsts<double>*>(&Kd)};
    cudaLaunchKernel(kernel, dim3(grid), dim3(static_cast<unsigned>(tpb)), kargs, smem, static_cast<cudaStream_t>(stream));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  rc = check_launch();
  if (rc != ATACOM_OK) {
    if (scratch) cudaFreeAsync(scratch, static_cast<cudaStream_t>(stream));
    return rc;
  }
  return fix_up();
}

template <class Env>
int launch_slack_init(const float* q, const float* dq, float* s, const uint8_t* mask, int64_t B,
                      const AtacomParams* p, void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if (B == 0) return ATACOM_OK;
  if (!q || !dq || !s) return ATACOM_ERR_NULL_POINTER;
  atacom_slack_init_kernel<Env><<<blocks_for(B), TPB, 0, static_cast<cudaStream_t>(stream)>>>(q, dq, s, mask, B,
                                                                                            as_params(p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

template <class Env, bool PARK>
int launch_substeps(const SubstepArgs& a, const ParamsT<float>& P, unsigned grid, int tpb, cudaStream_t st) {
  static std::atomic<int> cfgs[MAX_DEVICES];   // per instantiation and device
  std::atomic<int>& cfg = cfgs[current_device()];
  if (cfg.load(std::memory_order_acquire) == 0)
    cfg.store(cudaFuncSetAttribute(atacom_substeps_kernel<Env, PARK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(StepScratch<Env>::BYTES)) == cudaSuccess ? 1 : -1,
              std::memory_order_release);
  if (cfg.load(std::memory_order_acquire) < 0) return ATACOM_ERR_CUDA;
  using D = typename Env::D;
  atacom_substeps_kernel<Env, PARK><<<grid, tpb, StepScratch<Env>::bytes(tpb), st>>>(
      a, P, make_dual_consts<float, double>(P, D::F, D::G));
  return ATACOM_OK;
}

template <class Env>
int launch_stats(const float* q, const float* dq, float* per_env, double* stats, int64_t B,
                        const AtacomParams* p, void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if (B == 0) return ATACOM_OK;
  if (!q || !dq || (!per_env && !stats)) return ATACOM_ERR_NULL_POINTER;
  atacom_constraint_stats_kernel<Env><<<blocks_for(B), TPB, 0, static_cast<cudaStream_t>(stream)>>>(
      q, dq, per_env, stats, B, as_params(p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

void fill_common(AtacomParams* p) {
  for (size_t i = 0; i < sizeof(*p) / 4; ++i) reinterpret_cast<uint32_t*>(p)[i] = 0;
  p->rref_tol = 0.05f;  // atacom.py:128
  p->variant = ATACOM_VARIANT_ATACOM;
  // what the reference executes: first-order forwardKinematics leaves data.a at zero, so
  // getFrameClassicalAcceleration returns omega x v (iiwa_hit_atacom.py:87-89,122-128; atacom_air_hockey.py:94-96)
  p->bias_mode = ATACOM_BIAS_OMEGA_X_V;
  p->clip_acc = 1;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

const char* atacom_version(void) { return "atacom_b200 0.1.0 (sm_100a)"; }

const char* atacom_error_string(int code) {
  switch (code) {
    case ATACOM_OK: return "ok";
    case ATACOM_ERR_NULL_POINTER: return "null pointer argument";
    case ATACOM_ERR_BAD_DIMS: return "unsupported dimensions";
    case ATACOM_ERR_BAD_PARAM: return "invalid AtacomParams field";
    case ATACOM_ERR_CUDA: return "CUDA error";
    case ATACOM_ERR_NO_DEVICE: return "no CUDA device";
    case ATACOM_ERR_ALIGNMENT: return "pointer not 4-byte aligned";
    default: return "unknown error";
  }
}

int64_t atacom_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int atacom_circle_default_params(AtacomParams* p) {
  if (!p) return ATACOM_ERR_NULL_POINTER;
  fill_common(p);
  p->K_f[0] = 0.1f;  // circle_atacom.py:10
  p->K_g[0] = 2.f;   // circle_atacom.py:12
  for (int i = 0; i < 2; ++i) {
    p->K_c[i] = 100.f;  // circle_atacom.py:8 (Kc=100)
    p->K_q[i] = 20.f;   // circle_atacom.py:18
    p->vel_max[i] = 1.f;
    p->acc_max[i] = 10.f;
  }
  p->dt = 0.01f;
  p->env[0] = 10.0;   // CircularMotion.action_scale   circle_base.py:26
  p->env[1] = 0.01;   // CircularMotion.time_step      circle_base.py:12
  p->env[2] = 1.0;    // goal                          circle_base.py:65
  p->env[3] = 0.0;
  return ATACOM_OK;
}

int atacom_planar_default_params(AtacomParams* p) {
  if (!p) return ATACOM_ERR_NULL_POINTER;
  fill_common(p);
  const float vel[3] = {1.4835298641951802f, 1.4835298641951802f, 1.7453292519943295f};
  const double qmax[3] = {2.9670597283903604, 2.0943951023931953, 2.0943951023931953};
  for (int i = 0; i < 3; ++i) {
    p->K_g[i] = 0.5f;      // atacom_air_hockey.py:31
    p->K_g[3 + i] = 1.f;   // atacom_air_hockey.py:33
    p->acc_max[i] = 10.f;  // atacom_air_hockey.py:41
    p->vel_max[i] = vel[i];
    p->K_q[i] = 2.f * 10.f / vel[i];  // atacom_air_hockey.py:43
    p->env[5 + i] = qmax[i];
  }
  for (int i = 0; i < 6; ++i) p->K_c[i] = 240.f;
  p->dt = 1.f / 240.f;
  p->env[0] = 0.55;  // link lengths, base and table: recalled from mushroom-rl's planar URDF, unverified
  p->env[1] = 0.44;
  p->env[2] = 0.44;
  p->env[3] = -1.51;
  p->env[4] = 0.0;
  p->env[8] = 1.96 / 2 - 0.05;
  p->env[9] = 1.02 / 2 - 0.05;
  return ATACOM_OK;
}

int atacom_iiwa_default_params(AtacomParams* p, int n) {
  if (!p) return ATACOM_ERR_NULL_POINTER;
  if (n != 6 && n != 7) return ATACOM_ERR_BAD_DIMS;
  fill_common(p);
  const double vel[7] = {1.4835298641951802, 1.4835298641951802, 1.7453292519943295, 1.3089969389957472,
                         2.2689280275926285, 2.356194490192345,  2.356194490192345};  // urdf/iiwa_1.urdf:74..297
  const double qmax[7] = {2.9670597283903604, 2.0943951023931953, 2.9670597283903604, 2.0943951023931953,
                          2.9670597283903604, 2.0943951023931953, 3.0543261909900763};
  p->K_f[0] = 0.1f;  // iiwa_hit_atacom.py:26
  for (int i = 0; i < 5; ++i) p->K_g[i] = 0.5f;      // :31
  for (int i = 0; i < n; ++i) p->K_g[5 + i] = 1.f;   // :33
  for (int i = 0; i < 6 + n; ++i) p->K_c[i] = 240.f;  // :13
  for (int i = 0; i < n; ++i) {
    p->acc_max[i] = 10.f;                               // :38
    p->vel_max[i] = static_cast<float>(vel[i]);         // :39
    p->K_q[i] = static_cast<float>(4.0 * 10.0 / vel[i]);  // :40
  }
  for (int i = 0; i < 7; ++i) p->env[6 + i] = qmax[i];
  p->dt = 1.f / 240.f;
  p->env[0] = -1.51;              // env_base.py:50
  p->env[1] = 1.96 / 2 - 0.05;    // env_base.py:156,158
  p->env[2] = 1.02 / 2 - 0.05;
  p->env[3] = 0.1505;             // env_base.py:159
  p->env[4] = 0.36;               // iiwa_hit_atacom.py:106
  p->env[5] = 0.25;               // iiwa_hit_atacom.py:107
  return ATACOM_OK;
}

int atacom_point_reach_default_params(AtacomParams* p) {
  if (!p) return ATACOM_ERR_NULL_POINTER;
  fill_common(p);
  p->clip_acc = 0;
  p->dt = 0.01f;
  // rref default tolerance max(m,n)*eps*||V||_inf (null_space_coordinate.py:48-49) restated for fp32
  p->rref_tol = 8.f * 1.1920929e-07f * 2.5f;
  p->env[0] = 0.36;   // collision_avoidance_atacom.py:75
  p->env[1] = 0.5;    // :14
  p->env[2] = 100.0;  // :13
  p->env[3] = 10.0;   // PointGoalReach.action_scale             collision_avoidance_base.py:19
  p->env[4] = 0.01;   // time_step                               :7
  p->env[5] = 9.0;    // goal                                    :75
  p->env[6] = 9.0;
  p->env[7] = 0.0;    // walls                                   :50
  p->env[8] = 10.0;
  p->env[9] = 2.0;    // obstacle box                            :59,61
  p->env[10] = 10.0;
  p->env[11] = 10.0;  // obstacle random-walk action scale       :60
  p->env[12] = 1.0;   // obstacle speed limit                    :66
  p->env[13] = 2.0;   // obstacle circle radius                  :21
  return ATACOM_OK;
}

int atacom_circle_step(const float* q, const float* dq, const float* s_in, const float* alpha, float* ddq,
                       float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                       void* stream) {
  return launch_step<CircleEnv>(q, dq, s_in, alpha, ddq, s_out, status, w_dbg, B, p, stream);
}

int atacom_planar_step(const float* q, const float* dq, const float* s_in, const float* alpha, float* ddq,
                       float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                       void* stream) {
  return launch_step<PlanarEnv>(q, dq, s_in, alpha, ddq, s_out, status, w_dbg, B, p, stream);
}

int atacom_iiwa_step(int n, const float* q, const float* dq, const float* s_in, const float* alpha, float* ddq,
                     float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                     void* stream) {
  if (n == 6) return launch_step<IiwaEnv<6>>(q, dq, s_in, alpha, ddq, s_out, status, w_dbg, B, p, stream);
  if (n == 7) return launch_step<IiwaEnv<7>>(q, dq, s_in, alpha, ddq, s_out, status, w_dbg, B, p, stream);
  return ATACOM_ERR_BAD_DIMS;
}

int atacom_iiwa_step_gather(int n, const float* q, const float* dq, const float* s_in, const float* alpha,
                            float* ddq, float* s_out, uint8_t* status, int64_t B, const AtacomParams* p,
                            void* stream, float* const* peer_ddq, int world, int64_t row_offset) {
  if (world < 1) return ATACOM_ERR_BAD_DIMS;
  if (gather_bulk_stores()) {
    if (n == 6)
      return launch_step<IiwaEnv<6>, 2>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset);
    if (n == 7)
      return launch_step<IiwaEnv<7>, 2>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset);
  } else {
    if (n == 6)
      return launch_step<IiwaEnv<6>, 0>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset);
    if (n == 7)
      return launch_step<IiwaEnv<7>, 0>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset);
  }
  return ATACOM_ERR_BAD_DIMS;
}

int atacom_iiwa_step_gather_sync(int n, const float* q, const float* dq, const float* s_in, const float* alpha,
                                 float* ddq, float* s_out, uint8_t* status, int64_t B, const AtacomParams* p,
                                 void* stream, float* const* peer_ddq, int world, int64_t row_offset,
                                 uint32_t* const* peer_flags, uint32_t* local_sync, int rank) {
  if (world < 1 || !peer_flags || !local_sync) return ATACOM_ERR_BAD_DIMS;
  if (B == 0) return ATACOM_ERR_BAD_DIMS;   // an empty launch could not take part in the barrier
  if (gather_bulk_stores()) {
    if (n == 6)
      return launch_step<IiwaEnv<6>, 2>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset, peer_flags, local_sync, rank);
    if (n == 7)
      return launch_step<IiwaEnv<7>, 2>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset, peer_flags, local_sync, rank);
  } else {
    if (n == 6)
      return launch_step<IiwaEnv<6>, 0>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset, peer_flags, local_sync, rank);
    if (n == 7)
      return launch_step<IiwaEnv<7>, 0>(q, dq, s_in, alpha, ddq, s_out, status, nullptr, B, p, stream, peer_ddq,
                                        world, row_offset, peer_flags, local_sync, rank);
  }
  return ATACOM_ERR_BAD_DIMS;
}

int atacom_iiwa_substeps_workspace_doubles(int n) {
  if (n == 6) return Dual<double, IiwaEnv<6>::D, 6>::Y_SIZE + 6 + IiwaEnv<6>::D::C;
  if (n == 7) return Dual<double, IiwaEnv<7>::D, 7>::Y_SIZE + 7 + IiwaEnv<7>::D::C;
  return 0;
}

int atacom_iiwa_step_substeps(int n, int K, const float* q, const float* dq, const float* s_in, const float* alpha,
                              float* ddq, float* s_out, uint8_t* status, double* workspace, int64_t B,
                              const AtacomParams* p, void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if ((n != 6 && n != 7) || K < 1 || K > 64) return ATACOM_ERR_BAD_DIMS;
  if (B == 0) return ATACOM_OK;
  if (!q || !dq || !s_in || !alpha || !ddq || !s_out) return ATACOM_ERR_NULL_POINTER;
  if (p->basis_mode == ATACOM_BASIS_LAPACK && p->variant == ATACOM_VARIANT_ATACOM) {
    // every sub-step may leave environments to the fix-up kernel, and the next sub-step needs their slacks: K
    // two-pass steps back to back (the fused single launch is the canonical-basis mode's)
    for (int kk = 0; kk < K; ++kk) {
      rc = atacom_iiwa_step(n, q, dq, kk == 0 ? s_in : s_out, alpha, ddq + static_cast<int64_t>(kk) * B * n, s_out,
                            status, nullptr, B, p, stream);      // (status: the last sub-step's)
      if (rc != ATACOM_OK) return rc;
    }
    return ATACOM_OK;
  }
  SubstepArgs a{q, dq, s_in, alpha, ddq, s_out, status, workspace, B, K};
  const int tpb = step_block_size(B);
  const unsigned grid = static_cast<unsigned>((B + tpb - 1) / tpb);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ParamsT<float>& P = as_params(p);
  rc = n == 6 ? (workspace && K > 1 ? launch_substeps<IiwaEnv<6>, true>(a, P, grid, tpb, st)
                                    : launch_substeps<IiwaEnv<6>, false>(a, P, grid, tpb, st))
              : (workspace && K > 1 ? launch_substeps<IiwaEnv<7>, true>(a, P, grid, tpb, st)
                                    : launch_substeps<IiwaEnv<7>, false>(a, P, grid, tpb, st));
  if (rc != ATACOM_OK) return rc;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

int atacom_circle_slack_init(const float* q, const float* dq, float* s, const uint8_t* mask, int64_t B,
                             const AtacomParams* p, void* stream) {
  return launch_slack_init<CircleEnv>(q, dq, s, mask, B, p, stream);
}

int atacom_planar_slack_init(const float* q, const float* dq, float* s, const uint8_t* mask, int64_t B,
                             const AtacomParams* p, void* stream) {
  return launch_slack_init<PlanarEnv>(q, dq, s, mask, B, p, stream);
}

int atacom_iiwa_slack_init(int n, const float* q, const float* dq, float* s, const uint8_t* mask, int64_t B,
                           const AtacomParams* p, void* stream) {
  if (n == 6) return launch_slack_init<IiwaEnv<6>>(q, dq, s, mask, B, p, stream);
  if (n == 7) return launch_slack_init<IiwaEnv<7>>(q, dq, s, mask, B, p, stream);
  return ATACOM_ERR_BAD_DIMS;
}

#define ATACOM_POINT_DISPATCH(G_, CALL) \
  case G_: CALL(G_); break;

int atacom_point_reach_step(int n_objects, const float* q, const float* dq, const float* obs_p,
                            const float* obs_dp, const float* s_in, const float* action, float* w,
                            float* s_out, uint8_t* status, float* w_dbg, int64_t B, const AtacomParams* p,
                            void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if (B == 0) return ATACOM_OK;
  if (!q || !dq || !obs_p || !obs_dp || !s_in || !action || !w || !s_out) return ATACOM_ERR_NULL_POINTER;
  PointArgs a{q, dq, obs_p, obs_dp, s_in, action, w, s_out, status, w_dbg, B};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ParamsT<double> Pd = widen_params<double>(as_params(p));
#define ATACOM_PR_STEP(G_) point_reach_step_kernel<G_><<<blocks_for(B), TPB, 0, st>>>(a, Pd)
  switch (n_objects) {
    ATACOM_POINT_DISPATCH(1, ATACOM_PR_STEP)
    ATACOM_POINT_DISPATCH(2, ATACOM_PR_STEP)
    ATACOM_POINT_DISPATCH(3, ATACOM_PR_STEP)
    ATACOM_POINT_DISPATCH(4, ATACOM_PR_STEP)
    ATACOM_POINT_DISPATCH(6, ATACOM_PR_STEP)
    ATACOM_POINT_DISPATCH(8, ATACOM_PR_STEP)
    default: return ATACOM_ERR_BAD_DIMS;
  }
#undef ATACOM_PR_STEP
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

int atacom_point_reach_slack_init(int n_objects, const float* q, const float* obs_p, float* s,
                                  const uint8_t* mask, int64_t B, const AtacomParams* p, void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if (B == 0) return ATACOM_OK;
  if (!q || !obs_p || !s) return ATACOM_ERR_NULL_POINTER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define ATACOM_PR_INIT(G_) \
  point_reach_slack_init_kernel<G_><<<blocks_for(B), TPB, 0, st>>>(q, obs_p, s, mask, B, as_params(p))
  switch (n_objects) {
    ATACOM_POINT_DISPATCH(1, ATACOM_PR_INIT)
    ATACOM_POINT_DISPATCH(2, ATACOM_PR_INIT)
    ATACOM_POINT_DISPATCH(3, ATACOM_PR_INIT)
    ATACOM_POINT_DISPATCH(4, ATACOM_PR_INIT)
    ATACOM_POINT_DISPATCH(6, ATACOM_PR_INIT)
    ATACOM_POINT_DISPATCH(8, ATACOM_PR_INIT)
    default: return ATACOM_ERR_BAD_DIMS;
  }
#undef ATACOM_PR_INIT
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

// ------------------------------------------------------------------ constraint statistics, fused roll-outs
int atacom_circle_constraint_stats(const float* q, const float* dq, float* per_env, double* stats, int64_t B,
                                   const AtacomParams* p, void* stream) {
  return launch_stats<CircleEnv>(q, dq, per_env, stats, B, p, stream);
}
int atacom_planar_constraint_stats(const float* q, const float* dq, float* per_env, double* stats, int64_t B,
                                   const AtacomParams* p, void* stream) {
  return launch_stats<PlanarEnv>(q, dq, per_env, stats, B, p, stream);
}
int atacom_iiwa_constraint_stats(int n, const float* q, const float* dq, float* per_env, double* stats, int64_t B,
                                 const AtacomParams* p, void* stream) {
  if (n == 6) return launch_stats<IiwaEnv<6>>(q, dq, per_env, stats, B, p, stream);
  if (n == 7) return launch_stats<IiwaEnv<7>>(q, dq, per_env, stats, B, p, stream);
  return ATACOM_ERR_BAD_DIMS;
}

int atacom_circle_rollout(float* state, float* s, const float* actions, float* rewards, double* stats,
                          uint8_t* status, int64_t B, int T, const AtacomParams* p, void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if (T < 0) return ATACOM_ERR_BAD_DIMS;
  if (B == 0 || T == 0) return ATACOM_OK;
  if (!state || !s || !actions) return ATACOM_ERR_NULL_POINTER;
  RolloutArgs a{state, s, actions, nullptr, nullptr, rewards, stats, status, B, T, 0.0};
  const ParamsT<double> Pd = widen_params<double>(as_params(p));
  circle_rollout_kernel<<<blocks_for(B), TPB, 0, static_cast<cudaStream_t>(stream)>>>(
      a, Pd, make_dual_consts<double, double>(Pd, 1, 1));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

int atacom_point_reach_rollout(int n_objects, float* state, float* s, const float* actions,
                               const float* obstacle_draws, const float* obstacle_centers, double time0,
                               float* rewards, double* stats, uint8_t* status, int64_t B, int T,
                               const AtacomParams* p, void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if (T < 0) return ATACOM_ERR_BAD_DIMS;
  if (B == 0 || T == 0) return ATACOM_OK;
  if (!state || !s || !actions) return ATACOM_ERR_NULL_POINTER;
  RolloutArgs a{state, s, actions, obstacle_draws, obstacle_centers, rewards, stats, status, B, T, time0};
  const ParamsT<double> Pd = widen_params<double>(as_params(p));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define ATACOM_PR_ROLL(G_) point_reach_rollout_kernel<G_><<<blocks_for(B), TPB, 0, st>>>(a, Pd)
  switch (n_objects) {
    ATACOM_POINT_DISPATCH(1, ATACOM_PR_ROLL)
    ATACOM_POINT_DISPATCH(2, ATACOM_PR_ROLL)
    ATACOM_POINT_DISPATCH(3, ATACOM_PR_ROLL)
    ATACOM_POINT_DISPATCH(4, ATACOM_PR_ROLL)
    ATACOM_POINT_DISPATCH(6, ATACOM_PR_ROLL)
    ATACOM_POINT_DISPATCH(8, ATACOM_PR_ROLL)
    default: return ATACOM_ERR_BAD_DIMS;
  }
#undef ATACOM_PR_ROLL
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

// generic ConstraintsSet: the (n, F, G) shapes compiled in
#define ATACOM_GENERIC_SHAPES(X) \
  X(2, 1, 1) X(3, 0, 6) X(6, 1, 11) X(7, 1, 12) X(2, 0, 4) X(4, 2, 3) X(3, 1, 0) X(2, 1, 2) X(3, 1, 3) X(4, 1, 4)

int atacom_generic_supported(int n, int F, int G) {
#define X(n_, F_, G_) if (n == n_ && F == F_ && G == G_) return 1;
  ATACOM_GENERIC_SHAPES(X)
#undef X
  return 0;
}

int atacom_generic_step(int n, int F, int G, const float* c, const float* J, const float* b, const float* dq,
                        const float* s_in, const float* alpha, float* ddq, float* s_out, uint8_t* status,
                        float* w_dbg, int64_t B, const AtacomParams* p, void* stream) {
  int rc = check_common(B, p);
  if (rc) return rc;
  if (!atacom_generic_supported(n, F, G)) return ATACOM_ERR_BAD_DIMS;
  if (B == 0) return ATACOM_OK;
  if (!c || !J || !b || !dq || !ddq || (G > 0 && (!s_in || !s_out))) return ATACOM_ERR_NULL_POINTER;
  if ((p->variant == ATACOM_VARIANT_ERROR_CORRECTION || n - F > 0) && !alpha) return ATACOM_ERR_NULL_POINTER;
  if (B == 0) return ATACOM_OK;
  StepArgs a{nullptr, dq, s_in, alpha, ddq, s_out, status, w_dbg, B, {}, 0, 0, 0, {}, nullptr, 0, 0, nullptr, nullptr,
             nullptr};
  const ParamsT<double> Pd = widen_params<double>(as_params(p));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->basis_mode != ATACOM_BASIS_LAPACK && p->basis_mode != ATACOM_BASIS_CANONICAL) return ATACOM_ERR_BAD_PARAM;
  if (p->basis_mode == ATACOM_BASIS_LAPACK) {      // the reference's own null basis, every environment
    const DualConsts<double> Kd = make_dual_consts<double, double>(Pd, F, G);
#define X(n_, F_, G_)                                                                             \
  if (n == n_ && F == F_ && G == G_)                                                              \
    atacom_generic_lapack_kernel<n_, F_, G_><<<blocks_for(B), TPB, 0, st>>>(c, J, b, a, Pd, Kd);
    ATACOM_GENERIC_SHAPES(X)
#undef X
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch();
  }
#define X(n_, F_, G_)                                                                             \
  if (n == n_ && F == F_ && G == G_)                                                              \
    atacom_generic_kernel<n_, F_, G_><<<blocks_for(B), TPB, 0, st>>>(c, J, b, a, Pd);
  ATACOM_GENERIC_SHAPES(X)
#undef X
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch();
}

}  // extern "C"

// ------------------------------------------------------------------ host-buffer entry points
// The call a user of the NumPy reference binds: host arrays in, host arrays out, for every family.  The batch is
// cut into chunks that pipeline H2D copy -> kernel -> D2H copy over a few streams.  The whole pipeline of one call
// (every copy and every launch) is captured once into a CUDA graph and replayed while the caller keeps
// passing the same buffers, batch size and parameters: one driver call per step instead of ~30.
constexpr int HOST_MAX_IN = 8, HOST_MAX_OUT = 3;

struct HostArr {
  const void* h;      // caller's host array [B, dim] (null: absent optional array)
  size_t row_bytes;   // dim * sizeof(element)
};

// One call, family-independent: its arrays and how to launch the family's kernel on a chunk of them.
struct HostCall {
  int key[4];                        // family id and shape parameters
  int n_in, n_out;
  HostArr in[HOST_MAX_IN], out[HOST_MAX_OUT];
  // kernel launch on rows [e0, e0 + nb) given the bases of the (device-visible) arrays; `din` / `dout` follow
  // in[] / out[]; chunk starts are multiples of TPB, so every chunk's rows keep the 16-byte alignment of the base
  int (*launch)(const HostCall& c, void* const* din, void* const* dout, int64_t e0, int64_t nb, const AtacomParams* p,
                cudaStream_t st);
  // zero-copy launch (one kernel on the mapped aliases of the caller's buffers), or null when the family has none
  int (*zero_copy)(const HostCall& c, void* const* din, void* const* dout, int64_t B, const AtacomParams* p,
                   struct AtacomHostCtx* ctx);
};

struct AtacomHostCtx {
  int64_t max_B;
  int chunks;
  unsigned char* arena;      // device staging area, grown on demand (never while capturing)
  size_t arena_bytes;
  cudaStream_t streams[4];
  cudaEvent_t fork, join[4];
  // cached graph and the call it was captured for
  cudaGraphExec_t exec;
  int key[4];
  const void* key_ptr[HOST_MAX_IN + HOST_MAX_OUT];
  int64_t key_B;
  int key_launches;
  AtacomParams key_params;
  int mode;
  // zero-copy path: counters of the ordered admission of the bulk loads and the number of warps admitted at once
  uint32_t* gate;
  int zc_window, zc_tpb;
};

// Device-visible alias of a page-locked, mapped host pointer (nullptr if it is not one).
static void* mapped_alias(const void* h) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

static size_t round256(size_t v) { return (v + 255) / 256 * 256; }

// The staged pipeline of one call; `dout_direct` non-null (hybrid): device aliases of the caller's mapped output
// buffers — the kernels store their results there directly (over PCIe) and the D2H copies disappear.
static int host_pipeline(AtacomHostCtx* c, const HostCall& call, int64_t B, const AtacomParams* p, int n_streams,
                         void* const* dout_direct) {
  void* din[HOST_MAX_IN];
  void* dout[HOST_MAX_OUT];
  size_t off = 0;
  for (int i = 0; i < call.n_in; ++i) {
    din[i] = call.in[i].h ? c->arena + off : nullptr;
    off += round256(call.in[i].row_bytes * static_cast<size_t>(B));
  }
  for (int i = 0; i < call.n_out; ++i) {
    dout[i] = call.out[i].h ? (dout_direct ? dout_direct[i] : c->arena + off) : nullptr;
    off += round256(call.out[i].row_bytes * static_cast<size_t>(B));
  }
  int64_t per = (B + c->chunks - 1) / c->chunks;
  per = (per + TPB - 1) / TPB * TPB;
  int rc = ATACOM_OK;
  int ci = 0;
  for (int64_t e0 = 0; e0 < B && rc == ATACOM_OK; e0 += per, ++ci) {
    const int64_t nb = (B - e0) < per ? (B - e0) : per;
    cudaStream_t st = c->streams[ci % n_streams];
    for (int i = 0; i < call.n_in; ++i)
      if (call.in[i].h)
        cudaMemcpyAsync(static_cast<unsigned char*>(din[i]) + e0 * call.in[i].row_bytes,
                        static_cast<const unsigned char*>(call.in[i].h) + e0 * call.in[i].row_bytes,
                        call.in[i].row_bytes * nb, cudaMemcpyHostToDevice, st);
    rc = call.launch(call, din, dout, e0, nb, p, st);
    if (dout_direct) continue;
    for (int i = 0; i < call.n_out; ++i)
      if (call.out[i].h)
        cudaMemcpyAsync(static_cast<unsigned char*>(const_cast<void*>(call.out[i].h)) + e0 * call.out[i].row_bytes,
                        static_cast<unsigned char*>(dout[i]) + e0 * call.out[i].row_bytes, call.out[i].row_bytes * nb,
                        cudaMemcpyDeviceToHost, st);
  }
  return rc;
}

static int host_call(AtacomHostCtx* c, const HostCall& call, int64_t B, const AtacomParams* p) {
  if (!c || !p) return ATACOM_ERR_NULL_POINTER;
  if (B < 0 || B > c->max_B) return ATACOM_ERR_BAD_DIMS;
  int rc = check_common(B, p);
  if (rc) return rc;
  if (B == 0) return ATACOM_OK;
  NvtxRange range("atacom_step_host");
  void* alias_out[HOST_MAX_OUT] = {};
  void* const* dout_direct = nullptr;
  if (c->mode == ATACOM_HOST_HYBRID) {
    // hybrid: inputs through the copy engines (large PCIe reads), outputs stored by the kernels straight into
    // the caller's mapped buffers — no D2H stage at the end of the pipeline
    for (int i = 0; i < call.n_out; ++i) {
      if (!call.out[i].h) continue;
      alias_out[i] = mapped_alias(call.out[i].h);
      if (!alias_out[i]) return ATACOM_ERR_BAD_PARAM;
    }
    dout_direct = alias_out;
  }
  if (c->mode == ATACOM_HOST_ZERO_COPY || c->mode == ATACOM_HOST_AUTO) {
    // zero-copy: one launch on the device aliases of the caller's buffers; loads and stores cross PCIe
    bool all = call.zero_copy != nullptr;
    void* zin[HOST_MAX_IN] = {};
    void* zout[HOST_MAX_OUT] = {};
    for (int i = 0; i < call.n_in && all; ++i)
      if (call.in[i].h) all = (zin[i] = mapped_alias(call.in[i].h)) != nullptr;
    for (int i = 0; i < call.n_out && all; ++i)
      if (call.out[i].h) all = (zout[i] = mapped_alias(call.out[i].h)) != nullptr;
    if (!all && c->mode == ATACOM_HOST_ZERO_COPY) return ATACOM_ERR_BAD_PARAM;
    if (all) {
      rc = call.zero_copy(call, zin, zout, B, p, c);
      if (rc != ATACOM_OK) return rc;
      if (cudaStreamSynchronize(c->streams[0]) != cudaSuccess) return ATACOM_ERR_CUDA;
      return cudaGetLastError() == cudaSuccess ? ATACOM_OK : ATACOM_ERR_CUDA;
    }
  }
  // staging area (grown outside any capture; a larger call invalidates the cached graph, whose nodes point into it)
  size_t need = 0;
  for (int i = 0; i < call.n_in; ++i) need += round256(call.in[i].row_bytes * static_cast<size_t>(B));
  for (int i = 0; i < call.n_out; ++i) need += round256(call.out[i].row_bytes * static_cast<size_t>(B));
  if (need > c->arena_bytes) {
    if (c->exec) {
      cudaGraphExecDestroy(c->exec);
      c->exec = nullptr;
    }
    cudaStreamSynchronize(c->streams[0]);
    cudaFree(c->arena);
    c->arena = nullptr;
    c->arena_bytes = 0;
    if (cudaMalloc(&c->arena, need) != cudaSuccess) {
      cudaGetLastError();
      return ATACOM_ERR_CUDA;
    }
    c->arena_bytes = need;
  }
  const void* key[HOST_MAX_IN + HOST_MAX_OUT] = {};
  for (int i = 0; i < call.n_in; ++i) key[i] = call.in[i].h;
  for (int i = 0; i < call.n_out; ++i) key[HOST_MAX_IN + i] = call.out[i].h;
  bool hit = c->exec != nullptr && c->key_B == B && memcmp(c->key, call.key, sizeof(c->key)) == 0 &&
             memcmp(&c->key_params, p, sizeof(AtacomParams)) == 0 && memcmp(c->key_ptr, key, sizeof(key)) == 0;
  if (!hit) {
    if (c->exec) {
      cudaGraphExecDestroy(c->exec);
      c->exec = nullptr;
    }
    // capture: stream 0 is the origin, the others fork from it and join back
    cudaGraph_t graph = nullptr;
    const int ns = c->chunks < 4 ? c->chunks : 4;
    bool ok = cudaStreamBeginCapture(c->streams[0], cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      cudaEventRecord(c->fork, c->streams[0]);
      for (int i = 1; i < ns; ++i) cudaStreamWaitEvent(c->streams[i], c->fork, 0);
      const int64_t before = g_launches.load(std::memory_order_relaxed);
      rc = host_pipeline(c, call, B, p, ns, dout_direct);
      c->key_launches = static_cast<int>(g_launches.exchange(before, std::memory_order_relaxed) - before);   // captured, not run
      for (int i = 1; i < ns; ++i) {
        cudaEventRecord(c->join[i], c->streams[i]);
        cudaStreamWaitEvent(c->streams[0], c->join[i], 0);
      }
      ok = cudaStreamEndCapture(c->streams[0], &graph) == cudaSuccess && graph != nullptr && rc == ATACOM_OK;
    }
    if (ok) ok = cudaGraphInstantiate(&c->exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      c->exec = nullptr;
      return rc != ATACOM_OK ? rc : ATACOM_ERR_CUDA;
    }
    memcpy(c->key, call.key, sizeof(c->key));
    memcpy(c->key_ptr, key, sizeof(key));
    c->key_B = B;
    c->key_params = *p;
  }
  if (cudaGraphLaunch(c->exec, c->streams[0]) != cudaSuccess) return ATACOM_ERR_CUDA;
  g_launches.fetch_add(c->key_launches, std::memory_order_relaxed);   // kernel launches inside the replayed graph
  if (cudaStreamSynchronize(c->streams[0]) != cudaSuccess) return ATACOM_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? ATACOM_OK : ATACOM_ERR_CUDA;
}

// ---- the families' launches on a chunk.  Step families: in = { q, dq, s_in, alpha }, out = { ddq, s_out, status }.
template <class T> static T* at_row(void* base, int64_t e0, size_t row_bytes) {
  return base ? reinterpret_cast<T*>(static_cast<unsigned char*>(base) + e0 * row_bytes) : nullptr;
}

template <class Env, int IO>
static int host_launch_step(const HostCall& c, void* const* din, void* const* dout, int64_t e0, int64_t nb,
                            const AtacomParams* p, cudaStream_t st) {
  return launch_step<Env, IO>(at_row<const float>(din[0], e0, c.in[0].row_bytes), at_row<const float>(din[1], e0, c.in[1].row_bytes),
                              at_row<const float>(din[2], e0, c.in[2].row_bytes), at_row<const float>(din[3], e0, c.in[3].row_bytes),
                              at_row<float>(dout[0], e0, c.out[0].row_bytes), at_row<float>(dout[1], e0, c.out[1].row_bytes),
                              at_row<uint8_t>(dout[2], e0, 1), nullptr, nb, p, st);
}

template <class Env>
static int host_zero_copy_step(const HostCall&, void* const* din, void* const* dout, int64_t B, const AtacomParams* p,
                               AtacomHostCtx* ctx) {
  constexpr int HOST_IO = ATACOM_STEP_STAGED_IO != 0 ? 1 : 0;
  using D = typename Env::D;
  // the reference-exact null basis takes two passes: both in one launch, so that every row crosses PCIe once each way
  static const bool fused = getenv("ATACOM_ZC_FUSED") == nullptr || atoi(getenv("ATACOM_ZC_FUSED")) != 0;
  if constexpr (D::k > 1) {
    if (fused && p->basis_mode == ATACOM_BASIS_LAPACK && p->variant == ATACOM_VARIANT_ATACOM)
      return launch_zc_fused<Env>(static_cast<const float*>(din[0]), static_cast<const float*>(din[1]),
                                  static_cast<const float*>(din[2]), static_cast<const float*>(din[3]),
                                  static_cast<float*>(dout[0]), static_cast<float*>(dout[1]),
                                  static_cast<uint8_t*>(dout[2]), B, p, ctx->streams[0], ctx->gate, ctx->zc_window);
  }
  return launch_step<Env, HOST_IO>(static_cast<const float*>(din[0]), static_cast<const float*>(din[1]),
                                   static_cast<const float*>(din[2]), static_cast<const float*>(din[3]),
                                   static_cast<float*>(dout[0]), static_cast<float*>(dout[1]),
                                   static_cast<uint8_t*>(dout[2]), nullptr, B, p, ctx->streams[0], nullptr, 0, 0, nullptr,
                                   nullptr, 0, ctx->gate, ctx->zc_window, ctx->zc_tpb, /*mirror_inputs=*/true);
}

template <class Env>
static int step_host(AtacomHostCtx* c, int family_id, const float* q, const float* dq, const float* s_in,
                     const float* alpha, float* ddq, float* s_out, uint8_t* status, int64_t B, const AtacomParams* p) {
  using D = typename Env::D;
  if (!c || !p) return ATACOM_ERR_NULL_POINTER;
  const int na = p->variant == ATACOM_VARIANT_ERROR_CORRECTION ? D::n : D::k;
  if (!q || !dq || !ddq || (D::G > 0 && (!s_in || !s_out)) || (na > 0 && !alpha)) return ATACOM_ERR_NULL_POINTER;
  // every kernel instantiation the call may use opts in to its shared memory now, not while capturing
  constexpr int HOST_IO = ATACOM_STEP_STAGED_IO != 0 ? 1 : 0;
  constexpr bool CB = D::k > 1;
  if (!configure_step_kernel<Env, ATACOM_STEP_DEVICE_IO>() || !configure_step_kernel<Env, HOST_IO>() ||
      !configure_step_kernel<Env, 2>() || !configure_step_kernel<Env, ATACOM_STEP_DEVICE_IO, CB>() ||
      !configure_step_kernel<Env, HOST_IO, CB>() || !configure_step_kernel<Env, 2, CB>() || !configure_fix_kernel<Env>())
    return ATACOM_ERR_CUDA;
  step_block_size(1);
  keep_scratch_pool_warm();      // (touches the device's memory pool: not allowed while a stream is capturing either)
  HostCall call = {};
  call.key[0] = family_id;
  call.key[1] = D::n;
  call.key[2] = c->mode;
  call.n_in = 4;
  call.n_out = 3;
  call.in[0] = {q, sizeof(float) * D::n};
  call.in[1] = {dq, sizeof(float) * D::n};
  call.in[2] = {D::G > 0 ? s_in : nullptr, sizeof(float) * D::G};
  call.in[3] = {na > 0 ? alpha : nullptr, sizeof(float) * static_cast<size_t>(na)};
  call.out[0] = {ddq, sizeof(float) * D::n};
  call.out[1] = {D::G > 0 ? s_out : nullptr, sizeof(float) * D::G};
  call.out[2] = {status, 1};
  // staged: rows by direct loads / stores; hybrid: results leave as bulk stores over PCIe
  call.launch = c->mode == ATACOM_HOST_HYBRID ? host_launch_step<Env, 2> : host_launch_step<Env, ATACOM_STEP_DEVICE_IO>;
  call.zero_copy = host_zero_copy_step<Env>;
  return host_call(c, call, B, p);
}

// point reach: in = { q, dq, obs_p, obs_dp, s_in, action }, out = { w, s_out, status }
static int host_launch_point(const HostCall& c, void* const* din, void* const* dout, int64_t e0, int64_t nb,
                             const AtacomParams* p, cudaStream_t st) {
  return atacom_point_reach_step(c.key[1], at_row<const float>(din[0], e0, c.in[0].row_bytes),
                                 at_row<const float>(din[1], e0, c.in[1].row_bytes), at_row<const float>(din[2], e0, c.in[2].row_bytes),
                                 at_row<const float>(din[3], e0, c.in[3].row_bytes), at_row<const float>(din[4], e0, c.in[4].row_bytes),
                                 at_row<const float>(din[5], e0, c.in[5].row_bytes), at_row<float>(dout[0], e0, c.out[0].row_bytes),
                                 at_row<float>(dout[1], e0, c.out[1].row_bytes), at_row<uint8_t>(dout[2], e0, 1), nullptr, nb, p, st);
}

// generic ConstraintsSet: in = { c, J, b, dq, s_in, alpha }, out = { ddq, s_out, status }
static int host_launch_generic(const HostCall& c, void* const* din, void* const* dout, int64_t e0, int64_t nb,
                               const AtacomParams* p, cudaStream_t st) {
  return atacom_generic_step(c.key[1], c.key[2], c.key[3], at_row<const float>(din[0], e0, c.in[0].row_bytes),
                             at_row<const float>(din[1], e0, c.in[1].row_bytes), at_row<const float>(din[2], e0, c.in[2].row_bytes),
                             at_row<const float>(din[3], e0, c.in[3].row_bytes), at_row<const float>(din[4], e0, c.in[4].row_bytes),
                             at_row<const float>(din[5], e0, c.in[5].row_bytes), at_row<float>(dout[0], e0, c.out[0].row_bytes),
                             at_row<float>(dout[1], e0, c.out[1].row_bytes), at_row<uint8_t>(dout[2], e0, 1), nullptr, nb, p, st);
}

extern "C" {

int atacom_host_ctx_create(AtacomHostCtx** out, int64_t max_B, int chunks) {
  if (!out) return ATACOM_ERR_NULL_POINTER;
  if (max_B <= 0 || chunks < 1 || chunks > 64) return ATACOM_ERR_BAD_DIMS;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ATACOM_ERR_NO_DEVICE;
  AtacomHostCtx* c = new (std::nothrow) AtacomHostCtx();   // value-initialised: every member starts out null
  if (!c) return ATACOM_ERR_CUDA;
  c->max_B = max_B;
  c->chunks = chunks;
  bool ok = cudaMalloc(&c->gate, 16) == cudaSuccess && cudaMemset(c->gate, 0, 16) == cudaSuccess;
  // 96 warps = 336 KB of reads in flight: enough to keep PCIe busy, few enough that the loads arrive block after
  // block (measured sweep in DESIGN.md section 6).  ATACOM_ZC_WINDOW overrides; 0 = all loads at once.
  c->zc_window = 96;
  if (const char* f = getenv("ATACOM_ZC_WINDOW")) c->zc_window = atoi(f) > 0 ? atoi(f) : 0;
  // block size of the zero-copy launch: small blocks, so that what is left to compute after the last bytes have
  // crossed PCIe is short (0: the device path's choice); ATACOM_ZC_TPB overrides
  c->zc_tpb = 128;
  if (const char* f = getenv("ATACOM_ZC_TPB")) c->zc_tpb = atoi(f);
  for (int i = 0; i < 4 && ok; ++i) ok = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&c->join[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    atacom_host_ctx_destroy(c);   // frees whatever was allocated
    return ATACOM_ERR_CUDA;
  }
  *out = c;
  return ATACOM_OK;
}

int atacom_host_ctx_set_mode(AtacomHostCtx* c, int mode) {
  if (!c) return ATACOM_ERR_NULL_POINTER;
  if (mode != ATACOM_HOST_AUTO && mode != ATACOM_HOST_STAGED && mode != ATACOM_HOST_ZERO_COPY &&
      mode != ATACOM_HOST_HYBRID)
    return ATACOM_ERR_BAD_PARAM;
  if (mode != c->mode && c->exec) {   // the cached pipeline was captured for the old data path
    cudaGraphExecDestroy(c->exec);
    c->exec = nullptr;
  }
  c->mode = mode;
  return ATACOM_OK;
}

int atacom_host_ctx_destroy(AtacomHostCtx* c) {
  if (!c) return ATACOM_OK;
  if (c->exec) cudaGraphExecDestroy(c->exec);
  cudaFree(c->arena);     // cudaFree(nullptr) is a no-op
  cudaFree(c->gate);
  for (int i = 0; i < 4; ++i) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
  if (c->fork) cudaEventDestroy(c->fork);
  for (int i = 0; i < 4; ++i) if (c->join[i]) cudaEventDestroy(c->join[i]);
  cudaGetLastError();
  delete c;
  return ATACOM_OK;
}

int atacom_circle_step_host(AtacomHostCtx* c, const float* q, const float* dq, const float* s_in, const float* alpha,
                            float* ddq, float* s_out, uint8_t* status, int64_t B, const AtacomParams* p) {
  return step_host<CircleEnv>(c, 1, q, dq, s_in, alpha, ddq, s_out, status, B, p);
}

int atacom_planar_step_host(AtacomHostCtx* c, const float* q, const float* dq, const float* s_in, const float* alpha,
                            float* ddq, float* s_out, uint8_t* status, int64_t B, const AtacomParams* p) {
  return step_host<PlanarEnv>(c, 2, q, dq, s_in, alpha, ddq, s_out, status, B, p);
}

int atacom_iiwa_step_host(AtacomHostCtx* c, int n, const float* q, const float* dq, const float* s_in,
                          const float* alpha, float* ddq, float* s_out, uint8_t* status, int64_t B,
                          const AtacomParams* p) {
  if (n == 6) return step_host<IiwaEnv<6>>(c, 3, q, dq, s_in, alpha, ddq, s_out, status, B, p);
  if (n == 7) return step_host<IiwaEnv<7>>(c, 3, q, dq, s_in, alpha, ddq, s_out, status, B, p);
  return ATACOM_ERR_BAD_DIMS;
}

int atacom_point_reach_step_host(AtacomHostCtx* c, int n_objects, const float* q, const float* dq, const float* obs_p,
                                 const float* obs_dp, const float* s_in, const float* action, float* w, float* s_out,
                                 uint8_t* status, int64_t B, const AtacomParams* p) {
  if (!c || !p || !q || !dq || !obs_p || !obs_dp || !s_in || !action || !w || !s_out) return ATACOM_ERR_NULL_POINTER;
  if (n_objects < 1 || n_objects > 8) return ATACOM_ERR_BAD_DIMS;
  if (c->mode == ATACOM_HOST_ZERO_COPY || c->mode == ATACOM_HOST_HYBRID) return ATACOM_ERR_BAD_PARAM;   // staged only
  const size_t G = static_cast<size_t>(n_objects);
  HostCall call = {};
  call.key[0] = 4;
  call.key[1] = n_objects;
  call.n_in = 6;
  call.n_out = 3;
  call.in[0] = {q, 8};
  call.in[1] = {dq, 8};
  call.in[2] = {obs_p, 8 * G};
  call.in[3] = {obs_dp, 8 * G};
  call.in[4] = {s_in, 4 * G};
  call.in[5] = {action, 8};
  call.out[0] = {w, 8};
  call.out[1] = {s_out, 4 * G};
  call.out[2] = {status, 1};
  call.launch = host_launch_point;
  return host_call(c, call, B, p);
}

int atacom_generic_step_host(AtacomHostCtx* c, int n, int F, int G, const float* cc, const float* J, const float* b,
                             const float* dq, const float* s_in, const float* alpha, float* ddq, float* s_out,
                             uint8_t* status, int64_t B, const AtacomParams* p) {
  if (!c || !p) return ATACOM_ERR_NULL_POINTER;
  if (!atacom_generic_supported(n, F, G)) return ATACOM_ERR_BAD_DIMS;
  const int na = p->variant == ATACOM_VARIANT_ERROR_CORRECTION ? n : n - F;
  if (!cc || !J || !b || !dq || !ddq || (G > 0 && (!s_in || !s_out)) || (na > 0 && !alpha)) return ATACOM_ERR_NULL_POINTER;
  if (c->mode == ATACOM_HOST_ZERO_COPY || c->mode == ATACOM_HOST_HYBRID) return ATACOM_ERR_BAD_PARAM;   // staged only
  const size_t C = static_cast<size_t>(F + G);
  HostCall call = {};
  call.key[0] = 5;
  call.key[1] = n;
  call.key[2] = F;
  call.key[3] = G;
  call.n_in = 6;
  call.n_out = 3;
  call.in[0] = {cc, 4 * C};
  call.in[1] = {J, 4 * C * static_cast<size_t>(n)};
  call.in[2] = {b, 4 * C};
  call.in[3] = {dq, 4 * static_cast<size_t>(n)};
  call.in[4] = {G > 0 ? s_in : nullptr, 4 * static_cast<size_t>(G)};
  call.in[5] = {na > 0 ? alpha : nullptr, 4 * static_cast<size_t>(na)};
  call.out[0] = {ddq, 4 * static_cast<size_t>(n)};
  call.out[1] = {G > 0 ? s_out : nullptr, 4 * static_cast<size_t>(G)};
  call.out[2] = {status, 1};
  call.launch = host_launch_generic;
  return host_call(c, call, B, p);
}

int atacom_spin_timeouts(unsigned* out) {
  if (!out) return ATACOM_ERR_NULL_POINTER;
  return cudaMemcpyFromSymbol(out, g_spin_timeouts, sizeof(unsigned)) == cudaSuccess ? ATACOM_OK : ATACOM_ERR_CUDA;
}


}  // extern "C"
