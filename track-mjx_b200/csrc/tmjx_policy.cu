/*
 * tmjx_policy.cu — in-loop intention-network inference for the tracking env (BASELINE configs[2]).
 *
 * Replaces, for the acting loop only (no gradients),
 *   IntentionNetwork.__call__ / Encoder / Decoder / reparameterize   reference track_mjx/agent/mlp_ppo/intention_network.py:14-142
 *   make_inference_fn().policy (sample, log_prob, postprocess)       reference track_mjx/agent/mlp_ppo/ppo_networks.py:34-100
 *   running_statistics.normalize                                     reference track_mjx/agent/masked_running_statistics.py:217-236
 *   NormalTanhDistribution (upstream brax 0.12.3 training/distribution.py: scale = softplus(raw) + 0.001, tanh bijector)
 *
 * B200 mapping.  The ACTING path (tmjx_policy_act, tmjx_value_apply) is one persistent launch for the whole network, tmjx_chain.cuh.
 * This file holds the per-layer form -- the training forward / backward GEMMs of tmjx_train.cuh and the A/B reference of the fused
 * launch (TMJX_POLICY_FUSED=0): every Dense layer is one tcgen05 GEMM over the environment batch -- `tcgen05.mma.cta_group::1.kind::tf32`
 * (fp32 operands read as TF32 by the tensor core, fp32 accumulation in TMEM; XLA's default fp32 matmul precision on
 * NVIDIA GPUs is TF32 as well).  Three kernels live here, newest first:
 *   linear_tf32_tma_kernel<BN>  (default)  256 x BN tile = two M = 128 accumulators sharing one B tile, operands moved by TMA
 *                                (cp.async.bulk.tensor.2d, 128-byte swizzle, expect_tx mbarriers), one producer thread, one MMA
 *                                thread, 16-warp tcgen05.ld epilogue (bias, SiLU) with coalesced stores through a swizzled smem transpose;
 *   linear_tf32_v2_kernel<BN>   (TMJX_POLICY_V1=2)  the same tile with warp-specialised cp.async producers;
 *   linear_tf32_kernel          (TMJX_POLICY_V1=1)  128 x 128 tile, block-synchronous cp.async ring.
 * The older two are kept as A/B references (DESIGN.md 3b has the measurements that led from one to the next).  LayerNorm,
 * observation normalisation, the reparameterised latent and the tanh-normal action head are small row-wise kernels around
 * the GEMMs.  All launches are enqueued on the caller's stream; nothing is allocated per call.
 *
 * The learner-side kernels of a PPO update live here too: the value-network forward on the same
 * GEMM (tmjx_value_*), GAE (tmjx_gae, losses.py:39-101), the PPO loss head with its gradient seeds (tmjx_ppo_loss_head,
 * losses.py:154-245), the observation-normaliser update (tmjx_running_stats_*, masked_running_statistics.py:80-214) and the
 * optimiser step (tmjx_adam_step, ppo.py:517-520).  They are HBM-bound streaming kernels with fixed-order two-level reductions.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/tmjx.h"

namespace tmjx_policy {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3, THREADS = 256;
constexpr int kStageBytes = (BM + BN) * BK * 4;
constexpr int kSmemBytes = STAGES * kStageBytes;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// K-major, no-swizzle shared-memory matrix descriptor (sm_100 format, version 1): 8 x 16 B core matrices, LBO = byte
// distance between the two core matrices an MMA reads along K, SBO = byte distance between 8-row groups along M / N
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float silu(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.f + __expf(-v)); }   // MUFU.EX2 + MUFU.RCP, ~2 ulp
constexpr int kTmaThreads = 512;

// Y[M, ldy] (columns n0 .. n0 + 127 of this block) = act(X[M, K] * Wt[N, K]^T + bias).  X / Wt row pitches ldx / ldw are
// multiples of BK floats and zero padded; Wt and bias are padded to a multiple of BN rows.
__global__ void __launch_bounds__(THREADS) linear_tf32_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ Wt, int ldw,
                                                              const float* __restrict__ bias, float* __restrict__ Y, int ldy, int M, int Kpad,
                                                              int act, int desc_swap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_free[STAGES];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nk = Kpad / BK;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(uint32_t(BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bar_free[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  // one stage = A tile then B tile, each stored chunk-major: 16-byte K chunk kc of row r at (kc * ROWS + r) * 16
  auto load_stage = [&](int kt, int s) {
    const uint32_t sa = sbase + s * kStageBytes, sb = sa + BM * BK * 4;
#pragma unroll
    for (int i = 0; i < (BM + BN) * (BK / 4) / THREADS; ++i) {
      const int c = tid + THREADS * i;
      if (c < BM * (BK / 4)) {
        const int r = c >> 3, kc = c & 7;
        const bool ok = m0 + r < M;
        const float* src = X + size_t(ok ? m0 + r : 0) * ldx + kt * BK + kc * 4;
        cp_async16(sa + (kc * BM + r) * 16, src, ok ? 16u : 0u);
      } else {
        const int c2 = c - BM * (BK / 4), r = c2 >> 3, kc = c2 & 7;
        const float* src = Wt + size_t(n0 + r) * ldw + kt * BK + kc * 4;
        cp_async16(sb + (kc * BN + r) * 16, src, 16u);
      }
    }
  };
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt % STAGES;
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async wrote through the generic proxy; the MMA reads through the async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = sbase + s * kStageBytes, sb = sa + BM * BK * 4;
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {   // one MMA = K 8 (two 16-byte chunks)
        const uint32_t a_lbo = desc_swap ? 128u : uint32_t(BM * 16), a_sbo = desc_swap ? uint32_t(BM * 16) : 128u;
        const uint32_t b_lbo = desc_swap ? 128u : uint32_t(BN * 16), b_sbo = desc_swap ? uint32_t(BN * 16) : 128u;
        const uint64_t da = make_desc(sa + j * 2 * BM * 16, a_lbo, a_sbo);
        const uint64_t db = make_desc(sb + j * 2 * BN * 16, b_lbo, b_sbo);
        mma_tf32(tmem, da, db, idesc, (kt > 0 || j > 0) ? 1u : 0u);
      }
      mma_commit(&bar_free[s]);   // arrives when the MMAs above (and all earlier ones) have read their operands and finished
    }
    const int kn = kt + STAGES - 1;
    if (kn < nk) {
      const int sn = kn % STAGES;
      if (kt >= 1) mbar_wait(&bar_free[sn], uint32_t(((kt - 1) / STAGES) & 1));   // iteration kt - 1 read stage sn
      load_stage(kn, sn);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  mbar_wait(&bar_free[(nk - 1) % STAGES], uint32_t(((nk - 1) / STAGES) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (= output rows), column half w / 4
  const int lane_base = (warp & 3) * 32, col_base = (warp >> 2) * 64;
  const int row = m0 + lane_base + lane;
#pragma unroll
  for (int c = 0; c < 64; c += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + (uint32_t(lane_base) << 16) + uint32_t(col_base + c);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < M) {
      float* dst = Y + size_t(row) * ldy + n0 + col_base + c;
      const float* bb = bias + n0 + col_base + c;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = __uint_as_float(v[j]) + __ldg(bb + j);
        o.y = __uint_as_float(v[j + 1]) + __ldg(bb + j + 1);
        o.z = __uint_as_float(v[j + 2]) + __ldg(bb + j + 2);
        o.w = __uint_as_float(v[j + 3]) + __ldg(bb + j + 3);
        if (act) { o.x = silu(o.x); o.y = silu(o.y); o.z = silu(o.z); o.w = silu(o.w); }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(BN)) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------
// v2: 256 x BN output tile per CTA (two M = 128 accumulators sharing one B tile: the weight tile is fetched half as often
// and the activation tile N / BN times), warp-specialised: warps 0..6 are cp.async producers that signal a per-stage "full"
// mbarrier through cp.async.mbarrier.arrive.noinc (arrival = completion of the thread's copies), one thread of warp 7 waits on it, issues the
// eight MMAs of the K slice and tcgen05.commit's on the stage's "empty" mbarrier; there is no block barrier in the main loop.
template <int BN2>
struct V2 {
  static constexpr int BM2 = 256, STG = BN2 == 256 ? 3 : 4, NPROD = 224;
  static constexpr int kStage = (BM2 + BN2) * BK * 4;
  static constexpr int kSmem = STG * kStage;
};

template <int BN2>
__global__ void __launch_bounds__(THREADS, 1) linear_tf32_v2_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ Wt, int ldw,
                                                                    const float* __restrict__ bias, float* __restrict__ Y, int ldy, int M,
                                                                    int Kpad, int act) {
  using C2 = V2<BN2>;
  constexpr int BM2 = C2::BM2, STG = C2::STG, NPROD = C2::NPROD;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[STG], bar_empty[STG], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM2, n0 = blockIdx.y * BN2;
  const int nk = Kpad / BK;
  const uint32_t sbase = smem_u32(smem);
  constexpr uint32_t kTmemCols = 2 * BN2;   // 256 or 512: power of two

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < STG; ++s) { mbar_init(&bar_full[s], NPROD); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (tid < NPROD) {
    // ---- producers: never wait on their own loads -- `cp.async.mbarrier.arrive.noinc` makes the completion of this thread's
    // copies one of the NPROD expected arrivals of the stage's "full" barrier, so up to STG slices are in flight
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % STG;
      if (kt >= STG) mbar_wait(&bar_empty[s], uint32_t((kt / STG - 1) & 1));
      const uint32_t sa = sbase + s * C2::kStage, sb = sa + BM2 * BK * 4;
      for (int c = tid; c < (BM2 + BN2) * (BK / 4); c += NPROD) {
        if (c < BM2 * (BK / 4)) {
          const int r = c >> 3, kc = c & 7;
          const bool ok = m0 + r < M;
          cp_async16(sa + (kc * BM2 + r) * 16, X + size_t(ok ? m0 + r : 0) * ldx + kt * BK + kc * 4, ok ? 16u : 0u);
        } else {
          const int c2 = c - BM2 * (BK / 4), r = c2 >> 3, kc = c2 & 7;
          cp_async16(sb + (kc * BN2 + r) * 16, Wt + size_t(n0 + r) * ldw + kt * BK + kc * 4, 16u);
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_full[s])) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (tid == NPROD) {
    // ---- MMA issuer (one thread)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN2 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % STG;
      mbar_wait(&bar_full[s], uint32_t((kt / STG) & 1));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the landed cp.async data (generic proxy) -> tensor-core reads (async proxy)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = sbase + s * C2::kStage, sb = sa + BM2 * BK * 4;
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {
        const uint64_t db = make_desc(sb + j * 2 * BN2 * 16, uint32_t(BN2 * 16), 128u);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t da = make_desc(sa + j * 2 * BM2 * 16 + h * 128 * 16, uint32_t(BM2 * 16), 128u);
          mma_tf32(tmem + uint32_t(h * BN2), da, db, idesc, (kt > 0 || j > 0) ? 1u : 0u);
        }
      }
      mma_commit(&bar_empty[s]);
    }
    mma_commit(&bar_done);
  }
  mbar_wait(&bar_done, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncwarp();

  // ---- epilogue: warp w -> accumulator half w / 4, TMEM lanes 32 (w % 4) .. +31; thread = one output row, all BN2 columns
  const int h = warp >> 2, lane_base = (warp & 3) * 32;
  const int row = m0 + h * 128 + lane_base + lane;
#pragma unroll 1
  for (int c = 0; c < BN2; c += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + (uint32_t(lane_base) << 16) + uint32_t(h * BN2 + c);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < M) {
      float* dst = Y + size_t(row) * ldy + n0 + c;
      const float* bb = bias + n0 + c;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = __uint_as_float(v[j]) + __ldg(bb + j);
        o.y = __uint_as_float(v[j + 1]) + __ldg(bb + j + 1);
        o.z = __uint_as_float(v[j + 2]) + __ldg(bb + j + 2);
        o.w = __uint_as_float(v[j + 3]) + __ldg(bb + j + 3);
        if (act) { o.x = silu(o.x); o.y = silu(o.y); o.z = silu(o.z); o.w = silu(o.w); }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// v3 (default): the same 256 x BN tile and accumulator layout as v2, operands moved by TMA.  ncu on v2 showed the producer
// warps stalled on the scoreboard of their own LDGSTS address registers (one K slice in flight per thread, 15 GB/s per SM);
// with `cp.async.bulk.tensor.2d` ONE thread issues two bulk copies per K slice (A: 32 floats x 256 rows, B: 32 x BN) that
// land 128-byte-swizzled (CU_TENSOR_MAP_SWIZZLE_128B) and complete on the stage's "full" mbarrier (expect_tx); the MMA
// thread reads them through SWIZZLE_128B K-major descriptors (8-row groups 1024 B apart, K advance = +32 B inside the
// 128-byte atom).  Warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, all 16 warps = epilogue.
template <int BN2>
struct V3 {
  static constexpr int BM2 = 256, STG = BN2 == 256 ? 3 : 4;
  static constexpr int kABytes = BM2 * BK * 4, kBBytes = BN2 * BK * 4, kStage = kABytes + kBBytes;
  static constexpr int kSmem = STG * kStage + 1024;   // + slack for the 1024-byte alignment the swizzle needs
};
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// this thread's 32 values (row = TMEM lane, columns col .. col + 31) -> global, transposed through the warp's shared-memory tile so that
// every store instruction writes four complete 128-byte row segments (the direct form -- one 16-byte piece per lane, 32 rows per
// instruction -- made the epilogue LSU-bound: 16 k scattered requests per layer and CTA).
__device__ __forceinline__ void store_tile(const uint32_t (&v)[32], float* tile, int lane, float* __restrict__ g, int ld, size_t row0, int rows_valid,
                                           int col) {
#pragma unroll
  for (int j = 0; j < 8; ++j)   // row `lane`, chunk j -> chunk slot j ^ (lane & 7): a quarter-warp's eight 16-byte stores hit eight different bank groups
    *reinterpret_cast<float4*>(tile + lane * 32 + ((j ^ (lane & 7)) << 2)) =
        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
  __syncwarp();
  const int ch = lane & 7, rs = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rs;
    const float4 t = *reinterpret_cast<const float4*>(tile + r * 32 + ((ch ^ (r & 7)) << 2));
    if (r < rows_valid) *reinterpret_cast<float4*>(g + (row0 + r) * ld + col + ch * 4) = t;
  }
  __syncwarp();
}

// MN-major operands (the wgrad GEMMs: dW = x^T dH contracts over the ROWS of two row-major matrices, so the contraction index is the
// slow one of both operands).  The row-major matrix is read through a plain 2-D map with a 32 x 32 box, one box per 32-float MN chunk:
// 4-row groups 512 B apart along K (SBO), MN chunks 4096 B apart (LBO); the instruction descriptor's transpose bits
// (15: A, 16: B) select the MN-major read.  Replaces the two transposing copies per layer of the first backward pass.
// (32-bit MN-major operands have ONE legal shared-memory layout: 128-byte rows swizzled in 32-byte atoms over groups of FOUR rows --
//  layout type 1, TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; with the ordinary 128-byte swizzle the MMA silently produces zeros)
__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t saddr) {
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(4096 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(1) << 61);
}

template <int BN2, int kMN = 0>   // kMN bit 0: A operand MN-major, bit 1: B operand MN-major
__global__ void __launch_bounds__(kTmaThreads, 1) linear_tf32_tma_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                                                                     const float* __restrict__ bias, float* __restrict__ Y, int ldy, int M, int Kpad,
                                                                     int act, int nk_per_split = 0, size_t y_split_stride = 0) {
  using C3 = V3<BN2>;
  constexpr int BM2 = C3::BM2, STG = C3::STG;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[STG], bar_empty[STG], bar_done;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[BN2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM2, n0 = blockIdx.y * BN2;
  if (tid < BN2) sbias[tid] = __ldg(bias + n0 + tid);   // read by every epilogue thread: broadcast LDS instead of L2-latency loads (the L1 is ~30 KB here)
  // split-K (the wgrad GEMMs of the backward pass: small M x N, K = the minibatch rows): block z accumulates K slices
  // [z nk_per_split, (z + 1) nk_per_split) into its own output plane Y + z y_split_stride; the planes are summed in z order later
  int nk = Kpad / BK, kt0 = 0;
  if (nk_per_split > 0) {
    kt0 = int(blockIdx.z) * nk_per_split;
    nk = min(nk_per_split, nk - kt0);
    Y += size_t(blockIdx.z) * y_split_stride;
  }
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t kTmemCols = 2 * BN2;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < STG; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 64) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    // ---- TMA producer
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % STG;
      if (kt >= STG) mbar_wait(&bar_empty[s], uint32_t((kt / STG - 1) & 1));
      const uint32_t sa = sbase + s * C3::kStage, sb = sa + C3::kABytes;
      asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bar_full[s])),
                   "r"(uint32_t(C3::kStage))
                   : "memory");
      if (kMN & 1) {   // one 32 (K rows) x 32 (MN floats) box per MN chunk, 4096 B apart
#pragma unroll
        for (int b = 0; b < BM2 / 32; ++b) tma_load_2d(sa + b * 4096, &mapX, m0 + 32 * b, (kt0 + kt) * BK, &bar_full[s]);
      } else {
        tma_load_2d(sa, &mapX, (kt0 + kt) * BK, m0, &bar_full[s]);
      }
      if (kMN & 2) {
#pragma unroll
        for (int b = 0; b < BN2 / 32; ++b) tma_load_2d(sb + b * 4096, &mapW, n0 + 32 * b, (kt0 + kt) * BK, &bar_full[s]);
      } else {
        tma_load_2d(sb, &mapW, (kt0 + kt) * BK, n0, &bar_full[s]);
      }
    }
  } else if (tid == 32) {
    // ---- MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN2 >> 3) << 17) | (uint32_t(128 >> 4) << 24) | ((kMN & 1) ? (1u << 15) : 0u) | ((kMN & 2) ? (1u << 16) : 0u);
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % STG;
      mbar_wait(&bar_full[s], uint32_t((kt / STG) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = sbase + s * C3::kStage, sb = sa + C3::kABytes;
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {
        const uint64_t db = (kMN & 2) ? make_desc_sw128_mn(sb + j * 1024) : make_desc_sw128(sb + j * 32);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t da = (kMN & 1) ? make_desc_sw128_mn(sa + h * 128 * 128 + j * 1024) : make_desc_sw128(sa + h * 128 * 128 + j * 32);
          mma_tf32(tmem + uint32_t(h * BN2), da, db, idesc, (kt > 0 || j > 0) ? 1u : 0u);
        }
      }
      mma_commit(&bar_empty[s]);
    }
    mma_commit(&bar_done);
  }
  mbar_wait(&bar_done, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncwarp();

  // ---- epilogue, 16 warps: warp w reads TMEM lanes 32 (w % 4) .. +31 of accumulator (w / 4) % 2, column half w / 8; thread = one
  // output row.  Stores go through a per-warp 32 x 32 XOR-swizzled transpose tile (aliased onto the operand stages, which are dead once
  // bar_done has fired) so that every store instruction writes four complete 128-byte row segments: the direct form -- one 16-byte
  // piece per lane, 32 rows per instruction -- is LSU-request-bound (measured in the fused chain kernel: 16 k requests, 19 us per
  // 128 x 512 tile) and made this epilogue longer than the main loop.
  const int h = (warp >> 2) & 1, lane_base = (warp & 3) * 32, chalf = warp >> 3;
  const size_t row0 = size_t(m0) + h * 128 + lane_base;
  const int rows_valid = max(0, min(32, M - int(row0)));
  float* tile = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw))) + warp * 1024;
  const int c_beg = chalf * (BN2 / 2), c_end = (act & 4) ? c_beg : c_beg + BN2 / 2;   // bit 2: timing experiment only (skip the epilogue)
  act &= 1;
#pragma unroll 1
  for (int c = c_beg; c < c_end; c += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + (uint32_t(lane_base) << 16) + uint32_t(h * BN2 + c);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const float4* bb = reinterpret_cast<const float4*>(sbias + c);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = bb[j >> 2];
      float4 o;
      o.x = __uint_as_float(v[j]) + b4.x;
      o.y = __uint_as_float(v[j + 1]) + b4.y;
      o.z = __uint_as_float(v[j + 2]) + b4.z;
      o.w = __uint_as_float(v[j + 3]) + b4.w;
      if (act) { o.x = silu_fast(o.x); o.y = silu_fast(o.y); o.z = silu_fast(o.z); o.w = silu_fast(o.w); }
      v[j] = __float_as_uint(o.x); v[j + 1] = __float_as_uint(o.y); v[j + 2] = __float_as_uint(o.z); v[j + 3] = __float_as_uint(o.w);
    }
    store_tile(v, tile, lane, Y, ldy, row0, rows_valid, n0 + c);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// flax nn.LayerNorm (epsilon 1e-6, use_fast_variance: var = E[x^2] - E[x]^2 clipped at 0), in place; one warp per row
__global__ void layernorm_kernel(float* __restrict__ Y, int ldy, int n, const float* __restrict__ scale, const float* __restrict__ bias, int M) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  float* y = Y + size_t(row) * ldy;
  float s = 0.f, s2 = 0.f;
  for (int i = lane * 4; i < n; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(y + i);
    s += v.x + v.y + v.z + v.w;
    s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const float mean = s / float(n), var = fmaxf(0.f, s2 / float(n) - mean * mean), rstd = rsqrtf(var + 1e-6f);
  for (int i = lane * 4; i < n; i += 128) {
    float4 v = *reinterpret_cast<const float4*>(y + i);
    const float4 g = *reinterpret_cast<const float4*>(scale + i), b = *reinterpret_cast<const float4*>(bias + i);
    v.x = (v.x - mean) * rstd * g.x + b.x; v.y = (v.y - mean) * rstd * g.y + b.y;
    v.z = (v.z - mean) * rstd * g.z + b.z; v.w = (v.w - mean) * rstd * g.w + b.w;
    *reinterpret_cast<float4*>(y + i) = v;
  }
}

// (obs - mean) / std; reference part -> encoder input [M, ld_enc], proprioceptive part -> decoder input columns latent..
__global__ void obs_prep_kernel(const float* __restrict__ obs, int nobs, int nref, int latent, const float* __restrict__ mean,
                                const float* __restrict__ stdv, float* __restrict__ enc_in, int ld_enc, float* __restrict__ dec_in, int ld_dec, int M) {
  const int row = blockIdx.x;
  if (row >= M) return;
  const float* o = obs + size_t(row) * nobs;
  for (int i = threadIdx.x; i < nobs; i += blockDim.x) {
    const float v = (o[i] - mean[i]) / stdv[i];
    if (i < nref) enc_in[size_t(row) * ld_enc + i] = v;
    else dec_in[size_t(row) * ld_dec + latent + (i - nref)] = v;
  }
}

// z = mean + exp(logvar / 2) * eps (or the mean when deterministic) into decoder input columns 0 .. latent - 1
__global__ void latent_kernel(const float* __restrict__ head, int ld_head, int latent, const float* __restrict__ eps, int deterministic,
                              float* __restrict__ dec_in, int ld_dec, float* __restrict__ out_mean, float* __restrict__ out_logvar, int M) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * latent) return;
  const int row = idx / latent, j = idx % latent;
  const float mu = head[size_t(row) * ld_head + j], lv = head[size_t(row) * ld_head + latent + j];
  const float z = deterministic ? mu : mu + eps[idx] * expf(0.5f * lv);
  dec_in[size_t(row) * ld_dec + j] = z;
  if (out_mean) out_mean[idx] = mu;
  if (out_logvar) out_logvar[idx] = lv;
}

// NormalTanhDistribution: loc, scale = softplus(raw) + 0.001; raw_action = loc + scale * eps; action = tanh(raw_action);
// log_prob = sum_i [ N(raw; loc, scale) - log|d tanh / d raw| ],  log|.| = 2 (log 2 - x - softplus(-2 x))
__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__global__ void action_head_kernel(const float* __restrict__ lg, int ld_lg, int na, const float* __restrict__ eps, int deterministic,
                                   float* __restrict__ action, float* __restrict__ raw_action, float* __restrict__ log_prob,
                                   float* __restrict__ logits, int M) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* l = lg + size_t(row) * ld_lg;
  float lp = 0.f;
  for (int i = lane; i < na; i += 32) {
    const float loc = l[i], scale = softplus(l[na + i]) + 0.001f;
    const float raw = deterministic ? loc : loc + scale * eps[size_t(row) * na + i];
    action[size_t(row) * na + i] = tanhf(raw);
    if (raw_action) raw_action[size_t(row) * na + i] = raw;
    const float zn = (raw - loc) / scale;
    lp += -0.5f * zn * zn - logf(scale) - 0.91893853320467274f - 2.f * (0.69314718055994531f - raw - softplus(-2.f * raw));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
  if (log_prob && lane == 0) log_prob[row] = lp;
  if (logits) for (int i = lane; i < 2 * na; i += 32) logits[size_t(row) * 2 * na + i] = l[i];
}


// Generalised Advantage Estimation (reference losses.py:39-101 `compute_gae`): one thread per environment walks its T steps
// backwards; [T, B] arrays are read / written with consecutive threads on consecutive environments (coalesced), every
// element exactly once: 24 B per (t, env), HBM-bound.  The arithmetic follows the reference's float32 operation order with
// explicit round-to-nearest multiplies and adds (no fused multiply-add), so the result is bit-identical to the restatement.
__global__ void gae_kernel(const float* __restrict__ truncation, const float* __restrict__ termination, const float* __restrict__ rewards,
                           const float* __restrict__ values, const float* __restrict__ bootstrap, float lambda, float discount,
                           float* __restrict__ vs, float* __restrict__ advantages, int T, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float acc = 0.f, v_next = bootstrap[b], vs_next = bootstrap[b];
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = size_t(t) * B + b;
    const float tm = __fsub_rn(1.f, truncation[i]);
    const float dn = __fmul_rn(discount, __fsub_rn(1.f, termination[i]));   // discount * (1 - termination)
    const float r = rewards[i], v = values[i];
    const float delta = __fmul_rn(__fsub_rn(__fadd_rn(r, __fmul_rn(dn, v_next)), v), tm);
    acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(dn, tm), lambda), acc));
    const float vs_t = __fadd_rn(acc, v);
    advantages[i] = __fmul_rn(__fsub_rn(__fadd_rn(r, __fmul_rn(dn, vs_next)), v), tm);
    vs[i] = vs_t;
    v_next = v;
    vs_next = vs_t;
  }
}

// Value network input / output (brax make_value_network as used at ppo_networks.py:180-185): x = (obs - mean) / std into the
// K-padded GEMM input; value = column 0 of the last Dense (jnp.squeeze(..., axis=-1))
__global__ void value_prep_kernel(const float* __restrict__ obs, int obs_size, const float* __restrict__ mean, const float* __restrict__ stdv,
                                  float* __restrict__ x, int ldx, int n_env) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < obs_size; c += blockDim.x) x[size_t(r) * ldx + c] = (obs[size_t(r) * obs_size + c] - mean[c]) / stdv[c];
}
__global__ void value_out_kernel(const float* __restrict__ y, int ldy, float* __restrict__ value, int n_env) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_env) value[r] = y[size_t(r) * ldy];
}

// ---- PPO loss head (reference losses.py:104-245 after the network applications) ---------------------------------------
// Three passes over the [T, B] rollout, one warp per transition row in the two wide ones (lanes over the 38 actions and the 60
// latent dimensions, coalesced row reads, xor-shuffle sums: deterministic):
//   ppo_rows_kernel   target log-prob, entropy and latent-KL row sums + block partials of every mean that needs no normalised
//                     advantage                                                               (reads ~1.3 KB per row, HBM-bound)
//   ppo_reduce_kernel stage 1: advantage mean / std, value / entropy / KL terms; stage 2: policy term and total
//   ppo_grad_kernel   d loss / d (logits, latent mean, latent log-variance, baseline), the normalised advantages and the block
//                     partials of the clipped surrogate                                   (reads 1.3 KB, writes 0.8 KB per row)
// The gradients are what a backward pass through the policy / value networks starts from; vs and advantages carry no gradient
// (stop_gradient at losses.py:100), the bootstrap value therefore gets none.
struct PpoHyper { float entropy_cost, kl_weight, discounting, reward_scaling, gae_lambda, clipping_epsilon; int normalize_advantage; };
constexpr float kArAlpha = 0.95f, kArPriorVar = 1.f - 0.95f * 0.95f, kArInvPriorVar = 1.f / kArPriorVar;   // autoregressive latent prior, losses.py:201-202
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float log_det_tanh(float x) { return 2.f * (0.69314718055994531f - x - softplus(-2.f * x)); }

__global__ void ppo_prep_kernel(const float* __restrict__ reward, const float* __restrict__ discount, const float* __restrict__ truncation,
                                float reward_scaling, float* __restrict__ rewards, float* __restrict__ termination, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rewards[i] = reward[i] * reward_scaling;                          // losses.py:157
  termination[i] = (1.f - discount[i]) * (1.f - truncation[i]);     // :159
}

// Grid of the two row kernels: persistent, kPpoBlocks blocks of 8 warps, warp w of the grid takes rows w, w + nwarps, ...; the
// per-block partial sums (double) are merged by one small block afterwards -- the summation order depends on this constant
// only, never on the device or the batch
constexpr int kPpoBlocks = 148 * 8, kPpoWarps = 8, kPpoSums = 6;   // sums: adv, adv^2, entropy, kl_0, kl_t, (vs - baseline)^2
// Lanes per transition row: with 38 actions a full warp per row idles 26 of 64 lane slots; 8 lanes per row (four rows per warp)
// use 38 of 40 action slots and 60 of 64 latent slots
constexpr int kPpoLanes = 8, kPpoRowsPerWarp = 32 / kPpoLanes;
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = kPpoLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void block_partials(double* acc, int nacc, double* __restrict__ partial) {
  __shared__ double sh[kPpoWarps][kPpoSums];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < nacc; ++k) {                                  // the row-group leaders of the warp hold the sums
    double v = acc[k];
#pragma unroll
    for (int o = kPpoLanes; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < nacc) {
    double t = 0.0;
    for (int w = 0; w < kPpoWarps; ++w) t += sh[w][threadIdx.x];
    partial[size_t(blockIdx.x) * kPpoSums + threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32 * kPpoWarps) ppo_rows_kernel(const float* __restrict__ logits, const float* __restrict__ raw_action,
                                                       const float* __restrict__ eps, const float* __restrict__ lat_mean,
                                                       const float* __restrict__ lat_logvar, const float* __restrict__ adv,
                                                       const float* __restrict__ vs, const float* __restrict__ baseline, int T, int B, int A,
                                                       int L, float* __restrict__ logp, double* __restrict__ partial) {
  const size_t nrow = size_t(T) * B, stride = size_t(gridDim.x) * kPpoWarps * kPpoRowsPerWarp;
  const int lane = threadIdx.x & (kPpoLanes - 1);
  double acc[kPpoSums] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const float kLogPriorVar = logf(kArPriorVar);
  // every lane of a warp runs the same number of iterations (the shuffles need all of them); rows past the end are skipped
  for (size_t row0 = (size_t(blockIdx.x) * kPpoWarps + (threadIdx.x >> 5)) * kPpoRowsPerWarp; row0 < nrow; row0 += stride) {
    const size_t row = row0 + ((threadIdx.x & 31) / kPpoLanes);
    const bool live = row < nrow;
    float lp = 0.f, en = 0.f, k = 0.f;
    const bool first = row < size_t(B);                             // t == 0: standard normal prior, else AR(1) prior
    if (live) {
      const float* lg = logits + row * 2 * A;
      for (int i = lane; i < A; i += kPpoLanes) {
        const float loc = lg[i], scale = softplus(lg[A + i]) + 0.001f, raw = raw_action[row * A + i];
        const float z = (raw - loc) / scale, ls = logf(scale);
        lp += -0.5f * z * z - ls - 0.91893853320467274f - log_det_tanh(raw);
        en += 0.5f + 0.91893853320467274f + ls + log_det_tanh(fmaf(scale, eps[row * A + i], loc));
      }
      const float *mu = lat_mean + row * L, *lv = lat_logvar + row * L, *mp = mu - size_t(B) * L;
      for (int j = lane; j < L; j += kPpoLanes) {
        const float m = mu[j], v = lv[j];
        if (first) {
          k += 1.f + v - m * m - expf(v);                           // * -0.5 below (:206-208)
        } else {
          const float dm = kArAlpha * mp[j] - m;
          k += (expf(v) + dm * dm) * kArInvPriorVar - 1.f + (kLogPriorVar - v);   // * 0.5 below (:221-226)
        }
      }
    }
    lp = group_sum(lp); en = group_sum(en); k = group_sum(k);
    if (lane == 0 && live) {
      logp[row] = lp;
      const double a = adv[row], e = double(vs[row]) - double(baseline[row]);
      acc[0] += a; acc[1] += a * a; acc[2] += en;
      if (first) acc[3] += -0.5 * double(k); else acc[4] += 0.5 * double(k);
      acc[5] += e * e;
    }
  }
  block_partials(acc, kPpoSums, partial);
}

// out: [0] total [1] policy [2] value [3] latent KL [4] entropy loss [5] advantage mean [6] advantage std [7] 1 / (std + 1e-8).
// Stage 1 (after ppo_rows_kernel) fills [2..7]; stage 2 (after ppo_grad_kernel, whose partial[.][0] is the surrogate sum) [0..1]
__global__ void __launch_bounds__(256) ppo_reduce_kernel(const double* __restrict__ partial, int nblk, int stage, int T, int B, int L,
                                                         PpoHyper hp, float* __restrict__ out) {
  __shared__ double sh[256];
  double tot[kPpoSums];
  const int nsum = (stage == 1 || stage == 3) ? kPpoSums : (stage == 0 ? 2 : 1);
  for (int k = 0; k < nsum; ++k) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) s += partial[size_t(b) * kPpoSums + k];
    __syncthreads();
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if (int(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    tot[k] = sh[0];
  }
  if (threadIdx.x != 0) return;
  const double n = double(T) * B;
  if (stage == 0) {                                                 // one-pass path: advantage statistics only (slots 0, 1)
    float mean = 0.f, stdv = 1.f, inv = 1.f;
    if (hp.normalize_advantage) {
      const double m = tot[0] / n;
      mean = float(m);
      stdv = float(sqrt(fmax(tot[1] / n - m * m, 0.0)));
      inv = 1.f / (stdv + 1e-8f);
    }
    out[5] = mean; out[6] = stdv; out[7] = inv;
  } else if (stage == 3) {                                          // one-pass path: every loss term (slot 0 = surrogate sum, 2..5 as stage 1)
    const float kl0 = float(tot[3] / (double(B) * L));
    float klat = hp.kl_weight * kl0;
    if (T > 1) klat = hp.kl_weight * ((kl0 + float(tot[4] / (double(T - 1) * B * L)) * float(T - 1)) / float(T));
    out[2] = float(tot[5] / n) * 0.5f * 0.5f;
    out[3] = klat;
    out[4] = hp.entropy_cost * -float(tot[2] / n);
    out[1] = -float(tot[0] / n);
    out[0] = out[1] + out[2] + out[4] + out[3];
  } else if (stage == 1) {
    float mean = 0.f, stdv = 1.f, inv = 1.f;
    if (hp.normalize_advantage) {                                   // :175-176 (population standard deviation)
      const double m = tot[0] / n;
      mean = float(m);
      stdv = float(sqrt(fmax(tot[1] / n - m * m, 0.0)));
      inv = 1.f / (stdv + 1e-8f);
    }
    const float kl0 = float(tot[3] / (double(B) * L));
    float klat = hp.kl_weight * kl0;                                // :235
    if (T > 1) klat = hp.kl_weight * ((kl0 + float(tot[4] / (double(T - 1) * B * L)) * float(T - 1)) / float(T));   // :229-232
    out[2] = float(tot[5] / n) * 0.5f * 0.5f;                       // :187-188
    out[3] = klat;
    out[4] = hp.entropy_cost * -float(tot[2] / n);                  // :191-194
    out[5] = mean; out[6] = stdv; out[7] = inv;
  } else {
    out[1] = -float(tot[0] / n);                                    // :184
    out[0] = out[1] + out[2] + out[4] + out[3];                     // :237
  }
}

__global__ void __launch_bounds__(32 * kPpoWarps) ppo_grad_kernel(const float* __restrict__ logits, const float* __restrict__ raw_action,
                                                       const float* __restrict__ eps, const float* __restrict__ lat_mean,
                                                       const float* __restrict__ lat_logvar, const float* __restrict__ baseline,
                                                       const float* __restrict__ vs, const float* __restrict__ logp,
                                                       const float* __restrict__ behaviour_logp, const float* __restrict__ stats, int T, int B,
                                                       int A, int L, PpoHyper hp, const float* __restrict__ adv_raw, float* __restrict__ advantages,
                                                       float* __restrict__ d_logits,
                                                       float* __restrict__ d_mean, float* __restrict__ d_logvar, float* __restrict__ d_baseline,
                                                       double* __restrict__ partial) {
  const size_t nrow = size_t(T) * B, stride = size_t(gridDim.x) * kPpoWarps * kPpoRowsPerWarp;
  const int lane = threadIdx.x & (kPpoLanes - 1);
  const float inv_n = 1.f / float(nrow), adv_mean = stats[5], adv_inv = stats[7];
  const float ce = -hp.entropy_cost * inv_n;                        // d entropy_loss / d (row entropy)
  const float w = hp.kl_weight / (float(T) * float(B) * float(L));  // kl_0 and kl_t both end up with this weight per element
  double acc[1] = {0.0};
  for (size_t row0 = (size_t(blockIdx.x) * kPpoWarps + (threadIdx.x >> 5)) * kPpoRowsPerWarp; row0 < nrow; row0 += stride) {
    const size_t row = row0 + ((threadIdx.x & 31) / kPpoLanes);
    if (row >= nrow) continue;                                      // no warp-wide operation below
    const float a = (adv_raw[row] - adv_mean) * adv_inv, rho = expf(logp[row] - behaviour_logp[row]);   // :177
    const float clipped = fminf(fmaxf(rho, 1.f - hp.clipping_epsilon), 1.f + hp.clipping_epsilon);
    const float s1 = rho * a, s2 = clipped * a;                     // :179-182
    // d policy_loss / d log-prob: the unclipped branch is the minimum (or both coincide) -> -A rho / N, else 0
    const float cp = (s1 <= s2) ? -a * rho * inv_n : 0.f;
    const float* lg = logits + row * 2 * A;
    float* dl = d_logits + row * 2 * A;
    for (int i = lane; i < A; i += kPpoLanes) {
      const float loc = lg[i], sr = lg[A + i], scale = softplus(sr) + 0.001f, raw = raw_action[row * A + i], e = eps[row * A + i];
      const float z = (raw - loc) / scale, rs = 1.f / scale;
      const float dj = -2.f * tanhf(fmaf(scale, e, loc));           // d log_det_tanh(x) / dx
      const float sig = 1.f / (1.f + expf(-sr));                    // d softplus
      dl[i] = cp * (z * rs) + ce * dj;
      dl[A + i] = (cp * ((z * z - 1.f) * rs) + ce * (rs + dj * e)) * sig;
    }
    const int t = int(row / B);
    const float *mu = lat_mean + row * L, *lv = lat_logvar + row * L, *mp = mu - size_t(B) * L, *mn = mu + size_t(B) * L;
    for (int j = lane; j < L; j += kPpoLanes) {
      const float m = mu[j], ev = expf(lv[j]);
      float gm, gv;
      if (t == 0) { gm = m; gv = -0.5f * (1.f - ev); }
      else { gm = -(kArAlpha * mp[j] - m) * kArInvPriorVar; gv = 0.5f * (ev * kArInvPriorVar - 1.f); }
      if (t + 1 < T) gm += kArAlpha * (kArAlpha * m - mn[j]) * kArInvPriorVar;   // this row's mean is z_{t-1} of the next step's prior
      d_mean[row * L + j] = w * gm;
      d_logvar[row * L + j] = w * gv;
    }
    if (lane == 0) {
      d_baseline[row] = -0.5f * (vs[row] - baseline[row]) * inv_n;  // d [mean(e^2) / 4] / d baseline
      advantages[row] = a;
      acc[0] += double(fminf(s1, s2));
    }
  }
  block_partials(acc, 1, partial);
}

// One-pass loss head (default).  The advantage statistics only need the GAE output, so they are reduced first (adv_stats_kernel over
// the 4-byte-per-row advantage array); the row kernel then reads every input ONCE and writes every gradient once: per row group
// (8 lanes) the action terms are kept in registers between the log-prob sum and the gradient that needs it (rho = exp(logp - behaviour)).
// DRAM traffic = the algorithmic 1.9 KB per row (the two-pass form read the inputs twice: 1.76 x).
__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ adv, size_t n, double* __restrict__ partial) {
  __shared__ double sh[2][8];
  double a = 0.0, a2 = 0.0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) { const double v = adv[i]; a += v; a2 += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = a2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    partial[size_t(blockIdx.x) * kPpoSums + threadIdx.x] = t;
  }
}

constexpr int kPpoMaxPerLane = 8;   // action elements per lane of a row group: A <= 64
// kFast: MUFU-based exp / log / tanh (ex2.approx / lg2.approx, ~2 ulp; absolute error of a row's log-prob ~1e-6): the row kernel is
// issue-bound on the accurate libm chains (three softplus, a log, a tanh per action element), not on memory.  The reference's
// jnp.exp / log on GPU are XLA's own polynomial approximations of comparable accuracy.  TMJX_PPO_ACCURATE=1 selects libm.
template <bool kFast> __device__ __forceinline__ float exp_t(float x) { return kFast ? __expf(x) : expf(x); }
template <bool kFast> __device__ __forceinline__ float log_t(float x) { return kFast ? __logf(x) : logf(x); }
template <bool kFast> __device__ __forceinline__ float softplus_t(float x) {
  return kFast ? fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))) : softplus(x);
}
template <bool kFast> __device__ __forceinline__ float tanh_t(float x) {
  if (!kFast) return tanhf(x);
  const float e = __expf(-2.f * fabsf(x));                  // in (0, 1]: no overflow
  return copysignf(__fdividef(1.f - e, 1.f + e), x);
}
template <bool kFast> __device__ __forceinline__ float log_det_tanh_t(float x) { return 2.f * (0.69314718055994531f - x - softplus_t<kFast>(-2.f * x)); }

template <bool kFast>
__global__ void __launch_bounds__(32 * kPpoWarps) ppo_fused_kernel(const float* __restrict__ logits, const float* __restrict__ raw_action,
                                                        const float* __restrict__ eps, const float* __restrict__ lat_mean,
                                                        const float* __restrict__ lat_logvar, const float* __restrict__ baseline,
                                                        const float* __restrict__ vs, const float* __restrict__ behaviour_logp,
                                                        const float* __restrict__ stats, int T, int B, int A, int L, PpoHyper hp,
                                                        const float* __restrict__ adv_raw, float* __restrict__ advantages,
                                                        float* __restrict__ d_logits, float* __restrict__ d_mean, float* __restrict__ d_logvar,
                                                        float* __restrict__ d_baseline, double* __restrict__ partial) {
  const size_t nrow = size_t(T) * B, stride = size_t(gridDim.x) * kPpoWarps * kPpoRowsPerWarp;
  const int lane = threadIdx.x & (kPpoLanes - 1);
  const float inv_n = 1.f / float(nrow), adv_mean = stats[5], adv_inv = stats[7];
  const float ce = -hp.entropy_cost * inv_n;
  const float w = hp.kl_weight / (float(T) * float(B) * float(L));
  const float kLogPriorVar = logf(kArPriorVar);
  double acc[kPpoSums] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // 0 surrogate, 2 entropy, 3 kl_0, 4 kl_t, 5 (vs - baseline)^2
  for (size_t row0 = (size_t(blockIdx.x) * kPpoWarps + (threadIdx.x >> 5)) * kPpoRowsPerWarp; row0 < nrow; row0 += stride) {
    const size_t row = row0 + ((threadIdx.x & 31) / kPpoLanes);
    const bool live = row < nrow;
    const int t = live ? int(row / B) : 0;
    float lp = 0.f, en = 0.f, k = 0.f;
    float r_z[kPpoMaxPerLane], r_rs[kPpoMaxPerLane], r_dj[kPpoMaxPerLane], r_sig[kPpoMaxPerLane], r_e[kPpoMaxPerLane];
    if (live) {
      const float* lg = logits + row * 2 * A;
#pragma unroll
      for (int q = 0; q < kPpoMaxPerLane; ++q) {
        const int i = lane + q * kPpoLanes;
        if (i < A) {
          const float loc = lg[i], sr = lg[A + i], scale = softplus_t<kFast>(sr) + 0.001f, raw = raw_action[row * A + i], e = eps[row * A + i];
          const float z = (raw - loc) / scale, ls = log_t<kFast>(scale), x = fmaf(scale, e, loc);
          lp += -0.5f * z * z - ls - 0.91893853320467274f - log_det_tanh_t<kFast>(raw);
          en += 0.5f + 0.91893853320467274f + ls + log_det_tanh_t<kFast>(x);
          r_z[q] = z; r_rs[q] = 1.f / scale; r_dj[q] = -2.f * tanh_t<kFast>(x); r_sig[q] = __fdividef(1.f, 1.f + exp_t<kFast>(-sr)); r_e[q] = e;
        }
      }
      const float *mu = lat_mean + row * L, *lv = lat_logvar + row * L, *mp = mu - size_t(B) * L, *mn = mu + size_t(B) * L;
      for (int j = lane; j < L; j += kPpoLanes) {
        const float m = mu[j], v = lv[j], ev = exp_t<kFast>(v);
        float gm, gv;
        if (t == 0) {
          k += 1.f + v - m * m - ev;
          gm = m; gv = -0.5f * (1.f - ev);
        } else {
          const float dm = kArAlpha * mp[j] - m;
          k += (ev + dm * dm) * kArInvPriorVar - 1.f + (kLogPriorVar - v);
          gm = -dm * kArInvPriorVar; gv = 0.5f * (ev * kArInvPriorVar - 1.f);
        }
        if (t + 1 < T) gm += kArAlpha * (kArAlpha * m - mn[j]) * kArInvPriorVar;
        d_mean[row * L + j] = w * gm;
        d_logvar[row * L + j] = w * gv;
      }
    }
    lp = group_sum(lp); en = group_sum(en); k = group_sum(k);
    if (!live) continue;
    const float a = (adv_raw[row] - adv_mean) * adv_inv, rho = exp_t<kFast>(lp - behaviour_logp[row]);
    const float clipped = fminf(fmaxf(rho, 1.f - hp.clipping_epsilon), 1.f + hp.clipping_epsilon);
    const float s1 = rho * a, s2 = clipped * a;
    const float cp = (s1 <= s2) ? -a * rho * inv_n : 0.f;
    float* dl = d_logits + row * 2 * A;
#pragma unroll
    for (int q = 0; q < kPpoMaxPerLane; ++q) {
      const int i = lane + q * kPpoLanes;
      if (i < A) {
        dl[i] = cp * (r_z[q] * r_rs[q]) + ce * r_dj[q];
        dl[A + i] = (cp * ((r_z[q] * r_z[q] - 1.f) * r_rs[q]) + ce * (r_rs[q] + r_dj[q] * r_e[q])) * r_sig[q];
      }
    }
    if (lane == 0) {
      const double e = double(vs[row]) - double(baseline[row]);
      d_baseline[row] = -0.5f * (vs[row] - baseline[row]) * inv_n;
      advantages[row] = a;
      acc[0] += double(fminf(s1, s2));
      acc[2] += en;
      if (t == 0) acc[3] += -0.5 * double(k); else acc[4] += 0.5 * double(k);
      acc[5] += e * e;
    }
  }
  block_partials(acc, kPpoSums, partial);
}

// ---- Optimiser step (reference ppo.py:517-520: optax.chain(clip_by_global_norm(10.0), adam(lr)), optax 0.2.5) -------------
// Flat fp32 buffers (parameters, gradients, first and second moments), 16-byte accesses over the aligned body.  Two launches:
// block partial sums of g^2 (double, fixed order), then the update, every block re-reducing the kAdamBlocks partials for the
// global norm (1184 doubles from L2, cheaper than a third launch).  28 B per parameter + 4 B for the norm pass; at the
// intention network's 2.6 M parameters all of it is L2-resident and the step is launch-latency bound.
constexpr int kAdamBlocks = 148 * 8, kAdamThreads = 256;
__global__ void __launch_bounds__(kAdamThreads) sumsq_partial_kernel(const float* __restrict__ g, size_t n, float grad_scale,
                                                                    double* __restrict__ partial) {
  __shared__ double sh[kAdamThreads / 32];
  double acc = 0.0;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nthr = size_t(gridDim.x) * blockDim.x;
  const size_t n4 = (reinterpret_cast<uintptr_t>(g) % 16 == 0) ? n / 4 : 0;
  for (size_t i = tid; i < n4; i += nthr) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float a = v.x * grad_scale, b = v.y * grad_scale, c = v.z * grad_scale, d = v.w * grad_scale;
    acc += (double(a) * double(a) + double(b) * double(b)) + (double(c) * double(c) + double(d) * double(d));
  }
  for (size_t i = 4 * n4 + tid; i < n; i += nthr) {
    const float v = g[i] * grad_scale;
    acc += double(v) * double(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kAdamThreads / 32; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(kAdamThreads) adam_apply_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu,
                                                                 float* __restrict__ nu, size_t n, float lr, float b1, float b2, float eps,
                                                                 float max_norm, float grad_scale, int count, const double* __restrict__ partial,
                                                                 int npartial, float* __restrict__ norm_out) {
  __shared__ double sh[kAdamThreads];
  __shared__ float s_norm;
  if (max_norm > 0.f || norm_out) {
    double t = 0.0;
    for (int b = threadIdx.x; b < npartial; b += blockDim.x) t += partial[b];
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int o = kAdamThreads / 2; o > 0; o >>= 1) {
      if (int(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) { s_norm = float(sqrt(sh[0])); if (norm_out && blockIdx.x == 0) norm_out[0] = s_norm; }
    __syncthreads();
  }
  const float g_norm = (max_norm > 0.f) ? s_norm : 0.f;
  const bool clip = max_norm > 0.f && !(g_norm < max_norm);          // optax: select(g_norm < max_norm, g, (g / g_norm) * max_norm)
  const float c1 = 1.f - powf(b1, float(count)), c2 = 1.f - powf(b2, float(count));   // bias corrections, count after increment
  auto elem = [&](float gi, float& pi, float& mi, float& vi) {
    gi *= grad_scale;
    if (clip) gi = (gi / g_norm) * max_norm;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * (gi * gi);
    pi = pi + (-lr) * ((mi / c1) / (sqrtf(vi / c2) + eps));
  };
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nthr = size_t(gridDim.x) * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mu) |
                         reinterpret_cast<uintptr_t>(nu)) % 16) == 0;
  const size_t n4 = aligned ? n / 4 : 0;
  for (size_t i = tid; i < n4; i += nthr) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(mu)[i], v4 = reinterpret_cast<float4*>(nu)[i];
    elem(g4.x, p4.x, m4.x, v4.x); elem(g4.y, p4.y, m4.y, v4.y); elem(g4.z, p4.z, m4.z, v4.z); elem(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(mu)[i] = m4; reinterpret_cast<float4*>(nu)[i] = v4;
  }
  for (size_t i = 4 * n4 + tid; i < n; i += nthr) elem(g[i], p[i], mu[i], nu[i]);
}

// Observation-normaliser update (reference masked_running_statistics.py:80-214, called at ppo.py:357-361).  HBM-bound: ONE
// pass over the [N, D] batch (the reference reads it twice), consecutive threads on consecutive columns (coalesced rows), each
// block owns a contiguous slab of rows and keeps, per column, sum(x - p) and sum((x - p)^2) about a pivot p = the slab's first
// row (a sample of the distribution, so the two-sum variance has no catastrophic cancellation even when the running mean is
// far from the data, as it is on the first update).  Slab moments (mean, M2) are merged in a fixed block order with the
// pairwise-update formula (deterministic, no atomics).  The reference's sum((x - m)(x - m')) with m' = m + u equals
// M2 + n (xbar - m)(xbar - m'), which the two small kernels after the all-reduce of sum(x - m) evaluate (stats_mean_kernel).
constexpr int kStatThreads = 256, kStatColsPerThread = 4;   // D <= 1024
constexpr int kStatBlocks = 148 * 4;                          // row slabs: four resident blocks per SM
__global__ void __launch_bounds__(kStatThreads) stats_partial_kernel(const float* __restrict__ x, int N, int D, float* __restrict__ partial) {
  const int rows_per_block = (N + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  if (r0 >= r1) return;
  float p[kStatColsPerThread], s1[kStatColsPerThread], s2[kStatColsPerThread];
#pragma unroll
  for (int j = 0; j < kStatColsPerThread; ++j) {
    const int c = threadIdx.x + kStatThreads * j;
    p[j] = c < D ? x[size_t(r0) * D + c] : 0.f; s1[j] = 0.f; s2[j] = 0.f;
  }
  for (int r = r0 + 1; r < r1; ++r) {
    const float* row = x + size_t(r) * D;
#pragma unroll
    for (int j = 0; j < kStatColsPerThread; ++j) {
      const int c = threadIdx.x + kStatThreads * j;
      if (c < D) { const float d = row[c] - p[j]; s1[j] += d; s2[j] = fmaf(d, d, s2[j]); }
    }
  }
  const float n = float(r1 - r0);
#pragma unroll
  for (int j = 0; j < kStatColsPerThread; ++j) {
    const int c = threadIdx.x + kStatThreads * j;
    if (c < D) {
      const float dm = s1[j] / n;
      partial[(size_t(blockIdx.x) * 2) * D + c] = p[j] + dm;                              // slab mean
      partial[(size_t(blockIdx.x) * 2 + 1) * D + c] = fmaxf(s2[j] - dm * s1[j], 0.f);     // slab M2 = sum (x - slab mean)^2
    }
  }
}
// The same slab moments with 16-byte loads (D % 4 == 0, which the tracking observations satisfy: 696 = 4 x 174).  Block =
// 32 float4 columns (one 512-byte row segment per warp-row) x 8 row lanes, four independent loads in flight per thread; all
// lanes share the slab's first row as the pivot, so their sums simply add (fixed lane order, deterministic).  Grid = column
// chunks x kStatSlabs row slabs (6 x 192 = 1152 blocks for D = 696: ~8 resident blocks per SM).
constexpr int kStatTx = 32, kStatTy = 8, kStatSlabs = 192;
__device__ __forceinline__ void stat_acc(const float4 v, const float4 p, float4& s1, float4& s2) {
  const float dx = v.x - p.x, dy = v.y - p.y, dz = v.z - p.z, dw = v.w - p.w;
  s1.x += dx; s1.y += dy; s1.z += dz; s1.w += dw;
  s2.x = fmaf(dx, dx, s2.x); s2.y = fmaf(dy, dy, s2.y); s2.z = fmaf(dz, dz, s2.z); s2.w = fmaf(dw, dw, s2.w);
}
__global__ void __launch_bounds__(kStatTx* kStatTy) stats_partial_vec_kernel(const float4* __restrict__ x, int N, int D4, float* __restrict__ partial) {
  __shared__ float4 sh1[kStatTy][kStatTx], sh2[kStatTy][kStatTx];
  const int c4 = blockIdx.x * kStatTx + threadIdx.x;
  const int rows_per_slab = (N + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(N, r0 + rows_per_slab);
  if (r0 >= r1) return;
  const bool live = c4 < D4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 p = zero, s1 = zero, s2 = zero;
  if (live) {
    const float4* col = x + c4;
    p = __ldg(col + size_t(r0) * D4);
    int r = r0 + 1 + threadIdx.y;
    for (; r + 3 * kStatTy < r1; r += 4 * kStatTy) {
      const float4 v0 = __ldcs(col + size_t(r) * D4), v1 = __ldcs(col + size_t(r + kStatTy) * D4),
                   v2 = __ldcs(col + size_t(r + 2 * kStatTy) * D4), v3 = __ldcs(col + size_t(r + 3 * kStatTy) * D4);
      stat_acc(v0, p, s1, s2); stat_acc(v1, p, s1, s2); stat_acc(v2, p, s1, s2); stat_acc(v3, p, s1, s2);
    }
    for (; r < r1; r += kStatTy) stat_acc(__ldcs(col + size_t(r) * D4), p, s1, s2);
  }
  sh1[threadIdx.y][threadIdx.x] = s1; sh2[threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    for (int l = 1; l < kStatTy; ++l) {
      const float4 a = sh1[l][threadIdx.x], b = sh2[l][threadIdx.x];
      s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
      s2.x += b.x; s2.y += b.y; s2.z += b.z; s2.w += b.w;
    }
    const float n = float(r1 - r0);
    const float4 dm = make_float4(s1.x / n, s1.y / n, s1.z / n, s1.w / n);
    float4* out = reinterpret_cast<float4*>(partial + (size_t(blockIdx.y) * 2) * (4 * size_t(D4))) + c4;
    out[0] = make_float4(p.x + dm.x, p.y + dm.y, p.z + dm.z, p.w + dm.w);
    out[D4] = make_float4(fmaxf(s2.x - dm.x * s1.x, 0.f), fmaxf(s2.y - dm.y * s1.y, 0.f), fmaxf(s2.z - dm.z * s1.z, 0.f),
                          fmaxf(s2.w - dm.w * s1.w, 0.f));
  }
}
// Merge of the slab moments, one warp per column: lane l merges slabs l, l + 32, ... in order (its loads are independent of
// the merge chain, so they are all in flight at once), then a five-step xor tree merges the lanes -- a fixed order for a given
// slab count.  out[0..D) = sum(x - mean) = N (xbar - mean), out[D..2D) = xbar, out[2D..3D) = M2
__device__ __forceinline__ void moments_merge(float& n, float& mu, float& m2, float nb, float mb, float qb) {
  if (nb == 0.f) return;
  if (n == 0.f) { n = nb; mu = mb; m2 = qb; return; }
  const float nt = n + nb, delta = mb - mu;
  mu += delta * (nb / nt);
  m2 += qb + delta * delta * (n * nb / nt);
  n = nt;
}
__global__ void __launch_bounds__(256) stats_combine_kernel(const float* __restrict__ partial, int N, int nblk, int D,
                                                            const float* __restrict__ mean, float* __restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= D) return;
  const int rows_per_block = (N + nblk - 1) / nblk;
  float n = 0.f, mu = 0.f, m2 = 0.f;
#pragma unroll 4
  for (int b = lane; b < nblk; b += 32) {
    const int r0 = b * rows_per_block, r1 = min(N, r0 + rows_per_block);
    if (r0 < r1) moments_merge(n, mu, m2, float(r1 - r0), partial[(size_t(b) * 2) * D + c], partial[(size_t(b) * 2 + 1) * D + c]);
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mu, o), qb = __shfl_xor_sync(0xffffffffu, m2, o);
    // both partners must end up with the same bits: the lower lane's moments are always the left operand
    if (lane & o) { float n2 = nb, mu2 = mb, q2 = qb; moments_merge(n2, mu2, q2, n, mu, m2); n = n2; mu = mu2; m2 = q2; }
    else moments_merge(n, mu, m2, nb, mb, qb);
  }
  if (lane == 0) {
    out[c] = n * (mu - mean[c]);
    out[D + c] = mu;
    out[2 * D + c] = m2;
  }
}
// After sum(x - mean) [D] and the row count were summed over the GPUs: new count (to a side cell), mean update in place and this
// GPU's share of the variance update, left where sum(x - mean) was.  With d = xbar_local - m, g = S1 / N_total (global batch
// mean - m) and u = S1 / count' the reference's sum_local (x - m)(x - m - u) summed over the GPUs equals the sum of
//   M2_local + n_local ((d - g)^2 + g^2 count_old / count')
// -- every term non-negative: no cancellation against the rounding of the stored mean, which costs the reference's own float32
// evaluation up to 1e-4 of the variance on the first update
__global__ void stats_mean_kernel(float* __restrict__ sums, const float* __restrict__ increment, int n_local, int D, const float* __restrict__ count,
                                  float* __restrict__ mean, float* __restrict__ count_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const float cnt0 = count[0], inc = increment[0], cnt = cnt0 + inc;
  if (c < D) {
    const float s1 = sums[c], m = mean[c], d = sums[D + c] - m, g = s1 / inc, e = d - g;
    mean[c] = m + s1 / cnt;
    sums[c] = sums[2 * D + c] + float(n_local) * (e * e + g * g * (cnt0 / cnt));
  }
  if (c == 0) count_out[0] = cnt;   // a second cell, so that every thread reads the old count
}
// After the variance shares were summed over the GPUs: summed_variance += var; std = clip(sqrt(max(sv, 0) / count)); count commit
__global__ void stats_apply_kernel(const float* __restrict__ var, int D, float std_min, float std_max, const float* __restrict__ count_new,
                                   float* __restrict__ count, float* __restrict__ sv, float* __restrict__ stdv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const float cnt = count_new[0];
  if (c < D) {
    const float v = sv[c] + var[c];
    sv[c] = v;
    stdv[c] = fminf(fmaxf(sqrtf(fmaxf(v, 0.f) / cnt), std_min), std_max);
  }
  if (c == 0) count[0] = cnt;
}

struct Layer {
  int k = 0, n = 0, kpad = 0, npad = 0, act = 0, ln = 0;
  float *wt = nullptr, *bias = nullptr, *ln_scale = nullptr, *ln_bias = nullptr;
  // where the layer's parameters sit in the flat parameter vector (floats from its start): kernel [k, n1] + bias [n1], for the
  // fused (mean | logvar) head a second kernel [k, n - n1] + bias; LayerNorm scale / bias.  Used by tmjx_*_set_params and the trainer.
  size_t off_w = 0, off_b = 0, off_w2 = 0, off_b2 = 0, off_lns = 0, off_lnb = 0;
  int n1 = 0;
  CUtensorMap mapW;                 // [npad rows, kpad] fp32, box 32 x BN, 128-byte swizzle
  CUtensorMap mapX;                 // the layer's input buffer inside tmjx_policy_act ([max_env rows, kpad], box 32 x 256)
  CUtensorMap mapX128;              // the same buffer with a 32 x 128 box: the fused chain kernel (tmjx_chain.cuh) owns 128 rows per CTA
  // fused chain kernel, layers whose input is a LayerNorm output: the normalisation is applied THROUGH the weights (tmjx_chain.cuh header):
  // wt_f[j, i] = wt[j, i] g_prev[i], cvec[j] = sum_i wt[j, i] g_prev[i], bias_f[j] = bias[j] + sum_i wt[j, i] beta_prev[i]
  float *wt_f = nullptr, *bias_f = nullptr, *cvec = nullptr;
  CUtensorMap mapWf;
  CUtensorMap mapWc, mapWfc;        // the weight maps with a 32 x (cw / cluster size) box: a CTA's multicast share of a slice
  const float* x_bound = nullptr;   // the buffer mapX was encoded for
  int ldx_bound = 0;
};

// LayerNorm of the previous layer folded into this layer's operands (fused chain kernel): one warp per output row j
__global__ void fold_ln_kernel(const float* __restrict__ wt, const float* __restrict__ bias, const float* __restrict__ g, const float* __restrict__ beta,
                               int k, int kpad, int npad, float* __restrict__ wt_f, float* __restrict__ bias_f, float* __restrict__ cvec) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= npad) return;
  float c = 0.f, b = 0.f;
  for (int i = lane; i < kpad; i += 32) {
    const float w = wt[size_t(j) * kpad + i], gi = i < k ? g[i] : 0.f, bi = i < k ? beta[i] : 0.f;
    wt_f[size_t(j) * kpad + i] = w * gi;
    c = fmaf(w, gi, c); b = fmaf(w, bi, b);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (lane == 0) { cvec[j] = c; bias_f[j] = bias[j] + b; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// rows x cols fp32 matrix with row pitch ld (floats); box = 32 floats (one 128-byte swizzle atom) x box_rows
static bool encode_map(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(ld) * 4};
  const cuuint32_t box[2] = {32u, cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MN-major operand map of a row-major [rows, cols] matrix (see make_desc_sw128_mn): 32 (floats) x 32 (rows) box, 128-byte swizzle in 32-byte atoms
static bool encode_map_mn(CUtensorMap* m, const float* base, int rows, int cols, int ld) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(ld) * 4};
  const cuuint32_t box[2] = {32u, 32u};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tmjx_policy

#include "tmjx_chain.cuh"

using namespace tmjx_policy;

struct TmjxPolicy {
  TmjxPolicyDesc d;
  int device = 0, max_env = 0, desc_swap = 0, use_v1 = 0;   // use_v1: 0 = TMA kernel, 1 = 128 x 128 block-synchronous, 2 = cp.async warp-specialised
  int skip_epi = 0;   // timing experiment knob (TMJX_POLICY_SKIP_EPI=1): results are invalid
  int cluster = 1;    // thread-block cluster size of the fused launch (TMJX_CHAIN_CLUSTER = 1 / 2 / 4): weight slices are multicast inside a cluster
  int fused = 1;      // one persistent launch per act / value apply (tmjx_chain.cuh); TMJX_POLICY_FUSED=0 keeps the per-layer launches
  std::vector<Layer> enc, dec;   // enc: hidden layers + the fused (mean | logvar) head; dec: hidden layers + logits
  float *norm_mean = nullptr, *norm_std = nullptr;
  float* buf[2] = {nullptr, nullptr};   // ping-pong activations [max_env, ld_buf]
  float *enc_in = nullptr, *dec_in = nullptr;
  int ld_buf = 0, ld_enc = 0, ld_dec = 0;
  std::vector<void*> owned;
};

// (re)compute the folded operands of every layer that follows a LayerNorm layer in its stack; stream-ordered after the weight upload / repack
static void fold_layers(std::vector<Layer>& layers, cudaStream_t st) {
  for (size_t l = 1; l < layers.size(); ++l) {
    Layer& L = layers[l];
    const Layer& Pv = layers[l - 1];
    if (!Pv.ln || !L.wt_f) continue;
    fold_ln_kernel<<<(L.npad + 7) / 8, 256, 0, st>>>(L.wt, L.bias, Pv.ln_scale, Pv.ln_bias, L.k, L.kpad, L.npad, L.wt_f, L.bias_f, L.cvec);
  }
}

static thread_local std::string g_perr;
static int pfail(int code, const std::string& msg) { g_perr = msg; return code; }
#define PCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return pfail(TMJX_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)
static int pad_to(int x, int m) { return (x + m - 1) / m * m; }

extern "C" {

const char* tmjx_policy_last_error(void) { return g_perr.c_str(); }

void tmjx_policy_destroy(TmjxPolicy* p);

size_t tmjx_policy_param_count(const TmjxPolicyDesc* d) {
  if (!d) return 0;
  size_t n = 2 * size_t(d->obs_size);
  int k = d->reference_obs_size;
  for (int i = 0; i < d->n_encoder_layers; ++i) { n += size_t(k) * d->encoder_layers[i] + 3 * size_t(d->encoder_layers[i]); k = d->encoder_layers[i]; }
  n += 2 * (size_t(k) * d->latent_size + d->latent_size);
  k = d->latent_size + d->obs_size - d->reference_obs_size;
  for (int i = 0; i < d->n_decoder_layers; ++i) { n += size_t(k) * d->decoder_layers[i] + 3 * size_t(d->decoder_layers[i]); k = d->decoder_layers[i]; }
  n += size_t(k) * 2 * d->action_size + 2 * d->action_size;
  return n;
}

int tmjx_policy_create(const TmjxPolicyDesc* d, const float* params, size_t n_params, int device, int max_env, TmjxPolicy** out) {
  if (!d || !params || !out || max_env <= 0) return pfail(TMJX_E_ARG, "null argument");
  if (d->n_encoder_layers < 1 || d->n_encoder_layers > TMJX_POLICY_MAX_LAYERS || d->n_decoder_layers < 1 || d->n_decoder_layers > TMJX_POLICY_MAX_LAYERS)
    return pfail(TMJX_E_ARG, "layer count out of range");
  if (n_params != tmjx_policy_param_count(d)) return pfail(TMJX_E_ARG, "parameter vector has the wrong length");
  PCU(cudaSetDevice(device));
  auto* p = new TmjxPolicy();
  std::unique_ptr<TmjxPolicy, void (*)(TmjxPolicy*)> guard(p, tmjx_policy_destroy);   // error returns below free what was built
  p->d = *d; p->device = device; p->max_env = max_env;
  if (const char* e = std::getenv("TMJX_POLICY_DESC_SWAP")) p->desc_swap = atoi(e);
  if (const char* e = std::getenv("TMJX_POLICY_V1")) p->use_v1 = atoi(e);
  if (const char* e = std::getenv("TMJX_POLICY_SKIP_EPI")) p->skip_epi = atoi(e) ? 4 : 0;   // A/B knob: the 128 x 128 block-synchronous kernel
  if (const char* e = std::getenv("TMJX_POLICY_FUSED")) p->fused = atoi(e);
  if (const char* e = std::getenv("TMJX_CHAIN_CLUSTER")) p->cluster = atoi(e);
  if (p->cluster != 1 && p->cluster != 2 && p->cluster != 4) return pfail(TMJX_E_ARG, "TMJX_CHAIN_CLUSTER must be 1, 2 or 4");
  if (p->use_v1) p->fused = 0;
  const float* cur = params;
  auto upload = [&](const std::vector<float>& h, float** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(h.size(), 1) * 4);
    if (e != cudaSuccess) return e;
    p->owned.push_back(*dst);
    return cudaMemcpy(*dst, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  };
  {
    std::vector<float> m(cur, cur + d->obs_size); cur += d->obs_size;
    std::vector<float> s(cur, cur + d->obs_size); cur += d->obs_size;
    PCU(upload(m, &p->norm_mean)); PCU(upload(s, &p->norm_std));
  }
  // W is flax's Dense kernel [in, out] row-major; the tensor core wants both operands K-major: Wt[out (padded), in (padded)]
  auto dense = [&](int k, int n, const float* W, const float* b, Layer& L) -> cudaError_t {
    L.k = k; L.n = n; L.kpad = pad_to(k, BK); L.npad = n > 256 ? pad_to(n, 256) : pad_to(n, BN);
    std::vector<float> wt(size_t(L.npad) * L.kpad, 0.f), bb(L.npad, 0.f);
    for (int i = 0; i < k; ++i) for (int j = 0; j < n; ++j) wt[size_t(j) * L.kpad + i] = W[size_t(i) * n + j];
    for (int j = 0; j < n; ++j) bb[j] = b[j];
    cudaError_t e = upload(wt, &L.wt);
    if (e != cudaSuccess) return e;
    return upload(bb, &L.bias);
  };
  auto hidden = [&](int k, int n, Layer& L) -> cudaError_t {
    const float* W = cur; cur += size_t(k) * n;
    const float* b = cur; cur += n;
    cudaError_t e = dense(k, n, W, b, L);
    if (e != cudaSuccess) return e;
    L.act = 1; L.ln = 1;
    L.off_w = size_t(W - params); L.off_b = size_t(b - params); L.n1 = n;
    L.off_lns = size_t(cur - params); L.off_lnb = L.off_lns + size_t(n);
    std::vector<float> g(cur, cur + n); cur += n;
    std::vector<float> be(cur, cur + n); cur += n;
    g.resize(L.npad, 0.f); be.resize(L.npad, 0.f);
    e = upload(g, &L.ln_scale);
    if (e != cudaSuccess) return e;
    return upload(be, &L.ln_bias);
  };
  int k = d->reference_obs_size, widest = 0;
  for (int i = 0; i < d->n_encoder_layers; ++i) {
    Layer L; PCU(hidden(k, d->encoder_layers[i], L)); p->enc.push_back(L); k = d->encoder_layers[i]; widest = std::max(widest, L.npad);
  }
  {  // fc2_mean | fc2_logvar fused into one GEMM: columns 0 .. latent-1 mean, latent .. 2 latent - 1 logvar
    const int lat = d->latent_size;
    const float* Wm = cur; cur += size_t(k) * lat; const float* bm = cur; cur += lat;
    const float* Wl = cur; cur += size_t(k) * lat; const float* bl = cur; cur += lat;
    std::vector<float> W(size_t(k) * 2 * lat), b(2 * lat);
    for (int i = 0; i < k; ++i) for (int j = 0; j < lat; ++j) { W[size_t(i) * 2 * lat + j] = Wm[size_t(i) * lat + j]; W[size_t(i) * 2 * lat + lat + j] = Wl[size_t(i) * lat + j]; }
    for (int j = 0; j < lat; ++j) { b[j] = bm[j]; b[lat + j] = bl[j]; }
    Layer L; PCU(dense(k, 2 * lat, W.data(), b.data(), L));
    L.off_w = size_t(Wm - params); L.off_b = size_t(bm - params); L.off_w2 = size_t(Wl - params); L.off_b2 = size_t(bl - params); L.n1 = lat;
    p->enc.push_back(L); widest = std::max(widest, L.npad);
  }
  k = d->latent_size + d->obs_size - d->reference_obs_size;
  for (int i = 0; i < d->n_decoder_layers; ++i) {
    Layer L; PCU(hidden(k, d->decoder_layers[i], L)); p->dec.push_back(L); k = d->decoder_layers[i]; widest = std::max(widest, L.npad);
  }
  {
    const int n = 2 * d->action_size;
    const float* W = cur; cur += size_t(k) * n; const float* b = cur; cur += n;
    Layer L; PCU(dense(k, n, W, b, L)); L.off_w = size_t(W - params); L.off_b = size_t(b - params); L.n1 = n;
    p->dec.push_back(L); widest = std::max(widest, L.npad);
  }
  p->ld_buf = widest;
  p->ld_enc = pad_to(d->reference_obs_size, BK);
  p->ld_dec = pad_to(d->latent_size + d->obs_size - d->reference_obs_size, BK);
  for (int i = 0; i < 2; ++i) { PCU(cudaMalloc(&p->buf[i], size_t(max_env) * p->ld_buf * 4)); p->owned.push_back(p->buf[i]); PCU(cudaMemset(p->buf[i], 0, size_t(max_env) * p->ld_buf * 4)); }
  PCU(cudaMalloc(&p->enc_in, size_t(max_env) * p->ld_enc * 4)); p->owned.push_back(p->enc_in);
  PCU(cudaMalloc(&p->dec_in, size_t(max_env) * p->ld_dec * 4)); p->owned.push_back(p->dec_in);
  PCU(cudaMemset(p->enc_in, 0, size_t(max_env) * p->ld_enc * 4));   // the K padding columns stay zero
  PCU(cudaMemset(p->dec_in, 0, size_t(max_env) * p->ld_dec * 4));
  PCU(cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<256>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<128>::kSmem));
  if (!p->use_v1) {
    // weight tensor maps once; activation maps for the fixed buffer chain of tmjx_policy_act
    const float* x = p->enc_in;
    int ldx = p->ld_enc, pp = 0;
    auto bind = [&](Layer& L) -> bool {
      if (!encode_map(&L.mapW, L.wt, L.npad, L.kpad, L.kpad, L.npad >= 512 ? 256 : 128)) return false;
      if (!encode_map(&L.mapX, x, max_env, L.kpad, ldx, 256)) return false;
      if (!encode_map(&L.mapX128, x, max_env, L.kpad, ldx, 128)) return false;
      L.x_bound = x; L.ldx_bound = ldx;
      x = p->buf[pp]; ldx = p->ld_buf; pp ^= 1;
      return true;
    };
    for (Layer& L : p->enc) if (!bind(L)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
    x = p->dec_in; ldx = p->ld_dec;
    for (Layer& L : p->dec) if (!bind(L)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
    for (std::vector<Layer>* stack : {&p->enc, &p->dec}) {
      for (size_t l = 1; l < stack->size(); ++l) {
        Layer& L = (*stack)[l];
        if (!(*stack)[l - 1].ln) continue;
        PCU(cudaMalloc(&L.wt_f, size_t(L.npad) * L.kpad * 4)); p->owned.push_back(L.wt_f);
        PCU(cudaMalloc(&L.bias_f, size_t(L.npad) * 4)); p->owned.push_back(L.bias_f);
        PCU(cudaMalloc(&L.cvec, size_t(L.npad) * 4)); p->owned.push_back(L.cvec);
        if (!encode_map(&L.mapWf, L.wt_f, L.npad, L.kpad, L.kpad, L.npad >= 512 ? 256 : 128)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
        if (!encode_map(&L.mapWfc, L.wt_f, L.npad, L.kpad, L.kpad, (L.npad >= 512 ? 256 : 128) / p->cluster)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
      }
      for (Layer& L : *stack)
        if (!encode_map(&L.mapWc, L.wt, L.npad, L.kpad, L.kpad, (L.npad >= 512 ? 256 : 128) / p->cluster)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
      fold_layers(*stack, nullptr);
    }
    PCU(cudaDeviceSynchronize());
  }
  PCU(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem));
  if (int(p->enc.size() + p->dec.size()) > kChainMaxLayers) p->fused = 0;
  PCU(cudaFuncSetAttribute(linear_tf32_v2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, V2<256>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_v2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, V2<128>::kSmem));
  *out = guard.release();
  return TMJX_OK;
}

void tmjx_policy_destroy(TmjxPolicy* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (void* q : p->owned) cudaFree(q);
  delete p;
}

static int run_linear(const TmjxPolicy* p, const Layer& L, const float* x, int ldx, float* y, int ldy, int M, cudaStream_t st) {
  if (!p->use_v1) {
    CUtensorMap mx = L.mapX;
    if (x != L.x_bound || ldx != L.ldx_bound) {   // a caller-owned input (tmjx_policy_linear): encode its map on the fly
      if (!encode_map(&mx, x, M, L.kpad, ldx, 256)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
    }
    if (L.npad >= 512) {
      dim3 grid((M + 255) / 256, L.npad / 256);
      linear_tf32_tma_kernel<256><<<grid, kTmaThreads, V3<256>::kSmem, st>>>(mx, L.mapW, L.bias, y, ldy, M, L.kpad, L.act | p->skip_epi);
    } else {
      dim3 grid((M + 255) / 256, L.npad / 128);
      linear_tf32_tma_kernel<128><<<grid, kTmaThreads, V3<128>::kSmem, st>>>(mx, L.mapW, L.bias, y, ldy, M, L.kpad, L.act | p->skip_epi);
    }
  } else if (p->use_v1 == 1) {
    dim3 grid((M + BM - 1) / BM, L.npad / BN);
    linear_tf32_kernel<<<grid, THREADS, kSmemBytes, st>>>(x, ldx, L.wt, L.kpad, L.bias, y, ldy, M, L.kpad, L.act, p->desc_swap);
  } else if (L.npad >= 512) {   // wide layers: 256 x 256 tiles (one wave of <= 148 CTAs at 16384 rows x 512 columns)
    dim3 grid((M + 255) / 256, L.npad / 256);
    linear_tf32_v2_kernel<256><<<grid, THREADS, V2<256>::kSmem, st>>>(x, ldx, L.wt, L.kpad, L.bias, y, ldy, M, L.kpad, L.act);
  } else {
    dim3 grid((M + 255) / 256, L.npad / 128);
    linear_tf32_v2_kernel<128><<<grid, THREADS, V2<128>::kSmem, st>>>(x, ldx, L.wt, L.kpad, L.bias, y, ldy, M, L.kpad, L.act);
  }
  if (L.ln) layernorm_kernel<<<(M + 7) / 8, 256, 0, st>>>(y, ldy, L.n, L.ln_scale, L.ln_bias, M);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

static long long* g_chain_trace = nullptr;   // development only: tmjx_policy_chain_trace
// the layer list of a policy / value network as the fused chain kernel wants it: buffers in the ping-pong order the tensor maps were bound in
static void chain_fill(const TmjxPolicy* p, ChainParams& cp) {
  static_assert(sizeof(ChainParams) < 32000, "kernel parameter space");
  int pp = 0, i = 0;
  auto add = [&](const Layer& L) {
    ChainLayer& c = cp.L[i++];
    c.mapX = L.mapX128;
    if (L.wt_f) { c.mapW = p->cluster > 1 ? L.mapWfc : L.mapWf; c.bias = L.bias_f; c.cvec = L.cvec; }
    else { c.mapW = p->cluster > 1 ? L.mapWc : L.mapW; c.bias = L.bias; c.cvec = nullptr; }
    c.out = p->buf[pp]; c.ldo = p->ld_buf; c.save_h = nullptr; c.ldh = 0;
    c.kpad = L.kpad; c.npad = L.npad; c.n = L.n; c.cw = L.npad >= 512 ? 256 : 128; c.act = L.act; c.ln = L.ln; c.kind = 0;
    pp ^= 1;
  };
  for (const Layer& L : p->enc) add(L);
  for (const Layer& L : p->dec) add(L);
  cp.n_layers = i;
  cp.csz = p->cluster;
  if (const char* e = std::getenv("TMJX_CHAIN_DBG")) cp.dbg = atoi(e);
  if (const char* e = std::getenv("TMJX_CHAIN_STAGGER_NS")) cp.stagger_ns = unsigned(atoi(e));
  cp.trace = g_chain_trace;
}

// one launch of the fused chain: ceil(M / 128) CTAs rounded up to whole clusters (the spare CTAs of the last cluster run on rows >= M:
// out-of-range TMA boxes read zeros, every store is guarded)
static int launch_chain(const ChainParams& cp, int n_env, cudaStream_t st) {
  const int csz = cp.csz, ctas = ((n_env + 127) / 128 + csz - 1) / csz * csz;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(ctas)); cfg.blockDim = dim3(kChainThreads); cfg.dynamicSmemBytes = kChainSmem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(csz); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = csz > 1 ? 1 : 0;
  PCU(cudaLaunchKernelEx(&cfg, mlp_chain_kernel, cp));
  return TMJX_OK;
}

/* plain GEMM entry point (tests, microbenchmarks): y[M, ldy] = x[M, k] * W + b through the same tcgen05 kernel */
int tmjx_policy_linear(const TmjxPolicy* p, int which /* 0.. : encoder layers then decoder layers */, const float* x, int ldx, float* y, int ldy,
                       int n_env, void* stream) {
  if (!p || !x || !y) return pfail(TMJX_E_ARG, "null argument");
  const int ne = int(p->enc.size());
  if (which < 0 || which >= ne + int(p->dec.size())) return pfail(TMJX_E_ARG, "layer index out of range");
  const Layer& L = which < ne ? p->enc[which] : p->dec[which - ne];
  if (ldx < L.kpad || ldy < L.npad) return pfail(TMJX_E_ARG, "row pitch smaller than the padded layer width");
  PCU(cudaSetDevice(p->device));
  return run_linear(p, L, x, ldx, y, ldy, n_env, static_cast<cudaStream_t>(stream));
}

int tmjx_policy_act(const TmjxPolicy* p, const float* obs, const float* eps_latent, const float* eps_action, int deterministic, float* action,
                    float* raw_action, float* log_prob, float* logits, float* latent_mean, float* latent_logvar, int n_env, void* stream) {
  if (!p || !obs || !action) return pfail(TMJX_E_ARG, "null argument");
  if (!deterministic && (!eps_latent || !eps_action)) return pfail(TMJX_E_ARG, "stochastic acting needs eps_latent and eps_action");
  if (n_env <= 0 || n_env > p->max_env) return pfail(TMJX_E_ARG, "n_env exceeds the policy's max_env");
  PCU(cudaSetDevice(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TmjxPolicyDesc& d = p->d;
  if (p->fused) {   // ONE launch: normalise -> encoder -> latent -> decoder -> action rows (tmjx_chain.cuh)
    ChainParams cp;
    std::memset(&cp, 0, sizeof(cp));
    chain_fill(p, cp);
    cp.M = n_env;
    cp.obs = obs; cp.nobs = d.obs_size; cp.nref = d.reference_obs_size; cp.latent = d.latent_size;
    cp.mean = p->norm_mean; cp.stdv = p->norm_std; cp.enc_in = p->enc_in; cp.ld_enc = p->ld_enc; cp.dec_in = p->dec_in; cp.ld_dec = p->ld_dec;
    cp.eps_latent = eps_latent; cp.deterministic = deterministic; cp.out_mean = latent_mean; cp.out_logvar = latent_logvar;
    cp.na = d.action_size; cp.eps_action = eps_action; cp.action = action; cp.raw_action = raw_action; cp.log_prob = log_prob; cp.logits = logits;
    cp.L[p->enc.size() - 1].kind = 1;
    cp.L[cp.n_layers - 1].kind = 2;
    return launch_chain(cp, n_env, st);
  }
  obs_prep_kernel<<<n_env, 256, 0, st>>>(obs, d.obs_size, d.reference_obs_size, d.latent_size, p->norm_mean, p->norm_std, p->enc_in, p->ld_enc,
                                         p->dec_in, p->ld_dec, n_env);
  const float* x = p->enc_in;
  int ldx = p->ld_enc, pp = 0;
  for (const Layer& L : p->enc) {
    int rc = run_linear(p, L, x, ldx, p->buf[pp], p->ld_buf, n_env, st);
    if (rc) return rc;
    x = p->buf[pp]; ldx = p->ld_buf; pp ^= 1;
  }
  latent_kernel<<<(n_env * d.latent_size + 255) / 256, 256, 0, st>>>(x, ldx, d.latent_size, eps_latent, deterministic, p->dec_in, p->ld_dec,
                                                                      latent_mean, latent_logvar, n_env);
  x = p->dec_in; ldx = p->ld_dec;
  for (const Layer& L : p->dec) {
    int rc = run_linear(p, L, x, ldx, p->buf[pp], p->ld_buf, n_env, st);
    if (rc) return rc;
    x = p->buf[pp]; ldx = p->ld_buf; pp ^= 1;
  }
  action_head_kernel<<<(n_env + 7) / 8, 256, 0, st>>>(x, ldx, d.action_size, eps_action, deterministic, action, raw_action, log_prob, logits, n_env);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

/* Value network, see include/tmjx.h: the TmjxPolicy object with every layer in `enc` (Dense + SiLU, last layer Dense to 1). */
size_t tmjx_value_param_count(const TmjxValueDesc* d) {
  if (!d || d->n_hidden_layers < 0 || d->n_hidden_layers > TMJX_POLICY_MAX_LAYERS) return 0;
  size_t n = 2 * size_t(d->obs_size);
  int k = d->obs_size;
  for (int i = 0; i < d->n_hidden_layers; ++i) { n += size_t(k) * d->hidden_layers[i] + d->hidden_layers[i]; k = d->hidden_layers[i]; }
  return n + size_t(k) + 1;
}
int tmjx_value_create(const TmjxValueDesc* d, const float* params, size_t n_params, int device, int max_env, TmjxPolicy** out) {
  if (!d || !params || !out || max_env <= 0) return pfail(TMJX_E_ARG, "null argument");
  if (d->obs_size <= 0 || d->n_hidden_layers < 0 || d->n_hidden_layers > TMJX_POLICY_MAX_LAYERS) return pfail(TMJX_E_ARG, "bad value-network shape");
  if (n_params != tmjx_value_param_count(d)) return pfail(TMJX_E_ARG, "parameter vector has the wrong length");
  PCU(cudaSetDevice(device));
  auto* p = new TmjxPolicy();
  std::unique_ptr<TmjxPolicy, void (*)(TmjxPolicy*)> guard(p, tmjx_policy_destroy);
  p->d.obs_size = d->obs_size; p->device = device; p->max_env = max_env;
  const float* cur = params;
  auto upload = [&](const std::vector<float>& h, float** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(h.size(), 1) * 4);
    if (e != cudaSuccess) return e;
    p->owned.push_back(*dst);
    return cudaMemcpy(*dst, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  };
  {
    std::vector<float> m(cur, cur + d->obs_size); cur += d->obs_size;
    std::vector<float> sd(cur, cur + d->obs_size); cur += d->obs_size;
    PCU(upload(m, &p->norm_mean)); PCU(upload(sd, &p->norm_std));
  }
  int k = d->obs_size, widest = 0;
  for (int i = 0; i <= d->n_hidden_layers; ++i) {
    const int n = i < d->n_hidden_layers ? d->hidden_layers[i] : 1;
    if (n <= 0) return pfail(TMJX_E_ARG, "bad hidden layer size");
    const float* W = cur; cur += size_t(k) * n;
    const float* b = cur; cur += n;
    Layer L;
    L.k = k; L.n = n; L.kpad = pad_to(k, BK); L.npad = n > 256 ? pad_to(n, 256) : pad_to(n, BN);
    L.act = i < d->n_hidden_layers ? 1 : 0;                  // brax MLP: activation on every layer but the last, no LayerNorm
    L.off_w = size_t(W - params); L.off_b = size_t(b - params); L.n1 = n;
    std::vector<float> wt(size_t(L.npad) * L.kpad, 0.f), bb(L.npad, 0.f);
    for (int r = 0; r < k; ++r) for (int j = 0; j < n; ++j) wt[size_t(j) * L.kpad + r] = W[size_t(r) * n + j];
    for (int j = 0; j < n; ++j) bb[j] = b[j];
    PCU(upload(wt, &L.wt)); PCU(upload(bb, &L.bias));
    p->enc.push_back(L); k = n; widest = std::max(widest, L.npad);
  }
  p->ld_buf = widest;
  p->ld_enc = pad_to(d->obs_size, BK);
  for (int i = 0; i < 2; ++i) { PCU(cudaMalloc(&p->buf[i], size_t(max_env) * p->ld_buf * 4)); p->owned.push_back(p->buf[i]); PCU(cudaMemset(p->buf[i], 0, size_t(max_env) * p->ld_buf * 4)); }
  PCU(cudaMalloc(&p->enc_in, size_t(max_env) * p->ld_enc * 4)); p->owned.push_back(p->enc_in);
  PCU(cudaMemset(p->enc_in, 0, size_t(max_env) * p->ld_enc * 4));   // the K padding columns stay zero
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<256>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<128>::kSmem));
  const float* x = p->enc_in;
  int ldx = p->ld_enc, pp = 0;
  for (Layer& L : p->enc) {
    if (!encode_map(&L.mapW, L.wt, L.npad, L.kpad, L.kpad, L.npad >= 512 ? 256 : 128)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
    if (!encode_map(&L.mapX, x, max_env, L.kpad, ldx, 256)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
    if (!encode_map(&L.mapX128, x, max_env, L.kpad, ldx, 128)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
    L.x_bound = x; L.ldx_bound = ldx;
    x = p->buf[pp]; ldx = p->ld_buf; pp ^= 1;
  }
  PCU(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem));
  if (const char* e = std::getenv("TMJX_POLICY_FUSED")) p->fused = atoi(e);
  if (const char* e = std::getenv("TMJX_CHAIN_CLUSTER")) p->cluster = atoi(e);
  if (p->cluster != 1 && p->cluster != 2 && p->cluster != 4) return pfail(TMJX_E_ARG, "TMJX_CHAIN_CLUSTER must be 1, 2 or 4");
  for (Layer& L : p->enc)
    if (!encode_map(&L.mapWc, L.wt, L.npad, L.kpad, L.kpad, (L.npad >= 512 ? 256 : 128) / p->cluster)) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
  if (int(p->enc.size()) > kChainMaxLayers) p->fused = 0;
  *out = guard.release();
  return TMJX_OK;
}
int tmjx_value_apply(const TmjxPolicy* v, const float* obs, float* value, int n_env, void* stream) {
  if (!v || !obs || !value) return pfail(TMJX_E_ARG, "null argument");
  if (!v->dec.empty() || v->enc.empty() || v->enc.back().n != 1) return pfail(TMJX_E_ARG, "not a value network");
  if (n_env <= 0 || n_env > v->max_env) return pfail(TMJX_E_ARG, "n_env exceeds the network's max_env");
  PCU(cudaSetDevice(v->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (v->fused) {
    ChainParams cp;
    std::memset(&cp, 0, sizeof(cp));
    chain_fill(v, cp);
    cp.M = n_env;
    cp.obs = obs; cp.nobs = v->d.obs_size; cp.nref = v->d.obs_size; cp.mean = v->norm_mean; cp.stdv = v->norm_std; cp.enc_in = v->enc_in; cp.ld_enc = v->ld_enc;
    cp.value = value;
    cp.L[cp.n_layers - 1].kind = 3;
    return launch_chain(cp, n_env, st);
  }
  value_prep_kernel<<<n_env, 256, 0, st>>>(obs, v->d.obs_size, v->norm_mean, v->norm_std, v->enc_in, v->ld_enc, n_env);
  const float* x = v->enc_in;
  int ldx = v->ld_enc, pp = 0;
  for (const Layer& L : v->enc) {
    int rc = run_linear(v, L, x, ldx, v->buf[pp], v->ld_buf, n_env, st);
    if (rc) return rc;
    x = v->buf[pp]; ldx = v->ld_buf; pp ^= 1;
  }
  value_out_kernel<<<(n_env + 255) / 256, 256, 0, st>>>(x, ldx, value, n_env);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

int tmjx_gae(const float* truncation, const float* termination, const float* rewards, const float* values, const float* bootstrap_value,
             float lambda, float discount, float* vs, float* advantages, int T, int B, void* stream) {
  if (!truncation || !termination || !rewards || !values || !bootstrap_value || !vs || !advantages) return pfail(TMJX_E_ARG, "null argument");
  if (T <= 0 || B <= 0) return pfail(TMJX_E_ARG, "T and B must be positive");
  gae_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(truncation, termination, rewards, values, bootstrap_value, lambda,
                                                                           discount, vs, advantages, T, B);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

/* PPO loss head, see include/tmjx.h.  scratch: block partial sums (double), then 4 T B floats (scaled rewards, termination, row
 * log-prob, advantages before normalisation). */
size_t tmjx_ppo_loss_scratch_floats(int T, int B) { return 2 * size_t(kPpoBlocks) * kPpoSums + 4 * size_t(T) * size_t(B); }
int tmjx_ppo_loss_head(const float* logits, const float* latent_mean, const float* latent_logvar, const float* baseline,
                       const float* bootstrap_value, const float* reward, const float* discount, const float* truncation,
                       const float* raw_action, const float* behaviour_log_prob, const float* eps_entropy, int T, int B, int A, int L,
                       const TmjxPpoHyper* hyper, float* losses, float* vs, float* advantages, float* d_logits, float* d_latent_mean,
                       float* d_latent_logvar, float* d_baseline, float* scratch, void* stream) {
  if (!logits || !latent_mean || !latent_logvar || !baseline || !bootstrap_value || !reward || !discount || !truncation || !raw_action ||
      !behaviour_log_prob || !eps_entropy || !hyper || !losses || !vs || !advantages || !d_logits || !d_latent_mean || !d_latent_logvar ||
      !d_baseline || !scratch)
    return pfail(TMJX_E_ARG, "null argument");
  if (T <= 0 || B <= 0 || A <= 0 || L <= 0) return pfail(TMJX_E_ARG, "T, B, action and latent sizes must be positive");
  if (reinterpret_cast<uintptr_t>(scratch) % 8 != 0) return pfail(TMJX_E_ARG, "scratch must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = size_t(T) * B;
  double* partial = reinterpret_cast<double*>(scratch);
  float *rewards = scratch + 2 * size_t(kPpoBlocks) * kPpoSums, *termination = rewards + n, *logp = rewards + 2 * n, *adv_raw = rewards + 3 * n;
  PpoHyper hp{hyper->entropy_cost, hyper->kl_weight, hyper->discounting, hyper->reward_scaling, hyper->gae_lambda, hyper->clipping_epsilon,
              hyper->normalize_advantage};
  const size_t rows_per_block = size_t(kPpoWarps) * kPpoRowsPerWarp;
  const int nblk = int(std::min<size_t>(kPpoBlocks, (n + rows_per_block - 1) / rows_per_block));
  ppo_prep_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(reward, discount, truncation, hp.reward_scaling, rewards, termination, n);
  gae_kernel<<<(B + 127) / 128, 128, 0, st>>>(truncation, termination, rewards, baseline, bootstrap_value, hp.gae_lambda, hp.discounting, vs,
                                              adv_raw, T, B);
  static const bool two_pass = [] { const char* e = std::getenv("TMJX_PPO_TWO_PASS"); return e && atoi(e); }();   // A/B knob: the round-1 form
  if (!two_pass && A <= kPpoLanes * kPpoMaxPerLane) {
    const int nb0 = int(std::min<size_t>(kPpoBlocks, (n + 255) / 256));
    adv_stats_kernel<<<nb0, 256, 0, st>>>(adv_raw, n, partial);
    ppo_reduce_kernel<<<1, 256, 0, st>>>(partial, nb0, 0, T, B, L, hp, losses);
    static const bool accurate = [] { const char* e = std::getenv("TMJX_PPO_ACCURATE"); return e && atoi(e); }();
    if (accurate)
      ppo_fused_kernel<false><<<nblk, 32 * kPpoWarps, 0, st>>>(logits, raw_action, eps_entropy, latent_mean, latent_logvar, baseline, vs, behaviour_log_prob,
                                                                losses, T, B, A, L, hp, adv_raw, advantages, d_logits, d_latent_mean, d_latent_logvar,
                                                                d_baseline, partial);
    else
      ppo_fused_kernel<true><<<nblk, 32 * kPpoWarps, 0, st>>>(logits, raw_action, eps_entropy, latent_mean, latent_logvar, baseline, vs, behaviour_log_prob,
                                                               losses, T, B, A, L, hp, adv_raw, advantages, d_logits, d_latent_mean, d_latent_logvar,
                                                               d_baseline, partial);
    ppo_reduce_kernel<<<1, 256, 0, st>>>(partial, nblk, 3, T, B, L, hp, losses);
    PCU(cudaGetLastError());
    return TMJX_OK;
  }
  ppo_rows_kernel<<<nblk, 32 * kPpoWarps, 0, st>>>(logits, raw_action, eps_entropy, latent_mean, latent_logvar, adv_raw, vs, baseline, T, B,
                                                    A, L, logp, partial);
  ppo_reduce_kernel<<<1, 256, 0, st>>>(partial, nblk, 1, T, B, L, hp, losses);
  ppo_grad_kernel<<<nblk, 32 * kPpoWarps, 0, st>>>(logits, raw_action, eps_entropy, latent_mean, latent_logvar, baseline, vs, logp,
                                                    behaviour_log_prob, losses, T, B, A, L, hp, adv_raw, advantages, d_logits, d_latent_mean,
                                                    d_latent_logvar, d_baseline, partial);
  ppo_reduce_kernel<<<1, 256, 0, st>>>(partial, nblk, 2, T, B, L, hp, losses);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

/* Optimiser step, see include/tmjx.h. */
size_t tmjx_adam_scratch_floats(void) { return 2 * size_t(kAdamBlocks); }
int tmjx_adam_step(float* params, const float* grads, float* mu, float* nu, size_t n, float learning_rate, float b1, float b2, float eps,
                   float max_grad_norm, float grad_scale, int count, float* grad_norm_out, float* scratch, void* stream) {
  if (!params || !grads || !mu || !nu || !scratch) return pfail(TMJX_E_ARG, "null argument");
  if (n == 0 || count <= 0) return pfail(TMJX_E_ARG, "n and count (the 1-based step number) must be positive");
  if (reinterpret_cast<uintptr_t>(scratch) % 8 != 0) return pfail(TMJX_E_ARG, "scratch must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = int(std::min<size_t>(kAdamBlocks, (n + kAdamThreads - 1) / kAdamThreads));
  double* partial = reinterpret_cast<double*>(scratch);
  if (max_grad_norm > 0.f || grad_norm_out) sumsq_partial_kernel<<<nblk, kAdamThreads, 0, st>>>(grads, n, grad_scale, partial);
  adam_apply_kernel<<<nblk, kAdamThreads, 0, st>>>(params, grads, mu, nu, n, learning_rate, b1, b2, eps, max_grad_norm, grad_scale, count,
                                                   partial, nblk, grad_norm_out);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

/* Observation normaliser, see include/tmjx.h.  scratch: at least tmjx_running_stats_scratch_floats(D) floats. */
size_t tmjx_running_stats_scratch_floats(int D) { return size_t(kStatBlocks) * 2 * size_t(D) + 1; }
int tmjx_running_stats_sums(const float* batch, int N, int D, const float* mean, float* sums, float* scratch, void* stream) {
  if (!batch || !mean || !sums || !scratch) return pfail(TMJX_E_ARG, "null argument");
  if (N <= 0 || D <= 0 || D > kStatThreads * kStatColsPerThread) return pfail(TMJX_E_ARG, "bad batch shape (D <= 1024)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int nblk = std::min(kStatBlocks, N);
  if (D % 4 == 0 && reinterpret_cast<uintptr_t>(batch) % 16 == 0 && reinterpret_cast<uintptr_t>(scratch) % 16 == 0) {
    nblk = std::min(kStatSlabs, N);
    stats_partial_vec_kernel<<<dim3((D / 4 + kStatTx - 1) / kStatTx, nblk), dim3(kStatTx, kStatTy), 0, st>>>(
        reinterpret_cast<const float4*>(batch), N, D / 4, scratch);
  } else {
    stats_partial_kernel<<<nblk, kStatThreads, 0, st>>>(batch, N, D, scratch);
  }
  stats_combine_kernel<<<(D * 32 + 255) / 256, 256, 0, st>>>(scratch, N, nblk, D, mean, sums);
  PCU(cudaGetLastError());
  return TMJX_OK;
}
int tmjx_running_stats_mean(float* sums, const float* increment, int n_local, int D, const float* count, float* mean, float* scratch,
                            void* stream) {
  if (!sums || !increment || !count || !mean || !scratch) return pfail(TMJX_E_ARG, "null argument");
  if (n_local < 0 || D <= 0) return pfail(TMJX_E_ARG, "bad shape");
  stats_mean_kernel<<<(D + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(sums, increment, n_local, D, count, mean, scratch);
  PCU(cudaGetLastError());
  return TMJX_OK;
}
int tmjx_running_stats_apply(const float* var, int D, float std_min, float std_max, float* count, float* summed_variance, float* std,
                             const float* scratch, void* stream) {
  if (!var || !count || !summed_variance || !std || !scratch) return pfail(TMJX_E_ARG, "null argument");
  if (D <= 0) return pfail(TMJX_E_ARG, "bad shape");
  stats_apply_kernel<<<(D + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(var, D, std_min, std_max, scratch, count, summed_variance, std);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

/* development: per-layer clock64() marks of CTA 0 of the fused chain launch (tools/gpu_chain_trace.py).  enable != 0 allocates / arms the
 * buffer, out != null copies the 8 marks x kChainMaxLayers back.  Not part of the product surface (not declared in include/tmjx.h). */
int tmjx_policy_chain_trace(int enable, long long* out) {
  if (enable && !g_chain_trace) { PCU(cudaMalloc(&g_chain_trace, 8 * kChainMaxLayers * sizeof(long long))); }
  if (enable) PCU(cudaMemset(g_chain_trace, 0, 8 * kChainMaxLayers * sizeof(long long)));
  if (out && g_chain_trace) PCU(cudaMemcpy(out, g_chain_trace, 8 * kChainMaxLayers * sizeof(long long), cudaMemcpyDeviceToHost));
  if (!enable && !out && g_chain_trace) { cudaFree(g_chain_trace); g_chain_trace = nullptr; }
  return TMJX_OK;
}

int tmjx_policy_launches_per_act(const TmjxPolicy* p) {
  if (!p) return 0;
  if (p->fused) return 1;
  int n = 3;   // obs_prep, latent, action_head
  for (const Layer& L : p->enc) n += 1 + L.ln;
  for (const Layer& L : p->dec) n += 1 + L.ln;
  return n;
}

}  // extern "C"

#include "tmjx_train.cuh"
