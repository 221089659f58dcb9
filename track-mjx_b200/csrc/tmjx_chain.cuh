/*
 * tmjx_chain.cuh — one persistent launch for a whole stack of Dense (+ SiLU + LayerNorm) layers: the acting policy of
 * tmjx_policy_act (observation normalisation -> encoder -> reparameterised latent -> decoder -> tanh-normal head) and the value
 * network of tmjx_value_apply.  Included by tmjx_policy.cu (same namespace, same helpers).
 *
 * Replaces, for the acting / evaluation path, the per-layer launches of tmjx_policy.cu (12 GEMMs + 10 LayerNorms + 3 row kernels per
 * act): reference track_mjx/agent/mlp_ppo/intention_network.py:14-142, ppo_networks.py:34-100 ("persistent
 * fused-MLP policy kernel ... LayerNorm + SiLU in the epilogue ... 25 launches should become <= 3").
 *
 * Mapping.  An MLP is row-local: output row r of every layer depends on input row r only.  One CTA therefore owns 128 environments
 * (M = 128 = the TMEM lane count) and walks the WHOLE network for them:
 *   - layer input  : the CTA's 128 x K slab of the previous layer's output, re-read from global memory by TMA.  It was written a few
 *                    microseconds earlier by the same SM, 256 KB per layer, so it is an L2 hit -- fp32 activations of a 128-row slab
 *                    (256 - 512 KB) fit neither shared memory nor what TMEM has left beside the accumulators; L2 is the staging level;
 *   - weights      : streamed by TMA in 32-float K slices (128-byte swizzle), all CTAs read the same 10 MB, L2-resident;
 *   - accumulators : TMEM, two slots of 256 columns.  A layer wider than 256 is computed in 256-column chunks that alternate
 *                    between the slots, so chunk c + 1's MMAs run while chunk c is in the epilogue;
 *   - epilogue     : 16 warps (TMEM lane quarter = warp % 4, four column groups), thread = one row x 64 columns of a chunk, ONE pass:
 *                    tcgen05.ld, affine, SiLU, row sum / sum of squares, then the 32 x 32 register tile is transposed through a per-warp
 *                    shared-memory tile so that every store instruction writes four complete 128-byte row segments.
 *   - LayerNorm    : never materialised.  A hidden layer stores s = SiLU(h) un-normalised and keeps the row's (mean, rstd) in the
 *                    registers of the thread that owns the row; its consumer applies the normalisation through its operands:
 *                        LN(s) W + b = rstd (s (g . W)) - rstd mean (W^T g) + (W^T beta + b)
 *                    i.e. the GEMM runs on W' = diag(g) W (folded once per parameter update, `fold_ln_kernel`) and the epilogue's affine is
 *                    x = acc rstd - (rstd mean) cvec[col] + b'[col].  Same mathematics as flax's LayerNorm (fast variance, eps 1e-6) followed
 *                    by Dense, one rounding sequence apart (the TF32 operand rounding meets s g instead of LN(s)); what it buys: no second
 *                    epilogue pass over the accumulators (the two-pass form spent as long in the LayerNorm tail as in the MMAs).
 *   - heads        : the (mean | logvar) head writes z = mean + exp(logvar / 2) eps straight into the decoder input, the logits
 *                    head runs the tanh-normal sampling / log-prob rows -- both by the epilogue warps after a CTA-local barrier.
 * Roles: warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, warps 4..19 = epilogue.  Barriers: full / empty per operand stage,
 * tmem-full / tmem-empty per accumulator slot, and `ready` = "the previous layer's output is complete in global memory" (the epilogue
 * warps arrive after a generic->async proxy fence; the producer waits on it before the first activation slice of the next layer,
 * weight slices are prefetched ahead of it).
 */
#pragma once

namespace tmjx_policy {

constexpr int kChainMaxLayers = 14;
constexpr int kChainStages = 3;
constexpr int kChainABytes = 128 * BK * 4;                 // activation slice: 128 rows x 32 floats
constexpr int kChainWBytes = 256 * BK * 4;                 // weight slice: up to 256 output columns x 32 floats
constexpr int kChainStage = kChainABytes + kChainWBytes;   // 48 KB
constexpr int kChainTileBytes = 32 * 32 * 4;               // 32 x 32 transpose tile per epilogue warp, 16-byte chunks XOR-swizzled by the row
constexpr int kChainMaxN = 1024;                           // widest layer: its bias / cvec rows are staged in shared memory
constexpr int kChainSmem = kChainStages * kChainStage + 1024 + 16 * kChainTileBytes + 2 * kChainMaxN * 4;
constexpr int kChainThreads = 640;                         // warps 0 / 1: producer / MMA issuer, 2 - 3 idle, 4 - 19 epilogue
constexpr int kChainEpiThreads = 512;

struct alignas(64) ChainLayer {
  CUtensorMap mapX;   // the layer's input [rows, kpad] fp32, box 32 x 128
  CUtensorMap mapW;   // Wt [npad, kpad] fp32, box 32 x cw / csz (csz = cluster size: every CTA fetches its share of a weight slice and multicasts it)
  const float* bias;      // bias, or the folded b' = b + W^T beta_prev when the input is a LayerNorm output (see the header)
  const float* cvec;      // W^T g_prev per output column when the input is a LayerNorm output, else null
  float* out;         // [rows, ldo]: the layer's output = the next layer's input
  float* save_h;      // optional [rows, ldh]: the pre-activation x W + b (training forward), else null
  int ldo, ldh, kpad, npad, n, cw, act, ln;   // ln: this layer's output is layer-normalised (by its consumer, through cvec / the folded weights)
  int kind;           // 0 plain, 1 (mean | logvar) head -> latent, 2 logits head -> tanh-normal action rows, 3 value head (column 0)
};

struct ChainParams {
  int n_layers, M;
  int csz;               // thread-block cluster size (1, 2 or 4): the CTAs of a cluster share every weight slice through TMA multicast
  unsigned stagger_ns;   // experiment (TMJX_CHAIN_STAGGER_NS): odd CTAs start this much later, so that the chip's load and store bursts interleave
  long long* trace;   // development: CTA 0 records clock64() marks per layer (tmjx_policy_chain_trace), else null
  int dbg;   // timing experiments only (TMJX_CHAIN_DBG): 1 = no MMAs, 2 = no epilogue work, 4 = no operand loads; results are invalid
  // prologue: (obs - mean) / std; columns < nref -> enc_in, the rest -> dec_in columns latent.. (dec_in null: everything to enc_in)
  const float* obs; int nobs, nref, latent;
  const float* mean; const float* stdv;
  float* enc_in; int ld_enc;
  float* dec_in; int ld_dec;
  // kind 1
  const float* eps_latent; int deterministic;
  float* out_mean; float* out_logvar;
  // kind 2
  int na; const float* eps_action;
  float* action; float* raw_action; float* log_prob; float* logits;
  // kind 3: value network, column 0 of the last Dense
  float* value;
  ChainLayer L[kChainMaxLayers];
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait that parks the thread in hardware (suspend-time hint) instead of spinning: ncu showed 45 % of the launch's executed
// instructions in the epilogue warps' try_wait loops, stealing issue slots from the warps that had work
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity), "r"(20000u)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
               ::"r"(smem_u32(bar)), "r"(cta_rank) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kChainEpiThreads) : "memory"); }

// NormalTanhDistribution row (the body of action_head_kernel): one warp, lanes over the action dimensions
__device__ __forceinline__ void action_row(const float* l /* written earlier in this launch: no read-only path */, int na, const float* __restrict__ eps, int deterministic, size_t row, int lane,
                                           float* __restrict__ action, float* __restrict__ raw_action, float* __restrict__ log_prob,
                                           float* __restrict__ logits) {
  float lp = 0.f;
  for (int i = lane; i < na; i += 32) {
    const float loc = l[i], scale = softplus(l[na + i]) + 0.001f;
    const float raw = deterministic ? loc : loc + scale * eps[row * na + i];
    action[row * na + i] = tanhf(raw);
    if (raw_action) raw_action[row * na + i] = raw;
    const float zn = (raw - loc) / scale;
    lp += -0.5f * zn * zn - logf(scale) - 0.91893853320467274f - 2.f * (0.69314718055994531f - raw - softplus(-2.f * raw));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
  if (log_prob && lane == 0) log_prob[row] = lp;
  if (logits) for (int i = lane; i < 2 * na; i += 32) logits[row * 2 * na + i] = l[i];
}

// NormalTanhDistribution rows of the fused launch: a warp owns `nrows` (<= 8) rows first, first + stride, ...; the rows x na elements are spread
// flat over the lanes (10 rounds for 8 x 38 instead of 16 half-empty ones), every load is issued before the first use, and the
// transcendental parts use the MUFU forms (ex2 / lg2 / rcp: absolute error ~1e-7 on these O(1) terms; the per-layer launch path keeps
// the libm forms and tests/test_gpu_policy.py compares the two) -- the libm version was 80 KB of unrolled code fetched through a 32 KB
// instruction cache by every CTA at the same moment (ncu: 63 % of its stall samples `no_inst`), 20 us per launch.
__device__ __forceinline__ float softplus_fast(float x) { return fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }
__device__ __forceinline__ void action_rows_flat(const float* lg, int ld, int na, const float* __restrict__ eps, int deterministic, size_t first, int stride,
                                                 int M, int lane, float* lp_tile, float* __restrict__ action, float* __restrict__ raw_action,
                                                 float* __restrict__ log_prob, float* __restrict__ logits) {
  constexpr int kRounds = 10;          // 8 rows x na <= 320 elements
  const int tot = 8 * na;
  float loc[kRounds], rsc[kRounds], ep[kRounds];
#pragma unroll
  for (int k = 0; k < kRounds; ++k) {
    const int e = lane + 32 * k, t = e / na, i = e - t * na;
    const size_t gr = first + size_t(t) * stride;
    const bool ok = e < tot && gr < size_t(M);
    loc[k] = ok ? lg[gr * ld + i] : 0.f;
    rsc[k] = ok ? lg[gr * ld + na + i] : 0.f;
    ep[k] = ok && !deterministic ? eps[gr * na + i] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < kRounds; ++k) {
    const int e = lane + 32 * k, t = e / na, i = e - t * na;
    const size_t gr = first + size_t(t) * stride;
    if (e >= tot) continue;
    float lp = 0.f;
    if (gr < size_t(M)) {
      const float scale = softplus_fast(rsc[k]) + 0.001f;
      const float raw = deterministic ? loc[k] : fmaf(scale, ep[k], loc[k]);
      const float a2 = __expf(-2.f * fabsf(raw));
      action[gr * na + i] = copysignf(__fdividef(1.f - a2, 1.f + a2), raw);
      if (raw_action) raw_action[gr * na + i] = raw;
      const float zn = __fdividef(raw - loc[k], scale);
      lp = -0.5f * zn * zn - __logf(scale) - 0.91893853320467274f - 2.f * (0.69314718055994531f - raw - softplus_fast(-2.f * raw));
      if (logits) { logits[gr * 2 * na + i] = loc[k]; logits[gr * 2 * na + na + i] = rsc[k]; }
    }
    lp_tile[e] = lp;
  }
  __syncwarp();
  if (log_prob && lane < 8) {
    const size_t gr = first + size_t(lane) * stride;
    if (gr < size_t(M)) {
      float acc = 0.f;
      for (int i = 0; i < na; ++i) acc += lp_tile[lane * na + i];
      log_prob[gr] = acc;
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kChainThreads, 1) mlp_chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kChainStages], bar_empty[kChainStages], bar_free[kChainStages], bar_tfull[2], bar_tempty[2], bar_cready[4];
  __shared__ uint32_t tmem_slot;
  __shared__ float2 red[4][128];   // single-buffered: every layer starts with an epi_barrier (bias staging), so no warp writes layer l + 1's
                                   // partial sums before every warp has read layer l's
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < kChainStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); mbar_init(&bar_free[s], uint32_t(P.csz)); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_tfull[s], 1); mbar_init(&bar_tempty[s], 16); }
    for (int s = 0; s < 4; ++s) mbar_init(&bar_cready[s], 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  uint32_t crank = 0;
  if (P.csz > 1) {   // every CTA's barriers exist before a peer multicasts into this CTA or arrives on them
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    cluster_sync_all();
  }

  if (tid == 0) {
    // ---------------------------------------------------------------------------------------------------- TMA producer
    if (P.stagger_ns && (blockIdx.x & 1)) __nanosleep(P.stagger_ns);
    uint32_t it = 0, ruse[4] = {0u, 0u, 0u, 0u};
    for (int l = 0; l < P.n_layers; ++l) {
      const ChainLayer& L = P.L[l];
      const int nk = L.kpad / BK, ns = nk * (L.npad / L.cw), pre = min(ns, kChainStages);
      // readiness of the input slab, per 256- (128-) column chunk of the layer that produced it: this layer's first K slices only need
      // the producer's first chunk, which was stored while its later chunks were still in the tensor core -- the MMAs run across the
      // layer boundary instead of waiting for the previous layer's last epilogue pass
      const bool chunked = l > 0 && P.L[l - 1].kind == 0;
      const int in_chunks = chunked ? P.L[l - 1].npad / P.L[l - 1].cw : 1, in_cw = chunked ? P.L[l - 1].cw : (1 << 30);
      int have = 0;
      auto need = [&](int kt) {
        const int rc = min(in_chunks - 1, (kt * BK) / in_cw);
        while (have <= rc) { mbar_wait(&bar_cready[have], ruse[have] & 1u); ++ruse[have]; ++have; }
      };
      const uint32_t bytes = uint32_t(kChainABytes + L.cw * BK * 4);
      auto acquire = [&](uint32_t i) {
        const uint32_t s = i % kChainStages;
        if (i >= kChainStages) {
          const uint32_t par = ((i / kChainStages) - 1u) & 1u;
          mbar_wait(&bar_empty[s], par);
          if (P.csz > 1) {   // a peer's multicast lands in this stage of EVERY CTA: it is free when all of them have released it
            for (int r = 0; r < P.csz; ++r) mbar_arrive_remote(&bar_free[s], uint32_t(r));
            mbar_wait(&bar_free[s], par);
          }
        }
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bar_full[s])), "r"(bytes) : "memory");
        return s;
      };
      const int share = L.cw / P.csz;        // weight rows this CTA fetches per slice
      auto load_w = [&](uint32_t s, int j) {
        const uint32_t sb = sbase + s * kChainStage + kChainABytes;
        if (P.csz == 1) tma_load_2d(sb, &L.mapW, (j % nk) * BK, (j / nk) * L.cw, &bar_full[s]);
        else tma_load_2d_mc(sb + crank * share * (BK * 4), &L.mapW, (j % nk) * BK, (j / nk) * L.cw + int(crank) * share, &bar_full[s], uint16_t((1u << P.csz) - 1u));
      };
      if ((P.dbg & 4) && P.csz == 1) {
        for (int j = 0; j < ns; ++j, ++it) {
          const uint32_t s = acquire(it);
          if (j == 0) need(nk - 1);
          asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_full[s])), "r"(bytes) : "memory");
        }
        continue;
      }
      // the weight slices of the first stages do not depend on the previous layer: they fly while its epilogue is still running
      for (int j = 0; j < pre; ++j) {
        const uint32_t s = acquire(it + j);
        load_w(s, j);
      }
      for (int j = 0; j < pre; ++j) {
        need(j % nk);
        if (j == 0 && P.trace && blockIdx.x == 0) P.trace[l * 8 + 0] = clock64();
        tma_load_2d(sbase + ((it + j) % kChainStages) * kChainStage, &L.mapX, (j % nk) * BK, m0, &bar_full[(it + j) % kChainStages]);
      }
      it += pre;
      for (int j = pre; j < ns; ++j, ++it) {
        const uint32_t s = acquire(it);
        const uint32_t sa = sbase + s * kChainStage;
        load_w(s, j);
        if (j < nk) need(j);
        tma_load_2d(sa, &L.mapX, (j % nk) * BK, m0, &bar_full[s]);
      }
      need(nk - 1);   // (a one-chunk layer shorter than the prefetch depth)
    }
  } else if (tid == 32) {
    // ---------------------------------------------------------------------------------------------------- MMA issuer
    uint32_t it = 0, use[2] = {0u, 0u};
    for (int l = 0; l < P.n_layers; ++l) {
      const ChainLayer& L = P.L[l];
      const int nk = L.kpad / BK, nch = L.npad / L.cw;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(L.cw >> 3) << 17) | (uint32_t(128 >> 4) << 24);
      for (int c = 0; c < nch; ++c) {
        const int slot = c & 1;
        if (use[slot]) mbar_wait(&bar_tempty[slot], (use[slot] - 1u) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kt = 0; kt < nk; ++kt, ++it) {
          const uint32_t s = it % kChainStages;
          mbar_wait(&bar_full[s], (it / kChainStages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (P.trace && blockIdx.x == 0 && c == 0 && kt == 0) P.trace[l * 8 + 1] = clock64();
          const uint32_t sa = sbase + s * kChainStage, sb = sa + kChainABytes;
          if (!(P.dbg & 1)) {
#pragma unroll
            for (int j = 0; j < BK / 8; ++j)
              mma_tf32(tmem + uint32_t(slot * 256), make_desc_sw128(sa + j * 32), make_desc_sw128(sb + j * 32), idesc, (kt > 0 || j > 0) ? 1u : 0u);
          }
          mma_commit(&bar_empty[s]);
        }
        mma_commit(&bar_tfull[slot]);
        if (P.trace && blockIdx.x == 0 && c == nch - 1) P.trace[l * 8 + 2] = clock64();
        ++use[slot];
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------------------------------- epilogue warps
    const int ew = warp - 4, q = warp & 3, g = ew >> 2;      // TMEM lane quarter (= warp % 4), column group
    const int rl = q * 32 + lane;
    const size_t row0 = size_t(m0) + q * 32;                 // first row of this warp's TMEM lane quarter
    const int rows_valid = max(0, min(32, P.M - int(row0)));
    const uint32_t tlane = uint32_t(q * 32) << 16;
    float* tile = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + kChainStages * kChainStage) + ew * (32 * 32);
    float* sbias = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + kChainStages * kChainStage + 16 * kChainTileBytes);
    float* scvec = sbias + kChainMaxN;
    // prologue: normalised observations of the CTA's 128 rows; one warp per row, two rows' float4 loads in flight per thread,
    // 16- / 8-byte stores (scalar stores made this phase LSU-bound: 49 k partial-sector requests per CTA, 52 us)
    if (P.obs && !(P.dbg & 8)) {
      const int n4 = P.nobs >> 2, shift = P.latent - P.nref;
      if ((P.nobs & 3) == 0 && n4 <= 192) {
        const bool dec2 = P.dec_in && !(shift & 1) && !(P.ld_dec & 1);
        for (int r = ew; r < 128; r += 32) {
          float4 a[2][6];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const size_t gr = size_t(m0) + r + 16 * u;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              const int idx = lane + 32 * k;
              if (gr < size_t(P.M) && idx < n4) a[u][k] = __ldg(reinterpret_cast<const float4*>(P.obs + gr * P.nobs) + idx);
            }
          }
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int idx = lane + 32 * k, i0 = idx * 4;
            if (idx >= n4) continue;
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(P.mean) + idx), s4 = __ldg(reinterpret_cast<const float4*>(P.stdv) + idx);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const size_t gr = size_t(m0) + r + 16 * u;
              if (gr >= size_t(P.M)) continue;
              const float o[4] = {(a[u][k].x - m4.x) / s4.x, (a[u][k].y - m4.y) / s4.y, (a[u][k].z - m4.z) / s4.z, (a[u][k].w - m4.w) / s4.w};
              if (i0 + 3 < P.nref || !P.dec_in) {
                *reinterpret_cast<float4*>(P.enc_in + gr * P.ld_enc + i0) = make_float4(o[0], o[1], o[2], o[3]);
              } else if (i0 >= P.nref && dec2) {
                float* d = P.dec_in + gr * P.ld_dec + shift + i0;
                *reinterpret_cast<float2*>(d) = make_float2(o[0], o[1]);
                *reinterpret_cast<float2*>(d + 2) = make_float2(o[2], o[3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int i = i0 + e;
                  if (i < P.nref) P.enc_in[gr * P.ld_enc + i] = o[e];
                  else P.dec_in[gr * P.ld_dec + shift + i] = o[e];
                }
              }
            }
          }
        }
      } else {
        for (int r = ew; r < 128; r += 16) {
          const size_t gr = size_t(m0) + r;
          if (gr >= size_t(P.M)) break;
          const float* o = P.obs + gr * P.nobs;
          for (int i = lane; i < P.nobs; i += 32) {
            const float v = (o[i] - P.mean[i]) / P.stdv[i];
            if (i < P.nref || !P.dec_in) P.enc_in[gr * P.ld_enc + i] = v;
            else P.dec_in[gr * P.ld_dec + shift + i] = v;
          }
        }
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_cready[0]);

    uint32_t use[2] = {0u, 0u};
    float rs_in = 1.f, mr_in = 0.f;     // LayerNorm of this row's input, applied through the folded weights: x_ln W = rs (s W') - rs mean (W^T g) + ...
    for (int l = 0; l < P.n_layers; ++l) {
      const ChainLayer& L = P.L[l];
      const int nch = L.npad / L.cw, per = L.cw >> 2, nb = (P.dbg & 2) ? 0 : (per >> 5);   // columns per group in a chunk (64 / 32), x32 batches
      float s1 = 0.f, s2 = 0.f;
      // this layer's bias / cvec rows into shared memory (broadcast LDS in the epilogue instead of 32 L2-latency loads per 32 columns:
      // with 222 KB of shared memory the L1 is too small to keep them); overlaps the first chunk's MMAs
      for (int i = tid - 128; i < L.npad; i += kChainEpiThreads) { sbias[i] = __ldg(L.bias + i); scvec[i] = L.cvec ? __ldg(L.cvec + i) : 0.f; }
      epi_barrier();
      for (int c = 0; c < nch; ++c) {
        const int slot = c & 1;
        if (lane == 0) mbar_wait_park(&bar_tfull[slot], use[slot] & 1u);
        __syncwarp();
        ++use[slot];
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (P.trace && blockIdx.x == 0 && tid == 128 && (c == 0 || c == nch - 1)) P.trace[l * 8 + (c == nch - 1 ? 4 : 3)] = clock64();
        for (int b = 0; b < nb; ++b) {
          const int col = c * L.cw + g * per + b * 32;               // column of the layer
          uint32_t v[32];
          if (!(P.dbg & 128)) tmem_ld32(tmem + tlane + uint32_t(slot * 256 + g * per + b * 32), v);
          else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          const float4* bb = reinterpret_cast<const float4*>(sbias + col);
          const float4* cc = reinterpret_cast<const float4*>(scvec + col);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = bb[j >> 2], c4 = cc[j >> 2];
            v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), rs_in, fmaf(-mr_in, c4.x, b4.x)));
            v[j + 1] = __float_as_uint(fmaf(__uint_as_float(v[j + 1]), rs_in, fmaf(-mr_in, c4.y, b4.y)));
            v[j + 2] = __float_as_uint(fmaf(__uint_as_float(v[j + 2]), rs_in, fmaf(-mr_in, c4.z, b4.z)));
            v[j + 3] = __float_as_uint(fmaf(__uint_as_float(v[j + 3]), rs_in, fmaf(-mr_in, c4.w, b4.w)));
          }
          if (L.kind == 3 && col == 0 && lane < rows_valid) P.value[row0 + lane] = __uint_as_float(v[0]);
          if (L.save_h) store_tile(v, tile, lane, L.save_h, L.ldh, row0, rows_valid, col);
          if (L.act && !(P.dbg & 64)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(silu_fast(__uint_as_float(v[j])));
          }
          if (L.ln) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float x = __uint_as_float(v[j]); s1 += x; s2 = fmaf(x, x, s2); }
          }
          if (!(P.dbg & 32)) store_tile(v, tile, lane, L.out, L.ldo, row0, rows_valid, col);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (L.kind == 0) asm volatile("fence.proxy.async;" ::: "memory");   // plain layer: this chunk of the output slab may be loaded by the next layer
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bar_tempty[slot]); if (L.kind == 0) mbar_arrive(&bar_cready[c]); }
      }
      if (P.trace && blockIdx.x == 0 && tid == 128) P.trace[l * 8 + 5] = clock64();
      if (L.kind != 0 && (P.dbg & 16)) {
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_cready[0]);
      }
      if (L.ln) {
        red[g][rl] = make_float2(s1, s2);
        epi_barrier();
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 r2 = red[k][rl]; t1 += r2.x; t2 += r2.y; }
        const float mean = t1 / float(L.n), var = fmaxf(0.f, t2 / float(L.n) - mean * mean);
        rs_in = rsqrtf(var + 1e-6f);
        mr_in = rs_in * mean;
      } else {
        rs_in = 1.f; mr_in = 0.f;
        epi_barrier();   // (the LayerNorm branch has one): no warp restages sbias / scvec while another still reads this layer's
      }
      if (P.dbg & 16) continue;
      if (L.kind == 1) {
        // z = mean + exp(logvar / 2) eps into the decoder input; the head's 2 x latent columns were written by all column groups
        epi_barrier();
        const int et = tid - 128, lat = P.latent, tot = 128 * lat;
        for (int i0 = et; i0 < tot; i0 += 4 * kChainEpiThreads) {
          float mu[4], lv[4], ep[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kChainEpiThreads, r = i / lat, j = i - r * lat;
            const size_t gr = size_t(m0) + r;
            const bool ok = i < tot && gr < size_t(P.M);
            mu[u] = ok ? L.out[gr * L.ldo + j] : 0.f;
            lv[u] = ok ? L.out[gr * L.ldo + lat + j] : 0.f;
            ep[u] = ok && !P.deterministic ? P.eps_latent[gr * lat + j] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kChainEpiThreads, r = i / lat, j = i - r * lat;
            const size_t gr = size_t(m0) + r;
            if (i >= tot || gr >= size_t(P.M)) continue;
            P.dec_in[gr * P.ld_dec + j] = P.deterministic ? mu[u] : mu[u] + ep[u] * expf(0.5f * lv[u]);
            if (P.out_mean) P.out_mean[gr * lat + j] = mu[u];
            if (P.out_logvar) P.out_logvar[gr * lat + j] = lv[u];
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_cready[0]);
      } else if (L.kind != 0) {
        if (L.kind == 2) {
          epi_barrier();
          action_rows_flat(L.out, L.ldo, P.na, P.eps_action, P.deterministic, size_t(m0) + ew, 16, P.M, lane, tile, P.action, P.raw_action, P.log_prob, P.logits);
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_cready[0]);
      }
      if (P.trace && blockIdx.x == 0 && tid == 128) P.trace[l * 8 + 6] = clock64();
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (P.csz > 1) cluster_sync_all();   // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace tmjx_policy
