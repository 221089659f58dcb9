/*
 * tmjx_train.cuh — backward pass of the intention network and the value network (BASELINE configs[3]); included at
 * the end of tmjx_policy.cu (it uses that file's tcgen05 GEMM, Layer tables and row kernels).
 *
 * Replaces, for one PPO minibatch, what `jax.value_and_grad(compute_ppo_loss)` differentiates through the networks:
 *   gradient_update_fn / loss_and_pgrad      reference track_mjx/agent/mlp_ppo/ppo.py:263-272, 621-623 (upstream brax gradients.py)
 *   IntentionNetwork.__call__ (train mode)    reference track_mjx/agent/mlp_ppo/intention_network.py:90-142
 *   make_value_network(...).apply             reference track_mjx/agent/mlp_ppo/ppo_networks.py:180-185
 * The loss head (tmjx_ppo_loss_head) supplies d loss / d logits, d latent mean / logvar, d baseline; this file carries them back to
 * every parameter.  Hidden layer (Dense -> SiLU -> LayerNorm): the forward keeps the pre-activation H = x W + b and the output
 * A; the backward is
 *     row kernel   dH = LayerNorm'(SiLU(H)) * SiLU'(H) applied to dA;  column sums give d bias, d LayerNorm scale / bias
 *     wgrad        dW[k, n]  = x^T dH      tcgen05 GEMM with M = k, N = n, K = rows (operands = transposed copies of x and dH)
 *     dgrad        dx[rows,k] = dH W^T      tcgen05 GEMM with the un-transposed weights as the K-major B operand
 * all on the same `linear_tf32_tma_kernel` as the forward (kind::tf32, fp32 accumulation in TMEM).  Gradients land in ONE flat fp32
 * buffer with the layout of the flat parameter vector (policy vector, then value vector; the normaliser entries stay zero), so the
 * NCCL all-reduce and tmjx_adam_step see a single tensor.  Every reduction has a fixed order: results are bitwise reproducible.
 * The checker is oracle/mlp_grad.py (float64, equal to torch autograd to 1e-9); tests/test_gpu_train.py states the TF32 tolerance.
 */
struct BwdScratch;
namespace tmjx_policy {

constexpr int kTrainLd = 1024;       // row pitch of the gradient-activation buffers (widest layer / padded fan-in)
constexpr int kBwdWarps = 8;         // warps per block of the row kernels (8 x 3 x 1024 floats of column accumulators = 96 KB; 4 warps left the SMs at 8 resident warps: latency-bound)
constexpr int kBwdBlocks = 296;      // 2 x 148: partial column sums per block, reduced in a fixed order
constexpr int kWgradMaxSplits = 64;  // split-K planes of the wgrad GEMM (scratch: kWgradMaxSplits x the largest padded kernel)

__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

// A = LayerNorm(SiLU(H)) (has_ln) or SiLU(H); one warp per row.  Same arithmetic as the inference path (silu_fast in the GEMM
// epilogue + layernorm_kernel), so the acting policy and the training forward agree on identical parameters.
__global__ void silu_ln_fwd_kernel(const float* __restrict__ H, int ldh, int n, const float* __restrict__ scale, const float* __restrict__ bias,
                                   int has_ln, float* __restrict__ A, int lda, int M) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* h = H + size_t(row) * ldh;
  float* a = A + size_t(row) * lda;
  float s = 0.f, s2 = 0.f;
  for (int i = lane * 4; i < n; i += 128) {
    float4 v = *reinterpret_cast<const float4*>(h + i);
    v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w);
    s += v.x + v.y + v.z + v.w;
    s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    if (!has_ln) *reinterpret_cast<float4*>(a + i) = v;
  }
  if (!has_ln) return;
#pragma unroll
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const float mean = s / float(n), var = fmaxf(0.f, s2 / float(n) - mean * mean), rstd = rsqrtf(var + 1e-6f);
  for (int i = lane * 4; i < n; i += 128) {
    float4 v = *reinterpret_cast<const float4*>(h + i);
    const float4 g = *reinterpret_cast<const float4*>(scale + i), b = *reinterpret_cast<const float4*>(bias + i);
    v.x = (silu_fast(v.x) - mean) * rstd * g.x + b.x; v.y = (silu_fast(v.y) - mean) * rstd * g.y + b.y;
    v.z = (silu_fast(v.z) - mean) * rstd * g.z + b.z; v.w = (silu_fast(v.w) - mean) * rstd * g.w + b.w;
    *reinterpret_cast<float4*>(a + i) = v;
  }
}

// Backward of A = LayerNorm(SiLU(H)) * scale + bias (or A = SiLU(H)): dH from dA, one warp per row, rows strided over the grid.
//   s = SiLU(h), shat = (s - mean) rstd, g = dA scale:   ds = rstd (g - mean(g) - shat mean(g shat)),   dh = ds SiLU'(h)
// Column sums (d Dense bias = sum dh, d LN scale = sum dA shat, d LN bias = sum dA) are accumulated per warp in shared memory in
// row order, the warps of a block are added in warp order and the block's partial goes to partial[block][3][npad].
// Register form of ln_silu_bwd_kernel for rows of at most 128 NCH columns (every hidden layer of the shipped networks but the encoder's
// 1024): a lane keeps its NCH float4 chunks of h and dA in registers, so the row is read ONCE and SiLU / sigmoid are evaluated once
// (the shared-memory form below makes three passes over h and read-modify-writes three column accumulators per element through
// shared memory); the column sums live in registers over the warp's rows and meet in shared memory once, in warp order, at the end.
template <int NCH>
__global__ void __launch_bounds__(32 * kBwdWarps) ln_silu_bwd_reg_kernel(const float* __restrict__ dA, int ldda, const float* __restrict__ H, int ldh,
                                                                          int n, int npad, const float* __restrict__ scale, int has_ln,
                                                                          float* __restrict__ dH, int lddh, float* __restrict__ partial, int M) {
  extern __shared__ float acc[];   // [kBwdWarps][3][npad], written once
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_n = 1.f / float(n);
  float sc[NCH][4], ab[NCH][4], as_[NCH][4], al[NCH][4];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int i = lane * 4 + 128 * k;
    float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
    if (has_ln && i < n) v = *reinterpret_cast<const float4*>(scale + i);
    sc[k][0] = v.x; sc[k][1] = v.y; sc[k][2] = v.z; sc[k][3] = v.w;
#pragma unroll
    for (int c = 0; c < 4; ++c) ab[k][c] = as_[k][c] = al[k][c] = 0.f;
  }
  for (int row = blockIdx.x * kBwdWarps + warp; row < M; row += gridDim.x * kBwdWarps) {
    const float* h = H + size_t(row) * ldh;
    const float* da = dA + size_t(row) * ldda;
    float* dh = dH + size_t(row) * lddh;
    float hv[NCH][4], dv[NCH][4], sg[NCH][4];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int i = lane * 4 + 128 * k;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (i < n) { a = *reinterpret_cast<const float4*>(h + i); b = *reinterpret_cast<const float4*>(da + i); }
      hv[k][0] = a.x; hv[k][1] = a.y; hv[k][2] = a.z; hv[k][3] = a.w;
      dv[k][0] = b.x; dv[k][1] = b.y; dv[k][2] = b.z; dv[k][3] = b.w;
    }
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sg[k][c] = sigmoid_fast(hv[k][c]);
        const float a = lane * 4 + 128 * k < n ? hv[k][c] * sg[k][c] : 0.f;
        s += a; s2 = fmaf(a, a, s2);
      }
    float mean = 0.f, rstd = 1.f, m1 = 0.f, m2 = 0.f;
    if (has_ln) {
#pragma unroll
      for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      mean = s * inv_n;
      rstd = rsqrtf(fmaxf(0.f, s2 * inv_n - mean * mean) + 1e-6f);
      float g1 = 0.f, g2 = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (lane * 4 + 128 * k >= n) continue;
          const float shat = (hv[k][c] * sg[k][c] - mean) * rstd, g = dv[k][c] * sc[k][c];
          g1 += g; g2 = fmaf(g, shat, g2);
          as_[k][c] = fmaf(dv[k][c], shat, as_[k][c]);   // d LN scale
          al[k][c] += dv[k][c];                          // d LN bias
        }
#pragma unroll
      for (int o = 16; o; o >>= 1) { g1 += __shfl_xor_sync(0xffffffffu, g1, o); g2 += __shfl_xor_sync(0xffffffffu, g2, o); }
      m1 = g1 * inv_n; m2 = g2 * inv_n;
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int i = lane * 4 + 128 * k;
      if (i >= npad) continue;
      float o4[4] = {0.f, 0.f, 0.f, 0.f};
      if (i < n) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float ds = dv[k][c];
          if (has_ln) { const float shat = (hv[k][c] * sg[k][c] - mean) * rstd; ds = rstd * (dv[k][c] * sc[k][c] - m1 - shat * m2); }
          o4[c] = ds * (sg[k][c] * (1.f + hv[k][c] * (1.f - sg[k][c])));
          ab[k][c] += o4[c];                               // d Dense bias
        }
      }
      *reinterpret_cast<float4*>(dh + i) = make_float4(o4[0], o4[1], o4[2], o4[3]);   // padding columns (n <= i < npad) stay zero for the GEMMs
    }
  }
  float* my = acc + size_t(warp) * 3 * npad;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int i = lane * 4 + 128 * k;
    if (i >= npad) continue;
    *reinterpret_cast<float4*>(my + i) = make_float4(ab[k][0], ab[k][1], ab[k][2], ab[k][3]);
    *reinterpret_cast<float4*>(my + npad + i) = make_float4(as_[k][0], as_[k][1], as_[k][2], as_[k][3]);
    *reinterpret_cast<float4*>(my + 2 * npad + i) = make_float4(al[k][0], al[k][1], al[k][2], al[k][3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * npad; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kBwdWarps; ++w2) t += acc[size_t(w2) * 3 * npad + i];
    partial[size_t(blockIdx.x) * 3 * npad + i] = t;
  }
}

__global__ void __launch_bounds__(32 * kBwdWarps) ln_silu_bwd_kernel(const float* __restrict__ dA, int ldda, const float* __restrict__ H, int ldh,
                                                                      int n, int npad, const float* __restrict__ scale, int has_ln,
                                                                      float* __restrict__ dH, int lddh, float* __restrict__ partial, int M) {
  extern __shared__ float acc[];   // [kBwdWarps][3][npad]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = acc + size_t(warp) * 3 * npad;
  for (int i = lane; i < 3 * npad; i += 32) my[i] = 0.f;
  __syncwarp();
  const float inv_n = 1.f / float(n);
  for (int row = blockIdx.x * kBwdWarps + warp; row < M; row += gridDim.x * kBwdWarps) {
    const float* h = H + size_t(row) * ldh;
    const float* da = dA + size_t(row) * ldda;
    float* dh = dH + size_t(row) * lddh;
    float mean = 0.f, rstd = 1.f, m1 = 0.f, m2 = 0.f;
    if (has_ln) {
      float s = 0.f, s2 = 0.f;
      for (int i = lane * 4; i < n; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(h + i);
        const float a0 = silu_fast(v.x), a1 = silu_fast(v.y), a2 = silu_fast(v.z), a3 = silu_fast(v.w);
        s += a0 + a1 + a2 + a3;
        s2 += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      mean = s * inv_n;
      rstd = rsqrtf(fmaxf(0.f, s2 * inv_n - mean * mean) + 1e-6f);
      float g1 = 0.f, g2 = 0.f;
      for (int i = lane * 4; i < n; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(h + i), d = *reinterpret_cast<const float4*>(da + i),
                     sc = *reinterpret_cast<const float4*>(scale + i);
        const float hv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w}, sv[4] = {sc.x, sc.y, sc.z, sc.w};
        float4 as = *reinterpret_cast<float4*>(my + npad + i), ab = *reinterpret_cast<float4*>(my + 2 * npad + i);
        float ds4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float shat = (silu_fast(hv[c]) - mean) * rstd, g = dv[c] * sv[c];
          g1 += g; g2 += g * shat;
          ds4[c] = dv[c] * shat;
        }
        as.x += ds4[0]; as.y += ds4[1]; as.z += ds4[2]; as.w += ds4[3];       // d LN scale
        ab.x += dv[0]; ab.y += dv[1]; ab.z += dv[2]; ab.w += dv[3];           // d LN bias
        *reinterpret_cast<float4*>(my + npad + i) = as;
        *reinterpret_cast<float4*>(my + 2 * npad + i) = ab;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) { g1 += __shfl_xor_sync(0xffffffffu, g1, o); g2 += __shfl_xor_sync(0xffffffffu, g2, o); }
      m1 = g1 * inv_n; m2 = g2 * inv_n;
    }
    for (int i = lane * 4; i < n; i += 128) {
      const float4 v = *reinterpret_cast<const float4*>(h + i), d = *reinterpret_cast<const float4*>(da + i);
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
      if (has_ln) sc = *reinterpret_cast<const float4*>(scale + i);
      const float hv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w}, sv[4] = {sc.x, sc.y, sc.z, sc.w};
      float o4[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float sg = sigmoid_fast(hv[c]);
        float ds = dv[c];
        if (has_ln) { const float shat = (hv[c] * sg - mean) * rstd; ds = rstd * (dv[c] * sv[c] - m1 - shat * m2); }
        o4[c] = ds * (sg * (1.f + hv[c] * (1.f - sg)));
      }
      float4 a0 = *reinterpret_cast<float4*>(my + i);
      a0.x += o4[0]; a0.y += o4[1]; a0.z += o4[2]; a0.w += o4[3];             // d Dense bias
      *reinterpret_cast<float4*>(my + i) = a0;
      *reinterpret_cast<float4*>(dh + i) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
    // the GEMMs that consume dH read whole 32-float K slices up to npad: keep the padding columns zero
    for (int i = n + lane; i < npad; i += 32) dh[i] = 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * npad; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kBwdWarps; ++w2) t += acc[size_t(w2) * 3 * npad + i];
    partial[size_t(blockIdx.x) * 3 * npad + i] = t;
  }
}

// column sums of a plain [M, ld] matrix (d bias of the linear heads): per-block partials in the same layout (slot 0 only).
// Block = 32 columns x 8 row lanes; rows strided over (blockIdx.y, row lane); the 8 row lanes are added in lane order.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, int ld, int n, int npad, float* __restrict__ partial, int M) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
  float t = 0.f;
  if (c < n)
    for (int r = blockIdx.y * 8 + ty; r < M; r += gridDim.y * 8) t += X[size_t(r) * ld + c];
  red[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && c < n) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a += red[j][tx];
    partial[size_t(blockIdx.y) * 3 * npad + c] = a;
  }
}

// out_k[c] = sum over blocks (fixed order) of partial[b][k][c], k = 0 (d bias), 1 (d LN scale), 2 (d LN bias); a destination may
// be null.  The fused (mean | logvar) head splits slot 0 between two bias vectors at column n1.  Block = 32 columns x 8 lanes over
// the partial blocks (lane j adds blocks j, j + 8, ... in order, the 8 lanes are added in lane order).
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ partial, int nblk, int n, int npad, int n1,
                                                            float* __restrict__ d_bias, float* __restrict__ d_bias2, float* __restrict__ d_lns,
                                                            float* __restrict__ d_lnb) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx, k = blockIdx.y;
  float t = 0.f;
  if (c < n)
    for (int b = ty; b < nblk; b += 8) t += partial[(size_t(b) * 3 + k) * npad + c];
  red[ty][tx] = t;
  __syncthreads();
  if (ty != 0 || c >= n) return;
  float* dst = k == 0 ? (c < n1 ? d_bias : d_bias2) : (k == 1 ? d_lns : d_lnb);
  if (!dst) return;
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) a += red[j][tx];
  dst[k == 0 && c >= n1 ? c - n1 : c] = a;
}

// dst[c, r] = src[r, c] for r < M, c < ncols (32 x 32 tiles through shared memory); columns r in [M, Mpad) of dst are zeroed
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, int lds, int M, int Mpad, int ncols, float* __restrict__ dst, int ldd) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < M && c < ncols) ? src[size_t(r) * lds + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < ncols && r < Mpad) dst[size_t(c) * ldd + r] = tile[tx][j];
  }
}

// wgrad scratch (`planes` split-K planes of [k, ld], summed in plane order) -> the flat gradient vector: kernel [k, n1] (and
// [k, n - n1] for the second half of the fused head)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dWs, int ld, size_t plane_stride, int planes, int k, int n, int n1,
                                    float* __restrict__ g1, float* __restrict__ g2) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(k) * n) return;
  const int i = int(idx / n), j = int(idx % n);
  float v = 0.f;
  for (int z = 0; z < planes; ++z) v += dWs[size_t(z) * plane_stride + size_t(i) * ld + j];
  if (j < n1) g1[size_t(i) * n1 + j] = v;
  else g2[size_t(i) * (n - n1) + (j - n1)] = v;
}

// flat parameter vector -> GEMM operand layouts of one layer: wt[npad, kpad] (forward, K-major in the fan-in), wp[kNp, npad]
// (dgrad, K-major in the fan-out; may be null), bias[npad]
__global__ void repack_dense_kernel(const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                                    const float* __restrict__ b2, int k, int n, int n1, int kpad, int npad, float* __restrict__ wt,
                                    float* __restrict__ wp, float* __restrict__ bias, const float* __restrict__ lns_src,
                                    const float* __restrict__ lnb_src, float* __restrict__ lns, float* __restrict__ lnb) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < size_t(n)) {
    bias[idx] = idx < size_t(n1) ? b1[idx] : b2[idx - n1];
    if (lns) { lns[idx] = lns_src[idx]; lnb[idx] = lnb_src[idx]; }     // LayerNorm scale / bias ride along (were two more launches per layer)
  }
  if (idx >= size_t(k) * n) return;
  const int i = int(idx / n), j = int(idx % n);
  const float v = j < n1 ? W1[size_t(i) * n1 + j] : W2[size_t(i) * (n - n1) + (j - n1)];
  wt[size_t(j) * kpad + i] = v;
  if (wp) wp[size_t(i) * npad + j] = v;
}
__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
// dst[r, 0 .. npad) = src[r, 0 .. n) zero padded (gradient seeds into GEMM-shaped buffers); src pitch = n
__global__ void pad_rows_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int ldd, int npad, int M) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(M) * npad) return;
  const int r = int(idx / npad), c = int(idx % npad);
  dst[size_t(r) * ldd + c] = c < n ? src[size_t(r) * n + c] : 0.f;
}
__global__ void slice_rows_kernel(const float* __restrict__ src, int lds, int n, float* __restrict__ dst, int M) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(M) * n) return;
  const int r = int(idx / n), c = int(idx % n);
  dst[idx] = src[size_t(r) * lds + c];
}
// reparameterisation backward: z = mean + eps exp(logvar / 2) feeds the decoder, the KL term feeds mean / logvar directly
//   d head[:, 0..lat) = d_mean + dz,   d head[:, lat..2 lat) = d_logvar + dz eps exp(logvar / 2) / 2,   padding columns zero
__global__ void latent_bwd_kernel(const float* __restrict__ d_dec_in, int ldd, const float* __restrict__ d_mean, const float* __restrict__ d_logvar,
                                  const float* __restrict__ eps, const float* __restrict__ head, int ldh, int lat, float* __restrict__ d_head,
                                  int ldo, int npad, int M) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(M) * npad) return;
  const int r = int(idx / npad), c = int(idx % npad);
  float v = 0.f;
  if (c < lat) v = d_mean[size_t(r) * lat + c] + d_dec_in[size_t(r) * ldd + c];
  else if (c < 2 * lat) {
    const int j = c - lat;
    const float lv = head[size_t(r) * ldh + c];
    v = d_logvar[size_t(r) * lat + j] + d_dec_in[size_t(r) * ldd + j] * eps[size_t(r) * lat + j] * 0.5f * expf(0.5f * lv);
  }
  d_head[size_t(r) * ldo + c] = v;
}

// one stack of Dense layers (policy encoder, policy decoder, value network) with what its backward pass needs
struct TrainStack {
  std::vector<Layer>* layers = nullptr;
  const float* x0 = nullptr;           // the stack's input buffer ([max_rows, ldx0], K padding zero)
  int ldx0 = 0;
  size_t param_base = 0;               // offset of the owning network's parameter vector in the trainer's flat buffers
  ::BwdScratch* scr = nullptr;    // the owning network's backward scratch
  std::vector<float*> H, A, wp;        // pre-activations, outputs, un-transposed padded weights per layer
  std::vector<int> kNp;                // fan-in padded to the GEMM's N tile (dgrad output width)
  std::vector<CUtensorMap> mapX, mapDH, mapWp, mapXT, mapDHT;
  std::vector<float*> dWs, partial;   // per layer: split-K planes of the wgrad GEMM, per-block column-sum partials (reduced once per stack)
  std::vector<CUtensorMap> mapXmn, mapDHmn;   // MN-major wgrad operands straight from the row-major activations / gradients (no transposing copies)
};

}  // namespace tmjx_policy

// Backward scratch of ONE network.  The policy and the value network each own a set, so that their backward passes (and forwards: the
// activations are per stack anyway) can run on two streams at once: a 10240-row minibatch gives 80 CTAs per GEMM on 148 SMs, the two
// networks together fill the machine.
struct BwdScratch {
  float* dA[2] = {nullptr, nullptr};   // gradient w.r.t. a layer's output, ping-pong [max_rows, kTrainLd]
  float* dH = nullptr;                 // gradient w.r.t. a layer's pre-activation [max_rows, kTrainLd]
  float *xT = nullptr, *dhT = nullptr; // transposed operands of the wgrad GEMM [kTrainLd, rows_ld]
  float* dWs = nullptr;                // wgrad output: split-K planes of [kpad, npad]
  float* partial = nullptr;            // per-block column-sum partials
};

struct TmjxTrainer {
  TmjxPolicy* pol = nullptr;
  TmjxPolicy* val = nullptr;
  int device = 0, max_rows = 0, rows_ld = 0;
  size_t n_pol = 0, n_val = 0;
  float *params = nullptr, *grads = nullptr;
  TrainStack enc, dec, vnet;
  BwdScratch sp, sv;                   // policy (encoder + decoder) / value network
  int wgrad_mn = 3;                    // TMJX_WGRAD_MN=0: transposing copies + K-major operands (the first form; A/B)
  int bwd_smem_form = 0;               // TMJX_BWD_SMEM_FORM=1: the three-pass shared-memory row kernel everywhere (A/B)
  float *zeros = nullptr, *eps = nullptr;
  std::vector<void*> owned;
};

// splits > 1: split-K into `splits` output planes of plane_stride floats (see linear_tf32_tma_kernel); *splits_out = planes written

// ---- per-stack batched forms of the small per-layer kernels (one launch per stack instead of one per layer: a 10240-row minibatch
// update was ~60 launches of 5 - 9 us each for column sums, split-K unpacking and operand repacking).  blockIdx.z / .y = the layer;
// the arithmetic and its order are those of the per-layer kernels above.
constexpr int kMaxStackLayers = 12;
struct ColsumTask { const float* partial; float *d_bias, *d_bias2, *d_lns, *d_lnb; int nblk, n, npad, n1, nk; };
struct ColsumTable { ColsumTask t[kMaxStackLayers]; };
__global__ void __launch_bounds__(256) colsum_all_kernel(const __grid_constant__ ColsumTable T) {
  const ColsumTask& q = T.t[blockIdx.z];
  if (int(blockIdx.y) >= q.nk || int(blockIdx.x) * 32 >= q.n) return;
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx, k = blockIdx.y;
  float t = 0.f;
  if (c < q.n)
    for (int b = ty; b < q.nblk; b += 8) t += q.partial[(size_t(b) * 3 + k) * q.npad + c];
  red[ty][tx] = t;
  __syncthreads();
  if (ty != 0 || c >= q.n) return;
  float* dst = k == 0 ? (c < q.n1 ? q.d_bias : q.d_bias2) : (k == 1 ? q.d_lns : q.d_lnb);
  if (!dst) return;
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) a += red[j][tx];
  dst[k == 0 && c >= q.n1 ? c - q.n1 : c] = a;
}
struct UnpackTask { const float* dWs; float *g1, *g2; size_t plane_stride; int ld, planes, k, n, n1; };
struct UnpackTable { UnpackTask t[kMaxStackLayers]; };
__global__ void __launch_bounds__(256) unpack_all_kernel(const __grid_constant__ UnpackTable T) {
  const UnpackTask& q = T.t[blockIdx.y];
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(q.k) * q.n) return;
  const int i = int(idx / q.n), j = int(idx % q.n);
  float v = 0.f;
  for (int z = 0; z < q.planes; ++z) v += q.dWs[size_t(z) * q.plane_stride + size_t(i) * q.ld + j];
  if (j < q.n1) q.g1[size_t(i) * q.n1 + j] = v;
  else q.g2[size_t(i) * (q.n - q.n1) + (j - q.n1)] = v;
}
struct RepackTask { const float *W1, *b1, *W2, *b2, *lns_src, *lnb_src; float *wt, *wp, *bias, *lns, *lnb; int k, n, n1, kpad, npad; };
struct RepackTable { RepackTask t[kMaxStackLayers]; };
__global__ void __launch_bounds__(256) repack_all_kernel(const __grid_constant__ RepackTable T) {
  const RepackTask& q = T.t[blockIdx.y];
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < size_t(q.n)) {
    q.bias[idx] = idx < size_t(q.n1) ? q.b1[idx] : q.b2[idx - q.n1];
    if (q.lns) { q.lns[idx] = q.lns_src[idx]; q.lnb[idx] = q.lnb_src[idx]; }
  }
  if (idx >= size_t(q.k) * q.n) return;
  const int i = int(idx / q.n), j = int(idx % q.n);
  const float v = j < q.n1 ? q.W1[size_t(i) * q.n1 + j] : q.W2[size_t(i) * (q.n - q.n1) + (j - q.n1)];
  q.wt[size_t(j) * q.kpad + i] = v;
  if (q.wp) q.wp[size_t(i) * q.npad + j] = v;
}

static int train_gemm(const CUtensorMap& mapA, const CUtensorMap& mapB, const float* bias, float* Y, int ldy, int M, int Kpad, int Npad,
                      cudaStream_t st, int splits = 1, size_t plane_stride = 0, int* splits_out = nullptr, int mn = 0) {
  int nk_per = 0, nz = 1;
  if (splits > 1) {
    const int nk = Kpad / BK;
    nk_per = (nk + splits - 1) / splits;
    nz = (nk + nk_per - 1) / nk_per;
  }
  if (splits_out) *splits_out = nz;
  if (Npad >= 512) {
    dim3 grid((M + 255) / 256, Npad / 256, nz);
    if (mn == 3) linear_tf32_tma_kernel<256, 3><<<grid, kTmaThreads, V3<256>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
    else if (mn == 2) linear_tf32_tma_kernel<256, 2><<<grid, kTmaThreads, V3<256>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
    else if (mn == 1) linear_tf32_tma_kernel<256, 1><<<grid, kTmaThreads, V3<256>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
    else linear_tf32_tma_kernel<256><<<grid, kTmaThreads, V3<256>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
  } else {
    dim3 grid((M + 255) / 256, Npad / 128, nz);
    if (mn == 3) linear_tf32_tma_kernel<128, 3><<<grid, kTmaThreads, V3<128>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
    else if (mn == 2) linear_tf32_tma_kernel<128, 2><<<grid, kTmaThreads, V3<128>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
    else if (mn == 1) linear_tf32_tma_kernel<128, 1><<<grid, kTmaThreads, V3<128>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
    else linear_tf32_tma_kernel<128><<<grid, kTmaThreads, V3<128>::kSmem, st>>>(mapA, mapB, bias, Y, ldy, M, Kpad, 0, nk_per, plane_stride);
  }
  return cudaGetLastError() == cudaSuccess ? TMJX_OK : pfail(TMJX_E_CUDA, "GEMM launch failed");
}

static int repack_layers(std::vector<Layer>& layers, const float* flat, std::vector<float*>* wp, cudaStream_t st) {
  if (layers.size() > size_t(kMaxStackLayers)) return pfail(TMJX_E_ARG, "too many layers in one stack");
  RepackTable T;
  size_t most = 0;
  for (size_t l = 0; l < layers.size(); ++l) {
    Layer& L = layers[l];
    RepackTask& q = T.t[l];
    q.W1 = flat + L.off_w; q.b1 = flat + L.off_b;
    q.W2 = L.n1 < L.n ? flat + L.off_w2 : nullptr; q.b2 = L.n1 < L.n ? flat + L.off_b2 : nullptr;
    q.lns_src = L.ln ? flat + L.off_lns : nullptr; q.lnb_src = L.ln ? flat + L.off_lnb : nullptr;
    q.wt = L.wt; q.wp = wp ? (*wp)[l] : nullptr; q.bias = L.bias; q.lns = L.ln ? L.ln_scale : nullptr; q.lnb = L.ln ? L.ln_bias : nullptr;
    q.k = L.k; q.n = L.n; q.n1 = L.n1; q.kpad = L.kpad; q.npad = L.npad;
    most = std::max(most, size_t(L.k) * L.n);
  }
  repack_all_kernel<<<dim3(unsigned((most + 255) / 256), unsigned(layers.size())), 256, 0, st>>>(T);
  return cudaGetLastError() == cudaSuccess ? TMJX_OK : pfail(TMJX_E_CUDA, "repack launch failed");
}

static int stack_forward(TmjxTrainer* t, TrainStack& s, int rows, bool save, cudaStream_t st, const float** out, int* ld_out) {
  // save = false: inference-style ping-pong through the owning network's buffers is not needed here -- the same H / A buffers are
  // used and simply overwritten by the next call
  (void)save;
  const size_t m = s.layers->size();
  for (size_t l = 0; l < m; ++l) {
    Layer& L = (*s.layers)[l];
    int rc = train_gemm(s.mapX[l], L.mapW, L.bias, s.H[l], L.npad, rows, L.kpad, L.npad, st);
    if (rc) return rc;
    if (L.act) silu_ln_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(s.H[l], L.npad, L.n, L.ln_scale, L.ln_bias, L.ln, s.A[l], L.npad, rows);
  }
  const Layer& last = (*s.layers)[m - 1];
  *out = last.act ? s.A[m - 1] : s.H[m - 1];
  *ld_out = last.npad;
  return cudaGetLastError() == cudaSuccess ? TMJX_OK : pfail(TMJX_E_CUDA, "forward launch failed");
}

// split-K planes wanted for a layer's wgrad GEMM: the (M / 256) x (N / BN) output tiles x splits fill the 148 SMs
static int wgrad_splits(const Layer& L) {
  static const int sms = [] { const char* e = std::getenv("TMJX_WGRAD_SMS"); return e ? std::max(1, atoi(e)) : 74; }();
  // CTAs one wgrad GEMM aims for: half the SMs -- the policy's and the critic's backward passes run side by side on two streams (148
  // measured 2.5 % slower per update: twice the planes to write and re-read)
  const int tiles = ((L.k + 255) / 256) * (L.npad >= 512 ? L.npad / 256 : L.npad / 128);
  return std::max(1, std::min(kWgradMaxSplits, sms / tiles));
}

// backward through one stack.  dY: gradient w.r.t. the stack's output, in scr->dA[which] ([rows, kTrainLd], padding columns zero).
// On return (need_dx) scr->dA[*which] holds the gradient w.r.t. the stack's input.
static int stack_backward(TmjxTrainer* t, TrainStack& s, int rows, int* which, bool need_dx, cudaStream_t st) {
  BwdScratch& c = *s.scr;
  const int rows32 = (rows + 31) / 32 * 32;
  if (s.layers->size() > size_t(kMaxStackLayers)) return pfail(TMJX_E_ARG, "too many layers in one stack");
  ColsumTable CT;
  UnpackTable UT;
  size_t most = 0;
  for (int l = int(s.layers->size()) - 1; l >= 0; --l) {
    Layer& L = (*s.layers)[l];
    float* g = t->grads + s.param_base;
    const size_t m = s.layers->size();
    const float* dh = c.dA[*which];                          // linear layer: the incoming gradient IS d pre-activation
    const CUtensorMap* map_dh = &s.mapDH[l + m * (1 + *which)];
    const int nblk = std::min(kBwdBlocks, (rows + kBwdWarps - 1) / kBwdWarps);
    if (L.act) {
      const size_t sm = size_t(kBwdWarps) * 3 * L.npad * 4;
      if (L.npad <= 256 && !t->bwd_smem_form)
        ln_silu_bwd_reg_kernel<2><<<nblk, 32 * kBwdWarps, sm, st>>>(c.dA[*which], kTrainLd, s.H[l], L.npad, L.n, L.npad, L.ln_scale, L.ln, c.dH, kTrainLd, s.partial[l], rows);
      else if (L.npad <= 512 && !t->bwd_smem_form)
        ln_silu_bwd_reg_kernel<4><<<nblk, 32 * kBwdWarps, sm, st>>>(c.dA[*which], kTrainLd, s.H[l], L.npad, L.n, L.npad, L.ln_scale, L.ln, c.dH, kTrainLd, s.partial[l], rows);
      else
        ln_silu_bwd_kernel<<<nblk, 32 * kBwdWarps, sm, st>>>(c.dA[*which], kTrainLd, s.H[l], L.npad, L.n, L.npad, L.ln_scale, L.ln, c.dH, kTrainLd, s.partial[l], rows);
      CT.t[l] = ColsumTask{s.partial[l], g + L.off_b, nullptr, L.ln ? g + L.off_lns : nullptr, L.ln ? g + L.off_lnb : nullptr, nblk, L.n, L.npad, L.n1, L.ln ? 3 : 1};
      dh = c.dH;
      map_dh = &s.mapDH[l];
    } else {
      const int ny = std::min((rows + 7) / 8, 148);
      dim3 cg((L.n + 31) / 32, ny);
      colsum_partial_kernel<<<cg, 256, 0, st>>>(dh, kTrainLd, L.n, L.npad, s.partial[l], rows);
      CT.t[l] = ColsumTask{s.partial[l], g + L.off_b, L.n1 < L.n ? g + L.off_b2 : nullptr, nullptr, nullptr, ny, L.n, L.npad, L.n1, 1};
    }
    // wgrad: dW = x^T dH
    const float* x = l == 0 ? s.x0 : s.A[l - 1];
    const int ldx = l == 0 ? s.ldx0 : (*s.layers)[l - 1].npad;
    // MN-major operands read x and dH where they lie (whole 32-row K slices: rows % 32 == 0); else two transposing copies + K-major operands
    const int mn = rows % 32 == 0 ? t->wgrad_mn : 0;      // bit 0: x MN-major, bit 1: dH MN-major (3 = both; 1 / 2: bisecting knobs)
    if (!(mn & 1)) transpose_kernel<<<dim3(rows32 / 32, L.kpad / 32), 256, 0, st>>>(x, ldx, rows, rows32, L.kpad, c.xT, t->rows_ld);
    if (!(mn & 2)) transpose_kernel<<<dim3(rows32 / 32, L.npad / 32), 256, 0, st>>>(dh, kTrainLd, rows, rows32, L.npad, c.dhT, t->rows_ld);
    // split-K so that the (M / 256) x (N / BN) output tiles x splits fill the 148 SMs: K = the minibatch rows is the long dimension
    const int want = wgrad_splits(L);
    int planes = 1;
    const size_t dh_idx = L.act ? l : l + m * (1 + *which);       // dH, or the linear layer's incoming dA[which] (same indexing as mapDH)
    int rc = train_gemm((mn & 1) ? s.mapXmn[l] : s.mapXT[l], (mn & 2) ? s.mapDHmn[dh_idx] : s.mapDHT[l], t->zeros, s.dWs[l], L.npad, L.k, rows32, L.npad, st, want,
                        size_t(L.kpad) * L.npad, &planes, mn);
    if (rc) return rc;
    UT.t[l] = UnpackTask{s.dWs[l], g + L.off_w, L.n1 < L.n ? g + L.off_w2 : nullptr, size_t(L.kpad) * L.npad, L.npad, planes, L.k, L.n, L.n1};
    most = std::max(most, size_t(L.k) * L.n);
    // dgrad: dx = dH W^T
    if (l > 0 || need_dx) {
      rc = train_gemm(*map_dh, s.mapWp[l], t->zeros, c.dA[*which ^ 1], kTrainLd, rows, L.npad, s.kNp[l], st);
      if (rc) return rc;
      *which ^= 1;
    }
  }
  // every layer's column sums and split-K planes -> the flat gradient buffer: one launch each per stack
  const unsigned nl = unsigned(s.layers->size());
  colsum_all_kernel<<<dim3(kTrainLd / 32, 3, nl), 256, 0, st>>>(CT);
  unpack_all_kernel<<<dim3(unsigned((most + 255) / 256), nl), 256, 0, st>>>(UT);
  return cudaGetLastError() == cudaSuccess ? TMJX_OK : pfail(TMJX_E_CUDA, "backward launch failed");
}

extern "C" {

/* Refresh the GEMM operand copies of an acting policy / value network from a flat DEVICE parameter vector in the layout of
 * tmjx_policy_create / tmjx_value_create (normaliser mean, std first): call after every optimiser or normaliser update. */
int tmjx_policy_set_params(TmjxPolicy* p, const float* params_device, void* stream) {
  if (!p || !params_device) return pfail(TMJX_E_ARG, "null argument");
  PCU(cudaSetDevice(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = p->d.obs_size;
  copy_kernel<<<(D + 255) / 256, 256, 0, st>>>(params_device, p->norm_mean, D);
  copy_kernel<<<(D + 255) / 256, 256, 0, st>>>(params_device + D, p->norm_std, D);
  int rc = repack_layers(p->enc, params_device, nullptr, st);
  if (rc) return rc;
  rc = repack_layers(p->dec, params_device, nullptr, st);
  if (rc) return rc;
  fold_layers(p->enc, st);     // the fused chain kernel's LayerNorm-folded operands follow the new weights
  fold_layers(p->dec, st);
  return cudaGetLastError() == cudaSuccess ? TMJX_OK : pfail(TMJX_E_CUDA, "fold launch failed");
}

void tmjx_trainer_destroy(TmjxTrainer* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  for (void* q : t->owned) cudaFree(q);
  tmjx_policy_destroy(t->pol);
  tmjx_policy_destroy(t->val);
  delete t;
}

int tmjx_trainer_create(const TmjxPolicyDesc* pd, const TmjxValueDesc* vd, const float* policy_params, const float* value_params, int device,
                        int max_rows, TmjxTrainer** out) {
  if (!pd || !vd || !policy_params || !value_params || !out || max_rows <= 0) return pfail(TMJX_E_ARG, "null argument");
  auto* t = new TmjxTrainer();
  std::unique_ptr<TmjxTrainer, void (*)(TmjxTrainer*)> guard(t, tmjx_trainer_destroy);
  t->device = device; t->max_rows = max_rows; t->rows_ld = (max_rows + 31) / 32 * 32;
  t->n_pol = tmjx_policy_param_count(pd); t->n_val = tmjx_value_param_count(vd);
  int rc = tmjx_policy_create(pd, policy_params, t->n_pol, device, max_rows, &t->pol);
  if (rc) return rc;
  rc = tmjx_value_create(vd, value_params, t->n_val, device, max_rows, &t->val);
  if (rc) return rc;
  if (t->pol->use_v1) return pfail(TMJX_E_UNSUPPORTED, "the trainer needs the TMA GEMM (TMJX_POLICY_V1 must be unset)");
  auto alloc = [&](float** p, size_t n) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, std::max<size_t>(n, 4) * 4);
    if (e != cudaSuccess) return e;
    t->owned.push_back(*p);
    return cudaMemset(*p, 0, std::max<size_t>(n, 4) * 4);
  };
  const size_t n_all = t->n_pol + t->n_val;
  PCU(alloc(&t->params, n_all)); PCU(alloc(&t->grads, n_all));
  PCU(cudaMemcpy(t->params, policy_params, t->n_pol * 4, cudaMemcpyHostToDevice));
  PCU(cudaMemcpy(t->params + t->n_pol, value_params, t->n_val * 4, cudaMemcpyHostToDevice));
  for (BwdScratch* c : {&t->sp, &t->sv}) {
    for (int i = 0; i < 2; ++i) PCU(alloc(&c->dA[i], size_t(max_rows) * kTrainLd));
    PCU(alloc(&c->dH, size_t(max_rows) * kTrainLd));
    PCU(alloc(&c->xT, size_t(kTrainLd) * t->rows_ld)); PCU(alloc(&c->dhT, size_t(kTrainLd) * t->rows_ld));
  }
  PCU(alloc(&t->zeros, kTrainLd));
  PCU(alloc(&t->eps, size_t(max_rows) * std::max(1, pd->latent_size)));
  PCU(cudaFuncSetAttribute(ln_silu_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdWarps * 3 * kTrainLd * 4));
  PCU(cudaFuncSetAttribute(ln_silu_bwd_reg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdWarps * 3 * 256 * 4));
  PCU(cudaFuncSetAttribute(ln_silu_bwd_reg_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdWarps * 3 * 512 * 4));
  if (const char* e = std::getenv("TMJX_BWD_SMEM_FORM")) t->bwd_smem_form = atoi(e);
  if (const char* e = std::getenv("TMJX_WGRAD_MN")) t->wgrad_mn = atoi(e);
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<256>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<128>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<256>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<128>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<256>::kSmem));
  PCU(cudaFuncSetAttribute(linear_tf32_tma_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3<128>::kSmem));
  bool ok = true;
  auto setup = [&](TrainStack& s, std::vector<Layer>& layers, const float* x0, int ldx0, size_t base, BwdScratch* scr) -> cudaError_t {
    s.layers = &layers; s.x0 = x0; s.ldx0 = ldx0; s.param_base = base; s.scr = scr;
    const size_t m = layers.size();
    s.H.resize(m); s.A.resize(m); s.wp.resize(m); s.kNp.resize(m); s.dWs.resize(m); s.partial.resize(m);
    s.mapX.resize(m); s.mapDH.resize(3 * m); s.mapWp.resize(m); s.mapXT.resize(m); s.mapDHT.resize(m); s.mapXmn.resize(m); s.mapDHmn.resize(3 * m);
    for (size_t l = 0; l < m; ++l) {
      Layer& L = layers[l];
      if (L.npad > kTrainLd || L.kpad > kTrainLd) return cudaErrorInvalidValue;
      cudaError_t e = alloc(&s.H[l], size_t(max_rows) * L.npad);
      if (e != cudaSuccess) return e;
      s.A[l] = nullptr;
      if (L.act) { e = alloc(&s.A[l], size_t(max_rows) * L.npad); if (e != cudaSuccess) return e; }
      s.kNp[l] = L.k > 256 ? pad_to(L.k, 256) : pad_to(L.k, 128);
      e = alloc(&s.wp[l], size_t(s.kNp[l]) * L.npad);
      if (e != cudaSuccess) return e;
      e = alloc(&s.dWs[l], size_t(wgrad_splits(L)) * L.kpad * L.npad);
      if (e != cudaSuccess) return e;
      e = alloc(&s.partial[l], size_t(kBwdBlocks) * 3 * L.npad);
      if (e != cudaSuccess) return e;
      const float* x = l == 0 ? x0 : s.A[l - 1];
      const int ldx = l == 0 ? ldx0 : layers[l - 1].npad;
      ok = ok && encode_map(&s.mapX[l], x, max_rows, L.kpad, ldx, 256);
      // dgrad A operand = the gradient w.r.t. this layer's pre-activation, K extent npad: from dH (hidden layers) or dA[0/1] (linear)
      ok = ok && encode_map(&s.mapDH[l], scr->dH, max_rows, L.npad, kTrainLd, 256);
      ok = ok && encode_map(&s.mapDH[l + m], scr->dA[0], max_rows, L.npad, kTrainLd, 256);
      ok = ok && encode_map(&s.mapDH[l + 2 * m], scr->dA[1], max_rows, L.npad, kTrainLd, 256);
      ok = ok && encode_map(&s.mapWp[l], s.wp[l], s.kNp[l], L.npad, L.npad, s.kNp[l] >= 512 ? 256 : 128);
      ok = ok && encode_map(&s.mapXT[l], scr->xT, L.kpad, t->rows_ld, t->rows_ld, 256);
      ok = ok && encode_map(&s.mapDHT[l], scr->dhT, L.npad, t->rows_ld, t->rows_ld, L.npad >= 512 ? 256 : 128);
      ok = ok && encode_map_mn(&s.mapXmn[l], x, max_rows, L.kpad, ldx);
      const float* dsrc[3] = {scr->dH, scr->dA[0], scr->dA[1]};
      for (int q = 0; q < 3; ++q) ok = ok && encode_map_mn(&s.mapDHmn[l + q * m], dsrc[q], max_rows, L.npad, kTrainLd);
    }
    return cudaSuccess;
  };
  PCU(setup(t->enc, t->pol->enc, t->pol->enc_in, t->pol->ld_enc, 0, &t->sp));
  PCU(setup(t->dec, t->pol->dec, t->pol->dec_in, t->pol->ld_dec, 0, &t->sp));
  PCU(setup(t->vnet, t->val->enc, t->val->enc_in, t->val->ld_enc, t->n_pol, &t->sv));
  if (!ok) return pfail(TMJX_E_CUDA, "cuTensorMapEncodeTiled failed");
  for (TrainStack* s : {&t->enc, &t->dec}) { rc = repack_layers(*s->layers, t->params, &s->wp, nullptr); if (rc) return rc; }
  rc = repack_layers(*t->vnet.layers, t->params + t->n_pol, &t->vnet.wp, nullptr);
  if (rc) return rc;
  PCU(cudaDeviceSynchronize());
  *out = guard.release();
  return TMJX_OK;
}

size_t tmjx_trainer_param_count(const TmjxTrainer* t) { return t ? t->n_pol + t->n_val : 0; }
size_t tmjx_trainer_policy_param_count(const TmjxTrainer* t) { return t ? t->n_pol : 0; }

int tmjx_trainer_buffers(TmjxTrainer* t, float** params, float** grads) {
  if (!t) return pfail(TMJX_E_ARG, "null argument");
  if (params) *params = t->params;
  if (grads) *grads = t->grads;
  return TMJX_OK;
}

/* flat parameters (and the normaliser entries at the head of each network's vector) -> GEMM operand copies */
int tmjx_trainer_sync(TmjxTrainer* t, void* stream) {
  if (!t) return pfail(TMJX_E_ARG, "null argument");
  PCU(cudaSetDevice(t->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = t->pol->d.obs_size, Dv = t->val->d.obs_size;
  copy_kernel<<<(D + 255) / 256, 256, 0, st>>>(t->params, t->pol->norm_mean, D);
  copy_kernel<<<(D + 255) / 256, 256, 0, st>>>(t->params + D, t->pol->norm_std, D);
  copy_kernel<<<(Dv + 255) / 256, 256, 0, st>>>(t->params + t->n_pol, t->val->norm_mean, Dv);
  copy_kernel<<<(Dv + 255) / 256, 256, 0, st>>>(t->params + t->n_pol + Dv, t->val->norm_std, Dv);
  int rc = repack_layers(*t->enc.layers, t->params, &t->enc.wp, st);
  if (rc) return rc;
  rc = repack_layers(*t->dec.layers, t->params, &t->dec.wp, st);
  if (rc) return rc;
  return repack_layers(*t->vnet.layers, t->params + t->n_pol, &t->vnet.wp, st);
}

/* training-mode policy forward on `rows` observations: logits [rows, 2 A], latent_mean / latent_logvar [rows, L] (DEVICE); the
 * activations every layer's backward needs stay in the trainer until the next forward.  eps_latent [rows, L] ~ N(0, 1). */
int tmjx_trainer_policy_forward(TmjxTrainer* t, const float* obs, const float* eps_latent, int rows, float* logits, float* latent_mean,
                                float* latent_logvar, void* stream) {
  if (!t || !obs || !eps_latent || !logits) return pfail(TMJX_E_ARG, "null argument");
  if (rows <= 0 || rows > t->max_rows) return pfail(TMJX_E_ARG, "rows exceeds the trainer's max_rows");
  PCU(cudaSetDevice(t->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TmjxPolicy* p = t->pol;
  const TmjxPolicyDesc& d = p->d;
  obs_prep_kernel<<<rows, 256, 0, st>>>(obs, d.obs_size, d.reference_obs_size, d.latent_size, p->norm_mean, p->norm_std, p->enc_in, p->ld_enc, p->dec_in,
                                       p->ld_dec, rows);
  PCU(cudaMemcpyAsync(t->eps, eps_latent, size_t(rows) * d.latent_size * 4, cudaMemcpyDeviceToDevice, st));
  const float* y; int ldy;
  int rc = stack_forward(t, t->enc, rows, true, st, &y, &ldy);
  if (rc) return rc;
  latent_kernel<<<(rows * d.latent_size + 255) / 256, 256, 0, st>>>(y, ldy, d.latent_size, eps_latent, 0, p->dec_in, p->ld_dec, latent_mean, latent_logvar, rows);
  rc = stack_forward(t, t->dec, rows, true, st, &y, &ldy);
  if (rc) return rc;
  const int n = 2 * d.action_size;
  slice_rows_kernel<<<unsigned((size_t(rows) * n + 255) / 256), 256, 0, st>>>(y, ldy, n, logits, rows);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

/* gradient of the loss w.r.t. every policy parameter from the loss head's seeds (DEVICE, [rows, 2 A] and [rows, L]); writes the
 * policy part of the flat gradient buffer.  Must follow tmjx_trainer_policy_forward on the same rows. */
int tmjx_trainer_policy_backward(TmjxTrainer* t, const float* d_logits, const float* d_latent_mean, const float* d_latent_logvar, int rows, void* stream) {
  if (!t || !d_logits || !d_latent_mean || !d_latent_logvar) return pfail(TMJX_E_ARG, "null argument");
  if (rows <= 0 || rows > t->max_rows) return pfail(TMJX_E_ARG, "rows exceeds the trainer's max_rows");
  PCU(cudaSetDevice(t->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TmjxPolicyDesc& d = t->pol->d;
  int which = 0;
  const Layer& lg = t->dec.layers->back();
  pad_rows_kernel<<<unsigned((size_t(rows) * lg.npad + 255) / 256), 256, 0, st>>>(d_logits, 2 * d.action_size, t->sp.dA[which], kTrainLd, lg.npad, rows);
  int rc = stack_backward(t, t->dec, rows, &which, true, st);
  if (rc) return rc;
  const Layer& head = t->enc.layers->back();
  const float* headH = t->enc.H[t->enc.layers->size() - 1];
  latent_bwd_kernel<<<unsigned((size_t(rows) * head.npad + 255) / 256), 256, 0, st>>>(t->sp.dA[which], kTrainLd, d_latent_mean, d_latent_logvar, t->eps, headH,
                                                                                        head.npad, d.latent_size, t->sp.dA[which ^ 1], kTrainLd, head.npad, rows);
  which ^= 1;
  rc = stack_backward(t, t->enc, rows, &which, false, st);
  if (rc) return rc;
  PCU(cudaGetLastError());
  return TMJX_OK;
}

/* value network forward on `rows` observations (DEVICE) -> value [rows]; save != 0 keeps the activations for tmjx_trainer_value_backward
 * (the bootstrap rows of the loss, which carry no gradient, go through with save = 0 AFTER the saved call's backward, or through a
 * separate TmjxPolicy value object). */
int tmjx_trainer_value_forward(TmjxTrainer* t, const float* obs, int rows, float* value, void* stream) {
  if (!t || !obs || !value) return pfail(TMJX_E_ARG, "null argument");
  if (rows <= 0 || rows > t->max_rows) return pfail(TMJX_E_ARG, "rows exceeds the trainer's max_rows");
  PCU(cudaSetDevice(t->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TmjxPolicy* v = t->val;
  value_prep_kernel<<<rows, 256, 0, st>>>(obs, v->d.obs_size, v->norm_mean, v->norm_std, v->enc_in, v->ld_enc, rows);
  const float* y; int ldy;
  int rc = stack_forward(t, t->vnet, rows, true, st, &y, &ldy);
  if (rc) return rc;
  value_out_kernel<<<(rows + 255) / 256, 256, 0, st>>>(y, ldy, value, rows);
  PCU(cudaGetLastError());
  return TMJX_OK;
}

int tmjx_trainer_value_backward(TmjxTrainer* t, const float* d_value, int rows, void* stream) {
  if (!t || !d_value) return pfail(TMJX_E_ARG, "null argument");
  if (rows <= 0 || rows > t->max_rows) return pfail(TMJX_E_ARG, "rows exceeds the trainer's max_rows");
  PCU(cudaSetDevice(t->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int which = 0;
  const Layer& last = t->vnet.layers->back();
  pad_rows_kernel<<<unsigned((size_t(rows) * last.npad + 255) / 256), 256, 0, st>>>(d_value, 1, t->sv.dA[which], kTrainLd, last.npad, rows);
  int rc = stack_backward(t, t->vnet, rows, &which, false, st);
  if (rc) return rc;
  PCU(cudaGetLastError());
  return TMJX_OK;
}

}  // extern "C"
