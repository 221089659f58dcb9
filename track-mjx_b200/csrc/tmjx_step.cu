/*
 * tmjx_step.cu — sm_100a kernels + C ABI (include/tmjx.h) for the batched rodent-tracking env step.
 *
 * One WARP per environment, all per-env state staged in shared memory / registers, one launch per control
 * step: `physics_steps_per_control_step` x (forward dynamics + constraint solve + Euler) followed by the
 * tracking reward / termination / observation epilogue.  Replaces, for this path,
 *   SingleClipTracking.step / reset_from_clip   reference track_mjx/environment/task/single_clip_tracking.py:121-320
 *   compute_tracking_rewards                     reference track_mjx/environment/task/reward.py:359-485
 *   BaseWalker.compute_local_*                   reference track_mjx/environment/walker/base.py:170-258
 *   brax PipelineEnv.pipeline_step -> mjx.step   (upstream mujoco-mjx 3.3.2; called at single_clip_tracking.py:219)
 *
 * B200 design notes (DESIGN.md has the full story):
 *   - the path is FP32-latency bound, not HBM bound (15 KB of HBM traffic vs ~4 MFLOP per env step), so the
 *     design goal is resident environments per SM: the per-env shared-memory slice is ~17 KB by aliasing the
 *     smooth-dynamics transients with the two sparse LDL factors, and nv- / constraint-row vectors live in
 *     registers (lane l owns dofs l, l+32, l+64 and the 4 pyramid rows of contact l + 3 joint-limit rows);
 *   - structure-exploiting instead of MJX's dense algebra: sparse L^T D L of the joint-space inertia with the
 *     damping-augmented factor (Euler) built in the same sweep, scatter-form triangular solves without
 *     shuffles, and a Jacobian-free constraint operator (contact-point velocities from per-body spatial
 *     velocity sums; forces mapped back as wrenches about the subtree COM);
 *   - tree passes are level-parallel with deterministic child gathers (no atomics => bitwise reproducible,
 *     independent of grid size, so sharding over GPUs cannot change per-env results).
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "tmjx_tables.h"

namespace tmjx {

struct KArgs {
  DevModel m;
  const DevTask* task;
  const float* clips;
  int n_clips, clip_len;
  TmjxState st;
  TmjxOut out;
  const float* action;
  int n_env;
  unsigned flags;
  float* spill;   // per resident warp: m.spill_floats floats (compact CG layout), else unused
  int phase_offset_ns;   // two-blocks-per-SM variant: the upper half of the grid starts this much later (complementary phases)
};

#ifdef TMJX_VARIANT  // device code: compiled once per residency variant, in parallel (see __graft_entry__.build)
namespace {      // internal linkage: every variant translation unit carries its own copy of the device functions
#define FULLMASK 0xffffffffu
constexpr float kMinVal = 1e-15f;

// Phase timers (development builds only, -DTMJX_PHASE_TIMING; tools/gpu_phase_timing.py): warp 0 of block 0 accumulates the
// clock64() cycles it spends in each phase of a substep -- the critical path of ONE environment, which is what bounds the
// lock-step round time.
#ifdef TMJX_PHASE_TIMING
__device__ unsigned long long g_pt[64];
struct PhaseTimer {
  long long t; bool on;
  __device__ __forceinline__ PhaseTimer() : t(0), on(blockIdx.x == 0 && threadIdx.x == 0) { if (on) t = clock64(); }
  __device__ __forceinline__ void lap(int i) { if (on) { const long long n = clock64(); g_pt[i] += (unsigned long long)(n - t); t = n; } }
};
#define PT_DECL PhaseTimer pt_
#define PT_LAP(i) pt_.lap(i)
#else
#define PT_DECL
#define PT_LAP(i)
#endif

// ---------------------------------------------------------------------------------------------- device math
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
// Sum 9 values over the warp, all lanes get all totals, with 22 shuffles instead of 45: eight of the values are
// reduced by recursive halving (offsets 16, 8, 4: a lane keeps one half of its list and sends the other, so after three
// exchanges it owns ONE partial value), two butterflies finish that value, eight immediate-lane shuffles gather the
// totals; the ninth value takes the plain butterfly.  The shared-memory / shuffle pipe is the scarce unit of this kernel.
__device__ __forceinline__ void wsum9(float (&v)[9], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float y[4], z[2], w;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b4 ? v[2 * i] : v[2 * i + 1], keep = b4 ? v[2 * i + 1] : v[2 * i];
    y[i] = keep + __shfl_xor_sync(FULLMASK, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b3 ? y[2 * i] : y[2 * i + 1], keep = b3 ? y[2 * i + 1] : y[2 * i];
    z[i] = keep + __shfl_xor_sync(FULLMASK, send, 8);
  }
  {
    const float send = b2 ? z[0] : z[1], keep = b2 ? z[1] : z[0];
    w = keep + __shfl_xor_sync(FULLMASK, send, 4);
  }
  w += __shfl_xor_sync(FULLMASK, w, 2);
  w += __shfl_xor_sync(FULLMASK, w, 1);
  float t = v[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(FULLMASK, t, o);
  v[8] = t;
  // value i ended up in the lanes with 4*b2 + 2*b3 + b4 == i
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(FULLMASK, w, ((i >> 2) & 1) * 4 + ((i >> 1) & 1) * 8 + (i & 1) * 16);
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
  const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void rot(const float* v, const float* q, float* o) {  // mjx math.rotate
  const float s = q[0];
  const float* u = q + 1;
  float c[3];
  cross3(u, v, c);
  const float uv = dot3(u, v), uu = dot3(u, u), k = s * s - uu;
  const float r0 = 2.f * (uv * u[0]) + k * v[0] + 2.f * s * c[0];
  const float r1 = 2.f * (uv * u[1]) + k * v[1] + 2.f * s * c[1];
  const float r2 = 2.f * (uv * u[2]) + k * v[2] + 2.f * s * c[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}
__device__ __forceinline__ void qmul(const float* a, const float* b, float* o) {
  const float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  const float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  const float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  const float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
__device__ __forceinline__ void q2mat(const float* q, float* m) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
  m[3] = 2.f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.f * (y * z - w * x);
  m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
__device__ __forceinline__ float normalize3(float* x) {  // math.normalize_with_norm
  const float n = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const float d = n + (n == 0.f ? 1e-6f : 0.f);
  x[0] = x[0] / d; x[1] = x[1] / d; x[2] = x[2] / d;
  return n;
}
__device__ __forceinline__ void normalize4(float* x) {
  const float n = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
  const float d = n + (n == 0.f ? 1e-6f : 0.f);
  x[0] = x[0] / d; x[1] = x[1] / d; x[2] = x[2] / d; x[3] = x[3] / d;
}
__device__ __forceinline__ void inert_mul(const float* i, const float* v, float* o) {  // math.inert_mul
  float c1[3], c2[3];
  cross3(i + 6, v + 3, c1);
  cross3(i + 6, v, c2);
  const float r0 = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c1[0];
  const float r1 = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c1[1];
  const float r2 = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c1[2];
  const float r3 = i[9] * v[3] - c2[0], r4 = i[9] * v[4] - c2[1], r5 = i[9] * v[5] - c2[2];
  o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4; o[5] = r5;
}
__device__ __forceinline__ void motion_cross(const float* u, const float* v, float* o) {
  float a[3], b[3], c[3];
  cross3(u, v, a);
  cross3(u + 3, v, b);
  cross3(u, v + 3, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
__device__ __forceinline__ void motion_cross_force(const float* v, const float* f, float* o) {
  float a[3], b[3], c[3];
  cross3(v, f, a);
  cross3(v + 3, f + 3, b);
  cross3(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
// constraint._kbi impedance for |pos| given the pre-digested solimp (par = dmin dmax width mid power)
__device__ __forceinline__ float impedance(const float* par, float pos) {
  const float dmin = par[0], dmax = par[1], width = par[2], mid = par[3], power = par[4];
  const float x = fabsf(pos) / width;
  float a, b;
  if (power == 2.f) {
    a = (1.f / mid) * (x * x);
    b = 1.f - (1.f / (1.f - mid)) * ((1.f - x) * (1.f - x));
  } else {
    a = (1.f / powf(mid, power - 1.f)) * powf(x, power);
    b = 1.f - (1.f / powf(1.f - mid, power - 1.f)) * powf(1.f - x, power);
  }
  const float y = x < mid ? a : b;
  float imp = dmin + y * (dmax - dmin);
  imp = fminf(fmaxf(imp, dmin), dmax);
  if (x > 1.f) imp = dmax;
  return imp;
}

// Block lock-step.  Every warp of a block simulates its own environment, but all of them walk the (large, ~400 KB)
// step code phase by phase TOGETHER: a block barrier between phases keeps the warps inside the same few KB of code, so
// the 32 KB L1.5 instruction cache and the model-table lines in L1 are fetched once per block instead of once per
// warp (ncu: `stall_no_inst` was the top stall reason with free-running warps).  Barriers sit only in block-uniform
// control flow; environment-dependent loops (CG termination, line search) are masked, never broken out of.
// DevModel.sync_level selects how many of the barrier sites are live (0: the one at the top of every substep, which
// measured fastest; 1: major phases; 2: all).
__device__ __forceinline__ void phase_sync() { __syncthreads(); }

// ---------------------------------------------------------------------------------------------- per-warp context
struct Warp {
  const DevModel& m;
  float* s;  // this environment's shared-memory slice
  int lane;
  int dep[kNvSlots], rend[kNvSlots];  // depth / row end (sparse L) of the dofs this lane owns
  float* spill;                       // this warp's global parking slot for the head of the Euler factor (compact layout)
  __device__ __forceinline__ float* at(int off) const { return s + off; }
};

// registers owned by a lane across the solver
struct Rows {
  float D[kRowSlots], aref[kRowSlots];
  float mu;          // friction of contact `lane`
  bool cact;         // contact `lane` penetrating (J rows non-zero)
  float lsign[kLimSlots];  // +-1 / 0 sign of the limit Jacobian entry
};

// ---- nv-vector helpers: lane l owns dofs l + 32 q
__device__ __forceinline__ void vput(const Warp& w, float* dst, const float v[kNvSlots]) {
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) { const int d = w.lane + 32 * q; if (d < w.m.nv) dst[d] = v[q]; }
}
__device__ __forceinline__ void vget(const Warp& w, const float* src, float v[kNvSlots]) {
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) { const int d = w.lane + 32 * q; v[q] = d < w.m.nv ? src[d] : 0.f; }
}
__device__ __forceinline__ float vdot(const float a[kNvSlots], const float b[kNvSlots]) {
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) s += a[q] * b[q];
  return wsum(s);
}


// 6- and 10-float records (spatial vectors, cdof rows, composite inertias) are 8-byte aligned in the slice: move them as
// float2 (LDS.64 / STS.64 halve the shared-memory instruction count of the tree passes)
template <int NC>
__device__ __forceinline__ void ld_rec(const float* p, float (&v)[NC]) {
  static_assert(NC % 2 == 0, "even record length");
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) { const float2 t = reinterpret_cast<const float2*>(p)[k]; v[2 * k] = t.x; v[2 * k + 1] = t.y; }
}
template <int NC>
__device__ __forceinline__ void st_rec(float* p, const float (&v)[NC]) {
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) reinterpret_cast<float2*>(p)[k] = make_float2(v[2 * k], v[2 * k + 1]);
}

// ---------------------------------------------------------------------------------------------- smooth dynamics
// ---- log-depth tree scans (ancestor doubling).  Lane l owns bodies l, l+32, l+64 (kBodySlots).
constexpr int kBodySlots = 3;

// inclusive prefix sum over the ancestors of every body: buf[b] <- sum of buf[a] for a on the path root..b (NC floats per body)
template <int NC>
__device__ __forceinline__ void scan_ancestors(const Warp& w, float* buf) {
  const DevModel& m = w.m;
  float acc[kBodySlots][NC];
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[s][k] = b < m.nbody ? buf[b * NC + k] : 0.f;
  }
  for (int r = 0; r < m.nround; ++r) {
    float add[kBodySlots][NC];
    bool on[kBodySlots];
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      const int b = w.lane + 32 * s;
      const int a = b < m.nbody ? int(m.anc_pow[r * m.nbody + b]) : 0;
      on[s] = a != 0;
#pragma unroll
      for (int k = 0; k < NC; k += 2) {   // float2: half the LDS, no 2-way bank conflict of the stride-NC scalar form
        float2 t2 = make_float2(0.f, 0.f);
        if (on[s]) t2 = *reinterpret_cast<const float2*>(buf + a * NC + k);
        add[s][k] = t2.x; add[s][k + 1] = t2.y;
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      if (!on[s]) continue;
      const int b = w.lane + 32 * s;
#pragma unroll
      for (int k = 0; k < NC; k += 2) {
        acc[s][k] += add[s][k]; acc[s][k + 1] += add[s][k + 1];
        *reinterpret_cast<float2*>(buf + b * NC + k) = make_float2(acc[s][k], acc[s][k + 1]);
      }
    }
    __syncwarp();
  }
}

// subtree sums: buf[b] <- sum of buf[c] over the subtree rooted at b (b >= 1), by doubling over descendant distance.
// The per-(round, body) descriptors are fetched up front (independent loads) so that the rounds only touch shared memory.
constexpr int kMaxRounds = 6;  // trees up to 64 levels deep
template <int NC>
__device__ __forceinline__ void sum_subtrees_list(const Warp& w, float* buf);

// Default form: the round loop is NOT unrolled (the code of one round stays in the instruction cache for the other five;
// the unrolled form spent a third of its samples waiting for instruction fetch), the descriptor of the next round is
// fetched while the current one runs, and a body's (<= 4) descendants at distance 2^r come packed in one word: the
// gather is a warp-uniform loop of dsc_maxc[r][slot] predicated float2 accumulations instead of a divergent list walk.
template <int NC>
__device__ __forceinline__ void sum_subtrees(const Warp& w, float* buf) {
  const DevModel& m = w.m;
  if (!m.use_dsc4) { sum_subtrees_list<NC>(w, buf); return; }
  float acc[kBodySlots][NC];
  uint32_t pk[kBodySlots];
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
    pk[s] = (b < m.nbody && m.nround > 0) ? m.dsc4[b] : 0u;
    if (b < m.nbody) ld_rec<NC>(buf + b * NC, acc[s]);
    else {
#pragma unroll
      for (int k = 0; k < NC; ++k) acc[s][k] = 0.f;
    }
  }
#pragma unroll 1
  for (int r = 0; r < m.nround; ++r) {
    uint32_t nx[kBodySlots];
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      const int b = w.lane + 32 * s;
      nx[s] = (b < m.nbody && r + 1 < m.nround) ? m.dsc4[(r + 1) * m.nbody + b] : 0u;
    }
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      const int maxc = m.dsc_maxc[r][s];
      uint32_t p = pk[s];
      for (int i = 0; i < maxc; ++i, p >>= 8) {
        const int id = int(p & 0xffu);
        if (id) {
          const float2* src = reinterpret_cast<const float2*>(buf + id * NC);
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) { const float2 t2 = src[k]; acc[s][2 * k] += t2.x; acc[s][2 * k + 1] += t2.y; }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      if (!pk[s]) continue;
      const int b = w.lane + 32 * s;
#pragma unroll
      for (int k = 0; k < NC; k += 2) *reinterpret_cast<float2*>(buf + b * NC + k) = make_float2(acc[s][k], acc[s][k + 1]);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) pk[s] = nx[s];
  }
}

// list form (any number of descendants per distance): the fallback for trees the packed form cannot describe
template <int NC>
__device__ __forceinline__ void sum_subtrees_list(const Warp& w, float* buf) {
  const DevModel& m = w.m;
  float acc[kBodySlots][NC];
  uint32_t pack[kBodySlots][kMaxRounds];
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
#pragma unroll
    for (int r = 0; r < kMaxRounds; ++r) pack[s][r] = (b < m.nbody && r < m.nround) ? m.dsc_pack[r * m.nbody + b] : 0u;
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[s][k] = b < m.nbody ? buf[b * NC + k] : 0.f;
  }
#pragma unroll
  for (int r = 0; r < kMaxRounds; ++r) {
    if (r >= m.nround) break;
    bool on[kBodySlots];
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      const uint32_t p = pack[s][r];
      const int cnt = int(p >> 16);
      on[s] = cnt != 0;
      if (cnt == 1) {
        const float2* src = reinterpret_cast<const float2*>(buf + int(p & 0xffffu) * NC);
#pragma unroll
        for (int k = 0; k < NC / 2; ++k) { const float2 t2 = src[k]; acc[s][2 * k] += t2.x; acc[s][2 * k + 1] += t2.y; }
      } else if (cnt > 1) {
        const int e0 = int(p & 0xffffu);
        for (int e = e0; e < e0 + cnt; ++e) {
          const float2* src = reinterpret_cast<const float2*>(buf + int(m.dsc_list[e]) * NC);
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) { const float2 t2 = src[k]; acc[s][2 * k] += t2.x; acc[s][2 * k + 1] += t2.y; }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      if (!on[s]) continue;
      const int b = w.lane + 32 * s;
#pragma unroll
      for (int k = 0; k < NC; k += 2) *reinterpret_cast<float2*>(buf + b * NC + k) = make_float2(acc[s][k], acc[s][k + 1]);
    }
    __syncwarp();
  }
}

// smooth.kinematics.  Every body first composes its pose RELATIVE TO ITS PARENT (body offset, then its joints in order,
// exactly the per-body arithmetic of mjx), recording joint anchors / axes in the parent frame; world poses then follow
// from ancestor doubling (log2(depth) rounds of "compose with the ancestor 2^r levels up" instead of a 40-level walk),
// and anchors / axes are mapped to the world with the parent's final pose.
__device__ void kinematics(const Warp& w) {
  const DevModel& m = w.m;
  float* qpos = w.at(m.o_qpos);
  float* xpos = w.at(m.o_xpos);
  float* xquat = w.at(m.o_xquat);
  float* xipos = w.at(m.o_big + m.a_xipos);
  float* anchor = w.at(m.o_big + m.a_anchor);
  float* axis = w.at(m.o_big + m.a_axis);
  float P[kBodySlots][3], Q[kBodySlots][4];
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
    float pos[3] = {0.f, 0.f, 0.f}, quat[4] = {1.f, 0.f, 0.f, 0.f};
    if (b > 0 && b < m.nbody) {
      pos[0] = m.body_pos[b * 3]; pos[1] = m.body_pos[b * 3 + 1]; pos[2] = m.body_pos[b * 3 + 2];
      quat[0] = m.body_quat[b * 4]; quat[1] = m.body_quat[b * 4 + 1]; quat[2] = m.body_quat[b * 4 + 2]; quat[3] = m.body_quat[b * 4 + 3];
      const int ja = m.body_jntadr[b], jn = m.body_jntnum[b];
      for (int jj = 0; jj < jn; ++jj) {
        const int j = ja + jj, qa = m.jnt_qposadr[j];
        float t[3];
        if (m.jnt_type[j] == kJntFree) {  // the tree root: its pose is absolute (parent = world)
          pos[0] = qpos[qa]; pos[1] = qpos[qa + 1]; pos[2] = qpos[qa + 2];
          quat[0] = qpos[qa + 3]; quat[1] = qpos[qa + 4]; quat[2] = qpos[qa + 5]; quat[3] = qpos[qa + 6];
          normalize4(quat);
          qpos[qa + 3] = quat[0]; qpos[qa + 4] = quat[1]; qpos[qa + 5] = quat[2]; qpos[qa + 6] = quat[3];
          anchor[j * 3] = pos[0]; anchor[j * 3 + 1] = pos[1]; anchor[j * 3 + 2] = pos[2];
          axis[j * 3] = 0.f; axis[j * 3 + 1] = 0.f; axis[j * 3 + 2] = 1.f;
        } else {
          const float jp[3] = {m.jnt_pos[j * 3], m.jnt_pos[j * 3 + 1], m.jnt_pos[j * 3 + 2]};
          const float ja3[3] = {m.jnt_axis[j * 3], m.jnt_axis[j * 3 + 1], m.jnt_axis[j * 3 + 2]};
          float an[3];
          rot(jp, quat, t);
          an[0] = t[0] + pos[0]; an[1] = t[1] + pos[1]; an[2] = t[2] + pos[2];
          anchor[j * 3] = an[0]; anchor[j * 3 + 1] = an[1]; anchor[j * 3 + 2] = an[2];
          rot(ja3, quat, t);
          axis[j * 3] = t[0]; axis[j * 3 + 1] = t[1]; axis[j * 3 + 2] = t[2];
          float sn, cs;
          sincosf((qpos[qa] - m.jnt_qpos0[j]) * 0.5f, &sn, &cs);
          const float ql[4] = {cs, ja3[0] * sn, ja3[1] * sn, ja3[2] * sn};
          float q2[4];
          qmul(quat, ql, q2);
          quat[0] = q2[0]; quat[1] = q2[1]; quat[2] = q2[2]; quat[3] = q2[3];
          rot(jp, quat, t);
          pos[0] = an[0] - t[0]; pos[1] = an[1] - t[1]; pos[2] = an[2] - t[2];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) P[s][k] = pos[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) Q[s][k] = quat[k];
    if (b < m.nbody) {
      xpos[b * 3] = pos[0]; xpos[b * 3 + 1] = pos[1]; xpos[b * 3 + 2] = pos[2];
      *reinterpret_cast<float4*>(xquat + b * 4) = make_float4(quat[0], quat[1], quat[2], quat[3]);
    }
  }
  __syncwarp();
  for (int r = 0; r < m.nround; ++r) {
    float Pa[kBodySlots][3], Qa[kBodySlots][4];
    bool on[kBodySlots];
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      const int b = w.lane + 32 * s;
      const int a = b < m.nbody ? int(m.anc_pow[r * m.nbody + b]) : 0;
      on[s] = a != 0;
      if (on[s]) {
#pragma unroll
        for (int k = 0; k < 3; ++k) Pa[s][k] = xpos[a * 3 + k];
        { const float4 q4 = *reinterpret_cast<const float4*>(xquat + a * 4); Qa[s][0] = q4.x; Qa[s][1] = q4.y; Qa[s][2] = q4.z; Qa[s][3] = q4.w; }
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kBodySlots; ++s) {
      if (!on[s]) continue;
      const int b = w.lane + 32 * s;
      float t[3], q2[4];
      rot(P[s], Qa[s], t);
      qmul(Qa[s], Q[s], q2);
#pragma unroll
      for (int k = 0; k < 3; ++k) { P[s][k] = Pa[s][k] + t[k]; xpos[b * 3 + k] = P[s][k]; }
#pragma unroll
      for (int k = 0; k < 4; ++k) Q[s][k] = q2[k];
      *reinterpret_cast<float4*>(xquat + b * 4) = make_float4(q2[0], q2[1], q2[2], q2[3]);
    }
    __syncwarp();
  }
  // joint anchors / axes to the world frame with the parent body's pose; inertial frame origins
  for (int j = w.lane; j < m.njnt; j += 32) {
    if (m.jnt_type[j] == kJntFree) continue;
    const int p = m.body_parent[m.jnt_body[j]];
    float t[3];
    rot(anchor + j * 3, xquat + p * 4, t);
    anchor[j * 3] = xpos[p * 3] + t[0]; anchor[j * 3 + 1] = xpos[p * 3 + 1] + t[1]; anchor[j * 3 + 2] = xpos[p * 3 + 2] + t[2];
    rot(axis + j * 3, xquat + p * 4, t);
    axis[j * 3] = t[0]; axis[j * 3 + 1] = t[1]; axis[j * 3 + 2] = t[2];
  }
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
    if (b < m.nbody) {
      float t[3];
      rot(m.body_ipos + b * 3, Q[s], t);
      xipos[b * 3] = P[s][0] + t[0]; xipos[b * 3 + 1] = P[s][1] + t[1]; xipos[b * 3 + 2] = P[s][2] + t[2];
    }
  }
  __syncwarp();
}

// smooth.com_pos: COM of the moving tree (warp reduction), cinert per body, cdof per dof
__device__ void com_pos(const Warp& w, float com[3]) {
  const DevModel& m = w.m;
  const float* xquat = w.at(m.o_xquat);
  const float* xipos = w.at(m.o_big + m.a_xipos);
  const float* anchor = w.at(m.o_big + m.a_anchor);
  const float* axis = w.at(m.o_big + m.a_axis);
  const float* qpos = w.at(m.o_qpos);
  float* cin = w.at(m.o_cin);
  float* cdof = w.at(m.o_cdof);
  float px = 0.f, py = 0.f, pz = 0.f;
  for (int b = w.lane; b < m.nbody; b += 32) {
    const float mt = m.body_tree_mass[b];
    px += xipos[b * 3] * mt; py += xipos[b * 3 + 1] * mt; pz += xipos[b * 3 + 2] * mt;
  }
  com[0] = wsum(px) / m.tree_mass; com[1] = wsum(py) / m.tree_mass; com[2] = wsum(pz) / m.tree_mass;
  for (int b = w.lane; b < m.nbody; b += 32) {
    float q[4], R[9];
    qmul(xquat + b * 4, m.body_iquat + b * 4, q);
    q2mat(q, R);
    const float I0 = m.body_inertia[b * 3], I1 = m.body_inertia[b * 3 + 1], I2 = m.body_inertia[b * 3 + 2], ms = m.body_mass[b];
    const float off[3] = {xipos[b * 3] - com[0], xipos[b * 3 + 1] - com[1], xipos[b * 3 + 2] - com[2]};
    const float oo = dot3(off, off);
    float ci[10];
#define INRC(r, c) (R[r * 3] * I0 * R[c * 3] + R[r * 3 + 1] * I1 * R[c * 3 + 1] + R[r * 3 + 2] * I2 * R[c * 3 + 2])
    ci[0] = INRC(0, 0) + (oo - off[0] * off[0]) * ms;
    ci[1] = INRC(1, 1) + (oo - off[1] * off[1]) * ms;
    ci[2] = INRC(2, 2) + (oo - off[2] * off[2]) * ms;
    ci[3] = INRC(0, 1) + (0.f - off[0] * off[1]) * ms;
    ci[4] = INRC(0, 2) + (0.f - off[0] * off[2]) * ms;
    ci[5] = INRC(1, 2) + (0.f - off[1] * off[2]) * ms;
#undef INRC
    ci[6] = off[0] * ms; ci[7] = off[1] * ms; ci[8] = off[2] * ms; ci[9] = ms;
    st_rec<10>(cin + b * 10, ci);
  }
  for (int d = w.lane; d < m.nv; d += 32) {
    const int j = m.dof_jnt[d];
    float cd[6];
    const float off[3] = {com[0] - anchor[j * 3], com[1] - anchor[j * 3 + 1], com[2] - anchor[j * 3 + 2]};
    if (m.jnt_type[j] == kJntFree) {
      const int r = d - m.jnt_dofadr[j];
      if (r < 3) {
        cd[0] = cd[1] = cd[2] = 0.f;
        cd[3] = r == 0 ? 1.f : 0.f; cd[4] = r == 1 ? 1.f : 0.f; cd[5] = r == 2 ? 1.f : 0.f;
      } else {
        float R[9];
        q2mat(xquat + m.jnt_body[j] * 4, R);
        const int c = r - 3;
        const float a[3] = {R[c], R[3 + c], R[6 + c]};
        cd[0] = a[0]; cd[1] = a[1]; cd[2] = a[2];
        cross3(a, off, cd + 3);
      }
    } else {
      const float a[3] = {axis[j * 3], axis[j * 3 + 1], axis[j * 3 + 2]};
      cd[0] = a[0]; cd[1] = a[1]; cd[2] = a[2];
      cross3(a, off, cd + 3);
    }
    st_rec<6>(cdof + d * 6, cd);
  }
  (void)qpos;
  __syncwarp();
}

// smooth.com_vel + smooth.rne without level walks: body velocities / accelerations are prefix sums of per-body
// increments over the ancestor path (two doubling scans), cdof_dot is evaluated per dof from its parent body's velocity
// plus the earlier joints of the same body (mjx's in-body order), and the backward force accumulation is a subtree sum.
// Returns qfrc_bias per dof in registers.
__device__ void com_vel_rne(const Warp& w, float bias[kNvSlots]) {
  const DevModel& m = w.m;
  const float* qvel = w.at(m.o_qvel);
  const float* cdof = w.at(m.o_cdof);
  const float* cin = w.at(m.o_cin);
  float* cvel = w.at(m.o_big + m.a_cvel);
  float* cdd = w.at(m.o_big + m.a_cdofdot);
  float* cacc = w.at(m.o_big + m.a_cacc);
  // per-body velocity increment: sum of cdof * qvel over the body's own dofs
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
    if (b < m.nbody) {
      float dv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int da = m.body_dofadr[b], dn = m.body_dofnum[b];
      for (int d = da; d < da + dn; ++d) {
        const float v = qvel[d];
        float cd[6];
        ld_rec<6>(cdof + d * 6, cd);
#pragma unroll
        for (int k = 0; k < 6; ++k) dv[k] += cd[k] * v;
      }
      st_rec<6>(cvel + b * 6, dv);
    }
  }
  __syncwarp();
  scan_ancestors<6>(w, cvel);
  if (m.sync_level > 1) phase_sync();
  // cdof_dot per dof
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) {
    const int d = w.lane + 32 * q;
    if (d < m.nv) {
      const int b = m.dof_body[d], p = m.body_parent[b], j = m.dof_jnt[d], d0 = m.body_dofadr[b];
      float cv[6], t[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      ld_rec<6>(cvel + p * 6, cv);
      const bool is_free = m.jnt_type[j] == kJntFree;
      const int r = d - m.jnt_dofadr[j];
      // free joint: translations have cdof_dot = 0, the three rotations all see the velocity after the translations
      const int upto = is_free ? (r < 3 ? d0 : d0 + 3) : d;
      for (int e = d0; e < upto; ++e) {
        const float v = qvel[e];
        float ce[6];
        ld_rec<6>(cdof + e * 6, ce);
#pragma unroll
        for (int k = 0; k < 6; ++k) cv[k] += ce[k] * v;
      }
      if (!(is_free && r < 3)) { float cd[6]; ld_rec<6>(cdof + d * 6, cd); motion_cross(cv, cd, t); }
      st_rec<6>(cdd + d * 6, t);
    }
  }
  __syncwarp();
  // per-body acceleration increment, prefix sum, plus the root acceleration (-gravity) every body inherits
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
    if (b < m.nbody) {
      float da6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int da = m.body_dofadr[b], dn = m.body_dofnum[b];
      for (int d = da; d < da + dn; ++d) {
        const float v = qvel[d];
        float cd[6];
        ld_rec<6>(cdd + d * 6, cd);
#pragma unroll
        for (int k = 0; k < 6; ++k) da6[k] += cd[k] * v;
      }
      st_rec<6>(cacc + b * 6, da6);
    }
  }
  __syncwarp();
  scan_ancestors<6>(w, cacc);
  if (m.sync_level > 1) phase_sync();
  // local cfrc (in place of cacc)
#pragma unroll
  for (int s = 0; s < kBodySlots; ++s) {
    const int b = w.lane + 32 * s;
    if (b < m.nbody) {
      float ca[6], cv[6], ci[10], f1[6], f2[6], f3[6];
      ld_rec<6>(cacc + b * 6, ca);
      ld_rec<6>(cvel + b * 6, cv);
      ld_rec<10>(cin + b * 10, ci);
#pragma unroll
      for (int k = 3; k < 6; ++k) ca[k] = ca[k] + -m.gravity[k - 3];
      inert_mul(ci, ca, f1);
      inert_mul(ci, cv, f2);
      motion_cross_force(cv, f2, f3);
#pragma unroll
      for (int k = 0; k < 6; ++k) f1[k] = f1[k] + f3[k];
      st_rec<6>(cacc + b * 6, f1);
    }
  }
  __syncwarp();
  sum_subtrees<6>(w, cacc);
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) {
    const int d = w.lane + 32 * q;
    float acc = 0.f;
    if (d < m.nv) {
      float cf[6], cd[6];
      ld_rec<6>(cacc + m.dof_body[d] * 6, cf);
      ld_rec<6>(cdof + d * 6, cd);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc += cd[k] * cf[k];
    }
    bias[q] = acc;
  }
}

// passive.passive + forward.fwd_actuation; returns qfrc_actuator and qfrc_smooth in registers
__device__ void passive_actuation(const Warp& w, const float bias[kNvSlots], float qfa[kNvSlots], float qfs[kNvSlots],
                                  float actdot[2]) {
  const DevModel& m = w.m;
  const float* qpos = w.at(m.o_qpos);
  const float* qvel = w.at(m.o_qvel);
  const float* act = w.at(m.o_act);
  const float* ctrl = w.at(m.o_ctrl);
  float* force = w.at(m.o_big + m.a_force);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int u = w.lane + 32 * k;
    actdot[k] = 0.f;
    if (u < m.nu) {
      const int fl = m.act_flags[u];
      const float c = ctrl[u];
      float ca = c;
      if (fl & 2) { actdot[k] = (c - act[u]) / m.act_dyn_inv[u]; ca = act[u]; }
      float f = m.act_gain[u] * ca;
      if (fl & 4) {
        float len = 0.f, vel = 0.f;
        for (int e = m.act_mom_start[u]; e < m.act_mom_start[u + 1]; ++e) {
          const int d = m.act_mom_dof[e];
          len += m.act_mom_coef[e] * qpos[d + 1];
          vel += m.act_mom_coef[e] * qvel[d];
        }
        f += m.act_bias[u * 3] + m.act_bias[u * 3 + 1] * len + m.act_bias[u * 3 + 2] * vel;
      }
      if (fl & 8) f = fminf(fmaxf(f, m.act_force_lo[u]), m.act_force_hi[u]);
      force[u] = f;
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) {
    const int d = w.lane + 32 * q;
    float a = 0.f, pas = 0.f;
    if (d < m.nv) {
      for (int e = m.dof_act_start[d]; e < m.dof_act_start[d + 1]; ++e) a += m.dof_act_coef[e] * force[m.dof_act_id[e]];
      pas = -m.dof_damping[d] * qvel[d];
      const int qa = m.dof_qadr[d];
      if (qa >= 0) { const int j = m.dof_jnt[d]; pas += -m.jnt_stiffness[j] * (qpos[qa] - m.jnt_springref[j]); }
    }
    qfa[q] = a;
    qfs[q] = pas - bias[q] + a;
  }
}

// smooth.crb + support.make_m into the sparse rows of L1 (raw M), plus the damping-augmented copy in L2
__device__ void build_m(const Warp& w) {
  const DevModel& m = w.m;
  float* crb = w.at(m.o_cin);
  const float* cdof = w.at(m.o_cdof);
  float* L1 = w.at(m.o_big);
  float* L2 = w.at(m.o_L2);
  float* f = w.at(m.o_big + m.a_f);  // M-build scratch, dead before L2 is written
  sum_subtrees<10>(w, crb);
  for (int d = w.lane; d < m.nv; d += 32) {
    float t[6], ci[10], cd[6];
    ld_rec<10>(crb + m.dof_body[d] * 10, ci);
    ld_rec<6>(cdof + d * 6, cd);
    inert_mul(ci, cd, t);
    st_rec<6>(f + d * 6, t);
  }
  __syncwarp();
  __syncwarp();  // f (aliasing the o_L2 block) is complete
  // both matrices (and Newton's copy) from one pass; the entry values are kept in registers across the barrier that
  // retires the scratch f
  {
    constexpr int kMaxIt = 40;  // nM <= 1280
    float v[kMaxIt];
#pragma unroll
    for (int it = 0; it < kMaxIt; ++it) {
      const int e = w.lane + 32 * it;
      v[it] = 0.f;
      if (e < m.nM) {
        const uint32_t rc = (m.m_rc2[w.lane + 32 * (it >> 1)] >> (16 * (it & 1))) & 0xffffu;
        const int i = rc & 0xff, j = rc >> 8;
        float a = 0.f, fi[6], cj[6];
        ld_rec<6>(f + i * 6, fi);
        ld_rec<6>(cdof + j * 6, cj);
#pragma unroll
        for (int k = 0; k < 6; ++k) a += fi[k] * cj[k];
        v[it] = a;
      }
    }
    __syncwarp();  // f is dead from here
#pragma unroll
    for (int it = 0; it < kMaxIt; ++it) {
      const int e = w.lane + 32 * it;
      if (e < m.nM) {
        L1[e] = v[it];
        L2[e] = v[it];
        if (m.o_L != m.o_big) w.at(m.o_L)[e] = v[it];  // Newton: the raw inertia stays at o_big, its copy is factored
      }
    }
  }
  __syncwarp();
  // diagonal: + armature (all copies), + dt * damping (the Euler matrix)
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) {
    const int d = w.lane + 32 * q;
    if (d < m.nv) {
      const int e = w.rend[q] - w.dep[q];
      const float vd = L1[e] + m.dof_armature[d];
      L1[e] = vd;
      L2[e] = vd + __fmul_rn(m.dt, m.dof_damping[d]);
      if (m.o_L != m.o_big) w.at(m.o_L)[e] = vd;   // Newton
    }
  }
  __syncwarp();
}

// out = M x with the RAW sparse symmetric M in L1 (called before the factorisation overwrites it); x in shared memory
__device__ void mul_m_raw(const Warp& w, const float* x, float out[kNvSlots]) {
  const DevModel& m = w.m;
  const float* M = w.at(m.o_big);
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) {
    const int d = w.lane + 32 * q;
    float acc = 0.f;
    if (d < m.nv) {
      const int adr = m.dof_madr[d], c = m.dof_depth[d];
      for (int a = 0; a <= c; ++a) acc += M[adr + a] * x[m.m_anc[adr + a]];
      for (int t = m.desc_start[d]; t < m.desc_start[d + 1]; ++t) acc += M[m.desc_off[t]] * x[m.desc_dof[t]];
    }
    out[q] = acc;
  }
}

// both L^T D L factorisations (smooth.factor_m of M, and of M + dt*diag(damping) for forward.euler) in one sweep.
// The two matrices have the same structure, so they are STACKED in one warp: half-warp h = lane >> 4 works on matrix h,
// lane i = lane & 15 of each half owns the entries whose COLUMN dof has depth 16 j + i (column block j), for every row.
// Row k is held in registers (one register per column block), the pivot entry l_a of an ancestor row travels by one
// half-warp shuffle, the rank-1 update of row anc(k, a) is one LDS / FFMA / STS for BOTH matrices and touches
// consecutive addresses (conflict-free).  An entry is only ever read and written by the one lane that owns its
// (matrix, column depth), so the factorisation needs no warp barrier; the division 1/D is shared by the two matrices.
// The hot a-loop (a < 16: one column block) is software-pipelined four rows at a time; rows of one pivot never alias.
__device__ void factor_dual(const Warp& w) {
  const DevModel& m = w.m;
  const int lane = w.lane, i = lane & 15, hbit = lane & 16;
  float* ps = w.at(m.o_L) + (hbit ? m.nMpad : 0) - i;  // entry (row r, column depth 16 j + i) of my matrix: ps[rowend(r) - 16 j]
  for (int k = m.nv - 1; k >= 0; --k) {
    if ((k & 7) == 7) phase_sync();
    const int c = m.u_depth[k], re = m.u_rowend[k];
    float* rk = ps + re;
    float lb0 = 0.f, lb1 = 0.f, lb2 = 0.f;
    if (i <= c) lb0 = rk[0];
    if (c >= 16 && i + 16 <= c) lb1 = rk[-16];
    if (c >= 32 && i + 32 <= c) lb2 = rk[-32];
    const float dsel = c < 16 ? lb0 : (c < 32 ? lb1 : lb2);
    const float d = __shfl_sync(FULLMASK, dsel, (c & 15) | hbit);
    const float inv = 1.f / d;
    const float w0 = lb0 * inv, w1 = lb1 * inv, w2 = lb2 * inv;
    // scaled row + inverted diagonal (row k is not touched by its own rank-1 updates)
    if (i < c) rk[0] = w0; else if (i == c) rk[0] = inv;
    if (c >= 16) { if (i + 16 < c) rk[-16] = w1; else if (i + 16 == c) rk[-16] = inv; }
    if (c >= 32) { if (i + 32 < c) rk[-32] = w2; else if (i + 32 == c) rk[-32] = inv; }
    int al = c - 1, ti = re - al;  // u_ancre[ti] = row end of the ancestor at depth al
    for (; al >= 32; --al, ++ti) {  // deep chain tips only: three column blocks
      float* t = ps + int(m.u_ancre[ti]);
      const float a = __shfl_sync(FULLMASK, lb2, (al & 15) | hbit);
      t[0] = fmaf(-a, w0, t[0]);
      t[-16] = fmaf(-a, w1, t[-16]);
      if (i + 32 <= al) t[-32] = fmaf(-a, w2, t[-32]);
    }
    for (; al >= 16; --al, ++ti) {  // two column blocks
      float* t = ps + int(m.u_ancre[ti]);
      const float a = __shfl_sync(FULLMASK, lb1, (al & 15) | hbit);
      const float v0 = t[0];
      t[0] = fmaf(-a, w0, v0);
      if (i + 16 <= al) t[-16] = fmaf(-a, w1, t[-16]);
    }
    for (; al >= 3; al -= 4, ti += 4) {  // one column block, four rows in flight
      float* t0 = ps + int(m.u_ancre[ti]); float* t1 = ps + int(m.u_ancre[ti + 1]); float* t2 = ps + int(m.u_ancre[ti + 2]);
      float* t3 = ps + int(m.u_ancre[ti + 3]);
      const float a0 = __shfl_sync(FULLMASK, lb0, al | hbit), a1 = __shfl_sync(FULLMASK, lb0, (al - 1) | hbit);
      const float a2 = __shfl_sync(FULLMASK, lb0, (al - 2) | hbit), a3 = __shfl_sync(FULLMASK, lb0, (al - 3) | hbit);
      const bool p0 = i <= al, p1 = i <= al - 1, p2 = i <= al - 2, p3 = i <= al - 3;
      float v0, v1, v2, v3;
      if (p0) v0 = t0[0];
      if (p1) v1 = t1[0];
      if (p2) v2 = t2[0];
      if (p3) v3 = t3[0];
      if (p0) t0[0] = fmaf(-a0, w0, v0);
      if (p1) t1[0] = fmaf(-a1, w0, v1);
      if (p2) t2[0] = fmaf(-a2, w0, v2);
      if (p3) t3[0] = fmaf(-a3, w0, v3);
    }
    for (; al >= 0; --al, ++ti) {
      float* t = ps + int(m.u_ancre[ti]);
      const float a = __shfl_sync(FULLMASK, lb0, al | hbit);
      if (i <= al) t[0] = fmaf(-a, w0, t[0]);
    }
  }
  __syncwarp();  // consumers (solves, M products) use a dof-lane mapping
}

// x <- (L^T D L)^-1 x with x in REGISTERS (lane l owns dofs l + 32 t): mj_solveLD order, the pivot value travels by
// warp shuffle, ancestor / descendant tests are bit masks, L is read from shared memory; no barriers
template <int S>
__device__ __forceinline__ void solve_up(const DevModel& m, const float* L, float (&x)[kNvSlots], const uint32_t (&dm)[kNvSlots][kNvSlots],
                                         const int (&dep)[kNvSlots]) {
  const int hi = min(31, m.nv - 1 - 32 * S);
  for (int li = hi; li >= 0; --li) {
    const float xi = __shfl_sync(FULLMASK, x[S], li);
    const int re = m.u_rowend[li + 32 * S];
#pragma unroll
    for (int t = 0; t <= S; ++t)
      if (m.slot_used[t][S] && ((dm[t][S] >> li) & 1u)) x[t] -= L[re - dep[t]] * xi;
  }
}
template <int S>
__device__ __forceinline__ void solve_down(const DevModel& m, const float* L, float (&x)[kNvSlots], const uint32_t (&am)[kNvSlots][kNvSlots],
                                           const int (&rend)[kNvSlots]) {
  const int hi = min(31, m.nv - 1 - 32 * S);
  for (int lj = 0; lj <= hi; ++lj) {
    const float xj = __shfl_sync(FULLMASK, x[S], lj);
    const int dj = m.u_depth[lj + 32 * S];
#pragma unroll
    for (int t = S; t < kNvSlots; ++t)
      if (m.slot_used[S][t] && ((am[t][S] >> lj) & 1u)) x[t] -= L[rend[t] - dj] * xj;
  }
}
__device__ __noinline__ void solve_reg(const DevModel& m, const float* L, float (&x)[kNvSlots], int lane) {
  uint32_t dm[kNvSlots][kNvSlots], am[kNvSlots][kNvSlots];
  int dep[kNvSlots], rend[kNvSlots];
#pragma unroll
  for (int t = 0; t < kNvSlots; ++t) {
    dep[t] = m.depth_me[t * 32 + lane];
    rend[t] = m.rowend_me[t * 32 + lane];
#pragma unroll
    for (int s2 = 0; s2 < kNvSlots; ++s2) { dm[t][s2] = m.dmask[(t * 3 + s2) * 32 + lane]; am[t][s2] = m.amask[(t * 3 + s2) * 32 + lane]; }
  }
  solve_up<2>(m, L, x, dm, dep);
  solve_up<1>(m, L, x, dm, dep);
  solve_up<0>(m, L, x, dm, dep);
#pragma unroll
  for (int t = 0; t < kNvSlots; ++t)
    if (lane + 32 * t < m.nv) x[t] *= L[rend[t] - dep[t]];
  solve_down<0>(m, L, x, am, rend);
  solve_down<1>(m, L, x, am, rend);
  solve_down<2>(m, L, x, am, rend);
}

// solve with the factor at `L`: straight-line generated code when the dof tree is the one it was generated for
__device__ __forceinline__ void solve_ld(const Warp& w, const float* L, float (&x)[kNvSlots]) {
  if (w.m.use_gen) {
    gen::V3 v;
    v.a = x[0]; v.b = x[1]; v.c = x[2];
    v = gen::solve(L, v, w.lane, w.dep[0], w.dep[1], w.dep[2], w.rend[0], w.rend[1], w.rend[2]);
    x[0] = v.a; x[1] = v.b; x[2] = v.c;
  } else {
    solve_reg(w.m, L, x, w.lane);
  }
}

// ---------------------------------------------------------------------------------------------- constraints
// jv = J x for the rows this lane owns; x in shared memory
__device__ void apply_J(const Warp& w, const Rows& r, const float* x, float jv[kRowSlots]) {
  const DevModel& m = w.m;
  const float* cdof = w.at(m.o_cdof);
  float* sV = w.at(m.o_cin + m.c_sV);
  if (m.use_seg) {
    // segment partial sums, then each contact body adds the (<= 4) segments of its root path
    float* sP = w.at(m.o_cin + m.c_sP);
    const uint32_t st0 = m.seg_task[w.lane], st1 = m.seg_task[w.lane + 32], ct0 = m.cb_task[w.lane], ct1 = m.cb_task[w.lane + 32];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const uint32_t st = sl ? st1 : st0;
      const int d0 = st & 0xff, len = (st >> 8) & 0xf, k = (st >> 12) & 7;
      const float* cd = cdof + d0 * 6 + k;
      const float* xs = x + d0;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < len) acc = fmaf(cd[i * 6], xs[i], acc);
      if (st & 0x8000u) sP[w.lane + 32 * sl] = acc;
    }
    __syncwarp();
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const uint32_t ct = sl ? ct1 : ct0;
      const int k = (ct >> 16) & 7;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int sg = (ct >> (4 * j)) & 0xf; if (sg != 0xf) acc += sP[sg * 6 + k]; }
      if (ct & 0x80000u) sV[w.lane + 32 * sl] = acc;
    }
  } else {
    for (int t = w.lane; t < m.ncb * 6; t += 32) {
      const int cb = t / 6, k = t % 6;
      float acc = 0.f;
      for (int e = m.cb_chain_start[cb]; e < m.cb_chain_start[cb + 1]; ++e) { const int d = m.cb_chain_dof[e]; acc += cdof[d * 6 + k] * x[d]; }
      sV[t] = acc;
    }
  }
  __syncwarp();
  jv[0] = jv[1] = jv[2] = jv[3] = 0.f;
  if (w.lane < m.ncon && r.cact) {
    float V[6];
    ld_rec<6>(sV + m.con_cb[w.lane] * 6, V);
    const float* off = w.at(m.o_cin + m.c_off) + w.lane * 3;
    const float* t1 = w.at(m.o_cin + m.c_t1) + w.lane * 3;
    float vel[3], t2[3];
    cross3(V, off, vel);
    vel[0] += V[3]; vel[1] += V[4]; vel[2] += V[5];
    cross3(m.plane_n, t1, t2);
    const float jn = dot3(m.plane_n, vel), j1 = dot3(t1, vel), j2 = dot3(t2, vel);
    jv[0] = jn + j1 * r.mu; jv[1] = jn - j1 * r.mu; jv[2] = jn + j2 * r.mu; jv[3] = jn - j2 * r.mu;
  }
#pragma unroll
  for (int q = 0; q < kLimSlots; ++q) {
    const int l = w.lane + 32 * q;
    jv[4 + q] = l < m.nlimit ? r.lsign[q] * x[m.lim_dof[l]] : 0.f;
  }
  __syncwarp();
}

// out = J^T f for the dofs this lane owns
__device__ void apply_JT(const Warp& w, const Rows& r, const float f[kRowSlots], float out[kNvSlots]) {
  const DevModel& m = w.m;
  const float* cdof = w.at(m.o_cdof);
  float* sW = w.at(m.o_cin + m.c_sW);
  float* sWb = w.at(m.o_cin + m.c_sWb);
  float* lf = w.at(m.o_cin + m.c_lf);
  if (w.lane < m.ncon) {
    const float* off = w.at(m.o_cin + m.c_off) + w.lane * 3;
    const float* t1 = w.at(m.o_cin + m.c_t1) + w.lane * 3;
    float t2[3], fw[3], tau[3];
    cross3(m.plane_n, t1, t2);
    const float fn = f[0] + f[1] + f[2] + f[3], f1 = (f[0] - f[1]) * r.mu, f2 = (f[2] - f[3]) * r.mu;
    fw[0] = m.plane_n[0] * fn + t1[0] * f1 + t2[0] * f2;
    fw[1] = m.plane_n[1] * fn + t1[1] * f1 + t2[1] * f2;
    fw[2] = m.plane_n[2] * fn + t1[2] * f1 + t2[2] * f2;
    cross3(off, fw, tau);
    const float o6[6] = {tau[0], tau[1], tau[2], fw[0], fw[1], fw[2]};
    st_rec<6>(sW + w.lane * 6, o6);
  }
#pragma unroll
  for (int q = 0; q < kLimSlots; ++q) {
    const int l = w.lane + 32 * q;
    if (l < m.nlimit) lf[l] = r.lsign[q] * f[4 + q];
  }
  __syncwarp();
  if (m.use_seg) {
    // per contact body: sum of its contacts' wrenches; per segment: sum over the contact bodies below it; per dof: one
    // 6-vector product with its segment's wrench
    float* sP = w.at(m.o_cin + m.c_sP);
    const uint32_t st0 = m.seg_task[w.lane], st1 = m.seg_task[w.lane + 32], ct0 = m.cb_task[w.lane], ct1 = m.cb_task[w.lane + 32];
    const uint32_t ds3 = m.dof_seg3[w.lane];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const uint32_t ct = sl ? ct1 : ct0;
      const int k = (ct >> 16) & 7, c0 = (ct >> 20) & 0xff, cn = ct >> 28;
      const float* src = sW + c0 * 6 + k;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) if (i < cn) acc += src[i * 6];
      for (int i = 6; i < cn; ++i) acc += src[i * 6];
      if (ct & 0x80000u) sWb[w.lane + 32 * sl] = acc;
    }
    __syncwarp();
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const uint32_t st = sl ? st1 : st0;
      const int k = (st >> 12) & 7, b0 = (st >> 16) & 0xff, bn = (st >> 24) & 0xf;
      const float* src = sWb + b0 * 6 + k;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < bn) acc += src[i * 6];
      for (int i = 8; i < bn; ++i) acc += src[i * 6];
      if (st & 0x8000u) sP[w.lane + 32 * sl] = acc;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) {
      const int d = w.lane + 32 * q;
      const int sg = (ds3 >> (8 * q)) & 0xff;
      float acc = 0.f;
      if (sg != 0xff) {
        float W[6], cd[6];
        ld_rec<6>(sP + sg * 6, W);
        ld_rec<6>(cdof + d * 6, cd);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fmaf(cd[k], W[k], acc);
      }
      if (d < m.nv) { const int l = m.dof_limit[d]; if (l >= 0) acc += lf[l]; }
      out[q] = acc;
    }
    __syncwarp();
    return;
  }
  for (int t = w.lane; t < m.ncb * 6; t += 32) {
    const int cb = t / 6, k = t % 6;
    float acc = 0.f;
    for (int e = m.cb_con_start[cb]; e < m.cb_con_start[cb + 1]; ++e) acc += sW[m.cb_con[e] * 6 + k];
    sWb[t] = acc;
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) {
    const int d = w.lane + 32 * q;
    float acc = 0.f;
    if (d < m.nv) {
      for (int e = m.dof_cb_start[d]; e < m.dof_cb_start[d + 1]; ++e) {
        const float* W = sWb + m.dof_cb[e] * 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) acc += cdof[d * 6 + k] * W[k];
      }
      const int l = m.dof_limit[d];
      if (l >= 0) acc += lf[l];
    }
    out[q] = acc;
  }
  __syncwarp();
}

// collision_primitive.plane_{sphere,capsule,ellipsoid} + constraint.make_constraint (limits + pyramidal contacts)
__device__ void make_constraint(const Warp& w, const float com[3], Rows& r, float* dbg_dist) {
  const DevModel& m = w.m;
  const float* xpos = w.at(m.o_xpos);
  const float* xquat = w.at(m.o_xquat);
  const float* qpos = w.at(m.o_qpos);
  const float* n = m.plane_n;
  float cpos = 0.f;  // dist - includemargin of contact `lane`
  r.mu = 0.f; r.cact = false;
  float cD = 0.f, ck = 0.f, cb_ = 0.f, cimp = 0.f;
  if (w.lane < m.ncon) {
    const int c = w.lane, g = m.con_geom[c], b = m.cg_body[g], type = m.cg_type[g];
    const float* par = m.con_par + c * 12;
    float gpos[3], gq[4], t[3], pos[3], t1[3], dist;
    rot(m.cg_pos + g * 3, xquat + b * 4, t);
    gpos[0] = xpos[b * 3] + t[0]; gpos[1] = xpos[b * 3 + 1] + t[1]; gpos[2] = xpos[b * 3 + 2] + t[2];
    qmul(xquat + b * 4, m.cg_quat + g * 4, gq);
    const float sz0 = m.cg_size[g * 3], sz1 = m.cg_size[g * 3 + 1], sz2 = m.cg_size[g * 3 + 2];
    float R[9];
    q2mat(gq, R);
    if (type == kGeomEllipsoid) {
      float sv[3] = {(R[0] * n[0] + R[3] * n[1] + R[6] * n[2]) * sz0, (R[1] * n[0] + R[4] * n[1] + R[7] * n[2]) * sz1,
                     (R[2] * n[0] + R[5] * n[1] + R[8] * n[2]) * sz2};
      normalize3(sv);
      sv[0] = -sv[0] * sz0; sv[1] = -sv[1] * sz1; sv[2] = -sv[2] * sz2;
      pos[0] = gpos[0] + R[0] * sv[0] + R[1] * sv[1] + R[2] * sv[2];
      pos[1] = gpos[1] + R[3] * sv[0] + R[4] * sv[1] + R[5] * sv[2];
      pos[2] = gpos[2] + R[6] * sv[0] + R[7] * sv[1] + R[8] * sv[2];
      const float dl[3] = {pos[0] - m.plane_pos[0], pos[1] - m.plane_pos[1], pos[2] - m.plane_pos[2]};
      dist = dot3(n, dl);
      pos[0] = pos[0] - n[0] * dist * 0.5f; pos[1] = pos[1] - n[1] * dist * 0.5f; pos[2] = pos[2] - n[2] * dist * 0.5f;
    } else {
      float sp[3] = {gpos[0], gpos[1], gpos[2]};
      if (type == kGeomCapsule) {
        const float sg = float(m.con_side[c]);
        sp[0] += sg * R[2] * sz1; sp[1] += sg * R[5] * sz1; sp[2] += sg * R[8] * sz1;
      }
      const float dl[3] = {sp[0] - m.plane_pos[0], sp[1] - m.plane_pos[1], sp[2] - m.plane_pos[2]};
      dist = dot3(dl, n) - sz0;
      const float k = sz0 + 0.5f * dist;
      pos[0] = sp[0] - n[0] * k; pos[1] = sp[1] - n[1] * k; pos[2] = sp[2] - n[2] * k;
    }
    // tangent t1: capsule -> projected axis (fallback y/z), otherwise math.make_frame(n)[1]
    bool have = false;
    if (type == kGeomCapsule) {
      const float ax[3] = {R[2], R[5], R[8]};
      const float na = dot3(n, ax);
      t1[0] = ax[0] - n[0] * na; t1[1] = ax[1] - n[1] * na; t1[2] = ax[2] - n[2] * na;
      have = normalize3(t1) >= 0.5f;
      if (!have) {
        const bool usey = (-0.5f < n[1]) && (n[1] < 0.5f);
        t1[0] = 0.f; t1[1] = usey ? 1.f : 0.f; t1[2] = usey ? 0.f : 1.f;
      }
    } else {
      float a[3] = {n[0], n[1], n[2]};
      normalize3(a);
      const bool usey = (-0.5f < a[1]) && (a[1] < 0.5f);
      t1[0] = 0.f; t1[1] = usey ? 1.f : 0.f; t1[2] = usey ? 0.f : 1.f;
      const float ab = dot3(a, t1);
      t1[0] -= a[0] * ab; t1[1] -= a[1] * ab; t1[2] -= a[2] * ab;
      normalize3(t1);
    }
    float* off = w.at(m.o_cin + m.c_off) + c * 3;
    float* st1 = w.at(m.o_cin + m.c_t1) + c * 3;
    off[0] = pos[0] - com[0]; off[1] = pos[1] - com[1]; off[2] = pos[2] - com[2];
    st1[0] = t1[0]; st1[1] = t1[1]; st1[2] = t1[2];
    if (dbg_dist) dbg_dist[c] = dist;
    cpos = dist - par[9];
    r.cact = cpos < 0.f;
    r.mu = par[0];
    cimp = impedance(par + 4, cpos);
    ck = par[2]; cb_ = par[3];
    cD = 1.f / fmaxf(par[1] * (1.f - cimp) / cimp, kMinVal);
  }
  float lpos[kLimSlots], lk[kLimSlots], lb[kLimSlots], limp[kLimSlots];
#pragma unroll
  for (int q = 0; q < kLimSlots; ++q) {
    const int l = w.lane + 32 * q;
    r.lsign[q] = 0.f; lpos[q] = 0.f; lk[q] = lb[q] = limp[q] = 0.f;
    r.D[4 + q] = 0.f;
    if (l < m.nlimit) {
      const float* par = m.lim_par + l * 12;
      const float qv = qpos[m.lim_qadr[l]];
      const float dmin = qv - par[0], dmax = par[1] - qv;
      lpos[q] = fminf(dmin, dmax) - par[2];
      r.lsign[q] = lpos[q] < 0.f ? (dmin < dmax ? 1.f : -1.f) : 0.f;
      limp[q] = impedance(par + 6, lpos[q]);
      lk[q] = par[4]; lb[q] = par[5];
      r.D[4 + q] = 1.f / fmaxf(par[3] * (1.f - limp[q]) / limp[q], kMinVal);
    }
  }
  __syncwarp();
  float jq[kRowSlots];
  apply_J(w, r, w.at(m.o_qvel), jq);
#pragma unroll
  for (int k = 0; k < 4; ++k) { r.D[k] = cD; r.aref[k] = -cb_ * jq[k] - ck * cimp * cpos; }
#pragma unroll
  for (int q = 0; q < kLimSlots; ++q) r.aref[4 + q] = -lb[q] * jq[4 + q] - lk[q] * limp[q] * lpos[q];
  // rows that do not exist (lane >= ncon, l >= nlimit) must never activate: D = 0, aref = 0 => Jaref = 0 (not < 0)
  if (w.lane >= m.ncon) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { r.D[k] = 0.f; r.aref[k] = 0.f; }
  }
}

// ---------------------------------------------------------------------------------------------- Newton pieces
// L^T D L of ONE sparse matrix in place (depth-lane mapping: lane d owns column depth d, d + 32 in the hi slot);
// barrier-free for the same reason as factor_dual.  Used for the Newton Hessian, once per solver iteration.
__device__ void factor_single(const Warp& w, float* L) {
  const DevModel& m = w.m;
  const int lane = w.lane;
  float* pl = L - lane;
  __syncwarp();
  for (int k = m.nv - 1; k >= 0; --k) {
    const int c = m.u_depth[k], re = m.u_rowend[k];
    float* rk = pl + re;
    const bool hi = c >= 32;
    float l = 0.f, h = 0.f;
    if (lane <= c) l = rk[0];
    if (hi && lane + 32 <= c) h = rk[-32];
    const float d = __shfl_sync(FULLMASK, hi ? h : l, c & 31);
    const float inv = 1.f / d;
    const float wl = l * inv, wh = h * inv;
    if (lane < c) rk[0] = wl; else if (lane == c) rk[0] = inv;
    if (hi) { if (lane + 32 < c) rk[-32] = wh; else if (lane + 32 == c) rk[-32] = inv; }
    int ti = re - (c - 1);
    for (int al = c - 1; al >= 0; --al, ++ti) {
      float* t = pl + int(m.u_ancre[ti]);
      const float a = __shfl_sync(FULLMASK, al >= 32 ? h : l, al & 31);
      if (lane <= al) t[0] = fmaf(-a, wl, t[0]);
      if (al >= 32 && lane + 32 <= al) t[-32] = fmaf(-a, wh, t[-32]);
    }
  }
  __syncwarp();
}

// H = M + J^T diag(D * active) J in the sparse tree layout at o_L (solver.py Newton branch of update_gradient).
// Limit rows add D to a diagonal entry; the four pyramid rows of an active contact are supported on the contact body's
// root path, so J_r^T J_r only touches (dof, ancestor) pairs -- exactly the sparsity of M.  Depth-lane mapping: lane d
// holds the row values of the chain dof at depth d; each chain row is one broadcast + LDS / 4 FMA / STS.
__device__ void build_hessian(const Warp& w, const Rows& r, const float Jaref[kRowSlots]) {
  const DevModel& m = w.m;
  const int lane = w.lane;
  float* H = w.at(m.o_L);
  const float* M = w.at(m.o_big);
  const float* cdof = w.at(m.o_cdof);
  __syncwarp();
  for (int e = lane; e < m.nM; e += 32) H[e] = M[e];
  __syncwarp();
#pragma unroll
  for (int q = 0; q < kLimSlots; ++q) {
    const int l = lane + 32 * q;
    if (l < m.nlimit && Jaref[4 + q] < 0.f && r.lsign[q] != 0.f) H[m.dof_madr[m.lim_dof[l]]] += r.D[4 + q];
  }
  __syncwarp();
  // this lane's contact: D * active per pyramid row
  float dk[4];
  bool any = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const bool on = r.cact && Jaref[k] < 0.f; dk[k] = on ? r.D[k] : 0.f; any |= on; }
  unsigned todo = __ballot_sync(FULLMASK, any);
  float* pl = H - lane;
  while (todo) {
    const int c = __ffs(todo) - 1;
    todo &= todo - 1;
    const float d0 = __shfl_sync(FULLMASK, dk[0], c), d1 = __shfl_sync(FULLMASK, dk[1], c), d2 = __shfl_sync(FULLMASK, dk[2], c),
                d3 = __shfl_sync(FULLMASK, dk[3], c), mu = __shfl_sync(FULLMASK, r.mu, c);
    const int cb = m.con_cb[c], e0 = m.cb_chain_start[cb], len = m.cb_chain_start[cb + 1] - e0;
    const float* off = w.at(m.o_cin + m.c_off) + c * 3;
    const float* t1 = w.at(m.o_cin + m.c_t1) + c * 3;
    float j0 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f;  // the four row values at my chain dof
    if (lane < len) {
      const float* cd = cdof + m.cb_chain_dof[e0 + lane] * 6;
      float vel[3], t2[3];
      cross3(cd, off, vel);
      vel[0] += cd[3]; vel[1] += cd[4]; vel[2] += cd[5];
      cross3(m.plane_n, t1, t2);
      const float jn = dot3(m.plane_n, vel), ja = dot3(t1, vel), jb = dot3(t2, vel);
      j0 = jn + ja * mu; j1 = jn - ja * mu; j2 = jn + jb * mu; j3 = jn - jb * mu;
    }
    const float a0 = j0 * d0, a1 = j1 * d1, a2 = j2 * d2, a3 = j3 * d3;
    for (int al = 0; al < len; ++al) {
      const int re = m.u_rowend[m.cb_chain_dof[e0 + al]];
      const float b0 = __shfl_sync(FULLMASK, a0, al), b1 = __shfl_sync(FULLMASK, a1, al), b2 = __shfl_sync(FULLMASK, a2, al),
                  b3 = __shfl_sync(FULLMASK, a3, al);
      if (lane <= al) pl[re] += b0 * j0 + b1 * j1 + b2 * j2 + b3 * j3;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- solver.solve (CG)
struct LSPoint { float alpha, cost, d0, d1; };

// One line-search evaluation at N step sizes.  solver.py sums the per-row quadratic coefficients
// (q0, q1, q2) = (D Jaref^2 / 2, D jv Jaref, D jv^2 / 2) of the active rows and evaluates cost = a^2 q2 + a q1 + q0 and its
// derivatives; with t = Jaref + a jv (which the activity test needs anyway) the same three sums are
//   cost = sum D t^2 / 2,   d/da = sum D t jv,   d2/da2 = sum D jv^2
// -- identical in exact arithmetic, one FMUL + two predicated FFMA + one predicated FADD per (row, point) and two
// registers per row (Dj = D jv, Djj = D jv^2) instead of three.
template <int N>
__device__ __forceinline__ void ls_points(const float (&alpha)[N], const float Jaref[kRowSlots], const float jv[kRowSlots],
                                          const float D[kRowSlots], const float Dj[kRowSlots], const float Djj[kRowSlots],
                                          const float qg[3], LSPoint (&out)[N]) {
  float a0[N], a1[N], a2[N];
#pragma unroll
  for (int p = 0; p < N; ++p) {
    a0[p] = a1[p] = a2[p] = 0.f;
#pragma unroll
    for (int k = 0; k < kRowSlots; ++k) {
      const float t = Jaref[k] + alpha[p] * jv[k];
      const float u = D[k] * t;
      if (t < 0.f) { a0[p] = fmaf(u, t, a0[p]); a1[p] = fmaf(Dj[k], t, a1[p]); a2[p] += Djj[k]; }
    }
  }
  if constexpr (N == 3) {
    float v[9] = {a0[0], a1[0], a2[0], a0[1], a1[1], a2[1], a0[2], a1[2], a2[2]};
    wsum9(v, int(threadIdx.x) & 31);
#pragma unroll
    for (int p = 0; p < N; ++p) { a0[p] = v[3 * p]; a1[p] = v[3 * p + 1]; a2[p] = v[3 * p + 2]; }
  } else {
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
      for (int p = 0; p < N; ++p) {
        a0[p] += __shfl_xor_sync(FULLMASK, a0[p], o);
        a1[p] += __shfl_xor_sync(FULLMASK, a1[p], o);
        a2[p] += __shfl_xor_sync(FULLMASK, a2[p], o);
      }
  }
#pragma unroll
  for (int p = 0; p < N; ++p) {
    const float a = alpha[p];
    const float q2 = qg[2] + 0.5f * a2[p];   // the total quadratic coefficient (solver.py quad_total[2])
    out[p].alpha = a;
    out[p].cost = 0.5f * a0[p] + (a * a * qg[2] + a * qg[1] + qg[0]);
    out[p].d0 = a1[p] + (2.f * a * qg[2] + qg[1]);
    out[p].d1 = 2.f * q2 + (q2 == 0.f ? kMinVal : 0.f);
  }
}

struct SolverOut { float qacc[kNvSlots], qfc[kNvSlots], force[kRowSlots]; };

__device__ void solve_cg(const Warp& w, const Rows& r, const float qfs[kNvSlots], const float qas[kNvSlots],
                         const float Maw[kNvSlots], SolverOut& so) {
  const DevModel& m = w.m;
  float* sx = w.at(m.o_cin + m.c_sx);
  const float* L1 = w.at(m.o_L);
  // the residency variants are solver-specialised: 14 warps = CG only, 10 warps = Newton only (its slice carries a third
  // matrix), 4 warps = either (runtime); dropping the other solver's code shrinks the hot kernel's instruction footprint
#if TMJX_VARIANT >= 14 || TMJX_VARIANT == 7
  constexpr bool newton = false;
#elif TMJX_VARIANT == 10
  constexpr bool newton = true;
#else
  const bool newton = m.solver == TMJX_SOLVER_NEWTON;
#endif
  float warm[kNvSlots];
  vget(w, w.at(m.o_warm), warm);
  PT_DECL;

  // ---- warm-start choice: cost(qacc_warmstart) vs cost(qacc_smooth)
  float Jw[kRowSlots], Js[kRowSlots];
  apply_J(w, r, w.at(m.o_warm), Jw);
  vput(w, sx, qas);
  __syncwarp();
  apply_J(w, r, sx, Js);
  float cw = 0.f, cs = 0.f;
#pragma unroll
  for (int k = 0; k < kRowSlots; ++k) {
    Jw[k] -= r.aref[k]; Js[k] -= r.aref[k];
    cw += (Jw[k] < 0.f) ? r.D[k] * Jw[k] * Jw[k] : 0.f;
    cs += (Js[k] < 0.f) ? r.D[k] * Js[k] * Js[k] : 0.f;
  }
  float gw = 0.f;
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) gw += (Maw[q] - qfs[q]) * (warm[q] - qas[q]);
  cw = 0.5f * wsum(cw) + 0.5f * wsum(gw);
  cs = 0.5f * wsum(cs);  // gauss(qacc_smooth) == 0 exactly
  const bool use_warm = cw < cs;
  PT_LAP(16);

  float qacc[kNvSlots], Ma[kNvSlots], Jaref[kRowSlots];
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) { qacc[q] = use_warm ? warm[q] : qas[q]; Ma[q] = use_warm ? Maw[q] : qfs[q]; }
#pragma unroll
  for (int k = 0; k < kRowSlots; ++k) Jaref[k] = use_warm ? Jw[k] : Js[k];

  // mv = M search is carried by recurrence: search = -M^-1 grad + beta search  =>  M search = -grad + beta (M search_prev)
  float force[kRowSlots], qfc[kNvSlots], grad[kNvSlots], Mgrad[kNvSlots], search[kNvSlots], mv[kNvSlots];
  float gauss, cost = use_warm ? cw : cs, prev_cost = __int_as_float(0x7f800000);
  auto update_constraint = [&](bool with_cost) {
    float c = 0.f;
#pragma unroll
    for (int k = 0; k < kRowSlots; ++k) {
      const bool act = Jaref[k] < 0.f;
      force[k] = act ? r.D[k] * -Jaref[k] : 0.f;
      c += act ? r.D[k] * Jaref[k] * Jaref[k] : 0.f;
    }
    apply_JT(w, r, force, qfc);
    float g = 0.f;
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) g += (Ma[q] - qfs[q]) * (qacc[q] - qas[q]);
    gauss = 0.5f * wsum(g);
    if (with_cost) { prev_cost = cost; cost = 0.5f * wsum(c) + gauss; }
  };
  auto update_gradient = [&]() {
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) grad[q] = Ma[q] - qfs[q] - qfc[q];
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) Mgrad[q] = grad[q];
    if (newton) {  // H = M + J^T diag(D active) J, assembled and factored in the o_L block (same tree sparsity as M)
      build_hessian(w, r, Jaref);
      if (m.use_gen) { __syncwarp(); gen::factor_dual(w.at(m.o_L), w.lane, false, 0); __syncwarp(); }   // inside `if (active)`: no block barrier here
      else factor_single(w, w.at(m.o_L));
    }
    solve_ld(w, L1, Mgrad);
  };
  // Context.create: cost = inf -> update_constraint sets prev_cost = inf, cost = c
  {
    cost = __int_as_float(0x7f800000);
    update_constraint(true);
    update_gradient();
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) { search[q] = -Mgrad[q]; mv[q] = -grad[q]; }
  }
  PT_LAP(17);
  const float scale = m.meaninertia_scale;
  // The iteration count is block-uniform (phase barriers inside): an environment whose solver has terminated
  // (solver.py's while_loop cond) is masked for the remaining rounds instead of breaking out.
  const int max_iter = m.iterations != 1 ? m.iterations : 1;
  bool active = true;
  for (int iter = 0; iter < max_iter; ++iter) {
    if (m.sync_mask & 16) phase_sync();
    if (active && m.iterations != 1) {
      const float improvement = (prev_cost - cost) / scale;
      const float gradient = sqrtf(vdot(grad, grad)) / scale;
      if (improvement < m.tolerance || gradient < m.tolerance) active = false;
    }
    if (active) {
    // ---- _linesearch
    const float smag = sqrtf(vdot(search, search)) * scale;
    const float gtol = m.tolerance * m.ls_tolerance * smag;
    float jv[kRowSlots];
    vput(w, sx, search);
    __syncwarp();
    if (newton) {  // CG carries M*search by recurrence; Newton's direction has none
      if (m.use_gen) {
        gen::V3 v;
        v.a = search[0]; v.b = search[1]; v.c = search[2];
        v = gen::mul_m(w.at(m.o_big), v, w.lane, w.dep[0], w.dep[1], w.dep[2], w.rend[0], w.rend[1], w.rend[2]);
        mv[0] = v.a; mv[1] = v.b; mv[2] = v.c;
      } else {
        mul_m_raw(w, sx, mv);
      }
    }
    PT_LAP(18);
    apply_J(w, r, sx, jv);
    PT_LAP(19);
    float qg[3];
    {
      float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
      for (int q = 0; q < kNvSlots; ++q) { a += search[q] * Ma[q]; b += search[q] * qfs[q]; c += search[q] * mv[q]; }
      qg[0] = gauss; qg[1] = wsum(a) - wsum(b); qg[2] = 0.5f * wsum(c);
    }
    float Dj[kRowSlots], Djj[kRowSlots];
#pragma unroll
    for (int k = 0; k < kRowSlots; ++k) { Dj[k] = r.D[k] * jv[k]; Djj[k] = Dj[k] * jv[k]; }
    LSPoint p0, lo, hi;
    {
      const float a0[1] = {0.f};
      LSPoint o1[1];
      ls_points<1>(a0, Jaref, jv, r.D, Dj, Djj, qg, o1);
      p0 = o1[0];
      const float a1[1] = {p0.alpha - __fdividef(p0.d0, p0.d1)};   // Newton step on the 1-D cost; MUFU.RCP (2 ulp) instead of the IEEE sequence
      ls_points<1>(a1, Jaref, jv, r.D, Dj, Djj, qg, o1);
      lo = o1[0];
      const bool lesser = lo.d0 < p0.d0;
      hi = lesser ? p0 : lo;
      lo = lesser ? lo : p0;
    }
    bool swap = true;
    for (int it = 0;; ++it) {
      bool done = it >= m.ls_iterations;
      done |= !swap;
      done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
      done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
      if (done) break;
      const float al[3] = {lo.alpha - __fdividef(lo.d0, lo.d1), hi.alpha - __fdividef(hi.d0, hi.d1), 0.5f * (lo.alpha + hi.alpha)};
      LSPoint o3[3];
      ls_points<3>(al, Jaref, jv, r.D, Dj, Djj, qg, o3);
      const LSPoint lo_next = o3[0], hi_next = o3[1], mid = o3[2];
      const bool swap_lo_next = (lo.d0 > 0.f) || (lo.d0 < lo_next.d0);
      if (swap_lo_next) lo = lo_next;
      const bool swap_lo_mid = (mid.d0 < 0.f) && (lo.d0 < mid.d0);
      if (swap_lo_mid) lo = mid;
      const bool swap_hi_next = (hi.d0 < 0.f) || (hi.d0 > hi_next.d0);
      if (swap_hi_next) hi = hi_next;
      const bool swap_hi_mid = (mid.d0 > 0.f) && (hi.d0 > mid.d0);
      if (swap_hi_mid) hi = mid;
      swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
    }
    PT_LAP(20);
    const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
    const float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
    if (improved) {
#pragma unroll
      for (int q = 0; q < kNvSlots; ++q) { qacc[q] += search[q] * alpha; Ma[q] += mv[q] * alpha; }
#pragma unroll
      for (int k = 0; k < kRowSlots; ++k) Jaref[k] += jv[k] * alpha;
    }
    }
    if (m.sync_mask & 64) phase_sync();
    if (active) {
    // ---- body: update + Polak-Ribiere
    float pg[kNvSlots], pMg[kNvSlots];
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) { pg[q] = grad[q]; pMg[q] = Mgrad[q]; }
    PT_LAP(21);
    update_constraint(true);
    PT_LAP(22);
    update_gradient();
    PT_LAP(23);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) { num += grad[q] * (Mgrad[q] - pMg[q]); den += pg[q] * pMg[q]; }
    float beta = wsum(num) / fmaxf(kMinVal, wsum(den));
    beta = fmaxf(0.f, beta);
    if (newton) beta = 0.f;  // search = -H^-1 grad
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) { search[q] = -Mgrad[q] + beta * search[q]; mv[q] = -grad[q] + beta * mv[q]; }
    PT_LAP(24);
    }
  }
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) { so.qacc[q] = qacc[q]; so.qfc[q] = qfc[q]; }
#pragma unroll
  for (int k = 0; k < kRowSlots; ++k) so.force[k] = force[k];
}

// ---------------------------------------------------------------------------------------------- one mjx.forward
struct FwdOut {
  float qfa[kNvSlots], qfs[kNvSlots], qas[kNvSlots], bias[kNvSlots], actdot[2], com[3];
  SolverOut so;
};

__device__ void forward(const Warp& w, FwdOut& fo, float* dbg_dist, bool sync_here = true) {
  const DevModel& m = w.m;
  PT_DECL;
  if (m.sync_level >= 0 && sync_here) phase_sync();
  PT_LAP(0);
  kinematics(w);
  PT_LAP(1);
  if (m.sync_level > 1) phase_sync();
  com_pos(w, fo.com);
  PT_LAP(2);
  if (m.sync_level > 1) phase_sync();
  com_vel_rne(w, fo.bias);
  PT_LAP(3);
  if (m.sync_mask & 1) phase_sync();
  passive_actuation(w, fo.bias, fo.qfa, fo.qfs, fo.actdot);
  __syncwarp();
  PT_LAP(4);
  build_m(w);
  PT_LAP(5);
  if (m.sync_mask & 2) phase_sync();
  float Maw[kNvSlots];   // M qacc_warmstart, while o_big still holds the raw inertia
  if (m.use_gen) {
    float wv[kNvSlots];
    vget(w, w.at(m.o_warm), wv);
    gen::V3 v;
    v.a = wv[0]; v.b = wv[1]; v.c = wv[2];
    v = gen::mul_m(w.at(m.o_big), v, w.lane, w.dep[0], w.dep[1], w.dep[2], w.rend[0], w.rend[1], w.rend[2]);
    Maw[0] = v.a; Maw[1] = v.b; Maw[2] = v.c;
  } else {
    mul_m_raw(w, w.at(m.o_warm), Maw);
  }
  __syncwarp();
  PT_LAP(6);
  if (m.use_gen) { gen::factor_dual(w.at(m.o_L), w.lane, m.sync_level > 1, gen::kNMpad); __syncwarp(); } else factor_dual(w);
  PT_LAP(7);
  if (m.l2_spill) {
    // park the head of the Euler factor (it sits where the solver scratch is about to go) in global memory: written once,
    // read back once per substep by euler(), 2.7 KB per environment that never leaves the L2 cache
    const float4* src = reinterpret_cast<const float4*>(w.at(m.o_L2));
    float4* dst = reinterpret_cast<float4*>(w.spill);
    for (int i = w.lane; i < m.spill_floats / 4; i += 32) dst[i] = src[i];
    __syncwarp();
  }
  PT_LAP(8);
  if (m.sync_mask & 4) phase_sync();
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) fo.qas[q] = fo.qfs[q];
  solve_ld(w, w.at(m.o_L), fo.qas);
  PT_LAP(9);
  Rows r;
  make_constraint(w, fo.com, r, dbg_dist);
  PT_LAP(10);
  if (m.sync_mask & 8) phase_sync();
  solve_cg(w, r, fo.qfs, fo.qas, Maw, fo.so);
  PT_LAP(11);
  vput(w, w.at(m.o_warm), fo.so.qacc);
  __syncwarp();
}

// forward.euler + _advance
__device__ void euler(const Warp& w, const FwdOut& fo, float& time) {
  const DevModel& m = w.m;
  PT_DECL;
  float qacc[kNvSlots];
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) qacc[q] = fo.qfs[q] + fo.so.qfc[q];
  if (m.l2_spill) {   // the solver scratch is dead: bring the head of the Euler factor back behind its tail
    __syncwarp();
    const float4* src = reinterpret_cast<const float4*>(w.spill);
    float4* dst = reinterpret_cast<float4*>(w.at(m.o_L2));
    for (int i = w.lane; i < m.spill_floats / 4; i += 32) dst[i] = src[i];
    __syncwarp();
  }
  solve_ld(w, w.at(m.o_L2), qacc);
  float* qpos = w.at(m.o_qpos);
  float* qvel = w.at(m.o_qvel);
  float* act = w.at(m.o_act);
  if (m.na) {
#pragma unroll
    for (int k = 0; k < 2; ++k) { const int u = w.lane + 32 * k; if (u < m.nu) act[u] = act[u] + fo.actdot[k] * m.dt; }
  }
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) {
    const int d = w.lane + 32 * q;
    if (d < m.nv) {
      const float v = qvel[d] + qacc[q] * m.dt;
      qvel[d] = v;
      const int qa = m.dof_qadr[d];
      if (qa >= 0) qpos[qa] = qpos[qa] + m.dt * v;
    }
  }
  __syncwarp();
  for (int j = w.lane; j < m.njnt; j += 32) {
    if (m.jnt_type[j] != kJntFree) continue;
    const int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    qpos[qa] = qpos[qa] + m.dt * qvel[da]; qpos[qa + 1] = qpos[qa + 1] + m.dt * qvel[da + 1];
    qpos[qa + 2] = qpos[qa + 2] + m.dt * qvel[da + 2];
    float v[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]};
    const float nrm = normalize3(v);
    float sn, cs;
    sincosf(m.dt * nrm * 0.5f, &sn, &cs);
    const float qr[4] = {cs, v[0] * sn, v[1] * sn, v[2] * sn};
    float q2[4];
    qmul(qpos + qa + 3, qr, q2);
    normalize4(q2);
    qpos[qa + 3] = q2[0]; qpos[qa + 4] = q2[1]; qpos[qa + 5] = q2[2]; qpos[qa + 6] = q2[3];
  }
  time = __fadd_rn(time, m.dt);
  __syncwarp();
  PT_LAP(12);
}

// ---------------------------------------------------------------------------------------------- task layer
// jnp.minimum propagates NaN (fminf returns the non-NaN operand)
__device__ __forceinline__ float jminf(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float nan_to_num(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
  return x;
}
// _get_cur_frame (single_clip_tracking.py:452-454): unfused fp32 multiply then add, floor, int
__device__ __forceinline__ int cur_frame_of(float time, float mocap_hz, int start_frame) {
  return int(floorf(__fadd_rn(__fmul_rn(time, mocap_hz), float(start_frame))));
}

// _get_obs (single_clip_tracking.py:394-450) + walker transforms (walker/base.py:170-258); writes obs to global
__device__ void write_obs(const Warp& w, const DevTask& t, const float* __restrict__ clips, int clip_len, int clip, int frame,
                          const float qfa[kNvSlots], float* __restrict__ obs, bool& bad) {
  const DevModel& m = w.m;
  const TmjxTaskConfig& cfg = t.cfg;
  const float* qpos = w.at(m.o_qpos);
  const float* qvel = w.at(m.o_qvel);
  const float* xpos = w.at(m.o_xpos);
  const float* xquat = w.at(m.o_xquat);
  const int L = cfg.traj_length;
  const int start = min(max(frame + 1, 0), clip_len - L);  // dynamic_slice clamps the start
  const float quat[4] = {qpos[3], qpos[4], qpos[5], qpos[6]};
  const int nji = cfg.n_joint_idxs, nbi = cfg.n_body_idxs;
  float* o_track = obs;
  float* o_quat = o_track + 3 * L;
  float* o_joint = o_quat + 4 * L;
  float* o_body = o_joint + nji * L;
  float* o_prop = o_body + 3 * nbi * L;
  const size_t base = (size_t(clip) * clip_len + start) * t.frame_stride;
  for (int i = w.lane; i < L; i += 32) {
    const float* fr = clips + base + size_t(i) * t.frame_stride;
    const float dl[3] = {fr[t.o_pos] - qpos[0], fr[t.o_pos + 1] - qpos[1], fr[t.o_pos + 2] - qpos[2]};
    float r3[3];
    rot(dl, quat, r3);
    o_track[i * 3] = nan_to_num(r3[0]); o_track[i * 3 + 1] = nan_to_num(r3[1]); o_track[i * 3 + 2] = nan_to_num(r3[2]);
    const float rq[4] = {fr[t.o_quat], -fr[t.o_quat + 1], -fr[t.o_quat + 2], -fr[t.o_quat + 3]};
    float q4[4];
    qmul(quat, rq, q4);
    for (int k = 0; k < 4; ++k) o_quat[i * 4 + k] = nan_to_num(q4[k]);
  }
  for (int i = w.lane; i < L * nji; i += 32) {
    const int tt = i / nji, c = t.joint_col[i % nji];
    const float* fr = clips + base + size_t(tt) * t.frame_stride;
    o_joint[i] = nan_to_num(fr[t.o_joints + c] - qpos[7 + c]);
  }
  for (int i = w.lane; i < L * nbi; i += 32) {
    const int tt = i / nbi, bi = i % nbi;
    const float* fr = clips + base + size_t(tt) * t.frame_stride + t.o_bodies + 3 * t.body_slot[bi];
    const int row = t.body_row[bi] + 1;
    const float dl[3] = {fr[0] - xpos[row * 3], fr[1] - xpos[row * 3 + 1], fr[2] - xpos[row * 3 + 2]};
    float r3[3];
    rot(dl, quat, r3);
    o_body[i * 3] = nan_to_num(r3[0]); o_body[i * 3 + 1] = nan_to_num(r3[1]); o_body[i * 3 + 2] = nan_to_num(r3[2]);
  }
  // proprioception
  const int nj = m.nq - 7, nvj = m.nv - 6;
  for (int i = w.lane; i < nj; i += 32) o_prop[i] = nan_to_num(qpos[7 + i]);
  for (int i = w.lane; i < nvj; i += 32) o_prop[nj + i] = nan_to_num(qvel[6 + i]);
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) { const int d = w.lane + 32 * q; if (d < m.nv) o_prop[nj + nvj + d] = nan_to_num(qfa[q]); }
  float* o_tail = o_prop + nj + nvj + m.nv;
  const int tb = cfg.torso_body_id;
  float R[9];
  q2mat(xquat + tb * 4, R);
  if (w.lane == 0) {
    o_tail[0] = nan_to_num(xpos[tb * 3 + 2]);
    o_tail[1] = nan_to_num(R[6]); o_tail[2] = nan_to_num(R[7]); o_tail[3] = nan_to_num(R[8]);
  }
  for (int i = w.lane; i < cfg.n_appendages * 3; i += 32) {
    const int a = i / 3, c = i % 3, b = cfg.appendage_body_ids[a];
    const float dl[3] = {xpos[b * 3] - xpos[tb * 3], xpos[b * 3 + 1] - xpos[tb * 3 + 1], xpos[b * 3 + 2] - xpos[tb * 3 + 2]};
    o_tail[4 + i] = nan_to_num(dl[0] * R[c] + dl[1] * R[3 + c] + dl[2] * R[6 + c]);
  }
  (void)bad;
}


// kWPB warps (= environments) per block; (4, 3) and (7, 2) are the two residency points that matter on B200:
// 12 and 14 resident environments per SM (the latter needs <= 144 registers and fits 4096 envs in two waves of 148 SMs)
template <bool kStep, int kWPB, int kMinBlocks>
__global__ void __launch_bounds__(kWPB * 32, kMinBlocks) tmjx_env_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ float smem[];
  const DevModel& m = a.m;
  const DevTask& t = *a.task;
  const TmjxTaskConfig& cfg = t.cfg;
  const int wpb = kWPB, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Warp w{m, smem + size_t(warp) * m.smem_floats, lane, {0, 0, 0}, {0, 0, 0},
         a.spill + (size_t(blockIdx.x) * kWPB + warp) * size_t(m.spill_floats)};
#pragma unroll
  for (int q = 0; q < kNvSlots; ++q) { w.dep[q] = m.depth_me[q * 32 + lane]; w.rend[q] = m.rowend_me[q * 32 + lane]; }
  const int nu = m.nu, nobs = t.obs_size, W = cfg.var_window_size;
  // Environment e belongs to block e % gridDim.x (so a partly filled last round is spread over all SMs).  The number
  // of rounds is block-uniform; a warp without an environment in the last round recomputes the block's last one
  // (reads only, nothing is written) so that it keeps arriving at the phase barriers.
  const int G = gridDim.x, count = (a.n_env - int(blockIdx.x) + G - 1) / G, nrounds = (count + wpb - 1) / wpb;
  if (a.phase_offset_ns > 0 && int(blockIdx.x) >= (G + 1) / 2) {
    // the second block of an SM runs half a substep behind the first: the two blocks then sit in different phases (tree passes /
    // factorisation vs solver) and compete less for the shuffle / shared-memory pipe than 14 warps in one phase
    const long long t0 = clock64(), wait = (long long)(a.phase_offset_ns) * 2;   // ~2 cycles per ns
    while (clock64() - t0 < wait) __nanosleep(1000);
  }
  for (int rd = 0; rd < nrounds; ++rd) {
    const int slot = rd * wpb + warp;
    const bool live = slot < count;
    const int e = (live ? slot : count - 1) * G + int(blockIdx.x);
    // ---- stage the persistent state
    float* qpos = w.at(m.o_qpos);
    float* qvel = w.at(m.o_qvel);
    float* act = w.at(m.o_act);
    float* ctrl = w.at(m.o_ctrl);
    float* warm = w.at(m.o_warm);
    for (int i = lane; i < m.nq; i += 32) qpos[i] = a.st.qpos[size_t(e) * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) qvel[i] = a.st.qvel[size_t(e) * m.nv + i];
    float time;
    if (kStep) {
      for (int i = lane; i < m.na; i += 32) act[i] = a.st.act[size_t(e) * m.na + i];
      for (int i = lane; i < m.nv; i += 32) warm[i] = a.st.qacc_warmstart[size_t(e) * m.nv + i];
      for (int i = lane; i < nu; i += 32) {
        float c = a.action[size_t(e) * nu + i];
        if (m.act_flags[i] & 1) c = fminf(fmaxf(c, m.act_ctrl_lo[i]), m.act_ctrl_hi[i]);
        ctrl[i] = c;
      }
      time = a.st.time[e];
    } else {
      for (int i = lane; i < m.na; i += 32) act[i] = 0.f;
      for (int i = lane; i < m.nv; i += 32) warm[i] = 0.f;
      for (int i = lane; i < nu; i += 32) ctrl[i] = 0.f;
      time = 0.f;
    }
    float prev_done = 0.f;
    if (kStep && (a.flags & TMJX_F_AUTORESET)) prev_done = a.out.done[e];
    __syncwarp();

    PT_DECL;
    FwdOut fo;
    float* dbg_dist = (live && a.out.dbg_contact_dist) ? a.out.dbg_contact_dist + size_t(e) * m.ncon : nullptr;
    if (kStep && (a.flags & TMJX_F_EPILOGUE_ONLY)) {
      // test hook (block-uniform): the given state IS the post-physics state; only the task layer below runs on it
      for (int i = lane; i < m.nbody * 3; i += 32) w.at(m.o_xpos)[i] = a.st.xpos[size_t(e) * m.nbody * 3 + i];
      for (int i = lane; i < m.nbody * 4; i += 32) w.at(m.o_xquat)[i] = a.st.xquat[size_t(e) * m.nbody * 4 + i];
#pragma unroll
      for (int q = 0; q < kNvSlots; ++q) {
        const int d = lane + 32 * q;
        fo.qfa[q] = d < m.nv ? a.st.qfrc_actuator[size_t(e) * m.nv + d] : 0.f;
        fo.qfs[q] = fo.qas[q] = fo.bias[q] = fo.so.qacc[q] = fo.so.qfc[q] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < kRowSlots; ++k) fo.so.force[k] = 0.f;
      fo.actdot[0] = fo.actdot[1] = 0.f; fo.com[0] = fo.com[1] = fo.com[2] = 0.f;
      __syncwarp();
    } else if (kStep) {
      for (int f = 0; f < m.n_frames; ++f) {
        forward(w, fo, f == m.n_frames - 1 ? dbg_dist : nullptr, m.sync_every <= 1 || f % m.sync_every == 0);
        euler(w, fo, time);
      }
    } else {
      forward(w, fo, dbg_dist);
    }

    PT_LAP(13);
    if (live) {
    // ---- NaN scan over the state this build materialises (stand-in for ravel_pytree(data), :290-293)
    bool bad = isnan(time);
    for (int i = lane; i < m.nq; i += 32) bad |= isnan(qpos[i]);
    for (int i = lane; i < m.nv; i += 32) bad |= isnan(qvel[i]) || isnan(warm[i]);
    for (int i = lane; i < m.na; i += 32) bad |= isnan(act[i]);
    for (int i = lane; i < m.nbody * 3; i += 32) bad |= isnan(w.at(m.o_xpos)[i]);
    for (int i = lane; i < m.nbody * 4; i += 32) bad |= isnan(w.at(m.o_xquat)[i]);
#pragma unroll
    for (int q = 0; q < kNvSlots; ++q) bad |= isnan(fo.qfa[q]) || isnan(fo.so.qfc[q]) || isnan(fo.bias[q]) || isnan(fo.qas[q]);
#pragma unroll
    for (int k = 0; k < kRowSlots; ++k) bad |= isnan(fo.so.force[k]);
    bad = __any_sync(FULLMASK, bad);

    // jnp gathers clamp out-of-range indices (x[info["clip_idx"]], multi_clip_tracking.py:98-109): so does the table lookup here
    const int clip = min(max(a.st.clip_idx[e], 0), a.n_clips - 1), sf = a.st.start_frame[e];
    const int frame = cur_frame_of(time, cfg.mocap_hz, sf);
    float* obs = a.out.obs + size_t(e) * nobs;
    float done = 0.f;

    if (kStep) {
      const float* xpos = w.at(m.o_xpos);
      const int fcl = min(max(frame, 0), a.clip_len - 1);
      const float* fr = a.clips + (size_t(clip) * a.clip_len + fcl) * t.frame_stride;
      const float* action = a.action + size_t(e) * nu;
      // info updates (:227-234)
      float* buf = a.st.action_buffer + size_t(e) * W * nu;
      int idx = a.st.buffer_index[e];
      float a2 = 0.f;
      for (int i = lane; i < nu; i += 32) {
        const float v = action[i];
        a.st.prev_ctrl[size_t(e) * nu + i] = v;
        buf[idx * nu + i] = v;
        a2 += v * v;
      }
      const int idx_w = idx;
      idx = (idx + 1) % W;
      // ---- compute_tracking_rewards (reward.py:359-485)
      const float pd[3] = {qpos[0] - fr[t.o_pos], qpos[1] - fr[t.o_pos + 1], qpos[2] - fr[t.o_pos + 2]};
      const float pos_reward = cfg.pos_reward_weight * expf(-cfg.pos_reward_exp_scale * (pd[0] * pd[0] + pd[1] * pd[1] + pd[2] * pd[2]));
      float qs[4] = {qpos[3], qpos[4], qpos[5], qpos[6]}, qt[4] = {fr[t.o_quat], fr[t.o_quat + 1], fr[t.o_quat + 2], fr[t.o_quat + 3]};
      {
        const float ns = sqrtf(qs[0] * qs[0] + qs[1] * qs[1] + qs[2] * qs[2] + qs[3] * qs[3]);
        const float nt = sqrtf(qt[0] * qt[0] + qt[1] * qt[1] + qt[2] * qt[2] + qt[3] * qt[3]);
        for (int k = 0; k < 4; ++k) { qs[k] = qs[k] / ns; qt[k] = qt[k] / nt; }
      }
      const float qd = qs[0] * qt[0] + qs[1] * qt[1] + qs[2] * qt[2] + qs[3] * qt[3];
      const float bq = 0.5f * acosf(jminf(1.f, 2.f * qd * qd - 1.f));
      const float quat_distance = bq * bq;
      const float quat_reward = cfg.quat_reward_weight * expf(-cfg.quat_reward_exp_scale * quat_distance);
      float jd = 0.f;
      for (int j = lane; j < m.nq - 7; j += 32) { const float x = qpos[7 + j] - fr[t.o_joints + j]; jd += x * x; }
      const float joint_distance = wsum(jd);
      const float joint_reward = cfg.joint_reward_weight * expf(-cfg.joint_reward_exp_scale * joint_distance);
      float av = 0.f;
      for (int k = 0; k < 3; ++k) { const float x = qvel[3 + k] - fr[t.o_angvel + k]; av += x * x; }
      const float angvel_reward = cfg.angvel_reward_weight * expf(-cfg.angvel_reward_exp_scale * av);
      float be = 0.f, ee = 0.f;
      for (int i = lane; i < cfg.n_body_idxs * 3; i += 32) {
        const int bi = i / 3, k = i % 3;
        const float x = xpos[(t.body_row[bi] + 1) * 3 + k] - fr[t.o_bodies + 3 * t.body_slot[bi] + k];
        be += x * x;
      }
      for (int i = lane; i < cfg.n_endeff_idxs * 3; i += 32) {
        const int bi = i / 3, k = i % 3;
        const float x = xpos[(t.endeff_row[bi] + 1) * 3 + k] - fr[t.o_bodies + 3 * t.endeff_slot[bi] + k];
        ee += x * x;
      }
      const float bodypos_reward = cfg.bodypos_reward_weight * expf(-cfg.bodypos_reward_exp_scale * wsum(be));
      const float endeff_reward = cfg.endeff_reward_weight * expf(-cfg.endeff_reward_exp_scale * wsum(ee));
      const float ctrl_cost = cfg.ctrl_cost_weight * wsum(a2);
      const float ctrl_diff_cost = cfg.ctrl_diff_cost_weight * 0.f;  // prev_ctrl was overwritten first (:227)
      float en = 0.f;
#pragma unroll
      for (int q = 0; q < kNvSlots; ++q) { const int d = lane + 32 * q; if (d >= 6 && d < m.nv) en += fabsf(qvel[d]) * fabsf(fo.qfa[q]); }
      const float energy_cost = cfg.energy_cost_weight * jminf(wsum(en), 50.f);
      const float torso_z = xpos[cfg.torso_idx * 3 + 2];
      float healthy = torso_z < cfg.healthy_z_min ? 0.f : 1.f;
      if (torso_z > cfg.healthy_z_max) healthy = 0.f;
      const float fall = 1.f - healthy;
      float summed = 0.f;
      for (int k = 0; k < 3; ++k) { const float x = pd[k] * cfg.penalty_pos_distance_scale[k]; summed += x * x; }
      const float too_far = summed > cfg.too_far_dist ? 1.f : 0.f;
      const float bad_pose = joint_distance > cfg.bad_pose_dist ? 1.f : 0.f;
      const float bad_quat = quat_distance > cfg.bad_quat_dist ? 1.f : 0.f;
      // windowed variance + jerk over the ring buffer (reward.py:314-356); lane k owns columns k, k+32
      float var_sum = 0.f, jerk = 0.f;
      for (int k = lane; k < nu; k += 32) {
        const float mine = action[k];
        float mean = 0.f;
        for (int tt = 0; tt < W; ++tt) mean += (tt == idx_w) ? mine : buf[tt * nu + k];
        mean = mean / float(W);
        float v = 0.f;
        for (int tt = 0; tt < W; ++tt) { const float x = ((tt == idx_w) ? mine : buf[tt * nu + k]) - mean; v += x * x; }
        var_sum += v / float(W);
        float b0 = 0.f, b1 = 0.f;
        for (int tt = 0; tt < W; ++tt) {
          const int rr = (idx + tt) % W;
          const float b2 = (rr == idx_w) ? mine : buf[rr * nu + k];
          if (tt >= 2) { const float x = b2 - 2.f * b1 + b0; jerk += x * x; }
          b0 = b1; b1 = b2;
        }
      }
      const float var_cost = cfg.var_coeff * wsum(var_sum);
      const float jerk_cost = cfg.jerk_coeff * wsum(jerk);

      write_obs(w, t, a.clips, a.clip_len, clip, frame, fo.qfa, obs, bad);
      float reward = joint_reward + pos_reward + quat_reward + angvel_reward + bodypos_reward + endeff_reward - ctrl_cost -
                     ctrl_diff_cost - energy_cost - var_cost - jerk_cost;
      done = fmaxf(fmaxf(fall, too_far), fmaxf(bad_pose, bad_quat));
      reward = nan_to_num(reward);
      const float nanv = bad ? 1.f : 0.f;
      done = fmaxf(nanv, done);
      if (lane == 0) {
        float* mt = a.out.metrics + size_t(e) * TMJX_N_METRICS;
        mt[TMJX_M_POS_REWARD] = pos_reward; mt[TMJX_M_QUAT_REWARD] = quat_reward; mt[TMJX_M_JOINT_REWARD] = joint_reward;
        mt[TMJX_M_ANGVEL_REWARD] = angvel_reward; mt[TMJX_M_BODYPOS_REWARD] = bodypos_reward; mt[TMJX_M_ENDEFF_REWARD] = endeff_reward;
        mt[TMJX_M_CTRL_COST] = -ctrl_cost; mt[TMJX_M_CTRL_DIFF_COST] = -ctrl_diff_cost; mt[TMJX_M_ENERGY_COST] = -energy_cost;
        mt[TMJX_M_DONE] = done; mt[TMJX_M_TOO_FAR] = too_far; mt[TMJX_M_BAD_POSE] = bad_pose; mt[TMJX_M_BAD_QUAT] = bad_quat;
        mt[TMJX_M_FALL] = fall; mt[TMJX_M_NAN] = nanv; mt[TMJX_M_JOINT_DISTANCE] = joint_distance;
        mt[TMJX_M_SUMMED_POS_DISTANCE] = summed; mt[TMJX_M_QUAT_DISTANCE] = quat_distance; mt[TMJX_M_VAR_COST] = -var_cost;
        mt[TMJX_M_JERK_COST] = -jerk_cost;
        a.out.reward[e] = reward;
        a.out.cur_frame[e] = frame;
        a.st.buffer_index[e] = idx;
      }
    } else {
      write_obs(w, t, a.clips, a.clip_len, clip, frame, fo.qfa, obs, bad);
      for (int i = lane; i < TMJX_N_METRICS; i += 32) a.out.metrics[size_t(e) * TMJX_N_METRICS + i] = 0.f;
      for (int i = lane; i < W * nu; i += 32) a.st.action_buffer[size_t(e) * W * nu + i] = 0.f;
      for (int i = lane; i < nu; i += 32) a.st.prev_ctrl[size_t(e) * nu + i] = 0.f;
      if (lane == 0) { a.out.reward[e] = 0.f; a.out.cur_frame[e] = frame; a.st.buffer_index[e] = 0; }
    }

    // ---- debug taps
    if (a.out.dbg_qacc) for (int q = 0; q < kNvSlots; ++q) { const int d = lane + 32 * q; if (d < m.nv) a.out.dbg_qacc[size_t(e) * m.nv + d] = fo.so.qacc[q]; }
    if (a.out.dbg_qacc_smooth) for (int q = 0; q < kNvSlots; ++q) { const int d = lane + 32 * q; if (d < m.nv) a.out.dbg_qacc_smooth[size_t(e) * m.nv + d] = fo.qas[q]; }
    if (a.out.dbg_qfrc_bias) for (int q = 0; q < kNvSlots; ++q) { const int d = lane + 32 * q; if (d < m.nv) a.out.dbg_qfrc_bias[size_t(e) * m.nv + d] = fo.bias[q]; }
    if (a.out.dbg_qfrc_constraint) for (int q = 0; q < kNvSlots; ++q) { const int d = lane + 32 * q; if (d < m.nv) a.out.dbg_qfrc_constraint[size_t(e) * m.nv + d] = fo.so.qfc[q]; }
    if (a.out.dbg_efc_force) {  // MJX row order: limits, then 4 rows per contact
      float* ef = a.out.dbg_efc_force + size_t(e) * m.nefc;
      for (int q = 0; q < kLimSlots; ++q) { const int l = lane + 32 * q; if (l < m.nlimit) ef[l] = fo.so.force[4 + q]; }
      if (lane < m.ncon) for (int k = 0; k < 4; ++k) ef[m.nlimit + lane * 4 + k] = fo.so.force[k];
    }
    if (a.out.dbg_subtree_com && lane < 3) a.out.dbg_subtree_com[size_t(e) * 3 + lane] = fo.com[lane];

    // ---- wrappers (brax EpisodeWrapper + auto-reset, wrappers.py:104-144) when fused
    bool restore = false;
    if (kStep && (a.flags & TMJX_F_AUTORESET)) {
      float steps = a.st.steps[e];
      if (prev_done != 0.f) steps = 0.f;
      steps = steps + 1.f;
      const bool over = steps >= float(cfg.episode_length);
      const float trunc = over ? 1.f - done : 0.f;
      if (over) done = 1.f;
      if (lane == 0) { a.st.steps[e] = steps; a.st.truncation[e] = trunc; }
      restore = done != 0.f;
    }
    if (lane == 0) a.out.done[e] = done;

    // ---- write back the pipeline state (or the stored first state where done)
    __syncwarp();
    if (!restore) {
      for (int i = lane; i < m.nq; i += 32) a.st.qpos[size_t(e) * m.nq + i] = qpos[i];
      for (int i = lane; i < m.nv; i += 32) { a.st.qvel[size_t(e) * m.nv + i] = qvel[i]; a.st.qacc_warmstart[size_t(e) * m.nv + i] = warm[i]; }
      for (int i = lane; i < m.na; i += 32) a.st.act[size_t(e) * m.na + i] = act[i];
      for (int i = lane; i < m.nbody * 3; i += 32) a.st.xpos[size_t(e) * m.nbody * 3 + i] = w.at(m.o_xpos)[i];
      for (int i = lane; i < m.nbody * 4; i += 32) a.st.xquat[size_t(e) * m.nbody * 4 + i] = w.at(m.o_xquat)[i];
#pragma unroll
      for (int q = 0; q < kNvSlots; ++q) { const int d = lane + 32 * q; if (d < m.nv) a.st.qfrc_actuator[size_t(e) * m.nv + d] = fo.qfa[q]; }
      if (lane == 0) a.st.time[e] = time;
    } else {
      for (int i = lane; i < m.nq; i += 32) a.st.qpos[size_t(e) * m.nq + i] = a.st.first_qpos[size_t(e) * m.nq + i];
      for (int i = lane; i < m.nv; i += 32) {
        a.st.qvel[size_t(e) * m.nv + i] = a.st.first_qvel[size_t(e) * m.nv + i];
        a.st.qacc_warmstart[size_t(e) * m.nv + i] = a.st.first_qacc_warmstart[size_t(e) * m.nv + i];
        a.st.qfrc_actuator[size_t(e) * m.nv + i] = a.st.first_qfrc_actuator[size_t(e) * m.nv + i];
      }
      for (int i = lane; i < m.na; i += 32) a.st.act[size_t(e) * m.na + i] = a.st.first_act[size_t(e) * m.na + i];
      for (int i = lane; i < m.nbody * 3; i += 32) a.st.xpos[size_t(e) * m.nbody * 3 + i] = a.st.first_xpos[size_t(e) * m.nbody * 3 + i];
      for (int i = lane; i < m.nbody * 4; i += 32) a.st.xquat[size_t(e) * m.nbody * 4 + i] = a.st.first_xquat[size_t(e) * m.nbody * 4 + i];
      for (int i = lane; i < nobs; i += 32) obs[i] = a.st.first_obs[size_t(e) * nobs + i];
      for (int i = lane; i < nu; i += 32) a.st.prev_ctrl[size_t(e) * nu + i] = a.st.first_prev_ctrl[size_t(e) * nu + i];
      if (lane == 0) a.st.time[e] = a.st.first_time[e];
    }
    if (!kStep && (a.flags & TMJX_F_SNAPSHOT)) {
      for (int i = lane; i < m.nq; i += 32) a.st.first_qpos[size_t(e) * m.nq + i] = qpos[i];
      for (int i = lane; i < m.nv; i += 32) { a.st.first_qvel[size_t(e) * m.nv + i] = qvel[i]; a.st.first_qacc_warmstart[size_t(e) * m.nv + i] = warm[i]; }
      for (int i = lane; i < m.na; i += 32) a.st.first_act[size_t(e) * m.na + i] = act[i];
      for (int i = lane; i < m.nbody * 3; i += 32) a.st.first_xpos[size_t(e) * m.nbody * 3 + i] = w.at(m.o_xpos)[i];
      for (int i = lane; i < m.nbody * 4; i += 32) a.st.first_xquat[size_t(e) * m.nbody * 4 + i] = w.at(m.o_xquat)[i];
#pragma unroll
      for (int q = 0; q < kNvSlots; ++q) { const int d = lane + 32 * q; if (d < m.nv) a.st.first_qfrc_actuator[size_t(e) * m.nv + d] = fo.qfa[q]; }
      __syncwarp();
      for (int i = lane; i < nobs; i += 32) a.st.first_obs[size_t(e) * nobs + i] = obs[i];
      for (int i = lane; i < nu; i += 32) a.st.first_prev_ctrl[size_t(e) * nu + i] = 0.f;
      if (lane == 0) { a.st.first_time[e] = time; a.st.steps[e] = 0.f; a.st.truncation[e] = 0.f; }
    }
    }  // live
    __syncwarp();
    PT_LAP(14);
  }
}

}  // namespace

// per-variant entry points (one translation unit per residency variant; the C ABI below dispatches on envs_per_block)
#if TMJX_VARIANT >= 14
#define TMJX_WPB TMJX_VARIANT
#define TMJX_MINB 1
#elif TMJX_VARIANT == 7
#define TMJX_WPB 7
#define TMJX_MINB 2
#elif TMJX_VARIANT == 10
#define TMJX_WPB 10
#define TMJX_MINB 1
#else
#define TMJX_WPB 4
#define TMJX_MINB 3
#endif
#define TMJX_CAT2(a, b) a##b
#define TMJX_CAT(a, b) TMJX_CAT2(a, b)
#ifdef TMJX_PHASE_TIMING
extern "C" int tmjx_debug_phase_times(unsigned long long* out, int reset) {
  if (cudaMemcpyFromSymbol(out, g_pt, sizeof(g_pt)) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[64] = {0}; if (cudaMemcpyToSymbol(g_pt, z, sizeof(z)) != cudaSuccess) return -1; }
  return 0;
}
#endif
cudaError_t TMJX_CAT(variant_attr_, TMJX_VARIANT)(int dyn) {
  cudaError_t e = cudaFuncSetAttribute(tmjx_env_kernel<true, TMJX_WPB, TMJX_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(tmjx_env_kernel<false, TMJX_WPB, TMJX_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
}
cudaError_t TMJX_CAT(variant_launch_, TMJX_VARIANT)(bool step, const KArgs& a, int grid, size_t smem, cudaStream_t st) {
  // TMJX_DEBUG_WARPS = k (timing experiment only, results are INVALID): launch k of the block's warps; the environments of the
  // missing warps are skipped.  Gives the round time as a function of the number of resident warps (tools/gpu_warp_scaling.py).
  static const int dbg_warps = [] { const char* e = std::getenv("TMJX_DEBUG_WARPS"); return e ? atoi(e) : 0; }();
  const int threads = (dbg_warps > 0 && dbg_warps < TMJX_WPB ? dbg_warps : TMJX_WPB) * 32;
  if (step) tmjx_env_kernel<true, TMJX_WPB, TMJX_MINB><<<grid, threads, smem, st>>>(a);
  else tmjx_env_kernel<false, TMJX_WPB, TMJX_MINB><<<grid, threads, smem, st>>>(a);
  return cudaGetLastError();
}
}  // namespace tmjx
#else  // ------------------------------------------------------------------------------ main translation unit
// a development build may carry a subset of the variants (TMJX_BUILD_VARIANTS); a missing one fails loudly
#define TMJX_DECL_VARIANT(v) \
  cudaError_t variant_attr_##v(int dyn); \
  cudaError_t variant_launch_##v(bool step, const KArgs& a, int grid, size_t smem, cudaStream_t st);
#define TMJX_STUB_VARIANT(v) \
  static cudaError_t variant_attr_##v(int) { return cudaErrorNotSupported; } \
  static cudaError_t variant_launch_##v(bool, const KArgs&, int, size_t, cudaStream_t) { return cudaErrorNotSupported; }
#ifdef TMJX_HAVE_VARIANT_14
TMJX_DECL_VARIANT(14)
#else
TMJX_STUB_VARIANT(14)
#endif
#ifdef TMJX_HAVE_VARIANT_7
TMJX_DECL_VARIANT(7)
#else
TMJX_STUB_VARIANT(7)
#endif
#ifdef TMJX_HAVE_VARIANT_16
TMJX_DECL_VARIANT(16)
#else
TMJX_STUB_VARIANT(16)
#endif
#ifdef TMJX_HAVE_VARIANT_10
TMJX_DECL_VARIANT(10)
#else
TMJX_STUB_VARIANT(10)
#endif
#ifdef TMJX_HAVE_VARIANT_4
TMJX_DECL_VARIANT(4)
#else
TMJX_STUB_VARIANT(4)
#endif

// FP32 FMA-throughput microbenchmark (roofline denominator)
__global__ void fma_peak_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 1.000001f, c = 1e-7f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace tmjx

// =============================================================================================== C ABI
using namespace tmjx;

struct TmjxModel {
  DevModel dm;
  DevTask task;
  DevTask* d_task = nullptr;
  int* d_i32 = nullptr;
  uint16_t* d_u16 = nullptr;
  uint8_t* d_u8 = nullptr;
  float* d_f32 = nullptr;
  int device = 0, sm_count = 0, envs_per_block = 12, max_blocks_per_sm = 1;
  int envs_per_block_alt = 0;   // CG: the 16-warp residency variant, taken when it saves a lock-step round for the batch at hand
  int force_epb = 0;            // TMJX_ENVS_PER_BLOCK
  int phase_offset_ns = 0;      // TMJX_PHASE_OFFSET_US (7-warp x 2-block variant)
  size_t smem_per_env = 0;
  float* d_spill = nullptr;     // [sm_count * max_blocks_per_sm * max envs per block, spill_floats] (compact CG layout)
  TmjxTaskConfig cfg;
};
struct TmjxClips {
  float* d_table = nullptr;
  size_t bytes = 0;
  int n_clips = 0, clip_len = 0;
  int device = 0;
};

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(TMJX_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" {

int tmjx_abi_version(void) { return TMJX_ABI_VERSION; }
const char* tmjx_last_error(void) { return g_err.c_str(); }

void tmjx_model_destroy(TmjxModel* m);

int tmjx_model_create(const void* blob, size_t nbytes, const TmjxTaskConfig* cfg, int device, TmjxModel** out) {
  if (!blob || !cfg || !out) return fail(TMJX_E_ARG, "null argument");
  if (cfg->abi_version != TMJX_ABI_VERSION) return fail(TMJX_E_ARG, "TmjxTaskConfig.abi_version mismatch");
  HostTables t;
  try {
    Blob b(blob, nbytes);
    build_tables(b, *cfg, t);
  } catch (const std::exception& e) {
    const std::string msg = e.what();
    return fail(msg.find("unsupported") != std::string::npos || msg.find("only") != std::string::npos ? TMJX_E_UNSUPPORTED : TMJX_E_BLOB, msg);
  }
  CU(cudaSetDevice(device));
  auto* m = new TmjxModel();
  std::unique_ptr<TmjxModel, void (*)(TmjxModel*)> guard(m, tmjx_model_destroy);   // error returns below free what was built
  m->device = device;
  m->cfg = *cfg;
  CU(cudaMalloc(&m->d_i32, std::max<size_t>(t.i32.size(), 1) * 4));
  CU(cudaMalloc(&m->d_u16, std::max<size_t>(t.u16.size(), 1) * 2));
  CU(cudaMalloc(&m->d_u8, std::max<size_t>(t.u8.size(), 1)));
  CU(cudaMalloc(&m->d_f32, std::max<size_t>(t.f32.size(), 1) * 4));
  CU(cudaMemcpy(m->d_i32, t.i32.data(), t.i32.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(m->d_u16, t.u16.data(), t.u16.size() * 2, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(m->d_u8, t.u8.data(), t.u8.size(), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(m->d_f32, t.f32.data(), t.f32.size() * 4, cudaMemcpyHostToDevice));
  m->dm = t.dm;
  relocate(m->dm, m->d_i32, m->d_u16, m->d_u8, m->d_f32);
  build_task(m->dm, *cfg, /*n_ref_bodies=*/m->dm.nbody - 1, m->task);
  CU(cudaMalloc(&m->d_task, sizeof(DevTask)));
  CU(cudaMemcpy(m->d_task, &m->task, sizeof(DevTask), cudaMemcpyHostToDevice));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  m->sm_count = prop.multiProcessorCount;
  const size_t per_env = size_t(m->dm.smem_floats) * 4;
  m->smem_per_env = per_env;
  // One block per SM with all of its warps in lock-step (phase_sync).  CG (compact slice, 13.9 KB per environment): 14 warps per
  // block -- 194 KB of shared memory, which leaves the SM a 60 KB L1 for the model tables; a 16-warp variant (4 per scheduler, the
  // register file's limit at 128 registers) exists behind a knob.  Newton keeps a third sparse matrix: 10 warps.  Blocks of 4
  // warps are the generic fallback.
  const size_t optin = prop.sharedMemPerBlockOptin;
  const bool is_newton = m->dm.solver == TMJX_SOLVER_NEWTON;
  m->envs_per_block = (!is_newton && per_env * 14 <= optin) ? 14 : ((is_newton && per_env * 10 <= optin) ? 10 : 4);
  // measured (profiles/r2b_warp_scaling.txt): a round of 16 warps takes 1.14 x a round of 14, so the wide block only pays when it saves
  // more than one round in eight; it is opt-in (TMJX_WIDE_BLOCKS=1: launch() then takes it when it saves a round, e.g. 16384 envs)
  m->envs_per_block_alt = 0;
  if (const char* e = std::getenv("TMJX_WIDE_BLOCKS")) { if (atoi(e) && m->envs_per_block == 14 && per_env * 16 <= optin) m->envs_per_block_alt = 16; }
  if (const char* e = std::getenv("TMJX_ENVS_PER_BLOCK")) {   // tuning knob
    const int v = atoi(e);
    if (v == 4) { m->envs_per_block = 4; m->envs_per_block_alt = 0; }
    if (v == 14 && m->envs_per_block == 14) m->envs_per_block_alt = 0;
    if (v == 16 && m->envs_per_block == 14 && per_env * 16 <= optin) { m->envs_per_block = 16; m->envs_per_block_alt = 0; }
    if (v == 7 && m->envs_per_block == 14) { m->envs_per_block = 7; m->envs_per_block_alt = 0; }
  }
  if (const char* e = std::getenv("TMJX_PHASE_OFFSET_US")) m->phase_offset_ns = atoi(e) * 1000;   // tuning knob
  if (const char* e = std::getenv("TMJX_NO_GEN")) { if (atoi(e)) m->dm.use_gen = 0; }                    // tuning knob
  if (const char* e = std::getenv("TMJX_NO_SEG")) { if (atoi(e)) m->dm.use_seg = 0; }                    // tuning knob
  if (const char* e = std::getenv("TMJX_NO_DSC4")) { if (atoi(e)) m->dm.use_dsc4 = 0; }                  // tuning knob
  m->dm.sync_level = 0;  // measured: one barrier per substep keeps the block in lock-step; more only add skew
  if (const char* e = std::getenv("TMJX_SYNC")) m->dm.sync_level = atoi(e);                              // tuning knob
  // Newton: one barrier per solver iteration (the iteration re-runs the 70 KB generated factorisation; without it the
  // 10 warps drift apart inside a substep: ncu showed 28 % instruction-fetch + 28 % barrier stalls, +18 % with the barrier)
  m->dm.sync_mask = m->dm.sync_level > 0 ? 0x1f : (is_newton ? 16 : 0);
  if (const char* e = std::getenv("TMJX_SYNC_MASK")) m->dm.sync_mask = atoi(e);                          // tuning knob
  m->dm.sync_every = 1;
  if (const char* e = std::getenv("TMJX_SYNC_EVERY")) m->dm.sync_every = std::max(1, atoi(e));           // tuning knob
  if (per_env * m->envs_per_block > optin) return fail(TMJX_E_UNSUPPORTED, "model does not fit in shared memory (unsupported)");
  m->max_blocks_per_sm = m->envs_per_block == 7 ? 2 : (m->envs_per_block != 4 ? 1 : int(std::max<size_t>(1, prop.sharedMemPerMultiprocessor / (per_env * 4 + 1024))));
  for (int epb : {m->envs_per_block, m->envs_per_block_alt}) {
    if (!epb) continue;
    const int dyn = int(per_env * epb);
    CU(epb == 16 ? variant_attr_16(dyn) : (epb == 14 ? variant_attr_14(dyn) : (epb == 10 ? variant_attr_10(dyn) : (epb == 7 ? variant_attr_7(dyn) : variant_attr_4(dyn)))));
  }
  {
    const size_t slots = size_t(m->sm_count) * m->max_blocks_per_sm * std::max(m->envs_per_block, m->envs_per_block_alt);
    CU(cudaMalloc(&m->d_spill, std::max<size_t>(slots * size_t(m->dm.spill_floats), 4) * 4));
  }
  *out = guard.release();
  return TMJX_OK;
}

void tmjx_model_destroy(TmjxModel* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaFree(m->d_i32); cudaFree(m->d_u16); cudaFree(m->d_u8); cudaFree(m->d_f32); cudaFree(m->d_task); cudaFree(m->d_spill);
  delete m;
}

int tmjx_model_set_episode_length(TmjxModel* m, int episode_length) {
  if (!m || episode_length <= 0) return fail(TMJX_E_ARG, "episode_length must be positive");
  CU(cudaSetDevice(m->device));
  m->cfg.episode_length = episode_length;
  m->task.cfg.episode_length = episode_length;
  CU(cudaMemcpy(m->d_task, &m->task, sizeof(DevTask), cudaMemcpyHostToDevice));   // synchronous: ordered against every stream's later launches
  return TMJX_OK;
}

int tmjx_model_dims(const TmjxModel* m, TmjxDims* d) {
  if (!m || !d) return fail(TMJX_E_ARG, "null argument");
  d->nq = m->dm.nq; d->nv = m->dm.nv; d->nu = m->dm.nu; d->na = m->dm.na; d->nbody = m->dm.nbody; d->njnt = m->dm.njnt;
  d->ncon = m->dm.ncon; d->nefc = m->dm.nefc;
  d->obs_size = m->task.obs_size; d->reference_obs_size = m->task.ref_obs_size; d->proprioceptive_obs_size = m->task.prop_obs_size;
  d->var_window_size = m->cfg.var_window_size; d->n_metrics = TMJX_N_METRICS;
  d->smem_bytes_per_env = m->dm.smem_floats * 4; d->envs_per_block = m->envs_per_block; d->threads_per_env = 32;
  return TMJX_OK;
}

void tmjx_clips_destroy(TmjxClips* c);

int tmjx_clips_create(const TmjxModel* m, const float* position, const float* quaternion, const float* joints,
                      const float* body_positions, const float* velocity, const float* angular_velocity,
                      const float* joints_velocity, const float* body_quaternions, int n_clips, int clip_len,
                      int n_ref_bodies, TmjxClips** out) {
  (void)velocity; (void)joints_velocity; (void)body_quaternions;  // not read by reward / obs
  if (!m || !position || !quaternion || !joints || !body_positions || !angular_velocity || !out) return fail(TMJX_E_ARG, "null argument");
  if (n_clips <= 0 || clip_len < m->cfg.traj_length) return fail(TMJX_E_ARG, "bad clip table shape");
  if (n_ref_bodies != m->dm.nbody - 1) return fail(TMJX_E_ARG, "body_positions must have nbody-1 rows (stac-mjx layout without `floor`)");
  const DevTask& t = m->task;
  const int nj = m->dm.nq - 7;
  const std::vector<int> rows = task_rows(t);
  const size_t nf = size_t(n_clips) * clip_len;
  std::vector<float> tab(nf * t.frame_stride, 0.f);
  for (size_t f = 0; f < nf; ++f) {
    float* o = tab.data() + f * t.frame_stride;
    for (int k = 0; k < 3; ++k) o[t.o_pos + k] = position[f * 3 + k];
    for (int k = 0; k < 4; ++k) o[t.o_quat + k] = quaternion[f * 4 + k];
    for (int k = 0; k < 3; ++k) o[t.o_angvel + k] = angular_velocity[f * 3 + k];
    for (int k = 0; k < nj; ++k) o[t.o_joints + k] = joints[f * nj + k];
    for (int r = 0; r < t.n_rows; ++r)
      for (int k = 0; k < 3; ++k) o[t.o_bodies + 3 * r + k] = body_positions[(f * n_ref_bodies + rows[r]) * 3 + k];
  }
  CU(cudaSetDevice(m->device));
  auto* c = new TmjxClips();
  std::unique_ptr<TmjxClips, void (*)(TmjxClips*)> cguard(c, tmjx_clips_destroy);
  c->device = m->device; c->n_clips = n_clips; c->clip_len = clip_len; c->bytes = tab.size() * 4;
  CU(cudaMalloc(&c->d_table, c->bytes));
  CU(cudaMemcpy(c->d_table, tab.data(), c->bytes, cudaMemcpyHostToDevice));
  *out = cguard.release();
  return TMJX_OK;
}
void tmjx_clips_destroy(TmjxClips* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->d_table);
  delete c;
}
size_t tmjx_clips_device_bytes(const TmjxClips* c) { return c ? c->bytes : 0; }

}  // extern "C"

static int check_common(const TmjxModel* m, const TmjxClips* c, const TmjxState* s, const TmjxOut* o, int n_env) {
  if (!m || !c || !s || !o) return fail(TMJX_E_ARG, "null argument");
  if (n_env <= 0) return fail(TMJX_E_ARG, "n_env must be positive");
  if (!s->qpos || !s->qvel || !s->act || !s->time || !s->qacc_warmstart || !s->xpos || !s->xquat || !s->qfrc_actuator ||
      !s->clip_idx || !s->start_frame || !s->buffer_index || !s->prev_ctrl || !s->action_buffer)
    return fail(TMJX_E_ARG, "TmjxState has a null required buffer");
  if (!o->obs || !o->reward || !o->done || !o->metrics || !o->cur_frame) return fail(TMJX_E_ARG, "TmjxOut has a null required buffer");
  if (o->dbg_qM) return fail(TMJX_E_UNSUPPORTED, "dbg_qM is not produced by the CUDA path (the inertia is never dense) (unsupported)");
  return TMJX_OK;
}
static bool has_first(const TmjxState* s) {
  return s->steps && s->truncation && s->first_qpos && s->first_qvel && s->first_act && s->first_time && s->first_qacc_warmstart &&
         s->first_xpos && s->first_xquat && s->first_qfrc_actuator && s->first_obs && s->first_prev_ctrl;
}
template <bool kStep>
static int launch(const TmjxModel* m, const TmjxClips* c, const float* action, TmjxState* s, TmjxOut* o, int n_env, unsigned flags,
                  void* stream) {
  KArgs a;
  a.m = m->dm; a.task = m->d_task; a.clips = c->d_table; a.n_clips = c->n_clips; a.clip_len = c->clip_len;
  a.st = *s; a.out = *o; a.action = action; a.n_env = n_env; a.flags = flags;
  a.spill = m->d_spill;
  a.phase_offset_ns = m->envs_per_block == 7 ? m->phase_offset_ns : 0;
  // residency variant: the block runs ceil(envs of the block / envs per block) lock-step rounds of the same duration whatever the
  // number of warps, so the wider block is taken exactly when it saves a round (4096 envs on 148 SMs: 28 per SM = 2 rounds either
  // way -> 14 warps and the larger L1; 16384: 111 per SM = 8 rounds of 14 or 7 of 16)
  int epb = m->envs_per_block;
  if (m->envs_per_block_alt) {
    const int per_sm = (n_env + m->sm_count - 1) / m->sm_count, alt = m->envs_per_block_alt;
    if ((per_sm + alt - 1) / alt < (per_sm + epb - 1) / epb) epb = alt;
  }
  const int need = (n_env + epb - 1) / epb;
  const int grid = std::min(need, m->sm_count * m->max_blocks_per_sm);
  const size_t smem = m->smem_per_env * epb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(epb == 16 ? variant_launch_16(kStep, a, grid, smem, st)
               : (epb == 14 ? variant_launch_14(kStep, a, grid, smem, st)
                            : (epb == 10 ? variant_launch_10(kStep, a, grid, smem, st)
                                         : (epb == 7 ? variant_launch_7(kStep, a, grid, smem, st) : variant_launch_4(kStep, a, grid, smem, st)))));
  return TMJX_OK;
}

extern "C" {

int tmjx_forward(const TmjxModel* m, const TmjxClips* c, TmjxState* s, TmjxOut* o, int n_env, unsigned flags, void* stream) {
  int rc = check_common(m, c, s, o, n_env);
  if (rc) return rc;
  if ((flags & TMJX_F_SNAPSHOT) && !has_first(s)) return fail(TMJX_E_ARG, "TMJX_F_SNAPSHOT needs the first_* / steps / truncation buffers");
  return launch<false>(m, c, nullptr, s, o, n_env, flags, stream);
}

int tmjx_step(const TmjxModel* m, const TmjxClips* c, const float* action, TmjxState* s, TmjxOut* o, int n_env, unsigned flags,
              void* stream) {
  int rc = check_common(m, c, s, o, n_env);
  if (rc) return rc;
  if (!action) return fail(TMJX_E_ARG, "null action");
  if ((flags & TMJX_F_AUTORESET) && !has_first(s)) return fail(TMJX_E_ARG, "TMJX_F_AUTORESET needs the first_* / steps / truncation buffers");
  return launch<true>(m, c, action, s, o, n_env, flags, stream);
}

double tmjx_fp32_peak_tflops(int device, void* stream) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  float* d = nullptr;
  if (cudaMalloc(&d, size_t(blocks) * threads * 4) != cudaSuccess) return -1.0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fma_peak_kernel<<<blocks, threads, 0, st>>>(d, 1024);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    fma_peak_kernel<<<blocks, threads, 0, st>>>(d, iters);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * double(iters) * blocks * threads;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

}  // extern "C"
#endif  // TMJX_VARIANT
