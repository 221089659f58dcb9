/*
 * tmjx_oracle.cpp — CPU ORACLE (test infrastructure, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product path (track-mjx_b200/, csrc/) never links or calls it.
 *
 * PARITY UNPINNED: the reference (talmolab/track-mjx v0.0.2) ships no tests, golden vectors or fixtures for
 * the env path, and its arithmetic lives in third-party packages that are neither vendored under
 * /root/reference nor installable in this image (mujoco-mjx==3.3.2, mujoco==3.3.2, brax==0.12.3,
 * jax==0.6.2; reference pyproject.toml:17-27).  This file restates, from the published algorithms,
 *   - mujoco.mjx `forward` / `step` (mjx/_src/{forward,smooth,collision_primitive,constraint,solver,
 *     passive,support,math}.py @3.3.2) as exercised by brax `PipelineEnv.pipeline_init/pipeline_step`
 *     (reference call sites track_mjx/environment/task/single_clip_tracking.py:163 and :219), DENSE like
 *     MJX runs it (`opt.jacobian = 0`, single_clip_tracking.py:72): dense qM, dense Cholesky, dense efc_J;
 *   - the reference's own task code line by line: SingleClipTracking.step / reset_from_clip / _get_obs /
 *     _get_cur_frame (single_clip_tracking.py:121-454), compute_tracking_rewards (reward.py:57-485),
 *     BaseWalker.compute_local_* (walker/base.py:170-258), brax EpisodeWrapper + the auto-reset wrapper
 *     (wrappers.py:104-144, 288-310).
 * It is validated by physics invariants and analytic cases in tests/ (see DESIGN.md "Oracle").
 *
 * Templated on the scalar type: T=float mirrors the reference's fp32 arithmetic (evaluation order differs
 * from XLA's), T=double is the noise-free reference used to calibrate tolerances.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../include/tmjx.h"
#include "../include/tmjx_blob.h"

namespace {

constexpr double kMinVal = 1e-15;   // mjMINVAL
constexpr double kMinImp = 0.0001;  // mjMINIMP
constexpr double kMaxImp = 0.9999;  // mjMAXIMP
enum { kGeomSphere = 2, kGeomCapsule = 3, kGeomEllipsoid = 4 };
enum { kJntFree = 0, kJntHinge = 3 };

thread_local std::string g_err;

// ------------------------------------------------------------------ small math (mjx/_src/math.py)
// jnp.minimum / jnp.maximum propagate NaN (std::min / fminf do not)
template <class T> inline T jmin(T a, T b) { return (a != a || b != b) ? std::numeric_limits<T>::quiet_NaN() : (a < b ? a : b); }
template <class T> inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline void cross3(const T* a, const T* b, T* o) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
// math.rotate: 2(u.v)u + (s^2 - u.u)v + 2s(u x v)
template <class T> inline void rotate(const T* v, const T* q, T* o) {
  const T s = q[0];
  const T* u = q + 1;
  T c[3];
  cross3(u, v, c);
  const T uv = dot3(u, v), uu = dot3(u, u);
  T r[3];
  for (int i = 0; i < 3; ++i) r[i] = T(2) * (uv * u[i]) + (s * s - uu) * v[i] + T(2) * s * c[i];
  o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}
template <class T> inline void quat_mul(const T* a, const T* b, T* o) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
template <class T> inline void quat_to_mat(const T* q, T* m) {
  const T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = T(2) * (x * y - w * z); m[2] = T(2) * (x * z + w * y);
  m[3] = T(2) * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = T(2) * (y * z - w * x);
  m[6] = T(2) * (x * z - w * y); m[7] = T(2) * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
template <class T> inline void axis_angle_to_quat(const T* axis, T angle, T* q) {
  const T s = std::sin(angle * T(0.5)), c = std::cos(angle * T(0.5));
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// math.normalize_with_norm: x / (n + 1e-6 * (n == 0))
template <class T> inline T normalize(T* x, int n) {
  T s = 0;
  for (int i = 0; i < n; ++i) s += x[i] * x[i];
  const T nrm = std::sqrt(s);
  const T d = nrm + (nrm == T(0) ? T(1e-6) : T(0));
  for (int i = 0; i < n; ++i) x[i] = x[i] / d;
  return nrm;
}
// math.inert_mul: cinert 10-vector [Ixx Iyy Izz Ixy Ixz Iyz | m*off(3) | m] times motion vector [ang; lin]
template <class T> inline void inert_mul(const T* i, const T* v, T* o) {
  const T* pos = i + 6;
  const T mass = i[9];
  T c1[3], c2[3];
  cross3(pos, v + 3, c1);
  cross3(pos, v, c2);
  T r[6];
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c1[0];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c1[1];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c1[2];
  for (int k = 0; k < 3; ++k) r[3 + k] = mass * v[3 + k] - c2[k];
  for (int k = 0; k < 6; ++k) o[k] = r[k];
}
template <class T> inline void motion_cross(const T* u, const T* v, T* o) {
  T a[3], b[3], c[3];
  cross3(u, v, a);
  cross3(u + 3, v, b);
  cross3(u, v + 3, c);
  for (int k = 0; k < 3; ++k) { o[k] = a[k]; o[3 + k] = b[k] + c[k]; }
}
template <class T> inline void motion_cross_force(const T* v, const T* f, T* o) {
  T a[3], b[3], c[3];
  cross3(v, f, a);
  cross3(v + 3, f + 3, b);
  cross3(v, f + 3, c);
  for (int k = 0; k < 3; ++k) { o[k] = a[k] + b[k]; o[3 + k] = c[k]; }
}
// math.make_frame / orthogonals for a unit normal
template <class T> inline void make_frame(const T* n_in, T* fr) {
  T a[3] = {n_in[0], n_in[1], n_in[2]};
  normalize(a, 3);
  T b[3] = {0, 0, 0};
  if (T(-0.5) < a[1] && a[1] < T(0.5)) b[1] = 1; else b[2] = 1;
  const T ab = dot3(a, b);
  for (int k = 0; k < 3; ++k) b[k] -= a[k] * ab;
  normalize(b, 3);
  for (int k = 0; k < 3; ++k) { fr[k] = a[k]; fr[3 + k] = b[k]; }
  cross3(a, b, fr + 6);
}

// ------------------------------------------------------------------ model
template <class T> std::vector<T> conv(const std::vector<float>& v) { return std::vector<T>(v.begin(), v.end()); }

template <class T> struct Model {
  int nq, nv, nu, na, nbody, njnt, ncgeom, ncon, nefc, ntendon, npair, nlimit;
  T timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  std::vector<int32_t> body_parentid, body_rootid, body_jntadr, body_jntnum, body_dofadr, body_dofnum, jnt_type,
      jnt_qposadr, jnt_dofadr, jnt_bodyid, dof_bodyid, dof_jntid, dof_parentid, jnt_limited, actuator_ctrllimited,
      actuator_forcelimited, actuator_bias_affine, actuator_dyn_filter, cgeom_type, cgeom_bodyid, pair_cgeom;
  std::vector<T> body_pos, body_quat, body_ipos, body_iquat, body_mass, body_inertia, body_invweight0, jnt_pos,
      jnt_axis, jnt_range, jnt_stiffness, jnt_margin, jnt_solref, jnt_solimp, qpos0, qpos_spring, dof_armature,
      dof_damping, dof_invweight0, actuator_moment, actuator_gain, actuator_biasprm, actuator_dynprm,
      actuator_ctrlrange, actuator_forcerange, cgeom_pos, cgeom_quat, cgeom_size, plane, pair_friction, pair_solref,
      pair_solimp, pair_includemargin;
  int plane_bodyid;
  std::vector<int> limit_jnt;  // limited joints in id order -> efc rows [0, nlimit)

  void load(const tmjx::Blob& b) {
    auto d = b.i32("dims");
    nq = d[0]; nv = d[1]; nu = d[2]; na = d[3]; nbody = d[4]; njnt = d[5]; ncgeom = d[6]; ncon = d[7]; nefc = d[8];
    ntendon = d[9];
    auto o = b.f32("opt");
    timestep = o[0]; gravity[0] = o[1]; gravity[1] = o[2]; gravity[2] = o[3]; tolerance = o[4]; ls_tolerance = o[5];
    impratio = o[6]; meaninertia = o[7];
#define LI(x) x = b.i32(#x)
#define LF(x) x = conv<T>(b.f32(#x))
    LI(body_parentid); LI(body_rootid); LI(body_jntadr); LI(body_jntnum); LI(body_dofadr); LI(body_dofnum);
    LI(jnt_type); LI(jnt_qposadr); LI(jnt_dofadr); LI(jnt_bodyid); LI(dof_bodyid); LI(dof_jntid); LI(dof_parentid);
    LI(jnt_limited); LI(actuator_ctrllimited); LI(actuator_forcelimited); LI(actuator_bias_affine);
    LI(actuator_dyn_filter); LI(cgeom_type); LI(cgeom_bodyid); LI(pair_cgeom);
    LF(body_pos); LF(body_quat); LF(body_ipos); LF(body_iquat); LF(body_mass); LF(body_inertia); LF(body_invweight0);
    LF(jnt_pos); LF(jnt_axis); LF(jnt_range); LF(jnt_stiffness); LF(jnt_margin); LF(jnt_solref); LF(jnt_solimp);
    LF(qpos0); LF(qpos_spring); LF(dof_armature); LF(dof_damping); LF(dof_invweight0); LF(actuator_moment);
    LF(actuator_gain); LF(actuator_biasprm); LF(actuator_dynprm); LF(actuator_ctrlrange); LF(actuator_forcerange);
    LF(cgeom_pos); LF(cgeom_quat); LF(cgeom_size); LF(plane); LF(pair_friction); LF(pair_solref); LF(pair_solimp);
    LF(pair_includemargin);
#undef LI
#undef LF
    plane_bodyid = b.i32("plane_bodyid")[0];
    npair = int(pair_cgeom.size());
    for (int j = 0; j < njnt; ++j)
      if (jnt_limited[j] && jnt_type[j] == kJntHinge) limit_jnt.push_back(j);
    nlimit = int(limit_jnt.size());
    if (nlimit + 4 * ncon != nefc) throw std::runtime_error("oracle: nefc mismatch");
  }
};

struct Clips {
  std::vector<float> position, quaternion, joints, body_positions, angular_velocity;
  int n_clips = 0, clip_len = 0, n_ref_bodies = 0, n_joints = 0;
};

// ------------------------------------------------------------------ per-env workspace (the slice of mjx.Data used)
template <class T> struct Data {
  std::vector<T> qpos, qvel, act, ctrl, warm;
  T time = 0;
  std::vector<T> xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, subtree_com, cinert, cdof, crb, qM, qL, cvel,
      cdof_dot, qfrc_bias, qfrc_passive, qfrc_actuator, act_dot, qfrc_smooth, qacc_smooth, qacc, qfrc_constraint,
      con_dist, con_pos, con_frame, efc_J, efc_D, efc_aref, efc_force;
  std::vector<int> con_pair;
  explicit Data(const Model<T>& m) {
    qpos.resize(m.nq); qvel.resize(m.nv); act.resize(m.na); ctrl.resize(m.nu); warm.resize(m.nv);
    xpos.resize(m.nbody * 3); xquat.resize(m.nbody * 4); xmat.resize(m.nbody * 9); xipos.resize(m.nbody * 3);
    ximat.resize(m.nbody * 9); xanchor.resize(m.njnt * 3); xaxis.resize(m.njnt * 3);
    subtree_com.resize(m.nbody * 3); cinert.resize(m.nbody * 10); cdof.resize(m.nv * 6); crb.resize(m.nbody * 10);
    qM.resize(m.nv * m.nv); qL.resize(m.nv * m.nv); cvel.resize(m.nbody * 6); cdof_dot.resize(m.nv * 6);
    qfrc_bias.resize(m.nv); qfrc_passive.resize(m.nv); qfrc_actuator.resize(m.nv); act_dot.resize(m.na);
    qfrc_smooth.resize(m.nv); qacc_smooth.resize(m.nv); qacc.resize(m.nv); qfrc_constraint.resize(m.nv);
    con_dist.resize(m.ncon); con_pos.resize(m.ncon * 3); con_frame.resize(m.ncon * 9); con_pair.resize(m.ncon);
    efc_J.resize(size_t(m.nefc) * m.nv); efc_D.resize(m.nefc); efc_aref.resize(m.nefc); efc_force.resize(m.nefc);
  }
};

// dense Cholesky A = L L^T (lower), jax.scipy.linalg.cho_factor equivalent
template <class T> void cholesky(const T* a, T* l, int n) {
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j <= i; ++j) {
      T s = a[i * n + j];
      for (int k = 0; k < j; ++k) s -= l[i * n + k] * l[j * n + k];
      if (i == j) l[i * n + i] = std::sqrt(s);
      else l[i * n + j] = s / l[j * n + j];
    }
    for (int j = i + 1; j < n; ++j) l[i * n + j] = 0;
  }
}
template <class T> void cho_solve(const T* l, const T* b, T* x, int n) {
  for (int i = 0; i < n; ++i) {
    T s = b[i];
    for (int k = 0; k < i; ++k) s -= l[i * n + k] * x[k];
    x[i] = s / l[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    T s = x[i];
    for (int k = i + 1; k < n; ++k) s -= l[k * n + i] * x[k];
    x[i] = s / l[i * n + i];
  }
}
template <class T> void matvec(const T* a, const T* x, T* y, int rows, int cols) {
  for (int i = 0; i < rows; ++i) {
    T s = 0;
    for (int k = 0; k < cols; ++k) s += a[size_t(i) * cols + k] * x[k];
    y[i] = s;
  }
}
template <class T> T dotn(const T* a, const T* b, int n) {
  T s = 0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// ------------------------------------------------------------------ the simulator
template <class T> struct Sim {
  Model<T> m;
  TmjxTaskConfig cfg;
  T dt;

  // ---- smooth.kinematics
  void kinematics(Data<T>& d) const {
    for (int k = 0; k < 3; ++k) d.xpos[k] = 0;
    d.xquat[0] = 1; d.xquat[1] = d.xquat[2] = d.xquat[3] = 0;
    quat_to_mat(&d.xquat[0], &d.xmat[0]);
    for (int b = 1; b < m.nbody; ++b) {
      const int p = m.body_parentid[b];
      T pos[3], quat[4], t[3];
      rotate(&m.body_pos[b * 3], &d.xquat[p * 4], t);
      for (int k = 0; k < 3; ++k) pos[k] = d.xpos[p * 3 + k] + t[k];
      quat_mul(&d.xquat[p * 4], &m.body_quat[b * 4], quat);
      for (int jj = 0; jj < m.body_jntnum[b]; ++jj) {
        const int j = m.body_jntadr[b] + jj, qa = m.jnt_qposadr[j];
        if (m.jnt_type[j] == kJntFree) {
          for (int k = 0; k < 3; ++k) { d.xanchor[j * 3 + k] = d.qpos[qa + k]; d.xaxis[j * 3 + k] = (k == 2); }
          for (int k = 0; k < 3; ++k) pos[k] = d.qpos[qa + k];
          for (int k = 0; k < 4; ++k) quat[k] = d.qpos[qa + 3 + k];
          normalize(quat, 4);
          for (int k = 0; k < 4; ++k) d.qpos[qa + 3 + k] = quat[k];  // MJX writes the normalised quat back
        } else {
          rotate(&m.jnt_pos[j * 3], quat, t);
          for (int k = 0; k < 3; ++k) d.xanchor[j * 3 + k] = t[k] + pos[k];
          rotate(&m.jnt_axis[j * 3], quat, &d.xaxis[j * 3]);
          T qloc[4], q2[4];
          axis_angle_to_quat(&m.jnt_axis[j * 3], d.qpos[qa] - m.qpos0[qa], qloc);
          quat_mul(quat, qloc, q2);
          for (int k = 0; k < 4; ++k) quat[k] = q2[k];
          rotate(&m.jnt_pos[j * 3], quat, t);
          for (int k = 0; k < 3; ++k) pos[k] = d.xanchor[j * 3 + k] - t[k];
        }
      }
      for (int k = 0; k < 3; ++k) d.xpos[b * 3 + k] = pos[k];
      for (int k = 0; k < 4; ++k) d.xquat[b * 4 + k] = quat[k];
      quat_to_mat(quat, &d.xmat[b * 9]);
    }
    for (int b = 0; b < m.nbody; ++b) {  // support.local_to_global for the inertial frames
      T t[3], q[4];
      rotate(&m.body_ipos[b * 3], &d.xquat[b * 4], t);
      for (int k = 0; k < 3; ++k) d.xipos[b * 3 + k] = d.xpos[b * 3 + k] + t[k];
      quat_mul(&d.xquat[b * 4], &m.body_iquat[b * 4], q);
      quat_to_mat(q, &d.ximat[b * 9]);
    }
  }

  // ---- smooth.com_pos
  void com_pos(Data<T>& d) const {
    std::vector<T> pos(m.nbody * 3), mass(m.nbody);
    for (int b = 0; b < m.nbody; ++b) {
      mass[b] = m.body_mass[b];
      for (int k = 0; k < 3; ++k) pos[b * 3 + k] = d.xipos[b * 3 + k] * m.body_mass[b];
    }
    for (int b = m.nbody - 1; b > 0; --b) {
      const int p = m.body_parentid[b];
      mass[p] += mass[b];
      for (int k = 0; k < 3; ++k) pos[p * 3 + k] += pos[b * 3 + k];
    }
    for (int b = 0; b < m.nbody; ++b)
      for (int k = 0; k < 3; ++k)
        d.subtree_com[b * 3 + k] = (mass[b] < T(kMinVal)) ? d.xipos[b * 3 + k] : pos[b * 3 + k] / mass[b];
    for (int b = 0; b < m.nbody; ++b) {
      const T* root = &d.subtree_com[m.body_rootid[b] * 3];
      T off[3];
      for (int k = 0; k < 3; ++k) off[k] = d.xipos[b * 3 + k] - root[k];
      const T* R = &d.ximat[b * 9];
      const T* I = &m.body_inertia[b * 3];
      const T ms = m.body_mass[b];
      T in[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          T s = 0;
          for (int k = 0; k < 3; ++k) s += R[r * 3 + k] * I[k] * R[c * 3 + k];
          in[r * 3 + c] = s;
        }
      const T oo = dot3(off, off);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) in[r * 3 + c] += ((r == c ? oo : T(0)) - off[r] * off[c]) * ms;
      T* ci = &d.cinert[b * 10];
      ci[0] = in[0]; ci[1] = in[4]; ci[2] = in[8]; ci[3] = in[1]; ci[4] = in[2]; ci[5] = in[5];
      for (int k = 0; k < 3; ++k) ci[6 + k] = off[k] * ms;
      ci[9] = ms;
    }
    for (int j = 0; j < m.njnt; ++j) {  // cdof = [axis ; axis x (root_com - anchor)]
      const int b = m.jnt_bodyid[j], dofadr = m.jnt_dofadr[j];
      const T* root = &d.subtree_com[m.body_rootid[b] * 3];
      T off[3];
      for (int k = 0; k < 3; ++k) off[k] = root[k] - d.xanchor[j * 3 + k];
      if (m.jnt_type[j] == kJntFree) {
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 6; ++c) d.cdof[(dofadr + r) * 6 + c] = (c == 3 + r) ? T(1) : T(0);
        for (int r = 0; r < 3; ++r) {
          T a[3] = {d.xmat[b * 9 + 0 * 3 + r], d.xmat[b * 9 + 1 * 3 + r], d.xmat[b * 9 + 2 * 3 + r]};  // column r
          T* cd = &d.cdof[(dofadr + 3 + r) * 6];
          for (int k = 0; k < 3; ++k) cd[k] = a[k];
          cross3(a, off, cd + 3);
        }
      } else {
        T* cd = &d.cdof[dofadr * 6];
        for (int k = 0; k < 3; ++k) cd[k] = d.xaxis[j * 3 + k];
        cross3(&d.xaxis[j * 3], off, cd + 3);
      }
    }
  }

  // ---- smooth.crb + support.make_m (dense) + smooth.factor_m (dense Cholesky)
  void crb_and_factor(Data<T>& d) const {
    d.crb = d.cinert;
    for (int b = m.nbody - 1; b > 0; --b) {
      const int p = m.body_parentid[b];
      for (int k = 0; k < 10; ++k) d.crb[p * 10 + k] += d.crb[b * 10 + k];
    }
    for (int k = 0; k < 10; ++k) d.crb[k] = 0;
    const int nv = m.nv;
    std::fill(d.qM.begin(), d.qM.end(), T(0));
    for (int i = 0; i < nv; ++i) {
      T f[6];
      inert_mul(&d.crb[m.dof_bodyid[i] * 10], &d.cdof[i * 6], f);
      for (int j = i; j >= 0; j = m.dof_parentid[j]) {
        T s = 0;
        for (int k = 0; k < 6; ++k) s += f[k] * d.cdof[j * 6 + k];
        d.qM[i * nv + j] = s;
        d.qM[j * nv + i] = s;
      }
      d.qM[i * nv + i] += m.dof_armature[i];
    }
    cholesky(d.qM.data(), d.qL.data(), nv);
  }

  // ---- collision_driver.collision with collision_primitive.{plane_capsule, plane_ellipsoid, plane_sphere}
  void collision(Data<T>& d) const {
    const T* ppos = &m.plane[0];
    const T* n = &m.plane[3];
    int c = 0;
    for (int p = 0; p < m.npair; ++p) {
      const int g = m.pair_cgeom[p], b = m.cgeom_bodyid[g], type = m.cgeom_type[g];
      T gpos[3], gq[4], gmat[9], t[3];
      rotate(&m.cgeom_pos[g * 3], &d.xquat[b * 4], t);
      for (int k = 0; k < 3; ++k) gpos[k] = d.xpos[b * 3 + k] + t[k];
      quat_mul(&d.xquat[b * 4], &m.cgeom_quat[g * 4], gq);
      quat_to_mat(gq, gmat);
      const T* size = &m.cgeom_size[g * 3];
      auto plane_sphere = [&](const T* sp, T r, T* dist, T* pos) {
        T dlt[3] = {sp[0] - ppos[0], sp[1] - ppos[1], sp[2] - ppos[2]};
        *dist = dot3(dlt, n) - r;
        for (int k = 0; k < 3; ++k) pos[k] = sp[k] - n[k] * (r + T(0.5) * *dist);
      };
      if (type == kGeomSphere) {
        plane_sphere(gpos, size[0], &d.con_dist[c], &d.con_pos[c * 3]);
        make_frame(n, &d.con_frame[c * 9]);
        d.con_pair[c++] = p;
      } else if (type == kGeomCapsule) {
        T axis[3] = {gmat[2], gmat[5], gmat[8]};
        const T na = dot3(n, axis);
        T bb[3] = {axis[0] - n[0] * na, axis[1] - n[1] * na, axis[2] - n[2] * na};
        const T bn = normalize(bb, 3);
        if (bn < T(0.5)) {
          const bool usey = (T(-0.5) < n[1]) && (n[1] < T(0.5));
          bb[0] = 0; bb[1] = usey ? T(1) : T(0); bb[2] = usey ? T(0) : T(1);
        }
        T fr[9];
        for (int k = 0; k < 3; ++k) { fr[k] = n[k]; fr[3 + k] = bb[k]; }
        cross3(n, bb, fr + 6);
        for (int side = 0; side < 2; ++side) {
          const T sgn = side == 0 ? T(1) : T(-1);
          T sp[3];
          for (int k = 0; k < 3; ++k) sp[k] = gpos[k] + sgn * axis[k] * size[1];
          plane_sphere(sp, size[0], &d.con_dist[c], &d.con_pos[c * 3]);
          for (int k = 0; k < 9; ++k) d.con_frame[c * 9 + k] = fr[k];
          d.con_pair[c++] = p;
        }
      } else {  // ellipsoid
        T sv[3];
        for (int k = 0; k < 3; ++k) sv[k] = (gmat[0 * 3 + k] * n[0] + gmat[1 * 3 + k] * n[1] + gmat[2 * 3 + k] * n[2]) * size[k];
        normalize(sv, 3);
        for (int k = 0; k < 3; ++k) sv[k] = -sv[k] * size[k];
        T pos[3];
        for (int r = 0; r < 3; ++r) pos[r] = gpos[r] + gmat[r * 3] * sv[0] + gmat[r * 3 + 1] * sv[1] + gmat[r * 3 + 2] * sv[2];
        T dlt[3] = {pos[0] - ppos[0], pos[1] - ppos[1], pos[2] - ppos[2]};
        const T dist = dot3(n, dlt);
        d.con_dist[c] = dist;
        for (int k = 0; k < 3; ++k) d.con_pos[c * 3 + k] = pos[k] - n[k] * dist * T(0.5);
        make_frame(n, &d.con_frame[c * 9]);
        d.con_pair[c++] = p;
      }
    }
  }

  // constraint._kbi
  void kbi(const T* solref, const T* solimp, T pos, T* k, T* b, T* imp) const {
    T timeconst = solref[0], dampratio = solref[1];
    timeconst = std::max(timeconst, T(2) * dt);  // refsafe
    T dmin = std::min(std::max(solimp[0], T(kMinImp)), T(kMaxImp));
    T dmax = std::min(std::max(solimp[1], T(kMinImp)), T(kMaxImp));
    T width = std::max(T(kMinVal), solimp[2]);
    T mid = std::min(std::max(solimp[3], T(kMinImp)), T(kMaxImp));
    T power = std::max(T(1), solimp[4]);
    *k = T(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
    *b = T(2) / (dmax * timeconst);
    if (solref[0] <= 0) *k = -solref[0] / (dmax * dmax);
    if (solref[1] <= 0) *b = -solref[1] / dmax;
    const T x = std::abs(pos) / width;
    const T a_ = (T(1) / std::pow(mid, power - 1)) * std::pow(x, power);
    const T b_ = T(1) - (T(1) / std::pow(T(1) - mid, power - 1)) * std::pow(T(1) - x, power);
    const T y = x < mid ? a_ : b_;
    T im = dmin + y * (dmax - dmin);
    im = std::min(std::max(im, dmin), dmax);
    if (x > T(1)) im = dmax;
    *imp = im;
  }

  // ---- constraint.make_constraint: limits (hinge) then pyramidal contacts; dense efc_J
  void make_constraint(Data<T>& d) const {
    const int nv = m.nv;
    std::fill(d.efc_J.begin(), d.efc_J.end(), T(0));
    std::vector<T> pos(m.nefc), invw(m.nefc), solref(m.nefc * 2), solimp(m.nefc * 5);
    for (int r = 0; r < m.nlimit; ++r) {
      const int j = m.limit_jnt[r];
      const T q = d.qpos[m.jnt_qposadr[j]];
      const T dmin = q - m.jnt_range[j * 2], dmax = m.jnt_range[j * 2 + 1] - q;
      pos[r] = std::min(dmin, dmax) - m.jnt_margin[j];
      const bool active = pos[r] < 0;
      d.efc_J[size_t(r) * nv + m.jnt_dofadr[j]] = active ? (dmin < dmax ? T(1) : T(-1)) : T(0);
      invw[r] = m.dof_invweight0[m.jnt_dofadr[j]];
      for (int k = 0; k < 2; ++k) solref[r * 2 + k] = m.jnt_solref[j * 2 + k];
      for (int k = 0; k < 5; ++k) solimp[r * 5 + k] = m.jnt_solimp[j * 5 + k];
    }
    std::vector<T> jacp(nv * 3);
    for (int c = 0; c < m.ncon; ++c) {
      const int p = d.con_pair[c], body2 = m.cgeom_bodyid[m.pair_cgeom[p]], body1 = m.plane_bodyid;
      const T dist = d.con_dist[c] - m.pair_includemargin[p];
      const bool active = dist < 0;
      // support.jac_dif_pair: jacp(body2) - jacp(body1); the plane's body is static => jacp(body1) = 0
      std::fill(jacp.begin(), jacp.end(), T(0));
      T off[3];
      for (int k = 0; k < 3; ++k) off[k] = d.con_pos[c * 3 + k] - d.subtree_com[m.body_rootid[body2] * 3 + k];
      int b = body2;
      while (b > 0 && m.body_dofnum[b] == 0) b = m.body_parentid[b];
      if (b > 0)
        for (int i = m.body_dofadr[b] + m.body_dofnum[b] - 1; i >= 0; i = m.dof_parentid[i]) {
          T cr[3];
          cross3(&d.cdof[i * 6], off, cr);
          for (int k = 0; k < 3; ++k) jacp[i * 3 + k] = d.cdof[i * 6 + 3 + k] + cr[k];
        }
      const T t = m.body_invweight0[body1 * 2] + m.body_invweight0[body2 * 2];
      const T* fr = &d.con_frame[c * 9];
      const T* fric = &m.pair_friction[p * 5];
      for (int k = 0; k < 4; ++k) {
        const int r = m.nlimit + c * 4 + k;
        const int tan = 1 + k / 2;
        const T mu = fric[tan - 1], sgn = (k % 2 == 0) ? T(1) : T(-1);
        if (active)
          for (int i = 0; i < nv; ++i) {
            const T jn = dot3(fr, &jacp[i * 3]), jt = dot3(fr + 3 * tan, &jacp[i * 3]);
            d.efc_J[size_t(r) * nv + i] = jn + jt * mu * sgn;
          }
        pos[r] = dist;
        invw[r] = (t + mu * mu * t) * T(2) * mu * mu / m.impratio;
        for (int q = 0; q < 2; ++q) solref[r * 2 + q] = m.pair_solref[p * 2 + q];
        for (int q = 0; q < 5; ++q) solimp[r * 5 + q] = m.pair_solimp[p * 5 + q];
      }
    }
    for (int r = 0; r < m.nefc; ++r) {
      T k, b, imp;
      kbi(&solref[r * 2], &solimp[r * 5], pos[r], &k, &b, &imp);
      const T R = std::max(invw[r] * (T(1) - imp) / imp, T(kMinVal));
      d.efc_D[r] = T(1) / R;
      d.efc_aref[r] = -b * dotn(&d.efc_J[size_t(r) * nv], d.qvel.data(), nv) - k * imp * pos[r];
    }
  }

  // ---- smooth.com_vel
  void com_vel(Data<T>& d) const {
    for (int k = 0; k < 6; ++k) d.cvel[k] = 0;
    for (int b = 1; b < m.nbody; ++b) {
      T cv[6];
      for (int k = 0; k < 6; ++k) cv[k] = d.cvel[m.body_parentid[b] * 6 + k];
      for (int jj = 0; jj < m.body_jntnum[b]; ++jj) {
        const int j = m.body_jntadr[b] + jj, da = m.jnt_dofadr[j];
        if (m.jnt_type[j] == kJntFree) {
          for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 6; ++k) { cv[k] += d.cdof[(da + r) * 6 + k] * d.qvel[da + r]; d.cdof_dot[(da + r) * 6 + k] = 0; }
          for (int r = 3; r < 6; ++r) motion_cross(cv, &d.cdof[(da + r) * 6], &d.cdof_dot[(da + r) * 6]);
          for (int r = 3; r < 6; ++r)
            for (int k = 0; k < 6; ++k) cv[k] += d.cdof[(da + r) * 6 + k] * d.qvel[da + r];
        } else {
          motion_cross(cv, &d.cdof[da * 6], &d.cdof_dot[da * 6]);
          for (int k = 0; k < 6; ++k) cv[k] += d.cdof[da * 6 + k] * d.qvel[da];
        }
      }
      for (int k = 0; k < 6; ++k) d.cvel[b * 6 + k] = cv[k];
    }
  }

  // ---- passive.passive (joint springs + dampers) and smooth.rne
  void passive_and_rne(Data<T>& d) const {
    for (int i = 0; i < m.nv; ++i) d.qfrc_passive[i] = -m.dof_damping[i] * d.qvel[i];
    for (int j = 0; j < m.njnt; ++j)
      if (m.jnt_type[j] == kJntHinge) {
        const int qa = m.jnt_qposadr[j];
        d.qfrc_passive[m.jnt_dofadr[j]] += -m.jnt_stiffness[j] * (d.qpos[qa] - m.qpos_spring[qa]);
      }
    std::vector<T> cacc(m.nbody * 6), cfrc(m.nbody * 6);
    for (int k = 0; k < 3; ++k) { cacc[k] = 0; cacc[3 + k] = -m.gravity[k]; }
    for (int b = 1; b < m.nbody; ++b) {
      for (int k = 0; k < 6; ++k) cacc[b * 6 + k] = cacc[m.body_parentid[b] * 6 + k];
      for (int i = 0; i < m.body_dofnum[b]; ++i) {
        const int dof = m.body_dofadr[b] + i;
        for (int k = 0; k < 6; ++k) cacc[b * 6 + k] += d.cdof_dot[dof * 6 + k] * d.qvel[dof];
      }
    }
    for (int b = 0; b < m.nbody; ++b) {
      T f1[6], f2[6], f3[6];
      inert_mul(&d.cinert[b * 10], &cacc[b * 6], f1);
      inert_mul(&d.cinert[b * 10], &d.cvel[b * 6], f2);
      motion_cross_force(&d.cvel[b * 6], f2, f3);
      for (int k = 0; k < 6; ++k) cfrc[b * 6 + k] = f1[k] + f3[k];
    }
    for (int b = m.nbody - 1; b > 0; --b)
      for (int k = 0; k < 6; ++k) cfrc[m.body_parentid[b] * 6 + k] += cfrc[b * 6 + k];
    for (int i = 0; i < m.nv; ++i) d.qfrc_bias[i] = dotn(&d.cdof[i * 6], &cfrc[m.dof_bodyid[i] * 6], 6);
  }

  // ---- forward.fwd_actuation + fwd_acceleration
  void actuation_and_acceleration(Data<T>& d) const {
    const int nv = m.nv;
    std::fill(d.qfrc_actuator.begin(), d.qfrc_actuator.end(), T(0));
    for (int u = 0; u < m.nu; ++u) {
      T ctrl = d.ctrl[u];
      if (m.actuator_ctrllimited[u]) ctrl = std::min(std::max(ctrl, m.actuator_ctrlrange[u * 2]), m.actuator_ctrlrange[u * 2 + 1]);
      T ctrl_act = ctrl;
      if (m.actuator_dyn_filter[u]) {
        d.act_dot[u] = (ctrl - d.act[u]) / std::max(m.actuator_dynprm[u], T(kMinVal));
        ctrl_act = d.act[u];
      }
      T force = m.actuator_gain[u] * ctrl_act;
      if (m.actuator_bias_affine[u]) {
        T len = 0, vel = 0;
        for (int i = 6; i < nv; ++i) { len += m.actuator_moment[u * nv + i] * d.qpos[i + 1]; vel += m.actuator_moment[u * nv + i] * d.qvel[i]; }
        force += m.actuator_biasprm[u * 3] + m.actuator_biasprm[u * 3 + 1] * len + m.actuator_biasprm[u * 3 + 2] * vel;
      }
      if (m.actuator_forcelimited[u]) force = std::min(std::max(force, m.actuator_forcerange[u * 2]), m.actuator_forcerange[u * 2 + 1]);
      for (int i = 0; i < nv; ++i) d.qfrc_actuator[i] += m.actuator_moment[u * nv + i] * force;
    }
    for (int i = 0; i < nv; ++i) d.qfrc_smooth[i] = d.qfrc_passive[i] - d.qfrc_bias[i] + d.qfrc_actuator[i];
    cho_solve(d.qL.data(), d.qfrc_smooth.data(), d.qacc_smooth.data(), nv);
  }

  // ---- solver.solve (CG / Newton, primal, pyramidal cones; no equality / friction rows)
  struct Ctx {
    std::vector<T> qacc, qfrc_constraint, Jaref, efc_force, Ma, grad, Mgrad, search, H, HL;
    std::vector<char> active;
    T gauss = 0, cost = 0, prev_cost = 0;
    int niter = 0;
  };
  T rescale(T v) const { return v / (m.meaninertia * T(std::max(1, m.nv))); }

  void update_constraint(const Data<T>& d, Ctx& c) const {
    const int nv = m.nv, ne = m.nefc;
    for (int r = 0; r < ne; ++r) {
      c.active[r] = c.Jaref[r] < 0;
      c.efc_force[r] = d.efc_D[r] * -c.Jaref[r] * (c.active[r] ? T(1) : T(0));
    }
    for (int i = 0; i < nv; ++i) {
      T s = 0;
      for (int r = 0; r < ne; ++r) s += d.efc_J[size_t(r) * nv + i] * c.efc_force[r];
      c.qfrc_constraint[i] = s;
    }
    T g = 0;
    for (int i = 0; i < nv; ++i) g += (c.Ma[i] - d.qfrc_smooth[i]) * (c.qacc[i] - d.qacc_smooth[i]);
    c.gauss = T(0.5) * g;
    T s = 0;
    for (int r = 0; r < ne; ++r) s += d.efc_D[r] * c.Jaref[r] * c.Jaref[r] * (c.active[r] ? T(1) : T(0));
    c.prev_cost = c.cost;
    c.cost = T(0.5) * s + c.gauss;
  }
  void update_gradient(const Data<T>& d, Ctx& c) const {
    const int nv = m.nv, ne = m.nefc;
    for (int i = 0; i < nv; ++i) c.grad[i] = c.Ma[i] - d.qfrc_smooth[i] - c.qfrc_constraint[i];
    if (cfg.solver == TMJX_SOLVER_CG) {
      cho_solve(d.qL.data(), c.grad.data(), c.Mgrad.data(), nv);
    } else {
      c.H = d.qM;
      for (int r = 0; r < ne; ++r) {
        if (!c.active[r]) continue;
        const T* jr = &d.efc_J[size_t(r) * nv];
        for (int i = 0; i < nv; ++i) {
          if (jr[i] == T(0)) continue;
          const T a = jr[i] * d.efc_D[r];
          for (int j = 0; j < nv; ++j) c.H[i * nv + j] += a * jr[j];
        }
      }
      cholesky(c.H.data(), c.HL.data(), nv);
      cho_solve(c.HL.data(), c.grad.data(), c.Mgrad.data(), nv);
    }
  }
  void ctx_create(const Data<T>& d, const T* qacc, Ctx& c, bool grad) const {
    const int nv = m.nv, ne = m.nefc;
    c.qacc.assign(qacc, qacc + nv);
    c.qfrc_constraint.assign(nv, T(0)); c.Jaref.resize(ne); c.efc_force.assign(ne, T(0)); c.Ma.resize(nv);
    c.grad.assign(nv, T(0)); c.Mgrad.assign(nv, T(0)); c.search.assign(nv, T(0)); c.active.assign(ne, 0);
    if (cfg.solver == TMJX_SOLVER_NEWTON) { c.H.resize(nv * nv); c.HL.resize(nv * nv); }
    matvec(d.efc_J.data(), qacc, c.Jaref.data(), ne, nv);
    for (int r = 0; r < ne; ++r) c.Jaref[r] -= d.efc_aref[r];
    matvec(d.qM.data(), qacc, c.Ma.data(), nv, nv);
    c.gauss = 0; c.cost = std::numeric_limits<T>::infinity(); c.prev_cost = 0; c.niter = 0;
    update_constraint(d, c);
    if (grad) {
      update_gradient(d, c);
      for (int i = 0; i < nv; ++i) c.search[i] = -c.Mgrad[i];
    }
  }
  struct LSPoint { T alpha, cost, deriv_0, deriv_1; };
  LSPoint ls_point(const Data<T>& d, const Ctx& c, T alpha, const T* jv, const T* quad, const T* quad_gauss) const {
    T q0 = 0, q1 = 0, q2 = 0;
    for (int r = 0; r < m.nefc; ++r) {
      const T x = c.Jaref[r] + alpha * jv[r];
      if (x < 0) { q0 += quad[r * 3]; q1 += quad[r * 3 + 1]; q2 += quad[r * 3 + 2]; }
    }
    q0 += quad_gauss[0]; q1 += quad_gauss[1]; q2 += quad_gauss[2];
    LSPoint p;
    p.alpha = alpha;
    p.cost = alpha * alpha * q2 + alpha * q1 + q0;
    p.deriv_0 = T(2) * alpha * q2 + q1;
    p.deriv_1 = T(2) * q2 + (q2 == T(0) ? T(kMinVal) : T(0));
    return p;
  }
  void linesearch(const Data<T>& d, Ctx& c) const {
    const int nv = m.nv, ne = m.nefc;
    const T scale = m.meaninertia * T(std::max(1, nv));
    const T smag = std::sqrt(dotn(c.search.data(), c.search.data(), nv)) * scale;
    const T gtol = m.tolerance * m.ls_tolerance * smag;
    std::vector<T> mv(nv), jv(ne), quad(ne * 3);
    matvec(d.qM.data(), c.search.data(), mv.data(), nv, nv);
    matvec(d.efc_J.data(), c.search.data(), jv.data(), ne, nv);
    T quad_gauss[3] = {c.gauss, dotn(c.search.data(), c.Ma.data(), nv) - dotn(c.search.data(), d.qfrc_smooth.data(), nv),
                       T(0.5) * dotn(c.search.data(), mv.data(), nv)};
    for (int r = 0; r < ne; ++r) {
      quad[r * 3] = T(0.5) * c.Jaref[r] * c.Jaref[r] * d.efc_D[r];
      quad[r * 3 + 1] = jv[r] * c.Jaref[r] * d.efc_D[r];
      quad[r * 3 + 2] = T(0.5) * jv[r] * jv[r] * d.efc_D[r];
    }
    auto pt = [&](T a) { return ls_point(d, c, a, jv.data(), quad.data(), quad_gauss); };
    const LSPoint p0 = pt(T(0));
    LSPoint lo = pt(p0.alpha - p0.deriv_0 / p0.deriv_1), hi;
    const bool lesser = lo.deriv_0 < p0.deriv_0;
    hi = lesser ? p0 : lo;
    lo = lesser ? lo : p0;
    bool swap = true;
    int it = 0;
    while (true) {
      bool done = it >= cfg.ls_iterations;
      done |= !swap;
      done |= (lo.deriv_0 < 0) && (lo.deriv_0 > -gtol);
      done |= (hi.deriv_0 > 0) && (hi.deriv_0 < gtol);
      if (done) break;
      const LSPoint lo_next = pt(lo.alpha - lo.deriv_0 / lo.deriv_1);
      const LSPoint hi_next = pt(hi.alpha - hi.deriv_0 / hi.deriv_1);
      const LSPoint mid = pt(T(0.5) * (lo.alpha + hi.alpha));
      const bool swap_lo_next = (lo.deriv_0 > 0) || (lo.deriv_0 < lo_next.deriv_0);
      if (swap_lo_next) lo = lo_next;
      const bool swap_lo_mid = (mid.deriv_0 < 0) && (lo.deriv_0 < mid.deriv_0);
      if (swap_lo_mid) lo = mid;
      const bool swap_hi_next = (hi.deriv_0 < 0) || (hi.deriv_0 > hi_next.deriv_0);
      if (swap_hi_next) hi = hi_next;
      const bool swap_hi_mid = (mid.deriv_0 > 0) && (hi.deriv_0 > mid.deriv_0);
      if (swap_hi_mid) hi = mid;
      swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
      ++it;
    }
    const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
    const T alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
    if (improved)
      for (int i = 0; i < nv; ++i) { c.qacc[i] += c.search[i] * alpha; c.Ma[i] += mv[i] * alpha; }
    if (improved)
      for (int r = 0; r < ne; ++r) c.Jaref[r] += jv[r] * alpha;
  }
  void solve(Data<T>& d) const {
    const int nv = m.nv;
    Ctx warm, smth, c;
    ctx_create(d, d.warm.data(), warm, false);
    ctx_create(d, d.qacc_smooth.data(), smth, false);
    const T* start = warm.cost < smth.cost ? d.warm.data() : d.qacc_smooth.data();
    ctx_create(d, start, c, true);
    auto body = [&]() {
      linesearch(d, c);
      std::vector<T> prev_grad = c.grad, prev_Mgrad = c.Mgrad;
      update_constraint(d, c);
      update_gradient(d, c);
      if (cfg.solver == TMJX_SOLVER_NEWTON) {
        for (int i = 0; i < nv; ++i) c.search[i] = -c.Mgrad[i];
      } else {
        T num = 0;
        for (int i = 0; i < nv; ++i) num += c.grad[i] * (c.Mgrad[i] - prev_Mgrad[i]);
        T beta = num / std::max(T(kMinVal), dotn(prev_grad.data(), prev_Mgrad.data(), nv));
        beta = std::max(T(0), beta);
        for (int i = 0; i < nv; ++i) c.search[i] = -c.Mgrad[i] + beta * c.search[i];
      }
      ++c.niter;
    };
    if (cfg.iterations == 1) {
      body();
    } else {
      while (true) {
        const T improvement = rescale(c.prev_cost - c.cost);
        const T gradient = rescale(std::sqrt(dotn(c.grad.data(), c.grad.data(), nv)));
        bool done = c.niter >= cfg.iterations;
        done |= improvement < m.tolerance;
        done |= gradient < m.tolerance;
        if (done) break;
        body();
      }
    }
    d.qacc = c.qacc;
    d.warm = c.qacc;
    d.qfrc_constraint = c.qfrc_constraint;
    d.efc_force = c.efc_force;
  }

  // ---- forward.forward (sensors / camlight / tendon lengths are not consumed by the task and are omitted)
  void forward(Data<T>& d) const {
    kinematics(d);
    com_pos(d);
    crb_and_factor(d);
    collision(d);
    make_constraint(d);
    com_vel(d);
    passive_and_rne(d);
    actuation_and_acceleration(d);
    solve(d);
  }

  // ---- forward.euler + _advance
  void euler(Data<T>& d) const {
    const int nv = m.nv;
    std::vector<T> qM2 = d.qM, L2(nv * nv), rhs(nv), qacc(nv);
    for (int i = 0; i < nv; ++i) qM2[i * nv + i] += dt * m.dof_damping[i];
    cholesky(qM2.data(), L2.data(), nv);
    for (int i = 0; i < nv; ++i) rhs[i] = d.qfrc_smooth[i] + d.qfrc_constraint[i];
    cho_solve(L2.data(), rhs.data(), qacc.data(), nv);
    for (int u = 0; u < m.na; ++u) d.act[u] = d.act[u] + d.act_dot[u] * dt;
    for (int i = 0; i < nv; ++i) d.qvel[i] = d.qvel[i] + qacc[i] * dt;
    for (int j = 0; j < m.njnt; ++j) {
      const int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == kJntFree) {
        for (int k = 0; k < 3; ++k) d.qpos[qa + k] = d.qpos[qa + k] + dt * d.qvel[da + k];
        T v[3] = {d.qvel[da + 3], d.qvel[da + 4], d.qvel[da + 5]};  // math.quat_integrate
        const T nrm = normalize(v, 3);
        T qr[4], q2[4];
        axis_angle_to_quat(v, dt * nrm, qr);
        quat_mul(&d.qpos[qa + 3], qr, q2);
        normalize(q2, 4);
        for (int k = 0; k < 4; ++k) d.qpos[qa + 3 + k] = q2[k];
      } else {
        d.qpos[qa] = d.qpos[qa] + dt * d.qvel[da];
      }
    }
    d.time = d.time + dt;
  }

  // ------------------------------------------------------------------ task layer
  Clips clips;

  int clip_frame_clamped(int f) const { return std::min(std::max(f, 0), clips.clip_len - 1); }

  // _get_cur_frame (single_clip_tracking.py:452-454): floor(time * mocap_hz + start_frame) as int32, in fp32-style
  // unfused mul / add (volatile keeps the compiler from contracting into an FMA)
  int cur_frame(T time, int start_frame) const {
    volatile T prod = time * T(cfg.mocap_hz);
    volatile T sum = prod + T(start_frame);
    return int(std::floor(sum));
  }

  // _get_obs (single_clip_tracking.py:394-450) + walker transforms (walker/base.py:170-258)
  void get_obs(const Data<T>& d, int clip, int frame, T* obs) const {
    const int L = cfg.traj_length, nj = clips.n_joints, nrb = clips.n_ref_bodies;
    // dynamic_slice_in_dim(x, cur_frame + 1, L): the start is clamped so the slice stays in bounds
    const int start = std::min(std::max(frame + 1, 0), clips.clip_len - L);
    const T* root = &d.qpos[0];
    const T* quat = &d.qpos[3];
    T* o_track = obs;
    T* o_quat = o_track + 3 * L;
    T* o_joint = o_quat + 4 * L;
    T* o_body = o_joint + cfg.n_joint_idxs * L;
    T* o_prop = o_body + 3 * cfg.n_body_idxs * L;
    for (int t = 0; t < L; ++t) {
      const size_t fr = size_t(clip) * clips.clip_len + start + t;
      T dlt[3];
      for (int k = 0; k < 3; ++k) dlt[k] = T(clips.position[fr * 3 + k]) - root[k];
      rotate(dlt, quat, o_track + 3 * t);
      // relative_quat(ref, agent) = quat_mul(agent, ref * [1,-1,-1,-1])
      T rq[4] = {T(clips.quaternion[fr * 4]), -T(clips.quaternion[fr * 4 + 1]), -T(clips.quaternion[fr * 4 + 2]),
                 -T(clips.quaternion[fr * 4 + 3])};
      quat_mul(quat, rq, o_quat + 4 * t);
      for (int i = 0; i < cfg.n_joint_idxs; ++i) {
        // (ref_joints - qpos[7:])[:, joint_idxs - 1]; negative indices wrap, out-of-range clamps (jnp gather)
        int col = cfg.joint_idxs[i] - 1;
        if (col < 0) col += nj;
        col = std::min(std::max(col, 0), nj - 1);
        o_joint[t * cfg.n_joint_idxs + i] = T(clips.joints[fr * nj + col]) - d.qpos[7 + col];
      }
      for (int i = 0; i < cfg.n_body_idxs; ++i) {
        // (ref_positions - xpos[1:])[:, body_idxs]: 67-row arrays indexed by MODEL ids, clamped
        const int row = std::min(std::max(cfg.body_idxs[i], 0), nrb - 1);
        T dl[3];
        for (int k = 0; k < 3; ++k) dl[k] = T(clips.body_positions[(fr * nrb + row) * 3 + k]) - d.xpos[(row + 1) * 3 + k];
        rotate(dl, quat, o_body + (t * cfg.n_body_idxs + i) * 3);
      }
    }
    // _get_proprioception (single_clip_tracking.py:336-354)
    T* p = o_prop;
    for (int i = 7; i < m.nq; ++i) *p++ = d.qpos[i];
    for (int i = 6; i < m.nv; ++i) *p++ = d.qvel[i];
    for (int i = 0; i < m.nv; ++i) *p++ = d.qfrc_actuator[i];
    const int tb = cfg.torso_body_id;
    *p++ = d.xpos[tb * 3 + 2];
    for (int k = 0; k < 3; ++k) *p++ = d.xmat[tb * 9 + 6 + k];
    for (int a = 0; a < cfg.n_appendages; ++a) {  // dot(positions - torso.xpos, torso.xmat)
      const int b = cfg.appendage_body_ids[a];
      T dl[3];
      for (int k = 0; k < 3; ++k) dl[k] = d.xpos[b * 3 + k] - d.xpos[tb * 3 + k];
      for (int c = 0; c < 3; ++c) *p++ = dl[0] * d.xmat[tb * 9 + c] + dl[1] * d.xmat[tb * 9 + 3 + c] + dl[2] * d.xmat[tb * 9 + 6 + c];
    }
  }
  int obs_size() const {
    return cfg.traj_length * (3 + 4 + cfg.n_joint_idxs + 3 * cfg.n_body_idxs) + (m.nq - 7) + (m.nv - 6) + m.nv + 1 + 3 +
           3 * cfg.n_appendages;
  }

  static bool has_nan(const std::vector<T>& v) {
    for (T x : v) if (std::isnan(x)) return true;
    return false;
  }
  static T nan_to_num(T x) {
    if (std::isnan(x)) return T(0);
    if (std::isinf(x)) return x > 0 ? std::numeric_limits<T>::max() : std::numeric_limits<T>::lowest();
    return x;
  }
};

// ------------------------------------------------------------------ C entry points
struct Oracle {
  Sim<float> f;
  Sim<double> d;
};

template <class T> T* col(void* base, size_t e, size_t dim) { return static_cast<T*>(base) + e * dim; }

template <class T> void load_state(const Sim<T>& s, Data<T>& d, const TmjxState* st, size_t e) {
  const auto& m = s.m;
  std::memcpy(d.qpos.data(), col<T>(st->qpos, e, m.nq), sizeof(T) * m.nq);
  std::memcpy(d.qvel.data(), col<T>(st->qvel, e, m.nv), sizeof(T) * m.nv);
  std::memcpy(d.act.data(), col<T>(st->act, e, m.na), sizeof(T) * m.na);
  std::memcpy(d.warm.data(), col<T>(st->qacc_warmstart, e, m.nv), sizeof(T) * m.nv);
  d.time = *col<T>(st->time, e, 1);
}
template <class T> void store_state(const Sim<T>& s, const Data<T>& d, TmjxState* st, size_t e) {
  const auto& m = s.m;
  std::memcpy(col<T>(st->qpos, e, m.nq), d.qpos.data(), sizeof(T) * m.nq);
  std::memcpy(col<T>(st->qvel, e, m.nv), d.qvel.data(), sizeof(T) * m.nv);
  std::memcpy(col<T>(st->act, e, m.na), d.act.data(), sizeof(T) * m.na);
  std::memcpy(col<T>(st->qacc_warmstart, e, m.nv), d.warm.data(), sizeof(T) * m.nv);
  *col<T>(st->time, e, 1) = d.time;
  std::memcpy(col<T>(st->xpos, e, m.nbody * 3), d.xpos.data(), sizeof(T) * m.nbody * 3);
  std::memcpy(col<T>(st->xquat, e, m.nbody * 4), d.xquat.data(), sizeof(T) * m.nbody * 4);
  std::memcpy(col<T>(st->qfrc_actuator, e, m.nv), d.qfrc_actuator.data(), sizeof(T) * m.nv);
}
template <class T> void store_debug(const Sim<T>& s, const Data<T>& d, TmjxOut* o, size_t e) {
  const auto& m = s.m;
  if (o->dbg_qacc) std::memcpy(col<T>(o->dbg_qacc, e, m.nv), d.qacc.data(), sizeof(T) * m.nv);
  if (o->dbg_qacc_smooth) std::memcpy(col<T>(o->dbg_qacc_smooth, e, m.nv), d.qacc_smooth.data(), sizeof(T) * m.nv);
  if (o->dbg_qfrc_bias) std::memcpy(col<T>(o->dbg_qfrc_bias, e, m.nv), d.qfrc_bias.data(), sizeof(T) * m.nv);
  if (o->dbg_qfrc_constraint) std::memcpy(col<T>(o->dbg_qfrc_constraint, e, m.nv), d.qfrc_constraint.data(), sizeof(T) * m.nv);
  if (o->dbg_contact_dist) std::memcpy(col<T>(o->dbg_contact_dist, e, m.ncon), d.con_dist.data(), sizeof(T) * m.ncon);
  if (o->dbg_efc_force) std::memcpy(col<T>(o->dbg_efc_force, e, m.nefc), d.efc_force.data(), sizeof(T) * m.nefc);
  if (o->dbg_qM) std::memcpy(col<T>(o->dbg_qM, e, size_t(m.nv) * m.nv), d.qM.data(), sizeof(T) * m.nv * m.nv);
  if (o->dbg_subtree_com) {
    int root = 0;
    for (int b = 1; b < m.nbody; ++b) if (m.body_dofnum[b] > 0) { root = m.body_rootid[b]; break; }
    std::memcpy(col<T>(o->dbg_subtree_com, e, 3), &d.subtree_com[root * 3], sizeof(T) * 3);
  }
}

// reset path: reset_from_clip after qpos/qvel were chosen (single_clip_tracking.py:163-205)
template <class T> void run_forward(const Sim<T>& s, TmjxState* st, TmjxOut* o, int n_env, unsigned flags, int nthreads) {
  const auto& m = s.m;
  const auto& cfg = s.cfg;
  const int nobs = s.obs_size(), W = cfg.var_window_size;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int e = 0; e < n_env; ++e) {
    Data<T> d(m);
    load_state(s, d, st, e);
    std::fill(d.act.begin(), d.act.end(), T(0));
    std::fill(d.warm.begin(), d.warm.end(), T(0));
    std::fill(d.ctrl.begin(), d.ctrl.end(), T(0));
    d.time = 0;
    s.forward(d);
    store_state(s, d, st, e);
    const int clip = st->clip_idx[e], sf = st->start_frame[e];
    const int frame = s.cur_frame(d.time, sf);
    s.get_obs(d, clip, frame, col<T>(o->obs, e, nobs));
    *col<T>(o->reward, e, 1) = 0;
    *col<T>(o->done, e, 1) = 0;
    for (int k = 0; k < TMJX_N_METRICS; ++k) col<T>(o->metrics, e, TMJX_N_METRICS)[k] = 0;
    o->cur_frame[e] = frame;
    for (int k = 0; k < W * m.nu; ++k) col<T>(st->action_buffer, e, size_t(W) * m.nu)[k] = 0;
    for (int k = 0; k < m.nu; ++k) col<T>(st->prev_ctrl, e, m.nu)[k] = 0;
    st->buffer_index[e] = 0;
    store_debug(s, d, o, e);
    if (flags & TMJX_F_SNAPSHOT) {
      *col<T>(st->steps, e, 1) = 0;
      *col<T>(st->truncation, e, 1) = 0;
      std::memcpy(col<T>(st->first_qpos, e, m.nq), d.qpos.data(), sizeof(T) * m.nq);
      std::memcpy(col<T>(st->first_qvel, e, m.nv), d.qvel.data(), sizeof(T) * m.nv);
      std::memcpy(col<T>(st->first_act, e, m.na), d.act.data(), sizeof(T) * m.na);
      *col<T>(st->first_time, e, 1) = d.time;
      std::memcpy(col<T>(st->first_qacc_warmstart, e, m.nv), d.warm.data(), sizeof(T) * m.nv);
      std::memcpy(col<T>(st->first_xpos, e, m.nbody * 3), d.xpos.data(), sizeof(T) * m.nbody * 3);
      std::memcpy(col<T>(st->first_xquat, e, m.nbody * 4), d.xquat.data(), sizeof(T) * m.nbody * 4);
      std::memcpy(col<T>(st->first_qfrc_actuator, e, m.nv), d.qfrc_actuator.data(), sizeof(T) * m.nv);
      std::memcpy(col<T>(st->first_obs, e, nobs), col<T>(o->obs, e, nobs), sizeof(T) * nobs);
      for (int k = 0; k < m.nu; ++k) col<T>(st->first_prev_ctrl, e, m.nu)[k] = 0;
    }
  }
}

// SingleClipTracking.step (single_clip_tracking.py:207-320) [+ wrappers when TMJX_F_AUTORESET]
template <class T>
void run_step(const Sim<T>& s, const void* action_v, TmjxState* st, TmjxOut* o, int n_env, unsigned flags, int nthreads) {
  const auto& m = s.m;
  const auto& cfg = s.cfg;
  const auto& cl = s.clips;
  const int nobs = s.obs_size(), W = cfg.var_window_size, nu = m.nu;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int e = 0; e < n_env; ++e) {
    Data<T> d(m);
    load_state(s, d, st, e);
    const T* action = static_cast<const T*>(action_v) + size_t(e) * nu;
    if (flags & TMJX_F_AUTORESET) {  // wrappers.py:107-113: zero `steps` where the previous state was done
      if (*col<T>(o->done, e, 1) != T(0)) *col<T>(st->steps, e, 1) = 0;
    }
    // pipeline_step: n_frames x mjx.step with ctrl = action (brax PipelineEnv; called at :219)
    for (int k = 0; k < nu; ++k) d.ctrl[k] = action[k];
    for (int f = 0; f < cfg.physics_steps_per_control_step; ++f) {
      s.forward(d);
      s.euler(d);
    }
    const int clip = st->clip_idx[e], sf = st->start_frame[e];
    const int frame = s.cur_frame(d.time, sf);                   // :223-225
    const int fclamp = s.clip_frame_clamped(frame);              // jnp gather clamps out-of-range indices
    const size_t fr = size_t(clip) * cl.clip_len + fclamp;
    T* prev_ctrl = col<T>(st->prev_ctrl, e, nu);
    for (int k = 0; k < nu; ++k) prev_ctrl[k] = action[k];       // :227 (before the reward => ctrl_diff == 0)
    T* buf = col<T>(st->action_buffer, e, size_t(W) * nu);
    int idx = st->buffer_index[e];
    for (int k = 0; k < nu; ++k) buf[idx * nu + k] = action[k];  // :229-234
    idx = (idx + 1) % W;
    st->buffer_index[e] = idx;

    // compute_tracking_rewards (reward.py:359-485)
    T pos_dist[3], ssq = 0;
    for (int k = 0; k < 3; ++k) { pos_dist[k] = d.qpos[k] - T(cl.position[fr * 3 + k]); ssq += pos_dist[k] * pos_dist[k]; }
    const T pos_reward = T(cfg.pos_reward_weight) * std::exp(-T(cfg.pos_reward_exp_scale) * ssq);
    T qs[4], qt[4];
    for (int k = 0; k < 4; ++k) { qs[k] = d.qpos[3 + k]; qt[k] = T(cl.quaternion[fr * 4 + k]); }
    {  // _bounded_quat_dist: plain x / ||x||
      T ns = std::sqrt(qs[0] * qs[0] + qs[1] * qs[1] + qs[2] * qs[2] + qs[3] * qs[3]);
      T nt = std::sqrt(qt[0] * qt[0] + qt[1] * qt[1] + qt[2] * qt[2] + qt[3] * qt[3]);
      for (int k = 0; k < 4; ++k) { qs[k] /= ns; qt[k] /= nt; }
    }
    const T qd = qs[0] * qt[0] + qs[1] * qt[1] + qs[2] * qt[2] + qs[3] * qt[3];
    const T bq = T(0.5) * std::acos(jmin(T(1), T(2) * qd * qd - T(1)));
    const T quat_distance = bq * bq;
    const T quat_reward = T(cfg.quat_reward_weight) * std::exp(-T(cfg.quat_reward_exp_scale) * quat_distance);
    T joint_distance = 0;
    for (int j = 0; j < cl.n_joints; ++j) { const T x = d.qpos[7 + j] - T(cl.joints[fr * cl.n_joints + j]); joint_distance += x * x; }
    const T joint_reward = T(cfg.joint_reward_weight) * std::exp(-T(cfg.joint_reward_exp_scale) * joint_distance);
    T av = 0;
    for (int k = 0; k < 3; ++k) { const T x = d.qvel[3 + k] - T(cl.angular_velocity[fr * 3 + k]); av += x * x; }
    const T angvel_reward = T(cfg.angvel_reward_weight) * std::exp(-T(cfg.angvel_reward_exp_scale) * av);
    auto body_err = [&](const int32_t* ids, int n) {
      T sacc = 0;
      for (int i = 0; i < n; ++i) {
        const int row = std::min(std::max(ids[i], 0), cl.n_ref_bodies - 1);  // xpos[1:][ids], body_positions[ids]
        for (int k = 0; k < 3; ++k) {
          const T x = d.xpos[(row + 1) * 3 + k] - T(cl.body_positions[(fr * cl.n_ref_bodies + row) * 3 + k]);
          sacc += x * x;
        }
      }
      return sacc;
    };
    const T bodypos_reward = T(cfg.bodypos_reward_weight) * std::exp(-T(cfg.bodypos_reward_exp_scale) * body_err(cfg.body_idxs, cfg.n_body_idxs));
    const T endeff_reward = T(cfg.endeff_reward_weight) * std::exp(-T(cfg.endeff_reward_exp_scale) * body_err(cfg.endeff_idxs, cfg.n_endeff_idxs));
    T a2 = 0, ad = 0;
    for (int k = 0; k < nu; ++k) { a2 += action[k] * action[k]; const T x = prev_ctrl[k] - action[k]; ad += x * x; }
    const T ctrl_cost = T(cfg.ctrl_cost_weight) * a2;
    const T ctrl_diff_cost = T(cfg.ctrl_diff_cost_weight) * ad;
    T en = 0;
    for (int i = 6; i < m.nv; ++i) en += std::abs(d.qvel[i]) * std::abs(d.qfrc_actuator[i]);
    const T energy_cost = T(cfg.energy_cost_weight) * jmin(en, T(50));
    const T torso_z = d.xpos[cfg.torso_idx * 3 + 2];
    T healthy = torso_z < T(cfg.healthy_z_min) ? T(0) : T(1);
    if (torso_z > T(cfg.healthy_z_max)) healthy = 0;
    const T fall = T(1) - healthy;
    T summed = 0;
    for (int k = 0; k < 3; ++k) { const T x = pos_dist[k] * T(cfg.penalty_pos_distance_scale[k]); summed += x * x; }
    const T too_far = summed > T(cfg.too_far_dist) ? T(1) : T(0);
    const T bad_pose = joint_distance > T(cfg.bad_pose_dist) ? T(1) : T(0);
    const T bad_quat = quat_distance > T(cfg.bad_quat_dist) ? T(1) : T(0);
    T var_sum = 0;
    for (int k = 0; k < nu; ++k) {
      T mean = 0;
      for (int t = 0; t < W; ++t) mean += buf[t * nu + k];
      mean /= T(W);
      T v = 0;
      for (int t = 0; t < W; ++t) { const T x = buf[t * nu + k] - mean; v += x * x; }
      var_sum += v / T(W);
    }
    const T var_cost = T(cfg.var_coeff) * var_sum;
    T jerk = 0;
    for (int t = 0; t + 2 < W; ++t)
      for (int k = 0; k < nu; ++k) {
        const T b0 = buf[((idx + t) % W) * nu + k], b1 = buf[((idx + t + 1) % W) * nu + k], b2 = buf[((idx + t + 2) % W) * nu + k];
        const T x = b2 - T(2) * b1 + b0;
        jerk += x * x;
      }
    const T jerk_cost = T(cfg.jerk_coeff) * jerk;

    T* obs = col<T>(o->obs, e, nobs);
    s.get_obs(d, clip, frame, obs);
    T reward = joint_reward + pos_reward + quat_reward + angvel_reward + bodypos_reward + endeff_reward - ctrl_cost -
               ctrl_diff_cost - energy_cost - var_cost - jerk_cost;
    T done = std::max(std::max(fall, too_far), std::max(bad_pose, bad_quat));
    reward = Sim<T>::nan_to_num(reward);
    for (int k = 0; k < nobs; ++k) obs[k] = Sim<T>::nan_to_num(obs[k]);
    // ravel_pytree(data) NaN scan (:290-293), restricted to the mjx.Data fields this build materialises
    bool nanflag = Sim<T>::has_nan(d.qpos) || Sim<T>::has_nan(d.qvel) || Sim<T>::has_nan(d.act) || Sim<T>::has_nan(d.warm) ||
                   Sim<T>::has_nan(d.xpos) || Sim<T>::has_nan(d.xquat) || Sim<T>::has_nan(d.qfrc_actuator) ||
                   Sim<T>::has_nan(d.qacc) || Sim<T>::has_nan(d.qfrc_constraint) || Sim<T>::has_nan(d.efc_force) ||
                   Sim<T>::has_nan(d.qfrc_bias) || Sim<T>::has_nan(d.qacc_smooth) || std::isnan(d.time);
    const T nanv = nanflag ? T(1) : T(0);
    done = std::max(nanv, done);
    T* mt = col<T>(o->metrics, e, TMJX_N_METRICS);
    mt[TMJX_M_POS_REWARD] = pos_reward; mt[TMJX_M_QUAT_REWARD] = quat_reward; mt[TMJX_M_JOINT_REWARD] = joint_reward;
    mt[TMJX_M_ANGVEL_REWARD] = angvel_reward; mt[TMJX_M_BODYPOS_REWARD] = bodypos_reward; mt[TMJX_M_ENDEFF_REWARD] = endeff_reward;
    mt[TMJX_M_CTRL_COST] = -ctrl_cost; mt[TMJX_M_CTRL_DIFF_COST] = -ctrl_diff_cost; mt[TMJX_M_ENERGY_COST] = -energy_cost;
    mt[TMJX_M_DONE] = done; mt[TMJX_M_TOO_FAR] = too_far; mt[TMJX_M_BAD_POSE] = bad_pose; mt[TMJX_M_BAD_QUAT] = bad_quat;
    mt[TMJX_M_FALL] = fall; mt[TMJX_M_NAN] = nanv; mt[TMJX_M_JOINT_DISTANCE] = joint_distance;
    mt[TMJX_M_SUMMED_POS_DISTANCE] = summed; mt[TMJX_M_QUAT_DISTANCE] = quat_distance; mt[TMJX_M_VAR_COST] = -var_cost;
    mt[TMJX_M_JERK_COST] = -jerk_cost;
    o->cur_frame[e] = frame;
    *col<T>(o->reward, e, 1) = reward;
    store_debug(s, d, o, e);

    if (flags & TMJX_F_AUTORESET) {
      // brax EpisodeWrapper.step (action_repeat = 1) then wrappers.py:115-131
      T steps = *col<T>(st->steps, e, 1) + T(1);
      const bool over = steps >= T(cfg.episode_length);
      *col<T>(st->truncation, e, 1) = over ? T(1) - done : T(0);
      if (over) done = 1;
      *col<T>(st->steps, e, 1) = steps;
      if (done != T(0)) {
        std::memcpy(d.qpos.data(), col<T>(st->first_qpos, e, m.nq), sizeof(T) * m.nq);
        std::memcpy(d.qvel.data(), col<T>(st->first_qvel, e, m.nv), sizeof(T) * m.nv);
        std::memcpy(d.act.data(), col<T>(st->first_act, e, m.na), sizeof(T) * m.na);
        d.time = *col<T>(st->first_time, e, 1);
        std::memcpy(d.warm.data(), col<T>(st->first_qacc_warmstart, e, m.nv), sizeof(T) * m.nv);
        std::memcpy(d.xpos.data(), col<T>(st->first_xpos, e, m.nbody * 3), sizeof(T) * m.nbody * 3);
        std::memcpy(d.xquat.data(), col<T>(st->first_xquat, e, m.nbody * 4), sizeof(T) * m.nbody * 4);
        std::memcpy(d.qfrc_actuator.data(), col<T>(st->first_qfrc_actuator, e, m.nv), sizeof(T) * m.nv);
        std::memcpy(obs, col<T>(st->first_obs, e, nobs), sizeof(T) * nobs);
        std::memcpy(prev_ctrl, col<T>(st->first_prev_ctrl, e, nu), sizeof(T) * nu);
      }
    }
    *col<T>(o->done, e, 1) = done;
    store_state(s, d, st, e);
  }
}

}  // namespace

extern "C" {

const char* tmjx_oracle_last_error(void) { return g_err.c_str(); }

int tmjx_oracle_create(const void* blob, size_t nbytes, const TmjxTaskConfig* cfg, void** out) {
  try {
    if (!blob || !cfg || !out) throw std::runtime_error("null argument");
    if (cfg->abi_version != TMJX_ABI_VERSION) throw std::runtime_error("abi version mismatch");
    tmjx::Blob b(blob, nbytes);
    auto* o = new Oracle();
    o->f.m.load(b); o->d.m.load(b);
    o->f.cfg = *cfg; o->d.cfg = *cfg;
    o->f.m.timestep = cfg->mj_model_timestep; o->d.m.timestep = double(cfg->mj_model_timestep);
    o->f.dt = o->f.m.timestep; o->d.dt = o->d.m.timestep;
    *out = o;
    return TMJX_OK;
  } catch (const std::exception& e) { g_err = e.what(); return TMJX_E_BLOB; }
}
void tmjx_oracle_destroy(void* h) { delete static_cast<Oracle*>(h); }

int tmjx_oracle_set_clips(void* h, const float* position, const float* quaternion, const float* joints,
                          const float* body_positions, const float* angular_velocity, int n_clips, int clip_len,
                          int n_ref_bodies) {
  auto* o = static_cast<Oracle*>(h);
  Clips c;
  c.n_clips = n_clips; c.clip_len = clip_len; c.n_ref_bodies = n_ref_bodies; c.n_joints = o->f.m.nq - 7;
  const size_t nf = size_t(n_clips) * clip_len;
  c.position.assign(position, position + nf * 3);
  c.quaternion.assign(quaternion, quaternion + nf * 4);
  c.joints.assign(joints, joints + nf * c.n_joints);
  c.body_positions.assign(body_positions, body_positions + nf * n_ref_bodies * 3);
  c.angular_velocity.assign(angular_velocity, angular_velocity + nf * 3);
  o->f.clips = c;
  o->d.clips = c;
  return TMJX_OK;
}
int tmjx_oracle_obs_size(void* h) { return static_cast<Oracle*>(h)->f.obs_size(); }

/* dtype: 0 = fp32 buffers, 1 = fp64 buffers (every float* in TmjxState/TmjxOut is then a double*) */
int tmjx_oracle_forward(void* h, TmjxState* s, TmjxOut* out, int n_env, unsigned flags, int dtype, int nthreads) {
  auto* o = static_cast<Oracle*>(h);
  try {
    if (dtype == 0) run_forward(o->f, s, out, n_env, flags, nthreads);
    else run_forward(o->d, s, out, n_env, flags, nthreads);
    return TMJX_OK;
  } catch (const std::exception& e) { g_err = e.what(); return TMJX_E_ARG; }
}
int tmjx_oracle_step(void* h, const void* action, TmjxState* s, TmjxOut* out, int n_env, unsigned flags, int dtype,
                     int nthreads) {
  auto* o = static_cast<Oracle*>(h);
  try {
    if (dtype == 0) run_step(o->f, action, s, out, n_env, flags, nthreads);
    else run_step(o->d, action, s, out, n_env, flags, nthreads);
    return TMJX_OK;
  } catch (const std::exception& e) { g_err = e.what(); return TMJX_E_ARG; }
}

}  // extern "C"
