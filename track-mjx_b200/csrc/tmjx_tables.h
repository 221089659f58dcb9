/*
 * tmjx_tables.h — device-side model tables for the rodent-tracking step kernel, and the host code that
 * derives them from the model-constant blob (include/tmjx_blob.h).
 *
 * Everything the kernel needs from `mjx.Model` (reference single_clip_tracking.py:74,91) is re-laid-out here
 * for a warp-per-environment execution: bodies sorted by tree depth (level-parallel kinematics), child lists
 * (deterministic leaf-to-root gathers), MuJoCo-style sparse rows for the joint-space inertia (row i = dof i and
 * its ancestors), the column view of the same sparsity (descendant lists) for scatter-form triangular solves,
 * and a Jacobian-free description of the plane contacts (contact bodies + their dof chains).
 */
#ifndef TMJX_TABLES_H_
#define TMJX_TABLES_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tmjx.h"
#include "../../include/tmjx_blob.h"
#include "tmjx_gen_tree.cuh"

namespace tmjx {

constexpr int kNvSlots = 3;     // nv <= 96: lane l owns dofs l, l+32, l+64
constexpr int kLimSlots = 3;    // nlimit <= 96
constexpr int kMaxCon = 32;     // one lane per contact
constexpr int kRowSlots = 4 + kLimSlots;  // constraint rows owned by a lane: 4 pyramid rows of contact `lane` + 3 limits

enum { kGeomSphere = 2, kGeomCapsule = 3, kGeomEllipsoid = 4 };
enum { kJntFree = 0, kJntHinge = 3 };

/* Plain-old-data view passed to the kernel by value (lives in the constant bank). All pointers are device
 * pointers into one int32, one uint16, one uint8 and one float allocation. */
struct DevModel {
  int nq, nv, nu, na, nbody, njnt, nlevel, nM, ncg, ncon, ncb, nlimit, nefc, maxdepth;
  int n_frames, solver, iterations, ls_iterations;
  float dt, gravity[3], tolerance, ls_tolerance, meaninertia_scale /* meaninertia * max(1, nv) */, impratio;
  float plane_pos[3], plane_n[3];
  float tree_mass;  // total mass of the moving tree (fp32 sum in body order)
  int tree_root;    // body id of the root of the moving tree
  // ---- per body
  const int *body_parent, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum, *lvl_start, *lvl_body, *child_start,
      *child;
  const float *body_pos, *body_quat, *body_ipos, *body_iquat, *body_inertia, *body_mass, *body_tree_mass;
  /* log-depth tree scans: anc_pow[r * nbody + b] = ancestor of body b at distance 2^r (0 = none / world);
   * dsc_list[dsc_start[r * (nbody + 1) + b] ..) = descendants of b at distance exactly 2^r (b >= 1) */
  int nround;
  const uint8_t *anc_pow, *dsc_list;
  const uint16_t* dsc_start;
  /* dsc_pack[r * nbody + b] = count << 16 | (count == 1 ? the descendant : first index into dsc_list): one load per
   * (round, body) on the common path (chains: at most one descendant at each distance) */
  const uint32_t* dsc_pack;
  /* dsc4[r * nbody + b] = up to four descendant ids at distance 2^r, one per byte (0 = none); dsc_maxc[r][slot] = the largest
   * count among the bodies lane + 32 slot (warp-uniform trip count of the gather); use_dsc4 = every count <= 4 */
  const uint32_t* dsc4;
  uint8_t dsc_maxc[8][4];
  int use_dsc4;
  /* per sparse-inertia entry e = lane + 32 it: row | col << 8, two iterations per word:
   * m_rc2[lane + 32 h] = rc[lane + 32 (2 h)] | rc[lane + 32 (2 h + 1)] << 16 */
  const uint32_t* m_rc2;
  // ---- per joint
  const int *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_body;
  const float *jnt_pos, *jnt_axis, *jnt_stiffness, *jnt_qpos0, *jnt_springref;
  // ---- per dof
  const int *dof_body, *dof_jnt, *dof_madr, *dof_depth, *dof_limit /* limit row or -1 */, *dof_qadr /* hinge: qpos adr, else -1 */;
  const float *dof_armature, *dof_damping;
  // ---- sparse inertia structure
  const uint8_t *m_anc /* [nM] dof id at (row, position) */, *m_row, *m_col /* entry -> (i, j) */, *tri_a, *tri_b;
  /* register-resident triangular solves: lane l owns dofs l + 32 t.  dmask[(t*3+s)*32 + l] = bit li set when dof
   * li + 32 s is a DESCENDANT of dof l + 32 t; amask likewise for ANCESTORS; rowend_me[t*32+l] = madr + depth of
   * the lane's dof (index of L[i][j] is rowend(i) - depth(j)); depth_me its depth.  rowend / depth per dof id are
   * also kept in the kernel-parameter (constant) bank for uniform access. */
  const uint32_t *dmask, *amask;
  const int *rowend_me, *depth_me;
  uint16_t u_rowend[96];
  uint8_t u_depth[96];
  uint8_t slot_used[3][3];    /* [t][s]: some lane has a non-zero dmask word (same as amask[s][t]) */
  /* depth-lane factorisation: u_ancre[rowend(k) - a] = rowend of the ancestor of dof k that has depth a (a <= depth(k)),
   * i.e. the sparse position of entry (k, anc) maps to the row that the rank-1 update of pivot k touches. */
  uint16_t u_ancre[1280];
  int use_gen;                /* the dof tree equals the one csrc/tmjx_gen_tree.cuh was generated for */
  int sync_every;             /* the substep barrier is taken every sync_every-th substep (1 = every substep) */
  int sync_mask;              /* extra phase barriers: 1 after com_vel_rne, 2 after build_m, 4 after the factorisation, 8 before the solver, 16 per solver iteration (default for Newton), 64 after the line search */
  int sync_level;             /* 2: every phase barrier, 1: major phases only, 0: once per substep (block-uniform) */
  /* factorisation pair table: for step k the entries [pair_start[k], pair_start[k+1]) = a | b << 8 | target << 16 */
  const uint32_t* pair_tab;
  const int* pair_start;
  const int* desc_start;      /* [nv+1] */
  const uint8_t* desc_dof;    /* descendant dof id */
  const uint16_t* desc_off;   /* offset of L[desc][j] in the sparse array */
  // ---- actuators
  const float *act_gain, *act_ctrl_lo, *act_ctrl_hi, *act_dyn_inv /* max(dynprm[0], mjMINVAL): the filter time constant (divisor) */, *act_bias /* [nu,3] */,
      *act_force_lo, *act_force_hi;
  const int* act_flags;       /* bit0 ctrllimited, bit1 filter, bit2 affine bias, bit3 forcelimited */
  const int *dof_act_start, *dof_act_id;
  const float* dof_act_coef;  /* moment entries, CSR by dof */
  const int *act_mom_start, *act_mom_dof;
  const float* act_mom_coef;  /* same entries, CSR by actuator (affine bias length / velocity) */
  // ---- joint limits (row l)
  const int *lim_dof, *lim_qadr;
  const float* lim_par;       /* [nlimit, 12]: lo hi margin invweight k b dmin dmax width mid power pad */
  // ---- contacts
  const int *cg_type, *cg_body, *cg_cb;
  const float *cg_pos, *cg_quat, *cg_size;
  const int *con_geom, *con_side /* +1 / -1 capsule end, 0 single */, *con_cb;
  const float* con_par;       /* [ncon, 12]: mu invweight k b dmin dmax width mid power includemargin pad pad */
  const int *cb_body, *cb_chain_start, *cb_chain_dof, *dof_cb_start, *dof_cb, *cb_con_start, *cb_con;
  /* Contact-chain SEGMENTS (use_seg): the root paths of the contact bodies share long prefixes, so J x and J^T f are
   * evaluated on the maximal runs of consecutive chain dofs with the same set of contact bodies below them (10 runs of
   * <= 7 dofs for the rodent instead of 8 chains of 13..18).  Lane tasks t = lane + 32 * slot (two slots), packed words:
   *   seg_task[t]: bits 0-7 first dof, 8-11 length, 12-14 k (spatial component), 15 valid, 16-23 first contact body of the
   *                segment's subtree, 24-27 their count         (task = segment t / 6, component t % 6)
   *   cb_task[t] : bits 0-15 four 4-bit segment ids of the body's root path (0xf = none), 16-18 k, 19 valid,
   *                20-27 first contact of the body, 28-31 contact count   (task = contact body t / 6, component t % 6)
   *   dof_seg3[lane]: segment id (0xff = none) of dofs lane, lane + 32, lane + 64 in bytes 0..2 */
  int use_seg, nseg;
  const uint32_t *seg_task, *cb_task, *dof_seg3;
  // ---- shared-memory layout (float offsets inside one environment's slice)
  int o_qpos, o_qvel, o_act, o_ctrl, o_warm, o_xpos, o_xquat, o_cdof, o_cin, o_big, smem_floats;
  // phase A members of o_big
  int a_xipos, a_anchor, a_axis, a_cvel, a_cdofdot, a_cacc, a_force;
  // phase B members (sparse rows, nMpad floats each): raw inertia M at o_big; factor of M at o_L; factor of
  // M + dt*diag(damping) at o_L2 = o_L + nMpad.  CG: o_L = o_big (factored in place).  Newton keeps the raw M (it is
  // re-used every iteration for H = M + J^T D J and M*search) and factors a copy: o_L = o_big + nMpad, which is also
  // where H is assembled and factored.  f (M-build scratch) aliases the o_L2 block.
  int nMpad, o_L, o_L2;
  // Compact CG layout (l2_spill = 1): o_cin sits directly behind the factor of M (o_L2 = o_cin = o_L + nMpad), so the Euler
  // factor is BUILT over the dead composite-inertia block + a tail block; its first spill_floats floats (the part the solver
  // scratch needs back) are parked in a per-warp global (L2-cache resident) buffer between the factorisation and forward.euler.
  // a_f: M-build scratch (relative to o_big).
  int l2_spill, spill_floats, a_f;
  // solver-phase scratch inside o_cin
  int c_sx, c_sy, c_sD, c_sV, c_sW, c_sWb, c_off, c_t1, c_lf, c_sP, c_end;
};

/* Task-layer constants + packed clip table. */
struct DevTask {
  TmjxTaskConfig cfg;
  int obs_size, ref_obs_size, prop_obs_size, n_rows /* packed body rows per frame */, frame_stride;
  int o_pos, o_quat, o_angvel, o_joints, o_bodies;  // offsets inside a packed frame
  int body_slot[TMJX_MAX_IDX], endeff_slot[TMJX_MAX_IDX];  // index table entry -> packed row
  int body_row[TMJX_MAX_IDX], endeff_row[TMJX_MAX_IDX];    // index table entry -> clamped 67-row index
  int joint_col[TMJX_MAX_IDX];                              // joint_idxs - 1, wrapped + clamped
};

struct HostTables {
  std::vector<int32_t> i32;
  std::vector<uint16_t> u16;
  std::vector<uint8_t> u8;
  std::vector<float> f32;
  DevModel dm{};  // pointers hold OFFSETS until relocate()
  std::vector<int> packed_rows;  // unused here; task layer
};

namespace detail {
template <class T> size_t push(std::vector<T>& pool, const std::vector<T>& v, size_t align = 4) {
  while (pool.size() % align) pool.push_back(T());
  size_t off = pool.size();
  pool.insert(pool.end(), v.begin(), v.end());
  return off;
}
inline int pad4(int n) { return (n + 3) & ~3; }
}  // namespace detail

#define TMJX_OFF(T, off) reinterpret_cast<const T*>(static_cast<uintptr_t>(off))

/* Build all tables. Pointer members of `dm` temporarily hold element offsets into the four pools; the caller
 * uploads the pools and calls relocate(). Throws std::runtime_error for models outside the supported subset. */
inline void build_tables(const Blob& b, const TmjxTaskConfig& cfg, HostTables& t) {
  using detail::pad4;
  using detail::push;
  DevModel& m = t.dm;
  auto dims = b.i32("dims");
  m.nq = dims[0]; m.nv = dims[1]; m.nu = dims[2]; m.na = dims[3]; m.nbody = dims[4]; m.njnt = dims[5]; m.ncg = dims[6];
  m.ncon = dims[7]; m.nefc = dims[8];
  auto opt = b.f32("opt");
  m.dt = cfg.mj_model_timestep;
  m.gravity[0] = opt[1]; m.gravity[1] = opt[2]; m.gravity[2] = opt[3];
  m.tolerance = opt[4]; m.ls_tolerance = opt[5]; m.impratio = opt[6];
  m.meaninertia_scale = opt[7] * float(std::max(1, m.nv));
  m.n_frames = cfg.physics_steps_per_control_step; m.solver = cfg.solver; m.iterations = cfg.iterations;
  m.ls_iterations = cfg.ls_iterations;
  if (m.solver != TMJX_SOLVER_CG && m.solver != TMJX_SOLVER_NEWTON) throw std::runtime_error("unknown solver (unsupported)");
  if (m.nv > 32 * kNvSlots) throw std::runtime_error("nv > 96 unsupported");
  if (m.ncon > kMaxCon) throw std::runtime_error("ncon > 32 unsupported");
  if (m.na != 0 && m.na != m.nu) throw std::runtime_error("na must be 0 or nu");

  const int nbody = m.nbody, njnt = m.njnt, nv = m.nv, nu = m.nu;
  auto body_parent = b.i32("body_parentid"), body_rootid = b.i32("body_rootid"), body_jntadr = b.i32("body_jntadr"),
       body_jntnum = b.i32("body_jntnum"), body_dofadr = b.i32("body_dofadr"), body_dofnum = b.i32("body_dofnum"),
       jnt_type = b.i32("jnt_type"), jnt_qposadr = b.i32("jnt_qposadr"), jnt_dofadr = b.i32("jnt_dofadr"),
       jnt_bodyid = b.i32("jnt_bodyid"), dof_bodyid = b.i32("dof_bodyid"), dof_jntid = b.i32("dof_jntid"),
       dof_parentid = b.i32("dof_parentid"), jnt_limited = b.i32("jnt_limited");
  for (int j = 0; j < njnt; ++j)
    if (jnt_type[j] != kJntFree && jnt_type[j] != kJntHinge) throw std::runtime_error("only free/hinge joints");

  // ---- tree levels and child lists
  std::vector<int> depth(nbody, 0);
  int nlevel = 1;
  for (int i = 1; i < nbody; ++i) { depth[i] = depth[body_parent[i]] + 1; nlevel = std::max(nlevel, depth[i] + 1); }
  std::vector<int32_t> lvl_start(nlevel + 1, 0), lvl_body;
  for (int lv = 0; lv < nlevel; ++lv) {
    lvl_start[lv] = int(lvl_body.size());
    for (int i = 0; i < nbody; ++i) if (depth[i] == lv) lvl_body.push_back(i);
  }
  lvl_start[nlevel] = int(lvl_body.size());
  m.nlevel = nlevel;
  std::vector<int32_t> child_start(nbody + 1, 0), child;
  for (int i = 0; i < nbody; ++i) {
    child_start[i] = int(child.size());
    for (int c = 1; c < nbody; ++c) if (body_parent[c] == i) child.push_back(c);
  }
  child_start[nbody] = int(child.size());

  // ---- ancestor / descendant tables for the doubling scans
  int nround = 0;
  while ((1 << nround) < nlevel) ++nround;
  m.nround = nround;
  std::vector<uint8_t> anc_pow(size_t(std::max(nround, 1)) * nbody, 0), dsc_list;
  std::vector<uint16_t> dsc_start;
  for (int r = 0; r < nround; ++r) {
    for (int i = 0; i < nbody; ++i) {
      int a = i;
      for (int s = 0; s < (1 << r) && a > 0; ++s) a = body_parent[a];
      anc_pow[size_t(r) * nbody + i] = uint8_t(depth[i] >= (1 << r) ? a : 0);
    }
    for (int i = 0; i <= nbody; ++i) {
      dsc_start.push_back(uint16_t(dsc_list.size()));
      if (i == 0 || i == nbody) continue;
      for (int c = i + 1; c < nbody; ++c)
        if (depth[c] - depth[i] == (1 << r) && anc_pow[size_t(r) * nbody + c] == i) dsc_list.push_back(uint8_t(c));
    }
  }
  std::vector<int32_t> dsc_pack(size_t(std::max(nround, 1)) * nbody, 0);
  for (int r = 0; r < nround; ++r)
    for (int i = 0; i < nbody; ++i) {
      const int e0 = dsc_start[size_t(r) * (nbody + 1) + i], e1 = dsc_start[size_t(r) * (nbody + 1) + i + 1], cnt = e1 - e0;
      dsc_pack[size_t(r) * nbody + i] = int32_t((uint32_t(cnt) << 16) | uint32_t(cnt == 1 ? dsc_list[e0] : e0));
    }
  std::vector<int32_t> dsc4(size_t(std::max(nround, 1)) * nbody, 0);
  m.use_dsc4 = (nround <= 8 && nbody <= 96) ? 1 : 0;
  std::memset(m.dsc_maxc, 0, sizeof(m.dsc_maxc));
  for (int r = 0; r < nround && m.use_dsc4; ++r)
    for (int i = 0; i < nbody; ++i) {
      const int e0 = dsc_start[size_t(r) * (nbody + 1) + i], e1 = dsc_start[size_t(r) * (nbody + 1) + i + 1], cnt = e1 - e0;
      if (cnt > 4) { m.use_dsc4 = 0; break; }
      uint32_t pk = 0;
      for (int e = 0; e < cnt; ++e) pk |= uint32_t(dsc_list[e0 + e]) << (8 * e);
      dsc4[size_t(r) * nbody + i] = int32_t(pk);
      m.dsc_maxc[r][i / 32] = uint8_t(std::max<int>(m.dsc_maxc[r][i / 32], cnt));
    }
  if (dsc_start.empty()) dsc_start.push_back(0);
  if (dsc_list.empty()) dsc_list.push_back(0);
  if (nbody > 255) throw std::runtime_error("more than 255 bodies unsupported");

  // ---- moving tree
  int tree_root = -1;
  for (int d = 0; d < nv; ++d) {
    int r = body_rootid[dof_bodyid[d]];
    if (tree_root < 0) tree_root = r;
    if (r != tree_root) throw std::runtime_error("a single kinematic tree is supported");
  }
  m.tree_root = tree_root;
  auto body_mass = b.f32("body_mass");
  std::vector<float> tree_mass_v(nbody, 0.f);
  float tm = 0.f;
  for (int i = 0; i < nbody; ++i)
    if (body_rootid[i] == tree_root) { tree_mass_v[i] = body_mass[i]; tm += body_mass[i]; }
  m.tree_mass = tm;

  // ---- sparse inertia rows
  std::vector<int32_t> dof_madr(nv), dof_depth(nv);
  std::vector<uint8_t> m_anc, m_row, m_col;
  int maxdepth = 0;
  for (int i = 0; i < nv; ++i) {
    dof_madr[i] = int(m_anc.size());
    int c = 0;
    for (int j = i; j >= 0; j = dof_parentid[j]) { m_anc.push_back(uint8_t(j)); m_row.push_back(uint8_t(i)); m_col.push_back(uint8_t(j)); ++c; }
    dof_depth[i] = c - 1;
    maxdepth = std::max(maxdepth, c - 1);
  }
  m.nM = int(m_anc.size());
  m.maxdepth = maxdepth;
  if (m.nM > 1280) throw std::runtime_error("sparse inertia too large (unsupported)");
  if (maxdepth >= 64) throw std::runtime_error("dof chains deeper than 64 unsupported");
  m.use_gen = (nv == gen::kNv) ? 1 : 0;
  for (int i = 0; i < nv && m.use_gen; ++i) if (dof_parentid[i] != gen::kDofParent[i]) m.use_gen = 0;
  std::memset(m.u_ancre, 0, sizeof(m.u_ancre));
  for (int i = 0; i < nv; ++i)
    for (int a = 0; a <= dof_depth[i]; ++a) {
      const int j = m_anc[dof_madr[i] + a];  // ancestor with depth dof_depth[i] - a
      m.u_ancre[dof_madr[i] + a] = uint16_t(dof_madr[j] + dof_depth[j]);
    }
  std::vector<uint8_t> tri_a, tri_b;  // pair p = b(b+1)/2 + a, 0 <= a <= b < maxdepth
  for (int bb = 0; bb < maxdepth; ++bb) for (int a = 0; a <= bb; ++a) { tri_a.push_back(uint8_t(a)); tri_b.push_back(uint8_t(bb)); }
  std::vector<int32_t> desc_start(nv + 1, 0);
  std::vector<uint8_t> desc_dof;
  std::vector<uint16_t> desc_off;
  for (int j = 0; j < nv; ++j) {
    desc_start[j] = int(desc_dof.size());
    for (int i = j + 1; i < nv; ++i)
      for (int a = 1; a <= dof_depth[i]; ++a)
        if (m_anc[dof_madr[i] + a] == j) { desc_dof.push_back(uint8_t(i)); desc_off.push_back(uint16_t(dof_madr[i] + a)); }
  }
  desc_start[nv] = int(desc_dof.size());

  // ---- tables for the shuffle-based solves and the factorisation
  std::vector<int32_t> dmask(9 * 32, 0), amask(9 * 32, 0), rowend_me(3 * 32, 0), depth_me(3 * 32, 0);
  std::memset(m.u_rowend, 0, sizeof(m.u_rowend));
  std::memset(m.u_depth, 0, sizeof(m.u_depth));
  std::memset(m.slot_used, 0, sizeof(m.slot_used));
  for (int i = 0; i < nv; ++i) {
    m.u_rowend[i] = uint16_t(dof_madr[i] + dof_depth[i]);
    m.u_depth[i] = uint8_t(dof_depth[i]);
    rowend_me[(i / 32) * 32 + (i % 32)] = dof_madr[i] + dof_depth[i];
    depth_me[(i / 32) * 32 + (i % 32)] = dof_depth[i];
    for (int a = 1; a <= dof_depth[i]; ++a) {
      const int j = m_anc[dof_madr[i] + a];  // j is an ancestor of i
      dmask[((j / 32) * 3 + (i / 32)) * 32 + (j % 32)] |= int32_t(1u << (i % 32));
      amask[((i / 32) * 3 + (j / 32)) * 32 + (i % 32)] |= int32_t(1u << (j % 32));
      m.slot_used[j / 32][i / 32] = 1;
    }
  }
  std::vector<int32_t> pair_tab, pair_start(nv + 1, 0);
  for (int k = 0; k < nv; ++k) {
    pair_start[k] = int(pair_tab.size());
    const int c = dof_depth[k];
    for (int bb = 1; bb <= c; ++bb)
      for (int a = 1; a <= bb; ++a) {
        const int tgt = dof_madr[m_anc[dof_madr[k] + a]] + (bb - a);
        pair_tab.push_back(int32_t(uint32_t(a) | (uint32_t(bb) << 8) | (uint32_t(tgt) << 16)));
      }
  }
  pair_start[nv] = int(pair_tab.size());

  // ---- per dof helpers
  std::vector<int32_t> dof_qadr(nv, -1), dof_limit(nv, -1);
  for (int j = 0; j < njnt; ++j) if (jnt_type[j] == kJntHinge) dof_qadr[jnt_dofadr[j]] = jnt_qposadr[j];

  // ---- limits
  auto jnt_range = b.f32("jnt_range"), jnt_margin = b.f32("jnt_margin"), jnt_solref = b.f32("jnt_solref"),
       jnt_solimp = b.f32("jnt_solimp"), dof_invweight0 = b.f32("dof_invweight0");
  auto kb = [&](const float* solref, const float* solimp, float* out /* k b dmin dmax width mid power */) {
    float timeconst = std::max(solref[0], 2.f * m.dt), dampratio = solref[1];
    float dmin = std::min(std::max(solimp[0], 0.0001f), 0.9999f), dmax = std::min(std::max(solimp[1], 0.0001f), 0.9999f);
    float width = std::max(1e-15f, solimp[2]), mid = std::min(std::max(solimp[3], 0.0001f), 0.9999f);
    float power = std::max(1.f, solimp[4]);
    float k = 1.f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio), bb = 2.f / (dmax * timeconst);
    if (solref[0] <= 0) k = -solref[0] / (dmax * dmax);
    if (solref[1] <= 0) bb = -solref[1] / dmax;
    out[0] = k; out[1] = bb; out[2] = dmin; out[3] = dmax; out[4] = width; out[5] = mid; out[6] = power;
  };
  std::vector<int32_t> lim_dof, lim_qadr;
  std::vector<float> lim_par;
  for (int j = 0; j < njnt; ++j) {
    if (!jnt_limited[j] || jnt_type[j] != kJntHinge) continue;
    dof_limit[jnt_dofadr[j]] = int(lim_dof.size());
    lim_dof.push_back(jnt_dofadr[j]);
    lim_qadr.push_back(jnt_qposadr[j]);
    float p[12] = {jnt_range[j * 2], jnt_range[j * 2 + 1], jnt_margin[j], dof_invweight0[jnt_dofadr[j]]};
    kb(&jnt_solref[j * 2], &jnt_solimp[j * 5], p + 4);
    lim_par.insert(lim_par.end(), p, p + 12);
  }
  m.nlimit = int(lim_dof.size());
  if (m.nlimit > 32 * kLimSlots) throw std::runtime_error("more than 96 joint limits unsupported");
  if (m.nlimit + 4 * m.ncon != m.nefc) throw std::runtime_error("nefc mismatch");

  // ---- actuators
  auto act_moment = b.f32("actuator_moment"), act_gain = b.f32("actuator_gain"), act_biasprm = b.f32("actuator_biasprm"),
       act_dynprm = b.f32("actuator_dynprm"), act_ctrlrange = b.f32("actuator_ctrlrange"),
       act_forcerange = b.f32("actuator_forcerange");
  auto act_ctrllimited = b.i32("actuator_ctrllimited"), act_forcelimited = b.i32("actuator_forcelimited"),
       act_affine = b.i32("actuator_bias_affine"), act_filter = b.i32("actuator_dyn_filter");
  std::vector<float> a_lo(nu), a_hi(nu), a_dyninv(nu), a_flo(nu), a_fhi(nu);
  std::vector<int32_t> a_flags(nu);
  for (int u = 0; u < nu; ++u) {
    a_lo[u] = act_ctrlrange[u * 2]; a_hi[u] = act_ctrlrange[u * 2 + 1];
    a_flo[u] = act_forcerange[u * 2]; a_fhi[u] = act_forcerange[u * 2 + 1];
    a_dyninv[u] = std::max(act_dynprm[u], 1e-15f);  // the kernel divides, like MJX
    a_flags[u] = (act_ctrllimited[u] ? 1 : 0) | (act_filter[u] ? 2 : 0) | (act_affine[u] ? 4 : 0) | (act_forcelimited[u] ? 8 : 0);
    for (int d = 0; d < 6 && d < nv; ++d)
      if (act_moment[size_t(u) * nv + d] != 0.f) throw std::runtime_error("actuators on the free joint unsupported");
  }
  std::vector<int32_t> dof_act_start(nv + 1, 0), dof_act_id, act_mom_start(nu + 1, 0), act_mom_dof;
  std::vector<float> dof_act_coef, act_mom_coef;
  for (int d = 0; d < nv; ++d) {
    dof_act_start[d] = int(dof_act_id.size());
    for (int u = 0; u < nu; ++u)
      if (act_moment[size_t(u) * nv + d] != 0.f) { dof_act_id.push_back(u); dof_act_coef.push_back(act_moment[size_t(u) * nv + d]); }
  }
  dof_act_start[nv] = int(dof_act_id.size());
  for (int u = 0; u < nu; ++u) {
    act_mom_start[u] = int(act_mom_dof.size());
    for (int d = 0; d < nv; ++d)
      if (act_moment[size_t(u) * nv + d] != 0.f) { act_mom_dof.push_back(d); act_mom_coef.push_back(act_moment[size_t(u) * nv + d]); }
  }
  act_mom_start[nu] = int(act_mom_dof.size());

  // ---- contacts
  auto cg_type = b.i32("cgeom_type"), cg_body = b.i32("cgeom_bodyid"), pair_cgeom = b.i32("pair_cgeom");
  auto pair_friction = b.f32("pair_friction"), pair_solref = b.f32("pair_solref"), pair_solimp = b.f32("pair_solimp"),
       pair_margin = b.f32("pair_includemargin"), body_invweight0 = b.f32("body_invweight0"), plane = b.f32("plane");
  const int plane_body = b.i32("plane_bodyid")[0];
  for (int k = 0; k < 3; ++k) { m.plane_pos[k] = plane[k]; m.plane_n[k] = plane[3 + k]; }
  std::vector<int32_t> cb_body, cg_cb(m.ncg);
  for (int g = 0; g < m.ncg; ++g) cb_body.push_back(cg_body[g]);
  std::sort(cb_body.begin(), cb_body.end());   // body ids are depth-first: a subtree's contact bodies are one range
  cb_body.erase(std::unique(cb_body.begin(), cb_body.end()), cb_body.end());
  for (int g = 0; g < m.ncg; ++g) cg_cb[g] = int(std::lower_bound(cb_body.begin(), cb_body.end(), cg_body[g]) - cb_body.begin());
  m.ncb = int(cb_body.size());
  std::vector<int32_t> con_geom, con_side, con_cb;
  std::vector<float> con_par;
  for (size_t p = 0; p < pair_cgeom.size(); ++p) {
    const int g = pair_cgeom[p];
    const int n = cg_type[g] == kGeomCapsule ? 2 : 1;
    const float mu = pair_friction[p * 5];
    if (pair_friction[p * 5 + 1] != mu) throw std::runtime_error("anisotropic tangential friction unsupported");
    const float tw = body_invweight0[plane_body * 2] + body_invweight0[cg_body[g] * 2];
    for (int s = 0; s < n; ++s) {
      con_geom.push_back(g);
      con_side.push_back(n == 2 ? (s == 0 ? 1 : -1) : 0);
      con_cb.push_back(cg_cb[g]);
      float q[12] = {mu, (tw + mu * mu * tw) * 2.f * mu * mu / m.impratio};
      kb(&pair_solref[p * 2], &pair_solimp[p * 5], q + 2);
      q[9] = pair_margin[p];
      q[10] = q[11] = 0.f;
      con_par.insert(con_par.end(), q, q + 12);
    }
  }
  if (int(con_geom.size()) != m.ncon) throw std::runtime_error("ncon mismatch");
  std::vector<int32_t> cb_chain_start(m.ncb + 1, 0), cb_chain_dof, dof_cb_start(nv + 1, 0), dof_cb, cb_con_start(m.ncb + 1, 0), cb_con;
  std::vector<std::vector<int>> dof_cb_l(nv);
  for (int s = 0; s < m.ncb; ++s) {
    cb_chain_start[s] = int(cb_chain_dof.size());
    int bb = cb_body[s];
    while (bb > 0 && body_dofnum[bb] == 0) bb = body_parent[bb];
    std::vector<int> chain;
    if (bb > 0)
      for (int i = body_dofadr[bb] + body_dofnum[bb] - 1; i >= 0; i = dof_parentid[i]) chain.push_back(i);
    for (auto it = chain.rbegin(); it != chain.rend(); ++it) { cb_chain_dof.push_back(*it); dof_cb_l[*it].push_back(s); }
    cb_con_start[s] = int(cb_con.size());
    for (int c = 0; c < m.ncon; ++c) if (con_cb[c] == s) cb_con.push_back(c);
  }
  cb_chain_start[m.ncb] = int(cb_chain_dof.size());
  cb_con_start[m.ncb] = int(cb_con.size());
  for (int d = 0; d < nv; ++d) {
    dof_cb_start[d] = int(dof_cb.size());
    for (int s : dof_cb_l[d]) dof_cb.push_back(s);
  }
  dof_cb_start[nv] = int(dof_cb.size());
  // ---- contact-chain segments
  std::vector<int32_t> seg_task(64, 0), cb_task(64, 0), dof_seg3(32, 0x00ffffff);
  m.use_seg = 0; m.nseg = 0;
  {
    std::vector<int> seg_d0, seg_len, seg_of(nv, -1), seg_cb0, seg_cbn;
    bool ok = m.ncb * 6 <= 64 && m.ncb <= 32;
    std::vector<uint32_t> cset(nv, 0);
    if (ok) for (int d = 0; d < nv; ++d) for (int s2 : dof_cb_l[d]) cset[d] |= 1u << s2;
    for (int d = 0; d < nv && ok; ++d) {
      if (!cset[d]) continue;
      const bool cont = d > 0 && dof_parentid[d] == d - 1 && cset[d - 1] == cset[d] && seg_len.back() < 8;
      if (!cont) {
        seg_d0.push_back(d); seg_len.push_back(0);
        int lo = 0; while (!((cset[d] >> lo) & 1u)) ++lo;
        int n = 0; while (lo + n < 32 && ((cset[d] >> (lo + n)) & 1u)) ++n;
        if (cset[d] != (((n == 32 ? 0u : (1u << n)) - 1u) << lo)) ok = false;   // subtree bodies must be one range
        if (n > 15) ok = false;
        seg_cb0.push_back(lo); seg_cbn.push_back(n);
      }
      ++seg_len.back();
      seg_of[d] = int(seg_d0.size()) - 1;
    }
    const int nseg = int(seg_d0.size());
    if (nseg * 6 > 64 || nseg > 15 || nv > 96) ok = false;
    for (int s2 = 0; s2 < m.ncb && ok; ++s2) {
      // root path of the body in segments, contact range of the body
      std::vector<int> path;
      for (int e = cb_chain_start[s2 + 1] - 1; e >= cb_chain_start[s2]; --e) {
        const int sg = seg_of[cb_chain_dof[e]];
        if (path.empty() || path.back() != sg) path.push_back(sg);
      }
      const int c0 = cb_con_start[s2], cn = cb_con_start[s2 + 1] - c0;
      for (int e = 0; e < cn; ++e) if (cb_con[c0 + e] != cb_con[c0] + e) ok = false;   // contiguous contacts
      if (path.size() > 4 || cn > 15 || (cn > 0 && cb_con[c0] > 255)) ok = false;
      if (!ok) break;
      uint32_t pk = 0;
      for (int j = 0; j < 4; ++j) pk |= uint32_t(j < int(path.size()) ? path[j] : 0xf) << (4 * j);
      for (int k = 0; k < 6; ++k)
        cb_task[s2 * 6 + k] = int32_t(pk | (uint32_t(k) << 16) | (1u << 19) | (uint32_t(cn ? cb_con[c0] : 0) << 20) | (uint32_t(cn) << 28));
    }
    if (ok) {
      for (int sg = 0; sg < nseg; ++sg)
        for (int k = 0; k < 6; ++k)
          seg_task[sg * 6 + k] = int32_t(uint32_t(seg_d0[sg]) | (uint32_t(seg_len[sg]) << 8) | (uint32_t(k) << 12) | (1u << 15) |
                                         (uint32_t(seg_cb0[sg]) << 16) | (uint32_t(seg_cbn[sg]) << 24));
      for (int l = 0; l < 32; ++l) {
        uint32_t w3 = 0;
        for (int q = 0; q < 3; ++q) { const int d = l + 32 * q; w3 |= uint32_t(d < nv && seg_of[d] >= 0 ? seg_of[d] : 0xff) << (8 * q); }
        dof_seg3[l] = int32_t(w3);
      }
      m.use_seg = 1; m.nseg = nseg;
    }
  }

  // ---- per joint floats
  auto qpos0 = b.f32("qpos0"), qpos_spring = b.f32("qpos_spring");
  std::vector<float> jnt_qpos0(njnt), jnt_springref(njnt);
  for (int j = 0; j < njnt; ++j) { jnt_qpos0[j] = qpos0[jnt_qposadr[j]]; jnt_springref[j] = qpos_spring[jnt_qposadr[j]]; }

  // ---- pools (pointer members hold offsets for now)
#define PI(field, vec) m.field = TMJX_OFF(int, push(t.i32, vec))
#define PF(field, vec) m.field = TMJX_OFF(float, push(t.f32, vec))
#define P8(field, vec) m.field = TMJX_OFF(uint8_t, push(t.u8, vec))
  PI(body_parent, body_parent); PI(body_jntadr, body_jntadr); PI(body_jntnum, body_jntnum); PI(body_dofadr, body_dofadr);
  PI(body_dofnum, body_dofnum); PI(lvl_start, lvl_start); PI(lvl_body, lvl_body); PI(child_start, child_start); PI(child, child);
  PF(body_pos, b.f32("body_pos")); PF(body_quat, b.f32("body_quat")); PF(body_ipos, b.f32("body_ipos"));
  PF(body_iquat, b.f32("body_iquat")); PF(body_inertia, b.f32("body_inertia")); PF(body_mass, body_mass);
  PF(body_tree_mass, tree_mass_v);
  PI(jnt_type, jnt_type); PI(jnt_qposadr, jnt_qposadr); PI(jnt_dofadr, jnt_dofadr); PI(jnt_body, jnt_bodyid);
  PF(jnt_pos, b.f32("jnt_pos")); PF(jnt_axis, b.f32("jnt_axis")); PF(jnt_stiffness, b.f32("jnt_stiffness"));
  PF(jnt_qpos0, jnt_qpos0); PF(jnt_springref, jnt_springref);
  PI(dof_body, dof_bodyid); PI(dof_jnt, dof_jntid); PI(dof_madr, dof_madr); PI(dof_depth, dof_depth); PI(dof_limit, dof_limit);
  PI(dof_qadr, dof_qadr);
  PF(dof_armature, b.f32("dof_armature")); PF(dof_damping, b.f32("dof_damping"));
  P8(anc_pow, anc_pow); P8(dsc_list, dsc_list);
  m.dsc_pack = TMJX_OFF(uint32_t, push(t.i32, dsc_pack));
  m.dsc4 = TMJX_OFF(uint32_t, push(t.i32, dsc4));
  {
    if (m.nM > 1280) throw std::runtime_error("more than 1280 inertia entries unsupported");
    std::vector<int32_t> rc2(32 * 20, 0);
    for (int e = 0; e < m.nM; ++e) {
      const uint32_t rc = uint32_t(m_row[e]) | (uint32_t(m_col[e]) << 8);
      const int lane = e & 31, it = e >> 5;
      rc2[lane + 32 * (it >> 1)] |= int32_t(rc << (16 * (it & 1)));
    }
    m.m_rc2 = TMJX_OFF(uint32_t, push(t.i32, rc2));
  }
  m.dsc_start = TMJX_OFF(uint16_t, push(t.u16, dsc_start));
  P8(m_anc, m_anc); P8(m_row, m_row); P8(m_col, m_col); P8(tri_a, tri_a); P8(tri_b, tri_b);
  m.dmask = TMJX_OFF(uint32_t, push(t.i32, dmask)); m.amask = TMJX_OFF(uint32_t, push(t.i32, amask));
  PI(rowend_me, rowend_me); PI(depth_me, depth_me);
  m.pair_tab = TMJX_OFF(uint32_t, push(t.i32, pair_tab)); PI(pair_start, pair_start);
  PI(desc_start, desc_start); P8(desc_dof, desc_dof);
  m.desc_off = TMJX_OFF(uint16_t, push(t.u16, desc_off));
  PF(act_gain, act_gain); PF(act_ctrl_lo, a_lo); PF(act_ctrl_hi, a_hi); PF(act_dyn_inv, a_dyninv); PF(act_bias, act_biasprm);
  PF(act_force_lo, a_flo); PF(act_force_hi, a_fhi); PI(act_flags, a_flags);
  PI(dof_act_start, dof_act_start); PI(dof_act_id, dof_act_id); PF(dof_act_coef, dof_act_coef);
  PI(act_mom_start, act_mom_start); PI(act_mom_dof, act_mom_dof); PF(act_mom_coef, act_mom_coef);
  PI(lim_dof, lim_dof); PI(lim_qadr, lim_qadr); PF(lim_par, lim_par);
  PI(cg_type, cg_type); PI(cg_body, cg_body); PI(cg_cb, cg_cb);
  PF(cg_pos, b.f32("cgeom_pos")); PF(cg_quat, b.f32("cgeom_quat")); PF(cg_size, b.f32("cgeom_size"));
  PI(con_geom, con_geom); PI(con_side, con_side); PI(con_cb, con_cb); PF(con_par, con_par);
  PI(cb_body, cb_body); PI(cb_chain_start, cb_chain_start); PI(cb_chain_dof, cb_chain_dof); PI(dof_cb_start, dof_cb_start);
  PI(dof_cb, dof_cb); PI(cb_con_start, cb_con_start); PI(cb_con, cb_con);
  m.seg_task = TMJX_OFF(uint32_t, push(t.i32, seg_task)); m.cb_task = TMJX_OFF(uint32_t, push(t.i32, cb_task));
  m.dof_seg3 = TMJX_OFF(uint32_t, push(t.i32, dof_seg3));
#undef PI
#undef PF
#undef P8

  // ---- shared-memory layout
  int o = 0;
  auto take = [&](int n) { int r = o; o += pad4(n); return r; };
  m.o_qpos = take(m.nq); m.o_qvel = take(nv); m.o_act = take(std::max(m.na, 1)); m.o_ctrl = take(nu); m.o_warm = take(nv);
  m.o_xpos = take(nbody * 3); m.o_xquat = take(nbody * 4); m.o_cdof = take(nv * 6);
  m.nMpad = pad4(m.nM);
  int cin_size = 0;
  {
    int c = 0;
    auto tk = [&](int n) { int r = c; c += pad4(n); return r; };
    m.c_sx = tk(nv); m.c_sy = m.c_sD = 0; m.c_sV = tk(m.ncb * 6); m.c_sW = tk(m.ncon * 6); m.c_sWb = tk(m.ncb * 6);
    m.c_off = tk(m.ncon * 3); m.c_t1 = tk(m.ncon * 3); m.c_lf = tk(m.nlimit);
    m.c_sP = tk(64);
    m.c_end = c;
    cin_size = std::max(pad4(nbody * 10), c);
  }
  if (pad4(nv * 6) > m.nMpad) throw std::runtime_error("M-build scratch does not fit");
  m.l2_spill = (m.solver == TMJX_SOLVER_CG) ? 1 : 0;
  if (const char* e = std::getenv("TMJX_L2_SPILL")) m.l2_spill = (atoi(e) != 0 && m.solver == TMJX_SOLVER_CG) ? 1 : 0;   // tuning knob
  m.spill_floats = 0;
  if (m.l2_spill) {
    // [ L = factor of M (nMpad) | cin (cin_size) | tail ]   with the Euler factor L2 over cin + tail.
    // Phase A transients live in the L block and the tail (both dead until build_m writes the matrices):
    //   kinematics -> com_pos : anchor, axis (aliased later by cvel), xipos          (L block)
    //   com_vel_rne           : cvel, cacc (L block), cdof_dot (tail)
    //   passive_actuation     : force (L block);   build_m: f (tail, after cdof_dot died)
    m.o_big = o;
    m.o_L = m.o_big; m.o_cin = m.o_big + m.nMpad; m.o_L2 = m.o_cin;
    const int tail = std::max(std::max(m.nMpad - cin_size, 0), pad4(nv * 6));
    const int o_tail = m.nMpad + cin_size;       // relative to o_big
    int a = 0;
    auto tk = [&](int n) { int r = a; a += pad4(n); return r; };
    m.a_cvel = tk(nbody * 6); m.a_anchor = m.a_cvel; m.a_axis = m.a_anchor + pad4(njnt * 3);
    if (2 * pad4(njnt * 3) > pad4(nbody * 6)) a = m.a_anchor + 2 * pad4(njnt * 3);
    m.a_cacc = tk(nbody * 6); m.a_xipos = tk(nbody * 3); m.a_force = tk(nu);
    if (a > m.nMpad) throw std::runtime_error("smooth-dynamics transients do not fit the compact layout (unsupported)");
    m.a_cdofdot = o_tail; m.a_f = o_tail;
    m.spill_floats = cin_size;
    o += m.nMpad + cin_size + tail;
  } else {
    m.o_cin = o;
    o += cin_size;
    m.o_big = o;
    int a = 0;
    auto tk = [&](int n) { int r = a; a += pad4(n); return r; };
    m.a_xipos = tk(nbody * 3); m.a_anchor = tk(njnt * 3); m.a_axis = tk(njnt * 3); m.a_cvel = tk(nbody * 6);
    m.a_cdofdot = tk(nv * 6); m.a_cacc = tk(nbody * 6); m.a_force = tk(nu);
    const int nmat = m.solver == TMJX_SOLVER_NEWTON ? 3 : 2;
    m.o_L = m.o_big + (nmat - 2) * m.nMpad;
    m.o_L2 = m.o_L + m.nMpad;
    m.a_f = m.o_L2 - m.o_big;     // M-build scratch aliases the (not yet written) Euler-factor block
    o += std::max(a, nmat * m.nMpad);
  }
  if (m.solver == TMJX_SOLVER_NEWTON)
    for (int s = 0; s < m.ncb; ++s)
      if (cb_chain_start[s + 1] - cb_chain_start[s] > 32) throw std::runtime_error("Newton: contact chains longer than 32 dofs unsupported");
  m.smem_floats = o;
}

/* Turn the offsets stored in the pointer members into device pointers. */
inline void relocate(DevModel& m, const int* di, const uint16_t* d16, const uint8_t* d8, const float* df) {
#define RI(f) m.f = di + reinterpret_cast<uintptr_t>(m.f)
#define RF(f) m.f = df + reinterpret_cast<uintptr_t>(m.f)
#define R8(f) m.f = d8 + reinterpret_cast<uintptr_t>(m.f)
  RI(body_parent); RI(body_jntadr); RI(body_jntnum); RI(body_dofadr); RI(body_dofnum); RI(lvl_start); RI(lvl_body);
  RI(child_start); RI(child);
  RF(body_pos); RF(body_quat); RF(body_ipos); RF(body_iquat); RF(body_inertia); RF(body_mass); RF(body_tree_mass);
  RI(jnt_type); RI(jnt_qposadr); RI(jnt_dofadr); RI(jnt_body);
  RF(jnt_pos); RF(jnt_axis); RF(jnt_stiffness); RF(jnt_qpos0); RF(jnt_springref);
  RI(dof_body); RI(dof_jnt); RI(dof_madr); RI(dof_depth); RI(dof_limit); RI(dof_qadr);
  RF(dof_armature); RF(dof_damping);
  R8(anc_pow); R8(dsc_list);
  m.dsc_pack = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.dsc_pack);
  m.dsc4 = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.dsc4);
  m.seg_task = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.seg_task);
  m.cb_task = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.cb_task);
  m.dof_seg3 = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.dof_seg3);
  m.m_rc2 = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.m_rc2);
  m.dsc_start = d16 + reinterpret_cast<uintptr_t>(m.dsc_start);
  R8(m_anc); R8(m_row); R8(m_col); R8(tri_a); R8(tri_b);
  m.dmask = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.dmask);
  m.amask = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.amask);
  m.pair_tab = reinterpret_cast<const uint32_t*>(di) + reinterpret_cast<uintptr_t>(m.pair_tab);
  RI(rowend_me); RI(depth_me); RI(pair_start);
  RI(desc_start); R8(desc_dof);
  m.desc_off = d16 + reinterpret_cast<uintptr_t>(m.desc_off);
  RF(act_gain); RF(act_ctrl_lo); RF(act_ctrl_hi); RF(act_dyn_inv); RF(act_bias); RF(act_force_lo); RF(act_force_hi);
  RI(act_flags); RI(dof_act_start); RI(dof_act_id); RF(dof_act_coef); RI(act_mom_start); RI(act_mom_dof); RF(act_mom_coef);
  RI(lim_dof); RI(lim_qadr); RF(lim_par);
  RI(cg_type); RI(cg_body); RI(cg_cb); RF(cg_pos); RF(cg_quat); RF(cg_size);
  RI(con_geom); RI(con_side); RI(con_cb); RF(con_par);
  RI(cb_body); RI(cb_chain_start); RI(cb_chain_dof); RI(dof_cb_start); RI(dof_cb); RI(cb_con_start); RI(cb_con);
#undef RI
#undef RF
#undef R8
}

/* Task constants: obs sizes, packed clip-frame layout (the subset of ReferenceClip the step reads). */
inline void build_task(const DevModel& m, const TmjxTaskConfig& cfg, int n_ref_bodies, DevTask& t) {
  t.cfg = cfg;
  const int nj = m.nq - 7;
  t.ref_obs_size = cfg.traj_length * (3 + 4 + cfg.n_joint_idxs + 3 * cfg.n_body_idxs);
  t.prop_obs_size = (m.nq - 7) + (m.nv - 6) + m.nv + 1 + 3 + 3 * cfg.n_appendages;
  t.obs_size = t.ref_obs_size + t.prop_obs_size;
  std::vector<int> rows;
  auto slot_of = [&](int id) {
    const int row = std::min(std::max(id, 0), n_ref_bodies - 1);  // jnp gather clamps
    for (size_t s = 0; s < rows.size(); ++s) if (rows[s] == row) return int(s);
    rows.push_back(row);
    return int(rows.size()) - 1;
  };
  for (int i = 0; i < cfg.n_body_idxs; ++i) { t.body_slot[i] = slot_of(cfg.body_idxs[i]); t.body_row[i] = rows[t.body_slot[i]]; }
  for (int i = 0; i < cfg.n_endeff_idxs; ++i) { t.endeff_slot[i] = slot_of(cfg.endeff_idxs[i]); t.endeff_row[i] = rows[t.endeff_slot[i]]; }
  for (int i = 0; i < cfg.n_joint_idxs; ++i) {
    int col = cfg.joint_idxs[i] - 1;
    if (col < 0) col += nj;
    t.joint_col[i] = std::min(std::max(col, 0), nj - 1);
  }
  t.n_rows = int(rows.size());
  t.o_pos = 0; t.o_quat = 4; t.o_angvel = 8; t.o_joints = 12; t.o_bodies = 12 + detail::pad4(nj);
  t.frame_stride = detail::pad4(t.o_bodies + 3 * t.n_rows);
}
inline std::vector<int> task_rows(const DevTask& t) {
  std::vector<int> rows(t.n_rows, 0);
  for (int i = 0; i < t.cfg.n_body_idxs; ++i) rows[t.body_slot[i]] = t.body_row[i];
  for (int i = 0; i < t.cfg.n_endeff_idxs; ++i) rows[t.endeff_slot[i]] = t.endeff_row[i];
  return rows;
}

}  // namespace tmjx
#endif  // TMJX_TABLES_H_
