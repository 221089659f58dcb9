"""Generate the tree-specialised device code of the step kernel: csrc/tmjx_gen_tree.cuh.

    python tools/gen_tree_kernels.py            # rewrites track-mjx_b200/csrc/tmjx_gen_tree.cuh (committed)

The joint-space inertia of a kinematic tree factors as M = L^T D L with L unit lower triangular and non-zero only at
(dof, ancestor) (Featherstone; MuJoCo mj_factorM / mj_solveLD, which mjx.smooth.factor_m / solve_m restate in dense
form).  The triangular solves are the latency chain of every physics substep (8 per substep), so for the walker's
fixed tree they are emitted as STRAIGHT-LINE warp code: the right-hand side lives in registers (lane l owns dofs
l, l+32, l+64), each pivot value travels by one `__shfl_sync` with a compile-time source lane, ancestor / descendant
tests are compile-time lane masks, L is read from shared memory with immediate offsets.  Pivots are emitted level by
level (deepest first for L^T, root first for L) so that independent limbs interleave and the scheduler sees the ILP.

The same schedule is held as a tiny IR and executed by a numpy interpreter (`run_ir`), which is how
tests/test_gen_schedule.py validates it on the CPU against a dense solve before any GPU time is spent.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "track-mjx_b200", "csrc", "tmjx_gen_tree.cuh")


# ------------------------------------------------------------------------------------------------ tree structure
class Tree:
    def __init__(self, dof_parent):
        p = [int(x) for x in dof_parent]
        self.parent = p
        self.nv = nv = len(p)
        assert nv <= 96
        self.depth = [0] * nv
        for i in range(nv):
            assert p[i] < i
            self.depth[i] = 0 if p[i] < 0 else self.depth[p[i]] + 1
        self.madr, o = [], 0
        for i in range(nv):
            self.madr.append(o)
            o += self.depth[i] + 1
        self.nM = o
        self.rowend = [self.madr[i] + self.depth[i] for i in range(nv)]
        self.anc = []
        for i in range(nv):
            a, j = [], p[i]
            while j >= 0:
                a.append(j)
                j = p[j]
            self.anc.append(a)  # nearest first
        self.desc = [[] for _ in range(nv)]
        for i in range(nv):
            for j in self.anc[i]:
                self.desc[j].append(i)
        self.maxdepth = max(self.depth)

    def signature(self) -> str:
        return hashlib.sha256(bytes(np.asarray(self.parent, np.int16).tobytes())).hexdigest()[:16]


def lane_mask(dofs, slot):
    m = 0
    for d in dofs:
        if d // 32 == slot:
            m |= 1 << (d % 32)
    return m


# ------------------------------------------------------------------------------------------------ IR
# ("shfl", dst, src, lane)                    dst = src[lane]
# ("fnma_lds", x, ptr, imm, p, mask)          x = x - L[ptr(lane) + imm] * p     for lanes in mask
# ("mul_lds", x, ptr, imm, mask)              x = x * L[ptr(lane) + imm]         for lanes in mask
# pointers (per lane, in floats): "U<t>" = -depth_me[t], "D<t>" = rowend_me[t], "G<t>" = madr_me[t]
def build_solve_ir(t: Tree):
    nv = t.nv
    nslot = (nv + 31) // 32
    ir = []
    by_level = [[i for i in range(nv) if t.depth[i] == lv] for lv in range(t.maxdepth + 1)]
    # Pivots of one tree level are independent of each other (none is an ancestor of another), so a level first broadcasts ALL of
    # its pivot values (independent shuffles, issued back to back: one shuffle latency per level instead of one per pivot -- with a
    # single pivot register every shuffle had to wait for the predicated FFMA that last wrote its source register, a false
    # dependency) and then applies the updates.  The critical path is the deepest chain (36 levels), not the 73 pivots.
    # ---- x <- L^-T x: pivot i (deepest level first) updates its ancestors: x[j] -= L[i][j] x[i]
    for lv in range(t.maxdepth, 0, -1):
        piv = sorted(by_level[lv], reverse=True)
        for k, i in enumerate(piv):
            ir.append(("shfl", f"p{k}", f"x{i // 32}", i % 32))
        for k, i in enumerate(piv):
            for s in range(nslot):
                m = lane_mask(t.anc[i], s)
                if m:
                    ir.append(("fnma_lds", f"x{s}", f"U{s}", t.rowend[i], f"p{k}", m))
    # ---- x <- D^-1 x (the factorisation leaves 1/D on the diagonal)
    for s in range(nslot):
        ir.append(("mul_lds", f"x{s}", f"G{s}", 0, lane_mask(range(nv), s)))
    # ---- x <- L^-1 x: pivot j (root first) updates its descendants: x[i] -= L[i][j] x[j]
    for lv in range(0, t.maxdepth):
        piv = [j for j in by_level[lv] if t.desc[j]]
        for k, j in enumerate(piv):
            ir.append(("shfl", f"p{k}", f"x{j // 32}", j % 32))
        for k, j in enumerate(piv):
            for s in range(nslot):
                m = lane_mask(t.desc[j], s)
                if m:
                    ir.append(("fnma_lds", f"x{s}", f"D{s}", -t.depth[j], f"p{k}", m))
    return ir


# ------------------------------------------------------------------------------------------------ factorisation IR
# Stacked dual L^T D L: half-warp h = lane >> 4 owns matrix h (stored at L + h * off2), lane i = lane & 15 owns the
# entries whose column dof has depth 16 j + i (column block j).  Per-lane pointer ps = L + h * off2 - i, so entry
# (row r, column depth 16 j + i) is ps[rowend(r) - 16 j].
#   ("ld", reg, imm, mask)          reg = ps[imm] on lanes in mask, 0 elsewhere
#   ("shflh", dst, src, l15)        dst = src[l15 | (lane & 16)]           (half-warp broadcast)
#   ("rcp", dst, src) / ("mul", dst, a, b)
#   ("st", imm, reg, mask)          ps[imm] = reg on lanes in mask
#   ("fnma_reg", dst, a, w, mask)   dst = dst - a * w on lanes in mask (dst: a row register)
#   ("sync",)                       block phase barrier (instruction-cache lock-step)
def half_mask(n):
    """lanes i < n of both half-warps (n in 0..16)."""
    # (the row updates of the factorisation run unmasked, see build_factor_ir)
    n = max(0, min(16, n))
    m16 = (1 << n) - 1
    return m16 | (m16 << 16)


def chains_descending(t: Tree):
    """Maximal runs k, k-1, ... with parent(k) == k - 1, listed from the highest dof id down (every descendant chain of a
    chain comes before it, which is all the elimination order needs)."""
    out, k = [], t.nv - 1
    while k >= 0:
        run = [k]
        while t.parent[run[-1]] == run[-1] - 1 and run[-1] - 1 >= 0:
            run.append(run[-1] - 1)
        out.append(run)
        k = run[-1] - 1
    return out


K_BATCH = 6      # pivot-row broadcasts in flight per batch of the factorisation


def build_factor_ir(t: Tree):
    """Chain-at-a-time elimination: the rows a chain's pivots touch (the chain itself + the ancestors of its top end) are
    loaded into registers once, every pivot of the chain then updates them with shuffle + FMA only, and they are stored
    back once -- instead of one shared-memory load/store per (pivot, ancestor row)."""
    ir = []
    npiv = 0
    for run in chains_descending(t):
        bottom = run[-1]
        rows = list(run) + t.anc[bottom]          # every row this chain reads or updates
        ir.append(("sync",))
        for r in rows:
            for j in range(t.depth[r] // 16 + 1):
                ir.append(("ld", f"r{r}_{j}", t.rowend[r] - 16 * j, half_mask(t.depth[r] - 16 * j + 1)))
        for k in run:
            npiv += 1
            if npiv % 8 == 0:
                ir.append(("sync",))
            c, re = t.depth[k], t.rowend[k]
            nb = c // 16 + 1
            ir.append(("shflh", "d", f"r{k}_{c >> 4}", c & 15))
            ir.append(("rcp", "inv", "d"))
            for j in range(nb):
                ir.append(("mul", f"w{j}", f"r{k}_{j}", "inv"))
            for j in range(nb):
                lo = half_mask(c - 16 * j)            # columns with depth < c
                if lo:
                    ir.append(("st", re - 16 * j, f"w{j}", lo))
            jd = c >> 4
            ir.append(("st", re - 16 * jd, "inv", half_mask((c & 15) + 1) & ~half_mask(c & 15)))
            chain = [k] + t.anc[k]                    # chain[a] = a-th ancestor, depth c - a
            # the pivot-row broadcasts of a pivot are independent of each other (they read the finished row k): emit them in
            # batches of kBatch BEFORE the row updates that consume them, so that an in-order warp pays one shuffle latency per
            # batch instead of one per ancestor row (measured: the factorisation was 1046 exposed shuffle latencies)
            als = list(range(c - 1, -1, -1))
            for b0 in range(0, len(als), K_BATCH):
                batch = als[b0:b0 + K_BATCH]
                for q, al in enumerate(batch):
                    ir.append(("shflh", f"a{q}", f"r{k}_{al >> 4}", al & 15))
                for q, al in enumerate(batch):
                    row = chain[c - al]
                    for j in range(al // 16 + 1):
                        # UNMASKED: lanes beyond the row's last column (i > al - 16 j) get polluted, but a lane's entry is only
                        # ever combined with the same lane of other rows, shuffles read lanes <= depth and stores are masked,
                        # so the pollution never reaches a valid entry (run_factor_ir executes it on all 32 lanes, too)
                        ir.append(("fnma_reg", f"r{row}_{j}", f"a{q}", f"w{j}", 0xFFFFFFFF))
        for r in t.anc[bottom]:                       # ancestors of the chain: updated, not yet eliminated
            for j in range(t.depth[r] // 16 + 1):
                ir.append(("st", t.rowend[r] - 16 * j, f"r{r}_{j}", half_mask(t.depth[r] - 16 * j + 1)))
    return ir


def run_factor_ir(ir, t: Tree, M1: np.ndarray, M2: np.ndarray):
    """numpy interpreter of the stacked factorisation; M1 / M2: sparse rows [nM]; returns the two factors."""
    off2 = (t.nM + 3) & ~3
    pad = 64
    mem = np.full(pad + 2 * off2 + pad, np.nan, np.float32)
    mem[pad:pad + t.nM] = M1
    mem[pad + off2:pad + off2 + t.nM] = M2
    lanes = np.arange(32)
    base = pad + (lanes >> 4) * off2 - (lanes & 15)
    regs = {}
    for op in ir:
        if op[0] == "sync":
            continue
        if op[0] == "ld":
            _, r, imm, mask = op
            act = ((mask >> lanes) & 1).astype(bool)
            regs[r] = np.where(act, mem[base + imm], np.float32(0)).astype(np.float32)
        elif op[0] == "shflh":
            _, dst, src, l15 = op
            regs[dst] = regs[src][l15 | (lanes & 16)].astype(np.float32)
        elif op[0] == "rcp":
            regs[op[1]] = (np.float32(1) / regs[op[2]]).astype(np.float32)
        elif op[0] == "mul":
            regs[op[1]] = (regs[op[2]] * regs[op[3]]).astype(np.float32)
        elif op[0] == "st":
            _, imm, r, mask = op
            act = ((mask >> lanes) & 1).astype(bool)
            mem[(base + imm)[act]] = regs[r][act]
        elif op[0] == "fnma_reg":
            _, dst, a, w, mask = op
            act = ((mask >> lanes) & 1).astype(bool)
            with np.errstate(all="ignore"):
                regs[dst] = np.where(act, (regs[dst] - regs[a] * regs[w]).astype(np.float32), regs[dst])
        else:
            raise ValueError(op)
    return mem[pad:pad + t.nM].copy(), mem[pad + off2:pad + off2 + t.nM].copy()


def build_mulm_ir(t: Tree):
    """y = M x with the RAW sparse symmetric inertia (row i = dof i and its ancestors), x and y in dof-lane registers:
    pivot p's value travels by one shuffle; its descendants use M[i][p] (their own row), its ancestors M[p][j] (row p)."""
    nslot = (t.nv + 31) // 32
    ir = []
    for s in range(nslot):
        ir.append(("mulset_lds", f"y{s}", f"x{s}", f"G{s}", 0, lane_mask(range(t.nv), s)))
    piv = [p for p in range(t.nv) if t.desc[p] or t.anc[p]]
    for b0 in range(0, len(piv), K_BATCH):          # x is read-only: the broadcasts of a batch are independent, issue them first
        batch = piv[b0:b0 + K_BATCH]
        for q, p in enumerate(batch):
            ir.append(("shfl", f"p{q}", f"x{p // 32}", p % 32))
        for q, p in enumerate(batch):
            for s in range(nslot):
                m = lane_mask(t.desc[p], s)
                if m:
                    ir.append(("fma_lds", f"y{s}", f"D{s}", -t.depth[p], f"p{q}", m))
                m = lane_mask(t.anc[p], s)
                if m:
                    ir.append(("fma_lds", f"y{s}", f"U{s}", t.rowend[p], f"p{q}", m))
    return ir


def run_mulm_ir(ir, t: Tree, Ms: np.ndarray, x: np.ndarray) -> np.ndarray:
    nslot = (t.nv + 31) // 32
    pad = 64
    Lp = np.concatenate([np.full(pad, np.nan, np.float32), Ms.astype(np.float32), np.full(pad, np.nan, np.float32)])
    tabs = lane_tables(t)
    regs = {}
    for s in range(nslot):
        v = np.zeros(32, np.float32)
        n = min(32, t.nv - 32 * s)
        v[:n] = x[32 * s:32 * s + n]
        regs[f"x{s}"] = v
        regs[f"y{s}"] = np.zeros(32, np.float32)
    lanes = np.arange(32)
    for op in ir:
        if op[0] == "shfl":
            regs[op[1]] = np.full(32, regs[op[2]][op[3]], np.float32)
        elif op[0] == "mulset_lds":
            _, yr, xr, ptr, imm, mask = op
            act = ((mask >> lanes) & 1).astype(bool)
            regs[yr] = np.where(act, (regs[xr] * Lp[pad + tabs[ptr] + imm]).astype(np.float32), np.float32(0))
        elif op[0] == "fma_lds":
            _, yr, ptr, imm, p, mask = op
            act = ((mask >> lanes) & 1).astype(bool)
            regs[yr] = np.where(act, (regs[yr] + Lp[pad + tabs[ptr] + imm] * regs[p]).astype(np.float32), regs[yr])
    return np.concatenate([regs[f"y{s}"] for s in range(nslot)])[: t.nv]


def lane_tables(t: Tree):
    nslot = (t.nv + 31) // 32
    tabs = {}
    for s in range(nslot):
        dep = np.zeros(32, np.int64)
        rend = np.zeros(32, np.int64)
        madr = np.zeros(32, np.int64)
        for l in range(32):
            d = 32 * s + l
            if d < t.nv:
                dep[l], rend[l], madr[l] = t.depth[d], t.rowend[d], t.madr[d]
        tabs[f"U{s}"], tabs[f"D{s}"], tabs[f"G{s}"] = -dep, rend, madr
    return tabs


def run_ir(ir, t: Tree, L: np.ndarray, x: np.ndarray) -> np.ndarray:
    """numpy interpreter: L [nM] sparse factor (diagonal = 1/D), x [nv] -> solution [nv] (float32 arithmetic)."""
    nslot = (t.nv + 31) // 32
    pad = 64
    Lp = np.concatenate([np.full(pad, np.nan, np.float32), L.astype(np.float32), np.full(pad, np.nan, np.float32)])
    tabs = lane_tables(t)
    regs = {}
    for s in range(nslot):
        v = np.zeros(32, np.float32)
        n = min(32, t.nv - 32 * s)
        v[:n] = x[32 * s:32 * s + n]
        regs[f"x{s}"] = v
    lanes = np.arange(32)
    for op in ir:
        if op[0] == "shfl":
            regs[op[1]] = np.full(32, regs[op[2]][op[3]], np.float32)
        elif op[0] == "fnma_lds":
            _, xr, ptr, imm, p, mask = op
            act = ((mask >> lanes) & 1).astype(bool)
            val = Lp[pad + tabs[ptr] + imm]
            new = (regs[xr] - val * regs[p]).astype(np.float32)
            regs[xr] = np.where(act, new, regs[xr])
        elif op[0] == "mul_lds":
            _, xr, ptr, imm, mask = op
            act = ((mask >> lanes) & 1).astype(bool)
            val = Lp[pad + tabs[ptr] + imm]
            regs[xr] = np.where(act, (regs[xr] * val).astype(np.float32), regs[xr])
        else:
            raise ValueError(op)
    return np.concatenate([regs[f"x{s}"] for s in range(nslot)])[: t.nv]


# ------------------------------------------------------------------------------------------------ reference algebra
def sparse_from_dense(t: Tree, M: np.ndarray) -> np.ndarray:
    out = np.zeros(t.nM, M.dtype)
    for i in range(t.nv):
        out[t.madr[i]] = M[i, i]
        for a, j in enumerate(t.anc[i]):
            out[t.madr[i] + 1 + a] = M[i, j]
    return out


def factor_ref(t: Tree, Ms: np.ndarray) -> np.ndarray:
    """mj_factorM on the sparse rows; returns L with 1/D on the diagonal (what the kernel's factor_dual leaves)."""
    L = Ms.copy()
    for k in range(t.nv - 1, -1, -1):
        adr, c = t.madr[k], t.depth[k]
        inv = 1.0 / L[adr]
        for a in range(1, c + 1):
            ra = t.anc[k][a - 1]
            for b in range(a, c + 1):
                L[t.madr[ra] + (b - a)] -= L[adr + a] * L[adr + b] * inv
        L[adr + 1: adr + c + 1] *= inv
        L[adr] = inv
    return L


def random_tree_spd(t: Tree, rng) -> np.ndarray:
    """A random SPD matrix with the tree's sparsity (sum of chain outer products + diagonal)."""
    M = np.zeros((t.nv, t.nv))
    for i in range(t.nv):
        chain = [i] + t.anc[i]
        v = rng.normal(size=len(chain)) * rng.uniform(0.1, 1.0)
        for a, ia in enumerate(chain):
            for b, ib in enumerate(chain):
                M[ia, ib] += v[a] * v[b]
    M += np.diag(rng.uniform(0.05, 0.5, t.nv))
    return M


# ------------------------------------------------------------------------------------------------ CUDA printer

def _pred_fma(xr, ptr, imm, preg, mask, neg):
    """One predicated LDS + FFMA in place (inline PTX).  The C form `if (lb & m) x = fmaf(..)` compiled to a predicated FFMA
    into a temporary plus a predicated MOV and spilled predicates into a GPR bit-mask (ncu: 16 % IMAD.MOV/MOV, 12 % LOP3 of
    the solve's instructions); this pins it to LOP3(pred) + LDS + FFMA."""
    full = 0xFFFFFFFF
    sign = ""
    if neg:
        preg = "n" + preg   # the negated pivot value (one FADD per pivot instead of one operand negation per use)
    if mask == full:
        return (f'  asm("{{.reg .f32 t; ld.shared.f32 t, [%2+{4 * imm}]; {sign}fma.rn.f32 %0, t, %1, %0;}}" '
                f': "+f"({xr}) : "f"({preg}), "r"({ptr}));')
    return (f'  asm("{{.reg .pred q; .reg .b32 m; .reg .f32 t; and.b32 m, %3, 0x{mask:08x}; setp.ne.u32 q, m, 0; '
            f'ld.shared.f32 t, [%2+{4 * imm}]; {sign}@q fma.rn.f32 %0, t, %1, %0;}}" '
            f': "+f"({xr}) : "f"({preg}), "r"({ptr}), "r"(lb));')


def emit_cuda(t: Tree) -> str:
    ir = build_solve_ir(t)
    nslot = (t.nv + 31) // 32
    o = []
    w = o.append
    w("// GENERATED by tools/gen_tree_kernels.py -- do not edit; regenerate when the walker's dof tree changes.")
    w("// Straight-line L^T D L solves for one fixed kinematic tree (see the generator's docstring).")
    w("#ifndef TMJX_GEN_TREE_CUH_")
    w("#define TMJX_GEN_TREE_CUH_")
    w("namespace tmjx { namespace gen {")
    w(f"constexpr int kNv = {t.nv};")
    w(f"constexpr int kNM = {t.nM};")
    w(f"constexpr int kNMpad = {(t.nM + 3) & ~3};  // distance between the two factors in shared memory")
    w(f"constexpr int kNSlot = {nslot};")
    w(f"// dof_parentid of the tree this file was generated for (signature {t.signature()})")
    w("constexpr short kDofParent[kNv] = {" + ", ".join(str(p) for p in t.parent) + "};")
    w("struct V3 { float a, b, c; };")
    w("// the solve is called from four sites of a substep: as a call it costs ~24 caller-side register spills + reloads per call (the")
    w("// solver keeps ~90 live values per lane); -DTMJX_GEN_SOLVE_INLINE=__forceinline__ trades 2 k instructions of code for them")
    w("#ifndef TMJX_GEN_SOLVE_INLINE")
    w("#define TMJX_GEN_SOLVE_INLINE __noinline__")
    w("#endif")
    w("#ifdef __CUDACC__")
    w("// 1/d as MUFU.RCP + one Newton step (3 instructions, <= 1 ulp) instead of the IEEE division sequence with its")
    w("// slow-path call: 73 pivots per factorisation, every physics substep")
    w("static __device__ __forceinline__ float rcp_nr(float d) {")
    w('  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));')
    w("  return fmaf(fmaf(-d, r, 1.f), r, r);")
    w("}")
    w("// x <- (L^T D L)^-1 x.  L: this env's sparse factor in shared memory (diagonal holds 1/D);")
    w("// dep* / rend* : depth and row-end of the dofs this lane owns (lane, lane+32, lane+64).")
    w("static __device__ TMJX_GEN_SOLVE_INLINE V3 solve(const float* __restrict__ L, V3 xin, int lane, int dep0, int dep1, int dep2, int rend0, int rend1, int rend2) {")
    npiv = 1 + max(int(op[1][1:]) for op in ir if op[0] == "shfl")
    w("  float x0 = xin.a, x1 = xin.b, x2 = xin.c;")
    w("  float " + ", ".join(f"p{k}, np{k}" for k in range(npiv)) + ";")
    w("  const unsigned lb = 1u << lane;")
    w("  const unsigned sL = static_cast<unsigned>(__cvta_generic_to_shared(L));")
    for s in range(3):
        w(f"  const unsigned U{s} = sL - 4u * unsigned(dep{s});")
        w(f"  const unsigned D{s} = sL + 4u * unsigned(rend{s});")
        w(f"  const float* G{s} = L + (rend{s} - dep{s});")
    full = 0xFFFFFFFF
    for op in ir:
        if op[0] == "shfl":
            w(f"  {op[1]} = __shfl_sync(0xffffffffu, {op[2]}, {op[3]}); n{op[1]} = -{op[1]};")
        elif op[0] == "fnma_lds":
            _, xr, ptr, imm, p, mask = op
            w(_pred_fma(xr, ptr, imm, p, mask, True))
        elif op[0] == "mul_lds":
            _, xr, ptr, imm, mask = op
            guard = "" if mask == full else f"if (lb & 0x{mask:08x}u) "
            w(f"  {guard}{xr} *= {ptr}[{imm}];")
    w("  V3 r; r.a = x0; r.b = x1; r.c = x2;")
    w("  return r;")
    w("}")
    w("// y = M x with the raw sparse inertia at M (before it is factored); x, y: dof-lane registers.")
    w("static __device__ __noinline__ V3 mul_m(const float* __restrict__ L, V3 xin, int lane, int dep0, int dep1, int dep2, int rend0, int rend1, int rend2) {")
    w("  const float x0 = xin.a, x1 = xin.b, x2 = xin.c;")
    w("  float y0 = 0.f, y1 = 0.f, y2 = 0.f, " + ", ".join(f"p{q}" for q in range(K_BATCH)) + ";")
    w("  const unsigned lb = 1u << lane;")
    w("  const unsigned sL = static_cast<unsigned>(__cvta_generic_to_shared(L));")
    for s in range(3):
        w(f"  const unsigned U{s} = sL - 4u * unsigned(dep{s});")
        w(f"  const unsigned D{s} = sL + 4u * unsigned(rend{s});")
        w(f"  const float* G{s} = L + (rend{s} - dep{s});")
    for op in build_mulm_ir(t):
        if op[0] == "shfl":
            w(f"  {op[1]} = __shfl_sync(0xffffffffu, {op[2]}, {op[3]});")
        elif op[0] == "mulset_lds":
            _, yr, xr, ptr, imm, mask = op
            guard = "" if mask == full else f"if (lb & 0x{mask:08x}u) "
            w(f"  {guard}{yr} = {xr} * {ptr}[{imm}];")
        elif op[0] == "fma_lds":
            _, yr, ptr, imm, p, mask = op
            w(_pred_fma(yr, ptr, imm, p, mask, False))
    w("  V3 r; r.a = y0; r.b = y1; r.c = y2;")
    w("  return r;")
    w("}")
    w("// Both L^T D L factorisations (M at L, M + dt diag(damping) at L + kNMpad), stacked half-warp per matrix; see")
    w("// factor_dual in tmjx_step.cu for the loop form of the same schedule (used for other trees).")
    w("// off2 = kNMpad: two matrices; off2 = 0: both half-warps factor the SAME matrix at L (identical values, identical stores)")
    w("// -- the single-matrix factorisation Newton needs once per solver iteration, at the same instruction count.")
    w("static __device__ __noinline__ void factor_dual(float* __restrict__ L, int lane, bool sync, int off2) {")
    w("  const unsigned lb = 1u << lane;")
    w("  const int hbit = lane & 16;")
    w("  float* ps = L + (hbit ? off2 : 0) - (lane & 15);")
    fir = build_factor_ir(t)
    w("  float w0 = 0.f, w1 = 0.f, w2 = 0.f, d, inv, " + ", ".join(f"a{q}" for q in range(K_BATCH)) + ";")
    declared = set()
    for op in fir:
        if op[0] == "sync":
            w("  if (sync) __syncthreads();")
        elif op[0] == "ld":
            _, r, imm, mask = op
            decl = "" if r in declared else "float "
            declared.add(r)
            if mask == full:
                w(f"  {decl}{r} = ps[{imm}];")
            else:
                w(f"  {decl}{r} = (lb & 0x{mask:08x}u) ? ps[{imm}] : 0.f;")
        elif op[0] == "shflh":
            w(f"  {op[1]} = __shfl_sync(0xffffffffu, {op[2]}, {op[3]} | hbit);")
        elif op[0] == "rcp":
            w(f"  {op[1]} = rcp_nr({op[2]});")
        elif op[0] == "mul":
            w(f"  {op[1]} = {op[2]} * {op[3]};")
        elif op[0] == "st":
            _, imm, r, mask = op
            guard = "" if mask == full else f"if (lb & 0x{mask:08x}u) "
            w(f"  {guard}ps[{imm}] = {r};")
        elif op[0] == "fnma_reg":
            _, dst, a_, w_, mask = op
            guard = "" if mask == full else f"if (lb & 0x{mask:08x}u) "
            w(f"  {guard}{dst} = fmaf(-{a_}, {w_}, {dst});")
    w("}")
    w("#endif  // __CUDACC__")
    w("}}  // namespace tmjx::gen")
    w("#endif  // TMJX_GEN_TREE_CUH_")
    return "\n".join(o) + "\n"


def rodent_tree() -> Tree:
    sys.path.insert(0, ROOT)
    from track_mjx_b200.walker import Rodent

    w = Rodent(torque_actuators=True, rescale_factor=0.9)
    return Tree(np.asarray(w.sections["dof_parentid"]).astype(int))


def main():
    t = rodent_tree()
    src = emit_cuda(t)
    with open(OUT, "w") as f:
        f.write(src)
    ir = build_solve_ir(t)
    fir = build_factor_ir(t)
    print(f"wrote {OUT}: nv {t.nv} nM {t.nM} maxdepth {t.maxdepth}, solve IR {len(ir)} ops "
          f"({sum(1 for x in ir if x[0] == 'shfl')} shuffles), factor IR {len(fir)} ops "
          f"({sum(1 for x in fir if x[0] == 'fnma_reg')} row updates, {sum(1 for x in fir if x[0] in ('ld', 'st'))} LDS/STS)")


if __name__ == "__main__":
    main()
