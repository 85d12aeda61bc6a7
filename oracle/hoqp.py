"""Hierarchical QP (HoQP) restated for the CPU -- TEST INFRASTRUCTURE, groundwork for SURVEY 8f row N4 (not on the product path yet).

Follows legged_wbc/src/HoQp.cpp:20-198 (Bellicoso et al., "Perception-less terrain adaptation through whole body control and
hierarchical optimization"): each priority level solves
    min 1/2 ||A_k (x_prev + Z_prev z) - b_k||^2 + 1/2 ||v||^2   s.t.  D_k (x_prev + Z_prev z) - f_k <= v,  v >= 0,
                                                                     D_prev (x_prev + Z_prev z) - f_prev <= v_prev* (stacked, kept)
in the null space Z_prev of all higher-priority equality tasks; Z_k = Z_prev kernel(A_k Z_prev). The QP is solved with the oracle's
interior-point solver (the reference calls qpOASES with setToMPC, un-vendored). Task stacking as in legged_wbc/include/legged_wbc/Task.h:46-58.
"""
import numpy as np


class Task:
    def __init__(self, a=None, b=None, d=None, f=None, n=None):
        n = n if n is not None else (a.shape[1] if a is not None and a.size else d.shape[1])
        self.a = np.zeros((0, n)) if a is None else np.asarray(a, dtype=float).reshape(-1, n)
        self.b = np.zeros(0) if b is None else np.asarray(b, dtype=float).reshape(-1)
        self.d = np.zeros((0, n)) if d is None else np.asarray(d, dtype=float).reshape(-1, n)
        self.f = np.zeros(0) if f is None else np.asarray(f, dtype=float).reshape(-1)

    def __add__(self, other):
        return Task(np.vstack([self.a, other.a]), np.concatenate([self.b, other.b]), np.vstack([self.d, other.d]), np.concatenate([self.f, other.f]))

    def __mul__(self, w):
        return Task(self.a * w, self.b * w, self.d * w, self.f * w, n=self.a.shape[1])


def kernel_basis(M, n):
    """A basis of the null space of M (FullPivLU::kernel spans the same space; the hierarchy only depends on the space)."""
    if M.shape[0] == 0:
        return np.eye(n)
    u, s, vt = np.linalg.svd(M, full_matrices=True)
    tol = max(M.shape) * np.finfo(float).eps * (s[0] if s.size else 1.0)
    rank = int((s > tol).sum())
    return vt[rank:].T


class HoQp:
    def __init__(self, task, higher=None, qp_solve=None):
        from . import hbo
        self.task, self.higher = task, higher
        solve = qp_solve or (lambda H, g, A, lb, ub: hbo.qp_solve(H, g, A, lb, ub, 1e-10))
        nv = task.d.shape[0]
        if higher is not None:
            Zp, Tp, sp, xp = higher.stacked_z, higher.stacked_tasks, higher.stacked_slack, higher.x
        else:
            n = max(task.a.shape[1], task.d.shape[1])
            Zp, Tp, sp, xp = np.eye(n), Task(n=n), np.zeros(0), np.zeros(n)
        nx = Zp.shape[1]
        self.stacked_tasks = task + Tp
        if nx == 0:      # nothing left to decide: the higher-priority solution stands, the slacks absorb the violation
            self.z = np.zeros(0); self.slack = np.maximum(task.d @ xp - task.f, 0.0) if nv else np.zeros(0)
        else:
            AZ = task.a @ Zp
            H = np.zeros((nx + nv, nx + nv)); H[:nx, :nx] = AZ.T @ AZ + 1e-12 * np.eye(nx); H[nx:, nx:] = np.eye(nv)
            c = np.concatenate([AZ.T @ (task.a @ xp - task.b), np.zeros(nv)])
            npv = Tp.d.shape[0]
            D = np.zeros((2 * nv + npv, nx + nv))
            D[:nv, nx:] = -np.eye(nv)
            D[nv:nv + npv, :nx] = Tp.d @ Zp
            D[nv + npv:, :nx] = task.d @ Zp; D[nv + npv:, nx:] = -np.eye(nv)
            f = np.concatenate([np.zeros(nv), Tp.f - Tp.d @ xp + sp, task.f - task.d @ xp])
            sol, st, _ = solve(H, c, D, np.full(f.size, -1e20), f)
            if st != 0:
                raise RuntimeError("HoQP level did not solve, status %d" % st)
            self.z, self.slack = sol[:nx], sol[nx:]
        self.x = xp + Zp @ self.z
        self.stacked_z = Zp @ kernel_basis(task.a @ Zp, nx) if task.a.shape[0] > 0 and nx > 0 else Zp
        # Stacked in the same order as the stacked tasks ([this level; higher levels]). The reference concatenates the other way round
        # (HoQp.cpp:186-196) while stacking the tasks this way (Task.h:46-52); its own hierarchy only has inequalities on the first
        # level, where the two orders coincide.
        self.stacked_slack = np.concatenate([self.slack, sp]) if higher is not None else self.slack

    def solution(self):
        return self.x


def hierarchical_wbc(x_des, u_des, rbd, mode):
    """legged::HierarchicalWbc::update (legged_wbc/src/HierarchicalWbc.cpp:18-31) on the oracle's WBC pieces:
       task0 = floating-base EoM + torque limits + friction cone (incl. zero swing forces) + no contact motion,
       task1 = base acceleration (xy, height, angular),  task2 = 0.1 * contact force + swing leg.
    Decision vector x = [qdd(16), F(12), tau(10)] as in the weighted formulation. Returns (x, [HoQp levels])."""
    from . import hbo
    H, g, A, lb, ub = hbo.wbc_assemble(x_des, u_des, rbd, mode, False)
    Aw, bw, J, dJv = hbo.wbc_terms(x_des, u_des, rbd, mode)
    keep = np.abs(A).sum(axis=1) > 0
    A, lb, ub = A[keep], lb[keep], ub[keep]
    eq = lb == ub
    a0, b0 = A[eq], ub[eq]
    up = (~eq) & (ub < 1e19); dn = (~eq) & (lb > -1e19)
    d0 = np.vstack([A[up], -A[dn]]); f0 = np.concatenate([ub[up], -lb[dn]])
    stance = [mode in (2, 3), mode in (1, 3), mode in (2, 3), mode in (1, 3)]
    rows = [3 * c + k for c in range(4) if stance[c] for k in range(3)]
    ncm_a = np.zeros((len(rows), 38)); ncm_a[:, :16] = J[rows]; ncm_b = -dJv[rows]          # formulateNoContactMotionTask (WbcBase.cpp:169-188)
    task0 = Task(np.vstack([a0, ncm_a]), np.concatenate([b0, ncm_b]), d0, f0)
    nsw = 3 * (4 - sum(stance))
    w_swing, w_base = 100.0, 1.0                                                              # task.info:328-333 (weights divided out again)
    swing = Task(Aw[:nsw] / w_swing, bw[:nsw] / w_swing, n=38)
    base = Task(Aw[nsw:] / w_base, bw[nsw:] / w_base, n=38)
    cf_a = np.zeros((12, 38)); cf_a[:, 16:28] = np.eye(12)
    force = Task(cf_a, np.asarray(u_des[:12], dtype=float), n=38)                             # formulateContactForceTask (WbcBase.cpp:325-338)
    task1 = base
    task2 = force * 0.1 + swing * 1.0
    l0 = HoQp(task0); l1 = HoQp(task1, l0); l2 = HoQp(task2, l1)
    return l2.solution(), [l0, l1, l2], (task0, task1, task2)
