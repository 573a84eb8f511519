"""Independent numpy/scipy twin of the reference LM bundle adjustment, used only to pin oracle/slam_oracle.c.

Same control flow as g2o's OptimizationAlgorithmLevenberg (lambda init, gain ratio with the +1e-3, accept / reject,
nu doubling, ORB-SLAM's extra stop rule) but different numerics on purpose: poses are 4x4 matrices, the update is
scipy.linalg.expm of the twist, the FULL normal equations (poses and points together, no Schur complement) are
assembled as a scipy sparse matrix and solved with a sparse LU.  Agreement with the C oracle therefore checks its
Jacobians, robust weighting, Schur complement, profile LDLT, back-substitution and manifold update."""
import numpy as np
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg as spla

DELTA = float(np.float32(np.sqrt(5.991)))
DSQR = DELTA * DELTA


def _skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def _exp(u):
    A = np.zeros((4, 4)); A[:3, :3] = _skew(u[:3]); A[:3, 3] = u[3:]
    return scipy.linalg.expm(A)


class Twin:
    def __init__(self, poses, fixed, intr, points, kf, pt, uv, w):
        self.T = np.asarray(poses, np.float32).astype(np.float64).reshape(-1, 4, 4).copy()
        self.fixed = np.asarray(fixed).astype(bool)
        self.intr = np.asarray(intr, np.float64).reshape(-1, 4) if np.ndim(intr) == 2 else np.tile(np.asarray(intr, np.float64), (len(self.T), 1))
        self.X = np.asarray(points, np.float32).astype(np.float64).copy()
        self.kf, self.pt = np.asarray(kf), np.asarray(pt)
        self.uv = np.asarray(uv, np.float32).astype(np.float64); self.w = np.asarray(w, np.float32).astype(np.float64)
        E = len(self.kf)
        self.level = np.zeros(E, int); self.robust = np.ones(E, bool); self.err = np.zeros((E, 2))
        self.iterations = self.trials = 0

    def _cam(self, idx):
        R = self.T[self.kf[idx], :3, :3]; t = self.T[self.kf[idx], :3, 3]
        return np.einsum("eij,ej->ei", R, self.X[self.pt[idx]]) + t

    def compute_errors(self, idx):
        Xc = self._cam(idx); K = self.intr[self.kf[idx]]
        self.err[idx, 0] = self.uv[idx, 0] - (Xc[:, 0] / Xc[:, 2] * K[:, 0] + K[:, 2])
        self.err[idx, 1] = self.uv[idx, 1] - (Xc[:, 1] / Xc[:, 2] * K[:, 1] + K[:, 3])

    def chi2(self, idx):
        return self.w[idx] * (self.err[idx] ** 2).sum(1)

    def robust_chi2(self, idx):
        c = self.chi2(idx)
        out = np.where(c <= DSQR, c, 2 * np.sqrt(np.maximum(c, 1e-300)) * DELTA - DSQR)
        return float(np.where(self.robust[idx], out, c).sum())

    def setup(self):
        act = np.nonzero(self.level == 0)[0]
        self.act = act
        pa = np.zeros(len(self.T), bool); pa[self.kf[act]] = True
        la = np.zeros(len(self.X), bool); la[self.pt[act]] = True
        self.pose_ids = np.nonzero(pa & ~self.fixed)[0]; self.pt_ids = np.nonzero(la)[0]
        self.pidx = -np.ones(len(self.T), int); self.pidx[self.pose_ids] = np.arange(len(self.pose_ids))
        self.lidx = -np.ones(len(self.X), int); self.lidx[self.pt_ids] = np.arange(len(self.pt_ids))

    def build(self):
        idx = self.act
        Xc = self._cam(idx); x, y, z = Xc.T
        K = self.intr[self.kf[idx]]; fx, fy = K[:, 0], K[:, 1]
        n = len(idx)
        Jp = np.zeros((n, 2, 6))
        Jp[:, 0, 0] = x * y / z ** 2 * fx; Jp[:, 0, 1] = -(1 + x * x / z ** 2) * fx; Jp[:, 0, 2] = y / z * fx
        Jp[:, 0, 3] = -1 / z * fx; Jp[:, 0, 5] = x / z ** 2 * fx
        Jp[:, 1, 0] = (1 + y * y / z ** 2) * fy; Jp[:, 1, 1] = -x * y / z ** 2 * fy; Jp[:, 1, 2] = -x / z * fy
        Jp[:, 1, 4] = -1 / z * fy; Jp[:, 1, 5] = y / z ** 2 * fy
        M = np.zeros((n, 2, 3)); M[:, 0, 0] = fx; M[:, 0, 2] = -x / z * fx; M[:, 1, 1] = fy; M[:, 1, 2] = -y / z * fy
        Jl = -(1 / z)[:, None, None] * np.einsum("eij,ejk->eik", M, self.T[self.kf[idx], :3, :3])
        c = self.chi2(idx)
        rho1 = np.where(self.robust[idx] & (c > DSQR), DELTA / np.sqrt(np.maximum(c, 1e-300)), 1.0)
        wo = rho1 * self.w[idx]
        nP, nL = len(self.pose_ids), len(self.pt_ids)
        N = 6 * nP + 3 * nL
        rows, cols, vals = [], [], []
        ip = self.pidx[self.kf[idx]]; il = self.lidx[self.pt[idx]]
        J = np.zeros((n, 2, 9)); J[:, :, :6] = Jp; J[:, :, 6:] = Jl
        base = np.zeros((n, 9), int)
        base[:, :6] = 6 * ip[:, None] + np.arange(6); base[:, 6:] = 6 * nP + 3 * il[:, None] + np.arange(3)
        valid = np.ones((n, 9), bool); valid[:, :6] = (ip >= 0)[:, None]
        Hb = np.einsum("eia,e,eib->eab", J, wo, J)
        bb = -np.einsum("eia,e,ei->ea", J, wo, self.err[idx])
        m = valid[:, :, None] & valid[:, None, :]
        r = np.broadcast_to(base[:, :, None], Hb.shape)[m]; cc = np.broadcast_to(base[:, None, :], Hb.shape)[m]
        H = sp.coo_matrix((Hb[m], (r, cc)), shape=(N, N)).tocsc()
        b = np.zeros(N); np.add.at(b, base[valid], bb[valid])
        return H, b

    def lm_iteration(self, it):
        idx = self.act
        self.compute_errors(idx)
        cur = self.robust_chi2(idx); ini = cur
        H, b = self.build()
        if it == 0:
            self.lam = 1e-5 * np.abs(H.diagonal()).max(); self.ni = 2.0; self.nbad = 0
        rho, q = 0.0, 0
        nP = len(self.pose_ids)
        while True:
            Tb, Xb = self.T.copy(), self.X.copy()
            try:
                x = spla.spsolve(H + self.lam * sp.identity(H.shape[0], format="csc"), b)
                ok = np.all(np.isfinite(x))
            except Exception:
                ok = False; x = np.zeros_like(b)
            for i, k in enumerate(self.pose_ids):
                self.T[k] = _exp(x[6 * i:6 * i + 6]) @ self.T[k]
            self.X[self.pt_ids] += x[6 * nP:].reshape(-1, 3)
            self.compute_errors(idx)
            tmp = self.robust_chi2(idx) if ok else np.finfo(float).max
            rho = (cur - tmp) / (float(x @ (self.lam * x + b)) + 1e-3)
            if rho > 0 and np.isfinite(tmp):
                self.lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3)); self.ni = 2.0; cur = tmp
            else:
                self.lam *= self.ni; self.ni *= 2; self.T, self.X = Tb, Xb
            q += 1; self.trials += 1
            if not (rho < 0 and q < 10):
                break
        self.iterations += 1
        if q == 10 or rho == 0:
            return False
        self.nbad = self.nbad + 1 if (ini - cur) * 1e3 < ini else 0
        return self.nbad < 3

    def optimize(self, its):
        self.setup()
        if len(self.pose_ids) + len(self.pt_ids) == 0:
            return
        for i in range(its):
            if not self.lm_iteration(i):
                break

    def depth_ok(self):
        return self._cam(np.arange(len(self.kf)))[:, 2] > 0


def local_ba(poses, fixed, intr, points, kf, pt, uv, w, its0=5, its1=10):
    t = Twin(poses, fixed, intr, points, kf, pt, uv, w)
    t.optimize(its0)
    allidx = np.arange(len(t.kf))
    bad = (t.chi2(allidx) > 5.991) | ~t.depth_ok()
    t.level[bad] = 1; t.robust[:] = False
    t.optimize(its1)
    return t
