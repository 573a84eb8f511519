"""Independent numpy/scipy twin of Optimizer::OptimizeSim3 (Optimizer.cc:1348-1543), used only to pin oracle_optimize_sim3.

Same schedule (5 LM iterations, drop pairs with chi2 > th2 on either edge, 10 or 5 more, the < 10 correspondences rule) and the same
Levenberg control flow as g2o, but different numerics on purpose: the Sim3 is a 4x4 matrix [sR t; 0 1], the update is scipy.linalg.expm of
the sim(3) generator (not g2o's closed form), the inverse is a matrix inverse, and the Jacobians are central differences with a 1e-6 step
(g2o: 1e-9) -- so agreement checks the oracle's Sim3 exponential / inverse / product, error functions, Huber weighting and LM."""
import numpy as np
import scipy.linalg


def _gen(u):
    w, v, s = u[:3], u[3:6], u[6]
    A = np.zeros((4, 4))
    A[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) + s * np.eye(3)
    A[:3, 3] = v
    return A


def to_matrix(sim3):
    from scipy.spatial.transform import Rotation
    q, t, s = np.asarray(sim3[:4], float), np.asarray(sim3[4:7], float), float(sim3[7])
    T = np.eye(4); T[:3, :3] = s * Rotation.from_quat(q / np.linalg.norm(q)).as_matrix(); T[:3, 3] = t
    return T


class Twin:
    def __init__(self, sim3, valid, P1c, P2c, obs1, obs2, w1, w2, K1, K2, th2, fix_scale):
        self.T = to_matrix(sim3)
        f = lambda a: np.asarray(a, np.float32).astype(np.float64)
        self.P1, self.P2, self.o1, self.o2, self.w1, self.w2, self.K1, self.K2 = f(P1c), f(P2c), f(obs1), f(obs2), f(w1), f(w2), f(K1), f(K2)
        self.th2 = float(np.float32(th2)); self.delta = float(np.sqrt(np.float32(th2))); self.fix = bool(fix_scale)
        self.active = np.asarray(valid) > 0
        self.iterations = self.trials = 0

    def errors(self, T):
        a = self.active
        p = self.P2[a] @ T[:3, :3].T + T[:3, 3]
        e12 = self.o1[a] - np.stack([p[:, 0] / p[:, 2] * self.K1[0] + self.K1[2], p[:, 1] / p[:, 2] * self.K1[1] + self.K1[3]], 1)
        Ti = np.linalg.inv(T)
        p = self.P1[a] @ Ti[:3, :3].T + Ti[:3, 3]
        e21 = self.o2[a] - np.stack([p[:, 0] / p[:, 2] * self.K2[0] + self.K2[2], p[:, 1] / p[:, 2] * self.K2[1] + self.K2[3]], 1)
        return np.concatenate([e12, e21], 1)                      # [n, 4]

    def chi2(self, e):
        a = self.active
        return np.stack([self.w1[a] * (e[:, :2] ** 2).sum(1), self.w2[a] * (e[:, 2:] ** 2).sum(1)], 1)

    def robust(self, c):
        d2 = self.delta ** 2
        rho = np.where(c <= d2, c, 2 * np.sqrt(np.maximum(c, 1e-300)) * self.delta - d2)
        rho1 = np.where(c <= d2, 1.0, self.delta / np.sqrt(np.maximum(c, 1e-300)))
        return rho, rho1

    def oplus(self, T, u):
        u = np.array(u, float)
        if self.fix:
            u[6] = 0
        return scipy.linalg.expm(_gen(u)) @ T

    def optimize(self, iterations):
        if not self.active.any():
            return
        lam = ni = nbad = None
        a = self.active
        W = np.stack([self.w1[a], self.w1[a], self.w2[a], self.w2[a]], 1)
        for it in range(iterations):
            e = self.errors(self.T); c = self.chi2(e); rho, rho1 = self.robust(c)
            cur = ini = rho.sum()
            h = 1e-6
            J = np.zeros(e.shape + (7,))
            for d in range(7):
                u = np.zeros(7); u[d] = h
                J[:, :, d] = (self.errors(self.oplus(self.T, u)) - self.errors(self.oplus(self.T, -u))) / (2 * h)
            R1 = np.repeat(rho1, 2, axis=1)                         # [n, 4]: robust weight per residual
            H = np.einsum("nra,nr,nrb->ab", J, W * R1, J)
            b = -np.einsum("nra,nr->a", J, W * R1 * e)
            if it == 0:
                lam = 1e-5 * np.abs(np.diag(H)).max(); ni = 2.0; nbad = 0
            rho_gain, q = 0.0, 0
            while True:
                backup = self.T.copy()
                try:
                    x = np.linalg.solve(H + lam * np.eye(7), b); ok = True
                    np.linalg.cholesky(H + lam * np.eye(7))
                except np.linalg.LinAlgError:
                    x = np.zeros(7); ok = False
                self.T = self.oplus(self.T, x)
                tmp = self.robust(self.chi2(self.errors(self.T)))[0].sum() if ok else np.finfo(float).max
                rho_gain = (cur - tmp) / (x @ (lam * x + b) + 1e-3)
                if rho_gain > 0 and np.isfinite(tmp):
                    lam *= max(1 / 3, min(1 - (2 * rho_gain - 1) ** 3, 2 / 3)); ni = 2.0; cur = tmp
                else:
                    lam *= ni; ni *= 2; self.T = backup
                q += 1; self.trials += 1
                if not (rho_gain < 0 and q < 10):
                    break
            self.iterations += 1
            if q == 10 or rho_gain == 0:
                break
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            if nbad >= 3:
                break

    def run(self):
        n_corr = int(self.active.sum())
        self.optimize(5)
        c = self.chi2(self.errors(self.T))
        idx = np.where(self.active)[0]
        bad = (c > self.th2).any(1)
        inlier = self.active.copy(); inlier[idx[bad]] = False
        self.active = inlier.copy()
        if n_corr - int(bad.sum()) < 10:
            return None, inlier, 0
        self.optimize(10 if bad.any() else 5)
        c = self.chi2(self.errors(self.T))
        idx = np.where(self.active)[0]
        bad2 = (c > self.th2).any(1)
        inlier[idx[bad2]] = False
        return self.T, inlier, int((~bad2).sum())
