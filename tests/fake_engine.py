"""TEST DOUBLE (tests only): an object with the Engine's gram/factor/solve/residual interface whose
arithmetic is the CPU oracle, so that the host-side orchestration (`fit_rows`: shard -> Gram ->
all-reduce -> replicated solve -> refinement with all-reduced residual) can be exercised with
torch.distributed/gloo on a machine without a GPU.  Never imported by the product."""
import numpy as np
import scipy.linalg as sl
import torch

from oracle import linear_fit as lf


class _Factor:
    def __init__(self, chol, d, alpha):
        self.chol, self.d, self.alpha = chol, d, alpha
        self.info = torch.zeros(8, dtype=torch.int32)


class OracleEngine:
    launch_count = 0

    def gram(self, A, b, w, testing=None):
        a, bb, ww = A.numpy(), b.numpy(), w.numpy()
        t = None if testing is None else testing.numpy().astype(bool)
        G, c, btb, _ = lf.gram(a, bb, ww, t)
        k = a.shape[1]
        full = np.zeros((k + 1, k + 1))
        full[:k, :k], full[:k, k], full[k, :k], full[k, k] = G, c, c, btb
        return torch.from_numpy(full)

    def factor(self, gaug, alpha=0.0):
        g = gaug.numpy()
        k = g.shape[0] - 1
        G = g[:k, :k] + alpha * np.eye(k)
        d = 1.0 / np.sqrt(np.diag(G))
        return _Factor(np.linalg.cholesky(G * d[:, None] * d[None, :]), d, alpha)

    def solve(self, f, rhs, rhs_stride=1, x_in=None, out=None):
        r = rhs.numpy()[:len(f.d)].copy()
        x0 = np.zeros_like(r) if x_in is None else x_in.numpy()
        r = f.d * (r - f.alpha * x0)
        y = sl.solve_triangular(f.chol, r, lower=True)
        z = sl.solve_triangular(f.chol.T, y, lower=False)
        return torch.from_numpy(x0 + f.d * z)

    def residual(self, A, b, w, testing, x):
        t = None if testing is None else testing.numpy().astype(bool)
        aw, bw = lf.weighted_system(A.numpy(), b.numpy(), w.numpy(), t)
        return torch.from_numpy(aw.T @ (bw - aw @ x.numpy()))
