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
    device = torch.device("cpu")

    def to_device(self, arr, dtype=torch.float64, non_blocking=True):
        if isinstance(arr, torch.Tensor):
            return arr.to(dtype)
        return torch.from_numpy(np.ascontiguousarray(arr)).to(dtype)

    def fit(self, A, b, w, testing=None, alpha=0.0, refine=2, group=None, diagnostics=True):
        from fitsnap_b200.engine import fit_rows
        return fit_rows(self, A, b, w, testing, alpha=alpha, refine=refine, group=group, diagnostics=diagnostics)

    def scatter(self, batch, A=None, b=None, w=None, lda=None):
        """Row assembly through the oracle (lammps_snap.py:391-556 restated in oracle/linear_fit.py)."""
        nat = batch.natoms.numpy()
        roff = batch.raw_row_off.numpy()
        aoff = np.concatenate([[0], np.cumsum(nat.astype(np.int64))])
        raw = batch.raw.numpy()
        f = batch.flags
        cfgs = []
        for c in range(batch.ncfg):
            cfgs.append(dict(block=raw[roff[c]:roff[c + 1]], natoms=int(nat[c]), volume=float(batch.volume[c]),
                             energy=float(batch.energy[c]), forces=batch.forces.numpy()[3 * aoff[c]:3 * aoff[c + 1]],
                             stress=batch.stress.numpy()[c].reshape(3, 3), eweight=float(batch.eweight[c]),
                             fweight=float(batch.fweight[c]), vweight=float(batch.vweight[c]),
                             type_fraction=batch.type_fraction.numpy()[c]))
        if f & 16:
            for c in cfgs:
                c["block"] = np.nan_to_num(c["block"])
        a_, b_, w_ = lf.assemble(cfgs, batch.numtypes, batch.ncoeff, bool(f & 8), batch.blank2j.numpy(),
                                 bool(f & 1), bool(f & 2), bool(f & 4))
        bad = torch.tensor([int(not np.isfinite(raw).all())], dtype=torch.int32)
        if A is None:
            return torch.from_numpy(a_), torch.from_numpy(b_), torch.from_numpy(w_), bad
        A[batch.row_begin:batch.row_end] = torch.from_numpy(a_)
        b[batch.row_begin:batch.row_end] = torch.from_numpy(b_)
        w[batch.row_begin:batch.row_end] = torch.from_numpy(w_)
        return A, b, w, bad

    def group_stats(self, A, b, w, group_id, x, n_groups):
        a, t, wv, g = A.numpy(), b.numpy(), w.numpy(), group_id.numpy()
        res = t - a @ x.numpy()
        out = np.zeros((n_groups, 10))
        cols = [np.ones_like(t), np.abs(res), res ** 2, t, t * t, (wv != 0).astype(float), np.abs(wv * res),
                (wv * res) ** 2, wv * t, (wv * t) ** 2]
        for q, c in enumerate(cols):
            np.add.at(out[:, q], g, c)
        return torch.from_numpy(out)

    def predict(self, A, x):
        return torch.from_numpy(A.numpy() @ x.numpy())

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
