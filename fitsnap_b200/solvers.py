"""Host-side mirror of FitSNAP's linear `Solver` plugins (SVD, RIDGE, LASSO), backed by the
B200 engine.  Same class names, constructor and `perform_fit` signatures, same infile keys,
same result convention as the reference:

    fitsnap3lib/solvers/svd.py:13-54     class SVD(Solver).perform_fit(a, b, w, fs_dict, trainall)
    fitsnap3lib/solvers/ridge.py:6-60    class RIDGE(Solver).perform_fit(...)   [RIDGE] alpha, local_solver
    fitsnap3lib/solvers/lasso.py:9-30    class LASSO(Solver).perform_fit()      [LASSO] alpha, max_iter
    fitsnap3lib/solvers/anl.py:7-67      class ANL(Solver).perform_fit(a, b, w, trainall)  [SOLVER] cov_nugget, nsam
    fitsnap3lib/solvers/solver.py:19-46  Solver.__init__(name, pt, config, linear=True), fit_gather()

The classes here do NOT import fitsnap3lib (it is absent on a bare GPU box); they duck-type
`pt` (attributes `_rank`, `shared_arrays[...].array`, `fitsnap_dict`) and `config`
(`config.sections[...]`).  `fitsnap_b200.plugin.register()` grafts them under the reference's
`Solver` base so that `[SOLVER] solver = SVD|RIDGE|LASSO` resolves to them through the
reference's own factory (solvers/solver_factory.py:18-34).

Inputs are host numpy fp64 (as in the reference) or CUDA tensors (device-resident shards);
`self.fit` is a host `np.ndarray float64 (K,)` on rank 0 exactly as the reference leaves it.
There is no CPU fallback: without the CUDA extension these classes raise.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _engine
from .hostmirror import LazyHostMirror, shared_shape


def _section(config, name):
    secs = getattr(config, "sections", None)
    if secs is None or name not in secs:
        return None
    return secs[name]


def _get(sec, key, default):
    if sec is None:
        return default
    if isinstance(sec, dict):
        return sec.get(key, default)
    return getattr(sec, key, default)


class LinearSolverBase:
    """Common prologue of the three linear solvers (svd.py:31-46 == ridge.py:24-39 == lasso.py:19-21)."""

    #: refinement rounds: "auto" iterates until the correction stalls (one host sync per round)
    refine = "auto"
    max_refine = 10
    #: rank-deficient least squares (alpha = 0): reproduce lstsq's minimum-norm answer through G^+
    min_norm_fallback = True

    def __init__(self, name, pt, config, linear=True):
        self.name = name
        self.pt = pt
        self.config = config
        self.fit = None
        self.linear = linear
        self.errors = []
        self.engine = None
        self.last_result = None
        self.process_group = None      # set by fitsnap_b200.distributed for row-sharded fits

    # -- reference API -----------------------------------------------------------------
    def fit_gather(self):
        """solver.py:46 (no-op in the reference too)."""

    def _alpha(self):
        return 0.0

    def _training_mask(self, n, fs_dict, trainall):
        """svd.py:35-40: fs_dict['Testing'], else trainall, else pt.fitsnap_dict['Testing'].
        Unlike the explicit-array branch of svd.py:46 (which forgets to mask `w` and therefore
        raises a broadcast error as soon as a test row exists) the mask is applied to a, b AND w,
        which is what the shared-array branch svd.py:42-44 does."""
        if fs_dict is not None:
            testing = np.asarray(fs_dict["Testing"], dtype=bool)
        elif trainall:
            return None
        else:
            testing = np.asarray(self.pt.fitsnap_dict["Testing"], dtype=bool)
        if testing.shape[0] != n:
            raise ValueError("Testing mask has %d entries for %d rows" % (testing.shape[0], n))
        return testing if testing.any() else None

    def _shared_rows(self):
        """a, b, w when `perform_fit()` / `error_analysis()` are called without arrays (svd.py:42-44): the rows the
        drop-in calculator left on the device, unless the host copy of an array has been handed out since
        (`LazyHostMirror.exposed`: it may have been edited in place, e.g. bayesian_active_learning.py rescales `w`)
        -- that array is then re-read from the host.  Arrays filled by the stock calculator are host arrays."""
        sa = self.pt.shared_arrays
        dev = getattr(self.pt, "fitsnap_b200_device", None)
        sharded = self.process_group is not None
        out = []
        for name, key in (("a", "A"), ("b", "b"), ("w", "w")):
            m = sa[name]
            if (dev is not None and isinstance(m, LazyHostMirror) and not m.exposed and
                    (sharded or (dev["first_row"] == 0 and dev["n_rows"] == shared_shape(m)[0]))):
                out.append(dev[key])
            else:
                out.append(m.array)
        return out

    def _resolve_inputs(self, a, b, w, fs_dict, trainall):
        if a is None and b is None and w is None:          # svd.py:42-44
            a, b, w = self._shared_rows()
        if a is None or b is None or w is None:
            raise ValueError("a, b and w must be given together")
        n = a.shape[0]
        if isinstance(a, np.ndarray) and a.ndim == 1:      # StubsArray of width 1 (parallel_tools.py:1067)
            a = a.reshape(n, 1)
        return a, b, w, self._training_mask(n, fs_dict, trainall)

    def _engine(self):
        if self.engine is None:
            self.engine = _engine.default_engine()
        return self.engine

    def _to_device(self, a, b, w, testing):
        eng = self._engine()
        A = eng.to_device(a)
        B = eng.to_device(b).reshape(-1)
        W = eng.to_device(w).reshape(-1)
        T = None
        if testing is not None:
            T = eng.to_device(np.ascontiguousarray(testing, dtype=np.uint8), dtype=torch.uint8) \
                if not isinstance(testing, torch.Tensor) else testing.to(eng.device, torch.uint8)
        return A, B, W, T

    def _run_fit(self, A, B, W, T, alpha):
        eng = self._engine()
        if self.refine == "auto":
            res = _engine.fit_rows(eng, A, B, W, T, alpha=alpha, refine=2, group=self.process_group, diagnostics=True)
            # adaptive tail: keep refining while the correction still shrinks (ill-conditioned systems)
            last = float(res.last_correction) if res.last_correction is not None else 0.0
            rounds = 2
            while last > 1e-14 and rounds < self.max_refine:
                res2 = _engine.refine_rows(eng, A, B, W, T, res, group=self.process_group)
                new = float(res2.last_correction)
                rounds += 1
                if not (new <= last):       # the correction grew (or is NaN): keep the previous iterate
                    break
                res = res2
                if new > 0.5 * last:        # stalled at the attainable accuracy
                    break
                last = new
            res.extra["refine_rounds"] = rounds
            return res
        return _engine.fit_rows(eng, A, B, W, T, alpha=alpha, refine=int(self.refine), group=self.process_group,
                                diagnostics=False)

    def perform_fit(self, a=None, b=None, w=None, fs_dict=None, trainall=False):
        """Signature and result convention of svd.py:18 / ridge.py:11: `self.fit` <- np.float64 (K,)
        on rank 0; nothing is returned."""
        pt = self.pt
        sharded = self.process_group is not None
        if getattr(pt, "_rank", 0) != 0 and not sharded:      # svd.py:33: only rank 0 fits
            return
        a, b, w, testing = self._resolve_inputs(a, b, w, fs_dict, trainall)
        A, B, W, T = self._to_device(a, b, w, testing)
        res = self._run_fit(A, B, W, T, self._alpha())
        self.last_result = res
        self._check_info(res)
        if self.info["status"] != 0:
            res = self._rank_deficient(A, B, W, T, res)
        self.fit = res.coefficients()

    def _warn(self, msg):
        pr = getattr(self.pt, "single_print", None)
        (pr if pr is not None else print)(msg)

    def _rank_deficient(self, A, B, W, T, res):
        """The equilibrated Cholesky met a pivot below 64 k eps: the weighted design matrix has numerically
        dependent columns.  What the reference does then: `lstsq(aw, bw, 1e-13)` (svd.py:54) returns the
        minimum-norm solution, sklearn's Ridge falls back from Cholesky to an SVD-based solve of the same
        ridge problem.  Both are reproduced through the eigendecomposition of the Gram (`fsb_pinv_factor` on
        G + alpha I, refinement against A) and the user is told: the truncation acts on eigenvalues of
        G, i.e. singular values of aw below ~sqrt(k eps) sigma_max ~ 1e-7 sigma_max are cut, where gelsd cuts at
        1e-13 sigma_max."""
        alpha = self._alpha()
        k = res.gaug.shape[0] - 1
        self._warn("! WARNING (fitsnap_b200): %s: the weighted design matrix is numerically rank deficient (first "
                   "dependent column %d, %d dropped by the Cholesky factor); switching to the eigenvalue-truncated "
                   "%s solve of the Gram (singular values below ~%.0e sigma_max are cut; scipy's gelsd cuts at 1e-13)"
                   % (self.name, self.info["first_bad_column"], self.info["deficient"],
                      "minimum-norm" if alpha == 0.0 else "ridge", float(np.sqrt(32 * k * 2.220446049250313e-16))))
        if not self.min_norm_fallback:
            return res
        res2 = _engine.fit_rows_min_norm(self._engine(), A, B, W, T, res.gaug, refine=3, group=self.process_group,
                                         alpha=alpha)
        self.last_result = res2
        self.info["min_norm_rank"] = int(res2.info[0].item())
        return res2

    # -- error analysis ---------------------------------------------------
    def error_analysis(self, a=None, b=None, w=None, fs_dict=None):
        """Drop-in for the linear tail of `Solver.error_analysis` (solver.py:137-435, linear branch :368-435):
        `self.errors` gets the same (Group, Weighting, Testing, Subsystem) x (ncount, mae, rmse, rsq) table, and
        `_offset()` is applied for SNAP with bzeroflag (:432-435) -- but `preds = a @ fit` and the per-group sums come
        from ONE device pass over A (`fsb_group_stats`) instead of `DataFrame(a)` + groupby over every row and
        column.  `self.df` is built on demand and holds truths / preds / weights and the per-row lists (not the K
        descriptor columns).  When the frame itself is an output product (`[EXTRAS] dump_dataframe`) or the solver is
        not linear, the reference's implementation runs unchanged."""
        if not self.linear or _get(_section(self.config, "EXTRAS"), "dump_dataframe", False):
            return super().error_analysis(a=a, b=b, w=w, fs_dict=fs_dict)
        self.errors = []
        sharded = self.process_group is not None
        if getattr(self.pt, "_rank", 0) != 0 and not sharded:          # solver.py:160
            return
        if a is None and b is None and w is None and fs_dict is None:   # solver.py:368-372
            a, b, w = self._shared_rows()
            fs_dict = self.pt.fitsnap_dict
        fit0 = None if self.fit is None else np.array(self.fit, dtype=np.float64).reshape(-1)   # before _offset
        self._df, self._df_inputs = None, (a, b, w, fs_dict, fit0)
        true_multinode = bool(_get(_section(self.config, "SOLVER"), "true_multinode", False))
        if self.fit is not None and not true_multinode:
            from . import errors as _errors
            fit = np.asarray(self.fit, dtype=np.float64).reshape(-1)
            self.errors = _errors.linear_error_analysis(self._engine(), a, b, w, fs_dict, fit, group=self.process_group)
        if self.fit is not None:
            calc = _section(self.config, "CALCULATOR")
            bis = _section(self.config, "BISPECTRUM")
            if _get(calc, "calculator", None) == "LAMMPSSNAP" and _get(bis, "bzeroflag", False):
                self._offset()

    def error_analysis_device(self, a=None, b=None, w=None, fs_dict=None):
        """The device pass alone (no `_offset`, no frame): returns and stores the errors table."""
        from . import errors as _errors
        if a is None and b is None and w is None and fs_dict is None:
            a, b, w = self._shared_rows()
            fs_dict = self.pt.fitsnap_dict
        fit = np.asarray(self.fit, dtype=np.float64).reshape(-1)
        self.errors = _errors.linear_error_analysis(self._engine(), a, b, w, fs_dict, fit, group=self.process_group)
        return self.errors

    @property
    def df(self):
        """solver.py:374-382 frame, built on first access: truths, preds, weights + the per-row lists."""
        if getattr(self, "_df", None) is None and getattr(self, "_df_inputs", None) is not None:
            from . import errors as _errors
            a, b, w, fs_dict, fit = self._df_inputs
            self._df = _errors.light_frame(self._engine(), a, b, w, fs_dict, fit)
        return getattr(self, "_df", None)

    @df.setter
    def df(self, value):
        self._df, self._df_inputs = value, None

    def _offset(self):
        """solver.py:78-102 restated for stand-alone use of this mirror (no fitsnap3lib on the path): one zero
        coefficient in front of every type's block when bzeroflag = 1.  The classes built by plugin.register() take
        the reference's own `Solver._offset` instead."""
        nt = int(_get(_section(self.config, "BISPECTRUM"), "numtypes", 1))
        nc = int(_get(_section(self.config, "BISPECTRUM"), "ncoeff", 0))
        if nt > 1:
            fit = np.asarray(self.fit).reshape(nt, nc)
            self.fit = np.concatenate([np.zeros((nt, 1)), fit], axis=1).reshape((-1, 1))
        else:
            self.fit = np.insert(self.fit, 0, 0)
        fs = getattr(self, "fit_sam", None)
        if fs is not None:
            if nt > 1:
                nsam = fs.shape[0]
                fs3 = fs.reshape(nsam, nt, nc)
                self.fit_sam = np.concatenate([np.zeros((nsam, nt, 1)), fs3], axis=2).reshape(nsam, -1) + 0.0
            else:
                self.fit_sam = np.insert(fs, 0, 0, axis=1)

    def _check_info(self, res):
        info = res.info_host()
        self.info = {"status": int(info[0]), "first_bad_column": int(info[1]), "pinned": int(info[2]),
                     "deficient": int(info[3])}


class SVD(LinearSolverBase):
    """Drop-in for fitsnap3lib.solvers.svd.SVD: minimum of |aw x - bw| (scipy.linalg.lstsq(aw, bw,
    1e-13) in the reference, svd.py:54).  `[EXTRAS] apply_transpose` (svd.py:48-53) asks the
    reference to solve the same problem through aw^T aw -- which is what this path always does --
    so it changes nothing here."""

    def __init__(self, name, pt, config):
        super().__init__(name, pt, config)


class RIDGE(LinearSolverBase):
    """Drop-in for fitsnap3lib.solvers.ridge.RIDGE: (aw^T aw + alpha I) x = aw^T bw with
    alpha = [RIDGE] alpha (solver_sections/ridge.py:13).  `local_solver` selects between two
    implementations of the same formula in the reference (sklearn vs regressor.py:10-16); both
    map to the same device solve."""

    def __init__(self, name, pt, config):
        super().__init__(name, pt, config)

    def _alpha(self):
        sec = _section(self.config, "RIDGE")
        return float(_get(sec, "alpha", 1.0e-8))

    def perform_fit(self, a=None, b=None, w=None, fs_dict=None, trainall=False):
        extras = _section(self.config, "EXTRAS")
        if _get(extras, "apply_transpose", False):
            # ridge.py:41-43 + sklearn Ridge.fit(aw^T aw, aw^T bw): ridge on the NORMAL matrix,
            # (C^T C + alpha I) x = C^T d with C = aw^T aw, d = aw^T bw (SURVEY 3.3 semantic trap).
            pt = self.pt
            if getattr(pt, "_rank", 0) != 0 and self.process_group is None:
                return
            a, b, w, testing = self._resolve_inputs(a, b, w, fs_dict, trainall)
            A, B, W, T = self._to_device(a, b, w, testing)
            eng = self._engine()
            gaug = eng.gram(A, B, W, T)
            _engine._all_reduce(gaug, self.process_group, eng)
            k = gaug.shape[0] - 1
            C = gaug[:k, :k].contiguous()
            d = gaug[:k, k].contiguous()
            ones = torch.ones(k, dtype=torch.float64, device=eng.device)
            res = _engine.fit_rows(eng, C, d, ones, None, alpha=self._alpha(), refine=2, diagnostics=False)
            self.last_result = res
            self._check_info(res)
            self.fit = res.coefficients()
            return
        super().perform_fit(a, b, w, fs_dict, trainall)


class LASSO(LinearSolverBase):
    """Drop-in for fitsnap3lib.solvers.lasso.LASSO: argmin 1/(2n)|aw x - bw|^2 + alpha |x|_1 with
    alpha = [LASSO] alpha, at most [LASSO] max_iter sweeps (solver_sections/lasso.py:8-14).  The
    reference runs sklearn's coordinate descent on the tall system (or, with `apply_transpose`, on
    (aw^T aw, aw^T bw) -- lasso.py:22-24); here the Gram is formed once on the device and the
    coordinate descent runs on the k x k problem.  Iterated to a tight tolerance (1e-12 relative
    coordinate change), i.e. closer to the minimiser than sklearn's default tol = 1e-4 stop."""

    tol = 1e-12

    def __init__(self, name, pt, config):
        super().__init__(name, pt, config)

    def perform_fit(self, a=None, b=None, w=None, fs_dict=None, trainall=False):
        if getattr(self.pt, "_rank", 0) != 0 and self.process_group is None:
            return
        sec = _section(self.config, "LASSO")
        alpha = float(_get(sec, "alpha", 1.0e-8))
        max_iter = int(_get(sec, "max_iter", 2000))
        a, b, w, testing = self._resolve_inputs(a, b, w, fs_dict, trainall)
        A, B, W, T = self._to_device(a, b, w, testing)
        eng = self._engine()
        gaug = eng.gram(A, B, W, T)
        n_train = A.shape[0] if T is None else int(A.shape[0] - int(T.sum().item()))
        if self.process_group is not None:
            _engine._all_reduce(gaug, self.process_group, eng)
            nt = torch.tensor([n_train], dtype=torch.float64, device=gaug.device)
            _engine._all_reduce(nt, self.process_group, eng)
            n_train = int(nt.item())
        extras = _section(self.config, "EXTRAS")
        if _get(extras, "apply_transpose", False):
            # lasso.py:22-24: sklearn then sees X = aw^T aw (k rows), y = aw^T bw
            k = gaug.shape[0] - 1
            C = gaug[:k, :k].contiguous()
            d = gaug[:k, k].contiguous()
            gaug = eng.gram(C, d, torch.ones(k, dtype=torch.float64, device=C.device), None)
            n_train = k
        x, info = eng.lasso(gaug, n_train, alpha, max_iter, self.tol)
        self.info = {"not_converged": int(info[0].item()), "sweeps": int(info[1].item())}
        self.fit = x.detach().cpu().numpy().astype(np.float64, copy=True)


class ANL(LinearSolverBase):
    """Drop-in for fitsnap3lib.solvers.anl.ANL (analytical Bayesian linear regression, anl.py:13-67):
        invptp = sym(pinv(aw^T aw + nugget I)),   mean = invptp aw^T bw,   cov = sigmahat * invptp,
        sigmahat = (|bw - aw mean|^2 / 2) / ((n_train - k)/2 - 1),   fit_sam ~ N(mean, cov) x nsam,
    with nugget = [SOLVER] cov_nugget and nsam = [SOLVER] nsam; like the reference it drops `covariance.npy`
    and `mean.npy` into the working directory (anl.py:60-61).
    Device side: the two passes over A -- the fused mask/weight/Gram pass (`fsb_gram`) and |res|^2 with the
    training row count (`fsb_group_stats`).  Everything between them is k x k algebra on the all-reduced Gram and
    is kept as anl.py:41-44 states it (numpy `pinv`, default rcond 1e-15 ON THE EIGENVALUES of the Gram, i.e. a
    rank truncation at sigma/sigma_max ~ 3e-8 -- replacing it by the refined Cholesky solve of the SVD/RIDGE
    drop-ins would change the answer whenever that truncation bites).  `[EXTRAS] apply_transpose` (anl.py:32-37,
    marked "probably nonsense" there: it squares the Gram and ends with a negative variance) is not mirrored."""

    #: write covariance.npy / mean.npy like the reference (tests switch it off)
    save_files = True

    def __init__(self, name, pt, config):
        super().__init__(name, pt, config)
        self.cov = None
        self.fit_sam = None

    def _alpha(self):
        return float(_get(_section(self.config, "SOLVER"), "cov_nugget", 0.0))

    def perform_fit(self, a=None, b=None, w=None, trainall=False, fs_dict=None):
        pt = self.pt
        sharded = self.process_group is not None
        if getattr(pt, "_rank", 0) != 0 and not sharded:      # anl.py:14 sub_rank_zero
            return
        a, b, w, testing = self._resolve_inputs(a, b, w, fs_dict, trainall)
        A, B, W, T = self._to_device(a, b, w, testing)
        eng = self._engine()
        gaug = eng.gram(A, B, W, T)
        if sharded:
            _engine._all_reduce(gaug, self.process_group, eng)
        k = A.shape[1]
        gh = gaug.cpu().numpy()
        nugget = self._alpha()
        invptp = np.linalg.pinv(gh[:k, :k] + nugget * np.diag(np.ones((k,))))     # anl.py:41
        invptp = invptp * 0.5 + invptp.T * 0.5                                   # anl.py:42
        self.fit = np.dot(invptp, gh[:k, k])                                     # anl.py:44
        # |bw - aw mean|^2 and the training row count from one pass: group 0 = training, 1 = test rows
        gid = torch.zeros(A.shape[0], dtype=torch.int32, device=eng.device) if T is None else T.to(torch.int32)
        stats = eng.group_stats(A, B, W, gid, eng.to_device(self.fit), 2)
        if sharded:
            _engine._all_reduce(stats, self.process_group, eng)
        stats = stats.cpu().numpy()
        npt, wsq = float(stats[0, 0]), float(stats[0, 7])
        sigmahat = (wsq / 2.0) / ((npt - k) / 2.0 - 1.0)                         # anl.py:48-52
        self.cov = sigmahat * invptp                                             # anl.py:56
        if self.save_files and getattr(pt, "_rank", 0) == 0:
            np.save("covariance.npy", self.cov)
            np.save("mean.npy", self.fit)
        nsam = int(_get(_section(self.config, "SOLVER"), "nsam", 0))
        if nsam:
            self.fit_sam = np.random.multivariate_normal(self.fit, self.cov, size=(nsam,))   # anl.py:65
