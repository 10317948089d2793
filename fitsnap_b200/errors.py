"""Linear error analysis from device-side sums.

The reference builds a pandas frame of ALL rows and columns (`DataFrame(a)`), adds `preds = a @ fit`
and groups by (Groups, Testing, Row_Type) to get ncount / MAE / RMSE / R^2, weighted and unweighted
(fitsnap3lib/solvers/solver.py:108-133, 368-429).  At 1e7 rows that frame dwarfs the solve.  Here
one kernel pass (`fsb_group_stats`) returns ten sums per group, the metrics are closed-form
functions of those sums, and the result is laid out exactly like `Solver.errors`.
"""
from __future__ import annotations

import numpy as np

STAT_NAMES = ("n", "abs", "sq", "t", "tt", "nw", "wabs", "wsq", "wt", "wtwt")


def encode_groups(fs_dict, n_rows):
    """(Groups, Testing, Row_Type) per row -> dense int32 ids + the list of keys (solver.py:391-393)."""
    groups = fs_dict["Groups"]
    testing = fs_dict["Testing"]
    rtype = fs_dict["Row_Type"]
    assert len(groups) == len(testing) == len(rtype) == n_rows
    lut, keys = {}, []
    gid = np.empty(n_rows, dtype=np.int32)
    for i in range(n_rows):
        key = (groups[i], bool(testing[i]), rtype[i])
        j = lut.get(key)
        if j is None:
            j = lut[key] = len(keys)
            keys.append(key)
        gid[i] = j
    return gid, keys


def metrics_from_sums(s):
    """solver.py:108-133 `_ncount_mae_rmse_rsq_unweighted_and_weighted` expressed through sums."""
    n, sabs, ssq, st, stt, nw, wabs, wsq, swt, swtwt = (float(v) for v in s)
    with np.errstate(divide="ignore", invalid="ignore"):
        un = dict(ncount=int(round(n)), mae=sabs / n, rmse=np.sqrt(ssq / n),
                  rsq=1.0 - ssq / (stt - st * st / n))
        m = swt / nw if nw else np.nan                       # (w t / w_nconfig).sum()
        we = dict(ncount=int(round(nw)), mae=wabs / n, rmse=np.sqrt(wsq / nw) if nw else np.nan,
                  rsq=(1.0 - wsq / (swtwt - 2.0 * m * swt + n * m * m)) if nw else np.nan)
    return un, we


def errors_frame(stats, keys):
    """Lay the metrics out like `Solver.errors` (solver.py:395-429): index
    (Group, Weighting, Testing, Subsystem), '*ALL' block first, columns ncount/mae/rmse/rsq."""
    import pandas as pd
    stats = np.asarray(stats, dtype=np.float64)
    rows = {}
    all_sums = {}
    for (grp, test, rt), s in zip(keys, stats):
        un, we = metrics_from_sums(s)
        rows[(grp, "Unweighted", test, rt)] = un
        rows[(grp, "weighted", test, rt)] = we
        all_sums[(test, rt)] = all_sums.get((test, rt), 0.0) + s
    all_rows = {}
    for (test, rt), s in all_sums.items():
        un, we = metrics_from_sums(s)
        all_rows[("*ALL", "Unweighted", test, rt)] = un
        all_rows[("*ALL", "weighted", test, rt)] = we
    order = sorted(all_rows) + sorted(rows)
    data = {**all_rows, **rows}
    idx = pd.MultiIndex.from_tuples([(g, wgt, "Testing" if t else "Training", rt) for g, wgt, t, rt in order],
                                    names=["Group", "Weighting", "Testing", "Subsystem"])
    df = pd.DataFrame([data[o] for o in order], index=idx, columns=["ncount", "mae", "rmse", "rsq"])
    df["ncount"] = df["ncount"].astype(int)
    return df


def linear_error_analysis(engine, A, b, w, fs_dict, x):
    """Device pass + host formatting.  A, b, w, x: device tensors (or host arrays: uploaded)."""
    import torch
    A = engine.to_device(A)
    b = engine.to_device(b).reshape(-1)
    w = engine.to_device(w).reshape(-1)
    x = engine.to_device(np.asarray(x, dtype=np.float64).reshape(-1)) if not isinstance(x, torch.Tensor) else x
    gid, keys = encode_groups(fs_dict, A.shape[0])
    stats = engine.group_stats(A, b, w, engine.to_device(gid, dtype=torch.int32), x, len(keys))
    return errors_frame(stats.cpu().numpy(), keys)
