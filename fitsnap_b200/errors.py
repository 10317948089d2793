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
    import pandas as pd
    if n_rows == 0:
        return np.zeros(0, dtype=np.int32), []
    gc, gu = pd.factorize(np.asarray(groups, dtype=object))
    rc, ru = pd.factorize(np.asarray(rtype, dtype=object))
    tc = np.asarray(testing, dtype=bool).astype(np.int64)
    code = (gc.astype(np.int64) * 2 + tc) * len(ru) + rc
    uniq, gid = np.unique(code, return_inverse=True)
    keys = [(gu[u // (2 * len(ru))], bool((u // len(ru)) % 2), ru[u % len(ru)]) for u in uniq]
    return gid.astype(np.int32), keys


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


def _global_keys(keys, group):
    """Union of the (group, test, row type) keys of all ranks, in one order on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, keys, group=group)
    seen, out = set(), []
    for lst in gathered:
        for key in lst:
            if key not in seen:
                seen.add(key)
                out.append(key)
    return sorted(out)


def linear_error_analysis(engine, A, b, w, fs_dict, x, group=None):
    """Device pass + host formatting.  A, b, w, x: device tensors (or host arrays: uploaded).  With `group` (row-sharded
    fit, one shard per rank) the keys are unified across ranks, the per-group sums all-reduced, and every rank ends
    with the table of the whole data set."""
    import torch
    A = engine.to_device(A)
    b = engine.to_device(b).reshape(-1)
    w = engine.to_device(w).reshape(-1)
    x = engine.to_device(np.asarray(x, dtype=np.float64).reshape(-1)) if not isinstance(x, torch.Tensor) else x
    gid, keys = encode_groups(fs_dict, A.shape[0])
    sharded = False
    if group is not None:
        import torch.distributed as dist
        sharded = dist.is_initialized() and dist.get_world_size(group) > 1
    if sharded:
        all_keys = _global_keys(keys, group)
        pos = {key: i for i, key in enumerate(all_keys)}
        gid = np.asarray([pos[key] for key in keys], dtype=np.int32)[gid] if len(keys) else gid
        keys = all_keys
    stats = engine.group_stats(A, b, w, engine.to_device(gid, dtype=torch.int32), x, len(keys))
    if sharded:
        from .engine import _all_reduce
        _all_reduce(stats, group, engine)
    return errors_frame(stats.cpu().numpy(), keys)


def light_frame(engine, A, b, w, fs_dict, x):
    """The frame of solver.py:374-382 without its K descriptor columns: truths, preds (= a @ fit from the device),
    weights and every per-row list of `fs_dict`."""
    import pandas as pd
    import torch
    n = A.shape[0]
    to_host = lambda v: v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    df = pd.DataFrame({"truths": to_host(b).reshape(-1)})
    if x is not None:
        Ad = engine.to_device(A)
        xd = engine.to_device(np.asarray(x, dtype=np.float64).reshape(-1))
        df["preds"] = engine.predict(Ad, xd).cpu().numpy()
    df["weights"] = to_host(w).reshape(-1)
    for key, val in (fs_dict or {}).items():
        if isinstance(val, list) and len(val) == n:
            df[key] = val
    return df
