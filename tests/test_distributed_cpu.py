"""CPU, world_size 2, gloo: the N > 1 host logic -- row sharding, ONE all-reduce of the packed
Gram, replicated solve, all-reduced refinement residual -- gives the same coefficients as the
single-process oracle on the full matrix.  The arithmetic is a test double (tests/fake_engine.py);
the real kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fitsnap_b200.distributed import shard_configs_by_rows, shard_rows
from oracle import linear_fit as lf
from tests.synth import SOLVE_CASES, synth_system


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, alpha, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fitsnap_b200.engine import fit_rows
    from tests.fake_engine import OracleEngine
    a, b, w, t = synth_system(**SOLVE_CASES[case])
    lo, hi = shard_rows(a.shape[0], world, rank)
    A, B, W = torch.from_numpy(a[lo:hi]), torch.from_numpy(b[lo:hi]), torch.from_numpy(w[lo:hi])
    T = torch.from_numpy(t[lo:hi].astype(np.uint8))
    res = fit_rows(OracleEngine(), A, B, W, T, alpha=alpha, refine=2, group=dist.group.WORLD)
    np.save(os.path.join(out_dir, "x_%d.npy" % rank), res.x.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("case,alpha", [("well", 0.0), ("ill", 0.0), ("wide", 1e-6)])
def test_row_sharded_fit_matches_single_process(tmp_path, case, alpha):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, alpha, str(tmp_path)), nprocs=world, join=True)
    xs = [np.load(tmp_path / ("x_%d.npy" % r)) for r in range(world)]
    assert np.array_equal(xs[0], xs[1])                    # replicated solve: identical on every rank
    a, b, w, t = synth_system(**SOLVE_CASES[case])
    ref = lf.svd_fit(a, b, w, t) if alpha == 0.0 else lf.ridge_fit_exact(a, b, w, alpha, t)
    assert lf.coeff_rel_err(xs[0], ref)[0] < 1e-10


def _anl_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from fitsnap_b200.solvers import ANL
    from tests.fake_engine import OracleEngine
    a, b, w, t = synth_system(**SOLVE_CASES["well"])
    lo, hi = shard_rows(a.shape[0], world, rank)
    pt = SimpleNamespace(_rank=rank, shared_arrays={}, fitsnap_dict={"Testing": [bool(v) for v in t[lo:hi]]})
    cfg = SimpleNamespace(sections={"SOLVER": SimpleNamespace(cov_nugget=1e-8, nsam=0)})
    s = ANL("ANL", pt, cfg)
    s.engine = OracleEngine()
    s.process_group = dist.group.WORLD
    s.save_files = False
    s.perform_fit(a=a[lo:hi], b=b[lo:hi], w=w[lo:hi])
    np.savez(os.path.join(out_dir, "anl_%d.npz" % rank), mean=s.fit, cov=s.cov)
    dist.destroy_process_group()


def test_row_sharded_anl_matches_single_process(tmp_path):
    """ANL over two row shards: all-reduced Gram and all-reduced residual sums give the single-process
    posterior mean and covariance (anl.py:40-58) on every rank."""
    world = 2
    mp.spawn(_anl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = [np.load(tmp_path / ("anl_%d.npz" % r)) for r in range(world)]
    assert np.array_equal(out[0]["mean"], out[1]["mean"]) and np.array_equal(out[0]["cov"], out[1]["cov"])
    a, b, w, t = synth_system(**SOLVE_CASES["well"])
    mean, cov = lf.anl_fit(a, b, w, 1e-8, t)
    assert np.max(np.abs(out[0]["mean"] - mean)) < 1e-9 * np.max(np.abs(mean))
    assert np.max(np.abs(out[0]["cov"] - cov)) < 1e-7 * np.max(np.abs(cov))


def _stream_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fitsnap_b200.pipeline import StreamingLinearFit
    from tests.fake_engine import OracleEngine
    a, b, w, t = synth_system(**SOLVE_CASES["ill"])
    lo, hi = shard_rows(a.shape[0], world, rank)
    cuts = np.linspace(lo, hi, 4 + rank).astype(int)          # a different number of chunks on every rank
    chunks = [(a[i:j], b[i:j], w[i:j], t[i:j]) for i, j in zip(cuts[:-1], cuts[1:])]
    res = StreamingLinearFit(alpha=1e-6, refine=2, group=dist.group.WORLD, engine=OracleEngine()).fit(chunks)
    np.save(os.path.join(out_dir, "xs_%d.npy" % rank), res.x.numpy())
    dist.destroy_process_group()


def test_row_sharded_streaming_fit_matches_single_process(tmp_path):
    """Out-of-core mode under sharding: every rank streams its own rows in its own chunking; one all-reduce of the
    accumulated Gram and one per refinement round."""
    world = 2
    mp.spawn(_stream_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    xs = [np.load(tmp_path / ("xs_%d.npy" % r)) for r in range(world)]
    assert np.array_equal(xs[0], xs[1])
    a, b, w, t = synth_system(**SOLVE_CASES["ill"])
    assert lf.coeff_rel_err(xs[0], lf.ridge_fit_exact(a, b, w, 1e-6, t))[0] < 1e-10


def test_shard_rows_partition():
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            cuts = [shard_rows(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_configs_by_rows_balances_rows():
    rng = np.random.default_rng(0)
    natoms = rng.integers(1, 258, 5000)
    rows = 7 + 3 * natoms
    for world in (2, 4, 8):
        bnd = shard_configs_by_rows(rows, world)
        assert bnd[0] == 0 and bnd[-1] == len(rows) and np.all(np.diff(bnd) >= 0)
        per = np.array([rows[bnd[i]:bnd[i + 1]].sum() for i in range(world)])
        assert per.sum() == rows.sum()
        assert per.max() - per.min() <= 2 * rows.max()


def _plugin_worker(rank, world, port, out_dir):
    """One FitSnap-style flow per rank through the REFERENCE's factories after plugin.register(): every rank assembles
    its own share of the configurations, `distributed.attach` makes perform_fit / error_analysis row-sharded."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ref_driver as rd
    rd.install_fake_lammps()
    from fitsnap_b200 import distributed, plugin
    from tests.fake_engine import OracleEngine
    eng = OracleEngine()
    plugin.register(engine=eng)
    from fitsnap3lib.solvers.solver_factory import solver
    cfgs, blocks, vols, kw = _plugin_inputs()
    bnd = shard_configs_by_rows([7 + 3 * c["NumAtoms"] for c in cfgs], world)
    lo, hi = int(bnd[rank]), int(bnd[rank + 1])
    a, b, w, lists, cfg, pt, calc = rd.ref_scatter(cfgs[lo:hi], blocks[lo:hi], vols[lo:hi], use_factory=True, **kw)
    s = solver("SVD", pt, cfg)
    s.refine = 2
    distributed.attach(s, engine=eng)
    s.perform_fit()
    fit = s.fit.copy()
    s.error_analysis()
    s.errors.to_pickle(os.path.join(out_dir, "err_%d.pkl" % rank))
    np.save(os.path.join(out_dir, "fit_%d.npy" % rank), fit)
    dist.destroy_process_group()


def _plugin_inputs():
    from oracle import ref_driver as rd
    rng = np.random.default_rng(17)
    kw = dict(numtypes=1, types="Ta", twojmax="4", bzeroflag=0)
    _pt, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = [], [], []
    for i in range(60):
        n = int(rng.integers(1, 9))
        cfgs.append(rd.make_config_dict(n, 1, rng, ["Ta"], group="g%d" % (i % 3), fname="f%d" % i,
                                        eweight=float(10 ** rng.uniform(-2, 2)), fweight=float(10 ** rng.uniform(-2, 2)),
                                        vweight=float(10 ** rng.uniform(-9, -5)), test_bool=bool(i % 4 == 1)))
        blocks.append(rng.standard_normal((1 + 3 * n + 6, nc + 1)))
        vols.append(float(rng.uniform(20, 400)))
    return cfgs, blocks, vols, kw


def test_sharded_fit_through_the_reference_factories(tmp_path):
    import pandas as pd
    from oracle import ref_driver as rd
    if not rd.reference_available():
        pytest.skip("reference tree only exists in the build container")
    world = 2
    mp.spawn(_plugin_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    fits = [np.load(tmp_path / ("fit_%d.npy" % r)) for r in range(world)]
    assert np.array_equal(fits[0], fits[1])
    cfgs, blocks, vols, kw = _plugin_inputs()
    a, b, w, lists, *_ = rd.ref_scatter(cfgs, blocks, vols, **kw)                     # stock classes, one process
    x_ref, sref = rd.ref_fit("SVD", a, b, w, testing=np.array(lists["Testing"]))
    assert lf.coeff_rel_err(fits[0], x_ref)[0] < 1e-9
    sref.pt.fitsnap_dict.update({k_: v for k_, v in lists.items()})
    sref.error_analysis()
    errs = [pd.read_pickle(tmp_path / ("err_%d.pkl" % r)) for r in range(world)]
    assert errs[0].equals(errs[1])
    ref = sref.errors
    assert list(errs[0].index) == list(ref.index)
    assert np.array_equal(errs[0]["ncount"].values, ref["ncount"].values)
    for col in ("mae", "rmse", "rsq"):
        r, d = ref[col].values.astype(float), errs[0][col].values.astype(float)
        ok = np.isclose(d, r, rtol=1e-7, atol=1e-12) | (np.isnan(d) & np.isnan(r)) | (~np.isfinite(r) & ~np.isfinite(d))
        assert ok.all(), col
