"""Generate tests/golden/* by running the UNMODIFIED reference in the build container.

    python -m oracle.make_golden

Needs /root/reference (see oracle/ref_driver.py); the produced fixtures are committed so
that the GPU box (which has no reference tree) can check against them.

Fixtures
  ta_linear.npz       the reference's own golden triple (examples/Ta_Linear_JCP2014/
                      20May21_Standard/{Descriptors,Truth-Ref,Weights}.npy), the 31 golden
                      coefficients of Ta_pot.snapcoeff, and the reference SVD / RIDGE /
                      LASSO results on that triple (reference classes, this container).
  scatter_*.npz       synthetic raw LAMMPS blocks + the (A, b, w, Testing) the unmodified
                      LammpsSnap / LammpsPace calculators assemble from them.
  single_*.npz        the same kind of blocks through `process_single` (lammps_base.py:101-125): per-configuration
                      (a, b, w) of the unmodified `_collect_lammps_single`, incl. switched-off row families (zero rows)
                      and missing weight keys (default 1.0).
  solve_*.npz         reference SVD / RIDGE fits of the seeded synthetic systems of
                      tests/synth.py (well- and ill-conditioned, zero columns, k > 128);
                      the systems themselves are regenerated from the seed at test time.
"""
from __future__ import annotations

import os

import numpy as np

from . import ref_driver as rd
from tests.synth import SOLVE_CASES, synth_system

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TA_DIR = os.path.join(rd.REFERENCE_ROOT, "examples", "Ta_Linear_JCP2014", "20May21_Standard")


def read_snapcoeff(path):
    vals = []
    for line in open(path).read().splitlines()[4:]:
        s = line.strip()
        if s and not s.startswith("#"):
            vals.append(float(s.split()[0]))
    return np.array(vals, dtype=np.float64)


def ta_linear():
    a = np.load(os.path.join(TA_DIR, "Descriptors.npy"))
    b = np.load(os.path.join(TA_DIR, "Truth-Ref.npy"))
    w = np.load(os.path.join(TA_DIR, "Weights.npy"))
    gold = read_snapcoeff(os.path.join(TA_DIR, "Ta_pot.snapcoeff"))
    svd, _ = rd.ref_fit("SVD", a, b, w)
    ridge, _ = rd.ref_fit("RIDGE", a, b, w, ridge_alpha=1e-6)
    ridge_local, _ = rd.ref_fit("RIDGE", a, b, w, ridge_alpha=1e-6, ridge_local=1)
    # a deterministic test split (every 7th row is a test row) through fs_dict
    testing = (np.arange(a.shape[0]) % 7 == 3)
    svd_split, _ = rd.ref_fit("SVD", a, b, w, testing=testing)
    lasso, _ = rd.ref_fit("LASSO", a, b, w, lasso_alpha=1e-6, lasso_max_iter=20000)
    np.savez_compressed(os.path.join(OUT, "ta_linear.npz"), a=a, b=b, w=w, snapcoeff=gold, ref_svd=svd,
                        ref_ridge_1e6=ridge, ref_ridge_local_1e6=ridge_local, testing=testing,
                        ref_svd_split=svd_split, ref_lasso_1e6=lasso)
    print("ta_linear: ref-vs-gold max abs", np.max(np.abs(svd - gold)))


def solve_cases():
    """Only the reference coefficient vectors are stored; (A, b, w, Testing) are regenerated from
    the seeded generator in tests/synth.py (checked through the stored checksums)."""
    for name, kw in SOLVE_CASES.items():
        a, b, w, testing = synth_system(**kw)
        svd, _ = rd.ref_fit("SVD", a, b, w, testing=testing)
        svd_all, _ = rd.ref_fit("SVD", a, b, w)
        ridge, _ = rd.ref_fit("RIDGE", a, b, w, testing=testing, ridge_alpha=1e-6)
        np.savez_compressed(os.path.join(OUT, "solve_%s.npz" % name), ref_svd=svd, ref_svd_all=svd_all,
                            ref_ridge_1e6=ridge, checksum=np.array([a.sum(), b.sum(), w.sum(), testing.sum()]))
        print("solve_%s" % name, a.shape)


def anl_cases():
    """Posterior mean + covariance of the reference's ANL solver (anl.py) on two seeded systems."""
    for name, nugget in (("well", 0.0), ("zerocol", 1e-6)):
        a, b, w, testing = synth_system(**SOLVE_CASES[name])
        mean, cov = rd.ref_anl(a, b, w, testing=testing, cov_nugget=nugget)
        np.savez_compressed(os.path.join(OUT, "anl_%s.npz" % name), ref_mean=mean, ref_cov=cov,
                            cov_nugget=np.float64(nugget),
                            checksum=np.array([a.sum(), b.sum(), w.sum(), testing.sum()]))
        print("anl_%s" % name, cov.shape)


def scatter_cases():
    rng = np.random.default_rng(77)
    combos = [
        ("snap_b0_efs", dict(bzeroflag=0, twojmax="6 6", energy=1, force=1, stress=1)),
        ("snap_b1_efs", dict(bzeroflag=1, twojmax="6 4", energy=1, force=1, stress=1)),
        ("snap_b0_ef", dict(bzeroflag=0, twojmax="6 4", energy=1, force=1, stress=0)),
        ("snap_b1_es", dict(bzeroflag=1, twojmax="6 6", energy=1, force=0, stress=1)),
        ("snap_b0_f", dict(bzeroflag=0, twojmax="4 4", energy=0, force=1, stress=0)),
    ]
    names = ["In", "P"]
    for tag, kw in combos:
        pt, cfg = rd.make_reference_context(numtypes=2, types="In P", **kw)
        sec = cfg.sections["BISPECTRUM"]
        nc, b2j, tm = sec.ncoeff, np.array(sec.blank2J, dtype=np.float64), sec.type_mapping
        cfgs, blocks, vols = [], [], []
        for i in range(11):
            n = int(rng.integers(1, 14))
            c = rd.make_config_dict(n, 2, rng, names, group="grp%d" % (i % 4), fname="cfg%d.json" % i,
                                    eweight=float(10 ** rng.uniform(-3, 3)), fweight=float(10 ** rng.uniform(-3, 3)),
                                    vweight=float(10 ** rng.uniform(-12, -3)), test_bool=bool(i % 5 == 2))
            cfgs.append(c)
            blocks.append(rng.standard_normal((1 + 3 * n + 6, nc * 2 + 1)) * 10.0 ** rng.uniform(-4, 4, (1, nc * 2 + 1)))
            vols.append(float(rng.uniform(20, 4000)))
        a, b, w, lists, *_ = rd.ref_scatter(cfgs, blocks, vols, numtypes=2, types="In P", **kw)
        save_scatter(tag, cfgs, blocks, vols, a, b, w, lists, nc, 2, kw["bzeroflag"], b2j, tm, kw)

    # ACE-shaped (LammpsPace) case: [ACE] attributes injected (SURVEY 8c ACE caveat)
    nc, nt = 23, 2
    for tag, bz in (("pace_b0_efs", 0), ("pace_b1_efs", 1)):
        k = nc * nt + (0 if bz else nt)
        b2j = np.ones(k)
        b2j[rng.choice(k, 4, replace=False)] = 0.0
        tm = {"In": 1, "P": 2}
        ace = dict(numtypes=nt, ncoeff=nc, bzeroflag=bz, bikflag=0, dgradflag=0, blank2J=b2j, type_mapping=tm,
                   rcutfac=[4.0])
        cfgs, blocks, vols = [], [], []
        for i in range(9):
            n = int(rng.integers(1, 12))
            c = rd.make_config_dict(n, nt, rng, names, group="g%d" % (i % 3), fname="p%d" % i,
                                    eweight=float(10 ** rng.uniform(-2, 2)), fweight=float(10 ** rng.uniform(-2, 2)),
                                    vweight=float(10 ** rng.uniform(-9, -5)), test_bool=bool(i % 6 == 1))
            cfgs.append(c)
            blocks.append(rng.standard_normal((1 + 3 * n + 6, nc * nt + 1)) * 10.0 ** rng.uniform(-3, 3, (1, nc * nt + 1)))
            vols.append(float(rng.uniform(20, 4000)))
        kw = dict(bzeroflag=bz, twojmax="6 6", energy=1, force=1, stress=1)
        a, b, w, lists, *_ = rd.ref_scatter(cfgs, blocks, vols, calculator="LAMMPSPACE", ace=ace, numtypes=nt,
                                           types="In P", **kw)
        save_scatter(tag, cfgs, blocks, vols, a, b, w, lists, nc, nt, bz, b2j, tm, kw)


def save_scatter(tag, cfgs, blocks, vols, a, b, w, lists, ncoeff, numtypes, bzeroflag, b2j, tm, kw):
    natoms = np.array([c["NumAtoms"] for c in cfgs], dtype=np.int32)
    tf = np.zeros((len(cfgs), numtypes))
    for i, c in enumerate(cfgs):
        for at in c["AtomTypes"]:
            tf[i, tm[at] - 1] += 1
        tf[i] /= len(c["AtomTypes"])
    np.savez_compressed(
        os.path.join(OUT, "scatter_%s.npz" % tag),
        raw=np.concatenate(blocks, 0), natoms=natoms, volume=np.array(vols),
        energy=np.array([c["Energy"] for c in cfgs]),
        forces=np.concatenate([c["Forces"].reshape(-1) for c in cfgs]),
        stress=np.stack([c["Stress"] for c in cfgs]),
        eweight=np.array([c["eweight"] for c in cfgs]), fweight=np.array([c["fweight"] for c in cfgs]),
        vweight=np.array([c["vweight"] for c in cfgs]), type_fraction=tf, blank2j=b2j,
        ncoeff=ncoeff, numtypes=numtypes, bzeroflag=bzeroflag,
        use_energy=kw["energy"], use_force=kw["force"], use_stress=kw["stress"],
        test_bool=np.array([c["test_bool"] for c in cfgs]),
        ref_a=a, ref_b=b, ref_w=w, ref_testing=np.array(lists["Testing"], dtype=bool),
        ref_row_type=np.array(lists["Row_Type"]))
    print("scatter_%s" % tag, a.shape, "zero cols:", int((b2j == 0).sum()))


def single_cases():
    """`process_single` outputs of the unmodified reference (lammps_snap.py:224-389, lammps_pace.py:197-366)."""
    rng = np.random.default_rng(99)
    names = ["In", "P"]
    snap = [
        ("snap_b0_efs", dict(bzeroflag=0, twojmax="6 6", energy=1, force=1, stress=1), False),
        ("snap_b1_ef", dict(bzeroflag=1, twojmax="6 4", energy=1, force=1, stress=0), False),
        ("snap_b0_es", dict(bzeroflag=0, twojmax="4 4", energy=1, force=0, stress=1), False),
        ("snap_b1_fs", dict(bzeroflag=1, twojmax="4 4", energy=0, force=1, stress=1), True),
    ]
    for tag, kw, drop in snap:
        pt, cfg = rd.make_reference_context(numtypes=2, types="In P", **kw)
        sec = cfg.sections["BISPECTRUM"]
        nc, b2j, tm = sec.ncoeff, np.array(sec.blank2J, dtype=np.float64), sec.type_mapping
        cfgs, blocks, vols = _single_inputs(rng, nc, 2, names, 7)
        out, _ = rd.ref_single(cfgs, blocks, vols, numtypes=2, types="In P", drop_weights=drop, **kw)
        save_single(tag, cfgs, blocks, vols, out, nc, 2, kw["bzeroflag"], b2j, tm, kw, drop)
    nc, nt = 23, 2
    for tag, bz in (("pace_b0_efs", 0), ("pace_b1_ef", 1)):
        k = nc * nt + (0 if bz else nt)
        b2j = np.ones(k)
        b2j[rng.choice(k, 3, replace=False)] = 0.0
        tm = {"In": 1, "P": 2}
        ace = dict(numtypes=nt, ncoeff=nc, bzeroflag=bz, bikflag=0, dgradflag=0, blank2J=b2j, type_mapping=tm,
                   rcutfac=[4.0])
        kw = dict(bzeroflag=bz, twojmax="6 6", energy=1, force=1, stress=0 if bz else 1)
        cfgs, blocks, vols = _single_inputs(rng, nc, nt, names, 6)
        out, _ = rd.ref_single(cfgs, blocks, vols, calculator="LAMMPSPACE", ace=ace, numtypes=nt, types="In P", **kw)
        save_single(tag, cfgs, blocks, vols, out, nc, nt, bz, b2j, tm, kw, False)


def _single_inputs(rng, nc, nt, names, ncfg):
    cfgs, blocks, vols = [], [], []
    for i in range(ncfg):
        n = int(rng.integers(1, 11))
        cfgs.append(rd.make_config_dict(n, nt, rng, names, group="s%d" % (i % 2), fname="one%d" % i,
                                        eweight=float(10 ** rng.uniform(-2, 2)), fweight=float(10 ** rng.uniform(-2, 2)),
                                        vweight=float(10 ** rng.uniform(-9, -5))))
        blocks.append(rng.standard_normal((1 + 3 * n + 6, nc * nt + 1)) * 10.0 ** rng.uniform(-3, 3, (1, nc * nt + 1)))
        vols.append(float(rng.uniform(20, 4000)))
    return cfgs, blocks, vols


def save_single(tag, cfgs, blocks, vols, out, ncoeff, numtypes, bzeroflag, b2j, tm, kw, drop):
    natoms = np.array([c["NumAtoms"] for c in cfgs], dtype=np.int32)
    tf = np.zeros((len(cfgs), numtypes))
    for i, c in enumerate(cfgs):
        for at in c["AtomTypes"]:
            tf[i, tm[at] - 1] += 1
        tf[i] /= len(c["AtomTypes"])
    np.savez_compressed(
        os.path.join(OUT, "single_%s.npz" % tag),
        raw=np.concatenate(blocks, 0), natoms=natoms, volume=np.array(vols),
        energy=np.array([c["Energy"] for c in cfgs]),
        forces=np.concatenate([c["Forces"].reshape(-1) for c in cfgs]),
        stress=np.stack([c["Stress"] for c in cfgs]),
        atom_type_index=np.concatenate([[tm[a] for a in c["AtomTypes"]] for c in cfgs]).astype(np.int32),
        eweight=np.array([c["eweight"] for c in cfgs]), fweight=np.array([c["fweight"] for c in cfgs]),
        vweight=np.array([c["vweight"] for c in cfgs]), weights_dropped=bool(drop), type_fraction=tf, blank2j=b2j,
        ncoeff=ncoeff, numtypes=numtypes, bzeroflag=bzeroflag,
        use_energy=kw["energy"], use_force=kw["force"], use_stress=kw["stress"],
        rows_per_config=np.array([o[0].shape[0] for o in out], dtype=np.int64),
        ref_a=np.concatenate([o[0] for o in out], 0), ref_b=np.concatenate([o[1] for o in out]),
        ref_w=np.concatenate([o[2] for o in out]))
    print("single_%s" % tag, [o[0].shape for o in out][:3])


def group_weights():
    """(eweight, fweight, vweight) tables of the reference's WBe and InP examples ([GROUPS] of
    examples/WBe_PRB2019/WBe-example.in and examples/InP_JPCA2020/InP-example.in): the row-weight distributions of
    BASELINE configs[4] / configs[2], used by the -m gpu parity tests at those shapes."""
    import configparser
    import json
    out = {}
    for tag, rel in (("WBe", "examples/WBe_PRB2019/WBe-example.in"), ("InP", "examples/InP_JPCA2020/InP-example.in")):
        cp = configparser.ConfigParser(inline_comment_prefixes=("#",))
        cp.optionxform = str
        cp.read(os.path.join(rd.REFERENCE_ROOT, rel))
        rows = []
        for name, val in cp["GROUPS"].items():
            if name in ("group_sections", "group_types", "smartweights", "random_sampling", "BOLTZT"):
                continue
            f = val.split()
            rows.append([name, float(f[2]), float(f[3]), float(f[4])])
        out[tag] = rows
        print(tag, len(rows), "groups")
    with open(os.path.join(OUT, "group_weights.json"), "w") as f:
        json.dump(out, f, indent=0)


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    assert rd.reference_available(), "needs the reference tree at %s" % rd.REFERENCE_ROOT
    which = sys.argv[1:] or ["ta", "solve", "anl", "scatter", "single", "groups"]
    if "groups" in which:
        group_weights()
    if "ta" in which:
        ta_linear()
    if "solve" in which:
        solve_cases()
    if "anl" in which:
        anl_cases()
    if "scatter" in which:
        scatter_cases()
    if "single" in which:
        single_cases()


if __name__ == "__main__":
    main()
