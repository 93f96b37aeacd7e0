"""Shared helpers for the parity tests: run the oracle, feed the same state to the C-ABI."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refdump  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def cfg_for(rec_meta, extra=None):
    """cfg dict for mdacp_b200.driver.MDManager that reproduces the dump's SimulationInfo."""
    m = rec_meta
    cfg = {
        "SystemSizeX": repr(m["L"][0]), "SystemSizeY": repr(m["L"][1]), "SystemSizeZ": repr(m["L"][2]),
        "TimeStep": repr(m["dt"]), "AimedTemperature": repr(m["T0"]), "HeatbathGamma": repr(m["gamma"]),
        "HeatbathTau": repr(m["tau"]), "ControlTemperature": "yes" if m["control_temperature"] else "no",
        "HeatbathType": "Langevin" if m["heatbath_type"] == 1 else "NoseHoover",
        "CheckListLength": str(m["check_list_length"]),
        "LocalGridX": str(m["grid"][0]), "LocalGridY": str(m["grid"][1]), "LocalGridZ": str(m["grid"][2]),
    }
    if extra:
        cfg.update(extra)
    return cfg


def states_from_dump(rec, step, nunits):
    out = []
    for u in range(nunits):
        q = rec["q.s%d.u%d" % (step, u)]
        p = rec["p.s%d.u%d" % (step, u)]
        t = rec["type.s%d.u%d" % (step, u)]
        out.append({"q": q[:p.shape[0]], "p": p, "type": t})
    return out


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor) if floor else np.abs(b)
    return float(np.max(np.abs(a - b) / np.where(den == 0, 1.0, den))) if a.size else 0.0


def force_err(f, fref):
    """|dF| <= tol * max(|F_i|, F_rms): per-particle error normalised as SURVEY.md section 8a defines."""
    f = np.asarray(f)
    fref = np.asarray(fref)
    mag = np.sqrt((fref ** 2).sum(axis=1))
    rms = float(np.sqrt((mag ** 2).mean())) if mag.size else 1.0
    den = np.maximum(mag, rms)
    err = np.sqrt(((f - fref) ** 2).sum(axis=1)) / den
    return float(err.max()) if err.size else 0.0


# ---------------------------------------------------------------------------------------------
# record replay: the same reference records (live dump or committed golden .npz) check
# (a) the C oracle bit-for-bit and (b) the CUDA path within the stated tolerances
# ---------------------------------------------------------------------------------------------
import hashlib  # noqa: E402

TOL_OBS = 1e-9     # K, U, W, T, P, E: relative
TOL_FORCE = 1e-10  # per-particle impulse, relative to max(|F_i|, F_rms)
TOL_STATE = 1e-10  # q relative to box length, p relative to max(1, |p|max)


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    cfg = dict(zip([str(x) for x in rec["cfg_keys"]], [str(x) for x in rec["cfg_vals"]]))
    return rec, cfg, int(rec["steps"][0]), bool(rec["init_build"][0])


def pairs_match_set(rec, s, u, key, partner):
    """bit-exact equality of the pair SET with the reference's, from full arrays or from hashes"""
    tag = "s%d.u%d" % (s, u)
    mine = refdump.pair_set(key, partner)
    if ("pair_key." + tag) in rec:
        ref = refdump.pair_set(rec["pair_key." + tag], rec["pair_partner." + tag])
        return mine.shape == ref.shape and np.array_equal(mine, ref)
    if ("pairhash." + tag) in rec:
        return (len(key) == int(rec["paircount." + tag][0])
                and np.array_equal(_sha(mine.astype(np.int32)), rec["pairhash." + tag]))
    raise KeyError("no pair record for " + tag)


def has_pairs(rec, s, u):
    tag = "s%d.u%d" % (s, u)
    return ("pair_key." + tag) in rec or ("pairhash." + tag) in rec


def check_oracle_against_records(rec, steps, init_build, seed=None):
    """Replays reference records on the C oracle and asserts BIT equality everywhere."""
    from oracle import pyoracle
    meta = refdump.meta_of(rec)
    nu = meta["nunits"]
    cap = int(max(rec["tn"].max() * 2, 20000))
    osys = pyoracle.System.from_meta(meta, capacity=cap)
    try:
        for u, st in enumerate(states_from_dump(rec, 0, nu)):
            osys.set_unit(u, st["q"], st["p"], st["type"])
        if seed is not None:
            osys.set_seed(seed)
        if init_build:
            osys.make_pair_list()
        last = [0] * nu
        for s in range(steps + 1):
            if s > 0:
                osys.calculate(1)
            for u in range(nu):
                tag = "s%d.u%d" % (s, u)
                c = osys.counts(u)
                assert c["nbuilds"] == rec["builds"][s, u]
                assert c["pn"] == rec["pn"][s, u] and c["tn"] == rec["tn"][s, u]
                o = osys.observe(u)
                assert np.array_equal(o[:3], rec["obs_KUW"][s, u]), (s, u, o[:3], rec["obs_KUW"][s, u])
                if c["nbuilds"] != last[u]:
                    last[u] = c["nbuilds"]
                    k, p = osys.pairs(u)
                    if ("pair_key." + tag) in rec:  # same ORDER, not only the same set
                        assert np.array_equal(k, rec["pair_key." + tag])
                        assert np.array_equal(p, rec["pair_partner." + tag])
                    elif ("pairorder." + tag) in rec:
                        assert len(k) == int(rec["paircount." + tag][0])
                        assert np.array_equal(_sha(np.stack([k, p], axis=1).astype(np.int32)), rec["pairorder." + tag])
                if ("csr_list." + tag) in rec:
                    ptr, cnt, lst = osys.csr(u)
                    assert np.array_equal(ptr, rec["csr_pointer." + tag])
                    assert np.array_equal(cnt, rec["csr_count." + tag])
                    assert np.array_equal(lst, rec["csr_list." + tag])
                if ("impulse." + tag) in rec:
                    assert np.array_equal(osys.impulse(u), rec["impulse." + tag])
                if ("q." + tag) in rec and s > 0:
                    st = osys.state(u)
                    assert np.array_equal(st["q"], rec["q." + tag])
                    assert np.array_equal(st["p"], rec["p." + tag])
                    assert np.array_equal(st["type"], rec["type." + tag])
                    assert osys.zeta(u) == rec["zeta." + tag][0]
            if "change_scale" in rec and s == int(rec["change_scale"][0]):
                osys.change_scale(float(rec["change_scale"][1]))  # MDManager::ChangeScale after step s
    finally:
        osys.close()


def check_cuda_against_records(rec, steps, init_build, seed=None, observe_every=1):
    """Replays reference records through the C-ABI on the GPU.
    pair sets / rebuild schedule / pn / tn / reference index order: exact;
    impulse <= TOL_FORCE; K,U,W,T,P,E <= TOL_OBS; final q,p <= TOL_STATE."""
    from mdacp_b200 import driver
    meta = refdump.meta_of(rec)
    nunits = meta["nunits"]
    mdm = driver.MDManager(cfg_for(meta), num_procs=1, rank=0, num_threads=nunits, cutoff=meta["cutoff"])
    worst = {"K": 0.0, "U": 0.0, "W": 0.0, "F": 0.0}
    try:
        assert tuple(mdm.pinfo.GetGridSize()) == meta["grid"]
        mdm.upload(states_from_dump(rec, 0, nunits))
        if seed is not None:
            mdm.SetSeed(seed)
        builds_ref, pn_ref, tn_ref, obs_ref = rec["builds"], rec["pn"], rec["tn"], rec["obs_KUW"]
        if init_build:
            mdm.MakePairList()
        last_builds = [0] * nunits
        for s in range(steps + 1):
            if s > 0:
                mdm.Calculate()
            for u in range(nunits):
                unit = mdm.GetMDUnit(u)
                c = unit.counts()
                assert c["nbuilds"] == builds_ref[s, u], "rebuild schedule differs at step %d unit %d" % (s, u)
                assert c["pn"] == pn_ref[s, u] and c["tn"] == tn_ref[s, u], "pn/tn differ at step %d unit %d" % (s, u)
                rebuilt = c["nbuilds"] != last_builds[u]
                last_builds[u] = c["nbuilds"]
                # at every rebuild the record holds, and at any other step it holds a pair record for (the big
                # live cases dump the list of the last step only: DumpPairsAt, not DumpPairsAtBuild)
                if has_pairs(rec, s, u) and c["nbuilds"] > 0:
                    k, p = unit.pairs()
                    assert pairs_match_set(rec, s, u, k, p), "pair set differs at step %d unit %d" % (s, u)
                if ("impulse.s%d.u%d" % (s, u)) in rec:
                    e = force_err(unit.impulse(), rec["impulse.s%d.u%d" % (s, u)])
                    worst["F"] = max(worst["F"], e)
                    assert e <= TOL_FORCE, "impulse error %.3e at step %d unit %d" % (e, s, u)
                if s % observe_every == 0 or s == steps:
                    o = unit.observe()
                    r = obs_ref[s, u]
                    for idx, name in enumerate(("K", "U", "W")):
                        if c["nbuilds"] == 0 and idx > 0:
                            continue
                        den = abs(r[idx]) if r[idx] != 0 else 1.0
                        if name == "W":
                            # the virial is a sum with cancellation (near-zero pressure droplets): its error
                            # is measured on the scale it enters the pressure with, P = (2K/3 - W) / V
                            den = max(den, abs(r[0]))
                        worst[name] = max(worst[name], abs(o[idx] - r[idx]) / den)
                        assert worst[name] <= TOL_OBS, "%s differs by %.3e (relative) at step %d" % (
                            name, worst[name], s)
            if "change_scale" in rec and s == int(rec["change_scale"][0]):
                mdm.ChangeScale(float(rec["change_scale"][1]))  # MDManager::ChangeScale after step s
        # printed observables, combined exactly as mdmanager.cc:390-429
        tot = obs_ref[steps].sum(axis=0)
        N = float(pn_ref[steps].sum())
        Lref = list(meta["L"])
        if "change_scale" in rec and steps > int(rec["change_scale"][0]):
            Lref = [x * float(rec["change_scale"][1]) for x in Lref]
        ref_o = driver.observables_from_raw([tot[0], tot[1], tot[2], N], Lref)
        my_o = mdm.observables()
        for k in ("T", "E"):
            assert abs(my_o[k] - ref_o[k]) <= TOL_OBS * abs(ref_o[k]), (k, my_o[k], ref_o[k])
        pscale = max(abs(ref_o["P"]), ref_o["T"] * N / (Lref[0] * Lref[1] * Lref[2]))
        assert abs(my_o["P"] - ref_o["P"]) <= TOL_OBS * pscale, ("P", my_o["P"], ref_o["P"])
        for u in range(nunits):
            tag = "s%d.u%d" % (steps, u)
            if ("q." + tag) not in rec:
                continue
            d = mdm.GetMDUnit(u).download()
            qr, pr, tr = rec["q." + tag], rec["p." + tag], rec["type." + tag]
            assert np.array_equal(d["type"], tr), "reference index order differs (unit %d)" % u
            assert d["q"].shape == qr.shape
            assert np.max(np.abs(d["q"] - qr)) <= TOL_STATE * max(Lref), "q differs (unit %d)" % u
            assert np.max(np.abs(d["p"] - pr)) <= TOL_STATE * max(1.0, float(np.abs(pr).max())), "p differs"
        return worst
    finally:
        mdm.close()
