"""GPU parity of the READ-mode hot path (through the C ABI) against the CPU oracle.

Bar: called sites, MAP genotype triples and isDenovo bit-exact; p, LR and the
log-likelihoods within 1e-9 relative wherever the oracle's raw sums are normal
doubles, identical special values (0, +-inf, NaN) elsewhere."""
import json
import math
import os

import numpy as np
import pytest

import denovo_b200 as D
import oracle_lib as O
import synth_inputs as S

pytestmark = pytest.mark.gpu
G = D.GENOTYPES
REL_TOL = 1e-9
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_junit_cases.json")))
METHOD_FIELD = {"MAP": "denovo_map", "BAYES": "denovo_bayes", "LRT": "denovo_lrt"}


def close(a, b, tol=REL_TOL, log_domain=False):
    """1e-9 relative.  For log-likelihoods the bar is on the likelihood itself
    (|d logL| <= 1e-9 is a 1e-9 relative error of L): a logL within 1e-8 of zero
    is the log of a sum within 1e-8 of one, where the reference's own 1000-term
    sum already carries ~1e-16 absolute rounding."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    same_special = (np.isnan(a) & np.isnan(b)) | (a == b)
    with np.errstate(invalid="ignore", divide="ignore"):
        scale = np.maximum(np.abs(a), np.abs(b))
        if log_domain:
            scale = np.maximum(scale, 1.0)
        rel = np.abs(a - b) <= tol * scale
    return same_special | rel


def check_against_oracle(counts, eps=1e-2, mu=1e-8, lrt=1.0, method="BAYES", force_direct=False, dense=True):
    counts = np.ascontiguousarray(counts, np.uint16).reshape(-1, 3, 4)
    ref = O.infer(O.params(eps, mu, lrt), counts)
    with D.TrioEngine(eps, mu, lrt, method, force_direct=force_direct) as eng:
        calls, d = eng.infer(counts, dense=dense)
        stats = eng.stats()
        # calls only: BAYES / LRT then take the headline path (integer screen + list kernel), and the
        # compact uint8 input the same path with 12-byte sites
        variants = [eng.infer(counts)[0]]
        if counts.size and counts.max() < 256:
            variants.append(eng.infer(counts.astype(np.uint8))[0])
    for other in variants:
        for f in ("site", "dad", "mom", "child", "max_ll", "counts", "is_denovo"):
            assert np.array_equal(other[f], calls[f]), ("calls-only path differs from the dense path", f)
        same = (other["flags"] == 0) & (calls["flags"] == 0)  # hard sites: sequential vs warp-parallel exact sums
        for f in ("p", "lr", "log_men", "log_den"):
            a, b = other[f], calls[f]
            assert np.array_equal(a[same], b[same], equal_nan=True), ("calls-only path differs from the dense path", f)
            assert np.all(close(a[~same], b[~same], log_domain=f.startswith("log")) | (np.isnan(a[~same]) & np.isnan(b[~same])))
    want = ref[METHOD_FIELD[method]].astype(bool)
    # ---- called sites: bit-exact set, input order
    assert np.array_equal(calls["site"], np.nonzero(want)[0])
    r = ref[want]
    assert np.array_equal(calls["dad"], r["dad"]) and np.array_equal(calls["mom"], r["mom"])
    assert np.array_equal(calls["child"], r["child"])
    assert np.all(calls["is_denovo"] == 1)
    assert np.array_equal(calls["counts"].reshape(-1, 3, 4), counts[want])
    assert np.array_equal(calls["max_ll"], r["max_ll"])  # pure add/mul chain: bit-exact
    normal = (r["l_den"] >= 2.3e-308) & (r["l_men"] >= 2.3e-308)
    for f in ("p", "lr", "log_men", "log_den"):
        assert np.all(close(calls[f][normal], r[f][normal], log_domain=f.startswith("log"))), f
    # ---- dense record for every site
    if dense:
        assert np.array_equal(d["is_denovo"].astype(bool), want)
        assert np.array_equal(d["dad"], ref["dad"]) and np.array_equal(d["mom"], ref["mom"])
        assert np.array_equal(d["child"], ref["child"])
        assert np.array_equal(d["max_ll"], ref["max_ll"])
        normal = (ref["l_den"] >= 2.3e-308) & (ref["l_men"] >= 2.3e-308)
        for f in ("p", "lr", "log_men", "log_den"):
            ok = close(d[f][normal], ref[f][normal], log_domain=f.startswith("log"))
            assert np.all(ok), (f, np.nonzero(~ok)[0][:5], d[f][normal][~ok][:5], ref[f][normal][~ok][:5])
        # special values where the reference's raw sums flushed to exactly zero
        z = ~normal & ((ref["l_den"] == 0) | (ref["l_men"] == 0)) & ((ref["l_den"] == 0) | (ref["l_den"] >= 2.3e-308)) \
            & ((ref["l_men"] == 0) | (ref["l_men"] >= 2.3e-308))
        for f in ("p", "lr"):
            assert np.all(close(d[f][z], ref[f][z])), f
    return ref, calls, d, stats


def fixture_counts():
    return np.array([[fx["DAD"], fx["MOM"], fx["CHILD"]] for fx in GOLD["fixtures"].values()], np.uint16)


def test_tables_bit_identical_to_oracle():
    t = O.tables(O.params())
    with D.TrioEngine() as eng:
        g = eng.tables()
    assert np.array_equal(g["R"], t["R"])
    assert np.array_equal(g["BL"], t["BL"])
    assert np.array_equal(g["isden"], t["isden"])


@pytest.mark.parametrize("force_direct", [False, True])
@pytest.mark.parametrize("case", GOLD["infer_cases"], ids=lambda c: c["source"])
def test_reference_junit_cases_through_c_abi(case, force_direct):
    """The reference's own 21 known answers, asserted directly on the GPU path."""
    fx = case["counts"]
    c = np.array([[fx["DAD"], fx["MOM"], fx["CHILD"]]], np.uint16)
    p = GOLD["params"]
    with D.TrioEngine(p["seq_err_rate"], p["denovo_mut_rate"], p["lrt_threshold"], case["method"],
                      force_direct=force_direct) as eng:
        calls, d = eng.infer(c, dense=True)
    assert [G[d["dad"][0]], G[d["mom"][0]], G[d["child"][0]]] == case["expect_map_triple"]
    assert bool(d["is_denovo"][0]) == case["expect_is_denovo"]
    assert len(calls) == int(case["expect_is_denovo"])


@pytest.mark.parametrize("method", ["MAP", "BAYES", "LRT"])
@pytest.mark.parametrize("force_direct", [False, True])
def test_config1_10k_sites_30x(method, force_direct):
    """BASELINE config 1: 10k synthetic sites, 30x, eps 1e-2, mu 1e-8."""
    counts = O.synth_counts(0x5EED0001, 0, 10000, depth=30, seq_err=1e-2)
    ref, calls, d, stats = check_against_oracle(counts, method=method, force_direct=force_direct)
    assert len(calls) > 5
    if not force_direct:
        assert stats["n_hard"] <= 10


def test_lrt_threshold_other_than_one():
    counts = O.synth_counts(0x5EED0003, 0, 20000)
    for thr in (10.0, 0.1, 1e6):
        check_against_oracle(counts, lrt=thr, method="LRT")


@pytest.mark.parametrize("mu", [1e-10, 1e-9, 1e-8, 1e-7, 1e-6])
def test_config5_high_depth_underflow_path(mu):
    """BASELINE config 5: 200x, eps 1e-3, mu sweep; ~4.5 % of sites have l_den == 0
    in the reference (raw exp underflow) — decisions and special values must match."""
    counts = O.synth_counts(0x5EED0005, 0, 20000, depth=200, seq_err=1e-3)
    ref, calls, d, stats = check_against_oracle(counts, eps=1e-3, mu=mu, method="BAYES")
    under = (ref["l_den"] < 2.3e-308) | (ref["l_men"] < 2.3e-308)
    assert under.mean() > 0.01
    zero = ref["l_den"] == 0
    assert np.all(d["p"][zero & (ref["l_men"] > 0)] == 0.0)
    assert np.all(np.isneginf(d["log_den"][zero]))
    check_against_oracle(counts[:4000], eps=1e-3, mu=mu, method="LRT", lrt=3.0)
    check_against_oracle(counts[:4000], eps=1e-3, mu=mu, method="MAP")


def test_fixture_pileups_all_methods_and_eps():
    c = fixture_counts()
    for eps in (1e-2, 1e-3, 0.2):
        for method in ("MAP", "BAYES", "LRT"):
            check_against_oracle(c, eps=eps, method=method)


def test_edge_cases():
    rng = np.random.default_rng(7)
    blocks = [np.zeros((3, 3, 4), np.uint16),                         # zero depth
              np.full((2, 3, 4), 65535, np.uint16),                   # maximum counts: everything underflows
              rng.integers(0, 3, (500, 3, 4)).astype(np.uint16),      # very shallow
              rng.integers(0, 2000, (500, 3, 4)).astype(np.uint16),   # deep, noisy: gray zone / table fallback
              rng.integers(0, 40, (2000, 3, 4)).astype(np.uint16)]    # noisy 4-allele pileups
    one = np.zeros((64, 3, 4), np.uint16)                             # single-allele deep columns
    for i in range(64):
        one[i, :, i % 4] = [1000 + i, 1200, 900 + 7 * i]
    c = np.concatenate(blocks + [one])
    for method in ("BAYES", "MAP", "LRT"):
        ref, calls, d, stats = check_against_oracle(c, method=method)
    z = np.zeros((1, 3, 4), np.uint16)
    ref, calls, d, _ = check_against_oracle(z, method="MAP")
    assert [G[d["dad"][0]], G[d["mom"][0]], G[d["child"][0]]] == ["AC", "AC", "AC"]


def test_threshold_knife_edge_sites_go_to_exact_path():
    """Sites built so that p is as close to 0.5 as integer counts allow must agree
    with the oracle; any site within 1e-9 of the threshold is decided by the
    exact direct-form kernel (n_hard counts them)."""
    base = np.array([[0, 58, 0, 2], [0, 51, 0, 2], [0, 28, 0, 8]], np.uint16)  # p = 0.52199 fixture
    var = []
    for a in range(0, 12):
        for b in range(20, 40):
            c = base.copy()
            c[2, 3] = a
            c[2, 1] = b
            var.append(c)
    ref, calls, d, stats = check_against_oracle(np.array(var), method="BAYES")
    assert np.abs(ref["p"] - 0.5).min() < 0.2


def test_empty_and_capacity():
    with D.TrioEngine() as eng:
        calls, d = eng.infer(np.zeros((0, 3, 4), np.uint16), dense=True)
        assert len(calls) == 0 and len(d) == 0
        counts = O.synth_counts(0x5EED0001, 0, 10000)
        full, _ = eng.infer(counts)
        assert len(full) > 3
        with pytest.raises(D.DenovoError) as ei:
            eng.infer(counts, cap=2)
        assert ei.value.code == D.DNV_E_CAPACITY
        assert str(len(full)) in str(ei.value)  # required size is reported
        again, _ = eng.infer(counts)            # ctx still usable, idempotent
        assert again.tobytes() == full.tobytes()


def test_host_path_chunking_matches_single_chunk():
    """The host API pipelines 2M-site chunks; results must not depend on chunking."""
    counts = O.synth_counts(0x5EED0003, 0, 5000)
    big = np.tile(counts, (1000, 1, 1))  # 5M sites > 2 chunks
    with D.TrioEngine() as eng:
        small, _ = eng.infer(counts)
        calls, _ = eng.infer(big)
    assert len(calls) == 1000 * len(small)
    assert np.array_equal(calls["site"] % 5000, np.tile(small["site"], 1000))
    assert np.all(np.diff(calls["site"]) > 0)
    assert np.array_equal(calls["p"], np.tile(small["p"], 1000))


def test_uint8_compact_input_matches_uint16():
    """dnv_infer_counts8 (12 B/site) must give byte-identical results to the uint16 entry point."""
    counts = O.synth_counts(0x5EED0003, 0, 300_000)
    assert counts.max() < 256
    for method in ("BAYES", "MAP"):
        with D.TrioEngine(method=method) as eng:
            a, da = eng.infer(counts, dense=True)
            b, db = eng.infer(counts.astype(np.uint8), dense=True)
        assert a.tobytes() == b.tobytes() and da.tobytes() == db.tobytes() and len(a) > 100
    with D.TrioEngine(force_direct=True) as eng:
        c, _ = eng.infer(counts[:20000].astype(np.uint8))
        d, _ = eng.infer(counts[:20000])
    assert c.tobytes() == d.tobytes()


@pytest.mark.parametrize("method", ["BAYES", "MAP"])
def test_async_batches_match_synchronous_calls(method):
    """dnv_infer_counts_device_async: several batches in flight on one ctx, each with its own
    outputs, must give byte-identical records and counts to the synchronous entry point."""
    import torch
    sizes = [70_001, 0, 5, 200_000, 33]
    with D.TrioEngine(method=method) as eng:
        bufs, want = [], []
        for i, n in enumerate(sizes):
            c = torch.empty((max(n, 1), 12), dtype=torch.int16, device="cuda")
            if n:
                eng.synth_counts_device(0x5EED0003, 1000 * i, n, c, depth=30)
            bufs.append(c)
            calls = torch.zeros((max(n, 1), D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
            k = eng.infer_device(c, n, calls, max(n, 1))
            want.append(calls[:k].cpu().numpy().tobytes())
        h_nc = torch.full((len(sizes),), -1, dtype=torch.int64).pin_memory()
        outs = [torch.zeros((max(n, 1), D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda") for n in sizes]
        for i, n in enumerate(sizes):
            eng.infer_device_async(bufs[i], n, outs[i], max(n, 1), h_nc.data_ptr() + 8 * i)
        eng.sync()
        st = eng.stats()
        got_counts = h_nc.tolist()
        for i, n in enumerate(sizes):
            assert outs[i][:got_counts[i]].cpu().numpy().tobytes() == want[i], (i, n)
        assert st["n_sites"] == sum(sizes) and st["n_calls"] == sum(got_counts) and sum(got_counts) > 100
        # the ctx goes back to synchronous use without an explicit sync in between
        eng.infer_device_async(bufs[3], sizes[3], outs[3], sizes[3], h_nc.data_ptr())
        k = eng.infer_device(bufs[0], sizes[0], outs[0], sizes[0])
        assert outs[0][:k].cpu().numpy().tobytes() == want[0]


@pytest.mark.parametrize("eps,mu,lrt", [(1e-2, 1e-8, 1.0), (1e-3, 1e-10, 1.0), (1e-1, 1e-6, 0.05), (0.3, 1e-5, 20.0),
                                        (1e-2, 1e-3, 1.0), (2e-2, 1.7e-4, 1.0), (2e-2, 4.5e-5, 1.0)])
def test_screen_boundaries_against_oracle(eps, mu, lrt):
    """BAYES and LRT on adversarial pileups, for parameters where the integer screen is valid (small mu),
    barely valid (mu = 4.5e-5, just under its limit) and invalid (mu >= 1.7e-4: every site takes the FP64 path)."""
    rng = np.random.default_rng(int(eps * 1e6) + int(-np.log10(mu)))
    c = S.adversarial_counts(rng, 6000)
    for method in ("BAYES", "LRT"):
        check_against_oracle(c, eps=eps, mu=mu, lrt=lrt, method=method)


def test_multi_chunk_device_path_matches_single_chunk():
    """A device call walks its input in chunks of 2^27 sites; DNV_DEVICE_CHUNK shrinks the chunk so a
    small input crosses several boundaries (ragged last chunk, running call total, site offsets)."""
    import subprocess, sys, os
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import denovo_b200 as D
n = 300_001
out = []
for method in ("BAYES", "MAP"):
    with D.TrioEngine(method=method) as eng:
        c = torch.empty((n, 12), dtype=torch.int16, device="cuda")
        eng.synth_counts_device(0x5EED0003, 77, n, c, depth=30)
        calls = torch.zeros((n, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        for _ in range(3):  # the third call replays the graph captured by the second
            k = eng.infer_device(c, n, calls, n)
        st = eng.stats()
        out.append(calls[:k].cpu().numpy().tobytes())
        out.append(np.array([k, st["n_calls"], st["n_screened_out"]], np.int64).tobytes())
sys.stdout.buffer.write(b"".join(out))
""" % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "denovo-variant-caller-java_b200", "python")
    res = []
    for chunk in (None, "65536", "1280", "lanes"):
        env = dict(os.environ)
        if chunk == "lanes":  # a synchronous call split over the context's lanes (off by default: the whole call in one chunk is faster)
            env["DNV_SYNC_MIN_CHUNK"] = "65536"
        elif chunk:
            env["DNV_DEVICE_CHUNK"] = chunk
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        res.append(r.stdout)
    assert len(res[0]) > 1000 and res[0] == res[1] == res[2] == res[3]


def test_async_batches_over_the_lanes_match_synchronous_calls():
    """A stream of asynchronous batches (dnv_infer_counts_device_async: fed to the context's lanes in turn, so
    consecutive batches overlap on the GPU) must return, batch for batch, exactly what synchronous calls return —
    the first round eagerly, later rounds the same (graph replay of repeated arguments included)."""
    import torch
    n, nb = 150_000, 7
    with D.TrioEngine() as eng, D.TrioEngine() as eng2:
        bufs = []
        for i in range(nb):
            c = torch.empty((n, 12), dtype=torch.int16, device="cuda")
            eng.synth_counts_device(0x5EED0003, i * n, n, c, depth=30)
            bufs.append(c)
        torch.cuda.synchronize()
        want = []
        for c in bufs:
            calls = torch.zeros((n, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
            for _ in range(3):  # the third call replays the captured graph of the first two
                k = eng.infer_device(c, n, calls, n)
            want.append(calls[:k].cpu().numpy().tobytes())
        outs = [torch.zeros((n, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda") for _ in range(nb)]
        h_nc = torch.full((nb,), -1, dtype=torch.int64).pin_memory()
        for rnd in range(3):
            for o in outs:
                o.zero_()
            h_nc.fill_(-1)
            torch.cuda.synchronize()
            for i in range(nb):
                eng2.infer_device_async(bufs[i], n, outs[i], n, h_nc.data_ptr() + 8 * i)
            eng2.sync()
            st = eng2.stats()
            got = h_nc.tolist()
            assert st["n_sites"] == n * nb and st["n_calls"] == sum(got)
            for i in range(nb):
                assert outs[i][:got[i]].cpu().numpy().tobytes() == want[i], (rnd, i)
        # site_base shifts the site indices of the records, nothing else
        eng2.infer_device_async(bufs[2], n, outs[0], n, h_nc.data_ptr(), site_base=10**12)
        eng2.sync()
        a = outs[0][:h_nc[0]].cpu().numpy().view(D.CALL_DT).reshape(-1).copy()
        a["site"] -= 10**12
        assert a.tobytes() == want[2]


def test_device_entry_points_reject_bad_arguments():
    """Misaligned device pointers, a missing count slot and too many batches in flight fail with
    DNV_E_INVALID and a message; the context stays usable."""
    import ctypes as C
    import torch
    from denovo_b200 import _lib as L
    n = 1000
    with D.TrioEngine() as eng:
        c = torch.zeros((n + 1, 12), dtype=torch.int16, device="cuda")
        calls = torch.zeros((n + 1, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        h_nc = torch.zeros(1, dtype=torch.int64).pin_memory()
        nc = C.c_int64(0)
        lib = eng.lib
        assert lib.dnv_infer_counts_device(eng._h, C.c_void_p(c.data_ptr() + 2), n, None, C.c_void_p(calls.data_ptr()), n,
                                           C.byref(nc)) == D.DNV_E_INVALID
        assert b"aligned" in lib.dnv_last_error()
        assert lib.dnv_infer_counts_device(eng._h, C.c_void_p(c.data_ptr()), n, None, C.c_void_p(calls.data_ptr() + 8), n,
                                           C.byref(nc)) == D.DNV_E_INVALID
        assert b"16-byte" in lib.dnv_last_error()
        assert lib.dnv_infer_counts_device_async(eng._h, C.c_void_p(c.data_ptr()), n, C.c_void_p(calls.data_ptr()), n,
                                                 None) == D.DNV_E_INVALID
        for i in range(200):  # no limit on the number of batches between two dnv_sync calls
            eng.infer_device_async(c, n, calls, n, h_nc)
        eng.sync()
        assert eng.stats()["n_sites"] == 200 * n and eng.stats()["n_calls"] == 0
        assert eng.infer_device(c, n, calls, n) == 0  # all-zero pileups: nothing is called
