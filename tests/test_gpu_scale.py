"""Full-size (BASELINE config 3: 10M sites) checks through size-independent
properties: shard invariance, sortedness, agreement with the exact direct-form
kernel on the same device, and the oracle on a random subsample."""
import numpy as np
import pytest

import denovo_b200 as D
import oracle_lib as O

pytestmark = pytest.mark.gpu
N = 10_000_000


def test_synthetic_source_matches_its_cpu_specification():
    import torch
    n = 50_000
    for depth, eps, seed in ((30, 1e-2, 0x5EED0003), (200, 1e-3, 0x5EED0005)):
        thr = D.poisson_thresholds(depth)
        assert np.array_equal(thr, O.poisson_thresholds(depth))
        with D.TrioEngine(eps) as eng:
            buf = torch.empty((n, 12), dtype=torch.int16, device="cuda")
            eng.synth_counts_device(seed, 1234567, n, buf, depth=depth)
            got = buf.cpu().numpy().view(np.uint16).reshape(n, 3, 4)
        want = O.synth_counts(seed, 1234567, n, depth=depth, seq_err=eps)
        assert np.array_equal(got, want)
        assert abs(got.sum(axis=2).mean() - depth) < 0.2


def _values_close(got, want, ok):
    """p / LR / log-likelihoods of called sites: <= 1e-9 relative where the oracle's sums are normal
    (SURVEY.md A.6 value-parity rule); identical special values elsewhere."""
    for f in ("p", "lr", "log_men", "log_den"):
        g, w = got[f][ok], want[f][ok]
        assert np.all(np.abs(g - w) <= 1e-9 * np.maximum(1.0, np.abs(w))), f
    for f in ("p", "lr", "log_men", "log_den"):
        g, w = got[f][~ok], want[f][~ok]
        assert np.array_equal(np.isnan(g), np.isnan(w)) and np.array_equal(g[~np.isnan(g)], w[~np.isnan(w)]), f


@pytest.mark.parametrize("method,lrt", [("BAYES", 1.0), ("MAP", 1.0), ("LRT", 10.0)])
def test_10m_sites_every_site_against_the_oracle(method, lrt):
    """BASELINE config 3 at FULL size: all 10 M sites are compared with the threaded CPU oracle — the set of
    called sites, and for every called site the MAP triple and max_ll bit for bit, p / LR / log-likelihoods
    within 1e-9 — plus the size-independent properties (input order, shard invariance, agreement with the
    exact direct-form kernel on a 1 M-site window)."""
    import os
    import torch
    with D.TrioEngine(1e-2, 1e-8, lrt, method) as eng, D.TrioEngine(1e-2, 1e-8, lrt, method, force_direct=True) as exact:
        counts = torch.empty((N, 12), dtype=torch.int16, device="cuda")
        eng.synth_counts_device(0x5EED0003, 0, N, counts, depth=30)
        cap = 200_000
        calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        n_calls = eng.infer_device(counts, N, calls, cap)
        st = eng.stats()
        c_all = calls[:n_calls].cpu().numpy().view(D.CALL_DT).reshape(-1)
        assert 0.0002 * N < n_calls < 0.02 * N
        assert np.all(np.diff(c_all["site"]) > 0)            # input order
        assert st["n_hard"] < 1e-4 * N
        # every site against the oracle (all host cores)
        h_counts = counts.cpu().numpy().view(np.uint16).reshape(N, 3, 4)
        ref = O.infer(O.params(1e-2, 1e-8, lrt), h_counts, threads=os.cpu_count() or 1)
        want = ref[{"BAYES": "denovo_bayes", "MAP": "denovo_map", "LRT": "denovo_lrt"}[method]].astype(bool)
        assert np.array_equal(c_all["site"], np.nonzero(want)[0]), "called sites differ from the oracle"
        r = ref[c_all["site"]]
        for f in ("dad", "mom", "child"):
            assert np.array_equal(c_all[f], r[f]), f
        assert np.array_equal(c_all["max_ll"], r["max_ll"])
        assert np.array_equal(c_all["counts"].reshape(-1, 3, 4), h_counts[c_all["site"]])
        _values_close(c_all, r, (r["l_den"] >= 2.3e-308) & (r["l_men"] >= 2.3e-308))
        del ref, h_counts
        # shard invariance: two halves = whole (what interval sharding across ranks relies on)
        half = N // 2
        n1 = eng.infer_device(counts[:half], half, calls, cap)
        a = calls[:n1].cpu().numpy().view(D.CALL_DT).reshape(-1).copy()
        n2 = eng.infer_device(counts[half:], N - half, calls, cap)
        b = calls[:n2].cpu().numpy().view(D.CALL_DT).reshape(-1).copy()
        b["site"] += half
        assert np.concatenate([a, b]).tobytes() == c_all.tobytes()
        # fast path == exact direct-form kernel on a 1M-site window
        lo, m = 3_000_000, 1_000_000
        n3 = exact.infer_device(counts[lo:lo + m], m, calls, cap)
        e = calls[:n3].cpu().numpy().view(D.CALL_DT).reshape(-1).copy()
        e["site"] += lo
        sel = c_all[(c_all["site"] >= lo) & (c_all["site"] < lo + m)]
        for f in ("site", "dad", "mom", "child", "max_ll", "counts"):
            assert np.array_equal(e[f], sel[f]), f
        assert np.allclose(e["p"], sel["p"], rtol=1e-9, atol=0)


def _run_gather_workers(world, same_device):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if same_device:
        env["DNV_TEST_SAME_DEVICE"] = "1"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29533" if same_device else "29534",
                        "tests/multi_gpu_check.py"], cwd=root, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("multi_gpu_check OK") == 3  # two gather rounds + the collective fallback


@pytest.mark.parametrize("world", [2, 3])
def test_interprocess_gather_equals_single_context(world):
    """One process per rank, every rank on cuda:0 (runs on a one-GPU box): the interval shards, the IPC-mapped
    destination buffer, the kernels' remote stores, the release flags and the destination's wait must reproduce
    single-context inference over the whole range record for record."""
    _run_gather_workers(world, same_device=True)


def test_multi_gpu_gather_equals_single_gpu():
    """The same over NVLink on real devices (skipped on a 1-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_gather_workers(min(torch.cuda.device_count(), 8), same_device=False)


def test_gather_ends_in_one_process_and_slot_layout():
    """dnv_gather inside one process (peer access; both ends on cuda:0): slots are disjoint, 256-byte aligned,
    generation-major, records 16-byte aligned; two ranks' shards gathered = the whole range."""
    import ctypes as C
    import torch
    total, cap, gens, world = 700_001, 8000, 4, 2
    with D.TrioEngine() as e0, D.TrioEngine() as e1:
        ends = [D.Gather(e0, world, 0, cap, gens), D.Gather(e1, world, 1, cap, gens)]
        ends[1].attach(ends[0])
        seen = {}
        for gen in range(2 * gens):
            for r in range(world):
                rec, cnt = C.c_void_p(), C.c_void_p()
                assert e0.lib.dnv_gather_slot(ends[0]._h, r, gen, C.byref(rec), C.byref(cnt)) == 0
                assert cnt.value % 256 == 0 and rec.value == cnt.value + 64 and rec.value % 16 == 0
                if gen < gens:
                    assert all(abs(cnt.value - o) >= 64 + cap * 80 for o in seen.values())
                    seen[(r, gen)] = cnt.value
                else:
                    assert cnt.value == seen[(r, gen - gens)]  # generations wrap
        whole = torch.empty((total, 12), dtype=torch.int16, device="cuda")
        e0.synth_counts_device(0x5EED0003, 0, total, whole)
        ref = torch.zeros((2 * cap, 80), dtype=torch.uint8, device="cuda")
        k = e0.infer_device(whole, total, ref, 2 * cap)
        want = ref[:k].cpu().numpy().tobytes()
        half = total // 2
        for gen in (0, 1, 5):
            ends[0].submit(whole[:half], half, 0, gen)
            ends[1].submit(whole[half:], total - half, half, gen)
        ends[1].arrive()
        ends[0].arrive()
        ends[0].wait()
        for gen in (0, 1, 5):
            got, per = ends[0].fetch(gen)
            assert got.tobytes() == want and per.sum() == k
        e1.sync()
        with pytest.raises(D.DenovoError) as ei:   # too small a host buffer: the needed total is reported
            ends[0].fetch(0, cap=3)
        assert ei.value.code == D.DNV_E_CAPACITY and str(k) in str(ei.value)
        for g in ends:
            g.close()


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0, 0, 0]])
def test_device_group_equals_single_context(devices):
    """dnv_group (one process, N devices; here N contexts on cuda:0): host counts are split by interval, one host
    thread per device, records gathered through device 0 — equal to one context over the whole batch."""
    n = 1_000_003
    counts = O.synth_counts(0x5EED0001, 0, n, depth=30, seq_err=1e-2)
    with D.TrioEngine() as eng:
        want, _ = eng.infer(counts)
    with D.DeviceGroup(devices) as grp:
        got = grp.infer(counts)
        assert got.tobytes() == want.tobytes()
        st = grp.stats()
        assert st["n_sites"] == n and st["n_calls"] == len(want)
        got8 = grp.infer(counts.astype(np.uint8))
        assert got8.tobytes() == want.tobytes()
        # fewer sites than devices, no sites, too small a capacity
        few = grp.infer(counts[:3])
        assert len(few) == int((want["site"] < 3).sum())
        assert len(grp.infer(counts[:0])) == 0
        with pytest.raises(D.DenovoError) as ei:
            grp.infer(counts, cap=5)
        assert ei.value.code == D.DNV_E_CAPACITY and str(len(want)) in str(ei.value)
        assert grp.infer(counts).tobytes() == want.tobytes()   # still usable
