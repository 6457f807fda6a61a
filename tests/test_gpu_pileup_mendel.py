"""GPU parity (bit-exact) of the pileup kernel and the VARIANT-mode Mendelian filter
against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import denovo_b200 as D
import oracle_lib as O
import synth_inputs as S

pytestmark = pytest.mark.gpu


def pcalls_equal(a, b):
    return all(np.array_equal(a[f], b[f]) for f in ("position", "child_index", "child", "mom", "dad", "is_denovo"))


@pytest.mark.parametrize("var_len", [False, True])
def test_pileup_dense_sites(var_len):
    rs, bo, ba, pe = S.make_reads(11, 5000, 3000, depth=30, read_len=100, var_len=var_len)
    pos = np.arange(4900, 8200, dtype=np.int64)  # includes uncovered flanks (zero depth)
    rc, want = O.pileup(rs, bo, ba, pe, pos)
    assert rc == 0
    with D.TrioEngine() as eng:
        got = eng.pileup(rs, bo, ba, pe, pos)
    assert np.array_equal(got, want)
    assert got[0].sum() == 0 and got[-50:].sum() == 0 and 15 < got[500:2500].sum(axis=2).mean() < 35


def test_pileup_device_arrays_that_are_only_8_byte_aligned():
    """The pipeline kernel bulk-copies slices of read_start / base_off and wants them 16-byte aligned; arrays that are
    only 8-byte aligned (a view one element into a buffer) take the gather kernel: same counts."""
    import torch
    rs, bo, ba, pe = S.make_reads(31, 5000, 4000, depth=30)
    pos = np.arange(4900, 9100, dtype=np.int64)
    rc, want = O.pileup(rs, bo, ba, pe, pos)
    assert rc == 0
    with D.TrioEngine() as eng:
        pad = lambda a: torch.from_numpy(np.concatenate([a[:1], a])).cuda()[1:]
        d_rs, d_bo = pad(rs), pad(bo)
        assert d_rs.data_ptr() % 16 == 8 and d_bo.data_ptr() % 16 == 8
        d_ba, d_pe, d_pos = (torch.from_numpy(x).cuda() for x in (ba, pe, pos))
        out = torch.zeros((len(pos), 12), dtype=torch.int16, device="cuda")
        eng.pileup_device(d_rs, d_bo, d_ba, d_pe, len(rs), d_pos, len(pos), out)
        got = out.cpu().numpy().view(np.uint16).reshape(len(pos), 3, 4)
    assert np.array_equal(got, want)


def test_pileup_sparse_sites_and_edges():
    rs, bo, ba, pe = S.make_reads(12, 100000, 20000, depth=200, read_len=150, gap_rate=0.01)
    rng = np.random.default_rng(1)
    pos = np.unique(rng.integers(100000, 120000, 300)).astype(np.int64)
    rc, want = O.pileup(rs, bo, ba, pe, pos)
    with D.TrioEngine() as eng:
        assert np.array_equal(eng.pileup(rs, bo, ba, pe, pos), want)
        # no reads at all / one person missing
        z = eng.pileup(np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.uint8), np.zeros(0, np.uint8), pos)
        assert z.sum() == 0
        keep = pe != 1
        bo2 = np.zeros(keep.sum() + 1, np.int64)
        lens = np.diff(bo)[keep]
        np.cumsum(lens, out=bo2[1:])
        ba2 = np.concatenate([ba[bo[i]:bo[i + 1]] for i in np.nonzero(keep)[0]])
        got = eng.pileup(rs[keep], bo2, ba2, pe[keep], pos)
        assert np.array_equal(got[:, [0, 2]], want[:, [0, 2]]) and got[:, 1].sum() == 0
        # a base outside ACGT- is the reference's IllegalArgumentException
        bad = ba.copy()
        r = int(np.searchsorted(rs[pe == 0], pos[5], side="right")) - 1
        bad[bo[r] + (pos[5] - rs[r])] = ord("N")
        with pytest.raises(D.DenovoError) as ei:
            eng.pileup(rs, bo, bad, pe, pos)
        assert ei.value.code == D.DNV_E_BADBASE
        with pytest.raises(D.DenovoError) as ei:
            eng.pileup(rs[::-1].copy(), bo, ba, pe, pos)
        assert ei.value.code == D.DNV_E_UNSORTED


@pytest.mark.parametrize("n_sites", [1003, 1021, 8 * 128 + 1])
def test_pileup_dense_bad_base_only_at_requested_sites(n_sites):
    """Dense (staged) path, n_sites % 8 != 0: a non-ACGT base within the 7 positions AFTER the last site is
    not at a candidate position — the reference never looks at it (M/ReadSummary.java:48-55) — while the
    same base AT the last site is its IllegalArgumentException."""
    rs, bo, ba, pe = S.make_reads(14, 1000, 3000, depth=30)
    pos = np.arange(1000, 1000 + n_sites, dtype=np.int64)
    last = int(pos[-1])
    lens = np.diff(bo)
    near = ba.copy()
    for r in np.nonzero((rs <= last + 7) & (rs + lens > last + 1))[0]:   # reads reaching past the last site
        for q in range(last + 1, last + 8):
            off = q - rs[r]
            if 0 <= off < lens[r]:
                near[bo[r] + off] = ord("N")
    assert (near != ba).sum() > 50
    rc, want = O.pileup(rs, bo, near, pe, pos)
    assert rc == 0
    with D.TrioEngine() as eng:
        assert np.array_equal(eng.pileup(rs, bo, near, pe, pos), want)
        at = near.copy()
        r = int(np.nonzero((rs <= last) & (rs + lens > last))[0][0])
        at[bo[r] + (last - rs[r])] = ord("N")
        assert O.pileup(rs, bo, at, pe, pos)[0] < 0
        with pytest.raises(D.DenovoError) as ei:
            eng.pileup(rs, bo, at, pe, pos)
        assert ei.value.code == D.DNV_E_BADBASE


@pytest.mark.parametrize("kernel", ["pipe", "scatter", "stage"])
def test_all_pileup_kernels_dense_and_ragged(kernel, monkeypatch):
    """The producer / consumer pipeline (default), the read-major scatter kernel (DNV_PILEUP_KERNEL=scatter) and the
    position-major gather kernel (=stage) are all bit-exact against the oracle: dense sites, variable-length reads,
    200x depth (person-tiles cut into several chunks of reads / several rounds of 255 reads), gaps, an uncovered flank,
    a site count that is not a multiple of the tile, reads that end with the arrays."""
    import subprocess, sys, os
    code = r"""
import sys, numpy as np
sys.path.insert(0, "tests")
import denovo_b200 as D, oracle_lib as O, synth_inputs as S
for seed, depth, rl, var, first in ((21, 30, 100, False, 4800), (22, 200, 150, True, 4800), (23, 8, 37, True, 4800),
                                    (24, 30, 100, False, 5100), (25, 31, 99, True, 5101), (26, 700, 100, False, 5500)):
    rs, bo, ba, pe = S.make_reads(seed, 5000, 2500, depth=depth, read_len=rl, gap_rate=5e-3, var_len=var)
    pos = np.arange(first, 7777, dtype=np.int64)
    rc, want = O.pileup_sorted(rs, bo, ba, pe, pos)
    assert rc == 0
    with D.TrioEngine() as eng:
        got = eng.pileup(rs, bo, ba, pe, pos)
    assert np.array_equal(got, want), (seed, int((got != want).sum()))
print("ok")
"""
    env = dict(os.environ, DNV_PILEUP_KERNEL=kernel, PYTHONPATH=os.pathsep.join(sys.path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


@pytest.mark.parametrize("method,fused", [("BAYES", "0"), ("BAYES", "1"), ("MAP", "1"), ("LRT", "0"), ("LRT", "1")])
def test_reads_to_calls_equals_pileup_then_inference(method, fused, monkeypatch):
    """dnv_infer_reads — by default the pileup pipeline chained to the READ pipeline on the device, with
    DNV_READS_FUSED=1 one kernel (pileup in shared memory, screened there, counts only for undecided sites) — returns
    the records of dnv_infer_counts over dnv_pileup's counts, and those are the oracle's: dense every-position layout
    with planted de novo / inherited variants, a ragged tail, then a sparse candidate list (per-site path)."""
    monkeypatch.setenv("DNV_READS_FUSED", fused)
    n_sites = 300_000 + 77
    rs, bo, ba, pe = S.make_reads(17, 1, n_sites, depth=30, err=1e-2)
    # plant variants every 211 sites (vectorised): at even plants the child alone carries an alternative allele in
    # about half of its reads (the de novo pattern), at odd ones one parent carries it too (inherited)
    rng = np.random.default_rng(3)
    lens = np.diff(bo)
    posn = np.repeat(rs, lens) + (np.arange(bo[-1]) - np.repeat(bo[:-1], lens))
    per = np.repeat(pe, lens)
    k = (posn - 500) // 211
    planted = (posn >= 500) & ((posn - 500) % 211 == 0) & (ba != ord("-")) & (rng.random(len(posn)) < 0.5)
    carrier = (per == 2) | ((k % 2 == 1) & (per == (k // 2) % 2))
    sel = np.nonzero(planted & carrier)[0]
    alt = np.frombuffer(b"ACGT", np.uint8)[(k[sel] + 1) % 4]
    alt = np.where(alt == ba[sel], np.frombuffer(b"ACGT", np.uint8)[(k[sel] + 2) % 4], alt)
    ba[sel] = alt
    pos = np.arange(1, n_sites + 1, dtype=np.int64)
    lrt = 10.0 if method == "LRT" else 1.0
    with D.TrioEngine(1e-2, 1e-8, lrt, method) as eng:
        counts = eng.pileup(rs, bo, ba, pe, pos)
        want, _ = eng.infer(counts)
        got = eng.infer_reads(rs, bo, ba, pe, pos)
        assert len(want) > (20 if method == "LRT" else 100) and got.tobytes() == want.tobytes()
        st = eng.stats()
        assert st["n_sites"] == n_sites and st["n_calls"] == len(want)
        # sparse candidates (not consecutive positions): the per-site path of the fused kernel
        cand = np.unique(np.concatenate([want["site"][::3] + 1, rng.integers(1, n_sites, 5000)])).astype(np.int64)
        c2 = eng.pileup(rs, bo, ba, pe, cand)
        w2, _ = eng.infer(c2)
        g2 = eng.infer_reads(rs, bo, ba, pe, cand)
        assert len(w2) > 10 and g2.tobytes() == w2.tobytes()
        with pytest.raises(D.DenovoError) as ei:
            eng.infer_reads(rs, bo, ba, pe, pos, cap=7)
        assert ei.value.code == D.DNV_E_CAPACITY and str(len(want)) in str(ei.value)
        bad = ba.copy()
        bad[bo[len(rs) // 2] + 3] = ord("N")
        with pytest.raises(D.DenovoError) as ei:
            eng.infer_reads(rs, bo, bad, pe, pos)
        assert ei.value.code == D.DNV_E_BADBASE
    ref = O.infer(O.params(1e-2, 1e-8, lrt), O.pileup_sorted(rs, bo, ba, pe, pos)[1])
    field = {"BAYES": "denovo_bayes", "MAP": "denovo_map", "LRT": "denovo_lrt"}[method]
    assert np.array_equal(want["site"], np.nonzero(ref[field])[0])


def test_pileup_then_inference_end_to_end():
    """reads -> counts -> calls on the device equals oracle pileup -> oracle inference."""
    rs, bo, ba, pe = S.make_reads(13, 1, 20000, depth=30, err=2e-2)
    pos = np.arange(1, 20001, dtype=np.int64)
    rc, counts = O.pileup(rs, bo, ba, pe, pos)
    ref = O.infer(O.params(), counts)
    with D.TrioEngine() as eng:
        got_counts = eng.pileup(rs, bo, ba, pe, pos)
        calls, _ = eng.infer(got_counts)
    assert np.array_equal(got_counts, counts)
    assert np.array_equal(calls["site"], np.nonzero(ref["denovo_bayes"])[0])


@pytest.mark.parametrize("seed,n,adv", [(1, 200, False), (2, 200, True), (3, 5000, False), (4, 5000, True),
                                        (5, 60000, False)])
def test_mendel_filter_matches_streaming_oracle(seed, n, adv):
    calls = S.make_callset(seed, n, adversarial=adv)
    n_ref, want = O.mendel_stream(calls)
    assert n_ref >= 0
    with D.TrioEngine() as eng:
        got = eng.mendel_filter(calls)
    assert len(got) == n_ref and pcalls_equal(got, want)
    if n >= 5000:
        assert got["is_denovo"].sum() > 0


def test_mendel_filter_edges():
    with D.TrioEngine() as eng:
        assert len(eng.mendel_filter(np.zeros(0, D.VCALL_DT))) == 0
        calls = S.make_callset(7, 3000)
        only_child = calls[calls["person"] == 0]
        assert len(eng.mendel_filter(only_child)) == 0        # no parents: no PositionCall at all
        n_ref, want = O.mendel_stream(calls)
        with pytest.raises(D.DenovoError) as ei:
            eng.mendel_filter(calls, cap=3)
        assert ei.value.code == D.DNV_E_CAPACITY and str(n_ref) in str(ei.value)
        with pytest.raises(D.DenovoError) as ei:
            eng.mendel_filter(calls[::-1].copy())
        assert ei.value.code == D.DNV_E_UNSORTED
        bad = calls.copy()
        i = np.nonzero((bad["person"] == 0) & (bad["end"] == bad["start"] + 1) & (bad["filter_pass"] == 1)
                       & (bad["ngt"] == 2) & (bad["gt0"] >= 0) & (bad["gt1"] >= 0) & (bad["ref_len"] == 1)
                       & (bad["alt0_len"] <= 1) & (bad["alt1_len"] <= 1))[0][0]
        bad["ref_base"][i] = b"N"
        bad["gt0"][i] = 0
        assert O.mendel_stream(bad)[0] < 0
        with pytest.raises(D.DenovoError) as ei:
            eng.mendel_filter(bad)
        assert ei.value.code == D.DNV_E_BADBASE


def test_mendel_filter_full_size_whole_stream_against_the_oracle():
    """BASELINE config 2 at FULL size (>= 11 M (variant, call) records, ~2.3 M child / ~4.7 M records per
    parent): the WHOLE stream goes through the streaming CPU oracle (orc_mendel_stream) and every PositionCall
    is compared field by field; the tiling property (shifted copies of a block give shifted copies of its
    calls) is kept as a second, independent check."""
    import torch
    block = S.make_callset(31, 200_000)
    span = int(block["end"].max()) + 100
    reps = 30
    parts = []
    for r in range(reps):
        b = block.copy()
        b["start"] += r * span
        b["end"] += r * span
        b["variant_id"] += r * 200_000
        parts.append(b)
    calls = np.concatenate(parts)
    assert len(calls) > 11_000_000
    n_all, want_all = O.mendel_stream(calls)
    n_ref, want = O.mendel_stream(block)
    assert n_all == reps * n_ref
    with D.TrioEngine() as eng:
        d_calls = torch.from_numpy(calls.view(np.uint8).reshape(len(calls), 32)).cuda()
        d_out = torch.zeros((len(calls) // 3, 24), dtype=torch.uint8, device="cuda")
        n = eng.mendel_filter_device(d_calls, len(calls), d_out, d_out.shape[0])
        got = d_out[:n].cpu().numpy().view(D.PCALL_DT).reshape(-1)
    assert n == n_all and pcalls_equal(got, want_all)
    for r in (0, 1, reps - 1):
        g = got[r * n_ref:(r + 1) * n_ref]
        assert np.array_equal(g["position"], want["position"] + r * span)
        assert np.array_equal(g["child_index"], want["child_index"] + r * len(block))
        for f in ("child", "mom", "dad", "is_denovo"):
            assert np.array_equal(g[f], want[f])
    assert np.all(np.diff(got["child_index"]) > 0)


def test_pileup_million_sites_window_parity():
    """1M consecutive sites at 30x (the dense layout of configs 3-4): every non-gap base is counted
    exactly once and a 20k-site window in the middle is bit-exact against the oracle."""
    n_sites = 1_000_000
    rs, bo, ba, pe = S.make_reads(41, 1, n_sites, depth=30)
    pos = np.arange(1, n_sites + 1, dtype=np.int64)
    with D.TrioEngine() as eng:
        got = eng.pileup(rs, bo, ba, pe, pos)
    lens = np.diff(bo)
    posn = np.repeat(rs, lens) + (np.arange(bo[-1]) - np.repeat(bo[:-1], lens))
    on_site = (posn >= 1) & (posn <= n_sites) & (ba != ord("-"))
    assert int(got.astype(np.int64).sum()) == int(on_site.sum())
    sel = slice(n_sites // 2, n_sites // 2 + 20000)
    keep = (rs >= pos[sel][0] - 200) & (rs <= pos[sel][-1])
    idx = np.nonzero(keep)[0]
    bo2 = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(lens[idx], out=bo2[1:])
    ba2 = np.concatenate([ba[bo[i]:bo[i + 1]] for i in idx])
    rc, want = O.pileup(rs[idx], bo2, ba2, pe[idx], pos[sel])
    assert rc == 0 and np.array_equal(got[sel], want)
