"""CPU-only checks of the oracle's pileup and VARIANT-mode restatements against
independent formulations (the reference has no tests for these paths beyond one
PositionCall case, T/VariantsBufferTest.java:102-111)."""
import numpy as np

import oracle_lib as O
import synth_inputs as S


def pcall_tuples(pc):
    return [(int(r["position"]), int(r["child_index"]), int(r["child"]), int(r["mom"]), int(r["dad"]),
             int(r["is_denovo"])) for r in pc]


def test_position_call_case_from_reference_test():
    """T/VariantsBufferTest.java:102-111: (CHILD=AC, MOM=AA, DAD=AA) is denovo."""
    isden = O.tables(O.params())["isden"]
    G = O.GENOTYPES
    assert isden[G.index("AA") * 100 + G.index("AA") * 10 + G.index("AC")] == 1
    calls = np.zeros(3, S.VCALL_DT)
    for i, (p, gt) in enumerate(((1, (0, 0)), (2, (0, 0)), (0, (0, 1)))):
        calls[i] = (100, 101, 0, gt[0], gt[1], 2, p, 1, 1, b"A", 0, b"\0", 1 if gt[1] else 0, b"C" if gt[1] else b"\0", [0])
    n, pc = O.mendel_stream(calls)
    assert n == 1 and pcall_tuples(pc) == [(100, 2, G.index("AC"), 0, 0, 1)]


def test_streaming_equals_batch_definition():
    isden = O.tables(O.params())["isden"]
    for seed in range(12):
        calls = S.make_callset(seed, 300, adversarial=seed % 2 == 1)
        n, pc = O.mendel_stream(calls)
        assert n >= 0
        assert pcall_tuples(pc) == S.batch_mendel(calls, isden), seed
    calls = S.make_callset(99, 4000)
    n, pc = O.mendel_stream(calls)
    assert n > 100 and 0 < pc["is_denovo"].mean() < 1


def test_admission_rules():
    base = (10, 11, 0, 0, 1, 2, 0, 1, 1, b"A", 0, b"\0", 1, b"C", [0])
    def rec(**kw):
        r = np.zeros(1, S.VCALL_DT)
        r[0] = base
        for k, v in kw.items():
            r[k] = v
        return r
    assert O.mendel_admit(rec())[0] == 1
    assert O.mendel_admit(rec(filter_pass=0))[0] == 0            # FILTER != PASS
    assert O.mendel_admit(rec(gt0=-1))[0] == 0                   # './.'
    assert O.mendel_admit(rec(ngt=1))[0] == 0 and O.mendel_admit(rec(ngt=3))[0] == 0
    assert O.mendel_admit(rec(ref_len=2))[0] == 0                # deletion
    assert O.mendel_admit(rec(alt1_len=2))[0] == 0               # insertion
    assert O.mendel_admit(rec(end=15))[0] == 0                   # child must be a SNP
    assert O.mendel_admit(rec(end=15, person=1))[0] == 1         # parent reference block is admitted
    assert O.mendel_admit(rec(end=15, person=2, gt1=0, alt1_len=0))[0] == 1


def test_bad_base_is_an_error():
    calls = np.zeros(1, S.VCALL_DT)
    calls[0] = (10, 11, 0, 0, 1, 2, 0, 1, 1, b"N", 0, b"\0", 1, b"C", [0])
    n, _ = O.mendel_stream(calls)
    assert n == -1


def test_pileup_oracle_matches_numpy():
    rs, bo, ba, pe = S.make_reads(3, 1000, 400, depth=12, read_len=40, var_len=True, gap_rate=0.02)
    pos = np.arange(990, 1420, dtype=np.int64)
    rc, got = O.pileup(rs, bo, ba, pe, pos)
    assert rc == 0
    assert np.array_equal(got, S.numpy_pileup(rs, bo, ba, pe, pos))
    assert 8 < got[50:350].sum(axis=2).mean() < 16
    ba2 = ba.copy()
    ba2[bo[5] + 3] = ord("N")
    rc, _ = O.pileup(rs, bo, ba2, pe, np.array([rs[5] + 3]))
    assert rc == -6


def test_sorted_pileup_oracle_equals_the_scan_and_read_source_properties():
    """orc_pileup_sorted (binary-searched windows, for big inputs) == orc_pileup (the plain ReadSummary scan), on the
    synthetic read source; an isolated candidate's pileup equals the counts generator's, reads are sorted by
    (person, start) and every read covers its site."""
    import numpy as np
    import oracle_lib as O
    import synth_inputs as S
    idx = np.unique(np.concatenate([np.random.default_rng(5).integers(10_000, 50_000_000, 300),
                                    np.array([777, 800, 801, 850, 10**9, 10**9 + 99])])).astype(np.int64)
    rs, bo, ba, pe = O.synth_reads_at(0x5EED0001, idx, depth=30, read_len=100)
    assert np.all(np.diff(pe.astype(int)) >= 0) and all(np.all(np.diff(rs[pe == p]) >= 0) for p in range(3))
    assert set(np.unique(ba)) <= set(b"ACGT-") and (ba == ord("-")).mean() < 5e-3
    rc1, c1 = O.pileup(rs, bo, ba, pe, idx)
    rc2, c2 = O.pileup_sorted(rs, bo, ba, pe, idx)
    assert rc1 == 0 and rc2 == 0 and np.array_equal(c1, c2)
    counts = O.synth_counts_at(0x5EED0001, idx, depth=30)
    near = np.zeros(len(idx), bool)
    near[1:] |= np.diff(idx) < 100
    near[:-1] |= np.diff(idx) < 100
    assert np.array_equal(c1[~near], counts[~near]) and near.sum() >= 5
    assert np.all(c1.astype(int).sum(axis=2) >= counts.astype(int).sum(axis=2) - 0)   # neighbours only add reads
    # variable-length reads, unsorted sites
    rs, bo, ba, pe = S.make_reads(3, 100, 3000, depth=20, var_len=True)
    pos = np.random.default_rng(1).permutation(np.arange(50, 3200)).astype(np.int64)
    assert np.array_equal(O.pileup(rs, bo, ba, pe, pos)[1], O.pileup_sorted(rs, bo, ba, pe, pos)[1])
