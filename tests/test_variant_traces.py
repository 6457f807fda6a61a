"""VARIANT path pinned by HAND-TRACED scenarios.

The reference holds no test of checkAndAdd / canProcess / evictParents / retrieveNextCall / getMatchingPair
(SURVEY.md section 4), and no JVM exists here to run it, so each scenario below was traced by hand, statement
by statement, through the reference's own source — the trace is written next to the scenario with the line
it follows ("VB" = M/VariantsBuffer.java, "VC" = M/VariantCaller.java, "DU" = M/DenovoUtil.java) — and the
EXPECTED OUTPUT IS A LITERAL: the exact lines `String.format("%s,%d,%s%n", contig, position, positionCall)`
(VC:178-179,198-200) would append, with PositionCall.toString = "[" + position + "," + TreeMap.toString + "]"
(VB:358-360; TrioMember order CHILD, MOM, DAD, DU:135-139).  Nothing here is computed by tests/host_format.py.

CPU: the streaming oracle must reproduce the literals.  GPU: dnv_mendel_filter and the denovo_main CLI must.
Literal READ-mode lines for two of the reference's JUnit fixtures are at the end."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

G = O.GENOTYPES
PERSON = {"CHILD": 0, "MOM": 1, "DAD": 2}


def flatten(variants):
    """[(start, end, ref, [alts], [(person, [gt...], filter_values)])] -> dnv_vcall records, exactly the fields
    VariantsBuffer reads (include/denovo_b200.h: dnv_vcall)."""
    recs = []
    for vid, (start, end, ref, alts, calls) in enumerate(variants):
        for person, gt, filt in calls:
            def alt(k):
                g = gt[k] if k < len(gt) else 0
                a = alts[g - 1] if g > 0 else ""
                return len(a), (a[:1].encode() or b"\0")
            a0, a1 = alt(0), alt(1)
            recs.append((start, end, vid, gt[0] if len(gt) > 0 else 0, gt[1] if len(gt) > 1 else 0, len(gt),
                         PERSON[person], int(filt == ["PASS"]), min(len(ref), 255), ref[:1].encode(),
                         a0[0], a0[1], a1[0], a1[1], np.zeros(1, np.uint8)))
    out = np.zeros(len(recs), O.VCALL_DT)
    for i, r in enumerate(recs):
        out[i] = r
    return out


def lines(contig, pcalls):
    """VC:176-183 — only isDenovo() PositionCalls are written."""
    return ["%s,%d,[%d,{CHILD=%s, MOM=%s, DAD=%s}]" % (contig, p["position"], p["position"], G[p["child"]],
                                                       G[p["mom"]], G[p["dad"]]) for p in pcalls if p["is_denovo"]]


P = ["PASS"]

# --- S1: a filtered call is not queued and does NOT advance mostRecentStartPosition (VB:172-185, :56-59) ------
# V0 10-11 A>C: CHILD 0/1, MOM 0/0, DAD 0/0 but FILTER=[LowQual] -> passesFilter false (VB:245-249) -> checkAndAdd
#   returns before push (VB:181): DAD queue stays empty, mostRecentStart[DAD] stays 0 (VB:48).
#   canProcess (VB:74-78): child front 10, MOM 10 >= 10, DAD 0 >= 10 false -> nothing processed.
# V1 20-21 G>T: CHILD 0/1, MOM 0/0, DAD 0/0, all PASS.  mostRecentStart = 20 for all three.
#   canProcess true for child@10: evictParents (VB:147-157): MOM front V0 end 11 < 10? no; DAD front V1 21 < 10? no.
#     child genotype (VB:228-239): gt 0 -> REF A, gt 1 -> ALT[0] C -> AC.  MOM getMatchingPair(10) (VB:314-323): V0
#     [10,11) contains 10, isSnp -> 0/0 -> AA.  DAD queue = [V1 20-21]: no entry contains 10 -> null (VB:213-215): NO
#     PositionCall; the child entry is still popped (VC:186).   <- had the filtered DAD call been queued, this would
#     have been {CHILD=AC, MOM=AA, DAD=AA}: a de novo line at 10.
#   canProcess true for child@20: evict MOM front V0 (11 < 20) -> popped; V1 stays.  child G,T -> GT; MOM V1 0/0 -> GG;
#     DAD V1 -> GG.  isDenovo(dad GG, mom GG, child GT) (DU:41-50): (mom has G and dad has T) no, (mom has T and dad has
#     G) no -> de novo -> line.
S1 = [(10, 11, "A", ["C"], [("CHILD", [0, 1], P), ("MOM", [0, 0], P), ("DAD", [0, 0], ["LowQual"])]),
      (20, 21, "G", ["T"], [("CHILD", [0, 1], P), ("MOM", [0, 0], P), ("DAD", [0, 0], P)])]
S1_LINES = ["chr1,20,[20,{CHILD=GT, MOM=GG, DAD=GG}]"]

# --- S2: a parent REFERENCE BLOCK takes the CHILD variant's reference base (VB:217-219), flush at end of stream ---
# V0 100-200 ref T, no ALT: MOM 0/0, DAD 0/0 (end != start+1 -> not a SNP, VB:267; REF length 1 -> admitted).
# V1 150-151 A>G: CHILD 0/1.   canProcess: MOM mostRecentStart 100 >= 150 false -> waits.
# V2 160-161 C>G: CHILD 0/0 (isSnp holds: end = start+1, REF len 1, no called ALT) -> admitted, waits too.
# End of stream: flush loop (VC:193-209), no canProcess.
#   child@150: A,G -> AG.  MOM match: V0 [100,200) contains 150; isSnp(V0) false -> Genotype.valueOf(referenceBase +
#     referenceBase) with referenceBase = the CHILD's "A" (VB:201) -> AA (not TT, the block's own base).  DAD same.
#     isDenovo(AA, AA, AG): mom has A and dad has G? no; mom has G? no -> de novo -> line with MOM=AA, DAD=AA.
#   child@160: CC; parents from the block -> CC, CC -> Mendelian -> no line.
S2 = [(100, 200, "T", [], [("MOM", [0, 0], P), ("DAD", [0, 0], P)]),
      (150, 151, "A", ["G"], [("CHILD", [0, 1], P)]),
      (160, 161, "C", ["G"], [("CHILD", [0, 0], P)])]
S2_LINES = ["chr1,150,[150,{CHILD=AG, MOM=AA, DAD=AA}]"]

# --- S3: getMatchingPair returns the FIRST queued entry that covers the position (VB:314-323) -------------------
# V0 300-400 ref A: MOM 0/0 (block).
# V1 350-351 C>T: CHILD 0/1, MOM 1/1 (a SNP record saying TT), DAD 0/0.
#   after V1: mostRecentStart MOM 350, DAD 350 -> canProcess.  evict: MOM front V0 end 400 < 350? no.
#   child C,T -> CT.  MOM queue in admission order = [V0 300-400, V1 350-351]: the loop returns V0 (first whose
#   closedOpen range contains 350), which is not a SNP -> child's REF twice -> CC.  DAD V1 0/0 -> CC.
#   isDenovo(CC, CC, CT): mom has C and dad has T? no; mom has T? no -> de novo -> line.
#   <- had the later, exactly matching SNP record been preferred, MOM would be TT: (mom has T and dad has C) ->
#      Mendelian, no line.
S3 = [(300, 400, "A", [], [("MOM", [0, 0], P)]),
      (350, 351, "C", ["T"], [("CHILD", [0, 1], P), ("MOM", [1, 1], P), ("DAD", [0, 0], P)])]
S3_LINES = ["chr1,350,[350,{CHILD=CT, MOM=CC, DAD=CC}]"]

# --- S4: eviction pops from the FRONT only, while front.end < child.start (VB:147-157) ---------------------------
# V0 500-1000 ref G: DAD 0/0 (long block).          V1 510-511 A>C: DAD 0/1, MOM 0/0.
# V2 600-601 T>G: CHILD 0/1, MOM 0/0.               canProcess: DAD mostRecentStart 510 >= 600 false -> waits.
# V3 700-701 C>A: DAD 0/0 -> mostRecentStart[DAD] 700 -> canProcess for child@600.
#   evict MOM: front V1 end 511 < 600 -> pop; front V2 end 601 < 600? no.  evict DAD: front V0 end 1000 < 600? no ->
#   loop stops: V1 (end 511) stays queued BEHIND the block.
#   child T,G -> GT (valueOfPairAlleles sorts, DU:229-233).  MOM: V2 [600,601) -> 0/0 -> TT.
#   DAD queue [V0 500-1000, V1 510-511, V3 700-701]: V0 covers 600, not a SNP -> child's REF T twice -> TT.
#   isDenovo(TT, TT, GT) -> de novo -> line.
S4 = [(500, 1000, "G", [], [("DAD", [0, 0], P)]),
      (510, 511, "A", ["C"], [("DAD", [0, 1], P), ("MOM", [0, 0], P)]),
      (600, 601, "T", ["G"], [("CHILD", [0, 1], P), ("MOM", [0, 0], P)]),
      (700, 701, "C", ["A"], [("DAD", [0, 0], P)])]
S4_LINES = ["chr1,600,[600,{CHILD=GT, MOM=TT, DAD=TT}]"]

# --- S5: the interval is half open (Ranges.closedOpen, VB:318): end == position does not cover ---------------------
# V0 800-810 ref A: MOM 0/0, DAD 0/0.    V1 810-811 C>G: CHILD 0/1.    V2 820-821 T>A: MOM 0/0, DAD 0/0.
#   after V2: canProcess (820 >= 810).  evict: front V0 end 810 < 810? no (strict) -> stays, but [800,810) does not
#   contain 810 and [820,821) does not either -> MOM null -> no PositionCall at all.
S5 = [(800, 810, "A", [], [("MOM", [0, 0], P), ("DAD", [0, 0], P)]),
      (810, 811, "C", ["G"], [("CHILD", [0, 1], P)]),
      (820, 821, "T", ["A"], [("MOM", [0, 0], P), ("DAD", [0, 0], P)])]
S5_LINES = []

# --- S6: every admission branch of checkAndAdd (VB:175-180), one defect per position -------------------------------
# At p = 2000 + 100 k: variant p..p+1 A>C(,GT): CHILD 0/1, MOM 0/0, DAD <defect k>; then a clean closing variant at
# p+50 (G, all three 0/0) that makes canProcess true.  With a clean DAD 0/0 the call at p is {CHILD=AC, MOM=AA,
# DAD=AA} (de novo, control k = 0); with the defect the DAD call is rejected, DAD has no entry containing p
# (the closing variant is p+50..p+51) -> null -> no line at p.  The closing child (GG from GG, GG) is Mendelian.
#   k=1 FILTER absent (containsKey false, VB:246)      k=2 FILTER = [PASS, LowQual] (size != 1, VB:247)
#   k=3 genotype ./0 (contains -1, VB:255-258)          k=4 haploid call, genotype size 1 (VB:189-191)
#   k=5 triploid call, size 3                            k=6 called ALT "GT" (gt 0/2 -> isInsertion, VB:297-306)
#   k=7 REF "AT" (isDeletion, VB:288-291; end = start+2)
#   k=8 same two-ALT variant but DAD 0/1: the long ALT is NOT called -> admitted -> DAD = AC -> (mom has A and dad
#       has C) -> Mendelian: no line either, but for a different reason: the PositionCall exists (checked below).
#   k=9 the CHILD record itself is not a SNP (end = start+5) -> rejected (VB:175): nothing at p.


def s6():
    v, want = [], []
    for k in range(10):
        p = 2000 + 100 * k
        ref, alts, end = "A", ["C"], p + 1
        dad = ("DAD", [0, 0], P)
        child = ("CHILD", [0, 1], P)
        if k == 1: dad = ("DAD", [0, 0], None)
        if k == 2: dad = ("DAD", [0, 0], ["PASS", "LowQual"])
        if k == 3: dad = ("DAD", [-1, 0], P)
        if k == 4: dad = ("DAD", [0], P)
        if k == 5: dad = ("DAD", [0, 0, 0], P)
        if k == 6: alts, dad = ["C", "GT"], ("DAD", [0, 2], P)
        if k == 7: ref, end = "AT", p + 2
        if k == 8: alts, dad = ["C", "GT"], ("DAD", [0, 1], P)
        if k == 9: end = p + 5
        if k == 7:  # a deletion record rejects every call on it, the child's included: keep the child on its own SNP
            v.append((p, p + 1, "A", ["C"], [child, ("MOM", [0, 0], P)]))
            v.append((p, end, ref, ["A"], [dad]))
        else:
            v.append((p, end, ref, alts, [child, ("MOM", [0, 0], P), dad]))
        v.append((p + 50, p + 51, "G", ["T"], [("CHILD", [0, 0], P), ("MOM", [0, 0], P), ("DAD", [0, 0], P)]))
        if k == 0:
            want.append("chr1,%d,[%d,{CHILD=AC, MOM=AA, DAD=AA}]" % (p, p))
    return v, want


S6, S6_LINES = s6()

SCENARIOS = {"S1 filtered call": (S1, S1_LINES), "S2 block takes child REF": (S2, S2_LINES),
             "S3 first in queue": (S3, S3_LINES), "S4 front-only eviction": (S4, S4_LINES),
             "S5 half-open": (S5, S5_LINES), "S6 admission branches": (S6, S6_LINES)}
# all scenarios on one contig, in position order: the later scenarios do not disturb the earlier ones (every block
# ends before the next scenario's first position except DAD 500-1000, which covers no later MOM-covered child)
ALL = S1 + S2 + S3 + S4 + S5 + S6
ALL_LINES = S1_LINES + S2_LINES + S3_LINES + S4_LINES + S5_LINES + S6_LINES


@pytest.mark.parametrize("name", list(SCENARIOS) + ["ALL"])
def test_streaming_oracle_reproduces_the_hand_traces(name):
    variants, want = (ALL, ALL_LINES) if name == "ALL" else SCENARIOS[name]
    n, pc = O.mendel_stream(flatten(variants))
    assert n >= 0
    assert lines("chr1", pc) == want


def test_s6_position_calls_exist_exactly_where_traced():
    """k = 0 and k = 8 produce a PositionCall at p (de novo / Mendelian); every rejected DAD call leaves none;
    the closing variants at p+50 always produce one (GG, GG, GG)."""
    n, pc = O.mendel_stream(flatten(S6))
    got = {int(p["position"]): (G[p["child"]], G[p["mom"]], G[p["dad"]], bool(p["is_denovo"])) for p in pc}
    want = {2000: ("AC", "AA", "AA", True), 2800: ("AC", "AA", "AC", False)}
    for k in range(10):
        want[2050 + 100 * k] = ("GG", "GG", "GG", False)
    assert got == want


def test_admission_mask_of_s6_matches_the_trace():
    recs = flatten(S6)
    adm = O.mendel_admit(recs)
    dad = recs["person"] == 2
    first_dad = [int(np.nonzero(dad & (recs["start"] == 2000 + 100 * k))[0][0]) for k in range(10)]
    assert [int(adm[i]) for i in first_dad] == [1, 0, 0, 0, 0, 0, 0, 0, 1, 1]
    child9 = int(np.nonzero((recs["person"] == 0) & (recs["start"] == 2900))[0][0])
    assert adm[child9] == 0   # the child's own non-SNP record (VB:175)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCENARIOS) + ["ALL"])
def test_gpu_filter_reproduces_the_hand_traces(name):
    import denovo_b200 as D
    variants, want = (ALL, ALL_LINES) if name == "ALL" else SCENARIOS[name]
    with D.TrioEngine() as eng:
        pc = eng.mendel_filter(flatten(variants))
    assert lines("chr1", pc) == want


@pytest.mark.gpu
def test_cli_variant_mode_writes_the_hand_traced_file(tmp_path):
    """denovo_main --caller variant over the concatenated scenarios: the output FILE is the literal lines."""
    import host_format as H
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cs = tmp_path / "traces.calls"
    H.write_callset(str(cs), [("chr1", 5000, flatten(ALL))])
    out = tmp_path / "out.csv"
    r = subprocess.run([os.path.join(root, "denovo-variant-caller-java_b200", "bin", "denovo_main"), "--caller", "variant",
                        "--output_file", str(out), "--synthetic_callset", str(cs), "--dad_callset_name", "NA12877",
                        "--mom_callset_name", "NA12878", "--dataset_id", "x"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == "".join(l + "\n" for l in ALL_LINES)


# ---- literal READ-mode lines for two of the reference's JUnit fixtures -------------------------------------------
# Built by hand from M/BayesInfer.java:57-68 ("readCounts=%s,maxGenoType=%s,isDenovo=%b"; entries joined by ';' as
# "NAME:" + TreeMap<Allele,Integer>.toString over TrioMember order CHILD, MOM, DAD; only alleles that were seen) and
# M/ReadCaller.java:126-127 ("%s,%d,%s%n"), with the pileups of T/BayesInferTest.java:193-211 (Chr1Pos75884343) and
# :214-231 (Chr1Pos110583335) and the MAP triples those tests assert (T/BayesInferBayesTest.java:67-84).
READ_FIXTURES = [
    ("chr1", 75884343, [[0, 0, 1, 53], [0, 0, 0, 51], [0, 27, 0, 26]],
     "chr1,75884343,readCounts=CHILD:{C=27, T=26};MOM:{T=51};DAD:{G=1, T=53},maxGenoType=[TT, TT, CT],isDenovo=true"),
    ("chr1", 110583335, [[0, 0, 36, 0], [0, 0, 50, 0], [23, 0, 21, 0]],
     "chr1,110583335,readCounts=CHILD:{A=23, G=21};MOM:{G=50};DAD:{G=36},maxGenoType=[GG, GG, AG],isDenovo=true"),
]


def _read_line(chrom, pos, counts, dad, mom, child, is_denovo):
    def tm(c):
        return "{" + ", ".join("%s=%d" % (a, n) for a, n in zip("ACGT", c) if n) + "}"
    return "%s,%d,readCounts=CHILD:%s;MOM:%s;DAD:%s,maxGenoType=[%s, %s, %s],isDenovo=%s" % (
        chrom, pos, tm(counts[2]), tm(counts[1]), tm(counts[0]), G[dad], G[mom], G[child], "true" if is_denovo else "false")


@pytest.mark.parametrize("method", ["denovo_map", "denovo_bayes", "denovo_lrt"])
def test_oracle_read_lines_for_junit_fixtures_are_the_literals(method):
    for chrom, pos, counts, literal in READ_FIXTURES:
        r = O.infer(O.params(1e-2, 1e-8, 1.0), np.array([counts], np.uint16))[0]
        assert r[method] == 1   # all three methods call both fixtures (MapTest:101-118, BayesTest:67-84, LRTTest:101-118)
        assert _read_line(chrom, pos, counts, r["dad"], r["mom"], r["child"], True) == literal


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["MAP", "BAYES", "LRT"])
def test_gpu_read_lines_for_junit_fixtures_are_the_literals(method):
    import denovo_b200 as D
    counts = np.array([f[2] for f in READ_FIXTURES], np.uint16)
    with D.TrioEngine(1e-2, 1e-8, 1.0, method) as eng:
        calls, _ = eng.infer(counts)
    assert list(calls["site"]) == [0, 1]
    for c, (chrom, pos, cnt, literal) in zip(calls, READ_FIXTURES):
        assert _read_line(chrom, pos, c["counts"].reshape(3, 4), c["dad"], c["mom"], c["child"], bool(c["is_denovo"])) == literal
