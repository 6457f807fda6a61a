"""The CPU oracle against every known-answer case the reference's JUnit suite
holds for the hot path (tests/golden/reference_junit_cases.json)."""
import json
import math
import os

import numpy as np
import pytest

import oracle_lib as O

G = O.GENOTYPES
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_junit_cases.json")) as f:
    GOLD = json.load(f)
PRM = O.params(GOLD["params"]["seq_err_rate"], GOLD["params"]["denovo_mut_rate"],
               GOLD["params"]["lrt_threshold"])


def counts_of(fx):
    return np.array([fx["DAD"], fx["MOM"], fx["CHILD"]], np.uint16)[None]


@pytest.mark.parametrize("case", GOLD["infer_cases"], ids=lambda c: c["source"])
@pytest.mark.parametrize("faithful", [False, True])
def test_infer_cases(case, faithful):
    """T/BayesInfer{Map,Bayes,LRT}Test: MAP triple + isDenovo per method."""
    r = O.infer(PRM, counts_of(case["counts"]), faithful=faithful)[0]
    assert [G[r["dad"]], G[r["mom"]], G[r["child"]]] == case["expect_map_triple"]
    got = {"MAP": r["denovo_map"], "BAYES": r["denovo_bayes"], "LRT": r["denovo_lrt"]}[case["method"]]
    assert bool(got) == case["expect_is_denovo"]


def test_faithful_and_hoisted_modes_bit_identical():
    c = np.concatenate([counts_of(fx) for fx in GOLD["fixtures"].values()])
    a, b = O.infer(PRM, c, faithful=False), O.infer(PRM, c, faithful=True)
    assert a.tobytes() == b.tobytes()


def test_denovo_truth_cases():
    """T/DenovoUtilTest.java:49-96"""
    t = O.tables(PRM)["isden"]
    assert len(GOLD["denovo_truth_cases"]) == 21
    for c in GOLD["denovo_truth_cases"]:
        idx = G.index(c["dad"]) * 100 + G.index(c["mom"]) * 10 + G.index(c["child"])
        assert bool(t[idx]) == c["expect"], c
    assert int(t.sum()) == 750


def test_parent_prior_cases():
    """T/DenovoBayesNetTest.java:127-139"""
    P = O.tables(PRM)["P"]
    assert len(GOLD["parent_prior_cases"]) == 10
    for c in GOLD["parent_prior_cases"]:
        assert abs(P[G.index(c["genotype"])] - c["value"]) <= 1e-12
    assert abs(P.sum() - 1.0) <= 1e-12


def test_child_cpt_cases_and_row_sums():
    """T/DenovoBayesNetTest.java:157-198"""
    T = O.tables(PRM)["T"]
    assert len(GOLD["child_cpt_cases"]) == 23
    for c in GOLD["child_cpt_cases"]:
        v = T[G.index(c["dad"]) * 100 + G.index(c["mom"]) * 10 + G.index(c["child"])]
        assert abs(v - c["value"]) <= c["tol"], c
    assert np.all(np.abs(T.reshape(100, 10).sum(1) - 1.0) <= 1e-12)
    assert len(np.unique(T)) == 7


def test_base_loglik_formulas():
    """T/DenovoBayesNetTest.java:201-243 (EPS 1e-12, eps = 1e-2)"""
    BL = O.tables(PRM)["BL"]
    for g, name in enumerate(G):
        for a, al in enumerate("ACGT"):
            hom = name[0] == name[1]
            if hom and al in name:
                exp = math.log(0.99)
            elif not hom and al in name:
                exp = math.log(0.5 * (1 - 2 * 1e-2 / 3))
            else:
                exp = math.log(1e-2 / 3)
            assert abs(BL[g, a] - exp) <= 1e-12


def test_read_loglik():
    """T/DenovoBayesNetTest.java:246-277: sum count*baseLL; empty pileup -> 0.0"""
    BL = O.tables(PRM)["BL"]
    assert np.all(O.read_loglik(PRM, [0, 0, 0, 0]) == 0.0)
    for c in ([40, 0, 0, 0], [38, 2, 3, 0], [1, 245, 0, 10]):
        L = O.read_loglik(PRM, c)
        for g in range(10):
            assert abs(L[g] - sum(c[a] * BL[g, a] for a in range(4))) <= 1e-12 * max(1, abs(L[g]))


def test_max_likelihood_identity():
    """T/DenovoBayesNetTest.java:406-433: maxLL == 3*ll(AA) + log CPT terms."""
    t = O.tables(PRM)
    c = np.array([[40, 0, 0, 0]] * 3, np.uint16)[None]
    r = O.infer(PRM, c)[0]
    ll_aa = O.read_loglik(PRM, [40, 0, 0, 0])[0]
    expect = 3 * ll_aa + 2 * math.log(t["P"][0]) + math.log(t["T"][0])
    assert abs(r["max_ll"] - expect) <= 1e-12
    assert abs(r["max_ll"] - (-6.751217756899736)) <= 1e-12  # SURVEY Appendix B probe value


def test_zero_depth_site():
    """All-zero pileup: first-max tie-break gives (AC,AC,AC); p = mu-scale, not denovo."""
    r = O.infer(PRM, np.zeros((1, 3, 4), np.uint16))[0]
    assert [G[r["dad"]], G[r["mom"]], G[r["child"]]] == ["AC", "AC", "AC"]
    assert not r["denovo_bayes"] and not r["denovo_map"] and not r["denovo_lrt"]


def test_underflow_special_values():
    """raw exp sums, no max-shift -> 0 / inf / NaN rows."""
    deep = np.array([[[30000, 30000, 0, 0]] * 3], np.uint16)
    r = O.infer(PRM, deep)[0]
    assert r["l_den"] == 0.0 and r["l_men"] == 0.0
    assert math.isnan(r["p"]) and math.isnan(r["lr"])
    assert not r["denovo_bayes"] and not r["denovo_lrt"]
    assert r["log_den"] == -math.inf and r["log_men"] == -math.inf


def test_threads_do_not_change_results():
    c = O.synth_counts(0x5EED0001, 0, 2000)
    a = O.infer(PRM, c, threads=1)
    b = O.infer(PRM, c, threads=4)
    assert a.tobytes() == b.tobytes()
    rate = a["denovo_bayes"].mean()
    assert 0.0005 < rate < 0.02
