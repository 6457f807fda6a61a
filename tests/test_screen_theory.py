"""CPU check of the claim behind the integer screen (csrc/trio_screen.cuh) and its MAP variant, against the
oracle (the reference's algorithm): whenever a triple of per-person genotypes that are each within a factor 2
of that person's best likelihood is Mendelian-consistent, the reference does NOT call the site de novo —
for BAYES and LRT while 4 * 5334 mu / (1 - mu) < threshold, for MAP while mu < 0.01 (tables.cpp).
The GPU kernels decide ~99 % of all sites on this claim alone, so it is pinned here independently of them."""
import math

import numpy as np
import pytest

import oracle_lib as O
import synth_inputs as S

A0 = np.array([0, 0, 0, 0, 1, 1, 1, 2, 2, 3])
A1 = np.array([0, 1, 2, 3, 1, 2, 3, 2, 3, 3])


def genotype_log2_scores(counts, eps):
    """log2 likelihood of each genotype relative to 'every read is a mismatch' (exact reals, float64)."""
    c = counts.astype(np.float64)
    r_hom = math.log2((1 - eps) / (eps / 3))
    r_het = math.log2((0.5 - eps / 3) / (eps / 3))
    s = np.zeros(counts.shape[:2] + (10,))
    for g in range(10):
        s[:, :, g] = c[:, :, A0[g]] * r_hom if A0[g] == A1[g] else (c[:, :, A0[g]] + c[:, :, A1[g]]) * r_het
    return s


def mendelian(d, m, ch):
    has = lambda g, a: (A0[g] == a) | (A1[g] == a)  # noqa: E731
    c1, c2 = A0[ch], A1[ch]
    return (has(m, c1) & has(d, c2)) | (has(m, c2) & has(d, c1))


def screened_out(counts, eps):
    """True where some Mendelian triple uses, for every person, a genotype within 1 bit of the best."""
    s = genotype_log2_scores(counts, eps)
    near = s >= s.max(axis=2, keepdims=True) - 1.0  # [n, 3, 10]
    out = np.zeros(len(counts), bool)
    for d in range(10):
        for m in range(10):
            for ch in range(10):
                if mendelian(d, m, ch):
                    out |= near[:, 0, d] & near[:, 1, m] & near[:, 2, ch]
    return out


@pytest.mark.parametrize("eps,mu,lrt", [(1e-2, 1e-8, 1.0), (1e-3, 1e-10, 1.0), (1e-1, 1e-6, 0.05), (0.3, 1e-5, 20.0),
                                        (2e-2, 4.5e-5, 1.0), (1e-2, 1e-8, 1e-3)])
def test_mendelian_most_likely_triple_is_never_called(eps, mu, lrt):
    rng = np.random.default_rng(11)
    counts = np.concatenate([S.adversarial_counts(rng, 4000),
                             O.synth_counts(0x5EED0003, 0, 20000, depth=30, seq_err=1e-2),
                             O.synth_counts(0x5EED0005, 0, 4000, depth=200, seq_err=1e-3)])
    ref = O.infer(O.params(eps, mu, lrt), counts, threads=4)
    out = screened_out(counts, eps)
    assert out.sum() > 0.5 * len(counts)  # the claim covers most sites
    bound = 4 * 5334 * mu / (1 - mu)
    if bound < 0.999:
        assert not ref["denovo_bayes"][out].any()
    if bound < 0.999 * lrt:
        assert not ref["denovo_lrt"][out].any()
    if mu < 0.01:
        assert not ref["denovo_map"][out].any()
    # and the screen is not vacuous: it leaves every true call to the FP64 path
    for f in ("denovo_bayes", "denovo_lrt", "denovo_map"):
        assert ref[f].sum() > 0
