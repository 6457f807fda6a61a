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


# ---- level 3 of the screen (csrc/trio_screen.cuh: screen_level3), restated with the SAME integer arithmetic ---------
def level3_constants(eps, mu, thr):
    """tables.cpp: 16.16 fixed-point log2 ratios, the level-3 threshold and its depth limit."""
    log2r_hom = (math.log(1 - eps) - (math.log(eps) - math.log(3))) / math.log(2)
    log2r_het = ((math.log(1 - 2 * eps / 3) - math.log(2)) - (math.log(eps) - math.log(3))) / math.log(2)
    c_hom, c_het = round(log2r_hom * 65536), round(log2r_het * 65536)
    lam = math.log2(9 * thr * (1 - mu) / mu) - 0.0705
    room = 690.0 - abs(math.log(mu / 2304.0))
    nmax = int(room // abs(math.log(eps) - math.log(3))) if room > 0 else 0
    nmax = min(nmax, ((1 << 31) - (1 << 21)) // max(1, c_hom, c_het), 4000)
    if not (1.0 < lam < 30000.0 and nmax >= 3):
        return None
    return c_hom, c_het, math.floor(lam * 65536), nmax


def level3_not_denovo(n, consts):
    """n: int counts [3][4] (DAD, MOM, CHILD).  True = level 3 declares the site not de novo."""
    c_hom, c_het, lam, nmax = consts
    gap = 13 << 16
    if int(n.sum()) > nmax:
        return False
    x = []
    for p in range(2):
        xa = int(np.argmax(n[p]))            # first maximiser, as the device loop
        k1 = int(n[p][xa])
        k2 = max(int(n[p][a]) for a in range(4) if a != xa)
        if k1 * c_hom < max(k2 * c_hom, (k1 + k2) * c_het) + gap:
            return False
        x.append(xa)
    xd, xm = x
    c = [int(v) for v in n[2]]
    t1 = int(np.argmax(c)); k1 = c[t1]
    t2, k2 = -1, -1
    for a in range(4):
        if a != t1 and c[a] > k2:
            k2, t2 = c[a], a
    k3 = max([c[a] for a in range(4) if a != t1 and a != t2] + [0])
    hom_c, het_c = k1 * c_hom, (k1 + k2) * c_het
    het = het_c > hom_c
    e_star = het_c if het else hom_c
    second = max(hom_c, (k1 + k3) * c_het) if het else max(k2 * c_hom, het_c)
    if e_star < second + gap:
        return False
    a1, a2 = t1, (t2 if het else t1)
    e_alt = c[xd] * c_hom if xd == xm else (c[xd] + c[xm]) * c_het
    ratio = e_star - e_alt
    if a1 == xm or a2 == xm:
        u = a2 if a1 == xm else a1
        if u != xd:
            ratio = min(ratio, int(n[0][xd]) * c_hom - (int(n[0][xd]) + int(n[0][u])) * c_het)
    if a1 == xd or a2 == xd:
        u = a2 if a1 == xd else a1
        if u != xm:
            ratio = min(ratio, int(n[1][xm]) * c_hom - (int(n[1][xm]) + int(n[1][u])) * c_het)
    return ratio < lam


def denovo_pattern_counts(rng, n, depth_lo, depth_hi):
    """The sites level 3 is about: parents homozygous with 0-3 stray reads, child with an allele they lack, depths
    spread around the decision boundary."""
    out = np.zeros((n, 3, 4), np.int64)
    for i in range(n):
        xd = int(rng.integers(0, 4))
        xm = xd if rng.random() < 0.8 else int(rng.integers(0, 4))
        y = int((xd + rng.integers(1, 4)) % 4)
        for p, xa in ((0, xd), (1, xm)):
            d = int(rng.integers(depth_lo, depth_hi))
            stray = int(rng.integers(0, 4)) if rng.random() < 0.3 else 0
            out[i, p, xa] = max(d - stray, 0)
            out[i, p, y if rng.random() < 0.5 else int(rng.integers(0, 4))] += stray
        d = int(rng.integers(depth_lo, depth_hi))
        alt = int(rng.integers(0, d + 1))
        out[i, 2, xd if rng.random() < 0.7 else xm] += d - alt
        out[i, 2, y] += alt
    return out.astype(np.uint16)


@pytest.mark.parametrize("eps,mu,lrt", [(1e-2, 1e-8, 1.0), (1e-2, 1e-8, 10.0), (1e-3, 1e-10, 1.0), (1e-3, 1e-6, 1.0),
                                        (5e-2, 1e-7, 0.2), (1e-2, 1e-6, 1.0)])
def test_level3_never_discards_a_site_the_reference_calls(eps, mu, lrt):
    """Whenever level 3 declares a site not de novo, the oracle (reference algorithm) does not call it — for BAYES at
    threshold 1 and for LRT at the given threshold — and level 3 is not vacuous: it decides most of the de novo pattern
    sites the FP64 path would otherwise see, while every site the reference DOES call is left alone."""
    rng = np.random.default_rng(23)
    counts = np.concatenate([denovo_pattern_counts(rng, 6000, 5, 70), denovo_pattern_counts(rng, 3000, 20, 45),
                             S.adversarial_counts(rng, 3000),
                             O.synth_counts(0x5EED0003, 0, 30000, depth=30, seq_err=1e-2)])
    for method, thr in (("denovo_bayes", 1.0), ("denovo_lrt", lrt)):
        consts = level3_constants(eps, mu, thr)
        if 4 * 5334 * mu / (1 - mu) >= 0.999 * thr:   # the screen is off for these parameters (tables.cpp: early_ok)
            continue
        assert consts is not None
        ref = O.infer(O.params(eps, mu, thr), counts, threads=4)
        dec = np.array([level3_not_denovo(c.astype(np.int64), consts) for c in counts])
        assert not ref[method][dec].any(), "level 3 discarded a site the reference calls"
        shallow = counts[:9000].astype(np.int64).sum(axis=(1, 2)) <= consts[3]   # within level 3's depth limit
        assert shallow.sum() > 300 and dec[:9000][shallow].mean() > 0.1   # it decides a good share of the pattern sites
        assert ref[method][:9000].sum() > 100    # ... while plenty of them are real calls, none of which it touched
        if method == "denovo_bayes" and mu < 0.01:
            # MAP uses the T = 1 constants: a de novo argmax needs a de novo triple above every Mendelian one
            assert not ref["denovo_map"][dec].any(), "level 3 discarded a site whose MAP triple is de novo"
