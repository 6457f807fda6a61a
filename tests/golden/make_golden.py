#!/usr/bin/env python
"""Extract the reference's known-answer cases for the hot path into JSON.

Runs ONLY in the authoring container (it reads /root/reference, which does not
exist on the GPU box); the JSON it writes is committed next to it and is what
the tests load.  Nothing is invented here: every number is parsed out of the
reference's JUnit sources.

Sources ("T/" = src/test/java/com/google/cloud/genomics/denovo/):
  T/DenovoTest.java:67-69,91-103    shared params + the two helper pileups
  T/BayesInferTest.java:39-231      nine real-data pileup fixtures
  T/BayesInferMapTest.java          MAP expectations
  T/BayesInferBayesTest.java        BAYES expectations
  T/BayesInferLRTTest.java          LRT expectations
  T/DenovoUtilTest.java:49-96       denovo truth-table cases
  T/DenovoBayesNetTest.java:127-181 prior + child-CPT spot values
"""
import json
import os
import re
import sys

REF = "/root/reference/src/test/java/com/google/cloud/genomics/denovo"
OUT = os.path.dirname(os.path.abspath(__file__))
ALLELES = "ACGT"


def read(name):
    with open(os.path.join(REF, name)) as f:
        return f.read()


def method_bodies(src):
    """name -> body for every method whose header line ends with '{'."""
    out = {}
    pat = re.compile(r"^\s*(?:public\s+|static\s+)*[\w<>,\s\[\]]+?\s+(\w+)\s*\([^)]*\)\s*\{", re.M)
    for m in pat.finditer(src):
        depth, i = 1, m.end()
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        out[m.group(1)] = src[m.end():i - 1]
    return out


def parse_trio_fixture(body):
    """Fixture methods put per-person counts under `if (person == X)`."""
    counts = {"DAD": [0] * 4, "MOM": [0] * 4, "CHILD": [0] * 4}
    cur = None
    for line in body.splitlines():
        m = re.search(r"person\s*==\s*(DAD|MOM|CHILD)", line)
        if m:
            cur = m.group(1)
        for a, n in re.findall(r"baseCount\.put\(Allele\.([ACGT])\s*,\s*(\d+)\)", line):
            counts[cur][ALLELES.index(a)] = int(n)
        if re.match(r"\s*}\s*$", line) and "put" not in line:
            pass
    return counts


def parse_single_fixture(body):
    c = [0] * 4
    for a, n in re.findall(r"baseCount\.put\(Allele\.([ACGT])\s*,\s*(\d+)\)", body):
        c[ALLELES.index(a)] = int(n)
    return c


def main():
    base = method_bodies(read("DenovoTest.java"))
    fixt_src = method_bodies(read("BayesInferTest.java"))
    params_src = read("DenovoTest.java")
    params = {
        "seq_err_rate": float(re.search(r"getSequenceErrorRate\(\)\)\.thenReturn\(([^)]+)\)", params_src).group(1)),
        "denovo_mut_rate": float(re.search(r"getDenovoMutationRate\(\)\)\.thenReturn\(([^)]+)\)", params_src).group(1)),
        "lrt_threshold": float(re.search(r"getLrtThreshold\(\)\)\.thenReturn\(([^)]+)\)", params_src).group(1)),
    }
    fixtures = {}
    for name in ("createSameReadSummary", "createAlmostSameReadSummary"):
        c = parse_single_fixture(base[name])
        fixtures[name] = {"DAD": c, "MOM": c, "CHILD": c}
    for name, body in fixt_src.items():
        if name.startswith("createReadSummaryMap"):
            fixtures[name] = parse_trio_fixture(body)

    cases = []
    for fname, method in (("BayesInferMapTest.java", "MAP"),
                          ("BayesInferBayesTest.java", "BAYES"),
                          ("BayesInferLRTTest.java", "LRT")):
        for tname, body in method_bodies(read(fname)).items():
            if not tname.startswith("test"):
                continue
            fx = re.search(r"(create\w+)\s*\(\s*\)", body).group(1)
            meth = re.search(r"infer\(\s*\w+\s*,\s*(MAP|BAYES|LRT)\s*\)", body).group(1)
            assert meth == method
            is_denovo = bool(re.search(r"assertTrue\(result\.isDenovo\(\)\)", body))
            assert is_denovo != bool(re.search(r"assertFalse\(result\.isDenovo\(\)\)", body))
            tri = re.search(r"Arrays\.asList\(\s*(\w\w)\s*,\s*(\w\w)\s*,\s*(\w\w)\s*\)\s*,\s*result", body)
            cases.append({
                "source": "%s:%s" % (fname, tname), "fixture": fx, "method": meth,
                "counts": fixtures[fx], "expect_is_denovo": is_denovo,
                "expect_map_triple": [tri.group(1), tri.group(2), tri.group(3)],
            })

    util = read("DenovoUtilTest.java")
    truth = [{"label": lab, "expect": exp == "true", "dad": d, "mom": m, "child": c}
             for lab, exp, d, m, c in re.findall(
                 r'assertEquals\("([^"]+)",\s*(true|false),\s*DenovoUtil\.checkTrioGenoTypeIsDenovo\('
                 r'Arrays\.asList\((\w\w),\s*(\w\w),\s*(\w\w)\)\)\)', util)]

    net = read("DenovoBayesNetTest.java")
    prior = [{"genotype": g, "value": float(v)} for v, g in re.findall(
        r"assertEquals\(([\d.eE+-]+),\s*cpt\.get\(Collections\.singletonList\(Genotype\.(\w\w)\)\)", net)]
    cpt = [{"label": lab, "value": float(v), "dad": d, "mom": m, "child": c, "tol": float(t)}
           for lab, v, d, m, c, t in re.findall(
               r'assertEquals\("([^"]+)",\s*([\d.eE+-]+),\s*cpt\.get\(Arrays\.asList\((\w\w),\s*(\w\w),\s*(\w\w)\)\),\s*([\d.eE+-]+)\)', net)]

    doc = {"_generated_by": "tests/golden/make_golden.py from the reference's JUnit sources",
           "params": params, "fixtures": fixtures, "infer_cases": cases,
           "denovo_truth_cases": truth, "parent_prior_cases": prior, "child_cpt_cases": cpt}
    with open(os.path.join(OUT, "reference_junit_cases.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print("fixtures %d, infer cases %d, truth %d, prior %d, cpt %d" %
          (len(fixtures), len(cases), len(truth), len(prior), len(cpt)))


if __name__ == "__main__":
    sys.exit(main())
