"""Python restatement of the reference's output-file formats and of the host's
synthetic sources, used to check the C++ CLI (denovo_main) line by line.
  READ line     "%s,%d,%s%n" + BayesInfer details      M/ReadCaller.java:126-127, M/BayesInfer.java:57-68
  VARIANT line  "%s,%d,%s%n" + PositionCall.toString   M/VariantCaller.java:178-179, M/VariantsBuffer.java:358-360
"""
import struct

import numpy as np

import oracle_lib as O

G = O.GENOTYPES
CHROMS = ["CHR%d" % i for i in range(1, 23)] + ["CHRX", "CHRY", "CHRM"]
CONTIG_LEN = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022, 141213431,
              135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753, 81195210, 78077248,
              59128983, 63025520, 48129895, 51304566, 155270560, 59373566, 16571]
assert sum(CONTIG_LEN) == 3095693983


def global_index(chrom, pos):
    return sum(CONTIG_LEN[:CHROMS.index(chrom.upper())]) + pos


def tree_map(counts4):
    """TreeMap<Allele,Integer>.toString with only alleles seen at least once."""
    return "{" + ", ".join("%s=%d" % (a, n) for a, n in zip("ACGT", counts4) if n) + "}"


def read_line(chrom, pos, counts, dad, mom, child, is_denovo=True):
    return "%s,%d,readCounts=CHILD:%s;MOM:%s;DAD:%s,maxGenoType=[%s, %s, %s],isDenovo=%s" % (
        chrom, pos, tree_map(counts[2]), tree_map(counts[1]), tree_map(counts[0]), G[dad], G[mom], G[child],
        "true" if is_denovo else "false")


def expected_read_lines(cands, seed, depth=30, eps=1e-2, mu=1e-8, lrt=1.0, method="BAYES", chromosomes=None,
                        source="reads", read_len=100):
    """source="reads" (the CLI's default): the oracle's synthetic reads at the candidates -> oracle pileup
    (ReadSummary rule) -> oracle inference; source="counts": pileup counts straight from the generator."""
    cands = [(c, p) for c, p in cands if chromosomes is None or c.upper() in chromosomes]
    if not cands:
        return []
    idx = np.array([global_index(c, p) for c, p in cands], np.int64)
    if source == "counts":
        counts = O.synth_counts_at(seed, idx, depth=depth, seq_err=eps)
    else:
        uniq = np.unique(idx)
        rs, bo, ba, pe = O.synth_reads_at(seed, uniq, depth=depth, seq_err=eps, read_len=read_len)
        rc, ucounts = O.pileup_sorted(rs, bo, ba, pe, uniq)
        assert rc == 0
        counts = ucounts[np.searchsorted(uniq, idx)]
    ref = O.infer(O.params(eps, mu, lrt), counts)
    field = {"MAP": "denovo_map", "BAYES": "denovo_bayes", "LRT": "denovo_lrt"}[method]
    return [read_line(c, p, counts[i], ref["dad"][i], ref["mom"][i], ref["child"][i])
            for i, (c, p) in enumerate(cands) if ref[field][i]]


def variant_line(chrom, pc):
    return "%s,%d,[%d,{CHILD=%s, MOM=%s, DAD=%s}]" % (chrom, pc["position"], pc["position"], G[pc["child"]],
                                                      G[pc["mom"]], G[pc["dad"]])


def expected_variant_lines(contigs, num_threads=1, start=None, end=None, chromosomes=None):
    """contigs: [(name, upper_bound, records)].  Per-thread closed position strides with a fresh buffer
    each (M/VariantCaller.java:88-103); the source returns records overlapping the shard."""
    out = []
    for name, upper, recs in contigs:
        if chromosomes is not None and name.upper() not in chromosomes:
            continue
        s0 = 1 if start is None else start
        e0 = upper if end is None else end
        stride = (e0 - s0) // num_threads
        for t in range(num_threads):
            s = s0 + t * stride
            e = e0 if t == num_threads - 1 else s + stride - 1
            shard = recs[(recs["start"] <= e) & (recs["end"] > s)]
            n, pc = O.mendel_stream(shard)
            assert n >= 0
            out += [variant_line(name, r) for r in pc if r["is_denovo"]]
    return out


def write_callset(path, contigs):
    with open(path, "wb") as f:
        f.write(b"DNVCALL1")
        f.write(struct.pack("<i", len(contigs)))
        for name, upper, recs in contigs:
            f.write(name.encode().ljust(32, b"\0"))
            f.write(struct.pack("<qq", upper, len(recs)))
            f.write(np.ascontiguousarray(recs, O.VCALL_DT).tobytes())
