"""The C++ host CLI (denovo_main): DenovoMain's flags, caller modes and output-file formats.
CPU part: flag handling and loud failure without a GPU.  GPU part: output files line by line
against the Python restatement of the formats driven by the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import host_format as H
import oracle_lib as O
import synth_inputs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "denovo-variant-caller-java_b200", "bin", "denovo_main")
REQ = ["--dad_callset_name", "NA12877", "--mom_callset_name", "NA12878", "--child_callset_name", "NA12879",
       "--dataset_id", "3049512673186936334"]


def run(args, **kw):
    return subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600, **kw)


def test_binary_exists():
    assert os.path.exists(BIN), "run __graft_entry__.build()"


def test_missing_required_flag_prints_usage_and_exits_1():
    r = run(["--caller", "read"])
    assert r.returncode == 1
    assert 'Option "--output_file" is required' in r.stderr and "Usage: DenovoMain [flags...]" in r.stderr
    r = run(["--output_file", "x"] + REQ)
    assert r.returncode == 1 and 'Option "--caller" is required' in r.stderr
    r = run(["--caller", "bogus", "--output_file", "x"] + REQ)
    assert r.returncode == 1 and "is not a valid value" in r.stderr


def test_usage_lists_the_reference_flags():
    r = run(["--help"])
    assert r.returncode == 0
    for flag in ("--caller", "--inference_method", "--output_file", "--output_dir", "--input_calls_file",
                 "--dad_callset_name", "--mom_callset_name", "--child_callset_name", "--dataset_id", "--seq_err_rate",
                 "--denovo_mut_rate", "--lrt_threshold", "--num_threads", "--log_level", "--log_file", "--chromosome",
                 "--start_position", "--end_position", "--max_variant_results"):
        assert flag in r.stdout, flag


def test_read_mode_needs_input_file(tmp_path):
    r = run(["--caller", "read", "--output_file", str(tmp_path / "o")] + REQ)
    assert r.returncode != 0 and "Input calls file needed for read mode" in r.stderr


def test_unparsable_candidate_line_is_a_parse_exception(tmp_path):
    cand = tmp_path / "c.txt"
    cand.write_text("justonefield\n")
    r = run(["--caller", "READ", "--output_file", str(tmp_path / "o"), "--input_calls_file", str(cand)] + REQ)
    assert "Could not parse line : justonefield" in r.stderr


def test_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cand = tmp_path / "c.txt"
    cand.write_text("chr1,100\n")
    r = run(["--caller", "read", "--output_file", str(tmp_path / "o"), "--input_calls_file", str(cand)] + REQ)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def candidates(n, seed=0):
    """Candidate sites, some of them closer together than a read is long (their synthetic reads then contribute to
    each other's pileups, as real reads would), one of them listed twice."""
    rng = np.random.default_rng(seed)
    chroms = ["chr1", "chr2", "chrX", "CHR7", "chr21"]
    out = [(chroms[int(rng.integers(0, len(chroms)))], int(rng.integers(200, 40_000_000))) for _ in range(n)]
    for k in range(0, n - 1, 97):
        out[k + 1] = (out[k][0], out[k][1] + int(rng.integers(1, 60)))
    out[n // 2] = out[n // 3]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["bayes", "map", "lrt"])
def test_read_mode_output_file(tmp_path, method):
    """READ mode end to end on its default source: candidate file -> synthetic reads -> dnv_pileup (the pileup kernel)
    -> interval-sharded inference on two contexts (--devices 0,0) -> output file, line for line against the oracle
    chain (reads spec -> ReadSummary rule -> performInference -> format strings)."""
    cands = candidates(12000)
    cand = tmp_path / "cands.csv"
    cand.write_text("".join("%s,%d,extra,fields\n" % c for c in cands))
    out = tmp_path / "calls.csv"
    r = run(["--caller", "read", "--inference_method", method, "--output_file", str(out), "--input_calls_file",
             str(cand), "--synthetic_seed", str(0x5EED0001), "--log_level", "info", "--num_threads", "4",
             "--devices", "0,0"] + REQ)
    assert r.returncode == 0, r.stderr
    got = out.read_text().splitlines()
    want = H.expected_read_lines(cands, 0x5EED0001, method=method.upper())
    assert len(want) > 20
    assert got == want
    assert "INFO : ---- Starting Bayesian Read Caller ----" in r.stderr
    assert "INFO : 100 Read candidates processed" in r.stderr


@pytest.mark.gpu
def test_read_mode_counts_source_and_short_reads(tmp_path):
    """--synthetic_source counts (pileup counts straight from the generator) and a 40-base read source."""
    cands = candidates(20000, seed=9)
    cand = tmp_path / "cands.csv"
    cand.write_text("".join("%s,%d\n" % c for c in cands))
    out = tmp_path / "calls.csv"
    r = run(["--caller", "read", "--output_file", str(out), "--input_calls_file", str(cand), "--synthetic_source", "counts",
             "--synthetic_seed", "11"] + REQ)
    assert r.returncode == 0, r.stderr
    want = H.expected_read_lines(cands, 11, source="counts")
    assert out.read_text().splitlines() == want and len(want) > 10
    r = run(["--caller", "read", "--output_file", str(out), "--input_calls_file", str(cand), "--synthetic_read_length", "40",
             "--synthetic_seed", "11", "--devices", "0-0"] + REQ)
    assert r.returncode == 0, r.stderr
    want = H.expected_read_lines(cands, 11, read_len=40)
    assert out.read_text().splitlines() == want and len(want) > 10


def test_devices_flag_is_validated():
    r = run(["--caller", "read", "--output_file", "x", "--devices", "3-1"] + REQ)
    assert r.returncode == 1 and "--devices" in r.stderr


@pytest.mark.gpu
def test_read_mode_chromosome_filter_and_params(tmp_path):
    cands = candidates(20000, seed=3)
    cand = tmp_path / "cands.csv"
    cand.write_text("".join("%s,%d\n" % c for c in cands))
    out = tmp_path / "calls.csv"
    r = run(["--caller", "read", "--output_file", str(out), "--input_calls_file", str(cand), "--chromosome", "chr1",
             "--chromosome", "X", "--seq_err_rate", "5e-3", "--denovo_mut_rate", "1e-7", "--synthetic_depth", "60",
             "--synthetic_seed", "77"] + REQ)
    assert r.returncode == 0, r.stderr
    want = H.expected_read_lines(cands, 77, depth=60, eps=5e-3, mu=1e-7, chromosomes={"CHR1", "CHRX"})
    assert out.read_text().splitlines() == want and len(want) > 5


def make_contigs():
    a = S.make_callset(21, 20000)
    b = S.make_callset(22, 8000)
    c = S.make_callset(23, 3000)
    return [("chr1", int(a["end"].max()) + 10, a), ("chr2", int(b["end"].max()) + 10, b),
            ("chrX", int(c["end"].max()) + 10, c)]


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 4])
def test_variant_mode_output_file(tmp_path, threads):
    contigs = make_contigs()
    cs = tmp_path / "trio.calls"
    H.write_callset(str(cs), contigs)
    out = tmp_path / "variants.csv"
    r = run(["--caller", "variant", "--output_file", str(out), "--synthetic_callset", str(cs), "--num_threads",
             str(threads), "--chromosome", "chr1", "--chromosome", "chrX"] + REQ)
    assert r.returncode == 0, r.stderr
    want = H.expected_variant_lines(contigs, num_threads=threads, chromosomes={"CHR1", "CHRX"})
    assert len(want) > 50
    assert out.read_text().splitlines() == want
    assert "INFO : Processing Chromosome : chr1" in r.stderr and "chr2" not in r.stderr


@pytest.mark.gpu
def test_full_mode_chains_variant_then_read(tmp_path):
    contigs = make_contigs()[:1]
    cs = tmp_path / "trio.calls"
    H.write_callset(str(cs), contigs)
    out = tmp_path / "full.csv"
    r = run(["--caller", "full", "--output_file", str(out), "--synthetic_callset", str(cs), "--lrt_threshold", "0.01",
             "--inference_method", "LRT", "--synthetic_seed", "5"] + REQ)
    assert r.returncode == 0, r.stderr
    stage1 = H.expected_variant_lines(contigs)
    assert (tmp_path / "full.csv.tmp").read_text().splitlines() == stage1
    cands = [(l.split(",")[0], int(l.split(",")[1])) for l in stage1]
    want = H.expected_read_lines(cands, 5, lrt=0.01, method="LRT")
    assert out.read_text().splitlines() == want
