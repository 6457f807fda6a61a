"""Seeded synthetic inputs for the pileup kernel (reads) and the
VARIANT-mode filter (trio callsets).  numpy only; used by tests and bench."""
import numpy as np

VCALL_DT = np.dtype([("start", "<i8"), ("end", "<i8"), ("variant_id", "<i4"), ("gt0", "i1"), ("gt1", "i1"),
                     ("ngt", "u1"), ("person", "u1"), ("filter_pass", "u1"), ("ref_len", "u1"), ("ref_base", "S1"),
                     ("alt0_len", "u1"), ("alt0_base", "S1"), ("alt1_len", "u1"), ("alt1_base", "S1"),
                     ("pad_", "u1", (1,))])
BASES = np.frombuffer(b"ACGT", np.uint8)


def make_reads(seed, region_start, region_len, depth=30, read_len=100, gap_rate=1e-3, err=1e-2,
               var_len=False):
    """Reads of DAD, MOM, CHILD (packed person order 0,1,2) tiling [region_start, region_start+region_len)
    with Poisson(depth) coverage; returns read_start, base_off, bases(uint8 ASCII), person, sorted by
    (person, start).  Bases: a per-position reference base, sequencing errors at `err`, '-' at gap_rate."""
    rng = np.random.default_rng(seed)
    ref = BASES[rng.integers(0, 4, region_len + 2 * read_len)]
    starts, lens, persons = [], [], []
    for p in range(3):
        n = rng.poisson(depth * (region_len + read_len) / read_len)
        st = np.sort(rng.integers(region_start - read_len + 1, region_start + region_len, n))
        ln = rng.integers(read_len // 2, read_len + 1, n) if var_len else np.full(n, read_len)
        starts.append(st); lens.append(ln); persons.append(np.full(n, p, np.uint8))
    read_start = np.concatenate(starts).astype(np.int64)
    lens = np.concatenate(lens).astype(np.int64)
    person = np.concatenate(persons)
    base_off = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=base_off[1:])
    # gather reference bases for every read position
    idx = np.repeat(read_start - (region_start - read_len), lens) + (np.arange(base_off[-1]) - np.repeat(base_off[:-1], lens))
    bases = ref[idx].copy()
    e = rng.random(len(bases)) < err
    bases[e] = BASES[rng.integers(0, 4, int(e.sum()))]
    g = rng.random(len(bases)) < gap_rate
    bases[g] = ord("-")
    return read_start, base_off, bases, person


def numpy_pileup(read_start, base_off, bases, person, site_pos):
    """Independent vectorised restatement of ReadSummary for dense cross-checks."""
    out = np.zeros((len(site_pos), 3, 4), np.int64)
    lens = np.diff(base_off)
    pos_all = np.repeat(read_start, lens) + (np.arange(base_off[-1]) - np.repeat(base_off[:-1], lens))
    per_all = np.repeat(person, lens)
    site_index = {int(p): i for i, p in enumerate(site_pos)}
    lut = np.full(256, -1)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    a = lut[bases]
    keep = a >= 0
    sidx = np.array([site_index.get(int(p), -1) for p in pos_all[keep]])
    ok = sidx >= 0
    np.add.at(out, (sidx[ok], per_all[keep][ok], a[keep][ok]), 1)
    return out.astype(np.uint16)


def make_callset(seed, n_variants, span=None, adversarial=False):
    """One start-sorted variant-centric stream flattened to (variant, call) records.
    Mix: child/parent SNP records, parental reference blocks covering most positions,
    non-PASS, './.', indels, multi-allelic ALT index 2, haploid/triploid calls, duplicate starts."""
    rng = np.random.default_rng(seed)
    span = span or n_variants * 40
    starts = np.sort(rng.integers(1, span, n_variants)).astype(np.int64)
    if adversarial:  # many duplicate starts and abutting / nested intervals
        starts = np.sort(rng.integers(1, max(n_variants // 3, 4), n_variants)).astype(np.int64)
    recs = []
    vid = 0
    for s in starts:
        kind = rng.random()
        ref = bytes([BASES[rng.integers(0, 4)]])
        if kind < 0.45:  # SNP record with calls for a random subset of the trio
            alts = [bytes([b]) for b in BASES[rng.permutation(4)][:2] if bytes([b]) != ref][:2]
            members = [p for p in (0, 1, 2) if rng.random() < 0.8] or [0]
            rng.shuffle(members)
            end = s + 1
            ref_len = 1
            if rng.random() < 0.03:  # deletion: REF longer than one base
                ref_len = 3; end = s + 3
            for p in members:
                gt = list(rng.integers(0, len(alts) + 1, 2))
                r = rng.random()
                ngt = 2
                if r < 0.02: gt[rng.integers(0, 2)] = -1           # './.'
                elif r < 0.03: ngt = 1
                elif r < 0.04: ngt = 3
                a_len = [0, 0]; a_base = [b"\0", b"\0"]
                for k in range(2):
                    if gt[k] > 0:
                        a_len[k] = 1; a_base[k] = alts[gt[k] - 1]
                        if rng.random() < 0.02: a_len[k] = 2          # insertion: ALT longer than one base
                recs.append((s, end, vid, gt[0], gt[1], ngt, p, int(rng.random() > 0.03), ref_len, ref,
                             a_len[0], a_base[0], a_len[1], a_base[1]))
        else:  # gVCF-style reference block for one or both parents (sometimes the child: rejected, not a SNP)
            ln = int(rng.geometric(1 / 60.0)) + 1 if not adversarial else int(rng.integers(1, 6))
            members = [p for p in (1, 2) if rng.random() < 0.75] or [1]
            if rng.random() < 0.05: members.append(0)
            for p in members:
                recs.append((s, s + ln, vid, 0, 0, 2, p, int(rng.random() > 0.03), 1, ref, 0, b"\0", 0, b"\0"))
        vid += 1
    out = np.zeros(len(recs), VCALL_DT)
    for i, r in enumerate(recs):
        out[i] = r + (np.zeros(1, np.uint8),)
    return out


def batch_mendel(calls, isden):
    """The BATCH definition in plain Python: an independent check that the
    streaming oracle and the batch form agree on start-sorted input with end > start."""
    def alle(c):
        return b"ACGT".index(c)

    def geno(a, b):
        a, b = min(a, b), max(a, b)
        return a * 4 - a * (a - 1) // 2 + (b - a)

    adm = []
    for v in calls:
        ins = (v["ngt"] >= 1 and v["gt0"] > 0 and v["alt0_len"] != 1) or (v["ngt"] >= 2 and v["gt1"] > 0 and v["alt1_len"] != 1)
        dele = v["ref_len"] != 1
        snp = v["end"] == v["start"] + 1 and not dele and not ins
        dot = (v["ngt"] >= 1 and v["gt0"] == -1) or (v["ngt"] >= 2 and v["gt1"] == -1)
        adm.append(not ((v["person"] == 0 and not snp) or dot or v["ngt"] != 2 or not v["filter_pass"] or ins or dele))
    q = {p: [i for i, v in enumerate(calls) if adm[i] and v["person"] == p] for p in (0, 1, 2)}

    def g_of(v):
        a0 = alle(v["alt0_base"] if v["gt0"] > 0 else v["ref_base"])
        a1 = alle(v["alt1_base"] if v["gt1"] > 0 else v["ref_base"])
        return geno(a0, a1)

    out = []
    for ci in q[0]:
        c = calls[ci]
        pos = c["start"]
        g = {0: g_of(c)}
        ok = True
        for par in (1, 2):
            hit = next((i for i in q[par] if calls[i]["start"] <= pos < calls[i]["end"]), None)
            if hit is None:
                ok = False
                break
            h = calls[hit]
            g[par] = g_of(h) if h["end"] == h["start"] + 1 else geno(alle(c["ref_base"]), alle(c["ref_base"]))
        if ok:
            out.append((int(pos), int(ci), g[0], g[1], g[2], int(isden[g[2] * 100 + g[1] * 10 + g[0]])))
    return out


def adversarial_counts(rng, n):
    """Pileups built to sit on the integer screen's decision boundaries: one or two dominant alleles with
    0..6 stray reads, equal counts (ties between genotypes), depths around the uint8 and power-table
    limits, empty persons."""
    c = np.zeros((n, 3, 4), np.int64)
    kind = rng.integers(0, 8, n)
    depth = rng.choice([1, 2, 5, 12, 30, 31, 60, 127, 128, 254, 255, 256, 300, 1022, 1023, 1024, 1100, 4000], n)
    for i in range(n):
        x, y = rng.choice(4, 2, replace=False)
        for p in range(3):
            d = int(depth[i]) if kind[i] != 7 else int(rng.integers(0, 40))
            k = kind[i]
            if k == 0:      # everybody homozygous X with a few errors
                stray = int(rng.integers(0, 7))
                c[i, p, x] = max(d - stray, 0)
                c[i, p, rng.integers(0, 4)] += stray
            elif k == 1:    # one parent het, child anything
                if p == int(rng.integers(0, 2)):
                    c[i, p, x] = d // 2; c[i, p, y] = d - d // 2
                else:
                    c[i, p, x] = d
            elif k == 2:    # child het, parents hom: the de novo pattern, with 0..3 alt reads leaking into a parent
                if p == 2:
                    a = int(rng.integers(0, d + 1)); c[i, p, x] = d - a; c[i, p, y] = a
                else:
                    a = int(rng.integers(0, 4)); c[i, p, x] = max(d - a, 0); c[i, p, y] = min(a, d)
            elif k == 3:    # exact ties
                c[i, p, x] = d; c[i, p, y] = d
            elif k == 4:    # hom-vs-het boundary: n_X * log2 r_hom ~ N * log2 r_het
                a = int(round(d * rng.uniform(0.08, 0.16))); c[i, p, x] = d - a; c[i, p, y] = a
            elif k == 5:    # three alleles present
                parts = rng.multinomial(d, [0.5, 0.3, 0.2]); c[i, p, x] = parts[0]; c[i, p, y] = parts[1]
                c[i, p, min(set(range(4)) - {int(x), int(y)})] += parts[2]
            elif k == 6:    # an empty person
                if p != int(rng.integers(0, 3)):
                    c[i, p, x] = d
            else:           # uniform noise
                c[i, p] = rng.integers(0, 12, 4)
    return np.minimum(c, 65535).astype(np.uint16)
