"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): |dlog10 L| <= 1e-4 per pair vs the double-precision oracle,
identical FP32->FP64 fallback decisions, identical per-read best haplotype.  Stronger, by
construction of the kernels: raw FP32 sums bit-identical to the oracle's float twin and FP64
results equal to the double oracle to 1e-12.
"""
import numpy as np
import pytest

from falcon_genome_b200 import FlatBatch, PairHMM, Region, synth

pytestmark = pytest.mark.gpu

TOL = 1e-4  # |delta log10 L| per pair, north_star


def check_against_oracle(oracle, b, out, used, raw=None, simd=True):
    if simd:
        o_ref, u_ref, r_ref, _ = oracle.batch_simd(b)  # bit-identical to the scalar twin (tests/test_oracle.py)
        dbl = None
    else:
        o_ref, u_ref, r_ref, dbl = oracle.batch_scalar(b)
    assert np.array_equal(used, u_ref), f"fallback decisions differ on {(used != u_ref).sum()} of {len(used)} pairs"
    if raw is not None:
        assert np.array_equal(raw.view(np.uint32), r_ref.view(np.uint32)), "raw FP32 sums are not bit-identical to the float twin"
    fin = np.isfinite(o_ref)
    assert np.array_equal(fin, np.isfinite(out))
    err = np.abs(out[fin] - o_ref[fin])
    assert err.max() <= TOL, f"max |dlog10L| = {err.max()}"
    # FP64 path: same operation order as the oracle -> agreement far below the tolerance
    sel = (used == 1) & fin
    if sel.any():
        assert np.abs(out[sel] - o_ref[sel]).max() <= 1e-9
    if dbl is not None:
        assert np.abs(out[fin] - dbl[fin]).max() <= TOL  # vs the double-precision oracle for EVERY pair
    # per-read best haplotype (ties -> lowest h)
    for g in range(b.n_regions):
        nr, nh, o0 = int(b.reg_nreads[g]), int(b.reg_nhaps[g]), int(b.reg_out0[g])
        if nr and nh:
            a = out[o0:o0 + nr * nh].reshape(nr, nh)
            r = o_ref[o0:o0 + nr * nh].reshape(nr, nh)
            assert np.array_equal(a.argmax(1), r.argmax(1))
    return float(err.max())


def test_kat_single_cell(hmm):
    rd = (b"A", bytes([30]), bytes([45]), bytes([45]), bytes([10]))
    m = hmm.compute_likelihoods([rd], [b"A", b"C", b"AAAAA", b"N"])
    assert abs(m[0, 0] - np.log10(0.999 * 0.9)) < 5e-6
    assert abs(m[0, 1] - np.log10(0.001 / 3 * 0.9)) < 5e-6
    assert abs(m[0, 2] - np.log10(0.999 * 0.9)) < 5e-6
    assert abs(m[0, 3] - np.log10(0.999 * 0.9)) < 5e-6


def test_tiny_mixed_all_entry_points(hmm, oracle):
    b = synth.tiny_mixed()
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)
    out2, used2 = hmm.compute_regions(b)
    assert np.array_equal(out, out2) and np.array_equal(used, used2)
    rb = hmm.resident(b)
    rb.run()
    out3, used3, raw3 = rb.download(want_raw=True)
    rb.close()
    assert np.array_equal(out, out3) and np.array_equal(used, used3) and np.array_equal(raw, raw3)


@pytest.mark.parametrize("seed", [7, 8, 9, 10])
def test_ragged_seeds(hmm, oracle, seed):
    b = synth.tiny_mixed(seed=seed, n_regions=12)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)


def test_config1_golden_sample(hmm, oracle):
    b = synth.config1_golden(n_regions=24)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw)


def test_config2_sample(hmm, oracle):
    b = synth.config2_uniform(n_regions=6)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw)
    b = synth.config2_uniform(n_regions=3, random_quals=True)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw)


def test_config3_sample(hmm, oracle):
    b = synth.config3_wgs(n_regions=30)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw)


def test_config4_sample(hmm, oracle):
    b = synth.config4_mutect2(n_regions=1)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw)


def test_config5_underflow_fallback(hmm, oracle):
    b = synth.config5_underflow(n_regions=3)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw)
    assert used.mean() > 0.5, "the underflow stress config must mostly take the FP64 path"


def test_force_double_matches_double_oracle(oracle):
    b = synth.tiny_mixed(seed=3, n_regions=8)
    with PairHMM(use_double=True) as h:
        out, used = h.compute_flat(b)
    assert used.all()
    _, _, _, dbl = oracle.batch_scalar(b)
    assert np.abs(out - dbl).max() <= 1e-9


def test_every_read_length_class(hmm, oracle):
    """Read lengths sweep the (G, R) kernel classes, including exact tile fits G*R-1."""
    rng = np.random.default_rng(5)
    regs = []
    lens = [1, 2, 3, 15, 16, 23, 24, 47, 48, 63, 64, 95, 96, 103, 104, 127, 128, 151, 152, 159, 160, 191, 192, 207, 208,
            255, 256, 300, 382, 383]
    hap = bytes(rng.choice(list(b"ACGT"), 333).astype(np.uint8))
    for L in lens:
        s = int(rng.integers(0, max(1, 333 - L)))
        bases = (hap + hap + hap)[s:s + L]
        q = bytes(rng.integers(6, 42, L).astype(np.uint8))
        i = bytes(rng.integers(20, 46, L).astype(np.uint8))
        d = bytes(rng.integers(20, 46, L).astype(np.uint8))
        c = bytes(rng.integers(8, 12, L).astype(np.uint8))
        regs.append(Region([(bases, q, i, d, c)], [hap, hap[:100], hap[50:]]))
    b = FlatBatch.from_regions(regs)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)


def _uniform_indel_region(rng, L, n_reads, hap, ins=45, dele=45, gcp=10, n_rate=0.0):
    reads = []
    for _ in range(n_reads):
        s = int(rng.integers(0, max(1, len(hap) - L)))
        bases = bytearray((hap + hap + hap + hap)[s:s + L])
        for x in range(L):
            u = rng.random()
            if u < 0.02:
                bases[x] = int(rng.choice(list(b"ACGT")))
            elif u < 0.02 + n_rate:
                bases[x] = ord("N")
        q = bytes(rng.integers(6, 42, L).astype(np.uint8))
        reads.append((bytes(bases), q, bytes([ins]) * L, bytes([dele]) * L, bytes([gcp]) * L))
    return reads


def test_all_uniform_form_every_tile(hmm, oracle):
    """Reads whose insertion / deletion / continuation qualities are constant take the all-uniform
    kernels (G=4 up to 151 rows, G=8 up to 303): sweep every rows-per-lane class with full warps,
    exact tile fits, several quality triples in one call, N bases in reads and haplotypes, and
    groups that mix eligible and non-eligible reads (those must fall back to the other forms)."""
    rng = np.random.default_rng(2024)
    hap = bytes(rng.choice(list(b"ACGT"), 341).astype(np.uint8))
    hap_n = bytearray(hap[20:300]); hap_n[57] = ord("N"); hap_n[200] = ord("N")
    regs = []
    lens = [1, 3, 7, 8, 15, 16, 31, 39, 40, 47, 63, 64, 79, 80, 87, 88, 100, 103, 104, 111, 119, 120, 127, 135, 136, 143, 144,
            150, 151, 152, 159, 160, 167, 168, 175, 183, 184, 199, 200, 215, 216, 231, 239, 240, 250, 255, 256, 271, 287, 288,
            295, 296, 302, 303, 304, 320]
    for n, L in enumerate(lens):
        trip = [(45, 45, 10), (40, 40, 10), (30, 30, 12), (20, 45, 3)][n % 4]  # the last one (ins != del) must not take the UA kernels
        reads = _uniform_indel_region(rng, L, 8 if L < 152 else 4, hap, *trip, n_rate=0.01 if n % 5 == 0 else 0.0)
        haps = [hap, hap[:120], hap[33:]] + ([bytes(hap_n)] if n % 3 == 0 else [])
        regs.append(Region(reads, haps))
    # one region with two quality triples (sorted by length they interleave), one with a non-uniform read
    mixed = _uniform_indel_region(rng, 150, 9, hap, 45, 45, 10) + _uniform_indel_region(rng, 150, 9, hap, 44, 45, 10)
    regs.append(Region(mixed, [hap, hap[10:310]]))
    spoiled = _uniform_indel_region(rng, 150, 17, hap)
    b0, q0, i0, d0, c0 = spoiled[5]
    spoiled[5] = (b0, q0, i0[:70] + bytes([30]) + i0[71:], d0, c0)
    b1, q1, i1, d1, c1 = spoiled[11]
    spoiled[11] = (b1, q1, i1, d1[:-1] + bytes([31]), c1)  # differs in the very last byte only
    b2, q2, i2, d2, c2 = spoiled[2]
    spoiled[2] = (b2, q2, i2, d2, c2[:147] + bytes([11]) + c2[148:])
    regs.append(Region(spoiled, [hap, hap[5:250]]))
    b = FlatBatch.from_regions(regs)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)


def test_all_uniform_form_underflow_fallback(hmm, oracle):
    """All-uniform reads that underflow FP32 are queued for the FP64 kernels like any other read."""
    rng = np.random.default_rng(77)
    hap = bytes(rng.choice(list(b"ACGT"), 300).astype(np.uint8))
    other = bytes(rng.choice(list(b"ACGT"), 300).astype(np.uint8))  # unrelated haplotype: tiny likelihoods
    reads = _uniform_indel_region(rng, 250, 16, hap, 45, 45, 40)
    b = FlatBatch.from_regions([Region(reads, [hap, other, other[::-1]])])
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    assert used.any() and not used.all()
    check_against_oracle(oracle, b, out, used, raw, simd=False)


def _pcr_model_region(rng, L, n_reads, hap, gcp=10, same_indel=True):
    """Reads as GATK's PCR indel model leaves them: insertion == deletion quality per position (lower inside
    homopolymer runs), one gap-continuation quality; same_indel=False breaks the equality on some positions."""
    reads = []
    for _ in range(n_reads):
        s = int(rng.integers(0, max(1, len(hap) - L)))
        bases = bytearray((hap + hap + hap + hap)[s:s + L])
        for x in range(L):
            if rng.random() < 0.02:
                bases[x] = int(rng.choice(list(b"ACGTN")))
        q = bytes(rng.integers(6, 42, L).astype(np.uint8))
        i = np.full(L, 45, np.uint8)
        i[rng.random(L) < 0.1] = rng.integers(10, 40)
        d = i.copy()
        if not same_indel:
            d[rng.random(L) < 0.05] = 33
        reads.append((bytes(bases), q, bytes(i), bytes(d), bytes([gcp]) * L))
    return reads


def test_haplotype_pair_kernels(hmm, oracle):
    """Reads with one gap-continuation quality run against their haplotypes two at a time (packed f32x2 kernels):
    sweep the pair classes (exact tile fits included), odd and even haplotype counts, pairs of very different lengths,
    N in haplotypes, reads whose deletion qualities differ from the insertion qualities (four-plane blobs) and groups
    whose reads disagree on the continuation quality (scalar fallback).  Bit-identical raw sums are the bar."""
    from falcon_genome_b200 import plan_check

    rng = np.random.default_rng(31337)
    hap = bytes(rng.choice(list(b"ACGT"), 420).astype(np.uint8))
    hap_n = bytearray(hap[30:330]); hap_n[11] = ord("N"); hap_n[250] = ord("N")
    pool = [hap, hap[:100], hap[50:], bytes(hap_n), hap[200:], hap[5:395], hap[100:160]]
    lens = [40, 64, 95, 96, 103, 104, 127, 128, 150, 151, 152, 159, 160, 175, 191, 192, 200, 239, 250, 255, 256, 300, 319, 320, 383]
    regs = []
    for n, L in enumerate(lens):
        nh = [2, 3, 4, 5, 7][n % 5]
        reads = _pcr_model_region(rng, L, [8, 5, 9, 3][n % 4], hap, same_indel=(n % 3 != 0))
        regs.append(Region(reads, [pool[(n + j) % len(pool)] for j in range(nh)]))
    mixed = _pcr_model_region(rng, 150, 6, hap, gcp=10) + _pcr_model_region(rng, 150, 6, hap, gcp=12)
    regs.append(Region(mixed, [hap, hap[7:300], hap[60:]]))
    # filler: keeps the regions above out of the tail-shaping window (the last ~1.5 waves of a call use other classes)
    for _ in range(220):
        regs.append(Region(_pcr_model_region(rng, 150, 8, hap), [hap[:300], hap[20:310], hap[40:345], hap[3:290]]))
    b = FlatBatch.from_regions(regs)
    assert plan_check(b)["n_tasks_hap_pairs"] > 0
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)
    out2, used2 = hmm.compute_regions(b)  # the chunked path (tail shaping only on the last chunk)
    assert np.array_equal(out, out2) and np.array_equal(used, used2)


def test_golden_fixtures(hmm):
    """Committed fixtures (tools/make_golden.py, oracle-scored): bit-exact raw FP32 sums and fallback
    flags, FP64 reruns to 1e-9, everything within the 1e-4 contract of the double-precision value."""
    from helpers import load_golden, parse_kat

    for name in ("c1_sample.npz", "c2_sample.npz", "c5_sample.npz"):
        b, z = load_golden(name)
        out, used, raw = hmm.compute_flat(b, want_raw=True)
        assert np.array_equal(used, z["used_fp64"])
        assert np.array_equal(raw.view(np.uint32), z["raw_f32_bits"])
        assert np.abs(out - z["log10_double"]).max() <= TOL
        assert np.abs(out - z["out_log10"]).max() <= 4e-6
        if (used == 1).any():
            assert np.abs(out[used == 1] - z["out_log10"][used == 1]).max() <= 1e-9
    for read, hap, exp in parse_kat():
        assert abs(hmm.compute_likelihoods([read], [hap])[0, 0] - exp) < 5e-6


def test_long_reads_and_long_haplotypes_striped_path(hmm, oracle):
    """Shapes beyond the single-pass tiles (reads > 383 rows, haplotypes > 2000 columns) take the
    striped generic kernels: same bit-level contract, including the FP64 fallback."""
    rng = np.random.default_rng(17)

    def seq(n):
        return bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))

    def read_from(h, L, q_lo=6, q_hi=42):
        s0 = int(rng.integers(0, max(1, len(h) - L)))
        b = bytearray((h * (L // len(h) + 2))[s0:s0 + L])
        for k in rng.integers(0, L, max(1, L // 50)):
            b[k] = ord("ACGT"[int(rng.integers(0, 4))])
        return (bytes(b), bytes(rng.integers(q_lo, q_hi, L).astype(np.uint8)), bytes(rng.integers(20, 46, L).astype(np.uint8)),
                bytes(rng.integers(20, 46, L).astype(np.uint8)), bytes(rng.integers(8, 12, L).astype(np.uint8)))

    h1, h2, h3 = seq(900), seq(2500), seq(333)
    regs = [
        Region([read_from(h1, L) for L in (384, 511, 512, 513, 767, 1023, 1024, 1025, 1500)], [h1, h1[:300], h3]),
        Region([(h1[:L], bytes([30] * L), bytes([45] * L), bytes([45] * L), bytes([10] * L)) for L in (256, 512, 768)], [h1]),  # padding fills a stripe exactly; alignment starts at column 1
        Region([read_from(h2, L) for L in (5, 150, 400, 1200)], [h2, h2[100:2300], h3]),   # long haplotypes
        Region([read_from(h3, 150), read_from(h1, 600, 2, 8)], [h3, seq(700)]),               # mixed with a normal read; low quals -> FP64
    ]
    b = FlatBatch.from_regions(regs)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)
    assert used.any() and not used.all()
    with PairHMM(use_double=True) as hd:
        outd, usedd = hd.compute_flat(b)
    _, _, _, dbl = oracle.batch_scalar(b)
    fin = np.isfinite(dbl)  # some pairs underflow even the 2^1020-scaled double: -inf on both sides
    assert usedd.all() and np.array_equal(fin, np.isfinite(outd)) and np.abs(outd[fin] - dbl[fin]).max() <= 1e-9


def test_batching_invariance(hmm):
    """Shuffling regions, splitting a call, or shrinking chunks changes no output bit."""
    b = synth.config1_golden(n_regions=16, seed=11)
    out, used = hmm.compute_flat(b)
    perm = np.random.default_rng(0).permutation(b.n_regions)
    bp = b.select(perm)
    outp, usedp = hmm.compute_flat(bp)
    for k, g in enumerate(perm):
        n = int(b.reg_nreads[g]) * int(b.reg_nhaps[g])
        assert np.array_equal(out[b.reg_out0[g]:b.reg_out0[g] + n], outp[bp.reg_out0[k]:bp.reg_out0[k] + n])
    with PairHMM(max_chunk_cells=2_000_000, slots_per_device=2) as h2:
        out2, used2 = h2.compute_flat(b)
        st = h2.stats()
    assert st["chunks"] > 3
    assert np.array_equal(out, out2) and np.array_equal(used, used2)


def test_submit_wait(hmm):
    from falcon_genome_b200 import RegionArray

    b = synth.tiny_mixed(seed=21)
    ref, _ = hmm.compute_flat(b)
    ra = RegionArray(b)
    t = hmm.submit(ra)
    hmm.wait(t)
    assert np.array_equal(ra.out, ref)
    with pytest.raises(Exception):
        hmm.wait(t)


def test_errors(hmm):
    from falcon_genome_b200 import PairHMMError

    rd = (b"ACGT", bytes([30] * 4), bytes([45] * 4), bytes([45] * 4), bytes([10] * 4))
    with pytest.raises(PairHMMError) as e:
        hmm.compute_likelihoods([rd], [b""])
    assert e.value.code == -1
    # empty regions are fine
    b = FlatBatch.from_regions([Region([], [b"ACGT"]), Region([rd], []), Region([rd], [b"ACGT"])])
    out, used = hmm.compute_flat(b)
    assert out.shape == (1,) and np.isfinite(out[0])


def test_haplotype_bytes_outside_acgtn(hmm, oracle):
    """GKL compares raw bytes, so a haplotype may hold IUPAC codes, lower case or anything else (b37 / hg19
    carry such bases): the byte matches a read base only if the read holds the very same byte or an N.  The
    library must score such input like the oracle's raw-byte rule instead of refusing it -- single-pass
    kernels of every form, FP64 reruns and the striped path; bytes that occur on both sides get their own
    symbol rows (up to 8), the others share one."""
    from falcon_genome_b200 import PairHMMError

    rng = np.random.default_rng(606)

    def seq(n, alpha=b"ACGT"):
        return bytes(rng.choice(list(alpha), n).astype(np.uint8))

    def spoil(h, alpha, rate):
        h = bytearray(h)
        for k in np.nonzero(rng.random(len(h)) < rate)[0]:
            h[k] = int(rng.choice(list(alpha)))
        return bytes(h)

    regs = []
    # (a) haplotype-only foreign bytes (the realistic case): IUPAC + lower case + N, reads clean / with N
    h0 = seq(320)
    haps = [h0, spoil(h0, b"RYKMna", 0.03), spoil(h0[20:300], b"NRY", 0.05), spoil(h0[:150], b"*-.SW", 0.02)]
    regs.append(Region(_uniform_indel_region(rng, 150, 16, h0, n_rate=0.01), haps))          # all-uniform kernels
    regs.append(Region(_uniform_indel_region(rng, 101, 9, h0, 40, 38, 10), haps))            # uniform-GCP kernels
    # (b) the same foreign bytes in reads and haplotypes: equal bytes match (raw compare), different ones do not
    reads = []
    for _ in range(12):
        L = int(rng.integers(30, 260))
        s0 = int(rng.integers(0, 320 - 30))
        b = bytearray((haps[1] * 2)[s0:s0 + L])
        for k in rng.integers(0, L, max(1, L // 15)):
            b[k] = int(rng.choice(list(b"ACGTNRYKa")))
        reads.append((bytes(b), bytes(rng.integers(2, 42, L).astype(np.uint8)), bytes(rng.integers(10, 50, L).astype(np.uint8)),
                      bytes(rng.integers(10, 50, L).astype(np.uint8)), bytes(rng.integers(5, 30, L).astype(np.uint8))))
    regs.append(Region(reads, haps))                                                          # general form
    # (c) low qualities + unrelated foreign-byte haplotype: FP64 reruns see the same table
    lowq = [(seq(250), bytes(rng.integers(2, 8, 250).astype(np.uint8)), bytes([12] * 250), bytes([12] * 250), bytes([10] * 250)) for _ in range(6)]
    regs.append(Region(lowq, [spoil(seq(900), b"RYN", 0.05), spoil(seq(700), b"MK", 0.02)]))
    # (d) striped path: long read and long haplotype with foreign bytes on both sides
    hl = spoil(seq(2400), b"RYNn", 0.02)
    rl = bytearray(hl[100:700]); rl[17] = ord("R"); rl[300] = ord("y"); rl[301] = ord("N")
    regs.append(Region([(bytes(rl), bytes([25] * 600), bytes([40] * 600), bytes([40] * 600), bytes([10] * 600)),
                        (hl[5:155], bytes([30] * 150), bytes([45] * 150), bytes([45] * 150), bytes([10] * 150))], [hl, hl[50:2100]]))
    b = FlatBatch.from_regions(regs)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)
    assert used.any() and not used.all()
    # region by region (other chunk alphabets: a region alone may need fewer symbol rows) -> same bits
    for g in range(b.n_regions):
        bg = b.select([g])
        og, ug = hmm.compute_flat(bg)
        o0 = int(b.reg_out0[g])
        assert np.array_equal(og, out[o0:o0 + len(og)]) and np.array_equal(ug, used[o0:o0 + len(og)])
    # a known answer: 'R' in the haplotype matches a read 'R' and a read 'N', nothing else
    rd = lambda s: (s, bytes([30]), bytes([45]), bytes([45]), bytes([10]))  # noqa: E731
    m = np.array([hmm.compute_likelihoods([rd(x)], [b"R"])[0, 0] for x in (b"R", b"N", b"A", b"Y")])
    assert np.allclose(m[:2], np.log10(0.999 * 0.9), atol=5e-6) and np.allclose(m[2:], np.log10(0.001 / 3 * 0.9), atol=5e-6)
    # more than 8 distinct foreign byte values shared by reads and haplotypes of one chunk: refused, not mis-scored
    many = bytes(range(ord("a"), ord("a") + 12))
    with pytest.raises(PairHMMError) as e:
        hmm.compute_likelihoods([(many, bytes([30] * 12), bytes([45] * 12), bytes([45] * 12), bytes([10] * 12))], [many])
    assert e.value.code == -5


def test_finalize_epilogue_on_device(oracle):
    """f2 fused into the device pipeline: with set_finalize() every read's row is capped at best - 4.5 on the GPU and
    the poorly-modelled flag comes back per read.  Checked against (1) a restatement written here from GATK's
    description with numpy whole-matrix operations on the ORACLE's likelihoods, (2) the library's host function
    fcs_pairhmm_finalize_region applied to the unfinalized GPU result (bit-equal), through the streamed path with
    small chunks, the region API and a resident batch."""
    from falcon_genome_b200 import RegionArray
    from falcon_genome_b200.prepost import finalize_region

    b = synth.config1_golden(n_regions=30, seed=77)
    lowq = synth.config5_underflow(n_regions=2, seed=78)  # reads that are poorly modelled by every haplotype
    for bb in (b, lowq):
        with PairHMM(max_chunk_cells=20_000_000, slots_per_device=2) as h:
            plain, used0 = h.compute_flat(bb)
            h.set_finalize(True)
            out, used, poorly = h.compute_flat_finalized(bb)
            ra = RegionArray(bb)
            h.compute_regions(bb, ra)
            rb = h.resident(bb)
            rb.run()
            out_res, _ = rb.download()
            rb.close()
            h.set_finalize(False)
            again, _ = h.compute_flat(bb)
        assert np.array_equal(again, plain) and np.array_equal(used, used0)
        assert np.array_equal(ra.out, out) and np.array_equal(out_res, out)
        o_ref, _, _, _ = oracle.batch_simd(bb)
        flags_ref = np.zeros(bb.n_reads, np.uint8)
        for g in range(bb.n_regions):
            nr, nh, o0, r0 = int(bb.reg_nreads[g]), int(bb.reg_nhaps[g]), int(bb.reg_out0[g]), int(bb.reg_read0[g])
            m = plain[o0:o0 + nr * nh].reshape(nr, nh)
            lens = bb.rd_len[r0:r0 + nr]
            host, hflags = finalize_region(m, lens)
            assert np.array_equal(out[o0:o0 + nr * nh].reshape(nr, nh), host)           # device epilogue == host function, bit for bit
            assert np.array_equal(poorly[r0:r0 + nr], hflags)
            # restatement from the description, on the oracle's matrix: L'[r,h] = max(L[r,h], max_h L[r,:] - 4.5)
            mo = o_ref[o0:o0 + nr * nh].reshape(nr, nh)
            best = mo.max(axis=1, keepdims=True)
            assert np.abs(np.maximum(mo, best - 4.5) - out[o0:o0 + nr * nh].reshape(nr, nh)).max() <= 1e-4
            flags_ref[r0:r0 + nr] = best[:, 0] < np.minimum(2.0, np.ceil(lens * 0.02)) * -4.0
        near = np.zeros(bb.n_reads, bool)  # reads whose best likelihood sits within the tolerance of the flag threshold may differ
        assert np.array_equal(poorly[~near], flags_ref[~near])
    assert poorly.any()  # the low-quality batch has poorly modelled reads


def _full_size_parity(hmm, oracle, b):
    """The parity bars at benchmark size: decisions and raw FP32 sums against the float twin (bit level), every pair
    against the DOUBLE-precision oracle (1e-4), per-read best haplotype, bit-reproducibility."""
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    err = check_against_oracle(oracle, b, out, used, raw)      # float twin (SIMD port, IEEE subnormals) + arg-max
    dbl = oracle.batch_double(b)                                # the north_star's reference arithmetic, every pair
    fin = np.isfinite(dbl)
    assert np.array_equal(fin, np.isfinite(out)) and np.abs(out[fin] - dbl[fin]).max() <= TOL
    out2, used2 = hmm.compute_flat(b)
    assert np.array_equal(out, out2) and np.array_equal(used, used2)
    return out, used, err


def test_full_size_config2(hmm, oracle):
    """BASELINE config 2 at full size: 100 000 pairs, 150 x 300, uniform quals (the bench default)."""
    b = synth.config2_uniform()
    out, used, _ = _full_size_parity(hmm, oracle, b)
    assert b.n_pairs == 100_000 and used.sum() <= 5  # reads drawn from one haplotype rarely underflow against another
    m = out.reshape(100, 100, 10)
    assert (m.max(axis=2) > -45).all() and (m <= 0).all()


def test_full_size_config3_chunk(hmm, oracle):
    """One 2000-region chunk of the config-3 stream (405 k pairs, 24 Gcells, ragged lengths, ~9 % FP64 reruns)."""
    b = synth.config3_wgs(n_regions=2000, seed=3003, chunk=0)
    _, used, _ = _full_size_parity(hmm, oracle, b)
    assert b.n_regions == 2000 and 0.02 < used.mean() < 0.3


def test_full_size_config4_sample(hmm, oracle):
    """Twenty Mutect2-shaped regions (415 k pairs, 17 Gcells, deep pileups, per-position indel qualities)."""
    b = synth.config4_mutect2(n_regions=20, seed=4004)
    _full_size_parity(hmm, oracle, b)


def test_full_size_config5(hmm, oracle):
    """All 20 000 pairs of the underflow stress config: nearly every pair takes the FP64 path; results against the
    double oracle to 1e-9."""
    b = synth.config5_underflow()
    out, used, _ = _full_size_parity(hmm, oracle, b)
    assert b.n_pairs == 20_000 and used.mean() > 0.9
    dbl = oracle.batch_double(b)
    sel = (used == 1) & np.isfinite(dbl)
    assert np.abs(out[sel] - dbl[sel]).max() <= 1e-9


def test_in_process_multi_gpu_dispatch(oracle):
    """Engine::compute with every visible device: regions partitioned by cells, results gathered by
    index == single-device results, bit for bit (needs >= 2 GPUs; skipped otherwise)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    b = synth.config1_golden(n_regions=40, seed=21)
    with PairHMM(devices=[0]) as h1:
        o1, u1 = h1.compute_flat(b)
    with PairHMM(max_chunk_cells=50_000_000) as hn:
        assert hn.device_count >= 2
        on, un = hn.compute_flat(b)
        assert hn.stats()["chunks"] >= 2
    assert np.array_equal(o1, on) and np.array_equal(u1, un)


def test_concurrent_callers_on_one_handle(hmm):
    """GATK drives the native library from several threads (--native-pair-hmm-threads,
    /root/reference/src/workers/HTCWorker.cpp:85): concurrent compute() calls on one handle must be
    safe and give the single-threaded results (ctypes releases the GIL during the call)."""
    import threading

    batches = [synth.tiny_mixed(seed=50 + i, n_regions=5) for i in range(6)] + [synth.config1_golden(n_regions=6, seed=60)]
    ref = [hmm.compute_flat(b)[0] for b in batches]
    errs = []

    def work(tid):
        try:
            for rep in range(4):
                for k, b in enumerate(batches):
                    if (k + rep + tid) % 2 == 0:
                        out, _ = hmm.compute_flat(b)
                    else:
                        out, _ = hmm.compute_regions(b)
                    if not np.array_equal(out, ref[k]):
                        errs.append((tid, rep, k))
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[:3]


def test_pipelined_batches_and_a_failing_caller(hmm):
    """Concurrent multi-chunk calls: a batch's front phase (plan + pack + launch) overlaps the previous batch's tail on the
    device, chunks are retired by whoever needs their slot next.  Results must equal the single-threaded ones; a caller
    whose input the batcher refuses gets its error every time, alone (the merged batch is re-run call by call), and nobody hangs."""
    import threading

    from falcon_genome_b200 import PairHMMError

    batches = [synth.config3_wgs(n_regions=150, seed=900 + i) for i in range(4)] + [synth.config1_golden(n_regions=40, seed=77)]
    ref = [hmm.compute_flat(b) for b in batches]
    many = bytes(range(ord("a"), ord("a") + 12))  # > 8 foreign byte values shared by read and haplotype: FCS_PHMM_EUNSUPPORTED
    bad = FlatBatch.from_regions([Region([(many, bytes([30] * 12), bytes([45] * 12), bytes([45] * 12), bytes([10] * 12))], [many])])
    errs, refused = [], [0]

    def good(tid):
        try:
            for rep in range(5):
                for k in range(len(batches)):
                    kk = (k + tid) % len(batches)
                    out, used = hmm.compute_regions(batches[kk]) if (rep + tid) % 2 else hmm.compute_flat(batches[kk])
                    if not (np.array_equal(out, ref[kk][0]) and np.array_equal(used, ref[kk][1])):
                        errs.append((tid, rep, kk))
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    def faulty():
        for _ in range(25):
            try:
                hmm.compute_flat(bad)
                errs.append("the refused call went through")
            except PairHMMError as e:
                if e.code != -5:
                    errs.append(("code", e.code))
                refused[0] += 1

    th = [threading.Thread(target=good, args=(t,)) for t in range(5)] + [threading.Thread(target=faulty)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=240)
    assert not any(t.is_alive() for t in th), "a caller hangs"
    assert not errs, errs[:3]
    assert refused[0] == 25


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_fuzz_random_shapes_and_bytes(hmm, oracle, seed):
    """Randomised regions: any read length 1..420, haplotype length 1..700, quals over the whole byte
    range (the kernels mask & 127), N and non-ACGT read bytes, uniform and per-base gap-continuation
    quals mixed in one call (general and uniform-GCP kernels side by side)."""
    rng = np.random.default_rng(seed)
    regs = []
    for _ in range(40):
        nh = int(rng.integers(1, 6))
        alpha, pr = (b"ACGTN", [.24, .24, .24, .24, .04]) if rng.random() < 0.7 else (b"ACGTNRYa", [.23, .23, .23, .23, .04, .02, .01, .01])
        haps = [bytes(rng.choice(list(alpha), int(rng.integers(1, 700)), p=pr).astype(np.uint8)) for _ in range(nh)]
        reads = []
        for _r in range(int(rng.integers(1, 12))):
            L = int(rng.choice([rng.integers(1, 30), rng.integers(30, 200), rng.integers(200, 420)]))
            src = haps[int(rng.integers(0, nh))]
            b = bytearray((src * (L // len(src) + 2))[:L])
            for k in rng.integers(0, L, max(1, L // 20)):
                b[k] = int(rng.choice(list(b"ACGTNRYacgt")))
            q = rng.integers(0, 256, L).astype(np.uint8) if rng.random() < 0.3 else rng.integers(2, 42, L).astype(np.uint8)
            i = rng.integers(5, 60, L).astype(np.uint8)
            d = rng.integers(5, 60, L).astype(np.uint8)
            c = np.full(L, int(rng.integers(5, 40)), np.uint8) if rng.random() < 0.6 else rng.integers(5, 40, L).astype(np.uint8)
            reads.append((bytes(b), bytes(q), bytes(i), bytes(d), bytes(c)))
        regs.append(Region(reads, haps))
    b = FlatBatch.from_regions(regs)
    out, used, raw = hmm.compute_flat(b, want_raw=True)
    check_against_oracle(oracle, b, out, used, raw, simd=False)
