"""CPU tests of the host-side logic: flat batches, deterministic generators, region partition
over ranks (with a world_size-2 gloo run of the shard -> compute -> gather path)."""
import os
import subprocess
import sys

import numpy as np

from falcon_genome_b200 import FlatBatch, Region, partition_regions, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flatbatch_roundtrip():
    b = synth.tiny_mixed(seed=3)
    regs = [b.region(g) for g in range(b.n_regions)]
    b2 = FlatBatch.from_regions(regs)
    for f in ("read_bases", "read_q", "read_i", "read_d", "read_c", "rd_off", "rd_len", "hap_bases", "hp_off", "hp_len",
              "reg_read0", "reg_nreads", "reg_hap0", "reg_nhaps", "reg_out0"):
        assert np.array_equal(getattr(b, f), getattr(b2, f)), f
    assert b.n_pairs == sum(len(r.reads) * len(r.haps) for r in regs)
    assert b.cells == sum(sum(len(x[0]) for x in r.reads) * sum(len(h) for h in r.haps) for r in regs)


def test_generators_are_deterministic_and_shaped():
    a, b = synth.config2_uniform(n_regions=3), synth.config2_uniform(n_regions=3)
    assert np.array_equal(a.read_bases, b.read_bases) and np.array_equal(a.hap_bases, b.hap_bases)
    assert a.n_pairs == 3000 and set(a.rd_len) == {150} and set(a.hp_len) == {300}
    assert set(a.read_q) == {30} and set(a.read_i) == {45} and set(a.read_c) == {10}
    c1 = synth.config1_golden(n_regions=20)
    assert c1.rd_len.max() == 150 and c1.rd_len.min() >= 60 and 100 <= c1.hp_len.min() and c1.hp_len.max() <= 370
    assert (c1.read_bases == ord("N")).mean() < 0.03
    c5 = synth.config5_underflow(n_regions=2)
    assert set(c5.rd_len) == {250} and set(c5.hp_len) == {1000} and c5.read_q.max() <= 15
    full = synth.config2_uniform()
    assert full.n_pairs == 100_000 and full.cells == 4_500_000_000


def test_partition_regions_balanced_and_complete():
    b = synth.config1_golden(n_regions=60, seed=3)
    cells = b.region_cells()
    for world in (1, 2, 4, 8):
        parts = partition_regions(cells, world)
        allr = np.sort(np.concatenate(parts))
        assert np.array_equal(allr, np.arange(b.n_regions))
        loads = np.array([cells[p].sum() for p in parts], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.15
    assert all(np.array_equal(x, y) for x, y in zip(partition_regions(cells, 4), partition_regions(cells, 4)))


def test_select_rebuilds_dense_output_layout():
    b = synth.tiny_mixed(seed=5, n_regions=8)
    s = b.select([6, 1, 3])
    assert s.n_regions == 3 and s.reg_out0[0] == 0
    assert s.n_pairs == sum(int(b.reg_nreads[g]) * int(b.reg_nhaps[g]) for g in (6, 1, 3))
    assert s.region(0).haps == b.region(6).haps


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import _pkg; _pkg.load()
from falcon_genome_b200 import synth, partition_regions
from oracle import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
b = synth.tiny_mixed(seed=13, n_regions=9)
parts = partition_regions(b.region_cells(), world)
mine = b.select(parts[rank])
out, used, _, _ = O.batch_simd(mine, 1)      # stand-in for the per-GPU library call
gathered = [None] * world
dist.all_gather_object(gathered, (parts[rank].tolist(), out.tolist(), used.tolist()))
if rank == 0:
    full = np.full(b.n_pairs, np.nan); fu = np.zeros(b.n_pairs, np.uint8)
    for ids, o, u in gathered:
        sub = b.select(ids); o = np.asarray(o); u = np.asarray(u, dtype=np.uint8)
        for k, g in enumerate(ids):
            n = int(b.reg_nreads[g]) * int(b.reg_nhaps[g])
            full[b.reg_out0[g]:b.reg_out0[g]+n] = o[sub.reg_out0[k]:sub.reg_out0[k]+n]
            fu[b.reg_out0[g]:b.reg_out0[g]+n] = u[sub.reg_out0[k]:sub.reg_out0[k]+n]
    ref, ru, _, _ = O.batch_simd(b, 1)
    assert np.array_equal(full, ref) and np.array_equal(fu, ru)
    print("GLOO-OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_world_size_2_shard_compute_gather(tmp_path):
    """N>1 host path on CPU: shard regions over 2 ranks (gloo), score each shard independently,
    gather by region index == single-rank result.  No data-path collective exists."""
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(w), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO-OK" in r.stdout


def test_batcher_covers_every_pair_exactly_once():
    """fcs_pairhmm_plan_check runs the real planner + packer on the host and verifies that the FP32 tasks and
    the striped-path pair list cover every (read, hap) pair exactly once with classes that fit the reads."""
    from falcon_genome_b200 import plan_check

    rng = np.random.default_rng(9)
    regs = []
    for _ in range(60):
        nh = int(rng.integers(1, 9))
        haps = [bytes(rng.choice(list(b"ACGTN"), int(rng.choice([rng.integers(1, 50), rng.integers(50, 700), rng.integers(1900, 2300)]))).astype(np.uint8))
                for _ in range(nh)]
        reads = []
        for _r in range(int(rng.integers(1, 40))):
            L = int(rng.choice([rng.integers(1, 30), rng.integers(30, 260), rng.integers(260, 900)]))
            reads.append((bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8)), bytes([30] * L), bytes([45] * L), bytes([45] * L),
                          bytes([10] * L) if rng.random() < 0.7 else bytes(rng.integers(5, 30, L).astype(np.uint8))))
        regs.append(Region(reads, haps))
    b = FlatBatch.from_regions(regs)
    info = plan_check(b)
    assert info["n_pairs"] == b.n_pairs and info["n_generic_pairs"] > 0 and info["n_tasks"] > 0
    assert info["max_smem_bytes"] <= 227 * 1024 and info["n_sym"] == 6
    # (tail shaping trades geometric efficiency for a short last wave on the final ~24k pairs, so measure
    # the efficiency on batches well above that)
    for mk, n_regions, lo in ((synth.config2_uniform, 100, 0.95), (synth.config1_golden, 400, 0.90)):
        bb = mk(n_regions=n_regions)
        i2 = plan_check(bb)
        assert i2["n_pairs"] == bb.n_pairs and i2["n_generic_pairs"] == 0 and i2["geometric_efficiency"] >= lo and i2["n_sym"] == 5
    tiny = plan_check(synth.tiny_mixed(seed=2, n_regions=2))
    assert tiny["latency_mode"] == 1  # an under-filled call switches to the widest lane groups


def test_planner_picks_the_kernel_form_from_the_qualities():
    """Host-only (fcs_pairhmm_plan_check): constant insertion == deletion and continuation qualities -> all-uniform
    kernels for the full warps of reads; anything else -> uniform-GCP or general form; a chunk in which the eligible
    reads are a minority does not use the all-uniform kernels at all."""
    from falcon_genome_b200 import plan_check

    rng = np.random.default_rng(3)
    hap = bytes(rng.choice(list(b"ACGT"), 300).astype(np.uint8))

    def region(n_reads, L, ins, dele, gcp, spoil=None):
        reads = []
        for r in range(n_reads):
            i = bytearray([ins] * L)
            c = bytearray([gcp] * L)
            if spoil == "ins" and r % 2 == 0:
                i[L // 2] = ins - 5
            if spoil == "gcp":
                c[r % L] = gcp + 1
            reads.append((hap[:L], bytes([30] * L), bytes(i), bytes([dele] * L), bytes(c)))
        return Region(reads, [hap, hap[10:290], hap[5:280], hap[20:300]])

    def forms(regs):
        i = plan_check(FlatBatch.from_regions(regs))
        assert i["n_tasks"] == i["n_tasks_general"] + i["n_tasks_uniform_gcp"] + i["n_tasks_all_uniform"] + i["n_tasks_hap_pairs"]
        # (the haplotype-pair kernels are the throughput variant of the uniform-GCP form: counted with it)
        return i["n_tasks_general"], i["n_tasks_uniform_gcp"] + i["n_tasks_hap_pairs"], i["n_tasks_all_uniform"]

    big = 400  # regions: well above the tail-shaping window, which uses wide uniform-GCP classes
    g, u, a = forms([region(32, 150, 45, 45, 10) for _ in range(big)])
    assert g == 0 and a >= big  # (u > 0 too: the last half wave of a batch runs on wide uniform-GCP classes)
    g, u, a = forms([region(16, 150, 45, 40, 10) for _ in range(big)])  # ins != del: the shared product does not apply
    assert g == 0 and a == 0 and u > 0
    g, u, a = forms([region(16, 150, 45, 45, 10, spoil="gcp") for _ in range(big)])  # per-read GCP not constant
    assert u == 0 and a == 0 and g > 0
    g, u, a = forms([region(16, 150, 45, 45, 10, spoil="ins") for _ in range(big)])  # half of the reads eligible, interleaved
    assert g == 0 and u > 0  # groups that mix eligible and non-eligible reads fall back
    minority = [region(16, 150, 45, 45, 10) for _ in range(big // 4)] + [region(16, 150, 45, 40, 10) for _ in range(big)]
    g, u, a = forms(minority)
    assert a == 0 and u > 0


def test_compact_read_layouts_and_haplotype_pairs():
    """Host-only: the packer drops planes that carry no information (constant gap-continuation quality -> trailer;
    deletion plane == insertion plane -> not copied; all three constant -> two planes), and reads with one
    continuation quality take the haplotype-pair kernels for an even number of haplotypes, the scalar form for the odd one."""
    from falcon_genome_b200 import plan_check

    rng = np.random.default_rng(11)
    hap = bytes(rng.choice(list(b"ACGT"), 300).astype(np.uint8))
    L, n_reg, n_reads = 160, 300, 16  # L = 160: planes need no padding

    def batch(kind, haps):
        regs = []
        for _ in range(n_reg):
            reads = []
            for r in range(n_reads):
                i = bytearray([45] * L)
                d = bytearray([45] * L)
                c = bytearray([10] * L)
                if kind >= 1:
                    i[r + 3] = 30; d[r + 3] = 30  # per-position, insertion == deletion
                if kind >= 2:
                    d[r + 9] = 31                 # deletion plane differs
                if kind >= 3:
                    c[r + 1] = 11                 # continuation quality not constant
                reads.append((hap[:L], bytes([30] * L), bytes(i), bytes(d), bytes(c)))
            regs.append(Region(reads, haps))
        return plan_check(FlatBatch.from_regions(regs))

    one = [batch(k, [hap]) for k in range(4)]  # one haplotype per region: no pair kernels, same task count in every form
    n = n_reg * n_reads
    assert one[1]["n_tasks"] == one[2]["n_tasks"] == one[3]["n_tasks"]
    assert one[2]["in_bytes"] - one[1]["in_bytes"] == n * L          # + deletion plane
    assert one[3]["in_bytes"] - one[2]["in_bytes"] == n * (L - 16)   # + continuation plane - trailer
    assert one[1]["in_bytes"] > one[0]["in_bytes"]                   # + insertion plane (the all-uniform form cuts other tasks)
    assert all(o["n_tasks_hap_pairs"] == 0 for o in one)
    five = batch(1, [hap, hap[10:290], hap[5:280], hap[20:300], hap[1:299]])
    assert five["n_tasks_hap_pairs"] > 0 and five["n_tasks_uniform_gcp"] > 0 and five["n_tasks_general"] == 0
    four = batch(2, [hap, hap[10:290], hap[5:280], hap[20:300]])
    assert four["n_tasks_hap_pairs"] > 0
    assert batch(3, [hap, hap[10:290]])["n_tasks_hap_pairs"] == 0  # per-read continuation quality not constant


# ---- property test of the batcher (hypothesis): any ragged call shape is covered exactly once ----------
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

_read_len = st.one_of(st.integers(1, 24), st.integers(25, 260), st.integers(261, 420), st.integers(384, 1200))
_hap_len = st.one_of(st.integers(1, 40), st.integers(41, 700), st.integers(1990, 2100))


@st.composite
def _call_shape(draw):
    regions = []
    for _ in range(draw(st.integers(1, 12))):
        reads = draw(st.lists(st.tuples(_read_len, st.sampled_from(["ua", "ug", "general", "ins_ne_del"])), min_size=1, max_size=70))
        haps = draw(st.lists(st.tuples(_hap_len, st.booleans()), min_size=1, max_size=9))
        regions.append((reads, haps))
    return regions


@settings(max_examples=40, deadline=None)
@given(_call_shape(), st.integers(0, 2 ** 31 - 1))
def test_property_batcher_covers_any_call_shape(shape, seed):
    """fcs_pairhmm_plan_check (the real planner + packer, host only) accepts any mix of read / haplotype lengths and
    quality forms -- single-pass classes, the striped path for long reads and long haplotypes, leftover groups,
    latency mode -- and its own coverage check finds every (read, hap) pair exactly once."""
    from falcon_genome_b200 import plan_check

    rng = np.random.default_rng(seed)
    regs = []
    for reads, haps in shape:
        rr = []
        for L, form in reads:
            ins = bytearray([45] * L)
            dele = bytearray([45] * L)
            gcp = bytearray([10] * L)
            if form in ("ug", "general") and L > 1:
                ins[int(rng.integers(0, L))] = 30
                dele = bytearray(ins)
            if form == "general" and L > 1:
                gcp[int(rng.integers(0, L))] = 11
            if form == "ins_ne_del":
                dele = bytearray([40] * L)
            rr.append((bytes(rng.choice(list(b"ACGTN"), L).astype(np.uint8)), bytes(rng.integers(2, 42, L).astype(np.uint8)), bytes(ins), bytes(dele),
                       bytes(gcp)))
        hh = [bytes(rng.choice(list(b"ACGTN" if with_n else b"ACGT"), L).astype(np.uint8)) for L, with_n in haps]
        regs.append(Region(rr, hh))
    b = FlatBatch.from_regions(regs)
    info = plan_check(b)
    assert info["n_pairs"] == b.n_pairs
    assert info["n_tasks"] + info["n_generic_pairs"] > 0
    assert info["n_tasks"] == info["n_tasks_general"] + info["n_tasks_uniform_gcp"] + info["n_tasks_all_uniform"] + info["n_tasks_hap_pairs"]
    assert info["max_smem_bytes"] <= 227 * 1024
    assert 0.0 < info["geometric_efficiency"] <= 1.0
