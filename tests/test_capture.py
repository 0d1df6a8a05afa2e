"""Capture / replay formats: CPU round trips; the GPU test captures at the ABI."""
import numpy as np
import pytest

from falcon_genome_b200 import FlatBatch, load_capture, read_gkl_text, save_capture, synth, write_gkl_text
from helpers import GOLD


def same_batch(a: FlatBatch, b: FlatBatch):
    assert a.n_regions == b.n_regions
    for g in range(a.n_regions):
        ra, rb = a.region(g), b.region(g)
        assert ra.reads == rb.reads and ra.haps == rb.haps, g


def test_capture_roundtrip_through_library_loader(tmp_path):
    b = synth.tiny_mixed(seed=31, n_regions=7)
    p = str(tmp_path / "a.fcsphmm")
    save_capture(b, p)
    same_batch(b, load_capture(p))
    save_capture(b.select([2, 0]), p, append=True)  # a second block appends regions
    c = load_capture(p)
    assert c.n_regions == 9 and c.n_pairs == b.n_pairs + b.select([2, 0]).n_pairs
    assert np.array_equal(c.reg_out0, np.concatenate([[0], np.cumsum(c.reg_nreads.astype(np.int64) * c.reg_nhaps)])[:-1])


def test_capture_loader_rejects_garbage(tmp_path):
    from falcon_genome_b200 import PairHMMError

    p = tmp_path / "bad.bin"
    p.write_bytes(b"not a capture")
    with pytest.raises(PairHMMError):
        load_capture(str(p))
    q = tmp_path / "trunc.fcsphmm"
    save_capture(synth.tiny_mixed(seed=1, n_regions=2), str(q))
    q.write_bytes(q.read_bytes()[:-7])
    with pytest.raises(PairHMMError):
        load_capture(str(q))


def test_gkl_text_roundtrip(tmp_path, oracle):
    b, exp = read_gkl_text(GOLD + "/kat_closed_form.txt")
    assert b.n_pairs == len(exp) == 6 and np.isfinite(exp).all()
    out, _, _, _ = oracle.batch_scalar(b)
    assert np.abs(out - exp).max() < 5e-6
    b2 = synth.tiny_mixed(seed=4, n_regions=3)
    ref, _, _, _ = oracle.batch_scalar(b2)
    p = str(tmp_path / "t.txt")
    write_gkl_text(b2, p, ref)
    b3, exp3 = read_gkl_text(p)
    assert b3.n_pairs == b2.n_pairs
    out3, _, _, _ = oracle.batch_scalar(b3)
    assert np.abs(out3 - exp3).max() < 1e-9


@pytest.mark.gpu
def test_capture_at_the_abi_and_replay(tmp_path, hmm):
    b = synth.tiny_mixed(seed=41, n_regions=6)
    p = str(tmp_path / "live.fcsphmm")
    hmm.set_capture(p)
    try:
        out, used = hmm.compute_flat(b)
        out_r, _ = hmm.compute_regions(b.select([1, 3]))
    finally:
        hmm.set_capture(None)
    c = load_capture(p)
    assert c.n_regions == 8
    same_batch(b, c.select(range(6)))
    out2, used2 = hmm.compute_flat(c)
    assert np.array_equal(out2[: b.n_pairs], out) and np.array_equal(out2[b.n_pairs:], out_r)
