"""f3: the NAM-style daemon and its CUDA-free client (byte-stream and shared-memory transports)."""
import fcntl
import mmap
import os
import socket
import struct
import subprocess
import threading

import numpy as np
import pytest

from falcon_genome_b200 import synth
from falcon_genome_b200.remote import NAM_PATH, NamDaemon, RemotePairHMM


def test_daemon_refuses_to_start_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([NAM_PATH, str(tmp_path / "s.sock")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr  # 3 = the reference's exit code for a missing accelerator


def test_client_fails_loudly_without_daemon(tmp_path):
    from falcon_genome_b200 import PairHMMError

    with pytest.raises(PairHMMError) as e:
        RemotePairHMM(str(tmp_path / "nobody.sock"))
    assert e.value.code == -2


@pytest.mark.gpu
def test_daemon_serves_clients_and_stops_on_sigalrm(tmp_path, hmm):
    sock = str(tmp_path / "nam.sock")
    b1, b2 = synth.tiny_mixed(seed=71, n_regions=6), synth.config1_golden(n_regions=4, seed=72)
    ref1, u1 = hmm.compute_flat(b1)
    ref2, u2 = hmm.compute_flat(b2)
    with NamDaemon(sock, devices=1) as nam:
        with RemotePairHMM(sock) as c1, RemotePairHMM(sock) as c2:  # two client processes' worth of connections
            for _ in range(3):
                o1, f1 = c1.compute_flat(b1)
                o2, f2 = c2.compute_flat(b2)
                assert np.array_equal(o1, ref1) and np.array_equal(f1, u1)
                assert np.array_equal(o2, ref2) and np.array_equal(f2, u2)
            assert c1.uses_shm and c2.uses_shm  # the default transport: one sealed memfd segment per connection
            big = synth.config1_golden(n_regions=40, seed=73)  # outgrows the first segment: re-attach mid-connection
            ob, fb = c1.compute_flat(big)
            rb, ub = hmm.compute_flat(big)
            assert np.array_equal(ob, rb) and np.array_equal(fb, ub) and c1.uses_shm
        rc = nam.stop()
    assert rc == 0 and not os.path.exists(sock)


@pytest.mark.gpu
def test_client_builds_the_batch_in_the_segment(tmp_path, hmm):
    """fcs_pairhmm_remote_reserve / compute_reserved: the batch is written in place into the connection's segment (no
    client-side staging copy), the results are read in place; same bits as the in-process library, also after the
    segment had to grow and when the reservation is refilled."""
    sock = str(tmp_path / "nam.sock")
    small, big = synth.tiny_mixed(seed=75, n_regions=5), synth.config1_golden(n_regions=30, seed=76)
    with NamDaemon(sock, devices=1):
        with RemotePairHMM(sock) as c:
            for b in (small, big, small):
                ref, uref = hmm.compute_flat(b)
                out, used = c.compute_in_segment(b)
                assert np.array_equal(out, ref) and np.array_equal(used, uref)
            o2, u2 = c.compute_flat(big)  # the copying call still works on the same connection
            rb, ub = hmm.compute_flat(big)
            assert np.array_equal(o2, rb) and np.array_equal(u2, ub)


@pytest.mark.gpu
def test_byte_stream_transport_gives_the_same_results(tmp_path, hmm, monkeypatch):
    sock = str(tmp_path / "nam.sock")
    b = synth.tiny_mixed(seed=74, n_regions=5)
    ref, u = hmm.compute_flat(b)
    with NamDaemon(sock, devices=1):
        monkeypatch.setenv("FCS_PHMM_REMOTE_SHM", "0")
        with RemotePairHMM(sock) as c:
            o, f = c.compute_flat(b)
            assert not c.uses_shm
        assert np.array_equal(o, ref) and np.array_equal(f, u)


# ---- the shared-memory framing, spoken by hand ------------------------------------------------------

PHSM, PHSQ, PHRQ, PHRS, FSHM = 0x4D534850, 0x51534850, 0x51524850, 0x53524850, 0x4D485346
HDR = struct.Struct("<IIqqqQQQ" + "Q" * 17)  # ShmHeader of csrc/phmm_shm.h (192 bytes)
HDR_FIELDS = ["magic", "version", "n_regions", "n_reads", "n_haps", "n_pairs", "read_bytes", "hap_bytes", "off_read_bases",
              "off_read_q", "off_read_i", "off_read_d", "off_read_c", "off_rd_off", "off_rd_len", "off_hap_bases", "off_hp_off",
              "off_hp_len", "off_reg_read0", "off_reg_nreads", "off_reg_hap0", "off_reg_nhaps", "off_out", "off_used", "total_bytes"]


def _recv_exact(conn, n, fds=None):
    buf = b""
    while len(buf) < n:
        if fds is not None:
            chunk, got, _, _ = socket.recv_fds(conn, n - len(buf), 4)
            fds.extend(got)
        else:
            chunk = conn.recv(n - len(buf))
        if not chunk:
            raise EOFError
        buf += chunk
    return buf


class _FakeDaemon(threading.Thread):
    """Speaks the daemon's side of the protocol without a GPU: checks that the segment describes exactly the
    batch the client was given and answers with recognisable numbers."""

    def __init__(self, path, batch, accept_segment=True):
        super().__init__(daemon=True)
        self.batch, self.accept_segment = batch, accept_segment
        self.srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.srv.bind(path)
        self.srv.listen(1)
        self.seen = []
        self.error = None

    def run(self):
        try:
            conn, _ = self.srv.accept()
            seg = None
            while True:
                fds = []
                try:
                    tag, length = struct.unpack("<IQ", _recv_exact(conn, 12, fds))
                except EOFError:
                    break
                self.seen.append(tag)
                if tag == PHSM:
                    assert len(fds) == 1
                    if not self.accept_segment:
                        os.close(fds[0])
                        msg = b"no segments today"
                        conn.sendall(struct.pack("<IiQ", PHRS, -1, len(msg)) + msg)
                        continue
                    assert fcntl.fcntl(fds[0], fcntl.F_GET_SEALS) & fcntl.F_SEAL_SHRINK
                    assert os.fstat(fds[0]).st_size >= length
                    seg = mmap.mmap(fds[0], length)
                    os.close(fds[0])
                    conn.sendall(struct.pack("<IiQ", PHRS, 0, 0))
                elif tag == PHSQ:
                    h = dict(zip(HDR_FIELDS, HDR.unpack_from(seg, 0)))
                    b = self.batch
                    assert h["magic"] == FSHM and h["version"] == 1
                    assert (h["n_regions"], h["n_reads"], h["n_haps"], h["n_pairs"]) == (b.n_regions, b.n_reads, b.n_haps, b.n_pairs)
                    assert h["total_bytes"] <= len(seg) and all(h[k] % 64 == 0 for k in HDR_FIELDS if k.startswith("off_"))
                    rd_off = np.frombuffer(seg, np.int64, b.n_reads, h["off_rd_off"])
                    rd_len = np.frombuffer(seg, np.int32, b.n_reads, h["off_rd_len"])
                    assert np.array_equal(rd_len, b.rd_len)
                    for name, plane in (("off_read_bases", b.read_bases), ("off_read_q", b.read_q), ("off_read_i", b.read_i),
                                        ("off_read_d", b.read_d), ("off_read_c", b.read_c)):
                        for k in (0, b.n_reads // 2, b.n_reads - 1):
                            got = np.frombuffer(seg, np.uint8, int(rd_len[k]), h[name] + int(rd_off[k]))
                            assert np.array_equal(got, plane[b.rd_off[k]:b.rd_off[k] + b.rd_len[k]])
                    hp_off = np.frombuffer(seg, np.int64, b.n_haps, h["off_hp_off"])
                    hp_len = np.frombuffer(seg, np.int32, b.n_haps, h["off_hp_len"])
                    assert np.array_equal(hp_len, b.hp_len)
                    k = b.n_haps - 1
                    assert np.array_equal(np.frombuffer(seg, np.uint8, int(hp_len[k]), h["off_hap_bases"] + int(hp_off[k])),
                                          b.hap_bases[b.hp_off[k]:b.hp_off[k] + b.hp_len[k]])
                    assert np.array_equal(np.frombuffer(seg, np.int32, b.n_regions, h["off_reg_nreads"]), b.reg_nreads)
                    assert np.array_equal(np.frombuffer(seg, np.int32, b.n_regions, h["off_reg_nhaps"]), b.reg_nhaps)
                    np.frombuffer(seg, np.float64, b.n_pairs, h["off_out"])[:] = -np.arange(b.n_pairs)
                    np.frombuffer(seg, np.uint8, b.n_pairs, h["off_used"])[:] = np.arange(b.n_pairs) % 2
                    del rd_off, rd_len, hp_off, hp_len
                    conn.sendall(struct.pack("<IiQ", PHRS, 0, b.n_pairs))
                elif tag == PHRQ:
                    _recv_exact(conn, length)
                    msg = b"byte stream seen"
                    conn.sendall(struct.pack("<IiQ", PHRS, -1, len(msg)) + msg)
                else:
                    raise AssertionError(hex(tag))
            conn.close()
        except Exception as e:  # surfaced by the test
            self.error = e
        finally:
            self.srv.close()


def test_client_writes_the_batch_into_a_sealed_segment(tmp_path):
    """Host-only: the client's side of the shared-memory transport against a hand-written peer."""
    b = synth.tiny_mixed(seed=75, n_regions=4)
    sock = str(tmp_path / "fake.sock")
    d = _FakeDaemon(sock, b)
    d.start()
    with RemotePairHMM(sock) as c:
        for _ in range(2):  # the second call reuses the attached segment
            out, used = c.compute_flat(b)
            assert np.array_equal(out, -np.arange(b.n_pairs)) and np.array_equal(used, np.arange(b.n_pairs) % 2)
        assert c.uses_shm
    d.join(timeout=10)
    assert d.error is None, d.error
    assert d.seen == [PHSM, PHSQ, PHSQ]


def test_client_falls_back_to_the_byte_stream_when_the_segment_is_declined(tmp_path):
    from falcon_genome_b200 import PairHMMError

    b = synth.tiny_mixed(seed=76, n_regions=2)
    sock = str(tmp_path / "fake.sock")
    d = _FakeDaemon(sock, b, accept_segment=False)
    d.start()
    with RemotePairHMM(sock) as c:
        with pytest.raises(PairHMMError) as e:
            c.compute_flat(b)
        assert "byte stream seen" in str(e.value) and not c.uses_shm
    d.join(timeout=10)
    assert d.error is None, d.error
    assert d.seen == [PHSM, PHRQ]


def _attach(conn, fd, size):
    socket.send_fds(conn, [struct.pack("<IQ", PHSM, size)], [fd])
    tag, rc, n = struct.unpack("<IiQ", _recv_exact(conn, 16))
    assert tag == PHRS
    return rc, _recv_exact(conn, n).decode() if rc != 0 and n else ""


def _ring(conn):
    conn.sendall(struct.pack("<IQ", PHSQ, 0))
    tag, rc, n = struct.unpack("<IiQ", _recv_exact(conn, 16))
    assert tag == PHRS
    return rc, _recv_exact(conn, n).decode() if rc != 0 and n else ""


@pytest.mark.gpu
def test_daemon_rejects_malformed_segments_and_keeps_serving(tmp_path, hmm):
    sock = str(tmp_path / "nam.sock")
    b = synth.tiny_mixed(seed=77, n_regions=3)
    ref, u = hmm.compute_flat(b)
    size = 1 << 20
    with NamDaemon(sock, devices=1) as nam:
        conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        conn.connect(sock)
        rc, msg = _ring(conn)  # doorbell before any segment
        assert rc < 0 and "no segment" in msg
        fd = os.memfd_create("unsealed", os.MFD_CLOEXEC | os.MFD_ALLOW_SEALING)
        os.ftruncate(fd, size)
        rc, msg = _attach(conn, fd, size)  # not sealed against shrinking: SIGBUS bait
        assert rc < 0 and "sealed" in msg
        fcntl.fcntl(fd, fcntl.F_ADD_SEALS, fcntl.F_SEAL_SHRINK | fcntl.F_SEAL_GROW)
        rc, msg = _attach(conn, fd, size * 2)  # announces more than the file holds
        assert rc < 0 and "smaller" in msg
        rc, msg = _attach(conn, fd, size)
        assert rc == 0
        seg = mmap.mmap(fd, size)
        rc, msg = _ring(conn)  # all zeros
        assert rc < 0 and "header" in msg
        good = dict.fromkeys(HDR_FIELDS, 0)
        good.update(magic=FSHM, version=1, n_regions=1, n_reads=1, n_haps=1, n_pairs=1, read_bytes=4, hap_bytes=4, total_bytes=4096)
        for i, k in enumerate(f for f in HDR_FIELDS if f.startswith("off_")):
            good[k] = 256 + 64 * i
        cases = [
            ({"off_read_q": size - 2}, "outside the mapping"),
            ({"off_rd_off": 257}, "outside the mapping"),  # misaligned
            ({"n_reads": 1 << 40}, "out of range"),
            ({"n_pairs": 2}, "pair count"),
        ]
        def put(h, rd_len=4, hp_len=4, nreads=1):
            seg[:HDR.size] = HDR.pack(*[h[k] for k in HDR_FIELDS])
            struct.pack_into("<q", seg, good["off_rd_off"], 0)
            struct.pack_into("<i", seg, good["off_rd_len"], rd_len)
            struct.pack_into("<q", seg, good["off_hp_off"], 0)
            struct.pack_into("<i", seg, good["off_hp_len"], hp_len)
            struct.pack_into("<i", seg, good["off_reg_read0"], 0)
            struct.pack_into("<i", seg, good["off_reg_nreads"], nreads)
            struct.pack_into("<i", seg, good["off_reg_hap0"], 0)
            struct.pack_into("<i", seg, good["off_reg_nhaps"], 1)
        for patch, text in cases:
            put({**good, **patch})
            rc, msg = _ring(conn)
            assert rc < 0 and text in msg, (patch, msg)
        put(good, rd_len=5)
        rc, msg = _ring(conn)
        assert rc < 0 and "read outside" in msg
        put(good, hp_len=1 << 30)
        rc, msg = _ring(conn)
        assert rc < 0 and "haplotype outside" in msg
        put(good, nreads=2)
        rc, msg = _ring(conn)
        assert rc < 0 and "region outside" in msg
        # a well-formed one-pair batch through the same hand-made segment
        put(good)
        for k in ("off_read_bases", "off_hap_bases"):
            seg[good[k]:good[k] + 4] = b"ACGT"
        for k, q in (("off_read_q", 30), ("off_read_i", 45), ("off_read_d", 45), ("off_read_c", 10)):
            seg[good[k]:good[k] + 4] = bytes([q] * 4)
        rc, msg = _ring(conn)
        assert rc == 0, msg
        val = struct.unpack_from("<d", seg, good["off_out"])[0]
        assert -1.0 < val < 0.0
        seg.close()
        os.close(fd)
        conn.close()
        with RemotePairHMM(sock) as c:  # the daemon is still healthy
            o, f = c.compute_flat(b)
        assert np.array_equal(o, ref) and np.array_equal(f, u)
        assert nam.stop() == 0
