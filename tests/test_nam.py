"""f3: the NAM-style daemon and its CUDA-free client."""
import os
import subprocess

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
        rc = nam.stop()
    assert rc == 0 and not os.path.exists(sock)
