"""One process per GPU (-m gpu, needs two devices): pbte_cuda_comm_get_unique_id / comm_attach across processes — the exchange buffers
are shared as CUDA IPC handles — and pbte_cuda_solve with the device-side stop test, against one process on one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pbte_b200 as pb
from cases import host_case

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_processes_equal_one_device(tmp_path, peer):
    if pb.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    name = "c2_hex_p2_small"
    hc = host_case(name, face_rule=1)  # consistent traces: well-conditioned local systems, the two summation orders agree to rounding
    one = pb.Solver(hc.problem, device=0)
    one.build_schedule()
    it1, res1, hist1 = one.solve(6, 0.0)
    assert hist1[3] < hist1[:3].min()
    tol = float(0.5 * (hist1[3] + hist1[:3].min()))  # stops at the fourth iteration
    one.reset_state()
    it1, res1, hist1 = one.solve(10, tol)
    want = one.download_moments()
    one.close()
    hc.close()
    assert it1 == 4
    env = dict(os.environ, PBTE_PEER_ALLREDUCE=peer)
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "mp_worker.py"), str(r), "2", str(tmp_path), name, "10", repr(tol), "1"], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=120)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out.decode(errors="replace"))
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    got = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(2)]
    for g in got:
        assert int(g["it"]) == it1 and abs(float(g["res"]) - res1) <= 1e-10 * res1
        for k in ("Tc", "Qc", "Tv", "Qv"):
            assert rel(g[k], want[k]) <= 1e-12, (k, rel(g[k], want[k]))
    for k in ("Tc", "Qc", "Tv", "Qv"):
        assert np.array_equal(got[0][k], got[1][k]), k  # every rank ends with the same bits
