"""CPU, world_size 2, gloo: the multi-GPU host logic of bench.py — directions are partitioned across ranks, every rank
forms the PARTIAL moment vector of its directions, one all-reduce (sum) gives the global [Tc | Qc] on every rank."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import oracle as orc
    import pbte_b200 as pb
    from cases import host_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hc = host_case("quad_p2")
    P = hc.problem
    full = orc.solve(P, 2, 0.0)  # every rank can afford the tiny full solve: gives identical flux blocks everywhere
    k0, k1 = pb.partition_directions(P.ndir, world, rank)
    # partial moments of the owned directions (what pbte_cuda_sweep leaves in the moment buffer of one GPU)
    sub = pb.ProblemArrays(P.dim, P.ne, P.ndof, P.nfaces, k1 - k0, P.nband, P.omega, P.heat_cap_v,
                           **{f: getattr(P, f) for f in pb.ProblemArrays.FIELDS if f not in ("dir", "dir_w")},
                           dir=P.dir[k0:k1], dir_w=P.dir_w[k0:k1])
    part = orc.moments(sub, full["coeff"][k0:k1])
    buf = torch.from_numpy(np.concatenate([part["Tc"], part["Qc"].reshape(-1)]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    got = buf.numpy()
    want = np.concatenate([full["Tc"], full["Qc"].reshape(-1)])
    err = np.abs(got - want).max() / np.abs(want).max()
    q.put((rank, k0, k1, err))
    dist.destroy_process_group()


def test_partitioned_moments_allreduce_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == 0 and res[0][2] == res[1][1]
    for _, _, _, err in res:
        assert err < 1e-13
