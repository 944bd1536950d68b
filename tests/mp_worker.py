"""Worker of tests/test_gpu_two_processes.py: one process per GPU, as torchrun / mpirun would start them.
usage: mp_worker.py <rank> <world> <scratch dir> <case name> <max_iter> <tol> <face rule>
Rank 0 creates the communicator id and ships it through a file (the launcher's job: MPI_Bcast / torch.distributed.broadcast in a real
run); every rank owns a contiguous direction block on device `rank`, attaches, runs pbte_cuda_solve and stores what it ends with."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import pbte_b200 as pb
from cases import host_case

rank, world, scratch, name, max_iter, tol = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]), float(sys.argv[6])
hc = host_case(name, face_rule=int(sys.argv[7]))
P = hc.problem
s = pb.Solver(P, device=rank, dir_range=pb.partition_directions(P.ndir, world, rank))
s.build_schedule()
id_file = os.path.join(scratch, "comm_id.bin")
if rank == 0:
    with open(id_file + ".tmp", "wb") as f:
        f.write(pb.comm_unique_id())
    os.replace(id_file + ".tmp", id_file)
t0 = time.time()
while not os.path.exists(id_file):
    if time.time() - t0 > 120:
        raise RuntimeError("no communicator id from rank 0")
    time.sleep(0.01)
with open(id_file, "rb") as f:
    uid = f.read()
s.comm_attach(uid, rank, world)
print(f'rank {rank}: attached', flush=True)
it, res, hist = s.solve(max_iter, tol)
print(f'rank {rank}: solve returned after {it} iterations, residual {res}', flush=True)
m = s.download_moments()
np.savez(os.path.join(scratch, f"rank{rank}.npz"), it=it, res=res, hist=hist, Tc=m["Tc"], Qc=m["Qc"], Tv=m["Tv"], Qv=m["Qv"])
print(f'rank {rank}: results stored', flush=True)
s.comm_detach()
print(f'rank {rank}: detached', flush=True)
s.close()
hc.close()
