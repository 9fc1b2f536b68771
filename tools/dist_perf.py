"""Multi-GPU timing probe (under torchrun): the bench.py unit (C2 grown N-fold, rank-local generation, peer-to-peer plane), V-cycles
replayed as a graph, device time per cycle (max over ranks).  MGCFD_DIST_DEBUG switches parts of the protocol off (results wrong,
timing only).   usage: dist_perf.py [cycles [workload]]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
import mgcfd_b200 as M

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
dist.init_process_group("gloo")
torch.cuda.set_device(local)
cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 100
wl = sys.argv[2] if len(sys.argv) > 2 else "c2"
kind, dims, variant, _ = bench.WORKLOADS[wl]
gd = [[d[0] * world - (world - 1), d[1], d[2]] for d in dims]
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt = torch.frombuffer(bytearray(M.dist_unique_id()), dtype=torch.uint8).clone()
dist.broadcast(idt, 0)
s = M.Solver.generate_distributed(kind, gd, rank, world, bytes(idt.numpy().tobytes()), mesh_variant=variant, lengths=(float(world), 1.0, 1.0), device=local)
mine = s.p2p_prepare()
allp = [None] * world
dist.all_gather_object(allp, mine)
s.p2p_attach([a[0] for a in allp], [a[1] for a in allp])
try:
    s.run_cycles(5)
except M.MgcfdError:
    pass
stream = torch.cuda.ExternalStream(s.cuda_stream(), device=local)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
with torch.cuda.stream(stream):
    e0.record(stream); s.enqueue_cycles(cycles); e1.record(stream)
try:
    s.collect()
except M.MgcfdError:
    pass
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / cycles], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"ranks": world, "dbg": os.environ.get("MGCFD_DIST_DEBUG", "0"), "visit": os.environ.get("MGCFD_VISIT", "0"), "workload": wl, "ms_per_cycle": float(t)}), flush=True)
dist.destroy_process_group()
