"""Worker of tests/test_parallel.py::test_peer_memory_allreduce_adam_two_gpus (launched under torchrun, one process per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_mjx_lab_b200 import ppo  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
n, split = 297003, 200000
g = torch.Generator(device=dev).manual_seed(0)
p0 = torch.randn(n, device=dev, generator=g)
comm = ppo._PeerComm(n, dev)
pa = p0.clone()
opt = ppo._FlatAdam(pa, comm.grad, split, 3e-4, 1e-3, comm=comm)
pb1, pb2 = p0[:split].clone().requires_grad_(), p0[split:].clone().requires_grad_()
o1, o2 = torch.optim.Adam([pb1], lr=3e-4, eps=1e-8), torch.optim.Adam([pb2], lr=1e-3, eps=1e-8)
gr = torch.Generator(device=dev).manual_seed(100 + rank)          # different gradients on every rank
graph, static_g = None, torch.zeros(n, device=dev)
for t in range(12):
    grad = torch.randn(n, device=dev, generator=gr) * (1.0 + t)
    ref = grad.clone()
    dist.all_reduce(ref)
    ref /= dist.get_world_size()
    pb1.grad, pb2.grad = ref[:split].clone(), ref[split:].clone()
    o1.step(); o2.step()
    if t < 6:                                                      # eager launches
        comm.grad.copy_(grad)
        opt.step()
    else:                                                          # replayed from a CUDA graph
        static_g.copy_(grad)
        if graph is None:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                comm.grad.copy_(static_g)
                opt.step()
        graph.replay()
torch.cuda.synchronize()
assert comm.error() == 0
torch.testing.assert_close(pa[:split], pb1.detach(), rtol=1e-5, atol=2e-6)
torch.testing.assert_close(pa[split:], pb2.detach(), rtol=1e-5, atol=2e-6)
gathered = [torch.empty_like(pa) for _ in range(dist.get_world_size())]
dist.all_gather(gathered, pa)
assert all(torch.equal(gathered[0], x) for x in gathered), "replicas diverged"
assert float(opt.step_dev) == 12.0
if rank == 0:
    print("peer comm ok")
dist.destroy_process_group()
