"""(run by scripts/gpu_multi.sh) Times the accumulator exchange alone (64 MiB float32 grid of C3): this library's peer kernel at several CTA counts
vs torch.distributed.all_reduce (NCCL), under torchrun."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from raynet_b200 import sharding
dev = torch.device("cuda", local)
n = 256 ** 3
out = {"world": world}

def timed(fn, reps=40):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

x = torch.zeros(n, device=dev)
out["nccl_all_reduce_ms"] = timed(lambda: dist.all_reduce(x))
for ctas, mc in ((148, True), (96, True), (64, True), (32, True), (148, False), (96, False)):
    ex = sharding.PeerExchange(n, dev, n_ctas=ctas, multicast=mc)
    if mc and not ex.multicast:
        out["multicast"] = "unavailable"
        del ex
        continue
    out["peer_%s_%d_ctas_ms" % ("multicast" if mc else "p2p", ctas)] = timed(lambda: ex.allreduce(-2.9))
    # correctness: every partial holds rank + 1 -> result = prior + world (world + 1) / 2
    ex.partial.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
    r = ex.allreduce(0.5); torch.cuda.synchronize()
    assert float((r - (0.5 + world * (world + 1) / 2)).abs().max()) == 0.0
    del ex
if rank == 0:
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
