import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 4 * 1024 * 1024
t = symm_mem.empty(n, dtype=torch.float64, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
t.fill_(rank + 1.0)
torch.cuda.synchronize(); dist.barrier()
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (n,), torch.float64)
src = torch.full((n // 2,), 10.0 + rank, dtype=torch.float64, device=dev)
# push into the second half of the peer's buffer, then signal
for it in range(3):
    pb[n // 2:].copy_(src)
    hdl.put_signal(peer, channel=0)
    hdl.wait_signal((rank - 1) % world, channel=0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); e0.record()
for it in range(20):
    pb[n // 2:].copy_(src)
    hdl.put_signal(peer, channel=0)
    hdl.wait_signal((rank - 1) % world, channel=0)
e1.record(); torch.cuda.synchronize()
print(rank, "ok", float(t[0]), float(t[-1]), "ms per push+signal of", n // 2 * 8 / 1e6, "MB:", e0.elapsed_time(e1) / 20, flush=True)
dist.barrier(); dist.destroy_process_group()
