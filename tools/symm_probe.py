import os, torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1024, dtype=torch.int64, device=dev)
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "world", hdl.world_size, flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.int64)
print(rank, "peer value", int(peer[0].item()), flush=True)
peer[5] = 100 + rank
hdl.barrier()
torch.cuda.synchronize()
print(rank, "my slot 5", int(t[5].item()), flush=True)
dist.destroy_process_group()
