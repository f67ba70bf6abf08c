import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
t = symm.empty(1 << 20, dtype=torch.float32, device="cuda")
h = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr), "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs], "signal_pad_size", getattr(h, "signal_pad_size", None), flush=True)
dist.barrier(); dist.destroy_process_group()
