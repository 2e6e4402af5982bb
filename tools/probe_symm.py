"""Probe: does torch symmetric memory (NVLink peer memory + NVSwitch multicast) work on this box?  torchrun --nproc-per-node 2."""
import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    t = sm.empty(1 << 20, dtype=torch.float32, device=dev)
    h = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "backend", sm.get_backend(dev) if hasattr(sm, "get_backend") else None, "multicast", h.has_multicast_support, hex(h.multicast_ptr) if h.has_multicast_support else None,
          "bufs", [hex(p) for p in h.buffer_ptrs], "sigpad", h.signal_pad_size, flush=True)
    t.fill_(rank + 1)
    h.barrier(channel=0)
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer read", peer[:2].tolist(), flush=True)
    h.barrier(channel=0)
except Exception as ex:
    print(rank, "FAILED", repr(ex), flush=True)
dist.destroy_process_group()
