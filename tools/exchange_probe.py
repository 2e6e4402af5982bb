#!/usr/bin/env python3
"""Times the pieces of the multi-GPU exchange on their own (torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/exchange_probe.py [P W H]
rendezvous, ts2d_exchange_tiles (7 planes), ts2d_exchange_allreduce (contrib statistics; 16 P accumulators), the clone of the frame,
and the NCCL all-reduces they replace.  CUDA events on the launching stream, median of 20 after 5 warm-ups, max over ranks."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from triangle_splatting_b200 import _lib, distributed as tsd  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    P, W, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1_500_000, 1920, 1080)
    lib = _lib.load()
    tsd.enable_tile_sharding()
    fab = tsd.fabric(dev)
    assert fab is not None, "no multicast mapping"
    n_pix = W * H
    pad4 = lambda n: (n + 3) // 4 * 4
    o_sum = pad4(7 * n_pix)
    frame, mc, h = fab.buffer("frame", o_sum + 2 * pad4(P))
    acc, acc_mc, acc_h = fab.buffer("accumulators", 16 * P)
    frame.normal_(), acc.normal_()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    chunk = ((pad4(P) + world - 1) // world + 3) // 4 * 4
    first = min(rank * chunk, pad4(P))
    count = min(chunk, pad4(P) - first)
    achunk = ((16 * P + world - 1) // world + 3) // 4 * 4
    afirst = min(rank * achunk, 16 * P)
    acount = min(achunk, 16 * P - afirst)
    nccl_frame = torch.zeros(7 * n_pix + P, device=dev)
    nccl_max = torch.zeros(P, device=dev)
    nccl_acc = torch.zeros(16 * P, device=dev)
    pieces = {
        "rendezvous": lambda: h.barrier(channel=0),
        "exchange_tiles 7 planes": lambda: _lib.check(lib.ts2d_exchange_tiles(C.c_void_p(frame.data_ptr()), C.c_void_p(mc), 7, W, H, rank, world, stream), "x"),
        "allreduce contrib_sum (f32 add)": lambda: _lib.check(lib.ts2d_exchange_allreduce(C.c_void_p(mc), o_sum + first, count, 0, stream), "x"),
        "allreduce contrib_max (u32 max)": lambda: _lib.check(lib.ts2d_exchange_allreduce(C.c_void_p(mc), o_sum + pad4(P) + first, count, 1, stream), "x"),
        "allreduce accumulators (16 P f32)": lambda: _lib.check(lib.ts2d_exchange_allreduce(C.c_void_p(acc_mc), afirst, acount, 0, stream), "x"),
        "frame.clone()": lambda: frame.clone(),
        "NCCL all-reduce frame + contrib_sum": lambda: dist.all_reduce(nccl_frame),
        "NCCL all-reduce contrib_max": lambda: dist.all_reduce(nccl_max, op=dist.ReduceOp.MAX),
        "NCCL all-reduce accumulators": lambda: dist.all_reduce(nccl_acc),
    }
    for name, fn in pieces.items():
        ts = []
        for it in range(25):
            h.barrier(channel=0)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            if it >= 5:
                ts.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name:40s} {t.item() * 1e3:8.1f} us", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
