"""Probe (run under torchrun on >= 2 GPUs): does torch's symmetric memory rendezvous work on this box, does it expose
peer pointers / a multicast (NVLS) pointer, and what does NCCL's all-reduce of the 56 MB bucket cost."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 14_000_000
x = torch.randn(n, device=dev)
for _ in range(5):
    dist.all_reduce(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
e0.record()
for _ in range(20):
    dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"nccl all_reduce 56MB x{world}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(n + 4096, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print("symm ok: world", hdl.world_size, "rank", hdl.rank, "multicast_ptr", hex(hdl.multicast_ptr),
              "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_size", hdl.signal_pad_size,
              "has_multicast", getattr(hdl, "has_multicast_support", None), flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(f"rank {rank} reads peer value {float(peer[0])}", flush=True)
    hdl.barrier()
    for name in (("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_") if world < 8 else ()):
        try:
            op = getattr(torch.ops.symm_mem, name)
            for _ in range(3):
                op(t, "sum", dist.group.WORLD.group_name)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(20):
                op(t, "sum", dist.group.WORLD.group_name)
            e1.record()
            torch.cuda.synchronize()
            if rank == 0:
                print(f"torch symm_mem {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
        except Exception as ex:
            if rank == 0:
                print(f"torch symm_mem {name} failed: {str(ex)[:200]}", flush=True)
except Exception as ex:
    print(f"rank {rank}: symmetric memory unavailable: {repr(ex)[:400]}", flush=True)
try:
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
        sys.path.insert(0, p)
    from gsr_mapstep import PeerExchange
    sweep = [("peer", 64), ("nvls", 64)] if world < 8 else [("nvls", b) for b in (32, 48, 64, 96)]
    for mode, blocks in sweep:
        os.environ["GSR_AR_BLOCKS"] = str(blocks)
        ex = PeerExchange(n, dev, None, mode)
        ex.flat.copy_(x)
        for _ in range(5):
            ex.all_reduce()
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(20):
            ex.all_reduce()
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"library exchange mode={mode} blocks={blocks}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
        del ex
except Exception as ex_:
    print(f"rank {rank}: library exchange failed: {repr(ex_)[:400]}", flush=True)
dist.barrier()
dist.destroy_process_group()
