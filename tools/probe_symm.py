import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.int64, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hex(hdl.multicast_ptr) if hdl.multicast_ptr else 0, "signal pads", len(hdl.signal_pad_ptrs), flush=True)
    t.zero_(); hdl.barrier()
    # peer write test: every rank writes its rank+1 into element [rank] of every peer via a tensor view
    for p in range(world):
        buf = hdl.get_buffer(p, (1 << 20,), torch.int64)
        buf[rank] = rank + 1
    hdl.barrier()
    print(rank, "after peer writes", t[:world].tolist(), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
dist.destroy_process_group()
