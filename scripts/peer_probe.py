"""Probe: can two torchrun ranks map each other's CUDA buffers (CUDA IPC via torch storage sharing)?"""
import os, time, traceback
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1 << 30
send = torch.full((n,), rank + 1, dtype=torch.uint8, device=f"cuda:{local}")
recv = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local}")
try:
    handle = send.untyped_storage()._share_cuda_()
    print(rank, "handle fields", len(handle), [type(x).__name__ for x in handle], flush=True)
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, (handle, int(send.storage_offset()), int(send.numel())))
    peer = {}
    for r, (h, off, m) in enumerate(gathered):
        if r == rank:
            continue
        st = torch.UntypedStorage._new_shared_cuda(*h)
        peer[r] = torch.empty(0, dtype=torch.uint8, device=st.device).set_(st, off, (m,))
    dist.barrier()
    partner = rank ^ 1
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        recv.copy_(peer[partner], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(rank, "pull 1 GiB", f"{dt*1e3:.2f} ms", f"{n/dt/1e9:.0f} GB/s", "value", int(recv[12345]), flush=True)
    dist.barrier()
except Exception:
    traceback.print_exc()
dist.destroy_process_group()
