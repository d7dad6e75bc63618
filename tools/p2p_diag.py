"""p2p_diag.py - what connects the GPUs of this box: topology, peer-copy bandwidth, NCCL all-to-all bandwidth (run under torchrun)."""
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    n = torch.cuda.device_count()
    for a in range(n):
        for b in range(n):
            if a != b:
                print(f"can_access_peer {a}->{b}:", torch.cuda.can_device_access_peer(a, b))
    if n > 1:
        x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0"); y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:1")
        for _ in range(2):
            torch.cuda.synchronize(0); torch.cuda.synchronize(1); t0 = time.time()
            for _ in range(4):
                y.copy_(x, non_blocking=True)
            torch.cuda.synchronize(0); torch.cuda.synchronize(1)
            print("peer copy 0->1: %.1f GB/s" % (4 * (1 << 30) / (time.time() - t0) / 1e9))
        del x, y
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for mb in (64, 1024, 4096):
        n = mb << 20
        a = torch.empty(n // 8, dtype=torch.int64, device="cuda"); b = torch.empty_like(a)
        for _ in range(2):
            dist.all_to_all_single(b, a)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_to_all_single(b, a)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if rank == 0:
            print(f"all_to_all_single {mb} MB per rank: {ms:.2f} ms = {n * (world - 1) / world / ms / 1e6:.1f} GB/s sent per rank")
    dist.destroy_process_group()
