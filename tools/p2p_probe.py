"""Peer-copy bandwidth between GPU 0 and GPU 1 of the box (copy engine, one direction and both at once): the yardstick
for the fused exchange kernels of csrc/dp_fused.cu."""
import torch

assert torch.cuda.device_count() >= 2
print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
n = 64 << 20   # 256 MB of fp32
a0 = torch.empty(n, dtype=torch.float32, device="cuda:0")
b1 = torch.empty(n, dtype=torch.float32, device="cuda:1")
a1 = torch.empty(n, dtype=torch.float32, device="cuda:1")
b0 = torch.empty(n, dtype=torch.float32, device="cuda:0")
for size in (n, 7 << 20):     # 256 MB and 28 MB (one rank's share of the exchange at 2 ranks)
    for both in (False, True):
        s0 = torch.cuda.Stream(device=0)
        s1 = torch.cuda.Stream(device=1)
        for _ in range(3):
            with torch.cuda.stream(s0):
                b1[:size].copy_(a0[:size], non_blocking=True)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        with torch.cuda.stream(s0):
            e0.record(s0)
            for _ in range(iters):
                b1[:size].copy_(a0[:size], non_blocking=True)
            e1.record(s0)
        if both:
            torch.cuda.set_device(1)
            with torch.cuda.stream(s1):
                for _ in range(iters):
                    b0[:size].copy_(a1[:size], non_blocking=True)
            torch.cuda.set_device(0)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        ms = e0.elapsed_time(e1) / iters
        print(f"peer copy 0->1 of {size * 4 / 1e6:.0f} MB{' with 1->0 running' if both else ''}: {ms * 1e3:.1f} us = {size * 4 / ms / 1e6:.0f} GB/s")
