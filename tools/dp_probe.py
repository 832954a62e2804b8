"""Times the fused exchange kernels ALONE (nvqa_dp_rmsprop_step back to back, no backward in between) on N ranks:
what the NVLink reduce-scatter + RMSprop + all-gather of the whole flat vector costs when the ranks are in lockstep.
torchrun --nproc-per-node N tools/dp_probe.py   (prints one line per rank 0)"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import novel_vqa_b200 as nv          # noqa: E402  (alias package of novel-vqa_b200/, as bench.py imports it)
from novel_vqa_b200 import dp        # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cfg = nv.Arch1Config(B=500)
    model = nv.Arch1Model(cfg, device=local)
    enc, emb, mm = nv.synth_params(cfg, seed=123)
    for blk, w in ((nv.BLOCK_ENCODER, enc), (nv.BLOCK_EMBEDDING, emb), (nv.BLOCK_MULTIMODAL, mm)):
        model.set_params(blk, w)
    q, ln, fc7, lab = nv.synth_batch(cfg, 500, seed=123 + rank)
    stream = torch.cuda.Stream(device=local)
    nv._lib.check(model.lib.nvqa_set_stream(model.handle, ctypes.c_void_p(stream.cuda_stream)))
    dp.connect_fused(model, dist, rank, world)
    dq, dl, df, dy = (nv.DeviceBuffer(model, a) for a in (q, ln, fc7, lab))
    model.set_batch_device(dq, dl, df, dy, 500)
    with torch.cuda.stream(stream):
        for i in range(3):
            dp.fused_train_step(model, 3e-4, 10 + i)          # real gradients in place
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    lib = model.lib
    res = {}
    for label, iters in (("warm", 10), ("timed", 50)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(iters):
                nv._lib.check(lib.nvqa_dp_rmsprop_step(model.handle, 1e-6, 0.99, 1e-8, 0.0, 10.0))
            e1.record(stream)
        torch.cuda.synchronize()
        res[label] = e0.elapsed_time(e1) / iters * 1e3
    t = torch.tensor([res["timed"]], device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        nbytes = 4 * sum(model.param_count(b) for b in (nv.BLOCK_ENCODER, nv.BLOCK_EMBEDDING, nv.BLOCK_MULTIMODAL))
        us = float(t.item())
        print(f"dp_probe world={world} ctas={os.environ.get('NVQA_DP_MAIN_CTAS', 'default')}: {us:.1f} us per whole-vector exchange "
              f"({nbytes / 1e6:.1f} MB; per rank {nbytes * (world - 1) / world / 1e6:.1f} MB read + the same written over NVLink "
              f"= {nbytes * (world - 1) / world / us / 1e3:.0f} GB/s each way)", flush=True)
    model.sync()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
