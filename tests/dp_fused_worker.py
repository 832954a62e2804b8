"""Worker of tests/test_dp_fused_gpu.py: launched with torchrun, one rank per GPU.  Each rank runs a few fused
data-parallel steps (csrc/dp_fused.cu) on its own shard and checks, against the host restatement of
"sum over ranks / N -> clamp -> RMSprop" (oracle.arch1.rmsprop_update) fed with every rank's raw gradients, that
(1) the parameters after each step match, (2) all replicas are bit-identical."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import novel_vqa_b200 as nv
    from novel_vqa_b200 import dp
    from oracle import arch1 as A

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    cfg = nv.Arch1Config() if full else nv.Arch1Config(V=300, E=24, H=64, L=2, I=96, C=64, O=50, T=8, B=40)
    B = cfg.B
    m = nv.Arch1Model(cfg, precision=nv.PREC_BF16X2, device=local)
    blocks = (nv.BLOCK_ENCODER, nv.BLOCK_EMBEDDING, nv.BLOCK_MULTIMODAL)
    params = list(nv.synth_params(cfg, seed=1))                       # identical replicas
    for blk, w in zip(blocks, params):
        m.set_params(blk, w)
    rms = [np.zeros_like(w) for w in params]
    dp.connect_fused(m, dist, rank, world)
    lr = 3e-4
    ok = True
    for step in range(3):
        q, ln, fc7, lab = nv.synth_batch(cfg, B, seed=100 * step + rank, min_len=2)   # per-rank shard
        m.set_batch_host(q, ln, fc7, lab)
        dp.fused_train_step(m, lr, seed=7 + step)
        m.sync()
        mine = [m.get_grads(b) for b in blocks]
        allg = [None] * world
        dist.all_gather_object(allg, mine)
        for k, blk in enumerate(blocks):
            g = dp.average_then_update_reference([allg[r][k] for r in range(world)])
            A.rmsprop_update(params[k], g, rms[k], lr)
            got = m.get_params(blk)
            err = np.abs(got - params[k]).max() / np.abs(params[k]).max()
            upd = np.abs(got - params[k]).max() / (lr + 1e-30)        # error relative to the size of one update
            if not (err < 1e-6 and upd < 1e-2):
                ok = False
                print(f"rank {rank} step {step} block {blk}: rel err {err:.3e}, err/lr {upd:.3e}", flush=True)
            params[k] = got.copy()                                      # follow the device trajectory
        digest = hashlib.sha256(b"".join(m.get_params(b).tobytes() for b in blocks)).hexdigest()
        digs = [None] * world
        dist.all_gather_object(digs, digest)
        if len(set(digs)) != 1:
            ok = False
            print(f"rank {rank} step {step}: replicas diverged {digs}", flush=True)
        # RMSprop state is sharded: this rank owns [lo, hi) of the flat vector; check its own shard
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_FUSED_OK" if int(flag.item()) == 1 else "DP_FUSED_FAILED", flush=True)
    dist.barrier()
    m.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
