"""World-size-2 test of the data-parallel plumbing on CPU (gloo): the bucket order, async all-reduce and the
1/n_ranks -> clamp -> RMSprop sequence of novel-vqa_b200/dp.py reproduce the single-process step on the global batch
(gradient = mean over the global batch; clamp after the reduction)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


class OracleRankModel:
    """Stand-in for Arch1Model on a CPU rank: gradients come from the oracle, buckets are torch CPU tensors."""

    def __init__(self, cfg, params, batch, seed):
        import torch
        from oracle import arch1 as A
        from novel_vqa_b200 import api
        q, ln, fv, lab = batch
        self.f, grads, _, _ = A.jdj(cfg, *params, q, ln, fv, lab, seed=seed, clamp=None)
        self.views = {api.BLOCK_ENCODER: torch.from_numpy(grads[0].copy()), api.BLOCK_EMBEDDING: torch.from_numpy(grads[1].copy()),
                      api.BLOCK_MULTIMODAL: torch.from_numpy(grads[2].copy())}
        self.phases = []

    def backward(self, phase):
        self.phases.append(phase)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import make_golden as mg
    from oracle import arch1 as A
    from novel_vqa_b200 import api, dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, q, lengths, enc, emb, mm, fc7, labels = mg.make_inputs()
    q_ra = A.right_align(q, lengths)
    fv = A.l2_normalize_rows(fc7)
    half = q_ra.shape[0] // world
    sl = slice(rank * half, (rank + 1) * half)
    model = OracleRankModel(cfg, (enc, emb, mm), (q_ra[sl], lengths[sl], fv[sl], labels[sl]), seed=None)
    dp.backward_allreduce(model, model.views, dist, world, bind=False)   # fake CPU model: no CUDA stream to bind
    assert model.phases == [api.PHASE_HEAD, api.PHASE_LSTM, api.PHASE_EMBED]        # readiness order
    # what clamp_rmsprop does with grad_scale = 1/world
    g = [np.clip(model.views[b].numpy() * np.float32(1.0 / world), -10, 10) for b in
         (api.BLOCK_ENCODER, api.BLOCK_EMBEDDING, api.BLOCK_MULTIMODAL)]
    if rank == 0:
        np.savez(out, genc=g[0], gemb=g[1], gmm=g[2])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_global_batch(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    import make_golden as mg
    from oracle import arch1 as A
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    cfg, q, lengths, enc, emb, mm, fc7, labels = mg.make_inputs()
    q_ra = A.right_align(q, lengths)
    f, grads, _, _ = A.jdj(cfg, enc, emb, mm, q_ra, lengths, A.l2_normalize_rows(fc7), labels, seed=None)
    for k, g in zip(("genc", "gemb", "gmm"), grads):
        np.testing.assert_allclose(got[k], g, rtol=2e-4, atol=1e-7, err_msg=k)


def test_reference_reduction_helper():
    from novel_vqa_b200 import dp
    a = np.array([1.0, 30.0, -50.0], dtype=np.float32)
    b = np.array([3.0, 10.0, -10.0], dtype=np.float32)
    np.testing.assert_allclose(dp.average_then_update_reference([a, b]), [2.0, 10.0, -10.0])


def test_shard_partition_covers_every_element_once():
    """dp.shard_mask restates csrc/dp_fused.cu's partition (which also shards the RMSprop state): for every world size the
    ranks' shards are disjoint and cover the flat vector, in both forms (whole vector / two ranges), also when a range is
    not a multiple of world x 4 floats."""
    import importlib
    dp = importlib.import_module("novel_vqa_b200.dp")
    for off in ([0, 3563520, 6518320, 13836824], [0, 40, 52, 100], [0, 8, 8, 12]):
        for world in range(1, 9):
            for whole in (False, True):
                masks = [dp.shard_mask(off, r, world, whole) for r in range(world)]
                total = np.sum(np.stack(masks).astype(np.int32), axis=0)
                assert total.shape == (off[3],) and (total == 1).all(), (off, world, whole)
                if not whole:          # no shard straddles the boundary of the two ranges
                    for mk in masks:
                        idx = np.flatnonzero(mk)
                        lo, hi = idx[idx < off[2]], idx[idx >= off[2]]
                        for part in (lo, hi):
                            assert part.size == 0 or (np.diff(part) == 1).all()
