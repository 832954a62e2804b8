"""Data-parallel plumbing for the arch1 step (new functionality: the reference is single-GPU,
SURVEY 2.2).  One process per GPU; the batch is sharded (500 rows per rank, weak scaling); the flat
fp32 gradient is all-reduced (sum) in three buckets in gradient-readiness order -- multimodal
(ready after the head backward), encoder (after the LSTM backward), embedding (after the scatter) --
each issued asynchronously so that NCCL overlaps the next backward phase; 1/n_ranks is folded into
the clamp+RMSprop kernel (scale -> clamp -> update, i.e. single-process semantics at the global batch).

torch.distributed is plumbing only (rendezvous + ncclAllReduce on the library's device memory).
"""
import numpy as np

from . import api

# (phase to run, parameter block whose gradient becomes final in that phase)
PHASES = ((api.PHASE_HEAD, api.BLOCK_MULTIMODAL), (api.PHASE_LSTM, api.BLOCK_ENCODER),
          (api.PHASE_EMBED, api.BLOCK_EMBEDDING))


class _CudaArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def grad_bucket_views(model, device):
    """torch views (no copy) of the three gradient blocks of the library's flat gradient vector."""
    import torch
    _, g, off = model.device_views()
    views = {}
    for blk in (api.BLOCK_ENCODER, api.BLOCK_EMBEDDING, api.BLOCK_MULTIMODAL):
        n = model.param_count(blk)
        views[blk] = torch.as_tensor(_CudaArray(g + 4 * off[blk], n), device=f"cuda:{device}")
    return views


def bind_current_stream(model):
    """NCCL (through torch.distributed) orders its collectives against torch's CURRENT stream only, while the library
    enqueues on the model's own stream by default: an all_reduce on the gradient views could then read gradients the
    backward kernels have not written yet, and the optimizer could start before the all_reduce has finished.  The NCCL path
    therefore runs the model ON torch's current stream (nvqa_set_stream); called by train_step / backward_allreduce."""
    import ctypes as C
    import torch
    from . import _lib
    s = torch.cuda.current_stream().cuda_stream
    if s == 0:
        s = 1          # torch's default stream is the legacy default stream; nvqa_set_stream(NULL) means "the library's own
        #                stream", so name it explicitly: cudaStreamLegacy == (cudaStream_t)0x1
    if getattr(model, "_bound_stream", None) != s:
        model.sync()                                   # work already enqueued on the previous stream
        _lib.check(model.lib.nvqa_set_stream(model.handle, C.c_void_p(s)))
        model._bound_stream = s
    return s


def backward_allreduce(model, views, dist, world, bind=True):
    """backward in three phases; after each, all-reduce the bucket that just became final.
    ``dist`` is torch.distributed (or a stand-in with all_reduce(tensor, async_op=True)); bind=False only for such
    stand-ins (CPU tests), a real NCCL run must order the collectives against the model's stream."""
    if world == 1:
        model.backward(api.PHASE_ALL)
        return
    if bind:
        bind_current_stream(model)
    works = []
    for phase, blk in PHASES:
        model.backward(phase)
        works.append(dist.all_reduce(views[blk], async_op=True))
    for w in works:
        w.wait()


def train_step(model, views, dist, world, lr, seed, bind=True):
    """JdJ + all-reduce + clamp + RMSprop on an already-set batch (NCCL path: bucketed all-reduce overlapped with the
    backward phases, then the replicated optimizer kernel)."""
    if world == 1:
        model.train_step(lr, seed)         # single GPU: the library's own fused step (nvqa_train_step)
        return
    if bind:
        bind_current_stream(model)
    model.forward(api.MODE_TRAIN, seed)
    backward_allreduce(model, views, dist, world, bind)
    model.rmsprop_step(lr, grad_scale=1.0 / world)


def connect_fused(model, dist, rank, world):
    """Exchange the CUDA-IPC blobs of every rank's gradient / parameter / flag buffers (plumbing: torch.distributed
    all_gather_object) and map them, so that fused_train_step can run the collective inside the optimizer kernel."""
    import ctypes as C
    from . import _lib
    n = model.lib.nvqa_dp_blob_size()
    mine = C.create_string_buffer(n)
    _lib.check(model.lib.nvqa_dp_export(model.handle, mine))
    blobs = [None] * world
    if world > 1:
        dist.all_gather_object(blobs, bytes(mine.raw))
    else:
        blobs[0] = bytes(mine.raw)
    allb = C.create_string_buffer(b"".join(blobs), n * world)
    _lib.check(model.lib.nvqa_dp_connect(model.handle, rank, world, allb))
    if world > 1:
        dist.barrier()                 # nobody starts stepping before every rank has mapped its peers


def fused_train_step(model, lr, seed, alpha=0.99, eps=1e-8, wd=0.0, clamp=10.0):
    """JdJ with, per parameter block and as soon as its gradient is final, ONE kernel per rank: reduce-scatter over NVLink
    peer memory -> 1/world -> clamp -> RMSprop on the rank's shard -> all-gather of the updated parameters
    (csrc/dp_fused.cu; the multimodal block's exchange overlaps the LSTM backward).  No NCCL call on the step."""
    from . import _lib
    _lib.check(model.lib.nvqa_dp_train_step(model.handle, lr, seed, alpha, eps, wd, clamp))


def shard_mask(off, rank, world, whole_vector):
    """Which elements of the library's internal flat vector (block offsets ``off`` = nvqa_device_views' off4, in floats)
    rank ``rank`` of ``world`` reduces, updates and keeps the RMSprop state of -- the partition of csrc/dp_fused.cu
    (dp_range): every exchanged range [b0, b1) is cut into ``world`` runs of ceil(n / world) float4 units.
    ``whole_vector`` (nvqa_dp_layout): one range = the whole vector, else {encoder + embedding}, {multimodal}."""
    mine = np.zeros(int(off[3]), dtype=bool)
    for k0, k1 in (((0, 3),) if whole_vector else ((0, 2), (2, 3))):
        b0, b1 = int(off[k0]) // 4, int(off[k1]) // 4
        per = (b1 - b0 + world - 1) // world
        lo = min(b1, b0 + per * rank)
        mine[4 * lo:4 * min(b1, lo + per)] = True
    return mine


def average_then_update_reference(grads_per_rank, clamp=10.0):
    """What the collective + optimizer kernel compute, stated on host arrays (used by the gloo test):
    sum over ranks, scale 1/n, clamp."""
    n = len(grads_per_rank)
    s = np.sum(np.stack(grads_per_rank), axis=0, dtype=np.float32)
    return np.clip(s * np.float32(1.0 / n), -clamp, clamp)
