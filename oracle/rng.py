"""Counter-based dropout-mask generator shared (bit-for-bit) by the oracle and the CUDA kernels.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws its Dropout masks from
Torch7's MT19937 / cuRAND streams (nn.Dropout v2: ``mask ~ Bernoulli(1-p)/(1-p)``),
which cannot be reproduced without Torch7; parity tests therefore either pass masks explicitly
or use this stateless hash on both sides.  The CUDA twin is ``nvqa_keep_scale`` in
``novel-vqa_b200/csrc/common.cuh``.

One 32-bit hash word yields four 8-bit lanes; element ``idx`` uses lane ``idx & 3`` of word
``idx >> 2``.  An element is kept iff ``byte >= thresh``, ``thresh = min(255, round(p*256))``; kept elements
are scaled by ``256 / (256 - thresh)``, the exact reciprocal of the quantised keep probability, so E[mask] = 1
for every p (exactly 2.0 = 1/(1-p) for the reference's p = 0.5 everywhere).
"""
import numpy as np

# dropout sites ("streams") of the arch1 / arch2 step
STREAM_EMB = 1      # embedding Dropout(0.5)        002_train_baseline.lua:143
STREAM_LSTM0 = 2    # + (L-1): Dropout between LSTM layer L and L+1   misc/LSTM.lua:37
STREAM_AXB_Q = 16   # netdef.AxB Dropout on q       misc/netdef.lua:10
STREAM_AXB_I = 17   # netdef.AxB Dropout on i       misc/netdef.lua:11
STREAM_HEAD = 18    # Dropout before the classifier 002_train_baseline.lua:153

_M1 = np.uint32(0x85EBCA6B)
_M2 = np.uint32(0xC2B2AE35)
_G = np.uint32(0x9E3779B1)


def _mix32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= _M1
        x ^= x >> np.uint32(13)
        x *= _M2
        x ^= x >> np.uint32(16)
    return x


def stream_key(seed, stream):
    """32-bit key of (seed, stream); seed is a 64-bit integer (lo/hi words are both mixed in)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo = np.uint32(seed & 0xFFFFFFFF)
    hi = np.uint32(seed >> 32)
    with np.errstate(over="ignore"):
        k = _mix32(lo ^ _mix32(hi + np.uint32(stream) * _G))
    return np.uint32(k)


def keep_bytes(seed, stream, idx):
    """The 8-bit uniform of every element index in ``idx`` (any shape, int64)."""
    idx = np.asarray(idx, dtype=np.uint64)
    w = idx >> np.uint64(2)
    lane = (idx & np.uint64(3)).astype(np.uint32)
    wlo = (w & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    whi = (w >> np.uint64(32)).astype(np.uint32)
    key = stream_key(seed, stream)
    with np.errstate(over="ignore"):
        h = _mix32(wlo * _G + _mix32(whi ^ key))
    return ((h >> (lane * np.uint32(8))) & np.uint32(0xFF)).astype(np.uint32)


def keep_scale(seed, stream, idx, p, dtype=np.float32):
    """Dropout multiplier (0 or 1/(1-p)) for each element index."""
    thresh = np.uint32(min(255, int(np.float32(p) * np.float32(256.0) + np.float32(0.5))))
    scale = dtype(np.float32(256.0) / np.float32(256 - int(thresh)))
    return np.where(keep_bytes(seed, stream, idx) >= thresh, scale, dtype(0.0)).astype(dtype)
