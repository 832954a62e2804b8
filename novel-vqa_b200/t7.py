"""Torch7 ``torch.save`` / ``torch.load`` binary serialisation (host side) so that the
reference's checkpoints drop in: ``{encoder_w_q, embedding_w_q, multimodal_w}`` of 002_train_baseline.lua:401-402
(loaded by key in 004_eval_model.lua:154-163), ``{cnn_w, encoder_w_q, multimodal_w}`` of arch2, and the converted
autoencoder table ``{lookup, encoder}`` of 001_train_autoencoder/002_convert_text_model_arch1.lua:33-38.

Format (torch7 File.lua, binary mode, little-endian; the Torch7 sources are not in /root/reference, so this follows
the published layout restated in SURVEY App. B):

    object   := int32 type, payload
    type 0 nil | 1 number (float64) | 2 string (int32 len, bytes) | 5 boolean (int32)
    type 3 table  : int32 ref-index, then (first time) int32 count, count x (object key, object value)
    type 4 torch  : int32 ref-index, then (first time) string "V 1", string class-name, class payload
    type 6 / 7 / 8 function (dumped bytecode + upvalues): read and kept opaque
    Tensor payload  : int32 nDim, int64 size[nDim], int64 stride[nDim], int64 storageOffset (1-based), object storage
    Storage payload : int64 n, n raw elements
    any other torch class (nn modules ...): one object (the table of its fields)

Tensors are returned as NumPy arrays (views into their storage honouring size / stride / offset);
``torch.CudaTensor`` is accepted and read as float32 -- the reference saves the live flat CUDA views.
Pure host plumbing: no arithmetic of the hot path lives here.
"""
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, TYPE_RECUR_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION = 6, 8, 7

_DTYPES = {"Float": np.float32, "Double": np.float64, "Long": np.int64, "Int": np.int32, "Short": np.int16,
           "Byte": np.uint8, "Char": np.int8, "Cuda": np.float32, "CudaDouble": np.float64, "CudaLong": np.int64,
           "CudaInt": np.int32, "Half": np.float16, "CudaHalf": np.float16}
_NAME_OF = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long",
            np.dtype(np.int32): "Int", np.dtype(np.int16): "Short", np.dtype(np.uint8): "Byte", np.dtype(np.int8): "Char"}


class TorchObject(dict):
    """A torch class instance without a dedicated reader (e.g. nn.Sequential): its fields, plus the class name."""

    def __init__(self, torch_class, fields=None):
        super().__init__(fields or {})
        self.torch_class = torch_class

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class LuaFunction:
    def __init__(self, dumped, upvalues):
        self.dumped, self.upvalues = dumped, upvalues


class _Reader:
    def __init__(self, data):
        self.b, self.o, self.objects = memoryview(data), 0, {}

    def _take(self, fmt, n):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += n
        return v[0]

    def int(self):
        return self._take("i", 4)

    def long(self):
        return self._take("q", 8)

    def double(self):
        return self._take("d", 8)

    def string(self):
        n = self.int()
        s = bytes(self.b[self.o:self.o + n])
        self.o += n
        return s.decode("latin-1")

    def obj(self):
        t = self.int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            x = self.double()
            return int(x) if float(x).is_integer() and abs(x) < 2 ** 53 else x
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.int() == 1
        if t == TYPE_FUNCTION:
            n = self.int()
            dumped = bytes(self.b[self.o:self.o + n])
            self.o += n
            return LuaFunction(dumped, self.obj())
        if t not in (TYPE_TABLE, TYPE_TORCH, TYPE_RECUR_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION):
            raise ValueError(f"t7: unknown object type {t} at byte {self.o - 4}")
        idx = self.int()
        if idx in self.objects:
            return self.objects[idx]
        if t in (TYPE_RECUR_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION):
            n = self.int()
            dumped = bytes(self.b[self.o:self.o + n])
            self.o += n
            f = LuaFunction(dumped, None)
            self.objects[idx] = f
            f.upvalues = self.obj()
            return f
        if t == TYPE_TABLE:
            out = {}
            self.objects[idx] = out
            for _ in range(self.int()):
                k = self.obj()
                out[k] = self.obj()
            return out
        version = self.string()
        cls = self.string() if version.startswith("V ") else version
        if cls.startswith("torch.") and cls.endswith("Tensor"):
            kind = cls[len("torch."):-len("Tensor")]
            nd = self.int()
            size = [self.long() for _ in range(nd)]
            stride = [self.long() for _ in range(nd)]
            off = self.long() - 1
            storage = self.obj()
            if storage is None or nd == 0:
                arr = np.zeros([0] * max(nd, 1), dtype=_DTYPES.get(kind, np.float32))
            else:
                isz = storage.itemsize
                # a truncated / corrupt / hostile file must raise, not read out of bounds
                if off < 0 or any(n < 0 for n in size) or any(st < 0 for st in stride):
                    raise ValueError(f"{cls}: negative size, stride or storage offset")
                last = off + sum((n - 1) * st for n, st in zip(size, stride)) if all(n > 0 for n in size) else off
                if all(n > 0 for n in size) and last >= storage.size:
                    raise ValueError(f"{cls}: size {size} / stride {stride} / offset {off + 1} exceed the storage "
                                     f"({storage.size} elements)")
                if any(n == 0 for n in size):
                    arr = np.zeros(size, dtype=storage.dtype)
                else:
                    arr = np.lib.stride_tricks.as_strided(storage[off:], shape=size, strides=[st * isz for st in stride])
            self.objects[idx] = arr
            return arr
        if cls.startswith("torch.") and cls.endswith("Storage"):
            kind = cls[len("torch."):-len("Storage")]
            dt = np.dtype(_DTYPES[kind])
            n = self.long()
            arr = np.frombuffer(self.b, dtype=dt, count=n, offset=self.o).copy()
            self.o += n * dt.itemsize
            self.objects[idx] = arr
            return arr
        o = TorchObject(cls)
        self.objects[idx] = o
        fields = self.obj()
        if isinstance(fields, dict):
            o.update(fields)
        else:
            o["value"] = fields
        return o


def load(path):
    """torch.load(path) for binary files: tables -> dict, numbers -> int/float, tensors -> numpy arrays."""
    with open(path, "rb") as f:
        return _Reader(f.read()).obj()


class _Writer:
    def __init__(self, cuda):
        self.out, self.n, self.cuda = [], 0, cuda

    def int(self, v):
        self.out.append(struct.pack("<i", v))

    def long(self, v):
        self.out.append(struct.pack("<q", v))

    def string(self, s):
        b = s.encode("latin-1")
        self.int(len(b))
        self.out.append(b)

    def _index(self):
        self.n += 1
        self.int(self.n)

    def obj(self, v):
        if v is None:
            self.int(TYPE_NIL)
        elif isinstance(v, (bool, np.bool_)):
            self.int(TYPE_BOOLEAN)
            self.int(1 if v else 0)
        elif isinstance(v, (int, float, np.integer, np.floating)):
            self.int(TYPE_NUMBER)
            self.out.append(struct.pack("<d", float(v)))
        elif isinstance(v, str):
            self.int(TYPE_STRING)
            self.string(v)
        elif isinstance(v, np.ndarray):
            self.tensor(v)
        elif isinstance(v, dict):
            self.int(TYPE_TABLE)
            self._index()
            self.int(len(v))
            for k, x in v.items():
                self.obj(k)
                self.obj(x)
        elif isinstance(v, (list, tuple)):
            self.obj({i + 1: x for i, x in enumerate(v)})          # Lua arrays are 1-based tables
        else:
            raise TypeError(f"t7: cannot serialise {type(v)}")

    def tensor(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype not in _NAME_OF:
            raise TypeError(f"t7: no Torch7 tensor type for dtype {a.dtype}")
        kind = _NAME_OF[a.dtype]
        if self.cuda and kind == "Float":
            kind = "Cuda"                                           # what the scripts save from a GPU run
        self.int(TYPE_TORCH)
        self._index()
        self.string("V 1")
        self.string(f"torch.{kind}Tensor")
        self.int(a.ndim)
        for s in a.shape:
            self.long(s)
        for s in a.strides:
            self.long(s // a.itemsize)
        self.long(1)
        self.int(TYPE_TORCH)
        self._index()
        self.string("V 1")
        self.string(f"torch.{kind}Storage")
        self.long(a.size)
        self.out.append(a.tobytes())


def save(path, obj, cuda=False):
    """torch.save(path, obj): dict -> table, list -> 1-based table, numpy arrays -> torch.<Type>Tensor
    (float32 arrays as torch.CudaTensor when cuda=True, the class the reference's GPU runs write)."""
    w = _Writer(cuda)
    w.obj(obj)
    with open(path, "wb") as f:
        f.write(b"".join(w.out))


# ---- checkpoints of the path ---------------------------------------------------------------------------------------
ARCH1_KEYS = ("encoder_w_q", "embedding_w_q", "multimodal_w")          # 002_train_baseline.lua:401-402
ARCH2_KEYS = ("cnn_w", "encoder_w_q", "multimodal_w")                  # 003_train_vqa_arch2/002_train_baseline.lua
ARCH1_BLOCK = {"encoder_w_q": 0, "embedding_w_q": 1, "multimodal_w": 2}
ARCH2_BLOCK = {"cnn_w": 0, "encoder_w_q": 1, "multimodal_w": 2}


def save_checkpoint(path, model, cuda=True):
    """torch.save(path, {encoder_w_q=..., embedding_w_q=..., multimodal_w=...}) from a model handle."""
    blocks = ARCH2_BLOCK if model.ARCH == 2 else ARCH1_BLOCK
    save(path, {k: model.get_params(b) for k, b in blocks.items()}, cuda=cuda)


def load_checkpoint(path, model):
    """004_eval_model.lua:154-163: copy the three flat tensors into the nets, by key."""
    t = load(path)
    blocks = ARCH2_BLOCK if model.ARCH == 2 else ARCH1_BLOCK
    for k, b in blocks.items():
        w = np.ascontiguousarray(t[k], dtype=np.float32).ravel()
        if w.size != model.param_count(b):
            raise ValueError(f"checkpoint tensor {k} has {w.size} elements, the model expects {model.param_count(b)}")
        model.set_params(b, w)
    return t


def convert_autoencoder(lookup_table, encoder_flat):
    """001_train_autoencoder/002_convert_text_model_arch1.lua:33-38: saveModel = {lookup = LookupTable.weight:t(),
    encoder = encoder:getParameters()}.  lookup_table: [(V+1) x E] (or flat)."""
    enc = np.ascontiguousarray(encoder_flat, dtype=np.float32).ravel()
    return {"lookup": np.ascontiguousarray(np.asarray(lookup_table, dtype=np.float32).T), "encoder": enc.copy()}


def arch1_blocks_from_autoencoder(saved, V, E):
    """002_train_vqa_arch1/003_train_ae_based.lua:175-183: embedding Linear.weight [E x V] = lookup minus its last
    column (the START/END token), bias = 0; encoder_w_q = the saved encoder vector.  Returns (encoder_w_q, embedding_w_q)
    in Torch flat layout (multimodal_w stays uniform(-0.08, 0.08), :186)."""
    lookup = np.asarray(saved["lookup"], dtype=np.float32)
    if lookup.shape != (E, V + 1):
        raise ValueError(f"lookup has shape {lookup.shape}, expected {(E, V + 1)}")
    emb = np.concatenate([np.ascontiguousarray(lookup[:, :V]).ravel(), np.zeros(E, dtype=np.float32)])
    return np.ascontiguousarray(saved["encoder"], dtype=np.float32).ravel(), emb


def arch2_encoder_from_autoencoder(ae_encoder_flat, ae_lookup_flat, V=None, E=None):
    """003_train_vqa_arch2/003_train_ae_based.lua:150-152,191: ``encoder_model.encoder = modelT.ae.encoder:clone()``,
    ``encoder_model.lookup_table = modelT.ae.lookup_table:clone()``, then ``encoder_model:getParameters()`` -- the arch2
    ``encoder_w_q`` block is the autoencoder's encoder LSTM parameters followed by its LookupTable(V+1, E) weight
    (``misc/Encoder_lstm.lua:66-83``: core first, lookup table second), both taken as they are.  ``cnn_w`` and
    ``multimodal_w`` stay ``uniform(-0.08, 0.08)`` (``:188-194``).  Inputs: the flat parameter vectors of an autoencoder
    (``AEModel.get_params(BLOCK_AE_ENCODER)`` / ``(BLOCK_AE_LOOKUP)``, or tensors read with ``t7.load``)."""
    enc = np.ascontiguousarray(ae_encoder_flat, dtype=np.float32).ravel()
    lut = np.ascontiguousarray(ae_lookup_flat, dtype=np.float32).ravel()
    if V is not None and E is not None and lut.size != (V + 1) * E:
        raise ValueError(f"lookup table has {lut.size} elements, expected (V + 1) x E = {(V + 1) * E}")
    return np.concatenate([enc, lut])


def arch2_cnn_from_linear(weight, bias):
    """003_train_ae_based_wp_vgg.lua:174 / _wp_inc.lua: ``cnn_projection = nn.Sequential():add(modelT.cnn:get(40):clone())``
    -- the weakly-paired model's image projection nn.Linear(nhimage, E) becomes the arch2 ``cnn_w`` block
    (``getParameters()`` order: weight [E x nhimage], then bias [E])."""
    w = np.ascontiguousarray(weight, dtype=np.float32)
    b = np.ascontiguousarray(bias, dtype=np.float32).ravel()
    if w.ndim != 2 or w.shape[0] != b.size:
        raise ValueError(f"Linear weight {w.shape} and bias {b.shape} do not belong together")
    return np.concatenate([w.ravel(), b])

