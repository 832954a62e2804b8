"""Dataset loading and batch assembly of the stage scripts (host side).

Mirrors the ``dataset`` table of ``002_train_vqa_arch1/002_train_baseline.lua:84-123`` (train / val) and of
``004_eval_model.lua:68-95`` (test): same keys, same HDF5 dataset names, same JSON vocabulary, questions right-aligned
at load (``misc/RNNUtils.lua:54-61`` through the C ABI, ``nvqa_right_align``), plus ``next_batch`` (``:195-222``) and
``next_batch_val`` (``:227-259``).  Two deliberate differences, both on the reference's side of the boundary:

* the image features are NOT normalised here: with ``img_norm = 1`` the library normalises the gathered fc7 rows on
  the device inside the step (``imgnorm_drop_kernel``; row-wise, so gather-then-normalise equals the reference's
  normalise-then-gather ``:117-123``);
* ``torch.random(nqs)`` is Torch7's MT19937 stream, which cannot be reproduced here: the sampler takes a
  ``numpy.random.Generator`` -- the index *distribution* (uniform with replacement over 1..nqs) is the reference's.

Batches come back in the layout ``Arch1Model.set_batch_host`` / ``train_step_host`` take: right-aligned int32 questions
``[B x T]``, int32 lengths, float32 fc7 rows, int32 labels (1-based, as stored).  Indices stay 1-based where the
reference's are (``img_list``, ``answers``).  Plumbing only: the HDF5 parsing is ``h5lite`` (no libhdf5 in the image).
"""
import json

import numpy as np

from . import api, h5lite

_SPLIT_KEYS = {
    # reference key          -> HDF5 dataset              002_train_baseline.lua / 004_eval_model.lua
    "train": {"question": "ques_train", "lengths_q": "ques_length_train", "img_list": "img_pos_train",
              "answers": "answers", "ques_id": "question_id_train", "fv_im": "images_train"},
    "val": {"question": "ques_val", "lengths_q": "ques_length_val", "img_list": "img_pos_val",
            "answers": "answers_val", "ques_id": "question_id_val", "fv_im": "images_val"},
    "test": {"question": "ques_test", "lengths_q": "ques_length_test", "img_list": "img_pos_test",
             "ques_id": "question_id_test", "MC_ans_test": "MC_ans_test", "fv_im": "images_test"},
}


class VqaSplit:
    """One split of the reference's ``dataset`` table."""

    def __init__(self, name, fields):
        self.name = name
        for k, v in fields.items():
            setattr(self, k, v)
        self.fields = fields

    def __len__(self):
        return int(self.question.shape[0])

    def __getitem__(self, key):
        return self.fields[key]

    def batch(self, qinds):
        """Rows ``qinds`` (1-based, like the reference's LongTensor) -> (q_ra, lengths, fc7, labels | None)."""
        qi = np.asarray(qinds, dtype=np.int64) - 1
        if qi.size and (qi.min() < 0 or qi.max() >= len(self)):
            raise IndexError("question index out of range")
        iminds = self.img_list[qi].astype(np.int64) - 1                   # dataset['img_list'][qinds[i]], 1-based
        if iminds.size and (iminds.min() < 0 or iminds.max() >= self.fv_im.shape[0]):
            raise IndexError("img_pos points outside the image feature matrix")
        labels = self.fields["answers"][qi].astype(np.int32) if "answers" in self.fields else None
        return (np.ascontiguousarray(self.question[qi]), np.ascontiguousarray(self.lengths_q[qi]),
                np.ascontiguousarray(self.fv_im[iminds]), labels)


def _read(f, name, dtype):
    if name not in f:
        raise KeyError(f"{f.path}: no dataset '/{name}'")
    return np.ascontiguousarray(f[name].read().astype(dtype))


def load_split(input_ques_h5, input_img_h5, split):
    """The ``dataset[...]`` entries of one split; questions are right-aligned as at ``002_train_baseline.lua:113-114``."""
    names = _SPLIT_KEYS[split]
    fields = {}
    with h5lite.File(input_ques_h5) as fq:
        for key, ds in names.items():
            if key == "fv_im":
                continue
            if key == "ques_id" and ds not in fq:
                continue                                                   # the training script never reads the ids
            fields[key] = _read(fq, ds, np.int32)
    with h5lite.File(input_img_h5) as fi:
        fields["fv_im"] = _read(fi, names["fv_im"], np.float32)
    q, ln = fields["question"], fields["lengths_q"]
    if q.ndim != 2 or ln.shape != (q.shape[0],):
        raise ValueError(f"{input_ques_h5}: '{names['question']}' must be [n x T] with one length per row")
    if ln.size and (ln.min() < 0 or ln.max() > q.shape[1]):
        raise ValueError(f"{input_ques_h5}: question lengths outside 0..{q.shape[1]}")
    fields["question"] = api.right_align(q, ln)
    return VqaSplit(split, fields)


class VqaDataset:
    """``opt.input_json`` + ``opt.input_ques_h5`` + ``opt.input_img_h5`` of the stage scripts."""

    def __init__(self, input_json, input_ques_h5, input_img_h5, splits=("train", "val"), batch_size=500):
        with open(input_json) as f:
            self.json_file = json.load(f)
        self.ix_to_word = self.json_file["ix_to_word"]
        self.ix_to_ans = self.json_file.get("ix_to_ans", {})
        self.vocabulary_size_q = len(self.ix_to_word)                      # :125-127
        self.batch_size = batch_size
        self.splits = {s: load_split(input_ques_h5, input_img_h5, s) for s in splits}

    def __getitem__(self, split):
        return self.splits[split]

    @property
    def buffer_size_q(self):
        """``dataset['question']:size()[2]`` (``:134``): the T the model must be created with."""
        return int(next(iter(self.splits.values())).question.shape[1])

    def next_batch(self, rng, split="train"):
        """``dataset:next_batch()`` (``:195-222``): batch_size question rows drawn uniformly with replacement."""
        s = self.splits[split]
        qinds = rng.integers(1, len(s) + 1, size=self.batch_size)
        return s.batch(qinds)

    def next_batch_val(self, val_count, split="val"):
        """``dataset:next_batch_val(val_count)`` (``:227-259``): rows val_count+1 .. val_count+batch_size, the last batch
        shortened to what is left."""
        s = self.splits[split]
        n = min(self.batch_size, len(s) - val_count)
        if n <= 0:
            raise IndexError("val_count is past the end of the split")
        return s.batch(np.arange(val_count + 1, val_count + n + 1))

    def iter_eval(self, split="test"):
        """The evaluation loop of ``004_eval_model.lua:222-234``: consecutive batches over the whole split."""
        s = self.splits[split]
        for i0 in range(0, len(s), self.batch_size):
            qinds = np.arange(i0 + 1, min(i0 + self.batch_size, len(s)) + 1)
            yield qinds, s.batch(qinds)


def write_synthetic(input_json, input_ques_h5, input_img_h5, n_train=64, n_val=32, n_test=32, n_img=16, T=26, V=50, O=10,
                    I=32, seed=0, **h5_kwargs):
    """A small synthetic dataset in the reference's on-disk format (``000_prepro_vqa.py:274-292``, ``prepro_img.lua``):
    for tests and for trying the stage scripts without the VQA download."""
    rng = np.random.default_rng(seed)
    ques, img = {}, {}
    for split, n in (("train", n_train), ("val", n_val), ("test", n_test)):
        ln = rng.integers(1, T + 1, size=n)
        q = rng.integers(1, V + 1, size=(n, T))
        q[np.arange(T)[None, :] >= ln[:, None]] = 0                        # left-aligned, zero-padded, as prepro writes
        ques[f"ques_{split}"] = q.astype(np.uint32)
        ques[f"ques_length_{split}"] = ln.astype(np.uint32)
        ques[f"img_pos_{split}"] = rng.integers(1, n_img + 1, size=n).astype(np.uint32)
        ques[f"question_id_{split}"] = (np.arange(n) + {"train": 1000, "val": 2000, "test": 3000}[split]).astype(np.uint32)
        img[f"images_{split}"] = np.maximum(rng.standard_normal((n_img, I)), 0).astype(np.float32)
    ques["answers"] = rng.integers(1, O + 1, size=n_train).astype(np.uint32)
    ques["answers_val"] = rng.integers(1, O + 1, size=n_val).astype(np.uint32)
    mc = rng.integers(0, O + 1, size=(n_test, 18)).astype(np.uint32)
    ques["MC_ans_test"] = mc
    h5lite.write(input_ques_h5, ques, **h5_kwargs)
    h5lite.write(input_img_h5, img, **h5_kwargs)
    with open(input_json, "w") as f:
        json.dump({"ix_to_word": {str(i): f"w{i}" for i in range(1, V + 1)},
                   "ix_to_ans": {str(i): f"a{i}" for i in range(1, O + 1)}}, f)
