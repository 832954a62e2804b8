"""h5lite (HDF5 reader / writer of the reference's dataset files) and the dataset loader -- CPU only.

The reader is pinned against the one file in the image that libhdf5 itself wrote: scipy's MATLAB-7.3 test file (HDF5
behind a 512-byte user block); its dataset must equal the same variable of the MAT-5 twin file read by scipy.io.loadmat."""
import json
import os

import numpy as np
import pytest

from novel_vqa_b200 import data, h5lite


def _scipy_data():
    import scipy.io
    return os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data")


def test_reads_a_file_written_by_libhdf5():
    d = _scipy_data()
    h5, mat5 = os.path.join(d, "testhdf5_7.4_GLNX86.mat"), os.path.join(d, "testdouble_7.4_GLNX86.mat")
    if not (os.path.exists(h5) and os.path.exists(mat5)):
        pytest.skip("scipy's MATLAB test data is not installed")
    import scipy.io
    want = scipy.io.loadmat(mat5)["testdouble"]                       # 1 x 9: 0 : pi/4 : 2 pi
    with h5lite.File(h5) as f:
        assert f.superblock_version == 0 and f.keys() == ["testdouble"]
        ds = f["/testdouble"]
        assert ds.dtype == np.dtype("<f8") and ds.shape == (9, 1)     # MATLAB is column-major: 1 x 9 is stored as 9 x 1
        got = ds.read()
    assert np.array_equal(got.ravel(), want.ravel())
    assert np.allclose(got.ravel(), np.arange(9) * np.pi / 4)


def _sample():
    rng = np.random.default_rng(0)
    d = {"ques_train": rng.integers(0, 14773, (50, 26)).astype(np.uint32),
         "images_train": rng.standard_normal((7, 4096)).astype(np.float32),
         "answers": rng.integers(1, 1001, (50,)).astype(np.uint32),
         "empty": np.zeros((0, 26), np.uint32),
         "i64": np.arange(10, dtype=np.int64) - 5,
         "u8": np.arange(300, dtype=np.int64).astype(np.uint8),
         "f64": rng.standard_normal((3, 4, 5)),
         "big_endian": np.arange(6, dtype=">i4").reshape(2, 3)}
    for i in range(12):                                               # more entries than the default symbol-table node
        d[f"extra_{i:02d}"] = np.full((2,), i, np.int32)
    return d


@pytest.mark.parametrize("kw", [{}, {"chunks": 3}, {"chunks": 4, "compression": "gzip"}, {"chunks": 1000, "compression": "gzip"}])
def test_write_read_round_trip(tmp_path, kw):
    d = _sample()
    p = str(tmp_path / "a.h5")
    h5lite.write(p, d, **kw)
    with h5lite.File(p) as f:
        assert sorted(f.keys()) == sorted(d)
        for k, v in d.items():
            got = f[k].read()
            assert got.shape == v.shape and got.dtype == v.dtype.newbyteorder("="), k
            assert np.array_equal(got, v), k
            assert got.flags.c_contiguous
        assert "nope" not in f
        with pytest.raises(KeyError):
            f["nope"]


def test_rejects_what_it_does_not_know(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(str(p))
    with pytest.raises(h5lite.H5Error):
        h5lite.write(str(tmp_path / "c.h5"), {"c": np.zeros(3, np.complex64)})
    # a truncated file: the metadata still claims 4000 bytes of raw data, but the file ends early
    q = str(tmp_path / "t.h5")
    h5lite.write(q, {"zz": np.arange(1000, dtype=np.float32)})
    raw = open(q, "rb").read()
    with h5lite.File(q) as f:
        layout = f._find(f._messages(f._links["zz"]), 0x08)
    addr = int.from_bytes(layout[2:10], "little")
    f = h5lite.File(q)
    f._b = raw[:addr + 100] + raw[addr + 4000:]                       # same metadata, 3900 bytes of the data cut out
    f._b = f._b[:addr + 100]
    with pytest.raises(h5lite.H5Error):
        f._read_data(f._messages.__self__._messages(0) if False else [(0x08, 0, layout)], (1000,), np.dtype("<f4"))


def test_dataset_loader_mirrors_the_reference_tables(tmp_path):
    js, qh5, ih5 = (str(tmp_path / n) for n in ("data_prepro.json", "data_prepro.h5", "data_img.h5"))
    data.write_synthetic(js, qh5, ih5, n_train=64, n_val=21, n_test=10, n_img=9, T=26, V=50, O=10, I=32, seed=3)
    ds = data.VqaDataset(js, qh5, ih5, splits=("train", "val", "test"), batch_size=8)
    assert ds.vocabulary_size_q == 50 and ds.buffer_size_q == 26
    raw = h5lite.File(qh5)
    tr = ds["train"]
    q_raw, ln = raw["ques_train"].read().astype(np.int64), raw["ques_length_train"].read().astype(np.int64)
    # right_align (misc/RNNUtils.lua:54-61): row i shifted so that its last word sits in the last column
    for i in range(len(tr)):
        want = np.zeros(26, np.int64)
        want[26 - ln[i]:] = q_raw[i, :ln[i]]
        assert np.array_equal(tr.question[i], want)
    # next_batch: uniform with replacement, image rows gathered through img_pos (1-based)
    rng = np.random.default_rng(0)
    q, l, fc7, lab = ds.next_batch(rng)
    assert q.shape == (8, 26) and q.dtype == np.int32 and fc7.shape == (8, 32) and fc7.dtype == np.float32
    rng = np.random.default_rng(0)
    qinds = rng.integers(1, 65, size=8)
    img = h5lite.File(ih5)["images_train"].read()
    pos = raw["img_pos_train"].read().astype(np.int64)
    assert np.array_equal(fc7, img[pos[qinds - 1] - 1])
    assert np.array_equal(lab, raw["answers"].read().astype(np.int32)[qinds - 1])
    assert np.array_equal(l, ln[qinds - 1])
    # next_batch_val: consecutive rows, last batch shortened (:231-233)
    sizes, count = [], 0
    while count < len(ds["val"]):
        qv, lv, fv, yv = ds.next_batch_val(count)
        sizes.append(qv.shape[0])
        assert np.array_equal(yv, raw["answers_val"].read().astype(np.int32)[count:count + qv.shape[0]])
        count += qv.shape[0]
    assert sizes == [8, 8, 5]
    with pytest.raises(IndexError):
        ds.next_batch_val(21)
    # the evaluation loop covers every test question once, in order; labels are absent
    seen = []
    for qinds, (qt, lt, ft, yt) in ds.iter_eval("test"):
        assert yt is None
        seen.extend(qinds.tolist())
    assert seen == list(range(1, 11))
    assert ds["test"].MC_ans_test.shape == (10, 18)
    assert json.load(open(js))["ix_to_ans"]["3"] == "a3"


def test_loader_reads_compressed_chunked_files_too(tmp_path):
    js, qh5, ih5 = (str(tmp_path / n) for n in ("j.json", "q.h5", "i.h5"))
    data.write_synthetic(js, qh5, ih5, seed=5)
    js2, qh52, ih52 = (str(tmp_path / n) for n in ("j2.json", "q2.h5", "i2.h5"))
    data.write_synthetic(js2, qh52, ih52, seed=5, chunks=5, compression="gzip")
    a, b = data.VqaDataset(js, qh5, ih5), data.VqaDataset(js2, qh52, ih52)
    for s in ("train", "val"):
        for k in a[s].fields:
            assert np.array_equal(a[s][k], b[s][k]), (s, k)


def test_round_trip_property(tmp_path):
    """Random shapes / dtypes / storage layouts (hypothesis): what the writer stores the reader returns bit for bit."""
    from hypothesis import given, settings, strategies as st

    dtypes = st.sampled_from(["<u1", "<u2", "<u4", "<u8", "<i1", "<i2", "<i4", "<i8", "<f4", "<f8"])
    shapes = st.lists(st.integers(0, 7), min_size=1, max_size=3).map(tuple)
    counter = [0]

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(dtypes, shapes), min_size=1, max_size=6), st.sampled_from([None, 1, 2, 5]),
           st.booleans(), st.integers(0, 2 ** 31 - 1))
    def run(specs, chunks, gz, seed):
        rng = np.random.default_rng(seed)
        d = {}
        for i, (dt, shp) in enumerate(specs):
            a = rng.integers(0, 200, size=shp).astype(dt) if dt[1] in "ui" else rng.standard_normal(shp).astype(dt)
            d[f"ds_{i}_{'x' * (i * 3)}"] = a                          # names of different lengths (heap padding)
        counter[0] += 1
        p = str(tmp_path / f"p{counter[0]}.h5")
        kw = {} if chunks is None else {"chunks": chunks, "compression": "gzip" if gz else None}
        h5lite.write(p, d, **kw)
        with h5lite.File(p) as f:
            assert sorted(f.keys()) == sorted(d)
            for k, v in d.items():
                got = f[k].read()
                assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)

    run()
