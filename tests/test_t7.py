"""Torch7 .t7 serialisation (novel-vqa_b200/t7.py): byte layout against a hand-assembled
file, round trips, strided tensors / shared storages, CudaTensor class names, and the autoencoder -> arch1 converter."""
import struct

import numpy as np
import pytest

from novel_vqa_b200 import t7


def _s(s):
    return struct.pack("<i", len(s)) + s.encode()


def test_hand_assembled_file_is_read(tmp_path):
    """{encoder_w_q = FloatTensor{1,2,3}, n = 7, ok = true}: built byte by byte from the published layout."""
    storage = (struct.pack("<i", 4) + struct.pack("<i", 3) + _s("V 1") + _s("torch.FloatStorage") + struct.pack("<q", 3)
               + np.array([1, 2, 3], dtype="<f4").tobytes())
    tensor = (struct.pack("<i", 4) + struct.pack("<i", 2) + _s("V 1") + _s("torch.FloatTensor") + struct.pack("<i", 1)
              + struct.pack("<q", 3) + struct.pack("<q", 1) + struct.pack("<q", 1) + storage)
    body = (struct.pack("<i", 3) + struct.pack("<i", 1) + struct.pack("<i", 3)
            + struct.pack("<i", 2) + _s("encoder_w_q") + tensor
            + struct.pack("<i", 2) + _s("n") + struct.pack("<i", 1) + struct.pack("<d", 7.0)
            + struct.pack("<i", 2) + _s("ok") + struct.pack("<i", 5) + struct.pack("<i", 1))
    p = tmp_path / "hand.t7"
    p.write_bytes(body)
    t = t7.load(str(p))
    assert t["n"] == 7 and t["ok"] is True
    assert t["encoder_w_q"].dtype == np.float32 and t["encoder_w_q"].tolist() == [1.0, 2.0, 3.0]
    # and the writer produces the same bytes for the same table
    q = tmp_path / "again.t7"
    t7.save(str(q), {"encoder_w_q": np.array([1, 2, 3], dtype=np.float32), "n": 7, "ok": True})
    assert q.read_bytes() == body


def test_round_trip_nested_and_cuda_class(tmp_path):
    r = np.random.default_rng(0)
    obj = {"encoder_w_q": r.standard_normal(1000).astype(np.float32), "embedding_w_q": r.standard_normal((7, 5)).astype(np.float32),
           "meta": {"iter": 2500, "lr": 3e-4, "name": "model", 1: "first", 2: [1.5, None, False]},
           "idx": np.arange(6, dtype=np.int64).reshape(2, 3)}
    for cuda in (False, True):
        p = tmp_path / f"rt{cuda}.t7"
        t7.save(str(p), obj, cuda=cuda)
        raw = p.read_bytes()
        assert (b"torch.CudaTensor" in raw) == cuda and (b"torch.FloatTensor" in raw) != cuda
        back = t7.load(str(p))
        assert np.array_equal(back["encoder_w_q"], obj["encoder_w_q"]) and back["encoder_w_q"].dtype == np.float32
        assert np.array_equal(back["embedding_w_q"], obj["embedding_w_q"])
        assert np.array_equal(back["idx"], obj["idx"]) and back["idx"].dtype == np.int64
        assert back["meta"]["iter"] == 2500 and abs(back["meta"]["lr"] - 3e-4) < 1e-18 and back["meta"][1] == "first"
        assert back["meta"][2] == {1: 1.5, 2: None, 3: False}


def test_strided_view_and_shared_storage(tmp_path):
    """A transposed view (stride (1, 4)) at storage offset 3, and a second tensor referencing the SAME storage object."""
    data = np.arange(20, dtype="<f4")
    storage = (struct.pack("<i", 4) + struct.pack("<i", 3) + _s("V 1") + _s("torch.FloatStorage") + struct.pack("<q", 20) + data.tobytes())
    t_a = (struct.pack("<i", 4) + struct.pack("<i", 2) + _s("V 1") + _s("torch.FloatTensor") + struct.pack("<i", 2)
           + struct.pack("<qq", 4, 3) + struct.pack("<qq", 1, 4) + struct.pack("<q", 4) + storage)
    t_b = (struct.pack("<i", 4) + struct.pack("<i", 4) + _s("V 1") + _s("torch.CudaTensor") + struct.pack("<i", 1)
           + struct.pack("<q", 5) + struct.pack("<q", 1) + struct.pack("<q", 16) + struct.pack("<i", 4) + struct.pack("<i", 3))
    body = (struct.pack("<i", 3) + struct.pack("<i", 1) + struct.pack("<i", 2) + struct.pack("<i", 2) + _s("a") + t_a
            + struct.pack("<i", 2) + _s("b") + t_b)
    p = tmp_path / "strided.t7"
    p.write_bytes(body)
    t = t7.load(str(p))
    assert np.array_equal(t["a"], data[3:3 + 12].reshape(3, 4).T)
    assert np.array_equal(t["b"], data[15:20])


def test_generic_torch_object_fields(tmp_path):
    """An nn module without a dedicated reader is returned with its fields (how protos.ae.lookup_table is reached)."""
    w = np.array([[1, 2], [3, 4]], dtype="<f4")
    storage = (struct.pack("<i", 4) + struct.pack("<i", 4) + _s("V 1") + _s("torch.FloatStorage") + struct.pack("<q", 4) + w.tobytes())
    tensor = (struct.pack("<i", 4) + struct.pack("<i", 3) + _s("V 1") + _s("torch.FloatTensor") + struct.pack("<i", 2)
              + struct.pack("<qq", 2, 2) + struct.pack("<qq", 2, 1) + struct.pack("<q", 1) + storage)
    fields = struct.pack("<i", 3) + struct.pack("<i", 2) + struct.pack("<i", 1) + struct.pack("<i", 2) + _s("weight") + tensor
    body = struct.pack("<i", 4) + struct.pack("<i", 1) + _s("V 1") + _s("nn.LookupTable") + fields
    p = tmp_path / "mod.t7"
    p.write_bytes(body)
    m = t7.load(str(p))
    assert m.torch_class == "nn.LookupTable" and np.array_equal(m.weight, w)


def test_autoencoder_to_arch1_conversion():
    """002_convert_text_model_arch1.lua:33-38 then 003_train_ae_based.lua:175-183."""
    V, E = 11, 4
    r = np.random.default_rng(1)
    lut = r.standard_normal((V + 1, E)).astype(np.float32)
    enc = r.standard_normal(50).astype(np.float32)
    saved = t7.convert_autoencoder(lut, enc)
    assert saved["lookup"].shape == (E, V + 1) and np.array_equal(saved["lookup"][:, 3], lut[3])
    enc_w, emb_w = t7.arch1_blocks_from_autoencoder(saved, V, E)
    assert np.array_equal(enc_w, enc)
    W = emb_w[:V * E].reshape(E, V)                     # nn.Linear(V, E).weight
    assert np.array_equal(W, lut[:V].T) and np.all(emb_w[V * E:] == 0)
    # one-hot Linear with this weight == LookupTable gather of the first V rows
    onehot = np.eye(V, dtype=np.float32)[[2, 7]]
    assert np.allclose(onehot @ W.T, lut[[2, 7]])


def test_result_json_emission(tmp_path):
    """004_eval_model.lua:248-273: OpenEnded / MultipleChoice result files; 004_eval_model_lf.lua late fusion."""
    import json

    from novel_vqa_b200 import results
    r = np.random.default_rng(0)
    n, O = 20, 30
    scores = r.standard_normal((n, O)).astype(np.float32)
    scores[3, 5] = scores[3, 9] = scores[3].max() + 1           # tie: the first maximum wins (torch.max)
    ix_to_ans = {str(i): f"ans{i}" for i in range(1, O + 1)}
    mc = np.zeros((n, 18), dtype=np.int32)
    for i in range(n):
        mc[i, :4] = r.choice(np.arange(1, O + 1), 4, replace=False)
    qids = np.arange(1000, 1000 + n)
    oe, mcp = tmp_path / "oe.json", tmp_path / "mc.json"
    results.write_results(str(oe), str(mcp), qids, scores, ix_to_ans, mc_ids=mc)
    a = json.load(open(oe))
    assert a[3] == {"question_id": 1003, "answer": "ans6"} and len(a) == n
    b = json.load(open(mcp))
    for i in range(n):
        cand = [c for c in mc[i] if c]
        assert b[i]["answer"] == f"ans{cand[int(np.argmax([scores[i, c - 1] for c in cand]))]}"
    lf = results.late_fusion_scores(scores, 2 * scores, 0.25, 0.75)
    assert lf.dtype == np.float64 and np.allclose(lf, 1.75 * scores, rtol=1e-6)


def test_autoencoder_to_arch2_conversion():
    """003_train_vqa_arch2/003_train_ae_based.lua:150-152,191: encoder_w_q = [AE encoder LSTM | AE LookupTable], in the
    layout the arch2 oracle (and the library) splits it with; cnn_w from a Linear of the weakly-paired model."""
    from oracle import ae as AE, arch1 as A, arch2 as A2
    V, E, H = 13, 8, 8
    acfg = AE.AEConfig(V=V, E=E, H=H, L=1, T=5)
    r = np.random.default_rng(2)
    enc = r.standard_normal(acfg.n_enc).astype(np.float32)
    lut = r.standard_normal(acfg.n_lut).astype(np.float32)
    blk = t7.arch2_encoder_from_autoencoder(enc, lut, V=V, E=E)
    c2 = A2.Arch2Config(V=V, E=E, H=H, L=1, I=6, O=3, T=5)
    assert blk.size == c2.n_enc
    parts = A.split_flat(blk, c2.enc_layout())
    ae_parts = A.split_flat(enc, acfg.enc_layout())
    for k in ("Wi0", "bi0", "Wh0", "bh0"):
        assert np.array_equal(parts[k], ae_parts[k]), k
    assert np.array_equal(parts["lookup"], lut.reshape(V + 1, E))
    with pytest.raises(ValueError):
        t7.arch2_encoder_from_autoencoder(enc, lut[:-1], V=V, E=E)
    W, b = r.standard_normal((E, 6)).astype(np.float32), r.standard_normal(E).astype(np.float32)
    cnn = A.split_flat(t7.arch2_cnn_from_linear(W, b), c2.cnn_layout())
    assert np.array_equal(cnn["Wcnn"], W) and np.array_equal(cnn["bcnn"], b)


@pytest.mark.parametrize("size,stride,offset", [((4, 3), (1, 4), 12), ((30,), (1,), 1), ((5,), (-1,), 5), ((5,), (1,), 0)])
def test_corrupt_tensor_headers_raise_instead_of_reading_out_of_bounds(tmp_path, size, stride, offset):
    """size / stride / storageOffset fields that reach past the 20-element storage (or are negative) must raise."""
    data = np.arange(20, dtype="<f4")
    storage = (struct.pack("<i", 4) + struct.pack("<i", 3) + _s("V 1") + _s("torch.FloatStorage") + struct.pack("<q", 20) + data.tobytes())
    nd = len(size)
    t_a = (struct.pack("<i", 4) + struct.pack("<i", 2) + _s("V 1") + _s("torch.FloatTensor") + struct.pack("<i", nd)
           + struct.pack("<" + "q" * nd, *size) + struct.pack("<" + "q" * nd, *stride) + struct.pack("<q", offset) + storage)
    p = tmp_path / "bad.t7"
    p.write_bytes(t_a)
    with pytest.raises(ValueError):
        t7.load(str(p))
