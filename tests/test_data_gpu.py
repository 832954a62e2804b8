"""The stage-script flow from the reference's on-disk formats: data_prepro.h5 / data_img.h5 /
data_prepro.json -> dataset:next_batch() -> JdJ + rmsprop on the device (002_train_baseline.lua:195-222,272-335,408),
dataset:next_batch_val() -> validation loss (:337-371), and the evaluation loop with OpenEnded / MultipleChoice result JSON
(004_eval_model.lua:222-273) -- against the CPU oracle fed with the same arrays."""
import json

import numpy as np
import pytest

from conftest import assert_close, rel_err
from oracle import arch1 as A

pytestmark = pytest.mark.gpu


def test_train_validate_and_emit_answers_from_hdf5(tmp_path):
    import novel_vqa_b200 as nvm
    from novel_vqa_b200 import data, results
    if nvm.device_count() == 0:
        pytest.fail("no sm_100 device visible: GPU tests must run on the B200 box (no CPU fallback)")
    js, qh5, ih5 = (str(tmp_path / n) for n in ("data_prepro.json", "data_prepro.h5", "data_img.h5"))
    data.write_synthetic(js, qh5, ih5, n_train=96, n_val=40, n_test=50, n_img=12, T=12, V=60, O=10, I=40, seed=11)
    ds = data.VqaDataset(js, qh5, ih5, splits=("train", "val", "test"), batch_size=32)
    cfg = nvm.Arch1Config(V=ds.vocabulary_size_q, E=24, H=64, L=2, I=40, C=48, O=10, T=ds.buffer_size_q, B=32)
    oc = A.Arch1Config(V=cfg.V, E=cfg.E, H=cfg.H, L=cfg.L, I=cfg.I, C=cfg.C, O=cfg.O, T=cfg.T, p=cfg.dropout)
    enc, emb, mm = nvm.synth_params(cfg, seed=2)
    enc, emb, mm = enc * 3, emb * 3, mm * 3
    m = nvm.Arch1Model(cfg, precision=3)
    for blk, w in ((nvm.BLOCK_ENCODER, enc), (nvm.BLOCK_EMBEDDING, emb), (nvm.BLOCK_MULTIMODAL, mm)):
        m.set_params(blk, w)

    # ---- two training iterations on batches drawn like dataset:next_batch() ----
    rng = np.random.default_rng(7)
    w = [enc.copy(), emb.copy(), mm.copy()]
    rms = [np.zeros_like(x) for x in w]
    lr = 3e-4
    for it in range(2):
        q, ln, fc7, lab = ds.next_batch(rng)
        f, grads, _, _ = A.jdj(oc, w[0], w[1], w[2], q, ln, A.l2_normalize_rows(fc7), lab, seed=100 + it)
        got = m.train_step_host(q, ln, fc7, lab, lr, 100 + it)
        assert abs(got - f) <= 1e-4 * abs(f)
        for k in range(3):                                             # clamp + rmsprop (:329,408), oracle side
            g = np.clip(grads[k], -10, 10).astype(np.float32)
            rms[k] = (0.99 * rms[k] + 0.01 * g * g).astype(np.float32)
            w[k] = (w[k] - lr * g / (np.sqrt(rms[k]) + 1e-8)).astype(np.float32)
    # RMSprop's first steps move every weight by ~10 lr whatever its gradient: compare the weights in rel-L2 and the
    # *updates* separately (as test_training_trajectory_against_golden does)
    for blk, w0, want in zip((nvm.BLOCK_ENCODER, nvm.BLOCK_EMBEDDING, nvm.BLOCK_MULTIMODAL), (enc, emb, mm), w):
        got = m.get_params(blk)
        e2, _ = rel_err(got, want)
        assert e2 <= 1e-4, f"params {blk} after two iterations: rel-l2 {e2:.3e}"
        u2, _ = rel_err(got.astype(np.float64) - w0, want.astype(np.float64) - w0)
        assert u2 <= 5e-3, f"update {blk}: rel-l2 {u2:.3e}"
    w = [m.get_params(b) for b in (nvm.BLOCK_ENCODER, nvm.BLOCK_EMBEDDING, nvm.BLOCK_MULTIMODAL)]   # oracle continues from the device weights

    # ---- validation pass: consecutive batches, last one short ----
    count, losses = 0, []
    while count < len(ds["val"]):
        q, ln, fc7, lab = ds.next_batch_val(count)
        m.set_batch_host(q, ln, fc7, lab)
        m.forward(nvm.MODE_EVAL, 0)
        f, _, _, _ = A.jdj(oc, w[0], w[1], w[2], q, ln, A.l2_normalize_rows(fc7), lab, seed=None)
        assert abs(m.loss() - f) <= 1e-4 * abs(f)
        losses.append((m.loss(), q.shape[0]))
        count += q.shape[0]
    assert [n for _, n in losses] == [32, 8]

    # ---- evaluation loop + answer emission ----
    n = len(ds["test"])
    scores = np.zeros((n, cfg.O), np.float32)
    pred = np.zeros(n, np.int32)
    for qinds, (q, ln, fc7, _) in ds.iter_eval("test"):
        pred[qinds - 1] = m.eval_step_host(q, ln, fc7)
        scores[qinds - 1] = m.scores(q.shape[0])
    oe, mc = str(tmp_path / "OpenEnded.json"), str(tmp_path / "MultipleChoice.json")
    results.write_results(oe, mc, ds["test"].ques_id, scores, ds.ix_to_ans, mc_ids=ds["test"].MC_ans_test, pred=pred)
    out = json.load(open(oe))
    assert [r["question_id"] for r in out] == ds["test"].ques_id.tolist()
    assert np.array_equal(pred, scores.argmax(axis=1) + 1)              # torch.max: first maximum, 1-based
    assert [r["answer"] for r in out] == [f"a{p}" for p in pred]
    # oracle scores for the whole split (evaluate mode) -> same argmax wherever the top-2 margin is not a rounding tie
    t = ds["test"]
    _, _, want, _ = A.jdj(oc, w[0], w[1], w[2], t.question, t.lengths_q, A.l2_normalize_rows(t.fv_im[t.img_list - 1]),
                          np.ones(n, np.int32), seed=None)
    assert_close(scores, want, 1e-4, "test scores")
    srt = np.sort(want, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-4 * np.abs(srt[:, -1])
    assert np.array_equal(pred[clear], want.argmax(axis=1)[clear] + 1)
    # MultipleChoice: the best non-zero candidate of MC_ans_test (004_eval_model.lua:257-271)
    for r, s, cand in zip(json.load(open(mc)), scores, t.MC_ans_test):
        ids = [c for c in cand if c != 0]
        best = max(ids, key=lambda c: (s[c - 1], -ids.index(c)))
        assert r["answer"] == f"a{best}"
    m.close()
