"""The host-side mirror of the reference's Torch7 interface (novel-vqa_b200/torch7_mirror.py), exercised the way
002_train_vqa_arch1/002_train_baseline.lua uses the modules, against the oracle.  GPU only (every call lands in libnvqa)."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import arch1 as A

pytestmark = pytest.mark.gpu


def test_reference_style_forward_and_update():
    import novel_vqa_b200 as nvm
    if nvm.device_count() == 0:
        pytest.fail("needs the B200 box")
    from novel_vqa_b200.torch7_mirror import (LSTM, netdef, CrossEntropyCriterion, optim, right_align, rnn_forward,
                                              sort_encoding_onehot_right_align, split_vector, join_vector, inverse_mapping)
    E, H, n, I, Cc, O, T, V, B = 12, 16, 2, 24, 20, 15, 6, 40, 9
    oc = A.Arch1Config(V=V, E=E, H=H, L=n, I=I, C=Cc, O=O, T=T)
    r = np.random.default_rng(0)
    lengths = r.integers(1, T + 1, B).astype(np.int32)
    q = np.zeros((B, T), dtype=np.int32)
    for b in range(B):
        q[b, :lengths[b]] = r.integers(1, V + 1, lengths[b])
    # dataset['question'] = right_align(...)                                       002_train_baseline.lua:113
    q_ra = right_align(q, lengths)
    assert np.array_equal(q_ra, A.right_align(q, lengths))
    # fv_sorted_q = sort_encoding_onehot_right_align(...)                          :208
    fv_sorted_q = sort_encoding_onehot_right_align(q_ra, lengths, V)
    ow, os_, osi, oinv = A.sort_encoding_right_align(q_ra, lengths)
    assert np.array_equal(fv_sorted_q[0], ow) and np.array_equal(fv_sorted_q[1], os_)
    assert np.array_equal(fv_sorted_q[2], osi + 1) and np.array_equal(fv_sorted_q[3], oinv + 1)
    assert np.array_equal(inverse_mapping(fv_sorted_q[2]), fv_sorted_q[3])
    # encoder_net_q = LSTM.lstm_conventional(E, H, 1, n, 0.5); getParameters(); uniform(-0.08, 0.08)   :147,177-178
    encoder_net_q = LSTM.lstm_conventional(E, H, 1, n, 0.5)
    encoder_w_q, _ = encoder_net_q.getParameters()
    encoder_w_q[:] = r.uniform(-0.5, 0.5, encoder_w_q.size)
    assert encoder_w_q.size == oc.n_enc
    encoder_net_q.evaluate()
    # word embeddings come from the oracle here (the nn.Sequential embedding has no module-level entry point)
    emb = A.split_flat(r.uniform(-0.5, 0.5, oc.n_emb).astype(np.float32), oc.emb_layout())
    y = A.embedding_forward(emb, fv_sorted_q[0], None).astype(np.float32)
    word_embedding_q = split_vector(y, fv_sorted_q[1] * 1)                        # rows per step
    offs = np.concatenate([[0], np.cumsum(fv_sorted_q[1])])
    word_embedding_q = [y[offs[i]:offs[i + 1]] for i in range(len(fv_sorted_q[1]))]
    # states_q = rnn_forward(buffer, repeatTensor(dummy_state), word_embedding_q, sizes)            :303
    buffer = [[encoder_net_q] * len(word_embedding_q)]
    states_q = rnn_forward(buffer, np.zeros((B, 2 * n * H), np.float32), word_embedding_q, fv_sorted_q[1])
    enc = A.split_flat(encoder_w_q, oc.enc_layout())
    ref_states, _ = A.rnn_forward(oc, enc, np.zeros((B, oc.S), np.float32), word_embedding_q, fv_sorted_q[1], None)
    assert_close(states_q[-1], ref_states[-1], 1e-4, "final LSTM state")
    tv_q = states_q[-1][fv_sorted_q[3] - 1]                                        # :306
    # multimodal_net = AxB(2*H*n, I, C, 0.5) ...                                                 :151-154
    axb = netdef.AxB(2 * H * n, I, Cc, 0.5, rnn_size=H, rnn_layers=n)
    mm_w, _ = axb.getParameters()
    mm_w[:] = r.uniform(-0.5, 0.5, mm_w.size)
    axb.evaluate()
    fv_im = A.l2_normalize_rows(np.maximum(0, r.standard_normal((B, I))).astype(np.float32))
    z = axb.forward([tv_q, fv_im])
    mm = A.split_flat(mm_w[:Cc * 2 * H * n + Cc + Cc * I + Cc], [("Wq", (Cc, 2 * H * n)), ("bq", (Cc,)), ("Wv", (Cc, I)), ("bv", (Cc,))])
    z_ref = np.tanh(tv_q @ mm["Wq"].T + mm["bq"]) * np.tanh(fv_im @ mm["Wv"].T + mm["bv"])
    assert_close(z, z_ref, 1e-4, "AxB output")
    # criterion = nn.CrossEntropyCriterion(); f = criterion:forward(scores, labels); dscores = criterion:backward   :308-310
    scores = r.standard_normal((B, O)).astype(np.float32) * 2
    labels = r.integers(1, O + 1, B)
    criterion = CrossEntropyCriterion(O)
    f = criterion.forward(scores, labels)
    f_ref, d_ref = A.cross_entropy(scores, labels)
    assert abs(f - f_ref) <= 1e-5 * abs(f_ref)
    assert_close(criterion.backward(scores, labels), d_ref, 1e-5, "dscores")
    with pytest.raises(nvm.NvqaError):
        criterion.forward(scores, np.zeros(B, dtype=np.int64))                     # label 0: out of range (App. C-10)
    # optim.rmsprop(JdJ, x, config, state)                                                       :408
    x = r.standard_normal(1000).astype(np.float32)
    x_ref, m_ref = x.copy(), np.zeros(1000, np.float32)
    state, config = {}, {"learningRate": 3e-4}
    for it in range(3):
        g = r.standard_normal(1000).astype(np.float32)
        _, fx = optim.rmsprop(lambda xx: (1.5, g), x, config, state)
        A.rmsprop_update(x_ref, g, m_ref, 3e-4)
        assert fx == [1.5]
    assert_close(x, x_ref, 1e-6, "optim.rmsprop x")
    assert_close(state["m"], m_ref, 1e-6, "optim.rmsprop state.m")
    assert np.array_equal(join_vector([x[:3], x[3:5]]), x[:5])
