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


@pytest.mark.parametrize("prec_name,prec,tol", [("fp32_simt", 0, 1e-4), ("bf16x2", 3, 1e-4)])
@pytest.mark.parametrize("mode", ["evaluate", "training"])
def test_reference_shaped_jdj_through_module_calls(prec_name, prec, tol, mode):
    """JdJ of 002_train_baseline.lua:272-335 written the way the reference writes it -- embedding_net_q:forward ->
    rnn_forward over dupe_rnn clones -> multimodal_net:forward -> criterion -> multimodal_net:backward -> rnn_backward ->
    embedding_net_q:backward -> sum of the clones' dW -> clamp -> optim.rmsprop -- entirely through the module-level C-ABI
    entry points, against (a) the oracle and (b) the fused nvqa_train_step_host path."""
    import novel_vqa_b200 as nvm
    if nvm.device_count() == 0:
        pytest.fail("needs the B200 box")
    from novel_vqa_b200 import torch7_mirror as T7
    cfg = nvm.Arch1Config(V=60, E=12, H=16, L=2, I=24, C=20, O=15, T=6, B=16)
    oc = A.Arch1Config(V=cfg.V, E=cfg.E, H=cfg.H, L=cfg.L, I=cfg.I, C=cfg.C, O=cfg.O, T=cfg.T, p=cfg.dropout)
    B, seed, lr = 11, 77, 3e-4
    enc0, emb0, mm0 = nvm.synth_params(cfg, seed=3)
    enc0, emb0, mm0 = enc0 * 4, emb0 * 4, mm0 * 4
    q_ra, lengths, fc7, labels = nvm.synth_batch(cfg, B, seed=4, min_len=1)
    fv_im = A.l2_normalize_rows(fc7)                                               # :117-123
    train = mode == "training"

    # ---- model construction (:139-181)
    embedding_net_q, encoder_net_q, multimodal_net, criterion, _ = T7.build_arch1_nets(cfg, prec)
    embedding_w_q, embedding_dw_q = embedding_net_q.getParameters()
    encoder_w_q, encoder_dw_q = encoder_net_q.getParameters()
    multimodal_w, multimodal_dw = multimodal_net.getParameters()
    embedding_w_q[:], encoder_w_q[:], multimodal_w[:] = emb0, enc0, mm0
    sizes_w = [encoder_w_q.size, embedding_w_q.size, multimodal_w.size]
    optimize_parameters = T7.join_vector([encoder_w_q, embedding_w_q, multimodal_w])   # :183
    encoder_net_buffer_q = T7.dupe_rnn(encoder_net_q, cfg.T)                       # :269
    for net in (embedding_net_q, encoder_net_q, multimodal_net):
        net.training() if train else net.evaluate()
    for clone in encoder_net_buffer_q[0]:
        clone.training() if train else clone.evaluate()

    def JdJ(x):
        params = T7.split_vector(x, sizes_w)                                       # :274
        for i in range(cfg.T):                                                     # :275-280 (always re-copied, App. C-4)
            encoder_net_buffer_q[0][i].getParameters()[0][:] = params[0]
        embedding_w_q[:], multimodal_w[:] = params[1], params[2]
        for i in range(cfg.T):                                                     # :283-288
            encoder_net_buffer_q[0][i].zeroGradParameters()
        embedding_net_q.zeroGradParameters()
        multimodal_net.zeroGradParameters()
        fv_sorted_q = T7.sort_encoding_onehot_right_align(q_ra, lengths, cfg.V)    # :208
        sizes = fv_sorted_q[1]
        if train:   # Torch7's RNG cannot be reproduced: the masks of the shared counter hash, in the packed layout
            pm = A.build_masks(oc, seed, B, sizes, fv_sorted_q[2] - 1)
            embedding_net_q.masks = pm["emb"]
            for i in range(len(sizes)):
                encoder_net_buffer_q[0][i].masks = np.stack(pm["lstm"][i])
            multimodal_net.masks = (pm["q"], pm["i"], pm["z"])
        word_embedding_q = T7.split_vector(embedding_net_q.forward(fv_sorted_q[0]), sizes)         # :300
        init = np.zeros((B, 2 * cfg.L * cfg.H), np.float32)                                          # :303
        states_q = T7.rnn_forward(encoder_net_buffer_q, init, word_embedding_q, sizes)
        tv_q = states_q[-1][fv_sorted_q[3] - 1]                                                      # :306
        scores = multimodal_net.forward([tv_q, fv_im])                                               # :307
        f = criterion.forward(scores, labels)                                                        # :308
        dscores = criterion.backward(scores, labels)                                                 # :310
        tmp = multimodal_net.backward([tv_q, fv_im], dscores)                                        # :312
        dtv_q = tmp[0][fv_sorted_q[2] - 1]                                                           # :313
        _, dword_embedding_q = T7.rnn_backward(encoder_net_buffer_q, dtv_q, None, states_q, word_embedding_q, sizes)   # :316
        embedding_net_q.backward(fv_sorted_q[0], T7.join_vector([d.ravel() for d in dword_embedding_q]).reshape(-1, cfg.E))  # :319-320
        encoder_adw_q = np.zeros_like(encoder_dw_q)                                                  # :323-326
        for i in range(len(sizes)):
            encoder_adw_q += encoder_net_buffer_q[0][i].getParameters()[1]
        gradients = T7.join_vector([encoder_adw_q, embedding_dw_q, multimodal_dw])                   # :328
        np.clip(gradients, -10, 10, out=gradients)                                                   # :329
        JdJ.scores = scores
        return f, gradients

    # (a) against the oracle
    f, gradients = JdJ(optimize_parameters)
    f_ref, g_ref, scores_ref, _ = A.jdj(oc, enc0, emb0, mm0, q_ra, lengths, fv_im, labels, seed=seed if train else None)
    assert_close(JdJ.scores, scores_ref, tol, f"{prec_name} {mode} module-level scores")
    assert abs(f - f_ref) <= tol * abs(f_ref)
    for got, want, what in zip(T7.split_vector(gradients, sizes_w), g_ref, ("encoder", "embedding", "multimodal")):
        assert_close(got, want, tol, f"{prec_name} {mode} module-level {what} gradient")

    # (b) against the fused path of the same library: nvqa_train_step_host (hash masks = the explicit masks above)
    m = nvm.Arch1Model(cfg, precision=prec, img_norm=0)
    for blk, w in ((nvm.BLOCK_ENCODER, enc0), (nvm.BLOCK_EMBEDDING, emb0), (nvm.BLOCK_MULTIMODAL, mm0)):
        m.set_params(blk, w)
    m.set_batch_host(q_ra, lengths, fv_im, labels)
    m.forward(nvm.MODE_TRAIN if train else nvm.MODE_EVAL, seed)
    m.backward()
    assert abs(m.loss() - f) <= tol * abs(f)
    for blk, got, what in zip((nvm.BLOCK_ENCODER, nvm.BLOCK_EMBEDDING, nvm.BLOCK_MULTIMODAL), T7.split_vector(gradients, sizes_w),
                              ("encoder", "embedding", "multimodal")):
        assert_close(got, np.clip(m.get_grads(blk), -10, 10), tol, f"{prec_name} {mode} module-level vs fused {what} gradient")

    # optim.rmsprop(JdJ, optimize_parameters, optimize, state)  (:408) against the oracle's update
    state, optimize = {}, {"learningRate": lr}
    T7.optim.rmsprop(JdJ, optimize_parameters, optimize, state)
    w_ref = [enc0.copy(), emb0.copy(), mm0.copy()]
    for w, g in zip(w_ref, g_ref):
        A.rmsprop_update(w, g, np.zeros_like(w), lr)
    for got, w1, w0, g, what in zip(T7.split_vector(optimize_parameters, sizes_w), w_ref, (enc0, emb0, mm0), g_ref,
                                    ("encoder", "embedding", "multimodal")):
        big = np.abs(g) > 1e-5            # the first RMSprop step is lr * g / (0.1 |g| + eps): ill-conditioned where |g| ~ eps
        assert_close((got - w0)[big], (w1 - w0)[big], 50 * tol, f"{prec_name} {mode} update of {what}")
    m.close()


def test_axb_forward_twice_bf16x2_uses_fresh_planes():
    """the activation-plane cache must not serve stale planes when a module-level call rewrites qd / vd."""
    import novel_vqa_b200 as nvm
    if nvm.device_count() == 0:
        pytest.fail("needs the B200 box")
    from novel_vqa_b200.torch7_mirror import netdef
    r = np.random.default_rng(5)
    H, n, I, Cc, B = 64, 2, 72, 48, 40
    axb = netdef.AxB(2 * H * n, I, Cc, 0.5, rnn_size=H, rnn_layers=n)
    axb._model.close()
    axb._model = nvm.Arch1Model(nvm.Arch1Config(E=4, H=H, L=n, V=8, I=I, C=Cc, O=4, T=2, B=64), precision=nvm.PREC_BF16X2)
    w, _ = axb.getParameters()
    w[:] = r.uniform(-0.3, 0.3, w.size)
    axb.evaluate()
    mm = A.split_flat(w[:Cc * 2 * H * n + Cc + Cc * I + Cc], [("Wq", (Cc, 2 * H * n)), ("bq", (Cc,)), ("Wv", (Cc, I)), ("bv", (Cc,))])
    for k in range(3):
        q = r.standard_normal((B, 2 * H * n)).astype(np.float32)
        i = r.standard_normal((B, I)).astype(np.float32)
        z = axb.forward([q, i])
        z_ref = np.tanh(q @ mm["Wq"].T + mm["bq"]) * np.tanh(i @ mm["Wv"].T + mm["bv"])
        assert_close(z, z_ref, 1e-4, f"AxB forward call {k}")
