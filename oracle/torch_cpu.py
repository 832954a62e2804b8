"""PyTorch-CPU (MKL, all host threads) restatement of the reference's arch1 training step, used ONLY as the
timed CPU baseline of bench.py (``cpu_baseline`` and ``--impl reference``).  TEST/BENCH INFRASTRUCTURE --
see oracle/__init__.py.  Torch7 itself cannot run here, so this executes, op for op, what
002_train_baseline.lua:272-335,408 makes Torch7's CPU backend execute ("faithful" variant):

  * the embedding is nn.Linear(V,E) applied to a DENSE one-hot [N x V] matrix that is rebuilt every
    step (misc/RNNUtils.lua:42-47,123), including the backward into the one-hot input;
  * the LSTM runs as T un-shared clones: the flat weights are copied into every clone each step
    (:275-280) and the per-clone gradients are summed afterwards (:323-326);
  * i2h / h2h are separate Linear modules per layer and step, gate math is un-fused (misc/LSTM.lua:41-59);
  * the AxB backward also produces d fc7 (misc/netdef.lua:6-14);
  * clamp, then a 6-pass RMSprop over the joined parameter vector (misc/rmsprop_lrscale.lua:26-34).

``variant="gather"`` replaces only the one-hot Linear by an index (isolates the model cost).
Numerics are checked against oracle/arch1.py in tests/test_oracle.py::test_torch_cpu_port_matches_oracle.
"""
import time

import numpy as np
import torch

from . import arch1 as A


def _split(w, layout):
    out, off = {}, 0
    for name, shape in layout:
        n = int(np.prod(shape))
        out[name] = w[off:off + n].view(*shape)
        off += n
    return out


class TorchCpuArch1:
    def __init__(self, cfg: A.Arch1Config, enc, emb, mm, variant="faithful"):
        self.cfg, self.variant = cfg, variant
        self.enc = torch.tensor(np.asarray(enc, dtype=np.float32))
        self.emb = torch.tensor(np.asarray(emb, dtype=np.float32))
        self.mm = torch.tensor(np.asarray(mm, dtype=np.float32))
        self.x = torch.cat([self.enc, self.emb, self.mm])            # optimize.winit  (:190)
        self.m = None
        self.clone_w = [torch.empty_like(self.enc) for _ in range(cfg.T)]   # dupe_rnn (:269)

    def jdj(self, q_ra, lengths, fv_im, labels, seed=None):
        cfg = self.cfg
        H = cfg.H
        n_enc, n_emb = cfg.n_enc, cfg.n_emb
        # split_vector + copies into the nets and all clones (:273-286)
        self.enc.copy_(self.x[:n_enc])
        self.emb.copy_(self.x[n_enc:n_enc + n_emb])
        self.mm.copy_(self.x[n_enc + n_emb:])
        words, sizes, sort_index, inv = A.sort_encoding_right_align(q_ra, lengths)
        Lq = len(sizes)
        clones = []
        for t in range(Lq):
            self.clone_w[t].copy_(self.enc)
            clones.append(self.clone_w[t].clone().requires_grad_(True))
        emb_w = self.emb.clone().requires_grad_(True)
        mm_w = self.mm.clone().requires_grad_(True)
        B = q_ra.shape[0]
        masks = None if seed is None else A.build_masks(cfg, seed, B, sizes, sort_index)
        tm = (lambda a: torch.from_numpy(np.ascontiguousarray(a))) if masks is not None else None
        ep = _split(emb_w, cfg.emb_layout())
        wt = torch.from_numpy(words - 1)
        if self.variant == "faithful":
            onehot = torch.zeros(len(words), cfg.V)
            onehot.scatter_(1, wt.view(-1, 1), 1.0)
            onehot.requires_grad_(True)                            # nn.Linear:updateGradInput always runs
            pre = torch.nn.functional.linear(onehot, ep["We"], ep["be"])
        else:
            pre = ep["We"].t()[wt] + ep["be"]
        if masks is not None:
            pre = pre * tm(masks["emb"])
        y = torch.tanh(pre)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        state = torch.zeros(int(sizes[0]), cfg.S)
        for t in range(Lq):
            if t > 0 and sizes[t] > sizes[t - 1]:
                state = torch.cat([state, torch.zeros(int(sizes[t] - sizes[t - 1]), cfg.S)], 0)
            p = _split(clones[t], cfg.enc_layout())
            inp = y[offs[t]:offs[t + 1]]
            outs = []
            for l in range(cfg.L):
                c_prev = state[:, 2 * l * H:(2 * l + 1) * H]
                h_prev = state[:, (2 * l + 1) * H:(2 * l + 2) * H]
                if l > 0:
                    inp = outs[-1]
                    if masks is not None:
                        inp = inp * tm(masks["lstm"][t][l - 1])
                i2h = torch.nn.functional.linear(inp, p[f"Wi{l}"], p[f"bi{l}"])
                h2h = torch.nn.functional.linear(h_prev, p[f"Wh{l}"], p[f"bh{l}"])
                a = i2h + h2h
                sg = torch.sigmoid(a[:, :3 * H])
                g = torch.tanh(a[:, 3 * H:])
                c = sg[:, H:2 * H] * c_prev + sg[:, :H] * g
                h = sg[:, 2 * H:3 * H] * torch.tanh(c)
                outs += [c, h]
            state = torch.cat(outs, 1)
        tv_q = state[torch.from_numpy(inv)]
        fv = torch.from_numpy(np.ascontiguousarray(fv_im, dtype=np.float32)).requires_grad_(self.variant == "faithful")
        mp = _split(mm_w, cfg.mm_layout())
        qd = tv_q if masks is None else tv_q * tm(masks["q"])
        vd = fv if masks is None else fv * tm(masks["i"])
        z = torch.tanh(torch.nn.functional.linear(qd, mp["Wq"], mp["bq"])) * \
            torch.tanh(torch.nn.functional.linear(vd, mp["Wv"], mp["bv"]))
        if masks is not None:
            z = z * tm(masks["z"])
        scores = torch.nn.functional.linear(z, mp["Wc"], mp["bc"])
        f = torch.nn.functional.cross_entropy(scores, torch.from_numpy(np.asarray(labels, dtype=np.int64) - 1))
        f.backward()
        g_enc = torch.zeros_like(self.enc)
        for t in range(Lq):
            g_enc = g_enc + clones[t].grad                         # :323-326
        grads = torch.cat([g_enc, emb_w.grad, mm_w.grad]).clamp_(-10, 10)   # :328-329
        return float(f.detach()), grads, scores.detach()

    def step(self, batch, lr, seed=None):
        """optim.rmsprop(JdJ, x, ...) as the separate passes of misc/rmsprop_lrscale.lua:26-34."""
        f, g, _ = self.jdj(*batch, seed=seed)
        if self.m is None:
            self.m = torch.zeros_like(self.x)
        self.m.mul_(0.99)
        self.m.addcmul_(g, g, value=1.0 - 0.99)
        tmp = self.m.sqrt().add_(1e-8)
        self.x.add_(torch.div(g, tmp).mul_(-lr))
        return f


def time_steps(cfg, B, steps, warmup, variant="faithful", seed=123, threads=None):
    """Wall-clock samples/s of the CPU restatement on synthetic data of the BASELINE shape."""
    if threads:
        torch.set_num_threads(threads)
    r = np.random.default_rng(seed)
    enc = r.uniform(-0.08, 0.08, cfg.n_enc).astype(np.float32)
    emb = r.uniform(-0.08, 0.08, cfg.n_emb).astype(np.float32)
    mm = r.uniform(-0.08, 0.08, cfg.n_mm).astype(np.float32)
    q = r.integers(1, cfg.V + 1, (B, cfg.T)).astype(np.int64)
    lengths = np.full(B, cfg.T, dtype=np.int64)
    fv = A.l2_normalize_rows(np.maximum(0, r.standard_normal((B, cfg.I))).astype(np.float32))
    labels = r.integers(1, cfg.O + 1, B)
    model = TorchCpuArch1(cfg, enc, emb, mm, variant)
    lr = 3e-4
    for i in range(warmup):
        model.step((q, lengths, fv, labels), lr, seed=seed + i)
    t0 = time.perf_counter()
    for i in range(steps):
        model.step((q, lengths, fv, labels), lr, seed=seed + warmup + i)
    dt = time.perf_counter() - t0
    return {"samples_per_s": B * steps / dt, "s_per_step": dt / steps, "threads": torch.get_num_threads(),
            "variant": variant, "B": B, "steps": steps}
