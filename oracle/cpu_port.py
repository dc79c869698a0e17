"""CPU port of the reference's training step, used ONLY as bench.py's cpu_baseline / --impl reference arm.
TEST/BENCH INFRASTRUCTURE -- NOT PRODUCT CODE.

The reference's own implementation (TensorFlow 1.x graph driven from Python 2) cannot run in this image
(no TensorFlow, no network), so kind = "port": the same op sequence TensorFlow's CPU
kernels would execute for session.run([opt_step]) (multimodal_autoencoder.py:590) -- fp32 matmul + bias +
activation per layer (:454-518), the loss (:381-390), reverse-mode gradients, TF-formula Adam (:411) --
on torch's multithreaded CPU kernels, preceded by the reference's per-row NumPy noise loop (:668-702),
which is part of every reference training step (:568-569).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import mmae_oracle as O


class CpuPort:
    def __init__(self, cfg: O.OracleConfig, params, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.cfg = cfg
        self.P = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.t = 0

    def _act(self, z, a=None):
        a = a or self.cfg.activation
        if a == 'relu':
            return torch.relu(z)
        if a == 'tanh':
            return torch.tanh(z)
        if a == 'softsign':
            return torch.nn.functional.softsign(z)
        if a == 'softplus':
            return torch.nn.functional.softplus(z)
        return z

    def loss(self, noisy, X):
        cfg, P = self.cfg, self.P
        h = noisy
        for i in range(cfg.L):
            h = h @ P['weights%d' % i] + P['encode_biases%d' % i]
            if i < cfg.L - 1:
                h = self._act(h)
        for j in range(cfg.L):
            i = cfg.L - 1 - j
            W = P['weights%d' % i].t() if cfg.tie_weights else P['decode_weights%d' % i]
            h = h @ W + P['decode_biases%d' % i]
            if j < cfg.L - 1:
                h = self._act(h)
        if cfg.loss_func == 'mean_squared':
            rec = torch.sqrt(torch.mean((h - X) ** 2))
        else:
            rec = torch.nn.functional.binary_cross_entropy_with_logits(h, X, reduction='sum')
        reg = 0.0
        if cfg.weight_penalty:
            for k, w in P.items():
                if 'weights' in k:
                    reg = reg + (2.0 if (cfg.tie_weights and k.startswith('weights')) else 1.0) * 0.5 * (w ** 2).sum()
        return rec + cfg.weight_penalty * reg, rec

    def step(self, X64, rng=np.random, with_noise=True):
        """One reference training step on a float64 batch (the DataLoader's dtype): noise loop, f64->f32
        feed cast (:351-352), forward, backward, Adam.  Returns (loss, seconds in the noise loop)."""
        t0 = time.perf_counter()
        noisy64 = O.add_noise(self.cfg, X64, rng) if with_noise else X64
        t_noise = time.perf_counter() - t0
        X = torch.from_numpy(np.asarray(X64, np.float32))
        noisy = torch.from_numpy(np.asarray(noisy64, np.float32))
        total, rec = self.loss(noisy, X)
        for p in self.P.values():
            p.grad = None
        total.backward()
        self.t += 1
        c = self.cfg
        a = c.learning_rate * np.sqrt(1 - c.beta2 ** self.t) / (1 - c.beta1 ** self.t)
        with torch.no_grad():
            for k, p in self.P.items():
                if p.grad is None:
                    continue
                g = p.grad
                self.m[k] += (g - self.m[k]) * (1 - c.beta1)
                self.v[k] += (g * g - self.v[k]) * (1 - c.beta2)
                p -= a * self.m[k] / (torch.sqrt(self.v[k]) + c.adam_eps)
        return float(rec.detach()), t_noise

    def forward_only(self, X):
        with torch.no_grad():
            cfg, P = self.cfg, self.P
            h = X
            for i in range(cfg.L):
                h = h @ P['weights%d' % i] + P['encode_biases%d' % i]
                if i < cfg.L - 1:
                    h = self._act(h)
            for j in range(cfg.L):
                i = cfg.L - 1 - j
                W = P['weights%d' % i].t() if cfg.tie_weights else P['decode_weights%d' % i]
                h = h @ W + P['decode_biases%d' % i]
                if j < cfg.L - 1:
                    h = self._act(h)
            return torch.sigmoid(h) if cfg.loss_func == 'sigmoid_cross_entropy' else h

    def predict(self, X64):
        """fill_missing_data_in_file (:1167-1187): predict() over the whole matrix, then the reference's per-row
        fill loop (data_funcs.py:310-381).  Returns (filled, 0.0)."""
        Xbar = self.forward_only(torch.from_numpy(np.asarray(X64, np.float32))).numpy().astype(np.float64)
        return O.fill_missing(self.cfg, np.asarray(X64, np.float64), Xbar), 0.0

    def cls_step(self, X64, Y64, rng=np.random):
        """session.run([classification_opt_step]) (:647) in fp32 on torch's CPU kernels, like step(): noise loop, encoder
        + head forward (:520-540 incl. the :533 activation bound), mean sigmoid-CE / sparse-softmax loss (:431-441),
        reverse-mode gradients, the SECOND Adam (:443: own m / v / t; decoder untouched).
        Returns (loss, seconds in the noise loop)."""
        cfg, P = self.cfg, self.P
        if not hasattr(self, '_m1'):
            self._m1 = {k: torch.zeros_like(v) for k, v in P.items()}
            self._v1 = {k: torch.zeros_like(v) for k, v in P.items()}
            self._t1 = 0
        t0 = time.perf_counter()
        noisy64 = O.add_noise(cfg, X64, rng)
        t_noise = time.perf_counter() - t0
        h = torch.from_numpy(np.asarray(noisy64, np.float32))
        Y = torch.from_numpy(np.asarray(Y64, np.float32))
        for i in range(cfg.L):
            h = h @ P['weights%d' % i] + P['encode_biases%d' % i]
            if i < cfg.L - 1:
                h = self._act(h)
        nh = len(cfg.head_dims())
        for i in range(nh):
            h = h @ P['classification_weights%d' % i] + P['classification_biases%d' % i]
            if i < cfg.L - 1:                                   # :533
                h = self._act(h, cfg.cls_activation)
        if cfg.cls_loss == 'sigmoid_cross_entropy':
            loss = torch.nn.functional.binary_cross_entropy_with_logits(h, Y, reduction='mean')
        else:
            loss = torch.nn.functional.cross_entropy(h, Y.long(), reduction='mean')
        if cfg.cls_weight_penalty:
            loss = loss + cfg.cls_weight_penalty * sum(0.5 * (P['classification_weights%d' % i] ** 2).sum() for i in range(nh))
        for p in P.values():
            p.grad = None
        loss.backward()
        self._t1 += 1
        a = cfg.cls_learning_rate * np.sqrt(1 - cfg.beta2 ** self._t1) / (1 - cfg.beta1 ** self._t1)
        with torch.no_grad():
            for k, p in P.items():
                if p.grad is None:
                    continue
                g = p.grad
                self._m1[k] += (g - self._m1[k]) * (1 - cfg.beta1)
                self._v1[k] += (g * g - self._v1[k]) * (1 - cfg.beta2)
                p -= a * self._m1[k] / (torch.sqrt(self._v1[k]) + cfg.adam_eps)
        return float(loss.detach()), t_noise
