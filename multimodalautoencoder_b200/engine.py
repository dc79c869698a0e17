"""Thin Python handle over the C ABI (include/mmae_b200.h).  PyTorch supplies device buffers and
the current stream; all arithmetic happens in libmmae_b200.so."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _capi as capi
from .noise import DEFAULT_NOISE_P, DEFAULT_NOISE_TYPES, categorical_thresholds, type_masks_from_names


class EngineError(RuntimeError):
    pass


@dataclass
class EngineConfig:
    num_feats: int
    layer_sizes: List[int]
    modality_starts: List[int]
    modality_names: List[str]
    tie_weights: bool = True
    variational: bool = False
    activation: str = 'softsign'
    loss_func: str = 'sigmoid_cross_entropy'
    weight_penalty: float = 0.0
    learning_rate: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    cls_layer_sizes: Optional[List[int]] = None
    num_labels: Optional[int] = 3
    cls_activation: str = 'softsign'
    cls_loss: str = 'sigmoid_cross_entropy'
    cls_weight_penalty: float = 0.0
    cls_learning_rate: float = 1e-4
    mask_with: float = -1.0
    intelligent_noise: bool = True
    num_modalities_to_drop: int = 1
    noise_p: List[float] = field(default_factory=lambda: list(DEFAULT_NOISE_P))
    noise_types: List[List[str]] = field(default_factory=lambda: [list(t) for t in DEFAULT_NOISE_TYPES])
    seed: int = 0
    precision: str = 'tf32'
    max_batch: int = 256
    classifier_only: bool = False        # comparison_algorithms/neural_net.py: plain MLP classifier (every hidden layer activated)
    clip_norm: float = 0.0               # > 0: tf.clip_by_global_norm(gradients, clip_norm) in the head optimizer

    def head_widths(self):
        if self.cls_layer_sizes is None:
            return []
        return list(self.cls_layer_sizes) + [self.num_labels if self.num_labels is not None else 2]


def _arr(ctype, values):
    return (ctype * max(len(values), 1))(*values)


class Engine:
    def __init__(self, cfg: EngineConfig, device=None):
        import torch
        self._torch = torch
        self._h = None
        if not torch.cuda.is_available():
            raise EngineError('the MMAE engine needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = capi.load()
        self.cfg = cfg
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        act = lambda n: capi.ACT.get(n, 0)                       # unknown names are linear (:497)
        heads = cfg.head_widths()
        self._keep = dict(
            starts=_arr(C.c_int32, list(cfg.modality_starts)), layers=_arr(C.c_int32, list(cfg.layer_sizes)),
            heads=_arr(C.c_int32, heads))
        c = capi.Config()
        c.num_feats = cfg.num_feats
        c.num_modalities = len(cfg.modality_starts) - 1
        c.modality_starts = self._keep['starts']
        c.num_layers = len(cfg.layer_sizes)
        c.layer_sizes = self._keep['layers']
        c.tie_weights = int(cfg.tie_weights)
        c.variational = int(cfg.variational)
        c.activation = act(cfg.activation)
        c.loss_func = capi.LOSS[cfg.loss_func]
        c.weight_penalty, c.learning_rate = cfg.weight_penalty, cfg.learning_rate
        c.beta1, c.beta2, c.adam_eps = cfg.beta1, cfg.beta2, cfg.adam_eps
        c.num_head_layers = len(heads)
        c.head_sizes = self._keep['heads']
        c.head_activation = act(cfg.cls_activation)
        c.head_loss = capi.HEAD_LOSS['sigmoid_cross_entropy' if cfg.cls_loss == 'sigmoid_cross_entropy' else 'softmax']
        c.head_weight_penalty, c.head_learning_rate = cfg.cls_weight_penalty, cfg.cls_learning_rate
        c.mask_with = cfg.mask_with
        c.n_zero = int(cfg.num_feats * .05)
        if cfg.intelligent_noise:
            masks = type_masks_from_names(cfg.noise_types, list(cfg.modality_names))
            thr = [int(t) for t in categorical_thresholds(cfg.noise_p)]
            self._keep['masks'] = _arr(C.c_uint32, masks)
            self._keep['thr'] = _arr(C.c_uint32, thr)
            c.noise_mode = capi.NOISE_INTELLIGENT
            c.num_noise_types = len(masks)
            c.noise_type_masks = self._keep['masks']
            c.noise_thresholds = self._keep['thr']
            self.type_masks = masks
        else:
            c.noise_mode = capi.NOISE_UNIFORM
            c.num_noise_types = 0
            self.type_masks = None
        c.num_modalities_to_drop = cfg.num_modalities_to_drop
        c.seed = cfg.seed
        c.precision = capi.PREC[cfg.precision]
        c.max_batch = cfg.max_batch
        c.classifier_only = int(cfg.classifier_only)
        c.clip_norm = float(cfg.clip_norm)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.mmae_create(C.byref(c), C.byref(h))
        if rc != 0:
            raise ValueError('mmae_create failed (%d): %s' % (rc, self.lib.mmae_last_error(None).decode()))
        self._h = h
        self._names = None
        self.use_current_stream()

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.mmae_last_error(self._h).decode()
            raise (ValueError if rc in (-1, -3) else EngineError)('libmmae_b200 error %d: %s' % (rc, msg))

    def close(self):
        if getattr(self, '_h', None):
            self.lib.mmae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_current_stream(self):
        s = self._torch.cuda.current_stream(self.device).cuda_stream
        self._ck(self.lib.mmae_set_stream(self._h, C.c_void_p(s)))

    def synchronize(self):
        self._ck(self.lib.mmae_synchronize(self._h))

    def _dev(self, t, dtype=None):
        torch = self._torch
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float32), device=self.device)
        if t.device != self.device or t.dtype != (dtype or torch.float32) or not t.is_contiguous():
            t = t.to(device=self.device, dtype=dtype or torch.float32).contiguous()
        return t

    def _check_labels(self, Y, B):
        """The head reads B*C floats for a sigmoid head and B class indices for a softmax head (TensorFlow raises a
        shape error for anything else; here a wrong shape would be an out-of-bounds device read)."""
        heads = self.cfg.head_widths()
        if not heads:
            return                      # the C call reports the missing head (MMAE_ERR_STATE -> EngineError)
        Cn = heads[-1]
        n = int(np.prod(tuple(Y.shape) if hasattr(Y, 'shape') else np.shape(Y)))
        if self.cfg.cls_loss == 'sigmoid_cross_entropy':
            if n != B * Cn:
                raise ValueError('labels have %d values, the sigmoid head needs [%d, %d]' % (n, B, Cn))
        else:
            if n != B:
                raise ValueError('labels have %d values, the softmax head needs one class index per row (%d)' % (n, B))
            if not isinstance(Y, self._torch.Tensor):
                yi = np.asarray(Y)
                if yi.size and (yi.min() < 0 or yi.max() >= Cn):
                    raise ValueError('class indices must lie in [0, %d)' % Cn)

    # ------------------------------------------------------------------ variables
    def variables(self):
        if self._names is None:
            out = []
            buf = C.create_string_buffer(64)
            r, c = C.c_int64(), C.c_int64()
            for i in range(self.lib.mmae_num_variables(self._h)):
                self._ck(self.lib.mmae_variable_info(self._h, i, buf, 64, C.byref(r), C.byref(c)))
                out.append((buf.value.decode(), (r.value, c.value) if c.value > 0 else (r.value,)))
            self._names = out
        return self._names

    def shape_of(self, name):
        for n, s in self.variables():
            if n == name:
                return s
        raise ValueError('unknown variable ' + name)

    def set_variable(self, name, value):
        a = np.ascontiguousarray(value, dtype=np.float32)
        if tuple(a.shape) != tuple(self.shape_of(name)):
            raise ValueError('shape mismatch for %s: %s vs %s' % (name, a.shape, self.shape_of(name)))
        self._ck(self.lib.mmae_set_variable(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))

    def _get(self, fn, name):
        a = np.empty(self.shape_of(name), np.float32)
        self._ck(fn(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))
        return a

    def get_variable(self, name):
        return self._get(self.lib.mmae_get_variable, name)

    def get_gradient(self, name):
        return self._get(self.lib.mmae_get_gradient, name)

    def set_params(self, params):
        for k, v in params.items():
            self.set_variable(k, v)

    def get_params(self):
        return {n: self.get_variable(n) for n, _ in self.variables()}

    def get_opt_state(self, optimizer, name):
        m = np.empty(self.shape_of(name), np.float32)
        v = np.empty_like(m)
        t = C.c_int64()
        self._ck(self.lib.mmae_get_opt_state(self._h, optimizer, name.encode(), m.ctypes.data_as(C.c_void_p),
                                             v.ctypes.data_as(C.c_void_p), m.size, C.byref(t)))
        return m, v, t.value

    def set_opt_state(self, optimizer, name, m, v, t):
        m = np.ascontiguousarray(m, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        self._ck(self.lib.mmae_set_opt_state(self._h, optimizer, name.encode(), m.ctypes.data_as(C.c_void_p),
                                             v.ctypes.data_as(C.c_void_p), m.size, int(t)))

    # ------------------------------------------------------------------ noise
    def set_rng_step(self, step):
        self._ck(self.lib.mmae_set_rng_step(self._h, int(step)))

    def set_noise(self, zero_bits, mod_bits):
        z = np.ascontiguousarray(zero_bits, np.uint32)
        m = np.ascontiguousarray(mod_bits, np.uint32)
        self._ck(self.lib.mmae_set_noise(self._h, z.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), len(m)))

    def gen_noise(self, batch, first_row=0):
        self._ck(self.lib.mmae_gen_noise(self._h, int(batch), int(first_row)))

    def get_noise(self, batch):
        zw = (self.cfg.num_feats + 31) // 32
        z = np.empty((batch, zw), np.uint32)
        m = np.empty(batch, np.uint32)
        self._ck(self.lib.mmae_get_noise(self._h, z.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), batch))
        return z, m

    def apply_noise(self, X):
        X = self._dev(X)
        out = self._torch.empty_like(X)
        self._ck(self.lib.mmae_apply_noise(self._h, C.c_void_p(X.data_ptr()), X.shape[0], C.c_void_p(out.data_ptr())))
        return out

    # ------------------------------------------------------------------ forward / train
    def forward(self, X, target=None, labels=None, noise=False, keep=1.0, recon=False, embedding=False,
                head=False, loss=False, filled=False, head_loss=False):
        torch = self._torch
        X = self._dev(X)
        B = X.shape[0]
        tgt = None if target is None else self._dev(target)
        if labels is not None:
            self._check_labels(labels, B)
        lab = None if labels is None else self._dev(labels)
        want = 0
        o = capi.Outputs()
        res = {}
        E = self.cfg.layer_sizes[-1]
        if recon:
            want |= capi.WANT_RECON
            res['recon'] = torch.empty((B, self.cfg.num_feats), device=self.device)
            o.recon = res['recon'].data_ptr()
        if embedding:
            want |= capi.WANT_EMBEDDING
            res['embedding'] = torch.empty((B, E), device=self.device)
            o.embedding = res['embedding'].data_ptr()
        if head or head_loss:
            want |= capi.WANT_HEAD
            Cn = self.cfg.head_widths()[-1]
            res['logits'] = torch.empty((B, Cn), device=self.device)
            res['probs'] = torch.empty((B, Cn), device=self.device)
            pshape = (B, Cn) if self.cfg.cls_loss == 'sigmoid_cross_entropy' else (B,)
            res['preds'] = torch.empty(pshape, device=self.device, dtype=torch.int32)
            o.logits, o.probs, o.preds = res['logits'].data_ptr(), res['probs'].data_ptr(), res['preds'].data_ptr()
        if loss:
            want |= capi.WANT_LOSS
        if filled:
            want |= capi.WANT_FILLED
            res['filled'] = torch.empty((B, self.cfg.num_feats), device=self.device)
            o.filled = res['filled'].data_ptr()
        if head_loss:
            want |= capi.WANT_HEAD_LOSS
        self._ck(self.lib.mmae_forward(self._h, C.c_void_p(X.data_ptr()),
                                       C.c_void_p(tgt.data_ptr()) if tgt is not None else None,
                                       C.c_void_p(lab.data_ptr()) if lab is not None else None,
                                       B, int(bool(noise)), float(keep), want, C.byref(o)))
        res['_keepalive'] = (X, tgt, lab)
        return res

    def forward_into(self, X, recon=None, filled=None, embedding=None, loss=False, target=None):
        """forward() writing into caller-owned device tensors (no allocation inside the call)."""
        X = self._dev(X)
        want = 0
        o = capi.Outputs()
        if recon is not None:
            want |= capi.WANT_RECON
            o.recon = recon.data_ptr()
        if filled is not None:
            want |= capi.WANT_FILLED
            o.filled = filled.data_ptr()
        if embedding is not None:
            want |= capi.WANT_EMBEDDING
            o.embedding = embedding.data_ptr()
        if loss:
            want |= capi.WANT_LOSS
        tgt = None if target is None else self._dev(target)
        self._ck(self.lib.mmae_forward(self._h, C.c_void_p(X.data_ptr()),
                                       C.c_void_p(tgt.data_ptr()) if tgt is not None else None, None,
                                       X.shape[0], 0, 1.0, want, C.byref(o)))

    @staticmethod
    def _noise_flag(noise):
        """False: no noise;  True: apply the descriptor loaded by set_noise / gen_noise;  'gen': draw this step's Philox
        descriptor inside the call (one kernel draws it and writes the noisy batch)."""
        return 3 if noise == 'gen' else int(bool(noise))

    def train_step(self, X, noise=False, keep=1.0):
        X = self._dev(X)
        self._ck(self.lib.mmae_train_step(self._h, C.c_void_p(X.data_ptr()), X.shape[0], self._noise_flag(noise), float(keep)))

    def train_step_pair(self, X_in, target, noise=False, keep=1.0):
        X_in = self._dev(X_in)
        target = self._dev(target)
        self._ck(self.lib.mmae_train_step_pair(self._h, C.c_void_p(X_in.data_ptr()), C.c_void_p(target.data_ptr()),
                                               X_in.shape[0], self._noise_flag(noise), float(keep)))

    def cls_train_step(self, X, Y, noise=False, keep=1.0):
        X = self._dev(X)
        self._check_labels(Y, X.shape[0])
        Y = self._dev(Y)
        self._ck(self.lib.mmae_cls_train_step(self._h, C.c_void_p(X.data_ptr()), C.c_void_p(Y.data_ptr()),
                                              X.shape[0], self._noise_flag(noise), float(keep)))

    def train_step_host(self, X_host, gen_noise=False, keep=1.0, use_noise=False):
        """X_host: C-contiguous float32 ndarray or pinned CPU tensor (kept alive by the caller until synchronize()).
        gen_noise: Philox noise drawn on the device; use_noise: apply the descriptor last given to set_noise."""
        ptr, B = self._host_ptr(X_host)
        self._ck(self.lib.mmae_train_step_host(self._h, ptr, B, 1 if gen_noise else (2 if use_noise else 0), float(keep)))

    def cls_train_step_host(self, X_host, Y_host, gen_noise=False, keep=1.0, use_noise=False):
        ptr, B = self._host_ptr(X_host)
        self._check_labels(Y_host, B)
        yptr, _ = self._host_ptr(Y_host)
        self._ck(self.lib.mmae_cls_train_step_host(self._h, ptr, yptr, B, 1 if gen_noise else (2 if use_noise else 0), float(keep)))

    def _host_ptr(self, a):
        torch = self._torch
        if isinstance(a, torch.Tensor):
            assert a.device.type == 'cpu' and a.dtype == torch.float32 and a.is_contiguous()
            return C.c_void_p(a.data_ptr()), a.shape[0]
        assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
        return a.ctypes.data_as(C.c_void_p), a.shape[0]

    def forward_host(self, X_host, target_host=None, labels_host=None, noise=False, keep=1.0, recon=False,
                     embedding=False, head=False, loss=False, filled=False, head_loss=False, out=None):
        """predict()-style call: NumPy in, NumPy out, copies inside the C call.  `out` may map 'recon' / 'filled' /
        'embedding' to preallocated C-contiguous float32 arrays (e.g. views of pinned torch tensors): batches above
        131 072 rows then stream through an H2D / compute / D2H pipeline at PCIe speed."""
        X = np.ascontiguousarray(X_host, np.float32)
        out = out or {}
        B = X.shape[0]
        tgt = None if target_host is None else np.ascontiguousarray(target_host, np.float32)
        lab = None if labels_host is None else np.ascontiguousarray(labels_host, np.float32)
        if lab is not None:
            self._check_labels(lab, B)
        o = capi.Outputs()
        res = {}
        want = 0
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if recon:
            want |= capi.WANT_RECON
            res['recon'] = out.get('recon') if out.get('recon') is not None else np.empty((B, self.cfg.num_feats), np.float32)
            o.recon = res['recon'].ctypes.data
        if embedding:
            want |= capi.WANT_EMBEDDING
            res['embedding'] = (out.get('embedding') if out.get('embedding') is not None
                                else np.empty((B, self.cfg.layer_sizes[-1]), np.float32))
            o.embedding = res['embedding'].ctypes.data
        if head or head_loss:
            want |= capi.WANT_HEAD
            Cn = self.cfg.head_widths()[-1]
            res['logits'] = np.empty((B, Cn), np.float32)
            res['probs'] = np.empty((B, Cn), np.float32)
            res['preds'] = np.empty((B, Cn) if self.cfg.cls_loss == 'sigmoid_cross_entropy' else (B,), np.int32)
            o.logits, o.probs, o.preds = res['logits'].ctypes.data, res['probs'].ctypes.data, res['preds'].ctypes.data
        if loss:
            want |= capi.WANT_LOSS
        if filled:
            want |= capi.WANT_FILLED
            res['filled'] = out.get('filled') if out.get('filled') is not None else np.empty((B, self.cfg.num_feats), np.float32)
            o.filled = res['filled'].ctypes.data
        if head_loss:
            want |= capi.WANT_HEAD_LOSS
        self._ck(self.lib.mmae_forward_host(self._h, p(X), p(tgt) if tgt is not None else None,
                                            p(lab) if lab is not None else None, B, int(bool(noise)), float(keep),
                                            want, C.byref(o)))
        return res

    def backward(self, X, noise=False, keep=1.0, global_batch=0):
        X = self._dev(X)
        self._ck(self.lib.mmae_backward(self._h, C.c_void_p(X.data_ptr()), X.shape[0], int(global_batch),
                                        int(bool(noise)), float(keep)))

    def grad_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.mmae_grad_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def apply_update(self, optimizer=0):
        self._ck(self.lib.mmae_apply_update(self._h, optimizer))

    def set_dataset(self, slot, X, Y=None):
        X = np.ascontiguousarray(X, np.float32)
        yc = 0
        yp = None
        if Y is not None:
            Y = np.ascontiguousarray(Y, np.float32)
            self._check_labels(Y, X.shape[0])
            yc = 1 if Y.ndim == 1 else Y.shape[1]
            yp = Y.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.mmae_set_dataset(self._h, slot, X.ctypes.data_as(C.c_void_p), yp, X.shape[0], yc))

    def set_dataset_device(self, slot, X, Y=None):
        """set_dataset from torch CUDA tensors (copied device-to-device into engine-owned buffers)."""
        X = self._dev(X)
        yc, yp = 0, None
        if Y is not None:
            self._check_labels(Y, X.shape[0])
            Y = self._dev(Y)
            yc = 1 if Y.dim() == 1 else Y.shape[1]
            yp = C.c_void_p(Y.data_ptr())
        self._ck(self.lib.mmae_set_dataset_device(self._h, slot, C.c_void_p(X.data_ptr()), yp, X.shape[0], yc))

    def set_dataset_view(self, slot, rows=None):
        """Training view of the resident dataset in `slot`: the dataset rows the current cross-validation fold trains on
        (None: all rows).  Sampled / given indices address the view."""
        if rows is None:
            self._ck(self.lib.mmae_set_dataset_view(self._h, slot, None, 0))
            return
        r = np.ascontiguousarray(rows, np.int64)
        self._ck(self.lib.mmae_set_dataset_view(self._h, slot, r.ctypes.data_as(C.c_void_p), r.size))

    def train_step_resident(self, slot, batch, idx=None, gen_noise=True, keep=1.0, classification=False):
        ip = None
        if idx is not None:
            idx = np.ascontiguousarray(idx, np.int64)
            ip = idx.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.mmae_train_step_resident(self._h, slot, ip, int(batch), int(bool(gen_noise)), float(keep),
                                                   int(bool(classification))))

    def eval_resident(self, slot, batch, gen_noise=True, keep=1.0):
        """reconstruction_loss of a batch sampled on the device from the resident dataset (no update); read it with scalars()."""
        self._ck(self.lib.mmae_eval_resident(self._h, slot, int(batch), int(bool(gen_noise)), float(keep)))

    def modality_rmse(self, X_host):
        """get_reconstruction_loss_per_modality (:1189-1216) in one batched device pass: list of M RMSE values."""
        X = np.ascontiguousarray(X_host, np.float32)
        M = len(self.cfg.modality_starts) - 1
        out = (C.c_double * M)()
        self._ck(self.lib.mmae_modality_rmse(self._h, X.ctypes.data_as(C.c_void_p), X.shape[0], out))
        return [float(out[i]) for i in range(M)]

    def scalars(self):
        a = (C.c_double * capi.NUM_SCALARS)()
        self._ck(self.lib.mmae_read_scalars(self._h, a, capi.NUM_SCALARS))
        return {'recon_loss': a[0], 'kl_mean': a[1], 'sumsq': a[2], 'head_loss': a[3], 'head_acc': a[4],
                'grad_scale': a[5]}

    def get_buffer(self, name, shape):
        a = np.empty(shape, np.float32)
        self._ck(self.lib.mmae_get_buffer(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))
        return a

    def set_eps(self, eps):
        if eps is None:
            self._ck(self.lib.mmae_set_eps(self._h, None, 0))
        else:
            e = np.ascontiguousarray(eps, np.float32)
            self._ck(self.lib.mmae_set_eps(self._h, e.ctypes.data_as(C.c_void_p), e.size))

    def set_shard(self, global_batch, first_row):
        self._ck(self.lib.mmae_set_shard(self._h, int(global_batch), int(first_row)))

    def comm_init(self, unique_id: bytes, rank, world):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.mmae_comm_init(self._h, buf, rank, world))

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = capi.load()
        buf = C.create_string_buffer(128)
        rc = lib.mmae_comm_unique_id(buf)
        if rc != 0:
            raise EngineError('mmae_comm_unique_id failed: ' + lib.mmae_last_error(None).decode())
        return buf.raw

    def set_profiling(self, on):
        self._ck(self.lib.mmae_set_profiling(self._h, int(bool(on))))

    def read_profile(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self.lib.mmae_read_profile(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
        return {'gemm_ms': ms.value, 'gemm_flops': fl.value, 'gemm_launches': n.value}

    def read_scalars_async(self, pinned):
        """pinned: torch pinned CPU float64 tensor with >= 8 elements."""
        self._ck(self.lib.mmae_read_scalars_async(self._h, C.c_void_p(pinned.data_ptr()), capi.NUM_SCALARS))

    @property
    def kernel_launches(self):
        return self.lib.mmae_kernel_launches(self._h)

    @property
    def fused_noise_launches(self):
        """GEMM launches that applied the block mask + noise while loading their A operand (no noisy copy of X)."""
        return self.lib.mmae_fused_noise_launches(self._h)

    @property
    def graph_replays(self):
        """Train steps replayed from a captured CUDA graph."""
        return self.lib.mmae_graph_replays(self._h)

    @property
    def wgrad_group_launches(self):
        """How many of the launches were the grouped weight-gradient kernel (every dW of a step in one launch)."""
        return self.lib.mmae_wgrad_group_launches(self._h)

    @property
    def backward_chain_launches(self):
        """How many of the launches were the whole-backward (all dgrads of a step) kernel."""
        return self.lib.mmae_backward_chain_launches(self._h)

    @property
    def chain_launches(self):
        """How many of the launches were the whole-network (encode + decode + loss) kernel."""
        return self.lib.mmae_chain_launches(self._h)


def debug_gemm(A, B, transA=False, transB=False, bias=None, activation='linear', precision='tf32', C_init=None, beta=0.0):
    """C = op(A) op(B) through one of the engine's GEMM families (tests / bench only)."""
    import torch
    lib = capi.load()
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = B.shape[0] if transB else B.shape[1]
    out = torch.zeros((M, N), device=A.device, dtype=torch.float32) if C_init is None else C_init.clone()
    rc = lib.mmae_debug_gemm(capi.PREC[precision], int(transA), int(transB), M, N, K, C.c_void_p(A.data_ptr()),
                             A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0), C.c_void_p(out.data_ptr()), out.stride(0),
                             C.c_void_p(bias.data_ptr()) if bias is not None else None, capi.ACT[activation], float(beta),
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise EngineError('mmae_debug_gemm failed: %d' % rc)
    return out
