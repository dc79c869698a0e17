"""MultimodalAutoencoder -- the reference's model class, running on the B200 engine instead of TensorFlow.

Drop-in for /root/reference/multimodal_autoencoder.py (class MultimodalAutoencoder): same constructor
keywords and defaults (:59-70), same public methods and attributes, same training protocol
(:549-647).  Everything the reference did inside tf.Session.run happens in libmmae_b200.so; this file is
host glue: hyper-parameter bookkeeping, the RNG call order of the reference's training loop, loss-curve
lists, checkpoints, and a tiny `session.run` shim for callers that reached into the TensorFlow handles
(autoencoder_wrapper.py:212-226).

Extra keyword-only arguments (defaults preserve reference behaviour):
  rng_mode   'numpy'  -- batches and block-mask noise are drawn on the host from NumPy's global RandomState
                         in the reference's call order (bit-identical masks under the same np.random.seed);
             'philox' -- drawn on the device from a counter-based generator (no per-step host work).
  precision  'tf32' (tcgen05 tensor cores where shapes allow) or 'fp32' (CUDA-core FMA).
  seed       Philox key and weight-initialisation seed.
"""
from __future__ import annotations

import copy
import os

import numpy as np

from . import data_funcs
from .engine import Engine, EngineConfig
from .noise import (DEFAULT_NOISE_P, DEFAULT_NOISE_TYPES, apply_descriptor_host, numpy_descriptor,
                    type_masks_from_names)

DEFAULT_MAIN_DIRECTORY = '/Your/path/here/'


class _Handle:
    """Stand-in for a TensorFlow tensor / placeholder / op handle."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return '<mmae handle %s>' % self.name


class _Session:
    """session.run(fetches, feed_dict) over the handles the reference exposed (12 call shapes)."""

    def __init__(self, model):
        self._m = model

    def run(self, fetches, feed_dict=None):
        m = self._m
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        feed = feed_dict or {}
        X = feed.get(m.noisy_X)
        T = feed.get(m.true_X)
        Y = feed.get(m.true_Y) if m.true_Y is not None else None
        keep = float(feed.get(m.tf_dropout_prob, 1.0))
        names = [f.name for f in flist]
        if names == ['init']:
            return None if single else [None]
        out = {}
        if 'opt_step' in names:
            m.engine.train_step_pair(np.asarray(X, np.float32), np.asarray(T, np.float32), noise=False, keep=keep)
            out['opt_step'] = None
        if 'classification_opt_step' in names:
            m.engine.cls_train_step(np.asarray(X, np.float32), np.asarray(Y, np.float32), noise=False, keep=keep)
            out['classification_opt_step'] = None
        need = [n for n in names if n not in out and n != 'global_step']
        if need:
            want_head = any(n in ('predictions', 'class_probabilities', 'logits', 'classification_loss', 'accuracy') for n in need)
            want_hl = any(n in ('classification_loss', 'accuracy') for n in need)
            want_loss = 'reconstruction_loss' in need
            r = m.engine.forward_host(np.asarray(X, np.float32),
                                      target_host=None if T is None else np.asarray(T, np.float32),
                                      labels_host=None if (Y is None or not want_hl) else np.asarray(Y, np.float32),
                                      keep=keep, recon='decoded_X' in need, embedding='embedding' in need,
                                      head=want_head, loss=want_loss, head_loss=want_hl)
            sc = m.engine.scalars() if (want_loss or want_hl) else {}
            out.update(decoded_X=r.get('recon'), embedding=r.get('embedding'), predictions=r.get('preds'),
                       class_probabilities=r.get('probs'), logits=r.get('logits'),
                       reconstruction_loss=np.float32(sc.get('recon_loss', np.nan)),
                       classification_loss=np.float32(m._head_total_loss(sc)) if want_hl else None,
                       accuracy=np.float32(sc.get('head_acc', np.nan)))
        out['global_step'] = 0            # never incremented in the reference either (:356 vs :411)
        res = [out[n] for n in names]
        return res[0] if single else res


class MultimodalAutoencoder:
    def __init__(self, filename=None, layer_sizes=[128, 64, 32], variational=True, tie_weights=True, batch_size=10,
                 learning_rate=.0001, dropout_prob=1.0, weight_penalty=0.0, activation_func='softsign',
                 loss_func='sigmoid_cross_entropy', decay=True, decay_steps=1000, decay_rate=0.95,
                 clip_gradients=True, classification_layer_sizes=None, classification_filename=None,
                 weight_initialization='xavier', normalization='between_0_and_1', intelligent_noise=True,
                 num_modalities_to_drop=1, subdivide_physiology=True, fill_missing_with=0.0, mask_with=-1.0,
                 checkpoint_dir=DEFAULT_MAIN_DIRECTORY + 'temp_saved_models/', model_name='multimodal_autoencoder',
                 extra_data_filename=None, data_loader=None, classification_data_loader=None, verbose=True,
                 *, rng_mode='numpy', precision='tf32', seed=0, device=None):
        # hyper-parameters (:140-160).  decay / clip_gradients are accepted and stored but inert, exactly as
        # in the reference (global_step is never passed to minimize(), clip_gradients is never read).
        self.layer_sizes = list(layer_sizes)
        self.embedding_size = layer_sizes[-1]
        self.tie_weights = tie_weights
        self.variational = variational
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.dropout_prob = dropout_prob
        self.weight_penalty = weight_penalty
        self.weight_initialization = weight_initialization
        self.classification_layer_sizes = classification_layer_sizes
        self.classification_filename = classification_filename
        self.normalization = normalization
        self.fill_missing_with = fill_missing_with
        self.mask_with = mask_with
        self.clip_gradients = clip_gradients
        self.activation_func = activation_func
        self.loss_func = loss_func
        self.decay, self.decay_steps, self.decay_rate = decay, decay_steps, decay_rate
        self.optimizer = 'adam'
        self.checkpoint_dir = checkpoint_dir
        self.filename = filename
        self.model_name = model_name
        self.record_every_nth = 50
        self.save_every_nth = 100000
        self.subdivide_physiology = subdivide_physiology
        self.intelligent_noise = intelligent_noise
        self.num_modalities_to_drop = num_modalities_to_drop
        self.extra_data_filename = extra_data_filename
        self.verbose = verbose
        self.rng_mode, self.precision, self.seed, self._device = rng_mode, precision, seed, device
        if rng_mode not in ('numpy', 'philox'):
            raise ValueError("rng_mode must be 'numpy' or 'philox'")

        if self.variational:                       # :175-179
            if self.verbose:
                print("Building VAE. Will use 0-1 normalization, cross entropy loss, and will not tie weights.\n")
            self.tie_weights = False
            self.normalization = 'between_0_and_1'
            self.loss_func = 'sigmoid_cross_entropy'
        if self.normalization == 'z_score' and loss_func in ('cross_entropy', 'sigmoid_cross_entropy'):   # :181-184
            print("ERROR! Cannot use cross entropy loss with z-score data. Changing normalization method to 0-1")
            self.normalization = 'between_0_and_1'

        if data_loader is not None:
            self.data_loader = data_loader
        elif filename is not None:
            self.data_loader = data_funcs.DataLoader(filename, supervised=False,
                                                     subdivide_physiology_features=subdivide_physiology,
                                                     normalize_and_fill=False, normalization=self.normalization,
                                                     fill_missing_with=self.fill_missing_with)
        else:
            raise ValueError("Must set either filename or data_loader so that the MMAE has access to data.")
        self.extra_noisy_data_loader = None

        if self.intelligent_noise:
            self.noise_type_percentages = list(DEFAULT_NOISE_P)
            self.noise_types = [list(t) for t in DEFAULT_NOISE_TYPES]

        if self.classification_layer_sizes is not None:
            self.train_acc, self.val_acc = [], []
            self.classification_train_loss, self.classification_val_loss = [], []
            self.classification_learning_rate = .0001          # :215-220
            self.classification_batch_size = 100
            self.classification_dropout_prob = self.dropout_prob
            self.classification_activation_func = self.activation_func
            self.classification_weight_penalty = 0.0
            self.classification_loss_func = 'sigmoid_cross_entropy'
            if classification_data_loader is None:
                self.classification_data_loader = data_funcs.DataLoader(
                    self.classification_filename, supervised=True, subdivide_physiology_features=subdivide_physiology,
                    normalize_and_fill=False, normalization=self.normalization, fill_missing_with=self.fill_missing_with)
            else:
                self.classification_data_loader = classification_data_loader

        # handles the reference exposed as attributes
        for n in ('noisy_X', 'true_X', 'tf_dropout_prob', 'embedding', 'decoded_X', 'reconstruction_loss', 'opt_step',
                  'global_step', 'init'):
            setattr(self, n, _Handle(n))
        self.true_Y = None
        self.engine = None
        self.build_graph()
        self.initialize_session()
        self.train_loss, self.val_loss = [], []

    # ------------------------------------------------------------------ graph == engine
    def _engine_config(self):
        dl = self.data_loader
        head = self.classification_layer_sizes
        kw = {}
        if head is not None:
            kw = dict(cls_layer_sizes=list(head), num_labels=self.classification_data_loader.num_labels,
                      cls_activation=self.classification_activation_func,
                      cls_loss='sigmoid_cross_entropy' if self.classification_loss_func == 'sigmoid_cross_entropy' else 'softmax',
                      cls_weight_penalty=self.classification_weight_penalty,
                      cls_learning_rate=self.classification_learning_rate)
        cfg = EngineConfig(num_feats=dl.num_feats, layer_sizes=list(self.layer_sizes),
                           modality_starts=list(dl.modality_start_indices), modality_names=list(dl.modality_names),
                           tie_weights=bool(self.tie_weights), variational=bool(self.variational),
                           activation=self.activation_func, loss_func=self.loss_func, weight_penalty=self.weight_penalty,
                           learning_rate=self.learning_rate, mask_with=self.mask_with,
                           intelligent_noise=bool(self.intelligent_noise), num_modalities_to_drop=self.num_modalities_to_drop,
                           seed=self.seed, precision=self.precision, max_batch=max(self.batch_size, 200), **kw)
        if self.intelligent_noise:
            cfg.noise_p, cfg.noise_types = self.noise_type_percentages, self.noise_types
        return cfg

    def build_graph(self):
        """Creates the engine (tf.Graph + variables + optimizers in the reference, :344-452)."""
        if self.verbose:
            print('\nBuilding computation graph...')
        if self.engine is not None:
            self.engine.close()
        self.engine = Engine(self._engine_config(), device=self._device)
        if self.classification_layer_sizes is not None:
            for n in ('true_Y', 'logits', 'classification_loss', 'classification_opt_step', 'class_probabilities',
                      'predictions', 'accuracy'):
                setattr(self, n, _Handle(n))
        self._resident = {}
        self._resident_view = {}
        self._step_count = 0

    def initialize_network_weights(self):
        """Initial values with the reference's distributions (:22-56): 'xavier' U(+-sqrt(6/(in+out))), otherwise
        truncated normal (|z| <= 2) with sigma 1/sqrt(in); biases 0.1.  (TensorFlow's RNG stream itself is not
        reproducible; parity runs inject weights through engine.set_variable.)"""
        # every (re)build draws fresh weights, like a new TensorFlow initializer run: the stream is keyed by the seed AND
        # by how many times this model has been initialised (CV folds / grid settings do not restart from equal weights)
        self._init_count = getattr(self, '_init_count', 0) + 1
        rng = np.random.default_rng([int(self.seed), self._init_count - 1])
        for name, shp in self.engine.variables():
            if len(shp) == 1:
                w = np.full(shp, 0.1, np.float32)
            elif self.weight_initialization == 'xavier':
                lim = np.sqrt(6.0 / (shp[0] + shp[1]))
                w = rng.uniform(-lim, lim, shp).astype(np.float32)
            else:
                z = rng.standard_normal(shp)
                bad = np.abs(z) > 2.0
                while bad.any():
                    z[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(z) > 2.0
                w = (z / np.sqrt(float(shp[0]))).astype(np.float32)
            self.engine.set_variable(name, w)

    def initialize_session(self):
        self.initialize_network_weights()
        self.session = _Session(self)
        self.saver = self

    def rebuild_reinitialize(self):
        """Fresh engine + fresh weights; learned weights are discarded, as in the reference (:243-254)."""
        self.build_graph()
        self.initialize_session()
        self.train_loss, self.val_loss = [], []

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    def _head_total_loss(self, sc):
        """classification_loss as the graph defines it: data term + lambda_c * sum l2_loss(W_c) (:441)."""
        reg = 0.0
        if self.classification_weight_penalty:
            for name, shp in self.engine.variables():
                if name.startswith('classification_weights'):
                    reg += 0.5 * float(np.sum(self.engine.get_variable(name).astype(np.float64) ** 2))
        return sc['head_loss'] + self.classification_weight_penalty * reg

    # ------------------------------------------------------------------ noise (:649-702)
    def _descriptor(self, n_rows, missing_modes=()):
        dl = self.data_loader
        override = None
        if len(missing_modes) > 0:
            override = 0
            for m in missing_modes:
                override |= 1 << dl.modality_names.index(m)
        masks = type_masks_from_names(self.noise_types, list(dl.modality_names)) if self.intelligent_noise else None
        return numpy_descriptor(n_rows, dl.num_feats, dl.num_modalities, bool(self.intelligent_noise),
                                getattr(self, 'noise_type_percentages', None), masks, self.num_modalities_to_drop,
                                override_mask=override)

    def mask_modality(self, X, row, mod_i):
        s, e = self.data_loader.modality_start_indices[mod_i], self.data_loader.modality_start_indices[mod_i + 1]
        X[row, s:e] = self.mask_with
        return X

    def add_noise_to_batch(self, X, missing_modes=[]):
        """Noisy copy of X; host RNG in the reference's order (rng_mode='numpy') or the device generator."""
        X = np.asarray(X)
        if self.rng_mode == 'numpy' or len(missing_modes) > 0:
            zb, mb = self._descriptor(len(X), missing_modes)
        else:
            self.engine.set_rng_step(self._next_rng_step())
            self.engine.gen_noise(len(X))
            zb, mb = self.engine.get_noise(len(X))
        return apply_descriptor_host(X, zb, mb, self.data_loader.modality_start_indices, self.mask_with)

    def _next_rng_step(self):
        self._step_count += 1
        return self._step_count

    def _prepare_noise(self, n_rows):
        """Loads a fresh descriptor for n_rows rows into the engine (numpy: host draw + upload; philox: device)."""
        if self.rng_mode == 'numpy':
            zb, mb = self._descriptor(n_rows)
            self.engine.set_noise(zb, mb)
            self.engine.set_rng_step(self._next_rng_step())
        else:
            self.engine.set_rng_step(self._next_rng_step())
            self.engine.gen_noise(n_rows)

    # ------------------------------------------------------------------ training loops (:549-647)
    def set_record_save(self, record_every_nth, save_every_nth):
        if record_every_nth is not None:
            self.record_every_nth = record_every_nth
        if save_every_nth is not None:
            self.save_every_nth = save_every_nth

    def train(self, num_steps=30000, record_every_nth=None, save_every_nth=None):
        """Unsupervised training; RNG call order per step as in the reference: train indices, train noise rows,
        then (record steps only) validation indices and validation noise."""
        self.set_record_save(record_every_nth, save_every_nth)
        eng, dl = self.engine, self.data_loader
        if self.rng_mode == 'philox':
            return self._train_resident(num_steps, classification=False)
        for step in range(num_steps):
            X = np.ascontiguousarray(dl.get_unsupervised_train_batch(self.batch_size), np.float32)
            self._prepare_noise(len(X))
            if step % self.record_every_nth == 0:
                # train-loss run reuses the noisy training feed *with the training keep_prob* (:575, :726)
                Xd = eng._dev(X)
                eng.forward(Xd, target=Xd, noise=True, keep=self.dropout_prob, loss=True)
                train_loss = eng.scalars()['recon_loss']
                val_X = np.ascontiguousarray(dl.get_unsupervised_val_batch(200), np.float32)
                zb, mb = eng.get_noise(len(X))                  # keep the training descriptor for the optimizer step
                val_loss = self._loss_on(val_X, noise=True)
                eng.set_noise(zb, mb)
                # the evaluation forwards advanced the engine's Philox step past the host counter: take a fresh one so
                # that this optimizer step and the next do not share dropout masks / epsilon
                eng.set_rng_step(self._next_rng_step())
                if 'entropy' in self.loss_func:                 # :733-735
                    train_loss, val_loss = train_loss / len(X), val_loss / len(val_X)
                self.train_loss.append(train_loss)
                self.val_loss.append(val_loss)
                if self.verbose:
                    print("Training iteration", step)
                    print("\t Training loss", train_loss)
                    print("\t Validation loss", val_loss)
            if step > 0 and step % self.save_every_nth == 0:
                self.save_model()
            eng.train_step_host(X, use_noise=True, keep=self.dropout_prob)      # staged input: fixed device buffers -> graph replay

    def _train_resident(self, num_steps, classification):
        """rng_mode='philox': the training matrix lives on the device, every step samples its rows, draws its
        block-mask noise and runs fwd + bwd + Adam without touching the host; the host only joins at record steps."""
        eng = self.engine
        dl = self.classification_data_loader if classification else self.data_loader
        slot = 1 if classification else 0
        if getattr(dl, 'cross_validation', False) and getattr(dl, 'train_index', None) is not None:
            # cross validation: the base matrix goes to the device once, a fold is a row list (set_dataset_view)
            bX, bY = dl.cross_val_base()
            key = ('cv', id(bX), len(bX))
            if self._resident.get(slot) != key:
                eng.set_dataset(slot, bX, bY if classification else None)
                self._resident[slot] = key
                self._resident_view[slot] = None
            vkey = (dl.fold, len(dl.train_index))
            if self._resident_view.get(slot) != vkey:
                eng.set_dataset_view(slot, dl.train_index)
                self._resident_view[slot] = vkey
        else:
            key = (id(dl.train_X), getattr(dl, 'fold', None), len(dl.train_X))
            if self._resident.get(slot) != key:
                eng.set_dataset(slot, dl.train_X, dl.train_Y if classification else None)
                self._resident[slot] = key
                self._resident_view[slot] = None
        B = self.classification_batch_size if classification else self.batch_size
        keep = self.classification_dropout_prob if classification else self.dropout_prob
        for step in range(num_steps):
            if step % self.record_every_nth == 0:
                if classification:
                    tl, ta, vl, va = self.evaluate_classification_performance()
                    self.train_acc.append(ta); self.val_acc.append(va)
                    self.classification_train_loss.append(tl); self.classification_val_loss.append(vl)
                else:
                    # train loss: a batch sampled, noised and evaluated on the device (the training matrix is resident)
                    eng.set_rng_step(self._next_rng_step())
                    eng.eval_resident(slot, B, gen_noise=True, keep=keep)
                    tl = eng.scalars()['recon_loss']
                    val_X = dl.get_unsupervised_val_batch(200)
                    vl = self._loss_on(val_X, noise=True)
                    if 'entropy' in self.loss_func:
                        tl, vl = tl / B, vl / len(val_X)
                    self.train_loss.append(tl); self.val_loss.append(vl)
                if self.verbose:
                    print("Training iteration", step, "\t train", tl, "\t validation", vl)
            if step > 0 and step % self.save_every_nth == 0:
                self.save_model()
            eng.set_rng_step(self._next_rng_step())
            eng.train_step_resident(slot, B, idx=None, gen_noise=True, keep=keep, classification=classification)

    def _loss_on(self, X, noise, keep=1.0, target=None):
        """reconstruction_loss (graph value: a batch sum for the entropy losses) of X against target (default X)."""
        X = np.ascontiguousarray(X, np.float32)
        if noise:
            self._prepare_noise(len(X))
        Xd = self.engine._dev(X)
        Td = Xd if target is None else self.engine._dev(np.ascontiguousarray(target, np.float32))
        self.engine.forward(Xd, target=Td, noise=noise, keep=keep, loss=True)
        return self.engine.scalars()['recon_loss']

    def train_classification(self, num_steps=30000, record_every_nth=None, save_every_nth=None):
        self.set_record_save(record_every_nth, save_every_nth)
        eng, dl = self.engine, self.classification_data_loader
        if self.rng_mode == 'philox':
            return self._train_resident(num_steps, classification=True)
        for step in range(num_steps):
            X, Y = dl.get_supervised_train_batch(self.classification_batch_size)
            X = np.ascontiguousarray(X, np.float32)
            Y = np.ascontiguousarray(Y, np.float32)
            self._prepare_noise(len(X))
            if step % self.record_every_nth == 0:
                res = self.evaluate_classification_performance((X, Y, True, self.classification_dropout_prob))
                eng.set_rng_step(self._next_rng_step())         # see train(): evaluation forwards advanced the engine's step
                train_loss, train_acc, val_loss, val_acc = res
                self.train_acc.append(train_acc)
                self.val_acc.append(val_acc)
                self.classification_train_loss.append(train_loss)
                self.classification_val_loss.append(val_loss)
                if self.verbose:
                    print("Training iteration", step)
                    print("\t Training loss", train_loss, "\t Validation loss", val_loss)
                    print("\t Training accuracy", train_acc, "\t Validation accuracy", val_acc)
            if step > 0 and step % self.save_every_nth == 0:
                self.save_model()
            eng.cls_train_step_host(X, Y, use_noise=True, keep=self.classification_dropout_prob)

    def evaluate_performance(self, train_feed_dict=None):
        """(train loss, validation loss) on one batch each; validation batch of 200 with noise (:704-737)."""
        if train_feed_dict is None:
            X = self.data_loader.get_unsupervised_train_batch(self.batch_size)
            train_loss = self._loss_on(X, noise=False)
            n_train = len(X)
        else:
            X, T, keep = train_feed_dict[self.noisy_X], train_feed_dict[self.true_X], train_feed_dict[self.tf_dropout_prob]
            train_loss = self._loss_on(X, noise=False, keep=keep, target=T)
            n_train = len(T)
        val_X = self.data_loader.get_unsupervised_val_batch(200)
        val_loss = self._loss_on(val_X, noise=True)
        if 'entropy' in self.loss_func:
            train_loss, val_loss = train_loss / n_train, val_loss / len(val_X)
        return train_loss, val_loss

    def evaluate_classification_performance(self, train_feed_dict=None):
        """(train loss, train acc, val loss, val acc); validation X is *not* noised (:754-757)."""
        eng = self.engine
        if train_feed_dict is None:
            X, Y = self.classification_data_loader.get_supervised_train_batch(self.classification_batch_size)
            use_noise, keep = False, self.dropout_prob
        elif isinstance(train_feed_dict, dict):
            X, Y = train_feed_dict[self.noisy_X], train_feed_dict[self.true_Y]
            use_noise, keep = False, train_feed_dict[self.tf_dropout_prob]
        else:
            X, Y, use_noise, keep = train_feed_dict
        eng.forward(np.ascontiguousarray(X, np.float32), labels=np.ascontiguousarray(Y, np.float32), noise=use_noise,
                    keep=keep, head_loss=True)
        sc = eng.scalars()
        train_loss, train_acc = self._head_total_loss(sc), sc['head_acc']
        val_X, val_Y = self.classification_data_loader.get_supervised_val_batch(200)
        eng.forward(np.ascontiguousarray(val_X, np.float32), labels=np.ascontiguousarray(val_Y, np.float32), head_loss=True)
        sc = eng.scalars()
        return train_loss, train_acc, self._head_total_loss(sc), sc['head_acc']

    # ------------------------------------------------------------------ checkpoints (:766-896)
    def save_model(self, file_name=None, directory=None):
        """Parameters + both Adam states + loss curves + hyper-parameters in one .npz, variable names as in the
        reference's checkpoint (weights0, decode_weights0, encode_biases0, ...).  Returns the path."""
        if self.verbose:
            print("Saving model...")
        file_name = file_name or self.model_name
        if directory is None:
            directory = self.checkpoint_dir
        else:
            directory = os.path.join(directory + file_name, '')
        os.makedirs(directory, exist_ok=True)
        training_epochs = len(self.train_loss) * self.record_every_nth
        path = os.path.join(directory, '%s-%d.npz' % (file_name, training_epochs))
        blob = {}
        for name, _ in self.engine.variables():
            blob['var/' + name] = self.engine.get_variable(name)
            for opt in (0, 1):
                try:
                    m, v, t = self.engine.get_opt_state(opt, name)
                except (ValueError, RuntimeError):
                    continue
                blob['adam%d_m/%s' % (opt, name)] = m
                blob['adam%d_v/%s' % (opt, name)] = v
                blob['adam%d_t' % opt] = np.int64(t)
        np.savez(path, train_loss=self.train_loss, val_loss=self.val_loss, layer_sizes=self.layer_sizes,
                 variational=self.variational, dropout_prob=self.dropout_prob, weight_penalty=self.weight_penalty,
                 activation_func=self.activation_func, loss_func=self.loss_func,
                 weight_initialization=self.weight_initialization, tie_weights=self.tie_weights,
                 rng_step_count=np.int64(self._step_count), **blob)
        return path

    def load_saved_model(self, directory=None, checkpoint_name=None, npz_file_name=None):
        """Restores hyper-parameters, weights and optimizer state from a save_model() file."""
        directory = directory or self.checkpoint_dir          # the reference read an undefined self.output_dir (:818)
        name = checkpoint_name or npz_file_name
        if name is not None and not name.endswith('.npz'):
            # reference-style names: 'model.ckpt-N' (the TF checkpoint) or 'model-N' (its .npz twin) -> 'model-N.npz'
            name = name.replace('.ckpt-', '-') + '.npz'
        if name is None:
            cands = sorted((f for f in os.listdir(directory) if f.endswith('.npz')),
                           key=lambda f: os.path.getmtime(os.path.join(directory, f)))
            if not cands:
                print("Error! Cannot locate checkpoint in the directory")
                return
            name = cands[-1]
        z = np.load(os.path.join(directory, name), allow_pickle=False)
        self.train_loss, self.val_loss = list(z['train_loss']), list(z['val_loss'])
        rebuild = False
        for key in ('layer_sizes', 'variational', 'dropout_prob', 'weight_penalty', 'activation_func', 'loss_func',
                    'weight_initialization', 'tie_weights'):
            if self._print_if_saved_setting_differs(getattr(self, key), key, z):
                val = z[key]
                setattr(self, key, val.tolist() if val.ndim else val.item())
                rebuild = True
        if rebuild:
            self.embedding_size = self.layer_sizes[-1]
            tl, vl = self.train_loss, self.val_loss
            self.rebuild_reinitialize()
            self.train_loss, self.val_loss = tl, vl
        if 'rng_step_count' in z.files:                       # a resumed philox run continues its noise / dropout stream
            self._step_count = int(z['rng_step_count'])
        for name_, _ in self.engine.variables():
            self.engine.set_variable(name_, z['var/' + name_])
            for opt in (0, 1):
                k = 'adam%d_m/%s' % (opt, name_)
                if k in z:
                    self.engine.set_opt_state(opt, name_, z[k], z['adam%d_v/%s' % (opt, name_)], int(z['adam%d_t' % opt]))

    def _print_if_saved_setting_differs(self, class_var, setting_name, npz_file):
        if setting_name not in npz_file.keys():
            print("ERROR! The setting", setting_name, "is not in the saved model file. Using", class_var)
            return False
        saved = npz_file[setting_name]
        same = (list(np.atleast_1d(saved)) == list(np.atleast_1d(class_var)))
        if not same:
            print("WARNING! Saved setting for", setting_name, "differs:", class_var, "->", saved)
        return not same

    def set_classification_params(self, weight_penalty=None, learning_rate=None, dropout_prob=None, activation_func=None,
                                  batch_size=None, loss_func=None, suppress_warning=False):
        """Sets head hyper-parameters, then rebuilds and re-initialises everything (:898-930)."""
        pick = lambda new, old: new if new is not None else old
        self.classification_learning_rate = pick(learning_rate, self.classification_learning_rate)
        self.classification_dropout_prob = pick(dropout_prob, self.classification_dropout_prob)
        self.classification_activation_func = pick(activation_func, self.classification_activation_func)
        self.classification_weight_penalty = pick(weight_penalty, self.classification_weight_penalty)
        self.classification_batch_size = pick(batch_size, self.classification_batch_size)
        self.classification_loss_func = pick(loss_func, self.classification_loss_func)
        if not suppress_warning:
            print("The model will now be rebuilt and re-initialised. Unsaved changes will be lost.")
        self.rebuild_reinitialize()

    # ------------------------------------------------------------------ inference (:932-1216)
    def predict(self, X):
        """(reconstruction, loss): X is both the input and the target, as in the reference (:941-942)."""
        r = self.engine.forward_host(np.ascontiguousarray(X, np.float32), recon=True, loss=True)
        loss = self.engine.scalars()['recon_loss']
        if 'entropy' in self.loss_func:
            loss = loss / len(X)
        return r['recon'], loss

    def test_on_validation(self):
        loss = self.get_performance_on_data(self.data_loader.val_X)
        print("Final loss on validation data is:", loss)
        return loss

    def test_on_test(self):
        print("WARNING! Only test on the test set when you have finished choosing all of your hyperparameters!")
        loss = self.get_performance_on_data(self.data_loader.test_X)
        print("Final loss on test data is:", loss)
        return loss

    def get_performance_on_data(self, X):
        loss = self._loss_on(X, noise=False)
        return loss / len(X) if 'entropy' in self.loss_func else loss

    def get_performance_on_data_with_noise(self, X):
        loss = self._loss_on(X, noise=True)
        return loss / len(X) if 'entropy' in self.loss_func else loss

    def get_classification_predictions(self, X):
        return self.engine.forward_host(np.ascontiguousarray(X, np.float32), head=True)['preds']

    def get_classification_predictions_from_df(self):
        dl = self.classification_data_loader
        df = copy.deepcopy(dl.df)
        preds = self.get_classification_predictions(df[dl.wanted_feats].to_numpy())
        assert len(df) == len(preds)
        for i, label in enumerate(dl.wanted_labels):
            df['predictions_' + label] = preds[:, i] if preds.ndim > 1 else preds
        return df

    def get_embedding(self, X, add_noise=False):
        X = np.ascontiguousarray(X, np.float32)
        if add_noise:
            self._prepare_noise(len(X))
            return self.engine.forward(X, noise=True, embedding=True)['embedding'].cpu().numpy()
        return self.engine.forward_host(X, embedding=True)['embedding']

    def get_performance_on_extra_noisy_data(self):
        if self.extra_noisy_data_loader is None:
            if self.extra_data_filename is None:
                print("Error! Was not provided with location of extra data. Cannot perform this command")
                return
            self.extra_noisy_data_loader = data_funcs.DataLoader(
                self.extra_data_filename, normalize_and_fill=False, subdivide_physiology_features=self.subdivide_physiology,
                normalization=self.normalization, fill_missing_with=self.fill_missing_with, fill_gaps_with=self.mask_with)
        return self.get_performance_on_data(self.extra_noisy_data_loader.train_X)

    def convert_file_to_embeddings(self, filename, path, file_descriptor=""):
        """CSV -> CSV with ae_embedding_dim* columns (the reference wrote X[:, c] by mistake, :1164)."""
        import pandas as pd
        df = pd.read_csv(path + filename, index_col=0)
        feats = data_funcs.get_wanted_feats_from_df(df)
        emb = self.get_embedding(data_funcs.get_matrix_for_dataset(df, feats, dataset=None))
        out = df[[c for c in df.columns.values if c not in feats]].copy()
        for c in range(emb.shape[1]):
            out['ae_embedding_dim' + str(c)] = emb[:, c]
        out.to_csv(path + 'embedding-' + file_descriptor + filename)
        return out

    def fill_missing_data_in_file(self, filename, path, file_descriptor=""):
        """Reconstructs every row, replaces only the modality blocks that were missing (:1167-1187)."""
        import pandas as pd
        df = pd.read_csv(path + filename, index_col=0)
        dl = self.data_loader
        X = df[dl.wanted_feats].to_numpy(dtype=np.float64)
        filled = self.fill_missing(X)              # float32, device-side select
        # only the blocks that were missing are written back: every observed cell keeps its float64 value bit for bit
        # (fill_df_with_reconstruction touches nothing else, data_funcs.py:328-348); the rule runs on the float64 matrix
        cols = np.repeat(dl.missing_modality_mask(X), np.diff(dl.modality_start_indices), axis=1)
        df.loc[:, dl.wanted_feats] = np.where(cols, filled.astype(np.float64), X)
        df.to_csv(path + 'MMAE_filled-' + file_descriptor + filename)
        return df

    def fill_missing(self, X):
        """Device-side fill-in: reconstruction on missing blocks (sum == -width), original elsewhere."""
        return self.engine.forward_host(np.ascontiguousarray(X, np.float32), recon=True, filled=True)['filled']

    def get_reconstruction_loss_per_modality(self, X):
        """Per modality: mask it with literal -1.0 for every row, reconstruct, RMSE on that block (:1189-1216)."""
        dl = self.data_loader
        if len(X) == 0:
            return [np.nan] * len(dl.modality_names)
        # One batched device pass (mmae_modality_rmse): the M masked copies of the rows form one batch, one forward
        # reconstructs them all, the squared errors of each copy's own block are reduced on the device.  The mask value
        # is the literal -1.0, as in the reference (:1203), whatever mask_with is.
        rms = self.engine.modality_rmse(X)
        if self.verbose:
            for name, v in zip(dl.modality_names, rms):
                print("RMS for modality", name, "is", v)
        return rms

    # ------------------------------------------------------------------ plots (optional dependency)
    def _plt(self):
        try:
            import matplotlib.pyplot as plt
            return plt
        except Exception:
            print("matplotlib is not available; skipping the plot")
            return None

    def plot_training_progress(self):
        plt = self._plt()
        if plt is None:
            return
        x = [self.record_every_nth * i for i in range(len(self.train_loss))]
        plt.figure(); plt.plot(x, self.train_loss); plt.plot(x, self.val_loss)
        plt.legend(['Train', 'Validation'], loc='best'); plt.xlabel('Training epoch'); plt.ylabel('Loss'); plt.show()

    def plot_classification_training_progress(self):
        plt = self._plt()
        if plt is None:
            return
        x = [self.record_every_nth * i for i in range(len(self.train_acc))]
        for a, b, lab in ((self.train_acc, self.val_acc, 'Accuracy'),
                          (self.classification_train_loss, self.classification_val_loss, 'Classification loss')):
            plt.figure(); plt.plot(x, a); plt.plot(x, b)
            plt.legend(['Train', 'Validation'], loc='best'); plt.xlabel('Training epoch'); plt.ylabel(lab); plt.show()

    def view_reconstruction(self, dataset, with_noise=True):
        plt = self._plt()
        i = np.random.randint(0, len(dataset))
        X = np.reshape(dataset[i, :], [1, -1])
        noisy = self.add_noise_to_batch(X) if with_noise else X
        recon = self.engine.forward_host(np.ascontiguousarray(noisy, np.float32), recon=True)['recon']
        if plt is not None:
            plt.figure()
            if with_noise:
                plt.plot(np.reshape(noisy, -1))
            plt.plot(np.reshape(X, -1)); plt.plot(np.reshape(recon, -1), c='r'); plt.show()
        return recon


def get_rmse(x, y):
    """Root mean squared error between two arrays (:1218-1220; sklearn's mean_squared_error == plain mean)."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    return float(np.sqrt(np.mean((x - y) ** 2)))
