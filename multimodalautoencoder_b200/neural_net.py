"""NeuralNetwork -- the reference's plain MLP classifier (comparison_algorithms/neural_net.py:27-381) on the B200 engine.

Same constructor keywords, methods and attributes as the reference class; the graph it built in TensorFlow
(:136-198: relu on every hidden layer, dropout, linear logits, mean sigmoid cross-entropy + weight_penalty * sum
l2_loss(W), tf.clip_by_global_norm(gradients, 5), Adam with a live global_step) runs in libmmae_b200.so as the engine's
classifier-only mode: the encoder stack with every layer activated, one logits layer, the head optimizer with the L2
term on every weight matrix and the global-norm clip folded into the Adam pass (one reduction + one update kernel).

Variable names follow the reference checkpoint: weights{i} / biases{i}, i = 0 .. len(layer_sizes).
"""
from __future__ import annotations

import os

import numpy as np

from . import data_funcs
from .engine import Engine, EngineConfig

DEFAULT_MAIN_DIRECTORY = '/Your/path/here/'


class NeuralNetwork:
    def __init__(self, filename=None, layer_sizes=[128, 64], batch_size=20, learning_rate=.001, dropout_prob=1.0,
                 weight_penalty=0.0, model_name='NN', clip_gradients=True, data_loader=None,
                 checkpoint_dir=DEFAULT_MAIN_DIRECTORY + 'temp_saved_models/', verbose=True, *, precision='tf32', seed=0,
                 device=None):
        self.layer_sizes = list(layer_sizes)
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.dropout_prob = dropout_prob
        self.weight_penalty = weight_penalty
        self.clip_gradients = clip_gradients
        self.activation_func = 'relu'                     # :70
        self.optimizer = 'adam'
        self.checkpoint_dir = checkpoint_dir
        self.filename = filename
        self.model_name = model_name
        self.output_every_nth = 100
        self.verbose = verbose
        self.precision, self.seed, self._device = precision, seed, device
        self.data_loader = data_loader if data_loader is not None else data_funcs.DataLoader(filename)
        self.input_size = self.data_loader.get_feature_size()
        self.output_size = self.data_loader.num_labels
        if self.verbose:
            print("Input dimensions (number of features):", self.input_size)
            print("Number of classes/outputs:", self.output_size)
        self.engine = None
        self.build_graph()
        self.train_acc, self.val_acc = [], []

    # engine variable <-> reference checkpoint name
    def _names(self):
        L = len(self.layer_sizes)
        m = {}
        for i in range(L):
            m['weights%d' % i] = 'weights%d' % i
            m['biases%d' % i] = 'encode_biases%d' % i
        m['weights%d' % L] = 'classification_weights0'
        m['biases%d' % L] = 'classification_biases0'
        return m

    def build_graph(self):
        dl = self.data_loader
        if self.engine is not None:
            self.engine.close()
        starts = list(getattr(dl, 'modality_start_indices', None) or [0, self.input_size])
        names = list(getattr(dl, 'modality_names', None) or ['all'])
        cfg = EngineConfig(num_feats=self.input_size, layer_sizes=list(self.layer_sizes), modality_starts=starts,
                           modality_names=names, tie_weights=False, variational=False, activation=self.activation_func,
                           cls_layer_sizes=[], num_labels=self.output_size, cls_activation=self.activation_func,
                           cls_loss='sigmoid_cross_entropy', cls_weight_penalty=self.weight_penalty,
                           cls_learning_rate=self.learning_rate, intelligent_noise=False, seed=self.seed,
                           precision=self.precision, max_batch=max(self.batch_size, 256), classifier_only=True,
                           clip_norm=5.0 if self.clip_gradients else 0.0)
        self.engine = Engine(cfg, device=self._device)
        self.global_step = 0                               # live: apply_gradients(..., self.global_step) (:192-193)
        self.initialize_network_weights()

    def initialize_network_weights(self):
        """truncated normal (|z| <= 2) with sigma = 1/sqrt(in), biases 0.1 (:383-406)."""
        self._init_count = getattr(self, '_init_count', 0) + 1
        rng = np.random.default_rng([int(self.seed), self._init_count - 1])
        sizes = []
        for ref_name, eng_name in self._names().items():
            shp = self.engine.shape_of(eng_name)
            if len(shp) == 1:
                w = np.full(shp, 0.1, np.float32)
            else:
                z = rng.standard_normal(shp)
                bad = np.abs(z) > 2.0
                while bad.any():
                    z[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(z) > 2.0
                w = (z / np.sqrt(float(shp[0]))).astype(np.float32)
                sizes.append(('%dx%d' % shp, str(shp[1])))
            self.engine.set_variable(eng_name, w)
        if self.verbose:
            print("Okay, making a neural net with the following structure:")
            print(sizes)

    def get_variables(self):
        return {r: self.engine.get_variable(e) for r, e in self._names().items()}

    def set_variables(self, values):
        names = self._names()
        for r, v in values.items():
            self.engine.set_variable(names[r], v)

    def _loss_total(self, data_loss):
        reg = 0.0
        if self.weight_penalty:
            for r, e in self._names().items():
                if r.startswith('weights'):
                    reg += 0.5 * float(np.sum(self.engine.get_variable(e).astype(np.float64) ** 2))
        return data_loss + self.weight_penalty * reg

    def _evaluate(self, X, Y, keep):
        self.engine.forward(np.ascontiguousarray(X, np.float32), labels=np.ascontiguousarray(Y, np.float32), keep=keep,
                            head_loss=True)
        sc = self.engine.scalars()
        return sc['head_acc'], self._loss_total(sc['head_loss'])

    def train(self, num_steps=30000, output_every_nth=None):
        """The reference loop (:200-244): optimizer step first, then (every output_every_nth steps) accuracy / loss on the
        training feed (with the training dropout) and on the whole validation set, and a checkpoint."""
        if output_every_nth is not None:
            self.output_every_nth = output_every_nth
        eng, dl = self.engine, self.data_loader
        for step in range(num_steps):
            X, Y = dl.get_supervised_train_batch(self.batch_size)
            X = np.ascontiguousarray(X, np.float32)
            Y = np.ascontiguousarray(Y, np.float32)
            eng.set_rng_step(self.global_step + 1)
            eng.cls_train_step_host(X, Y, keep=self.dropout_prob)
            self.global_step += 1
            if step % self.output_every_nth == 0:
                val_X, val_Y = dl.get_val_data()
                train_score, _ = self._evaluate(X, Y, self.dropout_prob)
                val_score, loss = self._evaluate(val_X, val_Y, 1.0)
                if self.verbose:
                    print("Training iteration", step)
                    print("\t Training acc", train_score)
                    print("\t Validation acc", val_score)
                    print("\t Loss", loss)
                self.train_acc.append(train_score)
                self.val_acc.append(val_score)
                if self.checkpoint_dir and os.path.isdir(self.checkpoint_dir):
                    self.save_model()

    def predict(self, X, get_probabilities=False):
        r = self.engine.forward_host(np.ascontiguousarray(X, np.float32), head=True)
        return (r['preds'], r['probs']) if get_probabilities else r['preds']

    def get_performance_on_data(self, X, Y):
        return self._evaluate(X, Y, 1.0)[0]

    def test_on_validation(self):
        score = self.get_performance_on_data(self.data_loader.val_X, self.data_loader.val_Y)
        print("Final accuracy on validation data is:", score)
        return score

    def test_on_test(self):
        score = self.get_performance_on_data(self.data_loader.test_X, self.data_loader.test_Y)
        print("Final accuray on test data is:", score)
        return score

    def plot_training_progress(self):
        try:
            import matplotlib.pyplot as plt
        except Exception:
            print("matplotlib is not available; skipping the plot")
            return
        x = [self.output_every_nth * i for i in np.arange(len(self.train_acc))]
        plt.figure(); plt.plot(x, self.train_acc); plt.plot(x, self.val_acc)
        plt.legend(['Train', 'Validation'], loc='best'); plt.xlabel('Training epoch'); plt.ylabel('Accuracy'); plt.show()

    def save_model(self, file_name=None, directory=None):
        """Variables (reference names), the head optimizer's Adam state, global_step and the accuracy curves in one .npz."""
        if self.verbose:
            print("Saving model...")
        file_name = file_name or self.model_name
        if directory is None:
            directory = self.checkpoint_dir
        else:
            directory = os.path.join(directory + file_name, '')
        os.makedirs(directory, exist_ok=True)
        training_epochs = len(self.train_acc) * self.output_every_nth
        path = os.path.join(directory, '%s-%d.npz' % (file_name, training_epochs))
        blob = {}
        for r, e in self._names().items():
            blob['var/' + r] = self.engine.get_variable(e)
            m, v, t = self.engine.get_opt_state(1, e)
            blob['adam_m/' + r], blob['adam_v/' + r], blob['adam_t'] = m, v, np.int64(t)
        np.savez(path, train_acc=self.train_acc, val_acc=self.val_acc, global_step=np.int64(self.global_step), **blob)
        return path

    def load_saved_model(self, directory=None, checkpoint_name=None, npz_file_name=None):
        print("-----Loading saved model-----")
        directory = directory or self.checkpoint_dir
        name = checkpoint_name or npz_file_name
        if name is None:
            cands = sorted((f for f in os.listdir(directory) if f.endswith('.npz')),
                           key=lambda f: os.path.getmtime(os.path.join(directory, f)))
            if not cands:
                print("Error! Cannot locate checkpoint in the directory")
                return
            name = cands[-1]
        if not name.endswith('.npz'):
            name = name.replace('.ckpt-', '-') + '.npz'
        z = np.load(os.path.join(directory, name))
        self.train_acc, self.val_acc = list(z['train_acc']), list(z['val_acc'])
        self.build_graph()
        self.global_step = int(z['global_step'])
        for r, e in self._names().items():
            self.engine.set_variable(e, z['var/' + r])
            self.engine.set_opt_state(1, e, z['adam_m/' + r], z['adam_v/' + r], int(z['adam_t']))

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None


DEFAULT_NUM_CROSS_FOLDS = 5
LABELS_TO_PREDICT = ['happiness', 'health', 'calmness']


def _wrapper_base():
    from .generic_wrapper import ClassificationWrapper
    return ClassificationWrapper


class NNWrapper(_wrapper_base()):
    """Grid search over the MLP's hyper-parameters (comparison_algorithms/neural_net.py:407-631): architecture x
    dropout_prob x weight_penalty x learning_rate x batch_size, every setting trained on each cross-validation fold
    with a fresh `NeuralNetwork`, metrics per label written into the results row.

    Differences from the reference, both deliberate: the engine of the previous setting is destroyed before the next
    one is built (the reference leaks a TF graph per setting), and `test_on_test` evaluates on the TEST rows -- the
    reference passes `predict_on='test'`, which its own `train_and_predict` (:466-469, compares with 'Test') sends to
    the validation rows and then scores against `test_Y`.
    """

    def __init__(self, filename, layer_sizes=[[300, 200, 100], [200, 100], [128, 64], [200, 100, 50]],
                 dropout_probs=[0.5, 1.0], weight_penalties=[0.0, .01, .001, .0001], learning_rates=[.001],
                 batch_sizes=[100], num_steps=5000, output_every_nth=5001, cont=False, classifier_name='NN',
                 num_cross_folds=DEFAULT_NUM_CROSS_FOLDS, dropbox_path=DEFAULT_MAIN_DIRECTORY,
                 datasets_path='Data/Cleaned/', results_path=None, check_test=True, normalize_and_fill=False,
                 normalization='between_0_and_1', optimize_for='val_acc', min_or_max='max', save_results_every_nth=1,
                 check_noisy_data=True, cross_validation=True, shard=None, *, data_loader=None, precision='tf32', seed=0,
                 device=None):
        self._given_loader = data_loader
        self.layer_sizes = layer_sizes
        self.dropout_probs = dropout_probs
        self.weight_penalties = weight_penalties
        self.batch_sizes = batch_sizes
        self.learning_rates = learning_rates
        self.num_steps = num_steps
        self.output_every_nth = output_every_nth
        self.precision, self.seed, self._device = precision, seed, device
        self.model = None
        _wrapper_base().__init__(
            self, filename=filename, wanted_label=None, cont=cont, classifier_name=classifier_name,
            num_cross_folds=num_cross_folds, dropbox_path=dropbox_path, datasets_path=datasets_path,
            results_path=results_path, check_test=check_test, normalize_and_fill=normalize_and_fill,
            normalization=normalization, optimize_for=optimize_for, min_or_max=min_or_max,
            save_results_every_nth=save_results_every_nth, check_noisy_data=check_noisy_data,
            cross_validation=cross_validation, shard=shard)

    def load_data(self):
        if self._given_loader is not None:           # a loader built by the caller (tests, in-memory frames)
            self.data_loader = self._given_loader
        else:
            _wrapper_base().load_data(self)

    def define_params(self):
        self.params = {'architecture': self.layer_sizes, 'dropout_prob': self.dropout_probs,
                       'weight_penalty': self.weight_penalties, 'learning_rate': self.learning_rates,
                       'batch_size': self.batch_sizes}

    def predict_on_data(self, X):
        return self.model.predict(X)

    def make_model(self, param_dict):
        """One `NeuralNetwork` per (setting, fold); the previous engine goes first."""
        if self.model is not None:
            self.model.close()
        self.model = NeuralNetwork(layer_sizes=param_dict['architecture'], batch_size=int(param_dict['batch_size']),
                                   learning_rate=param_dict['learning_rate'], dropout_prob=param_dict['dropout_prob'],
                                   weight_penalty=param_dict['weight_penalty'], data_loader=self.data_loader,
                                   checkpoint_dir=None, verbose=False, precision=self.precision, seed=self.seed,
                                   device=self._device)
        return self.model

    def train_and_predict(self, param_dict, predict_on='Val'):
        predict_X = self.data_loader.test_X if predict_on == 'Test' else self.data_loader.val_X
        self.make_model(param_dict)
        self.model.train(num_steps=self.num_steps, output_every_nth=self.output_every_nth)
        if predict_on == 'df':
            return self.get_classification_predictions_from_df()
        return self.predict_on_data(predict_X)

    def test_on_test(self, param_dict):
        return self.train_and_predict(param_dict, predict_on='Test')

    @staticmethod
    def _metrics(preds, true_y):
        """[num_labels, 5] rows of (acc, auc, f1, precision, recall)."""
        from .generic_wrapper import compute_all_classification_metrics
        preds, true_y = np.asarray(preds), np.asarray(true_y)
        if preds.ndim == 1:
            preds, true_y = preds[:, None], true_y.reshape(len(true_y), -1)
        return np.array([compute_all_classification_metrics(preds[:, l], true_y[:, l])
                         for l in range(preds.shape[1])], dtype=np.float64)

    def get_cross_validation_results(self, param_dict):
        """Per-fold metrics for every label (:503-580); result columns keep the reference's names."""
        dl = self.data_loader
        n_labels = len(dl.wanted_labels)
        all_m = np.full((self.num_cross_folds, n_labels, 5), np.nan)
        noisy_m = np.full((self.num_cross_folds, n_labels, 5), np.nan)
        clean_m = np.full((self.num_cross_folds, n_labels, 5), np.nan)
        for f in range(self.num_cross_folds):
            dl.set_to_cross_validation_fold(f)
            preds = self.train_and_predict(param_dict)
            all_m[f] = self._metrics(preds, dl.val_Y)
            if self.check_noisy_data:
                noisy_m[f] = self._metrics(self.predict_on_data(dl.noisy_val_X), dl.noisy_val_Y)
                clean_m[f] = self._metrics(self.predict_on_data(dl.clean_val_X), dl.clean_val_Y)
        for j, k in enumerate(('acc', 'auc', 'f1', 'precision', 'recall')):
            param_dict['val_' + k] = np.nanmean(all_m[:, :, j])
        print("Finished training all folds, average acc was", param_dict['val_acc'])
        labels = LABELS_TO_PREDICT[:n_labels]
        for i, label in enumerate(labels):
            param_dict['val_acc_' + label] = np.nanmean(all_m[:, i, 0])
            param_dict['val_auc_' + label] = np.nanmean(all_m[:, i, 1])
        if self.check_noisy_data:
            param_dict['noisy_val_acc'] = np.nanmean(noisy_m[:, :, 0])
            param_dict['noisy_val_auc'] = np.nanmean(noisy_m[:, :, 1])
            param_dict['clean_val_acc'] = np.nanmean(clean_m[:, :, 0])
            param_dict['clean_val_auc'] = np.nanmean(clean_m[:, :, 1])
            for i, label in enumerate(labels):
                param_dict['noisy_val_acc_' + label] = np.nanmean(noisy_m[:, i, 0])
                param_dict['noisy_val_auc_' + label] = np.nanmean(noisy_m[:, i, 1])
                param_dict['clean_val_acc_' + label] = np.nanmean(clean_m[:, i, 0])
                param_dict['clean_val_auc_' + label] = np.nanmean(clean_m[:, i, 1])
        return param_dict

    def get_final_results(self):
        """Best setting by `optimize_for`, retrained and scored on the held-out test rows (:582-631). Returns the
        [num_labels, 5] test metrics (None when check_test is off)."""
        best_setting = self.find_best_setting()
        print("\nThe best", self.optimize_for, "was", best_setting[self.optimize_for])
        print("It was found with the following settings:")
        print(best_setting)
        if not self.check_test:
            print("check_test is set to false, Will not evaluate performance on held-out test set.")
            return None
        dl = self.data_loader
        preds = self.test_on_test(self.convert_param_dict_for_use(dict(best_setting)))
        m = self._metrics(preds, dl.test_Y)
        names = ('Acc:', 'AUC:', 'F1:', 'Precision:', 'Recall:')
        noisy = clean = None
        if self.check_noisy_data:
            noisy = self._metrics(self.predict_on_data(dl.noisy_test_X), dl.noisy_test_Y)
            clean = self._metrics(self.predict_on_data(dl.clean_test_X), dl.clean_test_Y)
        for i, label in enumerate(LABELS_TO_PREDICT[:len(m)]):
            print("\nFINAL TEST RESULTS ON ALL", label, "DATA:")
            print(*[x for pair in zip(names, m[i]) for x in pair])
            if noisy is not None:
                print("FINAL TEST RESULTS ON NOISY", label, "DATA:")
                print(*[x for pair in zip(names, noisy[i]) for x in pair])
                print("FINAL TEST RESULTS ON CLEAN", label, "DATA:")
                print(*[x for pair in zip(names, clean[i]) for x in pair])
        print("Overall:", 'Acc:', np.mean(m[:, 0]), 'AUC:', np.mean(m[:, 1]))
        return m


if __name__ == "__main__":
    import sys
    if len(sys.argv) < 2:
        print("usage: python -m multimodalautoencoder_b200.neural_net <filename> [<continue>] [<main dir>]")
        sys.exit()
    wrapper = NNWrapper(sys.argv[1], cont=len(sys.argv) >= 3 and sys.argv[2] == 'True',
                        dropbox_path=sys.argv[3] if len(sys.argv) >= 4 else DEFAULT_MAIN_DIRECTORY)
    print("\nThe validation results dataframe will be saved in:", wrapper.results_path + wrapper.save_prefix + '.csv')
    wrapper.run()
