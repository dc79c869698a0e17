"""MMAEWrapper: grid search + cross-validation over MMAE hyper-parameters, scoring each setting by its noisy
validation loss and by SVMs fitted on the embeddings (reference: autoencoder_wrapper.py)."""
from __future__ import annotations

import sys

import numpy as np

from . import data_funcs
from . import generic_wrapper as gen_wrap
from . import multimodal_autoencoder as mmae
from .generic_wrapper import Wrapper

DEFAULT_MAIN_DIRECTORY = '/Your/path/here/'
DEFAULT_NUM_CROSS_FOLDS = 5
LABELS_TO_PREDICT = ['happiness', 'health', 'calmness']


class MMAEWrapper(Wrapper):
    def __init__(self, filename, classification_filename='modalities_missing.csv',
                 layer_sizes=[[1000, 100], [500, 100], [300, 100]], tie_weights=[True, False], dropout_probs=[1.0, 0.5],
                 weight_penalties=[0.0, .01, .001], weight_initializers=['normal'], activation_funcs=['softsign', 'relu'],
                 test_variational=True, cont=False, classifier_name='MMAE', num_cross_folds=DEFAULT_NUM_CROSS_FOLDS,
                 dropbox_path=DEFAULT_MAIN_DIRECTORY, datasets_path='Data/Cleaned/', results_path=None,
                 temp_model_path='Results/temp_saved_models', check_test=False, optimize_for=None, min_or_max='min',
                 save_results_every_nth=1, shard=None, data_loader=None, classification_data_loader=None,
                 num_steps=15000, batch_size=20, model_kwargs=None, svm_scoring=True):
        self.temp_model_path = dropbox_path + temp_model_path
        self.classification_filename = filename if classification_filename is None else classification_filename
        self.layer_sizes = layer_sizes
        self.tie_weights = tie_weights
        self.dropout_probs = dropout_probs
        self.weight_penalties = weight_penalties
        self.weight_initializers = weight_initializers
        self.activation_funcs = activation_funcs
        self.test_variational = test_variational
        # fixed hyper-parameters (autoencoder_wrapper.py:80-92)
        self.loss_func = 'sigmoid_cross_entropy'
        self.learning_rate = .001
        self.clip_gradients = True
        self.normalization = 'between_0_and_1'
        self.mask_with = -1.0
        self.fill_missing = 0.0
        self.decay, self.decay_steps, self.decay_rate = True, 1000, 0.95
        self.batch_size = batch_size
        self.num_steps = num_steps
        self.model_kwargs = dict(model_kwargs or {})
        self.svm_scoring = svm_scoring
        self._given_loaders = (data_loader, classification_data_loader)
        self.model = None
        if optimize_for is None:
            optimize_for = 'val_' + self.loss_func
        Wrapper.__init__(self, filename=filename, cont=cont, classifier_name=classifier_name,
                         num_cross_folds=num_cross_folds, dropbox_path=dropbox_path, datasets_path=datasets_path,
                         results_path=results_path, check_test=check_test, optimize_for=optimize_for,
                         min_or_max=min_or_max, normalization=self.normalization,
                         save_results_every_nth=save_results_every_nth, shard=shard)
        if self.test_variational:
            self.add_extra_vae_params()

    def load_data(self):
        dl, cdl = self._given_loaders
        self.data_loader = dl if dl is not None else data_funcs.DataLoader(
            self.datasets_path + self.filename, normalize_and_fill=False, supervised=False, cross_validation=True,
            normalization=self.normalization, fill_missing_with=self.fill_missing)
        self.classification_data_loader = cdl if cdl is not None else data_funcs.DataLoader(
            self.datasets_path + self.classification_filename, normalize_and_fill=False, supervised=True,
            cross_validation=True, normalization=self.normalization, fill_missing_with=self.fill_missing,
            separate_noisy_data=True)

    def define_params(self):
        self.params = {'architecture': self.layer_sizes, 'tie_weights': self.tie_weights,
                       'dropout_prob': self.dropout_probs, 'weight_penalty': self.weight_penalties,
                       'weight_initialization': self.weight_initializers, 'activation_function': self.activation_funcs,
                       'variational': [False]}

    def add_extra_vae_params(self):
        """VAEs cannot tie weights, so their settings are appended separately (autoencoder_wrapper.py:138-155)."""
        for arch in self.layer_sizes:
            for act in self.activation_funcs:
                for dprob in self.dropout_probs:
                    for wpen in self.weight_penalties:
                        for winit in self.weight_initializers:
                            self.list_of_param_settings.append({
                                'activation_function': act, 'architecture': arch, 'dropout_prob': dprob,
                                'tie_weights': False, 'variational': True, 'weight_initialization': winit,
                                'weight_penalty': wpen})
        self.num_settings = len(self.list_of_param_settings)

    def initialize_model(self, param_dict):
        if self.model is not None:
            self.model.close()            # the reference leaked one tf.Session per fit
        self.model = mmae.MultimodalAutoencoder(
            batch_size=self.batch_size, learning_rate=self.learning_rate, decay=self.decay,
            decay_steps=self.decay_steps, decay_rate=self.decay_rate, clip_gradients=self.clip_gradients,
            normalization=self.normalization, subdivide_physiology=True, fill_missing_with=self.fill_missing,
            mask_with=self.mask_with, checkpoint_dir=self.temp_model_path, model_name='MMAE', loss_func=self.loss_func,
            verbose=False, layer_sizes=param_dict['architecture'], variational=param_dict['variational'],
            tie_weights=param_dict['tie_weights'], dropout_prob=param_dict['dropout_prob'],
            weight_penalty=param_dict['weight_penalty'], activation_func=param_dict['activation_function'],
            weight_initialization=param_dict['weight_initialization'], data_loader=self.data_loader, **self.model_kwargs)

    def train_and_predict(self, param_dict):
        self.initialize_model(param_dict)
        self.model.train(self.num_steps, record_every_nth=max(self.num_steps // 10, 1), save_every_nth=self.num_steps + 1)
        loss = self.model.get_performance_on_data_with_noise(self.data_loader.val_X)
        print("\tLoss on fold", self.model.data_loader.fold, "was", loss)
        return loss

    def test_embedding_classification_quality(self):
        """SVMs on the embeddings of the classification data: (acc, auc) on all / noisy / clean validation rows."""
        from sklearn.svm import SVC
        assert len(self.model.val_loss) > 0, "Model needs to be trained before embeddings can be tested"
        m, cdl = self.model, self.classification_data_loader
        # ONE embedding pass for the four row sets (the reference ran four session.run calls, autoencoder_wrapper.py:
        # 212-226): the rows are stacked, encoded once through the TensorFlow-handle shim, and split again.  Row-wise
        # results do not depend on what else is in the batch (a VAE's epsilon is the only batch-indexed draw).
        parts = [np.asarray(a, np.float64) for a in (cdl.train_X, cdl.val_X, cdl.clean_val_X, cdl.noisy_val_X)]
        emb_all = m.session.run(m.embedding, {m.noisy_X: np.concatenate(parts, axis=0), m.tf_dropout_prob: 1.0})
        cuts = np.cumsum([len(a) for a in parts])[:-1]
        emb_train, emb_val, emb_clean, emb_noisy = np.split(emb_all, cuts, axis=0)
        n = len(LABELS_TO_PREDICT)
        best = np.zeros((6, n))
        sets = ((emb_val, cdl.val_Y), (emb_noisy, cdl.noisy_val_Y), (emb_clean, cdl.clean_val_Y))
        for l in range(n):
            for C in [1.0, 10.0, 100.0]:
                for b in [.01, .001]:
                    try:
                        svm_model = SVC(C=C, kernel='rbf', gamma=b).fit(emb_train, cdl.train_Y[:, l])
                        for k, (X, Y) in enumerate(sets):
                            best[2 * k, l], best[2 * k + 1, l] = self.svm_pred_best_result(
                                svm_model, X, Y, l, best[2 * k, l], best[2 * k + 1, l])
                    except Exception as e:
                        print("Error! Could not fit SVM model:", e)
        return tuple(np.atleast_2d(best[i]) for i in range(6))

    def svm_pred_best_result(self, svm_model, X, Y, label, best_acc, best_auc):
        acc, auc = gen_wrap.compute_all_classification_metrics(svm_model.predict(X), Y[:, label])[:2]
        if acc > best_acc and auc > best_auc:
            return acc, auc
        return best_acc, best_auc

    def get_cross_validation_results(self, param_dict):
        losses, folds = [], [None] * 6
        for f in range(self.num_cross_folds):
            self.data_loader.set_to_cross_validation_fold(f)
            self.classification_data_loader.set_to_cross_validation_fold(f)
            losses.append(self.train_and_predict(param_dict))
            if self.svm_scoring:
                res = self.test_embedding_classification_quality()
                folds = [self.append_fold_results(a, r) for a, r in zip(folds, res)]
        print("Losses for each fold:", losses)
        param_dict[self.optimize_for] = np.mean(losses)
        if self.svm_scoring:
            accs, aucs, nacc, nauc, cacc, cauc = folds
            for i, label in enumerate(LABELS_TO_PREDICT):
                param_dict['svm_val_acc_' + label] = np.nanmean(accs[:, i])
                param_dict['svm_val_auc_' + label] = np.nanmean(aucs[:, i])
                param_dict['svm_noisy_val_acc_' + label] = np.nanmean(nacc[:, i])
                param_dict['svm_noisy_val_auc_' + label] = np.nanmean(nauc[:, i])
                param_dict['svm_clean_val_acc_' + label] = np.nanmean(cacc[:, i])
                param_dict['svm_clean_val_auc_' + label] = np.nanmean(cauc[:, i])
            for name, arr in (('svm_val_acc', accs), ('svm_val_auc', aucs), ('svm_noisy_val_acc', nacc),
                              ('svm_noisy_val_auc', nauc), ('svm_clean_val_acc', cacc), ('svm_clean_val_auc', cauc)):
                param_dict[name] = np.nanmean(arr)
        return param_dict

    def append_fold_results(self, all_results, fold_results):
        return fold_results if all_results is None else np.concatenate([all_results, fold_results], axis=0)

    def test_on_test(self, param_dict):
        self.train_and_predict(param_dict)
        loss = self.model.get_performance_on_data(self.data_loader.test_X)
        print("\nFINAL TEST RESULTS:", self.loss_func, loss)
        return loss

    def run(self):
        self.sweep_all_parameters()
        self.get_final_results()
        if self.svm_scoring:
            for metric in ['svm_val_acc', 'svm_val_auc']:
                self.find_best_setting(optimize_for=metric, min_or_max='max')


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print("usage: python -m multimodalautoencoder_b200.autoencoder_wrapper <filename> [True] [main_directory]")
        sys.exit()
    main_dir = sys.argv[3] if len(sys.argv) >= 4 else DEFAULT_MAIN_DIRECTORY
    MMAEWrapper(sys.argv[1], cont=(len(sys.argv) >= 3 and sys.argv[2] == 'True'), dropbox_path=main_dir).run()
