"""MMAEClassificationWrapper: two-phase fits (reconstruction, then the classification head) over a grid of
MMAE + head hyper-parameters (reference: autoencoder_classification_wrapper.py)."""
from __future__ import annotations

import sys

import numpy as np

from . import data_funcs
from . import generic_wrapper as gen_wrap
from . import multimodal_autoencoder as mmae
from .generic_wrapper import ClassificationWrapper

DEFAULT_MAIN_DIRECTORY = '/Your/path/here/'
DEFAULT_NUM_CROSS_FOLDS = 5
LABELS_TO_PREDICT = ['happiness', 'health', 'calmness']


class MMAEClassificationWrapper(ClassificationWrapper):
    def __init__(self, mmae_filename, classification_filename, mmae_layer_sizes=[[1000, 100], [200, 100], [500, 100]],
                 classification_layer_sizes=[[50, 20], [25, 10], [100, 50], [100]], tie_weights=[True, False],
                 mmae_dropout_probs=[1.0, 0.5], mmae_weight_penalties=[.01, .001], weight_initializers=['normal'],
                 mmae_activation_funcs=['relu'], mmae_test_variational=[True, False], weight_penalties=[0.0, .001],
                 dropout_probs=[0.5, 1.0], activation_funcs=['relu'], classification_learning_rate=.0001,
                 classification_batch_size=100, classification_num_steps=15000, cont=False,
                 classifier_name='MMAE_NN_classifier', num_cross_folds=DEFAULT_NUM_CROSS_FOLDS,
                 dropbox_path=DEFAULT_MAIN_DIRECTORY, datasets_path='Data/Cleaned/', results_path=None, check_test=False,
                 normalization='between_0_and_1', optimize_for='val_acc', min_or_max='max', save_results_every_nth=1,
                 check_noisy_data=True, wanted_label=None, shard=None, data_loader=None,
                 classification_data_loader=None, mmae_num_steps=15000, mmae_batch_size=20, model_kwargs=None):
        self.mmae_filename = mmae_filename
        self.classification_filename = classification_filename
        self.mmae_layer_sizes = mmae_layer_sizes
        self.classification_layer_sizes = classification_layer_sizes
        self.tie_weights = tie_weights
        self.mmae_dropout_probs = mmae_dropout_probs
        self.mmae_weight_penalties = mmae_weight_penalties
        self.weight_initializers = weight_initializers
        self.mmae_activation_funcs = mmae_activation_funcs
        self.mmae_test_variational = mmae_test_variational
        self.weight_penalties = weight_penalties
        self.dropout_probs = dropout_probs
        self.activation_funcs = activation_funcs
        self.classification_learning_rate = classification_learning_rate
        self.classification_num_steps = classification_num_steps
        self.classification_batch_size = classification_batch_size
        # fixed MMAE settings (autoencoder_classification_wrapper.py:90-97)
        self.mmae_loss_func = 'sigmoid_cross_entropy'
        self.mmae_learning_rate = .001
        self.mmae_num_steps = mmae_num_steps
        self.mmae_batch_size = mmae_batch_size
        self.model_kwargs = dict(model_kwargs or {})
        self._given_loaders = (data_loader, classification_data_loader)
        self.model = None
        ClassificationWrapper.__init__(
            self, filename=classification_filename, wanted_label=wanted_label, cont=cont,
            classifier_name=classifier_name, num_cross_folds=num_cross_folds, dropbox_path=dropbox_path,
            datasets_path=datasets_path, results_path=results_path, check_test=check_test, normalization=normalization,
            optimize_for=optimize_for, min_or_max=min_or_max, save_results_every_nth=save_results_every_nth,
            check_noisy_data=check_noisy_data, shard=shard)
        self.trim_extra_vae_params()

    def load_data(self):
        dl, cdl = self._given_loaders
        self.data_loader = dl if dl is not None else data_funcs.DataLoader(
            self.datasets_path + self.mmae_filename, normalize_and_fill=False, supervised=False, cross_validation=True,
            separate_noisy_data=self.check_noisy_data)
        self.classification_data_loader = cdl if cdl is not None else data_funcs.DataLoader(
            self.datasets_path + self.classification_filename, normalize_and_fill=False, cross_validation=True,
            supervised=True, separate_noisy_data=self.check_noisy_data, wanted_label=self.wanted_label)

    def define_params(self):
        self.params = {
            'mmae_architecture': self.mmae_layer_sizes, 'classification_layers': self.classification_layer_sizes,
            'tie_weights': self.tie_weights, 'mmae_dropout_prob': self.mmae_dropout_probs,
            'mmae_weight_penalty': self.mmae_weight_penalties, 'weight_initialization': self.weight_initializers,
            'mmae_activation_function': self.mmae_activation_funcs, 'variational': self.mmae_test_variational,
            'weight_penalty': self.weight_penalties, 'dropout_prob': self.dropout_probs,
            'activation_func': self.activation_funcs}

    def initialize_model(self, param_dict):
        if self.model is not None:
            self.model.close()
        self.model = mmae.MultimodalAutoencoder(
            batch_size=self.mmae_batch_size, learning_rate=self.mmae_learning_rate, model_name=self.classifier_name,
            verbose=False, loss_func=self.mmae_loss_func, checkpoint_dir=self.dropbox_path + 'temp_saved_models/',
            layer_sizes=param_dict['mmae_architecture'], classification_layer_sizes=param_dict['classification_layers'],
            variational=param_dict['variational'], tie_weights=param_dict['tie_weights'],
            dropout_prob=param_dict['mmae_dropout_prob'], weight_penalty=param_dict['mmae_weight_penalty'],
            activation_func=param_dict['mmae_activation_function'],
            weight_initialization=param_dict['weight_initialization'], data_loader=self.data_loader,
            classification_data_loader=self.classification_data_loader, **self.model_kwargs)
        # a single wanted label -> 2-logit sparse softmax head, otherwise per-label sigmoid (:169-172)
        classification_loss = 'cross_entropy' if self.wanted_label is not None else 'sigmoid_cross_entropy'
        self.model.set_classification_params(
            weight_penalty=param_dict['weight_penalty'], learning_rate=self.classification_learning_rate,
            dropout_prob=param_dict['dropout_prob'], activation_func=param_dict['activation_func'],
            batch_size=self.classification_batch_size, loss_func=classification_loss, suppress_warning=True)

    def trim_extra_vae_params(self):
        """Variational + tied weights is never tested (autoencoder_classification_wrapper.py:181-193)."""
        self.list_of_param_settings = [s for s in self.list_of_param_settings
                                       if not (s['variational'] is True and s['tie_weights'] is True)]
        self.num_settings = len(self.list_of_param_settings)

    def train_and_predict(self, param_dict, predict_on='Val'):
        if predict_on == 'Test':
            unsup_X, sup_X = self.data_loader.test_X, self.classification_data_loader.test_X
        else:
            unsup_X, sup_X = self.data_loader.val_X, self.classification_data_loader.val_X
        self.initialize_model(param_dict)
        self.model.train(self.mmae_num_steps, record_every_nth=max(self.mmae_num_steps // 10, 1),
                         save_every_nth=self.mmae_num_steps * 2)
        loss = self.model.get_performance_on_data_with_noise(unsup_X)
        self.model.train_classification(num_steps=self.classification_num_steps,
                                        record_every_nth=max(self.classification_num_steps // 10, 1),
                                        save_every_nth=self.classification_num_steps * 2)
        return loss, self.predict_on_data(sup_X)

    def predict_on_data(self, X):
        return self.model.get_classification_predictions(X)

    def _metrics(self, preds, true_y, n_labels):
        """[n_labels, 5] metric rows (acc, auc, f1, precision, recall)."""
        if self.wanted_label is None:
            return np.array([gen_wrap.compute_all_classification_metrics(preds[:, l], true_y[:, l]) for l in range(n_labels)])
        return np.array([gen_wrap.compute_all_classification_metrics(preds, true_y)])

    def get_cross_validation_results(self, param_dict):
        cdl = self.classification_data_loader
        n_labels = len(cdl.wanted_labels)
        all_m = np.full((self.num_cross_folds, n_labels, 5), np.nan)
        noisy_m = np.full((self.num_cross_folds, n_labels, 5), np.nan)
        clean_m = np.full((self.num_cross_folds, n_labels, 5), np.nan)
        all_loss = [np.nan] * self.num_cross_folds
        for f in range(self.num_cross_folds):
            self.data_loader.set_to_cross_validation_fold(f)
            cdl.set_to_cross_validation_fold(f)
            all_loss[f], preds = self.train_and_predict(param_dict)
            all_m[f] = self._metrics(preds, cdl.val_Y, n_labels)
            if self.check_noisy_data:
                noisy_m[f] = self._metrics(self.predict_on_data(cdl.noisy_val_X), cdl.noisy_val_Y, n_labels)
                clean_m[f] = self._metrics(self.predict_on_data(cdl.clean_val_X), cdl.clean_val_Y, n_labels)
        param_dict['val_loss'] = np.nanmean(all_loss)
        for j, k in enumerate(('acc', 'auc', 'f1', 'precision', 'recall')):
            param_dict['val_' + k] = np.nanmean(all_m[:, :, j])
        if self.wanted_label is None:
            for i, label in enumerate(LABELS_TO_PREDICT[:n_labels]):
                param_dict['val_acc_' + label] = np.nanmean(all_m[:, i, 0])
                param_dict['val_auc_' + label] = np.nanmean(all_m[:, i, 1])
        if self.check_noisy_data:
            param_dict['noisy_val_acc'] = np.nanmean(noisy_m[:, :, 0])
            param_dict['noisy_val_auc'] = np.nanmean(noisy_m[:, :, 1])
            param_dict['clean_val_acc'] = np.nanmean(clean_m[:, :, 0])
            param_dict['clean_val_auc'] = np.nanmean(clean_m[:, :, 1])
            if self.wanted_label is None:      # column names kept as the reference wrote them ('svm_' prefix, :318-326)
                for i, label in enumerate(LABELS_TO_PREDICT[:n_labels]):
                    param_dict['svm_noisy_val_acc_' + label] = np.nanmean(noisy_m[:, i, 0])
                    param_dict['svm_noisy_val_auc_' + label] = np.nanmean(noisy_m[:, i, 1])
                    param_dict['svm_clean_val_acc_' + label] = np.nanmean(clean_m[:, i, 0])
                    param_dict['svm_clean_val_auc_' + label] = np.nanmean(clean_m[:, i, 1])
        return param_dict

    def get_final_results(self):
        best_setting = self.find_best_setting()
        if not self.check_test:
            return
        loss, preds = self.test_on_test(self.convert_param_dict_for_use(best_setting.to_dict()))
        true_y = self.classification_data_loader.test_Y
        m = self._metrics(preds, true_y, len(self.classification_data_loader.wanted_labels))
        print("\nFINAL TEST RESULTS: loss", loss, "acc", np.nanmean(m[:, 0]), "auc", np.nanmean(m[:, 1]))
        return loss, m

    def test_on_test(self, param_dict):
        return self.train_and_predict(param_dict, predict_on='Test')


if __name__ == "__main__":
    if len(sys.argv) < 3:
        print("usage: python -m multimodalautoencoder_b200.autoencoder_classification_wrapper "
              "<mmae_filename> <classification_filename> [cont|<label>] [main_directory]")
        sys.exit()
    extra = sys.argv[3] if len(sys.argv) >= 4 else ''
    cont = 'true' in extra.lower() or 'cont' in extra.lower()
    label = None if (cont or not extra) else extra
    main_dir = sys.argv[4] if len(sys.argv) >= 5 else DEFAULT_MAIN_DIRECTORY
    MMAEClassificationWrapper(sys.argv[1], sys.argv[2], cont=cont, wanted_label=label, dropbox_path=main_dir).run()
