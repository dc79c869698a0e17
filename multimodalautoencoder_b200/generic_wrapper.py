"""Grid-search + k-fold cross-validation drivers (reference: generic_wrapper.py).

Same classes, constructor keywords, override points (define_params / train_and_predict / test_on_test /
predict_on_data) and results-CSV columns as the reference.  Python-3 / pandas >= 2 host code; the only
functional addition is `shard=(rank, world)`: sweep_all_parameters then tests the settings
i with i % world == rank and writes its own results CSV, so the independent fits of one grid spread over
the GPUs of a box without any collective (BASELINE.json config 3); merge_shard_results joins the CSVs.
"""
from __future__ import annotations

import ast
import copy
import itertools
import os
import sys
from time import time

import numpy as np
import pandas as pd
from sklearn.metrics import f1_score, precision_score, recall_score, roc_auc_score

from . import data_funcs
from . import helper_funcs as helper

DEFAULT_MAIN_DIRECTORY = '/Your/path/here/'
DEFAULT_NUM_CROSS_FOLDS = 5


class Wrapper:
    def __init__(self, filename, cont=False, classifier_name='MMAE', num_cross_folds=DEFAULT_NUM_CROSS_FOLDS,
                 dropbox_path=DEFAULT_MAIN_DIRECTORY, datasets_path='Data/', results_path=None, check_test=False,
                 normalize_and_fill=False, normalization='between_0_and_1', optimize_for='val_score', min_or_max='max',
                 save_results_every_nth=1, cross_validation=True, shard=None):
        self.filename = filename
        self.cont = cont
        self.classifier_name = classifier_name
        self.num_cross_folds = num_cross_folds
        self.dropbox_path = dropbox_path
        self.datasets_path = dropbox_path + datasets_path
        self.results_path = dropbox_path + ('Results/' + classifier_name + '/' if results_path is None else results_path)
        self.check_test = check_test
        self.save_results_every_nth = save_results_every_nth
        self.optimize_for = optimize_for
        self.normalize_and_fill = normalize_and_fill
        self.normalization = normalization
        self.min_or_max = min_or_max
        self.cross_validation = cross_validation
        self.shard = tuple(shard) if shard is not None else None

        self.save_prefix = self.get_save_prefix(filename, replace=cont)
        self.params = {}
        self.define_params()
        self.load_data()
        self.construct_list_of_params_to_test()
        self.num_settings = len(self.list_of_param_settings)

        self.time_sum = 0
        if cont and os.path.exists(self._results_file()):
            self.val_results_df = pd.read_csv(self._results_file(), index_col=0)
            print('\nPrevious validation results df loaded. It has', len(self.val_results_df), "rows")
        else:
            self.val_results_df = pd.DataFrame()
        self.started_from = len(self.val_results_df)

    # ---- to be provided by the child class
    def define_params(self):
        raise NotImplementedError("define_params should be overwritten in child class")

    def train_and_predict(self, param_dict):
        raise NotImplementedError("train_and_predict should be overwritten in child class")

    def test_on_test(self, param_dict):
        raise NotImplementedError("test_on_test should be overwritten in child class")

    # ---- shared machinery
    def load_data(self):
        self.data_loader = data_funcs.DataLoader(self.datasets_path + self.filename,
                                                 normalize_and_fill=self.normalize_and_fill,
                                                 cross_validation=self.cross_validation,
                                                 normalization=self.normalization)

    def construct_list_of_params_to_test(self):
        self.list_of_param_settings = []
        self.recurse_and_append_params(copy.deepcopy(self.params), {})

    def recurse_and_append_params(self, param_settings_left, this_param_dict, debug=False):
        """Every combination of the per-parameter value lists (the reference's recursive enumeration,
        generic_wrapper.py:149-185; row *order* depended on Python-2 hash order and is not a contract)."""
        keys = [k for k in self.params.keys() if k not in this_param_dict]
        for combo in itertools.product(*[param_settings_left[k] for k in keys]):
            d = dict(this_param_dict)
            d.update(zip(keys, combo))
            self.list_of_param_settings.append(d)

    def _results_file(self):
        suffix = '' if self.shard is None else '-rank%d' % self.shard[0]
        return self.results_path + self.save_prefix + suffix + '.csv'

    def get_save_prefix(self, filename, replace=False):
        prefix = self.classifier_name + '-' + filename[0:filename.find('.')]
        if not replace:
            while os.path.exists(self.results_path + prefix + '.csv'):
                prefix = prefix + '2'
        return prefix

    def setting_already_done(self, param_dict):
        mini = self.val_results_df
        for key, setting in param_dict.items():
            if key not in mini.columns:
                return False
            if isinstance(setting, list):
                setting = str(setting)
            mini = mini[mini[key] == setting]
            if len(mini) == 0:
                return False
        print("Setting already tested")
        return True

    def convert_param_dict_for_use(self, setting_dict):
        """Values read back from a results CSV arrive as strings; turn them into objects again."""
        setting_dict = dict(setting_dict)
        for key in ('architecture', 'mmae_architecture', 'classification_layers'):
            if isinstance(setting_dict.get(key), str):
                setting_dict[key] = ast.literal_eval(setting_dict[key])
        if 'batch_size' in setting_dict:
            setting_dict['batch_size'] = int(setting_dict['batch_size'])
        return setting_dict

    def my_settings(self):
        if self.shard is None:
            return list(self.list_of_param_settings)
        rank, world = self.shard
        return [s for i, s in enumerate(self.list_of_param_settings) if i % world == rank]

    def sweep_all_parameters(self):
        mine = self.my_settings()
        print("\nYou have chosen to test a total of", self.num_settings, "settings" +
              ('' if self.shard is None else ' (%d on this shard)' % len(mine)))
        sys.stdout.flush()
        for param_dict in mine:
            self.test_one_setting(param_dict)
        self._write_results()
        print("\n--------------PARAMETER SWEEP IS COMPLETE--------------")

    def _write_results(self):
        os.makedirs(self.results_path, exist_ok=True)
        self.val_results_df.to_csv(self._results_file())

    def merge_shard_results(self, world=None):
        """Rank-0 helper: concatenates the per-rank CSVs of a sharded sweep into the usual results file."""
        world = world or (self.shard[1] if self.shard else 1)
        frames = []
        for r in range(world):
            f = self.results_path + self.save_prefix + '-rank%d.csv' % r
            if os.path.exists(f):
                frames.append(pd.read_csv(f, index_col=0))
        merged = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
        merged.to_csv(self.results_path + self.save_prefix + '.csv')
        return merged

    def test_one_setting(self, param_dict):
        if self.cont and self.setting_already_done(param_dict):
            return
        t0 = time()
        results_dict = self.get_cross_validation_results(dict(param_dict))
        row = {k: (str(v) if isinstance(v, list) else v) for k, v in results_dict.items()}
        self.val_results_df = pd.concat([self.val_results_df, pd.DataFrame([row])], ignore_index=True)
        this_time = time() - t0
        self.time_sum += this_time
        print("\n", self.val_results_df.tail(n=1))
        print("It took", this_time, "seconds to obtain this result")
        self.print_time_estimate()
        sys.stdout.flush()
        if len(self.val_results_df) % self.save_results_every_nth == 0:
            self._write_results()

    def get_cross_validation_results(self, param_dict):
        scores = []
        for f in range(self.num_cross_folds):
            self.data_loader.set_to_cross_validation_fold(f)
            scores.append(self.train_and_predict(param_dict))
        print("Scores for each fold:", scores)
        param_dict[self.optimize_for] = np.mean(scores)
        return param_dict

    def print_time_estimate(self):
        num_done = len(self.val_results_df) - self.started_from
        num_remaining = len(self.my_settings()) - num_done - self.started_from
        avg_time = self.time_sum / max(num_done, 1)
        hours, mins, secs = helper.get_secs_mins_hours_from_secs(int(avg_time * max(num_remaining, 0)))
        print("\n", num_done, "settings processed so far,", num_remaining, "left to go")
        print("Estimated time remaining:", hours, "hours", mins, "mins", secs, "secs")

    def find_best_setting(self, optimize_for=None, min_or_max=None):
        optimize_for = optimize_for or self.optimize_for
        min_or_max = min_or_max or self.min_or_max
        scores = self.val_results_df[optimize_for].tolist()
        best_score = min(scores) if min_or_max == 'min' else max(scores)
        best_setting = self.val_results_df.iloc[scores.index(best_score)]
        print("\nThe best", optimize_for, "was", best_setting[optimize_for])
        print("It was found with the following settings:")
        print(best_setting, "\n")
        return best_setting

    def get_final_results(self):
        best_setting = self.find_best_setting()
        if not self.check_test:
            print("check_test is set to false, Will not evaluate performance on held-out test set.")
            return
        print("\nAbout to evaluate results on held-out test set with the best", self.optimize_for)
        test_score = self.test_on_test(self.convert_param_dict_for_use(best_setting.to_dict()))
        print("\nFINAL TEST RESULTS:", test_score)
        return test_score

    def run(self):
        self.sweep_all_parameters()
        self.get_final_results()


class ClassificationWrapper(Wrapper):
    def __init__(self, filename, wanted_label=None, cont=False, classifier_name='SVM',
                 num_cross_folds=DEFAULT_NUM_CROSS_FOLDS, dropbox_path=DEFAULT_MAIN_DIRECTORY, datasets_path='Data/',
                 results_path=None, check_test=False, normalize_and_fill=False, normalization='z_score',
                 optimize_for='val_acc', min_or_max='max', save_results_every_nth=1, check_noisy_data=False,
                 cross_validation=True, shard=None):
        self.wanted_label = wanted_label
        self.check_noisy_data = check_noisy_data
        Wrapper.__init__(self, filename=filename, cont=cont, classifier_name=classifier_name,
                         num_cross_folds=num_cross_folds, dropbox_path=dropbox_path, datasets_path=datasets_path,
                         results_path=results_path, check_test=check_test, normalize_and_fill=normalize_and_fill,
                         normalization=normalization, optimize_for=optimize_for, min_or_max=min_or_max,
                         save_results_every_nth=save_results_every_nth, cross_validation=cross_validation, shard=shard)

    def predict_on_data(self, X):
        raise NotImplementedError("predict_on_data should be overwritten in child class")

    def load_data(self):
        self.data_loader = data_funcs.DataLoader(self.datasets_path + self.filename,
                                                 normalize_and_fill=self.normalize_and_fill,
                                                 cross_validation=self.cross_validation, supervised=True,
                                                 wanted_label=self.wanted_label, normalization=self.normalization,
                                                 separate_noisy_data=self.check_noisy_data)

    def get_save_prefix(self, filename, replace=False):
        prefix = self.classifier_name + '-' + filename[0:filename.find('.')]
        if self.wanted_label is not None:
            prefix += '-' + helper.get_friendly_label_name(self.wanted_label)
        if not replace:
            while os.path.exists(self.results_path + prefix + '.csv'):
                prefix = prefix + '2'
        return prefix

    def get_cross_validation_results(self, param_dict):
        cols = {k: [] for k in ('acc', 'auc', 'f1', 'precision', 'recall')}
        noisy = {k: [] for k in ('noisy_acc', 'noisy_auc', 'clean_acc', 'clean_auc')}
        for f in range(self.num_cross_folds):
            self.data_loader.set_to_cross_validation_fold(f)
            preds = self.train_and_predict(param_dict)
            true_y = self.data_loader.val_Y
            if preds is None or true_y is None:
                continue
            for k, v in zip(cols, compute_all_classification_metrics(preds, true_y)):
                cols[k].append(v)
            if self.check_noisy_data:
                m = compute_all_classification_metrics(self.predict_on_data(self.data_loader.noisy_val_X),
                                                       self.data_loader.noisy_val_Y)
                noisy['noisy_acc'].append(m[0]); noisy['noisy_auc'].append(m[1])
                m = compute_all_classification_metrics(self.predict_on_data(self.data_loader.clean_val_X),
                                                       self.data_loader.clean_val_Y)
                noisy['clean_acc'].append(m[0]); noisy['clean_auc'].append(m[1])
        for k, v in cols.items():
            param_dict['val_' + k] = np.nanmean(v) if v else np.nan
        if self.check_noisy_data:
            for k, v in noisy.items():
                name = k.replace('_acc', '_val_acc').replace('_auc', '_val_auc')
                param_dict[name] = np.nanmean(v) if v else np.nan
        return param_dict

    def get_classification_predictions_from_df(self):
        df = copy.deepcopy(self.data_loader.df)
        preds = self.predict_on_data(df[self.data_loader.wanted_feats].to_numpy())
        assert len(df) == len(preds)
        for i, label in enumerate(self.data_loader.wanted_labels):
            df['predictions_' + label] = preds[:, i] if np.ndim(preds) > 1 else preds
        return df

    def get_final_results(self):
        best_setting = None
        for metric in ['val_acc', 'noisy_val_acc', 'clean_val_acc']:
            if metric in self.val_results_df.columns.values:
                best_setting = self.find_best_setting(optimize_for=metric, min_or_max='max')
        if not self.check_test or best_setting is None:
            print("check_test is set to false, Will not evaluate performance on held-out test set.")
            return
        preds = self.test_on_test(self.convert_param_dict_for_use(best_setting.to_dict()))
        res = compute_all_classification_metrics(preds, self.data_loader.test_Y)
        print("\nFINAL TEST RESULTS ON ALL DATA:", dict(zip(('acc', 'auc', 'f1', 'precision', 'recall'), res)))
        return res


def get_baseline(Y):
    """Share of the most frequent class."""
    Y = list(np.asarray(Y).tolist())
    p = float(Y.count(1.0)) / float(len(Y))
    return max(p, 1.0 - p)


def compute_classification_metric(metric, true_y, preds):
    try:
        return metric(true_y, preds)
    except Exception as e:       # e.g. AUC with one class present
        print("Error in computing metric:", e)
        return np.nan


def binary_accuracy(true_y, preds):
    assert len(preds) == len(true_y)
    return float(np.mean(np.asarray(preds) == np.asarray(true_y)))


def compute_all_classification_metrics(preds, true_y):
    """accuracy, AUC, F1, precision, recall (generic_wrapper.py:591-603)."""
    return tuple(compute_classification_metric(m, true_y, preds)
                 for m in (binary_accuracy, roc_auc_score, f1_score, precision_score, recall_score))
