"""The drop-in Python surface (MultimodalAutoencoder + wrappers) on the GPU engine, against the oracle driven
through the reference's own training protocol (multimodal_autoencoder.py:549-590)."""
import os

import numpy as np
import pytest

from oracle import mmae_oracle as O

pytestmark = pytest.mark.gpu


def _loaders(n=1500, seed=11):
    from multimodalautoencoder_b200.data_funcs import DataLoader
    from multimodalautoencoder_b200.synthetic import make_frame
    df = make_frame(n, seed=seed)
    dl = DataLoader(df=df, supervised=False, cross_validation=True, normalize_and_fill=False, suppress_output=True)
    cdl = DataLoader(df=df, supervised=True, cross_validation=True, normalize_and_fill=False, suppress_output=True)
    return df, dl, cdl


def _oracle_cfg(model):
    dl = model.data_loader
    return O.OracleConfig(num_feats=dl.num_feats, layer_sizes=list(model.layer_sizes),
                          modality_starts=list(dl.modality_start_indices), modality_names=list(dl.modality_names),
                          tie_weights=model.tie_weights, variational=model.variational, activation=model.activation_func,
                          loss_func=model.loss_func, weight_penalty=model.weight_penalty, learning_rate=model.learning_rate)


@pytest.mark.parametrize('loss', ['sigmoid_cross_entropy', 'mean_squared'])
def test_train_follows_reference_protocol(loss):
    """Same np.random.seed -> same batches, same masks, same loss curve as the reference's train() loop."""
    from multimodalautoencoder_b200 import MultimodalAutoencoder
    _, dl, _ = _loaders()
    m = MultimodalAutoencoder(data_loader=dl, layer_sizes=[128, 64], variational=False, tie_weights=True, batch_size=20,
                              learning_rate=1e-3, weight_penalty=0.001, weight_initialization='normal', loss_func=loss,
                              verbose=False, precision='fp32', rng_mode='numpy')
    cfg = _oracle_cfg(m)
    P = {k: v.astype(np.float64) for k, v in m.engine.get_params().items()}
    np.random.seed(7)
    m.train(40, record_every_nth=10, save_every_nth=1000)
    # the reference's loop, restated with the oracle
    np.random.seed(7)
    st = O.AdamState()
    tl, vl = [], []
    for step in range(40):
        X = dl.get_unsupervised_train_batch(20)
        noisy = O.add_noise(cfg, X)
        if step % 10 == 0:
            a = O.forward(cfg, P, noisy, X)['recon_loss']
            vX = dl.get_unsupervised_val_batch(200)
            b = O.forward(cfg, P, O.add_noise(cfg, vX), vX)['recon_loss']
            if 'entropy' in loss:
                a, b = a / len(X), b / len(vX)
            tl.append(a); vl.append(b)
        O.train_step(cfg, P, st, noisy, X)
    assert np.allclose(m.train_loss, tl, rtol=2e-4) and np.allclose(m.val_loss, vl, rtol=2e-4)
    for k in P:
        moved = np.abs(P[k] - m.engine.get_variable(k)).max()
        assert moved <= 2e-3 * max(np.abs(P[k]).max(), 1e-3), k
    m.close()


def test_inference_surface():
    from multimodalautoencoder_b200 import MultimodalAutoencoder, get_rmse
    _, dl, _ = _loaders()
    m = MultimodalAutoencoder(data_loader=dl, layer_sizes=[128, 64], variational=False, tie_weights=False,
                              loss_func='mean_squared', weight_initialization='normal', verbose=False, precision='fp32')
    cfg = _oracle_cfg(m)
    P = {k: v.astype(np.float64) for k, v in m.engine.get_params().items()}
    X = dl.val_X[:300]
    recon, loss = m.predict(X)
    c = O.forward(cfg, P, X, X)
    assert np.abs(recon - c['decoded']).max() < 1e-4 and abs(loss - c['recon_loss']) < 1e-5 * c['recon_loss']
    rms = m.get_reconstruction_loss_per_modality(X)
    want = O.reconstruction_loss_per_modality(cfg, P, X)
    assert np.allclose(rms, want, rtol=1e-4)
    emb = m.get_embedding(X)
    assert np.abs(emb - c['emb']).max() < 1e-4
    # the TensorFlow-handle shim used by autoencoder_wrapper.py:212-226
    emb2 = m.session.run(m.embedding, {m.noisy_X: X, m.tf_dropout_prob: 1.0})
    assert np.array_equal(emb, emb2)
    r2, l2 = m.session.run([m.decoded_X, m.reconstruction_loss], {m.noisy_X: X, m.true_X: X, m.tf_dropout_prob: 1.0})
    assert np.array_equal(r2, recon) and abs(l2 - c['recon_loss']) < 1e-4 * c['recon_loss']
    filled = m.fill_missing(X)
    assert np.abs(filled - O.fill_missing(cfg, X, c['decoded'])).max() < 1e-4
    assert abs(m.get_performance_on_data(X) - c['recon_loss']) < 1e-5 * c['recon_loss']
    assert np.isfinite(m.get_performance_on_data_with_noise(X))
    # session.run([opt_step]) with explicit noisy / true feeds == one oracle step on those feeds
    np.random.seed(3)
    noisy = m.add_noise_to_batch(X)
    st = O.AdamState()
    O.train_step(cfg, P, st, noisy, X)
    m.session.run([m.opt_step], {m.noisy_X: noisy, m.true_X: X, m.tf_dropout_prob: 1.0})
    for k in P:
        assert np.abs(P[k] - m.engine.get_variable(k)).max() <= 2e-6 + 1e-4 * np.abs(P[k]).max(), k
    assert get_rmse(np.zeros(4), np.ones(4)) == 1.0
    m.close()


def test_two_phase_classification_and_checkpoint(tmp_path):
    from multimodalautoencoder_b200 import MultimodalAutoencoder
    _, dl, cdl = _loaders()
    kw = dict(data_loader=dl, classification_data_loader=cdl, layer_sizes=[200, 100], classification_layer_sizes=[50, 20],
              variational=False, tie_weights=True, batch_size=20, learning_rate=1e-3, weight_initialization='normal',
              activation_func='relu', verbose=False, checkpoint_dir=str(tmp_path) + '/', rng_mode='philox')
    m = MultimodalAutoencoder(**kw)
    m.set_classification_params(weight_penalty=0.001, learning_rate=1e-3, dropout_prob=0.5, activation_func='relu',
                                batch_size=100, loss_func='sigmoid_cross_entropy', suppress_warning=True)
    m.train(60, record_every_nth=20, save_every_nth=1000)
    assert len(m.train_loss) == 3 and m.val_loss[-1] < m.val_loss[0]
    m.train_classification(60, record_every_nth=20, save_every_nth=1000)
    assert len(m.val_acc) == 3 and all(0.0 <= a <= 1.0 for a in m.val_acc)
    preds = m.get_classification_predictions(cdl.val_X)
    assert preds.shape == (len(cdl.val_X), 3) and preds.dtype == np.int32 and set(np.unique(preds)) <= {0, 1}
    path = m.save_model()
    before = m.engine.get_params()
    m2 = MultimodalAutoencoder(**kw)
    m2.load_saved_model(directory=str(tmp_path) + '/', checkpoint_name=os.path.basename(path))
    for k, v in before.items():
        assert np.array_equal(m2.engine.get_variable(k), v), k
    assert m2.engine.get_opt_state(0, 'weights0')[2] == 60 and m2.train_loss == m.train_loss
    m.close(); m2.close()


def test_wrapper_sweeps_a_small_grid(tmp_path):
    from multimodalautoencoder_b200.autoencoder_wrapper import MMAEWrapper
    from multimodalautoencoder_b200.autoencoder_classification_wrapper import MMAEClassificationWrapper
    _, dl, cdl = _loaders(900)
    w = MMAEWrapper('synthetic.csv', dropbox_path=str(tmp_path) + '/', data_loader=dl, classification_data_loader=cdl,
                    layer_sizes=[[64, 16]], tie_weights=[True], dropout_probs=[1.0, 0.5], weight_penalties=[0.001],
                    activation_funcs=['softsign'], test_variational=True, num_cross_folds=2, num_steps=30,
                    model_kwargs=dict(rng_mode='philox'))
    assert w.num_settings == 4
    w.run()
    df = w.val_results_df
    assert len(df) == 4 and 'val_sigmoid_cross_entropy' in df.columns and 'svm_val_acc_happiness' in df.columns
    assert os.path.exists(w.results_path + w.save_prefix + '.csv')
    c = MMAEClassificationWrapper('a.csv', 'b.csv', dropbox_path=str(tmp_path) + '/', data_loader=dl,
                                  classification_data_loader=cdl, mmae_layer_sizes=[[64, 16]],
                                  classification_layer_sizes=[[10]], tie_weights=[False], mmae_dropout_probs=[1.0],
                                  mmae_weight_penalties=[0.001], mmae_test_variational=[True, False], weight_penalties=[0.0],
                                  dropout_probs=[1.0], num_cross_folds=2, mmae_num_steps=20, classification_num_steps=20,
                                  model_kwargs=dict(rng_mode='philox'))
    assert c.num_settings == 2
    c.sweep_all_parameters()
    for col in ('val_loss', 'val_acc', 'val_auc', 'val_f1', 'noisy_val_acc', 'clean_val_auc', 'val_acc_happiness'):
        assert col in c.val_results_df.columns, col


def test_predict_right_after_resident_training_is_ordered():
    """Regression: the philox path gathers its batches into buffers of its own; a host-fed forward issued right after
    queued resident train steps must neither read training rows nor corrupt the last step's layer-0 weight gradient."""
    from multimodalautoencoder_b200 import MultimodalAutoencoder
    _, dl, _ = _loaders()
    kw = dict(data_loader=dl, layer_sizes=[128, 64], variational=False, tie_weights=True, batch_size=256, learning_rate=1e-3,
              weight_initialization='normal', verbose=False, precision='fp32', rng_mode='philox', seed=3)
    X = np.ascontiguousarray(dl.val_X[:200])
    runs = []
    for synced in (True, False):
        m = MultimodalAutoencoder(**kw)
        m.train(25, record_every_nth=1000, save_every_nth=10 ** 6)
        if synced:
            m.engine.synchronize()
        rec, loss = m.predict(X)                      # batch <= train batch: no workspace growth, no implicit sync
        runs.append((rec, loss, m.engine.get_params()))
        m.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]
    for k in runs[0][2]:
        assert np.array_equal(runs[0][2][k], runs[1][2][k]), k


def test_label_shapes_are_validated():
    from multimodalautoencoder_b200 import Engine
    from tests.helpers import make_cfgs
    _, ecfg = make_cfgs(head=[8], num_labels=None, cls_loss='softmax')
    e = Engine(ecfg)
    X = np.random.default_rng(0).uniform(size=(16, 320)).astype(np.float32)
    with pytest.raises(ValueError):
        e.cls_train_step(X, np.zeros((16, 2), np.float32))          # softmax head wants [B] class indices
    with pytest.raises(ValueError):
        e.cls_train_step(X, np.full(16, 2, np.float32))             # class index out of range
    e.cls_train_step(X, np.ones(16, np.float32))
    e.close()
    _, ecfg = make_cfgs(head=[8], num_labels=3)
    e = Engine(ecfg)
    with pytest.raises(ValueError):
        e.cls_train_step(X, np.zeros(16, np.float32))               # sigmoid head wants [B, 3]
    with pytest.raises(ValueError):
        e.set_dataset(1, X, np.zeros(16, np.float32))
    e.close()


def test_rebuild_draws_fresh_weights_and_record_steps_do_not_repeat_dropout(tmp_path):
    from multimodalautoencoder_b200 import MultimodalAutoencoder
    _, dl, _ = _loaders(600)
    m = MultimodalAutoencoder(data_loader=dl, layer_sizes=[32, 8], variational=False, tie_weights=True, batch_size=20,
                              dropout_prob=0.5, weight_initialization='normal', verbose=False, rng_mode='numpy',
                              checkpoint_dir=str(tmp_path) + '/')
    w_a = m.engine.get_variable('weights0')
    m.rebuild_reinitialize()
    assert not np.array_equal(w_a, m.engine.get_variable('weights0'))           # a rebuild is a new initializer run
    np.random.seed(0)
    m.train(3, record_every_nth=1, save_every_nth=10 ** 6)                        # every step is a record step
    # host counter and engine step agree after record steps: the next descriptor upload does not rewind the engine
    steps_before = m._step_count
    m.train(1, record_every_nth=1000, save_every_nth=10 ** 6)
    assert m._step_count > steps_before
    path = m.save_model()
    m2 = MultimodalAutoencoder(data_loader=dl, layer_sizes=[32, 8], variational=False, tie_weights=True, batch_size=20,
                               dropout_prob=0.5, weight_initialization='normal', verbose=False, rng_mode='numpy',
                               checkpoint_dir=str(tmp_path) + '/')
    ref_style = os.path.basename(path)[:-4].rsplit('-', 1)
    m2.load_saved_model(directory=str(tmp_path) + '/', checkpoint_name=ref_style[0] + '.ckpt-' + ref_style[1])
    assert m2._step_count == m._step_count
    assert np.array_equal(m2.engine.get_variable('weights0'), m.engine.get_variable('weights0'))
    m.close(); m2.close()
