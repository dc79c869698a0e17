"""NeuralNetwork (the reference's plain MLP classifier, comparison_algorithms/neural_net.py:27-381) on the engine, against
the reference's OWN class run through oracle/_ref + the TF-1 shim (fp64): same data, same initial weights, same
np.random stream -> same batches.  Checked: the clipped gradients of one step (tf.clip_by_global_norm, active and
inactive), the accuracy curves of the reference's train() protocol, the live global_step, the parameters' movement."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle.ref_loader import load_reference

pytestmark = pytest.mark.gpu
REF = load_reference(torch.float64)


def _pair(layer_sizes, lam, keep=1.0, lr=1e-3, seed=0, clip=True):
    from multimodalautoencoder_b200.data_funcs import DataLoader
    from multimodalautoencoder_b200.neural_net import NeuralNetwork
    from multimodalautoencoder_b200.synthetic import make_frame
    df = make_frame(900, seed=21)
    dl = DataLoader(df=df, supervised=True, cross_validation=False, normalize_and_fill=False, suppress_output=True)
    ours = NeuralNetwork(data_loader=dl, layer_sizes=layer_sizes, batch_size=20, learning_rate=lr, dropout_prob=keep,
                         weight_penalty=lam, clip_gradients=clip, verbose=False, checkpoint_dir='', precision='fp32', seed=seed)
    rdl = REF.make_loader(dl.train_X, dl.val_X, dl.modality_start_indices, dl.modality_names, dl.train_Y, dl.val_Y, 3)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = REF.neural_net.NeuralNetwork(data_loader=rdl, layer_sizes=layer_sizes, batch_size=20, learning_rate=lr,
                                           dropout_prob=keep, weight_penalty=lam, clip_gradients=clip, verbose=False,
                                           checkpoint_dir='/tmp/mmae_nn_ref_ckpt/')
    W0 = ours.get_variables()
    rv = REF.nn_variables(ref)
    assert set(rv) == set(W0)
    for k, v in W0.items():
        rv[k].load(v.astype(np.float64))
    return ours, ref, dl, W0


@pytest.mark.skipif(REF is None or REF.neural_net is None, reason='oracle/_ref/neural_net.py unavailable')
@pytest.mark.parametrize('lam,clipped', [(0.001, False), (1.0, True)])
def test_one_step_gradients_and_clip(lam, clipped):
    ours, ref, dl, W0 = _pair([64, 32], lam)
    X, Y = dl.train_X[:20], dl.train_Y[:20]
    ref.session.run([ref.opt_step], {ref.tf_X: X, ref.tf_Y: Y, ref.tf_dropout_prob: 1.0})
    want = ref.tf_optimizer.last_grads                      # what apply_gradients received: clipped, L2 term included
    ours.engine.cls_train_step(X.astype(np.float32), Y.astype(np.float32))
    names = ours._names()
    raw = {r: ours.engine.get_gradient(e).astype(np.float64) + (lam * W0[r] if r.startswith('weights') else 0.0) for r, e in names.items()}
    gnorm = np.sqrt(sum(np.sum(g * g) for g in raw.values()))
    assert (gnorm > 5.0) == clipped
    scale = 5.0 / max(gnorm, 5.0)
    for r in names:
        assert np.linalg.norm(raw[r] * scale - want[r]) <= 1e-4 * max(np.linalg.norm(want[r]), 1e-12), r
    # TF-Adam on the clipped gradient (first step: m = 0.1 g, v = 0.001 g^2)
    a = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    now = ours.get_variables()
    for r in names:
        g = raw[r] * scale
        exp = W0[r] - a * 0.1 * g / (np.sqrt(0.001 * g * g) + 1e-8)
        ok = np.abs(g) > 1e-4 * np.abs(g).max()
        assert np.abs(now[r] - exp)[ok].max() <= 2e-6 + 1e-6 * np.abs(W0[r]).max(), r
    ours.close()


@pytest.mark.skipif(REF is None or REF.neural_net is None, reason='oracle/_ref/neural_net.py unavailable')
def test_train_protocol_matches_reference():
    import os
    os.makedirs('/tmp/mmae_nn_ref_ckpt', exist_ok=True)
    ours, ref, dl, W0 = _pair([64, 32], 0.01, lr=1e-3)
    np.random.seed(4)
    ours.train(60, output_every_nth=20)
    np.random.seed(4)
    with contextlib.redirect_stdout(io.StringIO()):
        ref.train(60, output_every_nth=20)
    assert ours.global_step == 60 == int(ref.session.run(ref.global_step))                 # the global step is live (:192-193)
    assert len(ours.train_acc) == len(ref.train_acc) == 3
    assert np.allclose(ours.train_acc, ref.train_acc, atol=0.051) and np.allclose(ours.val_acc, ref.val_acc, atol=0.01)
    rv = REF.nn_variables(ref)
    now = ours.get_variables()
    for k in W0:
        moved = np.linalg.norm(rv[k].numpy() - W0[k])
        assert np.linalg.norm(now[k] - rv[k].numpy()) <= 0.05 * moved + 1e-6, k
    preds, probs = ours.predict(dl.val_X, get_probabilities=True)
    rp = ref.predict(dl.val_X)
    assert preds.shape == rp.shape and np.mean(preds == rp) > 0.99
    assert 0.0 <= ours.test_on_validation() <= 1.0
    ours.close()


def test_nn_wrapper_sweep_runs_the_engine(tmp_path):
    """NNWrapper (comparison_algorithms/neural_net.py:407-631) on the engine: a 2-setting, 2-fold sweep writes one row per
    setting with the reference's columns; what a fold trains is exactly a stand-alone NeuralNetwork with the same
    hyper-parameters on that fold (same np.random stream -> identical predictions); the previous engine is destroyed."""
    from multimodalautoencoder_b200.data_funcs import DataLoader
    from multimodalautoencoder_b200.neural_net import NeuralNetwork, NNWrapper
    from multimodalautoencoder_b200.synthetic import make_frame
    df = make_frame(600, seed=8)
    dl = DataLoader(df=df, supervised=True, cross_validation=True, normalize_and_fill=False, suppress_output=True)
    with contextlib.redirect_stdout(io.StringIO()):
        w = NNWrapper('synthetic.csv', layer_sizes=[[32, 16]], dropout_probs=[1.0], weight_penalties=[0.0, .001],
                      batch_sizes=[50], num_steps=40, num_cross_folds=2, dropbox_path=str(tmp_path) + '/',
                      data_loader=dl, check_test=True)
        w.run()
    res = w.val_results_df
    assert len(res) == 2 and {'val_acc', 'val_auc', 'noisy_val_acc', 'clean_val_acc', 'val_acc_happiness'} <= set(res.columns)
    assert ((res['val_acc'] >= 0) & (res['val_acc'] <= 1)).all()
    assert w.model.engine.kernel_launches > 0

    params = dict(w.list_of_param_settings[1])
    dl.set_to_cross_validation_fold(1)
    np.random.seed(3)
    first = w.model
    got = w.train_and_predict(params)
    assert first.engine is None                      # closed when the next model was built
    np.random.seed(3)
    alone = NeuralNetwork(data_loader=dl, layer_sizes=params['architecture'], batch_size=params['batch_size'],
                          learning_rate=params['learning_rate'], dropout_prob=params['dropout_prob'],
                          weight_penalty=params['weight_penalty'], verbose=False, checkpoint_dir=None)
    alone.train(num_steps=40, output_every_nth=5001)
    assert np.array_equal(got, alone.predict(dl.val_X))
    alone.close()
