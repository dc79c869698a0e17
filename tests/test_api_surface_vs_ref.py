"""Drop-in surface: every public class of the reference on this path, its constructor keywords WITH their defaults and
its public method names (with their positional parameters) exist in the package under the same names.  Read off the
reference's own classes (oracle/_ref, converted mechanically to Python 3) with `inspect`; nothing is instantiated, so
this runs without a GPU.  Keyword-only additions of the package are allowed, removals and changed defaults are not."""
import inspect

import pytest

from oracle.ref_loader import load_reference

REF = load_reference()
pytestmark = pytest.mark.skipif(REF is None, reason='oracle/_ref unavailable')


def pairs():
    from multimodalautoencoder_b200 import (autoencoder_classification_wrapper, autoencoder_wrapper, data_funcs,
                                            generic_wrapper, multimodal_autoencoder, neural_net)
    out = [('MultimodalAutoencoder', REF.mmae.MultimodalAutoencoder, multimodal_autoencoder.MultimodalAutoencoder),
           ('DataLoader', REF.data_funcs.DataLoader, data_funcs.DataLoader)]
    if REF.generic_wrapper is not None:
        out += [('Wrapper', REF.generic_wrapper.Wrapper, generic_wrapper.Wrapper),
                ('ClassificationWrapper', REF.generic_wrapper.ClassificationWrapper, generic_wrapper.ClassificationWrapper)]
    if REF.autoencoder_wrapper is not None:
        out += [('MMAEWrapper', REF.autoencoder_wrapper.MMAEWrapper, autoencoder_wrapper.MMAEWrapper),
                ('MMAEClassificationWrapper', REF.autoencoder_classification_wrapper.MMAEClassificationWrapper,
                 autoencoder_classification_wrapper.MMAEClassificationWrapper)]
    if REF.neural_net is not None:
        out += [('NeuralNetwork', REF.neural_net.NeuralNetwork, neural_net.NeuralNetwork),
                ('NNWrapper', REF.neural_net.NNWrapper, neural_net.NNWrapper)]
    return out


def positional(fn):
    return [p for p in inspect.signature(fn).parameters.values()
            if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.name != 'self']


# Graph-node builders: they take and return TensorFlow tensors while build_graph assembles the graph (:415-540) and
# cannot be called on data.  What they compute is reached at run time through get_embedding / predict /
# get_classification_predictions.
GRAPH_BUILDERS = {'MultimodalAutoencoder': {'apply_activation', 'build_classification_graph', 'classify', 'decode', 'encode'}}

# the reference's default main directory is a placeholder path the user edits; optimizers are TensorFlow classes
SKIP_DEFAULT = {'optimizer'}


@pytest.mark.parametrize('name', [p[0] for p in pairs()])
def test_constructor_keywords_and_defaults(name):
    _, ref, ours = next(p for p in pairs() if p[0] == name)
    rp, op = positional(ref.__init__), positional(ours.__init__)
    assert [p.name for p in op[:len(rp)]] == [p.name for p in rp], 'same keywords in the same positional order'
    for a, b in zip(rp, op):
        if a.name in SKIP_DEFAULT:
            continue
        if a.default is inspect.Parameter.empty:
            continue                       # required there; required or optional here (e.g. DataLoader(df=...) needs no file)
        assert b.default is not inspect.Parameter.empty, a.name
        if True:
            assert a.default == b.default, (name, a.name, a.default, b.default)
    for extra in op[len(rp):]:
        assert extra.default is not inspect.Parameter.empty, 'additions must be optional: %s' % extra.name


@pytest.mark.parametrize('name', [p[0] for p in pairs()])
def test_public_methods_and_their_parameters(name):
    _, ref, ours = next(p for p in pairs() if p[0] == name)
    missing, changed = [], []
    for m, fn in inspect.getmembers(ref, predicate=inspect.isfunction):
        if m.startswith('_') or m in GRAPH_BUILDERS.get(name, ()):
            continue
        if not hasattr(ours, m):
            missing.append(m)
            continue
        rp, op = positional(fn), positional(getattr(ours, m))
        if [p.name for p in op[:len(rp)]] != [p.name for p in rp]:
            changed.append((m, [p.name for p in rp], [p.name for p in op]))
        elif any(p.default is inspect.Parameter.empty for p in op[len(rp):]):
            changed.append((m, 'new required parameter'))
    assert not missing, '%s lacks the reference methods %s' % (name, missing)
    assert not changed, changed
