"""The tooling that makes the reference's own Python runnable here (oracle/build_ref.py, oracle/tf1_shim): unit checks of
the py2 -> py3 conversion rules and of the shim's op semantics against plain NumPy (the TF-1.x definitions listed in
oracle/mmae_oracle.py's header)."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


BR = _load('build_ref_under_test', os.path.join(ROOT, 'oracle', 'build_ref.py'))


def test_print_statements_become_calls():
    src = ('def f(x):\n'
           '    print "a", x\n'
           '    if x: print "b %d" % x\n'
           '    print\n'
           '    print "no newline",\n'
           '    print("already", "a call")\n'
           '    print ("tuple"), x\n'
           '    y = {"print": 1}\n'
           '    print "multi", (x +\n'
           '                    1)\n'
           '    return y\n')
    out = BR.convert(src)
    compile(out, 't', 'exec')
    assert 'print("a", x)' in out and 'if x: print("b %d" % x)' in out
    assert "end=' ')" in out and 'print()' in out
    assert 'print("already", "a call")' in out                      # a real call is left alone
    ns = {}
    exec(out, ns)
    import contextlib, io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ns['f'](3)
    assert buf.getvalue() == 'a 3\nb 3\n\nno newline already a call\ntuple 3\nmulti 4\n'


def test_except_clauses_and_tabs():
    src = ('def g():\n'
           '    """doc"""\n'
           '\ttry:\n'
           '\t\treturn 1 / 0\n'
           '\texcept ZeroDivisionError, e:\n'
           '\t\treturn str(e)\n')
    out = BR.convert(src)
    ns = {}
    exec(compile(out, 't', 'exec'), ns)
    assert 'division' in ns['g']()
    assert 'except ZeroDivisionError as e:' in out and '\t' not in out


def test_reload_is_imported():
    assert BR.convert('def r(m):\n    reload(m)\n').startswith('from importlib import reload\n')


def test_shim_op_semantics():
    tf = _load('tf1_shim_under_test', os.path.join(ROOT, 'oracle', 'tf1_shim', 'tensorflow', '__init__.py'))
    tf.set_default_dtype(torch.float64)
    rng = np.random.default_rng(0)
    l, z = rng.standard_normal((5, 7)), rng.uniform(size=(5, 7))
    g = tf.Graph()
    with g.as_default():
        L, Z, keep = tf.placeholder(tf.float32), tf.placeholder(tf.float32), tf.placeholder(tf.float32)
        sce = tf.nn.sigmoid_cross_entropy_with_logits(logits=L, labels=Z)
        drop = tf.nn.dropout(L, keep)
        w = tf.Variable(tf.constant(0.5, shape=[7, 3]), name='w')
        loss = tf.reduce_sum(tf.square(tf.matmul(L, w))) + 0.1 * tf.nn.l2_loss(w)
        opt = tf.train.AdamOptimizer(0.01)
        step = opt.minimize(loss)
        grads = tf.gradients(loss, [w])
        clipped, norm = tf.clip_by_global_norm(grads, 2.0)
        init = tf.global_variables_initializer()
    s = tf.Session(graph=g)
    s.run(init)
    got = s.run(sce, {L: l, Z: z})
    assert np.allclose(got, np.maximum(l, 0) - l * z + np.log1p(np.exp(-np.abs(l))))         # semantics (1)
    u = rng.uniform(size=(5, 7))
    tf.hooks.dropout_uniform = lambda shp, i: u
    got = s.run(drop, {L: l, keep: 0.6})
    tf.hooks.dropout_uniform = None
    assert np.allclose(got, l * np.floor(0.6 + u) / 0.6)                                       # semantics (6)
    W = np.full((7, 3), 0.5)
    gref = 2 * l.T @ (l @ W) + 0.1 * W                                                         # semantics (3): l2_loss = sum w^2 / 2
    gc, n = s.run([clipped[0], norm], {L: l})
    assert np.allclose(n, np.linalg.norm(gref)) and np.allclose(gc, gref * 2.0 / max(np.linalg.norm(gref), 2.0))
    s.run([step], {L: l})
    a = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)                                                  # semantics (4): t = 1
    m, v = 0.1 * gref, 0.001 * gref ** 2
    assert np.allclose(w.numpy(), W - a * m / (np.sqrt(v) + 1e-8))
    assert np.array_equal(s.run(tf.cast(tf.round(tf.constant([0.5, 1.5, 2.5, -0.5])), tf.int32)), [0, 2, 2, 0])   # (7) half to even
