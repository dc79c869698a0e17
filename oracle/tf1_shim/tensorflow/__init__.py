"""Minimal TensorFlow-1.x API on torch autograd.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Exists for one purpose: to let the reference's OWN graph-building code
(`oracle/_ref/multimodal_autoencoder.py`, converted mechanically from /root/reference by
`oracle/build_ref.py`) run in an image that has no TensorFlow, so that the fp64 restatement in
`oracle/mmae_oracle.py` can be pinned against the reference's wiring of the graph: which tensor
feeds which op, where the transposes / reversals / activations / dropouts sit, which variables
each optimizer touches, what each `session.run` returns.

Only the ~45 API names the reference calls are provided.  Per-op semantics are the published
TF-1.x ones (listed in oracle/mmae_oracle.py's header); they are restated here op by op, not
verified against a TensorFlow binary -- the graph *wiring* is the reference's, the op kernels are
torch's.  Evaluation is lazy: ops build a small symbolic graph, `Session.run` evaluates it with a
feed dict, `Optimizer.minimize` differentiates with torch.autograd and applies TF's ApplyAdam.

Arithmetic dtype: `set_default_dtype(torch.float64)` (the default, for pinning the fp64 oracle) or
torch.float32 (what TF computes in; used when this serves as the CPU baseline).

Hooks for parity tests ("identical inputs, weights and masks"):
  * `hooks.random_normal(shape, name)`    -> array or None   (inject the VAE epsilon)
  * `hooks.dropout_uniform(shape, index)` -> array or None   (inject U[0,1) of the index-th dropout
                                                             op evaluated in this run)
  * `Variable.load(value)`                                    (inject weights)
"""
from __future__ import annotations

import contextlib
import math

import numpy as np
import torch

float32 = 'float32'
float64 = 'float64'
int32 = 'int32'
int64 = 'int64'

_DTYPE = torch.float64
_rng = np.random.RandomState(0)


def set_default_dtype(dt):
    global _DTYPE
    _DTYPE = dt


def set_random_seed(seed):
    global _rng
    _rng = np.random.RandomState(seed)


class _Hooks:
    random_normal = None
    dropout_uniform = None


hooks = _Hooks()


# ----------------------------------------------------------------------------- graph nodes
class _Ctx:
    """One Session.run: feeds, memo table, and the order counter of the random ops."""

    def __init__(self, feeds, var_values=None):
        self.feeds = feeds
        self.cache = {}
        self.var_values = var_values       # Variable -> leaf tensor (inside minimize) or None
        self.rand = {}                     # id(random op) -> its draw: every fetch of one run sees the same randomness


def _to_t(v):
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, (np.ndarray, list, tuple, float, int, np.floating, np.integer)):
        a = np.asarray(v)
        if a.dtype.kind in 'iub':
            return torch.as_tensor(a)
        return torch.as_tensor(a, dtype=_DTYPE)
    raise TypeError('cannot convert %r' % (type(v),))


class Tensor:
    def __init__(self, fn, inputs=(), name=None):
        self._fn = fn
        self._inputs = tuple(inputs)
        self.name = name

    def _eval(self, ctx):
        k = id(self)
        if k in ctx.cache:
            return ctx.cache[k]
        args = [_ev(i, ctx) for i in self._inputs]
        v = self._fn(ctx, *args)
        ctx.cache[k] = v
        return v

    # operators (tf.Tensor overloads)
    def __add__(self, o): return _op(lambda a, b: a + b, self, o)
    def __radd__(self, o): return _op(lambda a, b: b + a, self, o)
    def __sub__(self, o): return _op(lambda a, b: a - b, self, o)
    def __rsub__(self, o): return _op(lambda a, b: b - a, self, o)
    def __mul__(self, o): return _op(lambda a, b: a * b, self, o)
    def __rmul__(self, o): return _op(lambda a, b: b * a, self, o)
    def __truediv__(self, o): return _op(lambda a, b: a / b, self, o)
    def __rtruediv__(self, o): return _op(lambda a, b: b / a, self, o)
    __div__ = __truediv__
    def __neg__(self): return _op(lambda a: -a, self)
    def __pow__(self, o): return _op(lambda a, b: a ** b, self, o)
    __hash__ = object.__hash__


def _ev(x, ctx):
    if isinstance(x, Tensor):
        return x._eval(ctx)
    if isinstance(x, (list, tuple)) and any(isinstance(i, Tensor) for i in x):
        return type(x)(_ev(i, ctx) for i in x)
    return x


def _op(fn, *inputs, name=None):
    def run(ctx, *args):
        return fn(*[_to_t(a) if isinstance(a, np.ndarray) else a for a in args])
    return Tensor(run, inputs, name)


class _Placeholder(Tensor):
    def __init__(self, dtype, name=None):
        super().__init__(None, (), name)
        self.dtype = dtype

    def _eval(self, ctx):
        if self not in ctx.feeds:
            raise ValueError('placeholder %r was not fed' % (self.name,))
        k = id(self)
        if k not in ctx.cache:
            v = ctx.feeds[self]
            a = np.asarray(v)
            # placeholders are tf.float32 in the reference (:351-352, :424): feeds are cast before any math
            ctx.cache[k] = torch.as_tensor(a.astype(np.float32) if _DTYPE == torch.float32 else a.astype(np.float64), dtype=_DTYPE)
        return ctx.cache[k]


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, name)


_ALL_VARIABLES = []          # per default graph, see Graph


class Variable(Tensor):
    def __init__(self, initial_value, name=None, trainable=True):
        super().__init__(None, (), name)
        self._initial = initial_value
        self.value = None
        self.trainable = trainable
        _current_graph().variables.append(self)

    def _eval(self, ctx):
        if ctx.var_values is not None and self in ctx.var_values:
            return ctx.var_values[self]
        if self.value is None:
            raise RuntimeError('variable %r used before initialisation' % (self.name,))
        return self.value

    def initialize(self):
        v = self._initial
        if isinstance(v, Tensor):
            v = v._eval(_Ctx({}))
        self.value = _to_t(v).clone()

    def load(self, value, session=None):
        t = _to_t(np.asarray(value))
        if self.value is not None and tuple(t.shape) != tuple(self.value.shape):
            raise ValueError('shape mismatch loading %r: %r vs %r' % (self.name, tuple(t.shape), tuple(self.value.shape)))
        self.value = t.clone() if t.dtype.is_floating_point is False else t.to(_DTYPE).clone()

    def numpy(self):
        return self.value.detach().cpu().numpy()

    @property
    def is_float(self):
        return self.value is not None and self.value.dtype.is_floating_point


class Graph:
    def __init__(self):
        self.variables = []

    @contextlib.contextmanager
    def as_default(self):
        _graph_stack.append(self)
        try:
            yield self
        finally:
            _graph_stack.pop()


_root_graph = Graph()
_graph_stack = [_root_graph]


def _current_graph():
    return _graph_stack[-1]


@contextlib.contextmanager
def name_scope(name):
    yield name


class _Init(Tensor):
    def __init__(self, graph):
        super().__init__(None, (), 'init')
        self._graph = graph

    def _eval(self, ctx):
        for v in self._graph.variables:
            v.initialize()
        return None


def global_variables_initializer():
    return _Init(_current_graph())


initialize_all_variables = global_variables_initializer


# ----------------------------------------------------------------------------- ops
def constant(value, shape=None, dtype=None, name=None):
    def run(ctx):
        a = np.asarray(value, dtype=np.float64 if dtype in (None, float32, float64) else np.int64)
        if shape is not None:
            a = np.broadcast_to(a, tuple(shape)).copy()
        return _to_t(a)
    return Tensor(run, (), name)


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=float32, seed=None, name=None):
    def run(ctx, shp):
        return _to_t(_rng.uniform(minval, maxval, size=tuple(int(s) for s in shp)))
    return Tensor(run, (shape,), name)


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    """Values beyond two standard deviations are re-drawn."""
    def run(ctx, shp):
        shp = tuple(int(s) for s in shp)
        w = _rng.standard_normal(shp)
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = _rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        return _to_t(mean + stddev * w)
    return Tensor(run, (shape,), name)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    node = None

    def run(ctx, shp):
        shp = tuple(int(s) for s in shp)
        if id(node) not in ctx.rand:
            inj = hooks.random_normal(shp, name) if hooks.random_normal is not None else None
            z = np.asarray(inj, np.float64) if inj is not None else _rng.standard_normal(shp)
            if z.shape != shp:
                raise ValueError('injected normal has shape %r, op wants %r' % (z.shape, shp))
            ctx.rand[id(node)] = z
        return _to_t(mean + stddev * ctx.rand[id(node)])
    node = Tensor(run, (shape,), name)
    return node


def shape(x, name=None):
    return Tensor(lambda ctx, a: tuple(a.shape), (x,), name)


def matmul(a, b, name=None):
    return _op(lambda x, y: x @ y, a, b, name=name)


def transpose(a, perm=None, name=None):
    return _op(lambda x: x.t() if perm is None else x.permute(*perm), a, name=name)


def exp(x, name=None): return _op(torch.exp, x, name=name)
def log(x, name=None): return _op(torch.log, x, name=name)
def sqrt(x, name=None): return _op(torch.sqrt, x, name=name)
def square(x, name=None): return _op(lambda a: a * a, x, name=name)
def pow(x, y, name=None): return _op(lambda a, b: a ** b, x, y, name=name)
def round(x, name=None): return _op(torch.round, x, name=name)          # half to even, as tf.round
def equal(a, b, name=None): return _op(lambda x, y: x == y, a, b, name=name)
def sigmoid(x, name=None): return _op(torch.sigmoid, x, name=name)


def _axis(kw, axis):
    for k in ('reduction_indices', 'axis'):
        if kw.get(k) is not None:
            return kw[k]
    return axis


def reduce_sum(x, axis=None, keep_dims=False, name=None, **kw):
    ax = _axis(kw, axis)
    return _op(lambda a: a.sum() if ax is None else a.sum(dim=ax, keepdim=keep_dims), x, name=name)


def reduce_mean(x, axis=None, keep_dims=False, name=None, **kw):
    ax = _axis(kw, axis)
    return _op(lambda a: a.mean() if ax is None else a.mean(dim=ax, keepdim=keep_dims), x, name=name)


def argmax(x, axis=None, name=None, dimension=None):
    ax = axis if axis is not None else dimension
    return _op(lambda a: torch.argmax(a, dim=ax), x, name=name)


def cast(x, dtype, name=None):
    def f(a):
        if not isinstance(a, torch.Tensor):
            a = _to_t(a)
        if dtype == int32:
            return a.to(torch.int32)          # float -> int truncates toward zero; bool -> 0/1
        if dtype == int64:
            return a.to(torch.int64)
        return a.to(_DTYPE)
    return _op(f, x, name=name)


class _NN:
    @staticmethod
    def relu(x, name=None): return _op(torch.relu, x, name=name)

    @staticmethod
    def tanh(x, name=None): return _op(torch.tanh, x, name=name)

    @staticmethod
    def softsign(x, name=None): return _op(lambda a: a / (1 + a.abs()), x, name=name)

    @staticmethod
    def softplus(x, name=None): return _op(torch.nn.functional.softplus, x, name=name)

    @staticmethod
    def sigmoid(x, name=None): return _op(torch.sigmoid, x, name=name)

    @staticmethod
    def l2_loss(t, name=None): return _op(lambda a: (a * a).sum() / 2, t, name=name)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        if _sentinel is not None:
            raise ValueError('call with named arguments (labels=..., logits=...)')
        # max(x, 0) - x * z + log(1 + exp(-abs(x)))
        return _op(lambda x, z: torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs())), logits, labels, name=name)

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        if _sentinel is not None:
            raise ValueError('call with named arguments (labels=..., logits=...)')
        def f(x, y):
            lse = torch.logsumexp(x, dim=-1)
            return lse - x.gather(-1, y.to(torch.int64).unsqueeze(-1)).squeeze(-1)
        return _op(f, logits, labels, name=name)

    @staticmethod
    def dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
        """TF 1.x: random_tensor = keep_prob + U[0,1); binary = floor(random_tensor); x / keep_prob * binary."""
        node = None

        def run(ctx, a, kp):
            shp = tuple(a.shape)
            if id(node) not in ctx.rand:
                idx = len([k for k in ctx.rand if isinstance(k, tuple)])
                inj = hooks.dropout_uniform(shp, idx) if hooks.dropout_uniform is not None else None
                ctx.rand[id(node)] = np.asarray(inj, np.float64) if inj is not None else _rng.uniform(0.0, 1.0, size=shp)
                ctx.rand[('dropout', idx)] = id(node)
            u = ctx.rand[id(node)]
            kp = kp if isinstance(kp, torch.Tensor) else _to_t(float(kp))
            binary = torch.floor(kp + _to_t(u))
            return a / kp * binary
        node = Tensor(run, (x, keep_prob), name)
        return node


nn = _NN()


# ----------------------------------------------------------------------------- training
class _SideEffectOp(Tensor):
    """Ops that assign variables: Session.run evaluates them after every plain fetch of the same run."""


class _MinimizeOp(_SideEffectOp):
    def __init__(self, opt, loss, var_list):
        super().__init__(None, (), 'minimize')
        self.opt, self.loss, self.var_list = opt, loss, var_list

    def _eval(self, ctx):
        k = id(self)
        if k not in ctx.cache:
            ctx.cache[k] = self.opt._apply(self.loss, self.var_list, ctx)
        return None


class _ApplyGradientsOp(_SideEffectOp):
    def __init__(self, opt, grads_and_vars, global_step):
        super().__init__(None, (), 'apply_gradients')
        self.opt, self.gv, self.global_step = opt, list(grads_and_vars), global_step

    def _eval(self, ctx):
        k = id(self)
        if k not in ctx.cache:
            grads = [g._eval(ctx) if isinstance(g, Tensor) else g for g, _ in self.gv]
            self.opt._adam([v for _, v in self.gv], grads, ctx)
            if self.global_step is not None:
                self.global_step.value = self.global_step.value + 1
            ctx.cache[k] = True
        return None


class _GradBundle(Tensor):
    """d loss / d params for a list of variables: one autograd pass shared by the per-variable gradient nodes."""

    def __init__(self, loss, params):
        super().__init__(None, (), 'gradients')
        self.loss, self.params = loss, list(params)

    def _eval(self, ctx):
        k = id(self)
        if k not in ctx.cache:
            leaves = {v: v.value.detach().clone().requires_grad_(True) for v in self.params}
            c2 = _Ctx(ctx.feeds, leaves)
            c2.rand = ctx.rand
            lv = self.loss._eval(c2)
            gs = torch.autograd.grad(lv, [leaves[v] for v in self.params], allow_unused=True)
            ctx.cache[k] = [None if g is None else g.detach() for g in gs]
        return ctx.cache[k]


def gradients(ys, xs, name=None):
    b = _GradBundle(ys, xs)
    return [Tensor(lambda ctx, bundle, i=i: bundle[i], (b,)) for i in range(len(b.params))]


def clip_by_global_norm(t_list, clip_norm, name=None):
    """t_list[i] * clip_norm / max(global_norm, clip_norm), global_norm = sqrt(sum ||t||^2)."""
    def norm_fn(ctx, *ts):
        return torch.sqrt(sum((t * t).sum() for t in ts if t is not None))
    norm = Tensor(norm_fn, tuple(t_list), 'global_norm')
    outs = [Tensor(lambda ctx, t, n: None if t is None else t * (clip_norm / torch.clamp(n, min=float(clip_norm))), (t, norm)) for t in t_list]
    return outs, norm


def trainable_variables():
    return [v for v in _current_graph().variables if v.trainable]


class _Train:
    @staticmethod
    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
        def f(step):
            p = float(step) / float(decay_steps)
            if staircase:
                p = math.floor(p)
            return learning_rate * decay_rate ** p
        return _op(f, global_step, name=name)

    class AdamOptimizer:
        def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False, name='Adam'):
            self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
            self.m, self.v = {}, {}
            self.b1_power, self.b2_power = beta1, beta2      # TF keeps beta^t as non-trainable variables, start = beta
            self.last_grads = {}                             # name -> numpy, for the parity tests
            self.graph = _current_graph()

        def minimize(self, loss, global_step=None, var_list=None, name=None):
            if var_list is None:
                var_list = list(self.graph.variables)       # default: every trainable variable of the graph
            return _MinimizeOp(self, loss, var_list)

        def apply_gradients(self, grads_and_vars, global_step=None, name=None):
            return _ApplyGradientsOp(self, grads_and_vars, global_step)

        def _adam(self, variables, grads, ctx):
            lr = self.lr._eval(_Ctx(ctx.feeds)) if isinstance(self.lr, Tensor) else self.lr
            alpha = float(lr) * math.sqrt(1.0 - self.b2_power) / (1.0 - self.b1_power)
            self.last_grads = {}
            with torch.no_grad():
                for v, g in zip(variables, grads):
                    if g is None:
                        continue
                    self.last_grads[v.name] = g.detach().cpu().numpy().copy()
                    m = self.m.setdefault(v, torch.zeros_like(v.value))
                    s = self.v.setdefault(v, torch.zeros_like(v.value))
                    m += (g - m) * (1.0 - self.b1)
                    s += (g * g - s) * (1.0 - self.b2)
                    v.value = v.value - alpha * m / (torch.sqrt(s) + self.eps)
            self.b1_power *= self.b1
            self.b2_power *= self.b2

        def _apply(self, loss, var_list, ctx):
            cand = [v for v in var_list if v.trainable and v.value is not None and v.value.dtype.is_floating_point]
            leaves = {v: v.value.detach().clone().requires_grad_(True) for v in cand}
            # same feeds and -- crucially -- the same random draws as every other fetch of this run
            c2 = _Ctx(ctx.feeds, leaves)
            c2.rand = ctx.rand
            lv = self.loss_value = loss._eval(c2)
            grads = torch.autograd.grad(lv, [leaves[v] for v in cand], allow_unused=True)
            lr = self.lr._eval(_Ctx(ctx.feeds)) if isinstance(self.lr, Tensor) else self.lr
            lr = float(lr)
            alpha = lr * math.sqrt(1.0 - self.b2_power) / (1.0 - self.b1_power)
            self.last_grads = {}
            with torch.no_grad():
                for v, g in zip(cand, grads):
                    if g is None:                            # compute_gradients: variables the loss does not reach are skipped
                        continue
                    self.last_grads[v.name] = g.detach().cpu().numpy().copy()
                    m = self.m.setdefault(v, torch.zeros_like(v.value))
                    s = self.v.setdefault(v, torch.zeros_like(v.value))
                    m += (g - m) * (1.0 - self.b1)
                    s += (g * g - s) * (1.0 - self.b2)
                    v.value = v.value - alpha * m / (torch.sqrt(s) + self.eps)
            self.b1_power *= self.b1
            self.b2_power *= self.b2
            return None

    class Saver:
        def __init__(self, *a, **k):
            self.graph = _current_graph()

        def save(self, sess, save_path, global_step=None):
            path = save_path + ('-%d' % global_step if global_step is not None else '')
            np.savez(path + '.shim.npz', **{v.name: v.numpy() for v in self.graph.variables if v.name})
            return path

        def restore(self, sess, save_path):
            z = np.load(save_path + '.shim.npz')
            for v in self.graph.variables:
                if v.name in z.files:
                    v.load(z[v.name])

    @staticmethod
    def get_checkpoint_state(directory):
        return None

    @staticmethod
    def latest_checkpoint(directory):
        return None


train = _Train()


class Session:
    def __init__(self, graph=None, config=None):
        self.graph = graph or _current_graph()

    def run(self, fetches, feed_dict=None):
        ctx = _Ctx(dict(feed_dict or {}))
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        # ops with side effects (minimize) see the pre-update variables, like every other fetch of the run
        plain = [f for f in fl if not isinstance(f, _SideEffectOp)]
        vals = {}
        for f in plain:
            vals[id(f)] = _out(f._eval(ctx)) if isinstance(f, Tensor) else f
        for f in fl:
            if isinstance(f, _SideEffectOp):
                f._eval(ctx)
                vals[id(f)] = None
        out = [vals[id(f)] for f in fl]
        return out[0] if single else out

    def close(self):
        pass


def _out(v):
    if isinstance(v, torch.Tensor):
        a = v.detach().cpu().numpy()
        if a.dtype == np.float64 and _DTYPE == torch.float32:
            a = a.astype(np.float32)
        return a[()] if a.ndim == 0 else a
    return v
