"""ctypes binding of libmmae_b200.so (include/mmae_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module
without the built .so raises ImportError with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmmae_b200.so')

# enums (mmae_b200.h)
ACT = {'linear': 0, 'relu': 1, 'tanh': 2, 'softsign': 3, 'softplus': 4}
LOSS = {'mean_squared': 0, 'sigmoid_cross_entropy': 1, 'cross_entropy': 2}
HEAD_LOSS = {'sigmoid_cross_entropy': 0, 'softmax': 1}
PREC = {'fp32': 0, 'tf32': 1}
NOISE_INTELLIGENT, NOISE_UNIFORM = 0, 1
WANT_RECON, WANT_EMBEDDING, WANT_HEAD, WANT_LOSS, WANT_FILLED, WANT_HEAD_LOSS = 1, 2, 4, 8, 16, 32
S_RECON_LOSS, S_KL_MEAN, S_SUMSQ, S_HEAD_LOSS, S_HEAD_ACC, S_GRAD_SCALE = 0, 1, 2, 3, 4, 5
NUM_SCALARS = 8


class Config(C.Structure):
    _fields_ = [
        ('num_feats', C.c_int32), ('num_modalities', C.c_int32), ('modality_starts', C.POINTER(C.c_int32)),
        ('num_layers', C.c_int32), ('layer_sizes', C.POINTER(C.c_int32)),
        ('tie_weights', C.c_int32), ('variational', C.c_int32), ('activation', C.c_int32), ('loss_func', C.c_int32),
        ('weight_penalty', C.c_float), ('learning_rate', C.c_float),
        ('beta1', C.c_float), ('beta2', C.c_float), ('adam_eps', C.c_float),
        ('num_head_layers', C.c_int32), ('head_sizes', C.POINTER(C.c_int32)),
        ('head_activation', C.c_int32), ('head_loss', C.c_int32),
        ('head_weight_penalty', C.c_float), ('head_learning_rate', C.c_float),
        ('mask_with', C.c_float), ('n_zero', C.c_int32), ('noise_mode', C.c_int32),
        ('num_noise_types', C.c_int32), ('noise_type_masks', C.POINTER(C.c_uint32)),
        ('noise_thresholds', C.POINTER(C.c_uint32)), ('num_modalities_to_drop', C.c_int32),
        ('seed', C.c_uint64), ('precision', C.c_int32), ('max_batch', C.c_int64),
        ('classifier_only', C.c_int32), ('clip_norm', C.c_float),
    ]


class Outputs(C.Structure):
    _fields_ = [('recon', C.c_void_p), ('embedding', C.c_void_p), ('logits', C.c_void_p),
                ('probs', C.c_void_p), ('preds', C.c_void_p), ('filled', C.c_void_p)]


# name -> (restype, argtypes); every symbol include/mmae_b200.h declares
_P, _I, _L, _F, _U = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32
PROTOTYPES = {
    'mmae_create': (_I, [C.POINTER(Config), C.POINTER(_P)]),
    'mmae_destroy': (None, [_P]),
    'mmae_last_error': (C.c_char_p, [_P]),
    'mmae_set_stream': (_I, [_P, _P]),
    'mmae_synchronize': (_I, [_P]),
    'mmae_num_variables': (_I, [_P]),
    'mmae_variable_info': (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_L), C.POINTER(_L)]),
    'mmae_set_variable': (_I, [_P, C.c_char_p, _P, _L]),
    'mmae_get_variable': (_I, [_P, C.c_char_p, _P, _L]),
    'mmae_get_gradient': (_I, [_P, C.c_char_p, _P, _L]),
    'mmae_get_opt_state': (_I, [_P, _I, C.c_char_p, _P, _P, _L, C.POINTER(_L)]),
    'mmae_set_opt_state': (_I, [_P, _I, C.c_char_p, _P, _P, _L, _L]),
    'mmae_set_rng_step': (_I, [_P, C.c_uint64]),
    'mmae_set_noise': (_I, [_P, _P, _P, _L]),
    'mmae_gen_noise': (_I, [_P, _L, _L]),
    'mmae_get_noise': (_I, [_P, _P, _P, _L]),
    'mmae_apply_noise': (_I, [_P, _P, _L, _P]),
    'mmae_forward': (_I, [_P, _P, _P, _P, _L, _I, _F, _U, C.POINTER(Outputs)]),
    'mmae_train_step': (_I, [_P, _P, _L, _I, _F]),
    'mmae_train_step_pair': (_I, [_P, _P, _P, _L, _I, _F]),
    'mmae_cls_train_step': (_I, [_P, _P, _P, _L, _I, _F]),
    'mmae_train_step_host': (_I, [_P, _P, _L, _I, _F]),
    'mmae_cls_train_step_host': (_I, [_P, _P, _P, _L, _I, _F]),
    'mmae_forward_host': (_I, [_P, _P, _P, _P, _L, _I, _F, _U, C.POINTER(Outputs)]),
    'mmae_backward': (_I, [_P, _P, _L, _L, _I, _F]),
    'mmae_grad_buffer': (_I, [_P, C.POINTER(_P), C.POINTER(_L)]),
    'mmae_apply_update': (_I, [_P, _I]),
    'mmae_set_dataset': (_I, [_P, _I, _P, _P, _L, C.c_int32]),
    'mmae_set_dataset_device': (_I, [_P, _I, _P, _P, _L, C.c_int32]),
    'mmae_set_dataset_view': (_I, [_P, _I, _P, _L]),
    'mmae_train_step_resident': (_I, [_P, _I, _P, _L, _I, _F, _I]),
    'mmae_eval_resident': (_I, [_P, _I, _L, _I, _F]),
    'mmae_modality_rmse': (_I, [_P, _P, _L, C.POINTER(C.c_double)]),
    'mmae_read_scalars': (_I, [_P, C.POINTER(C.c_double), _I]),
    'mmae_comm_unique_id': (_I, [_P]),
    'mmae_comm_init': (_I, [_P, _P, _I, _I]),
    'mmae_set_shard': (_I, [_P, _L, _L]),
    'mmae_kernel_launches': (_L, [_P]),
    'mmae_chain_launches': (_L, [_P]),
    'mmae_backward_chain_launches': (_L, [_P]),
    'mmae_wgrad_group_launches': (_L, [_P]),
    'mmae_graph_replays': (_L, [_P]),
    'mmae_fused_noise_launches': (_L, [_P]),
    'mmae_set_profiling': (_I, [_P, _I]),
    'mmae_read_profile': (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_L)]),
    'mmae_read_scalars_async': (_I, [_P, _P, _I]),
    'mmae_get_buffer': (_I, [_P, C.c_char_p, _P, _L]),
    'mmae_set_eps': (_I, [_P, _P, _L]),
    'mmae_debug_gemm': (_I, [_I, _I, _I, _L, _L, _L, _P, _L, _P, _L, _P, _L, _P, _I, _F, _P]),
}

_lib = None


def load():
    """dlopen the engine.  Raises ImportError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libmmae_b200.so is not built (%s). Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C multimodalautoencoder_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
