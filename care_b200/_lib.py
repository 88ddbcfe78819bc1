"""ctypes binding of libcare_b200.so (the C ABI declared in include/care_b200.h).

There is no fallback: if the shared library is missing or cannot be loaded this module raises,
and every op raises if the library reports an error (e.g. no sm_100 device).
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcare_b200.so")            # 16-bit operand type: IEEE fp16
LIB_PATH_BF16 = os.path.join(_HERE, "lib", "libcare_b200_bf16.so")  # same sources built with -DCARE_USE_BF16

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_RELU = 0, 1


class BeamState(Structure):
    """Mirror of `care_beam_state` (include/care_b200.h)."""
    _fields_ = [
        ("B", c_int32), ("K", c_int32), ("T_max", c_int32), ("V", c_int32), ("need", c_int32),
        ("scores", c_void_p), ("cur_tok", c_void_p), ("tok_hist", c_void_p), ("prev_ks", c_void_p),
        ("anc", c_void_p), ("fin_score", c_void_p), ("fin_t", c_void_p), ("fin_k", c_void_p),
        ("fin_count", c_void_p), ("done", c_void_p), ("n_done", c_void_p), ("scratch", c_void_p),
    ]


class NextStep(Structure):
    """Mirror of `care_next_step` (include/care_b200.h)."""
    _fields_ = [("word_emb", c_void_p), ("pos_emb", c_void_p), ("gsg", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("eps", c_float), ("d", c_int32), ("x0", c_void_p), ("x0_f32", c_void_p)]


_SIGNATURES = {
    "care_version": (c_int, []),
    "care_h16_dtype": (c_int, []),
    "care_last_error": (c_char_p, []),
    "care_ctx_create": (c_int, [POINTER(c_void_p), c_int]),
    "care_ctx_destroy": (None, [c_void_p]),
    "care_ctx_sm_count": (c_int, [c_void_p]),
    "care_ctx_share_tuning": (c_int, [c_void_p, c_void_p]),
    "care_ctx_set_next_step": (c_int, [c_void_p, POINTER(NextStep)]),
    "care_ctx_request_records": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "care_ctx_launch_count": (c_int64, [c_void_p]),
    "care_ctx_last_kernel": (c_char_p, [c_void_p, c_char_p]),
    "care_ctx_set_early_exit": (c_int, [c_void_p, c_void_p, c_int]),
    "care_ctx_counter": (c_int, [c_void_p, c_char_p, POINTER(c_int64)]),
    "care_ctx_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "care_gemm": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                          c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "care_split_f32_h16": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "care_encoder_ln_mean": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int,
                                     c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "care_encoder_highway_bn_mean": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_void_p, c_int,
                                             c_int, c_void_p, c_int64, c_int, c_void_p]),
    "care_concept_head": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_int, c_int, c_void_p]),
    "care_embed_ln": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "care_add_ln": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int,
                            c_void_p, c_void_p]),
    "care_gemm_add_ln": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "care_self_attn_step": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "care_cross_attn_step": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "care_group_attn": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "care_rows_mean": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "care_combine_means": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, POINTER(c_float), c_void_p, c_int64,
                                   c_void_p]),
    "care_nar_length_beam": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "care_nar_init": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "care_nar_best_logits": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "care_nar_teacher_probs": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p]),
    "care_nar_best_partials": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "care_nar_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_int, c_void_p]),
    "care_nar_remask": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                c_void_p]),
    "care_nar_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "care_beam_init": (c_int, [c_void_p, POINTER(BeamState), c_int, c_void_p]),
    "care_beam_step": (c_int, [c_void_p, POINTER(BeamState), c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p,
                               c_void_p]),
    "care_vocab_beam_nseg": (c_int, [c_void_p, c_int, c_int]),
    "care_vocab_beam_partials": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                         c_void_p, c_int, c_void_p]),
    "care_beam_step_partials": (c_int, [c_void_p, POINTER(BeamState), c_void_p, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p]),
    "care_beam_first_step_partials": (c_int, [c_void_p, POINTER(BeamState), c_void_p, c_int, c_int, c_void_p, c_void_p,
                                              c_void_p]),
    "care_ensemble_logprobs": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "care_beam_step_logprobs": (c_int, [c_void_p, POINTER(BeamState), c_void_p, c_int64, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p]),
    "care_beam_finalize": (c_int, [c_void_p, POINTER(BeamState), c_double, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())

_libs = {}


def load(h16="fp16"):
    """Loads the shared library whose 16-bit operand type is `h16` ("fp16": libcare_b200.so, "bf16":
    libcare_b200_bf16.so), once.  Raises if it has not been built."""
    lib = _libs.get(h16)
    if lib is not None:
        return lib
    if h16 not in ("fp16", "bf16"):
        raise ValueError("h16 must be 'fp16' or 'bf16'")
    path = LIB_PATH if h16 == "fp16" else LIB_PATH_BF16
    if not os.path.isfile(path):
        raise RuntimeError(
            "care_b200: %s is missing - build it with `python -m care_b200.build` "
            "(there is no CPU or PyTorch fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.care_h16_dtype() != (F16 if h16 == "fp16" else BF16):
        raise RuntimeError("care_b200: %s was built for another 16-bit type" % path)
    _libs[h16] = lib
    return lib


def check(rc, what="", lib=None):
    if rc != 0:
        msg = None
        for l in ([lib] if lib is not None else list(_libs.values())):
            msg = l.care_last_error()
            if msg:
                break
        raise RuntimeError("care_b200 %s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
