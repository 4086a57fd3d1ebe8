"""Host-side Karel DSL services of the evaluation path, served by the native library
(csrc/karel_dsl.cu): syntax check, program execution, exact-program comparison and the batch
execution-accuracy metrics the reference computes with per-step Python `py_func`s
(models/model_full.py:602-616, 712-727, 747-787, 870-897)."""
import ctypes as C

import numpy as np

from . import _lib


def _ip(a):
    return a.ctypes.data_as(C.c_void_p)


def _ids(tokens):
    return np.ascontiguousarray(np.asarray(tokens, dtype=np.int32).reshape(-1))


def check_syntax(tokens):
    """True when the token-id sequence parses (karel_env/dsl/dsl_parse.py:252-265)."""
    t = _ids(tokens)
    return _lib.load().d2p_karel_check_syntax(_ip(t), int(t.size)) == 1


def execute(tokens, state0, make_error=True, max_states=None):
    """Run a program on the initial state [h, w, 16].  Returns (status, s_h): status 1 = ran to
    completion, 0 = failed at run time (blocked move, marker error, call budget), -1 = does
    not parse; s_h = bool array [n_states, h, w, 16] (initial state first)."""
    t = _ids(tokens)
    s0 = np.ascontiguousarray(np.asarray(state0).astype(np.uint8))
    h, w, d = s0.shape
    if d != 16:
        raise ValueError('Karel states have 16 channels')
    cap = 256 if max_states is None else int(max_states)
    buf = np.zeros((cap, h, w, 16), np.uint8)
    n = C.c_int(0)
    rc = _lib.load().d2p_karel_execute(_ip(t), int(t.size), _ip(s0), h, w, int(bool(make_error)), cap,
                                       _ip(buf), C.addressof(n))
    if rc != 1:
        return rc, np.zeros((0, h, w, 16), bool)
    return 1, buf[:min(n.value, cap)].astype(bool)


def programs_equal(a, b):
    """dsl_enum_program canonical comparison: 1 / 0, -1 when a side is not a complete program."""
    a, b = _ids(a), _ids(b)
    return _lib.load().d2p_karel_programs_equal(_ip(a), int(a.size), _ip(b), int(b.size))


def eval_batch(tokens, lens, is_same_seq, demos, demo_len, make_error=True, nthreads=0):
    """tokens [B, L] ids, lens [B], is_same_seq [B], demos [B, k, T, h, w, 16], demo_len [B, k].
    Returns is_correct_syntax [B], is_correct_execution [B, k], num_correct_execution [B]."""
    tokens = np.ascontiguousarray(np.asarray(tokens, np.int32))
    B, L = tokens.shape
    lens = np.ascontiguousarray(np.asarray(lens, np.int32).reshape(B))
    same = np.ascontiguousarray(np.asarray(is_same_seq).reshape(B).astype(np.uint8))
    demos = np.ascontiguousarray(np.asarray(demos).astype(np.uint8))
    _, k, T, h, w, d = demos.shape
    if d != 16 or demos.shape[0] != B:
        raise ValueError('demos must be [B, k, T, h, w, 16]')
    demo_len = np.ascontiguousarray(np.asarray(demo_len, np.int32).reshape(B, k))
    syn = np.zeros(B, np.float32)
    exe = np.zeros((B, k), np.float32)
    num = np.zeros(B, np.float32)
    _lib.check(_lib.load().d2p_karel_eval_batch(_ip(tokens), _ip(lens), _ip(same), B, L, _ip(demos),
                                                _ip(demo_len), k, T, h, w, int(bool(make_error)), _ip(syn),
                                                _ip(exe), _ip(num), int(nthreads)), 'd2p_karel_eval_batch')
    return syn, exe, num
