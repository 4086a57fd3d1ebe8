"""ctypes binding of libd2p.so (the C ABI declared in include/d2p.h).

The product path has no CPU fallback: if the CUDA library is missing or fails
to load, every op raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libd2p.so')

D2P_U8, D2P_F32 = 0, 1
MAX_CONV_LAYERS = 5

_fp = C.c_void_p  # device pointers travel as integers


class ConvLayer(C.Structure):
    _fields_ = [('w', _fp), ('b', _fp), ('gamma', _fp), ('beta', _fp),
                ('moving_mean', _fp), ('moving_var', _fp),
                ('dw', _fp), ('db', _fp), ('dgamma', _fp), ('dbeta', _fp),
                ('cout', C.c_int), ('_pad', C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [('B', C.c_int), ('k', C.c_int), ('T', C.c_int), ('h', C.c_int),
                ('w', C.c_int), ('d', C.c_int), ('frames_dtype', C.c_int),
                ('n_layers', C.c_int), ('layers', ConvLayer * MAX_CONV_LAYERS)]


class FcBn(C.Structure):
    _fields_ = [('w', _fp), ('b', _fp), ('gamma', _fp), ('beta', _fp),
                ('moving_mean', _fp), ('moving_var', _fp),
                ('dw', _fp), ('db', _fp), ('dgamma', _fp), ('dbeta', _fp)]


class D2PError(RuntimeError):
    pass


_i, _f, _sz, _ll = C.c_int, C.c_float, C.c_size_t, C.c_longlong
_pd, _pf = C.POINTER(ConvDesc), C.POINTER(FcBn)

# name -> (restype, argtypes); mirrors include/d2p.h one to one
SIGNATURES = {
    'd2p_last_error': (C.c_char_p, []),
    'd2p_version': (_i, []),
    'd2p_launch_count': (_ll, []),
    'd2p_conv_encoder_saved_floats': (_sz, [_pd]),
    'd2p_conv_encoder_ws_bytes': (_sz, [_pd]),
    'd2p_conv_encoder_feature_dim': (_i, [_pd]),
    'd2p_conv_encoder_fwd': (_i, [_pd, _fp, _fp, _fp, _i, _fp, _sz, _fp]),
    'd2p_conv_encoder_bwd': (_i, [_pd, _fp, _fp, _fp, _i, _fp, _sz, _fp]),
    'd2p_lstm_seq_fwd': (_i, [_fp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _f,
                              _fp, _fp, _fp, _fp, _fp, _i, _fp]),
    'd2p_lstm_seq_bwd_ws_bytes': (_sz, [_i, _i, _i]),
    'd2p_lstm_seq_bwd': (_i, [_fp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp,
                              _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                              _sz, _i, _fp]),
    'd2p_embed_shifted': (_i, [_fp, _i, _i, _fp, _i, _i, _i, _fp, _fp]),
    'd2p_step_lens': (_i, [_fp, _i, _i, _fp, _fp]),
    'd2p_embed_shifted_step': (_i, [_fp, _i, _i, _fp, _i, _i, _i, _i, _fp, _fp]),
    'd2p_sched_sample_step': (_i, [_fp, _i, _i, _fp, _i, _i, _fp, _i, _f, C.c_uint, _i, _fp, _fp, _fp]),
    'd2p_sched_hash': (C.c_uint, [C.c_uint] * 6),
    'd2p_embed_shifted_bwd_ws_bytes': (_sz, [_i, _i, _i, _i]),
    'd2p_embed_shifted_bwd': (_i, [_fp, _i, _i, _fp, _i, _i, _i, _fp, _fp, _sz, _fp]),
    'd2p_seq_weights': (_i, [_fp, _i, _i, _f, _i, _fp, _fp, _fp]),
    'd2p_softmax_ce': (_i, [_fp, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                            _i, _fp]),
    'd2p_sigmoid_ce': (_i, [_fp, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                            _i, _fp]),
    'd2p_greedy_ws_bytes': (_sz, [_i, _i, _i]),
    'd2p_lstm_decoder_greedy': (_i, [_fp, _i, _i, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _fp, _fp,
                                     _fp, _fp, _fp, _fp, _sz, _fp]),
    'd2p_greedy_near_ties': (_i, [_fp, _i, _i, _i, _fp, _f, _fp, _fp]),
    'd2p_luong_pool_attention_ws_bytes': (_sz, [_i, _i, _i, _i]),
    'd2p_luong_pool_attention': (_i, [_fp, _i, _fp, _fp, _fp, _i, _i, _i, _i, _i, _fp, _i, _fp, _sz, _fp]),
    'd2p_induction_decode_ws_bytes': (_sz, [_i, _i, _i, _i]),
    'd2p_induction_decode': (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _fp, _fp, _fp, _i, _fp, _fp, _fp,
                                  _fp, _fp, _i, _fp, _fp, _fp, _fp, _sz, _fp]),
    'd2p_concat_cols': (_i, [_fp, _i, _fp, _i, _ll, _fp, _fp]),
    'd2p_luong_pool_attention_train_ws_bytes': (_sz, [_i, _i, _i, _i]),
    'd2p_luong_pool_attention_train_fwd': (_i, [_fp, _i, _fp, _fp, _fp, _i, _i, _i, _i, _i, _fp, _i, _fp, _fp, _sz,
                                                _fp]),
    'd2p_luong_pool_attention_train_bwd': (_i, [_fp, _i, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _fp, _i,
                                                _fp, _fp, _fp, _sz, _fp]),
    'd2p_split_cols': (_i, [_fp, _i, _i, _i, _ll, _fp, _fp]),
    'd2p_karel_check_syntax': (_i, [_fp, _i]),
    'd2p_karel_execute': (_i, [_fp, _i, _fp, _i, _i, _i, _i, _fp, _fp]),
    'd2p_karel_programs_equal': (_i, [_fp, _i, _fp, _i]),
    'd2p_karel_eval_batch': (_i, [_fp, _fp, _fp, _i, _i, _fp, _fp, _i, _i, _i, _i, _i, _fp, _fp, _fp, _i]),
    'd2p_fc_bn_saved_floats': (_sz, [_ll, _i, _i]),
    'd2p_fc_bn_ws_bytes': (_sz, [_ll, _i, _i]),
    'd2p_fc_bn_fwd': (_i, [_fp, _ll, _i, _i, _i, _i, _i, _pf, _fp, _fp, _i, _fp,
                           _sz, _fp]),
    'd2p_fc_bn_bwd': (_i, [_fp, _ll, _i, _i, _i, _i, _i, _pf, _fp, _fp, _fp, _i,
                           _fp, _sz, _fp]),
    'd2p_rn_pool_saved_floats': (_sz, [_i, _i, _i]),
    'd2p_rn_pool_ws_bytes': (_sz, [_i, _i, _i]),
    'd2p_rn_pool_fwd': (_i, [_fp, _i, _i, _i, _pf, _pf, _fp, _fp, _i, _fp, _sz,
                             _fp]),
    'd2p_rn_pool_bwd': (_i, [_fp, _i, _i, _i, _pf, _pf, _fp, _fp, _fp, _i, _fp,
                             _sz, _i, _fp, _fp]),
    'd2p_rn_pool_bwd_hold_floats': (_sz, [_i, _i, _i]),
    'd2p_group_sum': (_i, [_fp, _i, _i, _i, _f, _fp, _i, _fp]),
    'd2p_group_bcast': (_i, [_fp, _i, _i, _i, _f, _fp, _i, _fp]),
    'd2p_group_max': (_i, [_fp, _i, _i, _i, _fp, _fp, _fp]),
    'd2p_group_max_bwd': (_i, [_fp, _fp, _i, _i, _i, _fp, _fp]),
    'd2p_axpby': (_i, [_fp, _f, _fp, _f, _sz, _fp]),
    'd2p_add3': (_i, [_fp, _fp, _fp, _fp, _sz, _fp]),
    'd2p_logits_to_bvl': (_i, [_fp, _i, _i, _i, _fp, _fp]),
    'd2p_rtp_to_trp': (_i, [_fp, _i, _i, _i, _fp, _fp]),
    'd2p_len_to_int': (_i, [_fp, _fp, _i, _fp]),
    'd2p_adam_ws_bytes': (_sz, []),
    'd2p_clip_adam_step': (_i, [_fp, _fp, _fp, _fp, _sz, _f, _f, _f, _f, _f, _f,
                                _i, _fp, _fp, _sz, _fp]),
    'd2p_gemm_tc_ws_bytes': (_sz, [_i, _i, _i]),
    'd2p_gemm_tc': (_i, [_i, _i, _i, _i, _i, _f, _fp, _i, _fp, _i, _f, _fp, _i, _fp,
                         _fp, _sz, _fp]),
    'd2p_packed_bytes': (_sz, [_i, _i]),
    'd2p_pack_bf16': (_i, [_fp, _i, _i, _i, _i, _fp, _fp]),
    'd2p_gemm_tc_packed': (_i, [_fp, _fp, _i, _i, _i, _f, _f, _fp, _i, _fp, _i, _fp, _fp]),
    'd2p_tc_configure': (_i, [_fp, _sz, _fp, _sz, _i]),
    'd2p_tc_new_step': (_i, []),
    'd2p_tc_bind_stream': (_i, [_fp, _fp, _sz]),
    'd2p_debug_set_probe': (_i, [_fp]),
    'd2p_lstm_set_persistent': (_i, [_i]),
    'd2p_conv_set_fused': (_i, [_i]),
    'd2p_conv_set_tc': (_i, [_i]),
    'd2p_device_error': (_i, [_fp]),
    'd2p_device_error_async': (_i, [_fp, _fp]),
    'd2p_debug_inject_device_error': (_i, [_i]),
    'd2p_debug_stamp': (_i, [_fp, _i, _fp]),
    'd2p_gemm_set_persistent': (_i, [_i]),
    'd2p_crc32c': (C.c_uint32, [_fp, _sz, C.c_uint32]),
    'd2p_gemm': (_i, [_i, _i, _i, _i, _i, _f, _fp, _i, _fp, _i, _f, _fp, _i, _fp,
                      _fp]),
}

_lib = None


def load():
    """Load libd2p.so (once).  Raises D2PError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D2PError(
            'libd2p.so not found at %s - build it with '
            '`python -m demo2program_b200.build` (nvcc, sm_100a). There is no '
            'CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().d2p_last_error()
        raise D2PError('%s failed (%d): %s' % (what, rc, msg.decode() if msg else ''))


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


launch_count = 0   # incremented by the op layer: number of C-ABI op calls issued
