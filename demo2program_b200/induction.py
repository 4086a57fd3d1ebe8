"""Induction baseline (reference models/baselines/model_induction.py): encoder
(CNN features ++ perception vector -> LSTM), avg aggregate, k-way pooled Luong
attention decoder predicting the action sequences of the `test_k` unseen demos.

Inference (BASELINE.json configs[4]: greedy decode, batch 512): the teacher-forced forward pass
with its loss and the greedy decoder run in one fused device loop (d2p_induction_decode).
Training (reference trainer.py:102-109 over model_induction.py:788-819): `train_step` unrolls the
teacher-forced decoder step by step through the same C-ABI cell (d2p_lstm_seq_fwd/_bwd with T = 1,
input [embedding ; attention_{t-1}]), the training forms of the pooled attention
(d2p_luong_pool_attention_train_fwd/_bwd, csrc/attn_train.cu) and the shared clip + Adam step;
it is not a benchmarked configuration.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, check, ptr
from .manifest import build_manifests


class InductionEngine:
    def __init__(self, cfg, device='cuda:0', seed=0, is_train=False, frames_dtype=np.uint8,
                 flat_params=None, flat_state=None, use_tc=True, fold_memory_layer=False, **_):
        self.lib = _lib.load()
        # score = h . (values W_mem) = (W_mem h) . values: with the memory layer folded into the
        # query the keys tensor is never built (encode -0.3 ms at C5) but every decode step pays a
        # small product and the attention kernel is latency-bound (+0.4 ms): off by default
        self.fold_memory_layer = bool(fold_memory_layer)
        if not torch.cuda.is_available():
            raise _lib.D2PError('demo2program_b200 needs a CUDA device (no CPU fallback)')
        if cfg.model != 'induction_baseline':
            raise ValueError(cfg.model)
        for flag in ('pixel_input', 'state_encoder_fc', 'concat_state_feature_direct_prediction',
                     'stack_subsequent_state'):
            if getattr(cfg, flag):
                raise NotImplementedError('induction_baseline with %s=True' % flag)
        if cfg.attn_type != 'luong':
            raise ValueError('Unknown attention type')
        cfg.validate()
        self.cfg, self.dev = cfg, torch.device(device)
        torch.cuda.set_device(self.dev)
        self.is_train = bool(is_train)
        self.frames_u8 = np.dtype(frames_dtype) == np.uint8
        self.pm, self.sm = build_manifests(cfg)
        p0 = self.pm.init_flat(seed) if flat_params is None else np.asarray(flat_params, np.float32)
        s0 = self.sm.init_flat(seed) if flat_state is None else np.asarray(flat_state, np.float32)
        self.params = torch.from_numpy(p0.copy()).to(self.dev)
        self.state = torch.from_numpy(s0.copy()).to(self.dev)
        self.use_tc = bool(use_tc)
        self._alloc()

    def P(self, name):
        e = self.pm[name]
        return self.params[e.offset:e.offset + e.size]

    def S(self, name):
        e = self.sm[name]
        return self.state[e.offset:e.offset + e.size]

    def _alloc(self):
        cfg, lib, dev = self.cfg, self.lib, self.dev
        B, k, tk, T, H = cfg.batch_size, cfg.k, cfg.test_k, cfg.max_demo_len, cfg.num_lstm_cell_units
        A, Pd = cfg.action_space, cfg.per_dim
        R, R2 = B * k, B * tk
        self.B, self.k, self.tk, self.T, self.H, self.R, self.R2 = B, k, tk, T, H, R, R2
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        zi = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        fdt = torch.uint8 if self.frames_u8 else torch.float32
        self.d_frames = torch.zeros(B, k, T, cfg.h, cfg.w, cfg.depth, dtype=fdt, device=dev)
        self.d_per = z(R, T, Pd)
        self.d_demo_len_f, self.d_tlen_f = z(R), z(R2)
        self.d_demo_len, self.d_tlen = zi(R), zi(R2)
        self.d_ttok = zi(R2, T)
        d = ConvDesc()
        d.B, d.k, d.T, d.h, d.w, d.d = B, k, T, cfg.h, cfg.w, cfg.depth
        d.frames_dtype = _lib.D2P_U8 if self.frames_u8 else _lib.D2P_F32
        chans = cfg.conv_channels()
        d.n_layers = len(chans)
        for li, (_, cout) in enumerate(chans):
            sc = 'Demo_Encoder/State_Encoder/conv%d' % (li + 1)
            l, bn = d.layers[li], sc + '/bn_act/BatchNorm/'
            l.w, l.b = ptr(self.P(sc + '/Conv/weights')), ptr(self.P(sc + '/Conv/biases'))
            l.gamma, l.beta = ptr(self.P(bn + 'gamma')), ptr(self.P(bn + 'beta'))
            l.moving_mean, l.moving_var = ptr(self.S(bn + 'moving_mean')), ptr(self.S(bn + 'moving_variance'))
            l.cout = cout
        self.conv_desc = d
        F = lib.d2p_conv_encoder_feature_dim(C.byref(d))
        self.F = F
        self.conv_saved = z(lib.d2p_conv_encoder_saved_floats(C.byref(d)))
        ws = max(lib.d2p_conv_encoder_ws_bytes(C.byref(d)), lib.d2p_induction_decode_ws_bytes(B, k, tk, H))
        self.ws_bytes = (ws + 255) // 256 * 256
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.feat, self.per_tm, self.X = z(T, R, F), z(T, R, Pd), z(T, R, F + Pd)
        self.Y, self.keys = z(T, R, H), z(T, R, H)
        self.hT, self.cT = z(R, H), z(R, H)
        self.gates, self.cells = z(T, R, 4 * H), z(T, R, H)
        self.h_sum, self.c_sum = z(B, H), z(B, H)
        self.logits = z(T, R2, A)
        self.rowloss, self.w, self.runlen = z(T * R2), z(R2), zi(R2)
        self.loss = z(1)
        if self.use_tc:
            big = lib.d2p_gemm_tc_ws_bytes(T * R, 4 * H, max(H, F + Pd)) + (8 << 20)
            self.tc_scratch = torch.zeros(big, dtype=torch.uint8, device=dev)
            self.tc_cache = torch.zeros(12 * self.pm.total + (16 << 20), dtype=torch.uint8, device=dev)

    def _tc_bind(self, enabled=True):
        if self.use_tc and enabled:
            self.lib.d2p_tc_configure(ptr(self.tc_scratch), self.tc_scratch.numel(),
                                      ptr(self.tc_cache), self.tc_cache.numel(), 1)
        else:
            self.lib.d2p_tc_configure(None, 0, None, 0, 0)

    def _st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _call(self, name, *args):
        check(getattr(self.lib, name)(*args), name)

    def stage_batch(self, batch):
        """numpy feed (keys of model_induction.py:358-381) -> device."""
        dev = self.dev
        fr = np.asarray(batch['s_h']).astype(np.uint8 if self.frames_u8 else np.float32, copy=False)
        self.d_frames.copy_(torch.from_numpy(np.ascontiguousarray(fr)))
        self.d_per.copy_(torch.from_numpy(np.asarray(batch['per'], np.float32).reshape(self.R, self.T, -1)))
        self.d_demo_len_f.copy_(torch.from_numpy(np.asarray(batch['demo_len'], np.float32).reshape(-1)))
        self.d_tlen_f.copy_(torch.from_numpy(np.asarray(batch['test_demo_len'], np.float32).reshape(-1)))
        self.d_ttok.copy_(torch.from_numpy(
            np.asarray(batch['test_a_h_tokens'], np.int32).reshape(self.R2, self.T)))

    def encode(self, exact=False):
        """Demo_Encoder for all k demos + avg aggregate + attention keys."""
        cfg, st, call = self.cfg, self._st(), self._call
        B, k, T, H, R, F = self.B, self.k, self.T, self.H, self.R, self.F
        Pd = cfg.per_dim
        self._tc_bind(not exact)
        tr = int(self.is_train)
        call('d2p_len_to_int', ptr(self.d_demo_len_f), ptr(self.d_demo_len), R, st)
        call('d2p_len_to_int', ptr(self.d_tlen_f), ptr(self.d_tlen), self.R2, st)
        call('d2p_conv_encoder_fwd', C.byref(self.conv_desc), ptr(self.d_frames), ptr(self.feat),
             ptr(self.conv_saved), tr, ptr(self.ws), self.ws_bytes, st)
        call('d2p_rtp_to_trp', ptr(self.d_per), R, T, Pd, ptr(self.per_tm), st)
        call('d2p_concat_cols', ptr(self.feat), F, ptr(self.per_tm), Pd, T * R, ptr(self.X), st)
        sc = 'Demo_Encoder/rnn/basic_lstm_cell/'
        call('d2p_lstm_seq_fwd', ptr(self.X), T, R, F + Pd, H, ptr(self.d_demo_len), None, None,
             ptr(self.P(sc + 'kernel')), ptr(self.P(sc + 'bias')), 1.0, ptr(self.Y), ptr(self.hT),
             ptr(self.cT), ptr(self.gates), ptr(self.cells), 3, st)
        call('d2p_group_sum', ptr(self.hT), B, k, H, 1.0 / k, ptr(self.h_sum), 0, st)
        call('d2p_group_sum', ptr(self.cT), B, k, H, 1.0 / k, ptr(self.c_sum), 0, st)
        # values = Y (zero past len).  keys = values * W_mem (LuongAttention memory_layer) are only
        # materialised when the memory layer is NOT folded into the decoder's query
        if not self.fold_memory_layer:
            call('d2p_gemm', 0, 0, T * R, H, H, 1.0, ptr(self.Y), H,
                 ptr(self.P('AttnMechanism/memory_layer/kernel')), H, 0.0, ptr(self.keys), H, None, st)

    def _decode(self, tokens, out_tokens, lengths, exact):
        cfg, st = self.cfg, self._st()
        w = 'Manipulation/dynamic_decoder/pooling_attention_wrapper/'
        self._tc_bind(not exact)
        fold = self.fold_memory_layer
        self._call('d2p_induction_decode', None if fold else ptr(self.keys),
                   ptr(self.P('AttnMechanism/memory_layer/kernel')), ptr(self.Y), ptr(self.d_demo_len), self.B,
                   self.k, self.tk, self.T, self.H, ptr(self.h_sum), ptr(self.c_sum),
                   ptr(self.P('Manipulation/Token_Embedding/embedding_map')), cfg.action_space,
                   ptr(self.P(w + 'basic_lstm_cell/kernel')), ptr(self.P(w + 'basic_lstm_cell/bias')),
                   ptr(self.P(w + 'attention_layer/kernel')),
                   ptr(self.P('Manipulation/dynamic_decoder/output_projection/kernel')),
                   ptr(tokens), self.T, ptr(self.logits), ptr(out_tokens), ptr(lengths), ptr(self.ws),
                   self.ws_bytes, st)

    def forward_teacher(self, exact=False):
        """Teacher-forced decoders + loss = mean over test_k of the masked CE
        (model_induction.py:788-819).  Returns pred_action [B, test_k, T, A]."""
        A, T, R2, tk, st = self.cfg.action_space, self.T, self.R2, self.tk, self._st()
        self._decode(self.d_ttok, None, None, exact)
        self._call('d2p_seq_weights', ptr(self.d_tlen), R2, tk, 1.0 / tk, T, ptr(self.w), ptr(self.runlen), st)
        self._call('d2p_softmax_ce', ptr(self.logits), T, R2, A, ptr(self.d_ttok), ptr(self.d_tlen),
                   ptr(self.runlen), ptr(self.w), ptr(self.rowloss), None, ptr(self.loss), 0, st)
        return self.logits.permute(1, 0, 2).reshape(self.B, tk, T, A).contiguous()

    # ------------------------------------------------------------------ training
    def G(self, name):
        e = self.pm[name]
        return self.grads[e.offset:e.offset + e.size]

    def _alloc_train(self):
        if hasattr(self, 'grads'):
            return
        cfg, lib, dev = self.cfg, self.lib, self.dev
        B, k, tk, T, H, R, R2, F = self.B, self.k, self.tk, self.T, self.H, self.R, self.R2, self.F
        A, Pd = cfg.action_space, cfg.per_dim
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.grads, self.adam_m, self.adam_v = z(self.pm.total), z(self.pm.total), z(self.pm.total)
        self.adam_state = torch.zeros(8, dtype=torch.float64, device=dev)
        self.lr, self.clip = cfg.learning_rate, 20.0
        for li in range(self.conv_desc.n_layers):
            sc = 'Demo_Encoder/State_Encoder/conv%d' % (li + 1)
            l, bn = self.conv_desc.layers[li], sc + '/bn_act/BatchNorm/'
            l.dw, l.db = ptr(self.G(sc + '/Conv/weights')), ptr(self.G(sc + '/Conv/biases'))
            l.dgamma, l.dbeta = ptr(self.G(bn + 'gamma')), ptr(self.G(bn + 'beta'))
        self.tr = dict(
            emb=z(T, R2, H), xa=z(T, R2, 2 * H), hs=z(T, R2, H), cs=z(T, R2, H), gates=z(T, R2, 4 * H),
            ctx=z(R2, H), hc=z(T, R2, 2 * H), att=z(T, R2, H), alpha=z(T, B * k * tk * T),
            dlogits=z(T, R2, A), datt=z(T, R2, H), dhc=z(R2, 2 * H), dh=z(R2, H), dxa=z(R2, 2 * H), tmp=z(R2, H),
            dh0=z(R2, H), dc0=[z(R2, H), z(R2, H)], demb=z(T, R2, H), h0=z(R2, H), c0=z(R2, H), zero=z(R2, H),
            hT=z(R2, H), cT=z(R2, H), ones=torch.ones(R2, dtype=torch.int32, device=dev),
            dkeys=z(T, R, H), dvalues=z(T, R, H), dX=z(T, R, F + Pd), dfeat=z(T, R, F),
            dhT=z(R, H), dcT=z(R, H), dsum_h=z(B, H), dsum_c=z(B, H), edh0=z(R, H), edc0=z(R, H))
        ws = max(self.ws_bytes, lib.d2p_lstm_seq_bwd_ws_bytes(T, R, H), lib.d2p_lstm_seq_bwd_ws_bytes(1, R2, H),
                 lib.d2p_adam_ws_bytes(), lib.d2p_embed_shifted_bwd_ws_bytes(A + 1, H, R2, T),
                 lib.d2p_luong_pool_attention_train_ws_bytes(B, k, tk, H))
        if ws > self.ws_bytes:
            self.ws_bytes = (ws + 255) // 256 * 256
            self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)
        if self.use_tc:     # operand arena for the backward products ([T*R, 4H] dZ operands, K = T*R sums)
            big = lib.d2p_gemm_tc_ws_bytes(T * max(R, R2), 4 * H, 4 * H) + (16 << 20)
            if big > self.tc_scratch.numel():
                self.tc_scratch = torch.zeros(big, dtype=torch.uint8, device=dev)

    def forward_train(self):
        """Encoder + teacher-forced attention decoder, keeping what the backward pass needs.
        loss = mean over test_k of the masked cross-entropies (model_induction.py:788-819)."""
        self._alloc_train()
        cfg, st, call, tr = self.cfg, self._st(), self._call, self.tr
        B, k, tk, T, H, R, R2 = self.B, self.k, self.tk, self.T, self.H, self.R, self.R2
        A = cfg.action_space
        if self.fold_memory_layer:
            raise NotImplementedError('training needs the keys tensor (fold_memory_layer=False)')
        self.encode()
        w = 'Manipulation/dynamic_decoder/pooling_attention_wrapper/'
        K, bias = self.P(w + 'basic_lstm_cell/kernel'), self.P(w + 'basic_lstm_cell/bias')
        Wa = self.P(w + 'attention_layer/kernel')
        proj = self.P('Manipulation/dynamic_decoder/output_projection/kernel')
        table = self.P('Manipulation/Token_Embedding/embedding_map')
        # teacher-forced inputs: <s> (id A+1, out of the table's range -> zero row), then the gt tokens
        call('d2p_embed_shifted', ptr(table), A + 1, H, ptr(self.d_ttok), R2, T, A + 1, ptr(tr['emb']), st)
        call('d2p_seq_weights', ptr(self.d_tlen), R2, tk, 1.0 / tk, T, ptr(self.w), ptr(self.runlen), st)
        # the reference's swapped initial state (model_induction.py:674-676): cell c := h summary, h := c summary
        call('d2p_group_bcast', ptr(self.c_sum), B, tk, H, 1.0, ptr(tr['h0']), 0, st)
        call('d2p_group_bcast', ptr(self.h_sum), B, tk, H, 1.0, ptr(tr['c0']), 0, st)
        for t in range(T):
            att_prev = tr['att'][t - 1] if t else tr['zero']
            h_prev, c_prev = (tr['hs'][t - 1], tr['cs'][t - 1]) if t else (tr['h0'], tr['c0'])
            call('d2p_concat_cols', ptr(tr['emb'][t]), H, ptr(att_prev), H, R2, ptr(tr['xa'][t]), st)
            call('d2p_lstm_seq_fwd', ptr(tr['xa'][t]), 1, R2, 2 * H, H, ptr(tr['ones']), ptr(h_prev), ptr(c_prev),
                 ptr(K), ptr(bias), 1.0, ptr(tr['hs'][t]), ptr(tr['hT']), ptr(tr['cT']), ptr(tr['gates'][t]),
                 ptr(tr['cs'][t]), 3, st)
            call('d2p_luong_pool_attention_train_fwd', ptr(tr['hs'][t]), H, ptr(self.keys), ptr(self.Y),
                 ptr(self.d_demo_len), B, k, tk, T, H, ptr(tr['ctx']), H, ptr(tr['alpha'][t]), ptr(self.ws),
                 self.ws_bytes, st)
            call('d2p_concat_cols', ptr(tr['hs'][t]), H, ptr(tr['ctx']), H, R2, ptr(tr['hc'][t]), st)
            call('d2p_gemm', 0, 0, R2, H, 2 * H, 1.0, ptr(tr['hc'][t]), 2 * H, ptr(Wa), H, 0.0, ptr(tr['att'][t]), H,
                 None, st)
        call('d2p_gemm', 0, 0, T * R2, A, H, 1.0, ptr(tr['att']), H, ptr(proj), A, 0.0, ptr(self.logits), A, None, st)
        call('d2p_softmax_ce', ptr(self.logits), T, R2, A, ptr(self.d_ttok), ptr(self.d_tlen), ptr(self.runlen),
             ptr(self.w), ptr(self.rowloss), ptr(tr['dlogits']), ptr(self.loss), 0, st)

    def backward_train(self):
        """Back-propagation through the attention decoder (BPTT over the T decoder steps), the memory
        layer, the demonstration encoder and the frame encoder into self.grads."""
        cfg, st, call, tr = self.cfg, self._st(), self._call, self.tr
        B, k, tk, T, H, R, R2, F = self.B, self.k, self.tk, self.T, self.H, self.R, self.R2, self.F
        A, Pd = cfg.action_space, cfg.per_dim
        w = 'Manipulation/dynamic_decoder/pooling_attention_wrapper/'
        Kn, bn = w + 'basic_lstm_cell/kernel', w + 'basic_lstm_cell/bias'
        Wan, projn = w + 'attention_layer/kernel', 'Manipulation/dynamic_decoder/output_projection/kernel'
        self.grads.zero_()
        tr['dkeys'].zero_()
        tr['dvalues'].zero_()
        # output projection: logits = att * proj
        call('d2p_gemm', 1, 0, H, A, T * R2, 1.0, ptr(tr['att']), H, ptr(tr['dlogits']), A, 0.0, ptr(self.G(projn)), A,
             None, st)
        call('d2p_gemm', 0, 1, T * R2, H, A, 1.0, ptr(tr['dlogits']), A, ptr(self.P(projn)), A, 0.0, ptr(tr['datt']), H,
             None, st)
        for t in reversed(range(T)):
            last = t == T - 1
            h_prev, c_prev = (tr['hs'][t - 1], tr['cs'][t - 1]) if t else (tr['h0'], tr['c0'])
            if not last:    # attention_t also feeds the cell input of step t+1
                call('d2p_split_cols', ptr(tr['dxa']), 2 * H, H, H, R2, ptr(tr['tmp']), st)
                call('d2p_axpby', ptr(tr['tmp']), 1.0, ptr(tr['datt'][t]), 1.0, R2 * H, st)
            # attention layer [h ; ctx] * W_a
            call('d2p_gemm', 0, 1, R2, 2 * H, H, 1.0, ptr(tr['datt'][t]), H, ptr(self.P(Wan)), H, 0.0, ptr(tr['dhc']),
                 2 * H, None, st)
            call('d2p_split_cols', ptr(tr['dhc']), 2 * H, 0, H, R2, ptr(tr['dh']), st)
            if not last:
                call('d2p_axpby', ptr(tr['dh0']), 1.0, ptr(tr['dh']), 1.0, R2 * H, st)
            dctx = tr['dhc'].view(-1)[H:]          # columns H..2H of dhc, row stride 2H
            call('d2p_luong_pool_attention_train_bwd', ptr(tr['hs'][t]), H, ptr(self.keys), ptr(self.Y),
                 ptr(self.d_demo_len), ptr(tr['alpha'][t]), ptr(dctx), 2 * H, B, k, tk, T, H, ptr(tr['dh']), H,
                 ptr(tr['dkeys']), ptr(tr['dvalues']), ptr(self.ws), self.ws_bytes, st)
            dc_in = None if last else tr['dc0'][(t + 1) & 1]
            call('d2p_lstm_seq_bwd', ptr(tr['xa'][t]), 1, R2, 2 * H, H, ptr(tr['ones']), ptr(h_prev), ptr(c_prev),
                 ptr(self.P(Kn)), ptr(tr['hs'][t]), ptr(tr['gates'][t]), ptr(tr['cs'][t]), ptr(tr['dh']), None,
                 ptr(dc_in), ptr(tr['dxa']), ptr(self.G(Kn)), ptr(self.G(bn)), ptr(tr['dh0']), ptr(tr['dc0'][t & 1]),
                 ptr(self.ws), self.ws_bytes, 3, st)
            call('d2p_split_cols', ptr(tr['dxa']), 2 * H, 0, H, R2, ptr(tr['demb'][t]), st)
        call('d2p_gemm', 1, 0, 2 * H, H, T * R2, 1.0, ptr(tr['hc']), 2 * H, ptr(tr['datt']), H, 0.0, ptr(self.G(Wan)), H,
             None, st)
        call('d2p_embed_shifted_bwd', ptr(tr['demb']), A + 1, H, ptr(self.d_ttok), R2, T, A + 1,
             ptr(self.G('Manipulation/Token_Embedding/embedding_map')), ptr(self.ws), self.ws_bytes, st)
        # initial state (swapped): d c_summary from dh0, d h_summary from dc0; summaries = mean over k
        call('d2p_group_sum', ptr(tr['dh0']), B, tk, H, 1.0, ptr(tr['dsum_c']), 0, st)
        call('d2p_group_sum', ptr(tr['dc0'][0]), B, tk, H, 1.0, ptr(tr['dsum_h']), 0, st)
        call('d2p_group_bcast', ptr(tr['dsum_h']), B, k, H, 1.0 / k, ptr(tr['dhT']), 0, st)
        call('d2p_group_bcast', ptr(tr['dsum_c']), B, k, H, 1.0 / k, ptr(tr['dcT']), 0, st)
        # memory layer keys = values * W_mem
        Wm = 'AttnMechanism/memory_layer/kernel'
        call('d2p_gemm', 1, 0, H, H, T * R, 1.0, ptr(self.Y), H, ptr(tr['dkeys']), H, 0.0, ptr(self.G(Wm)), H, None, st)
        call('d2p_gemm', 0, 1, T * R, H, H, 1.0, ptr(tr['dkeys']), H, ptr(self.P(Wm)), H, 1.0, ptr(tr['dvalues']), H,
             None, st)
        # demonstration encoder (values = its outputs) and the frame encoder
        sc = 'Demo_Encoder/rnn/basic_lstm_cell/'
        call('d2p_lstm_seq_bwd', ptr(self.X), T, R, F + Pd, H, ptr(self.d_demo_len), None, None,
             ptr(self.P(sc + 'kernel')), ptr(self.Y), ptr(self.gates), ptr(self.cells), ptr(tr['dvalues']),
             ptr(tr['dhT']), ptr(tr['dcT']), ptr(tr['dX']), ptr(self.G(sc + 'kernel')), ptr(self.G(sc + 'bias')),
             ptr(tr['edh0']), ptr(tr['edc0']), ptr(self.ws), self.ws_bytes, 3, st)
        call('d2p_split_cols', ptr(tr['dX']), F + Pd, 0, F, T * R, ptr(tr['dfeat']), st)
        call('d2p_conv_encoder_bwd', C.byref(self.conv_desc), ptr(self.d_frames), ptr(tr['dfeat']),
             ptr(self.conv_saved), int(self.is_train), ptr(self.ws), self.ws_bytes, st)

    def optimizer_step(self, world=1):
        """clip_by_global_norm(20) + Adam, shared with the other models (reference trainer.py:102-109)."""
        from .dp import allreduce_flat_gradients
        scale = allreduce_flat_gradients(self.grads, world)
        decay = 10000 if self.cfg.lr_weight_decay else 0
        self._call('d2p_clip_adam_step', ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v),
                   self.pm.total, self.lr, 0.9, 0.999, 1e-8, self.clip, scale, decay, ptr(self.adam_state),
                   ptr(self.ws), self.ws_bytes, self._st())
        self.lib.d2p_tc_new_step()

    def train_step(self, batch, world=1):
        """Public API: host batch in, loss out."""
        self.stage_batch(batch)
        self.forward_train()
        self.backward_train()
        self.optimizer_step(world)
        loss = float(self.loss[0].item())
        flags = C.c_int(0)
        check(self.lib.d2p_device_error(C.byref(flags)), 'd2p_device_error')
        if flags.value:
            raise _lib.D2PError('a persistent-kernel step barrier timed out (flags=%d)' % flags.value)
        return loss

    def step_count(self):
        return int(self.adam_state[0].item()) if hasattr(self, 'adam_state') else 0

    def global_norm(self):
        return float(self.adam_state[3].item())

    TIE_TOL = 2e-4     # see Engine.TIE_TOL

    def greedy(self, exact=None):
        """Greedy action decode of every unseen demo: (logits [B,test_k,T,A], lengths [B,test_k]).
        exact=None (default): tensor-core engine, repeated on the exact fp32 engine if any executed
        arg-max has a top-2 margin within TIE_TOL (d2p_greedy_near_ties), so the token ids are those
        of fp32 arithmetic; True / False force the fp32 / tensor-core engine."""
        A, T, R2, tk = self.cfg.action_space, self.T, self.R2, self.tk
        toks = torch.zeros(T, R2, dtype=torch.int32, device=self.dev)
        lens = torch.zeros(R2, dtype=torch.int32, device=self.dev)
        on_fp32 = bool(exact) or not self.use_tc
        self._decode(None, toks, lens, on_fp32)
        self.greedy_path = 'fp32' if on_fp32 else 'tensor-core'
        if exact is None and self.use_tc:
            cnt = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self._call('d2p_greedy_near_ties', ptr(self.logits), T, R2, A, ptr(lens), self.TIE_TOL, ptr(cnt),
                       self._st())
            self.greedy_near_ties = int(cnt.item())
            if self.greedy_near_ties:
                # the whole chain again in fp32 arithmetic: encoder (BatchNorm moving statistics are
                # put back so that a train-mode model does not advance them twice) and decoder
                self.greedy_path = 'fp32 (re-evaluated: %d near-tie arg-maxes)' % self.greedy_near_ties
                state = self.state.clone()
                self.encode(exact=True)
                self.state.copy_(state)
                self._decode(None, toks, lens, True)
        self.greedy_tokens = toks
        return (self.logits.permute(1, 0, 2).reshape(self.B, tk, T, A).contiguous(),
                lens.view(self.B, tk))


class InductionModel(object):
    """Reference-facing facade for `--model induction_baseline`."""

    def __init__(self, config, debug_information=False, is_train=True, global_step=None, world_size=1, **kw):
        self.world = int(world_size)
        from .config import D2PConfig
        from .model import config_from_namespace
        self.config = config if isinstance(config, D2PConfig) else config_from_namespace(config)
        self.engine = InductionEngine(self.config, is_train=is_train, **kw)
        self.batch_size = self.config.batch_size
        self.loss, self.output = None, []
        self.report_loss, self.report_accuracy, self.report_hist = {}, {}, {}
        # program-related evaler fetches are empty lists in the reference (model_induction.py:850-875)
        self.pred_program = self.greedy_pred_program = []
        self.program_len = self.greedy_pred_program_len = []
        self.ground_truth_program = []

    def get_feed_dict(self, batch_chunk, step=None, is_training=True):
        for key in ('s_h', 'per', 'demo_len', 'test_a_h', 'test_a_h_tokens', 'test_demo_len'):
            if key not in batch_chunk:
                raise KeyError('batch_chunk is missing %s' % key)
        return batch_chunk

    def run_train_step(self, feed):
        self.loss = self.engine.train_step(feed, world=self.world)
        return self.loss

    def run_train_steps(self, feeds):
        for feed in feeds:
            yield self.run_train_step(feed)

    def run_eval_step(self, feed, greedy=True):
        eng = self.engine
        eng.stage_batch(feed)
        eng.encode()
        pred = eng.forward_teacher()
        torch.cuda.current_stream(eng.dev).synchronize()
        self.loss = float(eng.loss[0])
        self.report_loss = {'avg_action_loss': self.loss}
        self.output = [np.asarray(feed['test_a_h']), pred.cpu().numpy()]
        # report surface of models/baselines/model_induction.py:788-862
        from .metrics import demo_sequence_stats
        B, tk, T = eng.B, eng.tk, eng.T
        t_tok = torch.as_tensor(np.asarray(feed['test_a_h_tokens'])).long().to(eng.dev).view(B, tk, T)
        t_len = torch.as_tensor(np.asarray(feed['test_demo_len'])).long().to(eng.dev).view(B, tk)
        st = demo_sequence_stats(pred, t_tok, t_len, t_len)
        self.report_accuracy = {'avg_action_token_acc': st['token_acc'], 'avg_action_seq_acc': st['seq_acc'],
                                'avg_action_seq_all_acc': st['seq_all_acc']}
        self.report_hist = {}
        if greedy:
            g, gl = eng.greedy()
            self.greedy_pred_action, self.greedy_pred_action_len = g.cpu().numpy(), gl.cpu().numpy()
            gst = demo_sequence_stats(g, t_tok, gl.long(), t_len)
            self.report_loss['greedy_avg_action_loss'] = gst['loss']
            self.report_accuracy.update({'greedy_avg_action_token_acc': gst['token_acc'],
                                         'greedy_avg_action_seq_acc': gst['seq_acc'],
                                         'greedy_avg_action_seq_all_acc': gst['seq_all_acc']})
        return self.loss

    def state_dict(self):
        eng = self.engine
        p, s = eng.params.cpu().numpy(), eng.state.cpu().numpy()
        out = {e.name: p[e.offset:e.offset + e.size].reshape(e.shape).copy() for e in eng.pm}
        out.update({e.name: s[e.offset:e.offset + e.size].reshape(e.shape).copy() for e in eng.sm})
        return out

    def load_state_dict(self, d, trainable_only=False):
        eng = self.engine
        p, s = eng.params.cpu().numpy(), eng.state.cpu().numpy()
        for e in eng.pm:
            if e.name in d:
                p[e.offset:e.offset + e.size] = np.asarray(d[e.name], np.float32).reshape(-1)
        for e in eng.sm:
            if e.name in d and not trainable_only:
                s[e.offset:e.offset + e.size] = np.asarray(d[e.name], np.float32).reshape(-1)
        eng.params.copy_(torch.from_numpy(p))
        eng.state.copy_(torch.from_numpy(s))
