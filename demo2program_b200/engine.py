"""Host orchestration of the demo2program train / eval step on one B200.

Python only sequences C-ABI calls into libd2p.so (hand-written sm_100a CUDA);
PyTorch is the device allocator, the stream provider and the NCCL plumbing.
There is no autograd on the hot path: the backward pass is an explicit reverse
sequence of `*_bwd` ops over preallocated buffers, so the whole step (forward,
backward, clip + Adam) is a fixed launch sequence that is captured once into a
CUDA graph and replayed.

Graph restated (reference models/model_full.py:370-599, 918-1079; baselines
model_summarizer.py:363-397, model_synthesis.py:325-358):
  frames -> State_Encoder CNN -> Demo_Encoder LSTM (per demo, shared weights)
         -> [summarizer/full] avg (h,c) over k -> SecondPathEncoder LSTM
         -> rn_pool (+ mean in `full`) -> Program decoder (teacher forcing)
         -> [full] Action decoder + Per decoder from each demo's (h,c)
  loss = program + mean_k action + mean_k per.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, FcBn, check, ptr
from .manifest import build_manifests
from .dp import BucketedAllReduce, allreduce_flat_gradients


def _al(n, a=64):
    return (int(n) + a - 1) // a * a


class Engine:
    """Preallocated buffers + launch sequence for one model/config."""
    N_GRAD_STREAMS = 2

    def __init__(self, cfg, device='cuda:0', seed=0, is_train=True,
                 frames_dtype=np.uint8, flat_params=None, flat_state=None,
                 world_size=1, use_graph=True, use_tc=True, concurrent=True, token_tables=True,
                 compact_decoders=True, overlap_allreduce=None):
        self.lib = _lib.load()
        # token_tables: the teacher-forced token decoders feed embedding rows, so their input
        # products have at most V+1 distinct rows: gates = (E*Wx + b)[token] in the forward,
        # dWx = E^T*S and dE = S*Wx^T from the per-token sums S of dZ in the backward, instead of
        # [T*R]-row products (same arithmetic, different association; False = row-by-row products)
        self.token_tables = bool(token_tables)
        # compact_decoders: the action / perception decoder recurrences of `full` run in compact form
        # (32 CTAs each) side by side with the program decoder instead of one after the other
        self.compact_decoders = bool(compact_decoders) and os.environ.get('D2P_COMPACT_DECODERS', '1') != '0'
        # D2P_LSTM_WIDE for the encoder / second-path recurrences: D2P_WIDE_LSTM = 1 (default) forward only,
        # 2 forward and backward, 0 never.  Measured at C2 (profiles/r02p_*): 4 x 80-row tiles shorten each
        # of these recurrences by 5-17 us, but in the backward pass the 128 CTAs leave only 20 SMs to the
        # deferred weight-gradient products, whose tail then ends 80 us later - a net loss there.
        self.wide_lstm = os.environ.get('D2P_WIDE_LSTM', '1') != '0'
        self.wide_lstm_bwd = os.environ.get('D2P_WIDE_LSTM', '1') == '2'
        self._bwd_wide = False
        # scheduled sampling (reference models/model_full.py:59-67, 414-423): the program / action decoders
        # run step by step and feed, with the scheduled probability, a token sampled from their own
        # output instead of the ground truth.  sched_p_override >= 0 replaces the schedule (tests).
        self.sched = bool(cfg.scheduled_sampling) and bool(is_train)
        self.sched_seed = int(seed) & 0xFFFFFFFF
        self.sched_p_override = -1.0
        if self.sched:
            self.token_tables = True
        if not torch.cuda.is_available():
            raise _lib.D2PError('demo2program_b200 needs a CUDA device (no CPU '
                                'fallback)')
        if cfg.model == 'induction_baseline':
            raise NotImplementedError('induction_baseline uses InductionEngine')
        cfg.validate()
        self.cfg = cfg
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.is_train = bool(is_train)
        self.world = int(world_size)
        self.use_graph = use_graph
        self.frames_u8 = np.dtype(frames_dtype) == np.uint8
        self.pm, self.sm = build_manifests(cfg)
        f32 = dict(dtype=torch.float32, device=self.dev)
        p0 = self.pm.init_flat(seed) if flat_params is None else np.asarray(flat_params, np.float32)
        s0 = self.sm.init_flat(seed) if flat_state is None else np.asarray(flat_state, np.float32)
        self.params = torch.from_numpy(p0.copy()).to(self.dev)
        self.state = torch.from_numpy(s0.copy()).to(self.dev)
        self.grads = torch.zeros(self.pm.total, **f32)
        self.adam_m = torch.zeros(self.pm.total, **f32)
        self.adam_v = torch.zeros(self.pm.total, **f32)
        self.adam_state = torch.zeros(8, dtype=torch.float64, device=self.dev)
        self.lr, self.clip = cfg.learning_rate, 20.0
        self._alloc()
        # tensor-core engine arena: packed bf16 hi/lo operands (activations in
        # `scratch`, weights cached per step in `cache`)
        self.use_tc = bool(use_tc)
        if self.use_tc:
            R, T, H = self.R, max(self.T, cfg.max_program_len), self.H
            big = self.lib.d2p_gemm_tc_ws_bytes(T * R, 4 * H, 4 * H) + (8 << 20)
            self.tc_scratch = torch.zeros(big, dtype=torch.uint8, device=self.dev)
            self.tc_cache = torch.zeros(12 * self.pm.total + (16 << 20), dtype=torch.uint8,
                                        device=self.dev)
        # side streams for the independent branches (three decoders, two RN pools)
        self.concurrent = bool(concurrent)
        self.side_streams = [torch.cuda.Stream(self.dev, priority=-1) for _ in range(3)] if self.concurrent else []
        # the latency-bound chain runs at high priority; the throughput-bound weight-gradient
        # products (grad_stream, default = lowest priority) fill whatever SMs are left
        self.main_stream = torch.cuda.Stream(self.dev, priority=-1) if self.concurrent else None
        self.grad_streams = [torch.cuda.Stream(self.dev) for _ in range(self.N_GRAD_STREAMS)] if self.concurrent else []
        self._grad_rr = 0
        for s_ in self.side_streams + self.grad_streams:
            self._ws_side[s_.cuda_stream] = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.dev)
        self._all_side = self.side_streams + self.grad_streams
        self.tc_side = [torch.zeros_like(self.tc_scratch) for _ in self._all_side] if self.use_tc else []
        self._tc_bind()
        self._graph = None
        self._graph_key = None
        # data parallelism: the flat gradient buffer is summed over the ranks in two pieces on a
        # communication stream under the backward pass (dp.BucketedAllReduce), inside the step's
        # CUDA graph.  overlap_allreduce=False (or D2P_DP_OVERLAP=0): one all-reduce of the whole
        # buffer between a forward+backward graph and a clip+Adam graph.
        if overlap_allreduce is None:
            overlap_allreduce = os.environ.get('D2P_DP_OVERLAP', '1') != '0'
        self.dp = None
        self._dp_active = False
        if self.world > 1 and overlap_allreduce and self.concurrent:
            self.dp = BucketedAllReduce(self.grads, self.pm, self.world, self.dev)

    @property
    def ws(self):
        """Scratch buffer of the current stream (side branches have their own)."""
        return self._ws_side.get(torch.cuda.current_stream(self.dev).cuda_stream, self._ws_main)

    def _tc_bind(self):
        """The arena registration is library-global: bind this engine's buffers
        before issuing work (also invalidates the packed-weight cache)."""
        if self.use_tc and not getattr(self, '_force_fp32', False):
            self.lib.d2p_tc_configure(ptr(self.tc_scratch), self.tc_scratch.numel(),
                                      ptr(self.tc_cache), self.tc_cache.numel(), 1)
            for s_, arena in zip(self._all_side, self.tc_side):
                self.lib.d2p_tc_bind_stream(s_.cuda_stream, ptr(arena), arena.numel())
        else:
            self.lib.d2p_tc_configure(None, 0, None, 0, 0)

    # ------------------------------------------------------------------ params
    def P(self, name):
        e = self.pm[name]
        return self.params[e.offset:e.offset + e.size]

    def G(self, name):
        e = self.pm[name]
        return self.grads[e.offset:e.offset + e.size]

    def S(self, name):
        e = self.sm[name]
        return self.state[e.offset:e.offset + e.size]

    def _fcbn(self, scope):
        f, b = scope + '/fully_connected/', scope + '/bn_act/BatchNorm/'
        s = FcBn()
        s.w, s.b = ptr(self.P(f + 'weights')), ptr(self.P(f + 'biases'))
        s.gamma, s.beta = ptr(self.P(b + 'gamma')), ptr(self.P(b + 'beta'))
        s.moving_mean = ptr(self.S(b + 'moving_mean'))
        s.moving_var = ptr(self.S(b + 'moving_variance'))
        s.dw, s.db = ptr(self.G(f + 'weights')), ptr(self.G(f + 'biases'))
        s.dgamma, s.dbeta = ptr(self.G(b + 'gamma')), ptr(self.G(b + 'beta'))
        return s

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        cfg, lib = self.cfg, self.lib
        B, k, T, H = cfg.batch_size, cfg.k, cfg.max_demo_len, cfg.num_lstm_cell_units
        L, V, A, Pd = cfg.max_program_len, cfg.dim_program_token, cfg.action_space, cfg.per_dim
        R = B * k
        self.B, self.k, self.T, self.H, self.R = B, k, T, H, R
        dev = self.dev
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        zi = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        # ---- inputs (device) + pinned staging (host) ----
        fdt = torch.uint8 if self.frames_u8 else torch.float32
        self.d_frames = torch.zeros(B, k, T, cfg.h, cfg.w, cfg.depth, dtype=fdt, device=dev)
        self.d_demo_len_f = z(R)
        self.d_prog_len_f = z(B)
        self.d_demo_len = zi(R)
        self.d_prog_len = zi(B)
        self.d_prog_tok = zi(B, L)
        self.d_act_tok = zi(R, T)
        self.d_per = z(R, T, Pd)
        self.h_in = {
            's_h': torch.zeros(B, k, T, cfg.h, cfg.w, cfg.depth, dtype=fdt).pin_memory(),
            'demo_len': torch.zeros(R).pin_memory(),
            'program_len': torch.zeros(B).pin_memory(),
            'program_tokens': torch.zeros(B, L, dtype=torch.int32).pin_memory(),
            'a_h_tokens': torch.zeros(R, T, dtype=torch.int32).pin_memory(),
            'per': torch.zeros(R, T, Pd).pin_memory(),
        }
        self.h_loss = torch.zeros(4).pin_memory()
        # sticky step-barrier time-out words (LSTM recurrence, fused conv encoder): refreshed at the end of
        # every step on the device (d2p_device_error_async) and read back with the loss
        self.err_dev = torch.zeros(2, dtype=torch.int32, device=dev)
        self.h_err = torch.zeros(2, dtype=torch.int32).pin_memory()
        # ---- conv encoder ----
        d = ConvDesc()
        d.B, d.k, d.T, d.h, d.w, d.d = B, k, T, cfg.h, cfg.w, cfg.depth
        d.frames_dtype = _lib.D2P_U8 if self.frames_u8 else _lib.D2P_F32
        chans = cfg.conv_channels()
        d.n_layers = len(chans)
        for li, (_, cout) in enumerate(chans):
            sc = 'Demo_Encoder/State_Encoder/conv%d' % (li + 1)
            l = d.layers[li]
            l.w, l.b = ptr(self.P(sc + '/Conv/weights')), ptr(self.P(sc + '/Conv/biases'))
            bn = sc + '/bn_act/BatchNorm/'
            l.gamma, l.beta = ptr(self.P(bn + 'gamma')), ptr(self.P(bn + 'beta'))
            l.moving_mean = ptr(self.S(bn + 'moving_mean'))
            l.moving_var = ptr(self.S(bn + 'moving_variance'))
            l.dw, l.db = ptr(self.G(sc + '/Conv/weights')), ptr(self.G(sc + '/Conv/biases'))
            l.dgamma, l.dbeta = ptr(self.G(bn + 'gamma')), ptr(self.G(bn + 'beta'))
            l.cout = cout
        self.conv_desc = d
        self.F = lib.d2p_conv_encoder_feature_dim(C.byref(d))
        assert self.F == cfg.feature_dim(), (self.F, cfg.feature_dim())
        F = self.F
        self.conv_saved = z(lib.d2p_conv_encoder_saved_floats(C.byref(d)))
        ws = lib.d2p_conv_encoder_ws_bytes(C.byref(d))
        self.feat = z(T, R, F)
        self.dfeat = z(T, R, F)
        # ---- LSTMs ----
        def lstm_bufs(Tn, Rn, with_dx_in=None):
            b = {'y': z(Tn, Rn, H), 'hT': z(Rn, H), 'cT': z(Rn, H),
                 'gates': z(Tn, Rn, 4 * H), 'cells': z(Tn, Rn, H),
                 'dh0': z(Rn, H), 'dc0': z(Rn, H)}
            return b
        ws = max(ws, lib.d2p_lstm_seq_bwd_ws_bytes(max(T, L), R, H))
        self.enc = lstm_bufs(T, R)
        self.model = cfg.model
        two_pass = cfg.model in ('full', 'summarizer')
        if two_pass:
            self.sum1_h, self.sum1_c = z(B, H), z(B, H)
            self.init2_h, self.init2_c = z(R, H), z(R, H)
            self.sec = lstm_bufs(T, R)
            self.dy1 = z(T, R, H)
            self.rn_saved_h = z(lib.d2p_rn_pool_saved_floats(B, k, H))
            self.rn_saved_c = z(lib.d2p_rn_pool_saved_floats(B, k, H))
            self.rn_hold = {s_: z(lib.d2p_rn_pool_bwd_hold_floats(B, k, H)) for s_ in 'hc'}
            ws = max(ws, lib.d2p_rn_pool_ws_bytes(B, k, H))
            self.fc = {(s, f): self._fcbn('demo_%s_summary/rn_pool/%s' % (s, f))
                       for s in 'hc' for f in ('fc1', 'fc2')}
        self.dsum_h, self.dsum_c = z(B, H), z(B, H)       # decoder init state
        self.dh2, self.dc2 = z(R, H), z(R, H)             # grads wrt per-demo (h,c)
        self.pool_dh, self.pool_dc = z(R, H), z(R, H)     # rn_pool / mean contributions to them
        # ---- program decoder ----
        self.prog = lstm_bufs(L, B)
        self.prog.update(X=z(L, B, H), logits=z(L, B, V), dlogits=z(L, B, V),
                         dy=z(L, B, H), dX=z(L, B, H), rowloss=z(L * B), w=z(B),
                         runlen=zi(B))
        if cfg.model == 'full':
            self.act = lstm_bufs(T, R)
            self.act.update(X=z(T, R, H), logits=z(T, R, A), dlogits=z(T, R, A),
                            dy=z(T, R, H), dX=z(T, R, H), rowloss=z(T * R), w=z(R),
                            runlen=zi(R))
            self.per = lstm_bufs(T, R)
            self.per.update(X=z(T, R, H), logits=z(T, R, Pd), dlogits=z(T, R, Pd),
                            dy=z(T, R, H), dX=z(T, R, H), rowloss=z(T * R),
                            per_tm=z(T, R, Pd),
                            fc_saved=z(lib.d2p_fc_bn_saved_floats(T * R, H, k)))
            self.per_fc = self._fcbn('Per_Decoder/Per_Encoder/fc2')
            ws = max(ws, lib.d2p_fc_bn_ws_bytes(T * R, H, k))
        # per-token tables of the token decoders: [rows + 1 (the <s> row = bias), 4H] and dZ sums [rows, 4H]
        self.prog.update(tab=z(V + 2, 4 * H), tsum=z(V + 1, 4 * H))
        if cfg.model == 'full':
            self.act.update(tab=z(A + 2, 4 * H), tsum=z(A + 1, 4 * H))
        if self.sched:
            for b_, Rn, Ln in ([(self.prog, B, L)] + ([(self.act, R, T)] if cfg.model == 'full' else [])):
                b_.update(fed=zi(Rn, Ln), sampled=zi(Rn, Ln), steplen=zi(Ln, Rn),
                          hpp=[z(Rn, H), z(Rn, H)], cpp=[z(Rn, H), z(Rn, H)])
        ws = max(ws, lib.d2p_adam_ws_bytes(),
                 lib.d2p_embed_shifted_bwd_ws_bytes(V + 1, 4 * H, B, L),
                 lib.d2p_embed_shifted_bwd_ws_bytes(A + 1, 4 * H, R, T))
        self.ws_bytes = _al(ws, 256)
        self._ws_main = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)
        self._ws_side = {}
        self.loss = z(4)   # [total, program, action, per]
        self.out_bvl = None

    # ------------------------------------------------------------------ helpers
    def _st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _stamp(self, name):
        """Developer timeline (tools/timeline.py): with self.timeline set to an int64 device
        tensor, record the GPU global timer at this point of the current stream."""
        tl = getattr(self, 'timeline', None)
        if tl is None:
            return
        names = self.timeline_names
        if name not in names:
            names.append(name)
        check(self.lib.d2p_debug_stamp(ptr(tl), names.index(name), self._st()), 'stamp')

    def _call(self, name, *args):
        check(getattr(self.lib, name)(*args), name)

    def _lstm_fwd(self, X, Tn, Rn, In, lens, h0, c0, scope, b, phases=3, compact=False, wide=False):
        """compact: D2P_LSTM_COMPACT - the recurrence runs on a 32-CTA grid whose CTAs walk all row
        tiles, so that independent recurrences share the GPU (the library falls back to one CTA per
        row tile when the shape has a single tile).  wide: D2P_LSTM_WIDE - the recurrence runs alone
        (encoder / second-path LSTMs): rows split evenly over 4 x 32 CTAs."""
        if compact:
            phases |= 8
        if wide and self.wide_lstm:
            phases |= 16
        self._call('d2p_lstm_seq_fwd', ptr(X), Tn, Rn, In, self.H, ptr(lens), ptr(h0), ptr(c0),
                   ptr(self.P(scope + 'kernel')), ptr(self.P(scope + 'bias')), 1.0,
                   ptr(b['y']), ptr(b['hT']), ptr(b['cT']), ptr(b['gates']), ptr(b['cells']),
                   phases, self._st())

    def _lstm_bwd_call(self, X, Tn, Rn, In, lens, h0, c0, scope, b, dY, dhT, dcT, dX, phases):
        if self._bwd_wide and self.wide_lstm_bwd:
            phases |= 16
        self._call('d2p_lstm_seq_bwd', ptr(X), Tn, Rn, In, self.H, ptr(lens), ptr(h0), ptr(c0),
                   ptr(self.P(scope + 'kernel')), ptr(b['y']), ptr(b['gates']), ptr(b['cells']),
                   ptr(dY), ptr(dhT), ptr(dcT), ptr(dX), ptr(self.G(scope + 'kernel')),
                   ptr(self.G(scope + 'bias')), ptr(b['dh0']), ptr(b['dc0']), ptr(self.ws),
                   self.ws_bytes, phases, self._st())

    def _lstm_bwd(self, X, Tn, Rn, In, lens, h0, c0, scope, b, dY, dhT, dcT, dX, token_fn=None, wide=False):
        """BPTT recurrence (+ dX) on the current stream; the parameter-gradient products
        (dWx, dWh, db: large, throughput-bound) are handed to the gradient stream so they
        overlap with the latency-bound recurrences that follow.  token_fn (token decoders with
        token_tables): no dX, no [T*R]-row dWx product - token_fn() forms dWx and dE from the
        per-token sums of dZ, behind the dWh product on the same gradient stream."""
        nodwx = 4 if token_fn is not None else 0      # D2P_LSTM_BWD_NO_DWX
        self._bwd_wide = bool(wide)
        if token_fn is not None:
            dX = None
        if not self.concurrent:
            self._lstm_bwd_call(X, Tn, Rn, In, lens, h0, c0, scope, b, dY, dhT, dcT, dX, 3 | nodwx)
            if token_fn is not None:
                token_fn()
            return
        self._lstm_bwd_call(X, Tn, Rn, In, lens, h0, c0, scope, b, dY, dhT, dcT, dX, 1)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        gs = self._next_grad_stream()
        gs.wait_event(ev)
        with torch.cuda.stream(gs):
            self._lstm_bwd_call(X, Tn, Rn, In, lens, h0, c0, scope, b, dY, dhT, dcT, None, 2 | nodwx)
            if token_fn is not None:
                token_fn()

    def _next_grad_stream(self):
        """Gradient streams are used round-robin: the weight-gradient work of one LSTM (operand
        packing, two products, a column sum) is a serial chain, several of them overlap."""
        s_ = self.grad_streams[self._grad_rr % len(self.grad_streams)]
        self._grad_rr += 1
        return s_

    def _deferred(self, fn):
        """Run parameter-gradient-only work behind the current stream's work, on the
        (low-priority) gradient stream: it is joined once, before the optimizer."""
        if not self.concurrent:
            fn()
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        gs = self._next_grad_stream()
        gs.wait_event(ev)
        with torch.cuda.stream(gs):
            fn()

    def _gemm(self, ta, tb, M, N, K, alpha, A, lda, Bm, ldb, beta, Cm, ldc, bias=None):
        self._call('d2p_gemm', int(ta), int(tb), M, N, K, alpha, ptr(A), lda, ptr(Bm), ldb, beta,
                   ptr(Cm), ldc, ptr(bias), self._st())

    def _token_gates(self, emb_name, scope, rows, tokens, Rn, Tn, b):
        """Hoisted input product of a teacher-forced token decoder without the [T*R]-row GEMM:
        tab[v] = E[v]*Wx + bias for the `rows` embedding rows, tab[rows] = bias (the out-of-range
        <s> id embeds to zeros), gates[t, r] = tab[id(t, r)]."""
        H = self.H
        E, W, bias = self.P(emb_name), self.P(scope + 'kernel'), self.P(scope + 'bias')
        self._gemm(0, 0, rows, 4 * H, H, 1.0, E, H, W, 4 * H, 0.0, b['tab'], 4 * H, bias=bias)
        self._call('d2p_axpby', ptr(bias), 1.0, ptr(b['tab'][rows]), 0.0, 4 * H, self._st())
        if self.sched:      # the gate rows are gathered step by step from the tokens actually fed
            return
        self._call('d2p_embed_shifted', ptr(b['tab']), rows + 1, 4 * H, ptr(tokens), Rn, Tn, rows,
                   ptr(b['gates']), self._st())

    def _sched_decoder_fwd(self, b, rows, gt_tokens, Rn, Tn, h0, c0, scope, Wp, vocab, decoder_id):
        """Scheduled-sampling forward of a token decoder (seq2seq.ScheduledEmbeddingTrainingHelper,
        reference models/model_full.py:414-423): per step the hoisted gate row of the token fed
        (tab[token], <s> at t = 0), one recurrence step, the output projection, then the draw of
        the token fed to the next step (d2p_sched_sample_step).  Leaves y / gates / cells / logits
        of all steps exactly as the teacher-forced path does, and the fed tokens in b['fed'] - the
        backward pass is the teacher-forced one over those."""
        H, call, S = self.H, self._call, self._st
        W, bias = self.P(scope + 'kernel'), self.P(scope + 'bias')
        call('d2p_step_lens', ptr(b['runlen']), Rn, Tn, ptr(b['steplen']), S())
        for t in range(Tn):
            call('d2p_embed_shifted_step', ptr(b['tab']), rows + 1, 4 * H, ptr(b['fed']), Rn, Tn, t, rows,
                 ptr(b['gates'][t]), S())
            h_prev, c_prev = (h0, c0) if t == 0 else (b['hpp'][t & 1], b['cpp'][t & 1])
            call('d2p_lstm_seq_fwd', ptr(b['X']), 1, Rn, H, H, ptr(b['steplen'][t]), ptr(h_prev), ptr(c_prev),
                 ptr(W), ptr(bias), 1.0, ptr(b['y'][t]), ptr(b['hpp'][(t + 1) & 1]), ptr(b['cpp'][(t + 1) & 1]),
                 ptr(b['gates'][t]), ptr(b['cells'][t]), 2, S())
            self._gemm(0, 0, Rn, vocab, H, 1.0, b['y'][t], H, Wp, vocab, 0.0, b['logits'][t], vocab)
            call('d2p_sched_sample_step', ptr(b['logits'][t]), Rn, vocab, ptr(gt_tokens), Tn, t,
                 ptr(self.adam_state), int(self.cfg.scheduled_sampling_decay_steps), float(self.sched_p_override),
                 self.sched_seed, decoder_id, ptr(b['fed']), ptr(b['sampled']), S())

    def _token_grads(self, emb_name, scope, rows, tokens, Rn, Tn, b):
        """Backward of _token_gates from the dZ left in b['gates']: S[v] = sum of the dZ rows that
        were fed token v (fixed order), dE += S*Wx^T, dWx += E^T*S."""
        H = self.H
        E, W = self.P(emb_name), self.P(scope + 'kernel')
        b['tsum'].zero_()
        self._call('d2p_embed_shifted_bwd', ptr(b['gates']), rows, 4 * H, ptr(tokens), Rn, Tn, rows,
                   ptr(b['tsum']), ptr(self.ws), self.ws_bytes, self._st())
        self._gemm(0, 1, rows, H, 4 * H, 1.0, b['tsum'], 4 * H, W, 4 * H, 1.0, self.G(emb_name), H)
        self._gemm(1, 0, H, 4 * H, rows, 1.0, E, H, b['tsum'], 4 * H, 1.0, self.G(scope + 'kernel'), 4 * H)

    # ------------------------------------------------------------------ inputs
    def stage_batch(self, batch):
        """numpy batch (reference feed_dict keys, models/model_full.py:185-206)
        -> pinned host staging -> async H2D.  Returns bytes copied."""
        h = self.h_in
        fr = np.asarray(batch['s_h'])
        h['s_h'].numpy()[...] = fr.astype(np.uint8 if self.frames_u8 else np.float32, copy=False)
        h['demo_len'].numpy()[...] = np.asarray(batch['demo_len'], np.float32).reshape(-1)
        h['program_len'].numpy()[...] = np.asarray(batch['program_len'], np.float32).reshape(-1)
        h['program_tokens'].numpy()[...] = np.asarray(batch['program_tokens'], np.int32)
        if self.model == 'full':
            h['a_h_tokens'].numpy()[...] = np.asarray(batch['a_h_tokens'], np.int32).reshape(self.R, self.T)
            h['per'].numpy()[...] = np.asarray(batch['per'], np.float32).reshape(self.R, self.T, -1)
        return self.upload()

    def upload(self):
        """Async H2D of the staged batch (part of the timed e2e region)."""
        h = self.h_in
        n = 0
        pairs = [(self.d_frames, h['s_h']), (self.d_demo_len_f, h['demo_len']),
                 (self.d_prog_len_f, h['program_len']), (self.d_prog_tok, h['program_tokens'])]
        if self.model == 'full':
            pairs += [(self.d_act_tok, h['a_h_tokens']), (self.d_per, h['per'])]
        for d, s in pairs:
            d.copy_(s, non_blocking=True)
            n += s.numel() * s.element_size()
        return n

    # ------------------------------------------------------------------ branches
    def _parallel(self, fns, streams=None, join=True):
        """Run independent branches of the step concurrently: fns[0] on the current
        stream, the others on side streams that fork from / rejoin it (also valid
        inside CUDA-graph capture).  Every branch has its own scratch buffers."""
        if not self.concurrent or len(fns) == 1:
            for fn in fns:
                fn()
            return
        main = torch.cuda.current_stream(self.dev)
        ev = torch.cuda.Event()
        ev.record(main)
        used = []
        for fn, s in zip(fns[1:], streams or self.side_streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                fn()
            used.append(s)
        fns[0]()
        if not join:          # the caller joins later (forward tail overlapping the backward head)
            return used
        for s in used:
            main.wait_stream(s)

    # ------------------------------------------------------------------ forward
    def forward(self, train_stats=None, open_tail=False):
        """open_tail (train step only): do not wait for the action / per decoders at the end - the
        backward pass starts with the program decoder and the summary pools, which do not depend on
        them, and joins their stream itself."""
        cfg = self.cfg
        B, k, T, H, R, F = self.B, self.k, self.T, self.H, self.R, self.F
        L, V = cfg.max_program_len, cfg.dim_program_token
        tr = int(self.is_train if train_stats is None else train_stats)
        call, S = self._call, self._st
        self._tc_bind()              # (re)bind this engine's arenas; re-pack weights
        self._stamp('fwd start')
        call('d2p_len_to_int', ptr(self.d_demo_len_f), ptr(self.d_demo_len), R, S())
        call('d2p_len_to_int', ptr(self.d_prog_len_f), ptr(self.d_prog_len), B, S())
        if self.model == 'full':
            # instance normalisers / run lengths shared by the action and per decoders
            call('d2p_seq_weights', ptr(self.d_demo_len), R, k, 1.0 / k, T, ptr(self.act['w']),
                 ptr(self.act['runlen']), S())
        # The decoders' hoisted input products (teacher-forced embeddings x Wx) do not depend
        # on the encoder: issue them on the side streams now, under the encoder recurrence.
        p = self.prog

        def prog_in():
            # the gradient buffer is cleared here, on a side stream under the encoder recurrence, instead
            # of at the head of the backward pass (this branch is joined before the program decoder runs)
            self.grads.zero_()
            self._grads_zeroed = True
            if self.token_tables:
                self._token_gates('Program_Decoder/Token_Embedding/embedding_map',
                                  'Program_Decoder/dynamic_decoder/basic_lstm_cell/', V + 1,
                                  self.d_prog_tok, B, L, p)
                return
            call('d2p_embed_shifted', ptr(self.P('Program_Decoder/Token_Embedding/embedding_map')),
                 V + 1, H, ptr(self.d_prog_tok), B, L, V + 1, ptr(p['X']), S())
            self._lstm_fwd(p['X'], L, B, H, self.d_prog_len, None, None,
                           'Program_Decoder/dynamic_decoder/basic_lstm_cell/', p, phases=1)

        def act_in():
            a = self.act
            if self.token_tables:
                self._token_gates('Action_Decoder/Token_Embedding/embedding_map',
                                  'Action_Decoder/dynamic_decoder/basic_lstm_cell/', cfg.action_space + 1,
                                  self.d_act_tok, R, T, a)
                return
            call('d2p_embed_shifted', ptr(self.P('Action_Decoder/Token_Embedding/embedding_map')),
                 cfg.action_space + 1, H, ptr(self.d_act_tok), R, T, cfg.action_space + 1, ptr(a['X']), S())
            self._lstm_fwd(a['X'], T, R, H, self.d_demo_len, None, None,
                           'Action_Decoder/dynamic_decoder/basic_lstm_cell/', a, phases=1)

        def per_in():
            q = self.per
            call('d2p_rtp_to_trp', ptr(self.d_per), R, T, cfg.per_dim, ptr(q['per_tm']), S())
            call('d2p_fc_bn_fwd', ptr(q['per_tm']), T * R, cfg.per_dim, H, 1, k, 0, C.byref(self.per_fc),
                 ptr(q['X']), ptr(q['fc_saved']), tr, ptr(self.ws), self.ws_bytes, S())
            self._lstm_fwd(q['X'], T, R, H, self.d_demo_len, None, None,
                           'Per_Decoder/dynamic_decoder/basic_lstm_cell/', q, phases=1)

        # the frame encoder is one cooperative kernel over (nearly) all SMs: the side branches
        # fork behind it and overlap the encoder recurrence (96 of 148 SMs) instead
        call('d2p_conv_encoder_fwd', C.byref(self.conv_desc), ptr(self.d_frames), ptr(self.feat),
             ptr(self.conv_saved), tr, ptr(self.ws), self.ws_bytes, S())
        self._stamp('conv fwd done')
        self._ev_prog_in = None
        if self.concurrent:
            main = torch.cuda.current_stream(self.dev)
            ev0 = torch.cuda.Event()
            ev0.record(main)
            s1, s2 = self.side_streams[:2]
            s1.wait_event(ev0)
            with torch.cuda.stream(s1):
                if self.model == 'full':
                    act_in()
                prog_in()
                self._ev_prog_in = torch.cuda.Event()
                self._ev_prog_in.record(s1)
            if self.model == 'full':
                s2.wait_event(ev0)
                with torch.cuda.stream(s2):
                    per_in()
        else:
            prog_in()
            if self.model == 'full':
                act_in()
                per_in()
        self._lstm_fwd(self.feat, T, R, F, self.d_demo_len, None, None,
                       'Demo_Encoder/rnn/basic_lstm_cell/', self.enc, wide=True)
        self._stamp('encoder lstm fwd done')
        if self.model in ('full', 'summarizer'):
            call('d2p_group_sum', ptr(self.enc['hT']), B, k, H, 1.0 / k, ptr(self.sum1_h), 0, S())
            call('d2p_group_sum', ptr(self.enc['cT']), B, k, H, 1.0 / k, ptr(self.sum1_c), 0, S())
            call('d2p_group_bcast', ptr(self.sum1_h), B, k, H, 1.0, ptr(self.init2_h), 0, S())
            call('d2p_group_bcast', ptr(self.sum1_c), B, k, H, 1.0, ptr(self.init2_c), 0, S())
            self._lstm_fwd(self.enc['y'], T, R, H, self.d_demo_len, self.init2_h, self.init2_c,
                           'SecondPathEncoder/rnn/basic_lstm_cell/', self.sec, wide=True)
            self._stamp('second-path lstm fwd done')
            fin = self.sec
        else:
            fin = self.enc
        self.fin = fin
        full = self.model == 'full'
        if full:
            A, Pd = cfg.action_space, cfg.per_dim
            a, q = self.act, self.per
            side_by_side = self.concurrent and self.compact_decoders

            def act_fwd():
                Wa = self.P('Action_Decoder/dynamic_decoder/output_projection/kernel')
                if self.sched:
                    self._sched_decoder_fwd(a, A + 1, self.d_act_tok, R, T, fin['hT'], fin['cT'],
                                            'Action_Decoder/dynamic_decoder/basic_lstm_cell/', Wa, A, 2)
                else:
                    self._lstm_fwd(a['X'], T, R, H, a['runlen'], fin['hT'], fin['cT'],
                                   'Action_Decoder/dynamic_decoder/basic_lstm_cell/', a, phases=2,
                                   compact=side_by_side)
                    self._gemm(0, 0, T * R, A, H, 1.0, a['y'], H, Wa, A, 0.0, a['logits'], A)
                call('d2p_softmax_ce', ptr(a['logits']), T, R, A, ptr(self.d_act_tok), ptr(self.d_demo_len),
                     ptr(a['runlen']), ptr(a['w']), ptr(a['rowloss']), ptr(a['dlogits']),
                     ptr(self.loss[2:]), 0, S())
                self._stamp('action decoder fwd done')

            def per_fwd():
                self._lstm_fwd(q['X'], T, R, H, a['runlen'], fin['hT'], fin['cT'],
                               'Per_Decoder/dynamic_decoder/basic_lstm_cell/', q, phases=2, compact=side_by_side)
                Wq = self.P('Per_Decoder/dynamic_decoder/output_projection/kernel')
                self._gemm(0, 0, T * R, Pd, H, 1.0, q['y'], H, Wq, Pd, 0.0, q['logits'], Pd)
                call('d2p_sigmoid_ce', ptr(q['logits']), T, R, Pd, ptr(self.d_per), ptr(self.d_demo_len),
                     ptr(a['runlen']), ptr(a['w']), ptr(q['rowloss']), ptr(q['dlogits']),
                     ptr(self.loss[3:]), 0, S())
                self._stamp('per decoder fwd done')

            # The action and perception decoders (forward AND backward) depend only on the second-path
            # final states - not on the summary pools or the program decoder.  Their chain (two k*B-row
            # recurrences forward, two backward) is as long as pools -> program decoder forward ->
            # backward -> pools backward, so it forks HERE, on side stream 1, and is joined in the
            # backward pass where the three initial-state gradients meet.  Forward: in compact form
            # (32 CTAs each) the two recurrences run side by side (side streams 1 and 2) and leave 84
            # SMs to the pools and the 32-CTA program decoder; otherwise they are chained (two 96-CTA
            # cooperative grids cannot be resident together).
            if self.concurrent:
                main = torch.cuda.current_stream(self.dev)
                ev_fin = torch.cuda.Event()
                ev_fin.record(main)
                s1, s2 = self.side_streams[:2]
                s1.wait_event(ev_fin)
                if side_by_side:
                    s2.wait_event(ev_fin)
                    with torch.cuda.stream(s1):
                        act_fwd()
                    with torch.cuda.stream(s2):     # per_in ran on side stream 2 (stream order)
                        per_fwd()
                else:
                    with torch.cuda.stream(s1):
                        act_fwd()
                        s1.wait_stream(s2)          # the per decoder's hoisted input product
                        per_fwd()
        if self.model in ('full', 'summarizer'):
            def pool(s, out, saved):
                def run():
                    call('d2p_rn_pool_fwd', ptr(fin[s + 'T']), B, k, H, C.byref(self.fc[(s, 'fc1')]),
                         C.byref(self.fc[(s, 'fc2')]), ptr(out), ptr(saved), tr, ptr(self.ws),
                         self.ws_bytes, S())
                    if self.model == 'full':   # mean + rn_pool (model_full.py:357-359)
                        call('d2p_group_sum', ptr(fin[s + 'T']), B, k, H, 1.0 / k, ptr(out), 1, S())
                return run
            self._parallel([pool('h', self.dsum_h, self.rn_saved_h),
                            pool('c', self.dsum_c, self.rn_saved_c)], streams=self.side_streams[2:])
            self._stamp('pools fwd done')
        else:
            self._aggregate_fwd(fin)

        def prog_fwd():   # program decoder (teacher forcing)
            call('d2p_seq_weights', ptr(self.d_prog_len), B, 1, 1.0, L, ptr(p['w']), ptr(p['runlen']), S())
            if self._ev_prog_in is not None:
                torch.cuda.current_stream(self.dev).wait_event(self._ev_prog_in)
            Wp = self.P('Program_Decoder/dynamic_decoder/output_projection/kernel')
            if self.sched:
                self._sched_decoder_fwd(p, V + 1, self.d_prog_tok, B, L, self.dsum_h, self.dsum_c,
                                        'Program_Decoder/dynamic_decoder/basic_lstm_cell/', Wp, V, 1)
            else:
                self._lstm_fwd(p['X'], L, B, H, p['runlen'], self.dsum_h, self.dsum_c,
                               'Program_Decoder/dynamic_decoder/basic_lstm_cell/', p, phases=2)
                self._gemm(0, 0, L * B, V, H, 1.0, p['y'], H, Wp, V, 0.0, p['logits'], V)
            call('d2p_softmax_ce', ptr(p['logits']), L, B, V, ptr(self.d_prog_tok), ptr(self.d_prog_len),
                 ptr(p['runlen']), ptr(p['w']), ptr(p['rowloss']), ptr(p['dlogits']),
                 ptr(self.loss[1:]), 0, S())
            self._stamp('program decoder fwd done')

        prog_fwd()
        if not full:
            if self.concurrent:
                torch.cuda.current_stream(self.dev).wait_stream(self.side_streams[0])
            call('d2p_axpby', ptr(self.loss[1:]), 1.0, ptr(self.loss), 0.0, 1, S())
            return
        if not self.concurrent:
            act_fwd()
            per_fwd()
        self._fwd_open = bool(open_tail and self.concurrent)
        self._stamp('fwd done')
        if not self._fwd_open:
            self._join_act_per()
            self._total_loss()

    def _join_act_per(self):
        """Joins the action / per decoder branch (side streams 1 and 2) into the current stream."""
        if self.concurrent:
            main = torch.cuda.current_stream(self.dev)
            main.wait_stream(self.side_streams[0])
            main.wait_stream(self.side_streams[1])

    def _aggregate_fwd(self, fin):
        """synthesis_baseline: demo_aggregation over the k per-demonstration final states (reference
        models/baselines/model_synthesis.py:336-358).  'concat' hands a [B, k*H] state to a
        BasicLSTMCell of H units - the reference graph only builds for k == 1, where it is the
        identity; the same restriction applies here."""
        cfg = self.cfg
        B, k, H = self.B, self.k, self.H
        call, S = self._call, self._st
        agg = cfg.demo_aggregation
        if agg == 'concat':
            if k != 1:
                raise ValueError("demo_aggregation='concat' gives the decoder a [B, k*H] initial state; the "
                                 "reference's BasicLSTMCell(H) cannot take it either - use k == 1, avgpool or maxpool")
            agg = 'avgpool'
        if agg == 'avgpool':
            call('d2p_group_sum', ptr(fin['hT']), B, k, H, 1.0 / k, ptr(self.dsum_h), 0, S())
            call('d2p_group_sum', ptr(fin['cT']), B, k, H, 1.0 / k, ptr(self.dsum_c), 0, S())
        elif agg == 'maxpool':
            if not hasattr(self, 'agg_arg_h'):
                self.agg_arg_h = torch.zeros(B, H, dtype=torch.int32, device=self.dev)
                self.agg_arg_c = torch.zeros(B, H, dtype=torch.int32, device=self.dev)
            call('d2p_group_max', ptr(fin['hT']), B, k, H, ptr(self.dsum_h), ptr(self.agg_arg_h), S())
            call('d2p_group_max', ptr(fin['cT']), B, k, H, ptr(self.dsum_c), ptr(self.agg_arg_c), S())
        else:
            raise ValueError('Unknown demo aggregation type: %s' % agg)

    def _aggregate_bwd(self, dh, dc):
        """Gradient of _aggregate_fwd wrt the per-demonstration final states -> self.dh2 / self.dc2."""
        B, k, H = self.B, self.k, self.H
        call, S = self._call, self._st
        if self.cfg.demo_aggregation == 'maxpool':
            call('d2p_group_max_bwd', ptr(dh), ptr(self.agg_arg_h), B, k, H, ptr(self.dh2), S())
            call('d2p_group_max_bwd', ptr(dc), ptr(self.agg_arg_c), B, k, H, ptr(self.dc2), S())
        else:
            call('d2p_group_bcast', ptr(dh), B, k, H, 1.0 / k, ptr(self.dh2), 0, S())
            call('d2p_group_bcast', ptr(dc), B, k, H, 1.0 / k, ptr(self.dc2), 0, S())

    def _total_loss(self):
        """total = program + action + per"""
        call, S = self._call, self._st
        call('d2p_axpby', ptr(self.loss[1:]), 1.0, ptr(self.loss), 0.0, 1, S())
        call('d2p_axpby', ptr(self.loss[2:]), 1.0, ptr(self.loss), 1.0, 1, S())
        call('d2p_axpby', ptr(self.loss[3:]), 1.0, ptr(self.loss), 1.0, 1, S())

    # ------------------------------------------------------------------ backward
    def backward(self):
        cfg = self.cfg
        B, k, T, H, R, F = self.B, self.k, self.T, self.H, self.R, self.F
        L, V = cfg.max_program_len, cfg.dim_program_token
        tr = int(self.is_train)
        call, S = self._call, self._st
        fin = self.fin
        if not getattr(self, '_grads_zeroed', False):
            self.grads.zero_()
        self._grads_zeroed = False
        self._grad_rr = 0
        p = self.prog
        self._stamp('bwd start')

        def prog_bwd():
            Wp = self.P('Program_Decoder/dynamic_decoder/output_projection/kernel')
            self._deferred(lambda: self._gemm(
                1, 0, H, V, L * B, 1.0, p['y'], H, p['dlogits'], V, 1.0,
                self.G('Program_Decoder/dynamic_decoder/output_projection/kernel'), V))
            self._gemm(0, 1, L * B, H, V, 1.0, p['dlogits'], V, Wp, V, 0.0, p['dy'], H)
            if self.token_tables:
                self._lstm_bwd(p['X'], L, B, H, p['runlen'], self.dsum_h, self.dsum_c,
                               'Program_Decoder/dynamic_decoder/basic_lstm_cell/', p, p['dy'], None, None,
                               None, token_fn=lambda: self._token_grads(
                                   'Program_Decoder/Token_Embedding/embedding_map',
                                   'Program_Decoder/dynamic_decoder/basic_lstm_cell/', V + 1,
                                   p['fed'] if self.sched else self.d_prog_tok, B, L, p))
            else:
                self._lstm_bwd(p['X'], L, B, H, p['runlen'], self.dsum_h, self.dsum_c,
                               'Program_Decoder/dynamic_decoder/basic_lstm_cell/', p, p['dy'], None, None,
                               p['dX'])
                self._deferred(lambda: call(
                    'd2p_embed_shifted_bwd', ptr(p['dX']), V + 1, H, ptr(self.d_prog_tok), B, L, V + 1,
                    ptr(self.G('Program_Decoder/Token_Embedding/embedding_map')), ptr(self.ws),
                    self.ws_bytes, S()))
            # p['dh0'], p['dc0'] = grad wrt (demo_h_summary, demo_c_summary)
            self._stamp('program decoder bwd done')

        def pool_bwd(s, dsum, saved, dF):
            def run():
                if self.model == 'full':   # mean term
                    call('d2p_group_bcast', ptr(dsum), B, k, H, 1.0 / k, ptr(dF), 0, S())
                else:
                    dF.zero_()
                # data path now; the fc weight / bias gradients only read `hold`: gradient stream
                def part(ph):
                    call('d2p_rn_pool_bwd', ptr(fin[s + 'T']), B, k, H, C.byref(self.fc[(s, 'fc1')]),
                         C.byref(self.fc[(s, 'fc2')]), ptr(dsum), ptr(saved), ptr(dF), tr, ptr(self.ws),
                         self.ws_bytes, ph, ptr(self.rn_hold[s]), S())
                part(1)
                self._deferred(lambda: part(2))
            return run

        if self.model == 'full':
            A, Pd = cfg.action_space, cfg.per_dim
            a, q = self.act, self.per

            def act_bwd():
                Wa = self.P('Action_Decoder/dynamic_decoder/output_projection/kernel')
                self._deferred(lambda: self._gemm(
                    1, 0, H, A, T * R, 1.0, a['y'], H, a['dlogits'], A, 1.0,
                    self.G('Action_Decoder/dynamic_decoder/output_projection/kernel'), A))
                self._gemm(0, 1, T * R, H, A, 1.0, a['dlogits'], A, Wa, A, 0.0, a['dy'], H)
                if self.token_tables:
                    self._lstm_bwd(a['X'], T, R, H, a['runlen'], fin['hT'], fin['cT'],
                                   'Action_Decoder/dynamic_decoder/basic_lstm_cell/', a, a['dy'], None, None,
                                   None, token_fn=lambda: self._token_grads(
                                       'Action_Decoder/Token_Embedding/embedding_map',
                                       'Action_Decoder/dynamic_decoder/basic_lstm_cell/', A + 1,
                                       a['fed'] if self.sched else self.d_act_tok, R, T, a))
                else:
                    self._lstm_bwd(a['X'], T, R, H, a['runlen'], fin['hT'], fin['cT'],
                                   'Action_Decoder/dynamic_decoder/basic_lstm_cell/', a, a['dy'], None, None,
                                   a['dX'])
                    self._deferred(lambda: call(
                        'd2p_embed_shifted_bwd', ptr(a['dX']), A + 1, H, ptr(self.d_act_tok), R, T, A + 1,
                        ptr(self.G('Action_Decoder/Token_Embedding/embedding_map')), ptr(self.ws),
                        self.ws_bytes, S()))
                self._stamp('action decoder bwd done')

            def per_bwd():
                Wq = self.P('Per_Decoder/dynamic_decoder/output_projection/kernel')
                self._deferred(lambda: self._gemm(
                    1, 0, H, Pd, T * R, 1.0, q['y'], H, q['dlogits'], Pd, 1.0,
                    self.G('Per_Decoder/dynamic_decoder/output_projection/kernel'), Pd))
                self._gemm(0, 1, T * R, H, Pd, 1.0, q['dlogits'], Pd, Wq, Pd, 0.0, q['dy'], H)
                self._lstm_bwd(q['X'], T, R, H, a['runlen'], fin['hT'], fin['cT'],
                               'Per_Decoder/dynamic_decoder/basic_lstm_cell/', q, q['dy'], None, None,
                               q['dX'])
                # Per_Encoder fc + BatchNorm: parameter gradients only (its input is data)
                self._deferred(lambda: call(
                    'd2p_fc_bn_bwd', ptr(q['per_tm']), T * R, Pd, H, 1, k, 0, C.byref(self.per_fc),
                    ptr(q['dX']), ptr(q['fc_saved']), None, tr, ptr(self.ws), self.ws_bytes, S()))
                self._stamp('per decoder bwd done')

            # The summary pools only need the program decoder's (dh0, dc0): back-propagate them
            # right behind it (h on this stream, c on the third side stream) while the two
            # k*B-row decoders - which cannot share the SMs as persistent kernels - still run.
            def prog_then_pools():
                prog_bwd()
                self._parallel([pool_bwd('h', p['dh0'], self.rn_saved_h, self.pool_dh),
                                pool_bwd('c', p['dc0'], self.rn_saved_c, self.pool_dc)],
                               streams=self.side_streams[2:])
                self._stamp('pools bwd done')

            def act_then_per_bwd():
                act_bwd()
                if self.concurrent:   # the per decoder's forward may still run on side stream 2
                    torch.cuda.current_stream(self.dev).wait_stream(self.side_streams[1])
                per_bwd()
            if getattr(self, '_fwd_open', False):
                # the action / per branch forked in the forward pass (side stream 1) simply continues
                # with its backward - it never waits for the program decoder; joined below
                with torch.cuda.stream(self.side_streams[0]):
                    act_then_per_bwd()
                prog_then_pools()
                self._join_act_per()
                self._total_loss()
                self._fwd_open = False
            else:
                self._parallel([prog_then_pools, act_then_per_bwd])
            self._stamp('decoders + pools bwd joined')
            self._reduce_bucket(0)
            # dh2 = d(action init) + d(per init) + d(pools)
            call('d2p_add3', ptr(a['dh0']), ptr(q['dh0']), ptr(self.pool_dh), ptr(self.dh2), R * H, S())
            call('d2p_add3', ptr(a['dc0']), ptr(q['dc0']), ptr(self.pool_dc), ptr(self.dc2), R * H, S())
        else:
            prog_bwd()
            if self.model == 'summarizer':
                self._parallel([pool_bwd('h', p['dh0'], self.rn_saved_h, self.dh2),
                                pool_bwd('c', p['dc0'], self.rn_saved_c, self.dc2)])
            self._reduce_bucket(0)
        if self.model in ('full', 'summarizer'):
            sec = self.sec
            self._lstm_bwd(self.enc['y'], T, R, H, self.d_demo_len, self.init2_h, self.init2_c,
                           'SecondPathEncoder/rnn/basic_lstm_cell/', sec, None, self.dh2, self.dc2,
                           self.dy1, wide=True)
            self._stamp('second-path lstm bwd done')
            # init state = mean_i of first-pass finals, broadcast over i
            call('d2p_group_sum', ptr(sec['dh0']), B, k, H, 1.0, ptr(self.sum1_h), 0, S())
            call('d2p_group_sum', ptr(sec['dc0']), B, k, H, 1.0, ptr(self.sum1_c), 0, S())
            call('d2p_group_bcast', ptr(self.sum1_h), B, k, H, 1.0 / k, ptr(self.dh2), 0, S())
            call('d2p_group_bcast', ptr(self.sum1_c), B, k, H, 1.0 / k, ptr(self.dc2), 0, S())
            dY1 = self.dy1
        else:
            self._aggregate_bwd(p['dh0'], p['dc0'])
            dY1 = None
        self._lstm_bwd(self.feat, T, R, F, self.d_demo_len, None, None,
                       'Demo_Encoder/rnn/basic_lstm_cell/', self.enc, dY1, self.dh2, self.dc2,
                       self.dfeat, wide=True)
        self._stamp('encoder lstm bwd done')
        call('d2p_conv_encoder_bwd', C.byref(self.conv_desc), ptr(self.d_frames), ptr(self.dfeat),
             ptr(self.conv_saved), tr, ptr(self.ws), self.ws_bytes, S())
        self._stamp('conv bwd done')
        if self.concurrent:   # parameter-gradient products must land before the optimizer
            for gs in self.grad_streams:
                torch.cuda.current_stream(self.dev).wait_stream(gs)
        self._stamp('bwd done (weight-gradient stream joined)')
        self._reduce_bucket(1)

    def _reduce_bucket(self, b):
        """Data parallelism: everything that writes gradient bucket b has been enqueued (on this
        stream and the gradient streams) - sum it over the ranks on the communication stream."""
        if self.dp is not None and self._dp_active:
            self.dp.reduce(b, [torch.cuda.current_stream(self.dev)] + self.grad_streams)
            if getattr(self, 'timeline', None) is not None:
                with torch.cuda.stream(self.dp.comm):
                    self._stamp('gradient bucket %d summed over the ranks' % b)

    # ------------------------------------------------------------------ optimizer
    def optimizer_step(self):
        """clip_by_global_norm(20) + Adam (reference trainer.py:102-109); the
        gradient is first averaged over ranks with ONE all-reduce of the flat
        buffer when world_size > 1."""
        if self.dp is not None and self.dp.reduced:     # the backward pass reduced the buckets
            scale = self.dp.join(torch.cuda.current_stream(self.dev))
        else:
            scale = allreduce_flat_gradients(self.grads, self.world)
        decay = 10000 if self.cfg.lr_weight_decay else 0
        self._call('d2p_clip_adam_step', ptr(self.params), ptr(self.grads), ptr(self.adam_m),
                   ptr(self.adam_v), self.pm.total, self.lr, 0.9, 0.999, 1e-8, self.clip,
                   scale, decay, ptr(self.adam_state), ptr(self.ws), self.ws_bytes,
                   self._st())
        self._stamp('clip + adam done')

    # ------------------------------------------------------------------ steps
    def _step_body(self, with_opt):
        self._dp_active = bool(with_opt)      # gradient buckets are reduced only when an update follows
        if self.main_stream is None:
            self.forward()
            self.backward()
            if with_opt:
                self.optimizer_step()
            self._record_errors()
            return
        cur = torch.cuda.current_stream(self.dev)
        self.main_stream.wait_stream(cur)
        with torch.cuda.stream(self.main_stream):
            self.forward(open_tail=self.model == 'full')
            self.backward()
            if with_opt:
                self.optimizer_step()
            self._record_errors()
        cur.wait_stream(self.main_stream)

    def _record_errors(self):
        """Copy the sticky device-error words next to the loss (stream-ordered, graph-capturable)."""
        self._call('d2p_device_error_async', ptr(self.err_dev), self._st())

    def _raise_if_failed(self, words):
        """words: the two error words read back with a loss.  Raises if a step barrier timed out; the
        optimizer has skipped that step's update on the device (d2p_clip_adam_step), so parameters
        and Adam slots are those of the last good step."""
        if int(words[0]) or int(words[1]):
            self.check_device()      # synchronises, clears the sticky words and raises

    def _capture(self, fn):
        """Warm up `fn` on a side stream (lazy module loading), then capture it."""
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        n0 = self.lib.d2p_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g, self.lib.d2p_launch_count() - n0

    def _adam_only(self, scale):
        decay = 10000 if self.cfg.lr_weight_decay else 0
        self._call('d2p_clip_adam_step', ptr(self.params), ptr(self.grads), ptr(self.adam_m),
                   ptr(self.adam_v), self.pm.total, self.lr, 0.9, 0.999, 1e-8, self.clip,
                   scale, decay, ptr(self.adam_state), ptr(self.ws), self.ws_bytes, self._st())

    def train_step_device(self, with_opt=True):
        """One train step over the batch already resident in HBM.  The launch
        sequence is captured into CUDA graphs on first use and replayed: one graph
        for the whole step on a single GPU; with data parallelism a forward+backward
        graph, the single NCCL all-reduce of the flat gradients, and a clip+Adam graph."""
        if not self.use_graph:
            self._step_body(with_opt)
            return
        key = bool(with_opt)
        if self._graph is None or self._graph_key != key:
            snap = [t.clone() for t in (self.params, self.state, self.adam_m, self.adam_v,
                                        self.adam_state)]
            if self.world > 1 and (self.dp is None or not with_opt):
                g1, n1 = self._capture(lambda: self._step_body(False))
                g2, n2 = self._capture(lambda: self._adam_only(1.0 / self.world))
                self._graph, self.launches_per_step = (g1, g2), n1 + n2
            else:
                # one graph for the whole step; with data parallelism the bucketed NCCL all-reduces
                # are nodes of that graph (communication stream forked from / joined into the step)
                g, n = self._capture(lambda: self._step_body(with_opt))
                self._graph, self.launches_per_step = g, n
            for t, sv in zip((self.params, self.state, self.adam_m, self.adam_v, self.adam_state),
                             snap):
                t.copy_(sv)
            torch.cuda.synchronize(self.dev)
            self._graph_key = key
        if isinstance(self._graph, tuple):
            self._graph[0].replay()
            if with_opt:
                allreduce_flat_gradients(self.grads, self.world)
                self._graph[1].replay()
        else:
            self._graph.replay()

    def close(self):
        """Drops the captured CUDA graphs (with data parallelism they hold NCCL nodes: release them
        before torch.distributed.destroy_process_group, which otherwise may not return)."""
        torch.cuda.synchronize(self.dev)
        self._graph = None
        self._graph_key = None

    def check_device(self):
        """Fail loudly if a step barrier of a persistent kernel timed out (d2p_device_error):
        the step's results would be garbage.  Synchronises the device."""
        flags = C.c_int(0)
        check(self.lib.d2p_device_error(C.byref(flags)), 'd2p_device_error')
        if flags.value:
            raise _lib.D2PError('a persistent-kernel step barrier timed out (flags=%d: bit 0 LSTM '
                                'recurrence, bit 1 fused conv encoder); results are invalid - were two '
                                'persistent launches that cannot be co-resident issued on different '
                                'streams?' % flags.value)

    def train_step(self, batch):
        """Public API: host batch in, loss out (H2D + step + D2H of the loss)."""
        self.stage_batch(batch)
        self.train_step_device(True)
        self.h_loss.copy_(self.loss, non_blocking=True)
        self.h_err.copy_(self.err_dev, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        self._raise_if_failed(self.h_err)
        return float(self.h_loss[0])

    def _input_pairs(self):
        h = self.h_in
        pairs = [(self.d_frames, h['s_h']), (self.d_demo_len_f, h['demo_len']),
                 (self.d_prog_len_f, h['program_len']), (self.d_prog_tok, h['program_tokens'])]
        if self.model == 'full':
            pairs += [(self.d_act_tok, h['a_h_tokens']), (self.d_per, h['per'])]
        return pairs

    def train_steps(self, batches):
        """Public API for a training loop (reference trainer.py:126-160): iterates host
        batches and yields each step's loss.  The input path is double-buffered: while
        step i runs, batch i+1 is copied into the second pinned staging set and sent to
        the device on a copy stream, and step i's loss is read back after step i+1 has
        been enqueued.  Every step still pays its own H2D and its loss D2H, but they
        overlap the neighbouring steps' kernels and the device queue never drains."""
        dev = self.dev
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, '_pf'):
            pairs = self._input_pairs()
            self._pf = {
                'stream': torch.cuda.Stream(dev),
                'host': [[torch.empty_like(s).pin_memory() for _, s in pairs] for _ in range(2)],
                'dev': [[torch.empty_like(d) for d, _ in pairs] for _ in range(2)],
                'ready': [torch.cuda.Event() for _ in range(2)],
                'free': [torch.cuda.Event() for _ in range(2)],
                'done': [torch.cuda.Event() for _ in range(2)],
                'loss': [torch.zeros(4).pin_memory() for _ in range(2)],
                'err': [torch.zeros(2, dtype=torch.int32).pin_memory() for _ in range(2)],
            }
        pf = self._pf
        keys = ['s_h', 'demo_len', 'program_len', 'program_tokens'] + \
            (['a_h_tokens', 'per'] if self.model == 'full' else [])

        def prefetch(batch, slot, wait_free):
            hs, ds = pf['host'][slot], pf['dev'][slot]
            if wait_free:
                pf['free'][slot].synchronize()     # step that last read this slot has consumed it
            for key, hbuf in zip(keys, hs):
                hbuf.numpy()[...] = np.asarray(batch[key]).reshape(hbuf.shape)
            with torch.cuda.stream(pf['stream']):
                for hbuf, dbuf in zip(hs, ds):
                    dbuf.copy_(hbuf, non_blocking=True)
                pf['ready'][slot].record(pf['stream'])

        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        prefetch(nxt, 0, False)
        i = 0
        while nxt is not None:
            slot = i & 1
            # enqueue step i behind step i-1 (the device queue never drains)
            main.wait_event(pf['ready'][slot])
            for (d, _), sdev in zip(self._input_pairs(), pf['dev'][slot]):
                d.copy_(sdev, non_blocking=True)
            pf['free'][slot].record(main)
            self.train_step_device(True)
            pf['loss'][slot].copy_(self.loss, non_blocking=True)
            pf['err'][slot].copy_(self.err_dev, non_blocking=True)
            pf['done'][slot].record(main)
            # read back step i-1's loss (and its device-error words) while step i runs
            if i >= 1:
                pf['done'][slot ^ 1].synchronize()
                self._raise_if_failed(pf['err'][slot ^ 1])
                yield float(pf['loss'][slot ^ 1][0])
            try:
                nxt = next(it)
            except StopIteration:
                nxt = None
            if nxt is not None:
                prefetch(nxt, slot ^ 1, i >= 1)
            i += 1
        last = (i - 1) & 1
        pf['done'][last].synchronize()
        self._raise_if_failed(pf['err'][last])
        yield float(pf['loss'][last][0])

    # ------------------------------------------------------------------ outputs
    def pred_program(self):
        """[B, V, L] logits, the reference's `pred_program` layout
        (models/model_full.py:486-489)."""
        cfg = self.cfg
        L, V, B = cfg.max_program_len, cfg.dim_program_token, self.B
        out = torch.empty(B, V, L, dtype=torch.float32, device=self.dev)
        self._call('d2p_logits_to_bvl', ptr(self.prog['logits']), L, B, V, ptr(out), self._st())
        return out

    # ------------------------------------------------------------------ greedy decode
    # arg-max margin below which a tensor-core (bf16x3, ~5e-6 relative) greedy decode is repeated
    # on the exact fp32 engine: 40x the product error bound, relative to max(1, max|logit|)
    TIE_TOL = 2e-4

    def _near_ties(self, logits, lengths):
        """Executed positions of a greedy decode whose top-2 logit gap is within TIE_TOL
        (d2p_greedy_near_ties).  Synchronises the stream (one int comes back)."""
        cnt = torch.zeros(1, dtype=torch.int32, device=self.dev)
        Tn, rows, vocab = logits.shape
        self._call('d2p_greedy_near_ties', ptr(logits), Tn, rows, vocab, ptr(lengths), self.TIE_TOL,
                   ptr(cnt), self._st())
        return int(cnt.item())

    def _greedy(self, scope, vocab, end_id, max_len, rows, h0, c0, exact=None, nsl=1):
        """GreedyEmbeddingHelper decode with the decoder under `scope` (reference
        models/model_full.py:424-435).  exact=True runs every contraction on the fp32 SIMT
        engine; exact=False on the tensor-core engine; exact=None (default, what the facade
        ships): tensor cores, and if any executed arg-max has a top-2 margin within TIE_TOL the
        decode is repeated on the fp32 engine, so the token ids are those of fp32 arithmetic."""
        H = self.H
        logits = torch.zeros(max_len, rows, vocab, dtype=torch.float32, device=self.dev)
        tokens = torch.zeros(max_len, rows, dtype=torch.int32, device=self.dev)
        lengths = torch.zeros(rows, dtype=torch.int32, device=self.dev)
        wsb = self.lib.d2p_greedy_ws_bytes(rows, H, H)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        d = scope + '/dynamic_decoder/'

        def run(on_fp32):
            if on_fp32 or not self.use_tc:
                self.lib.d2p_tc_configure(None, 0, None, 0, 0)
            else:
                self._tc_bind()
            try:
                self._call('d2p_lstm_decoder_greedy',
                           ptr(self.P(scope + '/Token_Embedding/embedding_map')), vocab + 1, H,
                           ptr(self.P(d + 'basic_lstm_cell/kernel')), ptr(self.P(d + 'basic_lstm_cell/bias')),
                           ptr(self.P(d + 'output_projection/kernel')), rows, H, vocab, vocab, end_id,
                           max_len, nsl, ptr(h0), ptr(c0), ptr(logits), ptr(tokens), ptr(lengths), ptr(ws), wsb,
                           self._st())
            finally:
                self._tc_bind()
        self.greedy_path = 'fp32' if (exact or not self.use_tc) else 'tensor-core'
        run(bool(exact))
        if exact is None and self.use_tc:
            self.greedy_near_ties = self._near_ties(logits, lengths)
            if self.greedy_near_ties:
                # the whole chain again in fp32 arithmetic: the forward pass that produced (h0, c0)
                # (BatchNorm moving statistics and losses are put back) and the decode
                self.greedy_path = 'fp32 (re-evaluated: %d near-tie arg-maxes)' % self.greedy_near_ties
                state, loss = self.state.clone(), self.loss.clone()
                self._force_fp32 = True
                try:
                    self.forward()
                    run(True)
                finally:
                    self._force_fp32 = False
                    self._tc_bind()
                self.state.copy_(state)
                self.loss.copy_(loss)
        return logits, tokens, lengths

    def greedy_program(self, exact=None):
        """After forward(): (greedy_pred_program [B,V,L], greedy_pred_program_len [B,1],
        tokens [B,L]) - the reference's greedy program decoder outputs."""
        cfg = self.cfg
        L, V, B = cfg.max_program_len, cfg.dim_program_token, self.B
        logits, tokens, lengths = self._greedy('Program_Decoder', V, cfg.program_end_token, L, B,
                                               self.dsum_h, self.dsum_c, exact)
        out = torch.empty(B, V, L, dtype=torch.float32, device=self.dev)
        self._call('d2p_logits_to_bvl', ptr(logits), L, B, V, ptr(out), self._st())
        return out, lengths.view(B, 1), tokens.t().contiguous()

    def greedy_actions(self, exact=None):
        """`full` only: greedy action decoders of all k demos, batched as R = B*k rows.
        Returns (logits [B,k,T,A], lengths [B,k])."""
        cfg = self.cfg
        A, T = cfg.action_space, self.T
        logits, tokens, lengths = self._greedy('Action_Decoder', A, A - 1, T, self.R,
                                               self.fin['hT'], self.fin['cT'], exact, nsl=self.k)
        return (logits.permute(1, 0, 2).reshape(self.B, self.k, T, A).contiguous(),
                lengths.view(self.B, self.k))

    def global_norm(self):
        return float(self.adam_state[3].item())

    def step_count(self):
        return int(self.adam_state[0].item())
