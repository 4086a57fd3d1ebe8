"""TEST INFRASTRUCTURE - CPU oracle of the reference model graphs.

Restates, with torch CPU ops and the helpers in oracle/tf_ops.py, the numeric
path of
  full        reference models/model_full.py:208-599, 620-657, 918-1079
  summarizer  reference models/baselines/model_summarizer.py:264-397, 834-847
  synthesis   reference models/baselines/model_synthesis.py:212-489, 795-808
structured like the reference (k unrolled encoder copies with shared weights,
step-by-step LSTM loops) so it is also a fair stand-in for "the TF1 graph on
CPU" when timed.  PARITY UNPINNED: the reference has no tests for this path
and TF-1.3 is not runnable here (see oracle/tf_ops.py header).

Never imported by the product path.
"""
import numpy as np
import torch

from demo2program_b200.manifest import build_manifests
from . import tf_ops as T


class OracleModel:
    def __init__(self, cfg, flat_params, flat_state, dtype=torch.float64,
                 is_train=True):
        self.cfg = cfg
        self.dtype = dtype
        self.is_train = is_train
        self.pm, self.sm = build_manifests(cfg)
        self.flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
        self.flat.requires_grad_(True)
        self.state = torch.tensor(np.asarray(flat_state), dtype=dtype)
        self.new_state = self.state.clone()
        # test hook: conv_slope_hook(layer, pre_activation, lrelu_value) -> tensor with the same values whose
        # backward uses a slope pattern chosen by the caller (lrelu is not differentiable at 0: a path whose
        # forward rounds differently picks the other one-sided slope for activations within rounding of 0)
        self.conv_slope_hook = None

    # -- parameter access ---------------------------------------------------
    def p(self, name):
        e = self.pm[name]
        return self.flat[e.offset:e.offset + e.size].reshape(e.shape)

    def s(self, name):
        e = self.sm[name]
        return self.state[e.offset:e.offset + e.size].reshape(e.shape)

    def _set_state(self, name, val):
        e = self.sm[name]
        self.new_state[e.offset:e.offset + e.size] = val.detach().reshape(-1)

    def _bn(self, x, scope):
        """BN under `scope`; with k reuse-calls per step the moving stats are
        updated k times in sequence (reference models/ops.py:20-23 with
        updates_collections=None)."""
        b = scope + '/BatchNorm/'
        e_m, e_v = self.sm[b + 'moving_mean'], self.sm[b + 'moving_variance']
        mm = self.new_state[e_m.offset:e_m.offset + e_m.size]
        mv = self.new_state[e_v.offset:e_v.offset + e_v.size]
        y, nmm, nmv = T.batch_norm(x, self.p(b + 'gamma'), self.p(b + 'beta'),
                                   mm.clone(), mv.clone(), self.is_train)
        if self.is_train:
            self._set_state(b + 'moving_mean', nmm)
            self._set_state(b + 'moving_variance', nmv)
        return y

    # -- reference sub-graphs -------------------------------------------------
    def state_encoder(self, s):
        """State_Encoder, reference models/model_full.py:216-231: conv ->
        lrelu -> BN per layer.  s [N,h,w,d] -> [N,F]."""
        x = s
        for li in range(len(self.cfg.conv_channels())):
            sc = 'Demo_Encoder/State_Encoder/conv%d' % (li + 1)
            x = T.conv2d_3x3_s2_same(x, self.p(sc + '/Conv/weights'),
                                     self.p(sc + '/Conv/biases'))
            a = T.lrelu(x)
            if self.conv_slope_hook is not None:
                a = self.conv_slope_hook(li, x, a)
            x = self._bn(a, sc + '/bn_act')
        return x.reshape(x.shape[0], -1)

    def demo_encoder(self, s_h_i, len_i, per_i=None):
        """Demo_Encoder, reference models/model_full.py:235-258."""
        B, Tm = s_h_i.shape[:2]
        f = self.state_encoder(s_h_i.reshape((B * Tm,) + s_h_i.shape[2:]))
        f = f.reshape(B, Tm, -1)
        if per_i is not None:
            f = torch.cat([f, per_i], dim=-1)
        sc = 'Demo_Encoder/rnn/basic_lstm_cell/'
        return T.dynamic_rnn(f, len_i, self.p(sc + 'kernel'), self.p(sc + 'bias'))

    def rn_pool(self, feat, scope):
        """rn_pool, reference models/model_full.py:333-349. feat [B,k,H]."""
        B, k, H = feat.shape
        tile1 = feat.unsqueeze(1).expand(B, k, k, H)
        tile2 = feat.unsqueeze(2).expand(B, k, k, H)
        x = torch.cat([tile1, tile2], dim=3).reshape(B * k * k, 2 * H)
        for fc in ('fc1', 'fc2'):
            sc = '%s/rn_pool/%s' % (scope, fc)
            x = x @ self.p(sc + '/fully_connected/weights') + \
                self.p(sc + '/fully_connected/biases')
            x = self._bn(T.lrelu(x), sc + '/bn_act')
        return x.reshape(B, k, k, H).mean(1).mean(1)

    def token_decoder(self, scope, h0, c0, gt_tokens, seq_len, max_len, vocab):
        """LSTM_Decoder (teacher forcing), reference models/model_full.py:440-490."""
        start = torch.full_like(gt_tokens[:, :1], vocab + 1)   # out of range (F6)
        shifted = torch.cat([start, gt_tokens[:, :-1]], dim=1)
        emb = T.embedding_lookup_gpu(
            self.p(scope + '/Token_Embedding/embedding_map'), shifted)
        d = scope + '/dynamic_decoder/'
        return T.decode_training(
            emb, seq_len, c0, h0, self.p(d + 'basic_lstm_cell/kernel'),
            self.p(d + 'basic_lstm_cell/bias'),
            self.p(d + 'output_projection/kernel'), max_len)

    def token_decoder_scheduled(self, scope, h0, c0, gt_tokens, seq_len, max_len, vocab, sched, decoder, rows,
                                replay=None):
        """LSTM_Decoder with unroll_type='scheduled_sampling' (reference models/model_full.py:414-423)."""
        table = self.p(scope + '/Token_Embedding/embedding_map')
        d = scope + '/dynamic_decoder/'
        p = sched.get('p_override', -1.0)
        if p < 0:
            p = T.scheduled_sampling_prob(sched['step'], self.cfg.scheduled_sampling_decay_steps)
        return T.decode_scheduled(
            lambda ids: T.embedding_lookup_gpu(table, ids), vocab + 1, gt_tokens, seq_len, c0, h0,
            self.p(d + 'basic_lstm_cell/kernel'), self.p(d + 'basic_lstm_cell/bias'),
            self.p(d + 'output_projection/kernel'), max_len, p, sched['seed'], sched['step'], decoder, rows,
            replay=replay)

    def token_decoder_greedy(self, scope, h0, c0, max_len, vocab, end_id):
        table = self.p(scope + '/Token_Embedding/embedding_map')
        d = scope + '/dynamic_decoder/'
        return T.decode_greedy(
            lambda ids: T.embedding_lookup_gpu(table, ids), vocab, end_id,
            c0, h0, self.p(d + 'basic_lstm_cell/kernel'),
            self.p(d + 'basic_lstm_cell/bias'),
            self.p(d + 'output_projection/kernel'), max_len)

    # -- whole graphs -----------------------------------------------------------
    def forward(self, batch, greedy=False, sched=None):
        """sched (scheduled sampling): dict(step, seed[, p_override][, replay_program [B,L],
        replay_action [B,k,T]]); the tokens fed / draws taken come back in out['fed_*'],
        out['took_*'] and out['sched_mismatches']."""
        cfg, dt = self.cfg, self.dtype
        self.new_state = self.state.clone()   # moving stats advance once per forward
        s_h = torch.as_tensor(np.asarray(batch['s_h'])).to(dt)
        demo_len = torch.as_tensor(np.asarray(batch['demo_len'])).long()
        program_len = torch.as_tensor(np.asarray(batch['program_len'])).long()[:, 0]
        program = torch.as_tensor(np.asarray(batch['program'])).to(dt)
        ptoks = torch.as_tensor(np.asarray(batch['program_tokens'])).long()
        k = cfg.k
        out = {}
        # Demo -> demo feature (k copies, shared weights; model_full.py:373-379)
        hist1, h1, c1 = [], [], []
        for i in range(k):
            y, h, c = self.demo_encoder(s_h[:, i], demo_len[:, i])
            hist1.append(y), h1.append(h), c1.append(c)
        if cfg.model == 'synthesis_baseline':
            hs, cs = torch.stack(h1, 1), torch.stack(c1, 1)
            if cfg.demo_aggregation == 'avgpool':
                h_sum, c_sum = hs.mean(1), cs.mean(1)
            elif cfg.demo_aggregation == 'maxpool':
                h_sum, c_sum = hs.max(1).values, cs.max(1).values
            else:
                raise ValueError('Unknown demo aggregation type')
            demo_h, demo_c = h1, c1
        else:
            summary_h = torch.stack(h1, 1).mean(1)
            summary_c = torch.stack(c1, 1).mean(1)
            sc = 'SecondPathEncoder/rnn/basic_lstm_cell/'
            demo_h, demo_c = [], []
            for i in range(k):
                _, h, c = T.dynamic_rnn(hist1[i], demo_len[:, i],
                                        self.p(sc + 'kernel'), self.p(sc + 'bias'),
                                        c0=summary_c, h0=summary_h)
                demo_h.append(h), demo_c.append(c)
            hs, cs = torch.stack(demo_h, 1), torch.stack(demo_c, 1)
            h_sum = self.rn_pool(hs, 'demo_h_summary')
            c_sum = self.rn_pool(cs, 'demo_c_summary')
            if cfg.model == 'full':  # full: mean + rn (model_full.py:357-359)
                h_sum = hs.mean(1) + h_sum
                c_sum = cs.mean(1) + c_sum
        out['demo_h_summary'], out['demo_c_summary'] = h_sum, c_sum

        V, L = cfg.dim_program_token, cfg.max_program_len
        if sched is not None:
            out['sched_mismatches'] = []
            logits, fed, took, mm = self.token_decoder_scheduled(
                'Program_Decoder', h_sum, c_sum, ptoks, program_len, L, V, sched, 1,
                np.arange(ptoks.shape[0]), replay=sched.get('replay_program'))
            out['fed_program'], out['took_program'] = fed, took
            out['sched_mismatches'] += mm
        else:
            logits = self.token_decoder('Program_Decoder', h_sum, c_sum, ptoks,
                                        program_len, L, V)
        out['pred_program'] = logits.transpose(1, 2)           # [B,V,L]
        program_loss = T.softmax_ce_loss(logits, program.transpose(1, 2),
                                         program_len)
        out['program_loss'] = program_loss
        loss = program_loss
        if greedy:
            gl, glen, gtok = self.token_decoder_greedy(
                'Program_Decoder', h_sum, c_sum, L, V, cfg.program_end_token)
            out['greedy_pred_program'] = gl.transpose(1, 2)
            out['greedy_pred_program_len'] = glen.unsqueeze(1)
            out['greedy_program_tokens'] = gtok

        if cfg.model == 'full':
            A, Tm, P = cfg.action_space, cfg.max_demo_len, cfg.per_dim
            a_h = torch.as_tensor(np.asarray(batch['a_h'])).to(dt)
            a_tok = torch.as_tensor(np.asarray(batch['a_h_tokens'])).long()
            per = torch.as_tensor(np.asarray(batch['per'])).to(dt)
            act_loss, per_loss = 0, 0
            pa, pp, ga, galen = [], [], [], []
            for i in range(k):
                if sched is not None:
                    rp = sched.get('replay_action')
                    lg, fed, took, mm = self.token_decoder_scheduled(
                        'Action_Decoder', demo_h[i], demo_c[i], a_tok[:, i], demo_len[:, i], Tm, A, sched, 2,
                        np.arange(a_tok.shape[0]) * k + i, replay=None if rp is None else rp[:, i])
                    out.setdefault('fed_action', []).append(fed)
                    out.setdefault('took_action', []).append(took)
                    out['sched_mismatches'] += mm
                else:
                    lg = self.token_decoder('Action_Decoder', demo_h[i], demo_c[i],
                                            a_tok[:, i], demo_len[:, i], Tm, A)
                pa.append(lg)
                act_loss = act_loss + T.softmax_ce_loss(lg, a_h[:, i],
                                                        demo_len[:, i])
                if greedy:
                    g, gln, _ = self.token_decoder_greedy(
                        'Action_Decoder', demo_h[i], demo_c[i], Tm, A, A - 1)
                    ga.append(g), galen.append(gln)
            for i in range(k):
                # Per_Encoder: fc(act=None) + BN over [B,T] (model_full.py:308-316)
                sc = 'Per_Decoder/Per_Encoder/fc2'
                x = per[:, i] @ self.p(sc + '/fully_connected/weights') + \
                    self.p(sc + '/fully_connected/biases')
                x = self._bn(x, sc + '/bn_act')
                d = 'Per_Decoder/dynamic_decoder/'
                lg = T.decode_training(
                    x, demo_len[:, i], demo_c[i], demo_h[i],
                    self.p(d + 'basic_lstm_cell/kernel'),
                    self.p(d + 'basic_lstm_cell/bias'),
                    self.p(d + 'output_projection/kernel'), Tm)
                pp.append(lg)
                per_loss = per_loss + T.sigmoid_ce_loss(lg, per[:, i],
                                                        demo_len[:, i])
            out['avg_action_loss'] = act_loss / k
            out['avg_per_loss'] = per_loss / k
            out['pred_action'] = torch.stack(pa, 1)   # [B,k,T,A]
            out['pred_per'] = torch.stack(pp, 1)      # [B,k,T,P]
            if greedy:
                out['greedy_pred_action'] = torch.stack(ga, 1)
                out['greedy_pred_action_len'] = torch.stack(galen, 1)
            loss = loss + act_loss / k + per_loss / k
        out['loss'] = loss
        return out

    # -- induction baseline (inference path; reference model_induction.py:383-819) ------
    def forward_induction(self, batch, greedy=False):
        cfg, dt = self.cfg, self.dtype
        self.new_state = self.state.clone()
        s_h = torch.as_tensor(np.asarray(batch['s_h'])).to(dt)
        per = torch.as_tensor(np.asarray(batch['per'])).to(dt)
        demo_len = torch.as_tensor(np.asarray(batch['demo_len'])).long()
        t_len = torch.as_tensor(np.asarray(batch['test_demo_len'])).long()
        t_tok = torch.as_tensor(np.asarray(batch['test_a_h_tokens'])).long()
        t_oh = torch.as_tensor(np.asarray(batch['test_a_h'])).to(dt)
        k, tk, A, Tm = cfg.k, cfg.test_k, cfg.action_space, cfg.max_demo_len
        hist, hs, cs = [], [], []
        for i in range(k):
            y, h, c = self.demo_encoder(s_h[:, i], demo_len[:, i], per_i=per[:, i])
            hist.append(y), hs.append(h), cs.append(c)
        h_sum, c_sum = torch.stack(hs, 1).mean(1), torch.stack(cs, 1).mean(1)
        values = torch.stack(hist, 1)                                   # [B,k,T,H], zero past len
        keys = values @ self.p('AttnMechanism/memory_layer/kernel')
        w = 'Manipulation/dynamic_decoder/pooling_attention_wrapper/'
        kernel, bias = self.p(w + 'basic_lstm_cell/kernel'), self.p(w + 'basic_lstm_cell/bias')
        w_a = self.p(w + 'attention_layer/kernel')
        proj = self.p('Manipulation/dynamic_decoder/output_projection/kernel')
        table = self.p('Manipulation/Token_Embedding/embedding_map')
        B, H = h_sum.shape

        def step(x_emb, c, h, att):
            c, h = T.lstm_cell(torch.cat([x_emb, att], 1), c, h, kernel, bias)
            att = T.luong_pooled_attention(h, keys, values, demo_len, w_a)
            return c, h, att, att @ proj

        out = {'demo_h_summary': h_sum, 'demo_c_summary': c_sum}
        loss, logits_all, g_logits, g_lens = 0, [], [], []
        for j in range(tk):
            # swapped initial state: cell c := h-summary, cell h := c-summary
            c, h, att = h_sum, c_sum, h_sum.new_zeros(B, H)
            start = torch.full_like(t_tok[:, j, :1], A + 1)
            emb = T.embedding_lookup_gpu(table, torch.cat([start, t_tok[:, j, :-1]], 1))
            n_iter = max(int(min(int(t_len[:, j].max()), Tm)), 1)
            lg = []
            for t in range(n_iter):
                c, h, att, l = step(emb[:, t], c, h, att)
                lg.append(l)
            lg = torch.stack(lg, 1)
            if n_iter < Tm:
                lg = torch.cat([lg, lg.new_zeros(B, Tm - n_iter, A)], 1)
            logits_all.append(lg)
            loss = loss + T.softmax_ce_loss(lg, t_oh[:, j], t_len[:, j])
            if greedy:
                with torch.no_grad():
                    c, h, att = h_sum, c_sum, h_sum.new_zeros(B, H)
                    ids = torch.full((B,), A, dtype=torch.long)
                    fin = torch.zeros(B, dtype=torch.bool)
                    ln = torch.zeros(B, dtype=torch.long)
                    gl = []
                    for t in range(Tm):
                        c, h, att, l = step(T.embedding_lookup_gpu(table, ids), c, h, att)
                        ids = l.argmax(1)
                        nxt = fin | (ids == A - 1)
                        if t + 1 >= Tm:
                            nxt = torch.ones_like(nxt)
                        ln = torch.where(~fin & nxt, torch.full_like(ln, t + 1), ln)
                        fin = nxt
                        gl.append(l)
                        if bool(fin.all()):
                            break
                    gl = torch.stack(gl, 1)
                    if gl.shape[1] < Tm:
                        gl = torch.cat([gl, gl.new_zeros(B, Tm - gl.shape[1], A)], 1)
                    g_logits.append(gl), g_lens.append(ln)
        out['pred_action'] = torch.stack(logits_all, 1)                 # [B,test_k,T,A]
        out['avg_action_loss'] = out['loss'] = loss / tk
        if greedy:
            out['greedy_pred_action'] = torch.stack(g_logits, 1)
            out['greedy_pred_action_len'] = torch.stack(g_lens, 1)
        return out

    def loss_and_grad(self, batch, sched=None):
        """Returns (loss float, flat grad tensor, outputs)."""
        if self.flat.grad is not None:
            self.flat.grad = None
        if self.cfg.model == 'induction_baseline':
            out = self.forward_induction(batch)
        else:
            out = self.forward(batch, sched=sched)
        out['loss'].backward()
        return float(out['loss'].detach()), self.flat.grad.detach().clone(), out

    def commit_state(self):
        self.state = self.new_state.clone()


class OracleTrainer:
    """optimize_loss(Adam, clip_gradients=20.0) over the flat buffer
    (reference trainer.py:82-109; A.10)."""

    def __init__(self, cfg, flat_params, flat_state, dtype=torch.float64,
                 lr=1e-3, clip=20.0):
        self.model = OracleModel(cfg, flat_params, flat_state, dtype)
        self.m = torch.zeros_like(self.model.flat.detach())
        self.v = torch.zeros_like(self.m)
        self.step = 0
        self.lr, self.clip = lr, clip
        self.cfg = cfg

    def learning_rate(self):
        if self.cfg.lr_weight_decay:  # exponential_decay staircase (trainer.py:82-91)
            return self.lr * 0.5 ** (self.step // 10000)
        return self.lr

    def train_step(self, batch):
        loss, grad, out = self.model.loss_and_grad(batch)
        (grad,), norm = T.clip_by_global_norm([grad], self.clip)
        lr = self.learning_rate()
        self.step += 1
        with torch.no_grad():
            T.adam_step(self.model.flat, grad, self.m, self.v, self.step, lr=lr)
        self.model.commit_state()
        return loss, float(norm), out
