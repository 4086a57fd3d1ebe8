"""TEST INFRASTRUCTURE - CPU oracle, never imported by the product path.

Restatement (torch CPU ops, fp64 or fp32) of the TF-1.3 semantics the
reference's hot path relies on.  TF-1.3 itself (requirements.txt:1) is not in
/root/reference and cannot be installed offline, and the reference holds no
tests or golden vectors for this path: PARITY UNPINNED by the reference.  Each
function cites the reference call site and the TF-1.3 behaviour it restates
(SURVEY Appendix A).  A second, independent restatement in plain NumPy loops
lives in oracle/numpy_ref.py and the two are cross-checked in tests/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
may import this package.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3      # contrib.layers.batch_norm default epsilon (A.3)
BN_DECAY = 0.9     # reference models/ops.py:21


def lrelu(x, leak=0.2):
    """reference models/ops.py:7-11: f1*x + f2*abs(x); d/dx at 0 is f1 = 0.6."""
    f1 = 0.5 * (1 + leak)
    f2 = 0.5 * (1 - leak)
    return f1 * x + f2 * torch.abs(x)


def same_pad_3x3_s2(n):
    """TF SAME rule (A.1): out = ceil(n/2); pad_total = max((out-1)*2+3-n, 0);
    before = total // 2; after = total - before."""
    out = (n + 1) // 2
    total = max((out - 1) * 2 + 3 - n, 0)
    return out, total // 2, total - total // 2


def conv2d_3x3_s2_same(x, w, b):
    """slim.conv2d(x, C, [3,3], stride=2, activation_fn=None) at reference
    models/ops.py:30.  x NHWC, w HWIO, b [C]."""
    _, pt, pb = same_pad_3x3_s2(x.shape[1])
    _, pl, pr = same_pad_3x3_s2(x.shape[2])
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), b, stride=2)
    return y.permute(0, 2, 3, 1)


def batch_norm(x, gamma, beta, moving_mean, moving_var, is_training):
    """tf.contrib.layers.batch_norm(center, scale, decay=0.9, eps=1e-3,
    updates_collections=None) at reference models/ops.py:20-23 (A.3).
    Training: biased batch moments over all axes but the last; moving stats
    updated in line: moving -= (moving - batch) * (1 - decay).
    Returns (y, new_moving_mean, new_moving_var)."""
    if is_training:
        axes = tuple(range(x.dim() - 1))
        mean = x.mean(dim=axes)
        var = ((x - mean) ** 2).mean(dim=axes)
        y = (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta
        with torch.no_grad():
            nmm = moving_mean - (moving_mean - mean) * (1 - BN_DECAY)
            nmv = moving_var - (moving_var - var) * (1 - BN_DECAY)
        return y, nmm, nmv
    y = (x - moving_mean) * torch.rsqrt(moving_var + BN_EPS) * gamma + beta
    return y, moving_mean, moving_var


def lstm_cell(x, c, h, kernel, bias, forget_bias=1.0):
    """rnn.BasicLSTMCell (A.4): z = [x, h] @ kernel + bias; i, j, f, o = split;
    c' = c*sigmoid(f + 1) + sigmoid(i)*tanh(j); h' = tanh(c')*sigmoid(o)."""
    z = torch.cat([x, h], dim=1) @ kernel + bias
    i, j, f, o = torch.chunk(z, 4, dim=1)
    c2 = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
    h2 = torch.tanh(c2) * torch.sigmoid(o)
    return c2, h2


def dynamic_rnn(x, seq_len, kernel, bias, c0=None, h0=None):
    """tf.nn.dynamic_rnn(BasicLSTMCell, sequence_length=len) (A.5) at reference
    models/model_full.py:254-256,274-276.  x [R,T,In]; for t >= len[r] the
    output row is zero and the state is copied through.
    Returns (outputs [R,T,H], h_final, c_final)."""
    R, T, _ = x.shape
    H = kernel.shape[1] // 4
    c = x.new_zeros(R, H) if c0 is None else c0
    h = x.new_zeros(R, H) if h0 is None else h0
    outs = []
    for t in range(T):
        c2, h2 = lstm_cell(x[:, t], c, h, kernel, bias)
        live = (t < seq_len).unsqueeze(1)
        c = torch.where(live, c2, c)
        h = torch.where(live, h2, h)
        outs.append(torch.where(live, h2, torch.zeros_like(h2)))
    return torch.stack(outs, 1), h, c


def embedding_lookup_gpu(table, ids):
    """tf.nn.embedding_lookup with the TF *GPU* gather semantics the authors
    trained with: out-of-range ids yield a zero row (the teacher-forcing start
    id token_dim+1 is out of range for the [token_dim+1, E] table; reference
    models/model_full.py:288-291,448-450)."""
    ok = (ids >= 0) & (ids < table.shape[0])
    safe = torch.where(ok, ids, torch.zeros_like(ids))
    return table[safe] * ok.unsqueeze(-1).to(table.dtype)


def decode_training(inputs, seq_len, c0, h0, kernel, bias, proj, max_len):
    """dynamic_decode(BasicDecoder(cell, TrainingHelper(inputs, len), (c,h),
    Dense(V, no bias)), maximum_iterations=max_len), impute_finished=False
    (A.6) + the zero-pad to max_len at reference models/model_full.py:476-484.
    All rows keep stepping until every row is finished, i.e. for max_b len
    iterations; later positions are zero logits.
    inputs [R, max_len, E] -> logits [R, max_len, V]."""
    R = inputs.shape[0]
    n_iter = int(min(int(seq_len.max()), max_len)) if R else 0
    n_iter = max(n_iter, 1)  # the while_loop body always runs once
    c, h = c0, h0
    outs = []
    for t in range(n_iter):
        c, h = lstm_cell(inputs[:, t], c, h, kernel, bias)
        outs.append(h @ proj)
    logits = torch.stack(outs, 1)
    if n_iter < max_len:
        pad = logits.new_zeros(R, max_len - n_iter, logits.shape[2])
        logits = torch.cat([logits, pad], 1)
    return logits


def decode_greedy(embed_fn, start_id, end_id, c0, h0, kernel, bias, proj,
                  max_len):
    """dynamic_decode(BasicDecoder(cell, GreedyEmbeddingHelper(embed, start,
    end)), maximum_iterations=max_len) (A.6), reference
    models/model_full.py:424-435,513-521.  argmax ties -> lowest index;
    finished rows keep stepping on their own samples until all rows are
    finished; length = first finish step + 1, or max_len.
    Returns (logits [R,max_len,V] zero-padded, lengths [R], tokens [R,max_len])."""
    R = c0.shape[0]
    dev = c0.device
    ids = torch.full((R,), start_id, dtype=torch.long, device=dev)
    finished = torch.zeros(R, dtype=torch.bool, device=dev)
    lengths = torch.zeros(R, dtype=torch.long, device=dev)
    c, h = c0, h0
    outs, toks = [], []
    with torch.no_grad():
        for t in range(max_len):
            x = embed_fn(ids)
            c, h = lstm_cell(x, c, h, kernel, bias)
            logit = h @ proj
            ids = torch.argmax(logit, dim=1)
            nxt = finished | (ids == end_id)
            if t + 1 >= max_len:
                nxt = torch.ones_like(nxt)
            lengths = torch.where(~finished & nxt,
                                  torch.full_like(lengths, t + 1), lengths)
            finished = nxt
            outs.append(logit)
            toks.append(ids)
            if bool(finished.all()):
                break
    logits = torch.stack(outs, 1)
    tokens = torch.stack(toks, 1)
    n = logits.shape[1]
    if n < max_len:
        logits = torch.cat([logits, logits.new_zeros(R, max_len - n, logits.shape[2])], 1)
        tokens = torch.cat([tokens, tokens.new_zeros(R, max_len - n)], 1)
    return logits, lengths, tokens


# ---- scheduled sampling (reference models/model_full.py:59-67, 414-423) -----------------------
def _mix32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7feb352d) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846ca68b) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def sched_hash(seed, step, decoder, t, r, which):
    """The counter-based draw of csrc/sched_sample.cu (d2p_sched_hash), restated."""
    h = _mix32(seed + 0x9E3779B9)
    h = _mix32(h ^ step)
    h = _mix32(h + decoder)
    h = _mix32(h ^ t)
    h = _mix32(h + r)
    h = _mix32(h ^ which)
    return h


def scheduled_sampling_prob(step, decay_steps):
    """1 - tf.train.polynomial_decay(1.0, step, decay_steps, end_learning_rate=0.1, power=1.0):
    the probability of feeding a sampled token (ScheduledEmbeddingTrainingHelper's
    sampling_probability = 1 - sample_prob)."""
    frac = min(float(step), float(decay_steps)) / float(decay_steps)
    return 1.0 - ((1.0 - 0.1) * (1.0 - frac) + 0.1)


def decode_scheduled(embed_fn, start_id, gt_tokens, seq_len, c0, h0, kernel, bias, proj, max_len,
                     p, seed, step, decoder, rows, replay=None, tol=1e-4):
    """dynamic_decode(BasicDecoder(cell, ScheduledEmbeddingTrainingHelper(...))): like
    decode_training, but after step t row r feeds, if p > u1, a token drawn from
    Categorical(logits_t[r]) (inverse CDF at u2) to step t+1, else the ground-truth token.
    u1 / u2 = sched_hash(seed, step, decoder, t, rows[r], 0 / 1) / 2^32.
    replay [R, max_len] (the tokens another implementation fed): followed whenever this
    restatement's own draw differs; differences further than `tol` (of the total mass) from a CDF
    boundary are returned as mismatches.  Returns (logits, fed tokens, took mask, mismatches)."""
    import numpy as np
    R = c0.shape[0]
    n_iter = max(int(min(int(seq_len.max()), max_len)), 1)
    c, h = c0, h0
    ids = torch.full((R,), start_id, dtype=torch.long)
    fed = gt_tokens.clone().long()
    took = torch.zeros_like(fed)
    outs, mismatches = [], []
    for t in range(n_iter):
        c, h = lstm_cell(embed_fn(ids), c, h, kernel, bias)
        logit = h @ proj
        outs.append(logit)
        lg = logit.detach().double().numpy()
        for r in range(R):
            u1 = sched_hash(seed, step, decoder, t, int(rows[r]), 0) / 4294967296.0
            if not p > u1:
                continue
            took[r, t] = 1
            e = np.exp(lg[r] - lg[r].max())
            cdf = np.cumsum(e)
            u2 = float(np.float32(sched_hash(seed, step, decoder, t, int(rows[r]), 1) / 4294967296.0))
            target = u2 * cdf[-1]
            own = int(np.argmax(cdf > target)) if (cdf > target).any() else len(e) - 1
            tok = own
            if replay is not None and int(replay[r, t]) != own:
                tok = int(replay[r, t])
                lo, hi = min(own, tok), max(own, tok)
                # the draw sits within tol of every CDF boundary between the two answers
                if not all(abs(cdf[v] - target) <= tol * cdf[-1] for v in range(lo, hi)):
                    mismatches.append((decoder, t, r, own, tok))
            fed[r, t] = tok
        ids = fed[:, t]
    logits = torch.stack(outs, 1)
    if n_iter < max_len:
        logits = torch.cat([logits, logits.new_zeros(R, max_len - n_iter, logits.shape[2])], 1)
    return logits, fed, took, mismatches


def sequence_mask(lengths, max_len, dtype):
    return (torch.arange(max_len, device=lengths.device)[None, :] <
            lengths[:, None]).to(dtype)


def softmax_ce_loss(logits, onehot, gt_len):
    """Sequence_Loss for program/action (reference models/model_full.py:620-657,
    A.9): softmax_cross_entropy_with_logits(labels=onehot) per position, then
    sum(ce * gt_mask) / sum(gt_mask) over the whole batch.
    logits, onehot: [R, L, V]; gt_len [R]."""
    L = logits.shape[1]
    mask = sequence_mask(gt_len, L, logits.dtype)
    ce = -(onehot * F.log_softmax(logits, dim=-1)).sum(-1)
    return (ce * mask).sum() / mask.sum()


def sigmoid_ce_loss(logits, labels, gt_len):
    """Sequence_Loss for `per` (reference models/model_full.py:651-657):
    sigmoid_cross_entropy_with_logits = max(x,0) - x*z + log1p(exp(-|x|)),
    mean over per_dim, then the masked batch mean."""
    L = logits.shape[1]
    mask = sequence_mask(gt_len, L, logits.dtype)
    ce = torch.clamp(logits, min=0) - logits * labels + \
        torch.log1p(torch.exp(-torch.abs(logits)))
    ce = ce.mean(-1)
    return (ce * mask).sum() / mask.sum()


def fc_act_bn(x, w, b, gamma, beta, mm, mv, is_training, act=lrelu):
    """ops.fc (reference models/ops.py:149-155): slim.fully_connected ->
    activation -> BN."""
    y = x @ w + b
    if act is not None:
        y = act(y)
    return batch_norm(y, gamma, beta, mm, mv, is_training)


def clip_by_global_norm(grads, clip, extra_sq=0.0):
    """tf.clip_by_global_norm as used by optimize_loss(clip_gradients=20.0)
    (reference trainer.py:102-109, A.10): g * clip / max(norm, clip).
    `extra_sq` lets the caller restate TF's IndexedSlices norm quirk."""
    sq = sum((g.double() ** 2).sum() for g in grads) + extra_sq
    norm = torch.sqrt(sq)
    scale = clip / torch.clamp(norm, min=clip)
    return [g * scale.to(g.dtype) for g in grads], norm


def adam_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (A.10): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m / (sqrt(v) + eps)  (epsilon outside the bias correction).
    In-place on p, m, v; `step` is 1-based."""
    lr_t = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_(lr_t * m / (v.sqrt() + eps))


def luong_pooled_attention(h, keys, values, mem_len, w_a):
    """One PoolingAttentionWrapper step after the cell (reference
    models/baselines/model_induction.py:25-53, 156-169):
    for every memory i: score = h . keys_i^T (no scale), positions >= len -> -inf,
    softmax, context_i = alpha . values_i, attention_i = [h ; context_i] @ W_a (shared);
    output = mean_i attention_i.
    h [B,H]; keys/values [B,k,T,H]; mem_len [B,k]; w_a [2H,H]."""
    B, k, Tm, H = keys.shape
    score = torch.einsum('bh,bkth->bkt', h, keys)
    mask = torch.arange(Tm, device=h.device)[None, None, :] < mem_len[:, :, None]
    score = torch.where(mask, score, torch.full_like(score, float('-inf')))
    alpha = torch.softmax(score, dim=-1)
    ctx = torch.einsum('bkt,bkth->bkh', alpha, values)
    att = torch.cat([h.unsqueeze(1).expand(B, k, H), ctx], dim=-1) @ w_a     # [B,k,H]
    return att.mean(1)
