"""Host-side metric code of the Model facade against plain-NumPy restatements of the reference
definitions (runs on CPU; uses the GPU when there is one)."""
import numpy as np
import torch


def _reference_seq_stats(logits_bvl, gt_onehot_bvl, pred_len, gt_len):
    """Plain-NumPy restatement of the accuracy part of Sequence_Loss (reference
    models/model_full.py:626-683): label/logit argmax over the token axis, token accuracy =
    sum(equal * min_mask) / sum(max_mask), sequence equality on gt_mask-ed argmaxes AND equal
    lengths."""
    B, V, L = logits_bvl.shape
    labels = np.transpose(gt_onehot_bvl, (0, 2, 1)).reshape(B * L, V)
    logits = np.transpose(logits_bvl, (0, 2, 1)).reshape(B * L, V)
    ar = np.arange(L)[None]
    gt_mask = (ar < gt_len[:, None]).astype(np.float32).reshape(-1)
    max_mask = (ar < np.maximum(pred_len, gt_len)[:, None]).astype(np.float32).reshape(-1)
    min_mask = (ar < np.minimum(pred_len, gt_len)[:, None]).astype(np.float32).reshape(-1)
    la, lo = labels.argmax(-1), logits.argmax(-1)
    token_acc = ((la == lo).astype(np.float32) * min_mask).sum() / max_mask.sum()
    seq_equal = ((la.astype(np.float32) * gt_mask).reshape(B, -1) ==
                 (lo.astype(np.float32) * gt_mask).reshape(B, -1)).all(-1)
    same = np.logical_and(seq_equal, gt_len == pred_len).astype(np.float32)
    return float(token_acc), float(same.sum() / B), lo.reshape(B, L), same


def test_seq_stats_match_sequence_loss_definition():
    from demo2program_b200.model import _seq_stats
    rs = np.random.RandomState(4)
    B, V, L = 16, 50, 50
    for trial in range(4):
        gt_len = rs.randint(3, L + 1, size=B)
        pred_len = gt_len.copy() if trial == 0 else rs.randint(3, L + 1, size=B)
        gt_tok = rs.randint(0, V, size=(B, L)) * (np.arange(L)[None] < gt_len[:, None])
        onehot = np.zeros((B, V, L), np.float32)
        for b in range(B):
            onehot[b, gt_tok[b, :gt_len[b]], np.arange(gt_len[b])] = 1.0
        logits = rs.randn(B, V, L).astype(np.float32)
        # make most predictions right so that equal / unequal sequences both occur
        boost = rs.rand(B, L) < 0.97
        for b in range(B):
            for t in range(L):
                if boost[b, t]:
                    logits[b, gt_tok[b, t], t] += 20.0
        logits[0] = onehot[0] * 30.0            # one exactly right row
        if trial:
            pred_len[0] = gt_len[0]
        ta, sa, tok, same = _reference_seq_stats(logits, onehot, pred_len, gt_len)
        dev = 'cuda' if torch.cuda.is_available() else 'cpu'
        ta2, sa2, tok2, same2 = _seq_stats(torch.from_numpy(logits).to(dev), torch.from_numpy(gt_tok).long().to(dev),
                                           torch.from_numpy(pred_len).long().to(dev),
                                           torch.from_numpy(gt_len).long().to(dev))
        assert abs(ta - ta2) < 1e-6 and abs(sa - sa2) < 1e-6
        assert np.array_equal(tok, tok2.cpu().numpy())
        assert np.array_equal(same.astype(bool), same2.cpu().numpy())


def _reference_sequence_loss(logits_bvl, gt_onehot_bvl, gt_len):
    """Loss part of Sequence_Loss (reference models/model_full.py:639-655): softmax cross-entropy
    with the one-hot labels over [B*L] rows, masked by the gt mask, over the mask's sum."""
    B, V, L = logits_bvl.shape
    labels = np.transpose(gt_onehot_bvl, (0, 2, 1)).reshape(B * L, V).astype(np.float64)
    logits = np.transpose(logits_bvl, (0, 2, 1)).reshape(B * L, V).astype(np.float64)
    lse = np.log(np.exp(logits - logits.max(-1, keepdims=True)).sum(-1)) + logits.max(-1)
    ce = -(labels * (logits - lse[:, None])).sum(-1)
    mask = (np.arange(L)[None] < gt_len[:, None]).astype(np.float64).reshape(-1)
    return float((ce * mask).sum() / mask.sum())


def test_sequence_loss_and_demo_averages():
    from demo2program_b200.metrics import demo_sequence_stats, sequence_stats
    rs = np.random.RandomState(8)
    B, k, T, A = 7, 3, 12, 6
    gt_len = rs.randint(2, T + 1, size=(B, k))
    pred_len = rs.randint(2, T + 1, size=(B, k))
    pred_len[:, 0] = gt_len[:, 0]
    gt_tok = rs.randint(0, A, size=(B, k, T)) * (np.arange(T)[None, None] < gt_len[..., None])
    logits = rs.randn(B, k, T, A).astype(np.float32)
    for b in range(B):
        for t in range(T):
            logits[b, 0, t, gt_tok[b, 0, t]] += 25.0       # demonstration 0 is decoded exactly
    logits[0, 1] = -1.0
    logits[0, 1, np.arange(T), gt_tok[0, 1]] = 9.0
    pred_len[0, 1:] = gt_len[0, 1:]
    logits[0, 2] = -1.0
    logits[0, 2, np.arange(T), gt_tok[0, 2]] = 9.0         # batch element 0: all demonstrations right
    loss = tacc = sacc = 0.0
    same = []
    for i in range(k):
        onehot = np.zeros((B, A, T), np.float32)
        for b in range(B):
            onehot[b, gt_tok[b, i, :gt_len[b, i]], np.arange(gt_len[b, i])] = 1.0
        lg = np.transpose(logits[:, i], (0, 2, 1))
        ta, sa, _, sm = _reference_seq_stats(lg, onehot, pred_len[:, i], gt_len[:, i])
        loss += _reference_sequence_loss(lg, onehot, gt_len[:, i]) / k
        tacc += ta / k
        sacc += sa / k
        same.append(sm.astype(bool))
        one = sequence_stats(torch.from_numpy(lg), torch.from_numpy(gt_tok[:, i]), torch.from_numpy(pred_len[:, i]),
                             torch.from_numpy(gt_len[:, i]))
        assert abs(one['loss'] - _reference_sequence_loss(lg, onehot, gt_len[:, i])) < 1e-5
    got = demo_sequence_stats(torch.from_numpy(logits), torch.from_numpy(gt_tok), torch.from_numpy(pred_len),
                              torch.from_numpy(gt_len))
    assert abs(got['loss'] - loss) < 1e-5 and abs(got['token_acc'] - tacc) < 1e-6
    assert abs(got['seq_acc'] - sacc) < 1e-6
    all_same = np.stack(same, 1).all(1)
    assert all_same[0] and abs(got['seq_all_acc'] - all_same.mean()) < 1e-6
