"""Report metrics of the reference's `Sequence_Loss` (models/model_full.py:620-700) for the facade:
masked cross-entropy, token accuracy and sequence accuracy of a token sequence, and their means over
the k (or test_k) demonstrations the reference reports as `avg_action_*` / `greedy_avg_action_*`
(models/model_full.py:1014-1060, models/baselines/model_induction.py:788-846).  These are read-outs
of logits the CUDA path produced; they run as a handful of torch reductions on the device."""
import torch


def sequence_stats(logits_bvl, gt_tokens, pred_len, gt_len):
    """logits [B,V,L]; gt_tokens [B,L] (int); pred_len, gt_len [B] (int).
    Returns dict(loss, token_acc, seq_acc, pred_tokens [B,L], is_same_seq [B] bool).
      loss      = sum(CE * gt_mask) / sum(gt_mask)
      token_acc = sum(equal * min_mask) / sum(max_mask)          (min / max of the two lengths)
      seq_acc   = mean(all(argmax equal under gt_mask) and pred_len == gt_len)"""
    B, V, L = logits_bvl.shape
    gt_tokens = gt_tokens.long()
    pred_len, gt_len = pred_len.long(), gt_len.long()
    ar = torch.arange(L, device=logits_bvl.device)[None]
    gt_mask = (ar < gt_len[:, None]).float()
    max_mask = (ar < torch.maximum(pred_len, gt_len)[:, None]).float()
    min_mask = (ar < torch.minimum(pred_len, gt_len)[:, None]).float()
    pred = logits_bvl.argmax(1)
    # labels are one-hot of the gt token inside the gt length and all-zero beyond it; the argmax
    # of an all-zero label row is 0 (tf.argmax), which only matters under gt_mask == 0
    gt_idx = gt_tokens.clamp(0, V - 1)
    lsm = torch.log_softmax(logits_bvl.float(), dim=1)
    ce = -lsm.gather(1, gt_idx[:, None, :])[:, 0, :]
    loss = float((ce.double() * gt_mask.double()).sum() / gt_mask.double().sum().clamp(min=1))
    eq = (pred == gt_tokens).float()
    token_acc = float((eq * min_mask).sum() / max_mask.sum().clamp(min=1))
    seq_eq = ((pred.float() * gt_mask) == (gt_tokens.float() * gt_mask)).all(1) & (pred_len == gt_len)
    return {'loss': loss, 'token_acc': token_acc, 'seq_acc': float(seq_eq.float().mean()),
            'pred_tokens': pred, 'is_same_seq': seq_eq}


def demo_sequence_stats(logits_bkta, gt_tokens_bkt, pred_len_bk, gt_len_bk):
    """Per-demonstration Sequence_Loss statistics averaged over the demonstrations.
    logits [B,k,T,A]; gt_tokens [B,k,T]; lengths [B,k].  Returns dict(loss, token_acc, seq_acc,
    seq_all_acc) - seq_all_acc: all k sequences of a batch element match (induction baseline)."""
    B, k, T, A = logits_bkta.shape
    acc = {'loss': 0.0, 'token_acc': 0.0, 'seq_acc': 0.0}
    same = []
    for i in range(k):
        st = sequence_stats(logits_bkta[:, i].permute(0, 2, 1), gt_tokens_bkt[:, i], pred_len_bk[:, i],
                            gt_len_bk[:, i])
        for key in acc:
            acc[key] += st[key] / k
        same.append(st['is_same_seq'])
    acc['seq_all_acc'] = float(torch.stack(same, 1).all(1).float().mean())
    return acc
