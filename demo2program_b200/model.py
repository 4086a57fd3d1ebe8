"""Reference-facing `Model` facade (the drop-in seam).

Mirrors the contract `trainer.py` / `evaler.py` consume from
`models/model_full.py:Model` and the three baselines:
`get_model_class(name)`, `Model(config, debug_information, is_train,
global_step)`, `get_feed_dict(batch_chunk, step, is_training)`, and the
attributes `.loss .output .report_loss .report_accuracy .report_hist
.pred_program .program_len .greedy_pred_program .greedy_pred_program_len
.ground_truth_program`.  There is no TF session: `run_train_step(feed)` /
`run_eval_step(feed)` take the feed dict and refresh those attributes.

Metrics that need the DSL parser / interpreter on the host (syntax, exact-program and
execution accuracies, reference models/model_full.py:602-616, 713-916) are computed for
Karel by the native library (`karel_dsl.py` -> csrc/karel_dsl.cu); the ViZDoom
ones need the game engine and are reported as NaN / empty.
"""
import numpy as np
import torch

from .config import D2PConfig, MODEL_NAMES
from .engine import Engine
from .metrics import demo_sequence_stats, sequence_stats

FEED_KEYS = ('id', 'program', 'program_tokens', 's_h', 'a_h', 'a_h_tokens', 'program_len',
             'demo_len', 'test_s_h', 'test_demo_len', 'per', 'test_per')


def get_model_class(model_name):
    """reference trainer.py:18-30 / evaler.py:18-30."""
    if model_name not in MODEL_NAMES:
        raise ValueError(model_name)
    if model_name == 'induction_baseline':
        from .induction import InductionModel
        return InductionModel
    return Model


def config_from_namespace(ns):
    """argparse namespace (+ dataset-derived dims, reference trainer.py:312-335) -> D2PConfig."""
    kw = {}
    for f in D2PConfig.__dataclass_fields__:
        if hasattr(ns, f) and getattr(ns, f) is not None:
            kw[f] = getattr(ns, f)
    return D2PConfig(**kw).validate()


def _seq_stats(logits_bvl, gt_tokens, pred_len, gt_len):
    """token / sequence accuracy of Sequence_Loss (reference models/model_full.py:660-683).
    logits [B,V,L]; gt_tokens [B,L]; lengths [B]."""
    st = sequence_stats(logits_bvl, gt_tokens, pred_len, gt_len)
    return st['token_acc'], st['seq_acc'], st['pred_tokens'], st['is_same_seq']


def _acc_hist(num_correct, k):
    """CompareDemoAndExecution's histogram (reference models/model_full.py:889-895)."""
    return np.array([float(np.mean(num_correct == i)) for i in range(k + 1)], np.float32)


class Model(object):
    def __init__(self, config, debug_information=False, is_train=True, global_step=None, **engine_kw):
        self.config = config if isinstance(config, D2PConfig) else config_from_namespace(config)
        self.debug = debug_information
        self.global_step = global_step
        self.is_train = is_train
        self.engine = Engine(self.config, is_train=is_train, **engine_kw)
        self.batch_size = self.config.batch_size
        self.loss = None
        self.output = []
        self.report_loss, self.report_accuracy, self.report_hist = {}, {}, {}
        self.pred_program = self.program_len = self.ground_truth_program = None
        self.greedy_pred_program = self.greedy_pred_program_len = None
        # host-interpreter metrics: filled by run_eval_step for Karel, empty otherwise
        self.program_is_correct_syntax = self.greedy_program_is_correct_syntax = []
        self.program_num_execution_correct = self.program_is_correct_execution = []
        self.greedy_num_execution_correct = self.greedy_is_correct_execution = []

    def get_feed_dict(self, batch_chunk, step=None, is_training=True):
        """reference models/model_full.py:185-206 (same keys; lengths arrive as fp32)."""
        missing = [k for k in FEED_KEYS if k not in batch_chunk]
        if missing:
            raise KeyError('batch_chunk is missing %s' % missing)
        return batch_chunk

    # -- steps ------------------------------------------------------------------------
    def run_train_step(self, feed):
        self.loss = self.engine.train_step(feed)
        return self.loss

    def run_train_steps(self, feeds):
        """Training loop over an iterable of feed dicts through the engine's pipelined input path
        (Engine.train_steps): yields every step's loss, one step behind the device."""
        for loss in self.engine.train_steps(feeds):
            self.loss = loss
            yield loss

    def _common_outputs(self, feed, greedy):
        eng, cfg = self.engine, self.config
        dev = eng.dev
        self.ground_truth_program = np.asarray(feed['program'])
        self.program_len = np.asarray(feed['program_len']).astype(np.int32)
        gt_tok = torch.as_tensor(np.asarray(feed['program_tokens'])).long().to(dev)
        gt_len = torch.as_tensor(self.program_len[:, 0]).long().to(dev)
        pp = eng.pred_program()
        tacc, sacc, ptok, psame = _seq_stats(pp, gt_tok, gt_len, gt_len)
        losses = eng.loss.cpu().numpy()
        self.pred_program = pp.cpu().numpy()
        self.loss = float(losses[0])
        self.report_loss = {'program_loss': float(losses[1])}
        self.report_accuracy = {'program_token_acc': tacc, 'program_seq_acc': sacc,
                                'program_syntax_acc': float('nan'),
                                'pred_exact_program_accuracy': float('nan')}
        if cfg.model == 'full':
            # teacher-forced action decoders (models/model_full.py:1014-1036): statistics per
            # demonstration, averaged over the k demonstrations
            B, k, T, A = eng.B, eng.k, eng.T, cfg.action_space
            a_tok = torch.as_tensor(np.asarray(feed['a_h_tokens'])).long().to(dev).view(B, k, T)
            a_len = torch.as_tensor(np.asarray(feed['demo_len'])).long().to(dev).view(B, k)
            pa = eng.act['logits'].view(T, B, k, A).permute(1, 2, 0, 3)
            ast = demo_sequence_stats(pa, a_tok, a_len, a_len)
            self.pred_action = pa.contiguous().cpu().numpy()
            self.report_loss['avg_action_loss'] = float(losses[2])
            self.report_accuracy['avg_action_token_acc'] = ast['token_acc']
            self.report_accuracy['avg_action_seq_acc'] = ast['seq_acc']
        if greedy:
            gp, glen, _ = eng.greedy_program()
            gst = sequence_stats(gp, gt_tok, glen[:, 0].long(), gt_len)
            gtok, gsame = gst['pred_tokens'], gst['is_same_seq']
            self.greedy_pred_program = gp.cpu().numpy()
            self.greedy_pred_program_len = glen.cpu().numpy()
            self.report_loss['greedy_program_loss'] = gst['loss']
            self.report_accuracy.update({'greedy_program_token_acc': gst['token_acc'],
                                         'greedy_program_seq_acc': gst['seq_acc'],
                                         'greedy_program_syntax_acc': float('nan'),
                                         'greedy_exact_program_accuracy': float('nan')})
            if cfg.model == 'full':     # models/model_full.py:1038-1058
                ga, galen = eng.greedy_actions()
                gast = demo_sequence_stats(ga, a_tok, galen.long(), a_len)
                self.greedy_pred_action = ga.cpu().numpy()
                self.greedy_pred_action_len = galen.cpu().numpy()
                self.report_loss['greedy_avg_action_loss'] = gast['loss']
                self.report_accuracy['greedy_avg_action_token_acc'] = gast['token_acc']
                self.report_accuracy['greedy_avg_action_seq_acc'] = gast['seq_acc']
        self.report_hist = {}
        self.output = [self.ground_truth_program, self.pred_program]
        if cfg.dataset_type == 'karel':
            self._karel_metrics(feed, 'program', 'pred', ptok, self.program_len[:, 0], psame)
            if greedy:
                self._karel_metrics(feed, 'greedy', 'greedy', gtok, self.greedy_pred_program_len[:, 0], gsame)

    def _karel_metrics(self, feed, name, exact_name, pred_tokens, pred_len, is_same_seq):
        """Syntax / exact-program / execution accuracies of one decoded program batch
        (reference models/model_full.py:931-1013: the teacher-forced pass uses the
        ground-truth lengths, the greedy pass its own)."""
        from . import karel_dsl
        cfg = self.config
        tok = pred_tokens.cpu().numpy().astype(np.int32)
        same = is_same_seq.cpu().numpy().astype(np.uint8)
        plen = np.asarray(pred_len, np.int32)
        gt_tok = np.asarray(feed['program_tokens'], np.int32)
        gt_len = self.program_len[:, 0]
        make_error = cfg.env_type != 'no_error'
        prefix = 'program' if name == 'program' else 'greedy'
        for split, key, lkey in (('', 's_h', 'demo_len'), ('test_', 'test_s_h', 'test_demo_len')):
            syn, exe, num = karel_dsl.eval_batch(tok, plen, same, np.asarray(feed[key]) != 0,
                                                 np.asarray(feed[lkey]).astype(np.int32), make_error)
            k = exe.shape[1]
            if name == 'program':
                setattr(self, split + 'program_num_execution_correct', num)
                setattr(self, split + 'program_is_correct_execution', exe.astype(bool))
            else:
                setattr(self, split + 'greedy_num_execution_correct', num)
                setattr(self, split + 'greedy_is_correct_execution', exe.astype(bool))
            self.report_hist[split + ('program' if name == 'program' else 'greedy_program') +
                             '_execution_acc_hist'] = _acc_hist(num, k)
        setattr(self, prefix + '_program_is_correct_syntax' if name == 'greedy' else 'program_is_correct_syntax', syn)
        exact = np.array([float(syn[b] == 1 and karel_dsl.programs_equal(tok[b, :plen[b]], gt_tok[b, :gt_len[b]]) == 1)
                          for b in range(tok.shape[0])], np.float32)
        setattr(self, exact_name + '_exact_program_correct', exact)
        self.report_accuracy[prefix + '_syntax_acc' if name == 'program' else 'greedy_program_syntax_acc'] = \
            float(syn.mean())
        self.report_accuracy[exact_name + '_exact_program_accuracy'] = float(exact.mean())

    def run_eval_step(self, feed, greedy=True):
        """Forward only (evaler.py:253-280): BN uses moving statistics when the model
        was built with is_train=False (reference evaler.py:61)."""
        eng = self.engine
        eng.stage_batch(feed)
        eng.forward()
        torch.cuda.current_stream(eng.dev).synchronize()
        self._common_outputs(feed, greedy)
        return self.loss

    # -- checkpoints (flat buffers keyed by TF variable name) ---------------------------
    def state_dict(self):
        eng = self.engine
        p, s = eng.params.cpu().numpy(), eng.state.cpu().numpy()
        out = {e.name: p[e.offset:e.offset + e.size].reshape(e.shape).copy() for e in eng.pm}
        out.update({e.name: s[e.offset:e.offset + e.size].reshape(e.shape).copy() for e in eng.sm})
        out['global_step'] = np.array(eng.step_count())
        return out

    def load_state_dict(self, d, trainable_only=False):
        """trainable_only=True restates the reference's --checkpoint warm start
        (pretrain_saver over trainable vars only, trainer.py:98,115,142-147)."""
        eng = self.engine
        p, s = eng.params.cpu().numpy(), eng.state.cpu().numpy()
        for e in eng.pm:
            if e.name in d:
                p[e.offset:e.offset + e.size] = np.asarray(d[e.name], np.float32).reshape(-1)
        if not trainable_only:
            for e in eng.sm:
                if e.name in d:
                    s[e.offset:e.offset + e.size] = np.asarray(d[e.name], np.float32).reshape(-1)
        eng.params.copy_(torch.from_numpy(p))
        eng.state.copy_(torch.from_numpy(s))
