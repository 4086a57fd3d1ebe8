#!/usr/bin/env python
"""Evaluation driver with the reference's CLI (reference evaler.py:362-495):
restores a checkpoint, runs `max_steps` batches with is_train=False (BatchNorm on
moving statistics, evaler.py:61), aggregates report_loss / report_accuracy and
prints / writes the final report incl. the execution-accuracy histograms (evaler.py:292-359).
`--pred_program` writes out_<checkpoint>_<split>.{txt,hdf5,log} and `--result_data` the per-example
result file, like evaler.py:151-208 (HDF5 through hdf5_lite.write_hdf5).  `--id_list` and
`--unseen_test` are accepted and unused, exactly as in the reference (it never reads them)."""
import argparse
import glob
import logging
import os
import time

import numpy as np

from trainer import add_model_flags, set_data_dims

log = logging.getLogger('d2p')


class Evaler(object):
    @staticmethod
    def get_model_class(model_name):
        from demo2program_b200.model import get_model_class
        return get_model_class(model_name)

    def __init__(self, config, dataset):
        self.config = config
        self.train_dir = config.train_dir
        self.output_dir = getattr(config, 'output_dir', None)
        self.batch_size = config.batch_size
        from demo2program_b200.dataset import batches
        self.dataset = dataset
        self.batch = batches(dataset, self.batch_size, shuffle=False, epochs=None)
        Model = self.get_model_class(config.model)
        log.info("Using Model class: %s", Model)
        self.model = Model(config, is_train=False)
        self.checkpoint = config.checkpoint
        from demo2program_b200 import tf_checkpoint
        if self.checkpoint == '' and self.train_dir:
            # tf.train.latest_checkpoint(train_dir) (reference evaler.py:86), else the newest .npz
            self.checkpoint = tf_checkpoint.latest_checkpoint(self.train_dir) or ''
        if self.checkpoint == '' and self.train_dir:
            cands = sorted(glob.glob(os.path.join(self.train_dir, 'model-*.npz')),
                           key=lambda p: int(p.rsplit('-', 1)[1][:-4]))
            self.checkpoint = cands[-1] if cands else ''
        if self.checkpoint:
            if tf_checkpoint.is_tf_checkpoint(self.checkpoint):
                tf_checkpoint.load_model(self.checkpoint, self.model, restore_optimizer=False)
            else:
                self.model.load_state_dict(dict(np.load(self.checkpoint)))
            log.info("Loaded from checkpoint: %s", self.checkpoint)
        else:
            log.warning("No checkpoint given: evaluating the initial parameters")

    def eval_run(self):
        """reference evaler.py:96-248: per-batch report, optional dumps (`--pred_program`: text /
        HDF5 / log files named out_<checkpoint>_<split>.*; `--result_data`: one HDF5 group per
        example with the ground-truth and greedy programs and the demonstrations), final averages of
        losses, accuracies and execution histograms."""
        cfg = self.config
        max_steps = cfg.max_steps
        if max_steps <= 0:
            raise ValueError('nothing to evaluate: max_steps = len(split) // batch_size = 0 '
                             '(%d examples, batch_size %d)' % (len(self.dataset), self.batch_size))
        model_is_program = cfg.model != 'induction_baseline'
        vocab = text_file = log_file = None
        pred_tree, result_tree = {}, {}
        if cfg.pred_program and model_is_program:
            if not self.output_dir:
                raise ValueError('--pred_program needs --output_dir')
            if getattr(cfg, 'dataset_type', 'karel') != 'karel':
                raise ValueError('--pred_program: only the Karel vocabulary is available offline')
            from demo2program_b200.vocab import karel_vocab
            vocab = karel_vocab()
            os.makedirs(self.output_dir, exist_ok=True)
            base_name = os.path.join(self.output_dir, 'out_{}_{}'.format(
                os.path.basename(self.checkpoint) if self.checkpoint else 'init',
                getattr(cfg, 'dataset_split', 'test')))
            log.info("Output Dir: %s", self.output_dir)
            text_file = open('{}.txt'.format(base_name), 'w')
            log_file = open('{}.log'.format(base_name), 'w')
        final_msg = ''
        loss_all, acc_all, hist_all, time_all = [], [], {}, []
        loss = acc = {}
        for s in range(max_steps):
            step_time, loss, acc, hist, feed = self.run_single_step(self.batch)
            m = self.model
            step_msg = ''
            if not cfg.quiet:
                step_msg = self.log_step_message(s, loss, acc, hist, step_time)
            ids = [i.decode() if isinstance(i, bytes) else str(i) for i in np.asarray(feed['id']).reshape(-1)]
            if cfg.result_data and model_is_program:
                for i, pid in enumerate(ids):
                    if pid in result_tree:
                        print('Duplicates: {}'.format(pid))
                        continue
                    result_tree[pid] = {
                        'program': np.asarray(m.ground_truth_program[i]),
                        'pred_program': np.asarray(m.greedy_pred_program[i]),
                        'pred_program_len': np.int64(m.greedy_pred_program_len[i][0]),
                        # the reference re-reads the stored demonstrations (unpadded, all demos of
                        # the program); the batch holds the same frames padded to max_demo_len
                        's_h': self._stored(pid, 's_h', feed['s_h'][i]),
                        'test_s_h': self._stored(pid, 'test_s_h', feed['test_s_h'][i])}
            if vocab is not None:
                log_file.write('{}\n'.format(step_msg))
                correctness = ['wrong', 'correct']
                gt_tokens = np.asarray(feed['program_tokens'])
                for i, pid in enumerate(ids):
                    n, g = int(m.program_len[i, 0]), int(m.greedy_pred_program_len[i, 0])
                    pred_str = vocab.intseq2str(m.pred_program[i, :, :n].argmax(0))
                    greedy_str = vocab.intseq2str(m.greedy_pred_program[i, :, :g].argmax(0))
                    have = len(m.program_is_correct_syntax) > 0
                    psyn = int(m.program_is_correct_syntax[i]) if have else 0
                    gsyn = int(m.greedy_program_is_correct_syntax[i]) if have else 0
                    if pid not in pred_tree:
                        grp = {'program_prediction': pred_str, 'program_syntax': correctness[psyn],
                               'greedy_prediction': greedy_str, 'greedy_syntax': correctness[gsyn]}
                        if have:
                            grp.update({
                                'program_num_execution_correct': np.int64(m.program_num_execution_correct[i]),
                                'program_is_correct_execution': np.asarray(m.program_is_correct_execution[i]),
                                'greedy_num_execution_correct': np.int64(m.greedy_num_execution_correct[i]),
                                'greedy_is_correct_execution': np.asarray(m.greedy_is_correct_execution[i])})
                        pred_tree[pid] = grp
                    text_file.write('[id: {}]\ngt: {}\npred{}: {}\ngreedy{}: {}\n'.format(
                        pid, vocab.intseq2str(gt_tokens[i, :n]), '(error)' if psyn == 0 else '', pred_str,
                        '(error)' if gsyn == 0 else '', greedy_str))
            loss_all.append(loss); acc_all.append(acc); time_all.append(step_time)
            for hk, hv in hist.items():
                hist_all.setdefault(hk, []).append(np.asarray(hv))
        if not cfg.no_loss:
            lk, ak = sorted(loss), sorted(acc)
            avg_loss = [float(np.mean([l[k] for l in loss_all])) for k in lk]
            avg_acc = [float(np.nanmean([a[k] for a in acc_all])) if not all(
                np.isnan(a[k]) for a in acc_all) else float('nan') for k in ak]
            hist_avg = {hk: np.average(np.stack(hv), axis=0) for hk, hv in hist_all.items()}
            final_msg = self.log_final_message(avg_loss, lk, avg_acc, ak, hist_avg, sorted(hist_avg),
                                               float(np.sum(time_all)), write_summary=cfg.write_summary,
                                               summary_file=cfg.summary_file)
        from demo2program_b200.hdf5_lite import write_hdf5
        if cfg.result_data and model_is_program:
            write_hdf5(cfg.result_data_path, result_tree)
            log.info("Wrote evaluation results of %d programs: %s", len(result_tree), cfg.result_data_path)
        if vocab is not None:
            write_hdf5('{}.hdf5'.format(base_name), pred_tree)
            log_file.write('{}\n'.format(final_msg))
            log_file.write("Model class: {}\n".format(cfg.model))
            log_file.write("Checkpoint: {}\n".format(self.checkpoint))
            log_file.write("Dataset: {}\n".format(cfg.dataset_path))
            log_file.close()
            text_file.close()
        log.warning('Completed Evaluation.')

    def _stored(self, pid, key, fallback):
        """The example's demonstrations as stored in the dataset file (reference evaler.py:160-161
        copies data_file[id][key]); synthetic datasets have no file - the padded batch entry."""
        f = getattr(self.dataset, 'file', None) or getattr(self.dataset, 'data', None)
        try:
            return np.asarray(f[pid][key]) if f is not None else np.asarray(fallback)
        except Exception:
            return np.asarray(fallback)

    def run_single_step(self, batch):
        _start_time = time.time()
        feed = self.model.get_feed_dict(next(batch), is_training=False)
        self.model.run_eval_step(feed, greedy=True)
        return (time.time() - _start_time, self.model.report_loss, self.model.report_accuracy,
                self.model.report_hist, feed)

    def log_step_message(self, step, loss, acc, hist, step_time, is_train=False):
        if step_time == 0:
            step_time = 0.001
        loss_str = "".join("{}:{loss: .3f} ".format(k, loss=loss[k]) for k in sorted(loss))
        acc_str = "".join("{}:{acc: .3f} ".format(k, acc=acc[k]) for k in sorted(acc))
        hist_str = ""
        for k in sorted(hist):
            hist_str += "{}: [".format(k) + "".join("{acc: .3f}, ".format(acc=h) for h in hist[k]) + "] "
        msg = ("[{split_mode:5s} step {step:5d}] {loss_str}{acc_str}{hist_str}"
               "({sec_per_batch:.3f} sec/batch, {instance_per_sec:.3f} instances/sec)").format(
                   split_mode=(is_train and 'train' or 'val'), step=step, loss_str=loss_str,
                   acc_str=acc_str, hist_str=hist_str, sec_per_batch=step_time,
                   instance_per_sec=self.batch_size / step_time)
        log.info(msg)
        return msg

    def log_final_message(self, loss, loss_key, acc, acc_key, hist, hist_key, time_,
                          write_summary=False, summary_file=None, is_train=False):
        loss_str = "".join("{}:{loss: .3f} ".format(k, loss=v) for k, v in sorted(zip(loss_key, loss)))
        acc_str = "".join("{}:{acc: .3f}\n".format(k, acc=v) for k, v in sorted(zip(acc_key, acc)))
        hist_str = ""
        for key in sorted(hist_key):
            hist_str += "{}: [".format(key) + "".join("{acc: .3f}, ".format(acc=h) for h in hist[key]) + "]\n"
        msg = ("[Final Avg Report] \n[Loss] {}\n[Acc]  {}\n[Hist] {}\n[Time] ({:.3f} sec)").format(
            loss_str, acc_str[:-1], hist_str[:-1], time_)
        log.info(msg)
        log.info("Model class: %s", self.config.model)
        log.info("Checkpoint: %s", self.checkpoint)
        log.info("Dataset: %s", self.config.dataset_path)
        if write_summary:
            with open(summary_file, 'w') as f:
                f.write('Model class: {}\nCheckpoint: {}\nDataset: {}\n{}'.format(
                    self.config.model, self.checkpoint, self.config.dataset_path, msg))
        return msg


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--model', type=str, default='full',
                        choices=['synthesis_baseline', 'induction_baseline', 'summarizer', 'full'])
    parser.add_argument('--dataset_type', type=str, default='karel', choices=['karel', 'vizdoom'])
    parser.add_argument('--dataset_path', type=str, default='datasets/karel_dataset')
    parser.add_argument('--dataset_split', type=str, default='test', choices=['train', 'test', 'val'])
    parser.add_argument('--checkpoint', type=str, default='')
    parser.add_argument('--train_dir', type=str, default='')
    parser.add_argument('--output_dir', type=str, default=None)
    parser.add_argument('--max_steps', type=int, default=0)
    parser.add_argument('--num_k', type=int, default=10)
    parser.add_argument('--batch_size', type=int, default=20)
    add_model_flags(parser)
    parser.add_argument('--no_loss', action='store_true', default=False)
    parser.add_argument('--pred_program', action='store_true', default=False)
    parser.add_argument('--result_data', action='store_true', default=False)
    parser.add_argument('--result_data_path', type=str, default='result.hdf5')
    parser.add_argument('--id_list', type=str)
    parser.add_argument('--unseen_test', action='store_true', default=False)
    parser.add_argument('--quiet', action='store_true', default=False)
    parser.add_argument('--no_write_summary', action='store_true', default=False)
    parser.add_argument('--summary_file', type=str, default='report.txt')
    config = parser.parse_args(argv)
    config.write_summary = not config.no_write_summary
    from demo2program_b200 import dataset
    dataset_train, dataset_test, dataset_val = dataset.create_default_splits(
        config.dataset_path, num_k=config.num_k, is_train=False,
        dataset_type=getattr(config, "dataset_type", "karel"))
    ds = {'train': dataset_train, 'test': dataset_test, 'val': dataset_val}[config.dataset_split]
    if config.max_steps == 0:
        config.max_steps = int(len(ds) / config.batch_size)   # reference evaler.py:448-449
    set_data_dims(config, ds)
    config.learning_rate, config.lr_weight_decay = 0.001, False
    config.scheduled_sampling = False
    evaler = Evaler(config, ds)
    log.warning("dataset: %s", config.dataset_path)
    evaler.eval_run()


if __name__ == '__main__':
    main()
