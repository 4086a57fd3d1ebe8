#!/usr/bin/env python
"""Evaluation driver with the reference's CLI (reference evaler.py:362-495):
restores a checkpoint, runs `max_steps` batches with is_train=False (BatchNorm on
moving statistics, evaler.py:61), aggregates report_loss / report_accuracy and
prints / writes the final report (evaler.py:292-359).  Program dumps
(`--pred_program`) write predicted and ground-truth token strings."""
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
        max_steps = self.config.max_steps
        loss_all, acc_all, time_all = [], [], 0.0
        vocab = None
        if self.config.pred_program and self.output_dir:
            from demo2program_b200.vocab import karel_vocab
            vocab = karel_vocab()
            os.makedirs(self.output_dir, exist_ok=True)
        for s in range(max_steps):
            step_time, loss, acc, hist, feed = self.run_single_step(self.batch)
            loss_all.append(loss); acc_all.append(acc); time_all += step_time
            if not self.config.quiet:
                self.log_step_message(s, loss, acc, hist, step_time)
            if vocab is not None and self.model.greedy_pred_program is not None:
                with open(os.path.join(self.output_dir, 'out_%d.txt' % s), 'w') as f:
                    pred = self.model.greedy_pred_program.argmax(1)
                    for b in range(pred.shape[0]):
                        n = int(self.model.greedy_pred_program_len[b, 0])
                        g = int(self.model.program_len[b, 0])
                        f.write('[pred] %s\n[gt]   %s\n' % (
                            vocab.intseq2str(pred[b, :n]),
                            vocab.intseq2str(np.asarray(feed['program_tokens'])[b, :g])))
        lk, ak = sorted(loss_all[0]), sorted(acc_all[0])
        avg_loss = [float(np.mean([l[k] for l in loss_all])) for k in lk]
        avg_acc = [float(np.nanmean([a[k] for a in acc_all])) if not all(
            np.isnan(a[k]) for a in acc_all) else float('nan') for k in ak]
        self.log_final_message(avg_loss, lk, avg_acc, ak, {}, [], time_all,
                               write_summary=self.config.write_summary,
                               summary_file=self.config.summary_file)

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
        msg = ("[{split_mode:5s} step {step:5d}] {loss_str}{acc_str}"
               "({sec_per_batch:.3f} sec/batch, {instance_per_sec:.3f} instances/sec)").format(
                   split_mode=(is_train and 'train' or 'val'), step=step, loss_str=loss_str,
                   acc_str=acc_str, sec_per_batch=step_time,
                   instance_per_sec=self.batch_size / step_time)
        log.info(msg)
        return msg

    def log_final_message(self, loss, loss_key, acc, acc_key, hist, hist_key, time_,
                          write_summary=False, summary_file=None, is_train=False):
        loss_str = "".join("{}:{loss: .3f} ".format(k, loss=v) for k, v in zip(loss_key, loss))
        acc_str = "".join("{}:{acc: .3f}\n".format(k, acc=v) for k, v in zip(acc_key, acc))
        msg = ("[Final Avg Report] \n[Loss] {}\n[Acc]  {}\n[Hist] \n[Time] ({:.3f} sec)").format(
            loss_str, acc_str[:-1], time_)
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
