#!/usr/bin/env python
"""Training driver with the reference's CLI (reference trainer.py:243-344).

Same flags (--model/--dataset_type/--dataset_path/--checkpoint/--log_step/...),
same dataset-derived config fields (trainer.py:312-335), same log line
(trainer.py:227-240), same train_dir naming (trainer.py:37-53); the TF graph +
session is replaced by `demo2program_b200.model.Model` (libd2p on one B200, or one
process per GPU under torchrun with a single NCCL all-reduce of the gradients).
The train loop feeds `Engine.train_steps`, the pipelined public API `bench.py` times as `e2e`
(batch i+1 is staged and copied while step i runs; losses come back one step later), in runs that
end where a validation step or a checkpoint is due.
Deviations (documented in DESIGN.md): `--dataset_path synthetic[:N]` selects seeded
synthetic data; checkpoints are written both as TF-1.x tensor bundles (model-<step>.index/.data,
`checkpoint` state file) and as .npz keyed by TF variable name; TensorBoard summaries are not
produced.  Under torchrun every rank trains its own shard of the example ids; rank 0 alone owns
the train_dir, the log lines and the checkpoints.
"""
import argparse
import logging
import os
import time

import numpy as np

log = logging.getLogger('d2p')


class Trainer(object):
    @staticmethod
    def get_model_class(model_name):
        from demo2program_b200.model import get_model_class
        return get_model_class(model_name)

    def __init__(self, config, dataset, dataset_test):
        self.config = config
        hyper_parameter_str = 'bs_{}_lr_{}_{}_cell_{}'.format(
            config.batch_size, config.learning_rate, config.encoder_rnn_type,
            config.num_lstm_cell_units)
        if config.scheduled_sampling:
            hyper_parameter_str += '_sd_{}'.format(config.scheduled_sampling_decay_steps)
        hyper_parameter_str += '_k_{}'.format(config.num_k)
        self.train_dir = './train_dir/%s-%s-%s-%s-%s-%s' % (
            config.dataset_type, '_'.join(config.dataset_path.split('/')), config.model,
            config.prefix, hyper_parameter_str, time.strftime("%Y%m%d-%H%M%S"))
        self.is_chief = config.rank == 0
        if self.is_chief:       # one train_dir per job, not one per rank
            os.makedirs(self.train_dir, exist_ok=True)
            log.info("Train Dir: %s", self.train_dir)
        from demo2program_b200.dataset import batches
        self.batch_size = config.batch_size
        # data parallelism: rank r draws its batches from ids[r::world] (no example twice per epoch)
        self.batch_train = batches(dataset, self.batch_size, shuffle=True, seed=config.rank,
                                   workers=getattr(config, 'loader_workers', 0),
                                   rank=config.rank, world=config.world_size,
                                   copy=False)      # every batch is staged into pinned memory on receipt
        self.batch_test = batches(dataset_test, self.batch_size, shuffle=False)
        Model = self.get_model_class(config.model)
        log.info("Using Model class: %s", Model)
        self.model = Model(config, debug_information=config.debug, world_size=config.world_size,
                           device='cuda:%d' % config.local_rank)
        self.log_step = config.log_step
        self.test_sample_step = config.test_sample_step
        if config.checkpoint is not None:
            log.info("Checkpoint path: %s", config.checkpoint)
            # a TF checkpoint prefix (what the reference's Saver writes, e.g. .../model-5000) or a
            # .npz keyed by variable name; trainable variables only (pretrain_saver, trainer.py:115,145)
            from demo2program_b200 import tf_checkpoint
            if tf_checkpoint.is_tf_checkpoint(config.checkpoint):
                tf_checkpoint.load_model(config.checkpoint, self.model, trainable_only=True)
            else:
                self.model.load_state_dict(dict(np.load(config.checkpoint)), trainable_only=True)
            log.info("Loaded the pretrain parameters from the provided checkpoint path")

    def train(self, max_steps=1000000):
        log.info("Training Starts!")
        ckpt_save_step = 1000
        s = 0
        while s < max_steps:
            # a run of pipelined steps ends with the step after which a validation step or a
            # checkpoint is due (reference trainer.py:126-183 does both right after step s)
            due = lambda t: t % self.test_sample_step == 0 or t % ckpt_save_step == 0
            stop = s
            while not due(stop) and stop + 1 < max_steps:
                stop += 1
            for s, (step, loss, step_time) in zip(range(s, stop + 1), self.run_steps(self.batch_train, stop + 1 - s)):
                if s % self.log_step == 0 and self.is_chief:
                    self.log_step_message(step, loss, step_time)
            s = stop
            if s % self.test_sample_step == 0:
                step, test_loss, test_time = self.run_test(self.batch_test)
                if self.is_chief:
                    self.log_step_message(step, test_loss, test_time, is_train=False)
            if s % ckpt_save_step == 0 and self.is_chief:
                step = self.model.engine.step_count()
                log.info("Saved checkpoint at %d", s)
                np.savez(os.path.join(self.train_dir, 'model-%d.npz' % step), **self.model.state_dict())
                # and in the reference's own format (saver.save(..., 'model', global_step), trainer.py:182):
                # model-<step>.index / .data-00000-of-00001 + the `checkpoint` state file
                from demo2program_b200 import tf_checkpoint
                tf_checkpoint.save_model(os.path.join(self.train_dir, 'model-%d' % step), self.model)
            s += 1

    def run_steps(self, batch, n):
        """n train steps through the pipelined input path; yields (global step, loss, seconds)."""
        feeds = (self.model.get_feed_dict(next(batch), is_training=True) for _ in range(n))
        t0 = time.time()
        step0 = self.model.engine.step_count()
        for i, loss in enumerate(self.model.run_train_steps(feeds)):
            t1 = time.time()
            yield step0 + i + 1, loss, t1 - t0
            t0 = t1

    def run_single_step(self, batch, step=None, is_train=True):
        _start_time = time.time()
        batch_chunk = next(batch)
        feed = self.model.get_feed_dict(batch_chunk, step=step, is_training=is_train)
        loss = self.model.run_train_step(feed)
        return self.model.engine.step_count(), loss, time.time() - _start_time

    def run_test(self, batch):
        # like the reference (trainer.py:79-80) the trainer's model keeps batch statistics here
        _start_time = time.time()
        feed = self.model.get_feed_dict(next(batch), is_training=False)
        loss = self.model.run_eval_step(feed, greedy=False)
        return self.model.engine.step_count(), loss, time.time() - _start_time

    def log_step_message(self, step, loss, step_time, is_train=True):
        if step_time == 0:
            step_time = 0.001
        log.info((" [{split_mode:5s} step {step:4d}] " + "Loss: {loss:.5f} " +
                  "({sec_per_batch:.3f} sec/batch, {instance_per_sec:.3f} " + "instances/sec) "
                  ).format(split_mode=(is_train and 'train' or 'val'), step=step, loss=loss,
                           sec_per_batch=step_time, instance_per_sec=self.batch_size / step_time))


def add_model_flags(parser):
    parser.add_argument('--encoder_rnn_type', default='lstm', choices=['lstm', 'rnn', 'gru'])
    parser.add_argument('--num_lstm_cell_units', type=int, default=512)
    parser.add_argument('--demo_aggregation', type=str, default='avgpool',
                        choices=['concat', 'avgpool', 'maxpool'],
                        help='how to aggregate the demo features')
    # fields the reference's induction model reads but its CLI never defines
    # (models/baselines/model_induction.py:194-212) - added with defaults
    parser.add_argument('--pixel_input', action='store_true', default=False)
    parser.add_argument('--attn_type', type=str, default='luong')
    parser.add_argument('--state_encoder_fc', action='store_true', default=False)
    parser.add_argument('--concat_state_feature_direct_prediction', action='store_true', default=False)
    parser.add_argument('--stack_subsequent_state', action='store_true', default=False)


def set_data_dims(config, dataset_train):
    """reference trainer.py:305-335."""
    data_tuple = dataset_train.get_data(dataset_train.ids[0])
    program, _, s_h, test_s_h, a_h, _, _, _, program_len, demo_len, test_demo_len, per, test_per = \
        data_tuple[:13]
    config.dim_program_token = int(np.asarray(program.shape)[0])
    config.max_program_len = int(np.asarray(program.shape)[1])
    config.k = int(np.asarray(s_h.shape)[0])
    config.test_k = int(np.asarray(test_s_h.shape)[0])
    config.max_demo_len = int(np.asarray(s_h.shape)[1])
    config.h, config.w, config.depth = (int(x) for x in s_h.shape[2:5])
    config.action_space = int(np.asarray(a_h.shape)[2])
    config.per_dim = int(np.asarray(per.shape)[2])
    if config.dataset_type == 'karel':
        config.dsl_type = dataset_train.dsl_type
        config.env_type = dataset_train.env_type
        config.vizdoom_pos_keys = []
        config.vizdoom_max_init_pos_len = -1
        config.perception_type = ''
        config.level = None
    elif config.dataset_type == 'vizdoom':
        config.dsl_type = 'vizdoom_default'
        config.env_type = 'vizdoom_default'
        config.vizdoom_pos_keys = getattr(dataset_train, 'vizdoom_pos_keys', [])
        config.vizdoom_max_init_pos_len = getattr(dataset_train, 'vizdoom_max_init_pos_len', -1)
        config.perception_type = getattr(dataset_train, 'perception_type', '')
        config.level = getattr(dataset_train, 'level', None)
    else:
        raise ValueError(config.dataset_type)


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--debug', action='store_true', default=False)
    parser.add_argument('--prefix', type=str, default='default', help='a nickanme for the training')
    parser.add_argument('--model', type=str, default='full',
                        choices=['synthesis_baseline', 'induction_baseline', 'summarizer', 'full'])
    parser.add_argument('--dataset_type', type=str, default='karel', choices=['karel', 'vizdoom'])
    parser.add_argument('--dataset_path', type=str, default='datasets/karel_dataset')
    # not in the reference (it hard-codes 16 loader threads, input_ops_karel.py:118-123):
    parser.add_argument('--loader_workers', type=int, default=0,
                        help='forked loader processes assembling batches ahead of the step (0 = inline)')
    parser.add_argument('--checkpoint', type=str, default=None)
    parser.add_argument('--log_step', type=int, default=10)
    parser.add_argument('--write_summary_step', type=int, default=100)
    parser.add_argument('--test_sample_step', type=int, default=100)
    parser.add_argument('--num_k', type=int, default=10, help='the number of seen demonstrations')
    parser.add_argument('--batch_size', type=int, default=32)
    parser.add_argument('--learning_rate', type=float, default=0.001)
    parser.add_argument('--lr_weight_decay', action='store_true', default=False)
    parser.add_argument('--scheduled_sampling', action='store_true', default=False)
    parser.add_argument('--scheduled_sampling_decay_steps', type=int, default=20000)
    parser.add_argument('--max_steps', type=int, default=1000000,
                        help='(addition) stop after this many steps')
    add_model_flags(parser)
    config = parser.parse_args(argv)
    if config.scheduled_sampling and config.model == 'induction_baseline':
        raise ValueError('--scheduled_sampling: the step-by-step sampling decoder is built for the program / '
                         'action token decoders (full, summarizer, synthesis_baseline), not for the attention '
                         'decoder of induction_baseline')
    # one process per GPU (torchrun); single process otherwise
    config.rank = int(os.environ.get('RANK', '0'))
    config.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    config.world_size = int(os.environ.get('WORLD_SIZE', '1'))
    if config.world_size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(config.local_rank)
        from demo2program_b200.dp import init_nccl
        init_nccl('cuda:%d' % config.local_rank)
    from demo2program_b200 import dataset
    if config.dataset_type not in ('karel', 'vizdoom'):
        raise ValueError(config.dataset_type)
    dataset_train, dataset_test, dataset_val = dataset.create_default_splits(
        config.dataset_path, num_k=config.num_k, dataset_type=config.dataset_type)
    set_data_dims(config, dataset_train)
    trainer = Trainer(config, dataset_train, dataset_test)
    log.warning("dataset: %s, learning_rate: %f", config.dataset_path, config.learning_rate)
    trainer.train(config.max_steps)


if __name__ == '__main__':
    main()
