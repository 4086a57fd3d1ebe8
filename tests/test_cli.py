"""CPU tests of the reference-facing surface: CLI flags, config plumbing, dataset
tuples, model dispatch and error conventions (no GPU needed)."""
import argparse

import numpy as np
import pytest

from demo2program_b200 import dataset
from demo2program_b200.config import D2PConfig, karel_config
from demo2program_b200.model import config_from_namespace, get_model_class, FEED_KEYS


def test_dataset_tuple_matches_reference_order_and_dtypes():
    tr, te, va = dataset.create_default_splits('synthetic:16', num_k=3)
    t = tr.get_data(tr.ids[0])
    assert len(t) == 13
    program, ptok, s_h, test_s_h, a_h, a_tok, ta_h, ta_tok, plen, dlen, tdlen, per, tper = t
    assert program.shape == (50, 50) and program.dtype == bool
    assert s_h.shape == (3, 20, 8, 8, 16) and test_s_h.shape[0] == 5
    assert a_h.shape == (3, 20, 6) and plen.shape == (1,) and plen.dtype == np.float32
    b = next(dataset.batches(tr, 4, shuffle=False))
    assert set(FEED_KEYS) <= set(b)
    assert b['s_h'].dtype == np.uint8 and b['program_tokens'].dtype == np.int32
    assert b['demo_len'].dtype == np.float32            # lengths travel as fp32 (input_ops_karel.py:73-74)


def test_batches_are_deterministic():
    tr, _, _ = dataset.create_default_splits('synthetic:16', num_k=2)
    a = next(dataset.batches(tr, 4, seed=3))
    b = next(dataset.batches(tr, 4, seed=3))
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_set_data_dims_and_config():
    import trainer
    tr, _, _ = dataset.create_default_splits('synthetic:8', num_k=4)
    ns = argparse.Namespace(dataset_type='karel', model='summarizer', batch_size=4, num_k=4)
    trainer.set_data_dims(ns, tr)
    cfg = config_from_namespace(ns)
    assert (cfg.k, cfg.test_k, cfg.max_demo_len, cfg.h, cfg.w, cfg.depth) == (4, 5, 20, 8, 8, 16)
    assert (cfg.dim_program_token, cfg.max_program_len, cfg.action_space, cfg.per_dim) == (50, 50, 6, 5)


def test_error_conventions():
    with pytest.raises(ValueError):
        get_model_class('nope')                       # reference trainer.py:29
    with pytest.raises(ValueError):
        D2PConfig(encoder_rnn_type='gru').validate()  # rnn/gru crash in the reference
    with pytest.raises(ValueError):
        D2PConfig(dataset_type='atari').validate()


def test_cli_flag_surface_matches_reference():
    import trainer, evaler
    # every flag of reference trainer.py:247-289 / evaler.py:366-424 parses
    with pytest.raises(SystemExit):
        trainer.main(['--help'])
    p = argparse.ArgumentParser()
    trainer.add_model_flags(p)
    ns = p.parse_args(['--encoder_rnn_type', 'lstm', '--num_lstm_cell_units', '256',
                       '--demo_aggregation', 'maxpool'])
    assert ns.num_lstm_cell_units == 256 and ns.attn_type == 'luong' and ns.pixel_input is False


def test_conv_geometry():
    g = karel_config().conv_geometry()
    assert [(x[0], x[3]) for x in g] == [(8, 4), (4, 2), (2, 1)]
