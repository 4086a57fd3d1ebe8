"""Scheduled sampling (reference models/model_full.py:59-67, 414-423; trainer.py:278-281).

CPU part: the counter-based draws of csrc/sched_sample.cu against the oracle's restatement, the
polynomial_decay schedule, and the oracle's scheduled decoder at p = 0 against its teacher-forced one.
GPU part: the step-by-step CUDA decoders against the oracle - the Bernoulli draws bit for bit, the
sampled tokens up to CDF-boundary ties, loss and gradients over the tokens actually fed."""
import numpy as np
import pytest
import torch

from demo2program_b200.config import karel_config


def test_draw_hash_matches_the_library(lib):
    from oracle import tf_ops as T
    rs = np.random.RandomState(0)
    for _ in range(3000):
        a = [int(x) for x in rs.randint(0, 2 ** 31, size=6)]
        a[5] %= 2
        assert lib.d2p_sched_hash(*a) == T.sched_hash(*a), a
    u = np.array([T.sched_hash(7, 3, 1, t, r, 0) / 2 ** 32 for t in range(50) for r in range(64)])
    assert abs(u.mean() - 0.5) < 0.02 and abs((u < 0.25).mean() - 0.25) < 0.03


def test_schedule_is_the_reference_polynomial_decay():
    from oracle import tf_ops as T
    assert T.scheduled_sampling_prob(0, 20000) == 0.0                       # pure teacher forcing at step 0
    assert abs(T.scheduled_sampling_prob(10000, 20000) - 0.45) < 1e-12
    assert abs(T.scheduled_sampling_prob(20000, 20000) - 0.9) < 1e-12       # final teacher-forcing prob 0.1
    assert abs(T.scheduled_sampling_prob(10 ** 6, 20000) - 0.9) < 1e-12


def test_oracle_scheduled_decoder_at_p0_is_teacher_forcing():
    from oracle.models import OracleModel
    from demo2program_b200.manifest import build_manifests
    from demo2program_b200.synthetic import make_batch
    cfg = karel_config('full', batch_size=3, k=2, num_lstm_cell_units=16, max_program_len=9, max_demo_len=6)
    pm, sm = build_manifests(cfg)
    m = OracleModel(cfg, pm.init_flat(0), sm.init_flat(0))
    batch = make_batch(cfg, seed=4, min_demo_len=3, min_prog_len=5)
    with torch.no_grad():
        a = m.forward(batch)
        b = m.forward(batch, sched=dict(step=0, seed=5, p_override=0.0))
        c = m.forward(batch, sched=dict(step=0, seed=5, p_override=1.0))
    assert float(a['loss']) == float(b['loss']) and torch.equal(a['pred_program'], b['pred_program'])
    assert int(b['took_program'].sum()) == 0
    n_it = int(np.asarray(batch['program_len']).max())
    assert int(c['took_program'][:, :n_it].sum()) == 3 * n_it and float(c['loss']) != float(a['loss'])


@pytest.mark.gpu
@pytest.mark.parametrize('p', [0.0, 0.5, 1.0, None])
def test_scheduled_sampling_step_matches_oracle(p):
    from parity_util import oracle_and_engine
    cfg = karel_config('full', batch_size=4, k=3, scheduled_sampling=True, scheduled_sampling_decay_steps=200)
    orc, eng, batch, pm, sm = oracle_and_engine(cfg, use_graph=False)
    step = 0
    if p is None:                       # the schedule itself: global step 100 of 200 -> p = 0.45
        eng.adam_state[0] = 100.0
        step = 100
    else:
        eng.sched_p_override = p
    eng.stage_batch(batch)
    eng.forward()
    eng.backward()
    torch.cuda.synchronize()
    eng.check_device()
    B, k, T = cfg.batch_size, cfg.k, cfg.max_demo_len
    fed_p = eng.prog['fed'].cpu().numpy()
    fed_a = eng.act['fed'].cpu().numpy().reshape(B, k, T)
    sched = dict(step=step, seed=eng.sched_seed, replay_program=fed_p, replay_action=fed_a)
    if p is not None:
        sched['p_override'] = p
    loss_o, grad_o, out = orc.model.loss_and_grad(batch, sched=sched)
    assert not out['sched_mismatches'], out['sched_mismatches'][:5]
    # the Bernoulli draws: bit for bit over the executed steps
    n_p = int(np.asarray(batch['program_len']).max())
    assert np.array_equal(eng.prog['sampled'].cpu().numpy()[:, :n_p], out['took_program'].numpy()[:, :n_p])
    took_a = eng.act['sampled'].cpu().numpy().reshape(B, k, T)
    for i in range(k):
        n_a = int(np.asarray(batch['demo_len'])[:, i].max())
        assert np.array_equal(took_a[:, i, :n_a], out['took_action'][i].numpy()[:, :n_a])
        assert np.array_equal(fed_a[:, i, :n_a], out['fed_action'][i].numpy()[:, :n_a])
    assert np.array_equal(fed_p[:, :n_p], out['fed_program'].numpy()[:, :n_p])
    frac = eng.prog['sampled'].cpu().numpy()[:, :n_p].mean()
    want = {0.0: 0.0, 1.0: 1.0}.get(p)
    if want is not None:
        assert frac == want
    else:
        assert 0.2 < frac < 0.8
    # loss and gradients over the tokens actually fed
    assert abs(float(eng.loss[0]) - loss_o) < 1e-4
    g, go = eng.grads.cpu().numpy(), grad_o.numpy()
    gmax = np.abs(go).max()
    for e in pm:
        a, b = g[e.offset:e.offset + e.size], go[e.offset:e.offset + e.size]
        assert np.abs(a - b).max() < 1e-3 * np.abs(b).max() + 1e-5 * gmax, e.name


@pytest.mark.gpu
def test_scheduled_sampling_at_p0_equals_teacher_forcing_and_trains():
    """p = 0 feeds the ground truth everywhere: same loss / gradients as the hoisted teacher-forced
    path; and a few captured-graph train steps with the schedule run and move the sampling rate."""
    from demo2program_b200.engine import Engine
    from demo2program_b200.synthetic import make_batch
    cfg_s = karel_config('full', batch_size=4, k=3, scheduled_sampling=True, scheduled_sampling_decay_steps=4)
    cfg_t = karel_config('full', batch_size=4, k=3)
    batch = make_batch(cfg_t, seed=2)
    es, et = Engine(cfg_s, use_graph=False), Engine(cfg_t, use_graph=False)
    es.sched_p_override = 0.0
    for e in (es, et):
        e.stage_batch(batch); e.forward(); e.backward()
    torch.cuda.synchronize()
    # (same arithmetic up to the summation order of the kernels a one-step call selects)
    assert abs(float(es.loss[0]) - float(et.loss[0])) < 1e-5
    assert float((es.grads - et.grads).abs().max()) < 1e-4 * float(et.grads.abs().max())
    eg = Engine(cfg_s, use_graph=True)
    rates = []
    for i in range(5):
        assert np.isfinite(eg.train_step(batch))
        rates.append(float(eg.prog['sampled'].float().mean()))
    assert rates[0] == 0.0 and rates[-1] > 0.3        # step 0: p = 0; step >= 4: p = 0.9
