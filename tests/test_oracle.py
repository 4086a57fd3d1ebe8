"""CPU tests of the oracle: known-answer checks derived from the deterministic
artefacts of the reference (vocabulary table, SAME-padding geometry,
TF op definitions) and a cross-check of the two independent restatements
(torch ops vs NumPy loops), plus finite-difference gradient checks.
The reference holds no golden vectors for this path (parity unpinned)."""
import numpy as np
import torch

from oracle import numpy_ref as N
from oracle import tf_ops as T
from demo2program_b200.config import karel_config, vizdoom_config
from demo2program_b200.manifest import build_manifests
from demo2program_b200.synthetic import make_batch
from demo2program_b200.vocab import karel_vocab

rs = np.random.RandomState(0)
t64 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)


def test_vocab_table_known_answers():
    v = karel_vocab()
    assert len(v.int2token) == 50 == len(N.KAREL_VOCAB)
    assert v.int2token == N.KAREL_VOCAB
    # SURVEY 8c-1 known answers
    for tok, idx in [('DEF', 0), ('run', 1), ('m(', 2), ('m)', 3), ('move', 4), ('r)', 10),
                     ('R=0', 11), ('R=19', 30), ('REPEAT', 31), ('IF', 38), ('ELSE', 40),
                     ('frontIsClear', 41), ('not', 46), ('WHILE', 49)]:
        assert v.token2int[tok] == idx
    assert v.intseq2str([0, 1, 2, 4, 3]) == 'DEF run m( move m)'


def test_same_padding_geometry_known_answers():
    # Karel 8->4->2->1 and ViZDoom 80->40->20->10->5->3
    assert [T.same_pad_3x3_s2(n) for n in (8, 4, 2)] == [(4, 0, 1), (2, 0, 1), (1, 0, 1)]
    assert [T.same_pad_3x3_s2(n)[0] for n in (80, 40, 20, 10, 5)] == [40, 20, 10, 5, 3]
    assert T.same_pad_3x3_s2(5) == (3, 1, 1)        # ViZDoom L5 pads (1,1)
    assert T.same_pad_3x3_s2(80) == (40, 0, 1)
    assert karel_config().feature_dim() == 48
    assert vizdoom_config().feature_dim() == 432
    g = karel_config().conv_geometry()
    assert [(x[6], x[7]) for x in g] == [(0, 0)] * 3


def test_conv_matches_numpy_loops():
    for (h, w, cin, cout) in [(8, 8, 16, 16), (5, 5, 4, 8), (2, 2, 3, 4), (7, 6, 2, 4)]:
        x = rs.randn(2, h, w, cin); wt = rs.randn(3, 3, cin, cout); b = rs.randn(cout)
        a = T.conv2d_3x3_s2_same(t64(x), t64(wt), t64(b)).numpy()
        np.testing.assert_allclose(a, N.conv2d_3x3_s2_same(x, wt, b), rtol=1e-10, atol=1e-10)


def test_lrelu_definition_and_gradient_at_zero():
    x = torch.tensor([-2.0, 0.0, 3.0], dtype=torch.float64, requires_grad=True)
    y = T.lrelu(x)
    np.testing.assert_allclose(y.detach().numpy(), [-0.4, 0.0, 3.0])
    y.sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), [0.2, 0.6, 1.0])
    np.testing.assert_allclose(N.lrelu(np.array([-2.0, 0.0, 3.0])), [-0.4, 0.0, 3.0])


def test_batch_norm_train_and_moving_update():
    x = rs.randn(6, 3, 4) * 2 + 1
    g, b = rs.rand(4) + 0.5, rs.randn(4)
    y, mm, mv = T.batch_norm(t64(x), t64(g), t64(b), torch.zeros(4, dtype=torch.float64),
                             torch.ones(4, dtype=torch.float64), True)
    yn, mean, var = N.batch_norm_train(x, g, b)
    np.testing.assert_allclose(y.numpy(), yn, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(mm.numpy(), 0.1 * mean, rtol=1e-10)           # 0 - (0-mean)*0.1
    np.testing.assert_allclose(mv.numpy(), 1 - (1 - var) * 0.1, rtol=1e-10)  # biased var, no debias
    ye, _, _ = T.batch_norm(t64(x), t64(g), t64(b), t64(mean), t64(var), False)
    np.testing.assert_allclose(ye.numpy(), yn, rtol=1e-10, atol=1e-12)


def test_lstm_cell_known_answer():
    # zero kernel/bias: i=o=0.5, j=0, f=sigmoid(1) -> c' = c*sigmoid(1); h' = tanh(c')*0.5
    c = t64([[1.0, -2.0]]); h = t64([[0.3, 0.4]]); x = t64([[5.0]])
    c2, h2 = T.lstm_cell(x, c, h, torch.zeros(3, 8, dtype=torch.float64),
                         torch.zeros(8, dtype=torch.float64))
    s1 = 1 / (1 + np.exp(-1.0))
    np.testing.assert_allclose(c2.numpy(), [[s1, -2 * s1]], rtol=1e-12)
    np.testing.assert_allclose(h2.numpy(), np.tanh([[s1, -2 * s1]]) * 0.5, rtol=1e-12)


def test_dynamic_rnn_masking_matches_numpy():
    R, Tn, In, H = 5, 6, 3, 4
    x = rs.randn(R, Tn, In); k = rs.randn(In + H, 4 * H) * 0.3; b = rs.randn(4 * H) * 0.1
    lens = np.array([6, 3, 1, 0, 4]); c0 = rs.randn(R, H); h0 = rs.randn(R, H)
    y, h, c = T.dynamic_rnn(t64(x), torch.tensor(lens), t64(k), t64(b), t64(c0), t64(h0))
    yn, hn, cn = N.dynamic_rnn(x, lens, k, b, c0, h0)
    np.testing.assert_allclose(y.numpy(), yn, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(h.numpy(), hn, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(c.numpy(), cn, rtol=1e-10, atol=1e-12)
    assert np.all(y.numpy()[3] == 0) and np.all(h.numpy()[3] == h0[3])   # len 0: state copied
    assert np.all(y.numpy()[1, 3:] == 0)


def test_embedding_out_of_range_is_zero_row():
    tab = t64(rs.randn(7, 3))
    out = T.embedding_lookup_gpu(tab, torch.tensor([[0, 6, 7, 8]]))
    assert np.all(out.numpy()[0, 2:] == 0) and np.all(out.numpy()[0, 1] == tab.numpy()[6])


def test_losses_match_numpy_and_uniform_known_answer():
    R, L, V = 3, 5, 6
    lg = rs.randn(R, L, V); lab = np.eye(V)[rs.randint(0, V, (R, L))]; lens = np.array([5, 2, 3])
    a = float(T.softmax_ce_loss(t64(lg), t64(lab), torch.tensor(lens)))
    np.testing.assert_allclose(a, N.softmax_ce_loss(lg, lab, lens), rtol=1e-10)
    # all-zero logits -> log(V)
    z = float(T.softmax_ce_loss(torch.zeros(R, L, V, dtype=torch.float64), t64(lab), torch.tensor(lens)))
    np.testing.assert_allclose(z, np.log(V), rtol=1e-12)
    pl = rs.randn(R, L, 4); pz = (rs.rand(R, L, 4) > 0.5).astype(float)
    np.testing.assert_allclose(float(T.sigmoid_ce_loss(t64(pl), t64(pz), torch.tensor(lens))),
                               N.sigmoid_ce_loss(pl, pz, lens), rtol=1e-10)
    np.testing.assert_allclose(float(T.sigmoid_ce_loss(torch.zeros(R, L, 4, dtype=torch.float64),
                                                       t64(pz), torch.tensor(lens))), np.log(2.0))


def test_decode_training_runs_to_batch_max_then_zero_pads():
    R, L, E, H, V = 3, 6, 2, 3, 4
    x = t64(rs.randn(R, L, E)); k = t64(rs.randn(E + H, 4 * H)); b = t64(rs.randn(4 * H))
    pr = t64(rs.randn(H, V)); z = torch.zeros(R, H, dtype=torch.float64)
    lg = T.decode_training(x, torch.tensor([2, 4, 1]), z, z, k, b, pr, L).numpy()
    assert np.all(lg[:, 4:] == 0) and np.all(np.abs(lg[:, :4]).sum(-1) > 0)   # max len 4


def test_greedy_decode_lengths_and_stop():
    H, V = 3, 5
    table = t64(rs.randn(V + 1, H)); k = t64(rs.randn(2 * H, 4 * H)); b = t64(rs.randn(4 * H))
    pr = t64(rs.randn(H, V)); z = t64(rs.randn(4, H))
    lg, ln, tok = T.decode_greedy(lambda i: T.embedding_lookup_gpu(table, i), V, 3, z, z, k, b, pr, 7)
    assert lg.shape == (4, 7, V) and tok.shape == (4, 7)
    for r in range(4):
        hits = np.where(tok[r].numpy() == 3)[0]
        n = int(ln[r])
        assert 1 <= n <= 7
        if len(hits) and hits[0] + 1 <= n:
            assert n == hits[0] + 1
    n_exec = int(ln.max())
    assert np.all(lg.numpy()[:, n_exec:] == 0)


def test_adam_and_clip_match_numpy():
    p = rs.randn(50); g = rs.randn(50) * 10; m = np.zeros(50); v = np.zeros(50)
    tp, tm, tv = t64(p), t64(m), t64(v)
    for step in (1, 2, 3):
        (gc,), norm = T.clip_by_global_norm([t64(g)], 20.0)
        T.adam_step(tp, gc, tm, tv, step)
        p, m, v, nn = N.adam_clip_step(p, g, m, v, step)
        np.testing.assert_allclose(float(norm), nn, rtol=1e-12)
        np.testing.assert_allclose(tp.numpy(), p, rtol=1e-10)
    # first Adam step moves every coordinate by ~lr (epsilon outside the correction)
    p0 = t64(np.zeros(3)); T.adam_step(p0, t64([1.0, -2.0, 0.5]), t64(np.zeros(3)), t64(np.zeros(3)), 1)
    np.testing.assert_allclose(p0.numpy(), [-1e-3, 1e-3, -1e-3], rtol=1e-6)


def test_param_counts_match_survey_appendix_b():
    counts = {m: build_manifests(karel_config(m))[0].num_params()
              for m in ('full', 'summarizer', 'synthesis_baseline')}
    assert counts['full'] == 11210784
    assert counts['synthesis_baseline'] == 3320864
    pm, _ = build_manifests(karel_config('full'))
    assert pm['Demo_Encoder/rnn/basic_lstm_cell/kernel'].shape == (560, 2048)
    assert pm['Program_Decoder/Token_Embedding/embedding_map'].shape == (51, 512)
    assert all(e.offset % 4 == 0 for e in pm)


def test_synthetic_batch_shapes_and_loader_quirk():
    cfg = karel_config('full', batch_size=3, k=4)
    b = make_batch(cfg, seed=5)
    assert b['s_h'].shape == (3, 4, 20, 8, 8, 16) and b['s_h'].dtype == np.uint8
    assert b['program'].shape == (3, 50, 50) and b['a_h'].shape == (3, 4, 20, 6)
    s = b['s_h'].astype(int)
    dl = b['demo_len'].astype(int)
    for bi in range(3):
        n = int(b['program_len'][bi, 0])
        assert b['program_tokens'][bi, n - 1] == 3 and b['program'][bi, :, :n].sum() == n
        amax = dl[bi].max() - 1
        for i in range(4):
            live = s[bi, i, :dl[bi, i]]
            assert np.all(live[..., :4].sum(axis=(1, 2, 3)) == 1)      # one hero cell
            assert np.all(live[..., 5:].sum(-1) == 1)                  # one marker channel per cell
            assert s[bi, i, dl[bi, i]:].sum() == 0                     # zero padded
            # F10: <e> sits at the program-max action length, shorter demos padded with token 0
            assert b['a_h_tokens'][bi, i, amax] == 5
            assert np.all(b['a_h_tokens'][bi, i, dl[bi, i] - 1:amax] == 0)


def test_full_model_gradient_finite_difference():
    from oracle.models import OracleModel
    cfg = karel_config('full', batch_size=2, k=2, num_lstm_cell_units=8, max_program_len=6,
                       max_demo_len=4)
    pm, sm = build_manifests(cfg)
    batch = make_batch(cfg, seed=2, min_demo_len=2, min_prog_len=5)
    p0 = pm.init_flat(3).astype(np.float64)
    m = OracleModel(cfg, p0, sm.init_flat(3))
    loss, grad, _ = m.loss_and_grad(batch)
    r = np.random.RandomState(1)
    for name in ['Demo_Encoder/State_Encoder/conv2/Conv/weights',
                 'Demo_Encoder/rnn/basic_lstm_cell/kernel',
                 'demo_h_summary/rn_pool/fc1/fully_connected/weights',
                 'Action_Decoder/Token_Embedding/embedding_map',
                 'Per_Decoder/dynamic_decoder/output_projection/kernel']:
        e = pm[name]
        for _ in range(3):
            i = e.offset + r.randint(e.size)
            d = 1e-5
            pp, pn = p0.copy(), p0.copy()
            pp[i] += d; pn[i] -= d
            lp = float(OracleModel(cfg, pp, sm.init_flat(3)).forward(batch)['loss'])
            ln = float(OracleModel(cfg, pn, sm.init_flat(3)).forward(batch)['loss'])
            fd = (lp - ln) / (2 * d)
            assert abs(fd - float(grad[i])) < 1e-6 + 1e-4 * abs(fd), (name, fd, float(grad[i]))


def test_oracle_reproduces_golden():
    """tests/golden/oracle_golden.json (generated by tests/golden/make_oracle_golden.py) pins the
    oracle's arithmetic: loss, gradient norms per variable group, a projection of the flat gradient,
    logits / summary slices for the three teacher-forced model families."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, 'golden'))
    import make_oracle_golden as G
    golden = json.load(open(os.path.join(here, 'golden', 'oracle_golden.json')))
    assert [(g['model'], g['B'], g['k']) for g in golden] == G.CASES
    for g in golden[:2]:          # the two small families keep the CPU suite short
        r = G.oracle_case(g['model'], g['B'], g['k'])
        assert abs(r['loss'] - g['loss']) < 1e-10
        assert abs(r['grad_norm'] - g['grad_norm']) < 1e-9 * max(1.0, g['grad_norm'])
        assert abs(r['grad_projection'] - g['grad_projection']) < 1e-9
        for k_, v in g['grad_group_sqnorm'].items():
            assert abs(r['grad_group_sqnorm'][k_] - v) < 1e-9 * max(1.0, v), k_
        np.testing.assert_allclose(r['pred_program_slice'], g['pred_program_slice'], atol=1e-9)
        np.testing.assert_allclose(r['demo_h_summary_slice'], g['demo_h_summary_slice'], atol=1e-9)


def test_synthetic_state_sampler_matches_reference_golden():
    """tests/golden/karel_states_golden.json: states drawn by the REFERENCE's
    karel_env/generator.py KarelStateGenerator (run in the build container,
    tests/golden/make_state_golden.py); synthetic.KarelSim must draw bit-identical ones from the
    same RandomState stream."""
    import json
    import os
    from demo2program_b200.synthetic import KarelSim
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'karel_states_golden.json')))
    for seed, states in g.items():
        rng = np.random.RandomState(int(seed))
        for st in states:
            sim = KarelSim(rng, 8, 8, 0.1)
            want = np.unpackbits(np.asarray(st['bits'], np.uint8))[:8 * 8 * 16].reshape(8, 8, 16).astype(bool)
            assert np.array_equal(np.asarray(sim.s, bool), want), seed
            assert (sim.y, sim.x) == (st['y'], st['x']) and int(sim.s[:, :, 4].sum()) == st['walls']
