"""Karel DSL parser / interpreter / metrics: the native implementation behind
the C ABI against the Python restatement of the reference's closures (oracle/karel_dsl.py),
on seeded random programs, mutated (mostly invalid) token sequences and hand-made edge cases.
Runs on CPU (host code)."""
import numpy as np
import pytest

from demo2program_b200 import karel_dsl as kd
from demo2program_b200.synthetic import KarelSim
from demo2program_b200.vocab import karel_vocab
from oracle import karel_dsl as okd

V = karel_vocab()
ACTIONS = ['move', 'turnLeft', 'turnRight', 'pickMarker', 'putMarker']
PRIMS = ['frontIsClear', 'leftIsClear', 'rightIsClear', 'markersPresent', 'noMarkersPresent']


def rand_cond(rng, depth=0):
    if depth < 2 and rng.rand() < 0.3:
        return ['not', 'c('] + rand_cond(rng, depth + 1) + ['c)']
    return [PRIMS[rng.randint(5)]]


def rand_stmt(rng, depth):
    r = rng.rand()
    if depth >= 3 or r < 0.45:
        return [ACTIONS[rng.randint(5)]]
    if r < 0.65:
        return rand_stmt(rng, depth + 1) + rand_stmt(rng, depth + 1)
    if r < 0.75:
        return ['IF', 'c('] + rand_cond(rng) + ['c)', 'i('] + rand_stmt(rng, depth + 1) + ['i)']
    if r < 0.85:
        return (['IFELSE', 'c('] + rand_cond(rng) + ['c)', 'i('] + rand_stmt(rng, depth + 1) + ['i)', 'ELSE', 'e('] +
                rand_stmt(rng, depth + 1) + ['e)'])
    if r < 0.93:
        return ['WHILE', 'c('] + rand_cond(rng) + ['c)', 'w('] + rand_stmt(rng, depth + 1) + ['w)']
    return ['REPEAT', 'R=%d' % rng.randint(20), 'r('] + rand_stmt(rng, depth + 1) + ['r)']


def rand_program(rng):
    return ['DEF', 'run', 'm('] + rand_stmt(rng, 0) + ['m)']


def ids(words):
    return [V.token2int[w] for w in words]


def test_vocabulary_ids_the_parser_hardcodes():
    # csrc/karel_dsl.cu enum Tok mirrors the reference vocabulary order
    assert V.token2int['m)'] == 3 and V.token2int['R=0'] == 11 and V.token2int['R=19'] == 30
    assert V.token2int['REPEAT'] == 31 and V.token2int['IF'] == 38 and V.token2int['WHILE'] == 49
    assert V.token2int['turnRight'] == 5 and V.token2int['turnLeft'] == 6 and len(V.int2token) == 50


def test_syntax_matches_oracle_on_valid_and_mutated_programs():
    rng = np.random.RandomState(7)
    n_valid = n_invalid = 0
    for _ in range(400):
        words = rand_program(rng)
        assert okd.parse(words)[1] and kd.check_syntax(ids(words))
        n_valid += 1
        m = list(words)
        for _ in range(rng.randint(1, 3)):
            op = rng.randint(3)
            pos = rng.randint(len(m))
            if op == 0 and len(m) > 1:
                del m[pos]
            elif op == 1:
                m.insert(pos, V.int2token[rng.randint(50)])
            else:
                m[pos] = V.int2token[rng.randint(50)]
        want = okd.parse(m)[1]
        assert kd.check_syntax(ids(m)) == want, ' '.join(m)
        n_invalid += not want
    assert n_invalid > 100
    # quirks of the reference's loop: one leftover symbol of ANY kind is accepted
    for words, want in ((['move'], True), (['R=3'], True), (['frontIsClear'], True), (['DEF'], False),
                        (['move', 'move'], True), (['DEF', 'run', 'm(', 'm)'], False), ([], False)):
        assert okd.parse(words)[1] == want and kd.check_syntax(ids(words)) == want, words
    assert not kd.check_syntax([50]) and not kd.check_syntax([-1])


@pytest.mark.parametrize('make_error', [True, False])
def test_execution_matches_oracle(make_error):
    rng = np.random.RandomState(11 + make_error)
    outcomes = {1: 0, 0: 0}
    for it in range(300):
        words = rand_program(rng)
        s0 = KarelSim(rng).s.copy()
        if it % 7 == 0:
            s0[:, :, 5:] = False
            s0[:, :, 5 + rng.randint(8, 11)] = True       # near the marker cap
        st_o, sh_o = okd.execute(words, s0, make_error)
        st_n, sh_n = kd.execute(ids(words), s0, make_error)
        assert st_n == st_o, ' '.join(words)
        outcomes[st_o] += 1
        if st_o == 1:
            assert sh_n.shape[0] == len(sh_o) and np.array_equal(sh_n, np.stack(sh_o, 0)), ' '.join(words)
    assert outcomes[1] > 30 and outcomes[0] > (30 if make_error else 5)   # no_error: only time-outs fail


def test_call_budget_and_degenerate_roots():
    rng = np.random.RandomState(3)
    s0 = KarelSim(rng).s.copy()
    loop = 'DEF run m( WHILE c( not c( markersPresent c) c) w( turnLeft w) m)'.split()
    if not s0[:, :, 6:][np.where(s0[:, :, :4])[:2]].any():
        assert okd.execute(loop, s0)[0] == 0 and kd.execute(ids(loop), s0)[0] == 0       # time-out, not a hang
    deep = 'DEF run m( REPEAT R=19 r( REPEAT R=19 r( turnLeft r) r) m)'.split()
    assert okd.execute(deep, s0)[0] == kd.execute(ids(deep), s0)[0] == 0
    ok = 'DEF run m( REPEAT R=4 r( turnRight r) m)'.split()
    st, sh = kd.execute(ids(ok), s0)
    assert st == 1 and sh.shape[0] == 5 and np.array_equal(sh[0], sh[-1])
    assert kd.execute(ids(['move', 'turnLeft']), s0, make_error=False)[0] == \
        okd.execute(['move', 'turnLeft'], s0, make_error=False)[0] == 1
    assert kd.execute(ids(['R=3']), s0)[0] == okd.execute(['R=3'], s0)[0] == 0
    assert kd.execute(ids(['DEF']), s0)[0] == okd.execute(['DEF'], s0)[0] == -1
    # max_states truncation keeps the true count semantics of the reference's padding
    st, sh = kd.execute(ids(ok), s0, max_states=3)
    assert st == 1 and sh.shape[0] == 3


def test_exact_program_comparison_matches_oracle():
    rng = np.random.RandomState(5)
    P = lambda s: s.split()
    eq = [('DEF run m( REPEAT R=2 r( move r) m)', 'DEF run m( move move m)'),
          ('DEF run m( IFELSE c( frontIsClear c) i( move i) ELSE e( move e) m)', 'DEF run m( move m)'),
          ('DEF run m( IF c( not c( not c( leftIsClear c) c) c) i( putMarker i) m)',
           'DEF run m( IF c( leftIsClear c) i( putMarker i) m)'),
          ('DEF run m( IF c( noMarkersPresent c) i( move i) m)',
           'DEF run m( IF c( not c( markersPresent c) c) i( move i) m)'),
          ('DEF run m( IFELSE c( rightIsClear c) i( move i) ELSE e( turnLeft e) m)',
           'DEF run m( IF c( rightIsClear c) i( move i) IF c( not c( rightIsClear c) c) i( turnLeft i) m)')]
    for a, b in eq:
        assert okd.programs_equal(P(a), P(b)) == 1 and kd.programs_equal(ids(P(a)), ids(P(b))) == 1, (a, b)
    ne = [('DEF run m( move turnLeft m)', 'DEF run m( turnLeft move m)'),
          ('DEF run m( WHILE c( frontIsClear c) w( move w) m)', 'DEF run m( IF c( frontIsClear c) i( move i) m)')]
    for a, b in ne:
        assert okd.programs_equal(P(a), P(b)) == 0 and kd.programs_equal(ids(P(a)), ids(P(b))) == 0
    assert kd.programs_equal(ids(['move']), ids(P('DEF run m( move m)'))) == -1
    n_eq = 0
    for _ in range(300):
        a, b = rand_program(rng), rand_program(rng)
        if rng.rand() < 0.3:
            b = list(a)
        try:
            want = okd.programs_equal(a, b)
        except OverflowError:      # deeply nested loops: the flattened list is too long to materialise
            continue
        assert kd.programs_equal(ids(a), ids(b)) == want, (' '.join(a), ' '.join(b))
        n_eq += want == 1
    assert n_eq > 50
    # nested WHILEs: 100^3 copies - compared without materialising the list
    w3 = P('DEF run m( WHILE c( frontIsClear c) w( WHILE c( leftIsClear c) w( WHILE c( rightIsClear c) w( move w) '
           'w) w) m)')
    assert kd.programs_equal(ids(w3), ids(w3)) == 1


def test_batch_metrics_match_oracle():
    rng = np.random.RandomState(21)
    B, k, T, L = 24, 3, 12, 40
    tokens = np.zeros((B, L), np.int32)
    lens = np.zeros(B, np.int32)
    same = np.zeros(B, np.uint8)
    demos = np.zeros((B, k, T, 8, 8, 16), np.uint8)
    demo_len = np.zeros((B, k), np.int32)
    for b in range(B):
        while True:
            words = rand_program(rng)
            if len(words) <= L:
                break
        truth = list(words)
        kind = b % 4
        if kind == 1:      # a different but valid program
            words = rand_program(rng)[:L]
        elif kind == 2:    # broken syntax
            words = words[:-1]
        elif kind == 3:
            same[b] = 1
        tokens[b, :len(words)] = ids(words)
        lens[b] = len(words)
        for i in range(k):
            s0 = KarelSim(rng).s.copy()
            st, sh = okd.execute(truth, s0, True)
            if st == 1 and len(sh) <= T:
                demos[b, i, :len(sh)] = np.stack(sh, 0)
                demo_len[b, i] = len(sh)
            else:          # keep a demo the prediction cannot reproduce
                demos[b, i, 0] = s0
                demo_len[b, i] = 1
    want = okd.eval_batch(tokens, lens, same, demos, demo_len, True)
    for nthreads in (1, 4):
        got = kd.eval_batch(tokens, lens, same, demos, demo_len, True, nthreads=nthreads)
        for a, b_ in zip(got, want):
            assert np.array_equal(a, b_)
    assert 0 < want[1].sum() < B * k and want[0].sum() < B


# ---- pinned against the reference's own code ----------------------------------------------------
def _golden():
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'karel_dsl_golden.json')))
    states = [np.unpackbits(np.asarray(s, np.uint8))[:8 * 8 * 16].reshape(8, 8, 16).astype(bool) for s in g['states']]
    return g, states


def test_golden_vocabulary_is_the_reference_vocabulary():
    g, _ = _golden()
    assert g['vocab'] == V.int2token


def test_parser_interpreter_and_canonical_compare_match_reference_golden():
    """tests/golden/karel_dsl_golden.json holds outputs of the REFERENCE's karel_env/dsl/dsl_parse.py,
    karel_env/karel.py and karel_env/dsl/dsl_enum_program.py, run in the build container on programs
    from the reference's own sampler (tests/golden/make_karel_dsl_golden.py).  Both the oracle
    restatement and the native library must reproduce: the parse verdict of 330 token sequences, the
    run status / history length / sha256 of the state history of 708 executions (two initial states,
    make_error on and off), and 150 canonical-form equalities."""
    import hashlib
    g, states = _golden()
    n_runs = n_ok = 0
    for c in g['cases']:
        words = [V.int2token[t] for t in c['tokens']]
        assert okd.parse(words)[1] == c['syntax'], ' '.join(words)
        assert kd.check_syntax(c['tokens']) == c['syntax'], ' '.join(words)
        for r in c['runs']:
            s0 = states[r['state']]
            st_o, hist_o = okd.execute(words, s0, r['make_error'])
            st_n, hist_n = kd.execute(c['tokens'], s0, r['make_error'])
            assert (st_o == 1) == r['ok'] and (st_n == 1) == r['ok'], (' '.join(words), r)
            n_runs += 1
            if r['ok']:
                n_ok += 1
                ho = np.stack(hist_o, 0).astype(np.uint8)
                assert ho.shape[0] == r['len'] == hist_n.shape[0]
                assert hashlib.sha256(ho.tobytes()).hexdigest() == r['sha256']
                assert hashlib.sha256(np.ascontiguousarray(hist_n.astype(np.uint8)).tobytes()).hexdigest() == r['sha256']
    assert n_runs == 708 and n_ok > 400
    for p in g['pairs']:
        a, b = g['cases'][p['a']]['tokens'], g['cases'][p['b']]['tokens']
        wa, wb = [V.int2token[t] for t in a], [V.int2token[t] for t in b]
        assert okd.programs_equal(wa, wb) == int(p['equal'])
        assert kd.programs_equal(a, b) == int(p['equal'])
