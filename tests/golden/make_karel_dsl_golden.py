"""Generates tests/golden/karel_dsl_golden.json by running the REFERENCE's own Karel DSL code here:
karel_env/dsl/dsl_parse.py (shift-reduce parser + interpreter closures), karel_env/karel.py (the
world), karel_env/dsl/dsl_enum_program.py (canonical program comparison) and
karel_env/dsl/dsl_prob.py (`random_code`, the reference's seeded program sampler) - imported from
/root/reference, which is mounted read-only in the build container (NOT on the GPU box; the tests
only read the committed JSON).

The reference is Python 2 and needs `ply.lex` (not installed).  Two shims, both applied at run
time without touching the reference tree:
  * a stand-in `ply.lex` module (whitespace tokeniser over the DSL's own t_* definitions) - only
    dsl_prob's constructor needs it; the parser / interpreter under test (dsl_parse, karel,
    dsl_enum_program) do not use ply at all; the LALR tables come from the reference's vendored
    karel_env/dsl/third_party/yacc.py;
  * `zip` rebound to a list-returning version inside dsl_parse / dsl_enum_program
    (`zip(*t)[0]`, dsl_parse.py:8).

Content: seeded programs from the reference's sampler + mutated (mostly invalid) variants; for each
the parse verdict; for parsable ones the interpreter's (status, call counter, number of states,
sha256 of the state history) on two initial states with make_error on and off; and canonical-form
equality for program pairs.  tests/test_karel_dsl.py checks oracle/karel_dsl.py and the native
library (csrc/karel_dsl.cu) against it.
"""
import builtins
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def _install_ply_shim():
    import re

    class LexToken(object):
        pass

    class Lexer(object):
        def __init__(self, module):
            self.module, self.rules = module, []
            for name in module.tokens:
                t = getattr(module, 't_' + name)
                if callable(t):
                    self.rules.append((name, re.compile(t.__doc__ + r'\Z'), t))
                else:
                    self.rules.append((name, re.compile(t + r'\Z'), None))
            self.words, self.pos, self.lineno, self.lexpos = [], 0, 1, 0

        def input(self, s):
            self.words, self.pos = s.split(), 0

        def skip(self, n):
            pass

        def token(self):
            if self.pos >= len(self.words):
                return None
            w = self.words[self.pos]
            self.pos += 1
            tok = LexToken()
            tok.value, tok.lineno, tok.lexpos, tok.lexer = w, 1, self.pos, self
            hit = None
            for name, rx, fn in self.rules:
                if rx.match(w):
                    hit = (name, fn)
                    if fn is not None:
                        break
            if hit is None:
                tok.type = 'error'
                return self.module.t_error(tok)
            tok.type = hit[0]
            return hit[1](tok) if hit[1] is not None else tok

    ply = types.ModuleType('ply')
    lex = types.ModuleType('ply.lex')

    def lex_fn(module=None, **kw):
        lex.lexer = Lexer(module)
        return lex.lexer
    lex.lex, lex.Lexer = lex_fn, Lexer
    ply.lex = lex
    sys.modules['ply'], sys.modules['ply.lex'] = ply, lex


def load_reference():
    sys.dont_write_bytecode = True           # /root/reference is read-only
    _install_ply_shim()
    for p in (os.path.join(REF, 'karel_env', 'dsl'), os.path.join(REF, 'karel_env'), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    cwd = os.getcwd()
    os.chdir('/tmp')                         # yacc.py wants to write its table module: not into the tree
    try:
        import warnings
        warnings.simplefilter('ignore')
        import dsl_prob, dsl_parse, dsl_enum_program, karel
    finally:
        os.chdir(cwd)
    for m in (dsl_parse, dsl_enum_program):
        m.zip = lambda *a: list(builtins.zip(*a))
    return dsl_prob, dsl_parse, dsl_enum_program, karel


def while_depth(words):
    d = mx = 0
    for w in words:
        if w == 'w(':
            d += 1
            mx = max(mx, d)
        elif w == 'w)':
            d -= 1
    return mx


def main():
    sys.path.insert(0, os.path.join(HERE, '..', '..'))
    from demo2program_b200.synthetic import KarelSim
    dsl_prob, dsl_parse, dsl_enum_program, karel = load_reference()
    dsl = dsl_prob.KarelDSLProb(seed=123)
    vocab = dsl.int2token
    rng = np.random.RandomState(2024)
    programs = []
    for i in range(160):
        words = dsl.random_code().split()
        programs.append(words)
        m = list(words)
        for _ in range(rng.randint(1, 3)):      # mutated variant: deletion / insertion / substitution
            op, pos = rng.randint(3), rng.randint(len(m))
            if op == 0 and len(m) > 1:
                del m[pos]
            elif op == 1:
                m.insert(pos, vocab[rng.randint(50)])
            else:
                m[pos] = vocab[rng.randint(50)]
        programs.append(m)
    programs += [['move'], ['R=3'], ['frontIsClear'], ['DEF'], ['move', 'move'], ['DEF', 'run', 'm(', 'm)'],
                 'DEF run m( WHILE c( not c( frontIsClear c) c) w( turnLeft w) m)'.split(),
                 'DEF run m( REPEAT R=19 r( REPEAT R=19 r( putMarker r) r) m)'.split(),
                 'DEF run m( WHILE c( noMarkersPresent c) w( putMarker pickMarker w) m)'.split(),
                 'DEF run m( IFELSE c( markersPresent c) i( pickMarker i) ELSE e( putMarker e) m)'.split()]
    states = []
    for _ in range(2):
        sim = KarelSim(rng, 8, 8)
        for a in rng.randint(0, 5, size=3):
            sim.step(int(a))
        states.append(np.asarray(sim.s, dtype=bool).copy())
    cases = []
    for words in programs:
        code = ' '.join(words)
        ok = bool(dsl_parse.parse(code)[1])
        rec = {'tokens': [vocab.index(w) for w in words], 'syntax': ok, 'runs': []}
        # the reference's loop accepts ANY single leftover symbol; only complete programs are executable
        if ok and words[:3] == ['DEF', 'run', 'm('] and words[-1] == 'm)':
            for si, s0 in enumerate(states):
                for make_error in (True, False):
                    exe, _ = dsl_parse.parse(code)
                    w, n, s_run = exe(karel.Karel_world(s0.copy(), make_error=make_error), 0)
                    hist = np.stack(w.s_h, 0).astype(np.uint8)
                    rec['runs'].append({'state': si, 'make_error': make_error, 'ok': bool(s_run), 'n': int(n),
                                        'len': int(hist.shape[0]),
                                        'sha256': hashlib.sha256(hist.tobytes()).hexdigest()})
        cases.append(rec)
    # canonical-form comparison on pairs of complete programs (WHILE unrolls 100x per nesting level
    # in the reference: keep depth <= 1 so its lists stay small)
    full = [i for i, (c, w) in enumerate(zip(cases, programs))
            if c['syntax'] and w[:3] == ['DEF', 'run', 'm('] and w[-1] == 'm)' and while_depth(w) <= 1]
    pairs = []
    for _ in range(150):
        i = full[rng.randint(len(full))]
        j = i if rng.rand() < 0.25 else full[rng.randint(len(full))]
        a, _ = dsl_enum_program.parse(' '.join(programs[i]))
        b, _ = dsl_enum_program.parse(' '.join(programs[j]))
        pairs.append({'a': i, 'b': j, 'equal': bool(a == b)})
    out = {'vocab': vocab, 'states': [np.packbits(s.astype(np.uint8)).tolist() for s in states],
           'cases': cases, 'pairs': pairs}
    json.dump(out, open(os.path.join(HERE, 'karel_dsl_golden.json'), 'w'), separators=(',', ':'))
    nrun = sum(len(c['runs']) for c in cases)
    print('programs %d (parsable %d), runs %d (ok %d), pairs %d (equal %d)' % (
        len(cases), sum(c['syntax'] for c in cases), nrun,
        sum(r['ok'] for c in cases for r in c['runs']), len(pairs), sum(p['equal'] for p in pairs)))


if __name__ == '__main__':
    main()
