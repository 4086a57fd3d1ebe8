"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's Karel DSL parser, interpreter
and program-level metrics; the product path (demo2program_b200/csrc/karel_dsl.cu) never imports
this file.  PINNED against the reference itself: karel_env/dsl/dsl_parse.py, karel_env/karel.py and
karel_env/dsl/dsl_enum_program.py are plain Python and do run in the build container once
`zip(*t)[0]` (dsl_parse.py:8, a Python-2 idiom) is patched at run time;
tests/golden/make_karel_dsl_golden.py runs them on 330 token sequences and stores parse verdicts,
708 execution results (sha256 of the state histories) and 150 canonical-form equalities in
tests/golden/karel_dsl_golden.json, which this restatement and the native library both reproduce
bit-exactly (tests/test_karel_dsl.py).  The batch metric composition (`eval_batch`) follows
models/model_full.py, whose closures live inside the TF graph builder and cannot be imported.
The restatement follows the sources line by line:

  * parser / rule closures   karel_env/dsl/dsl_parse.py:4-13, 21-262
  * Karel world              karel_env/karel.py:33-185
  * exact-program canonical  karel_env/dsl/dsl_enum_program.py:4-222
  * metrics                  models/model_full.py:602-616, 712-727, 747-787, 870-897

Written as data-driven rule tables + small evaluator functions (Python 3).
"""
import numpy as np

from demo2program_b200.vocab import karel_vocab

MAX_FUNC_CALL = 100
MAX_WHILE = 100
MAX_NUM_MARKER = 10

# (right-hand side, left-hand side) in the reference's order
_STMT_KINDS = ['while_stmt', 'repeat_stmt', 'stmt_stmt', 'action', 'if_stmt', 'ifelse_stmt']
_CONDS = ['frontIsClear', 'leftIsClear', 'rightIsClear', 'markersPresent', 'noMarkersPresent']
_ACTIONS = ['move', 'turnLeft', 'turnRight', 'pickMarker', 'putMarker']     # karel.py action index order
RULES = [('DEF run m( stmt m)', 'prog')]
RULES += [(s, 'stmt') for s in _STMT_KINDS]
RULES += [('stmt stmt', 'stmt_stmt'),
          ('IF c( cond c) i( stmt i)', 'if_stmt'),
          ('IFELSE c( cond c) i( stmt i) ELSE e( stmt e)', 'ifelse_stmt'),
          ('WHILE c( cond c) w( stmt w)', 'while_stmt'),
          ('REPEAT cste r( stmt r)', 'repeat_stmt'),
          ('cond_without_not', 'cond'),
          ('not c( cond c)', 'cond')]
RULES += [(c, 'cond_without_not') for c in _CONDS]
RULES += [(a, 'action') for a in _ACTIONS]
RULES += [('R=%d' % i, 'cste') for i in range(20)]
RULES = [(rhs.split(), lhs) for rhs, lhs in RULES]


def parse(tokens):
    """tokens: list of token strings.  Returns (tree, ok); tree = (lhs, rhs_string, children)."""
    toks = list(tokens)[::-1]
    if not toks:
        return None, False           # the reference raises IndexError on an empty program
    stack = []
    applied = False
    while toks or len(stack) != 1:
        if applied:
            applied = False
        else:
            stack.append((toks.pop(), None, None))
        for rhs, lhs in RULES:
            n = len(rhs)
            if len(stack) >= n and [s[0] for s in stack[-n:]] == rhs:
                kids = stack[-n:]
                del stack[-n:]
                stack.append((lhs, ' '.join(rhs), kids))
                applied = True
                break
        if not applied and not toks:
            return None, False
    return stack[0], True


class World:
    """karel.py:33-185."""

    def __init__(self, s, make_error=True):
        self.s = np.array(s).astype(bool)
        self.h, self.w = self.s.shape[:2]
        self.make_error = make_error
        self.s_h = [self.s.copy()]

    def loc(self):
        x, y, z = np.where(self.s[:, :, :4] > 0)
        return int(x[0]), int(y[0]), int(z[0])

    def clear(self, face):
        x, y, z = self.loc()
        d = {'front': [(-1, 0), (0, 1), (1, 0), (0, -1)],
             'left': [(0, -1), (-1, 0), (0, 1), (1, 0)],
             'right': [(0, 1), (1, 0), (0, -1), (-1, 0)]}[face][z]
        nx, ny = x + d[0], y + d[1]
        if nx >= self.h or nx < 0 or ny >= self.w or ny < 0:
            return False
        return not self.s[nx, ny, 4]

    def perceive(self, name):
        if name == 'frontIsClear':
            return self.clear('front')
        if name == 'leftIsClear':
            return self.clear('left')
        if name == 'rightIsClear':
            return self.clear('right')
        x, y, _ = self.loc()
        m = int(np.sum(self.s[x, y, 6:]))
        return m > 0 if name == 'markersPresent' else m == 0

    def act(self, a):
        x, y, z = self.loc()
        if a == 0:
            if self.clear('front'):
                dx, dy = [(-1, 0), (0, 1), (1, 0), (0, -1)][z]
                self.s[x + dx, y + dy, :4] = self.s[x, y, :4]
                self.s[x, y, :4] = False
            else:
                if self.make_error:
                    raise RuntimeError('Failed to move.')
                v = np.zeros(4, bool)
                v[(z + 2) % 4] = True
                self.s[x, y, :4] = v
        elif a in (1, 2):
            v = np.zeros(4, bool)
            v[(a * 2 - 3 + z) % 4] = True
            self.s[x, y, :4] = v
        elif a in (3, 4):
            num = int(np.argmax(self.s[x, y, 5:]))
            new = a * 2 - 7 + num
            if new < 0 or new > MAX_NUM_MARKER - 1:
                if self.make_error:
                    raise RuntimeError('marker error')
                new = num
            v = np.zeros(MAX_NUM_MARKER + 1, bool)
            v[new] = True
            self.s[x, y, 5:] = v
        else:
            raise RuntimeError('Invalid action')
        self.s_h.append(self.s.copy())


def _cond(node, k, n):
    """-> (n, success, value)"""
    lhs, rhs, kids = node
    if lhs == 'cond' and rhs == 'cond_without_not':
        if n > MAX_FUNC_CALL:
            return n, False, False
        return _cond(kids[0], k, n)
    if lhs == 'cond':                      # not c( cond c)
        if n > MAX_FUNC_CALL:
            return n, False, False
        n, s, c = _cond(kids[2], k, n)
        return n, s, not c
    if n > MAX_FUNC_CALL:                  # cond_without_not (unreachable in the reference too)
        return n, False, False
    return n, True, k.perceive(rhs)


def _run(node, k, n):
    """-> (n, success)"""
    lhs, rhs, kids = node
    if n > MAX_FUNC_CALL:
        return n, False
    if lhs == 'prog':
        return _run(kids[3], k, n + 1)
    if lhs == 'stmt':
        return _run(kids[0], k, n + 1)
    if lhs == 'stmt_stmt':
        n, s = _run(kids[0], k, n + 1)
        if not s:
            return n, s
        if n > MAX_FUNC_CALL:
            return n, False
        return _run(kids[1], k, n)
    if lhs == 'if_stmt':
        n, s, c = _cond(kids[2], k, n + 1)
        if not s:
            return n, s
        return _run(kids[5], k, n) if c else (n, s)
    if lhs == 'ifelse_stmt':
        n, s, c = _cond(kids[2], k, n + 1)
        if not s:
            return n, s
        return _run(kids[5] if c else kids[9], k, n)
    if lhs == 'while_stmt':
        n, s, c = _cond(kids[2], k, n)
        if not s:
            return n, s
        while c:
            n, s = _run(kids[5], k, n)
            if not s:
                return n, s
            n, s, c = _cond(kids[2], k, n)
            if not s:
                return n, s
        return n, s
    if lhs == 'repeat_stmt':
        n += 1
        s = True
        for _ in range(int(kids[1][1][2:])):
            n, s = _run(kids[3], k, n)
            if not s:
                return n, s
        return n, s
    if lhs == 'action':
        try:
            k.act(_ACTIONS.index(rhs))
        except Exception:
            return n, False
        return n, True
    raise TypeError('not executable: ' + lhs)     # the reference raises as well (cste / cond roots)


def execute(tokens, state0, make_error=True):
    """-> (status, s_h): 1 ran to completion, 0 failed at run time, -1 does not parse."""
    tree, ok = parse(tokens)
    if not ok:
        return -1, []
    k = World(state0, make_error)
    try:
        n, s = _run(tree, k, 0)
    except (TypeError, IndexError):
        return 0, []
    return (1, k.s_h) if s else (0, [])


def _flat_cond(node):
    lhs, rhs, kids = node
    if lhs == 'cond' and rhs == 'cond_without_not':
        return _flat_cond(kids[0])
    if lhs == 'cond':
        c = _flat_cond(kids[2])
        return c[1:] if c[0] == 'not' else ['not'] + c
    return ['not', 'markersPresent'] if rhs == 'noMarkersPresent' else [rhs]


def flatten(node, cap=2_000_000):
    """dsl_enum_program.py canonical token list (raises OverflowError beyond `cap` words)."""
    lhs, rhs, kids = node
    if lhs == 'prog':
        out = flatten(kids[3], cap)
    elif lhs == 'stmt':
        out = flatten(kids[0], cap)
    elif lhs == 'stmt_stmt':
        out = flatten(kids[0], cap) + flatten(kids[1], cap)
    elif lhs == 'if_stmt':
        out = ['if'] + _flat_cond(kids[2]) + flatten(kids[5], cap)
    elif lhs == 'ifelse_stmt':
        s1, s2 = flatten(kids[5], cap), flatten(kids[9], cap)
        if s1 == s2:
            out = s1
        else:
            c = _flat_cond(kids[2])
            e = ['if'] + c[1:] if c[0] == 'not' else ['if', 'not'] + c
            out = ['if'] + c + s1 + e + s2
    elif lhs == 'while_stmt':
        unit = ['if'] + _flat_cond(kids[2]) + flatten(kids[5], cap)
        if len(unit) * MAX_WHILE > cap:
            raise OverflowError
        out = unit * MAX_WHILE
    elif lhs == 'repeat_stmt':
        unit = flatten(kids[3], cap)
        cnt = int(kids[1][1][2:])
        if len(unit) * cnt > cap:
            raise OverflowError
        out = unit * cnt
    elif lhs == 'action':
        out = [rhs]
    else:
        raise TypeError(lhs)
    if len(out) > cap:
        raise OverflowError
    return out


def programs_equal(a_tokens, b_tokens):
    """1 / 0, or -1 when either side is not a complete program."""
    ta, oka = parse(a_tokens)
    tb, okb = parse(b_tokens)
    if not (oka and okb) or ta[0] != 'prog' or tb[0] != 'prog':
        return -1
    return int(flatten(ta) == flatten(tb))


def eval_batch(tokens, lens, is_same_seq, demos, demo_len, make_error=True):
    """model_full.py:602-616, 747-787, 870-897 for a batch.  tokens [B, L] ids, demos
    [B, k, T, h, w, 16].  -> is_correct_syntax [B], is_correct_execution [B, k], num_correct [B]."""
    v = karel_vocab()
    B, k, T = demos.shape[:3]
    syn = np.zeros(B, np.float32)
    exe_ok = np.zeros((B, k), np.float32)
    for b in range(B):
        words = [v.int2token[int(t)] for t in tokens[b, :int(lens[b])]]
        same = bool(is_same_seq[b])
        syn[b] = 1.0 if same or parse(words)[1] else 0.0
        for i in range(k):
            pad = np.zeros(demos.shape[2:], bool)
            n_states = 0
            if not same and syn[b] == 1.0:
                status, s_h = execute(words, demos[b, i, 0], make_error)
                if status == 1:
                    n_states = len(s_h)
                    arr = np.stack(s_h, 0)[:T]
                    pad[:arr.shape[0]] = arr
            eq = bool(np.array_equal(pad, demos[b, i].astype(bool))) and n_states == int(demo_len[b, i])
            exe_ok[b, i] = 1.0 if (eq or same) else 0.0
    return syn, exe_ok, exe_ok.sum(-1)
