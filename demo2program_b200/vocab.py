"""Karel / ViZDoom DSL token tables (ids only; the parsers are out of scope).

Karel: token order of reference karel_env/dsl/dsl_prob.py:13-28 with `INT`
expanded to R=0..R=19 (karel_env/dsl/dsl_base.py:80-91): 50 tokens, `m)` = 3 is
the greedy-decode end token (reference models/model_full.py:428).
"""

KAREL_TOKENS = (
    ['DEF', 'run', 'm(', 'm)', 'move', 'turnRight', 'turnLeft', 'pickMarker',
     'putMarker', 'r(', 'r)'] + ['R=%d' % i for i in range(20)] +
    ['REPEAT', 'c(', 'c)', 'i(', 'i)', 'e(', 'e)', 'IF', 'IFELSE', 'ELSE',
     'frontIsClear', 'leftIsClear', 'rightIsClear', 'markersPresent',
     'noMarkersPresent', 'not', 'w(', 'w)', 'WHILE'])


class Vocab:
    def __init__(self, tokens):
        self.int2token = list(tokens)
        self.token2int = {t: i for i, t in enumerate(self.int2token)}

    def intseq2str(self, seq):
        return ' '.join(self.int2token[int(i)] for i in seq)

    def str2intseq(self, code):
        return [self.token2int[t] for t in code.split()]


def karel_vocab():
    return Vocab(KAREL_TOKENS)
