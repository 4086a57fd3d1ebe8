"""Synthetic batches with the reference's shapes, dtypes and loader quirks.

There is no dataset offline, so bench/tests use seeded synthetic examples whose
frames are genuine Karel states (one hero cell in channels 0-3, walls in
channel 4, exactly one marker-count channel 5-15 per cell), following the state
layout of reference karel_env/generator.py:18-44 and the transition rules of
karel_env/karel.py:138-185, and whose tensors are assembled exactly like
`Dataset.get_data` + `load_fn` (reference karel_env/dataset_karel.py:38-115,
karel_env/input_ops_karel.py:52-75), including the action-padding quirk
(SURVEY F10): action one-hots are built from the per-program zero-padded a_h
matrix, so shorter demos see token 0 as padding and `<e>` only at the
program-max position.
"""
import numpy as np

# heading -> (dy, dx); channel index = heading (N, E, S, W)
_DELTA = ((-1, 0), (0, 1), (1, 0), (0, -1))


class KarelSim:
    """Tiny Karel world used only to make realistic synthetic frames."""

    def __init__(self, rng, h=8, w=8, wall_prob=0.1):
        self.rng, self.h, self.w = rng, h, w
        s = np.zeros((h, w, 16), dtype=bool)
        s[:, :, 4] = rng.rand(h, w) > 1 - wall_prob
        s[0, :, 4] = s[h - 1, :, 4] = True
        s[:, 0, 4] = s[:, w - 1, 4] = True
        while True:
            y, x = rng.randint(0, h), rng.randint(0, w)
            if not s[y, x, 4]:
                break
        self.y, self.x, self.d = y, x, rng.randint(0, 4)
        s[y, x, self.d] = True
        s[:, :, 6] = (rng.rand(h, w) > 0.9) & ~s[:, :, 4]
        s[:, :, 5] = ~s[:, :, 6]
        self.s = s

    def _clear(self, d):
        dy, dx = _DELTA[d % 4]
        y, x = self.y + dy, self.x + dx
        if y < 0 or x < 0 or y >= self.h or x >= self.w:
            return False
        return not self.s[y, x, 4]

    def perception(self):
        markers = self.s[self.y, self.x, 6:].any()
        return np.array([self._clear(self.d), self._clear(self.d + 3),
                         self._clear(self.d + 1), markers, not markers])

    def step(self, a):
        s = self.s
        if a == 0:  # move (blocked -> turn around)
            if self._clear(self.d):
                dy, dx = _DELTA[self.d]
                s[self.y, self.x, :4] = False
                self.y += dy
                self.x += dx
            else:
                s[self.y, self.x, :4] = False
                self.d = (self.d + 2) % 4
            s[self.y, self.x, self.d] = True
        elif a in (1, 2):  # turns
            s[self.y, self.x, :4] = False
            self.d = (self.d + (a * 2 - 3)) % 4
            s[self.y, self.x, self.d] = True
        else:  # pick (3) / put (4) marker
            n = int(np.argmax(s[self.y, self.x, 5:]))
            m = n + (a * 2 - 7)
            if m < 0 or m > 10:
                m = n
            s[self.y, self.x, 5:] = False
            s[self.y, self.x, 5 + m] = True


def _demo_set(rng, n_demo, T, A_real, h, w, min_len):
    """frames [n,T,h,w,16] bool, lens [n], a_h zero-padded [n, max_a], per [n,T,5]."""
    frames = np.zeros((n_demo, T, h, w, 16), dtype=bool)
    per = np.zeros((n_demo, T, 5), dtype=np.float64)
    lens = rng.randint(min_len, T + 1, size=n_demo)
    acts = []
    for i in range(n_demo):
        sim = KarelSim(rng, h, w)
        frames[i, 0] = sim.s
        per[i, 0] = sim.perception()
        a_seq = rng.randint(0, A_real, size=lens[i] - 1)
        for t, a in enumerate(a_seq):
            sim.step(int(a))
            frames[i, t + 1] = sim.s
            per[i, t + 1] = sim.perception()
        acts.append(a_seq)
    return frames, lens, acts, per


def _action_onehot(acts_padded, T, A_real):
    """Restates reference karel_env/dataset_karel.py:67-79 (quirk F10)."""
    out = []
    for toks in acts_padded:
        a = np.zeros((T, A_real + 1), dtype=bool)
        a[np.arange(len(toks)), toks] = True
        a[len(toks), A_real] = True  # <e>
        out.append(a)
    a_h = np.stack(out, 0)
    return a_h, np.argmax(a_h, axis=2)


def make_batch(cfg, seed=123, batch_size=None, min_demo_len=8, min_prog_len=8,
               frames_dtype=np.uint8):
    """Build one synthetic batch dict with the reference's keys
    (reference models/model_full.py:185-206).  `s_h`/`test_s_h` are stored as
    `frames_dtype` (uint8 = the dataset's on-disk bool; float32 = what the
    reference load_fn feeds)."""
    if cfg.dataset_type != 'karel':
        return make_vizdoom_batch(cfg, seed, batch_size, frames_dtype)
    rng = np.random.RandomState(seed)
    B = batch_size or cfg.batch_size
    k, tk, T = cfg.k, cfg.test_k, cfg.max_demo_len
    V, L, A = cfg.dim_program_token, cfg.max_program_len, cfg.action_space
    out = {key: [] for key in (
        'id', 'program', 'program_tokens', 's_h', 'test_s_h', 'a_h',
        'a_h_tokens', 'test_a_h', 'test_a_h_tokens', 'program_len', 'demo_len',
        'test_demo_len', 'per', 'test_per')}
    min_demo_len = min(min_demo_len, T)
    for b in range(B):
        n = rng.randint(min(min_prog_len, L), L + 1)
        toks = np.concatenate([[0, 1, 2], rng.randint(4, V, size=n - 4), [3]])
        program = np.zeros((V, L), dtype=bool)
        program[toks, np.arange(n)] = True
        ptoks = np.zeros(L, dtype=np.int32)
        ptoks[:n] = toks
        for pre, nd in (('', k), ('test_', tk)):
            frames, lens, acts, per = _demo_set(rng, nd, T, A - 1, cfg.h, cfg.w,
                                                min_demo_len)
            # generator.py:119-123: a_h zero-padded to the program-max length
            max_a = max(len(a) for a in acts)
            padded = np.zeros((nd, max_a), dtype=np.int64)
            for i, a in enumerate(acts):
                padded[i, :len(a)] = a
            a_h, a_tok = _action_onehot(padded, T, A - 1)
            out[pre + 's_h'].append(frames)
            out[pre + 'a_h'].append(a_h)
            out[pre + 'a_h_tokens'].append(a_tok)
            out[pre + 'demo_len'].append(lens)
            out[pre + 'per'].append(per)
        out['id'].append(('synthetic_%06d' % b).encode())
        out['program'].append(program)
        out['program_tokens'].append(ptoks)
        out['program_len'].append(np.array([n], dtype=np.float32))
    batch = {
        'id': np.array(out['id']),
        'program': np.stack(out['program']).astype(np.float32),
        'program_tokens': np.stack(out['program_tokens']).astype(np.int32),
        's_h': np.stack(out['s_h']).astype(frames_dtype),
        'test_s_h': np.stack(out['test_s_h']).astype(frames_dtype),
        'a_h': np.stack(out['a_h']).astype(np.float32),
        'a_h_tokens': np.stack(out['a_h_tokens']).astype(np.int32),
        'test_a_h': np.stack(out['test_a_h']).astype(np.float32),
        'test_a_h_tokens': np.stack(out['test_a_h_tokens']).astype(np.int32),
        'program_len': np.stack(out['program_len']).astype(np.float32),
        'demo_len': np.stack(out['demo_len']).astype(np.float32),
        'test_demo_len': np.stack(out['test_demo_len']).astype(np.float32),
        'per': np.stack(out['per']).astype(np.float32),
        'test_per': np.stack(out['test_per']).astype(np.float32),
    }
    return batch


def make_vizdoom_batch(cfg, seed=123, batch_size=None, frames_dtype=np.uint8):
    """ViZDoom-shaped batch: frames are raw 0..255 pixel values, never
    normalised (reference vizdoom_env/dataset_vizdoom.py:48-140)."""
    rng = np.random.RandomState(seed)
    B = batch_size or cfg.batch_size
    k, tk, T = cfg.k, cfg.test_k, cfg.max_demo_len
    V, L, A, P = (cfg.dim_program_token, cfg.max_program_len, cfg.action_space,
                  cfg.per_dim)
    batch = {'id': np.array([('synthetic_%06d' % b).encode() for b in range(B)])}
    plen = rng.randint(min(8, L), L + 1, size=B)
    ptoks = np.zeros((B, L), dtype=np.int32)
    program = np.zeros((B, V, L), dtype=np.float32)
    for b in range(B):
        n = plen[b]
        t = np.concatenate([[0, 1, 2], rng.randint(4, V, size=n - 4), [3]])
        ptoks[b, :n] = t
        program[b, t, np.arange(n)] = 1
    batch['program'], batch['program_tokens'] = program, ptoks
    batch['program_len'] = plen.reshape(B, 1).astype(np.float32)
    for pre, nd in (('', k), ('test_', tk)):
        lens = rng.randint(2, T + 1, size=(B, nd))
        fr = rng.randint(0, 256, size=(B, nd, T, cfg.h, cfg.w, cfg.depth))
        per = (rng.rand(B, nd, T, P) > 0.5).astype(np.float32)
        a_h = np.zeros((B, nd, T, A), dtype=np.float32)
        for b in range(B):
            max_a = lens[b].max() - 1
            padded = np.zeros((nd, max_a), dtype=np.int64)
            for i in range(nd):
                padded[i, :lens[b, i] - 1] = rng.randint(0, A - 1, lens[b, i] - 1)
            oh, _ = _action_onehot(padded, T, A - 1)
            a_h[b] = oh
        tmask = (np.arange(T)[None, None] < lens[..., None])
        fr = fr * tmask[..., None, None, None]
        per = per * tmask[..., None]
        batch[pre + 's_h'] = fr.astype(frames_dtype)
        batch[pre + 'per'] = per.astype(np.float32)
        batch[pre + 'a_h'] = a_h
        batch[pre + 'a_h_tokens'] = np.argmax(a_h, -1).astype(np.int32)
        batch[pre + 'demo_len'] = lens.astype(np.float32)
    return batch


def program_tokens_in_batch(batch):
    """Metric unit: real target tokens = sum_b program_len[b]
    (the normaliser of reference models/model_full.py:656-657)."""
    return int(np.asarray(batch['program_len']).sum())
