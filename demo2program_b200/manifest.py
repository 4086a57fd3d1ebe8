"""Flat parameter buffer + name->(offset, shape) manifest.

One contiguous fp32 buffer holds every trainable variable in TF variable-name
order.  The same layout is used for the gradient buffer,
the Adam slots, the NCCL all-reduce unit and the checkpoint format.  A second
small buffer ("state") holds the non-trainable BatchNorm moving statistics.

Initialisers restate TF-1.3 defaults at the reference call sites:
  slim.conv2d / slim.fully_connected  -> xavier-uniform weights, zero biases
      (reference models/ops.py:30,152)
  BasicLSTMCell                       -> glorot-uniform kernel, zero bias
      (reference models/model_full.py:244)
  Token_Embedding                     -> U(-0.01, 0.01) (model_full.py:288-291)
  output_projection Dense(no bias)    -> glorot-uniform (model_full.py:463)
  BatchNorm                           -> gamma=1, beta=0, moving_mean=0,
      moving_variance=1 (models/ops.py:20-23)
"""
from collections import OrderedDict
import numpy as np


class Entry:
    __slots__ = ('name', 'shape', 'offset', 'size', 'init', 'fans')

    def __init__(self, name, shape, offset, init, fans=None):
        self.name = name
        self.shape = tuple(int(s) for s in shape)
        self.offset = int(offset)
        self.size = int(np.prod(self.shape))
        self.init = init
        self.fans = fans

    def __repr__(self):
        return 'Entry(%s, %s, @%d)' % (self.name, self.shape, self.offset)


class Manifest:
    """Ordered name -> Entry map over one flat buffer."""

    def __init__(self, align=4):
        self.entries = OrderedDict()
        self.total = 0
        self.align = align  # elements; keeps every tensor 16-byte aligned

    def add(self, name, shape, init, fans=None):
        if name in self.entries:
            raise KeyError('duplicate variable ' + name)
        e = Entry(name, shape, self.total, init, fans)
        self.entries[name] = e
        self.total += (e.size + self.align - 1) // self.align * self.align
        return e

    def __getitem__(self, name):
        return self.entries[name]

    def __contains__(self, name):
        return name in self.entries

    def __iter__(self):
        return iter(self.entries.values())

    def num_params(self):
        return sum(e.size for e in self)

    def view(self, flat, name):
        e = self.entries[name]
        return flat[e.offset:e.offset + e.size].reshape(e.shape)

    def init_flat(self, seed=0, dtype=np.float32):
        rs = np.random.RandomState(seed)
        flat = np.zeros(self.total, dtype=dtype)
        for e in self:
            if e.init == 'zeros':
                continue
            if e.init == 'ones':
                v = np.ones(e.size)
            elif e.init == 'glorot':
                fan_in, fan_out = e.fans
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                v = rs.uniform(-lim, lim, size=e.size)
            elif e.init == 'emb':
                v = rs.uniform(-0.01, 0.01, size=e.size)
            else:
                raise ValueError(e.init)
            flat[e.offset:e.offset + e.size] = v.astype(dtype)
        return flat


def _bn(m, s, scope, c):
    m.add(scope + '/BatchNorm/beta', (c,), 'zeros')
    m.add(scope + '/BatchNorm/gamma', (c,), 'ones')
    s.add(scope + '/BatchNorm/moving_mean', (c,), 'zeros')
    s.add(scope + '/BatchNorm/moving_variance', (c,), 'ones')


def _conv(m, s, scope, cin, cout):
    m.add(scope + '/Conv/weights', (3, 3, cin, cout), 'glorot',
          (9 * cin, 9 * cout))
    m.add(scope + '/Conv/biases', (cout,), 'zeros')
    _bn(m, s, scope + '/bn_act', cout)


def _fc(m, s, scope, cin, cout, bn=True):
    m.add(scope + '/fully_connected/weights', (cin, cout), 'glorot', (cin, cout))
    m.add(scope + '/fully_connected/biases', (cout,), 'zeros')
    if bn:
        _bn(m, s, scope + '/bn_act', cout)


def _lstm(m, scope, n_in, H):
    m.add(scope + '/basic_lstm_cell/kernel', (n_in + H, 4 * H), 'glorot',
          (n_in + H, 4 * H))
    m.add(scope + '/basic_lstm_cell/bias', (4 * H,), 'zeros')


def _rn_pool(m, s, scope, H):
    _fc(m, s, scope + '/rn_pool/fc1', 2 * H, H)
    _fc(m, s, scope + '/rn_pool/fc2', H, H)


def _token_decoder(m, scope, vocab, H):
    m.add(scope + '/Token_Embedding/embedding_map', (vocab + 1, H), 'emb')
    _lstm(m, scope + '/dynamic_decoder', H, H)
    m.add(scope + '/dynamic_decoder/output_projection/kernel', (H, vocab),
          'glorot', (H, vocab))


def build_manifests(cfg):
    """Return (params, state) manifests for cfg.model.

    Scope names follow the variable_scope nesting of the reference:
      full        models/model_full.py:216-599
      summarizer  models/baselines/model_summarizer.py
      synthesis   models/baselines/model_synthesis.py
      induction   models/baselines/model_induction.py:399-709
    """
    m, s = Manifest(), Manifest()
    H = cfg.num_lstm_cell_units
    enc = 'Demo_Encoder'
    for li, (cin, cout) in enumerate(cfg.conv_channels()):
        _conv(m, s, '%s/State_Encoder/conv%d' % (enc, li + 1), cin, cout)
    F = cfg.feature_dim()
    if cfg.model == 'induction_baseline':
        F = F + cfg.per_dim  # conv feats ++ perception vector (model_induction.py:399-474)
    _lstm(m, enc + '/rnn', F, H)
    if cfg.model in ('full', 'summarizer'):
        _lstm(m, 'SecondPathEncoder/rnn', H, H)
        _rn_pool(m, s, 'demo_h_summary', H)
        _rn_pool(m, s, 'demo_c_summary', H)
    if cfg.model in ('full', 'summarizer', 'synthesis_baseline'):
        _token_decoder(m, 'Program_Decoder', cfg.dim_program_token, H)
    if cfg.model == 'full':
        _token_decoder(m, 'Action_Decoder', cfg.action_space, H)
        _fc(m, s, 'Per_Decoder/Per_Encoder/fc2', cfg.per_dim, H)
        _lstm(m, 'Per_Decoder/dynamic_decoder', H, H)
        m.add('Per_Decoder/dynamic_decoder/output_projection/kernel',
              (H, cfg.per_dim), 'glorot', (H, cfg.per_dim))
    if cfg.model == 'induction_baseline':
        # scopes: AttnMechanism (memory_layer), Manipulation (decoder), reference
        # models/baselines/model_induction.py:638-709
        A = cfg.action_space
        m.add('AttnMechanism/memory_layer/kernel', (H, H), 'glorot', (H, H))
        sc = 'Manipulation'
        m.add(sc + '/Token_Embedding/embedding_map', (A + 1, H), 'emb')
        w = sc + '/dynamic_decoder/pooling_attention_wrapper/'
        # cell input = [emb ; attention] (2H) ++ h (H)
        m.add(w + 'basic_lstm_cell/kernel', (3 * H, 4 * H), 'glorot', (3 * H, 4 * H))
        m.add(w + 'basic_lstm_cell/bias', (4 * H,), 'zeros')
        m.add(w + 'attention_layer/kernel', (2 * H, H), 'glorot', (2 * H, H))
        m.add(sc + '/dynamic_decoder/output_projection/kernel', (H, A), 'glorot', (H, A))
    return m, s
