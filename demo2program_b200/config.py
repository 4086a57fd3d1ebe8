"""Model/data configuration for the demo2program hot path.

Mirrors the fields the reference `Model.__init__` copies out of the argparse
namespace (reference models/model_full.py:30-57) plus the dims `trainer.py`
sniffs from the dataset (reference trainer.py:312-335).
"""
from dataclasses import dataclass, field
from typing import List, Optional

MODEL_NAMES = ('synthesis_baseline', 'induction_baseline', 'summarizer', 'full')


@dataclass
class D2PConfig:
    # CLI surface (reference trainer.py:247-289)
    model: str = 'full'
    dataset_type: str = 'karel'
    batch_size: int = 32
    learning_rate: float = 1e-3
    lr_weight_decay: bool = False
    encoder_rnn_type: str = 'lstm'
    num_lstm_cell_units: int = 512
    demo_aggregation: str = 'avgpool'
    scheduled_sampling: bool = False
    scheduled_sampling_decay_steps: int = 20000
    # dataset-derived dims (reference trainer.py:312-321)
    dim_program_token: int = 50
    max_program_len: int = 50
    max_demo_len: int = 20
    k: int = 10
    test_k: int = 5
    h: int = 8
    w: int = 8
    depth: int = 16
    action_space: int = 6
    per_dim: int = 5
    dsl_type: str = 'prob'
    env_type: Optional[str] = None
    vizdoom_pos_keys: List[str] = field(default_factory=list)
    vizdoom_max_init_pos_len: int = -1
    perception_type: str = ''
    level: Optional[str] = None
    # induction-only fields nobody sets in the reference CLI
    # (reference models/baselines/model_induction.py:194-212)
    pixel_input: bool = False
    attn_type: str = 'luong'
    state_encoder_fc: bool = False
    concat_state_feature_direct_prediction: bool = False
    stack_subsequent_state: bool = False
    # end token of greedy program decode = vocab['m)'] (model_full.py:428)
    program_end_token: int = 3

    def validate(self):
        if self.model not in MODEL_NAMES:
            raise ValueError(self.model)
        if self.dataset_type not in ('karel', 'vizdoom'):
            raise ValueError(self.dataset_type)
        if self.encoder_rnn_type != 'lstm':
            # reference models/model_full.py:247-258: rnn/gru crash on
            # cell_state.h; reject up front.
            raise ValueError('Unknown encoder rnn type: only lstm is '
                             'supported (got %r)' % (self.encoder_rnn_type,))
        return self

    # conv stack: (cin, cout) per layer; reference model_full.py:216-231
    def conv_channels(self):
        chans = [16, 32, 48]
        if self.dataset_type == 'vizdoom':
            chans += [48, 48]
        out, cin = [], self.depth
        for c in chans:
            out.append((cin, c))
            cin = c
        return out

    def conv_geometry(self):
        """[(ih, iw, cin, oh, ow, cout, pad_top, pad_left)] for 3x3 stride-2 SAME
        (TF rule, SURVEY A.1: pad_before = pad_total // 2)."""
        geo = []
        ih, iw = self.h, self.w
        for cin, cout in self.conv_channels():
            oh, ow = (ih + 1) // 2, (iw + 1) // 2
            pt = max((oh - 1) * 2 + 3 - ih, 0) // 2
            pl = max((ow - 1) * 2 + 3 - iw, 0) // 2
            geo.append((ih, iw, cin, oh, ow, cout, pt, pl))
            ih, iw = oh, ow
        return geo

    def feature_dim(self):
        g = self.conv_geometry()[-1]
        return g[3] * g[4] * g[5]


def karel_config(model='full', batch_size=32, k=10, **kw):
    return D2PConfig(model=model, dataset_type='karel', batch_size=batch_size,
                     k=k, **kw).validate()


def vizdoom_config(model='full', batch_size=32, k=10, **kw):
    base = dict(dim_program_token=42, max_program_len=32, max_demo_len=20,
                test_k=10, h=80, w=80, depth=3, action_space=12, per_dim=6,
                dsl_type='vizdoom_default', env_type='vizdoom_default',
                perception_type='simple')
    base.update(kw)
    return D2PConfig(model=model, dataset_type='vizdoom', batch_size=batch_size,
                     k=k, **base).validate()
