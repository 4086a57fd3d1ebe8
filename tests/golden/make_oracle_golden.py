"""Generates tests/golden/oracle_golden.json: outputs of the CPU oracle (fp64) on seeded configs -
loss terms, gradient norms per variable group, a checksum of the flat gradient, a slice of the
program logits and the greedy program tokens.  The reference ships no golden vectors for this path
and cannot run offline (parity unpinned): these values pin the ORACLE (a change
of its arithmetic shows up in tests/test_oracle.py::test_oracle_reproduces_golden) and give the
GPU parity tests a committed fixture that does not depend on executing the oracle
(tests/test_gpu_parity.py::test_engine_matches_committed_golden).  Run from the repo root."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
sys.path.insert(0, os.path.join(HERE, '..'))

from demo2program_b200.config import karel_config
from demo2program_b200.manifest import build_manifests
from demo2program_b200.synthetic import make_batch

CASES = [('synthesis_baseline', 8, 2), ('summarizer', 3, 2), ('full', 4, 3)]


def oracle_case(model, B, k):
    from oracle.models import OracleTrainer
    cfg = karel_config(model, batch_size=B, k=k)
    pm, sm = build_manifests(cfg)
    p0, s0 = pm.init_flat(0), sm.init_flat(0)
    rs = np.random.RandomState(17)
    for e in pm:   # same non-trivial BN affine params / biases as tests/parity_util.py
        if e.name.endswith('/beta') or e.name.endswith('biases') or e.name.endswith('/bias'):
            p0[e.offset:e.offset + e.size] = rs.uniform(-0.1, 0.1, e.size)
        if e.name.endswith('/gamma'):
            p0[e.offset:e.offset + e.size] = rs.uniform(0.8, 1.2, e.size)
    batch = make_batch(cfg, seed=1)
    orc = OracleTrainer(cfg, p0, s0, dtype=torch.float64)
    loss, grad, out = orc.model.loss_and_grad(batch)
    g = grad.numpy().astype(np.float64)
    groups = {}
    for e in pm:
        top = e.name.split('/')[0]
        groups[top] = groups.get(top, 0.0) + float((g[e.offset:e.offset + e.size] ** 2).sum())
    w = np.cos(np.arange(g.size) * 0.37)         # fixed projection: a checksum that sees every element
    return {
        'model': model, 'B': B, 'k': k, 'loss': float(loss),
        'grad_norm': float(np.sqrt((g ** 2).sum())),
        'grad_group_sqnorm': {k_: v for k_, v in sorted(groups.items())},
        'grad_projection': float((g * w).sum()),
        'pred_program_slice': out['pred_program'].detach().numpy()[0, :6, :4].round(10).tolist(),
        'demo_h_summary_slice': out['demo_h_summary'].detach().numpy()[0, :6].round(10).tolist(),
    }


if __name__ == '__main__':
    res = [oracle_case(*c) for c in CASES]
    json.dump(res, open(os.path.join(HERE, 'oracle_golden.json'), 'w'), indent=1, sort_keys=True)
    for r in res:
        print(r['model'], r['loss'], r['grad_norm'], r['grad_projection'])
