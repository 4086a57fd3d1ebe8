"""Shared helpers for the GPU-vs-oracle parity tests."""
import numpy as np
import torch

from demo2program_b200.manifest import build_manifests
from demo2program_b200.synthetic import make_batch


def oracle_and_engine(cfg, seed=0, batch_seed=1, dtype=torch.float64, **engine_kw):
    from oracle.models import OracleTrainer
    from demo2program_b200.engine import Engine
    pm, sm = build_manifests(cfg)
    p0, s0 = pm.init_flat(seed), sm.init_flat(seed)
    # non-trivial BN affine params / biases so their gradients are exercised
    rs = np.random.RandomState(seed + 17)
    for e in pm:
        if e.name.endswith('/beta') or e.name.endswith('biases') or e.name.endswith('/bias'):
            p0[e.offset:e.offset + e.size] = rs.uniform(-0.1, 0.1, e.size)
        if e.name.endswith('/gamma'):
            p0[e.offset:e.offset + e.size] = rs.uniform(0.8, 1.2, e.size)
    batch = make_batch(cfg, seed=batch_seed)
    orc = OracleTrainer(cfg, p0, s0, dtype=dtype)
    eng = Engine(cfg, flat_params=p0, flat_state=s0, **engine_kw)
    return orc, eng, batch, pm, sm


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(np.abs(b).max(), 1e-12)
    return float(np.abs(a - b).max() / den)


def per_var_errors(pm, flat_a, flat_b):
    out = {}
    for e in pm:
        a = flat_a[e.offset:e.offset + e.size]
        b = flat_b[e.offset:e.offset + e.size]
        out[e.name] = rel_err(a, b)
    return out
