"""Generates tests/golden/karel_states_golden.json from the REFERENCE's state sampler,
karel_env/generator.py `KarelStateGenerator.generate_single_state` (legacy MT19937 RandomState
streams are reproducible under NumPy 2.x), imported from /root/reference with stand-ins for the
packages its module header imports but the sampler does not use (h5py, progressbar, colorlog, ply).
demo2program_b200.synthetic.KarelSim restates the sampler; tests/test_oracle.py checks it."""
import json
import logging
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
SEEDS = [0, 1, 7, 123, 2024, 99991]


def main():
    import make_karel_dsl_golden as M
    M.load_reference()                       # ply shim + sys.path for karel_env / karel_env/dsl
    for name in ('h5py', 'progressbar'):
        sys.modules.setdefault(name, types.ModuleType(name))
    cl = types.ModuleType('colorlog')
    cl.ColoredFormatter = lambda *a, **k: logging.Formatter('%(message)s')
    sys.modules['colorlog'] = cl
    cwd = os.getcwd()
    os.chdir('/tmp')
    try:
        import generator as ref
    finally:
        os.chdir(cwd)
    out = {}
    for seed in SEEDS:
        g = ref.KarelStateGenerator(seed=seed)
        states = []
        for _ in range(3):                   # consecutive draws from one stream
            s, y, x, walls, markers = g.generate_single_state(8, 8, 0.1)
            states.append({'bits': np.packbits(np.asarray(s, np.uint8)).tolist(), 'y': int(y), 'x': int(x),
                           'walls': int(walls), 'markers': int(markers)})
        out[str(seed)] = states
    json.dump(out, open(os.path.join(HERE, 'karel_states_golden.json'), 'w'), separators=(',', ':'))
    print({k: [(s['y'], s['x'], s['walls']) for s in v] for k, v in out.items()})


if __name__ == '__main__':
    main()
