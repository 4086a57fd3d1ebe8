"""Developer tool (GPU): persistent vs per-step recurrence on small Karel batches for several seeds
(how close the two land: loss difference, worst gradient difference and the variable it is in)."""
import sys
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np
from demo2program_b200.config import karel_config
from demo2program_b200.manifest import build_manifests
from demo2program_b200.synthetic import make_batch
from test_gpu_parity import _run_variant

cfg = karel_config('full', batch_size=4, k=3)
pm, _ = build_manifests(cfg)
for seed in (9, 1, 2, 3, 4):
    batch = make_batch(cfg, seed=seed)
    a = _run_variant(cfg, batch, 1, 0)
    b = _run_variant(cfg, batch, 1, 1)
    d = np.abs(a['grads'].astype(np.float64) - b['grads'])
    den = np.abs(a['grads']).max()
    i = int(d.argmax())
    name = [e.name for e in pm if e.offset <= i < e.offset + e.size][0]
    print('seed %d: dloss %.2e  grads rel %.2e (%s, |g|max %.2e)  saved equal %s' % (
        seed, abs(float(a['loss'][0] - b['loss'][0])), d.max() / den, name, den, np.array_equal(a['saved'], b['saved'])))
