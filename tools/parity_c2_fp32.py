"""Developer tool (GPU): C2-size gradient parity of the exact-fp32 SIMT engine (use_tc=False) and of
the default tensor-core engine against the fp64 oracle, variable by variable - separates the bf16x3
rounding of ill-conditioned sums (rn_pool) from algorithmic differences."""
import json
import sys

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np
import torch

from demo2program_b200.config import karel_config
from parity_util import oracle_and_engine


def main():
    cfg = karel_config('full', batch_size=32, k=10)
    out = {}
    go = None
    for name, kw in (('tc', dict(use_graph=False)), ('fp32', dict(use_graph=False, use_tc=False))):
        orc, eng, batch, pm, sm = oracle_and_engine(cfg, **kw)
        if go is None:
            loss_o, grad_o, _ = orc.model.loss_and_grad(batch)
            go = grad_o.numpy()
        eng.stage_batch(batch)
        eng.forward()
        eng.backward()
        torch.cuda.synchronize()
        g = eng.grads.cpu().numpy()
        gmax = np.abs(go).max()
        tab = {}
        for e in pm:
            a, b = g[e.offset:e.offset + e.size], go[e.offset:e.offset + e.size]
            tab[e.name] = (float(np.abs(a - b).max()), float(np.abs(b).max()))
        out[name] = {'loss_err': abs(float(eng.loss[0]) - loss_o), 'gmax': float(gmax),
                     'l2_rel': float(np.linalg.norm(g - go) / np.linalg.norm(go)), 'vars': tab}
        del eng
    print('loss err tc %.2e fp32 %.2e; flat-gradient relative L2 error tc %.2e fp32 %.2e; gmax %.3e' % (
        out['tc']['loss_err'], out['fp32']['loss_err'], out['tc']['l2_rel'], out['fp32']['l2_rel'], out['tc']['gmax']))
    rows = sorted(out['tc']['vars'], key=lambda n: -out['tc']['vars'][n][0] / (out['tc']['vars'][n][1] + 1e-30))
    for n in rows[:20]:
        dt, bm = out['tc']['vars'][n]
        df, _ = out['fp32']['vars'][n]
        print('%-70s max|g| %.2e  err tc %.2e (%.1e rel)  err fp32 %.2e (%.1e rel)' % (n, bm, dt, dt / (bm + 1e-30), df, df / (bm + 1e-30)))
    json.dump(out, open('gpurun_out/parity_c2_tc_vs_fp32.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
