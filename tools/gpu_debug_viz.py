import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from demo2program_b200.config import vizdoom_config
from parity_util import oracle_and_engine, rel_err, per_var_errors
for use_tc in (True, False):
    cfg = vizdoom_config('full', batch_size=2, k=2, max_demo_len=3, test_k=2, max_program_len=8)
    orc, eng, batch, pm, sm = oracle_and_engine(cfg, use_graph=False, use_tc=use_tc)
    loss_o, grad_o, out = orc.model.loss_and_grad(batch)
    eng.stage_batch(batch); eng.forward(); eng.backward(); torch.cuda.synchronize()
    print('use_tc', use_tc, 'loss', loss_o, float(eng.loss[0]))
    errs = per_var_errors(pm, eng.grads.cpu().numpy(), grad_o.numpy())
    for n, e in errs.items():
        if 'State_Encoder' in n or e > 1e-4:
            print('  %-70s %.3e' % (n, e))
