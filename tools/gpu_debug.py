"""Developer script (run under gpurun): engine vs oracle, per-variable errors."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from demo2program_b200.config import karel_config  # noqa: E402
from parity_util import oracle_and_engine, rel_err, per_var_errors  # noqa: E402


def run(model, B, k, steps=2, use_graph=False):
    cfg = karel_config(model, batch_size=B, k=k)
    orc, eng, batch, pm, sm = oracle_and_engine(cfg, use_graph=use_graph)
    print('==== %s B=%d k=%d graph=%s' % (model, B, k, use_graph))
    for s in range(steps):
        loss_o, grad_o, out = orc.model.loss_and_grad(batch)
        eng.stage_batch(batch)
        eng.forward()
        eng.backward()
        torch.cuda.synchronize()
        loss_e = eng.loss.cpu().numpy()
        print('step %d loss oracle %.7f engine %.7f (prog %.6f act %.6f per %.6f)' % (
            s, loss_o, loss_e[0], loss_e[1], loss_e[2], loss_e[3]))
        print('  oracle parts: prog %.6f act %s per %s' % (
            float(out['program_loss']), out.get('avg_action_loss'), out.get('avg_per_loss')))
        print('  pred_program rel err %.3e' % rel_err(eng.pred_program().cpu().numpy(),
                                                      out['pred_program'].detach().numpy()))
        print('  dsum_h rel err %.3e' % rel_err(eng.dsum_h.cpu().numpy(),
                                                out['demo_h_summary'].detach().numpy()))
        errs = per_var_errors(pm, eng.grads.cpu().numpy(), grad_o.numpy())
        for name, e in errs.items():
            flag = '' if e < 1e-3 else '   <<<<<<'
            print('  grad %-75s %.3e%s' % (name, e, flag))
        # advance both
        orc2_loss, norm_o, _ = orc.train_step(batch)
        eng.optimizer_step()
        torch.cuda.synchronize()
        print('  norm oracle %.6f engine %.6f' % (norm_o, eng.global_norm()))
        print('  params rel err after step: %.3e' % rel_err(eng.params.cpu().numpy(),
                                                          orc.model.flat.detach().numpy()))
        print('  state rel err after step: %.3e' % rel_err(eng.state.cpu().numpy(),
                                                         orc.model.state.numpy()))


if __name__ == '__main__':
    run('synthesis_baseline', 3, 2)
    run('summarizer', 3, 2)
    run('full', 4, 3)
    t = time.time()
    cfg = karel_config('full', batch_size=32, k=10)
    from demo2program_b200.engine import Engine
    from demo2program_b200.synthetic import make_batch
    eng = Engine(cfg, use_graph=True)
    batch = make_batch(cfg, seed=3)
    for i in range(5):
        print('C2 loss', eng.train_step(batch))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        eng.train_step_device(True)
    e1.record()
    torch.cuda.synchronize()
    print('C2 ms/step (graph replay):', e0.elapsed_time(e1) / 10, 'launches/step', eng.launches_per_step)
