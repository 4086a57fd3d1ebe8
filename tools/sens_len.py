"""Developer tool (GPU): sensitivity of the C2 step time to the program-decoder length
(how much of the critical path the 50 dependent decoder steps are)."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch
from tools.component_bench import graph_step_ms

for L in (50, 26, 12):
    cfg = karel_config('full', batch_size=32, k=10, max_program_len=L)
    eng = Engine(cfg, use_graph=True)
    eng.stage_batch(make_batch(cfg, seed=123))
    print('max_program_len', L, 'graph step ms', round(graph_step_ms(eng), 3), flush=True)
    del eng
    torch.cuda.empty_cache()
