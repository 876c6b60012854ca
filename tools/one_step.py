"""Two GEECO-F bf16 train steps at batch 64 (the bench workload) and nothing else: the target of ncu captures
(`ncu --set full -k regex:<kernel> -s <skip> -c 1 python tools/one_step.py`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geeco_b200 import create_e2evmc_config  # noqa: E402
from geeco_b200.data import synthetic_batch  # noqa: E402
from geeco_b200.engine import Engine  # noqa: E402

N = int(os.environ.get('GEECO_BATCH', '64'))
dev = torch.device('cuda:0')
cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=N))
eng = Engine(cfg, batch_size=N, precision='bf16', training=True, device=dev)
eng.init_params(seed=0)
f, l = synthetic_batch(N, seed=1, structured=False)
b = {k: torch.from_numpy(v).to(dev) for k, v in f.items() if k != 'step'}
b['cmd'] = torch.from_numpy(l['cmd']).to(dev)
for _ in range(int(os.environ.get('GEECO_STEPS', '2'))):
  eng.train_step(b, b)
torch.cuda.synchronize()
print('loss', float(eng.out_losses[5]))
