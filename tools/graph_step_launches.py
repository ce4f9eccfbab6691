#!/usr/bin/env python3
"""GPU, under `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv`: exactly ONE
replayed training step (three CUDA graphs + the bucketed AdamW launches) between cudaProfilerStart / Stop, so the launch
list holds the kernels of one step and nothing else.  usage: graph_step_launches.py [batch] [tf32|bf16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmfn_b200 import ops, synthetic  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from mmfn_b200.engine import BatchStager, TrainEngine  # noqa: E402
from mmfn_b200.model_rad import MMFN  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ops.set_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32")
dev = torch.device("cuda:0")
model = MMFN(GlobalConfig(), dev)
eng = TrainEngine(model)
hb = synthetic.synth_batch(B)
st = BatchStager(hb, dev)
st.stage(hb)
torch.cuda.synchronize()
eng.capture(st.dev_views)
for _ in range(2):
    eng.step_graph()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step_graph()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
