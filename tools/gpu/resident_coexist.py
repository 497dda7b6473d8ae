#!/usr/bin/env python
"""tools/gpu/resident_coexist.py -- (GPU box) a large device-resident launch right after per-block host calls (the resident CTA of
the per-block path still on the chip) against the same launch on a quiet context.  Prints ms per launch for both."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import doppler_b200
from doppler_b200 import F32, I16

n = 256_000_000
m = doppler_b200.Mixer(0)
m.tune(resident_idle_us=5_000_000)           # (stays for seconds unless told to leave)
d_in = torch.randint(-30000, 30000, (2 * n,), dtype=torch.int16, device="cuda")
d_out = torch.empty(2 * n, dtype=torch.int16, device="cuda")
blk = np.zeros(4096, np.int16).view(np.uint8)
st = torch.cuda.Stream()


def big():
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(st)
    m.mix_dev(d_in.data_ptr(), 4 * n, I16, I16, -15000.0, 256000, 0, d_out.data_ptr(), 4 * n, stream=st.cuda_stream)
    ev1.record(st)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1)


for _ in range(3):
    big()
quiet = [big() for _ in range(5)]
after = []
for _ in range(5):
    m.mix(blk, I16, I16, 5000.0, 1_024_000)  # a per-block call: the resident kernel is (back) on the chip
    after.append(big())
print({"launch_ms_quiet_context": [round(x, 3) for x in quiet], "launch_ms_right_after_a_per_block_call": [round(x, 3) for x in after]})
