import sys

import numpy as np
import torch.distributed as dist

from oracle import c_oracle as co
from spicey_b200 import workloads as w
from spicey_b200.parsing import parse_netlist
from spicey_b200.sharding import shard_range, sharded_sweep

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
ck = parse_netlist(w.rc_ladder(8, ppd=5))
freqs = np.logspace(0, 5, 37)
full = co.ac_solve(ck, freqs)[0]


def solve(lo, hi):
    return co.ac_solve(ck, freqs[lo:hi])[0]


out = sharded_sweep(solve, len(freqs))
local = sharded_sweep(solve, len(freqs), gather=False)
lo, hi = shard_range(len(freqs), rank, world)
assert np.array_equal(local, full[lo:hi])
if rank == 0:
    assert out.shape == full.shape and np.array_equal(out, full)
    print("GATHER_OK")
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
