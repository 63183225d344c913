"""N>1 path on CPU: two gloo ranks each simulate their contiguous shard of realizations (device counter
RNG, emulated build) exactly as bench.py shards them; the gathered ensemble equals the single-rank one."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'oracle')); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
    import gsp_b200 as gsp, gsp_oracle as O
    from helpers import iso
    from bench import shard_range
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = gsp.Library(os.path.join(%(root)r, 'tests', 'emu', 'libgspb200_emu.so'))
    plan = gsp.FFTPlan(lib, iso(O.SPHERICAL, 1.0, 3.0, 2), (12, 10), [0.0, 0.0], [1.0, 1.0])
    R = 7
    r0, r1 = shard_range(R, rank, world)
    Z = plan.sample(r1 - r0, None, seed=11, first_real=r0)
    parts = [None] * world
    dist.all_gather_object(parts, (r0, Z))
    t = torch.tensor([float(r1 - r0)])
    dist.all_reduce(t)
    if rank == 0:
        full = np.concatenate([z for _, z in sorted(parts, key=lambda p: p[0])])
        ref = plan.sample(R, None, seed=11, first_real=0)
        assert int(t.item()) == R
        assert np.array_equal(full, ref), 'sharded ensemble differs from the single-rank ensemble'
        print('OK')
    dist.destroy_process_group()
""")


def test_two_rank_sharding_gloo(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    from bench import shard_range

    for R in (1, 7, 64, 512):
        for world in (1, 2, 4, 8):
            spans = [shard_range(R, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
