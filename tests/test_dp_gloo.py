"""Multi-process (gloo, world_size 2, CPU) test of the data-parallel plumbing: sequence sharding + the single
trajectory all-gather at the end of a run."""
import os
import subprocess
import sys
import textwrap

import pytest

from util import ROOT

from uvltrack_b200 import dp


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from uvltrack_b200 import dp
    rank, world, _ = dp.init_process_group("gloo")
    n_seq, T = 5, 7                                   # ragged: rank 0 owns 3 sequences, rank 1 owns 2
    lo, hi = dp.shard_range(n_seq, rank, world)
    full = torch.arange(n_seq * T * 4, dtype=torch.float32).reshape(n_seq, T, 4)
    calls = []
    for name in ("all_gather", "all_gather_into_tensor", "all_reduce", "broadcast", "all_to_all"):
        orig = getattr(dist, name)
        setattr(dist, name, (lambda o, n: (lambda *a, **k: (calls.append(n), o(*a, **k))[1]))(orig, name))
    got = dp.gather_trajectories(full[lo:hi].clone(), n_sequences=n_seq)
    assert torch.equal(got, full), (rank, got.shape)
    assert calls == ["all_gather"], calls            # ONE collective: shard sizes come from shard_range, not a reduce
    # equal shards (the benchmark's layout) and the warm-up gather before a timed region
    del calls[:]
    eq = torch.arange(2 * 3 * T * 4, dtype=torch.float32).reshape(6, T, 4)
    dp.warmup_gather(eq[:3])
    got = dp.gather_trajectories(eq[3 * rank: 3 * rank + 3].clone(), n_sequences=6)
    assert torch.equal(got, eq) and calls == ["all_gather", "all_gather"], calls
    try:
        dp.gather_trajectories(eq[:2].clone(), n_sequences=6)   # a shard that disagrees with shard_range is an error
        raise SystemExit("expected RuntimeError")
    except RuntimeError:
        pass
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_gather_trajectories_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_single_process_is_identity():
    import torch

    x = torch.randn(3, 4, 4)
    assert dp.gather_trajectories(x) is x
