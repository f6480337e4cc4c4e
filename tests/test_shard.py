"""Host-side logic of reference-view sharding: partitioning, and the N>1 join path over
gloo with world_size 2 on CPU (the GPU box uses the same code over NCCL)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from deep3d_aerial_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["round_robin", "contiguous"])
@pytest.mark.parametrize("n,world", [(64, 1), (64, 8), (10, 4), (3, 8), (0, 2)])
def test_partition_covers_every_view_exactly_once(mode, n, world):
    ids = list(range(100, 100 + n))
    parts = [shard.partition(ids, world, r, mode) for r in range(world)]
    assert sorted(x for p in parts for x in p) == ids
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1
    if mode == "contiguous":
        assert [x for p in parts for x in p] == ids


def test_partition_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard.partition([1, 2], 2, 2)
    with pytest.raises(ValueError):
        shard.partition([1, 2], 2, 0, mode="zigzag")


def test_single_process_joins_are_identity():
    assert shard.world() == (0, 1, 0) or "RANK" in os.environ
    assert shard.join_max(3.5) == 3.5
    assert shard.join_sum(2.0) == 2.0
    assert shard.gather_objects({"a": 1}) == [{"a": 1}]
    out = shard.run_block([5, 6, 7], lambda v: v * v)
    assert out == {5: 25, 6: 36, 7: 49}
    with pytest.raises(RuntimeError, match="reference view 6 failed"):
        shard.run_block([5, 6], lambda v: 1 // (6 - v))


WORKER = textwrap.dedent("""
    import json, sys
    sys.path.insert(0, %r)
    from deep3d_aerial_b200 import shard
    rank, world, local = shard.init(backend="gloo")
    views = list(range(11))
    mine = shard.run_block(views, lambda v: {"view": v, "voxels": 1000 + v, "rank": rank})
    shard.barrier()
    ms = shard.join_max(10.0 * (rank + 1))          # slowest rank defines the job time
    total = shard.join_sum(sum(r["voxels"] for r in mine.values()))
    everyone = shard.gather_objects(sorted(mine))
    if rank == 0:
        print(json.dumps({"world": world, "ms": ms, "total": total, "owned": everyone}))
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [x for x in r.stdout.splitlines() if x.startswith("{")][-1]
    js = json.loads(line)
    assert js["world"] == 2
    assert js["ms"] == 20.0
    assert js["total"] == sum(1000 + v for v in range(11))
    assert js["owned"] == [[0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9]]


def test_reference_arm_rank_gate(tmp_path):
    """bench.py --impl reference: ranks other than 0 exit 0 without doing work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
