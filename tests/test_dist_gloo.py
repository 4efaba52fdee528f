"""N>1 host logic on CPU: world_size-2 gloo processes shard 256 images, derive world-size-independent
simplex seeds and reduce timings -- no data-path collective (detection is embarrassingly parallel)."""
import os
import subprocess
import sys
import textwrap

from tests.conftest import ROOT

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    import torch
    from anoddpm_b200 import dist as D
    rank, world, local = D.init_from_env(backend="gloo")
    lo, hi = D.shard_range(256, rank, world)
    seeds = D.seeds_for_step(1234, 5, range(lo, hi))
    mx = D.max_over_ranks(10.0 + rank, torch.device("cpu"))
    sm = D.sum_over_ranks(hi - lo, torch.device("cpu"))
    D.barrier()
    with open(os.path.join(os.environ["ADP_OUT"], f"rank{rank}.json"), "w") as f:   # (stdout of two ranks can interleave)
        json.dump({"rank": rank, "lo": lo, "hi": hi, "seeds": seeds[:4], "max": mx, "sum": sm}, f)
""") % ROOT


def test_two_rank_sharding_and_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, ADP_OUT=str(tmp_path)))
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    rows = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    assert [(r["lo"], r["hi"]) for r in rows] == [(0, 128), (128, 256)]
    assert all(r["max"] == 11.0 and r["sum"] == 256.0 for r in rows)
    # seeds depend on the global image index only, not on the world size / owner
    from anoddpm_b200.dist import seeds_for_step, shard_range
    assert rows[1]["seeds"] == seeds_for_step(1234, 5, range(128, 132))
    assert rows[0]["seeds"] == seeds_for_step(1234, 5, range(0, 4))
    assert [shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert all(-10000000000 <= s < 10000000000 and s != 0 for s in seeds_for_step(7, 0, range(100)))


REDUCE_WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    import torch
    from anoddpm_b200 import dist as D
    from anoddpm_b200.optim import GradReducer
    rank, world, local = D.init_from_env(backend="gloo")
    sizes = [100, 200, 300, 300, 100]
    offsets = [0, 100, 300, 600, 900]
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)      # rank r holds (r + 1) * i
    red = GradReducer(g, offsets, sizes, bucket_elems=350)
    assert red.buckets == [(600, 1000), (100, 600), (0, 100)], red.buckets
    launched = []
    red.mark_ready(900); launched.append(red.next_bucket)         # tail bucket not complete yet
    red.mark_ready(600); launched.append(red.next_bucket)         # bucket 0 goes out while "backward" continues
    red.mark_ready(300); launched.append(red.next_bucket)
    red.finish()                                                  # the rest + wait + divide by the world size
    # per-tensor completion in arbitrary order (autograd's post-accumulate hooks): a bucket goes out when ALL its tensors are final
    g2 = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    red2 = GradReducer(g2, offsets, sizes, bucket_elems=350)
    red2.track(offsets, sizes)
    seq = []
    for i in (3, 4, 1, 2, 0):                                     # tensor 4 = [900, 1000), 3 = [600, 900), ...
        red2.tensor_done(i)
        seq.append(red2.next_bucket)
    red2.finish()
    with open(os.path.join(os.environ["ADP_OUT"], f"red{rank}.json"), "w") as f:
        json.dump({"launched": launched, "g": g.tolist(), "seq": seq, "g2": g2.tolist()}, f)
    D.shutdown()
""") % ROOT


def test_two_rank_bucketed_gradient_average(tmp_path):
    """training row of SURVEY 8(e): the flat gradient is averaged bucket by bucket, tail (last layers) first; both ranks
    end with bit-identical means."""
    script = tmp_path / "r.py"
    script.write_text(REDUCE_WORKER)
    port = 31500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, ADP_OUT=str(tmp_path)))
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    rows = [json.load(open(tmp_path / f"red{r}.json")) for r in range(2)]
    assert rows[0]["launched"] == rows[1]["launched"] == [0, 1, 1]
    assert rows[0]["g"] == rows[1]["g"] == [1.5 * i for i in range(1000)]
    assert rows[0]["seq"] == rows[1]["seq"] == [0, 1, 1, 2, 3]    # bucket (600,1000) after tensors 3+4, (100,600) after 1+2, (0,100) last
    assert rows[0]["g2"] == rows[1]["g2"] == [1.5 * i for i in range(1000)]
