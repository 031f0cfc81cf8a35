"""N>1 plumbing of bench.py on the CPU: two gloo ranks each generate and slice-parse their own shard of
closed GOPs (weak scaling, no collective on the data path) and reduce counts / times."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--cpu-dryrun", "--workload", "720p420_ipb", "--steps", "1", "--gpus", str(nproc)]
    if nproc == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--cpu-dryrun", "--workload", "720p420_ipb", "--steps", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout      # rank 0 alone prints
    return json.loads(lines[0])


def test_two_ranks_decode_two_distinct_shards():
    one = run(1)
    two = run(2)
    assert one["frames_per_step"] == 30 and two["frames_per_step"] == 60
    assert two["n_ranks"] == 2 and two["scaling"] == "weak"
    # rank 1's shard is a different stream (seed = stream_id*1000 + config_id), not a copy of rank 0's
    assert two["coef_records_all_ranks"] != 2 * one["coef_records_all_ranks"]
    assert two["coef_records_all_ranks"] > one["coef_records_all_ranks"]
