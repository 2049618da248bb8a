"""Multi-process GPU test of the frame-sharded path (BASELINE configs[2]): two ranks run their blocks of
frames through the real engine, one collective gathers the pose rows, and the gathered tensor must be bit-identical to a
single-rank run of all frames (tests/mp_sharded_worker.py).  With fewer GPUs than ranks both ranks share cuda:0."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("num_frames", [8, 7])          # even blocks, ragged blocks
def test_sharded_forward_two_ranks_equals_single_rank(built_library, num_frames):
    world, port = 2, _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_sharded_worker.py"), str(num_frames)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out[-3000:])
    assert "gathered == single-rank: True" in outs[0]
