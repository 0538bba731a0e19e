"""world_size-2 gloo run of the multi-GPU host logic (scene sharding + throughput reduction), no GPU."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys; sys.path.insert(0, {ROOT!r})
        from scannet_b200 import dist
        g = dist.Group("gloo")
        seed = dist.scene_seed_for_rank(g.rank, 10)
        g.barrier()
        frames, ms = g.reduce_throughput(1000 + g.rank, 50.0 + 25.0 * g.rank)    # rank 1 is the slow one
        print(f"rank={{g.rank}} world={{g.world}} seed={{seed}} frames={{frames}} ms={{ms}}", flush=True)
        g.close()
    """))
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0)); port = so.getsockname()[1]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    import re                                   # the two ranks share one pipe: their lines can interleave
    lines = sorted(re.findall(r"rank=\d world=\d seed=\d+ frames=\d+ ms=[0-9.]+?(?=rank=|\s|$)", r.stdout))
    assert lines == ["rank=0 world=2 seed=10 frames=2001 ms=75.0", "rank=1 world=2 seed=11 frames=2001 ms=75.0"]
