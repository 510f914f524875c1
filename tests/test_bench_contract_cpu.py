"""CPU: bench.py's contract pieces that need no GPU — the reference arm (CPU port of the M1 step) prints one JSON line
with the required keys, and non-zero ranks of a torchrun launch exit quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_cores_is_positive_and_bounded():
    sys.path.insert(0, ROOT)
    import bench
    n = bench.host_cores()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_json_line(tmp_path):
    """One bounded step of the CPU arm (batch 1 of the c2 workload here to keep the CPU suite short; the driver's run
    uses the GPU arm's batch 8): the reference's own modules when /root/reference or oracle/_ref is present."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-batch", "1"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"] == "samples/sec perceiver+gated-xattn fwd+bwd" and d["value"] > 0
    sys.path.insert(0, ROOT)
    from oracle import ref_shims
    want_kind = "reference" if ref_shims.reference_available() else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["per_gpu_batch"] == 1 and d["cpu_baseline"]["batch"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["steps"] == 1


def test_allreduce_alone_record(monkeypatch):
    """bench.allreduce_alone (N > 1 only): field names and the bus-bandwidth formula, with CUDA events and the process
    group replaced by stand-ins (the real thing needs NCCL)."""
    import torch
    import torch.distributed as dist
    import bench

    class _Ev:
        def __init__(self, enable_timing=True):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 50.0                     # ms for the 10 iterations

    class _Flat:
        calls = 0

        def all_reduce(self):
            _Flat.calls += 1

        def comm_nbytes(self):
            return 2_000_000_000

    class _HP:
        flat = _Flat()

    monkeypatch.setattr(torch.cuda, "Event", _Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None: None)
    r = bench.allreduce_alone(_HP(), "cpu", 8, iters=10)
    assert _Flat.calls == 12                                    # 2 warm-up + 10 timed
    assert r["ms"] == 5.0 and r["wire_bytes"] == 2_000_000_000 and r["iters"] == 10
    assert abs(r["busbw_gbs"] - 2 * 7 / 8 * 2e9 / 5e-3 / 1e9) < 0.1      # 700 GB/s
