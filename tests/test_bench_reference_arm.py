"""`bench.py --impl reference`: the reference arm of the bench contract runs without a GPU (it times the unmodified
reference from oracle/_ref — or the oracle port when that library is absent — on a bounded pixel sample), prints
exactly one JSON line, and under torchrun only rank 0 prints."""
import json
import subprocess
import sys

from conftest import REPO


def _lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_line():
    p = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--workload", "config", "--ref-step-seconds", "0.2"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    (line,) = _lines(p.stdout)
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "Mrays/s"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    # beside the single-threaded arm: the same sample through one independent reference process per host core
    ac = line["all_cores"]
    if line["cpu_baseline"]["kind"] == "reference":
        assert ac["cores"] >= 1 and ac["value"] > 0 and ac["rays"] > 0 and "independent processes" in ac["sample"]
    else:
        assert ac is None


def test_reference_arm_under_torchrun_only_rank0_prints():
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(REPO / "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "config", "--ref-step-seconds", "0.2"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    (line,) = _lines(p.stdout)
    assert line["impl"] == "reference" and line["n_gpus"] == 2
