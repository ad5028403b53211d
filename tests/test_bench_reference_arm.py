"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one JSON line with
the contract's keys, printed by rank 0 only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-scale", "0.01", "--cpu-cores", "2"]


def _lines(out: str) -> list[dict]:
    return [json.loads(line) for line in out.splitlines() if line.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *ARGS], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = _lines(proc.stdout)
    assert len(lines) == 1
    line = lines[0]
    assert line["impl"] == "reference" and line["unit"] == "GCells/s" and line["value"] > 0
    assert line["metric"] == "allele-typing read x candidate GCells/s" and line["higher_is_better"] is True
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert line["cpu_baseline"]["cores"] == 2
    assert line["parity"]["genes"] == 34
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_reference_arm_under_torchrun_rank0_only():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "bench.py"),
                           "--gpus", "2", *ARGS], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = _lines(proc.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2


def test_reference_arm_worker_runs_the_imported_reference_without_the_cuda_library():
    """A worker of the CPU arm types its sample with the byte-compiled reference from oracle/_ref and never
    maps libgk_typing.so (the driver checks the arm's processes for exactly that)."""
    from oracle import ref_loader
    if not ref_loader.available():
        import pytest
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py where /root/reference exists)")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import bench\n"
        "cells, t_all, t_search, genes, ok = bench._cpu_type_sample((100, 0.01, 300, 'reference'))\n"
        "import graphkir.typing_mulit_allele as tma\n"
        "assert 'oracle/_ref' in tma.__file__.replace('\\\\', '/'), tma.__file__\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libgk_typing' not in maps and 'libcudart' not in maps, 'CUDA library mapped'\n"
        "print(cells, genes, ok)\n" % ROOT)
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    cells, genes, ok = proc.stdout.split()
    assert int(genes) == 17 and int(cells) > 0


def test_score_roofline_entries():
    """The roofline block of the GPU arm: the packed kernel is reported against one issue slot per cell
    (148 SM x 128 lanes), with the measured peak of its instruction mix and round 1's ALU-only denominator
    beside it; the FP32 kernel against the FP32 non-tensor peak."""
    sys.path.insert(0, ROOT)
    import bench
    cells, ms = 168.6e9, 7.62
    r = bench.score_roofline(cells, ms, True, 1965.0)
    peak = 148 * 128 * 1965e6 / 1e12
    assert r["bound"] == "issue_nontensor" and abs(r["peak"] - peak) < 1e-9
    assert abs(r["achieved"] - cells / (ms * 1e-3) / 1e12) < 1e-9 and abs(r["frac"] - r["achieved"] / peak) < 1e-12
    assert abs(r["alu_only_formulation"]["frac"] - 0.75 * r["achieved"] / (peak / 2)) < 1e-12
    assert abs(r["fp32_nontensor_equiv"]["frac"] - 2 * r["frac"]) < 1e-12
    assert r["measured_peak"]["peak"] > 25 and 0 < r["measured_peak"]["frac"] < 1
    f = bench.score_roofline(cells, ms, False, 1965.0)
    assert f["bound"] == "fp32_nontensor" and abs(f["frac"] - 2 * cells / (ms * 1e-3) / 1e12 / peak) < 1e-12


def test_secondary_blocks_are_guarded_on_one_gpu_only(capsys):
    """bench.guarded: on one GPU a failing secondary block becomes {"error": ...} (the headline line is still
    printed); with several ranks the failure stays fatal (a rank going on alone would leave the others in a
    collective)."""
    sys.path.insert(0, ROOT)
    import bench
    import pytest

    def boom(x, y=0):
        raise RuntimeError(f"no {x}{y}")

    assert bench.guarded(1, lambda a, b=1: a + b, 2, b=3) == 5
    block = bench.guarded(1, boom, "gpu", y=7)
    assert block == {"error": "RuntimeError: no gpu7"} and "RuntimeError" in capsys.readouterr().err
    with pytest.raises(RuntimeError):
        bench.guarded(2, boom, "gpu")
