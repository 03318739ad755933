"""The C++ benchmark drivers keep the reference's CLI and log format.

The parser below restates what the reference's postprocess.py does with a log
(benchmark04/postprocess.py:4-21, benchmark01/postprocess.py:10-20): keep the
lines that contain both marker words, x = split()[1], y = split()[3:], exactly
11 (b04/b05) or 5 (b01-03) numeric columns, title from the line with "NQ =".
"""
import math
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(binary, args=(), **env):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = os.path.join(ROOT, binary, "build", binary)
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", ROOT, "drivers"])
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    # stderr lands in the log too, as with run.sh's `&>`
    p = subprocess.run([exe, *map(str, args)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    return p.stdout.decode().splitlines()


def postprocess_parse(lines, key, metric, ncols):
    data = [l for l in lines if key in l and metric in l]
    title = [l for l in lines if "NQ =" in l]
    xs = [float(l.split()[1]) for l in data]
    ys = [[float(v) for v in l.split()[3:]] for l in data]
    assert all(len(y) == ncols for y in ys), [len(y) for y in ys]
    return xs, ys, title


def norms(lines, key):
    return {l.split()[1]: [float(v) for v in l.split()[3:]] for l in lines if l.startswith(key) and " norm:" in l}


@pytest.mark.parametrize("plan", [1, 0])
def test_benchmark04_cli_log_format_and_golden_norms(golden, plan):
    """plan = 1 (default): every repetition through a b200fe_plan; plan = 0: the per-call entry points"""
    lines = run("benchmark04", (4, 4), B200FE_NELMT="128,4096", B200FE_REPS=3, B200FE_CPU_REPS=1, B200FE_PLAN=plan)
    assert any(l.startswith("info") and ("via b200fe_plan" if plan else "per-call entry points") in l for l in lines)
    assert lines[1].startswith("Benchmark04 : BwdTrans (2D)")
    xs, ys, title = postprocess_parse(lines, "nelmt", "DOF/s", 11)
    assert xs == [128.0, 4096.0] and len(title) == 1 and "NQ = 4, 4" in title[0]
    assert all(v > 0 and math.isfinite(v) for y in ys for v in y)
    for n, cols in norms(lines, "nelmt").items():
        assert len(cols) == 11
        for got, want in zip(cols, golden["quad"]["4"][n]):
            assert abs(got - want) / want < 6e-10, (n, cols)
    # no extra line may look like data to postprocess.py
    assert not [l for l in lines if l.startswith("info") and "nelmt" in l and "DOF/s" in l]


def test_benchmark04_default_arguments_are_8_8():
    lines = run("benchmark04", (), B200FE_NELMT="128", B200FE_REPS=2, B200FE_SKIP_CPU=1)
    assert any("NQ = 8, 8" in l for l in lines)


@pytest.mark.parametrize("plan", [1, 0])
def test_benchmark05_log_format_and_golden_norms(golden, plan):
    lines = run("benchmark05", (4, 4, 4), B200FE_NELMT="128,1024", B200FE_REPS=3, B200FE_CPU_REPS=1, B200FE_PLAN=plan)
    assert any(l.startswith("info") and ("via b200fe_plan" if plan else "per-call entry points") in l for l in lines)
    xs, ys, title = postprocess_parse(lines, "nelmt", "DOF/s", 11)
    assert xs == [128.0, 1024.0] and "NQ = 4, 4, 4" in title[0]
    for n, cols in norms(lines, "nelmt").items():
        ref = golden["hex"]["4"][n]
        for c, (got, want) in enumerate(zip(cols, ref)):
            want = ref[0] if c == 6 else want  # col 7: reference bug (benchmark05.cc:193); ours is correct
            assert abs(got - want) / want < 6e-10, (n, c, got, want)


def test_benchmark05_float_and_unequal_nq():
    lines = run("benchmark05", (3, 4, 5), B200FE_NELMT="64", B200FE_REPS=2, B200FE_DTYPE="both", B200FE_CPU_REPS=1)
    xs, ys, _ = postprocess_parse(lines, "nelmt", "DOF/s", 11)
    assert xs == [64.0, 64.0]
    n = norms(lines, "nelmt")["64"]
    assert max(n) - min(n) < 1e-4 * max(n)


@pytest.mark.parametrize("binary,key,tol", [("benchmark01", "b01", 6e-10), ("benchmark02", "b02", 6e-10),
                                             ("benchmark03", "b03", 5e-9)])
def test_benchmark01_03_log_format_and_golden_norms(golden, binary, key, tol):
    sizes = "1024,65536" if key != "b03" else "128,512"
    lines = run(binary, (), B200FE_SIZES=sizes)
    xs, ys, _ = postprocess_parse(lines, "Size", "GB/s", 5)
    assert len(xs) == 2
    for n, cols in norms(lines, "Size").items():
        assert len(cols) == 5
        for got, want in zip(cols, golden[key][n]):
            assert abs(got - want) / want < tol, (binary, n, cols)


@pytest.mark.parametrize("bench,orders,fmt", [("benchmark04", (2, 4, 6, 8, 10, 12, 14, 16, 32), "nq{0}x{0}.log"),
                                              ("benchmark05", (2, 4, 6, 8, 10), "nq{0}x{0}x{0}.log")])
def test_reference_run_sh_unmodified_drives_the_built_binaries(golden, bench, orders, fmt, tmp_path):
    """The reference's run.sh, byte for byte (tests/test_driver_logs_cpu.py pins its hash): `cd build/`, the nq loop,
    CUDA_VISIBLE_DEVICES=1, `&> ../nq*.log` (benchmark04/run.sh:3-8).  It pins device 1, so this needs a box with at
    least two GPUs (gpurun --gpus 2); every log then goes through the parser half of postprocess.py."""
    import shutil
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("run.sh pins CUDA_VISIBLE_DEVICES=1: needs >= 2 GPUs")
    exe = os.path.join(ROOT, bench, "build", bench)
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", ROOT, "drivers"])
    # a scratch copy of the benchmark directory layout run.sh expects: ./run.sh, ./build/<binary>, logs land in ./
    work = tmp_path / bench
    (work / "build").mkdir(parents=True)
    shutil.copy2(os.path.join(ROOT, bench, "run.sh"), work / "run.sh")
    os.symlink(exe, work / "build" / bench)
    e = dict(os.environ)
    e.update({"B200FE_NELMT": "128,4096", "B200FE_REPS": "3", "B200FE_CPU_REPS": "1"})  # env knobs pass through run.sh
    e.pop("CUDA_VISIBLE_DEVICES", None)
    p = subprocess.run(["bash", "run.sh"], cwd=work, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    assert p.stdout.decode().split() == [f"nq={n}" for n in orders]
    kind = "quad" if bench == "benchmark04" else "hex"
    for nq in orders:
        lines = (work / fmt.format(nq)).read_text().splitlines()
        xs, ys, title = postprocess_parse(lines, "nelmt", "DOF/s", 11)
        assert xs == [128.0, 4096.0] and len(title) == 1 and f"NQ = {nq}, {nq}" in title[0], lines[-5:]
        assert all(v > 0 and math.isfinite(v) for y in ys for v in y)
        for n, cols in norms(lines, "nelmt").items():
            ref = golden[kind][str(nq)][n]
            for c, (got, want) in enumerate(zip(cols, ref)):
                want = ref[0] if (kind == "hex" and c == 6) else want
                assert abs(got - want) / want < 6e-10, (nq, n, c, got, want)


def test_column5_without_cublas(golden):
    """B200FE_COL5=gemm: column 5 from b200fe_gemm_bwdtrans_* (same factorisation as the reference's cuBLAS calls,
    benchmark04.cc:804-820 / benchmark05.cc:1128-1153), golden norms as every other column"""
    for binary, args, kind, key in (("benchmark04", (8, 8), "quad", "8"), ("benchmark05", (6, 6, 6), "hex", "6")):
        lines = run(binary, args, B200FE_NELMT="128,2048", B200FE_REPS=3, B200FE_SKIP_CPU=1, B200FE_SKIP_CUBLAS=1,
                    B200FE_COL5="gemm")
        xs, ys, _ = postprocess_parse(lines, "nelmt", "DOF/s", 11)
        assert all(y[4] > 0 and math.isfinite(y[4]) for y in ys)
        for n, cols in norms(lines, "nelmt").items():
            want = golden[kind][key][n][0]
            assert abs(cols[4] - want) / want < 6e-10, (binary, n, cols[4], want)
