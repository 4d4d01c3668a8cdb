"""The driver's three GPU entry points, dry-run on a machine without a GPU over the emulated C ABI (tests/_fake_cuda.py): the `-m gpu`
test files that had not met hardware when they were committed, ``__graft_entry__.smoke()``, and ``bench.py``'s own arm at a toy size.

None of the numbers mean anything (the "device" is a fiber emulation timed with the host clock); what is checked is that the scripts
run to the end, through the unmodified host package and the library's own C wrappers and kernels, and that the bench line carries
every key of the contract -- so a late edit to the measurement code cannot first fail on the round-end box."""
import json
import os
import subprocess
import sys

import pytest

import _kernel_emu as KE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not KE.available(), reason="needs g++ and the CUDA headers")

_PRELUDE = "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import _fake_cuda; _fake_cuda.enable(); " % (ROOT, os.path.join(ROOT, "tests"))


def _run(code, timeout=600):
    r = subprocess.run([sys.executable, "-c", _PRELUDE + code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return r.stdout


def test_gpu_suite_runs_on_the_emulated_gpu():
    """`pytest -m gpu --emulate-gpu`: every GPU test except the ones sized for the real machine (tests/_fake_cuda.py TOO_LARGE) --
    among them the files written after round 1's last GPU session (tests/test_gpu_zz_*.py)"""
    r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "--emulate-gpu", "-q", "-p", "no:cacheprovider"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail and "error" not in tail, r.stdout[-3000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 75, tail


def test_smoke_on_the_emulated_gpu():
    out = _run("import __graft_entry__ as g; g.smoke()")
    assert "smoke ok" in out and "transcripts bit-exact" in out


def test_bench_line_contract_on_the_emulated_gpu():
    code = ("import bench; bench.BATCH = 2; bench.UTT_SEC = 1; bench.LEGS = [('ragged', 'wav2letter', 2, True)]; "
            "bench.CONFIG5_POINTS = ((1, 20, 5), (3, 40, 8)); "
            "sys.argv = ['bench.py', '--steps', '2', '--warmup', '3', '--mid-layers', '2', '--skip-cpu']; bench.main()")
    line = json.loads(_run(code).strip().splitlines()[-1])
    assert line["metric"] == "audio-sec/sec per train step" and line["unit"] == "audio-s/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["dtype"] == "bf16" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["unit"] == "audio-s/s"
    assert line["gpu_launches"] > 0 and line["gpu_launches_per_step"] * 2 <= line["gpu_launches"] + 1
    roof = line["roofline"]
    assert set(roof) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s"
    assert roof["achieved"] > 0 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12 and "serialized" in roof
    layers = roof["serialized"]["by_layer"]                                           # per (pass, layer shape): L0 and the head stand apart
    assert isinstance(layers, list) and {r["pass"] for r in layers} >= {"conv1d_fwd", "conv1d_dgrad_wt", "conv1d_wgrad"}
    assert all(r["tflops"] > 0 and r["ms_per_step"] > 0 and r["calls_per_step"] >= 1 for r in layers)
    assert abs(sum(r["ms_per_step"] for r in layers) - roof["serialized"]["kernel_ms_per_step"]) < 1e-6 * roof["serialized"]["kernel_ms_per_step"] + 1e-9
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert line["config"]["ctc_schedule"].startswith("log-space")
    # (--skip-cpu: the cpu_baseline / config1 legs run BASELINE-sized CPU work; tests/test_host_cpu.py covers that arm's contract)
    assert set(line["hbm_kernels"]) == {"ctc_loss_raw", "greedy_decode", "bn_act_pad", "bn_act_bwd"}
    assert all(v["achieved"] > 0 and v["algorithmic_bytes_per_step"] > 0 for v in line["hbm_kernels"].values())
    assert "default_config" in line and line["default_config"]["value"] > 0          # the literal mid_layers=1 config beside the stack
    # secondary legs of the same line (the driver only runs `bench.py --gpus N`): ragged run, config-5 corners, reducer, loss trace
    rg = line["ragged"]
    assert rg["value"] > 0 and rg["conv"]["frac_sustained"] > 0 and rg["bn_act_bwd"]["GBps"] > 0 and rg["ctc_loss_raw"]["GBps"] > 0, rg
    assert len(line["config5"]["rows"]) == 2 and all(r["ctc_GBps"] > 0 and r["decode_GBps"] > 0 for r in line["config5"]["rows"])
    assert line["reducer"].startswith("none") and all(v == v for v in line["loss_last_step"]["all_timed_regions"])
    assert roof["frac_vs_burst"] > 0 and len(line["worst_layers"]) >= 1


def test_kernels_do_not_depend_on_the_thread_schedule():
    """racecheck on the host: the CTC wavefront (tagged shared-memory slots polled across warps), the GEMM's mbarrier pipeline and the
    block reductions must give the same answers when the fibers of a block are resumed in a random order every round
    (W2L_EMU_SCHEDULE, tests/kernel_emu_runtime.h) -- a missing barrier, or code counting on a warp running in lockstep, would not"""
    env = dict(os.environ, W2L_EMU_SCHEDULE="random:7")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_kernel_emu_ctc_decode.py",
                        "tests/test_kernel_emu_gemm.py", "tests/test_kernel_emu_elementwise.py", "-q", "-p", "no:cacheprovider", "-k", "not slab_mode"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT, env=env)
    tail = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:]
