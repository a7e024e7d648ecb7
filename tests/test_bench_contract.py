"""CPU: the reference arm of bench.py runs here (oracle/_ref, the reference's own code on the host cores)
and prints one JSON line with the contract's keys; the GPU arm's source carries the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "dedalus")),
                    reason="oracle/_ref not built (needs /root/reference: __graft_entry__.build())")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-n", "16", "--steps", "1",
                        "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] in (1, cb["threaded_fft_workers"]) and cb["host_cores"] >= cb["cores"]
    assert cb["single_thread_value"] > 0 and cb["threaded_fft_value"] > 0 and cb["value"] == max(cb["single_thread_value"], cb["threaded_fft_value"])
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_emits_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"clocks"', '"e2e"', '"gpu_launches"', '"roofline"',
                '"cpu_baseline"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"traffic"', '"frac"', '"peak"', '"bound"'):
        assert key in src, key
    assert "/root/reference" not in src


def test_gpu_arm_runs_against_the_package_under_host_emulation():
    """bench.py's GPU arm, line for line, in the GPU-less container (tests/bench_emul_child.py: the conftest harness plus CPU
    stand-ins for the torch.cuda calls the script makes itself).  Guards the script against drifting away from the package
    between GPU runs: same launch labels, 24 launches per RK4 step of 3-D MHD, every contract key on the line."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.pop("DEDALUS_DDL_LIB", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_emul_child.py"), "32"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["gpu_launches"] == 24 * line["steps"] and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert set(line["roofline"]["kernels"]) == {"z_inv", "y_inv", "x_fused", "y_fwd", "z_fwd", "assemble_stage"}
    assert all(k["launches"] == 8 for k in line["roofline"]["kernels"].values())          # 2 instrumented steps x 4 RHS
    # end-to-end leg: the retained box only (the short download was checked against the full one before the timed region) ...
    nbytes = 6 * line["config"]["N_k"] * 16
    assert line["e2e"]["transfer"].startswith("retained modes only")
    assert 0.2 * nbytes < line["e2e"]["h2d_bytes_per_step"] == line["e2e"]["d2h_bytes_per_step"] < 0.36 * nbytes
    # ... and the same trajectory as with the full arrays through comp['kspace']
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_emul_child.py"), "32", "--e2e-full"], stdout=subprocess.PIPE,
                        stderr=subprocess.PIPE, text=True, timeout=900, env=env, cwd=ROOT)
    assert r2.returncode == 0, r2.stderr[-3000:]
    full = json.loads(r2.stdout.strip().splitlines()[-1])
    assert full["e2e"]["transfer"].startswith("full arrays") and full["e2e"]["h2d_bytes_per_step"] == nbytes
    assert abs(full["e2e"]["ekin_after"] - line["e2e"]["ekin_after"]) < 1e-14 and 0 < line["e2e"]["ekin_after"] < line["invariants"]["ekin"]
    assert line["roofline"]["kernel"] in line["roofline"]["kernels"] and line["roofline"]["frac"] > 0
    assert 0 < line["invariants"]["ekin"] < 1 and 0 < line["invariants"]["emag"] < 1


def test_gpu_arm_under_torchrun_world_size_2_host_emulation():
    """The same under torchrun with two ranks (gloo standing in for nccl, collective exchange): rank 0 alone prints the line,
    the slab bookkeeping bench.py reads for its NVLink figures is still there, and the result is the same global field."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for k in ("DEDALUS_DDL_LIB", "DEDALUS_KY_LAYOUT", "DEDALUS_SLAB_EXCHANGE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "bench_emul_child.py"), "32"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and "cyclic" in line["config"]["parallelism"]
    assert line["roofline"]["nvlink"]["bytes_out_per_gpu_per_step"] > 0 and line["roofline"]["nvlink"]["transposes_per_step"] == 60
    assert line["e2e"]["transfer"].startswith("retained modes only")            # cyclic ky rows: still two runs of kept rows per rank
    assert 0.2 < line["e2e"]["h2d_bytes_per_step"] / (6 * line["config"]["N_k"] * 16) < 0.36
    assert abs(line["e2e"]["ekin_after"] - 0.4874384573027503) < 1e-9          # the single-rank run's value after the same steps
    # the single-rank run of the same script gives these invariants for the same synthetic field (make_state draws the global noise)
    assert abs(line["invariants"]["ekin"] - 0.4963064899312569) < 1e-9 and abs(line["invariants"]["emag"] - 0.5054009950719366) < 1e-9
    assert "cpu_baseline" not in line


def test_moved_bytes_model_matches_the_committed_ncu_traffic():
    """bench.py's roofline uses bytes computed from the plan's retained-mode counts; they must be what the hardware counters
    saw (profiles/ncu_traffic.json: dram__bytes_read + write of one launch per kernel, 512^3 MHD) to within 3 %."""
    sys.path.insert(0, ROOT)
    import bench
    traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["512"]
    model = bench.moved_bytes_per_rhs(512, (341, 341, 171))
    assert set(traffic) == {"z_inv", "y_inv", "x_fused", "y_fwd", "z_fwd", "assemble_stage"}
    for k, v in traffic.items():
        m = model["assemble_stage_middle" if k == "assemble_stage" else k]      # the capture is a middle stage of a step
        assert abs(m - v["dram_bytes"]) < 0.03 * v["dram_bytes"], (k, m, v["dram_bytes"])
        assert os.path.exists(os.path.join(ROOT, v["source"].split(" ")[0])), v["source"]
    # no kernel can move more than the SURVEY model's full-array bytes, and the step sum is the ~52 GB DESIGN.md quotes
    nk = 512 * 512 * 257
    for k in ("z_inv", "y_inv", "x_fused", "y_fwd", "z_fwd"):
        assert model[k] < bench.ALGO_BYTES_PER_MODE[k] * nk
    assert 50e9 < sum(model[k] for k in traffic) < 54e9


def test_smoke_body_runs_against_the_emulation_build():
    """__graft_entry__.smoke() checks two small RK4 runs against the oracle on cuda:0; its body (everything after the GPU asserts)
    is dry-run here so that the driver's GPU call is not spent on a Python error: 32^3 through the generic kernels, 16 x 256 x 128
    through the two-stage strided pass, the fused x pass and the fused assembly / stage kernel."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_emul_child.py"), os.path.join(ROOT, "tests", "smoke_emul_script.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("smoke:")]
    assert len(lines) == 2 and "16x256x128" in lines[1], r.stdout[-2000:]

