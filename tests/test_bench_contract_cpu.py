"""bench.py contract (CPU side): the reference arm prints ONE JSON line with the keys the driver reads, without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mh_proposals_per_sec" and d["unit"] == "proposals/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "mh_2olx65_chains1024_per_gpu"
    cb = d["cpu_baseline"]
    # the flow half runs the byte-compiled unmodified reference when oracle/_ref was built (python -m oracle.build_ref), else the port
    from oracle import ref_flow

    assert cb["kind"] == ("reference(flow)+port(energy)" if ref_flow.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert set(d["config"]) >= {"workload", "atoms", "chains_per_gpu", "model", "energy", "l2"}
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU reference; the other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_clock_sampler_rows_and_sources():
    """The clocks line of the bench contract: NVML in-process when a driver is present, nvidia-smi otherwise; both feed the same
    summary (median SM clock under load, max clock, active slowdown reasons)."""
    import bench

    smp = bench.ClockSampler(0)
    assert smp.source in ("nvml", "nvidia-smi")

    class FakeNvml:
        NVML_CLOCK_SM = 1

        @staticmethod
        def nvmlDeviceGetClockInfo(h, kind):
            return 1650

        @staticmethod
        def nvmlDeviceGetMaxClockInfo(h, kind):
            return 1965

        @staticmethod
        def nvmlDeviceGetCurrentClocksEventReasons(h):
            return 0x4 | 0x40  # sw_power_cap + hw_thermal_slowdown

        @staticmethod
        def nvmlDeviceGetPowerUsage(h):
            return 950_000

    smp._nvml, smp._h, smp.source = FakeNvml, object(), "nvml"
    smp.rows = [smp._sample_nvml(), smp._sample_nvml()]
    out = smp.summary()
    assert out["sm_mhz"] == 1650.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 2 and out["source"] == "nvml"
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    smp.rows = [["0", "1700", "1965", "900.0", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"]]  # an nvidia-smi row
    assert smp.summary()["reasons"] == []
