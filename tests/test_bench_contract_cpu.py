"""The JSON line bench.py prints is a contract with the driver: the committed lines of the last GPU run
(profiles/) must carry every key it names, and bench.py itself must parse its arguments without a GPU."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_our_arm_line_has_every_contract_key():
    files = [f for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_v*.json"))) if "reference" not in f and "gpu" not in f]
    d = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"] and d["unit"] == "audio-s/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and "workload" in d["config"] and "l2" in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"]  # DRAM bytes cannot undercut the algorithmic ones
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference") and c["gpu_output_bit_exact_on_sample"] is True
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the link ceiling the end-to-end number is read against
    assert e["link_ceiling_gbs"]["duplex"] > 0 and 0 < e["frac_of_link_ceiling"] <= 1.05
    assert e["host_output_identical_to_device_run"] is True
    # the other BASELINE configs ride in the same line, each with value, e2e and a parity flag
    for name in ("cfg1", "cfg3", "cfg4"):
        c = d["configs"][name]
        assert c["value"] > 0 and c["e2e"]["value"] > 0 and c["bit_exact_vs_oracle"] is True, name
    assert "whole stream: 620158 sound units" in d["configs"]["cfg3"]["parity_span"]
    assert d["configs"]["cfg3"]["short_block_frames"]["any_band"] > 0.01
    assert set(d["configs"]["cfg4"]["frames_per_call"]) == {"1", "8", "64"}
    for name in ("cfg1", "cfg3"):
        nt = d["configs"][name]["near_threshold"]
        assert nt["decisions"] > 0 and nt["within_1e-12"] <= nt["within_1e-9"] <= nt["decisions"]


def test_multi_gpu_lines_run_the_sharded_partition():
    """N > 1: the plan cuts frame ranges inside streams, the gathered shards equal the unsharded bytes."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*gpu_v*.json")))
    assert files
    for f in files:
        d = json.loads([ln for ln in open(f).read().splitlines() if ln.startswith("{")][-1])
        assert d["n_gpus"] > 1 and d["scaling"] == "weak"
        assert d["sharded_output_identical"] is True and d["config"]["frame_range_cuts"] >= 1
        assert max(d["config"]["shards_per_rank"]) >= 2
        det = d["sharded_check_detail"]
        assert det["gathered_shards_equal_unsharded"] and det["host_api_equals_device_on_every_rank"] and not det["mismatches"]
        assert abs(d["config"]["audio_seconds_per_gpu"] - 3600.0) < 1.0


def test_committed_ncu_capture_describes_this_build():
    """roofline.traffic and the FP64 instruction count come from a committed ncu launch list; that list must have been
    taken from the kernel sources as they are now (tools/gpu_check.sh + tools/traffic_from_launches.py refresh it)."""
    sys.path.insert(0, ROOT)
    import bench

    doc, path, fresh = bench.ncu_capture()
    assert doc is not None, "no profiles/r*_dram_traffic_*.json"
    assert fresh, "%s was captured from other kernel sources (source_sha %s, build %s): refresh it" % (
        path, doc.get("source_sha"), bench.kernel_source_sha())
    assert doc["sound_units"] == 620158
    names = {bench_name for bench_name in doc["kernels"]}
    assert {"alloc", "quant_pack", "unpack_dequant"} <= names
    per_unit, src = bench.fp64_per_unit(doc, fresh)
    assert "smsp__inst_executed_pipe_fp64" in src and 800 < per_unit < 2000


def test_reference_arm_line():
    d = latest("r*_bench_v*_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_bench_parses_without_a_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--impl" in out.stdout and "--gpus" in out.stdout
    import torch

    if torch.cuda.is_available():
        return
    # no CUDA device here: our arm must refuse loudly instead of falling back to the CPU
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_synthetic_streams_are_pure_functions_of_the_sample_index():
    """A rank that holds only a frame range of a stream has to generate exactly the samples the unsharded run sees."""
    import torch

    sys.path.insert(0, ROOT)
    import bench

    whole = bench.synth_cfg2_span(torch, 0xCA27A2 + 1, 37.5, 0, 60000, torch.device("cpu"))
    for a, b in ((0, 60000), (1, 513), (12345, 54321), (59999, 60000)):
        part = bench.synth_cfg2_span(torch, 0xCA27A2 + 1, 37.5, a, b, torch.device("cpu"))
        assert torch.equal(part, whole[:, a:b]), (a, b)
    other = bench.synth_cfg2_span(torch, 0xCA27A2 + 2, 37.5, 0, 4096, torch.device("cpu"))
    assert not torch.equal(other, whole[:, :4096])          # another stream, another signal
    z = bench.hashed_normal(torch, torch.arange(0, 1 << 18, dtype=torch.int64), 7)
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1.0) < 0.01


def test_stream_list_and_shard_staging_of_the_multi_gpu_run():
    sys.path.insert(0, ROOT)
    import bench
    from carta1_b200 import sharding

    for world in (1, 2, 3, 4, 8):
        lens = bench.stream_seconds(world, 3600.0)
        assert abs(sum(lens) - 3600.0 * world) < 1e-6 and len(lens) == world
        frames = [(int(round(s * 44100)) + 511) // 512 for s in lens]
        plan = sharding.plan(frames, world)
        sharding.check_plan(plan, frames)
        if world > 1:
            assert sum(1 for p in plan for sh in p if sh.begin > 0) >= world // 2   # frame-range cuts inside streams
            per_rank = [sum(sh.frames for sh in p) for p in plan]
            assert max(per_rank) - min(per_rank) <= 4                                   # balanced: 1 h each
    for begin in range(0, 12):
        if begin == 1:
            continue  # the plan never cuts at frame 1
        halo = 2 if begin else 0
        s0, e0, off, dec_halo = bench.shard_staging(begin, halo)
        assert 0 <= s0 <= e0 <= begin and (s0 == 0 or s0 == begin - 3)
        assert off == begin - halo - s0 and off >= 0
        assert dec_halo == begin - e0 and (dec_halo >= 1 if begin else dec_halo == 0)
        assert e0 - s0 in (0, 2)                                                        # the set-up encode's own halo
