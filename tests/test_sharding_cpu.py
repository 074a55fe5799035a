"""Multi-rank host logic on CPU (world_size 2, gloo): frame-wise sharding without a data-path
collective, timing / frame-count combination, and the rank-0-only reference arm of bench.py."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

import smh_b200  # noqa: F401
from squad_mortar_helper_b200 import shard

ROOT = Path(__file__).resolve().parent.parent


def test_shard_ranges_partition_the_batch():
    for world in (1, 2, 4, 8):
        for total in (0, 1, 7, 1024, 4096, 4097):
            cover = []
            for r in range(world):
                a, b = shard.shard_range(r, world, total)
                assert 0 <= a <= b <= total
                cover += list(range(a, b))
            assert cover == list(range(total)), (world, total)
    assert shard.weak_shard(3, 1024) == (3072, 4096)
    with pytest.raises(ValueError):
        shard.shard_range(2, 2, 10)


def test_python_and_c_abi_split_alike():
    """smh_b200_shard_range (what smh_b200_multi_process_host_batch gives each GPU) == shard.shard_range (what a
    torchrun rank takes)."""
    import ctypes as C
    from squad_mortar_helper_b200 import capi
    L = capi.lib()
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 1000, 4096, 4097):
            for r in range(world):
                a, c = C.c_uint32(), C.c_uint32()
                L.smh_b200_shard_range(total, world, r, C.byref(a), C.byref(c))
                assert (a.value, a.value + c.value) == shard.shard_range(r, world, total)


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import smh_b200  # noqa: F401
    from helpers import oracle_run
    from squad_mortar_helper_b200 import shard as S, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 5
    a, b = S.shard_range(rank, world, total)
    # every rank rasterises and analyses only its own frames (host rasteriser + CPU checker here;
    # on the GPU box the same indices go through smh_b200_process_batch)
    p = synth.SynthParams(1024, 768, n_segments=3, n_blips=2, n_labels=1)
    frames = synth.frames_host(p, a, b - a)
    n_lines = [int(oracle_run(f, True, 0, 15)["lines"].shape[0]) for f in frames]
    dist.barrier()
    # what bench.py does with its timings: gather, then MAX over ranks / SUM of frames
    import torch
    t = torch.tensor([float(b - a), 10.0 * (rank + 1)], dtype=torch.float64)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    frames_done, elapsed = int(sum(float(x[0]) for x in parts)), max(float(x[1]) for x in parts)
    json.dump({"range": [a, b], "n_lines": n_lines, "frames": frames_done, "elapsed": elapsed},
              open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    dist.destroy_process_group()


def test_two_ranks_gloo_shard_and_combine(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(world)]
    assert [r["range"] for r in res] == [[0, 2], [2, 5]]
    for r in res:  # both ranks see the combined numbers: sum of frames, MAX of the elapsed times
        assert r["frames"] == 5 and r["elapsed"] == 20.0
    # the shards together give exactly what one process gets for the whole batch
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import oracle_run
    from squad_mortar_helper_b200 import synth
    p = synth.SynthParams(1024, 768, n_segments=3, n_blips=2, n_labels=1)
    whole = [int(oracle_run(f, True, 0, 15)["lines"].shape[0]) for f in synth.frames_host(p, 0, 5)]
    assert res[0]["n_lines"] + res[1]["n_lines"] == whole


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, WORLD_SIZE="2", RANK="1", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "", r.stderr[-400:]
    env["RANK"] = "0"; env["LOCAL_RANK"] = "0"
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-400:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
