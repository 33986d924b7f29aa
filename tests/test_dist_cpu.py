"""The multi-GPU path's host logic on CPU: world_size-2 `gloo` processes shard a clip's segments and gather the decoded
frames on rank 0 (the NCCL/NVLink version of the same calls runs in bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nseg, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from latentsync_b200.pipeline import LipsyncPipeline, shard_segments

    mine = shard_segments(nseg, rank, world)
    counts = [len(shard_segments(nseg, r, world)) for r in range(world)]
    f = 4  # frames per segment in this test
    # stand-in for run_segments(): frame value encodes (segment, frame) so the gathered order can be checked
    frames = [torch.full((f, 3, 8, 8), float(s)) + torch.arange(f).view(f, 1, 1, 1) / 10 for s in mine]
    local = torch.cat(frames) if frames else torch.empty(0, 3, 8, 8)
    clip = LipsyncPipeline.gather_frames(local, counts, dst=0)
    if rank == 0:
        want = torch.cat([torch.full((f, 3, 8, 8), float(s)) + torch.arange(f).view(f, 1, 1, 1) / 10
                          for s in range(nseg)])
        out.put(bool(clip.shape == want.shape and torch.equal(clip, want)))
    else:
        assert clip is None
    dist.barrier()
    dist.destroy_process_group()


def _clip_worker(rank, world, port, nseg, out):
    """run_clip (the sharded clip entry, lipsync_pipeline.py:500-575 over ranks) with a CPU stand-in for run_segments:
    a callable segment source (only the rank's own shard is materialised), a shorter last segment, fp16 payload"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from latentsync_b200.pipeline import LipsyncPipeline, shard_segments

    pipe = LipsyncPipeline.__new__(LipsyncPipeline)
    pipe.device = torch.device("cpu")
    built = []

    def make(i):
        built.append(i)
        return {"i": i, "f": 4 if i < nseg - 1 else 2}

    def fake_run_segments(segs, steps, guidance, segments_per_batch=1):
        return [torch.full((s["f"], 3, 8, 8), float(s["i"])) + torch.arange(s["f"]).view(-1, 1, 1, 1) / 16
                for s in segs]

    pipe.run_segments = fake_run_segments
    clip = pipe.run_clip(make, num_segments=nseg, num_inference_steps=2, guidance_scale=1.5, out_dtype=torch.float16)
    assert built == list(shard_segments(nseg, rank, world))
    if rank == 0:
        want = torch.cat(fake_run_segments([{"i": i, "f": 4 if i < nseg - 1 else 2} for i in range(nseg)], 2, 1.5))
        out.put(bool(clip.dtype == torch.float16 and clip.shape == want.shape and torch.equal(clip, want.half())))
    else:
        assert clip is None
    dist.barrier()
    dist.destroy_process_group()


def _run(nseg, world=2, target=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target or _worker, args=(r, world, port, nseg, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_gloo_shard_and_gather_even():
    _run(8)


def test_gloo_shard_and_gather_ragged():
    _run(5)  # ranks hold 3 and 2 segments


def test_gloo_shard_and_gather_fewer_segments_than_ranks():
    _run(1)  # rank 1 holds nothing


def test_gloo_run_clip_sharded_entry():
    _run(5, target=_clip_worker)


def test_gloo_run_clip_more_ranks_than_segments():
    _run(1, target=_clip_worker)
