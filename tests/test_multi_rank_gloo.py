"""CPU tests of the N>1 path (`-m "not gpu"`): world_size-2 (and 3) `gloo` process groups exercise the host-side
sort-first plumbing bench.py uses on GPUs -- stripe ownership, chunk packing, in-place all-gather, unpacking --
with the REFERENCE's frame standing in for what each rank's device would have rendered into its own stripes.
The CUDA side of the same layout is checked on a GPU by
tests/test_gpu_parity.py::test_sort_first_partition_composes_to_single_gpu_image.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, have_ref
from malevich_b200 import partition


def test_partition_covers_every_tile_row_exactly_once():
    for h, n, s in ((720, 2, 1), (1080, 8, 1), (2160, 8, 4), (200, 3, 2), (2160, 5, 7)):
        rows = sorted(r for k in range(n) for r in partition.owned_tile_rows(h, n, k, s))
        assert rows == list(range(h // 8))
        img = np.arange(h * 16, dtype=np.uint32).reshape(h, 16)
        chunks = np.stack([partition.pack(img, n, k, s) for k in range(n)])
        assert chunks.shape[1] == partition.chunk_rows(h, n, s)
        assert np.array_equal(partition.unpack(chunks, h, n, s), img)


def _worker(rank, world, port, stripe_h, frame_path, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = np.load(frame_path)
        h, w = full.shape
        # this rank "renders" only the stripes it owns; everything else stays at a poison value
        mine = np.full_like(full, 0xDEADBEEF)
        for ty in partition.owned_tile_rows(h, world, rank, stripe_h):
            mine[ty * 8:(ty + 1) * 8] = full[ty * 8:(ty + 1) * 8]
        rows = partition.chunk_rows(h, world, stripe_h)
        gather = torch.zeros((world, rows, w), dtype=torch.int64)
        gather[rank] = torch.from_numpy(partition.pack(mine, world, rank, stripe_h).astype(np.int64))
        dist.all_gather_into_tensor(gather.view(-1), gather[rank].reshape(-1).clone())  # same call shape as bench.py's NCCL path
        image = partition.unpack(gather.numpy().astype(np.uint32), h, world, stripe_h)
        # Stats: bin/pair counters are per rank and must be summed; assembled triangles are replicated
        t = torch.tensor([len(partition.owned_tile_rows(h, world, rank, stripe_h))], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ok = np.array_equal(image, full) and int(t.item()) == h // 8
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write("OK" if ok else "MISMATCH")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,stripe_h", [(2, 1), (2, 4), (3, 1)])
def test_gloo_composite_reproduces_the_full_frame(world, stripe_h, tmp_path):
    if have_ref(320, 200):
        from malevich_b200 import scenes
        from oracle.ref_oracle import RefOracle
        sc = scenes.toon(320, 200)
        orc = RefOracle(320, 200, threads=1)
        orc.render(sc)
        frame = orc.colors()
    else:
        frame = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))["toon_320x200/colors"]
    frame_path = str(tmp_path / "frame.npy")
    np.save(frame_path, frame)
    port = 29600 + world * 10 + stripe_h
    mp.spawn(_worker, args=(world, port, stripe_h, frame_path, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "OK"


def _shard_worker(rank, world, port, sizes, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for n in sizes:
            host = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8)  # the replicated host buffer (same on every rank)
            sb = partition.shard_bytes(n, world)
            device = torch.full((sb * world,), 0xEE, dtype=torch.uint8)        # stands for the padded device buffer
            lo, count = partition.shard_range(n, world, rank)
            device[lo:lo + count] = torch.from_numpy(host[lo:lo + count])       # mlv_update_buffer_range: this rank's shard only
            dist.all_gather_into_tensor(device, device[rank * sb:(rank + 1) * sb].clone())  # NVLink replication (NCCL in bench.py)
            ok = ok and np.array_equal(device.numpy()[:n], host)
        with open(os.path.join(out_dir, f"shard{rank}.txt"), "w") as f:
            f.write("OK" if ok else "MISMATCH")
    finally:
        dist.destroy_process_group()


def test_shard_ranges_tile_the_buffer():
    for n in (1, 15, 16, 17, 4096, 20056032, 15000000):
        for world in (1, 2, 3, 4, 8):
            sb = partition.shard_bytes(n, world)
            assert sb % 16 == 0 and sb * world >= n
            pos = 0
            for r in range(world):
                lo, count = partition.shard_range(n, world, r)
                assert lo == min(pos, n) and lo % 16 == 0 or count == 0
                assert lo >= r * sb or count == 0
                pos = lo + count
            assert pos == n


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_upload_replicates_the_buffer(world, tmp_path):
    """The e2e upload path of bench.py at N > 1 with gloo standing in for NCCL: every rank contributes 1/N of each buffer."""
    mp.spawn(_shard_worker, args=(world, 29700 + world, (1, 17, 4096, 100003, 1200000), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"shard{r}.txt").read() == "OK"


def _host_composite_worker(rank, world, port, stripe_h, frame_path, name, out_dir):
    """What bench.py's end-to-end loop does at N > 1 with the frame composed in HOST memory: no collective on the data path.
    Every rank packs the stripes it owns (k_composite_pack's layout = partition.pack) and copies them to their rows of ONE
    shared frame (mlv_present_owned_rows_async's row arithmetic, restated here), publishes the frame, and rank 0 takes it
    once every rank has -- several frames in a row through the slots of the shared mapping."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from malevich_b200.hostframe import SharedHostFrames
        full = np.load(frame_path)
        h, w = full.shape
        ht = h // 8
        shared = SharedHostFrames(name, h, w, world, rank, slots=3, create=(rank == 0)) if rank == 0 else None
        dist.barrier()  # the mapping exists
        if rank != 0:
            shared = SharedHostFrames(name, h, w, world, rank, slots=3)
        shared.reset()
        dist.barrier()
        ok, frames = True, 5
        for f in range(frames + 1):
            if f > 0:
                shared.publish(f)  # (after mlv_present_wait: this rank's rows of frame f-1 are in host memory)
                if rank == 0:
                    got = shared.take(f - 1)
                    ok = ok and np.array_equal(got, full + np.uint32(f - 1))
            if f == frames:
                break
            rendered = np.full_like(full, 0xDEADBEEF)  # this rank's device renders only the stripes it owns
            for ty in partition.owned_tile_rows(h, world, rank, stripe_h):
                rendered[ty * 8:(ty + 1) * 8] = full[ty * 8:(ty + 1) * 8] + np.uint32(f)
            chunk = partition.pack(rendered, world, rank, stripe_h)
            dst = shared.slot_for_next()
            local = 0
            stripe = rank
            while stripe * stripe_h < ht:  # (ty / stripe_h) % world == rank, one copy per stripe
                ty0, ty1 = stripe * stripe_h, min(stripe * stripe_h + stripe_h, ht)
                dst[ty0 * 8:ty1 * 8] = chunk[local * stripe_h * 8:local * stripe_h * 8 + (ty1 - ty0) * 8]
                stripe += world
                local += 1
        dist.barrier()
        shared.close(unlink=(rank == 0))
        with open(os.path.join(out_dir, f"hostcomp{rank}.txt"), "w") as fo:
            fo.write("OK" if ok else "MISMATCH")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,stripe_h", [(2, 13), (2, 1), (3, 2)])
def test_gloo_host_memory_composite_reproduces_the_full_frame(world, stripe_h, tmp_path):
    frame = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))["toon_320x200/colors"]
    frame_path = str(tmp_path / "frame.npy")
    np.save(frame_path, frame)
    port = 29700 + world * 20 + stripe_h
    name = f"mlv_test_hostcomp_{os.getpid()}_{world}_{stripe_h}"
    mp.spawn(_host_composite_worker, args=(world, port, stripe_h, frame_path, name, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"hostcomp{r}.txt").read() == "OK"
