"""The shared host frame of a multi-process job (malevich_b200/hostframe.py) without any GPU: two producer processes write
the rows they own, the consumer takes a frame only when every rank has published it, and a slot is not overwritten before
the frame that used it has been consumed."""
import multiprocessing as mp
import os

import numpy as np

from malevich_b200.hostframe import SharedHostFrames

H, W, WORLD, FRAMES = 16, 8, 2, 7


def _rank(rank: int, name: str, out):
    sh = SharedHostFrames(name, H, W, WORLD, rank)
    rows = range(rank * H // WORLD, (rank + 1) * H // WORLD)
    seen = []
    for f in range(FRAMES):
        if f > 0:
            sh.publish(f)
            if rank == 0:
                seen.append(int(sh.take(f - 1).sum()))
        slot = sh.slot_for_next()
        for y in rows:
            slot[y, :] = 1000 * f + y
    sh.publish(FRAMES)
    if rank == 0:
        seen.append(int(sh.take(FRAMES - 1).sum()))
        out.put(seen)
    sh.close()


def test_two_processes_compose_frames_in_one_shared_mapping():
    name = f"mlv_test_frames_{os.getpid()}"
    creator = SharedHostFrames(name, H, W, WORLD, 0, create=True)
    creator.reset()
    try:
        ctx = mp.get_context("spawn")
        out = ctx.Queue()
        procs = [ctx.Process(target=_rank, args=(r, name, out)) for r in range(WORLD)]
        for p in procs:
            p.start()
        seen = out.get(timeout=60)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        want = [int(sum((1000 * f + y) * W for y in range(H))) for f in range(FRAMES)]
        assert seen == want
    finally:
        creator.close(unlink=True)


def test_slot_reuse_waits_for_the_consumer():
    name = f"mlv_test_frames_b_{os.getpid()}"
    sh = SharedHostFrames(name, H, W, 1, 0, create=True)
    sh.reset()
    try:
        sh.slot_for_next()
        sh.slot_for_next()
        try:
            sh.slot_for_next(timeout_s=0.05)  # frame 0 has not been consumed: its slot is not free
            raise AssertionError("expected a timeout")
        except TimeoutError:
            pass
        sh.delivered = 2
        sh.publish(2)
        sh.take(0)
        try:
            sh.slot_for_next(timeout_s=0.05)  # frame 0 is still in the consumer's hands
            raise AssertionError("expected a timeout")
        except TimeoutError:
            pass
        sh.delivered = 2
        sh.take(1)  # ... taking frame 1 lets go of it
        assert sh.slot_for_next(timeout_s=0.05) is sh.frames[0]
    finally:
        sh.close(unlink=True)
