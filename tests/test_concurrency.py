"""Concurrent callers: the handles are immutable after creation and every entry point may be called
from several host threads at once — the reference's get_cost is const and re-entrant (dpose_core.cpp:200) and the costmap
layers call into the library from their own threads.  ctypes releases the GIL around the C calls, so Python threads do run
them concurrently.  Every concurrent result must equal the serial one bit for bit."""
import threading

import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dp():
    import dpose_b200
    assert dpose_b200.device_count() > 0, "no CUDA device: GPU tests cannot run (no CPU fallback)"
    return dpose_b200


def run_threads(workers):
    errors = []

    def guard(fn):
        def inner():
            try:
                fn()
            except BaseException as e:  # noqa: BLE001 - reported below
                errors.append(repr(e))
        return inner
    threads = [threading.Thread(target=guard(w)) for w in workers]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


def test_concurrent_calls_match_serial(dp):
    m = fx.bernoulli_map(600, 500, 0.10, 51)
    cm = dp.Costmap(m)
    pg_rect = dp.pose_gradient(fx.RECT)
    fp_star, pad_star = fx.FOOTPRINTS["star40_pad2"]
    pg_star = dp.pose_gradient(fp_star, dp.pose_gradient.parameter(pad_star))
    poses_small = fx.random_poses(3000, 600, 500, 52)
    poses_big = fx.random_poses(300000, 600, 500, 53)  # more than one staging chunk
    ring = fx.rotated_ring(pg_rect.get_data().get_box(), (300.0, 250.0, 0.4))
    cells = dp.lethal_cells_within(cm, ring)
    want = {
        "small_rect": pg_rect.get_cost_batch(cm, poses_small),
        "big_rect": pg_rect.get_cost_batch(cm, poses_big),
        "small_star": pg_star.get_cost_batch(cm, poses_small),
        "single": pg_rect.get_cost([300.0, 250.0, 0.4], cells),
        "cells": cells,
        "check": dp.check_footprint_batch(cm, fx.RECT, poses_small),
    }
    rounds = 12

    def w_small_rect():
        for _ in range(rounds):
            assert np.array_equal(pg_rect.get_cost_batch(cm, poses_small), want["small_rect"])

    def w_big_rect():
        for _ in range(3):
            assert np.array_equal(pg_rect.get_cost_batch(cm, poses_big), want["big_rect"])

    def w_small_star():
        for _ in range(rounds):
            assert np.array_equal(pg_star.get_cost_batch(cm, poses_small), want["small_star"])

    def w_single():
        for _ in range(rounds * 20):
            c, J = pg_rect.get_cost([300.0, 250.0, 0.4], cells)
            assert c == want["single"][0] and np.array_equal(J, want["single"][1])

    def w_lethal():
        for _ in range(rounds * 4):
            assert np.array_equal(dp.lethal_cells_within(cm, ring), want["cells"])

    def w_check():
        for _ in range(rounds):
            assert np.array_equal(dp.check_footprint_batch(cm, fx.RECT, poses_small), want["check"])

    run_threads([w_small_rect, w_small_rect, w_big_rect, w_small_star, w_single, w_single, w_lethal, w_check, w_check])


def test_concurrent_first_use_of_a_fresh_map(dp):
    """several threads hit a map whose lethal planes have not been built yet (lazy build on first batched call)"""
    m = fx.bernoulli_map(800, 700, 0.10, 61)
    pg = dp.pose_gradient(fx.RECT)
    poses = fx.random_poses(20000, 800, 700, 62)
    want = pg.get_cost_batch(dp.Costmap(m), poses)
    for _ in range(5):
        cm = dp.Costmap(m)
        pgs = [dp.pose_gradient(fx.RECT) for _ in range(4)]  # one handle per thread: no serialisation on its staging
        outs = [None] * 4

        def worker(i):
            def run():
                outs[i] = pgs[i].get_cost_batch(cm, poses)
            return run
        run_threads([worker(i) for i in range(4)])
        for o in outs:
            assert np.array_equal(o, want)
