"""Happens-before race detection on the engine's static schedule (oracle/plan_race_check.cpp).  The numeric phase replays
the launch list on three streams (main, side = REST part of the look-ahead split, diag = diagonal blocks factorized ahead)
and every launch is a batch of concurrently running ops; the interpreter tests execute one or two legal orders, this pass
covers all of them: no two accesses to one element of the factor, one of them a write, may be unordered."""
import ctypes as C

import pytest

import sforacle
import sparseframe_b200 as sf
from conftest import build_matrix

LF_SIDE, LF_WAIT_SIDE, LF_DIAG_AHEAD, LF_WAIT_DIAG = 1, 2, 4, 8


def race_check(backend, by_segment=0, drop=0):
    fn = backend.L.sfcheck_race_check
    fn.restype = C.c_long
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return fn(backend.h, by_segment, drop)


# blockings: default (64 / 256 / 1024) on fronts wide enough for both look-aheads, and narrow ones that split small fronts
CASES = [("lap3d", 40, "cholesky", dict()), ("cd3d", 34, "lu", dict()), ("lap3d", 30, "lu", dict()),
         ("lap3d", 16, "cholesky", dict(nb=8, outer=32 * 100000 + 16, small_a=0, small_b=0)),
         ("cd3d", 14, "lu", dict(nb=8, outer=32 * 100000 + 16, small_a=0, small_b=0)),
         ("lap2d", 120, "cholesky", dict(nb=16, outer=64 * 100000 + 32, small_b=4096)),
         ("el3d", 7, "cholesky", dict(nb=16, outer=128 * 100000 + 32))]


@pytest.mark.parametrize("gen,g,kind,prm", CASES, ids=["%s%d-%s-%s" % (c[0], c[1], c[2], "narrow" if c[3] else "default") for c in CASES])
def test_schedule_is_race_free(gen, g, kind, prm):
    A, _ = build_matrix(gen, g)
    b = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, 0, 1, **prm)
    assert race_check(b) == 0
    b.close()


@pytest.mark.parametrize("gen,g,kind", [("lap3d", 20, "cholesky"), ("cd3d", 18, "lu")])
def test_detector_sees_the_races_the_wait_flags_prevent(gen, g, kind, capfd):
    """Without the joins the look-ahead split and the diagonal blocks factorized ahead ARE races: the detector must say so
    (this is what keeps the zero above meaningful)."""
    A, _ = build_matrix(gen, g)
    b = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, 0, 1, nb=8, outer=32 * 100000 + 16, small_a=0, small_b=0)
    assert race_check(b) == 0
    assert race_check(b, drop=LF_WAIT_SIDE) > 0
    assert race_check(b, drop=LF_WAIT_DIAG) > 0
    assert race_check(b, drop=LF_SIDE | LF_WAIT_SIDE | LF_DIAG_AHEAD | LF_WAIT_DIAG) == 0      # everything on one stream
    capfd.readouterr()
    b.close()


def test_detector_covers_the_streamed_copy_out(monkeypatch, capfd):
    """The host-buffer call sends finished block columns to the host while the level is still being factorized; the
    detector models that copy stream (reads) too.  Sending one block column too many must be reported."""
    A, _ = build_matrix("lap3d", 20)
    b = sforacle.CheckBackend("cholesky", A, sf.DEFAULT_DEVSLOT, 0, 1, nb=8, outer=32 * 100000 + 16, small_a=0, small_b=0)
    assert race_check(b) == 0
    monkeypatch.setenv("SFCHECK_RACE_COPY_AHEAD", "1")
    assert race_check(b) > 0
    monkeypatch.delenv("SFCHECK_RACE_COPY_AHEAD")
    monkeypatch.setenv("SFCHECK_RACE_NO_COPY", "1")
    assert race_check(b) == 0
    capfd.readouterr()
    b.close()


@pytest.mark.parametrize("gen,g,kind,world", [("lap3d", 24, "cholesky", 4), ("cd3d", 16, "lu", 3), ("lap3d", 20, "cholesky", 8)])
def test_rank_schedules_of_a_partition_are_race_free(gen, g, kind, world, monkeypatch):
    """Partitioned plans replay segment by segment (each one ends with the streams joined); group-factorized panels with a
    small threshold so that they occur at this size."""
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "64")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "32")
    A, _ = build_matrix(gen, g)
    for rank in range(world):
        b = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, rank, world, nb=8, outer=16, small_a=0, small_b=0)
        assert race_check(b, by_segment=1) == 0
        b.close()


@pytest.mark.parametrize("gen,g,kind", [("lap3d", 24, "cholesky"), ("lap2d", 150, "cholesky"), ("cd3d", 20, "lu"), ("lap3d", 14, "lu"),
                                        ("el3d", 7, "cholesky")])
def test_solve_schedule_has_no_conflicting_tasks_in_one_launch(gen, g, kind):
    """Device solve (row f1): the sweeps are one stream; inside a launch the tasks run concurrently and touch the
    right-hand side with plain read-modify-writes (diagonal blocks) and atomic adds (updates)."""
    A, _ = build_matrix(gen, g)
    b = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, 0, 1)
    fn = b.L.sfcheck_solve_race_check
    fn.restype = C.c_long
    fn.argtypes = [C.c_void_p]
    assert fn(b.h) == 0
    b.close()
