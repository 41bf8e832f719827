"""bench.py's Pacer (spacing of the starts of concurrent mem_process_seqs calls): it must never hold a call back for long, never
wait when the calls are out of phase already, and pull calls that run in phase apart."""
import importlib.util
import os
import threading
import time

from conftest import ROOT

spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def run(pacer, n_steps, inflight, step):
    todo, lock, starts = list(range(n_steps)), threading.Lock(), []
    pacer.new_run()

    def work(w):
        while True:
            with lock:
                if not todo:
                    return
                todo.pop()
            pacer.before_start()
            t = time.perf_counter()
            with lock:
                starts.append(t)
            step(w)
            pacer.after_end(time.perf_counter() - t)
    ths = [threading.Thread(target=work, args=(w,)) for w in range(inflight)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, sorted(starts)


def test_disabled_or_single_call_never_waits():
    for p in (bench.Pacer(3, enabled=False), bench.Pacer(1)):
        dt, starts = run(p, 6, p.inflight, lambda w: time.sleep(0.01))
        assert p.waited == 0.0 and len(starts) == 6


def test_calls_in_phase_are_pulled_apart_and_throughput_is_not_limited():
    """Three identical 60-ms calls that would start together wave after wave: after the first wave their starts are spaced, every
    step still runs exactly once, and the pacing costs less than one call in total."""
    p = bench.Pacer(3)
    dt, starts = run(p, 12, 3, lambda w: time.sleep(0.06))
    assert len(starts) == 12
    assert dt < 12 * 0.06 / 3 + 0.06 + 0.06          # 4 waves + at most one call of pacing and thread noise
    gaps = [b - a for a, b in zip(starts[3:], starts[4:])]   # after the (unpaced) first wave
    assert min(gaps) > 0.004, gaps                    # no two calls start together any more
    assert 0.0 <= p.gap() <= 0.2


def test_gap_stays_below_the_completion_interval():
    p = bench.Pacer(3)
    p.dmin = 0.3
    assert 0.0 < p.gap() <= 0.1
    p.done = [0.0, 0.1, 0.2, 0.3, 0.4]
    assert abs(p.gap() - 0.07) < 1e-9                 # 0.7 x the mean interval of the last completions
