"""The sleeping wait of ocg_dec_wait (ocg_set_blocking_sync(2)): waiters sleep on a semaphore, one poller thread
watches their completion flags.  Exercised here without a device: many threads wait on flag words that another
thread advances; nobody may miss a wake-up, wake early, or hang."""
import ctypes as C
import random
import threading
import time

from theora_b200 import abi


def test_sleeping_waiters_are_all_woken():
    L = abi.lib()
    n, rounds = 48, 60
    flags = (C.c_uint32 * (16 * n))()  # one cache line per waiter
    addr = C.addressof(flags)
    errors, done = [], [0] * n
    go = threading.Barrier(n + 1)

    def waiter(i):
        go.wait()
        for r in range(1, rounds + 1):
            ok = L.ocg_test_sleep_until(addr + 64 * i, r, 5)
            if not ok or flags[16 * i] < r:
                errors.append((i, r, ok, flags[16 * i]))
                return
            done[i] = r

    th = [threading.Thread(target=waiter, args=(i,)) for i in range(n)]
    for t in th:
        t.start()
    go.wait()
    rng = random.Random(5)
    t0 = time.time()
    for r in range(1, rounds + 1):
        order = list(range(n))
        rng.shuffle(order)
        for k, i in enumerate(order):
            flags[16 * i] = r
            if k % 7 == 0:
                time.sleep(0.0002)
        # let the slowest waiter of this round catch up before the flags move on (a waiter waits for r exactly once)
        while min(done) < r and not errors and time.time() - t0 < 60:
            time.sleep(0.0005)
    for t in th:
        t.join(timeout=30)
    assert not errors, errors[:5]
    assert all(not t.is_alive() for t in th), "a waiter hangs"
    assert done == [rounds] * n
    assert time.time() - t0 < 30, "wake-ups are being lost (waiters only return on their time-out)"


def test_sleeping_wait_times_out():
    L = abi.lib()
    flag = (C.c_uint32 * 16)()
    t0 = time.time()
    assert L.ocg_test_sleep_until(C.addressof(flag), 1, 1) == 0
    assert 0.9 < time.time() - t0 < 3.0
    flag[0] = 1
    assert L.ocg_test_sleep_until(C.addressof(flag), 1, 1) == 1
