"""Known-answer tests that pin the oracle's third-party arithmetic.

The reference holds no golden vectors for this path, and cannot be compiled here;
these are the public test vectors of SplitMix64 / xoshiro256++ (rand_xoshiro's own unit vector)
plus hand-derivable properties of rand 0.8's samplers and of Rust's BinaryHeap."""
import ctypes as C
import heapq

import numpy as np
import pytest


def _u64s(n):
    return (C.c_uint64 * n)()


def test_splitmix64_public_vector(oracle):
    out = _u64s(5)
    oracle.oracle().oracle_kat_splitmix64(C.c_uint64(1234567), 5, out)
    assert list(out) == [6457827717110365317, 3203168211198807973, 9817491932198370423, 4593380528125082431,
                         16408922859458223821]


def test_xoshiro256pp_reference_vector(oracle):
    # rand_xoshiro::Xoshiro256PlusPlus unit test vector: state [1,2,3,4]
    out = _u64s(10)
    st = (C.c_uint64 * 4)(1, 2, 3, 4)
    oracle.oracle().oracle_kat_xoshiro_state(st, 10, out)
    assert list(out) == [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
                         9973669472204895162, 14011001112246962877, 12406186145184390807, 15849039046786891736,
                         10450023813501588000]


def test_seed_from_u64(oracle):
    # [DERIVED] in SURVEY Appendix A.2 from the two public algorithms above
    st, out = _u64s(4), _u64s(4)
    oracle.oracle().oracle_kat_seed_from_u64(C.c_uint64(0), st, 4, out)
    assert [hex(x) for x in st] == ["0xe220a8397b1dcdaf", "0x6e789e6aa1b965f4", "0x6c45d188009454f", "0xf88bb8a8724c81ec"]
    assert list(out) == [5987356902031041503, 7051070477665621255, 6633766593972829180, 211316841551650330]
    for seed, first in ((1, 14971601782005023387), (2333, 14741951565609963218)):
        oracle.oracle().oracle_kat_seed_from_u64(C.c_uint64(seed), st, 1, out)
        assert out[0] == first


def _draws(oracle, seed, kind, lo, hi, p, n):
    out = _u64s(n)
    used = C.c_uint64(0)
    oracle.oracle().oracle_kat_draws(C.c_uint64(seed), kind, C.c_uint64(lo), C.c_uint64(hi), C.c_double(p), n, out,
                                     C.byref(used))
    return list(out), used.value


def test_default_latency_first_sample(oracle):
    # if the FIRST draw of seed 0 / seed 1 were a default-latency sample
    # (UniformDuration Small mode over 1 ms..10 ms) it would be 3 921 177 ns / 8 304 509 ns
    assert _draws(oracle, 0, 2, 1_000_000, 10_000_000, 0, 1)[0] == [3_921_177]
    assert _draws(oracle, 1, 2, 1_000_000, 10_000_000, 0, 1)[0] == [8_304_509]


def test_gen_range_matches_python_model(oracle):
    # UniformInt::sample_single: widening multiply + conservative zone (Appendix A.3)
    def xoshiro(seed):
        M = (1 << 64) - 1

        def sm():
            nonlocal x
            x = (x + 0x9E3779B97F4A7C15) & M
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
            return z ^ (z >> 31)

        x = seed
        s = [sm() for _ in range(4)]
        rotl = lambda v, k: ((v << k) | (v >> (64 - k))) & M
        while True:
            r = (rotl((s[0] + s[3]) & M, 23) + s[0]) & M
            t = (s[1] << 17) & M
            s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)
            yield r

    for seed, lo, hi in ((0, 0, 1), (7, 50, 100), (3, 0, 31_536_000), (11, 150, 300), (5, 0, 5)):
        g = xoshiro(seed)
        want, used = [], 0
        for _ in range(200):
            rng = hi - lo
            zone = ((rng << (64 - rng.bit_length())) - 1) & ((1 << 64) - 1)
            while True:
                v = next(g); used += 1
                m = v * rng
                if (m & ((1 << 64) - 1)) <= zone:
                    want.append(lo + (m >> 64)); break
        got, got_used = _draws(oracle, seed, 0, lo, hi, 0, 200)
        assert got == want and got_used == used
    # gen_range(0..1) rejects every v >= 2^63: mean 2 draws per sample (Appendix A.3)
    _, used = _draws(oracle, 42, 0, 0, 1, 0, 4000)
    assert 1.9 < used / 4000 < 2.1


def test_gen_bool_rules(oracle):
    # Bernoulli: p == 1.0 → true WITHOUT a draw; p == 0.0 draws and is false (Appendix A.5)
    vals, used = _draws(oracle, 9, 3, 0, 0, 1.0, 50)
    assert vals == [1] * 50 and used == 0
    vals, used = _draws(oracle, 9, 3, 0, 0, 0.0, 50)
    assert vals == [0] * 50 and used == 50
    vals, used = _draws(oracle, 9, 3, 0, 0, 0.25, 4000)
    assert used == 4000 and 0.2 < sum(vals) / 4000 < 0.3


def test_uniform_duration_modes(oracle):
    # Small with carry (1.5 s .. 2.2 s → secs fixed at 1, nanos up to 1 199 999 999) and Medium
    vals, _ = _draws(oracle, 3, 2, 1_500_000_000, 2_200_000_000, 0, 2000)
    assert min(vals) >= 1_500_000_000 and max(vals) < 2_200_000_000 and max(vals) > 2_000_000_000
    vals, _ = _draws(oracle, 3, 2, 1_000_000_000, 10_000_000_000, 0, 2000)
    assert min(vals) >= 1_000_000_000 and max(vals) < 10_000_000_000


def _rust_binary_heap_order(deadlines):
    """Pure-Python model of std::collections::BinaryHeap<Reverse-by-deadline> push / pop
    (sift_up strict; pop = swap last to root, sift_down_to_bottom preferring the RIGHT child on
    ties, then sift_up) — SURVEY Appendix A.6."""
    h = []

    def le(a, b):  # Ord reversed on deadline: a <= b  ⇔  a.deadline >= b.deadline
        return a[0] >= b[0]

    def sift_up(start, pos):
        e = h[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if le(e, h[parent]):
                break
            h[pos] = h[parent]; pos = parent
        h[pos] = e

    for i, d in enumerate(deadlines):
        h.append((d, i)); sift_up(0, len(h) - 1)
    order = []
    while h:
        item = h.pop()
        if h:
            item, h[0] = h[0], item
            end, pos = len(h), 0
            e = h[0]
            child = 1
            while child <= max(end - 2, 0) and end >= 2:
                if le(h[child], h[child + 1]):
                    child += 1
                h[pos] = h[child]; pos = child; child = 2 * pos + 1
            if child == end - 1:
                h[pos] = h[child]; pos = child
            h[pos] = e
            sift_up(0, pos)
        order.append(item[1])
    return order


@pytest.mark.parametrize("seed", range(6))
def test_timer_heap_tie_order(oracle, seed):
    rs = np.random.RandomState(seed)
    n = 40
    deadlines = [int(x) for x in rs.randint(0, 6, size=n)]  # many ties
    out = (C.c_uint32 * n)()
    oracle.oracle().oracle_kat_heap_order((C.c_uint64 * n)(*deadlines), n, None, out)
    got = list(out)
    assert got == _rust_binary_heap_order(deadlines)
    assert [deadlines[i] for i in got] == sorted(deadlines)
    # and it is NOT simply insertion order among ties (that is why a warp-reduce min cannot replace the heap)
    if seed == 0:
        stable = [i for _, i in sorted((d, i) for i, d in enumerate(deadlines))]
        assert got != stable


def test_appendix_b_worked_example(oracle):
    import madsim_b200 as m

    r, tr, _, _ = oracle.run_seed(m.Config(), m.Workload.sleep(10**9), 0, trace=True)
    assert (int(r["elapsed_ns"]), int(r["draws"]), int(r["ev_pop"]), int(r["ev_fire"])) == (1_000_000_101, 6, 2, 1)
    assert [(int(t["now_ns"]), int(t["kind_id"]) >> 24, int(t["draws"])) for t in tr] == [
        (0, 1, 2), (50, 5, 4), (1_000_000_050, 1, 5)]
    r, _, _, _ = oracle.run_seed(m.Config(), m.Workload.sleep(10**9), 1)
    assert (int(r["elapsed_ns"]), int(r["draws"])) == (1_000_000_129, 6)
