"""Bank-conflict model of the PADDED shared-memory layout of the shear FFTs (csrc/derotate.cu, ShearFftP<N, true>).

64-bit shared-memory accesses are served per half-warp: the 16 lanes must hit 16 distinct 8-byte banks (of 16).
For N = 512 ... 4096 (R3 = N/256 = 2 ... 16 threads per length-L1 block) this script enumerates every access pattern
of the forward / inverse transforms in the padded layout

    exchange A (stage 1 <-> 2):  row k1, column x                 ->  k1 * PITCH + x
    exchange B (stage 2 <-> 3):  block k1p, element q in [0, L1)  ->  k1p * PITCH + q + (q >> 4)        PITCH = 17 * R3

and asserts that no half-warp has two lanes on the same bank, and that distinct elements never alias.
    python tools/fft_pad_model.py
"""
import itertools


def check(N):
    T, R3 = N // 16, N // 256
    L1, L2 = N // 16, R3
    log_r3 = R3.bit_length() - 1
    pitch = 17 * R3

    def iA(k1, x):
        return k1 * pitch + x

    def iB2(k1p, k2, npp):
        return k1p * pitch + npp + k2 * L2 + ((k2 * L2) >> 4)

    def iB3(t, e):
        return (t >> log_r3) * pitch + 17 * (t & (R3 - 1)) + e

    def banks_ok(addr_of_lane, what):
        for hw in range(0, T, 16):
            banks = [addr_of_lane(t) % 16 for t in range(hw, min(T, hw + 16))]
            assert len(set(banks)) == len(banks), (N, what, hw, banks)

    for k1 in range(16):                                   # stage-1 write / inverse stage-1 read
        banks_ok(lambda t: iA(k1, t), f"A row {k1} by thread")
    for j in range(16):                                    # stage-2 read / inverse stage-2 write
        banks_ok(lambda t: iA(t >> log_r3, j * L2 + (t & (L2 - 1))), f"A stride read j={j}")
    for k2 in range(16):                                   # stage-2 write / inverse stage-2 read
        banks_ok(lambda t: iB2(t >> log_r3, k2, t & (L2 - 1)), f"B2 k2={k2}")
    for e in range(16):                                    # stage-3 read / inverse write / zbuf write
        banks_ok(lambda t: iB3(t, e), f"B3 e={e}")
    # the two descriptions of exchange B address the same element at the same place, and nothing aliases
    seen = {}
    for t, e in itertools.product(range(T), range(16)):
        flat = 16 * t + e
        k1p, q = flat // L1, flat % L1
        k2, npp = q // L2, q % L2
        assert iB3(t, e) == iB2(k1p, k2, npp) == k1p * pitch + q + (q >> 4), (N, t, e)
        assert iB3(t, e) not in seen
        seen[iB3(t, e)] = flat
    assert max(seen) < 16 * pitch
    a_addrs = {iA(k1, x) for k1 in range(16) for x in range(L1)}
    assert len(a_addrs) == N and max(a_addrs) < 16 * pitch
    return pitch


if __name__ == "__main__":
    for N in (512, 1024, 2048, 4096):
        print(f"N={N}: padded layout conflict-free, PITCH={check(N)}, buffer {16 * 17 * (N // 256) + 4} float2 "
              f"(swizzled layout: {N + 4})")
