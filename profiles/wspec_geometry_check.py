"""CPU check of the address arithmetic of the width-specialised integer scan (git branch r2-int-scan-wspec, k_int.cu
int_bits_fast_w): for T in {32, 64}, every W in 1..32, every step and lane, the field cut out by the kernel's immediates
equals the logical row the step covers, on chunks packed by the oracle's FastLanes restatement.
Run: PYTHONPATH=. python profiles/wspec_geometry_check.py"""
import numpy as np

from oracle.liquid_oracle import fl_pack_chunk


def brev2(x):
    return ((x & 1) << 1) | ((x >> 1) & 1)


def order(j):
    return (j & 7) * 4 + brev2(j >> 3)


rng = np.random.default_rng(0)
bad = 0
for T, U in ((64, np.uint64), (32, np.uint32)):
    for W in range(1, 33):
        vals = rng.integers(0, 1 << W, size=1024, dtype=np.uint64).astype(U)
        w32 = np.frombuffer(fl_pack_chunk(vals, W).tobytes(), dtype=np.uint32)
        dE = (W // 2) * 128 if W % 2 == 0 else ((W - 1) // 2) * 128 + 4
        dO = (W // 2) * 128 if W % 2 == 0 else ((W + 1) // 2) * 128 - 4
        for j in range(32):
            for lane in range(32):
                half = lane >> 4 if T == 64 else 0
                lbase = (lane & 15) * 8 if T == 64 else lane * 4
                baseE, baseO = lbase + (dE if half else 0), lbase + (dO if half else 0)
                r = ((j >> 3) * 8 + (j & 7)) if T == 64 else j
                b = r * W
                x, sh = b >> 5, b & 31

                def addr(y):
                    return ((baseO if y & 1 else baseE) + (y >> 1) * 128 + (y & 1) * 4) if T == 64 else lbase + y * 128

                w0 = int(w32[addr(x) // 4])
                w1 = int(w32[addr(x + 1) // 4]) if sh + W > 32 else 0
                bad += ((((w1 << 32) | w0) >> sh) & ((1 << W) - 1)) != int(vals[order(j) * 32 + lane])
print("mismatches", bad)
