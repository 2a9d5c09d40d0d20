"""CPU prototype of the single-binade integer box mean: inside one f32 binade the bit pattern is
linear in the value, so a window sum is an integer sum of bit patterns and the correctly rounded mean is an integer
division by N (N odd: no ties).  Checks, bit for bit against the oracle's exact box mean:
  * 1-D sliding sums on the raw bit patterns for N = 5, 17, 65 (vertical pass), reflect edges;
  * the 2-D chain (axis 0, f32 rounding, axis 1) when the intermediate stays in the binade;
  * the magic-multiply division floor(x / N) == (x * M) >> sh for every x < 2^31 (what the kernel would use).
Run: python scratch/int_binade_proto.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import terrain_oracle as orc

MAGIC = {5: (0xCCCCCCCD, 34), 17: (0xF0F0F0F1, 36), 65: (0xFC0FC0FD, 38)}


def box1d_bits(a: np.ndarray, n: int, axis: int) -> np.ndarray:
    """Exact box mean of width n along `axis` for an array inside ONE binade, computed on the bit patterns."""
    r = n // 2
    bits = a.view(np.int32).astype(np.int64)
    base = int(bits.min()) & ~0x7FFFFF                      # bit pattern of 2^e
    assert ((bits & ~0x7FFFFF) == base).all(), "not a single binade"
    d = np.moveaxis(bits - base, axis, 0)
    pad = np.concatenate([d[:r][::-1], d, d[-r:][::-1]], axis=0)     # scipy 'reflect' (edge-inclusive mirror)
    c = np.concatenate([np.zeros((1,) + pad.shape[1:], np.int64), np.cumsum(pad, axis=0)], axis=0)
    s = c[n:] - c[:-n]                                      # window sums, exact integers < n * 2^23 < 2^30
    m, sh = MAGIC[n]
    q = ((s + (n - 1) // 2) * m) >> sh                      # round-to-nearest division by the odd n
    out = (np.moveaxis(q, 0, axis) + base).astype(np.int32).view(np.float32)
    return out


def main():
    rng = np.random.default_rng(1)
    for n, (m, sh) in MAGIC.items():                        # exhaustive check of the magic division on [0, 2^31)
        for lo in range(0, 1 << 31, 1 << 26):
            x = np.arange(lo, lo + (1 << 26), dtype=np.uint64)
            assert np.array_equal((x * np.uint64(m)) >> np.uint64(sh), x // np.uint64(n)), (n, lo)
        print(f"magic division by {n}: exact on [0, 2^31)")
    for e_lo in (512.0, 1.0, 0.03125):
        a = (e_lo * (1.0 + rng.random((300, 257)))).astype(np.float32)
        a = np.minimum(a, np.nextafter(np.float32(2 * e_lo), np.float32(0)))
        for n in (5, 17, 65):
            want0 = orc.box_mean_exact(a, n)                # axis 0 then axis 1, f32 between (the oracle's 2-D form)
            v = box1d_bits(a, n, 0)
            h = box1d_bits(v, n, 1)                         # the vertical means stay in the binade of the inputs
            assert np.array_equal(h, want0), (e_lo, n, int((h != want0).sum()))
        print(f"binade [{e_lo}, {2*e_lo}): integer box means == exact box means for N = 5, 17, 65 (2-D, reflect)")
    # extremes: all-max mantissa rounds into the next binade cleanly
    top = np.full((70, 70), np.nextafter(np.float32(1024), np.float32(0)), np.float32)
    assert np.array_equal(box1d_bits(top, 65, 0), orc.box_mean_exact(top, 65))
    print("ok")


if __name__ == "__main__":
    main()
