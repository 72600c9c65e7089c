"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the device sampler eadgan_b200/csrc/sample.cu (Philox4x32-10,
Salmon et al. SC'11; counter / key layout documented there), used by tests/test_sampling_gpu.py to check the
device stream word for word.  Known-answer vectors of Philox4x32-10 from the Random123 distribution
(kat_vectors: counter 0 / key 0, all-ones, and the pi digits) pin this restatement in tests/test_cpu.py."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4]; key: (k0, k1) -> uint32 [..., 4]"""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            hi0, lo0 = p0 >> np.uint64(32), p0 & mask
            hi1, lo1 = p1 >> np.uint64(32), p1 & mask
            c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def words(seed, step, stream, row0, rows, cols):
    """the uint32 word of every element of rows [row0, row0 + rows) of the global [*, cols] array"""
    first, count = row0 * cols, rows * cols
    g0, g1 = first >> 2, (first + count + 3) >> 2
    g = np.arange(g0, g1, dtype=np.uint64)
    ctr = np.stack([(g & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                    ((g >> np.uint64(32)).astype(np.uint32) | np.uint32(stream << 24)),
                    np.full(g.shape, step & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(g.shape, (step >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=-1)
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)
    return w[first - 4 * g0: first - 4 * g0 + count].reshape(rows, cols), w, first - 4 * g0, count


def uniform(seed, step, stream, row0, rows, cols, lo, hi):
    w, _, _, _ = words(seed, step, stream, row0, rows, cols)
    u = (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return (np.float32(hi - lo) * u + np.float32(lo)).astype(np.float32)   # fmaf: exact here up to 1 ulp (checked 1e-7)


def randint(seed, step, stream, row0, rows, cols, n):
    w, _, _, _ = words(seed, step, stream, row0, rows, cols)
    return ((w.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def normal(seed, step, stream, row0, rows, cols):
    _, allw, off, count = words(seed, step, stream, row0, rows, cols)
    p = allw.reshape(-1, 2)
    u1 = ((p[:, 0] >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (p[:, 1] >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    r = np.sqrt(-2.0 * np.log(u1))
    z = np.stack([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)], axis=-1).reshape(-1)
    return z[off: off + count].reshape(rows, cols)
