"""Pure-Python Philox4x32-10 + Lemire rejection sampling, restating open_spiel_b200/csrc/common.cuh's
philox_uniform so the oracle can be driven with the same random decisions as the device rollout."""
M0, M1 = 0xD2511F53, 0xCD9E8D57
MASK = 0xFFFFFFFF


def philox4(seed, lane, ply, stream):
    c = [lane & MASK, (lane >> 32) & MASK, ply & MASK, stream & MASK]
    k = [seed & MASK, (seed >> 32) & MASK]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + 0x9E3779B9) & MASK, (k[1] + 0xBB67AE85) & MASK]
    return c


def philox_uniform(seed, lane, ply, n):
    thresh = ((1 << 32) - n) % n
    stream = 0
    while True:
        for r in philox4(seed, lane, ply, stream):
            m = r * n
            if (m & MASK) >= thresh:
                return m >> 32
        stream += 1
