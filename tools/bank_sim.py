"""Shared-memory bank conflicts of the bilinear gather, simulated on the CPU (no GPU needed).

A warp's 32 lanes sit on 32 consecutive columns of one row; lane i reads the staged DEM tile at column
i + floor(dx_i), row floor(dy_i) with dx, dy ~ N(0, sigma^2) clipped to +-8 (SURVEY 8d's synthetic offsets).  Counts the
data-stage wavefronts of one warp-wide load for the layouts that were considered ("measured and
rejected"): 32-bit loads from the fp32 tile (what the kernels do), 16-bit loads from the bf16 tile, and the
vertically paired layouts (v[r][q], v[r+1][q]) read with one 64-bit (fp32) / 32-bit (bf16) load per column, which
halve the number of loads per tap.  64-bit requests are served one half-warp at a time.
The fp32 / bf16 figures reproduce what ncu measures on the kernels (3.06 / 2.5 wavefronts per load at sigma 1.5).
"""
import numpy as np

SW = 144


def wavefronts(addr, banks):
    cnt = {}
    for a in set(addr.tolist()):
        cnt[a % banks] = cnt.get(a % banks, 0) + 1
    return max(cnt.values())


def sim(sigma, n=20000, seed=0):
    rng = np.random.default_rng(seed)
    dx = np.clip(rng.normal(0, sigma, (n, 32)), -8, 8)
    dy = np.clip(rng.normal(0, sigma, (n, 32)), -8, 8)
    q = np.floor(np.arange(32) + 9 + dx).astype(int)
    r = np.floor(10 + dy).astype(int)
    w32 = w64 = w16 = 0
    for i in range(n):
        a = r[i] * SW + q[i]
        w32 += wavefronts(a, 32)                                        # fp32 tile, LDS.32 (also: bf16 pairs, LDS.32)
        w64 += wavefronts(a[:16], 16) + wavefronts(a[16:], 16)          # fp32 vertical pairs, LDS.64, per half-warp
        w16 += wavefronts(a // 2, 32)                                   # bf16 tile, LDS.U16 (two columns per word)
    return w32 / n, w64 / n, w16 / n


if __name__ == "__main__":
    print("sigma | fp32 tile: 4 x LDS.32 per tap | fp32 vertical pairs: 2 x LDS.64 | bf16 tile: 4 x LDS.U16 | bf16 vertical pairs: 2 x LDS.32")
    for s in (0.5, 1.5, 4.0):
        a, b, c = sim(s)
        print(f"{s:5.1f} | {a:.2f} per load, {4 * a:5.2f} per tap | {b:.2f}, {2 * b:5.2f} | {c:.2f}, {4 * c:5.2f} | {a:.2f}, {2 * a:5.2f}")
