"""NumPy models of the data layouts the particle kernels use (csrc/particles.cu), the counterpart of
tests/fused_model.py for the solve.  They restate WHAT the kernels compute, element by element, so that the
identities the CUDA code relies on can be checked on a CPU box against the oracle:

* k_deposit_tiles + k_fold_tiles: the particle does not update grid nodes but a per-cell accumulator
  T[ix, iy, iz][k], k = kx + 2*ky = the four (x, y) corners of cell (ix, iy) in plane iz (one 32-byte sector in
  Float64); a particle adds ((q*wx)*wy)*wz to T[ix, iy, iz] (plane iz) and T[ix, iy, iz+1] (plane iz + 1); the fold
  forms rho[i, j, k] = T[i, j, k][0] + T[i-1, j, k][1] + T[i, j-1, k][2] + T[i-1, j-1, k][3] in this fixed order.
* k_pack_efield_f64 + k_interpolate_pair2_f64: node-major records {Ex, Ey, Ez, 0}; two lanes share a particle, lane
  kx sums its four (y, z) corners with weights wx_k*wy*wz in the order (0,0), (1,0), (0,1), (1,1), and the result is
  (x0 part) + (x1 part).
* k_pack_efield_f32 + k_interpolate_packed_f32: records {E(i), 0, E(i+1), 0}; one thread per particle, the eight
  products summed left to right in the reference's order (src/interpolation.jl:56-85) -- bit-identical to it."""
import numpy as np


def locate(p, lo, delta, n):
    """Same arithmetic as `locate` in particles.cu: true division, floor, clamp of the cell index to [0, n-2]."""
    t = (p - lo) / delta
    fl = np.minimum(np.maximum(np.floor(t), 0), n - 2)
    return fl.astype(np.int64), t - fl


def deposit_tiles(grid, lo, delta, x, y, z, q, T=np.float64):
    """Returns (tiles, rho): tiles[ix, iy, iz, k] as accumulated by k_deposit_tiles, rho as folded by k_fold_tiles."""
    nx, ny, nz = grid
    W = np.promote_types(x.dtype, T)
    ix, fx = locate(x.astype(W), W.type(lo[0]), W.type(delta[0]), nx)
    iy, fy = locate(y.astype(W), W.type(lo[1]), W.type(delta[1]), ny)
    iz, fz = locate(z.astype(W), W.type(lo[2]), W.type(delta[2]), nz)
    one = W.type(1)
    tiles = np.zeros((nx, ny, nz, 4), dtype=T)
    for k in range(4):
        kx, ky = k & 1, k >> 1
        qxy = q.astype(W) * (fx if kx else one - fx) * (fy if ky else one - fy)     # (charge * w_x) * w_y
        np.add.at(tiles, (ix, iy, iz, np.full_like(ix, k)), (qxy * (one - fz)).astype(T))
        np.add.at(tiles, (ix, iy, iz + 1, np.full_like(ix, k)), (qxy * fz).astype(T))
    rho = tiles[..., 0].copy()
    rho[1:, :, :] += tiles[:-1, :, :, 1]
    rho[:, 1:, :] += tiles[:, :-1, :, 2]
    rho[1:, 1:, :] += tiles[:-1, :-1, :, 3]
    return tiles, rho


def pack_f64(efield):
    """efield[ix, iy, iz, c] -> records[ix, iy, iz, 4] = {Ex, Ey, Ez, 0}"""
    rec = np.zeros(efield.shape[:3] + (4,), dtype=np.float64)
    rec[..., :3] = efield
    return rec


def gather_pair_f64(records, lo, delta, x, y, z):
    """k_interpolate_pair2_f64: per-lane partial sums over the four (y, z) corners, then x0 part + x1 part."""
    nx, ny, nz = records.shape[:3]
    ix, dx = locate(x.astype(np.float64), lo[0], delta[0], nx)
    iy, dy = locate(y.astype(np.float64), lo[1], delta[1], ny)
    iz, dz = locate(z.astype(np.float64), lo[2], delta[2], nz)
    parts = []
    for kx in (0, 1):
        wx = dx if kx else 1.0 - dx
        w00 = wx * (1.0 - dy) * (1.0 - dz)
        w10 = wx * dy * (1.0 - dz)
        w01 = wx * (1.0 - dy) * dz
        w11 = wx * dy * dz
        n00 = records[ix + kx, iy, iz]
        n10 = records[ix + kx, iy + 1, iz]
        n01 = records[ix + kx, iy, iz + 1]
        n11 = records[ix + kx, iy + 1, iz + 1]
        parts.append(n00[:, :3] * w00[:, None] + n10[:, :3] * w10[:, None] + n01[:, :3] * w01[:, None] + n11[:, :3] * w11[:, None])
    out = parts[0] + parts[1]
    return out[:, 0], out[:, 1], out[:, 2]


def pack_f32(efield):
    """efield[ix, iy, iz, c] (Float32) -> records[ix, iy, iz, 8] = {E(ix), 0, E(ix+1), 0}; the x+1 half of the last
    node of a row is zero (never read: the cell index is clamped to n-2)."""
    rec = np.zeros(efield.shape[:3] + (8,), dtype=np.float32)
    rec[..., 0:3] = efield
    rec[:-1, :, :, 4:7] = efield[1:]
    return rec


def gather_packed_f32(records, lo, delta, x, y, z):
    """k_interpolate_packed_f32 in promote(P, Float32): the reference's eight products, summed left to right."""
    nx, ny, nz = records.shape[:3]
    W = np.promote_types(x.dtype, np.float32)
    ix, dx = locate(x.astype(W), W.type(lo[0]), W.type(delta[0]), nx)
    iy, dy = locate(y.astype(W), W.type(lo[1]), W.type(delta[1]), ny)
    iz, dz = locate(z.astype(W), W.type(lo[2]), W.type(delta[2]), nz)
    one = W.type(1)
    w = {(a, b, c): (dx if a else one - dx) * (dy if b else one - dy) * (dz if c else one - dz)
         for a in (0, 1) for b in (0, 1) for c in (0, 1)}
    out = []
    for comp in range(3):
        acc = None
        for c in (0, 1):            # order 000, 100, 010, 110, 001, 101, 011, 111 (src/interpolation.jl:56-85)
            for b in (0, 1):
                for a in (0, 1):
                    term = records[ix, iy + b, iz + c, 4 * a + comp].astype(W) * w[(a, b, c)]
                    acc = term if acc is None else acc + term
        out.append(acc.astype(x.dtype))
    return tuple(out)


# ---- cell-ordered regime (csrc/sorted.cu) -----------------------------------------------------------------------------
def cell_keys(grid, lo, delta, x, y, z, T=np.float64):
    """k_cell_keys: linear index of the (clamped) cell, ix + nx*(iy + ny*iz)"""
    nx, ny, nz = grid
    W = np.promote_types(x.dtype, T)
    ix, _ = locate(x.astype(W), W.type(lo[0]), W.type(delta[0]), nx)
    iy, _ = locate(y.astype(W), W.type(lo[1]), W.type(delta[1]), ny)
    iz, _ = locate(z.astype(W), W.type(lo[2]), W.type(delta[2]), nz)
    return (ix + nx * (iy + ny * iz)).astype(np.uint32)


def radix_sort_pairs(keys, key_bits, tile=4096, warp_keys=256):
    """launch_sort_pairs: LSD passes of ceil(bits / passes) bits; inside a tile every warp ranks its `warp_keys`
    consecutive keys in rounds of 32 (rank = keys of the same digit in earlier rounds / lower lanes), the warps' counts
    are prefix-summed per digit, and the tile's digits start at the exclusive scan of the digit-major histogram table.
    Returns the permutation (index of the key that comes i-th)."""
    n = len(keys)
    passes = max(1, (max(key_bits, 1) + 7) // 8)
    bits = (max(key_bits, 1) + passes - 1) // passes
    mask = (1 << bits) - 1
    ntiles = (n + tile - 1) // tile
    cur_k = keys.astype(np.uint32).copy()
    cur_v = np.arange(n, dtype=np.uint32)
    for p in range(passes):
        shift = p * bits
        digit = (cur_k >> np.uint32(shift)) & np.uint32(mask)
        hist = np.zeros((mask + 1, ntiles), dtype=np.int64)                  # digit-major table of k_radix_hist
        for t in range(ntiles):
            hist[:, t] = np.bincount(digit[t * tile:(t + 1) * tile], minlength=mask + 1)
        offs = (np.cumsum(hist.reshape(-1)) - hist.reshape(-1)).reshape(hist.shape)   # ONE exclusive scan
        out_k, out_v = np.empty_like(cur_k), np.empty_like(cur_v)
        for t in range(ntiles):
            d = digit[t * tile:(t + 1) * tile]
            m = len(d)
            nwarps = (m + warp_keys - 1) // warp_keys
            whist = np.zeros((nwarps, mask + 1), dtype=np.int64)
            rank = np.zeros(m, dtype=np.int64)
            for w in range(nwarps):
                seg = d[w * warp_keys:(w + 1) * warp_keys]
                for r0 in range(0, len(seg), 32):                             # one round: __match_any_sync groups
                    rnd = seg[r0:r0 + 32]
                    for lane, dv in enumerate(rnd):
                        lower = int(np.sum(rnd[:lane] == dv))
                        rank[w * warp_keys + r0 + lane] = whist[w, dv] + lower
                    for dv, c in zip(*np.unique(rnd, return_counts=True)):
                        whist[w, dv] += c                                     # the leader's returning atomicAdd
            wbase = np.cumsum(whist, axis=0) - whist                          # exclusive prefix over the warps
            warp_of = np.arange(m) // warp_keys
            pos = offs[d, t] + wbase[warp_of, d] + rank
            out_k[pos] = cur_k[t * tile:(t + 1) * tile]
            out_v[pos] = cur_v[t * tile:(t + 1) * tile]
        cur_k, cur_v = out_k, out_v
    return cur_v


def deposit_runs(grid, lo, delta, x, y, z, q, T=np.float64, per_lane=8):
    """k_deposit_runs: every lane walks `per_lane` consecutive particles keeping the corner sums of its current cell; an
    interloper (differs from the run, next particle back in it) is flushed on its own; the warp's open runs are combined
    by a segmented scan over adjacent lanes with equal cells and the last lane of each segment flushes.  Returns
    (rho, flushes): the grid and the number of 8-value flushes that reached memory."""
    nx, ny, nz = grid
    W = np.promote_types(x.dtype, T)
    ix, fx = locate(x.astype(W), W.type(lo[0]), W.type(delta[0]), nx)
    iy, fy = locate(y.astype(W), W.type(lo[1]), W.type(delta[1]), ny)
    iz, fz = locate(z.astype(W), W.type(lo[2]), W.type(delta[2]), nz)
    one = W.type(1)
    qq = q.astype(W)
    qx = (qq * (one - fx), qq * fx)
    wy, wz = (one - fy, fy), (one - fz, fz)
    vals = np.stack([(qx[a] * wy[b]) * wz[c] for c in (0, 1) for b in (0, 1) for a in (0, 1)], axis=1)   # ((q*wx)*wy)*wz
    cell = ix + nx * (iy + ny * iz)
    rho = np.zeros(nx * ny * nz, dtype=T)
    corner = np.array([a + nx * (b + ny * c) for c in (0, 1) for b in (0, 1) for a in (0, 1)])
    flushes = 0

    def flush(c, s):
        nonlocal flushes
        flushes += 1
        for k in range(8):
            rho[c + corner[k]] += T(s[k])

    n = len(x)
    for wbase in range(0, n, 32 * per_lane):
        cur = np.full(32, -1, dtype=np.int64)
        sums = np.zeros((32, 8), dtype=W)
        for lane in range(32):
            i0 = wbase + lane * per_lane
            cnt = min(per_lane, max(0, n - i0))
            for j in range(cnt):
                c = cell[i0 + j]
                nxt = cell[i0 + j + 1] if j + 1 < cnt else -2
                if c != cur[lane] and cur[lane] >= 0 and nxt == cur[lane]:
                    flush(c, vals[i0 + j])
                else:
                    if c != cur[lane] and cur[lane] >= 0:
                        flush(cur[lane], sums[lane])
                    sums[lane] = vals[i0 + j] if c != cur[lane] else sums[lane] + vals[i0 + j]
                    cur[lane] = c
        # segmented inclusive scan over the lanes (Hillis-Steele, 5 steps), tails flush
        start = np.zeros(32, dtype=np.int64)
        for lane in range(32):
            start[lane] = lane if lane == 0 or cur[lane] != cur[lane - 1] else start[lane - 1]
        o = 1
        while o < 32:
            prev = sums.copy()
            for lane in range(32):
                if lane - o >= start[lane]:
                    sums[lane] = prev[lane] + prev[lane - o]
            o <<= 1
        for lane in range(32):
            if cur[lane] >= 0 and (lane == 31 or cur[lane + 1] != cur[lane]):
                flush(cur[lane], sums[lane])
    return rho.reshape(nz, ny, nx).transpose(2, 1, 0), flushes
