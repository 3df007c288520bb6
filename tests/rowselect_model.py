"""Host model of topn_rowselect_kernel's LOGIC (csrc/topn.cu): float4 strides over 512 threads, group extremes, 4-thread
leader merge, ranking of the 128 leaders on 32-bit stand-ins with a truncated threshold, survivors kept by the raw-value
bound, exact re-sweeps with a raised threshold after an overflow of the survivor buffer.  Test infrastructure: it lets
the CPU suite check the selection algorithm (not the CUDA code) against numpy's stable argsort, including the overflow
path, which a small `cap` forces."""
import numpy as np

NT, NL = 512, 128


def ord32(v):
    v = np.float32(v) + np.float32(0.0)
    b = int(np.array(v, np.float32).view(np.uint32))
    return b ^ (0xffffffff if b >> 31 else 0x80000000)


def make_key(v, idx, desc):
    u, t = ord32(v), int(idx)
    if not desc:
        u, t = (~u) & 0xffffffff, (~int(idx)) & 0xffffffff
    return (u << 32) | t


def decode(key, desc):
    t, u = key & 0xffffffff, key >> 32
    if not desc:
        t, u = (~t) & 0xffffffff, (~u) & 0xffffffff
    b = (u ^ 0x80000000) if (u & 0x80000000) else ((~u) & 0xffffffff)
    return t, np.array(b, np.uint32).view(np.float32)


def raw_bound(T, desc):
    if T == 0:
        return np.float32(-np.inf if desc else np.inf)
    return decode(T, desc)[1]


def rowselect(row, listed, only_listed, desc, n, cap=1024):
    row = row.astype(np.float32).copy()
    c = len(row)
    cr = (c + 3) & ~3
    row = np.concatenate([row, np.full(cr - c, np.nan, np.float32)])
    T = 0
    with np.errstate(invalid="ignore"):
        if not only_listed:
            for i in listed:
                if 0 <= i < c:
                    row[i] = np.nan
            r4 = row.reshape(-1, 4)
            ext = np.fmax.reduce(r4, axis=1) if desc else np.fmin.reduce(r4, axis=1)
            tk = []
            for tid in range(NT):
                best, bv = np.float32(-np.inf if desc else np.inf), -1
                for v in range(tid, cr // 4, NT):
                    g = ext[v]
                    if (g >= best) if desc else (g < best):
                        best, bv = g, v
                k = 0
                if bv >= 0:
                    q = r4[bv]
                    if desc:
                        j = 3 if q[3] == best else 2 if q[2] == best else 1 if q[1] == best else 0
                    else:
                        j = 0 if q[0] == best else 1 if q[1] == best else 2 if q[2] == best else 3
                    k = make_key(best, 4 * bv + j, desc)
                tk.append(k)
            lead = [max(tk[4 * g: 4 * g + 4]) for g in range(NL)]
            a = [((k >> 32) & ~0x7f & 0xffffffff) | g for g, k in enumerate(lead)]
            for g in range(NL):
                if sum(x > a[g] for x in a) == n - 1:
                    uw = a[g] & ~0x7f
                    T = (uw << 32) if uw > 0x007fffff else 0
        rounds = 0
        while True:
            zb = raw_bound(T, desc)
            keys, cnt, seen = [], 0, set()
            it = [i for i in listed if 0 <= i < c] if only_listed else range(c)
            for e in it:
                if only_listed:
                    if e in seen:
                        continue
                    seen.add(e)
                x = row[e]
                if not ((x >= zb) if desc else (x <= zb)):
                    continue
                k = make_key(x, e, desc)
                if rounds > 0 and k < T:
                    continue
                if cnt < cap:
                    keys.append(k)
                cnt += 1
            if cnt <= cap:
                break
            rounds += 1
            T = sorted(keys, reverse=True)[n - 1]
    keys.sort(reverse=True)
    out = [decode(k, desc)[0] for k in keys[:n]]
    return out, min(n, cnt), rounds, cnt


def expect(row, listed, only_listed, desc, n):
    c = len(row)
    order = row.argsort(kind="stable")
    if desc:
        order = order[::-1]
    inm = np.zeros(c, bool)
    inm[[i for i in listed if 0 <= i < c]] = True
    return [int(i) for i in order if inm[i] == only_listed][:n]
