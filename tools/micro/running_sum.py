"""CPU check of transport_grid_fast.cuh::CoordRun: the fp32 running position p += inc of the reference ray march
(heterogeneous.cpp:343-369) in closed form per binade, against sequential float32 accumulation."""
import numpy as np, math
f32 = np.float32
def seq(p0, inc, n):
    out = np.empty(n + 1, np.float32); p = f32(p0); out[0] = p
    for i in range(1, n + 1):
        p = f32(p + f32(inc)); out[i] = p
    return out
def closed(p0, inc, n):
    out = np.empty(n + 1, np.float32)
    p = f32(p0); inc = f32(inc); i = 0; runs = 0
    while i <= n:
        runs += 1
        a = abs(float(p))
        if a < 2.0 ** -60:
            out[i] = p; p = f32(p + inc); i += 1; continue
        mant, e = math.frexp(a); u = math.ldexp(1.0, e - 24)
        q = float(inc) / u
        fl = math.floor(q); frac = q - fl
        if frac == 0.5:
            # tie: ties-to-even keeps an even mantissa even, so the step is the even one of {fl, fl + 1} once p is even
            if (int(round(a / u)) & 1):              # odd mantissa: one true add makes it even (or leaves the binade)
                out[i] = p; p = f32(p + inc); i += 1; continue
            k = fl if (fl % 2 == 0) else fl + 1
        else:
            k = fl + 1 if frac > 0.5 else fl
        k = int(k)
        top = math.ldexp(1.0, e); bot = math.ldexp(1.0, e - 1)
        s = 1 if p > 0 else -1
        kk = k * s
        if kk > 0: m = int(((top - u) - a) / (kk * u))
        elif kk < 0: m = int((a - bot) / (-kk * u))
        else: m = n
        m = min(m, n - i)
        for j in range(m + 1):
            out[i + j] = f32(float(p) + (j * k) * u)
        i += m
        p = f32(out[i] + inc); i += 1
    return out, runs
rng = np.random.default_rng(0)
bad = 0; tot = 0; maxruns = 0
for t in range(20000):
    p0 = f32(rng.uniform(-0.001, 1.001)); n = int(rng.integers(2, 2000))
    length = rng.uniform(0.01, 1.7); d = rng.uniform(-1, 1)
    inc = f32(d * length / n)
    if t % 3 == 0: inc = f32(np.round(float(inc) * 2 ** 28) / 2 ** 28)      # few mantissa bits: many exact ties
    if not (-0.001 <= p0 + inc * n <= 1.001): continue
    a = seq(p0, inc, n); b, r = closed(p0, inc, n)
    tot += 1; maxruns = max(maxruns, r)
    if not np.array_equal(a, b):
        bad += 1
        if bad < 5:
            j = np.flatnonzero(a != b)[0]; print("mismatch", p0, inc, n, "first at", j, a[j], b[j], a[j-1])
print("cases", tot, "mismatching", bad, "max runs", maxruns)
