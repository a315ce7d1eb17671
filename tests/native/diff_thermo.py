import ctypes, numpy as np, sys, random
ref = ctypes.CDLL('/root/repo/oracle/_ref/libpcramp_ref.so')
hst = ctypes.CDLL('/root/repo/build/libhost_thermo.so')
def pack(strs):
    buf = np.zeros((len(strs), 33), dtype=np.uint8)
    for i, s in enumerate(strs):
        buf[i, :len(s)] = np.frombuffer(s.encode(), dtype=np.uint8)
    return buf
def run_ref(op, A, B, salt, strand):
    n = len(A); a = pack(A); b = pack(B) if B is not None else None
    out = np.zeros((n, 5), dtype=np.float32)
    st = np.ascontiguousarray(strand, dtype=np.float32)
    rc = ref.ref_thermo_batch(op, n, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p) if b is not None else None,
        ctypes.c_float(salt), st.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out
def run_host(op, A, B, salt, strand1):
    n = len(A); a = pack(A); b = pack(B) if B is not None else None
    out = np.zeros((n, 4), dtype=np.float32)
    st = np.ascontiguousarray(strand1, dtype=np.float32)
    cells = ctypes.c_longlong(0)
    rc = hst.host_thermo_batch(op, n, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p) if b is not None else None,
        ctypes.c_float(salt), st.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cells))
    assert rc == 0
    return out
def eff_strand(sa, sb):
    sa = np.float32(sa); sb = np.float32(sb)
    return np.where(sa > sb, sa - np.float32(0.5) * sb, sb - np.float32(0.5) * sa).astype(np.float32)
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
def rnd(lo=6, hi=32, alpha='ACGT'):
    return ''.join(rng.choice(alpha) for _ in range(rng.randint(lo, hi)))
def rc(s): return s[::-1].translate(str.maketrans('ACGT', 'TGCA'))
def mutate(s, k):
    s = list(s)
    for _ in range(k):
        i = rng.randrange(len(s)); r = rng.random()
        if r < 0.6: s[i] = rng.choice('ACGT')
        elif r < 0.8 and len(s) > 8: del s[i]
        elif len(s) < 32: s.insert(i, rng.choice('ACGT'))
    return ''.join(s)
salt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
for op in (0, 1, 2, 5, 3, 4):
    A = []; B = []
    for i in range(N):
        kind = rng.random()
        if kind < 0.4: a = rnd(15, 32)
        elif kind < 0.6: a = rnd(6, 32, rng.choice(['ACGT', 'AT', 'GC', 'ACGTGC']))
        elif kind < 0.8:
            h = rnd(4, 12); a = (rnd(0, 4) + h + rnd(3, 8) + mutate(rc(h), rng.randint(0, 2)) + rnd(0, 4))[:32]
        else:
            h = rnd(5, 15); a = (h + mutate(rc(h), rng.randint(0, 3)))[:32]
        if op in (3, 4):
            kb = rng.random()
            b = mutate(rc(a), rng.randint(0, 5))[:32] if kb < 0.6 else rnd(15, 32)
        else: b = a
        if len(a) < 5: a = a + 'ACGTA'
        A.append(a); B.append(b)
    sa = np.array([rng.choice([9e-7, 9e-7 / 4, 2e-7, 1e-6]) for _ in range(N)], dtype=np.float32)
    sb = np.array([rng.choice([9e-7, 9e-7 / 2, 1e-6]) for _ in range(N)], dtype=np.float32)
    strand = np.stack([sa, sb], 1)
    r = run_ref(op, A, B if op in (3, 4) else None, salt, strand)
    s1 = eff_strand(sa, sb) if op in (3, 4) else sa
    hop = {0: 0, 1: 1, 2: 2, 5: 5, 3: 3, 4: 4}[op]
    h = run_host(hop, A, B if op in (3, 4) else None, salt, s1)
    rr = r[:, [0, 1, 2, 4]]
    bad = np.where((rr.view(np.uint32) != h.view(np.uint32)).any(1))[0]
    print('op', op, 'n', N, 'mismatch', len(bad), 'nonzero tm', int((r[:, 0] > 0).sum()))
    for i in bad[:5]:
        print('   ', A[i], B[i], strand[i], 'ref', rr[i], 'host', h[i])
