"""The coverage kernel's exact shortcuts (csrc/ifx_cov.h) against the row-based Damerau they stand in for, through the test-only
emulation library: `damerau1_ascii` must agree with `damerau(.., 1)` on "is it <= 1, and if so which value", and a signature
rejection must imply that the row-based function also says "> maxd"."""
import ctypes, itertools, random
import numpy as np
from parity_util import emu_lib


def _lib():
    lib = ctypes.CDLL(emu_lib())
    u16p = ctypes.POINTER(ctypes.c_uint16)
    lib.ifx_emu_damerau.argtypes = [u16p, ctypes.c_int, u16p, ctypes.c_int, ctypes.c_int]
    lib.ifx_emu_damerau1_ascii.argtypes = [u16p, ctypes.c_int, u16p, ctypes.c_int]
    lib.ifx_emu_sig_far.argtypes = [u16p, ctypes.c_int, u16p, ctypes.c_int, ctypes.c_int]
    return lib


def _buf(s):
    a = np.array([ord(c) for c in s] or [0], dtype=np.uint16)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), len(s)


def _pairs():
    rng = random.Random(7)
    alpha = "abAB1"                                   # small alphabet: near misses are frequent; both cases; a non-letter
    for n in range(0, 5):                             # exhaustive over short strings
        for m in range(max(0, n - 2), n + 3):
            for s in itertools.product("abA", repeat=n):
                for t in itertools.product("abA", repeat=m):
                    yield "".join(s), "".join(t)
    words = ["gugruwo", "gufinsilve", "batman", "Batmna", "spider-man", "SPIDERMAN", "matrix", "matirx", "the", "teh", "a", ""]
    for s in words:
        for t in words:
            yield s, t
    for _ in range(60000):                            # random edits of random words
        n = rng.randint(1, 12); s = [rng.choice(alpha) for _ in range(n)]; t = list(s)
        for _ in range(rng.randint(0, 3)):
            op = rng.randint(0, 3); i = rng.randint(0, max(len(t) - 1, 0))
            if op == 0 and t: t[i] = rng.choice(alpha)
            elif op == 1: t.insert(i, rng.choice(alpha))
            elif op == 2 and t: del t[i]
            elif op == 3 and len(t) > 1 and i + 1 < len(t): t[i], t[i + 1] = t[i + 1], t[i]
        yield "".join(s), "".join(t)


def test_linear_damerau1_and_signature_reject_agree_with_the_row():
    lib = _lib(); n = 0
    for s, t in _pairs():
        sa, sp, sn = _buf(s); ta, tp, tn = _buf(t)
        ref1 = lib.ifx_emu_damerau(sp, sn, tp, tn, 1); fast = lib.ifx_emu_damerau1_ascii(sp, sn, tp, tn)
        assert (ref1 <= 1) == (fast <= 1) and (ref1 > 1 or ref1 == fast), (s, t, ref1, fast)
        for maxd in (1, 2):
            if lib.ifx_emu_sig_far(sp, sn, tp, tn, maxd):          # ASCII pairs: threshold maxd
                assert lib.ifx_emu_damerau(sp, sn, tp, tn, maxd) > maxd, (s, t, maxd)
        n += 1
    assert n > 60000
