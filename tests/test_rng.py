"""include/wrt_rng.h: the counter RNG behind the soft-shadow samples is Philox4x32-10
(known-answer vectors of the Random123 distribution), and the (u, v) mapping is the documented one."""
import ctypes as C

import numpy as np

import oracle_bindings as ob

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def philox(ctr, key):
    lib = ob.oracle()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.orc_philox4x32_10(c, k, o)
    return tuple(o)


def test_philox_known_answers():
    for ctr, key, want in KAT:
        assert philox(ctr, key) == want


def test_sample_uv_mapping():
    lib = ob.oracle()
    seed, pixel, path, light = 0x5EED, 123456, 5, 0
    for sample in range(50):
        uv = (C.c_float * 2)()
        lib.orc_light_sample_uv(seed, pixel, path, light, sample, uv)
        r = philox((pixel, path, light, sample >> 1), (seed, 0x57525421))
        a, b = (r[2], r[3]) if sample & 1 else (r[0], r[1])
        assert uv[0] == np.float32((a >> 8) / 16777216.0) and uv[1] == np.float32((b >> 8) / 16777216.0)
        assert 0.0 <= uv[0] < 1.0 and 0.0 <= uv[1] < 1.0
