"""CPU: the GPU depth decoder's source (scannet_b200/csrc/inflate.cu) compiled for the host with one lane, against zlib
streams of every deflate block type.  The reference inflates with stb_image's zlib decoder (sensorData.h:703-709); any
conforming inflate yields the same bytes, so zlib is the ground truth here."""
import zlib

import numpy as np
import pytest

from scannet_b200 import sens
from scannet_b200._lib import ScnError


def payloads():
    rng = np.random.default_rng(0)
    depth = (rng.integers(500, 4000, (120, 160)) // 8 * 8).astype("<u2"); depth[rng.random(depth.shape) < 0.1] = 0; depth[:30] = 0
    return {"depth": depth.tobytes(), "empty": b"", "one": b"a", "zeros": bytes(100000), "noise": rng.integers(0, 256, 70000).astype(np.uint8).tobytes(),
            "period3": b"abcabcabcabc" * 1000 + b"xyz", "period7": bytes(range(7)) * 5000}


def encoders():
    def z(level): return lambda raw: zlib.compress(raw, level)
    def strat(st, wbits=15, level=6):
        def f(raw):
            co = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, st); return co.compress(raw) + co.flush()
        return f
    def multi(raw):                                     # several blocks of different types in one stream
        co = zlib.compressobj(6); out = b""
        for i in range(0, len(raw), 10000):
            out += co.compress(raw[i:i + 10000]) + co.flush(zlib.Z_FULL_FLUSH if (i // 10000) % 2 else zlib.Z_SYNC_FLUSH)
        return out + co.flush()
    return {"stored": z(0), "fast": z(1), "default": z(6), "best": z(9), "fixed": strat(zlib.Z_FIXED), "huffman_only": strat(zlib.Z_HUFFMAN_ONLY),
            "window512": strat(zlib.Z_DEFAULT_STRATEGY, 9, 9), "rle": strat(zlib.Z_RLE), "multi_block": multi}


@pytest.mark.parametrize("enc", list(encoders()))
def test_host_build_matches_zlib(built, enc):
    f = encoders()[enc]
    for name, raw in payloads().items():
        assert sens.inflate_host(f(raw), len(raw)) == raw, (enc, name)


def test_longer_stream_is_truncated_to_the_frame(built):
    raw = payloads()["depth"]
    with pytest.raises(ScnError, match="longer than the frame"):          # generic hook reports it ...
        sens.inflate_host(zlib.compress(raw), len(raw) - 100)


@pytest.mark.parametrize("damage", ["header", "truncated", "bad_block", "bad_distance", "garbage"])
def test_corrupt_streams_are_errors(built, damage):
    raw = payloads()["depth"]; good = bytearray(zlib.compress(raw, 6))
    if damage == "header": bad = bytes([0x79]) + bytes(good[1:])
    elif damage == "truncated": bad = bytes(good[: len(good) // 2])
    elif damage == "bad_block": bad = bytes(good[:2]) + bytes([0x07]) + bytes(good[3:])            # BTYPE = 3
    elif damage == "bad_distance":                                                                   # fixed block: length 3, distance 1 with no output yet
        bad = bytes([0x78, 0x01]) + bytes([0b00000011, 0b00000010, 0]) + bytes(8)
    else: bad = bytes(good[:2]) + bytes(np.random.default_rng(1).integers(0, 256, 500).astype(np.uint8))
    with pytest.raises(ScnError):
        sens.inflate_host(bad, len(raw))


def test_fuzz_against_zlib(built):
    """random payload structure x random encoder settings: the host build of the GPU decoder must reproduce zlib's input"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(seed=st.integers(0, 2**31 - 1), size=st.integers(0, 40000), alphabet=st.sampled_from([1, 2, 4, 16, 64, 256]),
           repeat=st.integers(0, 3), level=st.integers(0, 9), strategy=st.sampled_from([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]),
           wbits=st.integers(9, 15), memlevel=st.integers(1, 9))
    def run(seed, size, alphabet, repeat, level, strategy, wbits, memlevel):
        rng = np.random.default_rng(seed)
        raw = rng.integers(0, alphabet, size).astype(np.uint8).tobytes()
        for _ in range(repeat):                                       # long-distance repeats -> matches at many distances
            cut = int(rng.integers(0, len(raw) + 1)); raw = raw + raw[:cut]
        co = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
        assert sens.inflate_host(co.compress(raw) + co.flush(), len(raw)) == raw
    run()


def test_corrupted_streams_never_escape_the_output_buffer(built):
    """6000 mutated zlib streams (all block types) through the host build of the GPU decoder source: each call returns bytes or
    raises ScnError, and a canary behind the output buffer stays intact (the kernel runs the same code with the same bounds)."""
    import ctypes as C
    from scannet_b200._lib import lib
    L = lib()
    L.scn_inflate_host.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    rng = np.random.default_rng(11)
    raws = [payloads()[k] for k in ("depth", "period3", "noise")]
    seeds = [(f(r), len(r)) for r in raws for f in encoders().values()]
    buf = np.zeros(max(n for _, n in seeds) + 64, np.uint8)
    for it in range(6000):
        comp, cap = seeds[it % len(seeds)]
        b = bytearray(comp)
        for _ in range(int(rng.integers(1, 5))):
            p = int(rng.integers(2, len(b)))
            op = int(rng.integers(0, 4))
            if op == 0: b[p] ^= 1 << int(rng.integers(0, 8))
            elif op == 1: b[p] = int(rng.integers(0, 256))
            elif op == 2: del b[p:p + int(rng.integers(1, 9))]
            else: b = b[:p]
            if len(b) < 3: break
        buf[cap:cap + 64] = 0xA5
        n = C.c_size_t()
        rc = L.scn_inflate_host(bytes(b), len(b), buf.ctypes.data, cap, C.byref(n))
        assert rc in (0, -4), rc                                            # SCN_OK or SCN_ERR_FORMAT
        assert n.value <= cap and (buf[cap:cap + 64] == 0xA5).all()


def truncated_dynamic_stream():
    """A dynamic-Huffman stream cut in the middle whose continuation by zero padding keeps decoding (the GPU path
    accepted such frames while the host path rejected them).  Noise-free data whose all-zero code is a literal."""
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 4, 60000).astype(np.uint8).tobytes()            # 2-bit alphabet: short codes, zero bits decode to a literal
    good = zlib.compress(raw, 6)
    return raw, good[: len(good) // 3]


def test_truncated_dynamic_block_is_an_error_on_the_host(built):
    raw, bad = truncated_dynamic_stream()
    with pytest.raises(ScnError, match="truncated"):
        sens.inflate_host(bad, len(raw))
