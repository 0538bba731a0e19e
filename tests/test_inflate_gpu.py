"""GPU parity: depth frames inflated on the device (one warp per frame, scannet_b200/csrc/inflate.cu) are byte-identical to
the host decode of the same .sens stream (which is pinned against the reference's stb decoder in test_sens_cpu.py) and to
zlib for every block type."""
import zlib

import numpy as np
import pytest
import torch

from scannet_b200 import sens, synth
from scannet_b200._lib import ScnError
from scannet_b200.sens import SensFile
from test_inflate_cpu import encoders, payloads

pytestmark = pytest.mark.gpu


def test_sens_stream_decodes_on_device(built, tmp_path):
    D, Cc, P, K = synth.make_frames(40, seed=3, width=160, height=120, loop_frames=80, noise_mm=2.0, drop=0.05)
    p = str(tmp_path / "s.sens")
    synth.write_sens(p, D, None, P, K, depth_comp=1, color_comp=0)
    f = SensFile(p)
    out = torch.zeros((40, 120, 160), dtype=torch.int16, device="cuda")
    f.decode_depth_device(0, 40, out.data_ptr())
    got = out.cpu().numpy().view(np.uint16)
    assert (got == D).all()
    assert all((f.depth(i) == got[i]).all() for i in range(40))
    # a window in the middle, and the raw (uncompressed) container type
    out2 = torch.zeros((7, 120, 160), dtype=torch.int16, device="cuda")
    f.decode_depth_device(11, 7, out2.data_ptr())
    assert (out2.cpu().numpy().view(np.uint16) == D[11:18]).all()
    with pytest.raises(ScnError, match="out of bounds"):
        f.decode_depth_device(38, 3, out2.data_ptr())
    q = str(tmp_path / "raw.sens")
    synth.write_sens(q, D[:5], None, P[:5], K, depth_comp=0, color_comp=0)
    g = SensFile(q); out3 = torch.zeros((5, 120, 160), dtype=torch.int16, device="cuda")
    g.decode_depth_device(0, 5, out3.data_ptr())
    assert (out3.cpu().numpy().view(np.uint16) == D[:5]).all()


@pytest.mark.parametrize("enc", list(encoders()))
def test_every_block_type_on_device(built, enc):
    f = encoders()[enc]
    raws = [r for r in payloads().values() if len(r) > 0]
    cap = max(len(r) for r in raws)
    padded = [r + bytes(cap - len(r)) for r in raws]                      # every stream must fill the frame
    out = torch.zeros((len(padded), cap), dtype=torch.uint8, device="cuda")
    sens.inflate_batch_device([f(r) for r in padded], cap, out.data_ptr())
    got = out.cpu().numpy()
    for i, r in enumerate(padded):
        assert got[i].tobytes() == r, (enc, i)


def test_full_resolution_batch(built):
    """640x480 frames, 300 streams in one launch (more warps than SMs), mixed encoders"""
    rng = np.random.default_rng(4)
    base = (rng.integers(400, 5000, (480, 640)) // 4 * 4).astype("<u2"); base[:, :100] = 0
    frames = [np.roll(base, 7 * i, axis=1).copy() for i in range(300)]
    encs = list(encoders().values())
    streams = [encs[i % len(encs)](fr.tobytes()) for i, fr in enumerate(frames)]
    out = torch.zeros((300, 480, 640), dtype=torch.int16, device="cuda")
    sens.inflate_batch_device(streams, 640 * 480 * 2, out.data_ptr())
    got = out.cpu().numpy().view(np.uint16)
    assert all((got[i] == frames[i]).all() for i in range(300))


def test_device_errors(built):
    raw = payloads()["depth"]; good = zlib.compress(raw)
    out = torch.zeros((3, len(raw)), dtype=torch.uint8, device="cuda")
    with pytest.raises(ScnError, match="corrupt zlib depth stream"):
        sens.inflate_batch_device([good, good[: len(good) // 2], good], len(raw), out.data_ptr())
    with pytest.raises(ScnError, match="need"):                          # stream shorter than the frame
        sens.inflate_batch_device([zlib.compress(raw[:1000])], len(raw), out.data_ptr())
    # a longer stream is cut at the frame size, like the host path
    sens.inflate_batch_device([zlib.compress(raw + b"tail" * 100)], len(raw), out.data_ptr())
    assert out[0].cpu().numpy().tobytes() == raw


def test_truncated_dynamic_block_is_an_error_on_the_device_too(built):
    """host and device must agree on a truncated stream (the device used to fill the frame from the zero padding)"""
    from test_inflate_cpu import truncated_dynamic_stream
    raw, bad = truncated_dynamic_stream()
    out = torch.zeros((2, len(raw)), dtype=torch.uint8, device="cuda")
    with pytest.raises(ScnError, match="truncated"):
        sens.inflate_batch_device([zlib.compress(raw, 6), bad], len(raw), out.data_ptr())


@pytest.mark.parametrize("window", ["ring", "hbm"])
def test_both_window_placements_give_the_same_bytes(built, window, monkeypatch):
    """the shared-memory-ring kernel (small launches) and the HBM-window kernel (large launches) are the same decoder with the
    deflate window in two places; SCN_INFLATE_WINDOW forces one"""
    monkeypatch.setenv("SCN_INFLATE_WINDOW", window)
    rng = np.random.default_rng(11)
    base = (rng.integers(400, 5000, (480, 640)) // 2 * 2).astype("<u2"); base[100:140] = 0
    frames = [np.roll(base, 13 * i, axis=1).copy() for i in range(48)]
    encs = list(encoders().values())
    streams = [encs[i % len(encs)](fr.tobytes()) for i, fr in enumerate(frames)]
    out = torch.zeros((len(frames), 480, 640), dtype=torch.int16, device="cuda")
    sens.inflate_batch_device(streams, 640 * 480 * 2, out.data_ptr())
    got = out.cpu().numpy().view(np.uint16)
    assert all((got[i] == frames[i]).all() for i in range(len(frames)))
    assert sens.inflate_last_timings()[2] == (window == "ring")
