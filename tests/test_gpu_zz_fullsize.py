"""GPU, BASELINE.json's full size (1 h stereo 44.1 kHz, configs[1]): size-independent properties of the path -- the oracle needs
minutes per minute of audio at this size, so parity is shown through invariants:
  * add -> get round trip: every real detection carries the payload, one A or B block every 51.7 s, the "all" pattern on top
  * the three ways audio reaches the kernels (device pointer, pinned host fp32 with pipelined copies, 16 bit PCM converted on
    the device) give identical results
  * digital silence stays digital silence (no watermark in zero frames, limiter idle)
  * embedding is local: changing the last minute of the input leaves the first 58 minutes of the output bit identical
(named zz so that it runs after the parity tests proper)."""
import numpy as np
import pytest

import awm_testlib as T
from audiowmark_b200 import hostapi as H

pytestmark = pytest.mark.gpu
RATE = 44100
N = 60 * 60 * RATE


def to_s16_like_the_reference(torch, f):
    """float -> 16 bit as the reference writes a WAV file: float_to_int_clip<32> (multiply by 2^31, clip, truncate toward zero,
    src/rawconverter.hh:34-50), then the 16 most significant bits (arithmetic shift, src/sfoutputstream.cc:148-155)"""
    sn = f * 2147483648.0                                   # exact scaling
    v = torch.trunc(sn).to(torch.int64)
    v = torch.where(sn >= 2147483648.0, torch.full_like(v, 2147483647), v)
    v = torch.where(sn <= -2147483648.0, torch.full_like(v, -2147483648), v)
    return (v >> 16).to(torch.int16)


@pytest.fixture(scope="module")
def hour():
    torch = pytest.importorskip("torch")
    g = torch.Generator(device="cuda")
    g.manual_seed(77)
    x = torch.rand((N, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5
    y = torch.empty_like(x)
    torch.cuda.synchronize()
    H.set_params()
    H.add(x.data_ptr(), T.PAYLOAD, None, y.data_ptr(), N, 2)
    H.synchronize()
    return torch, x, y


def test_round_trip_one_hour(hour):
    torch, x, y = hour
    doc = H.get(y.data_ptr(), n_frames=N, channels=2)
    real = [m for m in doc["matches"] if m["quality"] > 0.35]
    assert doc["length"] == "60:00"
    assert all(m["bits"] == T.PAYLOAD for m in real)
    blocks = [m for m in real if m["type"] in ("A", "B")]
    assert 66 <= len(blocks) <= 70                    # 3600 s / 51.69 s per block, minus the partial blocks at both ends
    assert sum(m["type"] == "ALL" for m in real) == 1 and any(m["type"] == "AB" for m in real)
    # watermark energy: the embedded signal differs from the input by about -30 dB relative (strength 10)
    d = (y - x)
    snr = 10 * float(torch.log10((x * x).sum() / (d * d).sum()))
    assert 28.0 < snr < 36.0


def test_device_host_and_16bit_paths_agree_at_full_size(hour):
    torch, x, y = hour
    doc_dev = H.get(y.data_ptr(), n_frames=N, channels=2)
    yp = torch.empty((N, 2), dtype=torch.float32).pin_memory()
    yp.copy_(y)
    torch.cuda.synchronize()
    assert H.get(yp.numpy()) == doc_dev
    # pipelined host add == device add, bit for bit
    xp = torch.empty((N, 2), dtype=torch.float32).pin_memory()
    xp.copy_(x)
    out = torch.empty((N, 2), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    H.add(xp.numpy(), T.PAYLOAD, None, out.numpy())
    assert torch.equal(out, yp)
    # 16 bit: some 16 bit input, then the s16 entry points against the float entry points fed with the same audio and the
    # reference's float -> int16 rule applied to their output
    x16 = torch.floor(x * 32768.0).clamp_(-32768, 32767).to(torch.int16)
    xq = x16.to(torch.float32) * (1.0 / 32768.0)
    yq = torch.empty_like(xq)
    torch.cuda.synchronize()
    H.add(xq.data_ptr(), T.PAYLOAD, None, yq.data_ptr(), N, 2)
    H.synchronize()
    want16 = to_s16_like_the_reference(torch, yq).cpu()
    x16p = torch.empty((N, 2), dtype=torch.int16).pin_memory()
    x16p.copy_(x16)
    y16p = torch.empty((N, 2), dtype=torch.int16).pin_memory()
    torch.cuda.synchronize()
    H.add_s16(x16p.numpy(), T.PAYLOAD, None, y16p.numpy())
    assert torch.equal(y16p, want16)
    y16f = (want16.to(torch.float32) * (1.0 / 32768.0)).cuda()
    torch.cuda.synchronize()
    assert H.get_s16(y16p.numpy()) == H.get(y16f.data_ptr(), n_frames=N, channels=2)


def test_silence_at_full_size(hour):
    torch, x, y = hour
    z = torch.zeros((N, 2), device="cuda", dtype=torch.float32)
    zo = torch.empty_like(z)
    torch.cuda.synchronize()
    H.add(z.data_ptr(), T.PAYLOAD, None, zo.data_ptr(), N, 2)
    H.synchronize()
    assert not bool(zo.any())
    doc = H.get(zo.data_ptr(), n_frames=N, channels=2)
    # like the reference on digital silence (tests/golden/golden_large.json: silence170): every sync quality is exactly 0, the n-best
    # rule still hands blocks to the decoder, whose soft bits are 0/0 -> all-zero payloads with error -1/858
    assert len(doc["matches"]) > 0
    assert all(m["quality"] == 0 and m["bits"] == "0" * 32 and m["error"] == -0.001166 for m in doc["matches"])


def test_locality_at_full_size(hour):
    """the synthesis window reaches one frame, the limiter one block (1 s) beyond a change"""
    torch, x, y = hour
    x2 = x.clone()
    x2[59 * 60 * RATE:] *= 0.5
    y2 = torch.empty_like(x2)
    torch.cuda.synchronize()
    H.add(x2.data_ptr(), T.PAYLOAD, None, y2.data_ptr(), N, 2)
    H.synchronize()
    keep = 58 * 60 * RATE
    assert torch.equal(y2[:keep], y[:keep])
    assert not torch.equal(y2[59 * 60 * RATE + RATE:], y[59 * 60 * RATE + RATE:])
