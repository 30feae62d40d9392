"""Shared helpers for the parity tests: oracle tables -> C-ABI structs, synthetic signals."""
from __future__ import annotations

import numpy as np

import awm_oracle as O
from audiowmark_b200 import capi

PAYLOAD = "0123456789abcdef0011223344556677"


def sync_entries(key, mode, P):
    sb = O.get_sync_bits(key, mode, P)
    ent = np.zeros(len(sb.frame), capi.SYNC_ENTRY)
    ent["frame"] = sb.frame
    ent["up"] = sb.up.reshape(-1, P.bands_per_frame)
    ent["down"] = sb.down.reshape(-1, P.bands_per_frame)
    return ent, sb.off.astype(np.int32)


def mix_entries(key, P):
    m = O.gen_mix_entries(key, P)
    ent = np.zeros(len(m), capi.MIX_ENTRY)
    ent["frame"] = [e[0] for e in m]
    ent["up"] = [e[1] for e in m]
    ent["down"] = [e[2] for e in m]
    n_coded = O.conv_code_size(O.A, P.payload_size)
    order = np.array(O.randomize_bit_order(key, list(range(n_coded)), True), np.uint16)
    return ent, order


def frame_mod_ab(key, payload, P):
    bitvec = O.parse_payload(payload, P)
    return np.stack([O.init_frame_mod(key, 0, bitvec, P), O.init_frame_mod(key, 1, bitvec, P)])


def setup_ctx(ctx, key, P, payload=None, key_slot=0):
    """Feed one key's tables (built by the ORACLE, so kernels are tested independently of host/ table code)."""
    for mode in (capi.MODE_BLOCK, capi.MODE_CLIP):
        ent, off = sync_entries(key, mode, P)
        ctx.set_sync_tables(key_slot, mode, ent, off)
    ent, order = mix_entries(key, P)
    ctx.set_mix_tables(key_slot, ent, order, P.frames_per_bit, O.frames_per_block(P))
    if payload is not None:
        ctx.set_embed_tables(frame_mod_ab(key, payload, P))


def noise(seconds, channels=2, seed=1234, amp=0.5):
    rng = np.random.default_rng(seed)
    n = int(seconds * 44100)
    return ((rng.random((n, channels), dtype=np.float32) - 0.5) * (2 * amp)).astype(np.float32)


def rms(x):
    x = np.asarray(x, np.float64)
    return float(np.sqrt(np.mean(x * x))) if x.size else 0.0


_SPEED_CACHE = {}


def watermarked_noise(seconds, payload="f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0"):
    """reference test signal: keyed noise (test-gen-noise), watermarked by the oracle, on the 16 bit grid"""
    k = ("wm", seconds, payload)
    if k not in _SPEED_CACHE:
        x = O.int16_to_float(O.quantize_sndfile16(O.gen_noise(seconds)))
        _SPEED_CACHE[k] = O.int16_to_float(O.quantize_sndfile16(O.embed(x, O.Key(), payload, O.Params()).samples))
    return _SPEED_CACHE[k]


def cli_float(v):
    """the reference CLI parses every floating point argument with strtof (src/audiowmark.cc:188-199)"""
    return float(np.float32(v))


def speed_changed(seconds, speed):
    """tests/detect-speed-test.sh input: test-change-speed = resample_ratio (1 / speed), saved as 16 bit"""
    k = ("sp", seconds, speed)
    if k not in _SPEED_CACHE:
        y = watermarked_noise(seconds)
        _SPEED_CACHE[k] = O.int16_to_float(O.quantize_sndfile16(O.resample_ratio(y, 1 / cli_float(speed))))
    return _SPEED_CACHE[k]
