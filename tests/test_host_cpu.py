"""CPU: host-side C++ (audiowmark_b200/host) against the oracle -- integer tables byte-exact -- plus
the boundary checks that need no GPU: every symbol of include/awm_b200.h is exported, the product
fails loudly without a CUDA device, WAV / raw stream round trips."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import awm_oracle as O
import awm_testlib as T
from audiowmark_b200 import capi, hostapi as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = O.Params()
KEYS = [O.Key(), O.Key.test_key(1), O.Key(bytes(range(16)), "k")]
HAS_GPU = os.path.exists("/dev/nvidiactl")


def test_capi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "awm_b200.h")).read()
    declared = set(re.findall(r"\b(awm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"awm_ctx"}
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    hlib = H.load()
    for name in H.EXPORTS:
        assert hasattr(hlib, name), name


@pytest.mark.skipif(HAS_GPU, reason="box has a GPU")
def test_no_cpu_fallback_without_device(tmp_path):
    with pytest.raises(capi.AwmError):
        capi.Context(0)
    x = T.noise(1.0)
    with pytest.raises(RuntimeError):
        H.add(x, T.PAYLOAD)
    with pytest.raises(RuntimeError):
        H.get(x)
    wav = tmp_path / "x.wav"
    O.write_wav16(str(wav), x)
    for args in (["add", str(wav), str(tmp_path / "y.wav"), "f0"], ["get", str(wav)]):
        p = subprocess.run([H.CLI_PATH] + args, capture_output=True, text=True)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr


def test_prng_words_and_noise():
    for key in KEYS:
        for seed, stream in ((0, 1), (0xF00F1234B00B5678, 5), (2225, 2)):
            r = O.Random(key, seed, stream)
            want = np.array([r() for _ in range(70)], np.uint64)
            assert np.array_equal(H.random_u64(key.aes_key, seed, stream, 70), want)
    # FIPS-197 / SP 800-38A pinned value (also printed by the reference's testrandom): zero key, stream 5
    assert H.random_u64(bytes(16), 0xF00F1234B00B5678, 5, 1)[0] == 0x8723958E3F2E0422
    assert np.array_equal(H.gen_noise(bytes(16), 5000), O.gen_noise(5000 / 2 / 44100.0 + 1e-9).reshape(-1)[:5000])


@pytest.mark.parametrize("ki", range(len(KEYS)))
def test_key_tables_byte_exact(ki):
    key = KEYS[ki]
    H.set_params()
    assert H.frames_per_block() == O.frames_per_block(P) == 2226
    assert H.n_coded_bits() == O.conv_code_size(O.A, 128) == 858
    for mode in (capi.MODE_BLOCK, capi.MODE_CLIP):
        ent, off = H.sync_table(key.aes_key, mode)
        went, woff = T.sync_entries(key, mode, P)
        assert ent.tobytes() == went.tobytes() and np.array_equal(off, woff)
    mix, order = H.mix_table(key.aes_key)
    wmix, worder = T.mix_entries(key, P)
    assert mix.tobytes() == wmix.tobytes() and np.array_equal(order, worder)
    for payload in (T.PAYLOAD, "f0", "ffffffffffffffffffffffffffffffff"):
        assert H.frame_mod(key.aes_key, payload).tobytes() == T.frame_mod_ab(key, payload, P).tobytes()


def test_conv_encoder():
    rng = np.random.default_rng(3)
    for n in (1, 16, 128):
        bits = rng.integers(0, 2, n)
        for bt in (O.A, O.B, O.AB):
            assert list(H.conv_encode(bt, bits)) == O.conv_encode(bt, [int(b) for b in bits])
    # src/testconvcode.cc:73-103: encode -> hard decode is the identity (decoder = oracle Viterbi here)
    msg = O.bit_str_to_vec("80f12381")
    for bt in (O.A, O.B, O.AB):
        enc = np.array(O.conv_encode(bt, msg), np.float32)
        assert O.conv_decode_soft(bt, enc)[0] == msg


def test_cli_wav_tools_match_oracle(tmp_path):
    """test-gen-noise / cut-start / test-info / test-snr of the CLI (no GPU involved)."""
    cli = H.CLI_PATH
    a = str(tmp_path / "a.wav")
    subprocess.check_call([cli, "test-gen-noise", a, "3", "44100"])
    x, rate, bits = O.read_wav(a)
    assert (rate, bits, x.shape) == (44100, 16, (3 * 44100, 2))
    assert np.array_equal(O.quantize_sndfile16(x), O.quantize_sndfile16(O.gen_noise(3)))
    b = str(tmp_path / "b.wav")
    subprocess.check_call([cli, "cut-start", a, b, "1000"])
    assert np.array_equal(O.read_wav(b)[0], x[1000:])
    assert subprocess.check_output([cli, "test-info", b, "frames"], text=True).strip() == str(3 * 44100 - 1000)
    y = O.int16_to_float(O.quantize_sndfile16(x * np.float32(0.99)))
    c = str(tmp_path / "c.wav")
    O.write_wav16(c, y)
    d = x.astype(np.float64) - y
    want = 10 * np.log10((x.astype(np.float64) ** 2).sum() / (d ** 2).sum())
    assert abs(float(subprocess.check_output([cli, "test-snr", a, c], text=True)) - want) < 1e-5


def test_raw_and_wav_pipe_streams(tmp_path):
    cli = H.CLI_PATH
    a = str(tmp_path / "a.wav")
    subprocess.check_call([cli, "test-gen-noise", a, "1", "44100"])
    x, _, _ = O.read_wav(a)
    # wav -> stdout wav-pipe -> file: 16 bit stays identical
    out = subprocess.run([cli, "cut-start", "--output-format", "wav-pipe", a, "-", "0"], capture_output=True)
    # cut-start has no format options in the reference either: it must be rejected the same way
    assert out.returncode == 1
    key = str(tmp_path / "k.key")
    subprocess.check_call([cli, "gen-key", key, "--name", 'a "b"'])
    txt = open(key).read()
    assert txt.startswith("# watermarking key for audiowmark\n\nkey ") and 'name "a \\"b\\""' in txt
    assert oct(os.stat(key).st_mode & 0o777) == "0o600"


def test_resampler_stream_counts_match_oracle():
    """host-side frame bookkeeping of the resampled paths (pure integer / double arithmetic, no GPU): streaming
    resampler availability, WavChunkLoader frame count, and the add loop's frame plan at other sample rates"""
    import random
    rnd = random.Random(5)
    for ratio in (44100 / 48000.0, 48000 / 44100.0, 44100 / 32000.0, 32000 / 44100.0, 44100 / 96000.0, 96000 / 44100.0, 0.4873, 1.0 / 1.01, 2.0, 0.5):
        for fed in [0, 1, 15, 16, 17, 31, 32, 33, 40, 100, 1024, 1025, 4096, 44100] + [rnd.randrange(1, 5_000_000) for _ in range(20)]:
            assert H.resample_stream_available(fed, ratio) == O.stream_avail(fed, ratio), (fed, ratio)
            assert H.resample_stream_frames(fed, ratio) == O.stream_out_count(fed, ratio), (fed, ratio)
    for no_lim in (False, True):
        H.set_params(test_no_limiter=no_lim)
        P = O.Params(test_no_limiter=no_lim)
        for rate in (8000, 22050, 32000, 33333, 48000, 88200, 96000):
            for n in (0, 1, 1023, 1024, 1025, 5000, rate, rate + 1, 7 * rate - 1, 361417, 6_400_000):
                assert H.resampled_add_plan(n, rate) == O.resampled_add_plan(n, rate, P), (rate, n, no_lim)
    H.set_params()


def test_short_payload_and_linear_tables_byte_exact():
    """--short 12/16/20 (block code in front of the convolutional code: different block length, sync positions and
    frame-mod tables) and --linear (un-mixed data frames): host tables against the oracle, block code round trip"""
    key = KEYS[0]
    rng = np.random.default_rng(11)
    try:
        for k, n, payload in ((12, 56, "abc"), (16, 61, "abcd"), (20, 65, "abcde")):
            H.set_params()
            H.set_short_payload(k)
            Pk = O.Params(payload_short=True, payload_size=k)
            assert H.frames_per_block() == O.frames_per_block(Pk) == 510 + 2 * 6 * (n + 15)
            assert H.n_coded_bits() == O.code_size(O.A, Pk)
            for _ in range(20):
                msg = rng.integers(0, 2, k)
                enc = H.short_encode(msg)
                assert list(enc) == O.short_encode_blk([int(b) for b in msg], k) and len(enc) == n
                assert list(H.short_decode(enc)) == [int(b) for b in msg]
                bad = enc.copy()
                bad[rng.integers(0, n)] ^= 1                       # minimum distance >= 20: a single bit error is no code word
                assert len(H.short_decode(bad)) == 0 and O.short_decode_blk([int(b) for b in bad], k) == []
            ent, off = H.sync_table(key.aes_key, capi.MODE_BLOCK)
            went, woff = T.sync_entries(key, capi.MODE_BLOCK, Pk)
            assert ent.tobytes() == went.tobytes() and np.array_equal(off, woff)
            assert H.frame_mod(key.aes_key, payload).tobytes() == T.frame_mod_ab(key, payload, Pk).tobytes()
        H.set_short_payload(0)
        H.set_params(mix=False)
        Pl = O.Params(mix=False)
        assert H.frame_mod(key.aes_key, T.PAYLOAD).tobytes() == T.frame_mod_ab(key, T.PAYLOAD, Pl).tobytes()
        mix, order = H.mix_table(key.aes_key)                       # --linear: the same triples, not shuffled
        udg, bpg = O.UpDownGen(key, O.STREAM_DATA_UP_DOWN, Pl), O.BitPosGen(key, Pl)
        want = [(bpg.data_frame(f), u, d) for f in range(O.mark_data_frame_count(Pl)) for u, d in zip(*udg.get(f))]
        assert [(int(m["frame"]), int(m["up"]), int(m["down"])) for m in mix] == [(int(a), int(b), int(c)) for a, b, c in want]
    finally:
        H.set_short_payload(0)
        H.set_params()


def test_cli_option_handling_of_speed_and_short_modes(tmp_path):
    """argv handling that needs no GPU (src/audiowmark.cc:389-397, 665-674, 831-854): option conflicts, unsupported sizes, test-speed"""
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audiowmark_b200", "bin", "audiowmark")
    p = subprocess.run([cli, "get", "--detect-speed", "--try-speed", "1.01", "x.wav"], capture_output=True, text=True)
    assert p.returncode == 1 and "can only use one option: --detect-speed or --detect-speed-patient or --try-speed" in p.stderr
    p = subprocess.run([cli, "add", "--short", "13", "a.wav", "b.wav", "abc"], capture_output=True, text=True)
    assert p.returncode == 1 and "unsupported short payload size 13" in p.stderr
    p = subprocess.run([cli, "get", "--strength", "15", "x.wav"], capture_output=True, text=True)      # add-only option in the reference
    assert p.returncode == 1 and "unsupported option '--strength' for command 'get'" in p.stderr
    # test-speed: one keyed PRNG draw mapped to [0.85, 1.15] (tests/detect-speed-test.sh's companion command)
    for seed in (0, 1, 42):
        out = subprocess.check_output([cli, "test-speed", "--test-key", "7", str(seed)], text=True)
        r = O.Random(O.Key.test_key(7), seed, O.STREAM_DATA_UP_DOWN)
        assert out == "%.6f\n" % (0.85 + (r() / float(2 ** 64 - 1)) * (1.15 - 0.85))


def test_result_set_merge_and_sort_against_their_definitions():
    """the indexed chunk merge and the tuple sort of host/awm_results.cc vs the literal definitions (scan with approx_match, compare
    key by key; src/wmget.cc:268-312) on random chunk results with colliding positions, combined patterns, two keys, stretched speeds"""
    import ctypes
    L = H.load()
    L.awmh_selftest_results.restype = ctypes.c_int
    for seed in (1, 2, 3, 4):
        assert L.awmh_selftest_results(ctypes.c_uint64(seed), ctypes.c_int(200)) == 0
