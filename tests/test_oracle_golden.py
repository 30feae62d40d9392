"""CPU: pins the oracle against the reference's own outputs.

tests/golden/golden.json was produced by tests/golden/make_golden.py with the unmodified reference
sources (oracle/_ref).  The oracle has to reproduce them exactly: PCM hashes bit-for-bit, pattern
lists with every printed digit.  These are the "golden vectors" of the path -- the reference tree
itself ships none."""
import hashlib
import json
import os

import numpy as np
import pytest

import awm_oracle as O
import awm_testlib as T

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
P = O.Params()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def q16(x):
    return O.int16_to_float(O.quantize_sndfile16(x))


def fmt_ref_json(doc):
    out = []
    for m in doc["matches"]:
        out.append({"key": m["key"], "pos": m["pos"], "bits": m["bits"], "quality": "%.5f" % m["quality"], "error": "%.6f" % m["error"],
                    "rating": "%.5f" % m["rating"], "type": m["type"], "speed": "%.6f" % m["speed"]})
    return {"length": doc["length"], "matches": out}


def test_keyed_noise_generator_matches_reference():
    n16 = O.quantize_sndfile16(O.gen_noise(20))
    assert [int(v) for v in n16.reshape(-1)[:32]] == G["gen_noise_20s"]["head"]
    assert sha(n16) == G["gen_noise_20s"]["sha256"]
    assert sha(O.quantize_sndfile16(O.gen_noise(2, key=O.Key.test_key(7)))) == G["gen_noise_2s_testkey7"]["sha256"]


CASES = {
    "clip10_mono": lambda: (q16(T.noise(10.0, 1, seed=1234)), O.Key()),
    "block115": lambda: (q16(T.noise(115.0, 2, seed=1234)), O.Key()),
    "block115_nolimiter": lambda: (q16(T.noise(115.0, 2, seed=1234)), O.Key()),
    "block115_testkey3": lambda: (q16(T.noise(115.0, 2, seed=1234)), O.Key.test_key(3)),
    "limiter30": lambda: (q16(T.noise(30.7, 2, seed=99, amp=1.0)), O.Key()),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_embed_bit_exact_vs_reference(name):
    g = G[name]
    x, key = CASES[name]()
    assert sha(O.quantize_sndfile16(x)) == g["input_sha256"]
    Pc = O.Params(test_no_limiter="--test-no-limiter" in g["add_args"])
    r = O.embed(x, key, g["payload"], Pc)
    out16 = O.quantize_sndfile16(r.samples)
    assert [int(v) for v in out16.reshape(-1)[:64]] == g["output_head"]
    assert sha(out16) == g["output_sha256"]
    assert ("Data Blocks:  %d\n" % r.data_blocks) in g["add_stderr"]
    if "--snr" in g["add_args"]:
        assert ("SNR:          %f dB\n" % r.snr_db) in g["add_stderr"]


@pytest.mark.parametrize("name", ["clip10_mono", "block115_testkey3"])
def test_get_exact_vs_reference(name):
    g = G[name]
    x, key = CASES[name]()
    y = q16(O.embed(x, key, g["payload"], P).samples)
    rs = O.get_watermark(y, [key], P)
    assert rs.json_doc(int(np.rint(len(y) / 44100.0))) == fmt_ref_json(g["json"])
    assert "\n".join(rs.lines()) + "\n" == g["get_stdout"]
    want = O.parse_payload(g["payload"], P)
    assert ("match_count %d %d\n" % (rs.match_count(want), len(rs.patterns))) in g["cmp_stdout"]


def test_sync_after_cut_vs_reference():
    """tests/sync-test.sh: 200 s reference noise, 882300 samples cut off -> 3 matches."""
    g = G["noise200"]
    y16 = O.quantize_sndfile16(T.watermarked_noise(200))        # shared with test_get_48000_vs_reference (cached)
    assert sha(y16) == g["wm_sha256"]
    y = O.int16_to_float(y16)[882300:]
    rs = O.ResultSet()
    O.block_decoder_run(O.Key(), y, rs, P)
    rs.sort([O.Key()])
    assert rs.json_doc(int(np.rint(len(y) / 44100.0))) == fmt_ref_json(g["cut_json"])
    assert "match_count 3 9" in g["cut_cmp_stdout"]
    assert rs.match_count(O.parse_payload("f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", P)) == 3


# ---- speed detection / resampler.  zita-resampler is not part of the reference tree: reference and oracle share the
# in-repo resampler (oracle/ref_shims/awm_vresampler.hh); everything else on this path is reference code.

def test_speed_changed_input_matches_reference():
    g = G["speed30"]
    assert sha(O.quantize_sndfile16(T.watermarked_noise(30))) == g["wm_sha256"]
    for c in g["cases"]:
        y = T.speed_changed(30, c["speed"])
        assert y.shape[0] == c["n_frames"]
        assert sha(O.quantize_sndfile16(y)) == c["input_sha256"], c["speed"]


@pytest.mark.parametrize("idx", [0, 2])
def test_detect_speed_exact_vs_reference(idx):
    """tests/detect-speed-test.sh: speed 0.9764 / 1.01, --detect-speed"""
    c = G["speed30"]["cases"][idx]
    assert c["opt"] == "--detect-speed"
    y = T.speed_changed(30, c["speed"])
    lines = []
    rs = O.get_watermark(y, [O.Key()], P, detect=True, test_speed=T.cli_float(c["speed"]), speed_lines=lines)
    out = "\n".join(lines + rs.lines()) + "\n"
    want = O.parse_payload("f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", P)
    out += "match_count %d %d\n" % (rs.match_count(want), len(rs.patterns))
    assert out == c["cmp_stdout"][:len(out)]
    assert rs.json_doc(int(np.rint(len(y) / 44100.0))) == fmt_ref_json(c["json"])


def test_detect_speed_lines_other_modes_vs_reference():
    """--detect-speed-patient at 1.01: the detect_speed line (speed 1.0, detected but not applied, is covered by the GPU test
    against the same golden)"""
    for idx in (3,):
        c = G["speed30"]["cases"][idx]
        info = O.detect_speed(O.Key(), T.speed_changed(30, c["speed"]), P, patient=c["opt"].endswith("patient"), test_speed=T.cli_float(c["speed"]))
        assert info.line + "\n" == c["cmp_stdout"].split("\n")[0] + "\n"
        assert info.accepted == ("speed " in c["cmp_stdout"].replace("detect_speed", ""))


def test_try_speed_vs_reference():
    g = G["speed30"]["try_speed_1.01"]
    y = T.speed_changed(30, 1.01)
    rs = O.get_watermark(y, [O.Key()], P, try_speed=T.cli_float(1.01))
    assert rs.json_doc(int(np.rint(len(y) / 44100.0))) == fmt_ref_json(g["json"])


def test_get_48000_vs_reference():
    """tests/sample-rate-test.sh (second half): the 44.1 kHz watermark resampled to 48 kHz decodes with 5 matches"""
    g = G["rate48000"]
    y = T.watermarked_noise(200)
    assert sha(O.quantize_sndfile16(y)) == G["noise200"]["wm_sha256"]
    z = O.int16_to_float(O.quantize_sndfile16(O.resample(y, 44100, 48000)))
    assert z.shape[0] == g["n_frames"] and sha(O.quantize_sndfile16(z)) == g["input_sha256"]
    rs = O.get_watermark(z, [O.Key()], P, rate=48000)
    n44 = O.stream_out_count(z.shape[0], 44100 / 48000.0)
    assert rs.json_doc(int(np.rint(n44 / 44100.0))) == fmt_ref_json(g["json"])
    assert "match_count 5 10" in g["cmp_stdout"]


@pytest.mark.parametrize("name", ["rate32000_add", "rate48000_add_nolimiter"])
def test_embed_other_sample_rates_bit_exact_vs_reference(name):
    """tests/sample-rate-test.sh, first half (and a 48 kHz / --test-no-limiter variant): add at a rate that needs the
    WatermarkResampler, output PCM, SNR and Data Blocks lines of the reference"""
    g = G[name]
    x = q16(O.gen_noise(g["seconds"], g["rate"]))
    assert sha(O.quantize_sndfile16(x)) == g["input_sha256"]
    Pc = O.Params(test_no_limiter="--test-no-limiter" in g["add_args"])
    r = O.embed(x, O.Key(), "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", Pc, g["rate"])
    assert sha(O.quantize_sndfile16(r.samples)) == g["output_sha256"]
    assert ("Data Blocks:  %d\n" % r.data_blocks) in g["add_stderr"]
    assert ("SNR:          %f dB\n" % r.snr_db) in g["add_stderr"]


def _opt_params(opts):
    Pc = O.Params()
    if "--linear" in opts:
        Pc.mix = False
    if "--short" in opts:
        Pc.payload_short, Pc.payload_size = True, int(opts[opts.index("--short") + 1])
    return Pc


@pytest.mark.parametrize("name", ["short12", "linear120"])      # short16 / short20: GPU tests against the same goldens + host table test
def test_short_payload_and_linear_exact_vs_reference(name):
    """tests/short-payload-test.sh (block code [56,12] / [65,20] in front of the convolutional code) and --linear"""
    g = G[name]
    Pc = _opt_params(g["opts"])
    x = q16(O.gen_noise(120))
    r = O.embed(x, O.Key(), g["payload"], Pc)
    y16 = O.quantize_sndfile16(r.samples)
    assert sha(y16) == g["output_sha256"]
    assert ("Data Blocks:  %d\n" % r.data_blocks) in g["add_stderr"]
    rs = O.get_watermark(O.int16_to_float(y16), [O.Key()], Pc)
    assert "\n".join(rs.lines()) + "\n" == g["get_stdout"]
    assert rs.json_doc(120) == fmt_ref_json(g["json"])
    assert g["cmp_rc"] == 0


def test_strength_frames_per_bit_hard_exact_vs_reference():
    """--strength 15 --frames-per-bit 3 on add (block length 3084 frames), --frames-per-bit 3 --hard on get"""
    g = G["strength15_fpb3_hard"]
    x = q16(O.gen_noise(130))
    Pa = O.Params(water_delta=float(np.float32(15.0)) / 1000, frames_per_bit=3)      # the CLI parses 15 as float, then / 1000
    r = O.embed(x, O.Key(), g["payload"], Pa)
    y16 = O.quantize_sndfile16(r.samples)
    assert sha(y16) == g["output_sha256"]
    assert ("Data Blocks:  %d\n" % r.data_blocks) in g["add_stderr"]
    rs = O.get_watermark(O.int16_to_float(y16), [O.Key()], O.Params(frames_per_bit=3, hard=True))
    assert "\n".join(rs.lines()) + "\n" == g["get_stdout"]
    assert rs.json_doc(130) == fmt_ref_json(g["json"])


def test_two_named_keys_exact_vs_reference():
    """tests/key-test.sh: a file watermarked twice with two keys (key files with names), `get --key k1 --key k2`"""
    g = G["two_keys30"]
    ka = O.Key(bytes.fromhex(g["keys"]["alpha"]), "alpha")
    kb = O.Key(bytes.fromhex(g["keys"]["beta"]), "beta")
    x = q16(O.gen_noise(30))
    y1 = q16(O.embed(x, ka, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", P).samples)
    y2_16 = O.quantize_sndfile16(O.embed(y1, kb, "0123456789abcdef0123456789abcdef", P).samples)
    assert sha(y2_16) == g["output_sha256"]
    rs = O.get_watermark(O.int16_to_float(y2_16), [ka, kb], P)
    assert "\n".join(rs.lines()) + "\n" == g["get_stdout"]
    assert rs.json_doc(30) == fmt_ref_json(g["json"])


def test_fft_shim_simd_bit_identical():
    """oracle/ref_shims/fftw_shim.cc runs its two narrow passes with explicit AVX2 where the CPU has it (runtime dispatch); the
    results must not depend on that: same IEEE operations per element, no FMA.  Checked through liboracle, which links the shim."""
    import ctypes
    L = O.lib()
    rng = np.random.default_rng(3)
    for n in (1024, 512, 64):
        L.fftwf_plan_dft_r2c_1d.restype = ctypes.c_void_p
        L.fftwf_plan_dft_c2r_1d.restype = ctypes.c_void_p
        fwd = ctypes.c_void_p(L.fftwf_plan_dft_r2c_1d(ctypes.c_int(n), None, None, ctypes.c_uint(0)))
        inv = ctypes.c_void_p(L.fftwf_plan_dft_c2r_1d(ctypes.c_int(n), None, None, ctypes.c_uint(0)))
        for scale in (1.0, 1e-4):
            x = ((rng.random(n + 2, dtype=np.float32) - 0.5) * scale).astype(np.float32)
            res = []
            for simd in (0, 1):
                L.awm_shim_set_simd(ctypes.c_int(simd))
                spec = np.zeros(n + 2, np.float32)
                back = np.zeros(n + 2, np.float32)
                L.fftwf_execute_dft_r2c(fwd, x.ctypes.data_as(ctypes.c_void_p), spec.ctypes.data_as(ctypes.c_void_p))
                L.fftwf_execute_dft_c2r(inv, spec.copy().ctypes.data_as(ctypes.c_void_p), back.ctypes.data_as(ctypes.c_void_p))
                res.append((spec.tobytes(), back.tobytes()))
            assert res[0] == res[1]
            ref = np.fft.rfft(x[:n].astype(np.float64))
            got = np.frombuffer(res[1][0], np.float32).astype(np.float64)
            assert np.abs(got[0::2] + 1j * got[1::2] - ref).max() < 2e-6 * np.abs(ref).max()
        L.fftwf_destroy_plan(fwd)
        L.fftwf_destroy_plan(inv)
    L.awm_shim_set_simd(ctypes.c_int(1))
