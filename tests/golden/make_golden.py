"""Generate tests/golden/golden.json with the UNMODIFIED reference built by oracle/Makefile.ref
(oracle/_ref/audiowmark; its FFT is the in-repo shim, everything else is reference code).

Run here (needs /root/reference to build the binary):   python tests/golden/make_golden.py
The fixtures store hashes / JSON documents only, inputs are regenerated from seeds by the tests.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import awm_oracle as O      # noqa: E402
import awm_testlib as T     # noqa: E402
import build_oracle         # noqa: E402

REF = build_oracle.build_reference()
assert REF and os.path.exists(REF), "reference binary not available"


def run(*args, ok_codes=(0,)):
    p = subprocess.run([REF] + [str(a) for a in args], capture_output=True, text=True)
    assert p.returncode in ok_codes, (args, p.returncode, p.stderr)
    return p


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pcm16(path):
    x, rate, bits = O.read_wav(path)
    return O.quantize_sndfile16(x)      # exact: file values are k / 32768


def case(tmp, name, x, payload, extra_add=(), extra_get=()):
    """x: float32 [n, ch] already on the s16 grid."""
    src, dst, js = (os.path.join(tmp, name + s) for s in (".wav", "_wm.wav", ".json"))
    O.write_wav16(src, x)
    p = run("add", *extra_add, src, dst, payload)
    out16 = pcm16(dst)
    g = run("get", *extra_get, "--json", js, dst)
    c = run("cmp", *extra_get, dst, payload, ok_codes=(0, 1))
    return {
        "input_sha256": sha(O.quantize_sndfile16(x)),
        "add_args": list(extra_add), "get_args": list(extra_get), "payload": payload,
        "add_stderr": p.stderr,
        "output_sha256": sha(out16),
        "output_head": [int(v) for v in out16.reshape(-1)[:64]],
        "get_stdout": g.stdout,
        "cmp_stdout": c.stdout, "cmp_rc": c.returncode,
        "json": json.load(open(js)),
    }, O.int16_to_float(out16)


def main():
    G = {"reference": "swesterfeld/audiowmark 0.6.5 sources compiled unmodified by oracle/Makefile.ref (FFT: oracle/ref_shims/fftw_shim.cc)"}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- reference noise generator (keyed PRNG + 16 bit quantisation), tests/test-common.sh.in
        nz = os.path.join(tmp, "noise.wav")
        run("test-gen-noise", nz, 20, 44100)
        n16 = pcm16(nz)
        G["gen_noise_20s"] = {"sha256": sha(n16), "head": [int(v) for v in n16.reshape(-1)[:32]]}
        run("test-gen-noise", "--test-key", 7, nz, 2, 44100)
        G["gen_noise_2s_testkey7"] = {"sha256": sha(pcm16(nz))}

        q = lambda x: O.int16_to_float(O.quantize_sndfile16(x))
        # ---- config[0]: 10 s mono, clip decoder
        G["clip10_mono"], _ = case(tmp, "clip10", q(T.noise(10.0, 1, seed=1234)), T.PAYLOAD)
        # ---- 115 s stereo: one full A block; limiter on / off; second key
        x115 = q(T.noise(115.0, 2, seed=1234))
        G["block115"], y115 = case(tmp, "block115", x115, T.PAYLOAD)
        G["block115_nolimiter"], _ = case(tmp, "block115nl", x115, T.PAYLOAD, extra_add=("--test-no-limiter", "--snr"))
        G["block115_testkey3"], _ = case(tmp, "block115k", x115, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0",
                                         extra_add=("--test-key", 3), extra_get=("--test-key", 3))
        # full-scale noise drives the limiter hard
        G["limiter30"], _ = case(tmp, "lim30", q(T.noise(30.7, 2, seed=99, amp=1.0)), "0f")
        # ---- sync test (tests/sync-test.sh): 200 s reference noise, cut 882300 samples -> 3 matches
        run("test-gen-noise", nz, 200, 44100)
        wm = os.path.join(tmp, "wm200.wav")
        cut = os.path.join(tmp, "cut200.wav")
        js = os.path.join(tmp, "cut.json")
        run("add", nz, wm, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0")
        c0 = run("cmp", wm, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0")
        run("cut-start", wm, cut, 882300)
        c1 = run("cmp", cut, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", "--test-cut", 882300, "--json", js)
        G["noise200"] = {"wm_sha256": sha(pcm16(wm)), "cmp_stdout": c0.stdout,
                         "cut_cmp_stdout": c1.stdout, "cut_json": json.load(open(js))}
        # ---- speed detection (tests/detect-speed-test.sh): 30 s reference noise, watermarked, speed changed.
        # NOTE: everything behind a resampler runs on oracle/ref_shims/awm_vresampler.hh (zita-resampler is absent).
        run("test-gen-noise", nz, 30, 44100)
        wm30 = os.path.join(tmp, "wm30.wav")
        sp30 = os.path.join(tmp, "sp30.wav")
        run("add", nz, wm30, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0")
        G["speed30"] = {"wm_sha256": sha(pcm16(wm30)), "cases": []}
        for speed, opt in ((0.9764, "--detect-speed"), (1.0, "--detect-speed"), (1.01, "--detect-speed"), (1.01, "--detect-speed-patient")):
            run("test-change-speed", wm30, sp30, speed)
            c = run("cmp", sp30, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", opt, "--test-speed", speed, "--json", js)
            G["speed30"]["cases"].append({"speed": speed, "opt": opt, "input_sha256": sha(pcm16(sp30)), "n_frames": int(pcm16(sp30).shape[0]),
                                          "cmp_stdout": c.stdout, "json": json.load(open(js))})
        t = run("cmp", sp30, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", "--try-speed", 1.01, "--json", js)
        G["speed30"]["try_speed_1.01"] = {"cmp_stdout": t.stdout, "json": json.load(open(js))}
        # ---- sample rate (tests/sample-rate-test.sh, shortened): 44.1 kHz watermark resampled to 48 kHz, get resamples back
        r48 = os.path.join(tmp, "r48.wav")
        run("test-resample", wm, r48, 48000)
        c48 = run("cmp", r48, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", "--json", js)
        G["rate48000"] = {"input_sha256": sha(pcm16(r48)), "n_frames": int(pcm16(r48).shape[0]), "cmp_stdout": c48.stdout, "json": json.load(open(js))}
        # ---- sample rate (tests/sample-rate-test.sh, first half): 200 s reference noise at 32 kHz, add + cmp -> 5 matches;
        # plus a short 48 kHz case without limiter and with --snr
        def rate_case(name, seconds, rate, extra_add):
            src, dst = os.path.join(tmp, name + ".wav"), os.path.join(tmp, name + "_wm.wav")
            run("test-gen-noise", src, seconds, rate)
            p = run("add", *extra_add, src, dst, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0")
            c = run("cmp", dst, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0", "--json", js, ok_codes=(0, 1))
            return {"seconds": seconds, "rate": rate, "add_args": list(extra_add), "input_sha256": sha(pcm16(src)), "output_sha256": sha(pcm16(dst)),
                    "add_stderr": p.stderr, "cmp_stdout": c.stdout, "json": json.load(open(js))}
        G["rate32000_add"] = rate_case("r32", 200, 32000, ("--snr",))
        G["rate48000_add_nolimiter"] = rate_case("r48nl", 11.3, 48000, ("--snr", "--test-no-limiter"))
        # ---- short payload (tests/short-payload-test.sh, 120 s instead of 200 s) and --linear
        def opt_case(name, payload, opts):
            src, dst = os.path.join(tmp, name + ".wav"), os.path.join(tmp, name + "_wm.wav")
            run("test-gen-noise", src, 120, 44100)
            p = run("add", *opts, src, dst, payload)
            g = run("get", *opts, "--json", js, dst)
            c = run("cmp", *opts, dst, payload, ok_codes=(0, 1))
            return {"opts": list(opts), "payload": payload, "output_sha256": sha(pcm16(dst)), "add_stderr": p.stderr, "get_stdout": g.stdout,
                    "cmp_stdout": c.stdout, "cmp_rc": c.returncode, "json": json.load(open(js))}
        G["short12"] = opt_case("s12", "abc", ("--short", 12))
        G["short16"] = opt_case("s16", "abcd", ("--short", 16))
        G["short20"] = opt_case("s20", "abcde", ("--short", 20))
        G["linear120"] = opt_case("lin", T.PAYLOAD, ("--linear",))
        # --strength (add only) / --frames-per-bit / --hard (get only)
        def opt_case2(name, payload, add_opts, get_opts):
            src, dst = os.path.join(tmp, name + ".wav"), os.path.join(tmp, name + "_wm.wav")
            run("test-gen-noise", src, 130, 44100)
            p = run("add", *add_opts, src, dst, payload)
            g = run("get", *get_opts, "--json", js, dst)
            return {"add_opts": list(add_opts), "get_opts": list(get_opts), "payload": payload, "output_sha256": sha(pcm16(dst)),
                    "add_stderr": p.stderr, "get_stdout": g.stdout, "json": json.load(open(js))}
        G["strength15_fpb3_hard"] = opt_case2("fpb3", T.PAYLOAD, ("--strength", 15, "--frames-per-bit", 3), ("--frames-per-bit", 3, "--hard"))
        # ---- two keys (tests/key-test.sh): 30 s noise watermarked twice with named keys from key files; get with both keys
        k1, k2 = os.path.join(tmp, "k1.key"), os.path.join(tmp, "k2.key")
        open(k1, "w").write('# watermarking key for audiowmark\n\nkey 000102030405060708090a0b0c0d0e0f\nname "alpha"\n')
        open(k2, "w").write('key 101112131415161718191a1b1c1d1e1f\nname "beta"\n')
        o1, o2 = os.path.join(tmp, "ko1.wav"), os.path.join(tmp, "ko2.wav")
        run("test-gen-noise", nz, 30, 44100)
        run("add", "--key", k1, nz, o1, "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0")
        run("add", "--key", k2, o1, o2, "0123456789abcdef0123456789abcdef")
        g2 = run("get", "--key", k1, "--key", k2, "--json", js, o2)
        G["two_keys30"] = {"keys": {"alpha": "000102030405060708090a0b0c0d0e0f", "beta": "101112131415161718191a1b1c1d1e1f"},
                           "output_sha256": sha(pcm16(o2)), "get_stdout": g2.stdout, "json": json.load(open(js))}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    json.dump(G, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
